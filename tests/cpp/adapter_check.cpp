// Drives KobayashiSimulation (the DXViewer ISimulation adapter) the way the viewer does — DX12App::setSimulation
// (iSetDXApp + iCreateObject), WM_CREATE, then a script of frames and control-panel messages — and prints, after every
// "dump", the state the panel shows and the colours of all objects.  tests/test_driver.py runs the SAME script through the
// unmodified reference class (oracle/_ref, ref_gui_* entry points) and compares.
// Compiled against the portable stand-in for the viewer headers (oracle/ref_harness/Win32App.h).
//
//   adapter_check N KERNEL  op op ...      ops: f = one viewer frame (DX12App::update + draw)
//                                               cK = WM_COMMAND K (9 Reset, 10 Play, 11 Stop, 12 Next step)
//                                               sI,C,P = WM_HSCROLL slider I, request C (SB_*), thumb position P
//                                               n = seed an extra off-centre nucleus (asymmetric picture)
//                                               d = dump
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "Win32App.h"
#include "KobayashiSimulation.hpp"

int main(int argc, char** argv) {
    const int n = argc > 1 ? std::atoi(argv[1]) : 48;
    const int kernel = argc > 2 && std::strcmp(argv[2], "strict") == 0 ? KOB_KERNEL_STRICT : KOB_KERNEL_FAST;
    try {
        KobayashiSimulation* ks = new KobayashiSimulation(n, n, 0.0001f, 0, kernel);
        ISimulation* sim = ks;
        DX12App app;
        app.setSimulation(sim);
        sim->iWMCreate(nullptr, nullptr);
        if ((int)app._constantBuffer.size() != n * n || sim->iGetVertexBufferSize() != 4 || sim->iGetIndexBufferSize() != 6) return 3;
        for (int a = 3; a < argc; ++a) {
            const char* op = argv[a];
            if (op[0] == 'f') { app.update(); app.draw(); }
            else if (op[0] == 'c') sim->iWMCommand(nullptr, 0, (WPARAM)std::atoi(op + 1), 0, nullptr);
            else if (op[0] == 's') {
                int idx = 0, code = 0, pos = 0;
                std::sscanf(op + 1, "%d,%d,%d", &idx, &code, &pos);
                sim->iWMHScroll(nullptr, (WPARAM)((unsigned)code | ((unsigned)pos << 16)), ks->sliderHandle(idx), nullptr);
            } else if (op[0] == 'n') ks->sim().createNucleus(n / 4, n / 2 + 5);
            else if (op[0] == 'd') {
                std::printf("state %d %lld", sim->iIsUpdated() ? 1 : 0, (long long)ks->simFrame());
                for (int i = 0; i < 9; ++i) std::printf(" %.9g", (double)ks->sliderValue(i));
                for (int i = 0; i < 9; ++i) std::printf(" %d", ks->sliderPosition(i));
                std::printf("\n");
                for (int i = 0; i < (int)app._constantBuffer.size(); ++i)
                    std::printf("%.6f %.6f %.6f\n", app._constantBuffer[i].color.x, app._constantBuffer[i].color.y, app._constantBuffer[i].color.z);
            }
        }
        sim->iWMDestory(nullptr);
        delete sim;
    } catch (const KobayashiError& e) {
        std::fprintf(stderr, "adapter_check: %s\n", e.what());
        return 1;
    }
    return 0;
}
