// Drives KobayashiSimulation (the DXViewer ISimulation adapter) the way DX12App does — iCreateObject once, then per
// frame iUpdate + iUpdateConstantBuffer for every object — and prints the colours of all objects after FRAMES frames.
// Compiled by tests/test_driver.py against the portable stand-in for the viewer headers (oracle/ref_harness/Win32App.h).
#include <cstdio>
#include <cstdlib>

#include "Win32App.h"
#include "KobayashiSimulation.hpp"

int main(int argc, char** argv) {
    const int n = argc > 1 ? std::atoi(argv[1]) : 48, frames = argc > 2 ? std::atoi(argv[2]) : 3;
    try {
        KobayashiSimulation* ks = new KobayashiSimulation(n, n, 0.0001f);
        ks->sim().createNucleus(n / 4, n / 2 + 5);          // off-centre: makes the picture asymmetric under transposition
        ISimulation* sim = ks;
        std::vector<ConstantBuffer> cb;
        sim->iCreateObject(cb);
        if ((int)cb.size() != n * n || sim->iGetVertexBufferSize() != 4 || sim->iGetIndexBufferSize() != 6) return 3;
        for (int f = 0; f < frames; ++f) {
            if (sim->iIsUpdated()) sim->iUpdate();
            for (int i = 0; i < (int)cb.size(); ++i) sim->iUpdateConstantBuffer(cb, i);
        }
        for (int i = 0; i < (int)cb.size(); ++i)
            std::printf("%d %d %d\n", (int)(cb[i].color.x * 255.0f + 0.5f), (int)(cb[i].color.y * 255.0f + 0.5f), (int)(cb[i].color.z * 255.0f + 0.5f));
        sim->iResetSimulationState(cb);
        delete sim;
    } catch (const KobayashiError& e) {
        std::fprintf(stderr, "adapter_check: %s\n", e.what());
        return 1;
    }
    return 0;
}
