// kob_bench — headless C++ driver: replaces the DXViewer render loop (ext/DXViewer/src/Win32App.cpp:170-193 ->
// DX12App::update -> ISimulation::iUpdate) for benchmarking and long runs.  It constructs the Kobayashi host class
// exactly like src/main.cpp:14-18 does (grid, dt), calls iUpdate() in a loop, and prints one JSON line.
//
//   kob_bench [--nx 250] [--ny 250] [--dt 1e-4] [--frames 200] [--kernel fast|strict] [--precision f32|f64]
//             [--anisotropy 6] [--noise 0] [--seed 0] [--nuclei N] [--ppm out.ppm] [--device 0]
//             [--resume in.kobck] [--save out.kobck] [--checkpoint-every FRAMES]   (long runs: SURVEY §8f rank 4)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "Kobayashi.hpp"

static uint32_t philox_word(uint32_t k, uint64_t seed, int which) {   // nucleus placement, same rule as strips.py
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    uint32_t c[4] = {k, 0, 0, 0}, key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ key[0], n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ key[1], n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3; key[0] += W0; key[1] += W1;
    }
    return c[which];
}

int main(int argc, char** argv) {
    int nx = 250, ny = 250, frames = 200, device = 0, nuclei = 0;
    double dt = 1e-4, aniso = 6.0, noise = 0.0;
    uint64_t seed = 0;
    std::string kernel = "fast", precision = "f32", ppm, resume, save;
    int ckpt_every = 0;
    for (int i = 1; i < argc; ++i) {
        auto next = [&]() -> const char* { if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", argv[i]); std::exit(2); } return argv[++i]; };
        if (!std::strcmp(argv[i], "--nx")) nx = std::atoi(next());
        else if (!std::strcmp(argv[i], "--ny")) ny = std::atoi(next());
        else if (!std::strcmp(argv[i], "--dt")) dt = std::atof(next());
        else if (!std::strcmp(argv[i], "--frames")) frames = std::atoi(next());
        else if (!std::strcmp(argv[i], "--kernel")) kernel = next();
        else if (!std::strcmp(argv[i], "--precision")) precision = next();
        else if (!std::strcmp(argv[i], "--anisotropy")) aniso = std::atof(next());
        else if (!std::strcmp(argv[i], "--noise")) noise = std::atof(next());
        else if (!std::strcmp(argv[i], "--seed")) seed = std::strtoull(next(), nullptr, 10);
        else if (!std::strcmp(argv[i], "--nuclei")) nuclei = std::atoi(next());
        else if (!std::strcmp(argv[i], "--ppm")) ppm = next();
        else if (!std::strcmp(argv[i], "--device")) device = std::atoi(next());
        else if (!std::strcmp(argv[i], "--resume")) resume = next();
        else if (!std::strcmp(argv[i], "--save")) save = next();
        else if (!std::strcmp(argv[i], "--checkpoint-every")) ckpt_every = std::atoi(next());
        else { std::fprintf(stderr, "unknown flag %s\n", argv[i]); return 2; }
    }
    try {
        const int prec = precision == "f64" ? KOB_F64 : KOB_F32;
        const int kern = kernel == "strict" || prec == KOB_F64 ? KOB_KERNEL_STRICT : KOB_KERNEL_FAST;
        Kobayashi sim(nx, ny, (float)dt, prec, kern, device, seed);   // src/main.cpp:14-18
        sim.setAnisotropy(aniso);
        sim.setNoiseAmplitude(noise);
        if (nuclei > 0) {
            sim.clear();
            for (int k = 0; k < nuclei; ++k)
                sim.createNucleus(8 + philox_word(k, seed, 0) % (nx - 16), 8 + philox_word(k, seed, 1) % (ny - 16));
        }
        if (!resume.empty()) sim.loadCheckpoint(resume);              // continue a long run exactly where it stopped
        else sim.iUpdate();                                           // warm-up frame
        sim.sync();
        const auto t0 = std::chrono::steady_clock::now();
        for (int f = 0; f < frames; ++f) {
            sim.iUpdate();                                            // the render loop's only simulation call
            if (ckpt_every > 0 && !save.empty() && (f + 1) % ckpt_every == 0) sim.saveCheckpoint(save);
        }
        sim.sync();
        if (!save.empty()) sim.saveCheckpoint(save);
        const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        const double cells = (double)nx * ny * 10.0 * frames;
        double solid = 0;
        if (prec == KOB_F32) { for (float v : sim.phi<float>()) solid += v > 0.5f; }
        else { for (double v : sim.phi<double>()) solid += v > 0.5; }
        if (!ppm.empty()) {
            const auto img = sim.renderRGBA();
            FILE* fp = std::fopen(ppm.c_str(), "wb");
            if (fp) {
                std::fprintf(fp, "P6\n%d %d\n255\n", nx, ny);
                for (size_t i = 0; i < (size_t)nx * ny; ++i) std::fwrite(&img[4 * i], 1, 3, fp);
                std::fclose(fp);
            }
        }
        std::printf("{\"driver\": \"kob_bench\", \"nx\": %d, \"ny\": %d, \"frames\": %d, \"substeps\": %d, \"kernel\": \"%s\", "
                    "\"precision\": \"%s\", \"sim_time_ms\": %.3f, \"sim_frames\": %lld, \"wall_s\": %.4f, \"gcell_per_s_device\": %.3f, "
                    "\"gcell_per_s_wall\": %.3f, \"solid_cells\": %.0f, \"launches\": %llu}\n",
                    nx, ny, frames, frames * 10, kernel.c_str(), precision.c_str(), sim.simTimeMs(), (long long)sim.simFrame(), wall,
                    (double)nx * ny * 10.0 * (double)sim.simFrame() / (sim.simTimeMs() * 1e-3) / 1e9,   // all iUpdate calls incl. warm-up
                    cells / wall / 1e9, solid, (unsigned long long)sim.launchCount());
    } catch (const KobayashiError& e) {
        std::fprintf(stderr, "kob_bench: %s\n", e.what());
        return 1;
    }
    return 0;
}
