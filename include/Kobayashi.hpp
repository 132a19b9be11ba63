// Kobayashi.hpp — C++ host class with the shape of the reference's `class Kobayashi` (src/Kobayashi.h:32-122),
// implemented entirely on top of the C ABI in kobayashi_c.h (libkobayashi_cuda.so).  Header only.
//
// Reference member                          ->  here
//   Kobayashi(int x, int y, float timeStep)     Kobayashi(x, y, timeStep [, precision, kernel, device, seed])
//   iUpdate()            (:227-239)             iUpdate()                 10 sub-steps + _simTime/_simFrame
//   iResetSimulationState(cb) (:241-249)        iResetSimulationState()   _vectorInit, parameters kept
//   _parameterInit()     (:73-96)               parameterInit()           defaults of :76-84
//   _createNucleus(x,y)  (:116-123)             createNucleus(x, y)       periodic instead of UB at the border
//   float members _tau ... _tEq (:94-105)       tau(), setTau(v) ...      setters do NOT reset (the GUI does, :616)
//   _phi / _t / _angl    (:107-115)             phi(), t(), angl()        host copies, index i + nx*j (:91)
//   iIsUpdated()/_updateFlag (:282-285)         isUpdated(), setUpdated()
// The DirectX/Win32 part of ISimulation (mesh, constant buffers, WndProc hooks) is not reproduced; a viewer reads
// phi() (what iUpdateConstantBuffer reads at :315) or renderRGBA() (the colour ramp of :318-342).
#ifndef KOBAYASHI_HPP
#define KOBAYASHI_HPP

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "kobayashi_c.h"

class KobayashiError : public std::runtime_error {
public:
    KobayashiError(int status, const std::string& what) : std::runtime_error(what), status_(status) {}
    int status() const { return status_; }
private:
    int status_;
};

class Kobayashi {
public:
    Kobayashi(int x, int y, float timeStep, int precision = KOB_F32, int kernel = -1, int device = 0,
              uint64_t seed = 0, int64_t ny_global = 0, int64_t y0 = 0)
        : nx_(x), ny_(y), precision_(precision), seed_(seed), ny_global_(ny_global ? ny_global : y), y0_(y0) {
        kob_params p;
        kob_default_params(&p, static_cast<double>(timeStep));   // float widened, as the reference stores _dt = timeStep
        kob_config c;
        kob_default_config(&c);
        c.precision = precision;
        c.kernel = kernel >= 0 ? kernel : (precision == KOB_F32 ? KOB_KERNEL_FAST : KOB_KERNEL_STRICT);
        c.device = device; c.seed = seed; c.ny_global = ny_global; c.y0 = y0;
        const int st = kob_create(&ctx_, x, y, &p, &c);
        if (st != KOB_OK) throw KobayashiError(st, std::string("kob_create: ") + kob_last_error(nullptr));
    }
    ~Kobayashi() { if (ctx_) kob_destroy(ctx_); }
    Kobayashi(const Kobayashi&) = delete;
    Kobayashi& operator=(const Kobayashi&) = delete;

    // ---- ISimulation: simulation methods ----
    void iUpdate() { if (updateFlag_) ck(kob_update(ctx_), "kob_update"); }
    void iResetSimulationState() { ck(kob_reset(ctx_), "kob_reset"); }
    bool iIsUpdated() const { return updateFlag_; }
    void setUpdated(bool f) { updateFlag_ = f; }          // Play / Stop buttons (src/Kobayashi.cpp:531-543)
    void nextStep() { ck(kob_update(ctx_), "kob_update"); }   // "Next step" button (:544-549)

    // ---- hot path, finer grained ----
    void step(int64_t n = 1) { ck(kob_step(ctx_, n), "kob_step"); }
    float stepTimed(int64_t n) { float ms = 0.f; ck(kob_step_timed(ctx_, n, &ms), "kob_step_timed"); return ms; }
    void sync() { ck(kob_sync(ctx_), "kob_sync"); }
    void clear() { ck(kob_clear(ctx_), "kob_clear"); }
    void createNucleus(int64_t x, int64_t y) { ck(kob_add_nucleus(ctx_, x, y), "kob_add_nucleus"); }

    // ---- parameters (reference members _tau ... _tEq; writes do not reset the fields by themselves) ----
    void parameterInit() { kob_params p = params(); const double dt = p.dt; kob_default_params(&p, dt); setParams(p); }
    kob_params params() const { kob_params p; ck(kob_get_params(ctx_, &p), "kob_get_params"); return p; }
    void setParams(const kob_params& p) { ck(kob_set_params(ctx_, &p), "kob_set_params"); }
#define KOB_PARAM(name, Setter, field)                                                       \
    double name() const { return params().field; }                                          \
    void Setter(double v, bool reset = false) { kob_params p = params(); p.field = v; setParams(p); if (reset) iResetSimulationState(); }
    KOB_PARAM(dx, setDx, dx) KOB_PARAM(dy, setDy, dy) KOB_PARAM(dt, setDt, dt) KOB_PARAM(tau, setTau, tau)
    KOB_PARAM(epsilonBar, setEpsilonBar, epsilon_bar) KOB_PARAM(mu, setMu, mu) KOB_PARAM(K, setK, K)
    KOB_PARAM(delta, setDelta, delta) KOB_PARAM(anisotropy, setAnisotropy, anisotropy) KOB_PARAM(alpha, setAlpha, alpha)
    KOB_PARAM(gamma, setGamma, gamma) KOB_PARAM(tEq, setTEq, t_eq) KOB_PARAM(theta0, setTheta0, theta0)
    KOB_PARAM(noiseAmplitude, setNoiseAmplitude, noise_a)
#undef KOB_PARAM

    // ---- field accessors: host copies in the reference layout i + nx*j ----
    template <typename real> std::vector<real> phi() { return field<real>(0); }
    template <typename real> std::vector<real> t() { return field<real>(1); }
    template <typename real> std::vector<real> angl() { return field<real>(2); }
    void getFields(void* phi, void* t, void* angl) { ck(kob_get_fields(ctx_, phi, t, angl), "kob_get_fields"); }
    void setFields(const void* phi, const void* t, const void* angl) { ck(kob_set_fields(ctx_, phi, t, angl), "kob_set_fields"); }
    void getFieldsAsync(void* phi, void* t, void* angl) { ck(kob_get_fields_async(ctx_, phi, t, angl), "kob_get_fields_async"); }
    void waitFields() { ck(kob_wait_fields(ctx_), "kob_wait_fields"); }
    void getWindow(int64_t x0, int64_t y0, int64_t w, int64_t h, void* phi, void* t, void* angl) { ck(kob_get_window(ctx_, x0, y0, w, h, phi, t, angl), "kob_get_window"); }
    void setWindow(int64_t x0, int64_t y0, int64_t w, int64_t h, const void* phi, const void* t, const void* angl) { ck(kob_set_window(ctx_, x0, y0, w, h, phi, t, angl), "kob_set_window"); }
    void setNoiseField(const float* r) { ck(kob_set_noise_field(ctx_, r), "kob_set_noise_field"); }
    std::vector<uint8_t> renderRGBA() {
        std::vector<uint8_t> img(4u * static_cast<size_t>(nx_) * static_cast<size_t>(ny_));
        ck(kob_render_rgba(ctx_, img.data()), "kob_render_rgba");
        return img;
    }

    // ---- checkpoint / resume (none in the reference).  File layout, little endian:
    //   0  char[8]  "KOBCKPT1"          8  u32 header bytes (256), u32 element bytes (4 | 8)
    //  16  i64 nx, ny, ny_global, y0   48  u64 step counter, u64 Philox seed, i64 sim frame
    //  72  f64[14] kob_params in declaration order, zero padding to 256, then phi, T, theta (ny*nx each, i + nx*j).
    // Resuming reproduces the uninterrupted run bit for bit (fields incl. theta, parameters, Philox step counter). ----
    void saveCheckpoint(const std::string& path) {
        const size_t e = precision_ == KOB_F64 ? 8u : 4u, n = static_cast<size_t>(nx_) * static_cast<size_t>(ny_);
        std::vector<unsigned char> buf(3 * n * e);
        getFields(buf.data(), buf.data() + n * e, buf.data() + 2 * n * e);
        unsigned char h[256] = {0};
        std::memcpy(h, "KOBCKPT1", 8);
        const uint32_t hb = 256, eb = static_cast<uint32_t>(e);
        std::memcpy(h + 8, &hb, 4); std::memcpy(h + 12, &eb, 4);
        const int64_t dims[4] = {nx_, ny_, ny_global_, y0_};
        std::memcpy(h + 16, dims, 32);
        const uint64_t sc = stepCounter();
        const int64_t fr = simFrame();
        std::memcpy(h + 48, &sc, 8); std::memcpy(h + 56, &seed_, 8); std::memcpy(h + 64, &fr, 8);
        const kob_params p = params();
        static_assert(sizeof(kob_params) == 14 * sizeof(double), "kob_params layout");
        std::memcpy(h + 72, &p, sizeof p);
        FILE* fp = std::fopen(path.c_str(), "wb");
        if (!fp) throw KobayashiError(KOB_ERR_INVALID_ARG, "cannot open " + path + " for writing");
        const bool ok = std::fwrite(h, 1, 256, fp) == 256 && std::fwrite(buf.data(), 1, buf.size(), fp) == buf.size();
        if (std::fclose(fp) != 0 || !ok) throw KobayashiError(KOB_ERR_INVALID_ARG, "short write to " + path);
    }
    void loadCheckpoint(const std::string& path) {
        FILE* fp = std::fopen(path.c_str(), "rb");
        if (!fp) throw KobayashiError(KOB_ERR_INVALID_ARG, "cannot open " + path);
        unsigned char h[256];
        const size_t e = precision_ == KOB_F64 ? 8u : 4u, n = static_cast<size_t>(nx_) * static_cast<size_t>(ny_);
        std::vector<unsigned char> buf(3 * n * e);
        const bool ok = std::fread(h, 1, 256, fp) == 256 && std::fread(buf.data(), 1, buf.size(), fp) == buf.size();
        std::fclose(fp);
        uint32_t hb, eb; int64_t dims[4], fr; uint64_t sc, seed; kob_params p;
        std::memcpy(&hb, h + 8, 4); std::memcpy(&eb, h + 12, 4); std::memcpy(dims, h + 16, 32);
        std::memcpy(&sc, h + 48, 8); std::memcpy(&seed, h + 56, 8); std::memcpy(&fr, h + 64, 8); std::memcpy(&p, h + 72, sizeof p);
        if (!ok || std::memcmp(h, "KOBCKPT1", 8) != 0 || hb != 256) throw KobayashiError(KOB_ERR_INVALID_ARG, path + " is not a KOBCKPT1 checkpoint");
        if (eb != e || dims[0] != nx_ || dims[1] != ny_ || dims[2] != ny_global_ || dims[3] != y0_ || seed != seed_)
            throw KobayashiError(KOB_ERR_INVALID_ARG, "checkpoint " + path + " was written for another grid / precision / strip / seed");
        setParams(p);
        setFields(buf.data(), buf.data() + n * e, buf.data() + 2 * n * e);
        sync();                                                    // kob_set_fields is asynchronous: buf must outlive the copies
        ck(kob_set_step_counter(ctx_, sc), "kob_set_step_counter");
        setSimCounters(fr < 0 ? 0 : fr, simTimeMs());              // _simFrame continues; on a linked strip follow with kob_halo_refresh on every strip
    }

    // ---- bookkeeping (_simTime, _simFrame, src/Kobayashi.cpp:237-238) ----
    int64_t simFrame() const { int64_t f = 0; ck(kob_sim_frame(ctx_, &f), "kob_sim_frame"); return f; }
    void setSimCounters(int64_t frames, double ms) { ck(kob_set_sim_counters(ctx_, frames, ms), "kob_set_sim_counters"); }
    double simTimeMs() const { double ms = 0; ck(kob_sim_time_ms(ctx_, &ms), "kob_sim_time_ms"); return ms; }
    uint64_t launchCount() const { uint64_t n = 0; ck(kob_launch_count(ctx_, &n), "kob_launch_count"); return n; }
    struct PathStats { uint64_t singleSteps, pairedSteps; double denseFraction; bool singleMode; uint64_t concurrentPairs; };
    PathStats pathStats() const {          // which step path ran (single-step kernel / two-step launch pairs), last density probe
        PathStats s{0, 0, 0.0, false, 0};
        int32_t m = 0;
        ck(kob_path_stats(ctx_, &s.singleSteps, &s.pairedSteps, &s.denseFraction, &m), "kob_path_stats");
        ck(kob_concurrent_pairs(ctx_, &s.concurrentPairs), "kob_concurrent_pairs");   // pairs whose general pass ran beside the far pass
        s.singleMode = m != 0;
        return s;
    }
    uint64_t stepCounter() const { uint64_t s = 0; ck(kob_get_step_counter(ctx_, &s), "kob_get_step_counter"); return s; }
    int nx() const { return nx_; }
    int ny() const { return ny_; }
    int precision() const { return precision_; }
    kob_ctx* handle() { return ctx_; }

private:
    void ck(int st, const char* what) const {
        if (st != KOB_OK) throw KobayashiError(st, std::string(what) + ": " + kob_last_error(ctx_) + " (" + kob_strerror(st) + ")");
    }
    template <typename real> std::vector<real> field(int which) {
        if (sizeof(real) != (precision_ == KOB_F64 ? 8u : 4u)) throw KobayashiError(KOB_ERR_INVALID_ARG, "element type does not match the context precision");
        std::vector<real> v(static_cast<size_t>(nx_) * static_cast<size_t>(ny_));
        ck(kob_get_fields(ctx_, which == 0 ? v.data() : nullptr, which == 1 ? v.data() : nullptr, which == 2 ? v.data() : nullptr), "kob_get_fields");
        return v;
    }
    kob_ctx* ctx_ = nullptr;
    int nx_, ny_, precision_;
    uint64_t seed_;
    int64_t ny_global_, y0_;
    bool updateFlag_ = true;
};

#endif  // KOBAYASHI_HPP
