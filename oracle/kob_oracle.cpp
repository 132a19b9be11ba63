// kob_oracle.cpp — CPU ORACLE for the Kobayashi step.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A from-scratch restatement of the reference's hot path, kept in the reference's own two-pass,
// scratch-array structure so that it can be proven BITWISE equal to the reference translation unit
// compiled in place (oracle/Makefile -> oracle/_ref/, test tests/test_oracle_vs_ref.py, gate G0):
//   state + init      src/Kobayashi.h:91,107-115 ; src/Kobayashi.cpp:98-123
//   pass 1            src/Kobayashi.cpp:125-175  (_computeGradientLaplacian)
//   pass 2            src/Kobayashi.cpp:177-221  (_evolution)
//   step driver       src/Kobayashi.cpp:227-239  (iUpdate: 10 sub-steps)
//   parameters        src/Kobayashi.cpp:61-63,76-84 ; PI_F ext/DXViewer/DXViewer-3.1.0/include/dx12header.h:22
// Parity status: PINNED — against the reference TU itself (bitwise, FP32 and FP64-typed) and against the
// known-answer values in SURVEY.md §8c (tests/golden/).  The reference has no tests of its own.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
// The product (libkobayashi_cuda.so) never links or calls it.
//
// Extensions over the reference (all bit-neutral when zero/off): theta0, Philox / injected noise,
// row strips with explicit ghost rows (used by the world_size-2 gloo tests), a portable math provider
// (crystalgrowth_b200/csrc/kob_math.h) that the CUDA STRICT kernel reproduces bit-for-bit, OpenMP.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <new>

#include "../include/kobayashi_c.h"               // kob_params (POD only)
#include "../crystalgrowth_b200/csrc/kob_math.h"  // portable atan/sin/cos, Philox, REF_PI_F, REF_DEADBAND

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

enum { MATH_LIBM = 0, MATH_PORTABLE = 1 };
constexpr int G = 2;  // ghost rows per side

template <typename real, int MATH>
struct M;
template <typename real>
struct M<real, MATH_LIBM> {
    // `atan(x)` with float x under `using namespace std` resolves to atanf (src/Kobayashi.cpp:162).
    static real atan_(real x) { return std::atan(x); }
    static real sin_(real x) { return std::sin(x); }
    static real cos_(real x) { return std::cos(x); }
};
template <typename real>
struct M<real, MATH_PORTABLE> {
    static real atan_(real x) { return kob::p_atan(x); }
    static real sin_(real x) { return kob::p_sin(x); }
    static real cos_(real x) { return kob::p_cos(x); }
};

struct OracleBase {
    virtual ~OracleBase() {}
    virtual void clear() = 0;
    virtual void add_nucleus(int64_t x, int64_t y) = 0;
    virtual void step(int64_t n) = 0;
    virtual void get(void* phi, void* t, void* angl) const = 0;
    virtual void set(const void* phi, const void* t, const void* angl) = 0;
    virtual void get_edge(int side, void* phi2, void* t2, void* angl2) const = 0;
    virtual void set_ghost(int side, const void* phi2, const void* t2, const void* angl2) = 0;
    int prec = 32, math = MATH_LIBM, threads = 1;
    int64_t nx = 0, ny = 0, ny_global = 0, y0 = 0;
    int64_t noise_x0 = 0, noise_y0 = 0;   // global cell of local (0, 0) for the Philox keys: a window of a larger torus
    kob_params p{};
    uint64_t seed = 0, step_counter = 0;
    bool self_periodic = true;
    std::vector<float> noise_field;  // injected r, reference layout, empty = Philox
};

template <typename real>
struct Oracle : OracleBase {
    // (ny + 2G) rows of nx; local row j lives at padded row j + G.
    std::vector<real> phi, t, angl, eps, epsd, gx, gy, lapphi, lapt;
    // parameters rounded ONCE to `real`, like the reference's float members
    real dx, dy, dt, tau, epsbar, K, delta, aniso, alpha, gamma_, teq, theta0, noise_a;

    size_t idx(int64_t i, int64_t jpad) const { return (size_t)i + (size_t)nx * (size_t)jpad; }

    void init(int64_t nx_, int64_t ny_, int64_t nyg, int64_t y0_, const kob_params& pp) {
        nx = nx_; ny = ny_; ny_global = nyg ? nyg : ny_; y0 = y0_;
        self_periodic = (ny_global == ny);
        set_params(pp);
        const size_t n = (size_t)nx * (size_t)(ny + 2 * G);
        phi.assign(n, 0); t.assign(n, 0); angl.assign(n, 0); eps.assign(n, 0); epsd.assign(n, 0);
        gx.assign(n, 0); gy.assign(n, 0); lapphi.assign(n, 0); lapt.assign(n, 0);
    }
    void set_params(const kob_params& pp) {
        p = pp;
        dx = (real)pp.dx; dy = (real)pp.dy; dt = (real)pp.dt; tau = (real)pp.tau;
        epsbar = (real)pp.epsilon_bar; K = (real)pp.K; delta = (real)pp.delta; aniso = (real)pp.anisotropy;
        alpha = (real)pp.alpha; gamma_ = (real)pp.gamma; teq = (real)pp.t_eq;
        theta0 = (real)pp.theta0; noise_a = (real)pp.noise_a;
    }
    // _vectorInit without the nucleus (src/Kobayashi.cpp:100-109)
    void clear() override {
        std::fill(phi.begin(), phi.end(), (real)0); std::fill(t.begin(), t.end(), (real)0);
        std::fill(angl.begin(), angl.end(), (real)0); std::fill(eps.begin(), eps.end(), (real)0);
        std::fill(epsd.begin(), epsd.end(), (real)0); std::fill(gx.begin(), gx.end(), (real)0);
        std::fill(gy.begin(), gy.end(), (real)0); std::fill(lapphi.begin(), lapphi.end(), (real)0);
        std::fill(lapt.begin(), lapt.end(), (real)0);
        step_counter = 0;
    }
    // _createNucleus (src/Kobayashi.cpp:116-123), global coordinates, periodic wrap instead of UB.
    void put(int64_t x, int64_t y) {
        x = ((x % nx) + nx) % nx;
        y = ((y % ny_global) + ny_global) % ny_global;
        const int64_t jl = y - y0;
        if (jl >= 0 && jl < ny) phi[idx(x, jl + G)] = (real)1;
    }
    void add_nucleus(int64_t x, int64_t y) override {
        put(x, y); put(x - 1, y); put(x + 1, y); put(x, y - 1); put(x, y + 1);
    }
    void wrap_ghosts() {
        for (int g = 0; g < G; ++g) {
            const int64_t lo_src = (((-(int64_t)G + g) % ny) + ny) % ny;  // local rows -2, -1
            const int64_t hi_src = (ny + g) % ny;                          // local rows ny, ny+1
            for (std::vector<real>* a : {&phi, &t, &angl}) {
                std::memcpy(&(*a)[idx(0, g)], &(*a)[idx(0, lo_src + G)], sizeof(real) * (size_t)nx);
                std::memcpy(&(*a)[idx(0, ny + G + g)], &(*a)[idx(0, hi_src + G)], sizeof(real) * (size_t)nx);
            }
        }
    }

    template <int MATH>
    void pass1_row(int64_t jp) {  // jp = padded row, G-1 <= jp <= ny+G
        const real e = (real)kob::REF_DEADBAND;
        const real pi = (real)kob::REF_PI_F;
        const real lapden = (real)3.0f * dx * dx;  // (3.0f * _dx) * _dx  (src/Kobayashi.cpp:146)
        for (int64_t i = 0; i < nx; ++i) {
            const int64_t ip = (i + 1) % nx, im = ((i - 1) + nx) % nx;  // src/Kobayashi.cpp:133-134
            const int64_t jpp = jp + 1, jm = jp - 1;                    // rows come from ghosts, :135-136
            const size_t c = idx(i, jp);
            const real gxv = (phi[idx(ip, jp)] - phi[idx(im, jp)]) / dx;   // :139 (no factor 1/2: reference quirk)
            const real gyv = (phi[idx(i, jpp)] - phi[idx(i, jm)]) / dy;    // :140
            gx[c] = gxv; gy[c] = gyv;
            lapphi[c] = ((real)2.0f * (phi[idx(ip, jp)] + phi[idx(im, jp)] + phi[idx(i, jpp)] + phi[idx(i, jm)])
                         + phi[idx(ip, jpp)] + phi[idx(im, jm)] + phi[idx(im, jpp)] + phi[idx(ip, jm)]
                         - (real)12.0f * phi[c]) / lapden;                  // :142-146
            lapt[c] = ((real)2.0f * (t[idx(ip, jp)] + t[idx(im, jp)] + t[idx(i, jpp)] + t[idx(i, jm)])
                       + t[idx(ip, jpp)] + t[idx(im, jm)] + t[idx(im, jpp)] + t[idx(ip, jm)]
                       - (real)12.0f * t[c]) / lapden;                      // :147-151
            // angle state machine, src/Kobayashi.cpp:154-167.  theta keeps its old value on the "else" paths.
            real th = angl[c];
            if (gxv <= e && gxv >= -e) {
                if (gyv < -e) th = (real)-0.5f * pi;
                else if (gyv > e) th = (real)0.5f * pi;
            }
            if (gxv > e) {
                if (gyv < -e) th = (real)2.0f * pi + M<real, MATH>::atan_(gyv / gxv);
                else if (gyv > e) th = M<real, MATH>::atan_(gyv / gxv);
            }
            if (gxv < -e) th = pi + M<real, MATH>::atan_(gyv / gxv);
            angl[c] = th;
            const real arg = (theta0 == (real)0) ? aniso * th : aniso * (th - theta0);
            eps[c] = epsbar * ((real)1.0f + delta * M<real, MATH>::cos_(arg));       // :170
            epsd[c] = -epsbar * aniso * delta * M<real, MATH>::sin_(arg);            // :171
        }
    }

    template <int MATH>
    void pass2_row(int64_t jp, std::vector<real>& phi_out, std::vector<real>& t_out) {  // G <= jp < ny+G
        const real pi = (real)kob::REF_PI_F;
        const bool noisy = (noise_a != (real)0);
        for (int64_t i = 0; i < nx; ++i) {
            const int64_t ip = (i + 1) % nx, im = ((i - 1) + nx) % nx;
            const int64_t jpp = jp + 1, jm = jp - 1;
            const size_t c = idx(i, jp), E = idx(ip, jp), W = idx(im, jp), N = idx(i, jpp), S = idx(i, jm);
            const real gepx = (eps[E] * eps[E] - eps[W] * eps[W]) / dx;              // :190-192
            const real gepy = (eps[N] * eps[N] - eps[S] * eps[S]) / dy;              // :193-195
            const real term1 = (eps[N] * epsd[N] * gx[N] - eps[S] * epsd[S] * gx[S]) / dy;   // :197-199
            const real term2 = -(eps[E] * epsd[E] * gy[E] - eps[W] * epsd[W] * gy[W]) / dx;  // :201-203
            const real term3 = gepx * gx[c] + gepy * gy[c];                          // :204
            const real m = alpha / pi * M<real, MATH>::atan_(gamma_ * (teq - t[c])); // :206
            const real op = phi[c], ot = t[c];
            const real q = op * ((real)1.0f - op);
            real sum = term1 + term2 + eps[c] * eps[c] * lapphi[c] + term3 + q * (op - (real)0.5f + m);  // :212-214
            if (noisy) {
                const int64_t jl = jp - G;
                const float r = noise_field.empty()
                                    ? kob::noise_r(seed, step_counter, (uint32_t)(noise_x0 + i), (uint32_t)(noise_y0 + y0 + jl))
                                    : noise_field[(size_t)i + (size_t)nx * (size_t)jl];
                sum = sum + (noise_a * q) * ((real)r - (real)0.5f);
            }
            const real np = op + sum * dt / tau;                                     // :211,214
            phi_out[c] = np;
            t_out[c] = ot + lapt[c] * dt + K * (np - op);                            // :215
        }
    }

    template <int MATH>
    void step_impl(int64_t n) {
        for (int64_t s = 0; s < n; ++s) {
            if (self_periodic) wrap_ghosts();
#pragma omp parallel for schedule(static) num_threads(threads)
            for (int64_t jp = G - 1; jp <= ny + G; ++jp) pass1_row<MATH>(jp);
            // pass 2 reads phi/t only at the centre cell, so the in-place update of the reference
            // (src/Kobayashi.cpp:211,215) is order independent; writing in place is exact.
#pragma omp parallel for schedule(static) num_threads(threads)
            for (int64_t jp = G; jp < ny + G; ++jp) pass2_row<MATH>(jp, phi, t);
            ++step_counter;
        }
    }
    void step(int64_t n) override {
        if (math == MATH_PORTABLE) step_impl<MATH_PORTABLE>(n);
        else step_impl<MATH_LIBM>(n);
    }
    void get(void* ph, void* tt, void* an) const override {
        const size_t bytes = sizeof(real) * (size_t)nx * (size_t)ny;
        if (ph) std::memcpy(ph, &phi[idx(0, G)], bytes);
        if (tt) std::memcpy(tt, &t[idx(0, G)], bytes);
        if (an) std::memcpy(an, &angl[idx(0, G)], bytes);
    }
    void set(const void* ph, const void* tt, const void* an) override {
        const size_t bytes = sizeof(real) * (size_t)nx * (size_t)ny;
        if (ph) std::memcpy(&phi[idx(0, G)], ph, bytes);
        if (tt) std::memcpy(&t[idx(0, G)], tt, bytes);
        if (an) std::memcpy(&angl[idx(0, G)], an, bytes);
    }
    // side 0: this strip's two LOWEST owned rows (local 0,1); side 1: two HIGHEST (ny-2, ny-1); 2*nx each.
    void get_edge(int side, void* ph, void* tt, void* an) const override {
        const int64_t r0 = side == 0 ? 0 : ny - G;
        const size_t bytes = sizeof(real) * (size_t)nx * G;
        if (ph) std::memcpy(ph, &phi[idx(0, r0 + G)], bytes);
        if (tt) std::memcpy(tt, &t[idx(0, r0 + G)], bytes);
        if (an) std::memcpy(an, &angl[idx(0, r0 + G)], bytes);
    }
    // side 0: ghost rows below (local -2,-1); side 1: ghost rows above (ny, ny+1).
    void set_ghost(int side, const void* ph, const void* tt, const void* an) override {
        const int64_t r0 = side == 0 ? 0 : ny + G;
        const size_t bytes = sizeof(real) * (size_t)nx * G;
        if (ph) std::memcpy(&phi[idx(0, r0)], ph, bytes);
        if (tt) std::memcpy(&t[idx(0, r0)], tt, bytes);
        if (an) std::memcpy(&angl[idx(0, r0)], an, bytes);
    }
};

}  // namespace

extern "C" {

void* kobo_create(int prec, int64_t nx, int64_t ny, int64_t ny_global, int64_t y0, const kob_params* p,
                  int math, uint64_t seed) {
    if (nx < 1 || ny < 1 || !p) return nullptr;
    OracleBase* o = nullptr;
    try {
        if (prec == 64) { auto* q = new Oracle<double>(); q->init(nx, ny, ny_global, y0, *p); o = q; }
        else { auto* q = new Oracle<float>(); q->init(nx, ny, ny_global, y0, *p); o = q; }
    } catch (const std::bad_alloc&) { return nullptr; }
    o->prec = prec == 64 ? 64 : 32; o->math = math; o->seed = seed;
    return o;
}
void kobo_destroy(void* h) { delete static_cast<OracleBase*>(h); }
void kobo_clear(void* h) { static_cast<OracleBase*>(h)->clear(); }
void kobo_add_nucleus(void* h, int64_t x, int64_t y) { static_cast<OracleBase*>(h)->add_nucleus(x, y); }
// _vectorInit (src/Kobayashi.cpp:98-114)
void kobo_reset(void* h) {
    auto* o = static_cast<OracleBase*>(h);
    o->clear();
    o->add_nucleus(o->nx / 2, o->ny_global / 2);
}
void kobo_set_params(void* h, const kob_params* p) {
    auto* o = static_cast<OracleBase*>(h);
    if (o->prec == 64) static_cast<Oracle<double>*>(o)->set_params(*p);
    else static_cast<Oracle<float>*>(o)->set_params(*p);
}
void kobo_step(void* h, int64_t n) { static_cast<OracleBase*>(h)->step(n); }
void kobo_get_fields(void* h, void* phi, void* t, void* angl) { static_cast<OracleBase*>(h)->get(phi, t, angl); }
void kobo_set_fields(void* h, const void* phi, const void* t, const void* angl) {
    static_cast<OracleBase*>(h)->set(phi, t, angl);
}
void kobo_get_edge(void* h, int side, void* phi, void* t, void* angl) {
    static_cast<OracleBase*>(h)->get_edge(side, phi, t, angl);
}
void kobo_set_ghost(void* h, int side, const void* phi, const void* t, const void* angl) {
    static_cast<OracleBase*>(h)->set_ghost(side, phi, t, angl);
}
void kobo_set_noise_field(void* h, const float* r) {
    auto* o = static_cast<OracleBase*>(h);
    if (!r) { o->noise_field.clear(); return; }
    o->noise_field.assign(r, r + (size_t)o->nx * (size_t)o->ny);
}
void kobo_set_step_counter(void* h, uint64_t s) { static_cast<OracleBase*>(h)->step_counter = s; }
uint64_t kobo_get_step_counter(void* h) { return static_cast<OracleBase*>(h)->step_counter; }
void kobo_set_threads(void* h, int n) { static_cast<OracleBase*>(h)->threads = n < 1 ? 1 : n; }
// The oracle grid is a WINDOW of a larger torus: local cell (0, 0) is global cell (x0, y0) for the Philox noise keys
// (full-size parity tests compare a window of a 16384^2 GPU run with a 250^2 oracle run).
void kobo_set_noise_origin(void* h, int64_t x0, int64_t y0) { OracleBase* o = static_cast<OracleBase*>(h); o->noise_x0 = x0; o->noise_y0 = y0; }
int kobo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void kobo_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    const kob::Philox4 p = kob::philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    for (int i = 0; i < 4; ++i) out[i] = p.w[i];
}
float kobo_noise_r(uint64_t seed, uint64_t step, uint32_t i, uint32_t j) { return kob::noise_r(seed, step, i, j); }
// portable math probes (accuracy tests against libm)
float kobo_p_atanf(float x) { return kob::p_atan(x); }
float kobo_p_sinf(float x) { return kob::p_sin(x); }
float kobo_p_cosf(float x) { return kob::p_cos(x); }
double kobo_p_atan(double x) { return kob::p_atan(x); }
double kobo_p_sin(double x) { return kob::p_sin(x); }
double kobo_p_cos(double x) { return kob::p_cos(x); }

}  // extern "C"
