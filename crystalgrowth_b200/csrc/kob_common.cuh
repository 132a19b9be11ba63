// kob_common.cuh — device-side data layout and helpers shared by all kernels of libkobayashi_cuda.
//
// HBM layout of one strip (one context = one GPU's rows [y0, y0+ny) of the nx x ny_global torus):
//
//   padded row:   | GX=4 ghost cols | nx cells | GXR=4 ghost cols | pad to 32 elems |
//   padded array: | GY=4 ghost rows | ny rows | GY=4 ghost rows |
// (the single-step kernels read ghosts to depth 2; the two-step kernel, whose composed stencil has radius 4, to depth 4)
//
// phi and T are double buffered (Jacobi step: read `cur`, write `cur^1`); theta is single buffered and
// updated in place (a cell's theta is either re-assigned — then nobody reads the old value — or held —
// then nobody writes it; both decisions are functions of the same phi bits in every tile that looks).
// Ghost columns/rows hold periodic copies.  They are never refreshed by a separate pass: the tile that
// OWNS an edge cell stores its new value to every alias (own ghost columns, and the ghost rows of the
// lower/upper neighbour strip — which is this same strip when P = 1, or a peer GPU's memory over NVLink).
// Loads therefore never wrap, which keeps them branch-free and TMA/float4 friendly.
//
// The reference indexes cells as i + nx*j with modulo wrap (src/Kobayashi.h:91, src/Kobayashi.cpp:133-136).
#ifndef KOB_COMMON_CUH
#define KOB_COMMON_CUH

#include <cuda_runtime.h>
#include <stdint.h>
#include "kob_math.h"

namespace kob {

constexpr int GX = 4;    // ghost columns on the left (keeps the interior 16/32-byte aligned)
constexpr int GXR = 4;   // ghost columns on the right
constexpr int GY = 4;    // ghost rows per side
constexpr int FBX = 128; // theta-flag block, padded coordinates
constexpr int FBY = 32;

// One strip's device arrays as seen by a kernel (its own, or a neighbour's over NVLink / same device).
template <typename real>
struct StripView {
    real* phi[2];
    real* t[2];
    real* theta;        // the current angle buffer (single-step kernels update it in place)
    real* theta_next;   // the other one: the two-step kernel reads `theta` and writes `theta_next`, then they swap
    uint32_t* tflags;   // [nfby][nfbx] : 1 = some theta in this block of the padded array may be non-zero
    uint32_t* arrive;   // [3]: arrive[0] written by the lower neighbour, arrive[1] by the upper neighbour,
                        //      arrive[2] = fault word (set when a wait on a neighbour timed out)
    long long ny;       // rows owned by that strip
};

// Parameters rounded once to `real` (the reference's float members, src/Kobayashi.h:94-105) plus the
// loop-invariant sub-expressions of the reference formulas, evaluated in the reference's order.
template <typename real>
struct KParams {
    real dx, dy, dt, tau, epsbar, K, delta, aniso, alpha, gamma, teq, theta0, noise_a;
    real lapden;         // (3.0f*dx)*dx                 src/Kobayashi.cpp:146
    real neg_ebjd;       // ((-epsbar)*aniso)*delta      src/Kobayashi.cpp:171
    real alpha_over_pi;  // alpha / PI_F                 src/Kobayashi.cpp:206
    // fast-kernel reciprocals (rounding-level substitutes for the divisions by loop constants)
    real inv_dx, inv_dy, inv_lapden, dt_over_tau;
    int jmode;           // integer anisotropy mode if aniso is a small integer, else 0
};

template <typename real>
struct StepArgs {
    StripView<real> self, lower, upper;
    KParams<real> prm;
    const float* noise_field;  // host-injected r (unpadded nx*ny) or nullptr -> Philox
    unsigned long long seed, step;
    unsigned int* ticket;      // CTA completion counter (linked strips only)
    long long pitch;           // elements per padded row
    int nx, ny;                // owned cells
    long long y0;              // global row of local row 0
    int nfbx, nfby;            // theta-flag blocks
    int cur;                   // buffer read by this step
    int tcur;                  // which theta buffer `theta` is (the two-step kernel swaps them)
    int linked;                // 1: neighbours are other strips -> wait/signal through `arrive`
    unsigned int epoch;        // number of SUB-STEPS this strip has completed before this launch
};

template <typename real>
__device__ __forceinline__ long long pidx(long long pitch, int i, int j) {  // local cell (i, j), ghosts allowed
    return (long long)(j + GY) * pitch + (i + GX);
}

// Store v for owned cell (i, j) to every alias: own array, own ghost columns, neighbours' ghost rows.
template <typename real>
__device__ __forceinline__ void store_aliases(real* __restrict__ self_buf, real* lower_buf, real* upper_buf,
                                              long long pitch, int nx, int ny, long long ny_lower, int i, int j,
                                              real v) {
    // x aliases: the cell itself, its right-hand ghost copy (columns 0..GXR-1 -> nx..), its left-hand ghost copy
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (k == 1 && !(i < GXR)) continue;
        if (k == 2 && !(i >= nx - GXR)) continue;
        const int x = k == 0 ? i : (k == 1 ? i + nx : i - nx);
        self_buf[pidx<real>(pitch, x, j)] = v;
        if (j < GY) lower_buf[(long long)(ny_lower + j + GY) * pitch + (x + GX)] = v;
        if (j >= ny - GY) upper_buf[(long long)(j - ny + GY) * pitch + (x + GX)] = v;
    }
}

// Mark every theta-flag block overlapped by the padded-coordinate rectangle [xlo,xhi] x [ylo,yhi].
__device__ __forceinline__ void mark_flags(uint32_t* tflags, int nfbx, int nfby, long long xlo, long long xhi,
                                           long long ylo, long long yhi) {
    int bx0 = (int)(xlo / FBX), bx1 = (int)(xhi / FBX), by0 = (int)(ylo / FBY), by1 = (int)(yhi / FBY);
    bx0 = max(bx0, 0); by0 = max(by0, 0); bx1 = min(bx1, nfbx - 1); by1 = min(by1, nfby - 1);
    for (int by = by0; by <= by1; ++by)
        for (int bx = bx0; bx <= bx1; ++bx) tflags[by * nfbx + bx] = 1u;
}

// OR of the flags overlapped by a padded-coordinate rectangle.
__device__ __forceinline__ uint32_t any_flags(const uint32_t* tflags, int nfbx, int nfby, long long xlo,
                                              long long xhi, long long ylo, long long yhi) {
    int bx0 = (int)(xlo / FBX), bx1 = (int)(xhi / FBX), by0 = (int)(ylo / FBY), by1 = (int)(yhi / FBY);
    bx0 = max(bx0, 0); by0 = max(by0, 0); bx1 = min(bx1, nfbx - 1); by1 = min(by1, nfby - 1);
    uint32_t f = 0;
    for (int by = by0; by <= by1; ++by)
        for (int bx = bx0; bx <= bx1; ++bx) f |= __ldcg(&tflags[by * nfbx + bx]);
    return f;
}

// A tile that assigned some theta marks its own block(s) and, for edge tiles, the blocks of every alias.
template <typename real>
__device__ __forceinline__ void mark_tile_flags(const StepArgs<real>& a, int x0, int y0, int tx, int ty) {
    const int x1 = min(x0 + tx, a.nx) - 1, y1 = min(y0 + ty, a.ny) - 1;
    // x alias ranges (padded coordinates)
    long long xr[3][2];
    int nxr = 0;
    xr[nxr][0] = x0 + GX; xr[nxr][1] = x1 + GX; ++nxr;
    if (x0 < GXR) { xr[nxr][0] = a.nx + GX; xr[nxr][1] = a.nx + GX + GXR - 1; ++nxr; }
    if (x1 >= a.nx - GXR) { xr[nxr][0] = GX - GXR; xr[nxr][1] = GX - 1; ++nxr; }
    for (int k = 0; k < nxr; ++k) {
        mark_flags(a.self.tflags, a.nfbx, a.nfby, xr[k][0], xr[k][1], y0 + GY, y1 + GY);
        if (y0 < GY) {
            const int nfby_l = (int)((a.lower.ny + 2 * GY + FBY - 1) / FBY);
            mark_flags(a.lower.tflags, a.nfbx, nfby_l, xr[k][0], xr[k][1], a.lower.ny + GY, a.lower.ny + 2 * GY - 1);
        }
        if (y1 >= a.ny - GY) {
            const int nfby_u = (int)((a.upper.ny + 2 * GY + FBY - 1) / FBY);
            mark_flags(a.upper.tflags, a.nfbx, nfby_u, xr[k][0], xr[k][1], 0, GY - 1);
        }
    }
}

// ---- cross-strip step flags (linked strips only) ---------------------------------------------------
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Edge tiles wait until the neighbour whose ghost rows they read (and whose ghost rows they will write)
// has completed `epoch` steps.  Called by all threads of the CTA.
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
constexpr unsigned long long WAIT_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;

// Bounded: a neighbour that never arrives (its process died, or strips were stepped out of order) raises the
// strip's fault word instead of hanging the GPU; the host reports it from kob_sync / kob_get_fields.
// `epoch` counts SUB-STEPS completed before this launch and `sub` the sub-steps this launch performs (1, or 2 for a two-step
// launch pair).  A neighbour is never more than one launch ahead, so its flag reads epoch (not there yet) or epoch + sub;
// anything else means the strips of the ring are not running the same launch sequence (one does pairs, the other single
// steps): that is reported at once through the fault word instead of after a 20 s spin.
__device__ __forceinline__ void wait_flag(const uint32_t* flag, uint32_t epoch, uint32_t* fault, uint32_t sub = 0u) {
    uint32_t v = ld_acquire_sys(flag);
    if ((int)(v - epoch) < 0) {
        const unsigned long long t0 = globaltimer_ns();
        while ((int)((v = ld_acquire_sys(flag)) - epoch) < 0) {
            __nanosleep(100);
            if (globaltimer_ns() - t0 > WAIT_TIMEOUT_NS) { atomicExch(fault, 1u); return; }
        }
        atomicAdd(fault + 2, (uint32_t)((globaltimer_ns() - t0) >> 10));   // diagnostics (kob_wait_stats): ~us spent waiting, waits
        atomicAdd(fault + 3, 1u);
    }
    if (sub != 0u && v != epoch && v != epoch + sub) atomicExch(fault, 2u);
}

template <typename real>
__device__ __forceinline__ void wait_neighbours(const StepArgs<real>& a, bool touches_low, bool touches_high) {
    if (!a.linked) return;
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        if (touches_low) wait_flag(&a.self.arrive[0], a.epoch, &a.self.arrive[2], 1u);
        if (touches_high) wait_flag(&a.self.arrive[1], a.epoch, &a.self.arrive[2], 1u);
    }
    __syncthreads();
}

// Last CTA of the grid publishes "this strip completed epoch + 1 sub-steps" into both neighbours (STRICT kernel; the FAST
// kernels publish each side as soon as the jobs touching it are done, kob_fast.cuh).
template <typename real>
__device__ __forceinline__ void signal_neighbours(const StepArgs<real>& a) {
    if (!a.linked) return;
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        __threadfence_system();
        const unsigned int total = gridDim.x * gridDim.y;
        const unsigned int done = atomicAdd(a.ticket, 1u) + 1u;
        if (done == total) {
            *a.ticket = 0u;
            __threadfence_system();
            st_release_sys(&a.lower.arrive[1], a.epoch + 1u);  // I am my lower neighbour's upper neighbour
            st_release_sys(&a.upper.arrive[0], a.epoch + 1u);
        }
    }
}

}  // namespace kob
#endif  // KOB_COMMON_CUH
