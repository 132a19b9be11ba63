// kob_aux.cuh — small kernels around the step: alias refresh, theta-flag rebuild, nuclei, colour ramp.
#ifndef KOB_AUX_CUH
#define KOB_AUX_CUH

#include "kob_common.cuh"

namespace kob {

// Re-store every edge cell (within GXR columns / GY rows of the strip border) of phi[cur], T[cur] and theta
// to all of its aliases.  Used after host writes (set_fields, nuclei, reset) — the step kernels keep the
// aliases current by themselves.  In linked mode it also publishes `epoch` to both neighbours.
template <typename real>
__global__ void kob_refresh_aliases(const StepArgs<real> a) {
    const long long nrow_items = 2LL * GY * a.nx;           // rows 0..GY-1 and ny-GY..ny-1, all columns
    const long long ncol_items = 2LL * GXR * a.ny;          // cols 0..GXR-1 and nx-GXR..nx-1, all rows
    const long long total = nrow_items + ncol_items;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total;
         k += (long long)gridDim.x * blockDim.x) {
        int i, j;
        if (k < nrow_items) {
            const int r = (int)(k / a.nx);
            i = (int)(k - (long long)r * a.nx);
            j = r < GY ? r : a.ny - 2 * GY + r;
        } else {
            const long long kk = k - nrow_items;
            const int c = (int)(kk / a.ny);
            j = (int)(kk - (long long)c * a.ny);
            i = c < GXR ? c : a.nx - 2 * GXR + c;
        }
        if (i < 0 || j < 0 || i >= a.nx || j >= a.ny) continue;
        const long long p = pidx<real>(a.pitch, i, j);
        store_aliases<real>(a.self.phi[a.cur], a.lower.phi[a.cur], a.upper.phi[a.cur], a.pitch, a.nx, a.ny,
                            a.lower.ny, i, j, a.self.phi[a.cur][p]);
        store_aliases<real>(a.self.t[a.cur], a.lower.t[a.cur], a.upper.t[a.cur], a.pitch, a.nx, a.ny,
                            a.lower.ny, i, j, a.self.t[a.cur][p]);
        const real th = a.self.theta[p];
        store_aliases<real>(a.self.theta, a.lower.theta, a.upper.theta, a.pitch, a.nx, a.ny, a.lower.ny, i, j, th);
        if (th != (real)0) mark_tile_flags(a, i, j, 1, 1);   // set-only: flags of every alias of this cell
    }
}

template <typename real>
__global__ void kob_publish_epoch(const StepArgs<real> a) {
    if (a.linked && blockIdx.x == 0 && threadIdx.x == 0) {
        __threadfence_system();
        st_release_sys(&a.lower.arrive[1], a.epoch);
        st_release_sys(&a.upper.arrive[0], a.epoch);
    }
}

// One CTA per theta-flag block: flag = any(theta != 0) over the block of the padded array.
template <typename real>
__global__ void kob_rebuild_flags(const real* __restrict__ theta, uint32_t* __restrict__ tflags, long long pitch,
                                  long long rows, int nfbx) {
    const int bx = blockIdx.x, by = blockIdx.y;
    int any = 0;
    for (int k = threadIdx.x; k < FBX * FBY; k += blockDim.x) {
        const long long xp = (long long)bx * FBX + (k % FBX), yp = (long long)by * FBY + (k / FBX);
        if (xp < pitch && yp < rows && theta[yp * pitch + xp] != (real)0) any = 1;
    }
    any = __syncthreads_or(any);
    // Blocks that contain ghost rows are only ever SET here: a linked neighbour's alias refresh may be setting the same flag
    // at this very moment (its theta rows land in my ghost rows), and clearing it after that store would leave held angles
    // unseen.  A stale set flag is just a conservative hint.
    const bool ghost_block = by * FBY < GY || (long long)(by + 1) * FBY > rows - GY;
    if (threadIdx.x == 0) {
        if (any) tflags[by * nfbx + bx] = 1u;
        else if (!ghost_block) tflags[by * nfbx + bx] = 0u;
    }
}

// _createNucleus (src/Kobayashi.cpp:116-123) at GLOBAL cell (x, y), periodic wrap, this strip's share only.
template <typename real>
__global__ void kob_nucleus(const StepArgs<real> a, long long x, long long y, long long ny_global) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const long long ox[5] = {0, -1, 1, 0, 0}, oy[5] = {0, 0, 0, -1, 1};
    for (int k = 0; k < 5; ++k) {
        const long long gx_ = (((x + ox[k]) % a.nx) + a.nx) % a.nx;
        const long long gy_ = (((y + oy[k]) % ny_global) + ny_global) % ny_global;
        const long long jl = gy_ - a.y0;
        if (jl < 0 || jl >= a.ny) continue;
        store_aliases<real>(a.self.phi[a.cur], a.lower.phi[a.cur], a.upper.phi[a.cur], a.pitch, a.nx, a.ny,
                            a.lower.ny, (int)gx_, (int)jl, (real)1);
    }
}

// Viewer colour ramp, iUpdateConstantBuffer (src/Kobayashi.cpp:315-344): piecewise-linear blend of four
// colours over phi <= 0.9, (0.9, 0.99], > 0.99; evaluated in float like the reference, then quantised to RGBA8.
template <typename real>
__global__ void kob_render(const real* __restrict__ phi, uint8_t* __restrict__ rgba, long long pitch, int nx,
                           int ny) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= nx || j >= ny) return;
    const float p = (float)phi[pidx<real>(pitch, i, j)];
    const float c0[3] = {0.0f, 0.0f, 0.0f};
    const float c1[3] = {0.2505490f, 0.5f, 0.9882353f};
    const float c2[3] = {0.3607843f, 1.0f, 0.9882353f};
    const float c3[3] = {0.9005490f, 1.0f, 0.9882353f};
    const float b1 = 0.9f, b2 = 0.99f, b3 = 1.0f;
    const float *lo, *hi;
    float ratio;
    if (p <= b1) { ratio = __fmul_rn(p, __fdiv_rn(1.0f, b1)); lo = c0; hi = c1; }
    else if (p <= b2) { ratio = __fmul_rn(__fsub_rn(p, b1), __fdiv_rn(1.0f, __fsub_rn(b2, b1))); lo = c1; hi = c2; }
    else { ratio = __fmul_rn(__fsub_rn(p, b2), __fdiv_rn(1.0f, __fsub_rn(b3, b2))); lo = c2; hi = c3; }
    uint8_t* o = rgba + 4LL * ((long long)i + (long long)nx * j);
    for (int c = 0; c < 3; ++c) {
        const float v = __fadd_rn(__fmul_rn(lo[c], __fsub_rn(1.0f, ratio)), __fmul_rn(hi[c], ratio));
        o[c] = (uint8_t)__float2int_rn(fminf(fmaxf(v, 0.0f), 1.0f) * 255.0f);
    }
    o[3] = 255;
}

}  // namespace kob
#endif  // KOB_AUX_CUH
