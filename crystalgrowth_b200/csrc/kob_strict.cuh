// kob_strict.cuh — STRICT fused Kobayashi step: one launch = pass 1 + pass 2 of the reference
// (src/Kobayashi.cpp:125-175 and :177-221) with NO scratch arrays in HBM.
//
// "Strict" = every floating-point operation of the reference, in the reference's order, executed with
// correctly rounded non-contracted primitives (rn_add/rn_mul/rn_div -> __fadd_rn/... ) and the portable
// atan/sin/cos of kob_math.h.  The CPU oracle built with the same math provider produces the same bits,
// which is what lets parity be checked bit-for-bit through the rounding-chaotic regime.
// Works for float and double.  Roofline kernel: see kob_fast.cuh.
//
// Tile: TX x TY owned cells per CTA.  Stage 0: phi tile with a 2-cell halo and T tile with a 1-cell halo
// go to shared memory (plain loads; ghost copies make wrap-free addressing possible).  Stage 1: pass 1 on
// the (TX+2) x (TY+2) ring (gradients, angle state machine, eps, eps') into shared memory.  Stage 2: pass 2
// on the owned cells; phi, T (and re-assigned theta) are written once, to every alias of the cell.
#ifndef KOB_STRICT_CUH
#define KOB_STRICT_CUH

#include "kob_common.cuh"

namespace kob {

// x / d for d > 0 (dx, dy, 3 dx^2, tau: validated at kob_create): a zero numerator — the far field, where this kernel spends most of
// its launches — gives the zero back with its sign, which is exactly what the correctly rounded division returns; everything else
// goes through it.  Same bits, ~10 instructions less per division on rows without crystal.
template <typename real>
__device__ __forceinline__ real div_pos(real x, real d) { return x == (real)0 ? x : rn_div(x, d); }
__device__ __forceinline__ bool is_plus_zero(float v) { return __float_as_uint(v) == 0u; }
__device__ __forceinline__ bool is_plus_zero(double v) { return __double_as_longlong(v) == 0ll; }

template <typename real, int TX, int TY, bool NOISE>
__global__ void __launch_bounds__(256) kob_step_strict(const StepArgs<real> a) {
    constexpr int PW = TX + 4, PH = TY + 4;   // phi tile
    constexpr int RW = TX + 2, RH = TY + 2;   // ring (pass-1) tile, also T tile
    __shared__ real s_phi[PH][PW];
    __shared__ real s_t[RH][RW];
    __shared__ real s_eps[RH][RW];
    __shared__ real s_epsd[RH][RW];
    __shared__ real s_gx[RH][RW];
    __shared__ real s_gy[RH][RW];
    __shared__ uint32_t s_flag, s_assigned;

    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int nthreads = blockDim.x * blockDim.y;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const long long pitch = a.pitch;
    const long long rows = (long long)a.ny + 2 * GY;
    const KParams<real>& P = a.prm;

    wait_neighbours(a, y0 == 0, y0 + TY >= a.ny - 1);

    const real* __restrict__ phi_in = a.self.phi[a.cur];
    const real* __restrict__ t_in = a.self.t[a.cur];

    if (tid == 0) {
        s_flag = any_flags(a.self.tflags, a.nfbx, a.nfby, x0 + GX - 1, x0 + GX + TX, y0 + GY - 1, y0 + GY + TY);
        s_assigned = 0u;
    }
    // ---- stage 0: tiles to shared memory ----
    for (int k = tid; k < PW * PH; k += nthreads) {
        const int ly = k / PW, lx = k - ly * PW;
        const long long xp = x0 + GX - 2 + lx, yp = y0 + GY - 2 + ly;
        s_phi[ly][lx] = (xp < pitch && yp < rows) ? phi_in[yp * pitch + xp] : (real)0;
    }
    for (int k = tid; k < RW * RH; k += nthreads) {
        const int ly = k / RW, lx = k - ly * RW;
        const long long xp = x0 + GX - 1 + lx, yp = y0 + GY - 1 + ly;
        s_t[ly][lx] = (xp < pitch && yp < rows) ? t_in[yp * pitch + xp] : (real)0;
    }
    __syncthreads();
    const bool theta_live = s_flag != 0u;
    // eps, eps' of a cell that holds theta = +0 (every cell that never saw an interface): the same expressions as below on the
    // same input, evaluated once per thread instead of once per cell
    real eps0, epsd0;
    {
        const real th0 = (real)0;
        const real arg0 = (P.theta0 == (real)0) ? rn_mul(P.aniso, th0) : rn_mul(P.aniso, rn_sub(th0, P.theta0));
        real sn0, cs0;
        p_sincos_core(arg0, &sn0, &cs0);
        eps0 = rn_mul(P.epsbar, rn_add((real)1.0f, rn_mul(P.delta, cs0)));
        epsd0 = rn_mul(P.neg_ebjd, sn0);
    }
    // m(T) of a cell whose temperature is still exactly +0 (far from every crystal): :206 on that input, once per thread
    const real m0 = rn_mul(P.alpha_over_pi, p_atan(rn_mul(P.gamma, rn_sub(P.teq, (real)0))));

    // ---- stage 1: pass 1 on the ring (src/Kobayashi.cpp:139-171) ----
    const real e = (real)REF_DEADBAND;
    const real pi = (real)REF_PI_F;
    bool assigned_any = false;
    for (int k = tid; k < RW * RH; k += nthreads) {
        const int ry = k / RW, rx = k - ry * RW;     // ring coordinates; cell = (x0-1+rx, y0-1+ry)
        const int px = rx + 1, py = ry + 1;          // same cell in the phi tile
        const int ci = x0 - 1 + rx, cj = y0 - 1 + ry;
        const real gx = div_pos(rn_sub(s_phi[py][px + 1], s_phi[py][px - 1]), P.dx);   // :139
        const real gy = div_pos(rn_sub(s_phi[py + 1][px], s_phi[py - 1][px]), P.dy);   // :140
        // angle state machine :154-167
        const bool gx_flat = (gx <= e) && (gx >= -e);
        const bool gy_neg = gy < -e, gy_pos = gy > e;
        const bool gx_pos = gx > e, gx_neg = gx < -e;
        const bool assigned = (gx_flat && (gy_neg || gy_pos)) || (gx_pos && (gy_neg || gy_pos)) || gx_neg;
        real th = (real)0;
        const bool in_grid = (ci < a.nx + GXR) && (cj < a.ny + GY);
        if (assigned) {
            if (gx_flat) th = gy_neg ? rn_mul((real)-0.5f, pi) : rn_mul((real)0.5f, pi);
            else {
                const real at = p_atan(rn_div(gy, gx));
                if (gx_pos) th = gy_neg ? rn_add(rn_mul((real)2.0f, pi), at) : at;
                else th = rn_add(pi, at);
            }
            const bool owned = rx >= 1 && rx <= TX && ry >= 1 && ry <= TY && ci < a.nx && cj < a.ny;
            if (owned) {
                store_aliases<real>(a.self.theta, a.lower.theta, a.upper.theta, pitch, a.nx, a.ny, a.lower.ny, ci, cj, th);
                assigned_any = true;
            }
        } else if (theta_live && in_grid) {
            th = a.self.theta[pidx<real>(pitch, ci, cj)];   // held: keep last angle
        }
        if (is_plus_zero(th)) {
            s_eps[ry][rx] = eps0;
            s_epsd[ry][rx] = epsd0;
        } else {
            const real arg = (P.theta0 == (real)0) ? rn_mul(P.aniso, th) : rn_mul(P.aniso, rn_sub(th, P.theta0));
            real sn, cs;
            p_sincos_core(arg, &sn, &cs);
            s_eps[ry][rx] = rn_mul(P.epsbar, rn_add((real)1.0f, rn_mul(P.delta, cs)));   // :170
            s_epsd[ry][rx] = rn_mul(P.neg_ebjd, sn);                                      // :171
        }
        s_gx[ry][rx] = gx;
        s_gy[ry][rx] = gy;
    }
    if (assigned_any) s_assigned = 1u;   // benign race: every writer stores 1
    __syncthreads();

    // ---- stage 2: pass 2 on the owned cells (src/Kobayashi.cpp:190-215) ----
    real* phi_out = a.self.phi[a.cur ^ 1];
    real* t_out = a.self.t[a.cur ^ 1];
    real* phi_lo = a.lower.phi[a.cur ^ 1];
    real* phi_hi = a.upper.phi[a.cur ^ 1];
    real* t_lo = a.lower.t[a.cur ^ 1];
    real* t_hi = a.upper.t[a.cur ^ 1];
    for (int k = tid; k < TX * TY; k += nthreads) {
        const int ly = k / TX, lx = k - ly * TX;
        const int i = x0 + lx, j = y0 + ly;
        if (i >= a.nx || j >= a.ny) continue;
        const int rx = lx + 1, ry = ly + 1, px = lx + 2, py = ly + 2;
        // Laplacians :142-151, summation order as written
        const real pE = s_phi[py][px + 1], pW = s_phi[py][px - 1], pN = s_phi[py + 1][px], pS = s_phi[py - 1][px];
        real lp = rn_mul((real)2.0f, rn_add(rn_add(rn_add(pE, pW), pN), pS));
        lp = rn_add(lp, s_phi[py + 1][px + 1]);
        lp = rn_add(lp, s_phi[py - 1][px - 1]);
        lp = rn_add(lp, s_phi[py + 1][px - 1]);
        lp = rn_add(lp, s_phi[py - 1][px + 1]);
        const real op = s_phi[py][px];
        lp = div_pos(rn_sub(lp, rn_mul((real)12.0f, op)), P.lapden);
        const real tE = s_t[ry][rx + 1], tW = s_t[ry][rx - 1], tN = s_t[ry + 1][rx], tS = s_t[ry - 1][rx];
        real lt = rn_mul((real)2.0f, rn_add(rn_add(rn_add(tE, tW), tN), tS));
        lt = rn_add(lt, s_t[ry + 1][rx + 1]);
        lt = rn_add(lt, s_t[ry - 1][rx - 1]);
        lt = rn_add(lt, s_t[ry + 1][rx - 1]);
        lt = rn_add(lt, s_t[ry - 1][rx + 1]);
        const real ot = s_t[ry][rx];
        lt = div_pos(rn_sub(lt, rn_mul((real)12.0f, ot)), P.lapden);

        const real eC = s_eps[ry][rx], eE = s_eps[ry][rx + 1], eW = s_eps[ry][rx - 1], eN = s_eps[ry + 1][rx], eS = s_eps[ry - 1][rx];
        const real gepx = div_pos(rn_sub(rn_mul(eE, eE), rn_mul(eW, eW)), P.dx);           // :190-192
        const real gepy = div_pos(rn_sub(rn_mul(eN, eN), rn_mul(eS, eS)), P.dy);           // :193-195
        const real term1 = div_pos(rn_sub(rn_mul(rn_mul(eN, s_epsd[ry + 1][rx]), s_gx[ry + 1][rx]),
                                         rn_mul(rn_mul(eS, s_epsd[ry - 1][rx]), s_gx[ry - 1][rx])), P.dy);   // :197-199
        const real term2 = div_pos(-rn_sub(rn_mul(rn_mul(eE, s_epsd[ry][rx + 1]), s_gy[ry][rx + 1]),
                                          rn_mul(rn_mul(eW, s_epsd[ry][rx - 1]), s_gy[ry][rx - 1])), P.dx);  // :201-203
        const real term3 = rn_add(rn_mul(gepx, s_gx[ry][rx]), rn_mul(gepy, s_gy[ry][rx]));                   // :204
        const real m = is_plus_zero(ot) ? m0 : rn_mul(P.alpha_over_pi, p_atan(rn_mul(P.gamma, rn_sub(P.teq, ot))));   // :206
        const real q = rn_mul(op, rn_sub((real)1.0f, op));
        real sum = rn_add(term1, term2);
        sum = rn_add(sum, rn_mul(rn_mul(eC, eC), lp));
        sum = rn_add(sum, term3);
        sum = rn_add(sum, rn_mul(q, rn_add(rn_sub(op, (real)0.5f), m)));                                     // :212-214
        if (NOISE) {
            // extension: + a*phi(1-phi)*(r - 1/2); skipped where q == 0 (adds an exact zero there)
            if (q != (real)0) {
                const float r = a.noise_field ? a.noise_field[(long long)i + (long long)a.nx * j]
                                              : noise_r(a.seed, a.step, (uint32_t)i, (uint32_t)(a.y0 + j));
                sum = rn_add(sum, rn_mul(rn_mul(P.noise_a, q), rn_sub((real)r, (real)0.5f)));
            }
        }
        const real np = rn_add(op, div_pos(rn_mul(sum, P.dt), P.tau));                                        // :211,214
        const real nt = rn_add(rn_add(ot, rn_mul(lt, P.dt)), rn_mul(P.K, rn_sub(np, op)));                   // :215
        store_aliases<real>(phi_out, phi_lo, phi_hi, pitch, a.nx, a.ny, a.lower.ny, i, j, np);
        store_aliases<real>(t_out, t_lo, t_hi, pitch, a.nx, a.ny, a.lower.ny, i, j, nt);
    }
    if (tid == 0 && s_assigned) mark_tile_flags(a, x0, y0, TX, TY);
    signal_neighbours(a);
}

}  // namespace kob
#endif  // KOB_STRICT_CUH
