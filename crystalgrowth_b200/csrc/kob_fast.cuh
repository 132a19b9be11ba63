// kob_fast.cuh — FAST fused Kobayashi step (FP32): the roofline kernel.  One launch = one explicit-Euler step
// = pass 1 + pass 2 of the reference (src/Kobayashi.cpp:125-175, :177-221), no scratch arrays in HBM.
//
// Same model as the STRICT kernel (dead-band angle state machine with carried theta, PI_F, 9-point Laplacians,
// Jacobi update); differences are rounding-level only: reciprocal multiplies for the divisions by loop
// constants, FMA contraction, re-associated Laplacian sums, and a TRIG-FREE anisotropy for integer mode j:
// cos(j*theta), sin(j*theta) = Re/Im ((gx + i gy)/|g|)^j — one rsqrt and a few FMAs instead of div+atan+sin+cos.
//
// Structure (B200-first):
//   * persistent grid, one warp = one independent worker; workers pull jobs (column strip x row segment) from a
//     global counter.  There is NO __syncthreads in the steady state.
//   * every worker owns a ring of NST shared-memory stages fed by TMA (cp.async.bulk.tensor.2d + mbarrier
//     complete_tx): a stage carries RB rows of phi and the RB rows of T one row behind it.  Loads are issued
//     NST chunks ahead by one lane; no LSU instruction or register is spent on input traffic.
//   * the warp MARCHES in y: lane L owns CPL = 2*NP adjacent cells of a row; all vertical neighbours (phi, T,
//     eps^2, eps*eps'*gx ...) are earlier rows kept in registers; horizontal neighbours of phi/T come from the
//     stage (LDS), horizontal neighbours of the pass-1 products from the adjacent lanes (SHFL).  Lane 0 and
//     lane 31 are halo lanes: they compute pass 1 for the strip's neighbours' edge cells and store nothing.
//   * results leave through coalesced 8/16-byte global stores; the cells on the strip/torus seams are stored to
//     every alias (own ghost columns, neighbour strips' ghost rows — peer memory over NVLink when P > 1).
//   * theta traffic is predicated: read only where the hold rule fires inside blocks flagged "theta may be
//     non-zero", written only where the state machine re-assigns it.
#ifndef KOB_FAST_CUH
#define KOB_FAST_CUH

#include <cuda.h>   // CUtensorMap (type only; the encode entry point is fetched at run time, no -lcuda)

#include "kob_common.cuh"

namespace kob {

struct FastMaps {
    CUtensorMap phi[2];
    CUtensorMap t[2];
};

struct FastArgs {
    unsigned long long* job_ctr;   // monotonically increasing across launches
    unsigned long long job_base;   // counter value at which this launch's job 0 sits
    int nstrips, nseg, yj;         // jobs = nstrips x nseg; rows per segment
    // constants of the far field / held-with-theta==0 cells: eps and eps' at theta = 0
    float eps0, epsd0;
    float cj0, sj0;                // cos(j*theta0), sin(j*theta0) for the theta0 rotation
    float ebd;                     // epsbar*delta
    float il_dt;                   // inv_lapden*dt
    float two_pi, half_pi;         // 2*PI_F, 0.5*PI_F
};

template <int NP>
struct FastGeom {
    static constexpr int CPL = 2 * NP;            // cells per lane
    static constexpr int WCOLS = 32 * CPL;        // pass-1 columns per warp
    static constexpr int OUTC = WCOLS - 2 * CPL;  // output columns per strip (lanes 1..30)
    static constexpr int BW = WCOLS + 2 * CPL;    // TMA box width (own cells of lane L at box column CPL*L + CPL)
};

constexpr int FAST_RB = 4;     // rows per TMA chunk
constexpr int FAST_NST = 4;    // stages per warp

// one stage = phi box + T box, each padded to a multiple of 128 bytes (TMA shared-memory destination alignment)
template <int NP>
__host__ __device__ constexpr int fast_box_floats() { return (FAST_RB * FastGeom<NP>::BW + 31) / 32 * 32; }
template <int NP>
__host__ __device__ constexpr int fast_stage_floats() { return 2 * fast_box_floats<NP>(); }
template <int NP>
__host__ __device__ constexpr int fast_warp_bytes() { return FAST_NST * fast_stage_floats<NP>() * 4; }

// ---- PTX helpers -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}

// (c + i s)^J by square-and-multiply, J a compile-time constant.
template <int J>
__device__ __forceinline__ void cpow(float c, float s, float& C, float& S) {
    if (J == 0) { C = 1.0f; S = 0.0f; return; }
    if (J == 1) { C = c; S = s; return; }
    float hc, hs;
    cpow<J / 2>(c, s, hc, hs);
    float qc = fmaf(hc, hc, -hs * hs), qs = 2.0f * hc * hs;
    if (J & 1) { C = fmaf(qc, c, -qs * s); S = fmaf(qc, s, qs * c); }
    else { C = qc; S = qs; }
}
__device__ __forceinline__ void cpow_rt(int j, float c, float s, float& C, float& S) {   // 0 <= j <= 16, warp-uniform
    float rc = 1.0f, rs = 0.0f;
#pragma unroll
    for (int bit = 4; bit >= 0; --bit) {
        const float qc = fmaf(rc, rc, -rs * rs), qs = 2.0f * rc * rs;
        rc = qc; rs = qs;
        if ((j >> bit) & 1) { const float tc = fmaf(rc, c, -rs * s), ts = fmaf(rc, s, rs * c); rc = tc; rs = ts; }
    }
    C = rc; S = rs;
}

// JM: 4 / 6 = compile-time integer mode, 0 = run-time integer mode (prm.jmode in 0..16), -1 = any real j (trig).
template <int NP, int JM, bool NOISE, bool ROT>
__global__ void __launch_bounds__(256, 2) kob_step_fast(const __grid_constant__ FastMaps maps,
                                                         const StepArgs<float> a, const FastArgs f) {
    using G = FastGeom<NP>;
    constexpr int CPL = G::CPL, BW = G::BW, RB = FAST_RB, NST = FAST_NST;
    constexpr int STAGE_FLOATS = fast_stage_floats<NP>(), BOX_FLOATS = fast_box_floats<NP>();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    float* stages = reinterpret_cast<float*>(smem_raw) + (size_t)warp * NST * STAGE_FLOATS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)nwarps * fast_warp_bytes<NP>()) + warp * NST;

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NST; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    const KParams<float>& P = a.prm;
    const CUtensorMap* map_phi = a.cur ? &maps.phi[1] : &maps.phi[0];
    const CUtensorMap* map_t = a.cur ? &maps.t[1] : &maps.t[0];
    float* __restrict__ phi_out = a.self.phi[a.cur ^ 1];
    float* __restrict__ t_out = a.self.t[a.cur ^ 1];
    const long long pitch = a.pitch;
    const float e = REF_DEADBAND, pi = REF_PI_F;
    const int njobs = f.nstrips * f.nseg;
    unsigned int gchunk = 0;   // chunks issued/consumed so far by this warp: stage = gchunk % NST, parity = (gchunk / NST) & 1

    for (;;) {
        unsigned long long jraw = 0;
        if (lane == 0) jraw = atomicAdd(f.job_ctr, 1ull) - f.job_base;
        jraw = __shfl_sync(0xffffffffu, jraw, 0);
        if (jraw >= (unsigned long long)njobs) break;
        const int job = (int)jraw;
        const int seg = job / f.nstrips, strip = job - seg * f.nstrips;
        const int y0 = seg * f.yj, y1 = min(y0 + f.yj, a.ny);
        const int xs = strip * G::OUTC - CPL;            // first pass-1 column of the warp (lane 0, halo)
        const int x = xs + CPL * lane;                   // first cell of this lane
        const bool out_lane = lane >= 1 && lane <= 30 && x < a.nx;
        const bool vec_ok = x + CPL <= a.nx;
        const bool edge_lane = x < GXR || x + CPL > a.nx - GXR;

        if (a.linked) {
            if (lane == 0) {
                if (y0 < GY + 1) wait_flag(&a.self.arrive[0], a.epoch, &a.self.arrive[2]);
                if (y1 > a.ny - GY - 1) wait_flag(&a.self.arrive[1], a.epoch, &a.self.arrive[2]);
            }
            __syncwarp();
        }
        // theta may be non-zero somewhere in the pass-1 footprint of this job?
        bool live;
        {
            const int bx0 = max((xs + GX) / FBX, 0), bx1 = min((xs + GX + G::WCOLS - 1) / FBX, a.nfbx - 1);
            const int by0 = max((y0 - 1 + GY) / FBY, 0), by1 = min((y1 + GY) / FBY, a.nfby - 1);
            const int nbx = bx1 - bx0 + 1, nb = nbx * (by1 - by0 + 1);
            uint32_t fl = 0;
            for (int k = lane; k < nb; k += 32) fl |= __ldcg(&a.self.tflags[(by0 + k / nbx) * a.nfbx + bx0 + k % nbx]);
            live = __any_sync(0xffffffffu, fl != 0u);
        }

        const int nrows = (y1 - y0) + 4;                 // streamed phi rows y0-2 .. y1+1
        const int nch = (nrows + RB - 1) / RB;
        const int box_x = xs - CPL + GX;                 // padded x of box column 0
        auto issue = [&](int c) {                        // lane 0: chunk c of this job -> stage (gchunk_issue % NST)
            const unsigned int gi = gchunk + (unsigned int)c;   // gchunk = global index of this job's chunk 0
            const int st = gi % NST;
            float* dst = stages + st * STAGE_FLOATS;
            mbar_expect_tx(&bars[st], 2 * RB * BW * 4);
            const int yr = y0 - 2 + c * RB + GY;         // padded row of the chunk's first phi row
            tma_load_2d(dst, map_phi, box_x, yr, &bars[st]);
            tma_load_2d(dst + BOX_FLOATS, map_t, box_x, yr - 1, &bars[st]);
        };
        if (lane == 0) {
            for (int c = 0; c < NST && c < nch; ++c) issue(c);
        }

        // ---- register state (per cell c of this lane) ----
        float po0[CPL], po1[CPL];          // phi rows r-2, r-1
        float gx1[CPL];                    // gx of row r-1 (pass 1 pending), then row r-2 (pass 2)
        float gx2[CPL], gy2[CPL];          // gx, gy of row r-2
        float u1[CPL], lp1[CPL];           // u(r-1), partial lap (r-1)
        float lap2[CPL];                   // lap phi (r-2) complete
        float tq1[CPL], tu1[CPL], tlp1[CPL];   // T: own (r-2), u (r-2), partial lap (r-2)   [T lags one row]
        float A1[CPL], A2[CPL], A3[CPL];   // eps^2 rows r-1 (new), r-2, r-3
        float P1[CPL], P3[CPL], P2[CPL];   // eps*eps'*gx rows r-1, r-2, r-3
        float Q2[CPL];                     // eps*eps'*gy row r-2
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            po0[c] = po1[c] = gx1[c] = gx2[c] = gy2[c] = u1[c] = lp1[c] = lap2[c] = 0.f;
            tq1[c] = tu1[c] = tlp1[c] = 0.f;
            A1[c] = A2[c] = A3[c] = P1[c] = P2[c] = P3[c] = Q2[c] = 0.f;
        }
        bool assigned_any = false;

        for (int c = 0; c < nch; ++c) {
            const unsigned int gi = gchunk + (unsigned int)c;
            const int st = gi % NST;
            mbar_wait(&bars[st], (gi / NST) & 1u);
            const float* sp = stages + st * STAGE_FLOATS;          // phi rows
            const float* stt = sp + BOX_FLOATS;                     // T rows (one row behind)
#pragma unroll
            for (int rr = 0; rr < RB; ++rr) {
                const int r = y0 - 2 + c * RB + rr;                // phi row streamed in this iteration
                if (r > y1 + 1) break;
                // ---- phi row r ----
                float pn[CPL], hsum[CPL], gxn[CPL];
                {
                    const float* row = sp + rr * BW + CPL * lane + CPL;
                    float w = row[-1], ee = row[CPL];
                    if (NP == 1) { const float2 v = *reinterpret_cast<const float2*>(row); pn[0] = v.x; pn[1] = v.y; }
                    else { const float4 v = *reinterpret_cast<const float4*>(row); pn[0] = v.x; pn[1] = v.y; pn[2 % CPL] = v.z; pn[3 % CPL] = v.w; }
#pragma unroll
                    for (int k = 0; k < CPL; ++k) {
                        const float l = k == 0 ? w : pn[k - 1], rgt = k == CPL - 1 ? ee : pn[k + 1];
                        hsum[k] = l + rgt;
                        gxn[k] = (rgt - l) * P.inv_dx;
                    }
                }
                // ---- pass 1 for row r-1 (needs phi rows r-2, r-1, r) ----
                float An[CPL], Pn[CPL], Qn[CPL], gyn[CPL];
                if (r >= y0) {
                    const int y = r - 1;
                    bool asg[CPL];
                    bool any_asg = false;
#pragma unroll
                    for (int k = 0; k < CPL; ++k) {
                        gyn[k] = (pn[k] - po0[k]) * P.inv_dy;
                        asg[k] = (gx1[k] < -e) || (fabsf(gyn[k]) > e);
                        any_asg |= asg[k];
                    }
                    float th_old[CPL];
#pragma unroll
                    for (int k = 0; k < CPL; ++k) th_old[k] = 0.f;
                    if (live) {   // warp-uniform
#pragma unroll
                        for (int k = 0; k < CPL; ++k)
                            if (!asg[k] && x + k < a.nx + GXR && x + k >= -GXR) th_old[k] = __ldg(&a.self.theta[pidx<float>(pitch, x + k, y)]);
                    }
                    if (!__any_sync(0xffffffffu, any_asg) && !live) {
                        // far field: every cell holds theta = 0
#pragma unroll
                        for (int k = 0; k < CPL; ++k) {
                            An[k] = f.eps0 * f.eps0;
                            const float B = f.eps0 * f.epsd0;
                            Pn[k] = B * gx1[k];
                            Qn[k] = B * gyn[k];
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < CPL; ++k) {
                            const float gx = gx1[k], gy = gyn[k];
                            float C, S;
                            if (asg[k]) {
                                const bool flat = (gx <= e) && (gx >= -e);                  // case A (:154-158)
                                float th;
                                if (JM >= 0) {
                                    const float rinv = rsqrtf(fmaf(gx, gx, gy * gy));
                                    const float c1 = flat ? 0.0f : gx * rinv;
                                    const float s1 = flat ? (gy < 0.f ? -1.0f : 1.0f) : gy * rinv;
                                    if (JM == 0) cpow_rt(P.jmode, c1, s1, C, S); else cpow<JM>(c1, s1, C, S);
                                    if (ROT) { const float c2 = fmaf(C, f.cj0, S * f.sj0), s2 = fmaf(S, f.cj0, -C * f.sj0); C = c2; S = s2; }
                                }
                                const bool owned = out_lane && y >= y0 && y < y1 && x + k < a.nx;
                                if (JM < 0 || owned) {
                                    if (flat) th = gy < 0.f ? -f.half_pi : f.half_pi;
                                    else {
                                        const float at = atanf(__fdiv_rn(gy, gx));
                                        th = gx > 0.f ? (gy < 0.f ? f.two_pi + at : at) : pi + at;     // :160-167
                                    }
                                    if (JM < 0) sincosf(P.aniso * (th - P.theta0), &S, &C);
                                    if (owned) {
                                        store_aliases<float>(a.self.theta, a.lower.theta, a.upper.theta, pitch, a.nx, a.ny,
                                                             a.lower.ny, x + k, y, th);
                                        assigned_any = true;
                                    }
                                }
                            } else if (th_old[k] != 0.f) {
                                sincosf(P.aniso * (th_old[k] - P.theta0), &S, &C);          // held, non-zero angle (rare)
                            } else {
                                C = 1.0f; S = 0.f;                                           // unused: eps0 / epsd0 below
                            }
                            const bool zero_hold = !asg[k] && th_old[k] == 0.f;
                            const float ep = zero_hold ? f.eps0 : fmaf(f.ebd, C, P.epsbar);                // :170
                            const float ed = zero_hold ? f.epsd0 : P.neg_ebjd * S;                          // :171
                            An[k] = ep * ep;
                            const float B = ep * ed;
                            Pn[k] = B * gx;
                            Qn[k] = B * gy;
                        }
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < CPL; ++k) { An[k] = Pn[k] = Qn[k] = gyn[k] = 0.f; }
                }
                // ---- T row r-1: own cells + horizontal neighbours ----
                float tn[CPL], thsum[CPL];
                {
                    const float* row = stt + rr * BW + CPL * lane + CPL;
                    float w = row[-1], ee = row[CPL];
                    if (NP == 1) { const float2 v = *reinterpret_cast<const float2*>(row); tn[0] = v.x; tn[1] = v.y; }
                    else { const float4 v = *reinterpret_cast<const float4*>(row); tn[0] = v.x; tn[1] = v.y; tn[2 % CPL] = v.z; tn[3 % CPL] = v.w; }
#pragma unroll
                    for (int k = 0; k < CPL; ++k) {
                        const float l = k == 0 ? w : tn[k - 1], rgt = k == CPL - 1 ? ee : tn[k + 1];
                        thsum[k] = l + rgt;
                    }
                }
                // ---- pass 2 for row r-2 ----
                if (r >= y0 + 2 && r - 2 < y1) {
                    const int y = r - 2;
                    // horizontal neighbours of the pass-1 products of row y (A2, Q2) from the adjacent lanes
                    const float A_w = __shfl_up_sync(0xffffffffu, A2[CPL - 1], 1);
                    const float A_e = __shfl_down_sync(0xffffffffu, A2[0], 1);
                    const float Q_w = __shfl_up_sync(0xffffffffu, Q2[CPL - 1], 1);
                    const float Q_e = __shfl_down_sync(0xffffffffu, Q2[0], 1);
                    float np_[CPL], nt_[CPL], q[CPL];
                    bool any_q = false;
#pragma unroll
                    for (int k = 0; k < CPL; ++k) { q[k] = fmaf(-po0[k], po0[k], po0[k]); any_q |= (q[k] != 0.f); }
                    // note: at this point po0 = phi(r-2) = phi(y), po1 = phi(r-1)
                    const bool active = __any_sync(0xffffffffu, any_q);
                    float rq[4];
                    if (NOISE && active) {
                        if (a.noise_field) {
#pragma unroll
                            for (int k = 0; k < CPL; ++k)
                                rq[k] = (q[k] != 0.f && x + k >= 0 && x + k < a.nx) ? __ldg(&a.noise_field[(long long)(x + k) + (long long)a.nx * y]) : 0.5f;
                        } else {
                            const Philox4 ph = philox4x32_10((uint32_t)x >> 2, (uint32_t)(a.y0 + y), (uint32_t)a.step,
                                                             (uint32_t)(a.step >> 32), (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
                            if (NP == 2) {
#pragma unroll
                                for (int k = 0; k < 4; ++k) rq[k] = noise_from_word(ph.w[k]);
                            } else {
                                const bool hi = (x & 2) != 0;
                                rq[0] = noise_from_word(hi ? ph.w[2] : ph.w[0]);
                                rq[1] = noise_from_word(hi ? ph.w[3] : ph.w[1]);
                            }
                        }
                    }
#pragma unroll
                    for (int k = 0; k < CPL; ++k) {
                        const float Aw = k == 0 ? A_w : A2[k - 1], Ae = k == CPL - 1 ? A_e : A2[k + 1];
                        const float Qw = k == 0 ? Q_w : Q2[k - 1], Qe = k == CPL - 1 ? Q_e : Q2[k + 1];
                        const float gEx = (Ae - Aw) * P.inv_dx;                               // :190-192
                        const float gEy = (An[k] - A3[k]) * P.inv_dy;                         // :193-195
                        const float t1 = (Pn[k] - P3[k]) * P.inv_dy;                          // :197-199
                        float sum = fmaf(-(Qe - Qw), P.inv_dx, t1);                           // + term2, :201-203
                        sum = fmaf(A2[k], lap2[k] * P.inv_lapden, sum);                       // eps^2 * lap(phi)
                        sum = fmaf(gEx, gx2[k], sum);                                         // term3, :204
                        sum = fmaf(gEy, gy2[k], sum);
                        const float op = po0[k], ot = tq1[k];
                        if (active) {
                            const float m = P.alpha_over_pi * atanf(P.gamma * (P.teq - ot));  // :206
                            sum = fmaf(q[k], (op - 0.5f) + m, sum);                           // :214
                            if (NOISE) sum = fmaf(P.noise_a * q[k], rq[k] - 0.5f, sum);
                        }
                        np_[k] = fmaf(sum, P.dt_over_tau, op);                                // :211
                        const float lapt = tlp1[k] + fmaf(2.0f, tn[k], thsum[k]);              // lap T (y) * 3dx^2
                        nt_[k] = fmaf(P.K, np_[k] - op, fmaf(lapt, f.il_dt, ot));             // :215
                    }
                    if (out_lane) {
                        const bool edge_row = y < GY || y >= a.ny - GY;
                        if (vec_ok && !edge_row && !edge_lane) {
                            const long long o = pidx<float>(pitch, x, y);
                            if (NP == 1) {
                                *reinterpret_cast<float2*>(phi_out + o) = make_float2(np_[0], np_[1]);
                                *reinterpret_cast<float2*>(t_out + o) = make_float2(nt_[0], nt_[1]);
                            } else {
                                *reinterpret_cast<float4*>(phi_out + o) = make_float4(np_[0], np_[1], np_[2 % CPL], np_[3 % CPL]);
                                *reinterpret_cast<float4*>(t_out + o) = make_float4(nt_[0], nt_[1], nt_[2 % CPL], nt_[3 % CPL]);
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < CPL; ++k)
                                if (x + k < a.nx) {
                                    store_aliases<float>(phi_out, a.lower.phi[a.cur ^ 1], a.upper.phi[a.cur ^ 1], pitch, a.nx, a.ny, a.lower.ny, x + k, y, np_[k]);
                                    store_aliases<float>(t_out, a.lower.t[a.cur ^ 1], a.upper.t[a.cur ^ 1], pitch, a.nx, a.ny, a.lower.ny, x + k, y, nt_[k]);
                                }
                        }
                    }
                }
                // ---- rotate the register windows ----
#pragma unroll
                for (int k = 0; k < CPL; ++k) {
                    // T (row r-1 becomes "r-2" of the next iteration)
                    const float tu_new = fmaf(2.0f, tn[k], thsum[k]);                        // u_T(r-1)
                    tlp1[k] = fmaf(2.0f, thsum[k], fmaf(-12.0f, tn[k], tu1[k]));             // c_T(r-1) + u_T(r-2)
                    tu1[k] = tu_new;
                    tq1[k] = tn[k];
                    // phi
                    const float u_new = fmaf(2.0f, pn[k], hsum[k]);                          // u(r)
                    lap2[k] = lp1[k] + u_new;                                                // lap(r-1) complete
                    lp1[k] = fmaf(2.0f, hsum[k], fmaf(-12.0f, pn[k], u1[k]));                // c(r) + u(r-1)
                    u1[k] = u_new;
                    gx2[k] = gx1[k]; gy2[k] = gyn[k]; gx1[k] = gxn[k];
                    po0[k] = po1[k]; po1[k] = pn[k];
                    A3[k] = A2[k]; A2[k] = An[k];
                    P3[k] = P2[k]; P2[k] = Pn[k];
                    Q2[k] = Qn[k];
                }
            }
            __syncwarp();
            if (lane == 0 && c + NST < nch) issue(c + NST);
        }
        gchunk += (unsigned int)nch;
        if (__any_sync(0xffffffffu, assigned_any) && lane == 0)
            mark_tile_flags(a, max(strip * G::OUTC, 0), y0, G::OUTC, y1 - y0);
    }
    signal_neighbours(a);
}

}  // namespace kob
#endif  // KOB_FAST_CUH
