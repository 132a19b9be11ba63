// kob_fast2.cuh — TWO explicit-Euler sub-steps per launch (temporal blocking).
//
// Same arithmetic, operation for operation, as two launches of kob_step_fast (kob_fast.cuh) — the results are
// bit-identical — but phi and T cross HBM once per TWO sub-steps: the warp that marches down a column strip keeps a
// second set of register windows and feeds the rows of the first sub-step (phi^1, T^1, never stored) straight into the
// second.  Where the single-step kernel is HBM bound (sparse fields: the far-field shortcut leaves only the T diffusion)
// this halves the bytes per cell-update: 16 B per cell per LAUNCH = 8 B per cell-update.
//
//   level 1 (sub-step s)   : phi^0 row r (TMA stage)        -> pass 1 of row r-1, pass 2 of row r-2  = phi^1, T^1 (r-2)
//   level 2 (sub-step s+1) : phi^1 row r-2 (registers+SHFL) -> pass 1 of row r-3, pass 2 of row r-4  = phi^2, T^2 (r-4)
//
// Geometry: lane L owns cells x = xs + 2L, xs = 56*strip - 4.  Level 1 is valid from lane 0's second cell to lane 31's
// first, level 2 on lanes 2..29: 56 output columns per 64 computed; a job streams (rows + 8) phi^0 rows.  The composed
// stencil has radius 4, hence ghost depth 4 in the layout (kob_common.cuh).
// theta: a cell that is HELD in sub-step s but RE-ASSIGNED in s+1 would be overwritten while a neighbouring warp's halo
// lane still needs its old value, so this kernel reads `theta` and writes every non-zero final angle to `theta_next`
// (the buffers swap after the launch); the angle after sub-step s travels from level 1 to level 2 in registers.
// Noise: level 1 draws Philox at step s, level 2 at step s+1 (global coordinates, as everywhere).
#ifndef KOB_FAST2_CUH
#define KOB_FAST2_CUH

#include "kob_fast.cuh"

namespace kob {

constexpr int F2_OUTC = 56;     // output columns per strip (lanes 2..29)
constexpr int F2_HALO = 4;      // pass-1 columns left of the first output column (lanes 0, 1)
#ifndef KOB_F2_WARPS
#define KOB_F2_WARPS 8
#endif
#ifndef KOB_F2_NST
#define KOB_F2_NST 8
#endif
constexpr int F2_WARPS = KOB_F2_WARPS;
constexpr int F2_NST = KOB_F2_NST;   // TMA stages per warp (one CTA per SM: a deeper ring than the single-step kernel's)
#ifndef KOB_F2_RANGES
#define KOB_F2_RANGES 4
#endif
constexpr int F2_RANGES = KOB_F2_RANGES;    // row ranges per job handed to the general pass (power of two; 2 and 8 measured slower)
__host__ __device__ constexpr int f2_range_rows(int rows) { return (((rows + F2_RANGES - 1) / F2_RANGES) + 3) & ~3; }
constexpr int F2_BW = 72;       // TMA box width: columns xs-4 .. xs+67 (the box must start on a 16-byte boundary)
constexpr int F2_BOX_FLOATS = (FAST_RB * F2_BW + 31) / 32 * 32;
constexpr int F2_STAGE_FLOATS = 2 * F2_BOX_FLOATS;
constexpr int F2_WARP_BYTES = F2_NST * F2_STAGE_FLOATS * 4;

// Register windows of one level; "r" is the phi row this level consumes in the current iteration.
struct F2Level {
    float2 po0, po1;            // phi rows r-2, r-1
    float2 gx1, gx2, gy2;       // gx(r-1); gx, gy (r-2)
    float2 u1, lp1, lap2;       // u(r-1), c(r-1)+u(r-2), complete 9-point sum of row r-2
    float2 tq1, tu1, tlp1;      // T(r-2), u_T(r-2), c_T(r-2)+u_T(r-3)      [T lags phi by a row]
    float2 A2, A3, P2, P3, Q2;  // eps^2 (r-2, r-3), eps*eps'*gx (r-2, r-3), eps*eps'*gy (r-2)
    uint32_t nxa, nxb;          // Philox words of the odd row, drawn at the even row
    bool have_next;
    __device__ __forceinline__ void clear() {
        po0 = po1 = gx1 = gx2 = gy2 = u1 = lp1 = lap2 = tq1 = tu1 = tlp1 = A2 = A3 = P2 = P3 = Q2 = make_float2(0.f, 0.f);
        nxa = nxb = 0u; have_next = false;
    }
};

// Loop constants of the row update, broadcast to pairs once per job.
struct F2Const {
    float2 idx2, idy2, il2, ildt2, dtt2, K2, B02;
    float A0, e, pi;
};

// T-only row of one level (far field: phi == +0 in the whole footprint): rotates the T windows, returns T+ of row r-2.
__device__ __forceinline__ float2 f2_row_tonly(F2Level& S, const F2Const& C, float2 tn, float tw, float te) {
    const float2 two2 = f2(2.0f), m12 = f2(-12.0f);
    const float2 thsum = make_float2(tw + tn.y, tn.x + te);
    const float2 tu_new = f2fma(two2, tn, thsum);
    const float2 lapt = f2add(S.tlp1, tu_new);
    const float2 nt = f2fma(C.K2, f2(0.f), f2fma(lapt, C.ildt2, S.tq1));           // :215 with phi+ - phi = +0
    S.tlp1 = f2fma(two2, thsum, f2fma(m12, tn, S.tu1));
    S.tu1 = tu_new;
    S.tq1 = tn;
    return nt;
}

// One full row of one level.  Inputs: phi row r (own pair pn, west, east), T row r-1 (tn, tw, te), th_old = angle a cell
// of row r-1 keeps if the state machine holds it.  Outputs: phi+/T+ of row r-2, th_eff = angle of row r-1 after this
// sub-step, any_asg |= some cell of row r-1 re-assigned.  `pc2/pc3` = Philox counter words (step) of this level,
// `yglob` = global row of pass 2 (r-2), `even` = first row of a row pair (Philox sharing), `wrap_noise` = seam job.
template <int JM, bool NOISE, bool ROT, bool GEN>
__device__ __forceinline__ void f2_row(F2Level& S, const F2Const& C, const KParams<float>& P, const FastArgs& f, float2 pn,
                                       float w, float ee, float2 tn, float tw, float te, const float (&th_old_in)[2],
                                       uint32_t pc2, uint32_t pc3, int x, long long yglob, bool even, int lane, bool wrap_noise,
                                       int nx, long long nyg, float2& np_, float2& nt_, float (&th_eff)[2], bool& any_asg) {
    const float2 two2 = f2(2.0f), m12 = f2(-12.0f);
    const float e = C.e, pi = C.pi;
    const float A_w = __shfl_up_sync(0xffffffffu, S.A2.y, 1);
    const float A_e = __shfl_down_sync(0xffffffffu, S.A2.x, 1);
    const float Q_w = __shfl_up_sync(0xffffffffu, S.Q2.y, 1);
    const float Q_e = __shfl_down_sync(0xffffffffu, S.Q2.x, 1);
    // ---- phi row r: horizontal sums and x-gradient; T row r-1: horizontal sums ----
    const float2 hsum = make_float2(w + pn.y, pn.x + ee);
    const float2 gxn = f2mul(make_float2(pn.y - w, ee - pn.x), C.idx2);                       // :139
    const float2 thsum = make_float2(tw + tn.y, tn.x + te);
    // ---- pass 1 for row r-1, far-field values first ----
    const float2 gyn = f2mul(f2sub(pn, S.po0), C.idy2);                                       // :140
    float2 An = f2(C.A0), Pn = f2mul(C.B02, S.gx1), Qn = f2mul(C.B02, gyn);                   // cells holding theta = 0
    const float2 q = f2fma(f2neg(S.po0), S.po0, S.po0);                                       // phi (1 - phi) of row r-2
    float2 radd = f2(0.f);
    bool asg[2];
    asg[0] = (S.gx1.x < -e) || (fabsf(gyn.x) > e);                                            // :154-167: theta re-assigned
    asg[1] = (S.gx1.y < -e) || (fabsf(gyn.y) > e);
    bool interesting = asg[0] || asg[1] || q.x != 0.f || q.y != 0.f;
    if (GEN) interesting |= th_old_in[0] != 0.f || th_old_in[1] != 0.f;
    th_eff[0] = GEN ? th_old_in[0] : 0.f;
    th_eff[1] = GEN ? th_old_in[1] : 0.f;
    if (even) S.have_next = false;
    if (__any_sync(0xffffffffu, interesting)) {
        float th_old[2];
        th_old[0] = (GEN && !asg[0]) ? th_old_in[0] : 0.f;
        th_old[1] = (GEN && !asg[1]) ? th_old_in[1] : 0.f;
        const float2 gx = S.gx1, gy = gyn;
        const float2 agx = make_float2(fabsf(gx.x), fabsf(gx.y)), agy = make_float2(fabsf(gy.x), fabsf(gy.y));
        const float2 mn = make_float2(fminf(agx.x, agy.x), fminf(agx.y, agy.y));
        const float2 mx = make_float2(fmaxf(agx.x, agy.x), fmaxf(agx.y, agy.y));
        float2 r = atan01_2(f2mul(mn, make_float2(rcp_approx(mx.x), rcp_approx(mx.y))));
        const bool sw0 = agy.x > agx.x, sw1 = agy.y > agx.y;
        r = f2fma(r, make_float2(sw0 ? -1.0f : 1.0f, sw1 ? -1.0f : 1.0f), make_float2(sw0 ? HALF_PI_TRUE : 0.0f, sw1 ? HALF_PI_TRUE : 0.0f));
        r.x = __uint_as_float(__float_as_uint(r.x) ^ ((__float_as_uint(gx.x) ^ __float_as_uint(gy.x)) & 0x80000000u));
        r.y = __uint_as_float(__float_as_uint(r.y) ^ ((__float_as_uint(gx.y) ^ __float_as_uint(gy.y)) & 0x80000000u));
        float2 th2 = f2add(make_float2(gx.x < 0.f ? pi : (gy.x < 0.f ? f.two_pi : 0.0f), gx.y < 0.f ? pi : (gy.y < 0.f ? f.two_pi : 0.0f)), r);
        const float2 r2 = f2fma(gx, gx, f2mul(gy, gy));
        const float2 rinv = make_float2(rsqrt_approx(r2.x), rsqrt_approx(r2.y));
        float2 c1 = f2mul(gx, rinv), s1 = f2mul(gy, rinv);
        const bool fl0 = asg[0] && agx.x <= e, fl1 = asg[1] && agx.y <= e;                    // case A (:154-158)
        const bool rare = fl0 || fl1 || (GEN && (th_old[0] != 0.f || th_old[1] != 0.f));
        if (__any_sync(0xffffffffu, rare)) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const bool fl = k ? fl1 : fl0;
                const bool held = GEN && th_old[k] != 0.f;
                const float sg = (k ? gy.y : gy.x) < 0.f ? -1.0f : 1.0f;
                float t = th_old[k];
                if (GEN) t = t > 3.14159265358979f ? fmaf(-1.0f, 6.28318548202514648f, t) + 1.74845553e-7f : t;
                const float ct = GEN ? __cosf(t) : 0.f, st = GEN ? __sinf(t) : 0.f;
                float& thk = k ? th2.y : th2.x;
                float& ck = k ? c1.y : c1.x;
                float& sk = k ? s1.y : s1.x;
                thk = fl ? sg * f.half_pi : thk;
                ck = fl ? 0.0f : (held ? ct : ck);
                sk = fl ? sg : (held ? st : sk);
            }
        }
        float2 Cc = f2(1.0f), Ss = f2(0.0f);
        if (JM >= 0) {
            if (JM == 0) cpow2_rt(P.jmode, c1, s1, Cc, Ss); else cpow2<(JM > 0 ? JM : 1)>(c1, s1, Cc, Ss);
            if (ROT) {
                const float2 c2 = f2fma(Cc, f2(f.cj0), f2mul(Ss, f2(f.sj0)));
                const float2 s2 = f2fma(Ss, f2(f.cj0), f2neg(f2mul(Cc, f2(f.sj0))));
                Cc = c2; Ss = s2;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const float th = asg[k] ? (k ? th2.y : th2.x) : th_old[k];
                if (asg[k] || th != 0.f) {
                    float Ck, Sk;
                    fast_sincos(P.aniso * (th - P.theta0), &Sk, &Ck);
                    if (k) { Cc.y = Ck; Ss.y = Sk; } else { Cc.x = Ck; Ss.x = Sk; }
                }
            }
        }
        th_eff[0] = asg[0] ? th2.x : th_eff[0];
        th_eff[1] = asg[1] ? th2.y : th_eff[1];
        any_asg |= asg[0] || asg[1];
        float2 ep = f2fma(f2(f.ebd), Cc, f2(P.epsbar));                                       // :170
        float2 ed = f2mul(f2(P.neg_ebjd), Ss);                                                // :171
        const bool d0 = !asg[0] && !(GEN && th_old[0] != 0.f), d1 = !asg[1] && !(GEN && th_old[1] != 0.f);
        ep = make_float2(d0 ? f.eps0 : ep.x, d1 ? f.eps0 : ep.y);
        ed = make_float2(d0 ? f.epsd0 : ed.x, d1 ? f.epsd0 : ed.y);
        An = f2mul(ep, ep);
        const float2 B = f2mul(ep, ed);
        Pn = f2mul(B, S.gx1);
        Qn = f2mul(B, gyn);
        // ---- reaction term q*((phi - 1/2) + m(T)) [+ noise] of row r-2, :206-214 ----
        float2 rq = f2(0.f);
        if (NOISE) {
            const bool hi = (x & 2) != 0;
            uint32_t wa, wb;
            if (wrap_noise) {
                // seam job: level 1 also updates cells it does not own (ghost columns / rows of the torus); their draw must
                // be the owner's, i.e. keyed on the WRAPPED global cell — one Philox block per cell, no sharing
                long long yw = yglob;
                yw = yw < 0 ? yw + nyg : (yw >= nyg ? yw - nyg : yw);
                uint32_t wk[2];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    int xw = x + k;
                    xw = xw < 0 ? xw + nx : (xw >= nx ? xw - nx : xw);
                    const Philox4 ph = fast_philox(f, (uint32_t)xw >> 2, (uint32_t)yw, pc2, pc3);
                    const uint32_t i3 = (uint32_t)xw & 3u;
                    wk[k] = i3 == 0u ? ph.w[0] : (i3 == 1u ? ph.w[1] : (i3 == 2u ? ph.w[2] : ph.w[3]));
                }
                wa = wk[0]; wb = wk[1];
            } else if (even) {
                const Philox4 ph = fast_philox(f, (uint32_t)x >> 2, (uint32_t)yglob + (hi ? 1u : 0u), pc2, pc3);
                const int partner = hi ? lane - 1 : lane + 1;
                const uint32_t ra = __shfl_sync(0xffffffffu, hi ? ph.w[0] : ph.w[2], partner);
                const uint32_t rb = __shfl_sync(0xffffffffu, hi ? ph.w[1] : ph.w[3], partner);
                wa = hi ? ra : ph.w[0]; wb = hi ? rb : ph.w[1];
                S.nxa = hi ? ph.w[2] : ra; S.nxb = hi ? ph.w[3] : rb;
                S.have_next = true;
            } else if (S.have_next) {
                wa = S.nxa; wb = S.nxb;
            } else {
                const Philox4 ph = fast_philox(f, (uint32_t)x >> 2, (uint32_t)yglob, pc2, pc3);
                wa = hi ? ph.w[2] : ph.w[0]; wb = hi ? ph.w[3] : ph.w[1];
            }
            rq = f2fma(make_float2((float)(wa >> 8), (float)(wb >> 8)), f2(5.9604644775390625e-8f), f2(-0.5f));
        }
        const float2 xa = f2mul(f2(P.gamma), f2sub(f2(P.teq), S.tq1));
        const float ax0 = fabsf(xa.x), ax1 = fabsf(xa.y);
        const bool b0 = ax0 > 1.0f, b1 = ax1 > 1.0f;
        const float2 ra_ = atan01_2(make_float2(b0 ? rcp_approx(ax0) : ax0, b1 ? rcp_approx(ax1) : ax1));
        float2 m = f2fma(ra_, make_float2(b0 ? -P.alpha_over_pi : P.alpha_over_pi, b1 ? -P.alpha_over_pi : P.alpha_over_pi),
                         make_float2(b0 ? f.m_off : 0.0f, b1 ? f.m_off : 0.0f));
        m.x = __uint_as_float(__float_as_uint(m.x) ^ (__float_as_uint(xa.x) & 0x80000000u));
        m.y = __uint_as_float(__float_as_uint(m.y) ^ (__float_as_uint(xa.y) & 0x80000000u));
        float2 rv = f2mul(q, f2add(f2sub(S.po0, f2(0.5f)), m));                               // :214
        if (NOISE) rv = f2fma(f2mul(f2(P.noise_a), q), rq, rv);
        radd = rv;
    }
    // ---- pass 2 for row r-2 ----
    const float2 dA = make_float2(S.A2.y - A_w, A_e - S.A2.x);                                // :190-192
    const float2 dQ = make_float2(Q_w - S.Q2.y, S.Q2.x - Q_e);                                // term2, :201-203
    const float2 gEx = f2mul(dA, C.idx2);
    const float2 gEy = f2mul(f2sub(An, S.A3), C.idy2);                                        // :193-195
    float2 sm = f2fma(f2sub(Pn, S.P3), C.idy2, radd);                                         // term1 (:197-199) + reaction
    sm = f2fma(dQ, C.idx2, sm);
    sm = f2fma(S.A2, f2mul(S.lap2, C.il2), sm);                                               // eps^2 * lap(phi)
    sm = f2fma(gEx, S.gx2, sm);                                                               // term3, :204
    sm = f2fma(gEy, S.gy2, sm);
    np_ = f2fma(sm, C.dtt2, S.po0);                                                           // :211
    const float2 tu_new = f2fma(two2, tn, thsum);                                             // u_T(r-1)
    const float2 lapt = f2add(S.tlp1, tu_new);                                                // 9-point sum of T at row r-2
    nt_ = f2fma(C.K2, f2sub(np_, S.po0), f2fma(lapt, C.ildt2, S.tq1));                        // :215
    // ---- rotate the windows ----
    S.tlp1 = f2fma(two2, thsum, f2fma(m12, tn, S.tu1));
    S.tu1 = tu_new;
    S.tq1 = tn;
    const float2 u_new = f2fma(two2, pn, hsum);
    S.lap2 = f2add(S.lp1, u_new);
    S.lp1 = f2fma(two2, hsum, f2fma(m12, pn, S.u1));
    S.u1 = u_new;
    S.gx2 = S.gx1; S.gy2 = gyn; S.gx1 = gxn;
    S.po0 = S.po1; S.po1 = pn;
    S.A3 = S.A2; S.A2 = An;
    S.P3 = S.P2; S.P2 = Pn;
    S.Q2 = Qn;
}

// ---- far pass --------------------------------------------------------------------------------------------------
// The two-step kernel needs ~230 registers (two sets of windows + the data-dependent block), i.e. 8 warps per SM — too
// few to keep HBM busy on the far-field rows, which need almost none of that state.  So a launch pair is used:
//   1. kob_far2 (this kernel, ~48 registers, 24 warps per SM) visits EVERY job.  Interior jobs whose theta flags are
//      clear are streamed (seam jobs too, with alias stores); every row goes through two T-diffusion sub-steps (level 2
//      from level 1's rows in registers) — the same instructions as kob_step_fast2's shortcut, bit for bit — and is
//      STORED where the 12 rows up to it are all +0 in phi (and, under a set theta flag, 0 in theta).  A job is cut
//      into F2_RANGES row ranges; ranges with unstored rows are appended to the work list (rows both passes store get the same
//      bits twice).
//   2. kob_step_fast2 then processes the work list.
struct Far2Args {
    int* list;                  // job ids for the general pass
    unsigned int* list_count;   // entries appended by this launch (zeroed by the host before it)
};

constexpr int FAR2_WARPS = 8;
constexpr int FAR2_NST = 4;
constexpr int FAR2_PBW = 64;    // phi box: exactly the 64 own columns (only looked at: are they all +0?)
constexpr int FAR2_PBOX_FLOATS = FAST_RB * FAR2_PBW;                       // 1024 B, a multiple of 128
constexpr int FAR2_STAGE_FLOATS = FAR2_PBOX_FLOATS + F2_BOX_FLOATS;       // phi box (64 wide) + T box (72 wide)
constexpr int FAR2_WARP_BYTES = FAR2_NST * FAR2_STAGE_FLOATS * 4;

__global__ void __launch_bounds__(32 * FAR2_WARPS, 3) kob_far2(const __grid_constant__ FastMaps maps, const StepArgs<float> a,
                                                              const FastArgs f, const Far2Args w) {
    constexpr int BW = F2_BW, RB = FAST_RB, NST = FAR2_NST;
    constexpr int STAGE_FLOATS = FAR2_STAGE_FLOATS, PBW = FAR2_PBW, PBOX_FLOATS = FAR2_PBOX_FLOATS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    float* stages = reinterpret_cast<float*>(smem_raw) + (size_t)warp * NST * STAGE_FLOATS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)nwarps * FAR2_WARP_BYTES) + warp * NST;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NST; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const CUtensorMap* map_phi = a.cur ? &maps.phi[1] : &maps.phi[0];
    const CUtensorMap* map_t = a.cur ? &maps.t[1] : &maps.t[0];
    float* __restrict__ phi_out = a.self.phi[a.cur ^ 1];
    float* __restrict__ t_out = a.self.t[a.cur ^ 1];
    const long long pitch = a.pitch;
    F2Const C;
    C.idx2 = C.idy2 = C.il2 = C.dtt2 = C.B02 = f2(0.f);
    C.ildt2 = f2(f.il_dt); C.K2 = f2(a.prm.K);
    C.A0 = 0.f; C.e = 0.f; C.pi = 0.f;
    unsigned int gchunk = 0;
    __shared__ unsigned long long s_job;
    const int nsp = f.cta_jobs ? f.nstrips_p : f.nstrips;     // strips per segment in the job numbering
    const int njobs_q = nsp * f.nseg;
    for (;;) {
        unsigned long long jraw = 0;
        if (f.cta_jobs) {                                    // a CTA claims 8 adjacent strips and keeps them in lock-step:
            __syncthreads();                                 // a grid row is then fetched as 8 x 224 contiguous bytes
            if (threadIdx.x == 0) s_job = atomicAdd(f.job_ctr, (unsigned long long)nwarps) - f.job_base;
            __syncthreads();
            if (s_job >= (unsigned long long)njobs_q) break;
            jraw = s_job + (unsigned long long)warp;
        } else {
            if (lane == 0) jraw = atomicAdd(f.job_ctr, 1ull) - f.job_base;
            jraw = __shfl_sync(0xffffffffu, jraw, 0);
            if (jraw >= (unsigned long long)njobs_q) break;
        }
        const int job = (int)jraw;
        const int strip = job - (job / nsp) * nsp;
        const int sq = job / nsp;
        const int seg_ = sq == 0 ? 0 : (sq == 1 ? f.nseg - 1 : sq - 1);
        const int y0 = seg_ < f.nseg_a ? seg_ * f.yj : f.nseg_a * f.yj + (seg_ - f.nseg_a) * f.yj_b;
        const int y1 = min(y0 + (seg_ < f.nseg_a ? f.yj : f.yj_b), a.ny);
        if (strip >= f.nstrips) {                        // padding warp of a CTA job: keep the barriers company
            const int nchp = ((y1 - y0) + 8 + RB - 1) / RB;
            for (int c = 0; c < nchp; ++c) __syncthreads();
            continue;
        }
        const int xs = strip * F2_OUTC - F2_HALO;
        const int x = xs + 2 * lane;
        // seam jobs store to the aliases as well (ghost columns / neighbour strips' ghost rows) and skip the ragged edge
        const bool seam = strip == 0 || (strip + 1) * F2_OUTC > a.nx - GXR || y0 < GY || y1 > a.ny - GY;
        if (a.linked) {
            if (lane == 0) {
                if (y0 < GY + 2) wait_flag(&a.self.arrive[0], a.epoch, &a.self.arrive[2]);
                if (y1 > a.ny - GY - 1) wait_flag(&a.self.arrive[1], a.epoch, &a.self.arrive[2]);
            }
            __syncwarp();
        }
        uint32_t need_out = 0u;                          // row ranges of this job that the general pass has to do
        bool live;                                       // some theta flag under the footprint is set -> look at theta itself
        {
            const int fby0 = max((y0 - 3 + GY) / FBY, 0);
            const int bx0 = max((xs + GX) / FBX, 0), bx1 = min((xs + GX + 63) / FBX, a.nfbx - 1);
            const int by1 = min((y1 + 2 + GY) / FBY, a.nfby - 1);
            const int nbx = bx1 - bx0 + 1, nby = by1 - fby0 + 1;
            uint32_t fl = 0;
            for (int i = lane; i < nby * nbx; i += 32) fl |= __ldcg(&a.self.tflags[(fby0 + i / nbx) * a.nfbx + bx0 + i % nbx]);
            live = __any_sync(0xffffffffu, fl != 0u);
        }
        const int nrows = (y1 - y0) + 8;                 // streamed phi^0 rows y0-4 .. y1+3
        const int nch = (nrows + RB - 1) / RB;
        int issued = 0;
        {
            const int box_x = xs - 4 + GX;
            auto issue = [&](int c) {
                const unsigned int gi = gchunk + (unsigned int)c;
                const int st = gi % NST;
                float* dst = stages + st * STAGE_FLOATS;
                mbar_expect_tx(&bars[st], RB * (PBW + BW) * 4);
                const int yr = y0 - 4 + c * RB + GY;
                tma_load_2d(dst, map_phi, box_x + 4, yr, &bars[st]);                 // phi: columns xs .. xs+63
                tma_load_2d(dst + PBOX_FLOATS, map_t, box_x, yr - 1, &bars[st]);     // T:   columns xs-4 .. xs+67
            };
            issued = min(NST, nch);
            if (lane == 0)
                for (int c = 0; c < issued; ++c) issue(c);
            F2Level L1, L2;
            L1.clear(); L2.clear();
            float2 t1_prev = f2(0.f);
            const long long o4 = pidx<float>(pitch, x, y0 - 8);
            float* pphi = phi_out + o4;
            float* ptt = t_out + o4;
            const unsigned int nvalid = (unsigned int)(y1 - y0);
            const unsigned int nstore = (lane >= 2 && lane <= 29) ? nvalid : 0u;
            const int q = f2_range_rows((int)nvalid);                // rows per row range (kob_step_fast2 decodes the same way)
            bool z1 = false, z2 = false;                             // the previous two chunks were all zero
            uint32_t need = 0u;                                      // row ranges (F2_RANGES per job) that need the general pass
            for (int c = 0; c < nch; ++c) {
                if (f.cta_jobs) __syncthreads();
                const unsigned int gi = gchunk + (unsigned int)c;
                const int st = gi % NST;
                mbar_wait(&bars[st], (gi / NST) & 1u);
                const float* sp = stages + st * STAGE_FLOATS + 2 * lane;           // this lane's own phi cells
                const float* stt = stages + st * STAGE_FLOATS + PBOX_FLOATS + 2 * lane + 4;   // ... and T cells (one row behind)
                uint32_t bits = 0u;
#pragma unroll
                for (int rr = 0; rr < RB; ++rr) bits |= __float_as_uint(sp[rr * PBW]) | __float_as_uint(sp[rr * PBW + 1]);
                if (live) {                              // the flags are coarse (128 x 32 blocks): check the angles of these rows
                    const int yr = y0 - 4 + c * RB + GY; // padded row of the chunk's first row
#pragma unroll
                    for (int rr = 0; rr < RB; ++rr)
                        if (yr + rr < a.ny + 2 * GY && x >= -GX && x + 1 < a.nx + GXR) {
                            const float2 v = __ldg(reinterpret_cast<const float2*>(a.self.theta + (long long)(yr + rr) * pitch + (x + GX)));
                            bits |= __float_as_uint(v.x) | __float_as_uint(v.y);
                        }
                }
                // Rows are always pushed through the T windows (they only see T^0, which is real data); a row's result is
                // STORED only if this chunk and the two before it are all zero — then phi^0 == +0 (and theta == 0) over the
                // 9 rows the two composed sub-steps look at, and T^1 in level 2's window carries no missing K (phi^1 - phi^0)
                // term.  Row ranges with unstored rows go to the general pass.
                const bool z0 = !__any_sync(0xffffffffu, bits != 0u);
                const bool ok = z0 && z1 && z2;
                z2 = z1; z1 = z0;
                const int yrel0 = c * RB - 8;
                if (!ok) {
                    const int ra = max(yrel0, 0), rb = min(yrel0 + RB - 1, (int)nvalid - 1);
                    if (ra <= rb) need |= (1u << (ra / q)) | (1u << (rb / q));
                }
#pragma unroll
                for (int rr = 0; rr < RB; ++rr) {
                    const unsigned int yrel = (unsigned int)(yrel0 + rr);
                    const float* row = stt + rr * BW;
                    const float2 tn = *reinterpret_cast<const float2*>(row);
                    const float2 t1 = f2_row_tonly(L1, C, tn, row[-1], row[2]);          // T^1 of row r-2
                    const float tw2 = __shfl_up_sync(0xffffffffu, t1_prev.y, 1);
                    const float te2 = __shfl_down_sync(0xffffffffu, t1_prev.x, 1);
                    const float2 t2 = f2_row_tonly(L2, C, t1_prev, tw2, te2);            // T^2 of row r-4
                    t1_prev = t1;
                    if (ok && yrel < nstore) {
                        if (!seam) {
                            *reinterpret_cast<float2*>(pphi) = f2(0.f);
                            *reinterpret_cast<float2*>(ptt) = t2;
                        } else {
                            const int y = y0 + (int)yrel;
#pragma unroll
                            for (int k = 0; k < 2; ++k) {
                                if (x + k < a.nx) {
                                    const float vt = k ? t2.y : t2.x;
                                    if (y < GY || y >= a.ny - GY) {
                                        fast_store_edge(phi_out, a.lower.phi[a.cur ^ 1], a.upper.phi[a.cur ^ 1], pitch, a.nx, a.ny, a.lower.ny, x + k, y, 0.f);
                                        fast_store_edge(t_out, a.lower.t[a.cur ^ 1], a.upper.t[a.cur ^ 1], pitch, a.nx, a.ny, a.lower.ny, x + k, y, vt);
                                    } else {
                                        pphi[k] = 0.f; ptt[k] = vt;
                                        if (x + k < GXR) { pphi[k + a.nx] = 0.f; ptt[k + a.nx] = vt; }
                                        if (x + k >= a.nx - GXR) { pphi[k - a.nx] = 0.f; ptt[k - a.nx] = vt; }
                                    }
                                }
                            }
                        }
                    }
                    pphi += pitch; ptt += pitch;
                }
                __syncwarp();
                if (c + NST < nch) {
                    if (lane == 0) issue(c + NST);
                    issued = c + NST + 1;
                }
            }
            __syncwarp();
            gchunk += (unsigned int)issued;
            need_out = need;
        }
        if (need_out) {                                  // append the row ranges that were not fully stored
            unsigned int pos = 0;
            if (lane == 0) pos = atomicAdd(w.list_count, (unsigned int)__popc(need_out));
            pos = __shfl_sync(0xffffffffu, pos, 0);
            if (lane < F2_RANGES && ((need_out >> lane) & 1u))
                w.list[pos + __popc(need_out & ((1u << lane) - 1u))] = (sq * f.nstrips + strip) * F2_RANGES + lane;   // unpadded job numbering
        }
    }
}

template <int JM, bool NOISE, bool ROT>
__global__ void __launch_bounds__(32 * F2_WARPS, 1) kob_step_fast2(const __grid_constant__ FastMaps maps, const StepArgs<float> a,
                                                                  const FastArgs f) {
    constexpr int BW = F2_BW, RB = FAST_RB, NST = F2_NST;
    constexpr int STAGE_FLOATS = F2_STAGE_FLOATS, BOX_FLOATS = F2_BOX_FLOATS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    float* stages = reinterpret_cast<float*>(smem_raw) + (size_t)warp * NST * STAGE_FLOATS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)nwarps * F2_WARP_BYTES) + warp * NST;
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
    F2Const C;
    C.idx2 = f2(P.inv_dx); C.idy2 = f2(P.inv_dy); C.il2 = f2(P.inv_lapden); C.ildt2 = f2(f.il_dt);
    C.dtt2 = f2(P.dt_over_tau); C.K2 = f2(P.K); C.B02 = f2(f.eps0 * f.epsd0);
    C.A0 = f.eps0 * f.eps0; C.e = REF_DEADBAND; C.pi = REF_PI_F;
    const uint32_t pc2b = (uint32_t)(a.step + 1ull), pc3b = (uint32_t)((a.step + 1ull) >> 32);   // level 2 = step + 1
    unsigned int gchunk = 0;

    __shared__ unsigned long long s_job;
    const int nsp = f.cta_jobs ? f.nstrips_p : f.nstrips;     // strips per segment in the job numbering
    const int njobs_q = nsp * f.nseg;
    for (;;) {
        unsigned long long jraw = 0;
        int sub = -1;
        if (f.cta_jobs) {                                    // a CTA claims 8 adjacent strips of one segment
            __syncthreads();
            if (threadIdx.x == 0) s_job = atomicAdd(f.job_ctr, (unsigned long long)nwarps) - f.job_base;
            __syncthreads();
            if (s_job >= (unsigned long long)njobs_q) break;
            jraw = s_job + (unsigned long long)warp;
        } else if (f.list) {                                 // general pass of a far/general launch pair: the work list
            unsigned int idx = 0;
            if (lane == 0) idx = atomicAdd(f.list_claim, 1u);
            idx = __shfl_sync(0xffffffffu, idx, 0);
            if (idx >= *f.list_count) break;
            jraw = (unsigned long long)f.list[idx];
            sub = (int)(jraw & (unsigned long long)(F2_RANGES - 1));   // the far pass cuts a job into F2_RANGES row ranges:
            jraw /= F2_RANGES;                               // short jobs keep the (latency-bound) general pass short
        } else {                                             // every warp claims its own job: no barrier anywhere
            if (lane == 0) jraw = atomicAdd(f.job_ctr, 1ull) - f.job_base;
            jraw = __shfl_sync(0xffffffffu, jraw, 0);
            if (jraw >= (unsigned long long)njobs_q) break;
        }
        const int job = (int)jraw;
        const int strip = job - (job / nsp) * nsp;
        const int sq = job / nsp;
        const int seg_ = sq == 0 ? 0 : (sq == 1 ? f.nseg - 1 : sq - 1);     // the two torus-seam segments first
        int y0 = seg_ < f.nseg_a ? seg_ * f.yj : f.nseg_a * f.yj + (seg_ - f.nseg_a) * f.yj_b;
        int y1 = min(y0 + (seg_ < f.nseg_a ? f.yj : f.yj_b), a.ny);
        if (sub >= 0) {
            const int q = f2_range_rows(y1 - y0);            // rows per sub-job, a multiple of 4
            y0 += sub * q;
            y1 = min(y0 + q, y1);
            if (y0 >= y1) continue;
        }
        const int xs = strip * F2_OUTC - F2_HALO;        // first pass-1 column of the warp
        const int x = xs + 2 * lane;                     // first cell of this lane
        const bool out_lane = lane >= 2 && lane <= 29;
        const bool real_job = strip < f.nstrips;

        if (a.linked && real_job) {
            if (lane == 0) {
                if (y0 < GY + 2) wait_flag(&a.self.arrive[0], a.epoch, &a.self.arrive[2]);
                if (y1 > a.ny - GY - 1) wait_flag(&a.self.arrive[1], a.epoch, &a.self.arrive[2]);
            }
            __syncwarp();
        }
        // theta may be non-zero somewhere in the pass-1 footprint (rows y0-3 .. y1+2)?  One bit per 32-row flag row.
        uint32_t livemask = 0;
        const int fby0 = max((y0 - 3 + GY) / FBY, 0);
        if (real_job) {
            const int bx0 = max((xs + GX) / FBX, 0), bx1 = min((xs + GX + 63) / FBX, a.nfbx - 1);
            const int by1 = min((y1 + 2 + GY) / FBY, a.nfby - 1);
            const int nbx = bx1 - bx0 + 1, nby = by1 - fby0 + 1;
            for (int i0 = 0; i0 < nby; i0 += 32) {
                uint32_t fl = 0;
                if (i0 + lane < nby)
                    for (int bx = 0; bx < nbx; ++bx) fl |= __ldcg(&a.self.tflags[(fby0 + i0 + lane) * a.nfbx + bx0 + bx]);
                const uint32_t m = __ballot_sync(0xffffffffu, fl != 0u);
                livemask |= i0 == 0 ? m : (m ? 0x80000000u : 0u);
            }
        }
        const bool live = livemask != 0u;
        const bool seam = strip == 0 || (strip + 1) * F2_OUTC > a.nx - GXR || y0 < GY || y1 > a.ny - GY;
        const bool lock = f.cta_jobs == 2 && !__syncthreads_or((live || seam) && real_job);   // far-field CTA jobs advance in lock-step

        const int nrows = (y1 - y0) + 8;                 // streamed phi^0 rows y0-4 .. y1+3
        const int nch = (nrows + RB - 1) / RB;
        if (!real_job) {
            if (lock)
                for (int c = 0; c < nch; ++c) __syncthreads();
            continue;
        }
        const int box_x = xs - 4 + GX;                   // padded x of box column 0 (a multiple of 4 elements)
        auto issue = [&](int c) {
            const unsigned int gi = gchunk + (unsigned int)c;
            const int st = gi % NST;
            float* dst = stages + st * STAGE_FLOATS;
            mbar_expect_tx(&bars[st], 2 * RB * BW * 4);
            const int yr = y0 - 4 + c * RB + GY;         // padded row of the chunk's first phi row
            tma_load_2d(dst, map_phi, box_x, yr, &bars[st]);
            tma_load_2d(dst + BOX_FLOATS, map_t, box_x, yr - 1, &bars[st]);
        };
        if (lane == 0) {
            for (int c = 0; c < NST && c < nch; ++c) issue(c);
        }

        bool assigned_any = false;
        auto body = [&](auto gen_tag) {
            constexpr bool GEN = decltype(gen_tag)::value;      // true: held theta is read, ragged edge + alias stores
            F2Level L1, L2;
            L1.clear(); L2.clear();
            float2 t1_prev = f2(0.f);                    // T^1 of row r-3 (level 2's T input lags its phi input by a row)
            float thp0[2] = {0.f, 0.f}, thp1[2] = {0.f, 0.f};      // theta^0 of rows r-1, r (held cells of level 1)
            float e1[2] = {0.f, 0.f}, e2[2] = {0.f, 0.f};          // angle after level 1 of rows r-2, r-3
            const long long o4 = pidx<float>(pitch, x, y0 - 8);    // cell (x, r-4) at the first iteration
            float* pphi = phi_out + o4;
            float* ptt = t_out + o4;
            float* pthn = a.self.theta_next + (o4 + pitch);        // theta_next of cell (x, r-3)
            const float* pth0 = a.self.theta + (o4 + 5 * pitch);   // theta^0 of cell (x, r+1): prefetch target
            const unsigned int nvalid = (unsigned int)(y1 - y0);
            const unsigned int nstore = out_lane ? nvalid : 0u;
            bool pz1 = false, pz2 = false;               // the previous / second previous chunk's phi^0 rows were all +0

            for (int c = 0; c < nch; ++c) {
                const unsigned int gi = gchunk + (unsigned int)c;
                const int st = gi % NST;
                if (lock) __syncthreads();
                mbar_wait(&bars[st], (gi / NST) & 1u);
                const float* sp = stages + st * STAGE_FLOATS + 2 * lane + 4;       // this lane's own phi cells
                const float* stt = sp + BOX_FLOATS;                                // T rows (one row behind)
                const int yrel0 = c * RB - 8;                                      // (r - 4) - y0 for rr = 0
                if (!GEN || !seam) {
                    // far field: phi^0 == +0 (and, where theta flags are live, theta^0 == 0) on the 64 own columns of this
                    // chunk and of the two chunks before it -> both levels only diffuse T (level 2 from level 1's rows in
                    // registers); bit-identical to the full path.  Seam jobs (alias stores) always take the full path.
                    uint32_t bits = 0u;
#pragma unroll
                    for (int rr = 0; rr < RB; ++rr) bits |= __float_as_uint(sp[rr * BW]) | __float_as_uint(sp[rr * BW + 1]);
                    if (GEN) {
                        const int yr = y0 - 4 + c * RB + GY;                         // padded row of the chunk's first row
#pragma unroll
                        for (int rr = 0; rr < RB; ++rr)
                            if (yr + rr < a.ny + 2 * GY) {
                                const float2 v = __ldg(reinterpret_cast<const float2*>(a.self.theta + (long long)(yr + rr) * pitch + (x + GX)));
                                bits |= __float_as_uint(v.x) | __float_as_uint(v.y);
                            }
                    }
                    const bool curz = !__any_sync(0xffffffffu, bits != 0u);
                    const bool skip = curz && pz1 && pz2 && !f.no_skip;
                    pz2 = pz1; pz1 = curz;
                    if (skip) {
#pragma unroll
                        for (int rr = 0; rr < RB; ++rr) {
                            const unsigned int yrel = (unsigned int)(yrel0 + rr);
                            const float* row = stt + rr * BW;
                            const float2 tn = *reinterpret_cast<const float2*>(row);
                            const float2 t1 = f2_row_tonly(L1, C, tn, row[-1], row[2]);          // T^1 of row r-2
                            const float tw2 = __shfl_up_sync(0xffffffffu, t1_prev.y, 1);
                            const float te2 = __shfl_down_sync(0xffffffffu, t1_prev.x, 1);
                            const float2 t2 = f2_row_tonly(L2, C, t1_prev, tw2, te2);            // T^2 of row r-4
                            t1_prev = t1;
                            if (yrel < nstore) {
                                *reinterpret_cast<float2*>(pphi) = f2(0.f);
                                *reinterpret_cast<float2*>(ptt) = t2;
                            }
                            pphi += pitch; ptt += pitch; pthn += pitch; pth0 += pitch;
                        }
                        if (GEN) {
                            // the angle pipelines skipped 4 rows that hold theta == 0; re-prime them for the next chunk's
                            // first row r': theta^0(r'-1) = 0 (this chunk), theta^0(r') is loaded now
                            e1[0] = e1[1] = e2[0] = e2[1] = 0.f;
                            thp0[0] = thp0[1] = 0.f;
                            thp1[0] = thp1[1] = 0.f;
                            if ((unsigned int)(yrel0 + RB) + 7u <= nvalid + 5u) {                // row r' <= y1 + 2
                                const float2 v = __ldg(reinterpret_cast<const float2*>(pth0 - pitch));
                                thp1[0] = v.x; thp1[1] = v.y;
                            }
                        }
                        __syncwarp();
                        if (lane == 0 && c + NST < nch) issue(c + NST);
                        continue;
                    }
                }
                bool lrow_c = false;          // GEN: some theta-flag row under this chunk's prefetch rows (r+1) is live
                if (GEN && live) {
                    const int f0 = ((y0 + yrel0 + 5 + GY) >> 5) - fby0, f1 = ((y0 + yrel0 + RB + 4 + GY) >> 5) - fby0;   // FBY == 32
                    lrow_c = ((livemask >> min(max(f0, 0), 31)) | (livemask >> min(max(f1, 0), 31))) & 1u;
                }
#pragma unroll
                for (int rr = 0; rr < RB; ++rr) {
                    const unsigned int yrel = (unsigned int)(yrel0 + rr);           // level-2 pass-2 row (r-4), relative to y0
                    const bool even = (rr & 1) == 0;
                    // ---- level 1: sub-step s ----
                    const float* row = sp + rr * BW;
                    const float* trow = stt + rr * BW;
                    const float2 pn = *reinterpret_cast<const float2*>(row);
                    const float2 tn = *reinterpret_cast<const float2*>(trow);
                    float2 p1, t1;
                    float th1[2];
                    bool dummy = false;
                    f2_row<JM, NOISE, ROT, GEN>(L1, C, P, f, pn, row[-1], row[2], tn, trow[-1], trow[2], thp0, f.pc2, f.pc3, x,
                                                a.y0 + y0 + (int)yrel + 2, even, lane, GEN && seam, a.nx, f.ny_global, p1, t1, th1, dummy);
                    // ---- level 2: sub-step s+1 on phi^1 row r-2, T^1 row r-3 ----
                    const float pw = __shfl_up_sync(0xffffffffu, p1.y, 1);
                    const float pe = __shfl_down_sync(0xffffffffu, p1.x, 1);
                    const float tw2 = __shfl_up_sync(0xffffffffu, t1_prev.y, 1);
                    const float te2 = __shfl_down_sync(0xffffffffu, t1_prev.x, 1);
                    float2 p2, t2;
                    float th2f[2];
                    bool asg2 = false;
                    f2_row<JM, NOISE, ROT, true>(L2, C, P, f, p1, pw, pe, t1_prev, tw2, te2, e2, pc2b, pc3b, x,
                                                 a.y0 + y0 + (int)yrel, even, lane, GEN && seam, a.nx, f.ny_global, p2, t2, th2f, asg2);
                    t1_prev = t1;
                    // ---- stores: phi^2, T^2 of row r-4; final angle of row r-3 ----
                    if (yrel < nstore) {
                        if (!GEN) {
                            *reinterpret_cast<float2*>(pphi) = p2;
                            *reinterpret_cast<float2*>(ptt) = t2;
                        } else {
                            const int y = y0 + (int)yrel;
#pragma unroll
                            for (int k = 0; k < 2; ++k) {
                                if (x + k < a.nx) {
                                    const float vp = k ? p2.y : p2.x, vt = k ? t2.y : t2.x;
                                    if (y < GY || y >= a.ny - GY) {          // rows on the strip seam: every alias
                                        fast_store_edge(phi_out, a.lower.phi[a.cur ^ 1], a.upper.phi[a.cur ^ 1], pitch, a.nx, a.ny, a.lower.ny, x + k, y, vp);
                                        fast_store_edge(t_out, a.lower.t[a.cur ^ 1], a.upper.t[a.cur ^ 1], pitch, a.nx, a.ny, a.lower.ny, x + k, y, vt);
                                    } else {
                                        pphi[k] = vp; ptt[k] = vt;
                                        if (x + k < GXR) { pphi[k + a.nx] = vp; ptt[k + a.nx] = vt; }
                                        if (x + k >= a.nx - GXR) { pphi[k - a.nx] = vp; ptt[k - a.nx] = vt; }
                                    }
                                }
                            }
                        }
                    }
                    if (yrel + 1u < nstore) {                               // row r-3 is owned
                        const int y = y0 + (int)yrel + 1;
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const float th = th2f[k];
                            if (th != 0.f && (!GEN || x + k < a.nx)) {
                                if (GEN && (y < GY || y >= a.ny - GY))
                                    fast_store_edge(a.self.theta_next, a.lower.theta_next, a.upper.theta_next, pitch, a.nx, a.ny, a.lower.ny, x + k, y, th);
                                else {
                                    pthn[k] = th;
                                    if (GEN) {
                                        if (x + k < GXR) pthn[k + a.nx] = th;
                                        if (x + k >= a.nx - GXR) pthn[k - a.nx] = th;
                                    }
                                }
                                assigned_any = true;
                            }
                        }
                    }
                    // ---- angle pipeline level 1 -> level 2, theta^0 prefetch ----
                    e2[0] = e1[0]; e2[1] = e1[1];
                    e1[0] = th1[0]; e1[1] = th1[1];
                    if (GEN) {
                        thp0[0] = thp1[0]; thp0[1] = thp1[1];
                        thp1[0] = thp1[1] = 0.f;
                        if (lrow_c && yrel + 8u <= nvalid + 5u) {                     // theta^0 rows y0-3 .. y1+2
#pragma unroll
                            for (int k = 0; k < 2; ++k)
                                if (x + k < a.nx + GXR && x + k >= -GXR) thp1[k] = __ldg(pth0 + k);
                        }
                    }
                    pphi += pitch; ptt += pitch; pthn += pitch; pth0 += pitch;
                }
                __syncwarp();
                if (lane == 0 && c + NST < nch) issue(c + NST);
            }
        };
        if (live || seam) body(std::true_type{}); else body(std::false_type{});

        gchunk += (unsigned int)nch;
        if (__any_sync(0xffffffffu, assigned_any) && lane == 0)
            fast_mark_flags(a.self.tflags, a.lower.tflags, a.upper.tflags, a.lower.ny, a.upper.ny, a.nx, a.ny, a.nfbx, a.nfby,
                            strip * F2_OUTC, y0, F2_OUTC, y1 - y0);
    }
    signal_neighbours(a);
}

}  // namespace kob
#endif  // KOB_FAST2_CUH
