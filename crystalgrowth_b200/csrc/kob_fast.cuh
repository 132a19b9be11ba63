// kob_fast.cuh — FAST fused Kobayashi step (FP32): the roofline kernel.  One launch = one explicit-Euler step
// = pass 1 + pass 2 of the reference (src/Kobayashi.cpp:125-175, :177-221), no scratch arrays in HBM.
//
// Same model as the STRICT kernel (dead-band angle state machine with carried theta, PI_F, 9-point Laplacians,
// Jacobi update); differences are rounding-level only: reciprocal multiplies for the divisions by loop
// constants, FMA contraction, re-associated Laplacian sums, and a TRIG-FREE anisotropy for integer mode j:
// cos(j*theta), sin(j*theta) = Re/Im ((gx + i gy)/|g|)^j — one rsqrt and a few FMAs instead of div+atan+sin+cos.
//
// Structure (B200-first):
//   * persistent grid; a CTA claims 8 adjacent column strips of one row segment from a global counter and hands one
//     strip to each warp, which then works on its own (TMA ring, registers, stores).  Far-field CTA jobs keep their
//     warps in lock-step (one barrier per chunk) so that a grid row is fetched as 1920 contiguous bytes; jobs with
//     data-dependent work free-run.  (cta_jobs = 0 falls back to per-warp jobs with no barrier at all.)
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
//   * far-field shortcut: a chunk whose phi rows (and the 4 rows before them) are all +0 only diffuses T — bit-identical
//     to the full path, and what makes sparse (seeded) fields purely HBM bound.
//   * the data-dependent block (angle, anisotropy, m(T), noise) is packed f32x2 as well: one minimax atan polynomial
//     serves both the angle and m(T); the Philox4x32-10 block of 4 cells is drawn once per lane pair.
#ifndef KOB_FAST_CUH
#define KOB_FAST_CUH

#include <cuda.h>   // CUtensorMap (type only; the encode entry point is fetched at run time, no -lcuda)

#include <type_traits>

#include "kob_common.cuh"

namespace kob {

struct FastMaps {
    CUtensorMap phi[2];
    CUtensorMap t[2];
};

struct FastArgs {
    unsigned long long* job_ctr;   // monotonically increasing across launches
    unsigned long long job_base;   // counter value at which this launch's job 0 sits
    // jobs = nstrips x nseg.  Row segments: nseg_a segments of yj rows, then segments of yj_b rows up to ny
    // (guided scheduling: big jobs first, small jobs last, so that the tail of the dynamic queue is short).
    int nstrips, nseg, yj, nseg_a, yj_b;
    int cta_jobs;                  // 1/2: a CTA claims 8 adjacent strips of one segment; 1 = always in lock-step, 2 = adaptive
    int nstrips_p;                 // strips padded to a multiple of the warps per CTA (cta_jobs only)
    int no_skip;                   // test knob: never take the far-field (phi == +0) chunk shortcut
    // constants of the far field / held-with-theta==0 cells: eps and eps' at theta = 0
    float eps0, epsd0;
    float cj0, sj0;                // cos(j*theta0), sin(j*theta0) for the theta0 rotation
    float ebd;                     // epsbar*delta
    float il_dt;                   // inv_lapden*dt
    float two_pi, half_pi;         // 2*PI_F, 0.5*PI_F
    float m_off;                   // (alpha/PI_F) * pi/2: m(T) for |gamma (T_eq - T)| -> inf
    uint32_t pk[20];               // Philox round keys: pk[2r] = seed_lo + r*W0, pk[2r+1] = seed_hi + r*W1
    uint32_t pc2, pc3;             // Philox counter words 2, 3 = (step_lo, step_hi)
    long long ny_global;           // rows of the whole torus (two-step kernel: noise of wrapped ghost rows)
    // two-step kernel, general pass: jobs come from the work list the far pass wrote (nullptr: all jobs, from job_ctr)
    const int* list;
    const unsigned int* list_count;
    unsigned int* list_claim;
    unsigned int* live_ctr;        // single-step kernel, probe launches only: counts the jobs that see live theta flags
};

// Philox4x32-10 with the per-round keys (key + r * Weyl) precomputed on the host into the constant bank and the
// (step_lo, step_hi) half of the counter taken from there too.  Same bits as kob_math.h's philox4x32_10.
__device__ __forceinline__ Philox4 fast_philox(const FastArgs& f, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0, p1 = (unsigned long long)0xCD9E8D57u * c2;
        c0 = (uint32_t)(p1 >> 32) ^ c1 ^ f.pk[2 * r];
        c2 = (uint32_t)(p0 >> 32) ^ c3 ^ f.pk[2 * r + 1];
        c1 = (uint32_t)p1; c3 = (uint32_t)p0;
    }
    Philox4 o;
    o.w[0] = c0; o.w[1] = c1; o.w[2] = c2; o.w[3] = c3;
    return o;
}
__device__ __forceinline__ Philox4 fast_philox(const FastArgs& f, uint32_t c0, uint32_t c1) { return fast_philox(f, c0, c1, f.pc2, f.pc3); }

template <int NP>
struct FastGeom {
    static constexpr int CPL = 2 * NP;            // cells per lane
    static constexpr int WCOLS = 32 * CPL;        // pass-1 columns per warp
    static constexpr int OUTC = WCOLS - 2 * CPL;  // output columns per strip (lanes 1..30)
    static constexpr int BW = WCOLS + 2 * CPL;    // TMA box width (own cells of lane L at box column CPL*L + CPL)
};

#ifndef KOB_FAST_RB
#define KOB_FAST_RB 4
#endif
#ifndef KOB_FAST_NST
#define KOB_FAST_NST 4
#endif
constexpr int FAST_RB = KOB_FAST_RB;     // rows per TMA chunk (= unroll of the row loop)
static_assert(FAST_RB >= 4, "the far-field shortcut needs a chunk to cover the 4-row history of the register windows");
constexpr int FAST_NST = KOB_FAST_NST;   // TMA stages per warp

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

// ---- packed FP32x2 arithmetic (Blackwell FADD2 / FMUL2 / FFMA2: two cells per issue slot) -----------------
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 f2add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 f2sub(float2 a, float2 b) {
    unsigned long long ra, rb, rc;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(rc) : "l"(ra), "l"(rb));
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(rc));
    return r;
}
__device__ __forceinline__ float2 f2neg(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float rsqrt_approx(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// atan on [0, 1] for a pair of cells: w + w*t*Q(t), t = w^2, Q = degree-7 minimax (max abs error 8.5e-8 in FP32,
// 1.4 ulp at pi/4 — the class of CUDA's atanf, at half the issue slots because both cells share every FFMA2).
__device__ __forceinline__ float2 atan01_2(float2 w) {
    const float2 t = f2mul(w, w);
    // Horner: an Estrin split (depth 4 instead of 7) measured 1.5 % slower — the block is issue-, not chain-limited
    float2 p = f2fma(f2(0.002622196450829506f), t, f2(-0.015132336877286434f));
    p = f2fma(p, t, f2(0.04112152010202408f));
    p = f2fma(p, t, f2(-0.0736667588353157f));
    p = f2fma(p, t, f2(0.10573916882276535f));
    p = f2fma(p, t, f2(-0.14185971021652222f));
    p = f2fma(p, t, f2(0.1999039649963379f));
    p = f2fma(p, t, f2(-0.33332985639572144f));
    return f2fma(f2mul(w, t), p, w);
}
constexpr float HALF_PI_TRUE = 1.57079632679489662f;   // atan(+inf): range reduction of the atan itself (not PI_F)

// (c + i s)^J for a pair of cells, J a compile-time constant
template <int J>
__device__ __forceinline__ void cpow2(float2 c, float2 s, float2& C, float2& S) {
    if (J == 0) { C = f2(1.0f); S = f2(0.0f); return; }
    if (J == 1) { C = c; S = s; return; }
    float2 hc, hs;
    cpow2<J / 2>(c, s, hc, hs);
    const float2 qc = f2fma(hc, hc, f2neg(f2mul(hs, hs))), qs = f2mul(f2add(hc, hc), hs);
    if (J & 1) { C = f2fma(qc, c, f2neg(f2mul(qs, s))); S = f2fma(qc, s, f2mul(qs, c)); }
    else { C = qc; S = qs; }
}
__device__ __forceinline__ void cpow2_rt(int j, float2 c, float2 s, float2& C, float2& S) {   // 0 <= j <= 16, warp-uniform
    float2 rc = f2(1.0f), rs = f2(0.0f);
#pragma unroll
    for (int bit = 4; bit >= 0; --bit) {
        const float2 qc = f2fma(rc, rc, f2neg(f2mul(rs, rs))), qs = f2mul(f2add(rc, rc), rs);
        rc = qc; rs = qs;
        if ((j >> bit) & 1) { const float2 tc = f2fma(rc, c, f2neg(f2mul(rs, s))), ts = f2fma(rc, s, f2mul(rs, c)); rc = tc; rs = ts; }
    }
    C = rc; S = rs;
}
// component k (compile-time) of an array of cell pairs
#define KOB_CX(arr, k) (((k) & 1) ? (arr)[(k) >> 1].y : (arr)[(k) >> 1].x)

// ---- rare paths, kept out of line so that the steady-state row loop stays small in the instruction cache ----
__device__ __noinline__ void fast_sincos(float arg, float* s, float* c) { sincosf(arg, s, c); }
// Store to every alias of an owned seam cell (own ghost columns, neighbour strips' ghost rows).
__device__ __noinline__ void fast_store_edge(float* self_buf, float* lower_buf, float* upper_buf, long long pitch, int nx,
                                             int ny, long long ny_lower, int i, int j, float v) {
    store_aliases<float>(self_buf, lower_buf, upper_buf, pitch, nx, ny, ny_lower, i, j, v);
}
__device__ __noinline__ void fast_mark_flags(uint32_t* self_f, uint32_t* lower_f, uint32_t* upper_f, long long ny_lower,
                                             long long ny_upper, int nx, int ny, int nfbx, int nfby, int x0, int y0,
                                             int tx, int ty) {
    StepArgs<float> a;
    a.self.tflags = self_f; a.lower.tflags = lower_f; a.upper.tflags = upper_f;
    a.lower.ny = ny_lower; a.upper.ny = ny_upper; a.nx = nx; a.ny = ny; a.nfbx = nfbx; a.nfby = nfby;
    mark_tile_flags<float>(a, x0, y0, tx, ty);
}

// JM: 4 / 6 = compile-time integer mode, 0 = run-time integer mode (prm.jmode in 0..16), -1 = any real j (trig).
// NP = 1: 8 warps x 2 CTAs per SM at <= 128 registers.  NP = 2 (4 cells per lane) needs ~2x the register windows: 12 warps
// x 1 CTA per SM at <= 168 registers.
#ifndef KOB_FAST_WARPS2
#define KOB_FAST_WARPS2 12
#endif
template <int NP>
__host__ __device__ constexpr int fast_warps() { return NP == 2 ? KOB_FAST_WARPS2 : 8; }

template <int NP, int JM, bool NOISE, bool ROT>
__global__ void __launch_bounds__(32 * fast_warps<NP>(), NP == 2 ? 1 : 2) kob_step_fast(const __grid_constant__ FastMaps maps, const StepArgs<float> a,
                                                         const FastArgs f) {
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
    const float A0 = f.eps0 * f.eps0, B0 = f.eps0 * f.epsd0;
    unsigned int gchunk = 0;   // chunks consumed so far by this warp: stage = gchunk % NST, parity = (gchunk / NST) & 1

    __shared__ unsigned long long s_job;
    const int nsp = f.cta_jobs ? f.nstrips_p : f.nstrips;     // strips per segment in the job numbering
    const int njobs_q = nsp * f.nseg;
    for (;;) {
        unsigned long long jraw = 0;
        if (f.cta_jobs) {
            __syncthreads();
            if (threadIdx.x == 0) s_job = atomicAdd(f.job_ctr, (unsigned long long)nwarps) - f.job_base;
            __syncthreads();
            jraw = s_job + (unsigned long long)warp;
            if (s_job >= (unsigned long long)njobs_q) break;
        } else {
            if (lane == 0) jraw = atomicAdd(f.job_ctr, 1ull) - f.job_base;
            jraw = __shfl_sync(0xffffffffu, jraw, 0);
            if (jraw >= (unsigned long long)njobs_q) break;
        }
        const int job = (int)jraw;
        const int strip = job - (job / nsp) * nsp;
        // queue order: the two segments on the torus seam first (they take the slower generic path), small ones last
        const int sq = job / nsp;
        const int seg_ = sq == 0 ? 0 : (sq == 1 ? f.nseg - 1 : sq - 1);
        const int y0 = seg_ < f.nseg_a ? seg_ * f.yj : f.nseg_a * f.yj + (seg_ - f.nseg_a) * f.yj_b;
        const int y1 = min(y0 + (seg_ < f.nseg_a ? f.yj : f.yj_b), a.ny);
        const int xs = strip * G::OUTC - CPL;            // first pass-1 column of the warp (lane 0, halo)
        const int x = xs + CPL * lane;                   // first cell of this lane
        const bool mid_lane = lane >= 1 && lane <= 30;

        if (a.linked) {
            if (lane == 0) {
                if (y0 < GY + 1) wait_flag(&a.self.arrive[0], a.epoch, &a.self.arrive[2]);
                if (y1 > a.ny - GY - 1) wait_flag(&a.self.arrive[1], a.epoch, &a.self.arrive[2]);
            }
            __syncwarp();
        }
        // theta may be non-zero somewhere in the pass-1 footprint of this job?  One bit per 32-row flag row.
        uint32_t livemask = 0;
        const int fby0 = max((y0 - 1 + GY) / FBY, 0);
        {
            const int bx0 = max((xs + GX) / FBX, 0), bx1 = min((xs + GX + G::WCOLS - 1) / FBX, a.nfbx - 1);
            const int by1 = min((y1 + GY) / FBY, a.nfby - 1);
            const int nbx = bx1 - bx0 + 1, nby = by1 - fby0 + 1;
            for (int i0 = 0; i0 < nby; i0 += 32) {       // lane i looks at flag row fby0 + i0 + i
                uint32_t fl = 0;
                if (i0 + lane < nby)
                    for (int bx = 0; bx < nbx; ++bx) fl |= __ldcg(&a.self.tflags[(fby0 + i0 + lane) * a.nfbx + bx0 + bx]);
                const uint32_t m = __ballot_sync(0xffffffffu, fl != 0u);
                livemask |= i0 == 0 ? m : (m ? 0x80000000u : 0u);   // flag rows beyond 31 fold into the last bit
            }
        }
        const bool live = livemask != 0u;
        if (f.live_ctr && live && lane == 0 && strip < f.nstrips) atomicAdd(f.live_ctr, 1u);     // density probe (adaptive policy)
        // seam job: touches the first/last GXR columns or GY rows -> alias stores, ragged right edge
        const bool seam = strip == 0 || (strip + 1) * G::OUTC > a.nx - GXR || y0 < GY || y1 > a.ny - GY;

        const int nrows = (y1 - y0) + 4;                 // streamed phi rows y0-2 .. y1+1
        const int nch = (nrows + RB - 1) / RB;
        // CTA-wide jobs: the 8 adjacent strips advance in lock-step (one barrier per chunk) so that a row is fetched as
        // 1920 contiguous bytes.  cta_jobs == 2: only while none of the 8 does data-dependent work (live theta / seam) —
        // there the rows take unequal time and the barrier would only add waiting.
        bool lock = f.cta_jobs == 1;
        if (f.cta_jobs == 2) lock = !__syncthreads_or((live || seam) && strip < f.nstrips);
        if (strip >= f.nstrips) {                        // padding job (cta_jobs): only keep the CTA's barriers company
            if (lock)
                for (int c = 0; c < nch; ++c) __syncthreads();
            continue;
        }
        const int box_x = xs - CPL + GX;                 // padded x of box column 0
        auto issue = [&](int c) {                        // lane 0: chunk c of this job -> stage ((gchunk + c) % NST)
            const unsigned int gi = gchunk + (unsigned int)c;
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

        bool assigned_any = false;

        // The row loop, instantiated twice: GEN = false is the lean steady-state path (interior strip, no theta
        // to read, no seams); GEN = true additionally reads held theta, handles the ragged edge and the aliases.
        // MODE 0 (lean): interior strip, theta all zero in the footprint.  MODE 1 (live): interior strip, held
        // theta is read.  MODE 2 (seam): additionally the ragged right edge and the alias stores.
        auto body = [&](auto mode_tag) {
            constexpr int MODE = decltype(mode_tag)::value;
            constexpr bool GEN = MODE != 0, SEAM = MODE == 2;
            // register windows, one float2 per pair of adjacent cells; "r" is the phi row streamed in this iteration
            float2 po0[NP], po1[NP];               // phi rows r-2, r-1
            float2 gx1[NP], gx2[NP], gy2[NP];      // gx(r-1); gx, gy (r-2)
            float2 u1[NP], lp1[NP], lap2[NP];      // u(r-1), c(r-1)+u(r-2), complete 9-point sum of row r-2
            float2 tq1[NP], tu1[NP], tlp1[NP];     // T(r-2), u_T(r-2), c_T(r-2)+u_T(r-3)        [T lags phi by a row]
            float2 A2[NP], A3[NP];                 // eps^2 rows r-2, r-3
            float2 P2[NP], P3[NP];                 // eps*eps'*gx rows r-2, r-3
            float2 Q2[NP];                         // eps*eps'*gy row r-2
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                po0[p] = po1[p] = gx1[p] = gx2[p] = gy2[p] = u1[p] = lp1[p] = lap2[p] = f2(0.f);
                tq1[p] = tu1[p] = tlp1[p] = f2(0.f);
                A2[p] = A3[p] = P2[p] = P3[p] = Q2[p] = f2(0.f);
            }
            // GEN: theta of the held cells, prefetched two rows ahead: thp0 = theta(r-1), thp1 = theta(r)
            float thp0[CPL], thp1[CPL];
#pragma unroll
            for (int k = 0; k < CPL; ++k) thp0[k] = thp1[k] = 0.f;
            // running pointers to cell (x, r-2) of the output arrays / theta
            const long long o2 = pidx<float>(pitch, x, y0 - 4);
            float* pphi = phi_out + o2;
            float* ptt = t_out + o2;
            const unsigned int nvalid = (unsigned int)(y1 - y0);
            const unsigned int nstore = mid_lane ? nvalid : 0u;    // rows this lane stores
            const float2 idx2 = f2(P.inv_dx), idy2 = f2(P.inv_dy), il2 = f2(P.inv_lapden), ildt2 = f2(f.il_dt);
            const float2 dtt2 = f2(P.dt_over_tau), K2 = f2(P.K), two2 = f2(2.0f), m12 = f2(-12.0f), B02 = f2(B0);
            bool prevz = false;                    // !GEN: the previous chunk's phi rows were all +0
            bool have_next = false;                // NP = 1 noise: the odd row's Philox words were drawn at the even row
            uint32_t nxa = 0u, nxb = 0u;
            float* pthe = a.self.theta + (o2 + pitch);   // theta of cell (x, r-1): the row pass 1 re-assigns
            const long long pitch2 = 2 * pitch;

            for (int c = 0; c < nch; ++c) {
                const unsigned int gi = gchunk + (unsigned int)c;
                const int st = gi % NST;
                if (lock) __syncthreads();                                          // adjacent strips advance together
                mbar_wait(&bars[st], (gi / NST) & 1u);
                const float* sp = stages + st * STAGE_FLOATS + CPL * lane + CPL;   // this lane's own phi cells
                const float* stt = sp + BOX_FLOATS;                                // T rows (one row behind)
                const int yrel0 = c * RB - 4;                                       // (r - 2) - y0 for rr = 0
                if (!GEN) {
                    // ---- far field: phi == +0 on all 64 own columns of this chunk's RB rows AND of the previous RB
                    // rows.  Every phi term of the step is then exactly +0 and the register windows already sit at
                    // their all-zero-input fixed point (each is a function of the last <= 5 streamed rows only), so
                    // the chunk reduces to the T diffusion; outputs are bit-identical to the full path. ----
                    uint32_t bits = 0u;
#pragma unroll
                    for (int rr = 0; rr < RB; ++rr) {
#pragma unroll
                        for (int k = 0; k < CPL; ++k) bits |= __float_as_uint(sp[rr * BW + k]);
                    }
                    const bool curz = !__any_sync(0xffffffffu, bits != 0u);
                    const bool skip = curz && prevz && !f.no_skip;
                    prevz = curz;
                    if (skip) {
#pragma unroll
                        for (int rr = 0; rr < RB; ++rr) {
                            const unsigned int yrel = (unsigned int)(yrel0 + rr);
                            const float* row = stt + rr * BW;
                            const float w = row[-1], ee = row[CPL];
                            float2 tn[NP], thsum[NP];
                            if (NP == 1) tn[0] = *reinterpret_cast<const float2*>(row);
                            else { const float4 v = *reinterpret_cast<const float4*>(row); tn[0] = make_float2(v.x, v.y); tn[NP - 1] = make_float2(v.z, v.w); }
#pragma unroll
                            for (int k = 0; k < CPL; ++k) {
                                const float l = k == 0 ? w : KOB_CX(tn, k - 1), rgt = k == CPL - 1 ? ee : KOB_CX(tn, k + 1);
                                KOB_CX(thsum, k) = l + rgt;
                            }
                            float2 nt_[NP];
#pragma unroll
                            for (int p = 0; p < NP; ++p) {
                                const float2 tu_new = f2fma(two2, tn[p], thsum[p]);
                                const float2 lapt = f2add(tlp1[p], tu_new);
                                nt_[p] = f2fma(K2, f2(0.f), f2fma(lapt, ildt2, tq1[p]));       // :215 with phi+ - phi = +0
                                tlp1[p] = f2fma(two2, thsum[p], f2fma(m12, tn[p], tu1[p]));
                                tu1[p] = tu_new;
                                tq1[p] = tn[p];
                            }
                            if (yrel < nstore) {
                                if (NP == 1) {
                                    *reinterpret_cast<float2*>(pphi) = f2(0.f);
                                    *reinterpret_cast<float2*>(ptt) = nt_[0];
                                } else {
                                    *reinterpret_cast<float4*>(pphi) = make_float4(0.f, 0.f, 0.f, 0.f);
                                    *reinterpret_cast<float4*>(ptt) = make_float4(nt_[0].x, nt_[0].y, nt_[NP - 1].x, nt_[NP - 1].y);
                                }
                            }
                            pphi += pitch;
                            ptt += pitch;
                            pthe += pitch;
                        }
                        __syncwarp();
                        if (lane == 0 && c + NST < nch) issue(c + NST);
                        continue;
                    }
                }
                bool lrow_c = false;          // GEN: some theta-flag row under this chunk's prefetch rows (r+1) is live
                if (GEN && live) {
                    const int f0 = ((y0 + yrel0 + 3 + GY) >> 5) - fby0, f1 = ((y0 + yrel0 + RB + 2 + GY) >> 5) - fby0;   // FBY == 32
                    lrow_c = ((livemask >> min(max(f0, 0), 31)) | (livemask >> min(max(f1, 0), 31))) & 1u;
                }
#pragma unroll
                for (int rr = 0; rr < RB; ++rr) {
                    const unsigned int yrel = (unsigned int)(yrel0 + rr);           // row of pass 2, relative to y0
                    // horizontal neighbours of the pass-1 products of row r-2, issued early: the shuffle latency hides
                    // behind this row's pass 1
                    const float A_w = __shfl_up_sync(0xffffffffu, A2[NP - 1].y, 1);
                    const float A_e = __shfl_down_sync(0xffffffffu, A2[0].x, 1);
                    const float Q_w = __shfl_up_sync(0xffffffffu, Q2[NP - 1].y, 1);
                    const float Q_e = __shfl_down_sync(0xffffffffu, Q2[0].x, 1);
                    // ---- phi row r: own cells, horizontal sums and x-gradient ----
                    float2 pn[NP], hsum[NP], gxn[NP];
                    {
                        const float* row = sp + rr * BW;
                        const float w = row[-1], ee = row[CPL];
                        if (NP == 1) pn[0] = *reinterpret_cast<const float2*>(row);
                        else { const float4 v = *reinterpret_cast<const float4*>(row); pn[0] = make_float2(v.x, v.y); pn[NP - 1] = make_float2(v.z, v.w); }
                        float2 gxd[NP];
#pragma unroll
                        for (int k = 0; k < CPL; ++k) {
                            const float l = k == 0 ? w : KOB_CX(pn, k - 1), rgt = k == CPL - 1 ? ee : KOB_CX(pn, k + 1);
                            KOB_CX(hsum, k) = l + rgt;
                            KOB_CX(gxd, k) = rgt - l;
                        }
#pragma unroll
                        for (int p = 0; p < NP; ++p) gxn[p] = f2mul(gxd[p], idx2);           // :139
                    }
                    // ---- T row r-1: own cells + horizontal sums ----
                    float2 tn[NP], thsum[NP];
                    {
                        const float* row = stt + rr * BW;
                        const float w = row[-1], ee = row[CPL];
                        if (NP == 1) tn[0] = *reinterpret_cast<const float2*>(row);
                        else { const float4 v = *reinterpret_cast<const float4*>(row); tn[0] = make_float2(v.x, v.y); tn[NP - 1] = make_float2(v.z, v.w); }
#pragma unroll
                        for (int k = 0; k < CPL; ++k) {
                            const float l = k == 0 ? w : KOB_CX(tn, k - 1), rgt = k == CPL - 1 ? ee : KOB_CX(tn, k + 1);
                            KOB_CX(thsum, k) = l + rgt;
                        }
                    }
                    // ---- pass 1 for row r-1 (phi rows r-2, r-1, r), far-field values first ----
                    float2 An[NP], Pn[NP], Qn[NP], gyn[NP], q[NP], radd[NP];
                    bool asg[CPL];
                    bool interesting = false;        // some cell re-assigns its angle (pass 1) or has phi(1-phi) != 0 (pass 2)
#pragma unroll
                    for (int p = 0; p < NP; ++p) {
                        gyn[p] = f2mul(f2sub(pn[p], po0[p]), idy2);                          // :140
                        An[p] = f2(A0); Pn[p] = f2mul(B02, gx1[p]); Qn[p] = f2mul(B02, gyn[p]);   // cells holding theta = 0
                        q[p] = f2fma(make_float2(-po0[p].x, -po0[p].y), po0[p], po0[p]);     // phi (1 - phi) of row r-2
                        radd[p] = f2(0.f);
                    }
#pragma unroll
                    for (int k = 0; k < CPL; ++k) {
                        asg[k] = (KOB_CX(gx1, k) < -e) || (fabsf(KOB_CX(gyn, k)) > e);      // :154-167: theta re-assigned
                        interesting |= asg[k] || (KOB_CX(q, k) != 0.f);
                    }
                    if (GEN) {
#pragma unroll
                        for (int k = 0; k < CPL; ++k) interesting |= thp0[k] != 0.f;       // a held cell may carry an angle
                    }
                    // ONE vote per row; everything data dependent lives in this cold block, as one straight-line
                    // region (angle, anisotropy, m(T) and the noise draw interleave in the issue stream)
                    if ((rr & 1) == 0) have_next = false;
                    if (__any_sync(0xffffffffu, interesting)) {
                        const bool row_owned = yrel + 1u < nvalid;
                        float th_old[CPL];
#pragma unroll
                        for (int k = 0; k < CPL; ++k) th_old[k] = (GEN && !asg[k]) ? thp0[k] : 0.f;
                        float2 c1[NP], s1[NP], th2[NP];
                        bool rare = false;
#pragma unroll
                        for (int p = 0; p < NP; ++p) {
                            const float2 gx = gx1[p], gy = gyn[p];
                            const float2 agx = make_float2(fabsf(gx.x), fabsf(gx.y)), agy = make_float2(fabsf(gy.x), fabsf(gy.y));
                            // reference angle (:154-167): atan(|gy|/|gx|) folded into [0, pi/4], then the quadrant
                            const float2 mn = make_float2(fminf(agx.x, agy.x), fminf(agx.y, agy.y));
                            const float2 mx = make_float2(fmaxf(agx.x, agy.x), fmaxf(agx.y, agy.y));
                            float2 r = atan01_2(f2mul(mn, make_float2(rcp_approx(mx.x), rcp_approx(mx.y))));
                            const bool sw0 = agy.x > agx.x, sw1 = agy.y > agx.y;
                            r = f2fma(r, make_float2(sw0 ? -1.0f : 1.0f, sw1 ? -1.0f : 1.0f),
                                      make_float2(sw0 ? HALF_PI_TRUE : 0.0f, sw1 ? HALF_PI_TRUE : 0.0f));
                            r.x = __uint_as_float(__float_as_uint(r.x) ^ ((__float_as_uint(gx.x) ^ __float_as_uint(gy.x)) & 0x80000000u));
                            r.y = __uint_as_float(__float_as_uint(r.y) ^ ((__float_as_uint(gx.y) ^ __float_as_uint(gy.y)) & 0x80000000u));
                            th2[p] = f2add(make_float2(gx.x < 0.f ? pi : (gy.x < 0.f ? f.two_pi : 0.0f),
                                                       gx.y < 0.f ? pi : (gy.y < 0.f ? f.two_pi : 0.0f)), r);
                            // unit vector of the gradient (trig-free anisotropy)
                            const float2 r2 = f2fma(gx, gx, f2mul(gy, gy));
                            const float2 rinv = make_float2(rsqrt_approx(r2.x), rsqrt_approx(r2.y));
                            c1[p] = f2mul(gx, rinv); s1[p] = f2mul(gy, rinv);
                        }
#pragma unroll
                        for (int k = 0; k < CPL; ++k)
                            rare |= (asg[k] && fabsf(KOB_CX(gx1, k)) <= e) || (GEN && th_old[k] != 0.f);
                        // rare cells: dead-band in gx (case A, :154-158: theta = +-PI_F/2) and held non-zero angles,
                        // whose unit vector is (cos theta, sin theta) by MUFU after folding theta into [-pi, pi]
                        if (__any_sync(0xffffffffu, rare)) {
#pragma unroll
                            for (int k = 0; k < CPL; ++k) {
                                const bool fl = asg[k] && fabsf(KOB_CX(gx1, k)) <= e;
                                const bool held = GEN && th_old[k] != 0.f;
                                const float sg = KOB_CX(gyn, k) < 0.f ? -1.0f : 1.0f;
                                float t = th_old[k];
                                if (GEN) t = t > 3.14159265358979f ? fmaf(-1.0f, 6.28318548202514648f, t) + 1.74845553e-7f : t;   // - 2 pi (hi, lo)
                                const float ct = GEN ? __cosf(t) : 0.f, st = GEN ? __sinf(t) : 0.f;
                                KOB_CX(th2, k) = fl ? sg * f.half_pi : KOB_CX(th2, k);
                                KOB_CX(c1, k) = fl ? 0.0f : (held ? ct : KOB_CX(c1, k));
                                KOB_CX(s1, k) = fl ? sg : (held ? st : KOB_CX(s1, k));
                            }
                        }
                        float2 Cc[NP], Ss[NP];
#pragma unroll
                        for (int p = 0; p < NP; ++p) {
                            Cc[p] = f2(1.0f); Ss[p] = f2(0.0f);
                            if (JM >= 0) {
                                if (JM == 0) cpow2_rt(P.jmode, c1[p], s1[p], Cc[p], Ss[p]); else cpow2<JM>(c1[p], s1[p], Cc[p], Ss[p]);
                                if (ROT) {
                                    const float2 c2 = f2fma(Cc[p], f2(f.cj0), f2mul(Ss[p], f2(f.sj0)));
                                    const float2 s2 = f2fma(Ss[p], f2(f.cj0), f2neg(f2mul(Cc[p], f2(f.sj0))));
                                    Cc[p] = c2; Ss[p] = s2;
                                }
                            }
                        }
                        if (JM < 0) {                                                       // any real j: trig on the angle
#pragma unroll
                            for (int k = 0; k < CPL; ++k) {
                                const float th = asg[k] ? KOB_CX(th2, k) : th_old[k];
                                if (asg[k] || th != 0.f) {
                                    float C, S;
                                    fast_sincos(P.aniso * (th - P.theta0), &S, &C);
                                    KOB_CX(Cc, k) = C; KOB_CX(Ss, k) = S;
                                }
                            }
                        }
                        // store the re-assigned angles of owned cells
                        if (row_owned && mid_lane) {
                            if (!SEAM) {
#pragma unroll
                                for (int k = 0; k < CPL; ++k) {
                                    if (asg[k]) pthe[k] = KOB_CX(th2, k);
                                    assigned_any |= asg[k];
                                }
                            } else {
#pragma unroll
                                for (int k = 0; k < CPL; ++k) {
                                    if (asg[k] && x + k < a.nx) {
                                        const float th = KOB_CX(th2, k);
                                        const int y = y0 + (int)yrel + 1;
                                        if (y < GY || y >= a.ny - GY)
                                            fast_store_edge(a.self.theta, a.lower.theta, a.upper.theta, pitch, a.nx, a.ny, a.lower.ny, x + k, y, th);
                                        else {
                                            float* pth = pthe + k;
                                            *pth = th;
                                            if (x + k < GXR) pth[a.nx] = th;
                                            if (x + k >= a.nx - GXR) pth[-a.nx] = th;
                                        }
                                        assigned_any = true;
                                    }
                                }
                            }
                        }
#pragma unroll
                        for (int p = 0; p < NP; ++p) {
                            float2 ep = f2fma(f2(f.ebd), Cc[p], f2(P.epsbar));              // :170
                            float2 ed = f2mul(f2(P.neg_ebjd), Ss[p]);                        // :171
                            const bool d0 = !asg[2 * p] && !(GEN && th_old[2 * p] != 0.f);   // holds theta = 0
                            const bool d1 = !asg[2 * p + 1] && !(GEN && th_old[2 * p + 1] != 0.f);
                            ep = make_float2(d0 ? f.eps0 : ep.x, d1 ? f.eps0 : ep.y);
                            ed = make_float2(d0 ? f.epsd0 : ed.x, d1 ? f.epsd0 : ed.y);
                            An[p] = f2mul(ep, ep);
                            const float2 B = f2mul(ep, ed);
                            Pn[p] = f2mul(B, gx1[p]);
                            Qn[p] = f2mul(B, gyn[p]);
                        }
                        // ---- reaction term q*((phi - 1/2) + m(T)) [+ noise] of row r-2, :206-214 ----
                        float2 rq[NP];                      // r - 1/2 of the noise draw
                        if (NOISE) {
                            const int y = y0 + (int)yrel;
                            if (a.noise_field) {
#pragma unroll
                                for (int k = 0; k < CPL; ++k)
                                    KOB_CX(rq, k) = ((KOB_CX(q, k) != 0.f && x + k >= 0 && x + k < a.nx && yrel < nvalid)
                                                ? __ldg(&a.noise_field[(long long)(x + k) + (long long)a.nx * y]) : 0.5f) - 0.5f;
                            } else if (NP == 2) {
                                const Philox4 ph = fast_philox(f, (uint32_t)x >> 2, (uint32_t)(a.y0 + y));
#pragma unroll
                                for (int k = 0; k < 4; ++k) KOB_CX(rq, k) = fmaf((float)(ph.w[k] >> 8), 5.9604644775390625e-8f, -0.5f);
                            } else {
                                // NP = 1: lanes (2m+1, 2m+2) share a Philox block (4 cells); over a row pair the low lane
                                // draws the block of row y, the high lane the block of row y+1, and they swap halves
                                const bool hi = (x & 2) != 0;
                                uint32_t wa, wb;
                                if ((rr & 1) == 0) {
                                    const Philox4 ph = fast_philox(f, (uint32_t)x >> 2, (uint32_t)(a.y0 + y) + (hi ? 1u : 0u));
                                    const int partner = hi ? lane - 1 : lane + 1;
                                    const uint32_t ra = __shfl_sync(0xffffffffu, hi ? ph.w[0] : ph.w[2], partner);
                                    const uint32_t rb = __shfl_sync(0xffffffffu, hi ? ph.w[1] : ph.w[3], partner);
                                    wa = hi ? ra : ph.w[0]; wb = hi ? rb : ph.w[1];
                                    nxa = hi ? ph.w[2] : ra; nxb = hi ? ph.w[3] : rb;
                                    have_next = true;
                                } else if (have_next) {
                                    wa = nxa; wb = nxb;
                                } else {
                                    const Philox4 ph = fast_philox(f, (uint32_t)x >> 2, (uint32_t)(a.y0 + y));
                                    wa = hi ? ph.w[2] : ph.w[0]; wb = hi ? ph.w[3] : ph.w[1];
                                }
                                rq[0] = f2fma(make_float2((float)(wa >> 8), (float)(wb >> 8)), f2(5.9604644775390625e-8f), f2(-0.5f));
                            }
                        }
#pragma unroll
                        for (int p = 0; p < NP; ++p) {
                            // m = (alpha/PI_F) atan(gamma (T_eq - T)), :206 — atan folded into [0, 1] by 1/|x|
                            const float2 xa = f2mul(f2(P.gamma), f2sub(f2(P.teq), tq1[p]));
                            const float ax0 = fabsf(xa.x), ax1 = fabsf(xa.y);
                            const bool b0 = ax0 > 1.0f, b1 = ax1 > 1.0f;
                            const float2 r = atan01_2(make_float2(b0 ? rcp_approx(ax0) : ax0, b1 ? rcp_approx(ax1) : ax1));
                            float2 m = f2fma(r, make_float2(b0 ? -P.alpha_over_pi : P.alpha_over_pi, b1 ? -P.alpha_over_pi : P.alpha_over_pi),
                                             make_float2(b0 ? f.m_off : 0.0f, b1 ? f.m_off : 0.0f));
                            m.x = __uint_as_float(__float_as_uint(m.x) ^ (__float_as_uint(xa.x) & 0x80000000u));
                            m.y = __uint_as_float(__float_as_uint(m.y) ^ (__float_as_uint(xa.y) & 0x80000000u));
                            float2 rv = f2mul(q[p], f2add(f2sub(po0[p], f2(0.5f)), m));                      // :214
                            if (NOISE) rv = f2fma(f2mul(f2(P.noise_a), q[p]), rq[p], rv);
                            radd[p] = rv;
                        }
                    }
                    // ---- pass 2 for row y = r-2 (computed unconditionally; stores predicated on the row being owned) ----
                    float2 tu_new[NP];
                    {
                        float2 dA[NP], dQ[NP], np_[NP], nt_[NP];
#pragma unroll
                        for (int k = 0; k < CPL; ++k) {
                            const float Aw = k == 0 ? A_w : KOB_CX(A2, k - 1), Ae = k == CPL - 1 ? A_e : KOB_CX(A2, k + 1);
                            const float Qw = k == 0 ? Q_w : KOB_CX(Q2, k - 1), Qe = k == CPL - 1 ? Q_e : KOB_CX(Q2, k + 1);
                            KOB_CX(dA, k) = Ae - Aw;                                         // :190-192
                            KOB_CX(dQ, k) = Qw - Qe;                                         // term2, :201-203
                        }
#pragma unroll
                        for (int p = 0; p < NP; ++p) {
                            const float2 gEx = f2mul(dA[p], idx2);
                            const float2 gEy = f2mul(f2sub(An[p], A3[p]), idy2);             // :193-195
                            float2 sm = f2fma(f2sub(Pn[p], P3[p]), idy2, radd[p]);           // term1 (:197-199) + reaction
                            sm = f2fma(dQ[p], idx2, sm);
                            sm = f2fma(A2[p], f2mul(lap2[p], il2), sm);                      // eps^2 * lap(phi)
                            sm = f2fma(gEx, gx2[p], sm);                                     // term3, :204
                            sm = f2fma(gEy, gy2[p], sm);
                            np_[p] = f2fma(sm, dtt2, po0[p]);                                // :211
                            tu_new[p] = f2fma(two2, tn[p], thsum[p]);                        // u_T(r-1)
                            const float2 lapt = f2add(tlp1[p], tu_new[p]);                   // 9-point sum of T at row y
                            nt_[p] = f2fma(K2, f2sub(np_[p], po0[p]), f2fma(lapt, ildt2, tq1[p]));   // :215
                        }
                        if (yrel < nstore) {
                            if (!SEAM) {
                                if (NP == 1) {
                                    *reinterpret_cast<float2*>(pphi) = np_[0];
                                    *reinterpret_cast<float2*>(ptt) = nt_[0];
                                } else {
                                    *reinterpret_cast<float4*>(pphi) = make_float4(np_[0].x, np_[0].y, np_[NP - 1].x, np_[NP - 1].y);
                                    *reinterpret_cast<float4*>(ptt) = make_float4(nt_[0].x, nt_[0].y, nt_[NP - 1].x, nt_[NP - 1].y);
                                }
                            } else {
                                const int y = y0 + (int)yrel;
                                if (y < GY || y >= a.ny - GY) {          // rows on the strip seam: every alias (rare)
#pragma unroll
                                    for (int k = 0; k < CPL; ++k)
                                        if (x + k < a.nx) {
                                            fast_store_edge(phi_out, a.lower.phi[a.cur ^ 1], a.upper.phi[a.cur ^ 1], pitch, a.nx, a.ny, a.lower.ny, x + k, y, KOB_CX(np_, k));
                                            fast_store_edge(t_out, a.lower.t[a.cur ^ 1], a.upper.t[a.cur ^ 1], pitch, a.nx, a.ny, a.lower.ny, x + k, y, KOB_CX(nt_, k));
                                        }
                                } else {                                 // interior rows: own cell + ghost-column copy
#pragma unroll
                                    for (int k = 0; k < CPL; ++k)
                                        if (x + k < a.nx) {
                                            pphi[k] = KOB_CX(np_, k);
                                            ptt[k] = KOB_CX(nt_, k);
                                            if (x + k < GXR) { pphi[k + a.nx] = KOB_CX(np_, k); ptt[k + a.nx] = KOB_CX(nt_, k); }
                                            if (x + k >= a.nx - GXR) { pphi[k - a.nx] = KOB_CX(np_, k); ptt[k - a.nx] = KOB_CX(nt_, k); }
                                        }
                                }
                            }
                        }
                    }
                    // ---- GEN: prefetch theta of row r+1 (pass-1 row of the iteration after next) ----
                    if (GEN) {
#pragma unroll
                        for (int k = 0; k < CPL; ++k) { thp0[k] = thp1[k]; thp1[k] = 0.f; }
                        if (lrow_c && yrel + 4u <= nvalid + 1u) {                            // theta rows y0-1 .. y1
                            const float* pf = pthe + pitch2;                                 // theta(x, r+1)
                            if (SEAM) {
#pragma unroll
                                for (int k = 0; k < CPL; ++k)
                                    if (x + k < a.nx + GXR && x + k >= -GXR) thp1[k] = __ldg(pf + k);
                            } else if (NP == 1) {
                                const float2 v = __ldg(reinterpret_cast<const float2*>(pf));
                                thp1[0] = v.x; thp1[1] = v.y;
                            } else {
                                const float4 v = __ldg(reinterpret_cast<const float4*>(pf));
                                thp1[0] = v.x; thp1[1] = v.y; thp1[CPL - 2] = v.z; thp1[CPL - 1] = v.w;
                            }
                        }
                    }
                    // ---- rotate the register windows ----
                    pphi += pitch;
                    ptt += pitch;
                    pthe += pitch;
#pragma unroll
                    for (int p = 0; p < NP; ++p) {
                        tlp1[p] = f2fma(two2, thsum[p], f2fma(m12, tn[p], tu1[p]));          // c_T(r-1) + u_T(r-2)
                        tu1[p] = tu_new[p];
                        tq1[p] = tn[p];
                        const float2 u_new = f2fma(two2, pn[p], hsum[p]);                    // u(r)
                        lap2[p] = f2add(lp1[p], u_new);                                      // 9-point sum of row r-1 complete
                        lp1[p] = f2fma(two2, hsum[p], f2fma(m12, pn[p], u1[p]));             // c(r) + u(r-1)
                        u1[p] = u_new;
                        gx2[p] = gx1[p]; gy2[p] = gyn[p]; gx1[p] = gxn[p];
                        po0[p] = po1[p]; po1[p] = pn[p];
                        A3[p] = A2[p]; A2[p] = An[p];
                        P3[p] = P2[p]; P2[p] = Pn[p];
                        Q2[p] = Qn[p];
                    }
                }
                __syncwarp();
                if (lane == 0 && c + NST < nch) issue(c + NST);
            }
        };
        if (seam) body(std::integral_constant<int, 2>{});
        else if (live) body(std::integral_constant<int, 1>{});
        else body(std::integral_constant<int, 0>{});

        gchunk += (unsigned int)nch;
        if (__any_sync(0xffffffffu, assigned_any) && lane == 0) fast_mark_flags(a.self.tflags, a.lower.tflags, a.upper.tflags, a.lower.ny, a.upper.ny, a.nx, a.ny, a.nfbx, a.nfby,
                            strip * G::OUTC, y0, G::OUTC, y1 - y0);
    }
    signal_neighbours(a);
}

}  // namespace kob
#endif  // KOB_FAST_CUH
