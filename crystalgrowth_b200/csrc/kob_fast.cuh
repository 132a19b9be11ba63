// kob_fast.cuh — FAST fused Kobayashi step (FP32): the roofline kernel.  One launch = one explicit-Euler step
// = pass 1 + pass 2 of the reference (src/Kobayashi.cpp:125-175, :177-221), no scratch arrays in HBM.
//
// Same model as the STRICT kernel (dead-band angle state machine with carried theta, PI_F, 9-point Laplacians,
// Jacobi update); differences are rounding-level only (kob_row.cuh: reciprocal multiplies for the divisions by loop
// constants, FMA contraction, re-associated Laplacian sums, trig-free anisotropy, minimax atan).
//
// Structure (B200-first):
//   * persistent grid; a CTA claims 8 adjacent column strips of one row segment from a global counter and hands one
//     strip to each warp, which then works on its own (TMA ring, registers, stores).  Far-field CTA jobs keep their
//     warps in lock-step (one barrier per chunk) so that a grid row is fetched as 1920 contiguous bytes; jobs with
//     data-dependent work free-run.  (cta_jobs = 0 falls back to per-warp jobs with no barrier at all.)
//   * every warp owns a ring of NST shared-memory stages fed by TMA (cp.async.bulk.tensor.2d + mbarrier
//     complete_tx): a stage carries RB rows of phi and the RB rows of T one row behind it.  Loads are issued
//     NST chunks ahead by one lane; no LSU instruction or register is spent on input traffic.
//   * the warp MARCHES in y: lane L owns 2 adjacent cells of a row as one packed float2; all vertical neighbours (phi, T,
//     eps^2, eps*eps'*gx ...) are earlier rows kept in registers (RowState); horizontal neighbours of phi/T come from the
//     stage (LDS), horizontal neighbours of the pass-1 products from the adjacent lanes (SHFL).  Lane 0 and
//     lane 31 are halo lanes: they compute pass 1 for the strip's neighbours' edge cells and store nothing.
//   * results leave through coalesced 8-byte global stores; the cells on the strip/torus seams are stored to
//     every alias (own ghost columns, neighbour strips' ghost rows — peer memory over NVLink when P > 1).
//   * theta traffic is predicated: read only inside blocks flagged "theta may be non-zero", written only where the
//     state machine re-assigns it.
//   * two tiers per 4-row chunk: (1) far-field shortcut — phi rows (and the 4 rows before them) all +0: only T
//     diffuses; (2) per-row vote — the data-dependent block runs for the rows that need it.  Same bits either way (the
//     shortcut only skips work that would produce +0 / constants).
//   * a CTA that met data-dependent work stops claiming CTA-wide jobs and lets every warp claim its own (no block-level
//     barrier any more: in developed fields rows take unequal time and the barriers cost ~10 %).
#ifndef KOB_FAST_CUH
#define KOB_FAST_CUH

#include <cuda.h>   // CUtensorMap (type only; the encode entry point is fetched at run time, no -lcuda)

#include <type_traits>

#include "kob_common.cuh"
#include "kob_row.cuh"

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
    int free_mode;                 // 1 (default): a CTA that met data-dependent work switches to per-warp claims; 0 never
    RowConst rc;                   // stencil constants
    ColdK ck;                      // constants of the data-dependent block
    uint32_t pk[20];               // Philox round keys: pk[2r] = seed_lo + r*W0, pk[2r+1] = seed_hi + r*W1
    uint32_t pc2, pc3;             // Philox counter words 2, 3 = (step_lo, step_hi)
    long long ny_global;           // rows of the whole torus (two-step kernel: noise of wrapped ghost rows)
    // two-step kernel, general pass: jobs come from the work list the far pass wrote (nullptr: all jobs, from job_ctr)
    // (header words before the list, LH_*: see kob_fast2.cuh)
    int* list;
    const unsigned int* list_count;
    unsigned int* list_claim;
    unsigned int list_cap;         // entries the list can hold
    int list_conc;                 // this general pass was launched BEFORE its far pass and runs beside it (tickets are waited for)
    int list_drain;                // ... and stays until every listed range is served (0: leaves when the far pass is done; the
                                   // closing launch, a full grid then, serves the rest)
    int list_rearm;                // this launch is the last of the pair: its last warp re-arms the header
    int list_hot_par;              // Far2Args::hot_par of the pair
    unsigned int* live_ctr;        // single-step kernel, probe launches only: counts the jobs that see live theta flags
    // linked strips: seam readiness is published per side as soon as the jobs touching that seam are done
    unsigned int* seam_ctr;        // [2]: low-side / high-side jobs completed in this launch (reset by the publisher)
    unsigned int seam_jobs[2];     // units of this launch (pair) that touch the low / high seam: jobs, or row ranges (two-step)
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

// Noise draw r - 1/2 for the lane's two cells (x, x+1) of global row y.  Lanes (2m+1, 2m+2) share a Philox block (4 cells);
// over a row pair the low lane draws the block of row y, the high lane the block of row y+1, and they swap halves with two
// SHFL: one block per lane per two rows.  `even` = first row of such a pair (every lane of the warp takes the same path).
__device__ __forceinline__ float2 fast_draw_shared(const FastArgs& f, RowState& S, int x, uint32_t yglob, uint32_t pc2, uint32_t pc3,
                                                   bool even, int lane) {
    const bool hi = (x & 2) != 0;
    uint32_t wa, wb;
    if (even) {
        const Philox4 ph = fast_philox(f, (uint32_t)x >> 2, yglob + (hi ? 1u : 0u), pc2, pc3);
        const int partner = hi ? lane - 1 : lane + 1;
        const uint32_t ra = __shfl_sync(0xffffffffu, hi ? ph.w[0] : ph.w[2], partner);
        const uint32_t rb = __shfl_sync(0xffffffffu, hi ? ph.w[1] : ph.w[3], partner);
        wa = hi ? ra : ph.w[0]; wb = hi ? rb : ph.w[1];
        S.nxa = hi ? ph.w[2] : ra; S.nxb = hi ? ph.w[3] : rb;
        S.have_next = true;
    } else if (S.have_next) {
        wa = S.nxa; wb = S.nxb;
    } else {
        const Philox4 ph = fast_philox(f, (uint32_t)x >> 2, yglob, pc2, pc3);
        wa = hi ? ph.w[2] : ph.w[0]; wb = hi ? ph.w[3] : ph.w[1];
    }
    return f2fma(make_float2((float)(wa >> 8), (float)(wb >> 8)), f2(5.9604644775390625e-8f), f2(-0.5f));
}

struct FastGeom {
    static constexpr int CPL = 2;                 // cells per lane
    static constexpr int WCOLS = 32 * CPL;        // pass-1 columns per warp
    static constexpr int OUTC = WCOLS - 2 * CPL;  // output columns per strip (lanes 1..30)
    static constexpr int BW = WCOLS + 2 * CPL;    // TMA box width (own cells of lane L at box column CPL*L + CPL)
};

#ifndef KOB_FAST_RB
#define KOB_FAST_RB 4
#endif
#ifndef KOB_FAST_WARPS
#define KOB_FAST_WARPS 8
#endif
#ifndef KOB_FAST_CTAS
#define KOB_FAST_CTAS 2
#endif
constexpr int FAST_RB = KOB_FAST_RB;     // rows per TMA chunk (= unroll of the row loop)
static_assert(FAST_RB >= 4, "the far-field shortcut needs a chunk to cover the 4-row history of the register windows");
#ifndef KOB_ROW_LOOP
#define KOB_ROW_LOOP 4      // rows of a chunk per loop iteration of the row update (4: the chunk is unrolled whole)
#endif
constexpr int FAST_WARPS = KOB_FAST_WARPS;

// an input box (4 rows x 68 columns), padded to a multiple of 128 bytes (TMA shared-memory destination alignment)
constexpr int FAST_BOX_FLOATS = (FAST_RB * FastGeom::BW + 31) / 32 * 32;

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

// The same with pre-computed 32-bit shared addresses (callers make them warp-uniform with a shuffle first: ptxas then moves each
// operand to a uniform register once, instead of wrapping every UTMALDG in an ELECT / R2UR.BROADCAST loop).
__device__ __forceinline__ void tma_load_2d_raw(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(map), "r"(x), "r"(y), "r"(bar)
        : "memory");
}
__device__ __forceinline__ int warp_uniform(int v) { return __shfl_sync(0xffffffffu, v, 0); }

// ---- rare paths, kept out of line so that the steady-state row loop stays small in the instruction cache ----
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

// Linked strips: a warp that finished work touching the low (side 0) / high (side 1) seam counts it (`n` units: 1 per job in
// the single-step kernel, 1 per row range in the two-step launch pair); the warp that completes the side's last unit of this
// launch (pair) publishes "epoch + sub" to that neighbour — its ghost rows are written and the ghost rows this strip read from
// it are no longer needed — without waiting for the rest of the grid to drain.
__device__ __forceinline__ void fast_seam_done(const StepArgs<float>& a, const FastArgs& f, bool low, bool high, unsigned int n,
                                               unsigned int sub, int lane) {
    __syncwarp();
    if (lane == 0 && n != 0u) {
        __threadfence_system();
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            if (!(side ? high : low)) continue;
            if (atomicAdd(&f.seam_ctr[side], n) + n == f.seam_jobs[side]) {
                f.seam_ctr[side] = 0u;
                __threadfence_system();
                st_release_sys(side ? &a.upper.arrive[0] : &a.lower.arrive[1], a.epoch + sub);
            }
        }
    }
}

// ---- shared-memory plan of one warp of kob_step_fast (FAST_WARP_REGION bytes) ------------------------------------------------
// Two layouts over the same bytes; a job uses one of them from its first TMA load to its last TMA store:
//   ring 0 (lean and seam jobs):  4 input stages [phi box | T box]; results leave by coalesced 8-byte stores from registers
//   ring 1 (live jobs)         :  3 input stages [phi box | T box | theta box]  + 1 output buffer  [phi+ | T+ | theta]
// Input boxes are 4 rows x 68 columns (they must start on a 16-byte boundary in global memory: box column 0 is cell
// 60*strip - 4), output boxes 4 rows x 60 columns, each padded to a multiple of 128 bytes (TMA shared-memory alignment).  The lean path is HBM bound and wants the deep ring; live jobs are issue bound and want the
// theta rows delivered by TMA instead of by LSU instructions and registers.
constexpr int FAST_OUTC = FastGeom::OUTC;
constexpr int FAST_OBOX_BYTES = (FAST_RB * FAST_OUTC * 4 + 127) / 128 * 128;           // 1024
constexpr int FAST_BOX_BYTES = FAST_BOX_FLOATS * 4;                                    // 1152
constexpr int FAST_R0_NST = 4, FAST_R1_NST = 3;
constexpr int FAST_R0_NOBUF = 0, FAST_R1_NOBUF = 1;   // output buffers (ring 1: the store of chunk c has been read by the time chunk c+1 writes)
constexpr int FAST_R0_STAGE = 2 * FAST_BOX_BYTES, FAST_R1_STAGE = 3 * FAST_BOX_BYTES;
constexpr int FAST_R0_OUT = FAST_R0_NST * FAST_R0_STAGE, FAST_R1_OUT = FAST_R1_NST * FAST_R1_STAGE;
constexpr int FAST_R0_OBUF = 2 * FAST_OBOX_BYTES, FAST_R1_OBUF = 3 * FAST_OBOX_BYTES;
constexpr int FAST_R0_BYTES = FAST_R0_OUT + FAST_R0_NOBUF * FAST_R0_OBUF, FAST_R1_BYTES = FAST_R1_OUT + FAST_R1_NOBUF * FAST_R1_OBUF;
constexpr int FAST_WARP_REGION = FAST_R0_BYTES > FAST_R1_BYTES ? FAST_R0_BYTES : FAST_R1_BYTES;
constexpr int FAST_NBARS = FAST_R0_NST + FAST_R1_NST;
static_assert(FAST_R0_OUT % 128 == 0 && FAST_R1_OUT % 128 == 0 && FAST_WARP_REGION % 128 == 0, "TMA boxes must sit on 128-byte boundaries");
__host__ __device__ constexpr int fast_smem_bytes(int warps) { return warps * FAST_WARP_REGION + warps * FAST_NBARS * 8; }

// TMA descriptors of the single-step kernel: inputs (68-wide phi / T / theta boxes) and outputs (60-wide boxes);
// one per ping-pong buffer (phi, T) / per theta buffer.
struct FastMapsIO {
    CUtensorMap phi_in[2], t_in[2], th_in[2];
    CUtensorMap phi_out[2], t_out[2], th_out[2];
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int x, int y, const void* src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(x), "r"(y),
                 "r"(smem_u32(src))
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// JM: 4 / 6 = compile-time integer mode, 0 = run-time integer mode (ck.jmode in 0..16), -1 = any real j (trig).
// NOISE: 0 = off, 1 = Philox stream, 2 = host-injected field (a.noise_field; parity option, JM <= 0 instantiations only).
template <int JM, int NOISE, bool ROT>
__global__ void __launch_bounds__(32 * FAST_WARPS, KOB_FAST_CTAS) kob_step_fast(const __grid_constant__ FastMapsIO maps, const StepArgs<float> a,
                                                                            const FastArgs f) {
    using G = FastGeom;
    constexpr int CPL = G::CPL, BW = G::BW, RB = FAST_RB, OUTC = G::OUTC;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // the warp index through a shuffle: tells the compiler it is warp-uniform, so that everything derived from it (shared-memory
    // stage addresses, barriers, job geometry, TMA operands) can live in uniform registers
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    unsigned char* region = smem_raw + (size_t)warp * FAST_WARP_REGION;
    uint64_t* bars0 = reinterpret_cast<uint64_t*>(smem_raw + (size_t)nwarps * FAST_WARP_REGION) + warp * FAST_NBARS;
    uint64_t* bars1 = bars0 + FAST_R0_NST;

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < FAST_NBARS; ++s) mbar_init(&bars0[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    const int tcur = a.tcur;                                 // which theta buffer is current (the two-step kernel flips them)
    const CUtensorMap* map_phi = &maps.phi_in[a.cur];
    const CUtensorMap* map_t = &maps.t_in[a.cur];
    const CUtensorMap* map_th = &maps.th_in[tcur];
    const CUtensorMap* omap_phi = &maps.phi_out[a.cur ^ 1];
    const CUtensorMap* omap_t = &maps.t_out[a.cur ^ 1];
    const CUtensorMap* omap_th = &maps.th_out[tcur];
    float* __restrict__ phi_out = a.self.phi[a.cur ^ 1];
    float* __restrict__ t_out = a.self.t[a.cur ^ 1];
    const long long pitch = a.pitch;
    unsigned int gch0 = 0, gch1 = 0;   // chunks consumed so far per ring: stage = g % NST, parity = (g / NST) & 1

    // Job claims.  Jobs are numbered (segment, strip) with the strips padded to a multiple of the CTA's warps; the global
    // counter is ALWAYS advanced by a whole group of `nwarps` adjacent strips of one segment, so that (a) the warps of a
    // lock-step group share the segment, hence the chunk count, and (b) every CTA overshoots the queue by exactly one group —
    // the host predicts the counter (job_base of the next launch) from that.  CTA mode hands a group out behind a block
    // barrier; FREE mode (after the CTA met data-dependent work) hands the slots of a group out to whichever warp comes
    // first, without any barrier: the warp that takes slot 0 fetches the group and publishes it in a small ring.
    __shared__ unsigned long long s_job;
    __shared__ unsigned long long s_gbase[8];
    __shared__ unsigned int s_gid[8];
    __shared__ unsigned int s_taken;
    if (threadIdx.x < 8) s_gid[threadIdx.x] = 0u;
    if (threadIdx.x == 0) s_taken = 0u;
    __syncthreads();
    const int nsp = f.cta_jobs ? f.nstrips_p : f.nstrips;     // strips per segment in the job numbering
    const int njobs_q = nsp * f.nseg;
    bool cta_claims = f.cta_jobs != 0;                       // CTA mode (block-uniform; may switch to FREE mode once)
    int busy_run = 0;                                        // consecutive CTA-mode groups with data-dependent work
    for (;;) {
        unsigned long long jraw = 0;
        if (cta_claims) {
            __syncthreads();
            if (threadIdx.x == 0) s_job = atomicAdd(f.job_ctr, (unsigned long long)nwarps) - f.job_base;
            __syncthreads();
            jraw = s_job + (unsigned long long)warp;
            if (s_job >= (unsigned long long)njobs_q) break;
        } else if (f.cta_jobs) {                             // FREE mode
            if (lane == 0) {
                const unsigned int slot = atomicAdd(&s_taken, 1u), g = slot / (unsigned int)nwarps, k = slot - g * (unsigned int)nwarps;
                volatile unsigned int* gid = &s_gid[g & 7u];
                volatile unsigned long long* gb = &s_gbase[g & 7u];
                if (k == 0u) {
                    *gb = atomicAdd(f.job_ctr, (unsigned long long)nwarps) - f.job_base;
                    __threadfence_block();
                    *gid = g + 1u;
                } else {
                    while (*gid != g + 1u) __nanosleep(20);
                    __threadfence_block();
                }
                jraw = *gb + (unsigned long long)k;
            }
            jraw = __shfl_sync(0xffffffffu, jraw, 0);
            if (jraw >= (unsigned long long)njobs_q) break;
        } else {                                             // narrow grids / KOB_FAST_CTA=0: every warp claims single jobs
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
        const int xs = strip * OUTC - CPL;               // first pass-1 column of the warp (lane 0, halo)
        const int x = xs + CPL * lane;                   // first cell of this lane
        const bool mid_lane = lane >= 1 && lane <= 30;
        const bool touch_low = y0 < GY + 1, touch_high = y1 > a.ny - GY - 1;

        if (a.linked) {
            if (lane == 0) {
                if (touch_low) wait_flag(&a.self.arrive[0], a.epoch, &a.self.arrive[2], 1u);
                if (touch_high) wait_flag(&a.self.arrive[1], a.epoch, &a.self.arrive[2], 1u);
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
        // seam job: touches the first/last GXR columns or GY rows (alias stores, ragged right edge), or its height is not a
        // multiple of the chunk (the TMA output boxes of the interior paths are whole chunks) -> generic LSU path
        const bool seam = strip == 0 || (strip + 1) * OUTC > a.nx - GXR || y0 < GY || y1 > a.ny - GY || ((y1 - y0) & (RB - 1)) != 0;

        const int nrows = (y1 - y0) + 4;                 // streamed phi rows y0-2 .. y1+1
        const int nch = (nrows + RB - 1) / RB;
        // CTA-wide jobs: the 8 adjacent strips advance in lock-step (one barrier per chunk) so that a row is fetched as
        // 1920 contiguous bytes.  cta_jobs == 2: only while none of the 8 does data-dependent work (live theta / seam) —
        // there the rows take unequal time and the barrier would only add waiting.
        bool lock = cta_claims && f.cta_jobs == 1;
        if (cta_claims && f.cta_jobs == 2) {
            const bool busy = __syncthreads_or((live || seam) && strip < f.nstrips);
            lock = !busy;
            // developed field (two busy groups in a row — isolated crystals in a sparse field do not do that): from the next
            // claim on the warps of this CTA take their slots without a barrier (FREE mode; same counter, same groups)
            busy_run = busy ? busy_run + 1 : 0;
            if (busy_run >= 2 && f.free_mode) cta_claims = false;
        }
        if (strip >= f.nstrips) {                        // padding job: only keep the CTA's barriers company
            if (lock)
                for (int c = 0; c < nch; ++c) __syncthreads();
            continue;
        }
        const int box_x = xs - CPL + GX;                 // padded x of the phi / T box column 0
        bool assigned_any = false;
        const unsigned int nvalid = (unsigned int)(y1 - y0);
        const unsigned int nstore = mid_lane ? nvalid : 0u;    // rows this lane stores

        // ---- interior jobs (MODE 0 lean: theta all zero in the footprint; MODE 1 live: held theta is read): every byte in and
        // out of the warp moves by TMA.  Chunk c consumes phi rows y0-2+4c .. +3 and produces rows y0+4(c-1) .. +3. ----
        auto body_io = [&](auto mode_tag) {
            constexpr int MODE = decltype(mode_tag)::value;
            constexpr bool GEN = MODE == 1;
            constexpr int NST = GEN ? FAST_R1_NST : FAST_R0_NST, STAGE = GEN ? FAST_R1_STAGE : FAST_R0_STAGE;
            constexpr int OUT0 = GEN ? FAST_R1_OUT : FAST_R0_OUT, OBUF = GEN ? FAST_R1_OBUF : FAST_R0_OBUF;
            uint64_t* bars = GEN ? bars1 : bars0;
            unsigned int& gchunk = GEN ? gch1 : gch0;
            const int box_xu = warp_uniform(box_x), ybase = warp_uniform(y0 - 2 + GY);
            const unsigned int gchunk_u = (unsigned int)warp_uniform((int)gchunk);
            const uint32_t region32 = (uint32_t)warp_uniform((int)smem_u32(region)), bars32 = (uint32_t)warp_uniform((int)smem_u32(bars));
            auto issue = [&](int c) {                    // chunk c of this job -> stage ((gchunk + c) % NST); called by the whole warp
                const int cu = warp_uniform(c);
                const unsigned int gi = gchunk_u + (unsigned int)cu;
                const int st = gi % NST;
                const uint32_t dst = region32 + st * STAGE, bar = bars32 + st * 8;
                const int yr = ybase + cu * RB;          // padded row of the chunk's first phi row
                if (lane == 0) {
                    mbar_expect_tx(&bars[st], (GEN ? 3 : 2) * RB * BW * 4);
                    tma_load_2d_raw(dst, map_phi, box_xu, yr, bar);
                    tma_load_2d_raw(dst + FAST_BOX_BYTES, map_t, box_xu, yr - 1, bar);
                    if (GEN) tma_load_2d_raw(dst + 2 * FAST_BOX_BYTES, map_th, box_xu, yr - 1, bar);   // theta rows = the T rows
                }
            };
            for (int c = 0; c < NST && c < nch; ++c) issue(c);
            RowState S;
            S.clear();
            bool prevz = false;                          // !GEN: the previous chunk's phi rows were all +0
            float2 th_hold = f2(0.f);                    // GEN: angle of the row produced by the last iteration of the previous chunk
            const int lane_out = (lane - 1) * CPL * 4;   // byte offset of this lane's pair in an output row (mid lanes)
            // !GEN: running pointers to cell (x, r-2) of the output arrays (coalesced 8-byte stores straight from registers)
            const long long o2 = pidx<float>(pitch, x, y0 - 4);
            float* pphi = phi_out + o2;
            float* ptt = t_out + o2;

            for (int c = 0; c < nch; ++c) {
                const unsigned int gi = gchunk + (unsigned int)c;
                const int st = gi % NST;
                if (lock) __syncthreads();                                          // adjacent strips advance together
                mbar_wait(&bars[st], (gi / NST) & 1u);
                const unsigned char* stage = region + st * STAGE;
                const float* sp = reinterpret_cast<const float*>(stage) + CPL * lane + CPL;   // this lane's own phi cells
                const float* stt = sp + FAST_BOX_FLOATS;                            // T rows (one row behind)
                const float* sth = stt + FAST_BOX_FLOATS;                           // GEN: theta rows (same rows as T)
                unsigned char* obuf = region + OUT0;                                // GEN: output boxes of this chunk
                const bool store = c > 0;                                           // chunk 0 only warms the windows up
                const int yrel0 = c * RB - 4;                                       // (r - 2) - y0 for rr = 0
                if (GEN && c > 1) {                      // single output buffer: the previous chunk's boxes must have been read
                    if (lane == 0) tma_store_wait_read<0>();
                    __syncwarp();
                }
                if (GEN && store && mid_lane) *reinterpret_cast<float2*>(obuf + 2 * FAST_OBOX_BYTES + lane_out) = th_hold;
                bool skipped = false;
                if (!GEN) {
                    // ---- far field: phi == +0 on all 64 own columns of this chunk's RB rows AND of the previous RB
                    // rows.  Every phi term of the step is then exactly +0 and the register windows already sit at
                    // their all-zero-input fixed point (each is a function of the last <= 5 streamed rows only), so
                    // the chunk reduces to the T diffusion; outputs are bit-identical to the full path. ----
                    uint32_t bits = 0u;
#pragma unroll
                    for (int rr = 0; rr < RB; ++rr) bits |= __float_as_uint(sp[rr * BW]) | __float_as_uint(sp[rr * BW + 1]);
                    const bool curz = !__any_sync(0xffffffffu, bits != 0u);
                    skipped = curz && prevz && !f.no_skip;
                    prevz = curz;
                    if (skipped) {
#pragma unroll
                        for (int rr = 0; rr < RB; ++rr) {
                            const float* trow = stt + rr * BW;
                            const float2 nt_ = row_tonly(S, f.rc, *reinterpret_cast<const float2*>(trow), trow[-1], trow[CPL]);
                            if ((unsigned int)(yrel0 + rr) < nstore) {
                                *reinterpret_cast<float2*>(pphi) = f2(0.f);
                                *reinterpret_cast<float2*>(ptt) = nt_;
                            }
                            pphi += pitch;
                            ptt += pitch;
                        }
                    }
                }
                if (!skipped) {
                    // (the chunk is unrolled whole: a loop over row pairs halves the code — and the instruction-fetch stalls that
                    // small grids with short jobs show — but ptxas then spends ~50 register moves per row at the back edge)
                    auto one_row = [&](const int rr, auto ro_tag) {
                        constexpr int ro = decltype(ro_tag)::value;
                        const unsigned int yrel = (unsigned int)(yrel0 + rr);       // row of pass 2, relative to y0
                        const float* prow = sp + rr * BW;
                        const float* trow = stt + rr * BW;
                        const float2 pn = *reinterpret_cast<const float2*>(prow);
                        const float2 tn = *reinterpret_cast<const float2*>(trow);
                        float th_in[CPL] = {0.f, 0.f};
                        if (GEN) {
                            const float2 v = *reinterpret_cast<const float2*>(sth + rr * BW);
                            th_in[0] = v.x; th_in[1] = v.y;
                        }
                        float2 np_, nt_, th2 = f2(0.f);
                        bool asg[CPL];
                        const int y = y0 + (int)yrel;
                        auto draw = [&]() -> float2 {
                            if (NOISE == 2) {
                                float2 rq;
                                rq.x = (yrel < nvalid ? __ldg(&a.noise_field[(long long)x + (long long)a.nx * y]) : 0.5f) - 0.5f;
                                rq.y = (yrel < nvalid ? __ldg(&a.noise_field[(long long)(x + 1) + (long long)a.nx * y]) : 0.5f) - 0.5f;
                                return rq;
                            }
                            return fast_draw_shared(f, S, x, (uint32_t)(a.y0 + y), f.pc2, f.pc3, ro == 0, lane);
                        };
                        if (ro == 0) S.have_next = false;
                        const bool vote = row_full<ro, JM, NOISE != 0, ROT, GEN>(S, f.rc, f.ck, pn, prow[-1], prow[CPL], tn, trow[-1], trow[CPL],
                                                                                 th_in, draw, np_, nt_, th2, asg);
                        if (GEN) {
                            if (store && mid_lane) {
                                *reinterpret_cast<float2*>(obuf + rr * (OUTC * 4) + lane_out) = np_;
                                *reinterpret_cast<float2*>(obuf + FAST_OBOX_BYTES + rr * (OUTC * 4) + lane_out) = nt_;
                            }
                        } else {
                            if (yrel < nstore) {
                                *reinterpret_cast<float2*>(pphi) = np_;
                                *reinterpret_cast<float2*>(ptt) = nt_;
                            }
                            pphi += pitch;
                            ptt += pitch;
                        }
                        if (GEN) {
                            // angle of row r-1 after this step: re-assigned or kept.  Rows are written back whole (a held cell gets
                            // the bits it had), one row ahead of phi/T: the last row of a chunk waits in th_hold for the next box.
                            const float2 thv = make_float2((vote && asg[0]) ? th2.x : th_in[0], (vote && asg[1]) ? th2.y : th_in[1]);
                            if (rr < RB - 1) {
                                if (store && mid_lane) *reinterpret_cast<float2*>(obuf + 2 * FAST_OBOX_BYTES + (rr + 1) * (OUTC * 4) + lane_out) = thv;
                            } else {
                                th_hold = thv;
                            }
                            if (vote && yrel + 1u < nstore) assigned_any |= asg[0] || asg[1];
                        } else if (vote && yrel + 1u < nstore) {
                            // lean job: theta is zero in the whole footprint and is only written where re-assigned (rare: the front
                            // of a crystal entering a far-field job)
                            float* pth = a.self.theta + pidx<float>(pitch, x, y + 1);
                            if (asg[0]) pth[0] = th2.x;
                            if (asg[1]) pth[1] = th2.y;
                            assigned_any |= asg[0] || asg[1];
                        }
                    };
#if KOB_ROW_LOOP == 2
#pragma unroll 1
                    for (int rp = 0; rp < RB; rp += 2) {
                        one_row(rp, std::integral_constant<int, 0>{});
                        one_row(rp + 1, std::integral_constant<int, 1>{});
                    }
#else
                    one_row(0, std::integral_constant<int, 0>{});
                    one_row(1, std::integral_constant<int, 1>{});
                    one_row(2, std::integral_constant<int, 0>{});
                    one_row(3, std::integral_constant<int, 1>{});
#endif
                }
                // ---- live jobs: this chunk's output boxes leave by TMA; the next chunk's input is requested ----
                if (GEN && store) fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                    if (GEN && store) {
                        const int ox = strip * OUTC + GX, oy = y0 + (c - 1) * RB + GY;
                        tma_store_2d(omap_phi, ox, oy, obuf);
                        tma_store_2d(omap_t, ox, oy, obuf + FAST_OBOX_BYTES);
                        tma_store_2d(omap_th, ox, oy, obuf + 2 * FAST_OBOX_BYTES);
                        tma_store_commit();
                    }
                }
                if (c + NST < nch) issue(c + NST);
                if (GEN) __syncwarp();
            }
            gchunk += (unsigned int)nch;
            if (GEN) {
                if (lane == 0) tma_store_wait_read<0>();     // the next job may lay the region out differently
                __syncwarp();
            }
        };

        // ---- seam jobs (MODE 2): the ragged right edge, the alias stores to ghost columns / neighbour strips, arbitrary job
        // heights; theta by LSU loads two rows ahead and predicated stores.  Ring 0 input stages. ----
        auto body_seam = [&]() {
            constexpr int NST = FAST_R0_NST, STAGE = FAST_R0_STAGE;
            uint64_t* bars = bars0;
            unsigned int& gchunk = gch0;
            auto issue = [&](int c) {
                const unsigned int gi = gchunk + (unsigned int)c;
                const int st = gi % NST;
                unsigned char* dst = region + st * STAGE;
                mbar_expect_tx(&bars[st], 2 * RB * BW * 4);
                const int yr = y0 - 2 + c * RB + GY;
                tma_load_2d(dst, map_phi, box_x, yr, &bars[st]);
                tma_load_2d(dst + FAST_BOX_BYTES, map_t, box_x, yr - 1, &bars[st]);
            };
            if (lane == 0) {
                for (int c = 0; c < NST && c < nch; ++c) issue(c);
            }
            RowState S;
            S.clear();
            // theta of the held cells, prefetched two rows ahead: thp0 = theta(r-1), thp1 = theta(r)
            float thp0[CPL] = {0.f, 0.f}, thp1[CPL] = {0.f, 0.f};
            // running pointers to cell (x, r-2) of the output arrays / theta
            const long long o2 = pidx<float>(pitch, x, y0 - 4);
            float* pphi = phi_out + o2;
            float* ptt = t_out + o2;
            float* pthe = a.self.theta + (o2 + pitch);   // theta of cell (x, r-1): the row pass 1 re-assigns
            const long long pitch2 = 2 * pitch;
            for (int c = 0; c < nch; ++c) {
                const unsigned int gi = gchunk + (unsigned int)c;
                const int st = gi % NST;
                mbar_wait(&bars[st], (gi / NST) & 1u);
                const float* sp = reinterpret_cast<const float*>(region + st * STAGE) + CPL * lane + CPL;
                const float* stt = sp + FAST_BOX_FLOATS;
                const int yrel0 = c * RB - 4;
                bool lrow_c = false;          // some theta-flag row under this chunk's prefetch rows (r+1) is live
                if (live) {
                    const int f0 = ((y0 + yrel0 + 3 + GY) >> 5) - fby0, f1 = ((y0 + yrel0 + RB + 2 + GY) >> 5) - fby0;   // FBY == 32
                    lrow_c = ((livemask >> min(max(f0, 0), 31)) | (livemask >> min(max(f1, 0), 31))) & 1u;
                }
#pragma unroll
                for (int rr = 0; rr < RB; ++rr) {
                    const unsigned int yrel = (unsigned int)(yrel0 + rr);
                    const float* prow = sp + rr * BW;
                    const float* trow = stt + rr * BW;
                    const float2 pn = *reinterpret_cast<const float2*>(prow);
                    const float2 tn = *reinterpret_cast<const float2*>(trow);
                    float2 np_, nt_, th2 = f2(0.f);
                    bool asg[CPL];
                    const int y = y0 + (int)yrel;
                    auto draw = [&]() -> float2 {
                        if (NOISE == 2) {
                            float2 rq;
                            rq.x = ((x >= 0 && x < a.nx && yrel < nvalid) ? __ldg(&a.noise_field[(long long)x + (long long)a.nx * y]) : 0.5f) - 0.5f;
                            rq.y = ((x + 1 >= 0 && x + 1 < a.nx && yrel < nvalid) ? __ldg(&a.noise_field[(long long)(x + 1) + (long long)a.nx * y]) : 0.5f) - 0.5f;
                            return rq;
                        }
                        return fast_draw_shared(f, S, x, (uint32_t)(a.y0 + y), f.pc2, f.pc3, (rr & 1) == 0, lane);
                    };
                    if ((rr & 1) == 0) S.have_next = false;
                    const bool vote = (rr & 1) == 0 ? row_full<0, JM, NOISE != 0, ROT, true>(S, f.rc, f.ck, pn, prow[-1], prow[CPL], tn, trow[-1], trow[CPL],
                                                                                               thp0, draw, np_, nt_, th2, asg)
                                                    : row_full<1, JM, NOISE != 0, ROT, true>(S, f.rc, f.ck, pn, prow[-1], prow[CPL], tn, trow[-1], trow[CPL],
                                                                                               thp0, draw, np_, nt_, th2, asg);
                    // ---- store the re-assigned angles of owned cells of row r-1 ----
                    if (vote && yrel + 1u < nstore) {
#pragma unroll
                        for (int k = 0; k < CPL; ++k) {
                            if (asg[k] && x + k < a.nx) {
                                const float th = k ? th2.y : th2.x;
                                const int yt = y + 1;
                                if (yt < GY || yt >= a.ny - GY)
                                    fast_store_edge(a.self.theta, a.lower.theta, a.upper.theta, pitch, a.nx, a.ny, a.lower.ny, x + k, yt, th);
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
                    // ---- store phi+, T+ of row r-2 ----
                    if (yrel < nstore) {
                        if (y < GY || y >= a.ny - GY) {          // rows on the strip seam: every alias (rare)
#pragma unroll
                            for (int k = 0; k < CPL; ++k)
                                if (x + k < a.nx) {
                                    fast_store_edge(phi_out, a.lower.phi[a.cur ^ 1], a.upper.phi[a.cur ^ 1], pitch, a.nx, a.ny, a.lower.ny, x + k, y, k ? np_.y : np_.x);
                                    fast_store_edge(t_out, a.lower.t[a.cur ^ 1], a.upper.t[a.cur ^ 1], pitch, a.nx, a.ny, a.lower.ny, x + k, y, k ? nt_.y : nt_.x);
                                }
                        } else {                                 // interior rows: own cell + ghost-column copy
#pragma unroll
                            for (int k = 0; k < CPL; ++k)
                                if (x + k < a.nx) {
                                    const float vp = k ? np_.y : np_.x, vt = k ? nt_.y : nt_.x;
                                    pphi[k] = vp;
                                    ptt[k] = vt;
                                    if (x + k < GXR) { pphi[k + a.nx] = vp; ptt[k + a.nx] = vt; }
                                    if (x + k >= a.nx - GXR) { pphi[k - a.nx] = vp; ptt[k - a.nx] = vt; }
                                }
                        }
                    }
                    // ---- prefetch theta of row r+1 (pass-1 row of the iteration after next) ----
                    thp0[0] = thp1[0]; thp0[1] = thp1[1];
                    thp1[0] = thp1[1] = 0.f;
                    if (lrow_c && yrel + 4u <= nvalid + 1u) {                            // theta rows y0-1 .. y1
                        const float* pf = pthe + pitch2;                                 // theta(x, r+1)
#pragma unroll
                        for (int k = 0; k < CPL; ++k)
                            if (x + k < a.nx + GXR && x + k >= -GXR) thp1[k] = __ldg(pf + k);
                    }
                    pphi += pitch;
                    ptt += pitch;
                    pthe += pitch;
                }
                __syncwarp();
                if (lane == 0 && c + NST < nch) issue(c + NST);
            }
            gchunk += (unsigned int)nch;
        };
#ifdef KOB_DEV_ONLY_LIVE   // developer harness (scripts/dev/sass_lines.sh): only the live row loop, for SASS inspection
        body_io(std::integral_constant<int, 1>{});
#else
        if (seam) body_seam();
        else if (live) body_io(std::integral_constant<int, 1>{});
        else body_io(std::integral_constant<int, 0>{});
#endif
        if (__any_sync(0xffffffffu, assigned_any) && lane == 0) fast_mark_flags(a.self.tflags, a.lower.tflags, a.upper.tflags, a.lower.ny, a.upper.ny, a.nx, a.ny, a.nfbx, a.nfby,
                            strip * OUTC, y0, OUTC, y1 - y0);
        if (a.linked && (touch_low || touch_high)) fast_seam_done(a, f, touch_low, touch_high, 1u, 1u, lane);
    }
    if (lane == 0) tma_store_wait_all();                 // every output box has landed before the warp retires
}

}  // namespace kob
#endif  // KOB_FAST_CUH
