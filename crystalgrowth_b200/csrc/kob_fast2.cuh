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

// One full row of one level (kob_row.cuh's row_full plus this kernel's noise addressing).  Inputs: phi row r (own pair pn,
// west, east), T row r-1 (tn, tw, te), th_old_in = angle a cell of row r-1 keeps if the state machine holds it.  Outputs:
// phi+/T+ of row r-2, th_eff = angle of row r-1 after this sub-step, any_asg |= some cell of row r-1 re-assigned.
// `pc2/pc3` = Philox counter words (step) of this level, `yglob` = global row of pass 2 (r-2), `even` = first row of a row
// pair (Philox sharing), `wrap_noise` = seam job.
template <int JM, bool NOISE, bool ROT, bool GEN>
__device__ __forceinline__ void f2_row(RowState& S, const FastArgs& f, float2 pn, float w, float ee, float2 tn, float tw, float te,
                                       const float (&th_old_in)[2], uint32_t pc2, uint32_t pc3, int x, long long yglob, bool even,
                                       int lane, bool wrap_noise, int nx, long long nyg, float2& np_, float2& nt_,
                                       float (&th_eff)[2], bool& any_asg) {
    auto draw = [&]() -> float2 {
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
            return f2fma(make_float2((float)(wk[0] >> 8), (float)(wk[1] >> 8)), f2(5.9604644775390625e-8f), f2(-0.5f));
        }
        return fast_draw_shared(f, S, x, (uint32_t)yglob, pc2, pc3, even, lane);
    };
    if (even) S.have_next = false;
    bool asg[2];
    float2 th2;
    // `even` is a compile-time fact after the callers' row loops are unrolled: the branch folds away
    if (even) row_full<0, JM, NOISE, ROT, GEN>(S, f.rc, f.ck, pn, w, ee, tn, tw, te, th_old_in, draw, np_, nt_, th2, asg);
    else row_full<1, JM, NOISE, ROT, GEN>(S, f.rc, f.ck, pn, w, ee, tn, tw, te, th_old_in, draw, np_, nt_, th2, asg);
    th_eff[0] = asg[0] ? th2.x : (GEN ? th_old_in[0] : 0.f);
    th_eff[1] = asg[1] ? th2.y : (GEN ? th_old_in[1] : 0.f);
    any_asg |= asg[0] || asg[1];
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
//   Hot units first: a claim unit (a CTA job of 8 strips, or a job) that listed something in the previous pair is visited before
//   all others (`hot_prev`, one byte per unit, written as `hot_next` by the previous far pass), so that the general pass — launched
//   BEFORE this kernel on a few SMs of its own and running beside it (programmatic dependent launch) — gets its work list within the
//   first microseconds and is done long before the far pass has streamed the rest of the grid.
// Work-list header (unsigned words in front of the list; all zero = armed, except LH_MIN = ~0):
enum : int {
    LH_COUNT = 0,     // entries appended by the far pass of this pair
    LH_CLAIM = 1,     // next ticket of the general pass
    LH_GEXITS = 2,    // warps of the running general-pass launch that have left
    LH_FEXITS = 4,    // far-pass warps that have left
    LH_FDONE = 5,     // 1: the far pass is complete, LH_COUNT is final
    LH_HOTCLAIM = 6,  // far pass, hot units first: next entry of the previous pair's hot list
    LH_STATE = 7,     // 0 / 1 = a general pass launched before its far pass gave up waiting for it / 2 = the far pass has started
    LH_MIN = 8,       // smallest ticket a general-pass warp left with unserved
    LH_HOTCNT = 9,    // [2] lengths of the two hot lists (the one being read, the one being written: Far2Args::hot_par)
    LH_LAST = 12,     // count of the previous pair (host density probe)
    LH_WORDS = 16
};
struct Far2Args {
    int* list;                  // job ids for the general pass (-1 = empty slot)
    unsigned int* hdr;          // work-list header
    const unsigned char* hot_prev;       // [units] 1 = the unit listed something in the previous pair
    unsigned char* hot_next;
    const unsigned int* hotlist_prev;    // those units, compact (one claim = one unit: spread over the whole grid)
    unsigned int* hotlist_next;
    int hot_par;                         // hdr[LH_HOTCNT + hot_par] = length of hotlist_prev, [.. + 1 - hot_par] of hotlist_next
    int hot_first;
};

__device__ __forceinline__ unsigned int ld_volatile_u32(const unsigned int* p) { return *reinterpret_cast<const volatile unsigned int*>(p); }
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

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
    unsigned int gchunk = 0;
    __shared__ unsigned long long s_job;
    __shared__ int s_hot;
    const int nsp = f.cta_jobs ? f.nstrips_p : f.nstrips;     // strips per segment in the job numbering
    const int njobs_q = nsp * f.nseg;
    const unsigned int nunits = f.cta_jobs ? (unsigned int)(njobs_q / nwarps) : (unsigned int)njobs_q;   // claim units
    if (threadIdx.x == 0) atomicCAS(&w.hdr[LH_STATE], 0u, 2u);       // the far pass has started (a concurrent general pass stops doubting)
    int phase = w.hot_first ? 0 : 1;                          // 0: the units that were hot in the previous pair, 1: all others
    for (;;) {
        unsigned long long jraw = 0;
        unsigned int unit = 0;
        if (phase == 0) {                                     // one claim = one hot unit
            unsigned int hi = 0;
            if (f.cta_jobs) {
                __syncthreads();
                if (threadIdx.x == 0) {
                    const unsigned int k = atomicAdd(&w.hdr[LH_HOTCLAIM], 1u);
                    s_job = k < w.hdr[LH_HOTCNT + w.hot_par] ? (unsigned long long)w.hotlist_prev[k] : ~0ull;
                }
                __syncthreads();
                hi = (unsigned int)s_job;
            } else {
                if (lane == 0) {
                    const unsigned int k = atomicAdd(&w.hdr[LH_HOTCLAIM], 1u);
                    hi = k < w.hdr[LH_HOTCNT + w.hot_par] ? w.hotlist_prev[k] : 0xffffffffu;
                }
                hi = __shfl_sync(0xffffffffu, hi, 0);
            }
            if (hi >= nunits) { phase = 1; continue; }
            unit = hi;
            jraw = f.cta_jobs ? (unsigned long long)unit * nwarps + warp : (unsigned long long)unit;
        } else if (f.cta_jobs) {                             // a CTA claims 8 adjacent strips and keeps them in lock-step:
            __syncthreads();                                 // a grid row is then fetched as 8 x 224 contiguous bytes
            if (threadIdx.x == 0) {
                s_job = atomicAdd(f.job_ctr, (unsigned long long)nwarps) - f.job_base;
                s_hot = (w.hot_first && s_job < (unsigned long long)njobs_q) ? w.hot_prev[s_job / nwarps] : 0;
            }
            __syncthreads();
            if (s_job >= (unsigned long long)njobs_q) break;
            if (s_hot) continue;                             // done in phase 0
            jraw = s_job + (unsigned long long)warp;
            unit = (unsigned int)(s_job / nwarps);
        } else {
            int hot = 0;
            if (lane == 0) {
                jraw = atomicAdd(f.job_ctr, 1ull) - f.job_base;
                hot = (w.hot_first && jraw < (unsigned long long)njobs_q) ? w.hot_prev[jraw] : 0;
            }
            jraw = __shfl_sync(0xffffffffu, jraw, 0);
            hot = __shfl_sync(0xffffffffu, hot, 0);
            if (jraw >= (unsigned long long)njobs_q) break;
            if (hot) continue;
            unit = (unsigned int)jraw;
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
            __syncthreads_or(0);
            continue;
        }
        const int xs = strip * F2_OUTC - F2_HALO;
        const int x = xs + 2 * lane;
        // seam jobs store to the aliases as well (ghost columns / neighbour strips' ghost rows) and skip the ragged edge
        const bool seam = strip == 0 || (strip + 1) * F2_OUTC > a.nx - GXR || y0 < GY || y1 > a.ny - GY;
        if (a.linked) {
            if (lane == 0) {
                if (y0 < GY + 2) wait_flag(&a.self.arrive[0], a.epoch, &a.self.arrive[2], 2u);
                if (y1 > a.ny - GY - 1) wait_flag(&a.self.arrive[1], a.epoch, &a.self.arrive[2], 2u);
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
            RowState L1, L2;
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
                    const float2 t1 = row_tonly(L1, f.rc, tn, row[-1], row[2]);          // T^1 of row r-2
                    const float tw2 = __shfl_up_sync(0xffffffffu, t1_prev.y, 1);
                    const float te2 = __shfl_down_sync(0xffffffffu, t1_prev.x, 1);
                    const float2 t2 = row_tonly(L2, f.rc, t1_prev, tw2, te2);            // T^2 of row r-4
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
            if (lane == 0) pos = atomicAdd(&w.hdr[LH_COUNT], (unsigned int)__popc(need_out));
            pos = __shfl_sync(0xffffffffu, pos, 0);
            if (lane < F2_RANGES && ((need_out >> lane) & 1u))                       // (a ticket holder polls its slot: -1 = not yet)
                *reinterpret_cast<volatile int*>(&w.list[pos + __popc(need_out & ((1u << lane) - 1u))]) =
                    (sq * f.nstrips + strip) * F2_RANGES + lane;                     // unpadded job numbering
        }
        {                                                // remember which units listed something: they go first in the next pair
            const int any = f.cta_jobs ? __syncthreads_or(need_out != 0u) : (int)(need_out != 0u);
            if (f.cta_jobs ? threadIdx.x == 0 : lane == 0) {
                w.hot_next[unit] = (unsigned char)(any != 0);
                if (any) w.hotlist_next[atomicAdd(&w.hdr[LH_HOTCNT + 1 - w.hot_par], 1u)] = unit;
            }
        }
        // linked strips: the row ranges of a seam job this pass completed itself count towards the side's early publish;
        // the listed ones are counted by the general pass
        if (a.linked && (y0 < GY + 2 || y1 > a.ny - GY - 1))
            fast_seam_done(a, f, y0 < GY + 2, y1 > a.ny - GY - 1, (unsigned int)(F2_RANGES - __popc(need_out)), 2u, lane);
    }
    // The last warp of the grid to leave closes the list: ticket holders beyond the final count may go.  Then wait for the
    // general pass if it runs beside this kernel (it was launched first; this kernel must not complete before it: the next
    // launch in the stream reads what it writes).  A no-op in a plain launch.
    __syncwarp();
    if (lane == 0) {
        __threadfence();
        if (atomicAdd(&w.hdr[LH_FEXITS], 1u) + 1u == gridDim.x * (unsigned int)nwarps) {
            __threadfence();
            *reinterpret_cast<volatile unsigned int*>(&w.hdr[LH_FDONE]) = 1u;
        }
        asm volatile("griddepcontrol.wait;" ::: "memory");
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

    const CUtensorMap* map_phi = a.cur ? &maps.phi[1] : &maps.phi[0];
    const CUtensorMap* map_t = a.cur ? &maps.t[1] : &maps.t[0];
    float* __restrict__ phi_out = a.self.phi[a.cur ^ 1];
    float* __restrict__ t_out = a.self.t[a.cur ^ 1];
    const long long pitch = a.pitch;
    const uint32_t pc2b = (uint32_t)(a.step + 1ull), pc3b = (uint32_t)((a.step + 1ull) >> 32);   // level 2 = step + 1
    unsigned int gchunk = 0;
    if (f.list_conc) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the far pass may start beside this kernel

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
            // A warp draws a ticket and waits until the far pass has filled that slot, or has finished with fewer entries.  When
            // this kernel follows its far pass in the stream nothing is ever waited for; launched before it (list_conc) the
            // tickets are served while the far pass is still streaming the rest of the grid.
            unsigned int* hdr = f.list_claim - LH_CLAIM;
            unsigned int idx = 0;
            int entry = -1;
            if (lane == 0) {
                unsigned long long t0 = 0ull;
                for (;;) {                                                   // tickets until one is served or the list is over
                    idx = atomicAdd(f.list_claim, 1u);
                    entry = -1;
                    for (unsigned int spin = 1u;; ++spin) {
                        if (idx < f.list_cap) entry = *reinterpret_cast<volatile int*>(&f.list[idx]);
                        if (entry >= 0) break;
                        if (ld_volatile_u32(&hdr[LH_FDONE]) != 0u) {         // far pass done: the count is final, all slots are written
                            __threadfence();
                            if (idx < f.list_cap) entry = *reinterpret_cast<volatile int*>(&f.list[idx]);
                            if (entry < 0) entry = idx < ld_volatile_u32(&hdr[LH_COUNT]) ? -3 : -2;   // -3: served by an earlier launch
                            break;
                        }
                        const unsigned int st = ld_volatile_u32(&hdr[LH_STATE]);
                        if (st == 1u) { entry = -2; break; }                 // this pass gave up (below); the closing launch serves the list
                        __nanosleep(400);
                        if ((spin & 255u) == 0u) {
                            // No far pass in sight for 20 ms: kernels are being serialised (a profiler, a debugger) and it cannot start
                            // before this kernel ends.  Give up — all warps or none (the far pass moves 0 -> 2, this moves 0 -> 1).
                            const unsigned long long now = global_timer_ns();
                            if (t0 == 0ull) t0 = now;
                            else if (st == 0u && now - t0 > 20000000ull) {
                                if (atomicCAS(&hdr[LH_STATE], 0u, 1u) != 2u) { entry = -2; break; }
                            } else if (now - t0 > 20000000000ull) {          // 20 s beside a running far pass: it is stuck
                                atomicExch(&a.self.arrive[2], 3u);
                                entry = -2; break;
                            }
                        }
                    }
                    // beside the far pass without `drain`: leave as soon as the far pass is done, the closing launch has the whole GPU
                    if (entry >= 0 && f.list_conc && !f.list_drain && ld_volatile_u32(&hdr[LH_FDONE]) != 0u) entry = -2;
                    if (entry != -3) break;
                }
                if (entry >= 0) *reinterpret_cast<volatile int*>(&f.list[idx]) = -1;    // slot consumed: empty for the next pair
                else atomicMin(&hdr[LH_MIN], idx);                                       // the closing launch starts there
            }
            entry = __shfl_sync(0xffffffffu, entry, 0);
            if (entry < 0) {
                // The LAST warp of the grid to get here: a closing launch re-arms the header for the next launch pair (the count is kept
                // for the host's density probe; no memset per pair); otherwise the tickets restart at the first unserved one.
                if (lane == 0) {
                    if (atomicAdd(&hdr[LH_GEXITS], 1u) + 1u == gridDim.x * (blockDim.x >> 5)) {
                        if (f.list_rearm) {
                            hdr[LH_LAST] = hdr[LH_COUNT];
                            hdr[LH_COUNT] = 0u; hdr[LH_CLAIM] = 0u; hdr[LH_FEXITS] = 0u; hdr[LH_FDONE] = 0u; hdr[LH_HOTCLAIM] = 0u; hdr[LH_STATE] = 0u;
                            hdr[LH_HOTCNT + f.list_hot_par] = 0u;            // the hot list read in this pair is the one written in the next
                        } else {
                            hdr[LH_CLAIM] = hdr[LH_MIN];
                        }
                        hdr[LH_MIN] = 0xffffffffu;
                        hdr[LH_GEXITS] = 0u;
                        __threadfence();
                    }
                }
                break;
            }
            jraw = (unsigned long long)entry;
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
        const bool job_low = y0 < GY + 2, job_high = y1 > a.ny - GY - 1;   // the JOB touches a strip seam (early publish, in row ranges)
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
                if (y0 < GY + 2) wait_flag(&a.self.arrive[0], a.epoch, &a.self.arrive[2], 2u);
                if (y1 > a.ny - GY - 1) wait_flag(&a.self.arrive[1], a.epoch, &a.self.arrive[2], 2u);
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
            RowState L1, L2;
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
                            const float2 t1 = row_tonly(L1, f.rc, tn, row[-1], row[2]);          // T^1 of row r-2
                            const float tw2 = __shfl_up_sync(0xffffffffu, t1_prev.y, 1);
                            const float te2 = __shfl_down_sync(0xffffffffu, t1_prev.x, 1);
                            const float2 t2 = row_tonly(L2, f.rc, t1_prev, tw2, te2);            // T^2 of row r-4
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
                    f2_row<JM, NOISE, ROT, GEN>(L1, f, pn, row[-1], row[2], tn, trow[-1], trow[2], thp0, f.pc2, f.pc3, x,
                                                a.y0 + y0 + (int)yrel + 2, even, lane, GEN && seam, a.nx, f.ny_global, p1, t1, th1, dummy);
                    // ---- level 2: sub-step s+1 on phi^1 row r-2, T^1 row r-3 ----
                    const float pw = __shfl_up_sync(0xffffffffu, p1.y, 1);
                    const float pe = __shfl_down_sync(0xffffffffu, p1.x, 1);
                    const float tw2 = __shfl_up_sync(0xffffffffu, t1_prev.y, 1);
                    const float te2 = __shfl_down_sync(0xffffffffu, t1_prev.x, 1);
                    float2 p2, t2;
                    float th2f[2];
                    bool asg2 = false;
                    f2_row<JM, NOISE, ROT, true>(L2, f, p1, pw, pe, t1_prev, tw2, te2, e2, pc2b, pc3b, x,
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
        if (a.linked && real_job && (job_low || job_high))
            fast_seam_done(a, f, job_low, job_high, sub >= 0 ? 1u : (unsigned int)F2_RANGES, 2u, lane);
    }
}

}  // namespace kob
#endif  // KOB_FAST2_CUH
