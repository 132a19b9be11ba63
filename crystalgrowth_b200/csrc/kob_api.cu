// kob_api.cu — implementation of the C ABI in include/kobayashi_c.h (libkobayashi_cuda.so).
// Host side: context, HBM allocation, streams/events, strip linking (same process or CUDA IPC), launches.
// There is NO CPU fallback: every entry point that needs the device fails with KOB_ERR_NO_DEVICE /
// KOB_ERR_CUDA when it is missing.
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <thread>

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/kobayashi_c.h"
#include "kob_aux.cuh"
#include "kob_common.cuh"
#include "kob_fast.cuh"
#include "kob_fast2.cuh"
#include "kob_strict.cuh"

using namespace kob;

namespace {

thread_local std::string g_create_error;

struct Layout {
    size_t elem;
    long long pitch, rows;
    int nfbx, nfby;
    size_t field_bytes, off_phi[2], off_t[2], off_theta[2], off_flags, off_arrive, off_ticket, total;
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

Layout make_layout(int64_t nx, int64_t ny, int prec) {
    Layout L;
    L.elem = prec == KOB_F64 ? 8 : 4;
    L.pitch = (long long)align_up((size_t)(GX + nx + GXR), 32);
    L.rows = ny + 2 * GY;
    L.nfbx = (int)((L.pitch + FBX - 1) / FBX);
    L.nfby = (int)((L.rows + FBY - 1) / FBY);
    L.field_bytes = align_up((size_t)L.pitch * (size_t)L.rows * L.elem, 256);
    size_t o = 0;
    L.off_phi[0] = o; o += L.field_bytes;
    L.off_phi[1] = o; o += L.field_bytes;
    L.off_t[0] = o; o += L.field_bytes;
    L.off_t[1] = o; o += L.field_bytes;
    L.off_theta[0] = o; o += L.field_bytes;
    L.off_theta[1] = o; o += L.field_bytes;
    L.off_flags = o; o += align_up((size_t)L.nfbx * L.nfby * 4, 256);
    L.off_arrive = o; o += 256;
    L.off_ticket = o; o += 256;
    L.total = o;
    return L;
}

struct Neighbour {
    char* base = nullptr;   // device pointer to the neighbour's allocation (own base when unlinked)
    long long ny = 0;
    bool ipc = false;       // opened with cudaIpcOpenMemHandle -> must be closed
};

struct IpcBlob {            // payload of kob_ipc_handle
    uint32_t magic, prec;
    int64_t nx, ny, ny_global, y0;
    cudaIpcMemHandle_t mem;
};
static_assert(sizeof(IpcBlob) <= sizeof(kob_ipc_handle), "kob_ipc_handle too small");
constexpr uint32_t IPC_MAGIC = 0x4b4f4231u;  // "KOB1"

}  // namespace

// Shared-memory segment of one ring (kob_ring_join): per strip a round counter and the density probe of the last two rounds.
struct RingShm {
    uint32_t magic, world;
    struct Slot { volatile uint64_t round; volatile double frac[2]; char pad[40]; } slot[64];
};
constexpr uint32_t RING_MAGIC = 0x4b52494eu;   // "KRIN"
constexpr int RING_PERIOD = 64;                // sub-steps between agreements
constexpr double RING_TO_SINGLE = 0.04, RING_TO_PAIRS = 0.03;

struct kob_ctx {
    int64_t nx = 0, ny = 0, ny_global = 0, y0 = 0;
    int prec = KOB_F32, kernel = KOB_KERNEL_FAST, device = 0;
    uint64_t seed = 0, step = 0;
    uint32_t epoch = 0;
    int cur = 0;
    int tcur = 0;                 // which theta buffer is current (flipped by every two-step launch)
    kob_params params{};
    Layout L{};
    char* base = nullptr;
    float* noise_field = nullptr;
    uint8_t* rgba = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // asynchronous readback (kob_get_fields_async): device snapshot + second stream
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_snap = nullptr, ev_copied = nullptr;
    char* snap = nullptr;             // 3 x nx*ny elements, unpadded
    bool copy_pending = false;
    Neighbour lower, upper;
    bool linked = false;
    uint64_t launches = 0;
    // FAST kernel: TMA descriptors of the four field buffers, job decomposition, job-counter bookkeeping
    FastMapsIO maps_io{};         // single-step kernel: input boxes (phi, T 68 wide; theta 64 wide) and output boxes (60 wide)
    FastMaps maps2{}, maps_far{};   // two-step boxes (72 wide); far pass: phi 64 wide + T 72 wide
    int fast_yj = 64, fast_yj_b = 32;
    int fast_free = 1;            // a CTA that met data-dependent work switches to per-warp job claims (0: never; test knob)
    int nsm = 0;                  // SMs of the device
    std::unordered_map<const void*, int> occ;   // resident CTAs per SM of each kernel instantiation this context launched
    bool fast_yj_env = false, fast2_yj_env = false;   // job heights given explicitly: no automatic shortening
    double fast_frac_a = 0.9;
    int fast_cta_jobs = 2;        // CTA-wide jobs: 8 adjacent strips (1920 B contiguous per row); 2 = lock-step only on far-field jobs
    int fast_no_skip = 0;
    // Two sub-steps per launch pair (kob_fast2.cuh) where kob_step(n >= 2) allows it.  KOB_FAST2 = 0 never, 1 always,
    // 2 (default) adaptive: the fraction of jobs with data-dependent work is read back asynchronously — the far pass's
    // work-list length in pair mode, a live-job count of every 8th launch in single-step mode — and the single-step
    // kernel is used while that fraction is high (the general pass of the pair runs at 8 warps/SM), with hysteresis.
    // Results do not depend on the choice: both paths are bit-identical.
    int fast2 = 2;
    unsigned int* h_count = nullptr;   // pinned: work-list length of the last probed launch pair
    cudaEvent_t ev_count = nullptr;
    bool count_pending = false;
    uint64_t probe_launch = 0;         // launch count at which the pending probe was issued
    long long count_total = 1;
    double general_frac = 0.0;
    bool single_mode = false;          // adaptive policy: the field is dense, use the single-step kernel
    uint64_t n_single = 0, n_paired = 0;
    // ring-wide step-path agreement (kob_ring_join): shared segment, this strip's slot, sub-steps since the join
    struct RingShm* ring = nullptr;
    int ring_rank = 0, ring_world = 0, ring_mode = 1;    // mode: 1 = two-step pairs (sparse field), 0 = single-step kernel
    uint64_t ring_steps = 0, ring_round = 0;
    std::string ring_name;
    int fast2_yj = 96, fast2_yj_b = 32, fast2_lock = 0, fast2_far = 1, fast2_far_cta = 1;
    int* worklist = nullptr;      // far/general launch pair: 8 header words (FastArgs::list), then the row-range ids (-1 = empty slot)
    long long worklist_cap = 0;
    unsigned char* hot = nullptr; // [2][units]: claim units of the far pass that listed something in the previous / this pair
    long long hot_units = 0;
    int hot_par = 0;
    // The general pass beside the far pass (programmatic dependent launch) while the work list is short: KOB_FAST2_CONC = 0 never,
    // otherwise the largest list (row ranges) it is used for.
    int fast2_conc = 1 << 30;
    double fast2_ticket_us = 85.0; // what a listed row range costs a warp of the general pass
    int fast2_conc_sm = 1 << 20;  // the most SMs the general pass may take from the far pass (default: 40 % of the device)
    int fast2_conc_margin = 100;  // per cent of the probed list length the SMs are asked for
    int fast2_conc_serial = 0;    // KOB_FAST2_CONC_SERIAL=1 (tests): serialise the early general pass and its far pass like a profiler would
    long long list_est = -1;      // length of the last probed work list (-1: none yet)
    bool probe_is_list = false;
    uint64_t n_conc = 0;
    bool registered = false;      // counted in g_ctx_on_device
    unsigned long long job_expected = 0;   // value of the device job counter before the next launch
    int64_t frames = 0;
    double sim_ms = 0.0;
    std::string err;
    // KOB_TRACE=<file>: CUDA events around the step kernels' launches (first 4096), written as CSV at kob_destroy
    std::string trace_path;
    std::vector<cudaEvent_t> trace_ev;
    std::vector<const char*> trace_name;
    std::vector<long long> trace_info;
};

namespace {

// contexts of this process per device: the concurrent general pass needs SMs of its own, which only a context that has the
// device to itself can count on
std::atomic<int> g_ctx_on_device[64];

int fail(kob_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}

// Launch trace (diagnostics): event before / after a kernel launch on the context's stream.
inline void trace_mark(kob_ctx* c, const char* name, long long info = 0) {
    if (c->trace_path.empty() || c->trace_ev.size() >= 8192) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, c->stream);
    c->trace_ev.push_back(e);
    c->trace_name.push_back(name);
    c->trace_info.push_back(info);
}
void trace_dump(kob_ctx* c) {
    if (c->trace_path.empty() || c->trace_ev.size() < 2) return;
    cudaStreamSynchronize(c->stream);
    if (FILE* fh = std::fopen(c->trace_path.c_str(), "w")) {
        std::fprintf(fh, "kernel,start_us,duration_us,gap_before_us,info\n");
        float prev_end = 0.f;
        for (size_t i = 0; i + 1 < c->trace_ev.size(); i += 2) {
            float t0 = 0.f, dt = 0.f;
            cudaEventElapsedTime(&t0, c->trace_ev[0], c->trace_ev[i]);
            cudaEventElapsedTime(&dt, c->trace_ev[i], c->trace_ev[i + 1]);
            std::fprintf(fh, "%s,%.2f,%.2f,%.2f,%lld\n", c->trace_name[i], t0 * 1e3f, dt * 1e3f, (t0 - prev_end) * 1e3f, c->trace_info[i]);
            prev_end = t0 + dt;
        }
        std::fclose(fh);
    }
    for (cudaEvent_t e : c->trace_ev) cudaEventDestroy(e);
    c->trace_ev.clear(); c->trace_name.clear(); c->trace_info.clear();
}

#define KOB_CUDA(c, call)                                                                         \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail((c), e_ == cudaErrorMemoryAllocation ? KOB_ERR_OOM : KOB_ERR_CUDA,        \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                      \
    } while (0)

#define KOB_TRY(expr) do { int rc_ = (expr); if (rc_ != KOB_OK) return rc_; } while (0)
#define KOB_DISPATCH(c, fn, ...) ((c)->prec == KOB_F64 ? fn<double>(__VA_ARGS__) : fn<float>(__VA_ARGS__))

template <typename real>
StripView<real> view_of(char* base, const Layout& L, long long ny, int tcur) {
    StripView<real> v;
    v.phi[0] = reinterpret_cast<real*>(base + L.off_phi[0]);
    v.phi[1] = reinterpret_cast<real*>(base + L.off_phi[1]);
    v.t[0] = reinterpret_cast<real*>(base + L.off_t[0]);
    v.t[1] = reinterpret_cast<real*>(base + L.off_t[1]);
    v.theta = reinterpret_cast<real*>(base + L.off_theta[tcur]);
    v.theta_next = reinterpret_cast<real*>(base + L.off_theta[tcur ^ 1]);
    v.tflags = reinterpret_cast<uint32_t*>(base + L.off_flags);
    v.arrive = reinterpret_cast<uint32_t*>(base + L.off_arrive);
    v.ny = ny;
    return v;
}

template <typename real>
KParams<real> kparams_of(const kob_params& p) {
    KParams<real> k;
    k.dx = (real)p.dx; k.dy = (real)p.dy; k.dt = (real)p.dt; k.tau = (real)p.tau;
    k.epsbar = (real)p.epsilon_bar; k.K = (real)p.K; k.delta = (real)p.delta; k.aniso = (real)p.anisotropy;
    k.alpha = (real)p.alpha; k.gamma = (real)p.gamma; k.teq = (real)p.t_eq;
    k.theta0 = (real)p.theta0; k.noise_a = (real)p.noise_a;
    // loop invariants in the reference's own evaluation order (host arithmetic is IEEE, no contraction)
    volatile real three_dx = (real)3.0f * k.dx;
    k.lapden = three_dx * k.dx;
    volatile real nebj = (-k.epsbar) * k.aniso;
    k.neg_ebjd = nebj * k.delta;
    k.alpha_over_pi = k.alpha / (real)REF_PI_F;
    k.inv_dx = (real)1 / k.dx; k.inv_dy = (real)1 / k.dy; k.inv_lapden = (real)1 / k.lapden;
    k.dt_over_tau = k.dt / k.tau;
    const double j = p.anisotropy;
    k.jmode = (j >= 0.0 && j <= 16.0 && std::floor(j) == j) ? (int)j : -1;
    return k;
}

// Neighbour layouts share nx/prec with ours; only ny (hence rows) may differ, and the field offsets
// depend on it, so views of neighbours are built from THEIR layout.
template <typename real>
StepArgs<real> args_of(kob_ctx* c) {
    StepArgs<real> a;
    a.self = view_of<real>(c->base, c->L, c->ny, c->tcur);
    const Layout Ll = make_layout(c->nx, c->lower.ny, c->prec), Lu = make_layout(c->nx, c->upper.ny, c->prec);
    a.lower = view_of<real>(c->lower.base, Ll, c->lower.ny, c->tcur);   // all strips of a ring run the same launch sequence
    a.upper = view_of<real>(c->upper.base, Lu, c->upper.ny, c->tcur);
    a.prm = kparams_of<real>(c->params);
    a.noise_field = c->noise_field;
    a.seed = c->seed; a.step = c->step;
    a.ticket = reinterpret_cast<unsigned int*>(c->base + c->L.off_ticket);
    a.pitch = c->L.pitch; a.nx = (int)c->nx; a.ny = (int)c->ny; a.y0 = c->y0;
    a.nfbx = c->L.nfbx; a.nfby = c->L.nfby;
    a.cur = c->cur; a.tcur = c->tcur; a.linked = c->linked ? 1 : 0; a.epoch = c->epoch;
    return a;
}

// ---- FAST kernel host side ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int build_fast_maps(kob_ctx* c) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn ||
        q != cudaDriverEntryPointSuccess)
        return fail(c, KOB_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled not available from this driver");
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
    const cuuint64_t dims[2] = {(cuuint64_t)c->L.pitch, (cuuint64_t)c->L.rows};
    const cuuint64_t strides[1] = {(cuuint64_t)c->L.pitch * 4};
    const cuuint32_t estr[2] = {1, 1};
    auto encode = [&](CUtensorMap* m, size_t off, int box_w) -> int {
        const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)FAST_RB};
        const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, c->base + off, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(c, KOB_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
        return KOB_OK;
    };
    for (int i = 0; i < 2; ++i) {
        // single-step kernel: 68-wide phi / T / theta input boxes, 60-wide output boxes
        KOB_TRY(encode(&c->maps_io.phi_in[i], c->L.off_phi[i], FastGeom::BW));
        KOB_TRY(encode(&c->maps_io.t_in[i], c->L.off_t[i], FastGeom::BW));
        KOB_TRY(encode(&c->maps_io.th_in[i], c->L.off_theta[i], FastGeom::BW));
        KOB_TRY(encode(&c->maps_io.phi_out[i], c->L.off_phi[i], FastGeom::OUTC));
        KOB_TRY(encode(&c->maps_io.t_out[i], c->L.off_t[i], FastGeom::OUTC));
        KOB_TRY(encode(&c->maps_io.th_out[i], c->L.off_theta[i], FastGeom::OUTC));
        // two-step kernel: 72-wide boxes; its far pass: phi 64 wide (only looked at), T 72 wide
        KOB_TRY(encode(&c->maps2.phi[i], c->L.off_phi[i], F2_BW));
        KOB_TRY(encode(&c->maps2.t[i], c->L.off_t[i], F2_BW));
        KOB_TRY(encode(&c->maps_far.phi[i], c->L.off_phi[i], FAR2_PBW));
        c->maps_far.t[i] = c->maps2.t[i];
    }
    return KOB_OK;
}



// Resident CTAs per SM of a kernel (and the opt-in to its dynamic shared memory), cached per context: a context is driven by
// one host thread at a time, so there is no shared mutable state between contexts stepped from different threads.
int kernel_occupancy(kob_ctx* c, const void* kern, int threads, int smem, int* out) {
    auto it = c->occ.find(kern);
    if (it == c->occ.end()) {
        int cps = 0;
        KOB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        KOB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cps, kern, threads, smem));
        if (cps < 1) return fail(c, KOB_ERR_CUDA, "kernel does not fit on an SM");
        it = c->occ.emplace(kern, cps).first;
    }
    *out = it->second;
    return KOB_OK;
}

// Rows per job: the configured height or a shorter one (halvings down to 4) — whichever minimises the expected makespan of the
// dynamically scheduled queue, (total rows incl. warm-up rows and per-job overhead) / warps + one job (the last one to finish).  On big grids that is
// the configured height; on small ones the jobs shrink until the tail is short and every warp of the persistent grid has work.
// Not applied when the height was set explicitly (environment).  Bit-neutral, like every job knob.
int auto_job_rows(int yj, bool fixed, int nstrips, long long ny, long long warps, int warm) {
    if (fixed) return yj;
    int best = yj;
    double best_cost = -1.0;
    for (int h = yj; h >= 4; h /= 2) {
        const long long jobs = (long long)nstrips * ((ny + h - 1) / h);
        const double per_job = (double)(h + warm) + 8.0;          // + claim, theta-flag scan and TMA ring fill: about 8 rows' worth
        const double cost = (double)jobs * per_job / (double)warps + per_job;
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = h; }
    }
    return std::max(4, best / 4 * 4);
}

template <int JM, int NOISE, bool ROT>
int launch_fast_t(kob_ctx* c, const StepArgs<float>& a, FastArgs f) {
    auto kern = kob_step_fast<JM, NOISE, ROT>;
    const int smem = fast_smem_bytes(FAST_WARPS);
    int cps = 0;
    KOB_TRY(kernel_occupancy(c, reinterpret_cast<const void*>(kern), FAST_WARPS * 32, smem, &cps));
    const int nsm = c->nsm;
    f.nstrips = (int)((c->nx + FastGeom::OUTC - 1) / FastGeom::OUTC);
    // guided job heights: fast_yj rows for the first fast_frac_a of the strip, fast_yj_b rows for the rest; on small
    // grids the jobs are made shorter until every warp of the persistent grid has one (a warp marches its job serially:
    // 250^2 with 64-row jobs is 20 warps at work, 62 us per launch; with 4-row jobs 12 us).  Bit-neutral.
    f.yj = auto_job_rows(c->fast_yj, c->fast_yj_env, f.nstrips, c->ny, (long long)nsm * cps * FAST_WARPS, 4);
    f.yj_b = std::min(c->fast_yj_b, f.yj);
    f.nseg_a = (int)(((double)c->ny * c->fast_frac_a) / f.yj);
    if ((long long)f.nseg_a * f.yj >= c->ny || f.yj_b == f.yj) f.nseg_a = (int)((c->ny + f.yj - 1) / f.yj);
    const long long rest = std::max<long long>(0, c->ny - (long long)f.nseg_a * f.yj);
    f.nseg = f.nseg_a + (int)((rest + f.yj_b - 1) / f.yj_b);
    f.cta_jobs = f.nstrips < FAST_WARPS ? 0 : c->fast_cta_jobs;   // narrow grids: a CTA job would be mostly padding warps
    f.no_skip = c->fast_no_skip;
    f.free_mode = c->fast_free;
    f.nstrips_p = (f.nstrips + FAST_WARPS - 1) / FAST_WARPS * FAST_WARPS;
    // linked strips: how many jobs touch the low / high seam (the kernel publishes a side when its jobs are done)
    f.seam_jobs[0] = f.seam_jobs[1] = 0;
    for (int seg = 0; seg < f.nseg; ++seg) {
        const long long y0 = seg < f.nseg_a ? (long long)seg * f.yj : (long long)f.nseg_a * f.yj + (long long)(seg - f.nseg_a) * f.yj_b;
        const long long y1 = std::min<long long>(y0 + (seg < f.nseg_a ? f.yj : f.yj_b), c->ny);
        if (y0 < GY + 1) f.seam_jobs[0] += (unsigned)f.nstrips;
        if (y1 > c->ny - GY - 1) f.seam_jobs[1] += (unsigned)f.nstrips;
    }
    const long long njobs = (long long)(f.cta_jobs ? f.nstrips_p : f.nstrips) * f.nseg;
    const int grid = (int)std::min<long long>((long long)nsm * cps, (njobs + FAST_WARPS - 1) / FAST_WARPS);
    f.job_base = c->job_expected;
    // adaptive policy, single-step mode: every 8th launch counts its live jobs (read back asynchronously)
    const bool probe = c->single_mode && !c->count_pending && (c->launches & 7) == 0;
    unsigned int* live_ctr = reinterpret_cast<unsigned int*>(c->base + c->L.off_ticket + 32);
    if (probe) {
        KOB_CUDA(c, cudaMemsetAsync(live_ctr, 0, sizeof(unsigned int), c->stream));
        f.live_ctr = live_ctr;
    }
    trace_mark(c, "kob_step_fast");
    kern<<<grid, FAST_WARPS * 32, smem, c->stream>>>(c->maps_io, a, f);
    trace_mark(c, "");
    if (probe) {
        KOB_CUDA(c, cudaMemcpyAsync(c->h_count, live_ctr, sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
        KOB_CUDA(c, cudaEventRecord(c->ev_count, c->stream));
        c->count_pending = true;
        c->probe_is_list = false;
        c->probe_launch = c->launches;
        c->count_total = (long long)f.nstrips * f.nseg;
    }
    c->job_expected += (unsigned long long)njobs + (unsigned long long)grid * FAST_WARPS;   // every warp overshoots once
    return KOB_OK;
}

FastArgs fast_args_of(kob_ctx* c, const StepArgs<float>& a) {
    const KParams<float>& P = a.prm;
    FastArgs f{};
    f.job_ctr = reinterpret_cast<unsigned long long*>(c->base + c->L.off_ticket + 8);
    f.seam_ctr = reinterpret_cast<unsigned int*>(c->base + c->L.off_ticket + 40);
    ColdK& K = f.ck;
    K.e = REF_DEADBAND;
    // re-assigned angle = quadrant offset + atan(w), w in [-1, 1] (kob_row.cuh): the reference's branch offsets
    // 0 / PI_F / 2 PI_F (src/Kobayashi.cpp:160-167) with the +-pi/4 of the argument fold, as a bilinear form in the signs
    const double q4 = 0.78539816339744830962, pif = (double)REF_PI_F;
    K.off_c = REF_PI_F; K.off_y = -0.5f * REF_PI_F; K.off_s = (float)(q4 - 0.5 * pif);
    K.half_pi = 0.5f * REF_PI_F;
    {   // case A cells: cos / sin (j * angl) with angl = +-0.5f * PI_F, float arithmetic like the reference (:156-158, :170-171)
        volatile float ap = P.aniso * (0.5f * REF_PI_F - P.theta0), am = P.aniso * (-0.5f * REF_PI_F - P.theta0);
        K.cfl_p = std::cos((float)ap); K.sfl_p = std::sin((float)ap); K.cfl_m = std::cos((float)am); K.sfl_m = std::sin((float)am);
        K.j_rev = (float)((double)P.aniso / 6.283185307179586);
        K.jth0_rev = (float)(-(double)P.aniso * (double)P.theta0 / 6.283185307179586);
    }
    // eps, eps' of a cell that holds theta = 0 (far field), evaluated like src/Kobayashi.cpp:170-171
    const double arg0 = (double)P.aniso * (0.0 - (double)P.theta0);
    volatile float c0 = (float)std::cos(arg0), s0 = (float)std::sin(arg0);
    volatile float dc = P.delta * c0;
    volatile float one_dc = 1.0f + dc;
    K.eps0 = P.epsbar * one_dc;
    K.epsd0 = P.neg_ebjd * s0;
    K.cj0 = (float)std::cos((double)P.aniso * (double)P.theta0);
    K.sj0 = (float)std::sin((double)P.aniso * (double)P.theta0);
    K.ebd = P.epsbar * P.delta;
    K.epsbar = P.epsbar;
    K.neg_ebjd = P.neg_ebjd;
    K.neg_gamma = -P.gamma;
    K.gamma_teq = P.gamma * P.teq;
    K.aop = P.alpha_over_pi;
    K.m_q = (float)((double)P.alpha_over_pi * q4);
    K.noise_a = P.noise_a;
    K.aniso = P.aniso; K.theta0 = P.theta0; K.jmode = P.jmode;
    RowConst& R = f.rc;
    R.idx = P.inv_dx; R.idy = P.inv_dy; R.il = P.inv_lapden; R.ildt = P.inv_lapden * P.dt; R.dtt = P.dt_over_tau; R.K = P.K;
    volatile float a0 = K.eps0 * K.eps0, b0 = K.eps0 * K.epsd0;
    R.A0 = a0; R.B0 = b0;
    for (int r = 0; r < 10; ++r) {
        f.pk[2 * r] = (uint32_t)a.seed + (uint32_t)r * 0x9E3779B9u;
        f.pk[2 * r + 1] = (uint32_t)(a.seed >> 32) + (uint32_t)r * 0xBB67AE85u;
    }
    f.pc2 = (uint32_t)a.step;
    f.pc3 = (uint32_t)(a.step >> 32);
    f.ny_global = c->ny_global;
    return f;
}

// SMs for a general pass that runs beside its far pass (0: keep the plain far -> general order).  Pure host arithmetic
// (kob_policy_conc_sms exposes it to the CPU tests): a ticket (one listed row range) takes a warp ~ticket_us, so the list needs
// tickets x ticket_us / (8 warps x g); the far pass streams ~2.75 ps per cell while it keeps >= ~85 % of the SMs (HBM bound) and slows
// in proportion beyond that.  The smallest g whose general pass finishes with the far pass (+40 us of slack: the closing launch
// serves a remainder), searched up to 40 % of the device.
int conc_sm_budget(double cells, int nsm, long long tickets, double ticket_us, int margin_pct, int cap) {
    if (nsm < 16 || tickets < 0) return 0;
    const double far_us = 2.75e-6 * cells;
    const double work_us = (double)(tickets * margin_pct / 100 + 8) * ticket_us / 8.0;
    const int cap_sm = std::max(2, std::min(cap, nsm * 2 / 5));
    for (int g = 2; g <= cap_sm; ++g) {
        const double far_g = std::max(far_us, 0.85 * far_us * nsm / (double)(nsm - g));
        if (work_us / g <= 0.95 * far_g + 40.0) return g;
    }
    return 0;
}

// ---- two sub-steps per launch (kob_fast2.cuh) ----
template <int JM, bool NOISE, bool ROT>
int launch_fast2_t(kob_ctx* c, const StepArgs<float>& a, FastArgs f) {
    auto kern = kob_step_fast2<JM, NOISE, ROT>;
    const int smem = F2_WARPS * F2_WARP_BYTES + F2_WARPS * F2_NST * 8;
    int cps = 0;
    KOB_TRY(kernel_occupancy(c, reinterpret_cast<const void*>(kern), F2_WARPS * 32, smem, &cps));
    const int nsm = c->nsm;
    f.nstrips = (int)((c->nx + F2_OUTC - 1) / F2_OUTC);
    f.yj = auto_job_rows(c->fast2_yj, c->fast2_yj_env, f.nstrips, c->ny, (long long)nsm * 3 * FAR2_WARPS, 8);
    f.yj_b = std::min(c->fast2_yj_b, f.yj);
    f.nseg_a = (int)(((double)c->ny * c->fast_frac_a) / f.yj);
    if ((long long)f.nseg_a * f.yj >= c->ny || f.yj_b == f.yj) f.nseg_a = (int)((c->ny + f.yj - 1) / f.yj);
    const long long rest = std::max<long long>(0, c->ny - (long long)f.nseg_a * f.yj);
    f.nseg = f.nseg_a + (int)((rest + f.yj_b - 1) / f.yj_b);
    f.cta_jobs = c->fast2_lock;
    f.no_skip = c->fast_no_skip;
    f.nstrips_p = (f.nstrips + F2_WARPS - 1) / F2_WARPS * F2_WARPS;
    // linked strips: row ranges (F2_RANGES per job) of the jobs that touch the low / high seam — a side is published as soon
    // as they are done, by whichever pass of the launch pair completes the last one
    f.seam_jobs[0] = f.seam_jobs[1] = 0;
    for (int seg = 0; seg < f.nseg; ++seg) {
        const long long y0 = seg < f.nseg_a ? (long long)seg * f.yj : (long long)f.nseg_a * f.yj + (long long)(seg - f.nseg_a) * f.yj_b;
        const long long y1 = std::min<long long>(y0 + (seg < f.nseg_a ? f.yj : f.yj_b), c->ny);
        if (y0 < GY + 2) f.seam_jobs[0] += (unsigned)(F2_RANGES * f.nstrips);
        if (y1 > c->ny - GY - 1) f.seam_jobs[1] += (unsigned)(F2_RANGES * f.nstrips);
    }
    const long long njobs = (long long)(f.cta_jobs ? f.nstrips_p : f.nstrips) * f.nseg;
    if (c->fast2_far && !f.cta_jobs) {
        // far/general launch pair: the light far pass visits every job and leaves the rest on the work list
        constexpr int HDR = LH_WORDS;
        if (c->worklist_cap < njobs) {
            if (c->worklist) { KOB_CUDA(c, cudaStreamSynchronize(c->stream)); cudaFree(c->worklist); c->worklist = nullptr; }
            KOB_CUDA(c, cudaMalloc((void**)&c->worklist, (size_t)(F2_RANGES * njobs + HDR) * sizeof(int)));
            KOB_CUDA(c, cudaMemsetAsync(c->worklist, 0, HDR * sizeof(int), c->stream));   // header: armed once; each pair's last launch re-arms it
            KOB_CUDA(c, cudaMemsetAsync(c->worklist + LH_MIN, 0xff, sizeof(int), c->stream));
            KOB_CUDA(c, cudaMemsetAsync(c->worklist + HDR, 0xff, (size_t)F2_RANGES * njobs * sizeof(int), c->stream));   // all slots empty
            c->worklist_cap = njobs;
        }
        unsigned int* hdr = reinterpret_cast<unsigned int*>(c->worklist);
        const int far_smem = FAR2_WARPS * FAR2_WARP_BYTES + FAR2_WARPS * FAR2_NST * 8;
        int fcps = 0;
        KOB_TRY(kernel_occupancy(c, reinterpret_cast<const void*>(kob_far2), FAR2_WARPS * 32, far_smem, &fcps));
        FastArgs ff = f;
        ff.cta_jobs = f.nstrips < FAR2_WARPS ? 0 : c->fast2_far_cta;      // narrow grids: a CTA job would be mostly padding
        const long long fjobs = (long long)(ff.cta_jobs ? f.nstrips_p : f.nstrips) * f.nseg;
        const long long units = ff.cta_jobs ? fjobs / FAR2_WARPS : fjobs;
        if (c->hot_units != units) {
            if (c->hot) { KOB_CUDA(c, cudaStreamSynchronize(c->stream)); cudaFree(c->hot); c->hot = nullptr; }
            const size_t ub = ((size_t)units + 15) / 16 * 16;                              // bytes of one flag array, 16-byte aligned
            KOB_CUDA(c, cudaMalloc((void**)&c->hot, 2 * ub + 2 * (size_t)units * sizeof(unsigned int)));   // 2 flag arrays + 2 compact lists
            KOB_CUDA(c, cudaMemsetAsync(c->hot, 0, 2 * ub + 2 * (size_t)units * sizeof(unsigned int), c->stream));
            KOB_CUDA(c, cudaMemsetAsync(hdr + LH_HOTCNT, 0, 2 * sizeof(unsigned int), c->stream));
            c->hot_units = units;
        }
        const size_t ub = ((size_t)units + 15) / 16 * 16;
        unsigned int* hotlists = reinterpret_cast<unsigned int*>(c->hot + 2 * ub);
        Far2Args w{c->worklist + HDR, hdr, c->hot + (size_t)c->hot_par * ub, c->hot + (size_t)(c->hot_par ^ 1) * ub,
                   hotlists + (size_t)c->hot_par * units, hotlists + (size_t)(c->hot_par ^ 1) * units, c->hot_par, 1};
        f.list_hot_par = c->hot_par;
        c->hot_par ^= 1;
        ff.job_base = c->job_expected;
        f.list = c->worklist + HDR; f.list_count = hdr + LH_COUNT; f.list_claim = hdr + LH_CLAIM;
        f.list_cap = (unsigned int)(F2_RANGES * njobs);
        // Short list (the last probe says so): the general pass goes FIRST, on a few SMs of its own, and serves the list while the far
        // pass — a programmatic dependent launch, started as soon as the general pass is resident — streams the grid.  One more
        // launch of the general pass closes the pair (normally it finds nothing to do; it serves what a too small estimate left
        // over, or everything if the kernels were serialised) and re-arms the list header.
        // How many SMs g: a ticket (one listed row range) takes a warp ~85 us, so the list needs est x 85 us / (8 warps x g); the far
        // pass streams ~2.75 ps per cell while it keeps >= ~85 % of the SMs (HBM bound: 24 of 148 SMs less cost it 4 %) and slows in
        // proportion beyond that (37 less: 13 %; both measured at 16384^2).  Take the smallest g whose general pass finishes with
        // the far pass; none up to 40 % of the SMs (a listed fraction of ~3 %): keep the plain far -> general order.
        int gsm = 0;
        if (c->fast2_conc > 0 && c->list_est >= 0 && c->list_est <= c->fast2_conc && c->device < 64 &&
            g_ctx_on_device[c->device].load() == 1)
            gsm = conc_sm_budget((double)c->nx * (double)c->ny, nsm, c->list_est, c->fast2_ticket_us, c->fast2_conc_margin, c->fast2_conc_sm);
        const bool conc = gsm > 0;
        // The early general pass leaves when the far pass is done; what it has not served by then (a stale, too small estimate)
        // is left to the closing launch, which therefore is a full grid: never much worse than the plain order.
        const bool drain = false;
        const int fgrid = (int)std::min<long long>((long long)(nsm - gsm) * fcps, (fjobs + FAR2_WARPS - 1) / FAR2_WARPS);
        trace_mark(c, conc ? "pair(general || far2)" : "kob_far2", c->list_est * 1000 + gsm);   // info: probed list length, SMs
        if (conc) {
            f.list_conc = 1; f.list_rearm = 0; f.list_drain = drain ? 1 : 0;
            kern<<<gsm * cps, F2_WARPS * 32, smem, c->stream>>>(c->maps2, a, f);
            if (c->fast2_conc_serial) KOB_CUDA(c, cudaStreamSynchronize(c->stream));   // test knob: what a profiler does to the pair
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3((unsigned)fgrid); cfg.blockDim = dim3(FAR2_WARPS * 32); cfg.dynamicSmemBytes = (size_t)far_smem; cfg.stream = c->stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            KOB_CUDA(c, cudaLaunchKernelEx(&cfg, kob_far2, c->maps_far, a, ff, w));
            f.list_conc = 0; f.list_rearm = 1; f.list_drain = 1;
            kern<<<drain ? 1 : nsm * cps, F2_WARPS * 32, smem, c->stream>>>(c->maps2, a, f);
            trace_mark(c, "");
            c->n_conc += 1;
        } else {
            kob_far2<<<fgrid, FAR2_WARPS * 32, far_smem, c->stream>>>(c->maps_far, a, ff, w);
            trace_mark(c, "");
            f.list_conc = 0; f.list_rearm = 1; f.list_drain = 1;
            trace_mark(c, "kob_step_fast2(list)");
            kern<<<nsm * cps, F2_WARPS * 32, smem, c->stream>>>(c->maps2, a, f);
            trace_mark(c, "");
        }
        // every far-pass warp / CTA overshoots the job counter once
        c->job_expected += (unsigned long long)fjobs + (unsigned long long)fgrid * FAR2_WARPS;
        c->launches += conc ? 2 : 1;
        if (!c->count_pending && c->h_count) {                            // density probe (adaptive policy / kob_path_stats)
            KOB_CUDA(c, cudaMemcpyAsync(c->h_count, hdr + LH_LAST, sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
            KOB_CUDA(c, cudaEventRecord(c->ev_count, c->stream));
            c->count_pending = true;
            c->probe_is_list = true;
            c->probe_launch = c->launches;
            c->count_total = (long long)F2_RANGES * f.nstrips * f.nseg;
        }
        return KOB_OK;
    }
    const int grid = (int)std::min<long long>((long long)nsm * cps, (njobs + F2_WARPS - 1) / F2_WARPS);
    f.job_base = c->job_expected;
    kern<<<grid, F2_WARPS * 32, smem, c->stream>>>(c->maps2, a, f);
    c->job_expected += (unsigned long long)njobs + (unsigned long long)grid * F2_WARPS;   // every warp / CTA overshoots once
    return KOB_OK;
}

// Can the next two sub-steps go through the two-step kernel?
bool fast2_eligible(const kob_ctx* c) {
    return c->kernel == KOB_KERNEL_FAST && c->prec == KOB_F32 && !c->noise_field &&
           c->nx >= 8 && c->ny >= 8;
}

int launch_two_steps_fast(kob_ctx* c) {
    const StepArgs<float> a = args_of<float>(c);
    const KParams<float>& P = a.prm;
    const FastArgs f = fast_args_of(c, a);
    const bool noise = c->params.noise_a != 0.0;
    const bool rot = P.theta0 != 0.0f;
    const int jm = P.jmode < 0 ? -1 : ((P.jmode == 4 || P.jmode == 6) && !rot ? P.jmode : 0);
#define KOB_FAST2_CASE(JM_, ROT_) \
    do { KOB_TRY((noise ? launch_fast2_t<JM_, true, ROT_>(c, a, f) : launch_fast2_t<JM_, false, ROT_>(c, a, f))); } while (0)
    if (jm == 4) KOB_FAST2_CASE(4, false);
    else if (jm == 6) KOB_FAST2_CASE(6, false);
    else if (jm == 0 && !rot) KOB_FAST2_CASE(0, false);
    else if (jm == 0 && rot) KOB_FAST2_CASE(0, true);
    else KOB_FAST2_CASE(-1, false);
#undef KOB_FAST2_CASE
    KOB_CUDA(c, cudaGetLastError());
    c->launches += 1;
    c->cur ^= 1; c->tcur ^= 1; c->step += 2; c->epoch += 2;
    return KOB_OK;
}

int launch_step_fast(kob_ctx* c, const StepArgs<float>& a, bool noise) {
    const KParams<float>& P = a.prm;
    FastArgs f = fast_args_of(c, a);
    const bool rot = P.theta0 != 0.0f;
    const bool field = noise && c->noise_field != nullptr;     // host-injected noise: the run-time-j instantiations only
    const int jm = P.jmode < 0 ? -1 : ((P.jmode == 4 || P.jmode == 6) && !rot && !field ? P.jmode : 0);
#define KOB_FAST_CASE(JM_, ROT_) \
    return noise ? launch_fast_t<JM_, 1, ROT_>(c, a, f) : launch_fast_t<JM_, 0, ROT_>(c, a, f)
    if (field) {
        if (jm == 0 && !rot) return launch_fast_t<0, 2, false>(c, a, f);
        if (jm == 0 && rot) return launch_fast_t<0, 2, true>(c, a, f);
        return launch_fast_t<-1, 2, false>(c, a, f);
    }
    if (jm == 4) { KOB_FAST_CASE(4, false); }
    if (jm == 6) { KOB_FAST_CASE(6, false); }
    if (jm == 0 && !rot) { KOB_FAST_CASE(0, false); }
    if (jm == 0 && rot) { KOB_FAST_CASE(0, true); }
    KOB_FAST_CASE(-1, false);
#undef KOB_FAST_CASE
}

template <typename real>
int launch_fast_dispatch(kob_ctx* c, const StepArgs<real>&, bool) { return fail(c, KOB_ERR_UNSUPPORTED, "the FAST kernel is FP32 only"); }
template <>
int launch_fast_dispatch<float>(kob_ctx* c, const StepArgs<float>& a, bool noise) { return launch_step_fast(c, a, noise); }

template <typename real>
int launch_one_step(kob_ctx* c) {
    StepArgs<real> a = args_of<real>(c);
    const bool noise = c->params.noise_a != 0.0;
    if (c->kernel == KOB_KERNEL_STRICT) {
        constexpr int TX = 32, TY = 16;
        dim3 block(32, 8), grid((unsigned)((c->nx + TX - 1) / TX), (unsigned)((c->ny + TY - 1) / TY));
        if (noise) kob_step_strict<real, TX, TY, true><<<grid, block, 0, c->stream>>>(a);
        else kob_step_strict<real, TX, TY, false><<<grid, block, 0, c->stream>>>(a);
    } else {
        KOB_TRY(launch_fast_dispatch<real>(c, a, noise));
    }
    KOB_CUDA(c, cudaGetLastError());
    c->launches += 1;
    c->cur ^= 1; c->step += 1; c->epoch += 1;
    return KOB_OK;
}

template <typename real>
int refresh_impl(kob_ctx* c) {
    StepArgs<real> a = args_of<real>(c);
    const long long items = 2LL * GY * c->nx + 2LL * GXR * c->ny;
    const int threads = 256;
    const int blocks = (int)std::min<long long>((items + threads - 1) / threads, 4096);
    kob_refresh_aliases<real><<<blocks, threads, 0, c->stream>>>(a);
    kob_publish_epoch<real><<<1, 32, 0, c->stream>>>(a);
    KOB_CUDA(c, cudaGetLastError());
    c->launches += 2;
    return KOB_OK;
}

template <typename real>
int rebuild_flags_impl(kob_ctx* c) {
    StepArgs<real> a = args_of<real>(c);
    dim3 grid(c->L.nfbx, c->L.nfby);
    kob_rebuild_flags<real><<<grid, 256, 0, c->stream>>>(a.self.theta, a.self.tflags, c->L.pitch, c->L.rows, c->L.nfbx);
    KOB_CUDA(c, cudaGetLastError());
    c->list_est = -1;             // the host wrote fields: what the last probe said about the work list no longer holds
    c->launches += 1;
    return KOB_OK;
}

template <typename real>
int nucleus_impl(kob_ctx* c, int64_t x, int64_t y) {
    StepArgs<real> a = args_of<real>(c);
    kob_nucleus<real><<<1, 32, 0, c->stream>>>(a, x, y, c->ny_global);
    KOB_CUDA(c, cudaGetLastError());
    c->launches += 1;
    return KOB_OK;
}

int set_device(kob_ctx* c) {
    KOB_CUDA(c, cudaSetDevice(c->device));
    return KOB_OK;
}

int zero_fields(kob_ctx* c) {
    // phi[2], T[2], theta[2], theta flags: one contiguous range
    KOB_CUDA(c, cudaMemsetAsync(c->base, 0, c->L.off_arrive, c->stream));
    return KOB_OK;
}


bool params_ok(const kob_params* p) {
    return p->dx > 0 && p->dy > 0 && p->dt > 0 && p->tau > 0 && std::isfinite(p->dx) && std::isfinite(p->dt) &&
           std::isfinite(p->tau) && std::isfinite(p->anisotropy);
}

}  // namespace

extern "C" {

int kob_abi_version(void) { return KOB_ABI_VERSION; }

const char* kob_strerror(int s) {
    switch (s) {
        case KOB_OK: return "ok";
        case KOB_ERR_INVALID_ARG: return "invalid argument";
        case KOB_ERR_CUDA: return "CUDA error";
        case KOB_ERR_NO_DEVICE: return "no CUDA device";
        case KOB_ERR_OOM: return "out of device memory";
        case KOB_ERR_STATE: return "invalid state";
        case KOB_ERR_UNSUPPORTED: return "unsupported";
        default: return "unknown status";
    }
}

const char* kob_last_error(const kob_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int kob_default_params(kob_params* p, double dt) {
    if (!p) return KOB_ERR_INVALID_ARG;
    p->dx = 0.03; p->dy = 0.03; p->dt = dt;                           // src/Kobayashi.cpp:61-63
    p->tau = 0.0003; p->epsilon_bar = 0.010; p->mu = 1.0; p->K = 1.6; // :76-79
    p->delta = 0.05; p->anisotropy = 6.0; p->alpha = 0.9; p->gamma = 10.0; p->t_eq = 1.0;  // :80-84
    p->theta0 = 0.0; p->noise_a = 0.0;
    return KOB_OK;
}

int kob_default_config(kob_config* c) {
    if (!c) return KOB_ERR_INVALID_ARG;
    std::memset(c, 0, sizeof(*c));
    c->precision = KOB_F32; c->kernel = KOB_KERNEL_FAST; c->device = 0;
    return KOB_OK;
}

int kob_create(kob_ctx** out, int64_t nx, int64_t ny, const kob_params* params, const kob_config* cfg) {
    if (!out) return fail(nullptr, KOB_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    kob_params p;
    kob_default_params(&p, 1e-4);
    if (params) p = *params;
    kob_config cf;
    kob_default_config(&cf);
    if (cfg) cf = *cfg;
    if (nx < 2 || ny < 2 || nx > (1LL << 30) || ny > (1LL << 30))
        return fail(nullptr, KOB_ERR_INVALID_ARG, "grid must be at least 2x2 per strip");
    if (!params_ok(&p)) return fail(nullptr, KOB_ERR_INVALID_ARG, "dx, dy, dt, tau must be positive and finite");
    if (cf.precision != KOB_F32 && cf.precision != KOB_F64) return fail(nullptr, KOB_ERR_INVALID_ARG, "bad precision");
    if (cf.kernel != KOB_KERNEL_STRICT && cf.kernel != KOB_KERNEL_FAST) return fail(nullptr, KOB_ERR_INVALID_ARG, "bad kernel");
    if (cf.kernel == KOB_KERNEL_FAST && cf.precision != KOB_F32)
        return fail(nullptr, KOB_ERR_UNSUPPORTED, "the FAST kernel is FP32 only; use KOB_KERNEL_STRICT for FP64");
    const int64_t nyg = cf.ny_global ? cf.ny_global : ny;
    if (cf.y0 < 0 || cf.y0 + ny > nyg) return fail(nullptr, KOB_ERR_INVALID_ARG, "strip [y0, y0+ny) outside ny_global");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, KOB_ERR_NO_DEVICE, "no CUDA device visible (this library has no CPU fallback)");
    if (cf.device < 0 || cf.device >= ndev) return fail(nullptr, KOB_ERR_INVALID_ARG, "device ordinal out of range");

    kob_ctx* c = new (std::nothrow) kob_ctx();
    if (!c) return fail(nullptr, KOB_ERR_OOM, "host allocation failed");
    c->nx = nx; c->ny = ny; c->ny_global = nyg; c->y0 = cf.y0;
    c->prec = cf.precision; c->kernel = cf.kernel; c->device = cf.device; c->seed = cf.seed;
    c->params = p;
    c->L = make_layout(nx, ny, c->prec);
    auto bail = [&](int code, const std::string& m) { g_create_error = m; kob_destroy(c); return code; };
    if (c->device < 64) { g_ctx_on_device[c->device].fetch_add(1); c->registered = true; }
    cudaError_t e;
    if ((e = cudaSetDevice(c->device)) != cudaSuccess) return bail(KOB_ERR_CUDA, cudaGetErrorString(e));
    if ((e = cudaDeviceGetAttribute(&c->nsm, cudaDevAttrMultiProcessorCount, c->device)) != cudaSuccess) return bail(KOB_ERR_CUDA, cudaGetErrorString(e));
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(KOB_ERR_CUDA, cudaGetErrorString(e));
    if ((e = cudaEventCreate(&c->ev0)) != cudaSuccess) return bail(KOB_ERR_CUDA, cudaGetErrorString(e));
    if ((e = cudaEventCreate(&c->ev1)) != cudaSuccess) return bail(KOB_ERR_CUDA, cudaGetErrorString(e));
    if ((e = cudaEventCreateWithFlags(&c->ev_count, cudaEventDisableTiming)) != cudaSuccess) return bail(KOB_ERR_CUDA, cudaGetErrorString(e));
    if ((e = cudaHostAlloc((void**)&c->h_count, 64, cudaHostAllocDefault)) != cudaSuccess) return bail(KOB_ERR_CUDA, cudaGetErrorString(e));
    c->h_count[0] = 0u;
    if ((e = cudaMalloc((void**)&c->base, c->L.total)) != cudaSuccess)
        return bail(e == cudaErrorMemoryAllocation ? KOB_ERR_OOM : KOB_ERR_CUDA,
                    std::string("cudaMalloc(") + std::to_string(c->L.total) + "): " + cudaGetErrorString(e));
    if ((e = cudaMemsetAsync(c->base, 0, c->L.total, c->stream)) != cudaSuccess) return bail(KOB_ERR_CUDA, cudaGetErrorString(e));
    c->lower.base = c->base; c->lower.ny = ny;
    c->upper.base = c->base; c->upper.ny = ny;
    if (c->kernel == KOB_KERNEL_FAST) {
        // tuning knobs (defaults are the measured best): cells per lane = 2*NP, rows per job
        if (const char* e_ = std::getenv("KOB_FAST_FREE")) c->fast_free = std::atoi(e_) ? 1 : 0;
        if (const char* e_ = std::getenv("KOB_FAST_YJ")) { c->fast_yj = c->fast_yj_b = std::max(4, std::atoi(e_)); c->fast_yj_env = true; }
        if (const char* e_ = std::getenv("KOB_FAST_CTA")) c->fast_cta_jobs = std::min(2, std::max(0, std::atoi(e_)));
        if (const char* e_ = std::getenv("KOB_FAST_NOSKIP")) c->fast_no_skip = std::atoi(e_) ? 1 : 0;
        if (const char* e_ = std::getenv("KOB_FAST2")) { c->fast2 = std::min(2, std::max(0, std::atoi(e_))); c->single_mode = c->fast2 == 0; }
        if (const char* e_ = std::getenv("KOB_FAST2_YJ")) { c->fast2_yj = c->fast2_yj_b = std::max(4, std::atoi(e_)); c->fast2_yj_env = true; }
        if (const char* e_ = std::getenv("KOB_FAST2_YJB")) c->fast2_yj_b = std::max(4, std::atoi(e_));
        if (const char* e_ = std::getenv("KOB_TRACE")) {          // "%p" in the file name = process id (one file per rank)
            c->trace_path = e_;
            const size_t at = c->trace_path.find("%p");
            if (at != std::string::npos) c->trace_path.replace(at, 2, std::to_string((long long)getpid()));
        }
        // A CUDA tool injected into this process is likely to serialise kernels (ncu sets NV_COMPUTE_PROFILER_PERFWORKS_DIR and
        // NV_NSIGHT_INJECTION_PORT_BASE in its target; other tools come in through CUDA_INJECTION64_PATH): keep the plain
        // far -> general order then, instead of letting every early general pass wait 20 ms for a far pass that cannot start.
        // An explicit KOB_FAST2_CONC wins; if a tool is not recognised the 20 ms fallback still gives the right answer.
        if (std::getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") || std::getenv("NV_NSIGHT_INJECTION_PORT_BASE") ||
            std::getenv("CUDA_INJECTION64_PATH"))
            c->fast2_conc = 0;
        if (const char* e_ = std::getenv("KOB_FAST2_CONC")) c->fast2_conc = std::max(0, std::atoi(e_));
        if (const char* e_ = std::getenv("KOB_FAST2_TICKET_US")) c->fast2_ticket_us = std::max(1.0, std::atof(e_));
        if (const char* e_ = std::getenv("KOB_FAST2_CONC_SM")) c->fast2_conc_sm = std::max(1, std::atoi(e_));
        if (const char* e_ = std::getenv("KOB_FAST2_CONC_SERIAL")) c->fast2_conc_serial = std::atoi(e_);
        if (const char* e_ = std::getenv("KOB_FAST2_CONC_MARGIN")) c->fast2_conc_margin = std::max(50, std::atoi(e_));
        if (const char* e_ = std::getenv("KOB_FAST2_FAR")) c->fast2_far = std::atoi(e_) ? 1 : 0;
        if (const char* e_ = std::getenv("KOB_FAST2_FAR_CTA")) c->fast2_far_cta = std::atoi(e_) ? 1 : 0;
        if (const char* e_ = std::getenv("KOB_FAST2_LOCK")) c->fast2_lock = std::min(2, std::max(0, std::atoi(e_)));   // 0 per-warp jobs, 1 CTA jobs, 2 CTA jobs in lock-step
        if (const char* e_ = std::getenv("KOB_FAST_YJB")) c->fast_yj_b = std::max(4, std::atoi(e_));
        if (const char* e_ = std::getenv("KOB_FAST_FRAC")) c->fast_frac_a = std::min(1.0, std::max(0.0, std::atof(e_)));
        int rcm = build_fast_maps(c);
        if (rcm != KOB_OK) return bail(rcm, c->err);
    }
    int rc = kob_reset(c);
    if (rc != KOB_OK) return bail(rc, c->err);
    if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) return bail(KOB_ERR_CUDA, cudaGetErrorString(e));
    *out = c;
    return KOB_OK;
}

int kob_destroy(kob_ctx* c) {
    if (!c) return KOB_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->registered) g_ctx_on_device[c->device].fetch_sub(1);
    trace_dump(c);
    if (c->lower.ipc && c->lower.base) cudaIpcCloseMemHandle(c->lower.base);
    if (c->upper.ipc && c->upper.base && c->upper.base != c->lower.base) cudaIpcCloseMemHandle(c->upper.base);
    if (c->ring) {
        munmap(c->ring, sizeof(RingShm));
        if (c->ring_rank == 0) shm_unlink(c->ring_name.c_str());
    }
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    if (c->ev_snap) cudaEventDestroy(c->ev_snap);
    if (c->ev_copied) cudaEventDestroy(c->ev_copied);
    if (c->snap) cudaFree(c->snap);
    if (c->noise_field) cudaFree(c->noise_field);
    if (c->rgba) cudaFree(c->rgba);
    if (c->worklist) cudaFree(c->worklist);
    if (c->hot) cudaFree(c->hot);
    if (c->h_count) cudaFreeHost(c->h_count);
    if (c->ev_count) cudaEventDestroy(c->ev_count);
    if (c->base) cudaFree(c->base);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return KOB_OK;
}

int kob_clear(kob_ctx* c) {
    if (!c) return KOB_ERR_INVALID_ARG;
    KOB_TRY(set_device(c));
    KOB_TRY(zero_fields(c));
    c->step = 0;
    return KOB_OK;
}

int kob_reset(kob_ctx* c) {
    if (!c) return KOB_ERR_INVALID_ARG;
    KOB_TRY(kob_clear(c));
    c->frames = 0; c->sim_ms = 0.0;                                   // src/Kobayashi.cpp:247-248
    return kob_add_nucleus(c, c->nx / 2, c->ny_global / 2);           // src/Kobayashi.cpp:113
}

int kob_add_nucleus(kob_ctx* c, int64_t x, int64_t y) {
    if (!c) return KOB_ERR_INVALID_ARG;
    KOB_TRY(set_device(c));
    return KOB_DISPATCH(c, nucleus_impl, c, x, y);
}

int kob_set_params(kob_ctx* c, const kob_params* p) {
    if (!c || !p) return KOB_ERR_INVALID_ARG;
    if (!params_ok(p)) return fail(c, KOB_ERR_INVALID_ARG, "dx, dy, dt, tau must be positive and finite");
    c->params = *p;
    return KOB_OK;
}

int kob_get_params(const kob_ctx* c, kob_params* p) {
    if (!c || !p) return KOB_ERR_INVALID_ARG;
    *p = c->params;
    return KOB_OK;
}

// Fold a landed density probe into the policy state (non-blocking).
// Break-even of the two paths (B200, 16384^2): a pair costs ~0.73 ms + 20 ms x g (general pass at 8 warps/SM), two single
// steps ~2 x (0.69 + 1.65 g) ms, g = fraction of row ranges with crystal  ->  g ~ 0.04.  The single-step probe counts jobs
// under a set theta flag (128 x 32 blocks), which over-estimates g: that is the hysteresis.
void poll_density_probe(kob_ctx* c) {
    constexpr double FAST2_TO_SINGLE = 0.04, FAST2_TO_PAIR = 0.04;
    if (!c->count_pending || cudaEventQuery(c->ev_count) != cudaSuccess) return;
    c->count_pending = false;
    c->general_frac = (double)c->h_count[0] / (double)c->count_total;
    c->list_est = c->probe_is_list ? (long long)c->h_count[0] : -1;       // a short list lets the general pass run beside the far pass
    if (c->fast2 == 2 && !c->linked) {
        if (!c->single_mode && c->general_frac > FAST2_TO_SINGLE) c->single_mode = true;
        else if (c->single_mode && c->general_frac < FAST2_TO_PAIR) c->single_mode = false;
    }
}

// Ring-wide agreement on the step path (every RING_PERIOD sub-steps since kob_ring_join): publish this strip's density probe,
// wait for every strip's, take the maximum, apply the hysteresis.  Every strip computes the same decision from the same numbers.
static int ring_agree(kob_ctx* c) {
    KOB_CUDA(c, cudaStreamSynchronize(c->stream));          // the asynchronous density probe has landed
    poll_density_probe(c);
    RingShm* R = c->ring;
    const uint64_t k = ++c->ring_round;
    R->slot[c->ring_rank].frac[k & 1] = c->general_frac;
    __sync_synchronize();
    R->slot[c->ring_rank].round = k;
    const auto t0 = std::chrono::steady_clock::now();
    double m = 0.0;
    for (int r = 0; r < c->ring_world; ++r) {
        while (R->slot[r].round < k) {
            std::this_thread::yield();
            if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(60))
                return fail(c, KOB_ERR_STATE, "kob_ring_join: a strip of the ring did not reach the step-path agreement within 60 s "
                                              "(all strips must be stepped by the same amounts)");
        }
        __sync_synchronize();
        m = std::max(m, (double)R->slot[r].frac[k & 1]);
    }
    if (c->ring_mode == 1 && m > RING_TO_SINGLE) c->ring_mode = 0;
    else if (c->ring_mode == 0 && m < RING_TO_PAIRS) c->ring_mode = 1;
    c->single_mode = c->ring_mode == 0;                     // single-step launches probe the live-job fraction
    return KOB_OK;
}

int kob_step(kob_ctx* c, int64_t nsteps) {
    if (!c || nsteps < 0) return KOB_ERR_INVALID_ARG;
    KOB_TRY(set_device(c));
    const bool ring = c->ring != nullptr && c->linked && c->fast2 == 2 && c->kernel == KOB_KERNEL_FAST;
    for (int64_t s = 0; s < nsteps;) {
        // Launches are queued far ahead of the GPU, and the density probe that steers the adaptive path comes back
        // asynchronously: without a bound, one long kob_step call would run entirely on the path chosen before it.  So the host
        // never runs more than 16 launches ahead of an outstanding probe (the GPU still has those 16 queued: no bubble).  Linked
        // strips that joined a ring (one process each) too — the probe sizes the early general pass there: every strip waits only
        // for work at least 16 launches behind its own queue front, which its neighbours — stepped concurrently, at most one launch
        // pair apart on the device — have long queued.  (Strips of one process are stepped one after the other: never block there.)
        if (c->count_pending && c->fast2 == 2 && (!c->linked || c->ring != nullptr) && c->launches - c->probe_launch >= 16)
            KOB_CUDA(c, cudaEventSynchronize(c->ev_count));
        poll_density_probe(c);
        // KOB_FAST2 / kob_set_path_mode: 0 = single-step kernel, 1 = pairs, 2 = adaptive.  Linked strips must all run the same
        // launch sequence: "adaptive" there means pairs, unless the ring agrees on a mode — through kob_ring_join (below), or
        // by the caller setting the same mode on every strip at the same step.
        bool want_single = c->fast2 == 0 || (c->fast2 == 2 && !c->linked && c->single_mode);
        int64_t left = nsteps - s;
        if (ring) {
            if (c->ring_steps > 0 && c->ring_steps % RING_PERIOD == 0 && c->ring_round < c->ring_steps / RING_PERIOD) KOB_TRY(ring_agree(c));
            want_single = c->ring_mode == 0;
            left = std::min<int64_t>(left, RING_PERIOD - (int64_t)(c->ring_steps % RING_PERIOD));   // pairs never straddle an agreement
        }
        int done = 1;
        if (left >= 2 && fast2_eligible(c) && !want_single) {
            KOB_TRY(launch_two_steps_fast(c));                  // temporal blocking: two sub-steps per launch pair
            done = 2; c->n_paired += 2;
        } else {
            KOB_TRY(KOB_DISPATCH(c, launch_one_step, c));
            c->n_single += 1;
        }
        s += done;
        if (ring) c->ring_steps += (uint64_t)done;
    }
    return KOB_OK;
}

int kob_step_timed(kob_ctx* c, int64_t nsteps, float* ms) {
    if (!c || !ms) return KOB_ERR_INVALID_ARG;
    KOB_TRY(set_device(c));
    KOB_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    KOB_TRY(kob_step(c, nsteps));
    KOB_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    KOB_CUDA(c, cudaEventSynchronize(c->ev1));
    KOB_CUDA(c, cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return KOB_OK;
}

int kob_update(kob_ctx* c) {
    float ms = 0.f;
    KOB_TRY(kob_step_timed(c, 10, &ms));                              // src/Kobayashi.cpp:230-234
    c->sim_ms += ms; c->frames += 1;                                  // :237-238
    return KOB_OK;
}

// Linked strips: did an edge tile give up waiting for a neighbour, or see it run another launch sequence (kob_common.cuh, wait_flag)?
static int fault_status(kob_ctx* c, uint32_t fault) {
    if (fault == 0) return KOB_OK;
    if (fault == 3) return fail(c, KOB_ERR_STATE, "the general pass of a launch pair waited 20 s beside its far pass; fields are invalid");
    return fail(c, KOB_ERR_STATE, fault == 2 ? "a neighbour strip runs a different launch sequence (single steps vs two-step pairs): "
                                               "set the same path mode on every strip of the ring; fields are invalid"
                                             : "a neighbour strip did not reach the expected step within 20 s "
                                               "(strips must be stepped together); fields are invalid");
}
static int check_fault(kob_ctx* c) {
    if (!c->linked && c->n_conc == 0) return KOB_OK;
    uint32_t fault = 0;
    KOB_CUDA(c, cudaMemcpyAsync(&fault, c->base + c->L.off_arrive + 8, 4, cudaMemcpyDeviceToHost, c->stream));
    KOB_CUDA(c, cudaStreamSynchronize(c->stream));
    return fault_status(c, fault);
}
// After the copy stream has drained: look at the fault word without touching the compute stream's queue.
static int check_fault_nosync(kob_ctx* c) {
    if (!c->linked && c->n_conc == 0) return KOB_OK;
    uint32_t fault = 0;
    KOB_CUDA(c, cudaMemcpyAsync(&fault, c->base + c->L.off_arrive + 8, 4, cudaMemcpyDeviceToHost, c->copy_stream));
    KOB_CUDA(c, cudaStreamSynchronize(c->copy_stream));
    return fault_status(c, fault);
}

int kob_sync(kob_ctx* c) {
    if (!c) return KOB_ERR_INVALID_ARG;
    KOB_TRY(set_device(c));
    KOB_CUDA(c, cudaStreamSynchronize(c->stream));
    poll_density_probe(c);
    return check_fault(c);
}

int kob_get_fields(kob_ctx* c, void* phi, void* t, void* angl) {
    if (!c) return KOB_ERR_INVALID_ARG;
    KOB_TRY(set_device(c));
    const size_t e = c->L.elem, w = (size_t)c->nx * e, sp = (size_t)c->L.pitch * e;
    const size_t o = ((size_t)GY * c->L.pitch + GX) * e;
    if (phi) KOB_CUDA(c, cudaMemcpy2DAsync(phi, w, c->base + c->L.off_phi[c->cur] + o, sp, w, c->ny, cudaMemcpyDeviceToHost, c->stream));
    if (t) KOB_CUDA(c, cudaMemcpy2DAsync(t, w, c->base + c->L.off_t[c->cur] + o, sp, w, c->ny, cudaMemcpyDeviceToHost, c->stream));
    if (angl) KOB_CUDA(c, cudaMemcpy2DAsync(angl, w, c->base + c->L.off_theta[c->tcur] + o, sp, w, c->ny, cudaMemcpyDeviceToHost, c->stream));
    KOB_CUDA(c, cudaStreamSynchronize(c->stream));
    return check_fault(c);
}

int kob_set_fields(kob_ctx* c, const void* phi, const void* t, const void* angl) {
    if (!c) return KOB_ERR_INVALID_ARG;
    KOB_TRY(set_device(c));
    const size_t e = c->L.elem, w = (size_t)c->nx * e, sp = (size_t)c->L.pitch * e;
    const size_t o = ((size_t)GY * c->L.pitch + GX) * e;
    if (phi) KOB_CUDA(c, cudaMemcpy2DAsync(c->base + c->L.off_phi[c->cur] + o, sp, phi, w, w, c->ny, cudaMemcpyHostToDevice, c->stream));
    if (t) KOB_CUDA(c, cudaMemcpy2DAsync(c->base + c->L.off_t[c->cur] + o, sp, t, w, w, c->ny, cudaMemcpyHostToDevice, c->stream));
    if (angl) {
        KOB_CUDA(c, cudaMemcpy2DAsync(c->base + c->L.off_theta[c->tcur] + o, sp, angl, w, w, c->ny, cudaMemcpyHostToDevice, c->stream));
        // the two-step kernel only writes non-zero angles into the other buffer: it must not hold stale ones
        KOB_CUDA(c, cudaMemsetAsync(c->base + c->L.off_theta[c->tcur ^ 1], 0, c->L.field_bytes, c->stream));
    }
    KOB_TRY(KOB_DISPATCH(c, refresh_impl, c));
    if (angl) KOB_TRY(KOB_DISPATCH(c, rebuild_flags_impl, c));
    return KOB_OK;
}

int kob_get_fields_async(kob_ctx* c, void* phi, void* t, void* angl) {
    if (!c) return KOB_ERR_INVALID_ARG;
    KOB_TRY(set_device(c));
    const size_t e = c->L.elem, w = (size_t)c->nx * e, sp = (size_t)c->L.pitch * e, n = w * (size_t)c->ny;
    const size_t o = ((size_t)GY * c->L.pitch + GX) * e;
    if (!c->copy_stream) {
        KOB_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        KOB_CUDA(c, cudaEventCreateWithFlags(&c->ev_snap, cudaEventDisableTiming));
        KOB_CUDA(c, cudaEventCreateWithFlags(&c->ev_copied, cudaEventDisableTiming));
        KOB_CUDA(c, cudaMalloc((void**)&c->snap, 3 * n));
    }
    if (c->copy_pending) KOB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_copied, 0));     // the snapshot buffer is free again
    // snapshot (padded -> packed) on the compute stream: the following steps may overwrite the ping-pong buffers at once
    if (phi) KOB_CUDA(c, cudaMemcpy2DAsync(c->snap, w, c->base + c->L.off_phi[c->cur] + o, sp, w, c->ny, cudaMemcpyDeviceToDevice, c->stream));
    if (t) KOB_CUDA(c, cudaMemcpy2DAsync(c->snap + n, w, c->base + c->L.off_t[c->cur] + o, sp, w, c->ny, cudaMemcpyDeviceToDevice, c->stream));
    if (angl) KOB_CUDA(c, cudaMemcpy2DAsync(c->snap + 2 * n, w, c->base + c->L.off_theta[c->tcur] + o, sp, w, c->ny, cudaMemcpyDeviceToDevice, c->stream));
    KOB_CUDA(c, cudaEventRecord(c->ev_snap, c->stream));
    KOB_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_snap, 0));
    if (phi) KOB_CUDA(c, cudaMemcpyAsync(phi, c->snap, n, cudaMemcpyDeviceToHost, c->copy_stream));
    if (t) KOB_CUDA(c, cudaMemcpyAsync(t, c->snap + n, n, cudaMemcpyDeviceToHost, c->copy_stream));
    if (angl) KOB_CUDA(c, cudaMemcpyAsync(angl, c->snap + 2 * n, n, cudaMemcpyDeviceToHost, c->copy_stream));
    KOB_CUDA(c, cudaEventRecord(c->ev_copied, c->copy_stream));
    c->copy_pending = true;
    return KOB_OK;
}

int kob_wait_fields(kob_ctx* c) {
    if (!c) return KOB_ERR_INVALID_ARG;
    if (!c->copy_pending) return KOB_OK;
    KOB_TRY(set_device(c));
    KOB_CUDA(c, cudaEventSynchronize(c->ev_copied));
    c->copy_pending = false;
    return check_fault_nosync(c);
}

static bool window_ok(const kob_ctx* c, int64_t x0, int64_t y0, int64_t w, int64_t h) {
    return x0 >= 0 && y0 >= 0 && w > 0 && h > 0 && x0 + w <= c->nx && y0 + h <= c->ny;
}

int kob_get_window(kob_ctx* c, int64_t x0, int64_t y0, int64_t w, int64_t h, void* phi, void* t, void* angl) {
    if (!c) return KOB_ERR_INVALID_ARG;
    if (!window_ok(c, x0, y0, w, h)) return fail(c, KOB_ERR_INVALID_ARG, "window outside the strip");
    KOB_TRY(set_device(c));
    const size_t e = c->L.elem, wb = (size_t)w * e, sp = (size_t)c->L.pitch * e;
    const size_t o = ((size_t)(GY + y0) * c->L.pitch + GX + (size_t)x0) * e;
    if (phi) KOB_CUDA(c, cudaMemcpy2DAsync(phi, wb, c->base + c->L.off_phi[c->cur] + o, sp, wb, h, cudaMemcpyDeviceToHost, c->stream));
    if (t) KOB_CUDA(c, cudaMemcpy2DAsync(t, wb, c->base + c->L.off_t[c->cur] + o, sp, wb, h, cudaMemcpyDeviceToHost, c->stream));
    if (angl) KOB_CUDA(c, cudaMemcpy2DAsync(angl, wb, c->base + c->L.off_theta[c->tcur] + o, sp, wb, h, cudaMemcpyDeviceToHost, c->stream));
    KOB_CUDA(c, cudaStreamSynchronize(c->stream));
    return check_fault(c);
}

int kob_set_window(kob_ctx* c, int64_t x0, int64_t y0, int64_t w, int64_t h, const void* phi, const void* t, const void* angl) {
    if (!c) return KOB_ERR_INVALID_ARG;
    if (!window_ok(c, x0, y0, w, h)) return fail(c, KOB_ERR_INVALID_ARG, "window outside the strip");
    KOB_TRY(set_device(c));
    const size_t e = c->L.elem, wb = (size_t)w * e, sp = (size_t)c->L.pitch * e;
    const size_t o = ((size_t)(GY + y0) * c->L.pitch + GX + (size_t)x0) * e;
    if (phi) KOB_CUDA(c, cudaMemcpy2DAsync(c->base + c->L.off_phi[c->cur] + o, sp, phi, wb, wb, h, cudaMemcpyHostToDevice, c->stream));
    if (t) KOB_CUDA(c, cudaMemcpy2DAsync(c->base + c->L.off_t[c->cur] + o, sp, t, wb, wb, h, cudaMemcpyHostToDevice, c->stream));
    if (angl) {
        KOB_CUDA(c, cudaMemcpy2DAsync(c->base + c->L.off_theta[c->tcur] + o, sp, angl, wb, wb, h, cudaMemcpyHostToDevice, c->stream));
        // the two-step kernel only writes non-zero angles into the other buffer: the window must not hold stale ones there
        KOB_CUDA(c, cudaMemset2DAsync(c->base + c->L.off_theta[c->tcur ^ 1] + o, sp, 0, wb, h, c->stream));
    }
    KOB_TRY(KOB_DISPATCH(c, refresh_impl, c));
    if (angl) KOB_TRY(KOB_DISPATCH(c, rebuild_flags_impl, c));
    return KOB_OK;
}

int kob_set_noise_field(kob_ctx* c, const float* r) {
    if (!c) return KOB_ERR_INVALID_ARG;
    // a host-injected field rules the two-step path out on THIS strip only; inside a ring that would desynchronise the launch
    // sequences, so the ring has to be in single-step mode (kob_set_path_mode(KOB_PATH_SINGLE) on every strip) first
    if (r && c->linked && c->kernel == KOB_KERNEL_FAST && c->fast2 != 0)
        return fail(c, KOB_ERR_STATE, "linked FAST strips: set KOB_PATH_SINGLE on every strip of the ring before injecting a noise field");
    KOB_TRY(set_device(c));
    if (!r) {
        if (c->noise_field) { KOB_CUDA(c, cudaStreamSynchronize(c->stream)); cudaFree(c->noise_field); c->noise_field = nullptr; }
        return KOB_OK;
    }
    const size_t bytes = sizeof(float) * (size_t)c->nx * (size_t)c->ny;
    if (!c->noise_field) KOB_CUDA(c, cudaMalloc((void**)&c->noise_field, bytes));
    KOB_CUDA(c, cudaMemcpyAsync(c->noise_field, r, bytes, cudaMemcpyHostToDevice, c->stream));
    KOB_CUDA(c, cudaStreamSynchronize(c->stream));
    return KOB_OK;
}

int kob_set_step_counter(kob_ctx* c, uint64_t s) { if (!c) return KOB_ERR_INVALID_ARG; c->step = s; return KOB_OK; }
int kob_get_step_counter(const kob_ctx* c, uint64_t* s) { if (!c || !s) return KOB_ERR_INVALID_ARG; *s = c->step; return KOB_OK; }

int kob_render_rgba(kob_ctx* c, uint8_t* rgba) {
    if (!c || !rgba) return KOB_ERR_INVALID_ARG;
    KOB_TRY(set_device(c));
    const size_t bytes = 4 * (size_t)c->nx * (size_t)c->ny;
    if (!c->rgba) KOB_CUDA(c, cudaMalloc((void**)&c->rgba, bytes));
    dim3 block(32, 8), grid((unsigned)((c->nx + 31) / 32), (unsigned)((c->ny + 7) / 8));
    if (c->prec == KOB_F64)
        kob_render<double><<<grid, block, 0, c->stream>>>(reinterpret_cast<double*>(c->base + c->L.off_phi[c->cur]), c->rgba, c->L.pitch, (int)c->nx, (int)c->ny);
    else
        kob_render<float><<<grid, block, 0, c->stream>>>(reinterpret_cast<float*>(c->base + c->L.off_phi[c->cur]), c->rgba, c->L.pitch, (int)c->nx, (int)c->ny);
    KOB_CUDA(c, cudaGetLastError());
    c->launches += 1;
    KOB_CUDA(c, cudaMemcpyAsync(rgba, c->rgba, bytes, cudaMemcpyDeviceToHost, c->stream));
    KOB_CUDA(c, cudaStreamSynchronize(c->stream));
    return KOB_OK;
}

int kob_sim_frame(const kob_ctx* c, int64_t* f) { if (!c || !f) return KOB_ERR_INVALID_ARG; *f = c->frames; return KOB_OK; }
int kob_sim_time_ms(const kob_ctx* c, double* ms) { if (!c || !ms) return KOB_ERR_INVALID_ARG; *ms = c->sim_ms; return KOB_OK; }
int kob_set_sim_counters(kob_ctx* c, int64_t frames, double ms) {
    if (!c || frames < 0) return KOB_ERR_INVALID_ARG;
    c->frames = frames; c->sim_ms = ms;
    return KOB_OK;
}
int kob_set_path_mode(kob_ctx* c, int32_t mode) {
    if (!c || mode < 0 || mode > 2) return KOB_ERR_INVALID_ARG;
    c->fast2 = mode;
    c->single_mode = mode == 0;        // single-step launches probe the live-job fraction, pairs the work-list length
    return KOB_OK;
}
int kob_path_stats(const kob_ctx* c, uint64_t* single_steps, uint64_t* paired_steps, double* dense_fraction, int32_t* single_mode) {
    if (!c) return KOB_ERR_INVALID_ARG;
    if (single_steps) *single_steps = c->n_single;
    if (paired_steps) *paired_steps = c->n_paired;
    if (dense_fraction) *dense_fraction = c->general_frac;
    if (single_mode) *single_mode = c->single_mode ? 1 : 0;
    return KOB_OK;
}
int kob_wait_stats(kob_ctx* c, uint64_t* waits, double* wait_ms) {
    if (!c) return KOB_ERR_INVALID_ARG;
    KOB_TRY(set_device(c));
    uint32_t w[2] = {0, 0};
    KOB_CUDA(c, cudaMemcpyAsync(w, c->base + c->L.off_arrive + 16, 8, cudaMemcpyDeviceToHost, c->stream));
    KOB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (waits) *waits = w[1];
    if (wait_ms) *wait_ms = (double)w[0] * 1.024e-3;
    return KOB_OK;
}
int kob_policy_conc_sms(int64_t nx, int64_t ny, int32_t sms, int64_t listed_ranges) {
    return conc_sm_budget((double)nx * (double)ny, sms, listed_ranges, 85.0, 100, 1 << 20);
}
int kob_concurrent_pairs(const kob_ctx* c, uint64_t* n) { if (!c || !n) return KOB_ERR_INVALID_ARG; *n = c->n_conc; return KOB_OK; }
int kob_launch_count(const kob_ctx* c, uint64_t* n) { if (!c || !n) return KOB_ERR_INVALID_ARG; *n = c->launches; return KOB_OK; }
int kob_get_dims(const kob_ctx* c, int64_t* nx, int64_t* ny, int64_t* nyg, int64_t* y0) {
    if (!c) return KOB_ERR_INVALID_ARG;
    if (nx) *nx = c->nx; if (ny) *ny = c->ny; if (nyg) *nyg = c->ny_global; if (y0) *y0 = c->y0;
    return KOB_OK;
}

int kob_host_alloc(void** p, size_t bytes) {
    if (!p) return KOB_ERR_INVALID_ARG;
    cudaError_t e = cudaHostAlloc(p, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return KOB_ERR_OOM; }
    return KOB_OK;
}
int kob_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? KOB_OK : KOB_ERR_CUDA; }

// CPUs of the NUMA node the device hangs off (sysfs), or an empty set.
static bool cpus_near_device(int device, cpu_set_t* set) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) return false;
    for (char* p = bus; *p; ++p) *p = (char)std::tolower((unsigned char)*p);
    char path[128];
    std::snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE* fp = std::fopen(path, "r");
    int node = -1;
    if (fp) { if (std::fscanf(fp, "%d", &node) != 1) node = -1; std::fclose(fp); }
    if (node < 0) return false;
    std::snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    fp = std::fopen(path, "r");
    if (!fp) return false;
    CPU_ZERO(set);
    int a = 0, b = 0, n = 0;
    char sep = 0;
    while (std::fscanf(fp, "%d", &a) == 1) {           // "0-31,64-95"
        b = a;
        int ch = std::fgetc(fp);
        if (ch == '-') { if (std::fscanf(fp, "%d", &b) != 1) b = a; ch = std::fgetc(fp); }
        for (int k = a; k <= b && k < CPU_SETSIZE; ++k) { CPU_SET(k, set); ++n; }
        sep = (char)ch;
        if (sep != ',') break;
    }
    std::fclose(fp);
    return n > 0;
}

int kob_host_alloc_near(kob_ctx* c, void** p, size_t bytes) {
    if (!c || !p) return KOB_ERR_INVALID_ARG;
    cpu_set_t near_set, old_set;
    const bool have = cpus_near_device(c->device, &near_set) && sched_getaffinity(0, sizeof old_set, &old_set) == 0;
    bool moved = false;
    if (have) {
        cpu_set_t both;                                 // stay inside the CPUs this process is allowed to use
        CPU_AND(&both, &near_set, &old_set);
        if (CPU_COUNT(&both) > 0) moved = sched_setaffinity(0, sizeof both, &both) == 0;
    }
    const int rc = kob_host_alloc(p, bytes);
    if (rc == KOB_OK && moved) std::memset(*p, 0, bytes);   // first touch from a near CPU
    if (moved) sched_setaffinity(0, sizeof old_set, &old_set);
    return rc;
}

// ---- strips ------------------------------------------------------------------------------------------

int kob_ipc_export(kob_ctx* c, kob_ipc_handle* out) {
    if (!c || !out) return KOB_ERR_INVALID_ARG;
    KOB_TRY(set_device(c));
    IpcBlob b;
    std::memset(&b, 0, sizeof(b));
    b.magic = IPC_MAGIC; b.prec = (uint32_t)c->prec; b.nx = c->nx; b.ny = c->ny; b.ny_global = c->ny_global; b.y0 = c->y0;
    KOB_CUDA(c, cudaIpcGetMemHandle(&b.mem, c->base));
    std::memset(out, 0, sizeof(*out));
    std::memcpy(out->bytes, &b, sizeof(b));
    return KOB_OK;
}

static int check_neighbour(kob_ctx* c, int64_t nx, int64_t ny, int64_t nyg, int64_t y0, int prec, bool is_lower) {
    if (nx != c->nx || prec != c->prec || nyg != c->ny_global) return fail(c, KOB_ERR_INVALID_ARG, "neighbour strip has different nx / precision / ny_global");
    const int64_t expect = is_lower ? ((c->y0 - ny) % nyg + nyg) % nyg : (c->y0 + c->ny) % nyg;
    if (y0 != expect) return fail(c, KOB_ERR_INVALID_ARG, is_lower ? "lower neighbour is not adjacent (y0 mismatch)" : "upper neighbour is not adjacent (y0 mismatch)");
    if (ny < 2) return fail(c, KOB_ERR_INVALID_ARG, "neighbour strip too thin");
    // every strip of a ring must run the same launch sequence; the FAST two-step path needs 8 rows, so a ring of FAST strips
    // with a thinner member would have that member fall back to single steps alone
    if (c->kernel == KOB_KERNEL_FAST && (ny < 8 || c->ny < 8))
        return fail(c, KOB_ERR_INVALID_ARG, "linked FAST strips need at least 8 rows each (use the STRICT kernel for thinner strips)");
    return KOB_OK;
}

int kob_ipc_link(kob_ctx* c, const kob_ipc_handle* lower, const kob_ipc_handle* upper) {
    if (!c || !lower || !upper) return KOB_ERR_INVALID_ARG;
    if (c->linked) return fail(c, KOB_ERR_STATE, "strip already linked");
    KOB_TRY(set_device(c));
    IpcBlob bl, bu;
    std::memcpy(&bl, lower->bytes, sizeof(bl));
    std::memcpy(&bu, upper->bytes, sizeof(bu));
    if (bl.magic != IPC_MAGIC || bu.magic != IPC_MAGIC) return fail(c, KOB_ERR_INVALID_ARG, "not a kob_ipc_handle");
    KOB_TRY(check_neighbour(c, bl.nx, bl.ny, bl.ny_global, bl.y0, (int)bl.prec, true));
    KOB_TRY(check_neighbour(c, bu.nx, bu.ny, bu.ny_global, bu.y0, (int)bu.prec, false));
    void* pl = nullptr; void* pu = nullptr;
    KOB_CUDA(c, cudaIpcOpenMemHandle(&pl, bl.mem, cudaIpcMemLazyEnablePeerAccess));
    const bool same = std::memcmp(&bl.mem, &bu.mem, sizeof(bl.mem)) == 0;   // P == 2: one neighbour on both sides
    if (same) pu = pl; else KOB_CUDA(c, cudaIpcOpenMemHandle(&pu, bu.mem, cudaIpcMemLazyEnablePeerAccess));
    c->lower.base = (char*)pl; c->lower.ny = bl.ny; c->lower.ipc = true;
    c->upper.base = (char*)pu; c->upper.ny = bu.ny; c->upper.ipc = !same;
    c->linked = true;
    return KOB_OK;
}

int kob_link_local(kob_ctx* c, kob_ctx* lower, kob_ctx* upper) {
    if (!c || !lower || !upper) return KOB_ERR_INVALID_ARG;
    if (c->linked) return fail(c, KOB_ERR_STATE, "strip already linked");
    KOB_TRY(check_neighbour(c, lower->nx, lower->ny, lower->ny_global, lower->y0, lower->prec, true));
    KOB_TRY(check_neighbour(c, upper->nx, upper->ny, upper->ny_global, upper->y0, upper->prec, false));
    KOB_TRY(set_device(c));
    for (kob_ctx* n : {lower, upper}) {
        if (n->device != c->device) {
            int can = 0;
            KOB_CUDA(c, cudaDeviceCanAccessPeer(&can, c->device, n->device));
            if (!can) return fail(c, KOB_ERR_UNSUPPORTED, "no peer access between the strips' devices");
            cudaError_t e = cudaDeviceEnablePeerAccess(n->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(c, KOB_ERR_CUDA, cudaGetErrorString(e));
            cudaGetLastError();
        }
    }
    c->lower.base = lower->base; c->lower.ny = lower->ny; c->lower.ipc = false;
    c->upper.base = upper->base; c->upper.ny = upper->ny; c->upper.ipc = false;
    c->linked = true;
    return KOB_OK;
}

int kob_ring_join(kob_ctx* c, const char* name, int32_t rank, int32_t world) {
    if (!c || !name || world < 1 || world > 64 || rank < 0 || rank >= world || std::strlen(name) > 47) return KOB_ERR_INVALID_ARG;
    if (c->ring) return fail(c, KOB_ERR_STATE, "strip already joined a ring");
    c->ring_name = std::string("/kobring_") + name;
    const int fd = shm_open(c->ring_name.c_str(), O_CREAT | O_RDWR, 0600);
    if (fd < 0) return fail(c, KOB_ERR_STATE, "shm_open(" + c->ring_name + ") failed");
    if (ftruncate(fd, sizeof(RingShm)) != 0) { close(fd); return fail(c, KOB_ERR_STATE, "ftruncate on the ring segment failed"); }
    void* p = mmap(nullptr, sizeof(RingShm), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return fail(c, KOB_ERR_STATE, "mmap of the ring segment failed");
    c->ring = static_cast<RingShm*>(p);                      // a fresh segment is zero-filled: every slot starts at round 0
    c->ring->magic = RING_MAGIC; c->ring->world = (uint32_t)world;
    c->ring_rank = rank; c->ring_world = world; c->ring_mode = 1; c->ring_steps = 0; c->ring_round = 0;
    c->ring->slot[rank].round = 0;
    return KOB_OK;
}

int kob_halo_refresh(kob_ctx* c) {
    if (!c) return KOB_ERR_INVALID_ARG;
    KOB_TRY(set_device(c));
    KOB_TRY(KOB_DISPATCH(c, refresh_impl, c));
    return KOB_DISPATCH(c, rebuild_flags_impl, c);
}

}  // extern "C"
