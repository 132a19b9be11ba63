/*
 * kobayashi_c.h — C ABI of libkobayashi_cuda.so
 *
 * Drop-in boundary for the ONE hot path of jklae/CrystalGrowth: the explicit-Euler
 * time step of Kobayashi's anisotropic phase-field model
 * (reference: src/Kobayashi.cpp:98-239, class interface src/Kobayashi.h:32-122).
 *
 * Every entry point is plain C: opaque context pointer, POD structs, raw host
 * pointers and sizes.  No CUDA, torch or C++ types cross this boundary.  All calls
 * return 0 on success and a negative kob_status on failure (the reference's methods
 * are all `void` and unchecked; out-of-range nuclei are UB there, src/Kobayashi.cpp:116-123).
 *
 * Host field layout is the reference's: element (i, j) of an nx-wide grid lives at
 * index i + nx*j (src/Kobayashi.h:91), x fastest, no padding.  Element type is
 * float (KOB_F32, what the reference stores, src/Kobayashi.h:107-115) or double (KOB_F64).
 *
 * Threading: one context is driven by one host thread at a time (the reference runs
 * everything on the Win32 message-loop thread).  kob_step is asynchronous on the
 * context's stream; kob_get_fields / kob_sync synchronise.
 */
#ifndef KOBAYASHI_C_H
#define KOBAYASHI_C_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KOB_ABI_VERSION 1

typedef struct kob_ctx kob_ctx;

typedef enum kob_status {
    KOB_OK = 0,
    KOB_ERR_INVALID_ARG = -1,
    KOB_ERR_CUDA = -2,
    KOB_ERR_NO_DEVICE = -3,
    KOB_ERR_OOM = -4,
    KOB_ERR_STATE = -5,
    KOB_ERR_UNSUPPORTED = -6
} kob_status;

/* Model parameters.  Replaces the float members _dx.._tEq (src/Kobayashi.h:94-105),
 * their defaults (src/Kobayashi.cpp:61-63, :76-84) and the slider writes
 * (src/Kobayashi.cpp:589-611).  Stored as double; a KOB_F32 context rounds each to
 * float exactly once, which reproduces the reference's float members when the caller
 * passes the reference's decimal literals (0.0003 -> 0.0003f).
 * theta0 and noise_a are extensions named by the north star; 0 makes them bit-neutral. */
typedef struct kob_params {
    double dx;          /* 0.03   src/Kobayashi.cpp:61 */
    double dy;          /* 0.03   src/Kobayashi.cpp:62 */
    double dt;          /* ctor arg, 1e-4 in src/main.cpp:16 */
    double tau;         /* 0.0003 src/Kobayashi.cpp:76 */
    double epsilon_bar; /* 0.010  :77 */
    double mu;          /* 1.0    :78 — carried but never used by the arithmetic (as in the reference) */
    double K;           /* 1.6    :79  latent heat */
    double delta;       /* 0.05   :80  anisotropy strength */
    double anisotropy;  /* 6.0    :81  mode number j */
    double alpha;       /* 0.9    :82 */
    double gamma;       /* 10.0   :83 */
    double t_eq;        /* 1.0    :84 */
    double theta0;      /* extension: preferred-direction offset, eps(theta - theta0); 0 = reference */
    double noise_a;     /* extension: amplitude a of a*phi*(1-phi)*(r-1/2); 0 = reference (no noise) */
} kob_params;

enum { KOB_F32 = 0, KOB_F64 = 1 };

/* Kernel variants.  Both implement the reference semantics exactly (dead-band angle
 * state machine with carried theta, PI_F, 9-point Laplacians, Jacobi update).
 *  STRICT: reference operation order, IEEE division, no FMA contraction, portable
 *          atan/sin/cos shared with the CPU oracle -> bit-identical to the oracle.
 *  FAST:   same model, rounding-level differences only (reciprocal multiplies, FMA,
 *          trig-free eps(theta), packed f32x2 math) -> within 1e-4 on the parity windows. */
enum { KOB_KERNEL_STRICT = 0, KOB_KERNEL_FAST = 1 };

typedef struct kob_config {
    int32_t precision;   /* KOB_F32 | KOB_F64 */
    int32_t kernel;      /* KOB_KERNEL_STRICT | KOB_KERNEL_FAST (FAST is F32 only) */
    int32_t device;      /* CUDA device ordinal */
    int32_t flags;       /* reserved, 0 */
    uint64_t seed;       /* Philox key for the noise extension */
    /* Row-strip decomposition: this context owns global rows [y0, y0+ny) of an
     * nx x ny_global torus.  ny_global == 0 means a single strip (ny_global = ny, y0 = 0). */
    int64_t ny_global;
    int64_t y0;
} kob_config;

/* Opaque blob another process needs to map this strip's ghost rows (CUDA IPC). */
typedef struct kob_ipc_handle {
    unsigned char bytes[128];
} kob_ipc_handle;

/* ---- lifecycle ------------------------------------------------------------------ */

/* Reference defaults: _parameterInit (src/Kobayashi.cpp:73-85) + dx,dy,dt (:61-63). */
int kob_default_params(kob_params* p, double dt);
int kob_default_config(kob_config* c);

/* Kobayashi::Kobayashi(int x, int y, float timeStep) (src/Kobayashi.cpp:7-67): allocates
 * device state and performs the reference's _vectorInit (fields zero, one nucleus at
 * (nx/2, ny_global/2)).  params == NULL -> defaults with dt = 1e-4; cfg == NULL -> F32, FAST, device 0. */
int kob_create(kob_ctx** out, int64_t nx, int64_t ny, const kob_params* params, const kob_config* cfg);
int kob_destroy(kob_ctx* ctx);

/* _vectorInit (src/Kobayashi.cpp:98-114): zero phi, T, theta; nucleus at the centre; step counter = 0.
 * Parameters are kept (iResetSimulationState, src/Kobayashi.cpp:241-249). */
int kob_reset(kob_ctx* ctx);
/* Zero all fields without seeding a nucleus. */
int kob_clear(kob_ctx* ctx);
/* _createNucleus(x, y) (src/Kobayashi.cpp:116-123): phi = 1 on the 5-cell plus centred at global
 * cell (x, y).  Unlike the reference, coordinates wrap periodically instead of indexing out of range. */
int kob_add_nucleus(kob_ctx* ctx, int64_t x, int64_t y);

/* Slider writes (src/Kobayashi.cpp:589-611).  As in the reference the caller is expected to
 * kob_reset afterwards (src/Kobayashi.cpp:616); the library does not force it. */
int kob_set_params(kob_ctx* ctx, const kob_params* p);
int kob_get_params(const kob_ctx* ctx, kob_params* p);

/* ---- the hot path --------------------------------------------------------------- */

/* nsteps x { _computeGradientLaplacian(); _evolution(); } (src/Kobayashi.cpp:230-234), asynchronous.
 * One fused kernel launch per sub-step; with the FAST kernel, pairs of sub-steps may instead run as one
 * two-step launch pair (phi and T cross HBM once per two sub-steps).  The choice (environment KOB_FAST2 =
 * 0 never | 1 always | 2 adaptive, the default) never shows in the results: both paths are bit-identical. */
int kob_step(kob_ctx* ctx, int64_t nsteps);
/* Kobayashi::iUpdate (src/Kobayashi.cpp:227-239): 10 sub-steps, then _simFrame++ and
 * _simTime += elapsed ms. */
int kob_update(kob_ctx* ctx);
/* kob_step bracketed by CUDA events on the context's stream; *ms = device time of the nsteps launches. */
int kob_step_timed(kob_ctx* ctx, int64_t nsteps, float* ms);
int kob_sync(kob_ctx* ctx);

/* ---- field access (src/Kobayashi.cpp:315 reads _phi; set/get of all three arrays = exact checkpoint) -- */

/* Any pointer may be NULL.  Buffers hold nx*ny elements of the context's precision, reference layout.
 * kob_get_fields returns after the copies have landed.  kob_set_fields is ASYNCHRONOUS on the context's stream: with pinned
 * buffers (kob_host_alloc) the caller must kob_sync before overwriting or freeing them (pageable buffers are staged before
 * the call returns). */
int kob_get_fields(kob_ctx* ctx, void* phi, void* t, void* angl);
int kob_set_fields(kob_ctx* ctx, const void* phi, const void* t, const void* angl);
/* Asynchronous readback: the current arrays are snapshotted on the device (ordered with the steps queued so far) and copied to
 * the host buffers on a second stream while kob_step / kob_update continue; kob_wait_fields returns when the last
 * kob_get_fields_async has landed.  One readback in flight per context (a second call waits for the first).  Use pinned
 * buffers (kob_host_alloc / kob_host_alloc_near).  This is how a viewer overlaps frame n's picture with frame n+1's steps. */
int kob_get_fields_async(kob_ctx* ctx, void* phi, void* t, void* angl);
int kob_wait_fields(kob_ctx* ctx);
/* The same for the w x h window whose lower-left cell is (x0, y0) (local rows; no wrap): buffers hold w*h elements, row
 * stride w.  What a viewer of a 65536^2 torus reads, and how state beyond 2^31 cells is inspected — the reference's `int`
 * index (src/Kobayashi.h:91) cannot address such grids at all.  kob_set_window is asynchronous like kob_set_fields and
 * refreshes the periodic aliases of the cells it touches. */
int kob_get_window(kob_ctx* ctx, int64_t x0, int64_t y0, int64_t w, int64_t h, void* phi, void* t, void* angl);
int kob_set_window(kob_ctx* ctx, int64_t x0, int64_t y0, int64_t w, int64_t h, const void* phi, const void* t, const void* angl);
/* Host-injected noise field r in [0,1) (nx*ny floats, reference layout) used instead of the Philox
 * stream on every following step; NULL returns to Philox.  Parity aid named by the north star. */
int kob_set_noise_field(kob_ctx* ctx, const float* r);
int kob_set_step_counter(kob_ctx* ctx, uint64_t step);
int kob_get_step_counter(const kob_ctx* ctx, uint64_t* step);

/* RGBA8 image of phi through the viewer's 4-colour ramp (iUpdateConstantBuffer,
 * src/Kobayashi.cpp:309-345), computed on the device; pixel (i, j) at 4*(i + nx*j). */
int kob_render_rgba(kob_ctx* ctx, uint8_t* rgba);

/* ---- bookkeeping ----------------------------------------------------------------- */

int kob_sim_frame(const kob_ctx* ctx, int64_t* frames);      /* _simFrame, src/Kobayashi.cpp:238 */
int kob_sim_time_ms(const kob_ctx* ctx, double* ms);         /* _simTime,  src/Kobayashi.cpp:237 */
/* Write _simFrame / _simTime: iResetSimulationState zeroes them AFTER the viewer's refresh (src/Kobayashi.cpp:247-248);
 * a checkpoint resume restores them. */
int kob_set_sim_counters(kob_ctx* ctx, int64_t frames, double ms);
int kob_launch_count(const kob_ctx* ctx, uint64_t* launches);/* kernels launched by this context so far */
/* Step-path bookkeeping (no reference counterpart): sub-steps done by the single-step kernel and by two-step launch
 * pairs so far, the last density probe (fraction of jobs with data-dependent work) and the adaptive policy's mode. */
/* Select the step path at run time: 0 single-step kernel, 1 two-step launch pairs, 2 adaptive (the default; KOB_FAST2
 * sets the initial value).  Strips of one ring must all be given the same mode at the same step (StripRing does that from
 * an all-reduced density probe); results never depend on it. */
enum { KOB_PATH_SINGLE = 0, KOB_PATH_PAIRS = 1, KOB_PATH_ADAPTIVE = 2 };
int kob_set_path_mode(kob_ctx* ctx, int32_t mode);
int kob_path_stats(const kob_ctx* ctx, uint64_t* single_steps, uint64_t* paired_steps, double* dense_fraction, int32_t* single_mode);
/* Launch pairs whose general pass ran BESIDE its far pass (programmatic dependent launch; used while the work list is short and the
 * context has its device to itself; KOB_FAST2_CONC=0 turns it off, e.g. under a profiler that serialises kernels — the library
 * then still gives the same results, only later).  Diagnostics; results never depend on it. */
int kob_concurrent_pairs(const kob_ctx* ctx, uint64_t* n);
/* The launch-pair policy as a pure function (no device needed): SMs the general pass gets beside the far pass of an nx x ny strip on
 * a device with `sms` SMs when `listed_ranges` row ranges were listed by the last probed pair; 0 = plain far -> general order. */
int kob_policy_conc_sms(int64_t nx, int64_t ny, int32_t sms, int64_t listed_ranges);
/* Linked strips: how often a job had to wait for a neighbour's seam flag since kob_create, and the summed waiting time of those
 * warps (several wait at once: divide by `waits` for the mean).  Synchronises the stream.  Diagnostics. */
int kob_wait_stats(kob_ctx* ctx, uint64_t* waits, double* wait_ms);
int kob_get_dims(const kob_ctx* ctx, int64_t* nx, int64_t* ny, int64_t* ny_global, int64_t* y0);
const char* kob_last_error(const kob_ctx* ctx);              /* ctx may be NULL: last create error */
const char* kob_strerror(int status);
int kob_abi_version(void);

/* Pinned host memory for the end-to-end (host buffer) path. */
int kob_host_alloc(void** p, size_t bytes);
int kob_host_free(void* p);
/* Pinned host memory on the NUMA node of the context's GPU (falls back to kob_host_alloc where the topology is not exposed):
 * with one process per GPU on a two-socket box this keeps every rank's PCIe copies off the inter-socket link. */
int kob_host_alloc_near(kob_ctx* ctx, void** p, size_t bytes);

/* ---- row strips (multi-GPU) ------------------------------------------------------ */

/* Periodic closure within one context (P = 1) is automatic.  With P > 1 each strip must be linked to
 * its lower (y0 - 1) and upper (y0 + ny) neighbour before stepping; the step kernel's edge tiles then
 * store their boundary rows straight into the neighbours' ghost rows (NVLink peer stores) and publish
 * a step flag; no separate exchange pass exists. */
int kob_ipc_export(kob_ctx* ctx, kob_ipc_handle* out);
int kob_ipc_link(kob_ctx* ctx, const kob_ipc_handle* lower, const kob_ipc_handle* upper);
/* Same-process variant (strips on one GPU, or peer-enabled GPUs of one process). */
int kob_link_local(kob_ctx* ctx, kob_ctx* lower, kob_ctx* upper);
/* Ring-wide adaptive step path without help from the caller.  Linked strips must all run the same launch sequence, so the
 * per-context adaptive choice between the single-step kernel and two-step launch pairs is off inside a ring.  After every
 * strip (one process / thread each) has called kob_ring_join with the same `name` (unique per ring on this host, <= 47 chars),
 * its `rank` and the ring size, kob_step agrees on the path across the ring every 64 sub-steps through a small POSIX
 * shared-memory segment: the maximum of the strips' density probes, with hysteresis — every strip switches on the same
 * sub-step.  All strips must be stepped by the same amounts (they must anyway).  Results never depend on the path. */
int kob_ring_join(kob_ctx* ctx, const char* name, int32_t rank, int32_t world);
/* Push this strip's current boundary rows into the linked neighbours' ghost rows (after kob_set_fields /
 * kob_reset / kob_add_nucleus on a linked strip).  Collective in spirit: call on every strip, then kob_sync. */
int kob_halo_refresh(kob_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* KOBAYASHI_C_H */
