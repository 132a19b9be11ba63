#!/usr/bin/env python
"""bench.py — headline benchmark of the fused Kobayashi step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[2], "16384^2 multi-seed with Philox noise a=0.01", FP32,
64 nuclei; with N > 1 the weak-scaling config[4]: 16384 x 16384 cells PER GPU, row strips of a
16384 x (16384*N) torus, 64*N nuclei, ring-linked through CUDA-IPC peer stores issued by the step kernel.

One "step" = one call of the reference's plugin entry point Kobayashi::iUpdate (src/Kobayashi.cpp:227-239)
= `--substeps` (10) explicit-Euler sub-steps = 10 launches of the fused single-step kernel, or 5 two-step launch pairs
(far pass + general pass; the library picks the path, results are bit-identical — kob_path_stats says which ran).

  value  Gcell-updates/s with the state resident in HBM, CUDA-event timed on the library's stream, max over ranks; the
         measurement (fresh seeded field, W warm-up steps, K timed steps) is run `--repeats` (3) times and the MEDIAN repeat is
         reported (all repeats under `repeats`; the crystals grow, so only identical step ranges are comparable)
  e2e    same metric through the host-buffer plugin call: every step copies phi, T, theta from pinned host
         memory to the device (kob_set_fields), runs the sub-steps, and reads phi and T back (kob_get_fields);
         `e2e_plugin` is the call shape of INTEGRATION.md §1 (state resident, kob_update + asynchronous phi readback)
  roofline  the headline path's kernel: `achieved` = bytes the launch has to move (phi and T read once + written once:
         16 B per cell per launch, FP32) / launch time, `frac` = achieved / measured copy peak (<= 1); with two-step launch
         pairs a launch covers TWO cell-updates per cell, so the figure in SURVEY §8d's units (16 B per cell-UPDATE) is
         reported as `frac_sec8d_units`.  `roofline.single_step` is the §8d kernel proper (one sub-step per launch,
         KOB_PATH_SINGLE) timed the same way; `roofline.dense_field` the same kernel on a developed field.
  cpu_baseline  the reference's own CPU loop (oracle/_ref when built, else the oracle port), 1 thread (the
         reference is single threaded), timed on this box on bounded samples: 1024^2 cold (the `value`), 4096^2 x 10
         sub-steps cold, and 1024^2 from a warm (500 sub-step) checkpoint — subnormals make the warm loop slower
`--impl reference` times only that CPU loop (the reference arm).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20260101
FALLBACK_HBM_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--substeps", type=int, default=10, help="sub-steps per step (iUpdate does 10)")
    ap.add_argument("--n", type=int, default=16384, help="grid edge: nx = n, ny = n per GPU")
    ap.add_argument("--kernel", default="fast", choices=["fast", "strict"])
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--noise", type=float, default=0.01)
    ap.add_argument("--nuclei", type=int, default=64, help="nuclei per GPU")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--repeats", type=int, default=3, help="the timed region is repeated this many times; the median is reported")
    ap.add_argument("--no-single", action="store_true", help="skip the single-step-kernel leg (roofline.single_step)")
    ap.add_argument("--no-invariance", action="store_true", help="N > 1: skip the bitwise shard-invariance check before the timed region")
    ap.add_argument("--strong-secondary", type=int, default=65536, help="N > 1: edge of the strong-scaling secondary measurement (0 = off)")
    ap.add_argument("--cpu-n", type=int, default=1024, help="edge of the CPU-baseline sample grid")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--field", default="seeded", choices=["seeded", "dense"],
                    help="seeded: the BASELINE workload (Philox-placed nuclei); dense: every cell on a diffuse interface "
                         "(worst case for the data-dependent path, reported as roofline.dense_field)")
    ap.add_argument("--no-dense", action="store_true", help="skip the secondary dense-field measurement")
    ap.add_argument("--strong", type=int, default=0,
                    help="strong scaling: a fixed STRONG x STRONG torus (BASELINE configs[3]: 65536) cut into row strips")
    return ap.parse_args()


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        d = json.load(open(p))
        for k in ("hbm_gbs", "hbm_GBs", "hbm_gb_s"):
            if k in d:
                return float(d[k]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------ CPU reference leg
def cpu_reference(n, substeps, state=None):
    """The reference's own CPU loop on an n x n sample of the workload, single threaded like the reference.  `state` =
    (phi, T, theta) to start from (a warm checkpoint), else the seeded cold start.  Returns (kind, one_step, sample)."""
    from oracle import pyoracle as po      # checker / baseline only — never on the product path
    po.build()
    if po.ref_available(32):
        kind = "reference"
        sim = po.Reference(n, n, 1e-4, prec=32)
    else:
        kind = "port"
        sim = po.Oracle(n, n, po.default_params(), prec=32, math=po.MATH_LIBM, threads=1)
    step = sim.step
    from crystalgrowth_b200.strips import nuclei_positions
    import numpy as np
    if state is not None:
        sim.set_fields(*state)
    else:
        # several nuclei, like the GPU workload's density (64 per 16384^2 is sparse; keep >= 4 on the sample)
        z = np.zeros((n, n), np.float32)
        sim.set_fields(z, z, z)
        for (x, y) in nuclei_positions(max(4, 64 * n * n // (16384 * 16384)), n, n, SEED):
            sim.add_nucleus(x, y)

    def one_step():
        t0 = time.perf_counter()
        step(substeps)
        return time.perf_counter() - t0

    return kind, one_step, f"{n}x{n} FP32, {substeps} sub-steps per step, 1 thread, same parameters (noise term absent in the reference)"


def cpu_baseline_legs(a, warm_state):
    """the reference loop on 1024^2 cold (the headline baseline), on 4096^2 x 10 sub-steps cold, and on 1024^2 from a
    warm checkpoint (`warm_state`: the same seeded sample after 500 sub-steps, produced by the GPU library) — the far-field
    tails of a developed field are subnormal, which slows the CPU loop down."""
    kind, one_step, sample = cpu_reference(a.cpu_n, a.substeps)
    one_step()
    t_cpu, n_cpu = 0.0, 0
    while t_cpu < 8.0 and n_cpu < 200:
        t_cpu += one_step()
        n_cpu += 1
    cpu = {"value": a.cpu_n * a.cpu_n * a.substeps * n_cpu / t_cpu / 1e9, "unit": "Gcell/s", "cores": 1, "kind": kind,
           "sample": sample + f"; cold start, {n_cpu} steps in {t_cpu:.1f} s", "host_cores_available": os.cpu_count(), "legs": {}}
    cpu["legs"][f"{a.cpu_n}x{a.cpu_n} cold"] = cpu["value"]
    big = 4096
    _, big_step, _ = cpu_reference(big, a.substeps)
    tb = big_step()
    cpu["legs"][f"{big}x{big} cold, {a.substeps} sub-steps"] = big * big * a.substeps / tb / 1e9
    if warm_state is not None:
        _, warm_step, _ = cpu_reference(a.cpu_n, a.substeps, state=warm_state)
        tw, nw = 0.0, 0
        while tw < 4.0 and nw < 50:
            tw += warm_step()
            nw += 1
        cpu["legs"][f"{a.cpu_n}x{a.cpu_n} warm (from sub-step 500), {nw} steps"] = a.cpu_n * a.cpu_n * a.substeps * nw / tw / 1e9
    return cpu


def run_reference_arm(a, rank):
    if rank != 0:
        return
    kind, one_step, sample = cpu_reference(a.cpu_n, a.substeps)
    for _ in range(a.warmup):
        one_step()
    t = 0.0
    for _ in range(a.steps):
        t += one_step()
    cells = a.cpu_n * a.cpu_n * a.substeps * a.steps
    v = cells / t / 1e9
    line = {"impl": "reference", "metric": "Gcell-updates/s", "value": v, "unit": "Gcell/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * t / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(a, a.gpus),
            "cpu_baseline": {"value": v, "unit": "Gcell/s", "cores": 1, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "Gcell/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "host_cores_available": os.cpu_count()}
    print(json.dumps(line), flush=True)


def workload_config(a, world):
    if a.strong:
        return {"workload": f"{a.strong}x{a.strong} torus strong-scaled over {world} GPU(s) ({a.strong}x{a.strong // world} cells per GPU), "
                            f"{a.nuclei * 16} Philox-placed nuclei, Philox noise a={a.noise}, Kobayashi-1993 defaults j=6, dt=1e-4",
                "baseline_config": "configs[3] 65536^2 strong scaling, row strips over 1/2/4/8 GPUs", "nx": a.strong,
                "ny_per_gpu": a.strong // world, "substeps_per_step": a.substeps, "kernel": a.kernel, "precision": a.precision,
                "parallelism": f"row strips x{world}, in-kernel NVLink peer stores for the 2-row halo" if world > 1 else "single GPU",
                "l2_policy": "inputs larger than L2; no flush needed"}
    return {"workload": f"{a.n}x{a.n * world} torus ({a.n}x{a.n} cells per GPU), {a.nuclei * world} Philox-placed nuclei, "
                        f"Philox noise a={a.noise}, Kobayashi-1993 defaults j=6, dt=1e-4",
            "baseline_config": "configs[2] 16384^2 multi-seed, a=0.01 (N=1); configs[4] weak scaling 16384x16384 per GPU (N>1)",
            "nx": a.n, "ny_per_gpu": a.n, "substeps_per_step": a.substeps, "kernel": a.kernel, "precision": a.precision,
            "parallelism": f"row strips x{world}, in-kernel NVLink peer stores for the 2-row halo" if world > 1 else "single GPU",
            "l2_policy": "inputs larger than L2 (4 GiB of phi/T ping-pong per GPU vs 126 MB L2); no flush needed"}


def dense_state(nx, ny, y0):
    """A developed-field stand-in: every cell sits on a diffuse interface (0.05 <= phi <= 0.95, non-flat gradient
    almost everywhere, T in [-0.3, 0.3]), so the data-dependent part of the step (angle re-assignment, anisotropy,
    m(T), noise draw) runs for every cell — the worst case, where the seeded BASELINE workload is the best case."""
    import numpy as np
    x = np.arange(nx, dtype=np.float64)
    y = np.arange(y0, y0 + ny, dtype=np.float64)
    sx, cy = np.sin(2 * np.pi * x / 97.0).astype(np.float32), np.cos(2 * np.pi * y / 61.0).astype(np.float32)
    cx, sy = np.cos(2 * np.pi * x / 53.0).astype(np.float32), np.sin(2 * np.pi * y / 131.0).astype(np.float32)
    phi = np.multiply.outer(cy, sx)
    phi *= np.float32(0.45)
    phi += np.float32(0.5)
    t = np.multiply.outer(sy, cx)
    t *= np.float32(0.3)
    return phi, t


# ------------------------------------------------------------------------------------------ native arm
def traffic_entry(key):
    """DRAM bytes per launch of the committed ncu capture (profiles/traffic.json), with the commit it was captured at."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        e = d.get(key)
        if isinstance(e, dict):
            return e.get("bytes"), e.get("captured_at"), e.get("source")
        return e, d.get("_captured_at"), None
    except Exception:
        return None, None, None


def median(xs):
    s = sorted(xs)
    return s[len(s) // 2] if len(s) % 2 else 0.5 * (s[len(s) // 2 - 1] + s[len(s) // 2])


def shard_invariance(torch, dist, cg, StripRing, nuclei_positions, rank, world, local):
    """N > 1, before the timed region: a 2500 x 2063 torus (strip seams at rows that divide nothing) is advanced 100 sub-steps
    by the linked strips and, on rank 0, by a single-GPU context; the gathered fields must be BITWISE equal."""
    import numpy as np
    nx, ny, steps = 2500, 2063, 100
    kw = dict(precision="f32", kernel="fast", seed=77, noise_a=0.01)
    ring = StripRing(nx, ny, 1e-4, rank=rank, world=world, device=local, **kw)
    pos = nuclei_positions(12, nx, ny, 5)
    for (y0, n_) in ring.parts:
        pos += [(nx // 3, y0), (2 * nx // 3, (y0 + n_ - 1) % ny), (0, (y0 + 1) % ny), (nx - 1, (y0 - 2) % ny)]
    ring.seed_nuclei(pos)
    for _ in range(steps // 10):
        ring.step(10)
    ring.strip.sync()
    ok = True
    mine = [torch.from_numpy(x).cuda() for x in ring.strip.fields()]
    want = None
    if rank == 0:
        single = cg.Kobayashi(nx, ny, 1e-4, device=local, **kw)
        single.clear()
        for (x, y) in pos:
            single.add_nucleus(x, y)
        single.step(steps)
        want = single.fields()
        single.close()
    for k in range(3):
        bufs = [torch.empty((n_, nx), dtype=torch.float32, device="cuda") for (_, n_) in ring.parts]
        for r in range(world):
            if r == rank:
                bufs[r].copy_(mine[k])
            dist.broadcast(bufs[r], r)
        if rank == 0:
            got = torch.cat(bufs, 0).cpu().numpy()
            ok &= bool(np.array_equal(got.view(np.uint8), want[k].view(np.uint8)))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    ring.close()
    return "bitwise" if int(flag.item()) else "FAILED"


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference_arm(a, rank)
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist
    import crystalgrowth_b200 as cg
    from crystalgrowth_b200.strips import StripRing, nuclei_positions

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(vals):
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    elem = 8 if a.precision == "f64" else 4
    peak, peak_src = hbm_peak()
    invariance = None
    if world > 1 and not a.no_invariance and a.kernel == "fast" and a.precision == "f32":
        invariance = shard_invariance(torch, dist, cg, StripRing, nuclei_positions, rank, world, local)

    nx, nyg = (a.strong, a.strong) if a.strong else (a.n, a.n * world)
    if a.strong:
        a.no_e2e = True            # 80 GiB of pinned host memory for a full-state round trip is not a sensible call
    ring = StripRing(nx, nyg, 1e-4, rank=rank, world=world, device=local, precision=a.precision, kernel=a.kernel,
                     seed=SEED, noise_a=a.noise)
    sim = ring.strip
    nuclei = nuclei_positions(a.nuclei * (16 if a.strong else world), nx, nyg, SEED)

    def fresh_field():
        """The workload's initial state (every repeat starts from it: the crystals grow, so later steps cost more)."""
        ring.seed_nuclei(nuclei)
        if a.field == "dense":
            sim.set_fields(*dense_state(nx, ring.ny, ring.y0), None)
            ring.refresh()

    total_cells_per_step = nx * nyg * a.substeps

    def timed_region():
        """K steps, CUDA events on the library's stream, barrier + synchronize on both sides; max over ranks."""
        barrier()
        w0 = time.perf_counter()
        ms = sim.step_timed(a.steps * a.substeps)             # events bracket exactly K*substeps sub-steps
        barrier()
        wall = 1e3 * (time.perf_counter() - w0)
        return allmax([ms, wall])

    # ---- `repeats` x (fresh seeded field, W warm-up steps, the timed region of K steps): every repeat times the SAME steps ----
    sampler = ClockSampler(local)
    sampler.start()
    reps, launches, paired, single = [], 0, 0, 0
    for _ in range(max(1, a.repeats)):
        fresh_field()
        for _ in range(max(a.warmup, 3)):
            sim.step(a.substeps)
        sim.sync()
        l0, p0 = sim.launch_count, sim.path_stats()
        w0 = sim.wait_stats() if world > 1 else None
        reps.append(timed_region())
        p1 = sim.path_stats()
        w1 = sim.wait_stats() if world > 1 else None
        launches = sim.launch_count - l0                      # kernels launched inside ONE timed region
        paired, single = p1["paired_steps"] - p0["paired_steps"], p1["single_steps"] - p0["single_steps"]
        conc_pairs = p1["concurrent_pairs"] - p0["concurrent_pairs"]
        listed_frac = p1["dense_fraction"]
    clocks = sampler.result()
    seam_waits = None
    if world > 1:
        # per rank, last repeat's timed region: how often a seam job waited for a neighbour's flag, summed waiting time of those warps
        t = torch.tensor([w1["waits"] - w0["waits"], w1["wait_ms"] - w0["wait_ms"]], dtype=torch.float64, device="cuda")
        g = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(g, t)
        seam_waits = {"per_rank_waits": [int(x[0]) for x in g], "per_rank_summed_wait_ms": [round(float(x[1]), 3) for x in g],
                      "what": "seam jobs that found the neighbour's flag not yet published, and the summed time those warps waited, in the last timed region"}
    ms_list = [r[0] for r in reps]
    ms_med = median(ms_list)
    wall_med = median([r[1] for r in reps])
    value = total_cells_per_step * a.steps / (ms_med * 1e-3) / 1e9
    rep_values = [total_cells_per_step * a.steps / (m * 1e-3) / 1e9 for m in ms_list]
    # The step path: single-step launches (one kernel per sub-step) and/or two-step launch pairs (far pass + general
    # pass = 2 kernels per 2 sub-steps).  A "launch" below is the unit that is timed: one sub-step for the single-step
    # kernel, one PAIR (two sub-steps) for the two-step path.
    two_step = paired >= single
    sub_per_launch = 2 if two_step else 1
    launch_ms = ms_med / (a.steps * a.substeps) * sub_per_launch
    cells = nx * ring.ny
    # bytes one launch has to move: phi and T read once + written once (16 B per cell, FP32)
    achieved = cells * 4 * elem / (launch_ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": None, "peak_source": peak_src,
            "kernel": ("kob_far2 + kob_step_fast2 (launch pair = 2 sub-steps)" if two_step else f"kob_step_{a.kernel}"),
            "algorithmic_bytes_per_cell_per_launch": 4 * elem, "substeps_per_launch": sub_per_launch, "launch_ms": launch_ms,
            "paired_substeps": paired, "single_substeps": single,
            "pairs_with_concurrent_general_pass": conc_pairs, "listed_range_fraction_last_probe": listed_frac,
            "frac_sec8d_units": achieved * sub_per_launch / peak,
            "note": ("frac = bytes the launch must move (16 B per cell) / launch time / measured copy peak; a two-step launch pair "
                     "advances every cell by TWO sub-steps for those bytes, so in SURVEY §8d's units (16 B per cell-UPDATE) the "
                     "same launch reads frac_sec8d_units" if two_step else
                     "one sub-step per launch: frac is SURVEY §8d's figure (16 B per cell-update / measured copy peak)")}
    key = f"{a.kernel}{'2' if two_step else ''}_{a.precision}_{a.n}"
    tr, tr_at, _ = traffic_entry(key)
    if tr:
        roof["traffic"] = tr
        roof["traffic_captured_at"] = tr_at
        roof["dram_achieved"] = tr / (launch_ms * 1e-3) / 1e9
        roof["dram_frac"] = roof["dram_achieved"] / peak

    def timed_leg(nsteps, warm, repeats):
        """median launch time (ms per sub-step launch) of `nsteps` sub-steps, after `warm` sub-steps"""
        sim.step(warm)
        sim.sync()
        out = []
        for _ in range(repeats):
            out.append(sim.step_timed(nsteps) / nsteps)
        return median(out), out

    # ---- the SURVEY §8d kernel proper: one sub-step per launch on the same (seeded) workload ----
    if world == 1 and a.kernel == "fast" and a.field == "seeded" and not a.no_single and not a.strong:
        s_all = []
        sim.set_path_mode(0)
        for _ in range(max(1, a.repeats)):                    # same protocol as the headline: fresh field, W warm-up steps, K timed steps
            fresh_field()
            sim.step(max(a.warmup, 3) * a.substeps)
            sim.sync()
            s_all.append(sim.step_timed(a.steps * a.substeps) / (a.steps * a.substeps))
        s_ms = median(s_all)
        sim.set_path_mode(2)
        s_ach = cells * 4 * elem / (s_ms * 1e-3) / 1e9
        roof["single_step"] = {"kernel": "kob_step_fast", "launch_ms": s_ms, "value": cells / (s_ms * 1e-3) / 1e9, "unit": "Gcell/s",
                               "achieved": s_ach, "frac": s_ach / peak, "repeats_launch_ms": s_all,
                               "what": "KOB_PATH_SINGLE: one fused sub-step per launch (the north-star kernel), same workload, same timing"}
        tr1, tr1_at, _ = traffic_entry(f"fast_{a.precision}_{a.n}")
        if tr1:
            roof["single_step"].update({"traffic": tr1, "traffic_captured_at": tr1_at, "dram_frac": tr1 / (s_ms * 1e-3) / 1e9 / peak})

    # ---- secondary: the same kernel on a dense (developed) field, N = 1 only ----
    warm_state = None
    if world == 1 and a.kernel == "fast" and a.field == "seeded" and not a.no_dense and not a.strong:
        saved = sim.fields()
        ctr = sim.step_counter
        sim.set_fields(*dense_state(nx, ring.ny, 0), np.zeros((ring.ny, nx), np.float32))
        sim.set_path_mode(0)                                  # the developed-field path is the single-step kernel
        d_steps = 5
        del_ms, d_all = timed_leg(d_steps * a.substeps, 3 * a.substeps, max(1, a.repeats))
        sim.set_path_mode(2)
        phi_d = sim.phi()
        roof["dense_field"] = {"value": cells / (del_ms * 1e-3) / 1e9, "unit": "Gcell/s", "launch_ms": del_ms,
                               "frac": cells * 4 * elem / (del_ms * 1e-3) / 1e9 / peak, "repeats_launch_ms": d_all,
                               "interface_fraction_after": float(((phi_d > 0.01) & (phi_d < 0.99)).mean()),
                               "what": "same grid started with every cell on a diffuse interface (angle re-assignment, anisotropy, m(T) "
                                       "and the noise draw run for every cell); 30 warm-up launches, then 50 launches per repeat — the "
                                       "field saturates as it evolves (interface_fraction_after), later repeats see more plateau rows"}
        del phi_d
        sim.set_fields(*saved)
        sim.step_counter = ctr
        del saved

    # ---- N > 1: the developed-field regime across the ring (BASELINE configs[4] lives there): every strip starts from the dense
    # state of its rows and all strips run the single-step kernel (same path mode on every strip at the same sub-step) ----
    if world > 1 and a.kernel == "fast" and a.field == "seeded" and not a.no_dense and not a.strong:
        sim.set_fields(*dense_state(nx, ring.ny, ring.y0), np.zeros((ring.ny, nx), np.float32))
        ring.refresh()
        sim.set_path_mode(0)
        sim.step(3 * a.substeps)
        d_all = []
        for _ in range(max(1, a.repeats)):
            barrier()
            ms_d = sim.step_timed(5 * a.substeps)
            barrier()
            d_all.append(allmax([ms_d])[0] / (5 * a.substeps))
        sim.set_path_mode(2)
        del_ms = median(d_all)
        roof["dense_field"] = {"value": cells * world / (del_ms * 1e-3) / 1e9, "unit": "Gcell/s", "launch_ms": del_ms,
                               "frac": cells * 4 * elem / (del_ms * 1e-3) / 1e9 / peak, "repeats_launch_ms": d_all,
                               "what": f"weak scaling in the developed-field regime: every one of the {world} linked strips starts with every cell on a "
                                       "diffuse interface and runs the single-step kernel; 30 warm-up launches, then 50 launches per repeat; "
                                       "max over ranks; value = all strips' cell-updates / that time"}

    # ---- end to end through the host-buffer plugin call ----
    e2e, e2e_plugin = None, None
    if not a.no_e2e:
        nbytes = nx * ring.ny * elem
        bufs = [sim.host_alloc_near(nbytes) for _ in range(5)]    # phi, T, theta in; phi, T out — pinned, NUMA-local to the GPU
        sim.get_fields_into(bufs[0], bufs[1], bufs[2])    # a valid evolved state as the host-resident input
        ring.refresh()
        barrier()
        for it in range(a.e2e_steps + 1):
            if it == 1:                                   # first iteration is warm-up
                sim.wait_fields()
                barrier()
                e0 = time.perf_counter()
            sim.set_fields_from(bufs[0], bufs[1], bufs[2])    # H2D of this step's inputs (asynchronous, library stream)
            if world > 1:
                ring.refresh()
            sim.step(a.substeps)
            # D2H of this step's result: snapshot on the device, copy on the library's second stream — it overlaps the NEXT
            # step's H2D (PCIe is full duplex); kob_wait_fields before the clock stops
            sim.get_fields_async(bufs[3], bufs[4], None)
        sim.wait_fields()
        barrier()
        e_ms = allmax([1e3 * (time.perf_counter() - e0)])[0]
        e2e = {"value": total_cells_per_step * a.e2e_steps / (e_ms * 1e-3) / 1e9, "unit": "Gcell/s",
               "h2d_bytes_per_step": 3 * nbytes * world, "d2h_bytes_per_step": 2 * nbytes * world,
               "steps": a.e2e_steps, "ms_per_step": e_ms / a.e2e_steps,
               "call": "per step: kob_set_fields(phi,T,theta) + kob_step(substeps) + kob_get_fields_async(phi,T); kob_wait_fields at the "
                       "end; pinned NUMA-local host buffers (step n's readback overlaps step n+1's upload)"}
        # the round-1 call shape for comparison on the same box: blocking readback, nothing overlaps
        barrier()
        s0_ = time.perf_counter()
        for it in range(a.e2e_steps):
            sim.set_fields_from(bufs[0], bufs[1], bufs[2])
            if world > 1:
                ring.refresh()
            sim.step(a.substeps)
            sim.get_fields_into(bufs[3], bufs[4], None)
        barrier()
        sm_ = allmax([1e3 * (time.perf_counter() - s0_)])[0]
        e2e["serial_value"] = total_cells_per_step * a.e2e_steps / (sm_ * 1e-3) / 1e9
        e2e["serial_call"] = "kob_set_fields + kob_step + blocking kob_get_fields (the round-1 measurement)"
        # the plugin's own call shape: state resident, iUpdate, then the picture's phi read back —
        # asynchronously, double buffered, so that frame n's copy overlaps frame n+1's sub-steps
        barrier()
        frames = a.e2e_steps + 2
        p0_ = time.perf_counter()
        for it in range(frames):
            sim.step(a.substeps)
            sim.get_fields_async(bufs[3 + (it & 1)], None, None)
        sim.wait_fields()
        barrier()
        pm = allmax([1e3 * (time.perf_counter() - p0_)])[0]
        e2e_plugin = {"value": total_cells_per_step * frames / (pm * 1e-3) / 1e9, "unit": "Gcell/s", "h2d_bytes_per_step": 0,
                      "d2h_bytes_per_step": nbytes * world, "steps": frames, "ms_per_step": pm / frames,
                      "call": "kob_update-shaped frame: kob_step(substeps) + kob_get_fields_async(phi) into alternating pinned buffers, "
                              "kob_wait_fields at the end (frame n's readback overlaps frame n+1's sub-steps)"}
        L = cg.load()
        for p_ in bufs:
            L.kob_host_free(p_)

    # ---- CPU baseline (rank 0, N = 1 only): bounded samples; the warm checkpoint comes from the GPU library ----
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        w = cg.Kobayashi(a.cpu_n, a.cpu_n, 1e-4, device=local, kernel="fast")
        w.clear()
        for (x, y) in nuclei_positions(max(4, 64 * a.cpu_n * a.cpu_n // (16384 * 16384)), a.cpu_n, a.cpu_n, SEED):
            w.add_nucleus(x, y)
        w.step(500)
        warm_state = w.fields()
        w.close()
        cpu = cpu_baseline_legs(a, warm_state)

    ring.close()

    # ---- N > 1 secondary: strong scaling of the 65536^2 torus (BASELINE configs[3]), 5 steps ----
    strong = None
    if world > 1 and a.strong_secondary and not a.strong:
        sn = a.strong_secondary
        sring = StripRing(sn, sn, 1e-4, rank=rank, world=world, device=local, precision=a.precision, kernel=a.kernel,
                          seed=SEED, noise_a=a.noise)
        sring.seed_nuclei(nuclei_positions(a.nuclei * 16, sn, sn, SEED))
        for _ in range(3):
            sring.strip.step(a.substeps)
        sring.strip.sync()
        barrier()
        sms = allmax([sring.strip.step_timed(5 * a.substeps)])[0]
        barrier()
        strong = {"value": sn * sn * a.substeps * 5 / (sms * 1e-3) / 1e9, "unit": "Gcell/s", "steps": 5, "ms_per_step": sms / 5,
                  "workload": f"{sn}x{sn} torus cut into {world} row strips ({sn}x{sn // world} cells per GPU), {a.nuclei * 16} nuclei, "
                              f"Philox noise a={a.noise}", "baseline_config": "configs[3] 65536^2 strong scaling"}
        sring.close()

    if rank == 0:
        line = {"metric": "Gcell-updates/s", "value": value, "unit": "Gcell/s", "n_gpus": world, "steps": a.steps,
                "warmup": max(a.warmup, 3), "ms_per_step": ms_med / a.steps, "higher_is_better": True,
                "scaling": "strong" if a.strong else "weak",
                "vs_baseline": None, "dtype": a.precision, "data": "synthetic", "config": workload_config(a, world),
                "repeats": {"n": len(reps), "statistic": "median", "values": rep_values, "min": min(rep_values), "max": max(rep_values)},
                "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "e2e_plugin": e2e_plugin, "gpu_launches": int(launches), "clocks": clocks,
                "wall_ms_per_step": wall_med / a.steps, "pct_of_hbm_roofline": 100.0 * roof["frac"]}
        if invariance is not None:
            line["shard_invariance"] = invariance
            line["shard_invariance_what"] = "2500x2063 torus, 100 sub-steps: the linked strips' gathered phi, T, theta vs one GPU, before the timed region"
        if strong is not None:
            line["strong_65536"] = strong
        if seam_waits is not None:
            line["seam_waits"] = seam_waits
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
