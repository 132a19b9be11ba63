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

  value  Gcell-updates/s with the state resident in HBM, CUDA-event timed on the library's stream, max over ranks
  e2e    same metric through the host-buffer plugin call: every step copies phi, T, theta from pinned host
         memory to the device (kob_set_fields), runs the sub-steps, and reads phi and T back (kob_get_fields)
  roofline  16 B per cell-update (phi and T read once + written once, FP32) x cell-updates per launch /
         avg launch duration (a "launch" is one sub-step, or one two-step PAIR); with pairs the measured DRAM traffic
         (dram_achieved / dram_frac) is reported next to the algorithmic figure, which can exceed the copy peak
  cpu_baseline  the reference's own CPU loop (oracle/_ref when built, else the oracle port), 1 thread (the
         reference is single threaded), timed on this box on a bounded sample
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
def cpu_reference(n, substeps, min_seconds=10.0, max_steps=50):
    """The reference's own CPU loop on an n x n sample of the workload (single nucleus field warmed up so the
    interface is populated), single threaded like the reference.  Returns (Gcell/s, kind, cores, sample, fn)."""
    from oracle import pyoracle as po      # checker / baseline only — never on the product path
    po.build()
    if po.ref_available(32):
        kind = "reference"
        sim = po.Reference(n, n, 1e-4, prec=32)
        step = sim.step
    else:
        kind = "port"
        sim = po.Oracle(n, n, po.default_params(), prec=32, math=po.MATH_LIBM, threads=1)
        step = sim.step
    # several nuclei, like the GPU workload's density (64 per 16384^2 is sparse; keep >= 4 on the sample)
    from crystalgrowth_b200.strips import nuclei_positions
    import numpy as np
    z = np.zeros((n, n), np.float32)
    sim.set_fields(z, z, z)
    for (x, y) in nuclei_positions(max(4, 64 * n * n // (16384 * 16384)), n, n, SEED):
        sim.add_nucleus(x, y)

    def one_step():
        t0 = time.perf_counter()
        step(substeps)
        return time.perf_counter() - t0

    return kind, one_step, f"{n}x{n} FP32, {substeps} sub-steps per step, 1 thread, same parameters (noise term absent in the reference)"


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

    nx, nyg = (a.strong, a.strong) if a.strong else (a.n, a.n * world)
    if a.strong:
        a.no_e2e = True            # 80 GiB of pinned host memory for a full-state round trip is not a sensible call
    ring = StripRing(nx, nyg, 1e-4, rank=rank, world=world, device=local, precision=a.precision, kernel=a.kernel,
                     seed=SEED, noise_a=a.noise)
    sim = ring.strip
    ring.seed_nuclei(nuclei_positions(a.nuclei * (16 if a.strong else world), nx, nyg, SEED))
    if a.field == "dense":
        sim.set_fields(*dense_state(nx, ring.ny, ring.y0), None)
        ring.refresh()
    cells_per_step = nx * ring.ny * a.substeps            # this rank
    total_cells_per_step = nx * nyg * a.substeps

    # ---- warm-up, then the timed region: K steps, CUDA events on the library's stream, max over ranks ----
    for _ in range(max(a.warmup, 3)):
        sim.step(a.substeps)
    sim.sync()
    sampler = ClockSampler(local)
    sampler.start()
    l0, p0 = sim.launch_count, sim.path_stats()
    barrier()
    w0 = time.perf_counter()
    ms = sim.step_timed(a.steps * a.substeps)             # events bracket exactly K*substeps sub-steps
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - w0)
    launches, p1 = sim.launch_count - l0, sim.path_stats()
    clocks = sampler.result()
    t = torch.tensor([ms, wall_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, wall_max = float(t[0]), float(t[1])
    value = total_cells_per_step * a.steps / (ms_max * 1e-3) / 1e9
    elem = 8 if a.precision == "f64" else 4
    peak, peak_src = hbm_peak()
    # The step path: single-step launches (one kernel per sub-step) and/or two-step launch pairs (far pass + general
    # pass = 2 kernels per 2 sub-steps).  A "launch" below is the unit that is timed: one sub-step for the single-step
    # kernel, one PAIR (two sub-steps) for the two-step path.
    paired = p1["paired_steps"] - p0["paired_steps"]
    single = p1["single_steps"] - p0["single_steps"]
    two_step = paired >= single
    sub_per_launch = 2 if two_step else 1
    launch_ms = ms / (a.steps * a.substeps) * sub_per_launch
    # algorithmic bytes: 16 B (FP32) per cell-update x the cell-updates one launch performs
    achieved = nx * ring.ny * 4 * elem * sub_per_launch / (launch_ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": None, "peak_source": peak_src,
            "kernel": ("kob_far2 + kob_step_fast2 (launch pair = 2 sub-steps)" if two_step else f"kob_step_{a.kernel}"),
            "algorithmic_bytes_per_cell": 4 * elem, "substeps_per_launch": sub_per_launch, "launch_ms": launch_ms,
            "paired_substeps": paired, "single_substeps": single,
            "achieved_incl_theta_state": nx * ring.ny * 6 * elem * sub_per_launch / (launch_ms * 1e-3) / 1e9}
    tr = os.path.join(ROOT, "profiles", "traffic.json")   # dram bytes per launch from the committed ncu capture
    try:
        key = f"{a.kernel}{'2' if two_step else ''}_{a.precision}_{a.n}"
        roof["traffic"] = json.load(open(tr)).get(key)
        if two_step and roof["traffic"]:
            # temporal blocking: phi and T cross HBM once per TWO sub-steps, so the algorithmic figure (16 B per
            # cell-update) exceeds what the launch pair actually moves; both are reported
            roof["dram_achieved"] = roof["traffic"] / (launch_ms * 1e-3) / 1e9
            roof["dram_frac"] = roof["dram_achieved"] / peak
            roof["note"] = ("two sub-steps per launch pair: compulsory traffic is 16 B per cell per PAIR = 8 B per cell-update, "
                            "so frac (algorithmic 16 B per cell-update / measured copy peak) can exceed 1; dram_frac is the "
                            "measured DRAM traffic of the pair over the same peak")
    except Exception:
        pass

    # ---- secondary: the same kernel on a dense (developed) field, N = 1 only ----
    if world == 1 and a.kernel == "fast" and a.field == "seeded" and not a.no_dense and not a.strong:
        saved = sim.fields()
        ctr = sim.step_counter
        sim.set_fields(*dense_state(nx, ring.ny, 0), np.zeros((ring.ny, nx), np.float32))
        sim.step(3 * a.substeps)
        sim.sync()
        d_steps = 5
        st0 = sim.path_stats()
        del_ms = sim.step_timed(d_steps * a.substeps) / (d_steps * a.substeps)
        st1 = sim.path_stats()
        roof["dense_field"] = {"value": nx * ring.ny / (del_ms * 1e-3) / 1e9, "unit": "Gcell/s", "launch_ms": del_ms,
                               "frac": nx * ring.ny * 4 * elem / (del_ms * 1e-3) / 1e9 / peak,
                               "single_steps": st1["single_steps"] - st0["single_steps"], "paired_steps": st1["paired_steps"] - st0["paired_steps"],
                               "what": "same grid, every cell on a diffuse interface (angle re-assignment, anisotropy, m(T) and the "
                                       "noise draw run for every cell); 50 launches after 30 warm-up launches"}
        sim.set_fields(*saved)
        sim.step_counter = ctr
        del saved

    # ---- end to end through the host-buffer plugin call ----
    e2e = None
    if not a.no_e2e:
        L = cg.load()
        nbytes = nx * ring.ny * elem
        bufs = []
        for _ in range(5):                                # phi, T, theta in; phi, T out — pinned
            p = C.c_void_p()
            if L.kob_host_alloc(C.byref(p), nbytes) != 0:
                raise SystemExit("pinned host allocation failed")
            bufs.append(p)
        sim.get_fields_into(bufs[0], bufs[1], bufs[2])    # a valid evolved state as the host-resident input
        ring.refresh()
        barrier()
        for it in range(a.e2e_steps + 1):
            if it == 1:                                   # first iteration is warm-up
                barrier()
                e0 = time.perf_counter()
            sim.set_fields_from(bufs[0], bufs[1], bufs[2])
            if world > 1:
                ring.refresh()
            sim.step(a.substeps)
            sim.get_fields_into(bufs[3], bufs[4], None)
        barrier()
        e_ms = 1e3 * (time.perf_counter() - e0)
        te = torch.tensor([e_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": total_cells_per_step * a.e2e_steps / (float(te[0]) * 1e-3) / 1e9, "unit": "Gcell/s",
               "h2d_bytes_per_step": 3 * nbytes * world, "d2h_bytes_per_step": 2 * nbytes * world,
               "steps": a.e2e_steps, "ms_per_step": float(te[0]) / a.e2e_steps,
               "call": "kob_set_fields(phi,T,theta) + kob_step(substeps) + kob_get_fields(phi,T), pinned host buffers"}
        for p in bufs:
            L.kob_host_free(p)

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        kind, one_step, sample = cpu_reference(a.cpu_n, a.substeps)
        one_step()
        t_cpu, n_cpu = 0.0, 0
        while t_cpu < 10.0 and n_cpu < 200:
            t_cpu += one_step()
            n_cpu += 1
        cpu = {"value": a.cpu_n * a.cpu_n * a.substeps * n_cpu / t_cpu / 1e9, "unit": "Gcell/s", "cores": 1, "kind": kind,
               "sample": sample + f"; {n_cpu} steps in {t_cpu:.1f} s", "host_cores_available": os.cpu_count()}

    ring.close()
    if rank == 0:
        line = {"metric": "Gcell-updates/s", "value": value, "unit": "Gcell/s", "n_gpus": world, "steps": a.steps,
                "warmup": max(a.warmup, 3), "ms_per_step": ms_max / a.steps, "higher_is_better": True,
                "scaling": "strong" if a.strong else "weak",
                "vs_baseline": None, "dtype": a.precision, "data": "synthetic", "config": workload_config(a, world),
                "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "wall_ms_per_step": wall_max / a.steps, "pct_of_hbm_roofline": 100.0 * achieved / peak}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
