"""A long dendrite-growth run on N GPUs (BASELINE configs[4]: weak scaling, n x n cells and `--nuclei` nuclei per GPU on one torus of
n x N*n), through the strip ring with the in-library ring-wide step-path policy.  Prints a markdown table (rank 0).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/long_run_ring.py --edge 8192 --steps 40000 --chunk 4000
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import crystalgrowth_b200 as cg  # noqa: E402,F401
from crystalgrowth_b200.strips import StripRing, nuclei_positions  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--edge", dest="n", type=int, default=8192)
ap.add_argument("--nuclei", type=int, default=64)
ap.add_argument("--steps", type=int, default=40000)
ap.add_argument("--chunk", type=int, default=4000)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
SEED = 20260101
ring = StripRing(a.n, a.n * world, 1e-4, rank=rank, world=world, device=local, kernel="fast", seed=SEED, noise_a=0.01)
sim = ring.strip
ring.seed_nuclei(nuclei_positions(a.nuclei * world, a.n, a.n * world, SEED))


def allred(vals, op):
    t = torch.tensor(vals, dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=op)
    return [float(x) for x in t]


if rank == 0:
    print(f"# {a.n} x {a.n * world} torus on {world} GPU(s) ({a.n}^2 cells, {a.nuclei} nuclei per GPU), Philox noise a = 0.01, j = 6: "
          f"{a.steps} sub-steps in chunks of {a.chunk}")
    print("| sub-steps done | Gcell-updates/s, all GPUs (chunk) | per GPU | paired / single sub-steps (rank 0) | density probe (rank 0) | solid fraction | seam waits in chunk (all ranks) |")
    print("|---|---|---|---|---|---|---|")
done, t_all = 0, 0.0
while done < a.steps:
    s0, w0 = sim.path_stats(), sim.wait_stats()
    if world > 1:
        dist.barrier()
    ms = sim.step_timed(a.chunk)
    if world > 1:
        dist.barrier()
    ms = allred([ms], dist.ReduceOp.MAX)[0]
    s1, w1 = sim.path_stats(), sim.wait_stats()
    done += a.chunk
    t_all += ms
    solid = allred([float((sim.phi() > 0.5).mean())], dist.ReduceOp.SUM)[0] / world
    waits = allred([w1["waits"] - w0["waits"]], dist.ReduceOp.SUM)[0]
    if rank == 0:
        rate = a.n * a.n * world * a.chunk / (ms * 1e-3) / 1e9
        print(f"| {done} | {rate:.1f} | {rate / world:.1f} | {s1['paired_steps'] - s0['paired_steps']} / {s1['single_steps'] - s0['single_steps']} | "
              f"{s1['dense_fraction']:.3f} | {solid:.4f} | {int(waits)} |", flush=True)
if rank == 0:
    rate = a.n * a.n * world * done / (t_all * 1e-3) / 1e9
    print(f"\ntotal: {rate:.1f} Gcell-updates/s ({rate / world:.1f} per GPU) over {done} sub-steps ({t_all / 1e3:.2f} s of device time, max over ranks)")
ring.close()
if world > 1:
    dist.destroy_process_group()
