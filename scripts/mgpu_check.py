"""Multi-GPU shard invariance (G5) on real GPUs.  Launch with torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        scripts/mgpu_check.py [--kernel fast|strict] [--nx 300] [--ny 257] [--steps 60]

Every rank owns one row strip linked to its ring neighbours through CUDA IPC (edge tiles store straight into the
neighbours' ghost rows over NVLink); rank 0 also runs the same global problem on a single GPU and compares the
gathered strips BITWISE."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crystalgrowth_b200 as cg  # noqa: E402
from crystalgrowth_b200.strips import StripRing, nuclei_positions  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kernel", default="fast")
    ap.add_argument("--precision", default="f32")
    ap.add_argument("--nx", type=int, default=300)
    ap.add_argument("--ny", type=int, default=257)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--noise", type=float, default=0.01)
    ap.add_argument("--dense", action="store_true", help="dense (developed) field: exercises the ring-wide step-path policy")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    kw = dict(precision=a.precision, kernel=a.kernel, seed=77, noise_a=a.noise)
    ring = StripRing(a.nx, a.ny, 1e-4, rank=rank, world=world, device=local, **kw)
    # nuclei on and next to every strip boundary, plus random ones
    pos = nuclei_positions(12, a.nx, a.ny, 5)
    for (y0, ny) in ring.parts:
        pos += [(a.nx // 3, y0), (2 * a.nx // 3, (y0 + ny - 1) % a.ny), (0, (y0 + 1) % a.ny), (a.nx - 1, (y0 - 2) % a.ny)]
    ring.seed_nuclei(pos)
    if a.dense:
        import bench
        phi, t = bench.dense_state(a.nx, ring.ny, ring.y0)
        ring.strip.set_fields(phi, t, np.zeros_like(phi))
        ring.refresh()
    for _ in range(a.steps // 10):
        ring.step(10)
    ring.step(a.steps % 10)
    ring.strip.sync()
    dtype = torch.float64 if a.precision == "f64" else torch.float32
    mine = [torch.from_numpy(x).cuda() for x in ring.strip.fields()]
    ok = True
    for k, name in enumerate(("phi", "T", "theta")):
        sizes = [n for (_, n) in ring.parts]
        bufs = [torch.empty((n, a.nx), dtype=dtype, device="cuda") for n in sizes]
        dist.all_gather(bufs, mine[k]) if len(set(sizes)) == 1 else _uneven_gather(bufs, mine[k], rank, world)
        if rank == 0:
            got = torch.cat(bufs, 0).cpu().numpy()
            if k == 0:
                single = cg.Kobayashi(a.nx, a.ny, 1e-4, device=local, **kw)
                single.clear()
                for (x, y) in pos:
                    single.add_nucleus(x, y)
                if a.dense:
                    phi, t = bench.dense_state(a.nx, a.ny, 0)
                    single.set_fields(phi, t, np.zeros_like(phi))
                single.step(a.steps)
                want = single.fields()
                print(f"[mgpu_check] step paths: ring rank 0 {ring.strip.path_stats()}, single GPU {single.path_stats()}", flush=True)
            same = np.array_equal(got.view(np.uint8), want[k].view(np.uint8))
            print(f"[mgpu_check] world={world} {a.kernel}/{a.precision} {a.nx}x{a.ny} steps={a.steps} {name}: "
                  f"bitwise={same} maxabs={np.abs(got.astype(np.float64) - want[k].astype(np.float64)).max():.3e}", flush=True)
            ok &= same
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    ring.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


def _uneven_gather(bufs, mine, rank, world):
    for r in range(world):
        if r == rank:
            bufs[r].copy_(mine)
        dist.broadcast(bufs[r], r)


if __name__ == "__main__":
    main()
