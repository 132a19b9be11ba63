"""Row-strip decomposition of the torus over the GPUs of one box: one process per GPU, one strip per process.

The reference wraps indices modulo the grid (src/Kobayashi.cpp:133-136); across GPUs that wrap becomes a ring
of strips.  The step kernel's edge tiles store their boundary rows straight into the neighbours' ghost rows
(peer memory mapped through CUDA IPC, i.e. NVLink stores) and publish a per-step flag, so there is no separate
exchange pass and no data-path collective.  `torch.distributed` (NCCL on the GPU box, gloo in CPU tests) is
used only for plumbing: exchanging the 128-byte IPC handles once, barriers, and gathering results.

Pure host logic (partition / neighbour ranks / handle exchange) is importable and testable without a GPU.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple


def partition(ny_global: int, world: int) -> List[Tuple[int, int]]:
    """[(y0, ny)] per rank: contiguous row strips, remainder rows to the LAST strips."""
    if world < 1 or ny_global < 2 * world:
        raise ValueError(f"cannot cut {ny_global} rows into {world} strips of at least 2 rows")
    base, rem = divmod(ny_global, world)
    out, y0 = [], 0
    for r in range(world):
        ny = base + (1 if r >= world - rem else 0)
        out.append((y0, ny))
        y0 += ny
    return out


def ring_neighbours(rank: int, world: int) -> Tuple[int, int]:
    """(lower, upper) ranks: the strip holding row y0-1 and the one holding row y0+ny, periodic closure."""
    return (rank - 1) % world, (rank + 1) % world


def philox4x32_10(ctr: Sequence[int], key: Sequence[int]) -> List[int]:
    """Host-side Philox4x32-10 (Salmon et al. 2011) for workload set-up; same function as kob_math.h."""
    M0, M1, W0, W1, mask = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xffffffff
    c, k = [int(x) & mask for x in ctr], [int(x) & mask for x in key]
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & mask, (p0 >> 32) ^ c[3] ^ k[1], p0 & mask]
        k = [(k[0] + W0) & mask, (k[1] + W1) & mask]
    return c


def nuclei_positions(n: int, nx: int, ny_global: int, seed: int, philox: Callable = philox4x32_10) -> List[Tuple[int, int]]:
    """Deterministic multi-seed layout of the C3-C5 workloads: nucleus k at
    (8 + w0 mod (nx-16), 8 + w1 mod (ny-16)) with (w0, w1, ..) = Philox4x32-10(ctr=(k,0,0,0), key=seed)."""
    out = []
    for k in range(n):
        w = philox([k, 0, 0, 0], [seed & 0xffffffff, (seed >> 32) & 0xffffffff])
        out.append((8 + w[0] % (nx - 16), 8 + w[1] % (ny_global - 16)))
    return out


def exchange_blobs(blob: bytes, rank: int, world: int, device=None) -> List[bytes]:
    """all_gather of fixed-size byte blobs over the default process group (nccl -> CUDA tensor, gloo -> CPU)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return [blob]
    dev = device if (device is not None and dist.get_backend() == "nccl") else "cpu"
    mine = torch.tensor(list(blob), dtype=torch.uint8, device=dev)
    outs = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(outs, mine)
    return [bytes(o.cpu().tolist()) for o in outs]


def next_path_mode(mode: int, dense_fraction: float, to_single: float, to_pairs: float) -> int:
    """Ring-wide step-path policy with hysteresis: 1 = two-step launch pairs (sparse field), 0 = single-step kernel."""
    if mode == 1 and dense_fraction > to_single:
        return 0
    if mode == 0 and dense_fraction < to_pairs:
        return 1
    return mode


class StripRing:
    """This rank's strip of an nx x ny_global torus, linked to its ring neighbours.

    make_strip(y0, ny) must return an object with the strip interface of crystalgrowth_b200.Kobayashi
    (ipc_export, ipc_link, halo_refresh, sync, step, ...); the default builds the CUDA strip.
    """

    def __init__(self, nx: int, ny_global: int, timeStep: float = 1e-4, *, rank: int = 0, world: int = 1,
                 device: int = 0, make_strip: Callable | None = None, **kw):
        self.nx, self.ny_global, self.rank, self.world = nx, ny_global, rank, world
        self.parts = partition(ny_global, world)
        self.y0, self.ny = self.parts[rank]
        self.lower_rank, self.upper_rank = ring_neighbours(rank, world)
        if make_strip is None:
            from .kobayashi import Kobayashi

            def make_strip(y0, ny):
                return Kobayashi(nx, ny, timeStep, device=device, ny_global=ny_global, y0=y0, **kw)
        self.strip = make_strip(self.y0, self.ny)
        self._mode, self._since_vote = 1, 0                         # ring-wide step path: pairs first (as the library does)
        if world > 1:
            import torch
            import torch.distributed as dist
            handles = exchange_blobs(self.strip.ipc_export(), rank, world,
                                     device=torch.device("cuda", device) if torch.cuda.is_available() else None)
            self.strip.ipc_link(handles[self.lower_rank], handles[self.upper_rank])
            # ring-wide step-path agreement inside the library (kob_ring_join): rank 0 names the ring, everybody joins
            self._lib_policy = False
            if hasattr(self.strip, "ring_join") and getattr(self.strip, "kernel", "") == "fast":
                import os
                import time
                token = (f"{os.getpid()}_{time.time_ns() & 0xffffffffff:x}".encode() + b"\0" * 48)[:48]
                name = exchange_blobs(token, rank, world, device=torch.device("cuda", device) if torch.cuda.is_available() else None)[0]
                self.strip.ring_join(name.rstrip(b"\0").decode(), rank, world)
                self._lib_policy = True
            dist.barrier()
            self.refresh()

    def refresh(self):
        """After host-side writes (nuclei, set_fields, reset): push boundary rows to the neighbours' ghost rows."""
        if self.world > 1:
            import torch.distributed as dist
            self.strip.sync()
            dist.barrier()          # everybody's interior is written ...
            self.strip.halo_refresh()
            self.strip.sync()
            dist.barrier()          # ... and everybody's ghost rows are current
        else:
            self.strip.sync()

    def seed_nuclei(self, positions: Sequence[Tuple[int, int]]):
        """clear + _createNucleus at GLOBAL positions; each strip keeps its share."""
        self.strip.clear()
        for (x, y) in positions:
            self.strip.add_nucleus(x, y)
        self.refresh()

    # Linked strips must run the same launch sequence, so the library's per-context adaptive choice between the single-step
    # kernel and two-step launch pairs is off inside a ring.  Strips that joined the ring in the library (kob_ring_join, the
    # normal case) agree on the path by themselves; for strip objects without it the same policy is run here: every
    # POLICY_CHUNK sub-steps the strips' density probes are max-reduced and every rank switches the mode on the same sub-step.
    POLICY_CHUNK = 64
    TO_SINGLE, TO_PAIRS = 0.04, 0.03

    def step(self, n: int = 1):
        if (self.world == 1 or getattr(self, "_lib_policy", False) or not hasattr(self.strip, "set_path_mode")
                or getattr(self.strip, "kernel", "") != "fast"):
            self.strip.step(n)                                      # the library agrees on the path across the ring by itself
            return
        import torch
        import torch.distributed as dist
        while n > 0:
            k = min(n, self.POLICY_CHUNK - self._since_vote)
            self.strip.step(k)
            n -= k
            self._since_vote += k
            if self._since_vote >= self.POLICY_CHUNK:
                self._since_vote = 0
                self.strip.sync()                                   # the asynchronous density probe has landed (and is folded in)
                t = torch.tensor([self.strip.path_stats()["dense_fraction"]], dtype=torch.float64,
                                 device="cuda" if dist.get_backend() == "nccl" else "cpu")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                self._mode = next_path_mode(self._mode, float(t[0]), self.TO_SINGLE, self.TO_PAIRS)
                self.strip.set_path_mode(self._mode)

    def close(self):
        if self.world > 1:
            import torch.distributed as dist
            self.strip.sync()
            dist.barrier()          # nobody unmaps memory a neighbour may still be storing into
        self.strip.close()

    def load_checkpoint(self, path: str):
        """Resume this rank's strip from its KOBCKPT1 file and bring the ring's ghost rows / theta flags up to date."""
        self.strip.load_checkpoint(path)
        self.refresh()
