"""Row-strip host logic (CPU): partition, ring closure, nucleus layout; and a world_size-2 gloo run in which
oracle strips exchange 2 ghost rows per side (the halo depth of the fused step) and must
reproduce the single-domain oracle bit-for-bit (G5 on the CPU side)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, bit_equal


def test_partition_and_ring():
    from crystalgrowth_b200.strips import partition, ring_neighbours
    assert partition(16, 1) == [(0, 16)]
    assert partition(16, 4) == [(0, 4), (4, 4), (8, 4), (12, 4)]
    assert partition(18, 4) == [(0, 4), (4, 4), (8, 5), (13, 5)]        # remainder rows to the last strips
    for ny, w in [(65536, 8), (1001, 8), (17, 8), (131072, 8)]:
        parts = partition(ny, w)
        assert parts[0][0] == 0 and sum(n for _, n in parts) == ny
        assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        assert max(n for _, n in parts) - min(n for _, n in parts) <= 1
    with pytest.raises(ValueError):
        partition(7, 4)
    assert ring_neighbours(0, 4) == (3, 1) and ring_neighbours(3, 4) == (2, 0) and ring_neighbours(0, 1) == (0, 0)
    assert ring_neighbours(0, 2) == (1, 1)


def test_nuclei_positions_deterministic(po):
    from crystalgrowth_b200.strips import nuclei_positions
    a = nuclei_positions(64, 16384, 16384, 20260101, po.philox)
    b = nuclei_positions(64, 16384, 16384, 20260101, po.philox)
    assert a == b and len(set(a)) == 64
    assert all(8 <= x < 16384 - 8 and 8 <= y < 16384 - 8 for x, y in a)
    # growing the torus in y for weak scaling keeps the x coordinates (same Philox words)
    c = nuclei_positions(64, 16384, 32768, 20260101, po.philox)
    assert [x for x, _ in a] == [x for x, _ in c]


def test_host_philox_equals_oracle_philox(po):
    from crystalgrowth_b200.strips import nuclei_positions, philox4x32_10
    for ctr, key in [([0, 0, 0, 0], [0, 0]), ([0xffffffff] * 4, [0xffffffff] * 2), ([5, 6, 7, 8], [9, 10]),
                     ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])]:
        assert philox4x32_10(ctr, key) == po.philox(ctr, key)
    assert nuclei_positions(16, 999, 777, 42) == nuclei_positions(16, 999, 777, 42, po.philox)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _exchange(o, rank, world):
    """Ring exchange of the 2 edge rows of phi, T, theta per side over gloo send/recv."""
    lower, upper = (rank - 1) % world, (rank + 1) % world
    lo = [torch.from_numpy(x) for x in o.edge(0)]     # my lowest two rows  -> lower neighbour's upper ghosts
    hi = [torch.from_numpy(x) for x in o.edge(1)]     # my highest two rows -> upper neighbour's lower ghosts
    send_lo, send_hi = torch.stack(lo), torch.stack(hi)
    recv_lo, recv_hi = torch.empty_like(send_hi), torch.empty_like(send_lo)
    reqs = [dist.isend(send_lo, lower, tag=1), dist.isend(send_hi, upper, tag=2),
            dist.irecv(recv_lo, lower, tag=2), dist.irecv(recv_hi, upper, tag=1)]
    for r in reqs:
        r.wait()
    o.set_ghost(0, *[x.numpy() for x in recv_lo])
    o.set_ghost(1, *[x.numpy() for x in recv_hi])


def _worker(rank, world, port, nx, nyg, steps, prec, noise_a, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from crystalgrowth_b200.strips import StripRing, exchange_blobs, nuclei_positions, partition
    from oracle import pyoracle as po
    y0, ny = partition(nyg, world)[rank]
    p = po.default_params(noise_a=noise_a)
    o = po.Oracle(nx, ny, p, prec=prec, seed=11, ny_global=nyg, y0=y0, reset=False)
    o.clear()
    for (x, y) in nuclei_positions(5, nx, nyg, 3, po.philox) + [(nx // 2, 0), (0, nyg // 2)]:
        o.add_nucleus(x, y)
    # plumbing used by the CUDA strips: fixed-size blob all_gather (here: a fake 128-byte handle)
    blobs = exchange_blobs(bytes([rank]) * 128, rank, world)
    assert [b[0] for b in blobs] == list(range(world))
    for _ in range(steps):
        _exchange(o, rank, world)
        o.step(1)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), *o.fields())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("prec,noise_a,nyg", [(32, 0.0, 48), (32, 0.02, 45), (64, 0.0, 48)])
def test_two_rank_gloo_strips_equal_single_domain(po, tmp_path, prec, noise_a, nyg):
    from crystalgrowth_b200.strips import nuclei_positions, partition
    nx, steps, world = 40, 25, 2
    mp.spawn(_worker, args=(world, _free_port(), nx, nyg, steps, prec, noise_a, str(tmp_path)), nprocs=world, join=True)
    ref = po.Oracle(nx, nyg, po.default_params(noise_a=noise_a), prec=prec, seed=11, reset=False)
    ref.clear()
    for (x, y) in nuclei_positions(5, nx, nyg, 3, po.philox) + [(nx // 2, 0), (0, nyg // 2)]:
        ref.add_nucleus(x, y)
    ref.step(steps)
    want = ref.fields()
    parts = [np.load(os.path.join(tmp_path, f"r{r}.npz")) for r in range(world)]
    for k in range(3):
        got = np.concatenate([p[f"arr_{k}"] for p in parts], axis=0)
        assert bit_equal(got, want[k])
    assert [n for _, n in partition(nyg, world)] == [p["arr_0"].shape[0] for p in parts]


# ---------------------------------------------------------------------------------------------- ring-wide step path
def test_next_path_mode_hysteresis():
    from crystalgrowth_b200.strips import next_path_mode
    assert next_path_mode(1, 0.01, 0.04, 0.03) == 1 and next_path_mode(1, 0.05, 0.04, 0.03) == 0
    assert next_path_mode(0, 0.035, 0.04, 0.03) == 0 and next_path_mode(0, 0.02, 0.04, 0.03) == 1


class _MockStrip:
    """Records the launch sequence a strip would run; its density probe is scripted per rank."""
    kernel = "fast"

    def __init__(self, density):
        self.density, self.log, self.steps, self.mode = density, [], 0, 1

    def step(self, n):
        self.log.append((self.mode, n))
        self.steps += n

    def sync(self):
        pass

    def path_stats(self):
        return {"dense_fraction": self.density(self.steps)}

    def set_path_mode(self, mode):
        self.mode = mode

    def ipc_export(self):
        return bytes(128)

    def ipc_link(self, lo, hi):
        pass

    def halo_refresh(self):
        pass


def _policy_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from crystalgrowth_b200.strips import StripRing
    # rank 1's strip becomes dense after 100 sub-steps and sparse again after 300; rank 0's never does
    density = (lambda s: 0.0) if rank == 0 else (lambda s: 0.5 if 100 <= s < 300 else 0.0)
    ring = StripRing(64, 64, 1e-4, rank=rank, world=world, make_strip=lambda y0, ny: _MockStrip(density))
    for n in (10, 70, 1, 200, 37, 130):
        ring.step(n)
    np.save(os.path.join(out_dir, f"p{rank}.npy"), np.array(ring.strip.log))
    dist.barrier()
    dist.destroy_process_group()


def test_ring_agrees_on_the_step_path(tmp_path):
    """Linked strips must run identical launch sequences: the ring max-reduces the density probes every 64 sub-steps and
    every rank switches between pairs (1) and the single-step kernel (0) on the same sub-step — although only ONE
    rank's strip is dense."""
    world = 2
    mp.spawn(_policy_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    logs = [np.load(os.path.join(tmp_path, f"p{r}.npy")) for r in range(world)]
    assert np.array_equal(logs[0], logs[1])
    modes = [int(m) for m, _ in logs[0]]
    assert modes[0] == 1 and 0 in modes and modes[-1] == 1            # pairs -> single (rank 1 dense) -> pairs again
    assert sum(int(n) for _, n in logs[0]) == 10 + 70 + 1 + 200 + 37 + 130
    assert all(int(n) <= 64 for _, n in logs[0])
