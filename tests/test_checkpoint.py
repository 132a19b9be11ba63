"""KOBCKPT1 checkpoint files (crystalgrowth_b200/checkpoint.py, Kobayashi::saveCheckpoint): format round trips on
the CPU; on the GPU a resumed run continues bit-identically, and files written by the C++ driver read back in Python."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, bit_equal

DRIVER = os.path.join(ROOT, "crystalgrowth_b200", "driver", "kob_bench")


def _header(ck, **over):
    kw = dict(elem_bytes=4, nx=37, ny=11, ny_global=53, y0=21, step_counter=2**40 + 7, seed=2**63 + 5, sim_frame=12,
              params=tuple(0.5 + i for i in range(14)))
    kw.update(over)
    return ck.CheckpointHeader(**kw)


def test_header_round_trip_and_layout():
    from crystalgrowth_b200 import checkpoint as ck
    h = _header(ck)
    b = ck.pack_header(h)
    assert len(b) == 256 and b[:8] == b"KOBCKPT1" and ck.unpack_header(b) == h
    assert int.from_bytes(b[16:24], "little") == 37 and int.from_bytes(b[48:56], "little") == 2**40 + 7
    assert np.frombuffer(b[72:72 + 112], "<f8").tolist() == list(h.params)
    with pytest.raises(ValueError):
        ck.unpack_header(b"NOTACKPT" + b[8:])
    with pytest.raises(ValueError):
        ck.unpack_header(b[:100])


@pytest.mark.parametrize("elem", [4, 8])
def test_file_round_trip(tmp_path, elem):
    from crystalgrowth_b200 import checkpoint as ck
    h = _header(ck, elem_bytes=elem)
    rng = np.random.default_rng(elem)
    arrs = [rng.standard_normal((h.ny, h.nx)).astype(h.dtype) for _ in range(3)]
    p = str(tmp_path / "a.kobck")
    ck.write_checkpoint(p, h, *arrs)
    assert os.path.getsize(p) == 256 + 3 * h.nx * h.ny * elem
    h2, *back = ck.read_checkpoint(p)
    assert h2 == h and all(bit_equal(a, b) for a, b in zip(arrs, back))
    with pytest.raises(ValueError):
        ck.write_checkpoint(p, h, arrs[0][:-1], arrs[1], arrs[2])
    with open(p, "r+b") as f:
        f.truncate(300)
    with pytest.raises(ValueError):
        ck.read_checkpoint(p)


@pytest.mark.gpu
@pytest.mark.parametrize("kernel,prec", [("fast", "f32"), ("strict", "f64")])
def test_resume_is_bit_identical(cg, tmp_path, kernel, prec):
    kw = dict(precision=prec, kernel=kernel, seed=77, noise_a=0.01, anisotropy=4.0)
    a = cg.Kobayashi(150, 90, 1e-4, **kw)
    a.step(25)
    path = str(tmp_path / "mid.kobck")
    a.save_checkpoint(path)
    a.step(25)
    b = cg.Kobayashi(150, 90, 1e-4, precision=prec, kernel=kernel, seed=77)      # default parameters: the file restores them
    b.load_checkpoint(path)
    assert b.step_counter == 25 and b.anisotropy == 4.0 and b.noise_a == 0.01
    b.step(25)
    assert all(bit_equal(x, y) for x, y in zip(a.fields(), b.fields()))
    c = cg.Kobayashi(150, 90, 1e-4, precision=prec, kernel=kernel, seed=78)
    with pytest.raises(ValueError):
        c.load_checkpoint(path)                                                # other Philox seed


@pytest.mark.gpu
def test_driver_checkpoints_read_back_in_python(cg, tmp_path):
    """kob_bench --save / --resume (C++ writer) against the Python mirror: same bytes on disk, same continuation."""
    from crystalgrowth_b200 import checkpoint as ck
    f1, f2 = str(tmp_path / "a.kobck"), str(tmp_path / "b.kobck")
    base = [DRIVER, "--nx", "120", "--ny", "80", "--noise", "0.01", "--seed", "5", "--kernel", "fast"]
    r = subprocess.run(base + ["--frames", "4", "--save", f1], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert r.returncode == 0, r.stdout
    g = cg.Kobayashi(120, 80, 1e-4, kernel="fast", seed=5, noise_a=0.01)
    g.step(50)                                     # warm-up frame + 4 frames of 10 sub-steps
    h, phi, t, th = ck.read_checkpoint(f1)
    assert (h.nx, h.ny, h.step_counter, h.seed, h.elem_bytes) == (120, 80, 50, 5, 4)
    assert all(bit_equal(x, y) for x, y in zip(g.fields(), (phi, t, th)))
    r = subprocess.run(base + ["--frames", "2", "--resume", f1, "--save", f2], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert r.returncode == 0, r.stdout
    assert json.loads(r.stdout.strip().splitlines()[-1])["frames"] == 2
    g.step(20)
    h2, phi2, t2, th2 = ck.read_checkpoint(f2)
    assert h2.step_counter == 70 and all(bit_equal(x, y) for x, y in zip(g.fields(), (phi2, t2, th2)))
