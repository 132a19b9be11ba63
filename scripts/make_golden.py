"""Generate tests/golden/*.npz by RUNNING THE REFERENCE ITSELF: /root/reference/src/Kobayashi.cpp compiled
unmodified into oracle/_ref/ (oracle/Makefile, g++ -O2 -ffp-contract=off, glibc libm).  Only works in the
authoring container, where /root/reference exists; the vectors it writes are committed and travel.

    python scripts/make_golden.py

Each file holds the reference's _phi, _t, _angl (shape (ny, nx) = index i + nx*j, src/Kobayashi.h:91) at the
listed sub-step counts, plus the parameters used.  FP64 files come from the FP64-typed build of the same
text (oracle/ref_harness/ref_driver.cpp, -DKOB_REF_FP64).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def run(name, nx, ny, steps, prec=32, nuclei=None, **params):
    # src/main.cpp:16 passes the float literal 0.0001f; the FP64-typed build widens that float
    r = po.Reference(nx, ny, float(np.float32(1e-4)) if prec == 64 else 1e-4, prec=prec, **params)
    if nuclei is not None:
        z = np.zeros((ny, nx), r.dtype)
        r.set_fields(z, z, z)
        for (x, y) in nuclei:
            r.add_nucleus(x, y)
    out = {"nx": nx, "ny": ny, "prec": prec, "steps": np.array(steps),
           "params": np.array(sorted(params.items()), dtype=object) if params else np.array([], dtype=object),
           "nuclei": np.array(nuclei if nuclei is not None else [(nx // 2, ny // 2)])}
    done = 0
    for s in steps:
        r.step(s - done)
        done = s
        phi, t, a = r.fields()
        out[f"phi_{s}"], out[f"t_{s}"], out[f"angl_{s}"] = phi, t, a
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, {s: float(out[f'phi_{s}'].astype(np.float64).sum()) for s in steps})


def run_colors(name, n):
    """The viewer's per-object colours (iUpdateConstantBuffer, src/Kobayashi.cpp:309-345) for a designed phi field:
    random values over [-0.05, 1.05] plus the ramp's break points, on a square grid (the reference derives the
    object -> cell mapping from sqrt(#objects), :311-313)."""
    r = po.Reference(n, n, 1e-4, prec=32)
    rng = np.random.default_rng(20260101)
    phi = (rng.random((n, n)) * 1.1 - 0.05).astype(np.float32)
    phi[0, :8] = [0.0, 0.9, 0.99, 1.0, np.nextafter(np.float32(0.9), np.float32(1)), np.nextafter(np.float32(0.99), np.float32(1)), 0.45, 0.995]
    z = np.zeros((n, n), np.float32)
    r.set_fields(phi, z, z)
    rgb = r.colors()                                    # object i -> colour, i = 0 .. n*n-1
    np.savez_compressed(os.path.join(OUT, name + ".npz"), n=n, phi=phi, rgb=rgb)
    print(name, rgb.shape, float(rgb.astype(np.float64).sum()))


if __name__ == "__main__":
    assert po.ref_available(32) and po.ref_available(64), "needs /root/reference (authoring container)"
    os.makedirs(OUT, exist_ok=True)
    run("ref_f32_n32_j6", 32, 32, [1, 2, 3])
    run("ref_f32_n64_j6", 64, 64, [4, 6, 10, 100])
    run("ref_f32_n64_j4", 64, 64, [4, 6, 10, 100], anisotropy=4.0)
    run("ref_f64_n64_j6", 64, 64, [10, 100, 300], prec=64)
    run("ref_f64_n64_j4", 64, 64, [10, 100, 300], prec=64, anisotropy=4.0)
    # ragged grid, several nuclei touching the periodic seams (kept >= 1 cell inside: the reference's
    # _createNucleus does not wrap, src/Kobayashi.cpp:116-123)
    run("ref_f32_96x40_multi", 96, 40, [5, 60], nuclei=[(1, 1), (94, 38), (50, 1), (1, 20)])
    run("ref_f64_37x53_j5", 37, 53, [5, 120], prec=64, nuclei=[(1, 1), (35, 51), (18, 26)], anisotropy=5.0,
        K=1.2, tau=0.0004, delta=0.03, alpha=1.1, gamma=15.0, t_eq=0.9, epsilon_bar=0.012)
    # warm checkpoints for the tolerance windows of SURVEY §4 (G1/G2): state at step 500 and +1/+100/+200
    run("ref_f32_n128_j6_warm", 128, 128, [500, 501, 600, 700])
    run("ref_f64_n128_j6_warm", 128, 128, [500, 501, 1000], prec=64)
    run_colors("ref_colors_n48", 48)
