"""dev: the two-step kernel (KOB_FAST2=1) must reproduce the single-step kernel bit for bit."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
import crystalgrowth_b200 as cg  # noqa: E402
from crystalgrowth_b200.strips import partition  # noqa: E402


def run(fast2, nx, ny, steps, nuclei, dense=False, chunks=(None,), **kw):
    os.environ["KOB_FAST2"] = str(fast2)
    g = cg.Kobayashi(nx, ny, 1e-4, kernel="fast", **kw)
    g.clear()
    for (x, y) in nuclei:
        g.add_nucleus(x, y)
    if dense:
        phi, t = bench.dense_state(nx, ny, 0)
        g.set_fields(phi, t, np.zeros((ny, nx), np.float32))
    for n in (chunks if chunks[0] is not None else (steps,)):
        g.step(n)
    out = g.fields()
    launches = g.launch_count
    g.close()
    return out, launches


def cmp(name, a, b):
    ok = True
    for nm, x, y in zip(("phi", "T", "theta"), a, b):
        same = np.array_equal(x.view(np.uint32), y.view(np.uint32))
        ok &= same
        if not same:
            d = np.abs(x.astype(np.float64) - y.astype(np.float64))
            bad = np.argwhere(x.view(np.uint32) != y.view(np.uint32))
            rows, cols = np.unique(bad[:, 0]), np.unique(bad[:, 1])
            print(f"  {name} {nm}: MISMATCH n={len(bad)} max={d.max():.3e} first={bad[:4].tolist()} rows[{rows.min()}..{rows.max()}] n={len(rows)} cols[{cols.min()}..{cols.max()}] n={len(cols)}")
    print(f"[fast2_check] {name}: {'bitwise OK' if ok else 'FAIL'}")
    return ok


cases = [
    ("64x64 single nucleus, no noise, 10 steps", dict(nx=64, ny=64, steps=10, nuclei=[(32, 32)])),
    ("700x300 5 nuclei noise 151 steps", dict(nx=700, ny=300, steps=151, nuclei=[(0, 0), (350, 150), (699, 299), (100, 40), (520, 222)], seed=11, noise_a=0.01)),
    ("150x90 noise 60 steps in chunks", dict(nx=150, ny=90, steps=60, nuclei=[(0, 0), (75, 45), (149, 89), (30, 7), (120, 8)], seed=9, noise_a=0.01, chunks=(10, 3, 7, 20, 20))),
    ("420x200 dense noise 40 steps", dict(nx=420, ny=200, steps=40, nuclei=[], dense=True, seed=21, noise_a=0.01)),
    ("420x200 dense j=4 40 steps", dict(nx=420, ny=200, steps=40, nuclei=[], dense=True, seed=21, noise_a=0.01, anisotropy=4.0)),
    ("420x200 dense j=5 theta0 30 steps", dict(nx=420, ny=200, steps=30, nuclei=[], dense=True, seed=21, noise_a=0.01, anisotropy=5.0, theta0=0.3)),
    ("420x200 dense j=5.5 30 steps", dict(nx=420, ny=200, steps=30, nuclei=[], dense=True, seed=21, noise_a=0.0, anisotropy=5.5)),
    ("250x250 reference default 400 steps", dict(nx=250, ny=250, steps=400, nuclei=[(125, 125)])),
    ("37x23 ragged tiny", dict(nx=37, ny=23, steps=30, nuclei=[(3, 3), (30, 20)], seed=2, noise_a=0.01)),
]
allok = True
for name, kw in cases:
    (a, la), (b, lb) = run(0, **kw), run(1, **kw)
    allok &= cmp(f"{name} (launches {la} vs {lb})", a, b)

# linked strips on one device
for nstrips, nyg in ((2, 64), (3, 70)):
    nx, steps = 200, 40
    nuclei = [(3, 0), (100, nyg // 2), (199, nyg - 1), (20, nyg // nstrips), (70, nyg // nstrips - 1)]
    outs = []
    for fast2 in (0, 1):
        os.environ["KOB_FAST2"] = str(fast2)
        strips = [cg.Kobayashi(nx, ny, 1e-4, kernel="fast", ny_global=nyg, y0=y0, seed=5, noise_a=0.01) for (y0, ny) in partition(nyg, nstrips)]
        for i, s in enumerate(strips):
            s.link_local(strips[(i - 1) % nstrips], strips[(i + 1) % nstrips])
            s.clear()
        for (x, y) in nuclei:
            for s in strips:
                s.add_nucleus(x, y)
        for s in strips:
            s.sync()
        for s in strips:
            s.halo_refresh()
        for s in strips:
            s.sync()
        for _ in range(steps // 2):
            for s in strips:
                s.step(2)
        outs.append([np.concatenate(parts, axis=0) for parts in zip(*[s.fields() for s in strips])])
        for s in strips:
            s.close()
    allok &= cmp(f"{nstrips} linked strips {nx}x{nyg}", outs[0], outs[1])
print("[fast2_check] ALL OK" if allok else "[fast2_check] FAILURES")
sys.exit(0 if allok else 1)
