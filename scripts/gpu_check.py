"""Dev script (GPU box): strict/fast kernels vs the CPU oracle; prints max-abs differences."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import crystalgrowth_b200 as cg
from oracle import pyoracle as po

def diff(a, b):
    return [float(np.abs(x.astype(np.float64) - y.astype(np.float64)).max()) for x, y in zip(a, b)]

def bit_eq(a, b):
    return all(np.array_equal(x.view(np.uint8), y.view(np.uint8)) for x, y in zip(a, b))

def case(nx, ny, steps, prec, kernel, j=6.0, a=0.0, seed=7, nuclei=None):
    p = po.default_params(anisotropy=j, noise_a=a)
    o = po.Oracle(nx, ny, p, prec=64 if prec == "f64" else 32, math=po.MATH_PORTABLE, seed=seed, threads=8)
    g = cg.Kobayashi(nx, ny, 1e-4, precision=prec, kernel=kernel, seed=seed, anisotropy=j, noise_a=a)
    if nuclei:
        o.clear(); g.clear()
        for (x, y) in nuclei:
            o.add_nucleus(x, y); g.add_nucleus(x, y)
    t0 = time.time(); o.step(steps); t1 = time.time()
    g.step(steps); fg = g.fields(); fo = o.fields()
    print(f"{nx}x{ny} steps={steps} {prec} {kernel} j={j} a={a}: bitwise={bit_eq(fg, fo)} maxdiff={diff(fg, fo)} (oracle {t1-t0:.2f}s)", flush=True)

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "strict"
    if which in ("strict", "all"):
        case(64, 64, 1, "f32", "strict")
        case(64, 64, 50, "f32", "strict")
        case(250, 250, 300, "f32", "strict")
        case(250, 250, 300, "f32", "strict", j=4.0)
        case(37, 53, 200, "f32", "strict", j=5.0)
        case(96, 40, 120, "f32", "strict", a=0.01, nuclei=[(1, 1), (95, 39), (50, 0), (0, 20)])
        case(64, 64, 100, "f64", "strict")
        case(130, 70, 150, "f64", "strict", a=0.02, nuclei=[(0, 0), (64, 35), (129, 69)])
        case(1024, 1024, 20, "f32", "strict")
    if which in ("fast", "all"):
        for st in (1, 3, 5):
            case(250, 250, st, "f32", "fast")
        case(96, 40, 5, "f32", "fast", a=0.01, nuclei=[(1, 1), (95, 39), (50, 0), (0, 20)])
