import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import crystalgrowth_b200 as cg
from oracle import pyoracle as po
o = po.Oracle(32, 32, math=po.MATH_LIBM)
g = cg.Kobayashi(32, 32, 1e-4, kernel="fast")
for s in range(1, 4):
    prev = o.fields()
    o.step(1); g.step(1)
    fo = o.fields(); fg = g.fields()
    d = np.abs(fg[2].astype(np.float64) - fo[2].astype(np.float64))
    bad = np.argwhere(d > 1e-5)
    print("step", s, "bad cells", len(bad), "phi maxdiff", np.abs(fg[0].astype(float)-fo[0].astype(float)).max())
    for (j, i) in bad[:12]:
        p = prev[0]
        gx = (p[j, (i+1) % 32] - p[j, (i-1) % 32]) / np.float32(0.03)
        gy = (p[(j+1) % 32, i] - p[(j-1) % 32, i]) / np.float32(0.03)
        print("  cell i,j", i, j, "theta gpu", fg[2][j, i], "oracle", fo[2][j, i], "prev theta", prev[2][j, i], "gx", gx, "gy", gy)
