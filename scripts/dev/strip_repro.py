"""dev: the nuclei of strip r of the N-GPU weak-scaling bench in ONE unlinked 16384^2 context: is the far pass slow because of the data?"""
import os, sys, csv
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import bench
import crystalgrowth_b200 as cg
from crystalgrowth_b200.strips import nuclei_positions

world, r = int(sys.argv[1]), int(sys.argv[2])
n = 16384
pos = [(x, y - r * n) for (x, y) in nuclei_positions(64 * world, n, n * world, bench.SEED) if r * n <= y < (r + 1) * n]
os.environ["KOB_FAST2"] = "1"
os.environ["KOB_FAST2_CONC"] = "0"
os.environ["KOB_TRACE"] = f"/tmp/strip_{r}.csv"
g = cg.Kobayashi(n, n, 1e-4, kernel="fast", seed=bench.SEED, noise_a=0.01)
g.clear()
for (x, y) in pos:
    g.add_nucleus(x, y)
g.step(50); g.sync()
ms = g.step_timed(200)
phi, t, th = g.fields()
nzb = (th.reshape(n // 32, 32, n // 128, 128) != 0).any(axis=(1, 3)).mean()
g.close()
rows = list(csv.DictReader(open(f"/tmp/strip_{r}.csv")))
far = [float(x["duration_us"]) for x in rows if x["kernel"] == "kob_far2"][-100:]
gen = [float(x["duration_us"]) for x in rows if x["kernel"].startswith("kob_step_fast2")][-100:]
print(f"strip {r}/{world}: {len(pos)} nuclei, pair {ms / 100 * 1e3:.1f} us, far2 mean {np.mean(far):.1f} (first {far[0]:.0f}, last {far[-1]:.0f}), general mean {np.mean(gen):.1f}, theta blocks non-zero {nzb:.4f}, solid {float((phi > 0.5).mean()):.5f}")
