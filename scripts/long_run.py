"""dev: a long dendrite-growth run (BASELINE configs[4] in small) that shows the adaptive step path at work: two-step launch
pairs while the field is sparse, the single-step kernel once the crystals fill the grid.  Prints a markdown table."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crystalgrowth_b200 as cg  # noqa: E402
from crystalgrowth_b200.strips import nuclei_positions  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--nuclei", type=int, default=16)
ap.add_argument("--steps", type=int, default=40000)
ap.add_argument("--chunk", type=int, default=2000)
a = ap.parse_args()
g = cg.Kobayashi(a.n, a.n, 1e-4, kernel="fast", seed=20260101, noise_a=0.01)
g.clear()
for (x, y) in nuclei_positions(a.nuclei, a.n, a.n, 20260101):
    g.add_nucleus(x, y)
g.step(10)
g.sync()
print(f"# {a.n}^2, {a.nuclei} nuclei, Philox noise a = 0.01, j = 6: {a.steps} sub-steps in chunks of {a.chunk} (KOB_FAST2 = {os.environ.get('KOB_FAST2', '2 (adaptive)')})")
print("| sub-steps done | Gcell-updates/s (chunk) | paired / single sub-steps in chunk | density probe | solid fraction (phi > 0.5) |")
print("|---|---|---|---|---|")
done, t_all = 10, 0.0
while done < a.steps:
    s0 = g.path_stats()
    ms = g.step_timed(a.chunk)
    s1 = g.path_stats()
    done += a.chunk
    t_all += ms
    phi = g.phi()
    print(f"| {done} | {a.n * a.n * a.chunk / (ms * 1e-3) / 1e9:.1f} | {s1['paired_steps'] - s0['paired_steps']} / {s1['single_steps'] - s0['single_steps']} | "
          f"{s1['dense_fraction']:.3f} | {float((phi > 0.5).mean()):.4f} |", flush=True)
print(f"\ntotal: {a.n * a.n * (done - 10) / (t_all * 1e-3) / 1e9:.1f} Gcell-updates/s over {done - 10} sub-steps ({t_all / 1e3:.2f} s of device time)")
