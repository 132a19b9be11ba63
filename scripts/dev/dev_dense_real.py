"""dev: (1) `make PATH`: grow a real developed-dendrite field and save it; (2) `run PATH [n]`: load it and time / step it
(for ncu captures of the single-step kernel in the dense regime)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import crystalgrowth_b200 as cg  # noqa: E402
from crystalgrowth_b200.strips import nuclei_positions  # noqa: E402

N, SEED = 4096, 20260101
mode, path = sys.argv[1], sys.argv[2]
g = cg.Kobayashi(N, N, 1e-4, kernel="fast", seed=SEED, noise_a=0.01)
if mode == "make":
    g.clear()
    for (x, y) in nuclei_positions(16, N, N, SEED):
        g.add_nucleus(x, y)
    g.step(int(sys.argv[3]) if len(sys.argv) > 3 else 16000)
    g.save_checkpoint(path)
    print("saved", g.step_counter)
else:
    os.environ.setdefault("KOB_FAST2", "0")
    g.load_checkpoint(path)
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    g.step(20)
    ms = g.step_timed(n)
    print(f"dense-real: {N * N * n / (ms * 1e-3) / 1e9:.1f} Gcell/s, {ms / n:.4f} ms/launch")
