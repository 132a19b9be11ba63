import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import crystalgrowth_b200 as cg
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
g = cg.Kobayashi(n, n, 1e-4, kernel="fast", noise_a=0.01, seed=3)
g.step(int(sys.argv[2]) if len(sys.argv) > 2 else 3)
g.sync()
print("ok", float(g.phi().sum()))
