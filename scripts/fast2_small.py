import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["KOB_FAST2"] = "1"
import crystalgrowth_b200 as cg
g = cg.Kobayashi(64, 64, 1e-4, kernel="fast")
g.step(2)
print("ok", float(g.phi().sum()))
