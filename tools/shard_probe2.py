"""Developer probe: B/F launch time of the first shard [0, X) vs an inner shard [500000, 500000 + X), Matern."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spnngp_b200 import Engine
from bench import make_workload
n, m = 1000000, 15
wl = make_workload(n, m)
eng = Engine(0)
idx, _, lu = eng.mk_nn_indx(wl["coords"], m, "cb")
th = [5, 1, 6, 0.75]
for X in (5000, 20000, 60000, 100000, 125000, 250000):
    row = []
    for lo in (0, 500000):
        eng.set_data(wl["coords"], idx, lu, wl["X"], wl["y"], m, loc_range=(lo, lo + X))
        eng.time_bf_loglik(2, [th], [wl["beta"]], iters=3)
        row.append(eng.time_bf_loglik(2, [th], [wl["beta"]], iters=30))
    print(f"X={X:7d}: head shard {row[0]:.4f} ms   inner shard {row[1]:.4f} ms", flush=True)
