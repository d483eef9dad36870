"""Developer probe: B/F launch time of the head shard [0, X) vs an inner shard of the same size, Matern, per series level
(matern_short = 2 / 0): what the slow correlation paths of the first locations cost."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spnngp_b200 import Engine
from bench import make_workload
n, m = 1000000, 15
wl = make_workload(n, m)
eng = Engine(0)
idx, _, lu = eng.mk_nn_indx(wl["coords"], m, "cb")
th = [5, 1, 6, 0.75]
for X in (4000, 32000, 125000):
    for level in (2, 0):
        eng.set_option("matern_short", level)
        row = []
        for lo in (0, 500000):
            eng.set_data(wl["coords"], idx, lu, wl["X"], wl["y"], m, loc_range=(lo, lo + X))
            eng.time_bf_loglik(2, [th], [wl["beta"]], iters=3)
            row.append(min(eng.time_bf_loglik(2, [th], [wl["beta"]], iters=30) for _ in range(2)))
        print(f"X={X:7d} matern_short={level}: head {row[0]*1e3:7.1f} us   inner {row[1]*1e3:7.1f} us   head - inner {1e3*(row[0]-row[1]):6.1f} us", flush=True)
