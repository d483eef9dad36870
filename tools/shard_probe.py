"""Developer probe: B/F launch time of each of N location shards on one GPU (is the head of the ordering slower?)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spnngp_b200 import Engine
from spnngp_b200.dist import shard_range
from bench import make_workload
n, m, N = 1000000, 15, int(sys.argv[1]) if len(sys.argv) > 1 else 8
wl = make_workload(n, m)
eng = Engine(0)
idx, _, lu = eng.mk_nn_indx(wl["coords"], m, "cb")
for cov, th in ((2, [5, 1, 6, 0.75]), (0, [5, 1, 6])):
    out = []
    for r in range(N):
        lo, hi = shard_range(n, m, r, N)
        eng.set_data(wl["coords"], idx, lu, wl["X"], wl["y"], m, loc_range=(lo, hi))
        eng.time_bf_loglik(cov, [th], [wl["beta"]], iters=3)
        out.append(eng.time_bf_loglik(cov, [th], [wl["beta"]], iters=30))
    print("cov", cov, "ms per shard:", " ".join(f"{x:.4f}" for x in out), flush=True)
