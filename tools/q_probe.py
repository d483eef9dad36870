"""Developer probe for ncu: Q-only passes at n = 10^6, m = 15."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spnngp_b200 import Engine
from bench import make_workload
n, m = 1000000, 15
wl = make_workload(n, m)
eng = Engine(0)
idx, dist, lu = eng.mk_nn_indx(wl["coords"], m, "cb")
eng.set_data(wl["coords"], idx, lu, wl["X"], wl["y"], m)
eng.bf_loglik(0, [[5, 1, 6]], [wl["beta"]], slots=[0])
for _ in range(3):
    eng.q_resid(0, wl["beta"]); eng.q_forms(0)
