"""Developer probe: time one replicated-data draw (nngp_rep_draw) at size n, and the reference's serial loop."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spnngp_b200 import Engine
from bench import make_workload
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
m = 15
wl = make_workload(n, m)
eng = Engine(0)
idx, dist, lu = eng.mk_nn_indx(wl["coords"], m, "cb")
eng.set_data(wl["coords"], idx, lu, wl["X"], wl["y"], m)
eng.bf_loglik(0, [[5, 1, 6]], [wl["beta"]], slots=[0])
z = np.random.default_rng(1).standard_normal(n)
for k in range(4):
    t = time.perf_counter(); y = eng.rep_draw(0, wl["beta"], z); print(f"rep_draw {k}: {(time.perf_counter()-t)*1e3:.2f} ms", flush=True)
B, F = eng.get_bf(0)
R = __import__("spnngp_b200")._lib.load_rshim()
t = time.perf_counter()
zz = np.array([R.rnorm(0.0, 1.0) for _ in range(100000)])
print(f"R-compatible rnorm through ctypes: {(time.perf_counter()-t)/100000*1e9:.0f} ns each (python call overhead included)")
