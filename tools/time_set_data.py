"""Developer probe: nngp_set_data wall time (host arrays in) at n = 10^6, m = 15; NNGP_TRACE=1 prints its phases."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spnngp_b200 import Engine
from bench import make_workload
wl = make_workload(1000000, 15)
eng = Engine(0)
idx, dist, lu = eng.mk_nn_indx(wl["coords"], 15, "cb")
for k in range(4):
    t = time.perf_counter(); eng.set_data(wl["coords"], idx, lu, wl["X"], wl["y"], 15); print(f"set_data {k}: {(time.perf_counter()-t)*1e3:.2f} ms", flush=True)
