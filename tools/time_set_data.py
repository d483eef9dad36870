import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spnngp_b200 import Engine
from bench import make_workload
n, m = 1000000, 15
wl = make_workload(n, m)
eng = Engine(0)
t = time.perf_counter(); idx, dist, lu = eng.mk_nn_indx(wl["coords"], m, "cb"); print("knn cb 1st", time.perf_counter() - t)
t = time.perf_counter(); idx, dist, lu = eng.mk_nn_indx(wl["coords"], m, "cb"); print("knn cb 2nd", time.perf_counter() - t)
for k in range(4):
    t = time.perf_counter(); eng.set_data(wl["coords"], idx, lu, wl["X"], wl["y"], m); print("set_data", k, time.perf_counter() - t)
import ctypes as C
from spnngp_b200.engine import colmajor, _i32, _f64
t = time.perf_counter(); c = colmajor(wl["coords"]); X = colmajor(wl["X"]); y = _f64(wl["y"]); i2 = _i32(idx); l2 = _i32(lu); print("python marshalling", time.perf_counter() - t)
