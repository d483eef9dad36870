import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spnngp_b200 import Engine, _lib
from bench import make_workload, sampler_settings
n, m = 1000000, 15
cov = int(sys.argv[1]) if len(sys.argv) > 1 else 0
wl = make_workload(n, m)
eng = Engine(0)
idx, dist, lu = eng.mk_nn_indx(wl["coords"], m, "cb")
starting, tuning, pri = sampler_settings(cov)
R = _lib.load_rshim()
eng.set_data(wl["coords"], idx, lu, wl["X"], wl["y"], m)
def chain(k, mode=0):
    R.rshim_set_seed(1)
    return eng.rnngp(cov, pri["sigmaSqIG"], pri["tauSqIG"], pri["phiUnif"], pri["nuUnif"], starting, tuning, k, R.rnorm, R.runif, mode=mode)
chain(3)
for k in (1, 5, 20, 20, 100):
    t = time.perf_counter(); chain(k); dt = time.perf_counter() - t
    print(f"chain({k}) {dt*1e3:.1f} ms  -> {dt/k*1e3:.2f} ms/iter", flush=True)
for rep in range(3):
    t = time.perf_counter(); eng.set_data(wl["coords"], idx, lu, wl["X"], wl["y"], m); t1 = time.perf_counter(); chain(20); t2 = time.perf_counter()
    print(f"e2e: set_data {1e3*(t1-t):.1f} ms, chain(20) {1e3*(t2-t1):.1f} ms")
