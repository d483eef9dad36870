"""Developer probe: wall time of repeated nngp_nn_query calls (grid search; NNGP_NN_QUERY=strip for the strip walk)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spnngp_b200 import Engine
eng = Engine(0)
rng = np.random.default_rng(5)
for n, q in ((151000, 37700), (1000000, 250000), (1000000, 1000000)):
    W = np.sqrt(n / 5000.0)
    c = rng.random((n, 2)) * W
    c0 = rng.random((q, 2)) * W
    ts = []
    for r in range(5):
        t = time.perf_counter(); out = eng.nn_query(c, c0, 15); ts.append(time.perf_counter() - t)
    print(f"n={n} q={q}: " + " ".join(f"{x*1e3:.1f}" for x in ts) + " ms", flush=True)
