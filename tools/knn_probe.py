"""Developer probe for ncu: one code-book search at size n, m."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spnngp_b200 import Engine
n = int(sys.argv[1]); m = int(sys.argv[2]); kind = sys.argv[3] if len(sys.argv) > 3 else "cb"
rng = np.random.default_rng(7)
c = rng.random((n, 2)) * np.sqrt(n / 5000.0)
c = c[np.argsort(c[:, 0], kind="stable")]
eng = Engine(0)
for _ in range(2):
    t = time.perf_counter(); eng.mk_nn_indx(c, m, kind); print(kind, n, m, time.perf_counter() - t, flush=True)
