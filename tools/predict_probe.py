"""Developer probe: time of the prediction solves per distinct theta (few samples, so the q x S transfers are small)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spnngp_b200 import Engine  # noqa: E402

n, m, q = 1000000, 15, int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
rng = np.random.default_rng(11)
W = np.sqrt(n / 5000.0)
c = rng.random((n, 2)) * W
c = c[np.argsort(c[:, 0], kind="stable")]
X = np.column_stack([np.ones(n), rng.standard_normal(n)])
y = X @ np.array([1.0, 5.0]) + rng.standard_normal(n)
eng = Engine(0)
idx, _, lu = eng.mk_nn_indx(c, m, "cb")
eng.set_data(c, idx, lu, X, y, m)
c0 = rng.random((q, 2)) * W
X0 = np.column_stack([np.ones(q), rng.standard_normal(q)])
nn0 = eng.nn_query(c, c0, m)
for cov, name in ((0, "exponential"), (2, "matern")):
    for S in (1, 9):
        bs = np.array([[1.0], [5.0]]) + 0.05 * rng.standard_normal((2, S))
        th = np.array([[5.0], [1.0], [6.0], [0.75]]) * (1 + 0.02 * rng.standard_normal((4, S)))
        z = rng.standard_normal(q * S)
        eng.rnngp_predict(cov, X0, c0, nn0, bs, th[: (4 if cov == 2 else 3)], z)
        t = time.perf_counter()
        eng.rnngp_predict(cov, X0, c0, nn0, bs, th[: (4 if cov == 2 else 3)], z)
        dt = time.perf_counter() - t
        print(f"{name:12s} S={S} distinct theta: {dt*1e3:8.2f} ms", flush=True)
