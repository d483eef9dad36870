"""Developer probe: the distance-triangle kernel (bf_tri_kernel.cu) against the gather kernels on the same inputs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spnngp_b200 import Engine  # noqa: E402
from tests.util import make_problem  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
ms = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [15]
eng = Engine(0)
eng.set_option("tri_force", 1)  # every m through the distance-triangle kernel, not only where it is fastest
for m in ms:
    prob = make_problem(n, m, seed=1, domain=np.sqrt(n / 5000.0))
    prob["nnIndx"], _, prob["nnIndxLU"] = eng.mk_nn_indx(prob["coords"], m, "cb")
    res = {}
    for tri in (0, 1):
        eng.set_option("tri_cache", tri)
        eng.set_data(prob["coords"], prob["nnIndx"], prob["nnIndxLU"], prob["X"], prob["y"], m)
        print("m", m, "tri", tri, "used", eng.get_option("tri_cache_used"), flush=True)
        for name, cov, theta in [("exponential", 0, [5, 1, 6, 0]), ("spherical", 1, [5, 1, 6, 0]), ("gaussian", 3, [5, 1, 6, 0]),
                                 ("matern", 2, [5, 1, 6, 0.75]), ("matern_hi", 2, [5, 1, 30, 1.5])]:
            th2 = [theta[:4], [4.0, 1.5, 7.0, theta[3]]]
            out = eng.bf_loglik(cov, th2, [prob["beta"], prob["beta"]], slots=[0, 1])
            B0, F0 = eng.get_bf(0)
            B1, F1 = eng.get_bf(1)
            res[(name, tri)] = (out.copy(), B0.copy(), F0.copy(), B1.copy(), F1.copy())
            print(f"  {name}: out={out.ravel()}", flush=True)
    for name in ("exponential", "spherical", "gaussian", "matern", "matern_hi"):
        a, b = res[(name, 0)], res[(name, 1)]
        eo = np.max(np.abs(a[0] - b[0]) / np.abs(a[0]))
        eB = max(np.max(np.abs(a[1] - b[1])), np.max(np.abs(a[3] - b[3])))
        eF = max(np.max(np.abs(a[2] - b[2]) / np.abs(a[2])), np.max(np.abs(a[4] - b[4]) / np.abs(a[4])))
        print(f"m={m} {name:12s} rel out {eo:.2e}  abs B {eB:.2e}  rel F {eF:.2e}", flush=True)
