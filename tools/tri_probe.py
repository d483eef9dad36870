"""Developer probe: B/F launch time with and without the quadratic form (no residual gathers), per model."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spnngp_b200 import Engine  # noqa: E402
from tests.util import make_problem  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 15
prob = make_problem(n, m, seed=1, domain=np.sqrt(n / 5000.0))
eng = Engine(0)
eng.set_option("tri_cache", int(os.environ.get("NNGP_TRI", "1")))
eng.set_option("tri_force", int(os.environ.get("NNGP_TRI_FORCE", "0")))
prob["nnIndx"], _, prob["nnIndxLU"] = eng.mk_nn_indx(prob["coords"], m, "cb")
eng.set_data(prob["coords"], prob["nnIndx"], prob["nnIndxLU"], prob["X"], prob["y"], m)
for name, cov, theta in [("exponential", 0, [5, 1, 6]), ("spherical", 1, [5, 1, 6]), ("matern", 2, [5, 1, 6, 0.75])]:
    for label, be in (("with Q", [prob["beta"]]), ("no Q", None)):
        eng.time_bf_loglik(cov, [theta], be, iters=3)
        ms = eng.time_bf_loglik(cov, [theta], be, iters=20)
        print(f"{name:12s} {label:7s}: {ms:.3f} ms/launch", flush=True)
