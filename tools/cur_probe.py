"""Developer probe: the sampler's fused launch (nngp_bf_loglik_cur: candidate solve + Q of the resident factors), wall time."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spnngp_b200 import Engine
from tests.util import make_problem
n, m = 1000000, 15
prob = make_problem(n, m, seed=1, domain=np.sqrt(n / 5000.0))
eng = Engine(0)
prob["nnIndx"], _, prob["nnIndxLU"] = eng.mk_nn_indx(prob["coords"], m, "cb")
eng.set_data(prob["coords"], prob["nnIndx"], prob["nnIndxLU"], prob["X"], prob["y"], m)
for name, cov, th in (("exponential", 0, [5, 1, 6]), ("matern", 2, [5, 1, 6, 0.75])):
    eng.bf_loglik(cov, [th], [prob["beta"]], slots=[0])
    cand = list(th); cand[2] *= 1.001
    for _ in range(5):
        eng.bf_loglik_cur(cov, cand, prob["beta"], 1, 0)
    t = time.perf_counter()
    for _ in range(40):
        out = eng.bf_loglik_cur(cov, cand, prob["beta"], 1, 0)
    print(f"{name:12s} fused launch (host wall, incl. residual kernel): {(time.perf_counter()-t)/40*1e3:.4f} ms  Qcur {out[2]:.6f}", flush=True)
