"""Developer probe (not the contract bench): time the fused B/F + log-likelihood launch at size n."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spnngp_b200 import Engine  # noqa: E402
from tests.util import make_problem  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 15
t = time.time()
prob = make_problem(n, m, seed=1, domain=np.sqrt(n / 5000.0))
eng = Engine(0)
eng.set_option("tri_cache", int(os.environ.get("NNGP_TRI", "1")))
eng.set_option("tri_force", int(os.environ.get("NNGP_TRI_FORCE", "0")))
prob["nnIndx"], _, prob["nnIndxLU"] = eng.mk_nn_indx(prob["coords"], m, "cb")
t = time.time()
eng.set_data(prob["coords"], prob["nnIndx"], prob["nnIndxLU"], prob["X"], prob["y"], m)
print(f"set_data {time.time()-t:.3f}s", flush=True)
res = {}
for name, cov, theta in [("exponential", 0, [5, 1, 6]), ("gaussian", 3, [5, 1, 6]), ("spherical", 1, [5, 1, 6]),
                         ("matern", 2, [5, 1, 6, 0.75])]:
    for nb in (1, 2):
        th = [theta] * nb
        be = [prob["beta"]] * nb
        eng.time_bf_loglik(cov, th, be, iters=3)
        ms = eng.time_bf_loglik(cov, th, be, iters=20)
        res[f"{name}_b{nb}"] = dict(ms=ms, loc_per_s=n * nb / ms * 1e3)
        print(f"{name:12s} batch={nb}: {ms:.3f} ms/launch  {n*nb/ms*1e3:.3e} loc/s", flush=True)
out = eng.bf_loglik(0, [[5, 1, 6]], [prob["beta"]], slots=[0])
t = time.time()
for _ in range(20):
    eng.q_resid(0, prob["beta"])
print(f"q_resid host-timed {(time.time()-t)/20*1e3:.3f} ms")
t = time.time()
for _ in range(20):
    eng.q_forms(0)
print(f"q_forms host-timed {(time.time()-t)/20*1e3:.3f} ms")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"quick_bench_n{n}_m{m}.json"), "w"), indent=1)
