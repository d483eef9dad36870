"""Short single-GPU workload for ncu captures: a few fused B/F + log-likelihood launches."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spnngp_b200 import Engine  # noqa: E402
from tests.util import make_problem  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 15
cov = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
prob = make_problem(n, m, seed=1, domain=np.sqrt(n / 5000.0))
eng = Engine(0)
prob["nnIndx"], _, prob["nnIndxLU"] = eng.mk_nn_indx(prob["coords"], m, "cb")
eng.set_data(prob["coords"], prob["nnIndx"], prob["nnIndxLU"], prob["X"], prob["y"], m)
theta = [5, 1, 6, 0.75][: 4 if cov == 2 else 3]
for _ in range(reps):
    out = eng.bf_loglik(cov, [theta], [prob["beta"]], slots=[0])
print(out)
