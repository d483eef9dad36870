"""Short single-GPU workload for ncu captures of every kernel of the path EXCEPT the B/F solve (which tools/ncu_probe.py
covers): ordered search (code-book, its head kernel, brute force), query search, Q passes, replicate recurrence,
prediction emit, distance-triangle build.  Sizes chosen so that one `ncu --set full` pass stays under a minute."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spnngp_b200 import Engine  # noqa: E402
from tests.util import make_problem  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
m = 15
rng = np.random.default_rng(3)
prob = make_problem(n, m, seed=1, domain=np.sqrt(n / 5000.0))
eng = Engine(0)
idx, _, lu = eng.mk_nn_indx(prob["coords"], m, "cb")                    # knn_cb_kernel, knn_cb_head_kernel
nb = min(n, 100000)
eng.mk_nn_indx(prob["coords"][:nb], m, "brute")                         # knn_brute_kernel
q = 200000
c0 = rng.random((q, 2)) * np.sqrt(n / 5000.0)
nn0 = eng.nn_query(prob["coords"], c0, m)                               # knn_grid_query_kernel
eng.set_data(prob["coords"], idx, lu, prob["X"], prob["y"], m)          # tri_build_kernel, xy_rows_kernel
eng.bf_loglik(0, [[5.0, 1.0, 6.0]], [prob["beta"]], slots=[0])
eng.q_resid(0, prob["beta"])                                            # resid, q_resid_kernel
eng.q_forms(0)                                                          # q_forms_kernel
eng.rep_draw(0, prob["beta"], rng.standard_normal(n))                   # rep_solve_kernel
S, qs = 100, 50000
X0 = np.column_stack([np.ones(qs), rng.standard_normal(qs)])
bs = prob["beta"][:, None] + 0.01 * rng.standard_normal((2, S))
th = np.tile(np.array([[5.0], [1.0], [6.0]]), (1, S))
eng.rnngp_predict(0, X0, c0[:qs], nn0.reshape(m, q)[:, :qs].ravel(), bs, th, rng.standard_normal(qs * S))  # site solve, emit
print("done")
