"""Timings of the other rows of SURVEY 8(d) on one GPU (not the contract bench): neighbour searches (C5),
prediction (C4-like), the conjugate model (C2-like), Q-only passes.  Writes profiles/r01_extra_timings.json.
Host wall clock around the public calls (host buffers in, results out), best of `reps`."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spnngp_b200 import Engine  # noqa: E402

out = {}
eng = Engine(0)
rng = np.random.default_rng(7)


def best(f, reps=2):
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); r = f(); ts.append(time.perf_counter() - t)
    return min(ts), r


def sorted_coords(n, scale=None):
    W = np.sqrt(n / 5000.0) if scale is None else scale
    c = rng.random((n, 2)) * W
    return c[np.argsort(c[:, 0], kind="stable")]


# ---- C5: ordered neighbour search -----------------------------------------------------------------------
knn = {}
for n, m, kinds in ((100000, 15, ("brute", "cb")), (100000, 32, ("brute", "cb")), (1000000, 15, ("cb",)),
                    (1000000, 32, ("cb",)), (10000000, 15, ("cb",)), (10000000, 32, ("cb",))):
    c = sorted_coords(n)
    for kind in kinds:
        t, r = best(lambda: eng.mk_nn_indx(c, m, kind), reps=2 if n < 10000000 else 1)
        knn[f"{kind}_n{n}_m{m}"] = dict(seconds=t, locations_per_s=n / t)
        print("knn", kind, n, m, f"{t:.3f}s", flush=True)
    if n == 100000:
        a = eng.mk_nn_indx(c, m, "brute"); b = eng.mk_nn_indx(c, m, "cb")
        knn[f"brute_equals_cb_n{n}_m{m}"] = bool(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]))
out["knn"] = knn

# ---- model at n = 10^6, m = 15 ------------------------------------------------------------------------------
n, m = 1000000, 15
c = sorted_coords(n)
X = np.column_stack([np.ones(n), rng.standard_normal(n)])
beta = np.array([1.0, 5.0])
y = X @ beta + rng.standard_normal(n)
idx, _, lu = eng.mk_nn_indx(c, m, "cb")
eng.set_data(c, idx, lu, X, y, m)
eng.bf_loglik(0, [[5, 1, 6]], [beta], slots=[0])
t, _ = best(lambda: eng.q_resid(0, beta), reps=5); out["q_resid_n1e6_m15_s"] = t
t, _ = best(lambda: eng.q_forms(0), reps=5); out["q_forms_n1e6_m15_p2_s"] = t
print("q", out["q_resid_n1e6_m15_s"], out["q_forms_n1e6_m15_p2_s"], flush=True)

# ---- C4-like: prediction, q sites x S samples ------------------------------------------------------------------
q, S, distinct = 200000, 500, 50
c0 = rng.random((q, 2)) * np.sqrt(n / 5000.0)
X0 = np.column_stack([np.ones(q), rng.standard_normal(q)])
t_nn, nn0 = best(lambda: eng.nn_query(c, c0, m), reps=1)
bs = beta[:, None] + 0.05 * rng.standard_normal((2, S))
th = np.repeat(np.array([[5.0], [1.0], [6.0]]) * (1 + 0.02 * rng.standard_normal((3, distinct))), S // distinct, axis=1)
z = rng.standard_normal(q * S)
t_pred, y0 = best(lambda: eng.rnngp_predict(0, X0, c0, nn0, bs, th, z), reps=1)
out["predict"] = dict(n=n, q=q, S=S, distinct_theta=distinct, nn_query_s=t_nn, predict_s=t_pred,
                      draws_per_s=q * S / t_pred, note="exponential; host buffers in (z: q*S doubles) and q x S draws out")
print("predict", out["predict"], flush=True)

# ---- C2-like: conjugate model, g = 25 parameter sets -------------------------------------------------------------
n2 = 105504
c2 = sorted_coords(n2, scale=20.0)
X2 = np.column_stack([np.ones(n2), rng.standard_normal(n2)])
y2 = X2 @ beta + rng.standard_normal(n2)
idx2, _, lu2 = eng.mk_nn_indx(c2, m, "cb")
eng.set_data(c2, idx2, lu2, X2, y2, m)
ta = np.array([[ph, al] for ph in (0.3, 0.6, 1.5, 3, 6) for al in (0.05, 0.1, 0.25, 0.5, 1)]).T
q2 = 20000
c02 = rng.random((q2, 2)) * 20.0
X02 = np.column_stack([np.ones(q2), rng.standard_normal(q2)])
nn02 = eng.nn_query(c2, c02, m)
t, _ = best(lambda: eng.cnngp(0, ta, [2.0, float(np.var(y2))], X02, c02, nn02), reps=2)
out["cnngp"] = dict(n=n2, g=25, n0=q2, seconds=t, sets_per_s=25 / t)
print("cnngp", out["cnngp"], flush=True)

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r01_extra_timings.json"), "w"), indent=1)
