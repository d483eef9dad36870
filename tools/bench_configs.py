"""BASELINE.json configs[1], [3], [4] at their full sizes on one GPU (not the contract bench, which is configs[2]).
Host wall clock around the public calls (host arrays in, results out).  Writes gpurun_out/r02_config_timings.json.

  configs[1]  spConjNNGP, BCEF shape: n = 188,717, theta.alpha 5 x 5 grid, k.fold = 5, m = 15 -> nngp_conj_kfold + final fit
  configs[3]  predict.NNGP: n0 = 10^6 sites x 1,000 posterior samples at n = 10^6 (four site slices of 250k)
  configs[4]  ordered search n = 10^7, m = 32 code-book; brute force at n = 10^6 (10^7 brute = 5 10^13 pairs, not run),
              brute == cb bit for bit at n = 10^6
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spnngp_b200 import Engine  # noqa: E402

which = set(sys.argv[1:]) or {"kfold", "predict", "knn"}
out = {}
eng = Engine(0)
rng = np.random.default_rng(11)


def timed(f):
    t = time.perf_counter(); r = f(); return time.perf_counter() - t, r


def by_x(c):
    return c[np.argsort(c[:, 0], kind="stable")]


if "kfold" in which:  # ---- configs[1] ----------------------------------------------------------------------------
    m, k_fold = 15, 5
    bcef = os.path.join(ROOT, "data", "bcef.npz")
    real = os.path.exists(bcef)
    if real:  # the reference's data set (tools/read_bcef.py): FCH ~ PTC on all 188,717 rows, ordered by x
        d = np.load(bcef)
        c = np.column_stack([d["x"], d["y"]]); o = np.argsort(c[:, 0], kind="stable"); c = c[o]
        y = d["FCH"][o]; X = np.column_stack([np.ones(len(y)), d["PTC"][o]])
        n = len(y)
    else:
        n = 188717
        c = by_x(rng.random((n, 2)) * np.array([21.5, 17.0]))
        X = np.column_stack([np.ones(n), rng.standard_normal(n)])
        y = X @ np.array([1.0, 5.0]) + 2.0 * np.sin(c[:, 0]) * np.cos(c[:, 1]) + rng.standard_normal(n)
    ta = np.array([[ph, al] for ph in (0.3, 0.6, 1.5, 3, 6) for al in (0.05, 0.1, 0.25, 0.5, 1)]).T
    n_ho = n // k_fold
    ho = np.stack([np.sort(rng.choice(n, n_ho, replace=False)) for _ in range(k_fold)]).astype(np.int32)
    ig = [2.0, float(np.var(y))]
    warm = np.sort(rng.choice(20000, 4000, replace=False))[None, :].astype(np.int32)
    eng.conj_kfold(y[:20000], X[:20000], c[:20000], m, 0, ta[:, :2], ig, warm, "cb")  # warm-up (allocations, module load)
    t, scores = timed(lambda: eng.conj_kfold(y, X, c, m, 0, ta, ig, ho, "cb"))
    best = int(np.argmin(scores[:, 1]))
    t_nn, (idx, _, lu) = timed(lambda: eng.mk_nn_indx(c, m, "cb"))
    eng.set_data(c, idx, lu, X, y, m)
    t_fit, fit = timed(lambda: eng.cnngp(0, ta[:, best:best + 1], ig))
    out["conj_kfold"] = dict(n=n, m=m, g=25, k_fold=k_fold, n_ho=n_ho, kfold_s=t, final_search_s=t_nn, final_fit_s=t_fit,
                             best_set=best, best_theta_alpha=ta[:, best].tolist(), crps_best=float(scores[best, 1]),
                             rmspe_best=float(scores[best, 0]), beta_hat=np.asarray(fit["beta_hat"]).ravel().tolist(),
                             fits_per_s=25 * k_fold / t,
                             data="BCEF (data/BCEF.rda of the reference package, FCH ~ PTC, all 188,717 rows)" if real else
                                  "synthetic sites on BCEF's 21.5 x 17 km extent",
                             note="exponential; 125 fits of ~151k locations + 125 x 37.7k hold-out predictions, scores "
                                  "formed on the device")
    print("kfold", out["conj_kfold"], flush=True)

if "predict" in which:  # ---- configs[3] --------------------------------------------------------------------------
    n, m, q, S, distinct, slices = 1000000, 15, 1000000, 1000, 400, 4
    W = np.sqrt(n / 5000.0)
    c = by_x(rng.random((n, 2)) * W)
    X = np.column_stack([np.ones(n), rng.standard_normal(n)])
    beta = np.array([1.0, 5.0])
    y = X @ beta + rng.standard_normal(n)
    idx, _, lu = eng.mk_nn_indx(c, m, "cb")
    eng.set_data(c, idx, lu, X, y, m)
    bs = beta[:, None] + 0.05 * rng.standard_normal((2, S))
    # a chain with 40 % acceptance: 400 distinct theta columns among the 1,000 samples
    base = np.array([[5.0], [1.0], [6.0], [0.75]]) * (1 + 0.02 * rng.standard_normal((4, distinct)))
    th = base[:, np.sort(rng.integers(0, distinct, S))]
    res = {}
    qs = q // slices
    zp = eng.pinned_array(qs * S); yp = eng.pinned_array(qs * S)  # page-locked arrays (nngp_host_alloc): direct DMA
    for cov, name in ((0, "exponential"), (2, "matern")):
        t_nn = t_pred = t_pin = 0.0
        chk = chk_pin = 0.0
        for s in range(slices):
            c0 = rng.random((qs, 2)) * W
            X0 = np.column_stack([np.ones(qs), rng.standard_normal(qs)])
            z = rng.standard_normal(qs * S)
            dt, nn0 = timed(lambda: eng.nn_query(c, c0, m)); t_nn += dt
            dt, y0 = timed(lambda: eng.rnngp_predict(cov, X0, c0, nn0, bs, th[: (4 if cov == 2 else 3)], z)); t_pred += dt
            chk += float(np.mean(y0))
            del y0
            zp[:] = z
            del z
            dt, y0 = timed(lambda: eng.rnngp_predict(cov, X0, c0, nn0, bs, th[: (4 if cov == 2 else 3)], zp, out=yp)); t_pin += dt
            chk_pin += float(np.mean(y0))
        res[name] = dict(nn_query_s=t_nn, predict_s=t_pred, draws_per_s=q * S / t_pred, mean_of_draws=chk / slices,
                         predict_page_locked_s=t_pin, draws_per_s_page_locked=q * S / t_pin,
                         transfer_GBps_each_way_page_locked=8.0 * q * S / t_pin / 1e9,
                         same_draws=bool(chk == chk_pin))
        print("predict", name, res[name], flush=True)
    eng.free_pinned(zp); eng.free_pinned(yp)
    out["predict"] = dict(n=n, q=q, S=S, distinct_theta=int(len(np.unique(th[2]))), site_slices=slices, **res,
                          note="host arrays in (z: q x S normals, 8 GB) and the q x S draws out (8 GB); the solves run "
                               "once per distinct theta.  predict_s: pageable arrays (what R hands over) staged through the "
                               "pinned rings; predict_page_locked_s: the same call with z and the result in page-locked "
                               "arrays (nngp_host_alloc / cudaHostRegister), moved by direct DMA, both directions at once")

if "knn" in which:  # ---- configs[4] ------------------------------------------------------------------------------
    knn = {}
    c = by_x(rng.random((1000000, 2)) * np.sqrt(1000000 / 5000.0))
    for m in (15, 32):
        tb, a = timed(lambda: eng.mk_nn_indx(c, m, "brute"))
        tc, b = timed(lambda: eng.mk_nn_indx(c, m, "cb"))
        knn[f"n1000000_m{m}"] = dict(brute_s=tb, cb_s=tc, brute_equals_cb=bool(
            np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.int64), b[1].view(np.int64))))
        print("knn", m, knn[f"n1000000_m{m}"], flush=True)
        del a, b
    c = by_x(rng.random((10000000, 2)) * np.sqrt(10000000 / 5000.0))
    tc, b = timed(lambda: eng.mk_nn_indx(c, 32, "cb"))
    knn["n10000000_m32"] = dict(cb_s=tc, locations_per_s=1e7 / tc)
    print("knn 1e7", knn["n10000000_m32"], flush=True)
    out["knn"] = knn

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
path = os.path.join(ROOT, "gpurun_out", "r02_config_timings.json")
old = json.load(open(path)) if os.path.exists(path) else {}
old.update(out)
json.dump(old, open(path, "w"), indent=1)
