"""BASELINE configs[1] on the REAL BCEF data (data/BCEF.rda of the reference package, read by tools/read_bcef.py into
data/bcef.npz; skipped when that file is absent): spConjNNGP's fold loop, R/spConjNNGP.R:287-349, cNNGP src/cNNGP.cpp:215-316."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BCEF = os.path.join(ROOT, "data", "bcef.npz")


def _bcef(fit_rows_only=True):
    if not os.path.exists(BCEF):
        pytest.skip("data/bcef.npz absent (python tools/read_bcef.py needs the reference's data/BCEF.rda)")
    d = np.load(BCEF)
    keep = d["holdout"] == 0 if fit_rows_only else np.ones(len(d["x"]), dtype=bool)
    coords = np.column_stack([d["x"][keep], d["y"][keep]])
    o = np.argsort(coords[:, 0], kind="stable")                      # ord = order(coords[,1]), R/spConjNNGP.R:96
    coords = coords[o]
    y = d["FCH"][keep][o]
    X = np.column_stack([np.ones(len(y)), d["PTC"][keep][o]])       # FCH ~ PTC
    return coords, X, y


def test_bcef_file_is_the_documented_data_set():
    """man/BCEF.rda.Rd:16-20,42-45: 188,717 rows; 105,504 model rows (holdout = 0), all x unique."""
    if not os.path.exists(BCEF):
        pytest.skip("data/bcef.npz absent")
    d = np.load(BCEF)
    assert len(d["x"]) == 188717 and int((d["holdout"] == 0).sum()) == 105504 and int((d["holdout"] == 1).sum()) == 83213
    assert len(np.unique(d["x"])) == 188717
    assert 258.8 < d["x"].min() < 258.9 and 280.4 < d["x"].max() < 280.5


@pytest.mark.gpu
def test_bcef_fold_against_reference(engine, ref, port):
    """One fold x the 5 x 5 theta.alpha grid of SURVEY 8(d) C2 on the model rows of BCEF (n = 105,504): the device's
    whole fold loop (ordered search, query search, 25 conjugate fits, 25 x 21,100 hold-out predictions, crps / rmspe)
    against the reference's own mkNNIndxCB + cNNGP driven through the same R-level loop, <= 1e-9."""
    from oracle import pyoracle as po
    coords, X, y = _bcef()
    n, m = len(y), 15
    ta = np.array([[ph, al] for ph in (0.3, 0.6, 1.5, 3.0, 6.0) for al in (0.05, 0.1, 0.25, 0.5, 1.0)]).T
    rng = np.random.default_rng(7)
    ho = np.sort(rng.choice(n, n // 5, replace=False))[None, :].astype(np.int32)
    ig = [2.0, float(np.var(y))]
    T = os.cpu_count() or 1
    port.set_threads(T)

    class RefT:  # the reference with all host threads
        def mk_nn_indx_cb(self, c, mm):
            return ref.mk_nn_indx_cb(c, mm, threads=T)

        def cnngp(self, *a):
            return ref.cnngp(*a, nThreads=T)

    want = po.conj_kfold(RefT(), port, y, X, coords, m, 0, ta, ig, ho, "cb")
    got = engine.conj_kfold(y, X, coords, m, 0, ta, ig, ho, "cb")
    assert np.all(np.isfinite(got)) and got.shape == (25, 2)
    assert np.max(np.abs(got - want) / np.abs(want)) < 1e-9
    assert int(np.argmin(got[:, 1])) == int(np.argmin(want[:, 1]))


@pytest.mark.gpu
def test_bcef_conjugate_fit_all_rows_against_reference(engine, ref):
    """The final fit of spConjNNGP on all 188,717 rows for three parameter sets: beta.hat, ab, bB.inv against the
    reference's cNNGP (src/cNNGP.cpp:215-250), neighbours by the reference's mkNNIndxCB vs the device's (bit-equal)."""
    coords, X, y = _bcef(fit_rows_only=False)
    n, m = len(y), 15
    T = os.cpu_count() or 1
    idx_r, _, lu_r = ref.mk_nn_indx_cb(coords, m, threads=T)
    idx, _, lu = engine.mk_nn_indx(coords, m, "cb")
    assert np.array_equal(idx, idx_r) and np.array_equal(lu, lu_r)
    ta = np.array([[0.6, 0.1], [3.0, 0.25], [6.0, 1.0]]).T
    ig = [2.0, float(np.var(y))]
    want = ref.cnngp(y, X, coords, ta, idx_r, lu_r, None, None, None, m, ig, 0, nThreads=T)
    engine.set_data(coords, idx, lu, X, y, m)
    got = engine.cnngp(0, ta, ig)
    for k_ref, k_got in (("beta.hat", "beta_hat"), ("ab", "ab"), ("bB.inv", "bB_inv"), ("bb", "bb")):
        a = np.asarray(want[k_ref]).ravel(order="F"); b = np.asarray(got[k_got]).ravel(order="F")  # R's storage order
        assert a.shape == b.shape, (k_ref, a.shape, b.shape)
        if k_ref == "bB.inv":  # dpotri('L') leaves the other triangle of each p x p block untouched (src/cNNGP.cpp:235-236)
            p = X.shape[1]
            keep = np.array([i >= j for _ in range(ta.shape[1]) for j in range(p) for i in range(p)])
            a, b = a[keep], b[keep]
        assert np.max(np.abs(a - b) / np.maximum(np.abs(a), 1e-300)) < 1e-9, (k_ref, a, b)
