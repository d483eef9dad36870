"""GPU parity: kNN + solve prediction paths (rNNGPPredict, cNNGP) through the C ABI vs the CPU oracle."""
import os

import numpy as np
import pytest

from tests.util import make_problem

pytestmark = pytest.mark.gpu


def _chain_like_samples(rng, S, cov, p):
    """theta columns with repeats (rejected proposals repeat the previous draw)."""
    nTheta = 4 if cov == 2 else 3
    th = np.zeros((nTheta, S)); cur = np.array([5.0, 1.0, 6.0, 0.75][:nTheta])
    for s in range(S):
        if rng.random() < 0.4:
            cur = cur * np.exp(0.05 * rng.standard_normal(nTheta))
            if cov == 2:
                cur[3] = min(max(cur[3], 0.2), 1.9)
        th[:, s] = cur
    beta = np.array([1.0, 5.0, -2.0][:p])[:, None] + 0.1 * rng.standard_normal((p, S))
    return beta, th


@pytest.mark.parametrize("cov,m", [(0, 10), (0, 15), (2, 10), (3, 20), (1, 5)])
def test_rnngp_predict_parity(engine, port, cov, m):
    rng = np.random.default_rng(100 + cov + m)
    prob = make_problem(1500, m, seed=31, port=port, domain=0.55)
    q, S = 333, 37
    coords0 = rng.random((q, 2)) * 0.55
    X0 = np.column_stack([np.ones(q), rng.standard_normal(q)])
    nn0 = port.nn_query(prob["coords"], coords0, m)
    beta, theta = _chain_like_samples(rng, S, cov, 2)
    z = rng.standard_normal(q * S)
    ref, st = port.predict(prob["X"], prob["y"], prob["coords"], m, X0, coords0, nn0, beta, theta, cov, z)
    assert st == 0
    engine.set_data(prob["coords"], prob["nnIndx"], prob["nnIndxLU"], prob["X"], prob["y"], m)
    got = engine.rnngp_predict(cov, X0, coords0, nn0, beta, theta, z)
    scale = np.abs(ref).max()
    assert np.max(np.abs(got - ref)) < 1e-10 * scale
    # site sharding: two halves reproduce the full result
    a = engine.rnngp_predict(cov, X0, coords0, nn0, beta, theta, z, site_range=(0, 100))
    b = engine.rnngp_predict(cov, X0, coords0, nn0, beta, theta, z, site_range=(100, q))
    assert np.array_equal(a[:100], got[:100]) and np.array_equal(b[100:], got[100:])
    # prediction-only context: no neighbour structure of the fit is needed (the C ABI's nnIndx = NULL mode)
    engine.set_data(prob["coords"], None, None, prob["X"], prob["y"], m)
    assert np.array_equal(engine.rnngp_predict(cov, X0, coords0, nn0, beta, theta, z), got)
    with pytest.raises(Exception):
        engine.bf_loglik(cov, [theta[:, 0]], None)  # a likelihood needs the neighbours


def test_rnngp_predict_many_chunks_full_duplex(engine, port):
    """Several site chunks (NNGP_PREDICT_CHUNK = 512 sites): the overlapped download of chunk k with the upload
    and solves of chunk k + 1 (two result buffers, second pinned ring) leaves every draw where the serial path puts it."""
    rng = np.random.default_rng(9)
    m, q, S = 10, 3000, 64
    prob = make_problem(1200, m, seed=33, port=port, domain=0.5)
    coords0 = rng.random((q, 2)) * 0.5
    X0 = np.column_stack([np.ones(q), rng.standard_normal(q)])
    nn0 = port.nn_query(prob["coords"], coords0, m)
    theta = np.tile(np.array([[5.0], [1.0], [6.0]]), (1, S)); theta[2, S // 2:] = 7.0
    os.environ["NNGP_PREDICT_CHUNK"] = "512"
    beta = np.array([[1.0], [5.0]]) + 0.1 * rng.standard_normal((2, S))
    z = rng.standard_normal(q * S)
    engine.set_data(prob["coords"], prob["nnIndx"], prob["nnIndxLU"], prob["X"], prob["y"], m)
    try:
        got = engine.rnngp_predict(0, X0, coords0, nn0, beta, theta, z)
        os.environ["NNGP_PREDICT_DUPLEX"] = "0"
        serial = engine.rnngp_predict(0, X0, coords0, nn0, beta, theta, z)
    finally:
        os.environ.pop("NNGP_PREDICT_DUPLEX", None)
        os.environ.pop("NNGP_PREDICT_CHUNK", None)
    assert np.array_equal(got, serial)
    sub = slice(0, 40)  # and against the oracle on a slice of the samples
    ref, st = port.predict(prob["X"], prob["y"], prob["coords"], m, X0, coords0, nn0, beta[:, sub], theta[:, sub], 0,
                           z.reshape(q, S)[:, sub].ravel())
    assert st == 0 and np.max(np.abs(got[:, sub] - ref)) < 1e-10 * np.abs(ref).max()


def test_rnngp_predict_page_locked_arrays(engine, port):
    """Normals and draws placed in page-locked arrays (nngp_host_alloc) travel by direct DMA instead of through the
    staging ring (host_stage.h): same draws bit for bit as with pageable arrays, one chunk and many chunks."""
    rng = np.random.default_rng(10)
    m, q, S = 10, 5000, 80  # z: 3.2 MB, above the 2 MB threshold of the direct path
    prob = make_problem(1500, m, seed=34, port=port, domain=0.5)
    coords0 = rng.random((q, 2)) * 0.5
    X0 = np.column_stack([np.ones(q), rng.standard_normal(q)])
    nn0 = port.nn_query(prob["coords"], coords0, m)
    beta, theta = _chain_like_samples(rng, S, 2, 2)
    engine.set_data(prob["coords"], prob["nnIndx"], prob["nnIndxLU"], prob["X"], prob["y"], m)
    z = rng.standard_normal(q * S)
    want = engine.rnngp_predict(2, X0, coords0, nn0, beta, theta, z)
    zp = engine.pinned_array(q * S); yp = engine.pinned_array(q * S)
    try:
        zp[:] = z
        yp[:] = -1.0
        got = engine.rnngp_predict(2, X0, coords0, nn0, beta, theta, zp, out=yp)
        assert np.array_equal(got, want)
        os.environ["NNGP_PREDICT_CHUNK"] = "1024"
        yp[:] = -1.0
        got = engine.rnngp_predict(2, X0, coords0, nn0, beta, theta, zp, out=yp)
        assert np.array_equal(got, want)
        got = got.copy()
    finally:
        os.environ.pop("NNGP_PREDICT_CHUNK", None)
        engine.free_pinned(zp); engine.free_pinned(yp)


@pytest.mark.parametrize("cov,m", [(0, 15), (2, 10), (3, 20)])
def test_cnngp_parity(engine, port, cov, m):
    rng = np.random.default_rng(7 + cov)
    prob = make_problem(2000, m, seed=33, port=port, domain=0.63)
    n0 = 400
    coords0 = rng.random((n0, 2)) * 0.63
    X0 = np.column_stack([np.ones(n0), rng.standard_normal(n0)])
    nn0 = port.nn_query(prob["coords"], coords0, m)
    phis = [3.0, 6.0, 12.0]; alphas = [0.1, 0.5]
    rows = [[ph, al] + ([0.8] if cov == 2 else []) for ph in phis for al in alphas]
    ta = np.array(rows).T  # nTheta x g
    ig = [2.0, float(np.var(prob["y"]))]
    ref = port.cnngp(prob["y"], prob["X"], prob["coords"], ta, prob["nnIndx"], prob["nnIndxLU"], X0, coords0, nn0, m, ig, cov)
    assert ref["status"] == 0
    engine.set_data(prob["coords"], prob["nnIndx"], prob["nnIndxLU"], prob["X"], prob["y"], m)
    got = engine.cnngp(cov, ta, ig, X0, coords0, nn0)
    g = ta.shape[1]
    assert np.allclose(got["beta_hat"], ref["beta_hat"], rtol=1e-9, atol=1e-12)
    assert np.allclose(got["ab"][: 2 * g], ref["ab"][: 2 * g], rtol=1e-10)
    assert np.allclose(got["bb"], ref["bb"], rtol=1e-10)
    p = 2
    for k in range(g):  # the reference only defines the lower triangle of bB.inv
        G = got["bB_inv"][:, k].reshape(p, p).T; R = ref["bB_inv"][:, k].reshape(p, p).T
        assert np.allclose(np.tril(G), np.tril(R), rtol=1e-9)
    assert np.allclose(got["y0_hat"], ref["y0_hat"], rtol=1e-9, atol=1e-10 * np.abs(ref["y0_hat"]).max())
    assert np.allclose(got["y0_hat_var"], ref["y0_hat_var"], rtol=1e-9)


def test_cnngp_without_prediction(engine, port):
    prob = make_problem(900, 10, seed=35, port=port, domain=0.4)
    ta = np.array([[6.0, 0.2], [3.0, 1.0]]).T
    ref = port.cnngp(prob["y"], prob["X"], prob["coords"], ta, prob["nnIndx"], prob["nnIndxLU"], None, None, None, 10,
                     [2.0, 1.0], 0)
    engine.set_data(prob["coords"], prob["nnIndx"], prob["nnIndxLU"], prob["X"], prob["y"], 10)
    got = engine.cnngp(0, ta, [2.0, 1.0])
    assert np.allclose(got["beta_hat"], ref["beta_hat"], rtol=1e-9)
    assert np.allclose(got["ab"][:4], ref["ab"][:4], rtol=1e-10)


@pytest.mark.parametrize("cov,m", [(0, 15), (2, 10)])
def test_cnngp_score_on_device(engine, port, cov, m):
    """crps / rmspe formed on the GPU (nngp_cnngp_score) vs the R-layer restatement applied to the oracle's cNNGP."""
    from oracle import pyoracle as po
    rng = np.random.default_rng(17 + cov)
    prob = make_problem(2000, m, seed=37, port=port, domain=0.63)
    n0 = 700
    coords0 = rng.random((n0, 2)) * 0.63
    X0 = np.column_stack([np.ones(n0), rng.standard_normal(n0)])
    y0 = X0 @ prob["beta"] + np.sqrt(5.0) * (np.sin(3 * coords0[:, 0]) + np.cos(2 * coords0[:, 1])) + rng.standard_normal(n0)
    nn0 = port.nn_query(prob["coords"], coords0, m)
    rows = [[ph, al] + ([0.8] if cov == 2 else []) for ph in (3.0, 12.0) for al in (0.1, 0.5)]
    ta = np.array(rows).T
    ig = [2.0, float(np.var(prob["y"]))]
    ref = port.cnngp(prob["y"], prob["X"], prob["coords"], ta, prob["nnIndx"], prob["nnIndxLU"], X0, coords0, nn0, m, ig, cov)
    engine.set_data(prob["coords"], prob["nnIndx"], prob["nnIndxLU"], prob["X"], prob["y"], m)
    crps, rmspe = engine.cnngp_score(cov, ta, ig, X0, coords0, nn0, y0)
    for k in range(ta.shape[1]):
        assert abs(crps[k] - po.crps(y0, ref["y0_hat"][:, k], ref["y0_hat_var"][:, k])) < 1e-10 * abs(crps[k])
        assert abs(rmspe[k] - po.rmspe(y0, ref["y0_hat"][:, k])) < 1e-10 * rmspe[k]
    again = engine.cnngp_score(cov, ta, ig, X0, coords0, nn0, y0)
    assert np.array_equal(again[0], crps) and np.array_equal(again[1], rmspe)  # deterministic reduction


def test_conj_kfold_golden(engine):
    """The whole k-fold loop in one call (nngp_conj_kfold) vs scores obtained with the reference's own mkNNIndxCB and
    cNNGP (tests/golden/kfold_n600_m10.npz); parameter-set shards add up to the whole."""
    import os
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kfold_n600_m10.npz"))
    for cov, name in ((0, "exp"), (2, "matern")):
        ta = d[f"{name}_theta_alpha"]
        got = engine.conj_kfold(d["y"], d["X"], d["coords"], int(d["m"]), cov, ta, d["sigmaSqIG"], d["ho_idx"], "cb")
        assert np.max(np.abs(got - d[f"{name}_scores"]) / np.abs(d[f"{name}_scores"])) < 1e-10
        a = engine.conj_kfold(d["y"], d["X"], d["coords"], int(d["m"]), cov, ta, d["sigmaSqIG"], d["ho_idx"], "cb", set_range=(0, 1))
        b = engine.conj_kfold(d["y"], d["X"], d["coords"], int(d["m"]), cov, ta, d["sigmaSqIG"], d["ho_idx"], "cb", set_range=(1, 4))
        assert np.array_equal(a + b, got)
    brute = engine.conj_kfold(d["y"], d["X"], d["coords"], int(d["m"]), 0, d["exp_theta_alpha"], d["sigmaSqIG"], d["ho_idx"], "brute")
    assert np.max(np.abs(brute - d["exp_scores"]) / np.abs(d["exp_scores"])) < 1e-10  # tie-free sites: brute == cb


@pytest.mark.parametrize("cov,search", [(0, "cb"), (3, "brute")])
def test_conj_kfold_midsize_vs_oracle(engine, port, cov, search):
    """nngp_conj_kfold at n = 12,000 (2 folds, 3 parameter sets) against the oracle's restatement of the R loop:
    ordered search, query search, fit and scores all go through the device path."""
    from oracle import pyoracle as po
    n, m, k_fold = 12000, 15, 2
    prob = make_problem(n, m, seed=51 + cov, domain=np.sqrt(n / 5000.0))
    rng = np.random.default_rng(52)
    ho = np.stack([np.sort(rng.choice(n, n // 4, replace=False)) for _ in range(k_fold)]).astype(np.int32)
    ta = np.array([[3.0, 0.1], [6.0, 0.3], [12.0, 1.0]]).T
    ig = [2.0, float(np.var(prob["y"]))]
    port.set_threads(8)
    ref = po.conj_kfold(port, port, prob["y"], prob["X"], prob["coords"], m, cov, ta, ig, ho, search)
    got = engine.conj_kfold(prob["y"], prob["X"], prob["coords"], m, cov, ta, ig, ho, search)
    assert np.max(np.abs(got - ref) / np.abs(ref)) < 1e-10
