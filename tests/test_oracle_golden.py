"""CPU: the plain-C oracle port against the committed golden vectors (outputs of the reference's own
C++, tests/golden/make_golden.py) -- this is what pins the oracle where /root/reference is absent."""
import os

import numpy as np

from tests.util import relerr, relerr_norm

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_bf_golden(port):
    d = np.load(os.path.join(G, "bf_n400_m10.npz"))
    n, m = 400, int(d["m"])
    r = d["X"] @ d["beta"] - d["y"]
    for cov in range(4):
        B, F, ld, info = port.update_bf(d["coords"], d["nnIndx"], d["nnIndxLU"], n, m, d[f"theta{cov}"], cov)
        tol = 1e-12 if cov == 2 else 0.0  # Matern: different Bessel-K implementations on the two sides
        assert info == 0
        assert relerr(F, d[f"F{cov}"]) <= tol and relerr_norm(B, d[f"B{cov}"]) <= tol
        assert abs(ld - float(d[f"logDet{cov}"])) <= max(tol, 1e-15) * abs(ld)
        q = port.q(B, F, r, r, n, d["nnIndx"], d["nnIndxLU"])
        assert abs(q - float(d[f"Q{cov}"])) <= max(tol, 1e-13) * abs(q)


def test_knn_golden(port):
    d = np.load(os.path.join(G, "knn.npz"))
    for kind, ms in (("rand", (5, 15)), ("grid", (8, 15, 32))):
        c = d[f"{kind}_coords"]
        for m in ms:
            a = port.mk_nn_indx(c, m); b = port.mk_nn_indx_cb(c, m)
            assert np.array_equal(a[0], d[f"{kind}_brute_m{m}_idx"])
            assert np.array_equal(a[1].view(np.int64), d[f"{kind}_brute_m{m}_dist"].view(np.int64))
            assert np.array_equal(b[0], d[f"{kind}_cb_m{m}_idx"])
            assert np.array_equal(b[1].view(np.int64), d[f"{kind}_cb_m{m}_dist"].view(np.int64))


def test_brute_equals_cb_without_ties(port):
    d = np.load(os.path.join(G, "knn.npz"))
    assert np.array_equal(d["rand_brute_m15_idx"], d["rand_cb_m15_idx"])          # man/spNNGP.Rd:110-119
    assert not np.array_equal(d["grid_brute_m15_idx"], d["grid_cb_m15_idx"])      # ties: the searches differ


def test_chain_golden(port):
    """Seeded sampler chain: the port consumes the same R-compatible RNG stream (librshim)."""
    from spnngp_b200 import _lib
    R = _lib.load_rshim()
    d = np.load(os.path.join(G, "chain_n300_m10.npz"))
    starting = {"beta": [0.0, 0.0], "sigma.sq": 5.0, "tau.sq": 1.0, "phi": 6.0, "nu": 0.75}
    tuning = {"sigma.sq": 0.05, "tau.sq": 0.05, "phi": 0.05, "nu": 0.05}
    for cov, name in ((0, "exp"), (2, "matern")):
        R.rshim_set_seed(1)
        b, t, acc = port.rnngp(d["y"], d["X"], d["coords"], cov, d["nnIndx"], d["nnIndxLU"], int(d["m"]), [2.0, 5.0],
                               [2.0, 1.0], [3.0, 300.0], [0.1, 2.0] if cov == 2 else None, starting, tuning, 40,
                               lambda mu, sd: R.rnorm(mu, sd), lambda a, bb: R.runif(a, bb))
        assert acc >= 0
        assert np.allclose(t, d[f"{name}_theta"], rtol=1e-9)
        assert np.allclose(b, d[f"{name}_beta"], rtol=1e-7, atol=1e-9)


def test_replicates_golden(port):
    """fit.rep draws and `rNNGPReplicated` of the port vs the reference's committed outputs (rep_n300_m10.npz)."""
    from spnngp_b200 import _lib
    R = _lib.load_rshim()
    d = np.load(os.path.join(G, "rep_n300_m10.npz"))
    starting = {"beta": [0.0, 0.0], "sigma.sq": 5.0, "tau.sq": 1.0, "phi": 6.0, "nu": 0.75}
    tuning = {"sigma.sq": 0.05, "tau.sq": 0.05, "phi": 0.05, "nu": 0.05}
    for cov, name in ((0, "exp"), (2, "matern")):
        R.rshim_set_seed(3)
        b, t, acc, yrep = port.rnngp(d["y"], d["X"], d["coords"], cov, d["nnIndx"], d["nnIndxLU"], int(d["m"]), [2.0, 5.0],
                                     [2.0, 1.0], [3.0, 300.0], [0.1, 2.0] if cov == 2 else None, starting, tuning, 30,
                                     lambda mu, sd: R.rnorm(mu, sd), lambda a, bb: R.runif(a, bb), rep_indx=d["rep_indx"])
        assert np.allclose(t, d[f"{name}_theta"], rtol=1e-9)
        assert np.allclose(yrep, d[f"{name}_yrep"], rtol=1e-8, atol=1e-8)
        R.rshim_set_seed(4)
        r = port.rnngp_replicated(d["X"], d["coords"], cov, d["nnIndx"], d["nnIndxLU"], int(d["m"]),
                                  d[f"{name}_replicated_beta"], d[f"{name}_replicated_theta"], lambda mu, sd: R.rnorm(mu, sd))
        assert np.allclose(r, d[f"{name}_replicated_yrep_seed4"], rtol=1e-9, atol=1e-10)


def test_conj_and_predict_golden(port):
    from spnngp_b200 import _lib
    R = _lib.load_rshim()
    d = np.load(os.path.join(G, "conj_predict_n500_m10.npz"))
    m = int(d["m"])
    for cov, name in ((0, "exp"), (3, "gauss")):
        out = port.cnngp(d["y"], d["X"], d["coords"], d["thetaAlpha"], d["nnIndx"], d["nnIndxLU"], d["X0"], d["coords0"],
                         d["nn0"], m, d["sigmaSqIG"], cov)
        assert out["status"] == 0
        assert np.allclose(out["beta_hat"], d[f"conj_{name}_beta_hat"], rtol=1e-11)
        assert np.allclose(out["ab"][:6], d[f"conj_{name}_ab"].T.ravel()[:6], rtol=1e-12)
        assert np.allclose(out["y0_hat"], d[f"conj_{name}_y_0_hat"], rtol=1e-10, atol=1e-12)
        assert np.allclose(out["y0_hat_var"], d[f"conj_{name}_y_0_hat_var"], rtol=1e-10)
    S = d["pred_beta"].shape[1]; q = d["X0"].shape[0]
    R.rshim_set_seed(7)
    z = np.array([R.rnorm(0.0, 1.0) for _ in range(q * S)])  # reference order z[i*S+s]
    y0, st = port.predict(d["X"], d["y"], d["coords"], m, d["X0"], d["coords0"], d["nn0"], d["pred_beta"],
                          d["pred_theta"], 0, z)
    assert st == 0 and np.allclose(y0, d["pred_y0_seed7"], rtol=1e-11, atol=1e-12)


def test_u_indx_golden(port):
    """`mkUIndx` of the port vs the reference's committed output (uindx_n600_m10.npz), all three search types."""
    d = np.load(os.path.join(G, "uindx_n600_m10.npz"))
    n, m = int(d["n"]), int(d["m"])
    for st in (0, 1, 2):
        u, lu, ui = port.mk_u_indx(n, m, d["nnIndx"], d["nnIndxLU"], search_type=st)
        assert np.array_equal(u, d[f"st{st}_uIndx"]) and np.array_equal(lu, d[f"st{st}_uIndxLU"])
        assert np.array_equal(ui, d[f"st{st}_uiIndx"])


def test_kfold_scores_golden(port):
    """k-fold loop of spConjNNGP + crps / rmspe (R/spConjNNGP.R:286-349): port-backed restatement vs the scores
    obtained with the reference's own mkNNIndxCB and cNNGP (fixture 7 of make_golden.py)."""
    from oracle import pyoracle as po
    d = np.load(os.path.join(G, "kfold_n600_m10.npz"))
    for cov, name in ((0, "exp"), (2, "matern")):
        got = po.conj_kfold(port, port, d["y"], d["X"], d["coords"], int(d["m"]), cov, d[f"{name}_theta_alpha"],
                            d["sigmaSqIG"], d["ho_idx"], "cb")
        assert relerr(got, d[f"{name}_scores"]) < (1e-10 if cov == 2 else 1e-12)


def test_score_rules_known_answers():
    """crps of a perfect forecast with unit variance: -(1/sqrt(pi) - 2 dnorm(0)) = (sqrt(2) - 1)/sqrt(pi)."""
    from oracle import pyoracle as po
    assert abs(po.crps([0.0, 3.0], [0.0, 3.0], [1.0, 1.0]) - (np.sqrt(2.0) - 1.0) / np.sqrt(np.pi)) < 1e-15
    assert abs(po.rmspe([1.0, 2.0, 3.0], [1.0, 2.0, 5.0]) - np.sqrt(4.0 / 3.0)) < 1e-15
