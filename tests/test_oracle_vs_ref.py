"""CPU: the oracle port against the reference's own C++ compiled here (oracle/_ref).  Skipped where
/root/reference was not available at build time (then tests/test_oracle_golden.py is the pin)."""
import numpy as np
import pytest

from tests.util import make_problem, relerr, relerr_norm


@pytest.mark.parametrize("cov,theta", [(0, [5, 1, 6]), (1, [5, 1, 6]), (2, [5, 1, 6, 0.75]), (2, [2, 0.3, 9, 2.3]),
                                       (3, [5, 1, 6])])
def test_update_bf_and_q(port, ref, cov, theta):
    prob = make_problem(1500, 15, seed=3, port=port, domain=0.55)
    n, m = 1500, 15
    B1, F1, l1, info = port.update_bf(prob["coords"], prob["nnIndx"], prob["nnIndxLU"], n, m, theta, cov)
    B2, F2, l2 = ref.update_bf(prob["coords"], prob["nnIndx"], prob["nnIndxLU"], n, m, theta, cov)
    tol = 2e-13 if cov == 2 else 0.0
    assert info == 0 and relerr(F1, F2) <= tol and relerr_norm(B1, B2) <= tol and abs(l1 - l2) <= 1e-14 * abs(l2) + tol
    u = prob["X"][:, 1].copy(); v = prob["y"]
    assert abs(port.q(B1, F1, u, v, n, prob["nnIndx"], prob["nnIndxLU"]) -
               ref.q(B2, F2, u, v, n, prob["nnIndx"], prob["nnIndxLU"])) <= 1e-12 * abs(l2)


def test_conj_bf(port, ref):
    prob = make_problem(800, 10, seed=4, port=port, domain=0.4)
    B1, F1, info = port.update_conj_bf(prob["coords"], prob["nnIndx"], prob["nnIndxLU"], 800, 10, 6.0, 0.3, 0.0, 0)
    B2, F2 = ref.update_conj_bf(prob["coords"], prob["nnIndx"], prob["nnIndxLU"], 800, 10, 6.0, 0.3, 0.0, 0)
    assert info == 0 and np.array_equal(F1, F2) and np.array_equal(B1, B2)


@pytest.mark.parametrize("m", [1, 8, 9, 15, 32])
def test_knn_port_equals_ref(port, ref, m):
    rng = np.random.default_rng(m)
    c = rng.random((1200, 2)); c = c[np.argsort(c[:, 0])]
    for f1, f2 in ((port.mk_nn_indx, ref.mk_nn_indx), (port.mk_nn_indx_cb, ref.mk_nn_indx_cb)):
        a, b = f1(c, m), f2(c, m)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.int64), b[1].view(np.int64))
        assert np.array_equal(a[2], b[2])


def test_chain_port_equals_ref(port, ref):
    """Whole `rNNGP` entry point of the reference through the shim vs the port's sampler, same seed."""
    prob = make_problem(400, 10, seed=5, port=port, domain=0.28)
    starting = {"beta": [0.0, 0.0], "sigma.sq": 5.0, "tau.sq": 1.0, "phi": 6.0}
    tuning = {"sigma.sq": 0.05, "tau.sq": 0.05, "phi": 0.05}
    ref.set_seed(11)
    out = ref.rnngp(prob["y"], prob["X"], prob["coords"], 0, prob["nnIndx"], prob["nnIndxLU"], 10, [2.0, 5.0], [2.0, 1.0],
                    [3.0, 300.0], None, starting, tuning, 50)
    ref.set_seed(11)
    L = ref.lib
    b, t, acc = port.rnngp(prob["y"], prob["X"], prob["coords"], 0, prob["nnIndx"], prob["nnIndxLU"], 10, [2.0, 5.0],
                           [2.0, 1.0], [3.0, 300.0], None, starting, tuning, 50,
                           lambda mu, sd: L.rnorm(mu, sd), lambda a, bb: L.runif(a, bb))
    assert np.allclose(t, out["p_theta_samples"], rtol=1e-10) and np.allclose(b, out["p_beta_samples"], rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("n,m", [(700, 10), (40, 15), (300, 1), (9, 3)])
def test_u_indx_port_equals_ref(port, ref, n, m):
    """`mkUIndx` (transposed neighbour graph), the reference's three search types (1 lists descending) vs the port."""
    prob = make_problem(n, m, seed=40 + m, port=port)
    for st in (0, 1, 2):
        want = port.mk_u_indx(n, m, prob["nnIndx"], prob["nnIndxLU"], search_type=st)
        got = ref.mk_u_indx(n, m, prob["nnIndx"], prob["nnIndxLU"], search_type=st)
        nI = len(prob["nnIndx"])
        assert np.array_equal(got[0][:nI], want[0][:nI]) and np.array_equal(got[1], want[1])
        assert np.array_equal(got[2][:nI], want[2][:nI])


def test_replicates_port_equals_ref(port, ref):
    """fit.rep inside `rNNGP` and the `rNNGPReplicated` entry point: reference vs port, same seed."""
    prob = make_problem(350, 8, seed=15, port=port, domain=0.27)
    starting = {"beta": [0.0, 0.0], "sigma.sq": 5.0, "tau.sq": 1.0, "phi": 6.0}
    tuning = {"sigma.sq": 0.05, "tau.sq": 0.05, "phi": 0.05}
    rep_indx = np.zeros(24, dtype=np.int32); rep_indx[8::3] = 1
    args = (prob["y"], prob["X"], prob["coords"], 0, prob["nnIndx"], prob["nnIndxLU"], 8, [2.0, 5.0], [2.0, 1.0],
            [3.0, 300.0], None, starting, tuning, 24)
    ref.set_seed(21)
    out = ref.rnngp(*args, rep_indx=rep_indx)
    assert out["y_rep_samples"].shape == (350, int(rep_indx.sum()))
    ref.set_seed(21)
    L = ref.lib
    b, t, acc, yrep = port.rnngp(*args, lambda mu, sd: L.rnorm(mu, sd), lambda a, bb: L.runif(a, bb), rep_indx=rep_indx)
    assert np.allclose(t, out["p_theta_samples"], rtol=1e-10)
    assert np.allclose(yrep, out["y_rep_samples"], rtol=1e-9, atol=1e-9)
    bsm = out["p_beta_samples"][:, -5:].T.copy(); tsm = out["p_theta_samples"][:, -5:].copy()
    ref.set_seed(22)
    r1 = ref.rnngp_replicated(prob["X"], prob["coords"], 0, prob["nnIndx"], prob["nnIndxLU"], 8, bsm, tsm)
    ref.set_seed(22)
    r2 = port.rnngp_replicated(prob["X"], prob["coords"], 0, prob["nnIndx"], prob["nnIndxLU"], 8, bsm, tsm,
                               lambda mu, sd: L.rnorm(mu, sd))
    assert r1.shape == (350, 5) and np.allclose(r1, r2, rtol=1e-10, atol=1e-11)


def test_reference_error_path(ref, port):
    """The reference aborts with 'dpotrf failed' on a singular neighbour covariance (src/rNNGP.cpp:86)."""
    prob = make_problem(60, 5, seed=6, port=port)
    coords = prob["coords"].copy(); coords[30] = coords[29]; coords[31] = coords[29]
    nnIndx, _, lu = port.mk_nn_indx(coords, 5)
    starting = {"beta": [0.0, 0.0], "sigma.sq": 5.0, "tau.sq": 0.0, "phi": 6.0}
    tuning = {"sigma.sq": 0.05, "tau.sq": 0.05, "phi": 0.05}
    with pytest.raises(RuntimeError, match="dpotrf failed"):
        ref.rnngp(prob["y"], prob["X"], coords, 0, nnIndx, lu, 5, [2.0, 5.0], [2.0, 1.0], [3.0, 300.0], None, starting,
                  tuning, 3)
