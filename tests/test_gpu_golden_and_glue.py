"""GPU: (1) the CUDA path against the committed golden vectors (outputs of the reference's own C++);
(2) the drop-in `.Call` layer (libspNNGP_b200.so: rNNGP, cNNGP, rNNGPPredict, mkNNIndx, mkNNIndxCB with the
reference's SEXP signatures) driven exactly like the reference's entry points, same seeds."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from tests.util import make_problem, relerr, relerr_norm

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def glue():
    lib = C.CDLL(os.path.join(ROOT, "spnngp_b200", "libspNNGP_b200.so"))
    return po._Sexp(lib)


def test_bf_against_golden(engine):
    d = np.load(os.path.join(G, "bf_n400_m10.npz"))
    m = int(d["m"])
    engine.set_data(d["coords"], d["nnIndx"], d["nnIndxLU"], d["X"], d["y"], m)
    for cov in range(4):
        out = engine.bf_loglik(cov, [d[f"theta{cov}"]], [d["beta"]], slots=[0])
        B, F = engine.get_bf(0)
        assert relerr(F, d[f"F{cov}"]) < 1e-10 and relerr_norm(B[: len(d["nnIndx"])], d[f"B{cov}"][: len(d["nnIndx"])]) < 1e-10
        assert abs(out[0, 0] - float(d[f"logDet{cov}"])) < 1e-10 * abs(out[0, 0])
        assert abs(out[0, 1] - float(d[f"Q{cov}"])) < 1e-10 * abs(out[0, 1])


def test_knn_against_golden(engine):
    d = np.load(os.path.join(G, "knn.npz"))
    for kind, ms in (("rand", (5, 15)), ("grid", (8, 15, 32))):
        c = d[f"{kind}_coords"]
        for m in ms:
            for name, st in (("brute", "brute"), ("cb", "cb")):
                idx, dist, _ = engine.mk_nn_indx(c, m, st)
                assert np.array_equal(idx, d[f"{kind}_{name}_m{m}_idx"])
                assert np.array_equal(dist.view(np.int64), d[f"{kind}_{name}_m{m}_dist"].view(np.int64))


def test_glue_rnngp_chain_against_golden(glue):
    """`rNNGP` .Call symbol, set.seed(1): retraces the reference's chain (golden = reference output)."""
    d = np.load(os.path.join(G, "chain_n300_m10.npz"))
    starting = {"beta": [0.0, 0.0], "sigma.sq": 5.0, "tau.sq": 1.0, "phi": 6.0, "nu": 0.75}
    tuning = {"sigma.sq": 0.05, "tau.sq": 0.05, "phi": 0.05, "nu": 0.05}
    for cov, name in ((0, "exp"), (2, "matern")):
        for mode in ("0", "1"):
            os.environ["NNGP_SAMPLER_MODE"] = mode
            glue.lib.rshim_set_seed(1)
            out = po.call_rnngp(glue, d["y"], d["X"], d["coords"], cov, d["nnIndx"], d["nnIndxLU"], int(d["m"]),
                                [2.0, 5.0], [2.0, 1.0], [3.0, 300.0], [0.1, 2.0] if cov == 2 else None, starting,
                                tuning, 40)
            assert np.allclose(out["p_theta_samples"], d[f"{name}_theta"], rtol=1e-9)
            assert np.allclose(out["p_beta_samples"], d[f"{name}_beta"], rtol=1e-7, atol=1e-9)
    os.environ.pop("NNGP_SAMPLER_MODE", None)


def test_glue_cnngp_and_predict_against_golden(glue):
    d = np.load(os.path.join(G, "conj_predict_n500_m10.npz"))
    m = int(d["m"])
    for cov, name in ((0, "exp"), (3, "gauss")):
        out = po.call_cnngp(glue, d["y"], d["X"], d["coords"], d["thetaAlpha"], d["nnIndx"], d["nnIndxLU"], d["X0"],
                            d["coords0"], d["nn0"], m, d["sigmaSqIG"], cov)
        assert np.allclose(out["beta.hat"], d[f"conj_{name}_beta_hat"], rtol=1e-9)
        assert np.allclose(out["ab"], d[f"conj_{name}_ab"], rtol=1e-10)
        assert np.allclose(out["bb"], d[f"conj_{name}_bb"], rtol=1e-10)
        assert np.allclose(out["y.0.hat"], d[f"conj_{name}_y_0_hat"], rtol=1e-9, atol=1e-10)
        assert np.allclose(out["y.0.hat.var"], d[f"conj_{name}_y_0_hat_var"], rtol=1e-9)
    glue.lib.rshim_set_seed(7)
    y0 = po.call_predict(glue, d["X"], d["y"], d["coords"], m, d["X0"], d["coords0"], d["nn0"], d["pred_beta"],
                         d["pred_theta"], 0)
    assert np.allclose(y0, d["pred_y0_seed7"], rtol=1e-9, atol=1e-10)


def test_glue_mk_nn_in_place(glue, port):
    """mkNNIndx / mkNNIndxCB write into their arguments and return R_NilValue (R/nn.R:45,71)."""
    c = np.random.default_rng(2).random((1500, 2)); c = c[np.argsort(c[:, 0])]
    for sym, fn in (("mkNNIndx", port.mk_nn_indx), ("mkNNIndxCB", port.mk_nn_indx_cb)):
        got = po.call_mk_nn(glue, c, 15, symbol=sym)
        ref = fn(c, 15)
        assert all(np.array_equal(a, b) for a, b in zip(got, ref))


def test_glue_mk_u_indx_against_golden(glue):
    """`mkUIndx` .Call symbol: in-place outputs equal the reference's (uindx_n600_m10.npz)."""
    d = np.load(os.path.join(G, "uindx_n600_m10.npz"))
    for st in (0, 1, 2):
        u, lu, ui = po.call_mk_u_indx(glue, int(d["n"]), int(d["m"]), d["nnIndx"], d["nnIndxLU"], search_type=st)
        assert np.array_equal(u, d[f"st{st}_uIndx"]) and np.array_equal(lu, d[f"st{st}_uIndxLU"])
        assert np.array_equal(ui, d[f"st{st}_uiIndx"])


def test_glue_error_text_matches_reference(glue, port):
    prob = make_problem(60, 5, seed=6, port=port)
    coords = prob["coords"].copy(); coords[30] = coords[29]; coords[31] = coords[29]
    nnIndx, _, lu = port.mk_nn_indx(coords, 5)
    starting = {"beta": [0.0, 0.0], "sigma.sq": 5.0, "tau.sq": 0.0, "phi": 6.0}
    tuning = {"sigma.sq": 0.05, "tau.sq": 0.05, "phi": 0.05}
    with pytest.raises(RuntimeError, match="dpotrf failed"):
        po.call_rnngp(glue, prob["y"], prob["X"], coords, 0, nnIndx, lu, 5, [2.0, 5.0], [2.0, 1.0], [3.0, 300.0], None,
                      starting, tuning, 3)
    with pytest.raises(RuntimeError, match="outside the hot path"):
        glue.call("PGLogit", [glue.scalar_int(1)] * 9)


def test_glue_vs_live_reference_when_built(glue, ref, port):
    """Where oracle/_ref exists: the two `.Call` libraries side by side on a fresh problem."""
    prob = make_problem(1200, 15, seed=8, port=port, domain=0.5)
    starting = {"beta": [0.0, 0.0], "sigma.sq": 5.0, "tau.sq": 1.0, "phi": 6.0}
    tuning = {"sigma.sq": 0.1, "tau.sq": 0.1, "phi": 0.1}
    args = (prob["y"], prob["X"], prob["coords"], 0, prob["nnIndx"], prob["nnIndxLU"], 15, [2.0, 5.0], [2.0, 1.0],
            [3.0, 300.0], None, starting, tuning, 80)
    ref.set_seed(5); a = ref.rnngp(*args)
    glue.lib.rshim_set_seed(5); b = po.call_rnngp(glue, *args)
    assert np.allclose(a["p_theta_samples"], b["p_theta_samples"], rtol=1e-9)
    assert np.allclose(a["p_beta_samples"], b["p_beta_samples"], rtol=1e-7, atol=1e-9)


def test_glue_kfold_symbol_against_golden(glue):
    """`cNNGPKFold` (the one added .Call symbol): 1-based n.ho x k.fold hold-out matrix in, k.fold.scores/k.fold out."""
    d = np.load(os.path.join(G, "kfold_n600_m10.npz"))
    ho = d["ho_idx"]; k_fold, n_ho = ho.shape
    y = d["y"]; n = len(y); X = d["X"]; p = X.shape[1]; ta = d["exp_theta_alpha"]; g = ta.shape[1]
    args = [glue.real(y), glue.real(np.ascontiguousarray(X.T).ravel()), glue.scalar_int(p), glue.scalar_int(n),
            glue.real(po.coords_flat(d["coords"])), glue.real(np.ascontiguousarray(ta.T).ravel()), glue.scalar_int(g),
            glue.scalar_int(int(d["m"])), glue.real(d["sigmaSqIG"]), glue.scalar_int(0), glue.scalar_int(1),
            glue.integer((ho + 1).ravel()), glue.scalar_int(n_ho), glue.scalar_int(k_fold)]
    out = glue.call("cNNGPKFold", args)
    got = np.ctypeslib.as_array(glue.lib.REAL(out), shape=(2 * g,)).copy().reshape(2, g).T
    glue.reset()
    assert np.max(np.abs(got - d["exp_scores"]) / np.abs(d["exp_scores"])) < 1e-10


def _config0_problem(port):
    """BASELINE configs[0] / SURVEY 8(d) C1: n = 5,000 sites ~ U(0,1)^2, x = [1, N(0,1)], beta = (1, 5), sigma.sq = 5,
    tau.sq = 1, phi = 6 (man/spNNGP.Rd:231-246), y = X beta + w + eps with w a dense exponential-GP draw; m = 15,
    neighbours by the code-book search."""
    n, m = 5000, 15
    rng = np.random.default_rng(1)
    coords = rng.random((n, 2))
    coords = coords[np.argsort(coords[:, 0], kind="stable")]
    X = np.column_stack([np.ones(n), rng.standard_normal(n)])
    d = np.sqrt(((coords[:, None, :] - coords[None, :, :]) ** 2).sum(axis=2))
    L = np.linalg.cholesky(5.0 * np.exp(-6.0 * d) + 1e-10 * np.eye(n))
    y = X @ np.array([1.0, 5.0]) + L @ rng.standard_normal(n) + rng.standard_normal(n)
    idx, dist, lu = port.mk_nn_indx_cb(coords, m)
    return dict(n=n, m=m, coords=coords, X=X, y=y, nnIndx=idx, nnIndxLU=lu)


@pytest.mark.parametrize("tune", [0.5, 0.05])
def test_config0_chain_1000_iterations_against_reference(glue, ref, port, tune):
    """BASELINE configs[0] verbatim: spNNGP method = 'response', n = 5,000, exponential, n.neighbors = 15, 1,000 MCMC
    samples, search.type = 'cb', set.seed(1).  The whole theta / beta chain of the `.Call` layer over the CUDA engine,
    in both sampler modes, against the reference's own `rNNGP` (src/rNNGP.cpp:272-378) run on the same inputs
    with the same RNG stream: identical accept decisions, theta to 1e-9, beta to 1e-7."""
    prob = _config0_problem(port)
    starting = {"beta": [0.0, 0.0], "sigma.sq": 5.0, "tau.sq": 1.0, "phi": 6.0}
    tuning = {"sigma.sq": tune, "tau.sq": tune, "phi": tune}
    args = (prob["y"], prob["X"], prob["coords"], 0, prob["nnIndx"], prob["nnIndxLU"], prob["m"], [2.0, 5.0], [2.0, 1.0],
            [3.0, 300.0], None, starting, tuning, 1000)
    ref.set_seed(1)
    want = ref.rnngp(*args, nThreads=os.cpu_count())
    moves = int((np.diff(want["p_theta_samples"][0]) != 0).sum())
    assert moves > (100 if tune < 0.1 else 3)  # the chain does move: accept decisions are being compared
    for mode in ("0", "1"):
        os.environ["NNGP_SAMPLER_MODE"] = mode
        try:
            glue.lib.rshim_set_seed(1)
            got = po.call_rnngp(glue, *args)
        finally:
            os.environ.pop("NNGP_SAMPLER_MODE", None)
        assert int((np.diff(got["p_theta_samples"][0]) != 0).sum()) == moves
        assert np.allclose(got["p_theta_samples"], want["p_theta_samples"], rtol=1e-9, atol=0)
        assert np.allclose(got["p_beta_samples"], want["p_beta_samples"], rtol=1e-7, atol=1e-9)


@pytest.mark.parametrize("tuning", [{"sigma.sq": 0.05, "tau.sq": 0.0, "phi": 0.05, "nu": 0.0},
                                    {"sigma.sq": 0.1, "tau.sq": 0.1, "phi": 0.1, "nu": 0.03}])
def test_matern_chain_with_predrawn_normals_against_reference(glue, ref, port, tuning):
    """The double-buffered sampler takes the next iteration's draws while the kernel runs (nngp_abi.cu: Predrawn).  The
    stream must stay the reference's: Matern chain, 120 iterations, with and without zero proposal scales (rnorm(mu, 0)
    returns mu without consuming a normal, nmath/rnorm.c), against the reference's own rNNGP and against the same
    engine with the pre-draw switched off; the RNG ends in the same state (the next uniform agrees)."""
    rng = np.random.default_rng(5)
    n, m = 900, 10
    coords = rng.random((n, 2)); coords = coords[np.argsort(coords[:, 0], kind="stable")]
    X = np.column_stack([np.ones(n), rng.standard_normal(n)])
    y = X @ np.array([1.0, 5.0]) + np.sin(4 * coords[:, 0]) + 0.5 * rng.standard_normal(n)
    idx, _, lu = port.mk_nn_indx_cb(coords, m)
    starting = {"beta": [0.0, 0.0], "sigma.sq": 2.0, "tau.sq": 0.5, "phi": 6.0, "nu": 0.8}
    args = (y, X, coords, 2, idx, lu, m, [2.0, 2.0], [2.0, 0.5], [3.0, 30.0], [0.3, 1.6], starting, tuning, 120)
    ref.set_seed(11)
    want = ref.rnngp(*args)
    u_ref = ref.runif() if hasattr(ref, "runif") else None
    got = {}
    for pre in ("on", "off"):
        if pre == "off":
            os.environ["NNGP_NO_PREDRAW"] = "1"
        try:
            glue.lib.rshim_set_seed(11)
            got[pre] = po.call_rnngp(glue, *args)
            got[pre + "_u"] = float(glue.lib.unif_rand()) if hasattr(glue.lib, "unif_rand") else None
        finally:
            os.environ.pop("NNGP_NO_PREDRAW", None)
    assert np.array_equal(got["on"]["p_theta_samples"], got["off"]["p_theta_samples"])
    assert np.array_equal(got["on"]["p_beta_samples"], got["off"]["p_beta_samples"])
    assert got["on_u"] == got["off_u"]  # same RNG state after the run
    assert int((np.diff(want["p_theta_samples"][0]) != 0).sum()) > 5
    assert np.allclose(got["on"]["p_theta_samples"], want["p_theta_samples"], rtol=1e-9, atol=0)
    assert np.allclose(got["on"]["p_beta_samples"], want["p_beta_samples"], rtol=1e-7, atol=1e-9)


def test_interrupt_unwinds_and_leaves_the_engine_usable(glue):
    """A user interrupt during `rNNGP`: the hook reports it (R_ToplevelExec contains R's jump), the engine returns
    NNGP_EINTR through its normal return path, the glue raises the R error after PutRNGstate -- and the next call works."""
    d = np.load(os.path.join(G, "chain_n300_m10.npz"))
    starting = {"beta": [0.0, 0.0], "sigma.sq": 5.0, "tau.sq": 1.0, "phi": 6.0}
    tuning = {"sigma.sq": 0.05, "tau.sq": 0.05, "phi": 0.05}
    args = (d["y"], d["X"], d["coords"], 0, d["nnIndx"], d["nnIndxLU"], int(d["m"]), [2.0, 5.0], [2.0, 1.0], [3.0, 300.0], None,
            starting, tuning, 40)
    glue.lib.rshim_set_seed(1)
    glue.lib.rshim_raise_interrupt()
    with pytest.raises(RuntimeError, match="interrupted"):
        po.call_rnngp(glue, *args)
    glue.reset()
    glue.lib.rshim_set_seed(1)
    out = po.call_rnngp(glue, *args)
    assert np.allclose(out["p_theta_samples"], d["exp_theta"], rtol=1e-9)
