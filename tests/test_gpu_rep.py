"""GPU: replicated-data draws -- fit.rep inside `rNNGP` and the `rNNGPReplicated` entry point --
against the reference's committed outputs, the live reference where built, and the CPU oracle."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from tests.util import make_problem

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def glue():
    lib = C.CDLL(os.path.join(ROOT, "spnngp_b200", "libspNNGP_b200.so"))
    return po._Sexp(lib)


def test_rep_draw_matches_oracle_recurrence(engine, port):
    """nngp_rep_draw on the B, F of a slot vs the serial recurrence v_i = B_i.v_N(i) + sqrt(F_i) z_i (src/rNNGP.cpp:24-42)."""
    for n, m, cov, theta in ((3000, 15, 0, [5, 1, 6]), (2500, 5, 3, [2, 0.5, 9]), (1500, 20, 2, [5, 1, 6, 0.75]), (40, 15, 0, [5, 1, 6])):
        prob = make_problem(n, m, seed=31 + m, port=port, domain=np.sqrt(n / 5000.0))
        engine.set_data(prob["coords"], prob["nnIndx"], prob["nnIndxLU"], prob["X"], prob["y"], m)
        engine.bf_loglik(cov, [theta], [prob["beta"]], slots=[1])
        B, F = engine.get_bf(1)
        rng = np.random.default_rng(5)
        for rep in range(2):  # second call reuses the level structure
            z = rng.standard_normal(n)
            got = engine.rep_draw(1, prob["beta"], z)
            lu = prob["nnIndxLU"]; idx = prob["nnIndx"]
            v = np.zeros(n)
            for i in range(n):
                o, c = lu[i], lu[n + i]
                v[i] = B[o:o + c] @ v[idx[o:o + c]] + np.sqrt(F[i]) * z[i]
            want = prob["X"] @ prob["beta"] + v
            assert np.allclose(got, want, rtol=1e-11, atol=1e-11)


def test_rep_rejects_a_shard(engine, port):
    prob = make_problem(500, 8, seed=3, port=port)
    engine.set_data(prob["coords"], prob["nnIndx"], prob["nnIndxLU"], prob["X"], prob["y"], 8, loc_range=(100, 500))
    engine.bf_loglik(0, [[5, 1, 6]], [prob["beta"]], slots=[0])
    with pytest.raises(RuntimeError, match="whole location range"):
        engine.rep_draw(0, prob["beta"], np.zeros(500))


def test_glue_replicates_against_golden(glue):
    """`rNNGP` with nRep > 0 and `rNNGPReplicated` through the .Call layer, R-compatible RNG stream, both sampler modes."""
    d = np.load(os.path.join(G, "rep_n300_m10.npz"))
    starting = {"beta": [0.0, 0.0], "sigma.sq": 5.0, "tau.sq": 1.0, "phi": 6.0, "nu": 0.75}
    tuning = {"sigma.sq": 0.05, "tau.sq": 0.05, "phi": 0.05, "nu": 0.05}
    for cov, name in ((0, "exp"), (2, "matern")):
        for mode in ("0", "1"):
            os.environ["NNGP_SAMPLER_MODE"] = mode
            glue.lib.rshim_set_seed(3)
            out = po.call_rnngp(glue, d["y"], d["X"], d["coords"], cov, d["nnIndx"], d["nnIndxLU"], int(d["m"]), [2.0, 5.0],
                                [2.0, 1.0], [3.0, 300.0], [0.1, 2.0] if cov == 2 else None, starting, tuning, 30,
                                rep_indx=d["rep_indx"])
            assert np.allclose(out["p_theta_samples"], d[f"{name}_theta"], rtol=1e-9)
            assert out["y_rep_samples"].shape == d[f"{name}_yrep"].shape
            assert np.allclose(out["y_rep_samples"], d[f"{name}_yrep"], rtol=1e-8, atol=1e-8)
        os.environ.pop("NNGP_SAMPLER_MODE", None)
        glue.lib.rshim_set_seed(4)
        r = po.call_rnngp_replicated(glue, d["X"], d["coords"], cov, d["nnIndx"], d["nnIndxLU"], int(d["m"]),
                                     d[f"{name}_replicated_beta"], d[f"{name}_replicated_theta"])
        assert np.allclose(r, d[f"{name}_replicated_yrep_seed4"], rtol=1e-9, atol=1e-10)


def test_glue_replicates_vs_live_reference(glue, ref, port):
    prob = make_problem(1500, 15, seed=18, port=port, domain=0.55)
    rng = np.random.default_rng(2)
    S = 5
    bsm = np.array([1.0, 5.0])[None, :] + 0.1 * rng.standard_normal((S, 2))
    tsm = np.tile(np.array([5.0, 1.0, 6.0])[:, None], (1, S)); tsm[:, 2:] *= 1.05  # runs of equal theta (rejected proposals)
    args = (prob["X"], prob["coords"], 0, prob["nnIndx"], prob["nnIndxLU"], 15, bsm, tsm)
    ref.set_seed(9); a = ref.rnngp_replicated(*args)
    glue.lib.rshim_set_seed(9); b = po.call_rnngp_replicated(glue, *args)
    assert a.shape == (1500, S) and np.allclose(a, b, rtol=1e-10, atol=1e-10)
