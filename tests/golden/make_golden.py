"""Generate tests/golden/*.npz from the reference's OWN C++ (oracle/_ref, built from /root/reference by
oracle/Makefile).  Run in the dev container only:  python tests/golden/make_golden.py

The reference ships no tests or golden vectors; these fixtures pin the CPU oracle port
(tests/test_oracle_golden.py) and, on the GPU box where /root/reference does not exist, the CUDA path.
Every input is generated from a recorded numpy seed; outputs are the reference's, run single-threaded.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from tests.util import make_problem  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
ref = po.Ref()

THETA = {0: [5.0, 1.0, 6.0], 1: [5.0, 1.0, 6.0], 2: [5.0, 1.0, 6.0, 0.75], 3: [5.0, 1.0, 6.0]}


def grid_sites(k):
    g = np.arange(k) / k
    xx, yy = np.meshgrid(g, g, indexing="ij")
    return np.column_stack([xx.ravel(), yy.ravel()])


# ---- 7. k-fold scores of the conjugate model: the reference's mkNNIndxCB + cNNGP per fold, the
#         R-layer loop and crps / rmspe restated in oracle/pyoracle.py (R/spConjNNGP.R:286-349, R/util.R:247-255)
def kfold_fixture():
    port = po.Port()
    n, m, k_fold = 600, 10, 3
    prob = make_problem(n, m, seed=20248, domain=0.35)
    rng = np.random.default_rng(20249)
    n_ho = n // k_fold
    ho = np.stack([np.sort(rng.choice(n, n_ho, replace=False)) for _ in range(k_fold)]).astype(np.int32)
    kf = dict(coords=prob["coords"], X=prob["X"], y=prob["y"], m=m, ho_idx=ho, sigmaSqIG=np.array([2.0, float(np.var(prob["y"]))]))
    for cov, name in ((0, "exp"), (2, "matern")):
        rows = [[ph, al] + ([0.8] if cov == 2 else []) for ph in (3.0, 9.0) for al in (0.1, 0.4)]
        ta = np.array(rows).T
        kf[f"{name}_theta_alpha"] = ta
        kf[f"{name}_scores"] = po.conj_kfold(ref, port, prob["y"], prob["X"], prob["coords"], m, cov, ta, kf["sigmaSqIG"], ho, "cb")
    np.savez_compressed(os.path.join(OUT, "kfold_n600_m10.npz"), **kf)


if "--kfold-only" in sys.argv:
    kfold_fixture()
    sys.exit(0)

# ---- 1. B, F, logDet, Q for the four covariance models -----------------------------------------------
n, m = 400, 10
prob = make_problem(n, m, seed=20241, domain=0.3)
nnIndx, nnDist, lu = ref.mk_nn_indx_cb(prob["coords"], m)
bf = dict(coords=prob["coords"], X=prob["X"], y=prob["y"], beta=prob["beta"], m=m, nnIndx=nnIndx, nnDist=nnDist, nnIndxLU=lu)
r = prob["X"] @ prob["beta"] - prob["y"]
for cov, th in THETA.items():
    B, F, ld = ref.update_bf(prob["coords"], nnIndx, lu, n, m, th, cov)
    bf[f"B{cov}"] = B; bf[f"F{cov}"] = F; bf[f"logDet{cov}"] = ld
    bf[f"Q{cov}"] = ref.q(B, F, r, r, n, nnIndx, lu)
    bf[f"theta{cov}"] = np.array(th)
np.savez_compressed(os.path.join(OUT, "bf_n400_m10.npz"), **bf)

# ---- 2. neighbour search: random sites and a tied grid, brute and code-book ---------------------------
knn = {}
rng = np.random.default_rng(20242)
c = rng.random((500, 2)); c = c[np.argsort(c[:, 0], kind="stable")]
knn["rand_coords"] = c
for mm in (5, 15):
    a = ref.mk_nn_indx(c, mm); b = ref.mk_nn_indx_cb(c, mm)
    knn[f"rand_brute_m{mm}_idx"] = a[0]; knn[f"rand_brute_m{mm}_dist"] = a[1]
    knn[f"rand_cb_m{mm}_idx"] = b[0]; knn[f"rand_cb_m{mm}_dist"] = b[1]
g = grid_sites(20)
knn["grid_coords"] = g
for mm in (8, 15, 32):
    a = ref.mk_nn_indx(g, mm); b = ref.mk_nn_indx_cb(g, mm)
    knn[f"grid_brute_m{mm}_idx"] = a[0]; knn[f"grid_brute_m{mm}_dist"] = a[1]
    knn[f"grid_cb_m{mm}_idx"] = b[0]; knn[f"grid_cb_m{mm}_dist"] = b[1]
np.savez_compressed(os.path.join(OUT, "knn.npz"), **knn)

# ---- 3. the response-model sampler: seeded chains (shim RNG = R's MT19937 + inversion) -----------------
chain = {}
n, m = 300, 10
prob = make_problem(n, m, seed=20243, domain=0.25)
nnIndx, _, lu = ref.mk_nn_indx_cb(prob["coords"], m)
chain.update(coords=prob["coords"], X=prob["X"], y=prob["y"], m=m, nnIndx=nnIndx, nnIndxLU=lu)
pri = dict(sigmaSqIG=[2.0, 5.0], tauSqIG=[2.0, 1.0], phiUnif=[3.0, 300.0])
for cov, name in ((0, "exp"), (2, "matern")):
    starting = {"beta": [0.0, 0.0], "sigma.sq": 5.0, "tau.sq": 1.0, "phi": 6.0, "nu": 0.75}
    tuning = {"sigma.sq": 0.05, "tau.sq": 0.05, "phi": 0.05, "nu": 0.05}
    ref.set_seed(1)
    out = ref.rnngp(prob["y"], prob["X"], prob["coords"], cov, nnIndx, lu, m, pri["sigmaSqIG"], pri["tauSqIG"],
                    pri["phiUnif"], [0.1, 2.0] if cov == 2 else None, starting, tuning, 40)
    chain[f"{name}_beta"] = out["p_beta_samples"]; chain[f"{name}_theta"] = out["p_theta_samples"]
np.savez_compressed(os.path.join(OUT, "chain_n300_m10.npz"), **chain)

# ---- 4. conjugate model and prediction -------------------------------------------------------------------
cp = {}
n, m, n0 = 500, 10, 120
prob = make_problem(n, m, seed=20244, domain=0.3)
nnIndx, _, lu = ref.mk_nn_indx_cb(prob["coords"], m)
rng = np.random.default_rng(20245)
coords0 = rng.random((n0, 2)) * 0.3
X0 = np.column_stack([np.ones(n0), rng.standard_normal(n0)])
nn0 = po.Port().nn_query(prob["coords"], coords0, m)  # exhaustive exact kNN (stand-in for RANN::nn2)
ta = np.array([[3.0, 0.1], [6.0, 0.5], [12.0, 1.0]]).T
ig = [2.0, float(np.var(prob["y"]))]
cp.update(coords=prob["coords"], X=prob["X"], y=prob["y"], m=m, nnIndx=nnIndx, nnIndxLU=lu, coords0=coords0, X0=X0,
          nn0=nn0, thetaAlpha=ta, sigmaSqIG=np.array(ig))
for cov, name in ((0, "exp"), (3, "gauss")):
    out = ref.cnngp(prob["y"], prob["X"], prob["coords"], ta, nnIndx, lu, X0, coords0, nn0, m, ig, cov)
    for k, v in out.items():
        cp[f"conj_{name}_{k.replace('.', '_')}"] = v
S = 12
beta = np.array([1.0, 5.0])[:, None] + 0.1 * rng.standard_normal((2, S))
theta = np.tile(np.array([5.0, 1.0, 6.0])[:, None], (1, S)); theta[:, 5:] *= 1.07; theta[:, 9:] *= 0.9
ref.set_seed(7)
cp["pred_beta"] = beta; cp["pred_theta"] = theta
cp["pred_y0_seed7"] = ref.predict(prob["X"], prob["y"], prob["coords"], m, X0, coords0, nn0, beta, theta, 0)
np.savez_compressed(os.path.join(OUT, "conj_predict_n500_m10.npz"), **cp)
# ---- 5. fit.rep / rNNGPReplicated: replicates from the reference, seeded ---------------
rp = {}
n, m = 300, 10
prob = make_problem(n, m, seed=20246, domain=0.25)
nnIndx, _, lu = ref.mk_nn_indx_cb(prob["coords"], m)
rp.update(coords=prob["coords"], X=prob["X"], y=prob["y"], m=m, nnIndx=nnIndx, nnIndxLU=lu)
starting = {"beta": [0.0, 0.0], "sigma.sq": 5.0, "tau.sq": 1.0, "phi": 6.0, "nu": 0.75}
tuning = {"sigma.sq": 0.05, "tau.sq": 0.05, "phi": 0.05, "nu": 0.05}
rep_indx = np.zeros(30, dtype=np.int32); rep_indx[10::4] = 1  # as R/spNNGP.R marks the sub.sample iterations
rp["rep_indx"] = rep_indx
for cov, name in ((0, "exp"), (2, "matern")):
    ref.set_seed(3)
    out = ref.rnngp(prob["y"], prob["X"], prob["coords"], cov, nnIndx, lu, m, pri["sigmaSqIG"], pri["tauSqIG"],
                    pri["phiUnif"], [0.1, 2.0] if cov == 2 else None, starting, tuning, 30, rep_indx=rep_indx)
    rp[f"{name}_beta"] = out["p_beta_samples"]; rp[f"{name}_theta"] = out["p_theta_samples"]
    rp[f"{name}_yrep"] = out["y_rep_samples"]
    ref.set_seed(4)
    S = 6
    bsm = out["p_beta_samples"][:, -S:].T.copy(); tsm = out["p_theta_samples"][:, -S:].copy()
    rp[f"{name}_replicated_beta"] = bsm; rp[f"{name}_replicated_theta"] = tsm
    rp[f"{name}_replicated_yrep_seed4"] = ref.rnngp_replicated(prob["X"], prob["coords"], cov, nnIndx, lu, m, bsm, tsm)
np.savez_compressed(os.path.join(OUT, "rep_n300_m10.npz"), **rp)

# ---- 6. mkUIndx: transposed neighbour graph from the reference, all three search types ----
ui = {}
n, m = 600, 10
prob = make_problem(n, m, seed=20247, domain=0.3)
nnIndx, _, lu = ref.mk_nn_indx_cb(prob["coords"], m)
ui.update(n=n, m=m, nnIndx=nnIndx, nnIndxLU=lu)
for st in (0, 1, 2):
    u, ulu, uii = ref.mk_u_indx(n, m, nnIndx, lu, search_type=st)
    ui[f"st{st}_uIndx"] = u; ui[f"st{st}_uIndxLU"] = ulu; ui[f"st{st}_uiIndx"] = uii
np.savez_compressed(os.path.join(OUT, "uindx_n600_m10.npz"), **ui)

kfold_fixture()

print("golden fixtures written to", OUT, [f for f in os.listdir(OUT) if f.endswith(".npz")])
