"""Shared synthetic-data helpers for the tests (numpy default_rng, seeds recorded by the callers)."""
import numpy as np


def make_problem(n, m, seed=0, p=2, domain=1.0, port=None, search="cb", sort=True):
    """coords ~ U(0,domain)^2 ordered by x (the reference's default `ord`, R/spNNGP.R:134),
    X = [1, N(0,1)...], y = X beta + smooth field + noise.  Neighbours from the CPU oracle."""
    rng = np.random.default_rng(seed)
    coords = rng.random((n, 2)) * domain
    if sort:
        coords = coords[np.argsort(coords[:, 0], kind="stable")]
    X = np.column_stack([np.ones(n)] + [rng.standard_normal(n) for _ in range(p - 1)])
    beta = np.array([1.0, 5.0, -2.0, 0.5][:p])
    w = np.sin(3 * coords[:, 0]) + np.cos(2 * coords[:, 1])
    y = X @ beta + np.sqrt(5.0) * w + rng.standard_normal(n)
    out = dict(n=n, m=m, p=p, coords=coords, X=X, y=y, beta=beta)
    if port is not None:
        fn = port.mk_nn_indx_cb if search == "cb" else port.mk_nn_indx
        nnIndx, nnDist, lu = fn(coords, m)
        out.update(nnIndx=nnIndx, nnDist=nnDist, nnIndxLU=lu)
    return out


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(b), 1e-300)
    return float(np.max(np.abs(a - b) / den)) if a.size else 0.0


def relerr_norm(a, b):
    """max |a-b| / max |b| (for vectors whose entries can legitimately be ~0, like B)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)) if a.size else 0.0
