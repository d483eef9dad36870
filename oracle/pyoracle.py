"""oracle/pyoracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes bindings for the two CPU checkers:

* ``Port``  -> oracle/liboracle_port.so, our plain-C restatement (oracle/nngp_oracle.c)
* ``Ref``   -> oracle/_ref/libspnngp_ref.so, the reference's own unmodified C++ compiled against
               the rshim stand-in for libR (oracle/Makefile); internals are reached through their
               C++-mangled names, the ``.Call`` entry points through ``rshim_call``.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
All matrices are column-major (R layout): coords is passed as a flat [x..., y...] vector.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "liboracle_port.so")
REF_SO = os.path.join(_HERE, "_ref", "libspnngp_ref.so")

EXPONENTIAL, SPHERICAL, MATERN, GAUSSIAN = 0, 1, 2, 3
COV_CODES = {"exponential": 0, "spherical": 1, "matern": 2, "gaussian": 3}

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
RNORM_FN = C.CFUNCTYPE(C.c_double, C.c_double, C.c_double)
RUNIF_FN = C.CFUNCTYPE(C.c_double, C.c_double, C.c_double)


def build(quiet: bool = True) -> None:
    """(Re)build the checkers. `_ref` is only rebuilt where /root/reference exists."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def dp(a):
    return a.ctypes.data_as(c_double_p)


def ip(a):
    return a.ctypes.data_as(c_int_p)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def coords_flat(coords):
    """n x 2 array -> column-major flat vector [x..., y...] (R's layout for a matrix)."""
    c = np.asarray(coords, dtype=np.float64)
    if c.ndim == 2:
        return np.ascontiguousarray(c.T).ravel()
    return np.ascontiguousarray(c)


def n_indx(n, m):
    """src/rNNGP.cpp:218"""
    return int((1 + m) / 2 * m + (n - m - 1) * m)


class Port:
    """Plain-C restatement."""

    def __init__(self, path: str = PORT_SO):
        if not os.path.exists(path):
            build()
        self.lib = L = C.CDLL(path)
        L.orc_update_bf.restype = C.c_double
        L.orc_q.restype = C.c_double
        L.orc_sp_cor.restype = C.c_double
        L.orc_sp_cor.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int]
        L.orc_bessel_k.restype = C.c_double
        L.orc_bessel_k.argtypes = [C.c_double, C.c_double]
        L.orc_dist2.restype = C.c_double
        L.orc_dist2.argtypes = [C.c_double] * 4
        L.orc_n_indx.restype = C.c_long
        L.orc_rnngp.argtypes = None
        L.orc_max_threads.restype = C.c_int

    def set_threads(self, n):
        self.lib.orc_set_threads(C.c_int(int(n)))

    def max_threads(self):
        return int(self.lib.orc_max_threads())

    def sp_cor(self, D, phi, nu, cov):
        return float(self.lib.orc_sp_cor(D, phi, nu, cov))

    def bessel_k(self, x, nu):
        return float(self.lib.orc_bessel_k(x, nu))

    def rsort_with_index(self, x, idx):
        x = f64(x).copy(); idx = i32(idx).copy()
        self.lib.orc_rsort_with_index(dp(x), ip(idx), C.c_int(len(x)))
        return x, idx

    def update_bf(self, coords, nnIndx, nnIndxLU, n, m, theta, cov):
        coords = coords_flat(coords); nnIndx = i32(nnIndx); nnIndxLU = i32(nnIndxLU)
        th = np.zeros(4); th[: len(theta)] = theta
        B = np.zeros(max(n_indx(n, m), 1)); F = np.zeros(n)
        info = C.c_int(0)
        ld = self.lib.orc_update_bf(dp(B), dp(F), dp(coords), ip(nnIndx), ip(nnIndxLU), C.c_int(n), C.c_int(m),
                                    dp(th), C.c_int(cov), C.byref(info))
        return B, F, float(ld), int(info.value)

    def update_conj_bf(self, coords, nnIndx, nnIndxLU, n, m, phi, alpha, nu, cov):
        coords = coords_flat(coords); nnIndx = i32(nnIndx); nnIndxLU = i32(nnIndxLU)
        B = np.zeros(max(n_indx(n, m), 1)); F = np.zeros(n)
        info = C.c_int(0)
        self.lib.orc_update_conj_bf(dp(B), dp(F), dp(coords), ip(nnIndx), ip(nnIndxLU), C.c_int(n), C.c_int(m),
                                    C.c_double(phi), C.c_double(alpha), C.c_double(nu), C.c_int(cov), C.byref(info))
        return B, F, int(info.value)

    def q(self, B, F, u, v, n, nnIndx, nnIndxLU):
        B = f64(B); F = f64(F); u = f64(u); v = f64(v); nnIndx = i32(nnIndx); nnIndxLU = i32(nnIndxLU)
        return float(self.lib.orc_q(dp(B), dp(F), dp(u), dp(v), C.c_int(n), ip(nnIndx), ip(nnIndxLU)))

    def _mk(self, fn, coords, m):
        coords = coords_flat(coords); n = len(coords) // 2
        nI = max(n_indx(n, m), 1)
        nnIndx = np.zeros(nI, dtype=np.int32); nnDist = np.zeros(nI); lu = np.zeros(2 * n, dtype=np.int32)
        fn(C.c_int(n), C.c_int(m), dp(coords), ip(nnIndx), dp(nnDist), ip(lu))
        return nnIndx, nnDist, lu

    def mk_nn_indx(self, coords, m):
        return self._mk(self.lib.orc_mk_nn_indx, coords, m)

    def mk_u_indx(self, n, m, nnIndx, nnIndxLU, search_type=2):
        nnIndx = i32(nnIndx); nnIndxLU = i32(nnIndxLU); nI = max(n_indx(n, m), 1)
        u = np.zeros(nI, dtype=np.int32); ui = np.zeros(nI, dtype=np.int32); lu = np.zeros(2 * n, dtype=np.int32)
        self.lib.orc_mk_u_indx(C.c_int(n), C.c_int(m), ip(nnIndx), ip(nnIndxLU), C.c_int(search_type), ip(u), ip(lu), ip(ui))
        return u, lu, ui

    def mk_nn_indx_cb(self, coords, m):
        return self._mk(self.lib.orc_mk_nn_indx_cb, coords, m)

    def nn_query(self, coords, coords0, m):
        coords = coords_flat(coords); coords0 = coords_flat(coords0)
        n, q = len(coords) // 2, len(coords0) // 2
        out = np.zeros(q * m, dtype=np.int32)
        self.lib.orc_nn_query(C.c_int(n), dp(coords), C.c_int(q), dp(coords0), C.c_int(m), ip(out))
        return out  # column-major q x m

    def rnngp(self, y, X, coords, cov, nnIndx, nnIndxLU, m, sigmaSqIG, tauSqIG, phiUnif, nuUnif, starting, tuning,
              nSamples, rnorm, runif, rep_indx=None):
        """starting/tuning: dict with beta (starting only), sigma.sq, tau.sq, phi, nu.
        rep_indx: 0/1 per iteration (fit.rep); then y.rep (n x nRep) is returned as a 4th item."""
        y = f64(y); n = len(y)
        X = np.asfortranarray(np.asarray(X, dtype=np.float64).reshape(n, -1)); p = X.shape[1]
        Xf = np.ascontiguousarray(X.T).ravel()
        coords = coords_flat(coords); nnIndx = i32(nnIndx); nnIndxLU = i32(nnIndxLU)
        nTheta = 4 if cov == MATERN else 3
        bs = np.zeros(p * nSamples); ts = np.zeros(nTheta * nSamples)
        beta0 = f64(starting["beta"])
        self._cb = (RNORM_FN(rnorm), RUNIF_FN(runif))
        self.lib.orc_rnngp_rep.restype = C.c_int
        ri = rep = None
        if rep_indx is not None:
            ri = i32(rep_indx); k = int((ri != 0).sum()); rep = np.zeros(max(k * n, 1))
        acc = self.lib.orc_rnngp_rep(
            dp(y), dp(Xf), C.c_int(p), C.c_int(n), C.c_int(m), dp(coords), C.c_int(cov), ip(nnIndx), ip(nnIndxLU),
            dp(f64(sigmaSqIG)), dp(f64(tauSqIG)), dp(f64(phiUnif)), dp(f64(nuUnif if nuUnif is not None else [0, 0])),
            dp(beta0), C.c_double(starting["sigma.sq"]), C.c_double(starting["tau.sq"]), C.c_double(starting["phi"]),
            C.c_double(starting.get("nu", 0.0)), C.c_double(tuning["sigma.sq"]), C.c_double(tuning["tau.sq"]),
            C.c_double(tuning["phi"]), C.c_double(tuning.get("nu", 0.0)), C.c_int(nSamples), self._cb[0], self._cb[1],
            dp(bs), dp(ts), ip(ri) if ri is not None else None, dp(rep) if rep is not None else None)
        res = (bs.reshape(nSamples, p).T.copy(), ts.reshape(nSamples, nTheta).T.copy(), int(acc))
        if rep is not None:
            res = res + (rep[: k * n].reshape(k, n).T.copy(),)
        return res

    def rnngp_replicated(self, X, coords, cov, nnIndx, nnIndxLU, m, beta_samples, theta_samples, rnorm):
        """src/rNNGPReplicated.cpp:183-201.  beta_samples S x p, theta_samples nTheta x S; returns n x S."""
        bsm = np.asarray(beta_samples, dtype=np.float64); S, p = bsm.shape
        coords = coords_flat(coords); n = coords.size // 2
        X = np.asarray(X, dtype=np.float64).reshape(n, p)
        Xf = np.ascontiguousarray(X.T).ravel(); bf = np.ascontiguousarray(bsm.T).ravel()
        tf = np.ascontiguousarray(np.asarray(theta_samples, dtype=np.float64).T).ravel()
        nnIndx = i32(nnIndx); nnIndxLU = i32(nnIndxLU)
        out = np.zeros(max(n * S, 1))
        self._cb = (RNORM_FN(rnorm),)
        self.lib.orc_rnngp_replicated.restype = C.c_int
        rc = self.lib.orc_rnngp_replicated(dp(Xf), C.c_int(p), C.c_int(n), C.c_int(m), dp(coords), C.c_int(cov), ip(nnIndx),
                                           ip(nnIndxLU), dp(bf), dp(tf), C.c_int(S), self._cb[0], dp(out))
        if rc:
            raise RuntimeError("orc_rnngp_replicated: factorisation failed")
        return out[: n * S].reshape(S, n).T.copy()

    def cnngp(self, y, X, coords, thetaAlpha, nnIndx, nnIndxLU, X0, coords0, nnIndx0, m, sigmaSqIG, cov):
        y = f64(y); n = len(y)
        X = np.asarray(X, dtype=np.float64).reshape(n, -1); p = X.shape[1]
        Xf = np.ascontiguousarray(X.T).ravel()
        coords = coords_flat(coords)
        ta = np.asarray(thetaAlpha, dtype=np.float64)  # nTheta x g
        nTheta, g = ta.shape
        taf = np.ascontiguousarray(ta.T).ravel()
        if X0 is None:
            n0 = 0; X0f = np.zeros(1); c0 = np.zeros(2); nn0 = np.zeros(1, dtype=np.int32)
        else:
            X0 = np.asarray(X0, dtype=np.float64); n0 = X0.shape[0]
            X0f = np.ascontiguousarray(X0.reshape(n0, p).T).ravel(); c0 = coords_flat(coords0); nn0 = i32(nnIndx0)
        betaHat = np.zeros(p * g); ab = np.zeros(nTheta * g); bBInv = np.zeros(p * p * g); bb = np.zeros(p * g)
        y0 = np.zeros(max(n0 * g, 1)); y0v = np.zeros(max(n0 * g, 1))
        st = self.lib.orc_cnngp(dp(y), dp(Xf), C.c_int(p), C.c_int(n), dp(coords), dp(taf), ip(i32(nnIndx)),
                                ip(i32(nnIndxLU)), dp(X0f), dp(c0), C.c_int(n0), ip(nn0), C.c_int(g), C.c_int(m),
                                dp(f64(sigmaSqIG)), C.c_int(cov), dp(betaHat), dp(ab), dp(bBInv), dp(bb), dp(y0), dp(y0v))
        return dict(status=int(st), beta_hat=betaHat.reshape(g, p).T.copy(), ab=ab, bB_inv=bBInv.reshape(g, p * p).T.copy(),
                    bb=bb.reshape(g, p).T.copy(), y0_hat=y0[: n0 * g].reshape(g, n0).T.copy(),
                    y0_hat_var=y0v[: n0 * g].reshape(g, n0).T.copy())

    def predict(self, X, y, coords, m, X0, coords0, nnIndx0, betaSamples, thetaSamples, cov, z, site_range=None):
        y = f64(y); n = len(y)
        X = np.asarray(X, dtype=np.float64).reshape(n, -1); p = X.shape[1]
        Xf = np.ascontiguousarray(X.T).ravel()
        X0 = np.asarray(X0, dtype=np.float64); q = X0.shape[0]
        X0f = np.ascontiguousarray(X0.reshape(q, p).T).ravel()
        bs = np.asarray(betaSamples, dtype=np.float64); S = bs.shape[1]
        bsf = np.ascontiguousarray(bs.T).ravel()
        tsf = np.ascontiguousarray(np.asarray(thetaSamples, dtype=np.float64).T).ravel()
        y0 = np.zeros(q * S)
        lo, hi = site_range if site_range else (0, q)
        st = self.lib.orc_rnngp_predict(dp(Xf), dp(y), dp(coords_flat(coords)), C.c_int(n), C.c_int(p), C.c_int(m),
                                        dp(X0f), dp(coords_flat(coords0)), C.c_int(q), ip(i32(nnIndx0)), dp(bsf), dp(tsf),
                                        C.c_int(S), C.c_int(cov), dp(f64(z)), dp(y0), C.c_int(lo), C.c_int(hi))
        return y0.reshape(S, q).T.copy(), int(st)


class _Sexp:
    """Helpers around the rshim SEXP model living inside a shim-linked library."""

    def __init__(self, lib):
        self.lib = lib
        vp = C.c_void_p
        lib.rshim_wrap_real.restype = vp; lib.rshim_wrap_real.argtypes = [c_double_p, C.c_ssize_t]
        lib.rshim_wrap_int.restype = vp; lib.rshim_wrap_int.argtypes = [c_int_p, C.c_ssize_t]
        lib.rshim_wrap_real_matrix.restype = vp; lib.rshim_wrap_real_matrix.argtypes = [c_double_p, C.c_int, C.c_int]
        lib.rshim_wrap_int_matrix.restype = vp; lib.rshim_wrap_int_matrix.argtypes = [c_int_p, C.c_int, C.c_int]
        lib.rshim_list_get.restype = vp; lib.rshim_list_get.argtypes = [vp, C.c_char_p]
        lib.rshim_call.restype = C.c_int; lib.rshim_call.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(vp)]
        lib.rshim_last_error.restype = C.c_char_p
        lib.REAL.restype = c_double_p; lib.REAL.argtypes = [vp]
        lib.INTEGER.restype = c_int_p; lib.INTEGER.argtypes = [vp]
        lib.Rf_xlength.restype = C.c_ssize_t; lib.Rf_xlength.argtypes = [vp]
        lib.Rf_nrows.restype = C.c_int; lib.Rf_nrows.argtypes = [vp]
        lib.Rf_ncols.restype = C.c_int; lib.Rf_ncols.argtypes = [vp]
        lib.rshim_set_seed.argtypes = [C.c_uint]
        lib.unif_rand.restype = C.c_double
        lib.norm_rand.restype = C.c_double
        lib.rnorm.restype = C.c_double; lib.rnorm.argtypes = [C.c_double, C.c_double]
        lib.runif.restype = C.c_double; lib.runif.argtypes = [C.c_double, C.c_double]
        self._keep = []

    def real(self, a):
        a = f64(a); self._keep.append(a)
        return self.lib.rshim_wrap_real(dp(a), a.size)

    def integer(self, a):
        a = i32(a); self._keep.append(a)
        return self.lib.rshim_wrap_int(ip(a), a.size)

    def scalar_int(self, v):
        return self.integer(np.array([v], dtype=np.int32))

    def scalar_real(self, v):
        return self.real(np.array([v], dtype=np.float64))

    def call(self, name, args):
        fn = C.cast(getattr(self.lib, name), C.c_void_p)
        arr = (C.c_void_p * len(args))(*args)
        out = C.c_void_p()
        rc = self.lib.rshim_call(fn, len(args), arr, C.byref(out))
        if rc != 0:
            raise RuntimeError(self.lib.rshim_last_error().decode())
        return out

    def get_real_matrix(self, lst, name):
        s = self.lib.rshim_list_get(lst, name.encode())
        if not s:
            return None
        n = self.lib.Rf_xlength(s); nr = self.lib.Rf_nrows(s); nc = self.lib.Rf_ncols(s)
        a = np.ctypeslib.as_array(self.lib.REAL(s), shape=(n,)).copy()
        return a.reshape(nc, nr).T.copy() if nr * nc == n else a

    def reset(self):
        self.lib.rshim_reset()
        self._keep.clear()


def call_rnngp(sx: _Sexp, y, X, coords, cov, nnIndx, nnIndxLU, m, sigmaSqIG, tauSqIG, phiUnif, nuUnif, starting,
               tuning, nSamples, nThreads=1, verbose=0, nReport=100, symbol="rNNGP", rep_indx=None):
    """Drive a 28-argument `rNNGP` .Call entry point (reference signature, src/spNNGP.h:9-13).
    rep_indx: 0/1 per iteration (R/spNNGP.R builds it from fit.rep / sub.sample); adds y_rep_samples (n x nRep)."""
    y = f64(y); n = len(y)
    X = np.asarray(X, dtype=np.float64).reshape(n, -1); p = X.shape[1]
    Xf = np.ascontiguousarray(X.T).ravel()
    nu_unif = nuUnif if nuUnif is not None else [0.0, 0.0]
    args = [sx.real(y), sx.real(Xf), sx.scalar_int(p), sx.scalar_int(n), sx.scalar_int(m), sx.real(coords_flat(coords)),
            sx.scalar_int(cov), sx.integer(nnIndx), sx.integer(nnIndxLU),
            sx.real(sigmaSqIG), sx.real(tauSqIG), sx.real(phiUnif), sx.real(nu_unif),
            sx.real(starting["beta"]), sx.scalar_real(starting["sigma.sq"]), sx.scalar_real(starting["tau.sq"]),
            sx.scalar_real(starting["phi"]), sx.scalar_real(starting.get("nu", 0.0)),
            sx.scalar_real(tuning["sigma.sq"]), sx.scalar_real(tuning["tau.sq"]), sx.scalar_real(tuning["phi"]),
            sx.scalar_real(tuning.get("nu", 0.0)),
            sx.scalar_int(nSamples), sx.scalar_int(nThreads), sx.scalar_int(verbose), sx.scalar_int(nReport),
            sx.scalar_int(0 if rep_indx is None else int((i32(rep_indx) != 0).sum())),
            sx.integer(np.zeros(max(nSamples, 1), dtype=np.int32) if rep_indx is None else i32(rep_indx))]
    out = sx.call(symbol, args)
    res = dict(p_beta_samples=sx.get_real_matrix(out, "p.beta.samples"),
               p_theta_samples=sx.get_real_matrix(out, "p.theta.samples"))
    if rep_indx is not None and int((i32(rep_indx) != 0).sum()) > 0:
        res["y_rep_samples"] = sx.get_real_matrix(out, "y.rep.samples")
    sx.reset()
    return res


def call_rnngp_replicated(sx: _Sexp, X, coords, cov, nnIndx, nnIndxLU, m, beta_samples, theta_samples, nThreads=1,
                          verbose=0, nReport=100, symbol="rNNGPReplicated"):
    """Drive the 14-argument `rNNGPReplicated` .Call entry point (src/spNNGP.h, src/rNNGPReplicated.cpp:76-78).
    beta_samples S x p (R passes the coda matrix as is), theta_samples nTheta x S.  Returns y.rep.samples n x S."""
    bsm = np.asarray(beta_samples, dtype=np.float64); S, p = bsm.shape
    cf = coords_flat(coords); n = cf.size // 2
    X = np.asarray(X, dtype=np.float64).reshape(n, p)
    args = [sx.real(np.ascontiguousarray(X.T).ravel()), sx.scalar_int(p), sx.scalar_int(n), sx.scalar_int(m), sx.real(cf),
            sx.scalar_int(cov), sx.integer(nnIndx), sx.integer(nnIndxLU), sx.real(np.ascontiguousarray(bsm.T).ravel()),
            sx.real(np.ascontiguousarray(np.asarray(theta_samples, dtype=np.float64).T).ravel()), sx.scalar_int(S),
            sx.scalar_int(nThreads), sx.scalar_int(verbose), sx.scalar_int(nReport)]
    out = sx.call(symbol, args)
    res = sx.get_real_matrix(out, "y.rep.samples")
    sx.reset()
    return res


def call_cnngp(sx: _Sexp, y, X, coords, thetaAlpha, nnIndx, nnIndxLU, X0, coords0, nnIndx0, m, sigmaSqIG, cov,
               nThreads=1, verbose=0, symbol="cNNGP"):
    """Drive an 18-argument `cNNGP` .Call entry point (src/spNNGP.h:21-23)."""
    y = f64(y); n = len(y)
    X = np.asarray(X, dtype=np.float64).reshape(n, -1); p = X.shape[1]
    ta = np.asarray(thetaAlpha, dtype=np.float64); nTheta, g = ta.shape
    if X0 is None:
        n0 = 0; X0f = np.zeros(1); c0 = np.zeros(2); nn0 = np.zeros(1, dtype=np.int32)
    else:
        X0 = np.asarray(X0, dtype=np.float64); n0 = X0.shape[0]
        X0f = np.ascontiguousarray(X0.reshape(n0, p).T).ravel(); c0 = coords_flat(coords0); nn0 = i32(nnIndx0)
    args = [sx.real(y), sx.real(np.ascontiguousarray(X.T).ravel()), sx.scalar_int(p), sx.scalar_int(n),
            sx.real(coords_flat(coords)), sx.real(np.ascontiguousarray(ta.T).ravel()), sx.integer(nnIndx),
            sx.integer(nnIndxLU), sx.real(X0f), sx.real(c0), sx.scalar_int(n0), sx.integer(nn0), sx.scalar_int(g),
            sx.scalar_int(m), sx.real(sigmaSqIG), sx.scalar_int(cov), sx.scalar_int(nThreads), sx.scalar_int(verbose)]
    out = sx.call(symbol, args)
    res = {k: sx.get_real_matrix(out, k) for k in ("beta.hat", "ab", "bB.inv", "bb", "y.0.hat", "y.0.hat.var")}
    sx.reset()
    return res


def call_predict(sx: _Sexp, X, y, coords, m, X0, coords0, nnIndx0, betaSamples, thetaSamples, cov, nThreads=1,
                 verbose=0, nReport=1000000, symbol="rNNGPPredict"):
    """Drive a 17-argument `rNNGPPredict` .Call entry point (src/spNNGP.h:36-37)."""
    y = f64(y); n = len(y)
    X = np.asarray(X, dtype=np.float64).reshape(n, -1); p = X.shape[1]
    X0 = np.asarray(X0, dtype=np.float64); q = X0.shape[0]
    bs = np.asarray(betaSamples, dtype=np.float64); S = bs.shape[1]
    ts = np.asarray(thetaSamples, dtype=np.float64)
    args = [sx.real(np.ascontiguousarray(X.T).ravel()), sx.real(y), sx.real(coords_flat(coords)), sx.scalar_int(n),
            sx.scalar_int(p), sx.scalar_int(m), sx.real(np.ascontiguousarray(X0.reshape(q, p).T).ravel()),
            sx.real(coords_flat(coords0)), sx.scalar_int(q), sx.integer(nnIndx0),
            sx.real(np.ascontiguousarray(bs.T).ravel()), sx.real(np.ascontiguousarray(ts.T).ravel()), sx.scalar_int(S),
            sx.scalar_int(cov), sx.scalar_int(nThreads), sx.scalar_int(verbose), sx.scalar_int(nReport)]
    out = sx.call(symbol, args)
    res = sx.get_real_matrix(out, "p.y.0")
    sx.reset()
    return res


def call_mk_nn(sx: _Sexp, coords, m, nThreads=1, symbol="mkNNIndxCB"):
    """Drive `mkNNIndx` / `mkNNIndxCB` (src/nn.h:25,38): outputs are written into the arguments."""
    coords = coords_flat(coords); n = len(coords) // 2
    nI = max(n_indx(n, m), 1)
    nnIndx = np.zeros(nI, dtype=np.int32); nnDist = np.zeros(nI); lu = np.zeros(2 * n, dtype=np.int32)
    args = [sx.scalar_int(n), sx.scalar_int(m), sx.real(coords), sx.lib.rshim_wrap_int(ip(nnIndx), nnIndx.size),
            sx.lib.rshim_wrap_real(dp(nnDist), nnDist.size), sx.lib.rshim_wrap_int(ip(lu), lu.size),
            sx.scalar_int(nThreads)]
    sx.call(symbol, args)
    sx.reset()
    return nnIndx, nnDist, lu


def call_mk_u_indx(sx: _Sexp, n, m, nnIndx, nnIndxLU, search_type=2, symbol="mkUIndx"):
    """Drive the 8-argument `mkUIndx` .Call entry point (src/nn.cpp:25; R/nn.R:19): outputs written into the arguments."""
    nnIndx = i32(nnIndx); nnIndxLU = i32(nnIndxLU); nI = max(n_indx(n, m), 1)
    u = np.zeros(nI, dtype=np.int32); ui = np.zeros(nI, dtype=np.int32); lu = np.zeros(2 * n, dtype=np.int32)
    args = [sx.scalar_int(n), sx.scalar_int(m), sx.integer(nnIndx), sx.lib.rshim_wrap_int(ip(u), u.size),
            sx.lib.rshim_wrap_int(ip(lu), lu.size), sx.lib.rshim_wrap_int(ip(ui), ui.size), sx.integer(nnIndxLU),
            sx.scalar_int(search_type)]
    sx.call(symbol, args)
    sx.reset()
    return u, lu, ui


class Ref:
    """The reference's own C++ (unmodified), compiled against rshim. None of this is product code."""

    SYM_UPDATE_BF = "_Z8updateBFPdS_S_S_S_PiS0_iiS_iiiiiS_d"      # src/rNNGP.cpp:46
    SYM_UPDATE_CONJ_BF = "_Z12updateConjBFPdS_S_S_S_PiS0_iidddiS_d"  # src/cNNGP.cpp:24
    SYM_Q = "_Z1QPdS_S_S_iPiS0_"                                  # src/util.cpp:261
    SYM_SPCOR = "_Z5spCorRdS_S_RiPd"                              # src/util.cpp:208

    def __init__(self, path: str = REF_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = L = C.CDLL(path)
        self.sx = _Sexp(L)
        getattr(L, self.SYM_UPDATE_BF).restype = C.c_double
        getattr(L, self.SYM_Q).restype = C.c_double
        getattr(L, self.SYM_SPCOR).restype = C.c_double

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_SO)

    def set_seed(self, seed):
        self.lib.rshim_set_seed(C.c_uint(seed))

    def sp_cor(self, D, phi, nu, cov):
        d = C.c_double(D); ph = C.c_double(phi); v = C.c_double(nu); cm = C.c_int(cov)
        bk = np.zeros(64)
        return float(getattr(self.lib, self.SYM_SPCOR)(C.byref(d), C.byref(ph), C.byref(v), C.byref(cm), dp(bk)))

    def update_bf(self, coords, nnIndx, nnIndxLU, n, m, theta, cov, threads=1):
        import ctypes
        coords = coords_flat(coords); nnIndx = i32(nnIndx); nnIndxLU = i32(nnIndxLU)
        th = np.zeros(4); th[: len(theta)] = theta
        B = np.zeros(max(n_indx(n, m), 1)); F = np.zeros(n)
        T = max(int(threads), 1)
        self._omp(T)
        c = np.zeros(m * T); Cm = np.zeros(m * m * T); nu_b = 8.0
        bk = np.zeros(T * (1 + int(nu_b)))
        ld = getattr(self.lib, self.SYM_UPDATE_BF)(
            dp(B), dp(F), dp(c), dp(Cm), dp(coords), ip(nnIndx), ip(nnIndxLU), C.c_int(n), C.c_int(m), dp(th),
            C.c_int(1), C.c_int(0), C.c_int(2), C.c_int(3), C.c_int(cov), dp(bk), C.c_double(nu_b))
        return B, F, float(ld)

    def update_conj_bf(self, coords, nnIndx, nnIndxLU, n, m, phi, alpha, nu, cov, threads=1):
        coords = coords_flat(coords); nnIndx = i32(nnIndx); nnIndxLU = i32(nnIndxLU)
        B = np.zeros(max(n_indx(n, m), 1)); F = np.zeros(n)
        T = max(int(threads), 1)
        self._omp(T)
        c = np.zeros(m * T); Cm = np.zeros(m * m * T); bk = np.zeros(T * 16)
        getattr(self.lib, self.SYM_UPDATE_CONJ_BF)(
            dp(B), dp(F), dp(c), dp(Cm), dp(coords), ip(nnIndx), ip(nnIndxLU), C.c_int(n), C.c_int(m),
            C.c_double(phi), C.c_double(alpha), C.c_double(nu), C.c_int(cov), dp(bk), C.c_double(8.0))
        return B, F

    def q(self, B, F, u, v, n, nnIndx, nnIndxLU):
        B = f64(B); F = f64(F); u = f64(u); v = f64(v); nnIndx = i32(nnIndx); nnIndxLU = i32(nnIndxLU)
        return float(getattr(self.lib, self.SYM_Q)(dp(B), dp(F), dp(u), dp(v), C.c_int(n), ip(nnIndx), ip(nnIndxLU)))

    def _omp(self, T):
        try:
            omp = C.CDLL("libgomp.so.1")
            omp.omp_set_num_threads(C.c_int(T))
        except OSError:
            pass

    def mk_nn_indx(self, coords, m, threads=1):
        return call_mk_nn(self.sx, coords, m, threads, "mkNNIndx")

    def mk_nn_indx_cb(self, coords, m, threads=1):
        return call_mk_nn(self.sx, coords, m, threads, "mkNNIndxCB")

    def rnngp(self, *a, **k):
        return call_rnngp(self.sx, *a, **k)

    def rnngp_replicated(self, *a, **k):
        return call_rnngp_replicated(self.sx, *a, **k)

    def mk_u_indx(self, *a, **k):
        return call_mk_u_indx(self.sx, *a, **k)

    def cnngp(self, *a, **k):
        return call_cnngp(self.sx, *a, **k)

    def predict(self, *a, **k):
        return call_predict(self.sx, *a, **k)


# ---- k-fold scoring of the conjugate model: a restatement of the reference's R layer --------------------
def crps(y, y_hat, y_var):
    """R/util.R:247-251.  dnorm / pnorm are R nmath in the reference; scipy's ndtr stands in (unpinned, ~1e-16)."""
    from scipy.special import ndtr
    y = np.asarray(y, dtype=np.float64); sd = np.sqrt(np.asarray(y_var, dtype=np.float64))
    z = (y - np.asarray(y_hat, dtype=np.float64)) / sd
    dn = np.exp(-0.5 * z * z) / np.sqrt(2.0 * np.pi)
    return float(np.mean(-sd * (1.0 / np.sqrt(np.pi) - 2.0 * dn - z * (2.0 * ndtr(z) - 1.0))))


def rmspe(y, y_hat):
    """R/util.R:253-255."""
    d = np.asarray(y, dtype=np.float64) - np.asarray(y_hat, dtype=np.float64)
    return float(np.sqrt(np.mean(d * d)))


def conj_kfold(backend, port, y, X, coords, m, cov, theta_alpha, sigmaSqIG, ho_idx, search_type="cb"):
    """The k-fold loop of spConjNNGP, R/spConjNNGP.R:286-349, with `backend` (Port or Ref) supplying mkNNIndx[CB] and
    cNNGP and `port.nn_query` standing in for RANN::nn2 (:311).  ho_idx: k_fold x n_ho, 0-based, ascending per fold
    (R: sort(sample(1:n, floor(n/k.fold))), :289).  Returns k.fold.scores/k.fold, g x 2 (columns rmspe, crps)."""
    y = f64(y); n = len(y)
    X = np.asarray(X, dtype=np.float64).reshape(n, -1)
    coords = np.asarray(coords, dtype=np.float64).reshape(n, 2)
    ta = np.asarray(theta_alpha, dtype=np.float64); g = ta.shape[1]
    ho_idx = np.asarray(ho_idx); k_fold = ho_idx.shape[0]
    scores = np.zeros((g, 2))
    for f in range(k_fold):
        ho = ho_idx[f]
        keep = np.ones(n, dtype=bool); keep[ho] = False
        X_mod, y_mod, c_mod = X[keep], y[keep], coords[keep]          # :291-294
        X_ho, y_ho, c_ho = X[ho], y[ho], coords[ho]                   # :296-299
        mk = backend.mk_nn_indx if search_type in (0, "brute") else backend.mk_nn_indx_cb   # :302-306
        nnIndx, _, lu = mk(c_mod, m)
        nn0 = port.nn_query(c_mod, c_ho, m)                           # :311
        out = backend.cnngp(y_mod, X_mod, c_mod, ta, nnIndx, lu, X_ho, c_ho, nn0, m, sigmaSqIG, cov)  # :322-328
        yh = out["y0_hat"] if "y0_hat" in out else out["y.0.hat"]
        yv = out["y0_hat_var"] if "y0_hat_var" in out else out["y.0.hat.var"]
        yh = np.asarray(yh).reshape(len(ho), g, order="F") if np.asarray(yh).ndim == 1 else np.asarray(yh)
        yv = np.asarray(yv).reshape(len(ho), g, order="F") if np.asarray(yv).ndim == 1 else np.asarray(yv)
        for j in range(g):                                            # :330-334
            scores[j, 1] += crps(y_ho, yh[:, j], yv[:, j])
            scores[j, 0] += rmspe(y_ho, yh[:, j])
    return scores / k_fold                                             # :345
