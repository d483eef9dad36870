"""Thin numpy-facing wrapper over the C ABI (include/nngp_b200.h). One Engine drives one GPU.

All numerics run in libnngp_b200.so on the device; this module only marshals arrays into the
reference's column-major layout and raises on non-zero status codes.
"""
from __future__ import annotations

import ctypes as C

import os

import numpy as np

from . import _lib
from ._lib import RnngpArgs, c_double_p, c_int_p

COV_MODELS = {"exponential": 0, "spherical": 1, "matern": 2, "gaussian": 3}  # R/spNNGP.R:172-178


class NNGPError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"nngp status {status}: {message}")
        self.status = status


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _ip(a):
    return a.ctypes.data_as(c_int_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def colmajor(a):
    """2-D array -> flat column-major vector (R matrix layout)."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        return np.ascontiguousarray(a)
    return np.ascontiguousarray(a.T).ravel()


def n_indx(n, m):
    """Length of nnIndx / B (reference src/rNNGP.cpp:218)."""
    return int((1 + m) / 2 * m + (n - m - 1) * m)


def _as_cfunc(fn, proto):
    """Python callable, CFUNCTYPE instance or a symbol of a loaded C library -> `proto` instance."""
    if isinstance(fn, proto):
        return fn
    if isinstance(fn, C._CFuncPtr):  # e.g. librshim's rnorm: pass the C address straight through
        return proto(C.cast(fn, C.c_void_p).value)
    return proto(fn)


class Engine:
    def __init__(self, device: int = -1, stream=None):
        self.lib = _lib.load()
        h = C.c_void_p()
        st = self.lib.nngp_create(C.c_int(device), C.c_void_p(stream) if stream else None, C.byref(h))
        if st != 0:
            raise NNGPError(st, self.lib.nngp_last_error(None).decode())
        self.h = h
        self.n = self.m = self.p = 0
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.lib.nngp_comm_destroy(self.h)
            self.lib.nngp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st != 0:
            raise NNGPError(st, self.lib.nngp_last_error(self.h).decode())

    # ---- data ------------------------------------------------------------------------------------
    def set_data(self, coords, nnIndx, nnIndxLU, X, y, m, loc_range=None):
        coords = colmajor(coords)
        n = coords.size // 2
        nnIndx = _i32(nnIndx) if nnIndx is not None else None  # None: prediction-only use (sites, X, y)
        lu = _i32(nnIndxLU) if nnIndxLU is not None else None
        if X is None:
            p = 0; Xf = None; yf = None
        else:
            X = np.asarray(X, dtype=np.float64).reshape(n, -1)
            p = X.shape[1]
            Xf = colmajor(X); yf = _f64(y)
        lo, hi = loc_range if loc_range is not None else (0, n)
        self._check(self.lib.nngp_set_data(self.h, n, m, p, _dp(coords), _ip(nnIndx) if nnIndx is not None else None, _ip(lu) if lu is not None else None,
                                           _dp(Xf) if Xf is not None else None, _dp(yf) if yf is not None else None,
                                           lo, hi))
        self.n, self.m, self.p = n, m, p
        self.loc_range = (lo, hi)

    # ---- hot path ----------------------------------------------------------------------------------
    def bf_loglik(self, cov, thetas, betas=None, slots=None):
        """thetas: (nb, <=4) rows (sigma.sq, tau.sq, phi[, nu]); betas: (nb, p) or None.
        Returns (nb, 2): sum log F and Q over this engine's shard."""
        th = np.zeros((np.atleast_2d(thetas).shape[0], 4))
        t2 = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
        th[:, : t2.shape[1]] = t2
        nb = th.shape[0]
        bp = None
        if betas is not None:
            b = _f64(np.atleast_2d(betas))
            assert b.shape == (nb, self.p)
            bp = _dp(b)
        sl = _i32(slots) if slots is not None else None
        out = np.zeros((nb, 2))
        self._check(self.lib.nngp_bf_loglik(self.h, cov, nb, _dp(th), bp, _ip(sl) if sl is not None else None, _dp(out)))
        return out

    def bf_loglik_cur(self, cov, theta, beta, slot_new, slot_cur):
        """One launch of the sampler's step (nngp_bf_loglik_cur): solves `theta` into slot_new and returns
        (sum log F(theta), Q(theta), Q of the factors resident in slot_cur), all with r = X beta - y."""
        th = np.zeros(4); t = np.asarray(theta, dtype=np.float64).ravel(); th[: t.size] = t
        b = _f64(beta)
        assert b.size == self.p
        out = np.zeros(3)
        self._check(self.lib.nngp_bf_loglik_cur(self.h, cov, _dp(th), _dp(b), int(slot_new), int(slot_cur), _dp(out)))
        return out

    def bf_loglik_vec(self, cov, thetas, u, slots=None):
        """As bf_loglik with the vector to whiten given directly: returns (nb, 2) = (sum log F, u' C~^-1 u).
        tau.sq = 0 and u = w is the latent model's likelihood piece (src/sNNGP.cpp:24-65, 323-400)."""
        t2 = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
        th = np.zeros((t2.shape[0], 4)); th[:, : t2.shape[1]] = t2
        nb = th.shape[0]
        uf = _f64(u)
        assert uf.size == self.n
        sl = _i32(slots) if slots is not None else None
        out = np.zeros((nb, 2))
        self._check(self.lib.nngp_bf_loglik_vec(self.h, cov, nb, _dp(th), _dp(uf), _ip(sl) if sl is not None else None, _dp(out)))
        return out

    def time_bf_loglik(self, cov, thetas, betas=None, iters=10, flush_l2=False):
        th = np.zeros((np.atleast_2d(thetas).shape[0], 4))
        t2 = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
        th[:, : t2.shape[1]] = t2
        nb = th.shape[0]
        bp = None
        if betas is not None:
            b = _f64(np.atleast_2d(betas)); bp = _dp(b)
        ms = C.c_float(0)
        self._check(self.lib.nngp_time_bf_loglik(self.h, cov, nb, _dp(th), bp, iters, 1 if flush_l2 else 0, C.byref(ms)))
        return float(ms.value)

    def q_resid(self, slot, beta):
        b = _f64(beta); out = np.zeros(1)
        self._check(self.lib.nngp_q_resid(self.h, slot, _dp(b), _dp(out)))
        return float(out[0])

    def q_forms(self, slot):
        K = self.p + 1
        S = np.zeros(K * K)
        self._check(self.lib.nngp_q_forms(self.h, slot, _dp(S)))
        return S.reshape(K, K)

    def get_bf(self, slot):
        B = np.zeros(max(n_indx(self.n, self.m), 1)); F = np.zeros(self.n)
        self._check(self.lib.nngp_get_bf(self.h, slot, _dp(B), _dp(F)))
        return B, F

    def comm_init_from_torch(self, group=None, nccl=True):
        """Join the ranks of torch.distributed's default (or given) group: afterwards likelihood results are
        rank-summed.  nccl=True: the engine's NCCL communicator plus the shared-memory mailbox (nngp_comm_init; rank 0
        makes the id, torch broadcasts it).  nccl=False: the mailbox alone (nngp_comm_init_local) -- the ranks of one
        node, which may even share a GPU; works over a gloo group."""
        import torch.distributed as dist
        world = dist.get_world_size(group); rank = dist.get_rank(group)
        if world == 1:
            return
        if not nccl:
            obj = [os.urandom(128) if rank == 0 else None]
            dist.broadcast_object_list(obj, src=0, group=group)
            self._check(self.lib.nngp_comm_init_local(self.h, obj[0], rank, world))
            return
        buf = C.create_string_buffer(128)
        if rank == 0:
            st = self.lib.nngp_comm_unique_id(buf)
            if st != 0:
                raise NNGPError(st, self.lib.nngp_last_error(None).decode())
        obj = [buf.raw if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0, group=group)
        self._check(self.lib.nngp_comm_init(self.h, obj[0], rank, world))

    def comm_allreduce(self, arr):
        a = _f64(arr).copy()
        self._check(self.lib.nngp_comm_allreduce_sum(self.h, _dp(a), a.size))
        return a

    def pinned_array(self, count, dtype=np.float64):
        """A page-locked numpy array (include/nngp_b200.h: nngp_host_alloc): inputs / outputs placed in such arrays are
        moved by direct DMA instead of through the staging ring.  Freed by `free_pinned` (or with the process)."""
        dt = np.dtype(dtype)
        ptr = C.c_void_p()
        self._check(self.lib.nngp_host_alloc(self.h, int(count) * dt.itemsize, C.byref(ptr)))
        buf = (C.c_char * (int(count) * dt.itemsize)).from_address(ptr.value)
        a = np.frombuffer(buf, dtype=dt, count=int(count))
        self._pinned = getattr(self, "_pinned", {})
        self._pinned[a.ctypes.data] = ptr
        return a

    def free_pinned(self, a):
        ptr = getattr(self, "_pinned", {}).pop(a.ctypes.data, None)
        if ptr is not None:
            self._check(self.lib.nngp_host_free(self.h, ptr))

    def fp64_peak_tflops(self):
        out = np.zeros(1)
        self._check(self.lib.nngp_fp64_peak(self.h, _dp(out)))
        return float(out[0])

    def set_option(self, name, value):
        """Tuning switch of the engine (include/nngp_b200.h: nngp_set_option), e.g. ("matern_short", 0)."""
        self._check(self.lib.nngp_set_option(self.h, name.encode(), int(value)))

    def get_option(self, name):
        v = C.c_int(0)
        self._check(self.lib.nngp_get_option(self.h, name.encode(), C.byref(v)))
        return v.value

    def launch_count(self):
        return int(self.lib.nngp_launch_count(self.h))

    # ---- entry points ------------------------------------------------------------------------------
    def rnngp(self, cov, sigmaSqIG, tauSqIG, phiUnif, nuUnif, starting, tuning, n_samples, rnorm, runif, mode=1,
              verbose=False, n_report=100, printer=None, allreduce=None, rep_indx=None, interrupt=None):
        """Response-model sampler on the data given to set_data (reference `rNNGP`, src/rNNGP.cpp:105).
        rnorm/runif: ctypes function pointers or Python callables (mu, sd) / (a, b).
        interrupt: callable polled once per iteration (R_CheckUserInterrupt's place); a nonzero return stops the chain.
        rep_indx: 0/1 per iteration (fit.rep, src/rNNGP.cpp:203-208); then a 4th result, y.rep n x nRep, is returned."""
        a = RnngpArgs()
        a.cov_model = cov
        a.sigmaSqIG[:] = list(sigmaSqIG); a.tauSqIG[:] = list(tauSqIG); a.phiUnif[:] = list(phiUnif)
        a.nuUnif[:] = list(nuUnif) if nuUnif is not None else [0.0, 0.0]
        b0 = _f64(starting["beta"]); a.betaStarting = _dp(b0)
        a.sigmaSqStarting = starting["sigma.sq"]; a.tauSqStarting = starting["tau.sq"]; a.phiStarting = starting["phi"]
        a.nuStarting = starting.get("nu", 0.0)
        a.sigmaSqTuning = tuning["sigma.sq"]; a.tauSqTuning = tuning["tau.sq"]; a.phiTuning = tuning["phi"]
        a.nuTuning = tuning.get("nu", 0.0)
        a.nSamples = n_samples; a.mode = mode; a.verbose = int(bool(verbose)); a.nReport = n_report
        a.rnorm = _as_cfunc(rnorm, _lib.RNORM_FN)
        a.runif = _as_cfunc(runif, _lib.RUNIF_FN)
        if printer is not None:
            a.print = _lib.PRINT_FN(lambda s: printer(s.decode()))
        if interrupt is not None:
            a.interrupt = _lib.INTERRUPT_FN(lambda: int(interrupt() or 0))
        if allreduce is not None:
            def _ar(buf, n, _user):
                arr = np.ctypeslib.as_array(buf, shape=(n,))
                allreduce(arr)
            a.allreduce = _lib.ALLREDUCE_FN(_ar)
        nTheta = 4 if cov == COV_MODELS["matern"] else 3
        bs = np.zeros(self.p * n_samples); ts = np.zeros(nTheta * n_samples)
        acc = C.c_int(0)
        rep = None
        if rep_indx is not None:
            ri = _i32(rep_indx)
            if ri.size != n_samples:
                raise ValueError("rep_indx needs one flag per iteration")
            rep = np.zeros(max(int((ri != 0).sum()) * self.n, 1))
            a.repIndx = _ip(ri); a.repSamples = _dp(rep)
        self._check(self.lib.nngp_rnngp(self.h, C.byref(a), _dp(bs), _dp(ts), C.byref(acc)))
        res = (bs.reshape(n_samples, self.p).T.copy(), ts.reshape(n_samples, nTheta).T.copy(), int(acc.value))
        if rep is not None:
            k = int((ri != 0).sum())
            res = res + (rep[: k * self.n].reshape(k, self.n).T.copy(),)
        return res

    def rep_draw(self, slot, beta, z):
        """One replicate y.rep = X beta + v from the B, F of `slot` and n standard normals z (reference `updateRep`,
        src/rNNGP.cpp:24-42)."""
        b = _f64(beta); zf = _f64(z); out = np.zeros(self.n)
        self._check(self.lib.nngp_rep_draw(self.h, slot, _dp(b), _dp(zf), _dp(out)))
        return out

    def rnngp_replicated(self, cov, beta_samples, theta_samples, rnorm):
        """Replicates for given posterior samples (reference `rNNGPReplicated`, src/rNNGPReplicated.cpp:76).
        beta_samples S x p (as R passes it, R/generics.R:252), theta_samples nTheta x S.  Returns n x S."""
        bs = np.asarray(beta_samples, dtype=np.float64); S = bs.shape[0]
        bsf = colmajor(bs); tsf = colmajor(np.asarray(theta_samples, dtype=np.float64))
        out = np.zeros(max(self.n * S, 1))
        self._check(self.lib.nngp_rnngp_replicated(self.h, cov, _dp(bsf), _dp(tsf), S, _as_cfunc(rnorm, _lib.RNORM_FN),
                                                   _lib.INTERRUPT_FN(0), _dp(out)))
        return out[: self.n * S].reshape(S, self.n).T.copy()

    def rnngp_predict(self, cov, X0, coords0, nnIndx0, beta_samples, theta_samples, z, site_range=None, out=None):
        """Posterior predictive draws (reference `rNNGPPredict`, src/rNNGPPredict.cpp:25).
        beta_samples p x S, theta_samples nTheta x S, nnIndx0 flat q*m column-major, z flat q*S (z[i*S+s]).
        Returns q x S.  `out`: optional flat q*S float64 array for the draws (e.g. `pinned_array`: direct DMA)."""
        X0 = np.asarray(X0, dtype=np.float64).reshape(-1, self.p); q = X0.shape[0]
        bs = np.asarray(beta_samples, dtype=np.float64); S = bs.shape[1]
        x0f = colmajor(X0); c0 = colmajor(coords0); nn0 = _i32(nnIndx0)
        bsf = colmajor(bs); tsf = colmajor(np.asarray(theta_samples, dtype=np.float64)); zf = _f64(z)
        if out is not None:
            assert out.dtype == np.float64 and out.size == q * S and out.flags.c_contiguous
            y0 = out
        else:
            y0 = np.zeros(q * S)
        lo, hi = site_range if site_range is not None else (0, q)
        self._check(self.lib.nngp_rnngp_predict(self.h, cov, q, _dp(x0f), _dp(c0), _ip(nn0), S, _dp(bsf), _dp(tsf),
                                                _dp(zf), _dp(y0), lo, hi))
        return y0.reshape(S, q).T  # q x S, column-major like the reference's p.y.0 (a view: no 8 GB transpose at C4 sizes)

    def cnngp(self, cov, theta_alpha, sigmaSqIG, X0=None, coords0=None, nnIndx0=None):
        """Conjugate model for g parameter sets (reference `cNNGP`, src/cNNGP.cpp:73).
        theta_alpha: nTheta x g rows (phi, alpha[, nu])."""
        ta = np.asarray(theta_alpha, dtype=np.float64); nTheta, g = ta.shape
        taf = colmajor(ta); ig = _f64(sigmaSqIG); p = self.p
        if X0 is None:
            n0 = 0; x0p = c0p = nnp = None
        else:
            X0 = np.asarray(X0, dtype=np.float64).reshape(-1, p); n0 = X0.shape[0]
            x0f = colmajor(X0); c0 = colmajor(coords0); nn0 = _i32(nnIndx0)
            x0p, c0p, nnp = _dp(x0f), _dp(c0), _ip(nn0)
        bh = np.zeros(p * g); ab = np.zeros(nTheta * g); bBi = np.zeros(p * p * g); bb = np.zeros(p * g)
        y0 = np.zeros(max(n0 * g, 1)); y0v = np.zeros(max(n0 * g, 1))
        self._check(self.lib.nngp_cnngp(self.h, cov, g, _dp(taf), _dp(ig), n0, x0p, c0p, nnp, _dp(bh), _dp(ab), _dp(bBi),
                                        _dp(bb), _dp(y0), _dp(y0v)))
        return dict(beta_hat=bh.reshape(g, p).T.copy(), ab=ab, bB_inv=bBi.reshape(g, p * p).T.copy(),
                    bb=bb.reshape(g, p).T.copy(), y0_hat=y0[: n0 * g].reshape(g, n0).T.copy(),
                    y0_hat_var=y0v[: n0 * g].reshape(g, n0).T.copy())

    def cnngp_score(self, cov, theta_alpha, sigmaSqIG, X0, coords0, nnIndx0, y0):
        """Hold-out scores of the conjugate model, computed on the device (R/spConjNNGP.R:322-334, R/util.R:247-255).
        Returns (crps[g], rmspe[g])."""
        ta = np.asarray(theta_alpha, dtype=np.float64); nTheta, g = ta.shape
        taf = colmajor(ta); ig = _f64(sigmaSqIG); p = self.p
        X0 = np.asarray(X0, dtype=np.float64).reshape(-1, p); n0 = X0.shape[0]
        x0f = colmajor(X0); c0 = colmajor(coords0); nn0 = _i32(nnIndx0); yo = _f64(y0)
        crps = np.zeros(g); rmspe = np.zeros(g)
        self._check(self.lib.nngp_cnngp_score(self.h, cov, g, _dp(taf), _dp(ig), n0, _dp(x0f), _dp(c0), _ip(nn0), _dp(yo),
                                              _dp(crps), _dp(rmspe)))
        return crps, rmspe

    def conj_kfold(self, y, X, coords, m, cov, theta_alpha, sigmaSqIG, ho_idx, search_type="cb", set_range=None):
        """The k-fold loop of spConjNNGP (R/spConjNNGP.R:286-349) in one call.  ho_idx: k_fold x n_ho hold-out rows
        (0-based, ascending per fold).  Returns g x 2 scores (columns rmspe, crps; means over folds); with
        set_range=(lo, hi) only those parameter sets are evaluated and the other rows are zero."""
        ta = np.asarray(theta_alpha, dtype=np.float64); nTheta, g = ta.shape
        taf = colmajor(ta); ig = _f64(sigmaSqIG)
        X = np.asarray(X, dtype=np.float64); n, p = X.shape
        xf = colmajor(X); cf = colmajor(coords); yf = _f64(y)
        ho = np.ascontiguousarray(np.asarray(ho_idx, dtype=np.int32)); k_fold, n_ho = ho.shape
        lo, hi = set_range if set_range is not None else (0, g)
        st = 0 if search_type in (0, "brute") else 1
        scores = np.zeros(2 * g)
        self._check(self.lib.nngp_conj_kfold(self.h, n, p, m, _dp(yf), _dp(xf), _dp(cf), cov, g, _dp(taf), _dp(ig), st,
                                             k_fold, n_ho, _ip(ho), lo, hi, _dp(scores)))
        self.n, self.m, self.p = n - n_ho, m, p
        return scores.reshape(2, g).T.copy()

    def mk_nn_indx(self, coords, m, search_type="cb"):
        """Ordered neighbour search (reference `mkNNIndx` / `mkNNIndxCB`, src/nn.cpp:63,214)."""
        cf = colmajor(coords); n = cf.size // 2
        nI = max(n_indx(n, m), 1)
        nnIndx = np.zeros(nI, dtype=np.int32); nnDist = np.zeros(nI); lu = np.zeros(2 * n, dtype=np.int32)
        ties = C.c_int(0)
        st = 0 if search_type in (0, "brute") else 1
        self._check(self.lib.nngp_mk_nn_indx(self.h, n, m, _dp(cf), st, _ip(nnIndx), _dp(nnDist), _ip(lu), C.byref(ties)))
        return nnIndx, nnDist, lu

    def mk_u_indx(self, n, m, nnIndx, search_type=2):
        """Transposed neighbour graph (reference `mkUIndx`, src/nn.cpp:25): returns uIndx, uIndxLU, uiIndx."""
        nn = _i32(nnIndx); nI = max(n_indx(n, m), 1)
        u = np.zeros(nI, dtype=np.int32); ui = np.zeros(nI, dtype=np.int32); lu = np.zeros(2 * n, dtype=np.int32)
        self._check(self.lib.nngp_mk_u_indx(self.h, n, m, _ip(nn), int(search_type), _ip(u), _ip(lu), _ip(ui)))
        return u, lu, ui

    def nn_query(self, coords, coords0, m):
        """Unordered query kNN, q x m column-major flat (stand-in for RANN::nn2(...)$nn.idx - 1)."""
        cf = colmajor(coords); c0 = colmajor(coords0); n = cf.size // 2; q = c0.size // 2
        out = np.zeros(q * m, dtype=np.int32)
        self._check(self.lib.nngp_nn_query(self.h, n, _dp(cf), q, _dp(c0), m, _ip(out)))
        return out
