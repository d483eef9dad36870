"""CPU: the libR stand-in (RNG, sort, BLAS/LAPACK subset) and the host builds of the device math."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def rshim():
    from spnngp_b200 import _lib
    return _lib.load_rshim()


def test_rng_matches_known_r_values(rshim):
    """set.seed(1); runif(1) -> 0.2655087, rnorm(1) -> -0.6264538; set.seed(42); runif(1) -> 0.914806
    (R's default Mersenne-Twister / Inversion; well-known values)."""
    rshim.rshim_set_seed(1); assert abs(rshim.unif_rand() - 0.2655087) < 5e-8
    rshim.rshim_set_seed(1); assert abs(rshim.norm_rand() - (-0.6264538)) < 5e-8
    rshim.rshim_set_seed(42); assert abs(rshim.unif_rand() - 0.914806) < 5e-7
    rshim.rshim_set_seed(123); x = [rshim.rnorm(0.0, 1.0) for _ in range(3)]
    assert np.allclose(x, [-0.56047565, -0.23017749, 1.55870831], atol=5e-8)  # set.seed(123); rnorm(3)


def test_qnorm_accuracy(rshim):
    from scipy.stats import norm
    rshim.qnorm.restype = C.c_double
    rshim.qnorm.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]
    p = np.concatenate([np.linspace(1e-12, 1 - 1e-12, 2001), 10.0 ** -np.arange(1, 300, 7)])
    got = np.array([rshim.qnorm(float(x), 0.0, 1.0, 1, 0) for x in p])
    assert np.allclose(got, norm.ppf(p), rtol=1e-13, atol=1e-14)


def test_rsort_with_index_semantics(rshim, port):
    """Shell sort with R's gap rule: n <= 8 is a stable insertion sort, larger n is unstable."""
    rshim.rsort_with_index.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_int]
    rng = np.random.default_rng(0)
    for n in (1, 2, 8, 9, 15, 32, 200):
        x = rng.integers(0, 5, n).astype(np.float64); idx = np.arange(n, dtype=np.int32)
        xs, ids = port.rsort_with_index(x, idx)
        x2, i2 = x.copy(), idx.copy()
        rshim.rsort_with_index(x2.ctypes.data_as(C.POINTER(C.c_double)), i2.ctypes.data_as(C.POINTER(C.c_int)), n)
        assert np.array_equal(xs, np.sort(x)) and np.array_equal(x2, xs) and np.array_equal(i2, ids)
        if n <= 8:
            assert np.array_equal(ids, np.argsort(x, kind="stable"))
    x = np.array([2.0, np.nan, 1.0, np.inf, -1.0]); xs, ids = port.rsort_with_index(x, np.arange(5))
    assert np.isnan(xs[-1]) and np.array_equal(xs[:4], [-1.0, 1.0, 2.0, np.inf])


def test_minilapack_against_numpy(rshim):
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    rng = np.random.default_rng(1)
    for n in (1, 2, 5, 15, 32):
        A = rng.standard_normal((n, n)); A = A @ A.T + n * np.eye(n)
        a = np.asfortranarray(A).ravel(order="F").copy()
        info = C.c_int(0); nn = C.c_int(n); lo = C.c_char_p(b"L")
        rshim.dpotrf_(lo, C.byref(nn), a.ctypes.data_as(dp), C.byref(nn), C.byref(info))
        assert info.value == 0
        Lc = np.tril(a.reshape(n, n, order="F"))
        assert np.allclose(Lc, np.linalg.cholesky(A), rtol=1e-12, atol=1e-13)
        rshim.dpotri_(lo, C.byref(nn), a.ctypes.data_as(dp), C.byref(nn), C.byref(info))
        Ai = np.tril(a.reshape(n, n, order="F"))
        assert np.allclose(Ai, np.tril(np.linalg.inv(A)), rtol=1e-10, atol=1e-13)
        x = rng.standard_normal(n); y = np.zeros(n); one = C.c_double(1.0); zero = C.c_double(0.0); inc = C.c_int(1)
        rshim.dsymv_(lo, C.byref(nn), C.byref(one), a.ctypes.data_as(dp), C.byref(nn), x.ctypes.data_as(dp), C.byref(inc),
                     C.byref(zero), y.ctypes.data_as(dp), C.byref(inc))
        assert np.allclose(y, np.linalg.inv(A) @ x, rtol=1e-10)
    bad = np.array([1.0, 2.0, 2.0, 1.0]); info = C.c_int(0); two = C.c_int(2)
    rshim.dpotrf_(C.c_char_p(b"L"), C.byref(two), bad.ctypes.data_as(dp), C.byref(two), C.byref(info))
    assert info.value == 2  # not positive definite: the reference turns this into "dpotrf failed"


def _build_host(src, out):
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-mfma",
                    os.path.join(ROOT, "tests", "csrc", src), "-o", out], check=True)
    return C.CDLL(out)


def test_device_besselk_on_host_vs_mpmath(tmp_path):
    """spnngp_b200/csrc/besselk.cuh compiled for the host: K_nu and the Matern correlation vs mpmath."""
    import mpmath as mp
    L = _build_host("besselk_host.cpp", str(tmp_path / "libbk.so"))
    for f in (L.host_besselk, L.host_matern):
        f.restype = C.c_double; f.argtypes = [C.c_double, C.c_double]
    mp.mp.dps = 40
    worst = 0.0
    for nu in (0.05, 0.25, 0.49, 0.5, 0.51, 0.75, 1.0, 1.5, 2.3, 2.5, 3.7, 4.5, 5.0):
        for x in (1e-8, 1e-4, 0.01, 0.1, 0.5, 1.0, 1.99, 2.0, 2.01, 3.0, 8.0, 20.0, 100.0, 600.0):
            k = mp.besselk(nu, x)
            worst = max(worst, float(abs((mp.mpf(L.host_besselk(x, nu)) - k) / k)))
            rho = mp.mpf(x) ** nu * k / (mp.mpf(2) ** (nu - 1) * mp.gamma(nu))
            if rho > 1e-290:
                worst = max(worst, float(abs((mp.mpf(L.host_matern(x, nu)) - rho) / rho)))
    assert worst < 5e-14
    for x in (0.01, 0.3, 1.7, 4.0):  # closed forms, SURVEY §8c KATs
        assert abs(L.host_matern(x, 0.5) - np.exp(-x)) < 1e-15
        assert abs(L.host_matern(x, 1.5) - (1 + x) * np.exp(-x)) < 2e-15
        assert abs(L.host_matern(x, 2.5) - (1 + x + x * x / 3) * np.exp(-x)) < 3e-15


def test_matern_gauss_laguerre_on_host_vs_mpmath(tmp_path):
    """Large-argument Matern path (spnngp_b200/csrc/matern_series.cuh: matern_gl), host build, vs mpmath; the
    Gauss-Laguerre rule itself vs scipy."""
    import mpmath as mp
    from scipy.special import roots_genlaguerre, gamma
    L = _build_host("besselk_host.cpp", str(tmp_path / "libbk.so"))
    L.host_matern_gl.restype = C.c_double; L.host_matern_gl.argtypes = [C.c_double, C.c_double]
    L.host_gl_rule.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    for nu in (0.1, 0.5, 1.0, 2.75):
        t = (C.c_double * 16)(); w = (C.c_double * 16)()
        L.host_gl_rule(nu, t, w)
        ts, ws = roots_genlaguerre(16, nu - 0.5)
        o = np.argsort(np.array(t[:]))
        assert np.allclose(np.array(t[:])[o], ts, rtol=1e-12, atol=0)
        assert np.allclose(np.array(w[:])[o], ws / gamma(nu + 0.5), rtol=1e-9, atol=1e-300)
    mp.mp.dps = 40
    worst2 = worst3 = worst4 = 0.0
    for nu in (0.05, 0.25, 0.5, 0.75, 1.0, 1.0005, 1.5, 2.0, 2.3, 3.0, 3.7, 5.0):
        for x in (2.0, 2.5, 3.0, 3.5, 4.0, 6.0, 9.0, 20.0, 55.0, 100.0, 400.0, 650.0):
            rho = mp.mpf(x) ** nu * mp.besselk(nu, x) / (mp.mpf(2) ** (nu - 1) * mp.gamma(nu))
            err = float(abs((mp.mpf(L.host_matern_gl(x, nu)) - rho) / rho))
            worst2 = max(worst2, err)
            if x >= 3.0: worst3 = max(worst3, err)
            if x >= 4.0: worst4 = max(worst4, err)
    assert worst2 < 1e-12 and worst3 < 1e-13 and worst4 < 5e-14, (worst2, worst3, worst4)
    assert L.host_matern_gl(800.0, 0.75) == 0.0  # underflows like the reference's bessel_k


def test_matern_short_series_on_host_vs_mpmath(tmp_path):
    """The 9-term series (matern_series.cuh, Cor<NNGP_COV_MATERN_SHORT>) over its whole range t <= t_short_max(nu):
    as accurate as the 14-term one (vs mpmath), and the range is wide enough to be useful."""
    import mpmath as mp
    L = _build_host("besselk_host.cpp", str(tmp_path / "libbk.so"))
    L.host_matern_series_short.restype = C.c_double; L.host_matern_series_short.argtypes = [C.c_double, C.c_double]
    L.host_t_short_max.restype = C.c_double; L.host_t_short_max.argtypes = [C.c_double]
    mp.mp.dps = 40
    worst = 0.0
    for nu in (0.1, 0.3, 0.5, 0.75, 0.9, 1.2, 1.5, 1.9, 2.5, 3.3):
        T = L.host_t_short_max(nu)
        assert 0.05 < T <= 1.0, (nu, T)
        for f in (1e-12, 1e-6, 1e-3, 0.1, 0.5, 0.9, 1.0):
            t = T * f
            x = 2.0 * np.sqrt(t)
            rho = mp.mpf(x) ** nu * mp.besselk(nu, x) / (mp.mpf(2) ** (nu - 1) * mp.gamma(nu))
            worst = max(worst, float(abs((mp.mpf(L.host_matern_series_short(t, nu)) - rho) / rho)))
    assert worst < 5e-14, worst
    assert L.host_t_short_max(1.0) == 0.0 and L.host_t_short_max(2.0005) == 0.0  # integer nu: no series at all


def test_matern_tiny_series_on_host_vs_mpmath(tmp_path):
    """The 5-term series (Cor<NNGP_COV_MATERN_TINY>, the variant BASELINE configs[2] runs: t ~ 3e-4 there) over its whole
    range t <= t_tiny_max(nu): as accurate as the 14-term one (vs mpmath); the range covers the benchmark's
    neighbourhoods with room to spare and lies inside the 9-term range."""
    import mpmath as mp
    L = _build_host("besselk_host.cpp", str(tmp_path / "libbk.so"))
    L.host_matern_series_tiny.restype = C.c_double; L.host_matern_series_tiny.argtypes = [C.c_double, C.c_double]
    L.host_t_tiny_max.restype = C.c_double; L.host_t_tiny_max.argtypes = [C.c_double]
    L.host_t_short_max.restype = C.c_double; L.host_t_short_max.argtypes = [C.c_double]
    mp.mp.dps = 40
    worst = 0.0
    for nu in (0.1, 0.3, 0.5, 0.75, 0.9, 1.2, 1.5, 1.9, 2.5, 3.3):
        T = L.host_t_tiny_max(nu)
        assert 1e-3 < T <= L.host_t_short_max(nu), (nu, T)
        for f in (1e-12, 1e-6, 1e-3, 0.1, 0.5, 0.9, 1.0):
            t = T * f
            x = 2.0 * np.sqrt(t)
            rho = mp.mpf(x) ** nu * mp.besselk(nu, x) / (mp.mpf(2) ** (nu - 1) * mp.gamma(nu))
            worst = max(worst, float(abs((mp.mpf(L.host_matern_series_tiny(t, nu)) - rho) / rho)))
    assert worst < 5e-14, worst
    assert L.host_t_tiny_max(1.0) == 0.0 and L.host_t_tiny_max(2.0005) == 0.0  # integer nu: no series at all


def test_integer_slow_path_threshold_is_conservative(tmp_path):
    """matern_set_phi: the producers test `hi32(d2) >= d2_slow_hi` instead of `phi^2 d2 / 4 > t_fast_max`.  Every d2 the
    integer test lets through must satisfy the floating-point bound (up to rounding), and it must not be grossly
    pessimistic (it may only flag d2 within one high-word step, 2^-20 relative, below the bound)."""
    L = _build_host("besselk_host.cpp", str(tmp_path / "libbk.so"))
    L.host_d2_slow_hi.restype = C.c_int
    L.host_d2_slow_hi.argtypes = [C.c_double, C.c_double, C.c_int, C.POINTER(C.c_double)]
    rng = np.random.default_rng(9)
    for nu in (0.3, 0.75, 1.5, 2.5):
        for phi in (0.01, 0.7, 6.0, 123.0, 4.0e4):
            for short in (0, 1):
                tmax = C.c_double(0.0)
                thr = L.host_d2_slow_hi(nu, phi, short, C.byref(tmax))
                bound = tmax.value / (0.25 * phi * phi)                       # d2 at which t reaches t_fast_max
                d2 = bound * np.exp(rng.uniform(-3e-6, 3e-6, 20000))           # a cloud around the bound
                d2 = np.concatenate([d2, bound * np.array([0.0, 1e-300, 0.5, 0.999999, 1.0, 1.000001, 2.0, 1e30])])
                hi = (d2.view(np.int64) >> 32).astype(np.int64)
                passed = hi < thr                                               # the integer test says "fast path"
                t = 0.25 * phi * phi * d2
                assert np.all(t[passed] <= tmax.value * (1 + 4e-16))
                assert np.all(passed[d2 < bound * (1 - 2.0 ** -19)])           # not pessimistic beyond one high-word step
    tmax = C.c_double(0.0)
    assert L.host_d2_slow_hi(1.0, 6.0, 0, C.byref(tmax)) == 0                   # integer nu: every entry takes slow()


def test_device_fastmath_on_host(tmp_path):
    """Branch-free sqrt / exp of spnngp_b200/csrc/fastmath.cuh (host build) vs libm."""
    L = _build_host("fastmath_host.cpp", str(tmp_path / "libfm.so"))
    for f in (L.host_nngp_sqrt, L.host_nngp_exp_nonpos, L.host_nngp_sqrt_eps):
        f.restype = C.c_double; f.argtypes = [C.c_double]
    rng = np.random.default_rng(0)
    xs = np.concatenate([10 ** rng.uniform(-12, 8, 20000), rng.random(10000)])
    s = np.array([L.host_nngp_sqrt(float(x)) for x in xs]); se = np.array([L.host_nngp_sqrt_eps(float(x)) for x in xs])
    assert np.max(np.abs(s - np.sqrt(xs)) / np.spacing(np.sqrt(xs))) <= 1.0
    assert np.max(np.abs(se - np.sqrt(xs)) / np.spacing(np.sqrt(xs))) <= 1.0
    assert L.host_nngp_sqrt(0.0) == 0.0 and abs(L.host_nngp_sqrt_eps(0.0) - 1e-150) < 1e-160
    xe = -np.concatenate([10 ** rng.uniform(-10, 2.85, 20000), rng.random(10000) * 5])
    e = np.array([L.host_nngp_exp_nonpos(float(x)) for x in xe]); r = np.exp(xe)
    assert np.max(np.abs(e - r) / r) < 3e-16
    assert L.host_nngp_exp_nonpos(0.0) == 1.0 and L.host_nngp_exp_nonpos(-800.0) == 0.0


def test_table_driven_exp_log_on_host(tmp_path):
    """Table-driven exp / log of fastmath.cuh + fasttab.cuh (host build) vs mpmath: absolute error ~1e-16."""
    import mpmath as mp
    L = _build_host("fastmath_host.cpp", str(tmp_path / "libfm2.so"))
    for f in (L.host_nngp_exp_tab, L.host_nngp_log_tab):
        f.restype = C.c_double; f.argtypes = [C.c_double]
    mp.mp.dps = 40
    rng = np.random.default_rng(3)
    xs = np.concatenate([-10 ** rng.uniform(-12, 2.838, 3000),  # results under 2^-1000 (x < -693.1) are flushed to 0
                          10 ** rng.uniform(-12, 2.8, 1500), rng.uniform(-2, 2, 1500)])
    worst = 0.0
    for x in xs:
        r = mp.exp(mp.mpf(float(x)))
        worst = max(worst, float(abs((mp.mpf(L.host_nngp_exp_tab(float(x))) - r) / r)))
    assert worst < 4e-16, worst
    assert L.host_nngp_exp_tab(0.0) == 1.0 and L.host_nngp_exp_tab(-800.0) == 0.0 and L.host_nngp_exp_tab(-1500.0) == 0.0
    assert 0.0 <= L.host_nngp_exp_tab(-700.0) < 1e-300
    ts = np.concatenate([10 ** rng.uniform(-300, 300, 3000), 1 + rng.uniform(-0.5, 1.0, 2000), 10 ** rng.uniform(-6, 0, 2000)])
    worst_abs = 0.0
    for t in ts:
        r = mp.log(mp.mpf(float(t)))
        err = float(abs(mp.mpf(L.host_nngp_log_tab(float(t))) - r))
        worst_abs = max(worst_abs, err / max(1.0, float(abs(r))))
    assert worst_abs < 3e-16, worst_abs
    assert abs(L.host_nngp_log_tab(1.0)) < 2e-17


def test_fused_power_kernel_on_host(tmp_path):
    """nngp_pow_tab (fastmath.cuh): x^nu = 2^(32 nu log2 x / 32) with the log's recombination and the exp's reduction
    fused -- vs mpmath over the producers' range (t = (phi D / 2)^2 from 1e-30 to 4, nu from 0.05 to 5), relative error of
    the order of |nu log x| ulps like exp(nu * log x); and the series evaluated through it (matern_series_tab) vs the
    Matern correlation itself for all three series lengths."""
    import mpmath as mp
    L = _build_host("fastmath_host.cpp", str(tmp_path / "libfm3.so"))
    L.host_nngp_pow_tab.restype = C.c_double; L.host_nngp_pow_tab.argtypes = [C.c_double, C.c_double]
    mp.mp.dps = 40
    rng = np.random.default_rng(5)
    worst = 0.0
    for nu in (0.05, 0.3, 0.5, 0.75, 1.5, 2.5, 5.0):
        for x in np.concatenate([10 ** rng.uniform(-30, 0.6, 600), 1 + rng.uniform(-0.5, 1.0, 100), [1.0, 0.5, 2.0, 4.0]]):
            r = mp.mpf(float(x)) ** mp.mpf(nu)
            err = float(abs((mp.mpf(L.host_nngp_pow_tab(float(x), nu)) - r) / r))
            worst = max(worst, err / max(1.0, abs(nu * np.log(x))))
    assert worst < 6e-16, worst  # exp_tab(nu * log_tab(x)) measures 3.2e-16 on the same points
    assert L.host_nngp_pow_tab(1.0, 0.75) == 1.0
    assert 0.0 <= L.host_nngp_pow_tab(0.0, 0.75) < 1e-200 and L.host_nngp_pow_tab(1e-300, 5.0) == 0.0
    Lb = _build_host("besselk_host.cpp", str(tmp_path / "libbk3.so"))
    Lb.host_matern_series_tab.restype = C.c_double; Lb.host_matern_series_tab.argtypes = [C.c_double, C.c_double, C.c_int]
    for name in ("host_t_tiny_max", "host_t_short_max"):
        getattr(Lb, name).restype = C.c_double; getattr(Lb, name).argtypes = [C.c_double]
    worst = 0.0
    for nu in (0.1, 0.3, 0.75, 0.9, 1.5, 2.5, 3.3):
        for terms, T in ((5, Lb.host_t_tiny_max(nu)), (9, Lb.host_t_short_max(nu)), (14, 1.0)):
            for f in (1e-12, 1e-6, 1e-3, 0.03, 0.1, 0.5, 0.9, 1.0):
                t = T * f
                x = 2.0 * np.sqrt(t)
                rho = mp.mpf(x) ** nu * mp.besselk(nu, x) / (mp.mpf(2) ** (nu - 1) * mp.gamma(nu))
                worst = max(worst, float(abs((mp.mpf(Lb.host_matern_series_tab(t, nu, terms)) - rho) / rho)))
    assert worst < 1e-13, worst


def test_kats_on_the_oracle(port):
    """Analytic known answers on the CPU oracle itself."""
    s2, t2, phi = 5.0, 1.0, 6.0
    coords = np.array([[0.1, 0.2], [0.4, 0.6]])
    idx, _, lu = port.mk_nn_indx(coords, 1)
    B, F, ld, info = port.update_bf(coords, idx, lu, 2, 1, [s2, t2, phi], 0)
    rho = np.exp(-phi * 0.5)
    assert abs(F[0] - 6.0) < 1e-15 and abs(B[0] - s2 * rho / 6.0) < 1e-15
    assert abs(F[1] - (6.0 - (s2 * rho) ** 2 / 6.0)) < 1e-14 and abs(ld - np.log(F).sum()) < 1e-14
    # tau.sq = 0, exponential, collinear equally spaced: Markov -> one weight e^{-phi h}, F = s2 (1 - e^{-2 phi h})
    h = 0.05; c = np.column_stack([np.arange(8) * h, np.zeros(8)])
    idx, _, lu = port.mk_nn_indx(c, 3)
    B, F, _, info = port.update_bf(c, idx, lu, 8, 3, [s2, 0.0, phi], 0)
    assert info == 0
    off, cnt = lu[7], lu[8 + 7]
    w = B[off:off + cnt]
    assert abs(w[0] - np.exp(-phi * h)) < 1e-9 and np.all(np.abs(w[1:]) < 1e-9)
    assert abs(F[7] - s2 * (1 - np.exp(-2 * phi * h))) < 1e-9
    # spherical: rho = 0 beyond the range, 1 at 0; Matern(1/2) == exponential
    assert port.sp_cor(1.0, 6.0, 0, 1) == 0.0 and port.sp_cor(0.0, 6.0, 0, 1) == 1.0
    assert abs(port.sp_cor(0.3, 6.0, 0.5, 2) - np.exp(-1.8)) < 1e-15
