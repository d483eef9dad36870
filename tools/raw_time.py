"""Developer probe: time the fused launch through the raw C ABI of any build of the library (old builds included).
    python tools/raw_time.py <lib.so> <n> <m>"""
import ctypes as C, sys
import numpy as np
lib = C.CDLL(sys.argv[1]); n = int(sys.argv[2]); m = int(sys.argv[3])
vp = C.c_void_p; dp = C.POINTER(C.c_double); ip = C.POINTER(C.c_int)
h = vp()
lib.nngp_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
assert lib.nngp_create(0, None, C.byref(h)) == 0
rng = np.random.default_rng(1)
c = rng.random((n, 2)) * np.sqrt(n / 5000.0); c = c[np.argsort(c[:, 0], kind="stable")]
cf = np.ascontiguousarray(c.T).ravel()
X = np.ascontiguousarray(np.column_stack([np.ones(n), rng.standard_normal(n)]).T).ravel(); y = rng.standard_normal(n)
nI = int((1 + m) / 2 * m + (n - m - 1) * m)
idx = np.zeros(nI, dtype=np.int32); dist = np.zeros(nI); lu = np.zeros(2 * n, dtype=np.int32); ties = C.c_int(0)
lib.nngp_mk_nn_indx.argtypes = [vp, C.c_int, C.c_int, dp, C.c_int, ip, dp, ip, ip]
assert lib.nngp_mk_nn_indx(h, n, m, cf.ctypes.data_as(dp), 1, idx.ctypes.data_as(ip), dist.ctypes.data_as(dp), lu.ctypes.data_as(ip), C.byref(ties)) == 0
lib.nngp_set_data.argtypes = [vp, C.c_int, C.c_int, C.c_int, dp, ip, ip, dp, dp, C.c_int, C.c_int]
assert lib.nngp_set_data(h, n, m, 2, cf.ctypes.data_as(dp), idx.ctypes.data_as(ip), lu.ctypes.data_as(ip), X.ctypes.data_as(dp), y.ctypes.data_as(dp), 0, n) == 0
lib.nngp_time_bf_loglik.argtypes = [vp, C.c_int, C.c_int, dp, dp, C.c_int, C.c_int, C.POINTER(C.c_float)]
beta = np.array([1.0, 5.0])
for name, cov in (("exponential", 0), ("spherical", 1), ("matern", 2), ("gaussian", 3)):
    th = np.array([5.0, 1.0, 6.0, 0.75]); ms = C.c_float(0)
    for it in (3, 20, 20):
        assert lib.nngp_time_bf_loglik(h, cov, 1, th.ctypes.data_as(dp), beta.ctypes.data_as(dp), it, 0, C.byref(ms)) == 0
    print(f"{name} {ms.value:.4f} ms", end="; ")
print()
