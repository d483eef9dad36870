"""ctypes loader for libnngp_b200.so (the C ABI declared in include/nngp_b200.h).

The product path has no CPU fallback: if the CUDA library is missing or no B200-class device is
visible, calls fail loudly (RuntimeError) instead of computing anything on the host.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# NNGP_B200_LIB: an alternative build of the same library (kernel A/B experiments, tools/ab_bench.py)
LIB_PATH = os.environ.get("NNGP_B200_LIB") or os.path.join(_HERE, "libnngp_b200.so")
RSHIM_PATH = os.path.join(_HERE, "librshim.so")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
RNORM_FN = C.CFUNCTYPE(C.c_double, C.c_double, C.c_double)
RUNIF_FN = C.CFUNCTYPE(C.c_double, C.c_double, C.c_double)
PRINT_FN = C.CFUNCTYPE(None, C.c_char_p)
INTERRUPT_FN = C.CFUNCTYPE(C.c_int)  # nonzero: stop (the call returns NNGP_EINTR)
ALLREDUCE_FN = C.CFUNCTYPE(None, c_double_p, C.c_int, C.c_void_p)


class RnngpArgs(C.Structure):
    """Mirror of nngp_rnngp_args (include/nngp_b200.h)."""
    _fields_ = [
        ("cov_model", C.c_int),
        ("sigmaSqIG", C.c_double * 2), ("tauSqIG", C.c_double * 2), ("phiUnif", C.c_double * 2),
        ("nuUnif", C.c_double * 2),
        ("betaStarting", c_double_p),
        ("sigmaSqStarting", C.c_double), ("tauSqStarting", C.c_double), ("phiStarting", C.c_double),
        ("nuStarting", C.c_double),
        ("sigmaSqTuning", C.c_double), ("tauSqTuning", C.c_double), ("phiTuning", C.c_double),
        ("nuTuning", C.c_double),
        ("nSamples", C.c_int), ("mode", C.c_int), ("verbose", C.c_int), ("nReport", C.c_int),
        ("rnorm", RNORM_FN), ("runif", RUNIF_FN), ("print", PRINT_FN), ("interrupt", INTERRUPT_FN),
        ("allreduce", ALLREDUCE_FN), ("allreduce_user", C.c_void_p),
        ("repIndx", C.POINTER(C.c_int)), ("repSamples", c_double_p),
    ]


# every symbol include/nngp_b200.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "nngp_abi_version", "nngp_device_count", "nngp_create", "nngp_destroy", "nngp_last_error", "nngp_sync",
    "nngp_mk_nn_indx", "nngp_mk_u_indx", "nngp_nn_query", "nngp_set_data", "nngp_bf_loglik", "nngp_bf_loglik_cur", "nngp_bf_loglik_vec", "nngp_q_resid", "nngp_q_forms",
    "nngp_get_bf", "nngp_rnngp", "nngp_rep_draw", "nngp_rnngp_replicated", "nngp_cnngp", "nngp_cnngp_score", "nngp_conj_kfold", "nngp_rnngp_predict", "nngp_time_bf_loglik", "nngp_launch_count", "nngp_set_option", "nngp_get_option",
    "nngp_fp64_peak", "nngp_host_alloc", "nngp_host_free", "nngp_comm_unique_id", "nngp_comm_init", "nngp_comm_init_local", "nngp_comm_allreduce_sum", "nngp_comm_destroy",
]

_lib = None
_rshim = None


def load():
    """Load libnngp_b200.so; raises RuntimeError with build instructions when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). spnngp_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.nngp_abi_version.restype = C.c_int
    L.nngp_device_count.restype = C.c_int
    L.nngp_create.restype = C.c_int
    L.nngp_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
    L.nngp_destroy.restype = None
    L.nngp_destroy.argtypes = [vp]
    L.nngp_last_error.restype = C.c_char_p
    L.nngp_last_error.argtypes = [vp]
    L.nngp_sync.argtypes = [vp]
    L.nngp_set_data.argtypes = [vp, C.c_int, C.c_int, C.c_int, c_double_p, c_int_p, c_int_p, c_double_p, c_double_p,
                                C.c_int, C.c_int]
    L.nngp_bf_loglik.argtypes = [vp, C.c_int, C.c_int, c_double_p, c_double_p, c_int_p, c_double_p]
    L.nngp_bf_loglik_cur.argtypes = [vp, C.c_int, c_double_p, c_double_p, C.c_int, C.c_int, c_double_p]
    L.nngp_bf_loglik_vec.argtypes = [vp, C.c_int, C.c_int, c_double_p, c_double_p, c_int_p, c_double_p]
    L.nngp_q_resid.argtypes = [vp, C.c_int, c_double_p, c_double_p]
    L.nngp_q_forms.argtypes = [vp, C.c_int, c_double_p]
    L.nngp_get_bf.argtypes = [vp, C.c_int, c_double_p, c_double_p]
    L.nngp_rnngp.argtypes = [vp, C.POINTER(RnngpArgs), c_double_p, c_double_p, c_int_p]
    L.nngp_mk_u_indx.argtypes = [vp, C.c_int, C.c_int, c_int_p, C.c_int, c_int_p, c_int_p, c_int_p]
    L.nngp_rep_draw.argtypes = [vp, C.c_int, c_double_p, c_double_p, c_double_p]
    L.nngp_rnngp_replicated.argtypes = [vp, C.c_int, c_double_p, c_double_p, C.c_int, RNORM_FN, INTERRUPT_FN, c_double_p]
    L.nngp_time_bf_loglik.argtypes = [vp, C.c_int, C.c_int, c_double_p, c_double_p, C.c_int, C.c_int,
                                      C.POINTER(C.c_float)]
    L.nngp_fp64_peak.argtypes = [vp, c_double_p]
    L.nngp_host_alloc.argtypes = [vp, C.c_size_t, C.POINTER(C.c_void_p)]
    L.nngp_host_free.argtypes = [vp, C.c_void_p]
    L.nngp_comm_unique_id.argtypes = [C.c_char_p]
    L.nngp_comm_init.argtypes = [vp, C.c_char_p, C.c_int, C.c_int]
    L.nngp_comm_init_local.argtypes = [vp, C.c_char_p, C.c_int, C.c_int]
    L.nngp_comm_allreduce_sum.argtypes = [vp, c_double_p, C.c_int]
    L.nngp_comm_destroy.restype = None
    L.nngp_comm_destroy.argtypes = [vp]
    L.nngp_set_option.argtypes = [vp, C.c_char_p, C.c_int]
    L.nngp_get_option.argtypes = [vp, C.c_char_p, c_int_p]
    L.nngp_launch_count.restype = C.c_longlong
    L.nngp_launch_count.argtypes = [vp]
    if hasattr(L, "nngp_mk_nn_indx"):
        L.nngp_mk_nn_indx.argtypes = [vp, C.c_int, C.c_int, c_double_p, C.c_int, c_int_p, c_double_p, c_int_p, c_int_p]
    if hasattr(L, "nngp_nn_query"):
        L.nngp_nn_query.argtypes = [vp, C.c_int, c_double_p, C.c_int, c_double_p, C.c_int, c_int_p]
    if hasattr(L, "nngp_cnngp"):
        L.nngp_cnngp.argtypes = [vp, C.c_int, C.c_int, c_double_p, c_double_p, C.c_int, c_double_p, c_double_p,
                                 c_int_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]
    L.nngp_cnngp_score.argtypes = [vp, C.c_int, C.c_int, c_double_p, c_double_p, C.c_int, c_double_p, c_double_p, c_int_p,
                                   c_double_p, c_double_p, c_double_p]
    L.nngp_conj_kfold.argtypes = [vp, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, C.c_int, C.c_int,
                                  c_double_p, c_double_p, C.c_int, C.c_int, C.c_int, c_int_p, C.c_int, C.c_int, c_double_p]
    if hasattr(L, "nngp_rnngp_predict"):
        L.nngp_rnngp_predict.argtypes = [vp, C.c_int, C.c_int, c_double_p, c_double_p, c_int_p, C.c_int, c_double_p,
                                         c_double_p, c_double_p, c_double_p, C.c_int, C.c_int]
    _lib = L
    return L


def load_rshim():
    """R-compatible RNG (Mersenne-Twister + inversion) used by the Python host layer."""
    global _rshim
    if _rshim is not None:
        return _rshim
    if not os.path.exists(RSHIM_PATH):
        raise RuntimeError(f"{RSHIM_PATH} is missing: run __graft_entry__.build()")
    R = C.CDLL(RSHIM_PATH)
    R.rshim_set_seed.argtypes = [C.c_uint]
    R.rnorm.restype = C.c_double
    R.rnorm.argtypes = [C.c_double, C.c_double]
    R.runif.restype = C.c_double
    R.runif.argtypes = [C.c_double, C.c_double]
    R.norm_rand.restype = C.c_double
    R.unif_rand.restype = C.c_double
    _rshim = R
    return R
