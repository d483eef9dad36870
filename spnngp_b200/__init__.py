"""spnngp_b200 -- B200-native (sm_100a) engine for the data-parallel hot path of spNNGP.

Product code lives in `csrc/` (CUDA kernels + the C ABI, built into libnngp_b200.so) and in the thin
host mirrors of the reference's interfaces (`engine.py`, `api.py`). There is no CPU fallback.
"""
from .engine import COV_MODELS, Engine, NNGPError  # noqa: F401

__all__ = ["Engine", "NNGPError", "COV_MODELS"]
