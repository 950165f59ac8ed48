"""hypermapper_b200: B200-native (sm_100a) surrogate prediction + acquisition scoring hot path of HyperMapper.

Python host code mirrors the reference's module-level API for that path (models / random_scalarizations /
local_search / compute_pareto); the arithmetic runs in hand-written CUDA kernels behind the C-ABI declared in
include/hm_b200.h.  There is no CPU fallback.
"""
from . import _lib  # noqa: F401
from ._lib import HmError, build  # noqa: F401

__all__ = ["HmError", "build"]
