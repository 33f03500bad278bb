"""riss-solver_b200: B200-native drop-in for Coprocessor's subsumption / strengthening / BVE hot path.

The product is the shared library libcp3b200.so (CUDA kernels for sm_100a + ordered-commit host engine +
C ABI, include/cp3b200.h).  This package is the thin host-side mirror of the reference's C interface on
top of it; importing it does not need a GPU, calling into the hot path does (no CPU fallback).
"""
from . import gen  # noqa: F401
from .coprocessor import Coprocessor, ScanService, nccl_unique_id, SUBSUME, STRENGTHEN  # noqa: F401
from ._lib import DEFAULT_LIB, CP_SYMBOLS, CP3G_SYMBOLS, load  # noqa: F401
