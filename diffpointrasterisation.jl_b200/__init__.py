"""B200-native (sm_100a) implementation of DiffPointRasterisation.jl's hot path: `raster` / `raster_pullback!`.

The product is the CUDA library `libdpr.so` (csrc/, C ABI in include/dpr.h).  This package is the thin host-side
mirror of the reference's interface used by the tests and the benchmark; Julia binds the same symbols with `ccall`
(see INTEGRATION.md and julia/).
"""
from . import _lib, autograd, build  # noqa: F401
from ._lib import DprError, kernel_launch_count, last_path, set_option, get_option  # noqa: F401
from .interface import (DimensionMismatch, PullbackResult, empty_f, fortran, is_fortran, raster, raster_,  # noqa: F401
                        raster_pullback_)

__all__ = ["raster", "raster_", "raster_pullback_", "PullbackResult", "DimensionMismatch", "DprError", "empty_f",
           "fortran", "is_fortran", "kernel_launch_count", "last_path", "set_option", "get_option"]
