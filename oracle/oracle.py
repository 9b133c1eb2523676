"""TEST INFRASTRUCTURE - ctypes front end of the C oracle (oracle/dpr_oracle.c).

The oracle restates the reference's CPU algorithm (src/raster.jl:5-108, src/raster_pullback.jl:2-160 under
/root/reference).  It is the checker for the CUDA library, never a product path: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.

Array convention (same as the product's host mirror): every array has the shape the reference's nd-array
flavour uses (docs/src/batch.md) - points (N_in, P), rotation (N_out, N_in, B), translation (N_out, B),
out / ds_dout (g_1..g_n, B), background / out_weight (B,), point_weight (P,) - and is handed to C in
column-major (Fortran) order, i.e. Julia's memory layout.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import NamedTuple, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libdpr_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with the committed recipe (oracle/Makefile)."""
    src_mtime = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("dpr_oracle.c", "dpr_oracle_impl.inc", "Makefile"))
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < src_mtime:
        subprocess.run(["make", "-C", _HERE, "-B", "libdpr_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.dpro_max_threads.restype = ctypes.c_int
    return _lib


def max_threads() -> int:
    return int(_load().dpro_max_threads())


class PullbackResult(NamedTuple):
    """Same field order as the reference's NamedTuple (src/raster_pullback.jl:140-147)."""
    points: np.ndarray
    rotation: np.ndarray
    translation: np.ndarray
    background: np.ndarray
    out_weight: np.ndarray
    point_weight: np.ndarray


def _f(a, dtype):
    return np.asfortranarray(np.asarray(a, dtype=dtype))


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _dims(points, rotation, translation):
    n_in, P = points.shape
    n_out, n_in_r, B = rotation.shape
    if n_in_r != n_in:
        raise ValueError("rotation/points dimension mismatch")
    if translation.shape != (n_out, B):
        raise ValueError("translation shape mismatch")
    return n_in, n_out, P, B


def raster(grid_size: Sequence[int], points, rotation, translation, background=None, out_weight=None,
           point_weight=None, *, dtype=np.float64, n_threads: int = 1, f64_accumulate: bool = False) -> np.ndarray:
    """Batched forward (src/raster.jl:5-66).  Returns out of shape (*grid_size, B), Fortran order."""
    lib = _load()
    dtype = np.dtype(dtype)
    points, rotation, translation = _f(points, dtype), _f(rotation, dtype), _f(translation, dtype)
    n_in, n_out, P, B = _dims(points, rotation, translation)
    if len(grid_size) != n_out:
        raise ValueError("grid_size/rotation dimension mismatch")
    bg = None if background is None else _f(background, dtype)
    ow = None if out_weight is None else _f(out_weight, dtype)
    pw = None if point_weight is None else _f(point_weight, dtype)
    for a, n in ((bg, B), (ow, B), (pw, P)):
        if a is not None and a.shape != (n,):
            raise ValueError("weight/background length mismatch")
    out = np.empty(tuple(grid_size) + (B,), dtype=dtype, order="F")
    grid = (ctypes.c_int64 * n_out)(*grid_size)
    fn = lib.dpro_raster_f32 if dtype == np.float32 else lib.dpro_raster_f64
    fn.restype = ctypes.c_int
    rc = fn(ctypes.c_int(n_in), ctypes.c_int(n_out), grid, ctypes.c_int64(P), ctypes.c_int64(B), _ptr(points),
            _ptr(rotation), _ptr(translation), _ptr(bg), _ptr(ow), _ptr(pw), _ptr(out), ctypes.c_int(n_threads),
            ctypes.c_int(int(f64_accumulate)))
    if rc != 0:
        raise RuntimeError(f"oracle raster failed rc={rc}")
    return out


def raster_pullback(ds_dout, points, rotation, translation, background=None, out_weight=None, point_weight=None, *,
                    dtype=np.float64, n_slabs: int = 1, f64_accumulate: bool = False) -> PullbackResult:
    """Batched pullback (src/raster_pullback.jl:85-148).  n_slabs = size(ds_dpoints, 3) of the reference
    (= min(B, nthreads()), src/interface.jl:402-412); it is also the number of OpenMP threads used."""
    lib = _load()
    dtype = np.dtype(dtype)
    points, rotation, translation = _f(points, dtype), _f(rotation, dtype), _f(translation, dtype)
    ds_dout = _f(ds_dout, dtype)
    n_in, n_out, P, B = _dims(points, rotation, translation)
    if ds_dout.ndim != n_out + 1 or ds_dout.shape[-1] != B:
        raise ValueError("ds_dout shape mismatch")
    ow = None if out_weight is None else _f(out_weight, dtype)
    pw = None if point_weight is None else _f(point_weight, dtype)
    grid_size = ds_dout.shape[:-1]
    grid = (ctypes.c_int64 * n_out)(*grid_size)
    res = PullbackResult(
        points=np.zeros((n_in, P), dtype=dtype, order="F"),
        rotation=np.zeros((n_out, n_in, B), dtype=dtype, order="F"),
        translation=np.zeros((n_out, B), dtype=dtype, order="F"),
        background=np.zeros((B,), dtype=dtype),
        out_weight=np.zeros((B,), dtype=dtype),
        point_weight=np.zeros((P,), dtype=dtype),
    )
    fn = lib.dpro_raster_pullback_f32 if dtype == np.float32 else lib.dpro_raster_pullback_f64
    fn.restype = ctypes.c_int
    rc = fn(ctypes.c_int(n_in), ctypes.c_int(n_out), grid, ctypes.c_int64(P), ctypes.c_int64(B), _ptr(points),
            _ptr(rotation), _ptr(translation), _ptr(ow), _ptr(pw), _ptr(ds_dout), _ptr(res.points),
            _ptr(res.rotation), _ptr(res.translation), _ptr(res.background), _ptr(res.out_weight),
            _ptr(res.point_weight), ctypes.c_int(n_slabs), ctypes.c_int(int(f64_accumulate)))
    if rc != 0:
        raise RuntimeError(f"oracle raster_pullback failed rc={rc}")
    return res


def voxel_shifts(n: int) -> np.ndarray:
    """voxel_shifts(Val(n)), src/util.jl:26-27; returns (2^n, n) int64."""
    out = np.zeros((1 << n, n), dtype=np.int64)
    _load().dpro_voxel_shifts(ctypes.c_int(n), out.ctypes.data_as(ctypes.c_void_p))
    return out
