"""ctypes binding of libdpr.so - the same symbols a Julia `ccall` layer binds (include/dpr.h).

There is no CPU fallback: if the library is missing it is built with nvcc; if that fails, or no sm_100 device is
present at call time, the call raises.
"""
from __future__ import annotations

import ctypes
import os

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdpr.so")

# every symbol include/dpr.h declares (tests/test_abi.py checks the header against this list and the .so)
SYMBOLS = [
    "dpr_version", "dpr_status_string", "dpr_last_error_message", "dpr_workspace_bytes",
    "dpr_raster_forward_f32", "dpr_raster_forward_f64", "dpr_raster_pullback_f32", "dpr_raster_pullback_f64",
    "dpr_raster_forward_host_f32", "dpr_raster_forward_host_f64", "dpr_raster_pullback_host_f32",
    "dpr_raster_pullback_host_f64", "dpr_host_alloc", "dpr_host_free", "dpr_host_release",
    "dpr_raster_forward_host_async_f32", "dpr_raster_forward_host_async_f64", "dpr_raster_pullback_host_async_f32",
    "dpr_raster_pullback_host_async_f64", "dpr_host_wait",
    "dpr_set_option", "dpr_get_option", "dpr_kernel_launch_count", "dpr_last_path",
    "dpr_profile_enable", "dpr_profile_count", "dpr_profile_get",
    "dpr_comm_unique_id", "dpr_comm_init_rank", "dpr_comm_destroy", "dpr_comm_allreduce_sum_f32", "dpr_comm_allreduce_sum_f64", "dpr_comm_uses_peer_memory",
]

OPT_FORWARD_ALGO, OPT_PULLBACK_ALGO, OPT_TILE_SMEM_BYTES, OPT_POINT_SPLIT, OPT_POSE_CHUNK, OPT_FORWARD_ACCUM, OPT_POINT_SORT, OPT_TILE3D_TMA, OPT_BINNING_CACHE, OPT_COMM_P2P = range(10)
OP_FORWARD, OP_PULLBACK = 0, 1

_lib = None


class DprError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libdpr status {status}: {message}")
        self.status = status


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    _build.build()          # no-op when libdpr.so is newer than every source it is built from (a stale binary would be
    lib = ctypes.CDLL(LIB_PATH)     # loaded silently against a newer header / SYMBOLS list otherwise)
    c_i, c_i64, c_p, c_sz = ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_size_t
    lib.dpr_version.restype = c_i
    lib.dpr_status_string.restype = ctypes.c_char_p
    lib.dpr_status_string.argtypes = [c_i]
    lib.dpr_last_error_message.restype = ctypes.c_char_p
    lib.dpr_workspace_bytes.restype = c_sz
    lib.dpr_workspace_bytes.argtypes = [c_i, c_i, c_i, c_p, c_i64, c_i64, c_i]
    head = [c_i, c_i, c_p, c_i64, c_i64]
    for suf in ("f32", "f64"):
        f = getattr(lib, f"dpr_raster_forward_{suf}")
        f.restype = c_i
        f.argtypes = head + [c_p] * 7 + [c_p, c_sz, c_p]
        f = getattr(lib, f"dpr_raster_pullback_{suf}")
        f.restype = c_i
        f.argtypes = head + [c_p] * 12 + [c_p, c_sz, c_p]
        f = getattr(lib, f"dpr_raster_forward_host_{suf}")
        f.restype = c_i
        f.argtypes = head + [c_p] * 7
        f = getattr(lib, f"dpr_raster_pullback_host_{suf}")
        f.restype = c_i
        f.argtypes = head + [c_p] * 12
        f = getattr(lib, f"dpr_raster_forward_host_async_{suf}")
        f.restype = c_i
        f.argtypes = head + [c_p] * 7 + [ctypes.POINTER(c_p)]
        f = getattr(lib, f"dpr_raster_pullback_host_async_{suf}")
        f.restype = c_i
        f.argtypes = head + [c_p] * 12 + [ctypes.POINTER(c_p)]
    lib.dpr_host_wait.restype = c_i
    lib.dpr_host_wait.argtypes = [c_p]
    lib.dpr_host_alloc.restype = c_i
    lib.dpr_host_alloc.argtypes = [ctypes.POINTER(c_p), c_sz]
    lib.dpr_host_free.restype = c_i
    lib.dpr_host_free.argtypes = [c_p]
    lib.dpr_host_release.restype = c_i
    lib.dpr_set_option.restype = c_i
    lib.dpr_set_option.argtypes = [c_i, c_i64]
    lib.dpr_get_option.restype = c_i64
    lib.dpr_get_option.argtypes = [c_i]
    lib.dpr_kernel_launch_count.restype = c_i64
    lib.dpr_profile_enable.restype = c_i
    lib.dpr_profile_enable.argtypes = [c_i]
    lib.dpr_profile_count.restype = c_i
    lib.dpr_profile_get.restype = c_i
    lib.dpr_profile_get.argtypes = [c_i, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_float)]
    lib.dpr_comm_unique_id.restype = c_i
    lib.dpr_comm_unique_id.argtypes = [c_p]
    lib.dpr_comm_init_rank.restype = c_i
    lib.dpr_comm_init_rank.argtypes = [ctypes.POINTER(c_p), c_i, c_i, c_p]
    lib.dpr_comm_destroy.restype = c_i
    lib.dpr_comm_destroy.argtypes = [c_p]
    lib.dpr_comm_uses_peer_memory.restype = c_i
    lib.dpr_comm_uses_peer_memory.argtypes = [c_p]
    for suf in ("f32", "f64"):
        f = getattr(lib, f"dpr_comm_allreduce_sum_{suf}")
        f.restype = c_i
        f.argtypes = [c_p, c_p, c_i64, c_p]
    lib.dpr_last_path.restype = ctypes.c_char_p
    lib.dpr_last_path.argtypes = [c_i]
    # The mirror owns its workspaces (interface._workspace: persistent per device and stream, header zeroed on allocation),
    # which is what DPR_OPT_BINNING_CACHE asks of a caller
    lib.dpr_set_option(OPT_BINNING_CACHE, 1)
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        lib = load()
        msg = lib.dpr_status_string(status).decode()
        detail = lib.dpr_last_error_message().decode()
        raise DprError(status, msg + (f" [{detail}]" if detail else ""))


def set_option(option: int, value: int) -> None:
    check(load().dpr_set_option(option, value))


def get_option(option: int) -> int:
    return int(load().dpr_get_option(option))


def kernel_launch_count() -> int:
    return int(load().dpr_kernel_launch_count())


def last_path(op: int) -> str:
    return load().dpr_last_path(op).decode()


def profile_enable(on: bool) -> None:
    check(load().dpr_profile_enable(1 if on else 0))


def profile_records():
    """[(kernel name, milliseconds)] for every launch since profile_enable(True)."""
    lib = load()
    out = []
    for i in range(lib.dpr_profile_count()):
        name, ms = ctypes.c_char_p(), ctypes.c_float()
        check(lib.dpr_profile_get(i, ctypes.byref(name), ctypes.byref(ms)))
        out.append((name.value.decode(), float(ms.value)))
    return out
