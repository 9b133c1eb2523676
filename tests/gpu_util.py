"""Helpers for the -m gpu parity tests: move NumPy fixtures to the device in the library's layout and back."""
import numpy as np
import torch

import dpr_b200


def to_dev(a, dtype=None):
    """NumPy array (any order) -> CUDA tensor of the same SHAPE with column-major strides."""
    if a is None:
        return None
    a = np.asarray(a)
    t = torch.from_numpy(np.ascontiguousarray(a)).to("cuda")
    if dtype is not None:
        t = t.to(dtype)
    return dpr_b200.fortran(t)


def to_np(t):
    return t.detach().cpu().numpy()


def dev_args(d, dtype):
    td = torch.float32 if np.dtype(dtype) == np.float32 else torch.float64
    return tuple(to_dev(d[k], td) for k in ("points", "rotation", "translation", "background", "out_weight", "point_weight"))


class forced:
    """Context manager forcing a library option for the duration of a test."""

    def __init__(self, **opts):
        self.opts = opts
        self.names = dict(forward_algo=dpr_b200._lib.OPT_FORWARD_ALGO, pullback_algo=dpr_b200._lib.OPT_PULLBACK_ALGO,
                          tile_smem_bytes=dpr_b200._lib.OPT_TILE_SMEM_BYTES, point_split=dpr_b200._lib.OPT_POINT_SPLIT,
                          pose_chunk=dpr_b200._lib.OPT_POSE_CHUNK, forward_accum=dpr_b200._lib.OPT_FORWARD_ACCUM, point_sort=dpr_b200._lib.OPT_POINT_SORT,
                          tile3d_tma=dpr_b200._lib.OPT_TILE3D_TMA, binning_cache=dpr_b200._lib.OPT_BINNING_CACHE)

    def __enter__(self):
        self.old = {k: dpr_b200.get_option(self.names[k]) for k in self.opts}
        for k, v in self.opts.items():
            dpr_b200.set_option(self.names[k], v)
        return self

    def __exit__(self, *exc):
        for k, v in self.old.items():
            dpr_b200.set_option(self.names[k], v)
