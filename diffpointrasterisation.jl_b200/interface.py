"""Host-side mirror of the reference's public interface for the hot path, over the C ABI of libdpr.so.

Mirrors (names, argument order and meaning, defaults, error behaviour) /root/reference:
  raster            src/interface.jl:62-77      -> raster(grid_size, points, rotation, translation, [background,
                                                   out_weight, point_weight])
  raster!           src/interface.jl:87-129     -> raster_(out, ...)            (trailing underscore = Julia's `!`)
  raster_pullback!  src/interface.jl:196-308    -> raster_pullback_(ds_dout, ...; points=..., rotation=..., ...)
and the batched canonical methods they end in (src/raster.jl:5-34, ext/DiffPointRasterisationCUDAExt.jl:231-321).

Arrays are torch CUDA tensors with the SHAPES of the reference's nd-array flavour (docs/src/batch.md) and
column-major (Fortran) strides, i.e. exactly Julia's memory:
  points (N_in, P), rotation (N_out, N_in, B), translation (N_out, B), background / out_weight (B,),
  point_weight (P,), out / ds_dout (g_1..g_n, B).
Tensors with other strides are copied into that layout (`fortran`).  PyTorch is used for device memory and
streams only; all arithmetic happens in the CUDA library.  There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

from typing import NamedTuple, Optional, Sequence

import torch

from . import _lib


class DimensionMismatch(ValueError):
    """Julia's DimensionMismatch (raised by the @argcheck's of src/raster.jl:14-23)."""


class PullbackResult(NamedTuple):
    """Field order of the reference's NamedTuple (ext/DiffPointRasterisationCUDAExt.jl:313-320)."""
    points: torch.Tensor
    rotation: torch.Tensor
    translation: torch.Tensor
    background: torch.Tensor
    out_weight: torch.Tensor
    point_weight: torch.Tensor


def empty_f(shape: Sequence[int], dtype: torch.dtype, device) -> torch.Tensor:
    """Uninitialised tensor of `shape` with column-major strides (Julia `similar`)."""
    shape = tuple(int(s) for s in shape)
    return torch.empty(shape[::-1], dtype=dtype, device=device).permute(*range(len(shape) - 1, -1, -1))


def is_fortran(t: torch.Tensor) -> bool:
    expected = 1
    for size, stride in zip(t.shape, t.stride()):
        if size != 1 and stride != expected:
            return False
        expected *= size
    return True


def fortran(t: torch.Tensor) -> torch.Tensor:
    """Return `t` in column-major memory (no copy if it already is)."""
    if is_fortran(t):
        return t
    out = empty_f(t.shape, t.dtype, t.device)
    out.copy_(t)
    return out


def _suffix(dtype: torch.dtype) -> str:
    if dtype == torch.float32:
        return "f32"
    if dtype == torch.float64:
        return "f64"
    raise TypeError(f"unsupported element type {dtype}: libdpr computes in Float32 or Float64")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _canonical(points, rotation, translation, background, out_weight, point_weight, dtype, device, batch_from=None):
    """Argument checks of the canonical method (src/raster.jl:14-23; ext/...CUDAExt.jl:247-262)."""
    for name, t in (("points", points), ("rotation", rotation), ("translation", translation)):
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name} must be a torch.Tensor")
        if not t.is_cuda:
            raise RuntimeError(f"{name} is not a CUDA tensor: this library has no CPU path (use the reference's CPU methods)")
    if points.dim() != 2 or rotation.dim() != 3 or translation.dim() != 2:
        raise DimensionMismatch("expected points (N_in, P), rotation (N_out, N_in, B), translation (N_out, B)")
    n_in, P = points.shape
    n_out, n_in_r, B = rotation.shape
    if n_in_r != n_in:
        raise DimensionMismatch(f"rotation is {n_out}x{n_in_r} but points are {n_in}-dimensional")
    if translation.shape[0] != n_out:
        # src/interface.jl:137-162: "Dimension of translation (..) and number of rows of rotation (..) do not match"
        raise DimensionMismatch(f"Dimension of translation ({translation.shape[0]}) and number of rows of rotation ({n_out}) do not match")
    if translation.shape[1] != B:
        raise DimensionMismatch("batch size of rotation and translation differ")
    for name, t, n in (("background", background, B), ("out_weight", out_weight, B), ("point_weight", point_weight, P)):
        if t is not None and tuple(t.shape) != (n,):
            raise DimensionMismatch(f"length of {name} is {tuple(t.shape)}, expected ({n},)")
    if batch_from is not None and batch_from != B:
        raise DimensionMismatch(f"batch dimension of the image array ({batch_from}) and of the poses ({B}) differ")
    conv = lambda t: None if t is None else fortran(t.to(device=device, dtype=dtype))
    return (n_in, n_out, int(P), int(B), conv(points), conv(rotation), conv(translation), conv(background),
            conv(out_weight), conv(point_weight))


def _grid_array(grid_size):
    import ctypes
    return (ctypes.c_int64 * len(grid_size))(*[int(g) for g in grid_size])


_workspaces = {}


def _workspace(op: int, n_in, n_out, grid, P, B, dtype, device):
    """Scratch buffer of dpr_workspace_bytes() for one call, cached per device and stream and SHARED by raster and
    raster_pullback! (the glue owns it, like the CuVector{UInt8} a Julia caller would allocate): with
    DPR_OPT_BINNING_CACHE the pullback that follows a forward on the same inputs - the rrule - reuses the point bins the
    forward left there.  The first 256 bytes are zeroed on allocation, as that option's contract asks."""
    lib = _lib.load()
    need = int(lib.dpr_workspace_bytes(op, n_in, n_out, grid, P, B, 4 if dtype == torch.float32 else 8))
    if need == 0:
        return None, 0
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < need:
        buf = torch.empty(max(need, 4096), dtype=torch.uint8, device=device)
        buf[:256].zero_()
        _workspaces[key] = buf
    return buf, buf.numel()


def _stream_handle(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _scalar_to_vec(v, dtype, device):
    """Single-image scalars (background::Number, out_weight::Number) -> length-1 vectors (src/interface.jl:100-120)."""
    if v is None:
        return None
    if isinstance(v, torch.Tensor):
        return v.reshape(1).to(device=device, dtype=dtype)
    return torch.tensor([float(v)], dtype=dtype, device=device)


def _is_single_image(rotation) -> bool:
    return isinstance(rotation, torch.Tensor) and rotation.dim() == 2


def raster_(out: torch.Tensor, points, rotation, translation, background=None, out_weight=None, point_weight=None):
    """`raster!(out, points, rotation, translation, [background, out_weight, point_weight])`.

    Batched when `rotation` is (N_out, N_in, B); single image when it is a matrix (N_out, N_in) - then `translation` is
    (N_out,), `background` / `out_weight` are numbers and `out` has no batch axis: exactly the wrapper of
    src/interface.jl:100-120 (a batch of one with a singleton trailing dimension, src/util.jl:98-102).

    `out` is (g_1..g_n, B) column-major and is overwritten (src/raster.jl:27).  Missing optional arguments are the
    reference's defaults: background 0, out_weight 1, point_weight 1 (src/interface.jl:87-92, :368-394).
    """
    if not out.is_cuda:
        raise RuntimeError("out is not a CUDA tensor: this library has no CPU path")
    if not is_fortran(out):
        raise ValueError("out must have column-major strides (use empty_f)")
    if _is_single_image(rotation):
        if translation.dim() != 1 or translation.shape[0] != rotation.shape[0]:
            raise DimensionMismatch(f"Dimension of translation ({tuple(translation.shape)}) and number of rows of rotation ({rotation.shape[0]}) do not match")
        raster_(out.unsqueeze(-1), points, rotation.unsqueeze(-1), translation.unsqueeze(-1),
                _scalar_to_vec(background, out.dtype, out.device), _scalar_to_vec(out_weight, out.dtype, out.device), point_weight)
        return out
    dtype, device = out.dtype, out.device
    suf = _suffix(dtype)
    n_out_p1 = out.dim()
    (n_in, n_out, P, B, points, rotation, translation, background, out_weight, point_weight) = _canonical(
        points, rotation, translation, background, out_weight, point_weight, dtype, device, batch_from=out.shape[-1])
    if n_out != n_out_p1 - 1:
        raise DimensionMismatch(f"out has {n_out_p1 - 1} grid dimensions but rotation has {n_out} rows")  # src/raster.jl:14
    lib = _lib.load()
    with torch.cuda.device(device):
        fn = getattr(lib, f"dpr_raster_forward_{suf}")
        garr = _grid_array(out.shape[:-1])
        ws, ws_bytes = _workspace(_lib.OP_FORWARD, n_in, n_out, garr, P, B, dtype, device)
        rc = fn(n_in, n_out, garr, P, B, _ptr(points), _ptr(rotation), _ptr(translation),
                _ptr(background), _ptr(out_weight), _ptr(point_weight), _ptr(out), _ptr(ws), ws_bytes,
                _stream_handle(device))
    _lib.check(rc)
    return out


def raster(grid_size: Sequence[int], points, rotation, translation, background=None, out_weight=None,
           point_weight=None) -> torch.Tensor:
    """`raster(grid_size, points, rotation, translation, [background, out_weight, point_weight])`, batched.

    Allocates `out` like src/interface.jl:62-77 (element type = promotion of the arguments' types, on the device of
    `points`) and calls raster_.  Returns (g_1..g_n, B), column-major.
    """
    if not isinstance(points, torch.Tensor):
        raise TypeError("points must be a torch.Tensor")
    dtype = points.dtype
    for t in (rotation, translation, background, out_weight, point_weight):
        if isinstance(t, torch.Tensor):
            dtype = torch.promote_types(dtype, t.dtype)
    if rotation.dim() == 2:     # single image (src/interface.jl:100-120)
        out = empty_f(tuple(grid_size), dtype, points.device)
        return raster_(out, points, rotation, translation, background, out_weight, point_weight)
    if rotation.dim() != 3:
        raise DimensionMismatch("rotation must be (N_out, N_in, B) or, for a single image, (N_out, N_in)")
    B = rotation.shape[-1]
    out = empty_f(tuple(grid_size) + (B,), dtype, points.device)
    return raster_(out, points, rotation, translation, background, out_weight, point_weight)


def raster_pullback_(ds_dout: torch.Tensor, points, rotation, translation, background=None, out_weight=None,
                     point_weight=None, *, points_out=None, rotation_out=None, translation_out=None,
                     background_out=None, out_weight_out=None, point_weight_out=None) -> PullbackResult:
    """`raster_pullback!(ds_dout, points, rotation, translation, [background, out_weight, point_weight]; kwargs...)`.

    Returns the sensitivities with respect to (points, rotation, translation, background, out_weight, point_weight)
    in the reference's order.  The keyword arguments are the pre-allocated outputs of src/interface.jl:278-291
    (`points=`, `rotation=`, ... there; suffixed `_out` here because the names are taken by the inputs).
    Default allocations follow the CuArray overrides ext/DiffPointRasterisationCUDAExt.jl:323-333:
    d_points is (N_in, P), d_point_weight is (P,).
    `background` does not enter the gradients (its own gradient is sum(ds_dout), src/raster_pullback.jl:78).
    """
    if not ds_dout.is_cuda:
        raise RuntimeError("ds_dout is not a CUDA tensor: this library has no CPU path")
    if _is_single_image(rotation):
        # Single image.  The reference's CUDA extension only has an error stub for it
        # (ext/DiffPointRasterisationCUDAExt.jl:213-228); the batch kernels handle B = 1 by splitting the points over
        # the SMs, so route it through a batch of one and drop the batch axis of the per-pose results
        # (same field meaning as the CPU method src/raster_pullback.jl:74-81).
        if any(v is not None for v in (rotation_out, translation_out, background_out, out_weight_out)):
            raise ValueError("pre-allocated per-pose outputs are only supported in batch mode")
        res = raster_pullback_(ds_dout.unsqueeze(-1), points, rotation.unsqueeze(-1), translation.unsqueeze(-1),
                               _scalar_to_vec(background, ds_dout.dtype, ds_dout.device),
                               _scalar_to_vec(out_weight, ds_dout.dtype, ds_dout.device), point_weight,
                               points_out=points_out, point_weight_out=point_weight_out)
        return PullbackResult(res.points, res.rotation[..., 0], res.translation[..., 0], res.background[0],
                              res.out_weight[0], res.point_weight)
    dtype, device = ds_dout.dtype, ds_dout.device
    for t in (points, rotation, translation, out_weight, point_weight):
        if isinstance(t, torch.Tensor):
            dtype = torch.promote_types(dtype, t.dtype)      # promote_type, ext/...CUDAExt.jl:246
    suf = _suffix(dtype)
    ds_dout = fortran(ds_dout.to(dtype))
    (n_in, n_out, P, B, points, rotation, translation, background, out_weight, point_weight) = _canonical(
        points, rotation, translation, background, out_weight, point_weight, dtype, device, batch_from=ds_dout.shape[-1])
    if n_out != ds_dout.dim() - 1:
        raise DimensionMismatch(f"ds_dout has {ds_dout.dim() - 1} grid dimensions but rotation has {n_out} rows")

    def out_buf(given, shape, name):
        if given is None:
            return empty_f(shape, dtype, device)
        if tuple(given.shape) != tuple(shape) or given.dtype != dtype or not given.is_cuda or not is_fortran(given):
            raise DimensionMismatch(f"pre-allocated {name} must be a column-major CUDA tensor of shape {tuple(shape)} and dtype {dtype}")
        return given

    res = PullbackResult(
        points=out_buf(points_out, (n_in, P), "points"),
        rotation=out_buf(rotation_out, (n_out, n_in, B), "rotation"),
        translation=out_buf(translation_out, (n_out, B), "translation"),
        background=out_buf(background_out, (B,), "background"),
        out_weight=out_buf(out_weight_out, (B,), "out_weight"),
        point_weight=out_buf(point_weight_out, (P,), "point_weight"),
    )
    lib = _lib.load()
    with torch.cuda.device(device):
        fn = getattr(lib, f"dpr_raster_pullback_{suf}")
        garr = _grid_array(ds_dout.shape[:-1])
        ws, ws_bytes = _workspace(_lib.OP_PULLBACK, n_in, n_out, garr, P, B, dtype, device)
        rc = fn(n_in, n_out, garr, P, B, _ptr(ds_dout), _ptr(points), _ptr(rotation),
                _ptr(translation), _ptr(out_weight), _ptr(point_weight), _ptr(res.points), _ptr(res.rotation),
                _ptr(res.translation), _ptr(res.background), _ptr(res.out_weight), _ptr(res.point_weight), _ptr(ws),
                ws_bytes, _stream_handle(device))
    _lib.check(rc)
    return res
