"""Pose-sharded multi-GPU driver (new; the reference is single-device).

Poses are independent units (src/raster_pullback.jl:117-138 already treats them so; the forward has no cross-pose
term), so the pose batch is split into contiguous ranges, one per rank (one process per GPU), with the points and
point weights replicated.  Per-pose outputs (`out`, d_rotation, d_translation, d_background, d_out_weight) stay
local to the rank that owns the pose.  The only cross-GPU dependency is the pose-sum in d_points
(src/raster_pullback.jl:141) and d_point_weight (:146): they are produced into ONE packed (N_in+1, P) buffer and
reduced with a single all-reduce (NCCL over NVLink on GPUs; gloo in the CPU tests).

`forward_fn` / `pullback_fn` default to the CUDA library; the CPU test-suite injects the oracle to exercise the
sharding, packing and collective logic under gloo with world_size 2.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

from . import interface
from .interface import PullbackResult


def pose_range(batch: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous pose range of `rank`: the first (batch % world_size) ranks get one extra pose."""
    base, rem = divmod(int(batch), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_poses(t: Optional[torch.Tensor], rank: int, world_size: int) -> Optional[torch.Tensor]:
    """Slice the trailing (batch) axis: contiguous in column-major memory, so a shard is a pointer offset."""
    if t is None:
        return None
    lo, hi = pose_range(t.shape[-1], rank, world_size)
    return t[..., lo:hi]


class DprComm:
    """The library's own communicator (dpr_comm_*, include/dpr.h): NCCL resolved inside libdpr.so.  torch.distributed
    is only used to ship the 128-byte unique id from rank 0 to the other ranks."""

    def __init__(self, device, group: Optional[dist.ProcessGroup] = None):
        import ctypes
        from . import _lib
        self._lib = _lib
        self.lib = _lib.load()
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        ident = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (ctypes.c_ubyte * 128)()
            _lib.check(self.lib.dpr_comm_unique_id(ctypes.cast(buf, ctypes.c_void_p)))
            ident = torch.tensor(list(buf), dtype=torch.uint8)
        ident = ident.to(device)
        dist.broadcast(ident, src=0, group=group)
        raw = bytes(ident.cpu().tolist())
        self.handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(self.lib.dpr_comm_init_rank(ctypes.byref(self.handle), world, rank, raw))
        self.device = device
        self.uses_peer_memory = bool(self.lib.dpr_comm_uses_peer_memory(self.handle))

    def all_reduce_(self, t: torch.Tensor) -> None:
        fn = self.lib.dpr_comm_allreduce_sum_f32 if t.dtype == torch.float32 else self.lib.dpr_comm_allreduce_sum_f64
        with torch.cuda.device(t.device):
            self._lib.check(fn(self.handle, t.data_ptr(), t.numel(), torch.cuda.current_stream(t.device).cuda_stream))

    def close(self) -> None:
        if self.handle:
            self.lib.dpr_comm_destroy(self.handle)
            self.handle = None


class PoseShardedRaster:
    def __init__(self, group: Optional[dist.ProcessGroup] = None, forward_fn: Callable = None,
                 pullback_fn: Callable = None, comm: Optional[DprComm] = None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.forward_fn = forward_fn or interface.raster
        self.pullback_fn = pullback_fn
        self.comm = comm          # None: torch.distributed all_reduce; DprComm: the library's NCCL entry point
        self._packed = None

    # ---- forward: no collective --------------------------------------------------------------------------
    def raster(self, grid_size, points, rotation, translation, background=None, out_weight=None, point_weight=None):
        """`rotation`, `translation`, `background`, `out_weight` are this rank's pose shard; returns its `out` shard."""
        return self.forward_fn(grid_size, points, rotation, translation, background, out_weight, point_weight)

    # ---- pullback: one all-reduce of the packed pose-summed gradients ---------------------------------------
    def packed_buffer(self, n_in: int, P: int, dtype, device) -> torch.Tensor:
        if (self._packed is None or self._packed.numel() != (n_in + 1) * P or self._packed.dtype != dtype
                or self._packed.device != torch.device(device)):
            self._packed = torch.empty((n_in + 1) * P, dtype=dtype, device=device)
        return self._packed

    def raster_pullback_(self, ds_dout, points, rotation, translation, background=None, out_weight=None,
                         point_weight=None, async_op: bool = False):
        """Local pullback on this rank's pose shard, then all-reduce(sum) of [d_points; d_point_weight].

        Returns (PullbackResult, work) where work is None unless async_op."""
        n_in, P = points.shape
        packed = self.packed_buffer(n_in, P, ds_dout.dtype, ds_dout.device)
        d_points = packed[: n_in * P].view(P, n_in).t()      # (N_in, P), column-major view of the packed buffer
        d_pw = packed[n_in * P:]
        if self.pullback_fn is None:
            res = interface.raster_pullback_(ds_dout, points, rotation, translation, background, out_weight,
                                             point_weight, points_out=d_points, point_weight_out=d_pw)
        else:
            res = self.pullback_fn(ds_dout, points, rotation, translation, background, out_weight, point_weight)
            d_points.copy_(res.points)
            d_pw.copy_(res.point_weight)
            res = PullbackResult(d_points, res.rotation, res.translation, res.background, res.out_weight, d_pw)
        work = None
        if self.world_size > 1:
            if self.comm is not None:
                self.comm.all_reduce_(packed)          # enqueued on the current stream, like the kernels
            else:
                work = dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
        return res, work
