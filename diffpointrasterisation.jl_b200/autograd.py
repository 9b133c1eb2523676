"""Reverse-mode rule for `raster`, mirroring the reference's ChainRules `rrule`
(/root/reference ext/DiffPointRasterisationChainRulesCoreExt.jl:6-27 single image, :48-74 batch) on top of the same
two library calls: the forward is `raster`, the pullback closure calls `raster_pullback!` with the incoming cotangent
and returns tangents for points, rotation, translation and for exactly the optional arguments that were passed
(`values(out_pb)[4:3+n_optional]`, ext :68-70).  `grid_size` gets no tangent (NoTangent).

PyTorch's autograd is only the host-side tape here (like Zygote on the Julia side); both passes run in libdpr.so.
"""
from __future__ import annotations

import torch

from . import interface


class _Raster(torch.autograd.Function):
    @staticmethod
    def forward(ctx, grid_size, points, rotation, translation, background, out_weight, point_weight):
        ctx.save_for_backward(*[t for t in (points, rotation, translation, background, out_weight, point_weight) if isinstance(t, torch.Tensor)])
        ctx.layout = [isinstance(t, torch.Tensor) for t in (points, rotation, translation, background, out_weight, point_weight)]
        ctx.scalars = [None if isinstance(t, torch.Tensor) else t for t in (points, rotation, translation, background, out_weight, point_weight)]
        with torch.no_grad():
            return interface.raster(grid_size, points, rotation, translation, background, out_weight, point_weight)

    @staticmethod
    def backward(ctx, ds_dout):
        saved = list(ctx.saved_tensors)
        args = [saved.pop(0) if is_t else s for is_t, s in zip(ctx.layout, ctx.scalars)]
        with torch.no_grad():
            pb = interface.raster_pullback_(interface.fortran(ds_dout), *args)
        grads = [pb.points, pb.rotation, pb.translation, pb.background, pb.out_weight, pb.point_weight]
        out = []
        for g, a in zip(grads, args):
            if isinstance(a, torch.Tensor) and a.requires_grad:
                out.append(g.reshape(a.shape).to(a.dtype))
            else:
                out.append(None)      # not passed / not differentiated: no tangent, like the rrule
        return (None, *out)


def raster(grid_size, points, rotation, translation, background=None, out_weight=None, point_weight=None):
    """Differentiable `raster` (batched or single image): gradients flow to every tensor argument that requires grad."""
    return _Raster.apply(tuple(int(g) for g in grid_size), points, rotation, translation, background, out_weight, point_weight)
