"""TEST INFRASTRUCTURE - second, independent restatement of the reference math in NumPy (small cases only).

Vectorised over points, one pose at a time, always Float64 accumulation via np.add.at.  Used by the CPU test
suite to cross-check the C oracle (two restatements written separately must agree) and for finite differences.
Follows the spec of /root/reference src/raster.jl:36-108 and src/raster_pullback.jl:39-78, 150-160.
"""
from __future__ import annotations

import itertools

import numpy as np


def _stencil(points, R, t, grid, dtype):
    """points (n_in,P), R (n_out,n_in), t (n_out,) -> ref (n_out,P) int64 1-based, dl (n_out,P)."""
    n_out, n_in = R.shape
    g = np.asarray(grid, dtype=dtype)
    scale = g / dtype(2)                                       # src/raster.jl:25
    proj = R[:, 0:1] * points[0:1, :]
    for j in range(1, n_in):                                   # left-to-right, unfused (src/raster.jl:88)
        proj = proj + R[:, j:j + 1] * points[j:j + 1, :]
    origin = (dtype(-1) - t)[:, None]                          # src/raster.jl:53
    coord = (proj - origin) * scale[:, None]                   # src/raster.jl:92
    r = np.ceil(coord - dtype(0.5))                            # src/raster.jl:94
    r = np.clip(r, -4e18, 4e18)
    dl = coord - (r - dtype(0.5))                              # src/raster.jl:97
    return r.astype(np.int64), dl.astype(dtype), scale


def raster(grid_size, points, rotation, translation, background=None, out_weight=None, point_weight=None,
           dtype=np.float64):
    dtype = np.dtype(dtype).type
    points = np.asarray(points, dtype=dtype)
    rotation = np.asarray(rotation, dtype=dtype)
    translation = np.asarray(translation, dtype=dtype)
    n_out, n_in, B = rotation.shape
    P = points.shape[1]
    bg = np.zeros(B, dtype) if background is None else np.asarray(background, dtype)
    ow = np.ones(B, dtype) if out_weight is None else np.asarray(out_weight, dtype)
    pw = np.ones(P, dtype) if point_weight is None else np.asarray(point_weight, dtype)
    out = np.empty(tuple(grid_size) + (B,), dtype=np.float64, order="F")
    for b in range(B):
        img = np.full(tuple(grid_size), float(bg[b]), dtype=np.float64, order="F")   # src/raster.jl:27
        ref, dl, _ = _stencil(points, rotation[:, :, b], translation[:, b], grid_size, dtype)
        du = dtype(1) - dl
        weight = ow[b] * pw                                                          # src/raster.jl:51
        for shift in itertools.product((0, 1), repeat=n_out):                        # any corner order: f64 sums
            idx = ref + np.asarray(shift)[:, None]
            inb = np.all((idx >= 1) & (idx <= np.asarray(grid_size)[:, None]), axis=0)  # src/raster.jl:62
            w = np.ones(P, dtype=dtype)
            for k in range(n_out):
                w = w * (dl[k] if shift[k] == 1 else du[k])                          # src/raster.jl:104-106
            val = (w * weight).astype(np.float64)
            np.add.at(img, tuple(idx[k, inb] - 1 for k in range(n_out)), val[inb])
        out[..., b] = img
    return out.astype(dtype)


def raster_pullback(ds_dout, points, rotation, translation, background=None, out_weight=None, point_weight=None,
                    dtype=np.float64):
    dtype = np.dtype(dtype).type
    ds_dout = np.asarray(ds_dout, dtype=dtype)
    points = np.asarray(points, dtype=dtype)
    rotation = np.asarray(rotation, dtype=dtype)
    translation = np.asarray(translation, dtype=dtype)
    n_out, n_in, B = rotation.shape
    P = points.shape[1]
    grid_size = ds_dout.shape[:-1]
    ow = np.ones(B, dtype) if out_weight is None else np.asarray(out_weight, dtype)
    pw = np.ones(P, dtype) if point_weight is None else np.asarray(point_weight, dtype)
    d_points = np.zeros((n_in, P))
    d_pw = np.zeros(P)
    d_rot = np.zeros((n_out, n_in, B))
    d_tr = np.zeros((n_out, B))
    d_bg = np.zeros(B)
    d_ow = np.zeros(B)
    x64 = points.astype(np.float64)
    for b in range(B):
        R = rotation[:, :, b]
        ref, dl, scale = _stencil(points, R, translation[:, b], grid_size, dtype)
        dl = dl.astype(np.float64)
        du = 1.0 - dl
        G_img = ds_dout[..., b].astype(np.float64)
        d_coord = np.zeros((n_out, P))
        for shift in itertools.product((0, 1), repeat=n_out):
            idx = ref + np.asarray(shift)[:, None]
            inb = np.all((idx >= 1) & (idx <= np.asarray(grid_size)[:, None]), axis=0)   # src/raster_pullback.jl:51
            G = np.zeros(P)
            G[inb] = G_img[tuple(idx[k, inb] - 1 for k in range(n_out))]
            w = np.ones(P)
            for k in range(n_out):
                w = w * (dl[k] if shift[k] == 1 else du[k])
            d_ow[b] += np.sum(w * G * pw)                                                # :57
            d_pw += w * G * float(ow[b])                                                 # :58
            factor = G * float(ow[b]) * pw                                               # :60
            for n in range(n_out):                                                       # :150-160
                iw = np.full(P, 1.0 if shift[n] == 1 else -1.0)
                for m in range(n_out):
                    if m != n:
                        iw = iw * (dl[m] if shift[m] == 1 else du[m])
                d_coord[n] += factor * iw
        scaled = d_coord * scale.astype(np.float64)[:, None]                             # :67
        d_tr[:, b] = scaled.sum(axis=1)                                                  # :68
        d_rot[:, :, b] = scaled @ x64.T                                                  # :69
        d_points += R.astype(np.float64).T @ scaled                                      # :70-71
        d_bg[b] = G_img.sum()                                                            # :78
    return dict(points=d_points.astype(dtype), rotation=d_rot.astype(dtype), translation=d_tr.astype(dtype),
                background=d_bg.astype(dtype), out_weight=d_ow.astype(dtype), point_weight=d_pw.astype(dtype))
