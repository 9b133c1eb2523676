"""Shared helpers for the test-suite: golden-case unpacking and the synthetic input distributions of the
reference's test fixtures (/root/reference test/data.jl:22-84), seeded."""
import numpy as np


def golden_forward_args(case, dtype=np.float64):
    """Turn one transcribed single-image known-answer case into batched (B=1) nd-array arguments."""
    pts = np.asarray(case["points"], dtype=dtype).T.copy(order="F")                 # (N_in, P)
    rot = np.asarray(case["rotation"], dtype=dtype)[:, :, None].copy(order="F")     # (N_out, N_in, 1)
    tr = np.asarray(case["translation"], dtype=dtype)[:, None].copy(order="F")      # (N_out, 1)
    bg = None if case["background"] is None else np.asarray([case["background"]], dtype=dtype)
    ow = None if case["out_weight"] is None else np.asarray([case["out_weight"]], dtype=dtype)
    pw = None if case["point_weight"] is None else np.asarray(case["point_weight"], dtype=dtype)
    return tuple(case["grid_size"]), pts, rot, tr, bg, ow, pw


def random_rotations(rng, n_in, n_out, B, dtype):
    """Haar-uniform rotations (test/data.jl:29-31); projections keep the first n_out rows (P*R, :13-16,:42-44)."""
    out = np.empty((n_out, n_in, B), dtype=dtype, order="F")
    n = max(n_in, n_out)                     # n_out > n_in (an embedding): the first n_in columns of a rotation
    for b in range(B):
        a = rng.standard_normal((n, n))
        q, r = np.linalg.qr(a)
        q = q * np.sign(np.diag(r))
        if n > 1 and np.linalg.det(q) < 0:
            q[:, 0] = -q[:, 0]
        out[:, :, b] = q[:n_out, :n_in]
    return out


def make_inputs(seed, n_in, n_out, P, B, grid, dtype, weights=True):
    """Synthetic inputs with the distributions of the reference fixtures (test/data.jl):
    points 0.4*randn (:22,:27), translations 0.1*randn (:54-64), backgrounds 1..B (:72),
    out weights 10*rand (:75), point weights rand normalised to sum 1 (:77-84), ds_dout randn (test/cuda.jl:46)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    points = np.asfortranarray((0.4 * rng.standard_normal((n_in, P))).astype(dtype))
    rotation = random_rotations(rng, n_in, n_out, B, dtype)
    translation = np.asfortranarray((0.1 * rng.standard_normal((n_out, B))).astype(dtype))
    ds_dout = np.asfortranarray(rng.standard_normal(tuple(grid) + (B,)).astype(dtype))
    if weights:
        background = np.arange(1, B + 1, dtype=dtype)
        out_weight = (10 * rng.random(B)).astype(dtype)
        w = rng.random(P)
        point_weight = (w / w.sum()).astype(dtype)
    else:
        background = out_weight = point_weight = None
    return dict(points=points, rotation=rotation, translation=translation, background=background,
                out_weight=out_weight, point_weight=point_weight, ds_dout=ds_dout, grid=tuple(grid))


def rel_l2(x, ref):
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    denom = np.linalg.norm(ref.ravel())
    return np.linalg.norm((x - ref).ravel()) / (denom if denom > 0 else 1.0)
