"""Full-size (BASELINE.json) runs on the GPU, checked through size-independent properties and pose subsets.

The oracle cannot redo 4e8..2e9 splats in a test, but poses are independent: a handful of poses of the full-size run
are compared with the oracle directly, the pose-summed gradients are checked by linearity over pose shards, and the
forward and the pullback are tied together by the adjoint identities
    <ds_dout_b, out_b - background_b> = out_weight_b * d_out_weight_b          (out - bg is linear in out_weight)
    sum_p point_weight_p * d_point_weight_p = sum_b out_weight_b * d_out_weight_b
    d_background_b = sum(ds_dout_b).
"""
import numpy as np
import pytest
import torch

import dpr_b200
from oracle import oracle
from tests.helpers import random_rotations, rel_l2

pytestmark = pytest.mark.gpu


def _full_inputs(n_in, P, B, grid, seed, weights):
    rng = np.random.Generator(np.random.PCG64(seed))
    pts = np.asfortranarray((0.4 * rng.standard_normal((n_in, P))).astype(np.float32))
    rot = random_rotations(rng, n_in, 2, B, np.float32)
    tr = np.asfortranarray((0.1 * rng.standard_normal((2, B))).astype(np.float32))
    bg = np.arange(1, B + 1, dtype=np.float32) if weights else None
    ow = (10 * rng.random(B)).astype(np.float32) if weights else None
    w = rng.random(P)
    pw = (w / w.sum() * P).astype(np.float32) if weights else None     # mean 1 keeps magnitudes comparable
    return pts, rot, tr, bg, ow, pw


def _dev(a):
    return None if a is None else dpr_b200.fortran(torch.from_numpy(np.ascontiguousarray(a)).cuda())


@pytest.mark.parametrize("name,n_in,P,B,grid,weights,do_fwd", [
    ("cfg2", 3, 100_000, 4096, (256, 256), False, True),
    ("cfg2w", 3, 100_000, 512, (256, 256), True, True),
    ("cfg5", 3, 1_000_000, 2048, (128, 128), False, False),
])
def test_full_size_properties(name, n_in, P, B, grid, weights, do_fwd):
    pts, rot, tr, bg, ow, pw = _full_inputs(n_in, P, B, grid, 2000 + len(name), weights)
    d = [_dev(a) for a in (pts, rot, tr, bg, ow, pw)]
    gen = torch.Generator(device="cuda").manual_seed(7)
    ds = dpr_b200.empty_f(grid + (B,), torch.float32, "cuda")
    ds.normal_(generator=gen)
    pb = dpr_b200.raster_pullback_(ds, *d)
    ow_t = d[4] if weights else torch.ones(B, device="cuda")
    pw_t = d[5] if weights else torch.ones(P, device="cuda")
    # d_background = sum(ds_dout) per pose
    ref_bg = ds.double().sum(dim=tuple(range(len(grid))))
    assert rel_l2(pb.background.cpu().numpy(), ref_bg.cpu().numpy()) < 1e-5
    # sum_p pw_p d_pw_p == sum_b ow_b d_ow_b
    lhs = float((pw_t.double() * pb.point_weight.double()).sum())
    rhs = float((ow_t.double() * pb.out_weight.double()).sum())
    assert abs(lhs - rhs) <= 2e-4 * max(abs(lhs), abs(rhs), float(pb.out_weight.double().abs().sum()) * 1e-2), (lhs, rhs)
    # pose subsets against the oracle (per-pose gradients) and linearity of the pose-summed gradients over shards
    sel = np.array([0, 1, B // 3, B // 2, B - 2, B - 1])
    sub = lambda a: None if a is None else np.asfortranarray(a[..., sel])
    ds_sub = np.asfortranarray(ds[..., torch.from_numpy(sel).cuda()].cpu().numpy())
    ref = oracle.raster_pullback(ds_sub, pts, sub(rot), sub(tr), sub(bg), sub(ow), pw, dtype=np.float32, n_slabs=6, f64_accumulate=True)
    for k in ("rotation", "translation", "background", "out_weight"):
        got = getattr(pb, k).cpu().numpy()[..., sel]
        assert rel_l2(got, getattr(ref, k)) <= 1e-5, (name, k)
    half = B // 2
    sh = lambda t, lo, hi: None if t is None else dpr_b200.fortran(t[..., lo:hi])
    acc_p = torch.zeros_like(pb.points, dtype=torch.float64)
    acc_w = torch.zeros_like(pb.point_weight, dtype=torch.float64)
    for lo, hi in ((0, half), (half, B)):
        part = dpr_b200.raster_pullback_(sh(ds, lo, hi), d[0], sh(d[1], lo, hi), sh(d[2], lo, hi), sh(d[3], lo, hi), sh(d[4], lo, hi), d[5])
        acc_p += part.points.double()
        acc_w += part.point_weight.double()
    assert rel_l2(pb.points.cpu().numpy(), acc_p.cpu().numpy()) <= 1e-5
    assert rel_l2(pb.point_weight.cpu().numpy(), acc_w.cpu().numpy()) <= 1e-5
    # the pose-summed d_points of the 6-pose subset alone must equal the oracle's (checks the d_points arithmetic)
    part = dpr_b200.raster_pullback_(_dev(ds_sub), d[0], _dev(sub(rot)), _dev(sub(tr)), _dev(sub(bg)), _dev(sub(ow)), d[5])
    assert rel_l2(part.points.cpu().numpy(), ref.points) <= 1e-5
    assert rel_l2(part.point_weight.cpu().numpy(), ref.point_weight) <= 1e-5
    if not do_fwd:
        return
    out = dpr_b200.raster(grid, *d)
    ref_out = oracle.raster(grid, pts, sub(rot), sub(tr), sub(bg), sub(ow), pw, dtype=np.float32, n_threads=6, f64_accumulate=True)
    assert rel_l2(out[..., torch.from_numpy(sel).cuda()].cpu().numpy(), ref_out) <= 1e-5
    # adjoint identity per pose: <ds_dout_b, out_b - bg_b> = ow_b * d_ow_b
    bg_t = d[3] if weights else torch.zeros(B, device="cuda")
    dims = tuple(range(len(grid)))
    inner = (ds.double() * (out.double() - bg_t.double().reshape((1,) * len(grid) + (B,)))).sum(dim=dims)
    want = ow_t.double() * pb.out_weight.double()
    scale = float(want.abs().mean())
    assert float((inner - want).abs().max()) <= 5e-4 * scale + 1e-6, name


def test_full_size_config3_volume():
    """BASELINE config 3 at full size - 1 M points, 16 poses, 256^3, Float32 - through the tile-binned 3-d kernels: a pose
    subset against the oracle (forward image and per-pose gradients), d_background, the pose-sum over shards, the adjoint
    identity per pose, and forward + pullback with the binning cache on (the bench's call order) against without."""
    from tests.gpu_util import forced
    n_in, P, B, grid = 3, 1_000_000, 16, (256, 256, 256)
    rng = np.random.Generator(np.random.PCG64(3003))
    pts = np.asfortranarray((0.4 * rng.standard_normal((n_in, P))).astype(np.float32))
    rot = random_rotations(rng, 3, 3, B, np.float32)
    tr = np.asfortranarray((0.1 * rng.standard_normal((3, B))).astype(np.float32))
    d = [_dev(a) for a in (pts, rot, tr, None, None, None)]
    gen = torch.Generator(device="cuda").manual_seed(11)
    ds = dpr_b200.empty_f(grid + (B,), torch.float32, "cuda")
    ds.normal_(generator=gen)
    with forced(binning_cache=0):
        out = dpr_b200.raster(grid, *d)
        assert dpr_b200.last_path(0).startswith("tile3d_binned")
        pb = dpr_b200.raster_pullback_(ds, *d)
        assert dpr_b200.last_path(1).startswith("tile3d_binned")
    sel = np.array([0, 7, 15])
    sub = lambda a: np.asfortranarray(a[..., sel])
    ref_out = oracle.raster(grid, pts, sub(rot), sub(tr), None, None, None, dtype=np.float32, n_threads=3, f64_accumulate=True)
    for i, b in enumerate(sel):
        assert rel_l2(out[..., int(b)].cpu().numpy(), ref_out[..., i]) <= 1e-5, b
    ds_sub = np.asfortranarray(ds[..., torch.from_numpy(sel).cuda()].cpu().numpy())
    ref = oracle.raster_pullback(ds_sub, pts, sub(rot), sub(tr), None, None, None, dtype=np.float32, n_slabs=3, f64_accumulate=True)
    for k in ("rotation", "translation", "background", "out_weight"):
        assert rel_l2(getattr(pb, k).cpu().numpy()[..., sel], getattr(ref, k)) <= 1e-5, k
    part = dpr_b200.raster_pullback_(_dev(ds_sub), d[0], _dev(sub(rot)), _dev(sub(tr)))
    assert rel_l2(part.points.cpu().numpy(), ref.points) <= 1e-5
    assert rel_l2(part.point_weight.cpu().numpy(), ref.point_weight) <= 1e-5
    # d_background, pose-sum linearity, adjoint identity
    ref_bg = torch.stack([ds[..., b].double().sum() for b in range(B)])
    assert rel_l2(pb.background.cpu().numpy(), ref_bg.cpu().numpy()) < 1e-5
    sh = lambda t, lo, hi: dpr_b200.fortran(t[..., lo:hi])
    acc = torch.zeros_like(pb.points, dtype=torch.float64)
    for lo, hi in ((0, 8), (8, 16)):
        acc += dpr_b200.raster_pullback_(sh(ds, lo, hi), d[0], sh(d[1], lo, hi), sh(d[2], lo, hi)).points.double()
    assert rel_l2(pb.points.cpu().numpy(), acc.cpu().numpy()) <= 1e-5
    for b in (0, 9):
        inner = float((ds[..., b].double() * out[..., b].double()).sum())
        want = float(pb.out_weight[b])
        assert abs(inner - want) <= 5e-4 * abs(want) + 1e-3, (b, inner, want)
    # the cached call order of the bench: forward, then pullback on the same inputs
    with forced(binning_cache=1):
        out_c = dpr_b200.raster(grid, *d)
        pb_c = dpr_b200.raster_pullback_(ds, *d)
    assert rel_l2(out_c.cpu().numpy(), out.cpu().numpy()) <= 1e-6
    for k in ("points", "rotation", "translation", "out_weight", "point_weight"):
        assert rel_l2(getattr(pb_c, k).cpu().numpy(), getattr(pb, k).cpu().numpy()) <= 2e-6, k


def test_full_size_config4_pose_subset():
    """BASELINE config 4 at full point count (1 M points in 2-d, 512 x 512, point weights + background): 24 of the 1024
    poses spread over the batch, forward and per-pose gradients against the oracle (the multi-slab culled forward and the
    sorted L1-gather pullback take exactly the paths of the full batch)."""
    n_in, P, B, grid = 2, 1_000_000, 24, (512, 512)
    rng = np.random.Generator(np.random.PCG64(4004))
    pts = np.asfortranarray((0.4 * rng.standard_normal((n_in, P))).astype(np.float32))
    ang = rng.uniform(0, 2 * np.pi, B)
    rot = np.empty((2, 2, B), dtype=np.float32, order="F")
    rot[0, 0], rot[0, 1], rot[1, 0], rot[1, 1] = np.cos(ang), -np.sin(ang), np.sin(ang), np.cos(ang)
    tr = np.asfortranarray((0.1 * rng.standard_normal((2, B))).astype(np.float32))
    bg = np.arange(1, B + 1, dtype=np.float32)
    ow = (10 * rng.random(B)).astype(np.float32)
    w = rng.random(P)
    pw = (w / w.sum()).astype(np.float32)
    d = [_dev(a) for a in (pts, rot, tr, bg, ow, pw)]
    gen = torch.Generator(device="cuda").manual_seed(13)
    ds = dpr_b200.empty_f(grid + (B,), torch.float32, "cuda")
    ds.normal_(generator=gen)
    out = dpr_b200.raster(grid, *d)
    assert "slabs" in dpr_b200.last_path(0)
    pb = dpr_b200.raster_pullback_(ds, *d)
    ref_out = oracle.raster(grid, pts, rot, tr, bg, ow, pw, dtype=np.float32, n_threads=8, f64_accumulate=True)
    assert rel_l2(out.cpu().numpy(), ref_out) <= 1e-5
    ref = oracle.raster_pullback(np.asfortranarray(ds.cpu().numpy()), pts, rot, tr, bg, ow, pw, dtype=np.float32, n_slabs=8, f64_accumulate=True)
    for k in ("points", "rotation", "translation", "background", "out_weight", "point_weight"):
        assert rel_l2(getattr(pb, k).cpu().numpy(), getattr(ref, k)) <= 1e-5, k


def test_tile3d_pose_groups():
    """The 3-d tile path bins at most 2^27 (point, pose) pairs per pass: 2^21 points x 70 poses run as two pose groups
    (64 + 6) that share the entry buffer - no bins can be cached, every group is binned again by the pullback.  Poses on both
    sides of the group boundary against the oracle, and the pose-summed gradients of the whole batch."""
    from tests.gpu_util import forced
    n_in, P, B, grid = 3, (1 << 21) + 5, 70, (64, 32, 48)
    rng = np.random.Generator(np.random.PCG64(5005))
    pts = np.asfortranarray((0.4 * rng.standard_normal((n_in, P))).astype(np.float32))
    rot = random_rotations(rng, 3, 3, B, np.float32)
    tr = np.asfortranarray((0.1 * rng.standard_normal((3, B))).astype(np.float32))
    ow = (0.5 + rng.random(B)).astype(np.float32)
    d = [_dev(a) for a in (pts, rot, tr, None, ow, None)]
    gen = torch.Generator(device="cuda").manual_seed(12)
    ds = dpr_b200.empty_f(grid + (B,), torch.float32, "cuda")
    ds.normal_(generator=gen)
    with forced(forward_algo=3, pullback_algo=7, binning_cache=1):
        out = dpr_b200.raster(grid, *d)
        assert dpr_b200.last_path(0).startswith("tile3d_binned")
        pb = dpr_b200.raster_pullback_(ds, *d)
        assert dpr_b200.last_path(1).startswith("tile3d_binned")
    sel = np.array([0, 63, 64, 69])
    sub = lambda a: np.asfortranarray(a[..., sel])
    ref_out = oracle.raster(grid, pts, sub(rot), sub(tr), None, ow[sel], None, dtype=np.float32, n_threads=4, f64_accumulate=True)
    assert rel_l2(out[..., torch.from_numpy(sel).cuda()].cpu().numpy(), ref_out) <= 1e-5
    ref = oracle.raster_pullback(np.asfortranarray(ds.cpu().numpy()), pts, rot, tr, None, ow, None, dtype=np.float32, n_slabs=8,
                                 f64_accumulate=True)
    for k in ("points", "rotation", "translation", "background", "out_weight", "point_weight"):
        assert rel_l2(getattr(pb, k).cpu().numpy(), getattr(ref, k)) <= 1e-5, k
