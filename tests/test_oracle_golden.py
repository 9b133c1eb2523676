"""Pins the CPU oracle (oracle/) against the reference's own known-answer tests - CPU only.

Golden vectors: tests/golden/reference_known_answers.json, transcribed from /root/reference
src/raster.jl:143-309, README.md:41-68, README.md:99-183 and src/util.jl:29-46 by tests/golden/make_golden.py.
"""
import numpy as np
import pytest

from oracle import oracle, oracle_np
from tests.helpers import golden_forward_args, make_inputs, rel_l2


def _cases(golden):
    return golden["forward"]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("impl", ["c", "c_f64acc", "numpy"])
def test_forward_known_answers(golden, dtype, impl):
    assert len(golden["forward"]) == 13
    for case in _cases(golden):
        grid, pts, rot, tr, bg, ow, pw = golden_forward_args(case, dtype)
        if impl == "numpy":
            out = oracle_np.raster(grid, pts, rot, tr, bg, ow, pw, dtype=dtype)
        else:
            out = oracle.raster(grid, pts, rot, tr, bg, ow, pw, dtype=dtype, f64_accumulate=(impl == "c_f64acc"))
        expected = np.asarray(case["expected"], dtype=np.float64)
        # the reference asserts `out ≈ expected` (isapprox, rtol sqrt(eps)); 0.4*2.5 etc. are not exact in binary
        tol = 1e-12 if dtype == np.float64 else 2e-6
        assert out.shape == (5, 5, 1)
        np.testing.assert_allclose(out[:, :, 0], expected, rtol=0, atol=tol * 4, err_msg=case["name"] + " " + case["cite"])


def test_voxel_shifts_known_answers(golden):
    for n, expected in golden["voxel_shifts"].items():
        np.testing.assert_array_equal(oracle.voxel_shifts(int(n)), np.asarray(expected))
    # digitstuple(5, Val(3)) == (1,0,1); digitstuple(2, Val(4)) == (0,1,0,0)   (src/util.jl:10-14)
    np.testing.assert_array_equal(oracle.voxel_shifts(3)[5], [1, 0, 1])
    np.testing.assert_array_equal(oracle.voxel_shifts(4)[2], [0, 1, 0, 0])


@pytest.mark.parametrize("impl", ["c", "c_f64acc", "numpy"])
def test_pullback_readme_known_answer(golden, impl):
    g = golden["pullback"]
    pts = np.asarray(g["points"]).T.copy(order="F")
    rot = np.asarray(g["rotation"])[:, :, None].copy(order="F")
    tr = np.asarray(g["translation"])[:, None].copy(order="F")
    ds_dout = np.asarray(g["ds_dout"])[:, :, None].copy(order="F")
    if impl == "numpy":
        r = oracle_np.raster_pullback(ds_dout, pts, rot, tr)
        d_points, d_rot, d_tr = r["points"], r["rotation"], r["translation"]
    else:
        r = oracle.raster_pullback(ds_dout, pts, rot, tr, f64_accumulate=(impl == "c_f64acc"))
        d_points, d_rot, d_tr = r.points, r.rotation, r.translation
    # ds_dout is printed with 6 significant digits (README.md:151-157): compare at that resolution
    np.testing.assert_allclose(d_points, np.asarray(g["d_points"]), rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(d_rot[:, :, 0], np.asarray(g["d_rotation"]), rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(d_tr[:, 0], np.asarray(g["d_translation"]), rtol=2e-5, atol=2e-5)
    # full-digit Zygote gradient (README.md:120-137) is minus the pullback
    np.testing.assert_allclose(d_points.T, -np.asarray(g["zygote_d_points"]), rtol=2e-5, atol=2e-5)


def test_readme_ds_dout_is_consistent_with_forward(golden):
    """ds_dout = 2 (target - raster(points)) (README.md:151): ties the forward and the pullback vectors together."""
    g = golden["pullback"]
    pts = np.asarray(g["points"]).T.copy(order="F")
    rot = np.asarray(g["rotation"])[:, :, None].copy(order="F")
    tr = np.asarray(g["translation"])[:, None].copy(order="F")
    out = oracle.raster((5, 5), pts, rot, tr)[:, :, 0]
    np.testing.assert_allclose(2 * (np.asarray(g["target_image"]) - out), np.asarray(g["ds_dout"]), atol=2e-5)


@pytest.mark.parametrize("n_in,n_out,grid", [(3, 2, (8, 8)), (3, 3, (8, 8, 8)), (2, 2, (16, 12)), (3, 1, (9,)),
                                             (4, 4, (5, 4, 6, 3)), (4, 2, (8, 8)), (2, 3, (6, 7, 5)), (1, 4, (3, 4, 5, 6))])
@pytest.mark.parametrize("weights", [True, False])
def test_c_oracle_matches_numpy_restatement(n_in, n_out, grid, weights):
    d = make_inputs(7, n_in, n_out, 300, 5, grid, np.float64, weights)
    args = (d["points"], d["rotation"], d["translation"], d["background"], d["out_weight"], d["point_weight"])
    out_c = oracle.raster(grid, *args)
    out_np = oracle_np.raster(grid, *args)
    assert rel_l2(out_c, out_np) < 1e-13
    pb_c = oracle.raster_pullback(d["ds_dout"], *args, n_slabs=3)
    pb_np = oracle_np.raster_pullback(d["ds_dout"], *args)
    for k in pb_c._fields:
        assert rel_l2(getattr(pb_c, k), pb_np[k]) < 1e-12, k


@pytest.mark.parametrize("n_in,n_out,grid", [(3, 2, (8, 8)), (3, 3, (8, 8, 8)), (2, 2, (8, 8))])
def test_batched_equals_singles_and_thread_invariance(n_in, n_out, grid):
    """batched == per-pose singles (src/raster.jl:383-431, src/raster_pullback.jl:271-345)."""
    d = make_inputs(11, n_in, n_out, 2000, 7, grid, np.float64)
    args = (d["points"], d["rotation"], d["translation"], d["background"], d["out_weight"], d["point_weight"])
    out = oracle.raster(grid, *args, n_threads=4)
    pb = oracle.raster_pullback(d["ds_dout"], *args, n_slabs=3)
    d_points = np.zeros_like(pb.points)
    d_pw = np.zeros_like(pb.point_weight)
    for b in range(7):
        sl = lambda a: None if a is None else a[..., b:b + 1]
        one = (d["points"], sl(d["rotation"]), sl(d["translation"]), sl(d["background"]), sl(d["out_weight"]), d["point_weight"])
        np.testing.assert_array_equal(oracle.raster(grid, *one)[..., 0], out[..., b])
        pb1 = oracle.raster_pullback(sl(d["ds_dout"]), *one)
        np.testing.assert_array_equal(pb1.rotation[..., 0], pb.rotation[..., b])
        np.testing.assert_array_equal(pb1.translation[..., 0], pb.translation[..., b])
        assert pb1.background[0] == pb.background[b] and pb1.out_weight[0] == pb.out_weight[b]
        d_points += pb1.points
        d_pw += pb1.point_weight
    assert rel_l2(pb.points, d_points) < 1e-13 and rel_l2(pb.point_weight, d_pw) < 1e-13


@pytest.mark.parametrize("n_in,n_out,grid", [(3, 2, (8, 8)), (3, 3, (8, 8, 8)), (4, 4, (5, 4, 6, 3)), (2, 3, (6, 7, 5))])
def test_pullback_is_gradient_of_forward_finite_differences(n_in, n_out, grid):
    """<ds_dout, raster(args)> differentiated numerically (test/chainrules.jl:6-89 does this with test_rrule)."""
    d = make_inputs(3, n_in, n_out, 12, 3, grid, np.float64)
    names = ["points", "rotation", "translation", "background", "out_weight", "point_weight"]
    base = [np.array(d[k], dtype=np.float64, order="F") for k in names]
    pb = oracle.raster_pullback(d["ds_dout"], *base)

    def loss(args):
        return float(np.sum(oracle.raster(grid, *args) * d["ds_dout"]))

    rng = np.random.default_rng(0)
    eps = 1e-7
    for i, k in enumerate(names):
        direction = rng.standard_normal(base[i].shape)
        plus = [a.copy() for a in base]
        minus = [a.copy() for a in base]
        plus[i] = plus[i] + eps * direction
        minus[i] = minus[i] - eps * direction
        fd = (loss(plus) - loss(minus)) / (2 * eps)
        an = float(np.sum(getattr(pb, k) * direction))
        assert abs(fd - an) <= 1e-5 * max(1.0, abs(an)), (k, fd, an)


def test_f32_faithful_vs_f64_accumulate_distance():
    """the faithful Float32 oracle carries its own accumulation noise; it must still be close."""
    d = make_inputs(5, 3, 2, 20000, 4, (32, 32), np.float32)
    args = (d["points"], d["rotation"], d["translation"], d["background"], d["out_weight"], d["point_weight"])
    a = oracle.raster((32, 32), *args, dtype=np.float32)
    b = oracle.raster((32, 32), *args, dtype=np.float32, f64_accumulate=True)
    assert rel_l2(a, b) < 1e-6
    pa = oracle.raster_pullback(d["ds_dout"], *args, dtype=np.float32, n_slabs=2)
    pb = oracle.raster_pullback(d["ds_dout"], *args, dtype=np.float32, n_slabs=2, f64_accumulate=True)
    for k in pa._fields:
        assert rel_l2(getattr(pa, k), getattr(pb, k)) < 2e-5, k


def test_far_away_and_boundary_points_are_clipped_per_corner():
    """Points outside the cube contribute only through in-bounds corners (src/raster.jl:62); huge coords must
    saturate, not wrap."""
    pts = np.asfortranarray(np.array([[1e30, -1e30, 0.999, -1.0, 1.0], [0.0, 0.0, 0.999, -1.0, 1.0]]))
    rot = np.eye(2)[:, :, None]
    tr = np.zeros((2, 1))
    out = oracle.raster((4, 4), pts, rot, tr)[:, :, 0]
    assert np.isfinite(out).all()
    # (0.999,0.999) -> coord 3.998: ref 4, dl 0.498 -> only corner (4,4) is in bounds, weight (1-dl)^2
    # (-1,-1) -> coord 0: ref 0, dl 0.5 -> only corner (1,1), weight 0.25 ; (1,1) -> coord 4: ref 4, dl .5 -> (4,4) .25
    assert abs(out[0, 0] - 0.25) < 1e-12
    assert abs(out[3, 3] - (0.25 + (1 - 0.498) ** 2)) < 1e-9
    assert abs(out.sum() - (0.25 + 0.25 + (1 - 0.498) ** 2)) < 1e-9
