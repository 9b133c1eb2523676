"""Host-side check of the argument behind dpr_forward_radial.cuh: for a pose with matrix rows r_0, r_1, every point with
|x| <= r_safe(pose) has all four corners of its stencil inside the on-chip rows - for ANY matrix (scaled, sheared, rank
deficient), in Float32 arithmetic with the reference's operation order (src/raster.jl:85-101).  The formulas below are the
kernel's, restated in NumPy Float32."""
import numpy as np

f32 = np.float32


def r_safe(R, t, g0, g1, ys, ye):
    """fwd_tile2d_radial_kernel: safe radius of one pose (rows [ys, ye) on chip)."""
    scale = (f32(g0) / f32(2), f32(g1) / f32(2))
    origin = (f32(-1) - f32(t[0]), f32(-1) - f32(t[1]))
    n0 = f32(np.sqrt(np.sum(R[0].astype(f32) ** 2, dtype=f32))) * scale[0]
    n1 = f32(np.sqrt(np.sum(R[1].astype(f32) ** 2, dtype=f32))) * scale[1]
    cx, cy = -origin[0] * scale[0], -origin[1] * scale[1]
    mx = min(cx - f32(0.5), f32(g0) - f32(0.5) - cx) - f32(0.05) - f32(1e-5) * f32(g0)
    my = min(cy - (f32(ys) + f32(0.5)), (f32(ye) - f32(0.5)) - cy) - f32(0.05) - f32(1e-5) * f32(g1)
    if not (mx > 0 and my > 0 and n0 < 1e30 and n1 < 1e30):
        return f32(-1)
    rs = min(mx / max(n0, f32(1e-30)), my / max(n1, f32(1e-30))) * f32(0.9999)
    return f32(min(rs, f32(1e30)))


def stencil_cell(R, t, x, g0, g1):
    """1-based lower-corner cell of src/raster.jl:88-94, Float32, left-to-right, unfused."""
    out = []
    for k, g in ((0, g0), (1, g1)):
        proj = f32(R[k, 0]) * f32(x[0])
        for j in range(1, len(x)):
            proj = f32(proj + f32(f32(R[k, j]) * f32(x[j])))
        origin = f32(-1) - f32(t[k])
        coord = f32(f32(proj - origin) * (f32(g) / f32(2)))
        out.append(int(np.ceil(f32(coord - f32(0.5)))))
    return out


def test_points_inside_the_safe_radius_are_interior():
    rng = np.random.default_rng(2026)
    checked = 0
    for trial in range(400):
        n_in = 2 + trial % 2
        g0, g1 = int(rng.integers(8, 600)), int(rng.integers(8, 600))
        rows = int(rng.integers(max(4, g1 // 2), g1 + 1))
        ys = int(rng.integers(0, g1 - rows + 1))
        ye = ys + rows
        R = rng.standard_normal((2, n_in)).astype(f32) * f32(rng.uniform(0.05, 3.0))
        if trial % 7 == 0:
            R[1] = 0                                        # rank deficient
        if trial % 11 == 0:
            q, _ = np.linalg.qr(rng.standard_normal((n_in, n_in)))
            R = q[:2].astype(f32)                           # a proper projection of a rotation
        t = (rng.standard_normal(2) * 0.4).astype(f32)
        rs = r_safe(R, t, g0, g1, ys, ye)
        if rs <= 0:
            continue
        for _ in range(40):
            d = rng.standard_normal(n_in)
            # points ON the safe sphere and inside it, including the directions that maximise |r_k . x|
            if _ % 4 == 0:
                d = R[_ % 8 // 4].astype(np.float64) + 1e-12
            x = (d / np.linalg.norm(d) * float(rs) * rng.choice([1.0, -1.0, 0.999, 0.5])).astype(f32)
            if f32(np.sqrt(np.sum(x.astype(f32) ** 2, dtype=f32))) > rs:
                x = (x * f32(0.99999)).astype(f32)          # the kernel compares the Float32 radius of the stored point
                if f32(np.sqrt(np.sum(x.astype(f32) ** 2, dtype=f32))) > rs:
                    continue
            rx, ry = stencil_cell(R, t, x, g0, g1)
            # all four corners on chip  <=>  1 <= rx <= g0 - 1  and  ys + 1 <= ry <= ye - 1   (1-based cells)
            assert 1 <= rx <= g0 - 1 and ys + 1 <= ry <= ye - 1, (trial, R, t, x, rs, rx, ry, g0, g1, ys, ye)
            checked += 1
    assert checked > 5000
