"""CPU-side model of the box-staged pullback's coverage: for config 2's synthetic cloud, sorted like
dpr_sort.cuh sorts it (counting sort by the Hilbert index of a uniform 2^bits grid over (-1.25, 1.25)^3), which fraction of
the warps (32 consecutive sorted points) has every 2 x 2 stencil inside a BOX x BOX pixel box centred on the projected
centroid of its CTA's run of points, over random poses.  A model, not a measurement: it explains the measured 56 % of the
L1 kernel's global-load sectors that pullback_box2d_kernel still issues, and sizes the alternatives for round 2.
Usage: python tools/box_coverage.py > profiles/box_coverage_r01.txt   (NumPy only, no GPU)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from tests.helpers import make_inputs


def hilbert_index(q, bits):
    """Skilling's transpose algorithm, vectorised; q: (P, N) integer cell coordinates -> (P,) index (dpr_sort.cuh::hilbert_index)."""
    X = q.astype(np.uint32).copy()
    N = X.shape[1]
    M = np.uint32(1 << (bits - 1))
    Q = M
    while Q > 1:
        Pm = np.uint32(Q - 1)
        for i in range(N):
            hit = (X[:, i] & Q) != 0
            X[hit, 0] ^= Pm
            t = (X[~hit, 0] ^ X[~hit, i]) & Pm
            X[~hit, 0] ^= t
            X[~hit, i] ^= t
        Q >>= 1
    for i in range(1, N):
        X[:, i] ^= X[:, i - 1]
    t = np.zeros(len(X), dtype=np.uint32)
    Q = M
    while Q > 1:
        hit = (X[:, N - 1] & Q) != 0
        t[hit] ^= np.uint32(Q - 1)
        Q >>= 1
    X ^= t[:, None]
    h = np.zeros(len(X), dtype=np.uint64)
    for b in range(bits - 1, -1, -1):
        for i in range(N):
            h = (h << np.uint64(1)) | ((X[:, i] >> np.uint32(b)) & np.uint32(1)).astype(np.uint64)
    return h


def sort_like_library(points, bits):
    cells = float(1 << bits)
    c = (points.T.astype(np.float32) + 1.25) * (cells / 2.5)
    q = np.clip(c, 0, cells - 1).astype(np.uint32)
    return np.argsort(hilbert_index(q, bits), kind="stable")


def coverage(points_sorted, rot, tr, grid, run, box, per_lane=False):
    """Fraction of warp slots (or lanes) on the shared-memory path for one pose."""
    g0, g1 = grid
    P = points_sorted.shape[1]
    proj = rot @ points_sorted                                    # (2, P)
    coord = (proj + 1.0 + tr[:, None]) * np.array([[g0 / 2], [g1 / 2]])
    ix = np.ceil(coord[0] - 0.5).astype(np.int64) - 1
    iy = np.ceil(coord[1] - 0.5).astype(np.int64) - 1
    n_run = (P + run - 1) // run
    pad = n_run * run - P
    pts = np.pad(points_sorted, ((0, 0), (0, pad)), mode="edge")
    ctr = pts.reshape(3, n_run, run).mean(axis=2)                # centroid of each run
    cc = (rot @ ctr + 1.0 + tr[:, None]) * np.array([[g0 / 2], [g1 / 2]])
    bx = np.clip((np.floor(cc[0]).astype(np.int64) - box // 2) & ~3, 0, g0 - box)
    by = np.clip(np.floor(cc[1]).astype(np.int64) - box // 2, 0, g1 - box)
    r = np.arange(P) // run
    inb = ((ix - bx[r]) >= 0) & ((ix - bx[r]) < box - 1) & ((iy - by[r]) >= 0) & ((iy - by[r]) < box - 1)
    if per_lane:
        return inb.mean()
    n_w = P // 32
    return inb[: n_w * 32].reshape(n_w, 32).all(axis=1).mean()


def main():
    P, B, grid = 100000, 64, (256, 256)
    d = make_inputs(1002, 3, 2, P, B, grid, np.float32, False)     # bench.py's config-2 generator and seed, 64 poses
    pts = d["points"].astype(np.float64)
    bits = 5                                                       # make_sort_plan: 16 key bits / 3 dimensions
    order = sort_like_library(d["points"], bits)
    ps = pts[:, order]
    print("# box coverage model for config 2 (100 k points 0.4*N(0,1), 256 x 256 image, 64 random poses); see tools/box_coverage.py")
    print("# run = points per CTA (equal-count runs of the Hilbert-sorted copy); warp = all 32 stencils of a warp slot in the box")
    print(f"{'run':>6} {'box':>5} {'warp-level':>11} {'lane-level':>11} {'box KB per (CTA,pose)':>22} {'L2->SM GB per launch (4096 poses)':>34}")
    for run in (512, 1024, 2048):
        for box in (48, 64, 96, 128):
            w = np.mean([coverage(ps, d["rotation"][:, :, b].astype(np.float64), d["translation"][:, b].astype(np.float64), grid, run, box) for b in range(B)])
            l = np.mean([coverage(ps, d["rotation"][:, :, b].astype(np.float64), d["translation"][:, b].astype(np.float64), grid, run, box, True) for b in range(B)])
            kb = box * box * 4 / 1024
            gb = kb * 1024 * ((P + run - 1) // run) * 4096 / 1e9
            print(f"{run:6d} {box:5d} {w:11.3f} {l:11.3f} {kb:22.1f} {gb:34.1f}")
    # how large are the runs? projected extent (max - min of the pixel coordinates) per 1024-point run, one pose
    rot, tr = d["rotation"][:, :, 0].astype(np.float64), d["translation"][:, 0].astype(np.float64)
    coord = (rot @ ps + 1.0 + tr[:, None]) * 128.0
    ext = []
    for r in range(0, P - 1023, 1024):
        c = coord[:, r:r + 1024]
        ext.append(max(c[0].max() - c[0].min(), c[1].max() - c[1].min()))
    ext = np.array(ext)
    print(f"# projected extent of the 1024-point runs (pixels, one pose): median {np.median(ext):.0f}, 25 % {np.percentile(ext, 25):.0f}, "
          f"75 % {np.percentile(ext, 75):.0f}, 90 % {np.percentile(ext, 90):.0f}, max {ext.max():.0f}")


if __name__ == "__main__":
    main()
