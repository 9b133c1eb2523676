"""Fixed vs per-entry cost of the 3-d tile kernels: config 3's grid and poses with 0 .. 2 M points.  Usage (GPU box): python tools/exp_tile3d_scaling.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dpr_b200
from dpr_b200 import _lib
from tests.helpers import random_rotations
grid, B = (256, 256, 256), 16
rng = np.random.Generator(np.random.PCG64(1))
rot = dpr_b200.fortran(torch.from_numpy(random_rotations(rng, 3, 3, B, np.float32)).cuda())
tr = dpr_b200.fortran(torch.from_numpy(np.asfortranarray((0.1 * rng.standard_normal((3, B))).astype(np.float32))).cuda())
ds = dpr_b200.empty_f(grid + (B,), torch.float32, "cuda"); ds.normal_()
_lib.set_option(_lib.OPT_FORWARD_ALGO, 3); _lib.set_option(_lib.OPT_PULLBACK_ALGO, 7)
for P in (1000, 125_000, 250_000, 500_000, 1_000_000, 2_000_000):
    pts = dpr_b200.fortran(torch.from_numpy(np.asfortranarray((0.4 * rng.standard_normal((3, P))).astype(np.float32))).cuda())
    for _ in range(2):
        out = dpr_b200.raster(grid, pts, rot, tr); pb = dpr_b200.raster_pullback_(ds, pts, rot, tr)
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    for _ in range(5):
        out = dpr_b200.raster(grid, pts, rot, tr); pb = dpr_b200.raster_pullback_(ds, pts, rot, tr)
    torch.cuda.synchronize()
    rec = {}
    for name, ms in _lib.profile_records():
        rec.setdefault(name, []).append(ms)
    _lib.profile_enable(False)
    print(P, {k: round(1e3 * sum(v) / 5, 1) for k, v in rec.items() if k in ("fwd_tile3d", "pullback_tile3d", "tile3_bin_count", "tile3_bin_scatter", "tile3_sort_scatter", "tile3_sort_count")}, flush=True)
