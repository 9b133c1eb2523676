"""Relative L2 distance of every output of the CUDA library to the CPU oracle, per BASELINE config (scaled to what
the oracle finishes in seconds), plus the faithful-Float32 oracle's own distance to the f64-accumulating oracle
(SURVEY.md 7 H5).  Usage (GPU box): python tools/accuracy_report.py > profiles/accuracy_r01.txt"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import dpr_b200
from oracle import oracle
from tests.helpers import make_inputs, rel_l2
from tests.gpu_util import dev_args, to_dev, to_np, forced

CASES = [("cfg1 (full size)", 3, 2, 10000, 64, (128, 128), np.float64, False),
         ("cfg2 (100k points, 16 of 4096 poses)", 3, 2, 100000, 16, (256, 256), np.float32, False),
         ("cfg3 (1M points, 2 of 16 poses, 128^3)", 3, 3, 1000000, 2, (128, 128, 128), np.float32, False),
         ("cfg4 (1M points, 4 of 1024 poses)", 2, 2, 1000000, 4, (512, 512), np.float32, True),
         ("cfg5 (1M points, 32 of 16384 poses)", 3, 2, 1000000, 32, (128, 128), np.float32, False),
         # dimension pairs beyond the BASELINE configs (every pair up to 4 x 4 is instantiated; generic kernels)
         ("4d->4d, 200k points, 4 poses, 24^4", 4, 4, 200000, 4, (24, 24, 24, 24), np.float32, True),
         ("4d->2d, 200k points, 8 poses, 128^2", 4, 2, 200000, 8, (128, 128), np.float32, False),
         ("2d->3d (embedding), 200k points, 4 poses, 64^3", 2, 3, 200000, 4, (64, 64, 64), np.float64, True)]
F = ("points", "rotation", "translation", "background", "out_weight", "point_weight")
for name, n_in, n_out, P, B, grid, dtype, weights in CASES:
    d = make_inputs(1000 + int(name[3]) if name.startswith("cfg") else 2000 + 10 * n_in + n_out, n_in, n_out, P, B, grid, dtype, weights)
    a = tuple(d[k] for k in F)
    acc = dtype == np.float32
    ref_out = oracle.raster(grid, *a, dtype=dtype, n_threads=8, f64_accumulate=acc)
    ref_pb = oracle.raster_pullback(d["ds_dout"], *a, dtype=dtype, n_slabs=8, f64_accumulate=acc)
    td = torch.float32 if acc else torch.float64
    args = dev_args(d, dtype)
    print(f"== {name}, {'Float32 vs f64-accumulate oracle (gate 1e-5)' if acc else 'Float64 vs faithful oracle (gate 1e-10)'}")
    if n_out == 2 and n_in in (2, 3) and acc:     # the tile kernels exist for 2-d and 3-d points
        for label, opts in (("fixed-point tile", dict(forward_accum=0)), ("float CAS tile", dict(forward_accum=1)), ("global REDG", dict(forward_algo=1))):
            with forced(**opts):
                out = dpr_b200.raster(grid, *args)
                print(f"   forward [{dpr_b200.last_path(0):28s}] {label:18s} rel L2 {rel_l2(to_np(out), ref_out):.2e}")
    else:
        out = dpr_b200.raster(grid, *args)
        print(f"   forward [{dpr_b200.last_path(0):28s}] rel L2 {rel_l2(to_np(out), ref_out):.2e}")
    pb = dpr_b200.raster_pullback_(to_dev(d["ds_dout"], td), *args)
    print(f"   pullback [{dpr_b200.last_path(1)}] " + "  ".join(f"{k} {rel_l2(to_np(getattr(pb, k)), getattr(ref_pb, k)):.1e}" for k in F))
    if acc:
        fo = oracle.raster(grid, *a, dtype=dtype, n_threads=8)
        fp = oracle.raster_pullback(d["ds_dout"], *a, dtype=dtype, n_slabs=8)
        print(f"   faithful Float32 oracle's own distance: forward {rel_l2(fo, ref_out):.1e}  " + "  ".join(f"{k} {rel_l2(getattr(fp, k), getattr(ref_pb, k)):.1e}" for k in F))
