"""Sweep a libdpr tuning option on a bench config and print kernel times. Usage: python tools/sweep_options.py cfg2 pose_chunk 0 42 84 171 342"""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import dpr_b200
from dpr_b200 import _lib
import bench
cfgname, opt, values = sys.argv[1], sys.argv[2], [int(v) for v in sys.argv[3:]]
cfg = bench.CONFIGS[cfgname]
names = dict(forward_algo=0, pullback_algo=1, tile_smem_bytes=2, point_split=3, pose_chunk=4, forward_accum=5, point_sort=6)
dev = torch.device("cuda", 0)
inputs = bench.synth_inputs(cfg, 1000 + int(cfgname[-1]), 0)
td = torch.float32 if cfg["dtype"] == "f32" else torch.float64
f = lambda a: None if a is None else dpr_b200.fortran(torch.from_numpy(np.ascontiguousarray(a)).to(dev))
args = [f(inputs[k]) for k in ("points", "rotation", "translation", "background", "out_weight", "point_weight")]
grid, B = tuple(cfg["grid"]), cfg["B"]
ds = dpr_b200.empty_f(grid + (B,), td, dev); ds.normal_()
out = dpr_b200.empty_f(grid + (B,), td, dev)
for v in values:
    _lib.set_option(names[opt], v)
    for _ in range(3):
        if "fwd" in cfg["ops"]: dpr_b200.raster_(out, *args)
        dpr_b200.raster_pullback_(ds, *args)
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    n = 5
    for _ in range(n):
        if "fwd" in cfg["ops"]: dpr_b200.raster_(out, *args)
        dpr_b200.raster_pullback_(ds, *args)
    torch.cuda.synchronize()
    rec = _lib.profile_records(); _lib.profile_enable(False)
    agg = {}
    for name, ms in rec: agg[name] = agg.get(name, 0) + ms / n
    print(opt, v, {k: round(x, 3) for k, x in agg.items() if x > 0.05}, dpr_b200.last_path(0), dpr_b200.last_path(1), flush=True)
