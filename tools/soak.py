"""Randomised soak of the auto-selected kernel paths against the oracle (GPU box).  Usage: python tools/soak.py [seconds] [seed]
Shapes are drawn so that every dispatch branch is reachable: 2-d tile / radial / slabs / global, TMA-staged and L1 pullbacks,
the 3-d tile path with and without TMA rows, generic dimension pairs; both element types; default and explicit weights;
sliced (element-aligned) buffers; repeated calls on one workspace (binning cache on, as the Python mirror sets it)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dpr_b200
from oracle import oracle
from tests.helpers import make_inputs, rel_l2
from tests.gpu_util import forced
FIELDS = ("points", "rotation", "translation", "background", "out_weight", "point_weight")
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2024)
t0, n, paths, worst = time.time(), 0, {}, 0.0
while time.time() - t0 < budget:
    dtype = np.float32 if rng.random() < 0.6 else np.float64
    td = torch.float32 if dtype == np.float32 else torch.float64
    kind = rng.integers(0, 5)
    if kind == 0:      # 3-d -> 3-d volumes around the tile size, dense enough for the tile path
        n_in, n_out = 3, 3
        grid = tuple(int(v) for v in rng.integers(8, 72, 3))
        if rng.random() < 0.5: grid = (grid[0] // 4 * 4 + 4,) + grid[1:]
        P, B = int(rng.integers(2000, 60000)), int(rng.integers(1, 7))
    elif kind == 1:    # 3-d -> 2-d, images that fit shared memory (TMA-staged pullback, radial / split forward)
        n_in, n_out = 3, 2
        grid = (int(rng.integers(2, 33)) * 4, int(rng.integers(8, 130)))
        P, B = int(rng.integers(500, 80000)), int(rng.integers(1, 40))
    elif kind == 2:    # 2-d -> 2-d, larger images (slabs, L1 gathers)
        n_in, n_out = 2, 2
        grid = (int(rng.integers(60, 400)), int(rng.integers(60, 400)))
        P, B = int(rng.integers(1000, 120000)), int(rng.integers(1, 12))
    elif kind == 3:    # generic dimension pairs
        n_in, n_out = int(rng.integers(1, 5)), int(rng.integers(1, 5))
        grid = tuple(int(v) for v in rng.integers(3, 14 if n_out > 2 else 60, n_out))
        P, B = int(rng.integers(10, 5000)), int(rng.integers(1, 6))
    else:              # 3-d -> 2-d, odd shapes
        n_in, n_out = 3, 2
        grid = (int(rng.integers(5, 300)), int(rng.integers(5, 300)))
        P, B = int(rng.integers(1, 30000)), int(rng.integers(1, 70))
    weights = bool(rng.random() < 0.5)
    seed = int(rng.integers(1 << 30))
    d = make_inputs(seed, n_in, n_out, P, B, grid, dtype, weights)
    far = bool(rng.random() < 0.3)
    if far: d["translation"] *= 4.0          # clouds partly outside
    acc = dtype == np.float32
    ref_out = oracle.raster(grid, *(d[k] for k in FIELDS), dtype=dtype, f64_accumulate=acc, n_threads=8)
    ref_pb = oracle.raster_pullback(d["ds_dout"], *(d[k] for k in FIELDS), dtype=dtype, f64_accumulate=acc, n_slabs=min(B, 8))
    sliced = []
    def dev(a):
        if a is None: return None
        t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
        sliced.append(bool(rng.random() < 0.3))
        if sliced[-1]:      # element-aligned slice of a larger buffer
            big = torch.zeros(t.shape[:-1] + (t.shape[-1] + 3,), dtype=t.dtype, device="cuda")
            big = dpr_b200.fortran(big); v = big[..., 1:1 + t.shape[-1]]; v.copy_(t); return v
        return dpr_b200.fortran(t)
    args = [dev(d[k]) for k in FIELDS]
    ds = dev(d["ds_dout"])
    tol = 1e-5 if acc else 1e-10
    force = dict(forward_algo=3, pullback_algo=7) if (kind == 0 and rng.random() < 0.6) else {}      # small volumes: force the tile path
    for rep in range(2):             # the second round runs on the workspace (and, for 3-d, the bins) the first left behind
        with forced(**force):
            out = dpr_b200.raster(grid, *args)
            pf = dpr_b200.last_path(0)
            pb = dpr_b200.raster_pullback_(ds, *args)
            pp = dpr_b200.last_path(1)
        torch.cuda.synchronize()
        errs = {"out": rel_l2(out.cpu().numpy(), ref_out)}
        for k in FIELDS: errs[k] = rel_l2(getattr(pb, k).cpu().numpy(), getattr(ref_pb, k))
        bad = {k: v for k, v in errs.items() if not (v <= tol)}
        if bad and acc:
            # Float32: the reference's own sequential Float32 sums are this far from the exact ones too - the tests' triangle
            # bound dist(GPU, f64acc) <= dist(faithful f32, f64acc) + tol
            f_out = oracle.raster(grid, *(d[k] for k in FIELDS), dtype=dtype, n_threads=1)
            f_pb = oracle.raster_pullback(d["ds_dout"], *(d[k] for k in FIELDS), dtype=dtype, n_slabs=1)
            own = {"out": rel_l2(f_out, ref_out)}
            for k in FIELDS: own[k] = rel_l2(getattr(f_pb, k), getattr(ref_pb, k))
            bad = {k: v for k, v in bad.items() if not (v <= own[k] + tol)}
            # a single number that is a sum of thousands of signed terms (one pose, one output dimension) carries the
            # cancellation of that sum: an order of magnitude more room for outputs with at most four elements
            bad = {k: v for k, v in bad.items() if not (np.size(ref_out if k == "out" else getattr(ref_pb, k)) <= 4 and v <= 10 * tol)}
        worst = max(worst, max(v / tol for v in errs.values()))
        if bad:
            print("FAIL", dict(n_in=n_in, n_out=n_out, grid=grid, P=P, B=B, dtype=dtype.__name__, weights=weights, rep=rep, seed=seed, far=far, sliced=sliced, n=n), pf, pp, bad, flush=True)
            sys.exit(1)
    paths[(pf, pp)] = paths.get((pf, pp), 0) + 1
    n += 1
print(f"soak ok: {n} problems in {time.time() - t0:.0f} s, worst error / tolerance {worst:.3f}")
for k, v in sorted(paths.items(), key=lambda kv: -kv[1]): print(f"  {v:4d}  {k[0]} / {k[1]}")
