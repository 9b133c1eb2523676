"""Multi-process (gloo, world_size 2) test of the pose-sharded driver's host logic on CPU.
The compute is injected from the oracle (tests may use it); the sharding, packing and the single all-reduce are
the code under test."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import make_inputs, rel_l2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import dpr_b200
    from dpr_b200 import sharded
    from oracle import oracle

    grid = (16, 16)
    d = make_inputs(21, 3, 2, 500, B, grid, np.float64)
    f = lambda a: None if a is None else dpr_b200.fortran(torch.from_numpy(np.ascontiguousarray(a)))

    def fwd(grid_size, points, rotation, translation, background, out_weight, point_weight):
        n = lambda t: None if t is None else t.numpy()
        return f(oracle.raster(grid_size, n(points), n(rotation), n(translation), n(background), n(out_weight), n(point_weight)))

    def pb(ds_dout, points, rotation, translation, background, out_weight, point_weight):
        n = lambda t: None if t is None else t.numpy()
        r = oracle.raster_pullback(n(ds_dout), n(points), n(rotation), n(translation), n(background), n(out_weight), n(point_weight))
        return dpr_b200.PullbackResult(*(f(x) for x in r))

    drv = sharded.PoseShardedRaster(forward_fn=fwd, pullback_fn=pb)
    sh = lambda k: sharded.shard_poses(f(d[k]), rank, world)
    out = drv.raster(grid, f(d["points"]), sh("rotation"), sh("translation"), sh("background"), sh("out_weight"), f(d["point_weight"]))
    res, _ = drv.raster_pullback_(sh("ds_dout"), f(d["points"]), sh("rotation"), sh("translation"), sh("background"),
                                  sh("out_weight"), f(d["point_weight"]))
    np.savez(os.path.join(result_dir, f"rank{rank}.npz"), out=out.numpy(), points=res.points.numpy(),
             point_weight=res.point_weight.numpy(), rotation=res.rotation.numpy(), translation=res.translation.numpy(),
             background=res.background.numpy(), out_weight=res.out_weight.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [6, 7])
def test_pose_sharded_two_ranks_match_unsharded_oracle(tmp_path, B):
    from oracle import oracle
    from dpr_b200 import sharded
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), B, str(tmp_path)), nprocs=world, join=True)
    grid = (16, 16)
    d = make_inputs(21, 3, 2, 500, B, grid, np.float64)
    args = tuple(d[k] for k in ("points", "rotation", "translation", "background", "out_weight", "point_weight"))
    out_ref = oracle.raster(grid, *args)
    pb_ref = oracle.raster_pullback(d["ds_dout"], *args)
    r = [np.load(tmp_path / f"rank{i}.npz") for i in range(world)]
    # pose-summed gradients are identical on every rank after the all-reduce and equal the unsharded result
    for i in range(world):
        assert rel_l2(r[i]["points"], pb_ref.points) < 1e-13
        assert rel_l2(r[i]["point_weight"], pb_ref.point_weight) < 1e-13
    # per-pose outputs stay local: concatenating the shards along the batch axis gives the full arrays
    assert rel_l2(np.concatenate([x["out"] for x in r], axis=-1), out_ref) == 0
    for k in ("rotation", "translation", "background", "out_weight"):
        assert rel_l2(np.concatenate([x[k] for x in r], axis=-1), getattr(pb_ref, k)) < 1e-14, k
    assert [sharded.pose_range(B, i, world) for i in range(world)] == ([(0, 3), (3, 6)] if B == 6 else [(0, 4), (4, 7)])


def test_pose_range_partitions():
    from dpr_b200 import sharded
    for B in (0, 1, 5, 16, 4097):
        for w in (1, 2, 3, 8):
            ranges = [sharded.pose_range(B, r, w) for r in range(w)]
            assert ranges[0][0] == 0 and ranges[-1][1] == B
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1
