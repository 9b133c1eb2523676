"""Run under torchrun on >= 2 GPUs: checks the library's NCCL all-reduce and the pose-sharded pullback against a
single-GPU run of the whole batch.  (Not a pytest file: the driver's -m gpu run has one GPU.)
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dpr_b200  # noqa: E402
from dpr_b200 import sharded  # noqa: E402
from tests.helpers import make_inputs, rel_l2  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
grid, B = (64, 64), 37
d = make_inputs(99, 3, 2, 30000, B, grid, np.float32)
f = lambda a: None if a is None else dpr_b200.fortran(torch.from_numpy(np.ascontiguousarray(a)).to(dev))
full = {k: f(v) for k, v in d.items() if k != "grid"}
ref = dpr_b200.raster_pullback_(full["ds_dout"], full["points"], full["rotation"], full["translation"], full["background"],
                                full["out_weight"], full["point_weight"])
comm = sharded.DprComm(dev)
t = torch.full((1000,), float(rank + 1), device=dev)
comm.all_reduce_(t)
torch.cuda.synchronize()
assert torch.all(t == sum(range(1, world + 1))), "dpr_comm_allreduce_sum_f32 wrong"
# the library's all-reduce against torch.distributed's: sizes that are / are not multiples of 16 bytes, both element
# types, a payload above the peer-memory capacity (NCCL path), and 60 calls back to back (the two symmetric buffers
# alternate, a rank may be a call ahead of its peers)
gen = torch.Generator(device=dev).manual_seed(100 + rank)
# (16 MB on >= 4 ranks takes the two-shot kernels - reduce-scatter + all-gather; `off` = 1 gives a buffer that is not
# 16-byte aligned)
for dtype, n, off in ((torch.float32, 400_003, 0), (torch.float32, 4 * 1_000_000, 0), (torch.float64, 300_001, 0),
                      (torch.float32, 5_000_000, 0), (torch.float64, 2_000_000, 0), (torch.float32, 3_999_998, 1),
                      (torch.float64, 1_999_999, 1)):
    x = torch.randn(n + off, generator=gen, device=dev, dtype=dtype)[off:]
    want = x.clone()
    dist.all_reduce(want)
    got = torch.empty(n + off, device=dev, dtype=dtype)[off:]
    got.copy_(x)
    comm.all_reduce_(got)
    torch.cuda.synchronize()
    assert rel_l2(got.cpu().numpy(), want.cpu().numpy()) < (1e-6 if dtype == torch.float32 else 1e-14), (dtype, n)
    ref_bits = got.clone()
    dist.broadcast(ref_bits, src=0)
    assert torch.equal(ref_bits, got), "ranks disagree bit-wise"            # same summation order on every rank
acc = torch.zeros(100_001, device=dev)
for it in range(60):
    y = torch.full((100_001,), float(it * world + rank), device=dev)
    comm.all_reduce_(y)
    acc += y
torch.cuda.synchronize()
want = sum(float(it * world + r) for it in range(60) for r in range(world))
assert torch.all(acc == want), "back-to-back all-reduces wrong"
if rank == 0:
    print("peer-memory all-reduce:", comm.uses_peer_memory)
for c in (comm, None):
    drv = sharded.PoseShardedRaster(comm=c)
    sh = lambda k: None if full[k] is None else dpr_b200.fortran(sharded.shard_poses(full[k], rank, world))
    res, _ = drv.raster_pullback_(sh("ds_dout"), full["points"], sh("rotation"), sh("translation"), sh("background"),
                                  sh("out_weight"), full["point_weight"])
    torch.cuda.synchronize()
    assert rel_l2(res.points.cpu().numpy(), ref.points.cpu().numpy()) < 1e-5
    assert rel_l2(res.point_weight.cpu().numpy(), ref.point_weight.cpu().numpy()) < 1e-5
    lo, hi = sharded.pose_range(B, rank, world)
    assert rel_l2(res.rotation.cpu().numpy(), ref.rotation[..., lo:hi].cpu().numpy()) < 1e-5
comm.close()
dist.barrier()
if rank == 0:
    print("multi_gpu_check ok: world", world)
dist.destroy_process_group()
