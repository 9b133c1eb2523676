#!/bin/bash
# N-GPU checks: tests/multi_gpu_check.py (NCCL + peer-memory all-reduce, sharded pullback) and the bench at N GPUs
N=${1:-2}; tag=${2:-r02}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py > gpurun_out/multi_gpu_check_${tag}_n$N.log 2>&1
tail -3 gpurun_out/multi_gpu_check_${tag}_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${tag}_n$N.json 2> gpurun_out/bench_${tag}_n$N.err
tail -3 gpurun_out/bench_${tag}_n$N.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${tag}_n$N.json").read().strip().splitlines()[-1])
    print("N", d["n_gpus"], "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 2) if d["e2e"] else None)
    print(d["config"]["collective"]); print(d["multi_gpu"]); print(d["other_configs"])
except Exception as e:
    print("bench parse failed", e)
PY
