# 2-GPU check of the final build on one box: NCCL all-reduce + pose-sharded pullback against a single-GPU run, then the
# driver's launch line of bench.py at 1 and 2 GPUs (weak scaling).  Usage: gpurun --gpus 2 -- bash tools/run_2gpu.sh <tag>
tag=${1:-v10}
mkdir -p gpurun_out/scale_r01_$tag
set -x
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py > gpurun_out/scale_r01_$tag/multi_gpu_check.log 2>&1; tail -2 gpurun_out/scale_r01_$tag/multi_gpu_check.log
timeout 120 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu > gpurun_out/scale_r01_$tag/scale_n1.json 2> gpurun_out/scale_r01_$tag/scale_n1.err
timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/scale_r01_$tag/scale_n2.json 2> gpurun_out/scale_r01_$tag/scale_n2.err
for n in 1 2; do python -c "
import json; d=json.load(open('gpurun_out/scale_r01_$tag/scale_n$n.json')); print($n, round(d['ms_per_step'],3), '%.3e' % d['value'], 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],1), d['e2e'] and '%.3e' % d['e2e']['value'])"; done
