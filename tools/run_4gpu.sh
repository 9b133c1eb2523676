# bench.py at 4 GPUs of one box with the driver's launch line (weak scaling).  Usage: gpurun --gpus 4 -- bash tools/run_4gpu.sh <tag>
tag=${1:-v10}
mkdir -p gpurun_out/scale_r01_$tag
set -x
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29604 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu > gpurun_out/scale_r01_$tag/scale_n4.json 2> gpurun_out/scale_r01_$tag/scale_n4.err
python -c "
import json; d=json.load(open('gpurun_out/scale_r01_$tag/scale_n4.json')); print(4, round(d['ms_per_step'],3), '%.3e' % d['value'], 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],1), d['e2e'] and '%.3e' % d['e2e']['value'])"
