#!/bin/bash
# round 2: tensor-map TMA probe + baseline numbers of the round-1 build on the same box
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version --format=csv > gpurun_out/r02_tma_probe.log 2>&1
echo "nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/probe_tma_tensor tools/probe_tma_tensor.cu -ldl" >> gpurun_out/r02_tma_probe.log
timeout 120 ./tools/probe_tma_tensor >> gpurun_out/r02_tma_probe.log 2>&1
echo "exit $?" >> gpurun_out/r02_tma_probe.log
echo "--- round-1 libcu++ probe (tools/probe_tma_box2.cu), rank 2 then 3" >> gpurun_out/r02_tma_probe.log
timeout 60 ./tools/probe_tma_box2 2 >> gpurun_out/r02_tma_probe.log 2>&1; echo "exit $?" >> gpurun_out/r02_tma_probe.log
timeout 60 ./tools/probe_tma_box2 3 >> gpurun_out/r02_tma_probe.log 2>&1; echo "exit $?" >> gpurun_out/r02_tma_probe.log
for c in cfg3 cfg2; do
  timeout 300 python bench.py --config $c --steps 10 --no-e2e --no-cpu > gpurun_out/r02_base_$c.json 2> gpurun_out/r02_base_$c.err
done
tail -3 gpurun_out/r02_tma_probe.log
