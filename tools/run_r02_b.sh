#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_tma_probe3.log
echo "probe_tma_tensor2 modes 5 (L2 cache hint form), 6 (shared::cta destination), 7 (explicit 1x1x1 cluster launch)" > $L
for m in 5 6 7; do timeout 60 ./tools/probe_tma_tensor2 $m >> $L 2>&1; echo "exit $?" >> $L; done
cat $L
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_gpu_b.log 2>&1
tail -5 gpurun_out/r02_pytest_gpu_b.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"tile3|bin_count|scan_lookback|sort_scatter4|tile_bin|unpermute" -c 16 -o gpurun_out/prof_r02_b_cfg3 -f python bench.py --config cfg3 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_r02_b_cfg3.log 2>&1
tail -3 gpurun_out/ncu_r02_b_cfg3.log
