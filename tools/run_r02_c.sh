#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tile3d" > gpurun_out/r02_pytest_tile3d_k.log 2>&1
tail -5 gpurun_out/r02_pytest_tile3d_k.log
timeout 300 python bench.py --config cfg3 --steps 10 --no-e2e --no-cpu > gpurun_out/r02_k_cfg3.json 2> gpurun_out/r02_k_cfg3.err
tail -3 gpurun_out/r02_k_cfg3.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"fwd_tile3d|pullback_tile3d" -c 2 -o gpurun_out/prof_r02_k_cfg3 -f python bench.py --config cfg3 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_r02_k_cfg3.log 2>&1
tail -2 gpurun_out/ncu_r02_k_cfg3.log
