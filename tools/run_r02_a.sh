#!/bin/bash
mkdir -p gpurun_out
bash tools/run_r02_probe_tma2.sh > /dev/null 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tile3d" > gpurun_out/r02_pytest_tile3d.log 2>&1
tail -15 gpurun_out/r02_pytest_tile3d.log
for c in cfg3; do
  timeout 300 python bench.py --config $c --steps 10 --no-e2e --no-cpu > gpurun_out/r02_a_$c.json 2> gpurun_out/r02_a_$c.err
  tail -3 gpurun_out/r02_a_$c.err
done
