#!/bin/bash
# config-3 bench (+ tile tests) for each experiment build in tools/exp/.  Usage: bash tools/run_r02_var.sh tag name...
tag=$1; shift
L=diffpointrasterisation.jl_b200/libdpr.so
cp $L /tmp/libdpr_main.so
for v in "$@"; do
  cp tools/exp/libdpr_$v.so $L
  timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tile3d" > gpurun_out/r02_${tag}_${v}_pytest.log 2>&1; tail -1 gpurun_out/r02_${tag}_${v}_pytest.log
  timeout 200 python bench.py --config cfg3 --steps 10 --no-e2e --no-cpu --no-others > gpurun_out/r02_${tag}_${v}_cfg3.json 2> gpurun_out/r02_${tag}_${v}_cfg3.err; tail -2 gpurun_out/r02_${tag}_${v}_cfg3.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r02_${tag}_${v}_cfg3.json").read().strip().splitlines()[-1])
print("$v cfg3", round(d["ms_per_step"], 4), {k: round(x, 4) for k, x in d["kernels_ms"].items()})
PY
done
cp /tmp/libdpr_main.so $L
