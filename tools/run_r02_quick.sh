#!/bin/bash
# quick check of a build: 3-d tile tests + host-buffer tests, config-3 bench.  Usage: bash tools/run_r02_quick.sh tag
tag=$1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tile3d or host or cache" > gpurun_out/r02_${tag}_pytest.log 2>&1; tail -2 gpurun_out/r02_${tag}_pytest.log
timeout 300 python bench.py --config cfg3 --steps 10 --no-e2e --no-cpu --no-others > gpurun_out/r02_${tag}_cfg3.json 2> gpurun_out/r02_${tag}_cfg3.err; tail -2 gpurun_out/r02_${tag}_cfg3.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_${tag}_cfg3.json").read().strip().splitlines()[-1])
print("cfg3", round(d["ms_per_step"], 4), {k: round(v, 4) for k, v in d["kernels_ms"].items()})
PY
