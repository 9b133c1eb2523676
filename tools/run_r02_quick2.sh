#!/bin/bash
# quick check of a build on the headline config: full parity file + config-2 bench (device-resident).  Usage: bash tools/run_r02_quick2.sh tag
tag=$1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -n 4 > gpurun_out/r02_${tag}_pytest.log 2>&1; tail -2 gpurun_out/r02_${tag}_pytest.log
timeout 300 python bench.py --config cfg2 --steps 20 --no-e2e --no-cpu --no-others > gpurun_out/r02_${tag}_cfg2.json 2> gpurun_out/r02_${tag}_cfg2.err; tail -2 gpurun_out/r02_${tag}_cfg2.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_${tag}_cfg2.json").read().strip().splitlines()[-1])
print("cfg2", round(d["ms_per_step"], 4), {k: round(v, 4) for k, v in d["kernels_ms"].items()})
PY
