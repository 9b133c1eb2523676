#!/bin/bash
# config-5 check of a build: the parity file, the full-size config-5 case, config 5 bench.  Usage: bash tools/run_r02_quick5.sh tag
tag=$1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -n 4 > gpurun_out/r02_${tag}_pytest.log 2>&1; tail -2 gpurun_out/r02_${tag}_pytest.log
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "cfg5" 2>&1 | tail -1
timeout 300 python bench.py --config cfg5 --steps 20 --no-e2e --no-cpu --no-others > gpurun_out/r02_${tag}_cfg5.json 2> gpurun_out/r02_${tag}_cfg5.err; tail -2 gpurun_out/r02_${tag}_cfg5.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_${tag}_cfg5.json").read().strip().splitlines()[-1])
print("cfg5", round(d["ms_per_step"], 4), {k: round(v, 4) for k, v in d["kernels_ms"].items()}, d["config"]["pullback_path"])
PY
