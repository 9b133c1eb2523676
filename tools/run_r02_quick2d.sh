#!/bin/bash
# config-4 bench + full parity file.  Usage: bash tools/run_r02_quick4.sh tag
tag=$1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -n 4 > gpurun_out/r02_${tag}_pytest.log 2>&1; tail -2 gpurun_out/r02_${tag}_pytest.log
for c in cfg2 cfg4 cfg5; do
timeout 300 python bench.py --config $c --steps 20 --no-e2e --no-cpu --no-others > gpurun_out/r02_${tag}_$c.json 2> gpurun_out/r02_${tag}_$c.err; tail -2 gpurun_out/r02_${tag}_$c.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_${tag}_$c.json").read().strip().splitlines()[-1])
print("$c", round(d["ms_per_step"], 4), {k: round(v, 4) for k, v in d["kernels_ms"].items()})
PY
done
