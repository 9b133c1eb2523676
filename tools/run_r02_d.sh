#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r02_v4.log 2>&1
tail -4 gpurun_out/pytest_gpu_r02_v4.log
for c in cfg5; do
timeout 300 python bench.py --config $c --steps 10 --no-e2e --no-cpu > gpurun_out/r02_d_$c.json 2> gpurun_out/r02_d_$c.err
tail -2 gpurun_out/r02_d_$c.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_d_$c.json").read().strip().splitlines()[-1])
    print("$c", round(d["ms_per_step"], 4), d["config"]["forward_path"], d["config"]["pullback_path"], {k: round(v, 4) for k, v in d["kernels_ms"].items()})
except Exception as e:
    print("bench parse failed", e)
PY
done
timeout 900 python bench.py --no-others > gpurun_out/bench_r02_v4.json 2> gpurun_out/bench_r02_v4.err; tail -2 gpurun_out/bench_r02_v4.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r02_v4.json").read().strip().splitlines()[-1])
print("cfg2", round(d["ms_per_step"], 4), "e2e", d["e2e"])
PY
