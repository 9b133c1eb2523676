#!/bin/bash
# full GPU check of the round-2 build: test-suite, smoke, the default bench line (with other_configs), the reference arm
tag=${1:-r02}
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$tag.log 2>&1
tail -4 gpurun_out/pytest_gpu_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; tail -2 gpurun_out/smoke_$tag.log
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -3 gpurun_out/bench_$tag.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
    print("cfg2", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 2) if d["e2e"] else None, {k: round(v, 4) for k, v in d["kernels_ms"].items()})
    for k, v in (d.get("other_configs") or {}).items():
        print(k, {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("ms_per_step", "whole_step_frac", "forward_path", "pullback_path", "error")})
except Exception as e:
    print("bench parse failed", e)
PY
