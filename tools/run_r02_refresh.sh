#!/bin/bash
# short evidence refresh of a build: GPU tests, smoke, the default bench line + reference arm.  Usage: bash tools/run_r02_refresh.sh tag
tag=${1:-r02_final3}
timeout 1500 python -m pytest tests -x -q -m gpu -n 4 --durations=8 > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$tag.log; tail -3 gpurun_out/pytest_gpu_$tag.log
timeout 200 python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' > gpurun_out/smoke_$tag.log 2>&1; tail -2 gpurun_out/smoke_$tag.log
timeout 600 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -2 gpurun_out/bench_$tag.err
timeout 400 python bench.py --impl reference > gpurun_out/bench_${tag}_ref.json 2> gpurun_out/bench_${tag}_ref.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
print("cfg2", round(d["ms_per_step"], 4), d["roofline"]["kernel_ms"], d["roofline"]["frac"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["dependent"]["ms_per_step"])
for k, v in (d.get("other_configs") or {}).items():
    print(k, round(v["ms_per_step"], 4), round(v["whole_step_frac"], 4))
PY
