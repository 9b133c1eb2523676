#!/bin/bash
# the 3-d pullback's ncu capture and a second default bench line.  Usage: bash tools/run_r02_extra.sh tag
tag=${1:-r02_final}
timeout 600 python bench.py > gpurun_out/bench_${tag}_b.json 2> gpurun_out/bench_${tag}_b.err; tail -2 gpurun_out/bench_${tag}_b.err
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"pullback_tile3d|unpermute" -c 2 -o gpurun_out/prof_${tag}_cfg3_pullback -f python bench.py --config cfg3 --steps 1 --warmup 1 --no-e2e --no-cpu --no-others > gpurun_out/ncu_full_${tag}_cfg3_pullback.log 2>&1
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${tag}_b.json").read().strip().splitlines()[-1])
print("cfg2", round(d["ms_per_step"], 4), d["roofline"]["kernel_ms"], d["roofline"].get("kernel_ms_median"), d["roofline"]["frac"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["dependent"]["ms_per_step"])
PY
