#!/bin/bash
# one bench line per BASELINE config and README-table row -> gpurun_out/configs_<tag>/ (+ table.md).  Usage: bash tools/run_r02_configs.sh tag
tag=${1:-r02}
d=gpurun_out/configs_$tag; mkdir -p $d
for c in cfg1 cfg2 cfg3 cfg4 cfg5 readme2 readme3 readme4 readme5; do
  timeout 400 python bench.py --config $c --steps 10 --no-e2e --no-others > $d/bench_$c.json 2> $d/bench_$c.err || echo "$c failed"
  python - <<PY
import json
try:
    d = json.loads(open("$d/bench_$c.json").read().strip().splitlines()[-1]); json.dump(d, open("$d/bench_$c.json", "w"))
    print("$c", round(d["ms_per_step"], 4), d["config"]["forward_path"], d["config"]["pullback_path"], (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print("$c parse failed", e)
PY
done
python tools/configs_table.py $d
