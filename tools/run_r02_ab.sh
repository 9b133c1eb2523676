#!/bin/bash
# A/B of experiment builds on one config, alternating, n rounds.  Usage: bash tools/run_r02_ab.sh cfg rounds name...
cfg=$1; n=$2; shift 2
L=diffpointrasterisation.jl_b200/libdpr.so
cp $L /tmp/libdpr_main.so
for r in $(seq 1 $n); do for v in "$@"; do
  cp tools/exp/libdpr_$v.so $L
  timeout 300 python bench.py --config $cfg --steps 20 --no-e2e --no-cpu --no-others > /tmp/ab.json 2> /tmp/ab.err
  python - <<PY
import json
d = json.loads(open("/tmp/ab.json").read().strip().splitlines()[-1])
print("$v $cfg", round(d["ms_per_step"], 4), {k: round(x, 4) for k, x in d["kernels_ms"].items() if x > 0.1})
PY
done; done
cp /tmp/libdpr_main.so $L
