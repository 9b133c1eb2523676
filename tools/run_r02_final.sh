#!/bin/bash
# Evidence run of the round-2 build on one B200: GPU tests, smoke(), the default bench line + reference arm, ncu launch
# lists and --set full captures of the dominant kernels of configs 2, 3 and 5.  Usage: bash tools/run_r02_final.sh <tag>
tag=${1:-r02_final}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -n 4 --durations=8 > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$tag.log; tail -5 gpurun_out/pytest_gpu_$tag.log
timeout 200 python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' > gpurun_out/smoke_$tag.log 2>&1; tail -2 gpurun_out/smoke_$tag.log
timeout 600 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -2 gpurun_out/bench_$tag.err
timeout 400 python bench.py --impl reference > gpurun_out/bench_${tag}_ref.json 2> gpurun_out/bench_${tag}_ref.err; tail -c 600 gpurun_out/bench_${tag}_ref.json
for cfg in cfg2 cfg3 cfg5; do
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}_$cfg.csv python bench.py --config $cfg --steps 2 --warmup 3 --no-e2e --no-cpu --no-others > gpurun_out/ncu_launch_${tag}_$cfg.log 2>&1
done
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"fwd_tile2d_radial|pullback_gather2d" -c 2 -o gpurun_out/prof_${tag}_cfg2 -f python bench.py --config cfg2 --steps 1 --warmup 1 --no-e2e --no-cpu --no-others > gpurun_out/ncu_full_${tag}_cfg2.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"fwd_tile3d|pullback_tile3d|tile_count|tile_scatter|sort_scatter4|hash_inputs|unpermute" -c 11 -o gpurun_out/prof_${tag}_cfg3 -f python bench.py --config cfg3 --steps 1 --warmup 1 --no-e2e --no-cpu --no-others > gpurun_out/ncu_full_${tag}_cfg3.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"pullback_tma2d" -c 1 -o gpurun_out/prof_${tag}_cfg5 -f python bench.py --config cfg5 --steps 1 --warmup 1 --no-e2e --no-cpu --no-others > gpurun_out/ncu_full_${tag}_cfg5.log 2>&1
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
    print("cfg2", round(d["ms_per_step"], 4), "e2e", d["e2e"], {k: round(v, 4) for k, v in d["kernels_ms"].items()})
    for k, v in (d.get("other_configs") or {}).items():
        print(k, {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("ms_per_step", "whole_step_frac", "forward_path", "pullback_path", "error")})
except Exception as e:
    print("bench parse failed", e)
PY
ls -la gpurun_out | grep $tag
