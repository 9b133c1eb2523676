# End-of-session evidence run on one B200 (under gpurun): tests, bench, ncu launch list + --set full captures, all configs.
# Usage: bash tools/run_profiles.sh <tag>      (results under gpurun_out/, copied to profiles/ by hand)
tag=${1:-v6}
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01_$tag.json 2> gpurun_out/bench_r01_$tag.err; tail -c 600 gpurun_out/bench_r01_$tag.json
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_r01_${tag}_ref.json 2>> gpurun_out/bench_r01_$tag.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_$tag.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch_$tag.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"fwd_tile2d_radial|pullback_gather2d|background_sum" -c 3 -o gpurun_out/prof_r01_$tag -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full_$tag.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"pullback_gather2d|fwd_tile2d_fast" -c 2 -o gpurun_out/prof_r01_${tag}_cfg4 -f python bench.py --config cfg4 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full_${tag}_cfg4.log 2>&1
for c in cfg1 cfg2 cfg3 cfg4 cfg5 readme2 readme3 readme4 readme5; do python bench.py --config $c --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_$c.json 2>> gpurun_out/bench_cfgs.err; done
ls -la gpurun_out | tail -12
