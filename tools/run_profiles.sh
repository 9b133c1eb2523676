set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01_v4.json 2> gpurun_out/bench_r01_v4.err; tail -c 600 gpurun_out/bench_r01_v4.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_v4.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch_v4.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"fwd_tile2d_radial|pullback_gather2d|background_sum" -c 3 -o gpurun_out/prof_r01_v4 -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full_v4.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"pullback_win2d|fwd_tile2d_fast" -c 2 -o gpurun_out/prof_r01_v4_cfg4 -f python bench.py --config cfg4 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full_v4_cfg4.log 2>&1
for c in cfg1 cfg2 cfg3 cfg4 cfg5 readme2 readme3 readme4 readme5; do python bench.py --config $c --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_$c.json 2>> gpurun_out/bench_cfgs.err; done
ls -la gpurun_out | tail -20
