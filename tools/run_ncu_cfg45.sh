# ncu --set full of config 5's and config 4's kernels with the current build (after the plain commands exited 0).
tag=${1:-v10}
mkdir -p gpurun_out
set -x
timeout 100 python bench.py --config cfg5 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_cfg5_$tag.json 2> gpurun_out/bench_cfg5_$tag.err && \
timeout 150 ncu --set full --import-source on --clock-control none -k regex:"pullback_tma2d" -c 1 -o gpurun_out/prof_r01_${tag}_cfg5 -f python bench.py --config cfg5 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full_${tag}_cfg5.log 2>&1
timeout 100 python bench.py --config cfg4 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_cfg4_$tag.json 2> gpurun_out/bench_cfg4_$tag.err && \
timeout 150 ncu --set full --import-source on --clock-control none -k regex:"pullback_gather2d|fwd_tile2d_fast" -c 2 -o gpurun_out/prof_r01_${tag}_cfg4 -f python bench.py --config cfg4 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full_${tag}_cfg4.log 2>&1
ls -la gpurun_out | tail -8
