# Time-boxed evidence run for one gpurun call (the round's last GPU minutes): GPU tests (4 workers), the plain bench,
# then the ncu launch list and one --set full capture of config 2's two kernels.  Usage: bash tools/run_final.sh <tag>
tag=${1:-v9}
mkdir -p gpurun_out
set -x
timeout 300 python -m pytest tests -m gpu -x -q -n 4 --durations=8 > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$tag.log; tail -14 gpurun_out/pytest_gpu_$tag.log
timeout 60 python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' > gpurun_out/smoke_$tag.log 2>&1; tail -2 gpurun_out/smoke_$tag.log
timeout 200 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01_$tag.json 2> gpurun_out/bench_r01_$tag.err; tail -c 900 gpurun_out/bench_r01_$tag.json
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_$tag.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch_$tag.log 2>&1
timeout 180 ncu --set full --import-source on --clock-control none -k regex:"fwd_tile2d_radial|pullback_gather2d" -c 2 -o gpurun_out/prof_r01_$tag -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full_$tag.log 2>&1
ls -la gpurun_out | tail -12
