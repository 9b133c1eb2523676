# Single time-boxed run for the box-staged pullback: whole GPU suite and bench with the box kernel selected automatically
# (DPR_BOX_AUTO=1), then one ncu --set full capture and the launch list.  Usage: bash tools/run_box.sh <tag>
tag=${1:-v11}
mkdir -p gpurun_out
export DPR_BOX_AUTO=1
set -x
timeout 70 python -m pytest tests -m gpu -x -q -n 4 > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$tag.log; tail -5 gpurun_out/pytest_gpu_$tag.log
timeout 40 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_r01_${tag}_resident.json 2> gpurun_out/bench_r01_$tag.err; tail -c 1500 gpurun_out/bench_r01_${tag}_resident.json
timeout 60 ncu --set full --import-source on --clock-control none -k regex:"pullback_box2d" -c 1 -o gpurun_out/prof_r01_$tag -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full_$tag.log 2>&1
timeout 40 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_$tag.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch_$tag.log 2>&1
ls -la gpurun_out | tail -6
