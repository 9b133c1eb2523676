# Quick verification of a build on one B200: GPU tests (4 workers), smoke(), accuracy report, one resident-only bench line.
tag=${1:-v10}
mkdir -p gpurun_out
set -x
timeout 300 python -m pytest tests -m gpu -x -q -n 4 > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$tag.log; tail -4 gpurun_out/pytest_gpu_$tag.log
timeout 60 python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' > gpurun_out/smoke_$tag.log 2>&1; tail -2 gpurun_out/smoke_$tag.log
timeout 240 python tools/accuracy_report.py > gpurun_out/accuracy_r01_$tag.txt 2>&1; tail -3 gpurun_out/accuracy_r01_$tag.txt
timeout 120 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_r01_${tag}_resident.json 2> gpurun_out/bench_r01_$tag.err; tail -c 400 gpurun_out/bench_r01_${tag}_resident.json
