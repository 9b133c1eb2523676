# One bench line per BASELINE config / README-table row with the current build, the reference arm, the accuracy report.
# Usage (one gpurun call): bash tools/run_configs.sh <tag>
tag=${1:-v9}
mkdir -p gpurun_out/configs_r01_$tag
set -x
for c in cfg1 cfg2 cfg3 cfg4 cfg5 readme2 readme3 readme4 readme5; do timeout 120 python bench.py --config $c --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/configs_r01_$tag/bench_$c.json 2>> gpurun_out/configs_r01_$tag/bench_cfgs.err; done
timeout 120 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_r01_${tag}_ref.json 2>> gpurun_out/configs_r01_$tag/bench_cfgs.err
timeout 240 python tools/accuracy_report.py > gpurun_out/accuracy_r01_$tag.txt 2>&1
tail -5 gpurun_out/accuracy_r01_$tag.txt; ls -la gpurun_out/configs_r01_$tag
