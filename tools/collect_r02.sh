#!/bin/bash
# Copies the evidence of a tools/run_r02_final.sh run (gpurun_out/, scratch) into profiles/ (tracked).  Usage: bash tools/collect_r02.sh <tag>
tag=${1:-r02_final}
cp gpurun_out/bench_$tag.json profiles/bench_r02_final.json
cp gpurun_out/bench_${tag}_ref.json profiles/bench_r02_final_ref.json
cp gpurun_out/pytest_gpu_$tag.log profiles/pytest_gpu_r02_final.log
cp gpurun_out/smoke_$tag.log profiles/smoke_r02_final.log
for c in cfg2 cfg3 cfg5; do
  cp gpurun_out/launches_${tag}_$c.csv profiles/launches_r02_$c.csv
  [ -f gpurun_out/prof_${tag}_$c.ncu-rep ] && python tools/ncu_to_profiles.py gpurun_out/prof_${tag}_$c.ncu-rep profiles/ncu_full_r02_${c}_summary.csv > profiles/ncu_full_r02_${c}_traffic.txt
done
cp gpurun_out/probe_pcie_r02.log profiles/probe_pcie_r02.log 2>/dev/null
python tools/sass_evidence.py > profiles/sass_evidence_r02.txt
