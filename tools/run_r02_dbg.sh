#!/bin/bash
for w in fwd bwd; do for t in f32 f64; do
  timeout 120 python tools/dbg_tile3d.py $w $t > gpurun_out/dbg_${w}_$t.log 2>&1; echo "$w $t exit $?"; grep -E "libdpr|ok" gpurun_out/dbg_${w}_$t.log | head -5
done; done
bash tools/run_r02_c.sh
