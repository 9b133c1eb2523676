#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tile3d" > gpurun_out/r02_pytest_tile3d_c.log 2>&1
tail -5 gpurun_out/r02_pytest_tile3d_c.log
timeout 300 python bench.py --config cfg3 --steps 10 --no-e2e --no-cpu > gpurun_out/r02_c_cfg3.json 2> gpurun_out/r02_c_cfg3.err
tail -3 gpurun_out/r02_c_cfg3.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"fwd_tile3d|pullback_tile3d" -c 2 -o gpurun_out/prof_r02_c_cfg3 -f python bench.py --config cfg3 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_r02_c_cfg3.log 2>&1
tail -2 gpurun_out/ncu_r02_c_cfg3.log
# does a CUTLASS-built tensor-map TMA kernel (vLLM's cutlass_scaled_mm, UTMALDG in its sm_100a SASS) run on this box?
timeout 400 python - > gpurun_out/r02_tma_vllm.log 2>&1 <<'PY'
import torch
try:
    from vllm import _custom_ops as ops
    a = (torch.randn(256, 512, device="cuda")).to(torch.float8_e4m3fn)
    b = (torch.randn(512, 256, device="cuda")).to(torch.float8_e4m3fn).t().contiguous().t()
    sa = torch.ones(1, device="cuda"); sb = torch.ones(1, device="cuda")
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        c = ops.cutlass_scaled_mm(a, b, sa, sb, torch.bfloat16)
        torch.cuda.synchronize()
    print("cutlass_scaled_mm ok", float(c.float().abs().mean()))
    for e in prof.key_averages(): print("kernel:", e.key[:160])
except Exception as e:
    import traceback; traceback.print_exc()
PY
tail -5 gpurun_out/r02_tma_vllm.log
