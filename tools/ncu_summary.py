"""Summarise an .ncu-rep: key metrics per kernel + the hottest SASS lines. Usage: python tools/ncu_summary.py rep [kernel-regex]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'smsp__thread_inst_executed_per_inst_executed.ratio','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
 'l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','lts__t_sectors_op_red.sum','lts__t_sectors_op_atom.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
 'smsp__average_warp_latency_per_inst_issued.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio','smsp__average_warps_issue_stalled_drain_per_issue_active.ratio']
for r in rows[2:]:
    print('=====', r[hdr.index('Kernel Name')][:70], 'grid', r[hdr.index('Grid Size')], 'block', r[hdr.index('Block Size')])
    for k in keys:
        if k in hdr:
            i = hdr.index(k); print(f"  {k:95s} {r[i]:>16s} {units[i]}")
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + sys.argv[2]], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[1]; ia, ie, isamp, ith = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples'), h.index('Avg. Threads Executed')
    data = [(r[ia].strip(), int(r[ie]), int(r[isamp]), float(r[ith])) for r in rows[2:] if len(r) > ie and r[ie].isdigit()]
    tot = sum(d[1] for d in data); ts = sum(d[2] for d in data)
    print("total warp-inst", tot, "samples", ts)
    thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.004
    for i, (s, e, sm, th) in enumerate(data):
        if e > tot * thr or sm > ts * 0.01:
            print(f"{i:4d} inst {e/tot*100:5.2f}% samp {sm/ts*100:5.2f}% thr {th:4.1f}  {s[:100]}")
