"""Per-CUDA-source-line instruction / stall-sample totals of one kernel in an .ncu-rep (needs -lineinfo + --import-source on).
Usage: python tools/ncu_lines.py rep kernel-regex [min_pct]"""
import csv, io, subprocess, sys
from collections import defaultdict
rep, kern = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.7
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = defaultdict(lambda: [0, 0, ""])
fname, hdr = None, None
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit(): continue
    ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    if not r[ie].isdigit() or not r[isamp].isdigit(): continue
    key = (fname, int(r[0]))
    agg[key][0] += int(r[ie]); agg[key][1] += int(r[isamp]); agg[key][2] = r[1].strip()
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values()) or 1
print("total warp-inst", tot, "samples", ts)
for (f, ln), (e, sm, src) in sorted(agg.items()):
    if e > tot * thr / 100 or sm > ts * thr / 100:
        print(f"{f:22s}:{ln:4d} inst {e/tot*100:5.2f}% samp {sm/ts*100:5.2f}%  {src[:110]}")
