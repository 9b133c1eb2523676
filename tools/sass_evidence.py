"""Mnemonic counts per kernel of libdpr.so (cuobjdump -sass) -> profiles/sass_evidence_r02.txt.  Runs without a GPU.
Usage: python tools/sass_evidence.py > profiles/sass_evidence_r02.txt"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "diffpointrasterisation.jl_b200", "libdpr.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTMALDG.3D", "UTMALDG.4D", "UBLKCP", "SYNCS", "LDGSTS", "LDGDEPBAR", "FMUL2", "FADD2", "FFMA2", "ATOMS.ADD", "ATOMS.CAST", "MATCH.ANY",
        "REDG.E.ADD.F32x4", "REDG.E.ADD.F32x2", "REDG.E.ADD.F32.",
        "REDG.E.ADD.F64", "ATOMG", "LDS", "LDG", "LD.E", "STG", "SHFL", "VOTE", "BAR.SYNC"]
print("# SASS evidence per kernel of libdpr.so (cuobjdump -sass, sm_100a, tools/sass_evidence.py); counts of selected mnemonics")
print("# UTMALDG = cp.async.bulk.tensor (tensor-map TMA), UBLKCP = cp.async.bulk (1-d bulk TMA), SYNCS = mbarrier ops, LDGSTS = cp.async (16-byte asynchronous copies), FMUL2/FADD2 = packed FP32x2,")
print("# ATOMS.ADD = native shared int atomic, ATOMS.CAST = shared float CAS loop, REDG / ATOMG ... RZ = global reduction without return")
rows = []
for chunk in sass.split("Function : ")[1:]:
    name, body = chunk.split("\n", 1)
    dem = subprocess.run(["c++filt", name.strip()], capture_output=True, text=True).stdout.strip().split("(")[0]
    ops = re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", body, flags=re.M)
    counts = []
    for k in KEYS:
        n = sum(1 for o in ops if o.startswith(k))
        if n:
            counts.append(f"{k}={n}")
    rows.append(f"{dem[:100]:100s} instr={len(ops):5d}  " + "  ".join(counts))
print("\n".join(sorted(set(rows))))
