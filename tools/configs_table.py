"""Build profiles/<dir>/table.md from the bench.py JSON lines of every config.  Usage: python tools/configs_table.py <dir>"""
import json, os, sys
d = sys.argv[1]
order = ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5", "readme2", "readme3", "readme4", "readme5"]
FWD = ("fwd_", "fill_background", "radial_", "chunk_aabb", "point_weight_stats")
rows = ["| config | workload | ms/step | fwd kernels ms | pullback kernels ms | splats/s | whole-step frac of HBM roofline | CPU port splats/s | kernel paths |",
        "|---|---|---|---|---|---|---|---|---|"]
for c in order:
    p = os.path.join(d, f"bench_{c}.json")
    if not os.path.exists(p):
        continue
    j = json.load(open(p))
    k = j["kernels_ms"]
    # sort kernels shared by both passes (bin_*) are attributed by the pass that ran them: split evenly when both sort
    fwd = sum(v for n, v in k.items() if n.startswith(FWD))
    bwd = sum(v for n, v in k.items() if n.startswith(("pullback_", "background_sum", "zero_gradients", "chunk_centroid", "accumulate")))
    rest = sum(k.values()) - fwd - bwd
    if "fwd" in j["config"]["ops"] and ("sorted" in j["config"]["forward_path"] or "culled" in j["config"]["forward_path"]) and "sorted" in j["config"]["pullback_path"]:
        fwd += rest / 2; bwd += rest / 2
    elif "sorted" in j["config"]["pullback_path"]:
        bwd += rest
    else:
        fwd += rest
    cpu = j.get("cpu_baseline") or {}
    rows.append(f"| {c} | {j['config']['workload'].split(': ', 1)[1]} | {j['ms_per_step']:.3f} | {fwd:.3f} | {bwd:.3f} | {j['value']:.3g} | "
                f"{100 * j['roofline']['whole_step']['frac']:.1f} % | " + (f"{cpu['value']:.3g} ({cpu.get('cores', '?')} cores)" if cpu.get('value') else "not re-run") + " | "
                f"{j['config']['forward_path']} / {j['config']['pullback_path']} |")
open(os.path.join(d, "table.md"), "w").write("\n".join(rows) + "\n")
print("\n".join(rows))
