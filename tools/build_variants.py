"""Experiment builds of libdpr.so with other 3-d tile shapes (tools/exp/libdpr_<name>.so; not part of the product).
Usage: python tools/build_variants.py name:DEF1=V,DEF2=V ..."""
import importlib.util, os, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("dpr_build", os.path.join(ROOT, "diffpointrasterisation.jl_b200", "build.py"))
b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
def one(arg):
    name, _, defs = arg.partition(":")
    out = os.path.join(ROOT, "tools", "exp", f"libdpr_{name}.so")
    b.build(out=out, defines=tuple(d for d in defs.split(",") if d))
    return out
with ThreadPoolExecutor(4) as ex:
    for o in ex.map(one, sys.argv[1:]): print(o)
