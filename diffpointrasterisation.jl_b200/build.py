"""Builds libdpr.so (the sm_100a CUDA library behind include/dpr.h) in-tree with nvcc."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdpr.so")
SOURCES = ["dpr_api.cu", "dpr_forward.cu", "dpr_pullback.cu", "dpr_comm.cu"]
HEADERS = ["dpr_common.cuh", "dpr_forward_fast.cuh", "dpr_forward_radial.cuh", "dpr_pullback_fast.cuh", "dpr_pullback_tma.cuh",
           "dpr_sort.cuh", "dpr_tile3d.cuh", "dpr_internal.h", os.path.join("..", "..", "include", "dpr.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=true",
    "-Xcompiler", "-fPIC,-O2,-Wall", "-shared", "-cudart", "shared", "-ldl", "--threads", "4",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libdpr.so cannot be built (there is no CPU fallback)")
    return exe


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, out: str | None = None, defines: tuple = ()) -> str:
    """out / defines: experiment builds (tools/): another output file, extra -D flags; the product is always LIB."""
    if out is None and not force and not needs_build():
        return LIB
    cmd = ([nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-D" + d for d in defines] + ["-o", out or LIB]
           + [os.path.join(CSRC, s) for s in SOURCES])
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libdpr.so")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return out or LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
