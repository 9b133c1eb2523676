"""CPU-only checks of the drop-in boundary: the C-ABI library builds, loads, and exports every symbol that
include/dpr.h declares; host-side argument checking mirrors the reference's errors; no CPU fallback exists."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import dpr_b200
from dpr_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "dpr.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dpr_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    lib = _lib.load()
    declared = _declared_symbols()
    assert declared, "no declarations parsed from include/dpr.h"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/dpr.h but not exported by libdpr.so"
    assert sorted(_lib.SYMBOLS) == declared


def test_library_is_sm100a_only():
    out = os.popen(f"cuobjdump -lelf {_lib.LIB_PATH} 2>/dev/null").read()
    if out.strip():
        assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_fast_forward_kernel_has_no_packed_fma():
    """ptxas 12.9 fuses mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (illegal for .rn, ignores -fmad=false); the cell a
    point lands in depends on the last bit of `coord`, so the fast forward kernel is written so that no such pair
    exists.  Guard: its SASS must not contain a single FFMA2, and must contain the packed multiplies and native ATOMS."""
    sass = os.popen(f"cuobjdump -sass {_lib.LIB_PATH} 2>/dev/null").read()
    if not sass.strip():
        pytest.skip("cuobjdump not available")
    chunks = sass.split("Function : ")
    fast = [c for c in chunks if "fwd_tile2d_fast_kernel" in c.split("\n", 1)[0]]
    assert len(fast) == 12, "expected 12 instantiations of the fast forward kernel"
    for body in fast:
        assert "FFMA2" not in body
        assert "FMUL2" in body and "ATOMS.ADD" in body
    radial = [c for c in chunks if "fwd_tile2d_radial_kernel" in c.split("\n", 1)[0]]
    assert len(radial) == 4      # N_in in {2,3} x point weights
    for body in radial:          # same rule; plus the streaming 16-byte point loads and the native integer atomics
        assert "FFMA2" not in body
        assert "FMUL2" in body and "ATOMS.ADD" in body and "LDG.E.NA." in body
    tma = [c for c in chunks if "pullback_tma2d_kernel" in c.split("\n", 1)[0]]
    assert len(tma) == 16      # N_in in {2,3} x point weights x stages {2,3} x {dense 1-d copies, padded tensor-map copies}
    for body in tma:   # TMA copies + mbarrier pipeline really are in the SASS, and no packed FMA
        assert ("UBLKCP" in body or "UTMALDG" in body) and "SYNCS" in body and "FFMA2" not in body
    assert sum("UTMALDG.3D" in body for body in tma) == 8 and sum("UBLKCP" in body for body in tma) == 8
    t3 = [c for c in chunks if "pullback_tile3d_kernel" in c.split("\n", 1)[0]]
    assert len(t3) == 4        # Float32 / Float64 x {tensor-map TMA tile loads, cooperative loads}
    assert sum("UTMALDG.4D" in body for body in t3) == 2          # cp.async.bulk.tensor.4d of the ds_dout tiles
    f3 = [c for c in chunks if "fwd_tile3d_kernelIf" in c.split("\n", 1)[0]]
    assert len(f3) == 1 and "ATOMS.ADD" in f3[0] and "STG.E" in f3[0]       # native integer shared atomics, 16-byte flush
    pb = [c for c in chunks if "pullback_gather2d_kernelIf" in c.split("\n", 1)[0]]
    assert len(pb) == 8   # N_in in {2,3} x point weights x paired loads
    for body in pb:
        assert "FFMA2" not in body and "FMUL2" in body


def test_status_strings_and_options():
    lib = _lib.load()
    assert lib.dpr_version() >= 100
    assert lib.dpr_status_string(0) == b"ok"
    for code in range(-8, 0):
        assert lib.dpr_status_string(code) not in (b"ok", b"unknown status")
    assert lib.dpr_set_option(99, 1) == -8 and lib.dpr_set_option(0, 7) == -8
    _lib.set_option(_lib.OPT_FORWARD_ALGO, 1)
    assert _lib.get_option(_lib.OPT_FORWARD_ALGO) == 1
    _lib.set_option(_lib.OPT_FORWARD_ALGO, 0)
    assert lib.dpr_workspace_bytes(0, 3, 2, (ctypes.c_int64 * 2)(8, 8), 10, 2, 4) <= 4096
    for bad in ((0, 2, 10, 2, 4), (3, 0, 10, 2, 4), (5, 2, 10, 2, 4), (3, 2, -1, 2, 4), (3, 2, 10, 2, 2)):   # never traps
        assert lib.dpr_workspace_bytes(1, bad[0], bad[1], (ctypes.c_int64 * 2)(8, 8), bad[2], bad[3], bad[4]) == 0


def test_argument_validation_without_gpu():
    """Validation happens before any CUDA call, so the status codes can be checked on a CPU-only box."""
    lib = _lib.load()
    grid = (ctypes.c_int64 * 2)(8, 8)
    f = lib.dpr_raster_forward_f32
    assert f(5, 2, grid, 1, 1, None, None, None, None, None, None, None, None, 0, None) == -2   # unsupported dims
    assert f(2, 5, grid, 1, 1, None, None, None, None, None, None, None, None, 0, None) == -2
    assert f(0, 2, grid, 1, 1, None, None, None, None, None, None, None, None, 0, None) == -2
    for n_in in range(1, 5):                 # every pair up to 4 x 4 passes the dimension check (NULL pointers next)
        for n_out in range(1, 5):
            g4 = (ctypes.c_int64 * 4)(8, 8, 8, 8)
            assert f(n_in, n_out, g4, 1, 1, None, None, None, None, None, None, None, None, 0, None) == -3
    assert f(3, 2, grid, -1, 1, None, None, None, None, None, None, None, None, 0, None) == -1  # bad dims
    assert f(3, 2, (ctypes.c_int64 * 2)(0, 8), 1, 1, None, None, None, None, None, None, None, None, 0, None) == -1
    assert f(3, 2, grid, 1, 1, None, None, None, None, None, None, None, None, 0, None) == -3   # NULL pointers
    assert f(3, 2, None, 1, 1, None, None, None, None, None, None, None, None, 0, None) == -3


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    pts = torch.zeros(3, 4)
    rot = torch.zeros(2, 3, 1)
    tr = torch.zeros(2, 1)
    with pytest.raises(RuntimeError, match="no CPU path"):
        dpr_b200.raster((8, 8), pts, rot, tr)
    with pytest.raises(RuntimeError, match="no CPU path"):
        dpr_b200.raster_pullback_(torch.zeros(8, 8, 1), pts, rot, tr)
    # the C entry point itself reports the missing device instead of computing anything
    lib = _lib.load()
    buf = np.zeros(64, dtype=np.float32)
    p = buf.ctypes.data_as(ctypes.c_void_p)
    rc = lib.dpr_raster_forward_f32(3, 2, (ctypes.c_int64 * 2)(4, 4), 1, 1, p, p, p, None, None, None, p, None, 0, None)
    assert rc in (-6, -5)


def test_fortran_helpers():
    t = dpr_b200.empty_f((3, 5, 2), torch.float32, "cpu")
    assert t.shape == (3, 5, 2) and t.stride() == (1, 3, 15) and dpr_b200.is_fortran(t)
    c = torch.arange(24.0).reshape(2, 3, 4)
    f = dpr_b200.fortran(c)
    assert dpr_b200.is_fortran(f) and torch.equal(f, c)
    assert np.array_equal(np.asarray(f.permute(2, 1, 0).contiguous()).ravel(), np.asarray(c).ravel(order="F"))


def _build_c_abi_smoke():
    """gcc + the CUDA runtime only: no nvcc, no C++ - what a `ccall` (or cgo / JNI) caller links against."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "c_abi_smoke")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    lib_dir = os.path.dirname(_lib.LIB_PATH)
    _lib.load()                                   # builds libdpr.so if it is stale
    cmd = [shutil.which("gcc") or "gcc", "-O1", "-Wall", "-std=c99", "-I", os.path.join(root, "include"), "-I", os.path.join(cuda, "include"),
           "-o", exe, os.path.join(root, "tests", "c_abi_smoke.c"), "-L", lib_dir, "-ldpr", "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lm",
           "-Wl,-rpath," + lib_dir, "-Wl,-rpath," + os.path.join(cuda, "lib64")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def test_c_abi_smoke_builds_with_plain_c():
    assert os.path.exists(_build_c_abi_smoke())


@pytest.mark.gpu
def test_c_abi_smoke_runs():
    """The reference's first known-answer forward (src/raster.jl:143-157) and its pullback from a C program holding raw
    cudaMalloc pointers - the stand-in for the Julia `ccall` glue."""
    import subprocess
    res = subprocess.run([_build_c_abi_smoke()], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "c_abi_smoke ok" in res.stdout, res.stdout + res.stderr
