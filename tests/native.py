"""Build/load the TEST-ONLY native helper (tests/csrc/m3dtest.cu)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_build", "libm3dtest.so")
SRC = os.path.join(_HERE, "csrc", "m3dtest.cu")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def build(force: bool = False) -> str:
    hdr = os.path.join(os.path.dirname(_HERE), "mandala-mapping_b200", "csrc", "m3dreg_kernels.cuh")
    stale = (not os.path.exists(SO)) or any(os.path.getmtime(f) > os.path.getmtime(SO) for f in (SRC, hdr))
    if force or stale:
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call([NVCC, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
                               "-Xcompiler", "-fPIC", "-shared", "-o", SO, SRC])
    return SO


def lib() -> C.CDLL:
    return C.CDLL(SO)

SHIM_EXE = os.path.join(_HERE, "_build", "shim_host")
SHIM_SRC = os.path.join(_HERE, "csrc", "shim_host.cpp")


def build_shim_host(force: bool = False) -> str:
    """C++ host program over include/cuda_wrapper_shim.hpp, linked against the product library (plain g++, no CUDA headers)."""
    root = os.path.dirname(_HERE)
    inc = os.path.join(root, "include")
    libdir = os.path.join(root, "mandala-mapping_b200")
    deps = [SHIM_SRC, os.path.join(inc, "cuda_wrapper_shim.hpp"), os.path.join(inc, "m3dreg.h"), os.path.join(libdir, "libm3dreg.so")]
    stale = (not os.path.exists(SHIM_EXE)) or any(os.path.getmtime(f) > os.path.getmtime(SHIM_EXE) for f in deps)
    if force or stale:
        os.makedirs(os.path.dirname(SHIM_EXE), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-Wextra", "-ffp-contract=off", "-I", inc, "-o", SHIM_EXE, SHIM_SRC,
                               "-L", libdir, "-lm3dreg", "-Wl,-rpath," + libdir])
    return SHIM_EXE


def build_shim_program(name: str, force: bool = False) -> str:
    """Any tests/csrc/<name>.cpp over the shim, linked against the product library (plain g++, no CUDA headers)."""
    root = os.path.dirname(_HERE)
    inc = os.path.join(root, "include")
    libdir = os.path.join(root, "mandala-mapping_b200")
    src, exe = os.path.join(_HERE, "csrc", name + ".cpp"), os.path.join(_HERE, "_build", name)
    deps = [src, os.path.join(inc, "cuda_wrapper_shim.hpp"), os.path.join(inc, "m3dreg.h"), os.path.join(libdir, "libm3dreg.so")]
    stale = (not os.path.exists(exe)) or any(os.path.getmtime(f) > os.path.getmtime(exe) for f in deps)
    if force or stale:
        os.makedirs(os.path.dirname(exe), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-Wextra", "-ffp-contract=off", "-I", inc, "-o", exe, src,
                               "-L", libdir, "-lm3dreg", "-Wl,-rpath," + libdir])
    return exe


EMUL_SO = os.path.join(_HERE, "_build", "libnn_emul.so")
EMUL_SRC = os.path.join(_HERE, "csrc", "nn_emul.cpp")


def build_nn_emul(force: bool = False) -> str:
    """Host instantiation of the device NN search logic (nn_core.cuh) for CPU-side checks against the oracle."""
    root = os.path.dirname(_HERE)
    core = os.path.join(root, "mandala-mapping_b200", "csrc", "nn_core.cuh")
    deps = [EMUL_SRC, core, os.path.join(root, "include", "m3dreg.h")]
    stale = (not os.path.exists(EMUL_SO)) or any(os.path.getmtime(f) > os.path.getmtime(EMUL_SO) for f in deps)
    if force or stale:
        os.makedirs(os.path.dirname(EMUL_SO), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared",
                               "-x", "c++", "-I", "/usr/local/cuda/include", "-o", EMUL_SO, EMUL_SRC])
    return EMUL_SO


def nn_emul_search(first, second, table, buckets, gp, radius, max_inner=100, max_outer=100, prune=True):
    """Run the emulated search; returns (nn, candidate evaluations)."""
    import numpy as np
    lib = C.CDLL(build_nn_emul())
    first = np.ascontiguousarray(first); second = np.ascontiguousarray(second)
    table = np.ascontiguousarray(table); buckets = np.ascontiguousarray(buckets); gp = np.ascontiguousarray(gp)
    nn = np.full(len(second), -7, dtype=np.int32)
    ev = C.c_longlong(0)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.emul_nn_search(vp(first), C.c_int(len(first)), vp(second), C.c_int(len(second)), vp(table), vp(buckets), vp(gp),
                            C.c_float(radius), C.c_int(max_inner), C.c_int(max_outer), C.c_int(1 if prune else 0), vp(nn), C.byref(ev))
    assert rc == 0
    return nn, int(ev.value)


def nn_emul_search_warp(first, second, table, buckets, gp, radius, cap=100, prune=True, skip_old_hull=False):
    """Warp-level emulation of the warp-shared search (tests/csrc/nn_emul.cpp): k_nn_search_hull, or round 1's
    k_nn_search_grid with skip_old_hull; returns (nn, fallback queries, re-scans)."""
    import numpy as np
    lib = C.CDLL(build_nn_emul())
    lib.emul_set_skip_old_hull(C.c_int(1 if skip_old_hull else 0))
    first = np.ascontiguousarray(first); second = np.ascontiguousarray(second)
    table = np.ascontiguousarray(table); buckets = np.ascontiguousarray(buckets); gp = np.ascontiguousarray(gp)
    nn = np.full(len(second), -7, dtype=np.int32)
    fb, rs = C.c_longlong(0), C.c_longlong(0)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.emul_nn_search_warp(vp(first), C.c_int(len(first)), vp(second), C.c_int(len(second)), vp(table), vp(buckets), vp(gp),
                                 C.c_float(radius), C.c_int(cap), C.c_int(1 if prune else 0), vp(nn), C.byref(fb), C.byref(rs))
    assert rc == 0
    return nn, int(fb.value), int(rs.value)
