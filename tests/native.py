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
