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
