import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module("mandala-mapping_b200")


@pytest.fixture(scope="session")
def synth(pkg):
    return pkg.synth


@pytest.fixture(scope="session")
def oracle():
    import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def ctx(pkg):
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    c = pkg.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def ref(oracle):
    """The reference's own kernels (oracle/_ref/libm3dref.so); skip where it was not built."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libm3dref.so not built")
    r = oracle.ref()
    if r.ref_device_count() <= 0:
        pytest.skip("no CUDA device for the reference kernels")
    return r


@pytest.fixture(scope="session")
def hdl_pair_small(synth):
    """16k-point HDL-32E-like pair."""
    return synth.scan_pair("hdl32", seed=11, n_azimuth=512)


@pytest.fixture(scope="session")
def hdl_pair(synth):
    """C1: 65 536-point HDL-32E-like pair."""
    return synth.scan_pair("hdl32", seed=42)


def dev(arr):
    """numpy structured array -> torch uint8 CUDA tensor holding the same bytes."""
    import torch
    return torch.from_numpy(np.frombuffer(arr.tobytes(), dtype=np.uint8).copy()).cuda()


def host(t, dtype, n):
    return np.frombuffer(t.cpu().numpy().tobytes(), dtype=dtype, count=n).copy()
