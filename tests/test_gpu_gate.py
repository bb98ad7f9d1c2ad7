"""Exhaustive proof (all 2^32 float bit patterns, on the device) that the product's angle gate and the oracle's
restatement equal the reference expression acos(dot)*180.0f/M_PI < 90 (lesson_16.cu:666-676)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_gate_exhaustive_on_device():
    from tests import native
    native.build()
    L = native.lib()
    out = (C.c_ulonglong * 6)()
    assert L.m3dtest_gate_exhaustive(out) == 0
    mm_product, mm_restated, accepted, lo, hi, first_bad = list(out)
    assert mm_product == 0 and mm_restated == 0, (mm_product, mm_restated, hex(first_bad - 1) if first_bad else None)
    # accepted set is one contiguous interval of positive floats [T, 1.0]
    assert hi == 0x3F800000 and lo == 0x328885AC
    assert accepted == hi - lo + 1
    t = np.array([lo], dtype=np.uint32).view(np.float32)[0]
    assert 0 < t < 1e-6
    print(f"angle gate accepts dot in [{t!r} (0x{lo:08X}), 1.0]")


def test_gate_oracle_matches_device_windows(oracle):
    """The CPU oracle's orc_angle_gate against the device evaluation of the reference expression on windows around
    every decision boundary plus a strided sample of all bit patterns."""
    from tests import native
    native.build()
    L = native.lib()
    gate = oracle.lib().orc_angle_gate
    gate_v = np.vectorize(lambda f: gate(float(f)), otypes=[np.uint8])
    windows = [(0x00000000, 1 << 12), (0x3F0F5C29 - 2048, 4096), (0x3F800000 - 2048, 4096), (0x80000000, 1 << 12),
               (0xBF0F5C29 - 2048, 4096), (0xBF800000 - 2048, 4096), (0x7F800000 - 16, 64), (0x33000000, 1 << 14)]
    out = (C.c_ulonglong * 6)()
    assert L.m3dtest_gate_exhaustive(out) == 0
    lo = int(out[3])
    windows.append((lo - 4096, 8192))
    for base, cnt in windows:
        buf = np.zeros(cnt, dtype=np.uint8)
        assert L.m3dtest_gate_window(C.c_uint32(base), C.c_uint32(cnt), buf.ctypes.data_as(C.c_void_p)) == 0
        vals = (np.arange(cnt, dtype=np.uint64) + base).astype(np.uint32).view(np.float32)
        assert np.array_equal(gate_v(vals), buf), hex(base)
