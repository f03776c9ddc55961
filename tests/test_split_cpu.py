"""The static sample split over devices: api.cu::split_rows against a pure-Python restatement of
the reference's distribute() (src/private.h:240-273), through the library's C ABI (no GPU needed)."""
import ctypes
import math
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def distribute_ref(amount, size_each, ndev):
    """reference src/private.h:240-273: equal shares whose starts are 512-byte aligned without breaking an element"""
    if ndev == 0:
        return []
    if ndev == 1:
        return [(0, amount)]
    a, b = size_each, 512
    while True:          # the reference's own gcd loop (not math.gcd: it starts with b %= a)
        if a == 0:
            gcd = b
            break
        b %= a
        if b == 0:
            gcd = a
            break
        a %= b
    stride = 512 // gcd
    offset, res = 0, []
    for i in range(ndev - 1):
        step = np.float32(np.float32(amount - offset) / np.float32(ndev - i))
        ln = int(np.float32(np.round(np.float32(step / np.float32(stride)))) * stride)   # roundf: half away from zero
        q = np.float32(step / np.float32(stride))
        ln = int(math.floor(float(q) + 0.5)) * stride if q >= 0 else ln
        ln = min(ln, amount - offset)
        res.append((offset, ln))
        offset += ln
    res.append((offset, amount - offset))
    return res


@pytest.fixture(scope="module")
def lib():
    import kmcuda_b200
    lb = kmcuda_b200._lib
    lb.kmcuda_b200_debug_split_rows.restype = ctypes.c_int32
    lb.kmcuda_b200_debug_split_rows.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p]
    return lb


@pytest.mark.parametrize("ndev", [1, 2, 3, 4, 7, 8])
@pytest.mark.parametrize("amount,row_bytes", [(8000000, 1024), (13000, 8), (100000, 1024), (4000000, 960),
                                              (999999, 12), (1024, 4), (5, 1024), (3000000, 1024), (77, 40)])
def test_split_rows_equals_reference_distribute(lib, amount, row_bytes, ndev):
    out = np.zeros(2 * ndev, np.uint32)
    assert lib.kmcuda_b200_debug_split_rows(amount, row_bytes, ndev, out.ctypes.data) == 0
    got = [(int(out[2 * i]), int(out[2 * i + 1])) for i in range(ndev)]
    assert got == distribute_ref(amount, row_bytes, ndev)
    # the properties the multi-GPU drivers rely on: a partition of [0, amount), aligned starts
    assert got[0][0] == 0 and sum(ln for _, ln in got) == amount
    for (o0, l0), (o1, _) in zip(got, got[1:]):
        assert o0 + l0 == o1
    for off, ln in got[:-1]:
        assert (off * row_bytes) % 512 == 0 or ln == 0 or off + ln == amount
