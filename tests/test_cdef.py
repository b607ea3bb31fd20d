"""CDEF parity (direction search and the three filter block sizes): product (CUDA, C ABI)
vs the oracle, inputs as in the reference's differential test
(tests/checkasm/cdef.c:42-144: init_tmp fill types, every direction x every edge combination)."""
import ctypes as C

import numpy as np
import pytest

from refharness import ptr


def _fill(rng, n, bdmax, dtype):
    t = int(rng.integers(0, 8))
    if t == 0:
        return rng.integers(0, 2, size=n).astype(dtype)            # underflow probe
    if t == 1:
        return (bdmax - rng.integers(0, 2, size=n)).astype(dtype)  # overflow probe
    return rng.integers(0, bdmax + 1, size=n).astype(dtype)


@pytest.mark.gpu
@pytest.mark.parametrize("bdmax", [255, 1023, 4095])
def test_cdef_dir(rb, ref, bdmax):
    rng = np.random.default_rng(bdmax)
    pdt = np.uint16 if bdmax > 255 else np.uint8
    for trial in range(64):
        src = _fill(rng, 64, bdmax, pdt).reshape(8, 8)
        if trial % 4 == 3:  # structured content: a ramp along some direction plus noise
            yy, xx = np.mgrid[0:8, 0:8]
            a, b = rng.integers(-3, 4, size=2)
            src = np.clip((a * xx + b * yy) * (bdmax // 48) + bdmax // 2 + rng.integers(-2, 3, size=(8, 8)), 0, bdmax).astype(pdt)
        v0, v1, d1 = C.c_uint(0), C.c_uint(0), C.c_int(0)
        d0 = ref.ref_cdef_dir(ptr(src), src.strides[0], C.byref(v0), bdmax)
        rb.check(rb.cdef_dir(ptr(src), src.strides[0], C.byref(v1), bdmax, C.byref(d1)))
        assert (d0, v0.value) == (d1.value, v1.value), (bdmax, trial)


@pytest.mark.gpu
@pytest.mark.parametrize("bdmax", [255, 1023, 4095])
@pytest.mark.parametrize("idx", [0, 1, 2])
def test_cdef_filter(rb, ref, bdmax, idx):
    w, h = (8, 8) if idx == 0 else (4, 8) if idx == 1 else (4, 4)
    rng = np.random.default_rng(bdmax * 3 + idx)
    pdt = np.uint16 if bdmax > 255 else np.uint8
    isz = np.dtype(pdt).itemsize
    bdmin8 = (bdmax + 1).bit_length() - 1 - 8
    stride = 16 * isz
    for s in (1, 2, 3):
        for dir_ in range(8):
            for edges in range(16):
                src = _fill(rng, 16 * 10 + 16, bdmax, pdt)
                top = _fill(rng, 16 * 2 + 16, bdmax, pdt)
                bot = _fill(rng, 16 * 2 + 16, bdmax, pdt)
                left = _fill(rng, 16, bdmax, pdt)
                a, b = src.copy(), src.copy()
                pri = (1 + int(rng.integers(0, 15))) << bdmin8 if s & 2 else 0
                sec = 1 << (int(rng.integers(0, 3)) + bdmin8) if s & 1 else 0
                damping = 3 + int(rng.integers(0, 4)) + bdmin8 - int(w == 4 or bool(rng.integers(0, 2)))
                args = (stride, ptr(left), C.c_void_p(top.ctypes.data + 8 * isz), C.c_void_p(bot.ctypes.data + 8 * isz),
                        pri, sec, dir_, damping, edges, bdmax)
                ref.ref_cdef_fb(idx, C.c_void_p(a.ctypes.data + 8 * isz), *args)
                rb.check(rb.cdef_fb(idx, C.c_void_p(b.ctypes.data + 8 * isz), *args))
                assert np.array_equal(a, b), (bdmax, idx, s, dir_, edges, pri, sec, damping)
