"""Loop-restoration parity (Wiener 7/5-tap, self-guided 5x5 / 3x3 / mix): product (CUDA,
C ABI) vs the oracle, inputs as in the reference's differential test
(tests/checkasm/looprestoration.c:41-196: checkerboard + noise, all 16 edge combinations,
w = 1..384, h = 1..64)."""
import ctypes as C

import numpy as np
import pytest

from refharness import ptr

SGR_PARAMS = {14: (56, 0), 10: (0, 2589), 0: (140, 3236)}   # dav1d_sgr_params rows used by checkasm


def _init_tmp(rng, h, w, bdmax, dtype):
    noise = bdmax >> 4
    xo, yo = rng.integers(0, 8, size=2)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.where(((xx + xo) ^ (yy + yo)) & 8, bdmax, 0)
    return (base ^ rng.integers(0, noise + 1, size=(h, w))).astype(dtype)


def _run(rb, ref, rng, kind, params, bdmax, sizes):
    pdt = np.uint16 if bdmax > 255 else np.uint8
    isz = np.dtype(pdt).itemsize
    src = _init_tmp(rng, 64 + 1, 448, bdmax, pdt).reshape(-1)[:448 * 64 + 64]
    edge = _init_tmp(rng, 8 + 1, 448, bdmax, pdt).reshape(-1)[:448 * 8 + 64]
    left = _init_tmp(rng, 64, 4, bdmax, pdt)
    stride = 448 * isz
    for edges in range(16):
        base_w, base_h = sizes
        w = 256 if edges & 2 else base_w
        h = 64 if edges & 8 else base_h
        a, b = src.copy(), src.copy()
        args = (stride, ptr(left), C.c_void_p(edge.ctypes.data + 64 * isz), w, h)
        ref.ref_lr(kind, C.c_void_p(a.ctypes.data + 64 * isz), *args, C.addressof(params), edges, bdmax)
        rb.check(rb.lr(kind, C.c_void_p(b.ctypes.data + 64 * isz), *args, C.byref(params), edges, bdmax))
        if not np.array_equal(a, b):
            bad = np.nonzero(a != b)[0] - 64
            raise AssertionError(f"kind {kind} bdmax {bdmax} {w}x{h} edges {edges:04b}: {bad.size} px differ, "
                                 f"first (x={bad[0] % 448}, y={bad[0] // 448})")


@pytest.mark.gpu
@pytest.mark.parametrize("bdmax", [255, 1023, 4095])
@pytest.mark.parametrize("taps5", [0, 1])
def test_wiener(rb, ref, bdmax, taps5):
    rng = np.random.default_rng(bdmax * 2 + taps5)
    for trial in range(4):
        p = rb.LrParams()
        for d in range(2):
            f0 = 0 if taps5 else int(rng.integers(0, 16)) - 5
            f1 = int(rng.integers(0, 32)) - 23
            f2 = int(rng.integers(0, 64)) - 17
            p.filter[d][0] = p.filter[d][6] = f0
            p.filter[d][1] = p.filter[d][5] = f1
            p.filter[d][2] = p.filter[d][4] = f2
            p.filter[d][3] = (128 if d else 0) - (f0 + f1 + f2) * 2
        if bdmax > 255:
            p.filter[0][3] += 128
        sizes = (1 + int(rng.integers(0, 384)), 1 + int(rng.integers(0, 64)))
        if trial == 0:
            sizes = (384, 64)
        if trial == 1:
            sizes = (1, 1)
        _run(rb, ref, rng, taps5, p, bdmax, sizes)


@pytest.mark.gpu
@pytest.mark.parametrize("bdmax", [255, 1023, 4095])
@pytest.mark.parametrize("kind,sgr_idx", [(2, 14), (3, 10), (4, 0)])
def test_sgr(rb, ref, bdmax, kind, sgr_idx):
    rng = np.random.default_rng(bdmax * 5 + kind)
    s0, s1 = SGR_PARAMS[sgr_idx]
    for trial in range(4):
        p = rb.LrParams()
        p.sgr.s0, p.sgr.s1 = s0, s1
        p.sgr.w0 = int(rng.integers(0, 128)) - 96 if s0 else 0
        p.sgr.w1 = (160 - int(rng.integers(0, 128)) if s1 else 33) - p.sgr.w0
        sizes = (1 + int(rng.integers(0, 384)), 1 + int(rng.integers(0, 64)))
        if trial == 0:
            sizes = (384, 64)
        if trial == 1:
            sizes = (2, 1)
        _run(rb, ref, rng, kind, p, bdmax, sizes)
