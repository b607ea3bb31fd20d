"""Intra-prediction parity: product (CUDA, C ABI) vs the oracle (reference C DSP) with the input recipes of the
reference's differential test (tests/checkasm/ipred.c:68-290): every mode, every block shape, random edges, random
directional angles with the smooth / edge-filter flags, random max_width / max_height for Z2, all five filter-intra
sets, every cfl_ac padding, the four cfl_pred variants and palette prediction."""
import ctypes as C

import numpy as np
import pytest

from refharness import ptr

BDS = [255, 1023, 4095]
Z_ANGLES = [3, 6, 9, 14, 17, 20, 23, 26, 29, 32, 36, 39, 42, 45, 48, 51, 54, 58, 61, 64, 67, 70, 73, 76, 81, 84, 87]
MODE_NAMES = ["dc", "v", "h", "dc_left", "dc_top", "dc_128", "z1", "z2", "z3", "smooth", "smooth_v", "smooth_h", "paeth", "filter"]


def _pdt(bdmax):
    return np.uint16 if bdmax > 255 else np.uint8


def _shapes(wmax, hmax=None):
    hmax = hmax or wmax
    w = 4
    while w <= wmax:
        h = max(w // 4, 4)
        while h <= min(w * 4, hmax):
            yield w, h
            h <<= 1
        w <<= 1


def test_ipred_mode_numbering_matches_the_reference(ref):
    from rav1d_b200 import lib
    ids = (C.c_int * 14)()
    assert ref.ref_ipred_mode_ids(ids) == 14
    assert list(ids) == [lib.DC_PRED, lib.VERT_PRED, lib.HOR_PRED, lib.PAETH_PRED, lib.SMOOTH_PRED, lib.SMOOTH_V_PRED, lib.SMOOTH_H_PRED,
                         lib.Z1_PRED, lib.Z2_PRED, lib.Z3_PRED, lib.LEFT_DC_PRED, lib.TOP_DC_PRED, lib.DC_128_PRED, lib.FILTER_PRED]


@pytest.mark.gpu
@pytest.mark.parametrize("bdmax", BDS)
@pytest.mark.parametrize("mode", range(14))
def test_intra_pred(rb, ref, mode, bdmax):
    rng = np.random.default_rng(100 * mode + bdmax)
    pdt = _pdt(bdmax)
    isz = np.dtype(pdt).itemsize
    n = 0
    for w, h in _shapes(32 if mode == rb.FILTER_PRED else 64):
        iters = 6 if rb.Z1_PRED <= mode <= rb.Z3_PRED else (5 if mode == rb.FILTER_PRED else 1)
        for it in range(iters):
            a = maxw = maxh = 0
            if rb.Z1_PRED <= mode <= rb.Z3_PRED:
                a = (90 * (mode - rb.Z1_PRED) + Z_ANGLES[int(rng.integers(0, 27))]) | (int(rng.integers(0, 4)) << 9)
                if mode == rb.Z2_PRED:
                    mw, mh = int(rng.integers(0, 8192)), int(rng.integers(0, 8192))
                    maxw = 1 + (mw & (4095 if mw & 4096 else w - 1))
                    maxh = 1 + (mh & (4095 if mh & 4096 else h - 1))
            elif mode == rb.FILTER_PRED:
                a = it | (int(rng.integers(0, 4)) << 9)
            edge = rng.integers(0, bdmax + 1, size=257).astype(pdt)
            if it == 1:
                edge[:] = rng.choice(np.array([0, bdmax], pdt), size=257)      # saturating edges
            tl = edge.ctypes.data + 128 * isz
            d0 = rng.integers(0, bdmax + 1, size=(h + 2, 80)).astype(pdt)
            d1 = d0.copy()
            off = (80 + 8) * isz
            ref.ref_ipred(mode, C.c_void_p(d0.ctypes.data + off), 80 * isz, C.c_void_p(tl), w, h, a, maxw, maxh, bdmax)
            rb.check(rb.ipred(mode, C.c_void_p(d1.ctypes.data + off), 80 * isz, C.c_void_p(tl), w, h, a, maxw, maxh, bdmax))
            assert np.array_equal(d0, d1), (MODE_NAMES[mode], w, h, a & 511, a >> 9, maxw, maxh, np.argwhere(d0 != d1)[:4])
            n += 1
    assert n >= 10


@pytest.mark.gpu
@pytest.mark.parametrize("bdmax", BDS)
def test_cfl_ac(rb, ref, bdmax):
    rng = np.random.default_rng(bdmax)
    pdt = _pdt(bdmax)
    isz = np.dtype(pdt).itemsize
    n = 0
    for ss in range(3):                        # layout - 1: 4:2:0, 4:2:2, 4:4:4
        ss_hor, ss_ver = int(ss != 2), int(ss == 0)
        h_step, v_step = 2 >> ss_hor, 2 >> ss_ver
        for w, h in _shapes(32 >> ss_hor, 32 >> ss_ver):
            for w_pad in range(max((w >> 2) - h_step, 0), -1, -h_step):
                for h_pad in range(max((h >> 2) - v_step, 0), -1, -v_step):
                    luma = rng.integers(0, bdmax + 1, size=(32, 32)).astype(pdt)
                    a0 = np.full(w * h, 77, np.int16); a1 = a0.copy()
                    ref.ref_cfl_ac(ss, ptr(a0), ptr(luma), 32 * isz, w_pad, h_pad, w, h, bdmax)
                    rb.check(rb.cfl_ac(ss, ptr(a1), ptr(luma), 32 * isz, w_pad, h_pad, w, h, bdmax))
                    assert np.array_equal(a0, a1), (ss, w, h, w_pad, h_pad)
                    n += 1
    assert n > 40


@pytest.mark.gpu
@pytest.mark.parametrize("bdmax", BDS)
def test_cfl_pred(rb, ref, bdmax):
    rng = np.random.default_rng(bdmax + 1)
    pdt = _pdt(bdmax)
    isz = np.dtype(pdt).itemsize
    for mode in (rb.DC_PRED, rb.LEFT_DC_PRED, rb.TOP_DC_PRED, rb.DC_128_PRED):
        for w, h in _shapes(32):
            alpha = (int(rng.integers(0, 16)) + 1) * (1 - 2 * int(rng.integers(0, 2)))
            edge = rng.integers(0, bdmax + 1, size=257).astype(pdt)
            tl = edge.ctypes.data + 128 * isz
            ac = rng.integers(0, (bdmax << 3) + 1, size=w * h).astype(np.int64)
            ac = (ac - (ac.sum() + (w * h >> 1)) // (w * h)).astype(np.int16)
            d0 = rng.integers(0, bdmax + 1, size=(h + 2, 48)).astype(pdt); d1 = d0.copy()
            off = (48 + 8) * isz
            ref.ref_cfl_pred(mode, C.c_void_p(d0.ctypes.data + off), 48 * isz, C.c_void_p(tl), w, h, ptr(ac), alpha, bdmax)
            rb.check(rb.cfl_pred(mode, C.c_void_p(d1.ctypes.data + off), 48 * isz, C.c_void_p(tl), w, h, ptr(ac), alpha, bdmax))
            assert np.array_equal(d0, d1), (mode, w, h, alpha)


@pytest.mark.gpu
@pytest.mark.parametrize("bdmax", BDS)
def test_pal_pred(rb, ref, bdmax):
    rng = np.random.default_rng(bdmax + 2)
    pdt = _pdt(bdmax)
    isz = np.dtype(pdt).itemsize
    for w, h in _shapes(64):
        pal = rng.integers(0, bdmax + 1, size=8).astype(pdt)
        idx = rng.integers(0, 8, size=w * h).astype(np.uint8)
        d0 = rng.integers(0, bdmax + 1, size=(h + 2, 80)).astype(pdt); d1 = d0.copy()
        off = (80 + 8) * isz
        ref.ref_pal_pred(C.c_void_p(d0.ctypes.data + off), 80 * isz, ptr(pal), ptr(idx), w, h, bdmax)
        rb.check(rb.pal_pred(C.c_void_p(d1.ctypes.data + off), 80 * isz, ptr(pal), ptr(idx), w, h, bdmax))
        assert np.array_equal(d0, d1), (w, h)


@pytest.mark.gpu
def test_intra_pred_dsp_table_slots(rb, ref):
    """rb200_intra_pred_dsp_init fills the members of Rav1dIntraPredDSPContext (src/ipred.rs:164-169) with callable slots."""
    ctx = rb.IntraPredDSPContext()
    rb.intra_pred_dsp_init(C.byref(ctx), 10)
    assert all(ctx.intra_pred[m] for m in range(14)) and all(ctx.cfl_ac[i] for i in range(3)) and ctx.pal_pred
    assert [bool(ctx.cfl_pred[m]) for m in range(6)] == [True, False, False, True, True, True]
    fn = C.CFUNCTYPE(None, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int)(ctx.intra_pred[rb.PAETH_PRED])
    rng = np.random.default_rng(3)
    edge = rng.integers(0, 1024, size=257).astype(np.uint16)
    d0 = np.zeros((16, 16), np.uint16); d1 = d0.copy()
    tl = edge.ctypes.data + 256
    ref.ref_ipred(rb.PAETH_PRED, ptr(d0), 32, C.c_void_p(tl), 16, 16, 0, 0, 0, 1023)
    fn(d1.ctypes.data, 32, tl, 16, 16, 0, 0, 0, 1023)
    assert np.array_equal(d0, d1)
