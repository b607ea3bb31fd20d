"""Slot-level checks of the four post-filter / film-grain DSP tables: every member that
rb200_{loop_filter,cdef,loop_restoration,film_grain}_dsp_init fills is called THROUGH the table, with the reference's own
function-pointer signature (src/loopfilter.rs:20-34, src/cdef.rs:35-56, src/looprestoration.rs:91-107,
src/filmgrain.rs:41-201), and compared with the oracle's same slot.  (itx, mc and ipred have theirs in test_itx.py,
test_mc.py, test_ipred.py.)"""
import ctypes as C

import numpy as np
import pytest

from refharness import ptr

VP, SS, I, U32, SZ = C.c_void_p, C.c_ssize_t, C.c_int, C.c_uint32, C.c_size_t
LPF_FN = C.CFUNCTYPE(None, VP, SS, VP, VP, SS, VP, I, I)
CDEF_FN = C.CFUNCTYPE(None, VP, SS, VP, VP, VP, I, I, I, I, U32, I)
CDEF_DIR_FN = C.CFUNCTYPE(I, VP, SS, C.POINTER(C.c_uint), I)
LR_FN = C.CFUNCTYPE(None, VP, SS, VP, VP, I, I, VP, U32, I)
GEN_Y_FN = C.CFUNCTYPE(None, VP, VP, I)
GEN_UV_FN = C.CFUNCTYPE(None, VP, VP, VP, SS, I)
FGY_FN = C.CFUNCTYPE(None, VP, VP, SS, VP, SZ, VP, VP, I, I, I)
FGUV_FN = C.CFUNCTYPE(None, VP, VP, SS, VP, SZ, VP, VP, I, I, VP, SS, I, I, I)


def _table(init, n, bpc):
    tbl = (C.c_void_p * n)()
    init(tbl, bpc)
    assert all(tbl[i] for i in range(n)), "unfilled slot"
    return tbl


@pytest.mark.gpu
@pytest.mark.parametrize("bdmax", [255, 1023])
def test_loop_filter_table_slots(rb, ref, bdmax):
    """loop_filter_sb[uv][dir]: a 32-unit (16 for chroma) edge strip through each of the four slots."""
    from rav1d_b200.synth.framegen import calc_eih
    tbl = _table(rb.loop_filter_dsp_init, 4, 8 if bdmax == 255 else 10)
    rng = np.random.default_rng(bdmax)
    pdt = np.uint16 if bdmax > 255 else np.uint8
    lut = calc_eih(2)
    for uv in range(2):
        for dir_ in range(2):
            n = 16 if uv else 32
            w, h = (n * 4, 16) if dir_ else (16, n * 4)
            # smooth content so that the filters trigger
            img = (np.add.outer(np.arange(h), np.arange(w)) // 3 + rng.integers(0, 3, size=(h, w)) + (bdmax >> 1)).astype(pdt)
            vmask = np.zeros(4, np.uint32)
            for j in range(n):
                vmask[int(rng.integers(0, 2 if uv else 3))] |= np.uint32(1 << j)
            lv = np.full((64, 4), 24, np.uint8)
            a, b = img.copy(), img.copy()
            isz = img.itemsize
            off = (8 * w if dir_ else 8) * isz
            lbase = lv.ctypes.data + (32 if dir_ else 1) * 4
            b4 = 32 if dir_ else 2
            ref.ref_lpf_sb(uv, dir_, C.c_void_p(a.ctypes.data + off), w * isz, ptr(vmask), C.c_void_p(lbase), b4, C.addressof(lut), n, bdmax)
            LPF_FN(tbl[uv * 2 + dir_])(b.ctypes.data + off, w * isz, vmask.ctypes.data, lbase, b4, C.addressof(lut), n, bdmax)
            assert np.array_equal(a, b), (bdmax, uv, dir_)
            assert not np.array_equal(a, img)


@pytest.mark.gpu
@pytest.mark.parametrize("bdmax", [255, 1023])
def test_cdef_table_slots(rb, ref, bdmax):
    """dir and fb[0..2] through the table."""
    tbl = _table(rb.cdef_dsp_init, 4, 8 if bdmax == 255 else 10)
    rng = np.random.default_rng(bdmax + 7)
    pdt = np.uint16 if bdmax > 255 else np.uint8
    isz = np.dtype(pdt).itemsize
    bdmin8 = (bdmax + 1).bit_length() - 1 - 8
    src = rng.integers(0, bdmax + 1, size=(8, 8)).astype(pdt)
    v0, v1 = C.c_uint(0), C.c_uint(0)
    d0 = ref.ref_cdef_dir(ptr(src), src.strides[0], C.byref(v0), bdmax)
    d1 = CDEF_DIR_FN(tbl[0])(src.ctypes.data, src.strides[0], C.byref(v1), bdmax)
    assert (d0, v0.value) == (d1, v1.value)
    stride = 16 * isz
    for idx in range(3):
        for edges in (0, 5, 10, 15):
            buf = rng.integers(0, bdmax + 1, size=16 * 10 + 16).astype(pdt)
            top = rng.integers(0, bdmax + 1, size=16 * 2 + 16).astype(pdt)
            bot = rng.integers(0, bdmax + 1, size=16 * 2 + 16).astype(pdt)
            left = rng.integers(0, bdmax + 1, size=16).astype(pdt)
            a, b = buf.copy(), buf.copy()
            args = (stride, left.ctypes.data, top.ctypes.data + 8 * isz, bot.ctypes.data + 8 * isz, 5 << bdmin8, 2 << bdmin8,
                    int(rng.integers(0, 8)), 5 + bdmin8, edges, bdmax)
            ref.ref_cdef_fb(idx, C.c_void_p(a.ctypes.data + 8 * isz), args[0], C.c_void_p(args[1]), C.c_void_p(args[2]), C.c_void_p(args[3]), *args[4:])
            CDEF_FN(tbl[1 + idx])(b.ctypes.data + 8 * isz, *args)
            assert np.array_equal(a, b), (bdmax, idx, edges)


@pytest.mark.gpu
@pytest.mark.parametrize("bdmax", [255, 1023])
def test_loop_restoration_table_slots(rb, ref, bdmax):
    """wiener[0..1] and sgr[0..2] through the table (kind = slot order of the reference's table)."""
    tbl = _table(rb.loop_restoration_dsp_init, 5, 8 if bdmax == 255 else 10)
    rng = np.random.default_rng(bdmax + 11)
    pdt = np.uint16 if bdmax > 255 else np.uint8
    isz = np.dtype(pdt).itemsize
    src = rng.integers(0, bdmax + 1, size=448 * 64 + 64).astype(pdt)
    edge = rng.integers(0, bdmax + 1, size=448 * 8 + 64).astype(pdt)
    left = rng.integers(0, bdmax + 1, size=(64, 4)).astype(pdt)
    for kind in range(5):
        p = rb.LrParams()
        if kind < 2:
            for d in range(2):
                f0, f1, f2 = (0 if kind else 3), -9, 21
                p.filter[d][0] = p.filter[d][6] = f0
                p.filter[d][1] = p.filter[d][5] = f1
                p.filter[d][2] = p.filter[d][4] = f2
                p.filter[d][3] = (128 if d else 0) - (f0 + f1 + f2) * 2
            if bdmax > 255:
                p.filter[0][3] += 128
        else:
            s0, s1 = {2: (56, 0), 3: (0, 2589), 4: (140, 3236)}[kind]
            p.sgr.s0, p.sgr.s1 = s0, s1
            p.sgr.w0 = -30 if s0 else 0
            p.sgr.w1 = (70 if s1 else 33) - p.sgr.w0
        for edges, w, h in ((15, 96, 40), (0, 33, 7), (6, 256, 64)):
            a, b = src.copy(), src.copy()
            ref.ref_lr(kind, C.c_void_p(a.ctypes.data + 64 * isz), 448 * isz, ptr(left), C.c_void_p(edge.ctypes.data + 64 * isz), w, h,
                       C.addressof(p), edges, bdmax)
            LR_FN(tbl[kind])(b.ctypes.data + 64 * isz, 448 * isz, left.ctypes.data, edge.ctypes.data + 64 * isz, w, h, C.addressof(p), edges, bdmax)
            assert np.array_equal(a, b), (bdmax, kind, edges, w, h)


@pytest.mark.gpu
@pytest.mark.parametrize("bdmax", [255, 1023])
def test_film_grain_table_slots(rb, ref, bdmax):
    """generate_grain_y, generate_grain_uv[3], fgy_32x32xn, fguv_32x32xn[3] through the table."""
    from rav1d_b200.synth import framegen
    tbl = _table(rb.film_grain_dsp_init, 8, 8 if bdmax == 255 else 10)
    rng = np.random.default_rng(bdmax + 13)
    pdt = np.uint16 if bdmax > 255 else np.uint8
    isz = np.dtype(pdt).itemsize
    edt = np.int16 if bdmax > 255 else np.int8
    GH, GW = 73, 82
    d = framegen.random_film_grain(rng, lag=2, overlap=1)
    a = np.zeros((GH + 1, GW), edt); b = a.copy()
    ref.ref_fg_gen_y(ptr(a), C.addressof(d), bdmax)
    GEN_Y_FN(tbl[0])(b.ctypes.data, C.addressof(d), bdmax)
    assert np.array_equal(a[:GH], b[:GH])
    luts_c = []
    for layout in (1, 2, 3):
        ca = np.zeros((GH + 1, GW), edt); cb = ca.copy()
        ref.ref_fg_gen_uv(layout - 1, ptr(ca), ptr(a), C.addressof(d), 1, bdmax)
        GEN_UV_FN(tbl[layout])(cb.ctypes.data, a.ctypes.data, C.addressof(d), 1, bdmax)
        assert np.array_equal(ca, cb), layout
        luts_c.append(ca)
    bitdepth = 8 if bdmax == 255 else 10
    scaling = np.zeros(4096 if bdmax > 255 else 256, np.uint8)
    parr = np.array([[d.y_points[i][0], d.y_points[i][1]] for i in range(14)], np.uint8)
    rb.check(rb.generate_scaling(bitdepth, ptr(parr), d.num_y_points, ptr(scaling)))
    src = rng.integers(0, bdmax + 1, size=(32, 128)).astype(pdt)
    x, y = np.zeros((32, 128), pdt), np.zeros((32, 128), pdt)
    ref.ref_fgy(ptr(x), ptr(src), 128 * isz, C.addressof(d), 100, ptr(scaling), ptr(a), 32, 3, bdmax)
    FGY_FN(tbl[4])(y.ctypes.data, src.ctypes.data, 128 * isz, C.addressof(d), 100, scaling.ctypes.data, a.ctypes.data, 32, 3, bdmax)
    assert np.array_equal(x, y)
    luma = rng.integers(0, bdmax + 1, size=(64, 256)).astype(pdt)
    for layout in (1, 2, 3):
        sx, sy = int(layout != 3), int(layout == 1)
        w, h = 100 >> sx, 32 >> sy
        x, y = np.zeros((32, 128), pdt), np.zeros((32, 128), pdt)
        ref.ref_fguv(layout - 1, ptr(x), ptr(src), 128 * isz, C.addressof(d), w, ptr(scaling), ptr(luts_c[layout - 1]), h, 3, ptr(luma),
                     256 * isz, 1, 0, bdmax)
        FGUV_FN(tbl[4 + layout])(y.ctypes.data, src.ctypes.data, 128 * isz, C.addressof(d), w, scaling.ctypes.data,
                                 luts_c[layout - 1].ctypes.data, h, 3, luma.ctypes.data, 256 * isz, 1, 0, bdmax)
        assert np.array_equal(x, y), layout
