"""Motion-compensation parity: product (CUDA, C ABI) vs the oracle (reference C DSP), with
the input recipes of the reference's differential test (tests/checkasm/mc.c:58-790)."""
import ctypes as C

import numpy as np
import pytest

from refharness import ptr

BDS = [255, 1023, 4095]


def _h_next(h):
    if h in (4, 8, 16):
        return (h * 3) >> 1
    if h in (6, 12, 24):
        return (h & (h - 1)) * 2
    return h * 2


def _sizes():
    w = 2
    while w <= 128:
        h = 2 if w <= 32 else w // 4
        h_max = max(min(w * 4, 128), 32)
        while h <= h_max:
            yield w, h
            h = _h_next(h)
        w <<= 1


def _pdt(bdmax):
    return np.uint16 if bdmax > 255 else np.uint8


def _mct_input(rng, bdmax):
    """generate_mct_input: worst-case pattern in the top-left corner (tests/checkasm/mc.c:113-121)."""
    pat = np.array([-1, 0, -1, 0, 0, -1, 0, -1])
    sign = -int(rng.integers(0, 2))
    buf = rng.integers(0, bdmax + 1, size=(135, 135))
    yy, xx = np.mgrid[0:135, 0:135]
    corner = (xx | yy) < 8
    buf[corner] = ((pat[xx % 8] ^ pat[yy % 8] ^ sign) & bdmax)[corner]
    return buf.astype(_pdt(bdmax))


@pytest.mark.gpu
@pytest.mark.parametrize("bdmax", BDS)
def test_mc_put(rb, ref, bdmax):
    rng = np.random.default_rng(bdmax)
    pdt = _pdt(bdmax); isz = np.dtype(pdt).itemsize
    for filt in range(10):
        for w, h in _sizes():
            for mxy in range(4):
                if (filt not in (0, 5, 9)) and (w + h + mxy) % 3:   # thin out the middle filters
                    continue
                mx = int(rng.integers(1, 16)) if mxy & 1 else 0
                my = int(rng.integers(1, 16)) if mxy & 2 else 0
                src = rng.integers(0, bdmax + 1, size=(135, 135)).astype(pdt)
                a = np.zeros((h + 2, w + 16), pdt); b = a.copy()
                sp = C.c_void_p(src.ctypes.data + (135 * 3 + 3) * isz)
                ref.ref_mc(filt, C.c_void_p(a.ctypes.data + (a.shape[1] + 8) * isz), a.strides[0], sp, 135 * isz, w, h, mx, my, bdmax)
                rb.check(rb.mc(filt, C.c_void_p(b.ctypes.data + (b.shape[1] + 8) * isz), b.strides[0], sp, 135 * isz, w, h, mx, my, bdmax))
                assert np.array_equal(a, b), ("mc", filt, w, h, mx, my, bdmax)


@pytest.mark.gpu
@pytest.mark.parametrize("bdmax", BDS)
def test_mc_prep(rb, ref, bdmax):
    rng = np.random.default_rng(bdmax + 1)
    isz = np.dtype(_pdt(bdmax)).itemsize
    for filt in range(10):
        for w, h in _sizes():
            if w < 4:
                continue
            for mxy in range(4):
                if (filt not in (0, 5, 9)) and (w + h + mxy) % 3:
                    continue
                mx = int(rng.integers(1, 16)) if mxy & 1 else 0
                my = int(rng.integers(1, 16)) if mxy & 2 else 0
                src = _mct_input(rng, bdmax)
                a = np.full(w * h + 8, 12345, np.int16); b = a.copy()
                sp = C.c_void_p(src.ctypes.data + (135 * 3 + 3) * isz)
                ref.ref_mct(filt, ptr(a), sp, 135 * isz, w, h, mx, my, bdmax)
                rb.check(rb.mct(filt, ptr(b), sp, 135 * isz, w, h, mx, my, bdmax))
                assert np.array_equal(a, b), ("mct", filt, w, h, mx, my, bdmax)


@pytest.mark.gpu
@pytest.mark.parametrize("bdmax", BDS)
def test_mc_negative_strides(rb, ref, bdmax):
    rng = np.random.default_rng(bdmax + 2)
    pdt = _pdt(bdmax); isz = np.dtype(pdt).itemsize
    for w, h, mx, my in ((16, 16, 5, 9), (8, 4, 0, 3), (32, 8, 7, 0)):
        src = rng.integers(0, bdmax + 1, size=(135, 135)).astype(pdt)
        a = np.zeros((h, w), pdt); b = a.copy()
        sp = C.c_void_p(src.ctypes.data + (135 * (135 - 4) + 3) * isz)  # row 3 counted from the bottom
        for fn, d in ((ref.ref_mc, a), (rb.mc, b)):
            rc = fn(2, C.c_void_p(d.ctypes.data + (h - 1) * w * isz), -w * isz, sp, -135 * isz, w, h, mx, my, bdmax)
            assert not rc
        assert np.array_equal(a, b), ("neg stride", w, h)


@pytest.mark.gpu
@pytest.mark.parametrize("bdmax", BDS)
@pytest.mark.parametrize("prep", [0, 1])
def test_mc_scaled(rb, ref, bdmax, prep):
    rng = np.random.default_rng(bdmax + 3 + prep)
    pdt = _pdt(bdmax); isz = np.dtype(pdt).itemsize
    for filt in (0, 2, 4, 7, 9):
        for w, h in _sizes():
            if prep and w < 4:
                continue
            for p in range(3):
                if (w + h + p + filt) % 2:
                    continue
                mx, my = int(rng.integers(0, 1024)), int(rng.integers(0, 1024))
                dx = int(rng.integers(1, 2049))
                dy = int(rng.integers(1, 2049)) if p == 0 else p << 10
                src = rng.integers(0, bdmax + 1, size=(263 + 8, 263 + 8)).astype(pdt)
                sp = C.c_void_p(src.ctypes.data + (src.shape[1] * 3 + 3) * isz)
                ss = src.strides[0]
                if prep:
                    a = np.full(w * h + 8, 777, np.int16); b = a.copy()
                    ref.ref_mct_scaled(filt, ptr(a), sp, ss, w, h, mx, my, dx, dy, bdmax)
                    rb.check(rb.mct_scaled(filt, ptr(b), sp, ss, w, h, mx, my, dx, dy, bdmax))
                else:
                    a = np.zeros((h + 2, w + 16), pdt); b = a.copy()
                    ref.ref_mc_scaled(filt, C.c_void_p(a.ctypes.data + (a.shape[1] + 8) * isz), a.strides[0], sp, ss, w, h, mx, my, dx, dy, bdmax)
                    rb.check(rb.mc_scaled(filt, C.c_void_p(b.ctypes.data + (b.shape[1] + 8) * isz), b.strides[0], sp, ss, w, h, mx, my, dx, dy, bdmax))
                assert np.array_equal(a, b), ("scaled", prep, filt, w, h, mx, my, dx, dy, bdmax)


def _compound_inputs(ref, rng, bdmax):
    """init_tmp (tests/checkasm/mc.c:275-285): two worst-case-seeded 128x128 preps from the reference."""
    isz = np.dtype(_pdt(bdmax)).itemsize
    out = []
    for _ in range(2):
        src = _mct_input(rng, bdmax)
        t = np.zeros(128 * 128, np.int16)
        ref.ref_mct(5, ptr(t), C.c_void_p(src.ctypes.data + (135 * 3 + 3) * isz), 135 * isz, 128, 128, 8, 8, bdmax)
        out.append(t)
    return out


def _cmp_sizes():
    w = 4
    while w <= 128:
        h = max(w // 4, 4)
        while h <= min(w * 4, 128):
            yield w, h
            h <<= 1
        w <<= 1


@pytest.mark.gpu
@pytest.mark.parametrize("bdmax", BDS)
def test_compound(rb, ref, bdmax):
    rng = np.random.default_rng(bdmax + 4)
    pdt = _pdt(bdmax); isz = np.dtype(pdt).itemsize
    for w, h in _cmp_sizes():
        t1, t2 = _compound_inputs(ref, rng, bdmax)
        def dst():
            d = np.zeros((h + 2, w + 16), pdt)
            return d, C.c_void_p(d.ctypes.data + (d.shape[1] + 8) * isz), d.strides[0]
        a, ap, st = dst(); b, bp, _ = dst()
        ref.ref_avg(ap, st, ptr(t1), ptr(t2), w, h, bdmax)
        rb.check(rb.avg(bp, st, ptr(t1), ptr(t2), w, h, bdmax))
        assert np.array_equal(a, b), ("avg", w, h, bdmax)
        weight = int(rng.integers(1, 16))
        a, ap, st = dst(); b, bp, _ = dst()
        ref.ref_w_avg(ap, st, ptr(t1), ptr(t2), w, h, weight, bdmax)
        rb.check(rb.w_avg(bp, st, ptr(t1), ptr(t2), w, h, weight, bdmax))
        assert np.array_equal(a, b), ("w_avg", w, h, weight, bdmax)
        m = rng.integers(0, 65, size=w * h).astype(np.uint8)
        a, ap, st = dst(); b, bp, _ = dst()
        ref.ref_mask(ap, st, ptr(t1), ptr(t2), w, h, ptr(m), bdmax)
        rb.check(rb.mask(bp, st, ptr(t1), ptr(t2), w, h, ptr(m), bdmax))
        assert np.array_equal(a, b), ("mask", w, h, bdmax)
        for ss in range(3):
            sign = int(rng.integers(0, 2))
            m0 = np.full(128 * 128, 99, np.uint8); m1 = m0.copy()
            a, ap, st = dst(); b, bp, _ = dst()
            ref.ref_w_mask(ss, ap, st, ptr(t1), ptr(t2), w, h, ptr(m0), sign, bdmax)
            rb.check(rb.w_mask(ss, bp, st, ptr(t1), ptr(t2), w, h, ptr(m1), sign, bdmax))
            assert np.array_equal(a, b), ("w_mask dst", ss, w, h, sign, bdmax)
            assert np.array_equal(m0, m1), ("w_mask mask", ss, w, h, sign, bdmax)


@pytest.mark.gpu
@pytest.mark.parametrize("bdmax", BDS)
def test_blend(rb, ref, bdmax):
    rng = np.random.default_rng(bdmax + 5)
    pdt = _pdt(bdmax); isz = np.dtype(pdt).itemsize
    cases = []
    w = 4
    while w <= 32:                                   # blend: tests/checkasm/mc.c:445-484
        h = max(w // 2, 4)
        while h <= min(w * 2, 32):
            cases.append((0, w, h)); h <<= 1
        w <<= 1
    w = 2
    while w <= 32:                                   # blend_v: :486-523
        h = 2
        while h <= (64 if w == 2 else 128):
            cases.append((1, w, h)); h <<= 1
        w <<= 1
    w = 2
    while w <= 128:                                  # blend_h: :525-561
        h = 4 if w == 128 else 2
        while h <= 32:
            cases.append((2, w, h)); h <<= 1
        w <<= 1
    for dir_, w, h in cases:
        tmp = rng.integers(0, bdmax + 1, size=w * h).astype(pdt)
        m = rng.integers(0, 65, size=w * h).astype(np.uint8)
        a = np.zeros((h + 2, w + 16), pdt)
        a[1:h + 1, 8:8 + w] = rng.integers(0, bdmax + 1, size=(h, w))
        b = a.copy()
        off = (a.shape[1] + 8) * isz
        ref.ref_blend(dir_, C.c_void_p(a.ctypes.data + off), a.strides[0], ptr(tmp), w, h, ptr(m), bdmax)
        rb.check(rb.blend(dir_, C.c_void_p(b.ctypes.data + off), b.strides[0], ptr(tmp), w, h, ptr(m), bdmax))
        assert np.array_equal(a, b), ("blend", dir_, w, h, bdmax)


@pytest.mark.gpu
@pytest.mark.parametrize("bdmax", BDS)
def test_warp8x8(rb, ref, bdmax):
    rng = np.random.default_rng(bdmax + 6)
    pdt = _pdt(bdmax); isz = np.dtype(pdt).itemsize
    for trial in range(48):
        mx, my = (int(v) - 0xa00 for v in rng.integers(0, 0x2000, size=2))
        abcd = (rng.integers(0, 0x2000, size=4) - 0xa00).astype(np.int16)
        src = rng.integers(0, bdmax + 1, size=(15, 15)).astype(pdt)
        sp = C.c_void_p(src.ctypes.data + (15 * 3 + 3) * isz)
        a = np.zeros((10, 24), pdt); b = a.copy()
        ref.ref_warp8x8(C.c_void_p(a.ctypes.data + (24 + 8) * isz), a.strides[0], sp, 15 * isz, ptr(abcd), mx, my, bdmax)
        rb.check(rb.warp8x8(C.c_void_p(b.ctypes.data + (24 + 8) * isz), b.strides[0], sp, 15 * isz, ptr(abcd), mx, my, bdmax))
        assert np.array_equal(a, b), ("warp8x8", trial, bdmax)
        ta = np.full(8 * 16 + 8, 321, np.int16); tb = ta.copy()
        ref.ref_warp8x8t(ptr(ta), 16, sp, 15 * isz, ptr(abcd), mx, my, bdmax)
        rb.check(rb.warp8x8t(ptr(tb), 16, sp, 15 * isz, ptr(abcd), mx, my, bdmax))
        assert np.array_equal(ta, tb), ("warp8x8t", trial, bdmax)


@pytest.mark.gpu
@pytest.mark.parametrize("bdmax", [255, 1023])
def test_emu_edge(rb, ref, bdmax):
    rng = np.random.default_rng(bdmax + 7)
    pdt = _pdt(bdmax); isz = np.dtype(pdt).itemsize
    src = rng.integers(0, bdmax + 1, size=(160, 160)).astype(pdt)

    def off(edge, lo_flag, hi_flag, b):
        e = edge & (lo_flag | hi_flag)
        i = 160 if e else 1 + int(rng.integers(0, b - 2))
        if e == (lo_flag | hi_flag):
            pos = int(rng.integers(0, i - b + 1))
        elif e == lo_flag:
            pos = (i - b) + 1 + int(rng.integers(0, b - 1))
        elif e == hi_flag:
            pos = -(1 + int(rng.integers(0, b - 1)))
        else:
            pos = -(1 + int(rng.integers(0, b - i - 1)))
        return pos, i

    for w, h in _cmp_sizes():
        for edge in range(15):
            bw, bh = w + int(rng.integers(0, 8)), h + int(rng.integers(0, 8))
            x, iw = off(edge, 4, 8, bw)
            y, ih = off(edge, 1, 2, bh)
            a = np.zeros((135, 192), pdt); b = a.copy()
            ref.ref_emu_edge(bw, bh, iw, ih, x, y, ptr(a), 192 * isz, ptr(src), 160 * isz, bdmax)
            rb.check(rb.emu_edge(bw, bh, iw, ih, x, y, ptr(b), 192 * isz, ptr(src), 160 * isz, bdmax))
            assert np.array_equal(a, b), ("emu_edge", bw, bh, iw, ih, x, y)


@pytest.mark.gpu
@pytest.mark.parametrize("bdmax", BDS)
def test_resize(rb, ref, bdmax):
    rng = np.random.default_rng(bdmax + 8)
    pdt = _pdt(bdmax); isz = np.dtype(pdt).itemsize
    for trial in range(8):
        src = rng.integers(0, bdmax + 1, size=(64, 512)).astype(pdt)
        w_den = 9 + int(rng.integers(0, 8))
        src_w = 16 + int(rng.integers(0, 512 - 16 + 1))
        dst_w = w_den * src_w >> 3
        dx = ((src_w << 14) + (dst_w >> 1)) // dst_w
        err = dst_w * dx - (src_w << 14)
        num = -((dst_w - src_w) << 13) + (dst_w >> 1)
        q = abs(num) // dst_w * (1 if num >= 0 else -1)      # C division truncates toward zero
        mx0 = (q + 128 - (err >> 1)) & 0x3fff
        a = np.zeros((64, 1024 + 32), pdt); b = a.copy()
        ref.ref_resize(C.c_void_p(a.ctypes.data + 16 * isz), a.strides[0], ptr(src), 512 * isz, dst_w, 64, src_w, dx, mx0, bdmax)
        rb.check(rb.resize(C.c_void_p(b.ctypes.data + 16 * isz), b.strides[0], ptr(src), 512 * isz, dst_w, 64, src_w, dx, mx0, bdmax))
        assert np.array_equal(a, b), ("resize", src_w, dst_w, dx, mx0, bdmax)


@pytest.mark.gpu
def test_mc_dsp_table(rb, ref):
    """rb200_mc_dsp_init fills every slot of the Rav1dMCDSPContext mirror."""
    n_ptrs = 10 * 4 + 3 + 3 + 3 + 2 + 2
    tbl = (C.c_void_p * n_ptrs)()
    rb.mc_dsp_init(tbl, 8)
    assert all(tbl[i] for i in range(n_ptrs))
    fn = C.CFUNCTYPE(None, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int)(tbl[0])
    rng = np.random.default_rng(0)
    src = rng.integers(0, 256, size=(32, 32)).astype(np.uint8)
    a = np.zeros((8, 8), np.uint8); b = a.copy()
    sp = C.c_void_p(src.ctypes.data + 32 * 8 + 8)
    ref.ref_mc(0, ptr(a), 8, sp, 32, 8, 8, 3, 11, 255)
    fn(b.ctypes.data, 8, sp, 32, 8, 8, 3, 11, 255)
    assert np.array_equal(a, b)


@pytest.mark.gpu
def test_wedge_masks(rb, ref):
    """All of dav1d_wedge_masks: 9 block sizes x 3 layouts x 2 signs x 16 wedges, against the table the
    reference builds at start-up (src/wedge.c:216-243)."""
    import ctypes as C
    for w in (8, 16, 32):
        for h in (8, 16, 32):
            for ss in range(3):
                cw, ch = w >> (ss > 0), h >> (ss == 2)
                for sign in range(2):
                    for idx in range(16):
                        exp = np.ctypeslib.as_array(C.cast(ref.ref_wedge_mask(w, h, ss, sign, idx), C.POINTER(C.c_uint8)), shape=(ch, cw))
                        got = np.zeros((ch, cw), np.uint8)
                        rb.check(rb.wedge_mask(w, h, ss, sign, idx, got.ctypes.data))
                        assert np.array_equal(exp, got), (w, h, ss, sign, idx)
