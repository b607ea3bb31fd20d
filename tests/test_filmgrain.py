"""Film-grain parity: grain LUT generation, scaling LUTs, fgy / fguv row application (per
call, recipes of tests/checkasm/filmgrain.c:49-401) and the whole-picture stage against the
reference's dav1d_apply_grain (src/fg_apply_tmpl.c:229-245)."""
import ctypes as C

import numpy as np
import pytest

import framecheck
import refharness
from refharness import ptr

BDS = [255, 1023, 4095]
GW, GH = 82, 73


def _entry(bdmax):
    return np.int16 if bdmax > 255 else np.int8


def _rand_fg(rb, rng, **kw):
    from rav1d_b200.synth import framegen
    return framegen.random_film_grain(rng, **kw)


def test_film_grain_struct_layout(ref):
    from rav1d_b200 import lib
    assert ref.ref_sizeof_film_grain_data() == C.sizeof(lib.FilmGrainData)


@pytest.mark.gpu
@pytest.mark.parametrize("bdmax", BDS)
def test_generate_grain(rb, ref, bdmax):
    rng = np.random.default_rng(bdmax)
    for lag in range(4):
        for rep in range(3):
            d = _rand_fg(rb, rng, lag=lag, luma_points=bool(rep & 1))
            a = np.full((GH + 1, GW), -1, _entry(bdmax)); b = a.copy()
            ref.ref_fg_gen_y(ptr(a), C.addressof(d), bdmax)
            rb.check(rb.generate_grain_y(ptr(b), C.byref(d), bdmax))
            assert np.array_equal(a[:GH], b[:GH]), ("grain_y", lag, bdmax)
            for layout in (1, 2, 3):
                uv = int(rng.integers(0, 2))
                ca = np.full((GH + 1, GW), -1, _entry(bdmax)); cb = ca.copy()
                ref.ref_fg_gen_uv(layout - 1, ptr(ca), ptr(a), C.addressof(d), uv, bdmax)
                rb.check(rb.generate_grain_uv(layout, ptr(cb), ptr(a), C.byref(d), uv, bdmax))
                assert np.array_equal(ca, cb), ("grain_uv", lag, layout, uv, bdmax)


def _ref_scaling(ref, rb, d, which, bdmax):
    """Scaling LUT via the reference: it has no exported generate_scaling, so use the product's and
    cross-check it against a numpy restatement of src/fg_apply_tmpl.c:41-96."""
    bitdepth = (bdmax + 1).bit_length() - 1
    pts, num = (d.y_points, d.num_y_points) if which == 0 else (d.uv_points[which - 1], d.num_uv_points[which - 1])
    size = 4096 if bdmax > 255 else 256
    got = np.zeros(size, np.uint8)
    parr = np.array([[pts[i][0], pts[i][1]] for i in range(14 if which == 0 else 10)], np.uint8)
    rb.check(rb.generate_scaling(bitdepth, ptr(parr), num, ptr(got)))
    exp = np.zeros(size, np.int64)
    shift = bitdepth - 8
    n_used = 1 << bitdepth
    if num:
        exp[:int(parr[0, 0]) << shift] = parr[0, 1]
        for i in range(num - 1):
            bx, by, ex, ey = (int(v) for v in (*parr[i], *parr[i + 1]))
            dx, dy = ex - bx, ey - by
            delta = dy * ((0x10000 + (dx >> 1)) // dx)
            dd = 0x8000
            for x in range(dx):
                exp[(bx + x) << shift] = by + (dd >> 16)
                dd += delta
        n = int(parr[num - 1, 0]) << shift
        exp[n:n_used] = parr[num - 1, 1]
        if shift:
            pad, rnd = 1 << shift, (1 << shift) >> 1
            for i in range(num - 1):
                bx, ex = int(parr[i, 0]) << shift, int(parr[i + 1, 0]) << shift
                for x in range(0, ex - bx, pad):
                    rng_ = (exp[bx + x + pad] & 0xff) - (exp[bx + x] & 0xff)
                    r = rnd
                    for k in range(1, pad):
                        r += rng_
                        exp[bx + x + k] = (exp[bx + x] & 0xff) + (r >> shift)
    assert np.array_equal(got[:n_used], (exp[:n_used] & 0xff).astype(np.uint8)), ("generate_scaling", which, bdmax)
    return got


@pytest.mark.gpu
@pytest.mark.parametrize("bdmax", BDS)
def test_fgy_rows(rb, ref, bdmax):
    rng = np.random.default_rng(bdmax + 1)
    pdt = np.uint16 if bdmax > 255 else np.uint8
    isz = np.dtype(pdt).itemsize
    for trial in range(6):
        d = _rand_fg(rb, rng)
        lut = np.zeros((GH + 1, GW), _entry(bdmax))
        ref.ref_fg_gen_y(ptr(lut), C.addressof(d), bdmax)
        scaling = _ref_scaling(ref, rb, d, 0, bdmax)
        for overlap in (0, 1):
            d.overlap_flag = overlap
            for i in range(1 + 2 * overlap):
                if overlap:
                    w = 35 + int(rng.integers(0, 93))
                    if i == 0:
                        row_num, h = 0, 1 + int(rng.integers(0, 32))
                    else:
                        row_num = 1 + int(rng.integers(0, 0x800))
                        h = 3 + int(rng.integers(0, 30)) if i == 1 else 1 + int(rng.integers(0, 2))
                else:
                    w, h, row_num = 1 + int(rng.integers(0, 128)), 1 + int(rng.integers(0, 32)), int(rng.integers(0, 0x800))
                src = rng.integers(0, bdmax + 1, size=(32, 128)).astype(pdt)
                a = np.zeros((32, 128), pdt); b = a.copy()
                ref.ref_fgy(ptr(a), ptr(src), 128 * isz, C.addressof(d), w, ptr(scaling), ptr(lut), h, row_num, bdmax)
                rb.check(rb.fgy_32x32xn(ptr(b), ptr(src), 128 * isz, C.byref(d), w, ptr(scaling), ptr(lut), h, row_num, bdmax))
                assert np.array_equal(a, b), ("fgy", trial, overlap, i, w, h, row_num, bdmax)


@pytest.mark.gpu
@pytest.mark.parametrize("bdmax", BDS)
@pytest.mark.parametrize("layout", [1, 2, 3])
def test_fguv_rows(rb, ref, bdmax, layout):
    rng = np.random.default_rng(bdmax + 10 * layout)
    pdt = np.uint16 if bdmax > 255 else np.uint8
    isz = np.dtype(pdt).itemsize
    ss_x, ss_y = int(layout != 3), int(layout == 1)
    for csfl in (0, 1):
        for trial in range(3):
            d = _rand_fg(rb, rng, csfl=csfl, luma_points=bool(csfl))
            uv_pl, is_id = int(rng.integers(0, 2)), int(rng.integers(0, 2))
            luty = np.zeros((GH + 1, GW), _entry(bdmax)); lutc = luty.copy()
            ref.ref_fg_gen_y(ptr(luty), C.addressof(d), bdmax)
            ref.ref_fg_gen_uv(layout - 1, ptr(lutc), ptr(luty), C.addressof(d), uv_pl, bdmax)
            scaling = _ref_scaling(ref, rb, d, 0 if csfl else 1 + uv_pl, bdmax)
            for overlap in (0, 1):
                d.overlap_flag = overlap
                for i in range(1 + 2 * overlap):
                    if overlap:
                        w = (36 >> ss_x) + int(rng.integers(0, 92 >> ss_x))
                        if i == 0:
                            row_num, h = 0, 1 + int(rng.integers(0, 32 >> ss_y))
                        else:
                            row_num = 1 + int(rng.integers(0, 0x800))
                            h = ((2 if ss_y else 3) + int(rng.integers(0, 15 if ss_y else 30))) if i == 1 else (1 if ss_y else 1 + int(rng.integers(0, 2)))
                    else:
                        w, h = 1 + int(rng.integers(0, 128 >> ss_x)), 1 + int(rng.integers(0, 32 >> ss_y))
                        row_num = int(rng.integers(0, 0x800))
                    src = rng.integers(0, bdmax + 1, size=(32, 128)).astype(pdt)
                    luma = rng.integers(0, bdmax + 1, size=(32, 128)).astype(pdt)
                    a = np.zeros((32, 128), pdt); b = a.copy()
                    ref.ref_fguv(layout - 1, ptr(a), ptr(src), 128 * isz, C.addressof(d), w, ptr(scaling), ptr(lutc), h,
                                 row_num, ptr(luma), 128 * isz, uv_pl, is_id, bdmax)
                    rb.check(rb.fguv_32x32xn(layout, ptr(b), ptr(src), 128 * isz, C.byref(d), w, ptr(scaling), ptr(lutc), h,
                                             row_num, ptr(luma), 128 * isz, uv_pl, is_id, bdmax))
                    assert np.array_equal(a, b), ("fguv", layout, csfl, trial, overlap, i, w, h, row_num, bdmax)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,bpc", [(200, 120, 8), (333, 201, 10), (264, 200, 12), (1920, 1080, 10)])
def test_frame_film_grain_stage(rb, ref, w, h, bpc):
    """Whole-picture grain (prep + every 32-row band) on top of the filtered picture."""
    from rav1d_b200.synth import framegen
    rng = np.random.default_rng(w + bpc)
    s = framegen.generate(w, h, bpc, seed=w)
    start = framegen.recon_input_planes(s)
    variants = [dict(), dict(csfl=1), dict(luma_points=False, uv_points=(True, False)), dict(overlap=1, lag=3)]
    for kw in variants:
        d = _rand_fg(rb, rng, **kw)
        is_id = int(rng.integers(0, 2))
        cur = refharness.RefFrame(ref, s, 1)
        try:
            cur.load_filter_meta(); cur.set_planes(start); cur.filter(6)
            exp = framecheck.visible(s, cur.apply_grain(d, is_id))
            pre = framecheck.visible(s, cur.get_planes())
        finally:
            cur.close()
        dev = framegen.DeviceFrame(s)
        try:
            dev.load_batch(); dev.upload(0, start)
            rb.check(rb.frame_set_film_grain(dev.h, C.byref(d), is_id))
            dev.submit(6 | rb.STAGE_FILM_GRAIN); dev.wait()
            got = framecheck.visible(s, dev.readback())
        finally:
            dev.close()
        framecheck.assert_planes_equal(exp, got, f"film grain {w}x{h}@{bpc} {kw}")
        assert not np.array_equal(exp[0], pre[0]) or not d.num_y_points
