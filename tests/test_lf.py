"""loop_filter_sb parity: product (CUDA, C ABI) vs the oracle (reference C DSP) on
inputs built like the reference's own differential test
(tests/checkasm/loopfilter.c:35-203: init_lpf_border / check_lpf_sb)."""
import ctypes as C

import numpy as np
import pytest

from refharness import ptr


def _lut(lib, sharp):
    from rav1d_b200.synth.framegen import calc_eih
    return calc_eih(sharp)


def _border(rng, col, E, I, bdmax):
    """init_lpf_border: col is a 16-sample view across the edge (index 8 = first q sample)."""
    bdmin8 = (bdmax + 1).bit_length() - 1 - 8
    F = 1 << bdmin8
    E <<= bdmin8
    I <<= bdmin8
    clip = lambda v: min(max(int(v), 0), bdmax)
    kind = int(rng.integers(0, 4))
    edge_diff = int(rng.integers(0, (E + 2) * 4)) - 2 * (E + 2)
    r = lambda: int(rng.integers(0, bdmax + 1))
    d = lambda i: 8 + i
    if kind == 0:
        for i in range(-8, 8):
            col[d(i)] = r()
        return
    n_flat = 7 if kind == 1 else 4
    if kind == 1:
        col[d(-8)] = r(); col[d(7)] = r()
    else:
        for i in range(4, 8):
            col[d(-(1 + i))] = r(); col[d(i)] = r()
    col[d(0)] = r()
    col[d(-1)] = clip(int(col[d(0)]) + edge_diff)
    for i in range(1, n_flat):
        if kind == 3:
            col[d(-(1 + i))] = clip(int(col[d(-i)]) + int(rng.integers(0, 2 * (I + 1))) - (I + 1))
            col[d(i)] = clip(int(col[d(i - 1)]) + int(rng.integers(0, 2 * (I + 1))) - (I + 1))
        else:
            col[d(-(1 + i))] = clip(int(col[d(-1)]) + int(rng.integers(0, 2 * (F + 1))) - (F + 1))
            col[d(i)] = clip(int(col[d(0)]) + int(rng.integers(0, 2 * (F + 1))) - (F + 1))


@pytest.mark.gpu
@pytest.mark.parametrize("bdmax", [255, 1023, 4095])
@pytest.mark.parametrize("uv,dir_", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_loop_filter_sb(rb, ref, bdmax, uv, dir_):
    rng = np.random.default_rng(bdmax * 4 + uv * 2 + dir_)
    pdt = np.uint16 if bdmax > 255 else np.uint8
    n_blks = 16 if uv else 32
    for trial in range(24):
        lut = _lut(rb, int(rng.integers(0, 8)))
        n_str = 2 if uv else 3
        i = trial % n_str
        vmask = np.zeros(4, np.uint32)
        lf_idx = int(rng.integers(0, 4))
        lv = np.zeros((64, 4), np.uint8)
        for j in range(n_blks):
            idx = int(rng.integers(0, i + 2))
            if idx:
                vmask[idx - 1] |= np.uint32(1 << j)
            if dir_:
                lv[j, lf_idx] = rng.integers(0, 64); lv[j + 32, lf_idx] = rng.integers(0, 64)
            else:
                lv[j * 2, lf_idx] = rng.integers(0, 64); lv[j * 2 + 1, lf_idx] = rng.integers(0, 64)
        if trial % 5 == 4:   # own level 0 -> neighbour fallback, and both 0 -> skipped
            lv[rng.random(64) < 0.4] = 0
        if dir_:
            w, h = n_blks * 4, 16
            img = np.zeros((h, w), pdt)
        else:
            w, h = 16, n_blks * 4
            img = np.zeros((h, w), pdt)
        for k in range(4 * n_blks):
            x = k >> 2
            if dir_:
                L = int(lv[32 + x, lf_idx]) or int(lv[x, lf_idx])
                col = img[:, k]
            else:
                L = int(lv[2 * x + 1, lf_idx]) or int(lv[2 * x, lf_idx])
                col = img[k, :]
            _border(rng, col, lut.e[L], lut.i[L], bdmax)
        a, b = img.copy(), img.copy()
        isz = img.itemsize
        off = (8 * w if dir_ else 8) * isz
        lbase = lv.ctypes.data + ((32 if dir_ else 1) * 4 + lf_idx)
        b4 = 32 if dir_ else 2
        ref.ref_lpf_sb(uv, dir_, C.c_void_p(a.ctypes.data + off), w * isz, ptr(vmask), C.c_void_p(lbase), b4,
                       C.addressof(lut), n_blks, bdmax)
        rb.check(rb.loop_filter_sb(uv, dir_, C.c_void_p(b.ctypes.data + off), w * isz, ptr(vmask), C.c_void_p(lbase), b4,
                                   C.byref(lut), n_blks, bdmax))
        assert np.array_equal(a, b), (bdmax, uv, dir_, trial, np.argwhere(a != b)[:8])
        assert not np.array_equal(a, img)   # the filter did something
