"""itxfm_add parity: product (CUDA, through the C ABI) vs the oracle
(reference C DSP, oracle/_ref) on checkasm-style inputs
(tests/checkasm/itx.c:242-310 in the reference)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import refharness
from refharness import ptr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KIND_NAMES = {0: "dct", 1: "adst", 2: "flipadst", 3: "identity", 4: "wht"}


def _hostsim():
    src = os.path.join(ROOT, "tests", "hostsim", "hostsim_itx1d.cpp")
    so = os.path.join(ROOT, "tests", "hostsim", "libhostsim_itx1d.so")
    if not os.path.exists(so) or os.path.getmtime(so) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(ROOT, "rav1d_b200/csrc/itx1d.cuh"))):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-o", so, src])
    return C.CDLL(so)


def test_itx1d_host_build_matches_reference_1d(ref):
    """The product's __host__ __device__ 1-D transforms (same code the kernels
    inline) against the reference's dav1d_inv_*_1d_c, incl. clip-saturating inputs."""
    hs = _hostsim()
    rng = np.random.default_rng(7)
    for kind in range(5):
        for n in (4, 8, 16, 32, 64):
            name = f"dav1d_inv_{KIND_NAMES[kind]}{n}_1d_c"
            if not hasattr(ref, name):
                continue
            f = getattr(ref, name)
            for bits in (16, 18, 20):
                lo, hi = -(1 << (bits - 1)), (1 << (bits - 1)) - 1
                for trial in range(400):
                    mag = [1 << (bits - 1), 1 << (bits - 4), 300][trial % 3]
                    x = rng.integers(-mag, mag, size=n).astype(np.int32)
                    if trial % 8 == 7:
                        x = rng.choice(np.array([lo, hi, 0, hi // 2], dtype=np.int32), size=n)
                    if n == 64:
                        x[32:] = 0
                    a, b = x.copy(), x.copy()
                    if kind == 4:
                        f(ptr(a), C.c_ssize_t(1))
                    else:
                        f(ptr(a), C.c_ssize_t(1), lo, hi)
                    assert hs.hostsim_itx1d(n, kind, ptr(b), lo, hi) == 0
                    assert np.array_equal(a, b), (name, bits, x)


def _one_case(rb, ref, rng, tx, txtp, bdmax, variant, neg_stride=False):
    from rav1d_b200.lib import TX_DIMS
    from rav1d_b200.synth.itxgen import gen_coefs
    w, h = TX_DIMS[tx]
    hbd = bdmax > 255
    pdt = np.uint16 if hbd else np.uint8
    cdt = np.int32 if hbd else np.int16
    coef, eob = gen_coefs(rng, tx, txtp, bdmax, 1, variant)
    if not hbd:
        coef = np.clip(coef, -32768, 32767)
    c0 = np.zeros(32 * 32, dtype=cdt)
    c0[:coef.shape[1]] = coef[0].astype(cdt)
    c0[coef.shape[1]:] = rng.integers(-100, 100, size=32 * 32 - coef.shape[1])  # guard: must stay untouched
    c1 = c0.copy()
    pitch = 80
    d0 = rng.integers(0, bdmax + 1, size=(h + 4, pitch)).astype(pdt)
    d1 = d0.copy()
    isz = d0.itemsize
    if neg_stride:
        off = ((h + 1) * pitch + 8) * isz
        stride = -pitch * isz
    else:
        off = (2 * pitch + 8) * isz
        stride = pitch * isz
    ref.ref_itxfm_add(tx, txtp, C.c_void_p(d0.ctypes.data + off), stride, ptr(c0), int(eob[0]), bdmax)
    rb.check(rb.itxfm_add(tx, txtp, C.c_void_p(d1.ctypes.data + off), stride, ptr(c1), int(eob[0]), bdmax))
    assert np.array_equal(d0, d1), (tx, txtp, bdmax, variant)
    assert np.array_equal(c0, c1), ("coef zeroing", tx, txtp, bdmax, variant)


@pytest.mark.gpu
@pytest.mark.parametrize("bdmax", [255, 1023, 4095])
def test_itxfm_add_all_types_sizes(rb, ref, bdmax):
    from rav1d_b200.synth.itxgen import valid_txtps
    rng = np.random.default_rng(bdmax)
    n = 0
    for tx in range(19):
        for txtp in range(17):
            assert bool(rb.itx_valid(tx, txtp)) == bool(ref.ref_itx_has(tx, txtp)) == (txtp in valid_txtps(tx))
            if not rb.itx_valid(tx, txtp):
                continue
            for variant in ("dc", "sub", "full", "full", "extreme"):
                _one_case(rb, ref, rng, tx, txtp, bdmax, variant)
                n += 1
            _one_case(rb, ref, rng, tx, txtp, bdmax, "full", neg_stride=True)
    assert n == 156 * 5


@pytest.mark.gpu
def test_config1_itx_batch_8bit(rb, ref):
    """BASELINE.json configs[0] through the batch path: every valid (size, type) pair, N blocks each with the
    checkasm coefficient recipe (5 variants), laid out in one 4096x4096 8-bit plane whose pixels start random;
    the whole plane must equal the reference's itxfm_add applied block by block."""
    from rav1d_b200.lib import TX_DIMS
    from rav1d_b200.synth.itxgen import gen_coefs, valid_txtps
    N, W = 384, 4096
    rng = np.random.default_rng(1)
    hdr = rb.FrameHeader()
    hdr.width, hdr.height, hdr.bpc, hdr.layout = W, W, 8, rb.LAYOUT_I400
    plane = rng.integers(0, 256, size=(W, W)).astype(np.uint8)
    exp = plane.copy()
    items, coefs, counts = [], [], np.zeros(19, np.int32)
    cf_off, cx, cy, band = 0, 0, 0, 0
    variants = ("dc", "sub", "full", "full", "extreme")
    for tx in range(19):
        w, h = TX_DIMS[tx]
        sw, sh = min(w, 32), min(h, 32)
        for txtp in valid_txtps(tx):
            c, eob = gen_coefs(rng, tx, txtp, 255, N, "full")
            for k, v in enumerate(variants):               # a fifth of the blocks per variant
                ck, ek = gen_coefs(rng, tx, txtp, 255, N // 5, v)
                c[k * (N // 5):(k + 1) * (N // 5)] = ck
                eob[k * (N // 5):(k + 1) * (N // 5)] = ek
            c = np.clip(c, -32768, 32767).astype(np.int16)
            it = np.zeros(N, rb.ITX_ITEM_DT)
            for i in range(N):
                if cx + w > W:
                    cx, cy = 0, cy + band
                    band = 0
                band = max(band, h)
                it[i] = (cf_off + i * sw * sh, cx, cy, 0, tx, txtp, 0, eob[i], 0)
                cx += w
            assert cy + band <= W
            items.append(it); coefs.append(c.reshape(-1)); counts[tx] += N
            cf_off += N * sw * sh
            cw = c.copy()
            for i in range(N):
                ref.ref_itxfm_add(tx, txtp, C.c_void_p(exp.ctypes.data + int(it["y"][i]) * W + int(it["x"][i])), W, ptr(cw[i]), int(eob[i]), 255)
            assert not cw.any()                              # the reference consumed (zeroed) every block
    items, coefs = np.concatenate(items), np.concatenate(coefs)
    assert len(items) == 156 * N
    f = C.c_void_p()
    rb.check(rb.frame_create(C.byref(f), C.byref(hdr), len(coefs), len(items), 1))
    try:
        rb.np_view(rb.frame_coef_buffer(f), np.int16, len(coefs))[:] = coefs
        rb.np_view(rb.frame_itx_items(f), rb.ITX_ITEM_DT, len(items))[:] = items
        data = (C.c_void_p * 3)(plane.ctypes.data, None, None)
        strides = (C.c_ssize_t * 2)(W, 0)
        rb.check(rb.frame_upload_planes(f, 0, data, strides))
        rb.check(rb.frame_submit(f, len(coefs), (C.c_int32 * 19)(*[int(v) for v in counts]), 0, rb.STAGE_RECON, 1))
        rb.check(rb.frame_wait(f))
        got = np.zeros_like(plane)
        data = (C.c_void_p * 3)(got.ctypes.data, None, None)
        rb.check(rb.frame_readback(f, data, strides))
    finally:
        rb.frame_destroy(f)
    assert np.array_equal(exp, got), np.argwhere(exp != got)[:5]


@pytest.mark.gpu
def test_itx_dsp_table_slots(rb, ref):
    """rb200_itx_dsp_init fills the same slots as the reference's init and each slot works."""
    from rav1d_b200.lib import InvTxfmDSPContext, TX_DIMS
    from rav1d_b200.synth.itxgen import gen_coefs
    ctx = InvTxfmDSPContext()
    rb.itx_dsp_init(C.byref(ctx), 10)
    rng = np.random.default_rng(3)
    for tx in range(19):
        for txtp in range(17):
            fn = ctx.itxfm_add[tx][txtp]
            assert bool(fn) == bool(ref.ref_itx_has(tx, txtp))
            if not fn or (tx * 17 + txtp) % 7:
                continue
            w, h = TX_DIMS[tx]
            coef, eob = gen_coefs(rng, tx, txtp, 1023, 1, "full")
            c0 = np.zeros(32 * 32, dtype=np.int32); c0[:coef.shape[1]] = coef[0]
            c1 = c0.copy()
            d0 = rng.integers(0, 1024, size=(h, w)).astype(np.uint16); d1 = d0.copy()
            ref.ref_itxfm_add(tx, txtp, ptr(d0), w * 2, ptr(c0), int(eob[0]), 1023)
            fn(d1.ctypes.data, w * 2, c1.ctypes.data, int(eob[0]), 1023)
            assert np.array_equal(d0, d1) and np.array_equal(c0, c1)


@pytest.mark.gpu
def test_itx_invalid_combination_is_an_error(rb):
    d = np.zeros((64, 64), dtype=np.uint8); c = np.zeros(1024, dtype=np.int16)
    assert rb.itxfm_add(4, 1, ptr(d), 64, ptr(c), 1, 255) != 0   # 64x64 ADST_DCT does not exist
    assert b"no such transform" in rb.last_error()
