"""Post-filter parity on REAL AV1 streams.

tests/golden/streams.npz (tools/make_stream_fixtures.py) holds, for frames of 16 conformance
streams of the reference tree (8 / 10 / 12 bit; 4:0:0, 4:2:0, 4:2:2, 4:4:4; 64- and 128-pixel
superblocks; multi-tile; odd sizes), the picture before the in-loop filters, the filter metadata
exactly as the reference decoder built it, and the decoder's picture after deblock + CDEF + loop
restoration -- captured by interposing dav1d_filter_sbrow_* in the reference decoder
(oracle/ref_dump.c).  The CUDA frame path must turn the former into the latter, bit for bit.
"""
import os

import numpy as np
import pytest

import refharness
import streamdump

GOLD = streamdump.load_golden() if os.path.exists(streamdump.GOLDEN) else []
IDS = [k for k, _ in GOLD]


def _assert_equal(s, exp, got, what):
    for p, (a, b) in enumerate(zip(streamdump.visible(s, exp), streamdump.visible(s, got))):
        if not np.array_equal(a, b):
            bad = np.argwhere(a != b)
            raise AssertionError(f"{what}: plane {p} differs at {len(bad)} px, first (x={bad[0][1]}, y={bad[0][0]}): "
                                 f"expected {a[tuple(bad[0])]} got {b[tuple(bad[0])]}")


def test_golden_covers_the_formats():
    assert len(GOLD) >= 30
    assert {s.bpc for _, s in GOLD} == {8, 10, 12}
    assert {s.layout for _, s in GOLD} == {0, 1, 2, 3}
    assert {s.hdr.sb128 for _, s in GOLD} == {0, 1}
    assert any(s.tiles != (1, 1) for _, s in GOLD) and any(s.w % 8 for _, s in GOLD)
    assert {st for _, s in GOLD for st in (2, 4, 8) if s.stages & st} == {2, 4, 8}


@pytest.mark.parametrize("key", IDS)
def test_golden_replays_through_the_reference_drivers(ref, key):
    """The fixture is self-consistent under the oracle's frame harness (CPU): pins both."""
    s = dict(GOLD)[key]
    cur = refharness.RefFrame(ref, s, 1)
    try:
        cur.load_filter_meta(); cur.set_planes(s.pre); cur.filter(s.stages)
        _assert_equal(s, s.post, cur.get_planes(), key)
    finally:
        cur.close()


@pytest.mark.skipif(not streamdump.available(), reason="reference tree / oracle build not present (GPU box)")
def test_golden_matches_a_fresh_dump_of_the_reference_decoder():
    for rel in ("8-bit/data/00000658.ivf", "10-bit/data/00000676.ivf", "12-bit/argon/test15240.obu"):
        for s in streamdump.dump(os.path.join(streamdump.REF_DATA, rel), 2):
            g = dict(GOLD)[f"{rel}#{s.index}"]
            assert np.array_equal(s.masks.view(np.uint8), g.masks.view(np.uint8)) and np.array_equal(s.levels, g.levels)
            for a, b in zip(streamdump.visible(s, s.post), streamdump.visible(g, g.post)):
                assert np.array_equal(a, b)
            for a, b in zip(streamdump.visible(s, s.pre), streamdump.visible(g, g.pre)):
                assert np.array_equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("key", IDS)
def test_stream_postfilters_bit_exact(rb, key):
    from rav1d_b200.synth import framegen
    s = dict(GOLD)[key]
    d = framegen.DeviceFrame(s)
    try:
        d.load_batch()
        d.upload(0, s.pre)
        d.submit(s.stages); d.wait()
        _assert_equal(s, s.post, d.readback(), key)
    finally:
        d.close()


# ---------------------------------------------------------------- the reference's picture-size sweep
GOLD_SIZES = streamdump.load_golden(streamdump.GOLDEN_SIZES) if os.path.exists(streamdump.GOLDEN_SIZES) else []


@pytest.mark.parametrize("key", [k for k, _ in GOLD_SIZES])
def test_size_sweep_replays_through_the_reference_drivers(ref, key):
    s = dict(GOLD_SIZES)[key]
    cur = refharness.RefFrame(ref, s, 1)
    try:
        cur.load_filter_meta(); cur.set_planes(s.pre); cur.filter(s.stages)
        _assert_equal(s, s.post, cur.get_planes(), key)
    finally:
        cur.close()


@pytest.mark.gpu
@pytest.mark.parametrize("key", [k for k, _ in GOLD_SIZES])
def test_stream_size_sweep_bit_exact(rb, key):
    """Pictures of 16 .. 66 and 196 .. 226 pixels a side (tests/dav1d-test-data/8-bit/size): every edge rule of the
    post-filters at sizes that are not multiples of 8, 64 or 128."""
    from rav1d_b200.synth import framegen
    s = dict(GOLD_SIZES)[key]
    d = framegen.DeviceFrame(s)
    try:
        d.load_batch()
        d.upload(0, s.pre)
        d.submit(s.stages); d.wait()
        _assert_equal(s, s.post, d.readback(), key)
    finally:
        d.close()


# ---------------------------------------------------------------- super-resolution frames of real streams
GOLD_SR = streamdump.load_golden(streamdump.GOLDEN_SR) if os.path.exists(streamdump.GOLDEN_SR) else []


def test_super_resolution_fixtures_cover_ratios_and_depths():
    assert len(GOLD_SR) >= 10 and {s.bpc for _, s in GOLD_SR} == {8, 10}
    assert len({(s.w, s.out_w) for _, s in GOLD_SR}) >= 6 and all(s.out_w > s.w for _, s in GOLD_SR)
    assert any(s.stages & 8 for _, s in GOLD_SR) and any(not s.stages & 8 for _, s in GOLD_SR)


@pytest.mark.skipif(not streamdump.available(), reason="reference tree / oracle build not present (GPU box)")
def test_super_resolution_fixtures_match_a_fresh_dump():
    for rel in ("8-bit/data/00000855.ivf", "10-bit/data/00000832.ivf"):
        for s in streamdump.dump(os.path.join(streamdump.REF_DATA, rel), 2, sr=True):
            g = dict(GOLD_SR)[f"{rel}#{s.index}"]
            assert (s.w, s.out_w, s.stages) == (g.w, g.out_w, g.stages)   # (unused lr_mask entries are uninitialised memory)
            for a, b in zip(streamdump.visible(s, s.post, out=True), streamdump.visible(g, g.post, out=True)):
                assert np.array_equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("key", [k for k, _ in GOLD_SR])
def test_stream_super_resolution_bit_exact(rb, key):
    """Deblock + CDEF + horizontal upscaling + loop restoration of frames coded with super-resolution:
    the coded-width picture before the filters must become the reference decoder's upscaled output."""
    from rav1d_b200.synth import framegen
    s = dict(GOLD_SR)[key]
    d = framegen.DeviceFrame(s)
    try:
        d.load_batch()
        d.upload(0, s.pre)
        d.submit(s.stages); d.wait()
        got = d.readback()
    finally:
        d.close()
    for p, (a, b) in enumerate(zip(streamdump.visible(s, s.post, out=True), streamdump.visible(s, got, out=True))):
        if not np.array_equal(a, b):
            bad = np.argwhere(a != b)
            raise AssertionError(f"{key}: plane {p} differs at {len(bad)} px, first (x={bad[0][1]}, y={bad[0][0]}): "
                                 f"expected {a[tuple(bad[0])]} got {b[tuple(bad[0])]}; rows {np.unique(bad[:, 0])[:10]}")


@pytest.mark.gpu
@pytest.mark.parametrize("key", [k for k, _ in GOLD_SR][:4])
def test_super_resolution_then_film_grain(rb, ref, key):
    """Film grain on a picture coded with super-resolution goes onto the UPSCALED output (rav1d_apply_grain runs on the
    output picture, src/lib.rs:604): the stage chain deblock + CDEF + upscaling + LR + grain against the reference's grain
    applied to the reference decoder's upscaled output."""
    import ctypes as C
    import refharness
    from rav1d_b200.synth import framegen
    s = dict(GOLD_SR)[key]
    if s.layout == 0:
        pytest.skip("monochrome")
    out_w = getattr(s, "out_w", s.w)
    rng = np.random.default_rng(out_w)
    d = framegen.random_film_grain(rng, lag=2, overlap=1)
    # the oracle: a frame object of the output size carries the upscaled picture through dav1d_apply_grain
    carrier = framegen.generate(out_w, s.h, s.bpc, seed=1)
    if carrier.hdr.layout != s.layout:
        pytest.skip("the synthetic carrier frame is 4:2:0 only")
    cur = refharness.RefFrame(ref, carrier, 1)
    try:
        post = streamdump.visible(s, s.post, out=True)
        planes = cur.get_planes()
        for p in range(3):
            planes[p][:] = 0
            planes[p][:post[p].shape[0], :post[p].shape[1]] = post[p]
        cur.set_planes(planes)
        exp = [a[:post[p].shape[0], :post[p].shape[1]] for p, a in enumerate(cur.apply_grain(d, 0))]
    finally:
        cur.close()
    dev = framegen.DeviceFrame(s)
    try:
        dev.load_batch()
        dev.upload(0, s.pre)
        rb.check(rb.frame_set_film_grain(dev.h, C.byref(d), 0))
        dev.submit(s.stages | rb.STAGE_FILM_GRAIN); dev.wait()
        got = streamdump.visible(s, dev.readback(), out=True)
    finally:
        dev.close()
    for p in range(3):
        assert np.array_equal(exp[p], got[p]), (key, p, int((exp[p] != got[p]).sum()))
    assert not np.array_equal(exp[0], post[0])


# ---------------------------------------------------------------- film grain on real streams
def _load_grain():
    if not os.path.exists(streamdump.GOLDEN_GRAIN):
        return []
    z = np.load(streamdump.GOLDEN_GRAIN, allow_pickle=False)
    return [(str(k), z) for k in z["index"]]


GRAIN = _load_grain()


@pytest.mark.gpu
@pytest.mark.parametrize("key", [k for k, _ in GRAIN])
def test_stream_film_grain_bit_exact(rb, key):
    """dav1d_apply_grain's input / parameters / output captured from the reference decoder on the
    film-grain conformance streams (tests/dav1d-test-data/{8,10}-bit/film_grain); the frame-level
    RB200_STAGE_FILM_GRAIN must reproduce the output."""
    import ctypes as C
    from rav1d_b200 import lib
    from rav1d_b200.synth import framegen
    z = dict(GRAIN)[key]
    w, h, bpc, layout, is_id = (int(v) for v in z[f"{key}/ints"])
    fg = lib.FilmGrainData.from_buffer_copy(z[f"{key}/fg"].tobytes())
    s = streamdump.StreamFrame()
    s.w, s.h, s.bpc, s.layout = w, h, bpc, layout
    s.aw, s.ah = (w + 127) & ~127, (h + 127) & ~127
    hdr = lib.FrameHeader()
    hdr.width, hdr.height, hdr.bpc, hdr.layout = w, h, bpc, layout
    s.hdr = hdr
    ss_ver, ss_hor = int(layout == 1), int(layout != 3)
    n_planes = 1 if layout == 0 else 3
    pdt = np.uint16 if bpc > 8 else np.uint8

    def padded(kind):
        out = []
        for p in range(n_planes):
            a = np.zeros((s.ah >> ss_ver if p else s.ah, s.aw >> ss_hor if p else s.aw), pdt)
            c = z[f"{key}/{kind}{p}"]
            a[:c.shape[0], :c.shape[1]] = c
            out.append(a)
        return out
    inp, exp = padded("in"), padded("out")
    s.ref = inp
    s.n_coefs, s.itx_items, s.mc_items = 0, np.zeros(0, lib.ITX_ITEM_DT), np.zeros(0, lib.MC_ITEM_DT)
    d = framegen.DeviceFrame(s)
    try:
        d.upload(0, inp)
        lib.check(lib.frame_set_film_grain(d.h, C.byref(fg), is_id))
        counts = (C.c_int32 * 19)()
        lib.check(lib.frame_submit(d.h, 0, counts, 0, lib.STAGE_FILM_GRAIN, 0), "frame_submit")
        d.wait()
        _assert_equal(s, exp, d.readback(), key)
        assert not np.array_equal(exp[0], inp[0])
    finally:
        d.close()
