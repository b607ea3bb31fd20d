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
