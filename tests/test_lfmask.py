"""Loop-filter masks and levels built on the device from per-block records.

Reference: rav1d_create_lf_mask_intra / _inter (src/lf_mask.rs:380-606; C: src/lf_mask.c:286-406), the noskip mask of
decode_b (src/decode.c:1996-2005) and the tile-edge fix-ups of src/lf_apply_tmpl.c:331-400.

Three anchors:
  * tests/golden/streams_lfb.npz -- the arguments of every create_lf_mask call the REFERENCE DECODER made for 31 frames of
    16 conformance streams (oracle/ref_dump.c interposes the two functions), next to the Av1Filter / level arrays the same
    decoder ended up with (tests/golden/streams.npz).  CPU: the oracle replay reproduces the decoder's arrays (pins the
    replay rules).  GPU: the kernels reproduce them too, multi-tile streams included.
  * synthetic partition trees in every pixel layout, both superblock sizes and odd picture sizes: kernels vs the
    oracle (the reference's own functions replayed in decode order, oracle/ref_lfmask.c), records shuffled.
  * the frame path fed with records instead of masks delivers the decoder's filtered picture.
"""
import ctypes as C
import os
import types

import numpy as np
import pytest

import streamdump

HAVE = os.path.exists(streamdump.GOLDEN) and os.path.exists(streamdump.GOLDEN_LFB)
GOLD = dict(streamdump.load_golden()) if HAVE else {}
LFB = streamdump.load_golden_lfb() if HAVE else {}
KEYS = list(LFB)


def oracle_lf(ref, blocks, w, h, layout, sb128):
    """The reference's functions replayed over `blocks`.  Records are first put in the decode order of a one-tile frame
    (superblocks in raster order, order inside a superblock kept): the left context then carries across what were
    tile-column boundaries and the above context across tile rows, which is what the decoder's fix-ups arrive at."""
    from rav1d_b200 import lib
    bw, bh = ((w + 7) >> 3) << 1, ((h + 7) >> 3) << 1
    w4, h4 = (w + 3) >> 2, (h + 3) >> 2
    sb128w, sb128h, b4_stride = (bw + 31) >> 5, (bh + 31) >> 5, (bw + 31) & ~31
    sh = 5 if sb128 else 4
    order = np.argsort((blocks["by"].astype(np.int64) >> sh) * 4096 + (blocks["bx"] >> sh), kind="stable")
    blk = np.ascontiguousarray(blocks[order])
    masks = np.zeros(sb128w * sb128h, lib.AV1_FILTER_DT)
    levels = np.zeros((32 * sb128h, b4_stride, 4), np.uint8)
    r = ref.ref_lf_build(blk.ctypes.data, len(blk), w4, h4, layout, sb128, b4_stride, sb128w, sb128h, masks.ctypes.data,
                         levels.ctypes.data)
    assert r == 0
    return masks, levels


def in_frame_levels(lv, w, h, layout):
    w4, h4 = (w + 3) >> 2, (h + 3) >> 2
    ss_hor, ss_ver = int(layout != 3), int(layout == 1)
    out = [lv[:h4, :w4, :2]]
    if layout != 0:
        out.append(lv[:(h4 + ss_ver) >> ss_ver, :(w4 + ss_hor) >> ss_hor, 2:])
    return out


def assert_lf_equal(exp_masks, exp_levels, got_masks, got_levels, w, h, layout, what, noskip=True):
    for name in ("filter_y", "filter_uv") + (("noskip_mask",) if noskip else ()):
        a, b = exp_masks[name], got_masks[name]
        if not np.array_equal(a, b):
            bad = np.argwhere(a != b)
            raise AssertionError(f"{what}: {name} differs in {len(bad)} words, first {tuple(bad[0])}: "
                                 f"expected {a[tuple(bad[0])]:#06x} got {b[tuple(bad[0])]:#06x}")
    for k, (a, b) in enumerate(zip(in_frame_levels(exp_levels, w, h, layout), in_frame_levels(got_levels, w, h, layout))):
        if not np.array_equal(a, b):
            bad = np.argwhere(a != b)
            raise AssertionError(f"{what}: level bytes {'y' if not k else 'uv'} differ at {len(bad)} places, first {tuple(bad[0])}")


def with_intra_skip(blocks, masks, sb128w):
    """The intra call has no skip argument; it only matters for the noskip mask, so take it from the decoder's mask:
    a block whose bit is clear was skipped (a set bit is reproduced by marking the block not skipped)."""
    from rav1d_b200 import lib
    b = blocks.copy()
    bx, by = b["bx"].astype(np.int64), b["by"].astype(np.int64)
    word = masks["noskip_mask"][(by >> 5) * sb128w + (bx >> 5), (by & 31) >> 1, (bx & 16) >> 4]
    clear = ((word >> (bx & 15)) & 1) == 0
    intra = (b["flags"] & lib.LFB_INTRA) != 0
    b["flags"][intra & clear] |= lib.LFB_SKIP
    return b


def test_fixture_covers_tiles_layouts_and_block_kinds():
    from rav1d_b200 import lib
    assert len(KEYS) >= 25
    fr = [GOLD[k] for k in KEYS]
    assert {s.layout for s in fr} == {0, 1, 2, 3} and {s.hdr.sb128 for s in fr} == {0, 1}
    assert any(s.tiles != (1, 1) for s in fr)
    allb = np.concatenate([LFB[k] for k in KEYS])
    assert (allb["flags"] & lib.LFB_INTRA).any() and not (allb["flags"] & lib.LFB_INTRA).all()
    assert (allb["tx_split"] != 0).any() and len(np.unique(allb["bs"])) >= 15


@pytest.mark.parametrize("key", KEYS)
def test_oracle_replay_reproduces_the_decoders_masks(ref, key):
    s = GOLD[key]
    m, lv = oracle_lf(ref, LFB[key], s.w, s.h, s.layout, s.hdr.sb128)
    n = s.geom.sb128w * s.geom.sb128h
    assert_lf_equal(s.masks[:n], s.levels, m, lv, s.w, s.h, s.layout, key, noskip=False)


def test_oracle_noskip_follows_the_skip_flags(ref):
    from rav1d_b200 import lib
    from rav1d_b200.synth import framegen
    b = framegen.generate_lf_blocks(200, 136, 1, 1, seed=5)
    m, _ = oracle_lf(ref, b, 200, 136, 1, 1)
    b2 = b.copy(); b2["flags"] |= lib.LFB_SKIP
    m2, _ = oracle_lf(ref, b2, 200, 136, 1, 1)
    assert m["noskip_mask"].any() and not m2["noskip_mask"].any()


def _bare_frame(w, h, bpc, layout, sb128, blocks):
    """The least a DeviceFrame needs: header, empty batch, zero masks (cdef_idx random: it must survive)."""
    from rav1d_b200 import lib
    from rav1d_b200.synth import framegen
    s = types.SimpleNamespace()
    hdr = lib.FrameHeader()
    hdr.width, hdr.height, hdr.bpc, hdr.layout, hdr.sb128 = w, h, bpc, layout, sb128
    hdr.lf_level_y[0], hdr.lf_level_y[1], hdr.lf_level_u, hdr.lf_level_v = 20, 20, 10, 10
    hdr.cdef_damping = 3
    s.hdr, s.w, s.h, s.layout = hdr, w, h, layout
    s.n_coefs, s.coef = 0, np.zeros(0, np.int32 if bpc > 8 else np.int16)
    s.itx_items, s.mc_items = np.zeros(0, lib.ITX_ITEM_DT), np.zeros(0, lib.MC_ITEM_DT)
    s.itx_counts = np.zeros(19, np.int32)
    g = framegen.geometry(hdr)
    n = g.sb128w * g.sb128h
    s.masks = np.zeros(n, lib.AV1_FILTER_DT)
    s.masks["cdef_idx"] = np.random.default_rng(1).integers(-1, 8, size=(n, 4))
    s.masks["filter_y"] = 0x5a5a          # must be ignored and overwritten on the device
    s.levels = np.full((32 * g.sb128h, g.b4_stride, 4), 63, np.uint8)
    s.lut = framegen.calc_eih(0)
    s.lr_masks = np.zeros(n, lib.AV1_RESTORATION_DT)
    s.lf_blocks = blocks
    return s


CASES = [(200, 136, 1, 1), (200, 136, 1, 0), (130, 66, 3, 1), (258, 130, 2, 0), (96, 96, 0, 1), (1922, 1082, 1, 1), (37, 21, 1, 0)]


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,layout,sb128", CASES)
def test_device_masks_match_the_oracle_on_partition_trees(ref, rb, w, h, layout, sb128):
    from rav1d_b200 import lib
    from rav1d_b200.synth import framegen
    blocks = framegen.generate_lf_blocks(w, h, layout, sb128, seed=w + h + layout)
    em, el = oracle_lf(ref, blocks, w, h, layout, sb128)
    shuffled = blocks[np.random.default_rng(3).permutation(len(blocks))]      # the result may not depend on the order
    s = _bare_frame(w, h, 8, layout, sb128, shuffled)
    d = framegen.DeviceFrame(s)
    try:
        d.load_batch()
        d.submit(lib.STAGE_DEBLOCK, 1)
        gm, gl = d.download_lf()
        assert_lf_equal(em, el, gm, gl, w, h, layout, f"{w}x{h} layout {layout} sb128 {sb128}")
        assert np.array_equal(gm["cdef_idx"], s.masks["cdef_idx"])
        # a second frame with fewer blocks on the same context: nothing of the first may linger in the masks
        few = framegen.generate_lf_blocks(w, h, layout, sb128, seed=99, min_log=3)
        em2, el2 = oracle_lf(ref, few, w, h, layout, sb128)
        lib.check(lib.frame_reserve_lf_blocks(d.h, len(few)))
        lib.np_view(lib.frame_lf_blocks(d.h), lib.LF_BLOCK_DT, len(few))[:] = few
        lib.check(lib.frame_set_lf_block_count(d.h, len(few)))
        d.submit(lib.STAGE_DEBLOCK, 1)
        gm2, gl2 = d.download_lf()
        assert_lf_equal(em2, el2, gm2, gl2, w, h, layout, "second frame")
    finally:
        d.close()


@pytest.mark.gpu
@pytest.mark.parametrize("key", KEYS)
def test_device_masks_match_the_reference_decoder(rb, key):
    """Records the decoder produced -> the arrays the decoder produced (tile-edge fix-ups included), and the frame path fed
    with the records delivers the decoder's filtered picture."""
    from rav1d_b200 import lib
    from rav1d_b200.synth import framegen
    import copy
    s = copy.copy(GOLD[key])
    n = s.geom.sb128w * s.geom.sb128h
    s.lf_blocks = with_intra_skip(LFB[key], s.masks, s.geom.sb128w)
    host_masks = s.masks
    s.masks = host_masks.copy()
    s.masks["filter_y"] = 0; s.masks["filter_uv"] = 0; s.masks["noskip_mask"] = 0     # only cdef_idx is left to the host
    s.levels = np.zeros_like(s.levels)
    d = framegen.DeviceFrame(s)
    try:
        d.load_batch()
        d.upload(0, s.pre)
        d.submit(s.stages, 1)
        d.wait()
        gm, gl = d.download_lf()
        assert_lf_equal(host_masks[:n], GOLD[key].levels, gm[:n], gl, s.w, s.h, s.layout, key)
        assert np.array_equal(gm["cdef_idx"][:n], host_masks["cdef_idx"][:n])
        got = d.readback()
        for p, (a, b) in enumerate(zip(streamdump.visible(s, s.post), streamdump.visible(s, got))):
            assert np.array_equal(a, b), f"{key}: filtered plane {p} differs"
    finally:
        d.close()


@pytest.mark.parametrize("w,h,bpc", [(200, 120, 8), (640, 360, 10)])
def test_synthetic_frame_records_describe_its_masks(ref, w, h, bpc):
    """The block records framegen.generate() emits, run through the reference's functions, give the masks and levels the
    generator wrote by hand -- so a bench or test may hand either form to the frame path."""
    from rav1d_b200.synth import framegen
    s = framegen.generate(w, h, bpc, seed=w)
    m, lv = oracle_lf(ref, s.lf_block_records, w, h, 1, 1)
    assert_lf_equal(s.masks, s.levels, m, lv, w, h, 1, f"{w}x{h}")


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,bpc", [(200, 120, 8), (640, 360, 10), (264, 200, 12)])
def test_frame_path_with_records_matches_the_reference_drivers(ref, rb, w, h, bpc):
    """Recon + deblock + CDEF + LR with the masks built on the device against the reference's frame drivers."""
    import framecheck
    from rav1d_b200.synth import framegen
    s = framegen.generate(w, h, bpc, seed=w + bpc)
    a = framecheck.oracle_frame(ref, s, 15)
    s.lf_blocks = s.lf_block_records
    keep = s.masks
    s.masks = keep.copy()
    s.masks["filter_y"] = 0; s.masks["filter_uv"] = 0; s.masks["noskip_mask"] = 0
    s.levels = np.zeros_like(s.levels)
    b = framecheck.product_frame(s, 15)
    framecheck.assert_planes_equal(a, b, f"{w}x{h}@{bpc} with lf records")
