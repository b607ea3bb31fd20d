"""Frame-level parity: the product's stream-ordered stage launches (rb200_frame_*, C ABI)
against the reference's own frame drivers (dav1d_filter_sbrow_* and the DSP tables, run by
oracle/ref_frame.c) on the same synthetic frame batch -- stage by stage, then end to end."""
import ctypes as C

import numpy as np
import pytest

import framecheck
from rav1d_b200.synth import framegen

R, D, Cd, L = 1, 2, 4, 8

SMALL = [(176, 144, 8), (200, 120, 10), (264, 200, 12), (100, 68, 8), (64, 64, 10), (24, 16, 8)]


def _check(ref, s, stages, start=None):
    a = framecheck.oracle_frame(ref, s, stages, start_planes=start)
    b = framecheck.product_frame(s, stages, start_planes=start)
    framecheck.assert_planes_equal(a, b, f"{s.w}x{s.h}@{s.bpc} stages={stages}")
    return a


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,bpc", SMALL)
def test_recon_stage(rb, ref, w, h, bpc):
    s = framegen.generate(w, h, bpc, seed=w + bpc)
    _check(ref, s, R)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,bpc", SMALL)
@pytest.mark.parametrize("stages", [D, Cd, L, D | Cd, Cd | L, D | Cd | L])
def test_filter_stages(rb, ref, w, h, bpc, stages):
    s = framegen.generate(w, h, bpc, seed=w + bpc + stages)
    start = framegen.recon_input_planes(s)
    out = _check(ref, s, stages, start)
    if w >= 100:
        assert not np.array_equal(out[0], start[0][:h, :w])   # the stage did something


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,bpc", SMALL + [(640, 360, 10)])
def test_all_stages(rb, ref, w, h, bpc):
    s = framegen.generate(w, h, bpc, seed=w * 3 + bpc)
    _check(ref, s, R | D | Cd | L)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,bpc", [(176, 144, 8), (200, 120, 10), (264, 200, 12), (640, 360, 10)])
def test_compound_blocks(rb, ref, w, h, bpc):
    """Half of the blocks predicted from two references (avg / w_avg / segmentation mask / wedge), BASELINE config 5."""
    s = framegen.generate(w, h, bpc, seed=w + 1, comp_frac=0.5)
    assert len(s.comp_items) > 10 and set(s.comp_items["comp_type"]) == {0, 1, 2, 3}
    _check(ref, s, R)
    _check(ref, s, R | D | Cd | L)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,bpc", [(256, 192, 8), (320, 256, 10), (256, 192, 12)])
def test_compound_block_sizes(rb, ref, w, h, bpc):
    """Compound blocks from 8x8 to 32x32 (the generator only makes 16x16): smaller ones leave part of
    their cell to the start picture, larger ones overwrite translational neighbours (compound blocks run
    after the single-reference ones in both the oracle replay and the batch)."""
    s = framegen.generate(w, h, bpc, seed=w + 7, comp_frac=0.4)
    rng = np.random.default_rng(w + bpc)
    it = s.comp_items
    cells = {(int(x) // 16, int(y) // 16) for x, y in zip(it["x"], it["y"])}
    grown = set()
    for k in range(len(it)):
        cx, cy = int(it["x"][k]) // 16, int(it["y"][k]) // 16
        big = [(32, 32), (32, 16), (16, 32)][int(rng.integers(0, 3))]
        need = {(cx + dx, cy + dy) for dx in range(big[0] // 16) for dy in range(big[1] // 16)} - {(cx, cy)}
        if (rng.random() < 0.5 and not (need & cells) and not (need & grown)
                and it["x"][k] + big[0] <= (w & ~15) and it["y"][k] + big[1] <= (h & ~15)):
            it["w"][k], it["h"][k] = big
            grown |= need
        elif (cx, cy) not in grown:
            it["w"][k], it["h"][k] = [(8, 8), (8, 16), (16, 8), (16, 16)][int(rng.integers(0, 4))]
    assert {(int(a), int(b)) for a, b in zip(it["w"], it["h"])} >= {(8, 8), (32, 32), (16, 32), (32, 16)}
    _check(ref, s, R)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,bpc", [(176, 144, 8), (208, 128, 10), (256, 192, 12)])
def test_warped_blocks(rb, ref, w, h, bpc):
    """Affine-warped blocks next to translational and compound ones, all planes."""
    s = framegen.generate(w, h, bpc, seed=w + 2, comp_frac=0.3, warp_frac=0.3)
    assert len(s.warp_items) > 5
    _check(ref, s, R)
    _check(ref, s, R | D | Cd | L)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,bpc,rsize", [(176, 144, 8, (128, 112)), (208, 128, 10, (320, 160)), (256, 192, 12, (200, 300)),
                                           (320, 192, 10, (640, 96))])
def test_scaled_reference_blocks(rb, ref, w, h, bpc, rsize):
    """Blocks predicted from a reference of another size (mc_scaled through the batch), next to
    translational and compound ones; down- and up-scaling, different ratios per axis."""
    s = framegen.generate(w, h, bpc, seed=w + 4, comp_frac=0.2, scaled_frac=0.4, scaled_size=rsize)
    assert len(s.scaled_items) > 30
    _check(ref, s, R)
    _check(ref, s, R | D | Cd | L)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,bpc,rsize", [(320, 192, 8, (256, 160)), (256, 160, 10, (384, 200)), (208, 144, 12, (104, 72))])
def test_obmc_strips_from_scaled_references(rb, ref, w, h, bpc, rsize):
    """OBMC strips whose neighbour predicts from a reference of another size: mc_scaled for the strip (its prediction
    height picks the filters), then blend_h / blend_v (src/recon.rs:2205-2309 with the scaled branch of mc())."""
    s = framegen.generate(w, h, bpc, seed=w + 9, scaled_frac=0.5, scaled_size=rsize, scaled_obmc_frac=0.6)
    assert s.n_scaled[1] > 10 and s.n_scaled[2] > 10
    _check(ref, s, R)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,bpc,rsize", [(320, 192, 8, (256, 160)), (256, 160, 10, (384, 200)), (208, 144, 12, (104, 72))])
def test_compound_blocks_from_scaled_references(rb, ref, w, h, bpc, rsize):
    """Compound blocks with one or both predictions from a reference of another size (prep_8tap_scaled inside the compound
    kernel, position and step from the two sizes), under every compound type."""
    s = framegen.generate(w, h, bpc, seed=w + 13, comp_frac=0.5, scaled_frac=0.2, scaled_size=rsize, comp_scaled_frac=0.7)
    assert (s.comp_items["ref"] == 2).any(axis=1).sum() > 10
    _check(ref, s, R)


@pytest.mark.gpu
@pytest.mark.parametrize("layout", [0, 1, 2, 3])
@pytest.mark.parametrize("w,h,bpc", [(192, 128, 8), (256, 160, 10), (176, 144, 12)])
def test_recon_all_layouts(rb, ref, layout, w, h, bpc):
    """Prediction (translational, compound incl. wedge / segmentation masks, warped, OBMC) and residual in 4:0:0,
    4:2:0, 4:2:2 and 4:4:4: chroma block sizes, vectors, phases, mask sub-sampling and transform sizes all change."""
    s = framegen.generate_recon_layout(w, h, bpc, layout, seed=w + 10 * layout + bpc, warp_frac=0.15)
    assert len(s.comp_items) > 5 and len(s.warp_items) >= 2 and min(s.n_obmc) > 3
    a = framecheck.oracle_frame(ref, s, R)
    b = framecheck.product_frame(s, R)
    assert len(a) == (1 if layout == 0 else 3)
    framecheck.assert_planes_equal(a, b, f"layout {layout} {w}x{h}@{bpc}")


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,bpc", [(176, 144, 8), (208, 128, 10), (256, 192, 12)])
def test_obmc_strips(rb, ref, w, h, bpc):
    """Overlapped block MC on top of translational, compound and warped neighbours (BASELINE config 5)."""
    s = framegen.generate(w, h, bpc, seed=w + 3, comp_frac=0.25, warp_frac=0.1, obmc_frac=0.5)
    assert s.n_obmc[0] > 10 and s.n_obmc[1] > 10
    _check(ref, s, R)
    _check(ref, s, R | D | Cd | L)


@pytest.mark.gpu
@pytest.mark.parametrize("unit_log2", [7, 8])
def test_lr_unit_sizes(rb, ref, unit_log2):
    """Restoration units larger than the 64-pixel default, incl. the 1.5x last unit."""
    s = framegen.generate(424, 296, 10, seed=unit_log2)
    s.hdr.lr_unit_size_log2[0] = unit_log2
    s.hdr.lr_unit_size_log2[1] = unit_log2 - 1
    _check(ref, s, D | Cd | L, framegen.recon_input_planes(s))


@pytest.mark.gpu
def test_stage_switches(rb, ref):
    """Header switches: no chroma deblock, plane-wise restoration types, CDEF presets that skip."""
    s = framegen.generate(200, 136, 8, seed=11)
    s.hdr.lf_level_u = s.hdr.lf_level_v = 0
    s.hdr.lr_type[1] = 0
    for i in range(8):
        s.hdr.cdef_uv_strength[i] = 0 if i & 1 else s.hdr.cdef_uv_strength[i]
    _check(ref, s, D | Cd | L, framegen.recon_input_planes(s))


@pytest.mark.gpu
def test_config2_1080p_8bit(rb, ref):
    """BASELINE.json configs[1]: itx + 8-tap MC + deblock on a 1080p 8-bit frame."""
    s = framegen.generate(1920, 1080, 8, seed=2)
    _check(ref, s, R | D)


@pytest.mark.gpu
def test_config3_4k_10bit(rb, ref):
    """BASELINE.json configs[2]: full recon + deblock + CDEF + Wiener/SGR on a 4K 10-bit frame."""
    s = framegen.generate(3840, 2160, 10, seed=3)
    _check(ref, s, R | D | Cd | L)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,bpc,inter_frac", [(96, 64, 8, 0.0), (160, 128, 10, 0.0), (192, 128, 12, 0.3), (256, 192, 8, 0.5),
                                               (320, 256, 10, 0.15)])
def test_intra_blocks_wavefront(rb, ref, w, h, bpc, inter_frac):
    """Intra-predicted transform blocks reconstructed level by level on the device (edge preparation, all 14 modes,
    angle deltas, filter-intra, availability flags) against the reference's decode-order loop
    (rav1d_prepare_intra_edges + intra_pred + itxfm_add per block); also next to inter blocks."""
    s = framegen.generate_intra(w, h, bpc, seed=w + bpc, inter_frac=inter_frac)
    assert len(s.intra_counts) > 10 and len(set(s.intra_items["mode"])) >= 12      # (every mode: see test_intra_modes_all_present)
    assert ((s.intra_items["plane"] > 0) & (s.intra_items["mode"] == 13)).sum() > 3          # chroma-from-luma blocks
    assert (s.intra_items["mode"] == 14).sum() >= 2 and (s.intra_items["mode"] == 15).sum() >= 1   # palette blocks
    if inter_frac >= 0.3:                                                                      # inter-intra: mask and wedge kinds
        ii = s.intra_items[(s.intra_items["flags"] & 64) != 0]
        assert (ii["angle"] < 0).any() and (ii["angle"] >= 0).any()
    INTRA = rb.STAGE_INTRA
    a = framecheck.oracle_frame(ref, s, R)
    b = framecheck.product_frame(s, R | INTRA)
    framecheck.assert_planes_equal(a, b, f"intra {w}x{h}@{bpc}")


@pytest.mark.gpu
def test_intra_wavefront_per_level_launches(rb):
    """The alternative schedule of the same wavefront (one launch per level with programmatic dependent launch,
    RB200_INTRA_LEVEL_LAUNCHES=1; the default is one cooperative launch) stays bit-exact.  The switch is read once per
    process, hence the child process."""
    import os
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, 'tests'); import refharness, framecheck\n"
            "from rav1d_b200 import lib; from rav1d_b200.synth import framegen\n"
            "lib.check(lib.init(0)); ref = refharness.load()\n"
            "for w, h, bpc, fr in ((96, 64, 8, 0.0), (192, 128, 10, 0.3)):\n"
            "    s = framegen.generate_intra(w, h, bpc, seed=w + bpc, inter_frac=fr)\n"
            "    a = framecheck.oracle_frame(ref, s, lib.STAGE_RECON)\n"
            "    d = framegen.DeviceFrame(s); d.load_batch(); d.set_ref_from_host(s.ref)\n"
            "    d.submit(lib.STAGE_RECON | lib.STAGE_INTRA, 1); d.wait()\n"
            "    assert lib.frame_last_launches(d.h) > len(s.intra_counts) // 2, 'expected one launch per level'\n"
            "    b = framecheck.visible(s, d.readback()); d.close()\n"
            "    framecheck.assert_planes_equal(a, b, 'per-level intra')\n"
            "print('per-level ok')\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=dict(os.environ, RB200_INTRA_LEVEL_LAUNCHES="1"),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "per-level ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("layout", [0, 1, 2, 3])
def test_intra_levels_host_helper(layout):
    """rb200_intra_assign_levels (host code) reproduces the generator's wavefront levels from the items in decode order,
    and its level-sorted order is the one the frames are submitted in."""
    import ctypes as C
    from rav1d_b200 import lib
    s = framegen.generate_intra(192, 160, 10, seed=90 + layout, inter_frac=0.3, layout=layout, ibc_frac=0.2 * (layout & 1))
    items = s.intra_items_decode.copy()
    want = items["level"].copy()
    items["level"] = 0xffff
    n = len(items)
    order = np.zeros(n, np.int32); counts = np.zeros(4096, np.int32); nl = C.c_int()
    g = s.geom
    lib.check(lib.intra_assign_levels(items.ctypes.data, n, g.bw, g.bh, g.ss_hor, g.ss_ver, order.ctypes.data, counts.ctypes.data, 4096, C.byref(nl)))
    assert np.array_equal(items["level"], want)
    assert nl.value == len(s.intra_counts) and np.array_equal(counts[:nl.value], s.intra_counts)
    assert np.array_equal(items[order], s.intra_items)
    assert lib.intra_assign_levels(items.ctypes.data, n, g.bw, g.bh, g.ss_hor, g.ss_ver, order.ctypes.data, counts.ctypes.data, 3, C.byref(nl)) != 0


def test_intra_modes_all_present():
    """The frames of test_intra_blocks_wavefront together exercise every coded mode."""
    modes = set()
    for w, h, bpc, fr in [(96, 64, 8, 0.0), (160, 128, 10, 0.0), (192, 128, 12, 0.3)]:
        modes |= set(int(m) for m in framegen.generate_intra(w, h, bpc, seed=w + bpc, inter_frac=fr).intra_items["mode"])
    assert modes == set(range(16))


@pytest.mark.gpu
@pytest.mark.parametrize("layout", [0, 2, 3])
def test_intra_blocks_other_layouts(rb, ref, layout):
    """The intra wavefront in 4:0:0, 4:2:2 and 4:4:4 (chroma block and transform sizes follow the sub-sampling)."""
    s = framegen.generate_intra(160, 128, 10, seed=40 + layout, inter_frac=0.25, layout=layout)
    a = framecheck.oracle_frame(ref, s, R)
    b = framecheck.product_frame(s, R | rb.STAGE_INTRA)
    framecheck.assert_planes_equal(a, b, f"intra layout {layout}")


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,bpc,layout", [(192, 160, 8, 1), (160, 128, 10, 1), (192, 128, 12, 2), (160, 96, 10, 3), (96, 96, 8, 0)])
def test_intra_block_copy(rb, ref, w, h, bpc, layout):
    """Intra block copy (src/recon.rs:3196-3240: bilinear mc() from the picture being reconstructed, whole luma pixels, half
    chroma pixels where chroma is sub-sampled, source windows that leave the picture) as items of the intra wavefront,
    one level above whatever wrote the source area, with copies of copies and residuals on top."""
    s = framegen.generate_intra(w, h, bpc, seed=w + bpc + layout, inter_frac=0.15, layout=layout, ibc_frac=0.35)
    ibc = s.intra_items[s.intra_items["mode"] == 16]
    assert len(ibc) > 10 and ibc["level"].max() > 3
    if layout == 1:
        assert set(int(a) & 0xff for a in ibc["angle"]) == {0, 8, 128, 136}
    a = framecheck.oracle_frame(ref, s, R)
    b = framecheck.product_frame(s, R | rb.STAGE_INTRA)
    framecheck.assert_planes_equal(a, b, f"intra block copy {w}x{h}@{bpc} layout {layout}")


@pytest.mark.gpu
@pytest.mark.parametrize("bpc", [8, 10])
def test_empty_batch_and_idle_filters(rb, bpc):
    """No work items and every filter switched off in the header: each stage is a no-op, the picture comes back
    untouched, and nothing reads the (empty) lists."""
    import ctypes as C
    s = framegen.generate(96, 80, bpc, seed=9)
    start = framegen.recon_input_planes(s)
    hdr = rb.FrameHeader.from_buffer_copy(bytes(s.hdr))
    hdr.lf_level_y[0] = hdr.lf_level_y[1] = 0
    for i in range(8):
        hdr.cdef_y_strength[i] = hdr.cdef_uv_strength[i] = 0
    for p in range(3):
        hdr.lr_type[p] = rb.RESTORATION_NONE
    f = C.c_void_p()
    rb.check(rb.frame_create(C.byref(f), C.byref(hdr), 0, 0, 0))
    try:
        data, strides = framegen._plane_args(start)
        rb.check(rb.frame_upload_planes(f, 0, data, strides))
        for stages in (R, R | D | Cd | L):
            rb.check(rb.frame_submit(f, 0, (C.c_int32 * 19)(), 0, stages, 1))
            rb.check(rb.frame_wait(f))
            assert rb.frame_last_launches(f) <= 3          # at most the CDEF pass-through
            out = [np.zeros_like(p) for p in start]
            data2, strides2 = framegen._plane_args(out)
            rb.check(rb.frame_readback(f, data2, strides2))
            framecheck.assert_planes_equal(framecheck.visible(s, start), framecheck.visible(s, out), f"idle stages {stages}")
        # more items than the frame was created for is an error, not a crash
        assert rb.frame_submit(f, 0, (C.c_int32 * 19)(), 5, R, 1) != 0
        assert b"batch larger" in rb.last_error()
    finally:
        rb.frame_destroy(f)


@pytest.mark.gpu
def test_config5_4k_10bit(rb, ref):
    """BASELINE.json configs[4] (one GPU's stream): 50 % compound, 5 % warped, 10 % OBMC blocks, all
    post-filters and film grain on a 4K 10-bit frame."""
    import ctypes as C
    import refharness
    s = framegen.generate(3840, 2160, 10, seed=5, comp_frac=0.5, warp_frac=0.05, obmc_frac=0.1)
    fg = framegen.random_film_grain(np.random.default_rng(5), lag=3, overlap=1)
    cur, rf, rf2 = (refharness.RefFrame(ref, s, n) for n in (8, 1, 1))
    try:
        rf.set_planes(s.ref); rf2.set_planes(s.ref2)
        cur.load_filter_meta()
        cur.recon(rf, n_threads=8, ref_frame2=rf2)
        cur.filter(D | Cd | L, n_threads=8)
        exp = framecheck.visible(s, cur.apply_grain(fg, 0))
    finally:
        for f in (cur, rf, rf2):
            f.close()
    dev = framegen.DeviceFrame(s)
    try:
        dev.load_batch(); dev.set_ref_from_host(s.ref); dev.set_ref_slot(1, s.ref2)
        rb.check(rb.frame_set_film_grain(dev.h, C.byref(fg), 0))
        dev.submit(R | D | Cd | L | rb.STAGE_FILM_GRAIN); dev.wait()
        got = framecheck.visible(s, dev.readback())
    finally:
        dev.close()
    framecheck.assert_planes_equal(exp, got, "config 5")


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,bpc", [(200, 120, 8), (640, 360, 10)])
def test_zero_copy_coefficients(rb, ref, w, h, bpc):
    """RB200_UPLOAD_ZERO_COPY_COEF: the itx kernels read the pinned coefficient staging directly,
    bounded by Rb200ItxItem.ncols; RB200_UPLOAD_GATHER_COEF: a gather kernel pulls the same columns into
    the device mirror first; and ncols = 0 (unknown) still reads whole blocks."""
    s = framegen.generate(w, h, bpc, seed=21)
    a = framecheck.oracle_frame(ref, s, R | D)
    framecheck.assert_planes_equal(a, framecheck.product_frame(s, R | D, upload=2), "zero-copy")
    framecheck.assert_planes_equal(a, framecheck.product_frame(s, R | D, upload=3), "gather")
    s.itx_items["ncols"] = 0
    framecheck.assert_planes_equal(a, framecheck.product_frame(s, R | D, upload=2), "zero-copy, ncols unknown")
    framecheck.assert_planes_equal(a, framecheck.product_frame(s, R | D, upload=3), "gather, ncols unknown")
    framecheck.assert_planes_equal(a, framecheck.product_frame(s, R | D, upload=1), "copy, ncols unknown")


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,bpc", [(640, 360, 10), (200, 120, 12)])
def test_int16_coefficient_transport(rb, ref, w, h, bpc):
    """RB200_UPLOAD_GATHER_COEF16: the coefficients of a 16-bit picture cross PCIe as int16; those that do not fit travel
    as {index, value} escapes (forced here: real residuals of that size are rare) and are patched in on the device."""
    s = framegen.generate(w, h, bpc, seed=23)
    a = framecheck.oracle_frame(ref, s, R | D)
    framecheck.assert_planes_equal(a, framecheck.product_frame(s, R | D, upload=rb.UPLOAD_GATHER_COEF16), "int16 transport")
    framecheck.assert_planes_equal(a, framecheck.product_frame(s, R | D, upload=rb.UPLOAD_PACKED_COEF16), "packed int16 stream")
    # large coefficients: every 97th non-zero one pushed to the edge of the legal range (src/recon.rs:1417)
    lim = 128 << bpc
    nz = np.flatnonzero(s.coef)[::97]
    s.coef = s.coef.copy()
    s.coef[nz] = np.where(s.coef[nz] > 0, lim - 1 - (nz % 1000), -lim + (nz % 1000)).astype(s.coef.dtype)
    assert (np.abs(s.coef.astype(np.int64)) > 32767).sum() > 50
    a = framecheck.oracle_frame(ref, s, R | D)
    framecheck.assert_planes_equal(a, framecheck.product_frame(s, R | D, upload=1), "large coefficients, copy")
    framecheck.assert_planes_equal(a, framecheck.product_frame(s, R | D, upload=rb.UPLOAD_GATHER_COEF16), "large coefficients, int16 + escapes")
    framecheck.assert_planes_equal(a, framecheck.product_frame(s, R | D, upload=rb.UPLOAD_PACKED_COEF16), "large coefficients, packed stream + escapes")
    s.itx_items = s.itx_items.copy(); s.itx_items["ncols"] = 0
    framecheck.assert_planes_equal(a, framecheck.product_frame(s, R | D, upload=rb.UPLOAD_PACKED_COEF16), "packed stream, ncols unknown")


@pytest.mark.gpu
def test_int16_transport_is_refused_for_8bit(rb):
    s = framegen.generate(64, 64, 8, seed=3)
    d = framegen.DeviceFrame(s)
    try:
        d.load_batch(); d.set_ref_from_host(s.ref)
        assert not rb.frame_coef16_buffer(d.h)
        assert rb.frame_pack_coef16(d.h, s.n_coefs) != 0
        counts = (C.c_int32 * 19)(*[int(c) for c in s.itx_counts])
        assert rb.frame_submit(d.h, s.n_coefs, counts, len(s.mc_items), R, rb.UPLOAD_GATHER_COEF16) != 0
    finally:
        d.close()


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,bpc", [(424, 300, 10), (200, 120, 8), (264, 136, 12)])
def test_luma_and_chroma_reconstruction_chains(rb, ref, w, h, bpc):
    """rb200_frame_set_plane_counts: lists sorted luma first, reconstruction and post-filters as two chains on two streams;
    several submits of the same context back to back (the chains join at the end of every submit)."""
    s = framegen.generate(w, h, bpc, seed=31)
    a = framecheck.oracle_frame(ref, s, 15)
    assert framegen.sort_luma_first(s)
    d = framegen.DeviceFrame(s)
    try:
        d.load_batch(); d.set_ref_from_host(s.ref)
        for it in range(3):
            d.submit(15, upload=it == 0)
        d.wait()
        framecheck.assert_planes_equal(a, framecheck.visible(s, d.readback()), "plane chains")
        rb.check(rb.frame_set_plane_streams(d.h, 0))
        d.submit(15, upload=False); d.wait()
        framecheck.assert_planes_equal(a, framecheck.visible(s, d.readback()), "one stream, sorted lists")
    finally:
        d.close()


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,bpc", [(424, 300, 10), (200, 136, 8), (264, 200, 12)])
def test_compound_blocks_with_global_motion_warp(rb, ref, w, h, bpc):
    """GLOBALMV_GLOBALMV compound blocks: one or both predictions are the reference's global-motion warp (warp8x8t with
    frame_hdr.gmv, src/recon.rs:3253-3268,3352-3369), luma and chroma, under every compound type."""
    s = framegen.generate(w, h, bpc, seed=41, comp_frac=0.6, gmv_frac=0.7)
    assert (s.comp_items["warp_mask"] != 0).sum() > 10 and len({int(x) for x in s.comp_items["warp_mask"]}) >= 3
    a = framecheck.oracle_frame(ref, s, R)
    framecheck.assert_planes_equal(a, framecheck.product_frame(s, R), "compound with warped predictions")


@pytest.mark.gpu
def test_validate_names_bad_records(rb):
    """rb200_frame_validate: the staged batch as rb200_frame_submit will read it; a bad record is an error code, not a fault."""
    s = framegen.generate(200, 120, 10, seed=9, obmc_frac=0.2)
    d = framegen.DeviceFrame(s)
    try:
        d.load_batch(); d.set_ref_from_host(s.ref)
        counts = (C.c_int32 * 19)(*[int(c) for c in s.itx_counts])
        args = (d.h, s.n_coefs, counts, len(s.mc_items), R)
        assert rb.frame_validate(*args) == 0, rb.last_error()
        mc = rb.np_view(rb.frame_mc_items(d.h), rb.MC_ITEM_DT, len(s.mc_items))
        itx = rb.np_view(rb.frame_itx_items(d.h), rb.ITX_ITEM_DT, len(s.itx_items))
        for arr, field, bad, word in ((mc, "ref", 5, b"reference slot"), (mc, "dst_x", 4000, b"leaves plane"), (mc, "filter2d", 11, b"filter"),
                                      (itx, "cf_off", s.n_coefs, b"coefficients"), (itx, "txtp", 17, b"type"), (itx, "x", 3000, b"leaves plane")):
            keep = arr[field][3].copy()
            arr[field][3] = bad
            assert rb.frame_validate(*args) != 0 and word in rb.last_error(), (field, rb.last_error())
            arr[field][3] = keep
        assert rb.frame_validate(*args) == 0
    finally:
        d.close()


@pytest.mark.gpu
def test_validate_names_bad_intra_records(rb):
    """rb200_frame_validate over the intra wavefront's records: mode, block size, origin and level."""
    s = framegen.generate_intra(160, 128, 10, seed=3, inter_frac=0.2, ibc_frac=0.2)
    d = framegen.DeviceFrame(s)
    try:
        d.load_batch(); d.set_ref_from_host(s.ref)
        counts = (C.c_int32 * 19)(*[int(c) for c in s.itx_counts])
        args = (d.h, s.n_coefs, counts, len(s.mc_items), R | rb.STAGE_INTRA)
        assert rb.frame_validate(*args) == 0, rb.last_error()
        items = rb.np_view(rb.frame_intra_items(d.h), rb.INTRA_ITEM_DT, len(s.intra_items))
        for field, bad, word in (("mode", 17, b"mode"), ("tw4", 0, b"size"), ("th4", 17, b"size"), ("x4", 4000, b"outside plane"),
                                 ("plane", 3, b"plane"), ("level", 0xffff, b"level")):
            keep = items[field][5].copy()
            items[field][5] = bad
            assert rb.frame_validate(*args) != 0 and word in rb.last_error(), (field, rb.last_error())
            items[field][5] = keep
        assert rb.frame_validate(*args) == 0
    finally:
        d.close()


@pytest.mark.gpu
def test_resubmit_is_idempotent(rb, ref):
    """Submitting the same batch twice gives the same picture (recon overwrites, filters are
    out of place or restart from recon)."""
    from rav1d_b200.synth.framegen import DeviceFrame
    s = framegen.generate(200, 120, 10, seed=5)
    d = DeviceFrame(s)
    try:
        d.load_batch(); d.set_ref_from_host(s.ref)
        d.submit(15); d.wait(); a = framecheck.visible(s, d.readback())
        d.submit(15, upload=False); d.wait(); b = framecheck.visible(s, d.readback())
        framecheck.assert_planes_equal(a, b, "resubmit")
    finally:
        d.close()
