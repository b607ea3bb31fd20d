"""Frame-level parity: the product's stream-ordered stage launches (rb200_frame_*, C ABI)
against the reference's own frame drivers (dav1d_filter_sbrow_* and the DSP tables, run by
oracle/ref_frame.c) on the same synthetic frame batch -- stage by stage, then end to end."""
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
    """Half of the blocks predicted from two references (avg / w_avg / segmentation mask), BASELINE config 5."""
    s = framegen.generate(w, h, bpc, seed=w + 1, comp_frac=0.5)
    assert len(s.comp_items) > 10 and set(s.comp_items["comp_type"]) == {0, 1, 2}
    _check(ref, s, R)
    _check(ref, s, R | D | Cd | L)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,bpc", [(176, 144, 8), (208, 128, 10), (256, 192, 12)])
def test_warped_blocks(rb, ref, w, h, bpc):
    """Affine-warped blocks next to translational and compound ones, all planes."""
    s = framegen.generate(w, h, bpc, seed=w + 2, comp_frac=0.3, warp_frac=0.3)
    assert len(s.warp_items) > 5
    _check(ref, s, R)
    _check(ref, s, R | D | Cd | L)


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
@pytest.mark.parametrize("w,h,bpc", [(200, 120, 8), (640, 360, 10)])
def test_zero_copy_coefficients(rb, ref, w, h, bpc):
    """RB200_UPLOAD_ZERO_COPY_COEF: the itx kernels read the pinned coefficient staging directly,
    bounded by Rb200ItxItem.ncols; and ncols = 0 (unknown) still reads whole blocks."""
    s = framegen.generate(w, h, bpc, seed=21)
    a = framecheck.oracle_frame(ref, s, R | D)
    framecheck.assert_planes_equal(a, framecheck.product_frame(s, R | D, upload=2), "zero-copy")
    s.itx_items["ncols"] = 0
    framecheck.assert_planes_equal(a, framecheck.product_frame(s, R | D, upload=2), "zero-copy, ncols unknown")
    framecheck.assert_planes_equal(a, framecheck.product_frame(s, R | D, upload=1), "copy, ncols unknown")


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
