"""Shared helpers: run one synthetic frame through the oracle (reference drivers) and
through the product (CUDA, C ABI) and compare visible planes."""
import numpy as np

import refharness


def oracle_frame(ref, s, stages, n_tc=1, start_planes=None):
    """Returns the visible planes after `stages`. start_planes: pre-filter picture when RECON is not in stages."""
    from rav1d_b200 import lib
    cur = refharness.RefFrame(ref, s, n_tc)
    rf = refharness.RefFrame(ref, s, 1)
    rf2 = refharness.RefFrame(ref, s, 1) if hasattr(s, "ref2") else None
    rf3 = None
    if hasattr(s, "ref3"):      # a reference of another size: its own header
        import copy
        s3 = copy.copy(s)
        s3.hdr = lib.FrameHeader.from_buffer_copy(bytes(s.hdr))
        s3.hdr.width, s3.hdr.height = s.scaled_ref_size
        s3.w, s3.h = s.scaled_ref_size
        s3.aw, s3.ah = (s3.w + 127) & ~127, (s3.h + 127) & ~127
        rf3 = refharness.RefFrame(ref, s3, 1)
    try:
        rf.set_planes(s.ref)
        if rf2:
            rf2.set_planes(s.ref2)
        if rf3:
            rf3.set_planes(s.ref3)
        cur.load_filter_meta()
        if stages & lib.STAGE_RECON:
            cur.recon(rf, n_threads=n_tc, ref_frame2=rf2, ref_frame3=rf3)
        else:
            cur.set_planes(start_planes)
        if stages & ~lib.STAGE_RECON:
            cur.filter(stages, n_threads=n_tc)
        return visible(s, cur.get_planes())
    finally:
        cur.close()
        rf.close()
        if rf2:
            rf2.close()
        if rf3:
            rf3.close()


def product_frame(s, stages, start_planes=None, upload=1):
    from rav1d_b200 import lib
    from rav1d_b200.synth.framegen import DeviceFrame
    d = DeviceFrame(s)
    try:
        d.load_batch()
        if stages & lib.STAGE_RECON:
            d.set_ref_from_host(s.ref)
            if hasattr(s, "ref2"):
                d.set_ref_slot(1, s.ref2)
            if hasattr(s, "ref3"):
                d.set_ref_slot(2, s.ref3, size=s.scaled_ref_size)
        else:
            d.upload(0, start_planes)
        if upload == lib.UPLOAD_GATHER_COEF16:      # what a front end does when it fills the staging: int16 + escapes
            lib.check(lib.frame_pack_coef16(d.h, s.n_coefs), "frame_pack_coef16")
        if upload == lib.UPLOAD_PACKED_COEF16:      # ... or one contiguous int16 stream of the blocks' leading columns
            import ctypes as C
            counts = (C.c_int32 * 19)(*[int(c) for c in s.itx_counts])
            lib.check(lib.frame_pack_coef_stream(d.h, s.n_coefs, counts, stages), "frame_pack_coef_stream")
        d.submit(stages, upload)
        d.wait()
        return visible(s, d.readback())
    finally:
        d.close()


def visible(s, planes):
    layout = getattr(s, "layout", 1)
    ssx, ssy = int(layout != 3), int(layout == 1)
    return [planes[0][:s.h, :s.w]] + [planes[p][:(s.h + ssy) >> ssy, :(s.w + ssx) >> ssx] for p in range(1, len(planes))]


def assert_planes_equal(a, b, what=""):
    assert len(a) == len(b)
    for p in range(len(a)):
        if not np.array_equal(a[p], b[p]):
            bad = np.argwhere(a[p] != b[p])
            y, x = bad[0]
            raise AssertionError(f"{what}: plane {p} differs at {len(bad)} px, first (x={x}, y={y}): "
                                 f"oracle {a[p][y, x]} vs product {b[p][y, x]}; rows {np.unique(bad[:, 0])[:12]}")
