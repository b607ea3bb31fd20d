"""ctypes access to the ORACLE: oracle/_ref/libdav1d_ref.so, the reference's
own portable C DSP and frame drivers compiled in place from /root/reference by
oracle/Makefile (plus oracle/ref_harness.c and oracle/ref_frame.c glue).
Test infrastructure only."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libdav1d_ref.so")

_lib = None


def available():
    return os.path.exists(REF_SO)


def load():
    global _lib
    if _lib is None:
        if not available():
            import pytest
            pytest.skip("oracle/_ref/libdav1d_ref.so not built (run `make -C oracle` where /root/reference exists)")
        _lib = C.CDLL(REF_SO)
        _lib.ref_init()
        vp, i, ss, sz = C.c_void_p, C.c_int, C.c_ssize_t, C.c_size_t

        def sig(name, res, *args):
            f = getattr(_lib, name)
            f.restype = res
            f.argtypes = list(args)

        sig("ref_itxfm_add", None, i, i, vp, ss, vp, i, i)
        sig("ref_itx_has", i, i, i)
        sig("ref_itxfm_add_many", None, i, vp, vp, ss, vp, i)
        sig("ref_mc", None, i, vp, ss, vp, ss, i, i, i, i, i)
        sig("ref_mct", None, i, vp, vp, ss, i, i, i, i, i)
        sig("ref_mc_scaled", None, i, vp, ss, vp, ss, i, i, i, i, i, i, i)
        sig("ref_mct_scaled", None, i, vp, vp, ss, i, i, i, i, i, i, i)
        sig("ref_avg", None, vp, ss, vp, vp, i, i, i)
        sig("ref_w_avg", None, vp, ss, vp, vp, i, i, i, i)
        sig("ref_mask", None, vp, ss, vp, vp, i, i, vp, i)
        sig("ref_w_mask", None, i, vp, ss, vp, vp, i, i, vp, i, i)
        sig("ref_blend", None, i, vp, ss, vp, i, i, vp, i)
        sig("ref_warp8x8", None, vp, ss, vp, ss, vp, i, i, i)
        sig("ref_warp8x8t", None, vp, ss, vp, ss, vp, i, i, i)
        sig("ref_emu_edge", None, ss, ss, ss, ss, ss, ss, vp, ss, vp, ss, i)
        sig("ref_resize", None, vp, ss, vp, ss, i, i, i, i, i, i)
        sig("ref_lpf_sb", None, i, i, vp, ss, vp, vp, ss, vp, i, i)
        sig("ref_cdef_dir", i, vp, ss, C.POINTER(C.c_uint), i)
        sig("ref_cdef_fb", None, i, vp, ss, vp, vp, vp, i, i, i, i, i, i)
        sig("ref_lr", None, i, vp, ss, vp, vp, i, i, vp, i, i)
        sig("ref_calc_eih", None, vp, i)
        sig("ref_sizeof", sz, i)
        sig("ref_frame_new", vp, vp, i)
        sig("ref_lf_build", i, vp, i, i, i, i, i, i, i, i, vp, vp)
        sig("ref_frame_free", None, vp)
        sig("ref_frame_plane", vp, vp, i)
        sig("ref_frame_stride", ss, vp, i)
        sig("ref_frame_masks", vp, vp)
        sig("ref_frame_levels", vp, vp)
        sig("ref_frame_lut", vp, vp)
        sig("ref_frame_lr_masks", vp, vp)
        sig("ref_frame_sbh", i, vp)
        sig("ref_frame_filter", None, vp, i, i)
        sig("ref_frame_recon", None, vp, vp, i, vp, i, vp, i, vp, i)
        sig("ref_frame_recon_comp", None, vp, vp, i, vp, i, i)
        sig("ref_frame_recon_warp", None, vp, vp, i, vp, i, i)
        sig("ref_frame_recon_obmc", None, vp, vp, i, vp, i)
        sig("ref_wedge_mask", vp, i, i, i, i, i)
        sig("ref_frame_recon_scaled", None, vp, vp, i, vp, i)
        sig("ref_frame_recon_intra", None, vp, vp, i, vp, vp, vp, vp)
        sig("ref_frame_apply_grain", None, vp, vp, i)
        sig("ref_frame_set_gmv", None, vp, i, vp, vp)
        sig("ref_frame_grain_plane", vp, vp, i)
        sig("ref_fg_gen_y", None, vp, vp, i)
        sig("ref_fg_gen_uv", None, i, vp, vp, vp, ss, i)
        sig("ref_fgy", None, vp, vp, ss, vp, sz, vp, vp, i, i, i)
        sig("ref_fguv", None, i, vp, vp, ss, vp, sz, vp, vp, i, i, vp, ss, i, i, i)
        sig("ref_sizeof_film_grain_data", sz)
    return _lib


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class RefFrame:
    """One picture inside the reference's own Dav1dFrameContext (oracle/ref_frame.c)."""

    def __init__(self, ref, s, n_tc=1):
        self.ref, self.s = ref, s
        self.h = ref.ref_frame_new(C.addressof(s.hdr), n_tc)
        self.px = 2 if s.bpc > 8 else 1
        self.strides = [ref.ref_frame_stride(self.h, 0), ref.ref_frame_stride(self.h, 1)]

    def close(self):
        if self.h:
            self.ref.ref_frame_free(self.h)
            self.h = None

    def plane_view(self, p, grain=False):
        s = self.s
        layout = s.hdr.layout
        ss_ver, ss_hor = int(layout == 1), int(layout != 3)
        rows = s.ah if p == 0 else s.ah >> ss_ver
        stride = self.strides[1 if p else 0]
        addr = self.ref.ref_frame_grain_plane(self.h, p) if grain else self.ref.ref_frame_plane(self.h, p)
        buf = (C.c_ubyte * (stride * rows)).from_address(addr)
        a = np.frombuffer(buf, dtype=np.uint8).reshape(rows, stride)
        cols = s.aw if p == 0 else s.aw >> ss_hor
        return a[:, :cols * self.px].view(np.uint16 if self.px == 2 else np.uint8)

    def n_planes(self):
        return 1 if self.s.hdr.layout == 0 else 3

    def set_planes(self, planes):
        for p in range(self.n_planes()):
            self.plane_view(p)[:] = planes[p]

    def get_planes(self):
        return [self.plane_view(p).copy() for p in range(self.n_planes())]

    def load_filter_meta(self):
        from rav1d_b200 import lib
        s, ref = self.s, self.ref
        g = s.geom
        n = g.sb128w * g.sb128h
        assert ref.ref_sizeof(0) == lib.AV1_FILTER_DT.itemsize and ref.ref_sizeof(1) == lib.AV1_RESTORATION_DT.itemsize
        assert ref.ref_sizeof(2) == C.sizeof(lib.FilterLUT)
        lib.np_view(ref.ref_frame_masks(self.h), lib.AV1_FILTER_DT, n)[:] = s.masks
        lib.np_view(ref.ref_frame_levels(self.h), np.uint8, s.levels.size)[:] = s.levels.reshape(-1)
        C.memmove(ref.ref_frame_lut(self.h), C.byref(s.lut), C.sizeof(lib.FilterLUT))
        lib.np_view(ref.ref_frame_lr_masks(self.h), lib.AV1_RESTORATION_DT, n)[:] = s.lr_masks

    def recon(self, ref_frame, n_threads=1, coef_work=None, ref_frame2=None, ref_frame3=None):
        """Prediction (put items, then compound blocks) and residual, in the reference's DSP calls."""
        s = self.s
        cw = s.coef.copy() if coef_work is None else coef_work
        frames = [ref_frame] + ([ref_frame2] if ref_frame2 is not None else [])
        if ref_frame3 is not None:       # slot 2: the scaled reference
            frames = (frames + [ref_frame])[:2] + [ref_frame3]
        refs = (C.c_void_p * len(frames))(*[f.h for f in frames])
        mc = np.ascontiguousarray(s.mc_items)
        itx = np.ascontiguousarray(s.itx_items)
        comp = np.ascontiguousarray(getattr(s, "comp_items", np.zeros(0, np.uint8)))
        if hasattr(s, "gmv_matrix"):
            for slot in range(7):
                self.ref.ref_frame_set_gmv(self.h, slot, ptr(np.ascontiguousarray(s.gmv_matrix[slot])), ptr(np.ascontiguousarray(s.gmv_abcd[slot])))
        warp = np.ascontiguousarray(getattr(s, "warp_items", np.zeros(0, np.uint8)))
        if len(warp):
            self.ref.ref_frame_recon_warp(self.h, refs, len(frames), ptr(warp), len(warp), n_threads)
        obmc = np.ascontiguousarray(getattr(s, "obmc_items", np.zeros(0, np.uint8)))
        self.ref.ref_frame_recon(self.h, refs, len(frames), ptr(mc), len(mc), ptr(itx), 0, ptr(cw), n_threads)
        scaled = np.ascontiguousarray(getattr(s, "scaled_items", np.zeros(0, np.uint8)))
        if len(scaled):
            self.ref.ref_frame_recon_scaled(self.h, refs, len(frames), ptr(scaled), len(scaled))
        if len(comp):
            self.ref.ref_frame_recon_comp(self.h, refs, len(frames), ptr(comp), len(comp), n_threads)
        if len(obmc):
            self.ref.ref_frame_recon_obmc(self.h, refs, len(frames), ptr(obmc), len(obmc))
        n_inter_itx = int(np.sum(s.itx_counts))      # the residuals of intra blocks follow in the same list
        self.ref.ref_frame_recon(self.h, refs, len(frames), ptr(mc), 0, ptr(itx), n_inter_itx, ptr(cw), n_threads)
        intra = getattr(s, "intra_items_decode", None)
        if intra is not None and len(intra):
            intra = np.ascontiguousarray(intra); of = np.ascontiguousarray(s.intra_itx_of)
            pal = np.ascontiguousarray(getattr(s, "palette", np.zeros(16, np.uint8)))
            self.ref.ref_frame_recon_intra(self.h, ptr(intra), len(intra), ptr(of), ptr(itx), ptr(cw), ptr(pal))
        return cw

    def apply_grain(self, fg, is_id=0):
        self.ref.ref_frame_apply_grain(self.h, C.addressof(fg), is_id)
        return [self.plane_view(p, grain=True).copy() for p in range(3)]

    def apply_grain_inplace(self, fg, is_id=0):
        """As apply_grain, without copying the grained planes out (bench timing)."""
        self.ref.ref_frame_apply_grain(self.h, C.addressof(fg), is_id)

    def filter(self, stages, n_threads=1):
        self.ref.ref_frame_filter(self.h, stages, n_threads)
