"""ctypes binding of librav1d_b200.so (C ABI: include/rav1d_b200.h).

Fails loudly when the CUDA extension has not been built: there is no CPU path.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librav1d_b200.so")

# RB200_LIB_TYPES_ONLY=1: only the record types / numpy dtypes / constants of this module are wanted (bench.py's
# --impl reference arm builds the synthetic frame with them and must not map the product library); every function
# then raises when called.
TYPES_ONLY = os.environ.get("RB200_LIB_TYPES_ONLY") == "1"

if not TYPES_ONLY and not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or `make -C rav1d_b200/csrc`). rav1d_b200 has no CPU fallback.")

cdll = None if TYPES_ONLY else C.CDLL(LIB_PATH)

N_RECT_TX_SIZES = 19
N_TX_TYPES_PLUS_LL = 17

TX_NAMES = ["TX_4X4", "TX_8X8", "TX_16X16", "TX_32X32", "TX_64X64", "RTX_4X8", "RTX_8X4", "RTX_8X16",
            "RTX_16X8", "RTX_16X32", "RTX_32X16", "RTX_32X64", "RTX_64X32", "RTX_4X16", "RTX_16X4",
            "RTX_8X32", "RTX_32X8", "RTX_16X64", "RTX_64X16"]
TX_DIMS = [(4, 4), (8, 8), (16, 16), (32, 32), (64, 64), (4, 8), (8, 4), (8, 16), (16, 8), (16, 32), (32, 16),
           (32, 64), (64, 32), (4, 16), (16, 4), (8, 32), (32, 8), (16, 64), (64, 16)]
TXTP_NAMES = ["DCT_DCT", "ADST_DCT", "DCT_ADST", "ADST_ADST", "FLIPADST_DCT", "DCT_FLIPADST",
              "FLIPADST_FLIPADST", "ADST_FLIPADST", "FLIPADST_ADST", "IDTX", "V_DCT", "H_DCT", "V_ADST",
              "H_ADST", "V_FLIPADST", "H_FLIPADST", "WHT_WHT"]


class Planes(C.Structure):
    _fields_ = [("data", C.c_void_p * 3), ("stride", C.c_int64 * 3)]


class ItxItem(C.Structure):
    _fields_ = [("cf_off", C.c_uint32), ("x", C.c_uint16), ("y", C.c_uint16), ("plane", C.c_uint8),
                ("tx", C.c_uint8), ("txtp", C.c_uint8), ("ncols", C.c_uint8), ("eob", C.c_int16),
                ("pad", C.c_int16)]


ITXFM_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_int, C.c_int)


class InvTxfmDSPContext(C.Structure):
    _fields_ = [("itxfm_add", (ITXFM_FN * N_TX_TYPES_PLUS_LL) * N_RECT_TX_SIZES)]


def _sig(name, restype, *argtypes):
    if cdll is None:
        def unavailable(*a, **k):
            raise RuntimeError(f"{name}: rav1d_b200.lib was imported with RB200_LIB_TYPES_ONLY=1 (no library mapped)")
        return unavailable
    f = getattr(cdll, name)
    f.restype = restype
    f.argtypes = list(argtypes)
    return f


abi_version = _sig("rb200_abi_version", C.c_int)
init = _sig("rb200_init", C.c_int, C.c_int)
last_error = _sig("rb200_last_error", C.c_char_p)
malloc = _sig("rb200_malloc", C.c_int, C.POINTER(C.c_void_p), C.c_size_t)
free = _sig("rb200_free", C.c_int, C.c_void_p)
malloc_host = _sig("rb200_malloc_host", C.c_int, C.POINTER(C.c_void_p), C.c_size_t)
free_host = _sig("rb200_free_host", C.c_int, C.c_void_p)
memcpy_h2d = _sig("rb200_memcpy_h2d", C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)
memcpy_d2h = _sig("rb200_memcpy_d2h", C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)
memset = _sig("rb200_memset", C.c_int, C.c_void_p, C.c_int, C.c_size_t, C.c_void_p)
stream_sync = _sig("rb200_stream_sync", C.c_int, C.c_void_p)

itx_dsp_init = _sig("rb200_itx_dsp_init", None, C.POINTER(InvTxfmDSPContext), C.c_int)
itxfm_add = _sig("rb200_itxfm_add", C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_int,
                 C.c_int)
itx_valid = _sig("rb200_itx_valid", C.c_int, C.c_int, C.c_int)
itx_add_batch = _sig("rb200_itx_add_batch", C.c_int, C.POINTER(Planes), C.c_void_p, C.c_void_p,
                     C.POINTER(C.c_int32), C.c_int, C.c_void_p)


class Rb200Error(RuntimeError):
    pass


def check(rc, what=""):
    if rc != 0:
        raise Rb200Error(f"{what or 'rav1d_b200'} failed ({rc}): {last_error().decode(errors='replace')}")
    return rc


# ---------------------------------------------------------------- mc / filters / frame
import numpy as _np

N_2D_FILTERS = 10
FILTER_2D_NAMES = ["8TAP_REGULAR", "8TAP_REGULAR_SMOOTH", "8TAP_REGULAR_SHARP", "8TAP_SHARP_REGULAR",
                   "8TAP_SHARP_SMOOTH", "8TAP_SHARP", "8TAP_SMOOTH_REGULAR", "8TAP_SMOOTH", "8TAP_SMOOTH_SHARP",
                   "BILINEAR"]
STAGE_RECON, STAGE_DEBLOCK, STAGE_CDEF, STAGE_LR, STAGE_FILM_GRAIN, STAGE_SUPER_RES, STAGE_INTRA = 1, 2, 4, 8, 16, 32, 64
STAGE_ALL = 15
LAYOUT_I400, LAYOUT_I420, LAYOUT_I422, LAYOUT_I444 = 0, 1, 2, 3
RESTORATION_NONE, RESTORATION_SWITCHABLE, RESTORATION_WIENER, RESTORATION_SGRPROJ = 0, 1, 2, 3

# numpy mirrors of the C records (include/rav1d_b200.h)
MC_ITEM_DT = _np.dtype([("dst_x", "<i2"), ("dst_y", "<i2"), ("src_x", "<i2"), ("src_y", "<i2"), ("w", "u1"),
                        ("h", "u1"), ("plane", "u1"), ("ref", "u1"), ("mx", "u1"), ("my", "u1"),
                        ("filter2d", "u1"), ("flags", "u1")])
ITX_ITEM_DT = _np.dtype([("cf_off", "<u4"), ("x", "<u2"), ("y", "<u2"), ("plane", "u1"), ("tx", "u1"),
                         ("txtp", "u1"), ("ncols", "u1"), ("eob", "<i2"), ("pad", "<i2")])
AV1_FILTER_DT = _np.dtype([("filter_y", "<u2", (2, 32, 3, 2)), ("filter_uv", "<u2", (2, 32, 2, 2)),
                           ("cdef_idx", "i1", (4,)), ("noskip_mask", "<u2", (16, 2))])
LR_UNIT_DT = _np.dtype([("type", "u1"), ("filter_h", "i1", (3,)), ("filter_v", "i1", (3,)),
                        ("sgr_weights", "i1", (2,))])
AV1_RESTORATION_DT = _np.dtype([("lr", LR_UNIT_DT, (3, 4))])
COMP_ITEM_DT = _np.dtype([("x", "<i2"), ("y", "<i2"), ("w", "u1"), ("h", "u1"), ("ref", "u1", (2,)), ("mv", "<i2", (2, 2)),
                          ("filter2d", "u1"), ("comp_type", "u1"), ("jnt_weight", "u1"), ("mask_sign", "u1"),
                          ("wedge_idx", "u1"), ("warp_mask", "u1"), ("pad", "u1", (10,))])
SCALED_ITEM_DT = _np.dtype([("dst_x", "<i2"), ("dst_y", "<i2"), ("w", "u1"), ("h", "u1"), ("plane", "u1"), ("ref", "u1"),
                            ("pos_x", "<i4"), ("pos_y", "<i4"), ("step_x", "<i4"), ("step_y", "<i4"), ("filter2d", "u1"),
                            ("flags", "u1"), ("pad", "u1", (6,))])
assert SCALED_ITEM_DT.itemsize == 32
INTRA_ITEM_DT = _np.dtype([("x4", "<u2"), ("y4", "<u2"), ("w4_end", "<u2"), ("h4_end", "<u2"), ("plane", "u1"), ("tw4", "u1"),
                           ("th4", "u1"), ("mode", "u1"), ("angle", "i1"), ("flags", "u1"), ("level", "<u2")])
assert INTRA_ITEM_DT.itemsize == 16
WARP_ITEM_DT = _np.dtype([("x", "<i2"), ("y", "<i2"), ("w", "u1"), ("h", "u1"), ("ref", "u1"), ("pad0", "u1"),
                          ("matrix", "<i4", (6,)), ("abcd", "<i2", (4,)), ("pad", "u1", (8,))])
assert WARP_ITEM_DT.itemsize == 48
LF_BLOCK_DT = _np.dtype([("bx", "<u2"), ("by", "<u2"), ("bs", "u1"), ("flags", "u1"), ("ytx", "u1"), ("uvtx", "u1"),
                         ("tx_split", "<u2", (2,)), ("lvl", "u1", (4,))])
assert LF_BLOCK_DT.itemsize == 16
LFB_INTRA, LFB_SKIP, LFB_HAS_CHROMA = 1, 2, 4
COMP_AVG, COMP_WEIGHTED_AVG, COMP_SEG, COMP_WEDGE = 0, 1, 2, 3
MC_PUT, MC_OBMC_ABOVE, MC_OBMC_LEFT = 0, 1, 2
assert MC_ITEM_DT.itemsize == 16 and ITX_ITEM_DT.itemsize == 16 and COMP_ITEM_DT.itemsize == 32
assert AV1_FILTER_DT.itemsize == 1348 and AV1_RESTORATION_DT.itemsize == 108


class FilterLUT(C.Structure):
    _fields_ = [("e", C.c_uint8 * 64), ("i", C.c_uint8 * 64), ("sharp", C.c_uint64 * 2)]


class LrParams(C.Union):
    class _Sgr(C.Structure):
        _fields_ = [("s0", C.c_uint32), ("s1", C.c_uint32), ("w0", C.c_int16), ("w1", C.c_int16)]
    _fields_ = [("filter", (C.c_int16 * 8) * 2), ("sgr", _Sgr), ("_align", C.c_byte * 32)]


class FilmGrainData(C.Structure):
    """Rb200FilmGrainData == Dav1dFilmGrainData (include/dav1d/headers.rs:1610-1661)"""
    _fields_ = [("seed", C.c_uint), ("num_y_points", C.c_int), ("y_points", (C.c_uint8 * 2) * 14),
                ("chroma_scaling_from_luma", C.c_int), ("num_uv_points", C.c_int * 2),
                ("uv_points", ((C.c_uint8 * 2) * 10) * 2), ("scaling_shift", C.c_int), ("ar_coeff_lag", C.c_int),
                ("ar_coeffs_y", C.c_int8 * 24), ("ar_coeffs_uv", (C.c_int8 * 28) * 2), ("ar_coeff_shift", C.c_uint64),
                ("grain_scale_shift", C.c_int), ("uv_mult", C.c_int * 2), ("uv_luma_mult", C.c_int * 2),
                ("uv_offset", C.c_int * 2), ("overlap_flag", C.c_int), ("clip_to_restricted_range", C.c_int)]


GRAIN_WIDTH, GRAIN_HEIGHT = 82, 73


class FrameHeader(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("bpc", C.c_int32), ("layout", C.c_int32),
                ("sb128", C.c_int32), ("lf_level_y", C.c_int32 * 2), ("lf_level_u", C.c_int32),
                ("lf_level_v", C.c_int32), ("cdef_damping", C.c_int32), ("cdef_y_strength", C.c_int32 * 8),
                ("cdef_uv_strength", C.c_int32 * 8), ("lr_type", C.c_int32 * 3),
                ("lr_unit_size_log2", C.c_int32 * 2), ("upscaled_width", C.c_int32)]


class FrameGeometry(C.Structure):
    _fields_ = [("bw", C.c_int32), ("bh", C.c_int32), ("w4", C.c_int32), ("h4", C.c_int32),
                ("sb128w", C.c_int32), ("sb128h", C.c_int32), ("sbh", C.c_int32), ("b4_stride", C.c_int32),
                ("ss_hor", C.c_int32), ("ss_ver", C.c_int32), ("stride", C.c_int64 * 2),
                ("plane_h", C.c_int32 * 2), ("n_planes", C.c_int32)]


_vp, _i, _ss, _sz = C.c_void_p, C.c_int, C.c_ssize_t, C.c_size_t
mc = _sig("rb200_mc", _i, _i, _vp, _ss, _vp, _ss, _i, _i, _i, _i, _i)
mct = _sig("rb200_mct", _i, _i, _vp, _vp, _ss, _i, _i, _i, _i, _i)
mc_scaled = _sig("rb200_mc_scaled", _i, _i, _vp, _ss, _vp, _ss, _i, _i, _i, _i, _i, _i, _i)
mct_scaled = _sig("rb200_mct_scaled", _i, _i, _vp, _vp, _ss, _i, _i, _i, _i, _i, _i, _i)
avg = _sig("rb200_avg", _i, _vp, _ss, _vp, _vp, _i, _i, _i)
w_avg = _sig("rb200_w_avg", _i, _vp, _ss, _vp, _vp, _i, _i, _i, _i)
mask = _sig("rb200_mask", _i, _vp, _ss, _vp, _vp, _i, _i, _vp, _i)
w_mask = _sig("rb200_w_mask", _i, _i, _vp, _ss, _vp, _vp, _i, _i, _vp, _i, _i)
blend = _sig("rb200_blend", _i, _i, _vp, _ss, _vp, _i, _i, _vp, _i)
warp8x8 = _sig("rb200_warp8x8", _i, _vp, _ss, _vp, _ss, _vp, _i, _i, _i)
warp8x8t = _sig("rb200_warp8x8t", _i, _vp, _ss, _vp, _ss, _vp, _i, _i, _i)
emu_edge = _sig("rb200_emu_edge", _i, _ss, _ss, _ss, _ss, _ss, _ss, _vp, _ss, _vp, _ss, _i)
resize = _sig("rb200_resize", _i, _vp, _ss, _vp, _ss, _i, _i, _i, _i, _i, _i)
mc_batch = _sig("rb200_mc_batch", _i, C.POINTER(Planes), C.POINTER(Planes), _i, _i, _i, _i, _i, _vp, _i, _i, _vp)
loop_filter_sb = _sig("rb200_loop_filter_sb", _i, _i, _i, _vp, _ss, _vp, _vp, _ss, C.POINTER(FilterLUT), _i, _i)
cdef_dir = _sig("rb200_cdef_dir", _i, _vp, _ss, C.POINTER(C.c_uint), _i, C.POINTER(C.c_int))
cdef_fb = _sig("rb200_cdef_fb", _i, _i, _vp, _ss, _vp, _vp, _vp, _i, _i, _i, _i, C.c_uint32, _i)
lr = _sig("rb200_lr", _i, _i, _vp, _ss, _vp, _vp, _i, _i, C.POINTER(LrParams), C.c_uint32, _i)
mc_dsp_init = _sig("rb200_mc_dsp_init", None, _vp, _i)
loop_filter_dsp_init = _sig("rb200_loop_filter_dsp_init", None, _vp, _i)
cdef_dsp_init = _sig("rb200_cdef_dsp_init", None, _vp, _i)
loop_restoration_dsp_init = _sig("rb200_loop_restoration_dsp_init", None, _vp, _i)

film_grain_dsp_init = _sig("rb200_film_grain_dsp_init", None, _vp, _i)
generate_grain_y = _sig("rb200_generate_grain_y", _i, _vp, C.POINTER(FilmGrainData), _i)
generate_grain_uv = _sig("rb200_generate_grain_uv", _i, _i, _vp, _vp, C.POINTER(FilmGrainData), _ss, _i)
fgy_32x32xn = _sig("rb200_fgy_32x32xn", _i, _vp, _vp, _ss, C.POINTER(FilmGrainData), _sz, _vp, _vp, _i, _i, _i)
fguv_32x32xn = _sig("rb200_fguv_32x32xn", _i, _i, _vp, _vp, _ss, C.POINTER(FilmGrainData), _sz, _vp, _vp, _i, _i,
                    _vp, _ss, _i, _i, _i)
generate_scaling = _sig("rb200_generate_scaling", _i, _i, _vp, _i, _vp)
frame_set_film_grain = _sig("rb200_frame_set_film_grain", _i, _vp, C.POINTER(FilmGrainData), _i)
frame_display_planes = _sig("rb200_frame_display_planes", _i, _vp, C.POINTER(Planes))

frame_create = _sig("rb200_frame_create", _i, C.POINTER(_vp), C.POINTER(FrameHeader), _sz, _i, _i)
frame_destroy = _sig("rb200_frame_destroy", _i, _vp)
frame_geometry = _sig("rb200_frame_geometry", _i, _vp, C.POINTER(FrameGeometry))
frame_coef_buffer = _sig("rb200_frame_coef_buffer", _vp, _vp)
frame_itx_items = _sig("rb200_frame_itx_items", _vp, _vp)
frame_mc_items = _sig("rb200_frame_mc_items", _vp, _vp)
frame_lf_masks = _sig("rb200_frame_lf_masks", _vp, _vp)
frame_lf_levels = _sig("rb200_frame_lf_levels", _vp, _vp)
frame_lf_lut = _sig("rb200_frame_lf_lut", _vp, _vp)
frame_lr_masks = _sig("rb200_frame_lr_masks", _vp, _vp)
frame_reserve_comp_items = _sig("rb200_frame_reserve_comp_items", _i, _vp, _i)
frame_comp_items = _sig("rb200_frame_comp_items", _vp, _vp)
frame_set_comp_count = _sig("rb200_frame_set_comp_count", _i, _vp, _i)
ipred = _sig("rb200_ipred", _i, _i, _vp, _ss, _vp, _i, _i, _i, _i, _i, _i)
cfl_ac = _sig("rb200_cfl_ac", _i, _i, _vp, _vp, _ss, _i, _i, _i, _i, _i)
cfl_pred = _sig("rb200_cfl_pred", _i, _i, _vp, _ss, _vp, _i, _i, _vp, _i, _i)
pal_pred = _sig("rb200_pal_pred", _i, _vp, _ss, _vp, _vp, _i, _i, _i)
(DC_PRED, VERT_PRED, HOR_PRED, LEFT_DC_PRED, TOP_DC_PRED, DC_128_PRED, Z1_PRED, Z2_PRED, Z3_PRED, SMOOTH_PRED, SMOOTH_V_PRED,
 SMOOTH_H_PRED, PAETH_PRED, FILTER_PRED) = range(14)


class IntraPredDSPContext(C.Structure):
    _fields_ = [("intra_pred", C.c_void_p * 14), ("cfl_ac", C.c_void_p * 3), ("cfl_pred", C.c_void_p * 6), ("pal_pred", C.c_void_p)]


intra_pred_dsp_init = _sig("rb200_intra_pred_dsp_init", None, C.POINTER(IntraPredDSPContext), _i)
wedge_mask = _sig("rb200_wedge_mask", _i, _i, _i, _i, _i, _i, _vp)
frame_reserve_intra_items = _sig("rb200_frame_reserve_intra_items", _i, _vp, _i, _i)
frame_intra_items = _sig("rb200_frame_intra_items", _vp, _vp)
frame_intra_itx_index = _sig("rb200_frame_intra_itx_index", _vp, _vp)
intra_assign_levels = _sig("rb200_intra_assign_levels", _i, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, C.POINTER(_i))
frame_reserve_palette = _sig("rb200_frame_reserve_palette", _i, _vp, _sz)
frame_palette_buffer = _sig("rb200_frame_palette_buffer", _vp, _vp)
frame_set_palette_bytes = _sig("rb200_frame_set_palette_bytes", _i, _vp, _sz)
frame_set_intra_levels = _sig("rb200_frame_set_intra_levels", _i, _vp, _i, C.POINTER(C.c_int32), C.POINTER(C.c_int32))
frame_reserve_scaled_items = _sig("rb200_frame_reserve_scaled_items", _i, _vp, _i)
frame_scaled_items = _sig("rb200_frame_scaled_items", _vp, _vp)
frame_set_scaled_count = _sig("rb200_frame_set_scaled_count", _i, _vp, _i)
frame_set_ref_size = _sig("rb200_frame_set_ref_size", _i, _vp, _i, _i, _i)
frame_reserve_obmc_items = _sig("rb200_frame_reserve_obmc_items", _i, _vp, _i)
frame_obmc_items = _sig("rb200_frame_obmc_items", _vp, _vp)
frame_set_obmc_counts = _sig("rb200_frame_set_obmc_counts", _i, _vp, _i, _i)
frame_reserve_warp_items = _sig("rb200_frame_reserve_warp_items", _i, _vp, _i)
frame_warp_items = _sig("rb200_frame_warp_items", _vp, _vp)
frame_set_warp_count = _sig("rb200_frame_set_warp_count", _i, _vp, _i)
frame_set_ref = _sig("rb200_frame_set_ref", _i, _vp, _i, C.POINTER(Planes))
frame_upload_planes = _sig("rb200_frame_upload_planes", _i, _vp, _i, C.POINTER(_vp), C.POINTER(_ss))
frame_output_planes = _sig("rb200_frame_output_planes", _i, _vp, C.POINTER(Planes))
frame_stage_planes = _sig("rb200_frame_stage_planes", _i, _vp, _i, C.POINTER(Planes))
frame_submit = _sig("rb200_frame_submit", _i, _vp, _sz, C.POINTER(C.c_int32), _i, _i, _i)
frame_wait = _sig("rb200_frame_wait", _i, _vp)
frame_readback = _sig("rb200_frame_readback", _i, _vp, C.POINTER(_vp), C.POINTER(_ss))
frame_readback_async = _sig("rb200_frame_readback_async", _i, _vp, C.POINTER(_vp), C.POINTER(_ss))
frame_stream = _sig("rb200_frame_stream", _vp, _vp)
frame_set_band = _sig("rb200_frame_set_band", _i, _vp, _i, _i)
frame_band_rows = _sig("rb200_frame_band_rows", _i, _vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i))
frame_upload_rows = _sig("rb200_frame_upload_rows", _i, _vp, _i, C.POINTER(_vp), C.POINTER(_ss), _i, _i)
frame_readback_rows = _sig("rb200_frame_readback_rows", _i, _vp, C.POINTER(_vp), C.POINTER(_ss), _i, _i)
frame_plane_block = _sig("rb200_frame_plane_block", _i, _vp, _i, C.POINTER(_vp), C.POINTER(_sz))
frame_pull_rows = _sig("rb200_frame_pull_rows", _i, _vp, _i, _vp, _i, _i)
ipc_get_handle = _sig("rb200_ipc_get_handle", _i, _vp, _vp)
ipc_open_handle = _sig("rb200_ipc_open_handle", _i, _vp, C.POINTER(_vp))
ipc_close_handle = _sig("rb200_ipc_close_handle", _i, _vp)
enable_peer_access = _sig("rb200_enable_peer_access", _i, _i)
frame_set_stream = _sig("rb200_frame_set_stream", _i, _vp, _vp)
frame_depend = _sig("rb200_frame_depend", _i, _vp, _vp)
frame_set_scaled_obmc_counts = _sig("rb200_frame_set_scaled_obmc_counts", _i, _vp, _i, _i)
frame_set_ref_gmv = _sig("rb200_frame_set_ref_gmv", _i, _vp, _i, C.POINTER(C.c_int32), C.POINTER(C.c_int16))
frame_set_plane_counts = _sig("rb200_frame_set_plane_counts", _i, _vp, _i, C.POINTER(C.c_int32))
frame_validate = _sig("rb200_frame_validate", _i, _vp, _sz, C.POINTER(C.c_int32), _i, _i)
flag_signal = _sig("rb200_flag_signal", _i, _vp, _vp, C.c_uint32)
flag_wait = _sig("rb200_flag_wait", _i, _vp, _vp, C.c_uint32)
UPLOAD_GATHER_COEF16 = 4
UPLOAD_PACKED_COEF16 = 5
frame_pack_coef_stream = _sig("rb200_frame_pack_coef_stream", _i, _vp, _sz, C.POINTER(C.c_int32), _i)
frame_coef_stream = _sig("rb200_frame_coef_stream", _vp, _vp)
frame_coef_stream_offsets = _sig("rb200_frame_coef_stream_offsets", _vp, _vp)
frame_set_coef_stream_length = _sig("rb200_frame_set_coef_stream_length", _i, _vp, _sz)
frame_coef16_buffer = _sig("rb200_frame_coef16_buffer", _vp, _vp)
frame_pack_coef16 = _sig("rb200_frame_pack_coef16", _i, _vp, _sz)
frame_reserve_coef_escapes = _sig("rb200_frame_reserve_coef_escapes", _i, _vp, _i)
frame_coef_escapes = _sig("rb200_frame_coef_escapes", _vp, _vp)
frame_set_coef_escape_count = _sig("rb200_frame_set_coef_escape_count", _i, _vp, _i)
frame_set_plane_streams = _sig("rb200_frame_set_plane_streams", _i, _vp, _i)
frame_enable_timing = _sig("rb200_frame_enable_timing", _i, _vp, _i)
frame_stage_times = _sig("rb200_frame_stage_times", _i, _vp, C.POINTER(C.c_float))
frame_last_launches = _sig("rb200_frame_last_launches", _i, _vp)
frame_reserve_lf_blocks = _sig("rb200_frame_reserve_lf_blocks", _i, _vp, _i)
frame_lf_blocks = _sig("rb200_frame_lf_blocks", _vp, _vp)
frame_set_lf_block_count = _sig("rb200_frame_set_lf_block_count", _i, _vp, _i)
frame_download_lf = _sig("rb200_frame_download_lf", _i, _vp, _vp, _vp)


def np_view(ptr, dtype, count):
    """numpy view of `count` records of `dtype` at raw address `ptr` (pinned staging owned by the library)."""
    dtype = _np.dtype(dtype)
    buf = (C.c_ubyte * (dtype.itemsize * count)).from_address(ptr)
    return _np.frombuffer(buf, dtype=dtype, count=count)
