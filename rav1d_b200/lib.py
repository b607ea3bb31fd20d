"""ctypes binding of librav1d_b200.so (C ABI: include/rav1d_b200.h).

Fails loudly when the CUDA extension has not been built: there is no CPU path.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librav1d_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or `make -C rav1d_b200/csrc`). rav1d_b200 has no CPU fallback.")

cdll = C.CDLL(LIB_PATH)

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
                ("tx", C.c_uint8), ("txtp", C.c_uint8), ("flags", C.c_uint8), ("eob", C.c_int16),
                ("pad", C.c_int16)]


ITXFM_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_int, C.c_int)


class InvTxfmDSPContext(C.Structure):
    _fields_ = [("itxfm_add", (ITXFM_FN * N_TX_TYPES_PLUS_LL) * N_RECT_TX_SIZES)]


def _sig(name, restype, *argtypes):
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
