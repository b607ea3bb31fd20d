"""ctypes access to the ORACLE: oracle/_ref/libdav1d_ref.so, the reference's
own portable C DSP compiled in place from /root/reference by oracle/Makefile
(plus oracle/ref_harness.c glue).  Test infrastructure only."""
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
        vp, i, ss = C.c_void_p, C.c_int, C.c_ssize_t
        _lib.ref_itxfm_add.argtypes = [i, i, vp, ss, vp, i, i]
        _lib.ref_itxfm_add.restype = None
        _lib.ref_itx_has.argtypes = [i, i]
        _lib.ref_itx_has.restype = i
        _lib.ref_itxfm_add_many.argtypes = [i, vp, vp, ss, vp, i]
        _lib.ref_itxfm_add_many.restype = None
    return _lib


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)
