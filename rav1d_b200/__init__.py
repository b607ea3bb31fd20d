"""rav1d_b200: B200 (sm_100a) implementation of rav1d's reconstruction and
post-filter DSP behind the reference's own DSP function-pointer surface.

The product is the CUDA library ``librav1d_b200.so`` (C ABI declared in
``include/rav1d_b200.h``).  This package is the thin Python host side used by
tests and ``bench.py``: a ctypes binding (``rav1d_b200.lib``), the synthetic
frame generators (``rav1d_b200.synth``) and the band-split helper
(``rav1d_b200.multigpu``).  The host layer that replaces the reference's pass 2
is C, in ``rav1d_b200/host/``.  There is no CPU fallback: importing
``rav1d_b200.lib`` raises if the library is missing.
"""
__all__ = ["lib", "multigpu", "synth"]
