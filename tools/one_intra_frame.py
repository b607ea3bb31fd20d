#!/usr/bin/env python3
"""One key frame (every block intra) through RB200_STAGE_RECON | RB200_STAGE_INTRA, for an ncu capture of the wavefront
kernel.   python tools/one_intra_frame.py [w h bpc]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rav1d_b200 import lib  # noqa: E402
from rav1d_b200.synth import framegen  # noqa: E402

w, h, bpc = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (3840, 2176, 10)
lib.check(lib.init(0))
s = framegen.generate_intra(w, h, bpc, seed=1, inter_frac=0.0)
d = framegen.DeviceFrame(s)
d.load_batch(); d.set_ref_from_host(s.ref)
for i in range(2):
    d.submit(lib.STAGE_RECON | lib.STAGE_INTRA, 1 if i == 0 else 0); d.wait()
print(len(s.intra_items), "items in", len(s.intra_counts), "levels")
d.close()
