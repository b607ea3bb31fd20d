#!/usr/bin/env python3
"""One submit of a bench workload's frame (for `ncu --metrics gpu__time_duration.sum` launch lists).
    python tools/one_frame.py [workload] [n_submits]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rav1d_b200 import lib  # noqa: E402
from rav1d_b200.synth import framegen  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "4k10"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
wl = bench.WORKLOADS[name]
s = framegen.generate(wl[0], wl[1], wl[2], seed=1, **bench.GEN_ARGS.get(name, {}))
lib.check(lib.init(0))
if os.environ.get("RB200_LF_RECORDS"):      # masks and levels built on the device (lf_cells_kernel + lf_words_kernel)
    s.lf_blocks = s.lf_block_records
d = framegen.DeviceFrame(s)
d.load_batch()
if wl[3] & 1:
    d.set_ref_from_host(s.ref)
    if hasattr(s, "ref2"):
        d.set_ref_slot(1, s.ref2)
else:
    d.upload(0, framegen.recon_input_planes(s))
if wl[3] & 16:
    fg = framegen.random_film_grain(np.random.default_rng(7), lag=3, overlap=1)
    lib.check(lib.frame_set_film_grain(d.h, C.byref(fg), 0))
for i in range(n):
    d.submit(wl[3], 1 if i == 0 else 0)
    d.wait()
print("launches per frame:", lib.frame_last_launches(d.h))
d.close()
