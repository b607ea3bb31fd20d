#!/usr/bin/env python3
"""Time the reconstruction of frames with intra blocks (RB200_STAGE_RECON | RB200_STAGE_INTRA) on the device and in the
reference's decode-order loop on the CPU.   python tools/time_intra.py [w h bpc]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from rav1d_b200 import lib  # noqa: E402
from rav1d_b200.synth import framegen  # noqa: E402

w, h, bpc = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (3840, 2176, 10)
lib.check(lib.init(0))
try:
    import refharness
    ref = refharness.load()
except Exception:      # no oracle on this box: device timing only
    ref = None
for inter_frac in (0.9, 0.5, 0.0):
    s = framegen.generate_intra(w, h, bpc, seed=1, inter_frac=inter_frac)
    d = framegen.DeviceFrame(s)
    d.load_batch(); d.set_ref_from_host(s.ref)
    stages = lib.STAGE_RECON | lib.STAGE_INTRA
    d.submit(stages, 1); d.wait()
    n = 5
    t0 = time.perf_counter()
    for _ in range(n):
        d.submit(stages, 0)
    d.wait()
    gpu_ms = (time.perf_counter() - t0) / n * 1e3
    launches = lib.frame_last_launches(d.h)
    d.close()
    cpu_ms = float("nan")
    if ref is not None:
        cur, rf = refharness.RefFrame(ref, s, 1), refharness.RefFrame(ref, s, 1)
        rf.set_planes(s.ref); cur.load_filter_meta()
        t0 = time.perf_counter(); cur.recon(rf); cpu_ms = (time.perf_counter() - t0) * 1e3
        cur.close(); rf.close()
    print(f"{w}x{h}@{bpc} intra blocks {100 - inter_frac * 100:.0f} %: {len(s.intra_items)} transform blocks in {len(s.intra_counts)} levels, "
          f"{launches} launches, device {gpu_ms:.2f} ms / frame, reference loop (1 thread) {cpu_ms:.1f} ms")
