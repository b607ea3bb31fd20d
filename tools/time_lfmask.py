#!/usr/bin/env python3
"""Cost of building the loop-filter masks / levels on the device for one frame:
    python tools/time_lfmask.py [w h layout]
Times deblock-only submits of a zero picture with (a) host masks uploaded, (b) block records uploaded and the masks built
by lf_cells_kernel + lf_words_kernel, (c) the same resident (no upload), and prints the oracle's time for the same records."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from rav1d_b200 import lib  # noqa: E402
from rav1d_b200.synth import framegen  # noqa: E402
import refharness  # noqa: E402
import test_lfmask as T  # noqa: E402


def main():
    w, h, layout = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (3840, 2160, 1)
    lib.check(lib.init(0), "init")
    ref = refharness.load()
    blocks = framegen.generate_lf_blocks(w, h, layout, 1, seed=7)
    t0 = time.perf_counter()
    em, el = T.oracle_lf(ref, blocks, w, h, layout, 1)
    t_ref = time.perf_counter() - t0
    s = T._bare_frame(w, h, 10, layout, 1, blocks)
    s.masks, s.levels = em.copy(), el.copy()
    res = {}
    for name, use_blocks, upload in (("host masks, uploaded", False, 1), ("block records, uploaded", True, 1), ("block records, resident", True, 0)):
        s.lf_blocks = blocks if use_blocks else None
        d = framegen.DeviceFrame(s)
        try:
            d.load_batch()
            for _ in range(5):
                d.submit(lib.STAGE_DEBLOCK, 1)
            d.wait()
            n = 200
            t0 = time.perf_counter()
            for _ in range(n):
                d.submit(lib.STAGE_DEBLOCK, upload)
            d.wait()
            res[name] = (time.perf_counter() - t0) / n * 1e3
            if use_blocks:
                gm, gl = d.download_lf()
                T.assert_lf_equal(em, el, gm, gl, w, h, layout, name)
        finally:
            d.close()
    print(f"{w}x{h} layout {layout}: {len(blocks)} blocks ({len(blocks) * 16 / 1e6:.2f} MB of records vs "
          f"{(em.nbytes + el.nbytes) / 1e6:.2f} MB of masks + levels); reference functions on one core {t_ref * 1e3:.2f} ms")
    for k, v in res.items():
        print(f"  deblock-only submit, {k}: {v:.3f} ms / frame")


if __name__ == "__main__":
    main()
