"""GPU-box debug aid: deblock-only parity on the real-stream fixtures, product vs the oracle's frame harness."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import refharness, streamdump
from rav1d_b200 import lib
from rav1d_b200.synth import framegen
lib.check(lib.init(0))
ref = refharness.load()
for key, s in streamdump.load_golden():
    if not s.stages & 2: continue
    cur = refharness.RefFrame(ref, s, 1)
    cur.load_filter_meta(); cur.set_planes(s.pre); cur.filter(2)
    exp = streamdump.visible(s, cur.get_planes()); cur.close()
    d = framegen.DeviceFrame(s); d.load_batch(); d.upload(0, s.pre); d.submit(2); d.wait()
    got = streamdump.visible(s, d.readback()); d.close()
    pre = streamdump.visible(s, s.pre)
    for p, (a, b) in enumerate(zip(exp, got)):
        bad = np.argwhere(a != b)
        if len(bad):
            print(key, "plane", p, len(bad), "px:", [(int(x), int(y), int(a[y, x]), int(b[y, x]), int(pre[p][y, x])) for y, x in bad[:24]], "(x, y, expected, got, before)")
print("done")
