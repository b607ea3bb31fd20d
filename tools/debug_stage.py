"""GPU-box debug aid: one small synthetic frame, stage subsets, product vs oracle.  usage: debug_stage.py [w h bpc]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import refharness, framecheck
from rav1d_b200 import lib
from rav1d_b200.synth import framegen
w, h, bpc = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (208, 128, 10)
lib.check(lib.init(0))
ref = refharness.load()
s = framegen.generate(w, h, bpc, seed=1)
for stages in (1, 3, 7, 15):
    a = framecheck.oracle_frame(ref, s, stages)
    b = framecheck.product_frame(s, stages)
    bad = [int((x != y).sum()) for x, y in zip(a, b)]
    print("stages", stages, "mismatching pixels per plane", bad, flush=True)
