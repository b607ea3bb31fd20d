#!/usr/bin/env python3
"""Where the time of the one-launch intra wavefront goes.  RB200_INTRA_TRACE=<file> makes intra_levels_kernel stamp, per
item, when it saw its level released and when its own release was done; this script runs one 4K key frame with the
trace on and prints the per-level period split into hand-off (last release of level l -> first item of level l + 1
released from its wait) and critical section (wait seen -> own release), by coded mode.
    python tools/intra_trace.py [w h bpc]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
path = "/tmp/rb200_intra_trace.bin"
os.environ["RB200_INTRA_TRACE"] = path
from rav1d_b200 import lib  # noqa: E402
from rav1d_b200.synth import framegen  # noqa: E402

w, h, bpc = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (3840, 2176, 10)
lib.check(lib.init(0))
s = framegen.generate_intra(w, h, bpc, seed=1, inter_frac=0.0)
d = framegen.DeviceFrame(s)
d.load_batch(); d.set_ref_from_host(s.ref)
for i in range(2):
    d.submit(lib.STAGE_RECON | lib.STAGE_INTRA, 1 if i == 0 else 0); d.wait()
d.close()
t = np.fromfile(path, np.uint64).reshape(-1, 2).astype(np.int64)
it = s.intra_items
assert len(t) == len(it), (len(t), len(it))
lv = it["level"].astype(np.int64)
n_levels = int(lv.max()) + 1
seen, rel = t[:, 0], t[:, 1]
ok = lv > 0
crit = (rel - seen)[ok]
print(f"{len(it)} items, {n_levels} levels; kernel span {(rel.max() - rel[rel > 0].min()) / 1e6:.2f} ms")
print(f"critical section (wait seen -> own release), all items: median {np.median(crit):.0f} ns, p90 {np.percentile(crit, 90):.0f}, p99 {np.percentile(crit, 99):.0f}, max {crit.max()}")
first_seen = np.full(n_levels, np.iinfo(np.int64).max); last_seen = np.zeros(n_levels, np.int64); last_rel = np.zeros(n_levels, np.int64)
np.minimum.at(first_seen, lv[ok], seen[ok]); np.maximum.at(last_seen, lv[ok], seen[ok]); np.maximum.at(last_rel, lv, rel)
L = np.arange(2, n_levels)
period = last_rel[L] - last_rel[L - 1]
handoff = first_seen[L] - last_rel[L - 1]
spread = last_seen[L] - first_seen[L]
slowest = np.zeros(n_levels, np.int64); np.maximum.at(slowest, lv[ok], (rel - seen)[ok])
print(f"per level: period median {np.median(period):.0f} ns (mean {period.mean():.0f}); hand-off last release -> first seen median {np.median(handoff):.0f}; "
      f"first -> last item to see it {np.median(spread):.0f}; slowest critical section median {np.median(slowest[L]):.0f}")
key = it["mode"].astype(np.int64) * 100 + it["tw4"].astype(np.int64) * 4
for k in np.unique(key[ok]):
    m = ok & (key == k)
    c = (rel - seen)[m]
    print(f"  mode {k // 100:2d} size {k % 100:2d}: {m.sum():6d} items, critical section median {np.median(c):5.0f} ns, p90 {np.percentile(c, 90):5.0f}")
