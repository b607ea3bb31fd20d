"""GPU-box debug aid: decode the given fixture streams concurrently (N processes at a time, R rounds) and compare every
output with the reference CLI's; prints where they differ.  usage: stress_stream_gpu.py N R <relpath> ..."""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import cmp_stream as cs
FIX = os.path.join(cs.ROOT, "tests", "golden", "conformance")
os.makedirs("/tmp/dbg", exist_ok=True)
N, R = int(sys.argv[1]), int(sys.argv[2])
rels = sys.argv[3:]
refs = {}
for i, rel in enumerate(rels):
    cs.run("dav1d_ref", os.path.join(FIX, rel), f"/tmp/dbg/ref{i}.y4m", [])
    refs[rel] = cs.frames(f"/tmp/dbg/ref{i}.y4m")
def one(job):
    k, rel = job
    out = f"/tmp/dbg/new{k}.y4m"
    env = dict(os.environ)
    rc, err = cs.run("dav1d_b200", os.path.join(FIX, rel), out, [])
    b = cs.frames(out)
    a = refs[rel]
    msgs = []
    for i, (fa, fb) in enumerate(zip(a, b)):
        for p, (pa, pb) in enumerate(zip(fa, fb)):
            d = np.argwhere(pa != pb)
            if len(d):
                cells = sorted(set((int(yy) // 8 * 8, int(xx) // 8 * 8) for yy, xx in d))
                msgs.append(f"f{i} p{p}: {len(d)} px, bbox x {d[:,1].min()}..{d[:,1].max()} y {d[:,0].min()}..{d[:,0].max()}, {len(cells)} 8x8 cells {cells[:6]}, maxdiff {int(np.abs(pa.astype(int) - pb.astype(int)).max())}")
    return rel, rc, msgs
jobs = [(k, rels[k % len(rels)]) for k in range(N * R)]
with ThreadPoolExecutor(max_workers=N) as ex:
    for rel, rc, msgs in ex.map(one, jobs):
        print(rel, "rc", rc, "IDENTICAL" if not msgs else "DIFF")
        for m in msgs[:10]: print("    ", m)
