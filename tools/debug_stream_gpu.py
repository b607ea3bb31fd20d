"""GPU-box debug aid: for each given fixture stream, decode with the reference CLI and with dav1d_b200, report the first
differing frame / plane / pixel, then decode again with RB200_HOST_DUMP at that pixel to list the records covering it.
usage: debug_stream_gpu.py <fixture relpath> ..."""
import os, subprocess, sys, hashlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import cmp_stream as cs
ROOT = cs.ROOT
FIX = os.path.join(ROOT, "tests", "golden", "conformance")
os.makedirs("/tmp/dbg", exist_ok=True)
for rel in sys.argv[1:]:
    path = os.path.join(FIX, rel)
    print("=====", rel)
    cs.run("dav1d_ref", path, "/tmp/dbg/ref.y4m", [])
    md5s = []
    for k in range(2):
        rc, err = cs.run("dav1d_b200", path, f"/tmp/dbg/new{k}.y4m", [])
        md5s.append(hashlib.md5(open(f"/tmp/dbg/new{k}.y4m", "rb").read()).hexdigest())
    print("deterministic:", md5s[0] == md5s[1], "rc", rc, err[-200:])
    a, b = cs.frames("/tmp/dbg/ref.y4m"), cs.frames("/tmp/dbg/new0.y4m")
    found = None
    for i, (fa, fb) in enumerate(zip(a, b)):
        for p, (pa, pb) in enumerate(zip(fa, fb)):
            d = np.argwhere(pa != pb)
            if len(d):
                y, x = d[0]
                cells = sorted(set((int(yy) // 4 * 4, int(xx) // 4 * 4) for yy, xx in d))
                print(f"output frame {i} plane {p}: {len(d)} px differ, first x={x} y={y} ref {pa[y, x]} new {pb[y, x]}; bbox x {d[:,1].min()}..{d[:,1].max()} y {d[:,0].min()}..{d[:,0].max()}; {len(cells)} cells, first {cells[:12]}")
                print("   row ref:", pa[y, max(x - 2, 0):x + 14].tolist()); print("   row new:", pb[y, max(x - 2, 0):x + 14].tolist())
                if found is None: found = (p, int(x), int(y))
        if found: break
    if not found:
        print("identical"); continue
    env = dict(os.environ, RB200_HOST_DUMP="%d,%d,%d" % found)
    r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "dav1d_b200"), "-q", "-i", path, "--muxer", "null", "-o", "/dev/null"], capture_output=True, text=True, env=env)
    print(r.stderr[-6000:])
