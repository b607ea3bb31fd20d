"""Debug aid: decode a stream with the reference CLI and with the batch-path CLI (CPU checker by default), compare
the output frame by frame and report where they first differ.  usage: cmp_stream.py stream.ivf [gpu] [extra CLI args]"""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
def run(exe, path, out, extra):
    cmd = [os.path.join(ROOT, "oracle", "_ref", exe), "-q", "-i", path, "--muxer", "yuv4mpeg2", "-o", out] + extra
    if exe == "dav1d_ref": cmd += ["--threads", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return r.returncode, r.stderr
def frames(path):
    data = open(path, "rb").read()
    hdr_end = data.index(b"\n")
    hdr = data[:hdr_end].decode().split()
    w = int([t for t in hdr if t[0] == "W"][0][1:]); h = int([t for t in hdr if t[0] == "H"][0][1:])
    cs = [t for t in hdr if t[0] == "C"]
    cs = cs[0][1:] if cs else "420"
    hbd = any(s in cs for s in ("p10", "p12"))
    px = 2 if hbd else 1
    if cs.startswith("mono"): dims = [(w, h)]
    elif cs.startswith("420"): dims = [(w, h), ((w + 1) // 2, (h + 1) // 2)] * 1 + [((w + 1) // 2, (h + 1) // 2)]
    elif cs.startswith("422"): dims = [(w, h), ((w + 1) // 2, h), ((w + 1) // 2, h)]
    else: dims = [(w, h)] * 3
    pos = hdr_end + 1
    out = []
    while pos < len(data):
        pos = data.index(b"\n", pos) + 1
        planes = []
        for (pw, ph) in dims:
            n = pw * ph * px
            planes.append(np.frombuffer(data[pos:pos + n], dtype=np.uint16 if hbd else np.uint8).reshape(ph, pw))
            pos += n
        out.append(planes)
    return out
if __name__ == "__main__":
    path = sys.argv[1]
    exe = "dav1d_b200" if len(sys.argv) > 2 and sys.argv[2] == "gpu" else "dav1d_b200_cpucheck"
    extra = [a for a in sys.argv[2:] if a != "gpu"]
    rc0, e0 = run("dav1d_ref", path, "/tmp/dbg/ref.y4m", extra)
    rc1, e1 = run(exe, path, "/tmp/dbg/new.y4m", extra)
    print("ref rc", rc0, e0.strip()[-200:], "| new rc", rc1, e1.strip()[-300:])
    a, b = frames("/tmp/dbg/ref.y4m"), frames("/tmp/dbg/new.y4m")
    print(len(a), "vs", len(b), "frames")
    for i, (fa, fb) in enumerate(zip(a, b)):
        for p, (pa, pb) in enumerate(zip(fa, fb)):
            if pa.shape != pb.shape: print("frame", i, "plane", p, "shape", pa.shape, pb.shape); sys.exit(1)
            d = np.argwhere(pa != pb)
            if len(d):
                y, x = d[0]
                print(f"frame {i} plane {p}: {len(d)} px differ, first at x={x} y={y} (ref {pa[y, x]} new {pb[y, x]}); bbox x {d[:,1].min()}..{d[:,1].max()} y {d[:,0].min()}..{d[:,0].max()}")
                sys.exit(1)
    print("identical")
