#!/usr/bin/env python3
"""All streams of the reference's conformance manifests through oracle/_ref/dav1d_b200 (every frame reconstructed and
filtered on the GPU), including the ones too large to be committed as fixtures.  The GPU box has no reference tree, so:

    python tools/conformance_all_gpu.py stage      # here: copies the streams + manifest into _allstreams/ (git-ignored)
    gpurun -- python tools/conformance_all_gpu.py  # on the box: decodes, compares the MD5s, writes gpurun_out/conformance_all_gpu.txt
    rm -r _allstreams

profiles/r02j_conformance_all_gpu.txt is the result of such a run (789 of 789)."""
import collections
import json
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ALL = os.path.join(ROOT, "_allstreams")


def stage():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sweep_conformance as sc
    ents = sc.manifest()
    for rel, _, _ in ents:
        dst = os.path.join(ALL, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst):
            shutil.copyfile(os.path.join(sc.DATA, rel), dst)
    json.dump(ents, open(os.path.join(ALL, "manifest.json"), "w"))
    print(len(ents), "streams staged in", ALL)


def run(e):
    rel, md5, extra = e
    cmd = [os.path.join(ROOT, "oracle", "_ref", "dav1d_b200"), "-q", "-i", os.path.join(ALL, rel), "--muxer", "md5", "-o", "-"] + extra
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    except subprocess.TimeoutExpired:
        return rel, "timeout", ""
    got = r.stdout.split()[0] if r.stdout.split() else ""
    if r.returncode == 3:
        return rel, "unsupported", r.stderr.strip()[-200:]
    if r.returncode != 0:
        return rel, "error", f"rc {r.returncode} {r.stderr.strip()[-200:]}"
    return rel, "ok" if got == md5 else "mismatch", got


if __name__ == "__main__":
    if sys.argv[1:] == ["stage"]:
        stage()
        sys.exit(0)
    ents = json.load(open(os.path.join(ALL, "manifest.json")))
    with ThreadPoolExecutor(max_workers=6) as ex:
        res = list(ex.map(run, ents))
    c = collections.Counter(st for _, st, _ in res)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "conformance_all_gpu.txt"), "w") as f:
        f.write(f"oracle/_ref/dav1d_b200 (every frame reconstructed and filtered on the B200) over the reference's conformance manifests: {dict(c)} of {len(res)}\n")
        for rel, st, info in res:
            f.write(f"{st} {rel} {info if st != 'ok' else ''}\n")
    print(dict(c), "of", len(res))
    sys.exit(0 if c.get("ok", 0) == len(res) else 1)
