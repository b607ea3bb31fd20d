"""Runs the batch-path decoder over every stream of the reference's conformance manifests
(/root/reference/tests/dav1d-test-data/**/meson.build) and compares the decoded-frame MD5.
usage: sweep_conformance.py [cpucheck|gpu] [substring filter] -> prints a summary and writes /tmp/dbg/sweep_<mode>.json"""
import json, os, re, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = "/root/reference/tests/dav1d-test-data"

def manifest():
    out = []
    for top in ("8-bit", "10-bit", "12-bit"):
        for dirpath, _, files in os.walk(os.path.join(DATA, top)):
            if "meson.build" not in files: continue
            txt = open(os.path.join(dirpath, "meson.build")).read()
            for m in re.finditer(r"files\('([^']+)'\),\s*'([0-9a-f]{32})'", txt):
                p = os.path.join(dirpath, m.group(1))
                if os.path.exists(p): out.append((os.path.relpath(p, DATA), m.group(2), []))
            # tests with extra arguments: args: dav1d_test_args + ['-i', files('x'), '--filmgrain', '1', '--verify', 'md5']
            for m in re.finditer(r"files\('([^']+)'\),((?:\s*'[^']*',)*?)\s*'--verify',\s*'([0-9a-f]{32})'", txt):
                p = os.path.join(dirpath, m.group(1))
                extra = re.findall(r"'([^']*)'", m.group(2))
                if os.path.exists(p): out.append((os.path.relpath(p, DATA), m.group(3), extra))
    return out

def run(exe, rel, md5, extra):
    cmd = [os.path.join(ROOT, "oracle", "_ref", exe), "-q", "-i", os.path.join(DATA, rel), "--muxer", "md5", "-o", "-"] + extra
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    except subprocess.TimeoutExpired:
        return rel, "timeout", ""
    got = r.stdout.split()[0] if r.stdout.split() else ""
    if r.returncode == 3: return rel, "unsupported", r.stderr.strip().split("UNSUPPORTED:")[-1].strip()
    if r.returncode != 0: return rel, "error", f"rc {r.returncode} {r.stderr.strip()[-200:]}"
    return rel, "ok" if got == md5 else "mismatch", got

if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "cpucheck"
    filt = sys.argv[2] if len(sys.argv) > 2 else ""
    exe = "dav1d_b200" if mode == "gpu" else "dav1d_b200_cpucheck"
    ents = [e for e in manifest() if filt in e[0]]
    with ThreadPoolExecutor(max_workers=int(os.environ.get("JOBS", "4"))) as ex:
        res = list(ex.map(lambda e: run(exe, *e), ents))
    counts = {}
    for rel, st, info in res:
        counts[st] = counts.get(st, 0) + 1
        if st != "ok": print(f"{st:12s} {rel}  {info}")
    print(counts, "of", len(res))
    os.makedirs("/tmp/dbg", exist_ok=True)
    json.dump(res, open(f"/tmp/dbg/sweep_{mode}.json", "w"), indent=0)
