"""Runs oracle/_ref/dav1d_b200 (GPU) or dav1d_b200_cpucheck over tests/golden/conformance and prints / saves a summary.
usage: run_conformance.py [gpu|cpucheck] [substring] [--jobs N] [--out file.json]"""
import json, os, subprocess, sys, time
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, "tests", "golden", "conformance")

def run(exe, ent):
    cmd = [os.path.join(ROOT, "oracle", "_ref", exe), "-q", "-i", os.path.join(FIX, ent["path"]), "--muxer", "md5", "-o", "-"] + ent["args"]
    t0 = time.time()
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, RB200_HOST_STATS="1"))
    except subprocess.TimeoutExpired:
        return {"path": ent["path"], "status": "timeout"}
    got = r.stdout.split()[0] if r.stdout.split() else ""
    st = "ok" if r.returncode == 0 and got == ent["md5"] else ("unsupported" if r.returncode == 3 else "mismatch" if r.returncode == 0 else "error")
    return {"path": ent["path"], "status": st, "rc": r.returncode, "got": got, "stderr": r.stderr.strip()[-400:], "s": round(time.time() - t0, 2)}

if __name__ == "__main__":
    args = sys.argv[1:]
    jobs = int(args[args.index("--jobs") + 1]) if "--jobs" in args else 6
    out = args[args.index("--out") + 1] if "--out" in args else None
    pos = [a for i, a in enumerate(args) if not a.startswith("--") and (i == 0 or args[i - 1] not in ("--jobs", "--out"))]
    mode = pos[0] if pos else "gpu"
    filt = pos[1] if len(pos) > 1 else ""
    exe = "dav1d_b200" if mode == "gpu" else "dav1d_b200_cpucheck"
    ents = [e for e in json.load(open(os.path.join(FIX, "manifest.json")))["streams"] if filt in e["path"]]
    t0 = time.time()
    with ThreadPoolExecutor(max_workers=jobs) as ex:
        res = list(ex.map(lambda e: run(exe, e), ents))
    counts = {}
    for r in res:
        counts[r["status"]] = counts.get(r["status"], 0) + 1
    bad = [r for r in res if r["status"] != "ok"]
    for r in bad[:40]: print(r["status"], r["path"], r.get("rc"), r.get("stderr", "")[-200:].replace("\n", " | "))
    print(counts, "of", len(res), f"in {time.time() - t0:.1f} s")
    if out: json.dump({"counts": counts, "results": res}, open(out, "w"), indent=0)
