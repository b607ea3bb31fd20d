"""Builds tests/golden/conformance/: the reference's own conformance bitstreams (inputs) and the decoded-frame MD5 its
manifests hold for them (tests/dav1d-test-data/**/meson.build; extra CLI arguments such as --filmgrain 1 included).
A stream is taken if it is at most MAX_BYTES long and the batch path can express every block of it (the CPU checker
oracle/_ref/dav1d_b200_cpucheck decodes it to the manifest's MD5 here).  Run in the build container (needs /root/reference).
usage: python tools/make_conformance_fixtures.py"""
import json, os, shutil, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from concurrent.futures import ThreadPoolExecutor
import sweep_conformance as sw

MAX_BYTES = 100000
OUT = os.path.join(sw.ROOT, "tests", "golden", "conformance")

if __name__ == "__main__":
    ents = [e for e in sw.manifest() if os.path.getsize(os.path.join(sw.DATA, e[0])) <= MAX_BYTES]
    with ThreadPoolExecutor(max_workers=4) as ex:
        res = list(ex.map(lambda e: sw.run("dav1d_b200_cpucheck", *e), ents))
    if os.path.isdir(OUT): shutil.rmtree(OUT)
    man, skipped = [], {}
    for (rel, md5, extra), (_, st, info) in zip(ents, res):
        if st != "ok":
            skipped[rel] = f"{st}: {info}"
            continue
        dst = os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(sw.DATA, rel), dst)
        man.append({"path": rel, "md5": md5, "args": extra})
    json.dump({"source": "tests/dav1d-test-data/**/meson.build of the reference", "streams": man, "not_expressible": skipped},
              open(os.path.join(OUT, "manifest.json"), "w"), indent=1)
    print(len(man), "streams,", sum(os.path.getsize(os.path.join(OUT, m["path"])) for m in man), "bytes;", len(skipped), "skipped")
