"""Decoded-frame MD5 on the reference's conformance streams, through the batch path end to end.

The reference decoder (oracle/_ref/libdav1d_ref.so, its own C sources compiled in place) does what stays on the CPU:
OBU parsing and entropy decoding (pass 1 of its two-pass frame threading).  Pass 2 and the filter tasks are the
host layer of rav1d_b200/host/: recon_b_intra / recon_b_inter append records to the frame's batch
(src/recon.rs:2402-4045), the frame is handed to a backend when its tasks are done, pictures are read back when they
are output.  Two executables link that host layer behind the reference's CLI (tools/dav1d.c, md5 muxer):

  dav1d_b200            the product backend: every frame through rb200_frame_* on the GPU            (-m gpu)
  dav1d_b200_cpucheck   the CPU checker (oracle/ref_backend.c): the same batches executed with the
                        reference's DSP functions in the GPU's stage order                           (-m "not gpu")

Fixtures: tests/golden/conformance/ (bitstreams + manifest.json with the MD5s of the reference's meson.build
manifests), made by tools/make_conformance_fixtures.py.  Bit-exact = the MD5 of every output frame's visible pixels
(tools/output/md5.rs:541-574 hashes plane rows in order) equals the manifest.
"""
import json
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, "tests", "golden", "conformance")
BIN = os.path.join(ROOT, "oracle", "_ref")
MANIFEST = json.load(open(os.path.join(FIX, "manifest.json")))["streams"]


def _run_all(exe, ents, jobs):
    """Decode `ents` with the multi-stream runner `exe` (jobs processes, one CUDA context each); returns the lines that
    are not "ok".  Streams with extra CLI arguments all ask for --filmgrain 1; without it the md5 muxer of the reference's CLI
    decodes with film grain off (tools/dav1d_cli_parse.c:420-426), and so does the runner."""
    assert all(e["args"] in ([], ["--filmgrain", "1"]) for e in ents)
    import tempfile
    chunks = [ents[i::jobs] for i in range(jobs) if ents[i::jobs]]

    def one(chunk):
        with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
            # film grain only where the manifest asks for it: under the md5 muxer the reference's CLI leaves it off otherwise
            f.write("".join(f"{os.path.join(FIX, e['path'])} {e['md5']} {int('--filmgrain' in e['args'])}\n" for e in chunk))
        try:
            r = subprocess.run([os.path.join(BIN, exe), f.name], capture_output=True, text=True, timeout=1200)
        finally:
            os.unlink(f.name)
        lines = r.stdout.splitlines()
        assert len(lines) == len(chunk), (r.returncode, r.stdout[-400:], r.stderr[-400:])
        return lines

    with ThreadPoolExecutor(max_workers=jobs) as ex:
        lines = [l for ls in ex.map(one, chunks) for l in ls]
    return lines, [l for l in lines if not l.startswith("ok ")]


def _groups():
    g = {}
    for e in MANIFEST:
        g.setdefault(os.path.dirname(e["path"]), []).append(e)
    return g


def test_fixture_set_covers_the_manifests():
    g = _groups()
    assert len(MANIFEST) >= 500
    for d in ("8-bit/data", "10-bit/data", "12-bit/data", "8-bit/quantizer", "10-bit/quantizer", "8-bit/size", "8-bit/resize",
              "8-bit/film_grain", "10-bit/film_grain", "8-bit/features"):
        assert g.get(d), d
    assert any("--filmgrain" in e["args"] for e in MANIFEST)


needs_cpucheck = pytest.mark.skipif(not os.path.exists(os.path.join(BIN, "dav1d_b200_multi_cpucheck")),
                                    reason="oracle/_ref/dav1d_b200_multi_cpucheck not built")


@needs_cpucheck
def test_batch_builder_md5_cpu_checker():
    """Every fixture stream through the CPU checker: pins the batch builder, the level assignment and the stage order."""
    res, bad = _run_all("dav1d_b200_multi_cpucheck", MANIFEST, jobs=4)
    assert len(res) == len(MANIFEST)
    assert not bad, (len(bad), bad[:8])


def test_the_cli_form_matches_too():
    """The reference's own CLI with the host layer linked in (what `dav1d --verify` runs), on a few streams."""
    exe = os.path.join(BIN, "dav1d_b200_cpucheck")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/dav1d_b200_cpucheck not built")
    for e in MANIFEST[::97]:
        r = subprocess.run([exe, "-q", "-i", os.path.join(FIX, e["path"]), "--verify", e["md5"]] + e["args"], capture_output=True, text=True)
        assert r.returncode == 0, (e["path"], r.stderr[-300:])


@pytest.mark.gpu
def test_conformance_md5_gpu():
    """All fixture streams, every frame reconstructed and filtered on the GPU, MD5 against the reference's manifests."""
    assert os.path.exists(os.path.join(BIN, "dav1d_b200_multi")), "built by __graft_entry__.build() in the build container"
    res, bad = _run_all("dav1d_b200_multi", MANIFEST, jobs=int(os.environ.get("RB200_CONFORMANCE_JOBS", "4")))
    assert len(res) == len(MANIFEST) >= 500
    assert not bad, (len(bad), bad[:8])


@pytest.mark.gpu
def test_conformance_cli_gpu():
    """`dav1d_b200 --verify md5` (the reference's CLI over the GPU path) on a few streams of every bit depth."""
    exe = os.path.join(BIN, "dav1d_b200")
    for e in MANIFEST[::61]:
        r = subprocess.run([exe, "-q", "-i", os.path.join(FIX, e["path"]), "--verify", e["md5"]] + e["args"], capture_output=True, text=True)
        assert r.returncode == 0, (e["path"], r.stderr[-300:])
