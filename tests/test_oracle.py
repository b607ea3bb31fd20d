"""CPU-only checks of the test infrastructure and of the drop-in boundary.

* The oracle is the reference's own C DSP compiled in place (oracle/Makefile ->
  oracle/_ref/).  It is pinned against the reference's golden vectors: the decoded-frame
  MD5 manifests of its conformance streams (tests/dav1d-test-data/**/meson.build),
  which exercise exactly the DSP tables the parity tests call.
* The frame harness (oracle/ref_frame.c) gives the same picture in the reference's
  single-thread sbrow order and in its tile-thread buffer layout with stages flattened.
* librav1d_b200.so loads and exports every symbol include/rav1d_b200.h declares.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import framecheck
import refharness

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DATA = "/root/reference/tests/dav1d-test-data"
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "dav1d_ref")


def _manifest(subdir):
    """(ivf path, md5) pairs of one manifest."""
    d = os.path.join(REF_DATA, subdir)
    txt = open(os.path.join(d, "meson.build")).read()
    return [(os.path.join(d, m.group(1)), m.group(2))
            for m in re.finditer(r"files\('([^']+)'\),\s*'([0-9a-f]{32})'", txt)]


def _decode_md5(path, *extra):
    out = subprocess.run([REF_BIN, "-q", "--threads", "1", "-i", path, "-o", "-", "--muxer", "md5", *extra],
                         check=True, capture_output=True, text=True).stdout
    return out.split()[0]


needs_ref = pytest.mark.skipif(not (os.path.isdir(REF_DATA) and os.path.exists(REF_BIN)),
                               reason="reference tree / oracle build not present (GPU box)")


@needs_ref
@pytest.mark.parametrize("subdir,step", [("8-bit/quantizer", 8), ("10-bit/quantizer", 8), ("8-bit/data", 6),
                                         ("8-bit/size", 16), ("8-bit/resize", 4), ("8-bit/features", 1),
                                         ("8-bit/intra", 1), ("8-bit/mv", 1), ("8-bit/mfmv", 1),
                                         ("10-bit/data", 1), ("10-bit/features", 1),
                                         ("12-bit/data", 1), ("12-bit/features", 1)])
def test_oracle_matches_reference_md5_manifests(subdir, step):
    if not os.path.exists(os.path.join(REF_DATA, subdir, "meson.build")):
        pytest.skip("no manifest")
    entries = [e for e in _manifest(subdir) if os.path.exists(e[0])][::step]
    assert entries
    for path, md5 in entries:
        assert _decode_md5(path) == md5, path


@needs_ref
def test_oracle_film_grain_md5():
    # tests/dav1d-test-data/10-bit/meson.build:43-47
    p = os.path.join(REF_DATA, "10-bit/film_grain/av1-1-b10-23-film_grain-50.ivf")
    assert _decode_md5(p, "--filmgrain", "1") == "be596f5921854b9a9a5be81c302a5327"


def test_oracle_frame_harness_schedules_agree(ref):
    from rav1d_b200.synth import framegen
    for (w, h, bpc, stages) in ((200, 120, 8, 15), (264, 200, 10, 14), (100, 68, 12, 15), (424, 296, 10, 6)):
        s = framegen.generate(w, h, bpc, seed=w)
        start = framegen.recon_input_planes(s)
        a = framecheck.oracle_frame(ref, s, stages, n_tc=1, start_planes=start)
        b = framecheck.oracle_frame(ref, s, stages, n_tc=3, start_planes=start)
        framecheck.assert_planes_equal(a, b, f"schedules {w}x{h}")


def test_oracle_frame_harness_is_stateless(ref):
    """Filtering does not edit the masks (single tile) and stage-at-a-time equals stages together."""
    from rav1d_b200 import lib
    from rav1d_b200.synth import framegen
    s = framegen.generate(1000, 600, 10, seed=9)
    start = framegen.recon_input_planes(s)
    cur = refharness.RefFrame(ref, s, 1)
    try:
        cur.load_filter_meta(); cur.set_planes(start); cur.filter(2)
        n = s.geom.sb128w * s.geom.sb128h
        assert np.array_equal(lib.np_view(ref.ref_frame_masks(cur.h), lib.AV1_FILTER_DT, n), s.masks)
        cur.filter(4)
        a = framecheck.visible(s, cur.get_planes())
    finally:
        cur.close()
    b = framecheck.oracle_frame(ref, s, 6, start_planes=start)
    framecheck.assert_planes_equal(a, b, "D then C vs D|C")


def test_oracle_stages_change_the_picture(ref):
    """The synthetic frame makes every stage fire (otherwise parity would be vacuous)."""
    from rav1d_b200.synth import framegen
    s = framegen.generate(264, 200, 10, seed=4)
    r = framecheck.oracle_frame(ref, s, 1)
    d = framecheck.oracle_frame(ref, s, 3)
    c = framecheck.oracle_frame(ref, s, 7)
    l = framecheck.oracle_frame(ref, s, 15)
    for p in range(3):
        assert (r[p] != s.ref[p][:r[p].shape[0], :r[p].shape[1]]).mean() > 0.5
        assert (d[p] != r[p]).mean() > 0.02
        assert (c[p] != d[p]).mean() > 0.05
        assert (l[p] != c[p]).mean() > 0.05


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "rav1d_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(rb200_[a-z0-9_]+)\s*\(", hdr))
    declared |= set(re.findall(r"\(\*(rb200_[a-z0-9_]+)\(", hdr))      # functions returning array pointers
    declared -= set(re.findall(r"\(\*(rb200_[a-z0-9_]+)\)\s*\(", hdr))  # fn-pointer typedefs
    assert len(declared) >= 55, sorted(declared)
    from rav1d_b200 import lib
    missing = [n for n in sorted(declared) if not hasattr(lib.cdll, n)]
    assert not missing, missing
    assert lib.abi_version() == 2


def test_record_layouts_match_the_reference(ref):
    """Batch records that reuse the reference's structs have the reference's sizes."""
    from rav1d_b200 import lib
    assert ref.ref_sizeof(0) == lib.AV1_FILTER_DT.itemsize            # Av1Filter
    assert ref.ref_sizeof(1) == lib.AV1_RESTORATION_DT.itemsize       # Av1Restoration
    assert ref.ref_sizeof(2) == C.sizeof(lib.FilterLUT)               # Av1FilterLUT
    assert ref.ref_sizeof(3) == lib.LR_UNIT_DT.itemsize               # Av1RestorationUnit


def test_constant_tables_match_the_reference(ref):
    """rav1d_b200/csrc/tables_data.inc (generated by tools/gen_tables.py) against the
    reference's tables as compiled into the oracle."""
    r = subprocess.run(["python", os.path.join(ROOT, "tools", "gen_tables.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
