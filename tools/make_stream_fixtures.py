#!/usr/bin/env python3
"""Generate tests/golden/streams.npz: for a few small conformance streams of the reference tree,
per decoded frame, the pre-filter picture, the decoder's filter metadata and the post-filter picture,
as dumped by oracle/_ref/ref_dump (the reference decoder with dav1d_filter_sbrow_* interposed).
Planes are cropped to the picture rounded up to 8 pixels.  Run where /root/reference exists:

    python tools/make_stream_fixtures.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import streamdump  # noqa: E402

STREAMS = [  # (path under tests/dav1d-test-data, frames)
    ("8-bit/data/00000658.ivf", 2), ("8-bit/data/00000664.ivf", 2), ("8-bit/data/00000668.ivf", 2),
    ("8-bit/data/00000593.ivf", 2), ("8-bit/issues/309_odd_width.ivf", 2), ("8-bit/data/00000527.ivf", 3),
    ("10-bit/data/00000671.ivf", 2), ("10-bit/data/00000676.ivf", 2), ("10-bit/data/00000682.ivf", 2),
    ("10-bit/argon/test185_302.obu", 2), ("10-bit/argon/test5606.obu", 2),
    ("12-bit/data/00000686.ivf", 2), ("12-bit/data/00000692.ivf", 2), ("12-bit/data/00000696.ivf", 2),
    ("12-bit/argon/test15240.obu", 2), ("multi-bit/argon/test10218_6914.obu", 2),
]


def pack(s):
    """One StreamFrame -> dict of arrays (see tests/streamdump.py unpack())."""
    h = s.hdr
    ints = [s.w, s.h, s.bpc, s.layout, h.sb128, h.lf_level_y[0], h.lf_level_y[1], h.lf_level_u, h.lf_level_v, s.cdef_on,
            h.cdef_damping, *h.cdef_y_strength, *h.cdef_uv_strength, *h.lr_type, *h.lr_unit_size_log2, *s.tiles, s.ah, s.stages]
    g = s.geom
    out = {"ints": np.array(ints, np.int32), "masks": s.masks.view(np.uint8), "levels": s.levels,
           "lut": np.frombuffer(bytes(s.lut), np.uint8), "lr_masks": s.lr_masks.view(np.uint8)}
    out_w = getattr(s, "out_w", s.w)
    if out_w != s.w:                       # super-resolution: the filtered picture is wider than the coded one
        out["ints"] = np.array(ints + [out_w], np.int32)
    for k, planes in (("pre", s.pre), ("post", s.post)):
        for p, a in enumerate(planes):
            wk = (g.bw * 4) if k == "pre" or out_w == s.w else (out_w + 7) & ~7
            rows, cols = (g.bh * 4) >> (g.ss_ver if p else 0), wk >> (g.ss_hor if p else 0)
            out[f"{k}{p}"] = a[:rows, :cols]
    return out


SR_STREAMS = [  # frames coded with super-resolution (8/10 bit, several scaling ratios, Wiener / SGR / no restoration)
    ("8-bit/data/00000802.ivf", 2), ("8-bit/data/00000855.ivf", 3), ("8-bit/issues/323_tennis.ivf", 2),
    ("8-bit/data/00000863.ivf", 1), ("10-bit/data/00000826.ivf", 2), ("10-bit/data/00000832.ivf", 2),
]


SIZE_STREAMS = [  # the reference's picture-size sweep (tests/dav1d-test-data/8-bit/size): 16 .. 66 and 196 .. 226 pixels
    (f"8-bit/size/av1-1-b8-01-size-{w}x{h}.ivf", 2)
    for w, h in ((16, 16), (16, 18), (18, 34), (34, 16), (32, 66), (66, 18), (64, 64), (66, 66), (196, 198), (202, 210), (226, 196), (226, 226))
]


def main(streams=STREAMS, path=streamdump.GOLDEN, sr=False):
    blob, index = {}, []
    for rel, n in streams:
        frames = streamdump.dump(os.path.join(streamdump.REF_DATA, rel), n, sr=sr)
        assert frames, rel
        for s in frames:
            key = f"{rel}#{s.index}"
            index.append(key)
            for k, v in pack(s).items():
                blob[f"{key}/{k}"] = v
    blob["index"] = np.array(index)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez_compressed(path, **blob)
    print(f"{len(index)} frames -> {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


GRAIN_STREAMS = [("10-bit/film_grain/av1-1-b10-23-film_grain-50.ivf", 2), ("8-bit/film_grain/av1-1-b8-23-film_grain-50.ivf", 2)]


def main_grain():
    blob, index = {}, []
    for rel, n in GRAIN_STREAMS:
        recs = streamdump.dump_grain(os.path.join(streamdump.REF_DATA, rel), n)
        assert recs, rel
        for i, r in enumerate(recs):
            key = f"{rel}#{i}"
            index.append(key)
            blob[f"{key}/ints"] = np.array([r["w"], r["h"], r["bpc"], r["layout"], r["is_id"]], np.int32)
            blob[f"{key}/fg"] = r["fg"]
            for p in range(len(r["inp"])):
                blob[f"{key}/in{p}"] = r["inp"][p]; blob[f"{key}/out{p}"] = r["out"][p]
    blob["index"] = np.array(index)
    np.savez_compressed(streamdump.GOLDEN_GRAIN, **blob)
    print(f"{len(index)} film-grain frames -> {streamdump.GOLDEN_GRAIN} ({os.path.getsize(streamdump.GOLDEN_GRAIN) / 1e6:.2f} MB)")


def main_lfb():
    """tests/golden/streams_lfb.npz: the per-block create_lf_mask arguments of the deblocked frames of STREAMS
    (same keys as streams.npz, which holds the masks and levels the decoder derived from them)."""
    blob, index = {}, []
    for rel, n in STREAMS:
        for idx, rec in sorted(streamdump.dump_lfb(os.path.join(streamdump.REF_DATA, rel), n).items()):
            if len(rec):
                key = f"{rel}#{idx}"
                index.append(key)
                blob[key] = rec.view(np.uint8)
    blob["index"] = np.array(index)
    np.savez_compressed(streamdump.GOLDEN_LFB, **blob)
    print(f"{len(index)} frames -> {streamdump.GOLDEN_LFB} ({os.path.getsize(streamdump.GOLDEN_LFB) / 1e6:.2f} MB)")


if __name__ == "__main__":
    if sys.argv[1:] == ["lfb"]:
        main_lfb()
        sys.exit(0)
    main()
    main(SR_STREAMS, streamdump.GOLDEN_SR, sr=True)
    main(SIZE_STREAMS, streamdump.GOLDEN_SIZES)
    main_grain()
    main_lfb()
