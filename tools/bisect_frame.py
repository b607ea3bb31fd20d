#!/usr/bin/env python3
"""Debug aid: run one synthetic frame through oracle and product stage by stage, feeding each
product stage the ORACLE's previous-stage picture so a mismatch is attributed to one stage.
usage: python tools/bisect_frame.py W H BPC [seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import framecheck  # noqa: E402
import refharness  # noqa: E402
from rav1d_b200 import lib  # noqa: E402
from rav1d_b200.synth import framegen  # noqa: E402


def full(s, vis):
    out = [np.zeros_like(p) for p in s.ref]
    for p in range(3):
        out[p][:vis[p].shape[0], :vis[p].shape[1]] = vis[p]
    return out


def report(name, a, b):
    for p in range(3):
        bad = np.argwhere(a[p] != b[p])
        if len(bad):
            print(f"  {name} plane {p}: {len(bad)} px differ; x [{bad[:, 1].min()}, {bad[:, 1].max()}] "
                  f"y [{bad[:, 0].min()}, {bad[:, 0].max()}]; first {bad[0][::-1]} oracle {a[p][tuple(bad[0])]} "
                  f"product {b[p][tuple(bad[0])]}")
            ys, xs = bad[:, 0], bad[:, 1]
            print("    distinct 64x64 cells:", sorted(set(zip((ys // 64).tolist(), (xs // 64).tolist())))[:12])
        else:
            print(f"  {name} plane {p}: equal")


def main():
    w, h, bpc = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    seed = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    lib.check(lib.init(0))
    ref = refharness.load()
    s = framegen.generate(w, h, bpc, seed=seed)
    # NOTE: the oracle's full pictures incl. the 8-px-aligned padding are needed as stage inputs,
    # so run it on full planes
    def oracle(stages, start=None):
        cur = refharness.RefFrame(ref, s, 1); rf = refharness.RefFrame(ref, s, 1)
        try:
            rf.set_planes(s.ref); cur.load_filter_meta()
            if stages & 1:
                cur.recon(rf)
            else:
                cur.set_planes(start)
            if stages & ~1:
                cur.filter(stages)
            return cur.get_planes()
        finally:
            cur.close(); rf.close()

    def product(stages, start=None):
        d = framegen.DeviceFrame(s)
        try:
            d.load_batch()
            if stages & 1:
                d.set_ref_from_host(s.ref)
            else:
                d.upload(0, start)
            d.submit(stages); d.wait()
            return d.readback()
        finally:
            d.close()

    vis = lambda pl: framecheck.visible(s, pl)
    o_r = oracle(1)
    report("recon", vis(o_r), vis(product(1)))
    o_d = oracle(2, o_r)
    report("deblock(from oracle recon)", vis(o_d), vis(product(2, o_r)))
    o_c = oracle(4, o_d)
    report("cdef(from oracle deblock)", vis(o_c), vis(product(4, o_d)))
    o_dc = oracle(6, o_r)
    report("oracle cdef-alone == deblock+cdef", vis(o_c), vis(o_dc))
    o_l = oracle(14, o_r)
    report("deblock+cdef+lr(from oracle recon)", vis(o_l), vis(product(14, o_r)))
    report("all", vis(oracle(15)), vis(product(15)))


if __name__ == "__main__":
    main()
