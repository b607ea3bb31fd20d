#!/usr/bin/env python3
"""Where the end-to-end time of bench.py goes: per-direction variants of the e2e loop and host submit cost."""
import ctypes as C
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rav1d_b200 import lib  # noqa: E402
from rav1d_b200.synth import framegen  # noqa: E402

wl = bench.WORKLOADS["4k10"]
w, h, bpc, stages = wl[:4]
s = framegen.generate(w, h, bpc, seed=1)
lib.check(lib.init(0))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ctxs, outs = [], []
px = 2
ob = [w * h * px, (w // 2) * (h // 2) * px, (w // 2) * (h // 2) * px]
for i in range(N):
    d = framegen.DeviceFrame(s)
    d.load_batch(); d.set_ref_from_host(s.ref)
    ctxs.append(d)
    ptrs = []
    for nb in ob:
        p = C.c_void_p(); lib.check(lib.malloc_host(C.byref(p), nb)); ptrs.append(p.value)
    outs.append(((C.c_void_p * 3)(*ptrs), (C.c_ssize_t * 2)(w * px, (w // 2) * px)))
counts = (C.c_int32 * 19)(*[int(c) for c in s.itx_counts])
n_mc = len(s.mc_items)
for d in ctxs:
    lib.check(lib.frame_submit(d.h, s.n_coefs, counts, n_mc, stages, 1)); d.wait()
    lib.check(lib.frame_pack_coef16(d.h, s.n_coefs))
    lib.check(lib.frame_pack_coef_stream(d.h, s.n_coefs, counts, stages))
CHAIN = os.environ.get("RB200_PROBE_CHAIN", "1") == "1"
if CHAIN:      # frame i predicts from the output of frame i - 1, as in bench.py
    outs_pl = []
    for d in ctxs:
        pl = lib.Planes(); lib.check(lib.frame_output_planes(d.h, C.byref(pl))); outs_pl.append(pl)
    for i, d in enumerate(ctxs):
        lib.check(lib.frame_set_ref(d.h, 0, C.byref(outs_pl[i - 1])))


def run(upload, readback, frames=96):
    torch.cuda.synchronize()
    host = 0.0
    t0 = time.perf_counter()
    for i in range(frames):
        d = ctxs[i % N]
        d.wait()
        a = time.perf_counter()
        if CHAIN:
            lib.check(lib.frame_depend(d.h, ctxs[(i - 1) % N].h))
        lib.check(lib.frame_submit(d.h, s.n_coefs, counts, n_mc, stages, upload))
        if readback:
            lib.check(lib.frame_readback_async(d.h, outs[i % N][0], outs[i % N][1]))
        host += time.perf_counter() - a
    for d in ctxs:
        d.wait()
    dt = time.perf_counter() - t0
    return dt / frames * 1e3, host / frames * 1e3


for name, up, rb in [("packed16+readback", 5, True), ("packed16 only", 5, False), ("gather16+readback", 4, True), ("gather16 only", 4, False), ("gather+readback", 3, True), ("gather only", 3, False), ("zerocopy only", 2, False), ("copy only", 1, False),
                     ("readback only", 0, True), ("resident", 0, False)]:
    run(up, rb, 32)
    ms, hostms = run(up, rb)
    print(f"{name:18s} {ms:.3f} ms/frame  ({w * h / ms / 1e3:.0f} Mpx/s)  host submit {hostms:.3f} ms/frame")
