#!/usr/bin/env python3
"""One picture's post-filters split over the ranks of a torchrun job (one process per GPU):
CUDA-IPC mapping of the neighbours' planes, NVLink peer pulls of the halo rows, deblock + CDEF +
LR on each band, bit-exact check against the oracle's whole-picture result and timing.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/run_band_split.py [W H BPC] [--steps K]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("dims", nargs="*", type=int, default=[7680, 4320, 10])
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--no-check", action="store_true")
    args = ap.parse_args()
    w, h, bpc = args.dims
    import torch
    import torch.distributed as dist
    from rav1d_b200 import lib, multigpu as mg
    from rav1d_b200.synth import framegen
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib.check(lib.init(local))
    s = framegen.generate(w, h, bpc, seed=4)          # same seed on every rank: the same picture
    start = framegen.recon_input_planes(s)
    ranges = mg.split_stripes(h, world)
    d = framegen.DeviceFrame(s)
    d.load_batch()
    band = mg.BandContext(s.hdr, ranges, rank, frame_handle=d.h)
    band.upload_owned(start)
    handles = mg.exchange_bytes(band.ipc_handle())
    band.open_peers(handles)
    torch.cuda.synchronize(); dist.barrier()           # every rank's own rows are resident

    stream = torch.cuda.ExternalStream(lib.frame_stream(d.h))
    flag = torch.zeros(1, device="cuda")
    def step():
        band.pull_halo()                                # P2P over NVLink, on the frame's stream
        # deblocking is in place, so no rank may start it while a neighbour is still reading its rows:
        # a 4-byte all-reduce enqueued behind the pulls is the stream-ordered cross-GPU barrier (no host sync)
        with torch.cuda.stream(stream):
            dist.all_reduce(flag)
        if not band.empty:
            d.submit(14, upload=False)
    d.submit(14, upload=True); d.wait()                 # metadata resident; warm-up
    for _ in range(3):
        step()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / args.steps

    ok = None
    if not args.no_check:
        # the deblock stage is in place: restore the inputs, run once more, compare this band
        band.upload_owned(start)
        torch.cuda.synchronize(); dist.barrier()
        step(); d.wait()
        out = [np.zeros_like(p) for p in s.ref]
        band.readback_owned(out)
        import framecheck, refharness
        if rank == 0:
            ref = refharness.load()
            exp = framecheck.oracle_frame(ref, s, 14, n_tc=os.cpu_count() or 1, start_planes=start)
            blob = [e.copy() for e in exp]
        else:
            blob = None
        lst = [blob]
        dist.broadcast_object_list(lst, src=0)
        exp = lst[0]
        lo, hi = band.out_rows
        good = True
        for p in range(3):
            a, b = (lo, hi) if p == 0 else (lo >> 1, (hi + 1) >> 1)
            good &= bool(np.array_equal(exp[p][a:b], out[p][a:b, :exp[p].shape[1]]))
        good_t = torch.tensor([1 if good else 0], device="cuda")
        dist.all_reduce(good_t, op=dist.ReduceOp.MIN)
        ok = bool(good_t.item())
    hb = torch.tensor([band.halo_bytes()], dtype=torch.float64, device="cuda")
    dist.all_reduce(hb, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"what": "post-filters of one picture split by stripe rows with P2P halo pulls", "width": w,
                          "height": h, "bpc": bpc, "n_gpus": world, "ms_per_frame": ms, "mpixel_per_s": w * h / ms / 1e3,
                          "max_halo_bytes_per_rank": hb.item(), "bit_exact_vs_oracle": ok}))
    band.close()
    dist.destroy_process_group()
    if ok is False:
        sys.exit(1)


if __name__ == "__main__":
    main()
