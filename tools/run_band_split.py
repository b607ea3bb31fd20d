#!/usr/bin/env python3
"""One picture's post-filters split over the ranks of a torchrun job (one process per GPU): CUDA-IPC mapping of the
neighbours' planes, NVLink peer pulls of the halo rows ordered by cross-GPU flags, deblock + CDEF + LR on each band,
bit-exact check against the oracle's whole-picture result and timing (rav1d_b200.multigpu.run_band_split).

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/run_band_split.py [W H BPC] [--steps K] [--in-flight F]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("dims", nargs="*", type=int, default=[7680, 4320, 10])
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--in-flight", type=int, default=2)
    ap.add_argument("--no-check", action="store_true")
    args = ap.parse_args()
    w, h, bpc = args.dims
    import torch
    import torch.distributed as dist
    from rav1d_b200 import lib, multigpu as mg
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib.check(lib.init(local))
    res = mg.run_band_split(w, h, bpc, steps=args.steps, in_flight=args.in_flight, check=not args.no_check)
    if rank == 0:
        print(json.dumps(res))
    dist.destroy_process_group()
    if res["bit_exact_vs_oracle"] is False:
        sys.exit(1)


if __name__ == "__main__":
    main()
