#!/usr/bin/env python3
"""Host<->device copy bandwidth with every GPU of the box busy at once: run under torchrun (one process per GPU); each
rank copies 32 MB pinned chunks H2D and D2H simultaneously between barriers; rank 0 prints per-rank and aggregate GB/s
(the ceiling of bench.py's end-to-end leg at N GPUs).  Also prints the host's NUMA layout and the GPUs' CPU affinity."""
import os
import subprocess

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 32 << 20
h_in = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(4)]
h_out = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(4)]
d_in = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(4)]
d_out = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(4)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=60):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    s1.wait_event(e0); s2.wait_event(e0)
    for i in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in[i % 4].copy_(h_in[i % 4], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out[i % 4].copy_(d_out[i % 4], non_blocking=True)
    e1.record(s1); e2.record(s2)
    torch.cuda.synchronize()
    ms = max(e0.elapsed_time(e1), e0.elapsed_time(e2))
    return reps * n / ms / 1e6


res = []
for _ in range(2):
    res = [run(True, False), run(False, True), run(True, True)]
t = torch.tensor(res, dtype=torch.float64, device="cuda")
if world > 1:
    allr = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allr, t)
else:
    allr = [t]
if rank == 0:
    rows = [[float(x) for x in a.tolist()] for a in allr]
    for r, (a, b, c) in enumerate(rows):
        print(f"rank {r}: H2D alone {a:.1f} GB/s, D2H alone {b:.1f} GB/s, both at once {c:.1f} + {c:.1f} GB/s")
    print(f"{world} ranks at once, aggregate: H2D {sum(r[0] for r in rows):.1f} GB/s, D2H {sum(r[1] for r in rows):.1f} GB/s, "
          f"both {sum(r[2] for r in rows):.1f} + {sum(r[2] for r in rows):.1f} GB/s")
    for cmd in (["nvidia-smi", "topo", "-m"], ["lscpu"]):
        try:
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=30).stdout
            print("\n".join(l for l in out.splitlines() if cmd[0] != "lscpu" or any(k in l for k in ("NUMA", "Socket", "Model name", "CPU(s):"))))
        except Exception as e:  # noqa: BLE001
            print(cmd, "failed:", e)
if world > 1:
    dist.destroy_process_group()
