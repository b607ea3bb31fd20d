#!/usr/bin/env python3
"""Host<->device copy bandwidth of the box: H2D alone, D2H alone, both at once (pinned memory, 32 MB chunks)."""
import torch

n = 32 << 20
h_in = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(4)]
h_out = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(4)]
d_in = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(4)]
d_out = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(4)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=40):
    torch.cuda.synchronize()
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


for _ in range(2):
    a, b, c = run(True, False), run(False, True), run(True, True)
print(f"H2D alone {a:.1f} GB/s, D2H alone {b:.1f} GB/s, both at once {c:.1f} + {c:.1f} GB/s")
