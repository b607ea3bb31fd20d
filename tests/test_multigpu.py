"""Splitting one picture's post-filters over several GPUs by superblock rows (BASELINE config 4).

CPU: the stripe partition / halo plan, and the handle exchange over torch.distributed (gloo,
world_size 2).  GPU: the banded contexts and the peer pulls, emulated as N contexts on one device
(bit-exact against the oracle run on the whole picture); with >= 2 GPUs the same through CUDA IPC
is exercised by tools/run_band_split.py under torchrun.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import framecheck

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_split_covers_picture_without_overlap():
    from rav1d_b200 import multigpu as mg
    for h in (64, 120, 1080, 2160, 4320, 57):
        for n in (1, 2, 3, 4, 8):
            ranges = mg.split_stripes(h, n)
            assert len(ranges) == n and ranges[0][0] == 0 and ranges[-1][1] == (h + 8 + 63) // 64
            rows = [mg.owned_rows(h, *r) for r in ranges]
            covered = 0
            for (lo, hi), (s0, s1) in zip(rows, ranges):
                if s1 > s0:
                    assert lo == covered
                    covered = hi
            assert covered == h


def test_halo_plan_covers_needed_rows():
    from rav1d_b200 import multigpu as mg
    h, n = 4320, 8
    ranges = mg.split_stripes(h, n)
    for rank in range(n):
        own = mg.owned_rows(h, *ranges[rank])
        need = (max(own[0] - 84, 0), min(own[1] + 24, h))
        plan = mg.halo_plan(h, ranges, rank, need)
        got = np.zeros(h, bool)
        got[own[0]:own[1]] = True
        for peer, a, b in plan:
            assert peer != rank and mg.owned_rows(h, *ranges[peer])[0] <= a < b <= mg.owned_rows(h, *ranges[peer])[1]
            got[a:b] = True
        assert got[need[0]:need[1]].all()
        assert all(abs(peer - rank) == 1 for peer, _, _ in plan)   # 8K over 8 GPUs: nearest neighbours only


def test_handle_exchange_over_gloo():
    """exchange_bytes() is what carries the CUDA IPC handles; run it with world_size 2 on CPU."""
    code = (
        "import os, sys\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import torch.distributed as dist\n"
        "from rav1d_b200 import multigpu as mg\n"
        "dist.init_process_group('gloo')\n"
        "r = dist.get_rank()\n"
        "out = mg.exchange_bytes(bytes([r]) * 64)\n"
        "assert out == [bytes([0]) * 64, bytes([1]) * 64], out\n"
        "ranges = mg.split_stripes(2160, dist.get_world_size())\n"
        "assert mg.owned_rows(2160, *ranges[r]) == ((0, 1080) if r == 0 else (1080, 2160)), mg.owned_rows(2160, *ranges[r])\n"
        "dist.destroy_process_group()\n"
        "print('ok', r)\n")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "_gloo_exchange.py")],
                       capture_output=True, text=True, timeout=240,
                       env={**os.environ, "RB200_GLOO_CODE": code})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    # the two ranks share stdout: their lines may interleave
    assert r.stdout.count("ok") == 2 and "0" in r.stdout and "1" in r.stdout, r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,bpc,n", [(424, 600, 10, 3), (200, 330, 8, 2), (264, 1000, 12, 4), (640, 360, 10, 8)])
def test_banded_filters_match_whole_picture(rb, ref, w, h, bpc, n):
    from rav1d_b200 import multigpu as mg
    from rav1d_b200.synth import framegen
    s = framegen.generate(w, h, bpc, seed=n + bpc)
    start = framegen.recon_input_planes(s)
    exp = framecheck.oracle_frame(ref, s, 14, start_planes=start)
    ranges = mg.split_stripes(h, n)
    ctxs = []
    try:
        for rank in range(n):
            d = framegen.DeviceFrame(s)
            d.load_batch()
            ctxs.append((d, mg.BandContext(s.hdr, ranges, rank, frame_handle=d.h)))
        for d, b in ctxs:
            b.upload_owned(start)
        for d, b in ctxs:                       # same process: the peers' blocks are plain device pointers
            for peer, _, _ in b.plan():
                b.peer_bases[peer] = ctxs[peer][1].plane_block()[0]
            b.pull_halo()
        for d, b in ctxs:                       # deblocking is in place: every pull is complete before any band filters
            d.wait()                            # (between GPUs this is the flag protocol of multigpu.BandRing)
        for d, b in ctxs:
            if not b.empty:
                d.submit(14)
        out = [np.zeros_like(p) for p in s.ref]
        for d, b in ctxs:
            if not b.empty:
                d.wait()
                b.readback_owned(out)
        framecheck.assert_planes_equal(exp, framecheck.visible(s, out), f"{n} bands {w}x{h}@{bpc}")
    finally:
        for d, b in ctxs:
            b.peer_bases = {}
            b.h = None
            d.close()


def test_pullers_are_the_mirror_of_the_halo_plans():
    """A rank waits for exactly the ranks whose halo plan names it (multigpu.pullers_of)."""
    from rav1d_b200 import multigpu as mg
    for height, n in ((4320, 8), (2160, 4), (1080, 3), (600, 8), (100, 4)):
        ranges = mg.split_stripes(height, n)
        # stand-in for rb200_frame_band_rows(): what a band reads = its own rows widened by the stages' reach
        in_rows = []
        for r in range(n):
            lo, hi = mg.owned_rows(height, *ranges[r])
            in_rows.append((max(lo - 84, 0), min(hi + 24, height)) if hi > lo else (0, 0))
        plans = [mg.halo_plan(height, ranges, r, in_rows[r]) if ranges[r][1] > ranges[r][0] else [] for r in range(n)]
        for r in range(n):
            expect = sorted(q for q in range(n) if any(peer == r for peer, _, _ in plans[q]))
            assert sorted(mg.pullers_of(height, ranges, r, in_rows)) == expect, (height, n, r)


@pytest.mark.gpu
def test_band_split_across_gpus():
    """The flag-ordered band split on real peers: one process per GPU (torchrun), every band bit-exact against the oracle's
    whole-picture result.  Needs two devices; the single-GPU test box skips it (tools/run_band_split.py is the same run)."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    n = min(n, 4)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", "29534", os.path.join(ROOT, "tools", "run_band_split.py"), "1280", "720", "10", "--steps", "4"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert '"bit_exact_vs_oracle": true' in r.stdout, r.stdout[-1000:]
