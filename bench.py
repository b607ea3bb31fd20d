#!/usr/bin/env python3
"""Throughput of the recon + post-filter path on synthetic frames (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload 4k10|1080p8|8k10|4k10c5]

A *step* is one pass of the hot path over one batch of FRAMES_PER_STEP synthetic frames:
recon (8-tap MC + itx add) -> deblock -> CDEF -> Wiener/SGR loop restoration (4k10: BASELINE
configs[2], the configuration the metric is quoted on).

* `value`  : Mpixel/s (luma pixels of output frames / s) of ONE video stream with the batch resident in
             HBM, timed with CUDA events on the launching streams: frames cycle over N_CTX frame
             contexts (working sets > 190 MB each at 4K, beyond the 126 MB L2), and every frame
             predicts from the OUTPUT of the frame before it (rb200_frame_depend: its kernels wait
             for the previous frame's last kernel) -- the dependency chain of real inter frames.
* `value_unchained`: the same contexts with a static reference and no dependency, i.e. N_CTX
             independent streams sharing the GPU (what round 1 reported as `value`).
* `e2e`    : `value`'s chained stream through the C ABI with HOST buffers: per frame H2D of
             coefficients, work items and filter metadata from pinned staging, the stage launches, D2H
             of the output picture into pinned host planes; N_CTX frames in flight on their own
             streams (uploads and read-backs of neighbouring frames overlap the kernels).
* `roofline`: the dominant kernel's algorithmic bytes / its CUDA-event duration (stage marks of
             rb200_frame_stage_times, recorded inside the timed region) against MEASURED_PEAKS.json.
* `cpu_baseline`: the reference's own C DSP + frame drivers (oracle/_ref, all host threads) on a
             bounded sample of the same workload.
* `--impl reference`: only that CPU arm, in the same JSON shape.

N > 1 (torchrun): independent streams, one per GPU, no data-path collective;
value = pixels of all ranks / max-over-ranks device time; "scaling": "weak".
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (w, h, bpc, stages, description)
    "4k10": (3840, 2160, 10, 15, "4K 10-bit 4:2:0 synthetic inter frame: recon (8-tap/bilinear MC + itx) + deblock + CDEF + Wiener/SGR LR"),
    "1080p8": (1920, 1080, 8, 3, "1080p 8-bit 4:2:0 synthetic inter frame: recon (8-tap/bilinear MC + itx) + deblock"),
    "8k10": (7680, 4320, 10, 14, "8K 10-bit 4:2:0 post-filters only: deblock + CDEF + Wiener/SGR LR"),
    # BASELINE configs[4] per GPU: 50 % compound blocks (avg / w_avg / segmentation mask), 5 % warped, OBMC on 10 %, film grain
    "4k10c5": (3840, 2160, 10, 31, "4K 10-bit 4:2:0 synthetic inter frame, 50% compound / 5% warped / 10% OBMC blocks: "
                                   "recon + deblock + CDEF + Wiener/SGR LR + film grain"),
}
GEN_ARGS = {"4k10c5": dict(comp_frac=0.5, warp_frac=0.05, obmc_frac=0.1)}
FRAMES_PER_STEP = int(os.environ.get("RB200_BENCH_FRAMES_PER_STEP", "256"))   # x steps: a timed region of > 1 s
N_CTX = int(os.environ.get("RB200_BENCH_CTX", "8"))
# BASELINE.json's metric, verbatim; `value` is its Mpixel/s part (luma pixels of output frames per second), the
# "% of HBM roofline" part is `frame_roofline_frac` (whole frame) and `roofline` (dominant kernel).
METRIC = "Mpixel/s recon+post-filter (itx+MC+LF/CDEF/LR) 4K 10-bit; % of HBM roofline"


def metric_name(workload):
    return METRIC if workload.startswith("4k10") else METRIC.replace("4K 10-bit", {"1080p8": "1080p 8-bit", "8k10": "8K 10-bit"}[workload])


STAGE_NAMES = ["h2d", "mc", "itx", "deblock", "cdef", "lr", "film_grain"]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(w, h, bpc, comp_frac=0.0):
    """SURVEY 8(d) / BASELINE.md 4: S = one 4:2:0 plane set, C = dense coefficients; a compound block
    reads two references."""
    px, cs = (2, 4) if bpc > 8 else (1, 2)
    S = w * h * 3 // 2 * px
    Cb = w * h * 3 // 2 * cs
    mc = int((2 + comp_frac) * S)
    return {"mc": mc, "itx": Cb + 2 * S, "recon": mc + Cb, "deblock": 2 * S, "cdef": 2 * S, "lr": 2 * S,
            "film_grain": 2 * S, "S": S, "C": Cb}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        gid = vis.split(",")[index] if vis else str(index)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", gid, f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 6]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------ CPU arm
def cpu_frames(s, stages, n_frames, n_threads):
    """Time `n_frames` passes of the reference's C path (oracle/_ref) over the synthetic frame.
    Returns seconds (restoring the consumed coefficient buffer is not timed)."""
    import refharness
    ref = refharness.load()
    cur = refharness.RefFrame(ref, s, max(n_threads, 2))
    rf = refharness.RefFrame(ref, s, 1)
    rf2 = refharness.RefFrame(ref, s, 1) if hasattr(s, "ref2") else None
    start = None
    if not stages & 1:
        from rav1d_b200.synth import framegen
        start = framegen.recon_input_planes(s)
    try:
        rf.set_planes(s.ref)
        if rf2:
            rf2.set_planes(s.ref2)
        cur.load_filter_meta()
        cw = s.coef.copy()
        total = 0.0
        for _ in range(n_frames):
            if stages & 1:
                np.copyto(cw, s.coef)
            else:
                cur.set_planes(start)
            t0 = time.perf_counter()
            if stages & 1:
                cur.recon(rf, n_threads=n_threads, coef_work=cw, ref_frame2=rf2)
            if stages & 14:
                cur.filter(stages & 14, n_threads=n_threads)
            if stages & 16:
                cur.apply_grain_inplace(s.film_grain, 0)
            total += time.perf_counter() - t0
        return total
    finally:
        cur.close()
        rf.close()
        if rf2:
            rf2.close()


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference(args, s, wl):
    w, h, bpc, stages, desc = wl[:5]
    cores = os.cpu_count() or 1
    t1 = cpu_frames(s, stages, 1, cores)                       # warm-up + estimate
    per_step = max(1, min(FRAMES_PER_STEP, int(1.5 / max(t1, 1e-3))))   # ~1.5 s of CPU work per step
    for _ in range(max(args.warmup - 1, 0)):
        cpu_frames(s, stages, 1, cores)
    t = 0.0
    for _ in range(args.steps):
        t += cpu_frames(s, stages, per_step, cores)
    frames = per_step * args.steps
    mpx = frames * w * h / t / 1e6
    sample = f"{frames} frames of the workload ({per_step} per step), {cores} threads, sbrow-parallel stages"
    line = {"impl": "reference", "metric": metric_name(args.workload), "value": mpx, "unit": "Mpixel/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": desc, "frames_per_step": per_step, "width": w, "height": h, "bpc": bpc,
                       "cpu": cpu_model(), "note": "reference's portable C DSP (no asm: nasm unavailable), -O3 x86-64-v3"},
            "cpu_baseline": {"value": mpx, "unit": "Mpixel/s", "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": mpx, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------ GPU arm
def run_gpu(args, s, wl):
    import torch
    import torch.distributed as dist
    from rav1d_b200 import lib
    from rav1d_b200.synth import framegen

    w, h, bpc, stages, desc = wl
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; rav1d_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib.check(lib.init(local), "rb200_init")

    ctxs = []
    start_planes = None if stages & 1 else framegen.recon_input_planes(s)
    if args.lf == "records":
        s.lf_blocks = s.lf_block_records
    for i in range(N_CTX):
        d = framegen.DeviceFrame(s)
        d.load_batch()
        if stages & 1:
            d.set_ref_from_host(s.ref)
            if hasattr(s, "ref2"):
                d.set_ref_slot(1, s.ref2)
        else:
            d.upload(0, start_planes)
        if stages & 16:
            lib.check(lib.frame_set_film_grain(d.h, C.byref(s.film_grain), 0))
        lib.check(lib.frame_enable_timing(d.h, 1))
        ctxs.append(d)
    # pinned host output planes for the e2e leg
    px = 2 if bpc > 8 else 1
    out_bytes = [w * h * px, ((w + 1) // 2) * ((h + 1) // 2) * px, ((w + 1) // 2) * ((h + 1) // 2) * px]
    host_out = []
    for d in ctxs:
        ptrs = []
        for nb in out_bytes:
            p = C.c_void_p()
            lib.check(lib.malloc_host(C.byref(p), nb))
            ptrs.append(p.value)
        host_out.append(((C.c_void_p * 3)(*ptrs), (C.c_ssize_t * 2)(w * px, ((w + 1) // 2) * px)))

    counts = (C.c_int32 * 19)(*[int(c) for c in s.itx_counts])
    n_mc = len(s.mc_items)

    def submit(d, upload):
        lib.check(lib.frame_submit(d.h, s.n_coefs, counts, n_mc, stages, int(upload)), "frame_submit")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- resident legs: CUDA events around K steps.
    # Each frame context runs on its own stream.  "chained": frame i predicts from the output of frame i - 1 (the
    # context before it in the cycle) and its kernels wait for that frame's last kernel (rb200_frame_depend) -- one
    # video stream, the way real inter frames depend on each other.  "unchained": a static reference and no
    # dependency, i.e. N_CTX independent streams sharing the GPU.  The timed region is bracketed by one start event
    # and one end event per stream; the slowest stream decides.
    main = torch.cuda.Stream()
    one_stream = args.streams == 1           # debugging / ncu launch lists: every context on one CUDA stream
    for d in ctxs:
        lib.check(lib.frame_set_plane_streams(d.h, args.plane_streams))
        lib.check(lib.frame_set_stream(d.h, C.c_void_p(main.cuda_stream) if one_stream else None))
        submit(d, True)                      # batch becomes resident (not timed)
    rstreams = [main] if one_stream else [torch.cuda.ExternalStream(lib.frame_stream(d.h)) for d in ctxs]
    barrier()
    static_ref = lib.Planes()
    chain_refs = []
    if stages & 1:
        lib.check(lib.frame_stage_planes(ctxs[0].ref_handle, 0, C.byref(static_ref)))
        for d in ctxs:
            pl = lib.Planes()
            lib.check(lib.frame_output_planes(d.h, C.byref(pl)))
            chain_refs.append(pl)

    def set_chain(on):
        """Reference slot 0 of context i: the output planes of context i - 1 (chained) or the uploaded static picture."""
        if not stages & 1:
            return
        for i, d in enumerate(ctxs):
            lib.check(lib.frame_set_ref(d.h, 0, C.byref(chain_refs[i - 1] if on else static_ref)))

    def frame(i, chained, upload=False):
        d = ctxs[i % N_CTX]
        if chained and stages & 1:
            lib.check(lib.frame_depend(d.h, ctxs[(i - 1) % N_CTX].h))
        submit(d, upload)
        return d

    def timed_resident(chained):
        set_chain(chained)
        for i in range(args.warmup * FRAMES_PER_STEP):
            frame(i, chained)
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1s = [torch.cuda.Event(enable_timing=True) for _ in rstreams]
        n_launch = 0
        e0.record(rstreams[0])
        for st in rstreams[1:]:
            st.wait_event(e0)
        for i in range(args.steps * FRAMES_PER_STEP):
            n_launch += lib.frame_last_launches(frame(i, chained).h)
        for st, ev in zip(rstreams, e1s):
            ev.record(st)
        barrier()
        ms = max(e0.elapsed_time(ev) for ev in e1s)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), n_launch

    sampler = ClockSampler(local) if rank == 0 else None      # started before the warm-up: sampled every 100 ms from here on
    frames_total = world * args.steps * FRAMES_PER_STEP
    ms_unchained, _ = timed_resident(False)
    value_unchained = frames_total * w * h / (ms_unchained * 1e-3) / 1e6
    if stages & 1:
        ms_max, launches = timed_resident(True)
    else:                                    # post-filters only: nothing predicts from anything, the frames are independent
        ms_max, launches = timed_resident(False)
    value = frames_total * w * h / (ms_max * 1e-3) / 1e6
    clocks = sampler.stop() if sampler else None
    # ---- per-stage kernel durations: with frames on several streams the stage marks of one frame can include other
    # frames' kernels, so the stage table (and the roofline of the dominant kernel) comes from a pass of the same
    # chained frames with every context on ONE stream (kernels strictly serial).
    stage_ms = np.zeros(7)
    stage_n = 0
    buf = (C.c_float * 7)()
    for d in ctxs:
        lib.check(lib.frame_set_stream(d.h, C.c_void_p(main.cuda_stream)))
        lib.check(lib.frame_set_plane_streams(d.h, 0))     # luma and chroma chains on the one stream as well
    for i in range(FRAMES_PER_STEP):
        frame(i, False)
    barrier()
    s0e, s1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_steps = max(1, min(args.steps, 4))
    s0e.record(main)
    for _ in range(stage_steps):
        for i in range(FRAMES_PER_STEP):
            frame(i, False)
        main.synchronize()                   # read the stage marks of this step's last N_CTX frames
        for d in ctxs:
            lib.check(lib.frame_stage_times(d.h, buf))
            stage_ms += np.array(buf[:])
            stage_n += 1
    s1e.record(main)
    barrier()
    value_serial = stage_steps * FRAMES_PER_STEP * w * h / (s0e.elapsed_time(s1e) * 1e-3) / 1e6
    stage_ms = stage_ms / max(stage_n, 1)

    # ---------------- e2e leg: host buffers, the chained stream with N_CTX frames in flight on their own streams
    for d in ctxs:
        lib.check(lib.frame_set_stream(d.h, None))
        lib.check(lib.frame_set_plane_streams(d.h, args.plane_streams))
    streams = [torch.cuda.ExternalStream(lib.frame_stream(d.h)) for d in ctxs]
    if args.coefs in ("gather16", "packed16") and bpc == 8:
        args.coefs = "gather"                # 8-bit pictures carry int16 coefficients anyway
    e2e_upload = {"gather": 3, "zerocopy": 2, "copy": 1, "gather16": 4, "packed16": 5}[args.coefs]
    if args.coefs == "gather16":             # what the front end does when it fills the staging: int16 + escapes
        for d in ctxs:
            lib.check(lib.frame_pack_coef16(d.h, s.n_coefs), "frame_pack_coef16")
    if args.coefs == "packed16":             # ... or one contiguous int16 stream of the blocks' leading columns
        for d in ctxs:
            lib.check(lib.frame_pack_coef_stream(d.h, s.n_coefs, counts, stages), "frame_pack_coef_stream")

    def e2e_frame(i):
        d = ctxs[i % N_CTX]
        d.wait()                             # the context's previous frame (incl. its readback) is done;
        frame(i, True, e2e_upload)           # the front end would refill the pinned staging here
        lib.check(lib.frame_readback_async(d.h, host_out[i % N_CTX][0], host_out[i % N_CTX][1]))

    e2e_frames = args.steps * FRAMES_PER_STEP
    for i in range(max(min(args.warmup * FRAMES_PER_STEP, e2e_frames), 256)):   # >= 0.25 s of traffic: lets the host link and clocks settle
        e2e_frame(i)
    barrier()
    s0 = torch.cuda.Event(enable_timing=True)
    ends = [torch.cuda.Event(enable_timing=True) for _ in ctxs]
    s0.record(streams[0])
    for st in streams[1:]:
        st.wait_event(s0)
    for i in range(e2e_frames):
        e2e_frame(i)
    for st, ev in zip(streams, ends):
        ev.record(st)
    barrier()
    e2e_ms = max(s0.elapsed_time(ev) for ev in ends)
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = frames_total * w * h / (float(t.item()) * 1e-3) / 1e6
    g = ctxs[0].g
    n_sb = g.sb128w * g.sb128h
    h2d = 0
    if stages & 1:
        cs = 4 if bpc > 8 else 2
        if args.coefs != "copy":   # only the leading ncols columns of every block cross PCIe
            from rav1d_b200.lib import TX_DIMS
            sh = np.array([min(TX_DIMS[t][1], 32) for t in range(19)])[s.itx_items["tx"]]
            per_block = s.itx_items["ncols"].astype(np.int64) * sh
            if args.coefs == "packed16":
                h2d += int(((per_block + 7) & ~7).sum()) * 2 + 4 * len(s.itx_items)   # stream (blocks padded to 16 bytes) + offsets
            else:
                h2d += int(per_block.sum()) * (2 if args.coefs == "gather16" else cs)
            if args.coefs in ("gather16", "packed16"):
                h2d += 8 * int((np.abs(s.coef.astype(np.int64)) > 32767).sum())      # escapes (upper bound: whole buffer)
        else:
            h2d += s.n_coefs * cs
        h2d += 16 * (len(s.itx_items) + n_mc + len(getattr(s, "obmc_items", ()))) + 32 * len(getattr(s, "comp_items", ())) \
            + 48 * len(getattr(s, "warp_items", ()))
    if args.lf == "records":
        h2d += 16 * len(s.lf_blocks) + n_sb * 4 + 144 + n_sb * 108
    else:
        h2d += n_sb * 1348 + (g.b4_stride * 32 * g.sb128h + 32) * 4 + 144 + n_sb * 108
    d2h = sum(out_bytes)

    line = None
    if rank == 0:
        peak, peak_src = peaks()
        ab = algorithmic_bytes(w, h, bpc, GEN_ARGS.get(args.workload, {}).get('comp_frac', 0.0))
        if stages & 1:      # the transforms fetch only the leading ncols columns of a block (Rb200ItxItem.ncols)
            from rav1d_b200.lib import TX_DIMS
            sh_ = np.array([min(TX_DIMS[t][1], 32) for t in range(19)])[s.itx_items["tx"]]
            ab["itx"] = int((s.itx_items["ncols"].astype(np.int64) * sh_).sum()) * (4 if bpc > 8 else 2) + 2 * ab["S"]
        per_stage = {}
        for name, t_ms in zip(STAGE_NAMES, stage_ms):
            bit = {"mc": 1, "itx": 1, "deblock": 2, "cdef": 4, "lr": 8, "film_grain": 16}.get(name, 0)
            if t_ms <= 0 or not stages & bit:
                continue
            per_stage[name] = {"ms": round(float(t_ms), 4), "algorithmic_bytes": ab[name],
                               "gbs": round(ab[name] / (t_ms * 1e-3) / 1e9, 1)}
        dom = max(per_stage, key=lambda k: per_stage[k]["ms"]) if per_stage else None
        roofline = None
        if dom:
            ach = per_stage[dom]["gbs"]
            roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s",
                        "frac": round(ach / peak, 4), "traffic": TRAFFIC.get((args.workload, dom)),
                        "peak_source": peak_src, "share_of_step": round(per_stage[dom]["ms"] / sum(v["ms"] for v in per_stage.values()), 3)}
        issue = None
        if args.workload in WARP_INST_PER_FRAME and clocks and clocks.get("sm_mhz"):
            peak_issue = 148 * 4 * clocks["sm_mhz"] * 1e6          # warp instructions / s
            wi = WARP_INST_PER_FRAME[args.workload]
            issue = {"warp_inst_per_frame": wi, "peak_warp_inst_per_s": peak_issue, "source": "profiles/r02i_ncu_full_summary_4k10.csv" if args.workload == "4k10" else "profiles/r02i_ncu_full_summary_4k10c5.csv",
                     "frac": round(wi * (value * 1e6 / (w * h)) / world / peak_issue, 4)}
        # frame-level algorithmic bytes as BASELINE.md 4 counts them (MC + itx = one fused recon stage: 2S + C)
        frame_bytes = (ab["recon"] if stages & 1 else 0) + sum(ab[k] for k in ("deblock", "cdef", "lr", "film_grain") if k in per_stage)
        line = {"metric": metric_name(args.workload), "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "int32", "data": "synthetic",
                "config": {"workload": desc, "frames_per_step": FRAMES_PER_STEP, "width": w, "height": h, "bpc": bpc,
                           "l2": f"inputs larger than L2: {N_CTX} frame contexts cycled, > {ab['S'] * 4 // 1000000} MB working set each",
                           "frames_in_flight": N_CTX,
                           "plane_streams": "luma and chroma post-filter chains of a frame on two CUDA streams" if args.plane_streams else "off",
                           "dependency": "every frame predicts from the previous frame's output and waits for it on the device (rb200_frame_depend)"
                                         if stages & 1 else "none (post-filters only)",
                           "lf_metadata": "block records, masks built on the device" if args.lf == "records" else "masks and levels uploaded",
                           "parallelism": f"{world} independent video streams (one per GPU)" if world > 1 else "1 video stream"},
                "fps": value * 1e6 / (w * h),
                "value_unchained": value_unchained,
                "value_unchained_note": f"{N_CTX} independent streams per GPU: same contexts, static reference, no frame-to-frame dependency",
                "value_serial_kernels": value_serial,
                "stage_timing": "CUDA-event marks of a one-CUDA-stream pass over the same frames (kernels strictly serial)",
                "frame_roofline_frac": round(frame_bytes * (value * 1e6 / (w * h)) / world / (peak * 1e9), 4),
                "stages": per_stage, "roofline": roofline, "issue": issue, "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "Mpixel/s", "h2d_bytes_per_step": h2d * FRAMES_PER_STEP,
                        "d2h_bytes_per_step": d2h * FRAMES_PER_STEP,
                        "coefficients": {"gather": "gather kernel over pinned host memory, column-bounded", "copy": "H2D copy",
                                         "gather16": "int16 transport + escape list: gather kernel over pinned host memory, column-bounded, widened on the device",
                                         "packed16": "one contiguous int16 stream of the blocks' non-zero columns + escape list: DMA copy, spread into the int32 layout on the device",
                                         "zerocopy": "transforms read pinned host memory, column-bounded"}[args.coefs]},
                "gpu_launches": launches}
    for d in ctxs:
        d.close()
    # BASELINE configs[3]: one 8K 10-bit picture's post-filters split by superblock rows over the N GPUs of this job, halo rows
    # pulled peer to peer over NVLink, ordered by cross-GPU flags (no collective) -- reported beside the headline at N > 1
    band = None
    if world > 1 and args.workload == "4k10" and not args.no_band_split:
        from rav1d_b200 import multigpu
        torch.cuda.synchronize()
        dist.barrier()
        band = multigpu.run_band_split(7680, 4320, 10, steps=40, in_flight=2, check=False)
        if line is not None:
            line["band_split_8k10"] = band
    if world > 1:
        dist.barrier()
    if rank == 0:
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            t1 = cpu_frames(s, stages, 1, cores)
            n = max(2, min(64, int(12.0 / max(t1, 1e-3))))
            tt = cpu_frames(s, stages, n, cores)
            line["cpu_baseline"] = {"value": n * w * h / tt / 1e6, "unit": "Mpixel/s", "cores": cores, "kind": "reference",
                                    "sample": f"{n} frames of the workload, {cores} threads, reference C DSP (no asm), {cpu_model()}"}
        else:
            line["cpu_baseline"] = None
        emit(line)
    if world > 1:
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full`
# capture of the dominant kernel (profiles/), keyed by (workload, stage); None until captured.
TRAFFIC = {
    # profiles/r02i_ncu_full_summary_4k10.csv: cdef_dir_frame_kernel 13.9 MB + cdef_filter_tma_kernel 17.7 (luma) + 9.4 (chroma) MB
    ("4k10", "cdef"): 40.95e6,
    ("4k10", "mc"): 26.71e6,
    ("4k10", "lr"): 27.32e6,       # 17.79 + 4.76 + 4.77 MB (three planes)
    # profiles/r02i_ncu_full_summary_4k10c5.csv: mc_batch 25.5 + mc_comp_batch 55.3 + warp 5.8 + obmc 8.4 + 13.9 MB
    ("4k10c5", "mc"): 108.9e6,
}
# Executed warp instructions per frame (smsp__inst_executed.sum over one frame's launches, same captures): the path is
# integer-issue bound, so value x this / (SMs x 4 schedulers x clock) says how full the issue slots are.
WARP_INST_PER_FRAME = {"4k10": 164.3e6, "4k10c5": 267.8e6}   # r02i captures (round 1: 209 M / 293.9 M)


_JSON_OUT = None


def emit(line):
    """The one JSON line, on the process's original stdout."""
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # stdout carries exactly one JSON line: everything else that writes to fd 1 (NCCL prints its version banner there
    # from C) is sent to stderr, the line itself goes to a duplicate of the original descriptor.
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="4k10", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--lf", default="masks", choices=["masks", "records"],
                    help="loop-filter metadata: Av1Filter masks + levels uploaded, or per-block records uploaded and the masks built on the device")
    ap.add_argument("--streams", type=int, default=N_CTX, help="resident legs: 1 = every frame context on one CUDA stream (launch lists), otherwise one stream per context")
    ap.add_argument("--no-band-split", action="store_true", help="N > 1: skip the 8K band-split leg (BASELINE configs[3])")
    ap.add_argument("--plane-streams", type=int, default=1, help="1 (default): luma and chroma post-filter chains of a frame on two streams; 0: one stream")
    ap.add_argument("--coefs", default="packed16", choices=["packed16", "gather16", "gather", "zerocopy", "copy"],
                    help="e2e leg, how coefficients cross PCIe: as one contiguous int16 stream of the blocks' non-zero columns + escapes, "
                         "DMA-copied and spread out on the device (default for 16-bit pictures); as int16 + escapes in the block layout, a "
                         "gather kernel pulling each block's non-zero columns into HBM and widening them; the same gather on the int32 staging; the "
                         "transforms read pinned memory directly; or the whole buffer is H2D-copied")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    rank = int(os.environ.get("RANK", 0))
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        if rank != 0:
            return
        os.environ["RB200_LIB_TYPES_ONLY"] = "1"   # the generator needs the record types only: the product library is not mapped
    from rav1d_b200.synth import framegen
    s = framegen.generate(wl[0], wl[1], wl[2], seed=1 + (rank if args.impl == "b200" else 0), **GEN_ARGS.get(args.workload, {}))
    if wl[3] & 16:
        s.film_grain = framegen.random_film_grain(np.random.default_rng(7), lag=3, overlap=1)
    if args.impl == "b200" and args.plane_streams and wl[3] & 1:
        framegen.sort_luma_first(s)          # what a front end does for free: two lists instead of one (rb200_frame_set_plane_counts)
    if args.impl == "reference":
        run_reference(args, s, wl)
    else:
        run_gpu(args, s, wl)


if __name__ == "__main__":
    main()
