"""One picture's post-filters split over several GPUs by superblock rows (BASELINE config 4,
SURVEY 8e).

The picture is cut at boundaries of the 64-row loop-restoration stripe grid (stripe s = luma rows
64 s - 8 .. 64 s + 55, src/lr_apply.rs:47-54).  Every rank owns the reconstructed rows of its
stripes, pulls the halo its earlier stages read (rb200_frame_band_rows) from the ranks that own
those rows -- a peer-to-peer copy over NVLink, no collective -- and then runs deblock, CDEF and
loop restoration on its band only.  Ranks are one process per GPU; peers' plane memory is mapped
with CUDA IPC handles that travel through torch.distributed (any backend: object all-gather).

Ordering between GPUs (deblocking is in place, so a rank may not start it while a neighbour still reads
its rows): a flag per (context, peer) in the waiting GPU's memory.  A rank that has pulled its halo out of
peer p writes the frame's epoch into p's flag through the peer mapping (rb200_flag_signal, stream-ordered
behind the copies); before its own filters it waits for the flags of the ranks that pull from it
(rb200_flag_wait).  No host synchronisation, no collective; BandRing keeps several pictures in flight.
"""
import ctypes as C

from . import lib


def split_stripes(height, n_ranks):
    """Stripe ranges [(s0, s1)] per rank, as even as the stripe grid allows; ranks beyond the
    number of stripes get an empty range."""
    n_stripes = (height + 8 + 63) // 64
    base, extra = divmod(n_stripes, n_ranks)
    out, s = [], 0
    for r in range(n_ranks):
        k = base + (1 if r < extra else 0)
        out.append((s, s + k))
        s += k
    return out


def owned_rows(height, s0, s1):
    """Luma rows delivered (and, as input, owned) by stripes [s0, s1)."""
    if s1 <= s0:
        return (0, 0)
    n_stripes = (height + 8 + 63) // 64
    lo = max(64 * s0 - 8, 0)
    hi = height if s1 >= n_stripes else 64 * s1 - 8
    return (lo, hi)


def halo_plan(height, ranges, rank, in_rows, padded_height=None):
    """[(peer, row_begin, row_end)]: which rows `rank` must pull from which owner to cover in_rows.
    Rows at or below the picture height that lie in the allocation padding belong to the last owner."""
    lo, hi = in_rows
    own = owned_rows(height, *ranges[rank])
    plan = []
    last = max(r for r in range(len(ranges)) if ranges[r][1] > ranges[r][0])
    for peer, (s0, s1) in enumerate(ranges):
        if peer == rank or s1 <= s0:
            continue
        p_lo, p_hi = owned_rows(height, s0, s1)
        if peer == last and padded_height:
            p_hi = padded_height
        a, b = max(lo, p_lo), min(hi, p_hi)
        if b > a and not (a >= own[0] and b <= own[1]):
            plan.append((peer, a, b))
    return plan


def exchange_bytes(payload: bytes):
    """All-gather one bytes object per rank over torch.distributed (works on gloo and nccl)."""
    import torch.distributed as dist
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, payload)
    return out


class BandContext:
    """One rank's share of a picture: a frame context restricted to its stripes."""

    def __init__(self, hdr, ranges, rank, frame_handle=None, max_coefs=1, max_itx=1, max_mc=1):
        self.hdr, self.ranges, self.rank = hdr, ranges, rank
        self.h = frame_handle or C.c_void_p()
        if not frame_handle:
            lib.check(lib.frame_create(C.byref(self.h), C.byref(hdr), max_coefs, max_itx, max_mc), "frame_create")
        s0, s1 = ranges[rank]
        self.empty = s1 <= s0
        lib.check(lib.frame_set_band(self.h, s0, s1 if not self.empty else s0), "frame_set_band")
        v = [C.c_int() for _ in range(4)]
        lib.check(lib.frame_band_rows(self.h, *[C.byref(x) for x in v]))
        self.in_rows = (v[0].value, v[1].value)
        self.out_rows = (v[2].value, v[3].value)
        self.own_rows = owned_rows(hdr.height, s0, s1)
        g = lib.FrameGeometry()
        lib.check(lib.frame_geometry(self.h, C.byref(g)))
        self.geom = g
        self.peer_bases = {}

    def plane_block(self, which=0):
        base, size = C.c_void_p(), C.c_size_t()
        lib.check(lib.frame_plane_block(self.h, which, C.byref(base), C.byref(size)))
        return base.value, size.value

    def ipc_handle(self, which=0):
        h = (C.c_uint8 * 64)()
        lib.check(lib.ipc_get_handle(C.c_void_p(self.plane_block(which)[0]), h), "ipc_get_handle")
        return bytes(h)

    def open_peers(self, handles):
        """handles[r]: IPC handle bytes of rank r's plane block (from exchange_bytes)."""
        for peer, _, _ in self.plan():
            if peer not in self.peer_bases:
                buf = (C.c_uint8 * 64).from_buffer_copy(handles[peer])
                p = C.c_void_p()
                lib.check(lib.ipc_open_handle(buf, C.byref(p)), "ipc_open_handle")
                self.peer_bases[peer] = p.value

    def plan(self):
        if self.empty:
            return []
        return halo_plan(self.hdr.height, self.ranges, self.rank, self.in_rows, self.geom.plane_h[0])

    def upload_owned(self, planes, which=0):
        """Host picture -> this rank's own rows (plus, on the last band, the allocation padding the
        8-pixel-aligned filters may read)."""
        if self.empty:
            return
        data = (C.c_void_p * 3)(*[p.ctypes.data for p in planes])
        strides = (C.c_ssize_t * 2)(planes[0].strides[0], planes[1].strides[0])
        lib.check(lib.frame_upload_rows(self.h, which, data, strides, self.own_rows[0], self.own_rows[1]), "upload_rows")

    def pull_halo(self, which=0):
        """Queue the peer-to-peer copies of the halo rows on the frame's stream (asynchronous)."""
        n = 0
        for peer, a, b in self.plan():
            lib.check(lib.frame_pull_rows(self.h, which, C.c_void_p(self.peer_bases[peer]), a, b), "pull_rows")
            n += 1
        return n

    def halo_bytes(self):
        g = self.geom
        total = 0
        for _, a, b in self.plan():
            total += (b - a) * g.stride[0]
            if g.n_planes > 1:
                total += 2 * (((b + g.ss_ver) >> g.ss_ver) - (a >> g.ss_ver)) * g.stride[1]
        return total

    def readback_owned(self, out_planes):
        if self.empty:
            return
        data = (C.c_void_p * 3)(*[p.ctypes.data for p in out_planes])
        strides = (C.c_ssize_t * 2)(out_planes[0].strides[0], out_planes[1].strides[0])
        lib.check(lib.frame_readback_rows(self.h, data, strides, self.out_rows[0], self.out_rows[1]), "readback_rows")

    def close(self):
        for p in self.peer_bases.values():
            lib.ipc_close_handle(C.c_void_p(p))
        self.peer_bases = {}
        if self.h:
            lib.frame_destroy(self.h)
            self.h = None


def pullers_of(height, ranges, rank, in_rows_of, padded_height=None):
    """Ranks whose halo plan names `rank` (they read its rows, so it waits for them before filtering in place)."""
    return [r for r in range(len(ranges)) if r != rank and ranges[r][1] > ranges[r][0] and
            any(peer == rank for peer, _, _ in halo_plan(height, ranges, r, in_rows_of[r], padded_height))]


class BandRing:
    """K pictures in flight on one rank of a band split: K BandContexts with their flags, the IPC exchange and the
    per-picture protocol pull -> signal -> wait -> filter.  `make_frame(k)` returns a DeviceFrame-like object (attributes
    h, submit, wait) whose batch is loaded; `dist` is torch.distributed (initialised)."""

    def __init__(self, hdr, world, rank, make_frame, in_flight=2):
        import numpy as np
        self.world, self.rank, self.k = world, rank, in_flight
        self.ranges = split_stripes(hdr.height, world)
        self.frames = [make_frame(i) for i in range(in_flight)]
        self.bands = [BandContext(hdr, self.ranges, rank, frame_handle=f.h) for f in self.frames]
        # in_rows of every rank (the plan is a pure function of the geometry, but band_rows() lives in the library)
        mine = list(self.bands[0].in_rows)
        self.in_rows_of = [tuple(x) for x in _all_gather(mine)]
        self.epoch = [0] * in_flight
        # flags: uint32[world] per context, in this GPU's memory
        self.flags = []
        for _ in range(in_flight):
            p = C.c_void_p()
            lib.check(lib.malloc(C.byref(p), 4 * max(world, 64)), "malloc(flags)")
            lib.check(lib.memset(p, 0, 4 * max(world, 64), None))
            self.flags.append(p.value)
        lib.check(lib.stream_sync(None))
        payload = []
        for i in range(in_flight):
            h = (C.c_uint8 * 64)()
            lib.check(lib.ipc_get_handle(C.c_void_p(self.flags[i]), h), "ipc_get_handle(flags)")
            payload.append((self.bands[i].ipc_handle(), bytes(h)))
        everyone = _all_gather(payload)
        self.peer_flags = [dict() for _ in range(in_flight)]
        b0 = self.bands[0]
        self.pull_from = sorted({peer for peer, _, _ in b0.plan()})
        self.pulled_by = pullers_of(hdr.height, self.ranges, rank, self.in_rows_of, b0.geom.plane_h[0])
        for i in range(in_flight):
            self.bands[i].open_peers([everyone[r][i][0] for r in range(world)])
            for peer in self.pull_from:
                buf = (C.c_uint8 * 64).from_buffer_copy(everyone[peer][i][1])
                q = C.c_void_p()
                lib.check(lib.ipc_open_handle(buf, C.byref(q)), "ipc_open_handle(flags)")
                self.peer_flags[i][peer] = q.value

    def step(self, i, stages=14):
        """One picture on context i: pull the halo, tell the owners, wait for the ranks that read this rank's rows, filter."""
        band, f = self.bands[i], self.frames[i]
        self.epoch[i] += 1
        e = self.epoch[i]
        st = C.c_void_p(lib.frame_stream(f.h))
        band.pull_halo()
        for peer in self.pull_from:
            lib.check(lib.flag_signal(st, C.c_void_p(self.peer_flags[i][peer] + 4 * self.rank), e), "flag_signal")
        for peer in self.pulled_by:
            lib.check(lib.flag_wait(st, C.c_void_p(self.flags[i] + 4 * peer), e), "flag_wait")
        if not band.empty:
            f.submit(stages, upload=False)

    def close(self):
        for i in range(self.k):
            for q in self.peer_flags[i].values():
                lib.ipc_close_handle(C.c_void_p(q))
            self.bands[i].close()
            self.frames[i].h = None
            lib.free(C.c_void_p(self.flags[i]))


def _all_gather(obj):
    import torch.distributed as dist
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def run_band_split(w, h, bpc, steps=10, in_flight=2, check=True, seed=4):
    """The post-filters (deblock + CDEF + LR) of one w x h picture split over the ranks of the current torch.distributed
    job (one process per GPU, NCCL or gloo for the set-up only).  Returns, on every rank, a dict with the time per picture
    (max over ranks, CUDA events), the halo bytes and -- with check -- whether every band equals the oracle's
    whole-picture result bit for bit (rank 0 runs the oracle)."""
    import os
    import sys
    import numpy as np
    import torch
    import torch.distributed as dist
    from .synth import framegen
    rank, world = dist.get_rank(), dist.get_world_size()
    s = framegen.generate(w, h, bpc, seed=seed)          # same seed on every rank: the same picture
    start = framegen.recon_input_planes(s)

    def make_frame(_):
        d = framegen.DeviceFrame(s)
        d.load_batch()
        return d

    ring = BandRing(s.hdr, world, rank, make_frame, in_flight)
    for i in range(in_flight):
        ring.bands[i].upload_owned(start)
        ring.frames[i].submit(14, upload=True)          # metadata resident (also a warm-up of the band's kernels)
        ring.frames[i].wait()
        ring.bands[i].upload_owned(start)
    torch.cuda.synchronize(); dist.barrier()             # every rank's own rows are resident
    streams = [torch.cuda.ExternalStream(lib.frame_stream(f.h)) for f in ring.frames]
    for n in range(2 * in_flight):
        ring.step(n % in_flight)
    torch.cuda.synchronize(); dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = [torch.cuda.Event(enable_timing=True) for _ in streams]
    e0.record(streams[0])
    for st in streams[1:]:
        st.wait_event(e0)
    for n in range(steps):
        ring.step(n % in_flight)
    for st, ev in zip(streams, e1):
        ev.record(st)
    torch.cuda.synchronize(); dist.barrier()
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([max(e0.elapsed_time(ev) for ev in e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    ok = None
    if check:
        # the deblock stage is in place: restore the inputs, run one picture, compare this band
        ring.bands[0].upload_owned(start)
        torch.cuda.synchronize(); dist.barrier()
        ring.step(0); ring.frames[0].wait()
        out = [np.zeros_like(p) for p in s.ref]
        ring.bands[0].readback_owned(out)
        lst = [None]
        if rank == 0:
            root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
            sys.path.insert(0, os.path.join(root, "tests"))
            import framecheck, refharness            # test infrastructure: the checker, not the product path
            exp = framecheck.oracle_frame(refharness.load(), s, 14, n_tc=os.cpu_count() or 1, start_planes=start)
            lst = [[e.copy() for e in exp]]
        dist.broadcast_object_list(lst, src=0)
        exp = lst[0]
        lo, hi = ring.bands[0].out_rows
        good = True
        for p in range(3):
            a, b = (lo, hi) if p == 0 else (lo >> 1, (hi + 1) >> 1)
            good &= bool(np.array_equal(exp[p][a:b], out[p][a:b, :exp[p].shape[1]]))
        g = torch.tensor([1 if good else 0], device=dev)
        dist.all_reduce(g, op=dist.ReduceOp.MIN)
        ok = bool(g.item())
        torch.cuda.synchronize(); dist.barrier()
    hb = torch.tensor([float(ring.bands[0].halo_bytes())], dtype=torch.float64, device=dev)
    dist.all_reduce(hb, op=dist.ReduceOp.MAX)
    ring.close()
    return {"what": "post-filters (deblock + CDEF + LR) of one picture split by stripe rows over the GPUs, halo rows pulled peer to peer, "
                    "flag-ordered (no collective)", "width": w, "height": h, "bpc": bpc, "n_gpus": world, "pictures_in_flight": in_flight,
            "ms_per_frame": ms, "mpixel_per_s": w * h / ms / 1e3, "max_halo_bytes_per_rank": hb.item(), "bit_exact_vs_oracle": ok}
