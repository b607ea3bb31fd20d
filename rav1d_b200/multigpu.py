"""One picture's post-filters split over several GPUs by superblock rows (BASELINE config 4,
SURVEY 8e).

The picture is cut at boundaries of the 64-row loop-restoration stripe grid (stripe s = luma rows
64 s - 8 .. 64 s + 55, src/lr_apply.rs:47-54).  Every rank owns the reconstructed rows of its
stripes, pulls the halo its earlier stages read (rb200_frame_band_rows) from the ranks that own
those rows -- a peer-to-peer copy over NVLink, no collective -- and then runs deblock, CDEF and
loop restoration on its band only.  Ranks are one process per GPU; peers' plane memory is mapped
with CUDA IPC handles that travel through torch.distributed (any backend: object all-gather).
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
