"""Reader of oracle/_ref/ref_dump output: per decoded frame of a real AV1 stream, the picture
before the in-loop filters, the decoder's own filter metadata and the picture after deblock +
CDEF + loop restoration (see oracle/ref_dump.c).  Test infrastructure."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DUMP = os.path.join(ROOT, "oracle", "_ref", "ref_dump")
REF_DATA = "/root/reference/tests/dav1d-test-data"
N_HDR = 45


class StreamFrame:
    """Duck-types rav1d_b200.synth.framegen.SynthFrame for the filter stages."""


def available():
    return os.path.exists(REF_DUMP) and os.path.isdir(REF_DATA)


def dump(path, max_frames=4, sr=False):
    """Decode `path` with the reference decoder and return the list of StreamFrames.
    sr: only the frames coded with super-resolution (pre = coded width, post = upscaled width)."""
    from rav1d_b200 import lib
    with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as tf:
        out = tf.name
    try:
        r = subprocess.run([REF_DUMP, path, out, str(max_frames)] + (["0", "1"] if sr else []), capture_output=True, text=True)
        if r.returncode != 0 or not os.path.getsize(out):
            return []
        buf = np.fromfile(out, dtype=np.uint8)
    finally:
        os.unlink(out)
    return parse(buf)


def parse(buf):
    from rav1d_b200 import lib
    frames, off = [], 0
    while off + 4 * N_HDR <= buf.size:
        hd = buf[off:off + 4 * N_HDR].view("<i4")
        off += 4 * N_HDR
        assert hd[0] in (0x52423230, 0x52425352), "bad dump magic"
        is_sr = hd[0] == 0x52425352
        (idx, w, h, bpc, layout, sb128, ly0, ly1, lu, lv, cdef_on, damping) = (int(v) for v in hd[1:13])
        ystr, uvstr = [int(v) for v in hd[13:21]], [int(v) for v in hd[21:29]]
        lr_type = [int(v) for v in hd[29:32]]
        unit = [int(v) for v in hd[32:34]]
        tiles = (int(hd[34]), int(hd[35]))
        b4_stride, sb128w, sb128h, stride_y, stride_uv, ah, n_planes, frame_type, show = (int(v) for v in hd[36:45])
        n_sb = sb128w * sb128h
        out_w, out_stride_y, out_stride_uv, sr_sb128w = w, stride_y, stride_uv, sb128w
        if is_sr:
            ex = [int(v) for v in buf[off:off + 32].view("<i4")]
            off += 32
            out_w, out_stride_y, out_stride_uv, sr_sb128w = ex[:4]
        s = StreamFrame()
        s.out_w, s.resize = out_w, (tuple(ex[4:8]) if is_sr else None)
        s.index, s.tiles, s.layout, s.cdef_on, s.frame_type = idx, tiles, layout, cdef_on, frame_type
        s.w, s.h, s.bpc, s.bdmax = w, h, bpc, (1 << bpc) - 1
        s.aw, s.ah = (w + 127) & ~127, ah
        hdr = lib.FrameHeader()
        hdr.width, hdr.height, hdr.bpc, hdr.layout, hdr.sb128 = w, h, bpc, layout, sb128
        hdr.lf_level_y[0], hdr.lf_level_y[1], hdr.lf_level_u, hdr.lf_level_v = ly0, ly1, lu, lv
        hdr.cdef_damping = damping
        for i in range(8):
            hdr.cdef_y_strength[i], hdr.cdef_uv_strength[i] = ystr[i], uvstr[i]
        for i in range(3):
            hdr.lr_type[i] = lr_type[i]
        hdr.lr_unit_size_log2[0], hdr.lr_unit_size_log2[1] = unit
        hdr.upscaled_width = out_w if is_sr else 0
        s.hdr = hdr
        s.masks = buf[off:off + n_sb * 1348].view(lib.AV1_FILTER_DT).copy(); off += n_sb * 1348
        nl = b4_stride * 32 * sb128h
        s.levels = buf[off:off + nl * 4].reshape(32 * sb128h, b4_stride, 4).copy(); off += nl * 4
        s.lut = lib.FilterLUT.from_buffer_copy(buf[off:off + 144].tobytes()); off += 144
        n_lr = sr_sb128w * sb128h
        s.lr_masks = buf[off:off + n_lr * 108].view(lib.AV1_RESTORATION_DT).copy(); off += n_lr * 108
        pdt = np.uint16 if bpc > 8 else np.uint8
        px = 2 if bpc > 8 else 1
        ss_ver, ss_hor = int(layout == 1), int(layout != 3)

        def planes(sy, suv, aw):
            nonlocal off
            out = []
            for pl in range(n_planes):
                stride = suv if pl else sy
                rows = ah >> ss_ver if pl else ah
                cols = aw >> ss_hor if pl else aw
                a = buf[off:off + stride * rows].reshape(rows, stride)[:, :cols * px].view(pdt).copy()
                off += stride * rows
                out.append(a)
            return out
        s.pre = planes(stride_y, stride_uv, s.aw)
        s.post = planes(out_stride_y, out_stride_uv, (out_w + 127) & ~127)
        s.ref = s.pre
        s.readback_like = s.post                       # shapes for DeviceFrame.readback
        g = lib.FrameGeometry()
        g.bw = ((w + 7) >> 3) << 1; g.bh = ((h + 7) >> 3) << 1
        g.w4 = (w + 3) >> 2; g.h4 = (h + 3) >> 2
        g.sb128w, g.sb128h, g.b4_stride = sb128w, sb128h, b4_stride
        g.ss_hor, g.ss_ver, g.n_planes = ss_hor, ss_ver, n_planes
        s.geom = g
        # no reconstruction batch: the pre-filter picture is the input
        s.n_coefs = 0
        s.coef = np.zeros(0, np.int32 if bpc > 8 else np.int16)
        s.itx_items = np.zeros(0, lib.ITX_ITEM_DT); s.mc_items = np.zeros(0, lib.MC_ITEM_DT)
        s.itx_counts = np.zeros(19, np.int32)
        # the stages the decoder ran for this frame
        s.stages = (2 if (ly0 or ly1) else 0) | (4 if cdef_on else 0) | (8 if any(lr_type) else 0) | (32 if is_sr else 0)
        frames.append(s)
    return frames


def dump_grain(path, max_frames=2):
    """Film-grain records of `path` decoded with apply_grain = 1: [dict(w, h, bpc, layout, is_id, fg (bytes), inp, out)]."""
    with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as tf:
        out = tf.name
    try:
        r = subprocess.run([REF_DUMP, path, out, "0", str(max_frames)], capture_output=True, text=True)
        if r.returncode != 0 or not os.path.getsize(out):
            return []
        buf = np.fromfile(out, dtype=np.uint8)
    finally:
        os.unlink(out)
    recs, off = [], 0
    while off + 36 <= buf.size:
        hd = buf[off:off + 36].view("<i4"); off += 36
        assert hd[0] == 0x52424647, "bad film-grain record"
        w, h, bpc, layout, is_id, stride_y, stride_uv, fgsz = (int(v) for v in hd[1:9])
        fg = buf[off:off + fgsz].tobytes(); off += fgsz
        ss_ver, ss_hor = int(layout == 1), int(layout != 3)
        n_planes = 1 if layout == 0 else 3
        pdt, px = (np.uint16, 2) if bpc > 8 else (np.uint8, 1)
        pics = []
        for _ in range(2):
            pl = []
            for p in range(n_planes):
                stride = stride_uv if p else stride_y
                rows = (h + ss_ver) >> ss_ver if p else h
                cols = (w + ss_hor) >> ss_hor if p else w
                pl.append(buf[off:off + stride * rows].reshape(rows, stride)[:, :cols * px].view(pdt).copy())
                off += stride * rows
            pics.append(pl)
        recs.append(dict(w=w, h=h, bpc=bpc, layout=layout, is_id=is_id, fg=np.frombuffer(fg, np.uint8).copy(), inp=pics[0], out=pics[1]))
    return recs


GOLDEN = os.path.join(ROOT, "tests", "golden", "streams.npz")
GOLDEN_GRAIN = os.path.join(ROOT, "tests", "golden", "film_grain.npz")
GOLDEN_SR = os.path.join(ROOT, "tests", "golden", "streams_sr.npz")
GOLDEN_SIZES = os.path.join(ROOT, "tests", "golden", "streams_sizes.npz")
GOLDEN_LFB = os.path.join(ROOT, "tests", "golden", "streams_lfb.npz")


def dump_lfb(path, max_frames=4):
    """{frame index: Rb200LfBlock records} -- the arguments of every dav1d_create_lf_mask_{intra,inter} call the reference
    decoder made for the frame, in decode order (oracle/ref_dump.c, RB200_DUMP_LFB).  Frames without deblocking have none."""
    from rav1d_b200 import lib
    with tempfile.TemporaryDirectory() as td:
        out, lfb = os.path.join(td, "o.bin"), os.path.join(td, "lfb.bin")
        r = subprocess.run([REF_DUMP, path, out, str(max_frames)], capture_output=True, text=True, env=dict(os.environ, RB200_DUMP_LFB=lfb))
        if r.returncode != 0 or not os.path.exists(lfb):
            return {}
        buf = np.fromfile(lfb, dtype=np.uint8)
    res, off = {}, 0
    while off + 12 <= buf.size:
        magic, idx, n = (int(v) for v in buf[off:off + 12].view("<i4"))
        assert magic == 0x52424c42, "bad lf-block dump magic"
        off += 12
        res[idx] = buf[off:off + 16 * n].view(lib.LF_BLOCK_DT).copy()
        off += 16 * n
    return res


def load_golden_lfb(path=GOLDEN_LFB):
    """{key: records} for the frames of tests/golden/streams.npz that were deblocked."""
    from rav1d_b200 import lib
    z = np.load(path, allow_pickle=False)
    return {str(k): z[str(k)].view(lib.LF_BLOCK_DT).reshape(-1) for k in z["index"]}


def load_golden(path=GOLDEN):
    """[(key, StreamFrame)] from tests/golden/streams*.npz (tools/make_stream_fixtures.py)."""
    from rav1d_b200 import lib
    z = np.load(path, allow_pickle=False)
    out = []
    for key in z["index"]:
        key = str(key)
        it = [int(v) for v in z[f"{key}/ints"]]
        s = StreamFrame()
        (s.w, s.h, s.bpc, s.layout, sb128, ly0, ly1, lu, lv, s.cdef_on, damping) = it[:11]
        ystr, uvstr, lr_type, unit, tiles, s.ah, s.stages = it[11:19], it[19:27], it[27:30], it[30:32], it[32:34], it[34], it[35]
        s.tiles, s.bdmax, s.aw, s.index = tuple(tiles), (1 << s.bpc) - 1, (s.w + 127) & ~127, int(key.rsplit("#", 1)[1])
        s.out_w = it[36] if len(it) > 36 else s.w
        hdr = lib.FrameHeader()
        hdr.width, hdr.height, hdr.bpc, hdr.layout, hdr.sb128 = s.w, s.h, s.bpc, s.layout, sb128
        hdr.lf_level_y[0], hdr.lf_level_y[1], hdr.lf_level_u, hdr.lf_level_v = ly0, ly1, lu, lv
        hdr.cdef_damping = damping
        for i in range(8):
            hdr.cdef_y_strength[i], hdr.cdef_uv_strength[i] = ystr[i], uvstr[i]
        for i in range(3):
            hdr.lr_type[i] = lr_type[i]
        hdr.lr_unit_size_log2[0], hdr.lr_unit_size_log2[1] = unit
        hdr.upscaled_width = s.out_w if s.out_w != s.w else 0
        s.hdr = hdr
        s.masks = z[f"{key}/masks"].view(lib.AV1_FILTER_DT)
        s.levels = z[f"{key}/levels"]
        s.lut = lib.FilterLUT.from_buffer_copy(z[f"{key}/lut"].tobytes())
        s.lr_masks = z[f"{key}/lr_masks"].view(lib.AV1_RESTORATION_DT)
        ss_ver, ss_hor = int(s.layout == 1), int(s.layout != 3)
        n_planes = 1 if s.layout == 0 else 3
        pdt = np.uint16 if s.bpc > 8 else np.uint8

        def planes(kind):
            res = []
            aw = s.aw if kind == "pre" else (s.out_w + 127) & ~127
            for p in range(n_planes):
                a = np.zeros((s.ah >> ss_ver if p else s.ah, aw >> ss_hor if p else aw), pdt)
                c = z[f"{key}/{kind}{p}"]
                a[:c.shape[0], :c.shape[1]] = c
                res.append(a)
            return res
        s.pre, s.post = planes("pre"), planes("post")
        s.ref = s.pre
        s.readback_like = s.post
        g = lib.FrameGeometry()
        g.bw = ((s.w + 7) >> 3) << 1; g.bh = ((s.h + 7) >> 3) << 1
        g.w4 = (s.w + 3) >> 2; g.h4 = (s.h + 3) >> 2
        g.sb128w = (g.bw + 31) >> 5; g.sb128h = (g.bh + 31) >> 5; g.b4_stride = (g.bw + 31) & ~31
        g.ss_hor, g.ss_ver, g.n_planes = ss_hor, ss_ver, n_planes
        s.geom = g
        s.n_coefs = 0
        s.coef = np.zeros(0, np.int32 if s.bpc > 8 else np.int16)
        s.itx_items = np.zeros(0, lib.ITX_ITEM_DT); s.mc_items = np.zeros(0, lib.MC_ITEM_DT)
        s.itx_counts = np.zeros(19, np.int32)
        out.append((key, s))
    return out


def visible(s, planes, out=False):
    """Visible part of each plane of a picture of StreamFrame `s` (out: the picture after super-resolution)."""
    ss_ver, ss_hor = int(s.layout == 1), int(s.layout != 3)
    w = getattr(s, "out_w", s.w) if out else s.w
    return [a[:(s.h + ss_ver) >> ss_ver if p else s.h, :(w + ss_hor) >> ss_hor if p else w] for p, a in enumerate(planes)]
