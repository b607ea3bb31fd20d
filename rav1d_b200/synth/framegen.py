"""Synthetic inter frame for the recon + post-filter path (BASELINE.json configs 2-4).

Produces, with numpy only, every array the reference's pass 1 would hand to pass 2
and to the filter tasks, in the reference's own formats:

* a smooth-plus-noise reference picture (so that deblock / CDEF / LR decisions
  actually fire, unlike on white noise);
* per-block motion vectors -> `McItem`s, following recon.rs `mc()` addressing
  (src/recon.rs:2047-2055: mx = mvx & (15 >> !ss_hor), dx = bx*h_mul + (mvx >> (3 + ss_hor)));
* residual coefficients from the checkasm `ftx` recipe (rav1d_b200.synth.itxgen) packed
  back to back like `f.frame_thread.cf` (src/recon.rs:1706-1707), `ItxItem`s sorted by tx size;
* `Av1Filter` edge masks built with the min-of-both-sides rule of mask_edges_inter
  (src/lf_mask.rs:230-330), per-4x4 `level[4]`, `Av1FilterLUT` (rav1d_calc_eih, src/lf_mask.rs:608-626);
* `cdef_idx` / `noskip_mask` (src/decode.rs:1996-2004) and per-unit `Av1RestorationUnit`s with the
  tap / weight ranges of read_restoration_info (src/decode.rs:2559-2620).

Layout is 4:2:0; luma prediction blocks are 16x16 (8x8 chroma).
"""
import ctypes as C

import numpy as np

from .. import lib
from .itxgen import gen_coefs

BLK = 16  # luma prediction block edge


def calc_eih(sharp):
    """rav1d_calc_eih (src/lf_mask.rs:608-626) -> lib.FilterLUT"""
    lut = lib.FilterLUT()
    for level in range(64):
        limit = level
        if sharp > 0:
            limit >>= (sharp + 3) >> 2
            limit = min(limit, 9 - sharp)
        limit = max(limit, 1)
        lut.i[level] = limit
        lut.e[level] = 2 * (level + 2) + limit
    lut.sharp[0] = (sharp + 3) >> 2
    lut.sharp[1] = 9 - sharp if sharp else 0xff
    return lut


def smooth_plane(rng, h, w, bdmax, cell=16, noise=2):
    """Low-frequency content plus a little noise."""
    gh, gw = h // cell + 2, w // cell + 2
    coarse = rng.integers(bdmax // 8, bdmax - bdmax // 8, size=(gh, gw)).astype(np.float32)
    # bilinear upsample
    ys = (np.arange(h) + 0.5) / cell
    xs = (np.arange(w) + 0.5) / cell
    y0 = np.floor(ys).astype(np.int64); x0 = np.floor(xs).astype(np.int64)
    fy = (ys - y0)[:, None].astype(np.float32); fx = (xs - x0)[None, :].astype(np.float32)
    a = coarse[y0][:, x0]; b = coarse[y0][:, x0 + 1]; c = coarse[y0 + 1][:, x0]; d = coarse[y0 + 1][:, x0 + 1]
    img = (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy
    img += rng.integers(-noise, noise + 1, size=(h, w))
    return np.clip(np.rint(img), 0, bdmax)


class SynthFrame:
    pass


def make_header(w, h, bpc, rng):
    hd = lib.FrameHeader()
    hd.width, hd.height, hd.bpc, hd.layout, hd.sb128 = w, h, bpc, lib.LAYOUT_I420, 0
    hd.lf_level_y[0] = hd.lf_level_y[1] = 32
    hd.lf_level_u = hd.lf_level_v = 32
    hd.cdef_damping = int(rng.integers(3, 7))
    for i in range(8):
        hd.cdef_y_strength[i] = int(rng.integers(0, 64))
        hd.cdef_uv_strength[i] = int(rng.integers(0, 64))
    hd.cdef_y_strength[1] = 0          # one preset without luma filtering
    hd.cdef_uv_strength[2] = 0         # one without chroma
    hd.cdef_y_strength[3] = 3          # secondary only
    for p in range(3):
        hd.lr_type[p] = lib.RESTORATION_SWITCHABLE
    hd.lr_unit_size_log2[0] = 6
    hd.lr_unit_size_log2[1] = 6
    return hd


def geometry(hd):
    w, h = hd.width, hd.height
    g = lib.FrameGeometry()
    g.bw = ((w + 7) >> 3) << 1; g.bh = ((h + 7) >> 3) << 1
    g.w4 = (w + 3) >> 2; g.h4 = (h + 3) >> 2
    g.sb128w = (g.bw + 31) >> 5; g.sb128h = (g.bh + 31) >> 5
    g.sbh = (g.bh + (16 << hd.sb128) - 1) >> (4 + hd.sb128)
    g.b4_stride = (g.bw + 31) & ~31
    g.ss_hor = g.ss_ver = 1
    g.n_planes = 3
    return g


def scale_mv(val, scale):
    """scale_mv of the reference's mc() (src/recon.rs:2128-2140; C: src/recon_tmpl.c:1018-1021)."""
    tmp = val.astype(np.int64) * scale + (scale - 0x4000) * 8
    return (np.sign(tmp) * ((np.abs(tmp) + 128) >> 8) + 32).astype(np.int64)


def generate(w, h, bpc, seed=1, res_amp_shift=4, skip_frac=0.1, comp_frac=0.0, warp_frac=0.0, obmc_frac=0.0,
             scaled_frac=0.0, scaled_size=None, gmv_frac=0.0, scaled_obmc_frac=0.0, comp_scaled_frac=0.0):
    """Returns a SynthFrame with numpy arrays; see module docstring."""
    rng = np.random.default_rng(seed)
    bdmax = (1 << bpc) - 1
    hbd = bpc > 8
    pdt = np.uint16 if hbd else np.uint8
    cdt = np.int32 if hbd else np.int16
    s = SynthFrame()
    s.w, s.h, s.bpc, s.bdmax = w, h, bpc, bdmax
    s.hdr = make_header(w, h, bpc, rng)
    g = s.geom = geometry(s.hdr)
    aw, ah = (w + 127) & ~127, (h + 127) & ~127
    s.aw, s.ah = aw, ah

    # ---- reference picture (visible area w x h; the padding is left zero and never read: MC clamps)
    s.ref = [np.zeros((ah, aw), pdt), np.zeros((ah // 2, aw // 2), pdt), np.zeros((ah // 2, aw // 2), pdt)]
    s.ref[0][:h, :w] = smooth_plane(rng, h, w, bdmax)
    for p in (1, 2):
        s.ref[p][:(h + 1) // 2, :(w + 1) // 2] = smooth_plane(rng, (h + 1) // 2, (w + 1) // 2, bdmax, cell=8)

    # ---- blocks
    nbx, nby = (w + BLK - 1) // BLK, (h + BLK - 1) // BLK
    nb = nbx * nby
    by, bx = np.divmod(np.arange(nb), nbx)
    mvx = rng.integers(-512, 513, size=nb)      # 1/8 luma pel, +-64 px
    mvy = rng.integers(-512, 513, size=nb)
    zero = rng.random(nb) < 0.15                # some full-pel / zero-phase vectors (copy and 1-D paths)
    mvx[zero] &= ~7
    zero = rng.random(nb) < 0.15
    mvy[zero] &= ~7
    f2d = rng.integers(0, 10, size=nb)          # 9 8-tap combinations + bilinear
    skip = rng.random(nb) < skip_frac
    ytx_split = (rng.random(nb) < 0.3) & ~skip  # luma: TX_16X16 or 4 x TX_8X8
    ctx_split = (rng.random(nb) < 0.2) & ~skip  # chroma: TX_8X8 or 4 x TX_4X4
    s.n_blocks = nb

    # compound blocks (BASELINE config 5): a second reference, a second vector, avg / w_avg / seg
    is_comp = rng.random(nb) < comp_frac
    mvx2 = rng.integers(-512, 513, size=nb); mvy2 = rng.integers(-512, 513, size=nb)
    mvx2[rng.random(nb) < 0.15] &= ~7
    mvy2[rng.random(nb) < 0.15] &= ~7
    # warped blocks: a random near-identity affine matrix per block and shear parameters in the range the
    # reference's own test uses (tests/checkasm/mc.c:563-600)
    is_warp = (rng.random(nb) < warp_frac) & ~is_comp
    wi = np.nonzero(is_warp)[0]
    warp = np.zeros(wi.size, lib.WARP_ITEM_DT)
    warp["x"] = bx[wi] * BLK; warp["y"] = by[wi] * BLK; warp["w"] = BLK; warp["h"] = BLK; warp["ref"] = 0
    mat = np.zeros((wi.size, 6), np.int64)
    mat[:, 0] = rng.integers(-(40 << 16), 40 << 16, size=wi.size)
    mat[:, 1] = rng.integers(-(40 << 16), 40 << 16, size=wi.size)
    mat[:, 2] = (1 << 16) + rng.integers(-3000, 3000, size=wi.size)
    mat[:, 3] = rng.integers(-3000, 3000, size=wi.size)
    mat[:, 4] = rng.integers(-3000, 3000, size=wi.size)
    mat[:, 5] = (1 << 16) + rng.integers(-3000, 3000, size=wi.size)
    # keep the block near its own position: fold the linear part's effect at the block centre into the offset
    mat[:, 0] -= (mat[:, 2] - (1 << 16)) * (bx[wi] * BLK) + mat[:, 3] * (by[wi] * BLK)
    mat[:, 1] -= mat[:, 4] * (bx[wi] * BLK) + (mat[:, 5] - (1 << 16)) * (by[wi] * BLK)
    warp["matrix"] = mat.astype(np.int32)
    warp["abcd"] = (rng.integers(0, 0x2000, size=(wi.size, 4)) - 0xa00).astype(np.int16)
    s.warp_items = warp
    comp = np.zeros(int(is_comp.sum()), lib.COMP_ITEM_DT)
    ci = np.nonzero(is_comp)[0]
    comp["x"] = bx[ci] * BLK; comp["y"] = by[ci] * BLK; comp["w"] = BLK; comp["h"] = BLK
    comp["ref"][:, 0] = 0; comp["ref"][:, 1] = 1
    comp["mv"][:, 0, 0] = mvy[ci]; comp["mv"][:, 0, 1] = mvx[ci]
    comp["mv"][:, 1, 0] = mvy2[ci]; comp["mv"][:, 1, 1] = mvx2[ci]
    comp["filter2d"] = f2d[ci]
    comp["comp_type"] = rng.integers(0, 4, size=ci.size)     # avg, w_avg, segmentation mask, wedge
    comp["jnt_weight"] = rng.choice(np.array([3, 5, 7, 9, 11, 13]), size=ci.size)   # dav1d quant_dist_lookup_table values
    comp["mask_sign"] = rng.integers(0, 2, size=ci.size)
    comp["wedge_idx"] = rng.integers(0, 16, size=ci.size)
    if gmv_frac > 0:
        # GLOBALMV_GLOBALMV blocks: one or both predictions are the reference's global-motion warp (luma, and chroma too:
        # the 16x16 blocks have 8x8 chroma); per-reference parameters near the identity (frame_hdr.gmv)
        comp["warp_mask"] = np.where(rng.random(ci.size) < gmv_frac, rng.integers(1, 4, size=ci.size) * 5, 0)
        gm = np.zeros((8, 6), np.int64)
        gm[:2, 0] = rng.integers(-(6 << 16), 6 << 16, size=2); gm[:2, 1] = rng.integers(-(6 << 16), 6 << 16, size=2)
        gm[:2, 2] = (1 << 16) + rng.integers(-600, 600, size=2); gm[:2, 3] = rng.integers(-600, 600, size=2)
        gm[:2, 4] = rng.integers(-600, 600, size=2); gm[:2, 5] = (1 << 16) + rng.integers(-600, 600, size=2)
        s.gmv_matrix = gm.astype(np.int32)
        s.gmv_abcd = np.zeros((8, 4), np.int16)
        s.gmv_abcd[:2] = (rng.integers(0, 0x800, size=(2, 4)) - 0x400).astype(np.int16)
    if comp_scaled_frac > 0:
        # compound blocks with one or both predictions from the reference of another size (slot 2; needs scaled_frac > 0)
        crng = np.random.default_rng(seed + 80)
        pick = crng.random(ci.size) < comp_scaled_frac
        which = crng.integers(1, 4, size=ci.size)
        comp["ref"][:, 0] = np.where(pick & ((which & 1) != 0), 2, comp["ref"][:, 0])
        comp["ref"][:, 1] = np.where(pick & ((which & 2) != 0), 2, comp["ref"][:, 1])
    s.comp_items = comp
    if comp_frac > 0:
        s.ref2 = [np.zeros((ah, aw), pdt), np.zeros((ah // 2, aw // 2), pdt), np.zeros((ah // 2, aw // 2), pdt)]
        s.ref2[0][:h, :w] = smooth_plane(rng, h, w, bdmax)
        for p in (1, 2):
            s.ref2[p][:(h + 1) // 2, :(w + 1) // 2] = smooth_plane(rng, (h + 1) // 2, (w + 1) // 2, bdmax, cell=8)

    mc = np.zeros(nb * 3, lib.MC_ITEM_DT)
    y_it = mc[:nb]
    y_it["dst_x"] = bx * BLK; y_it["dst_y"] = by * BLK
    y_it["src_x"] = bx * BLK + (mvx >> 3); y_it["src_y"] = by * BLK + (mvy >> 3)
    y_it["w"] = BLK; y_it["h"] = BLK; y_it["plane"] = 0
    y_it["mx"] = (mvx & 7) << 1; y_it["my"] = (mvy & 7) << 1
    y_it["filter2d"] = f2d
    for p in (1, 2):
        c_it = mc[p * nb:(p + 1) * nb]
        c_it["dst_x"] = bx * (BLK // 2); c_it["dst_y"] = by * (BLK // 2)
        c_it["src_x"] = bx * (BLK // 2) + (mvx >> 4); c_it["src_y"] = by * (BLK // 2) + (mvy >> 4)
        c_it["w"] = BLK // 2; c_it["h"] = BLK // 2; c_it["plane"] = p
        c_it["mx"] = mvx & 15; c_it["my"] = mvy & 15
        c_it["filter2d"] = f2d
    # OBMC (src/recon.rs:2205-2309) on some translational blocks: one strip from the block above
    # (blend_h) and one from the block to the left (blend_v), per plane, predicted with that neighbour's
    # vector and filter.  16x16 blocks: above strip 16x8 (4:2:0 chroma 8x4), left strip 8x16 (4x8).
    is_obmc = (rng.random(nb) < obmc_frac) & ~is_comp & ~is_warp
    def strips(above):
        sel = np.nonzero(is_obmc & ((by > 0) if above else (bx > 0)))[0]
        nbr = sel - nbx if above else sel - 1
        out = np.zeros(sel.size * 3, lib.MC_ITEM_DT)
        for p in range(3):
            sh = 0 if p == 0 else 1
            o = out[p * sel.size:(p + 1) * sel.size]
            bw_, bh_ = (BLK, BLK // 2) if above else (BLK // 2, BLK)
            o["dst_x"] = bx[sel] * BLK >> sh; o["dst_y"] = by[sel] * BLK >> sh
            o["src_x"] = (bx[sel] * BLK >> sh) + (mvx[nbr] >> (3 + sh)); o["src_y"] = (by[sel] * BLK >> sh) + (mvy[nbr] >> (3 + sh))
            o["w"] = bw_ >> sh; o["h"] = bh_ >> sh; o["plane"] = p
            o["mx"] = ((mvx[nbr] & 7) << 1) if p == 0 else (mvx[nbr] & 15)
            o["my"] = ((mvy[nbr] & 7) << 1) if p == 0 else (mvy[nbr] & 15)
            o["filter2d"] = f2d[nbr]
            o["flags"] = lib.MC_OBMC_ABOVE if above else lib.MC_OBMC_LEFT
        return out
    ab, lf_ = strips(True), strips(False)
    s.obmc_items = np.concatenate([ab, lf_])
    s.n_obmc = (len(ab), len(lf_))
    # Blocks predicted from a reference of another size (slot 2), src/recon.rs:2116-2199
    is_scaled = np.zeros(nb, bool)
    if scaled_frac > 0:
        is_scaled = (np.random.default_rng(seed + 77).random(nb) < scaled_frac) & ~is_comp & ~is_warp & ~is_obmc
        rw, rh = scaled_size
        s.scaled_ref_size = (rw, rh)
        sc = [((rw << 14) + (w >> 1)) // w, ((rh << 14) + (h >> 1)) // h]         # scale_fac, src/decode.rs (svc)
        st = [(sc[0] + 8) >> 4, (sc[1] + 8) >> 4]
        si = np.nonzero(is_scaled)[0]
        out = np.zeros(si.size * 3, lib.SCALED_ITEM_DT)
        for p in range(3):
            sh = 0 if p == 0 else 1
            o = out[p * si.size:(p + 1) * si.size]
            o["dst_x"] = bx[si] * BLK >> sh; o["dst_y"] = by[si] * BLK >> sh
            o["w"] = BLK >> sh; o["h"] = BLK >> sh; o["plane"] = p; o["ref"] = 2
            # orig_pos = (b4 * mul << 4) + mv * (1 << !ss): block position in 1/16 pel of this plane
            o["pos_x"] = scale_mv(((bx[si] * BLK >> sh) << 4) + mvx[si] * (2 >> sh), sc[0])
            o["pos_y"] = scale_mv(((by[si] * BLK >> sh) << 4) + mvy[si] * (2 >> sh), sc[1])
            o["step_x"] = st[0]; o["step_y"] = st[1]
            o["filter2d"] = f2d[si]
        s.scaled_items = out
        s.n_scaled = (len(out), 0, 0)
        if scaled_obmc_frac > 0:
            # OBMC strips predicted from the scaled reference, on top of some of the scaled blocks: the blend area of an ABOVE
            # strip is the block's upper half, of a LEFT strip its left half (obmc(), src/recon.rs:2205-2309); the strip's
            # own vector (the neighbour's) differs from the block's
            srng = np.random.default_rng(seed + 79)
            strips = []
            for kind, flag in ((0, lib.MC_OBMC_ABOVE), (1, lib.MC_OBMC_LEFT)):
                pick = si[srng.random(si.size) < scaled_obmc_frac]
                dvx, dvy = srng.integers(-40, 41, size=pick.size), srng.integers(-40, 41, size=pick.size)
                for p in range(3):
                    sh = 0 if p == 0 else 1
                    o = np.zeros(pick.size, lib.SCALED_ITEM_DT)
                    o["dst_x"] = bx[pick] * BLK >> sh; o["dst_y"] = by[pick] * BLK >> sh
                    o["w"] = (BLK >> sh) >> (kind == 1); o["h"] = (BLK >> sh) >> (kind == 0); o["plane"] = p; o["ref"] = 2
                    o["pos_x"] = scale_mv(((bx[pick] * BLK >> sh) << 4) + (mvx[pick] + dvx) * (2 >> sh), sc[0])
                    o["pos_y"] = scale_mv(((by[pick] * BLK >> sh) << 4) + (mvy[pick] + dvy) * (2 >> sh), sc[1])
                    o["step_x"] = st[0]; o["step_y"] = st[1]
                    o["filter2d"] = f2d[pick]; o["flags"] = flag
                    strips.append(o)
            above, left = np.concatenate(strips[:3]), np.concatenate(strips[3:])
            s.scaled_items = np.concatenate([out, above, left])
            s.n_scaled = (len(out), len(above), len(left))
        raw, rah = (rw + 127) & ~127, (rh + 127) & ~127
        rrng = np.random.default_rng(seed + 78)
        s.ref3 = [np.zeros((rah, raw), pdt), np.zeros((rah // 2, raw // 2), pdt), np.zeros((rah // 2, raw // 2), pdt)]
        s.ref3[0][:rh, :rw] = smooth_plane(rrng, rh, rw, bdmax)
        for p in (1, 2):
            s.ref3[p][:(rh + 1) // 2, :(rw + 1) // 2] = smooth_plane(rrng, (rh + 1) // 2, (rw + 1) // 2, bdmax, cell=8)
    keep = np.tile(~(is_comp | is_warp | is_scaled), 3)
    s.mc_items = np.ascontiguousarray(mc[keep])

    # ---- transform blocks: (plane, x, y, tx, txtp)
    TX_4X4, TX_8X8, TX_16X16 = 0, 1, 2
    recs = []
    ns = ~skip
    def add(plane, xs, ys, tx):
        recs.append((np.full(xs.size, plane), xs, ys, np.full(xs.size, tx)))
    sel = ns & ~ytx_split
    add(0, bx[sel] * 16, by[sel] * 16, TX_16X16)
    sel = ns & ytx_split
    for oy in (0, 8):
        for ox in (0, 8):
            add(0, bx[sel] * 16 + ox, by[sel] * 16 + oy, TX_8X8)
    for p in (1, 2):
        sel = ns & ~ctx_split
        add(p, bx[sel] * 8, by[sel] * 8, TX_8X8)
        sel = ns & ctx_split
        for oy in (0, 4):
            for ox in (0, 4):
                add(p, bx[sel] * 8 + ox, by[sel] * 8 + oy, TX_4X4)
    plane = np.concatenate([r[0] for r in recs]); xs = np.concatenate([r[1] for r in recs])
    ys = np.concatenate([r[2] for r in recs]); tx = np.concatenate([r[3] for r in recs])
    n_itx = tx.size
    r = rng.random(n_itx)
    ntypes = np.where(tx == TX_16X16, 12, 16)
    txtp = np.where(r < 0.6, 0, rng.integers(0, 16, size=n_itx) % ntypes)
    # the appender buckets transform blocks by (size, type): a counting sort on the host that keeps
    # the 1-D transform kind uniform across the threads of a warp (csrc/itx.cu)
    order = np.argsort(tx * 32 + txtp, kind="stable")
    plane, xs, ys, tx, txtp = plane[order], xs[order], ys[order], tx[order], txtp[order]
    itx = np.zeros(n_itx, lib.ITX_ITEM_DT)
    itx["x"] = xs; itx["y"] = ys; itx["plane"] = plane; itx["tx"] = tx; itx["txtp"] = txtp
    counts = np.zeros(19, np.int32)
    for t in range(19):
        counts[t] = int((tx == t).sum())
    dims = {TX_4X4: 16, TX_8X8: 64, TX_16X16: 256}
    ncoef = np.array([dims[int(t)] for t in (TX_4X4, TX_8X8, TX_16X16)])
    per = np.where(tx == TX_4X4, 16, np.where(tx == TX_8X8, 64, 256))
    cf_off = np.concatenate([[0], np.cumsum(per)[:-1]]) if n_itx else np.zeros(0, np.int64)
    itx["cf_off"] = cf_off
    n_coefs = int(per.sum())
    coef = np.zeros(n_coefs, cdt)
    eobs = np.zeros(n_itx, np.int32)
    ncols = np.zeros(n_itx, np.int32)
    res_max = max(bdmax >> res_amp_shift, 2)
    for t in (TX_4X4, TX_8X8, TX_16X16):
        for tp in range(16):
            idx = np.nonzero((tx == t) & (txtp == tp))[0]
            if not idx.size:
                continue
            c, e = gen_coefs(rng, t, tp, res_max, idx.size, "full")
            # eob uniform in [0, full]: zero everything above a random scan position
            n = dims[t]
            cut = rng.integers(0, n, size=idx.size)
            c[np.arange(n)[None, :] > cut[:, None]] = 0
            nz = c != 0
            last = np.where(nz.any(axis=1), n - 1 - np.argmax(nz[:, ::-1], axis=1), 0)
            if tp == 0:
                e = last  # DCT_DCT: eob 0 takes the dc-only shortcut (src/itx.rs:90)
            else:
                e = np.maximum(last, 1)
            eobs[idx] = e
            # leading coefficient columns that hold a non-zero value (column-major blocks of `sh` rows)
            sh = {TX_4X4: 4, TX_8X8: 8, TX_16X16: 16}[t]
            ncols[idx] = np.maximum(((np.arange(n) // sh + 1)[None, :] * nz).max(axis=1), 1)
            dst = cf_off[idx][:, None] + np.arange(n)[None, :]
            coef[dst.ravel()] = c.ravel().astype(cdt)
    itx["eob"] = eobs
    itx["ncols"] = ncols
    s.itx_items, s.itx_counts, s.coef, s.n_coefs = itx, counts, coef, n_coefs

    # ---- per-4x4 transform-size maps -> edge masks
    H4, W4 = g.sb128h * 32, g.b4_stride
    ylog = np.full((H4, W4), 2, np.int8)        # log2 of luma tx edge in 4-px units (2 = 16)
    yb = np.zeros((H4, W4), bool)               # 4x4 columns / rows that start a tx block
    split_map = np.zeros((nby, nbx), bool); split_map.ravel()[:] = ytx_split
    big = np.kron(split_map, np.ones((4, 4), bool))
    ylog[:big.shape[0], :big.shape[1]][big] = 1
    y4, x4 = np.mgrid[0:H4, 0:W4]
    step = (1 << ylog.astype(np.int64))
    col_edge = (x4 % step) == 0
    row_edge = (y4 % step) == 0
    inside = (y4 < g.h4) & (x4 < g.w4)
    masks = np.zeros(g.sb128w * g.sb128h, lib.AV1_FILTER_DT)
    fy = masks["filter_y"]
    left = np.empty_like(ylog); left[:, 1:] = ylog[:, :-1]; left[:, 0] = ylog[:, 0]
    top = np.empty_like(ylog); top[1:, :] = ylog[:-1, :]; top[0, :] = ylog[0, :]
    def setbits(arr, d, sel, a_idx, bitpos, idx, sb):
        halfw = 16
        yy = np.nonzero(sel)
        sbv = sb[yy]; a = a_idx[yy]; b = bitpos[yy]; i = idx[yy]
        np.bitwise_or.at(arr, (sbv, d, a, i, b // halfw), (1 << (b % halfw)).astype(np.uint16))
    sb = (y4 >> 5) * g.sb128w + (x4 >> 5)
    setbits(fy, 0, col_edge & inside, x4 & 31, y4 & 31, np.minimum(ylog, left).astype(np.int64), sb)
    setbits(fy, 1, row_edge & inside, y4 & 31, x4 & 31, np.minimum(ylog, top).astype(np.int64), sb)
    # chroma (4:2:0): units are chroma 4x4; sb128 holds 16 x 16 of them
    CH4, CW4 = H4 // 2, W4 // 2
    clog = np.full((CH4, CW4), 1, np.int8)      # 8x8 chroma tx
    csplit = np.zeros((nby, nbx), bool); csplit.ravel()[:] = ctx_split
    cbig = np.kron(csplit, np.ones((2, 2), bool))
    clog[:cbig.shape[0], :cbig.shape[1]][cbig] = 0
    cy4, cx4 = np.mgrid[0:CH4, 0:CW4]
    cstep = (1 << clog.astype(np.int64))
    ccol = (cx4 % cstep) == 0; crow = (cy4 % cstep) == 0
    cinside = (cy4 < (g.h4 + 1) // 2) & (cx4 < (g.w4 + 1) // 2)
    cleft = np.empty_like(clog); cleft[:, 1:] = clog[:, :-1]; cleft[:, 0] = clog[:, 0]
    ctop = np.empty_like(clog); ctop[1:, :] = clog[:-1, :]; ctop[0, :] = clog[0, :]
    csb = (cy4 >> 4) * g.sb128w + (cx4 >> 4)
    fuv = masks["filter_uv"]
    def setbits_uv(d, sel, a_idx, bitpos, idx):
        yy = np.nonzero(sel)
        sbv = csb[yy]; a = a_idx[yy]; b = bitpos[yy]; i = idx[yy]
        np.bitwise_or.at(fuv, (sbv, d, a, i, b // 8), (1 << (b % 8)).astype(np.uint16))
    setbits_uv(0, ccol & cinside, cx4 & 15, cy4 & 15, np.minimum(clog, cleft).astype(np.int64))
    setbits_uv(1, crow & cinside, cy4 & 15, cx4 & 15, np.minimum(clog, ctop).astype(np.int64))

    # ---- levels: per prediction block, a few zeros to exercise the neighbour fallback
    lv = rng.integers(0, 64, size=(4, nby, nbx)).astype(np.uint8)
    lv[rng.random(lv.shape) < 0.08] = 0
    levels = np.zeros((H4, W4, 4), np.uint8)
    for k in (0, 1):
        up = np.kron(lv[k], np.ones((4, 4), np.uint8))
        levels[:up.shape[0], :up.shape[1], k] = up[:H4, :W4]
    for k in (2, 3):
        up = np.kron(lv[k], np.ones((2, 2), np.uint8))
        levels[:up.shape[0], :up.shape[1], k] = up[:H4, :W4]
    s.levels = levels
    s.lut = calc_eih(int(rng.integers(0, 3)))

    # ---- CDEF: cdef_idx per 64x64, noskip bits per non-skip block
    n64y, n64x = g.sb128h * 2, g.sb128w * 2
    cidx = rng.integers(0, 8, size=(n64y, n64x)).astype(np.int8)
    cidx[rng.random(cidx.shape) < 0.05] = -1
    masks["cdef_idx"] = cidx.reshape(g.sb128h, 2, g.sb128w, 2).transpose(0, 2, 1, 3).reshape(-1, 4)
    nsk = masks["noskip_mask"]
    bsel = np.nonzero(ns)[0]
    bx4 = bx[bsel] * 4; by4 = by[bsel] * 4
    sbb = (by4 >> 5) * g.sb128w + (bx4 >> 5)
    for dy in (0, 2):
        np.bitwise_or.at(nsk, (sbb, ((by4 + dy) & 31) >> 1, (bx4 & 16) >> 4), (0xF << (bx4 & 15)).astype(np.uint16))
    s.masks = masks
    # The same facts as per-block records, the arguments rav1d_create_lf_mask_inter would get for these blocks
    # (SURVEY 8 row f2): assign to s.lf_blocks to have the masks and levels built on the device instead.
    rec = np.zeros(nb, lib.LF_BLOCK_DT)
    rec["bx"], rec["by"], rec["bs"] = bx * 4, by * 4, 12                     # BS_16x16
    rec["flags"] = np.where(skip, lib.LFB_SKIP, 0) | lib.LFB_HAS_CHROMA
    rec["ytx"], rec["uvtx"] = TX_16X16, np.where(ctx_split, TX_4X4, TX_8X8)
    rec["tx_split"][:, 0] = ytx_split                                        # one split: four TX_8X8
    rec["lvl"] = lv.reshape(4, nb).T
    s.lf_block_records = rec

    # ---- loop restoration units
    lrm = np.zeros(g.sb128w * g.sb128h, lib.AV1_RESTORATION_DT)
    u = lrm["lr"]
    shape = u["type"].shape
    kind = rng.random(shape)
    sgr_idx = rng.integers(0, 16, size=shape)
    typ = np.where(kind < 0.1, 0, np.where(kind < 0.55, lib.RESTORATION_WIENER, lib.RESTORATION_SGRPROJ + sgr_idx))
    u["type"] = typ.astype(np.uint8)
    for name in ("filter_h", "filter_v"):
        t0 = rng.integers(-5, 11, size=shape); t1 = rng.integers(-23, 9, size=shape); t2 = rng.integers(-17, 47, size=shape)
        t0[rng.random(shape) < 0.3] = 0           # 5-tap variant
        t0[:, 1:, :] = 0                          # chroma has no outer tap (src/decode.rs:2578)
        u[name] = np.stack([t0, t1, t2], axis=-1).astype(np.int8)
    sp = np.array([[140, 3236], [112, 2158], [93, 1618], [80, 1438], [70, 1295], [58, 1177], [47, 1079], [37, 996],
                   [30, 925], [25, 863], [0, 2589], [0, 1618], [0, 1177], [0, 925], [56, 0], [22, 0]])
    w0 = np.where(sp[sgr_idx, 0] != 0, rng.integers(-96, 32, size=shape), 0)
    w1 = np.where(sp[sgr_idx, 1] != 0, rng.integers(-32, 96, size=shape), 95)
    u["sgr_weights"] = np.stack([w0, w1], axis=-1).astype(np.int8)
    s.lr_masks = lrm
    return s


def generate_recon_layout(w, h, bpc, layout, seed=1, comp_frac=0.3, warp_frac=0.1, obmc_frac=0.2):
    """A reconstruction-only batch (prediction + residual, no filter metadata) for any pixel layout: 4:0:0, 4:2:0,
    4:2:2 or 4:4:4.  16x16 luma blocks; chroma blocks, vectors, phases, OBMC strips and transform sizes follow the
    sub-sampling as recon.rs `mc()` / `obmc()` derive them (src/recon.rs:2047-2055,2100-2101,2205-2309)."""
    rng = np.random.default_rng(seed)
    bdmax = (1 << bpc) - 1
    pdt, cdt = (np.uint16, np.int32) if bpc > 8 else (np.uint8, np.int16)
    ssx, ssy = int(layout != lib.LAYOUT_I444), int(layout == lib.LAYOUT_I420)
    n_planes = 1 if layout == lib.LAYOUT_I400 else 3
    s = SynthFrame()
    s.w, s.h, s.bpc, s.bdmax, s.layout = w, h, bpc, bdmax, layout
    hd = lib.FrameHeader()
    hd.width, hd.height, hd.bpc, hd.layout = w, h, bpc, layout
    s.hdr = hd
    g = s.geom = geometry(hd)
    g.ss_hor, g.ss_ver, g.n_planes = (ssx, ssy, 3) if n_planes == 3 else (0, 0, 1)
    aw, ah = (w + 127) & ~127, (h + 127) & ~127
    s.aw, s.ah = aw, ah

    def picture():
        pl = [np.zeros((ah, aw), pdt)]
        pl[0][:h, :w] = smooth_plane(rng, h, w, bdmax)
        for _ in range(n_planes - 1):
            c = np.zeros((ah >> ssy, aw >> ssx), pdt)
            ch, cw = (h + ssy) >> ssy, (w + ssx) >> ssx
            c[:ch, :cw] = smooth_plane(rng, ch, cw, bdmax, cell=8)
            pl.append(c)
        return pl
    s.ref, s.ref2 = picture(), picture()
    nbx, nby = w // BLK, h // BLK                 # whole blocks only (warp / OBMC need them)
    nb = nbx * nby
    by, bx = np.divmod(np.arange(nb), nbx)
    mvx = rng.integers(-256, 257, size=nb); mvy = rng.integers(-256, 257, size=nb)
    mvx[rng.random(nb) < 0.15] &= ~7; mvy[rng.random(nb) < 0.15] &= ~7
    mvx2 = rng.integers(-256, 257, size=nb); mvy2 = rng.integers(-256, 257, size=nb)
    f2d = rng.integers(0, 10, size=nb)
    is_comp = rng.random(nb) < comp_frac
    is_warp = (rng.random(nb) < warp_frac) & ~is_comp
    is_obmc = (rng.random(nb) < obmc_frac) & ~is_comp & ~is_warp

    def plane_items(sel, p, bw, bh, vx, vy, fl, flags=0, dst_off=(0, 0)):
        """Rb200McItems of plane p for blocks `sel`: a bw x bh (luma units) area at the block origin + dst_off, moved by (vx, vy)."""
        sx, sy = (ssx, ssy) if p else (0, 0)
        o = np.zeros(sel.size, lib.MC_ITEM_DT)
        x0 = (bx[sel] * BLK + dst_off[0]) >> sx; y0 = (by[sel] * BLK + dst_off[1]) >> sy
        o["dst_x"] = x0; o["dst_y"] = y0
        o["src_x"] = x0 + (vx >> (3 + sx)); o["src_y"] = y0 + (vy >> (3 + sy))
        o["w"] = bw >> sx; o["h"] = bh >> sy; o["plane"] = p
        o["mx"] = (vx & (15 >> (1 - sx))) << (1 - sx); o["my"] = (vy & (15 >> (1 - sy))) << (1 - sy)
        o["filter2d"] = fl; o["flags"] = flags
        return o
    put = np.nonzero(~(is_comp | is_warp))[0]
    s.mc_items = np.concatenate([plane_items(put, p, BLK, BLK, mvx[put], mvy[put], f2d[put]) for p in range(n_planes)])
    ab = np.nonzero(is_obmc & (by > 0))[0]; lf_ = np.nonzero(is_obmc & (bx > 0))[0]
    above = [plane_items(ab, p, BLK, BLK // 2, mvx[ab - nbx], mvy[ab - nbx], f2d[ab - nbx], lib.MC_OBMC_ABOVE) for p in range(n_planes)]
    left = [plane_items(lf_, p, BLK // 2, BLK, mvx[lf_ - 1], mvy[lf_ - 1], f2d[lf_ - 1], lib.MC_OBMC_LEFT) for p in range(n_planes)]
    s.obmc_items = np.concatenate(above + left)
    s.n_obmc = (sum(len(a) for a in above), sum(len(a) for a in left))
    ci = np.nonzero(is_comp)[0]
    comp = np.zeros(ci.size, lib.COMP_ITEM_DT)
    comp["x"] = bx[ci] * BLK; comp["y"] = by[ci] * BLK; comp["w"] = BLK; comp["h"] = BLK
    comp["ref"][:, 1] = 1
    comp["mv"][:, 0, 0] = mvy[ci]; comp["mv"][:, 0, 1] = mvx[ci]; comp["mv"][:, 1, 0] = mvy2[ci]; comp["mv"][:, 1, 1] = mvx2[ci]
    comp["filter2d"] = f2d[ci]; comp["comp_type"] = rng.integers(0, 4, size=ci.size)
    comp["jnt_weight"] = rng.choice(np.array([3, 5, 7, 9, 11, 13]), size=ci.size)
    comp["mask_sign"] = rng.integers(0, 2, size=ci.size); comp["wedge_idx"] = rng.integers(0, 16, size=ci.size)
    s.comp_items = comp
    wi = np.nonzero(is_warp)[0]
    warp = np.zeros(wi.size, lib.WARP_ITEM_DT)
    warp["x"] = bx[wi] * BLK; warp["y"] = by[wi] * BLK; warp["w"] = BLK; warp["h"] = BLK
    mat = np.zeros((wi.size, 6), np.int64)
    mat[:, 0] = rng.integers(-(20 << 16), 20 << 16, size=wi.size); mat[:, 1] = rng.integers(-(20 << 16), 20 << 16, size=wi.size)
    mat[:, 2] = (1 << 16) + rng.integers(-3000, 3000, size=wi.size); mat[:, 3] = rng.integers(-3000, 3000, size=wi.size)
    mat[:, 4] = rng.integers(-3000, 3000, size=wi.size); mat[:, 5] = (1 << 16) + rng.integers(-3000, 3000, size=wi.size)
    mat[:, 0] -= (mat[:, 2] - (1 << 16)) * (bx[wi] * BLK) + mat[:, 3] * (by[wi] * BLK)
    mat[:, 1] -= mat[:, 4] * (bx[wi] * BLK) + (mat[:, 5] - (1 << 16)) * (by[wi] * BLK)
    warp["matrix"] = mat.astype(np.int32)
    warp["abcd"] = (rng.integers(0, 0x2000, size=(wi.size, 4)) - 0xa00).astype(np.int16)
    s.warp_items = warp

    # residuals: one luma transform per block, one chroma transform of the sub-sampled block size per plane
    from rav1d_b200.lib import TX_DIMS
    TX_8X8, TX_16X16 = 1, 2
    ctx = {(0, 0): TX_16X16, (1, 1): TX_8X8, (1, 0): next(t for t in range(19) if TX_DIMS[t] == (8, 16))}[(ssx, ssy)]
    recs = [(0, bx * BLK, by * BLK, np.full(nb, TX_16X16))]
    for p in range(1, n_planes):
        recs.append((p, bx * BLK >> ssx, by * BLK >> ssy, np.full(nb, ctx)))
    plane = np.concatenate([np.full(nb, r[0]) for r in recs]); xs = np.concatenate([r[1] for r in recs])
    ys = np.concatenate([r[2] for r in recs]); tx = np.concatenate([r[3] for r in recs])
    keep = rng.random(tx.size) > 0.2
    plane, xs, ys, tx = plane[keep], xs[keep], ys[keep], tx[keep]
    from rav1d_b200.synth.itxgen import valid_txtps
    txtp = np.zeros(tx.size, np.int64)
    for t in np.unique(tx):
        m = tx == t
        v = np.array([tp for tp in valid_txtps(int(t)) if tp < 16])
        txtp[m] = np.where(rng.random(int(m.sum())) < 0.5, 0, rng.choice(v, size=int(m.sum())))
    order = np.argsort(tx * 32 + txtp, kind="stable")
    plane, xs, ys, tx, txtp = plane[order], xs[order], ys[order], tx[order], txtp[order]
    itx = np.zeros(tx.size, lib.ITX_ITEM_DT)
    itx["x"] = xs; itx["y"] = ys; itx["plane"] = plane; itx["tx"] = tx; itx["txtp"] = txtp
    per = np.array([min(TX_DIMS[int(t)][0], 32) * min(TX_DIMS[int(t)][1], 32) for t in tx], np.int64)
    cf_off = np.concatenate([[0], np.cumsum(per)[:-1]]) if tx.size else np.zeros(0, np.int64)
    itx["cf_off"] = cf_off
    coef = np.zeros(int(per.sum()), cdt)
    for t in np.unique(tx):
        for tp in np.unique(txtp[tx == t]):
            idx = np.nonzero((tx == t) & (txtp == tp))[0]
            c, e = gen_coefs(rng, int(t), int(tp), max(bdmax >> 4, 2), idx.size, "full")
            n = c.shape[1]
            cut = rng.integers(0, n, size=idx.size)
            c[np.arange(n)[None, :] > cut[:, None]] = 0
            nz = c != 0
            last = np.where(nz.any(axis=1), n - 1 - np.argmax(nz[:, ::-1], axis=1), 0)
            itx["eob"][idx] = last if tp == 0 else np.maximum(last, 1)
            coef[(cf_off[idx][:, None] + np.arange(n)[None, :]).ravel()] = c.ravel().astype(cdt)
    s.itx_items, s.coef, s.n_coefs = itx, coef, int(per.sum())
    s.itx_counts = np.array([int((tx == t).sum()) for t in range(19)], np.int32)
    n_sb = g.sb128w * g.sb128h
    s.masks = np.zeros(n_sb, lib.AV1_FILTER_DT); s.lr_masks = np.zeros(n_sb, lib.AV1_RESTORATION_DT)
    s.levels = np.zeros((g.sb128h * 32, g.b4_stride, 4), np.uint8)
    s.lut = calc_eih(0)
    return s


def generate_intra(w, h, bpc, seed=1, inter_frac=0.0, layout=None, cfl_frac=0.25, ii_frac=0.3, pal_frac=0.1, ibc_frac=0.0):
    """A frame (4:2:0 unless `layout` says otherwise) whose 16x16 blocks are intra predicted (a fraction `inter_frac` of them translational inter blocks):
    per transform block a coded mode, angle delta, edge-availability flags consistent with the decode order (raster over
    blocks; inside a block luma transform blocks in raster order, then U, then V), the wavefront level the batch
    path needs, and a residual.  Reconstruction only (no filter metadata)."""
    from rav1d_b200.lib import TX_DIMS
    from rav1d_b200.synth.itxgen import valid_txtps
    rng = np.random.default_rng(seed)
    bdmax = (1 << bpc) - 1
    pdt, cdt = (np.uint16, np.int32) if bpc > 8 else (np.uint8, np.int16)
    s = SynthFrame()
    layout = lib.LAYOUT_I420 if layout is None else layout
    ssx, ssy = int(layout != lib.LAYOUT_I444), int(layout == lib.LAYOUT_I420)
    n_planes = 1 if layout == lib.LAYOUT_I400 else 3
    s.w, s.h, s.bpc, s.bdmax, s.layout = w, h, bpc, bdmax, layout
    hd = lib.FrameHeader()
    hd.width, hd.height, hd.bpc, hd.layout = w, h, bpc, layout
    s.hdr = hd
    g = s.geom = geometry(hd)
    g.ss_hor, g.ss_ver, g.n_planes = (ssx, ssy, 3) if n_planes == 3 else (0, 0, 1)
    aw, ah = (w + 127) & ~127, (h + 127) & ~127
    s.aw, s.ah = aw, ah
    s.ref = [np.zeros((ah, aw), pdt)] + [np.zeros((ah >> ssy, aw >> ssx), pdt) for _ in range(n_planes - 1)]
    s.ref[0][:h, :w] = smooth_plane(rng, h, w, bdmax)
    for p in range(1, n_planes):
        s.ref[p][:(h + ssy) >> ssy, :(w + ssx) >> ssx] = smooth_plane(rng, (h + ssy) >> ssy, (w + ssx) >> ssx, bdmax, cell=8)
    nbx, nby = w // BLK, h // BLK
    eief = int(rng.integers(0, 2))
    TXS = {4: 0, 8: 1, 16: 2}
    # per-plane maps in 4x4 cells: decode index of the transform block that owns the cell, its level
    psx = [0] + [ssx] * (n_planes - 1); psy = [0] + [ssy] * (n_planes - 1)
    pw4 = [g.bw >> psx[p] for p in range(n_planes)]; ph4 = [g.bh >> psy[p] for p in range(n_planes)]
    BIG = 1 << 60                                                         # not decoded yet: never available
    dec = [np.full((ph4[p], pw4[p]), BIG, np.int64) for p in range(n_planes)]
    lvl = [np.full((ph4[p], pw4[p]), -1, np.int64) for p in range(n_planes)]
    is_inter = rng.random((nby, nbx)) < inter_frac
    is_ii = is_inter & (rng.random((nby, nbx)) < ii_frac)                 # inter-intra: inter prediction blended with an intra one
    for byi, bxi in zip(*np.nonzero(is_inter & ~is_ii)):                  # plain inter blocks are reconstructed before any intra block
        for p in range(n_planes):
            cx, cy = BLK // 4 >> psx[p], BLK // 4 >> psy[p]
            dec[p][byi * cy:(byi + 1) * cy, bxi * cx:(bxi + 1) * cx] = -1
    items, itx_rows, mc_rows = [], [], []      # decode order
    pal_records, pal_bytes = [], 0             # palette blocks: { 8 entries padded to 16 bytes, w * h indices }
    inter_itx = []
    for byi in range(nby):
        for bxi in range(nbx):
            if is_inter[byi, bxi]:
                mvx, mvy, f2d = int(rng.integers(-128, 129)), int(rng.integers(-128, 129)), int(rng.integers(0, 10))
                for p in range(n_planes):
                    sx, sy = psx[p], psy[p]
                    x0, y0 = bxi * BLK >> sx, byi * BLK >> sy
                    mc_rows.append((x0, y0, x0 + (mvx >> (3 + sx)), y0 + (mvy >> (3 + sy)), BLK >> sx, BLK >> sy, p, 0,
                                    (mvx & (15 >> (1 - sx))) << (1 - sx), (mvy & (15 >> (1 - sy))) << (1 - sy), f2d, 0))
                    rect = {(0, 0): 2, (1, 1): 1, (1, 0): next(t for t in range(19) if TX_DIMS[t] == (8, 16))}[(sx, sy)]
                    if is_ii[byi, bxi]:
                        # the intra half: one item per plane over the whole block (src/recon.rs:3475-3550); its residual hangs on it
                        x4, y4, tw4, th4 = x0 // 4, y0 // 4, (BLK >> sx) // 4, (BLK >> sy) // 4
                        idx = len(items)
                        if p == 0:
                            ii_mode = int(rng.choice([0, 1, 2, 9])); ii_wedge = int(rng.integers(0, 16)) if rng.random() < 0.5 else -1
                        have_left, have_top = int(x4 > 0), int(y4 > 0)
                        W4, H4 = pw4[p], ph4[p]
                        level = 0
                        for (xa, xb, ya, yb) in ((x4 - 1, x4, y4 - 1, y4 + th4), (x4, x4 + tw4, y4 - 1, y4)):
                            lv = lvl[p][max(ya, 0):min(yb, H4), max(xa, 0):min(xb, W4)]
                            if lv.size: level = max(level, int(lv.max()) + 1)
                        dec[p][y4:y4 + th4, x4:x4 + tw4] = idx
                        lvl[p][y4:y4 + th4, x4:x4 + tw4] = level
                        items.append((x4, y4, W4, H4, p, tw4, th4, ii_mode, ii_wedge, have_left | have_top << 1 | 64, level))
                        itx_rows.append((p, x0, y0, rect) if rng.random() < 0.7 else None)
                    elif rng.random() < 0.7:
                        inter_itx.append((p, x0, y0, rect))
                continue
            if ibc_frac and (byi or bxi) and rng.random() < ibc_frac:
                # intra block copy (src/recon.rs:3196-3240): per plane one item of mode 16 over the whole block, predicted from
                # an area of the picture that is already decoded (rows of blocks above, or -- in the first row -- blocks to the
                # left), a whole number of luma pixels away; the residual's transform blocks follow as items of mode 15
                if byi: sxl, syl = int(rng.integers(-8, w - 4)), int(rng.integers(-8, (byi - 1) * BLK - 1))
                else: sxl, syl = int(rng.integers(-8, (bxi - 1) * BLK - 1)), int(rng.integers(-6, -1))
                mvx, mvy = 8 * (sxl - bxi * BLK), 8 * (syl - byi * BLK)
                for p in range(n_planes):
                    sx, sy = psx[p], psy[p]
                    bwp, bhp = BLK >> sx, BLK >> sy
                    x0, y0 = bxi * BLK >> sx, byi * BLK >> sy
                    x4, y4, tw4, th4 = x0 // 4, y0 // 4, bwp // 4, bhp // 4
                    dx, dy = x0 + (mvx >> (3 + sx)), y0 + (mvy >> (3 + sy))
                    frac = ((mvx & (15 >> (1 - sx))) << (1 - sx)) | ((mvy & (15 >> (1 - sy))) << (1 - sy)) << 4
                    W4, H4 = pw4[p], ph4[p]
                    xa, ya = min(max(dx >> 2, 0), W4 - 1), min(max(dy >> 2, 0), H4 - 1)
                    xb, yb = max((dx + bwp + 4) >> 2, xa + 1), max((dy + bhp + 4) >> 2, ya + 1)
                    assert (dec[p][ya:min(yb, H4), xa:min(xb, W4)] < len(items)).all()
                    level = int(lvl[p][ya:min(yb, H4), xa:min(xb, W4)].max()) + 1
                    dec[p][y4:y4 + th4, x4:x4 + tw4] = len(items)
                    lvl[p][y4:y4 + th4, x4:x4 + tw4] = level
                    items.append((x4, y4, dx & 0xffff, dy & 0xffff, p, tw4, th4, 16, frac - 256 if frac > 127 else frac, 0, level))
                    itx_rows.append(None)
                    tsz = int(rng.choice([t for t in (4, 8, 16) if t <= min(bwp, bhp)]))
                    for ty in range(0, th4, tsz // 4):
                        for tx_ in range(0, tw4, tsz // 4):
                            if rng.random() < 0.6:
                                items.append((x4 + tx_, y4 + ty, 0, 0, p, tsz // 4, tsz // 4, 15, 0, 0, level + 1))
                                itx_rows.append((p, (x4 + tx_) * 4, (y4 + ty) * 4, TXS[tsz]))
                                lvl[p][y4 + ty:y4 + ty + tsz // 4, x4 + tx_:x4 + tx_ + tsz // 4] = level + 1
                continue
            uv_cfl = rng.random() < cfl_frac                       # chroma from luma for this block (needs its luma first)
            pal_y, pal_uv = rng.random() < pal_frac, rng.random() < pal_frac
            for p in range(n_planes):
                bwp, bhp = BLK >> psx[p], BLK >> psy[p]            # block size in this plane, pixels
                if (pal_y and p == 0) or (pal_uv and p > 0):
                    # palette prediction of the whole block (no neighbours needed), then its residual transform blocks
                    x4, y4, tw4, th4 = bxi * bwp // 4, byi * bhp // 4, bwp // 4, bhp // 4
                    rec = np.zeros(16 + ((bwp * bhp + 15) & ~15), np.uint8)
                    rec[:8 * np.dtype(pdt).itemsize] = rng.integers(0, bdmax + 1, size=8).astype(pdt).view(np.uint8)
                    rec[16:16 + bwp * bhp] = rng.integers(0, 8, size=bwp * bhp)
                    off16 = pal_bytes // 16
                    pal_records.append(rec); pal_bytes += rec.size
                    tsz = int(rng.choice([t for t in (4, 8, 16) if t <= min(bwp, bhp)]))
                    whole = tsz == bwp == bhp
                    dec[p][y4:y4 + th4, x4:x4 + tw4] = len(items)
                    lvl[p][y4:y4 + th4, x4:x4 + tw4] = 0
                    items.append((x4, y4, off16 & 0xffff, off16 >> 16, p, tw4, th4, 14, 0, 0, 0))
                    itx_rows.append((p, x4 * 4, y4 * 4, TXS[tsz]) if whole and rng.random() < 0.8 else None)
                    if not whole:
                        for ty in range(0, th4, tsz // 4):
                            for tx_ in range(0, tw4, tsz // 4):
                                if rng.random() < 0.8:
                                    items.append((x4 + tx_, y4 + ty, 0, 0, p, tsz // 4, tsz // 4, 15, 0, 0, 1))
                                    itx_rows.append((p, (x4 + tx_) * 4, (y4 + ty) * 4, TXS[tsz]))
                                    lvl[p][y4 + ty:y4 + ty + tsz // 4, x4 + tx_:x4 + tx_ + tsz // 4] = 1
                    continue
                if p and uv_cfl:
                    # one item per plane over the whole chroma block; alpha 0 = plain DC prediction (src/recon.rs, CFL branch)
                    alpha = int(rng.integers(-16, 17))
                    x4, y4, tw4, th4 = bxi * bwp // 4, byi * bhp // 4, bwp // 4, bhp // 4
                    idx = len(items)
                    have_left, have_top = int(x4 > 0), int(y4 > 0)
                    W4, H4 = pw4[p], ph4[p]
                    level = int(lvl[0][byi * 4:(byi + 1) * 4, bxi * 4:(bxi + 1) * 4].max()) + 1 if alpha else 0
                    for (xa, xb, ya, yb) in ((x4 - 1, x4, y4 - 1, y4 + th4), (x4, x4 + tw4, y4 - 1, y4)):
                        lv = lvl[p][max(ya, 0):min(yb, H4), max(xa, 0):min(xb, W4)]
                        if lv.size: level = max(level, int(lv.max()) + 1)
                    dec[p][y4:y4 + th4, x4:x4 + tw4] = idx
                    lvl[p][y4:y4 + th4, x4:x4 + tw4] = level
                    w_pad = int(rng.integers(0, tw4)) if rng.random() < 0.3 else 0
                    h_pad = int(rng.integers(0, th4)) if rng.random() < 0.3 else 0
                    flags = have_left | have_top << 1 | eief << 5
                    items.append((x4, y4, W4 | (w_pad << 13 if alpha else 0), H4 | (h_pad << 13 if alpha else 0), p, tw4, th4,
                                  13 if alpha else 0, alpha, flags, level))
                    rect = next(t for t in range(19) if TX_DIMS[t] == (bwp, bhp))
                    itx_rows.append((p, x4 * 4, y4 * 4, rect) if rng.random() < 0.8 else None)
                    continue
                tsz = int(rng.choice([t for t in (4, 8, 16) if t <= min(bwp, bhp)]))
                mode = int(rng.integers(0, 14 if p == 0 else 13))  # filter-intra is luma only
                delta = int(rng.integers(-3, 4)) if 1 <= mode <= 8 else (int(rng.integers(0, 5)) if mode == 13 else 0)
                is_sm = int(rng.integers(0, 2))
                x0, y0 = bxi * bwp // 4, byi * bhp // 4            # 4-px units
                for ty in range(0, bhp // 4, tsz // 4):
                    for tx_ in range(0, bwp // 4, tsz // 4):
                        x4, y4, t4 = x0 + tx_, y0 + ty, tsz // 4
                        idx = len(items)
                        have_left, have_top = int(x4 > 0), int(y4 > 0)
                        W4, H4 = pw4[p], ph4[p]
                        deps = []

                        def cells(xa, xb, ya, yb):
                            return dec[p][max(ya, 0):min(yb, H4), max(xa, 0):min(xb, W4)], lvl[p][max(ya, 0):min(yb, H4), max(xa, 0):min(xb, W4)]
                        tr_ok = have_top and x4 + t4 < W4 and (cells(x4 + t4, x4 + 2 * t4, y4 - 1, y4)[0] < idx).all()
                        bl_ok = have_left and y4 + t4 < H4 and (cells(x4 - 1, x4, y4 + t4, y4 + 2 * t4)[0] < idx).all()
                        has_tr = int(tr_ok and rng.random() < 0.9)
                        has_bl = int(bl_ok and rng.random() < 0.9)
                        rects = [(x4 - 1, x4, y4 - 1, y4 + t4), (x4, x4 + t4, y4 - 1, y4)]
                        if has_tr: rects.append((x4 + t4, x4 + 2 * t4, y4 - 1, y4))
                        if has_bl: rects.append((x4 - 1, x4, y4 + t4, y4 + 2 * t4))
                        level = 0
                        for r_ in rects:
                            lv = cells(*r_)[1]
                            if lv.size: level = max(level, int(lv.max()) + 1)
                        dec[p][y4:y4 + t4, x4:x4 + t4] = idx
                        lvl[p][y4:y4 + t4, x4:x4 + t4] = level
                        flags = have_left | have_top << 1 | has_tr << 2 | has_bl << 3 | is_sm << 4 | eief << 5
                        items.append((x4, y4, W4, H4, p, t4, t4, mode, delta, flags, level))
                        itx_rows.append((p, x4 * 4, y4 * 4, TXS[tsz]) if rng.random() < 0.8 else None)
    out_items = np.zeros(len(items), lib.INTRA_ITEM_DT)
    for i, it in enumerate(items):
        out_items[i] = it[:10] + (it[10],)
    s.intra_items_decode = out_items
    s.palette = np.concatenate(pal_records) if pal_records else np.zeros(16, np.uint8)
    # ---- residuals: inter ones first (bucketed by size / type), then the intra ones level by level
    def make_itx(rows):
        plane = np.array([r[0] for r in rows], np.int64); xs = np.array([r[1] for r in rows], np.int64)
        ys = np.array([r[2] for r in rows], np.int64); tx = np.array([r[3] for r in rows], np.int64)
        txtp = np.zeros(len(rows), np.int64)
        for t in np.unique(tx):
            m = tx == t
            v = np.array([tp for tp in valid_txtps(int(t)) if tp < 16])
            txtp[m] = np.where(rng.random(int(m.sum())) < 0.5, 0, rng.choice(v, size=int(m.sum())))
        return plane, xs, ys, tx, txtp
    n_intra = len(items)
    levels = out_items["level"].astype(np.int64)
    n_levels = int(levels.max()) + 1 if n_intra else 0
    order_items = np.argsort(levels, kind="stable")
    s.intra_items = out_items[order_items]
    s._order_items = order_items
    s.intra_counts = np.bincount(levels, minlength=n_levels).astype(np.int32) if n_intra else np.zeros(0, np.int32)
    ip, ix, iy, itx_, itp = make_itx(inter_itx) if inter_itx else (np.zeros(0, np.int64),) * 5
    o = np.argsort(itx_ * 32 + itp, kind="stable")
    ip, ix, iy, itx_, itp = ip[o], ix[o], iy[o], itx_[o], itp[o]
    s.itx_counts = np.array([int((itx_ == t).sum()) for t in range(19)], np.int32)
    have = [i for i, r in enumerate(itx_rows) if r is not None]
    ap, ax, ay, atx, atp = make_itx([itx_rows[i] for i in have]) if have else (np.zeros(0, np.int64),) * 5
    alv = levels[have] if have else np.zeros(0, np.int64)
    o2 = np.argsort(alv * 1024 + atx * 32 + atp, kind="stable")
    s.intra_itx_counts = np.zeros((n_levels, 19), np.int32)
    for l_, t in zip(alv, atx):
        s.intra_itx_counts[int(l_), int(t)] += 1
    n_inter_itx = len(ip)
    plane = np.concatenate([ip, ap[o2]]); xs = np.concatenate([ix, ax[o2]]); ys = np.concatenate([iy, ay[o2]])
    tx = np.concatenate([itx_, atx[o2]]); txtp = np.concatenate([itp, atp[o2]])
    s.intra_itx_of = np.full(n_intra, -1, np.int32)
    for pos, k in enumerate(o2):
        s.intra_itx_of[have[int(k)]] = n_inter_itx + pos
    s.intra_itx_of_sorted = s.intra_itx_of[s._order_items]       # same, in the level-sorted order of intra_items
    itx = np.zeros(tx.size, lib.ITX_ITEM_DT)
    itx["x"] = xs; itx["y"] = ys; itx["plane"] = plane; itx["tx"] = tx; itx["txtp"] = txtp
    per = np.array([TX_DIMS[int(t)][0] * TX_DIMS[int(t)][1] for t in tx], np.int64)
    cf_off = np.concatenate([[0], np.cumsum(per)[:-1]]) if tx.size else np.zeros(0, np.int64)
    itx["cf_off"] = cf_off
    coef = np.zeros(int(per.sum()), cdt)
    for t in np.unique(tx):
        for tp in np.unique(txtp[tx == t]):
            idx = np.nonzero((tx == t) & (txtp == tp))[0]
            c, e = gen_coefs(rng, int(t), int(tp), max(bdmax >> 3, 2), idx.size, "full")
            n = c.shape[1]
            cut = rng.integers(0, n, size=idx.size)
            c[np.arange(n)[None, :] > cut[:, None]] = 0
            nz = c != 0
            last = np.where(nz.any(axis=1), n - 1 - np.argmax(nz[:, ::-1], axis=1), 0)
            itx["eob"][idx] = last if tp == 0 else np.maximum(last, 1)
            coef[(cf_off[idx][:, None] + np.arange(n)[None, :]).ravel()] = c.ravel().astype(cdt)
    s.itx_items, s.coef, s.n_coefs = itx, coef, int(per.sum())
    mc = np.zeros(len(mc_rows), lib.MC_ITEM_DT)
    for i, r_ in enumerate(mc_rows):
        mc[i] = r_
    s.mc_items = mc
    n_sb = g.sb128w * g.sb128h
    s.masks = np.zeros(n_sb, lib.AV1_FILTER_DT); s.lr_masks = np.zeros(n_sb, lib.AV1_RESTORATION_DT)
    s.levels = np.zeros((g.sb128h * 32, g.b4_stride, 4), np.uint8)
    s.lut = calc_eih(0)
    return s


def random_film_grain(rng, lag=None, luma_points=True, csfl=0, uv_points=(True, True), overlap=None):
    """Random Dav1dFilmGrainData within the ranges of tests/checkasm/filmgrain.c:156-190."""
    d = lib.FilmGrainData()
    d.seed = int(rng.integers(0, 0x10000))
    d.grain_scale_shift = int(rng.integers(0, 4))
    d.ar_coeff_shift = int(rng.integers(0, 4)) + 6
    d.ar_coeff_lag = int(rng.integers(0, 4)) if lag is None else lag
    n_y = 2 * d.ar_coeff_lag * (d.ar_coeff_lag + 1)
    for n in range(n_y):
        d.ar_coeffs_y[n] = int(rng.integers(0, 256)) - 128
    for uv in range(2):
        for n in range(n_y + 1):
            d.ar_coeffs_uv[uv][n] = int(rng.integers(0, 256)) - 128

    def points(arr, num):
        pad = 0xff // num
        for n in range(num):
            arr[n][0] = 0xff * n // num + int(rng.integers(0, pad))
            arr[n][1] = int(rng.integers(0, 256))
    if luma_points:
        d.num_y_points = 2 + int(rng.integers(0, 13))
        points(d.y_points, d.num_y_points)
    d.chroma_scaling_from_luma = csfl
    for uv in range(2):
        if uv_points[uv] and not csfl:
            d.num_uv_points[uv] = 2 + int(rng.integers(0, 9))
            points(d.uv_points[uv], d.num_uv_points[uv])
        d.uv_mult[uv] = int(rng.integers(0, 256)) - 128
        d.uv_luma_mult[uv] = int(rng.integers(0, 256)) - 128
        d.uv_offset[uv] = int(rng.integers(0, 512)) - 256
    d.clip_to_restricted_range = int(rng.integers(0, 2))
    d.scaling_shift = int(rng.integers(0, 4)) + 8
    d.overlap_flag = int(rng.integers(0, 2)) if overlap is None else overlap
    return d


def recon_input_planes(s, rng=None):
    """A plausible pre-filter picture for post-filter-only tests (BASELINE config 4)."""
    rng = rng or np.random.default_rng(7)
    pdt = s.ref[0].dtype
    out = [np.zeros_like(p) for p in s.ref]
    out[0][:s.h, :s.w] = smooth_plane(rng, s.h, s.w, s.bdmax, noise=3)
    for p in (1, 2):
        out[p][:(s.h + 1) // 2, :(s.w + 1) // 2] = smooth_plane(rng, (s.h + 1) // 2, (s.w + 1) // 2, s.bdmax, cell=8, noise=3)
    return [o.astype(pdt) for o in out]


def _plane_args(planes):
    """ctypes (data[3], stride[2]) of 1 (4:0:0) or 3 numpy planes."""
    ptrs = [p.ctypes.data for p in planes] + [None] * (3 - len(planes))
    data = (C.c_void_p * 3)(*ptrs)
    strides = (C.c_ssize_t * 2)(planes[0].strides[0], planes[1].strides[0] if len(planes) > 1 else 0)
    return data, strides


def sort_luma_first(s):
    """Reorders the put items and every transform-size bucket of the residual items luma first and records the counts
    rb200_frame_set_plane_counts wants (s.n_mc_luma, s.itx_luma_counts), so that the reconstruction can run as a luma and
    a chroma chain.  Frames with intra items are left alone (their items name residuals by index)."""
    if len(getattr(s, "intra_items", ())):
        return False
    mc = s.mc_items
    s.mc_items = mc[np.argsort(mc["plane"] != 0, kind="stable")]
    s.n_mc_luma = int((mc["plane"] == 0).sum())
    itx = s.itx_items.copy()
    luma = np.zeros(19, np.int32)
    off = 0
    for t in range(19):
        n = int(s.itx_counts[t])
        b = itx[off:off + n]
        itx[off:off + n] = b[np.argsort(b["plane"] != 0, kind="stable")]
        luma[t] = int((b["plane"] == 0).sum())
        off += n
    s.itx_items = itx
    s.itx_luma_counts = luma
    return True


class DeviceFrame:
    """Host-side driver of one rb200 frame object: fills the pinned staging from a SynthFrame."""

    def __init__(self, s):
        self.s = s
        self.h = C.c_void_p()
        lib.check(lib.frame_create(C.byref(self.h), C.byref(s.hdr), max(s.n_coefs, 1), max(len(s.itx_items), 1),
                                   max(len(s.mc_items), 1)), "frame_create")
        g = lib.FrameGeometry()
        lib.check(lib.frame_geometry(self.h, C.byref(g)))
        self.g = g
        self.ref_handle = None
        self.ref_handles = {}

    def close(self):
        if self.h:
            lib.frame_destroy(self.h)
            self.h = None
        if self.ref_handle:
            lib.frame_destroy(self.ref_handle)
            self.ref_handle = None
        for hnd in self.ref_handles.values():
            lib.frame_destroy(hnd)
        self.ref_handles = {}

    def load_batch(self):
        s, g = self.s, self.g
        lib.np_view(lib.frame_coef_buffer(self.h), s.coef.dtype, max(s.n_coefs, 1))[:s.n_coefs] = s.coef
        lib.np_view(lib.frame_itx_items(self.h), lib.ITX_ITEM_DT, max(len(s.itx_items), 1))[:len(s.itx_items)] = s.itx_items
        lib.np_view(lib.frame_mc_items(self.h), lib.MC_ITEM_DT, max(len(s.mc_items), 1))[:len(s.mc_items)] = s.mc_items
        n = g.sb128w * g.sb128h
        lib.np_view(lib.frame_lf_masks(self.h), lib.AV1_FILTER_DT, n)[:] = s.masks
        lv = lib.np_view(lib.frame_lf_levels(self.h), np.uint8, g.b4_stride * 32 * g.sb128h * 4)
        lv[:] = s.levels.reshape(-1)
        C.memmove(lib.frame_lf_lut(self.h), C.byref(s.lut), C.sizeof(lib.FilterLUT))
        lib.np_view(lib.frame_lr_masks(self.h), lib.AV1_RESTORATION_DT, len(s.lr_masks))[:] = s.lr_masks   # sr_sb128w * sb128h
        obmc = getattr(s, "obmc_items", None)
        if obmc is not None and len(obmc):
            lib.check(lib.frame_reserve_obmc_items(self.h, len(obmc)), "reserve_obmc_items")
            lib.np_view(lib.frame_obmc_items(self.h), lib.MC_ITEM_DT, len(obmc))[:] = obmc
            lib.check(lib.frame_set_obmc_counts(self.h, *s.n_obmc))
        intra = getattr(s, "intra_items", None)
        if intra is not None and len(intra):
            lib.check(lib.frame_reserve_intra_items(self.h, len(intra), len(s.intra_counts)), "reserve_intra_items")
            lib.np_view(lib.frame_intra_items(self.h), lib.INTRA_ITEM_DT, len(intra))[:] = intra
            lib.np_view(lib.frame_intra_itx_index(self.h), np.int32, len(intra))[:] = s.intra_itx_of_sorted
            pal = getattr(s, "palette", None)
            if pal is not None and len(pal):
                lib.check(lib.frame_reserve_palette(self.h, len(pal)), "reserve_palette")
                lib.np_view(lib.frame_palette_buffer(self.h), np.uint8, len(pal))[:] = pal
                lib.check(lib.frame_set_palette_bytes(self.h, len(pal)))
            lib.check(lib.frame_set_intra_levels(self.h, len(s.intra_counts), s.intra_counts.ctypes.data_as(C.POINTER(C.c_int32)),
                                                 np.ascontiguousarray(s.intra_itx_counts).ctypes.data_as(C.POINTER(C.c_int32))))
        scaled = getattr(s, "scaled_items", None)
        if scaled is not None and len(scaled):
            lib.check(lib.frame_reserve_scaled_items(self.h, len(scaled)), "reserve_scaled_items")
            lib.np_view(lib.frame_scaled_items(self.h), lib.SCALED_ITEM_DT, len(scaled))[:] = scaled
            n_put, n_above, n_left = getattr(s, "n_scaled", (len(scaled), 0, 0))
            lib.check(lib.frame_set_scaled_count(self.h, n_put))
            lib.check(lib.frame_set_scaled_obmc_counts(self.h, n_above, n_left))
        warp = getattr(s, "warp_items", None)
        if warp is not None and len(warp):
            lib.check(lib.frame_reserve_warp_items(self.h, len(warp)), "reserve_warp_items")
            lib.np_view(lib.frame_warp_items(self.h), lib.WARP_ITEM_DT, len(warp))[:] = warp
            lib.check(lib.frame_set_warp_count(self.h, len(warp)))
        lfb = getattr(s, "lf_blocks", None)
        if lfb is not None and len(lfb):       # masks and levels built on the device
            lib.check(lib.frame_reserve_lf_blocks(self.h, len(lfb)), "reserve_lf_blocks")
            lib.np_view(lib.frame_lf_blocks(self.h), lib.LF_BLOCK_DT, len(lfb))[:] = lfb
            lib.check(lib.frame_set_lf_block_count(self.h, len(lfb)))
        if hasattr(s, "gmv_matrix"):               # global-motion parameters of the references (compound warp predictions)
            for slot in range(8):
                lib.check(lib.frame_set_ref_gmv(self.h, slot, s.gmv_matrix[slot].ctypes.data_as(C.POINTER(C.c_int32)),
                                                s.gmv_abcd[slot].ctypes.data_as(C.POINTER(C.c_int16))))
        if hasattr(s, "itx_luma_counts"):          # lists sorted luma first (sort_luma_first)
            lib.check(lib.frame_set_plane_counts(self.h, s.n_mc_luma, s.itx_luma_counts.ctypes.data_as(C.POINTER(C.c_int32))))
        comp = getattr(s, "comp_items", None)
        if comp is not None and len(comp):
            lib.check(lib.frame_reserve_comp_items(self.h, len(comp)), "reserve_comp_items")
            lib.np_view(lib.frame_comp_items(self.h), lib.COMP_ITEM_DT, len(comp))[:] = comp
            lib.check(lib.frame_set_comp_count(self.h, len(comp)))

    def upload(self, which, planes):
        data, strides = _plane_args(planes)
        lib.check(lib.frame_upload_planes(self.h, which, data, strides), "frame_upload_planes")

    def set_ref_from_host(self, planes):
        """Reference pictures live in a second frame object's plane set 0 (device resident)."""
        if self.ref_handle is None:
            self.ref_handle = C.c_void_p()
            lib.check(lib.frame_create(C.byref(self.ref_handle), C.byref(self.s.hdr), 1, 1, 1), "frame_create(ref)")
        data, strides = _plane_args(planes)
        lib.check(lib.frame_upload_planes(self.ref_handle, 0, data, strides))
        pl = lib.Planes()
        lib.check(lib.frame_stage_planes(self.ref_handle, 0, C.byref(pl)))
        lib.check(lib.frame_set_ref(self.h, 0, C.byref(pl)))

    def set_ref_slot(self, slot, planes, size=None):
        """Upload a reference picture into its own device-resident frame object and bind it to `slot`.
        size = (w, h): a reference of another size than the current picture (scaled prediction)."""
        if slot not in self.ref_handles:
            hnd = C.c_void_p()
            hdr = self.s.hdr
            if size is not None:
                hdr = lib.FrameHeader.from_buffer_copy(bytes(self.s.hdr))
                hdr.width, hdr.height = size
            lib.check(lib.frame_create(C.byref(hnd), C.byref(hdr), 1, 1, 1), "frame_create(ref)")
            self.ref_handles[slot] = hnd
        hnd = self.ref_handles[slot]
        data, strides = _plane_args(planes)
        lib.check(lib.frame_upload_planes(hnd, 0, data, strides))
        pl = lib.Planes()
        lib.check(lib.frame_stage_planes(hnd, 0, C.byref(pl)))
        lib.check(lib.frame_set_ref(self.h, slot, C.byref(pl)))
        if size is not None:
            lib.check(lib.frame_set_ref_size(self.h, slot, size[0], size[1]))

    def submit(self, stages, upload=True):
        """upload: False / 0 = batch already on the device, True / 1 = copy the batch, 2 = copy everything but the
        coefficients, which the itx kernels read straight from the pinned staging (zero-copy)."""
        s = self.s
        counts = (C.c_int32 * 19)(*[int(c) for c in s.itx_counts])
        lib.check(lib.frame_submit(self.h, s.n_coefs, counts, len(s.mc_items), stages, int(upload)), "frame_submit")

    def wait(self):
        lib.check(lib.frame_wait(self.h), "frame_wait")

    def download_lf(self):
        """(Av1Filter[sb128h * sb128w], level[32 sb128h][b4_stride][4]) as the device holds them after the last submit."""
        g = self.g
        masks = np.zeros(g.sb128w * g.sb128h, lib.AV1_FILTER_DT)
        levels = np.zeros((32 * g.sb128h, g.b4_stride, 4), np.uint8)
        lib.check(lib.frame_download_lf(self.h, masks.ctypes.data, levels.ctypes.data), "frame_download_lf")
        return masks, levels

    def readback(self):
        s = self.s
        out = [np.zeros_like(p) for p in getattr(s, "readback_like", s.ref)]
        data, strides = _plane_args(out)
        lib.check(lib.frame_readback(self.h, data, strides), "frame_readback")
        return out


# ---- loop-filter block records: a random partition tree per superblock, in decode order
_BS_DIMS = [(32, 32), (32, 16), (16, 32), (16, 16), (16, 8), (16, 4), (8, 16), (8, 8), (8, 4), (8, 2), (4, 16), (4, 8), (4, 4),
            (4, 2), (4, 1), (2, 8), (2, 4), (2, 2), (2, 1), (1, 4), (1, 2), (1, 1)]     # BlockSize -> (w4, h4)
_TX_OF_LOG = {(0, 0): 0, (1, 1): 1, (2, 2): 2, (3, 3): 3, (4, 4): 4, (0, 1): 5, (1, 0): 6, (1, 2): 7, (2, 1): 8, (2, 3): 9,
              (3, 2): 10, (3, 4): 11, (4, 3): 12, (0, 2): 13, (2, 0): 14, (1, 3): 15, (3, 1): 16, (2, 4): 17, (4, 2): 18}


def _tx_for(lw, lh, cap):
    """Largest transform (RectTxfmSize) that fits a 2^lw x 2^lh (4-px units) area, sides capped, aspect <= 4:1."""
    lw, lh = min(lw, cap), min(lh, cap)
    lw, lh = min(lw, lh + 2), min(lh, lw + 2)
    return _TX_OF_LOG[(lw, lh)]


def generate_lf_blocks(w, h, layout=1, sb128=1, seed=1, intra_frac=0.3, skip_frac=0.3, min_log=0):
    """Block records (lib.LF_BLOCK_DT) covering a w x h picture in decode order (one tile): every superblock is split
    by a random partition tree (none / horizontal / vertical / 4-way horizontal / 4-way vertical / quad), blocks whose
    origin lies outside the 8-pixel aligned picture are not coded (src/decode.rs decode_sb), sizes are real BlockSizes.
    Inter blocks carry random tx_split words, max_ytx / uvtx follow the block size; levels are random with some zeros."""
    rng = np.random.default_rng(seed)
    ss_hor, ss_ver = int(layout not in (0, 3)), int(layout == 1)
    bw, bh = ((w + 7) >> 3) << 1, ((h + 7) >> 3) << 1
    sb = 32 if sb128 else 16
    bs_of = {d: i for i, d in enumerate(_BS_DIMS)}
    out = []

    def emit(x, y, bw4, bh4):
        if x >= bw or y >= bh:
            return
        lw, lh = bw4.bit_length() - 1, bh4.bit_length() - 1
        intra = rng.random() < intra_frac
        has_chroma = layout != 0 and (bw4 > ss_hor or (x & 1)) and (bh4 > ss_ver or (y & 1))
        flags = (lib.LFB_INTRA if intra else 0) | (lib.LFB_SKIP if rng.random() < skip_frac else 0) | \
                (lib.LFB_HAS_CHROMA if has_chroma else 0)
        ytx = _tx_for(lw, lh, 4)
        if intra and rng.random() < 0.5:        # intra blocks may use a smaller uniform transform
            k = int(rng.integers(0, 3))
            ytx = _tx_for(max(lw - k, 0), max(lh - k, 0), 4)
        uvtx = _tx_for(max(lw - ss_hor, 0), max(lh - ss_ver, 0), 3)
        split = rng.integers(0, 1 << 16, size=2) & rng.integers(0, 1 << 16, size=2) if rng.random() < 0.7 else (0, 0)
        lvl = rng.integers(0, 64, size=4)
        lvl[rng.random(4) < 0.1] = 0
        out.append((x, y, bs_of[(bw4, bh4)], flags, ytx, uvtx, (int(split[0]), int(split[1])), tuple(int(v) for v in lvl)))

    def part(x, y, n):      # n: side of the square in 4-px units
        if x >= bw or y >= bh:
            return
        r = rng.random()
        if n == 1 or (r < (0.08 if n >= 16 else 0.25) and n > (1 << min_log)):
            return emit(x, y, n, n)
        hn = n >> 1
        if n == 2 and ss_hor + ss_ver and r < 0.5:      # 8x8 stays whole half of the time (chroma 4x4)
            return emit(x, y, n, n)
        if r < 0.35:
            emit(x, y, n, hn); emit(x, y + hn, n, hn)
        elif r < 0.45:
            emit(x, y, hn, n); emit(x + hn, y, hn, n)
        elif r < 0.50 and n >= 4 and n <= 16:
            q = n >> 2
            for i in range(4):
                emit(x, y + i * q, n, q)
        elif r < 0.55 and n >= 4 and n <= 16:
            q = n >> 2
            for i in range(4):
                emit(x + i * q, y, q, n)
        else:
            for dy in (0, hn):
                for dx in (0, hn):
                    part(x + dx, y + dy, hn)

    for sy in range(0, bh, sb):
        for sx in range(0, bw, sb):
            part(sx, sy, sb)
    return np.array(out, dtype=lib.LF_BLOCK_DT)
