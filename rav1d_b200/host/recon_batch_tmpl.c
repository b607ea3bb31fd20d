/*
 * rav1d_b200 host layer: the per-bit-depth members of Rav1dFrameContext_bd_fn
 * (src/internal.rs:350-395; C twin src/internal.h:232-247) for the batch path.
 *
 *   recon_b_intra / recon_b_inter   src/recon.rs:2402-3160 / 3162-4045  (C: src/recon_tmpl.c:1200-1603 / 1605-2051)
 *   mc() / obmc() / warp_affine()   src/recon.rs:2025-2203 / 2205-2309 / 2311-2400 (C: :962-1074 / 1076-1137 / 1139-1198)
 *   read_coef_tree (pass 2 branch)  src/recon.rs:1544-1790 (C: :726-824)
 *   filter_sbrow_*                  src/recon.rs:4047-4338 (C: :2053-2175)
 *   backup_ipred_edge               src/recon.rs:4340-4400 (C: :2177-2201)
 *
 * Same signatures, same block walk, same context reads -- but where the reference computes pixels (emu_edge + mc /
 * mct / avg / w_mask / blend / warp8x8 / prepare_intra_edges + intra_pred / pal_pred / cfl / itxfm_add) these
 * functions append one fixed-size record to the frame's batch (include/rav1d_b200.h) and return.  They run in pass 2
 * of the reference's two-pass frame threading (src/decode.rs:1203-1326): pass 1 has filled f.frame_thread.{b, cbi, cf,
 * pal, pal_idx}; the tile state's cf / pal_idx cursors are advanced exactly as the reference advances them.
 * The filter entry points do nothing per superblock row: the whole frame's filters are launched by the submit.
 *
 * Compiled twice (BITDEPTH = 8 / 16) against the reference's headers, like the reference's own *_tmpl.c files.
 */
#include "config.h"

#include <limits.h>
#include <string.h>

#include "common/attributes.h"
#include "common/bitdepth.h"
#include "common/frame.h"
#include "common/intops.h"

#include "src/internal.h"
#include "src/ipred_prepare.h"
#include "src/recon.h"
#include "src/tables.h"
#include "src/levels.h"

#include "host_frame.h"

/* flags of Rb200IntraItem */
enum { II_HAVE_LEFT = 1, II_HAVE_TOP = 2, II_TOP_HAS_RIGHT = 4, II_LEFT_HAS_BOTTOM = 8, II_SMOOTH = 16, II_EDGE_FILTER = 32,
       II_INTER_INTRA = 64 };
enum { MODE_CFL = 13, MODE_PAL = 14, MODE_RESIDUAL = 15, MODE_INTRABC = 16 };

static inline int plane_ss_hor(const Dav1dFrameContext *const f, const int pl) {
    return pl && f->cur.p.layout != DAV1D_PIXEL_LAYOUT_I444;
}
static inline int plane_ss_ver(const Dav1dFrameContext *const f, const int pl) {
    return pl && f->cur.p.layout == DAV1D_PIXEL_LAYOUT_I420;
}

/* ---- residuals: one Rb200ItxItem per transform block with eob >= 0 (the itxfm_add call sites) ---- */
static void fill_itx(Rb200ItxItem *const it, const Dav1dFrameContext *const f, const coef *const cf, const int pl,
                     const int x, const int y, const int tx, const int txtp, const int eob)
{
    it->cf_off = (uint32_t)(cf - (const coef *)f->frame_thread.cf);
    it->x = (uint16_t)x; it->y = (uint16_t)y;
    it->plane = (uint8_t)pl; it->tx = (uint8_t)tx; it->txtp = (uint8_t)txtp;
    it->ncols = 0;
    it->eob = (int16_t)eob; it->pad = 0;
}

/* A residual that must follow a prediction of the intra wavefront (or stand alone in it). */
static void push_intra_residual_only(RbHostBatch *const B, const Dav1dFrameContext *const f, const Dav1dTileState *const ts,
                                     const coef *const cf, const int pl, const int x4, const int y4, const int tx,
                                     const int txtp, const int eob)
{
    const TxfmInfo *const td = &dav1d_txfm_dimensions[tx];
    Rb200IntraItem *const it = RB_PUSH(B->intra);
    memset(it, 0, sizeof(*it));
    it->x4 = (uint16_t)x4; it->y4 = (uint16_t)y4;
    it->w4_end = (uint16_t)(ts->tiling.col_end >> plane_ss_hor(f, pl));
    it->h4_end = (uint16_t)(ts->tiling.row_end >> plane_ss_ver(f, pl));
    it->plane = (uint8_t)pl; it->tw4 = td->w; it->th4 = td->h; it->mode = MODE_RESIDUAL;
    *RB_PUSH(B->intra_itx) = B->iitx.n;
    fill_itx(RB_PUSH(B->iitx), f, cf, pl, 4 * x4, 4 * y4, tx, txtp, eob);
}

/* Intra block copy of one plane: mc() (src/recon_tmpl.c:962-1014) with refp == the current picture, as an item of the
 * intra wavefront.  x4 / y4: destination in plane 4-pixel units; bw4 / bh4, bx / by: mc()'s arguments. */
static void push_intrabc(RbHostBatch *const B, const Dav1dFrameContext *const f, const int pl, const int x4, const int y4,
                         const int bw4, const int bh4, const int bx, const int by, const mv mv)
{
    const int ss_ver = plane_ss_ver(f, pl), ss_hor = plane_ss_hor(f, pl);
    const int h_mul = 4 >> ss_hor, v_mul = 4 >> ss_ver;
    const int mx = mv.x & (15 >> !ss_hor), my = mv.y & (15 >> !ss_ver);
    const int w = f->bw * 4 >> ss_hor, h = f->bh * 4 >> ss_ver;
    /* emu_edge replicates the border: a window that lies wholly outside can be moved next to the picture */
    const int dx = iclip(bx * h_mul + (mv.x >> (3 + ss_hor)), -(bw4 * h_mul + 16), w + 16);
    const int dy = iclip(by * v_mul + (mv.y >> (3 + ss_ver)), -(bh4 * v_mul + 16), h + 16);
    if (w + 16 > 32767 || h + 16 > 32767) { rb_batch_unsupported(B, "intra block copy in a picture beyond 32K"); return; }
    Rb200IntraItem *const it = RB_PUSH(B->intra);
    memset(it, 0, sizeof(*it));
    it->x4 = (uint16_t)x4; it->y4 = (uint16_t)y4;
    it->w4_end = (uint16_t)(int16_t)dx; it->h4_end = (uint16_t)(int16_t)dy;
    it->plane = (uint8_t)pl; it->tw4 = (uint8_t)(bw4 * h_mul >> 2); it->th4 = (uint8_t)(bh4 * v_mul >> 2);
    it->mode = MODE_INTRABC;
    it->angle = (int8_t)(uint8_t)((mx << !ss_hor) | (my << !ss_ver) << 4);
    *RB_PUSH(B->intra_itx) = -1;
}

/* ---- mc(): src/recon.rs:2025-2203 (C: src/recon_tmpl.c:962-1074).
 * list: 0 = the block's own prediction, 1 / 2 = OBMC strip from the above / left neighbour (then w x h below is the
 * blend area and pred_bh4 the rows the reference predicts).  dst_x / dst_y: destination in plane pixels. */
static void push_mc(RbHostBatch *const B, const Dav1dFrameContext *const f, const int list, const int dst_x, const int dst_y,
                    const int bw4, const int bh4, const int bx, const int by, const int pl, const mv mv, const int refidx,
                    const enum Filter2d filter_2d, const int blend_w, const int blend_h)
{
    const int ss_ver = plane_ss_ver(f, pl), ss_hor = plane_ss_hor(f, pl);
    const int h_mul = 4 >> ss_hor, v_mul = 4 >> ss_ver;
    const int mvx = mv.x, mvy = mv.y;
    const int mx = mvx & (15 >> !ss_hor), my = mvy & (15 >> !ss_ver);
    const Dav1dThreadPicture *const refp = &f->refp[refidx];

    if (refp->p.p.w == f->cur.p.w && refp->p.p.h == f->cur.p.h) {
        const int w = (f->cur.p.w + ss_hor) >> ss_hor, h = (f->cur.p.h + ss_ver) >> ss_ver;
        int dx = bx * h_mul + (mvx >> (3 + ss_hor));
        int dy = by * v_mul + (mvy >> (3 + ss_ver));
        /* emu_edge replicates the border: a window that lies wholly outside can be moved next to the picture */
        dx = iclip(dx, -(bw4 * h_mul + 16), w + 16);
        dy = iclip(dy, -(bh4 * v_mul + 16), h + 16);
        Rb200McItem *const it = list == 0 ? RB_PUSH(B->mc) : list == 1 ? RB_PUSH(B->obmc_above) : RB_PUSH(B->obmc_left);
        it->dst_x = (int16_t)dst_x; it->dst_y = (int16_t)dst_y;
        it->src_x = (int16_t)dx; it->src_y = (int16_t)dy;
        it->w = (uint8_t)(list ? blend_w : bw4 * h_mul);
        it->h = (uint8_t)(list ? blend_h : bh4 * v_mul);
        it->plane = (uint8_t)pl; it->ref = (uint8_t)refidx;
        it->mx = (uint8_t)(mx << !ss_hor); it->my = (uint8_t)(my << !ss_ver);
        it->filter2d = (uint8_t)filter_2d;
        it->flags = (uint8_t)(list == 0 ? RB200_MC_PUT : list == 1 ? RB200_MC_OBMC_ABOVE : RB200_MC_OBMC_LEFT);
    } else {
        /* reference of another size: the position arithmetic of the scaled branch stays on the host */
        const int orig_pos_y = (by * v_mul << 4) + mvy * (1 << !ss_ver);
        const int orig_pos_x = (bx * h_mul << 4) + mvx * (1 << !ss_hor);
#define scale_mv(res, val, scale) do { \
            const int64_t tmp = (int64_t)(val) * scale + (scale - 0x4000) * 8; \
            res = apply_sign64((int) ((llabs(tmp) + 128) >> 8), tmp) + 32;     \
        } while (0)
        int pos_y, pos_x;
        scale_mv(pos_x, orig_pos_x, f->svc[refidx][0].scale);
        scale_mv(pos_y, orig_pos_y, f->svc[refidx][1].scale);
#undef scale_mv
        Rb200McScaledItem *const it = list == 0 ? RB_PUSH(B->scaled) : list == 1 ? RB_PUSH(B->scaled_obmc_above) : RB_PUSH(B->scaled_obmc_left);
        memset(it, 0, sizeof(*it));
        it->dst_x = (int16_t)dst_x; it->dst_y = (int16_t)dst_y;
        it->w = (uint8_t)(list ? blend_w : bw4 * h_mul); it->h = (uint8_t)(list ? blend_h : bh4 * v_mul);
        it->flags = (uint8_t)(list == 0 ? RB200_MC_PUT : list == 1 ? RB200_MC_OBMC_ABOVE : RB200_MC_OBMC_LEFT);
        it->plane = (uint8_t)pl; it->ref = (uint8_t)refidx;
        it->pos_x = pos_x; it->pos_y = pos_y;
        it->step_x = f->svc[refidx][0].step; it->step_y = f->svc[refidx][1].step;
        it->filter2d = (uint8_t)filter_2d;
    }
}

/* ---- obmc(): src/recon.rs:2205-2309 (C: src/recon_tmpl.c:1076-1137) ---- */
static void push_obmc(RbHostBatch *const B, Dav1dTaskContext *const t, const int dst_x, const int dst_y,
                      const uint8_t *const b_dim, const int pl, const int bx4, const int by4, const int w4, const int h4)
{
    const Dav1dFrameContext *const f = t->f;
    /*const*/ refmvs_block **r = &t->rt.r[(t->by & 31) + 5];
    const int ss_ver = plane_ss_ver(f, pl), ss_hor = plane_ss_hor(f, pl);
    const int h_mul = 4 >> ss_hor, v_mul = 4 >> ss_ver;

    if (t->by > t->ts->tiling.row_start &&
        (!pl || b_dim[0] * h_mul + b_dim[1] * v_mul >= 16))
    {
        for (int i = 0, x = 0; x < w4 && i < imin(b_dim[2], 4); ) {
            // only odd blocks are considered for overlap handling, hence +1
            const refmvs_block *const a_r = &r[-1][t->bx + x + 1];
            const uint8_t *const a_b_dim = dav1d_block_dimensions[a_r->bs];
            const int step4 = iclip(a_b_dim[0], 2, 16);

            if (a_r->ref.ref[0] > 0) {
                const int ow4 = imin(step4, b_dim[0]);
                const int oh4 = imin(b_dim[1], 16) >> 1;
                push_mc(B, f, 1, dst_x + x * h_mul, dst_y, ow4, (oh4 * 3 + 3) >> 2, t->bx + x, t->by, pl,
                        a_r->mv.mv[0], a_r->ref.ref[0] - 1,
                        dav1d_filter_2d[t->a->filter[1][bx4 + x + 1]][t->a->filter[0][bx4 + x + 1]],
                        h_mul * ow4, v_mul * oh4);
                i++;
            }
            x += step4;
        }
    }

    if (t->bx > t->ts->tiling.col_start)
        for (int i = 0, y = 0; y < h4 && i < imin(b_dim[3], 4); ) {
            // only odd blocks are considered for overlap handling, hence +1
            const refmvs_block *const l_r = &r[y + 1][t->bx - 1];
            const uint8_t *const l_b_dim = dav1d_block_dimensions[l_r->bs];
            const int step4 = iclip(l_b_dim[1], 2, 16);

            if (l_r->ref.ref[0] > 0) {
                const int ow4 = imin(b_dim[0], 16) >> 1;
                const int oh4 = imin(step4, b_dim[1]);
                push_mc(B, f, 2, dst_x, dst_y + y * v_mul, ow4, oh4, t->bx, t->by + y, pl,
                        l_r->mv.mv[0], l_r->ref.ref[0] - 1,
                        dav1d_filter_2d[t->l.filter[1][by4 + y + 1]][t->l.filter[0][by4 + y + 1]],
                        h_mul * ow4, v_mul * oh4);
                i++;
            }
            y += step4;
        }
}

/* ---- warp_affine(): src/recon.rs:2311-2400 (C: src/recon_tmpl.c:1139-1198).  One record per block; the kernel
 * walks the 8x8s of every plane whose block is at least 8x8 (the reference's imin(cbw4, cbh4) > 1 test). ---- */
static void push_warp(RbHostBatch *const B, const Dav1dTaskContext *const t, const uint8_t *const b_dim, const int refidx,
                      const Dav1dWarpedMotionParams *const wmp)
{
    const Dav1dFrameContext *const f = t->f;
    const Dav1dThreadPicture *const refp = &f->refp[refidx];
    if (refp->p.p.w != f->cur.p.w || refp->p.p.h != f->cur.p.h) {
        rb_batch_unsupported(B, "warped block from a reference of another size");
        return;
    }
    Rb200WarpItem *const it = RB_PUSH(B->warp);
    memset(it, 0, sizeof(*it));
    it->x = (int16_t)(t->bx * 4); it->y = (int16_t)(t->by * 4);
    it->w = (uint8_t)(b_dim[0] * 4); it->h = (uint8_t)(b_dim[1] * 4);
    it->ref = (uint8_t)refidx;
    for (int i = 0; i < 6; i++) it->matrix[i] = wmp->matrix[i];
    for (int i = 0; i < 4; i++) it->abcd[i] = wmp->u.abcd[i];
}

/* The intra half of an inter-intra block (src/recon.rs:3475-3550 luma, :3742-3850 chroma): one wavefront item per plane. */
static void push_interintra(RbHostBatch *const B, const Dav1dTaskContext *const t, const Av1Block *const b, const int pl,
                            const int tw4, const int th4)
{
    const Dav1dFrameContext *const f = t->f;
    const Dav1dTileState *const ts = t->ts;
    const int ss_ver = plane_ss_ver(f, pl), ss_hor = plane_ss_hor(f, pl);
    Rb200IntraItem *const it = RB_PUSH(B->intra);
    memset(it, 0, sizeof(*it));
    it->x4 = (uint16_t)(t->bx >> ss_hor); it->y4 = (uint16_t)(t->by >> ss_ver);
    it->w4_end = (uint16_t)(ts->tiling.col_end >> ss_hor); it->h4_end = (uint16_t)(ts->tiling.row_end >> ss_ver);
    it->plane = (uint8_t)pl; it->tw4 = (uint8_t)tw4; it->th4 = (uint8_t)th4;
    it->mode = b->interintra_mode == II_SMOOTH_PRED ? SMOOTH_PRED : b->interintra_mode;
    it->angle = b->interintra_type == INTER_INTRA_BLEND ? -1 : (int8_t)b->wedge_idx;
    it->flags = (uint8_t)(((t->bx >> ss_hor) > (ts->tiling.col_start >> ss_hor) ? II_HAVE_LEFT : 0) |
                          ((t->by >> ss_ver) > (ts->tiling.row_start >> ss_ver) ? II_HAVE_TOP : 0) | II_INTER_INTRA);
    *RB_PUSH(B->intra_itx) = -1;
}

/* read_coef_tree, pass 2 (C: src/recon_tmpl.c:726-824): walk the transform split tree of an inter block */
static void push_coef_tree(RbHostBatch *const B, Dav1dTaskContext *const t, const int as_intra,
                           const enum RectTxfmSize ytx, const int depth, const uint16_t *const tx_split,
                           const int x_off, const int y_off)
{
    const Dav1dFrameContext *const f = t->f;
    Dav1dTileState *const ts = t->ts;
    const TxfmInfo *const t_dim = &dav1d_txfm_dimensions[ytx];
    const int txw = t_dim->w, txh = t_dim->h;

    if (depth < 2 && tx_split[depth] &&
        tx_split[depth] & (1 << (y_off * 4 + x_off)))
    {
        const enum RectTxfmSize sub = t_dim->sub;
        const TxfmInfo *const sub_t_dim = &dav1d_txfm_dimensions[sub];
        const int txsw = sub_t_dim->w, txsh = sub_t_dim->h;

        push_coef_tree(B, t, as_intra, sub, depth + 1, tx_split, x_off * 2 + 0, y_off * 2 + 0);
        t->bx += txsw;
        if (txw >= txh && t->bx < f->bw)
            push_coef_tree(B, t, as_intra, sub, depth + 1, tx_split, x_off * 2 + 1, y_off * 2 + 0);
        t->bx -= txsw;
        t->by += txsh;
        if (txh >= txw && t->by < f->bh) {
            push_coef_tree(B, t, as_intra, sub, depth + 1, tx_split, x_off * 2 + 0, y_off * 2 + 1);
            t->bx += txsw;
            if (txw >= txh && t->bx < f->bw)
                push_coef_tree(B, t, as_intra, sub, depth + 1, tx_split, x_off * 2 + 1, y_off * 2 + 1);
            t->bx -= txsw;
        }
        t->by -= txsh;
    } else {
        const coef *const cf = ts->frame_thread[0].cf;
        ts->frame_thread[0].cf = (coef *)cf + imin(t_dim->w, 8) * imin(t_dim->h, 8) * 16;
        const int cbi = f->frame_thread.cbi[t->by * f->b4_stride + t->bx][0];
        const int eob = cbi >> 5, txtp = cbi & 0x1f;
        if (eob >= 0) {
            if (as_intra) push_intra_residual_only(B, f, ts, cf, 0, t->bx, t->by, ytx, txtp, eob);
            else fill_itx(RB_PUSH(B->itx), f, cf, 0, 4 * t->bx, 4 * t->by, ytx, txtp, eob);
        }
    }
}

/* ------------------------------------------------------------------------------------------------ */
void bytefn(dav1d_recon_b_intra)(Dav1dTaskContext *const t, const enum BlockSize bs,
                                 const enum EdgeFlags intra_edge_flags,
                                 const Av1Block *const b)
{
    Dav1dTileState *const ts = t->ts;
    const Dav1dFrameContext *const f = t->f;
    RbHostBatch *const B = rb_host_block_begin(f);
    const int bx4 = t->bx & 31, by4 = t->by & 31;
    const int ss_ver = f->cur.p.layout == DAV1D_PIXEL_LAYOUT_I420;
    const int ss_hor = f->cur.p.layout != DAV1D_PIXEL_LAYOUT_I444;
    const int cbx4 = bx4 >> ss_hor, cby4 = by4 >> ss_ver;
    const uint8_t *const b_dim = dav1d_block_dimensions[bs];
    const int bw4 = b_dim[0], bh4 = b_dim[1];
    const int w4 = imin(bw4, f->bw - t->bx), h4 = imin(bh4, f->bh - t->by);
    const int cw4 = (w4 + ss_hor) >> ss_hor, ch4 = (h4 + ss_ver) >> ss_ver;
    const int has_chroma = f->cur.p.layout != DAV1D_PIXEL_LAYOUT_I400 &&
                           (bw4 > ss_hor || t->bx & 1) &&
                           (bh4 > ss_ver || t->by & 1);
    const TxfmInfo *const t_dim = &dav1d_txfm_dimensions[b->tx];
    const TxfmInfo *const uv_t_dim = &dav1d_txfm_dimensions[b->uvtx];
    const int cbw4 = (bw4 + ss_hor) >> ss_hor, cbh4 = (bh4 + ss_ver) >> ss_ver;
    const int edge_filter = f->seq_hdr->intra_edge_filter ? II_EDGE_FILTER : 0;

    for (int init_y = 0; init_y < h4; init_y += 16) {
        const int sub_h4 = imin(h4, 16 + init_y);
        const int sub_ch4 = imin(ch4, (init_y + 16) >> ss_ver);
        for (int init_x = 0; init_x < w4; init_x += 16) {
            if (b->pal_sz[0]) {
                /* pal_pred over the whole block: a palette record { pal[8] padded to 16 bytes, bw x bh index bytes } */
                const uint8_t *const pal_idx = ts->frame_thread[0].pal_idx;
                ts->frame_thread[0].pal_idx += bw4 * bh4 * 16;
                const pixel *const pal =
                    ((pixel (*)[3][8])f->frame_thread.pal)[((t->by >> 1) + (t->bx & 1)) * (f->b4_stride >> 1) +
                                                           ((t->bx >> 1) + (t->by & 1))][0];
                rb_host_push_palette(B, t->bx, t->by, ts->tiling.col_end, ts->tiling.row_end, 0, bw4, bh4,
                                     pal, 8 * sizeof(pixel), pal_idx);
            }

            const int sm = (sm_flag(t->a, bx4) | sm_flag(&t->l, by4)) ? II_SMOOTH : 0;
            const int sb_has_tr = init_x + 16 < w4 ? 1 : init_y ? 0 :
                              intra_edge_flags & EDGE_I444_TOP_HAS_RIGHT;
            const int sb_has_bl = init_x ? 0 : init_y + 16 < h4 ? 1 :
                              intra_edge_flags & EDGE_I444_LEFT_HAS_BOTTOM;
            int y, x;
            const int sub_w4 = imin(w4, init_x + 16);
            for (y = init_y, t->by += init_y; y < sub_h4;
                 y += t_dim->h, t->by += t_dim->h)
            {
                for (x = init_x, t->bx += init_x; x < sub_w4;
                     x += t_dim->w, t->bx += t_dim->w)
                {
                    const coef *cf = NULL;
                    int eob = -1, txtp = 0;
                    if (!b->skip) {
                        cf = ts->frame_thread[0].cf;
                        ts->frame_thread[0].cf += imin(t_dim->w, 8) * imin(t_dim->h, 8) * 16;
                        const int cbi = f->frame_thread.cbi[t->by * f->b4_stride + t->bx][0];
                        eob  = cbi >> 5;
                        txtp = cbi & 0x1f;
                    }
                    if (b->pal_sz[0]) {
                        if (eob >= 0) push_intra_residual_only(B, f, ts, cf, 0, t->bx, t->by, b->tx, txtp, eob);
                        continue;
                    }
                    const enum EdgeFlags edge_flags =
                        (((y > init_y || !sb_has_tr) && (x + t_dim->w >= sub_w4)) ?
                             0 : EDGE_I444_TOP_HAS_RIGHT) |
                        ((x > init_x || (!sb_has_bl && y + t_dim->h >= sub_h4)) ?
                             0 : EDGE_I444_LEFT_HAS_BOTTOM);
                    Rb200IntraItem *const it = RB_PUSH(B->intra);
                    memset(it, 0, sizeof(*it));
                    it->x4 = (uint16_t)t->bx; it->y4 = (uint16_t)t->by;
                    it->w4_end = (uint16_t)ts->tiling.col_end; it->h4_end = (uint16_t)ts->tiling.row_end;
                    it->plane = 0; it->tw4 = t_dim->w; it->th4 = t_dim->h;
                    it->mode = b->y_mode; it->angle = b->y_angle;
                    it->flags = (uint8_t)((t->bx > ts->tiling.col_start ? II_HAVE_LEFT : 0) |
                                          (t->by > ts->tiling.row_start ? II_HAVE_TOP : 0) |
                                          (edge_flags & EDGE_I444_TOP_HAS_RIGHT ? II_TOP_HAS_RIGHT : 0) |
                                          (edge_flags & EDGE_I444_LEFT_HAS_BOTTOM ? II_LEFT_HAS_BOTTOM : 0) |
                                          sm | edge_filter);
                    if (eob >= 0) {
                        *RB_PUSH(B->intra_itx) = B->iitx.n;
                        fill_itx(RB_PUSH(B->iitx), f, cf, 0, 4 * t->bx, 4 * t->by, b->tx, txtp, eob);
                    } else {
                        *RB_PUSH(B->intra_itx) = -1;
                    }
                }
                t->bx -= x;
            }
            t->by -= y;

            if (!has_chroma) continue;

            if (b->uv_mode == CFL_PRED) {
                /* cfl_ac over the block's luma + cfl_pred per plane with a non-zero alpha: one item per plane */
                const int furthest_r =
                    ((cw4 << ss_hor) + t_dim->w - 1) & ~(t_dim->w - 1);
                const int furthest_b =
                    ((ch4 << ss_ver) + t_dim->h - 1) & ~(t_dim->h - 1);
                const int w_pad = cbw4 - (furthest_r >> ss_hor), h_pad = cbh4 - (furthest_b >> ss_ver);
                if (uv_t_dim->w != cbw4 || uv_t_dim->h != cbh4) {
                    /* (lossless blocks: the reference predicts one 4x4 from an AC buffer of the whole block) */
                    if (b->cfl_alpha[0] || b->cfl_alpha[1])
                        rb_batch_unsupported(B, "chroma-from-luma block whose chroma transform is smaller than the block");
                }
                for (int pl = 0; pl < 2; pl++) {
                    if (!b->cfl_alpha[pl]) continue;
                    const int xpos = t->bx >> ss_hor, ypos = t->by >> ss_ver;
                    const int xstart = ts->tiling.col_start >> ss_hor;
                    const int ystart = ts->tiling.row_start >> ss_ver;
                    Rb200IntraItem *const it = RB_PUSH(B->intra);
                    memset(it, 0, sizeof(*it));
                    it->x4 = (uint16_t)xpos; it->y4 = (uint16_t)ypos;
                    it->w4_end = (uint16_t)((ts->tiling.col_end >> ss_hor) | (w_pad << 13));
                    it->h4_end = (uint16_t)((ts->tiling.row_end >> ss_ver) | (h_pad << 13));
                    it->plane = (uint8_t)(1 + pl); it->tw4 = uv_t_dim->w; it->th4 = uv_t_dim->h;
                    it->mode = MODE_CFL; it->angle = b->cfl_alpha[pl];
                    it->flags = (uint8_t)((xpos > xstart ? II_HAVE_LEFT : 0) | (ypos > ystart ? II_HAVE_TOP : 0));
                    *RB_PUSH(B->intra_itx) = -1;
                }
            } else if (b->pal_sz[1]) {
                const pixel (*const pal)[8] =
                    ((pixel (*)[3][8])f->frame_thread.pal)[((t->by >> 1) + (t->bx & 1)) * (f->b4_stride >> 1) +
                                                           ((t->bx >> 1) + (t->by & 1))];
                const uint8_t *const pal_idx = ts->frame_thread[0].pal_idx;
                ts->frame_thread[0].pal_idx += cbw4 * cbh4 * 16;
                for (int pl = 1; pl <= 2; pl++)
                    rb_host_push_palette(B, t->bx >> ss_hor, t->by >> ss_ver, ts->tiling.col_end >> ss_hor,
                                         ts->tiling.row_end >> ss_ver, pl, cbw4, cbh4, pal[pl], 8 * sizeof(pixel), pal_idx);
            }

            const int sm_uv = (sm_uv_flag(t->a, cbx4) | sm_uv_flag(&t->l, cby4)) ? II_SMOOTH : 0;
            const int uv_sb_has_tr =
                ((init_x + 16) >> ss_hor) < cw4 ? 1 : init_y ? 0 :
                intra_edge_flags & (EDGE_I420_TOP_HAS_RIGHT >> (f->cur.p.layout - 1));
            const int uv_sb_has_bl =
                init_x ? 0 : ((init_y + 16) >> ss_ver) < ch4 ? 1 :
                intra_edge_flags & (EDGE_I420_LEFT_HAS_BOTTOM >> (f->cur.p.layout - 1));
            const int sub_cw4 = imin(cw4, (init_x + 16) >> ss_hor);
            for (int pl = 0; pl < 2; pl++) {
                for (y = init_y >> ss_ver, t->by += init_y; y < sub_ch4;
                     y += uv_t_dim->h, t->by += uv_t_dim->h << ss_ver)
                {
                    for (x = init_x >> ss_hor, t->bx += init_x; x < sub_cw4;
                         x += uv_t_dim->w, t->bx += uv_t_dim->w << ss_hor)
                    {
                        const coef *cf = NULL;
                        int eob = -1, txtp = 0;
                        if (!b->skip) {
                            cf = ts->frame_thread[0].cf;
                            ts->frame_thread[0].cf += uv_t_dim->w * uv_t_dim->h * 16;
                            const int cbi = f->frame_thread.cbi[t->by * f->b4_stride + t->bx][pl + 1];
                            eob  = cbi >> 5;
                            txtp = cbi & 0x1f;
                        }
                        const int xpos = t->bx >> ss_hor, ypos = t->by >> ss_ver;
                        if ((b->uv_mode == CFL_PRED && b->cfl_alpha[pl]) || b->pal_sz[1]) {
                            if (eob >= 0) push_intra_residual_only(B, f, ts, cf, 1 + pl, xpos, ypos, b->uvtx, txtp, eob);
                            continue;
                        }
                        // this probably looks weird because we're using
                        // luma flags in a chroma loop, but that's because
                        // prepare_intra_edges() expects luma flags as input
                        const enum EdgeFlags edge_flags =
                            (((y > (init_y >> ss_ver) || !uv_sb_has_tr) &&
                              (x + uv_t_dim->w >= sub_cw4)) ?
                                 0 : EDGE_I444_TOP_HAS_RIGHT) |
                            ((x > (init_x >> ss_hor) ||
                              (!uv_sb_has_bl && y + uv_t_dim->h >= sub_ch4)) ?
                                 0 : EDGE_I444_LEFT_HAS_BOTTOM);
                        const enum IntraPredMode uv_mode =
                             b->uv_mode == CFL_PRED ? DC_PRED : b->uv_mode;
                        const int xstart = ts->tiling.col_start >> ss_hor;
                        const int ystart = ts->tiling.row_start >> ss_ver;
                        Rb200IntraItem *const it = RB_PUSH(B->intra);
                        memset(it, 0, sizeof(*it));
                        it->x4 = (uint16_t)xpos; it->y4 = (uint16_t)ypos;
                        it->w4_end = (uint16_t)(ts->tiling.col_end >> ss_hor);
                        it->h4_end = (uint16_t)(ts->tiling.row_end >> ss_ver);
                        it->plane = (uint8_t)(1 + pl); it->tw4 = uv_t_dim->w; it->th4 = uv_t_dim->h;
                        it->mode = (uint8_t)uv_mode; it->angle = b->uv_angle;
                        it->flags = (uint8_t)((xpos > xstart ? II_HAVE_LEFT : 0) | (ypos > ystart ? II_HAVE_TOP : 0) |
                                              (edge_flags & EDGE_I444_TOP_HAS_RIGHT ? II_TOP_HAS_RIGHT : 0) |
                                              (edge_flags & EDGE_I444_LEFT_HAS_BOTTOM ? II_LEFT_HAS_BOTTOM : 0) |
                                              sm_uv | edge_filter);
                        if (eob >= 0) {
                            *RB_PUSH(B->intra_itx) = B->iitx.n;
                            fill_itx(RB_PUSH(B->iitx), f, cf, 1 + pl, 4 * xpos, 4 * ypos, b->uvtx, txtp, eob);
                        } else {
                            *RB_PUSH(B->intra_itx) = -1;
                        }
                    }
                    t->bx -= x << ss_hor;
                }
                t->by -= y << ss_ver;
            }
        }
    }
    rb_host_block_end(f);
}

int bytefn(dav1d_recon_b_inter)(Dav1dTaskContext *const t, const enum BlockSize bs,
                                const Av1Block *const b)
{
    Dav1dTileState *const ts = t->ts;
    const Dav1dFrameContext *const f = t->f;
    RbHostBatch *const B = rb_host_block_begin(f);
    const int bx4 = t->bx & 31, by4 = t->by & 31;
    const int ss_ver = f->cur.p.layout == DAV1D_PIXEL_LAYOUT_I420;
    const int ss_hor = f->cur.p.layout != DAV1D_PIXEL_LAYOUT_I444;
    const uint8_t *const b_dim = dav1d_block_dimensions[bs];
    const int bw4 = b_dim[0], bh4 = b_dim[1];
    const int w4 = imin(bw4, f->bw - t->bx), h4 = imin(bh4, f->bh - t->by);
    const int has_chroma = f->cur.p.layout != DAV1D_PIXEL_LAYOUT_I400 &&
                           (bw4 > ss_hor || t->bx & 1) &&
                           (bh4 > ss_ver || t->by & 1);
    const int cbh4 = (bh4 + ss_ver) >> ss_ver, cbw4 = (bw4 + ss_hor) >> ss_hor;
    /* destinations in plane pixels */
    const int dx0 = 4 * t->bx, dy0 = 4 * t->by;
    const int cdx0 = 4 * (t->bx >> ss_hor), cdy0 = 4 * (t->by >> ss_ver);
    int is_interintra = 0;

    // prediction
    if (IS_KEY_OR_INTRA(f->frame_hdr)) {
        /* intra block copy (src/recon_tmpl.c:1631-1645) predicts from the picture being reconstructed: it joins the intra
         * wavefront, one level above whatever wrote its source area, and its residual follows there too */
        push_intrabc(B, f, 0, t->bx, t->by, bw4, bh4, t->bx, t->by, b->mv[0]);
        if (has_chroma) for (int pl = 1; pl < 3; pl++)
            push_intrabc(B, f, pl, t->bx >> ss_hor, t->by >> ss_ver, bw4 << (bw4 == ss_hor), bh4 << (bh4 == ss_ver),
                         t->bx & ~ss_hor, t->by & ~ss_ver, b->mv[0]);
        is_interintra = 1;
    } else if (b->comp_type == COMP_INTER_NONE) {
        const enum Filter2d filter_2d = b->filter2d;
        const int warped = (b->inter_mode == GLOBALMV && f->gmv_warp_allowed[b->ref[0]]) ||
                           (b->motion_mode == MM_WARP && t->warpmv.type > DAV1D_WM_TYPE_TRANSLATION);
        const Dav1dWarpedMotionParams *const wmp = b->motion_mode == MM_WARP ? &t->warpmv : &f->frame_hdr->gmv[b->ref[0]];
        const int warp_luma = imin(bw4, bh4) > 1 && warped;

        if (warp_luma) {
            push_warp(B, t, b_dim, b->ref[0], wmp);
        } else {
            push_mc(B, f, 0, dx0, dy0, bw4, bh4, t->bx, t->by, 0, b->mv[0], b->ref[0], filter_2d, 0, 0);
            if (b->motion_mode == MM_OBMC)
                push_obmc(B, t, dx0, dy0, b_dim, 0, bx4, by4, w4, h4);
        }
        if (b->interintra_type) {
            is_interintra = 1;
            push_interintra(B, t, b, 0, bw4, bh4);
        }

        if (!has_chroma) goto skip_inter_chroma_pred;

        // sub8x8 derivation
        int is_sub8x8 = bw4 == ss_hor || bh4 == ss_ver;
        refmvs_block *const *r;
        if (is_sub8x8) {
            r = &t->rt.r[(t->by & 31) + 5];
            if (bw4 == 1) is_sub8x8 &= r[0][t->bx - 1].ref.ref[0] > 0;
            if (bh4 == ss_ver) is_sub8x8 &= r[-1][t->bx].ref.ref[0] > 0;
            if (bw4 == 1 && bh4 == ss_ver)
                is_sub8x8 &= r[-1][t->bx - 1].ref.ref[0] > 0;
        }

        // chroma prediction
        if (is_sub8x8) {
            /* the chroma block of a sub-8x8 luma block is assembled from the vectors of the up to four luma
             * blocks it covers (src/recon.rs:3552-3700) */
            int h_off = 0, v_off = 0;
            if (bw4 == 1 && bh4 == ss_ver) {
                for (int pl = 0; pl < 2; pl++)
                    push_mc(B, f, 0, cdx0, cdy0, bw4, bh4, t->bx - 1, t->by - 1, 1 + pl,
                            r[-1][t->bx - 1].mv.mv[0], r[-1][t->bx - 1].ref.ref[0] - 1,
                            f->frame_thread.b[((t->by - 1) * f->b4_stride) + t->bx - 1].filter2d, 0, 0);
                v_off = 2;
                h_off = 2;
            }
            if (bw4 == 1) {
                for (int pl = 0; pl < 2; pl++)
                    push_mc(B, f, 0, cdx0, cdy0 + v_off, bw4, bh4, t->bx - 1, t->by, 1 + pl,
                            r[0][t->bx - 1].mv.mv[0], r[0][t->bx - 1].ref.ref[0] - 1,
                            f->frame_thread.b[(t->by * f->b4_stride) + t->bx - 1].filter2d, 0, 0);
                h_off = 2;
            }
            if (bh4 == ss_ver) {
                for (int pl = 0; pl < 2; pl++)
                    push_mc(B, f, 0, cdx0 + h_off, cdy0, bw4, bh4, t->bx, t->by - 1, 1 + pl,
                            r[-1][t->bx].mv.mv[0], r[-1][t->bx].ref.ref[0] - 1,
                            f->frame_thread.b[((t->by - 1) * f->b4_stride) + t->bx].filter2d, 0, 0);
                v_off = 2;
            }
            for (int pl = 0; pl < 2; pl++)
                push_mc(B, f, 0, cdx0 + h_off, cdy0 + v_off, bw4, bh4, t->bx, t->by, 1 + pl, b->mv[0],
                        b->ref[0], filter_2d, 0, 0);
        } else {
            /* chroma planes of at least 8x8 belong to the block's warp record (imin(cbw4, cbh4) > 1 implies that luma
             * was warped too); smaller chroma blocks of a warped block are translated like any other */
            if (!(imin(cbw4, cbh4) > 1 && warped)) {
                for (int pl = 0; pl < 2; pl++) {
                    push_mc(B, f, 0, cdx0, cdy0, bw4 << (bw4 == ss_hor), bh4 << (bh4 == ss_ver),
                            t->bx & ~ss_hor, t->by & ~ss_ver, 1 + pl, b->mv[0], b->ref[0], filter_2d, 0, 0);
                    if (b->motion_mode == MM_OBMC)
                        push_obmc(B, t, cdx0, cdy0, b_dim, 1 + pl, bx4, by4, w4, h4);
                }
            }
            if (b->interintra_type)
                for (int pl = 0; pl < 2; pl++)
                    push_interintra(B, t, b, 1 + pl, cbw4, cbh4);
        }

    skip_inter_chroma_pred: {}
    } else {
        /* compound: one record per block; both predictions, the blend and the mask stay on the device */
        const int same_size[2] = {
            f->refp[b->ref[0]].p.p.w == f->cur.p.w && f->refp[b->ref[0]].p.p.h == f->cur.p.h,
            f->refp[b->ref[1]].p.p.w == f->cur.p.w && f->refp[b->ref[1]].p.p.h == f->cur.p.h };
        /* a prediction of a GLOBALMV_GLOBALMV block is the reference's global-motion warp where the reference allows it
         * (src/recon.rs:3253-3268; chroma only when the chroma block is at least 8x8, :3352-3369) */
        const int gmv_warp[2] = { b->inter_mode == GLOBALMV_GLOBALMV && f->gmv_warp_allowed[b->ref[0]],
                                  b->inter_mode == GLOBALMV_GLOBALMV && f->gmv_warp_allowed[b->ref[1]] };
        if ((gmv_warp[0] && !same_size[0]) || (gmv_warp[1] && !same_size[1])) {
            rb_batch_unsupported(B, "warped block from a reference of another size");
        } else {      /* (predictions from references of another size: the kernel scales them, rb200_frame_set_ref_size) */
            Rb200CompItem *const it = RB_PUSH(B->comp);
            memset(it, 0, sizeof(*it));
            it->x = (int16_t)dx0; it->y = (int16_t)dy0;
            it->w = (uint8_t)(bw4 * 4); it->h = (uint8_t)(bh4 * 4);
            for (int i = 0; i < 2; i++) {
                it->ref[i] = (uint8_t)b->ref[i];
                it->mv[i][0] = b->mv[i].y; it->mv[i][1] = b->mv[i].x;
            }
            it->filter2d = b->filter2d;
            {
                const int chroma_8x8 = imin(cbw4, cbh4) > 1;
                it->warp_mask = (uint8_t)(gmv_warp[0] | gmv_warp[1] << 1 | (gmv_warp[0] && chroma_8x8) << 2 | (gmv_warp[1] && chroma_8x8) << 3);
            }
            switch (b->comp_type) {
            case COMP_INTER_AVG: it->comp_type = RB200_COMP_AVG; break;
            case COMP_INTER_WEIGHTED_AVG:
                it->comp_type = RB200_COMP_WEIGHTED_AVG;
                it->jnt_weight = f->jnt_weights[b->ref[0]][b->ref[1]];
                break;
            case COMP_INTER_SEG: it->comp_type = RB200_COMP_SEG; it->mask_sign = b->mask_sign; break;
            case COMP_INTER_WEDGE:
                it->comp_type = RB200_COMP_WEDGE; it->mask_sign = b->mask_sign; it->wedge_idx = b->wedge_idx;
                break;
            }
        }
    }

    if (b->skip) {
        rb_host_block_end(f);
        return 0;
    }

    const int cw4 = (w4 + ss_hor) >> ss_hor, ch4 = (h4 + ss_ver) >> ss_ver;
    const TxfmInfo *const uvtx = &dav1d_txfm_dimensions[b->uvtx];
    const TxfmInfo *const ytx = &dav1d_txfm_dimensions[b->max_ytx];
    const uint16_t tx_split[2] = { b->tx_split0, b->tx_split1 };

    for (int init_y = 0; init_y < bh4; init_y += 16) {
        for (int init_x = 0; init_x < bw4; init_x += 16) {
            // coefficient coding & inverse transforms
            int y_off = !!init_y, y;
            for (y = init_y, t->by += init_y; y < imin(h4, init_y + 16);
                 y += ytx->h, y_off++)
            {
                int x, x_off = !!init_x;
                for (x = init_x, t->bx += init_x; x < imin(w4, init_x + 16);
                     x += ytx->w, x_off++)
                {
                    push_coef_tree(B, t, is_interintra, b->max_ytx, 0, tx_split, x_off, y_off);
                    t->bx += ytx->w;
                }
                t->bx -= x;
                t->by += ytx->h;
            }
            t->by -= y;

            // chroma coefs and inverse transform
            if (has_chroma) for (int pl = 0; pl < 2; pl++) {
                for (y = init_y >> ss_ver, t->by += init_y;
                     y < imin(ch4, (init_y + 16) >> ss_ver); y += uvtx->h)
                {
                    int x;
                    for (x = init_x >> ss_hor, t->bx += init_x;
                         x < imin(cw4, (init_x + 16) >> ss_hor); x += uvtx->w)
                    {
                        const coef *const cf = ts->frame_thread[0].cf;
                        ts->frame_thread[0].cf += uvtx->w * uvtx->h * 16;
                        const int cbi = f->frame_thread.cbi[t->by * f->b4_stride + t->bx][pl + 1];
                        const int eob = cbi >> 5, txtp = cbi & 0x1f;
                        if (eob >= 0) {
                            const int px = cdx0 + 4 * x, py = cdy0 + 4 * y;
                            if (is_interintra) push_intra_residual_only(B, f, ts, cf, 1 + pl, px >> 2, py >> 2, b->uvtx, txtp, eob);
                            else fill_itx(RB_PUSH(B->itx), f, cf, 1 + pl, px, py, b->uvtx, txtp, eob);
                        }
                        t->bx += uvtx->w << ss_hor;
                    }
                    t->bx -= x << ss_hor;
                    t->by += uvtx->h << ss_ver;
                }
                t->by -= y << ss_ver;
            }
        }
    }
    rb_host_block_end(f);
    return 0;
}

/* The filters run as whole-frame launches after the frame's last block was appended (host_frame.c); per superblock
 * row there is nothing to do, and nothing to back up: reconstruction is complete before any filter runs and CDEF /
 * loop restoration are out of place on the device, so neither ipred_edge nor the lpf line buffers exist. */
void bytefn(dav1d_filter_sbrow_deblock_cols)(Dav1dFrameContext *const f, const int sby) { (void)f; (void)sby; }
void bytefn(dav1d_filter_sbrow_deblock_rows)(Dav1dFrameContext *const f, const int sby) { (void)f; (void)sby; }
void bytefn(dav1d_filter_sbrow_cdef)(Dav1dTaskContext *const tc, const int sby) { (void)tc; (void)sby; }
void bytefn(dav1d_filter_sbrow_resize)(Dav1dFrameContext *const f, const int sby) { (void)f; (void)sby; }
void bytefn(dav1d_filter_sbrow_lr)(Dav1dFrameContext *const f, const int sby) { (void)f; (void)sby; }
void bytefn(dav1d_filter_sbrow)(Dav1dFrameContext *const f, const int sby) { (void)f; (void)sby; }
void bytefn(dav1d_backup_ipred_edge)(Dav1dTaskContext *const t) { (void)t; }
