/*
 * oracle/ref_frame.c -- TEST INFRASTRUCTURE ONLY (linked into oracle/_ref/libdav1d_ref.so).
 *
 * Drives the REFERENCE'S OWN frame-level drivers over a synthetic frame:
 *   dav1d_filter_sbrow_{deblock_cols,deblock_rows,cdef,lr}_{8,16}bpc  (src/recon_tmpl.c:2053-2176)
 *     -> dav1d_loopfilter_sbrow_cols/rows, dav1d_copy_lpf        (src/lf_apply_tmpl.c)
 *     -> dav1d_cdef_brow                                        (src/cdef_apply_tmpl.c:98)
 *     -> dav1d_lr_sbrow                                         (src/lr_apply_tmpl.c:162)
 * by building the minimal Dav1dContext / Dav1dFrameContext those functions read
 * (buffer set-up restated from dav1d_decode_frame_init, src/decode.c:2911-3004).
 * Reconstruction (prediction + residual) is replayed item by item through the
 * reference's DSP tables with the addressing rules of recon_tmpl.c `mc()`
 * (src/recon_tmpl.c:1036-1106: emu_edge when the block window leaves the picture).
 *
 * Two schedules:
 *   n_threads == 1: the reference's single-thread order, dav1d_filter_sbrow(f, sby) per sbrow
 *                   (n_tc == 1: rolling 12-line lpf buffer, src/decode.c:2972);
 *   n_threads  > 1: the reference's tile-thread buffers (n_tc > 1: per-sbrow lpf / cdef
 *                   lines) with every stage run as a parallel-for over superblock rows and
 *                   a join between stages -- the dependency structure of
 *                   src/thread_task.c:761-830 with the wavefront flattened.  This is the CPU
 *                   baseline bench.py times.
 */
#include "config.h"

#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "src/internal.h"
#include "src/wedge.h"
#include "src/lf_mask.h"
#include "src/levels.h"
#include "src/tables.h"
#include "src/recon.h"

#include "../include/rav1d_b200.h"

#include "ref_frame.h"

void ref_init(void);
void dav1d_filter_sbrow_8bpc(Dav1dFrameContext *f, int sby);
void dav1d_filter_sbrow_16bpc(Dav1dFrameContext *f, int sby);
void dav1d_filter_sbrow_deblock_cols_8bpc(Dav1dFrameContext *f, int sby);
void dav1d_filter_sbrow_deblock_cols_16bpc(Dav1dFrameContext *f, int sby);
void dav1d_filter_sbrow_deblock_rows_8bpc(Dav1dFrameContext *f, int sby);
void dav1d_filter_sbrow_deblock_rows_16bpc(Dav1dFrameContext *f, int sby);
void dav1d_filter_sbrow_cdef_8bpc(Dav1dTaskContext *tc, int sby);
void dav1d_filter_sbrow_cdef_16bpc(Dav1dTaskContext *tc, int sby);
void dav1d_filter_sbrow_lr_8bpc(Dav1dFrameContext *f, int sby);
void dav1d_filter_sbrow_lr_16bpc(Dav1dFrameContext *f, int sby);
void dav1d_film_grain_dsp_init_8bpc(Dav1dFilmGrainDSPContext *c);
void dav1d_film_grain_dsp_init_16bpc(Dav1dFilmGrainDSPContext *c);
void dav1d_intra_pred_dsp_init_8bpc(Dav1dIntraPredDSPContext *c);
void dav1d_intra_pred_dsp_init_16bpc(Dav1dIntraPredDSPContext *c);
void dav1d_apply_grain_8bpc(const Dav1dFilmGrainDSPContext *dsp, Dav1dPicture *out, const Dav1dPicture *in);
void dav1d_apply_grain_16bpc(const Dav1dFilmGrainDSPContext *dsp, Dav1dPicture *out, const Dav1dPicture *in);
void dav1d_itx_dsp_init_8bpc(Dav1dInvTxfmDSPContext *c, int bpc);
void dav1d_itx_dsp_init_16bpc(Dav1dInvTxfmDSPContext *c, int bpc);
void dav1d_mc_dsp_init_8bpc(Dav1dMCDSPContext *c);
void dav1d_mc_dsp_init_16bpc(Dav1dMCDSPContext *c);
void dav1d_loop_filter_dsp_init_8bpc(Dav1dLoopFilterDSPContext *c);
void dav1d_loop_filter_dsp_init_16bpc(Dav1dLoopFilterDSPContext *c);
void dav1d_cdef_dsp_init_8bpc(Dav1dCdefDSPContext *c);
void dav1d_cdef_dsp_init_16bpc(Dav1dCdefDSPContext *c);
void dav1d_loop_restoration_dsp_init_8bpc(Dav1dLoopRestorationDSPContext *c, int bpc);
void dav1d_loop_restoration_dsp_init_16bpc(Dav1dLoopRestorationDSPContext *c, int bpc);

static void *zalloc(size_t n) { void *p = NULL; if (posix_memalign(&p, 64, n ? n : 64)) return NULL; memset(p, 0, n); return p; }

void ref_frame_free(RefFrame *r) {
    if (!r) return;
    if (r->f) {
        free(r->f->lf.cdef_line_buf); free(r->f->lf.lr_line_buf); free(r->f->lf.mask); free(r->f->lf.lr_mask);
        free(r->f->lf.tx_lpf_right_edge[0]);
    }
    free(r->grain_mem); free(r->lvl_mem); free(r->plane_mem); free(r->tc); free(r->f); free(r->c); free(r);
}

/* layout: 0 I400, 1 I420, 2 I422, 3 I444 (same values as Dav1dPixelLayout) */
RefFrame *ref_frame_new(const Rb200FrameHeader *h, int n_tc) {
    ref_init();
    RefFrame *r = zalloc(sizeof(*r));
    Dav1dContext *c = r->c = zalloc(sizeof(Dav1dContext));
    Dav1dFrameContext *f = r->f = zalloc(sizeof(Dav1dFrameContext));
    if (n_tc < 1) n_tc = 1;
    r->n_tc = n_tc;
    r->tc = zalloc(sizeof(Dav1dTaskContext) * n_tc);
    r->hbd = h->bpc > 8;
    r->bdmax = (1 << h->bpc) - 1;
    for (int i = 0; i < n_tc; i++) { r->tc[i].c = c; r->tc[i].f = f; }
    c->tc = r->tc; c->n_tc = n_tc; c->fc = f; c->n_fc = 1;
    c->inloop_filters = DAV1D_INLOOPFILTER_ALL;
    Dav1dDSPContext *dsp = &c->dsp[(h->bpc >> 1) - 4];
    if (r->hbd) {
        dav1d_itx_dsp_init_16bpc(&dsp->itx, h->bpc); dav1d_mc_dsp_init_16bpc(&dsp->mc);
        dav1d_loop_filter_dsp_init_16bpc(&dsp->lf); dav1d_cdef_dsp_init_16bpc(&dsp->cdef);
        dav1d_loop_restoration_dsp_init_16bpc(&dsp->lr, h->bpc);
        dav1d_film_grain_dsp_init_16bpc(&dsp->fg); dav1d_intra_pred_dsp_init_16bpc(&dsp->ipred);
    } else {
        dav1d_film_grain_dsp_init_8bpc(&dsp->fg); dav1d_intra_pred_dsp_init_8bpc(&dsp->ipred);
        dav1d_itx_dsp_init_8bpc(&dsp->itx, 8); dav1d_mc_dsp_init_8bpc(&dsp->mc);
        dav1d_loop_filter_dsp_init_8bpc(&dsp->lf); dav1d_cdef_dsp_init_8bpc(&dsp->cdef);
        dav1d_loop_restoration_dsp_init_8bpc(&dsp->lr, 8);
    }
    f->c = c; f->dsp = dsp; f->seq_hdr = &r->seq; f->frame_hdr = &r->hdr;
    f->bitdepth_max = r->bdmax;
    r->seq.sb128 = h->sb128; r->seq.cdef = 1; r->seq.restoration = 1;
    r->seq.layout = h->layout; r->seq.hbd = h->bpc == 8 ? 0 : h->bpc == 10 ? 1 : 2;
    r->hdr.width[0] = r->hdr.width[1] = h->width; r->hdr.height = h->height;
    r->hdr.tiling.cols = r->hdr.tiling.rows = 1;
    r->hdr.loopfilter.level_y[0] = h->lf_level_y[0]; r->hdr.loopfilter.level_y[1] = h->lf_level_y[1];
    r->hdr.loopfilter.level_u = h->lf_level_u; r->hdr.loopfilter.level_v = h->lf_level_v;
    r->hdr.cdef.damping = h->cdef_damping;
    for (int i = 0; i < 8; i++) { r->hdr.cdef.y_strength[i] = h->cdef_y_strength[i]; r->hdr.cdef.uv_strength[i] = h->cdef_uv_strength[i]; }
    for (int i = 0; i < 3; i++) r->hdr.restoration.type[i] = h->lr_type[i];
    r->hdr.restoration.unit_size[0] = h->lr_unit_size_log2[0]; r->hdr.restoration.unit_size[1] = h->lr_unit_size_log2[1];

    /* geometry: dav1d_submit_frame, src/decode.c:3560-3580 */
    f->bw = ((h->width + 7) >> 3) << 1; f->bh = ((h->height + 7) >> 3) << 1;
    f->w4 = (h->width + 3) >> 2; f->h4 = (h->height + 3) >> 2;
    f->sb128w = (f->bw + 31) >> 5; f->sb128h = (f->bh + 31) >> 5; f->sr_sb128w = f->sb128w;
    f->sb_shift = 4 + h->sb128; f->sb_step = 16 << h->sb128;
    f->sbh = (f->bh + f->sb_step - 1) >> f->sb_shift;
    f->b4_stride = (f->bw + 31) & ~31;
    /* one tile: col/row_start_sb[n_tiles] = picture size in superblocks (src/obu.c parse_tile_hdr);
     * the loop-filter driver walks col_start_sb[] until it passes the right edge (src/lf_apply_tmpl.c:335) */
    r->hdr.tiling.col_start_sb[0] = 0; r->hdr.tiling.col_start_sb[1] = (f->bw + f->sb_step - 1) >> f->sb_shift;
    r->hdr.tiling.row_start_sb[0] = 0; r->hdr.tiling.row_start_sb[1] = f->sbh;

    /* picture: dav1d_default_picture_alloc, src/picture.c:47-90 */
    const int px = r->hbd ? 2 : 1;
    const int aw = (h->width + 127) & ~127, ah = (h->height + 127) & ~127;
    const int has_chroma = h->layout != 0, ss_ver = h->layout == 1, ss_hor = h->layout != 3;
    ptrdiff_t y_stride = (ptrdiff_t)aw * px, uv_stride = has_chroma ? y_stride >> ss_hor : 0;
    if (!(y_stride & 1023)) y_stride += 64;
    if (has_chroma && !(uv_stride & 1023)) uv_stride += 64;
    const size_t ysz = (size_t)y_stride * ah, uvsz = (size_t)uv_stride * (ah >> ss_ver);
    r->plane_bytes = ysz + 2 * uvsz + 64;
    r->plane_mem = zalloc(r->plane_bytes);
    f->cur.data[0] = r->plane_mem;
    f->cur.data[1] = has_chroma ? r->plane_mem + ysz : NULL;
    f->cur.data[2] = has_chroma ? r->plane_mem + ysz + uvsz : NULL;
    f->cur.stride[0] = y_stride; f->cur.stride[1] = uv_stride;
    f->cur.p.w = h->width; f->cur.p.h = h->height; f->cur.p.layout = h->layout; f->cur.p.bpc = h->bpc;
    f->sr_cur.p = f->cur;
    for (int i = 0; i < 3; i++) f->lf.p[i] = f->lf.sr_p[i] = f->cur.data[i];

    /* filter metadata */
    const int num_sb128 = f->sb128w * f->sb128h;
    f->lf.mask = zalloc(sizeof(Av1Filter) * num_sb128);
    f->lf.lr_mask = zalloc(sizeof(Av1Restoration) * num_sb128);
    r->lvl_mem = zalloc(((size_t)f->b4_stride * 32 * f->sb128h + 32 + 3) * 4);
    f->lf.level = (uint8_t(*)[4])r->lvl_mem + 32;
    f->lf.tx_lpf_right_edge[0] = zalloc((size_t)f->sb128h * 32 * 2 + 64);
    f->lf.tx_lpf_right_edge[1] = f->lf.tx_lpf_right_edge[0] + f->sb128h * 32;
    f->lf.start_of_tile_row = r->start_of_tile_row;
    f->lf.restore_planes = ((h->lr_type[0] != 0) << 0) | ((has_chroma && h->lr_type[1] != 0) << 1) |
                           ((has_chroma && h->lr_type[2] != 0) << 2);

    /* cdef / lr line buffers: src/decode.c:2911-3004 (positive strides, no super-res) */
    {
        size_t alloc_sz = 64 + (size_t)y_stride * 4 * f->sbh + (size_t)uv_stride * 8 * f->sbh;
        uint8_t *ptr = f->lf.cdef_line_buf = zalloc(alloc_sz + 64);
        ptr += 32;
        f->lf.cdef_line[0][0] = ptr; f->lf.cdef_line[1][0] = ptr + y_stride * 2;
        ptr += y_stride * f->sbh * 4;
        f->lf.cdef_line[0][1] = ptr; f->lf.cdef_line[0][2] = ptr + uv_stride * 2;
        f->lf.cdef_line[1][1] = ptr + uv_stride * 4; f->lf.cdef_line[1][2] = ptr + uv_stride * 6;
        const int num_lines = n_tc > 1 ? f->sbh * 4 << h->sb128 : 12;
        alloc_sz = 128 + (size_t)y_stride * num_lines + (size_t)uv_stride * num_lines * 2;
        ptr = f->lf.lr_line_buf = zalloc(alloc_sz + 64);
        ptr += 64;
        f->lf.lr_lpf_line[0] = ptr; ptr += y_stride * num_lines;
        f->lf.lr_lpf_line[1] = ptr; f->lf.lr_lpf_line[2] = ptr + uv_stride * num_lines;
    }
    return r;
}

void *ref_frame_plane(RefFrame *r, int pl) { return r->f->cur.data[pl]; }
ptrdiff_t ref_frame_stride(RefFrame *r, int uv) { return r->f->cur.stride[uv]; }
void *ref_frame_masks(RefFrame *r) { return r->f->lf.mask; }
void *ref_frame_levels(RefFrame *r) { return r->f->lf.level; }
void *ref_frame_lut(RefFrame *r) { return &r->f->lf.lim_lut; }
void *ref_frame_lr_masks(RefFrame *r) { return r->f->lf.lr_mask; }
int ref_frame_sbh(RefFrame *r) { return r->f->sbh; }
void ref_calc_eih(void *lut, int sharpness) { dav1d_calc_eih((Av1FilterLUT *)lut, sharpness); }
size_t ref_sizeof(int what) {
    switch (what) {
    case 0: return sizeof(Av1Filter);
    case 1: return sizeof(Av1Restoration);
    case 2: return sizeof(Av1FilterLUT);
    case 3: return sizeof(Av1RestorationUnit);
    }
    return 0;
}

/* ---------------------------------------------------------------- threads */
typedef struct Job { void (*fn)(RefFrame *, int tid, int idx, void *arg); RefFrame *r; void *arg; int n, tid; int *next; } Job;
static void *job_main(void *p) {
    Job *j = p;
    /* dynamic schedule (one shared counter), like the reference's task queue */
    for (;;) {
        const int i = __atomic_fetch_add(j->next, 1, __ATOMIC_RELAXED);
        if (i >= j->n) break;
        j->fn(j->r, j->tid, i, j->arg);
    }
    return NULL;
}
static void parallel_for(RefFrame *r, int nthr, int n, void (*fn)(RefFrame *, int, int, void *), void *arg) {
    if (nthr > n) nthr = n;
    if (nthr <= 1) { for (int i = 0; i < n; i++) fn(r, 0, i, arg); return; }
    pthread_t th[256]; Job jobs[256];
    int next = 0;
    if (nthr > 256) nthr = 256;
    for (int t = 0; t < nthr; t++) {
        jobs[t] = (Job){ fn, r, arg, n, t, &next };
        pthread_create(&th[t], NULL, job_main, &jobs[t]);
    }
    for (int t = 0; t < nthr; t++) pthread_join(th[t], NULL);
}

static void st_cols(RefFrame *r, int tid, int sby, void *a) { (void)tid; (void)a; if (r->hbd) dav1d_filter_sbrow_deblock_cols_16bpc(r->f, sby); else dav1d_filter_sbrow_deblock_cols_8bpc(r->f, sby); }
static void st_rows(RefFrame *r, int tid, int sby, void *a) { (void)tid; (void)a; if (r->hbd) dav1d_filter_sbrow_deblock_rows_16bpc(r->f, sby); else dav1d_filter_sbrow_deblock_rows_8bpc(r->f, sby); }
static void st_cdef(RefFrame *r, int tid, int sby, void *a) { (void)a; if (r->hbd) dav1d_filter_sbrow_cdef_16bpc(&r->tc[tid], sby); else dav1d_filter_sbrow_cdef_8bpc(&r->tc[tid], sby); }
static void st_lr(RefFrame *r, int tid, int sby, void *a) { (void)tid; (void)a; if (r->hbd) dav1d_filter_sbrow_lr_16bpc(r->f, sby); else dav1d_filter_sbrow_lr_8bpc(r->f, sby); }

/* stages: RB200_STAGE_DEBLOCK | _CDEF | _LR */
void ref_frame_filter(RefFrame *r, int stages, int n_threads) {
    Dav1dFrameContext *f = r->f;
    r->c->inloop_filters = ((stages & RB200_STAGE_DEBLOCK) ? DAV1D_INLOOPFILTER_DEBLOCK : 0) |
                           ((stages & RB200_STAGE_CDEF) ? DAV1D_INLOOPFILTER_CDEF : 0) |
                           ((stages & RB200_STAGE_LR) ? DAV1D_INLOOPFILTER_RESTORATION : 0);
    r->seq.cdef = !!(stages & RB200_STAGE_CDEF);
    const int restore_planes = f->lf.restore_planes;
    if (!(stages & RB200_STAGE_LR)) f->lf.restore_planes = 0;
    for (int i = 0; i < r->n_tc; i++) r->tc[i].top_pre_cdef_toggle = 0;
    if (r->n_tc == 1) {
        for (int sby = 0; sby < f->sbh; sby++) {
            if (r->hbd) dav1d_filter_sbrow_16bpc(f, sby); else dav1d_filter_sbrow_8bpc(f, sby);
        }
    } else {
        if (n_threads > r->n_tc) n_threads = r->n_tc;
        parallel_for(r, n_threads, f->sbh, st_cols, NULL);
        parallel_for(r, n_threads, f->sbh, st_rows, NULL);   /* + dav1d_copy_lpf */
        if (r->seq.cdef) parallel_for(r, n_threads, f->sbh, st_cdef, NULL);
        if (f->lf.restore_planes) parallel_for(r, n_threads, f->sbh, st_lr, NULL);
    }
    f->lf.restore_planes = restore_planes;
}

/* ------------------------------------------------------------------ recon */
typedef struct ReconArgs {
    const Rb200McItem *mc; const Rb200ItxItem *itx; const void *coef_src; void *coef_work;
    RefFrame *refs[8]; int chunk;
    int n_mc, n_itx;
} ReconArgs;

typedef void (*mc_fn8)(void *, ptrdiff_t, const void *, ptrdiff_t, int, int, int, int);
typedef void (*mc_fn16)(void *, ptrdiff_t, const void *, ptrdiff_t, int, int, int, int, int);
typedef void (*itx_fn8)(void *, ptrdiff_t, void *, int);
typedef void (*itx_fn16)(void *, ptrdiff_t, void *, int, int);

static void do_mc_chunk(RefFrame *r, int tid, int chunk, void *arg) {
    (void)tid;
    ReconArgs *a = arg;
    Dav1dFrameContext *f = r->f;
    const int px = r->hbd ? 2 : 1;
    const int ss_ver_l = f->cur.p.layout == DAV1D_PIXEL_LAYOUT_I420, ss_hor_l = f->cur.p.layout != DAV1D_PIXEL_LAYOUT_I444;
    uint8_t emu[320 * (256 + 7) * 2];
    const int lo = chunk * a->chunk, hi = lo + a->chunk < a->n_mc ? lo + a->chunk : a->n_mc;
    for (int i = lo; i < hi; i++) {
        const Rb200McItem *it = &a->mc[i];
        const int pl = it->plane, ss_hor = pl && ss_hor_l, ss_ver = pl && ss_ver_l;
        const Dav1dFrameContext *rf = a->refs[it->ref]->f;
        const int w = (f->cur.p.w + ss_hor) >> ss_hor, h = (f->cur.p.h + ss_ver) >> ss_ver;
        const int mx = it->mx, my = it->my, dx = it->src_x, dy = it->src_y, bw = it->w, bh = it->h;
        ptrdiff_t ref_stride = rf->cur.stride[!!pl];
        const uint8_t *ref;
        /* src/recon_tmpl.c:1065-1080 */
        if (dx < !!mx * 3 || dy < !!my * 3 || dx + bw + !!mx * 4 > w || dy + bh + !!my * 4 > h) {
            ((void (*)(intptr_t, intptr_t, intptr_t, intptr_t, intptr_t, intptr_t, void *, ptrdiff_t, const void *, ptrdiff_t))
                 f->dsp->mc.emu_edge)(bw + !!mx * 7, bh + !!my * 7, w, h, dx - !!mx * 3, dy - !!my * 3, emu,
                                      192 * px, rf->cur.data[pl], ref_stride);
            ref = emu + (192 * !!my * 3 + !!mx * 3) * px;
            ref_stride = 192 * px;
        } else {
            ref = (const uint8_t *)rf->cur.data[pl] + ref_stride * dy + (ptrdiff_t)dx * px;
        }
        uint8_t *dst = (uint8_t *)f->cur.data[pl] + f->cur.stride[!!pl] * it->dst_y + (ptrdiff_t)it->dst_x * px;
        if (r->hbd) ((mc_fn16)f->dsp->mc.mc[it->filter2d])(dst, f->cur.stride[!!pl], ref, ref_stride, bw, bh, mx, my, r->bdmax);
        else ((mc_fn8)f->dsp->mc.mc[it->filter2d])(dst, f->cur.stride[!!pl], ref, ref_stride, bw, bh, mx, my);
    }
}

static void do_itx_chunk(RefFrame *r, int tid, int chunk, void *arg) {
    (void)tid;
    ReconArgs *a = arg;
    Dav1dFrameContext *f = r->f;
    const int px = r->hbd ? 2 : 1, cs = r->hbd ? 4 : 2;
    const int lo = chunk * a->chunk, hi = lo + a->chunk < a->n_itx ? lo + a->chunk : a->n_itx;
    for (int i = lo; i < hi; i++) {
        const Rb200ItxItem *it = &a->itx[i];
        uint8_t *dst = (uint8_t *)f->cur.data[it->plane] + f->cur.stride[!!it->plane] * it->y + (ptrdiff_t)it->x * px;
        void *cf = (uint8_t *)a->coef_work + (size_t)it->cf_off * cs;
        if (r->hbd) ((itx_fn16)f->dsp->itx.itxfm_add[it->tx][it->txtp])(dst, f->cur.stride[!!it->plane], cf, it->eob, r->bdmax);
        else ((itx_fn8)f->dsp->itx.itxfm_add[it->tx][it->txtp])(dst, f->cur.stride[!!it->plane], cf, it->eob);
    }
}

/* Prediction + residual for one frame.  `coef_work` is consumed (zeroed) like the
 * reference's cf buffer; the caller restores it between timed iterations. */
void ref_frame_recon(RefFrame *r, RefFrame *const refs[], int n_refs, const Rb200McItem *mc, int n_mc,
                     const Rb200ItxItem *itx, int n_itx, void *coef_work, int n_threads) {
    ReconArgs a;
    memset(&a, 0, sizeof(a));
    a.mc = mc; a.itx = itx; a.coef_work = coef_work; a.n_mc = n_mc; a.n_itx = n_itx;
    for (int i = 0; i < n_refs && i < 8; i++) a.refs[i] = refs[i];
    a.chunk = 256;
    parallel_for(r, n_threads, (n_mc + a.chunk - 1) / a.chunk, do_mc_chunk, &a);
    parallel_for(r, n_threads, (n_itx + a.chunk - 1) / a.chunk, do_itx_chunk, &a);
}

/* ------------------------------------------------------------- film grain */
/* dav1d_apply_grain (src/fg_apply_tmpl.c:229-245) from the frame's current picture into a
 * separate output picture, as rav1d_apply_grain does on output (src/lib.c output_image). */
void ref_frame_apply_grain(RefFrame *r, const Dav1dFilmGrainData *data, int is_identity) {
    Dav1dFrameContext *f = r->f;
    if (!r->grain_mem) r->grain_mem = zalloc(r->plane_bytes);
    r->hdr.film_grain.data = *data;
    r->seq.mtrx = is_identity ? DAV1D_MC_IDENTITY : DAV1D_MC_BT709;
    Dav1dPicture in;
    memset(&in, 0, sizeof(in));
    in.data[0] = f->cur.data[0]; in.data[1] = f->cur.data[1]; in.data[2] = f->cur.data[2];
    in.stride[0] = f->cur.stride[0]; in.stride[1] = f->cur.stride[1];
    in.p = f->cur.p; in.seq_hdr = &r->seq; in.frame_hdr = &r->hdr;
    r->grain_out = in;
    for (int i = 0; i < 3; i++)
        r->grain_out.data[i] = f->cur.data[i] ? r->grain_mem + ((uint8_t *)f->cur.data[i] - r->plane_mem) : NULL;
    if (r->hbd) dav1d_apply_grain_16bpc(&f->dsp->fg, &r->grain_out, &in);
    else dav1d_apply_grain_8bpc(&f->dsp->fg, &r->grain_out, &in);
}
void *ref_frame_grain_plane(RefFrame *r, int pl) { return r->grain_out.data[pl]; }

/* ------------------------------------------------------- compound blocks */
/* The compound branch of dav1d_recon_b_inter (src/recon_tmpl.c:1836-1921) for avg / w_avg / seg,
 * with the addressing of mc() (src/recon_tmpl.c:1036-1106). */
typedef void (*mct_fn8)(int16_t *, const void *, ptrdiff_t, int, int, int, int);
typedef void (*mct_fn16)(int16_t *, const void *, ptrdiff_t, int, int, int, int, int);
typedef struct CompArgs { const Rb200CompItem *it; RefFrame *refs[8]; int n, chunk; } CompArgs;

static void comp_prep(RefFrame *r, const RefFrame *rfr, int16_t *tmp, int pl, int px0, int py0, int bw, int bh,
                      int mvx, int mvy, int filter2d, uint8_t *emu) {
    Dav1dFrameContext *f = r->f;
    const Dav1dFrameContext *rf = rfr->f;
    const int px = r->hbd ? 2 : 1;
    const int ss_hor = pl && f->cur.p.layout != DAV1D_PIXEL_LAYOUT_I444, ss_ver = pl && f->cur.p.layout == DAV1D_PIXEL_LAYOUT_I420;
    const int w = (f->cur.p.w + ss_hor) >> ss_hor, h = (f->cur.p.h + ss_ver) >> ss_ver;
    const int mx = (mvx & (15 >> !ss_hor)) << !ss_hor, my = (mvy & (15 >> !ss_ver)) << !ss_ver;
    const int dx = px0 + (mvx >> (3 + ss_hor)), dy = py0 + (mvy >> (3 + ss_ver));
    ptrdiff_t ref_stride = rf->cur.stride[!!pl];
    const uint8_t *ref;
    if (dx < !!mx * 3 || dy < !!my * 3 || dx + bw + !!mx * 4 > w || dy + bh + !!my * 4 > h) {
        ((void (*)(intptr_t, intptr_t, intptr_t, intptr_t, intptr_t, intptr_t, void *, ptrdiff_t, const void *, ptrdiff_t))
             f->dsp->mc.emu_edge)(bw + !!mx * 7, bh + !!my * 7, w, h, dx - !!mx * 3, dy - !!my * 3, emu, 192 * px,
                                  rf->cur.data[pl], ref_stride);
        ref = emu + (192 * !!my * 3 + !!mx * 3) * px;
        ref_stride = 192 * px;
    } else {
        ref = (const uint8_t *)rf->cur.data[pl] + ref_stride * dy + (ptrdiff_t)dx * px;
    }
    if (r->hbd) ((mct_fn16)f->dsp->mc.mct[filter2d])(tmp, ref, ref_stride, bw, bh, mx, my, r->bdmax);
    else ((mct_fn8)f->dsp->mc.mct[filter2d])(tmp, ref, ref_stride, bw, bh, mx, my);
}

/* The prediction of a GLOBALMV_GLOBALMV compound block from a reference whose global motion may be warped: warp_affine
 * with dst16 (src/recon_tmpl.c:1139-1198), i.e. warp8x8t per 8x8 of the block's plane, into tmp (row pitch bw). */
static void comp_warp_prep(RefFrame *r, const RefFrame *rfr, int16_t *tmp, int pl, int bx_luma, int by_luma, int bw, int bh,
                           const Dav1dWarpedMotionParams *wmp, uint8_t *emu) {
    Dav1dFrameContext *f = r->f;
    const Dav1dFrameContext *rf = rfr->f;
    const int px = r->hbd ? 2 : 1;
    const int ss_hor = pl && f->cur.p.layout != DAV1D_PIXEL_LAYOUT_I444, ss_ver = pl && f->cur.p.layout == DAV1D_PIXEL_LAYOUT_I420;
    const int width = (rf->cur.p.w + ss_hor) >> ss_hor, height = (rf->cur.p.h + ss_ver) >> ss_ver;
    const int32_t *mat = wmp->matrix;
    for (int y = 0; y < bh; y += 8) {
        const int src_y = by_luma + ((y + 4) << ss_ver);
        const int64_t mat3_y = (int64_t)mat[3] * src_y + mat[0], mat5_y = (int64_t)mat[5] * src_y + mat[1];
        for (int x = 0; x < bw; x += 8) {
            const int src_x = bx_luma + ((x + 4) << ss_hor);
            const int64_t mvx = ((int64_t)mat[2] * src_x + mat3_y) >> ss_hor, mvy = ((int64_t)mat[4] * src_x + mat5_y) >> ss_ver;
            const int dx = (int)(mvx >> 16) - 4, dy = (int)(mvy >> 16) - 4;
            const int mx = (((int)mvx & 0xffff) - wmp->u.abcd[0] * 4 - wmp->u.abcd[1] * 7) & ~0x3f;
            const int my = (((int)mvy & 0xffff) - wmp->u.abcd[2] * 4 - wmp->u.abcd[3] * 4) & ~0x3f;
            const uint8_t *ref_ptr;
            ptrdiff_t ref_stride = rf->cur.stride[!!pl];
            if (dx < 3 || dx + 8 + 4 > width || dy < 3 || dy + 8 + 4 > height) {
                ((void (*)(intptr_t, intptr_t, intptr_t, intptr_t, intptr_t, intptr_t, void *, ptrdiff_t, const void *, ptrdiff_t))
                     f->dsp->mc.emu_edge)(15, 15, width, height, dx - 3, dy - 3, emu, 32 * px, rf->cur.data[pl], ref_stride);
                ref_ptr = emu + (32 * 3 + 3) * px;
                ref_stride = 32 * px;
            } else {
                ref_ptr = (const uint8_t *)rf->cur.data[pl] + ref_stride * dy + (ptrdiff_t)dx * px;
            }
            int16_t *d = tmp + y * bw + x;
            if (r->hbd) ((void (*)(int16_t *, ptrdiff_t, const void *, ptrdiff_t, const int16_t *, int, int, int))f->dsp->mc.warp8x8t)(d, bw, ref_ptr, ref_stride, wmp->u.abcd, mx, my, r->bdmax);
            else ((void (*)(int16_t *, ptrdiff_t, const void *, ptrdiff_t, const int16_t *, int, int))f->dsp->mc.warp8x8t)(d, bw, ref_ptr, ref_stride, wmp->u.abcd, mx, my);
        }
    }
}

/* The prediction of a compound block from a reference of another size: the scaled branch of mc() with dst16
 * (src/recon_tmpl.c:1010-1073), scale factors as dav1d_submit_frame derives them from the two sizes. */
typedef void (*mcts_fn8)(int16_t *, const void *, ptrdiff_t, int, int, int, int, int, int);
typedef void (*mcts_fn16)(int16_t *, const void *, ptrdiff_t, int, int, int, int, int, int, int);
static int comp_scale_mv(int v, int scale) {
    const int64_t t = (int64_t)v * scale + (int64_t)(scale - 0x4000) * 8;
    const int64_t a = ((t < 0 ? -t : t) + 128) >> 8;
    return (int)(t < 0 ? -a : a) + 32;
}
static void comp_scaled_prep(RefFrame *r, const RefFrame *rfr, int16_t *tmp, int pl, int px0, int py0, int bw, int bh,
                             int mvx, int mvy, int filter2d, uint8_t *emu) {
    Dav1dFrameContext *f = r->f;
    const Dav1dFrameContext *rf = rfr->f;
    const int px = r->hbd ? 2 : 1;
    const int ss_hor = pl && f->cur.p.layout != DAV1D_PIXEL_LAYOUT_I444, ss_ver = pl && f->cur.p.layout == DAV1D_PIXEL_LAYOUT_I420;
    const int scale_x = ((rf->cur.p.w << 14) + (f->cur.p.w >> 1)) / f->cur.p.w, scale_y = ((rf->cur.p.h << 14) + (f->cur.p.h >> 1)) / f->cur.p.h;
    const int step_x = (scale_x + 8) >> 4, step_y = (scale_y + 8) >> 4;
    const int pos_x = comp_scale_mv((px0 << 4) + mvx * (1 << !ss_hor), scale_x), pos_y = comp_scale_mv((py0 << 4) + mvy * (1 << !ss_ver), scale_y);
    const int left = pos_x >> 10, top = pos_y >> 10;
    const int right = ((pos_x + (bw - 1) * step_x) >> 10) + 1, bottom = ((pos_y + (bh - 1) * step_y) >> 10) + 1;
    const int w = (rf->cur.p.w + ss_hor) >> ss_hor, h = (rf->cur.p.h + ss_ver) >> ss_ver;
    ptrdiff_t ref_stride = rf->cur.stride[!!pl];
    const uint8_t *ref;
    if (left < 3 || top < 3 || right + 4 > w || bottom + 4 > h) {
        ((void (*)(intptr_t, intptr_t, intptr_t, intptr_t, intptr_t, intptr_t, void *, ptrdiff_t, const void *, ptrdiff_t))
             f->dsp->mc.emu_edge)(right - left + 7, bottom - top + 7, w, h, left - 3, top - 3, emu, 320 * px, rf->cur.data[pl], ref_stride);
        ref = emu + (320 * 3 + 3) * px;
        ref_stride = 320 * px;
    } else {
        ref = (const uint8_t *)rf->cur.data[pl] + ref_stride * top + (ptrdiff_t)left * px;
    }
    if (r->hbd) ((mcts_fn16)f->dsp->mc.mct_scaled[filter2d])(tmp, ref, ref_stride, bw, bh, pos_x & 0x3ff, pos_y & 0x3ff, step_x, step_y, r->bdmax);
    else ((mcts_fn8)f->dsp->mc.mct_scaled[filter2d])(tmp, ref, ref_stride, bw, bh, pos_x & 0x3ff, pos_y & 0x3ff, step_x, step_y);
}

static int wedge_bs(int w, int h) {
    switch (w << 8 | h) {
    case 32 << 8 | 32: return BS_32x32; case 32 << 8 | 16: return BS_32x16; case 32 << 8 | 8: return BS_32x8;
    case 16 << 8 | 32: return BS_16x32; case 16 << 8 | 16: return BS_16x16; case 16 << 8 | 8: return BS_16x8;
    case 8 << 8 | 32: return BS_8x32; case 8 << 8 | 16: return BS_8x16; default: return BS_8x8;
    }
}

static void do_comp_chunk(RefFrame *r, int tid, int chunk, void *arg) {
    (void)tid;
    CompArgs *a = arg;
    Dav1dFrameContext *f = r->f;
    const int px = r->hbd ? 2 : 1;
    const int layout = f->cur.p.layout;
    const int ss_hor_c = layout != DAV1D_PIXEL_LAYOUT_I444, ss_ver_c = layout == DAV1D_PIXEL_LAYOUT_I420;
    const int n_planes = layout == DAV1D_PIXEL_LAYOUT_I400 ? 1 : 3;
    const int chr_layout_idx = layout == DAV1D_PIXEL_LAYOUT_I400 ? 0 : DAV1D_PIXEL_LAYOUT_I444 - layout;
    uint8_t *emu = malloc(320 * (256 + 7) * 2);
    int16_t (*tmp)[128 * 128] = malloc(2 * sizeof(*tmp));
    uint8_t *seg_mask = malloc(128 * 128);
    const int lo = chunk * a->chunk, hi = lo + a->chunk < a->n ? lo + a->chunk : a->n;
    for (int i = lo; i < hi; i++) {
        const Rb200CompItem *it = &a->it[i];
        for (int pl = 0; pl < n_planes; pl++) {
            const int ss_hor = pl && ss_hor_c, ss_ver = pl && ss_ver_c;
            const int bw = it->w >> ss_hor, bh = it->h >> ss_ver, x0 = it->x >> ss_hor, y0 = it->y >> ss_ver;
            for (int k = 0; k < 2; k++) {
                if ((it->warp_mask >> ((pl ? 2 : 0) + k)) & 1)      /* the reference's global-motion warp (frame_hdr.gmv) */
                    comp_warp_prep(r, a->refs[it->ref[k]], tmp[k], pl, it->x, it->y, bw, bh, &f->frame_hdr->gmv[it->ref[k]], emu);
                else if (a->refs[it->ref[k]]->f->cur.p.w != f->cur.p.w || a->refs[it->ref[k]]->f->cur.p.h != f->cur.p.h)
                    comp_scaled_prep(r, a->refs[it->ref[k]], tmp[k], pl, x0, y0, bw, bh, it->mv[k][1], it->mv[k][0], it->filter2d, emu);
                else
                    comp_prep(r, a->refs[it->ref[k]], tmp[k], pl, x0, y0, bw, bh, it->mv[k][1], it->mv[k][0], it->filter2d, emu);
            }
            uint8_t *dst = (uint8_t *)f->cur.data[pl] + f->cur.stride[!!pl] * y0 + (ptrdiff_t)x0 * px;
            const ptrdiff_t ds = f->cur.stride[!!pl];
            const int s = it->mask_sign;
#define CALL8(fn, ...) ((void (*)())(fn))(__VA_ARGS__)
            if (it->comp_type == RB200_COMP_AVG) {
                if (r->hbd) ((void (*)(void *, ptrdiff_t, const int16_t *, const int16_t *, int, int, int))f->dsp->mc.avg)(dst, ds, tmp[0], tmp[1], bw, bh, r->bdmax);
                else ((void (*)(void *, ptrdiff_t, const int16_t *, const int16_t *, int, int))f->dsp->mc.avg)(dst, ds, tmp[0], tmp[1], bw, bh);
            } else if (it->comp_type == RB200_COMP_WEIGHTED_AVG) {
                if (r->hbd) ((void (*)(void *, ptrdiff_t, const int16_t *, const int16_t *, int, int, int, int))f->dsp->mc.w_avg)(dst, ds, tmp[0], tmp[1], bw, bh, it->jnt_weight, r->bdmax);
                else ((void (*)(void *, ptrdiff_t, const int16_t *, const int16_t *, int, int, int))f->dsp->mc.w_avg)(dst, ds, tmp[0], tmp[1], bw, bh, it->jnt_weight);
            } else if (it->comp_type == RB200_COMP_WEDGE) {
                /* src/recon_tmpl.c:1874-1881,1913-1919 */
                const int bs = wedge_bs(it->w, it->h);
                const uint8_t *mask = pl ? dav1d_wedge_masks[bs][chr_layout_idx][s][it->wedge_idx]
                                         : dav1d_wedge_masks[bs][0][0][it->wedge_idx];
                if (r->hbd) ((void (*)(void *, ptrdiff_t, const int16_t *, const int16_t *, int, int, const uint8_t *, int))f->dsp->mc.mask)(dst, ds, tmp[s], tmp[!s], bw, bh, mask, r->bdmax);
                else ((void (*)(void *, ptrdiff_t, const int16_t *, const int16_t *, int, int, const uint8_t *))f->dsp->mc.mask)(dst, ds, tmp[s], tmp[!s], bw, bh, mask);
            } else if (pl == 0) {
                if (r->hbd) ((void (*)(void *, ptrdiff_t, const int16_t *, const int16_t *, int, int, uint8_t *, int, int))f->dsp->mc.w_mask[chr_layout_idx])(dst, ds, tmp[s], tmp[!s], bw, bh, seg_mask, s, r->bdmax);
                else ((void (*)(void *, ptrdiff_t, const int16_t *, const int16_t *, int, int, uint8_t *, int))f->dsp->mc.w_mask[chr_layout_idx])(dst, ds, tmp[s], tmp[!s], bw, bh, seg_mask, s);
            } else {
                if (r->hbd) ((void (*)(void *, ptrdiff_t, const int16_t *, const int16_t *, int, int, const uint8_t *, int))f->dsp->mc.mask)(dst, ds, tmp[s], tmp[!s], bw, bh, seg_mask, r->bdmax);
                else ((void (*)(void *, ptrdiff_t, const int16_t *, const int16_t *, int, int, const uint8_t *))f->dsp->mc.mask)(dst, ds, tmp[s], tmp[!s], bw, bh, seg_mask);
            }
#undef CALL8
        }
    }
    free(seg_mask); free(tmp); free(emu);
}

/* frame_hdr.gmv[slot] of the harness frame (real decoder frames carry their own) */
void ref_frame_set_gmv(RefFrame *r, int slot, const int32_t matrix[6], const int16_t abcd[4]) {
    memcpy(r->hdr.gmv[slot].matrix, matrix, sizeof(r->hdr.gmv[slot].matrix));
    memcpy(r->hdr.gmv[slot].u.abcd, abcd, sizeof(r->hdr.gmv[slot].u.abcd));
}

void ref_frame_recon_comp(RefFrame *r, RefFrame *const refs[], int n_refs, const Rb200CompItem *items, int n, int n_threads) {
    CompArgs a;
    static int wedge_ready;
    if (!wedge_ready) { dav1d_init_wedge_masks(); wedge_ready = 1; }   /* dav1d_open does this once (src/lib.c:59) */
    memset(&a, 0, sizeof(a));
    a.it = items; a.n = n; a.chunk = 64;
    for (int i = 0; i < n_refs && i < 8; i++) a.refs[i] = refs[i];
    parallel_for(r, n_threads, (n + a.chunk - 1) / a.chunk, do_comp_chunk, &a);
}

/* --------------------------------------------------------- warped blocks */
/* warp_affine (src/recon_tmpl.c:1139-1198) for put, driven by Rb200WarpItem. */
typedef struct WarpArgs { const Rb200WarpItem *it; RefFrame *refs[8]; int n; } WarpArgs;
static void do_warp_item(RefFrame *r, int tid, int i, void *arg) {
    (void)tid;
    WarpArgs *a = arg;
    Dav1dFrameContext *f = r->f;
    const Rb200WarpItem *it = &a->it[i];
    const Dav1dFrameContext *rf = a->refs[it->ref]->f;
    const int px = r->hbd ? 2 : 1;
    const int layout = f->cur.p.layout;
    const int n_planes = layout == DAV1D_PIXEL_LAYOUT_I400 ? 1 : 3;
    uint8_t emu[32 * 32 * 2];
    for (int pl = 0; pl < n_planes; pl++) {
        const int ss_ver = pl && layout == DAV1D_PIXEL_LAYOUT_I420, ss_hor = pl && layout != DAV1D_PIXEL_LAYOUT_I444;
        const int32_t *mat = it->matrix;
        /* chroma is warped only when the chroma block is at least 8x8 (imin(cbw4, cbh4) > 1, src/recon_tmpl.c:1755); smaller
         * chroma blocks of a warped block are ordinary Rb200McItems */
        if (pl && ((it->w >> ss_hor) < 8 || (it->h >> ss_ver) < 8)) continue;
        const int width = (rf->cur.p.w + ss_hor) >> ss_hor, height = (rf->cur.p.h + ss_ver) >> ss_ver;
        uint8_t *dst8 = (uint8_t *)f->cur.data[pl] + f->cur.stride[!!pl] * (it->y >> ss_ver) + (ptrdiff_t)(it->x >> ss_hor) * px;
        for (int y = 0; y < it->h >> ss_ver; y += 8) {
            const int src_y = it->y + ((y + 4) << ss_ver);
            const int64_t mat3_y = (int64_t)mat[3] * src_y + mat[0];
            const int64_t mat5_y = (int64_t)mat[5] * src_y + mat[1];
            for (int x = 0; x < it->w >> ss_hor; x += 8) {
                const int src_x = it->x + ((x + 4) << ss_hor);
                const int64_t mvx = ((int64_t)mat[2] * src_x + mat3_y) >> ss_hor;
                const int64_t mvy = ((int64_t)mat[4] * src_x + mat5_y) >> ss_ver;
                const int dx = (int)(mvx >> 16) - 4;
                const int mx = (((int)mvx & 0xffff) - it->abcd[0] * 4 - it->abcd[1] * 7) & ~0x3f;
                const int dy = (int)(mvy >> 16) - 4;
                const int my = (((int)mvy & 0xffff) - it->abcd[2] * 4 - it->abcd[3] * 4) & ~0x3f;
                const uint8_t *ref_ptr;
                ptrdiff_t ref_stride = rf->cur.stride[!!pl];
                if (dx < 3 || dx + 8 + 4 > width || dy < 3 || dy + 8 + 4 > height) {
                    ((void (*)(intptr_t, intptr_t, intptr_t, intptr_t, intptr_t, intptr_t, void *, ptrdiff_t, const void *, ptrdiff_t))
                         f->dsp->mc.emu_edge)(15, 15, width, height, dx - 3, dy - 3, emu, 32 * px, rf->cur.data[pl], ref_stride);
                    ref_ptr = emu + (32 * 3 + 3) * px;
                    ref_stride = 32 * px;
                } else {
                    ref_ptr = (const uint8_t *)rf->cur.data[pl] + ref_stride * dy + (ptrdiff_t)dx * px;
                }
                uint8_t *d = dst8 + f->cur.stride[!!pl] * y + (ptrdiff_t)x * px;
                if (r->hbd) ((void (*)(void *, ptrdiff_t, const void *, ptrdiff_t, const int16_t *, int, int, int))f->dsp->mc.warp8x8)(d, f->cur.stride[!!pl], ref_ptr, ref_stride, it->abcd, mx, my, r->bdmax);
                else ((void (*)(void *, ptrdiff_t, const void *, ptrdiff_t, const int16_t *, int, int))f->dsp->mc.warp8x8)(d, f->cur.stride[!!pl], ref_ptr, ref_stride, it->abcd, mx, my);
            }
        }
    }
}
void ref_frame_recon_warp(RefFrame *r, RefFrame *const refs[], int n_refs, const Rb200WarpItem *items, int n, int n_threads) {
    WarpArgs a;
    memset(&a, 0, sizeof(a));
    a.it = items; a.n = n;
    for (int i = 0; i < n_refs && i < 8; i++) a.refs[i] = refs[i];
    parallel_for(r, n_threads, n, do_warp_item, &a);
}

/* ------------------------------------------------------------ OBMC strips */
/* obmc() (src/recon_tmpl.c:1076-1137): the neighbour's prediction of a strip into `lap`, then
 * blend_h (above) / blend_v (left).  Items: all ABOVE first, then LEFT; run in list order. */
void ref_frame_recon_obmc(RefFrame *r, RefFrame *const refs[], int n_refs, const Rb200McItem *items, int n) {
    (void)n_refs;
    Dav1dFrameContext *f = r->f;
    const int px = r->hbd ? 2 : 1;
    const int ss_ver_l = f->cur.p.layout == DAV1D_PIXEL_LAYOUT_I420, ss_hor_l = f->cur.p.layout != DAV1D_PIXEL_LAYOUT_I444;
    uint8_t *emu = malloc(320 * (256 + 7) * 2), *lap = malloc(128 * 128 * 2);
    for (int i = 0; i < n; i++) {
        const Rb200McItem *it = &items[i];
        const int pl = it->plane, ss_hor = pl && ss_hor_l, ss_ver = pl && ss_ver_l;
        const Dav1dFrameContext *rf = refs[it->ref]->f;
        const int w = (f->cur.p.w + ss_hor) >> ss_hor, h = (f->cur.p.h + ss_ver) >> ss_ver;
        const int above = it->flags == RB200_MC_OBMC_ABOVE;
        const int v_mul = 4 >> ss_ver;
        const int bw = it->w, bh = above ? (((it->h / v_mul) * 3 + 3) >> 2) * v_mul : it->h;
        const int mx = it->mx, my = it->my, dx = it->src_x, dy = it->src_y;
        ptrdiff_t ref_stride = rf->cur.stride[!!pl];
        const uint8_t *ref;
        if (dx < !!mx * 3 || dy < !!my * 3 || dx + bw + !!mx * 4 > w || dy + bh + !!my * 4 > h) {
            ((void (*)(intptr_t, intptr_t, intptr_t, intptr_t, intptr_t, intptr_t, void *, ptrdiff_t, const void *, ptrdiff_t))
                 f->dsp->mc.emu_edge)(bw + !!mx * 7, bh + !!my * 7, w, h, dx - !!mx * 3, dy - !!my * 3, emu, 192 * px,
                                      rf->cur.data[pl], ref_stride);
            ref = emu + (192 * !!my * 3 + !!mx * 3) * px;
            ref_stride = 192 * px;
        } else {
            ref = (const uint8_t *)rf->cur.data[pl] + ref_stride * dy + (ptrdiff_t)dx * px;
        }
        if (r->hbd) ((mc_fn16)f->dsp->mc.mc[it->filter2d])(lap, (ptrdiff_t)bw * px, ref, ref_stride, bw, bh, mx, my, r->bdmax);
        else ((mc_fn8)f->dsp->mc.mc[it->filter2d])(lap, (ptrdiff_t)bw * px, ref, ref_stride, bw, bh, mx, my);
        uint8_t *dst = (uint8_t *)f->cur.data[pl] + f->cur.stride[!!pl] * it->dst_y + (ptrdiff_t)it->dst_x * px;
        ((void (*)(void *, ptrdiff_t, const void *, int, int))(above ? f->dsp->mc.blend_h : f->dsp->mc.blend_v))(
            dst, f->cur.stride[!!pl], lap, it->w, it->h);
    }
    free(lap); free(emu);
}

/* dav1d_wedge_masks accessor (tests compare the closed-form device masks through frames; this lets a
 * CPU test look at the table itself). layout_idx: 0 = 4:4:4 (luma), 1 = 4:2:2, 2 = 4:2:0. */
const uint8_t *ref_wedge_mask(int w, int h, int layout_idx, int sign, int idx) {
    static int ready;
    if (!ready) { dav1d_init_wedge_masks(); ready = 1; }
    return dav1d_wedge_masks[wedge_bs(w, h)][layout_idx][sign][idx];
}

/* ------------------------------------------------------ scaled references */
/* The scaled branch of mc() (src/recon_tmpl.c:1014-1071) from the point where pos_x / pos_y are known:
 * window bounds, emu_edge with the 320-pixel scratch stride, mc_scaled[filter2d]. */
typedef void (*mcs_fn8)(void *, ptrdiff_t, const void *, ptrdiff_t, int, int, int, int, int, int);
typedef void (*mcs_fn16)(void *, ptrdiff_t, const void *, ptrdiff_t, int, int, int, int, int, int, int);
void ref_frame_recon_scaled(RefFrame *r, RefFrame *const refs[], int n_refs, const Rb200McScaledItem *items, int n) {
    (void)n_refs;
    Dav1dFrameContext *f = r->f;
    const int px = r->hbd ? 2 : 1;
    const int ss_ver_l = f->cur.p.layout == DAV1D_PIXEL_LAYOUT_I420, ss_hor_l = f->cur.p.layout != DAV1D_PIXEL_LAYOUT_I444;
    uint8_t *emu = malloc(320 * (256 + 7) * 2), *lap = malloc(128 * 128 * 2);
    for (int i = 0; i < n; i++) {
        const Rb200McScaledItem *it = &items[i];
        const int pl = it->plane, ss_hor = pl && ss_hor_l, ss_ver = pl && ss_ver_l;
        const Dav1dFrameContext *rf = refs[it->ref]->f;
        /* an OBMC strip: the neighbour's prediction (w x pred_h, src/recon_tmpl.c:1100-1130) into `lap`, then blend_h / blend_v */
        const int strip = it->flags != RB200_MC_PUT, above = it->flags == RB200_MC_OBMC_ABOVE, v_mul = 4 >> ss_ver;
        const int pw = it->w, ph = strip && above ? (((it->h / v_mul) * 3 + 3) >> 2) * v_mul : it->h;
        const int left = it->pos_x >> 10, top = it->pos_y >> 10;
        const int right = ((it->pos_x + (pw - 1) * it->step_x) >> 10) + 1;
        const int bottom = ((it->pos_y + (ph - 1) * it->step_y) >> 10) + 1;
        const int w = (rf->cur.p.w + ss_hor) >> ss_hor, h = (rf->cur.p.h + ss_ver) >> ss_ver;
        ptrdiff_t ref_stride = rf->cur.stride[!!pl];
        const uint8_t *ref;
        if (left < 3 || top < 3 || right + 4 > w || bottom + 4 > h) {
            ((void (*)(intptr_t, intptr_t, intptr_t, intptr_t, intptr_t, intptr_t, void *, ptrdiff_t, const void *, ptrdiff_t))
                 f->dsp->mc.emu_edge)(right - left + 7, bottom - top + 7, w, h, left - 3, top - 3, emu, 320 * px,
                                      rf->cur.data[pl], ref_stride);
            ref = emu + (320 * 3 + 3) * px;
            ref_stride = 320 * px;
        } else {
            ref = (const uint8_t *)rf->cur.data[pl] + ref_stride * top + (ptrdiff_t)left * px;
        }
        uint8_t *dst = (uint8_t *)f->cur.data[pl] + f->cur.stride[!!pl] * it->dst_y + (ptrdiff_t)it->dst_x * px;
        uint8_t *out = strip ? lap : dst;
        const ptrdiff_t out_stride = strip ? (ptrdiff_t)pw * px : f->cur.stride[!!pl];
        if (r->hbd) ((mcs_fn16)f->dsp->mc.mc_scaled[it->filter2d])(out, out_stride, ref, ref_stride, pw, ph,
                                                                  it->pos_x & 0x3ff, it->pos_y & 0x3ff, it->step_x, it->step_y, r->bdmax);
        else ((mcs_fn8)f->dsp->mc.mc_scaled[it->filter2d])(out, out_stride, ref, ref_stride, pw, ph,
                                                           it->pos_x & 0x3ff, it->pos_y & 0x3ff, it->step_x, it->step_y);
        if (strip)
            ((void (*)(void *, ptrdiff_t, const void *, int, int))(above ? f->dsp->mc.blend_h : f->dsp->mc.blend_v))(
                dst, f->cur.stride[!!pl], lap, it->w, it->h);
    }
    free(lap); free(emu);
}

/* ------------------------------------------------------------ intra blocks */
/* The intra-prediction half of dav1d_recon_b_intra (src/recon_tmpl.c:1254-1345): per transform block, in DECODE
 * order, dav1d_prepare_intra_edges, the predictor it selects, then the block's residual.  items[] is in decode
 * order; itx_of[i] is the index of block i's residual in itx[] or -1.  The in-loop filters have not run, so the
 * saved pre-filter superblock edge is the picture itself (prefilter_toplevel_sb_edge = NULL). */
#include "src/intra_edge.h"
enum IntraPredMode dav1d_prepare_intra_edges_8bpc(int x, int have_left, int y, int have_top, int w, int h, enum EdgeFlags edge_flags,
                                                  const uint8_t *dst, ptrdiff_t stride, const uint8_t *prefilter_toplevel_sb_edge,
                                                  enum IntraPredMode mode, int *angle, int tw, int th, int filter_edge,
                                                  uint8_t *topleft_out);
enum IntraPredMode dav1d_prepare_intra_edges_16bpc(int x, int have_left, int y, int have_top, int w, int h, enum EdgeFlags edge_flags,
                                                   const uint16_t *dst, ptrdiff_t stride, const uint16_t *prefilter_toplevel_sb_edge,
                                                   enum IntraPredMode mode, int *angle, int tw, int th, int filter_edge,
                                                   uint16_t *topleft_out, int bitdepth_max);
void ref_frame_recon_intra(RefFrame *r, const Rb200IntraItem *items, int n, const int32_t *itx_of, const Rb200ItxItem *itx,
                           void *coef_work, const uint8_t *pal_buf) {
    Dav1dFrameContext *f = r->f;
    const int px = r->hbd ? 2 : 1, cs = r->hbd ? 4 : 2;
    const int ss_ver_l = f->cur.p.layout == DAV1D_PIXEL_LAYOUT_I420, ss_hor_l = f->cur.p.layout != DAV1D_PIXEL_LAYOUT_I444;
    uint16_t edge_buf[257 + 32];
    for (int i = 0; i < n; i++) {
        const Rb200IntraItem *it = &items[i];
        const int pl = it->plane, ss_hor = pl && ss_hor_l, ss_ver = pl && ss_ver_l;
        const ptrdiff_t stride = f->cur.stride[!!pl];
        uint8_t *dst = (uint8_t *)f->cur.data[pl] + stride * (it->y4 * 4) + (ptrdiff_t)(it->x4 * 4) * px;
        const int have_left = it->flags & 1, have_top = (it->flags >> 1) & 1;
        const int ef = (it->flags & 4 ? EDGE_I444_TOP_HAS_RIGHT : 0) | (it->flags & 8 ? EDGE_I444_LEFT_HAS_BOTTOM : 0);
        const int is_sm = (it->flags >> 4) & 1, eief = (it->flags >> 5) & 1;
        int angle = it->angle;
        const int max_w = ((4 * f->bw) >> ss_hor) - 4 * it->x4, max_h = ((4 * f->bh) >> ss_ver) - 4 * it->y4;
        const int intra_flags = (is_sm << 9) | (eief << 10);
        const int w4_end = it->w4_end & 0x1fff, h4_end = it->h4_end & 0x1fff;
        if (it->mode >= 14) {        /* palette block (src/recon_tmpl.c:1231-1252,1429-1462) / residual-only transform block */
            if (it->mode == 14) {
                const uint8_t *rec = pal_buf + ((size_t)it->w4_end | (size_t)it->h4_end << 16) * 16;
                ((void (*)(void *, ptrdiff_t, const void *, const uint8_t *, int, int))f->dsp->ipred.pal_pred)(dst, stride, rec, rec + 16,
                                                                                                         it->tw4 * 4, it->th4 * 4);
            } else if (it->mode == 16) {   /* intra block copy: mc() from the current picture, src/recon_tmpl.c:1631-1645,979-1009 */
                static uint8_t emu[192 * (128 + 7) * 2];
                const int dx = (int16_t)it->w4_end, dy = (int16_t)it->h4_end, mx = (uint8_t)it->angle & 15, my = (uint8_t)it->angle >> 4;
                const int bw = it->tw4 * 4, bh = it->th4 * 4, w = f->bw * 4 >> ss_hor, h = f->bh * 4 >> ss_ver;
                const uint8_t *ref;
                ptrdiff_t ref_stride = stride;
                if (dx < !!mx * 3 || dy < !!my * 3 || dx + bw + !!mx * 4 > w || dy + bh + !!my * 4 > h) {
                    ((void (*)(intptr_t, intptr_t, intptr_t, intptr_t, intptr_t, intptr_t, void *, ptrdiff_t, const void *, ptrdiff_t))
                         f->dsp->mc.emu_edge)(bw + !!mx * 7, bh + !!my * 7, w, h, dx - !!mx * 3, dy - !!my * 3, emu,
                                              192 * px, f->cur.data[pl], stride);
                    ref = emu + (192 * !!my * 3 + !!mx * 3) * px;
                    ref_stride = 192 * px;
                } else {
                    ref = (const uint8_t *)f->cur.data[pl] + stride * dy + (ptrdiff_t)dx * px;
                }
                if (r->hbd) ((mc_fn16)f->dsp->mc.mc[FILTER_2D_BILINEAR])(dst, stride, ref, ref_stride, bw, bh, mx, my, r->bdmax);
                else ((mc_fn8)f->dsp->mc.mc[FILTER_2D_BILINEAR])(dst, stride, ref, ref_stride, bw, bh, mx, my);
            }
        } else if (pl && it->mode == 13) {   /* chroma from luma, src/recon_tmpl.c:1376-1422 */
            int16_t ac[32 * 32];
            const uint8_t *ysrc = (const uint8_t *)f->cur.data[0] + f->cur.stride[0] * ((it->y4 << ss_ver) * 4) +
                                  (ptrdiff_t)((it->x4 << ss_hor) * 4) * px;
            ((void (*)(int16_t *, const void *, ptrdiff_t, int, int, int, int))f->dsp->ipred.cfl_ac[f->cur.p.layout - 1])(
                ac, ysrc, f->cur.stride[0], it->w4_end >> 13, it->h4_end >> 13, it->tw4 * 4, it->th4 * 4);
            int a0 = 0;
            if (r->hbd) {
                uint16_t *edge = edge_buf + 128 + 16;
                const int m = dav1d_prepare_intra_edges_16bpc(it->x4, have_left, it->y4, have_top, w4_end, h4_end, 0, (const uint16_t *)dst, stride,
                                                              NULL, DC_PRED, &a0, it->tw4, it->th4, 0, edge, r->bdmax);
                ((void (*)(void *, ptrdiff_t, const void *, int, int, const int16_t *, int, int))f->dsp->ipred.cfl_pred[m])(
                    dst, stride, edge, it->tw4 * 4, it->th4 * 4, ac, it->angle, r->bdmax);
            } else {
                uint8_t *edge = (uint8_t *)edge_buf + 128 + 16;
                const int m = dav1d_prepare_intra_edges_8bpc(it->x4, have_left, it->y4, have_top, w4_end, h4_end, 0, dst, stride, NULL, DC_PRED,
                                                             &a0, it->tw4, it->th4, 0, edge);
                ((void (*)(void *, ptrdiff_t, const void *, int, int, const int16_t *, int))f->dsp->ipred.cfl_pred[m])(
                    dst, stride, edge, it->tw4 * 4, it->th4 * 4, ac, it->angle);
            }
        } else if (it->flags & 64) {   /* inter-intra, src/recon_tmpl.c:1665-1692,1795-1831 */
            static int ii_ready;
            if (!ii_ready) { dav1d_init_interintra_masks(); dav1d_init_wedge_masks(); ii_ready = 1; }
            uint16_t tmp[32 * 32];
            const int bw = it->tw4 * 4, bh = it->th4 * 4;
            const int bs = wedge_bs(bw << ss_hor, bh << ss_ver);
            const int li = pl ? DAV1D_PIXEL_LAYOUT_I444 - f->cur.p.layout : 0;
            const int ii_mode = it->mode == SMOOTH_PRED ? II_SMOOTH_PRED : it->mode;
            const uint8_t *mask = it->angle < 0 ? dav1d_ii_masks[bs][li][ii_mode] : dav1d_wedge_masks[bs][li][0][it->angle & 15];
            int a0 = 0;
            if (r->hbd) {
                uint16_t *edge = edge_buf + 128 + 16;
                const int m = dav1d_prepare_intra_edges_16bpc(it->x4, have_left, it->y4, have_top, w4_end, h4_end, 0, (const uint16_t *)dst, stride,
                                                              NULL, it->mode, &a0, it->tw4, it->th4, 0, edge, r->bdmax);
                ((void (*)(void *, ptrdiff_t, const void *, int, int, int, int, int, int))f->dsp->ipred.intra_pred[m])(tmp, bw * 2, edge, bw, bh, 0, 0, 0, r->bdmax);
            } else {
                uint8_t *edge = (uint8_t *)edge_buf + 128 + 16;
                const int m = dav1d_prepare_intra_edges_8bpc(it->x4, have_left, it->y4, have_top, w4_end, h4_end, 0, dst, stride, NULL, it->mode,
                                                             &a0, it->tw4, it->th4, 0, edge);
                ((void (*)(void *, ptrdiff_t, const void *, int, int, int, int, int))f->dsp->ipred.intra_pred[m])(tmp, bw, edge, bw, bh, 0, 0, 0);
            }
            ((void (*)(void *, ptrdiff_t, const void *, int, int, const uint8_t *))f->dsp->mc.blend)(dst, stride, tmp, bw, bh, mask);
        } else if (r->hbd) {
            uint16_t *edge = edge_buf + 128 + 16;
            const int m = dav1d_prepare_intra_edges_16bpc(it->x4, have_left, it->y4, have_top, w4_end, h4_end, ef, (const uint16_t *)dst,
                                                          stride, NULL, it->mode, &angle, it->tw4, it->th4, eief, edge, r->bdmax);
            ((void (*)(void *, ptrdiff_t, const void *, int, int, int, int, int, int))f->dsp->ipred.intra_pred[m])(
                dst, stride, edge, it->tw4 * 4, it->th4 * 4, angle | intra_flags, max_w, max_h, r->bdmax);
        } else {
            uint8_t *edge = (uint8_t *)edge_buf + 128 + 16;
            const int m = dav1d_prepare_intra_edges_8bpc(it->x4, have_left, it->y4, have_top, w4_end, h4_end, ef, dst, stride, NULL,
                                                         it->mode, &angle, it->tw4, it->th4, eief, edge);
            ((void (*)(void *, ptrdiff_t, const void *, int, int, int, int, int))f->dsp->ipred.intra_pred[m])(
                dst, stride, edge, it->tw4 * 4, it->th4 * 4, angle | intra_flags, max_w, max_h);
        }
        if (itx_of && itx_of[i] >= 0) {
            const Rb200ItxItem *t = &itx[itx_of[i]];
            uint8_t *d = (uint8_t *)f->cur.data[t->plane] + f->cur.stride[!!t->plane] * t->y + (ptrdiff_t)t->x * px;
            void *cf = (uint8_t *)coef_work + (size_t)t->cf_off * cs;
            if (r->hbd) ((itx_fn16)f->dsp->itx.itxfm_add[t->tx][t->txtp])(d, f->cur.stride[!!t->plane], cf, t->eob, r->bdmax);
            else ((itx_fn8)f->dsp->itx.itxfm_add[t->tx][t->txtp])(d, f->cur.stride[!!t->plane], cf, t->eob);
        }
    }
}
