/*
 * rav1d_b200 host layer: the frame life cycle around the batch builder (recon_batch_tmpl.c).
 *
 * What changes in the reference when reconstruction and the post-filters run on the GPU, and where:
 *
 *   rav1d_open                (src/lib.rs:  n_tc / n_fc, allocator)        two-pass frame threading is required (the batch
 *                                                                          IS pass 2), pictures come from our allocator pair
 *   Rav1dPicAllocator pair    (include/dav1d/picture.rs:300-322)           host planes from the backend (pinned) + a handle
 *                                                                          that owns the picture's device planes
 *   rav1d_decode_frame_init   (src/decode.rs:4069)                         start the frame's batch
 *   rav1d_create_lf_mask_*    (src/lf_mask.rs:380-606, called in pass 1)   + one Rb200LfBlock record per coded block
 *   rav1d_decode_frame_exit   (src/decode.rs:4602)                         hand the batch to the backend, in decode order
 *   rav1d_get_picture         (src/lib.rs:449)                             pinned-host read-back of the picture being output
 *   rav1d_task_delayed_fg     (src/thread_task.rs:  film grain on output)  the grained picture is read back instead
 *
 * Each of these is the reference's own function, interposed: the executable that links this file exports the same
 * symbol, the reference library (position-independent, not -Bsymbolic) resolves its calls here, and the original is
 * reached with dlsym(RTLD_NEXT).  The Rust edit each one stands for is listed in INTEGRATION.md.
 */
#define _GNU_SOURCE
#include "config.h"

#include <dlfcn.h>
#include <errno.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "dav1d/dav1d.h"
#include "src/internal.h"
#include "src/lf_mask.h"
#include "src/picture.h"

#include "host_frame.h"

#define EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------ pictures */
typedef struct RbPic {
    void *backend;          /* backend handle (device planes) */
    void *host_mem;
    int pending;            /* a frame is being decoded into it and has not been handed over yet */
    int have_frame;         /* was handed over: its pixels live in the backend */
    int fetched;            /* host planes are valid */
    int failed;
} RbPic;

typedef struct RbJob {
    pthread_mutex_t lock;
    RbHostBatch batch;
    int active, refs_ready;
} RbJob;

static const RbHostBackend *g_be;
static Dav1dContext *g_c;
static RbJob *g_jobs;
static int g_n_jobs;
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_cond = PTHREAD_COND_INITIALIZER;
static int g_unsupported, g_keep_going;
static char g_why[128];
static long g_frames, g_items[8];

static void die(const char *what) {
    fprintf(stderr, "rav1d_b200 host: %s%s%s\n", what, g_be && g_be->last_error ? ": " : "",
            g_be && g_be->last_error ? g_be->last_error() : "");
    exit(4);
}

static void *real_sym(const char *name) {
    void *p = dlsym(RTLD_NEXT, name);
    if (!p) { fprintf(stderr, "rav1d_b200 host: the reference does not export %s\n", name); exit(4); }
    return p;
}

/* Dav1dPicAllocator.alloc_picture_callback: the layout of dav1d_default_picture_alloc (src/picture.c:46-88; 128-aligned
 * size, anti-aliasing pad of the stride), the memory from the backend, and the handle in allocator_data. */
static int pic_alloc(Dav1dPicture *const p, void *const cookie) {
    (void)cookie;
    const int hbd = p->p.bpc > 8;
    const int aligned_w = (p->p.w + 127) & ~127;
    const int aligned_h = (p->p.h + 127) & ~127;
    const int has_chroma = p->p.layout != DAV1D_PIXEL_LAYOUT_I400;
    const int ss_ver = p->p.layout == DAV1D_PIXEL_LAYOUT_I420;
    const int ss_hor = p->p.layout != DAV1D_PIXEL_LAYOUT_I444;
    ptrdiff_t y_stride = aligned_w << hbd;
    ptrdiff_t uv_stride = has_chroma ? y_stride >> ss_hor : 0;
    if (!(y_stride & 1023)) y_stride += DAV1D_PICTURE_ALIGNMENT;
    if (!(uv_stride & 1023) && has_chroma) uv_stride += DAV1D_PICTURE_ALIGNMENT;
    p->stride[0] = y_stride;
    p->stride[1] = uv_stride;
    const size_t y_sz = y_stride * aligned_h;
    const size_t uv_sz = uv_stride * (aligned_h >> ss_ver);
    RbPic *const rp = calloc(1, sizeof(*rp));
    if (!rp) return DAV1D_ERR(ENOMEM);
    rp->host_mem = g_be->host_alloc(y_sz + 2 * uv_sz + DAV1D_PICTURE_ALIGNMENT);
    rp->backend = g_be->pic_new();
    if (!rp->host_mem || !rp->backend) { free(rp); return DAV1D_ERR(ENOMEM); }
    rp->fetched = g_be->host_pixels;
    uint8_t *const data = rp->host_mem;
    p->data[0] = data;
    p->data[1] = has_chroma ? data + y_sz : NULL;
    p->data[2] = has_chroma ? data + y_sz + uv_sz : NULL;
    p->allocator_data = rp;
    return 0;
}

static void pic_release(Dav1dPicture *const p, void *const cookie) {
    (void)cookie;
    RbPic *const rp = p->allocator_data;
    if (!rp) return;
    g_be->pic_free(rp->backend);
    g_be->host_free(rp->host_mem);
    free(rp);
}

/* ------------------------------------------------------------------------------------------- open / close */
EXPORT int dav1d_open(Dav1dContext **const c_out, const Dav1dSettings *const s) {
    static int (*real)(Dav1dContext **, const Dav1dSettings *);
    if (!real) real = real_sym("dav1d_open");
    if (g_c) { fprintf(stderr, "rav1d_b200 host: one decoder instance per process\n"); return DAV1D_ERR(EINVAL); }
    if (!g_be) {
        g_be = rb200_host_backend();
        if (g_be->init()) die("backend init failed");
    }
    g_unsupported = 0; g_why[0] = 0;
    Dav1dSettings s2 = *s;
    /* the batch replaces pass 2 of the two-pass frame threading: at least two frame contexts */
    if (s2.n_threads < 2) s2.n_threads = 2;
    if (s2.max_frame_delay == 1 || !s2.max_frame_delay) s2.max_frame_delay = 2;
    s2.allocator.cookie = NULL;
    s2.allocator.alloc_picture_callback = pic_alloc;
    s2.allocator.release_picture_callback = pic_release;
    const int r = real(c_out, &s2);
    if (r) return r;
    g_c = *c_out;
    g_n_jobs = g_c->n_fc;
    g_jobs = calloc(g_n_jobs, sizeof(*g_jobs));
    for (int i = 0; i < g_n_jobs; i++) pthread_mutex_init(&g_jobs[i].lock, NULL);
    return 0;
}

EXPORT void dav1d_close(Dav1dContext **const c_out) {
    static void (*real)(Dav1dContext **);
    if (!real) real = real_sym("dav1d_close");
    real(c_out);
    for (int i = 0; i < g_n_jobs; i++) rb_batch_free(&g_jobs[i].batch);
    free(g_jobs);
    g_jobs = NULL; g_n_jobs = 0; g_c = NULL;
    if (getenv("RB200_HOST_STATS"))
        fprintf(stderr, "rav1d_b200 host [%s]: %ld frames; items: mc %ld scaled %ld comp %ld warp %ld obmc %ld itx %ld intra %ld lf blocks %ld\n",
                g_be->name, g_frames, g_items[0], g_items[1], g_items[2], g_items[3], g_items[4], g_items[5], g_items[6], g_items[7]);
    if (g_unsupported && !g_keep_going) {
        /* the run cannot claim parity: some block needed a record the batch formats do not have */
        fprintf(stderr, "rav1d_b200 host: UNSUPPORTED: %s\n", g_why);
        fflush(NULL);
        _exit(3);
    }
}

/* For a host that decodes several streams in one process (conformance_main.c): ask instead of exiting. */
int rb200_host_unsupported(const char **const why) {
    g_keep_going = 1;
    if (why) *why = g_why;
    return g_unsupported;
}

/* ------------------------------------------------------------------------------------------ frame set-up */
static RbJob *job_of(const Dav1dFrameContext *const f) { return &g_jobs[f - f->c->fc]; }

/* dav1d_thread_picture_alloc (src/picture.c:192-233) allocates the frame's output picture in submit_frame, before any
 * task of the frame exists: from here on the picture is "pending" until its frame is handed to the backend. */
EXPORT int dav1d_thread_picture_alloc(Dav1dContext *const c, Dav1dFrameContext *const f, const int bpc) {
    static int (*real)(Dav1dContext *, Dav1dFrameContext *, int);
    if (!real) real = real_sym("dav1d_thread_picture_alloc");
    const int r = real(c, f, bpc);
    if (!r && f->sr_cur.p.allocator_data) {
        pthread_mutex_lock(&g_lock);
        ((RbPic *)f->sr_cur.p.allocator_data)->pending = 1;
        pthread_mutex_unlock(&g_lock);
    }
    return r;
}

EXPORT int dav1d_decode_frame_init(Dav1dFrameContext *const f) {
    static int (*real)(Dav1dFrameContext *);
    if (!real) real = real_sym("dav1d_decode_frame_init");
    const int r = real(f);
    RbJob *const j = job_of(f);
    pthread_mutex_lock(&j->lock);
    rb_batch_reset(&j->batch);
    j->active = !r;
    j->refs_ready = 0;
    pthread_mutex_unlock(&j->lock);
    return r;
}

RbHostBatch *rb_host_block_begin(const Dav1dFrameContext *const f) {
    RbJob *const j = job_of(f);
    pthread_mutex_lock(&j->lock);
    if (!j->refs_ready) {
        /* decode order of the hand-over: no block of this frame is appended before the frames it predicts from have
         * been handed over (their task threads never wait for this frame, so this cannot deadlock) */
        pthread_mutex_lock(&g_lock);
        for (int i = 0; i < 7; i++) {
            const RbPic *const rp = f->refp[i].p.data[0] ? f->refp[i].p.allocator_data : NULL;
            while (rp && rp->pending) pthread_cond_wait(&g_cond, &g_lock);
        }
        pthread_mutex_unlock(&g_lock);
        j->refs_ready = 1;
    }
    return &j->batch;
}

void rb_host_block_end(const Dav1dFrameContext *const f) { pthread_mutex_unlock(&job_of(f)->lock); }

void rb_host_push_palette(RbHostBatch *const b, const int x4, const int y4, const int w4_end, const int h4_end, const int pl,
                          const int bw4, const int bh4, const void *const pal, const int pal_bytes, const uint8_t *const idx)
{
    (void)w4_end; (void)h4_end;
    const int n_idx = bw4 * bh4 * 16, rec = 16 + ((n_idx + 15) & ~15);
    const int off = b->pal.n;
    if (b->pal.n + rec > b->pal.cap) b->pal.v = rb_vec_grow(b->pal.v, &b->pal.cap, b->pal.n + rec, 1);
    b->pal.n += rec;
    memset(b->pal.v + off, 0, rec);
    memcpy(b->pal.v + off, pal, pal_bytes);
    memcpy(b->pal.v + off + 16, idx, n_idx);
    Rb200IntraItem *const it = RB_PUSH(b->intra);
    memset(it, 0, sizeof(*it));
    it->x4 = (uint16_t)x4; it->y4 = (uint16_t)y4;
    it->w4_end = (uint16_t)((off >> 4) & 0xffff); it->h4_end = (uint16_t)((off >> 4) >> 16);
    it->plane = (uint8_t)pl; it->tw4 = (uint8_t)bw4; it->th4 = (uint8_t)bh4; it->mode = 14;
    *RB_PUSH(b->intra_itx) = -1;
}

/* ------------------------------------------------------------------ loop-filter block records (pass 1) */
static RbJob *job_of_level_cache(uint8_t (*const level_cache)[4], const Dav1dFrameContext **const f_out) {
    for (int i = 0; i < g_n_jobs; i++)
        if (g_c->fc[i].lf.level == level_cache) { *f_out = &g_c->fc[i]; return &g_jobs[i]; }
    return NULL;
}

static void lfb_push(uint8_t (*const level_cache)[4], const int bx, const int by, const int bs, int flags, const int ytx, const int uvtx,
                     const uint16_t *const tx_masks, const uint8_t (*const filter_level)[8][2], const int intra)
{
    const Dav1dFrameContext *f;
    RbJob *const j = job_of_level_cache(level_cache, &f);
    if (!j) return;
    /* the intra call has no skip argument; the block itself (pass 1 keeps it in f.frame_thread.b) does */
    if (intra && f->frame_thread.b && f->frame_thread.b[by * f->b4_stride + bx].skip) flags |= RB200_LFB_SKIP;
    pthread_mutex_lock(&j->lock);
    Rb200LfBlock *const b = RB_PUSH(j->batch.lfb);
    b->bx = (uint16_t)bx; b->by = (uint16_t)by; b->bs = (uint8_t)bs; b->flags = (uint8_t)flags;
    b->ytx = (uint8_t)ytx; b->uvtx = (uint8_t)uvtx;
    b->tx_split[0] = tx_masks ? tx_masks[0] : 0; b->tx_split[1] = tx_masks ? tx_masks[1] : 0;
    for (int k = 0; k < 4; k++) b->lvl[k] = filter_level[k][0][0];
    pthread_mutex_unlock(&j->lock);
}

EXPORT void dav1d_create_lf_mask_intra(Av1Filter *lflvl, uint8_t (*level_cache)[4], ptrdiff_t b4_stride,
        const uint8_t (*filter_level)[8][2], int bx, int by, int iw, int ih, enum BlockSize bs, enum RectTxfmSize ytx,
        enum RectTxfmSize uvtx, enum Dav1dPixelLayout layout, uint8_t *ay, uint8_t *ly, uint8_t *auv, uint8_t *luv) {
    static void (*real)(Av1Filter *, uint8_t (*)[4], ptrdiff_t, const uint8_t (*)[8][2], int, int, int, int, enum BlockSize,
                        enum RectTxfmSize, enum RectTxfmSize, enum Dav1dPixelLayout, uint8_t *, uint8_t *, uint8_t *, uint8_t *);
    if (!real) real = real_sym("dav1d_create_lf_mask_intra");
    real(lflvl, level_cache, b4_stride, filter_level, bx, by, iw, ih, bs, ytx, uvtx, layout, ay, ly, auv, luv);
    lfb_push(level_cache, bx, by, bs, RB200_LFB_INTRA | (auv ? RB200_LFB_HAS_CHROMA : 0), ytx, uvtx, NULL, filter_level, 1);
}

EXPORT void dav1d_create_lf_mask_inter(Av1Filter *lflvl, uint8_t (*level_cache)[4], ptrdiff_t b4_stride,
        const uint8_t (*filter_level)[8][2], int bx, int by, int iw, int ih, int skip, enum BlockSize bs,
        enum RectTxfmSize max_ytx, const uint16_t *tx_masks, enum RectTxfmSize uvtx, enum Dav1dPixelLayout layout,
        uint8_t *ay, uint8_t *ly, uint8_t *auv, uint8_t *luv) {
    static void (*real)(Av1Filter *, uint8_t (*)[4], ptrdiff_t, const uint8_t (*)[8][2], int, int, int, int, int, enum BlockSize,
                        enum RectTxfmSize, const uint16_t *, enum RectTxfmSize, enum Dav1dPixelLayout, uint8_t *, uint8_t *,
                        uint8_t *, uint8_t *);
    if (!real) real = real_sym("dav1d_create_lf_mask_inter");
    real(lflvl, level_cache, b4_stride, filter_level, bx, by, iw, ih, skip, bs, max_ytx, tx_masks, uvtx, layout, ay, ly, auv, luv);
    lfb_push(level_cache, bx, by, bs, (skip ? RB200_LFB_SKIP : 0) | (auv ? RB200_LFB_HAS_CHROMA : 0), max_ytx, uvtx, tx_masks,
             filter_level, 0);
}

/* ------------------------------------------------------------------------------------------ frame hand-over */
static int has_grain(const Dav1dFrameHeader *const h) {
    const Dav1dFilmGrainData *const fg = &h->film_grain.data;
    return fg->num_y_points || fg->num_uv_points[0] || fg->num_uv_points[1] ||
           (fg->clip_to_restricted_range && fg->chroma_scaling_from_luma);
}

/* Debug aid (RB200_HOST_DUMP="plane,x,y"): every record of the frame that covers that pixel. */
static void dump_records(const Dav1dFrameContext *const f, const RbHostBatch *const B, const RbHostFinal *const fin) {
    int pl, x, y;
    if (sscanf(getenv("RB200_HOST_DUMP"), "%d,%d,%d", &pl, &x, &y) != 3) return;
    const int ss_hor = pl && f->cur.p.layout != DAV1D_PIXEL_LAYOUT_I444, ss_ver = pl && f->cur.p.layout == DAV1D_PIXEL_LAYOUT_I420;
    fprintf(stderr, "== frame %ld (type %d, show %d, offset %d, %dx%d) plane %d pixel %d,%d\n", g_frames, f->frame_hdr->frame_type,
            f->frame_hdr->show_frame, f->frame_hdr->frame_offset, f->cur.p.w, f->cur.p.h, pl, x, y);
#define IN(x0, y0, w, h) (x >= (x0) && x < (x0) + (w) && y >= (y0) && y < (y0) + (h))
    for (int i = 0; i < B->mc.n; i++) { const Rb200McItem *it = &B->mc.v[i];
        if (it->plane == pl && IN(it->dst_x, it->dst_y, it->w, it->h))
            fprintf(stderr, "  mc[%d] dst %d,%d src %d,%d %dx%d ref %d mx %d my %d f2d %d\n", i, it->dst_x, it->dst_y, it->src_x, it->src_y, it->w, it->h, it->ref, it->mx, it->my, it->filter2d); }
    for (int i = 0; i < B->scaled.n; i++) { const Rb200McScaledItem *it = &B->scaled.v[i];
        if (it->plane == pl && IN(it->dst_x, it->dst_y, it->w, it->h))
            fprintf(stderr, "  scaled[%d] dst %d,%d %dx%d ref %d pos %d,%d step %d,%d f2d %d\n", i, it->dst_x, it->dst_y, it->w, it->h, it->ref, it->pos_x, it->pos_y, it->step_x, it->step_y, it->filter2d); }
    for (int i = 0; i < B->comp.n; i++) { const Rb200CompItem *it = &B->comp.v[i];
        if (IN(it->x >> ss_hor, it->y >> ss_ver, it->w >> ss_hor, it->h >> ss_ver))
            fprintf(stderr, "  comp[%d] %d,%d %dx%d refs %d,%d mv (%d,%d) (%d,%d) f2d %d type %d w %d sign %d wedge %d\n", i, it->x, it->y, it->w, it->h, it->ref[0], it->ref[1],
                    it->mv[0][0], it->mv[0][1], it->mv[1][0], it->mv[1][1], it->filter2d, it->comp_type, it->jnt_weight, it->mask_sign, it->wedge_idx); }
    for (int i = 0; i < B->warp.n; i++) { const Rb200WarpItem *it = &B->warp.v[i];
        if (IN(it->x >> ss_hor, it->y >> ss_ver, it->w >> ss_hor, it->h >> ss_ver))
            fprintf(stderr, "  warp[%d] %d,%d %dx%d ref %d\n", i, it->x, it->y, it->w, it->h, it->ref); }
    for (int k = 0; k < 2; k++) for (int i = 0; i < (k ? B->obmc_left.n : B->obmc_above.n); i++) { const Rb200McItem *it = k ? &B->obmc_left.v[i] : &B->obmc_above.v[i];
        if (it->plane == pl && IN(it->dst_x, it->dst_y, it->w, it->h))
            fprintf(stderr, "  obmc_%s[%d] dst %d,%d src %d,%d %dx%d ref %d mx %d my %d f2d %d\n", k ? "left" : "above", i, it->dst_x, it->dst_y, it->src_x, it->src_y, it->w, it->h, it->ref, it->mx, it->my, it->filter2d); }
    static const uint8_t txw[19] = {4,8,16,32,64,4,8,8,16,16,32,32,64,4,16,8,32,16,64}, txh[19] = {4,8,16,32,64,8,4,16,8,32,16,64,32,16,4,32,8,64,16};
    for (int i = 0; i < fin->n_itx; i++) { const Rb200ItxItem *it = &fin->itx[i];
        if (it->plane == pl && IN(it->x, it->y, txw[it->tx], txh[it->tx]))
            fprintf(stderr, "  itx[%d%s] %d,%d tx %d txtp %d eob %d cf_off %u\n", i, i >= fin->n_itx_inter ? " intra" : "", it->x, it->y, it->tx, it->txtp, it->eob, it->cf_off); }
    for (int i = 0; i < fin->n_intra; i++) { const Rb200IntraItem *it = &fin->intra[i];
        if (it->plane == pl && IN(it->x4 * 4, it->y4 * 4, it->tw4 * 4, it->th4 * 4))
            fprintf(stderr, "  intra[%d] %d,%d %dx%d mode %d angle %d flags 0x%x level %d itx %d ends %d,%d\n", i, it->x4 * 4, it->y4 * 4, it->tw4 * 4, it->th4 * 4, it->mode, it->angle, it->flags, it->level, fin->intra_itx[i], it->w4_end, it->h4_end); }
#undef IN
}

static void hand_over(Dav1dFrameContext *const f, RbJob *const j, RbPic *const rp) {
    const Dav1dContext *const c = f->c;
    const Dav1dFrameHeader *const h = f->frame_hdr;
    RbHostBatch *const B = &j->batch;
    const int hbd = f->cur.p.bpc > 8;
    const int ss_ver = f->cur.p.layout == DAV1D_PIXEL_LAYOUT_I420, ss_hor = f->cur.p.layout != DAV1D_PIXEL_LAYOUT_I444;

    /* coefficients used: the end of pass 1's cursor; pass 2 must have walked exactly the same distance */
    size_t n_coefs = 0;
    for (int i = 0; i < f->n_ts; i++) {
        const uint8_t *const base = (const uint8_t *)f->frame_thread.cf;
        const size_t e1 = ((const uint8_t *)f->ts[i].frame_thread[1].cf - base) >> (1 + hbd);
        const size_t e2 = ((const uint8_t *)f->ts[i].frame_thread[0].cf - base) >> (1 + hbd);
        if (e1 != e2) { fprintf(stderr, "rav1d_b200 host: tile %d: pass 2 consumed %zu coefficients, pass 1 wrote %zu\n", i, e2, e1); exit(4); }
        if (e1 > n_coefs) n_coefs = e1;
    }

    if (B->unsupported) {
        pthread_mutex_lock(&g_lock);
        if (!g_unsupported) snprintf(g_why, sizeof(g_why), "%s", B->why);
        g_unsupported = 1;
        pthread_mutex_unlock(&g_lock);
    }

    RbHostFrameDesc d;
    memset(&d, 0, sizeof(d));
    d.hdr.width = f->cur.p.w; d.hdr.height = f->cur.p.h; d.hdr.bpc = f->cur.p.bpc; d.hdr.layout = f->cur.p.layout;
    d.hdr.sb128 = f->seq_hdr->sb128;
    d.hdr.lf_level_y[0] = h->loopfilter.level_y[0]; d.hdr.lf_level_y[1] = h->loopfilter.level_y[1];
    d.hdr.lf_level_u = h->loopfilter.level_u; d.hdr.lf_level_v = h->loopfilter.level_v;
    d.hdr.cdef_damping = h->cdef.damping;
    for (int i = 0; i < 8; i++) { d.hdr.cdef_y_strength[i] = h->cdef.y_strength[i]; d.hdr.cdef_uv_strength[i] = h->cdef.uv_strength[i]; }
    for (int i = 0; i < 3; i++) d.hdr.lr_type[i] = f->lf.restore_planes & (1 << i) ? h->restoration.type[i] : 0;
    d.hdr.lr_unit_size_log2[0] = h->restoration.unit_size[0]; d.hdr.lr_unit_size_log2[1] = h->restoration.unit_size[1];
    d.hdr.upscaled_width = h->width[0] != h->width[1] ? h->width[1] : 0;
    d.stages = RB200_STAGE_RECON | (B->intra.n ? RB200_STAGE_INTRA : 0);
    if (c->inloop_filters & DAV1D_INLOOPFILTER_DEBLOCK) d.stages |= RB200_STAGE_DEBLOCK;
    if (f->seq_hdr->cdef && (c->inloop_filters & DAV1D_INLOOPFILTER_CDEF)) d.stages |= RB200_STAGE_CDEF;
    if (d.hdr.upscaled_width) d.stages |= RB200_STAGE_SUPER_RES;
    if (f->lf.restore_planes && (c->inloop_filters & DAV1D_INLOOPFILTER_RESTORATION)) d.stages |= RB200_STAGE_LR;
    d.cur = rp->backend;
    for (int i = 0; i < 7; i++) {
        if (!f->refp[i].p.data[0] || !f->refp[i].p.allocator_data) continue;
        d.ref[i] = ((RbPic *)f->refp[i].p.allocator_data)->backend;
        d.ref_w[i] = f->refp[i].p.p.w; d.ref_h[i] = f->refp[i].p.p.h;
    }
    for (int i = 0; i < 7; i++) {
        memcpy(d.gmv_matrix[i], h->gmv[i].matrix, sizeof(d.gmv_matrix[i]));
        memcpy(d.gmv_abcd[i], h->gmv[i].u.abcd, sizeof(d.gmv_abcd[i]));
    }
    d.coef = f->frame_thread.cf; d.n_coefs = n_coefs;
    d.masks = (const Rb200Av1Filter *)f->lf.mask; d.n_masks = f->sb128w * f->sb128h;
    d.levels = (const uint8_t (*)[4])f->lf.level; d.n_levels = (size_t)f->b4_stride * 32 * f->sb128h;
    d.lut = (const Rb200Av1FilterLUT *)&f->lf.lim_lut;
    d.lr = (const Rb200Av1Restoration *)f->lf.lr_mask; d.n_lr = f->sr_sb128w * f->sb128h;
    Rb200FilmGrainData fg;
    if (c->apply_grain && has_grain(h)) {
        _Static_assert(sizeof(Rb200FilmGrainData) == sizeof(Dav1dFilmGrainData), "film grain parameter layout");
        memcpy(&fg, &h->film_grain.data, sizeof(fg));
        d.fg = &fg;
        d.fg_is_identity = f->seq_hdr->mtrx == DAV1D_MC_IDENTITY;
        d.stages |= RB200_STAGE_FILM_GRAIN;
    }
    d.decoder_frame = f;

    RbHostFinal fin;
    if (rb_batch_finalize(B, &fin, f->bw, f->bh, ss_hor, ss_ver)) { fprintf(stderr, "rav1d_b200 host: %s\n", B->why); exit(4); }
    if (getenv("RB200_HOST_DUMP")) dump_records(f, B, &fin);
    if (g_be->frame_stage(&d, B, &fin)) die("frame_stage failed");
    rb_final_free(&fin);
    if (g_be->frame_submit(rp->backend)) die("frame_submit failed");

    /* the reference's itxfm_add leaves the coefficients it consumed zeroed for the next frame's pass 1 */
    memset(f->frame_thread.cf, 0, n_coefs << (1 + hbd));

    g_frames++;
    g_items[0] += B->mc.n; g_items[1] += B->scaled.n; g_items[2] += B->comp.n; g_items[3] += B->warp.n;
    g_items[4] += B->obmc_above.n + B->obmc_left.n; g_items[5] += B->itx.n + B->iitx.n; g_items[6] += B->intra.n;
    g_items[7] += B->lfb.n;
}

EXPORT void dav1d_decode_frame_exit(Dav1dFrameContext *const f, const int retval) {
    static void (*real)(Dav1dFrameContext *, int);
    if (!real) real = real_sym("dav1d_decode_frame_exit");
    RbJob *const j = g_jobs ? job_of(f) : NULL;
    RbPic *const rp = f->sr_cur.p.data[0] ? f->sr_cur.p.allocator_data : NULL;
    if (j && j->active) {
        pthread_mutex_lock(&j->lock);
        if (!retval && rp) {
            hand_over(f, j, rp);
            rp->have_frame = 1;
            rp->fetched = g_be->host_pixels;
        } else if (rp) {
            rp->failed = 1;
        }
        j->active = 0;
        pthread_mutex_unlock(&j->lock);
    }
    if (rp) {
        pthread_mutex_lock(&g_lock);
        rp->pending = 0;
        pthread_cond_broadcast(&g_cond);
        pthread_mutex_unlock(&g_lock);
    }
    real(f, retval);
}

/* ------------------------------------------------------------------------------------------------ output */
static void fetch(Dav1dPicture *const p, const int grain) {
    RbPic *const rp = p->allocator_data;
    if (!rp || rp->fetched || !rp->have_frame) return;
    void *data[3] = { p->data[0], p->data[1], p->data[2] };
    const ptrdiff_t stride[2] = { p->stride[0], p->stride[1] };
    if (g_be->pic_fetch(rp->backend, data, stride, grain)) die("pic_fetch failed");
    rp->fetched = 1;
}

EXPORT int dav1d_get_picture(Dav1dContext *const c, Dav1dPicture *const out) {
    static int (*real)(Dav1dContext *, Dav1dPicture *);
    if (!real) real = real_sym("dav1d_get_picture");
    const int r = real(c, out);
    if (!r && out->data[0]) fetch(out, 0);
    return r;
}

/* Film grain on output.  With task threads the reference spreads dav1d_prep_grain / dav1d_apply_grain_row over them
 * (src/thread_task.c dav1d_task_delayed_fg); here the grained picture was produced by the frame's own submit
 * (RB200_STAGE_FILM_GRAIN) and is read back into `out`. */
EXPORT void dav1d_task_delayed_fg(Dav1dContext *const c, Dav1dPicture *const out, const Dav1dPicture *const in) {
    RbPic *const rin = in->allocator_data, *const rout = out->allocator_data;
    if (g_be->host_pixels) {
        typedef void (*apply_fn)(const Dav1dFilmGrainDSPContext *, Dav1dPicture *, const Dav1dPicture *);
        static apply_fn real8, real16;
        if (!real8) { real8 = (apply_fn)real_sym("dav1d_apply_grain_8bpc"); real16 = (apply_fn)real_sym("dav1d_apply_grain_16bpc"); }
        if (out->p.bpc == 8) real8(&c->dsp[0].fg, out, in);
        else real16(&c->dsp[(out->p.bpc >> 1) - 4].fg, out, in);
        return;
    }
    if (!rin || !rin->have_frame || !rout) die("film grain on a picture that was not decoded here");
    void *data[3] = { out->data[0], out->data[1], out->data[2] };
    const ptrdiff_t stride[2] = { out->stride[0], out->stride[1] };
    if (g_be->pic_fetch(rin->backend, data, stride, 1)) die("pic_fetch (film grain) failed");
    rout->fetched = 1;
}
