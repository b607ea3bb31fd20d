/*
 * rav1d_b200 host layer: the product backend -- every frame goes through the C ABI of include/rav1d_b200.h
 * (rb200_frame_*), nothing else.  No reference types, no CPU path: if the library or the GPU is missing, init fails.
 *
 * Pictures: a picture owns an Rb200Frame context (three device plane sets + the batch staging) from the moment a
 * frame is decoded into it until the Dav1dPicAllocator release callback; its output planes are what later frames
 * name with rb200_frame_set_ref.  Contexts are pooled by geometry, like the reference's picture memory pool
 * (src/mem.rs / Dav1dMemPool).  All frames of the stream run on ONE CUDA stream, so a frame's launches are ordered
 * behind the frames it predicts from without host synchronisation; the host only waits when a picture is output
 * (rb200_frame_readback into the pinned planes the allocator handed to the decoder).
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rb200_host.h"

typedef struct GpuPic {
    Rb200Frame *fr;
    Rb200FrameHeader hdr;
    size_t n_coefs;
    int32_t itx_counts[RB200_N_RECT_TX_SIZES];
    int n_mc, stages, staged;
} GpuPic;

typedef struct PoolEntry { Rb200Frame *fr; Rb200FrameHeader hdr; struct PoolEntry *next; } PoolEntry;

static void *g_stream;
static PoolEntry *g_pool;
static pthread_mutex_t g_pool_lock = PTHREAD_MUTEX_INITIALIZER;   /* pictures are released from any thread */
static char g_err[600];

static int fail(const char *what) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, rb200_last_error());
    return -1;
}
static const char *gpu_last_error(void) { return g_err; }

static int gpu_init(void) {
    if (rb200_abi_version() != RB200_ABI_VERSION) { snprintf(g_err, sizeof(g_err), "librav1d_b200 ABI mismatch"); return -1; }
    if (rb200_init(-1)) return fail("rb200_init");
    if (rb200_stream_create(&g_stream)) return fail("rb200_stream_create");
    return 0;
}

static void *gpu_host_alloc(size_t bytes) {
    void *p = NULL;
    if (rb200_malloc_host(&p, bytes)) return NULL;
    return p;
}
static void gpu_host_free(void *p) { if (p) rb200_free_host(p); }

static void *gpu_pic_new(void) { return calloc(1, sizeof(GpuPic)); }

static int same_geometry(const Rb200FrameHeader *a, const Rb200FrameHeader *b) {
    return a->width == b->width && a->height == b->height && a->bpc == b->bpc && a->layout == b->layout &&
           a->sb128 == b->sb128 && a->upscaled_width == b->upscaled_width;
}

static void gpu_pic_free(void *pic) {
    GpuPic *const p = pic;
    if (!p) return;
    if (p->fr) {    /* back to the pool; the next user waits for the stream before it touches the staging */
        PoolEntry *const e = malloc(sizeof(*e));
        if (e) {
            e->fr = p->fr; e->hdr = p->hdr;
            pthread_mutex_lock(&g_pool_lock);
            e->next = g_pool; g_pool = e;
            pthread_mutex_unlock(&g_pool_lock);
        } else rb200_frame_destroy(p->fr);
    }
    free(p);
}

static Rb200Frame *acquire(const Rb200FrameHeader *const h) {
    pthread_mutex_lock(&g_pool_lock);
    for (PoolEntry **pe = &g_pool; *pe; pe = &(*pe)->next) {
        if (!same_geometry(&(*pe)->hdr, h)) continue;
        PoolEntry *const e = *pe;
        Rb200Frame *const fr = e->fr;
        *pe = e->next;
        pthread_mutex_unlock(&g_pool_lock);
        free(e);
        return fr;
    }
    pthread_mutex_unlock(&g_pool_lock);
    /* worst case of a frame: one residual per 4x4 of every plane, one prediction per 4x4 of every plane */
    const int bw = ((h->width + 7) >> 3) << 1, bh = ((h->height + 7) >> 3) << 1;
    const size_t sb128 = (size_t)((bw + 31) >> 5) * ((bh + 31) >> 5);
    const size_t per_sb = h->layout == RB200_LAYOUT_I400 ? 1 * 16384 : h->layout == RB200_LAYOUT_I420 ? 24576 :
                          h->layout == RB200_LAYOUT_I422 ? 2 * 16384 : 3 * 16384;
    Rb200Frame *fr = NULL;
    if (rb200_frame_create(&fr, h, sb128 * per_sb, 3 * bw * bh + 64, 3 * bw * bh + 64)) return NULL;
    if (rb200_frame_set_stream(fr, g_stream)) { rb200_frame_destroy(fr); return NULL; }
    return fr;
}

#define CHECK(call) do { if (call) return fail(#call); } while (0)

static int gpu_frame_stage(const RbHostFrameDesc *const d, const RbHostBatch *const B, const RbHostFinal *const fin) {
    GpuPic *const p = d->cur;
    if (!p->fr) {
        p->fr = acquire(&d->hdr);
        if (!p->fr) return fail("rb200_frame_create");
    }
    Rb200Frame *const fr = p->fr;
    p->hdr = d->hdr;
    CHECK(rb200_frame_wait(fr));              /* an earlier use of this context may still be reading its staging */
    CHECK(rb200_frame_set_params(fr, &d->hdr));
    Rb200FrameGeometry g;
    CHECK(rb200_frame_geometry(fr, &g));
    const size_t cs = d->hdr.bpc > 8 ? 4 : 2;

    /* reconstruction batch */
    if (d->n_coefs) memcpy(rb200_frame_coef_buffer(fr), d->coef, d->n_coefs * cs);
    if (fin->n_itx) memcpy(rb200_frame_itx_items(fr), fin->itx, (size_t)fin->n_itx * sizeof(Rb200ItxItem));
    /* put predictions luma first, and the promise that goes with it: frames without compound / warped / OBMC / scaled /
     * intra items are then reconstructed as a luma and a chroma chain */
    {
        Rb200McItem *const dst = rb200_frame_mc_items(fr);
        int n = 0;
        for (int i = 0; i < B->mc.n; i++) if (!B->mc.v[i].plane) dst[n++] = B->mc.v[i];
        const int n_luma = n;
        for (int i = 0; i < B->mc.n; i++) if (B->mc.v[i].plane) dst[n++] = B->mc.v[i];
        CHECK(rb200_frame_set_plane_counts(fr, n_luma, fin->itx_luma_counts));
    }
    CHECK(rb200_frame_reserve_comp_items(fr, B->comp.n));
    if (B->comp.n) memcpy(rb200_frame_comp_items(fr), B->comp.v, (size_t)B->comp.n * sizeof(Rb200CompItem));
    CHECK(rb200_frame_set_comp_count(fr, B->comp.n));
    CHECK(rb200_frame_reserve_warp_items(fr, B->warp.n));
    if (B->warp.n) memcpy(rb200_frame_warp_items(fr), B->warp.v, (size_t)B->warp.n * sizeof(Rb200WarpItem));
    CHECK(rb200_frame_set_warp_count(fr, B->warp.n));
    /* predictions from references of another size: the blocks' own, then the OBMC strips (above, left) */
    CHECK(rb200_frame_reserve_scaled_items(fr, B->scaled.n + B->scaled_obmc_above.n + B->scaled_obmc_left.n));
    if (B->scaled.n) memcpy(rb200_frame_scaled_items(fr), B->scaled.v, (size_t)B->scaled.n * sizeof(Rb200McScaledItem));
    if (B->scaled_obmc_above.n)
        memcpy(rb200_frame_scaled_items(fr) + B->scaled.n, B->scaled_obmc_above.v, (size_t)B->scaled_obmc_above.n * sizeof(Rb200McScaledItem));
    if (B->scaled_obmc_left.n)
        memcpy(rb200_frame_scaled_items(fr) + B->scaled.n + B->scaled_obmc_above.n, B->scaled_obmc_left.v,
               (size_t)B->scaled_obmc_left.n * sizeof(Rb200McScaledItem));
    CHECK(rb200_frame_set_scaled_count(fr, B->scaled.n));
    CHECK(rb200_frame_set_scaled_obmc_counts(fr, B->scaled_obmc_above.n, B->scaled_obmc_left.n));
    const int n_obmc = B->obmc_above.n + B->obmc_left.n;
    CHECK(rb200_frame_reserve_obmc_items(fr, n_obmc));
    if (B->obmc_above.n) memcpy(rb200_frame_obmc_items(fr), B->obmc_above.v, (size_t)B->obmc_above.n * sizeof(Rb200McItem));
    if (B->obmc_left.n) memcpy(rb200_frame_obmc_items(fr) + B->obmc_above.n, B->obmc_left.v, (size_t)B->obmc_left.n * sizeof(Rb200McItem));
    CHECK(rb200_frame_set_obmc_counts(fr, B->obmc_above.n, B->obmc_left.n));
    /* intra wavefront */
    CHECK(rb200_frame_reserve_intra_items(fr, fin->n_intra, fin->n_levels));
    if (fin->n_intra) {
        memcpy(rb200_frame_intra_items(fr), fin->intra, (size_t)fin->n_intra * sizeof(Rb200IntraItem));
        memcpy(rb200_frame_intra_itx_index(fr), fin->intra_itx, (size_t)fin->n_intra * sizeof(int32_t));
    }
    CHECK(rb200_frame_set_intra_levels(fr, fin->n_levels, fin->level_counts, fin->level_itx_counts));
    CHECK(rb200_frame_reserve_palette(fr, (size_t)B->pal.n));
    if (B->pal.n) memcpy(rb200_frame_palette_buffer(fr), B->pal.v, (size_t)B->pal.n);
    CHECK(rb200_frame_set_palette_bytes(fr, (size_t)B->pal.n));
    /* references: the output planes of the pictures' own contexts, at their own sizes */
    for (int i = 0; i < 7; i++) {
        const GpuPic *const r = d->ref[i];
        if (!r || !r->fr) continue;
        Rb200Planes planes;
        CHECK(rb200_frame_output_planes(r->fr, &planes));
        CHECK(rb200_frame_set_ref(fr, i, &planes));
        CHECK(rb200_frame_set_ref_size(fr, i, d->ref_w[i], d->ref_h[i]));
        CHECK(rb200_frame_set_ref_gmv(fr, i, d->gmv_matrix[i], d->gmv_abcd[i]));
    }
    /* filter metadata: masks (cdef_idx; the rest is rebuilt on the device from the block records when there are any),
     * levels, limits, restoration units */
    memcpy(rb200_frame_lf_masks(fr), d->masks, (size_t)d->n_masks * sizeof(Rb200Av1Filter));
    CHECK(rb200_frame_reserve_lf_blocks(fr, B->lfb.n));
    if (B->lfb.n) memcpy(rb200_frame_lf_blocks(fr), B->lfb.v, (size_t)B->lfb.n * sizeof(Rb200LfBlock));
    else memcpy(rb200_frame_lf_levels(fr), d->levels, d->n_levels * 4);
    CHECK(rb200_frame_set_lf_block_count(fr, B->lfb.n));
    memcpy(rb200_frame_lf_lut(fr), d->lut, sizeof(Rb200Av1FilterLUT));
    memcpy(rb200_frame_lr_masks(fr), d->lr, (size_t)d->n_lr * sizeof(Rb200Av1Restoration));
    if (d->fg) CHECK(rb200_frame_set_film_grain(fr, d->fg, d->fg_is_identity));

    p->n_coefs = d->n_coefs;
    memcpy(p->itx_counts, fin->itx_counts, sizeof(p->itx_counts));
    p->n_mc = B->mc.n;
    p->stages = d->stages;
    p->staged = 1;
    (void)g;
    return 0;
}

static int gpu_frame_submit(void *const cur) {
    GpuPic *const p = cur;
    if (!p || !p->staged) { snprintf(g_err, sizeof(g_err), "frame_submit: nothing staged"); return -1; }
    p->staged = 0;
    static int validate = -1;      /* RB200_VALIDATE=1: check every record before the kernels trust it (untrusted bitstreams, debugging) */
    if (validate < 0) validate = getenv("RB200_VALIDATE") && atoi(getenv("RB200_VALIDATE"));
    if (validate) CHECK(rb200_frame_validate(p->fr, p->n_coefs, p->itx_counts, p->n_mc, p->stages));
    CHECK(rb200_frame_submit(p->fr, p->n_coefs, p->itx_counts, p->n_mc, p->stages, RB200_UPLOAD_ALL));
    return 0;
}

static int gpu_pic_fetch(void *const pic, void *const data[3], const ptrdiff_t stride[2], const int grain) {
    GpuPic *const p = pic;
    (void)grain;   /* rb200_frame_readback returns the grained picture when the frame ran RB200_STAGE_FILM_GRAIN */
    if (!p || !p->fr) { snprintf(g_err, sizeof(g_err), "pic_fetch: no frame was decoded into this picture"); return -1; }
    CHECK(rb200_frame_readback(p->fr, data, stride));
    CHECK(rb200_frame_wait(p->fr));
    return 0;
}

static const RbHostBackend g_backend = {
    "gpu", 0, gpu_init, gpu_host_alloc, gpu_host_free, gpu_pic_new, gpu_pic_free, gpu_frame_stage, gpu_frame_submit,
    gpu_pic_fetch, gpu_last_error,
};
const RbHostBackend *rb200_host_backend(void) { return &g_backend; }
