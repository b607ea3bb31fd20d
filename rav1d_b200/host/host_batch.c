/*
 * rav1d_b200 host layer: the per-frame batch as a set of growable record lists, and its conversion
 * into the layout rb200_frame_submit consumes (include/rav1d_b200.h): inter residuals bucketed by
 * transform size, the intra wavefront sorted by dependency level (rb200_intra_assign_levels) with the
 * residuals of each level behind the inter ones.  No reference types, no GPU calls.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rb200_host.h"

void *rb_vec_grow(void *v, int *cap, int need, size_t elem) {
    int n = *cap ? *cap : 256;
    while (n < need) n *= 2;
    void *p = realloc(v, (size_t)n * elem);
    if (!p) { fprintf(stderr, "rav1d_b200 host: out of memory\n"); abort(); }
    *cap = n;
    return p;
}

void rb_batch_reset(RbHostBatch *b) {
    b->mc.n = b->scaled.n = b->comp.n = b->warp.n = b->obmc_above.n = b->obmc_left.n = 0;
    b->scaled_obmc_above.n = b->scaled_obmc_left.n = 0;
    b->itx.n = b->intra.n = b->intra_itx.n = b->iitx.n = b->pal.n = b->lfb.n = 0;
    b->unsupported = 0;
    b->why[0] = 0;
}

void rb_batch_free(RbHostBatch *b) {
    free(b->mc.v); free(b->scaled.v); free(b->comp.v); free(b->warp.v); free(b->obmc_above.v); free(b->obmc_left.v);
    free(b->scaled_obmc_above.v); free(b->scaled_obmc_left.v);
    free(b->itx.v); free(b->intra.v); free(b->intra_itx.v); free(b->iitx.v); free(b->pal.v); free(b->lfb.v);
    memset(b, 0, sizeof(*b));
}

void rb_batch_unsupported(RbHostBatch *b, const char *why) {
    if (!b->unsupported) snprintf(b->why, sizeof(b->why), "%s", why);
    b->unsupported = 1;
}

void rb_final_free(RbHostFinal *f) {
    free(f->itx); free(f->intra); free(f->intra_itx); free(f->level_counts); free(f->level_itx_counts);
    memset(f, 0, sizeof(*f));
}

int rb_batch_finalize(RbHostBatch *b, RbHostFinal *out, int frame_w4, int frame_h4, int ss_hor, int ss_ver) {
    memset(out, 0, sizeof(*out));
    const int n_inter = b->itx.n, n_intra = b->intra.n, n_iitx = b->iitx.n;
    out->itx = malloc(sizeof(Rb200ItxItem) * (size_t)(n_inter + n_iitx + 1));
    out->intra = malloc(sizeof(Rb200IntraItem) * (size_t)(n_intra + 1));
    out->intra_itx = malloc(sizeof(int32_t) * (size_t)(n_intra + 1));
    if (!out->itx || !out->intra || !out->intra_itx) return -12;

    /* inter residuals: stable bucket sort by transform size (one launch per size), luma first within a size
     * (rb200_frame_set_plane_counts: the luma and the chroma reconstruction can then run as two chains) */
    int start[RB200_N_RECT_TX_SIZES + 1] = { 0 }, cstart[RB200_N_RECT_TX_SIZES];
    for (int i = 0; i < n_inter; i++) {
        out->itx_counts[b->itx.v[i].tx]++;
        if (!b->itx.v[i].plane) out->itx_luma_counts[b->itx.v[i].tx]++;
    }
    for (int t = 0; t < RB200_N_RECT_TX_SIZES; t++) {
        start[t + 1] = start[t] + out->itx_counts[t];
        cstart[t] = start[t] + out->itx_luma_counts[t];
    }
    for (int i = 0; i < n_inter; i++) {
        const int t = b->itx.v[i].tx;
        out->itx[b->itx.v[i].plane ? cstart[t]++ : start[t]++] = b->itx.v[i];
    }
    out->n_itx_inter = n_inter;
    out->n_itx = n_inter;
    out->n_intra = n_intra;
    if (!n_intra) return 0;

    /* dependency levels from the items in decode order, then the level-sorted order */
    const int max_levels = n_intra + 1;
    int32_t *order = malloc(sizeof(int32_t) * (size_t)n_intra);
    out->level_counts = calloc((size_t)max_levels, sizeof(int32_t));
    if (!order || !out->level_counts) { free(order); return -12; }
    int n_levels = 0;
    const int r = rb200_intra_assign_levels(b->intra.v, n_intra, frame_w4, frame_h4, ss_hor, ss_ver, order, out->level_counts,
                                            max_levels, &n_levels);
    if (r) { free(order); snprintf(b->why, sizeof(b->why), "intra levels: %s", rb200_last_error()); return r; }
    out->n_levels = n_levels;
    out->level_itx_counts = calloc((size_t)n_levels * RB200_N_RECT_TX_SIZES + 1, sizeof(int32_t));
    if (!out->level_itx_counts) { free(order); return -12; }
    int k = 0, itx_pos = n_inter;
    for (int l = 0; l < n_levels; l++) {
        int32_t *cnt = out->level_itx_counts + (size_t)l * RB200_N_RECT_TX_SIZES;
        const int k0 = k, k1 = k + out->level_counts[l];
        for (int j = k0; j < k1; j++) {
            const int src = b->intra_itx.v[order[j]];
            if (src >= 0) cnt[b->iitx.v[src].tx]++;
        }
        int lstart[RB200_N_RECT_TX_SIZES + 1];
        lstart[0] = itx_pos;
        for (int t = 0; t < RB200_N_RECT_TX_SIZES; t++) lstart[t + 1] = lstart[t] + cnt[t];
        for (int j = k0; j < k1; j++) {
            const int i = order[j], src = b->intra_itx.v[i];
            out->intra[j] = b->intra.v[i];
            if (src >= 0) {
                const int dst = lstart[b->iitx.v[src].tx]++;
                out->itx[dst] = b->iitx.v[src];
                out->intra_itx[j] = dst;
            } else {
                out->intra_itx[j] = -1;
            }
        }
        itx_pos = lstart[RB200_N_RECT_TX_SIZES];
        k = k1;
    }
    out->n_itx = itx_pos;
    free(order);
    return 0;
}
