/*
 * rav1d_b200 host layer -- what the reference's pass 2 and filter tasks become when the
 * reconstruction and the post-filters run on the GPU.
 *
 * The reference keeps its per-bit-depth frame drivers behind Rav1dFrameContext_bd_fn
 * (src/internal.rs:350-395; C twin: Dav1dFrameContext_bd_fn, src/internal.h:232-247):
 * recon_b_intra, recon_b_inter, filter_sbrow_{deblock_cols,deblock_rows,cdef,resize,lr},
 * backup_ipred_edge.  This directory re-implements exactly those entry points so that they
 * APPEND to a per-frame batch instead of calling the DSP (recon_batch_tmpl.c), plus the frame
 * life cycle around them (host_frame.c): batch hand-over when the frame's tasks are done,
 * device-resident reference pictures tied to the Dav1dPicAllocator pair, pinned-host
 * read-back when a picture is output, film grain on output.
 *
 * No Rust toolchain exists in the build image, so the host side is written against the
 * reference's C twin (same structures, same call sites) and hooked into the compiled reference
 * decoder by symbol interposition; INTEGRATION.md shows the same edits in recon.rs.
 *
 * This header is free of reference types: it is the seam between the batch builder and a
 * backend.  The product backend (backend_gpu.c) drives include/rav1d_b200.h; tests link a CPU
 * checker (oracle/ref_backend.c) behind the same seam to pin the batch builder without a GPU.
 */
#ifndef RB200_HOST_H
#define RB200_HOST_H

#include <stddef.h>
#include <stdint.h>

#include "../../include/rav1d_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Growable record list */
#define RB_VEC(T) struct { T *v; int n, cap; }

/* One frame's batch, in the order the reference's pass 2 visits the blocks. */
typedef struct RbHostBatch {
    RB_VEC(Rb200McItem) mc;          /* single-reference predictions, same-size reference */
    RB_VEC(Rb200McScaledItem) scaled;/* ... reference of another size */
    RB_VEC(Rb200McScaledItem) scaled_obmc_above, scaled_obmc_left;   /* OBMC strips from references of another size */
    RB_VEC(Rb200CompItem) comp;      /* compound blocks */
    RB_VEC(Rb200WarpItem) warp;      /* warped blocks */
    RB_VEC(Rb200McItem) obmc_above, obmc_left;
    RB_VEC(Rb200ItxItem) itx;        /* residuals of inter blocks, decode order */
    RB_VEC(Rb200IntraItem) intra;    /* intra / inter-intra / palette / residual-only items, decode order */
    RB_VEC(int32_t) intra_itx;       /* per intra item: index into iitx or -1 */
    RB_VEC(Rb200ItxItem) iitx;       /* residuals attached to intra items */
    RB_VEC(uint8_t) pal;             /* palette records (16-byte units) */
    RB_VEC(Rb200LfBlock) lfb;        /* one record per coded block (pass 1) */
    int unsupported;                 /* a block needed something the batch formats cannot express */
    char why[96];
} RbHostBatch;

void rb_batch_reset(RbHostBatch *b);
void rb_batch_free(RbHostBatch *b);
void rb_batch_unsupported(RbHostBatch *b, const char *why);
void *rb_vec_grow(void *v, int *cap, int need, size_t elem);
#define RB_PUSH(vec) (((vec).n == (vec).cap ? ((vec).v = rb_vec_grow((vec).v, &(vec).cap, (vec).n + 1, sizeof(*(vec).v))) : 0), &(vec).v[(vec).n++])

/* The batch in the layout rb200_frame_submit consumes (host_batch.c): inter residuals bucketed by transform size, the
 * intra wavefront sorted by dependency level with its residuals behind the inter ones. */
typedef struct RbHostFinal {
    Rb200ItxItem *itx; int n_itx_inter, n_itx;        /* [n_itx]: inter (bucketed), then the intra levels' residuals */
    int32_t itx_counts[RB200_N_RECT_TX_SIZES];        /* inter only */
    int32_t itx_luma_counts[RB200_N_RECT_TX_SIZES];   /* of those, plane 0: every bucket holds its luma items first */
    Rb200IntraItem *intra; int32_t *intra_itx; int n_intra;   /* level order; intra_itx = index into itx or -1 */
    int n_levels; int32_t *level_counts, *level_itx_counts;   /* [n_levels], [n_levels][RB200_N_RECT_TX_SIZES] */
} RbHostFinal;
/* frame_w4 / frame_h4: f.bw / f.bh.  Returns 0 or a negative code (message in b->why). */
int rb_batch_finalize(RbHostBatch *b, RbHostFinal *out, int frame_w4, int frame_h4, int ss_hor, int ss_ver);
void rb_final_free(RbHostFinal *f);

/* Everything a backend needs to run one frame, free of reference types. */
typedef struct RbHostFrameDesc {
    Rb200FrameHeader hdr;
    int stages;                           /* RB200_STAGE_* */
    void *cur;                            /* backend handle of the output picture */
    void *ref[7];                         /* backend handles of f.refp[0..6] (NULL = unused) */
    int ref_w[7], ref_h[7];
    int32_t gmv_matrix[7][6]; int16_t gmv_abcd[7][4];   /* frame_hdr.gmv[i]: global-motion warp of reference i */
    const void *coef; size_t n_coefs;     /* f.frame_thread.cf and the number of coefficients used */
    const Rb200Av1Filter *masks; int n_masks;           /* f.lf.mask (cdef_idx; everything when there are no records) */
    const uint8_t (*levels)[4]; size_t n_levels;        /* f.lf.level */
    const Rb200Av1FilterLUT *lut;
    const Rb200Av1Restoration *lr; int n_lr;            /* f.lf.lr_mask, sr_sb128w * sb128h */
    const Rb200FilmGrainData *fg; int fg_is_identity;   /* NULL: no grain on output */
    void *decoder_frame;                  /* the reference's frame context (used by the CPU checker only) */
} RbHostFrameDesc;

/* A backend: the product one (backend_gpu.c) or the CPU checker of the tests. */
typedef struct RbHostBackend {
    const char *name;
    int host_pixels;   /* 1: the pictures' host planes hold the pixels at all times (CPU checker); 0: they live on the device */
    int (*init)(void);
    /* host memory of a picture (pinned for the GPU backend) */
    void *(*host_alloc)(size_t bytes);
    void (*host_free)(void *p);
    /* per-picture handle, tied to the Dav1dPicAllocator alloc / release pair */
    void *(*pic_new)(void);
    void (*pic_free)(void *pic);
    /* stage: copy the batch and the filter metadata out of decoder-owned memory (must not block on other frames);
     * submit: queue the frame's work; every reference of the frame has been submitted before */
    int (*frame_stage)(const RbHostFrameDesc *d, const RbHostBatch *b, const RbHostFinal *fin);
    int (*frame_submit)(void *cur);
    /* make the host pixels of a picture valid (data / stride: the picture's host planes) */
    int (*pic_fetch)(void *pic, void *const data[3], const ptrdiff_t stride[2], int grain);
    const char *(*last_error)(void);
} RbHostBackend;

const RbHostBackend *rb200_host_backend(void);

#ifdef __cplusplus
}
#endif
#endif
