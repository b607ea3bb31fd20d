/*
 * rav1d_b200 -- C ABI of the B200 (sm_100a) implementation of rav1d's
 * per-block reconstruction and in-loop post-filter DSP.
 *
 * Plain C, plain pointers and sizes; no CUDA or torch types.  `void *stream`
 * arguments are a cudaStream_t (NULL = default stream).  Unless stated
 * otherwise functions return 0 on success or a negative errno-style code and
 * leave a message retrievable with rb200_last_error().
 *
 * Two layers:
 *  (1) per-call entry points with the reference's own DSP function-pointer
 *      signatures (host pointers, synchronous, re-entrant), plus *_dsp_init()
 *      functions that fill tables laid out exactly like the reference's
 *      `Rav1dDSPContext` members -- the drop-in for `f.dsp`
 *      (src/internal.rs:111-121, filled at src/decode.rs:4739-4774);
 *  (2) a device-resident batch / frame API (rb200_*_batch, rb200_frame_*)
 *      which is what the modified recon.rs / *_apply.rs drivers call: the host
 *      appends coefficients, modes and motion vectors, the GPU runs each stage
 *      as frame-level launches.
 *
 * All strides are in BYTES and may be negative in layer (1), as in the
 * reference (include/common/bitdepth.rs:113-120).  Every pixel function takes
 * `bitdepth_max` (255, 1023 or 4095) like the Rust fn-pointer ABI
 * (include/common/bitdepth.rs:187-195): <=255 selects the 8-bit class
 * (pixel = uint8_t, coef = int16_t), otherwise the 16-bit class
 * (pixel = uint16_t, coef = int32_t).
 */
#ifndef RAV1D_B200_H
#define RAV1D_B200_H

#include <stddef.h>
#include <stdint.h>

/* Batch records are 16 or 32 bytes and live in 16-byte aligned arrays (the library allocates them); the
 * attribute only tells compilers so -- it does not change any field offset or size. */
#if defined(__GNUC__) || defined(__clang__) || defined(__CUDACC__)
#define RB200_ALIGN16 __attribute__((aligned(16)))
#else
#define RB200_ALIGN16
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ core */
#define RB200_ABI_VERSION 2
int rb200_abi_version(void);
/* Bind the calling thread to a CUDA device (default: current device). */
int rb200_init(int device);
/* Last error message of the calling thread ("" if none). */
const char *rb200_last_error(void);
/* The reference's DSP functions return `()` and cannot fail (SURVEY 8b
 * "Errors"); a GPU failure inside a table slot is reported out-of-band through
 * this callback (default: message to stderr + abort()). */
typedef void (*rb200_error_cb)(void *cookie, int code, const char *msg);
void rb200_set_error_callback(rb200_error_cb cb, void *cookie);
void rb200_report_fatal(const char *where);
/* Device memory helpers so that non-CUDA hosts (Rust, Python ctypes) can drive layer (2). */
int rb200_malloc(void **dptr, size_t bytes);
int rb200_free(void *dptr);
int rb200_malloc_host(void **hptr, size_t bytes); /* pinned */
int rb200_free_host(void *hptr);
int rb200_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream);
int rb200_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream);
int rb200_memset(void *dst, int value, size_t bytes, void *stream);
int rb200_stream_sync(void *stream);
int rb200_stream_create(void **stream);   /* non-blocking stream; hand it to rb200_frame_set_stream to chain frames */
int rb200_stream_destroy(void *stream);

/* Up to three device planes of one picture (Y, U, V); stride in bytes. */
typedef struct Rb200Planes {
    void *data[3];
    int64_t stride[3];
} Rb200Planes;

/* ------------------------------------------------------------------- itx */
/* enum RectTxfmSize, src/levels.rs:31-59 */
enum {
    RB200_TX_4X4, RB200_TX_8X8, RB200_TX_16X16, RB200_TX_32X32, RB200_TX_64X64,
    RB200_RTX_4X8, RB200_RTX_8X4, RB200_RTX_8X16, RB200_RTX_16X8, RB200_RTX_16X32,
    RB200_RTX_32X16, RB200_RTX_32X64, RB200_RTX_64X32, RB200_RTX_4X16, RB200_RTX_16X4,
    RB200_RTX_8X32, RB200_RTX_32X8, RB200_RTX_16X64, RB200_RTX_64X16, RB200_N_RECT_TX_SIZES
};
/* enum TxfmType, src/levels.rs:63-82 */
enum {
    RB200_DCT_DCT, RB200_ADST_DCT, RB200_DCT_ADST, RB200_ADST_ADST, RB200_FLIPADST_DCT,
    RB200_DCT_FLIPADST, RB200_FLIPADST_FLIPADST, RB200_ADST_FLIPADST, RB200_FLIPADST_ADST,
    RB200_IDTX, RB200_V_DCT, RB200_H_DCT, RB200_V_ADST, RB200_H_ADST, RB200_V_FLIPADST,
    RB200_H_FLIPADST, RB200_WHT_WHT, RB200_N_TX_TYPES_PLUS_LL
};

/* itxfm_fn, src/itx.rs:190-191:
 *   (dst, dst_stride, coeff, eob, bitdepth_max) -> ()
 * dst += inverse transform of coeff (column-major sw x sh coefficients), and the
 * consumed coefficients are zeroed (src/itx.rs:94,152-158). */
typedef void (*rb200_itxfm_fn)(void *dst, ptrdiff_t dst_stride, void *coeff, int eob, int bitdepth_max);
/* Rav1dInvTxfmDSPContext, src/itx.rs:193-196 */
typedef struct Rb200InvTxfmDSPContext {
    rb200_itxfm_fn itxfm_add[RB200_N_RECT_TX_SIZES][RB200_N_TX_TYPES_PLUS_LL];
} Rb200InvTxfmDSPContext;
/* rav1d_itx_dsp_init, src/itx.rs:1072-1105.  Slots the reference leaves unset stay NULL. */
void rb200_itx_dsp_init(Rb200InvTxfmDSPContext *c, int bpc);
/* Same operation with the table indices as arguments; returns an error code. */
int rb200_itxfm_add(int tx, int txtp, void *dst, ptrdiff_t dst_stride, void *coeff, int eob, int bitdepth_max);
int rb200_itx_valid(int tx, int txtp);

/* Batch form: one record per transform block, what recon.rs appends instead of
 * calling itxfm_add (call sites src/recon.rs:1781,2674,3116,4013). */
typedef struct Rb200ItxItem {
    uint32_t cf_off;  /* offset of the block's coefficients in the frame's coef buffer, in coefs */
    uint16_t x, y;    /* top-left, pixels, in `plane` */
    uint8_t plane;    /* 0..2 */
    uint8_t tx;       /* RB200_TX_* / RB200_RTX_* */
    uint8_t txtp;     /* RB200_*_* */
    uint8_t ncols;    /* number of leading coefficient columns that can be non-zero (1 + the largest x of a
                         non-zero coefficient, as the front end saw while writing them); 0 = unknown, read
                         all min(w, 32).  Lets the kernel skip the zero tail of the column-major block. */
    int16_t eob;
    int16_t pad;
} RB200_ALIGN16 Rb200ItxItem;       /* 16 bytes */
/* d_items (device) sorted by tx size; counts[t] = number of items of size t.
 * Coefficients are read, not zeroed: the host zeroes its own staging copy. */
int rb200_itx_add_batch(const Rb200Planes *planes, const void *d_coef, const Rb200ItxItem *d_items,
                        const int32_t counts[RB200_N_RECT_TX_SIZES], int bitdepth_max, void *stream);

/* -------------------------------------------------------------------- mc */
/* enum Filter2d, src/levels.rs:172-183 */
enum {
    RB200_FILTER_2D_8TAP_REGULAR, RB200_FILTER_2D_8TAP_REGULAR_SMOOTH, RB200_FILTER_2D_8TAP_REGULAR_SHARP,
    RB200_FILTER_2D_8TAP_SHARP_REGULAR, RB200_FILTER_2D_8TAP_SHARP_SMOOTH, RB200_FILTER_2D_8TAP_SHARP,
    RB200_FILTER_2D_8TAP_SMOOTH_REGULAR, RB200_FILTER_2D_8TAP_SMOOTH, RB200_FILTER_2D_8TAP_SMOOTH_SHARP,
    RB200_FILTER_2D_BILINEAR, RB200_N_2D_FILTERS
};
/* fn-pointer types of Rav1dMCDSPContext, src/mc.rs:1174-1320 (trailing bitdepth_max in every
 * pixel function, as in the Rust ABI; blend* and emu_edge take none, as in the reference). */
typedef void (*rb200_mc_fn)(void *dst, ptrdiff_t dst_stride, const void *src, ptrdiff_t src_stride,
                            int w, int h, int mx, int my, int bitdepth_max);
typedef void (*rb200_mc_scaled_fn)(void *dst, ptrdiff_t dst_stride, const void *src, ptrdiff_t src_stride,
                                   int w, int h, int mx, int my, int dx, int dy, int bitdepth_max);
typedef void (*rb200_mct_fn)(int16_t *tmp, const void *src, ptrdiff_t src_stride,
                             int w, int h, int mx, int my, int bitdepth_max);
typedef void (*rb200_mct_scaled_fn)(int16_t *tmp, const void *src, ptrdiff_t src_stride,
                                    int w, int h, int mx, int my, int dx, int dy, int bitdepth_max);
typedef void (*rb200_avg_fn)(void *dst, ptrdiff_t dst_stride, const int16_t *tmp1, const int16_t *tmp2,
                             int w, int h, int bitdepth_max);
typedef void (*rb200_w_avg_fn)(void *dst, ptrdiff_t dst_stride, const int16_t *tmp1, const int16_t *tmp2,
                               int w, int h, int weight, int bitdepth_max);
typedef void (*rb200_mask_fn)(void *dst, ptrdiff_t dst_stride, const int16_t *tmp1, const int16_t *tmp2,
                              int w, int h, const uint8_t *mask, int bitdepth_max);
typedef void (*rb200_w_mask_fn)(void *dst, ptrdiff_t dst_stride, const int16_t *tmp1, const int16_t *tmp2,
                                int w, int h, uint8_t *mask, int sign, int bitdepth_max);
typedef void (*rb200_blend_fn)(void *dst, ptrdiff_t dst_stride, const void *tmp, int w, int h,
                               const uint8_t *mask);
typedef void (*rb200_blend_dir_fn)(void *dst, ptrdiff_t dst_stride, const void *tmp, int w, int h);
typedef void (*rb200_warp8x8_fn)(void *dst, ptrdiff_t dst_stride, const void *src, ptrdiff_t src_stride,
                                 const int16_t *abcd, int mx, int my, int bitdepth_max);
typedef void (*rb200_warp8x8t_fn)(int16_t *tmp, ptrdiff_t tmp_stride, const void *src, ptrdiff_t src_stride,
                                  const int16_t *abcd, int mx, int my, int bitdepth_max);
typedef void (*rb200_emu_edge_fn)(intptr_t bw, intptr_t bh, intptr_t iw, intptr_t ih, intptr_t x, intptr_t y,
                                  void *dst, ptrdiff_t dst_stride, const void *ref, ptrdiff_t ref_stride);
typedef void (*rb200_resize_fn)(void *dst, ptrdiff_t dst_stride, const void *src, ptrdiff_t src_stride,
                                int dst_w, int h, int src_w, int dx, int mx, int bitdepth_max);
/* Rav1dMCDSPContext, src/mc.rs:1321-1338 (same member order) */
typedef struct Rb200MCDSPContext {
    rb200_mc_fn mc[RB200_N_2D_FILTERS];
    rb200_mc_scaled_fn mc_scaled[RB200_N_2D_FILTERS];
    rb200_mct_fn mct[RB200_N_2D_FILTERS];
    rb200_mct_scaled_fn mct_scaled[RB200_N_2D_FILTERS];
    rb200_avg_fn avg;
    rb200_w_avg_fn w_avg;
    rb200_mask_fn mask;
    rb200_w_mask_fn w_mask[3]; /* 444, 422, 420 */
    rb200_blend_fn blend;
    rb200_blend_dir_fn blend_v;
    rb200_blend_dir_fn blend_h;
    rb200_warp8x8_fn warp8x8;
    rb200_warp8x8t_fn warp8x8t;
    rb200_emu_edge_fn emu_edge;
    rb200_resize_fn resize;
} Rb200MCDSPContext;
/* rav1d_mc_dsp_init, src/mc.rs:2495-2566.  `bpc` selects pixel size for the functions
 * that take no bitdepth_max (blend, blend_v, blend_h, emu_edge): 8 -> u8, 10/12 -> u16. */
void rb200_mc_dsp_init(Rb200MCDSPContext *c, int bpc);
/* Index-based entry points (error code instead of the table's void). */
int rb200_mc(int filter2d, void *dst, ptrdiff_t dst_stride, const void *src, ptrdiff_t src_stride,
             int w, int h, int mx, int my, int bitdepth_max);
int rb200_mct(int filter2d, int16_t *tmp, const void *src, ptrdiff_t src_stride,
              int w, int h, int mx, int my, int bitdepth_max);
int rb200_mc_scaled(int filter2d, void *dst, ptrdiff_t dst_stride, const void *src, ptrdiff_t src_stride,
                    int w, int h, int mx, int my, int dx, int dy, int bitdepth_max);
int rb200_mct_scaled(int filter2d, int16_t *tmp, const void *src, ptrdiff_t src_stride,
                     int w, int h, int mx, int my, int dx, int dy, int bitdepth_max);
int rb200_avg(void *dst, ptrdiff_t dst_stride, const int16_t *tmp1, const int16_t *tmp2, int w, int h,
              int bitdepth_max);
int rb200_w_avg(void *dst, ptrdiff_t dst_stride, const int16_t *tmp1, const int16_t *tmp2, int w, int h,
                int weight, int bitdepth_max);
int rb200_mask(void *dst, ptrdiff_t dst_stride, const int16_t *tmp1, const int16_t *tmp2, int w, int h,
               const uint8_t *mask, int bitdepth_max);
int rb200_w_mask(int ss /* 0:444 1:422 2:420 */, void *dst, ptrdiff_t dst_stride, const int16_t *tmp1,
                 const int16_t *tmp2, int w, int h, uint8_t *mask, int sign, int bitdepth_max);
/* One entry of dav1d_wedge_masks[bs][ss][sign][wedge_idx] (src/wedge.rs:377; C: src/wedge.c:83-243) as the compound
 * kernel evaluates it; w, h in {8, 16, 32}; mask receives (w >> ss_hor) * (h >> ss_ver) bytes. */
int rb200_wedge_mask(int w, int h, int ss /* 0:444 1:422 2:420 */, int sign, int wedge_idx, uint8_t *mask);
int rb200_blend(int dir /* 0:mask 1:v 2:h */, void *dst, ptrdiff_t dst_stride, const void *tmp, int w, int h,
                const uint8_t *mask, int bitdepth_max);
int rb200_warp8x8(void *dst, ptrdiff_t dst_stride, const void *src, ptrdiff_t src_stride,
                  const int16_t *abcd, int mx, int my, int bitdepth_max);
int rb200_warp8x8t(int16_t *tmp, ptrdiff_t tmp_stride, const void *src, ptrdiff_t src_stride,
                   const int16_t *abcd, int mx, int my, int bitdepth_max);
int rb200_emu_edge(intptr_t bw, intptr_t bh, intptr_t iw, intptr_t ih, intptr_t x, intptr_t y,
                   void *dst, ptrdiff_t dst_stride, const void *ref, ptrdiff_t ref_stride, int bitdepth_max);
int rb200_resize(void *dst, ptrdiff_t dst_stride, const void *src, ptrdiff_t src_stride,
                 int dst_w, int h, int src_w, int dx, int mx, int bitdepth_max);

/* Batch form: one record per prediction block and plane, what recon.rs `mc()`
 * (src/recon.rs:2025-2203) appends instead of calling emu_edge + mc/mct.  Source
 * coordinates may leave the reference picture: the kernel clamps them, which is
 * what emu_edge materialises (src/mc.rs:1032-1112). */
typedef struct Rb200McItem {
    int16_t dst_x, dst_y; /* top-left in the destination plane, pixels */
    int16_t src_x, src_y; /* integer source position `dx`, `dy` (src/recon.rs:2052-2055), pixels */
    uint8_t w, h;         /* block size in pixels, 2..128 */
    uint8_t plane;        /* 0..2 */
    uint8_t ref;          /* reference slot, 0..7 */
    uint8_t mx, my;       /* sub-pel phase 0..15 (already << !ss, src/recon.rs:2100-2101) */
    uint8_t filter2d;     /* RB200_FILTER_2D_* */
    uint8_t flags;        /* RB200_MC_* */
} RB200_ALIGN16 Rb200McItem;            /* 16 bytes */
enum { RB200_MC_PUT = 0, RB200_MC_OBMC_ABOVE = 1, RB200_MC_OBMC_LEFT = 2 };
/* refs[slot]: device planes of reference pictures; ref_w/ref_h: picture size of plane 0 in pixels;
 * ss_hor/ss_ver: chroma subsampling.  dst: device planes of the current picture. */
int rb200_mc_batch(const Rb200Planes *dst, const Rb200Planes *refs, int n_refs, int ref_w, int ref_h,
                   int ss_hor, int ss_ver, const Rb200McItem *d_items, int n_items, int bitdepth_max,
                   void *stream);

/* Compound prediction, one record per BLOCK (all planes): recon.rs rav1d_recon_b_inter's compound
 * branch (src/recon.rs:3290-3346 luma, :3742-3850 chroma; C: src/recon_tmpl.c:1836-1921) -- two
 * `mct` predictions per plane combined by avg / w_avg / w_mask (+ `mask` with the luma-derived,
 * sub-sampled segmentation mask for chroma).  The two int16 predictions and the mask never leave
 * shared memory.  Motion vectors are the block's own (1/8 luma pel, {y, x} like Av1Block.mv); the
 * kernel derives the per-plane position and phase as recon.rs `mc()` does (:2047-2055). */
enum { RB200_COMP_AVG = 0, RB200_COMP_WEIGHTED_AVG = 1, RB200_COMP_SEG = 2, RB200_COMP_WEDGE = 3 };
typedef struct Rb200CompItem {
    int16_t x, y;        /* top-left of the block in the luma plane, pixels */
    uint8_t w, h;        /* luma block size, 8..128 */
    uint8_t ref[2];      /* reference slots */
    int16_t mv[2][2];    /* mv[i] = {y, x}, 1/8 luma pel */
    uint8_t filter2d;
    uint8_t comp_type;   /* RB200_COMP_* */
    uint8_t jnt_weight;  /* w_avg weight, f.jnt_weights[ref0][ref1] (src/decode.rs:4354-4386) */
    uint8_t mask_sign;
    uint8_t wedge_idx;   /* RB200_COMP_WEDGE: 0..15, block sizes 8..32 (dav1d_wedge_masks, src/wedge.rs:377) */
    uint8_t warp_mask;   /* GLOBALMV_GLOBALMV blocks whose reference allows it (f.gmv_warp_allowed, src/recon.rs:3253-3268,3352-3369):
                            bit i = the luma prediction from ref[i] is the reference's global-motion warp (rb200_frame_set_ref_gmv),
                            bit 2 + i = the chroma predictions too (chroma block at least 8x8) */
    uint8_t pad[10];
} RB200_ALIGN16 Rb200CompItem;         /* 32 bytes */

/* Warped (affine) prediction, one record per BLOCK (all planes): recon.rs `warp_affine`
 * (src/recon.rs:2311-2400; C: src/recon_tmpl.c:1139-1198) -- per 8x8 the position and phase follow from
 * the warp matrix, then warp8x8 with the shear parameters.  Plane sizes must be multiples of 8. */
typedef struct Rb200WarpItem {
    int16_t x, y;        /* top-left of the block in the luma plane, pixels */
    uint8_t w, h;        /* luma block size, 8..128 (chroma must come out >= 8: 16 with sub-sampling) */
    uint8_t ref;
    uint8_t pad0;
    int32_t matrix[6];   /* Dav1dWarpedMotionParams.matrix */
    int16_t abcd[4];     /* alpha, beta, gamma, delta */
    uint8_t pad[8];
} RB200_ALIGN16 Rb200WarpItem;         /* 48 bytes */

/* ------------------------------------------------------------ loop filter */
/* Av1FilterLUT, src/lf_mask.rs:24-28 */
typedef struct Rb200Av1FilterLUT {
    uint8_t e[64];
    uint8_t i[64];
    uint64_t sharp[2];
} Rb200Av1FilterLUT;
/* Av1Filter, src/lf_mask.rs:43-52: one per 128x128 luma area */
typedef struct Rb200Av1Filter {
    uint16_t filter_y[2 /* 0=col, 1=row */][32][3][2];
    uint16_t filter_uv[2][32][2][2];
    int8_t cdef_idx[4]; /* -1 = unset */
    uint16_t noskip_mask[16][2];
} Rb200Av1Filter;
/* Av1RestorationUnit / Av1Restoration, src/lf_mask.rs:32-41,56-58 */
typedef struct Rb200Av1RestorationUnit {
    uint8_t type; /* RB200_RESTORATION_*; SGR: RB200_RESTORATION_SGRPROJ + sgr_idx */
    int8_t filter_h[3];
    int8_t filter_v[3];
    int8_t sgr_weights[2];
} Rb200Av1RestorationUnit;
typedef struct Rb200Av1Restoration {
    Rb200Av1RestorationUnit lr[3][4];
} Rb200Av1Restoration;
enum { RB200_RESTORATION_NONE, RB200_RESTORATION_SWITCHABLE, RB200_RESTORATION_WIENER, RB200_RESTORATION_SGRPROJ };

/* loopfilter_sb_fn, src/loopfilter.rs:20-29 */
typedef void (*rb200_loopfilter_sb_fn)(void *dst, ptrdiff_t stride, const uint32_t *mask,
                                       const uint8_t (*lvl)[4], ptrdiff_t lvl_stride,
                                       const Rb200Av1FilterLUT *lut, int w_or_h, int bitdepth_max);
/* Rav1dLoopFilterDSPContext, src/loopfilter.rs:31-34: [0=y,1=uv][0=h (column edges),1=v (row edges)] */
typedef struct Rb200LoopFilterDSPContext {
    rb200_loopfilter_sb_fn loop_filter_sb[2][2];
} Rb200LoopFilterDSPContext;
void rb200_loop_filter_dsp_init(Rb200LoopFilterDSPContext *c, int bpc);
int rb200_loop_filter_sb(int uv, int dir, void *dst, ptrdiff_t stride, const uint32_t *mask,
                         const uint8_t (*lvl)[4], ptrdiff_t lvl_stride, const Rb200Av1FilterLUT *lut,
                         int w_or_h, int bitdepth_max);

/* ------------------------------------------------------------------- cdef */
enum { RB200_CDEF_HAVE_LEFT = 1, RB200_CDEF_HAVE_RIGHT = 2, RB200_CDEF_HAVE_TOP = 4, RB200_CDEF_HAVE_BOTTOM = 8 };
/* cdef_fn / cdef_dir_fn, src/cdef.rs:35-50 */
typedef void (*rb200_cdef_fn)(void *dst, ptrdiff_t stride, const void *left /* [8][2] px */, const void *top,
                              const void *bottom, int pri_strength, int sec_strength, int dir, int damping,
                              uint32_t edges, int bitdepth_max);
typedef int (*rb200_cdef_dir_fn)(const void *src, ptrdiff_t stride, unsigned *var, int bitdepth_max);
/* Rav1dCdefDSPContext, src/cdef.rs:52-56 */
typedef struct Rb200CdefDSPContext {
    rb200_cdef_dir_fn dir;
    rb200_cdef_fn fb[3]; /* 8x8, 4x8, 4x4 */
} Rb200CdefDSPContext;
void rb200_cdef_dsp_init(Rb200CdefDSPContext *c, int bpc);
int rb200_cdef_dir(const void *src, ptrdiff_t stride, unsigned *var, int bitdepth_max, int *dir_out);
int rb200_cdef_fb(int idx /* 0:8x8 1:4x8 2:4x4 */, void *dst, ptrdiff_t stride, const void *left,
                  const void *top, const void *bottom, int pri_strength, int sec_strength, int dir,
                  int damping, uint32_t edges, int bitdepth_max);

/* ------------------------------------------------------- loop restoration */
enum { RB200_LR_HAVE_LEFT = 1, RB200_LR_HAVE_RIGHT = 2, RB200_LR_HAVE_TOP = 4, RB200_LR_HAVE_BOTTOM = 8 };
/* LooprestorationParams, src/looprestoration.rs:76-89 */
typedef union Rb200LooprestorationParams {
    int16_t filter[2][8] __attribute__((aligned(16)));
    struct { uint32_t s0, s1; int16_t w0, w1; } sgr;
} Rb200LooprestorationParams;
/* looprestorationfilter_fn, src/looprestoration.rs:91-101 */
typedef void (*rb200_lr_fn)(void *dst, ptrdiff_t stride, const void *left /* [h][4] px */, const void *lpf,
                            int w, int h, const Rb200LooprestorationParams *params, uint32_t edges,
                            int bitdepth_max);
/* Rav1dLoopRestorationDSPContext, src/looprestoration.rs:103-107 */
typedef struct Rb200LoopRestorationDSPContext {
    rb200_lr_fn wiener[2]; /* 7-tap, 5-tap */
    rb200_lr_fn sgr[3];    /* 5x5, 3x3, mix */
} Rb200LoopRestorationDSPContext;
void rb200_loop_restoration_dsp_init(Rb200LoopRestorationDSPContext *c, int bpc);
int rb200_lr(int kind /* 0 wiener7, 1 wiener5, 2 sgr5x5, 3 sgr3x3, 4 sgr mix */, void *dst, ptrdiff_t stride,
             const void *left, const void *lpf, int w, int h, const Rb200LooprestorationParams *params,
             uint32_t edges, int bitdepth_max);

/* ------------------------------------------------------------ film grain */
/* Dav1dFilmGrainData / Rav1dFilmGrainData, include/dav1d/headers.rs:1610-1661 (same layout) */
typedef struct Rb200FilmGrainData {
    unsigned seed;
    int num_y_points;
    uint8_t y_points[14][2]; /* value, scaling */
    int chroma_scaling_from_luma;
    int num_uv_points[2];
    uint8_t uv_points[2][10][2];
    int scaling_shift;
    int ar_coeff_lag;
    int8_t ar_coeffs_y[24];
    int8_t ar_coeffs_uv[2][25 + 3];
    uint64_t ar_coeff_shift;
    int grain_scale_shift;
    int uv_mult[2];
    int uv_luma_mult[2];
    int uv_offset[2];
    int overlap_flag;
    int clip_to_restricted_range;
} Rb200FilmGrainData;
#define RB200_GRAIN_WIDTH 82
#define RB200_GRAIN_HEIGHT 73
/* fn-pointer types of Rav1dFilmGrainDSPContext, src/filmgrain.rs:41-143.  `buf` / `grain_lut` are
 * entry[GRAIN_HEIGHT + 1][GRAIN_WIDTH] with entry = int8_t (8 bpc) or int16_t; scaling is
 * uint8_t[256] (8 bpc) or uint8_t[4096]. */
typedef void (*rb200_generate_grain_y_fn)(void *buf, const Rb200FilmGrainData *data, int bitdepth_max);
typedef void (*rb200_generate_grain_uv_fn)(void *buf, const void *buf_y, const Rb200FilmGrainData *data,
                                           intptr_t uv, int bitdepth_max);
typedef void (*rb200_fgy_32x32xn_fn)(void *dst_row, const void *src_row, ptrdiff_t stride,
                                     const Rb200FilmGrainData *data, size_t pw, const uint8_t *scaling,
                                     const void *grain_lut, int bh, int row_num, int bitdepth_max);
typedef void (*rb200_fguv_32x32xn_fn)(void *dst_row, const void *src_row, ptrdiff_t stride,
                                      const Rb200FilmGrainData *data, size_t pw, const uint8_t *scaling,
                                      const void *grain_lut, int bh, int row_num, const void *luma_row,
                                      ptrdiff_t luma_stride, int uv_pl, int is_id, int bitdepth_max);
/* Rav1dFilmGrainDSPContext, src/filmgrain.rs:195-201: [layout - 1] = 420, 422, 444 */
typedef struct Rb200FilmGrainDSPContext {
    rb200_generate_grain_y_fn generate_grain_y;
    rb200_generate_grain_uv_fn generate_grain_uv[3];
    rb200_fgy_32x32xn_fn fgy_32x32xn;
    rb200_fguv_32x32xn_fn fguv_32x32xn[3];
} Rb200FilmGrainDSPContext;
void rb200_film_grain_dsp_init(Rb200FilmGrainDSPContext *c, int bpc);
int rb200_generate_grain_y(void *buf, const Rb200FilmGrainData *data, int bitdepth_max);
int rb200_generate_grain_uv(int layout /* RB200_LAYOUT_I420.. */, void *buf, const void *buf_y,
                            const Rb200FilmGrainData *data, intptr_t uv, int bitdepth_max);
int rb200_fgy_32x32xn(void *dst_row, const void *src_row, ptrdiff_t stride, const Rb200FilmGrainData *data,
                      size_t pw, const uint8_t *scaling, const void *grain_lut, int bh, int row_num,
                      int bitdepth_max);
int rb200_fguv_32x32xn(int layout, void *dst_row, const void *src_row, ptrdiff_t stride,
                       const Rb200FilmGrainData *data, size_t pw, const uint8_t *scaling, const void *grain_lut,
                       int bh, int row_num, const void *luma_row, ptrdiff_t luma_stride, int uv_pl, int is_id,
                       int bitdepth_max);
/* ---------------------------------------------------------- intra prediction */
/* The per-call slots of Rav1dIntraPredDSPContext (src/ipred.rs:47-169; C: src/ipred.h:37-93, src/ipred_tmpl.c).
 * First step of the "next" row f1: function-level parity of every predictor.  `topleft` points INTO the caller's
 * edge buffer (left neighbours at negative, top neighbours at positive indices), as in the reference; the call
 * mirrors topleft[-(h + min(w, h)) .. w + min(w, h)], the range the reference's predictors may read.
 * enum IntraPredMode incl. the implementation modes (src/levels.rs:85-130): */
enum { RB200_DC_PRED = 0, RB200_VERT_PRED, RB200_HOR_PRED, RB200_LEFT_DC_PRED, RB200_TOP_DC_PRED, RB200_DC_128_PRED,
       RB200_Z1_PRED, RB200_Z2_PRED, RB200_Z3_PRED, RB200_SMOOTH_PRED, RB200_SMOOTH_V_PRED, RB200_SMOOTH_H_PRED,
       RB200_PAETH_PRED, RB200_FILTER_PRED, RB200_N_IMPL_INTRA_PRED_MODES };
typedef void (*rb200_angular_ipred_fn)(void *dst, ptrdiff_t stride, const void *topleft, int width, int height, int angle,
                                       int max_width, int max_height, int bitdepth_max);
typedef void (*rb200_cfl_ac_fn)(int16_t *ac, const void *y, ptrdiff_t stride, int w_pad, int h_pad, int cw, int ch);
typedef void (*rb200_cfl_pred_fn)(void *dst, ptrdiff_t stride, const void *topleft, int width, int height, const int16_t *ac,
                                  int alpha, int bitdepth_max);
typedef void (*rb200_pal_pred_fn)(void *dst, ptrdiff_t stride, const void *pal, const uint8_t *idx, int w, int h);
typedef struct Rb200IntraPredDSPContext {
    rb200_angular_ipred_fn intra_pred[RB200_N_IMPL_INTRA_PRED_MODES];
    rb200_cfl_ac_fn cfl_ac[3];          /* [layout - 1] = 420, 422, 444 */
    rb200_cfl_pred_fn cfl_pred[6];      /* DC, LEFT_DC, TOP_DC, DC_128 are set */
    rb200_pal_pred_fn pal_pred;
} Rb200IntraPredDSPContext;
void rb200_intra_pred_dsp_init(Rb200IntraPredDSPContext *c, int bpc);
/* angle: the reference's packed argument (angle | is_smooth << 9 | enable_intra_edge_filter << 10; the filter set for
 * RB200_FILTER_PRED) */
int rb200_ipred(int mode, void *dst, ptrdiff_t stride, const void *topleft, int width, int height, int angle,
                int max_width, int max_height, int bitdepth_max);
int rb200_cfl_ac(int ss /* layout - 1 */, int16_t *ac, const void *y, ptrdiff_t stride, int w_pad, int h_pad, int cw,
                 int ch, int bitdepth_max);
int rb200_cfl_pred(int mode, void *dst, ptrdiff_t stride, const void *topleft, int width, int height, const int16_t *ac,
                   int alpha, int bitdepth_max);
int rb200_pal_pred(void *dst, ptrdiff_t stride, const void *pal, const uint8_t *idx, int w, int h, int bitdepth_max);

/* generate_scaling, src/fg_apply.rs:14-72: piecewise-linear scaling LUT (256 or 4096 bytes). */
int rb200_generate_scaling(int bitdepth, const uint8_t points[][2], int num, uint8_t *scaling);

/* ------------------------------------------------------------ frame level */
/* The coarser drop-in: Rav1dFrameContext_bd_fn.filter_sbrow_{deblock_cols,deblock_rows,
 * cdef,lr} (src/internal.rs:368-395, bodies src/recon.rs:4047-4338 driving
 * src/lf_apply.rs:597,763, src/cdef_apply.rs:159, src/lr_apply.rs:261) and the
 * pass-2 reconstruction calls of recon.rs become frame-level launches over
 * device-resident planes.  The host hands over the arrays the reference's
 * pass 1 already produces, in the reference's own layouts. */
enum { RB200_LAYOUT_I400, RB200_LAYOUT_I420, RB200_LAYOUT_I422, RB200_LAYOUT_I444 };
enum { RB200_STAGE_RECON = 1, RB200_STAGE_DEBLOCK = 2, RB200_STAGE_CDEF = 4, RB200_STAGE_LR = 8,
       RB200_STAGE_FILM_GRAIN = 16 /* needs rb200_frame_set_film_grain */,
       RB200_STAGE_INTRA = 64      /* intra-predicted blocks, level by level (rb200_frame_set_intra_levels) */,
       RB200_STAGE_SUPER_RES = 32  /* frames created with upscaled_width > width: horizontal upscaling between CDEF
                                      and loop restoration (rav1d_filter_sbrow_resize, src/recon.rs:4215-4281) */ };
typedef struct Rb200FrameHeader {
    int32_t width, height;       /* picture size in pixels (plane 0) */
    int32_t bpc;                 /* 8, 10 or 12 */
    int32_t layout;              /* RB200_LAYOUT_* */
    int32_t sb128;               /* Dav1dSequenceHeader.sb128 */
    int32_t lf_level_y[2];       /* frame_hdr.loopfilter.level_y: deblock runs if either is non-zero */
    int32_t lf_level_u, lf_level_v; /* chroma deblock runs if either is non-zero (src/lf_apply.rs:731) */
    int32_t cdef_damping;        /* frame_hdr.cdef.damping (3..6) */
    int32_t cdef_y_strength[8];
    int32_t cdef_uv_strength[8];
    int32_t lr_type[3];          /* frame_hdr.restoration.type[plane]; NONE = plane not restored */
    int32_t lr_unit_size_log2[2];/* frame_hdr.restoration.unit_size[y, uv] */
    int32_t upscaled_width;      /* frame_hdr.width[1] when super-resolution is on (> width), else 0: the picture is
                                    coded `width` wide and upscaled after CDEF; loop restoration units, the output
                                    picture and rb200_frame_readback then use the upscaled width */
} Rb200FrameHeader;
typedef struct Rb200Frame Rb200Frame;

/* Geometry derived from the header exactly as src/decode.rs:4880-4905 does. */
typedef struct Rb200FrameGeometry {
    int32_t bw, bh;         /* 4-pixel units, rounded up to 8 pixels */
    int32_t w4, h4;
    int32_t sb128w, sb128h, sbh;
    int32_t b4_stride;
    int32_t ss_hor, ss_ver;
    int64_t stride[2];      /* device plane strides in bytes (y, uv) */
    int32_t plane_h[2];     /* allocated rows (y, uv) */
    int32_t n_planes;
} Rb200FrameGeometry;

int rb200_frame_create(Rb200Frame **out, const Rb200FrameHeader *hdr, size_t max_coefs, int max_itx_items,
                       int max_mc_items);
int rb200_frame_destroy(Rb200Frame *f);
/* A decoder reuses a context for the next picture of the same geometry: everything in the header but width, height,
 * bpc, layout, sb128 and upscaled_width may change (levels, CDEF strengths, restoration types and unit sizes). */
int rb200_frame_set_params(Rb200Frame *f, const Rb200FrameHeader *hdr);
int rb200_frame_geometry(const Rb200Frame *f, Rb200FrameGeometry *g);
/* Pinned host staging the front end writes into (the batch the north star describes). */
void *rb200_frame_coef_buffer(Rb200Frame *f);                 /* coef[max_coefs] (i16 / i32) */
Rb200ItxItem *rb200_frame_itx_items(Rb200Frame *f);           /* [max_itx_items], sorted by tx size */
Rb200McItem *rb200_frame_mc_items(Rb200Frame *f);             /* [max_mc_items] */
Rb200Av1Filter *rb200_frame_lf_masks(Rb200Frame *f);          /* [sb128h * sb128w] */
uint8_t (*rb200_frame_lf_levels(Rb200Frame *f))[4];           /* [b4_stride * 32 * sb128h] */
Rb200Av1FilterLUT *rb200_frame_lf_lut(Rb200Frame *f);
Rb200Av1Restoration *rb200_frame_lr_masks(Rb200Frame *f);     /* [sb128h * sb128w] */
/* Compound blocks: reserve the staging once, then set the count before each submit that has
 * RB200_STAGE_RECON (blocks of put items and compound items must not overlap). */
int rb200_frame_reserve_comp_items(Rb200Frame *f, int max_comp_items);
Rb200CompItem *rb200_frame_comp_items(Rb200Frame *f);
int rb200_frame_set_comp_count(Rb200Frame *f, int n_comp_items);
/* Overlapped block motion compensation, recon.rs `obmc()` (src/recon.rs:2205-2309; C: src/recon_tmpl.c:1076-1137):
 * one Rb200McItem per neighbour strip and plane, flags = RB200_MC_OBMC_ABOVE (w x h = the blend_h area,
 * h_mul * ow4 by v_mul * oh4; the neighbour's vector is predicted over w x ((oh4 * 3 + 3) >> 2) * v_mul and
 * blended with dav1d_obmc_masks over the first h * 3 / 4 rows) or RB200_MC_OBMC_LEFT (blend_v over the
 * first w * 3 / 4 columns).  The list holds all ABOVE strips first, then all LEFT strips; they run after
 * every block's own prediction and before the residuals, in that order (the reference's order per block). */
int rb200_frame_reserve_obmc_items(Rb200Frame *f, int max_obmc_items);
Rb200McItem *rb200_frame_obmc_items(Rb200Frame *f);
int rb200_frame_set_obmc_counts(Rb200Frame *f, int n_above, int n_left);
/* Prediction from a reference picture of another size (scaled references: the second branch of recon.rs
 * `mc()`, src/recon.rs:2116-2199; C: src/recon_tmpl.c:1014-1071 -> mc.mc_scaled[filter2d], src/mc.rs:212-275,
 * 496-541).  The host does the position arithmetic of that branch (scale_mv) and hands over pos_x / pos_y,
 * the 1/1024-sample position of the block's first pixel in the reference plane, and the per-pixel steps
 * f.svc[ref][0 / 1].step; the kernel clamps source coordinates to the reference plane (emu_edge).  The
 * reference's size is given with rb200_frame_set_ref_size. */
typedef struct Rb200McScaledItem {
    int16_t dst_x, dst_y;    /* in plane `plane` of the current picture, pixels */
    uint8_t w, h, plane, ref;
    int32_t pos_x, pos_y;
    int32_t step_x, step_y;
    uint8_t filter2d;
    uint8_t flags;           /* RB200_MC_PUT, or RB200_MC_OBMC_ABOVE / _LEFT: an OBMC strip predicted from a reference of
                                another size (w x h = the blend area, as for the Rb200McItem strips) */
    uint8_t pad[6];
} RB200_ALIGN16 Rb200McScaledItem;         /* 32 bytes */
int rb200_frame_set_ref_size(Rb200Frame *f, int slot, int width, int height);   /* luma size of reference `slot` */
/* Global-motion parameters of reference `slot` (frame_hdr.gmv[slot]: Rav1dWarpedMotionParams.matrix and alpha / beta /
 * gamma / delta), used by compound blocks whose Rb200CompItem.warp_mask asks for the warped prediction. */
int rb200_frame_set_ref_gmv(Rb200Frame *f, int slot, const int32_t matrix[6], const int16_t abcd[4]);
int rb200_frame_reserve_scaled_items(Rb200Frame *f, int max_scaled_items);
Rb200McScaledItem *rb200_frame_scaled_items(Rb200Frame *f);
int rb200_frame_set_scaled_count(Rb200Frame *f, int n);
/* OBMC strips from references of another size: they follow the n put items of rb200_frame_set_scaled_count in the same
 * list (reserve n + n_above + n_left), all ABOVE strips first, then all LEFT strips; they run with the ordinary strips
 * of their kind (above before left). */
int rb200_frame_set_scaled_obmc_counts(Rb200Frame *f, int n_above, int n_left);
/* Frame level (stage RB200_STAGE_INTRA): the intra half of rav1d_recon_b_intra (src/recon.rs:2402-3160) as a
 * wavefront.  One record per intra-predicted TRANSFORM block; the host gives each its dependency level (a block
 * needs the reconstructed pixels left of, above, above-right and below-left of it, so level = 1 + the highest
 * level among the intra blocks that own those pixels; blocks that only touch inter-predicted or absent
 * neighbours are level 0).  Per level the library predicts every block (edge preparation as
 * rav1d_prepare_intra_edges, src/ipred_prepare.rs:118, then the predictor, then the block's residual) in ONE launch.
 * Chroma from luma: a chroma item with mode 13 (UV_CFL_PRED) covers the whole chroma block, `angle` carries cfl_alpha of
 * its plane (non-zero; a plane with alpha 0 is a plain DC item) and bits 13-15 of w4_end / h4_end the w_pad / h_pad
 * arguments of cfl_ac; its level must exceed the levels of the block's luma transform blocks.
 * Inter-intra (src/recon.rs:3475-3550,3742-3850): flags bit 6; one item per plane over the whole block with mode DC / VERT /
 * HOR / SMOOTH, angle = -1 for the inter-intra mask of that mode (dav1d_ii_masks) or the wedge index 0 .. 15; the block's
 * inter prediction comes from the ordinary Rb200McItem lists, its residual must be attached to the intra item.
 * Palette blocks (pal_pred): mode 14, tw4 x th4 = the whole block, w4_end | h4_end << 16 = offset in 16-byte units of the
 * block's record { 8 palette entries padded to 16 bytes, w * h index bytes } in rb200_frame_palette_buffer(); the further
 * transform blocks of such a block are items of mode 15 (no prediction, only the residual).
 * Intra block copy (src/recon.rs:3196-3240 == src/recon_tmpl.c:1631-1645: mc() with the bilinear filter from the picture
 * being reconstructed) is one item of mode 16 per plane: tw4 x th4 is the whole block, (int16) w4_end / h4_end the source
 * position in plane pixels, angle = mx | my << 4 (the 1/16-pixel fractions mc() hands the filter), no residual of its
 * own -- the block's transform blocks follow as items of mode 15.  Its level is one above everything in the source area. */
typedef struct Rb200IntraItem {
    uint16_t x4, y4;          /* block position in `plane`, 4-pixel units (t.bx, t.by; >> ss for chroma) */
    uint16_t w4_end, h4_end;  /* tile end in the same units: the `w`, `h` arguments of rav1d_prepare_intra_edges (bits 0-12) */
    uint8_t plane;
    uint8_t tw4, th4;         /* transform block size, 4-pixel units (1 .. 16) */
    uint8_t mode;             /* coded IntraPredMode: DC 0, VERT 1, HOR 2, DIAG_DOWN_LEFT 3 .. VERT_LEFT 8, SMOOTH 9,
                                 SMOOTH_V 10, SMOOTH_H 11, PAETH 12, FILTER 13 */
    int8_t angle;             /* angle_delta -3 .. 3 (directional modes) or the filter-intra set 0 .. 4 */
    uint8_t flags;            /* 1 have_left, 2 have_top, 4 EDGE_TOP_HAS_RIGHT, 8 EDGE_LEFT_HAS_BOTTOM, 16 smooth
                                 neighbour (sm_flag), 32 seq_hdr.intra_edge_filter, 64 inter-intra */
    uint16_t level;
} RB200_ALIGN16 Rb200IntraItem;   /* 16 bytes */
int rb200_frame_reserve_intra_items(Rb200Frame *f, int max_items, int max_levels);
Rb200IntraItem *rb200_frame_intra_items(Rb200Frame *f);   /* sorted by level */
int32_t *rb200_frame_intra_itx_index(Rb200Frame *f);      /* per intra item: index of its residual in the frame's
                                                             Rb200ItxItem list (>= the inter count), or -1 */
/* item_counts[level]; itx_counts[level][RB200_N_RECT_TX_SIZES]: the residuals of that level's blocks, stored in the
 * frame's Rb200ItxItem list right after the inter ones (i.e. from index sum(itx_counts of rb200_frame_submit)),
 * level by level and bucketed by size within a level. */
int rb200_frame_set_intra_levels(Rb200Frame *f, int n_levels, const int32_t *item_counts, const int32_t *itx_counts);
/* Host helper (no GPU work): fills Rb200IntraItem.level of `items`, given in DECODE order, from their positions, sizes and
 * availability flags (see above), and returns the level-sorted order (order[k] = index of the k-th item to append) and the
 * number of items per level.  frame_w4 / frame_h4: luma picture size in 4-pixel units (f.bw, f.bh). */
int rb200_intra_assign_levels(Rb200IntraItem *items, int n, int frame_w4, int frame_h4, int ss_hor, int ss_ver,
                              int32_t *order, int32_t *level_counts, int max_levels, int *n_levels);
int rb200_frame_reserve_palette(Rb200Frame *f, size_t bytes);
uint8_t *rb200_frame_palette_buffer(Rb200Frame *f);
int rb200_frame_set_palette_bytes(Rb200Frame *f, size_t bytes);

int rb200_frame_reserve_warp_items(Rb200Frame *f, int max_warp_items);
Rb200WarpItem *rb200_frame_warp_items(Rb200Frame *f);
int rb200_frame_set_warp_count(Rb200Frame *f, int n_warp_items);
/* Reference pictures: device planes (layout of rb200_frame_geometry) that stay resident. */
int rb200_frame_set_ref(Rb200Frame *f, int slot, const Rb200Planes *planes);
/* Upload a host picture into one of the frame's own plane sets (0 = current/recon).  The copy covers the
 * picture rounded up to 8 pixels in both directions (the part of the allocation reconstruction writes
 * and CDEF reads), so the host planes must be at least that large (the reference's are 128-aligned). */
int rb200_frame_upload_planes(Rb200Frame *f, int which, const void *const data[3], const ptrdiff_t stride[2]);
/* Film grain on output (rav1d_apply_grain, src/lib.rs:604 -> src/fg_apply.rs:272): parameters of the
 * RB200_STAGE_FILM_GRAIN stage.  The grained picture is a separate plane set (it is never used as a
 * reference); rb200_frame_readback returns it when the stage ran, rb200_frame_output_planes stays
 * the reference picture. */
int rb200_frame_set_film_grain(Rb200Frame *f, const Rb200FilmGrainData *data, int is_identity_matrix);
int rb200_frame_display_planes(Rb200Frame *f, Rb200Planes *out);
/* The device planes holding the result of the last submit (valid after rb200_frame_wait). */
int rb200_frame_output_planes(Rb200Frame *f, Rb200Planes *out);
int rb200_frame_stage_planes(Rb200Frame *f, int which, Rb200Planes *out);
/* Launch: H2D of `n_coefs` coefficients, the item lists and the filter metadata, then one
 * kernel sequence per stage on the frame's stream.  Asynchronous.
 * upload: RB200_UPLOAD_NONE (batch already on the device), RB200_UPLOAD_ALL, or
 * RB200_UPLOAD_ZERO_COPY_COEF: everything but the coefficients is copied; the inverse-transform
 * kernels read the coefficients directly from the pinned staging buffer, column-bounded by
 * Rb200ItxItem.ncols, so only the non-zero part of each block crosses PCIe;
 * RB200_UPLOAD_GATHER_COEF: the same bytes cross PCIe, but a gather kernel pulls each block's leading
 * columns into the device mirror with wide loads first and the transforms read device memory;
 * RB200_UPLOAD_GATHER_COEF16 (16-bit pictures): like RB200_UPLOAD_GATHER_COEF, but the coefficients cross PCIe as
 * int16 from rb200_frame_coef16_buffer() -- half the bytes; the few that do not fit (AV1 allows 18 + bits at 10 / 12 bpc)
 * travel as {index, value} records in rb200_frame_coef_escapes() and are patched in on the device;
 * RB200_UPLOAD_PACKED_COEF16 (16-bit pictures): the coefficients travel as ONE contiguous int16 stream -- each block's
 * leading ncols columns back to back (rb200_frame_coef_stream(), block i at rb200_frame_coef_stream_offsets()[i],
 * offsets multiples of 8) -- so a plain DMA copy moves exactly the bytes that matter at the link's full rate; a kernel
 * then spreads the blocks into the int32 device array.  Escapes as for RB200_UPLOAD_GATHER_COEF16. */
enum { RB200_UPLOAD_NONE = 0, RB200_UPLOAD_ALL = 1, RB200_UPLOAD_ZERO_COPY_COEF = 2, RB200_UPLOAD_GATHER_COEF = 3,
       RB200_UPLOAD_GATHER_COEF16 = 4, RB200_UPLOAD_PACKED_COEF16 = 5 };
/* int16 transport of the coefficients of a 16-bit picture.  The staging has the indexing of rb200_frame_coef_buffer()
 * (element i of one is element i of the other).  A front end writes it directly -- decode_coefs stores `dq as i16` and
 * pushes an escape when dq does not fit (src/recon.rs:1417: |dq| < 128 << bitdepth) -- or lets
 * rb200_frame_pack_coef16 narrow the first n_coefs elements of the int32 staging (it sizes the escape list itself). */
typedef struct Rb200CoefEscape { uint32_t index; int32_t value; } Rb200CoefEscape;
int16_t *rb200_frame_coef16_buffer(Rb200Frame *f);           /* [max_coefs], pinned, allocated on first use; NULL: 8-bit picture */
int rb200_frame_reserve_coef_escapes(Rb200Frame *f, int max_escapes);
Rb200CoefEscape *rb200_frame_coef_escapes(Rb200Frame *f);
int rb200_frame_set_coef_escape_count(Rb200Frame *f, int n);
int rb200_frame_pack_coef16(Rb200Frame *f, size_t n_coefs);
/* The packed stream of RB200_UPLOAD_PACKED_COEF16.  A front end appends to it directly (decode_coefs knows a block's
 * last non-zero column when it is done with the block); rb200_frame_pack_coef_stream builds it -- stream, offsets,
 * escapes, length -- from the int32 staging and the staged residual items (all of them: itx_counts plus the intra
 * levels' residuals when stages has RB200_STAGE_INTRA). */
int16_t *rb200_frame_coef_stream(Rb200Frame *f);             /* [max_coefs + 8 * max_itx_items], pinned, allocated on first use */
uint32_t *rb200_frame_coef_stream_offsets(Rb200Frame *f);    /* [max_itx_items], parallel to rb200_frame_itx_items() */
int rb200_frame_set_coef_stream_length(Rb200Frame *f, size_t n_elements);
int rb200_frame_pack_coef_stream(Rb200Frame *f, size_t n_coefs, const int32_t itx_counts[RB200_N_RECT_TX_SIZES], int stages);
int rb200_frame_submit(Rb200Frame *f, size_t n_coefs, const int32_t itx_counts[RB200_N_RECT_TX_SIZES],
                       int n_mc_items, int stages, int upload);
int rb200_frame_wait(Rb200Frame *f);
/* Checks the staged batch the way rb200_frame_submit will read it -- reference slots that are set, blocks inside the
 * picture allocation, coefficient offsets inside the buffer, transform sizes / types that exist, intra items' residual
 * indices -- and names the first bad record in rb200_last_error().  The kernels trust the records (a bad one is an
 * out-of-bounds access on the device, i.e. a sticky CUDA error); a front end calls this in debug builds or on
 * untrusted bitstreams.  Arguments as for rb200_frame_submit; costs one pass over the host-side lists. */
int rb200_frame_validate(Rb200Frame *f, size_t n_coefs, const int32_t itx_counts[RB200_N_RECT_TX_SIZES], int n_mc_items, int stages);
/* D2H of the output picture into host planes (stride in bytes, may be negative). */
int rb200_frame_readback(Rb200Frame *f, void *const data[3], const ptrdiff_t stride[2]);
/* Same, queued on the frame's stream without waiting (use pinned host planes, then rb200_frame_wait). */
int rb200_frame_readback_async(Rb200Frame *f, void *const data[3], const ptrdiff_t stride[2]);
/* ---- one picture split over several GPUs by superblock rows (post-filters) ----
 * A context can be restricted to a band named by the 64-row loop-restoration stripes
 * [stripe_begin, stripe_end) it must deliver (stripe s = luma rows 64 s - 8 .. 64 s + 55; (0, 0) = the
 * whole picture).  rb200_frame_band_rows reports the reconstructed luma rows the band reads (its own
 * rows plus the halo the earlier stages need) and the rows it delivers.  The halo comes from the
 * neighbour GPU over NVLink: rb200_frame_pull_rows copies rows of a plane set from a peer copy of the
 * same picture, addressed by the base of its memory block (rb200_frame_plane_block, mapped into this
 * process with rb200_ipc_open_handle when the peer is another process). */
int rb200_frame_set_band(Rb200Frame *f, int stripe_begin, int stripe_end);
int rb200_frame_band_rows(const Rb200Frame *f, int *in_row_begin, int *in_row_end, int *out_row_begin,
                          int *out_row_end);
int rb200_frame_upload_rows(Rb200Frame *f, int which, const void *const data[3], const ptrdiff_t stride[2],
                            int row_begin, int row_end);
int rb200_frame_readback_rows(Rb200Frame *f, void *const data[3], const ptrdiff_t stride[2], int row_begin,
                              int row_end);
int rb200_frame_plane_block(Rb200Frame *f, int which, void **base, size_t *bytes);
int rb200_frame_pull_rows(Rb200Frame *f, int which, const void *peer_base, int row_begin, int row_end);
#define RB200_IPC_HANDLE_BYTES 64
int rb200_ipc_get_handle(void *dptr, uint8_t handle[RB200_IPC_HANDLE_BYTES]);
int rb200_ipc_open_handle(const uint8_t handle[RB200_IPC_HANDLE_BYTES], void **dptr);
int rb200_ipc_close_handle(void *dptr);
int rb200_enable_peer_access(int peer_device);
/* Cross-GPU ordering without the host and without a collective (the band split's "my neighbours have pulled their halo
 * rows out of my picture, I may deblock it in place"): `flag` is a uint32 in the WAITING GPU's memory (rb200_malloc +
 * rb200_memset; shared like the planes with rb200_ipc_get_handle / _open_handle).  rb200_flag_signal writes `value`
 * into it once everything queued on `stream` so far is done -- through the peer mapping when the stream belongs to
 * another GPU; rb200_flag_wait holds `stream` until the flag has reached `value` (wrap-safe >=). */
int rb200_flag_signal(void *stream, uint32_t *flag, uint32_t value);
int rb200_flag_wait(void *stream, const uint32_t *flag, uint32_t value);
void *rb200_frame_stream(Rb200Frame *f);
/* Run the frame's copies and launches on `stream` instead of the frame's own (NULL restores it). */
int rb200_frame_set_stream(Rb200Frame *f, void *stream);
/* Stream-ordered dependency between frame contexts that run on different streams: the kernels of `f`'s NEXT
 * rb200_frame_submit wait on the device for the last kernel of `producer`'s most recent rb200_frame_submit (not for a
 * read-back queued behind it); the submit's own batch uploads are not held back.  `producer` must stay alive until
 * that submit.  This is what the reference's per-picture progress waits become
 * (src/thread_task.rs:1255-1323, rav1d_thread_picture_wait): a frame that predicts from `producer`'s output names it
 * here before its submit; no host synchronisation.  Up to 8 producers per submit; a no-op when both contexts share a
 * stream. */
int rb200_frame_depend(Rb200Frame *f, Rb200Frame *producer);
/* The post-filters of a frame run as a luma chain and a chroma chain on two streams (they only meet at the CDEF
 * direction search), joined before the submit ends; on by default.  0 = every launch on the frame's one stream, which
 * is what the per-stage times below need to mean anything. */
int rb200_frame_set_plane_streams(Rb200Frame *f, int on);
/* Optional promise about the order of the staged lists: the put items [0, n_mc_luma) are plane 0 and the rest chroma, and
 * every transform-size bucket t of the residual items holds its itx_luma_counts[t] plane-0 items first.  The
 * reconstruction of a frame that has nothing but put predictions and residuals (no compound / warped / OBMC / scaled /
 * intra items: those work on several planes at once) then runs as a luma chain and a chroma chain on the two streams as
 * well, and the chains continue into the post-filters without meeting.  Holds for every following submit;
 * n_mc_luma < 0 withdraws it. */
int rb200_frame_set_plane_counts(Rb200Frame *f, int n_mc_luma, const int32_t itx_luma_counts[RB200_N_RECT_TX_SIZES]);
/* Per-stage device times of the last submit (CUDA events on the frame's stream), the analogue
 * of the reference CLI's --frametimes (tools/dav1d.rs:127-150).  ms[] = H2D, MC, itx, deblock,
 * CDEF, LR, film grain; valid after rb200_frame_wait(). */
#define RB200_N_FRAME_MARKS 8
int rb200_frame_enable_timing(Rb200Frame *f, int on);
int rb200_frame_stage_times(Rb200Frame *f, float ms[RB200_N_FRAME_MARKS - 1]);
/* ---- loop-filter masks and levels generated on the device ----
 * rav1d_create_lf_mask_intra / rav1d_create_lf_mask_inter (src/lf_mask.rs:380-606; C: src/lf_mask.c:286-406)
 * run once per coded block inside decode_b (src/decode.rs; C: src/decode.c:1260-1271,1926-1947) and carry the
 * above / left transform-size contexts from block to block.  Here the front end appends one 16-byte record per
 * block -- exactly the arguments of those two calls, plus b->skip for the noskip mask (src/decode.c:1996-2005) --
 * and the submit builds Av1Filter.filter_y / filter_uv / noskip_mask and lf.level[][4] on the device in two
 * launches: a scatter of per-4x4 cell facts, then a gather that assembles every mask word without atomics.
 * The result is the mask set the filters consume, i.e. AFTER the tile-edge fix-ups of
 * src/lf_apply.rs (C: src/lf_apply_tmpl.c:331-400): a block edge is min(own, neighbour's) transform size whether
 * or not a tile boundary lies between them.  Records may come in any order.  cdef_idx is still taken from
 * rb200_frame_lf_masks(); with records set, that array's masks and rb200_frame_lf_levels() are not read. */
enum { RB200_LFB_INTRA = 1 /* create_lf_mask_intra: ytx = b->tx, no split, inner edges even if skipped */,
       RB200_LFB_SKIP = 2 /* b->skip */, RB200_LFB_HAS_CHROMA = 4 /* auv / luv passed (has_chroma) */ };
typedef struct Rb200LfBlock {
    uint16_t bx, by;        /* t->bx, t->by: block origin in 4-pixel luma units */
    uint8_t bs;             /* enum BlockSize, src/levels.rs (BS_128x128 = 0 .. BS_4x4 = 21) */
    uint8_t flags;          /* RB200_LFB_* */
    uint8_t ytx, uvtx;      /* RB200_TX_* : b->tx (intra) or max_ytx (inter); uvtx */
    uint16_t tx_split[2];   /* b->tx_split0, b->tx_split1 (inter) */
    uint8_t lvl[4];         /* filter_level[0..3][0][0]: y column edges, y row edges, u, v */
} RB200_ALIGN16 Rb200LfBlock;           /* 16 bytes */
int rb200_frame_reserve_lf_blocks(Rb200Frame *f, int max_blocks);
Rb200LfBlock *rb200_frame_lf_blocks(Rb200Frame *f);
int rb200_frame_set_lf_block_count(Rb200Frame *f, int n_blocks);   /* 0 = masks and levels come from the host arrays */
/* Test / debugging aid: copies the device's mask and level arrays of the last submit back (waits for the stream). */
int rb200_frame_download_lf(Rb200Frame *f, Rb200Av1Filter *masks, uint8_t (*levels)[4]);
/* Number of kernels launched by the last submit (bench bookkeeping). */
int rb200_frame_last_launches(const Rb200Frame *f);

#ifdef __cplusplus
}
#endif
#endif /* RAV1D_B200_H */
