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

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ core */
#define RB200_ABI_VERSION 1
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
    uint8_t flags;    /* reserved, 0 */
    int16_t eob;
    int16_t pad;
} Rb200ItxItem;       /* 16 bytes */
/* d_items (device) sorted by tx size; counts[t] = number of items of size t.
 * Coefficients are read, not zeroed: the host zeroes its own staging copy. */
int rb200_itx_add_batch(const Rb200Planes *planes, const void *d_coef, const Rb200ItxItem *d_items,
                        const int32_t counts[RB200_N_RECT_TX_SIZES], int bitdepth_max, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RAV1D_B200_H */
