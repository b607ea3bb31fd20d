/*
 * oracle/ref_harness.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Thin C glue that is linked INTO oracle/_ref/libdav1d_ref.so (the reference's
 * own portable C DSP compiled in place from /root/reference, see Makefile).
 * It gives the Python tests and bench.py's cpu_baseline leg index-based access
 * to the reference's DSP function-pointer tables and to its (hidden-visibility)
 * constant tables.  Nothing here is part of the product; the product library
 * (rav1d_b200/csrc) never links or loads it.
 *
 * The DSP tables are the C twins of the Rust ones:
 *   Dav1dInvTxfmDSPContext          src/itx.h:42-44   (Rust: src/itx.rs:193-196)
 *   Dav1dMCDSPContext               src/mc.h:114-130  (Rust: src/mc.rs:1321-1338)
 *   Dav1dLoopFilterDSPContext       src/loopfilter.h:45-53
 *   Dav1dCdefDSPContext             src/cdef.h:64-67
 *   Dav1dLoopRestorationDSPContext  src/looprestoration.h:72-75
 *   Dav1dFilmGrainDSPContext        src/filmgrain.h:74-80
 *   Dav1dIntraPredDSPContext        src/ipred.h:82-93  (Rust: src/ipred.rs:164-169)
 */
#include "config.h"

#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "src/levels.h"
#include "src/tables.h"
#include "src/lf_mask.h"
#include "src/wedge.h"
#include "dav1d/headers.h"

/* ---- table sizes (enum-sized arrays, bit-depth independent layout) ---- */
typedef void (*fnptr)(void);

typedef struct { fnptr itxfm_add[N_RECT_TX_SIZES][N_TX_TYPES_PLUS_LL]; } ItxCtx;
typedef struct {
    fnptr mc[N_2D_FILTERS], mc_scaled[N_2D_FILTERS], mct[N_2D_FILTERS], mct_scaled[N_2D_FILTERS];
    fnptr avg, w_avg, mask, w_mask[3], blend, blend_v, blend_h, warp8x8, warp8x8t, emu_edge, resize;
} McCtx;
typedef struct { fnptr loop_filter_sb[2][2]; } LfCtx;
typedef struct { fnptr dir; fnptr fb[3]; } CdefCtx;
typedef struct { fnptr wiener[2]; fnptr sgr[3]; } LrCtx;
typedef struct { fnptr generate_grain_y; fnptr generate_grain_uv[3]; fnptr fgy_32x32xn; fnptr fguv_32x32xn[3]; } FgCtx;
typedef struct { fnptr intra_pred[N_IMPL_INTRA_PRED_MODES]; fnptr cfl_ac[3]; fnptr cfl_pred[DC_128_PRED + 1]; fnptr pal_pred; } IpredCtx;

void dav1d_itx_dsp_init_8bpc(ItxCtx *c, int bpc);
void dav1d_itx_dsp_init_16bpc(ItxCtx *c, int bpc);
void dav1d_mc_dsp_init_8bpc(McCtx *c);
void dav1d_mc_dsp_init_16bpc(McCtx *c);
void dav1d_loop_filter_dsp_init_8bpc(LfCtx *c);
void dav1d_loop_filter_dsp_init_16bpc(LfCtx *c);
void dav1d_cdef_dsp_init_8bpc(CdefCtx *c);
void dav1d_cdef_dsp_init_16bpc(CdefCtx *c);
void dav1d_loop_restoration_dsp_init_8bpc(LrCtx *c, int bpc);
void dav1d_loop_restoration_dsp_init_16bpc(LrCtx *c, int bpc);
void dav1d_film_grain_dsp_init_8bpc(FgCtx *c);
void dav1d_film_grain_dsp_init_16bpc(FgCtx *c);
void dav1d_intra_pred_dsp_init_8bpc(IpredCtx *c);
void dav1d_intra_pred_dsp_init_16bpc(IpredCtx *c);

static struct Tables {
    int ready;
    ItxCtx itx[3]; McCtx mc[2]; LfCtx lf[2]; CdefCtx cdef[2]; LrCtx lr[3]; FgCtx fg[2]; IpredCtx ipred[2];
} T;

static void init_once(void) {
    if (T.ready) return;
    dav1d_itx_dsp_init_8bpc(&T.itx[0], 8);
    dav1d_itx_dsp_init_16bpc(&T.itx[1], 10);
    dav1d_itx_dsp_init_16bpc(&T.itx[2], 12);
    dav1d_mc_dsp_init_8bpc(&T.mc[0]);
    dav1d_mc_dsp_init_16bpc(&T.mc[1]);
    dav1d_loop_filter_dsp_init_8bpc(&T.lf[0]);
    dav1d_loop_filter_dsp_init_16bpc(&T.lf[1]);
    dav1d_cdef_dsp_init_8bpc(&T.cdef[0]);
    dav1d_cdef_dsp_init_16bpc(&T.cdef[1]);
    dav1d_loop_restoration_dsp_init_8bpc(&T.lr[0], 8);
    dav1d_loop_restoration_dsp_init_16bpc(&T.lr[1], 10);
    dav1d_loop_restoration_dsp_init_16bpc(&T.lr[2], 12);
    dav1d_film_grain_dsp_init_8bpc(&T.fg[0]);
    dav1d_film_grain_dsp_init_16bpc(&T.fg[1]);
    dav1d_intra_pred_dsp_init_8bpc(&T.ipred[0]);
    dav1d_intra_pred_dsp_init_16bpc(&T.ipred[1]);
    dav1d_init_wedge_masks();
    T.ready = 1;
}

void ref_init(void) { init_once(); }

static inline int hbd(int bdmax) { return bdmax > 255; }
static inline int bdidx(int bdmax) { return bdmax > 255 ? (bdmax > 1023 ? 2 : 1) : 0; }

/* ------------------------------- itx --------------------------------- */
int ref_itx_has(int tx, int txtp) { init_once(); return T.itx[0].itxfm_add[tx][txtp] != NULL; }

void ref_itxfm_add(int tx, int txtp, void *dst, ptrdiff_t stride, void *coeff, int eob, int bdmax) {
    init_once();
    if (!hbd(bdmax))
        ((void (*)(void *, ptrdiff_t, void *, int))T.itx[0].itxfm_add[tx][txtp])(dst, stride, coeff, eob);
    else
        ((void (*)(void *, ptrdiff_t, void *, int, int))T.itx[bdidx(bdmax)].itxfm_add[tx][txtp])(dst, stride, coeff, eob, bdmax);
}

/* many blocks in one call (used for timing the CPU baseline and bulk parity) */
void ref_itxfm_add_many(int n, const int32_t *items /* n x 6: tx,txtp,x,y,eob,cf_off */,
                        void *plane, ptrdiff_t stride, void *cf, int bdmax) {
    init_once();
    const int px = hbd(bdmax) ? 2 : 1, cs = hbd(bdmax) ? 4 : 2;
    for (int i = 0; i < n; i++) {
        const int32_t *it = items + 6 * i;
        uint8_t *d = (uint8_t *)plane + (ptrdiff_t)it[3] * stride + (ptrdiff_t)it[2] * px;
        ref_itxfm_add(it[0], it[1], d, stride, (uint8_t *)cf + (size_t)(uint32_t)it[5] * cs, it[4], bdmax);
    }
}

/* -------------------------------- mc ---------------------------------- */
void ref_mc(int filt, void *dst, ptrdiff_t ds, const void *src, ptrdiff_t ss, int w, int h, int mx, int my, int bdmax) {
    init_once();
    if (!hbd(bdmax)) ((void (*)(void *, ptrdiff_t, const void *, ptrdiff_t, int, int, int, int))T.mc[0].mc[filt])(dst, ds, src, ss, w, h, mx, my);
    else ((void (*)(void *, ptrdiff_t, const void *, ptrdiff_t, int, int, int, int, int))T.mc[1].mc[filt])(dst, ds, src, ss, w, h, mx, my, bdmax);
}
void ref_mct(int filt, int16_t *tmp, const void *src, ptrdiff_t ss, int w, int h, int mx, int my, int bdmax) {
    init_once();
    if (!hbd(bdmax)) ((void (*)(int16_t *, const void *, ptrdiff_t, int, int, int, int))T.mc[0].mct[filt])(tmp, src, ss, w, h, mx, my);
    else ((void (*)(int16_t *, const void *, ptrdiff_t, int, int, int, int, int))T.mc[1].mct[filt])(tmp, src, ss, w, h, mx, my, bdmax);
}
void ref_mc_scaled(int filt, void *dst, ptrdiff_t ds, const void *src, ptrdiff_t ss, int w, int h, int mx, int my, int dx, int dy, int bdmax) {
    init_once();
    if (!hbd(bdmax)) ((void (*)(void *, ptrdiff_t, const void *, ptrdiff_t, int, int, int, int, int, int))T.mc[0].mc_scaled[filt])(dst, ds, src, ss, w, h, mx, my, dx, dy);
    else ((void (*)(void *, ptrdiff_t, const void *, ptrdiff_t, int, int, int, int, int, int, int))T.mc[1].mc_scaled[filt])(dst, ds, src, ss, w, h, mx, my, dx, dy, bdmax);
}
void ref_mct_scaled(int filt, int16_t *tmp, const void *src, ptrdiff_t ss, int w, int h, int mx, int my, int dx, int dy, int bdmax) {
    init_once();
    if (!hbd(bdmax)) ((void (*)(int16_t *, const void *, ptrdiff_t, int, int, int, int, int, int))T.mc[0].mct_scaled[filt])(tmp, src, ss, w, h, mx, my, dx, dy);
    else ((void (*)(int16_t *, const void *, ptrdiff_t, int, int, int, int, int, int, int))T.mc[1].mct_scaled[filt])(tmp, src, ss, w, h, mx, my, dx, dy, bdmax);
}
void ref_avg(void *dst, ptrdiff_t ds, const int16_t *t1, const int16_t *t2, int w, int h, int bdmax) {
    init_once();
    if (!hbd(bdmax)) ((void (*)(void *, ptrdiff_t, const int16_t *, const int16_t *, int, int))T.mc[0].avg)(dst, ds, t1, t2, w, h);
    else ((void (*)(void *, ptrdiff_t, const int16_t *, const int16_t *, int, int, int))T.mc[1].avg)(dst, ds, t1, t2, w, h, bdmax);
}
void ref_w_avg(void *dst, ptrdiff_t ds, const int16_t *t1, const int16_t *t2, int w, int h, int weight, int bdmax) {
    init_once();
    if (!hbd(bdmax)) ((void (*)(void *, ptrdiff_t, const int16_t *, const int16_t *, int, int, int))T.mc[0].w_avg)(dst, ds, t1, t2, w, h, weight);
    else ((void (*)(void *, ptrdiff_t, const int16_t *, const int16_t *, int, int, int, int))T.mc[1].w_avg)(dst, ds, t1, t2, w, h, weight, bdmax);
}
void ref_mask(void *dst, ptrdiff_t ds, const int16_t *t1, const int16_t *t2, int w, int h, const uint8_t *m, int bdmax) {
    init_once();
    if (!hbd(bdmax)) ((void (*)(void *, ptrdiff_t, const int16_t *, const int16_t *, int, int, const uint8_t *))T.mc[0].mask)(dst, ds, t1, t2, w, h, m);
    else ((void (*)(void *, ptrdiff_t, const int16_t *, const int16_t *, int, int, const uint8_t *, int))T.mc[1].mask)(dst, ds, t1, t2, w, h, m, bdmax);
}
void ref_w_mask(int ss, void *dst, ptrdiff_t ds, const int16_t *t1, const int16_t *t2, int w, int h, uint8_t *m, int sign, int bdmax) {
    init_once();
    if (!hbd(bdmax)) ((void (*)(void *, ptrdiff_t, const int16_t *, const int16_t *, int, int, uint8_t *, int))T.mc[0].w_mask[ss])(dst, ds, t1, t2, w, h, m, sign);
    else ((void (*)(void *, ptrdiff_t, const int16_t *, const int16_t *, int, int, uint8_t *, int, int))T.mc[1].w_mask[ss])(dst, ds, t1, t2, w, h, m, sign, bdmax);
}
void ref_blend(int dir /*0 mask,1 v,2 h*/, void *dst, ptrdiff_t ds, const void *tmp, int w, int h, const uint8_t *m, int bdmax) {
    init_once();
    McCtx *c = &T.mc[hbd(bdmax)];
    if (dir == 0) ((void (*)(void *, ptrdiff_t, const void *, int, int, const uint8_t *))c->blend)(dst, ds, tmp, w, h, m);
    else ((void (*)(void *, ptrdiff_t, const void *, int, int))(dir == 1 ? c->blend_v : c->blend_h))(dst, ds, tmp, w, h);
}
void ref_warp8x8(void *dst, ptrdiff_t ds, const void *src, ptrdiff_t ss, const int16_t *abcd, int mx, int my, int bdmax) {
    init_once();
    if (!hbd(bdmax)) ((void (*)(void *, ptrdiff_t, const void *, ptrdiff_t, const int16_t *, int, int))T.mc[0].warp8x8)(dst, ds, src, ss, abcd, mx, my);
    else ((void (*)(void *, ptrdiff_t, const void *, ptrdiff_t, const int16_t *, int, int, int))T.mc[1].warp8x8)(dst, ds, src, ss, abcd, mx, my, bdmax);
}
void ref_warp8x8t(int16_t *tmp, ptrdiff_t ts, const void *src, ptrdiff_t ss, const int16_t *abcd, int mx, int my, int bdmax) {
    init_once();
    if (!hbd(bdmax)) ((void (*)(int16_t *, ptrdiff_t, const void *, ptrdiff_t, const int16_t *, int, int))T.mc[0].warp8x8t)(tmp, ts, src, ss, abcd, mx, my);
    else ((void (*)(int16_t *, ptrdiff_t, const void *, ptrdiff_t, const int16_t *, int, int, int))T.mc[1].warp8x8t)(tmp, ts, src, ss, abcd, mx, my, bdmax);
}
void ref_emu_edge(intptr_t bw, intptr_t bh, intptr_t iw, intptr_t ih, intptr_t x, intptr_t y,
                  void *dst, ptrdiff_t ds, const void *src, ptrdiff_t ss, int bdmax) {
    init_once();
    ((void (*)(intptr_t, intptr_t, intptr_t, intptr_t, intptr_t, intptr_t, void *, ptrdiff_t, const void *, ptrdiff_t))
         T.mc[hbd(bdmax)].emu_edge)(bw, bh, iw, ih, x, y, dst, ds, src, ss);
}
void ref_resize(void *dst, ptrdiff_t ds, const void *src, ptrdiff_t ss, int dst_w, int h, int src_w, int dx, int mx, int bdmax) {
    init_once();
    if (!hbd(bdmax)) ((void (*)(void *, ptrdiff_t, const void *, ptrdiff_t, int, int, int, int, int))T.mc[0].resize)(dst, ds, src, ss, dst_w, h, src_w, dx, mx);
    else ((void (*)(void *, ptrdiff_t, const void *, ptrdiff_t, int, int, int, int, int, int))T.mc[1].resize)(dst, ds, src, ss, dst_w, h, src_w, dx, mx, bdmax);
}

/* ----------------------------- loopfilter ------------------------------ */
void ref_lpf_sb(int pl_uv, int dir /*0: col edges (h), 1: row edges (v)*/, void *dst, ptrdiff_t stride,
                const uint32_t *mask, const uint8_t (*lvl)[4], ptrdiff_t lvl_stride,
                const Av1FilterLUT *lut, int wh, int bdmax) {
    init_once();
    if (!hbd(bdmax)) ((void (*)(void *, ptrdiff_t, const uint32_t *, const uint8_t (*)[4], ptrdiff_t, const Av1FilterLUT *, int))
                          T.lf[0].loop_filter_sb[pl_uv][dir])(dst, stride, mask, lvl, lvl_stride, lut, wh);
    else ((void (*)(void *, ptrdiff_t, const uint32_t *, const uint8_t (*)[4], ptrdiff_t, const Av1FilterLUT *, int, int))
              T.lf[1].loop_filter_sb[pl_uv][dir])(dst, stride, mask, lvl, lvl_stride, lut, wh, bdmax);
}

/* -------------------------------- cdef --------------------------------- */
int ref_cdef_dir(const void *src, ptrdiff_t stride, unsigned *var, int bdmax) {
    init_once();
    if (!hbd(bdmax)) return ((int (*)(const void *, ptrdiff_t, unsigned *))T.cdef[0].dir)(src, stride, var);
    return ((int (*)(const void *, ptrdiff_t, unsigned *, int))T.cdef[1].dir)(src, stride, var, bdmax);
}
void ref_cdef_fb(int idx /*0: 8x8, 1: 4x8, 2: 4x4*/, void *dst, ptrdiff_t stride, const void *left,
                 const void *top, const void *bottom, int pri, int sec, int dir, int damping, int edges, int bdmax) {
    init_once();
    if (!hbd(bdmax)) ((void (*)(void *, ptrdiff_t, const void *, const void *, const void *, int, int, int, int, int))
                          T.cdef[0].fb[idx])(dst, stride, left, top, bottom, pri, sec, dir, damping, edges);
    else ((void (*)(void *, ptrdiff_t, const void *, const void *, const void *, int, int, int, int, int, int))
              T.cdef[1].fb[idx])(dst, stride, left, top, bottom, pri, sec, dir, damping, edges, bdmax);
}

/* ------------------------- loop restoration ---------------------------- */
void ref_lr(int kind /*0 wiener7,1 wiener5,2 sgr5x5,3 sgr3x3,4 sgrmix*/, void *dst, ptrdiff_t stride,
            const void *left, const void *lpf, int w, int h, const void *params, int edges, int bdmax) {
    init_once();
    LrCtx *c = &T.lr[bdidx(bdmax)];
    fnptr f = kind < 2 ? c->wiener[kind] : c->sgr[kind - 2];
    if (!hbd(bdmax)) ((void (*)(void *, ptrdiff_t, const void *, const void *, int, int, const void *, int))f)(dst, stride, left, lpf, w, h, params, edges);
    else ((void (*)(void *, ptrdiff_t, const void *, const void *, int, int, const void *, int, int))f)(dst, stride, left, lpf, w, h, params, edges, bdmax);
}

/* ------------------------------ film grain ----------------------------- */
void ref_fg_gen_y(void *buf, const Dav1dFilmGrainData *d, int bdmax) {
    init_once();
    if (!hbd(bdmax)) ((void (*)(void *, const Dav1dFilmGrainData *))T.fg[0].generate_grain_y)(buf, d);
    else ((void (*)(void *, const Dav1dFilmGrainData *, int))T.fg[1].generate_grain_y)(buf, d, bdmax);
}
void ref_fg_gen_uv(int ss /* layout - 1: 0 = 4:2:0, 1 = 4:2:2, 2 = 4:4:4 */, void *buf, const void *buf_y, const Dav1dFilmGrainData *d, intptr_t uv, int bdmax) {
    init_once();
    if (!hbd(bdmax)) ((void (*)(void *, const void *, const Dav1dFilmGrainData *, intptr_t))T.fg[0].generate_grain_uv[ss])(buf, buf_y, d, uv);
    else ((void (*)(void *, const void *, const Dav1dFilmGrainData *, intptr_t, int))T.fg[1].generate_grain_uv[ss])(buf, buf_y, d, uv, bdmax);
}
void ref_fgy(void *dst, const void *src, ptrdiff_t stride, const Dav1dFilmGrainData *d, size_t pw,
             const uint8_t *scaling, const void *grain_lut, int bh, int row_num, int bdmax) {
    init_once();
    if (!hbd(bdmax)) ((void (*)(void *, const void *, ptrdiff_t, const Dav1dFilmGrainData *, size_t, const uint8_t *, const void *, int, int))
                          T.fg[0].fgy_32x32xn)(dst, src, stride, d, pw, scaling, grain_lut, bh, row_num);
    else ((void (*)(void *, const void *, ptrdiff_t, const Dav1dFilmGrainData *, size_t, const uint8_t *, const void *, int, int, int))
              T.fg[1].fgy_32x32xn)(dst, src, stride, d, pw, scaling, grain_lut, bh, row_num, bdmax);
}
void ref_fguv(int ss, void *dst, const void *src, ptrdiff_t stride, const Dav1dFilmGrainData *d, size_t pw,
              const uint8_t *scaling, const void *grain_lut, int bh, int row_num,
              const void *luma, ptrdiff_t luma_stride, int uv_pl, int is_id, int bdmax) {
    init_once();
    if (!hbd(bdmax)) ((void (*)(void *, const void *, ptrdiff_t, const Dav1dFilmGrainData *, size_t, const uint8_t *, const void *, int, int, const void *, ptrdiff_t, int, int))
                          T.fg[0].fguv_32x32xn[ss])(dst, src, stride, d, pw, scaling, grain_lut, bh, row_num, luma, luma_stride, uv_pl, is_id);
    else ((void (*)(void *, const void *, ptrdiff_t, const Dav1dFilmGrainData *, size_t, const uint8_t *, const void *, int, int, const void *, ptrdiff_t, int, int, int))
              T.fg[1].fguv_32x32xn[ss])(dst, src, stride, d, pw, scaling, grain_lut, bh, row_num, luma, luma_stride, uv_pl, is_id, bdmax);
}
/* ------------------------------ intra prediction ------------------------ */
/* mode: enum IntraPredMode incl. the implementation modes (src/levels.h:85-107): DC 0, VERT 1, HOR 2, ..., Z1 = 6 .. */
void ref_ipred(int mode, void *dst, ptrdiff_t stride, const void *topleft, int w, int h, int angle, int max_w, int max_h, int bdmax) {
    init_once();
    if (bdmax > 255) ((void (*)(void *, ptrdiff_t, const void *, int, int, int, int, int, int))T.ipred[1].intra_pred[mode])(dst, stride, topleft, w, h, angle, max_w, max_h, bdmax);
    else ((void (*)(void *, ptrdiff_t, const void *, int, int, int, int, int))T.ipred[0].intra_pred[mode])(dst, stride, topleft, w, h, angle, max_w, max_h);
}
void ref_cfl_ac(int ss /* layout - 1 */, int16_t *ac, const void *y, ptrdiff_t stride, int w_pad, int h_pad, int cw, int ch, int bdmax) {
    init_once();
    ((void (*)(int16_t *, const void *, ptrdiff_t, int, int, int, int))T.ipred[bdmax > 255].cfl_ac[ss])(ac, y, stride, w_pad, h_pad, cw, ch);
}
void ref_cfl_pred(int mode, void *dst, ptrdiff_t stride, const void *topleft, int w, int h, const int16_t *ac, int alpha, int bdmax) {
    init_once();
    if (bdmax > 255) ((void (*)(void *, ptrdiff_t, const void *, int, int, const int16_t *, int, int))T.ipred[1].cfl_pred[mode])(dst, stride, topleft, w, h, ac, alpha, bdmax);
    else ((void (*)(void *, ptrdiff_t, const void *, int, int, const int16_t *, int))T.ipred[0].cfl_pred[mode])(dst, stride, topleft, w, h, ac, alpha);
}
void ref_pal_pred(void *dst, ptrdiff_t stride, const void *pal, const uint8_t *idx, int w, int h, int bdmax) {
    init_once();
    ((void (*)(void *, ptrdiff_t, const void *, const uint8_t *, int, int))T.ipred[bdmax > 255].pal_pred)(dst, stride, pal, idx, w, h);
}
int ref_ipred_mode_ids(int *out /* DC, VERT, HOR, PAETH, SMOOTH, SMOOTH_V, SMOOTH_H, Z1, Z2, Z3, LEFT_DC, TOP_DC, DC_128, FILTER */) {
    const int ids[14] = { DC_PRED, VERT_PRED, HOR_PRED, PAETH_PRED, SMOOTH_PRED, SMOOTH_V_PRED, SMOOTH_H_PRED, Z1_PRED, Z2_PRED, Z3_PRED,
                          LEFT_DC_PRED, TOP_DC_PRED, DC_128_PRED, FILTER_PRED };
    for (int i = 0; i < 14; i++) out[i] = ids[i];
    return N_IMPL_INTRA_PRED_MODES;
}

size_t ref_sizeof_film_grain_data(void) { return sizeof(Dav1dFilmGrainData); }

/* ------------------------- constant tables ----------------------------- */
/* AV1-normative constant tables (hidden visibility in the library): exposed so
 * that tools/gen_tables.py can cross-check the product's generated tables. */
const void *ref_table(const char *name, size_t *size) {
#define TAB(sym) if (!strcmp(name, #sym)) { *size = sizeof(dav1d_##sym); return dav1d_##sym; }
    TAB(mc_subpel_filters) TAB(mc_warp_filter) TAB(resize_filter)
    TAB(sgr_params) TAB(sgr_x_by_x) TAB(cdef_directions)
    TAB(obmc_masks) TAB(gaussian_sequence) TAB(txfm_dimensions)
    TAB(block_dimensions) TAB(filter_2d) TAB(filter_dir)
    TAB(sm_weights) TAB(dr_intra_derivative) TAB(filter_intra_taps)
#undef TAB
    *size = 0;
    return NULL;
}
