/*
 * oracle/ref_dump.c -- TEST INFRASTRUCTURE ONLY (builds oracle/_ref/ref_dump).
 *
 * Decodes an AV1 stream with the REFERENCE decoder (libdav1d_ref.so, the reference's own C
 * sources compiled in place) and, for every decoded frame, writes what the post-filter stage
 * consumed and produced:
 *
 *   - the reconstructed picture BEFORE the in-loop filters,
 *   - the filter metadata exactly as the decoder built it (Av1Filter masks incl. the tile-edge
 *     fix-ups of src/lf_apply_tmpl.c:331-395, level[4], Av1FilterLUT, Av1Restoration units, header
 *     fields of cdef / loop restoration),
 *   - the picture AFTER deblock + CDEF + loop restoration.
 *
 * It does so by interposing dav1d_filter_sbrow_{8,16}bpc (src/recon_tmpl.c:2166-2175), the
 * function the single-threaded decoder calls once per superblock row after reconstruction
 * (src/decode.c:3247): this executable exports its own definition, which the dynamic linker
 * prefers over the library's when src/decode.c takes the function's address; the original is
 * reached through dlsym(RTLD_NEXT).  Rows of superblock row `sby` are still pristine when
 * filter_sbrow(f, sby) is entered, so copying them there assembles the pre-filter picture even
 * though the CPU filters in place with a lag.
 *
 * tests/test_streams.py replays the dump through the CUDA post-filter path and compares.
 *
 * usage: ref_dump <in.ivf|.obu> <out.bin> [max_frames [max_grain_frames [sr_only]]]
 */
#define _GNU_SOURCE
#include "config.h"

#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "dav1d/dav1d.h"
#include "src/internal.h"
#include "src/lf_mask.h"
#include "input/input.h"

#include "../include/rav1d_b200.h"

static FILE *g_out;
static int g_frames, g_max_frames = 1 << 30, g_sr_seen;
static int g_sr_mode;   /* 1: dump only super-resolution frames (record "RBSR": 8 more header ints, post = sr_cur) */
static uint8_t *g_pre[3];
static size_t g_pre_sz[3];

typedef void (*filter_sbrow_fn)(Dav1dFrameContext *f, int sby);

/* ---- the arguments of every dav1d_create_lf_mask_{intra,inter} call of the frame (src/decode.c:
 * 1260-1271,1926-1947), interposed like filter_sbrow (the library calls them through its PLT).  With
 * RB200_DUMP_LFB=<file> each dumped frame also appends "RBLB", frame index, n, n Rb200LfBlock records there. */
static FILE *g_lfb_out;
static Rb200LfBlock *g_lfb;
static int g_n_lfb, g_cap_lfb;
static void lfb_push(int bx, int by, int bs, int flags, int ytx, int uvtx, const uint16_t *tx_masks,
                     const uint8_t (*filter_level)[8][2]) {
    if (g_n_lfb == g_cap_lfb) { g_cap_lfb = g_cap_lfb ? 2 * g_cap_lfb : 4096; g_lfb = realloc(g_lfb, (size_t)g_cap_lfb * sizeof(*g_lfb)); }
    Rb200LfBlock *b = &g_lfb[g_n_lfb++];
    b->bx = bx; b->by = by; b->bs = bs; b->flags = flags; b->ytx = ytx; b->uvtx = uvtx;
    b->tx_split[0] = tx_masks ? tx_masks[0] : 0; b->tx_split[1] = tx_masks ? tx_masks[1] : 0;
    for (int k = 0; k < 4; k++) b->lvl[k] = filter_level[k][0][0];
}
typedef void (*lf_intra_fn)(Av1Filter *, uint8_t (*)[4], ptrdiff_t, const uint8_t (*)[8][2], int, int, int, int,
                            enum BlockSize, enum RectTxfmSize, enum RectTxfmSize, enum Dav1dPixelLayout, uint8_t *, uint8_t *,
                            uint8_t *, uint8_t *);
typedef void (*lf_inter_fn)(Av1Filter *, uint8_t (*)[4], ptrdiff_t, const uint8_t (*)[8][2], int, int, int, int, int,
                            enum BlockSize, enum RectTxfmSize, const uint16_t *, enum RectTxfmSize, enum Dav1dPixelLayout,
                            uint8_t *, uint8_t *, uint8_t *, uint8_t *);
__attribute__((visibility("default"))) void dav1d_create_lf_mask_intra(Av1Filter *lflvl, uint8_t (*level_cache)[4], ptrdiff_t b4_stride,
        const uint8_t (*filter_level)[8][2], int bx, int by, int iw, int ih, enum BlockSize bs, enum RectTxfmSize ytx,
        enum RectTxfmSize uvtx, enum Dav1dPixelLayout layout, uint8_t *ay, uint8_t *ly, uint8_t *auv, uint8_t *luv) {
    static lf_intra_fn real;
    if (!real) real = (lf_intra_fn)dlsym(RTLD_NEXT, "dav1d_create_lf_mask_intra");
    if (!real) { fprintf(stderr, "ref_dump: cannot find dav1d_create_lf_mask_intra\n"); exit(2); }
    real(lflvl, level_cache, b4_stride, filter_level, bx, by, iw, ih, bs, ytx, uvtx, layout, ay, ly, auv, luv);
    if (g_lfb_out) lfb_push(bx, by, bs, RB200_LFB_INTRA | (auv ? RB200_LFB_HAS_CHROMA : 0), ytx, uvtx, NULL, filter_level);
}
__attribute__((visibility("default"))) void dav1d_create_lf_mask_inter(Av1Filter *lflvl, uint8_t (*level_cache)[4], ptrdiff_t b4_stride,
        const uint8_t (*filter_level)[8][2], int bx, int by, int iw, int ih, int skip, enum BlockSize bs,
        enum RectTxfmSize max_ytx, const uint16_t *tx_masks, enum RectTxfmSize uvtx, enum Dav1dPixelLayout layout,
        uint8_t *ay, uint8_t *ly, uint8_t *auv, uint8_t *luv) {
    static lf_inter_fn real;
    if (!real) real = (lf_inter_fn)dlsym(RTLD_NEXT, "dav1d_create_lf_mask_inter");
    if (!real) { fprintf(stderr, "ref_dump: cannot find dav1d_create_lf_mask_inter\n"); exit(2); }
    real(lflvl, level_cache, b4_stride, filter_level, bx, by, iw, ih, skip, bs, max_ytx, tx_masks, uvtx, layout, ay, ly, auv, luv);
    if (g_lfb_out) lfb_push(bx, by, bs, (skip ? RB200_LFB_SKIP : 0) | (auv ? RB200_LFB_HAS_CHROMA : 0), max_ytx, uvtx, tx_masks, filter_level);
}

static void put_i32(int32_t v) { fwrite(&v, 4, 1, g_out); }

static void hook(Dav1dFrameContext *const f, const int sby, const char *const sym) {
    static filter_sbrow_fn real8, real16;
    filter_sbrow_fn *real = f->cur.p.bpc > 8 ? &real16 : &real8;
    if (!*real) *real = (filter_sbrow_fn)dlsym(RTLD_NEXT, sym);
    if (!*real) { fprintf(stderr, "ref_dump: cannot find %s\n", sym); exit(2); }
    const int hbd = f->cur.p.bpc > 8, px = hbd ? 2 : 1;
    const int layout = f->cur.p.layout;
    const int ss_ver = layout == DAV1D_PIXEL_LAYOUT_I420, ss_hor = layout != DAV1D_PIXEL_LAYOUT_I444;
    const int n_planes = layout == DAV1D_PIXEL_LAYOUT_I400 ? 1 : 3;
    const int ah = (f->cur.p.h + 127) & ~127;
    const int is_sr = f->frame_hdr->width[0] != f->frame_hdr->width[1];
    const int dump = g_frames < g_max_frames && f->cur.stride[0] > 0 && is_sr == g_sr_mode;
    if (sby == 0 && f->frame_hdr->width[0] != f->frame_hdr->width[1]) g_sr_seen++;
    if (dump) {
        /* copy the pristine rows of this superblock row (to the end of the allocation on the last one) */
        const int sbsz = f->sb_step * 4;
        const int y0 = sby * sbsz, y1 = sby + 1 == f->sbh ? ah : (sby + 1) * sbsz;
        for (int pl = 0; pl < n_planes; pl++) {
            const ptrdiff_t stride = f->cur.stride[!!pl];
            const int rows = pl ? ah >> ss_ver : ah;
            const size_t need = (size_t)stride * rows;
            if (g_pre_sz[pl] < need) { g_pre[pl] = realloc(g_pre[pl], need); g_pre_sz[pl] = need; }
            const int a = pl ? y0 >> ss_ver : y0, b = pl ? y1 >> ss_ver : y1;
            memcpy(g_pre[pl] + (size_t)a * stride, (const uint8_t *)f->cur.data[pl] + (size_t)a * stride, (size_t)(b - a) * stride);
        }
    }
    (*real)(f, sby);
    if (sby + 1 == f->sbh) {    /* the frame's blocks are complete */
        if (dump && g_lfb_out) {
            int32_t hd[3] = { 0x52424c42, g_frames, g_n_lfb };
            fwrite(hd, 4, 3, g_lfb_out);
            fwrite(g_lfb, sizeof(*g_lfb), g_n_lfb, g_lfb_out);
        }
        g_n_lfb = 0;
    }
    if (!dump || sby + 1 != f->sbh) return;

    /* ---- the frame is complete: header, metadata, pre and post pictures */
    const Dav1dFrameHeader *const h = f->frame_hdr;
    const int n_sb128 = f->sb128w * f->sb128h;
    put_i32(is_sr ? 0x52425352 : 0x52423230);  /* "RBSR" / "RB20" */
    put_i32(g_frames);
    put_i32(f->cur.p.w); put_i32(f->cur.p.h); put_i32(f->cur.p.bpc); put_i32(layout);
    put_i32(f->seq_hdr->sb128);
    put_i32(h->loopfilter.level_y[0]); put_i32(h->loopfilter.level_y[1]);
    put_i32(h->loopfilter.level_u); put_i32(h->loopfilter.level_v);
    put_i32(f->seq_hdr->cdef);
    put_i32(h->cdef.damping);
    for (int i = 0; i < 8; i++) put_i32(h->cdef.y_strength[i]);
    for (int i = 0; i < 8; i++) put_i32(h->cdef.uv_strength[i]);
    for (int i = 0; i < 3; i++) put_i32(f->lf.restore_planes & (1 << i) ? h->restoration.type[i] : 0);
    put_i32(h->restoration.unit_size[0]); put_i32(h->restoration.unit_size[1]);
    put_i32(h->tiling.cols); put_i32(h->tiling.rows);
    put_i32((int32_t)f->b4_stride); put_i32(f->sb128w); put_i32(f->sb128h);
    put_i32((int32_t)f->cur.stride[0]); put_i32((int32_t)f->cur.stride[1]);
    put_i32(ah); put_i32(n_planes);
    put_i32(h->frame_type); put_i32(h->show_frame);
    if (is_sr) {
        put_i32(h->width[1]); put_i32((int32_t)f->sr_cur.p.stride[0]); put_i32((int32_t)f->sr_cur.p.stride[1]);
        put_i32(f->sr_sb128w);
        put_i32(f->resize_step[0]); put_i32(f->resize_step[1]); put_i32(f->resize_start[0]); put_i32(f->resize_start[1]);
    }
    fwrite(f->lf.mask, sizeof(Av1Filter), n_sb128, g_out);
    fwrite(f->lf.level, 4, (size_t)f->b4_stride * 32 * f->sb128h, g_out);
    fwrite(&f->lf.lim_lut, sizeof(Av1FilterLUT), 1, g_out);
    fwrite(f->lf.lr_mask, sizeof(Av1Restoration), (size_t)f->sr_sb128w * f->sb128h, g_out);
    for (int pl = 0; pl < n_planes; pl++) {
        const int rows = pl ? ah >> ss_ver : ah;
        fwrite(g_pre[pl], 1, (size_t)f->cur.stride[!!pl] * rows, g_out);
    }
    for (int pl = 0; pl < n_planes; pl++) {   /* the filtered picture: sr_cur is cur itself without super-resolution */
        const int rows = pl ? ah >> ss_ver : ah;
        fwrite(f->sr_cur.p.data[pl], 1, (size_t)f->sr_cur.p.stride[!!pl] * rows, g_out);
    }
    (void)px; (void)ss_hor;
    g_frames++;
}

/* ---- film grain on output: dav1d_apply_grain_{8,16}bpc (src/fg_apply_tmpl.c:229-245), called from
 * dav1d_apply_grain (src/lib.c:482-516).  Record: "RBFG", w, h, bpc, layout, is_identity, stride_y,
 * stride_uv, sizeof(Dav1dFilmGrainData), the struct, the input planes, the output planes (h rows each). */
typedef void (*apply_grain_fn)(const Dav1dFilmGrainDSPContext *dsp, Dav1dPicture *out, const Dav1dPicture *in);
static int g_grain_frames, g_max_grain = 0;
static void grain_hook(const Dav1dFilmGrainDSPContext *dsp, Dav1dPicture *out, const Dav1dPicture *in, const char *sym) {
    apply_grain_fn real = (apply_grain_fn)dlsym(RTLD_NEXT, sym);
    if (!real) { fprintf(stderr, "ref_dump: cannot find %s\n", sym); exit(2); }
    real(dsp, out, in);
    if (g_grain_frames >= g_max_grain || in->stride[0] <= 0) return;
    const int layout = in->p.layout;
    const int ss_ver = layout == DAV1D_PIXEL_LAYOUT_I420;
    const int n_planes = layout == DAV1D_PIXEL_LAYOUT_I400 ? 1 : 3;
    put_i32(0x52424647);
    put_i32(in->p.w); put_i32(in->p.h); put_i32(in->p.bpc); put_i32(layout);
    put_i32(out->seq_hdr->mtrx == DAV1D_MC_IDENTITY);
    put_i32((int32_t)in->stride[0]); put_i32((int32_t)in->stride[1]);
    put_i32((int32_t)sizeof(Dav1dFilmGrainData));
    fwrite(&out->frame_hdr->film_grain.data, sizeof(Dav1dFilmGrainData), 1, g_out);
    for (int k = 0; k < 2; k++) {
        const Dav1dPicture *p = k ? out : in;
        for (int pl = 0; pl < n_planes; pl++) {
            const int rows = pl ? (in->p.h + ss_ver) >> ss_ver : in->p.h;
            fwrite(p->data[pl], 1, (size_t)p->stride[!!pl] * rows, g_out);
        }
    }
    g_grain_frames++;
}
__attribute__((visibility("default"))) void dav1d_apply_grain_8bpc(const Dav1dFilmGrainDSPContext *dsp, Dav1dPicture *out, const Dav1dPicture *in) { grain_hook(dsp, out, in, "dav1d_apply_grain_8bpc"); }
__attribute__((visibility("default"))) void dav1d_apply_grain_16bpc(const Dav1dFilmGrainDSPContext *dsp, Dav1dPicture *out, const Dav1dPicture *in) { grain_hook(dsp, out, in, "dav1d_apply_grain_16bpc"); }

/* exported so that the library's address-of in src/decode.c:3457 binds here */
__attribute__((visibility("default"))) void dav1d_filter_sbrow_8bpc(Dav1dFrameContext *f, int sby) { hook(f, sby, "dav1d_filter_sbrow_8bpc"); }
__attribute__((visibility("default"))) void dav1d_filter_sbrow_16bpc(Dav1dFrameContext *f, int sby) { hook(f, sby, "dav1d_filter_sbrow_16bpc"); }

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s in.ivf out.bin [max_frames]\n", argv[0]); return 2; }
    if (argc > 3) g_max_frames = atoi(argv[3]);
    if (argc > 4) g_max_grain = atoi(argv[4]);      /* film-grain records (decoded with apply_grain = 1) */
    if (argc > 5) g_sr_mode = atoi(argv[5]) != 0;   /* 1: super-resolution frames only */
    if (getenv("RB200_DUMP_LFB")) g_lfb_out = fopen(getenv("RB200_DUMP_LFB"), "wb");
    g_out = fopen(argv[2], "wb");
    if (!g_out) { perror(argv[2]); return 2; }
    DemuxerContext *in;
    unsigned fps[2], total, timebase[2];
    if (input_open(&in, NULL, argv[1], fps, &total, timebase) < 0) return 2;
    Dav1dSettings s;
    dav1d_default_settings(&s);
    s.n_threads = 1;            /* the single-thread path is the one that calls filter_sbrow per sbrow */
    s.max_frame_delay = 1;
    s.apply_grain = g_max_grain > 0;
    Dav1dContext *c;
    if (dav1d_open(&c, &s)) return 2;
    Dav1dData data;
    memset(&data, 0, sizeof(data));
    if (input_read(in, &data) < 0) return 2;
    int res = 0;
    do {
        Dav1dPicture p;
        memset(&p, 0, sizeof(p));
        res = dav1d_send_data(c, &data);
        if (res < 0 && res != DAV1D_ERR(EAGAIN)) { dav1d_data_unref(&data); if (res != DAV1D_ERR(EINVAL)) break; }
        res = dav1d_get_picture(c, &p);
        if (res >= 0) dav1d_picture_unref(&p);
        else if (res != DAV1D_ERR(EAGAIN) && res != DAV1D_ERR(EINVAL)) break;
        if (g_frames >= g_max_frames && g_grain_frames >= g_max_grain) break;
    } while (data.sz > 0 || !input_read(in, &data));
    if (data.sz > 0) dav1d_data_unref(&data);
    for (;;) {
        Dav1dPicture p;
        memset(&p, 0, sizeof(p));
        if (dav1d_get_picture(c, &p) < 0) break;
        dav1d_picture_unref(&p);
    }
    input_close(in);
    dav1d_close(&c);
    fclose(g_out);
    if (g_lfb_out) fclose(g_lfb_out);
    fprintf(stderr, "ref_dump: %d frames, %d film-grain records, %d super-resolution frames seen\n", g_frames, g_grain_frames, g_sr_seen);
    return g_frames + g_grain_frames > 0 ? 0 : 1;
}
