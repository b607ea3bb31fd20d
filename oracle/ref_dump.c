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

static FILE *g_out;
static int g_frames, g_max_frames = 1 << 30, g_sr_seen;
static int g_sr_mode;   /* 1: dump only super-resolution frames (record "RBSR": 8 more header ints, post = sr_cur) */
static uint8_t *g_pre[3];
static size_t g_pre_sz[3];

typedef void (*filter_sbrow_fn)(Dav1dFrameContext *f, int sby);

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
    fprintf(stderr, "ref_dump: %d frames, %d film-grain records, %d super-resolution frames seen\n", g_frames, g_grain_frames, g_sr_seen);
    return g_frames + g_grain_frames > 0 ? 0 : 1;
}
