/*
 * oracle/ref_backend.c -- TEST INFRASTRUCTURE ONLY (linked into oracle/_ref/dav1d_b200_cpucheck).
 *
 * A CPU checker behind the host layer's backend seam (rav1d_b200/host/rb200_host.h): it executes a frame's batch
 * -- the records rav1d_b200/host/recon_batch_tmpl.c appended -- with the REFERENCE'S OWN DSP functions, in the stage
 * order of rb200_frame_submit (all predictions, then compound, scaled, warped, OBMC above, OBMC left, the inter
 * residuals by transform size, then the intra wavefront level by level), and then runs the reference's own
 * filter_sbrow_* drivers over the whole frame.  A stream decoded this way and matching the conformance MD5 pins the
 * batch builder, the level assignment and the stage order without a GPU; the product never links this file.
 */
#define _GNU_SOURCE
#include "config.h"

#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "src/internal.h"
#include "ref_frame.h"
#include "../rav1d_b200/host/rb200_host.h"

static char g_err[256];
static const char *cpu_last_error(void) { return g_err; }
static int cpu_init(void) { return 0; }
static void *cpu_host_alloc(size_t bytes) { void *p = NULL; if (posix_memalign(&p, 64, bytes)) return NULL; memset(p, 0, bytes); return p; }
static void cpu_host_free(void *p) { free(p); }
static void *cpu_pic_new(void) { return malloc(8); }
static void cpu_pic_free(void *p) { free(p); }

typedef void (*sbrow_fn)(Dav1dFrameContext *, int);
typedef void (*sbrow_cdef_fn)(Dav1dTaskContext *, int);
static void *sym(const char *base, int hbd) {
    char name[96];
    snprintf(name, sizeof(name), "%s_%dbpc", base, hbd ? 16 : 8);
    void *p = dlsym(RTLD_NEXT, name);
    if (!p) { fprintf(stderr, "ref_backend: cannot find %s\n", name); exit(4); }
    return p;
}

static int cpu_frame_stage(const RbHostFrameDesc *const d, const RbHostBatch *const B, const RbHostFinal *const fin) {
    Dav1dFrameContext *const f = d->decoder_frame;
    const int hbd = f->cur.p.bpc > 8;
    RefFrame cur;
    memset(&cur, 0, sizeof(cur));
    cur.c = (Dav1dContext *)f->c; cur.f = f; cur.n_tc = 1; cur.hbd = hbd; cur.bdmax = f->bitdepth_max;
    /* reference pictures: shells whose `cur` is the reference picture (that is all the replays read) */
    static Dav1dFrameContext *shell_fc;
    static RefFrame shell[7];
    if (!shell_fc) shell_fc = calloc(7, sizeof(*shell_fc));
    RefFrame *refs[8] = { 0 };
    for (int i = 0; i < 7; i++) {
        shell_fc[i].cur = f->refp[i].p;
        shell[i].f = &shell_fc[i]; shell[i].hbd = hbd; shell[i].bdmax = f->bitdepth_max;
        refs[i] = &shell[i];
    }
    void *const cf = f->frame_thread.cf;

    if (d->stages & RB200_STAGE_RECON) {
        ref_frame_recon(&cur, refs, 7, B->mc.v, B->mc.n, NULL, 0, cf, 1);
        ref_frame_recon_comp(&cur, refs, 7, B->comp.v, B->comp.n, 1);
        ref_frame_recon_scaled(&cur, refs, 7, B->scaled.v, B->scaled.n);
        ref_frame_recon_warp(&cur, refs, 7, B->warp.v, B->warp.n, 1);
        ref_frame_recon_obmc(&cur, refs, 7, B->obmc_above.v, B->obmc_above.n);
        ref_frame_recon_scaled(&cur, refs, 7, B->scaled_obmc_above.v, B->scaled_obmc_above.n);
        ref_frame_recon_obmc(&cur, refs, 7, B->obmc_left.v, B->obmc_left.n);
        ref_frame_recon_scaled(&cur, refs, 7, B->scaled_obmc_left.v, B->scaled_obmc_left.n);
        ref_frame_recon(&cur, refs, 7, NULL, 0, fin->itx, fin->n_itx_inter, cf, 1);
        if (d->stages & RB200_STAGE_INTRA)
            ref_frame_recon_intra(&cur, fin->intra, fin->n_intra, fin->intra_itx, fin->itx, cf, B->pal.v);
    }

    /* the reference's filter drivers, stage by stage over all superblock rows (the decoder runs with task threads, so
     * its lpf / cdef line buffers hold every row: src/decode.c:2911-3004) */
    static Dav1dTaskContext *tc;
    if (!tc) tc = calloc(1, sizeof(*tc));
    tc->c = f->c; tc->f = f; tc->top_pre_cdef_toggle = 0;
    const sbrow_fn cols = (sbrow_fn)sym("dav1d_filter_sbrow_deblock_cols", hbd), rows = (sbrow_fn)sym("dav1d_filter_sbrow_deblock_rows", hbd),
                   resize = (sbrow_fn)sym("dav1d_filter_sbrow_resize", hbd), lr = (sbrow_fn)sym("dav1d_filter_sbrow_lr", hbd);
    const sbrow_cdef_fn cdef = (sbrow_cdef_fn)sym("dav1d_filter_sbrow_cdef", hbd);
    for (int sby = 0; sby < f->sbh; sby++) cols(f, sby);
    for (int sby = 0; sby < f->sbh; sby++) rows(f, sby);
    if (f->seq_hdr->cdef) for (int sby = 0; sby < f->sbh; sby++) cdef(tc, sby);
    if (f->frame_hdr->width[0] != f->frame_hdr->width[1]) for (int sby = 0; sby < f->sbh; sby++) resize(f, sby);
    if (f->lf.restore_planes) for (int sby = 0; sby < f->sbh; sby++) lr(f, sby);
    return 0;
}

static int cpu_frame_submit(void *cur) { (void)cur; return 0; }
static int cpu_pic_fetch(void *pic, void *const data[3], const ptrdiff_t stride[2], int grain) {
    (void)pic; (void)data; (void)stride; (void)grain;
    return 0;
}

static const RbHostBackend g_backend = {
    "cpu-check", 1, cpu_init, cpu_host_alloc, cpu_host_free, cpu_pic_new, cpu_pic_free, cpu_frame_stage, cpu_frame_submit,
    cpu_pic_fetch, cpu_last_error,
};
const RbHostBackend *rb200_host_backend(void) { return &g_backend; }
