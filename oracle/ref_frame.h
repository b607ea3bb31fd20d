/*
 * oracle/ref_frame.h -- TEST INFRASTRUCTURE ONLY.  The frame harness of oracle/ref_frame.c as seen by
 * oracle/ref_backend.c (the CPU checker behind the host layer's backend seam).
 */
#ifndef ORACLE_REF_FRAME_H
#define ORACLE_REF_FRAME_H

#include "src/internal.h"
#include "../include/rav1d_b200.h"

typedef struct RefFrame {
    Dav1dContext *c;
    Dav1dFrameContext *f;
    Dav1dSequenceHeader seq;
    Dav1dFrameHeader hdr;
    Dav1dTaskContext *tc;   /* n_tc entries */
    int n_tc, hbd, bdmax;
    uint8_t *plane_mem;
    size_t plane_bytes;
    uint8_t *lvl_mem;
    uint8_t start_of_tile_row[1024];
    uint8_t *grain_mem;     /* output picture of dav1d_apply_grain */
    Dav1dPicture grain_out;
} RefFrame;

/* Replays of the batch records through the reference's own DSP tables (see ref_frame.c). */
void ref_frame_recon(RefFrame *r, RefFrame *const refs[], int n_refs, const Rb200McItem *mc, int n_mc,
                     const Rb200ItxItem *itx, int n_itx, void *coef_work, int n_threads);
void ref_frame_recon_comp(RefFrame *r, RefFrame *const refs[], int n_refs, const Rb200CompItem *items, int n, int n_threads);
void ref_frame_recon_warp(RefFrame *r, RefFrame *const refs[], int n_refs, const Rb200WarpItem *items, int n, int n_threads);
void ref_frame_recon_obmc(RefFrame *r, RefFrame *const refs[], int n_refs, const Rb200McItem *items, int n);
void ref_frame_recon_scaled(RefFrame *r, RefFrame *const refs[], int n_refs, const Rb200McScaledItem *items, int n);
void ref_frame_recon_intra(RefFrame *r, const Rb200IntraItem *items, int n, const int32_t *itx_of, const Rb200ItxItem *itx,
                           void *coef_work, const uint8_t *pal_buf);
#endif
