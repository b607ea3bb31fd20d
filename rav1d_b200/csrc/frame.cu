// Frame-level pipeline: device-resident planes, per-frame batch staging and
// stream-ordered stage launches.
//
// This is the B200 counterpart of the reference's pass-2 reconstruction plus the
// filter tasks of its scheduler (TileReconstruction -> DeblockCols -> DeblockRows
// -> Cdef -> LoopRestoration, src/thread_task.rs:1048-1241; single-thread form
// rav1d_decode_frame_main src/decode.rs:4497-4570 calling filter_sbrow
// src/recon.rs:4319-4338).  The per-sbrow wavefront of the CPU becomes
// whole-frame launches ordered on one CUDA stream:
//
//   H2D(coef, items, masks)  ->  MC (prediction into `cur`)  ->  itx add (per tx size)
//   -> deblock column edges -> deblock row edges (in place on `cur`)
//   -> CDEF (cur -> p2, out of place) -> loop restoration (p2 + cur -> p3)
//
// and the in-place backups of the CPU (cdef_line_buf, lr_line_buf, left[][])
// become three ping-pong plane sets.  The front end (entropy decode, mode
// parsing, lf-mask generation) stays on the host and fills the pinned staging
// arrays in the reference's own formats (Av1Filter, level[4], Av1Restoration,
// packed cf).
#include "common.cuh"
#include "tma.cuh"
#include "stages.cuh"
#include <new>
#include <stdio.h>
#include <stdlib.h>

struct Rb200Frame {
    Rb200FrameHeader hdr;
    Rb200FrameGeometry g;
    int bdmax;
    size_t px, cs;
    cudaStream_t stream;
    // device plane sets: 0 = cur (recon, deblocked in place), 1 = CDEF output, 2 = LR output
    uint8_t *plane_mem[3];
    Rb200Planes planes[3];
    Rb200Planes out;
    Rb200Planes refs[8];
    int n_refs;
    // batch staging: pinned host + device mirror
    size_t max_coefs; int max_itx, max_mc;
    void *h_coef, *d_coef;
    int16_t *h_coef16;                              // int16 transport of a 16-bit picture's coefficients (allocated on first use)
    Rb200CoefEscape *h_esc, *d_esc; int max_esc, n_esc;
    int16_t *h_pk, *d_pk; uint32_t *h_pkoff, *d_pkoff; size_t n_pk;   // packed int16 coefficient stream (RB200_UPLOAD_PACKED_COEF16)
    Rb200ItxItem *h_itx, *d_itx;
    Rb200McItem *h_mc, *d_mc;
    Rb200CompItem *h_comp, *d_comp; int max_comp, n_comp;
    Rb200WarpItem *h_warp, *d_warp; int max_warp, n_warp;
    Rb200McItem *h_obmc, *d_obmc; int max_obmc, n_obmc_above, n_obmc_left;
    Rb200McScaledItem *h_scaled, *d_scaled; int max_scaled, n_scaled, n_scaled_above, n_scaled_left;   // put items, then OBMC strips
    rb200::McRefDims ref_dims;
    rb200::McGmvSet ref_gmv;      // rb200_frame_set_ref_gmv
    // intra blocks, level by level (RB200_STAGE_INTRA)
    Rb200IntraItem *h_intra, *d_intra; int max_intra, max_levels, n_levels;
    int32_t *h_intra_itx, *d_intra_itx;           // per intra item: index of its residual in the itx list, -1 = none
    uint8_t *h_pal, *d_pal; size_t max_pal, n_pal; // palette records of the palette blocks
    int32_t *intra_counts, *intra_itx_counts;     // [max_levels], [max_levels][RB200_N_RECT_TX_SIZES]
    int32_t *h_level_off, *d_level_off;           // [max_levels + 1] item offsets of the levels (one-launch wavefront)
    unsigned *d_intra_sync, *h_intra_sync;        // arrival counter + time-limit flag of that launch
    int intra_widest, n_levels_nonempty; bool intra_check;
    // super-resolution (hdr.upscaled_width > hdr.width): plane sets at the upscaled width --
    // 0 = upscaled CDEF output, 1 = upscaled deblocked picture (what lr_line_buf holds on the CPU), 2 = LR output
    bool sr;
    int sr_w, out_w;                    // upscaled luma width; luma width of `out` / `display`
    int resize_step[2], resize_start[2];
    uint8_t *sr_mem[3];
    Rb200Planes sr_planes[3];
    Rb200Av1Filter *h_masks, *d_masks;
    uint8_t (*h_lvl)[4], (*d_lvl)[4];
    Rb200Av1FilterLUT *h_lut, *d_lut;
    Rb200Av1Restoration *h_lr, *d_lr;
    // loop-filter masks / levels generated on the device from per-block records (lfmask.cu)
    Rb200LfBlock *h_lfb, *d_lfb; int max_lfb, n_lfb;
    uint8_t *d_lf_cells;                // per-4x4 cell facts, luma + chroma
    int8_t *h_cdef_idx, *d_cdef_idx;    // [n_masks][4], the only Av1Filter member still uploaded in that mode
    cudaStream_t lf_stream;             // records upload + mask build run beside the reconstruction
    cudaEvent_t lf_fork, lf_join;
    void *d_cdef_blk;   // per-8x8 CDEF decisions (direction, strengths), device only
    CUtensorMap tm_cdef[3]; bool tm_cdef_ok;
    CUtensorMap tm_lr_main[2][3], tm_lr_halo[3]; bool tm_lr_ok;   // loop-restoration windows: plane sets 0 / 1 (8-row boxes), set 0 (2-row boxes)
    rb200::McRefMapCache tm_refs;      // tensor maps of the reference planes for the prediction kernel's window fetches   // tensor maps of plane set 0 for the CDEF tile loads (16-bit pictures)
    int *d_counters;    // work dispensers of the batch kernels (one int each)
    int band_s0, band_s1;   // loop-restoration stripes this context produces (0, 0 = whole picture)
    size_t n_masks, n_lvl;
    int launches;
    // film grain on output (src/fg_apply.rs): parameters, LUTs, display planes
    bool fg_set; int fg_is_id;
    Rb200FilmGrainData fg;
    uint8_t *fg_mem;            // device: 3 grain LUTs, 3 scaling LUTs, points, offsets
    cudaStream_t fg_stream;     // LUT / scaling / offset preparation runs beside the reconstruction
    cudaEvent_t fg_fork, fg_join;
    cudaStream_t up_stream;     // high priority: the coefficient gather over PCIe (RB200_UPLOAD_GATHER_COEF)
    cudaEvent_t up_fork, up_join;
    uint8_t *plane_mem_fg; Rb200Planes planes_fg, display;
    // optional per-stage timing (the analogue of the reference CLI's --frametimes, tools/dav1d.rs:127-150)
    cudaStream_t own_stream;
    // luma | chroma post-filter chains on two streams (rb200_frame_set_plane_streams)
    bool plane_split;
    bool plane_counts; int n_mc_luma; int32_t itx_luma[RB200_N_RECT_TX_SIZES];   // rb200_frame_set_plane_counts
    cudaStream_t uv_stream;
    cudaEvent_t uv_fork, uv_dir, uv_join, itx_fork, itx_join;
    cudaEvent_t done_event;     // recorded behind the last kernel of every submit (what rb200_frame_depend waits for)
    cudaEvent_t dep_events[8];  // rb200_frame_depend: the producers' done events, waited for by the next submit
    int n_deps;
    bool timing;
    cudaEvent_t ev[RB200_N_FRAME_MARKS];
    bool ev_valid[RB200_N_FRAME_MARKS];
};

using namespace rb200;

namespace {

int alloc_planes(Rb200Frame *f, int which) {
    const size_t ysz = (size_t)f->g.stride[0] * f->g.plane_h[0];
    const size_t uvsz = f->g.n_planes > 1 ? (size_t)f->g.stride[1] * f->g.plane_h[1] : 0;
    RB_CUDA(cudaMalloc((void **)&f->plane_mem[which], ysz + 2 * uvsz + 256));
    RB_CUDA(cudaMemsetAsync(f->plane_mem[which], 0, ysz + 2 * uvsz + 256, f->stream));
    Rb200Planes &p = f->planes[which];
    p.data[0] = f->plane_mem[which]; p.stride[0] = f->g.stride[0];
    p.data[1] = uvsz ? f->plane_mem[which] + ysz : nullptr;
    p.data[2] = uvsz ? f->plane_mem[which] + ysz + uvsz : nullptr;
    p.stride[1] = p.stride[2] = uvsz ? f->g.stride[1] : 0;
    return 0;
}

// src/decode.rs (rav1d_submit_frame) == src/decode.c:3325-3329,3509-3510,3567-3576
int upscale_x0(int in_w, int out_w, int step) {
    const int err = out_w * step - (in_w << 14);
    const int x0 = (-((out_w - in_w) << 13) + (out_w >> 1)) / out_w + 128 - (err / 2);
    return x0 & 0x3fff;
}
int alloc_sr_planes(Rb200Frame *f, int which) {
    const int aw = (f->sr_w + 127) & ~127;
    const int64_t sy = (int64_t)aw * (int64_t)f->px, suv = f->g.n_planes > 1 ? (int64_t)(aw >> f->g.ss_hor) * (int64_t)f->px : 0;
    const size_t ysz = (size_t)sy * f->g.plane_h[0], uvsz = (size_t)suv * f->g.plane_h[1];
    RB_CUDA(cudaMalloc((void **)&f->sr_mem[which], ysz + 2 * uvsz + 256));
    RB_CUDA(cudaMemsetAsync(f->sr_mem[which], 0, ysz + 2 * uvsz + 256, f->stream));
    Rb200Planes &p = f->sr_planes[which];
    p.data[0] = f->sr_mem[which]; p.stride[0] = sy;
    p.data[1] = uvsz ? f->sr_mem[which] + ysz : nullptr;
    p.data[2] = uvsz ? f->sr_mem[which] + ysz + uvsz : nullptr;
    p.stride[1] = p.stride[2] = suv;
    return 0;
}

template <typename T>
int alloc_pair(T **h, T **d, size_t n) {
    const size_t bytes = (n ? n : 1) * sizeof(T);
    RB_CUDA(cudaMallocHost((void **)h, bytes));
    memset(*h, 0, bytes);
    RB_CUDA(cudaMalloc((void **)d, bytes));
    // cudaMemset on device memory is asynchronous with respect to the host and runs on the legacy default stream, which
    // the frames' non-blocking streams do not wait for: without the synchronize a late memset can land on top of the
    // first batch uploaded into this buffer (seen as garbage frames when several processes share the GPU).
    RB_CUDA(cudaMemset(*d, 0, bytes));
    RB_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
    return 0;
}

}  // namespace

extern "C" int rb200_frame_create(Rb200Frame **out, const Rb200FrameHeader *hdr, size_t max_coefs, int max_itx,
                                  int max_mc) {
    if (!out || !hdr) return set_error(-22, "frame_create: null argument");
    if (hdr->width < 1 || hdr->height < 1 || hdr->width > 16384 || hdr->height > 16384 ||
        (hdr->bpc != 8 && hdr->bpc != 10 && hdr->bpc != 12) || hdr->layout < 0 || hdr->layout > 3)
        return set_error(-22, "frame_create: bad header");
    Rb200Frame *f = new (std::nothrow) Rb200Frame();
    if (!f) return set_error(-12, "frame_create: out of memory");
    memset(f, 0, sizeof(*f));
    f->hdr = *hdr;
    f->bdmax = (1 << hdr->bpc) - 1;
    f->px = hdr->bpc > 8 ? 2 : 1;
    f->cs = hdr->bpc > 8 ? 4 : 2;
    Rb200FrameGeometry &g = f->g;
    // src/decode.rs:4880-4905 (dav1d_submit_frame): bw/bh in 4-px units rounded to 8 px
    g.bw = ((hdr->width + 7) >> 3) << 1;
    g.bh = ((hdr->height + 7) >> 3) << 1;
    g.w4 = (hdr->width + 3) >> 2;
    g.h4 = (hdr->height + 3) >> 2;
    g.sb128w = (g.bw + 31) >> 5;
    g.sb128h = (g.bh + 31) >> 5;
    const int sb_shift = 4 + hdr->sb128;
    g.sbh = (g.bh + (1 << sb_shift) - 1) >> sb_shift;
    g.b4_stride = (g.bw + 31) & ~31;
    g.ss_hor = hdr->layout != RB200_LAYOUT_I444 && hdr->layout != RB200_LAYOUT_I400;
    g.ss_ver = hdr->layout == RB200_LAYOUT_I420;
    g.n_planes = hdr->layout == RB200_LAYOUT_I400 ? 1 : 3;
    // picture allocation: dimensions aligned up to 128 (src/picture.rs:91-137); the device
    // planes are the library's own, so no extra 64-byte anti-aliasing pad is needed.
    const int aw = (hdr->width + 127) & ~127, ah = (hdr->height + 127) & ~127;
    g.stride[0] = (int64_t)aw * (int64_t)f->px;
    g.stride[1] = g.n_planes > 1 ? (int64_t)(aw >> g.ss_hor) * (int64_t)f->px : 0;
    g.plane_h[0] = ah;
    g.plane_h[1] = g.n_planes > 1 ? ah >> g.ss_ver : 0;
    f->max_coefs = max_coefs; f->max_itx = max_itx; f->max_mc = max_mc;
    f->sr = hdr->upscaled_width > hdr->width;
    if (hdr->upscaled_width && (hdr->upscaled_width < hdr->width || hdr->upscaled_width > 2 * hdr->width)) {
        delete f;
        return set_error(-22, "frame_create: upscaled_width must be in [width, 2 * width]");
    }
    f->sr_w = f->sr ? hdr->upscaled_width : hdr->width;
    f->out_w = hdr->width;
    if (f->sr) {
        const int in_cw = (hdr->width + g.ss_hor) >> g.ss_hor, out_cw = (f->sr_w + g.ss_hor) >> g.ss_hor;
        f->resize_step[0] = ((hdr->width << 14) + (f->sr_w >> 1)) / f->sr_w;
        f->resize_step[1] = ((in_cw << 14) + (out_cw >> 1)) / out_cw;
        f->resize_start[0] = upscale_x0(hdr->width, f->sr_w, f->resize_step[0]);
        f->resize_start[1] = upscale_x0(in_cw, out_cw, f->resize_step[1]);
    }
    // loop-restoration units are indexed in upscaled coordinates (f.sr_sb128w)
    f->n_masks = (size_t)imax(g.sb128w, (f->sr_w + 127) >> 7) * g.sb128h;
    f->n_lvl = (size_t)g.b4_stride * 32 * g.sb128h + 32;  // + guard for the level fallback of row/col 0
    int r = 0;
    cudaError_t e = cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { r = cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__); delete f; return r; }
    f->own_stream = f->stream;
    f->plane_split = true;
    e = cudaEventCreateWithFlags(&f->done_event, cudaEventDisableTiming);
    if (e != cudaSuccess) r = cuda_fail(e, "cudaEventCreate", __FILE__, __LINE__);
    for (int i = 0; i < 3 && !r; i++) r = alloc_planes(f, i);
    for (int i = 0; i < 3 && !r && f->sr; i++) r = alloc_sr_planes(f, i);
    if (!r) {
        e = cudaMallocHost(&f->h_coef, (max_coefs ? max_coefs : 1) * f->cs);
        if (e == cudaSuccess) {
            memset(f->h_coef, 0, (max_coefs ? max_coefs : 1) * f->cs);
            e = cudaMalloc(&f->d_coef, (max_coefs ? max_coefs : 1) * f->cs);
        }
        if (e != cudaSuccess) r = cuda_fail(e, "coefficient staging", __FILE__, __LINE__);
    }
    if (!r) r = alloc_pair(&f->h_itx, &f->d_itx, (size_t)max_itx);
    if (!r) r = alloc_pair(&f->h_mc, &f->d_mc, (size_t)max_mc);
    if (!r) r = alloc_pair(&f->h_masks, &f->d_masks, f->n_masks);
    if (!r) r = alloc_pair((uint8_t **)&f->h_lvl, (uint8_t **)&f->d_lvl, f->n_lvl * 4);
    if (!r) r = alloc_pair(&f->h_lut, &f->d_lut, 1);
    if (!r) r = alloc_pair(&f->h_lr, &f->d_lr, f->n_masks);
    if (!r) { e = cudaMalloc(&f->d_cdef_blk, (size_t)(g.bw >> 1) * (g.bh >> 1) * 8 + 64); if (e != cudaSuccess) r = cuda_fail(e, "cudaMalloc", __FILE__, __LINE__); }
    if (!r) { e = cudaMalloc((void **)&f->d_counters, 64); if (e != cudaSuccess) r = cuda_fail(e, "cudaMalloc", __FILE__, __LINE__); }
    if (!r) { e = cudaStreamSynchronize(f->stream); if (e != cudaSuccess) r = cuda_fail(e, "sync", __FILE__, __LINE__); }
    if (!r && hdr->bpc > 8) {
        CdefFrameParams P = {};
        P.bw = g.bw; P.bh = g.bh; P.ss_hor = g.ss_hor; P.ss_ver = g.ss_ver; P.n_planes = g.n_planes;
        r = cdef_encode_maps(f->tm_cdef, f->planes[0], P);
        f->tm_cdef_ok = !r;
        for (int p = 0; p < g.n_planes && !r && !f->sr; p++) {
            const int ssh = p ? g.ss_hor : 0, ssv = p ? g.ss_ver : 0;
            const int pw = (hdr->width + ssh) >> ssh, ph = (hdr->height + ssv) >> ssv;
            r = lr_encode_maps(&f->tm_lr_main[0][p], &f->tm_lr_halo[p], f->planes[0].data[p], f->planes[0].data[p], f->planes[0].stride[p], pw, ph);
            CUtensorMap unused;
            if (!r) r = lr_encode_maps(&f->tm_lr_main[1][p], &unused, f->planes[1].data[p], f->planes[0].data[p], f->planes[1].stride[p], pw, ph);
        }
        f->tm_lr_ok = !r && !f->sr;
    }
    if (r) { rb200_frame_destroy(f); return r; }
    f->out = f->planes[0];
    f->display = f->out;
    *out = f;
    return 0;
}

extern "C" int rb200_frame_destroy(Rb200Frame *f) {
    if (!f) return 0;
    if (f->stream) cudaStreamSynchronize(f->stream);
    for (int i = 0; i < RB200_N_FRAME_MARKS; i++) if (f->ev[i]) cudaEventDestroy(f->ev[i]);
    for (int i = 0; i < 3; i++) if (f->plane_mem[i]) cudaFree(f->plane_mem[i]);
    if (f->fg_mem) cudaFree(f->fg_mem);
    if (f->fg_stream) { cudaStreamSynchronize(f->fg_stream); cudaStreamDestroy(f->fg_stream); }
    if (f->lf_stream) { cudaStreamSynchronize(f->lf_stream); cudaStreamDestroy(f->lf_stream); }
    if (f->lf_fork) cudaEventDestroy(f->lf_fork);
    if (f->lf_join) cudaEventDestroy(f->lf_join);
    if (f->up_stream) { cudaStreamSynchronize(f->up_stream); cudaStreamDestroy(f->up_stream); }
    if (f->up_fork) cudaEventDestroy(f->up_fork);
    if (f->up_join) cudaEventDestroy(f->up_join);
    if (f->fg_fork) cudaEventDestroy(f->fg_fork);
    if (f->fg_join) cudaEventDestroy(f->fg_join);
    if (f->plane_mem_fg) cudaFree(f->plane_mem_fg);
    if (f->h_coef) cudaFreeHost(f->h_coef);
    if (f->h_coef16) cudaFreeHost(f->h_coef16);
    if (f->h_pk) cudaFreeHost(f->h_pk);
    if (f->d_pk) cudaFree(f->d_pk);
    if (f->h_pkoff) cudaFreeHost(f->h_pkoff);
    if (f->d_pkoff) cudaFree(f->d_pkoff);
    if (f->h_esc) cudaFreeHost(f->h_esc);
    if (f->d_esc) cudaFree(f->d_esc);
    if (f->d_coef) cudaFree(f->d_coef);
    if (f->h_itx) cudaFreeHost(f->h_itx);
    if (f->d_itx) cudaFree(f->d_itx);
    for (int i = 0; i < 3; i++) if (f->sr_mem[i]) cudaFree(f->sr_mem[i]);
    if (f->h_obmc) cudaFreeHost(f->h_obmc);
    if (f->d_obmc) cudaFree(f->d_obmc);
    if (f->h_scaled) cudaFreeHost(f->h_scaled);
    if (f->d_scaled) cudaFree(f->d_scaled);
    if (f->h_intra) cudaFreeHost(f->h_intra);
    if (f->d_intra) cudaFree(f->d_intra);
    if (f->h_intra_itx) cudaFreeHost(f->h_intra_itx);
    if (f->d_intra_itx) cudaFree(f->d_intra_itx);
    if (f->h_pal) cudaFreeHost(f->h_pal);
    if (f->d_pal) cudaFree(f->d_pal);
    free(f->intra_counts); free(f->intra_itx_counts);
    if (f->h_warp) cudaFreeHost(f->h_warp);
    if (f->d_warp) cudaFree(f->d_warp);
    if (f->h_comp) cudaFreeHost(f->h_comp);
    if (f->d_comp) cudaFree(f->d_comp);
    if (f->h_mc) cudaFreeHost(f->h_mc);
    if (f->d_mc) cudaFree(f->d_mc);
    if (f->d_counters) cudaFree(f->d_counters);
    if (f->h_level_off) cudaFreeHost(f->h_level_off);
    if (f->d_level_off) cudaFree(f->d_level_off);
    if (f->h_intra_sync) cudaFreeHost(f->h_intra_sync);
    if (f->d_intra_sync) cudaFree(f->d_intra_sync);
    if (f->h_lfb) cudaFreeHost(f->h_lfb);
    if (f->d_lfb) cudaFree(f->d_lfb);
    if (f->d_lf_cells) cudaFree(f->d_lf_cells);
    if (f->h_cdef_idx) cudaFreeHost(f->h_cdef_idx);
    if (f->d_cdef_idx) cudaFree(f->d_cdef_idx);
    if (f->h_masks) cudaFreeHost(f->h_masks);
    if (f->d_masks) cudaFree(f->d_masks);
    if (f->h_lvl) cudaFreeHost(f->h_lvl);
    if (f->d_lvl) cudaFree(f->d_lvl);
    if (f->h_lut) cudaFreeHost(f->h_lut);
    if (f->d_lut) cudaFree(f->d_lut);
    if (f->h_lr) cudaFreeHost(f->h_lr);
    if (f->d_lr) cudaFree(f->d_lr);
    if (f->d_cdef_blk) cudaFree(f->d_cdef_blk);
    if (f->done_event) cudaEventDestroy(f->done_event);
    if (f->uv_fork) cudaEventDestroy(f->uv_fork);
    if (f->uv_dir) cudaEventDestroy(f->uv_dir);
    if (f->uv_join) cudaEventDestroy(f->uv_join);
    if (f->itx_fork) cudaEventDestroy(f->itx_fork);
    if (f->itx_join) cudaEventDestroy(f->itx_join);
    if (f->uv_stream) cudaStreamDestroy(f->uv_stream);
    if (f->own_stream) cudaStreamDestroy(f->own_stream);
    delete f;
    return 0;
}

// Per-frame parameters of a context that is reused for the next picture of the same geometry (a decoder keeps a pool of
// contexts; only the filter parameters change from frame to frame).
extern "C" int rb200_frame_set_params(Rb200Frame *f, const Rb200FrameHeader *hdr) {
    if (!f || !hdr) return set_error(-22, "frame_set_params: null argument");
    const Rb200FrameHeader &o = f->hdr;
    if (hdr->width != o.width || hdr->height != o.height || hdr->bpc != o.bpc || hdr->layout != o.layout ||
        hdr->sb128 != o.sb128 || hdr->upscaled_width != o.upscaled_width)
        return set_error(-22, "frame_set_params: the geometry of a frame context cannot change");
    f->hdr = *hdr;
    return 0;
}

extern "C" int rb200_frame_geometry(const Rb200Frame *f, Rb200FrameGeometry *g) {
    if (!f || !g) return set_error(-22, "frame_geometry: null argument");
    *g = f->g;
    return 0;
}
extern "C" void *rb200_frame_coef_buffer(Rb200Frame *f) { return f ? f->h_coef : nullptr; }
extern "C" Rb200ItxItem *rb200_frame_itx_items(Rb200Frame *f) { return f ? f->h_itx : nullptr; }
extern "C" Rb200McItem *rb200_frame_mc_items(Rb200Frame *f) { return f ? f->h_mc : nullptr; }
extern "C" Rb200Av1Filter *rb200_frame_lf_masks(Rb200Frame *f) { return f ? f->h_masks : nullptr; }
// The level array is handed out one 4x4 row (b4_stride entries) past its start so that the
// reference's `l[-b4_stride]` fallback for the first row stays inside the allocation.
extern "C" uint8_t (*rb200_frame_lf_levels(Rb200Frame *f))[4] { return f ? f->h_lvl + 32 : nullptr; }
extern "C" Rb200Av1FilterLUT *rb200_frame_lf_lut(Rb200Frame *f) { return f ? f->h_lut : nullptr; }
extern "C" Rb200Av1Restoration *rb200_frame_lr_masks(Rb200Frame *f) { return f ? f->h_lr : nullptr; }
extern "C" void *rb200_frame_stream(Rb200Frame *f) { return f ? (void *)f->stream : nullptr; }
extern "C" int rb200_frame_last_launches(const Rb200Frame *f) { return f ? f->launches : 0; }

constexpr size_t FG_LUT_BYTES = (size_t)(RB200_GRAIN_HEIGHT + 1) * RB200_GRAIN_WIDTH * 2;
constexpr size_t FG_OFF_LUT = 0, FG_OFF_SCALING = 3 * FG_LUT_BYTES + 64, FG_OFF_PTS = FG_OFF_SCALING + 3 * 4096,
                 FG_OFF_OFFSETS = FG_OFF_PTS + 128;

extern "C" int rb200_frame_set_film_grain(Rb200Frame *f, const Rb200FilmGrainData *data, int is_id) {
    if (!f || !data) return set_error(-22, "frame_set_film_grain: null argument");
    if (data->num_y_points < 0 || data->num_y_points > 14 || data->num_uv_points[0] < 0 || data->num_uv_points[0] > 10 ||
        data->num_uv_points[1] < 0 || data->num_uv_points[1] > 10 || data->ar_coeff_lag < 0 || data->ar_coeff_lag > 3)
        return set_error(-22, "frame_set_film_grain: bad parameters");
    if (!f->fg_mem) {
        // grain goes onto the OUTPUT picture: at the upscaled width when the frame is coded with super-resolution
        const int ow = f->sr ? f->sr_w : f->hdr.width, aw = (ow + 127) & ~127;
        const int rows = (f->hdr.height + 31) >> 5, cols = (ow + 31) >> 5;
        RB_CUDA(cudaMalloc((void **)&f->fg_mem, FG_OFF_OFFSETS + (size_t)rows * (cols + 1) + 64));
        RB_CUDA(cudaStreamCreateWithFlags(&f->fg_stream, cudaStreamNonBlocking));
        RB_CUDA(cudaEventCreateWithFlags(&f->fg_fork, cudaEventDisableTiming));
        RB_CUDA(cudaEventCreateWithFlags(&f->fg_join, cudaEventDisableTiming));
        const int64_t sy = (int64_t)aw * (int64_t)f->px, suv = f->g.n_planes > 1 ? (int64_t)(aw >> f->g.ss_hor) * (int64_t)f->px : 0;
        const size_t ysz = (size_t)sy * f->g.plane_h[0];
        const size_t uvsz = (size_t)suv * f->g.plane_h[1];
        RB_CUDA(cudaMalloc((void **)&f->plane_mem_fg, ysz + 2 * uvsz + 256));
        Rb200Planes &p = f->planes_fg;
        p.data[0] = f->plane_mem_fg; p.stride[0] = sy;
        p.data[1] = uvsz ? f->plane_mem_fg + ysz : nullptr;
        p.data[2] = uvsz ? f->plane_mem_fg + ysz + uvsz : nullptr;
        p.stride[1] = p.stride[2] = suv;
    }
    f->fg = *data; f->fg_is_id = is_id; f->fg_set = true;
    return 0;
}

extern "C" int rb200_frame_display_planes(Rb200Frame *f, Rb200Planes *out) {
    if (!f || !out) return set_error(-22, "frame_display_planes: bad argument");
    *out = f->display;
    return 0;
}

// rav1d_prep_grain (src/fg_apply.rs:74-172): grain LUTs, scaling LUTs and block offsets.  None of it
// depends on the picture, so it is launched on the frame's side stream at submit time and overlaps
// the reconstruction and the in-loop filters; film_grain_apply joins it.
static int film_grain_prepare(Rb200Frame *f, cudaStream_t st, int out_w) {
    const Rb200FilmGrainData &d = f->fg;
    const Rb200FrameGeometry &g = f->g;
    const int w = out_w, h = f->hdr.height, bpc = f->hdr.bpc;
    uint8_t *lut[3], *sc[3];
    for (int i = 0; i < 3; i++) { lut[i] = f->fg_mem + FG_OFF_LUT + i * FG_LUT_BYTES; sc[i] = f->fg_mem + FG_OFF_SCALING + i * 4096; }
    uint8_t *d_off = f->fg_mem + FG_OFF_OFFSETS;
    const bool chroma = g.n_planes > 1;
    const bool do_uv[2] = { chroma && (d.num_uv_points[0] || d.chroma_scaling_from_luma),
                            chroma && (d.num_uv_points[1] || d.chroma_scaling_from_luma) };
    int r;
    if ((r = fg_generate(lut[0], nullptr, d, -1, 0, 0, f->bdmax, st))) return r;
    f->launches++;
    if (do_uv[0] || do_uv[1]) {   // both chroma LUTs in one launch (they only depend on the luma LUT)
        const int first = do_uv[0] ? 0 : 1, n = (int)do_uv[0] + (int)do_uv[1];
        if ((r = fg_generate(lut[1 + first], lut[0], d, first, g.ss_hor, g.ss_ver, f->bdmax, st, n, FG_LUT_BYTES))) return r;
        f->launches++;
    }
    if (d.num_y_points || d.chroma_scaling_from_luma) { if ((r = fg_scaling(sc[0], d.y_points, d.num_y_points, bpc, st))) return r; f->launches++; }
    for (int pl = 0; pl < 2; pl++)
        if (chroma && d.num_uv_points[pl]) { if ((r = fg_scaling(sc[1 + pl], d.uv_points[pl], d.num_uv_points[pl], bpc, st))) return r; f->launches++; }
    const int rows = (h + 31) >> 5, ncols = ((w + 31) >> 5) + 1;
    if ((r = fg_offsets(d_off, d.seed, 0, rows, ncols, st))) return r;
    f->launches++;
    return 0;
}

// rav1d_apply_grain_row over the whole picture (src/fg_apply.rs:174-284)
static int film_grain_apply(Rb200Frame *f, cudaStream_t st) {
    const Rb200FilmGrainData &d = f->fg;
    const Rb200FrameGeometry &g = f->g;
    const int w = f->out_w, h = f->hdr.height;      // the output picture (upscaled width under super-resolution)
    uint8_t *lut[3], *sc[3];
    for (int i = 0; i < 3; i++) { lut[i] = f->fg_mem + FG_OFF_LUT + i * FG_LUT_BYTES; sc[i] = f->fg_mem + FG_OFF_SCALING + i * 4096; }
    uint8_t *d_off = f->fg_mem + FG_OFF_OFFSETS;
    const bool chroma = g.n_planes > 1;
    const bool do_uv[2] = { chroma && (d.num_uv_points[0] || d.chroma_scaling_from_luma),
                            chroma && (d.num_uv_points[1] || d.chroma_scaling_from_luma) };
    const int ncols = ((w + 31) >> 5) + 1;
    int r;
    const Rb200Planes in = f->out;
    f->display = in;
    if (d.num_y_points) {
        FgApplyParams P = fg_params(d, 0, 0, 0, 0, f->fg_is_id);
        P.pw = w; P.ph = h; P.row0 = 0; P.off_row0 = 0; P.ncols = ncols; P.luma_w = w;
        if ((r = fg_apply((uint8_t *)f->planes_fg.data[0], (const uint8_t *)in.data[0], in.stride[0], nullptr, 0, P, sc[0],
                          lut[0], d_off, f->bdmax, st))) return r;
        f->launches++;
        f->display.data[0] = f->planes_fg.data[0]; f->display.stride[0] = f->planes_fg.stride[0];
    }
    for (int pl = 0; pl < 2; pl++) {
        if (!do_uv[pl]) continue;
        FgApplyParams P = fg_params(d, 1, pl, g.ss_hor, g.ss_ver, f->fg_is_id);
        P.pw = (w + g.ss_hor) >> g.ss_hor; P.ph = (h + g.ss_ver) >> g.ss_ver;
        P.row0 = 0; P.off_row0 = 0; P.ncols = ncols; P.luma_w = w;
        if ((r = fg_apply((uint8_t *)f->planes_fg.data[1 + pl], (const uint8_t *)in.data[1 + pl], in.stride[1 + pl],
                          (const uint8_t *)in.data[0], in.stride[0], P,
                          d.chroma_scaling_from_luma ? sc[0] : sc[1 + pl], lut[1 + pl], d_off, f->bdmax, st))) return r;
        f->launches++;
        f->display.data[1 + pl] = f->planes_fg.data[1 + pl]; f->display.stride[1 + pl] = f->planes_fg.stride[1 + pl];
    }
    return 0;
}

// Uploads cover the picture rounded up to 8 pixels (f.bw * 4 x f.bh * 4): the reconstruction of blocks that
// straddle the picture edge lives there and CDEF reads it (src/cdef_apply.rs:159-507 walks bw / bh).
static int upload_rows(const Rb200Frame *f, int p) { return (f->g.bh * 4) >> (p ? f->g.ss_ver : 0); }
static size_t upload_row_bytes(const Rb200Frame *f, int p) { return (size_t)((f->g.bw * 4) >> (p ? f->g.ss_hor : 0)) * f->px; }

// Row ranges of one band (multi-GPU split of one picture by superblock rows).
// The band is named by the 64-row loop-restoration stripes [s0, s1) it must deliver; each earlier
// stage is widened by exactly what the next one reads: CDEF whole 64-row tiles, deblocked rows
// +-2 around them, row edges +-8 around those, column edges +-8 more.
struct BandRows {
    int s0, s1;           // stripes
    int t0, t1;           // CDEF tile rows
    int y4b, y4e;         // deblock row-edge unit rows (luma 4-px units)
    int in0, in1;         // reconstructed luma rows read
    int out0, out1;       // luma rows delivered
};
static BandRows band_rows(const Rb200Frame *f) {
    const int h = f->hdr.height;
    const int n_stripes = (h + 8 + 63) / 64, n_tiles = (f->g.bh * 4 + 63) / 64;
    BandRows b;
    const bool all = f->band_s1 <= f->band_s0;
    b.s0 = all ? 0 : imax(f->band_s0, 0);
    b.s1 = all ? n_stripes : imin(f->band_s1, n_stripes);
    b.out0 = imax(64 * b.s0 - 8, 0);
    b.out1 = b.s1 >= n_stripes ? h : 64 * b.s1 - 8;
    b.t0 = b.out0 / 64;
    b.t1 = imin((b.out1 + 63) / 64, n_tiles);
    if (b.s1 >= n_stripes) b.t1 = n_tiles;
    b.y4b = imax(16 * b.t0 - 3, 0);
    b.y4e = b.t1 >= n_tiles ? f->g.h4 : imin(16 * b.t1 + 3, f->g.h4);
    b.in0 = imax(4 * (b.y4b - 2), 0);
    b.in1 = b.y4e >= f->g.h4 ? f->g.plane_h[0] : imin(4 * (b.y4e + 2), f->g.plane_h[0]);
    return b;
}

extern "C" int rb200_frame_set_band(Rb200Frame *f, int stripe_begin, int stripe_end) {
    if (!f || stripe_begin < 0 || stripe_end < 0) return set_error(-22, "frame_set_band: bad argument");
    f->band_s0 = stripe_begin; f->band_s1 = stripe_end;
    return 0;
}

extern "C" int rb200_frame_band_rows(const Rb200Frame *f, int *in_begin, int *in_end, int *out_begin, int *out_end) {
    if (!f) return set_error(-22, "frame_band_rows: null frame");
    const BandRows b = band_rows(f);
    if (in_begin) *in_begin = b.in0;
    if (in_end) *in_end = b.in1;
    if (out_begin) *out_begin = b.out0;
    if (out_end) *out_end = b.out1;
    return 0;
}

// Copy luma rows [row_begin, row_end) (and the chroma rows they cover) between a host picture and a plane set.
static int copy_rows(Rb200Frame *f, const Rb200Planes &dev, void *const data[3], const ptrdiff_t stride[2], int row_begin,
                     int row_end, bool to_device) {
    for (int p = 0; p < f->g.n_planes; p++) {
        const int ss = p ? f->g.ss_ver : 0;
        const int full = to_device ? upload_rows(f, p) : (p ? (f->hdr.height + ss) >> ss : f->hdr.height);
        const int r0 = imax(row_begin >> ss, 0);
        // the band that owns the last picture row also owns the rows up to the 8-pixel boundary
        const int r1 = (to_device && row_end >= f->hdr.height) ? full : imin((row_end + ss) >> ss, full);
        if (r1 <= r0) continue;
        const ptrdiff_t hs = stride[p ? 1 : 0];
        if (hs < 0) return set_error(-22, "row-range copies need a positive host stride");
        const size_t rb = to_device ? upload_row_bytes(f, p)
                                    : (size_t)(p ? (f->hdr.width + f->g.ss_hor) >> f->g.ss_hor : f->hdr.width) * f->px;
        uint8_t *h = (uint8_t *)data[p] + (int64_t)r0 * hs;
        uint8_t *d = (uint8_t *)dev.data[p] + (int64_t)r0 * dev.stride[p];
        if (to_device) RB_CUDA(cudaMemcpy2DAsync(d, (size_t)dev.stride[p], h, (size_t)hs, rb, r1 - r0, cudaMemcpyHostToDevice, f->stream));
        else RB_CUDA(cudaMemcpy2DAsync(h, (size_t)hs, d, (size_t)dev.stride[p], rb, r1 - r0, cudaMemcpyDeviceToHost, f->stream));
    }
    return 0;
}

extern "C" int rb200_frame_upload_rows(Rb200Frame *f, int which, const void *const data[3], const ptrdiff_t stride[2],
                                       int row_begin, int row_end) {
    if (!f || which < 0 || which > 2 || !data || !stride) return set_error(-22, "frame_upload_rows: bad argument");
    const int r = copy_rows(f, f->planes[which], (void *const *)data, stride, row_begin, row_end, true);
    if (r) return r;
    RB_CUDA(cudaStreamSynchronize(f->stream));
    return 0;
}

extern "C" int rb200_frame_readback_rows(Rb200Frame *f, void *const data[3], const ptrdiff_t stride[2], int row_begin,
                                         int row_end) {
    if (!f || !data || !stride) return set_error(-22, "frame_readback_rows: bad argument");
    const int r = copy_rows(f, f->display, data, stride, row_begin, row_end, false);
    if (r) return r;
    RB_CUDA(cudaStreamSynchronize(f->stream));
    return 0;
}

// The device memory block behind plane set `which` (one allocation: Y, U, V back to back), for
// peer access from another process / device (rb200_ipc_*).
extern "C" int rb200_frame_plane_block(Rb200Frame *f, int which, void **base, size_t *bytes) {
    if (!f || which < 0 || which > 2 || !base || !bytes) return set_error(-22, "frame_plane_block: bad argument");
    const size_t ysz = (size_t)f->g.stride[0] * f->g.plane_h[0];
    const size_t uvsz = f->g.n_planes > 1 ? (size_t)f->g.stride[1] * f->g.plane_h[1] : 0;
    *base = f->plane_mem[which];
    *bytes = ysz + 2 * uvsz + 256;
    return 0;
}

// Pull luma rows [row_begin, row_end) (and their chroma rows) of plane set `which` from a peer copy of
// the same picture whose memory block starts at `peer_base` (same geometry) -- the halo exchange.
extern "C" int rb200_frame_pull_rows(Rb200Frame *f, int which, const void *peer_base, int row_begin, int row_end) {
    if (!f || which < 0 || which > 2 || !peer_base) return set_error(-22, "frame_pull_rows: bad argument");
    for (int p = 0; p < f->g.n_planes; p++) {
        const int ss = p ? f->g.ss_ver : 0;
        const int full = f->g.plane_h[p ? 1 : 0];
        const int r0 = imax(row_begin >> ss, 0), r1 = imin((row_end + ss) >> ss, full);
        if (r1 <= r0) continue;
        const int64_t off = (uint8_t *)f->planes[which].data[p] - f->plane_mem[which] + (int64_t)r0 * f->planes[which].stride[p];
        RB_CUDA(cudaMemcpyAsync(f->plane_mem[which] + off, (const uint8_t *)peer_base + off,
                                (size_t)(r1 - r0) * f->planes[which].stride[p], cudaMemcpyDefault, f->stream));
    }
    return 0;
}

extern "C" int rb200_frame_set_stream(Rb200Frame *f, void *stream) {
    if (!f) return set_error(-22, "frame_set_stream: null frame");
    RB_CUDA(cudaStreamSynchronize(f->stream));
    f->stream = stream ? (cudaStream_t)stream : f->own_stream;
    return 0;
}

extern "C" int16_t *rb200_frame_coef16_buffer(Rb200Frame *f) {
    if (!f || f->cs != 4) return nullptr;
    if (!f->h_coef16) {
        if (cudaMallocHost((void **)&f->h_coef16, (f->max_coefs ? f->max_coefs : 1) * sizeof(int16_t)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        memset(f->h_coef16, 0, (f->max_coefs ? f->max_coefs : 1) * sizeof(int16_t));
    }
    return f->h_coef16;
}
extern "C" int rb200_frame_reserve_coef_escapes(Rb200Frame *f, int max_escapes) {
    if (!f || max_escapes < 0) return set_error(-22, "frame_reserve_coef_escapes: bad argument");
    if (max_escapes <= f->max_esc) return 0;
    RB_CUDA(cudaStreamSynchronize(f->stream));
    if (f->up_stream) RB_CUDA(cudaStreamSynchronize(f->up_stream));
    Rb200CoefEscape *h = nullptr, *d = nullptr;
    RB_CUDA(cudaMallocHost((void **)&h, (size_t)max_escapes * sizeof(Rb200CoefEscape)));
    cudaError_t e = cudaMalloc((void **)&d, (size_t)max_escapes * sizeof(Rb200CoefEscape));
    if (e != cudaSuccess) { cudaFreeHost(h); return cuda_fail(e, "cudaMalloc(escapes)", __FILE__, __LINE__); }
    if (f->n_esc) memcpy(h, f->h_esc, (size_t)f->n_esc * sizeof(Rb200CoefEscape));
    if (f->h_esc) cudaFreeHost(f->h_esc);
    if (f->d_esc) cudaFree(f->d_esc);
    f->h_esc = h; f->d_esc = d; f->max_esc = max_escapes;
    return 0;
}
extern "C" Rb200CoefEscape *rb200_frame_coef_escapes(Rb200Frame *f) { return f ? f->h_esc : nullptr; }
extern "C" int rb200_frame_set_coef_escape_count(Rb200Frame *f, int n) {
    if (!f || n < 0 || n > f->max_esc) return set_error(-22, "frame_set_coef_escape_count: bad count");
    f->n_esc = n;
    return 0;
}
extern "C" int rb200_frame_pack_coef16(Rb200Frame *f, size_t n_coefs) {
    if (!f || n_coefs > f->max_coefs) return set_error(-22, "frame_pack_coef16: bad argument");
    if (f->cs != 4) return set_error(-22, "frame_pack_coef16: 8-bit pictures carry int16 coefficients already");
    int16_t *o = rb200_frame_coef16_buffer(f);
    if (!o) return set_error(-12, "frame_pack_coef16: out of pinned memory");
    const int32_t *c = (const int32_t *)f->h_coef;
    f->n_esc = 0;
    for (size_t i = 0; i < n_coefs; i++) {
        const int32_t v = c[i];
        o[i] = (int16_t)v;
        if (v != (int16_t)v) {
            if (f->n_esc == f->max_esc) {
                const int r = rb200_frame_reserve_coef_escapes(f, f->max_esc ? 2 * f->max_esc : 4096);
                if (r) return r;
            }
            f->h_esc[f->n_esc].index = (uint32_t)i; f->h_esc[f->n_esc].value = v; f->n_esc++;
        }
    }
    return 0;
}

static int ensure_coef_stream(Rb200Frame *f) {
    if (f->h_pk) return 0;
    const size_t n = f->max_coefs + 8 * (size_t)f->max_itx + 8;
    RB_CUDA(cudaMallocHost((void **)&f->h_pk, n * sizeof(int16_t)));
    RB_CUDA(cudaMalloc((void **)&f->d_pk, n * sizeof(int16_t)));
    RB_CUDA(cudaMallocHost((void **)&f->h_pkoff, ((size_t)f->max_itx + 1) * sizeof(uint32_t)));
    RB_CUDA(cudaMalloc((void **)&f->d_pkoff, ((size_t)f->max_itx + 1) * sizeof(uint32_t)));
    return 0;
}
extern "C" int16_t *rb200_frame_coef_stream(Rb200Frame *f) { return f && f->cs == 4 && !ensure_coef_stream(f) ? f->h_pk : nullptr; }
extern "C" uint32_t *rb200_frame_coef_stream_offsets(Rb200Frame *f) { return f && f->cs == 4 && !ensure_coef_stream(f) ? f->h_pkoff : nullptr; }
extern "C" int rb200_frame_set_coef_stream_length(Rb200Frame *f, size_t n) {
    if (!f || n > f->max_coefs + 8 * (size_t)f->max_itx) return set_error(-22, "frame_set_coef_stream_length: bad length");
    f->n_pk = n;
    return 0;
}
extern "C" int rb200_frame_pack_coef_stream(Rb200Frame *f, size_t n_coefs, const int32_t itx_counts[RB200_N_RECT_TX_SIZES], int stages) {
    if (!f || !itx_counts || n_coefs > f->max_coefs) return set_error(-22, "frame_pack_coef_stream: bad argument");
    if (f->cs != 4) return set_error(-22, "frame_pack_coef_stream: 8-bit pictures carry int16 coefficients already");
    int r = ensure_coef_stream(f);
    if (r) return r;
    int n_itx = 0;
    for (int t = 0; t < RB200_N_RECT_TX_SIZES; t++) n_itx += itx_counts[t];
    if (stages & RB200_STAGE_INTRA)
        for (int l = 0; l < f->n_levels; l++)
            for (int t = 0; t < RB200_N_RECT_TX_SIZES; t++) n_itx += f->intra_itx_counts[l * RB200_N_RECT_TX_SIZES + t];
    if (n_itx > f->max_itx) return set_error(-22, "frame_pack_coef_stream: more residual items than the frame holds");
    static const uint8_t txw[RB200_N_RECT_TX_SIZES] = {4, 8, 16, 32, 64, 4, 8, 8, 16, 16, 32, 32, 64, 4, 16, 8, 32, 16, 64};
    static const uint8_t txh[RB200_N_RECT_TX_SIZES] = {4, 8, 16, 32, 64, 8, 4, 16, 8, 32, 16, 64, 32, 16, 4, 32, 8, 64, 16};
    const int32_t *c = (const int32_t *)f->h_coef;
    size_t pos = 0;
    f->n_esc = 0;
    for (int i = 0; i < n_itx; i++) {
        const Rb200ItxItem &it = f->h_itx[i];
        if (it.tx >= RB200_N_RECT_TX_SIZES) return set_error(-22, "frame_pack_coef_stream: item %d has transform size %d", i, it.tx);
        const int sw = txw[it.tx] < 32 ? txw[it.tx] : 32, sh = txh[it.tx] < 32 ? txh[it.tx] : 32;
        const int nc = it.ncols && it.ncols < sw ? it.ncols : sw, count = nc * sh;
        if ((size_t)it.cf_off + count > n_coefs) return set_error(-22, "frame_pack_coef_stream: item %d reads beyond the coefficients", i);
        f->h_pkoff[i] = (uint32_t)pos;
        for (int k = 0; k < count; k++) {
            const int32_t v = c[it.cf_off + k];
            f->h_pk[pos + k] = (int16_t)v;
            if (v != (int16_t)v) {
                if (f->n_esc == f->max_esc && (r = rb200_frame_reserve_coef_escapes(f, f->max_esc ? 2 * f->max_esc : 4096))) return r;
                f->h_esc[f->n_esc].index = it.cf_off + (uint32_t)k; f->h_esc[f->n_esc].value = v; f->n_esc++;
            }
        }
        for (int k = count; k & 7; k++) f->h_pk[pos + k] = 0;
        pos += (size_t)(count + 7) & ~(size_t)7;
    }
    f->n_pk = pos;
    return 0;
}

extern "C" int rb200_frame_depend(Rb200Frame *f, Rb200Frame *producer) {
    if (!f || !producer) return set_error(-22, "frame_depend: null frame");
    if (f == producer || f->stream == producer->stream) return 0;   // one stream: already ordered
    if (f->n_deps >= 8) return set_error(-22, "frame_depend: more than 8 producers before a submit");
    f->dep_events[f->n_deps++] = producer->done_event;
    return 0;
}

extern "C" int rb200_frame_set_plane_streams(Rb200Frame *f, int on) {
    if (!f) return set_error(-22, "frame_set_plane_streams: null frame");
    f->plane_split = on != 0;
    return 0;
}

extern "C" int rb200_frame_set_plane_counts(Rb200Frame *f, int n_mc_luma, const int32_t itx_luma_counts[RB200_N_RECT_TX_SIZES]) {
    if (!f) return set_error(-22, "frame_set_plane_counts: null frame");
    f->plane_counts = n_mc_luma >= 0 && itx_luma_counts;
    if (f->plane_counts) {
        f->n_mc_luma = n_mc_luma;
        for (int t = 0; t < RB200_N_RECT_TX_SIZES; t++) {
            if (itx_luma_counts[t] < 0) { f->plane_counts = false; return set_error(-22, "frame_set_plane_counts: negative count"); }
            f->itx_luma[t] = itx_luma_counts[t];
        }
    }
    return 0;
}

extern "C" int rb200_frame_enable_timing(Rb200Frame *f, int on) {
    if (!f) return set_error(-22, "frame_enable_timing: null frame");
    if (on)
        for (int i = 0; i < RB200_N_FRAME_MARKS; i++)
            if (!f->ev[i]) RB_CUDA(cudaEventCreate(&f->ev[i]));
    f->timing = on != 0;
    return 0;
}

// ms between consecutive marks of the last submit: [0] H2D, [1] MC, [2] itx, [3] deblock, [4] CDEF, [5] LR, [6] film grain.
// A stage that did not run reports 0.  Valid after rb200_frame_wait().
extern "C" int rb200_frame_stage_times(Rb200Frame *f, float ms[RB200_N_FRAME_MARKS - 1]) {
    if (!f || !ms) return set_error(-22, "frame_stage_times: bad argument");
    if (!f->timing) return set_error(-22, "frame_stage_times: timing not enabled");
    for (int i = 0; i + 1 < RB200_N_FRAME_MARKS; i++) {
        ms[i] = 0.f;
        if (f->ev_valid[i] && f->ev_valid[i + 1]) RB_CUDA(cudaEventElapsedTime(&ms[i], f->ev[i], f->ev[i + 1]));
    }
    return 0;
}

extern "C" int rb200_frame_reserve_comp_items(Rb200Frame *f, int max_comp) {
    if (!f || max_comp < 0) return set_error(-22, "frame_reserve_comp_items: bad argument");
    if (max_comp <= f->max_comp) return 0;
    RB_CUDA(cudaStreamSynchronize(f->stream));
    if (f->h_comp) cudaFreeHost(f->h_comp);
    if (f->d_comp) cudaFree(f->d_comp);
    f->h_comp = nullptr; f->d_comp = nullptr; f->max_comp = 0; f->n_comp = 0;
    const int r = alloc_pair(&f->h_comp, &f->d_comp, (size_t)max_comp);
    if (r) return r;
    f->max_comp = max_comp;
    return 0;
}
extern "C" Rb200CompItem *rb200_frame_comp_items(Rb200Frame *f) { return f ? f->h_comp : nullptr; }
extern "C" int rb200_frame_set_comp_count(Rb200Frame *f, int n) {
    if (!f || n < 0 || n > f->max_comp) return set_error(-22, "frame_set_comp_count: more items than reserved");
    f->n_comp = n;
    return 0;
}

extern "C" int rb200_frame_reserve_warp_items(Rb200Frame *f, int max_warp) {
    if (!f || max_warp < 0) return set_error(-22, "frame_reserve_warp_items: bad argument");
    if (max_warp <= f->max_warp) return 0;
    RB_CUDA(cudaStreamSynchronize(f->stream));
    if (f->h_warp) cudaFreeHost(f->h_warp);
    if (f->d_warp) cudaFree(f->d_warp);
    f->h_warp = nullptr; f->d_warp = nullptr; f->max_warp = 0; f->n_warp = 0;
    const int r = alloc_pair(&f->h_warp, &f->d_warp, (size_t)max_warp);
    if (r) return r;
    f->max_warp = max_warp;
    return 0;
}
extern "C" Rb200WarpItem *rb200_frame_warp_items(Rb200Frame *f) { return f ? f->h_warp : nullptr; }
extern "C" int rb200_frame_set_warp_count(Rb200Frame *f, int n) {
    if (!f || n < 0 || n > f->max_warp) return set_error(-22, "frame_set_warp_count: more items than reserved");
    f->n_warp = n;
    return 0;
}

extern "C" int rb200_frame_reserve_scaled_items(Rb200Frame *f, int max_scaled) {
    if (!f || max_scaled < 0) return set_error(-22, "frame_reserve_scaled_items: bad argument");
    if (max_scaled <= f->max_scaled) return 0;
    RB_CUDA(cudaStreamSynchronize(f->stream));
    if (f->h_scaled) cudaFreeHost(f->h_scaled);
    if (f->d_scaled) cudaFree(f->d_scaled);
    f->h_scaled = nullptr; f->d_scaled = nullptr; f->max_scaled = 0; f->n_scaled = 0;
    const int r = alloc_pair(&f->h_scaled, &f->d_scaled, (size_t)max_scaled);
    if (r) return r;
    f->max_scaled = max_scaled;
    return 0;
}
extern "C" Rb200McScaledItem *rb200_frame_scaled_items(Rb200Frame *f) { return f ? f->h_scaled : nullptr; }
extern "C" int rb200_frame_set_scaled_count(Rb200Frame *f, int n) {
    if (!f || n < 0 || n > f->max_scaled) return set_error(-22, "frame_set_scaled_count: more items than reserved");
    f->n_scaled = n;
    f->n_scaled_above = f->n_scaled_left = 0;
    return 0;
}
extern "C" int rb200_frame_set_ref_gmv(Rb200Frame *f, int slot, const int32_t matrix[6], const int16_t abcd[4]) {
    if (!f || slot < 0 || slot > 7 || !matrix || !abcd) return set_error(-22, "frame_set_ref_gmv: bad argument");
    for (int k = 0; k < 6; k++) f->ref_gmv.g[slot].matrix[k] = matrix[k];
    for (int k = 0; k < 4; k++) f->ref_gmv.g[slot].abcd[k] = abcd[k];
    return 0;
}

extern "C" int rb200_frame_set_scaled_obmc_counts(Rb200Frame *f, int n_above, int n_left) {
    if (!f || n_above < 0 || n_left < 0 || f->n_scaled + n_above + n_left > f->max_scaled)
        return set_error(-22, "frame_set_scaled_obmc_counts: more items than reserved");
    f->n_scaled_above = n_above; f->n_scaled_left = n_left;
    return 0;
}

extern "C" int rb200_frame_set_ref_size(Rb200Frame *f, int slot, int width, int height) {
    if (!f || slot < 0 || slot > 7 || width < 1 || height < 1) return set_error(-22, "frame_set_ref_size: bad argument");
    f->ref_dims.w[slot] = width; f->ref_dims.h[slot] = height;
    return 0;
}

extern "C" int rb200_frame_reserve_intra_items(Rb200Frame *f, int max_items, int max_levels) {
    if (!f || max_items < 0 || max_levels < 0) return set_error(-22, "frame_reserve_intra_items: bad argument");
    if (max_items > f->max_intra) {
        RB_CUDA(cudaStreamSynchronize(f->stream));
        if (f->h_intra) cudaFreeHost(f->h_intra);
        if (f->d_intra) cudaFree(f->d_intra);
        if (f->h_intra_itx) cudaFreeHost(f->h_intra_itx);
        if (f->d_intra_itx) cudaFree(f->d_intra_itx);
        f->h_intra = nullptr; f->d_intra = nullptr; f->h_intra_itx = nullptr; f->d_intra_itx = nullptr; f->max_intra = 0; f->n_levels = 0;
        int r = alloc_pair(&f->h_intra, &f->d_intra, (size_t)max_items);
        if (!r) r = alloc_pair(&f->h_intra_itx, &f->d_intra_itx, (size_t)max_items);
        if (r) return r;
        for (int i = 0; i < max_items; i++) f->h_intra_itx[i] = -1;
        f->max_intra = max_items;
    }
    if (max_levels > f->max_levels) {
        free(f->intra_counts); free(f->intra_itx_counts);
        f->intra_counts = nullptr; f->intra_itx_counts = nullptr; f->max_levels = 0; f->n_levels = 0;
        RB_CUDA(cudaStreamSynchronize(f->stream));
        if (f->h_level_off) cudaFreeHost(f->h_level_off);
        if (f->d_level_off) cudaFree(f->d_level_off);
        f->h_level_off = nullptr; f->d_level_off = nullptr;
        int r = alloc_pair(&f->h_level_off, &f->d_level_off, (size_t)max_levels + 1);
        if (f->h_intra_sync) cudaFreeHost(f->h_intra_sync);
        if (f->d_intra_sync) cudaFree(f->d_intra_sync);
        f->h_intra_sync = nullptr; f->d_intra_sync = nullptr;
        if (!r) r = alloc_pair(&f->h_intra_sync, &f->d_intra_sync, (size_t)max_levels + 1);
        if (r) return r;
        f->intra_counts = (int32_t *)calloc((size_t)max_levels, sizeof(int32_t));
        f->intra_itx_counts = (int32_t *)calloc((size_t)max_levels * RB200_N_RECT_TX_SIZES, sizeof(int32_t));
        if (!f->intra_counts || !f->intra_itx_counts) return set_error(-12, "frame_reserve_intra_items: out of memory");
        f->max_levels = max_levels; f->n_levels = 0;
    }
    return 0;
}
extern "C" Rb200IntraItem *rb200_frame_intra_items(Rb200Frame *f) { return f ? f->h_intra : nullptr; }
extern "C" int32_t *rb200_frame_intra_itx_index(Rb200Frame *f) { return f ? f->h_intra_itx : nullptr; }
extern "C" int rb200_frame_reserve_palette(Rb200Frame *f, size_t bytes) {
    if (!f) return set_error(-22, "frame_reserve_palette: null frame");
    if (bytes <= f->max_pal) return 0;
    RB_CUDA(cudaStreamSynchronize(f->stream));
    if (f->h_pal) cudaFreeHost(f->h_pal);
    if (f->d_pal) cudaFree(f->d_pal);
    f->h_pal = nullptr; f->d_pal = nullptr; f->max_pal = 0; f->n_pal = 0;
    const int r = alloc_pair(&f->h_pal, &f->d_pal, bytes);
    if (r) return r;
    f->max_pal = bytes;
    return 0;
}
extern "C" uint8_t *rb200_frame_palette_buffer(Rb200Frame *f) { return f ? f->h_pal : nullptr; }
extern "C" int rb200_frame_set_palette_bytes(Rb200Frame *f, size_t bytes) {
    if (!f || bytes > f->max_pal) return set_error(-22, "frame_set_palette_bytes: more than reserved");
    f->n_pal = bytes;
    return 0;
}
extern "C" int rb200_frame_set_intra_levels(Rb200Frame *f, int n_levels, const int32_t *item_counts, const int32_t *itx_counts) {
    if (!f || n_levels < 0 || n_levels > f->max_levels || (n_levels && (!item_counts || !itx_counts)))
        return set_error(-22, "frame_set_intra_levels: more levels than reserved");
    int64_t n = 0;
    for (int l = 0; l < n_levels; l++) {
        if (item_counts[l] < 0) return set_error(-22, "frame_set_intra_levels: negative count");
        n += item_counts[l];
    }
    if (n > f->max_intra) return set_error(-22, "frame_set_intra_levels: more items than reserved");
    if (n_levels) {
        memcpy(f->intra_counts, item_counts, (size_t)n_levels * sizeof(int32_t));
        memcpy(f->intra_itx_counts, itx_counts, (size_t)n_levels * RB200_N_RECT_TX_SIZES * sizeof(int32_t));
    }
    // item offsets of the non-empty levels (the one-launch wavefront chains each level to the one before it)
    f->intra_widest = 0; f->n_levels_nonempty = 0;
    if (f->h_level_off) f->h_level_off[0] = 0;
    for (int l = 0; l < n_levels; l++) {
        if (!item_counts[l]) continue;
        const int k = f->n_levels_nonempty++;
        f->h_level_off[k + 1] = f->h_level_off[k] + item_counts[l];
        f->intra_widest = imax(f->intra_widest, item_counts[l]);
    }
    f->n_levels = n_levels;
    return 0;
}

extern "C" int rb200_frame_reserve_lf_blocks(Rb200Frame *f, int max_blocks) {
    if (!f || max_blocks < 0) return set_error(-22, "frame_reserve_lf_blocks: bad argument");
    if (max_blocks <= f->max_lfb) return 0;
    RB_CUDA(cudaStreamSynchronize(f->stream));
    if (f->h_lfb) cudaFreeHost(f->h_lfb);
    if (f->d_lfb) cudaFree(f->d_lfb);
    f->h_lfb = nullptr; f->d_lfb = nullptr; f->max_lfb = 0; f->n_lfb = 0;
    int r = alloc_pair(&f->h_lfb, &f->d_lfb, (size_t)max_blocks);
    if (!r && !f->d_lf_cells) {
        RB_CUDA(cudaMalloc((void **)&f->d_lf_cells, (size_t)2 * 32 * f->g.sb128w * 32 * f->g.sb128h));
        r = alloc_pair(&f->h_cdef_idx, &f->d_cdef_idx, f->n_masks * 4);
        RB_CUDA(cudaStreamCreateWithFlags(&f->lf_stream, cudaStreamNonBlocking));
        RB_CUDA(cudaEventCreateWithFlags(&f->lf_fork, cudaEventDisableTiming));
        RB_CUDA(cudaEventCreateWithFlags(&f->lf_join, cudaEventDisableTiming));
    }
    if (r) return r;
    f->max_lfb = max_blocks;
    return 0;
}
extern "C" Rb200LfBlock *rb200_frame_lf_blocks(Rb200Frame *f) { return f ? f->h_lfb : nullptr; }
extern "C" int rb200_frame_set_lf_block_count(Rb200Frame *f, int n) {
    if (!f || n < 0 || n > f->max_lfb) return set_error(-22, "frame_set_lf_block_count: more blocks than reserved");
    f->n_lfb = n;
    return 0;
}
extern "C" int rb200_frame_download_lf(Rb200Frame *f, Rb200Av1Filter *masks, uint8_t (*levels)[4]) {
    if (!f) return set_error(-22, "frame_download_lf: null frame");
    RB_CUDA(cudaStreamSynchronize(f->stream));
    if (masks) RB_CUDA(cudaMemcpy(masks, f->d_masks, (size_t)f->g.sb128w * f->g.sb128h * sizeof(Rb200Av1Filter), cudaMemcpyDeviceToHost));
    if (levels) RB_CUDA(cudaMemcpy(levels, f->d_lvl + 32, (f->n_lvl - 32) * 4, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int rb200_frame_reserve_obmc_items(Rb200Frame *f, int max_obmc) {
    if (!f || max_obmc < 0) return set_error(-22, "frame_reserve_obmc_items: bad argument");
    if (max_obmc <= f->max_obmc) return 0;
    RB_CUDA(cudaStreamSynchronize(f->stream));
    if (f->h_obmc) cudaFreeHost(f->h_obmc);
    if (f->d_obmc) cudaFree(f->d_obmc);
    f->h_obmc = nullptr; f->d_obmc = nullptr; f->max_obmc = 0; f->n_obmc_above = f->n_obmc_left = 0;
    const int r = alloc_pair(&f->h_obmc, &f->d_obmc, (size_t)max_obmc);
    if (r) return r;
    f->max_obmc = max_obmc;
    return 0;
}
extern "C" Rb200McItem *rb200_frame_obmc_items(Rb200Frame *f) { return f ? f->h_obmc : nullptr; }
extern "C" int rb200_frame_set_obmc_counts(Rb200Frame *f, int n_above, int n_left) {
    if (!f || n_above < 0 || n_left < 0 || n_above + n_left > f->max_obmc)
        return set_error(-22, "frame_set_obmc_counts: more items than reserved");
    f->n_obmc_above = n_above; f->n_obmc_left = n_left;
    return 0;
}

extern "C" int rb200_frame_set_ref(Rb200Frame *f, int slot, const Rb200Planes *planes) {
    if (!f || slot < 0 || slot > 7 || !planes) return set_error(-22, "frame_set_ref: bad argument");
    f->refs[slot] = *planes;
    f->ref_dims.w[slot] = f->hdr.width; f->ref_dims.h[slot] = f->hdr.height;   // same size unless rb200_frame_set_ref_size says otherwise
    if (slot + 1 > f->n_refs) f->n_refs = slot + 1;
    return 0;
}

extern "C" int rb200_frame_stage_planes(Rb200Frame *f, int which, Rb200Planes *out) {
    if (!f || which < 0 || which > 2 || !out) return set_error(-22, "frame_stage_planes: bad argument");
    *out = f->planes[which];
    return 0;
}
extern "C" int rb200_frame_output_planes(Rb200Frame *f, Rb200Planes *out) {
    if (!f || !out) return set_error(-22, "frame_output_planes: bad argument");
    *out = f->out;
    return 0;
}

static int plane_rows(const Rb200Frame *f, int p) {
    return p ? (f->hdr.height + f->g.ss_ver) >> f->g.ss_ver : f->hdr.height;
}
static size_t plane_row_bytes(const Rb200Frame *f, int p) {   // of the output picture
    return (size_t)(p ? (f->out_w + f->g.ss_hor) >> f->g.ss_hor : f->out_w) * f->px;
}

extern "C" int rb200_frame_upload_planes(Rb200Frame *f, int which, const void *const data[3],
                                         const ptrdiff_t stride[2]) {
    if (!f || which < 0 || which > 2 || !data || !stride) return set_error(-22, "frame_upload_planes: bad argument");
    for (int p = 0; p < f->g.n_planes; p++) {
        const int rows = upload_rows(f, p);
        const ptrdiff_t hs = stride[p ? 1 : 0];
        const uint8_t *src = (const uint8_t *)data[p];
        if (hs < 0) src += (int64_t)(rows - 1) * hs;  // lowest address; copy flipped below
        if (hs >= 0) {
            RB_CUDA(cudaMemcpy2DAsync(f->planes[which].data[p], (size_t)f->planes[which].stride[p], src, (size_t)hs,
                                      upload_row_bytes(f, p), rows, cudaMemcpyHostToDevice, f->stream));
        } else {
            for (int y = 0; y < rows; y++)
                RB_CUDA(cudaMemcpyAsync((uint8_t *)f->planes[which].data[p] + (int64_t)y * f->planes[which].stride[p],
                                        (const uint8_t *)data[p] + (int64_t)y * hs, upload_row_bytes(f, p),
                                        cudaMemcpyHostToDevice, f->stream));
        }
    }
    RB_CUDA(cudaStreamSynchronize(f->stream));
    return 0;
}

extern "C" int rb200_frame_readback_async(Rb200Frame *f, void *const data[3], const ptrdiff_t stride[2]) {
    if (!f || !data || !stride) return set_error(-22, "frame_readback: bad argument");
    for (int p = 0; p < f->g.n_planes; p++) {
        const int rows = plane_rows(f, p);
        const ptrdiff_t hs = stride[p ? 1 : 0];
        if (hs >= 0 && (size_t)hs == plane_row_bytes(f, p) && (int64_t)hs == f->display.stride[p]) {
            // rows back to back on both sides: one linear copy
            RB_CUDA(cudaMemcpyAsync(data[p], f->display.data[p], (size_t)hs * rows, cudaMemcpyDeviceToHost, f->stream));
        } else if (hs >= 0) {
            RB_CUDA(cudaMemcpy2DAsync(data[p], (size_t)hs, f->display.data[p], (size_t)f->display.stride[p],
                                      plane_row_bytes(f, p), rows, cudaMemcpyDeviceToHost, f->stream));
        } else {
            for (int y = 0; y < rows; y++)
                RB_CUDA(cudaMemcpyAsync((uint8_t *)data[p] + (int64_t)y * hs,
                                        (const uint8_t *)f->display.data[p] + (int64_t)y * f->display.stride[p],
                                        plane_row_bytes(f, p), cudaMemcpyDeviceToHost, f->stream));
        }
    }
    return 0;
}

extern "C" int rb200_frame_readback(Rb200Frame *f, void *const data[3], const ptrdiff_t stride[2]) {
    const int r = rb200_frame_readback_async(f, data, stride);
    if (r) return r;
    RB_CUDA(cudaStreamSynchronize(f->stream));
    return 0;
}

extern "C" int rb200_frame_wait(Rb200Frame *f) {
    if (!f) return set_error(-22, "frame_wait: null frame");
    RB_CUDA(cudaStreamSynchronize(f->stream));
    RB_CUDA(cudaGetLastError());
    if (f->intra_check) {       // the one-launch intra wavefront reports a wait that ran into its time limit
        f->intra_check = false;
        RB_CUDA(cudaMemcpy(f->h_intra_sync, f->d_intra_sync, sizeof(unsigned), cudaMemcpyDeviceToHost));
        if (f->h_intra_sync[0]) return set_error(-110, "frame_wait: the intra wavefront timed out waiting for a level");
    }
    return 0;
}

extern "C" int rb200_frame_validate(Rb200Frame *f, size_t n_coefs, const int32_t itx_counts[RB200_N_RECT_TX_SIZES], int n_mc, int stages) {
    if (!f) return set_error(-22, "frame_validate: null frame");
    if (!(stages & RB200_STAGE_RECON)) return 0;
    if (!itx_counts) return set_error(-22, "frame_validate: itx_counts required");
    const Rb200FrameGeometry &g = f->g;
    static const uint8_t txw[RB200_N_RECT_TX_SIZES] = {4, 8, 16, 32, 64, 4, 8, 8, 16, 16, 32, 32, 64, 4, 16, 8, 32, 16, 64};
    static const uint8_t txh[RB200_N_RECT_TX_SIZES] = {4, 8, 16, 32, 64, 8, 4, 16, 8, 32, 16, 64, 32, 16, 4, 32, 8, 64, 16};
    auto plane_w = [&](int p) { return (int)(g.stride[p ? 1 : 0] / (int64_t)f->px); };
    auto plane_h = [&](int p) { return g.plane_h[p ? 1 : 0]; };
    if (n_coefs > f->max_coefs) return set_error(-22, "frame_validate: %zu coefficients, the frame holds %zu", n_coefs, f->max_coefs);
    int n_itx = 0, idx = 0;
    for (int t = 0; t < RB200_N_RECT_TX_SIZES; t++) {
        if (itx_counts[t] < 0) return set_error(-22, "frame_validate: negative itx count");
        n_itx += itx_counts[t];
    }
    int n_itx_all = n_itx;
    if (stages & RB200_STAGE_INTRA)
        for (int l = 0; l < f->n_levels; l++)
            for (int t = 0; t < RB200_N_RECT_TX_SIZES; t++) n_itx_all += f->intra_itx_counts[l * RB200_N_RECT_TX_SIZES + t];
    if (n_itx_all > f->max_itx) return set_error(-22, "frame_validate: %d residual items, the frame holds %d", n_itx_all, f->max_itx);
    for (int t = 0; t < RB200_N_RECT_TX_SIZES; t++)
        for (int k = 0; k < itx_counts[t]; k++, idx++) {
            const Rb200ItxItem &it = f->h_itx[idx];
            const int sw = txw[t] < 32 ? txw[t] : 32, sh = txh[t] < 32 ? txh[t] : 32;
            if (it.tx != t) return set_error(-22, "frame_validate: itx item %d is in the bucket of size %d but says %d", idx, t, it.tx);
            if (it.plane >= g.n_planes || !rb200_itx_valid(it.tx, it.txtp)) return set_error(-22, "frame_validate: itx item %d: plane %d / type %d", idx, it.plane, it.txtp);
            if ((size_t)it.cf_off + (size_t)sw * sh > n_coefs) return set_error(-22, "frame_validate: itx item %d reads coefficients %u.. beyond %zu", idx, it.cf_off, n_coefs);
            if (it.x + txw[t] > plane_w(it.plane) || it.y + txh[t] > plane_h(it.plane)) return set_error(-22, "frame_validate: itx item %d (%d, %d) leaves plane %d", idx, it.x, it.y, it.plane);
            if (it.ncols > sw) return set_error(-22, "frame_validate: itx item %d: ncols %d of %d", idx, it.ncols, sw);
        }
    for (int i = n_itx; i < n_itx_all; i++) {
        const Rb200ItxItem &it = f->h_itx[i];
        if (it.tx >= RB200_N_RECT_TX_SIZES || it.plane >= g.n_planes || !rb200_itx_valid(it.tx, it.txtp)) return set_error(-22, "frame_validate: intra residual %d: size %d / plane %d / type %d", i, it.tx, it.plane, it.txtp);
        const int sw = txw[it.tx] < 32 ? txw[it.tx] : 32, sh = txh[it.tx] < 32 ? txh[it.tx] : 32;
        if ((size_t)it.cf_off + (size_t)sw * sh > n_coefs) return set_error(-22, "frame_validate: intra residual %d reads coefficients beyond %zu", i, n_coefs);
        if (it.x + txw[it.tx] > plane_w(it.plane) || it.y + txh[it.tx] > plane_h(it.plane)) return set_error(-22, "frame_validate: intra residual %d leaves plane %d", i, it.plane);
    }
    if (n_mc < 0 || n_mc > f->max_mc) return set_error(-22, "frame_validate: %d prediction items, the frame holds %d", n_mc, f->max_mc);
    auto check_mc = [&](const Rb200McItem &m, int i, const char *list) -> int {
        if (m.plane >= g.n_planes || m.ref >= f->n_refs || !f->refs[m.ref].data[m.plane]) return set_error(-22, "frame_validate: %s item %d: plane %d, reference slot %d (%d set)", list, i, m.plane, m.ref, f->n_refs);
        if (m.w < 2 || m.h < 2 || m.w > 128 || m.h > 128 || m.mx > 15 || m.my > 15 || m.filter2d > 9) return set_error(-22, "frame_validate: %s item %d: %d x %d, phase (%d, %d), filter %d", list, i, m.w, m.h, m.mx, m.my, m.filter2d);
        if (m.dst_x < 0 || m.dst_y < 0 || m.dst_x + m.w > plane_w(m.plane) || m.dst_y + m.h > plane_h(m.plane)) return set_error(-22, "frame_validate: %s item %d (%d, %d) %d x %d leaves plane %d", list, i, m.dst_x, m.dst_y, m.w, m.h, m.plane);
        return 0;
    };
    int r;
    for (int i = 0; i < n_mc; i++) if ((r = check_mc(f->h_mc[i], i, "mc"))) return r;
    for (int i = 0; i < f->n_obmc_above + f->n_obmc_left; i++) if ((r = check_mc(f->h_obmc[i], i, "obmc"))) return r;
    if (stages & RB200_STAGE_INTRA) {
        int n_in = 0;
        for (int l = 0; l < f->n_levels; l++) n_in += f->intra_counts[l];
        for (int i = 0; i < n_in; i++) {
            const Rb200IntraItem &it = f->h_intra[i];
            if (it.plane >= g.n_planes) return set_error(-22, "frame_validate: intra item %d: plane %d", i, it.plane);
            if (it.mode > 16 || it.tw4 < 1 || it.tw4 > 16 || it.th4 < 1 || it.th4 > 16)
                return set_error(-22, "frame_validate: intra item %d: mode %d, size %d x %d", i, it.mode, it.tw4, it.th4);
            if (it.x4 * 4 >= plane_w(it.plane) || it.y4 * 4 >= plane_h(it.plane))      // the block may hang over, its origin may not
                return set_error(-22, "frame_validate: intra item %d starts outside plane %d", i, it.plane);
            if (it.level >= f->n_levels) return set_error(-22, "frame_validate: intra item %d: level %d of %d", i, it.level, f->n_levels);
            if (f->h_intra_itx[i] >= n_itx_all || f->h_intra_itx[i] < -1) return set_error(-22, "frame_validate: intra item %d names residual %d of %d", i, f->h_intra_itx[i], n_itx_all);
        }
    }
    return 0;
}

extern "C" int rb200_frame_submit(Rb200Frame *f, size_t n_coefs, const int32_t itx_counts[RB200_N_RECT_TX_SIZES],
                                  int n_mc, int stages, int upload) {
    if (!f) return set_error(-22, "frame_submit: null frame");
    const Rb200FrameHeader &h = f->hdr;
    const Rb200FrameGeometry &g = f->g;
    cudaStream_t st = f->stream;
    int n_itx = 0;
    if (stages & RB200_STAGE_RECON) {
        if (!itx_counts) return set_error(-22, "frame_submit: itx_counts required for the recon stage");
        for (int t = 0; t < RB200_N_RECT_TX_SIZES; t++) {
            if (itx_counts[t] < 0) return set_error(-22, "frame_submit: negative itx count");
            n_itx += itx_counts[t];
        }
        if (stages & RB200_STAGE_INTRA)   // the residuals of the intra blocks follow the inter ones in the same list
            for (int l = 0; l < f->n_levels; l++)
                for (int t = 0; t < RB200_N_RECT_TX_SIZES; t++) n_itx += f->intra_itx_counts[l * RB200_N_RECT_TX_SIZES + t];
        if (n_coefs > f->max_coefs || n_itx > f->max_itx || n_mc > f->max_mc || n_mc < 0)
            return set_error(-22, "frame_submit: batch larger than the frame was created for");
        if ((n_mc || f->n_comp || f->n_warp || f->n_scaled || f->n_scaled_above || f->n_scaled_left || f->n_obmc_above || f->n_obmc_left) && f->n_refs < 1) return set_error(-22, "frame_submit: no reference picture set");
    }
    f->launches = 0;
    const BandRows band = band_rows(f);
    for (int i = 0; i < RB200_N_FRAME_MARKS; i++) f->ev_valid[i] = false;
#define MARK(i) do { if (f->timing) { RB_CUDA(cudaEventRecord(f->ev[i], st)); f->ev_valid[i] = true; } } while (0)
    MARK(0);
    if (stages & RB200_STAGE_FILM_GRAIN) {
        if (!f->fg_set) return set_error(-22, "frame_submit: RB200_STAGE_FILM_GRAIN without rb200_frame_set_film_grain");
        // fork: the previous frame's grain application (on st) still reads the tables prepared here
        RB_CUDA(cudaEventRecord(f->fg_fork, st));
        RB_CUDA(cudaStreamWaitEvent(f->fg_stream, f->fg_fork, 0));
        int r0;
        if ((r0 = film_grain_prepare(f, f->fg_stream, f->sr && (stages & RB200_STAGE_SUPER_RES) ? f->sr_w : f->hdr.width))) return r0;
        RB_CUDA(cudaEventRecord(f->fg_join, f->fg_stream));
    }
    const bool do_lf = (stages & RB200_STAGE_DEBLOCK) && (h.lf_level_y[0] || h.lf_level_y[1]);
    const bool do_cdef = (stages & RB200_STAGE_CDEF) != 0;
    const bool do_sr = f->sr && (stages & RB200_STAGE_SUPER_RES);
    int restore_planes = 0;
    if (stages & RB200_STAGE_LR)
        for (int p = 0; p < g.n_planes; p++) if (h.lr_type[p] != RB200_RESTORATION_NONE) restore_planes |= 1 << p;

    // Loop-filter masks / levels from block records (lfmask.cu): nothing in them depends on the picture, so the records'
    // upload and the two launches run on a side stream beside the reconstruction.  Fork here: the previous frame's
    // filters (on st) still read the arrays that are rebuilt.
    const bool build_lf = f->n_lfb && (do_lf || do_cdef);
    if (build_lf) {
        RB_CUDA(cudaEventRecord(f->lf_fork, st));
        RB_CUDA(cudaStreamWaitEvent(f->lf_stream, f->lf_fork, 0));
    }
    // ---- host -> device: the per-frame batch
    if (upload) {
        if (stages & RB200_STAGE_RECON) {
            if (n_coefs && upload == RB200_UPLOAD_ALL)
                RB_CUDA(cudaMemcpyAsync(f->d_coef, f->h_coef, n_coefs * f->cs, cudaMemcpyHostToDevice, st));
            if (upload == RB200_UPLOAD_GATHER_COEF16 && (f->cs != 4 || !f->h_coef16))
                return set_error(-22, "frame_submit: RB200_UPLOAD_GATHER_COEF16 needs a 16-bit picture whose rb200_frame_coef16_buffer() was filled");
            if (upload == RB200_UPLOAD_PACKED_COEF16 && (f->cs != 4 || !f->h_pk))
                return set_error(-22, "frame_submit: RB200_UPLOAD_PACKED_COEF16 needs a 16-bit picture whose rb200_frame_coef_stream() was filled");
            const bool gather = (upload == RB200_UPLOAD_GATHER_COEF || upload == RB200_UPLOAD_GATHER_COEF16 || upload == RB200_UPLOAD_PACKED_COEF16) && n_itx;
            if (gather) {
                // The coefficient gather runs on a high-priority side stream: its CTAs are scheduled ahead of the
                // stage kernels of other frames, so the PCIe link stays busy, and it overlaps this frame's
                // prediction.  Fork after the previous frame's transforms (they read what is overwritten here).
                if (!f->up_stream) {
                    int lo = 0, hi = 0;
                    RB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
                    RB_CUDA(cudaStreamCreateWithPriority(&f->up_stream, cudaStreamNonBlocking, hi));
                    RB_CUDA(cudaEventCreateWithFlags(&f->up_fork, cudaEventDisableTiming));
                    RB_CUDA(cudaEventCreateWithFlags(&f->up_join, cudaEventDisableTiming));
                }
                RB_CUDA(cudaEventRecord(f->up_fork, st));
                RB_CUDA(cudaStreamWaitEvent(f->up_stream, f->up_fork, 0));
                RB_CUDA(cudaMemcpyAsync(f->d_itx, f->h_itx, (size_t)n_itx * sizeof(Rb200ItxItem), cudaMemcpyHostToDevice, f->up_stream));
                int rg;
                if (upload == RB200_UPLOAD_PACKED_COEF16) {
                    // plain DMA of the stream and its offsets, then the blocks are spread out on the device
                    if (f->n_pk) RB_CUDA(cudaMemcpyAsync(f->d_pk, f->h_pk, f->n_pk * sizeof(int16_t), cudaMemcpyHostToDevice, f->up_stream));
                    RB_CUDA(cudaMemcpyAsync(f->d_pkoff, f->h_pkoff, (size_t)n_itx * sizeof(uint32_t), cudaMemcpyHostToDevice, f->up_stream));
                    if (f->n_esc)
                        RB_CUDA(cudaMemcpyAsync(f->d_esc, f->h_esc, (size_t)f->n_esc * sizeof(Rb200CoefEscape), cudaMemcpyHostToDevice, f->up_stream));
                    if ((rg = coef_expand16_launch(f->d_pk, f->d_pkoff, (int32_t *)f->d_coef, f->d_itx, n_itx, f->d_esc, f->n_esc, f->up_stream, &f->launches))) return rg;
                } else if (upload == RB200_UPLOAD_GATHER_COEF16) {
                    if (f->n_esc)
                        RB_CUDA(cudaMemcpyAsync(f->d_esc, f->h_esc, (size_t)f->n_esc * sizeof(Rb200CoefEscape), cudaMemcpyHostToDevice, f->up_stream));
                    if ((rg = coef_gather16_launch(f->h_coef16, (int32_t *)f->d_coef, f->d_itx, n_itx, f->d_esc, f->n_esc, f->up_stream, &f->launches))) return rg;
                } else {
                    if ((rg = coef_gather_launch(f->h_coef, f->d_coef, f->d_itx, n_itx, f->bdmax, f->up_stream))) return rg;
                    f->launches++;
                }
                RB_CUDA(cudaEventRecord(f->up_join, f->up_stream));
            } else if (n_itx) {
                RB_CUDA(cudaMemcpyAsync(f->d_itx, f->h_itx, (size_t)n_itx * sizeof(Rb200ItxItem), cudaMemcpyHostToDevice, st));
            }
            if (n_mc) RB_CUDA(cudaMemcpyAsync(f->d_mc, f->h_mc, (size_t)n_mc * sizeof(Rb200McItem), cudaMemcpyHostToDevice, st));
            if (f->n_comp) RB_CUDA(cudaMemcpyAsync(f->d_comp, f->h_comp, (size_t)f->n_comp * sizeof(Rb200CompItem), cudaMemcpyHostToDevice, st));
            if (f->n_warp) RB_CUDA(cudaMemcpyAsync(f->d_warp, f->h_warp, (size_t)f->n_warp * sizeof(Rb200WarpItem), cudaMemcpyHostToDevice, st));
            if ((stages & RB200_STAGE_INTRA) && f->n_levels) {
                int n_in = 0;
                for (int l = 0; l < f->n_levels; l++) n_in += f->intra_counts[l];
                if (n_in) {
                    RB_CUDA(cudaMemcpyAsync(f->d_intra, f->h_intra, (size_t)n_in * sizeof(Rb200IntraItem), cudaMemcpyHostToDevice, st));
                    RB_CUDA(cudaMemcpyAsync(f->d_intra_itx, f->h_intra_itx, (size_t)n_in * sizeof(int32_t), cudaMemcpyHostToDevice, st));
                    RB_CUDA(cudaMemcpyAsync(f->d_level_off, f->h_level_off, (size_t)(f->n_levels_nonempty + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
                    if (f->n_pal) RB_CUDA(cudaMemcpyAsync(f->d_pal, f->h_pal, f->n_pal, cudaMemcpyHostToDevice, st));
                }
            }
            if (f->n_scaled + f->n_scaled_above + f->n_scaled_left)
                RB_CUDA(cudaMemcpyAsync(f->d_scaled, f->h_scaled, (size_t)(f->n_scaled + f->n_scaled_above + f->n_scaled_left) * sizeof(Rb200McScaledItem), cudaMemcpyHostToDevice, st));
            if (f->n_obmc_above + f->n_obmc_left)
                RB_CUDA(cudaMemcpyAsync(f->d_obmc, f->h_obmc, (size_t)(f->n_obmc_above + f->n_obmc_left) * sizeof(Rb200McItem), cudaMemcpyHostToDevice, st));
        }
        if (build_lf) {
            // masks and levels are built on the device from the block records; only cdef_idx travels
            const size_t n_sb = (size_t)g.sb128w * g.sb128h;
            for (size_t i = 0; i < n_sb; i++) memcpy(f->h_cdef_idx + 4 * i, f->h_masks[i].cdef_idx, 4);
            RB_CUDA(cudaMemcpyAsync(f->d_cdef_idx, f->h_cdef_idx, n_sb * 4, cudaMemcpyHostToDevice, f->lf_stream));
            RB_CUDA(cudaMemcpyAsync(f->d_lfb, f->h_lfb, (size_t)f->n_lfb * sizeof(Rb200LfBlock), cudaMemcpyHostToDevice, f->lf_stream));
        } else if (do_lf || do_cdef)
            RB_CUDA(cudaMemcpyAsync(f->d_masks, f->h_masks, f->n_masks * sizeof(Rb200Av1Filter), cudaMemcpyHostToDevice, st));
        if (do_lf) {
            if (!f->n_lfb) RB_CUDA(cudaMemcpyAsync(f->d_lvl, f->h_lvl, f->n_lvl * 4, cudaMemcpyHostToDevice, st));
            RB_CUDA(cudaMemcpyAsync(f->d_lut, f->h_lut, sizeof(Rb200Av1FilterLUT), cudaMemcpyHostToDevice, st));
        }
        if (restore_planes)
            RB_CUDA(cudaMemcpyAsync(f->d_lr, f->h_lr, f->n_masks * sizeof(Rb200Av1Restoration), cudaMemcpyHostToDevice, st));
    }

    // producers named with rb200_frame_depend: the batch uploads above do not touch pictures and stay ahead of the wait
    for (int i = 0; i < f->n_deps; i++) RB_CUDA(cudaStreamWaitEvent(st, f->dep_events[i], 0));
    f->n_deps = 0;
    MARK(1);
    int r;
    if (build_lf) {
        if ((r = lf_build_launch(f->d_lfb, f->n_lfb, g.w4, g.h4, g.sb128w, g.sb128h, g.b4_stride, g.ss_hor, g.ss_ver, g.n_planes,
                                 f->d_lf_cells, f->d_cdef_idx, f->d_masks, f->d_lvl + 32, f->lf_stream))) return r;
        f->launches += 2;
        RB_CUDA(cudaEventRecord(f->lf_join, f->lf_stream));
    }
    // ---- reconstruction: prediction, then residual add.
    // With the lists sorted luma first (rb200_frame_set_plane_counts) and nothing but put predictions and residuals in the
    // frame, the luma and the chroma reconstruction are two chains on two streams that run on into the post-filters.
    bool recon_split = f->plane_split && f->plane_counts && g.n_planes > 1 && (stages & RB200_STAGE_RECON) && !f->n_comp && !f->n_warp &&
                       !f->n_scaled && !f->n_scaled_above && !f->n_scaled_left && !f->n_obmc_above && !f->n_obmc_left && !((stages & RB200_STAGE_INTRA) && f->n_levels) && !do_sr &&
                       !(f->band_s1 > f->band_s0) && f->n_mc_luma <= n_mc && (do_lf || do_cdef || restore_planes);
    if (recon_split)
        for (int t = 0; t < RB200_N_RECT_TX_SIZES; t++) if (f->itx_luma[t] > itx_counts[t]) recon_split = false;
    if (recon_split) {
        if (!f->uv_stream) RB_CUDA(cudaStreamCreateWithFlags(&f->uv_stream, cudaStreamNonBlocking));
        if (!f->uv_fork) {
            RB_CUDA(cudaEventCreateWithFlags(&f->uv_fork, cudaEventDisableTiming));
            RB_CUDA(cudaEventCreateWithFlags(&f->uv_dir, cudaEventDisableTiming));
            RB_CUDA(cudaEventCreateWithFlags(&f->uv_join, cudaEventDisableTiming));
        }
        cudaStream_t su = f->uv_stream;
        RB_CUDA(cudaEventRecord(f->uv_fork, st));          // behind the batch uploads and the producers' events
        RB_CUDA(cudaStreamWaitEvent(su, f->uv_fork, 0));
        const int n_luma = f->n_mc_luma, n_chroma = n_mc - f->n_mc_luma;
        if (n_luma) {
            if ((r = mc_batch_launch(f->planes[0], f->refs, f->n_refs, h.width, h.height, g.ss_hor, g.ss_ver, f->d_mc, n_luma, f->bdmax, st,
                                     f->d_counters, &f->tm_refs))) return r;
            f->launches++;
        }
        if (n_chroma) {
            if ((r = mc_batch_launch(f->planes[0], f->refs, f->n_refs, h.width, h.height, g.ss_hor, g.ss_ver, f->d_mc + n_luma, n_chroma, f->bdmax,
                                     su, f->d_counters + 1, &f->tm_refs))) return r;
            f->launches++;
        }
        if ((upload == RB200_UPLOAD_GATHER_COEF || upload == RB200_UPLOAD_GATHER_COEF16 || upload == RB200_UPLOAD_PACKED_COEF16) && n_itx) {
            RB_CUDA(cudaStreamWaitEvent(st, f->up_join, 0));
            RB_CUDA(cudaStreamWaitEvent(su, f->up_join, 0));
        }
        const void *cf = upload == RB200_UPLOAD_ZERO_COPY_COEF ? f->h_coef : f->d_coef;
        int off = 0;
        for (int t = 0; t < RB200_N_RECT_TX_SIZES; t++) {
            const int nl = f->itx_luma[t], nc = itx_counts[t] - nl;
            if (nl) { if ((r = itx_launch(t, f->planes[0], cf, f->d_itx + off, nl, f->bdmax, st))) return r; f->launches++; }
            if (nc) { if ((r = itx_launch(t, f->planes[0], cf, f->d_itx + off + nl, nc, f->bdmax, su))) return r; f->launches++; }
            off += itx_counts[t];
        }
        if (build_lf) RB_CUDA(cudaStreamWaitEvent(su, f->lf_join, 0));
    } else if (stages & RB200_STAGE_RECON) {
        if (n_mc) {
            if ((r = mc_batch_launch(f->planes[0], f->refs, f->n_refs, h.width, h.height, g.ss_hor, g.ss_ver, f->d_mc,
                                     n_mc, f->bdmax, st, f->d_counters, &f->tm_refs))) return r;
            f->launches++;
        }
        if (f->n_comp) {
            if ((r = mc_comp_batch_launch_gmv(f->planes[0], f->refs, f->n_refs, h.width, h.height, h.layout, f->d_comp, f->n_comp,
                                              f->bdmax, st, f->ref_gmv, &f->ref_dims))) return r;
            f->launches++;
        }
        if (f->n_scaled) {
            if ((r = mc_scaled_batch_launch(f->planes[0], f->refs, f->n_refs, f->ref_dims, g.ss_hor, g.ss_ver, f->d_scaled,
                                            f->n_scaled, f->bdmax, st))) return r;
            f->launches++;
        }
        if (f->n_warp) {
            if ((r = mc_warp_batch_launch(f->planes[0], f->refs, f->n_refs, h.width, h.height, h.layout, f->d_warp, f->n_warp,
                                          f->bdmax, st))) return r;
            f->launches++;
        }
        // OBMC: every ABOVE strip, then every LEFT strip, on top of the finished predictions
        if (f->n_obmc_above) {
            if ((r = mc_obmc_batch_launch(f->planes[0], f->refs, f->n_refs, h.width, h.height, g.ss_hor, g.ss_ver, f->d_obmc,
                                          f->n_obmc_above, f->bdmax, st))) return r;
            f->launches++;
        }
        if (f->n_scaled_above) {      // ... strips from references of another size with the strips of their kind
            if ((r = mc_scaled_batch_launch(f->planes[0], f->refs, f->n_refs, f->ref_dims, g.ss_hor, g.ss_ver, f->d_scaled + f->n_scaled,
                                            f->n_scaled_above, f->bdmax, st))) return r;
            f->launches++;
        }
        if (f->n_obmc_left) {
            if ((r = mc_obmc_batch_launch(f->planes[0], f->refs, f->n_refs, h.width, h.height, g.ss_hor, g.ss_ver,
                                          f->d_obmc + f->n_obmc_above, f->n_obmc_left, f->bdmax, st))) return r;
            f->launches++;
        }
        if (f->n_scaled_left) {
            if ((r = mc_scaled_batch_launch(f->planes[0], f->refs, f->n_refs, f->ref_dims, g.ss_hor, g.ss_ver,
                                            f->d_scaled + f->n_scaled + f->n_scaled_above, f->n_scaled_left, f->bdmax, st))) return r;
            f->launches++;
        }
        MARK(2);
        if ((upload == RB200_UPLOAD_GATHER_COEF || upload == RB200_UPLOAD_GATHER_COEF16 || upload == RB200_UPLOAD_PACKED_COEF16) && n_itx) RB_CUDA(cudaStreamWaitEvent(st, f->up_join, 0));
        // The residual launches (one per transform size present) write disjoint pixels: they are dealt to two streams,
        // the heavier buckets first, so that they overlap and fill each other's tails.
        {
            const void *cf = upload == RB200_UPLOAD_ZERO_COPY_COEF ? f->h_coef : f->d_coef;
            static const int tx_area_log2[RB200_N_RECT_TX_SIZES] = {4, 6, 8, 10, 12, 5, 5, 7, 7, 9, 9, 11, 11, 6, 6, 8, 8, 10, 10};
            int order[RB200_N_RECT_TX_SIZES], offs[RB200_N_RECT_TX_SIZES], n_b = 0, off = 0;
            for (int t = 0; t < RB200_N_RECT_TX_SIZES; t++) { offs[t] = off; off += itx_counts[t]; if (itx_counts[t]) order[n_b++] = t; }
            auto weight = [&](int t) { return (int64_t)itx_counts[t] << tx_area_log2[t]; };
            for (int i = 1; i < n_b; i++)
                for (int j = i; j > 0 && weight(order[j]) > weight(order[j - 1]); j--) { const int x = order[j]; order[j] = order[j - 1]; order[j - 1] = x; }
            const bool two = f->plane_split && n_b > 1;
            if (two) {
                if (!f->uv_stream) RB_CUDA(cudaStreamCreateWithFlags(&f->uv_stream, cudaStreamNonBlocking));
                if (!f->itx_fork) {
                    RB_CUDA(cudaEventCreateWithFlags(&f->itx_fork, cudaEventDisableTiming));
                    RB_CUDA(cudaEventCreateWithFlags(&f->itx_join, cudaEventDisableTiming));
                }
                RB_CUDA(cudaEventRecord(f->itx_fork, st));
                RB_CUDA(cudaStreamWaitEvent(f->uv_stream, f->itx_fork, 0));
            }
            int64_t load[2] = {0, 0};
            for (int i = 0; i < n_b; i++) {
                const int t = order[i], which = two && load[1] < load[0] ? 1 : 0;
                load[which] += weight(t);
                // zero-copy: the kernel pulls exactly the coefficient columns it needs (Rb200ItxItem.ncols)
                // over PCIe from the pinned staging instead of a full H2D copy first
                if ((r = itx_launch(t, f->planes[0], cf, f->d_itx + offs[t], itx_counts[t], f->bdmax, which ? f->uv_stream : st))) return r;
                f->launches++;
            }
            if (two) {
                RB_CUDA(cudaEventRecord(f->itx_join, f->uv_stream));
                RB_CUDA(cudaStreamWaitEvent(st, f->itx_join, 0));
            }
        }
        // ---- intra blocks: one launch per dependency level; a CTA prepares a block's edge from the reconstructed
        // picture, predicts it and adds its residual
        if ((stages & RB200_STAGE_INTRA) && f->n_levels) {
            const void *cf = upload == RB200_UPLOAD_ZERO_COPY_COEF ? f->h_coef : f->d_coef;
            static const bool per_level = getenv("RB200_INTRA_LEVEL_LAUNCHES") && atoi(getenv("RB200_INTRA_LEVEL_LAUNCHES"));
            if (!per_level) {
                // the whole wavefront in one cooperative launch; levels are separated by an arrival counter in global memory
                if ((r = intra_levels_launch(f->planes[0], f->d_intra, f->d_intra_itx, f->d_itx, cf, f->d_pal, f->d_level_off, f->n_levels_nonempty,
                                             f->intra_widest, g.bw, g.bh, g.ss_hor, g.ss_ver, f->bdmax, f->d_intra_sync, st))) return r;
                f->launches++;
                f->intra_check = true;
            } else {
                int ioff = 0;
                for (int l = 0; l < f->n_levels; l++) {
                    if (!f->intra_counts[l]) continue;
                    if ((r = intra_items_launch(f->planes[0], f->d_intra + ioff, f->d_intra_itx + ioff, f->d_itx, cf, f->d_pal, f->intra_counts[l],
                                                g.bw, g.bh, g.ss_hor, g.ss_ver, f->bdmax, st))) return r;
                    f->launches++;
                    ioff += f->intra_counts[l];
                }
            }
        }
    }
    f->out = f->planes[0];
    if (stages & RB200_STAGE_RECON) MARK(3); else { MARK(2); MARK(3); }
    if (build_lf) RB_CUDA(cudaStreamWaitEvent(st, f->lf_join, 0));
    // The post-filters never mix planes (only the chroma CDEF wants the luma direction search), so the luma chain and
    // the chroma chain of a frame run on two streams: each fills the issue slots the other leaves idle.
    const bool split = recon_split || (f->plane_split && g.n_planes > 1 && !do_sr && !(f->band_s1 > f->band_s0) && (do_lf || do_cdef || restore_planes));
    cudaStream_t su = st;
    if (split) {
        if (!f->uv_stream) RB_CUDA(cudaStreamCreateWithFlags(&f->uv_stream, cudaStreamNonBlocking));
        if (!f->uv_fork) {
            RB_CUDA(cudaEventCreateWithFlags(&f->uv_fork, cudaEventDisableTiming));
            RB_CUDA(cudaEventCreateWithFlags(&f->uv_dir, cudaEventDisableTiming));
            RB_CUDA(cudaEventCreateWithFlags(&f->uv_join, cudaEventDisableTiming));
        }
        su = f->uv_stream;
        if (!recon_split) {      // (a split reconstruction has forked the chroma chain already)
            RB_CUDA(cudaEventRecord(f->uv_fork, st));
            RB_CUDA(cudaStreamWaitEvent(su, f->uv_fork, 0));
        }
    }
    // ---- deblock (in place): all column edges, then all row edges (src/recon.rs:4047-4170)
    if (do_lf) {
        const int uv_mask = (h.lf_level_u || h.lf_level_v) && g.n_planes > 1 ? 6 : 0;
        if ((r = deblock_planes_launch(f->planes[0], g.n_planes, g.w4, g.h4, g.sb128w, g.b4_stride, g.ss_hor, g.ss_ver, split ? 1 : 1 | uv_mask,
                                       f->d_masks, f->d_lvl + 32, f->d_lut, f->bdmax, st, &f->launches, band.y4b, band.y4e))) return r;
        if (split && uv_mask &&
            (r = deblock_planes_launch(f->planes[0], g.n_planes, g.w4, g.h4, g.sb128w, g.b4_stride, g.ss_hor, g.ss_ver, uv_mask, f->d_masks,
                                       f->d_lvl + 32, f->d_lut, f->bdmax, su, &f->launches, band.y4b, band.y4e))) return r;
    }
    MARK(4);
    // ---- CDEF: cur -> p2 (src/recon.rs:4172-4213)
    if (do_cdef) {
        CdefFrameParams P;
        P.bw = g.bw; P.bh = g.bh; P.sb128w = g.sb128w; P.ss_hor = g.ss_hor; P.ss_ver = g.ss_ver;
        P.n_planes = g.n_planes; P.bdmin8 = h.bpc - 8; P.damping = h.cdef_damping + P.bdmin8;
        for (int i = 0; i < 8; i++) { P.y_strength[i] = h.cdef_y_strength[i]; P.uv_strength[i] = h.cdef_uv_strength[i]; }
        P.layout_422 = h.layout == RB200_LAYOUT_I422;
        const CUtensorMap *maps = f->tm_cdef_ok ? f->tm_cdef : nullptr;
        if (!split) {
            if ((r = cdef_planes_launch(f->planes[0], f->planes[1], P, f->d_masks, f->d_cdef_blk, f->bdmax, st, band.t0, band.t1, maps, 7, 3,
                                        &f->launches))) return r;
        } else {
            // per-block decisions (direction search on the deblocked luma), then the two filter launches
            if ((r = cdef_planes_launch(f->planes[0], f->planes[1], P, f->d_masks, f->d_cdef_blk, f->bdmax, st, band.t0, band.t1, maps, 0, 1,
                                        &f->launches))) return r;
            RB_CUDA(cudaEventRecord(f->uv_dir, st));
            RB_CUDA(cudaStreamWaitEvent(su, f->uv_dir, 0));
            if ((r = cdef_planes_launch(f->planes[0], f->planes[1], P, f->d_masks, f->d_cdef_blk, f->bdmax, st, band.t0, band.t1, maps, 1, 2,
                                        &f->launches))) return r;
            if ((r = cdef_planes_launch(f->planes[0], f->planes[1], P, f->d_masks, f->d_cdef_blk, f->bdmax, su, band.t0, band.t1, maps, 6, 2,
                                        &f->launches))) return r;
        }
        f->out = f->planes[1];
    }
    MARK(5);
    // ---- super-resolution: the CDEF output and -- for the rows loop restoration takes from outside its
    // stripes -- the deblocked picture are upscaled horizontally (rav1d_filter_sbrow_resize src/recon.rs:4215-4281,
    // backup_lpf with resize src/lf_apply.rs:24-141)
    f->out_w = h.width;
    if (do_sr && f->band_s1 > f->band_s0)
        return set_error(-38, "frame_submit: super-resolution together with a band restriction is not implemented");
    if (do_sr) {
        const Rb200Planes cdefp = f->out;
        for (int p = 0; p < g.n_planes; p++) {
            const int ssh = p ? g.ss_hor : 0, ssv = p ? g.ss_ver : 0;
            const int dst_w = (f->sr_w + ssh) >> ssh, src_w = (4 * g.bw + ssh) >> ssh, rows = (h.height + ssv) >> ssv;
            if ((r = resize_plane_launch(f->sr_planes[0].data[p], f->sr_planes[0].stride[p], cdefp.data[p], cdefp.stride[p], dst_w,
                                         rows, src_w, f->resize_step[p ? 1 : 0], f->resize_start[p ? 1 : 0], f->bdmax, st))) return r;
            f->launches++;
            if ((restore_planes & (1 << p)) && cdefp.data[p] != f->planes[0].data[p]) {
                if ((r = resize_plane_launch(f->sr_planes[1].data[p], f->sr_planes[1].stride[p], f->planes[0].data[p],
                                             f->planes[0].stride[p], dst_w, rows, src_w, f->resize_step[p ? 1 : 0],
                                             f->resize_start[p ? 1 : 0], f->bdmax, st))) return r;
                f->launches++;
            }
        }
        f->out = f->sr_planes[0];
        f->out_w = f->sr_w;
    }
    // ---- loop restoration: (p2 | cur) + cur -> p3 for the restored planes (src/recon.rs:4283-4317)
    if (restore_planes) {
        const Rb200Planes cdefp = f->out;
        for (int p = 0; p < g.n_planes; p++) {
            if (!(restore_planes & (1 << p))) continue;
            LrFrameParams P;
            P.plane = p;
            P.ss_hor = p ? g.ss_hor : 0; P.ss_ver = p ? g.ss_ver : 0;
            P.w = ((do_sr ? f->sr_w : h.width) + P.ss_hor) >> P.ss_hor; P.h = (h.height + P.ss_ver) >> P.ss_ver;
            P.unit_log2 = h.lr_unit_size_log2[p ? 1 : 0];
            P.sb128 = h.sb128; P.sbh = g.sbh; P.sr_sb128w = do_sr ? (f->sr_w + 127) >> 7 : g.sb128w;
            P.stripe_first = band.s0; P.stripe_end = band.s1;
            // rows outside a stripe come from the deblocked picture (upscaled alike under super-resolution;
            // without CDEF the two are the same plane)
            const Rb200Planes &dbl = !do_sr ? f->planes[0] : (cdefp.data[p] == f->sr_planes[0].data[p] && !do_cdef ? f->sr_planes[0] : f->sr_planes[1]);
            const Rb200Planes &dst = do_sr ? f->sr_planes[2] : f->planes[2];
            const CUtensorMap *m_main = nullptr, *m_halo = nullptr;     // windows by TMA when the planes are the context's own
            if (f->tm_lr_ok && !do_sr && dbl.data[p] == f->planes[0].data[p]) {
                if (cdefp.data[p] == f->planes[1].data[p]) m_main = &f->tm_lr_main[1][p];
                else if (cdefp.data[p] == f->planes[0].data[p]) m_main = &f->tm_lr_main[0][p];
                if (m_main) m_halo = &f->tm_lr_halo[p];
            }
            if ((r = lr_plane_launch_tma((const uint8_t *)cdefp.data[p], (const uint8_t *)dbl.data[p], (uint8_t *)dst.data[p], dst.stride[p], P,
                                         f->d_lr, f->bdmax, p ? su : st, m_main, m_halo)))
                return r;
            f->launches++;
            f->out.data[p] = dst.data[p];
            f->out.stride[p] = dst.stride[p];
        }
    }
    if (split) {
        RB_CUDA(cudaEventRecord(f->uv_join, su));
        RB_CUDA(cudaStreamWaitEvent(st, f->uv_join, 0));
    }
    MARK(6);
    f->display = f->out;
    if (stages & RB200_STAGE_FILM_GRAIN) {
        if (!f->fg_set) return set_error(-22, "frame_submit: RB200_STAGE_FILM_GRAIN without rb200_frame_set_film_grain");
        RB_CUDA(cudaStreamWaitEvent(st, f->fg_join, 0));
        if ((r = film_grain_apply(f, st))) return r;
    }
    MARK(7);
#undef MARK
    RB_CUDA(cudaEventRecord(f->done_event, st));   // a read-back queued after this does not hold consumers up
    return 0;
}
static_assert(sizeof(Rb200WarpItem) == 48 && sizeof(Rb200CompItem) == 32 && sizeof(Rb200McItem) == 16 && sizeof(Rb200ItxItem) == 16 && sizeof(Rb200McScaledItem) == 32 && sizeof(Rb200IntraItem) == 16 && sizeof(Rb200LfBlock) == 16, "batch record sizes");
