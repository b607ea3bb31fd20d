// Stage launchers that take a plane subset (bit p = plane p): the frame path runs the luma and the chroma post-filter
// chains of a frame on two streams, so that one chain fills the issue slots the other leaves idle.
#pragma once
#include "common.cuh"

namespace rb200 {

int deblock_planes_launch(const Rb200Planes &pl, int n_planes, int w4, int h4, int sb128w, int b4_stride, int ss_hor,
                          int ss_ver, int plane_mask, const Rb200Av1Filter *masks, const uint8_t (*lvl)[4],
                          const Rb200Av1FilterLUT *lut, int bdmax, cudaStream_t st, int *launches, int y4b, int y4e);
// what: bit 0 = the per-block decisions (direction search; needs the luma plane), bit 1 = filter the planes of plane_mask
int cdef_planes_launch(const Rb200Planes &src, const Rb200Planes &dst, const CdefFrameParams &P, const Rb200Av1Filter *masks,
                       void *blk_scratch, int bdmax, cudaStream_t st, int t0, int t1, const CUtensorMap_st *maps, int plane_mask,
                       int what, int *launches);

// compound blocks, with the references' global-motion parameters for predictions that are warps (Rb200CompItem.warp_mask)
struct McGmv { int32_t matrix[6]; int16_t abcd[4]; };
struct McGmvSet { McGmv g[8]; };
// dims: luma size of every reference slot; a prediction from a slot whose size differs from ref_w x ref_h is a scaled one
int mc_comp_batch_launch_gmv(const Rb200Planes &dst, const Rb200Planes *refs, int n_refs, int ref_w, int ref_h, int layout,
                             const Rb200CompItem *d_items, int n, int bdmax, cudaStream_t st, const McGmvSet &gmv,
                             const McRefDims *dims = nullptr);
// loop restoration with the window fetched by TMA (lr.cu)
int lr_encode_maps(CUtensorMap_st *map_main, CUtensorMap_st *map_halo, const void *cdef, const void *dbl, int64_t stride, int w, int h);
int lr_plane_launch_tma(const uint8_t *cdef, const uint8_t *dbl, uint8_t *out, int64_t stride, const LrFrameParams &P,
                        const Rb200Av1Restoration *lrm, int bdmax, cudaStream_t st, const CUtensorMap_st *map_main,
                        const CUtensorMap_st *map_halo);
// int16 coefficient transport (RB200_UPLOAD_GATHER_COEF16): each block's leading columns pulled from the pinned int16
// staging and widened into the int32 device array, then the escapes patched in
int coef_gather16_launch(const int16_t *h_cf16, int32_t *d_cf, const Rb200ItxItem *d_items, int n, const Rb200CoefEscape *d_esc,
                         int n_esc, cudaStream_t st, int *launches);

// packed int16 stream (RB200_UPLOAD_PACKED_COEF16): block i's leading columns at d_stream + d_off[i] -> int32 at its cf_off
int coef_expand16_launch(const int16_t *d_stream, const uint32_t *d_off, int32_t *d_cf, const Rb200ItxItem *d_items, int n,
                         const Rb200CoefEscape *d_esc, int n_esc, cudaStream_t st, int *launches);

}  // namespace rb200
