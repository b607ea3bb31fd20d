// Wedge and inter-intra blend masks in closed form (shared by the compound kernel, mc.cu, and the inter-intra path
// of the intra kernel, ipred.cu).
#pragma once
#include "common.cuh"

namespace rb200 {

// ---- wedge masks (dav1d_wedge_masks, src/wedge.rs:377 == src/wedge.c:83-243), evaluated in closed form.
// The reference fills 64x64 master templates -- a vertical edge, an oblique one that moves one column
// every two rows (with different 8-entry ramps on even and odd rows), their transposes and mirror
// images -- and copies a w x h window per codebook entry.  A mask value is therefore a ramp lookup at a
// position that is linear in (x, y); flipped when the sign table says the wedge is stored inverted.
static __device__ const uint8_t k_wedge_ramp[3][8] = {
    { 1, 2, 6, 18, 37, 53, 60, 63 },     // odd rows of the oblique master
    { 1, 4, 11, 27, 46, 58, 62, 63 },    // even rows
    { 0, 2, 7, 21, 43, 57, 62, 64 },     // vertical / horizontal master
};
// codebooks {direction, x_offset, y_offset} packed dir | xo << 4 | yo << 8; directions: 0 horizontal,
// 1 vertical, 2 oblique 27, 3 oblique 63, 4 oblique 117, 5 oblique 153 (AV1 spec 7.11.3.11 Wedge_Codebook)
#define WC(d, x, y) ((d) | (x) << 4 | (y) << 8)
static __device__ const uint16_t k_wedge_codebook[3][16] = {
    // h > w
    { WC(2,4,4), WC(3,4,4), WC(4,4,4), WC(5,4,4), WC(0,4,2), WC(0,4,4), WC(0,4,6), WC(1,4,4),
      WC(2,4,2), WC(2,4,6), WC(5,4,2), WC(5,4,6), WC(3,2,4), WC(3,6,4), WC(4,2,4), WC(4,6,4) },
    // h < w
    { WC(2,4,4), WC(3,4,4), WC(4,4,4), WC(5,4,4), WC(1,2,4), WC(1,4,4), WC(1,6,4), WC(0,4,4),
      WC(2,4,2), WC(2,4,6), WC(5,4,2), WC(5,4,6), WC(3,2,4), WC(3,6,4), WC(4,2,4), WC(4,6,4) },
    // h == w
    { WC(2,4,4), WC(3,4,4), WC(4,4,4), WC(5,4,4), WC(0,4,2), WC(0,4,6), WC(1,2,4), WC(1,6,4),
      WC(2,4,2), WC(2,4,6), WC(5,4,2), WC(5,4,6), WC(3,2,4), WC(3,6,4), WC(4,2,4), WC(4,6,4) },
};
#undef WC
__device__ __forceinline__ int wedge_ramp(int line, int pos, int ctr) {
    const int d = pos - ctr + 4;
    return d < 0 ? 0 : (d >= 8 ? 64 : k_wedge_ramp[line][d]);
}
// oblique-63 master at column x, row y: the edge centre starts at column 48 and moves left one column per row pair
__device__ __forceinline__ int wedge_o63(int x, int y) {
    return (y & 1) ? wedge_ramp(0, x, 47 - (y >> 1)) : wedge_ramp(1, x, 48 - (y >> 1));
}
// which wedges are stored inverted, per block size (src/wedge.c:232-240)
__device__ __forceinline__ unsigned wedge_signs(int w, int h) {
    if (w == h) return 0x7bfb;
    if (w == 32 && h == 8) return 0x6beb;
    if (w == 8 && h == 32) return 0x7aeb;
    return 0x7beb;
}
__device__ __forceinline__ int wedge_mask_at(int w, int h, int idx, int x, int y) {
    const unsigned cb = k_wedge_codebook[h > w ? 0 : (h < w ? 1 : 2)][idx];
    const int dir = cb & 15, xo = (cb >> 4) & 15, yo = cb >> 8;
    const int mx = x + 32 - ((w * xo) >> 3), my = y + 32 - ((h * yo) >> 3);
    int v;
    switch (dir) {
    case 0: v = wedge_ramp(2, my, 32); break;
    case 1: v = wedge_ramp(2, mx, 32); break;
    case 2: v = wedge_o63(my, mx); break;            // transpose of oblique 63
    case 3: v = wedge_o63(mx, my); break;
    case 4: v = wedge_o63(63 - mx, my); break;       // mirror image of oblique 63
    default: v = wedge_o63(my, 63 - mx); break;      // mirror image of oblique 27
    }
    return ((wedge_signs(w, h) >> idx) & 1) ? 64 - v : v;
}

// Inter-intra blend masks (dav1d_ii_masks, src/wedge.rs == src/wedge.c:262-340): a 1-D weight ramp sampled with a step
// that depends on the larger block side; DC blends evenly.  mode: 0 DC, 1 vertical, 2 horizontal, 3 smooth.
static __device__ const uint8_t k_ii_weights_1d[32] = { 60, 52, 45, 39, 34, 30, 26, 22, 19, 17, 15, 13, 11, 10, 8, 7,
                                                        6, 6, 5, 4, 4, 3, 3, 2, 2, 2, 2, 1, 1, 1, 1, 1 };
__device__ __forceinline__ int ii_mask_at(int w, int h, int mode, int x, int y) {
    if (mode == 0) return 32;
    const int step = 32 / imax(w, h);
    return k_ii_weights_1d[(mode == 1 ? y : (mode == 2 ? x : imin(x, y))) * step];
}

}  // namespace rb200
