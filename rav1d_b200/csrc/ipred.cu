// Intra prediction for sm_100a: the per-call slots of Rav1dIntraPredDSPContext
// { intra_pred[14], cfl_ac[3], cfl_pred[6], pal_pred } (src/ipred.rs:164-169; bodies src/ipred.rs:171-1500 ==
// src/ipred_tmpl.c:39-735).  First step of SURVEY row f1: function-level parity of every predictor; the
// frame-level wavefront over transform blocks is the next step.
//
// One CTA predicts one block.  The caller's edge (`topleft[-(h + min(w, h)) .. w + min(w, h)]`, everything the
// reference may read) is mirrored into shared memory; the directional predictors prepare their filtered or
// upsampled edge there in parallel (each output sample of filter_edge / upsample_edge is independent), then
// every pixel is one thread's closed-form lookup -- the reference's running xpos / ypos accumulators become
// dx * (y + 1) etc.  Filter-intra, the only predictor with a dependency inside the block, runs as a wavefront
// over its 4x2 sub-blocks (a sub-block needs the ones to its left, above and above-left).
#include "common.cuh"
#include <cstdio>
#include <vector>
#include "tables.cuh"
#include "itx_block.cuh"
#include "wedge.cuh"

namespace rb200 {

// enum IntraPredMode with the implementation modes (src/levels.rs:85-130)
enum { IP_DC = 0, IP_VERT, IP_HOR, IP_LEFT_DC, IP_TOP_DC, IP_DC_128, IP_Z1, IP_Z2, IP_Z3, IP_SMOOTH, IP_SMOOTH_V, IP_SMOOTH_H,
       IP_PAETH, IP_FILTER, IP_N_MODES };

constexpr int IP_EC = 128;        // index of topleft inside the mirrored edge (the reference's edge[257] / topleft = edge + 128)
constexpr int IP_EDGE = 2 * IP_EC + 1;

__device__ __forceinline__ int ip_filter_strength(int wh, int angle, int is_sm) {   // src/ipred_tmpl.c:327-360
    if (is_sm) {
        if (wh <= 8) { if (angle >= 64) return 2; if (angle >= 40) return 1; }
        else if (wh <= 16) { if (angle >= 48) return 2; if (angle >= 20) return 1; }
        else if (wh <= 24) { if (angle >= 4) return 3; }
        else return 3;
    } else {
        if (wh <= 8) { if (angle >= 56) return 1; }
        else if (wh <= 16) { if (angle >= 40) return 1; }
        else if (wh <= 24) { if (angle >= 32) return 3; if (angle >= 16) return 2; if (angle >= 8) return 1; }
        else if (wh <= 32) { if (angle >= 32) return 3; if (angle >= 4) return 2; return 1; }
        else return 3;
    }
    return 0;
}
__device__ __forceinline__ int ip_upsample(int wh, int angle, int is_sm) { return angle < 40 && wh <= (16 >> is_sm); }

// out[i], i < sz: the 5-tap smoothing of in[] inside [lim_from, lim_to), a clamped copy outside (filter_edge, :362-385)
template <typename P>
__device__ void ip_filter_edge(P *out, int sz, int lim_from, int lim_to, const P *in, int from, int to, int strength) {
    const int k0 = strength == 3 ? 2 : 0, k1 = strength == 1 ? 4 : (strength == 2 ? 5 : 4), k2 = strength == 1 ? 8 : (strength == 2 ? 6 : 4);
    for (int i = threadIdx.x; i < sz; i += blockDim.x) {
        int v;
        if (i < imin(sz, lim_from) || i >= imin(lim_to, sz)) {
            v = in[iclip(i, from, to - 1)];
        } else {
            const int s = k0 * (in[iclip(i - 2, from, to - 1)] + in[iclip(i + 2, from, to - 1)]) +
                          k1 * (in[iclip(i - 1, from, to - 1)] + in[iclip(i + 1, from, to - 1)]) + k2 * in[iclip(i, from, to - 1)];
            v = (s + 8) >> 4;
        }
        out[i] = (P)v;
    }
}
// out[0 .. 2 * hsz - 2]: in[] with a (-1, 9, 9, -1) / 16 sample between neighbours (upsample_edge, :391-406)
template <typename P>
__device__ void ip_upsample_edge(P *out, int hsz, const P *in, int from, int to, int bdmax) {
    for (int i = threadIdx.x; i < hsz; i += blockDim.x) {
        out[2 * i] = in[iclip(i, from, to - 1)];
        if (i < hsz - 1) {
            const int s = 9 * (in[iclip(i, from, to - 1)] + in[iclip(i + 1, from, to - 1)]) - in[iclip(i - 1, from, to - 1)] -
                          in[iclip(i + 2, from, to - 1)];
            out[2 * i + 1] = (P)iclip((s + 8) >> 4, 0, bdmax);
        }
    }
}

// Scratch of one block prediction (shared memory of the calling CTA).
template <typename P>
struct IpScratch {
    P e[IP_EDGE];             // the block's edge, topleft at e[IP_EC]
    P work[2 * IP_EC + 2];    // prepared edge of the directional modes
    P tile[33 * 33];          // filter-intra block with its top row / left column border; inter-intra prediction
    int dc;
    // the predictors' tables, copied once per kernel (ip_load_tables): the level waits of the frame-level wavefront
    // invalidate L1, and a table fetched from L2 after the wait would sit on the critical path of every level
    uint8_t sm_w[128];
    uint16_t dr[44];
    int8_t flt[320];
};
template <typename P>
__device__ __forceinline__ void ip_load_tables(IpScratch<P> &S) {      // followed by a __syncthreads of the caller
    for (int i = threadIdx.x; i < 128; i += blockDim.x) S.sm_w[i] = tab::k_sm_weights[i];
    for (int i = threadIdx.x; i < 44; i += blockDim.x) S.dr[i] = tab::k_dr_intra_derivative[i];
    for (int i = threadIdx.x; i < 320; i += blockDim.x) S.flt[i] = tab::k_filter_intra_taps[i];
}

// Predict one w x h block at dst8 from the edge in S.e (all threads of the CTA take part; S.e is complete and
// visible on entry).  mode: IP_*; angle: the reference's packed argument.
template <typename BD>
__device__ void ipred_block(IpScratch<typename BD::pixel> &S, int mode, uint8_t *dst8, int64_t stride, int w, int h, int angle,
                            int max_w, int max_h, int bdmax, const int16_t *cfl_ac = nullptr, int cfl_alpha = 0) {
    using pixel = typename BD::pixel;
    pixel *work = S.work, *tile = S.tile;
    int &dc_s = S.dc;
    const int tid = threadIdx.x;
    const pixel *tl = S.e + IP_EC;
    auto put = [&](int x, int y, int v) { ((pixel *)(dst8 + (int64_t)y * stride))[x] = (pixel)v; };
    const int n = w * h, lw = 31 - __clz(w);       // block sizes are powers of two: no integer divisions on this path

    switch (mode) {
    case IP_DC: case IP_TOP_DC: case IP_LEFT_DC: case IP_DC_128: {
        // the edge sums by the first warp (every caller runs at least one full warp), the rest by its first lane
        unsigned top_sum = 0, left_sum = 0;
        if (tid < 32 && mode != IP_DC_128) {
            if (mode != IP_LEFT_DC) for (int i = tid; i < w; i += 32) top_sum += tl[1 + i];
            if (mode != IP_TOP_DC) for (int i = tid; i < h; i += 32) left_sum += tl[-(1 + i)];
            for (int o = 16; o; o >>= 1) {
                top_sum += __shfl_xor_sync(0xffffffffu, top_sum, o);
                left_sum += __shfl_xor_sync(0xffffffffu, left_sum, o);
            }
        }
        if (tid == 0) {
            unsigned dc;
            if (mode == IP_DC_128) {
                dc = BD::hbd ? (unsigned)(bdmax + 1) >> 1 : 128;
            } else if (mode == IP_TOP_DC) {
                dc = ((w >> 1) + top_sum) >> ulog2(w);
            } else if (mode == IP_LEFT_DC) {
                dc = ((h >> 1) + left_sum) >> ulog2(h);
            } else {   // dc_gen, src/ipred_tmpl.c:150-166
                dc = ((w + h) >> 1) + top_sum + left_sum;
                dc >>= __ffs(w + h) - 1;
                if (w != h) {
                    const bool x4 = w > h * 2 || h > w * 2;
                    dc *= BD::hbd ? (x4 ? 0x6667u : 0xAAABu) : (x4 ? 0x3334u : 0x5556u);
                    dc >>= BD::hbd ? 17 : 16;
                }
            }
            dc_s = (int)dc;
        }
        __syncthreads();
        if (cfl_ac) {   // chroma from luma: dc + alpha * ac (cfl_pred, src/ipred_tmpl.c:71-84)
            for (int i = tid; i < n; i += blockDim.x) {
                const int diff = cfl_alpha * cfl_ac[i];
                const int m = (abs(diff) + 32) >> 6;
                put((i & (w - 1)), (i >> lw), iclip(dc_s + (diff < 0 ? -m : m), 0, bdmax));
            }
        } else {
            for (int i = tid; i < n; i += blockDim.x) put((i & (w - 1)), (i >> lw), dc_s);
        }
        break;
    }
    case IP_VERT:
        for (int i = tid; i < n; i += blockDim.x) put((i & (w - 1)), (i >> lw), tl[1 + (i & (w - 1))]);
        break;
    case IP_HOR:
        for (int i = tid; i < n; i += blockDim.x) put((i & (w - 1)), (i >> lw), tl[-(1 + (i >> lw))]);
        break;
    case IP_PAETH:
        for (int i = tid; i < n; i += blockDim.x) {
            const int x = (i & (w - 1)), y = (i >> lw);
            const int left = tl[-(y + 1)], top = tl[1 + x], topleft = tl[0];
            const int base = left + top - topleft;
            const int ld = abs(left - base), td = abs(top - base), tld = abs(topleft - base);
            put(x, y, ld <= td && ld <= tld ? left : (td <= tld ? top : topleft));
        }
        break;
    case IP_SMOOTH: case IP_SMOOTH_V: case IP_SMOOTH_H: {
        const uint8_t *wh = S.sm_w + w, *wv = S.sm_w + h;
        const int right = tl[w], bottom = tl[-h];
        for (int i = tid; i < n; i += blockDim.x) {
            const int x = (i & (w - 1)), y = (i >> lw);
            int v;
            if (mode == IP_SMOOTH)
                v = (wv[y] * tl[1 + x] + (256 - wv[y]) * bottom + wh[x] * tl[-(1 + y)] + (256 - wh[x]) * right + 256) >> 9;
            else if (mode == IP_SMOOTH_V)
                v = (wv[y] * tl[1 + x] + (256 - wv[y]) * bottom + 128) >> 8;
            else
                v = (wh[x] * tl[-(1 + y)] + (256 - wh[x]) * right + 128) >> 8;
            put(x, y, v);
        }
        break;
    }
    case IP_Z1: {   // src/ipred_tmpl.c:408-460
        const int is_sm = (angle >> 9) & 1, eief = angle >> 10;
        angle &= 511;
        int dx = S.dr[angle >> 1];
        const int ups = eief ? ip_upsample(w + h, 90 - angle, is_sm) : 0;
        const pixel *top;
        int max_base_x;
        if (ups) {
            ip_upsample_edge(work, w + h, tl + 1, -1, w + imin(w, h), bdmax);
            top = work; max_base_x = 2 * (w + h) - 2; dx <<= 1;
        } else {
            const int fs = eief ? ip_filter_strength(w + h, 90 - angle, is_sm) : 0;
            if (fs) {
                ip_filter_edge(work, w + h, 0, w + h, tl + 1, -1, w + imin(w, h), fs);
                top = work; max_base_x = w + h - 1;
            } else {
                top = tl + 1; max_base_x = w + imin(w, h) - 1;
            }
        }
        __syncthreads();
        const int inc = 1 + ups;
        for (int i = tid; i < n; i += blockDim.x) {
            const int x = (i & (w - 1)), y = (i >> lw);
            const int xpos = dx * (y + 1), frac = xpos & 0x3E, base = (xpos >> 6) + inc * x;
            put(x, y, base < max_base_x ? (top[base] * (64 - frac) + top[base + 1] * frac + 32) >> 6 : top[max_base_x]);
        }
        break;
    }
    case IP_Z2: {   // src/ipred_tmpl.c:462-540
        const int is_sm = (angle >> 9) & 1, eief = angle >> 10;
        angle &= 511;
        int dy = S.dr[(angle - 90) >> 1], dx = S.dr[(180 - angle) >> 1];
        const int ul = eief ? ip_upsample(w + h, 180 - angle, is_sm) : 0;
        const int ua = eief ? ip_upsample(w + h, angle - 90, is_sm) : 0;
        pixel *t2 = work + IP_EC;                 // the prepared corner: t2[0] = topleft
        if (ua) {
            ip_upsample_edge(t2, w + 1, tl, 0, w + 1, bdmax);
            dx <<= 1;
        } else {
            const int fs = eief ? ip_filter_strength(w + h, angle - 90, is_sm) : 0;
            if (fs) ip_filter_edge(t2 + 1, w, 0, max_w, tl + 1, -1, w, fs);
            else for (int i = tid; i < w; i += blockDim.x) t2[1 + i] = tl[1 + i];
        }
        if (ul) {
            ip_upsample_edge(t2 - 2 * h, h + 1, tl - h, 0, h + 1, bdmax);
            dy <<= 1;
        } else {
            const int fs = eief ? ip_filter_strength(w + h, 180 - angle, is_sm) : 0;
            if (fs) ip_filter_edge(t2 - h, h, h - max_h, h, tl - h, 0, h + 1, fs);
            else for (int i = tid; i < h; i += blockDim.x) t2[-h + i] = tl[-h + i];
        }
        __syncthreads();
        if (tid == 0) t2[0] = tl[0];
        __syncthreads();
        const int incx = 1 + ua;
        const pixel *left = t2 - (1 + ul);
        for (int i = tid; i < n; i += blockDim.x) {
            const int x = (i & (w - 1)), y = (i >> lw);
            const int xpos = ((1 + ua) << 6) - dx * (y + 1);
            const int base_x = (xpos >> 6) + incx * x;
            int v;
            if (base_x >= 0) {
                const int fx = xpos & 0x3E;
                v = t2[base_x] * (64 - fx) + t2[base_x + 1] * fx;
            } else {
                const int ypos = (y << (6 + ul)) - dy * (x + 1);
                const int base_y = ypos >> 6, fy = ypos & 0x3E;
                v = left[-base_y] * (64 - fy) + left[-(base_y + 1)] * fy;
            }
            put(x, y, (v + 32) >> 6);
        }
        break;
    }
    case IP_Z3: {   // src/ipred_tmpl.c:542-600
        const int is_sm = (angle >> 9) & 1, eief = angle >> 10;
        angle &= 511;
        int dy = S.dr[(270 - angle) >> 1];
        const int ups = eief ? ip_upsample(w + h, angle - 180, is_sm) : 0;
        const pixel *left;
        int max_base_y;
        if (ups) {
            ip_upsample_edge(work, w + h, tl - (w + h), imax(w - h, 0), w + h + 1, bdmax);
            left = work + 2 * (w + h) - 2; max_base_y = 2 * (w + h) - 2; dy <<= 1;
        } else {
            const int fs = eief ? ip_filter_strength(w + h, angle - 180, is_sm) : 0;
            if (fs) {
                ip_filter_edge(work, w + h, 0, w + h, tl - (w + h), imax(w - h, 0), w + h + 1, fs);
                left = work + w + h - 1; max_base_y = w + h - 1;
            } else {
                left = tl - 1; max_base_y = h + imin(w, h) - 1;
            }
        }
        __syncthreads();
        const int inc = 1 + ups;
        for (int i = tid; i < n; i += blockDim.x) {
            const int x = (i & (w - 1)), y = (i >> lw);
            const int ypos = dy * (x + 1), frac = ypos & 0x3E, base = (ypos >> 6) + inc * y;
            put(x, y, base < max_base_y ? (left[-base] * (64 - frac) + left[-(base + 1)] * frac + 32) >> 6 : left[-max_base_y]);
        }
        break;
    }
    case IP_FILTER: {   // src/ipred_tmpl.c:618-655; up to 32x32
        const int8_t *flt = S.flt + (angle & 511) * 64 + (tid & 7);      // this thread's output of the 4x2 sub-block
        const int f0 = flt[0], f1 = flt[8], f2 = flt[16], f3 = flt[24], f4 = flt[32], f5 = flt[40], f6 = flt[48];
        const int nbx = w >> 2, nby = h >> 1;
        // the tile carries the edge as its row -1 / column -1, so the seven inputs of a sub-block are plain loads
        constexpr int TS = 33;
        pixel *tb = tile + TS + 1;                                     // (0, 0) of the block
        for (int i = tid; i <= w; i += blockDim.x) tile[i] = tl[i];    // row -1: topleft, top
        for (int i = tid; i < h; i += blockDim.x) tile[(i + 1) * TS] = tl[-(1 + i)];
        __syncthreads();
        for (int t = 0; t < nbx + nby - 1; t++) {
            // sub-blocks on the anti-diagonal bx + by = t; 8 threads (outputs) per sub-block
            const int sb = tid >> 3, o = tid & 7;
            for (int by = sb; by < nby; by += blockDim.x >> 3) {
                const int bx = t - by;
                if (bx < 0 || bx >= nbx) continue;
                const int x0 = 4 * bx, y0 = 2 * by;
                auto px = [&](int x, int y) -> int { return tb[y * TS + x]; };
                const int p0 = px(x0 - 1, y0 - 1), p1 = px(x0, y0 - 1), p2 = px(x0 + 1, y0 - 1), p3 = px(x0 + 2, y0 - 1),
                          p4 = px(x0 + 3, y0 - 1), p5 = px(x0 - 1, y0), p6 = px(x0 - 1, y0 + 1);
                const int acc = f0 * p0 + f1 * p1 + f2 * p2 + f3 * p3 + f4 * p4 + f5 * p5 + f6 * p6;
                tb[(y0 + (o >> 2)) * TS + x0 + (o & 3)] = (pixel)iclip((acc + 8) >> 4, 0, bdmax);
            }
            __syncthreads();
        }
        for (int i = tid; i < n; i += blockDim.x) put((i & (w - 1)), (i >> lw), tb[((i >> lw)) * TS + (i & (w - 1))]);
        break;
    }
    default: break;
    }
}

template <typename BD>
__global__ void __launch_bounds__(256)
ipred_kernel(int mode, uint8_t *dst8, int64_t stride, const typename BD::pixel *__restrict__ edge_in, int lo, int hi, int w,
             int h, int angle, int max_w, int max_h, int bdmax) {
    __shared__ IpScratch<typename BD::pixel> S;
    for (int i = threadIdx.x; i < lo + hi + 1; i += blockDim.x) S.e[IP_EC - lo + i] = edge_in[i];
    ip_load_tables(S);
    __syncthreads();
    ipred_block<BD>(S, mode, dst8, stride, w, h, angle, max_w, max_h, bdmax);
}

// ---------------------------------------------------------------- frame level: one wavefront level of intra blocks
// rav1d_prepare_intra_edges (src/ipred_prepare.rs:118-330 == src/ipred_prepare_tmpl.c:77-204) restated per thread
// index: which neighbouring pixels of the reconstructed picture form the edge, how the unavailable parts are
// replicated, and which implementation mode the coded mode becomes; then the block is predicted in place.
// The picture rows above a superblock row are still unfiltered when this runs (the in-loop filters come after
// the whole reconstruction), so the reference's saved pre-filter edge (f.ipred_edge) is the picture itself.
template <typename P>
struct IntraSmem {
    IpScratch<P> S;
    int itile[65 * 32];          // the residual's transform tile (largest: 64 x 32 + padding)
    int res_s[64 * 64];          // the residual itself, computed while waiting for the neighbours
    int16_t ac_s[32 * 32];       // chroma-from-luma: the sub-sampled, zero-mean luma of the block
    int red_s[4];
};
// Picture reads of the persistent kernel bypass L1 (CG): another SM wrote those pixels during this very launch.
template <bool CG, typename T>
__device__ __forceinline__ T pic_ld(const T *p) { return CG ? __ldcg(p) : *p; }

// cfl_ac's sub-sampling sum of one output (src/ipred_tmpl.c:670-685): 1, 2 or 4 luma pixels
template <bool CG, typename P>
__device__ __forceinline__ int cfl_luma_sum(const uint8_t *ypx, int64_t lstride, int cx, int cy, int ss_hor, int ss_ver) {
    const P *p = (const P *)(ypx + (int64_t)(cy << ss_ver) * lstride) + (cx << ss_hor);
    int sum = pic_ld<CG>(&p[0]);
    if (ss_hor) sum += pic_ld<CG>(&p[1]);
    if (ss_ver) {
        const P *q = (const P *)((const uint8_t *)p + lstride);
        sum += pic_ld<CG>(&q[0]);
        if (ss_hor) sum += pic_ld<CG>(&q[1]);
    }
    return sum;
}

// One intra item, all threads of the CTA.  `wait()` is called once, after everything that does not depend on the picture
// (the item's set-up and the residual's transform) and before the first picture read: it returns when every item of the
// earlier levels is in the picture.
template <typename BD, bool CG, typename WaitFn>
__device__ __forceinline__ void intra_item(IntraSmem<typename BD::pixel> &M, const Rb200Planes &cur, const Rb200IntraItem it, const int ti,
                                           const Rb200ItxItem *__restrict__ itx, const typename BD::coef *__restrict__ cf,
                                           const uint8_t *__restrict__ pal_buf, int frame_w4, int frame_h4, int ss_hor_c, int ss_ver_c,
                                           int bdmax, WaitFn wait) {
    using pixel = typename BD::pixel;
    IpScratch<pixel> &S = M.S;
    int (&itile)[65 * 32] = M.itile;
    int *res_s = M.res_s;
    int16_t *ac_s = M.ac_s;
    int *red_s = M.red_s;
    const int tid = threadIdx.x;
    const int ss_hor = it.plane ? ss_hor_c : 0, ss_ver = it.plane ? ss_ver_c : 0;
    const int64_t stride = plane_stride(cur, it.plane);
    uint8_t *dst8 = plane_ptr(cur, it.plane) + (int64_t)(it.y4 * 4) * stride + (int64_t)(it.x4 * 4) * sizeof(pixel);
    const int64_t ps = stride / (int64_t)sizeof(pixel);
    const pixel *dst = (const pixel *)dst8;
    const int have_left = it.flags & 1, have_top = (it.flags >> 1) & 1;
    const int top_has_right = (it.flags >> 2) & 1, left_has_bottom = (it.flags >> 3) & 1;
    const int is_sm = (it.flags >> 4) & 1, eief = (it.flags >> 5) & 1;
    const int tw = it.tw4, th = it.th4, x = it.x4, y = it.y4, w = it.w4_end & 0x1fff, h = it.h4_end & 0x1fff;
    const bool cfl = it.plane && it.mode == 13;     // UV_CFL_PRED shares the number of luma's FILTER_PRED
    const bool inter_intra = (it.flags >> 6) & 1;
    const int bitdepth = BD::hbd ? bpc_from_max(bdmax) : 8;
    // ---- coded mode -> implementation mode (src/ipred_prepare_tmpl.c:89-116); coded numbering: DC 0, VERT 1, HOR 2,
    // DIAG_DOWN_LEFT 3, DIAG_DOWN_RIGHT 4, VERT_RIGHT 5, HOR_DOWN 6, HOR_UP 7, VERT_LEFT 8, SMOOTH 9 .. PAETH 12, FILTER 13
    int mode = cfl ? IP_DC : it.mode, angle = (cfl || inter_intra) ? 0 : it.angle;
    if (mode >= 1 && mode <= 8) {
        const int map[8] = { 90, 180, 45, 135, 113, 157, 203, 67 };
        angle = map[mode - 1] + 3 * angle;
        if (angle <= 90) mode = (angle < 90 && have_top) ? IP_Z1 : IP_VERT;
        else if (angle < 180) mode = IP_Z2;
        else mode = (angle > 180 && have_left) ? IP_Z3 : IP_HOR;
    } else if (mode == IP_DC) {
        mode = have_left ? (have_top ? IP_DC : IP_LEFT_DC) : (have_top ? IP_TOP_DC : IP_DC_128);
    } else if (mode == IP_PAETH) {
        mode = have_left ? (have_top ? IP_PAETH : IP_HOR) : (have_top ? IP_VERT : IP_DC_128);
    }
    // needs_{left, top, topleft, topright, bottomleft} per implementation mode (:52-75)
    const bool n_left = mode == IP_DC || mode == IP_HOR || mode == IP_LEFT_DC || mode == IP_Z2 || mode == IP_Z3 || (mode >= IP_SMOOTH && mode <= IP_FILTER);
    const bool n_top = mode == IP_DC || mode == IP_VERT || mode == IP_TOP_DC || mode == IP_Z1 || mode == IP_Z2 || (mode >= IP_SMOOTH && mode <= IP_FILTER);
    const bool n_tl = mode == IP_Z1 || mode == IP_Z2 || mode == IP_Z3 || mode == IP_PAETH || mode == IP_FILTER;
    const bool n_tr = mode == IP_Z1, n_bl = mode == IP_Z3;
    const pixel *dst_top = dst - ps;      // only dereferenced when have_top
    pixel *tl = S.e + IP_EC;
    // The residual depends on the coefficients only: both transform passes run BEFORE the wait, i.e. while the previous
    // level is still reconstructing, and leave the term to add in shared memory.  After the wait the critical path is
    // edge -> prediction -> add.
    int res_w = 0, res_h = 0, res_plane = 0, res_x = 0, res_y = 0;
    if (ti >= 0) {
        const Rb200ItxItem t = itx[ti];
        res_plane = t.plane; res_x = t.x; res_y = t.y;
        switch (t.tx) {
#define CASE(TX) case TX: itx_residual_block<BD, TX>(itile, res_s, tid, tid < ItxGeom<TX>::T, t, cf, bdmax); res_w = tx_w(TX); res_h = tx_h(TX); break;
            CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9)
            CASE(10) CASE(11) CASE(12) CASE(13) CASE(14) CASE(15) CASE(16) CASE(17) CASE(18)
#undef CASE
        default: break;
        }
    }
    // dst = clip(dst + residual), rows of the block coalesced (call after a __syncthreads that follows the prediction)
    auto add_residual = [&]() {
        const int64_t rstride = plane_stride(cur, res_plane);
        uint8_t *rbase = plane_ptr(cur, res_plane) + (int64_t)res_y * rstride;
        const int lw = 31 - __clz(res_w);
        for (int i = tid; i < res_w * res_h; i += blockDim.x) {
            pixel *d = (pixel *)(rbase + (int64_t)(i >> lw) * rstride) + res_x + (i & (res_w - 1));
            *d = (pixel)iclip((int)pic_ld<CG>(d) + res_s[i], 0, bdmax);
        }
    };
    wait();                                                 // everything before this level is in the picture now
    if (it.mode >= 14) {
        // 14: palette block (pal_pred, src/ipred_tmpl.c:717-729): w4_end | h4_end << 16 is the offset, in 16-byte units, of
        // { 8 palette entries (16 bytes), w * h index bytes } in the frame's palette buffer.  15: no prediction, only the
        // residual (the further transform blocks of a palette block).
        if (it.mode == 14) {
            const uint8_t *rec = pal_buf + ((size_t)it.w4_end | (size_t)it.h4_end << 16) * 16;
            const pixel *pal = (const pixel *)rec;
            const uint8_t *idx = rec + 16;
            const int lpw = 31 - __clz(tw * 4);
            for (int i = tid; i < tw * th * 16; i += blockDim.x)
                ((pixel *)(dst8 + (int64_t)((i >> lpw)) * stride))[(i & (tw * 4 - 1))] = pal[idx[i] & 7];
        }
        // 16: intra block copy (src/recon_tmpl.c:1631-1645): mc() with the bilinear filter from the picture being
        // reconstructed.  (int16) w4_end / h4_end = the source position in plane pixels, angle = mx | my << 4 in 1/16
        // pixels (non-zero only for sub-sampled chroma); source pixels clamp to the picture rounded up to 8 luma pixels
        // (emu_edge with w = f.bw * 4 >> ss_hor, :988-989).  put_bilin_c (src/mc_tmpl.c:146-203) per pixel.
        if (it.mode == 16) {
            const int sx = (int16_t)it.w4_end, sy = (int16_t)it.h4_end;
            const int mx = (uint8_t)it.angle & 15, my = ((uint8_t)it.angle >> 4) & 15;
            const int pw = (frame_w4 * 4) >> ss_hor, ph = (frame_h4 * 4) >> ss_ver;
            const uint8_t *sbase = plane_ptr(cur, it.plane);
            const int ib = BD::hbd ? 14 - bitdepth : 4;
            auto px = [&](int xx, int yy) -> int {
                xx = iclip(xx, 0, pw - 1); yy = iclip(yy, 0, ph - 1);
                return (int)pic_ld<CG>((const pixel *)(sbase + (int64_t)yy * stride) + xx);
            };
            const int bw = tw * 4, n = bw * th * 4, lbw = 31 - __clz(bw);
            for (int i = tid; i < n; i += blockDim.x) {
                const int xx = sx + (i & (bw - 1)), yy = sy + (i >> lbw);
                int v;
                if (mx) {
                    const int a = px(xx, yy), b = px(xx + 1, yy);
                    const int m0 = (16 * a + mx * (b - a) + ((1 << (4 - ib)) >> 1)) >> (4 - ib);
                    if (my) {
                        const int c = px(xx, yy + 1), d = px(xx + 1, yy + 1);
                        const int m1 = (16 * c + mx * (d - c) + ((1 << (4 - ib)) >> 1)) >> (4 - ib);
                        v = (16 * m0 + my * (m1 - m0) + ((1 << (4 + ib)) >> 1)) >> (4 + ib);
                    } else {
                        v = (m0 + ((1 << ib) >> 1)) >> ib;
                    }
                } else if (my) {
                    const int a = px(xx, yy), c = px(xx, yy + 1);
                    v = (16 * a + my * (c - a) + 8) >> 4;
                } else {
                    v = px(xx, yy);
                }
                ((pixel *)(dst8 + (int64_t)((i >> lbw)) * stride))[(i & (bw - 1))] = (pixel)iclip(v, 0, bdmax);
            }
        }
        if (ti < 0) return;
        __syncthreads();
        add_residual();
        return;
    }
    // ---- the edge: left column, top row and their bottom-left / top-right extensions.  Every segment is at most 64
    // pixels, one per thread; all picture loads of a thread are issued back to back into registers (one global round
    // trip for the whole edge), then stored; only the replication of a missing extension waits for them.
    const int szl = th << 2, szt = tw << 2;
    int cfl_first[2] = { 0, 0 };
    const int have_bl = (n_left && n_bl && have_left && y + th < h) ? left_has_bottom : 0;
    const int have_tr = (n_top && n_tr && have_top && x + tw < w) ? top_has_right : 0;
    {
        const bool ld_l = n_left && have_left && tid < szl, ld_t = n_top && have_top && tid < szt;
        const bool ld_bl = have_bl && tid < szl, ld_tr = have_tr && tid < szt;
        pixel r_l = 0, r_t = 0, r_bl = 0, r_tr = 0, r_fl = 0, r_ft = 0;
        if (cfl) {
            // chroma from luma: the block's own luma (this thread's first two outputs) in the same round trip as the edge
            const int w_pad = it.w4_end >> 13, h_pad = it.h4_end >> 13;
            const int cw = tw * 4, ch = th * 4, vw = cw - 4 * w_pad, vh = ch - 4 * h_pad, lcw = 31 - __clz(cw);
            const int64_t lstride = plane_stride(cur, 0);
            const uint8_t *ypx = plane_ptr(cur, 0) + (int64_t)((y << ss_ver) * 4) * lstride + (int64_t)((x << ss_hor) * 4) * sizeof(pixel);
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const int i = tid + k * (int)blockDim.x;
                if (i < cw * ch) cfl_first[k] = cfl_luma_sum<CG, pixel>(ypx, lstride, imin((i & (cw - 1)), vw - 1), imin((i >> lcw), vh - 1), ss_hor, ss_ver);
            }
        }
        if (ld_l) r_l = pic_ld<CG>(&dst[(int64_t)imin(tid, imin(szl, (h - y) << 2) - 1) * ps - 1]);
        if (ld_t) r_t = pic_ld<CG>(&dst_top[imin(tid, imin(szt, (w - x) << 2) - 1)]);
        if (ld_bl) r_bl = pic_ld<CG>(&dst[(int64_t)(szl + imin(tid, imin(szl, (h - y - th) << 2) - 1)) * ps - 1]);
        if (ld_tr) r_tr = pic_ld<CG>(&dst_top[szt + imin(tid, imin(szt, (w - x - tw) << 2) - 1)]);
        if (n_left && !have_left) r_fl = have_top ? pic_ld<CG>(&dst_top[0]) : (pixel)(((1 << bitdepth) >> 1) + 1);
        if (n_top && !have_top) r_ft = have_left ? pic_ld<CG>(&dst[-1]) : (pixel)(((1 << bitdepth) >> 1) - 1);
        if (n_left && tid < szl) tl[-(1 + tid)] = have_left ? r_l : r_fl;
        if (n_top && tid < szt) tl[1 + tid] = have_top ? r_t : r_ft;
        if (ld_bl) tl[-(szl + 1 + tid)] = r_bl;
        if (ld_tr) tl[1 + szt + tid] = r_tr;
    }
    int corner = 0;
    if (n_tl && tid == 0) {
        if (have_left) corner = have_top ? pic_ld<CG>(&dst_top[-1]) : pic_ld<CG>(&dst[-1]);
        else corner = have_top ? pic_ld<CG>(&dst_top[0]) : (1 << bitdepth) >> 1;
    }
    __syncthreads();
    if (n_left && n_bl && !have_bl) {
        const int sz = th << 2;
        const pixel v = tl[-sz];
        for (int i = tid; i < sz; i += blockDim.x) tl[-(sz + 1 + i)] = v;
    }
    if (n_top && n_tr && !have_tr) {
        const int sz = tw << 2;
        const pixel v = tl[sz];
        for (int i = tid; i < sz; i += blockDim.x) tl[1 + sz + i] = v;
    }
    if (n_tl && tid == 0) {
        int v = corner;
        if (mode == IP_Z2 && tw + th >= 6 && eief) v = ((tl[-1] + tl[1]) * 5 + v * 6 + 8) >> 4;
        tl[0] = (pixel)v;
    }
    __syncthreads();
    const int max_w = ((frame_w4 * 4) >> ss_hor) - 4 * x, max_h = ((frame_h4 * 4) >> ss_ver) - 4 * y;
    if (cfl) {
        // cfl_ac (src/ipred_tmpl.c:657-703) on the block's reconstructed luma; the padding counts ride in the top bits
        // of w4_end / h4_end
        const int w_pad = it.w4_end >> 13, h_pad = it.h4_end >> 13;
        const int cw = tw * 4, ch = th * 4, vw = cw - 4 * w_pad, vh = ch - 4 * h_pad, lcw = 31 - __clz(cw);
        const int64_t lstride = plane_stride(cur, 0);
        const uint8_t *ypx = plane_ptr(cur, 0) + (int64_t)((y << ss_ver) * 4) * lstride + (int64_t)((x << ss_hor) * 4) * sizeof(pixel);
        int part = 0;
        for (int i = tid, k = 0; i < cw * ch; i += blockDim.x, k++) {
            int sum = k == 0 ? cfl_first[0] : (k == 1 ? cfl_first[1] : cfl_luma_sum<CG, pixel>(ypx, lstride, imin((i & (cw - 1)), vw - 1), imin((i >> lcw), vh - 1), ss_hor, ss_ver));
            sum <<= 1 + !ss_ver + !ss_hor;
            ac_s[i] = (int16_t)sum;
            part += sum;
        }
        for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if ((tid & 31) == 0) red_s[tid >> 5] = part;
        __syncthreads();
        const int log2sz = ulog2(cw) + ulog2(ch);
        const int mean = (((1 << log2sz) >> 1) + red_s[0] + red_s[1] + red_s[2] + red_s[3]) >> log2sz;
        for (int i = tid; i < cw * ch; i += blockDim.x) ac_s[i] = (int16_t)(ac_s[i] - mean);
        __syncthreads();
        if (ti >= 0 && res_plane == it.plane && res_x == x * 4 && res_y == y * 4 && res_w == cw && res_h == ch) {
            // with a residual over the whole block: prediction into shared memory, one write (as for the plain modes below)
            pixel *pred = reinterpret_cast<pixel *>(itile);
            ipred_block<BD>(S, mode, (uint8_t *)pred, (int64_t)cw * sizeof(pixel), cw, ch, 0, max_w, max_h, bdmax, ac_s, it.angle);
            __syncthreads();
            for (int i = tid; i < cw * ch; i += blockDim.x)
                ((pixel *)(dst8 + (int64_t)(i >> lcw) * stride))[i & (cw - 1)] = (pixel)iclip((int)pred[i] + res_s[i], 0, bdmax);
            return;
        }
        ipred_block<BD>(S, mode, dst8, stride, cw, ch, 0, max_w, max_h, bdmax, ac_s, it.angle);
    } else if (inter_intra) {
        // inter-intra (src/recon.rs:3475-3550, chroma :3742-3850): the intra prediction of the whole block goes into shared
        // memory and is blended over the inter prediction already in the picture, with the inter-intra mask of the
        // mode (it.angle < 0) or a wedge mask (it.angle = wedge index)
        const int bw = tw * 4, bh = th * 4, lbw = 31 - __clz(bw);
        ipred_block<BD>(S, mode, (uint8_t *)S.tile, (int64_t)bw * sizeof(pixel), bw, bh, 0, 0, 0, bdmax);
        __syncthreads();
        const int ii_mode = it.mode == 9 ? 3 : it.mode;          // SMOOTH_PRED is II_SMOOTH_PRED
        const int lw = bw << ss_hor, lh = bh << ss_ver;            // the luma block the wedge is defined on
        for (int i = tid; i < bw * bh; i += blockDim.x) {
            const int px = (i & (bw - 1)), py = (i >> lbw);
            int m;
            if (it.angle < 0) {
                m = ii_mask_at(bw, bh, ii_mode, px, py);
            } else {
                m = wedge_mask_at(lw, lh, it.angle & 15, px << ss_hor, py << ss_ver);
                if (ss_hor) {
                    m += wedge_mask_at(lw, lh, it.angle & 15, (px << 1) + 1, py << ss_ver) + 1;
                    if (ss_ver) m += wedge_mask_at(lw, lh, it.angle & 15, px << 1, (py << 1) + 1) +
                                     wedge_mask_at(lw, lh, it.angle & 15, (px << 1) + 1, (py << 1) + 1) + 1;
                    m >>= 1 + ss_ver;
                }
            }
            pixel *d = (pixel *)(dst8 + (int64_t)py * stride) + px;
            *d = (pixel)(((int)pic_ld<CG>(d) * (64 - m) + (int)S.tile[py * bw + px] * m + 32) >> 6);
        }
    } else if (ti >= 0 && res_plane == it.plane && res_x == x * 4 && res_y == y * 4 && res_w == tw * 4 && res_h == th * 4) {
        // the usual case, a transform block with a residual: predict into shared memory (the transform tile is free by
        // now) and write prediction + residual once, instead of writing the prediction and reading it back
        static_assert(sizeof(itile) >= 64 * 64 * sizeof(pixel), "prediction tile");
        pixel *pred = reinterpret_cast<pixel *>(itile);
        ipred_block<BD>(S, mode, (uint8_t *)pred, (int64_t)res_w * sizeof(pixel), res_w, res_h,
                        mode == IP_FILTER ? (it.angle & 7) : (angle | (is_sm << 9) | (eief << 10)), max_w, max_h, bdmax);
        __syncthreads();
        const int lw = 31 - __clz(res_w);
        for (int i = tid; i < res_w * res_h; i += blockDim.x)
            ((pixel *)(dst8 + (int64_t)(i >> lw) * stride))[i & (res_w - 1)] = (pixel)iclip((int)pred[i] + res_s[i], 0, bdmax);
        return;
    } else {
        ipred_block<BD>(S, mode, dst8, stride, tw * 4, th * 4, mode == IP_FILTER ? (it.angle & 7) : (angle | (is_sm << 9) | (eief << 10)),
                        max_w, max_h, bdmax);
    }
    // ---- the block's residual on top of its prediction, in the same launch (the prediction is visible to the CTA)
    if (ti < 0) return;
    __syncthreads();
    add_residual();
}

// (a) one launch per level, with programmatic dependent launch: the next level's grid is scheduled while this one runs,
// does its picture-independent part and waits for this grid to complete.  Kept for RB200_INTRA_LEVEL_LAUNCHES=1.
template <typename BD>
__global__ void __launch_bounds__(128)
intra_items_kernel(Rb200Planes cur, const Rb200IntraItem *__restrict__ items, const int32_t *__restrict__ itx_of,
                   const Rb200ItxItem *__restrict__ itx, const typename BD::coef *__restrict__ cf, const uint8_t *__restrict__ pal_buf,
                   int frame_w4, int frame_h4, int ss_hor_c, int ss_ver_c, int bdmax) {
    __shared__ IntraSmem<typename BD::pixel> M;
    asm volatile("griddepcontrol.launch_dependents;");
    ip_load_tables(M.S);
    __syncthreads();
    intra_item<BD, false>(M, cur, items[blockIdx.x], itx_of ? itx_of[blockIdx.x] : -1, itx, cf, pal_buf, frame_w4, frame_h4, ss_hor_c,
                          ss_ver_c, bdmax, [] { asm volatile("griddepcontrol.wait;" ::: "memory"); });
}

// (b) ALL levels in one cooperative launch: the grid (every CTA resident) walks the levels; CTA b takes the items
// i = b (mod gridDim) of the level-sorted list.  Instead of a kernel boundary per level there is one counter per level in global memory:
// the CTA that has written an item of level l adds one to done[l], and before the first picture read of a level-l item
// a CTA waits until done[l - 1] equals the number of items of level l - 1.  That is enough: by induction every item of
// level l - 1 was itself only started when level l - 2 was complete (levels are never empty -- the host drops empty
// ones).  CTAs without an item in a level neither wait nor count.  Between its items a CTA already transforms the
// residual of the next one.  Saves the launch gap per level and ~2,600 host launches per 4K key frame.
// sync[0] = set if a wait ever ran into its time limit (a bug, not a state: reported by rb200_frame_wait), sync[1 + l] = done[l].
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
template <typename BD>
__global__ void __launch_bounds__(128)
intra_levels_kernel(Rb200Planes cur, const Rb200IntraItem *__restrict__ items, const int32_t *__restrict__ itx_of,
                    const Rb200ItxItem *__restrict__ itx, const typename BD::coef *__restrict__ cf, const uint8_t *__restrict__ pal_buf,
                    const int32_t *__restrict__ level_off, int n_levels, int frame_w4, int frame_h4, int ss_hor_c, int ss_ver_c,
                    int bdmax, unsigned *sync, unsigned long long *trace) {
    __shared__ IntraSmem<typename BD::pixel> M;
    __shared__ int abort_s;
    const int tid = threadIdx.x;
    unsigned *done = sync + 1;
    if (tid == 0) abort_s = 0;
    ip_load_tables(M.S);
    __syncthreads();
    for (int l = 0; l < n_levels; l++) {
        const int beg = level_off[l], end = level_off[l + 1];
        bool waited = false;
        int cur_i = 0;
        auto level_wait = [&] {
            if (waited || l == 0) return;
            waited = true;
            if (tid == 0) {
                const unsigned target = (unsigned)(beg - level_off[l - 1]);
                const long long t0 = clock64();
                unsigned seen;
                // a wait that runs into its time limit (a bug, not a state) ends this CTA; the CTAs that depend on it run
                // into their own limit at about the same time, so the flag is not polled on the way
                // relaxed polls and ONE acquire fence at the end: an acquire load is LDG + CCTL.IVALL, i.e. every turn of the
                // loop would empty the L1 of an SM whose other CTAs are fetching their next items
                do {
                    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(done + l - 1) : "memory");
                    if (seen < target && clock64() - t0 > 4000000000LL) { atomicExch(sync, 1u); abort_s = 1; break; }
                } while (seen < target);
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
                if (trace) trace[2 * cur_i] = global_ns();      // RB200_INTRA_TRACE: when this item saw its level released ...
            }
            __syncthreads();
        };
        // Item i of the level-sorted list belongs to CTA i % grid: consecutive levels then sit side by side on the ring of
        // CTAs instead of all starting at CTA 0, so the CTAs of the next few levels have fetched their items and transformed
        // their residuals by the time this level completes -- the per-level period is wait -> edge -> prediction -> release,
        // not a whole item.
        const int G = (int)gridDim.x;
        int first = (int)blockIdx.x - beg % G;
        if (first < 0) first += G;
        for (int i = beg + first; i < end; i += G) {
            cur_i = i;
            intra_item<BD, true>(M, cur, items[i], itx_of ? itx_of[i] : -1, itx, cf, pal_buf, frame_w4, frame_h4, ss_hor_c, ss_ver_c, bdmax,
                                 level_wait);
            __syncthreads();            // the item is in the picture (as far as this CTA is concerned) and its shared memory is free
            if (abort_s) return;
            // release at gpu scope, cumulative over the barrier above: every thread's stores of this item are visible to
            // whoever acquires the counter -- no separate fence
            if (tid == 0) {
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(done + l) : "memory");
                if (trace) trace[2 * i + 1] = global_ns();          // ... and when it had released its own
            }
        }
    }
}

int intra_items_launch(const Rb200Planes &cur, const Rb200IntraItem *d_items, const int32_t *d_itx_of, const Rb200ItxItem *d_itx,
                       const void *cf, const uint8_t *d_pal, int n, int frame_w4, int frame_h4, int ss_hor, int ss_ver, int bdmax,
                       cudaStream_t st) {
    if (n <= 0) return 0;
    // launched with programmatic stream serialization: consecutive levels overlap their launch latency and set-up
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n); cfg.blockDim = dim3(128); cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (bdmax > 255) RB_CUDA(cudaLaunchKernelEx(&cfg, intra_items_kernel<BD16>, cur, d_items, d_itx_of, d_itx, (const int32_t *)cf, d_pal, frame_w4, frame_h4, ss_hor, ss_ver, bdmax));
    else RB_CUDA(cudaLaunchKernelEx(&cfg, intra_items_kernel<BD8>, cur, d_items, d_itx_of, d_itx, (const int16_t *)cf, d_pal, frame_w4, frame_h4, ss_hor, ss_ver, bdmax));
    RB_LAUNCH_CHECK();
    return 0;
}

// d_level_off: device, n_levels + 1 item offsets of NON-EMPTY levels; d_sync: device, n_levels + 1 words, zeroed here on the stream.
// max_items_per_level sizes the grid (no more CTAs than the widest level needs).
int intra_levels_launch(const Rb200Planes &cur, const Rb200IntraItem *d_items, const int32_t *d_itx_of, const Rb200ItxItem *d_itx,
                        const void *cf, const uint8_t *d_pal, const int32_t *d_level_off, int n_levels, int max_items_per_level,
                        int frame_w4, int frame_h4, int ss_hor, int ss_ver, int bdmax, unsigned *d_sync, cudaStream_t st) {
    if (n_levels <= 0) return 0;
    static int per_sm[2] = { 0, 0 }, n_sm = 0;
    const int hbd = bdmax > 255;
    if (!per_sm[hbd]) {
        int dev = 0, occ = 0;
        RB_CUDA(cudaGetDevice(&dev));
        RB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        if (hbd) RB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, intra_levels_kernel<BD16>, 128, 0));
        else RB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, intra_levels_kernel<BD8>, 128, 0));
        if (occ < 1) return set_error(-12, "intra_levels_launch: the kernel does not fit an SM");
        per_sm[hbd] = occ < 4 ? occ : 4;
    }
    // As many CTAs as fit (four per SM at 128 registers) unless RB200_INTRA_CTAS_PER_SM says otherwise, and never more
    // than the widest level has items: the more CTAs on the ring, the more levels are prepared ahead.
    static const int want = getenv("RB200_INTRA_CTAS_PER_SM") ? atoi(getenv("RB200_INTRA_CTAS_PER_SM")) : 4;
    int grid = imin(imax(want, 1), per_sm[hbd]) * n_sm;
    if (max_items_per_level > 0 && max_items_per_level < grid) grid = max_items_per_level;
    RB_CUDA(cudaMemsetAsync(d_sync, 0, ((size_t)n_levels + 1) * sizeof(unsigned), st));
    // RB200_INTRA_TRACE=<file> (a debugging aid): two time stamps per item -- level seen released, own release done -- written
    // to <file> after the launch (tools/intra_trace.py reads it)
    static const char *trace_path = getenv("RB200_INTRA_TRACE");
    unsigned long long *d_trace = nullptr;
    int32_t n_items_total = 0;
    if (trace_path) {
        RB_CUDA(cudaStreamSynchronize(st));
        RB_CUDA(cudaMemcpy(&n_items_total, d_level_off + n_levels, sizeof(int32_t), cudaMemcpyDeviceToHost));
        RB_CUDA(cudaMalloc(&d_trace, (size_t)n_items_total * 16));
        RB_CUDA(cudaMemsetAsync(d_trace, 0, (size_t)n_items_total * 16, st));
    }
    const int32_t *cf32 = (const int32_t *)cf; const int16_t *cf16 = (const int16_t *)cf;
    Rb200Planes cur_v = cur;
    void *args16[] = { &cur_v, &d_items, &d_itx_of, &d_itx, &cf32, &d_pal, &d_level_off, &n_levels, &frame_w4, &frame_h4, &ss_hor, &ss_ver, &bdmax, &d_sync, &d_trace };
    void *args8[] = { &cur_v, &d_items, &d_itx_of, &d_itx, &cf16, &d_pal, &d_level_off, &n_levels, &frame_w4, &frame_h4, &ss_hor, &ss_ver, &bdmax, &d_sync, &d_trace };
    if (hbd) RB_CUDA(cudaLaunchCooperativeKernel((const void *)intra_levels_kernel<BD16>, dim3(grid), dim3(128), args16, 0, st));
    else RB_CUDA(cudaLaunchCooperativeKernel((const void *)intra_levels_kernel<BD8>, dim3(grid), dim3(128), args8, 0, st));
    RB_LAUNCH_CHECK();
    if (d_trace) {
        std::vector<unsigned long long> h((size_t)n_items_total * 2);
        RB_CUDA(cudaStreamSynchronize(st));
        RB_CUDA(cudaMemcpy(h.data(), d_trace, h.size() * 8, cudaMemcpyDeviceToHost));
        RB_CUDA(cudaFree(d_trace));
        if (FILE *fp = fopen(trace_path, "wb")) { fwrite(h.data(), 8, h.size(), fp); fclose(fp); }
    }
    return 0;
}

// cfl_ac (src/ipred_tmpl.c:657-703): sub-sampled luma, edge replication of the padded part, zero mean
template <typename BD>
__global__ void __launch_bounds__(256)
cfl_ac_kernel(int16_t *__restrict__ ac, const uint8_t *__restrict__ ypx, int64_t stride, int w_pad, int h_pad, int width,
              int height, int ss_hor, int ss_ver) {
    using pixel = typename BD::pixel;
    __shared__ int16_t a[32 * 32];
    __shared__ int red[8];
    const int tid = threadIdx.x, n = width * height;
    const int vw = width - 4 * w_pad, vh = height - 4 * h_pad;
    int part = 0;
    for (int i = tid; i < n; i += blockDim.x) {
        const int x = imin(i % width, vw - 1), y = imin(i / width, vh - 1);     // padding repeats the last valid column / row
        const pixel *p = (const pixel *)(ypx + (int64_t)(y << ss_ver) * stride) + (x << ss_hor);
        int s = p[0];
        if (ss_hor) s += p[1];
        if (ss_ver) {
            const pixel *q = (const pixel *)((const uint8_t *)p + stride);
            s += q[0];
            if (ss_hor) s += q[1];
        }
        s <<= 1 + !ss_ver + !ss_hor;
        a[i] = (int16_t)s;
        part += s;
    }
    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((tid & 31) == 0) red[tid >> 5] = part;
    __syncthreads();
    const int log2sz = ulog2(width) + ulog2(height);
    int sum = (1 << log2sz) >> 1;
    for (int k = 0; k < 8; k++) sum += red[k];
    sum >>= log2sz;
    for (int i = tid; i < n; i += blockDim.x) ac[i] = (int16_t)(a[i] - sum);
}

// cfl_pred (src/ipred_tmpl.c:71-84 with the dc of the four variants)
template <typename BD>
__global__ void __launch_bounds__(256)
cfl_pred_kernel(int mode, uint8_t *dst8, int64_t stride, const typename BD::pixel *__restrict__ edge_in, int lo, int w, int h,
                const int16_t *__restrict__ ac, int alpha, int bdmax) {
    using pixel = typename BD::pixel;
    __shared__ int dc_s;
    const pixel *tl = edge_in + lo;
    if (threadIdx.x == 0) {
        unsigned dc;
        if (mode == IP_DC_128) {
            dc = BD::hbd ? (unsigned)(bdmax + 1) >> 1 : 128;
        } else if (mode == IP_TOP_DC) {
            dc = w >> 1;
            for (int i = 0; i < w; i++) dc += tl[1 + i];
            dc >>= ulog2(w);
        } else if (mode == IP_LEFT_DC) {
            dc = h >> 1;
            for (int i = 0; i < h; i++) dc += tl[-(1 + i)];
            dc >>= ulog2(h);
        } else {
            dc = (w + h) >> 1;
            for (int i = 0; i < w; i++) dc += tl[1 + i];
            for (int i = 0; i < h; i++) dc += tl[-(1 + i)];
            dc >>= __ffs(w + h) - 1;
            if (w != h) {
                const bool x4 = w > h * 2 || h > w * 2;
                dc *= BD::hbd ? (x4 ? 0x6667u : 0xAAABu) : (x4 ? 0x3334u : 0x5556u);
                dc >>= BD::hbd ? 17 : 16;
            }
        }
        dc_s = (int)dc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < w * h; i += blockDim.x) {
        const int diff = alpha * ac[i];
        const int m = (abs(diff) + 32) >> 6;
        ((pixel *)(dst8 + (int64_t)(i / w) * stride))[i % w] = (pixel)iclip(dc_s + (diff < 0 ? -m : m), 0, bdmax);
    }
}

template <typename BD>
__global__ void __launch_bounds__(256)
pal_pred_kernel(uint8_t *dst8, int64_t stride, const typename BD::pixel *__restrict__ pal, const uint8_t *__restrict__ idx, int w,
                int h) {
    using pixel = typename BD::pixel;
    for (int i = threadIdx.x + blockIdx.x * blockDim.x; i < w * h; i += blockDim.x * gridDim.x)
        ((pixel *)(dst8 + (int64_t)(i / w) * stride))[i % w] = pal[idx[i] & 7];
}

}  // namespace rb200

using namespace rb200;

namespace {
inline size_t ip_px(int bdmax) { return bdmax > 255 ? 2 : 1; }
inline bool ip_size_ok(int v) { return v == 4 || v == 8 || v == 16 || v == 32 || v == 64; }
}  // namespace

// ---------------------------------------------------------------- C ABI (per call)
extern "C" int rb200_ipred(int mode, void *dst, ptrdiff_t stride, const void *topleft, int w, int h, int angle, int max_w,
                           int max_h, int bdmax) {
    if (mode < 0 || mode >= IP_N_MODES || !dst || !topleft || !ip_size_ok(w) || !ip_size_ok(h) ||
        (mode == IP_FILTER && (w > 32 || h > 32 || (angle & 511) > 4)))
        return set_error(-22, "ipred: bad argument");
    const int a = angle & 511;
    if ((mode == IP_Z1 && !(a > 0 && a < 90)) || (mode == IP_Z2 && !(a > 90 && a < 180)) || (mode == IP_Z3 && !(a > 180 && a < 270)))
        return set_error(-22, "ipred: angle %d outside the range of the directional mode", a);
    const size_t px = ip_px(bdmax);
    const int lo = h + (w < h ? w : h), hi = w + (w < h ? w : h);   // everything the reference's predictors may read
    HostCall hc(2 * DevRect::bytes_for(w * px, h) + (size_t)(lo + hi + 1) * px * 2 + 256);
    DevRect drect;
    if (hc.rect_up(drect, dst, stride, w * px, h)) return hc.err;
    const void *de = hc.up((const uint8_t *)topleft - (size_t)lo * px, (size_t)(lo + hi + 1) * px);
    if (hc.err) return hc.err;
    if (bdmax > 255) ipred_kernel<BD16><<<1, 256, 0, hc.stream()>>>(mode, drect.dptr, drect.dpitch, (const uint16_t *)de, lo, hi, w, h, angle, max_w, max_h, bdmax);
    else ipred_kernel<BD8><<<1, 256, 0, hc.stream()>>>(mode, drect.dptr, drect.dpitch, (const uint8_t *)de, lo, hi, w, h, angle, max_w, max_h, bdmax);
    hc.rect_down(drect);
    if (hc.sync()) return hc.err;
    drect.finish(dst);
    return 0;
}

extern "C" int rb200_cfl_ac(int ss, int16_t *ac, const void *ypx, ptrdiff_t stride, int w_pad, int h_pad, int cw, int ch,
                            int bdmax) {
    if (ss < 0 || ss > 2 || !ac || !ypx || cw < 4 || ch < 4 || cw > 32 || ch > 32 || (cw & (cw - 1)) || (ch & (ch - 1)) ||
        w_pad < 0 || h_pad < 0 || w_pad * 4 >= cw || h_pad * 4 >= ch)
        return set_error(-22, "cfl_ac: bad argument");
    const int ss_hor = ss != 2, ss_ver = ss == 0;          // ss = layout - 1: 0 = 4:2:0, 1 = 4:2:2, 2 = 4:4:4
    const size_t px = ip_px(bdmax);
    const int lw = (cw - 4 * w_pad) << ss_hor, lh = (ch - 4 * h_pad) << ss_ver;
    HostCall hc(2 * DevRect::bytes_for(lw * px, lh) + (size_t)cw * ch * 4 + 256);
    DevRect yrect;
    if (hc.rect_up(yrect, ypx, stride, lw * px, lh)) return hc.err;
    int16_t *dac = (int16_t *)hc.dev((size_t)cw * ch * 2);
    if (hc.err) return hc.err;
    if (bdmax > 255) cfl_ac_kernel<BD16><<<1, 256, 0, hc.stream()>>>(dac, yrect.dptr, yrect.dpitch, w_pad, h_pad, cw, ch, ss_hor, ss_ver);
    else cfl_ac_kernel<BD8><<<1, 256, 0, hc.stream()>>>(dac, yrect.dptr, yrect.dpitch, w_pad, h_pad, cw, ch, ss_hor, ss_ver);
    void *s = hc.down(dac, (size_t)cw * ch * 2);
    if (hc.sync()) return hc.err;
    memcpy(ac, s, (size_t)cw * ch * 2);
    return 0;
}

extern "C" int rb200_cfl_pred(int mode, void *dst, ptrdiff_t stride, const void *topleft, int w, int h, const int16_t *ac,
                              int alpha, int bdmax) {
    if ((mode != IP_DC && mode != IP_LEFT_DC && mode != IP_TOP_DC && mode != IP_DC_128) || !dst || !topleft || !ac ||
        !ip_size_ok(w) || !ip_size_ok(h) || w > 32 || h > 32)
        return set_error(-22, "cfl_pred: bad argument");
    const size_t px = ip_px(bdmax);
    HostCall hc(2 * DevRect::bytes_for(w * px, h) + (size_t)(w + h + 1) * px * 2 + (size_t)w * h * 4 + 256);
    DevRect drect;
    if (hc.rect_up(drect, dst, stride, w * px, h)) return hc.err;
    const void *de = hc.up((const uint8_t *)topleft - (size_t)h * px, (size_t)(w + h + 1) * px);
    const int16_t *dac = (const int16_t *)hc.up(ac, (size_t)w * h * 2);
    if (hc.err) return hc.err;
    if (bdmax > 255) cfl_pred_kernel<BD16><<<1, 256, 0, hc.stream()>>>(mode, drect.dptr, drect.dpitch, (const uint16_t *)de, h, w, h, dac, alpha, bdmax);
    else cfl_pred_kernel<BD8><<<1, 256, 0, hc.stream()>>>(mode, drect.dptr, drect.dpitch, (const uint8_t *)de, h, w, h, dac, alpha, bdmax);
    hc.rect_down(drect);
    if (hc.sync()) return hc.err;
    drect.finish(dst);
    return 0;
}

extern "C" int rb200_pal_pred(void *dst, ptrdiff_t stride, const void *pal, const uint8_t *idx, int w, int h, int bdmax) {
    if (!dst || !pal || !idx || w < 4 || h < 4 || w > 64 || h > 64) return set_error(-22, "pal_pred: bad argument");
    const size_t px = ip_px(bdmax);
    HostCall hc(2 * DevRect::bytes_for(w * px, h) + (size_t)w * h * 2 + 64 + 256);
    DevRect drect;
    if (hc.rect_up(drect, dst, stride, w * px, h)) return hc.err;
    const void *dp = hc.up(pal, 8 * px);
    const uint8_t *di = (const uint8_t *)hc.up(idx, (size_t)w * h);
    if (hc.err) return hc.err;
    if (bdmax > 255) pal_pred_kernel<BD16><<<(w * h + 255) / 256, 256, 0, hc.stream()>>>(drect.dptr, drect.dpitch, (const uint16_t *)dp, di, w, h);
    else pal_pred_kernel<BD8><<<(w * h + 255) / 256, 256, 0, hc.stream()>>>(drect.dptr, drect.dpitch, (const uint8_t *)dp, di, w, h);
    hc.rect_down(drect);
    if (hc.sync()) return hc.err;
    drect.finish(dst);
    return 0;
}

// ---- function-pointer table (drop-in for rav1d_intra_pred_dsp_init, src/ipred.rs:1502-1560) ----
namespace {
#define IP_FATAL_IF(x, name) do { if (x) rb200_report_fatal(name); } while (0)
template <int MODE> void ipred_slot(void *d, ptrdiff_t s, const void *tl, int w, int h, int a, int mw, int mh, int bd) {
    IP_FATAL_IF(rb200_ipred(MODE, d, s, tl, w, h, a, mw, mh, bd), "intra_pred");
}
template <int SS, int BDMAX> void cfl_ac_slot(int16_t *ac, const void *y, ptrdiff_t s, int wp, int hp, int cw, int ch) {
    IP_FATAL_IF(rb200_cfl_ac(SS, ac, y, s, wp, hp, cw, ch, BDMAX), "cfl_ac");
}
template <int MODE> void cfl_pred_slot(void *d, ptrdiff_t s, const void *tl, int w, int h, const int16_t *ac, int alpha, int bd) {
    IP_FATAL_IF(rb200_cfl_pred(MODE, d, s, tl, w, h, ac, alpha, bd), "cfl_pred");
}
template <int BDMAX> void pal_pred_slot(void *d, ptrdiff_t s, const void *pal, const uint8_t *idx, int w, int h) {
    IP_FATAL_IF(rb200_pal_pred(d, s, pal, idx, w, h, BDMAX), "pal_pred");
}
template <int... M>
void fill_ipred(Rb200IntraPredDSPContext *c, std::integer_sequence<int, M...>) { ((c->intra_pred[M] = &ipred_slot<M>), ...); }
}  // namespace

extern "C" void rb200_intra_pred_dsp_init(Rb200IntraPredDSPContext *c, int bpc) {
    memset(c, 0, sizeof(*c));
    fill_ipred(c, std::make_integer_sequence<int, IP_N_MODES>{});
    if (bpc > 8) {   // cfl_ac and pal_pred carry no bitdepth_max: 16-bit pixels for 10 / 12 bpc
        c->cfl_ac[0] = &cfl_ac_slot<0, 1023>; c->cfl_ac[1] = &cfl_ac_slot<1, 1023>; c->cfl_ac[2] = &cfl_ac_slot<2, 1023>;
        c->pal_pred = &pal_pred_slot<1023>;
    } else {
        c->cfl_ac[0] = &cfl_ac_slot<0, 255>; c->cfl_ac[1] = &cfl_ac_slot<1, 255>; c->cfl_ac[2] = &cfl_ac_slot<2, 255>;
        c->pal_pred = &pal_pred_slot<255>;
    }
    c->cfl_pred[IP_DC] = &cfl_pred_slot<IP_DC>; c->cfl_pred[IP_LEFT_DC] = &cfl_pred_slot<IP_LEFT_DC>;
    c->cfl_pred[IP_TOP_DC] = &cfl_pred_slot<IP_TOP_DC>; c->cfl_pred[IP_DC_128] = &cfl_pred_slot<IP_DC_128>;
}
