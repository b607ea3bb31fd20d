// Deblocking loop filter for sm_100a.
//
// Replaces Rav1dLoopFilterDSPContext.loop_filter_sb[2][2] (src/loopfilter.rs:20-34; core `loop_filter`
// src/loopfilter.rs:397, sb drivers :500-1079) and, at frame level, the per-sbrow drivers
// rav1d_loopfilter_sbrow_cols / _rows (src/lf_apply.rs:597,763).
//
// The edge filter is written from the AV1 specification (7.14.6), not from the reference's expanded sums:
//   * the filter mask and the flatness masks are "largest difference <= limit" tests (one max chain each);
//   * the narrow filter (7.14.6.3) is the spec's single formula with the high-edge-variance term selected;
//   * the wide filters (7.14.6.4) are a sliding window -- out[i] = Round2(W[i] + centre taps, log2) with
//     W[i] = sum_{j=-n..n} x[clip(i + j)], and W[i + 1] = W[i] + x[clip(i + n + 1)] - x[clip(i - n)] -- two adds per
//     output instead of one add per tap (13 / 7 / 5 taps).
// Mapping at frame level: within a pass every edge is independent of every other (the filter length is bounded by the
// transform size on both sides), so a pass is one launch over a plane set: all column edges, then all row
// edges.  A CTA takes a 128x128 area, turns the set bits of its Av1Filter mask words into a compact shared-memory list
// of edge units (pop counts + one prefix sum) and filters list entries with full warps (see deblock_units_kernel).
#include "common.cuh"
#include "stages.cuh"
#include <stdlib.h>

namespace rb200 {

__device__ __forceinline__ int lf_absdiff(int a, int b) { return __sad(a, b, 0); }

// Wide filter of 2 * N outputs around the edge: x[8 + i], i = -N .. N - 1; samples are clipped to [-(N + 1), N].
// N2 = 1: the taps at i - 1, i, i + 1 count twice (13-tap luma, 5-tap chroma); N2 = 0: only the centre tap (7-tap luma).
template <int N, int N2, int LOG2>
__device__ __forceinline__ void lf_smooth(int (&x)[16]) {
    constexpr int LO = 8 - (N + 1), HI = 8 + N;
    int out[2 * N];
    int win = 0;
#pragma unroll
    for (int j = -N; j <= N; j++) { const int k = 8 - N + j; win += x[k < LO ? LO : (k > HI ? HI : k)]; }
#pragma unroll
    for (int i = -N; i < N; i++) {
        const int c = 8 + i;
        int centre = x[c];
        if (N2) centre += x[c - 1 < LO ? LO : c - 1] + x[c + 1 > HI ? HI : c + 1];
        out[i + N] = (win + centre + (1 << (LOG2 - 1))) >> LOG2;
        const int in = c + N + 1, gone = c - N;
        win += x[in > HI ? HI : in] - x[gone < LO ? LO : gone];
    }
#pragma unroll
    for (int i = 0; i < 2 * N; i++) x[8 - N + i] = out[i];
}

// One line across an edge, in registers: x[0..15] = p7 .. p0 | q0 .. q7 (x[7] = p0, x[8] = q0); only the samples the
// filter width needs have to be valid.  E, I, H are already scaled by << (bitdepth - 8).  Returns how many samples on
// each side may have changed (0, 1, 2, 3 or 6).  WD: 4 (both planes), 6 (chroma), 8 / 16 (luma).
template <int WD>
__device__ __forceinline__ int lf_edge(int (&x)[16], const int E, const int I, const int H, const int bdmin8, const int bdmax) {
    const int p0 = x[7], q0 = x[8], p1 = x[6], q1 = x[9];
    // filter mask (7.14.6.2): no step between neighbouring samples larger than I, and the step across the edge within E
    const int step01 = max(lf_absdiff(p1, p0), lf_absdiff(q1, q0));
    int step = step01;
    if (WD > 4) step = max(step, max(lf_absdiff(x[5], p1), lf_absdiff(x[10], q1)));
    if (WD > 6) step = max(step, max(lf_absdiff(x[4], x[5]), lf_absdiff(x[11], x[10])));
    if (step > I || lf_absdiff(p0, q0) * 2 + (lf_absdiff(p1, q1) >> 1) > E) return 0;
    // flatness: every sample of a side within 1 << (bitdepth - 8) of the sample next to the edge
    if (WD >= 6) {
        const int F = 1 << bdmin8;
        int dev = max(step01, max(lf_absdiff(x[5], p0), lf_absdiff(x[10], q0)));
        if (WD >= 8) dev = max(dev, max(lf_absdiff(x[4], p0), lf_absdiff(x[11], q0)));
        if (dev <= F) {
            if (WD == 16) {
                int far = max(max(lf_absdiff(x[3], p0), lf_absdiff(x[2], p0)), lf_absdiff(x[1], p0));
                far = max(far, max(max(lf_absdiff(x[12], q0), lf_absdiff(x[13], q0)), lf_absdiff(x[14], q0)));
                if (far <= F) { lf_smooth<6, 1, 4>(x); return 6; }
            }
            if (WD >= 8) { lf_smooth<3, 0, 3>(x); return 3; }
            lf_smooth<2, 1, 3>(x);
            return 2;
        }
    }
    // narrow filter (7.14.6.3)
    const int lo = -(128 << bdmin8), hi = (128 << bdmin8) - 1;
    const bool hev = step01 > H;
    int f = hev ? iclip(p1 - q1, lo, hi) : 0;
    f = iclip(f + 3 * (q0 - p0), lo, hi);
    const int f1 = imin(f + 4, hi) >> 3, f2 = imin(f + 3, hi) >> 3;
    x[8] = iclip(q0 - f1, 0, bdmax);
    x[7] = iclip(p0 + f2, 0, bdmax);
    if (hev) return 1;
    const int f3 = (f1 + 1) >> 1;
    x[9] = iclip(q1 - f3, 0, bdmax);
    x[6] = iclip(p1 + f3, 0, bdmax);
    return 2;
}

__device__ __forceinline__ int lf_edge_wd(int wd, int (&x)[16], int E, int I, int H, int bdmin8, int bdmax) {
    switch (wd) {
    case 16: return lf_edge<16>(x, E, I, H, bdmin8, bdmax);
    case 8: return lf_edge<8>(x, E, I, H, bdmin8, bdmax);
    case 6: return lf_edge<6>(x, E, I, H, bdmin8, bdmax);
    default: return lf_edge<4>(x, E, I, H, bdmin8, bdmax);
    }
}

struct LfPlane {
    uint8_t *base;
    int64_t stride;
    int w4, h4;         // plane size in 4-pixel units that are filtered (luma: f.w4 / f.h4; chroma: rounded-up halves)
    int ss_hor, ss_ver; // of this plane
    int lvl_idx;        // byte of level[][4]: luma col 0, luma row 1, u 2, v 3
    int uv;             // 0 luma, 1 chroma
    int y4_begin, y4_end;   // unit rows processed in this launch (band restriction)
};
struct LfPlaneSet {
    LfPlane p[3];
    int n_planes;
    int sb128w, b4_stride;
    int sby_first;      // first 128-row superblock row of the launch
};

constexpr int LF_THREADS = 256;
constexpr int LF_LIST = 3 * 32 * 32;   // edge units of one 128x128 area: luma + two chroma planes (4:4:4 at most)

// Frame pass, one direction: a CTA takes one 128x128 luma area (and the co-located chroma).
//  1. Edge list.  The Av1Filter words ARE the list of edges: filter_y[dir][a][idx][half] is a 16-bit word whose bits run
//     along the edge direction.  A thread takes one (plane, a, half), ORs the words of the filter widths, masks what lies
//     outside the picture / the band / on the picture border, and the CTA turns the set bits into a compact list of
//     {x4, y4, width index, plane} (one block-wide prefix sum over the pop counts) -- no per-unit test, no division.
//  2. Filtering.  Column edges (DIR 0): four lanes per list entry, one LINE each -- the line's 8 / 16 / 32 pixels
//     straddling the edge as aligned 4-pixel words, only the words the filter may have modified written back.
//     Row edges (DIR 1): one lane per entry; entries of a mask word are horizontally adjacent units, so every stencil
//     row is one coalesced run of 4-pixel words.
template <typename BD, int DIR>
__global__ void __launch_bounds__(LF_THREADS, DIR ? 3 : 4)
deblock_units_kernel(LfPlaneSet S, const Rb200Av1Filter *__restrict__ masks, const uint8_t (*__restrict__ lvl)[4],
                     const Rb200Av1FilterLUT *__restrict__ lut, int bdmax) {
    using pixel = typename BD::pixel;
    __shared__ uint32_t list[LF_LIST];            // x4 | y4 << 12 | idx << 24 | plane << 26
    __shared__ int warp_tot[LF_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sbx = blockIdx.x, sby = S.sby_first + blockIdx.y;
    const Rb200Av1Filter &M = masks[sby * S.sb128w + sbx];
    // ---- 1. edge list
    // task t: plane pi, position a across the edges, half
    int n_task[3], t0 = 0, my_pi = -1, my_t = 0;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        n_task[i] = 0;
        if (i < S.n_planes) {
            const LfPlane &P = S.p[i];
            n_task[i] = 2 * (32 >> (DIR == 0 ? P.ss_hor : P.ss_ver));
            if (my_pi < 0 && tid < t0 + n_task[i]) { my_pi = i; my_t = tid - t0; }
            t0 += n_task[i];
        }
    }
    unsigned bits = 0, w1 = 0, w2 = 0;
    int ex4 = 0, ey4 = 0;               // unit of bit 0 of this thread's word; bits advance along y (DIR 0) or x (DIR 1)
    if (my_pi >= 0) {
        const LfPlane &P = my_pi == 0 ? S.p[0] : (my_pi == 1 ? S.p[1] : S.p[2]);
        const int a = my_t >> 1, half = my_t & 1;
        const int nb = 16 >> (DIR == 0 ? P.ss_ver : P.ss_hor);      // bits per half-word
        unsigned w0;
        if (!P.uv) { w0 = M.filter_y[DIR][a][0][half]; w1 = M.filter_y[DIR][a][1][half]; w2 = M.filter_y[DIR][a][2][half]; }
        else { w0 = M.filter_uv[DIR][a][0][half]; w1 = M.filter_uv[DIR][a][1][half]; w2 = 0; }
        bits = (w0 | w1 | w2) & ((1u << nb) - 1);
        const int x4s = sbx << (5 - P.ss_hor), y4s = sby << (5 - P.ss_ver);      // first unit of the area in this plane
        if (DIR == 0) { ex4 = x4s + a; ey4 = y4s + half * nb; }
        else { ex4 = x4s + half * nb; ey4 = y4s + a; }
        // what lies outside the picture or the rows of this launch, and the picture's own border, carries no edge
        const int along0 = DIR == 0 ? ey4 : ex4, along_lo = DIR == 0 ? P.y4_begin : 0, along_hi = DIR == 0 ? P.y4_end : P.w4;
        const int lo = imax(along_lo - along0, 0), hi = imin(along_hi - along0, nb);
        bits = hi > lo ? bits & ((1u << hi) - (1u << lo)) : 0u;
        const int across = DIR == 0 ? ex4 : ey4;
        if (across == 0 || across >= (DIR == 0 ? P.w4 : P.h4) || (DIR == 1 && (across < P.y4_begin || across >= P.y4_end))) bits = 0;
    }
    // block-wide exclusive prefix sum of the pop counts
    const int cnt = __popc(bits);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int off = incl - cnt, n = 0;
#pragma unroll
    for (int w = 0; w < LF_THREADS / 32; w++) { const int v = warp_tot[w]; if (w < warp) off += v; n += v; }
    while (bits) {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1;
        const unsigned idx = ((w2 >> b) & 1) ? 2u : ((w1 >> b) & 1);
        list[off++] = (unsigned)(DIR == 0 ? ex4 : ex4 + b) | (unsigned)(DIR == 0 ? ey4 + b : ey4) << 12 | idx << 24 | (unsigned)my_pi << 26;
    }
    __syncthreads();
    // ---- 2. filter the list
    const int bdmin8 = BD::hbd ? bpc_from_max(bdmax) - 8 : 0;
    constexpr int WPR = BD::hbd ? 2 : 1;            // 32-bit words per 4-pixel group
    auto unpack4 = [](const unsigned *w, int *v) {
        if (BD::hbd) { v[0] = w[0] & 0xffff; v[1] = w[0] >> 16; v[2] = w[WPR - 1] & 0xffff; v[3] = w[WPR - 1] >> 16; }
        else { v[0] = w[0] & 0xff; v[1] = (w[0] >> 8) & 0xff; v[2] = (w[0] >> 16) & 0xff; v[3] = w[0] >> 24; }
    };
    auto pack4 = [](const int *v, unsigned *w) {
        if (BD::hbd) { w[0] = (unsigned)v[0] | (unsigned)v[1] << 16; w[WPR - 1] = (unsigned)v[2] | (unsigned)v[3] << 16; }
        else w[0] = (unsigned)v[0] | (unsigned)v[1] << 8 | (unsigned)v[2] << 16 | (unsigned)v[3] << 24;
    };
    auto ld4 = [](const pixel *p, unsigned *w) {
        if (BD::hbd) { const uint2 q = *(const uint2 *)p; w[0] = q.x; w[WPR - 1] = q.y; } else w[0] = *(const unsigned *)p;
    };
    auto st4 = [](pixel *p, const unsigned *w) {
        if (BD::hbd) *(uint2 *)p = make_uint2(w[0], w[WPR - 1]); else *(unsigned *)p = w[0];
    };
    constexpr int LANES_PER_ENTRY = DIR == 0 ? 4 : 1;
#pragma unroll 1
    for (int e = tid; e < n * LANES_PER_ENTRY; e += LF_THREADS) {
        const uint32_t entry = list[DIR == 0 ? e >> 2 : e];
        const int x4 = entry & 0xfff, y4 = (entry >> 12) & 0xfff, idx = (entry >> 24) & 3, pi = entry >> 26;
        const LfPlane &P = pi == 0 ? S.p[0] : (pi == 1 ? S.p[1] : S.p[2]);
        const uint8_t(*l)[4] = lvl + (int64_t)y4 * S.b4_stride + x4;
        int L = l[0][P.lvl_idx];
        if (!L) L = DIR == 0 ? l[-1][P.lvl_idx] : l[-(int64_t)S.b4_stride][P.lvl_idx];
        if (!L) continue;
        const int64_t ps = P.stride / (int64_t)sizeof(pixel);
        pixel *base = (pixel *)P.base + (int64_t)(y4 * 4) * ps + x4 * 4;   // q0 of line 0
        const int H = (L >> 4) << bdmin8, E = (int)lut->e[L] << bdmin8, I = (int)lut->i[L] << bdmin8;
        const int wd = P.uv ? 4 + 2 * idx : 4 << idx;
        const int ng = wd == 16 ? 2 : 1;           // 4-pixel groups on each side of the edge
        if (DIR == 0) {
            pixel *p = base + (int64_t)(e & 3) * ps;
            unsigned w[4][WPR];
            int x[16];
            ld4(p - 4, w[1]); ld4(p, w[2]);
            if (ng == 2) { ld4(p - 8, w[0]); ld4(p + 4, w[3]); unpack4(w[0], x); unpack4(w[3], x + 12); }
            unpack4(w[1], x + 4); unpack4(w[2], x + 8);
            const int m = lf_edge_wd(wd, x, E, I, H, bdmin8, bdmax);
            if (!m) continue;
            if (wd == 4) {
                // the neighbouring edges may be only 4 pixels away and own p3 / p2 and q2 / q3: write p1 p0 | q0 q1 only
                if (BD::hbd) { *(unsigned *)(p - 2) = (unsigned)x[6] | (unsigned)x[7] << 16; *(unsigned *)p = (unsigned)x[8] | (unsigned)x[9] << 16; }
                else { *(uint16_t *)(p - 2) = (uint16_t)(x[6] | x[7] << 8); *(uint16_t *)p = (uint16_t)(x[8] | x[9] << 8); }
                continue;
            }
            // wider filters: both transform blocks are at least 8 (16) pixels, so the 4-pixel words are this edge's alone
            pack4(x + 4, w[1]); pack4(x + 8, w[2]); st4(p - 4, w[1]); st4(p, w[2]);
            if (m > 4) { pack4(x, w[0]); pack4(x + 12, w[3]); st4(p - 8, w[0]); st4(p + 4, w[3]); }
        } else {
            // rows y - 4 ng .. y + 4 ng - 1 of the unit's 4 columns, one packed word group per row
            unsigned rows[16][WPR];
#pragma unroll
            for (int r = 4; r < 12; r++) ld4(base + (int64_t)(r - 8) * ps, rows[r]);
            if (ng == 2) {
#pragma unroll
                for (int r = 0; r < 16; r++) if (r < 4 || r >= 12) ld4(base + (int64_t)(r - 8) * ps, rows[r]);
            }
            auto get = [&](int r, int c) -> int {
                if (BD::hbd) return (int)((rows[r][c >> 1] >> (16 * (c & 1))) & 0xffff);
                return (int)((rows[r][0] >> (8 * c)) & 0xff);
            };
            auto put = [&](int r, int c, int v) {
                if (BD::hbd) rows[r][c >> 1] = (rows[r][c >> 1] & ~(0xffffu << (16 * (c & 1)))) | ((unsigned)v << (16 * (c & 1)));
                else rows[r][0] = (rows[r][0] & ~(0xffu << (8 * c))) | ((unsigned)v << (8 * c));
            };
            int mmax = 0;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                int x[16];
#pragma unroll
                for (int r = 0; r < 16; r++) x[r] = get(r, c);
                const int m = lf_edge_wd(wd, x, E, I, H, bdmin8, bdmax);
                mmax = imax(mmax, m);
#pragma unroll
                for (int i = 0; i < 6; i++)
                    if (i < m) { put(7 - i, c, x[7 - i]); put(8 + i, c, x[8 + i]); }
            }
#pragma unroll
            for (int i = 0; i < 6; i++)
                if (i < mmax) { st4(base + (int64_t)(-1 - i) * ps, rows[7 - i]); st4(base + (int64_t)i * ps, rows[8 + i]); }
        }
    }
}

// Per-call form of loop_filter_{h,v}_sb128{y,uv}: explicit mask words and level pointer.
// thread = (unit along the edge 0..31, line 0..3)
template <typename BD>
__global__ void lpf_sb_kernel(uint8_t *dst, int64_t stride, int uv, int dir, uint32_t m0, uint32_t m1, uint32_t m2,
                              const uint8_t *__restrict__ lvl /* row 0 = l[-1] resp. l[-b4_stride] */, int lvl_pitch,
                              Rb200Av1FilterLUT lut, int bdmax) {
    using pixel = typename BD::pixel;
    const int u = threadIdx.x >> 2, line = threadIdx.x & 3;
    const uint32_t bit = 1u << u;
    if (!((m0 | m1 | m2) & bit)) return;
    // staged levels: [unit][2] = {neighbour (l[-1] / l[-b4_stride]), own}, one byte each
    int L = lvl[u * lvl_pitch + 1];
    if (!L) L = lvl[u * lvl_pitch + 0];
    if (!L) return;
    const int idx = uv ? ((m1 & bit) ? 1 : 0) : ((m2 & bit) ? 2 : ((m1 & bit) ? 1 : 0));
    const int wd = uv ? 4 + 2 * idx : 4 << idx;
    const int bdmin8 = BD::hbd ? bpc_from_max(bdmax) - 8 : 0;
    const int H = (L >> 4) << bdmin8, E = (int)lut.e[L] << bdmin8, I = (int)lut.i[L] << bdmin8;
    const int64_t ps = stride / (int64_t)sizeof(pixel);
    pixel *q0 = dir == 0 ? (pixel *)dst + (int64_t)(u * 4 + line) * ps : (pixel *)dst + u * 4 + line;
    const int64_t sb = dir == 0 ? 1 : ps;           // element stride across the edge
    const int reach = wd == 16 ? 7 : (wd >> 1);     // samples read on each side
    int x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = (i >= 8 - reach && i < 8 + reach) ? (int)q0[(int64_t)(i - 8) * sb] : 0;
    const int m = lf_edge_wd(wd, x, E, I, H, bdmin8, bdmax);
#pragma unroll
    for (int i = 0; i < 6; i++)
        if (i < m) { q0[(int64_t)(-1 - i) * sb] = (pixel)x[7 - i]; q0[(int64_t)i * sb] = (pixel)x[8 + i]; }
}

// Whole-frame deblock: column edges of every plane (one launch), then row edges (one launch).
// y4b / y4e: luma 4-pixel unit rows whose ROW edges are filtered (the whole picture: 0, h4); the
// column-edge pass covers two more unit rows on each side, which is everything those row edges read.
int deblock_frame_launch(const Rb200Planes &pl, int n_planes, int w4, int h4, int sb128w, int b4_stride, int ss_hor,
                         int ss_ver, bool do_uv, const Rb200Av1Filter *masks, const uint8_t (*lvl)[4],
                         const Rb200Av1FilterLUT *lut, int bdmax, cudaStream_t st, int *launches, int y4b, int y4e) {
    return deblock_planes_launch(pl, n_planes, w4, h4, sb128w, b4_stride, ss_hor, ss_ver, do_uv ? 7 : 1, masks, lvl, lut, bdmax, st,
                                 launches, y4b, y4e);
}

int deblock_planes_launch(const Rb200Planes &pl, int n_planes, int w4, int h4, int sb128w, int b4_stride, int ss_hor,
                          int ss_ver, int plane_mask, const Rb200Av1Filter *masks, const uint8_t (*lvl)[4],
                          const Rb200Av1FilterLUT *lut, int bdmax, cudaStream_t st, int *launches, int y4b, int y4e) {
    static const int only_dir = getenv("RB200_LF_ONLY_DIR") ? atoi(getenv("RB200_LF_ONLY_DIR")) : -1;   // debugging aid
    for (int dir = 0; dir < 2; dir++) {
        if (only_dir >= 0 && dir != only_dir) continue;
        LfPlaneSet S = {};
        int n = 0, y_lo = 1 << 30, y_hi = 0;     // luma unit rows touched by this launch
        for (int p = 0; p < n_planes; p++) {
            if (!((plane_mask >> p) & 1)) continue;
            LfPlane &P = S.p[n];
            P.uv = p ? 1 : 0;
            P.ss_hor = p ? ss_hor : 0; P.ss_ver = p ? ss_ver : 0;
            P.w4 = (w4 + P.ss_hor) >> P.ss_hor; P.h4 = (h4 + P.ss_ver) >> P.ss_ver;
            P.lvl_idx = p ? 1 + p : dir;
            P.base = (uint8_t *)pl.data[p]; P.stride = pl.stride[p];
            // band in this plane's unit rows (chroma rows round outwards); the column-edge pass covers two more unit
            // rows on each side, which is everything the row edges of the band read
            int b = dir ? y4b : y4b - 2, e = dir ? y4e : y4e + 2;
            b = imax(b >> P.ss_ver, 0); e = imin((e + P.ss_ver) >> P.ss_ver, P.h4);
            P.y4_begin = b; P.y4_end = e;
            if (e > b) { y_lo = imin(y_lo, b << P.ss_ver); y_hi = imax(y_hi, e << P.ss_ver); }
            n++;
        }
        S.n_planes = n; S.sb128w = sb128w; S.b4_stride = b4_stride;
        if (!n || y_hi <= y_lo) continue;
        S.sby_first = y_lo >> 5;
        const dim3 grid(sb128w, ((y_hi + 31) >> 5) - S.sby_first);
#define L(BD, D) deblock_units_kernel<BD, D><<<grid, LF_THREADS, 0, st>>>(S, masks, lvl, lut, bdmax)
        if (bdmax > 255) { if (dir) L(BD16, 1); else L(BD16, 0); } else { if (dir) L(BD8, 1); else L(BD8, 0); }
#undef L
        RB_LAUNCH_CHECK();
        if (launches) ++*launches;
        static const int sync_between = getenv("RB200_LF_SYNC_BETWEEN") ? atoi(getenv("RB200_LF_SYNC_BETWEEN")) : 0;   // debugging aid
        if (sync_between) RB_CUDA(cudaStreamSynchronize(st));
    }
    return 0;
}

}  // namespace rb200

using namespace rb200;

extern "C" int rb200_loop_filter_sb(int uv, int dir, void *dst, ptrdiff_t stride, const uint32_t *mask,
                                    const uint8_t (*lvl)[4], ptrdiff_t lvl_stride, const Rb200Av1FilterLUT *lut,
                                    int w_or_h, int bdmax) {
    (void)w_or_h;  // unused by the reference's C/Rust bodies too (src/loopfilter_tmpl.c:163-240)
    if (!dst || !mask || !lvl || !lut || uv < 0 || uv > 1 || dir < 0 || dir > 1)
        return set_error(-22, "loop_filter_sb: bad argument");
    const uint32_t m0 = mask[0], m1 = mask[1], m2 = uv ? 0 : mask[2];
    const uint32_t vm = m0 | m1 | m2;
    if (!vm) return 0;
    const int units = 32 - __builtin_clz(vm);  // the reference stops at the highest set bit
    const size_t px = bdmax > 255 ? 2 : 1;
    // pixels touched: up to 8 before and 8 after the edge across it, units*4 along it
    const int across = 16, along = units * 4;
    const int cols = dir == 0 ? across : along, rows = dir == 0 ? along : across;
    uint8_t *row0 = (uint8_t *)dst - (dir == 0 ? (int64_t)8 * (int64_t)px : (int64_t)8 * stride);
    HostCall hc(2 * DevRect::bytes_for(cols * px, rows) + 4096);
    DevRect rect;
    if (hc.rect_up(rect, row0, stride, cols * px, rows)) return hc.err;
    // levels: {neighbour, own} per unit; h filter walks l += b4_stride per unit with neighbour l[-1];
    // v filter walks l++ per unit with neighbour l[-b4_stride]
    uint8_t hl[64];
    for (int u = 0; u < units; u++) {
        const uint8_t(*l)[4] = dir == 0 ? lvl + (int64_t)u * lvl_stride : lvl + u;
        // the reference only dereferences the neighbour when the own level is 0 and the unit is selected
        const bool sel = (vm >> u) & 1;
        hl[2 * u + 1] = sel ? l[0][0] : 0;
        hl[2 * u] = (sel && !l[0][0]) ? (dir == 0 ? l[-1][0] : l[-lvl_stride][0]) : 0;
    }
    const uint8_t *dl = (const uint8_t *)hc.up(hl, 2 * units);
    if (hc.err) return hc.err;
    uint8_t *d = rect.dptr + (dir == 0 ? (int64_t)8 * (int64_t)px : (int64_t)8 * rect.dpitch);
    if (bdmax > 255) lpf_sb_kernel<BD16><<<1, units * 4, 0, hc.stream()>>>(d, rect.dpitch, uv, dir, m0, m1, m2, dl, 2, *lut, bdmax);
    else lpf_sb_kernel<BD8><<<1, units * 4, 0, hc.stream()>>>(d, rect.dpitch, uv, dir, m0, m1, m2, dl, 2, *lut, bdmax);
    hc.rect_down(rect);
    if (hc.sync()) return hc.err;
    rect.finish(row0);
    return 0;
}

namespace {
template <int UV, int DIR>
void lpf_slot(void *dst, ptrdiff_t stride, const uint32_t *mask, const uint8_t (*lvl)[4], ptrdiff_t ls,
              const Rb200Av1FilterLUT *lut, int wh, int bd) {
    if (rb200_loop_filter_sb(UV, DIR, dst, stride, mask, lvl, ls, lut, wh, bd)) rb200_report_fatal("loop_filter_sb");
}
}  // namespace

extern "C" void rb200_loop_filter_dsp_init(Rb200LoopFilterDSPContext *c, int bpc) {
    (void)bpc;
    c->loop_filter_sb[0][0] = &lpf_slot<0, 0>;
    c->loop_filter_sb[0][1] = &lpf_slot<0, 1>;
    c->loop_filter_sb[1][0] = &lpf_slot<1, 0>;
    c->loop_filter_sb[1][1] = &lpf_slot<1, 1>;
}
