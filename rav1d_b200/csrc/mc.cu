// Motion compensation for sm_100a: put / prep with 8-tap and bilinear filters,
// scaled variants, compound combines (avg, w_avg, mask, w_mask), OBMC /
// inter-intra blends, 8x8 affine warp, edge emulation and super-res resize.
//
// Replaces Rav1dMCDSPContext (src/mc.rs:1321-1338; kernels src/mc.rs:28-1172 ==
// src/mc_tmpl.c) and the addressing half of recon.rs `mc()` (src/recon.rs:2025-2203).
//
// Mapping: one warp owns one prediction tile of at most 16x16 pixels.  The
// source window (tile + 7 rows/columns of filter support) is staged into a
// warp-private shared-memory tile with coordinates clamped to the reference
// picture -- the clamp *is* emu_edge (src/mc.rs:1032-1112), so no edge buffer is
// ever materialised -- then filtered horizontally into an int16 intermediate
// (same truncating store as the reference's `mid`, src/mc.rs:161) and
// vertically into the destination.  Larger blocks are split into 16x16 tiles;
// the split is exact because filtering is separable per pixel and the 4-tap
// selection rule (w <= 4 / h <= 4, src/mc.rs:119-127) can only apply to blocks
// that are not split in that dimension.
#include "common.cuh"
#include "tables.cuh"
#include "wedge.cuh"
#include "tma.cuh"
#include "stages.cuh"
#include <mutex>
#include <utility>

namespace rb200 {

template <typename BD>
struct McBits {
    // intermediate_bits: 4 for 8 and 10 bpc, 2 for 12 bpc (include/common/bitdepth.rs:315,395)
    static __device__ __forceinline__ int ib(int bdmax) { return BD::hbd ? 14 - bpc_from_max(bdmax) : 4; }
    static constexpr int prep_bias = BD::hbd ? 8192 : 0;  // bitdepth.rs:320,401
};

constexpr int MC_TILE = 16;
constexpr int MC_WIN_ROWS = MC_TILE + 7;
constexpr int MC_WIN_PITCH = MC_TILE + 8;  // 24 columns: 23 used
constexpr int MC_WARPS = 4;

struct McSmem {
    uint16_t win[MC_WIN_ROWS * MC_WIN_PITCH];
    int16_t mid[MC_WIN_ROWS * MC_TILE];
};

struct McRef {
    const uint8_t *base;  // row 0 of the reference plane
    int64_t stride;       // bytes
    int w, h;             // clamp bounds in pixels
};

__device__ __forceinline__ void load_taps(int f[8], const int8_t *p) {
#pragma unroll
    for (int k = 0; k < 8; k++) f[k] = p[k];
}

// One warp: w x h tile (<= 16x16) whose top-left source sample is (sx, sy) in `ref`.
// bw/bh: size of the whole prediction block (selects 4-tap filters).
// PREP: write int16 `(value) - PREP_BIAS` to out16 (row pitch out_pitch elements);
// else write clipped pixels to out8/out16 as pixels (row pitch in bytes).
template <typename BD, bool PREP>
__device__ void mc_tile(McSmem &sm, const McRef ref, int sx, int sy, int w, int h, int bw, int bh,
                        int mx, int my, int filter2d, void *out, int64_t out_pitch, int bdmax) {
    using pixel = typename BD::pixel;
    const int lane = threadIdx.x & 31;
    const int ib = McBits<BD>::ib(bdmax);
    const bool bilin = filter2d == RB200_FILTER_2D_BILINEAR;
    const bool fh = mx != 0, fv = my != 0;
    const int x0 = fh ? (bilin ? 0 : -3) : 0, y0 = fv ? (bilin ? 0 : -3) : 0;
    const int ncols = w + (fh ? (bilin ? 1 : 7) : 0), nrows = h + (fv ? (bilin ? 1 : 7) : 0);

    // ---- stage the clamped source window
    for (int i = lane; i < nrows * ncols; i += 32) {
        const int r = i / ncols, c = i - r * ncols;
        const int yy = iclip(sy + y0 + r, 0, ref.h - 1), xx = iclip(sx + x0 + c, 0, ref.w - 1);
        sm.win[r * MC_WIN_PITCH + c] = ((const pixel *)(ref.base + (int64_t)yy * ref.stride))[xx];
    }
    __syncwarp();

    auto store = [&](int r, int c, int v) {  // v: final value (pixel or prep)
        if (PREP) ((int16_t *)out)[(int64_t)r * out_pitch + c] = (int16_t)v;
        else ((pixel *)((uint8_t *)out + (int64_t)r * out_pitch))[c] = (pixel)v;
    };

    if (bilin) {
        // put_bilin / prep_bilin, src/mc.rs:431-652 == src/mc_tmpl.c:354-540
        if (fh && fv) {
            const int sh1 = 4 - ib, r1 = (1 << sh1) >> 1;
            for (int i = lane; i < (h + 1) * w; i += 32) {
                const int r = i / w, c = i - r * w;
                const int a = sm.win[r * MC_WIN_PITCH + c], b = sm.win[r * MC_WIN_PITCH + c + 1];
                sm.mid[r * MC_TILE + c] = (int16_t)((16 * a + mx * (b - a) + r1) >> sh1);
            }
            __syncwarp();
            for (int i = lane; i < h * w; i += 32) {
                const int r = i / w, c = i - r * w;
                const int a = sm.mid[r * MC_TILE + c], b = sm.mid[(r + 1) * MC_TILE + c];
                const int s = 16 * a + my * (b - a);
                if (PREP) store(r, c, ((s + 8) >> 4) - McBits<BD>::prep_bias);
                else store(r, c, iclip((s + ((1 << (4 + ib)) >> 1)) >> (4 + ib), 0, bdmax));
            }
        } else if (fh || fv) {
            const int step = fh ? 1 : MC_WIN_PITCH, m = fh ? mx : my;
            for (int i = lane; i < h * w; i += 32) {
                const int r = i / w, c = i - r * w;
                const int a = sm.win[r * MC_WIN_PITCH + c], b = sm.win[r * MC_WIN_PITCH + c + step];
                const int s = 16 * a + m * (b - a);
                if (PREP) {
                    store(r, c, ((s + ((1 << (4 - ib)) >> 1)) >> (4 - ib)) - McBits<BD>::prep_bias);
                } else if (fh) {
                    const int px = (s + ((1 << (4 - ib)) >> 1)) >> (4 - ib);
                    store(r, c, iclip((px + ((1 << ib) >> 1)) >> ib, 0, bdmax));
                } else {
                    store(r, c, iclip((s + 8) >> 4, 0, bdmax));
                }
            }
        } else {
            for (int i = lane; i < h * w; i += 32) {
                const int r = i / w, c = i - r * w;
                const int a = sm.win[r * MC_WIN_PITCH + c];
                store(r, c, PREP ? (a << ib) - McBits<BD>::prep_bias : a);
            }
        }
        __syncwarp();
        return;
    }

    // 8-tap: filter_type = h | v << 2 with (0 regular, 1 smooth, 2 sharp)   (src/mc_tmpl.c:330-338)
    //   Filter2d order: REG, REG_SMOOTH, REG_SHARP, SHARP_REG, SHARP_SMOOTH, SHARP, SMOOTH_REG, SMOOTH, SMOOTH_SHARP
    //   name = <h>_<v>:  h type: 0,0,0,2,2,2,1,1,1   v type: 0,1,2,0,1,2,0,1,2
    const int th_ = (0x111222000LL >> (4 * filter2d)) & 3, tv_ = (0x210210210LL >> (4 * filter2d)) & 3;
    int FH[8], FV[8];
    if (fh) load_taps(FH, tab::subpel(bw > 4 ? th_ : 3 + (th_ & 1), mx - 1));
    if (fv) load_taps(FV, tab::subpel(bh > 4 ? tv_ : 3 + (tv_ & 1), my - 1));

    if (fh && fv) {
        const int sh1 = 6 - ib, r1 = (1 << sh1) >> 1;
        for (int i = lane; i < (h + 7) * w; i += 32) {
            const int r = i / w, c = i - r * w;
            const uint16_t *s = sm.win + r * MC_WIN_PITCH + c;
            int acc = r1;
#pragma unroll
            for (int k = 0; k < 8; k++) acc += FH[k] * (int)s[k];
            sm.mid[r * MC_TILE + c] = (int16_t)(acc >> sh1);
        }
        __syncwarp();
        for (int i = lane; i < h * w; i += 32) {
            const int r = i / w, c = i - r * w;
            const int16_t *s = sm.mid + r * MC_TILE + c;
            int acc = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) acc += FV[k] * (int)s[k * MC_TILE];
            if (PREP) store(r, c, ((acc + 32) >> 6) - McBits<BD>::prep_bias);
            else store(r, c, iclip((acc + ((1 << (6 + ib)) >> 1)) >> (6 + ib), 0, bdmax));
        }
    } else if (fh) {
        for (int i = lane; i < h * w; i += 32) {
            const int r = i / w, c = i - r * w;
            const uint16_t *s = sm.win + r * MC_WIN_PITCH + c;
            int acc = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) acc += FH[k] * (int)s[k];
            if (PREP) store(r, c, ((acc + ((1 << (6 - ib)) >> 1)) >> (6 - ib)) - McBits<BD>::prep_bias);
            else store(r, c, iclip((acc + 32 + ((1 << (6 - ib)) >> 1)) >> 6, 0, bdmax));
        }
    } else if (fv) {
        for (int i = lane; i < h * w; i += 32) {
            const int r = i / w, c = i - r * w;
            const uint16_t *s = sm.win + r * MC_WIN_PITCH + c;
            int acc = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) acc += FV[k] * (int)s[k * MC_WIN_PITCH];
            if (PREP) store(r, c, ((acc + ((1 << (6 - ib)) >> 1)) >> (6 - ib)) - McBits<BD>::prep_bias);
            else store(r, c, iclip((acc + 32) >> 6, 0, bdmax));
        }
    } else {
        for (int i = lane; i < h * w; i += 32) {
            const int r = i / w, c = i - r * w;
            const int a = sm.win[r * MC_WIN_PITCH + c];
            store(r, c, PREP ? (a << ib) - McBits<BD>::prep_bias : a);
        }
    }
    __syncwarp();
}

// ---------------------------------------------------------------- fast 8-tap tile (batch path)
// Same arithmetic as mc_tile, restated for the integer dot-product unit: the source window is
// staged as 32-bit words holding two horizontally adjacent pixels, the horizontal pass packs
// its int16 results as vertical pairs, and both passes run on IDP.2A (dp2a: two 16-bit x 8-bit
// multiply-adds per instruction).  An output whose first tap falls on the low half of a word
// needs 4 IDP, one that starts on the high half 5 (with the taps shifted by one byte), so a
// pair of outputs costs 9 instead of 16 IMADs.  Rows of the window are clamped to the
// reference (free); windows that leave it horizontally are staged with the clamped per-pixel
// gather (emu_edge, src/mc.rs:1032-1112) into the same layout.
constexpr int MCF_WROWS = 24;   // window rows (16 + 7, rounded to pairs)
constexpr int MCF_WPW = 20;     // words per window row = the TMA box width of 40 pixels (7 of alignment + 23 + spare); row
                                // pairs are 40 words apart, so the 4 row pairs x 8 column pairs a warp reads at once hit 32 banks
constexpr int MCF_TMA_W = 2 * MCF_WPW, MCF_TMA_H = MCF_WROWS;   // box of the reference-plane tensor maps (16-bit pixels)
constexpr int MCF_MPW = 17;     // words per row pair of the intermediate

struct McFastSmem {
    __align__(128) uint32_t win[2][MCF_WROWS * MCF_WPW];   // double buffered: the next item's window streams in during the arithmetic (TMA destination: 128-byte aligned)
    uint32_t midv[(MCF_WROWS / 2) * MCF_MPW];   // [row pair][column] = (mid[2j][c], mid[2j + 1][c])
    __align__(16) uint16_t out[MC_TILE * MC_TILE];
};

struct McTaps {
    int e0, e1;        // first tap on the low half:  (F0 F1 F2 F3) (F4 F5 F6 F7)
    int o0, o1, o2;    // first tap on the high half: (0 F0 F1 F2) (F3 F4 F5 F6) (F7 0 0 0)
    int pad[3];
};
// the 6 x 15 sub-pel filters of tab::k_subpel_filters in packed form (mc_pack_taps), built once
__device__ McTaps g_subpel_packed[6 * 15];
__device__ __forceinline__ McTaps mc_pack_taps(const int8_t *f) {
    unsigned b[8];
#pragma unroll
    for (int k = 0; k < 8; k++) b[k] = (unsigned)(uint8_t)f[k];
    McTaps t;
    t.e0 = (int)(b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24));
    t.e1 = (int)(b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24));
    t.o0 = (int)((b[0] << 8) | (b[1] << 16) | (b[2] << 24));
    t.o1 = (int)(b[3] | (b[4] << 8) | (b[5] << 16) | (b[6] << 24));
    t.o2 = (int)b[7];
    t.pad[0] = t.pad[1] = t.pad[2] = 0;
    return t;
}
__device__ __forceinline__ int mc_fir_even(const McTaps &t, unsigned w0, unsigned w1, unsigned w2, unsigned w3, int init) {
    int a = __dp2a_lo((int)w0, t.e0, init);
    a = __dp2a_hi((int)w1, t.e0, a);
    a = __dp2a_lo((int)w2, t.e1, a);
    return __dp2a_hi((int)w3, t.e1, a);
}
__device__ __forceinline__ int mc_fir_odd(const McTaps &t, unsigned w0, unsigned w1, unsigned w2, unsigned w3, unsigned w4, int init) {
    int a = __dp2a_lo((int)w0, t.o0, init);
    a = __dp2a_hi((int)w1, t.o0, a);
    a = __dp2a_lo((int)w2, t.o1, a);
    a = __dp2a_hi((int)w3, t.o1, a);
    return __dp2a_lo((int)w4, t.o2, a);
}

__global__ void mc_pack_table_kernel() {
    const int i = threadIdx.x;
    if (i < 6 * 15) g_subpel_packed[i] = mc_pack_taps(tab::k_subpel_filters + i * 8);
}
__device__ __forceinline__ McTaps mc_load_taps(int set, int phase_minus1) {
    const int4 *p = (const int4 *)&g_subpel_packed[set * 15 + phase_minus1];
    const int4 a = p[0]; const int b = ((const int *)p)[4];
    McTaps t; t.e0 = a.x; t.e1 = a.y; t.o0 = a.z; t.o1 = a.w; t.o2 = b;
    return t;
}

// Bilinear is the 2-tap FIR (16 - m, m) at tap positions 0 and 1 (its window has no left/top margin) with a gain of 16 instead of 64
// (src/mc.rs:431-541); its two-step roundings collapse to the same single-shift forms.
__device__ __forceinline__ McTaps mc_bilin_taps(int m) {
    McTaps t;
    t.e0 = (int)((unsigned)(16 - m) | ((unsigned)m << 8)); t.e1 = 0;
    t.o0 = (int)(((unsigned)(16 - m) << 8) | ((unsigned)m << 16)); t.o1 = 0; t.o2 = 0;
    t.pad[0] = t.pad[1] = t.pad[2] = 0;
    return t;
}

// Window geometry of one tile: the filter support (8-tap: -3 .. +4, bilinear: 0 .. +1) decides
// where the window starts; it is staged from the even column at or left of it.
struct McWin {
    int xs, ys, ncols, nrows2, xa, par, nw;
    bool fh, fv, inside;
};
__device__ __forceinline__ McWin mc_window(const McRef &ref, int sx, int sy, int w, int h, int mx, int my, int filter2d) {
    McWin W;
    const int pre = filter2d == RB200_FILTER_2D_BILINEAR ? 0 : 3, ext = filter2d == RB200_FILTER_2D_BILINEAR ? 1 : 7;
    W.fh = mx != 0; W.fv = my != 0;
    W.xs = sx - (W.fh ? pre : 0); W.ys = sy - (W.fv ? pre : 0);
    W.ncols = w + (W.fh ? ext : 0);
    const int nrows = h + (W.fv ? ext : 0);
    W.nrows2 = (nrows + 1) & ~1;
    W.xa = W.xs & ~1; W.par = W.xs & 1;
    W.nw = (W.par + W.ncols + 1) >> 1;
    W.inside = W.xs >= 0 && W.xs + W.ncols <= ref.w;
    return W;
}

__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Stage the window of a tile: word k of row r = pixels (xa + 2k, xa + 2k + 1) of source row ys + r.
// 16-bit pixels of a window inside the picture go global -> shared asynchronously (LDGSTS), so
// the caller can overlap the fetch with the previous tile's arithmetic; the caller commits/waits.
// Other cases are staged synchronously.
template <typename BD>
__device__ __forceinline__ void mc_stage(uint32_t *win, const McRef &ref, const McWin &W) {
    using pixel = typename BD::pixel;
    const int lane = threadIdx.x & 31;
    if (W.inside) {
        const int k = lane & 15;
        if (k < W.nw) {
            if (BD::hbd && W.ys >= 0 && W.ys + W.nrows2 <= ref.h) {
                // no row is clamped: one pointer bump per row
                const int r0 = lane >> 4;
                const uint8_t *src = ref.base + (int64_t)(W.ys + r0) * ref.stride + (int64_t)(W.xa + 2 * k) * 2;
                uint32_t *d = win + r0 * MCF_WPW + k;
                const int64_t step = 2 * ref.stride;
                for (int r = r0; r < W.nrows2; r += 2, src += step, d += 2 * MCF_WPW) cp_async4(d, src);
            } else {
                for (int r = lane >> 4; r < W.nrows2; r += 2) {
                    const int yy = iclip(W.ys + r, 0, ref.h - 1);
                    const pixel *row = (const pixel *)(ref.base + (int64_t)yy * ref.stride) + W.xa + 2 * k;
                    if (BD::hbd) cp_async4(win + r * MCF_WPW + k, row);
                    else { const unsigned q = *(const uint16_t *)row; win[r * MCF_WPW + k] = (q & 0xff) | ((q & 0xff00) << 8); }
                }
            }
        }
    } else {
        uint16_t *we = (uint16_t *)win;
        const int ne = 2 * W.nw;
        for (int i = lane; i < W.nrows2 * ne; i += 32) {
            const int r = i / ne, e = i - r * ne;
            const int yy = iclip(W.ys + r, 0, ref.h - 1), xx = iclip(W.xa + e, 0, ref.w - 1);
            we[r * (2 * MCF_WPW) + e] = ((const pixel *)(ref.base + (int64_t)yy * ref.stride))[xx];
        }
    }
}

// 8-bit windows of the frame batch: the raw bytes go global -> shared asynchronously in 4-byte pieces (so that,
// like the 16-bit windows, they stream in behind the previous item's arithmetic); mc_expand8 then widens them
// into the pixel-pair words the filter passes read.
constexpr int MCF_RAW8_WPR = 7;   // 28 bytes per window row: up to 3 bytes of alignment + 23 pixels
__device__ __forceinline__ bool mc_raw8_ok(const McRef &ref, const McWin &W) {
    return W.inside && (W.xs & ~3) + 4 * MCF_RAW8_WPR <= ref.stride && !(((uintptr_t)ref.base | (uintptr_t)ref.stride) & 3);
}
__device__ __forceinline__ void mc_stage8_async(uint32_t *raw8, const McRef &ref, const McWin &W) {
    const int lane = threadIdx.x & 31, k = lane & 7;
    const int xa4 = W.xs & ~3, nw4 = ((W.xs & 3) + W.ncols + 3) >> 2;
    if (k < nw4) {
        for (int r = lane >> 3; r < W.nrows2; r += 4) {
            const int yy = iclip(W.ys + r, 0, ref.h - 1);
            cp_async4(raw8 + r * MCF_RAW8_WPR + k, ref.base + (int64_t)yy * ref.stride + xa4 + 4 * k);
        }
    }
}
__device__ __forceinline__ void mc_expand8(uint32_t *win, const uint32_t *raw8, const McWin &W) {
    const int lane = threadIdx.x & 31, k = lane & 15;
    const uint8_t *rb = (const uint8_t *)raw8 + ((W.xs & 3) & ~1) + 2 * k;   // the window's pair k within a raw row
    if (k < W.nw) {
        for (int r = lane >> 4; r < W.nrows2; r += 2) {
            const unsigned q = *(const uint16_t *)(rb + r * (4 * MCF_RAW8_WPR));
            win[r * MCF_WPW + k] = (q & 0xff) | ((q & 0xff00) << 8);
        }
    }
}

// put, w and h even and <= 16; `win` already staged (mc_stage) and visible to the warp.
// TW / TH: compile-time tile size (0 = use the run-time w / h) -- the common 16x16 and 8x8 tiles get
// fully unrolled task loops with constant index arithmetic.
// PREP: the `prep` form (int16, no clip, minus PREP_BIAS; src/mc.rs:277-349,543-606) is left in `tile`
// (row pitch MC_TILE) instead of pixels going to the picture.
template <typename BD, int TW, int TH, bool PREP = false>
__device__ void mc_tile_fast(McFastSmem &sm, const uint32_t *win, const McWin &W, int w_rt, int h_rt, int bw, int bh, int mx,
                             int my, int filter2d, uint8_t *out, int64_t out_pitch, int bdmax, uint16_t *tile = nullptr) {
    if (!PREP) tile = sm.out;
    constexpr int pb = PREP ? McBits<BD>::prep_bias : 0;
    const int w = TW ? TW : w_rt, h = TH ? TH : h_rt;
    const int lane = threadIdx.x & 31;
    const int ib = McBits<BD>::ib(bdmax);
    const bool fh = W.fh, fv = W.fv;
    const int nrows2 = W.nrows2, par = W.par;

    const int th_ = (0x111222000LL >> (4 * filter2d)) & 3, tv_ = (0x210210210LL >> (4 * filter2d)) & 3;
    const int ncp = w >> 1, cp_shift = 31 - __clz(ncp), w_shift = cp_shift + 1;
    uint32_t *out32 = (uint32_t *)tile;   // row pitch MC_TILE pixels = 8 words

    const bool bilin = filter2d == RB200_FILTER_2D_BILINEAR;
    const int base = bilin ? 4 : 6;
    if (fh) {
        const McTaps T = bilin ? mc_bilin_taps(mx) : mc_load_taps(bw > 4 ? th_ : 3 + (th_ & 1), mx - 1);
        const int sh1 = base - ib, r1 = (1 << sh1) >> 1;
        const int n_tasks = (nrows2 >> 1) << cp_shift;
        const int init = (fv || PREP) ? r1 : (1 << (base - 1)) + r1, sh = (fv || PREP) ? sh1 : base;
        constexpr int MAX_IT = TW ? ((TH + 8) / 2 * (TW / 2) + 31) / 32 : 3;
#pragma unroll
        for (int it = 0; it < MAX_IT; it++) {
            const int t = lane + 32 * it;
            if (t >= n_tasks) break;
            const int cp = t & (ncp - 1), rp = t >> cp_shift;
            int a[2][2];
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                const uint32_t *wv = win + (2 * rp + rr) * MCF_WPW + cp;
                const unsigned w0 = wv[0], w1 = wv[1], w2 = wv[2], w3 = wv[3], w4 = wv[4];
                if (par == 0) { a[rr][0] = mc_fir_even(T, w0, w1, w2, w3, init); a[rr][1] = mc_fir_odd(T, w0, w1, w2, w3, w4, init); }
                else { a[rr][0] = mc_fir_odd(T, w0, w1, w2, w3, w4, init); a[rr][1] = mc_fir_even(T, w1, w2, w3, w4, init); }
            }
            if (fv) {
#pragma unroll
                for (int cc = 0; cc < 2; cc++)
                    sm.midv[rp * MCF_MPW + 2 * cp + cc] = __byte_perm((unsigned)(a[0][cc] >> sh), (unsigned)(a[1][cc] >> sh), 0x5410);
            } else {
#pragma unroll
                for (int rr = 0; rr < 2; rr++) {
                    const unsigned p0 = PREP ? (unsigned)((a[rr][0] >> sh) - pb) : (unsigned)iclip(a[rr][0] >> sh, 0, bdmax);
                    const unsigned p1 = PREP ? (unsigned)((a[rr][1] >> sh) - pb) : (unsigned)iclip(a[rr][1] >> sh, 0, bdmax);
                    out32[(2 * rp + rr) * (MC_TILE / 2) + cp] = (p0 & 0xffff) | (p1 << 16);
                }
            }
        }
    } else if (fv) {
        // no horizontal filter: re-pack the raw pixels as vertical pairs
        const int n_tasks = (nrows2 >> 1) << cp_shift;
        for (int t = lane; t < n_tasks; t += 32) {
            const int cp = t & (ncp - 1), rp = t >> cp_shift;
            const uint32_t *r0 = win + (2 * rp) * MCF_WPW + cp, *r1p = r0 + MCF_WPW;
            unsigned a0, a1;   // pixels (c, c + 1) of the two rows
            if (par == 0) { a0 = r0[0]; a1 = r1p[0]; }
            else { a0 = __funnelshift_r(r0[0], r0[1], 16); a1 = __funnelshift_r(r1p[0], r1p[1], 16); }
            sm.midv[rp * MCF_MPW + 2 * cp] = __byte_perm(a0, a1, 0x5410);
            sm.midv[rp * MCF_MPW + 2 * cp + 1] = __byte_perm(a0, a1, 0x7632);
        }
    } else {
        const int n_tasks = h << cp_shift;
        for (int t = lane; t < n_tasks; t += 32) {
            const int cp = t & (ncp - 1), r = t >> cp_shift;
            const uint32_t *wv = win + r * MCF_WPW + cp;
            const unsigned v = par == 0 ? wv[0] : __funnelshift_r(wv[0], wv[1], 16);
            if (PREP) {
                const unsigned p0 = (unsigned)((int)((v & 0xffff) << ib) - pb), p1 = (unsigned)((int)((v >> 16) << ib) - pb);
                out32[r * (MC_TILE / 2) + cp] = (p0 & 0xffff) | (p1 << 16);
            } else {
                out32[r * (MC_TILE / 2) + cp] = v;
            }
        }
    }
    __syncwarp();
    if (fv) {
        const McTaps T = bilin ? mc_bilin_taps(my) : mc_load_taps(bh > 4 ? tv_ : 3 + (tv_ & 1), my - 1);
        const int sh2 = PREP ? (fh ? base : base - ib) : (fh ? base + ib : base), r2 = (1 << sh2) >> 1;
        const int n_tasks = (h >> 1) << w_shift;
        constexpr int MAX_IT = TW ? (TH / 2 * TW + 31) / 32 : 4;
#pragma unroll
        for (int it = 0; it < MAX_IT; it++) {
            const int t = lane + 32 * it;
            if (t >= n_tasks) break;
            const int c = t & (w - 1), rp = t >> w_shift;
            const uint32_t *mv = sm.midv + rp * MCF_MPW + c;
            const unsigned v0 = mv[0], v1 = mv[MCF_MPW], v2 = mv[2 * MCF_MPW], v3 = mv[3 * MCF_MPW], v4 = mv[4 * MCF_MPW];
            const int a0 = mc_fir_even(T, v0, v1, v2, v3, r2), a1 = mc_fir_odd(T, v0, v1, v2, v3, v4, r2);
            tile[(2 * rp) * MC_TILE + c] = PREP ? (uint16_t)((a0 >> sh2) - pb) : (uint16_t)iclip(a0 >> sh2, 0, bdmax);
            tile[(2 * rp + 1) * MC_TILE + c] = PREP ? (uint16_t)((a1 >> sh2) - pb) : (uint16_t)iclip(a1 >> sh2, 0, bdmax);
        }
        __syncwarp();
    }
    if (PREP) return;
    // ---- tile -> picture, widest aligned stores the row allows
    if (BD::hbd) {
        const int row_bytes = w * 2;
        if (row_bytes >= 16 && !(((uintptr_t)out | (uintptr_t)out_pitch) & 15)) {
            const int upr = row_bytes >> 4;   // 16-byte units per row (1 or 2)
            for (int t = lane; t < h * upr; t += 32) {
                const int r = upr == 2 ? t >> 1 : t, u = upr == 2 ? t & 1 : 0;
                *(uint4 *)(out + (int64_t)r * out_pitch + u * 16) = *(const uint4 *)(sm.out + r * MC_TILE + u * 8);
            }
        } else {
            for (int t = lane; t < h << cp_shift; t += 32) {
                const int cp = t & (ncp - 1), r = t >> cp_shift;
                *(uint32_t *)(out + (int64_t)r * out_pitch + cp * 4) = out32[r * (MC_TILE / 2) + cp];
            }
        }
    } else {
        for (int t = lane; t < h << cp_shift; t += 32) {
            const int cp = t & (ncp - 1), r = t >> cp_shift;
            const unsigned v = out32[r * (MC_TILE / 2) + cp];
            *(uint16_t *)(out + (int64_t)r * out_pitch + cp * 2) = (uint16_t)((v & 0xff) | ((v >> 8) & 0xff00));
        }
    }
    __syncwarp();
}

struct McRefSet {
    Rb200Planes p[8];
};

// One work item unpacked into registers together with everything derived from it once: the reference
// plane and the window of its first tile.  (The 16-byte record is fetched with one 128-bit load.)
#ifndef MC_BATCH_CTAS
#define MC_BATCH_CTAS 6      // resident CTAs per SM the batch kernel is compiled for (register budget 80)
#endif
struct McJob {
    int dst_x, dst_y, src_x, src_y, w, h, plane, slot, mx, my, filter2d;
    McRef ref;
    McWin W;      // first tile
    bool fast;
};
__device__ __forceinline__ McJob mc_load_job(const Rb200McItem *__restrict__ items, int idx, const McRefSet &refs, int ref_w,
                                             int ref_h, int ss_hor, int ss_ver) {
    const uint4 q = __ldg((const uint4 *)(items + idx));
    McJob j;
    j.dst_x = (int)(short)(q.x & 0xffff); j.dst_y = (int)q.x >> 16;
    j.src_x = (int)(short)(q.y & 0xffff); j.src_y = (int)q.y >> 16;
    j.w = q.z & 0xff; j.h = (q.z >> 8) & 0xff; j.plane = (q.z >> 16) & 0xff;
    const int slot = (q.z >> 24) & 7;
    j.slot = slot;
    j.mx = q.w & 0xff; j.my = (q.w >> 8) & 0xff; j.filter2d = (q.w >> 16) & 0xff;
    const Rb200Planes &rp = refs.p[slot];
    j.ref.base = plane_ptr(rp, j.plane);
    j.ref.stride = plane_stride(rp, j.plane);
    j.ref.w = j.plane ? (ref_w + ss_hor) >> ss_hor : ref_w;
    j.ref.h = j.plane ? (ref_h + ss_ver) >> ss_ver : ref_h;
    j.fast = !((j.w | j.h) & 1) && (j.w >= MC_TILE || !(j.w & (j.w - 1)));
    j.W = mc_window(j.ref, j.src_x, j.src_y, imin(MC_TILE, j.w), imin(MC_TILE, j.h), j.mx, j.my, j.filter2d);
    return j;
}

// Frame batch.  Warps take chunks of MC_CHUNK consecutive items from a shared counter.  The per-item set-up
// (unpack, reference plane, window geometry, destination pointer) is the same ~150 instructions whether one
// lane or thirty-two execute it, so it is done lane-parallel -- lane k prepares item k of the chunk -- and
// parked in shared memory; the warp then walks the chunk reading each prepared job back with six broadcast
// 128-bit loads.  While item k is filtered, the first window of item k + 1 is already in flight (cp.async).
// Items wider/taller than 16 are walked tile by tile.
constexpr int MC_CHUNK = 16;
struct __align__(16) McJobS {
    const uint8_t *rbase; int64_t rstride;
    uint8_t *dst; int64_t dstride;
    int rw, rh, src_x, src_y;
    int w, h, phase /* mx | my << 8 | filter2d << 16 | fast << 24 */, xs;
    int ys, ncols, nrows2, xa;
    int par, nw, flags /* fh | fv << 1 | inside << 2 */, map /* tensor map of the reference plane (slot * 3 + plane), -1: none */;
};
// Tensor maps of the reference planes, [slot * 3 + plane], boxes of MCF_TMA_W x MCF_TMA_H pixels (16-bit pictures).
struct McTmaMaps { CUtensorMap m[24]; };
static_assert(sizeof(McJobS) == 96, "six 128-bit words");

template <typename BD>
__global__ void __launch_bounds__(MC_WARPS * 32, MC_BATCH_CTAS)
mc_batch_kernel(Rb200Planes dst, McRefSet refs, int ref_w, int ref_h, int ss_hor, int ss_ver,
                const Rb200McItem *__restrict__ items, int n_items, int bdmax, int *__restrict__ chunk_counter, int chunk,
                const __grid_constant__ McTmaMaps maps, unsigned tma_mask) {
    __shared__ struct { McFastSmem fast; McSmem slow; } smem[MC_WARPS];   // not a union: a prefetch may be in flight
    __shared__ __align__(8) uint64_t tma_bar[MC_WARPS][2];                // one per window buffer of each warp
    __shared__ McJobS jobs[MC_WARPS][MC_CHUNK];
    __shared__ uint32_t raw8_s[BD::hbd ? 1 : MC_WARPS][2][BD::hbd ? 1 : MCF_WROWS * MCF_RAW8_WPR];   // 8-bit only
    uint32_t (*raw8)[BD::hbd ? 1 : MCF_WROWS * MCF_RAW8_WPR] = raw8_s[BD::hbd ? 0 : threadIdx.x >> 5];
    auto stage_first_gather = [&](int b, const McRef &ref, const McWin &W) {
        if (!BD::hbd && mc_raw8_ok(ref, W)) mc_stage8_async(raw8[b], ref, W);
        else mc_stage<BD>(smem[threadIdx.x >> 5].fast.win[b], ref, W);
    };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    McFastSmem &sm = smem[warp].fast;
    // Windows that lie inside the reference are fetched by the copy engine: lane 0 names the box (first column rounded
    // down to a 16-byte boundary of the plane, MCF_TMA_W x MCF_TMA_H pixels) and the warp later waits on the buffer's
    // mbarrier; no lane computes an address.  Windows that reach over the picture edge need replicated pixels, which
    // the copy engine does not make (it fills with zeros): those keep the clamped gather of mc_stage.
    unsigned tma_phase = 0;                                   // bit b: parity the next wait on buffer b's barrier uses
    if (BD::hbd && tma_mask) {
        if (lane == 0) { mbar_init(&tma_bar[warp][0], 1); mbar_init(&tma_bar[warp][1], 1); mbar_fence_init(); }
        __syncwarp();
    }
    auto tma_window = [&](int map, const McWin &W, int rh) { return BD::hbd && map >= 0 && W.inside && W.ys >= 0 && W.ys + W.nrows2 <= rh; };
    auto tma_issue = [&](int b, int map, const McWin &W) {
        fence_proxy_async_smem();      // this lane's earlier writes to the buffer (a gathered window) are ordered before the engine's
        __syncwarp();
        if (lane == 0) {
            mbar_arrive_expect_tx(&tma_bar[warp][b], MCF_TMA_W * MCF_TMA_H * 2);
            tma_load_2d(sm.win[b], &maps.m[map], W.xs & ~7, W.ys, &tma_bar[warp][b]);
        }
    };
    auto tma_wait = [&](int b) {
        mbar_wait(&tma_bar[warp][b], (tma_phase >> b) & 1u);
        tma_phase ^= 1u << b;
    };
    auto read_job = [&](int k) {
        McJobS J;
        const uint4 *p = (const uint4 *)&jobs[warp][k];
        uint4 *q = (uint4 *)&J;
#pragma unroll
        for (int i = 0; i < 6; i++) q[i] = p[i];
        return J;
    };
    auto win_of = [](const McJobS &J) {
        McWin W;
        W.xs = J.xs; W.ys = J.ys; W.ncols = J.ncols; W.nrows2 = J.nrows2; W.xa = J.xa; W.par = J.par; W.nw = J.nw;
        W.fh = J.flags & 1; W.fv = (J.flags >> 1) & 1; W.inside = (J.flags >> 2) & 1;
        return W;
    };
    auto ref_of = [](const McJobS &J) { McRef r; r.base = J.rbase; r.stride = J.rstride; r.w = J.rw; r.h = J.rh; return r; };
    int c_static = blockIdx.x * MC_WARPS + warp;      // without a dispenser: chunks dealt round-robin
    for (;;) {
        int c = c_static;
        if (chunk_counter) {
            if (lane == 0) c = atomicAdd(chunk_counter, 1);
            c = __shfl_sync(0xffffffffu, c, 0);
        }
        c_static += gridDim.x * MC_WARPS;
        const int base = c * chunk;
        if (base >= n_items) break;
        const int cnt = imin(chunk, n_items - base);
        if (lane < cnt) {
            const McJob j = mc_load_job(items, base + lane, refs, ref_w, ref_h, ss_hor, ss_ver);
            McJobS J;
            J.rbase = j.ref.base; J.rstride = j.ref.stride; J.rw = j.ref.w; J.rh = j.ref.h;
            J.dstride = plane_stride(dst, j.plane);
            J.dst = plane_ptr(dst, j.plane) + (int64_t)j.dst_y * J.dstride + (int64_t)j.dst_x * sizeof(typename BD::pixel);
            J.src_x = j.src_x; J.src_y = j.src_y; J.w = j.w; J.h = j.h;
            J.phase = j.mx | (j.my << 8) | (j.filter2d << 16) | ((int)j.fast << 24);
            J.xs = j.W.xs; J.ys = j.W.ys; J.ncols = j.W.ncols; J.nrows2 = j.W.nrows2; J.xa = j.W.xa; J.par = j.W.par; J.nw = j.W.nw;
            J.flags = (int)j.W.fh | ((int)j.W.fv << 1) | ((int)j.W.inside << 2);
            J.map = ((tma_mask >> (j.slot * 3 + j.plane)) & 1u) ? j.slot * 3 + j.plane : -1;
            jobs[warp][lane] = J;
        }
        __syncwarp();
        auto stage_first = [&](int b, const McJobS &J) {
            if (!(J.phase >> 24)) return;
            const McWin W = win_of(J);
            if (tma_window(J.map, W, J.rh)) tma_issue(b, J.map, W);
            else stage_first_gather(b, ref_of(J), W);
        };
        int buf = 0;
        stage_first(0, read_job(0));
        cp_async_commit();
        for (int k = 0; k < cnt; k++) {
            if (k + 1 < cnt) stage_first(buf ^ 1, read_job(k + 1));   // only one prepared job is held in registers at a time
            cp_async_commit();
            cp_async_wait<1>();      // the current item's first window has landed (gathered windows)
            __syncwarp();
            const McJobS cur = read_job(k);
            const int mx = cur.phase & 0xff, my = (cur.phase >> 8) & 0xff, filter2d = (cur.phase >> 16) & 0xff;
            const bool fast = cur.phase >> 24;
            const bool first_by_tma = fast && tma_window(cur.map, win_of(cur), cur.rh);
            if (first_by_tma) tma_wait(buf);                          // (windows fetched by the copy engine)
            // a fetched window starts at the 16-byte boundary at or left of it: skip to the word of pixel xs & ~1
            const uint32_t *win0 = sm.win[buf] + (first_by_tma ? (cur.xs & 6) >> 1 : 0);
            if (!BD::hbd && fast && mc_raw8_ok(ref_of(cur), win_of(cur))) {   // raw bytes have landed: widen them
                mc_expand8(sm.win[buf], raw8[buf], win_of(cur));
                __syncwarp();
            }
            if (fast && cur.w <= MC_TILE && cur.h <= MC_TILE) {
                // the common case: the whole block is the prefetched tile
                const McWin W = win_of(cur);
                if (cur.w == 16 && cur.h == 16)
                    mc_tile_fast<BD, 16, 16>(sm, win0, W, 16, 16, 16, 16, mx, my, filter2d, cur.dst, cur.dstride, bdmax);
                else if (cur.w == 8 && cur.h == 8)
                    mc_tile_fast<BD, 8, 8>(sm, win0, W, 8, 8, 8, 8, mx, my, filter2d, cur.dst, cur.dstride, bdmax);
                else
                    mc_tile_fast<BD, 0, 0>(sm, win0, W, cur.w, cur.h, cur.w, cur.h, mx, my, filter2d, cur.dst, cur.dstride, bdmax);
            } else {
                const McRef ref = ref_of(cur);
                for (int ty = 0; ty < cur.h; ty += MC_TILE) {
                    for (int tx = 0; tx < cur.w; tx += MC_TILE) {
                        uint8_t *o = cur.dst + (int64_t)ty * cur.dstride + (int64_t)tx * sizeof(typename BD::pixel);
                        const int tw = imin(MC_TILE, cur.w - tx), th = imin(MC_TILE, cur.h - ty);
                        if (fast) {
                            const McWin W = mc_window(ref, cur.src_x + tx, cur.src_y + ty, tw, th, mx, my, filter2d);
                            const uint32_t *wt = win0;
                            if (tx | ty) {   // further tiles of a large block: staged in place, not prefetched
                                if (tma_window(cur.map, W, cur.rh)) {
                                    tma_issue(buf, cur.map, W);
                                    tma_wait(buf);
                                    wt = sm.win[buf] + ((W.xs & 6) >> 1);
                                } else {
                                    mc_stage<BD>(sm.win[buf], ref, W);
                                    cp_async_commit();
                                    cp_async_wait<0>();
                                    wt = sm.win[buf];
                                }
                                __syncwarp();
                            }
                            if (tw == 16 && th == 16)
                                mc_tile_fast<BD, 16, 16>(sm, wt, W, tw, th, cur.w, cur.h, mx, my, filter2d, o, cur.dstride, bdmax);
                            else
                                mc_tile_fast<BD, 0, 0>(sm, wt, W, tw, th, cur.w, cur.h, mx, my, filter2d, o, cur.dstride, bdmax);
                        } else {
                            mc_tile<BD, false>(smem[warp].slow, ref, cur.src_x + tx, cur.src_y + ty, tw, th, cur.w, cur.h, mx, my,
                                               filter2d, o, cur.dstride, bdmax);
                        }
                    }
                }
            }
            buf ^= 1;
        }
        cp_async_wait<0>();
        __syncwarp();      // every lane is done with jobs[] before the next chunk overwrites it
    }
}

template <typename BD, bool PREP>
__device__ __forceinline__ int mc_scaled_px(const McRef &ref, int ox, int oy, int w, int h, int mx, int my, int dx, int dy,
                                            int filter2d, int x, int y, int bdmax);

// ---------------------------------------------------------------- compound blocks (batch)
// One warp per block.  Per 16x16 luma tile: both references are predicted in `prep` form into
// shared memory, combined into the luma plane, and -- for the segmentation compound -- the full
// resolution blend mask stays in shared memory for the co-located chroma tiles, which derive
// their sub-sampled mask exactly as w_mask does (src/mc.rs:812-883).
struct McCompSmem {
    McFastSmem fast;
    McSmem slow;
    __align__(16) int16_t tmp[2][MC_TILE * MC_TILE];
    uint8_t mask[MC_TILE * MC_TILE];
};

// One prediction of the compound walk: block `idx`, luma tile (tx, ty), plane pl, reference i.
struct McCompJob {
    Rb200CompItem it;
    int idx, tx, ty, pl, i;
    bool valid;
};

template <typename BD>
__global__ void __launch_bounds__(MC_WARPS * 32)
mc_comp_batch_kernel(Rb200Planes dst, McRefSet refs, int ref_w, int ref_h, int layout,
                     const Rb200CompItem *__restrict__ items, int n_items, int bdmax, McGmvSet gmv, McRefDims dims) {
    using pixel = typename BD::pixel;
    __shared__ McCompSmem smem[MC_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_warps = gridDim.x * MC_WARPS;
    McCompSmem &sm = smem[warp];
    const int n_planes = layout == RB200_LAYOUT_I400 ? 1 : 3;
    const int ss_hor_c = layout != RB200_LAYOUT_I444, ss_ver_c = layout == RB200_LAYOUT_I420;
    const int ib = McBits<BD>::ib(bdmax), pb = McBits<BD>::prep_bias;
    const int bitdepth = BD::hbd ? bpc_from_max(bdmax) : 8;
    const int mask_sh = bitdepth + ib - 4, mask_rnd = 1 << (mask_sh - 5);

    // geometry of a job: reference plane, source position and phase (src/recon.rs:2047-2055,2100-2101)
    struct Geo { McRef ref; int pw, ph, bw, bh, px0, py0, sx, sy, mx, my; bool fast, warped, scaled; int pos_x, pos_y, step_x, step_y; };
    auto slot_dim = [&](int slot, bool height) {          // (no run-time index into the parameter struct)
        int v = 0;
#pragma unroll
        for (int s_ = 0; s_ < 8; s_++) if (s_ == slot) v = height ? dims.h[s_] : dims.w[s_];
        return v;
    };
    auto geo = [&](const McCompJob &j) {
        Geo g;
        // the prediction is the reference's global-motion warp (warp_affine with frame_hdr.gmv[ref], src/recon.rs:3253-3268;
        // chroma only when the chroma block is at least 8x8, :3352-3369)
        g.warped = (j.it.warp_mask >> ((j.pl ? 2 : 0) + j.i)) & 1;
        const int ss_hor = j.pl ? ss_hor_c : 0, ss_ver = j.pl ? ss_ver_c : 0;
        const Rb200Planes &rp = refs.p[(j.i ? j.it.ref[1] : j.it.ref[0]) & 7];
        g.ref.base = plane_ptr(rp, j.pl);
        g.ref.stride = plane_stride(rp, j.pl);
        g.ref.w = j.pl ? (ref_w + ss_hor) >> ss_hor : ref_w;
        g.ref.h = j.pl ? (ref_h + ss_ver) >> ss_ver : ref_h;
        g.pw = imin(MC_TILE, j.it.w - j.tx) >> ss_hor; g.ph = imin(MC_TILE, j.it.h - j.ty) >> ss_ver;
        g.bw = j.it.w >> ss_hor; g.bh = j.it.h >> ss_ver;
        g.px0 = (j.it.x + j.tx) >> ss_hor; g.py0 = (j.it.y + j.ty) >> ss_ver;
        const int mvy = j.i ? j.it.mv[1][0] : j.it.mv[0][0], mvx = j.i ? j.it.mv[1][1] : j.it.mv[0][1];
        g.mx = (mvx & (15 >> !ss_hor)) << !ss_hor; g.my = (mvy & (15 >> !ss_ver)) << !ss_ver;
        g.sx = g.px0 + (mvx >> (3 + ss_hor)); g.sy = g.py0 + (mvy >> (3 + ss_ver));
        // a reference of another size (src/recon.rs:2116-2199): position and step of the scaled prediction from the
        // slot's size -- scale = ((ref << 14) + (cur >> 1)) / cur, step = (scale + 8) >> 4 (src/decode.rs, f.svc) -- and
        // scale_mv() of the block's 1/16-pel position in this plane
        const int slot = (j.i ? j.it.ref[1] : j.it.ref[0]) & 7;
        const int rw = slot_dim(slot, false), rh = slot_dim(slot, true);
        g.scaled = rw > 0 && rh > 0 && (rw != ref_w || rh != ref_h);
        if (g.scaled) {
            g.ref.w = j.pl ? (rw + ss_hor) >> ss_hor : rw;
            g.ref.h = j.pl ? (rh + ss_ver) >> ss_ver : rh;
            const int scale_x = ((rw << 14) + (ref_w >> 1)) / ref_w, scale_y = ((rh << 14) + (ref_h >> 1)) / ref_h;
            g.step_x = (scale_x + 8) >> 4; g.step_y = (scale_y + 8) >> 4;
            auto scale_mv = [](int v, int scale) {
                const long long t = (long long)v * scale + (long long)(scale - 0x4000) * 8;
                const long long a = ((t < 0 ? -t : t) + 128) >> 8;
                return (int)(t < 0 ? -a : a) + 32;
            };
            g.pos_x = scale_mv(((j.it.x >> ss_hor) << 4) + mvx * (1 << !ss_hor), scale_x);
            g.pos_y = scale_mv(((j.it.y >> ss_ver) << 4) + mvy * (1 << !ss_ver), scale_y);
        }
        g.fast = !((g.pw | g.ph) & 1) && !(g.pw & (g.pw - 1)) && !g.warped && !g.scaled;
        return g;
    };
    // warp8x8t over the 8x8s of a tile into tmp (prep form), one warp: src/recon.rs:2311-2400, src/mc.rs:958-1030
    auto warp_tile = [&](const McCompJob &j, const Geo &g, int16_t *tmp) {
        const int ss_hor = j.pl ? ss_hor_c : 0, ss_ver = j.pl ? ss_ver_c : 0;
        const int slot = (j.i ? j.it.ref[1] : j.it.ref[0]) & 7;
            int32_t mat[6]; int ab[4];
#pragma unroll
        for (int s_ = 0; s_ < 8; s_++)
            if (s_ == slot) {
#pragma unroll
                for (int k = 0; k < 6; k++) mat[k] = gmv.g[s_].matrix[k];
#pragma unroll
                for (int k = 0; k < 4; k++) ab[k] = gmv.g[s_].abcd[k];
            }
        int16_t *mid = (int16_t *)sm.fast.midv;          // 15 x 8 intermediates (free: no mc_tile_fast is running)
        for (int sy8 = 0; sy8 < g.ph; sy8 += 8)
            for (int sx8 = 0; sx8 < g.pw; sx8 += 8) {
                // position of this 8x8 inside the block's plane
                const int bx = (j.tx >> ss_hor) + sx8, by = (j.ty >> ss_ver) + sy8;
                const int src_y = j.it.y + ((by + 4) << ss_ver), src_x = j.it.x + ((bx + 4) << ss_hor);
                const int64_t mvx = ((int64_t)mat[2] * src_x + (int64_t)mat[3] * src_y + mat[0]) >> ss_hor;
                const int64_t mvy = ((int64_t)mat[4] * src_x + (int64_t)mat[5] * src_y + mat[1]) >> ss_ver;
                const int dx = (int)(mvx >> 16) - 4, dy = (int)(mvy >> 16) - 4;
                const int mx = (((int)mvx & 0xffff) - ab[0] * 4 - ab[1] * 7) & ~0x3f;
                const int my = (((int)mvy & 0xffff) - ab[2] * 4 - ab[3] * 4) & ~0x3f;
                for (int e = lane; e < 15 * 8; e += 32) {
                    const int yy = e >> 3, xx = e & 7;
                    const int tmx = mx + yy * ab[1] + xx * ab[0];
                    const int8_t *f = tab::k_warp_filter + (64 + ((tmx + 512) >> 10)) * 8;
                    int acc = (1 << (7 - ib)) >> 1;
                    const pixel *row = (const pixel *)(g.ref.base + (int64_t)iclip(dy + yy - 3, 0, g.ref.h - 1) * g.ref.stride);
#pragma unroll
                    for (int k = 0; k < 8; k++) acc += f[k] * (int)row[iclip(dx + xx + k - 3, 0, g.ref.w - 1)];
                    mid[e] = (int16_t)(acc >> (7 - ib));
                }
                __syncwarp();
                for (int e = lane; e < 64; e += 32) {
                    const int yy = e >> 3, xx = e & 7;
                    const int tmy = my + yy * ab[3] + xx * ab[2];
                    const int8_t *f = tab::k_warp_filter + (64 + ((tmy + 512) >> 10)) * 8;
                    int acc = 0;
#pragma unroll
                    for (int k = 0; k < 8; k++) acc += f[k] * (int)mid[(yy + k) * 8 + xx];
                    tmp[(sy8 + yy) * MC_TILE + sx8 + xx] = (int16_t)(((acc + 64) >> 7) - pb);
                }
                __syncwarp();
            }
    };
    auto advance = [&](McCompJob &j) {
        if (++j.i < 2) return;
        j.i = 0;
        if (++j.pl < n_planes) return;
        j.pl = 0;
        if ((j.tx += MC_TILE) < j.it.w) return;
        j.tx = 0;
        if ((j.ty += MC_TILE) < j.it.h) return;
        j.ty = 0;
        j.idx += n_warps;
        j.valid = j.idx < n_items;
        if (j.valid) j.it = items[j.idx];
    };
    auto stage = [&](const McCompJob &j, uint32_t *win) {
        const Geo g = geo(j);
        if (g.fast) mc_stage<BD>(win, g.ref, mc_window(g.ref, g.sx, g.sy, g.pw, g.ph, g.mx, g.my, j.it.filter2d));
    };

    McCompJob cur;
    cur.idx = blockIdx.x * MC_WARPS + warp;
    if (cur.idx >= n_items) return;
    cur.it = items[cur.idx];
    cur.tx = cur.ty = cur.pl = cur.i = 0; cur.valid = true;
    int buf = 0;
    stage(cur, sm.fast.win[0]);
    cp_async_commit();
    while (cur.valid) {
        McCompJob nxt = cur;
        advance(nxt);
        if (nxt.valid) stage(nxt, sm.fast.win[buf ^ 1]);    // streams in during this job's arithmetic
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        const Geo g = geo(cur);
        const Rb200CompItem &it = cur.it;
        if (g.fast) {
            const McWin W = mc_window(g.ref, g.sx, g.sy, g.pw, g.ph, g.mx, g.my, it.filter2d);
            if (g.pw == 16 && g.ph == 16)
                mc_tile_fast<BD, 16, 16, true>(sm.fast, sm.fast.win[buf], W, g.pw, g.ph, g.bw, g.bh, g.mx, g.my, it.filter2d, nullptr, 0, bdmax, (uint16_t *)sm.tmp[cur.i]);
            else if (g.pw == 8 && g.ph == 8)
                mc_tile_fast<BD, 8, 8, true>(sm.fast, sm.fast.win[buf], W, g.pw, g.ph, g.bw, g.bh, g.mx, g.my, it.filter2d, nullptr, 0, bdmax, (uint16_t *)sm.tmp[cur.i]);
            else
                mc_tile_fast<BD, 0, 0, true>(sm.fast, sm.fast.win[buf], W, g.pw, g.ph, g.bw, g.bh, g.mx, g.my, it.filter2d, nullptr, 0, bdmax, (uint16_t *)sm.tmp[cur.i]);
        } else if (g.warped) {
            warp_tile(cur, g, sm.tmp[cur.i]);
        } else if (g.scaled) {
            // prep_8tap_scaled per pixel (rare path): the tile's pixels of the block-wide scaled prediction
            const int ox = g.pos_x >> 10, oy = g.pos_y >> 10, smx = g.pos_x & 0x3ff, smy = g.pos_y & 0x3ff;
            const int bx0 = g.px0 - (it.x >> (cur.pl ? ss_hor_c : 0)), by0 = g.py0 - (it.y >> (cur.pl ? ss_ver_c : 0));   // tile origin inside the block
            for (int e = lane; e < g.pw * g.ph; e += 32) {
                const int r = e / g.pw, c = e - r * g.pw;
                sm.tmp[cur.i][r * MC_TILE + c] = (int16_t)mc_scaled_px<BD, true>(g.ref, ox, oy, g.bw, g.bh, smx, smy, g.step_x, g.step_y,
                                                                                 it.filter2d, bx0 + c, by0 + r, bdmax);
            }
        } else {
            mc_tile<BD, true>(sm.slow, g.ref, g.sx, g.sy, g.pw, g.ph, g.bw, g.bh, g.mx, g.my, it.filter2d, sm.tmp[cur.i], MC_TILE, bdmax);
        }
        if (cur.i == 1) {
            // ---- both predictions of this plane's tile are in tmp[]: combine into the picture
            __syncwarp();
            const int pl = cur.pl, pw = g.pw, ph = g.ph;
            const int ss_hor = pl ? ss_hor_c : 0, ss_ver = pl ? ss_ver_c : 0;
            const int s1 = it.comp_type == RB200_COMP_AVG ? 0 : it.mask_sign;   // tmp1 = tmp[mask_sign] for masks
            uint8_t *dbase = plane_ptr(dst, pl);
            const int64_t dstride = plane_stride(dst, pl);
            const int16_t *t1 = sm.tmp[s1], *t2 = sm.tmp[s1 ^ 1];
            const int pw_shift = 31 - __clz(pw);
            const bool pow2 = !(pw & (pw - 1));
            for (int e = lane; e < pw * ph; e += 32) {
                const int r = pow2 ? e >> pw_shift : e / pw, c = e - r * pw;
                const int a = t1[r * MC_TILE + c], b = t2[r * MC_TILE + c];
                int v;
                if (it.comp_type == RB200_COMP_AVG) {
                    v = (a + b + (1 << ib) + pb * 2) >> (ib + 1);
                } else if (it.comp_type == RB200_COMP_WEIGHTED_AVG) {
                    // w_avg(tmp[0], tmp[1], weight): no sign swap
                    const int a0 = sm.tmp[0][r * MC_TILE + c], b0 = sm.tmp[1][r * MC_TILE + c];
                    v = (a0 * it.jnt_weight + b0 * (16 - it.jnt_weight) + (8 << ib) + pb * 16) >> (ib + 4);
                } else {
                    int m;
                    if (pl == 0) {
                        if (it.comp_type == RB200_COMP_WEDGE) {
                            m = wedge_mask_at(it.w, it.h, it.wedge_idx & 15, cur.tx + c, cur.ty + r);
                        } else {
                            const int d = a - b;
                            m = imin(38 + (((d < 0 ? -d : d) + mask_rnd) >> mask_sh), 64);
                        }
                        sm.mask[r * MC_TILE + c] = (uint8_t)m;
                    } else {
                        const uint8_t *mp = sm.mask + (r << ss_ver) * MC_TILE + (c << ss_hor);
                        if (ss_hor && ss_ver) m = (mp[0] + mp[1] + mp[MC_TILE] + mp[MC_TILE + 1] + 2 - it.mask_sign) >> 2;
                        else if (ss_hor) m = (mp[0] + mp[1] + 1 - it.mask_sign) >> 1;
                        else m = mp[0];
                    }
                    v = (a * m + b * (64 - m) + (32 << ib) + pb * 64) >> (ib + 6);
                }
                ((pixel *)(dbase + (int64_t)(g.py0 + r) * dstride))[g.px0 + c] = (pixel)iclip(v, 0, bdmax);
            }
            __syncwarp();
        }
        cur = nxt;
        buf ^= 1;
    }
    cp_async_wait<0>();
}

// The packed sub-pel tap table (g_subpel_packed) is filled once per device; every launcher whose kernel goes through
// mc_load_taps calls this first (the put batch and the compound batch -- a frame may hold only compound blocks).
static int mc_ensure_packed_taps(cudaStream_t st) {
    static std::mutex mu;
    static bool packed[64] = {};
    std::lock_guard<std::mutex> lock(mu);
    int dev = 0;
    RB_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !packed[dev]) {
        mc_pack_table_kernel<<<1, 96, 0, st>>>();
        RB_LAUNCH_CHECK();
        RB_CUDA(cudaStreamSynchronize(st));   // once per device: visible to every stream that follows
        packed[dev] = true;
    }
    return 0;
}

int mc_comp_batch_launch(const Rb200Planes &dst, const Rb200Planes *refs, int n_refs, int ref_w, int ref_h, int layout,
                         const Rb200CompItem *d_items, int n, int bdmax, cudaStream_t st) {
    static const McGmvSet none = {};
    return mc_comp_batch_launch_gmv(dst, refs, n_refs, ref_w, ref_h, layout, d_items, n, bdmax, st, none);
}

int mc_comp_batch_launch_gmv(const Rb200Planes &dst, const Rb200Planes *refs, int n_refs, int ref_w, int ref_h, int layout,
                             const Rb200CompItem *d_items, int n, int bdmax, cudaStream_t st, const McGmvSet &gmv, const McRefDims *dims_in) {
    if (n <= 0) return 0;
    McRefDims dims = {};
    if (dims_in) dims = *dims_in;
    McRefSet rs = {};
    for (int i = 0; i < n_refs && i < 8; i++) rs.p[i] = refs[i];
    { const int r = mc_ensure_packed_taps(st); if (r) return r; }
    const int grid = imin((n + MC_WARPS - 1) / MC_WARPS, 148 * 4);   // persistent warps walk the list
    if (bdmax > 255) mc_comp_batch_kernel<BD16><<<grid, MC_WARPS * 32, 0, st>>>(dst, rs, ref_w, ref_h, layout, d_items, n, bdmax, gmv, dims);
    else mc_comp_batch_kernel<BD8><<<grid, MC_WARPS * 32, 0, st>>>(dst, rs, ref_w, ref_h, layout, d_items, n, bdmax, gmv, dims);
    RB_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------- warped blocks (batch)
// One CTA per block, one warp per 8x8 (all planes): position / phase per 8x8 as warp_affine
// (src/recon.rs:2311-2400), the 15x8 horizontal pass into warp-private shared memory, then the
// vertical pass; source coordinates are clamped (emu_edge).
template <typename BD>
__global__ void __launch_bounds__(MC_WARPS * 32)
mc_warp_batch_kernel(Rb200Planes dst, McRefSet refs, int ref_w, int ref_h, int layout,
                     const Rb200WarpItem *__restrict__ items, int n_items, int bdmax) {
    using pixel = typename BD::pixel;
    __shared__ int16_t mid_s[MC_WARPS][15 * 8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if ((int)blockIdx.x >= n_items) return;
    const Rb200WarpItem it = items[blockIdx.x];
    int16_t *mid = mid_s[warp];
    const int n_planes = layout == RB200_LAYOUT_I400 ? 1 : 3;
    const int ss_hor_c = layout != RB200_LAYOUT_I444, ss_ver_c = layout == RB200_LAYOUT_I420;
    const int ib = McBits<BD>::ib(bdmax);
    const Rb200Planes &rp = refs.p[it.ref & 7];
    const int a0 = it.abcd[0], a1 = it.abcd[1], a2 = it.abcd[2], a3 = it.abcd[3];
    // enumerate the 8x8s of all planes
    const int nbx0 = it.w >> 3, nby0 = it.h >> 3;
    const int nbxc = it.w >> (3 + ss_hor_c), nbyc = it.h >> (3 + ss_ver_c);
    const int n0 = nbx0 * nby0, nc = n_planes > 1 ? nbxc * nbyc : 0;
    for (int t = warp; t < n0 + 2 * nc; t += MC_WARPS) {
        const int pl = t < n0 ? 0 : (t < n0 + nc ? 1 : 2);
        const int u = t - (pl == 0 ? 0 : (pl == 1 ? n0 : n0 + nc));
        const int ss_hor = pl ? ss_hor_c : 0, ss_ver = pl ? ss_ver_c : 0;
        const int nbx = pl ? nbxc : nbx0;
        const int x = (u % nbx) * 8, y = (u / nbx) * 8;
        const int width = pl ? (ref_w + ss_hor) >> ss_hor : ref_w, height = pl ? (ref_h + ss_ver) >> ss_ver : ref_h;
        const int src_y = it.y + ((y + 4) << ss_ver), src_x = it.x + ((x + 4) << ss_hor);
        const int64_t mvx = ((int64_t)it.matrix[2] * src_x + (int64_t)it.matrix[3] * src_y + it.matrix[0]) >> ss_hor;
        const int64_t mvy = ((int64_t)it.matrix[4] * src_x + (int64_t)it.matrix[5] * src_y + it.matrix[1]) >> ss_ver;
        const int dx = (int)(mvx >> 16) - 4, dy = (int)(mvy >> 16) - 4;
        const int mx = (((int)mvx & 0xffff) - a0 * 4 - a1 * 7) & ~0x3f;
        const int my = (((int)mvy & 0xffff) - a2 * 4 - a3 * 4) & ~0x3f;
        const uint8_t *rbase = plane_ptr(rp, pl);
        const int64_t rstride = plane_stride(rp, pl);
        for (int i = lane; i < 15 * 8; i += 32) {
            const int yy = i >> 3, xx = i & 7;
            const int tmx = mx + yy * a1 + xx * a0;
            const int8_t *f = tab::k_warp_filter + (64 + ((tmx + 512) >> 10)) * 8;
            int acc = (1 << (7 - ib)) >> 1;
            const pixel *row = (const pixel *)(rbase + (int64_t)iclip(dy + yy - 3, 0, height - 1) * rstride);
#pragma unroll
            for (int k = 0; k < 8; k++) acc += f[k] * (int)row[iclip(dx + xx + k - 3, 0, width - 1)];
            mid[i] = (int16_t)(acc >> (7 - ib));
        }
        __syncwarp();
        uint8_t *dbase = plane_ptr(dst, pl);
        const int64_t dstride = plane_stride(dst, pl);
        const int px0 = (it.x >> ss_hor) + x, py0 = (it.y >> ss_ver) + y;
        for (int i = lane; i < 64; i += 32) {
            const int yy = i >> 3, xx = i & 7;
            const int tmy = my + yy * a3 + xx * a2;
            const int8_t *f = tab::k_warp_filter + (64 + ((tmy + 512) >> 10)) * 8;
            int acc = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) acc += f[k] * (int)mid[(yy + k) * 8 + xx];
            ((pixel *)(dbase + (int64_t)(py0 + yy) * dstride))[px0 + xx] =
                (pixel)iclip((acc + ((1 << (7 + ib)) >> 1)) >> (7 + ib), 0, bdmax);
        }
        __syncwarp();
    }
}

int mc_warp_batch_launch(const Rb200Planes &dst, const Rb200Planes *refs, int n_refs, int ref_w, int ref_h, int layout,
                         const Rb200WarpItem *d_items, int n, int bdmax, cudaStream_t st) {
    if (n <= 0) return 0;
    McRefSet rs = {};
    for (int i = 0; i < n_refs && i < 8; i++) rs.p[i] = refs[i];
    if (bdmax > 255) mc_warp_batch_kernel<BD16><<<n, MC_WARPS * 32, 0, st>>>(dst, rs, ref_w, ref_h, layout, d_items, n, bdmax);
    else mc_warp_batch_kernel<BD8><<<n, MC_WARPS * 32, 0, st>>>(dst, rs, ref_w, ref_h, layout, d_items, n, bdmax);
    RB_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------- OBMC strips (batch)
// One warp per strip: the neighbour's prediction of the strip goes into a warp-private tile, then
// blend_h / blend_v (src/mc.rs:742-810) onto the block's own prediction.
template <typename BD>
__global__ void __launch_bounds__(MC_WARPS * 32)
mc_obmc_batch_kernel(Rb200Planes dst, McRefSet refs, int ref_w, int ref_h, int ss_hor_c, int ss_ver_c,
                     const Rb200McItem *__restrict__ items, int n_items, int bdmax) {
    using pixel = typename BD::pixel;
    __shared__ struct { McSmem slow; pixel lap[MC_TILE * MC_TILE]; } smem[MC_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int idx = blockIdx.x * MC_WARPS + warp;
    if (idx >= n_items) return;
    const Rb200McItem it = items[idx];
    const Rb200Planes &rp = refs.p[it.ref & 7];
    const int ss_hor = it.plane ? ss_hor_c : 0, ss_ver = it.plane ? ss_ver_c : 0;
    McRef ref;
    ref.base = plane_ptr(rp, it.plane);
    ref.stride = plane_stride(rp, it.plane);
    ref.w = it.plane ? (ref_w + ss_hor) >> ss_hor : ref_w;
    ref.h = it.plane ? (ref_h + ss_ver) >> ss_ver : ref_h;
    const bool above = it.flags == RB200_MC_OBMC_ABOVE;
    const int v_mul = 4 >> ss_ver;
    // the prediction the reference makes for an ABOVE strip is ((oh4 * 3 + 3) >> 2) units tall
    const int pred_h = above ? (((it.h / v_mul) * 3 + 3) >> 2) * v_mul : it.h;
    const int lim_r = above ? (it.h * 3) >> 2 : it.h, lim_c = above ? it.w : (it.w * 3) >> 2;
    uint8_t *dbase = plane_ptr(dst, it.plane);
    const int64_t dstride = plane_stride(dst, it.plane);
    for (int ty = 0; ty < pred_h; ty += MC_TILE) {
        for (int tx = 0; tx < it.w; tx += MC_TILE) {
            const int tw = imin(MC_TILE, it.w - tx), th = imin(MC_TILE, pred_h - ty);
            mc_tile<BD, false>(smem[warp].slow, ref, it.src_x + tx, it.src_y + ty, tw, th, it.w, pred_h, it.mx, it.my,
                               it.filter2d, smem[warp].lap, MC_TILE * sizeof(pixel), bdmax);
            __syncwarp();
            for (int e = lane; e < tw * th; e += 32) {
                const int r = e / tw, c = e - r * tw;
                const int row = ty + r, col = tx + c;
                if (row >= lim_r || col >= lim_c) continue;
                const int m = above ? tab::k_obmc_masks[it.h + row] : tab::k_obmc_masks[it.w + col];
                pixel *d = (pixel *)(dbase + (int64_t)(it.dst_y + row) * dstride) + it.dst_x + col;
                *d = (pixel)(((int)*d * (64 - m) + (int)smem[warp].lap[r * MC_TILE + c] * m + 32) >> 6);
            }
            __syncwarp();
        }
    }
}

int mc_obmc_batch_launch(const Rb200Planes &dst, const Rb200Planes *refs, int n_refs, int ref_w, int ref_h, int ss_hor,
                         int ss_ver, const Rb200McItem *d_items, int n, int bdmax, cudaStream_t st) {
    if (n <= 0) return 0;
    McRefSet rs = {};
    for (int i = 0; i < n_refs && i < 8; i++) rs.p[i] = refs[i];
    const int grid = (n + MC_WARPS - 1) / MC_WARPS;
    if (bdmax > 255) mc_obmc_batch_kernel<BD16><<<grid, MC_WARPS * 32, 0, st>>>(dst, rs, ref_w, ref_h, ss_hor, ss_ver, d_items, n, bdmax);
    else mc_obmc_batch_kernel<BD8><<<grid, MC_WARPS * 32, 0, st>>>(dst, rs, ref_w, ref_h, ss_hor, ss_ver, d_items, n, bdmax);
    RB_LAUNCH_CHECK();
    return 0;
}

// Per-call: a single prediction block over a staged source rectangle; one warp per 16x16 tile.
template <typename BD, bool PREP>
__global__ void __launch_bounds__(MC_WARPS * 32)
mc_one_kernel(McRef ref, int ox, int oy, int w, int h, int mx, int my, int filter2d, void *out, int64_t out_pitch,
              int bdmax) {
    __shared__ McSmem smem[MC_WARPS];
    const int warp = threadIdx.x >> 5;
    const int tiles_x = (w + MC_TILE - 1) / MC_TILE, tiles_y = (h + MC_TILE - 1) / MC_TILE;
    const int t = blockIdx.x * MC_WARPS + warp;
    if (t >= tiles_x * tiles_y) return;
    const int ty = (t / tiles_x) * MC_TILE, tx = (t % tiles_x) * MC_TILE;
    void *o = PREP ? (void *)((int16_t *)out + (int64_t)ty * out_pitch + tx)
                   : (void *)((uint8_t *)out + (int64_t)ty * out_pitch + (int64_t)tx * sizeof(typename BD::pixel));
    mc_tile<BD, PREP>(smem[warp], ref, ox + tx, oy + ty, imin(MC_TILE, w - tx), imin(MC_TILE, h - ty), w, h, mx, my,
                      filter2d, o, out_pitch, bdmax);
}

// ------------------------------------------------------------------ scaled
// put_8tap_scaled / prep_8tap_scaled / bilin scaled: src/mc.rs:212-275,351-429,496-541,608-652.
// One thread per output pixel; position (mx + x*dx) is the closed form of the
// reference's running imx / ioff accumulation.
template <typename BD, bool PREP>
__device__ __forceinline__ int mc_scaled_px(const McRef &ref, int ox, int oy, int w, int h, int mx, int my, int dx, int dy,
                                            int filter2d, int x, int y, int bdmax) {
    using pixel = typename BD::pixel;
    const int ib = McBits<BD>::ib(bdmax);
    const int px = mx + x * dx, py = my + y * dy;
    const int ix = px >> 10, iy = py >> 10, fx = (px & 0x3ff) >> 6, fy = (py & 0x3ff) >> 6;
    auto at = [&](int r, int c) -> int {
        const int yy = iclip(oy + r, 0, ref.h - 1), xx = iclip(ox + c, 0, ref.w - 1);
        return ((const pixel *)(ref.base + (int64_t)yy * ref.stride))[xx];
    };
    int v;
    if (filter2d == RB200_FILTER_2D_BILINEAR) {
        int m[2];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int a = at(iy + k, ix), b = at(iy + k, ix + 1);
            m[k] = (int16_t)((16 * a + fx * (b - a) + ((1 << (4 - ib)) >> 1)) >> (4 - ib));
        }
        const int s = 16 * m[0] + fy * (m[1] - m[0]);
        v = PREP ? ((s + 8) >> 4) - McBits<BD>::prep_bias : iclip((s + ((1 << (4 + ib)) >> 1)) >> (4 + ib), 0, bdmax);
    } else {
        const int th_ = (0x111222000LL >> (4 * filter2d)) & 3, tv_ = (0x210210210LL >> (4 * filter2d)) & 3;
        const int8_t *FH = fx ? tab::subpel(w > 4 ? th_ : 3 + (th_ & 1), fx - 1) : nullptr;
        const int8_t *FV = fy ? tab::subpel(h > 4 ? tv_ : 3 + (tv_ & 1), fy - 1) : nullptr;
        int m[8];
        // the source pointer of the reference is pre-advanced by -3 rows; row iy+k-3 <-> mid row (iy + k)
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (!FV && k != 3) { m[k] = 0; continue; }
            const int r = iy + k - 3;
            if (FH) {
                int acc = (1 << (6 - ib)) >> 1;
#pragma unroll
                for (int j = 0; j < 8; j++) acc += FH[j] * at(r, ix + j - 3);
                m[k] = (int16_t)(acc >> (6 - ib));
            } else {
                m[k] = (int16_t)(at(r, ix) << ib);
            }
        }
        if (FV) {
            int acc = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) acc += FV[k] * m[k];
            v = PREP ? ((acc + 32) >> 6) - McBits<BD>::prep_bias
                     : iclip((acc + ((1 << (6 + ib)) >> 1)) >> (6 + ib), 0, bdmax);
        } else {
            v = PREP ? m[3] - McBits<BD>::prep_bias : iclip((m[3] + ((1 << ib) >> 1)) >> ib, 0, bdmax);
        }
    }
    return v;
}

template <typename BD, bool PREP>
__global__ void mc_scaled_kernel(McRef ref, int ox, int oy, int w, int h, int mx, int my, int dx, int dy,
                                 int filter2d, void *out, int64_t out_pitch, int bdmax) {
    using pixel = typename BD::pixel;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w || y >= h) return;
    const int v = mc_scaled_px<BD, PREP>(ref, ox, oy, w, h, mx, my, dx, dy, filter2d, x, y, bdmax);
    if (PREP) ((int16_t *)out)[(int64_t)y * out_pitch + x] = (int16_t)v;
    else ((pixel *)((uint8_t *)out + (int64_t)y * out_pitch))[x] = (pixel)v;
}

// Frame batch of scaled predictions: one CTA per item, a thread per pixel.  (Scaled references are rare
// -- spatial scalability and reference scaling -- so this path favours simplicity over the IDP pipeline.)
template <typename BD>
__global__ void __launch_bounds__(128)
mc_scaled_batch_kernel(Rb200Planes dst, McRefSet refs, McRefDims dims, int ss_hor_c, int ss_ver_c,
                       const Rb200McScaledItem *__restrict__ items, int bdmax) {
    using pixel = typename BD::pixel;
    const Rb200McScaledItem it = items[blockIdx.x];
    const int ss_hor = it.plane ? ss_hor_c : 0, ss_ver = it.plane ? ss_ver_c : 0;
    const Rb200Planes &rp = refs.p[it.ref & 7];
    McRef ref;
    ref.base = plane_ptr(rp, it.plane);
    ref.stride = plane_stride(rp, it.plane);
    ref.w = (dims.w[it.ref & 7] + ss_hor) >> ss_hor;
    ref.h = (dims.h[it.ref & 7] + ss_ver) >> ss_ver;
    uint8_t *dbase = plane_ptr(dst, it.plane);
    const int64_t dstride = plane_stride(dst, it.plane);
    // src/recon_tmpl.c:1029-1062: the block starts at (pos >> 10), phase pos & 0x3ff
    const int ox = it.pos_x >> 10, oy = it.pos_y >> 10, mx = it.pos_x & 0x3ff, my = it.pos_y & 0x3ff;
    if (it.flags == RB200_MC_PUT) {
        for (int i = threadIdx.x; i < it.w * it.h; i += blockDim.x) {
            const int y = i / it.w, x = i - y * it.w;
            const int v = mc_scaled_px<BD, false>(ref, ox, oy, it.w, it.h, mx, my, it.step_x, it.step_y, it.filter2d, x, y, bdmax);
            ((pixel *)(dbase + (int64_t)(it.dst_y + y) * dstride))[it.dst_x + x] = (pixel)v;
        }
        return;
    }
    // OBMC strip (obmc(), src/recon.rs:2205-2309) from a reference of another size: w x h is the blend area; the
    // neighbour's prediction is ((oh4 * 3 + 3) >> 2) units tall for an ABOVE strip (that size picks the 4-tap filters);
    // blend_h / blend_v touch 3/4 of the rows / columns.  Every thread predicts and blends its own pixel.
    const bool above = it.flags == RB200_MC_OBMC_ABOVE;
    const int v_mul = 4 >> ss_ver;
    const int pred_h = above ? (((it.h / v_mul) * 3 + 3) >> 2) * v_mul : it.h;
    const int lim_r = above ? (it.h * 3) >> 2 : it.h, lim_c = above ? it.w : (it.w * 3) >> 2;
    for (int i = threadIdx.x; i < lim_c * lim_r; i += blockDim.x) {
        const int y = i / lim_c, x = i - y * lim_c;
        const int v = mc_scaled_px<BD, false>(ref, ox, oy, it.w, pred_h, mx, my, it.step_x, it.step_y, it.filter2d, x, y, bdmax);
        const int m = above ? tab::k_obmc_masks[it.h + y] : tab::k_obmc_masks[it.w + x];
        pixel *d = (pixel *)(dbase + (int64_t)(it.dst_y + y) * dstride) + it.dst_x + x;
        *d = (pixel)(((int)*d * (64 - m) + v * m + 32) >> 6);
    }
}

int mc_scaled_batch_launch(const Rb200Planes &dst, const Rb200Planes *refs, int n_refs, const McRefDims &dims, int ss_hor,
                           int ss_ver, const Rb200McScaledItem *d_items, int n, int bdmax, cudaStream_t st) {
    if (n <= 0) return 0;
    McRefSet rs = {};
    for (int i = 0; i < n_refs && i < 8; i++) rs.p[i] = refs[i];
    if (bdmax > 255) mc_scaled_batch_kernel<BD16><<<n, 128, 0, st>>>(dst, rs, dims, ss_hor, ss_ver, d_items, bdmax);
    else mc_scaled_batch_kernel<BD8><<<n, 128, 0, st>>>(dst, rs, dims, ss_hor, ss_ver, d_items, bdmax);
    RB_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------- compound
// avg / w_avg / mask / w_mask: src/mc.rs:654-740,812-883 == src/mc_tmpl.c:542-604,643-699
enum { CMP_AVG, CMP_WAVG, CMP_MASK };
template <typename BD>
__global__ void compound_kernel(int mode, void *dst, int64_t dstride, const int16_t *__restrict__ t1,
                                const int16_t *__restrict__ t2, int w, int h, int weight,
                                const uint8_t *__restrict__ mask, int bdmax) {
    using pixel = typename BD::pixel;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w * h) return;
    const int y = i / w, x = i - y * w;
    const int ib = McBits<BD>::ib(bdmax), pb = McBits<BD>::prep_bias;
    const int a = t1[i], b = t2[i];
    int v;
    if (mode == CMP_AVG) v = (a + b + (1 << ib) + pb * 2) >> (ib + 1);
    else if (mode == CMP_WAVG) v = (a * weight + b * (16 - weight) + (8 << ib) + pb * 16) >> (ib + 4);
    else { const int m = mask[i]; v = (a * m + b * (64 - m) + (32 << ib) + pb * 64) >> (ib + 6); }
    ((pixel *)((uint8_t *)dst + (int64_t)y * dstride))[x] = (pixel)iclip(v, 0, bdmax);
}

// w_mask: one thread per (1 << ss_hor) x (1 << ss_ver) pixel group, so that the
// sub-sampled mask value is produced without the reference's even/odd row carry.
template <typename BD>
__global__ void w_mask_kernel(void *dst, int64_t dstride, const int16_t *__restrict__ t1,
                              const int16_t *__restrict__ t2, int w, int h, uint8_t *__restrict__ mask, int sign,
                              int ss_hor, int ss_ver, int bdmax) {
    using pixel = typename BD::pixel;
    const int gw = w >> ss_hor, gh = h >> ss_ver;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= gw * gh) return;
    const int gy = g / gw, gx = g - gy * gw;
    const int ib = McBits<BD>::ib(bdmax), pb = McBits<BD>::prep_bias;
    const int bitdepth = BD::hbd ? bpc_from_max(bdmax) : 8;
    const int sh = ib + 6, rnd = (32 << ib) + pb * 64;
    const int mask_sh = bitdepth + ib - 4, mask_rnd = 1 << (mask_sh - 5);
    int msum = 0;
    for (int j = 0; j <= ss_ver; j++) {
        for (int k = 0; k <= ss_hor; k++) {
            const int y = (gy << ss_ver) + j, x = (gx << ss_hor) + k;
            const int a = t1[y * w + x], b = t2[y * w + x];
            const int d = a - b;
            const int m = imin(38 + (((d < 0 ? -d : d) + mask_rnd) >> mask_sh), 64);
            ((pixel *)((uint8_t *)dst + (int64_t)y * dstride))[x] =
                (pixel)iclip((a * m + b * (64 - m) + rnd) >> sh, 0, bdmax);
            msum += m;
        }
    }
    // 444: m ; 422: (m+n+1-sign)>>1 ; 420: (m+n+m'+n'+2-sign)>>2
    const int n = (1 + ss_hor) * (1 + ss_ver);
    mask[g] = (uint8_t)(n == 1 ? msum : n == 2 ? (msum + 1 - sign) >> 1 : (msum + 2 - sign) >> 2);
}

// blend / blend_v / blend_h: src/mc.rs:742-810 == src/mc_tmpl.c:606-641
template <typename BD>
__global__ void blend_kernel(int dir, void *dst, int64_t dstride, const void *__restrict__ tmp, int w, int h,
                             const uint8_t *__restrict__ mask) {
    using pixel = typename BD::pixel;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w * h) return;
    const int y = i / w, x = i - y * w;
    int m;
    if (dir == 0) m = mask[i];
    else if (dir == 1) { if (x >= (w * 3) >> 2) return; m = tab::k_obmc_masks[w + x]; }
    else { if (y >= (h * 3) >> 2) return; m = tab::k_obmc_masks[h + y]; }
    pixel *d = (pixel *)((uint8_t *)dst + (int64_t)y * dstride) + x;
    const int a = *d, b = ((const pixel *)tmp)[i];
    *d = (pixel)((a * (64 - m) + b * m + 32) >> 6);
}

// warp_affine_8x8 / 8x8t: src/mc.rs:885-1030 == src/mc_tmpl.c:701-772.  One block of 64 threads
// per 8x8; 15x8 horizontal pass into shared int16, then vertical pass.
template <typename BD, bool PREP>
__global__ void __launch_bounds__(64)
warp8x8_kernel(McRef ref, int ox, int oy, void *out, int64_t out_pitch, const int16_t *__restrict__ abcd_d,
               int mx, int my, int bdmax) {
    using pixel = typename BD::pixel;
    __shared__ int16_t mid[15 * 8];
    const int ib = McBits<BD>::ib(bdmax);
    const int a0 = abcd_d[0], a1 = abcd_d[1], a2 = abcd_d[2], a3 = abcd_d[3];
    for (int i = threadIdx.x; i < 15 * 8; i += 64) {
        const int y = i >> 3, x = i & 7;
        const int tmx = mx + y * a1 + x * a0;
        const int8_t *f = tab::k_warp_filter + (64 + ((tmx + 512) >> 10)) * 8;
        int acc = (1 << (7 - ib)) >> 1;
        const int yy = iclip(oy + y - 3, 0, ref.h - 1);
        const pixel *row = (const pixel *)(ref.base + (int64_t)yy * ref.stride);
#pragma unroll
        for (int k = 0; k < 8; k++) acc += f[k] * (int)row[iclip(ox + x + k - 3, 0, ref.w - 1)];
        mid[i] = (int16_t)(acc >> (7 - ib));
    }
    __syncthreads();
    const int y = threadIdx.x >> 3, x = threadIdx.x & 7;
    const int tmy = my + y * a3 + x * a2;
    const int8_t *f = tab::k_warp_filter + (64 + ((tmy + 512) >> 10)) * 8;
    int acc = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) acc += f[k] * (int)mid[(y + k) * 8 + x];
    if (PREP) ((int16_t *)out)[(int64_t)y * out_pitch + x] = (int16_t)(((acc + 64) >> 7) - McBits<BD>::prep_bias);
    else ((pixel *)((uint8_t *)out + (int64_t)y * out_pitch))[x] =
             (pixel)iclip((acc + ((1 << (7 + ib)) >> 1)) >> (7 + ib), 0, bdmax);
}

// emu_edge: src/mc.rs:1032-1112 == src/mc_tmpl.c:774-826 (a clamped gather)
template <typename BD>
__global__ void emu_edge_kernel(McRef ref, int x0, int y0, int bw, int bh, void *dst, int64_t dstride) {
    using pixel = typename BD::pixel;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= bw || y >= bh) return;
    const int yy = iclip(y0 + y, 0, ref.h - 1), xx = iclip(x0 + x, 0, ref.w - 1);
    ((pixel *)((uint8_t *)dst + (int64_t)y * dstride))[x] = ((const pixel *)(ref.base + (int64_t)yy * ref.stride))[xx];
}

// resize (super-resolution upscaling): src/mc.rs:1114-1172 == src/mc_tmpl.c:828-858
template <typename BD>
__global__ void resize_kernel(void *dst, int64_t dstride, const void *src, int64_t sstride, int dst_w, int h,
                              int src_w, int dx, int mx0, int bdmax) {
    using pixel = typename BD::pixel;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dst_w || y >= h) return;
    const int64_t pos = (int64_t)mx0 + (int64_t)x * dx;  // running mx: src_x = -1 + (pos >> 14), phase pos & 0x3fff
    const int src_x = -1 + (int)(pos >> 14), ph = (int)(pos & 0x3fff);
    const int8_t *F = tab::k_resize_filter + (ph >> 8) * 8;
    const pixel *s = (const pixel *)((const uint8_t *)src + (int64_t)y * sstride);
    int acc = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) acc += F[k] * (int)s[iclip(src_x + k - 3, 0, src_w - 1)];
    ((pixel *)((uint8_t *)dst + (int64_t)y * dstride))[x] = (pixel)iclip((-acc + 64) >> 7, 0, bdmax);
}

int resize_plane_launch(void *dst, int64_t dstride, const void *src, int64_t sstride, int dst_w, int h, int src_w, int dx,
                        int mx0, int bdmax, cudaStream_t st) {
    dim3 grid((dst_w + 127) / 128, h);
    if (bdmax > 255) resize_kernel<BD16><<<grid, 128, 0, st>>>(dst, dstride, src, sstride, dst_w, h, src_w, dx, mx0, bdmax);
    else resize_kernel<BD8><<<grid, 128, 0, st>>>(dst, dstride, src, sstride, dst_w, h, src_w, dx, mx0, bdmax);
    RB_LAUNCH_CHECK();
    return 0;
}

// Tensor maps of the reference planes for the batch kernel (16-bit pictures).  `cache` remembers what each map was
// encoded for, so a context whose references did not move re-encodes nothing.  A plane the copy engine cannot address
// (base or stride not 16-byte aligned: only possible through rb200_mc_batch with foreign planes) gets no map.
void mc_ref_maps_update(McRefMapCache &c, const Rb200Planes *refs, int n_refs, int ref_w, int ref_h, int ss_hor, int ss_ver) {
    static_assert(sizeof(c.maps) == sizeof(McTmaMaps), "McRefMapCache::maps holds a McTmaMaps");
    McTmaMaps &M = *(McTmaMaps *)c.maps;
    for (int sl = 0; sl < 8; sl++)
        for (int p = 0; p < 3; p++) {
            const int i = sl * 3 + p;
            const void *base = sl < n_refs ? refs[sl].data[p] : nullptr;
            const int64_t stride = sl < n_refs ? refs[sl].stride[p] : 0;
            const int w = p ? (ref_w + ss_hor) >> ss_hor : ref_w, h = p ? (ref_h + ss_ver) >> ss_ver : ref_h;
            if (c.base[i] == base && c.stride[i] == stride && c.w[i] == w && c.h[i] == h) continue;
            c.base[i] = base; c.stride[i] = stride; c.w[i] = w; c.h[i] = h;
            c.mask &= ~(1u << i);
            if (!base || ((uintptr_t)base & 15) || (stride & 15) || stride <= 0) continue;
            if (tma_encode_plane(&M.m[i], base, 2, w, h, stride, MCF_TMA_W, MCF_TMA_H) == 0) c.mask |= 1u << i;
        }
}

int mc_batch_launch(const Rb200Planes &dst, const Rb200Planes *refs, int n_refs, int ref_w, int ref_h, int ss_hor,
                    int ss_ver, const Rb200McItem *d_items, int n, int bdmax, cudaStream_t st, int *counter, McRefMapCache *map_cache) {
    if (n <= 0) return 0;
    McRefSet rs = {};
    for (int i = 0; i < n_refs && i < 8; i++) rs.p[i] = refs[i];
    McRefMapCache local;
    if (bdmax > 255) {
        if (!map_cache) { memset(&local, 0, sizeof(local)); map_cache = &local; }
        mc_ref_maps_update(*map_cache, refs, n_refs, ref_w, ref_h, ss_hor, ss_ver);
    }
    static const McTmaMaps no_maps = {};
    const McTmaMaps &maps = bdmax > 255 ? *(const McTmaMaps *)map_cache->maps : no_maps;
    const unsigned tma_mask = bdmax > 255 ? map_cache->mask : 0u;
    { const int r = mc_ensure_packed_taps(st); if (r) return r; }
    // persistent warps: 148 SMs x 6 resident CTAs, capped by the item count
    // chunk size: as large as MC_CHUNK when there is enough work to give every resident warp a few chunks
    const int resident_warps = 148 * MC_BATCH_CTAS * MC_WARPS;
    const int chunk = imax(2, imin(MC_CHUNK, n / (2 * resident_warps)));
    const int n_chunks = (n + chunk - 1) / chunk;
    const int grid = imin((n_chunks + MC_WARPS - 1) / MC_WARPS, 148 * MC_BATCH_CTAS);
    // counter: chunk dispenser of this launch (4 bytes of device memory owned by the caller, e.g. one per frame
    // context -- launches of different contexts overlap); nullptr = static round-robin
    if (counter) RB_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));
    if (bdmax > 255) mc_batch_kernel<BD16><<<grid, MC_WARPS * 32, 0, st>>>(dst, rs, ref_w, ref_h, ss_hor, ss_ver, d_items, n, bdmax, counter, chunk, maps, tma_mask);
    else mc_batch_kernel<BD8><<<grid, MC_WARPS * 32, 0, st>>>(dst, rs, ref_w, ref_h, ss_hor, ss_ver, d_items, n, bdmax, counter, chunk, maps, tma_mask);
    RB_LAUNCH_CHECK();
    return 0;
}

}  // namespace rb200

using namespace rb200;

// ---------------------------------------------------------------- C ABI
extern "C" int rb200_mc_batch(const Rb200Planes *dst, const Rb200Planes *refs, int n_refs, int ref_w, int ref_h,
                              int ss_hor, int ss_ver, const Rb200McItem *d_items, int n_items, int bitdepth_max,
                              void *stream) {
    if (!dst || !refs || n_refs < 1 || n_refs > 8) return set_error(-22, "mc_batch: bad argument");
    return mc_batch_launch(*dst, refs, n_refs, ref_w, ref_h, ss_hor, ss_ver, d_items, n_items, bitdepth_max,
                           (cudaStream_t)stream);
}

namespace {

inline size_t pxsz(int bdmax) { return bdmax > 255 ? 2 : 1; }

// Stage the source rectangle a (w x h) prediction at phase (mx, my) reads:
// rows -3..h+3 / columns -3..w+3 for 8-tap, +1 for bilinear, nothing extra for phase 0.
struct SrcStage {
    DevRect rect;
    McRef ref;
    int ox, oy;
    int setup(HostCall &hc, const void *src, ptrdiff_t ss, int w, int h, int before_x, int after_x, int before_y,
              int after_y, int bdmax) {
        const size_t px = pxsz(bdmax);
        const uint8_t *row0 = (const uint8_t *)src - (int64_t)before_y * ss - (int64_t)before_x * (int64_t)px;
        const int cols = w + before_x + after_x, rows = h + before_y + after_y;
        if (hc.rect_up(rect, row0, ss, cols * px, rows)) return hc.err;
        ref.base = rect.dptr; ref.stride = rect.dpitch; ref.w = cols; ref.h = rows;
        ox = before_x; oy = before_y;
        return 0;
    }
};

int mc_common(bool prep, bool scaled, int filter2d, void *dst, ptrdiff_t ds, const void *src, ptrdiff_t ss, int w,
              int h, int mx, int my, int dx, int dy, int bdmax) {
    if (filter2d < 0 || filter2d >= RB200_N_2D_FILTERS || w < 1 || h < 1 || w > 128 || h > 256 || !dst || !src)
        return set_error(-22, "mc: bad argument");
    const bool bilin = filter2d == RB200_FILTER_2D_BILINEAR;
    const size_t px = pxsz(bdmax);
    int bx = 0, ax = 0, by = 0, ay = 0;
    if (scaled) {
        // reference reads rows -3 .. ((h-1)*dy+my >> 10) + 4 and columns -3 .. ((w-1)*dx+mx >> 10) + 4
        const int lastx = ((w - 1) * dx + mx) >> 10, lasty = ((h - 1) * dy + my) >> 10;
        if (bilin) { ax = lastx + 2 - w; ay = lasty + 2 - h; }
        else { bx = by = 3; ax = lastx + 5 - w; ay = lasty + 5 - h; }
    } else {
        if (mx) { if (bilin) ax = 1; else { bx = 3; ax = 4; } }
        if (my) { if (bilin) ay = 1; else { by = 3; ay = 4; } }
    }
    const size_t src_bytes = DevRect::bytes_for((size_t)(w + bx + ax) * px, h + by + ay);
    const size_t dst_bytes = prep ? (size_t)w * h * 2 : DevRect::bytes_for(w * px, h);
    HostCall hc(2 * (src_bytes + dst_bytes));
    SrcStage s;
    if (s.setup(hc, src, ss, w, h, bx, ax, by, ay, bdmax)) return hc.err;
    DevRect drect;
    void *out; int64_t pitch;
    if (prep) { out = hc.dev((size_t)w * h * 2); pitch = w; }
    else { if (hc.rect_up(drect, dst, ds, w * px, h)) return hc.err; out = drect.dptr; pitch = drect.dpitch; }
    if (hc.err) return hc.err;
    const bool hbd = bdmax > 255;
    if (scaled) {
        dim3 grid((w + 63) / 64, h), blk(64);
#define L(BD, P) mc_scaled_kernel<BD, P><<<grid, blk, 0, hc.stream()>>>(s.ref, s.ox, s.oy, w, h, mx, my, dx, dy, filter2d, out, pitch, bdmax)
        if (hbd) { if (prep) L(BD16, true); else L(BD16, false); } else { if (prep) L(BD8, true); else L(BD8, false); }
#undef L
    } else {
        const int tiles = ((w + MC_TILE - 1) / MC_TILE) * ((h + MC_TILE - 1) / MC_TILE);
        const int grid = (tiles + MC_WARPS - 1) / MC_WARPS;
#define L(BD, P) mc_one_kernel<BD, P><<<grid, MC_WARPS * 32, 0, hc.stream()>>>(s.ref, s.ox, s.oy, w, h, mx, my, filter2d, out, pitch, bdmax)
        if (hbd) { if (prep) L(BD16, true); else L(BD16, false); } else { if (prep) L(BD8, true); else L(BD8, false); }
#undef L
    }
    void *stage = nullptr;
    if (prep) stage = hc.down(out, (size_t)w * h * 2); else hc.rect_down(drect);
    if (hc.sync()) return hc.err;
    if (prep) memcpy(dst, stage, (size_t)w * h * 2); else drect.finish(dst);
    return 0;
}

}  // namespace

extern "C" int rb200_mc(int filter2d, void *dst, ptrdiff_t ds, const void *src, ptrdiff_t ss, int w, int h, int mx,
                        int my, int bdmax) {
    return mc_common(false, false, filter2d, dst, ds, src, ss, w, h, mx, my, 0, 0, bdmax);
}
extern "C" int rb200_mct(int filter2d, int16_t *tmp, const void *src, ptrdiff_t ss, int w, int h, int mx, int my,
                         int bdmax) {
    return mc_common(true, false, filter2d, tmp, 0, src, ss, w, h, mx, my, 0, 0, bdmax);
}
extern "C" int rb200_mc_scaled(int filter2d, void *dst, ptrdiff_t ds, const void *src, ptrdiff_t ss, int w, int h,
                               int mx, int my, int dx, int dy, int bdmax) {
    return mc_common(false, true, filter2d, dst, ds, src, ss, w, h, mx, my, dx, dy, bdmax);
}
extern "C" int rb200_mct_scaled(int filter2d, int16_t *tmp, const void *src, ptrdiff_t ss, int w, int h, int mx,
                                int my, int dx, int dy, int bdmax) {
    return mc_common(true, true, filter2d, tmp, 0, src, ss, w, h, mx, my, dx, dy, bdmax);
}

static int compound_common(int mode, void *dst, ptrdiff_t ds, const int16_t *t1, const int16_t *t2, int w, int h,
                           int weight, const uint8_t *mask, int bdmax) {
    if (!dst || !t1 || !t2 || w < 1 || h < 1 || w > 128 || h > 128) return set_error(-22, "compound: bad argument");
    const size_t px = pxsz(bdmax), n = (size_t)w * h;
    HostCall hc(2 * (DevRect::bytes_for(w * px, h) + n * 5));
    DevRect drect;
    if (hc.rect_up(drect, dst, ds, w * px, h)) return hc.err;
    const int16_t *d1 = (const int16_t *)hc.up(t1, n * 2), *d2 = (const int16_t *)hc.up(t2, n * 2);
    const uint8_t *dm = mask ? (const uint8_t *)hc.up(mask, n) : nullptr;
    if (hc.err) return hc.err;
    const int grid = (int)((n + 127) / 128);
    if (bdmax > 255) compound_kernel<BD16><<<grid, 128, 0, hc.stream()>>>(mode, drect.dptr, drect.dpitch, d1, d2, w, h, weight, dm, bdmax);
    else compound_kernel<BD8><<<grid, 128, 0, hc.stream()>>>(mode, drect.dptr, drect.dpitch, d1, d2, w, h, weight, dm, bdmax);
    hc.rect_down(drect);
    if (hc.sync()) return hc.err;
    drect.finish(dst);
    return 0;
}
extern "C" int rb200_avg(void *dst, ptrdiff_t ds, const int16_t *t1, const int16_t *t2, int w, int h, int bdmax) {
    return compound_common(CMP_AVG, dst, ds, t1, t2, w, h, 0, nullptr, bdmax);
}
extern "C" int rb200_w_avg(void *dst, ptrdiff_t ds, const int16_t *t1, const int16_t *t2, int w, int h, int weight,
                           int bdmax) {
    return compound_common(CMP_WAVG, dst, ds, t1, t2, w, h, weight, nullptr, bdmax);
}
extern "C" int rb200_mask(void *dst, ptrdiff_t ds, const int16_t *t1, const int16_t *t2, int w, int h,
                          const uint8_t *mask, int bdmax) {
    if (!mask) return set_error(-22, "mask: null mask");
    return compound_common(CMP_MASK, dst, ds, t1, t2, w, h, 0, mask, bdmax);
}
extern "C" int rb200_w_mask(int ss, void *dst, ptrdiff_t ds, const int16_t *t1, const int16_t *t2, int w, int h,
                            uint8_t *mask, int sign, int bdmax) {
    if (ss < 0 || ss > 2 || !dst || !t1 || !t2 || !mask || w < 2 || h < 2 || w > 128 || h > 128)
        return set_error(-22, "w_mask: bad argument");
    const int ss_hor = ss > 0, ss_ver = ss == 2;
    const size_t px = pxsz(bdmax), n = (size_t)w * h, nm = (size_t)(w >> ss_hor) * (h >> ss_ver);
    HostCall hc(2 * (DevRect::bytes_for(w * px, h) + n * 5));
    DevRect drect;
    if (hc.rect_up(drect, dst, ds, w * px, h)) return hc.err;
    const int16_t *d1 = (const int16_t *)hc.up(t1, n * 2), *d2 = (const int16_t *)hc.up(t2, n * 2);
    uint8_t *dm = (uint8_t *)hc.dev(nm);
    if (hc.err) return hc.err;
    const int grid = (int)((nm + 127) / 128);
    if (bdmax > 255) w_mask_kernel<BD16><<<grid, 128, 0, hc.stream()>>>(drect.dptr, drect.dpitch, d1, d2, w, h, dm, sign, ss_hor, ss_ver, bdmax);
    else w_mask_kernel<BD8><<<grid, 128, 0, hc.stream()>>>(drect.dptr, drect.dpitch, d1, d2, w, h, dm, sign, ss_hor, ss_ver, bdmax);
    hc.rect_down(drect);
    void *ms = hc.down(dm, nm);
    if (hc.sync()) return hc.err;
    drect.finish(dst);
    memcpy(mask, ms, nm);
    return 0;
}
// dav1d_wedge_masks[bs][layout][sign][idx] as the compound kernel evaluates it (luma in closed form, chroma
// averaged from it as init_chroma, src/wedge.c:149-163)
__global__ void wedge_mask_kernel(uint8_t *out, int w, int h, int ss_hor, int ss_ver, int sign, int idx) {
    const int cw = w >> ss_hor, ch = h >> ss_ver;
    for (int i = threadIdx.x; i < cw * ch; i += blockDim.x) {
        const int y = i / cw, x = i - y * cw;
        int m = wedge_mask_at(w, h, idx, x << ss_hor, y << ss_ver);
        if (ss_hor) {
            m += wedge_mask_at(w, h, idx, (x << 1) + 1, y << ss_ver) + 1;
            if (ss_ver) m += wedge_mask_at(w, h, idx, x << 1, (y << 1) + 1) + wedge_mask_at(w, h, idx, (x << 1) + 1, (y << 1) + 1) + 1;
            m = (m - sign) >> (1 + ss_ver);
        }
        out[i] = (uint8_t)m;
    }
}
extern "C" int rb200_wedge_mask(int w, int h, int ss, int sign, int wedge_idx, uint8_t *mask) {
    auto ok = [](int v) { return v == 8 || v == 16 || v == 32; };
    if (!ok(w) || !ok(h) || ss < 0 || ss > 2 || wedge_idx < 0 || wedge_idx > 15 || !mask)
        return set_error(-22, "wedge_mask: bad argument");
    const int ss_hor = ss > 0, ss_ver = ss == 2;
    const size_t n = (size_t)(w >> ss_hor) * (h >> ss_ver);
    HostCall hc(2 * n + 256);
    uint8_t *d = (uint8_t *)hc.dev(n);
    if (hc.err) return hc.err;
    wedge_mask_kernel<<<1, 256, 0, hc.stream()>>>(d, w, h, ss_hor, ss_ver, sign & 1, wedge_idx);
    void *m = hc.down(d, n);
    if (hc.sync()) return hc.err;
    memcpy(mask, m, n);
    return 0;
}
extern "C" int rb200_blend(int dir, void *dst, ptrdiff_t ds, const void *tmp, int w, int h, const uint8_t *mask,
                           int bdmax) {
    if (dir < 0 || dir > 2 || !dst || !tmp || (dir == 0 && !mask) || w < 1 || h < 1 || w > 128 || h > 128)
        return set_error(-22, "blend: bad argument");
    const size_t px = pxsz(bdmax), n = (size_t)w * h;
    HostCall hc(2 * (DevRect::bytes_for(w * px, h) + n * 3));
    DevRect drect;
    if (hc.rect_up(drect, dst, ds, w * px, h)) return hc.err;
    const void *dt = hc.up(tmp, n * px);
    const uint8_t *dm = dir == 0 ? (const uint8_t *)hc.up(mask, n) : nullptr;
    if (hc.err) return hc.err;
    const int grid = (int)((n + 127) / 128);
    if (bdmax > 255) blend_kernel<BD16><<<grid, 128, 0, hc.stream()>>>(dir, drect.dptr, drect.dpitch, dt, w, h, dm);
    else blend_kernel<BD8><<<grid, 128, 0, hc.stream()>>>(dir, drect.dptr, drect.dpitch, dt, w, h, dm);
    hc.rect_down(drect);
    if (hc.sync()) return hc.err;
    drect.finish(dst);
    return 0;
}

static int warp_common(bool prep, void *dst, ptrdiff_t ds, const void *src, ptrdiff_t ss, const int16_t *abcd, int mx,
                       int my, int bdmax) {
    if (!dst || !src || !abcd) return set_error(-22, "warp8x8: bad argument");
    const size_t px = pxsz(bdmax);
    HostCall hc(2 * (DevRect::bytes_for(15 * px, 15) + 8 * 8 * 2 + 64));
    SrcStage s;
    if (s.setup(hc, src, ss, 8, 8, 3, 4, 3, 4, bdmax)) return hc.err;
    const int16_t *dabcd = (const int16_t *)hc.up(abcd, 8);
    DevRect drect;
    void *out; int64_t pitch;
    if (prep) { out = hc.dev(8 * 8 * 2); pitch = 8; }
    else { if (hc.rect_up(drect, dst, ds, 8 * px, 8)) return hc.err; out = drect.dptr; pitch = drect.dpitch; }
    if (hc.err) return hc.err;
#define L(BD, P) warp8x8_kernel<BD, P><<<1, 64, 0, hc.stream()>>>(s.ref, s.ox, s.oy, out, pitch, dabcd, mx, my, bdmax)
    if (bdmax > 255) { if (prep) L(BD16, true); else L(BD16, false); } else { if (prep) L(BD8, true); else L(BD8, false); }
#undef L
    void *stage = nullptr;
    if (prep) stage = hc.down(out, 128); else hc.rect_down(drect);
    if (hc.sync()) return hc.err;
    if (prep) { for (int y = 0; y < 8; y++) memcpy((int16_t *)dst + (int64_t)y * ds, (int16_t *)stage + y * 8, 16); }
    else drect.finish(dst);
    return 0;
}
extern "C" int rb200_warp8x8(void *dst, ptrdiff_t ds, const void *src, ptrdiff_t ss, const int16_t *abcd, int mx,
                             int my, int bdmax) {
    return warp_common(false, dst, ds, src, ss, abcd, mx, my, bdmax);
}
extern "C" int rb200_warp8x8t(int16_t *tmp, ptrdiff_t tmp_stride, const void *src, ptrdiff_t ss, const int16_t *abcd,
                              int mx, int my, int bdmax) {
    return warp_common(true, tmp, tmp_stride, src, ss, abcd, mx, my, bdmax);
}

extern "C" int rb200_emu_edge(intptr_t bw, intptr_t bh, intptr_t iw, intptr_t ih, intptr_t x, intptr_t y, void *dst,
                              ptrdiff_t ds, const void *ref, ptrdiff_t rs, int bdmax) {
    if (!dst || !ref || bw < 1 || bh < 1 || iw < 1 || ih < 1 || bw > 1024 || bh > 1024)
        return set_error(-22, "emu_edge: bad argument");
    const size_t px = pxsz(bdmax);
    // only the visible part of the reference that the block touches is uploaded
    const int x0 = iclip((int)x, 0, (int)iw - 1), y0 = iclip((int)y, 0, (int)ih - 1);
    const int x1 = iclip((int)(x + bw - 1), 0, (int)iw - 1), y1 = iclip((int)(y + bh - 1), 0, (int)ih - 1);
    const int cw = x1 - x0 + 1, ch = y1 - y0 + 1;
    HostCall hc(2 * (DevRect::bytes_for(cw * px, ch) + DevRect::bytes_for(bw * px, (int)bh)));
    DevRect srect, drect;
    if (hc.rect_up(srect, (const uint8_t *)ref + (int64_t)y0 * rs + (int64_t)x0 * (int64_t)px, rs, cw * px, ch)) return hc.err;
    if (hc.rect_up(drect, dst, ds, bw * px, (int)bh)) return hc.err;
    McRef r; r.base = srect.dptr; r.stride = srect.dpitch; r.w = cw; r.h = ch;
    dim3 grid(((int)bw + 63) / 64, (int)bh);
    if (bdmax > 255) emu_edge_kernel<BD16><<<grid, 64, 0, hc.stream()>>>(r, (int)x - x0, (int)y - y0, (int)bw, (int)bh, drect.dptr, drect.dpitch);
    else emu_edge_kernel<BD8><<<grid, 64, 0, hc.stream()>>>(r, (int)x - x0, (int)y - y0, (int)bw, (int)bh, drect.dptr, drect.dpitch);
    hc.rect_down(drect);
    if (hc.sync()) return hc.err;
    drect.finish(dst);
    return 0;
}

extern "C" int rb200_resize(void *dst, ptrdiff_t ds, const void *src, ptrdiff_t ss, int dst_w, int h, int src_w,
                            int dx, int mx, int bdmax) {
    if (!dst || !src || dst_w < 1 || h < 1 || src_w < 1) return set_error(-22, "resize: bad argument");
    const size_t px = pxsz(bdmax);
    HostCall hc(2 * (DevRect::bytes_for(src_w * px, h) + DevRect::bytes_for(dst_w * px, h)));
    DevRect srect, drect;
    if (hc.rect_up(srect, src, ss, src_w * px, h)) return hc.err;
    if (hc.rect_up(drect, dst, ds, dst_w * px, h)) return hc.err;
    dim3 grid((dst_w + 127) / 128, h);
    if (bdmax > 255) resize_kernel<BD16><<<grid, 128, 0, hc.stream()>>>(drect.dptr, drect.dpitch, srect.dptr, srect.dpitch, dst_w, h, src_w, dx, mx, bdmax);
    else resize_kernel<BD8><<<grid, 128, 0, hc.stream()>>>(drect.dptr, drect.dpitch, srect.dptr, srect.dpitch, dst_w, h, src_w, dx, mx, bdmax);
    hc.rect_down(drect);
    if (hc.sync()) return hc.err;
    drect.finish(dst);
    return 0;
}

// ---- function-pointer table (drop-in for rav1d_mc_dsp_init, src/mc.rs:2495-2566) ----
namespace {
#define FATAL_IF(x, name) do { if (x) rb200_report_fatal(name); } while (0)
template <int F> void mc_slot(void *d, ptrdiff_t ds, const void *s, ptrdiff_t ss, int w, int h, int mx, int my, int bd) { FATAL_IF(rb200_mc(F, d, ds, s, ss, w, h, mx, my, bd), "mc"); }
template <int F> void mct_slot(int16_t *t, const void *s, ptrdiff_t ss, int w, int h, int mx, int my, int bd) { FATAL_IF(rb200_mct(F, t, s, ss, w, h, mx, my, bd), "mct"); }
template <int F> void mcs_slot(void *d, ptrdiff_t ds, const void *s, ptrdiff_t ss, int w, int h, int mx, int my, int dx, int dy, int bd) { FATAL_IF(rb200_mc_scaled(F, d, ds, s, ss, w, h, mx, my, dx, dy, bd), "mc_scaled"); }
template <int F> void mcts_slot(int16_t *t, const void *s, ptrdiff_t ss, int w, int h, int mx, int my, int dx, int dy, int bd) { FATAL_IF(rb200_mct_scaled(F, t, s, ss, w, h, mx, my, dx, dy, bd), "mct_scaled"); }
void avg_slot(void *d, ptrdiff_t ds, const int16_t *a, const int16_t *b, int w, int h, int bd) { FATAL_IF(rb200_avg(d, ds, a, b, w, h, bd), "avg"); }
void w_avg_slot(void *d, ptrdiff_t ds, const int16_t *a, const int16_t *b, int w, int h, int wt, int bd) { FATAL_IF(rb200_w_avg(d, ds, a, b, w, h, wt, bd), "w_avg"); }
void mask_slot(void *d, ptrdiff_t ds, const int16_t *a, const int16_t *b, int w, int h, const uint8_t *m, int bd) { FATAL_IF(rb200_mask(d, ds, a, b, w, h, m, bd), "mask"); }
template <int SS> void w_mask_slot(void *d, ptrdiff_t ds, const int16_t *a, const int16_t *b, int w, int h, uint8_t *m, int sign, int bd) { FATAL_IF(rb200_w_mask(SS, d, ds, a, b, w, h, m, sign, bd), "w_mask"); }
template <int BDMAX> void blend_slot(void *d, ptrdiff_t ds, const void *t, int w, int h, const uint8_t *m) { FATAL_IF(rb200_blend(0, d, ds, t, w, h, m, BDMAX), "blend"); }
template <int BDMAX, int DIR> void blend_dir_slot(void *d, ptrdiff_t ds, const void *t, int w, int h) { FATAL_IF(rb200_blend(DIR, d, ds, t, w, h, nullptr, BDMAX), "blend_dir"); }
void warp_slot(void *d, ptrdiff_t ds, const void *s, ptrdiff_t ss, const int16_t *abcd, int mx, int my, int bd) { FATAL_IF(rb200_warp8x8(d, ds, s, ss, abcd, mx, my, bd), "warp8x8"); }
void warpt_slot(int16_t *t, ptrdiff_t ts, const void *s, ptrdiff_t ss, const int16_t *abcd, int mx, int my, int bd) { FATAL_IF(rb200_warp8x8t(t, ts, s, ss, abcd, mx, my, bd), "warp8x8t"); }
template <int BDMAX> void emu_slot(intptr_t bw, intptr_t bh, intptr_t iw, intptr_t ih, intptr_t x, intptr_t y, void *d, ptrdiff_t ds, const void *r, ptrdiff_t rs) { FATAL_IF(rb200_emu_edge(bw, bh, iw, ih, x, y, d, ds, r, rs, BDMAX), "emu_edge"); }
void resize_slot(void *d, ptrdiff_t ds, const void *s, ptrdiff_t ss, int dw, int h, int sw, int dx, int mx, int bd) { FATAL_IF(rb200_resize(d, ds, s, ss, dw, h, sw, dx, mx, bd), "resize"); }

template <int... F>
void fill_mc(Rb200MCDSPContext *c, std::integer_sequence<int, F...>) {
    ((c->mc[F] = &mc_slot<F>, c->mct[F] = &mct_slot<F>, c->mc_scaled[F] = &mcs_slot<F>, c->mct_scaled[F] = &mcts_slot<F>), ...);
}
}  // namespace

extern "C" void rb200_mc_dsp_init(Rb200MCDSPContext *c, int bpc) {
    fill_mc(c, std::make_integer_sequence<int, RB200_N_2D_FILTERS>{});
    c->avg = &avg_slot; c->w_avg = &w_avg_slot; c->mask = &mask_slot;
    c->w_mask[0] = &w_mask_slot<0>; c->w_mask[1] = &w_mask_slot<1>; c->w_mask[2] = &w_mask_slot<2>;
    if (bpc > 8) {
        c->blend = &blend_slot<1023>; c->blend_v = &blend_dir_slot<1023, 1>; c->blend_h = &blend_dir_slot<1023, 2>;
        c->emu_edge = &emu_slot<1023>;
    } else {
        c->blend = &blend_slot<255>; c->blend_v = &blend_dir_slot<255, 1>; c->blend_h = &blend_dir_slot<255, 2>;
        c->emu_edge = &emu_slot<255>;
    }
    c->warp8x8 = &warp_slot; c->warp8x8t = &warpt_slot; c->resize = &resize_slot;
}
