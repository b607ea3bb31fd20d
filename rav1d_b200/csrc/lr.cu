// Loop restoration (Wiener and self-guided filters) for sm_100a.
//
// Replaces Rav1dLoopRestorationDSPContext {wiener[2], sgr[3]} (src/looprestoration.rs:91-107;
// wiener_rust :299, selfguided_filter :566, sgr_5x5/3x3/mix :710,785,855 ==
// src/looprestoration_tmpl.c:41-520) and, at frame level, the per-sbrow driver
// rav1d_lr_sbrow / lr_sbrow / lr_stripe (src/lr_apply.rs:28-329 == src/lr_apply_tmpl.c:36-202)
// together with the stripe-boundary row backup rav1d_copy_lpf (src/lf_apply.rs:24-226).
//
// Frame kernel: one CTA per (32-pixel column tile) x (64-row stripe, offset by 8
// luma rows as in the reference).  The tile plus a 3-pixel halo is staged in shared
// memory through one sampling rule that states which version of each neighbour
// is read: CDEF output inside the stripe, *deblocked pre-CDEF* rows
// (two real rows, the outer one duplicated) above / below it, replicated frame
// edges, columns clamped to the picture.  The CPU needs lr_line_buf / left[]
// backups for that because it filters in place; here the stage is out of place
// and the deblocked plane is simply still alive.
#include "common.cuh"
#include "tables.cuh"
#include "tma.cuh"
#include "stages.cuh"

namespace rb200 {

constexpr int LR_TW = 32;             // tile width
constexpr int LR_TH = 64;             // max stripe height
constexpr int LR_WP = LR_TW + 6;      // window pitch (38)
constexpr int LR_WROWS = LR_TH + 6;   // 70
constexpr int LR_AP = LR_TW + 2;      // A/B pitch (34)
constexpr int LR_AROWS = LR_TH + 2;   // 66

struct LrSmem {
    uint16_t win[LR_WROWS * LR_WP];   // padded source window
    union {
        uint16_t hor[LR_WROWS * LR_TW];                                 // Wiener intermediate
        struct { int32_t A[LR_AROWS * LR_AP]; int32_t B[LR_AROWS * LR_AP]; } s;  // SGR
    } u;
};

struct LrUnit {
    int kind;          // 0 none, 1 wiener, 2 sgr 5x5, 3 sgr 3x3, 4 sgr mix
    int16_t fh[8], fv[8];
    uint32_t s0, s1;
    int w0, w1;
};

// lr_stripe's parameter packing, src/lr_apply.rs:59-90 == src/lr_apply_tmpl.c:55-85
__device__ __forceinline__ LrUnit lr_unpack(const Rb200Av1RestorationUnit &u) {
    LrUnit r;
    r.kind = 0; r.s0 = r.s1 = 0; r.w0 = r.w1 = 0;
    if (u.type == RB200_RESTORATION_NONE) return r;
    if (u.type == RB200_RESTORATION_WIENER) {
        r.kind = 1;
        r.fh[0] = r.fh[6] = u.filter_h[0]; r.fh[1] = r.fh[5] = u.filter_h[1]; r.fh[2] = r.fh[4] = u.filter_h[2];
        r.fh[3] = (int16_t)(-(u.filter_h[0] + u.filter_h[1] + u.filter_h[2]) * 2 + 128);  // +128 folded in for every bpc
        r.fv[0] = r.fv[6] = u.filter_v[0]; r.fv[1] = r.fv[5] = u.filter_v[1]; r.fv[2] = r.fv[4] = u.filter_v[2];
        r.fv[3] = (int16_t)(128 - (u.filter_v[0] + u.filter_v[1] + u.filter_v[2]) * 2);
        r.fh[7] = r.fv[7] = 0;
        return r;
    }
    const int idx = u.type - RB200_RESTORATION_SGRPROJ;
    r.s0 = tab::k_sgr_params[idx * 2]; r.s1 = tab::k_sgr_params[idx * 2 + 1];
    r.w0 = u.sgr_weights[0];
    r.w1 = 128 - (u.sgr_weights[0] + u.sgr_weights[1]);
    r.kind = 1 + (!!r.s0 + !!r.s1 * 2);  // 2: 5x5 only, 3: 3x3 only, 4: mix
    return r;
}

// ---- Wiener on a staged window.  out(r, c, v) stores pixel v.   src/looprestoration_tmpl.c:139-195
template <typename BD, typename Out>
__device__ void lr_wiener_tile(LrSmem &sm, int tw, int th, const int16_t *fh, const int16_t *fv, int bdmax, Out out) {
    const int bitdepth = BD::hbd ? bpc_from_max(bdmax) : 8;
    const int rbh = 3 + (bitdepth == 12) * 2, rbv = 11 - (bitdepth == 12) * 2;
    const int clip_limit = 1 << (bitdepth + 1 + 7 - rbh);
    int FH[7], FV[7];
#pragma unroll
    for (int k = 0; k < 7; k++) { FH[k] = fh[k]; FV[k] = fv[k]; }
    for (int i = threadIdx.x; i < (th + 6) * tw; i += blockDim.x) {
        const int r = i / tw, c = i - r * tw;
        const uint16_t *s = sm.win + r * LR_WP + c;
        int sum = 1 << (bitdepth + 6);
#pragma unroll
        for (int k = 0; k < 7; k++) sum += (int)s[k] * FH[k];
        sm.u.hor[r * LR_TW + c] = (uint16_t)iclip((sum + (1 << (rbh - 1))) >> rbh, 0, clip_limit - 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < th * tw; i += blockDim.x) {
        const int r = i / tw, c = i - r * tw;
        int sum = -(1 << (bitdepth + (rbv - 1)));
#pragma unroll
        for (int k = 0; k < 7; k++) sum += (int)sm.u.hor[(r + k) * LR_TW + c] * FV[k];
        out(r, c, iclip((sum + (1 << (rbv - 1))) >> rbv, 0, bdmax));
    }
}

// ---- self-guided filter: fills A/B for box size n (25 or 9) with strength s.
// A[j][i], B[j][i] for i in [-1, tw], j in [-1, th]  (stored at [(j+1)*LR_AP + i+1])
template <typename BD>
__device__ void sgr_ab(LrSmem &sm, int tw, int th, int n, unsigned s, int bdmin8) {
    const unsigned one_by_x = n == 25 ? 164 : 455;
    const int rad = n == 25 ? 2 : 1, step = n == 25 ? 2 : 1;
    const int cols = tw + 2, rows = (th + 2 + step - 1) / step;
    for (int t = threadIdx.x; t < rows * cols; t += blockDim.x) {
        const int jr = t / cols, i = t - jr * cols - 1;
        const int j = jr * step - 1;
        // box centred on pixel (i, j): window coordinates (j + 3, i + 3)
        int sum = 0, sumsq = 0;
        for (int dy = -rad; dy <= rad; dy++) {
            const uint16_t *row = sm.win + (j + 3 + dy) * LR_WP + (i + 3);
            for (int dx = -rad; dx <= rad; dx++) { const int v = row[dx]; sum += v; sumsq += v * v; }
        }
        const int a = (sumsq + ((1 << (2 * bdmin8)) >> 1)) >> (2 * bdmin8);
        const int b = (sum + ((1 << bdmin8) >> 1)) >> bdmin8;
        const unsigned p = (unsigned)imax(a * n - b * b, 0);
        const unsigned z = (p * s + (1u << 19)) >> 20;
        const unsigned x = tab::k_sgr_x_by_x[z < 255 ? z : 255];
        sm.u.s.A[(j + 1) * LR_AP + i + 1] = (int32_t)((x * (unsigned)sum * one_by_x + (1u << 11)) >> 12);
        sm.u.s.B[(j + 1) * LR_AP + i + 1] = (int32_t)x;
    }
}

// value of the filter output `dst[j][i]` (coef, truncated to int16 at 8 bpc)
template <typename BD>
__device__ __forceinline__ int sgr_out(const LrSmem &sm, int n, int j, int i) {
    const int32_t *A = sm.u.s.A + (j + 1) * LR_AP + i + 1, *B = sm.u.s.B + (j + 1) * LR_AP + i + 1;
    const int src = sm.win[(j + 3) * LR_WP + i + 3];
    int v;
    if (n == 25) {
        if (!(j & 1)) {
            const int a = (B[-LR_AP] + B[LR_AP]) * 6 + (B[-1 - LR_AP] + B[-1 + LR_AP] + B[1 - LR_AP] + B[1 + LR_AP]) * 5;
            const int b = (A[-LR_AP] + A[LR_AP]) * 6 + (A[-1 - LR_AP] + A[-1 + LR_AP] + A[1 - LR_AP] + A[1 + LR_AP]) * 5;
            v = (b - a * src + (1 << 8)) >> 9;
        } else {
            const int a = B[0] * 6 + (B[-1] + B[1]) * 5;
            const int b = A[0] * 6 + (A[-1] + A[1]) * 5;
            v = (b - a * src + (1 << 7)) >> 8;
        }
    } else {
        const int a = (B[0] + B[-1] + B[1] + B[-LR_AP] + B[LR_AP]) * 4 +
                      (B[-1 - LR_AP] + B[-1 + LR_AP] + B[1 - LR_AP] + B[1 + LR_AP]) * 3;
        const int b = (A[0] + A[-1] + A[1] + A[-LR_AP] + A[LR_AP]) * 4 +
                      (A[-1 - LR_AP] + A[-1 + LR_AP] + A[1 - LR_AP] + A[1 + LR_AP]) * 3;
        v = (b - a * src + (1 << 8)) >> 9;
    }
    return BD::hbd ? v : (int)(int16_t)v;
}

constexpr int LR_PX_PER_THREAD = LR_TW * LR_TH / 256;  // 8

// kind: 2 5x5, 3 3x3, 4 mix.  src/looprestoration_tmpl.c:446-520
template <typename BD, typename Out>
__device__ void lr_sgr_tile(LrSmem &sm, int tw, int th, int kind, unsigned s0, unsigned s1, int w0, int w1, int bdmax,
                            Out out) {
    const int bdmin8 = BD::hbd ? bpc_from_max(bdmax) - 8 : 0;
    int acc[LR_PX_PER_THREAD];
#pragma unroll
    for (int k = 0; k < LR_PX_PER_THREAD; k++) acc[k] = 0;
    if (kind != 3) {
        sgr_ab<BD>(sm, tw, th, 25, s0, bdmin8);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < LR_PX_PER_THREAD; k++) {
            const int i = threadIdx.x + k * 256;
            const int r = i / LR_TW, c = i - r * LR_TW;
            if (r < th && c < tw) acc[k] = w0 * sgr_out<BD>(sm, 25, r, c);
        }
        __syncthreads();
    }
    if (kind != 2) {
        sgr_ab<BD>(sm, tw, th, 9, s1, bdmin8);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < LR_PX_PER_THREAD; k++) {
            const int i = threadIdx.x + k * 256;
            const int r = i / LR_TW, c = i - r * LR_TW;
            if (r < th && c < tw) acc[k] += w1 * sgr_out<BD>(sm, 9, r, c);
        }
    }
#pragma unroll
    for (int k = 0; k < LR_PX_PER_THREAD; k++) {
        const int i = threadIdx.x + k * 256;
        const int r = i / LR_TW, c = i - r * LR_TW;
        if (r < th && c < tw) {
            const int px = sm.win[(r + 3) * LR_WP + c + 3];
            out(r, c, iclip(px + ((acc[k] + (1 << 10)) >> 11), 0, bdmax));
        }
    }
}


// ---------------------------------------------------------------- frame level
// Layout of the frame kernel's shared memory.  Window element e of a row holds picture column x0 - 8 + e (48 columns:
// the box the copy engine delivers starts on a 16-byte boundary of the plane; the filters reach 3 columns out, i.e.
// elements 5 .. 42 are used) and window row r holds picture row top - 3 + r.  Output column i therefore reads its 7 Wiener
// taps at e = i + 5 .. i + 11; four outputs are served from the aligned 12 elements 4g + 4 .. 4g + 15.  A/B column c
// holds box column i = c - 1.
constexpr int L2_WP = 48;                 // window pitch, pixels
constexpr int L2_X0 = 8;                  // element of picture column x0
constexpr int L2_WROWS = LR_TH + 6;       // 70
constexpr int L2_AP = 36;                 // A/B pitch, ints
constexpr int L2_AROWS = LR_TH + 2;       // 66

struct Lr2Smem {
    // rows 3 .. of the window are TMA destinations (8-row boxes of 768 bytes): row 3 must sit on a 128-byte boundary,
    // so the window starts 96 bytes into the aligned block (3 rows x 96 bytes + 96 = 384)
    __align__(128) uint16_t pad_[48];
    uint16_t win[L2_WROWS * L2_WP];
    __align__(128) uint16_t above[2 * L2_WP + 32];   // deblocked rows top - 2, top - 1 (box of 2 rows; 256 bytes)
    __align__(128) uint16_t below[2 * L2_WP + 32];   // deblocked rows bot, bot + 1
    union {
        __align__(16) uint16_t hor[L2_WROWS * LR_TW];
        struct { __align__(16) int32_t A[L2_AROWS * L2_AP]; __align__(16) int32_t B[L2_AROWS * L2_AP]; } s;
    } u;
};

// v[0 .. N-1] = elements 1 .. N of the 12 starting at the 8-byte aligned address p (N <= 10): the window's columns sit
// one element past the alignment (L2_X0 - 3 = 5)
template <int N>
__device__ __forceinline__ void lr2_unpack_odd(const uint16_t *p, int *v) {
    const uint2 a = *(const uint2 *)p, b = *(const uint2 *)(p + 4);
    v[0] = a.x >> 16; v[1] = a.y & 0xffff; v[2] = a.y >> 16;
    v[3] = b.x & 0xffff; v[4] = b.x >> 16; v[5] = b.y & 0xffff;
    if (N > 6) v[6] = b.y >> 16;
    if (N > 7) {
        const uint2 c = *(const uint2 *)(p + 8);
        v[7] = c.x & 0xffff;
        if (N > 8) v[8] = c.x >> 16;
        if (N > 9) v[9] = c.y & 0xffff;
    }
}

template <typename BD>
__device__ __forceinline__ void lr2_store4(typename BD::pixel *o, const int *v, int n) {
    if (n >= 4) {
        if (BD::hbd) *(uint2 *)o = make_uint2(v[0] | (v[1] << 16), v[2] | (v[3] << 16));
        else *(unsigned *)o = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
    } else {
        for (int i = 0; i < n; i++) o[i] = (typename BD::pixel)v[i];
    }
}

// Wiener: each thread produces 4 horizontally adjacent samples per pass from registers.
template <typename BD>
__device__ void lr2_wiener(Lr2Smem &sm, int tw, int th, const int16_t *fh, const int16_t *fv, int bdmax,
                           typename BD::pixel *o, int64_t ps) {
    const int bitdepth = BD::hbd ? bpc_from_max(bdmax) : 8;
    const int rbh = 3 + (bitdepth == 12) * 2, rbv = 11 - (bitdepth == 12) * 2;
    const int clip_limit = 1 << (bitdepth + 1 + 7 - rbh);
    int F[7];
#pragma unroll
    for (int k = 0; k < 7; k++) F[k] = fh[k];
    for (int t = threadIdx.x; t < (th + 6) * 8; t += 256) {
        const int g = t & 7, r = t >> 3;
        int px[10];      // picture columns x0 + 4g - 3 .. x0 + 4g + 6
        lr2_unpack_odd<10>(sm.win + r * L2_WP + 4 * g + 4, px);
        unsigned outv[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int sum = 1 << (bitdepth + 6);
#pragma unroll
            for (int k = 0; k < 7; k++) sum += px[j + k] * F[k];
            outv[j] = (unsigned)iclip((sum + (1 << (rbh - 1))) >> rbh, 0, clip_limit - 1);
        }
        *(uint2 *)(sm.u.hor + r * LR_TW + 4 * g) = make_uint2(outv[0] | (outv[1] << 16), outv[2] | (outv[3] << 16));
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 7; k++) F[k] = fv[k];
    {
        const int g = threadIdx.x & 7, rp = threadIdx.x >> 3;   // 8 column groups x 32 row pairs
        const int r0 = 2 * rp;
        if (r0 < th && 4 * g < tw) {
            int hv[8][4];
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const uint2 q = *(const uint2 *)(sm.u.hor + (r0 + r) * LR_TW + 4 * g);
                hv[r][0] = q.x & 0xffff; hv[r][1] = q.x >> 16; hv[r][2] = q.y & 0xffff; hv[r][3] = q.y >> 16;
            }
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                if (r0 + rr >= th) break;
                int outv[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    int sum = -(1 << (bitdepth + (rbv - 1)));
#pragma unroll
                    for (int k = 0; k < 7; k++) sum += hv[rr + k][j] * F[k];
                    outv[j] = iclip((sum + (1 << (rbv - 1))) >> rbv, 0, bdmax);
                }
                lr2_store4<BD>(o + (int64_t)(r0 + rr) * ps + 4 * g, outv, tw - 4 * g);
            }
        }
    }
}

// Box sums + the a/b -> (x, x * sum * one_by_x) transform for four horizontally adjacent boxes.
// N = 9: every row j in [-1, th]; N = 25: rows j = -1, 1, 3, ...   src/looprestoration_tmpl.c:282-373
template <typename BD, int N>
__device__ void lr2_sgr_ab(Lr2Smem &sm, int th, unsigned s, int bdmin8) {
    constexpr int RAD = N == 25 ? 2 : 1, STEP = N == 25 ? 2 : 1;
    constexpr unsigned one_by_x = N == 25 ? 164 : 455;
    const int rows = (th + 2 + STEP - 1) / STEP;
    for (int t = threadIdx.x; t < rows * 9; t += 256) {
        const int jr = t / 9, k = t - jr * 9;
        const int j = jr * STEP - 1;
        int cs[8], cq[8];
#pragma unroll
        for (int c = 0; c < 8; c++) { cs[c] = 0; cq[c] = 0; }
#pragma unroll
        for (int dy = -RAD; dy <= RAD; dy++) {
            int px[8];   // picture columns x0 + 4k - 3 .. x0 + 4k + 4
            lr2_unpack_odd<8>(sm.win + (j + 3 + dy) * L2_WP + 4 * k + 4, px);
#pragma unroll
            for (int c = (N == 25 ? 0 : 1); c < (N == 25 ? 8 : 7); c++) { cs[c] += px[c]; cq[c] += px[c] * px[c]; }
        }
        int av[4], bv[4];
#pragma unroll
        for (int m = 0; m < 4; m++) {
            int sum, sumsq;
            if (N == 25) { sum = cs[m] + cs[m + 1] + cs[m + 2] + cs[m + 3] + cs[m + 4]; sumsq = cq[m] + cq[m + 1] + cq[m + 2] + cq[m + 3] + cq[m + 4]; }
            else { sum = cs[m + 1] + cs[m + 2] + cs[m + 3]; sumsq = cq[m + 1] + cq[m + 2] + cq[m + 3]; }
            const int a = (sumsq + ((1 << (2 * bdmin8)) >> 1)) >> (2 * bdmin8);
            const int b = (sum + ((1 << bdmin8) >> 1)) >> bdmin8;
            const unsigned p = (unsigned)imax(a * N - b * b, 0);
            const unsigned z = (p * s + (1u << 19)) >> 20;
            const unsigned x = tab::k_sgr_x_by_x[z < 255 ? z : 255];
            av[m] = (int)((x * (unsigned)sum * one_by_x + (1u << 11)) >> 12);
            bv[m] = (int)x;
        }
        *(int4 *)(sm.u.s.A + (j + 1) * L2_AP + 4 * k) = make_int4(av[0], av[1], av[2], av[3]);
        *(int4 *)(sm.u.s.B + (j + 1) * L2_AP + 4 * k) = make_int4(bv[0], bv[1], bv[2], bv[3]);
    }
}

// acc[m] += wgt * (weighted-neighbourhood output) for the 4 pixels (j, 4g .. 4g+3).
template <typename BD, int N>
__device__ __forceinline__ void lr2_sgr_out4(const Lr2Smem &sm, int j, int g, int wgt, const int *src, int *acc) {
    auto row6 = [&](const int32_t *base, int row, int *v) {   // columns i = 4g-1 .. 4g+4 of A/B row `row`
        const int32_t *p = base + (row + 1) * L2_AP + 4 * g;
        const int4 q = *(const int4 *)p; const int2 q2 = *(const int2 *)(p + 4);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; v[4] = q2.x; v[5] = q2.y;
    };
    int a[4], b[4], sh, rnd;
    if (N == 9) {
        int au[6], am[6], ad[6];
        row6(sm.u.s.B, j - 1, au); row6(sm.u.s.B, j, am); row6(sm.u.s.B, j + 1, ad);
#pragma unroll
        for (int c = 0; c < 6; c++) au[c] += am[c] + ad[c];
#pragma unroll
        for (int m = 0; m < 4; m++) a[m] = 3 * (au[m] + au[m + 1] + au[m + 2]) + am[m] + au[m + 1] + am[m + 2];
        row6(sm.u.s.A, j - 1, au); row6(sm.u.s.A, j, am); row6(sm.u.s.A, j + 1, ad);
#pragma unroll
        for (int c = 0; c < 6; c++) au[c] += am[c] + ad[c];
#pragma unroll
        for (int m = 0; m < 4; m++) b[m] = 3 * (au[m] + au[m + 1] + au[m + 2]) + am[m] + au[m + 1] + am[m + 2];
        sh = 9; rnd = 1 << 8;
    } else if (!(j & 1)) {
        int u[6], d[6];
        row6(sm.u.s.B, j - 1, u); row6(sm.u.s.B, j + 1, d);
#pragma unroll
        for (int c = 0; c < 6; c++) u[c] += d[c];
#pragma unroll
        for (int m = 0; m < 4; m++) a[m] = u[m + 1] * 6 + (u[m] + u[m + 2]) * 5;
        row6(sm.u.s.A, j - 1, u); row6(sm.u.s.A, j + 1, d);
#pragma unroll
        for (int c = 0; c < 6; c++) u[c] += d[c];
#pragma unroll
        for (int m = 0; m < 4; m++) b[m] = u[m + 1] * 6 + (u[m] + u[m + 2]) * 5;
        sh = 9; rnd = 1 << 8;
    } else {
        int u[6];
        row6(sm.u.s.B, j, u);
#pragma unroll
        for (int m = 0; m < 4; m++) a[m] = u[m + 1] * 6 + (u[m] + u[m + 2]) * 5;
        row6(sm.u.s.A, j, u);
#pragma unroll
        for (int m = 0; m < 4; m++) b[m] = u[m + 1] * 6 + (u[m] + u[m + 2]) * 5;
        sh = 8; rnd = 1 << 7;
    }
#pragma unroll
    for (int m = 0; m < 4; m++) {
        const int v = (b[m] - a[m] * src[m] + rnd) >> sh;
        acc[m] += wgt * (BD::hbd ? v : (int)(int16_t)v);
    }
}

// kind: 2 5x5, 3 3x3, 4 mix.  src/looprestoration_tmpl.c:446-520
template <typename BD>
__device__ void lr2_sgr(Lr2Smem &sm, int tw, int th, int kind, unsigned s0, unsigned s1, int w0, int w1, int bdmax,
                        typename BD::pixel *o, int64_t ps) {
    const int bdmin8 = BD::hbd ? bpc_from_max(bdmax) - 8 : 0;
    const int g = threadIdx.x & 7, r_lo = threadIdx.x >> 3;   // rows r_lo and r_lo + 32
    int acc[2][4], src[2][4];
#pragma unroll
    for (int it = 0; it < 2; it++) {
        const int r = r_lo + 32 * it;
#pragma unroll
        for (int m = 0; m < 4; m++) { acc[it][m] = 0; src[it][m] = sm.win[(r + 3) * L2_WP + L2_X0 + 4 * g + m]; }
    }
    if (kind != 3) {
        lr2_sgr_ab<BD, 25>(sm, th, s0, bdmin8);
        __syncthreads();
#pragma unroll
        for (int it = 0; it < 2; it++) {
            const int r = r_lo + 32 * it;
            if (r < th && 4 * g < tw) lr2_sgr_out4<BD, 25>(sm, r, g, w0, src[it], acc[it]);
        }
        __syncthreads();
    }
    if (kind != 2) {
        lr2_sgr_ab<BD, 9>(sm, th, s1, bdmin8);
        __syncthreads();
#pragma unroll
        for (int it = 0; it < 2; it++) {
            const int r = r_lo + 32 * it;
            if (r < th && 4 * g < tw) lr2_sgr_out4<BD, 9>(sm, r, g, w1, src[it], acc[it]);
        }
    }
#pragma unroll
    for (int it = 0; it < 2; it++) {
        const int r = r_lo + 32 * it;
        if (r < th && 4 * g < tw) {
            int outv[4];
#pragma unroll
            for (int m = 0; m < 4; m++) outv[m] = iclip(src[it][m] + ((acc[it][m] + (1 << 10)) >> 11), 0, bdmax);
            lr2_store4<BD>(o + (int64_t)r * ps + 4 * g, outv, tw - 4 * g);
        }
    }
}

// cdef: CDEF output (the picture being restored); dbl: deblocked pre-CDEF picture; out: restored picture.
// TMA (16-bit pictures): the window comes from the copy engine -- ceil(th / 8) boxes of 48 x 8 pixels of the CDEF output
// into rows 3 .., one box of 48 x 2 deblocked rows above and one below into side buffers -- requested by one thread;
// what the boxes cannot express (the stripe rule's repeated rows, replicated pixels at the picture's edges: the copy
// engine fills with zeros) is patched by a few threads afterwards.
template <typename BD, bool TMA>
__global__ void __launch_bounds__(256, 6)
lr_frame_kernel(const uint8_t *__restrict__ cdef, const uint8_t *__restrict__ dbl, uint8_t *__restrict__ outp,
                int64_t stride, LrFrameParams P, const Rb200Av1Restoration *__restrict__ lrm, int bdmax,
                const __grid_constant__ CUtensorMap map_main, const __grid_constant__ CUtensorMap map_halo) {
    using pixel = typename BD::pixel;
    __shared__ Lr2Smem sm;
    __shared__ __align__(8) uint64_t bar;
    const int x0 = blockIdx.x * LR_TW, s = P.stripe_first + blockIdx.y;
    const int sh = 64 >> P.ss_ver, off = 8 >> P.ss_ver;
    const int top = imax(0, s * sh - off), bot = imin(P.h, (s + 1) * sh - off);
    const int tw = imin(LR_TW, P.w - x0), th = bot - top;
    if (th <= 0) return;

    // ---- which restoration unit (lr_sbrow, src/lr_apply_tmpl.c:117-160)
    const int unit = 1 << P.unit_log2, half = unit >> 1;
    const int sby = imin(s >> P.sb128, P.sbh - 1);
    const int row_y = sby << (6 - P.ss_ver + P.sb128);
    int aligned = row_y & ~(unit - 1);
    if (aligned && aligned + half > P.h) aligned -= unit;
    aligned <<= P.ss_ver;
    const int sb_idx = (aligned >> 7) * P.sr_sb128w;
    const int unit_idx = ((aligned >> 6) & 1) << 1;
    const int n_units = imax(1, (P.w + half) >> P.unit_log2);
    const int ux = imin(x0 >> P.unit_log2, n_units - 1) << P.unit_log2;  // unit start x
    const int shift_hor = 7 - P.ss_hor;
    const Rb200Av1RestorationUnit ru =
        lrm[sb_idx + (ux >> shift_hor)].lr[P.plane][unit_idx + ((ux >> (shift_hor - 1)) & 1)];
    const LrUnit U = lr_unpack(ru);

    const int64_t ps = stride / (int64_t)sizeof(pixel);
    pixel *o = (pixel *)outp + (int64_t)top * ps + x0;
    if (U.kind == 0) {  // unit not restored: copy through
        const pixel *c = (const pixel *)cdef + (int64_t)top * ps + x0;
        for (int i = threadIdx.x; i < th * 8; i += 256) {
            const int r = i >> 3, g = i & 7;
            if (4 * g >= tw) continue;
            int v[4];
#pragma unroll
            for (int m = 0; m < 4; m++) v[m] = 4 * g + m < tw ? c[(int64_t)r * ps + 4 * g + m] : 0;
            lr2_store4<BD>(o + (int64_t)r * ps + 4 * g, v, tw - 4 * g);
        }
        return;
    }
    // ---- stage the padded window (padding(), src/looprestoration_tmpl.c:41-137, with
    //      lpf rows = deblocked rows saved by backup_lpf, src/lf_apply_tmpl.c:41-106)
    if (TMA) {
        const bool have_above = top > 0, have_below = bot < P.h;
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            mbar_fence_init();
            const int n_main = (th + 7) >> 3;
            mbar_arrive_expect_tx(&bar, (unsigned)(n_main * 8 + (have_above ? 2 : 0) + (have_below ? 2 : 0)) * L2_WP * 2);
            for (int k = 0; k < n_main; k++) tma_load_2d(sm.win + (3 + 8 * k) * L2_WP, &map_main, x0 - L2_X0, top + 8 * k, &bar);
            if (have_above) tma_load_2d(sm.above, &map_halo, x0 - L2_X0, top - 2, &bar);
            if (have_below) tma_load_2d(sm.below, &map_halo, x0 - L2_X0, bot, &bar);
        }
        __syncthreads();
        mbar_wait(&bar, 0);
        // rows above / below the stripe: deblocked rows top - 2, top - 2, top - 1 and bot, bot + 1, bot + 1 (the last one
        // clamped to the picture); at the picture's top / bottom the first / last row of the stripe, three times
        if (threadIdx.x < 6 * (L2_WP / 2)) {
            const int r = threadIdx.x / (L2_WP / 2), k = threadIdx.x - r * (L2_WP / 2);      // 6 rows x 24 words
            const unsigned *src;
            int dst_row;
            if (r < 3) {
                dst_row = r;
                src = have_above ? (const unsigned *)(sm.above + (r == 2 ? L2_WP : 0)) : (const unsigned *)(sm.win + 3 * L2_WP);
            } else {
                dst_row = th + r;
                const bool second = r > 3 && bot + 1 < P.h;
                src = have_below ? (const unsigned *)(sm.below + (second ? L2_WP : 0)) : (const unsigned *)(sm.win + (th + 2) * L2_WP);
            }
            const unsigned v = src[k];
            ((unsigned *)(sm.win + dst_row * L2_WP))[k] = v;
        }
        if (x0 == 0 || x0 + LR_TW + 3 > P.w) {      // replicate the picture's first / last column into the 3 columns beyond it
            __syncthreads();
            const int e_last = P.w - 1 - x0 + L2_X0;            // element of the last picture column (>= L2_X0)
            for (int i = threadIdx.x; i < (th + 6) * 8; i += 256) {
                const int r = i >> 3, j = i & 7;
                uint16_t *row = sm.win + r * L2_WP;
                if (j < 3) { if (x0 == 0) row[L2_X0 - 1 - j] = row[L2_X0]; }
                else if (j < 6 && x0 + LR_TW + 3 > P.w && e_last + (j - 2) < L2_WP) row[e_last + (j - 2)] = row[e_last];
            }
        }
    } else {
        // a warp owns rows r = warp, warp + 8, ...; lanes own columns e and e + 32
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const int xa = iclip(x0 + lane - 3, 0, P.w - 1), xb = iclip(x0 + lane + 29, 0, P.w - 1);
        constexpr int NIT = (L2_WROWS + 7) / 8;
        pixel va[NIT], vb[NIT];
        // all loads of the thread are issued before the first store: the window is fetched with
        // NIT-fold memory-level parallelism instead of one DRAM/L2 round trip per row
#pragma unroll
        for (int it = 0; it < NIT; it++) {
            const int r = warp + 8 * it;
            va[it] = 0; vb[it] = 0;
            if (r < th + 6) {
                const int yy = top + r - 3;
                const pixel *row;
                if (yy < top) row = top == 0 ? (const pixel *)cdef : (const pixel *)dbl + (int64_t)imax(yy, top - 2) * ps;
                else if (yy >= bot) row = bot >= P.h ? (const pixel *)cdef + (int64_t)(bot - 1) * ps
                                                     : (const pixel *)dbl + (int64_t)imin(imin(yy, bot + 1), P.h - 1) * ps;
                else row = (const pixel *)cdef + (int64_t)yy * ps;
                va[it] = row[xa];
                if (lane < 6) vb[it] = row[xb];
            }
        }
#pragma unroll
        for (int it = 0; it < NIT; it++) {
            const int r = warp + 8 * it;
            if (r < th + 6) {
                sm.win[r * L2_WP + L2_X0 - 3 + lane] = va[it];
                if (lane < 8) sm.win[r * L2_WP + L2_X0 + 29 + lane] = vb[it];
            }
        }
    }
    __syncthreads();
    if (U.kind == 1) lr2_wiener<BD>(sm, tw, th, U.fh, U.fv, bdmax, o, ps);
    else lr2_sgr<BD>(sm, tw, th, U.kind, U.s0, U.s1, U.w0, U.w1, bdmax, o, ps);
}

// ---- per-call: window from an explicit padded buffer tmp[(h+6)][pitch], built by lr_pad_kernel
template <typename BD>
__global__ void lr_pad_kernel(uint8_t *tmp, int pitch, const uint8_t *p, int64_t stride, const uint8_t *left,
                              const uint8_t *lpf, int64_t lpf_stride, int w, int h, unsigned edges) {
    using pixel = typename BD::pixel;
    // tmp(r, c), r in [0, h+6), c in [0, w+6) <-> picture (y = r-3, x = c-3); same cases as padding()
    const int have_left = !!(edges & RB200_LR_HAVE_LEFT), have_right = !!(edges & RB200_LR_HAVE_RIGHT);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (h + 6) * (w + 6); i += gridDim.x * blockDim.x) {
        const int r = i / (w + 6), c = i - r * (w + 6);
        int x = c - 3, y = r - 3;
        if (!have_left && x < 0) x = 0;
        if (!have_right && x >= w) x = w - 1;
        pixel v;
        auto P = [&](int yy, int xx) -> pixel {  // picture rows: left[] supplies x < 0
            if (xx < 0) return ((const pixel *)left)[yy * 4 + 4 + xx];
            return ((const pixel *)(p + (int64_t)yy * stride))[xx];
        };
        // lpf is staged with its x = -3 column at offset 0 when have_left
        auto Lp = [&](int row, int xx) -> pixel { return ((const pixel *)(lpf + (int64_t)row * lpf_stride))[xx + 3 * have_left]; };
        if (y < 0) {
            if (edges & RB200_LR_HAVE_TOP) v = Lp(y == -1 ? 1 : 0, x);
            else v = P(0, x);
        } else if (y >= h) {
            if (edges & RB200_LR_HAVE_BOTTOM) v = Lp(y == h ? 2 : 3, x);
            else v = P(h - 1, x);
        } else {
            v = P(y, x);
        }
        ((pixel *)tmp)[r * pitch + c] = v;
    }
}

template <typename BD>
__global__ void __launch_bounds__(256)
lr_call_kernel(const uint8_t *tmp, int pitch, uint8_t *dst, int64_t stride, int w, int h, LrUnit U, int bdmax) {
    using pixel = typename BD::pixel;
    __shared__ LrSmem sm;
    const int x0 = blockIdx.x * LR_TW;
    const int tw = imin(LR_TW, w - x0), th = h;
    for (int i = threadIdx.x; i < (th + 6) * (tw + 6); i += 256) {
        const int r = i / (tw + 6), c = i - r * (tw + 6);
        sm.win[r * LR_WP + c] = ((const pixel *)tmp)[r * pitch + x0 + c];
    }
    __syncthreads();
    const int64_t ps = stride / (int64_t)sizeof(pixel);
    pixel *o = (pixel *)dst + x0;
    auto out = [&](int r, int c, int v) { o[(int64_t)r * ps + c] = (pixel)v; };
    if (U.kind == 1) lr_wiener_tile<BD>(sm, tw, th, U.fh, U.fv, bdmax, out);
    else lr_sgr_tile<BD>(sm, tw, th, U.kind, U.s0, U.s1, U.w0, U.w1, bdmax, out);
}

// P.stripe_first / P.stripe_end: stripes to produce (whole plane: 0, number of stripes; end <= 0 means all)
int lr_plane_launch(const uint8_t *cdef, const uint8_t *dbl, uint8_t *out, int64_t stride, const LrFrameParams &P,
                    const Rb200Av1Restoration *lrm, int bdmax, cudaStream_t st) {
    return lr_plane_launch_tma(cdef, dbl, out, stride, P, lrm, bdmax, st, nullptr, nullptr);
}

// map_main: tensor map of the `cdef` plane (P.w x P.h pixels, boxes of 48 x 8); map_halo: of the `dbl` plane (boxes of
// 48 x 2).  Both null: the window is gathered with loads (8-bit pictures, super-resolution planes).
int lr_encode_maps(CUtensorMap *map_main, CUtensorMap *map_halo, const void *cdef, const void *dbl, int64_t stride, int w, int h) {
    int r = tma_encode_plane(map_main, cdef, 2, w, h, stride, L2_WP, 8);
    if (!r) r = tma_encode_plane(map_halo, dbl, 2, w, h, stride, L2_WP, 2);
    return r;
}

int lr_plane_launch_tma(const uint8_t *cdef, const uint8_t *dbl, uint8_t *out, int64_t stride, const LrFrameParams &P,
                        const Rb200Av1Restoration *lrm, int bdmax, cudaStream_t st, const CUtensorMap *map_main,
                        const CUtensorMap *map_halo) {
    const int sh = 64 >> P.ss_ver, off = 8 >> P.ss_ver;
    const int n_stripes = (P.h + off + sh - 1) / sh;
    const int s1 = P.stripe_end > 0 ? imin(P.stripe_end, n_stripes) : n_stripes;
    if (s1 <= P.stripe_first) return 0;
    dim3 grid((P.w + LR_TW - 1) / LR_TW, s1 - P.stripe_first);
    static const CUtensorMap none = {};
    if (bdmax > 255 && map_main && map_halo)
        lr_frame_kernel<BD16, true><<<grid, 256, 0, st>>>(cdef, dbl, out, stride, P, lrm, bdmax, *map_main, *map_halo);
    else if (bdmax > 255) lr_frame_kernel<BD16, false><<<grid, 256, 0, st>>>(cdef, dbl, out, stride, P, lrm, bdmax, none, none);
    else lr_frame_kernel<BD8, false><<<grid, 256, 0, st>>>(cdef, dbl, out, stride, P, lrm, bdmax, none, none);
    RB_LAUNCH_CHECK();
    return 0;
}

}  // namespace rb200

using namespace rb200;

extern "C" int rb200_lr(int kind, void *dst, ptrdiff_t stride, const void *left, const void *lpf, int w, int h,
                        const Rb200LooprestorationParams *params, uint32_t edges, int bdmax) {
    if (kind < 0 || kind > 4 || !dst || !params || w < 1 || w > 384 || h < 1 || h > 64)
        return set_error(-22, "lr: bad argument");
    const size_t px = bdmax > 255 ? 2 : 1;
    const int hl = !!(edges & RB200_LR_HAVE_LEFT), hr = !!(edges & RB200_LR_HAVE_RIGHT);
    const int cols = w + 3 * hr;
    const int lcols = w + 3 * hl + 3 * hr;
    HostCall hc(2 * (DevRect::bytes_for(cols * px, h) + 4 * DevRect::pitch_for(lcols * px) + (size_t)h * 4 * px) +
                (size_t)(h + 6) * (w + 6) * px + 8192);
    DevRect rect;
    if (hc.rect_up(rect, dst, stride, cols * px, h)) return hc.err;
    // lpf rows the reference may read: 0, 1 (above, with LR_HAVE_TOP) and 6, 7 (below, with LR_HAVE_BOTTOM)
    const size_t lp = DevRect::pitch_for(lcols * px);
    uint8_t *hl_rows = (uint8_t *)hc.st.halloc(4 * lp);
    uint8_t *dl_rows = (uint8_t *)hc.dev(4 * lp);
    if (!hl_rows || hc.err) return hc.err ? hc.err : set_error(-12, "staging arena too small");
    memset(hl_rows, 0, 4 * lp);
    for (int k = 0; k < 4; k++) {
        const bool need = k < 2 ? (edges & RB200_LR_HAVE_TOP) : (edges & RB200_LR_HAVE_BOTTOM);
        if (!need || !lpf) continue;
        const int row = k < 2 ? k : 4 + k;  // 0, 1, 6, 7
        memcpy(hl_rows + k * lp, (const uint8_t *)lpf + (int64_t)row * stride - (int64_t)(3 * hl) * (int64_t)px, (size_t)lcols * px);
    }
    RB_CUDA(cudaMemcpyAsync(dl_rows, hl_rows, 4 * lp, cudaMemcpyHostToDevice, hc.stream()));
    uint8_t lbuf[64 * 4 * 2] = {};
    if (hl && left) memcpy(lbuf, left, (size_t)h * 4 * px);
    const uint8_t *d_left = (const uint8_t *)hc.up(lbuf, sizeof(lbuf));
    const int pitch = w + 6;
    uint8_t *d_tmp = (uint8_t *)hc.dev((size_t)(h + 6) * pitch * px);
    if (hc.err) return hc.err;
    LrUnit U = {};
    if (kind < 2) {
        U.kind = 1;
        for (int k = 0; k < 8; k++) { U.fh[k] = params->filter[0][k]; U.fv[k] = params->filter[1][k]; }
        if (bdmax <= 255) U.fh[3] += 128;  // the 8 bpc convention keeps the +128 out of the table (src/lr_apply.rs:67-72)
    } else {
        U.kind = kind;  // 2: 5x5, 3: 3x3, 4: mix
        U.s0 = params->sgr.s0; U.s1 = params->sgr.s1; U.w0 = params->sgr.w0; U.w1 = params->sgr.w1;
    }
    const int grid = (w + LR_TW - 1) / LR_TW;
    if (bdmax > 255) {
        lr_pad_kernel<BD16><<<8, 256, 0, hc.stream()>>>(d_tmp, pitch, rect.dptr, rect.dpitch, d_left, dl_rows, (int64_t)lp, w, h, edges);
        lr_call_kernel<BD16><<<grid, 256, 0, hc.stream()>>>(d_tmp, pitch, rect.dptr, rect.dpitch, w, h, U, bdmax);
    } else {
        lr_pad_kernel<BD8><<<8, 256, 0, hc.stream()>>>(d_tmp, pitch, rect.dptr, rect.dpitch, d_left, dl_rows, (int64_t)lp, w, h, edges);
        lr_call_kernel<BD8><<<grid, 256, 0, hc.stream()>>>(d_tmp, pitch, rect.dptr, rect.dpitch, w, h, U, bdmax);
    }
    hc.rect_down(rect);
    if (hc.sync()) return hc.err;
    rect.row_bytes = w * px;  // only the unit itself is written back
    rect.finish(dst);
    return 0;
}

namespace {
template <int KIND>
void lr_slot(void *dst, ptrdiff_t stride, const void *left, const void *lpf, int w, int h,
             const Rb200LooprestorationParams *params, uint32_t edges, int bd) {
    if (rb200_lr(KIND, dst, stride, left, lpf, w, h, params, edges, bd)) rb200_report_fatal("lr");
}
}  // namespace

extern "C" void rb200_loop_restoration_dsp_init(Rb200LoopRestorationDSPContext *c, int bpc) {
    (void)bpc;
    c->wiener[0] = &lr_slot<0>; c->wiener[1] = &lr_slot<1>;
    c->sgr[0] = &lr_slot<2>; c->sgr[1] = &lr_slot<3>; c->sgr[2] = &lr_slot<4>;
}
