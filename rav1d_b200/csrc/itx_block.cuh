// The transform-block body of the inverse transforms, shared by the per-size kernels (itx.cu) and the fused
// intra-block kernel (ipred.cu).  See itx.cu for the mapping.
#pragma once
#include "common.cuh"
#include "itx1d.cuh"

namespace rb200 {

// (w, h, shift) per RectTxfmSize, index = enum order of src/levels.rs:31-59.
// shift values: src/itx.rs:439-457 == src/itx_tmpl.c:153-171.
__host__ __device__ constexpr int tx_w(int tx) {
    constexpr int W[19] = {4, 8, 16, 32, 64, 4, 8, 8, 16, 16, 32, 32, 64, 4, 16, 8, 32, 16, 64};
    return W[tx];
}
__host__ __device__ constexpr int tx_h(int tx) {
    constexpr int H[19] = {4, 8, 16, 32, 64, 8, 4, 16, 8, 32, 16, 64, 32, 16, 4, 32, 8, 64, 16};
    return H[tx];
}
__host__ __device__ constexpr int tx_shift(int tx) {
    constexpr int S[19] = {0, 1, 2, 2, 2, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2};
    return S[tx];
}

// txtp -> 1-D kinds.  "A_B" = A vertical (columns, second pass), B horizontal
// (rows, first pass): src/levels.rs:63-82, assignment src/itx.rs:978-1068.
// packed 3-bit kinds per txtp (index order of RB200_* txtp enum):
//   rows: DCT DCT ADST ADST DCT FLIPADST FLIPADST FLIPADST ADST IDENTITY IDENTITY DCT IDENTITY ADST IDENTITY FLIPADST WHT
//   cols: DCT ADST DCT ADST FLIPADST DCT FLIPADST ADST FLIPADST IDENTITY DCT IDENTITY ADST IDENTITY FLIPADST IDENTITY WHT
__host__ __device__ inline int txtp_row_kind(int txtp) { return (int)((0x44cb0d9490240ULL >> (3 * txtp)) & 7); }
__host__ __device__ inline int txtp_col_kind(int txtp) { return (int)((0x469961a282208ULL >> (3 * txtp)) & 7); }

template <int N>
__device__ __forceinline__ void run_kind(int kind, int *x, int lo, int hi) {
    if constexpr (N == 64) {
        itx1d<64, T1_DCT>(x, lo, hi);
    } else if constexpr (N == 32) {
        if (kind == T1_DCT) itx1d<32, T1_DCT>(x, lo, hi);
        else itx1d<32, T1_IDENTITY>(x, lo, hi);
    } else {
        switch (kind) {
        case T1_DCT: itx1d<N, T1_DCT>(x, lo, hi); break;
        case T1_ADST: itx1d<N, T1_ADST>(x, lo, hi); break;
        case T1_FLIPADST: itx1d<N, T1_FLIPADST>(x, lo, hi); break;
        case T1_IDENTITY: itx1d<N, T1_IDENTITY>(x, lo, hi); break;
        default: if constexpr (N == 4) wht4(x); break;
        }
    }
}

template <int TX>
struct ItxGeom {
    static constexpr int W = tx_w(TX), H = tx_h(TX);
    static constexpr int SW = W < 32 ? W : 32, SH = H < 32 ? H : 32;
    static constexpr int T = W > SH ? W : SH;   // threads per transform block
    static constexpr int CTA = 128;
    static constexpr int PER_CTA = CTA / T;
    static constexpr int PITCH = W + 1;         // padded row pitch (ints) -> conflict-free both passes
    static constexpr int TILE = PITCH * SH;
};

// One transform block: `lane` = thread index inside the block's group of ItxGeom<TX>::T threads, `tl` = the group's
// shared-memory tile (ItxGeom<TX>::TILE ints), `live` = false for idle groups.  EVERY thread of the CTA must call this
// (it contains a __syncthreads between the passes).
template <typename BD, int TX>
__device__ __forceinline__ void itx_add_block(int *tl, int lane, bool live, const Rb200ItxItem &it, const Rb200Planes &planes,
                                              const typename BD::coef *__restrict__ cf, int bdmax) {
    using G = ItxGeom<TX>;
    using pixel = typename BD::pixel;
    constexpr int W = G::W, H = G::H, SW = G::SW, SH = G::SH;
    constexpr int shift = tx_shift(TX);
    constexpr bool rect2 = (W * 2 == H) || (H * 2 == W);
    const int row_lo = BD::hbd ? (int)((unsigned)~bdmax << 7) : -32768;
    const int col_lo = BD::hbd ? (int)((unsigned)~bdmax << 5) : -32768;
    const int row_hi = ~row_lo, col_hi = ~col_lo;
    const typename BD::coef *c = cf + it.cf_off;
    const bool dconly = live && it.txtp == RB200_DCT_DCT && it.eob < 1;
    const bool wht = it.txtp == RB200_WHT_WHT;

    // The destination column of the second pass does not depend on the transform: fetch it now so that
    // its latency overlaps the coefficient loads and the first pass (one global round trip less on the
    // critical path of this latency-bound kernel).
    const int64_t bstride = plane_stride(planes, it.plane);
    const int64_t pstride = bstride / (int64_t)sizeof(pixel);
    pixel *dst = (pixel *)(plane_ptr(planes, it.plane) + (int64_t)it.y * bstride) + it.x + lane;
    pixel dpx[H];
    if (live && lane < W) {
#pragma unroll
        for (int i = 0; i < H; i++) dpx[i] = dst[i * pstride];
    }

    // ---- first pass: rows ----
    if (live && !dconly && lane < SH) {
        int x[W];
        const int nc = it.ncols ? it.ncols : SW;   // columns beyond nc hold zeros by contract and are not read
        if (wht) {
#pragma unroll
            for (int i = 0; i < SW; i++) x[i] = i < nc ? (int)c[lane + i * SH] >> 2 : 0;
        } else {
#pragma unroll
            for (int i = 0; i < SW; i++) {
                const int v = i < nc ? (int)c[lane + i * SH] : 0;
                x[i] = rect2 ? (v * 181 + 128) >> 8 : v;
            }
        }
        run_kind<W>(txtp_row_kind(it.txtp), x, row_lo, row_hi);
        constexpr int rnd = (1 << shift) >> 1;
        if (wht) {
#pragma unroll
            for (int i = 0; i < W; i++) tl[lane * G::PITCH + i] = x[i];
        } else {
#pragma unroll
            for (int i = 0; i < W; i++) tl[lane * G::PITCH + i] = iclip((x[i] + rnd) >> shift, col_lo, col_hi);
        }
    }
    __syncthreads();
    if (!live || lane >= W) return;

    // ---- second pass: columns, add to destination ----
    if (dconly) {
        // src/itx.rs:90-111
        int dc = c[0];
        if (rect2) dc = (dc * 181 + 128) >> 8;
        dc = (dc * 181 + 128) >> 8;
        dc = (dc + ((1 << shift) >> 1)) >> shift;
        dc = (dc * 181 + 128 + 2048) >> 12;
#pragma unroll
        for (int i = 0; i < H; i++) dst[i * pstride] = (pixel)iclip((int)dpx[i] + dc, 0, bdmax);
        return;
    }
    int v[H];
#pragma unroll
    for (int i = 0; i < SH; i++) v[i] = tl[i * G::PITCH + lane];
    run_kind<H>(txtp_col_kind(it.txtp), v, col_lo, col_hi);
    if (wht) {
#pragma unroll
        for (int i = 0; i < H; i++) dst[i * pstride] = (pixel)iclip((int)dpx[i] + v[i], 0, bdmax);
    } else {
#pragma unroll
        for (int i = 0; i < H; i++) dst[i * pstride] = (pixel)iclip((int)dpx[i] + ((v[i] + 8) >> 4), 0, bdmax);
    }
}

// The same transform, but the term that itx_add_block adds to the destination is left in shared memory instead
// (`res`, W x H ints, row-major): the fused intra kernel computes it while it waits for the block's neighbours and adds
// it once the prediction is in place.  EVERY thread of the CTA must call this.
template <typename BD, int TX>
__device__ __forceinline__ void itx_residual_block(int *tl, int *res, int lane, bool live, const Rb200ItxItem &it,
                                                   const typename BD::coef *__restrict__ cf, int bdmax) {
    using G = ItxGeom<TX>;
    constexpr int W = G::W, H = G::H, SW = G::SW, SH = G::SH;
    constexpr int shift = tx_shift(TX);
    constexpr bool rect2 = (W * 2 == H) || (H * 2 == W);
    const int row_lo = BD::hbd ? (int)((unsigned)~bdmax << 7) : -32768;
    const int col_lo = BD::hbd ? (int)((unsigned)~bdmax << 5) : -32768;
    const int row_hi = ~row_lo, col_hi = ~col_lo;
    const typename BD::coef *c = cf + it.cf_off;
    const bool dconly = live && it.txtp == RB200_DCT_DCT && it.eob < 1;
    const bool wht = it.txtp == RB200_WHT_WHT;
    if (live && !dconly && lane < SH) {
        int x[W];
        const int nc = it.ncols ? it.ncols : SW;
        if (wht) {
#pragma unroll
            for (int i = 0; i < SW; i++) x[i] = i < nc ? (int)c[lane + i * SH] >> 2 : 0;
        } else {
#pragma unroll
            for (int i = 0; i < SW; i++) {
                const int v = i < nc ? (int)c[lane + i * SH] : 0;
                x[i] = rect2 ? (v * 181 + 128) >> 8 : v;
            }
        }
        run_kind<W>(txtp_row_kind(it.txtp), x, row_lo, row_hi);
        constexpr int rnd = (1 << shift) >> 1;
        if (wht) {
#pragma unroll
            for (int i = 0; i < W; i++) tl[lane * G::PITCH + i] = x[i];
        } else {
#pragma unroll
            for (int i = 0; i < W; i++) tl[lane * G::PITCH + i] = iclip((x[i] + rnd) >> shift, col_lo, col_hi);
        }
    }
    __syncthreads();
    if (!live || lane >= W) return;
    if (dconly) {   // src/itx.rs:90-111
        int dc = c[0];
        if (rect2) dc = (dc * 181 + 128) >> 8;
        dc = (dc * 181 + 128) >> 8;
        dc = (dc + ((1 << shift) >> 1)) >> shift;
        dc = (dc * 181 + 128 + 2048) >> 12;
#pragma unroll
        for (int i = 0; i < H; i++) res[i * W + lane] = dc;
        return;
    }
    int v[H];
#pragma unroll
    for (int i = 0; i < SH; i++) v[i] = tl[i * G::PITCH + lane];
    run_kind<H>(txtp_col_kind(it.txtp), v, col_lo, col_hi);
#pragma unroll
    for (int i = 0; i < H; i++) res[i * W + lane] = wht ? v[i] : (v[i] + 8) >> 4;
}

}  // namespace rb200
