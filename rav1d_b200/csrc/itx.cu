// Inverse transform + add-to-destination ("itxfm_add") for sm_100a.
//
// Replaces Rav1dInvTxfmDSPContext.itxfm_add[19][17] (src/itx.rs:190-196; body
// inv_txfm_add_rust src/itx.rs:64-188 == inv_txfm_add_c src/itx_tmpl.c:40-113,
// WHT src/itx.rs:475-526 == src/itx_tmpl.c:173-193).
//
// Mapping: one GPU thread owns one row (first pass) and then one column (second
// pass) of a transform block, with the whole 1-D transform unrolled into
// registers (itx1d.cuh); the two passes meet in a padded shared-memory tile.
// A block of W x H uses T = max(W, min(H,32)) threads; a 128-thread CTA holds
// 128/T transform blocks.  Kernels are specialised per transform size; the
// 1-D kind (DCT/ADST/flipADST/identity/WHT) is a warp-uniform runtime switch,
// so the host buckets items by (size, type).  No tensor cores: this is an
// integer butterfly network, not a contraction.
#include "common.cuh"
#include <stdlib.h>
#include "stages.cuh"
#include "itx1d.cuh"
#include "itx_block.cuh"
#include <string.h>
#include <utility>

namespace rb200 {

static bool itx_valid(int tx, int txtp) {
    if (tx < 0 || tx >= 19 || txtp < 0 || txtp >= 17) return false;
    const int w = tx_w(tx), h = tx_h(tx), m = w > h ? w : h;
    if (txtp == RB200_WHT_WHT) return tx == RB200_TX_4X4;
    if (m == 64) return txtp == RB200_DCT_DCT;
    if (m == 32) return txtp == RB200_DCT_DCT || txtp == RB200_IDTX;
    if (w == 16 && h == 16) return txtp <= RB200_H_DCT;  // 12 types: no V/H (flip)ADST
    return true;
}

template <typename BD, int TX>
__global__ void __launch_bounds__(128)
itx_add_kernel(Rb200Planes planes, const typename BD::coef *__restrict__ cf,
               const Rb200ItxItem *__restrict__ items, int n_items, int bdmax) {
    using G = ItxGeom<TX>;
    using pixel = typename BD::pixel;
    constexpr int W = G::W, H = G::H, SW = G::SW, SH = G::SH;
    constexpr int shift = tx_shift(TX);
    constexpr bool rect2 = (W * 2 == H) || (H * 2 == W);
    __shared__ int tile[G::PER_CTA * G::TILE];

    const int slot = threadIdx.x / G::T, lane = threadIdx.x % G::T;
    const int idx = blockIdx.x * G::PER_CTA + slot;
    const bool live = idx < n_items;
    int *tl = tile + slot * G::TILE;

    Rb200ItxItem it;
    if (live) it = items[idx];
    else { it.cf_off = 0; it.x = it.y = 0; it.plane = 0; it.tx = TX; it.txtp = 0; it.eob = 0; it.ncols = 0; }

    itx_add_block<BD, TX>(tl, lane, live, it, planes, cf, bdmax);
}

template <typename BD, int TX>
static int launch_one(const Rb200Planes &planes, const void *cf, const Rb200ItxItem *items, int n, int bdmax,
                      cudaStream_t st) {
    using G = ItxGeom<TX>;
    if (n <= 0) return 0;
    const int grid = (n + G::PER_CTA - 1) / G::PER_CTA;
    itx_add_kernel<BD, TX><<<grid, G::CTA, 0, st>>>(planes, (const typename BD::coef *)cf, items, n, bdmax);
    RB_LAUNCH_CHECK();
    return 0;
}

template <typename BD>
static int launch_tx(int tx, const Rb200Planes &planes, const void *cf, const Rb200ItxItem *items, int n, int bdmax,
                     cudaStream_t st) {
    switch (tx) {
#define CASE(TX) case TX: return launch_one<BD, TX>(planes, cf, items, n, bdmax, st);
        CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9)
        CASE(10) CASE(11) CASE(12) CASE(13) CASE(14) CASE(15) CASE(16) CASE(17) CASE(18)
#undef CASE
    }
    return set_error(-22, "itx: bad transform size %d", tx);
}

// The part of each block's coefficients that can be non-zero -- its leading `ncols` columns, a contiguous
// prefix of the column-major block -- pulled from the pinned staging buffer into the device mirror.  One
// warp per block, 16 bytes per lane and load: many wide reads in flight keep the PCIe link busy, which the
// transform kernels' own 32..128-byte row reads would not.
template <typename coef>
__global__ void __launch_bounds__(256)
coef_gather_kernel(const coef *__restrict__ h_cf, coef *__restrict__ d_cf, const Rb200ItxItem *__restrict__ items, int n) {
    const int lane = threadIdx.x & 31;
    const int n_warps = (int)((gridDim.x * (unsigned)blockDim.x) >> 5);
    constexpr int PER16 = 16 / (int)sizeof(coef);
    // persistent warps; two blocks' loads are issued before either is stored (more PCIe reads in flight)
    for (int idx = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5); idx < n; idx += 2 * n_warps) {
        const coef *src[2]; coef *dst[2]; int count[2]; uint4 v[2][2];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int i = idx + k * n_warps;
            count[k] = 0;
            if (i >= n) continue;
            const Rb200ItxItem it = items[i];
            const int w = tx_w(it.tx), h = tx_h(it.tx), sw = w < 32 ? w : 32, sh = h < 32 ? h : 32;
            const int nc = it.ncols && it.ncols < sw ? it.ncols : sw;
            count[k] = nc * sh;
            src[k] = h_cf + it.cf_off; dst[k] = d_cf + it.cf_off;
        }
#pragma unroll
        for (int k = 0; k < 2; k++) {
            if (!count[k] || ((uintptr_t)src[k] & 15)) continue;
            const int n16 = count[k] / PER16;
#pragma unroll
            for (int u = 0; u < 2; u++)
                if (lane + 32 * u < n16) v[k][u] = __ldcs((const uint4 *)src[k] + lane + 32 * u);
        }
#pragma unroll
        for (int k = 0; k < 2; k++) {
            if (!count[k]) continue;
            if ((uintptr_t)src[k] & 15) {
                for (int i = lane; i < count[k]; i += 32) dst[k][i] = src[k][i];
                continue;
            }
            const int n16 = count[k] / PER16;
#pragma unroll
            for (int u = 0; u < 2; u++)
                if (lane + 32 * u < n16) ((uint4 *)dst[k])[lane + 32 * u] = v[k][u];
            for (int i = lane + 64; i < n16; i += 32) ((uint4 *)dst[k])[i] = __ldcs((const uint4 *)src[k] + i);   // > 1 KB blocks
            for (int i = n16 * PER16 + lane; i < count[k]; i += 32) dst[k][i] = src[k][i];
        }
    }
}

// Resident CTAs of the gather kernels: they sit on the device for as long as the link takes to deliver a frame's
// coefficients, so they should hold as few thread slots as keep the link full (a warp has 2 KB of reads in flight;
// the link's bandwidth-delay product is ~100 KB).
static int gather_ctas() {
    static const int n = getenv("RB200_GATHER_CTAS") ? atoi(getenv("RB200_GATHER_CTAS")) : 148;
    return n > 0 ? n : 148;
}

int coef_gather_launch(const void *h_cf, void *d_cf, const Rb200ItxItem *d_items, int n, int bdmax, cudaStream_t st) {
    if (n <= 0) return 0;
    const int grid = imin((n + 15) / 16, gather_ctas());
    if (bdmax > 255) coef_gather_kernel<int32_t><<<grid, 256, 0, st>>>((const int32_t *)h_cf, (int32_t *)d_cf, d_items, n);
    else coef_gather_kernel<int16_t><<<grid, 256, 0, st>>>((const int16_t *)h_cf, (int16_t *)d_cf, d_items, n);
    RB_LAUNCH_CHECK();
    return 0;
}

// int16 transport: 8 coefficients per 16-byte load over PCIe, two 16-byte stores of int32 into HBM.  One warp per block,
// two blocks' loads in flight before either is stored, like coef_gather_kernel.
__global__ void __launch_bounds__(256)
coef_gather16_kernel(const int16_t *__restrict__ h_cf, int32_t *__restrict__ d_cf, const Rb200ItxItem *__restrict__ items, int n) {
    const int lane = threadIdx.x & 31;
    const int n_warps = (int)((gridDim.x * (unsigned)blockDim.x) >> 5);
    for (int idx = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5); idx < n; idx += 2 * n_warps) {
        const int16_t *src[2]; int32_t *dst[2]; int count[2]; uint4 v[2][2];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int i = idx + k * n_warps;
            count[k] = 0;
            if (i >= n) continue;
            const Rb200ItxItem it = items[i];
            const int w = tx_w(it.tx), h = tx_h(it.tx), sw = w < 32 ? w : 32, sh = h < 32 ? h : 32;
            const int nc = it.ncols && it.ncols < sw ? it.ncols : sw;
            count[k] = nc * sh;
            src[k] = h_cf + it.cf_off; dst[k] = d_cf + it.cf_off;
        }
#pragma unroll
        for (int k = 0; k < 2; k++) {
            if (!count[k] || ((uintptr_t)src[k] & 15)) continue;
            const int n16 = count[k] >> 3;
#pragma unroll
            for (int u = 0; u < 2; u++)
                if (lane + 32 * u < n16) v[k][u] = __ldcs((const uint4 *)src[k] + lane + 32 * u);
        }
        auto widen = [](int32_t *d, const uint4 q) {
            ((int4 *)d)[0] = make_int4((int)(short)(q.x & 0xffff), (int)q.x >> 16, (int)(short)(q.y & 0xffff), (int)q.y >> 16);
            ((int4 *)d)[1] = make_int4((int)(short)(q.z & 0xffff), (int)q.z >> 16, (int)(short)(q.w & 0xffff), (int)q.w >> 16);
        };
#pragma unroll
        for (int k = 0; k < 2; k++) {
            if (!count[k]) continue;
            if ((uintptr_t)src[k] & 15) {      // (block offsets are multiples of 16 coefficients: not taken in practice)
                for (int i = lane; i < count[k]; i += 32) dst[k][i] = src[k][i];
                continue;
            }
            const int n16 = count[k] >> 3;
#pragma unroll
            for (int u = 0; u < 2; u++)
                if (lane + 32 * u < n16) widen(dst[k] + 8 * (lane + 32 * u), v[k][u]);
            for (int i = lane + 64; i < n16; i += 32) widen(dst[k] + 8 * i, __ldcs((const uint4 *)src[k] + i));   // > 512-coefficient prefixes
            for (int i = n16 * 8 + lane; i < count[k]; i += 32) dst[k][i] = src[k][i];
        }
    }
}
__global__ void coef_escape_kernel(int32_t *__restrict__ d_cf, const Rb200CoefEscape *__restrict__ esc, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) d_cf[esc[i].index] = esc[i].value;
}

// packed stream already in HBM: one warp per block, 16-byte loads, two 16-byte stores per load
__global__ void __launch_bounds__(256)
coef_expand16_kernel(const int16_t *__restrict__ stream, const uint32_t *__restrict__ offs, int32_t *__restrict__ d_cf,
                     const Rb200ItxItem *__restrict__ items, int n) {
    const int lane = threadIdx.x & 31;
    const int n_warps = (int)((gridDim.x * (unsigned)blockDim.x) >> 5);
    for (int i = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5); i < n; i += n_warps) {
        const Rb200ItxItem it = items[i];
        const int w = tx_w(it.tx), h = tx_h(it.tx), sw = w < 32 ? w : 32, sh = h < 32 ? h : 32;
        const int nc = it.ncols && it.ncols < sw ? it.ncols : sw;
        const int count = nc * sh;                      // a multiple of 4; the stream pads every block to a multiple of 8
        const uint4 *src = (const uint4 *)(stream + offs[i]);
        int32_t *dst = d_cf + it.cf_off;
        for (int k = lane; 8 * k < count; k += 32) {
            const uint4 q = __ldcs(src + k);
            ((int4 *)dst)[2 * k] = make_int4((int)(short)(q.x & 0xffff), (int)q.x >> 16, (int)(short)(q.y & 0xffff), (int)q.y >> 16);
            if (8 * k + 4 < count)
                ((int4 *)dst)[2 * k + 1] = make_int4((int)(short)(q.z & 0xffff), (int)q.z >> 16, (int)(short)(q.w & 0xffff), (int)q.w >> 16);
        }
    }
}

int coef_expand16_launch(const int16_t *d_stream, const uint32_t *d_off, int32_t *d_cf, const Rb200ItxItem *d_items, int n,
                         const Rb200CoefEscape *d_esc, int n_esc, cudaStream_t st, int *launches) {
    if (n <= 0) return 0;
    coef_expand16_kernel<<<imin((n + 7) / 8, 148 * 8), 256, 0, st>>>(d_stream, d_off, d_cf, d_items, n);
    if (launches) ++*launches;
    if (n_esc > 0) {
        coef_escape_kernel<<<imin((n_esc + 255) / 256, 148), 256, 0, st>>>(d_cf, d_esc, n_esc);
        if (launches) ++*launches;
    }
    RB_LAUNCH_CHECK();
    return 0;
}

int coef_gather16_launch(const int16_t *h_cf16, int32_t *d_cf, const Rb200ItxItem *d_items, int n, const Rb200CoefEscape *d_esc,
                         int n_esc, cudaStream_t st, int *launches) {
    if (n <= 0) return 0;
    const int grid = imin((n + 15) / 16, gather_ctas());
    coef_gather16_kernel<<<grid, 256, 0, st>>>(h_cf16, d_cf, d_items, n);
    if (launches) ++*launches;
    if (n_esc > 0) {
        coef_escape_kernel<<<imin((n_esc + 255) / 256, 148), 256, 0, st>>>(d_cf, d_esc, n_esc);
        if (launches) ++*launches;
    }
    RB_LAUNCH_CHECK();
    return 0;
}

int itx_launch(int tx, const Rb200Planes &planes, const void *cf, const Rb200ItxItem *items, int n, int bdmax,
               cudaStream_t st) {
    return bdmax > 255 ? launch_tx<BD16>(tx, planes, cf, items, n, bdmax, st)
                       : launch_tx<BD8>(tx, planes, cf, items, n, bdmax, st);
}

}  // namespace rb200

using namespace rb200;

// ---------------------------------------------------------------- C ABI
extern "C" int rb200_itx_valid(int tx, int txtp) { return itx_valid(tx, txtp) ? 1 : 0; }

extern "C" int rb200_itx_add_batch(const Rb200Planes *planes, const void *d_coef, const Rb200ItxItem *d_items,
                                   const int32_t counts[19], int bitdepth_max, void *stream) {
    if (!planes || !counts) return set_error(-22, "itx_add_batch: null argument");
    int off = 0;
    for (int tx = 0; tx < 19; tx++) {
        const int n = counts[tx];
        if (n < 0) return set_error(-22, "itx_add_batch: negative count");
        if (n) {
            int r = itx_launch(tx, *planes, d_coef, d_items + off, n, bitdepth_max, (cudaStream_t)stream);
            if (r) return r;
        }
        off += n;
    }
    return 0;
}

// Host-pointer, synchronous, one transform block: exact semantics of one
// itxfm_add[tx][txtp] call including zeroing of the consumed coefficients
// (src/itx.rs:94 dc-only zeroes coeff[0]; :152-158 full path zeroes sw*sh).
extern "C" int rb200_itxfm_add(int tx, int txtp, void *dst, ptrdiff_t stride, void *coeff, int eob,
                               int bitdepth_max) {
    if (!itx_valid(tx, txtp)) return set_error(-22, "itxfm_add: no such transform tx=%d txtp=%d", tx, txtp);
    if (eob < 0 || !dst || !coeff) return set_error(-22, "itxfm_add: bad argument");
    const bool hbd = bitdepth_max > 255;
    const int w = tx_w(tx), h = tx_h(tx), sw = w < 32 ? w : 32, sh = h < 32 ? h : 32;
    const size_t px = hbd ? 2 : 1, cs = hbd ? 4 : 2;
    const size_t ncoef = (size_t)sw * sh;
    Staging &st = staging();
    int r = st.begin(DevRect::bytes_for(w * px, h) + ncoef * cs + sizeof(Rb200ItxItem) + 1024,
                     DevRect::bytes_for(w * px, h) + ncoef * cs + sizeof(Rb200ItxItem) + 1024);
    if (r) return r;
    DevRect rect;
    if ((r = rect.upload(st, dst, stride, w * px, h))) return r;
    void *d_cf = st.dalloc(ncoef * cs);
    void *h_cf = st.halloc(ncoef * cs);
    memcpy(h_cf, coeff, ncoef * cs);
    RB_CUDA(cudaMemcpyAsync(d_cf, h_cf, ncoef * cs, cudaMemcpyHostToDevice, st.stream));
    Rb200ItxItem *h_it = (Rb200ItxItem *)st.halloc(sizeof(Rb200ItxItem));
    Rb200ItxItem *d_it = (Rb200ItxItem *)st.dalloc(sizeof(Rb200ItxItem));
    memset(h_it, 0, sizeof(*h_it));
    h_it->tx = (uint8_t)tx; h_it->txtp = (uint8_t)txtp; h_it->eob = (int16_t)(eob > 32767 ? 32767 : eob);
    RB_CUDA(cudaMemcpyAsync(d_it, h_it, sizeof(*h_it), cudaMemcpyHostToDevice, st.stream));
    Rb200Planes pl = {};
    pl.data[0] = rect.dptr; pl.stride[0] = rect.dpitch;
    if ((r = itx_launch(tx, pl, d_cf, d_it, 1, bitdepth_max, st.stream))) return r;
    if ((r = rect.download(st))) return r;
    RB_CUDA(cudaStreamSynchronize(st.stream));
    rect.finish(dst);
    if (txtp == RB200_DCT_DCT && eob < 1) memset(coeff, 0, cs);
    else memset(coeff, 0, ncoef * cs);
    return 0;
}

// ---- function-pointer table (drop-in for rav1d_itx_dsp_init, src/itx.rs:1072-1105) ----
namespace {
template <int TX, int TXTP>
void itx_slot(void *dst, ptrdiff_t stride, void *coeff, int eob, int bitdepth_max) {
    if (rb200_itxfm_add(TX, TXTP, dst, stride, coeff, eob, bitdepth_max)) rb200_report_fatal("itxfm_add");
}
template <int TX, int... TP>
void fill_row(Rb200InvTxfmDSPContext *c, std::integer_sequence<int, TP...>) {
    ((c->itxfm_add[TX][TP] = itx_valid(TX, TP) ? &itx_slot<TX, TP> : nullptr), ...);
}
template <int... TX>
void fill_all(Rb200InvTxfmDSPContext *c, std::integer_sequence<int, TX...>) {
    (fill_row<TX>(c, std::make_integer_sequence<int, 17>{}), ...);
}
}  // namespace

extern "C" void rb200_itx_dsp_init(Rb200InvTxfmDSPContext *c, int bpc) {
    (void)bpc;  // the bit depth travels in each call's bitdepth_max, as in the Rust fn-pointer ABI
    fill_all(c, std::make_integer_sequence<int, 19>{});
}
