// Film grain synthesis for sm_100a.
//
// Replaces Rav1dFilmGrainDSPContext {generate_grain_y, generate_grain_uv[3], fgy_32x32xn,
// fguv_32x32xn[3]} (src/filmgrain.rs:195-201; generate_grain_y_rust :298, _uv_rust :345,
// fgy_32x32xn_rust :549, fguv_32x32xn_rust :678 == src/filmgrain_tmpl.c:37-400) and the
// driver half of src/fg_apply.rs (generate_scaling :14, rav1d_prep_grain :74,
// rav1d_apply_grain_row :174 == src/fg_apply_tmpl.c:41-245).
//
// The serial pieces of the CPU algorithm are restated so that they parallelise without
// changing a bit:
//  * the 16-bit LFSR (get_random_number, src/filmgrain_tmpl.c:38-44) is linear over GF(2),
//    so the state at the start of LUT row y is J^y * seed with J = (one step)^row_width;
//    every row of the grain LUT is then drawn by its own thread;
//  * the causal AR filter (lag <= 3) is a skewed wavefront: row y may compute column x as
//    soon as row y-1 has finished column x+3, i.e. at step t = x + 4y;
//  * the per-32x32-block offsets (one LFSR per block row, stepped once per block column)
//    are expanded into a small table by one thread per block row, after which grain
//    application is independent per pixel.
#include "common.cuh"
#include "tables.cuh"

namespace rb200 {

constexpr int GW = RB200_GRAIN_WIDTH, GH = RB200_GRAIN_HEIGHT;
constexpr int SUB_GW = 44, SUB_GH = 38;

__host__ __device__ __forceinline__ unsigned fg_lfsr_step(unsigned r) {
    const unsigned bit = ((r >> 0) ^ (r >> 1) ^ (r >> 3) ^ (r >> 12)) & 1;
    return (r >> 1) | (bit << 15);
}
__device__ __forceinline__ int fg_round2(int x, int shift) { return (x + ((1 << shift) >> 1)) >> shift; }

template <typename BD> struct FgEntry { using type = int8_t; };
template <> struct FgEntry<BD16> { using type = int16_t; };

// y = M * s over GF(2), M given as the images of the 16 basis vectors
__device__ __forceinline__ unsigned fg_gf2_apply(const uint16_t *M, unsigned s) {
    unsigned n = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) n ^= ((s >> i) & 1) ? M[i] : 0;
    return n;
}

// One CTA generates one grain LUT.  uv0 < 0: luma; otherwise CTA b generates chroma plane uv0 + b into
// lut + b * lut_pitch.  lut_y: finished luma LUT (chroma only).  LAG = ar_coeff_lag.
template <typename BD, int LAG>
__global__ void __launch_bounds__(96)
fg_generate_kernel(typename FgEntry<BD>::type *__restrict__ lut, const typename FgEntry<BD>::type *__restrict__ lut_y,
                   Rb200FilmGrainData d, int uv0, int subx, int suby, int bdmax, int lut_pitch) {
    __shared__ int16_t buf[GH][GW];
    __shared__ int16_t lum[GH][GW];
    __shared__ uint16_t Jp[7][16];
    const int tid = threadIdx.x;
    const int uv = uv0 < 0 ? -1 : uv0 + (int)blockIdx.x;
    lut += (size_t)blockIdx.x * lut_pitch;
    const int bdmin8 = BD::hbd ? bpc_from_max(bdmax) - 8 : 0;
    const int W = (uv >= 0 && subx) ? SUB_GW : GW, H = (uv >= 0 && suby) ? SUB_GH : GH;
    const int shift = 4 - bdmin8 + d.grain_scale_shift;
    const int grain_ctr = 128 << bdmin8, grain_min = -grain_ctr, grain_max = grain_ctr - 1;
    // J = (LFSR step)^W and its squarings J^2, J^4, .. J^64
    if (tid < 16) {
        unsigned v = 1u << tid;
        for (int i = 0; i < W; i++) v = fg_lfsr_step(v);
        Jp[0][tid] = (uint16_t)v;
    }
    __syncthreads();
    for (int k = 1; k < 7; k++) {
        if (tid < 16) Jp[k][tid] = (uint16_t)fg_gf2_apply(Jp[k - 1], Jp[k - 1][tid]);
        __syncthreads();
    }
    if (tid < H) {
        unsigned s = uv < 0 ? d.seed : d.seed ^ (uv ? 0x49d8u : 0xb524u);
        for (int k = 0; k < 7; k++)
            if ((tid >> k) & 1) s = fg_gf2_apply(Jp[k], s);          // s = J^tid * seed
        for (int x = 0; x < W; x++) {
            s = fg_lfsr_step(s);
            buf[tid][x] = (int16_t)fg_round2(tab::k_gaussian_sequence[(s >> 5) & 0x7ff], shift);
        }
    }
    // auto-regressive filter, raster-causal (src/filmgrain_tmpl.c:66-84,112-148)
    constexpr int NC = 2 * LAG * (LAG + 1);
    int coef[NC + 1];
#pragma unroll
    for (int i = 0; i <= NC; i++) coef[i] = uv < 0 ? (i < 24 ? d.ar_coeffs_y[i < 24 ? i : 0] : 0) : d.ar_coeffs_uv[uv][i];
    const int ar_shift = (int)d.ar_coeff_shift;
    const bool luma_term = uv >= 0 && d.num_y_points != 0;
    if (luma_term) {   // the luma-grain term of every chroma entry, up front
        for (int i = tid; i < (H - 3) * (W - 6); i += blockDim.x) {
            const int y = 3 + i / (W - 6), x = 3 + i % (W - 6);
            const int lx = ((x - 3) << subx) + 3, ly = ((y - 3) << suby) + 3;
            int l = 0;
            for (int a = 0; a <= suby; a++)
                for (int b = 0; b <= subx; b++) l += lut_y[(ly + a) * GW + lx + b];
            lum[y][x] = (int16_t)fg_round2(l, subx + suby);
        }
    }
    __syncthreads();
    // row y computes column x at step t = x - 3 + (LAG + 1) * (y - 3): row y - 1 has then finished x + LAG
    constexpr int SKEW = LAG + 1;
    const int y = tid;
    const int n_steps = LAG ? SKEW * (H - 4) + W - 6 : W - 6;
    for (int t = 0; t < n_steps; t++) {
        const int x = LAG ? t - SKEW * (y - 3) + 3 : t + 3;
        if (y >= 3 && y < H && x >= 3 && x < W - 3) {
            int v[NC ? NC : 1];
            int c = 0;
#pragma unroll
            for (int dy = -LAG; dy <= 0; dy++) {
#pragma unroll
                for (int dx = -LAG; dx <= LAG; dx++) {
                    if (dy == 0 && dx >= 0) continue;
                    v[c++] = buf[y + dy][x + dx];
                }
            }
            int sum = luma_term ? (int)lum[y][x] * coef[NC] : 0;
#pragma unroll
            for (int i = 0; i < NC; i++) sum += coef[i] * v[i];
            buf[y][x] = (int16_t)iclip(buf[y][x] + fg_round2(sum, ar_shift), grain_min, grain_max);
        }
        if (LAG) __syncthreads();
    }
    __syncthreads();
    for (int i = tid; i < H * W; i += blockDim.x) {
        const int r = i / W, c = i - r * W;
        lut[r * GW + c] = (typename FgEntry<BD>::type)buf[r][c];
    }
}

// generate_scaling, src/fg_apply_tmpl.c:41-96.  One CTA of 256 threads; thread x owns 8-bit index x.
__global__ void __launch_bounds__(256)
fg_scaling_kernel(uint8_t *__restrict__ scaling, const FgPoints pts /* [num][2] */, int num, int bitdepth) {
    __shared__ uint8_t v8[257];
    const uint8_t *points = pts.p;
    const int x = threadIdx.x;
    const int shift_x = bitdepth - 8, pad = 1 << shift_x, rnd = pad >> 1;
    if (num == 0) {
        for (int n = 0; n < pad; n++) scaling[(x << shift_x) + n] = 0;
        return;
    }
    const int first = points[0], last = points[2 * (num - 1)];
    int v;
    if (x < first) v = points[1];
    else if (x >= last) v = points[2 * (num - 1) + 1];
    else {
        int i = 0;
        while (i < num - 2 && x >= points[2 * (i + 1)]) i++;
        const int bx = points[2 * i], by = points[2 * i + 1], ex = points[2 * i + 2], ey = points[2 * i + 3];
        const int dx = ex - bx, dy = ey - by;
        const int delta = dy * ((0x10000 + (dx >> 1)) / dx);
        v = by + ((0x8000 + delta * (x - bx)) >> 16);
    }
    v8[x] = (uint8_t)v;
    if (x == 0) v8[256] = points[2 * (num - 1) + 1];
    __syncthreads();
    if (shift_x == 0) { scaling[x] = v8[x]; return; }
    if (x >= first && x < last) {
        const int base = v8[x], range = (int)v8[x + 1] - base;
        for (int n = 0; n < pad; n++) scaling[(x << shift_x) + n] = (uint8_t)(base + (n ? (rnd + n * range) >> shift_x : 0));
    } else {
        for (int n = 0; n < pad; n++) scaling[(x << shift_x) + n] = v8[x];
    }
}

// Block offsets: off[r * ncols + c] = the 8-bit draw of block row (row0 + r), block column c
// (src/filmgrain_tmpl.c:184-209).
__global__ void fg_offsets_kernel(uint8_t *__restrict__ off, unsigned seed0, int row0, int nrows, int ncols) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    const int row = row0 + r;
    unsigned s = seed0;
    s ^= (unsigned)(((row * 37 + 178) & 0xFF) << 8);
    s ^= (unsigned)((row * 173 + 105) & 0xFF);
    for (int c = 0; c < ncols; c++) {
        s = fg_lfsr_step(s);
        off[r * ncols + c] = (uint8_t)((s >> 8) & 0xff);
    }
}

template <typename BD>
__global__ void __launch_bounds__(256)
fg_apply_kernel(uint8_t *__restrict__ dst, const uint8_t *__restrict__ src, int64_t stride,
                const uint8_t *__restrict__ luma, int64_t luma_stride, FgApplyParams P,
                const uint8_t *__restrict__ scaling, const typename FgEntry<BD>::type *__restrict__ lut,
                const uint8_t *__restrict__ off, int bdmax) {
    using pixel = typename BD::pixel;
    const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= P.pw || y >= P.ph) return;
    const int bdmin8 = BD::hbd ? bpc_from_max(bdmax) - 8 : 0;
    const int grain_ctr = 128 << bdmin8, grain_min = -grain_ctr, grain_max = grain_ctr - 1;
    const int BW = 32 >> P.sx, BH = 32 >> P.sy;
    const int bx = x / BW, lx = x - bx * BW, byr = y / BH, ly = y - byr * BH;
    const int row = P.row0 + byr;
    const int bw = imin(BW, P.pw - bx * BW), bh = imin(BH, P.ph - byr * BH);
    const int ystart = (P.overlap && row) ? imin(2 >> P.sy, bh) : 0;
    const int xstart = (P.overlap && bx) ? imin(2 >> P.sx, bw) : 0;
    const uint8_t *orow = off + (row - P.off_row0) * P.ncols;
    auto sample = [&](int bxi, int byi) -> int {
        const int rv = (byi ? orow - P.ncols : orow)[bx - bxi];
        const int offx = 3 + (2 >> P.sx) * (3 + (rv >> 4)), offy = 3 + (2 >> P.sy) * (3 + (rv & 0xF));
        return lut[(offy + ly + BH * byi) * GW + offx + lx + BW * bxi];
    };
    // overlap weights: full-resolution {27,17},{17,27}; sub-sampled {23,22} (src/filmgrain_tmpl.c:211,325-328)
    auto wgt = [](int sub, int i, int k) -> int { return sub ? (k ? 22 : 23) : ((i ^ k) ? 17 : 27); };
    auto blend = [&](int old, int cur, int sub, int i) -> int {
        return iclip(fg_round2(old * wgt(sub, i, 0) + cur * wgt(sub, i, 1), 5), grain_min, grain_max);
    };
    int grain = sample(0, 0);
    if (ly >= ystart) {
        if (lx < xstart) grain = blend(sample(1, 0), grain, P.sx, lx);
    } else {
        if (lx >= xstart) {
            grain = blend(sample(0, 1), grain, P.sy, ly);
        } else {
            const int top = blend(sample(1, 1), sample(0, 1), P.sx, lx);
            grain = blend(sample(1, 0), grain, P.sx, lx);
            grain = blend(top, grain, P.sy, ly);
        }
    }
    const int64_t ps = stride / (int64_t)sizeof(pixel);
    const int s = ((const pixel *)src)[(int64_t)y * ps + x];
    int min_value = 0, max_value = bdmax;
    if (P.clip) { min_value = 16 << bdmin8; max_value = ((P.chroma && !P.is_id) ? 240 : 235) << bdmin8; }
    int val = s;
    if (P.chroma) {
        const pixel *l = (const pixel *)luma + (int64_t)(y << P.sy) * (luma_stride / (int64_t)sizeof(pixel));
        const int l0 = x << P.sx;
        int avg = l[l0];
        if (P.sx) avg = (avg + l[imin(l0 + 1, P.luma_w - 1)] + 1) >> 1;
        val = avg;
        if (!P.csfl) {
            const int combined = avg * P.uv_luma_mult + s * P.uv_mult;
            val = iclip((combined >> 6) + P.uv_offset * (1 << bdmin8), 0, bdmax);
        }
    }
    const int noise = fg_round2(scaling[val] * grain, P.scaling_shift);
    ((pixel *)dst)[(int64_t)y * ps + x] = (pixel)iclip(s + noise, min_value, max_value);
}

// Two horizontally adjacent pixels per thread, sub-sampling known at compile time (block geometry is
// shifts, pixels move as 32-bit pairs).  Same arithmetic as fg_apply_kernel; used whenever the planes
// are aligned for pair accesses (always for the frame stage).
template <typename BD, int SX, int SY, bool CHROMA>
__global__ void __launch_bounds__(256)
fg_apply_pair_kernel(uint8_t *__restrict__ dst, const uint8_t *__restrict__ src, int64_t stride,
                     const uint8_t *__restrict__ luma, int64_t luma_stride, FgApplyParams P,
                     const uint8_t *__restrict__ scaling, const typename FgEntry<BD>::type *__restrict__ lut,
                     const uint8_t *__restrict__ off, int bdmax) {
    using pixel = typename BD::pixel;
    using pair_t = typename BD::pair;            // two pixels in one word
    constexpr int BWL = 5 - SX, BHL = 5 - SY, BW = 1 << BWL, BH = 1 << BHL;
    const int x = blockIdx.x * 128 + (threadIdx.x & 63) * 2, y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= P.pw || y >= P.ph) return;
    const int bdmin8 = BD::hbd ? bpc_from_max(bdmax) - 8 : 0;
    const int grain_ctr = 128 << bdmin8, grain_min = -grain_ctr, grain_max = grain_ctr - 1;
    const int bx = x >> BWL, lx0 = x & (BW - 1), byr = y >> BHL, ly = y & (BH - 1);
    const int row = P.row0 + byr;
    const int bw = imin(BW, P.pw - (bx << BWL)), bh = imin(BH, P.ph - (byr << BHL));
    const int ystart = (P.overlap && row) ? imin(2 >> SY, bh) : 0;
    const int xstart = (P.overlap && bx) ? imin(2 >> SX, bw) : 0;
    const uint8_t *orow = off + (row - P.off_row0) * P.ncols;
    const int rv = orow[bx];
    const int offx = 3 + (2 >> SX) * (3 + (rv >> 4)), offy = 3 + (2 >> SY) * (3 + (rv & 0xF));
    const typename FgEntry<BD>::type *lp = lut + (offy + ly) * GW + offx + lx0;
    int grain[2] = { lp[0], lp[1] };
    if (ly < ystart || lx0 < xstart) {
        // overlap with the block above and/or to the left (2 rows / columns at full resolution, 1 sub-sampled)
        auto sample = [&](int lx, int bxi, int byi) -> int {
            const int r = (byi ? orow - P.ncols : orow)[bx - bxi];
            const int ox = 3 + (2 >> SX) * (3 + (r >> 4)), oy = 3 + (2 >> SY) * (3 + (r & 0xF));
            return lut[(oy + ly + BH * byi) * GW + ox + lx + BW * bxi];
        };
        auto wgt = [](int sub, int i, int k) -> int { return sub ? (k ? 22 : 23) : ((i ^ k) ? 17 : 27); };
        auto blend = [&](int old, int cur, int sub, int i) -> int {
            return iclip(fg_round2(old * wgt(sub, i, 0) + cur * wgt(sub, i, 1), 5), grain_min, grain_max);
        };
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int lx = lx0 + k;
            int g = grain[k];
            if (ly >= ystart) {
                if (lx < xstart) g = blend(sample(lx, 1, 0), g, SX, lx);
            } else if (lx >= xstart) {
                g = blend(sample(lx, 0, 1), g, SY, ly);
            } else {
                const int top = blend(sample(lx, 1, 1), sample(lx, 0, 1), SX, lx);
                g = blend(sample(lx, 1, 0), g, SX, lx);
                g = blend(top, g, SY, ly);
            }
            grain[k] = g;
        }
    }
    const bool two = x + 1 < P.pw;
    const pixel *srow = (const pixel *)(src + (int64_t)y * stride);
    pixel *drow = (pixel *)(dst + (int64_t)y * stride);
    int s[2];
    if (two) { const pair_t q = *(const pair_t *)(srow + x); s[0] = BD::lo(q); s[1] = BD::hi(q); }
    else { s[0] = srow[x]; s[1] = 0; }
    int min_value = 0, max_value = bdmax;
    if (P.clip) { min_value = 16 << bdmin8; max_value = ((CHROMA && !P.is_id) ? 240 : 235) << bdmin8; }
    int val[2] = { s[0], s[1] };
    if (CHROMA) {
        const pixel *l = (const pixel *)(luma + (int64_t)(y << SY) * luma_stride);
        int avg[2];
        if (SX) {
            const int l0 = x << 1;
            // four luma samples; the last one of a row is clamped (odd luma widths)
            const int a = l[l0], b = l[imin(l0 + 1, P.luma_w - 1)], c = l[imin(l0 + 2, P.luma_w - 1)], d = l[imin(l0 + 3, P.luma_w - 1)];
            avg[0] = (a + b + 1) >> 1; avg[1] = (c + d + 1) >> 1;
        } else {
            avg[0] = l[x]; avg[1] = l[imin(x + 1, P.luma_w - 1)];
        }
#pragma unroll
        for (int k = 0; k < 2; k++) {
            val[k] = avg[k];
            if (!P.csfl) {
                const int combined = avg[k] * P.uv_luma_mult + s[k] * P.uv_mult;
                val[k] = iclip((combined >> 6) + P.uv_offset * (1 << bdmin8), 0, bdmax);
            }
        }
    }
    int o[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int noise = fg_round2(scaling[val[k]] * grain[k], P.scaling_shift);
        o[k] = iclip(s[k] + noise, min_value, max_value);
    }
    if (two) *(pair_t *)(drow + x) = BD::pack(o[0], o[1]);
    else drow[x] = (pixel)o[0];
}

// ---- launch helpers shared by the per-call entry points and the frame stage
template <typename BD>
static int fg_generate_launch(void *lut, const void *lut_y, const Rb200FilmGrainData &d, int uv, int n_luts, int lut_pitch_bytes,
                              int subx, int suby, int bdmax, cudaStream_t st) {
    using E = typename FgEntry<BD>::type;
    const int pitch = lut_pitch_bytes / (int)sizeof(E);
    switch (d.ar_coeff_lag) {
    case 0: fg_generate_kernel<BD, 0><<<n_luts, 96, 0, st>>>((E *)lut, (const E *)lut_y, d, uv, subx, suby, bdmax, pitch); break;
    case 1: fg_generate_kernel<BD, 1><<<n_luts, 96, 0, st>>>((E *)lut, (const E *)lut_y, d, uv, subx, suby, bdmax, pitch); break;
    case 2: fg_generate_kernel<BD, 2><<<n_luts, 96, 0, st>>>((E *)lut, (const E *)lut_y, d, uv, subx, suby, bdmax, pitch); break;
    case 3: fg_generate_kernel<BD, 3><<<n_luts, 96, 0, st>>>((E *)lut, (const E *)lut_y, d, uv, subx, suby, bdmax, pitch); break;
    default: return set_error(-22, "film grain: ar_coeff_lag > 3");
    }
    RB_LAUNCH_CHECK();
    return 0;
}

// uv < 0: the luma LUT.  Otherwise n_luts (1 or 2) chroma LUTs for planes uv, uv + 1 at lut, lut + lut_pitch_bytes.
int fg_generate(void *lut, const void *lut_y, const Rb200FilmGrainData &d, int uv, int subx, int suby, int bdmax,
                cudaStream_t st, int n_luts, int lut_pitch_bytes) {
    return bdmax > 255 ? fg_generate_launch<BD16>(lut, lut_y, d, uv, n_luts, lut_pitch_bytes, subx, suby, bdmax, st)
                       : fg_generate_launch<BD8>(lut, lut_y, d, uv, n_luts, lut_pitch_bytes, subx, suby, bdmax, st);
}

template <typename BD>
static void fg_apply_dispatch(uint8_t *dst, const uint8_t *src, int64_t stride, const uint8_t *luma, int64_t luma_stride,
                              const FgApplyParams &P, const uint8_t *scaling, const void *lut, const uint8_t *off, int bdmax,
                              cudaStream_t st) {
    using E = typename FgEntry<BD>::type;
    constexpr uintptr_t A = 2 * sizeof(typename BD::pixel) - 1;
    const bool aligned = !(((uintptr_t)dst | (uintptr_t)src | (uintptr_t)stride) & A);
    if (!aligned) {
        dim3 grid((P.pw + 63) / 64, (P.ph + 3) / 4);
        fg_apply_kernel<BD><<<grid, 256, 0, st>>>(dst, src, stride, luma, luma_stride, P, scaling, (const E *)lut, off, bdmax);
        return;
    }
    dim3 grid((P.pw + 127) / 128, (P.ph + 3) / 4);
#define L(SX, SY, C) fg_apply_pair_kernel<BD, SX, SY, C><<<grid, 256, 0, st>>>(dst, src, stride, luma, luma_stride, P, scaling, (const E *)lut, off, bdmax)
    if (!P.chroma) L(0, 0, false);
    else if (P.sx && P.sy) L(1, 1, true);
    else if (P.sx) L(1, 0, true);
    else L(0, 0, true);
#undef L
}

int fg_apply(uint8_t *dst, const uint8_t *src, int64_t stride, const uint8_t *luma, int64_t luma_stride,
             const FgApplyParams &P, const uint8_t *scaling, const void *lut, const uint8_t *off, int bdmax,
             cudaStream_t st) {
    if (bdmax > 255) fg_apply_dispatch<BD16>(dst, src, stride, luma, luma_stride, P, scaling, lut, off, bdmax, st);
    else fg_apply_dispatch<BD8>(dst, src, stride, luma, luma_stride, P, scaling, lut, off, bdmax, st);
    RB_LAUNCH_CHECK();
    return 0;
}

int fg_offsets(uint8_t *off, unsigned seed, int row0, int nrows, int ncols, cudaStream_t st) {
    fg_offsets_kernel<<<(nrows + 63) / 64, 64, 0, st>>>(off, seed, row0, nrows, ncols);
    RB_LAUNCH_CHECK();
    return 0;
}

int fg_scaling(uint8_t *scaling, const uint8_t points[][2], int num, int bitdepth, cudaStream_t st) {
    FgPoints pts = {};
    if (num > 0) memcpy(pts.p, points, (size_t)(num > 14 ? 14 : num) * 2);
    fg_scaling_kernel<<<1, 256, 0, st>>>(scaling, pts, num, bitdepth);
    RB_LAUNCH_CHECK();
    return 0;
}

FgApplyParams fg_params(const Rb200FilmGrainData &d, int chroma, int uv, int sx, int sy, int is_id) {
    FgApplyParams P = {};
    P.sx = sx; P.sy = sy; P.chroma = chroma; P.uv = uv; P.is_id = is_id;
    P.overlap = d.overlap_flag; P.clip = d.clip_to_restricted_range; P.scaling_shift = d.scaling_shift;
    P.csfl = d.chroma_scaling_from_luma;
    if (chroma) { P.uv_mult = d.uv_mult[uv]; P.uv_luma_mult = d.uv_luma_mult[uv]; P.uv_offset = d.uv_offset[uv]; }
    return P;
}

}  // namespace rb200

using namespace rb200;

// ---------------------------------------------------------------- C ABI (per call)
extern "C" int rb200_generate_grain_y(void *buf, const Rb200FilmGrainData *data, int bdmax) {
    if (!buf || !data) return set_error(-22, "generate_grain_y: null argument");
    const size_t es = bdmax > 255 ? 2 : 1, n = (size_t)GH * GW * es;
    HostCall hc(2 * n);
    void *d = hc.dev(n);
    if (hc.err) return hc.err;
    if (fg_generate(d, nullptr, *data, -1, 0, 0, bdmax, hc.stream())) return -5;
    void *s = hc.down(d, n);
    if (hc.sync()) return hc.err;
    memcpy(buf, s, n);
    return 0;
}

extern "C" int rb200_generate_grain_uv(int layout, void *buf, const void *buf_y, const Rb200FilmGrainData *data,
                                       intptr_t uv, int bdmax) {
    if (!buf || !buf_y || !data || layout < RB200_LAYOUT_I420 || layout > RB200_LAYOUT_I444 || uv < 0 || uv > 1)
        return set_error(-22, "generate_grain_uv: bad argument");
    const int subx = layout != RB200_LAYOUT_I444, suby = layout == RB200_LAYOUT_I420;
    const size_t es = bdmax > 255 ? 2 : 1, n = (size_t)GH * GW * es;
    HostCall hc(4 * n);
    const void *dy = hc.up(buf_y, n);
    void *d = hc.dev(n);
    if (hc.err) return hc.err;
    if (fg_generate(d, dy, *data, (int)uv, subx, suby, bdmax, hc.stream())) return -5;
    void *s = hc.down(d, n);
    if (hc.sync()) return hc.err;
    // only the chromaH x chromaW corner is produced; the rest of the caller's buffer is left alone
    const int W = subx ? SUB_GW : GW, H = suby ? SUB_GH : GH;
    for (int y = 0; y < H; y++) memcpy((uint8_t *)buf + (size_t)y * GW * es, (uint8_t *)s + (size_t)y * GW * es, W * es);
    return 0;
}

extern "C" int rb200_generate_scaling(int bitdepth, const uint8_t points[][2], int num, uint8_t *scaling) {
    if (!scaling || num < 0 || num > 14 || (num && !points) || (bitdepth != 8 && bitdepth != 10 && bitdepth != 12))
        return set_error(-22, "generate_scaling: bad argument");
    const size_t n = (size_t)1 << bitdepth;
    HostCall hc(2 * n + 4096);
    uint8_t *d = (uint8_t *)hc.dev(n);
    if (hc.err) return hc.err;
    if (fg_scaling(d, points, num, bitdepth, hc.stream())) return -5;
    void *s = hc.down(d, n);
    if (hc.sync()) return hc.err;
    memcpy(scaling, s, n);
    return 0;
}

static int fg_rows_common(int chroma, int layout, void *dst_row, const void *src_row, ptrdiff_t stride,
                          const Rb200FilmGrainData *data, size_t pw, const uint8_t *scaling, const void *grain_lut,
                          int bh, int row_num, const void *luma_row, ptrdiff_t luma_stride, int uv_pl, int is_id,
                          int bdmax) {
    if (!dst_row || !src_row || !data || !scaling || !grain_lut || pw < 1 || pw > 16384 || bh < 1 || bh > 32 ||
        row_num < 0 || (chroma && (!luma_row || uv_pl < 0 || uv_pl > 1)))
        return set_error(-22, "fg_32x32xn: bad argument");
    const int sx = chroma && layout != RB200_LAYOUT_I444, sy = chroma && layout == RB200_LAYOUT_I420;
    const size_t px = bdmax > 255 ? 2 : 1, es = px;
    const size_t lut_bytes = (size_t)GH * GW * es, sc_bytes = bdmax > 255 ? 4096 : 256;
    const int luma_cols = chroma ? (int)(pw << sx) : 0, luma_rows = chroma ? ((bh - 1) << sy) + 1 : 0;
    const int ncols = (int)((pw + (32 >> sx) - 1) / (32 >> sx));
    HostCall hc(2 * (2 * DevRect::bytes_for(pw * px, bh) + DevRect::bytes_for(luma_cols * px, luma_rows) + lut_bytes +
                     sc_bytes + 2 * ncols) + 8192);
    DevRect rs, rd, rl;
    if (hc.rect_up(rs, src_row, stride, pw * px, bh)) return hc.err;
    if (hc.rect_up(rd, dst_row, stride, pw * px, bh)) return hc.err;
    if (chroma && hc.rect_up(rl, luma_row, luma_stride, luma_cols * px, luma_rows)) return hc.err;
    const uint8_t *d_sc = (const uint8_t *)hc.up(scaling, sc_bytes);
    const void *d_lut = hc.up(grain_lut, lut_bytes);
    const int off_row0 = row_num > 0 ? row_num - 1 : 0, nrows = row_num - off_row0 + 1;
    uint8_t *d_off = (uint8_t *)hc.dev((size_t)nrows * ncols);
    if (hc.err) return hc.err;
    if (fg_offsets(d_off, data->seed, off_row0, nrows, ncols, hc.stream())) return -5;
    FgApplyParams P = fg_params(*data, chroma, uv_pl, sx, sy, is_id);
    P.pw = (int)pw; P.ph = bh; P.row0 = row_num; P.off_row0 = off_row0; P.ncols = ncols; P.luma_w = luma_cols;
    if (fg_apply(rd.dptr, rs.dptr, rs.dpitch, chroma ? rl.dptr : nullptr, chroma ? rl.dpitch : 0, P, d_sc, d_lut, d_off,
                 bdmax, hc.stream())) return -5;
    hc.rect_down(rd);
    if (hc.sync()) return hc.err;
    rd.finish(dst_row);
    return 0;
}

extern "C" int rb200_fgy_32x32xn(void *dst_row, const void *src_row, ptrdiff_t stride, const Rb200FilmGrainData *data,
                                 size_t pw, const uint8_t *scaling, const void *grain_lut, int bh, int row_num,
                                 int bdmax) {
    return fg_rows_common(0, 0, dst_row, src_row, stride, data, pw, scaling, grain_lut, bh, row_num, nullptr, 0, 0, 0, bdmax);
}
extern "C" int rb200_fguv_32x32xn(int layout, void *dst_row, const void *src_row, ptrdiff_t stride,
                                  const Rb200FilmGrainData *data, size_t pw, const uint8_t *scaling,
                                  const void *grain_lut, int bh, int row_num, const void *luma_row,
                                  ptrdiff_t luma_stride, int uv_pl, int is_id, int bdmax) {
    if (layout < RB200_LAYOUT_I420 || layout > RB200_LAYOUT_I444) return set_error(-22, "fguv_32x32xn: bad layout");
    return fg_rows_common(1, layout, dst_row, src_row, stride, data, pw, scaling, grain_lut, bh, row_num, luma_row,
                          luma_stride, uv_pl, is_id, bdmax);
}

namespace {
void gen_y_slot(void *buf, const Rb200FilmGrainData *d, int bd) { if (rb200_generate_grain_y(buf, d, bd)) rb200_report_fatal("generate_grain_y"); }
template <int LAYOUT> void gen_uv_slot(void *buf, const void *by, const Rb200FilmGrainData *d, intptr_t uv, int bd) {
    if (rb200_generate_grain_uv(LAYOUT, buf, by, d, uv, bd)) rb200_report_fatal("generate_grain_uv");
}
void fgy_slot(void *dst, const void *src, ptrdiff_t stride, const Rb200FilmGrainData *d, size_t pw, const uint8_t *sc,
              const void *lut, int bh, int row, int bd) {
    if (rb200_fgy_32x32xn(dst, src, stride, d, pw, sc, lut, bh, row, bd)) rb200_report_fatal("fgy_32x32xn");
}
template <int LAYOUT> void fguv_slot(void *dst, const void *src, ptrdiff_t stride, const Rb200FilmGrainData *d, size_t pw,
                                     const uint8_t *sc, const void *lut, int bh, int row, const void *luma,
                                     ptrdiff_t ls, int uv, int is_id, int bd) {
    if (rb200_fguv_32x32xn(LAYOUT, dst, src, stride, d, pw, sc, lut, bh, row, luma, ls, uv, is_id, bd)) rb200_report_fatal("fguv_32x32xn");
}
}  // namespace

extern "C" void rb200_film_grain_dsp_init(Rb200FilmGrainDSPContext *c, int bpc) {
    (void)bpc;
    c->generate_grain_y = &gen_y_slot;
    c->generate_grain_uv[0] = &gen_uv_slot<RB200_LAYOUT_I420>;
    c->generate_grain_uv[1] = &gen_uv_slot<RB200_LAYOUT_I422>;
    c->generate_grain_uv[2] = &gen_uv_slot<RB200_LAYOUT_I444>;
    c->fgy_32x32xn = &fgy_slot;
    c->fguv_32x32xn[0] = &fguv_slot<RB200_LAYOUT_I420>;
    c->fguv_32x32xn[1] = &fguv_slot<RB200_LAYOUT_I422>;
    c->fguv_32x32xn[2] = &fguv_slot<RB200_LAYOUT_I444>;
}
