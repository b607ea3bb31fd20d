// CDEF (constrained directional enhancement filter) for sm_100a.
//
// Replaces Rav1dCdefDSPContext {dir, fb[3]} (src/cdef.rs:35-56; cdef_find_dir_rust
// src/cdef.rs:921, cdef_filter_block_c src/cdef.rs:668 == src/cdef_tmpl.c:104-300) and,
// at frame level, the per-sbrow driver rav1d_cdef_brow (src/cdef_apply.rs:159-507 ==
// src/cdef_apply_tmpl.c:98-309).
//
// Frame kernel: one CTA per 64x64 luma area (one cdef_idx unit).  The pre-CDEF
// (deblocked) pixels of the area plus a 2-pixel halo are staged in shared memory as
// int16 with the reference's INT16_MIN sentinel outside the frame; 64 threads
// each run the direction search of one 8x8 block out of registers; then all 256
// threads filter, out of place, so every tap reads pre-CDEF data without the
// reference's line/column backups.  Blocks that are skipped are
// copied through.  Chroma planes reuse the tile buffer and the luma directions.
#include "common.cuh"
#include "tables.cuh"

namespace rb200 {

constexpr int CDEF_T = 64;               // luma tile edge
constexpr int CDEF_PITCH = CDEF_T + 4;   // 68
constexpr int16_t CDEF_SENTINEL = -32768;

__device__ __forceinline__ int cdef_constrain(int diff, int threshold, int shift) {
    const int adiff = diff < 0 ? -diff : diff;
    const int v = imin(adiff, imax(0, threshold - (adiff >> shift)));
    return diff < 0 ? -v : v;
}

// One pixel of cdef_filter_block_c.  t points at the pixel inside an int16 tile of row pitch `pitch`.
__device__ __forceinline__ int cdef_filter_px(const int16_t *t, int pitch, int pri, int sec, int dir, int damping,
                                              int bdmin8) {
    const int px = t[0];
    int sum = 0;
    if (pri) {
        const int pri_tap = 4 - ((pri >> bdmin8) & 1);
        const int pri_shift = imax(0, damping - ulog2(pri));
        if (sec) {
            const int sec_shift = damping - ulog2(sec);
            int mx = px, mn = px;
            int ptk = pri_tap;
#pragma unroll
            for (int k = 0; k < 2; k++) {
                int dy, dx;
                tab::cdef_dir_off(dir, k, dy, dx);
                const int o1 = dy * pitch + dx;
                const int p0 = t[o1], p1 = t[-o1];
                sum += ptk * cdef_constrain(p0 - px, pri, pri_shift);
                sum += ptk * cdef_constrain(p1 - px, pri, pri_shift);
                ptk = (ptk & 3) | 2;
                mn = (int)umin((unsigned)p0, (unsigned)mn); mx = imax(p0, mx);
                mn = (int)umin((unsigned)p1, (unsigned)mn); mx = imax(p1, mx);
                tab::cdef_dir_off((dir + 2) & 7, k, dy, dx);
                const int o2 = dy * pitch + dx;
                tab::cdef_dir_off((dir + 6) & 7, k, dy, dx);
                const int o3 = dy * pitch + dx;
                const int s0 = t[o2], s1 = t[-o2], s2 = t[o3], s3 = t[-o3];
                const int sec_tap = 2 - k;
                sum += sec_tap * cdef_constrain(s0 - px, sec, sec_shift);
                sum += sec_tap * cdef_constrain(s1 - px, sec, sec_shift);
                sum += sec_tap * cdef_constrain(s2 - px, sec, sec_shift);
                sum += sec_tap * cdef_constrain(s3 - px, sec, sec_shift);
                mn = (int)umin((unsigned)s0, (unsigned)mn); mx = imax(s0, mx);
                mn = (int)umin((unsigned)s1, (unsigned)mn); mx = imax(s1, mx);
                mn = (int)umin((unsigned)s2, (unsigned)mn); mx = imax(s2, mx);
                mn = (int)umin((unsigned)s3, (unsigned)mn); mx = imax(s3, mx);
            }
            return iclip(px + ((sum - (sum < 0) + 8) >> 4), mn, mx);
        }
        int ptk = pri_tap;
#pragma unroll
        for (int k = 0; k < 2; k++) {
            int dy, dx;
            tab::cdef_dir_off(dir, k, dy, dx);
            const int o = dy * pitch + dx;
            sum += ptk * cdef_constrain(t[o] - px, pri, pri_shift);
            sum += ptk * cdef_constrain(t[-o] - px, pri, pri_shift);
            ptk = (ptk & 3) | 2;
        }
        return px + ((sum - (sum < 0) + 8) >> 4);
    }
    const int sec_shift = damping - ulog2(sec);
#pragma unroll
    for (int k = 0; k < 2; k++) {
        int dy, dx;
        tab::cdef_dir_off((dir + 2) & 7, k, dy, dx);
        const int o1 = dy * pitch + dx;
        tab::cdef_dir_off((dir + 6) & 7, k, dy, dx);
        const int o2 = dy * pitch + dx;
        const int sec_tap = 2 - k;
        sum += sec_tap * cdef_constrain(t[o1] - px, sec, sec_shift);
        sum += sec_tap * cdef_constrain(t[-o1] - px, sec, sec_shift);
        sum += sec_tap * cdef_constrain(t[o2] - px, sec, sec_shift);
        sum += sec_tap * cdef_constrain(t[-o2] - px, sec, sec_shift);
    }
    return px + ((sum - (sum < 0) + 8) >> 4);
}

// cdef_find_dir for one 8x8 block held in an int16 tile; fully unrolled so that the 90
// partial sums live in registers.  src/cdef_tmpl.c:223-300
__device__ __forceinline__ int cdef_find_dir(const int16_t *t, int pitch, int bdmin8, unsigned *var) {
    int hv0[8] = {0}, hv1[8] = {0}, dg0[15] = {0}, dg1[15] = {0}, al0[11] = {0}, al1[11] = {0}, al2[11] = {0},
        al3[11] = {0};
#pragma unroll
    for (int y = 0; y < 8; y++) {
#pragma unroll
        for (int x = 0; x < 8; x++) {
            const int px = ((int)t[y * pitch + x] >> bdmin8) - 128;
            dg0[y + x] += px;
            al0[y + (x >> 1)] += px;
            hv0[y] += px;
            al1[3 + y - (x >> 1)] += px;
            dg1[7 + y - x] += px;
            al2[3 - (y >> 1) + x] += px;
            hv1[x] += px;
            al3[(y >> 1) + x] += px;
        }
    }
    unsigned cost[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int n = 0; n < 8; n++) {
        cost[2] += hv0[n] * hv0[n];
        cost[6] += hv1[n] * hv1[n];
    }
    cost[2] *= 105; cost[6] *= 105;
    constexpr int div_table[7] = {840, 420, 280, 210, 168, 140, 120};
#pragma unroll
    for (int n = 0; n < 7; n++) {
        const int d = div_table[n];
        cost[0] += (dg0[n] * dg0[n] + dg0[14 - n] * dg0[14 - n]) * d;
        cost[4] += (dg1[n] * dg1[n] + dg1[14 - n] * dg1[14 - n]) * d;
    }
    cost[0] += dg0[7] * dg0[7] * 105;
    cost[4] += dg1[7] * dg1[7] * 105;
    auto alt_cost = [&](const int *a) -> unsigned {
        unsigned c = 0;
#pragma unroll
        for (int m = 0; m < 5; m++) c += a[3 + m] * a[3 + m];
        c *= 105;
#pragma unroll
        for (int m = 0; m < 3; m++) c += (a[m] * a[m] + a[10 - m] * a[10 - m]) * div_table[2 * m + 1];
        return c;
    };
    cost[1] = alt_cost(al0); cost[3] = alt_cost(al1); cost[5] = alt_cost(al2); cost[7] = alt_cost(al3);
    int best = 0;
    unsigned best_cost = cost[0];
#pragma unroll
    for (int n = 1; n < 8; n++)
        if (cost[n] > best_cost) { best_cost = cost[n]; best = n; }
    unsigned opp = 0;
#pragma unroll
    for (int n = 0; n < 8; n++) if (n == (best ^ 4)) opp = cost[n];
    *var = (best_cost - opp) >> 10;
    return best;
}

__device__ __forceinline__ int cdef_adjust_strength(int strength, unsigned var) {  // src/cdef_apply_tmpl.c:92-96
    if (!var) return 0;
    const int i = (var >> 6) ? imin(ulog2(var >> 6), 12) : 0;
    return (strength * (4 + i) + 8) >> 4;
}


// ---------------------------------------------------------------- frame level
// Two launches per frame:
//  (1) cdef_dir_frame_kernel: one thread per 8x8 luma block reads the block straight from
//      global memory (rows are coalesced across the 32 blocks of a warp), runs the direction
//      search out of registers and resolves everything rav1d_cdef_brow decides per block
//      (cdef_idx, noskip bits, strengths, adjust_strength, chroma direction) into an 8-byte
//      record;
//  (2) cdef_filter_frame_kernel: one CTA per 64x64 luma area and its chroma.  The pre-CDEF
//      pixels (+2 halo) are staged in shared memory twice, the second copy shifted by one
//      pixel, so that any horizontally adjacent pixel PAIR is one aligned 32-bit word in one
//      of the copies.  A thread filters one row of one 8x8 block, two pixels per register:
//      the constrain() of both pixels is evaluated with the packed 16x2 min/max/add
//      instructions of sm_100a (VIMNMX.S16x2, VIADDMNMX.S16x2, VIMNMX3), the weighted sum
//      with one IMAD per tap on values biased to stay positive in both halves.
//      Blocks on the frame border (the only ones that can see the reference's INT16_MIN
//      "unavailable" sentinel) take the scalar path.
struct CdefBlk {  // per 8x8 luma block, written by (1), read by (2)
    uint8_t y_pri, y_sec, uv_pri, uv_sec;  // y_pri already through adjust_strength
    uint8_t dir, uvdir, do_y, do_uv;
};

template <typename BD>
__global__ void __launch_bounds__(128)
cdef_dir_frame_kernel(Rb200Planes src, CdefFrameParams P, const Rb200Av1Filter *__restrict__ masks,
                      CdefBlk *__restrict__ out, int nbx, int nby, int by_first, int by_end) {
    using pixel = typename BD::pixel;
    const int bx = blockIdx.x * 32 + threadIdx.x, by = by_first + blockIdx.y * 4 + threadIdx.y;
    if (bx >= nbx || by >= by_end) return;
    CdefBlk b = {};
    const Rb200Av1Filter &m = masks[(by >> 4) * P.sb128w + (bx >> 4)];
    const int cdef_idx = m.cdef_idx[(((by >> 3) & 1) << 1) + ((bx >> 3) & 1)];
    const int y_lvl = cdef_idx >= 0 ? P.y_strength[cdef_idx] : 0;
    const int uv_lvl = cdef_idx >= 0 ? P.uv_strength[cdef_idx] : 0;
    if (cdef_idx >= 0 && (y_lvl || uv_lvl)) {
        const uint16_t *row = m.noskip_mask[by & 15];
        const unsigned nm = ((unsigned)row[1] << 16) | row[0];
        if (nm & (3u << ((bx & 15) * 2))) {
            const int y_pri = (y_lvl >> 2) << P.bdmin8;
            int y_sec = y_lvl & 3; y_sec += y_sec == 3; y_sec <<= P.bdmin8;
            const int uv_pri = (uv_lvl >> 2) << P.bdmin8;
            int uv_sec = uv_lvl & 3; uv_sec += uv_sec == 3; uv_sec <<= P.bdmin8;
            int dir = 0; unsigned var = 0;
            if (y_pri || uv_pri) {
                // 8 rows of 8 pixels -> int16 registers
                int16_t t[64];
                const uint8_t *p = (const uint8_t *)src.data[0] + (int64_t)(by * 8) * src.stride[0] + (int64_t)bx * 8 * sizeof(pixel);
#pragma unroll
                for (int y = 0; y < 8; y++) {
                    if (BD::hbd) {
                        const uint4 v = *(const uint4 *)(p + (int64_t)y * src.stride[0]);
                        t[y * 8 + 0] = (int16_t)(v.x & 0xffff); t[y * 8 + 1] = (int16_t)(v.x >> 16);
                        t[y * 8 + 2] = (int16_t)(v.y & 0xffff); t[y * 8 + 3] = (int16_t)(v.y >> 16);
                        t[y * 8 + 4] = (int16_t)(v.z & 0xffff); t[y * 8 + 5] = (int16_t)(v.z >> 16);
                        t[y * 8 + 6] = (int16_t)(v.w & 0xffff); t[y * 8 + 7] = (int16_t)(v.w >> 16);
                    } else {
                        const uint2 v = *(const uint2 *)(p + (int64_t)y * src.stride[0]);
#pragma unroll
                        for (int x = 0; x < 4; x++) {
                            t[y * 8 + x] = (int16_t)((v.x >> (8 * x)) & 0xff);
                            t[y * 8 + 4 + x] = (int16_t)((v.y >> (8 * x)) & 0xff);
                        }
                    }
                }
                dir = cdef_find_dir(t, 8, P.bdmin8, &var);
            }
            if (y_pri) {
                const int adj = cdef_adjust_strength(y_pri, var);
                if (adj || y_sec) { b.do_y = 1; b.y_pri = (uint8_t)adj; b.y_sec = (uint8_t)y_sec; b.dir = (uint8_t)dir; }
            } else if (y_sec) {
                b.do_y = 1; b.y_pri = 0; b.y_sec = (uint8_t)y_sec; b.dir = 0;
            }
            if (uv_lvl) {
                b.do_uv = 1; b.uv_pri = (uint8_t)uv_pri; b.uv_sec = (uint8_t)uv_sec;
                const int d422 = (0x66654207 >> (4 * dir)) & 7;  // 4:2:2 remap, src/cdef_apply_tmpl.c:113-115
                b.uvdir = (uint8_t)(uv_pri ? (P.layout_422 ? d422 : dir) : 0);
            }
        }
    }
    out[by * nbx + bx] = b;
}

constexpr int CDEF_TP = 74;                        // tile pitch in pixels: 37 words, odd -> rows spread over banks
constexpr int CDEF_TROWS = 68;
constexpr int CDEF_COPY = CDEF_TROWS * CDEF_TP;    // elements per copy
constexpr int CDEF_X0 = 4;                         // tile column of the area's first pixel (even; 4 = one load group)
constexpr unsigned CDEF_BIAS = 256;                // per-half bias of the constrained differences (|c| <= 240)

// Stage rows y0-2 .. y0+th+1, columns x0-4 .. x0+tw+3 of `plane` (4-pixel groups, vector loads)
// into copy A (tile[]) and the one-pixel-shifted copy B (tile[CDEF_COPY + i] = A[i + 1]).
template <typename BD>
__device__ __forceinline__ void cdef_stage2(int16_t *tile, const uint8_t *plane, int64_t stride, int x0, int y0, int tw,
                                            int th, int fw, int fh) {
    using pixel = typename BD::pixel;
    const int groups = (tw + 8) >> 2, rows = th + 4;
    constexpr int NIT = (68 * 18 + 255) / 256;   // 256 threads
    uint2 q[NIT];
    // loads first, then stores: NIT loads in flight per thread
#pragma unroll
    for (int it = 0; it < NIT; it++) {
        const int i = threadIdx.x + it * 256;
        q[it] = make_uint2(0, 0);
        if (i < rows * groups) {
            const int r = i / groups, g = i - r * groups;
            const int y = y0 + r - 2, x = x0 - 4 + g * 4;
            if (y >= 0 && y < fh && x >= 0 && x < fw) {  // fw is a multiple of 8: a group is inside or outside as a whole
                const uint8_t *p = plane + (int64_t)y * stride + (int64_t)x * sizeof(pixel);
                if (BD::hbd) q[it] = *(const uint2 *)p;
                else q[it].x = *(const unsigned *)p;
            } else {
                q[it].y = 0xffffffffu;   // marks "outside the frame" (a pixel never has the top bit set)
            }
        }
    }
#pragma unroll
    for (int it = 0; it < NIT; it++) {
        const int i = threadIdx.x + it * 256;
        if (i >= rows * groups) continue;
        const int r = i / groups, g = i - r * groups;
        int v[4];
        if (q[it].y == 0xffffffffu) {
            v[0] = v[1] = v[2] = v[3] = CDEF_SENTINEL;
        } else if (BD::hbd) {
            v[0] = q[it].x & 0xffff; v[1] = q[it].x >> 16; v[2] = q[it].y & 0xffff; v[3] = q[it].y >> 16;
        } else {
            v[0] = q[it].x & 0xff; v[1] = (q[it].x >> 8) & 0xff; v[2] = (q[it].x >> 16) & 0xff; v[3] = q[it].x >> 24;
        }
        int16_t *a = tile + r * CDEF_TP + g * 4;
        // the odd word pitch makes rows only 4-byte aligned
        ((unsigned *)a)[0] = (unsigned)(v[0] & 0xffff) | ((unsigned)v[1] << 16);
        ((unsigned *)a)[1] = (unsigned)(v[2] & 0xffff) | ((unsigned)v[3] << 16);
        int16_t *bcopy = a + CDEF_COPY - 1;
        if (g) bcopy[0] = (int16_t)v[0];
        bcopy[1] = (int16_t)v[1]; bcopy[2] = (int16_t)v[2]; bcopy[3] = (int16_t)v[3];
    }
}

// byte offset (from the tile base) of the aligned word holding pixels (s, s + 1) of a row-major element index s
__device__ __forceinline__ int cdef_pair_addr(int s) { return (s & 1) ? 2 * (CDEF_COPY + s - 1) : 2 * s; }

// Filter NP pixel pairs of one block row.  `s0`: element index of the first pixel (even).
// Returns the filtered pixels in out[2 * NP].  Interior blocks only (no sentinel in reach).
template <int NP>
__device__ __forceinline__ void cdef_row_packed(const int16_t *tile, int s0, int pri, int sec, int dir, int damping,
                                                int bdmin8, int *out) {
    const char *base = (const char *)tile;
    unsigned px2[NP], cst[NP], sum[NP], mn[NP], mx[NP];
#pragma unroll
    for (int j = 0; j < NP; j++) {
        px2[j] = *(const unsigned *)(base + 2 * (s0 + 2 * j));
        cst[j] = __vneg2(px2[j]);   // -px per half
        sum[j] = 0; mn[j] = px2[j]; mx[j] = px2[j];
    }
    const unsigned B2 = CDEF_BIAS * 0x10001u;
    int ktot = 0;
    auto tap = [&](int off_el, int thr, int shift, int w, bool track) {
        const unsigned m = (0xffffu >> shift) * 0x10001u;
        const unsigned thr1 = (unsigned)(thr + 1) * 0x10001u;
        const unsigned bmt = (unsigned)(CDEF_BIAS - thr) * 0x10001u;
#pragma unroll
        for (int sgn = 0; sgn < 2; sgn++) {
            const int a0 = cdef_pair_addr(s0 + (sgn ? -off_el : off_el));
#pragma unroll
            for (int j = 0; j < NP; j++) {
                const unsigned p2 = *(const unsigned *)(base + a0 + 4 * j);
                const unsigned d = __vadd2(p2, cst[j]);                 // p - px
                const unsigned a = __vmaxs2(d, __vneg2(d));             // |d|
                const unsigned x = a >> shift;
                const unsigned sft = x & m;                             // |d| >> shift, per half
                const unsigned t = __viaddmax_s16x2(~sft, thr1, 0u);    // max(thr - s, 0)
                const unsigned ntb = __viaddmin_s16x2(sft, bmt, B2);    // BIAS - t
                const unsigned c = __viaddmax_s16x2(__vmins2(d, t), B2, ntb);  // BIAS + clamp(d, -t, t)
                sum[j] += (unsigned)w * c;
                if (track) { mn[j] = __vminu2(mn[j], p2); mx[j] = __vmaxs2(mx[j], p2); }
            }
        }
        ktot += 2 * w;
    };
    const bool both = pri && sec;
    if (pri) {
        const int pri_tap = 4 - ((pri >> bdmin8) & 1);
        const int pri_shift = imax(0, damping - ulog2(pri));
        int dy, dx;
        tab::cdef_dir_off(dir, 0, dy, dx);
        tap(dy * CDEF_TP + dx, pri, pri_shift, pri_tap, both);
        tab::cdef_dir_off(dir, 1, dy, dx);
        tap(dy * CDEF_TP + dx, pri, pri_shift, (pri_tap & 3) | 2, both);
    }
    if (sec) {
        const int sec_shift = damping - ulog2(sec);
#pragma unroll
        for (int k = 0; k < 2; k++) {
            int dy, dx;
            tab::cdef_dir_off((dir + 2) & 7, k, dy, dx);
            tap(dy * CDEF_TP + dx, sec, sec_shift, 2 - k, both);
            tab::cdef_dir_off((dir + 6) & 7, k, dy, dx);
            tap(dy * CDEF_TP + dx, sec, sec_shift, 2 - k, both);
        }
    }
    const int kb = ktot * (int)CDEF_BIAS;
#pragma unroll
    for (int j = 0; j < NP; j++) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int s = (int)((sum[j] >> (16 * h)) & 0xffff) - kb;
            const int px = (int)((px2[j] >> (16 * h)) & 0xffff);
            int v = px + ((s - (s < 0) + 8) >> 4);
            if (both) v = iclip(v, (int)((mn[j] >> (16 * h)) & 0xffff), (int)((mx[j] >> (16 * h)) & 0xffff));
            out[2 * j + h] = v;
        }
    }
}

// One plane of the area.  bw/bh: block size in this plane (8 or 4); tile staged already.
template <typename BD, int BW>
__device__ __forceinline__ void cdef_filter_plane(const int16_t *tile, const CdefBlk *blk, uint8_t *dbase, int64_t dstride,
                                                  int px0, int py0, int tw, int th, int bh_log2, bool chroma, int damping,
                                                  int bdmin8, int nbx, int nby, int gbx0, int gby0) {
    using pixel = typename BD::pixel;
    constexpr int NP = BW / 2;
    const int bh = 1 << bh_log2;
    const int n_tasks = 64 << bh_log2;   // 64 blocks x rows
    for (int t = threadIdx.x; t < n_tasks; t += blockDim.x) {
        // a warp covers 8 blocks along x times 4 rows: conflict-free centre loads with the odd word pitch
        const int lane = t & 31, grp = t >> 5;
        const int bxl = lane & 7, rlow = lane >> 3;
        const int rows_per_blk_grp = bh >> 2;                  // 4-row groups per block (1 or 2)
        const int byl = grp / rows_per_blk_grp, r = rlow + 4 * (grp - byl * rows_per_blk_grp);
        const int x = bxl * BW, y = byl * bh + r;
        if (x >= tw || y >= th) continue;
        const CdefBlk b = blk[byl * 8 + bxl];
        const int s0 = (2 + y) * CDEF_TP + CDEF_X0 + x;
        int out[BW];
        const bool on = chroma ? b.do_uv : b.do_y;
        const int pri = chroma ? b.uv_pri : b.y_pri, sec = chroma ? b.uv_sec : b.y_sec, dir = chroma ? b.uvdir : b.dir;
        const int gbx = gbx0 + bxl, gby = gby0 + byl;
        const bool border = gbx == 0 || gby == 0 || gbx == nbx - 1 || gby == nby - 1;
        if (!on) {
#pragma unroll
            for (int i = 0; i < BW; i++) out[i] = tile[s0 + i];
        } else if (border) {
#pragma unroll
            for (int i = 0; i < BW; i++) out[i] = cdef_filter_px(tile + s0 + i, CDEF_TP, pri, sec, dir, damping, bdmin8);
        } else {
            cdef_row_packed<NP>(tile, s0, pri, sec, dir, damping, bdmin8, out);
        }
        pixel *d = (pixel *)(dbase + (int64_t)(py0 + y) * dstride) + px0 + x;
        if (BD::hbd) {
            if (BW == 8) *(uint4 *)d = make_uint4(out[0] | (out[1] << 16), out[2] | (out[3] << 16), out[4] | (out[5] << 16), out[6] | (out[7] << 16));
            else *(uint2 *)d = make_uint2(out[0] | (out[1] << 16), out[2] | (out[3] << 16));
        } else {
            if (BW == 8) *(uint2 *)d = make_uint2(out[0] | (out[1] << 8) | (out[2] << 16) | (out[3] << 24), out[4] | (out[5] << 8) | (out[6] << 16) | (out[7] << 24));
            else *(unsigned *)d = out[0] | (out[1] << 8) | (out[2] << 16) | (out[3] << 24);
        }
    }
}

template <typename BD>
__global__ void __launch_bounds__(256)
cdef_filter_frame_kernel(Rb200Planes src, Rb200Planes dst, CdefFrameParams P, const CdefBlk *__restrict__ blocks, int nbx,
                         int nby, int tile_row_first) {
    __shared__ __align__(16) int16_t tile[2 * CDEF_COPY + 8];
    __shared__ CdefBlk blk[64];
    const int sbx = blockIdx.x, sby = tile_row_first + blockIdx.y;
    const int x0 = sbx * 64, y0 = sby * 64;
    const int fw = P.bw * 4, fh = P.bh * 4;
    const int tw = imin(64, fw - x0), th = imin(64, fh - y0);
    if (threadIdx.x < 64) {
        const int bx = sbx * 8 + (threadIdx.x & 7), by = sby * 8 + (threadIdx.x >> 3);
        CdefBlk b = {};
        if (bx < nbx && by < nby) b = blocks[by * nbx + bx];
        blk[threadIdx.x] = b;
    }
    cdef_stage2<BD>(tile, (const uint8_t *)src.data[0], src.stride[0], x0, y0, tw, th, fw, fh);
    __syncthreads();
    cdef_filter_plane<BD, 8>(tile, blk, (uint8_t *)dst.data[0], dst.stride[0], x0, y0, tw, th, 3, false, P.damping, P.bdmin8,
                             nbx, nby, sbx * 8, sby * 8);
    for (int p = 1; p < P.n_planes; p++) {
        __syncthreads();
        const int cx0 = x0 >> P.ss_hor, cy0 = y0 >> P.ss_ver;
        const int ctw = tw >> P.ss_hor, cth = th >> P.ss_ver, cfw = fw >> P.ss_hor, cfh = fh >> P.ss_ver;
        cdef_stage2<BD>(tile, (const uint8_t *)src.data[p], src.stride[p], cx0, cy0, ctw, cth, cfw, cfh);
        __syncthreads();
        if (P.ss_hor)
            cdef_filter_plane<BD, 4>(tile, blk, (uint8_t *)dst.data[p], dst.stride[p], cx0, cy0, ctw, cth, 3 - P.ss_ver, true,
                                     P.damping - 1, P.bdmin8, nbx, nby, sbx * 8, sby * 8);
        else
            cdef_filter_plane<BD, 8>(tile, blk, (uint8_t *)dst.data[p], dst.stride[p], cx0, cy0, ctw, cth, 3, true,
                                     P.damping - 1, P.bdmin8, nbx, nby, sbx * 8, sby * 8);
    }
}

// ---- per-call kernels
template <typename BD>
__global__ void cdef_dir_kernel(const uint8_t *src, int64_t stride, int bdmax, int *out /* dir, var */) {
    using pixel = typename BD::pixel;
    __shared__ int16_t t[64];
    if (threadIdx.x < 64) t[threadIdx.x] = (int16_t)((const pixel *)(src + (int64_t)(threadIdx.x >> 3) * stride))[threadIdx.x & 7];
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned var;
        out[0] = cdef_find_dir(t, 8, BD::hbd ? bpc_from_max(bdmax) - 8 : 0, &var);
        out[1] = (int)var;
    }
}

// padding() + filter of one w x h block.  src/cdef_tmpl.c:55-102
template <typename BD>
__global__ void cdef_fb_kernel(uint8_t *dst, int64_t stride, const uint8_t *left /* [h][2] */,
                               const uint8_t *top /* 2 rows x (w+4), from x=-2 */, const uint8_t *bottom, int w, int h,
                               int pri, int sec, int dir, int damping, unsigned edges, int bdmax) {
    using pixel = typename BD::pixel;
    __shared__ int16_t tile[12 * 12];
    const int cols = w + 4, rows = h + 4;
    for (int i = threadIdx.x; i < rows * cols; i += blockDim.x) {
        const int r = i / cols, c = i - r * cols;
        const int y = r - 2, x = c - 2;
        bool avail = true;
        if (y < 0 && !(edges & RB200_CDEF_HAVE_TOP)) avail = false;
        if (y >= h && !(edges & RB200_CDEF_HAVE_BOTTOM)) avail = false;
        if (x < 0 && !(edges & RB200_CDEF_HAVE_LEFT)) avail = false;
        if (x >= w && !(edges & RB200_CDEF_HAVE_RIGHT)) avail = false;
        int16_t v = CDEF_SENTINEL;
        if (avail) {
            if (y < 0) v = (int16_t)((const pixel *)top)[(y + 2) * cols + c];
            else if (y >= h) v = (int16_t)((const pixel *)bottom)[(y - h) * cols + c];
            else if (x < 0) v = (int16_t)((const pixel *)left)[y * 2 + 2 + x];
            else v = (int16_t)((const pixel *)(dst + (int64_t)y * stride))[x];  // x >= w: right neighbour, in dst rect
        }
        tile[r * 12 + c] = v;
    }
    __syncthreads();
    const int bdmin8 = BD::hbd ? bpc_from_max(bdmax) - 8 : 0;
    for (int i = threadIdx.x; i < w * h; i += blockDim.x) {
        const int r = i / w, c = i - r * w;
        const int v = cdef_filter_px(tile + (2 + r) * 12 + 2 + c, 12, pri, sec, dir, damping, bdmin8);
        ((pixel *)(dst + (int64_t)r * stride))[c] = (pixel)v;
    }
}

// t0 / t1: 64-row tile rows to produce (whole picture: 0, ceil(height / 64))
int cdef_frame_launch(const Rb200Planes &src, const Rb200Planes &dst, const CdefFrameParams &P,
                      const Rb200Av1Filter *masks, void *blk_scratch, int bdmax, cudaStream_t st, int t0, int t1) {
    const int nbx = P.bw >> 1, nby = P.bh >> 1;
    CdefBlk *blocks = (CdefBlk *)blk_scratch;
    const int by0 = t0 * 8, by1 = imin(t1 * 8, nby);
    if (by1 <= by0) return 0;
    dim3 g1((nbx + 31) / 32, (by1 - by0 + 3) / 4), b1(32, 4);
    dim3 grid((P.bw * 4 + 63) / 64, t1 - t0);
    if (bdmax > 255) {
        cdef_dir_frame_kernel<BD16><<<g1, b1, 0, st>>>(src, P, masks, blocks, nbx, nby, by0, by1);
        cdef_filter_frame_kernel<BD16><<<grid, 256, 0, st>>>(src, dst, P, blocks, nbx, nby, t0);
    } else {
        cdef_dir_frame_kernel<BD8><<<g1, b1, 0, st>>>(src, P, masks, blocks, nbx, nby, by0, by1);
        cdef_filter_frame_kernel<BD8><<<grid, 256, 0, st>>>(src, dst, P, blocks, nbx, nby, t0);
    }
    RB_LAUNCH_CHECK();
    return 0;
}

}  // namespace rb200

using namespace rb200;

extern "C" int rb200_cdef_dir(const void *src, ptrdiff_t stride, unsigned *var, int bdmax, int *dir_out) {
    if (!src || !var || !dir_out) return set_error(-22, "cdef_dir: bad argument");
    const size_t px = bdmax > 255 ? 2 : 1;
    HostCall hc(4096);
    DevRect rect;
    if (hc.rect_up(rect, src, stride, 8 * px, 8)) return hc.err;
    int *d_out = (int *)hc.dev(8);
    if (hc.err) return hc.err;
    if (bdmax > 255) cdef_dir_kernel<BD16><<<1, 64, 0, hc.stream()>>>(rect.dptr, rect.dpitch, bdmax, d_out);
    else cdef_dir_kernel<BD8><<<1, 64, 0, hc.stream()>>>(rect.dptr, rect.dpitch, bdmax, d_out);
    int *h_out = (int *)hc.down(d_out, 8);
    if (hc.sync()) return hc.err;
    *dir_out = h_out[0];
    *var = (unsigned)h_out[1];
    return 0;
}

extern "C" int rb200_cdef_fb(int idx, void *dst, ptrdiff_t stride, const void *left, const void *top,
                             const void *bottom, int pri, int sec, int dir, int damping, uint32_t edges, int bdmax) {
    if (idx < 0 || idx > 2 || !dst || (!pri && !sec) || dir < 0 || dir > 7) return set_error(-22, "cdef_fb: bad argument");
    const int w = idx == 0 ? 8 : 4, h = idx == 2 ? 4 : 8;
    const size_t px = bdmax > 255 ? 2 : 1;
    HostCall hc(8192);
    // destination rectangle incl. the 2 right-neighbour columns when they exist
    const int cols = w + ((edges & RB200_CDEF_HAVE_RIGHT) ? 2 : 0);
    DevRect rect;
    if (hc.rect_up(rect, dst, stride, cols * px, h)) return hc.err;
    // top / bottom: 2 rows of (w + 4) pixels starting at x = -2, gathered where the reference may read them
    uint8_t rows2[2][2][12 * 2] = {};
    const int xs = (edges & RB200_CDEF_HAVE_LEFT) ? -2 : 0, xe = w + ((edges & RB200_CDEF_HAVE_RIGHT) ? 2 : 0);
    for (int tb = 0; tb < 2; tb++) {
        const uint8_t *p = (const uint8_t *)(tb ? bottom : top);
        if (!(edges & (tb ? RB200_CDEF_HAVE_BOTTOM : RB200_CDEF_HAVE_TOP)) || !p) continue;
        for (int r = 0; r < 2; r++)
            memcpy(&rows2[tb][r][(xs + 2) * px], p + (int64_t)r * stride + (int64_t)xs * (int64_t)px, (size_t)(xe - xs) * px);
    }
    uint8_t lbuf[8 * 2 * 2] = {};
    if ((edges & RB200_CDEF_HAVE_LEFT) && left) memcpy(lbuf, left, (size_t)h * 2 * px);
    // repack rows2 to pitch (w+4) pixels
    uint8_t tbuf[2][2 * 12 * 2];
    for (int tb = 0; tb < 2; tb++)
        for (int r = 0; r < 2; r++) memcpy(&tbuf[tb][(size_t)r * (w + 4) * px], rows2[tb][r], (size_t)(w + 4) * px);
    const uint8_t *d_left = (const uint8_t *)hc.up(lbuf, sizeof(lbuf));
    const uint8_t *d_top = (const uint8_t *)hc.up(tbuf[0], sizeof(tbuf[0]));
    const uint8_t *d_bot = (const uint8_t *)hc.up(tbuf[1], sizeof(tbuf[1]));
    if (hc.err) return hc.err;
    if (bdmax > 255) cdef_fb_kernel<BD16><<<1, 64, 0, hc.stream()>>>(rect.dptr, rect.dpitch, d_left, d_top, d_bot, w, h, pri, sec, dir, damping, edges, bdmax);
    else cdef_fb_kernel<BD8><<<1, 64, 0, hc.stream()>>>(rect.dptr, rect.dpitch, d_left, d_top, d_bot, w, h, pri, sec, dir, damping, edges, bdmax);
    // only the w x h block is written back
    DevRect out = rect;
    hc.rect_down(rect);
    if (hc.sync()) return hc.err;
    rect.row_bytes = w * px;
    rect.finish(dst);
    (void)out;
    return 0;
}

namespace {
int dir_slot(const void *src, ptrdiff_t stride, unsigned *var, int bd) {
    int d = 0;
    if (rb200_cdef_dir(src, stride, var, bd, &d)) rb200_report_fatal("cdef_dir");
    return d;
}
template <int IDX>
void fb_slot(void *dst, ptrdiff_t stride, const void *left, const void *top, const void *bottom, int pri, int sec,
             int dir, int damping, uint32_t edges, int bd) {
    if (rb200_cdef_fb(IDX, dst, stride, left, top, bottom, pri, sec, dir, damping, edges, bd)) rb200_report_fatal("cdef_fb");
}
}  // namespace

extern "C" void rb200_cdef_dsp_init(Rb200CdefDSPContext *c, int bpc) {
    (void)bpc;
    c->dir = &dir_slot;
    c->fb[0] = &fb_slot<0>; c->fb[1] = &fb_slot<1>; c->fb[2] = &fb_slot<2>;
}
