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
#include "tma.cuh"
#include "stages.cuh"

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

// Direction search of one 8x8 block held in an int16 tile, as the AV1 specification states it (7.15.2, "CDEF direction
// process"): every pixel is added to one line sum per direction -- line(d, i, j) below is the specification's index
// into partial[d][] -- and a direction's cost is the sum of its squared line sums, each weighted by 840 / (number of pixels
// on the line) (Div_Table).  Fully unrolled: the indices are compile-time constants and the 8 x 15 sums live in registers
// (the entries a direction never touches fold away).  Replaces cdef_find_dir_rust, src/cdef.rs:921.
__device__ __forceinline__ constexpr int cdef_line(int d, int i, int j) {
    return d == 0 ? i + j : d == 1 ? i + j / 2 : d == 2 ? i : d == 3 ? 3 + i - j / 2 : d == 4 ? 7 + i - j : d == 5 ? 3 - i / 2 + j
         : d == 6 ? j : i / 2 + j;
}
__device__ __forceinline__ int cdef_find_dir(const int16_t *t, int pitch, int bdmin8, unsigned *var) {
    int partial[8][15] = {};
#pragma unroll
    for (int i = 0; i < 8; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int x = ((int)t[i * pitch + j] >> bdmin8) - 128;
#pragma unroll
            for (int d = 0; d < 8; d++) partial[d][cdef_line(d, i, j)] += x;
        }
    }
    // pixels on line k of a direction: diagonals (0, 4) 1..8..1 over 15 lines; the half-step directions (odd) 2, 4, 6, then
    // 8 on the 5 middle lines of 11; rows and columns (2, 6) 8 on each of 8 lines.  weight = 840 / pixels.
    unsigned cost[8];
#pragma unroll
    for (int d = 0; d < 8; d++) {
        unsigned c = 0;
#pragma unroll
        for (int k = 0; k < 15; k++) {
            int n;   // pixels on the line
            if (d == 2 || d == 6) n = k < 8 ? 8 : 0;
            else if (d & 1) n = k < 3 ? 2 * (k + 1) : (k < 8 ? 8 : (k < 11 ? 2 * (11 - k) : 0));
            else n = k < 8 ? k + 1 : 15 - k;
            if (n) c += (unsigned)(partial[d][k] * partial[d][k]) * (unsigned)(840 / n);
        }
        cost[d] = c;
    }
    int best = 0;
    unsigned best_cost = cost[0];
#pragma unroll
    for (int d = 1; d < 8; d++)
        if (cost[d] > best_cost) { best_cost = cost[d]; best = d; }
    unsigned across = 0;      // the cost of the direction at right angles to the best one
#pragma unroll
    for (int d = 0; d < 8; d++) if (d == (best ^ 4)) across = cost[d];
    *var = (best_cost - across) >> 10;
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

// ---- the filter proper: pixel PAIRS in 32-bit registers, one lane per pair of a tile row
//
// A warp covers one tile row (luma: 32 pairs) or two (4:2:0 chroma: 16 pairs each) and walks down the rows of ONE
// row of 8x8 blocks, so the per-block parameters are set up once per thread; consecutive lanes read consecutive
// words of shared memory (no bank conflicts whatever the row pitch) and write consecutive words of the picture.
//
// constrain(d) = sign(d) * min(|d|, max(0, thr - (|d| >> shift)))  (src/cdef.rs:57-66) on two unsigned halves:
//   dpos = max(p, px) - px,  dneg = px - min(p, px)     one of them is 0; plain 32-bit subtractions cannot borrow
//   a    = dpos + dneg                                   = |d|
//   mag  = relu(min(thr - (a >> shift), a))              one VIADDMNMX.S16x2.RELU
//   the tap adds w * min(dpos, mag) to one sum and w * min(dneg, mag) to another; the sums (< 2^12) meet in the
//   epilogue, which rounds, adds the centre pixel and clips to the tracked min / max, all still packed.
// A tap that is "unavailable" (the reference's INT16_MIN sentinel, 0x8000 as an unsigned half) gives a >= 0x7001, so
// mag = 0 and the tap drops out exactly as in the reference; it never wins the unsigned min or the signed max either.
// Lanes whose block has no primary (secondary) strength run those taps with thr = 0, i.e. mag = 0.
template <typename BD, bool PRI, bool SEC>
__device__ __forceinline__ void cdef_rows(const char *cb, int row_step, int n_rows, const int (&aoff)[12], unsigned wp0,
                                          unsigned wp1, unsigned pthr1, unsigned pmask, int pshift, unsigned sthr1,
                                          unsigned smask, int sshift, bool clip, uint8_t *d, int64_t dstep, bool store) {
    constexpr bool TRACK = PRI && SEC;
    const unsigned clip_lo = clip ? 0u : 0xffffffffu, clip_hi = clip ? 0u : 0x7fff7fffu;
#pragma unroll 1
    for (int i = 0; i < n_rows; i++, cb += row_step, d += dstep) {
        const unsigned px2 = *(const unsigned *)cb;
        unsigned sum_p = 0, sum_n = 0, mn = px2, mx = px2;
        auto tap = [&](int off, unsigned thr1, unsigned m, int shift, unsigned w) {
            const unsigned p2 = *(const unsigned *)(cb + off);
            const unsigned dpos = __vmaxu2(p2, px2) - px2, dneg = px2 - __vminu2(p2, px2);
            const unsigned a = dpos + dneg;
            const unsigned mag = __viaddmin_s16x2_relu(~((a >> shift) & m), thr1, a);
            sum_p += w * __vminu2(dpos, mag);
            sum_n += w * __vminu2(dneg, mag);
            if (TRACK) { mn = __vminu2(mn, p2); mx = __vmaxs2(mx, p2); }
        };
        if (PRI) {
            tap(aoff[0], pthr1, pmask, pshift, wp0); tap(aoff[1], pthr1, pmask, pshift, wp0);
            tap(aoff[2], pthr1, pmask, pshift, wp1); tap(aoff[3], pthr1, pmask, pshift, wp1);
        }
        if (SEC) {
#pragma unroll
            for (int k = 4; k < 8; k++) tap(aoff[k], sthr1, smask, sshift, 2u);
#pragma unroll
            for (int k = 8; k < 12; k++) tap(aoff[k], sthr1, smask, sshift, 1u);
        }
        // (sum - (sum < 0) + 8) >> 4 on both halves: sums biased by 4096 (a multiple of 16, > any |sum|)
        const unsigned tb = sum_p + (0x10001000u - sum_n);
        const unsigned neg = (~tb >> 12) & 0x00010001u;
        const unsigned r = ((tb + 0x00080008u - neg) >> 4) & 0x0fff0fffu;   // 256 + rounded sum
        unsigned vb = px2 + r;                                              // 256 + result, never negative
        if (TRACK) vb = __vminu2(__vmaxu2(vb, (mn & ~clip_lo) + 0x01000100u), (mx | clip_hi) + 0x01000100u);
        const unsigned v = vb - 0x01000100u;
        if (store) {
            if (BD::hbd) *(unsigned *)d = v;
            else *(uint16_t *)d = (uint16_t)((v & 0xff) | ((v >> 8) & 0xff00));
        }
    }
}

// One plane (luma) or two planes that share the block parameters (U and V) of the 64x64 area: `n_sub` staged tiles of
// the same geometry.  tile pitch (pixels, even), copy_el: elements from copy A to the one-pixel-shifted copy B
// (B[i] = A[i + 1]); the area's first pixel sits at row 2, column col0 (even) of a tile.
template <typename BD>
__device__ __forceinline__ void cdef_filter_tiles(const uint16_t *tile0, const uint16_t *tile1, uint8_t *dbase0, uint8_t *dbase1,
                                                  int64_t dstride0, int64_t dstride1, int n_sub, int pitch, int copy_el, int col0,
                                                  const CdefBlk *blk, int px0, int py0, int tw, int th, int ssh, int ssv, bool chroma,
                                                  int damping, int bdmin8) {
    using pixel = typename BD::pixel;
    const int lane = threadIdx.x & 31, byl = threadIdx.x >> 5;
    const int bh = 8 >> ssv;
    if (byl * bh >= th) return;                          // warp-uniform: block rows below the picture
    const int ppr_log2 = 5 - ssh;                         // pairs per tile row: 32 or 16
    const int pair = lane & ((1 << ppr_log2) - 1), rsub = lane >> ppr_log2, rows_per_it = 1 << ssh;
    const int x = 2 * pair;
    const CdefBlk b = blk[byl * 8 + (x >> (3 - ssh))];
    const bool on = (chroma ? b.do_uv : b.do_y) && x < tw;
    const int pri = on ? (chroma ? b.uv_pri : b.y_pri) : 0, sec = on ? (chroma ? b.uv_sec : b.y_sec) : 0;
    const int dir = chroma ? b.uvdir : b.dir;
    // byte offsets of the 12 taps from the centre pair: an odd column offset reads copy B
    int aoff[12];
    {
        const int odd_adj = 2 * (copy_el - 1);
        auto set = [&](int slot, int d, int k) {
            int dy, dx;
            tab::cdef_dir_off(d, k, dy, dx);
            const int o = 2 * (dy * pitch + dx), adj = (dx & 1) ? odd_adj : 0;
            aoff[slot] = o + adj; aoff[slot + 1] = adj - o;
        };
        set(0, dir, 0); set(2, dir, 1);
        set(4, (dir + 2) & 7, 0); set(6, (dir + 6) & 7, 0);
        set(8, (dir + 2) & 7, 1); set(10, (dir + 6) & 7, 1);
    }
    const unsigned pri_tap = 4 - ((pri >> bdmin8) & 1);
    const unsigned wp0 = pri_tap, wp1 = (pri_tap & 3) | 2;
    const int pshift = pri ? imax(0, damping - ulog2(pri)) : 0, sshift = sec ? damping - ulog2(sec) : 0;
    const unsigned pthr1 = (unsigned)(pri + 1) * 0x10001u, sthr1 = (unsigned)(sec + 1) * 0x10001u;
    const unsigned pmask = (0xffffu >> pshift) * 0x10001u, smask = (0xffffu >> sshift) * 0x10001u;
    const bool clip = pri && sec;
    const bool any_pri = __any_sync(0xffffffffu, pri != 0), any_sec = __any_sync(0xffffffffu, sec != 0);
    const int y0r = byl * bh + rsub;                     // first row of this lane, then every rows_per_it-th
    const int n_rows = bh >> ssh;
    const bool store = x < tw;
#pragma unroll
    for (int sub = 0; sub < 2; sub++) {
        if (sub >= n_sub) break;
        const char *cb = (const char *)((sub ? tile1 : tile0) + (2 + y0r) * pitch + col0 + x);
        const int row_step = 2 * pitch * rows_per_it;
        const int64_t ds = sub ? dstride1 : dstride0;
        uint8_t *d = (sub ? dbase1 : dbase0) + (int64_t)(py0 + y0r) * ds + (int64_t)(px0 + x) * sizeof(pixel);
        const int64_t dstep = ds * rows_per_it;
        if (any_pri && any_sec)
            cdef_rows<BD, true, true>(cb, row_step, n_rows, aoff, wp0, wp1, pthr1, pmask, pshift, sthr1, smask, sshift, clip, d, dstep, store);
        else if (any_pri)
            cdef_rows<BD, true, false>(cb, row_step, n_rows, aoff, wp0, wp1, pthr1, pmask, pshift, sthr1, smask, sshift, false, d, dstep, store);
        else if (any_sec)
            cdef_rows<BD, false, true>(cb, row_step, n_rows, aoff, wp0, wp1, pthr1, pmask, pshift, sthr1, smask, sshift, false, d, dstep, store);
        else
            cdef_rows<BD, false, false>(cb, row_step, n_rows, aoff, wp0, wp1, pthr1, pmask, pshift, sthr1, smask, sshift, false, d, dstep, store);
    }
}

// 8-bit pictures: tiles widened to 16 bits by cdef_stage2, one plane at a time through one buffer.
template <typename BD>
__global__ void __launch_bounds__(256)
cdef_filter_frame_kernel(Rb200Planes src, Rb200Planes dst, CdefFrameParams P, const CdefBlk *__restrict__ blocks, int nbx,
                         int nby, int tile_row_first, int plane_mask) {
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
    bool first = true;
    for (int p = 0; p < P.n_planes; p++) {
        if (!((plane_mask >> p) & 1)) continue;
        if (!first) __syncthreads();
        first = false;
        const int ssh = p ? P.ss_hor : 0, ssv = p ? P.ss_ver : 0;
        cdef_stage2<BD>(tile, (const uint8_t *)plane_ptr(src, p), plane_stride(src, p), x0 >> ssh, y0 >> ssv, tw >> ssh, th >> ssv, fw >> ssh, fh >> ssv);
        __syncthreads();
        cdef_filter_tiles<BD>((const uint16_t *)tile, nullptr, (uint8_t *)plane_ptr(dst, p), nullptr, plane_stride(dst, p), 0, 1, CDEF_TP, CDEF_COPY, CDEF_X0, blk, x0 >> ssh, y0 >> ssv, tw >> ssh, th >> ssv, ssh, ssv,
                              p != 0, P.damping - (p ? 1 : 0), P.bdmin8);
    }
}

// 16-bit pictures: the tiles are fetched by the copy engine.  Per plane one box of (tile + 16) x (tile + 4) pixels at
// (x0 - 8, y0 - 2) -- the copy engine wants the first column of a box 16-byte aligned, hence 8 columns of halo where
// the taps reach 2 -- lands in shared memory as dense rows (copy A); all three are requested by one thread at kernel
// start.  What lies outside the picture arrives as zeros and is overwritten with the reference's "unavailable" sentinel
// by the (few) CTAs on the picture border.  Copy B (one pixel to the left, so that odd column offsets are aligned words
// too) is made from copy A in shared memory: one funnel shift per word.
struct CdefTmaGeom {                     // (scalars, not arrays: a run-time index into a kernel parameter costs a local copy)
    int pitch_y, rows_y, copy_y;         // luma tile geometry (pixels / elements)
    int pitch_c, rows_c, copy_c;         // chroma
    int off_y, off_u, off_v;             // byte offset of a plane's copy A in dynamic shared memory
};
constexpr int CDEF_TMA_X0 = 8;
__device__ __forceinline__ void cdef_patch_outside(uint16_t *tile, int pitch, int rows, int x0, int y0, int fw, int fh) {
    for (int i = threadIdx.x; i < rows * pitch; i += blockDim.x) {
        const int r = i / pitch, c = i - r * pitch;
        const int y = y0 - 2 + r, x = x0 - CDEF_TMA_X0 + c;
        if (y < 0 || y >= fh || x < 0 || x >= fw) tile[i] = (uint16_t)CDEF_SENTINEL;
    }
}
__device__ __forceinline__ void cdef_shifted_copy(uint16_t *tile, int n_el, int copy_el) {
    const unsigned *a = (const unsigned *)tile;
    unsigned *b = (unsigned *)(tile + copy_el);
    const int nw = n_el >> 1;
    for (int i = threadIdx.x; i < nw; i += blockDim.x) b[i] = __funnelshift_r(a[i], i + 1 < nw ? a[i + 1] : 0u, 16);
}
template <typename BD>
__global__ void __launch_bounds__(256, 3)
cdef_filter_tma_kernel(const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_u,
                       const __grid_constant__ CUtensorMap map_v, Rb200Planes dst, CdefFrameParams P, CdefTmaGeom G,
                       const CdefBlk *__restrict__ blocks, int nbx, int nby, int tile_row_first, int plane_mask) {
    extern __shared__ uint8_t cdef_dyn_smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ CdefBlk blk[64];
    uint8_t *sm = cdef_dyn_smem + ((128u - (smem_addr(cdef_dyn_smem) & 127u)) & 127u);   // TMA destinations: 128-byte aligned
    const int sbx = blockIdx.x, sby = tile_row_first + blockIdx.y;
    const int x0 = sbx * 64, y0 = sby * 64;
    const int fw = P.bw * 4, fh = P.bh * 4;
    const int tw = imin(64, fw - x0), th = imin(64, fh - y0);
    const bool luma = plane_mask & 1, chroma = P.n_planes > 1 && (plane_mask & 6);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        const unsigned bytes = 2u * (unsigned)((luma ? G.pitch_y * G.rows_y : 0) + (chroma ? 2 * G.pitch_c * G.rows_c : 0));
        mbar_arrive_expect_tx(&bar, bytes);
        if (luma) tma_load_2d(sm + G.off_y, &map_y, x0 - CDEF_TMA_X0, y0 - 2, &bar);
        if (chroma) {
            const int cx = (x0 >> P.ss_hor) - CDEF_TMA_X0, cy = (y0 >> P.ss_ver) - 2;
            tma_load_2d(sm + G.off_u, &map_u, cx, cy, &bar);
            tma_load_2d(sm + G.off_v, &map_v, cx, cy, &bar);
        }
    }
    if (threadIdx.x < 64) {
        const int bx = sbx * 8 + (threadIdx.x & 7), by = sby * 8 + (threadIdx.x >> 3);
        CdefBlk b = {};
        if (bx < nbx && by < nby) b = blocks[by * nbx + bx];
        blk[threadIdx.x] = b;
    }
    __syncthreads();     // barrier initialised, block records visible
    mbar_wait(&bar, 0);
    if (x0 == 0 || y0 == 0 || x0 + 64 >= fw || y0 + 64 >= fh) {      // the tile's halo leaves the picture
        if (luma) cdef_patch_outside((uint16_t *)(sm + G.off_y), G.pitch_y, G.rows_y, x0, y0, fw, fh);
        if (chroma) {
            cdef_patch_outside((uint16_t *)(sm + G.off_u), G.pitch_c, G.rows_c, x0 >> P.ss_hor, y0 >> P.ss_ver, fw >> P.ss_hor, fh >> P.ss_ver);
            cdef_patch_outside((uint16_t *)(sm + G.off_v), G.pitch_c, G.rows_c, x0 >> P.ss_hor, y0 >> P.ss_ver, fw >> P.ss_hor, fh >> P.ss_ver);
        }
        __syncthreads();
    }
    if (luma) cdef_shifted_copy((uint16_t *)(sm + G.off_y), G.pitch_y * G.rows_y, G.copy_y);
    if (chroma) {
        cdef_shifted_copy((uint16_t *)(sm + G.off_u), G.pitch_c * G.rows_c, G.copy_c);
        cdef_shifted_copy((uint16_t *)(sm + G.off_v), G.pitch_c * G.rows_c, G.copy_c);
    }
    __syncthreads();
    if (luma)
        cdef_filter_tiles<BD>((const uint16_t *)(sm + G.off_y), nullptr, (uint8_t *)dst.data[0], nullptr, dst.stride[0], 0, 1, G.pitch_y, G.copy_y,
                              CDEF_TMA_X0, blk, x0, y0, tw, th, 0, 0, false, P.damping, P.bdmin8);
    if (chroma)
        cdef_filter_tiles<BD>((const uint16_t *)(sm + G.off_u), (const uint16_t *)(sm + G.off_v), (uint8_t *)dst.data[1], (uint8_t *)dst.data[2],
                              dst.stride[1], dst.stride[2], 2, G.pitch_c, G.copy_c, CDEF_TMA_X0, blk, x0 >> P.ss_hor, y0 >> P.ss_ver,
                              tw >> P.ss_hor, th >> P.ss_ver, P.ss_hor, P.ss_ver, true, P.damping - 1, P.bdmin8);
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

// Tensor maps of the three source planes for cdef_filter_tma_kernel (16-bit pictures); the frame context encodes them
// once, its planes never move.
int cdef_encode_maps(CUtensorMap maps[3], const Rb200Planes &src, const CdefFrameParams &P) {
    for (int p = 0; p < P.n_planes; p++) {
        const int ssh = p ? P.ss_hor : 0, ssv = p ? P.ss_ver : 0;
        const int r = tma_encode_plane(&maps[p], src.data[p], 2, (P.bw * 4) >> ssh, (P.bh * 4) >> ssv, src.stride[p], (64 >> ssh) + 16,
                                       (64 >> ssv) + 4);
        if (r) return r;
    }
    for (int p = P.n_planes; p < 3; p++) maps[p] = maps[0];
    return 0;
}

// t0 / t1: 64-row tile rows to produce (whole picture: 0, ceil(height / 64)).  maps: cdef_encode_maps() of `src`
// (16-bit pictures; null = stage with loads).
int cdef_frame_launch(const Rb200Planes &src, const Rb200Planes &dst, const CdefFrameParams &P,
                      const Rb200Av1Filter *masks, void *blk_scratch, int bdmax, cudaStream_t st, int t0, int t1,
                      const CUtensorMap *maps) {
    return cdef_planes_launch(src, dst, P, masks, blk_scratch, bdmax, st, t0, t1, maps, 7, 3, nullptr);
}

int cdef_planes_launch(const Rb200Planes &src, const Rb200Planes &dst, const CdefFrameParams &P, const Rb200Av1Filter *masks,
                       void *blk_scratch, int bdmax, cudaStream_t st, int t0, int t1, const CUtensorMap *maps, int plane_mask,
                       int what, int *launches) {
    const int nbx = P.bw >> 1, nby = P.bh >> 1;
    CdefBlk *blocks = (CdefBlk *)blk_scratch;
    const int by0 = t0 * 8, by1 = imin(t1 * 8, nby);
    if (by1 <= by0) return 0;
    dim3 g1((nbx + 31) / 32, (by1 - by0 + 3) / 4), b1(32, 4);
    dim3 grid((P.bw * 4 + 63) / 64, t1 - t0);
    if (what & 1) {
        if (bdmax > 255) cdef_dir_frame_kernel<BD16><<<g1, b1, 0, st>>>(src, P, masks, blocks, nbx, nby, by0, by1);
        else cdef_dir_frame_kernel<BD8><<<g1, b1, 0, st>>>(src, P, masks, blocks, nbx, nby, by0, by1);
        if (launches) ++*launches;
    }
    if ((what & 2) && plane_mask) {
        if (bdmax > 255 && maps) {
            const bool luma = plane_mask & 1, chroma = P.n_planes > 1 && (plane_mask & 6);
            CdefTmaGeom G;
            G.pitch_y = 64 + 16; G.rows_y = 64 + 4;
            G.pitch_c = (64 >> P.ss_hor) + 16; G.rows_c = (64 >> P.ss_ver) + 4;
            G.copy_y = (G.pitch_y * G.rows_y + 63) & ~63;      // copy B starts 128-byte aligned as well
            G.copy_c = (G.pitch_c * G.rows_c + 63) & ~63;
            G.off_y = 0; G.off_u = luma ? 4 * G.copy_y : 0; G.off_v = G.off_u + 4 * G.copy_c;
            const int smem = (chroma ? G.off_v + 4 * G.copy_c : G.off_u) + 128;
            static int smem_set[64] = {};     // per device: the opt-in above 48 KB (4:4:4) is a per-context function attribute
            int dev = 0;
            RB_CUDA(cudaGetDevice(&dev));
            if (smem > 48 * 1024 && smem > smem_set[dev & 63]) {
                RB_CUDA(cudaFuncSetAttribute(cdef_filter_tma_kernel<BD16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                smem_set[dev & 63] = smem;
            }
            cdef_filter_tma_kernel<BD16><<<grid, 256, smem, st>>>(maps[0], maps[1], maps[2], dst, P, G, blocks, nbx, nby, t0, plane_mask);
        } else if (bdmax > 255) {
            cdef_filter_frame_kernel<BD16><<<grid, 256, 0, st>>>(src, dst, P, blocks, nbx, nby, t0, plane_mask);
        } else {
            cdef_filter_frame_kernel<BD8><<<grid, 256, 0, st>>>(src, dst, P, blocks, nbx, nby, t0, plane_mask);
        }
        if (launches) ++*launches;
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
