// Loop-filter masks and levels built on the device.
//
// Reference: rav1d_create_lf_mask_intra / rav1d_create_lf_mask_inter, src/lf_mask.rs:380-606 (C twin
// src/lf_mask.c:39-406), called per block from decode_b (src/decode.c:1260-1271,1926-1947), the noskip mask of
// src/decode.c:1996-2005, and the tile-edge fix-ups of src/lf_apply.rs (C: src/lf_apply_tmpl.c:331-400).
//
// The reference is serial: every call ORs bits into Av1Filter words and updates the above / left transform-size
// contexts (`a`, `l`) the next block reads.  Those contexts only ever hold "the clamped transform size of the 4x4
// cell just above / left of this one", so the whole computation is a pure function of per-cell facts:
//
//   pass 1 (scatter, lf_cells_kernel): an 8-lane group per block record (the whole warp for big blocks) walks its 4x4 cells, finds the
//     transform leaf that covers each cell (the tx_split tree of decomp_tx evaluated in closed form: at most two
//     splits, each halving the longer side(s)), and writes one byte per luma cell and one per chroma cell:
//       bits 0-1 min(2, lw) (chroma: min(1, lw))   bits 2-3 the same for lh
//       bit 4    the cell's left boundary is a filtered edge (block edge, or leaf edge of a non-skipped block)
//       bit 5    the same for its top boundary      bit 6  cell belongs to a non-skipped block (luma only)
//       bit 7    cell written (inside the picture)
//     plus the block's levels into lf.level[cell][0..1] (luma cells) and [2..3] (chroma-coordinate cells).
//   pass 2 (gather, lf_words_kernel): a CTA per 128x128 area stages its 32x32 cell bytes (and the column left /
//     row above) in shared memory; one thread per (direction, position, half) assembles the 2 or 3 mask words of
//     that position bit by bit: bit set in word min(own size, neighbour's size) -- the neighbour's size is 2 (luma) /
//     1 (chroma) at the picture's edge, the reset value of the contexts (src/decode.c:2451-2452).  No atomics, every
//     word written exactly once, so the arrays need no clearing and the result does not depend on record order.
//
// Across tile boundaries the reference first uses the reset value and later corrects the word with the real
// neighbour (the fix-ups); the gather produces the corrected value directly.
#include "common.cuh"

namespace rb200 {
namespace {

// log2 of the transform width / height in 4-pixel units, indexed by RectTxfmSize (src/levels.rs:31-59)
__constant__ uint8_t c_tx_lw[RB200_N_RECT_TX_SIZES] = { 0, 1, 2, 3, 4, 0, 1, 1, 2, 2, 3, 3, 4, 0, 2, 1, 3, 2, 4 };
__constant__ uint8_t c_tx_lh[RB200_N_RECT_TX_SIZES] = { 0, 1, 2, 3, 4, 1, 0, 2, 1, 3, 2, 4, 3, 2, 0, 3, 1, 4, 2 };
// block width / height in 4-pixel units, indexed by BlockSize (BS_128x128 .. BS_4x4)
__constant__ uint8_t c_bs_w4[22] = { 32, 32, 16, 16, 16, 16, 8, 8, 8, 8, 4, 4, 4, 4, 4, 2, 2, 2, 2, 1, 1, 1 };
__constant__ uint8_t c_bs_h4[22] = { 32, 16, 32, 16, 8, 4, 16, 8, 4, 2, 16, 8, 4, 2, 1, 8, 4, 2, 1, 4, 2, 1 };

constexpr int CELL_V = 16, CELL_H = 32, CELL_NOSKIP = 64, CELL_VALID = 128;
constexpr int GROUP = 8;    // lanes per block record

struct LfGeom {
    int w4, h4;             // picture size in 4-px luma units (f.w4, f.h4)
    int cw4, ch4;           // the same for chroma
    int ss_hor, ss_ver, chroma;
    int cell_stride;        // 32 * sb128w
    int b4_stride;
    int sb128w, sb128h;
};

// One split step of decomp_tx (src/lf_mask.c:45-63): the longer side(s) are halved; returns the leaf after at most
// two steps for the cell (y, x) of the block.  ox / oy: origin of the leaf inside the block.
__device__ __forceinline__ void tx_leaf(int lw, int lh, unsigned split0, unsigned split1, int x, int y, int &olw, int &olh,
                                        int &ox, int &oy) {
    int x_off = x >> lw, y_off = y >> lh;
    ox = x_off << lw; oy = y_off << lh;
#pragma unroll
    for (int depth = 0; depth < 2; depth++) {
        if (!(lw | lh)) break;
        const unsigned m = depth ? split1 : split0;
        if (!((m >> (y_off * 4 + x_off)) & 1)) break;
        const int sw = lw >= lh, sh = lh >= lw;
        const int nlw = lw - sw, nlh = lh - sh;
        const int sx = sw ? ((x - ox) >> nlw) & 1 : 0, sy = sh ? ((y - oy) >> nlh) & 1 : 0;
        ox += sx << nlw; oy += sy << nlh;
        x_off = x_off * 2 + sx; y_off = y_off * 2 + sy;
        lw = nlw; lh = nlh;
    }
    olw = lw; olh = lh;
}

// All cells of one block record, strided over `nlanes` cooperating lanes.
__device__ __forceinline__ void lf_block_cells(const uint4 raw, int lane, int nlanes, const LfGeom &G, uint8_t *__restrict__ cell_y,
                                               uint8_t *__restrict__ cell_uv, uint8_t (*__restrict__ lvl)[4]) {
    const int bx = raw.x & 0xffff, by = raw.x >> 16;
    const int bs = raw.y & 0xff, flags = (raw.y >> 8) & 0xff, ytx = (raw.y >> 16) & 0xff, uvtx = raw.y >> 24;
    unsigned split0 = raw.z & 0xffff, split1 = raw.z >> 16;
    const int intra = flags & RB200_LFB_INTRA, skip = (flags & RB200_LFB_SKIP) != 0;
    if (intra) split0 = split1 = 0;
    const int inner = intra || !skip;                 // mask_edges_intra has no skip test (src/lf_mask.c:176-197)
    const int w4 = c_bs_w4[bs], h4 = c_bs_h4[bs];
    const int lw = c_tx_lw[ytx], lh = c_tx_lh[ytx];
    const uint16_t lv01 = raw.w & 0xffff, lv23 = raw.w >> 16;
    // ---- luma cells (noskip covers the whole block, everything else only the part inside the picture)
    const int lw4 = 31 - __clz(w4);
    for (int c = lane; c < w4 * h4; c += nlanes) {
        const int x = c & (w4 - 1), y = c >> lw4;
        const int fx = bx + x, fy = by + y;
        if (fx >= G.cell_stride || fy >= G.sb128h * 32) continue;
        uint8_t v = skip ? 0 : CELL_NOSKIP;
        if (fx < G.w4 && fy < G.h4) {
            int llw, llh, ox, oy;
            tx_leaf(lw, lh, split0, split1, x, y, llw, llh, ox, oy);
            v |= CELL_VALID | min(2, llw) | (min(2, llh) << 2);
            if (x == 0 || (inner && x == ox)) v |= CELL_V;
            if (y == 0 || (inner && y == oy)) v |= CELL_H;
            *reinterpret_cast<uint16_t *>(&lvl[(size_t)fy * G.b4_stride + fx][0]) = lv01;
        }
        cell_y[(size_t)fy * G.cell_stride + fx] = v;
    }
    // ---- chroma cells (src/lf_mask.c:318-343,380-405; mask_edges_chroma :215-284); skip_inter = 0 for intra blocks
    if (!G.chroma || !(flags & RB200_LFB_HAS_CHROMA)) return;
    const int cbx = bx >> G.ss_hor, cby = by >> G.ss_ver;
    const int cbw4 = min(G.cw4 - cbx, (w4 + G.ss_hor) >> G.ss_hor), cbh4 = min(G.ch4 - cby, (h4 + G.ss_ver) >> G.ss_ver);
    if (cbw4 <= 0 || cbh4 <= 0) return;
    const int ulw = c_tx_lw[uvtx], ulh = c_tx_lh[uvtx];
    const uint8_t base = CELL_VALID | min(1, ulw) | (min(1, ulh) << 2);
    for (int c = lane; c < cbw4 * cbh4; c += nlanes) {
        const int x = c % cbw4, y = c / cbw4;
        uint8_t v = base;
        if (x == 0 || (inner && !(x & ((1 << ulw) - 1)))) v |= CELL_V;
        if (y == 0 || (inner && !(y & ((1 << ulh) - 1)))) v |= CELL_H;
        cell_uv[(size_t)(cby + y) * G.cell_stride + cbx + x] = v;
        *reinterpret_cast<uint16_t *>(&lvl[(size_t)(cby + y) * G.b4_stride + cbx + x][2]) = lv23;
    }
}

// An 8-lane group per record; blocks of more than 32 cells are then walked by the whole warp, one after the other, so a
// 128x128 block costs 32 iterations instead of 128 and does not hold three idle groups hostage.
__global__ void __launch_bounds__(256) lf_cells_kernel(const Rb200LfBlock *__restrict__ blocks, int n, LfGeom G,
                                                       uint8_t *__restrict__ cell_y, uint8_t *__restrict__ cell_uv,
                                                       uint8_t (*__restrict__ lvl)[4]) {
    const int gid = (blockIdx.x * blockDim.x + threadIdx.x) / GROUP, lane = threadIdx.x % GROUP;
    uint4 raw = make_uint4(0, 0, 0, 0);
    bool valid = gid < n;
    if (valid) {
        raw = __ldg(reinterpret_cast<const uint4 *>(blocks) + gid);
        valid = (raw.y & 0xff) < 22 && ((raw.y >> 16) & 0xff) < RB200_N_RECT_TX_SIZES && (raw.y >> 24) < RB200_N_RECT_TX_SIZES;
    }
    const bool big = valid && c_bs_w4[raw.y & 0xff] * c_bs_h4[raw.y & 0xff] > 32;
    if (valid && !big) lf_block_cells(raw, lane, GROUP, G, cell_y, cell_uv, lvl);
    unsigned m = __ballot_sync(0xffffffffu, big) & 0x01010101u;
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        uint4 r;
        r.x = __shfl_sync(0xffffffffu, raw.x, src); r.y = __shfl_sync(0xffffffffu, raw.y, src);
        r.z = __shfl_sync(0xffffffffu, raw.z, src); r.w = __shfl_sync(0xffffffffu, raw.w, src);
        lf_block_cells(r, threadIdx.x & 31, 32, G, cell_y, cell_uv, lvl);
    }
}

// tile[1 + y][1 + x]: cell bytes of one 128x128 area, row / column 0 = the neighbours above / left (0 outside the map:
// not valid, so the callers read the reset value).
__device__ __forceinline__ void stage_cells(uint8_t (*tile)[36], const uint8_t *__restrict__ cells, int stride, int rows, int x0,
                                            int y0, int nx, int ny) {
    for (int i = threadIdx.x; i < (ny + 1) * (nx + 1); i += blockDim.x) {
        const int ty = i / (nx + 1), tx = i % (nx + 1);
        const int fx = x0 + tx - 1, fy = y0 + ty - 1;
        tile[ty][tx] = fx >= 0 && fy >= 0 && fx < stride && fy < rows ? cells[(size_t)fy * stride + fx] : 0;
    }
}

__global__ void __launch_bounds__(288) lf_words_kernel(LfGeom G, const uint8_t *__restrict__ cell_y, const uint8_t *__restrict__ cell_uv,
                                                       const int8_t *__restrict__ cdef_idx, Rb200Av1Filter *__restrict__ masks) {
    __shared__ uint8_t ty[33][36], tuv[33][36];
    const int sbx = blockIdx.x, sby = blockIdx.y;
    Rb200Av1Filter &M = masks[sby * G.sb128w + sbx];
    stage_cells(ty, cell_y, G.cell_stride, G.sb128h * 32, sbx * 32, sby * 32, 32, 32);
    const int cnx = 32 >> G.ss_hor, cny = 32 >> G.ss_ver;
    if (G.chroma) stage_cells(tuv, cell_uv, G.cell_stride, G.sb128h * 32, sbx * cnx, sby * cny, cnx, cny);
    __syncthreads();
    const int t = threadIdx.x;
    if (t < 128) {
        // filter_y[dir][pos][0..2][half]: dir 0 = column edges (bit = row), dir 1 = row edges (bit = column)
        const int dir = t >> 6, pos = (t >> 1) & 31, half = t & 1;
        unsigned w[3] = { 0, 0, 0 };
#pragma unroll 4
        for (int b = 0; b < 16; b++) {
            const int q = half * 16 + b;
            const int y = dir ? pos : q, x = dir ? q : pos;
            const unsigned c = ty[1 + y][1 + x];
            if (!(c & CELL_VALID) || !(c & (dir ? CELL_H : CELL_V))) continue;
            const unsigned nb = dir ? ty[y][1 + x] : ty[1 + y][x];
            const int own = dir ? (c >> 2) & 3 : c & 3;
            const int other = (nb & CELL_VALID) ? (dir ? (nb >> 2) & 3 : nb & 3) : 2;
            w[min(own, other)] |= 1u << b;
        }
        M.filter_y[dir][pos][0][half] = (uint16_t)w[0];
        M.filter_y[dir][pos][1][half] = (uint16_t)w[1];
        M.filter_y[dir][pos][2][half] = (uint16_t)w[2];
    } else if (t < 256) {
        // filter_uv[dir][pos][0..1][half]; halves are 16 >> ss bits wide (src/lf_mask.c:227-244)
        const int u = t - 128, dir = u >> 6, pos = (u >> 1) & 31, half = u & 1;
        unsigned w[2] = { 0, 0 };
        const int npos = dir ? cny : cnx, hbits = dir ? 16 >> G.ss_hor : 16 >> G.ss_ver;
        if (G.chroma && pos < npos) {
            for (int b = 0; b < hbits; b++) {
                const int q = half * hbits + b;
                const int y = dir ? pos : q, x = dir ? q : pos;
                const unsigned c = tuv[1 + y][1 + x];
                if (!(c & CELL_VALID) || !(c & (dir ? CELL_H : CELL_V))) continue;
                const unsigned nb = dir ? tuv[y][1 + x] : tuv[1 + y][x];
                const int own = dir ? (c >> 2) & 1 : c & 1;
                const int other = (nb & CELL_VALID) ? (dir ? (nb >> 2) & 1 : nb & 1) : 1;
                w[min(own, other)] |= 1u << b;
            }
        }
        M.filter_uv[dir][pos][0][half] = (uint16_t)w[0];
        M.filter_uv[dir][pos][1][half] = (uint16_t)w[1];
    } else {
        // noskip_mask[row of 8 pixels][half]: a bit per 4-pixel column, set if either 4-pixel row is not skipped
        const int u = t - 256, r = u >> 1, half = u & 1;
        unsigned w = 0;
#pragma unroll 4
        for (int b = 0; b < 16; b++) {
            const int x = half * 16 + b;
            if ((ty[1 + 2 * r][1 + x] | ty[2 + 2 * r][1 + x]) & CELL_NOSKIP) w |= 1u << b;
        }
        M.noskip_mask[r][half] = (uint16_t)w;
        if (u < 4) M.cdef_idx[u] = cdef_idx[(sby * G.sb128w + sbx) * 4 + u];
    }
}

}  // namespace

// cells: device scratch of 2 * (32 sb128w) * (32 sb128h) bytes; cdef_idx: device, [sb128h * sb128w][4]
int lf_build_launch(const Rb200LfBlock *d_blocks, int n, int w4, int h4, int sb128w, int sb128h, int b4_stride, int ss_hor,
                    int ss_ver, int n_planes, uint8_t *cells, const int8_t *cdef_idx, Rb200Av1Filter *masks, uint8_t (*lvl)[4],
                    cudaStream_t st) {
    LfGeom G;
    G.w4 = w4; G.h4 = h4;
    G.ss_hor = ss_hor; G.ss_ver = ss_ver; G.chroma = n_planes > 1;
    G.cw4 = (w4 + ss_hor) >> ss_hor; G.ch4 = (h4 + ss_ver) >> ss_ver;
    G.cell_stride = 32 * sb128w; G.b4_stride = b4_stride; G.sb128w = sb128w; G.sb128h = sb128h;
    const size_t plane = (size_t)G.cell_stride * 32 * sb128h;
    RB_CUDA(cudaMemsetAsync(cells, 0, 2 * plane, st));
    if (n > 0) {
        const int per_cta = 256 / GROUP;
        lf_cells_kernel<<<(n + per_cta - 1) / per_cta, 256, 0, st>>>(d_blocks, n, G, cells, cells + plane, lvl);
        RB_LAUNCH_CHECK();
    }
    lf_words_kernel<<<dim3(sb128w, sb128h), 288, 0, st>>>(G, cells, cells + plane, cdef_idx, masks);
    RB_LAUNCH_CHECK();
    return 0;
}

}  // namespace rb200
