// Deblocking loop filter for sm_100a.
//
// Replaces Rav1dLoopFilterDSPContext.loop_filter_sb[2][2] (src/loopfilter.rs:20-34;
// core `loop_filter` src/loopfilter.rs:397 == src/loopfilter_tmpl.c:36-161, sb
// drivers src/loopfilter_tmpl.c:163-240) and, at frame level, the per-sbrow drivers
// rav1d_loopfilter_sbrow_cols / _rows (src/lf_apply.rs:597,763 ==
// src/lf_apply_tmpl.c:327-466).
//
// Mapping: one thread filters one line (one row of a column edge, one column of
// a row edge) of one 4-pixel edge unit.  Within a pass every edge is independent
// of every other (read / write sets never overlap because the filter length is
// bounded by the transform size on both sides), so a whole plane is
// one launch per direction: all column edges first, then all row edges, which
// is the order the reference's sbrow loop produces.  Threads of a warp are laid
// along the edge-normal for column edges and along the edge for row edges so
// that global accesses stay contiguous per row.
#include "common.cuh"

namespace rb200 {

// Filter one line across an edge.  p points at q0; `sb` = element stride across the edge.
// E, I, H already scaled by << (bpc - 8).   src/loopfilter_tmpl.c:48-160
template <typename BD>
__device__ __forceinline__ void lf_line(typename BD::pixel *p, const int64_t sb, int E, int I, int H, const int wd,
                                        const int bdmin8, const int bdmax) {
    using pixel = typename BD::pixel;
    auto A = [](int v) { return v < 0 ? -v : v; };
    const int F = 1 << bdmin8;
    int p6, p5, p4, p3, p2, q2, q3, q4, q5, q6;
    const int p1 = p[sb * -2], p0 = p[sb * -1], q0 = p[0], q1 = p[sb];
    bool fm = A(p1 - p0) <= I && A(q1 - q0) <= I && A(p0 - q0) * 2 + (A(p1 - q1) >> 1) <= E;
    if (wd > 4) {
        p2 = p[sb * -3]; q2 = p[sb * 2];
        fm = fm && A(p2 - p1) <= I && A(q2 - q1) <= I;
        if (wd > 6) {
            p3 = p[sb * -4]; q3 = p[sb * 3];
            fm = fm && A(p3 - p2) <= I && A(q3 - q2) <= I;
        }
    }
    if (!fm) return;
    bool flat8out = false, flat8in = false;
    if (wd >= 16) {
        p6 = p[sb * -7]; p5 = p[sb * -6]; p4 = p[sb * -5];
        q4 = p[sb * 4]; q5 = p[sb * 5]; q6 = p[sb * 6];
        flat8out = A(p6 - p0) <= F && A(p5 - p0) <= F && A(p4 - p0) <= F && A(q4 - q0) <= F && A(q5 - q0) <= F &&
                   A(q6 - q0) <= F;
    }
    if (wd >= 6) flat8in = A(p2 - p0) <= F && A(p1 - p0) <= F && A(q1 - q0) <= F && A(q2 - q0) <= F;
    if (wd >= 8) flat8in = flat8in && A(p3 - p0) <= F && A(q3 - q0) <= F;

    if (wd >= 16 && flat8out && flat8in) {
        p[sb * -6] = (pixel)((p6 + p6 + p6 + p6 + p6 + p6 * 2 + p5 * 2 + p4 * 2 + p3 + p2 + p1 + p0 + q0 + 8) >> 4);
        p[sb * -5] = (pixel)((p6 + p6 + p6 + p6 + p6 + p5 * 2 + p4 * 2 + p3 * 2 + p2 + p1 + p0 + q0 + q1 + 8) >> 4);
        p[sb * -4] = (pixel)((p6 + p6 + p6 + p6 + p5 + p4 * 2 + p3 * 2 + p2 * 2 + p1 + p0 + q0 + q1 + q2 + 8) >> 4);
        p[sb * -3] = (pixel)((p6 + p6 + p6 + p5 + p4 + p3 * 2 + p2 * 2 + p1 * 2 + p0 + q0 + q1 + q2 + q3 + 8) >> 4);
        p[sb * -2] = (pixel)((p6 + p6 + p5 + p4 + p3 + p2 * 2 + p1 * 2 + p0 * 2 + q0 + q1 + q2 + q3 + q4 + 8) >> 4);
        p[sb * -1] = (pixel)((p6 + p5 + p4 + p3 + p2 + p1 * 2 + p0 * 2 + q0 * 2 + q1 + q2 + q3 + q4 + q5 + 8) >> 4);
        p[0]       = (pixel)((p5 + p4 + p3 + p2 + p1 + p0 * 2 + q0 * 2 + q1 * 2 + q2 + q3 + q4 + q5 + q6 + 8) >> 4);
        p[sb * 1]  = (pixel)((p4 + p3 + p2 + p1 + p0 + q0 * 2 + q1 * 2 + q2 * 2 + q3 + q4 + q5 + q6 + q6 + 8) >> 4);
        p[sb * 2]  = (pixel)((p3 + p2 + p1 + p0 + q0 + q1 * 2 + q2 * 2 + q3 * 2 + q4 + q5 + q6 + q6 + q6 + 8) >> 4);
        p[sb * 3]  = (pixel)((p2 + p1 + p0 + q0 + q1 + q2 * 2 + q3 * 2 + q4 * 2 + q5 + q6 + q6 + q6 + q6 + 8) >> 4);
        p[sb * 4]  = (pixel)((p1 + p0 + q0 + q1 + q2 + q3 * 2 + q4 * 2 + q5 * 2 + q6 + q6 + q6 + q6 + q6 + 8) >> 4);
        p[sb * 5]  = (pixel)((p0 + q0 + q1 + q2 + q3 + q4 * 2 + q5 * 2 + q6 * 2 + q6 + q6 + q6 + q6 + q6 + 8) >> 4);
    } else if (wd >= 8 && flat8in) {
        p[sb * -3] = (pixel)((p3 + p3 + p3 + 2 * p2 + p1 + p0 + q0 + 4) >> 3);
        p[sb * -2] = (pixel)((p3 + p3 + p2 + 2 * p1 + p0 + q0 + q1 + 4) >> 3);
        p[sb * -1] = (pixel)((p3 + p2 + p1 + 2 * p0 + q0 + q1 + q2 + 4) >> 3);
        p[0]       = (pixel)((p2 + p1 + p0 + 2 * q0 + q1 + q2 + q3 + 4) >> 3);
        p[sb * 1]  = (pixel)((p1 + p0 + q0 + 2 * q1 + q2 + q3 + q3 + 4) >> 3);
        p[sb * 2]  = (pixel)((p0 + q0 + q1 + 2 * q2 + q3 + q3 + q3 + 4) >> 3);
    } else if (wd == 6 && flat8in) {
        p[sb * -2] = (pixel)((p2 + 2 * p2 + 2 * p1 + 2 * p0 + q0 + 4) >> 3);
        p[sb * -1] = (pixel)((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
        p[0]       = (pixel)((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
        p[sb * 1]  = (pixel)((p0 + 2 * q0 + 2 * q1 + 2 * q2 + q2 + 4) >> 3);
    } else {
        const bool hev = A(p1 - p0) > H || A(q1 - q0) > H;
        const int lo = -128 * (1 << bdmin8), hi = 128 * (1 << bdmin8) - 1;
        if (hev) {
            int f = iclip(p1 - q1, lo, hi);
            f = iclip(3 * (q0 - p0) + f, lo, hi);
            const int f1 = imin(f + 4, hi) >> 3, f2 = imin(f + 3, hi) >> 3;
            p[sb * -1] = (pixel)iclip(p0 + f2, 0, bdmax);
            p[0] = (pixel)iclip(q0 - f1, 0, bdmax);
        } else {
            int f = iclip(3 * (q0 - p0), lo, hi);
            const int f1 = imin(f + 4, hi) >> 3, f2 = imin(f + 3, hi) >> 3;
            p[sb * -1] = (pixel)iclip(p0 + f2, 0, bdmax);
            p[0] = (pixel)iclip(q0 - f1, 0, bdmax);
            f = (f1 + 1) >> 1;
            p[sb * -2] = (pixel)iclip(p1 + f, 0, bdmax);
            p[sb * 1] = (pixel)iclip(q1 - f, 0, bdmax);
        }
    }
}

struct LfGeom {
    int w4, h4;         // plane size in 4-pixel units that are filtered (luma: f.w4/f.h4; chroma: rounded-up halves)
    int sb128w;
    int b4_stride;
    int ss_hor, ss_ver; // of this plane
    int lvl_idx;        // byte of level[][4]: luma col 0, luma row 1, u 2, v 3
    int uv;             // 0 luma, 1 chroma
};

// Filter width index of the edge unit (x4, y4) from the Av1Filter masks, or -1.
// Luma:   filter_y[dir][a][idx][half] bit b;  chroma: filter_uv[dir][a][idx][half] bit b, where for
// column edges a = x4 in sb128, bits run over y4; for row edges a = y4 in sb128, bits over x4
// (src/lf_apply_tmpl.c:176-325: hmask/vmask assembly).
__device__ __forceinline__ int lf_mask_idx(const Rb200Av1Filter *__restrict__ masks, const LfGeom &g, int dir, int x4,
                                           int y4) {
    const int shx = 5 - g.ss_hor, shy = 5 - g.ss_ver;
    const Rb200Av1Filter &m = masks[(y4 >> shy) * g.sb128w + (x4 >> shx)];
    const int xi = x4 & ((1 << shx) - 1), yi = y4 & ((1 << shy) - 1);
    // `a` indexes the array, `b` is the bit position along the edge; halves hold 16 >> ss bits each
    const int a = dir == 0 ? xi : yi, b = dir == 0 ? yi : xi;
    const int hsh = dir == 0 ? 4 - g.ss_ver : 4 - g.ss_hor;
    const int half = b >> hsh, bit = b & ((1 << hsh) - 1);
    if (!g.uv) {
        if ((m.filter_y[dir][a][2][half] >> bit) & 1) return 2;
        if ((m.filter_y[dir][a][1][half] >> bit) & 1) return 1;
        if ((m.filter_y[dir][a][0][half] >> bit) & 1) return 0;
    } else {
        if ((m.filter_uv[dir][a][1][half] >> bit) & 1) return 1;
        if ((m.filter_uv[dir][a][0][half] >> bit) & 1) return 0;
    }
    return -1;
}

// Register form of the edge filter for one line: P[i] = p_i, Q[i] = q_i (i = 0 next to the edge).
// Returns how many samples on each side may have changed (0, 1, 2, 3 or 6).  Same arithmetic as
// lf_line / src/loopfilter_tmpl.c:48-160.
__device__ __forceinline__ int lf_line_regs(int *P, int *Q, int E, int I, int H, const int wd, const int bdmin8,
                                            const int bdmax) {
    auto A = [](int v) { return v < 0 ? -v : v; };
    const int F = 1 << bdmin8;
    const int p0 = P[0], p1 = P[1], q0 = Q[0], q1 = Q[1];
    bool fm = A(p1 - p0) <= I && A(q1 - q0) <= I && A(p0 - q0) * 2 + (A(p1 - q1) >> 1) <= E;
    if (wd > 4) {
        fm = fm && A(P[2] - p1) <= I && A(Q[2] - q1) <= I;
        if (wd > 6) fm = fm && A(P[3] - P[2]) <= I && A(Q[3] - Q[2]) <= I;
    }
    if (!fm) return 0;
    bool flat8out = false, flat8in = false;
    if (wd >= 16)
        flat8out = A(P[6] - p0) <= F && A(P[5] - p0) <= F && A(P[4] - p0) <= F && A(Q[4] - q0) <= F && A(Q[5] - q0) <= F &&
                   A(Q[6] - q0) <= F;
    if (wd >= 6) flat8in = A(P[2] - p0) <= F && A(p1 - p0) <= F && A(q1 - q0) <= F && A(Q[2] - q0) <= F;
    if (wd >= 8) flat8in = flat8in && A(P[3] - p0) <= F && A(Q[3] - q0) <= F;
    if (wd >= 16 && flat8out && flat8in) {
        const int p6 = P[6], p5 = P[5], p4 = P[4], p3 = P[3], p2 = P[2], q2 = Q[2], q3 = Q[3], q4 = Q[4], q5 = Q[5], q6 = Q[6];
        P[5] = (p6 + p6 + p6 + p6 + p6 + p6 * 2 + p5 * 2 + p4 * 2 + p3 + p2 + p1 + p0 + q0 + 8) >> 4;
        P[4] = (p6 + p6 + p6 + p6 + p6 + p5 * 2 + p4 * 2 + p3 * 2 + p2 + p1 + p0 + q0 + q1 + 8) >> 4;
        P[3] = (p6 + p6 + p6 + p6 + p5 + p4 * 2 + p3 * 2 + p2 * 2 + p1 + p0 + q0 + q1 + q2 + 8) >> 4;
        P[2] = (p6 + p6 + p6 + p5 + p4 + p3 * 2 + p2 * 2 + p1 * 2 + p0 + q0 + q1 + q2 + q3 + 8) >> 4;
        P[1] = (p6 + p6 + p5 + p4 + p3 + p2 * 2 + p1 * 2 + p0 * 2 + q0 + q1 + q2 + q3 + q4 + 8) >> 4;
        P[0] = (p6 + p5 + p4 + p3 + p2 + p1 * 2 + p0 * 2 + q0 * 2 + q1 + q2 + q3 + q4 + q5 + 8) >> 4;
        Q[0] = (p5 + p4 + p3 + p2 + p1 + p0 * 2 + q0 * 2 + q1 * 2 + q2 + q3 + q4 + q5 + q6 + 8) >> 4;
        Q[1] = (p4 + p3 + p2 + p1 + p0 + q0 * 2 + q1 * 2 + q2 * 2 + q3 + q4 + q5 + q6 + q6 + 8) >> 4;
        Q[2] = (p3 + p2 + p1 + p0 + q0 + q1 * 2 + q2 * 2 + q3 * 2 + q4 + q5 + q6 + q6 + q6 + 8) >> 4;
        Q[3] = (p2 + p1 + p0 + q0 + q1 + q2 * 2 + q3 * 2 + q4 * 2 + q5 + q6 + q6 + q6 + q6 + 8) >> 4;
        Q[4] = (p1 + p0 + q0 + q1 + q2 + q3 * 2 + q4 * 2 + q5 * 2 + q6 + q6 + q6 + q6 + q6 + 8) >> 4;
        Q[5] = (p0 + q0 + q1 + q2 + q3 + q4 * 2 + q5 * 2 + q6 * 2 + q6 + q6 + q6 + q6 + q6 + 8) >> 4;
        return 6;
    }
    if (wd >= 8 && flat8in) {
        const int p3 = P[3], p2 = P[2], q2 = Q[2], q3 = Q[3];
        P[2] = (p3 + p3 + p3 + 2 * p2 + p1 + p0 + q0 + 4) >> 3;
        P[1] = (p3 + p3 + p2 + 2 * p1 + p0 + q0 + q1 + 4) >> 3;
        P[0] = (p3 + p2 + p1 + 2 * p0 + q0 + q1 + q2 + 4) >> 3;
        Q[0] = (p2 + p1 + p0 + 2 * q0 + q1 + q2 + q3 + 4) >> 3;
        Q[1] = (p1 + p0 + q0 + 2 * q1 + q2 + q3 + q3 + 4) >> 3;
        Q[2] = (p0 + q0 + q1 + 2 * q2 + q3 + q3 + q3 + 4) >> 3;
        return 3;
    }
    if (wd == 6 && flat8in) {
        const int p2 = P[2], q2 = Q[2];
        P[1] = (p2 + 2 * p2 + 2 * p1 + 2 * p0 + q0 + 4) >> 3;
        P[0] = (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3;
        Q[0] = (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3;
        Q[1] = (p0 + 2 * q0 + 2 * q1 + 2 * q2 + q2 + 4) >> 3;
        return 2;
    }
    const bool hev = A(p1 - p0) > H || A(q1 - q0) > H;
    const int lo = -128 * (1 << bdmin8), hi = 128 * (1 << bdmin8) - 1;
    if (hev) {
        int f = iclip(p1 - q1, lo, hi);
        f = iclip(3 * (q0 - p0) + f, lo, hi);
        const int f1 = imin(f + 4, hi) >> 3, f2 = imin(f + 3, hi) >> 3;
        P[0] = iclip(p0 + f2, 0, bdmax);
        Q[0] = iclip(q0 - f1, 0, bdmax);
        return 1;
    }
    int f = iclip(3 * (q0 - p0), lo, hi);
    const int f1 = imin(f + 4, hi) >> 3, f2 = imin(f + 3, hi) >> 3;
    P[0] = iclip(p0 + f2, 0, bdmax);
    Q[0] = iclip(q0 - f1, 0, bdmax);
    f = (f1 + 1) >> 1;
    P[1] = iclip(p1 + f, 0, bdmax);
    Q[1] = iclip(q1 - f, 0, bdmax);
    return 2;
}

struct LfPlaneSet {
    uint8_t *plane[3];
    int64_t stride[3];
    LfGeom g[3];
    int unit_start[4];   // prefix sums of the unit counts of the planes in this launch
    int y4_first[3];     // first 4-pixel unit row processed in each plane (band restriction)
    int n_planes;
};

// One thread per 4-pixel edge unit, 4 lines each, all planes of a pass in one launch.
// DIR 0: column edges -- per line the thread loads the 4/8/16 pixels straddling the edge with
// aligned vector loads and writes back only what the filter may modify.
// DIR 1: row edges -- the thread owns 4 adjacent columns; every row of the stencil is one
// aligned 4-pixel load, coalesced across the warp.
template <typename BD, int DIR>
__global__ void __launch_bounds__(128, 8)
deblock_units_kernel(LfPlaneSet S, const Rb200Av1Filter *__restrict__ masks, const uint8_t (*__restrict__ lvl)[4],
                     const Rb200Av1FilterLUT *__restrict__ lut, int bdmax) {
    using pixel = typename BD::pixel;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= S.unit_start[S.n_planes]) return;
    const int pi = (S.n_planes > 1 && gid >= S.unit_start[1]) ? ((S.n_planes > 2 && gid >= S.unit_start[2]) ? 2 : 1) : 0;
    const LfGeom g = pi == 0 ? S.g[0] : (pi == 1 ? S.g[1] : S.g[2]);
    uint8_t *plane = pi == 0 ? S.plane[0] : (pi == 1 ? S.plane[1] : S.plane[2]);
    const int64_t stride = pi == 0 ? S.stride[0] : (pi == 1 ? S.stride[1] : S.stride[2]);
    const int u = gid - (pi == 0 ? 0 : (pi == 1 ? S.unit_start[1] : S.unit_start[2]));
    const int yl = u / g.w4, x4 = u - yl * g.w4;
    const int y4 = yl + (pi == 0 ? S.y4_first[0] : (pi == 1 ? S.y4_first[1] : S.y4_first[2]));
    if (DIR == 0 ? x4 == 0 : y4 == 0) return;  // have_left / have_top
    const int64_t ps = stride / (int64_t)sizeof(pixel);
    pixel *base = (pixel *)plane + (int64_t)(y4 * 4) * ps + x4 * 4;   // q0 of line 0
    // Row-edge pass: the 8 rows every filter width needs are requested before the mask / level lookups
    // they would otherwise wait behind (the pass is latency-bound; units without an edge waste the loads,
    // which L2 serves).
    constexpr int WPR = BD::hbd ? 2 : 1;            // 32-bit words per 4-pixel row
    unsigned rows[16][WPR];
    if (DIR == 1) {
#pragma unroll
        for (int r = 4; r < 12; r++) {
            const pixel *p = base + (int64_t)(r - 8) * ps;
            if (BD::hbd) { const uint2 q = *(const uint2 *)p; rows[r][0] = q.x; rows[r][WPR - 1] = q.y; }
            else rows[r][0] = *(const unsigned *)p;
        }
    }
    const int idx = lf_mask_idx(masks, g, DIR, x4, y4);
    if (idx < 0) return;
    const uint8_t(*l)[4] = lvl + (int64_t)y4 * g.b4_stride + x4;
    int L = l[0][g.lvl_idx];
    if (!L) L = DIR == 0 ? l[-1][g.lvl_idx] : l[-(int64_t)g.b4_stride][g.lvl_idx];
    if (!L) return;
    const int bdmin8 = BD::hbd ? bpc_from_max(bdmax) - 8 : 0;
    const int H = (L >> 4) << bdmin8, E = (int)lut->e[L] << bdmin8, I = (int)lut->i[L] << bdmin8;
    const int wd = g.uv ? 4 + 2 * idx : 4 << idx;
    const int ng = wd == 16 ? 2 : 1;           // 4-pixel groups loaded on each side of the edge
    auto load4 = [](const pixel *p, int *v) {
        if (BD::hbd) { const uint2 q = *(const uint2 *)p; v[0] = q.x & 0xffff; v[1] = q.x >> 16; v[2] = q.y & 0xffff; v[3] = q.y >> 16; }
        else { const unsigned q = *(const unsigned *)p; v[0] = q & 0xff; v[1] = (q >> 8) & 0xff; v[2] = (q >> 16) & 0xff; v[3] = q >> 24; }
    };
    if (DIR == 0) {
#pragma unroll 1
        for (int line = 0; line < 4; line++) {
            pixel *p = base + (int64_t)line * ps;
            int P[8], Q[8], t[4];
            load4(p - 4, t); P[0] = t[3]; P[1] = t[2]; P[2] = t[1]; P[3] = t[0];
            load4(p, Q);
            if (ng == 2) { load4(p - 8, t); P[4] = t[3]; P[5] = t[2]; P[6] = t[1]; P[7] = t[0]; load4(p + 4, Q + 4); }
            const int n = lf_line_regs(P, Q, E, I, H, wd, bdmin8, bdmax);
#pragma unroll
            for (int i = 0; i < 6; i++)
                if (i < n) { p[-1 - i] = (pixel)P[i]; p[i] = (pixel)Q[i]; }
        }
    } else {
        // rows y-4ng .. y+4ng-1, 4 columns each, kept PACKED (one or two registers per row) to leave room
        // for more resident warps; column c is unpacked, filtered and re-packed
        if (ng == 2) {
#pragma unroll
            for (int r = 0; r < 16; r++) {
                if (r >= 4 && r < 12) continue;
                const pixel *p = base + (int64_t)(r - 8) * ps;
                if (BD::hbd) { const uint2 q = *(const uint2 *)p; rows[r][0] = q.x; rows[r][WPR - 1] = q.y; }
                else rows[r][0] = *(const unsigned *)p;
            }
        }
        auto get = [&](int r, int c) -> int {
            if (BD::hbd) return (int)((rows[r][c >> 1] >> (16 * (c & 1))) & 0xffff);
            return (int)((rows[r][0] >> (8 * c)) & 0xff);
        };
        auto put = [&](int r, int c, int v) {
            if (BD::hbd) rows[r][c >> 1] = (rows[r][c >> 1] & ~(0xffffu << (16 * (c & 1)))) | ((unsigned)v << (16 * (c & 1)));
            else rows[r][0] = (rows[r][0] & ~(0xffu << (8 * c))) | ((unsigned)v << (8 * c));
        };
        int nmax = 0;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            int P[8], Q[8];
#pragma unroll
            for (int i = 0; i < 8; i++) { P[i] = get(7 - i, c); Q[i] = get(8 + i, c); }
            const int n = lf_line_regs(P, Q, E, I, H, wd, bdmin8, bdmax);
            nmax = imax(nmax, n);
#pragma unroll
            for (int i = 0; i < 6; i++)
                if (i < n) { put(7 - i, c, P[i]); put(8 + i, c, Q[i]); }
        }
#pragma unroll
        for (int i = 0; i < 6; i++) {
            if (i < nmax) {
                pixel *pu = base + (int64_t)(-1 - i) * ps, *pd = base + (int64_t)i * ps;
                if (BD::hbd) { *(uint2 *)pu = make_uint2(rows[7 - i][0], rows[7 - i][WPR - 1]); *(uint2 *)pd = make_uint2(rows[8 + i][0], rows[8 + i][WPR - 1]); }
                else { *(unsigned *)pu = rows[7 - i][0]; *(unsigned *)pd = rows[8 + i][0]; }
            }
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
    pixel *p = (pixel *)dst;
    if (dir == 0) lf_line<BD>(p + (int64_t)(u * 4 + line) * ps, 1, E, I, H, wd, bdmin8, bdmax);
    else lf_line<BD>(p + u * 4 + line, ps, E, I, H, wd, bdmin8, bdmax);
}

// Whole-frame deblock: column edges of every plane (one launch), then row edges (one launch).
// y4b / y4e: luma 4-pixel unit rows whose ROW edges are filtered (the whole picture: 0, h4); the
// column-edge pass covers two more unit rows on each side, which is everything those row edges read.
int deblock_frame_launch(const Rb200Planes &pl, int n_planes, int w4, int h4, int sb128w, int b4_stride, int ss_hor,
                         int ss_ver, bool do_uv, const Rb200Av1Filter *masks, const uint8_t (*lvl)[4],
                         const Rb200Av1FilterLUT *lut, int bdmax, cudaStream_t st, int *launches, int y4b, int y4e) {
    for (int dir = 0; dir < 2; dir++) {
        LfPlaneSet S = {};
        int n = 0, total = 0;
        for (int p = 0; p < n_planes; p++) {
            if (p && !do_uv) continue;
            LfGeom &g = S.g[n];
            g.uv = p ? 1 : 0;
            g.ss_hor = p ? ss_hor : 0; g.ss_ver = p ? ss_ver : 0;
            g.w4 = (w4 + g.ss_hor) >> g.ss_hor; g.h4 = (h4 + g.ss_ver) >> g.ss_ver;
            g.sb128w = sb128w; g.b4_stride = b4_stride;
            g.lvl_idx = p ? 1 + p : dir;
            S.plane[n] = (uint8_t *)pl.data[p]; S.stride[n] = pl.stride[p];
            S.unit_start[n] = total;
            // band in this plane's unit rows; chroma rows round outwards
            int b = dir ? y4b : y4b - 2, e = dir ? y4e : y4e + 2;
            b = imax(b >> g.ss_ver, 0); e = imin((e + g.ss_ver) >> g.ss_ver, g.h4);
            S.y4_first[n] = b;
            total += g.w4 * imax(e - b, 0);
            n++;
        }
        S.unit_start[n] = total;
        for (int k = n + 1; k < 4; k++) S.unit_start[k] = total;
        S.n_planes = n;
        const int grid = (total + 127) / 128;
#define L(BD, D) deblock_units_kernel<BD, D><<<grid, 128, 0, st>>>(S, masks, lvl, lut, bdmax)
        if (bdmax > 255) { if (dir) L(BD16, 1); else L(BD16, 0); } else { if (dir) L(BD8, 1); else L(BD8, 0); }
#undef L
        RB_LAUNCH_CHECK();
        if (launches) ++*launches;
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
