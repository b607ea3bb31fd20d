// Host-side helper of the intra wavefront: dependency levels of intra-predicted transform blocks.
//
// The reference reconstructs blocks in decode order, so an intra block simply finds its neighbours finished
// (rav1d_recon_b_intra, src/recon.rs:2402-3160).  The batch path runs all blocks of one dependency level in one
// launch; this function derives the levels from the same per-block facts the reference hands to
// rav1d_prepare_intra_edges (position, size, availability flags): a block needs the reconstructed pixels of the
// column left of it (and the corner), the row above it, the above-right extension when EDGE_TOP_HAS_RIGHT is set and
// the below-left extension when EDGE_LEFT_HAS_BOTTOM is set; a chroma-from-luma block also needs its own luma; a
// residual-only item needs the prediction underneath it.  level = 1 + the highest level among the intra items that
// own those 4x4 cells, 0 if there is none (inter blocks are reconstructed before every level).  Pure host code.
#include "common.cuh"
#include <algorithm>
#include <vector>

using namespace rb200;

extern "C" int rb200_intra_assign_levels(Rb200IntraItem *items, int n, int frame_w4, int frame_h4, int ss_hor, int ss_ver,
                                         int32_t *order, int32_t *level_counts, int max_levels, int *n_levels_out) {
    if (!items || n < 0 || frame_w4 < 1 || frame_h4 < 1 || !order || !level_counts || !n_levels_out)
        return set_error(-22, "intra_assign_levels: bad argument");
    const int pw[3] = { frame_w4, (frame_w4 + ss_hor) >> ss_hor, (frame_w4 + ss_hor) >> ss_hor };
    const int ph[3] = { frame_h4, (frame_h4 + ss_ver) >> ss_ver, (frame_h4 + ss_ver) >> ss_ver };
    std::vector<int32_t> lvl[3];
    for (int p = 0; p < 3; p++) lvl[p].assign((size_t)pw[p] * ph[p], -1);
    auto rect_max = [&](int p, int xa, int xb, int ya, int yb) {
        int m = -1;
        xa = std::max(xa, 0); ya = std::max(ya, 0); xb = std::min(xb, pw[p]); yb = std::min(yb, ph[p]);
        for (int y = ya; y < yb; y++)
            for (int x = xa; x < xb; x++) m = std::max(m, lvl[p][(size_t)y * pw[p] + x]);
        return m;
    };
    int n_levels = 0;
    for (int i = 0; i < n; i++) {
        Rb200IntraItem &it = items[i];
        const int p = it.plane;
        // a transform block may hang over the right / bottom picture edge (the reference reconstructs into the
        // padding of the 128-aligned allocation); only its origin has to lie inside
        if (p > 2 || it.tw4 < 1 || it.th4 < 1 || it.x4 >= pw[p] || it.y4 >= ph[p])
            return set_error(-22, "intra_assign_levels: item %d outside the picture", i);
        const int x = it.x4, y = it.y4, tw = it.tw4, th = it.th4;
        int dep = -1;
        if (it.mode == 14) {
            dep = -1;                                                    // palette: no neighbours
        } else if (it.mode == 15) {
            dep = rect_max(p, x, x + tw, y, y + th);                     // residual on top of an earlier prediction
        } else if (it.mode == 16) {                                      // intra block copy: the source area (+ 1 pixel
            const int sx = (int16_t)it.w4_end, sy = (int16_t)it.h4_end;  // right / below for the bilinear filter)
            const int xa = std::min(std::max(sx >> 2, 0), pw[p] - 1), ya = std::min(std::max(sy >> 2, 0), ph[p] - 1);
            dep = rect_max(p, xa, std::max((sx + tw * 4 + 4) >> 2, xa + 1), ya, std::max((sy + th * 4 + 4) >> 2, ya + 1));
        } else {
            dep = std::max(rect_max(p, x - 1, x, y - 1, y + th), rect_max(p, x, x + tw, y - 1, y));
            const bool plain = !(it.flags & 64) && !(p && it.mode == 13);
            if (plain && (it.flags & 4)) dep = std::max(dep, rect_max(p, x + tw, x + 2 * tw, y - 1, y));
            if (plain && (it.flags & 8)) dep = std::max(dep, rect_max(p, x - 1, x, y + th, y + 2 * th));
            if (p && it.mode == 13)                                      // chroma from luma: the block's own luma
                dep = std::max(dep, rect_max(0, x << ss_hor, (x + tw) << ss_hor, y << ss_ver, (y + th) << ss_ver));
        }
        const int level = dep + 1;
        if (level > 0xffff || level >= max_levels) return set_error(-34, "intra_assign_levels: more than %d levels", max_levels);
        it.level = (uint16_t)level;
        n_levels = std::max(n_levels, level + 1);
        for (int yy = y; yy < std::min(y + th, ph[p]); yy++)
            for (int xx = x; xx < std::min(x + tw, pw[p]); xx++) lvl[p][(size_t)yy * pw[p] + xx] = level;
    }
    // stable counting sort by level
    std::fill(level_counts, level_counts + n_levels, 0);
    for (int i = 0; i < n; i++) level_counts[items[i].level]++;
    std::vector<int32_t> start((size_t)n_levels + 1, 0);
    for (int l = 0; l < n_levels; l++) start[l + 1] = start[l] + level_counts[l];
    for (int i = 0; i < n; i++) order[start[items[i].level]++] = i;
    *n_levels_out = n_levels;
    return 0;
}
