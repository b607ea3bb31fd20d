/*
 * oracle/ref_lfmask.c -- TEST INFRASTRUCTURE ONLY (linked into oracle/_ref/libdav1d_ref.so).
 *
 * Oracle of SURVEY 8 row f2: replays the REFERENCE'S OWN dav1d_create_lf_mask_intra / dav1d_create_lf_mask_inter
 * (src/lf_mask.c:286-406, compiled in place) over a list of block records in decode order, exactly as decode_b
 * calls them (src/decode.c:1260-1271,1926-1947): the above context is per 128-pixel column (BlockContext
 * f->a[], tx_lpf_y / tx_lpf_uv reset to 2 / 1 at the start of a tile, src/decode.c:2451-2452), the left context
 * is reset at the start of every superblock row (dav1d_decode_tile_sbrow, src/decode.c:2627).  One tile, so none
 * of the tile-edge fix-ups of src/lf_apply_tmpl.c:331-400 apply.  The noskip mask is decode_b's own loop
 * (src/decode.c:1996-2005) restated.
 */
#include "config.h"

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "src/internal.h"
#include "src/lf_mask.h"
#include "src/levels.h"
#include "src/tables.h"

#include "../include/rav1d_b200.h"

int ref_lf_build(const Rb200LfBlock *blk, int n, int w4, int h4, int layout, int sb128, int b4_stride, int sb128w,
                 int sb128h, Av1Filter *masks, uint8_t (*lvl)[4]) {
    const int ss_ver = layout == DAV1D_PIXEL_LAYOUT_I420, ss_hor = layout != DAV1D_PIXEL_LAYOUT_I444;
    const int sb_shift = 4 + !!sb128;
    uint8_t (*a_y)[32] = malloc((size_t)sb128w * 32), (*a_uv)[32] = malloc((size_t)sb128w * 32);
    uint8_t l_y[32], l_uv[32];
    if (!a_y || !a_uv) return -1;
    memset(a_y, 2, (size_t)sb128w * 32);
    memset(a_uv, 1, (size_t)sb128w * 32);
    for (int i = 0; i < sb128w * sb128h; i++) {
        memset(masks[i].filter_y, 0, sizeof(masks[i].filter_y));
        memset(masks[i].filter_uv, 0, sizeof(masks[i].filter_uv));
        memset(masks[i].noskip_mask, 0, sizeof(masks[i].noskip_mask));
    }
    int cur_sby = -1;
    for (int i = 0; i < n; i++) {
        const Rb200LfBlock *const b = &blk[i];
        const int bx = b->bx, by = b->by, bx4 = bx & 31, by4 = by & 31;
        const int cbx4 = bx4 >> ss_hor, cby4 = by4 >> ss_ver;
        if ((by >> sb_shift) != cur_sby) {
            cur_sby = by >> sb_shift;
            memset(l_y, 2, 32);
            memset(l_uv, 1, 32);
        }
        Av1Filter *const m = &masks[(by >> 5) * sb128w + (bx >> 5)];
        uint8_t fl[4][8][2];
        memset(fl, 0, sizeof(fl));
        for (int k = 0; k < 4; k++) fl[k][0][0] = b->lvl[k];
        const int has_chroma = !!(b->flags & RB200_LFB_HAS_CHROMA) && layout != DAV1D_PIXEL_LAYOUT_I400;
        uint8_t *const auv = has_chroma ? &a_uv[bx >> 5][cbx4] : NULL, *const luv = has_chroma ? &l_uv[cby4] : NULL;
        if (b->flags & RB200_LFB_INTRA)
            dav1d_create_lf_mask_intra(m, lvl, b4_stride, (const uint8_t (*)[8][2])fl, bx, by, w4, h4, b->bs, b->ytx, b->uvtx,
                                       layout, &a_y[bx >> 5][bx4], &l_y[by4], auv, luv);
        else
            dav1d_create_lf_mask_inter(m, lvl, b4_stride, (const uint8_t (*)[8][2])fl, bx, by, w4, h4,
                                       !!(b->flags & RB200_LFB_SKIP), b->bs, b->ytx, b->tx_split, b->uvtx, layout,
                                       &a_y[bx >> 5][bx4], &l_y[by4], auv, luv);
        if (!(b->flags & RB200_LFB_SKIP)) {     /* src/decode.c:1996-2005 */
            const int bw4 = dav1d_block_dimensions[b->bs][0], bh4 = dav1d_block_dimensions[b->bs][1];
            uint16_t (*noskip_mask)[2] = &m->noskip_mask[by4 >> 1];
            const unsigned mask = (~0U >> (32 - bw4)) << (bx4 & 15);
            const int bx_idx = (bx4 & 16) >> 4;
            for (int y = 0; y < bh4; y += 2, noskip_mask++) {
                (*noskip_mask)[bx_idx] |= mask;
                if (bw4 == 32) (*noskip_mask)[1] |= mask;
            }
        }
    }
    free(a_y); free(a_uv);
    return 0;
}
