// AV1-normative constant tables on the device (values generated into
// tables_data.inc by tools/gen_tables.py).  Tables whose index is uniform
// across a warp live in __constant__ memory (broadcast); per-lane indexed ones
// are plain __device__ arrays served by L1.
#pragma once
#include <stdint.h>

namespace rb200 {
namespace tab {

#define RAV1D_B200_TABLE(type, name, n) static __device__ const type name[n]
#include "tables_data.inc"
#undef RAV1D_B200_TABLE

// 8-tap sub-pel filter row.  set: 0 regular, 1 smooth, 2 sharp, 3 regular4, 4 smooth4, 5 (unused)
// (dav1d_mc_subpel_filters, src/tables.rs:748; selection rule get_filter src/mc.rs:119-127).
__device__ __forceinline__ const int8_t *subpel(int set, int phase_minus1) {
    return k_subpel_filters + (set * 15 + phase_minus1) * 8;
}

// CDEF tap offsets as (dy, dx) for direction d (0..7) and tap distance k (0..1);
// the reference stores them as dy*12+dx (dav1d_cdef_directions, src/tables.rs:396).
__device__ __forceinline__ void cdef_dir_off(int d, int k, int &dy, int &dx) {
    // packed nibbles, biased by +2 so that -2..2 fits:  [d][k] = (dy+2) | (dx+2) << 4
    // d:   0          1         2         3         4         5         6         7
    // k=0 (-1,1)    (0,1)     (0,1)     (0,1)     (1,1)     (1,0)     (1,0)     (1,0)
    // k=1 (-2,2)    (-1,2)    (0,2)     (1,2)     (2,2)     (2,1)     (2,0)     (2,-1)
    constexpr unsigned DY0 = 0x33332221u;       // nibble d = dy+2 for k=0 : 1,2,2,2,3,3,3,3
    constexpr unsigned DX0 = 0x22233333u;       // dx+2 for k=0 : 3,3,3,3,3,2,2,2
    constexpr unsigned DY1 = 0x44443210u;       // dy+2 for k=1 : 0,1,2,3,4,4,4,4
    constexpr unsigned DX1 = 0x12344444u;       // dx+2 for k=1 : 4,4,4,4,4,3,2,1
    const int sh = d * 4;
    dy = (int)(((k ? DY1 : DY0) >> sh) & 15) - 2;
    dx = (int)(((k ? DX1 : DX0) >> sh) & 15) - 2;
}

}  // namespace tab
}  // namespace rb200
