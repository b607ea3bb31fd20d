// Shared declarations of the rav1d_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "../../include/rav1d_b200.h"

struct CUtensorMap_st;   // <cuda.h>; tma.cuh

namespace rb200 {

// Bit-depth classes, mirroring the reference's BitDepth trait
// (include/common/bitdepth.rs:277-289 BitDepth8, :355-366 BitDepth16).
struct BD8 {
    using pixel = uint8_t;
    using coef = int16_t;
    using pair = uint16_t;      // two horizontally adjacent pixels
    static constexpr bool hbd = false;
    static __host__ __device__ __forceinline__ int lo(pair q) { return q & 0xff; }
    static __host__ __device__ __forceinline__ int hi(pair q) { return q >> 8; }
    static __host__ __device__ __forceinline__ pair pack(int a, int b) { return (pair)(a | (b << 8)); }
};
struct BD16 {
    using pixel = uint16_t;
    using coef = int32_t;
    using pair = uint32_t;
    static constexpr bool hbd = true;
    static __host__ __device__ __forceinline__ int lo(pair q) { return (int)(q & 0xffff); }
    static __host__ __device__ __forceinline__ int hi(pair q) { return (int)(q >> 16); }
    static __host__ __device__ __forceinline__ pair pack(int a, int b) { return (pair)a | ((pair)b << 16); }
};

__host__ __device__ __forceinline__ int imin(int a, int b) { return a < b ? a : b; }
__host__ __device__ __forceinline__ int imax(int a, int b) { return a > b ? a : b; }
// lo <= hi at every call site: max-then-min is the same value and compiles to two VIMNMX (or VIADDMNMX + VIMNMX
// when v is a sum) instead of compare + select + min
__host__ __device__ __forceinline__ int iclip(int v, int lo, int hi) { return imin(imax(v, lo), hi); }
__host__ __device__ __forceinline__ int ulog2(unsigned v) {
#ifdef __CUDA_ARCH__
    return 31 - __clz(v);
#else
    return 31 - __builtin_clz(v);
#endif
}
// bits per component from bitdepth_max (include/common/bitdepth.rs:590 bitdepth_from_max)
__host__ __device__ __forceinline__ int bpc_from_max(int bdmax) { return bdmax > 255 ? 32 -
#ifdef __CUDA_ARCH__
    __clz(bdmax)
#else
    __builtin_clz(bdmax)
#endif
    : 8; }

// Plane selection by branches: indexing a by-value kernel parameter struct with a
// runtime index would force a local-memory copy of it.
__host__ __device__ __forceinline__ uint8_t *plane_ptr(const Rb200Planes &p, int pl) {
    return (uint8_t *)(pl == 0 ? p.data[0] : (pl == 1 ? p.data[1] : p.data[2]));
}
__host__ __device__ __forceinline__ int64_t plane_stride(const Rb200Planes &p, int pl) {
    return pl == 0 ? p.stride[0] : (pl == 1 ? p.stride[1] : p.stride[2]);
}

// ---- stage launchers (defined in itx.cu, mc.cu, lf.cu, cdef.cu, lr.cu; used by frame.cu)
struct CdefFrameParams {
    int bw, bh;        // frame size in 4-pixel units (even)
    int sb128w;
    int ss_hor, ss_ver;
    int n_planes;
    int damping;       // frame_hdr.cdef.damping + bitdepth_min_8
    int bdmin8;
    int y_strength[8], uv_strength[8];
    int layout_422;
};
struct LrFrameParams {
    int w, h;             // plane size in pixels
    int ss_hor, ss_ver;   // of this plane
    int plane;
    int unit_log2;
    int sb128, sbh, sr_sb128w;
    int stripe_first, stripe_end;   // band restriction (all stripes: 0, 0)
};
int coef_gather_launch(const void *h_cf, void *d_cf, const Rb200ItxItem *d_items, int n, int bdmax, cudaStream_t st);
int itx_launch(int tx, const Rb200Planes &planes, const void *cf, const Rb200ItxItem *items, int n, int bdmax,
               cudaStream_t st);
// Tensor maps of up to 8 x 3 reference planes (mc.cu: McTmaMaps) and what each was encoded for; zero-initialise.
struct alignas(64) McRefMapCache {
    unsigned char maps[24 * 128];
    const void *base[24]; int64_t stride[24]; int w[24], h[24];
    unsigned mask;
};
int mc_batch_launch(const Rb200Planes &dst, const Rb200Planes *refs, int n_refs, int ref_w, int ref_h, int ss_hor,
                    int ss_ver, const Rb200McItem *d_items, int n, int bdmax, cudaStream_t st, int *counter = nullptr,
                    McRefMapCache *map_cache = nullptr);
int mc_comp_batch_launch(const Rb200Planes &dst, const Rb200Planes *refs, int n_refs, int ref_w, int ref_h, int layout,
                         const Rb200CompItem *d_items, int n, int bdmax, cudaStream_t st);
int mc_warp_batch_launch(const Rb200Planes &dst, const Rb200Planes *refs, int n_refs, int ref_w, int ref_h, int layout,
                         const Rb200WarpItem *d_items, int n, int bdmax, cudaStream_t st);
struct McRefDims { int w[8], h[8]; };   // luma size of each reference slot
int mc_scaled_batch_launch(const Rb200Planes &dst, const Rb200Planes *refs, int n_refs, const McRefDims &dims, int ss_hor,
                           int ss_ver, const Rb200McScaledItem *d_items, int n, int bdmax, cudaStream_t st);
int intra_levels_launch(const Rb200Planes &cur, const Rb200IntraItem *d_items, const int32_t *d_itx_of, const Rb200ItxItem *d_itx,
                        const void *cf, const uint8_t *d_pal, const int32_t *d_level_off, int n_levels, int max_items_per_level,
                        int frame_w4, int frame_h4, int ss_hor, int ss_ver, int bdmax, unsigned *d_sync, cudaStream_t st);
int intra_items_launch(const Rb200Planes &cur, const Rb200IntraItem *d_items, const int32_t *d_itx_of, const Rb200ItxItem *d_itx,
                       const void *cf, const uint8_t *d_pal, int n, int frame_w4, int frame_h4, int ss_hor, int ss_ver, int bdmax,
                       cudaStream_t st);
// super-resolution: one plane, all rows (src/mc.rs:1114-1172 resize)
int resize_plane_launch(void *dst, int64_t dstride, const void *src, int64_t sstride, int dst_w, int h, int src_w, int dx,
                        int mx0, int bdmax, cudaStream_t st);
int mc_obmc_batch_launch(const Rb200Planes &dst, const Rb200Planes *refs, int n_refs, int ref_w, int ref_h, int ss_hor,
                         int ss_ver, const Rb200McItem *d_items, int n, int bdmax, cudaStream_t st);
int deblock_frame_launch(const Rb200Planes &pl, int n_planes, int w4, int h4, int sb128w, int b4_stride, int ss_hor,
                         int ss_ver, bool do_uv, const Rb200Av1Filter *masks, const uint8_t (*lvl)[4],
                         const Rb200Av1FilterLUT *lut, int bdmax, cudaStream_t st, int *launches, int y4b, int y4e);
// lfmask.cu: Av1Filter masks + lf.level built from per-block records (cells: 2 * 32 sb128w * 32 sb128h bytes of scratch)
int lf_build_launch(const Rb200LfBlock *d_blocks, int n, int w4, int h4, int sb128w, int sb128h, int b4_stride, int ss_hor,
                    int ss_ver, int n_planes, uint8_t *cells, const int8_t *cdef_idx, Rb200Av1Filter *masks, uint8_t (*lvl)[4],
                    cudaStream_t st);
// blk_scratch: device, 8 bytes per 8x8 luma block ((bw / 2) * (bh / 2) records)
int cdef_frame_launch(const Rb200Planes &src, const Rb200Planes &dst, const CdefFrameParams &P,
                      const Rb200Av1Filter *masks, void *blk_scratch, int bdmax, cudaStream_t st, int t0, int t1,
                      const CUtensorMap_st *maps = nullptr);
int cdef_encode_maps(CUtensorMap_st maps[3], const Rb200Planes &src, const CdefFrameParams &P);
int lr_plane_launch(const uint8_t *cdef, const uint8_t *dbl, uint8_t *out, int64_t stride, const LrFrameParams &P,
                    const Rb200Av1Restoration *lrm, int bdmax, cudaStream_t st);

// film grain (fg.cu), used by the frame stage
struct FgApplyParams {
    int pw, ph;            // plane region in pixels
    int row0;              // block-row number of y == 0 (row_num of the per-call form)
    int off_row0, ncols;   // first block row held by the offsets table, its row pitch
    int sx, sy;            // subsampling of this plane (0 for luma)
    int chroma;            // 0 luma, 1 chroma
    int uv;                // chroma plane index 0/1
    int is_id;
    int luma_w;            // luma columns readable (lx + 1 is clamped to luma_w - 1)
    int overlap, clip, scaling_shift, csfl;
    int uv_mult, uv_luma_mult, uv_offset;
};
int fg_generate(void *lut, const void *lut_y, const Rb200FilmGrainData &d, int uv, int subx, int suby, int bdmax,
                cudaStream_t st, int n_luts = 1, int lut_pitch_bytes = 0);
int fg_apply(uint8_t *dst, const uint8_t *src, int64_t stride, const uint8_t *luma, int64_t luma_stride,
             const FgApplyParams &P, const uint8_t *scaling, const void *lut, const uint8_t *off, int bdmax,
             cudaStream_t st);
int fg_offsets(uint8_t *off, unsigned seed, int row0, int nrows, int ncols, cudaStream_t st);
struct FgPoints { uint8_t p[32]; };   // up to 14 (x, y) scaling points, passed by value
int fg_scaling(uint8_t *scaling, const uint8_t points[][2], int num, int bitdepth, cudaStream_t st);
FgApplyParams fg_params(const Rb200FilmGrainData &d, int chroma, int uv, int sx, int sy, int is_id);

// ---- error plumbing --------------------------------------------------
int set_error(int code, const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define RB_CUDA(call)                                                        \
    do {                                                                     \
        cudaError_t e_ = (call);                                             \
        if (e_ != cudaSuccess) return ::rb200::cuda_fail(e_, #call, __FILE__, __LINE__); \
    } while (0)
#define RB_LAUNCH_CHECK() RB_CUDA(cudaGetLastError())

// ---- per-thread staging used by the host-pointer (per-call) entry points ----
// The reference DSP functions are synchronous and re-entrant
// (SURVEY 8b "Threading"); each host thread gets its own stream + arenas.
struct Staging {
    cudaStream_t stream = nullptr;
    uint8_t *dev = nullptr;
    size_t dev_cap = 0, dev_used = 0;
    uint8_t *host = nullptr;  // pinned
    size_t host_cap = 0, host_used = 0;
    int device = -1;

    int begin(size_t dev_bytes, size_t host_bytes);
    void *dalloc(size_t bytes);   // 256-byte aligned bump allocation (valid until next begin)
    void *halloc(size_t bytes);
    ~Staging();
};
Staging &staging();

// A host rectangle (row 0 pointer, signed byte stride) mirrored on the device
// as densely pitched rows.  Handles negative strides.
struct DevRect {
    uint8_t *dptr = nullptr;  // device, row 0
    int64_t dpitch = 0;       // bytes
    uint8_t *hstage = nullptr;
    const uint8_t *hsrc = nullptr;
    int64_t hstride = 0;
    size_t row_bytes = 0;
    int rows = 0;
    static size_t pitch_for(size_t row_bytes) { return (row_bytes + 63) & ~size_t(63); }
    static size_t bytes_for(size_t row_bytes, int rows) { return pitch_for(row_bytes) * (size_t)rows; }
    // allocate from st and upload
    int upload(Staging &st, const void *host_row0, int64_t stride, size_t row_bytes, int rows);
    // queue the copy back to the same host rectangle (call st sync afterwards, then finish())
    int download(Staging &st);
    void finish(void *host_row0);  // scatter staged rows back
};


// Scope of one synchronous host-pointer DSP call: reserves the calling thread's
// staging arenas, mirrors host buffers on the device and copies results back.
struct HostCall {
    Staging &st;
    int err = 0;
    explicit HostCall(size_t bytes) : st(staging()) { err = st.begin(bytes + 4096, bytes + 4096); }
    cudaStream_t stream() const { return st.stream; }
    // host array -> device (returns device pointer or nullptr on error)
    void *up(const void *h, size_t n) {
        if (err) return nullptr;
        void *d = st.dalloc(n), *s = st.halloc(n);
        if (!d || !s) { err = set_error(-12, "staging arena too small"); return nullptr; }
        memcpy(s, h, n);
        cudaError_t e = cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, st.stream);
        if (e != cudaSuccess) { err = cuda_fail(e, "cudaMemcpyAsync(H2D)", __FILE__, __LINE__); return nullptr; }
        return d;
    }
    // device scratch whose content is fetched back with down()
    void *dev(size_t n) {
        if (err) return nullptr;
        void *d = st.dalloc(n);
        if (!d) err = set_error(-12, "staging arena too small");
        return d;
    }
    // queue D2H of n bytes into a pinned stage; valid after sync()
    void *down(const void *d, size_t n) {
        if (err) return nullptr;
        void *s = st.halloc(n);
        if (!s) { err = set_error(-12, "staging arena too small"); return nullptr; }
        cudaError_t e = cudaMemcpyAsync(s, d, n, cudaMemcpyDeviceToHost, st.stream);
        if (e != cudaSuccess) { err = cuda_fail(e, "cudaMemcpyAsync(D2H)", __FILE__, __LINE__); return nullptr; }
        return s;
    }
    int rect_up(DevRect &r, const void *row0, int64_t stride, size_t row_bytes, int rows) {
        if (err) return err;
        return err = r.upload(st, row0, stride, row_bytes, rows);
    }
    int rect_down(DevRect &r) { if (err) return err; return err = r.download(st); }
    int sync() {
        if (err) return err;
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(st.stream);
        if (e != cudaSuccess) err = cuda_fail(e, "kernel/stream sync", __FILE__, __LINE__);
        return err;
    }
};

}  // namespace rb200
