// Shared declarations of the rav1d_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/rav1d_b200.h"

namespace rb200 {

// Bit-depth classes, mirroring the reference's BitDepth trait
// (include/common/bitdepth.rs:277-289 BitDepth8, :355-366 BitDepth16).
struct BD8 {
    using pixel = uint8_t;
    using coef = int16_t;
    static constexpr bool hbd = false;
};
struct BD16 {
    using pixel = uint16_t;
    using coef = int32_t;
    static constexpr bool hbd = true;
};

__host__ __device__ __forceinline__ int imin(int a, int b) { return a < b ? a : b; }
__host__ __device__ __forceinline__ int imax(int a, int b) { return a > b ? a : b; }
__host__ __device__ __forceinline__ int iclip(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
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

}  // namespace rb200
