// Core plumbing of the rav1d_b200 library: error reporting, per-thread staging
// arenas for the host-pointer entry points, and small device-memory helpers so
// that a non-CUDA host (Rust FFI, Python ctypes) can drive the batch API.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tma.cuh"

namespace rb200 {

static thread_local char g_err[512];
static rb200_error_cb g_cb = nullptr;
static void *g_cb_cookie = nullptr;
static thread_local int g_last_code = 0;

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    g_last_code = code;
    return code;
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    const char *base = strrchr(file, '/');
    return set_error(e == cudaErrorMemoryAllocation ? -12 : -5, "CUDA error %d (%s) at %s:%d: %s", (int)e,
                     cudaGetErrorString(e), base ? base + 1 : file, line, what);
}

// ------------------------------------------------------------ Staging
Staging::~Staging() {
    // Process teardown order vs. the CUDA runtime is undefined; leak on exit.
}

int Staging::begin(size_t dev_bytes, size_t host_bytes) {
    if (!stream) {
        RB_CUDA(cudaGetDevice(&device));
        RB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    }
    if (dev_bytes > dev_cap) {
        if (dev) RB_CUDA(cudaFree(dev));
        dev = nullptr;
        size_t cap = dev_bytes < (1u << 20) ? (1u << 20) : dev_bytes * 2;
        RB_CUDA(cudaMalloc((void **)&dev, cap));
        dev_cap = cap;
    }
    if (host_bytes > host_cap) {
        if (host) RB_CUDA(cudaFreeHost(host));
        host = nullptr;
        size_t cap = host_bytes < (1u << 20) ? (1u << 20) : host_bytes * 2;
        RB_CUDA(cudaMallocHost((void **)&host, cap));
        host_cap = cap;
    }
    dev_used = host_used = 0;
    return 0;
}

void *Staging::dalloc(size_t bytes) {
    size_t off = (dev_used + 255) & ~size_t(255);
    if (off + bytes > dev_cap) return nullptr;
    dev_used = off + bytes;
    return dev + off;
}
void *Staging::halloc(size_t bytes) {
    size_t off = (host_used + 63) & ~size_t(63);
    if (off + bytes > host_cap) return nullptr;
    host_used = off + bytes;
    return host + off;
}

Staging &staging() {
    static thread_local Staging s;
    return s;
}

// ------------------------------------------------------------ DevRect
int DevRect::upload(Staging &st, const void *host_row0, int64_t stride, size_t rb, int nrows) {
    row_bytes = rb; rows = nrows; hstride = stride; hsrc = (const uint8_t *)host_row0;
    dpitch = (int64_t)pitch_for(rb);
    const size_t total = (size_t)dpitch * rows;
    dptr = (uint8_t *)st.dalloc(total);
    hstage = (uint8_t *)st.halloc(total);
    if (!dptr || !hstage) return set_error(-12, "staging arena too small (%zu bytes)", total);
    for (int y = 0; y < rows; y++) memcpy(hstage + (size_t)y * dpitch, hsrc + (int64_t)y * stride, rb);
    RB_CUDA(cudaMemcpyAsync(dptr, hstage, total, cudaMemcpyHostToDevice, st.stream));
    return 0;
}
int DevRect::download(Staging &st) {
    RB_CUDA(cudaMemcpyAsync(hstage, dptr, (size_t)dpitch * rows, cudaMemcpyDeviceToHost, st.stream));
    return 0;
}
void DevRect::finish(void *host_row0) {
    uint8_t *d = (uint8_t *)host_row0;
    for (int y = 0; y < rows; y++) memcpy(d + (int64_t)y * hstride, hstage + (size_t)y * dpitch, row_bytes);
}

// ---- cross-GPU flags: a stream of one GPU tells a stream of another that it is done with something (halo rows pulled),
// without the host and without a collective.  The flag lives in the waiting GPU's memory; the signalling GPU writes it
// through its peer mapping (NVLink).
__global__ void flag_signal_kernel(volatile uint32_t *flag, uint32_t value) {
    __threadfence_system();      // (the copies queued before this launch are complete: stream order)
    *flag = value;
    __threadfence_system();
}
__global__ void flag_wait_kernel(const volatile uint32_t *flag, uint32_t value) {
    while ((int32_t)(*flag - value) < 0) __nanosleep(200);
    __threadfence_system();
}

// ---- tensor maps (tma.cuh).  cuTensorMapEncodeTiled is a driver entry point; it is looked up through the runtime so
// that the library needs no link-time libcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int tma_encode_plane(CUtensorMap *out, const void *base, int elem_bytes, int width, int height, int64_t stride_bytes,
                     int box_w, int box_h) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        RB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) return set_error(-38, "cuTensorMapEncodeTiled is not available from this driver");
        fn = (EncodeTiledFn)p;
    }
    if (!out || !base || (elem_bytes != 1 && elem_bytes != 2) || width < 1 || height < 1 || box_w < 1 || box_h < 1 || box_w > 256 ||
        box_h > 256 || ((box_w * elem_bytes) & 15) || (stride_bytes & 15) || ((uintptr_t)base & 15))
        return set_error(-22, "tma_encode_plane: bad geometry (%d x %d, stride %lld, box %d x %d)", width, height, (long long)stride_bytes, box_w, box_h);
    const cuuint64_t dims[2] = {(cuuint64_t)width, (cuuint64_t)height};
    const cuuint64_t strides[1] = {(cuuint64_t)stride_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h}, estr[2] = {1, 1};
    const CUresult r = fn(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *)base, dims, strides,
                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(-22, "cuTensorMapEncodeTiled failed (%d) for %d x %d, stride %lld, box %d x %d", (int)r, width, height,
                                            (long long)stride_bytes, box_w, box_h);
    return 0;
}

}  // namespace rb200

using namespace rb200;

extern "C" int rb200_abi_version(void) { return RB200_ABI_VERSION; }

extern "C" int rb200_init(int device) {
    int n = 0;
    RB_CUDA(cudaGetDeviceCount(&n));
    if (n <= 0) return set_error(-19, "no CUDA device visible: rav1d_b200 has no CPU fallback");
    if (device < 0) RB_CUDA(cudaGetDevice(&device));
    if (device >= n) return set_error(-22, "device %d out of range (%d visible)", device, n);
    RB_CUDA(cudaSetDevice(device));
    cudaDeviceProp p;
    RB_CUDA(cudaGetDeviceProperties(&p, device));
    if (p.major != 10) return set_error(-19, "device %d is sm_%d%d; this library is built for sm_100a only", device, p.major, p.minor);
    RB_CUDA(cudaFree(0));
    return 0;
}

extern "C" const char *rb200_last_error(void) { return g_err; }

extern "C" void rb200_set_error_callback(rb200_error_cb cb, void *cookie) { g_cb = cb; g_cb_cookie = cookie; }

extern "C" void rb200_report_fatal(const char *where) {
    if (g_cb) { g_cb(g_cb_cookie, g_last_code, g_err); return; }
    fprintf(stderr, "rav1d_b200: fatal GPU error in %s: %s\n", where, g_err);
    abort();
}

extern "C" int rb200_malloc(void **dptr, size_t bytes) { RB_CUDA(cudaMalloc(dptr, bytes)); return 0; }
extern "C" int rb200_free(void *dptr) { RB_CUDA(cudaFree(dptr)); return 0; }
extern "C" int rb200_malloc_host(void **hptr, size_t bytes) { RB_CUDA(cudaMallocHost(hptr, bytes)); return 0; }
extern "C" int rb200_free_host(void *hptr) { RB_CUDA(cudaFreeHost(hptr)); return 0; }
extern "C" int rb200_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream) {
    RB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream)); return 0;
}
extern "C" int rb200_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream) {
    RB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream)); return 0;
}
extern "C" int rb200_memset(void *dst, int value, size_t bytes, void *stream) {
    RB_CUDA(cudaMemsetAsync(dst, value, bytes, (cudaStream_t)stream)); return 0;
}
extern "C" int rb200_stream_sync(void *stream) { RB_CUDA(cudaStreamSynchronize((cudaStream_t)stream)); return 0; }
extern "C" int rb200_stream_create(void **stream) {
    if (!stream) return set_error(-22, "stream_create: null argument");
    cudaStream_t s;
    RB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = (void *)s;
    return 0;
}
extern "C" int rb200_stream_destroy(void *stream) { RB_CUDA(cudaStreamDestroy((cudaStream_t)stream)); return 0; }

// ---- CUDA IPC: map another process's device allocation (halo exchange between ranks, one process per GPU)
static_assert(sizeof(cudaIpcMemHandle_t) == RB200_IPC_HANDLE_BYTES, "ipc handle size");
extern "C" int rb200_ipc_get_handle(void *dptr, uint8_t handle[RB200_IPC_HANDLE_BYTES]) {
    cudaIpcMemHandle_t h;
    RB_CUDA(cudaIpcGetMemHandle(&h, dptr));
    memcpy(handle, &h, sizeof(h));
    return 0;
}
extern "C" int rb200_ipc_open_handle(const uint8_t handle[RB200_IPC_HANDLE_BYTES], void **dptr) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    RB_CUDA(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
extern "C" int rb200_ipc_close_handle(void *dptr) { RB_CUDA(cudaIpcCloseMemHandle(dptr)); return 0; }
extern "C" int rb200_flag_signal(void *stream, uint32_t *flag, uint32_t value) {
    if (!flag) return set_error(-22, "flag_signal: null flag");
    flag_signal_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(flag, value);
    RB_LAUNCH_CHECK();
    return 0;
}
extern "C" int rb200_flag_wait(void *stream, const uint32_t *flag, uint32_t value) {
    if (!flag) return set_error(-22, "flag_wait: null flag");
    flag_wait_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(flag, value);
    RB_LAUNCH_CHECK();
    return 0;
}

extern "C" int rb200_enable_peer_access(int peer_device) {
    int can = 0, dev = 0;
    RB_CUDA(cudaGetDevice(&dev));
    RB_CUDA(cudaDeviceCanAccessPeer(&can, dev, peer_device));
    if (!can) return set_error(-19, "device %d cannot access peer %d", dev, peer_device);
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return 0; }
    RB_CUDA(e);
    return 0;
}
