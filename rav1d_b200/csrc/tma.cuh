// Tensor Memory Accelerator plumbing for the picture-tile kernels (sm_100a): 2-D tensor maps over picture planes,
// mbarrier completion, bulk tensor loads issued by one elected thread.
//
// A picture plane is a 2-D tensor (width x height elements, row pitch = plane stride).  A kernel names a tile by
// its element coordinates -- also negative ones and tiles that hang over the picture: the copy engine does the address
// arithmetic and fills what lies outside the tensor with zeros -- and the tile lands in shared memory as dense rows of
// the box width.  One rule found the hard way (tools/ubench/tma_probe.cu; the fault is "illegal instruction"): the
// first column of a box must start on a 16-byte boundary of the plane (x a multiple of 8 for 16-bit pixels, of 16 for
// 8-bit ones); rows are free.  The staging loops this replaces spent 4-5 instructions per 4-byte word on address
// arithmetic and bounds tests (ncu: a quarter of the CDEF kernel's instructions, 40 % LSU pipe in the MC kernel).
#pragma once
#include <cuda.h>          // CUtensorMap (types only; the encoder is reached through cudaGetDriverEntryPoint)
#include <cuda_runtime.h>
#include <stdint.h>

namespace rb200 {

// Host: tensor map of a plane of `width` x `height` elements of elem_bytes (1 or 2), rows `stride_bytes` apart
// (multiple of 16, base 16-byte aligned), for boxes of box_w x box_h elements (box_w * elem_bytes a multiple of 16,
// both <= 256).  Returns 0 or a negative errno (rb200_last_error() says why).
int tma_encode_plane(CUtensorMap *out, const void *base, int elem_bytes, int width, int height, int64_t stride_bytes,
                     int box_w, int box_h);

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_addr(bar)), "r"(arrivals));
}
// makes the initialised barrier visible to the async proxy (the copy engine signals it)
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
// orders this thread's earlier generic-proxy accesses to shared memory before later async-proxy (TMA) accesses
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
// one arrival + the number of bytes the copy engine will deliver to this phase
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// box of the map at element coordinates (x, y), x * elem_bytes a multiple of 16 -> dense rows at smem_dst (128-byte
// aligned); completion on `bar`
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, int x, int y, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
            smem_addr(smem_dst)),
        "l"((uint64_t)map), "r"(smem_addr(bar)), "r"(x), "r"(y)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"((uint64_t)map) : "memory");
}
#endif

}  // namespace rb200
