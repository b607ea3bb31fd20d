// How fast can kernels read pinned host memory over PCIe?  (a) plain 16-byte loads, grid-stride;
// (b) TMA bulk copies (cp.async.bulk) global(host) -> shared -> global(device).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o zerocopy_bw zerocopy_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void copy_ld(const uint4 *__restrict__ src, uint4 *__restrict__ dst, size_t n, int unroll_dummy) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        const uint4 a = __ldcs(src + i), b = __ldcs(src + i + stride), c = __ldcs(src + i + 2 * stride), d = __ldcs(src + i + 3 * stride);
        dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
    }
    for (; i < n; i += stride) dst[i] = __ldcs(src + i);
}

// one CTA streams chunks of CH bytes: TMA load into smem (mbarrier), TMA store to device memory
template <int CH, int NBUF>
__global__ void __launch_bounds__(32) copy_tma(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, size_t bytes) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) unsigned long long bar[NBUF];
    const size_t n_chunks = bytes / CH;
    if (threadIdx.x == 0) {
        for (int b = 0; b < NBUF; b++) {
            unsigned a = (unsigned)__cvta_generic_to_shared(&bar[b]);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        unsigned phase[NBUF] = {};
        size_t issued = 0, c = blockIdx.x;
        // prologue
        for (int b = 0; b < NBUF && c < n_chunks; b++, c += gridDim.x, issued++) {
            unsigned a = (unsigned)__cvta_generic_to_shared(&bar[b]), s = (unsigned)__cvta_generic_to_shared(sm + b * CH);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(CH) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(s), "l"(src + c * CH), "r"(CH), "r"(a) : "memory");
        }
        size_t done = 0, cc = blockIdx.x;
        while (done < issued) {
            const int b = (int)(done % NBUF);
            unsigned a = (unsigned)__cvta_generic_to_shared(&bar[b]), s = (unsigned)__cvta_generic_to_shared(sm + b * CH);
            unsigned ok = 0;
            while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                     : "=r"(ok) : "r"(a), "r"(phase[b]) : "memory");
            phase[b] ^= 1;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + cc * CH), "r"(s), "r"(CH) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // smem buffer free again
            done++; cc += gridDim.x;
            if (c < n_chunks) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(CH) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(s), "l"(src + c * CH), "r"(CH), "r"(a) : "memory");
                c += gridDim.x; issued++;
            }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

int main() {
    const size_t bytes = 256u << 20;
    uint8_t *h, *d;
    CK(cudaMallocHost(&h, bytes));
    CK(cudaMalloc(&d, bytes));
    for (size_t i = 0; i < bytes; i += 4096) h[i] = (uint8_t)i;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms;
    for (int blocks : {148, 148 * 4, 148 * 16}) {
        for (int rep = 0; rep < 2; rep++) {
            CK(cudaEventRecord(e0));
            copy_ld<<<blocks, 256>>>((const uint4 *)h, (uint4 *)d, bytes / 16, 0);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        }
        printf("ld.128 x4   grid %5d x 256: %.1f GB/s\n", blocks, bytes / ms / 1e6);
    }
    CK(cudaEventRecord(e0));
    CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice));
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("cudaMemcpyAsync H2D: %.1f GB/s\n", bytes / ms / 1e6);
    {
        constexpr int CH = 8192, NB = 4;
        CK(cudaFuncSetAttribute(copy_tma<CH, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH * NB));
        for (int blocks : {148, 148 * 2, 148 * 4}) {
            for (int rep = 0; rep < 2; rep++) {
                CK(cudaEventRecord(e0));
                copy_tma<CH, NB><<<blocks, 32, CH * NB>>>(h, d, bytes);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
            }
            printf("TMA bulk %d B x %d bufs, grid %4d: %.1f GB/s\n", CH, NB, blocks, bytes / ms / 1e6);
        }
    }
    {
        constexpr int CH = 512, NB = 8;
        for (int blocks : {148 * 4, 148 * 16}) {
            for (int rep = 0; rep < 2; rep++) {
                CK(cudaEventRecord(e0));
                copy_tma<CH, NB><<<blocks, 32, CH * NB>>>(h, d, bytes);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
            }
            printf("TMA bulk %d B x %d bufs, grid %4d: %.1f GB/s\n", CH, NB, blocks, bytes / ms / 1e6);
        }
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
