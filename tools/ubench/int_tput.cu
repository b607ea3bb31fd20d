// Micro-benchmark: issue throughput of the integer instructions the DSP kernels lean on
// (sm_100a).  Prints thread-ops per clock per SM for dependent-chain-free streams.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_tput int_tput.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N_ITERS 4096
#define DECL(name, body)                                                              \
    __global__ void k_##name(unsigned *out, unsigned seed) {                          \
        unsigned a0 = threadIdx.x + seed, a1 = a0 * 3 + 1, a2 = a0 * 5 + 2, a3 = a0 * 7 + 3; \
        unsigned a4 = a0 ^ 0x1234, a5 = a1 ^ 0x4321, a6 = a2 + 99, a7 = a3 + 77;       \
        unsigned b = seed * 0x9e3779b9u | 1, c = seed + 12345;                        \
        for (int i = 0; i < N_ITERS; i++) {                                           \
            body(a0) body(a1) body(a2) body(a3) body(a4) body(a5) body(a6) body(a7)   \
        }                                                                             \
        out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7; \
    }
#define B_IMAD(x) x = x * b + c;
#define B_IADD(x) x = x + b;
#define B_LOP3(x) x = (x & b) ^ c;
#define B_SHF(x) x = (x >> 3) | (x << 29);
#define B_VIADD2(x) x = __vadd2(x, b);
#define B_VMAX2(x) x = __vmaxs2(x, b);
#define B_VMINU2(x) x = __vminu2(x, b);
#define B_VIADDMAX(x) x = __viaddmax_s16x2(x, b, c);
#define B_VIMAX3(x) x = __vimax3_s16x2(x, b, c);
#define B_DP2A(x) x = __dp2a_lo((int)x, (int)b, (int)c);
#define B_DP4A(x) x = __dp4a((int)x, (int)b, (int)c);
#define B_IMNMX(x) x = max((int)x, (int)b);
#define B_IABS(x) x = abs((int)x) + b;
#define B_PRMT(x) x = __byte_perm(x, b, 0x5432);
#define B_MIX(x) x = __vmaxs2(x * b + c, b);
DECL(imad, B_IMAD) DECL(iadd, B_IADD) DECL(lop3, B_LOP3) DECL(shf, B_SHF) DECL(viadd2, B_VIADD2)
DECL(vmax2, B_VMAX2) DECL(vminu2, B_VMINU2) DECL(viaddmax, B_VIADDMAX) DECL(vimax3, B_VIMAX3)
DECL(dp2a, B_DP2A) DECL(dp4a, B_DP4A) DECL(imnmx, B_IMNMX) DECL(iabs_add, B_IABS) DECL(prmt, B_PRMT) DECL(imad_vmax, B_MIX)
template <typename K> void run(const char *name, K k, int ops_per_body) {
    unsigned *out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 8, threads = 256;
    k<<<blocks, threads>>>(out, 1); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<<<blocks, threads>>>(out, 2); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double ops = (double)blocks * threads * N_ITERS * 8 * ops_per_body;
    printf("%-10s %8.3f ms  %7.1f Gthread-op/s  %6.1f thread-ops/clk/SM (at %d MHz nominal)\n", name, ms, ops / ms / 1e6,
           ops / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000);
    cudaFree(out);
}
int main() {
    run("imad", k_imad, 1); run("iadd", k_iadd, 1); run("lop3", k_lop3, 1); run("shf", k_shf, 1);
    run("viadd.16x2", k_viadd2, 1); run("vimnmx.s16x2", k_vmax2, 1); run("vimnmx.u16x2", k_vminu2, 1);
    run("viaddmnmx", k_viaddmax, 1); run("vimnmx3", k_vimax3, 1); run("idp.2a", k_dp2a, 1); run("idp.4a", k_dp4a, 1);
    run("imnmx", k_imnmx, 1); run("iabs+iadd", k_iabs_add, 2); run("prmt", k_prmt, 1); run("imad+vmax2", k_imad_vmax, 2);
    return 0;
}
