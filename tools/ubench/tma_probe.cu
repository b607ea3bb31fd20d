// Probe of the TMA helpers in rav1d_b200/csrc/tma.cuh: loads boxes of a u16 plane at assorted coordinates (also outside the
// tensor) and compares with a host gather (zero fill).  nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../rav1d_b200/csrc/tma.cuh"
using namespace rb200;
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void probe(const __grid_constant__ CUtensorMap map, int x, int y, int box_el, uint16_t *out, int variant) {
    extern __shared__ uint8_t dyn[];
    __shared__ __align__(8) uint64_t bar;
    uint8_t *sm = dyn + ((128u - (smem_addr(dyn) & 127u)) & 127u);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar, (unsigned)box_el * 2);
        tma_load_2d(sm, &map, x, y, &bar);
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < box_el; i += blockDim.x) out[i] = ((uint16_t *)sm)[i];
}

int main(int argc, char **argv) {
    const int W = 208, H = 128, stride = 256;   // elements
    std::vector<uint16_t> h(stride * H);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint16_t)(i * 7 + 3);
    uint16_t *d, *dout;
    CK(cudaMalloc(&d, h.size() * 2)); CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dout, 256 * 256 * 2));
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncodeTiledFn fn = (EncodeTiledFn)p;
    // one case per process (a faulting launch poisons the context): tma_probe box_w box_h x y
    if (argc < 5) { printf("usage: tma_probe box_w box_h x y\n"); return 1; }
    const int boxes[][2] = {{atoi(argv[1]), atoi(argv[2])}};
    const int coords[][2] = {{atoi(argv[3]), atoi(argv[4])}};
    int bad = 0;
    for (auto &b : boxes) {
        CUtensorMap map;
        const cuuint64_t dims[2] = {W, H}, strides[1] = {stride * 2};
        const cuuint32_t box[2] = {(cuuint32_t)b[0], (cuuint32_t)b[1]}, es[2] = {1, 1};
        CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) { printf("encode %dx%d failed %d\n", b[0], b[1], (int)r); bad++; continue; }
        for (auto &c : coords) {
            const int n = b[0] * b[1];
            CK(cudaMemset(dout, 0xff, n * 2));
            probe<<<1, 128, n * 2 + 128>>>(map, c[0], c[1], n, dout, 0);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("box %dx%d at (%d,%d): %s\n", b[0], b[1], c[0], c[1], cudaGetErrorString(e)); return 2; }
            std::vector<uint16_t> o(n);
            CK(cudaMemcpy(o.data(), dout, n * 2, cudaMemcpyDeviceToHost));
            int miss = 0;
            for (int yy = 0; yy < b[1]; yy++)
                for (int xx = 0; xx < b[0]; xx++) {
                    const int gx = c[0] + xx, gy = c[1] + yy;
                    const uint16_t e2 = (gx < 0 || gx >= W || gy < 0 || gy >= H) ? 0 : h[gy * stride + gx];
                    miss += o[yy * b[0] + xx] != e2;
                }
            printf("box %dx%d at (%d,%d): %d mismatches\n", b[0], b[1], c[0], c[1], miss);
            bad += miss != 0;
        }
    }
    printf(bad ? "FAILED\n" : "all ok\n");
    return bad != 0;
}
