// Test-only host build of the product's 1-D transform header (the same
// __host__ __device__ code the CUDA kernels inline).  Lets the CPU test suite
// pin the arithmetic against the reference without a GPU.  Never loaded by
// the rav1d_b200 package.
#include "../../rav1d_b200/csrc/itx1d.cuh"
using namespace rb200;
template <int N, int K> static void go(int *x, int lo, int hi) { itx1d<N, K>(x, lo, hi); }
extern "C" int hostsim_itx1d(int n, int kind, int *x, int lo, int hi) {
#define C(N, K) if (n == N && kind == K) { go<N, K>(x, lo, hi); return 0; }
    C(4, 0) C(8, 0) C(16, 0) C(32, 0) C(64, 0)
    C(4, 1) C(8, 1) C(16, 1) C(4, 2) C(8, 2) C(16, 2)
    C(4, 3) C(8, 3) C(16, 3) C(32, 3) C(4, 4)
    return -1;
}
