// 1-D inverse transforms of the AV1 reconstruction path, written for registers.
//
// Every transform works on a small array that the compiler keeps entirely in
// registers once the (compile-time bounded) loops are unrolled: one GPU thread
// owns one row or one column of a transform block.  The flow graphs are
// expressed generically (bit-reversal order input rotations, alternating
// Hadamard stages, inner-pair rotations) instead of one hand-expanded listing
// per size; constants come from one cos(k*pi/128) table.
//
// Arithmetic contract (bit-exact with the reference, whose 1-D kernels are
//   src/itx_1d.rs:6-1140 == src/itx_1d.c:66-1034):
//   * every rotation is  (a*ca + b*cb + 2048) >> 12  evaluated exactly
//     (no 32-bit wrap for |a|,|b| < 2^20), arithmetic shift;
//   * every add/sub butterfly output is clipped to [lo, hi]      (itx_1d.c:38);
//   * pi/4 rotations are ((a +- b) * 181 + 128) >> 8;
//   * ADST-4 and identity transforms do not clip   (itx_1d.c:783-802,980-1018);
//   * 64-point DCT reads only its first 32 inputs  (itx_1d.c:435-483).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define RB_HD __host__ __device__ __forceinline__
#else
#define RB_HD inline __attribute__((always_inline))
#endif

namespace rb200 {

// round(4096 * cos(k*pi/128)), k = 0..64  (AV1 spec Cos128 lookup).
RB_HD constexpr int cos128(int k) {
    // a switch (not a table) so that the value folds to an immediate once loops are unrolled
    switch (k) {
    case 0: return 4096; case 1: return 4095; case 2: return 4091; case 3: return 4085; case 4: return 4076; case 5: return 4065;
    case 6: return 4052; case 7: return 4036; case 8: return 4017; case 9: return 3996; case 10: return 3973; case 11: return 3948;
    case 12: return 3920; case 13: return 3889; case 14: return 3857; case 15: return 3822; case 16: return 3784; case 17: return 3745;
    case 18: return 3703; case 19: return 3659; case 20: return 3612; case 21: return 3564; case 22: return 3513; case 23: return 3461;
    case 24: return 3406; case 25: return 3349; case 26: return 3290; case 27: return 3229; case 28: return 3166; case 29: return 3102;
    case 30: return 3035; case 31: return 2967; case 32: return 2896; case 33: return 2824; case 34: return 2751; case 35: return 2675;
    case 36: return 2598; case 37: return 2520; case 38: return 2440; case 39: return 2359; case 40: return 2276; case 41: return 2191;
    case 42: return 2106; case 43: return 2019; case 44: return 1931; case 45: return 1842; case 46: return 1751; case 47: return 1660;
    case 48: return 1567; case 49: return 1474; case 50: return 1380; case 51: return 1285; case 52: return 1189; case 53: return 1092;
    case 54: return 995; case 55: return 897; case 56: return 799; case 57: return 700; case 58: return 601; case 59: return 501;
    case 60: return 401; case 61: return 301; case 62: return 201; case 63: return 101; case 64: return 0;
    default: return 0;
    }
}
RB_HD constexpr int sin128(int k) { return cos128(64 - k); }

RB_HD constexpr int brev(int bits, int v) {
    int r = 0;
    for (int i = 0; i < bits; i++) r |= ((v >> i) & 1) << (bits - 1 - i);
    return r;
}
RB_HD constexpr int ilog2c(int v) { return v <= 1 ? 0 : 1 + ilog2c(v >> 1); }

RB_HD int clip3(int v, int lo, int hi) { const int m = v > lo ? v : lo; return m < hi ? m : hi; }   // lo <= hi: max, then min

// exact (a*ca + b*cb + 2048) >> 12 without 32-bit overflow for 20-bit inputs:
// a coefficient above 2048 in magnitude is folded by +-4096 and its operand is
// added back after the shift (x*4096 >> 12 == x exactly); two even
// coefficients are halved instead and the shift becomes 11.
RB_HD constexpr int fold_c(int c) { return c > 2048 ? c - 4096 : (c < -2048 ? c + 4096 : c); }
RB_HD constexpr int fold_s(int c) { return c > 2048 ? 1 : (c < -2048 ? -1 : 0); }

RB_HD int rot(int a, int ca, int b, int cb) {
    if (((ca | cb) & 1) == 0 && (ca > 2048 || ca < -2048 || cb > 2048 || cb < -2048))
        return (a * (ca / 2) + b * (cb / 2) + 1024) >> 11;
    return ((a * fold_c(ca) + b * fold_c(cb) + 2048) >> 12) + fold_s(ca) * a + fold_s(cb) * b;
}
RB_HD int rot1(int a, int ca) { return (a * ca + 2048) >> 12; }  // |ca| <= 4096, single term
RB_HD int mul181(int v) { return (v * 181 + 128) >> 8; }

// ------------------------------------------------------------------ DCT
// x[0..N) natural order in, natural order out.  HALF: inputs N/2.. are zero
// (only used below a 64-point transform) and are never read.
template <int N, bool HALF>
struct Dct {
    static RB_HD void run(int *x, const int lo, const int hi) {
        constexpr int M = N / 2;
        constexpr int LN = ilog2c(N);
        int e[M];
#pragma unroll
        for (int i = 0; i < M; i++) e[i] = (HALF && 2 * i >= M) ? 0 : x[2 * i];
        Dct<M, HALF>::run(e, lo, hi);

        int t[M];  // the odd half, t[i] == "t(M+i)" of the textbook numbering
        // A: input rotations, pairs (i, M-1-i), odd inputs visited in bit-reversed order
#pragma unroll
        for (int i = 0; i < M / 2; i++) {
            const int p = brev(LN, i) + 1, q = N - p;
            const int k = p * (128 / (2 * N)) ;  // angle p*pi/(2N) in units of pi/128
            const int s = sin128(k), c = cos128(k);
            if (HALF) {
                if (p < M) { t[i] = rot1(x[p], s);  t[M - 1 - i] = rot1(x[p], c); }
                else       { t[i] = rot1(x[q], -c); t[M - 1 - i] = rot1(x[q], s); }
            } else {
                t[i]         = rot(x[p], s, x[q], -c);
                t[M - 1 - i] = rot(x[p], c, x[q], s);
            }
        }
#pragma unroll
        for (int g = 2; g < M; g *= 2) {
            // Hadamard over groups of g: even groups (a,b)<-(a+b,a-b), odd groups (a,b)<-(b-a,b+a)
#pragma unroll
            for (int k = 0; k < M / g; k++) {
#pragma unroll
                for (int j = 0; j < g / 2; j++) {
                    const int a = k * g + j, b = k * g + g - 1 - j;
                    const int s = clip3(t[a] + t[b], lo, hi);
                    const int d = (k & 1) ? clip3(t[b] - t[a], lo, hi) : clip3(t[a] - t[b], lo, hi);
                    if (k & 1) { t[a] = d; t[b] = s; } else { t[a] = s; t[b] = d; }
                }
            }
            // inner-pair rotations with period G = 2g, pairs (i, M-1-i)
            const int G = 2 * g;
#pragma unroll
            for (int i = 0; i < M / 2; i++) {
                const int r = i % G;
                if (r < G / 4 || r >= 3 * G / 4) continue;
                const int a = i, b = M - 1 - i;
                const int ta = t[a], tb = t[b];
                if (G == M) {
                    t[a] = mul181(tb - ta);
                    t[b] = mul181(tb + ta);
                } else {
                    const int NS = N / G;  // constants of the (N/G)-point stage A
                    const int p = brev(ilog2c(NS), i / G) + 1;
                    const int k = p * (128 / (2 * NS));
                    const int s = sin128(k), c = cos128(k);
                    if (r < G / 2) { t[a] = rot(tb, s, ta, -c);  t[b] = rot(tb, c, ta, s); }
                    else           { t[a] = rot(tb, -c, ta, -s); t[b] = rot(tb, s, ta, -c); }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < M; i++) {
            x[i]         = clip3(e[i] + t[M - 1 - i], lo, hi);
            x[N - 1 - i] = clip3(e[i] - t[M - 1 - i], lo, hi);
        }
    }
};

template <bool HALF>
struct Dct<2, HALF> {
    static RB_HD void run(int *x, const int, const int) {
        if (HALF) { x[0] = x[1] = mul181(x[0]); }
        else { const int a = x[0], b = x[1]; x[0] = mul181(a + b); x[1] = mul181(a - b); }
    }
};

// ------------------------------------------------------------------ ADST
RB_HD int dot4(int a, int ca, int b, int cb, int c, int cc, int d, int cd) {
    return ((a * fold_c(ca) + b * fold_c(cb) + c * fold_c(cc) + d * fold_c(cd) + 2048) >> 12) +
           fold_s(ca) * a + fold_s(cb) * b + fold_s(cc) * c + fold_s(cd) * d;
}

// o[] may alias nothing; caller copies (flip handled by caller).  itx_1d.c:783-802
RB_HD void adst4(const int *in, int *o) {
    const int a = in[0], b = in[1], c = in[2], d = in[3];
    o[0] = dot4(a, 1321, b, 3344, c, 3803, d, 2482);
    o[1] = dot4(a, 2482, b, 3344, c, -1321, d, -3803);
    o[2] = (209 * (a - c + d) + 128) >> 8;
    o[3] = dot4(a, 3803, b, -3344, c, 2482, d, -1321);
}

// N = 8 or 16.  itx_1d.c:804-856 (8), :858-974 (16)
template <int N>
RB_HD void adstN(const int *in, int *o, const int lo, const int hi) {
    int t[N];
    // stage 1: (t[2i], t[2i+1]) from (in[N-1-2i], in[2i]), angle (4i+1)*pi/(4N)
#pragma unroll
    for (int i = 0; i < N / 2; i++) {
        const int k = (4 * i + 1) * (128 / (4 * N));
        const int s = sin128(k), c = cos128(k);
        const int u = in[N - 1 - 2 * i], v = in[2 * i];
        t[2 * i]     = rot(u, c, v, s);
        t[2 * i + 1] = rot(u, s, v, -c);
    }
    // add/sub at distance h, then rotate pairs inside the upper half of every
    // 2h block (first half of the pairs "forward", second half mirrored)
#pragma unroll
    for (int h = N / 2; h >= 4; h /= 2) {
#pragma unroll
        for (int base = 0; base < N; base += 2 * h) {
#pragma unroll
            for (int j = 0; j < h; j++) {
                const int a = t[base + j], b = t[base + h + j];
                t[base + j]     = clip3(a + b, lo, hi);
                t[base + h + j] = clip3(a - b, lo, hi);
            }
            const int half = h / 4;  // pairs per direction
#pragma unroll
            for (int m = 0; m < h / 2; m++) {
                const int ia = base + h + 2 * m, ib = ia + 1;
                const int kk = (4 * (m % half) + 1) * (64 / h);  // angle in units of pi/128
                const int c = cos128(kk), s = sin128(kk);
                const int a = t[ia], b = t[ib];
                if (m < half) { t[ia] = rot(a, c, b, s);  t[ib] = rot(a, s, b, -c); }
                else          { t[ia] = rot(b, c, a, -s); t[ib] = rot(b, s, a, c); }
            }
        }
    }
    // h == 2 stage: add/sub at distance 2 producing half of the outputs, pi/4 on the rest
    // output order follows the transform's bit-reversed sign pattern.
    if (N == 8) {
        const int ord[4] = {0, 7, 1, 6};   // sums of (0,2),(1,3),(4,6),(5,7)
        const int sgn[4] = {1, -1, -1, 1};
        const int mo[4]  = {3, 4, 2, 5};   // pi/4 outputs: -(a+b), (a-b), (c+d), -(c-d)
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int b0 = 4 * q;
            const int s0 = clip3(t[b0] + t[b0 + 2], lo, hi), s1 = clip3(t[b0 + 1] + t[b0 + 3], lo, hi);
            const int d0 = clip3(t[b0] - t[b0 + 2], lo, hi), d1 = clip3(t[b0 + 1] - t[b0 + 3], lo, hi);
            o[ord[2 * q]]     = sgn[2 * q] * s0;
            o[ord[2 * q + 1]] = sgn[2 * q + 1] * s1;
            if (q == 0) { o[mo[0]] = -mul181(d0 + d1); o[mo[1]] = mul181(d0 - d1); }
            else        { o[mo[2]] = mul181(d0 + d1);  o[mo[3]] = -mul181(d0 - d1); }
        }
    } else {
        const int ord[8] = {0, 15, 3, 12, 1, 14, 2, 13};
        const int sgn[8] = {1, -1, -1, 1, -1, 1, 1, -1};
        // pi/4 outputs per quad: (idx of +(d0+d1) or -(..), idx of (d0-d1)), with signs
        const int pa[4] = {7, 4, 6, 5},  pas[4] = {-1, 1, 1, -1};
        const int pb[4] = {8, 11, 9, 10}, pbs[4] = {1, -1, -1, 1};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int b0 = 4 * q;
            const int s0 = clip3(t[b0] + t[b0 + 2], lo, hi), s1 = clip3(t[b0 + 1] + t[b0 + 3], lo, hi);
            const int d0 = clip3(t[b0] - t[b0 + 2], lo, hi), d1 = clip3(t[b0 + 1] - t[b0 + 3], lo, hi);
            o[ord[2 * q]]     = sgn[2 * q] * s0;
            o[ord[2 * q + 1]] = sgn[2 * q + 1] * s1;
            o[pa[q]] = pas[q] * mul181(d0 + d1);
            o[pb[q]] = pbs[q] * mul181(d0 - d1);
        }
    }
}

// ------------------------------------------------------------------ identity / WHT
RB_HD int identity4(int v)  { return v + ((v * 1697 + 2048) >> 12); }
RB_HD int identity8(int v)  { return v * 2; }
RB_HD int identity16(int v) { return 2 * v + ((v * 1697 + 1024) >> 11); }
RB_HD int identity32(int v) { return v * 4; }

RB_HD void wht4(int *x) {  // itx_1d.c:1020-1034
    const int a = x[0] + x[1], c = x[2] - x[3];
    const int m = (a - c) >> 1;
    const int b = m - x[3], d = m - x[1];
    x[0] = a - b; x[1] = b; x[2] = d; x[3] = c + d;
}

// ------------------------------------------------------------------ dispatch
enum Tx1d { T1_DCT = 0, T1_ADST = 1, T1_FLIPADST = 2, T1_IDENTITY = 3, T1_WHT = 4 };

// In-place N-point transform of kind K on x[0..N).  For N == 64 only x[0..32) is read.
template <int N, int K>
RB_HD void itx1d(int *x, const int lo, const int hi) {
    if constexpr (K == T1_DCT) {
        Dct<N, N == 64>::run(x, lo, hi);
    } else if constexpr (K == T1_ADST || K == T1_FLIPADST) {
        static_assert(N <= 16, "ADST exists for 4, 8 and 16 points only");
        int o[N];
        if constexpr (N == 4) adst4(x, o);
        else adstN<N>(x, o, lo, hi);
#pragma unroll
        for (int i = 0; i < N; i++) x[i] = o[K == T1_FLIPADST ? N - 1 - i : i];
    } else if constexpr (K == T1_IDENTITY) {
#pragma unroll
        for (int i = 0; i < N; i++)
            x[i] = N == 4 ? identity4(x[i]) : N == 8 ? identity8(x[i]) : N == 16 ? identity16(x[i]) : identity32(x[i]);
    } else {
        wht4(x);
    }
}

}  // namespace rb200
