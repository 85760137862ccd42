// Fast forward transform + quantisation with a proven guard band and an exact fallback.
//
// The reference evaluates every coefficient as a 2*N*N-operation binary64 chain in a fixed order (algo.cpp:309-331) and
// then std::round(e / m) (Block.cpp:152).  Doing that for every coefficient costs 128 DP instructions per pixel at 8x8.
// Instead:
//   1. fast path: separable even/odd factorised DCT in FP32 (FMA allowed), q~ = Y * (C(u)C(v)/Q[u][v]);
//      r = rn(q~) by the 1.5*2^23 magic constant inside one FFMA, d = q~ - r.
//   2. guard: the fast value differs from the exact real quotient by at most delta_uv (bound below).  If
//      |d| < 0.5 - delta_uv the reference's rounding cannot disagree with r (its own deviation from the real value is
//      ~1e-13, far inside delta).  Every true tie (real quotient exactly k + 0.5, where the reference's rounding noise
//      decides) lies inside the band for any delta > 0.
//   3. fallback: coefficients inside the band are recomputed in the reference's exact order and precision
//      (fdct_coef_exact, __dmul_rn/__dadd_rn/__ddiv_rn, host-built tables) -- so the result is exact, not statistical.
//
// Error bound (X = max |x|, u = 2^-24): every 1-D output is an expression of depth <= 6 roundings (constants included)
// over intermediates bounded by 8X (pass 1) and 64X (pass 2); pass-1 errors are amplified by sum|cos| <= 8 in pass 2:
//   |Y_fast - Y_real| <= 8 * 6u * 8X + 6u * 64X = 768 u X ;  k_float = k (1 + eps), |eps| <= u adds |Y| k u <= 64 X k u.
//   delta_uv = 1.25 * (768 + 64) * u * X * k_uv      (X = 128 for pixels - 128, 383 for P-frame residuals - 128)
// Measured worst case over adversarial sign patterns: 3.7e-6 X (tools/ + DESIGN.md), 12x below the bound used.
#pragma once
#include "common.cuh"
#include "transform.cuh"

namespace ie {

struct FastQuant {
    float k[kMaxNN];        // C(u)C(v) / Q[u][v]
    float thr[kMaxNN];      // 0.5 - delta_uv
};

#ifdef __CUDACC__

constexpr float kMagic = 12582912.0f;          // 1.5 * 2^23
constexpr int kMagicBits = 0x4B400000;

// cos(m*pi/16)
#define IE_C1 0.98078528040323043f
#define IE_C2 0.92387953251128674f
#define IE_C3 0.83146961230254524f
#define IE_C4 0.70710678118654752f
#define IE_C5 0.55557023301960218f
#define IE_C6 0.38268343236508977f
#define IE_C7 0.19509032201612825f

// unnormalised 8-point DCT-II: y[k] = sum_n x[n] cos((2n+1) k pi / 16), in place, stride `S` between elements
template <int S>
__device__ __forceinline__ void dct8_inplace(float *v) {
    const float s0 = v[0 * S] + v[7 * S], d0 = v[0 * S] - v[7 * S];
    const float s1 = v[1 * S] + v[6 * S], d1 = v[1 * S] - v[6 * S];
    const float s2 = v[2 * S] + v[5 * S], d2 = v[2 * S] - v[5 * S];
    const float s3 = v[3 * S] + v[4 * S], d3 = v[3 * S] - v[4 * S];
    const float t0 = s0 + s3, t1 = s1 + s2, t2 = s0 - s3, t3 = s1 - s2;
    v[0 * S] = t0 + t1;
    v[4 * S] = (t0 - t1) * IE_C4;
    v[2 * S] = fmaf(t2, IE_C2, t3 * IE_C6);
    v[6 * S] = fmaf(t2, IE_C6, -(t3 * IE_C2));
    v[1 * S] = fmaf(d0, IE_C1, fmaf(d1, IE_C3, fmaf(d2, IE_C5, d3 * IE_C7)));
    v[3 * S] = fmaf(d0, IE_C3, fmaf(d1, -IE_C7, fmaf(d2, -IE_C1, d3 * -IE_C5)));
    v[5 * S] = fmaf(d0, IE_C5, fmaf(d1, -IE_C1, fmaf(d2, IE_C7, d3 * IE_C3)));
    v[7 * S] = fmaf(d0, IE_C7, fmaf(d1, -IE_C5, fmaf(d2, IE_C3, d3 * -IE_C1)));
}

// unnormalised 4-point DCT-II: y[k] = sum_n x[n] cos((2n+1) k pi / 8)
template <int S>
__device__ __forceinline__ void dct4_inplace(float *v) {
    const float s0 = v[0 * S] + v[3 * S], d0 = v[0 * S] - v[3 * S];
    const float s1 = v[1 * S] + v[2 * S], d1 = v[1 * S] - v[2 * S];
    v[0 * S] = s0 + s1;
    v[2 * S] = (s0 - s1) * IE_C4;
    v[1 * S] = fmaf(d0, IE_C2, d1 * IE_C6);
    v[3 * S] = fmaf(d0, IE_C6, -(d1 * IE_C2));
}

// inverse of the above (unnormalised DCT-III): x[i] = sum_u y[u] cos((2i+1) u pi / 16)
template <int S>
__device__ __forceinline__ void idct8_inplace(float *v) {
    const float y0 = v[0 * S], y1 = v[1 * S], y2 = v[2 * S], y3 = v[3 * S], y4 = v[4 * S], y5 = v[5 * S], y6 = v[6 * S], y7 = v[7 * S];
    const float t4 = y4 * IE_C4;
    const float a = y0 + t4, b = y0 - t4;
    const float p = fmaf(y2, IE_C2, y6 * IE_C6), q = fmaf(y2, IE_C6, -(y6 * IE_C2));
    const float e0 = a + p, e3 = a - p, e1 = b + q, e2 = b - q;
    const float o0 = fmaf(y1, IE_C1, fmaf(y3, IE_C3, fmaf(y5, IE_C5, y7 * IE_C7)));
    const float o1 = fmaf(y1, IE_C3, fmaf(y3, -IE_C7, fmaf(y5, -IE_C1, y7 * -IE_C5)));
    const float o2 = fmaf(y1, IE_C5, fmaf(y3, -IE_C1, fmaf(y5, IE_C7, y7 * IE_C3)));
    const float o3 = fmaf(y1, IE_C7, fmaf(y3, -IE_C5, fmaf(y5, IE_C3, y7 * -IE_C1)));
    v[0 * S] = e0 + o0; v[7 * S] = e0 - o0;
    v[1 * S] = e1 + o1; v[6 * S] = e1 - o1;
    v[2 * S] = e2 + o2; v[5 * S] = e2 - o2;
    v[3 * S] = e3 + o3; v[4 * S] = e3 - o3;
}

template <int S>
__device__ __forceinline__ void idct4_inplace(float *v) {
    const float y0 = v[0 * S], y1 = v[1 * S], y2 = v[2 * S], y3 = v[3 * S];
    const float t2 = y2 * IE_C4;
    const float a = y0 + t2, b = y0 - t2;
    const float p = fmaf(y1, IE_C2, y3 * IE_C6), q = fmaf(y1, IE_C6, -(y3 * IE_C2));
    v[0 * S] = a + p; v[3 * S] = a - p; v[1 * S] = b + q; v[2 * S] = b - q;
}

template <int N>
__device__ __forceinline__ void idct2d_fast(float *x) {
#pragma unroll
    for (int u = 0; u < N; u++) {
        if (N == 8) idct8_inplace<1>(x + u * N); else idct4_inplace<1>(x + u * N);
    }
#pragma unroll
    for (int j = 0; j < N; j++) {
        if (N == 8) idct8_inplace<N>(x + j); else idct4_inplace<N>(x + j);
    }
}

template <int N>
__device__ __forceinline__ void fdct2d_fast(float *x) {
#pragma unroll
    for (int i = 0; i < N; i++) {
        if (N == 8) dct8_inplace<1>(x + i * N); else dct4_inplace<1>(x + i * N);
    }
#pragma unroll
    for (int j = 0; j < N; j++) {
        if (N == 8) dct8_inplace<N>(x + j); else dct4_inplace<N>(x + j);
    }
}

#endif
}  // namespace ie
