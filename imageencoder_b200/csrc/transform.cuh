// Per-block arithmetic in the reference's precision and summation order.
//
//   forward : x = pixel + (-128);  t = +0.0;  for i: for j: t = t + ((cs[i][u]*cs[j][v]) * x[i][j]);
//             e = t * (C(u)*C(v));  q = round_half_away(e / Q[u][v])          (Block.cpp:138-153, algo.cpp:309-331)
//   inverse : d = coef * Q;  X[i][j] = +0.0;  for u: for v: X[i][j] += ((((C(u)*C(v))*cs[i][u])*cs[j][v]) * d[u][v]);
//             X += 128;  pixel = (uint8)clamp(X, 0, 255)   -- truncation      (Block.cpp:162-177,99-119, algo.cpp:343-363)
//
// Every product/sum is an individually rounded binary64 operation (the reference is built without FMA contraction;
// one fused multiply-add changes the bitstream) -> __dmul_rn / __dadd_rn / __ddiv_rn, never a*b+c.
// The table products fl(cs*cs) and fl(fl(cc*cs)*cs) are formed on the host in the reference's association order.
#pragma once
#include "common.cuh"

namespace ie {
#ifdef __CUDACC__

// std::round (half away from zero) of a double whose magnitude is far below 2^51
__device__ __forceinline__ double round_half_away(double v) {
    const double a = fabs(v);
    double r = floor(a);
    if (__dsub_rn(a, r) >= 0.5) r = __dadd_rn(r, 1.0);
    return copysign(r, v);
}

// Exact-order forward coefficient (u,v) of one block.  x[] already holds pixel-128 (exact small integers).
template <int NN>
__device__ __forceinline__ double fdct_coef_exact(const double *__restrict__ fw_uv, const double (&x)[NN], double cc_uv) {
    double acc = 0.0;
#pragma unroll
    for (int ij = 0; ij < NN; ij++) acc = __dadd_rn(acc, __dmul_rn(__ldg(fw_uv + ij), x[ij]));
    return __dmul_rn(acc, cc_uv);
}

__device__ __forceinline__ uint8_t clamp_trunc_u8(double v) {      // Block.cpp:103: uint8_t(std::clamp(v, 0.0, 255.0))
    v = fmin(fmax(v, 0.0), 255.0);
    return (uint8_t)__double2int_rz(v);
}

#endif
}  // namespace ie
