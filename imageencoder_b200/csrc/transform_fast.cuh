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
#include <cmath>
#include <cstring>
#include <utility>

// host + device code (the transform and quantisation arithmetic is also run on the CPU by tests/host/lean_check.cu)
#if defined(__CUDACC__)
#define IE_HD __host__ __device__ __forceinline__
#else
#define IE_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define IE_UNROLL _Pragma("unroll")
#else
#define IE_UNROLL
#endif

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
IE_HD void dct8_inplace(float *v) {
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
IE_HD void dct4_inplace(float *v) {
    const float s0 = v[0 * S] + v[3 * S], d0 = v[0 * S] - v[3 * S];
    const float s1 = v[1 * S] + v[2 * S], d1 = v[1 * S] - v[2 * S];
    v[0 * S] = s0 + s1;
    v[2 * S] = (s0 - s1) * IE_C4;
    v[1 * S] = fmaf(d0, IE_C2, d1 * IE_C6);
    v[3 * S] = fmaf(d0, IE_C6, -(d1 * IE_C2));
}

// inverse of the above (unnormalised DCT-III): x[i] = sum_u y[u] cos((2i+1) u pi / 16)
template <int S>
IE_HD void idct8_inplace(float *v) {
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
IE_HD void idct4_inplace(float *v) {
    const float y0 = v[0 * S], y1 = v[1 * S], y2 = v[2 * S], y3 = v[3 * S];
    const float t2 = y2 * IE_C4;
    const float a = y0 + t2, b = y0 - t2;
    const float p = fmaf(y1, IE_C2, y3 * IE_C6), q = fmaf(y1, IE_C6, -(y3 * IE_C2));
    v[0 * S] = a + p; v[3 * S] = a - p; v[1 * S] = b + q; v[2 * S] = b - q;
}

template <int N>
IE_HD void idct2d_fast(float *x) {
IE_UNROLL
    for (int u = 0; u < N; u++) {
        if (N == 8) idct8_inplace<1>(x + u * N); else idct4_inplace<1>(x + u * N);
    }
IE_UNROLL
    for (int j = 0; j < N; j++) {
        if (N == 8) idct8_inplace<N>(x + j); else idct4_inplace<N>(x + j);
    }
}

template <int N>
IE_HD void fdct2d_fast(float *x) {
IE_UNROLL
    for (int i = 0; i < N; i++) {
        if (N == 8) dct8_inplace<1>(x + i * N); else dct4_inplace<1>(x + i * N);
    }
IE_UNROLL
    for (int j = 0; j < N; j++) {
        if (N == 8) dct8_inplace<N>(x + j); else dct4_inplace<N>(x + j);
    }
}

#endif

// ---------------------------------------------------------------------------------------------------------------------
// Experimental instantiations of the tile kernel's per-block arithmetic (ie_set_option("encode_variant", 1 | 2); the
// default path is untouched).  Same arithmetic as the default path of encode_tiles_kernel -- rn(q~) by the magic constant,
// guard-band test, exact integer DC -- with fewer instructions around it:
//   * two coefficients are quantised per packed f32x2 instruction (FFMA2 x3; every component is an individually rounded
//     FP32 operation) and a zigzag pair goes to the staging area as ONE 32-bit word (PRMT of the two mantissas: the low
//     16 bits of 1.5*2^23 + q ARE q in two's complement, no subtraction of the magic bits);
//   * max bits_needed comes from a packed s16x2 running max / min of those words (VIMNMX3.S16x2, two words per
//     instruction) instead of q ^ (q >> 31) per coefficient: bits_needed is monotone on either side of zero, so
//     max_q bits_needed(q) = max(bits_needed(q_max), bits_needed(q_min))   (utils.hpp:226-243);
//   * non-zero detection per 8-coefficient zigzag segment = OR of the segment's four pair words;
//   * the guard-band mask is updated by one predicated instruction per coefficient;
//   * variant 2: the factorised transform itself in packed operations (below).
// Static SASS of encode_tiles_kernel<8,1>: 3528 instructions (default), 3272 (variant 1), 2984 (variant 2).
// Host + device: tests/host/lean_check.cu runs exactly this code on the CPU against a transcription of the default path
// (the device-only instructions have bit-identical host shims below).
// ---------------------------------------------------------------------------------------------------------------------

namespace lean {

constexpr float kMagicF = 12582912.0f;         // 1.5 * 2^23

// zigzag position -> raster index (algo.cpp:68-87), compile-time copies for the unrolled loops
struct ZZ4 { static constexpr unsigned char t[16] = {0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15}; };
struct ZZ8 {
    static constexpr unsigned char t[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                            41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                            30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
};

IE_HD unsigned f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    unsigned u; memcpy(&u, &f, 4); return u;
#endif
}
// low halfword of a in the low half, low halfword of b in the high half
IE_HD unsigned pack_lo16(unsigned a, unsigned b) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, 0x5410);
#else
    return (a & 0xffffu) | (b << 16);
#endif
}
IE_HD unsigned max3_s16x2(unsigned a, unsigned b, unsigned c) {
#if defined(__CUDA_ARCH__)
    return __vimax3_s16x2(a, b, c);
#else
    auto mx = [](int x, int y, int z) { int m = x > y ? x : y; return m > z ? m : z; };
    const int lo = mx((short)(a & 0xffff), (short)(b & 0xffff), (short)(c & 0xffff));
    const int hi = mx((short)(a >> 16), (short)(b >> 16), (short)(c >> 16));
    return ((unsigned)lo & 0xffffu) | ((unsigned)hi << 16);
#endif
}
IE_HD unsigned min3_s16x2(unsigned a, unsigned b, unsigned c) {
#if defined(__CUDA_ARCH__)
    return __vimin3_s16x2(a, b, c);
#else
    auto mn = [](int x, int y, int z) { int m = x < y ? x : y; return m < z ? m : z; };
    const int lo = mn((short)(a & 0xffff), (short)(b & 0xffff), (short)(c & 0xffff));
    const int hi = mn((short)(a >> 16), (short)(b >> 16), (short)(c >> 16));
    return ((unsigned)lo & 0xffffu) | ((unsigned)hi << 16);
#endif
}
// mask |= bit if |d| >= thr   (one FSETP + one predicated LOP3)
template <unsigned BIT>
IE_HD void or_if_near(unsigned &mask, float d, float thr) {
#if defined(__CUDA_ARCH__)
    asm("{\n\t.reg .pred p;\n\tsetp.ge.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(mask) : "f"(fabsf(d)), "f"(thr), "n"(BIT));
#else
    if (fabsf(d) >= thr) mask |= BIT;
#endif
}

// exact integer DC (see encode_tiles_kernel): q = round_half_away(S / (4 Q00)), S = the block's sample sum (an integer)
IE_HD int dc_quant(float y0, int dc_den2, float dc_rcp) {
    const int S = (int)y0;
    const int n = S < 0 ? -S : S;
    const int num = 2 * n + (dc_den2 >> 1);
    int qq = (int)((float)num * dc_rcp);
    const int rem = num - qq * dc_den2;
    qq += (rem >= dc_den2) ? 1 : 0;
    qq -= (rem < 0) ? 1 : 0;
    return (S < 0) ? -qq : qq;
}

// running state of quantise_block
struct QuantAcc {
    unsigned near_lo, near_hi;      // guard-band mask, bit uv (raster)
    unsigned mx, mn;                // packed s16x2 running max (halves >= 0) / min (halves <= 0)
    unsigned wprev;
};

// ---------------------------------------------------------------------------------------------------------------------
// Variant 2: the factorised transform itself in packed f32x2 operations (FADD2 / FMUL2 / FFMA2: two independent,
// individually rounded FP32 operations per instruction; constants are immediates).  The operation DAG is exactly the one of
// dct8_inplace / dct4_inplace above, so every output is bit-identical to the default fast path's.
//   pass 1 (along a row):     rows 2r and 2r+1 ride in the two halves -> N/2 packed 1-D transforms;
//   pass 2 (along a column):  the first butterfly stage is scalar (it reads single halves of the pass-1 pairs and may write
//                             its results wherever it likes, i.e. into pairs made of columns 2c and 2c+1), the rest is packed;
//   quantisation:             on the raster pairs (u, 2c), (u, 2c+1) pass 2 leaves behind; the zigzag pair words are then
//                             formed by PRMT from any two of the results.
// ---------------------------------------------------------------------------------------------------------------------
IE_HD float2 add2(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
    return __fadd2_rn(a, b);
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
IE_HD float2 sub2(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
    return __fadd2_rn(a, make_float2(-b.x, -b.y));
#else
    return make_float2(a.x - b.x, a.y - b.y);
#endif
}
IE_HD float2 mulc2(float2 a, float c) {
#if defined(__CUDA_ARCH__)
    return __fmul2_rn(a, make_float2(c, c));
#else
    return make_float2(a.x * c, a.y * c);
#endif
}
IE_HD float2 fmac2(float2 a, float c, float2 b) {      // a * c + b
#if defined(__CUDA_ARCH__)
    return __ffma2_rn(a, make_float2(c, c), b);
#else
    return make_float2(fmaf(a.x, c, b.x), fmaf(a.y, c, b.y));
#endif
}

// what follows the first butterfly stage of dct8_inplace / dct4_inplace: s[i] = v[i] + v[N-1-i], d[i] = v[i] - v[N-1-i]
template <int N>
IE_HD void dct_tail2(const float2 *s, const float2 *d, float2 *o) {
    if (N == 8) {
        const float2 t0 = add2(s[0], s[3]), t1 = add2(s[1], s[2]), t2 = sub2(s[0], s[3]), t3 = sub2(s[1], s[2]);
        o[0] = add2(t0, t1);
        o[4] = mulc2(sub2(t0, t1), IE_C4);
        o[2] = fmac2(t2, IE_C2, mulc2(t3, IE_C6));
        o[6] = fmac2(t2, IE_C6, mulc2(t3, -IE_C2));
        o[1] = fmac2(d[0], IE_C1, fmac2(d[1], IE_C3, fmac2(d[2], IE_C5, mulc2(d[3], IE_C7))));
        o[3] = fmac2(d[0], IE_C3, fmac2(d[1], -IE_C7, fmac2(d[2], -IE_C1, mulc2(d[3], -IE_C5))));
        o[5] = fmac2(d[0], IE_C5, fmac2(d[1], -IE_C1, fmac2(d[2], IE_C7, mulc2(d[3], IE_C3))));
        o[7] = fmac2(d[0], IE_C7, fmac2(d[1], -IE_C5, fmac2(d[2], IE_C3, mulc2(d[3], -IE_C1))));
    } else {
        o[0] = add2(s[0], s[1]);
        o[2] = mulc2(sub2(s[0], s[1]), IE_C4);
        o[1] = fmac2(d[0], IE_C2, mulc2(d[1], IE_C6));
        o[3] = fmac2(d[0], IE_C6, mulc2(d[1], -IE_C2));
    }
}

// X2[r * N + k] = (x[2r][k], x[2r+1][k]): the block's samples, two rows per pair.
// Y2[u * (N/2) + c] = (Y[u][2c], Y[u][2c+1]), the unnormalised 2-D DCT outputs.
template <int N>
IE_HD void fdct2d_packed(const float2 *X2, float2 *Y2) {
    constexpr int H = N / 2;
    float2 T2[H][N];                                     // T2[r][v] = (T[2r][v], T[2r+1][v])
IE_UNROLL
    for (int r = 0; r < H; r++) {
        float2 s[H], d[H];
IE_UNROLL
        for (int i = 0; i < H; i++) {
            const float2 a = X2[r * N + i], b = X2[r * N + (N - 1 - i)];
            s[i] = add2(a, b);
            d[i] = sub2(a, b);
        }
        dct_tail2<N>(s, d, T2[r]);
    }
IE_UNROLL
    for (int c = 0; c < H; c++) {
        float2 s[H], d[H], o[N];
IE_UNROLL
        for (int i = 0; i < H; i++) {
            const int j = N - 1 - i;
            // scalar butterflies on single halves of the pass-1 pairs; the results pair up columns 2c and 2c+1
            const float2 pi0 = T2[i >> 1][2 * c], pi1 = T2[i >> 1][2 * c + 1], pj0 = T2[j >> 1][2 * c], pj1 = T2[j >> 1][2 * c + 1];
            const float ti0 = (i & 1) ? pi0.y : pi0.x, ti1 = (i & 1) ? pi1.y : pi1.x;
            const float tj0 = (j & 1) ? pj0.y : pj0.x, tj1 = (j & 1) ? pj1.y : pj1.x;
            s[i] = make_float2(ti0 + tj0, ti1 + tj1);
            d[i] = make_float2(ti0 - tj0, ti1 - tj1);
        }
        dct_tail2<N>(s, d, o);
IE_UNROLL
        for (int u = 0; u < N; u++) Y2[u * H + c] = o[u];
    }
}

// raster pair P = coefficients 2P, 2P+1: bits[] receive 1.5*2^23 + q as raw words (low 16 bits = q)
template <int N, int P>
IE_HD void quantise_raster_pair(const float2 *Y2, const FastQuant &fq, int dc_den2, float dc_rcp, unsigned *bits, QuantAcc &a) {
    constexpr int ua = 2 * P, ub = 2 * P + 1;
    const float2 yy = Y2[P];
#if defined(__CUDA_ARCH__)
    const float2 kk = make_float2(fq.k[ua], fq.k[ub]);
    const float2 rr = __ffma2_rn(yy, kk, make_float2(kMagicF, kMagicF));
    const float2 nrf = __ffma2_rn(rr, make_float2(-1.0f, -1.0f), make_float2(kMagicF, kMagicF));      // -(rr - magic), exact
    const float2 dd = __ffma2_rn(yy, kk, nrf);
    const float rra = rr.x, rrb = rr.y, da = dd.x, db = dd.y;
#else
    const float rra = fmaf(yy.x, fq.k[ua], kMagicF), rrb = fmaf(yy.y, fq.k[ub], kMagicF);
    const float da = fmaf(yy.x, fq.k[ua], -(rra - kMagicF)), db = fmaf(yy.y, fq.k[ub], -(rrb - kMagicF));
#endif
    if (P == 0) {
        bits[0] = (unsigned)dc_quant(yy.x, dc_den2, dc_rcp);
    } else {
        bits[ua] = f2u(rra);
        or_if_near<(1u << (ua & 31))>(ua < 32 ? a.near_lo : a.near_hi, da, fq.thr[ua]);
    }
    bits[ub] = f2u(rrb);
    or_if_near<(1u << (ub & 31))>(ub < 32 ? a.near_lo : a.near_hi, db, fq.thr[ub]);
}

template <int N, int M>
IE_HD void pack_zigzag_pair(const unsigned *bits, unsigned *cfw, unsigned *orseg, QuantAcc &a) {
    constexpr int ua = (N == 8) ? ZZ8::t[2 * M] : ZZ4::t[2 * M];
    constexpr int ub = (N == 8) ? ZZ8::t[2 * M + 1] : ZZ4::t[2 * M + 1];
    const unsigned w = pack_lo16(bits[ua], bits[ub]);
    cfw[M] = w;
    if ((M & 3) == 0) orseg[M >> 2] = w; else orseg[M >> 2] |= w;
    if (M & 1) { a.mx = max3_s16x2(a.mx, a.wprev, w); a.mn = min3_s16x2(a.mn, a.wprev, w); }
    a.wprev = w;
}

template <int N, int... P>
IE_HD void quantise_block_packed_impl(const float2 *Y2, const FastQuant &fq, int dc_den2, float dc_rcp, unsigned *cfw, unsigned *orseg,
                                      QuantAcc &a, std::integer_sequence<int, P...>) {
    unsigned bits[N * N];
    (quantise_raster_pair<N, P>(Y2, fq, dc_den2, dc_rcp, bits, a), ...);
    (pack_zigzag_pair<N, P>(bits, cfw, orseg, a), ...);
}

// Y2[P] = (Y[2P], Y[2P+1]): the NN unnormalised 2-D DCT outputs of the block as raster pairs.  Writes the NN/2 zigzag pair
// words to cfw, returns the guard-band mask (bit uv, raster) in near_lo/near_hi, the per-segment OR words in orseg[NN/8]
// (non-zero <=> the segment holds a non-zero coefficient) and `orbits` with bit_length(orbits) + 1 = max bits_needed over
// the block (0 -> all zero).
template <int N>
IE_HD void quantise_block_packed(const float2 *Y2, const FastQuant &fq, int dc_den2, float dc_rcp, unsigned *cfw, unsigned &near_lo,
                                 unsigned &near_hi, unsigned *orseg, unsigned &orbits) {
    QuantAcc a;
    a.near_lo = 0; a.near_hi = 0; a.mx = 0; a.mn = 0; a.wprev = 0;
    quantise_block_packed_impl<N>(Y2, fq, dc_den2, dc_rcp, cfw, orseg, a, std::make_integer_sequence<int, N * N / 2>{});
    near_lo = a.near_lo; near_hi = a.near_hi;
    const int hmx = (short)(a.mx >> 16), lmx = (short)(a.mx & 0xffffu), hmn = (short)(a.mn >> 16), lmn = (short)(a.mn & 0xffffu);
    const int qmax = hmx > lmx ? hmx : lmx, qmin = hmn < lmn ? hmn : lmn;
    orbits = (unsigned)qmax | (unsigned)(qmin ^ (qmin >> 31));
}

// ---------------------------------------------------------------------------------------------------------------------
// Decode variant 1 (ie_set_option("decode_variant", 1), experimental): the inverse transform and the pixel stage of
// decode_blocks_fast_kernel in packed f32x2 operations -- the operation DAG of idct8_inplace / idct4_inplace above, bit-identical
// outputs.  Mirror image of the forward trick: pass 1 (along a row, rows 2r and 2r+1 in the two halves) is packed up to its LAST
// butterfly stage, which is done on single halves and writes pairs of columns 2c and 2c+1 of one row; pass 2 (along a column)
// is then packed over those column pairs and leaves pairs of neighbouring pixels of a row behind, which is what the pixel stage
// wants.
// ---------------------------------------------------------------------------------------------------------------------
IE_HD float2 fmac2_rd_one(float2 a, float c) {            // a + c per half, rounded towards -infinity (a * 1 is exact)
#if defined(__CUDA_ARCH__)
    return __ffma2_rd(a, make_float2(1.0f, 1.0f), make_float2(c, c));
#else
    // only used with a in [0.5, 255.5] and c = 2^23, where the result is exactly 2^23 + floor(a)
    return make_float2(c + floorf(a.x), c + floorf(a.y));
#endif
}

// everything of idct8_inplace / idct4_inplace before the final butterflies: e[i], o[i] with v[i] = e[i] + o[i],
// v[N-1-i] = e[i] - o[i]
template <int N>
IE_HD void idct_head2(const float2 *y, float2 *e, float2 *o) {
    if (N == 8) {
        const float2 t4 = mulc2(y[4], IE_C4);
        const float2 a = add2(y[0], t4), b = sub2(y[0], t4);
        const float2 pp = fmac2(y[2], IE_C2, mulc2(y[6], IE_C6)), qq = fmac2(y[2], IE_C6, mulc2(y[6], -IE_C2));
        e[0] = add2(a, pp); e[3] = sub2(a, pp); e[1] = add2(b, qq); e[2] = sub2(b, qq);
        o[0] = fmac2(y[1], IE_C1, fmac2(y[3], IE_C3, fmac2(y[5], IE_C5, mulc2(y[7], IE_C7))));
        o[1] = fmac2(y[1], IE_C3, fmac2(y[3], -IE_C7, fmac2(y[5], -IE_C1, mulc2(y[7], -IE_C5))));
        o[2] = fmac2(y[1], IE_C5, fmac2(y[3], -IE_C1, fmac2(y[5], IE_C7, mulc2(y[7], IE_C3))));
        o[3] = fmac2(y[1], IE_C7, fmac2(y[3], -IE_C5, fmac2(y[5], IE_C3, mulc2(y[7], -IE_C1))));
    } else {
        const float2 t2 = mulc2(y[2], IE_C4);
        e[0] = add2(y[0], t2); e[1] = sub2(y[0], t2);                                           // a, b
        o[0] = fmac2(y[1], IE_C2, mulc2(y[3], IE_C6)); o[1] = fmac2(y[1], IE_C6, mulc2(y[3], -IE_C2));   // p, q
    }
}

// X2[r * N + v] = (x[2r][v], x[2r+1][v]): the dequantised coefficients, two rows per pair.
// P2[i * (N/2) + c] = (X[i][2c], X[i][2c+1]): the block's samples before the +128, two neighbouring pixels per pair.
template <int N>
IE_HD void idct2d_packed(const float2 *X2, float2 *P2) {
    constexpr int H = N / 2;
    float2 C2[N][H];                                     // C2[u][c] = (T[u][2c], T[u][2c+1]) after pass 1
IE_UNROLL
    for (int r = 0; r < H; r++) {
        float2 e[H], o[H];
        idct_head2<N>(X2 + r * N, e, o);
        // final butterflies on single halves: row 2r from the .x halves, row 2r+1 from the .y halves
        float lo[N], hi[N];
IE_UNROLL
        for (int i = 0; i < H; i++) {
            lo[i] = e[i].x + o[i].x; lo[N - 1 - i] = e[i].x - o[i].x;
            hi[i] = e[i].y + o[i].y; hi[N - 1 - i] = e[i].y - o[i].y;
        }
IE_UNROLL
        for (int c = 0; c < H; c++) {
            C2[2 * r][c] = make_float2(lo[2 * c], lo[2 * c + 1]);
            C2[2 * r + 1][c] = make_float2(hi[2 * c], hi[2 * c + 1]);
        }
    }
IE_UNROLL
    for (int c = 0; c < H; c++) {
        float2 y[N], e[H], o[H];
IE_UNROLL
        for (int u = 0; u < N; u++) y[u] = C2[u][c];
        idct_head2<N>(y, e, o);
IE_UNROLL
        for (int i = 0; i < H; i++) {
            P2[i * H + c] = add2(e[i], o[i]);
            P2[(N - 1 - i) * H + c] = sub2(e[i], o[i]);
        }
    }
}

// Pixel stage of the fast decode for two neighbouring pixels (decode_blocks_fast_kernel): v = x + 128, clamp to [0.5, 255.5],
// floor by a round-down add of 2^23 (the pixel is the low byte of the result), `unsure` if an integer boundary lies within
// the block's error bound.  Returns the two words whose low bytes are the pixels.
template <unsigned BIT_A, unsigned BIT_B>
IE_HD void pixel_pair(float2 x, float hi_thr, unsigned &fa, unsigned &fb, unsigned &unsure) {
    const float2 v = add2(x, make_float2(128.0f, 128.0f));
    float2 u;
    u.x = fminf(fmaxf(v.x, 0.5f), 255.5f);
    u.y = fminf(fmaxf(v.y, 0.5f), 255.5f);
    const float2 fm = fmac2_rd_one(u, 8388608.0f);                                  // 2^23 + floor(u)
    const float2 fl = add2(fm, make_float2(-8388608.0f, -8388608.0f));              // floor(u), exact
    const float2 g = add2(sub2(u, fl), make_float2(-0.5f, -0.5f));                  // frac - 0.5 (both steps exact)
    or_if_near<BIT_A>(unsure, g.x, hi_thr);
    or_if_near<BIT_B>(unsure, g.y, hi_thr);
    fa = f2u(fm.x); fb = f2u(fm.y);
}

// the same with a per-pixel addend in place of the 128 (P-frames: reference pixel + 128, Block.cpp:110-119)
template <unsigned BIT_A, unsigned BIT_B>
IE_HD void pixel_pair_add(float2 x, float2 addend, float hi_thr, unsigned &fa, unsigned &fb, unsigned &unsure) {
    const float2 v = add2(x, addend);
    float2 u;
    u.x = fminf(fmaxf(v.x, 0.5f), 255.5f);
    u.y = fminf(fmaxf(v.y, 0.5f), 255.5f);
    const float2 fm = fmac2_rd_one(u, 8388608.0f);
    const float2 fl = add2(fm, make_float2(-8388608.0f, -8388608.0f));
    const float2 g = add2(sub2(u, fl), make_float2(-0.5f, -0.5f));
    or_if_near<BIT_A>(unsure, g.x, hi_thr);
    or_if_near<BIT_B>(unsure, g.y, hi_thr);
    fa = f2u(fm.x); fb = f2u(fm.y);
}

// row y of the block: N pixels -> N/4 output words (4 pixels each) + the row's bits of the `unsure` mask (bit = raster index)
template <int N, int Y, int Q4>
IE_HD unsigned pixel_quad(const float2 *P2, float hi_thr, unsigned &unsure_lo, unsigned &unsure_hi) {
    constexpr int ij = Y * N + Q4 * 4;                      // raster index of the quad's first pixel
    unsigned f0, f1, f2, f3;
    unsigned &un = (ij < 32) ? unsure_lo : unsure_hi;        // a quad never straddles bit 32
    pixel_pair<(1u << (ij & 31)), (1u << ((ij + 1) & 31))>(P2[Y * (N / 2) + Q4 * 2], hi_thr, f0, f1, un);
    pixel_pair<(1u << ((ij + 2) & 31)), (1u << ((ij + 3) & 31))>(P2[Y * (N / 2) + Q4 * 2 + 1], hi_thr, f2, f3, un);
#if defined(__CUDA_ARCH__)
    return __byte_perm(__byte_perm(f0, f1, 0x0040), __byte_perm(f2, f3, 0x0040), 0x5410);
#else
    return (f0 & 0xffu) | ((f1 & 0xffu) << 8) | ((f2 & 0xffu) << 16) | ((f3 & 0xffu) << 24);
#endif
}

template <int N, int... I>
IE_HD void pixel_stage_impl(const float2 *P2, float hi_thr, unsigned *outw, unsigned &unsure_lo, unsigned &unsure_hi,
                            std::integer_sequence<int, I...>) {
    ((outw[I] = pixel_quad<N, I / (N / 4), I % (N / 4)>(P2, hi_thr, unsure_lo, unsure_hi)), ...);
}

// outw[y * (N/4) + q4]: the block's pixels, 4 per word (as decode_blocks_fast_kernel keeps them); unsure: bit = raster index
template <int N>
IE_HD void pixel_stage(const float2 *P2, float hi_thr, unsigned *outw, unsigned &unsure_lo, unsigned &unsure_hi) {
    unsure_lo = 0; unsure_hi = 0;
    pixel_stage_impl<N>(P2, hi_thr, outw, unsure_lo, unsure_hi, std::make_integer_sequence<int, N * (N / 4)>{});
}

// ---------------------------------------------------------------------------------------------------------------------
// Exact-queue variant "fast64" (encode_variant 5 / 6 / 7, experimental): a guard-band coefficient is first evaluated in binary64
// in a SHORT chain -- N row dot products b . x[y] (independent chains of N fused multiply-adds), then their combination with
// a[y] (N more) -- instead of the reference's 2*N*N sequentially rounded operations.  Both are evaluations of the same real
// number sum_y sum_k a[y] b[k] x[y][k]; with |a|, |b| <= 1 and |x| <= 383 (P-frame residual - 128) the reference's chain
// (algo.cpp:309-331: N*N rounded factor products, N*N rounded products with x, N*N rounded additions) deviates from it by at
// most (N*N + 2) u S and this one by at most (2N + 2) u S, S = sum |a b x| <= N*N * 383, u = 2^-53: together < 2.2e-10 on the
// sum at N = 8, < 1.2e-10 on the quotient t = sum * cc / m (cc <= 0.5, m >= 1; the three roundings of t add < 1e-11).  So when
// t lies further than 1e-8 (80x that) from every rounding boundary k + 0.5, std::round of the reference's value and of this one
// agree; otherwise -- true ties, SURVEY 0.3, and nothing else in practice -- the caller falls back to the exact-order chain.
// tests/host/lean_check.cu runs this arithmetic on the CPU against the exact chain (random, tie-heavy and saturated blocks).
// ---------------------------------------------------------------------------------------------------------------------
template <int N>
IE_HD double row_dot64(const double *b, const int *xr) {
    double r = 0.0;
IE_UNROLL
    for (int k = 0; k < N; k++) r = fma(b[k], (double)xr[k], r);
    return r;
}

constexpr double kFast64Guard = 1e-8;

// acc = the short-chain sum; returns false if the quotient is within the guard of a rounding boundary (q untouched)
IE_HD bool decide64(double acc, double cc, double m, int &q) {
    const double t = acc * cc / m;
    const double at = fabs(t), fl = floor(at), fr = at - fl;           // fr exact
    if (fabs(fr - 0.5) <= kFast64Guard) return false;
    const double r = (fr > 0.5) ? fl + 1.0 : fl;
    q = (int)(short)(int)(t < 0.0 ? -r : r);                            // Block.cpp:205: int16_t(double)
    return true;
}

}  // namespace lean
}  // namespace ie
