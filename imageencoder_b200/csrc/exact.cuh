// Exact-order evaluation of single coefficients (the fallback behind the guard band of transform_fast.cuh), shared by the
// tile kernels of encode_image.cu and the fused stream kernel of encode_fused.cu.
#pragma once
#include "encode_image.cuh"
#include "transform_fast.cuh"

namespace ie {

// Sample (pixel, or pixel - reference pixel for P-frames) of block (byi, bxi) at raster position ij, minus 128, re-read
// from global memory (used by the exact fallback, which runs on a different thread than the one that owns the block).
struct ExactCtx {              // the few launch constants the exact path needs (passed by value: no param-space copy)
    const uint8_t *src, *ref;
    const short *res_coord;
    const BlockTables *tab;
    size_t pitch;
    unsigned bx, mbx;
};

template <int N, bool PF>
__device__ __forceinline__ double exact_sample(const ExactCtx &p, const uint8_t *src, unsigned byi, unsigned bxi, int rx, int ry, int ij) {
    const int y = ij / N, x = ij % N;
    int v = (int)__ldg(src + (size_t)(byi * N + y) * p.pitch + (size_t)bxi * N + x);
    if (PF) v -= (int)__ldg(p.ref + (size_t)(ry + y) * p.pitch + rx + x);                 // Block.cpp:262
    return __dadd_rn((double)v, -128.0);                                                    // Block.cpp:141-143
}

// One coefficient of one block in the reference's exact order and precision (algo.cpp:309-331, Block.cpp:152).
// The 2*N*N-operation chain is inherently sequential (every partial sum is rounded); everything that does not depend on
// the running sum is fetched before it starts: the N pixel rows and the 2N cosines whose rounded products
// fl(cs[i][u] * cs[j][v]) are the reference's factors (algo.cpp:318-319) -- one round of loads instead of N*N table reads.
template <int N, bool PF>
__device__ __noinline__ int exact_coefficient(const ExactCtx p, unsigned gb, int uv, double m_uv) {
    const unsigned byi = gb / p.bx, bxi = gb - byi * p.bx;
    int rx = 0, ry = 0;
    if (PF) {
        const unsigned mb = (byi >> 2) * p.mbx + (bxi >> 2);
        rx = p.res_coord[2 * mb] + (int)(bxi & 3) * 4;
        ry = p.res_coord[2 * mb + 1] + (int)(byi & 3) * 4;
    }
    const int u = uv / N, v = uv % N;
    const uint8_t *blk = p.src + (size_t)(byi * N) * p.pitch + (size_t)bxi * N;
    unsigned lo[N], hi[N];
    double a[N], b[N];
#pragma unroll
    for (int y = 0; y < N; y++) {
        hi[y] = 0;
        if (N == 8) { const uint2 w = *reinterpret_cast<const uint2 *>(blk + (size_t)y * p.pitch); lo[y] = w.x; hi[y] = w.y; }
        else lo[y] = *reinterpret_cast<const unsigned *>(blk + (size_t)y * p.pitch);
        a[y] = __ldg(p.tab->cs + y * N + u);
        b[y] = __ldg(p.tab->cs + y * N + v);
    }
    double acc = 0.0;
#pragma unroll
    for (int y = 0; y < N; y++) {
#pragma unroll
        for (int k = 0; k < N; k++) {
            int px = (int)(((k < 4 ? lo[y] : hi[y]) >> (8 * (k & 3))) & 0xffu);
            if (PF) px -= (int)__ldg(p.ref + (size_t)(ry + y) * p.pitch + rx + k);                  // Block.cpp:262
            const double x = (double)(px - 128);                                                   // Block.cpp:141-143 (exact)
            acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(a[y], b[k]), x));                             // algo.cpp:318-320
        }
    }
    const double e = __dmul_rn(acc, p.tab->cc[uv]);
    return (int)(short)__double2int_rz(round_half_away(__ddiv_rn(e, m_uv)));
}

// Short-chain binary64 evaluation of one queued coefficient (transform_fast.cuh, "fast64"): true and q set if the quotient is
// clear of every rounding boundary, false if only the exact-order chain (exact_coefficient) can decide.
template <int N, bool PF>
__device__ __forceinline__ bool fast64_coefficient(const ExactCtx &p, unsigned gb, int uv, double m_uv, int &q) {
    const unsigned byi = gb / p.bx, bxi = gb - byi * p.bx;
    int rx = 0, ry = 0;
    if (PF) {
        const unsigned mb = (byi >> 2) * p.mbx + (bxi >> 2);
        rx = p.res_coord[2 * mb] + (int)(bxi & 3) * 4;
        ry = p.res_coord[2 * mb + 1] + (int)(byi & 3) * 4;
    }
    const int u = uv / N, v = uv % N;
    const uint8_t *blk = p.src + (size_t)(byi * N) * p.pitch + (size_t)bxi * N;
    unsigned lo[N], hi[N];
    double a[N], b[N];
#pragma unroll
    for (int y = 0; y < N; y++) {
        hi[y] = 0;
        if (N == 8) { const uint2 w = *reinterpret_cast<const uint2 *>(blk + (size_t)y * p.pitch); lo[y] = w.x; hi[y] = w.y; }
        else lo[y] = *reinterpret_cast<const unsigned *>(blk + (size_t)y * p.pitch);
        a[y] = __ldg(p.tab->cs + y * N + u);
        b[y] = __ldg(p.tab->cs + y * N + v);
    }
    double acc = 0.0;
#pragma unroll
    for (int y = 0; y < N; y++) {
        int xr[N];
#pragma unroll
        for (int k = 0; k < N; k++) {
            int px = (int)(((k < 4 ? lo[y] : hi[y]) >> (8 * (k & 3))) & 0xffu);
            if (PF) px -= (int)__ldg(p.ref + (size_t)(ry + y) * p.pitch + rx + k);                  // Block.cpp:262
            xr[k] = px - 128;                                                                      // Block.cpp:141-143
        }
        acc = fma(a[y], lean::row_dot64<N>(b, xr), acc);
    }
    return lean::decide64(acc, p.tab->cc[uv], m_uv, q);
}

}  // namespace ie
