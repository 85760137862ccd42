// Shared declarations of the sm_100a block-codec kernels and their host launchers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <atomic>
#include <string>

#include "../../include/imageencoder_b200.h"

namespace ie {

constexpr int kMaxN = 8;
constexpr int kMaxNN = 64;
constexpr int kMB = 16;                 // MacroBlock size (Block.hpp:14)
constexpr int kHdrWordsMax = 40;        // 1 + 5 + 64*16 + 31 + 45 bits < 40 words

// ---------------------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------------------
void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

#define IE_CUDA(call)                                                              \
    do {                                                                           \
        cudaError_t e__ = (call);                                                  \
        if (e__ != cudaSuccess) return ::ie::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)
#define IE_TRY(call)                  \
    do {                              \
        int rc__ = (call);            \
        if (rc__ != IE_OK) return rc__; \
    } while (0)

// ---------------------------------------------------------------------------------------------------------
// Per-block-size arithmetic tables, resident in HBM (read through L1, warp-uniform addresses).
//   fw [uv*NN + ij] = fl(cs[i][u] * cs[j][v])                 forward product  (algo.cpp:318-319)
//   inv[uv*NN + ij] = fl(fl(cc[uv] * cs[i][u]) * cs[j][v])    inverse product  (algo.cpp:352-354)
//   cc [uv]         = fl(C(u) * C(v))                         (algo.cpp:294-297,325)
//   zz [k]          = raster index of the k-th zigzag position (algo.cpp:68-87); izz = inverse
// ---------------------------------------------------------------------------------------------------------
struct BlockTables {
    double fw[kMaxNN * kMaxNN];
    double inv[kMaxNN * kMaxNN];
    double cc[kMaxNN];
    double cs[kMaxNN];
    uint8_t zz[kMaxNN];
    uint8_t izz[kMaxNN];
};

// One pixel of the inverse transform in the reference's order and precision (algo.cpp:343-363: u outer, v inner; Block.cpp:165-168
// multiplies the coefficient by the quantiser first; zero coefficients contribute exact zeroes and are skipped).  `nzmask` holds
// the raster positions of the block's non-zero coefficients, `cf` its zigzag-ordered staging row.  The additions are one
// dependent chain in the reference's order; everything else of a term (zigzag index, coefficient, quantiser, table entry: the
// table read goes to L2) does not depend on the running sum, so the terms are fetched R per round and their latencies overlap
// (R = 4 pays for 8x8 blocks and in the P-frame tiles; the 4x4 block decoder, few non-zeros per block, is fastest with R = 1).
#ifdef __CUDACC__
template <int NN, int R, typename MaskT, typename QuantT>
__device__ __forceinline__ double exact_inverse_pixel(MaskT nzmask, int ij, const short *cf, const BlockTables *tab, const QuantT &quant) {
    double acc = 0.0;
    MaskT nz = nzmask;
    if constexpr (R == 1) {                                                                     // term by term
        while (nz) {
            const int uv = (sizeof(MaskT) == 8 ? __ffsll((long long)nz) : __ffs((int)nz)) - 1;
            nz &= nz - 1;
            const double d = __dmul_rn((double)(int)cf[tab->izz[uv]], quant.m[uv]);            // Block.cpp:165-168
            acc = __dadd_rn(acc, __dmul_rn(__ldg(tab->inv + uv * NN + ij), d));                 // algo.cpp:352-355
        }
        return acc;
    } else {
    while (nz) {
        int uv[R];
#pragma unroll
        for (int t = 0; t < R; t++) {
            uv[t] = (sizeof(MaskT) == 8 ? __ffsll((long long)nz) : __ffs((int)nz)) - 1;      // -1 once the mask is empty
            nz &= nz - 1;
        }
        double a[R], d[R];
#pragma unroll
        for (int t = 0; t < R; t++) {
            const int k = max(uv[t], 0);
            a[t] = __ldg(tab->inv + k * NN + ij);
            d[t] = __dmul_rn((double)(int)cf[tab->izz[k]], quant.m[k]);                        // Block.cpp:165-168
        }
#pragma unroll
        for (int t = 0; t < R; t++)
            if (uv[t] >= 0) acc = __dadd_rn(acc, __dmul_rn(a[t], d[t]));                       // algo.cpp:352-355
    }
    return acc;
    }
}
#endif

struct HostTables {
    BlockTables t4, t8;
};
const HostTables &host_tables();

// Quantisation matrix as the kernels take it (by value, in the kernel parameter block).
struct QuantParam {
    double m[kMaxNN];      // double(Q[u][v])   (MatrixReader.cpp:128)
};

// Header bits prepared on the host (ImageEncoder.cpp:84-94, MatrixReader.cpp:144-158, VideoEncoder.cpp:60-73),
// MSB-first, zero padded.
struct HeaderParam {
    uint32_t words[kHdrWordsMax];
    uint32_t bits;
};
// lead_bit: the '0' "no Huffman" bit of non-Huffman builds.  video: append frames/gop/merange (15 bits each).
int build_header(HeaderParam &h, int N, const uint16_t *quant, int use_rle, uint32_t W, uint32_t H, int lead_bit,
                 int video, uint32_t frames, uint32_t gop, uint32_t merange);

struct DeviceState {
    int device = -1;
    BlockTables *d_t4 = nullptr, *d_t8 = nullptr;
    int sm_count = 0;
};
int get_device_state(DeviceState **out);      // ie_init(current device) if needed

// ---------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
// util::bits_needed (utils.hpp:226-243) for a value that fits int16: minimal two's complement width.
__device__ __forceinline__ int dev_bits_needed(int v) { return 33 - __clz(v ^ (v >> 31)); }
// util::ffs (utils.hpp:210-216) with ffs(0) = 0
__device__ __forceinline__ int dev_ffs(unsigned v) { return 32 - __clz(v); }

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
#endif

#ifdef __CUDACC__
// Programmatic dependent launch: the grid is set up while its predecessor in the stream drains; the kernel must execute
// pdl_wait() before it touches anything an earlier kernel (or copy) of the stream produced.  The decode path is a chain of
// short, latency-bound kernels (a dozen per video frame), where the launch latency otherwise adds up.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif
}  // namespace ie
