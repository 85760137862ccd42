// Image-block decode.
//   parse_blocks_kernel : finds every block's first bit.  The format has no markers and each block's size is in its own
//                         header (Block.cpp:443-444), so the offsets form a dependent chain (ImageDecoder.cpp:89-92
//                         "Reading raw must happen in sequence"); one lane per stream walks it.
//   decode_blocks_kernel: lane per block: read fields, sign-extend (utils.hpp:265-269), x Q, inverse DCT in the
//                         reference's summation order with exact zero-skipping, +128, clamp, truncate to u8
//                         (Block.cpp:441-472, 162-177, 99-107; algo.cpp:343-363).
#include "decode_image.cuh"
#include "stage.cuh"
#include "transform.cuh"
#include "transform_fast.cuh"
#include <atomic>

namespace ie {

__device__ constexpr unsigned char kZigzagInvD4[16] = {0, 1, 5, 6, 2, 4, 7, 12, 3, 8, 11, 13, 9, 10, 14, 15};
__device__ constexpr unsigned char kZigzagInvD8[64] = {
    0,  1,  5,  6,  14, 15, 27, 28, 2,  4,  7,  13, 16, 26, 29, 42, 3,  8,  12, 17, 25, 30, 41, 43, 9,  11, 18, 24, 31, 40, 44, 53,
    10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38, 46, 51, 55, 60, 21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63};

// n <= 25 bits at bit position p (MSB-first).  Bits past the end read as 0 (BitStream.cpp:17-20).
__device__ __forceinline__ unsigned read_bits(const uint8_t *__restrict__ s, unsigned long long total_bits, unsigned long long p, int n) {
    if (n == 0) return 0u;
    const unsigned long long nbytes = (total_bits + 7) >> 3;
    const unsigned long long b = p >> 3;
    unsigned v = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const unsigned long long bi = b + i;
        const unsigned byte = (bi < nbytes) ? (unsigned)__ldg(s + bi) : 0u;
        v = (v << 8) | byte;
    }
    const int sh = 32 - (int)(p & 7) - n;
    return (v >> sh) & ((n >= 32) ? 0xffffffffu : ((1u << n) - 1u));
}

__global__ void parse_blocks_kernel(const DecodeParams p) {
    const unsigned img = blockIdx.x;
    if (threadIdx.x != 0) return;
    const uint8_t *s = p.enc + (size_t)img * p.enc_stride;
    const unsigned long long total = p.enc_bits[img];
    unsigned long long pos = p.cursor ? (*p.cursor + p.skip_bits) : p.start_bit[img];
    unsigned long long *off = p.block_off + (size_t)img * (p.block_off_stride ? p.block_off_stride : (size_t)p.nblocks + 1);
    const int NN = p.N * p.N;
    for (unsigned k = 0; k < p.nblocks; k++) {
        off[k] = pos;
        const unsigned w = read_bits(s, total, pos, 4);
        unsigned long long q = min(pos + 4, total);
        unsigned len = NN;
        if (p.use_rle) { len = read_bits(s, total, q, (int)w); q = min(q + w, total); }
        q = min(q + (unsigned long long)len * w, total);        // reads past the end do not advance (BitStream.cpp:17-20)
        pos = q;
    }
    off[p.nblocks] = pos;
    if (p.cursor) *p.cursor = pos;
}

// four bytes at any alignment from the two aligned words that hold them
__device__ __forceinline__ unsigned load_u8x4_any(const uint8_t *p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const unsigned *wp = reinterpret_cast<const unsigned *>(a & ~(uintptr_t)3);
    const unsigned k = (unsigned)(a & 3);
    const unsigned w0 = __ldg(wp), w1 = k ? __ldg(wp + 1) : 0u;
    return __byte_perm(w0, w1, 0x3210u + 0x1111u * k);
}

template <int N, bool ADD>
__global__ void __launch_bounds__(256) decode_blocks_kernel(const DecodeParams p) {
    constexpr int NN = N * N;
    constexpr int STRIDE = NN + 2;
    __shared__ short s_coef[256 * STRIDE];
    const unsigned img = blockIdx.y;
    const unsigned gb = p.block_base + blockIdx.x * 256 + threadIdx.x;
    if (gb >= (p.block_end ? p.block_end : p.nblocks)) return;
    const uint8_t *s = p.enc + (size_t)img * p.enc_stride;
    const unsigned long long total = p.enc_bits[img];
    const unsigned long long *off = p.block_off + (size_t)img * (p.block_off_stride ? p.block_off_stride : (size_t)p.nblocks + 1);
    const BlockTables *tab = p.tab;

    unsigned long long pos = off[gb];
    const int w = (int)read_bits(s, total, pos, 4);
    pos = min(pos + 4, total);
    int len = NN;
    if (p.use_rle) { len = (int)read_bits(s, total, pos, w); pos = min(pos + w, total); }
    if (len > NN) { atomicExch(p.err, IE_EFORMAT); len = NN; }   // the reference indexes out of bounds here
    short *cf = s_coef + threadIdx.x * STRIDE;
    for (int k = 0; k < NN; k++) {
        int v = 0;
        if (k < len) {
            const unsigned raw = read_bits(s, total, pos, w);
            pos = min(pos + w, total);
            v = (int)(short)(unsigned short)(raw << (16 - w)) >> (16 - w);      // util::shift_signed<int16_t>
            if (w == 0) v = 0;
        }
        cf[k] = (short)v;
    }
    double X[NN];
#pragma unroll
    for (int i = 0; i < NN; i++) X[i] = 0.0;
#pragma unroll 1
    for (int uv = 0; uv < NN; uv++) {
        const int c = cf[tab->izz[uv]];
        if (c != 0) {                                                            // adding +-0 never changes the sum
            const double d = __dmul_rn((double)c, p.quant.m[uv]);                // Block.cpp:165-168
            const double *t = tab->inv + uv * NN;
#pragma unroll
            for (int ij = 0; ij < NN; ij++) X[ij] = __dadd_rn(X[ij], __dmul_rn(__ldg(t + ij), d));   // algo.cpp:352-355
        }
    }
    const unsigned byi = gb / p.bx, bxi = gb - byi * p.bx;
    uint8_t *dst = p.out + (size_t)img * p.out_stride;
    // P-frames of the whole-stream video decode: the prediction comes from the reference frame (DecodeParams::mc_coord)
    const uint8_t *pred = nullptr;
    if (ADD && N == 4 && p.mc_coord) {
        const unsigned mb = (byi >> 2) * p.mbx + (bxi >> 2);
        const short *mc = p.mc_coord + ((size_t)img * p.mc_stride + mb) * 2;
        pred = dst - p.ref_delta + (size_t)(mc[1] + (int)(byi & 3) * 4) * p.pitch + (mc[0] + (int)(bxi & 3) * 4);
    }
#pragma unroll
    for (int y = 0; y < N; y++) {
        unsigned lo = 0, hi = 0;
        uint8_t *row = dst + (size_t)(byi * N + y) * p.pitch + (size_t)bxi * N;
        unsigned cur_lo = 0;
        if (ADD) cur_lo = pred ? load_u8x4_any(pred + (size_t)y * p.pitch) : *reinterpret_cast<const unsigned *>(row);
#pragma unroll
        for (int x = 0; x < N; x++) {
            double v = __dadd_rn(X[y * N + x], 128.0);                           // Block.cpp:173-175
            if (ADD) v = __dadd_rn((double)(int)((cur_lo >> (8 * (x & 3))) & 0xff), v);   // Block.cpp:114-116
            const unsigned px = clamp_trunc_u8(v);                               // Block.cpp:99-107 (truncation)
            if (x < 4) lo |= px << (8 * x); else hi |= px << (8 * (x - 4));
        }
        if (N == 8) *reinterpret_cast<uint2 *>(row) = make_uint2(lo, hi);
        else *reinterpret_cast<unsigned *>(row) = lo;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Fast decode: fields through a 64-bit bit buffer fed by aligned word loads; separable FP32 inverse DCT; every pixel
// whose fast value lies within the block's error bound of an integer boundary (where the reference's truncation,
// Block.cpp:103, could go either way) is recomputed in the reference's exact order over the non-zero coefficients.
//   error bound, u = 2^-24, S = sum |C(u)C(v) Q c| (the dequantised inputs), |cos| <= 1:
//     one 1-D pass on inputs y with A = sum |y_i|: the even part carries <= 4 u A_even (stored constant + product + two sums),
//     the odd fma chain <= 5 u A_odd, the final butterfly adds u A -- at most 6 u A per output.  Pass 2 adds 6 u S of its own
//     and carries pass 1's 6 u (row sums) through |cos| <= 1: 12 u S; the inputs themselves (float(Q C C) times the
//     coefficient) are off by <= 2 u each: |X_fast - X_real| <= 14 u S (1 + O(u)).
//     Measured worst case over 2 M random / sparse / single-coefficient / sign-aligned blocks: 2.8 u S on exact inputs (tools/idct_error_bound.py).
//   The +128 / +prediction adds round at <= u (S + 383) each.  delta = 18 u S + 2 u (S + 383) + 2e-6 is used (round 1 used
//   32 u S: twice as many pixels took the exact path, a third of this kernel's stall samples).
//   Pixels that are certainly clamped (v < -delta or v > 255 + delta) need no check.
// ---------------------------------------------------------------------------------------------------------
__constant__ unsigned char c_zz4[16] = {0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15};
__constant__ unsigned char c_zz8[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                        41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                        30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct BitReader {
    const unsigned *words;
    unsigned long long nwords, total;
    unsigned long long wi;       // next word to load
    unsigned long long buf;      // upcoming bits, left aligned
    int n;                       // valid bits in buf
    __device__ __forceinline__ unsigned load(unsigned long long i) const {
        if (i >= nwords) return 0u;
        unsigned w = __byte_perm(__ldg(words + i), 0, 0x0123);
        if ((i + 1) * 32 > total) w &= ~(0xFFFFFFFFu >> (unsigned)(total - i * 32));      // bits past the end read 0
        return w;
    }
    __device__ __forceinline__ void init(const uint8_t *s, unsigned long long total_bits, unsigned long long pos) {
        words = reinterpret_cast<const unsigned *>(s);
        total = total_bits; nwords = (total_bits + 31) >> 5;
        wi = pos >> 5;
        const unsigned sh = (unsigned)(pos & 31);
        buf = ((unsigned long long)load(wi) << 32) | load(wi + 1);
        buf <<= sh;
        n = 64 - (int)sh;
        wi += 2;
    }
    __device__ __forceinline__ unsigned get(int k) {          // k <= 16
        if (k == 0) return 0u;
        if (n < 32) { buf |= (unsigned long long)load(wi++) << (32 - n); n += 32; }
        const unsigned v = (unsigned)(buf >> (64 - k));
        buf <<= k;
        n -= k;
        return v;
    }
};

#ifndef IE_DEC_R8
#define IE_DEC_R8 4
#endif
#ifndef IE_DEC_R4
#define IE_DEC_R4 1
#endif
#define IE_DEC_R(N) ((N) == 8 ? IE_DEC_R8 : IE_DEC_R4)
#ifndef IE_DEC_LOWSPLIT
#define IE_DEC_LOWSPLIT 1
#endif
template <int N, bool ADD, int VAR>
__device__ __forceinline__ void decode_blocks_fast_body(const DecodeParams &p) {
    pdl_wait();
    constexpr int NN = N * N;
    constexpr int STRIDE = NN + 2;
    // the bits of the CTA's 128 blocks are one contiguous span of the stream (<= 128 * (4 + 16 + 16 NN) bits): staged with
    // 16-byte cp.async, then every lane reads its block's fields from shared memory (no bounds or end-of-stream tests)
    // (+ one maximal block: in a truncated stream the last block's fields run past the end of the stream, where they must
    // read as zero bits, BitStream.cpp:17-20, not as whatever the shared memory held)
    constexpr unsigned kMaxBlockBits = 4 + 16 + 16 * NN;
    constexpr unsigned kStage = (129u * kMaxBlockBits + 31) / 32 + 12;
    __shared__ short s_coef[128 * STRIDE];
    __shared__ __align__(16) unsigned s_bits[kStage];
    const unsigned img = blockIdx.y;
    const unsigned first = p.block_base + blockIdx.x * 128;
    const unsigned gb = first + threadIdx.x;
    const unsigned range_end = p.block_end ? p.block_end : p.nblocks;
    const uint8_t *s = p.enc + (size_t)img * p.enc_stride;
    const unsigned long long total = p.enc_bits[img];
    const unsigned long long *off = p.block_off + (size_t)img * (p.block_off_stride ? p.block_off_stride : (size_t)p.nblocks + 1);
    const BlockTables *tab = p.tab;
    const StagedStream st = stage_stream(s_bits, kStage, s, total, off[first], off[min(first + 128u, range_end)] + kMaxBlockBits);
    if (gb >= range_end) return;

    // ---- fields (Block.cpp:441-472) -------------------------------------------------------------------------
    // a valid chain keeps the CTA's 128 blocks within 128 maximal blocks of its first one.  A malformed stream does not: once
    // the parser meets a length field > N*N (IE_EFORMAT) every later block is placed at the end of the stream, which can lie
    // far outside the staged span -- such a block reads (unspecified, the call fails) bits from inside the staging area
    const unsigned long long rel = off[gb] - st.base;
    unsigned pos = (unsigned)min(rel, (unsigned long long)(kStage * 32u - kMaxBlockBits - 64u));
    auto peek = [&](unsigned at) -> unsigned {                 // 32 stream bits starting at `at`, MSB first
        const unsigned i = at >> 5;
        return __funnelshift_l(__byte_perm(st.w[i + 1], 0, 0x0123), __byte_perm(st.w[i], 0, 0x0123), at & 31u);
    };
    const unsigned head = peek(pos);
    const int w = (int)(head >> 28);
    int len = NN;
    if (p.use_rle) len = w ? (int)((head << 4) >> (32 - w)) : 0;
    pos += 4u + (p.use_rle ? (unsigned)w : 0u);
    if (len > NN) { atomicExch(p.err, IE_EFORMAT); len = NN; }   // the reference indexes out of bounds here
    short *cf = s_coef + threadIdx.x * STRIDE;
    unsigned *cfw = reinterpret_cast<unsigned *>(cf);
#pragma unroll
    for (int j = 0; j < NN / 2; j++) cfw[j] = 0u;
    if (w != 0) {
        const int sh = 32 - w;
        // two fields per 32-bit window (w <= 16), stored as one word: util::shift_signed<int16_t> = the top w bits, sign extended
        int k = 0;
        for (; k + 1 < len; k += 2) {
            const unsigned x = peek(pos);
            const int v0 = (int)x >> sh, v1 = (int)(x << w) >> sh;
            pos += 2u * (unsigned)w;
            cfw[k >> 1] = __byte_perm((unsigned)v0, (unsigned)v1, 0x5410);
        }
        if (k < len) {
            const int v = (int)peek(pos) >> sh;
            if (v != 0) cf[k] = (short)v;
        }
    }
    // ---- fast inverse transform --------------------------------------------------------------------------------
    float x[NN];
    float S = 0.f;
    unsigned long long nzmask = 0;
    // 8x8: the first 36 zigzag positions are the diagonals u + v <= 7.  When no block of the warp is longer (the usual case with
    // RLE), the other 28 coefficients are known zeroes: their conversion is skipped (warp-uniform branch).
    constexpr int kLow = (N == 8 && IE_DEC_LOWSPLIT) ? 36 : NN;
    const bool any_high = (kLow < NN) ? (__any_sync(__activemask(), len > kLow) != 0) : true;
#pragma unroll
    for (int uv = 0; uv < NN; uv++) {
        const int k = (N == 8) ? kZigzagInvD8[uv] : kZigzagInvD4[uv];
        if (k >= kLow) { x[uv] = 0.f; continue; }
        const int c = cf[k];
        if (c != 0) nzmask |= 1ull << uv;                                     // raster positions of the non-zero coefficients
        const float d = (float)c * p.k2[uv];                                  // coefficient * Q * C(u)C(v)
        x[uv] = d;
        S += fabsf(d);
    }
    if (kLow < NN && any_high) {
#pragma unroll
        for (int uv = 0; uv < NN; uv++) {
            const int k = (N == 8) ? kZigzagInvD8[uv] : kZigzagInvD4[uv];
            if (k < kLow) continue;
            const int c = cf[k];
            if (c != 0) nzmask |= 1ull << uv;
            const float d = (float)c * p.k2[uv];
            x[uv] = d;
            S += fabsf(d);
        }
    }
    if (VAR != 1 || ADD) idct2d_fast<N>(x);
    const float delta = (18.f * S + 2.f * (S + 383.f)) * 5.9604645e-8f * 1.0001f + 2e-6f;
    // unsure: |frac - 0.5| >= 0.5 - delta; absurd coefficients (delta >= 0.49) take the exact path everywhere
    const float hi_thr = (delta < 0.49f) ? 0.5f - delta : 0.f;
    const unsigned byi = gb / p.bx, bxi = gb - byi * p.bx;
    uint8_t *dst = p.out + (size_t)img * p.out_stride + (size_t)(byi * N) * p.pitch + (size_t)bxi * N;
    unsigned outw[N * (N / 4)];
    unsigned long long unsure = 0;
    // ADD mode (N = 4, P-frames): where the prediction comes from, and the four words read from there (the exact path needs them)
    const uint8_t *pred = nullptr;
    unsigned predw[N];
    if (ADD && N == 4 && p.mc_coord) {
        const unsigned mb = (byi >> 2) * p.mbx + (bxi >> 2);
        const short *mc = p.mc_coord + ((size_t)img * p.mc_stride + mb) * 2;
        pred = p.out + (size_t)img * p.out_stride - p.ref_delta + (size_t)(mc[1] + (int)(byi & 3) * 4) * p.pitch + (mc[0] + (int)(bxi & 3) * 4);
    }
    if (VAR == 1 && !ADD) {
        // variant 1 (transform_fast.cuh, lean::): inverse transform and pixel stage in packed f32x2 operations, same bits
        float2 x2[NN / 2], p2[NN / 2];
#pragma unroll
        for (int r2 = 0; r2 < N / 2; r2++)
#pragma unroll
            for (int v = 0; v < N; v++) x2[r2 * N + v] = make_float2(x[(2 * r2) * N + v], x[(2 * r2 + 1) * N + v]);
        lean::idct2d_packed<N>(x2, p2);
        unsigned ulo, uhi;
        lean::pixel_stage<N>(p2, hi_thr, outw, ulo, uhi);
        unsure = ((unsigned long long)uhi << 32) | ulo;
    } else
#pragma unroll
    for (int y = 0; y < N; y++) {
        unsigned curw[N / 4];
        if (ADD) {
            if (N == 8) { const uint2 c2 = *reinterpret_cast<const uint2 *>(dst + (size_t)y * p.pitch); curw[0] = c2.x; curw[N / 4 - 1] = c2.y; }
            else if (pred) curw[0] = load_u8x4_any(pred + (size_t)y * p.pitch);            // Block.cpp:481-496 folded in
            else curw[0] = *reinterpret_cast<const unsigned *>(dst + (size_t)y * p.pitch);
            if (N == 4) predw[y] = curw[0];
        }
#pragma unroll
        for (int q4 = 0; q4 < N / 4; q4++) {
            unsigned fl[4];
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int ij = y * N + q4 * 4 + b;
                float v = x[ij] + 128.f;
                if (ADD) v += (float)((curw[q4] >> (8 * b)) & 0xffu);
                // the reference truncates after clamping to [0, 255] (Block.cpp:103): floor(u) with u clamped to
                // [0.5, 255.5] is the same pixel, and u sits exactly on x.5 wherever the clamp decided
                const float u = fminf(fmaxf(v, 0.5f), 255.5f);
                const float fm = __fadd_rd(u, 8388608.0f);                    // 2^23 + floor(u): the pixel is the low byte
                const float frac = u - (fm - 8388608.0f);                     // [0, 1), exact
                // an integer boundary within delta of u could flip the truncation
                if (fabsf(frac - 0.5f) >= hi_thr) unsure |= 1ull << ij;
                fl[b] = __float_as_uint(fm);
            }
            const unsigned word = __byte_perm(__byte_perm(fl[0], fl[1], 0x0040), __byte_perm(fl[2], fl[3], 0x0040), 0x5410);
            outw[y * (N / 4) + q4] = word;
        }
    }
    // ---- exact recomputation of the uncertain pixels (algo.cpp:343-363 order: u outer, v inner; zeros skipped) -------
    while (unsure) {
        const int ij = __ffsll((long long)unsure) - 1;
        unsure &= unsure - 1;
        const double acc = exact_inverse_pixel<NN, IE_DEC_R(N)>(nzmask, ij, cf, tab, p.quant);
        double v = __dadd_rn(acc, 128.0);                                               // Block.cpp:173-175
        if (ADD) {                                                                       // Block.cpp:114-116
            unsigned pw = 0;
            if (N == 4) {
#pragma unroll
                for (int r = 0; r < N; r++) if (r == ij / N) pw = predw[r];
            }
            const int pp = (N == 4) ? (int)((pw >> (8 * (ij % N))) & 0xffu) : (int)dst[(size_t)(ij / N) * p.pitch + (ij % N)];
            v = __dadd_rn((double)pp, v);
        }
        const unsigned px = clamp_trunc_u8(v);
        const int wi = ij >> 2, bsh = 8 * (ij & 3);
#pragma unroll
        for (int r = 0; r < N * (N / 4); r++)
            if (r == wi) outw[r] = (outw[r] & ~(0xffu << bsh)) | (px << bsh);
    }
#pragma unroll
    for (int y = 0; y < N; y++) {
        if (N == 8) *reinterpret_cast<uint2 *>(dst + (size_t)y * p.pitch) = make_uint2(outw[2 * y], outw[2 * y + 1]);
        else *reinterpret_cast<unsigned *>(dst + (size_t)y * p.pitch) = outw[y];
    }
}

template <int N, bool ADD>
__global__ void __launch_bounds__(128) decode_blocks_fast_kernel(const DecodeParams p) {
    decode_blocks_fast_body<N, ADD, 0>(p);
}
// decode variant 1: same body with the packed inverse transform + pixel stage, held to 6 CTAs per SM like the default
template <int N>
__global__ void __launch_bounds__(128, 6) decode_blocks_lean_kernel(const DecodeParams p) {
    decode_blocks_fast_body<N, false, 1>(p);
}

extern std::atomic<int> g_exact_transform;
// 1 = inverse transform + pixel stage of the fast decode in packed f32x2 operations (image blocks and I-frames; the default since
// round 2: arithmetic checked on the CPU, tests/host/lean_check.cu, pixel-identical on the B200 in every decode test, 3 % faster),
// 0 = the scalar kernel
std::atomic<int> g_decode_variant{1};

int launch_parse_blocks(const DecodeParams &p, unsigned images, cudaStream_t stream) {
    parse_blocks_kernel<<<images, 32, 0, stream>>>(p);
    count_launch();
    IE_CUDA(cudaGetLastError());
    return IE_OK;
}

int launch_decode_blocks(const DecodeParams &p, unsigned images, cudaStream_t stream) {
    if (!g_exact_transform.load()) {
        const unsigned nrange = (p.block_end ? p.block_end : p.nblocks) - p.block_base;
        dim3 gridf((nrange + 127) / 128, images);
        const bool lean_dec = g_decode_variant.load() == 1;
        if (p.N == 8 && lean_dec) IE_CUDA(launch_pdl(decode_blocks_lean_kernel<8>, gridf, dim3(128), 0, stream, p));
        else if (p.N == 4 && !p.add_mode && lean_dec) IE_CUDA(launch_pdl(decode_blocks_lean_kernel<4>, gridf, dim3(128), 0, stream, p));
        else if (p.N == 8) IE_CUDA(launch_pdl(decode_blocks_fast_kernel<8, false>, gridf, dim3(128), 0, stream, p));
        else if (p.N == 4 && p.add_mode) IE_CUDA(launch_pdl(decode_blocks_fast_kernel<4, true>, gridf, dim3(128), 0, stream, p));
        else if (p.N == 4) IE_CUDA(launch_pdl(decode_blocks_fast_kernel<4, false>, gridf, dim3(128), 0, stream, p));
        else { set_error("block size must be 4 or 8"); return IE_EINVAL; }
        count_launch();
        IE_CUDA(cudaGetLastError());
        return IE_OK;
    }
    dim3 grid(((p.block_end ? p.block_end : p.nblocks) - p.block_base + 255) / 256, images);
    if (p.N == 8) decode_blocks_kernel<8, false><<<grid, 256, 0, stream>>>(p);
    else if (p.N == 4 && p.add_mode) decode_blocks_kernel<4, true><<<grid, 256, 0, stream>>>(p);
    else if (p.N == 4) decode_blocks_kernel<4, false><<<grid, 256, 0, stream>>>(p);
    else { set_error("block size must be 4 or 8"); return IE_EINVAL; }
    count_launch();
    IE_CUDA(cudaGetLastError());
    return IE_OK;
}

}  // namespace ie
