// Image-block decode.
//   parse_blocks_kernel : finds every block's first bit.  The format has no markers and each block's size is in its own
//                         header (Block.cpp:443-444), so the offsets form a dependent chain (ImageDecoder.cpp:89-92
//                         "Reading raw must happen in sequence"); one lane per stream walks it.
//   decode_blocks_kernel: lane per block: read fields, sign-extend (utils.hpp:265-269), x Q, inverse DCT in the
//                         reference's summation order with exact zero-skipping, +128, clamp, truncate to u8
//                         (Block.cpp:441-472, 162-177, 99-107; algo.cpp:343-363).
#include "decode_image.cuh"
#include "transform.cuh"

namespace ie {

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
    unsigned long long *off = p.block_off + (size_t)img * (p.nblocks + 1);
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

template <int N, bool ADD>
__global__ void __launch_bounds__(256) decode_blocks_kernel(const DecodeParams p) {
    constexpr int NN = N * N;
    constexpr int STRIDE = NN + 2;
    __shared__ short s_coef[256 * STRIDE];
    const unsigned img = blockIdx.y;
    const unsigned gb = blockIdx.x * 256 + threadIdx.x;
    if (gb >= p.nblocks) return;
    const uint8_t *s = p.enc + (size_t)img * p.enc_stride;
    const unsigned long long total = p.enc_bits[img];
    const unsigned long long *off = p.block_off + (size_t)img * (p.nblocks + 1);
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
#pragma unroll
    for (int y = 0; y < N; y++) {
        unsigned lo = 0, hi = 0;
        uint8_t *row = dst + (size_t)(byi * N + y) * p.pitch + (size_t)bxi * N;
        unsigned cur_lo = 0;
        if (ADD) cur_lo = *reinterpret_cast<const unsigned *>(row);
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

int launch_parse_blocks(const DecodeParams &p, unsigned images, cudaStream_t stream) {
    parse_blocks_kernel<<<images, 32, 0, stream>>>(p);
    count_launch();
    IE_CUDA(cudaGetLastError());
    return IE_OK;
}

int launch_decode_blocks(const DecodeParams &p, unsigned images, cudaStream_t stream) {
    dim3 grid((p.nblocks + 255) / 256, images);
    if (p.N == 8) decode_blocks_kernel<8, false><<<grid, 256, 0, stream>>>(p);
    else if (p.N == 4 && p.add_mode) decode_blocks_kernel<4, true><<<grid, 256, 0, stream>>>(p);
    else if (p.N == 4) decode_blocks_kernel<4, false><<<grid, 256, 0, stream>>>(p);
    else { set_error("block size must be 4 or 8"); return IE_EINVAL; }
    count_launch();
    IE_CUDA(cudaGetLastError());
    return IE_OK;
}

}  // namespace ie
