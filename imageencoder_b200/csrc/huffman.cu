// Byte-wise Huffman stage (Huffman.cpp / Huffman.hpp of the reference) on a device-resident stream.
//
//   byte_hist_kernel      256-bin histogram, warp-aggregated, plus the position of every symbol's FIRST occurrence
//                         (Huffman.cpp:236-243: the reference's tie-breaking depends on insertion order)
//   host                  the 256-entry tree and dictionary, built with the same libstdc++ containers fed in the same
//                         order as the reference (Huffman.cpp:246-272) -- tie order is *defined* by those containers
//   huff_encode_kernel    per-byte code length -> CTA scan -> look-back -> chunk-centric pack after the dictionary
//                         header (Huffman.cpp:314-319), same append contract as the block packer (pack.cuh)
//   shift_copy_kernel     the "no extra compression" revert: '0' bit + the input bytes (Huffman.cpp:329-341)
//   huff_decode_kernel    table-driven decode (Huffman.cpp:190-204, 376-383)
#include <algorithm>
#include <cstring>
#include <memory_resource>
#include <queue>
#include <unordered_map>
#include <vector>

#include "api_internal.cuh"

namespace ie {

// 1 (default) = the round-2 kernels (span histogram, bits / scan / input-centric pack), 0 = the round-1 kernels (cross-check)
std::atomic<int> g_huffman_variant{1};

constexpr int kHuffTileBytes = kThreads * 16;

// ---------------------------------------------------------------------------------------------------------
// histogram + first occurrence
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) byte_hist_kernel(const uint8_t *__restrict__ in, size_t n, unsigned *hist,
                                                        unsigned long long *first_pos) {
    __shared__ unsigned s_hist[256];
    __shared__ unsigned long long s_first[256];
    s_hist[threadIdx.x] = 0;
    s_first[threadIdx.x] = ~0ull;
    __syncthreads();
    const size_t nvec = (n + 15) / 16;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec + 31; v += (size_t)gridDim.x * blockDim.x) {
        // keep whole warps in the loop so the match below is convergent
        const size_t wbase = v - (threadIdx.x & 31);
        if (wbase >= nvec) break;
        uint4 q = make_uint4(0, 0, 0, 0);
        int valid = 0;
        if (v < nvec) {
            const size_t b = v * 16;
            if (b + 16 <= n) { q = __ldg(reinterpret_cast<const uint4 *>(in) + v); valid = 16; }
            else {
                unsigned w[4] = {0, 0, 0, 0};
                valid = (int)(n - b);
                for (int i = 0; i < valid; i++) w[i >> 2] |= (unsigned)in[b + i] << (8 * (i & 3));
                q = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
        const unsigned words[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const unsigned byte = (words[i >> 2] >> (8 * (i & 3))) & 0xffu;
            const bool ok = i < valid;
            // warp aggregation: lanes holding the same byte value elect their lowest lane (= lowest position)
            const unsigned key = ok ? byte : 0x100u + (threadIdx.x & 31);
            const unsigned m = __match_any_sync(0xffffffffu, key);
            if (ok && (unsigned)(threadIdx.x & 31) == (unsigned)(__ffs(m) - 1)) {
                atomicAdd(&s_hist[byte], (unsigned)__popc(m));
                atomicMin(&s_first[byte], (unsigned long long)(v * 16 + i));
            }
        }
    }
    __syncthreads();
    if (s_hist[threadIdx.x]) {
        atomicAdd(&hist[threadIdx.x], s_hist[threadIdx.x]);
        atomicMin(&first_pos[threadIdx.x], s_first[threadIdx.x]);
    }
}

// ---------------------------------------------------------------------------------------------------------
// encode
// ---------------------------------------------------------------------------------------------------------
struct HuffCodes {
    unsigned word[256];
    unsigned char len[256];
};

struct ByteCodeTile {
    const unsigned char *bytes;      // shared: the tile's input bytes
    int nbytes;
    const unsigned *grp_off;         // shared: [kThreads + 1] exclusive bit offsets of the 16-byte groups
    const unsigned *word;            // shared [256]
    const unsigned char *len;        // shared [256]
};

__device__ __forceinline__ uint4 gather_chunk(const ByteCodeTile &t, long long ls) {
    unsigned ow0 = 0, ow1 = 0, ow2 = 0, ow3 = 0;
    int widx = 0, nacc = 0;
    unsigned long long acc = 0;
    if (ls < 0) {
        const int skip = (int)(-ls);
        widx = skip >> 5;
        nacc = skip & 31;
        ls = 0;
    }
    const unsigned total = t.grp_off[kThreads];
    if ((unsigned long long)ls < (unsigned long long)total) {
        int lo = 0, hi = kThreads;                       // largest g with grp_off[g] <= ls
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (t.grp_off[mid] <= (unsigned)ls) lo = mid; else hi = mid;
        }
        int i = lo * 16;
        unsigned pos = t.grp_off[lo];
        while (i < t.nbytes && pos + t.len[t.bytes[i]] <= (unsigned)ls) { pos += t.len[t.bytes[i]]; i++; }
        int fo = (int)((unsigned)ls - pos);
        for (; i < t.nbytes && widx < 4; i++) {
            const unsigned sym = t.bytes[i];
            int fw = t.len[sym];
            unsigned long long v = t.word[sym];
            if (fo) { fw -= fo; v &= (1ull << fw) - 1ull; fo = 0; }
            acc = (acc << fw) | v;
            nacc += fw;
            while (nacc >= 32 && widx < 4) {
                const unsigned word = (unsigned)(acc >> (nacc - 32));
                if (widx == 0) ow0 = word; else if (widx == 1) ow1 = word; else if (widx == 2) ow2 = word; else ow3 = word;
                widx++;
                nacc -= 32;
            }
        }
    }
    if (widx < 4 && nacc > 0) {
        const unsigned word = (unsigned)(acc << (32 - nacc));
        if (widx == 0) ow0 = word; else if (widx == 1) ow1 = word; else if (widx == 2) ow2 = word; else ow3 = word;
    }
    return make_uint4(__byte_perm(ow0, 0, 0x0123), __byte_perm(ow1, 0, 0x0123), __byte_perm(ow2, 0, 0x0123), __byte_perm(ow3, 0, 0x0123));
}

struct HuffEncodeParams {
    const uint8_t *in;
    size_t n;
    unsigned ntiles;
    const HuffCodes *codes;          // device
    uint8_t *out;
    size_t out_cap;
    unsigned long long *bit_counter;
    int *err;
    ScanState scan;
};

__global__ void __launch_bounds__(kThreads) huff_encode_kernel(const HuffEncodeParams p) {
    __shared__ __align__(16) unsigned char s_bytes[kHuffTileBytes];
    __shared__ unsigned s_grp[kThreads + 1];
    __shared__ unsigned s_word[256];
    __shared__ unsigned char s_len[256];
    __shared__ unsigned s_warp[kThreads / 32 + 1];
    __shared__ unsigned long long s_bcast;
    __shared__ unsigned s_tile;
    ScanState st = p.scan;
    if (threadIdx.x == 0) s_tile = atomicAdd(st.ticket, 1u);
    for (int i = threadIdx.x; i < 256; i += kThreads) { s_word[i] = p.codes->word[i]; s_len[i] = p.codes->len[i]; }
    __syncthreads();
    const unsigned tile = s_tile;
    const size_t base_byte = (size_t)tile * kHuffTileBytes;
    const int nbytes = (int)min((size_t)kHuffTileBytes, p.n - base_byte);

    // stage 16 bytes per thread, sum their code lengths
    unsigned bits = 0;
    {
        const int b0 = threadIdx.x * 16;
        uint4 q = make_uint4(0, 0, 0, 0);
        if (b0 + 16 <= nbytes && ((uintptr_t)(p.in + base_byte) % 16 == 0)) {
            q = __ldg(reinterpret_cast<const uint4 *>(p.in + base_byte) + threadIdx.x);
        } else if (b0 < nbytes) {
            unsigned w[4] = {0, 0, 0, 0};
            for (int i = 0; i < 16 && b0 + i < nbytes; i++) w[i >> 2] |= (unsigned)p.in[base_byte + b0 + i] << (8 * (i & 3));
            q = make_uint4(w[0], w[1], w[2], w[3]);
        }
        reinterpret_cast<uint4 *>(s_bytes)[threadIdx.x] = q;
        const unsigned words[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 16; i++)
            if (b0 + i < nbytes) bits += s_len[(words[i >> 2] >> (8 * (i & 3))) & 0xffu];
    }
    unsigned T;
    const unsigned excl = cta_exclusive_scan(bits, s_warp, &T);
    s_grp[threadIdx.x] = excl;
    if (threadIdx.x == kThreads - 1) s_grp[kThreads] = T;
    unsigned long long base = 0;
    if (tile == 0) base = *p.bit_counter;
    const unsigned long long G = tile_lookback(st, tile, (unsigned long long)T + base, &s_bcast) + (tile == 0 ? base : 0ull);
    const bool last_tile = (tile == p.ntiles - 1);
    ByteCodeTile bt;
    bt.bytes = s_bytes; bt.nbytes = nbytes; bt.grp_off = s_grp; bt.word = s_word; bt.len = s_len;
    tile_write_chunks(bt, st, tile, tile == 0, last_tile, G, T, p.out, p.out_cap, p.err);
    if (last_tile && threadIdx.x == 0) {
        *p.bit_counter = G + T;
        atomicExch(st.ticket, 0u);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Round 2: the Huffman stage as bandwidth-shaped kernels (the kernels above took 128 + 188 us on the 27.7 MB stream of
// config 2; they stay as the cross-check: ie_set_option("huffman_variant", 0)).
//   byte_hist_span_kernel  every CTA counts one contiguous span of the input in per-warp shared-memory histograms (plain
//                          shared atomics: the stream's bytes are close to uniform) and notes, per symbol, the first CTA
//                          that saw it
//   byte_first_kernel      a warp per symbol scans that CTA's span for the symbol's first position -- the reference's
//                          tie-breaking depends on first-occurrence order (Huffman.cpp:236-243)
//   huff_bits_kernel       code bits per 4 KiB tile              (the dictionary is known by now: built on the host)
//   huff_scan_kernel       one CTA: exclusive scan of the tile totals -> every tile's first bit, the stream's length
//   huff_pack_kernel       input-centric: a thread concatenates the codes of its 16 bytes in registers and writes whole words
//                          into a shared-memory image of the tile's bits (shared atomicOr only where two threads meet); the
//                          image goes out through the same chunk writer as the block encoder.  No look-back: the offsets
//                          are known before the kernel starts.
// ---------------------------------------------------------------------------------------------------------
constexpr int kHist2Threads = 256;
__global__ void __launch_bounds__(kHist2Threads) byte_hist_span_kernel(const uint8_t *__restrict__ in, size_t n, size_t span, unsigned *hist,
                                                                     unsigned *first_cta) {
    __shared__ unsigned s_h[kHist2Threads / 32][256];
    for (int i = threadIdx.x; i < (kHist2Threads / 32) * 256; i += kHist2Threads) (&s_h[0][0])[i] = 0;
    __syncthreads();
    const size_t b0 = (size_t)blockIdx.x * span, b1 = min(n, b0 + span);           // span is a multiple of 16
    unsigned *h = s_h[threadIdx.x >> 5];
    for (size_t b = b0 + (size_t)threadIdx.x * 16; b < b1; b += (size_t)kHist2Threads * 16) {
        if (b + 16 <= b1) {
            const uint4 q = __ldg(reinterpret_cast<const uint4 *>(in + b));
            const unsigned words[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < 16; i++) atomicAdd(&h[(words[i >> 2] >> (8 * (i & 3))) & 0xffu], 1u);
        } else {
            for (size_t i = b; i < b1; i++) atomicAdd(&h[in[i]], 1u);
        }
    }
    __syncthreads();
    unsigned c = 0;
#pragma unroll
    for (int w = 0; w < kHist2Threads / 32; w++) c += s_h[w][threadIdx.x];
    if (c) {
        atomicAdd(&hist[threadIdx.x], c);
        atomicMin(&first_cta[threadIdx.x], blockIdx.x);
    }
}

// first_pos[sym] = position of the symbol's first occurrence: it lies in the span of the first CTA that counted it
__global__ void __launch_bounds__(256) byte_first_kernel(const uint8_t *__restrict__ in, size_t n, size_t span, const unsigned *first_cta,
                                                         unsigned long long *first_pos) {
    const unsigned sym = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (sym >= 256) return;
    const unsigned cta = first_cta[sym];
    if (cta == 0xFFFFFFFFu) { if (lane == 0) first_pos[sym] = ~0ull; return; }
    const size_t b0 = (size_t)cta * span, b1 = min(n, b0 + span);
    // 2 KiB of the span per iteration: lane l looks at bytes [64 l, 64 l + 64) of it
    for (size_t wb = b0; wb < b1; wb += 2048) {
        const size_t b = wb + (size_t)lane * 64;
        unsigned hit = 64;
        if (b + 64 <= b1) {
            uint4 q[4];
#pragma unroll
            for (int k = 0; k < 4; k++) q[k] = __ldg(reinterpret_cast<const uint4 *>(in + b) + k);
#pragma unroll
            for (int k = 3; k >= 0; k--) {
                const unsigned words[4] = {q[k].x, q[k].y, q[k].z, q[k].w};
#pragma unroll
                for (int i = 15; i >= 0; i--) if (((words[i >> 2] >> (8 * (i & 3))) & 0xffu) == sym) hit = 16 * k + i;
            }
        } else {
            for (size_t i = b; i < b1; i++) if (in[i] == sym) { hit = (unsigned)(i - b); break; }
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit < 64);
        if (m) {
            const int src = __ffs((int)m) - 1;
            const unsigned h = __shfl_sync(0xffffffffu, hit, src);
            if (lane == 0) first_pos[sym] = (unsigned long long)(wb + (size_t)src * 64 + h);
            return;
        }
    }
    if (lane == 0) first_pos[sym] = ~0ull;                 // cannot happen: the CTA counted the symbol
}

constexpr int kPackThreads = 256;
constexpr int kPackTileBytes = kPackThreads * 16;                           // 4 KiB of input per CTA
constexpr int kPackImageWords = kPackTileBytes;                            // worst case: 32-bit codes -> 32 bits per input byte

struct HuffPackParams {
    const uint8_t *in;
    size_t n;
    unsigned ntiles;
    const HuffCodes *codes;
    unsigned *tile_bits;                 // [ntiles]
    unsigned long long *tile_off;        // [ntiles] first bit of the tile in the stream
    uint8_t *out;
    size_t out_cap;
    unsigned long long *bit_counter;     // in: first free bit of the stream; out: one past the last bit
    int *err;
    ScanState scan;                      // tile-boundary hand-off records
};

__device__ __forceinline__ uint4 load_tile_bytes(const uint8_t *in, size_t base_byte, int nbytes, int b0) {
    uint4 q = make_uint4(0, 0, 0, 0);
    if (b0 + 16 <= nbytes && ((uintptr_t)(in + base_byte) % 16 == 0)) {
        q = __ldg(reinterpret_cast<const uint4 *>(in + base_byte) + (b0 >> 4));
    } else if (b0 < nbytes) {
        unsigned w[4] = {0, 0, 0, 0};
        for (int i = 0; i < 16 && b0 + i < nbytes; i++) w[i >> 2] |= (unsigned)in[base_byte + b0 + i] << (8 * (i & 3));
        q = make_uint4(w[0], w[1], w[2], w[3]);
    }
    return q;
}

__global__ void __launch_bounds__(kPackThreads) huff_bits_kernel(const HuffPackParams p) {
    __shared__ unsigned char s_len[256];
    __shared__ unsigned s_part[kPackThreads / 32];
    s_len[threadIdx.x] = p.codes->len[threadIdx.x];
    __syncthreads();
    for (unsigned tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const size_t base_byte = (size_t)tile * kPackTileBytes;
        const int nbytes = (int)min((size_t)kPackTileBytes, p.n - base_byte);
        const int b0 = threadIdx.x * 16;
        const uint4 q = load_tile_bytes(p.in, base_byte, nbytes, b0);
        const unsigned words[4] = {q.x, q.y, q.z, q.w};
        unsigned bits = 0;
#pragma unroll
        for (int i = 0; i < 16; i++)
            if (b0 + i < nbytes) bits += s_len[(words[i >> 2] >> (8 * (i & 3))) & 0xffu];
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) bits += __shfl_xor_sync(0xffffffffu, bits, d);
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = bits;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned t = 0;
            for (int w = 0; w < kPackThreads / 32; w++) t += s_part[w];
            p.tile_bits[tile] = t;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024) huff_scan_kernel(const HuffPackParams p) {
    __shared__ unsigned long long s_sum[1024];
    const unsigned per = (p.ntiles + 1023) / 1024;
    const unsigned t0 = threadIdx.x * per, t1 = min(t0 + per, p.ntiles);
    unsigned long long sum = 0;
    for (unsigned t = t0; t < t1; t++) sum += p.tile_bits[t];
    s_sum[threadIdx.x] = sum;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        const unsigned long long v = (threadIdx.x >= (unsigned)d) ? s_sum[threadIdx.x - d] : 0ull;
        __syncthreads();
        s_sum[threadIdx.x] += v;
        __syncthreads();
    }
    const unsigned long long start = *p.bit_counter;
    unsigned long long base = start + s_sum[threadIdx.x] - sum;
    for (unsigned t = t0; t < t1; t++) { p.tile_off[t] = base; base += p.tile_bits[t]; }
    __syncthreads();
    if (threadIdx.x == 1023) *p.bit_counter = start + s_sum[1023];
}

__global__ void __launch_bounds__(kPackThreads) huff_pack_kernel(const HuffPackParams p) {
    __shared__ __align__(16) unsigned s_img[kPackImageWords + 8];
    __shared__ unsigned s_word[256];
    __shared__ unsigned char s_len[256];
    __shared__ unsigned s_warp[kPackThreads / 32];
    s_word[threadIdx.x] = p.codes->word[threadIdx.x];
    s_len[threadIdx.x] = p.codes->len[threadIdx.x];
    const unsigned tile = blockIdx.x;
    const size_t base_byte = (size_t)tile * kPackTileBytes;
    const int nbytes = (int)min((size_t)kPackTileBytes, p.n - base_byte);
    const int b0 = threadIdx.x * 16;
    const uint4 q = load_tile_bytes(p.in, base_byte, nbytes, b0);
    const unsigned words[4] = {q.x, q.y, q.z, q.w};
    const unsigned T = p.tile_bits[tile];
    const unsigned nw = (T + 31) / 32;
    for (unsigned i = threadIdx.x; i < (nw + 3) / 4; i += kPackThreads) reinterpret_cast<uint4 *>(s_img)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    // this thread's bits and where they start inside the tile
    unsigned bits = 0;
#pragma unroll
    for (int i = 0; i < 16; i++)
        if (b0 + i < nbytes) bits += s_len[(words[i >> 2] >> (8 * (i & 3))) & 0xffu];
    unsigned inc = bits;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    unsigned pos = inc - bits;
    for (int w = 0; w < wid; w++) pos += s_warp[w];
    // concatenate the codes (<= 32 bits each, Huffman.cpp:86-88) and write whole words; the first and the last word may be
    // shared with the neighbouring threads
    unsigned *ow = s_img + (pos >> 5);
    unsigned long long acc = 0;
    int nacc = (int)(pos & 31u);
    bool first = nacc != 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        if (b0 + i < nbytes) {
            const unsigned sym = (words[i >> 2] >> (8 * (i & 3))) & 0xffu;
            const int l = s_len[sym];
            acc = (acc << l) | (unsigned long long)s_word[sym];
            nacc += l;
            if (nacc >= 32) {
                nacc -= 32;
                const unsigned w = (unsigned)(acc >> nacc);
                if (first) { atomicOr(ow, w); first = false; } else *ow = w;
                ow++;
            }
        }
    }
    if (nacc > 0) atomicOr(ow, (unsigned)(acc << (32 - nacc)));
    __syncthreads();
    SmemStreamTile t;
    t.words = s_img; t.nwords = nw;
    tile_write_chunks_n<kPackThreads>(t, p.scan, tile, tile == 0, tile + 1 == p.ntiles, p.tile_off[tile], T, p.out, p.out_cap, p.err);
}

// out bits [shift, shift + 8 n) = in bytes; out bits [0, shift) = 0; whole 32-bit words, zero padded.
__global__ void shift_copy_kernel(const uint8_t *__restrict__ in, size_t n, uint8_t *out, unsigned shift) {
    const size_t nwords = (n * 8 + shift + 31) / 32;
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += (size_t)gridDim.x * blockDim.x) {
        // out word w covers stream bits [32w, 32w+32) = in bits [32w - shift, ...)
        unsigned long long window = 0;     // in bytes 4w-1 .. 4w+3 (5 bytes), MSB first
        for (int i = -1; i < 4; i++) {
            const long long bi = (long long)w * 4 + i;
            const unsigned byte = (bi >= 0 && (size_t)bi < n) ? in[bi] : 0u;
            window = (window << 8) | byte;
        }
        // window (40 bits) starts at in bit 32w - 8; skip 8 - shift bits, keep 32, drop the low `shift` bits
        const unsigned v = (unsigned)((window >> shift) & 0xffffffffull);
        reinterpret_cast<unsigned *>(out)[w] = __byte_perm(v, 0, 0x0123);
    }
}

// Multi-GPU stitch helper: re-aligns a shard's stream to the chunk grid of the global stream.
// params[0] = nbits of the shard stream (it starts at bit 0 of `in`), params[1] = the shard's global bit offset.
// out bit (params[1] % 128 + i) = in bit i; everything else in the written chunks is 0, so neighbouring shards can be
// OR-merged at their one shared chunk.  Thread per 32-bit output word.
__global__ void stream_shift_kernel(const uint8_t *__restrict__ in, const unsigned long long *__restrict__ params, uint8_t *out,
                                    size_t out_cap) {
    const unsigned long long nbits = params[0];
    const unsigned shift = (unsigned)(params[1] % kChunkBits);
    const unsigned long long nwords = ((nbits + shift + kChunkBits - 1) / kChunkBits) * 4;
    for (unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w < nwords;
         w += (unsigned long long)gridDim.x * blockDim.x) {
        if ((w + 1) * 4 > out_cap) return;
        const long long b0 = (long long)(w * 32) - (long long)shift;          // first input bit of this word
        unsigned v = 0;
        if (b0 + 32 > 0 && b0 < (long long)nbits) {
            // 5 input bytes starting at floor(b0 / 8)
            const long long by = (b0 >= 0) ? (b0 >> 3) : -((-b0 + 7) >> 3);
            const unsigned bo = (unsigned)(b0 - by * 8);                       // 0..7
            unsigned long long win = 0;
            for (int i = 0; i < 5; i++) {
                const long long bi = by + i;
                const unsigned byte = (bi >= 0 && (unsigned long long)bi * 8 < nbits) ? in[bi] : 0u;
                win = (win << 8) | byte;
            }
            v = (unsigned)((win >> (8 - bo)) & 0xffffffffull);
            // clear bits at or beyond nbits
            const long long over = b0 + 32 - (long long)nbits;
            if (over > 0) v &= (over >= 32) ? 0u : (0xffffffffu << over);
            if (b0 < 0) v &= (0xffffffffu >> (unsigned)(-b0));
        }
        reinterpret_cast<unsigned *>(out)[w] = __byte_perm(v, 0, 0x0123);
    }
}

// ---------------------------------------------------------------------------------------------------------
// decode (one lane; 12-bit primary lookup table, tree walk for longer codes)
// ---------------------------------------------------------------------------------------------------------
struct HuffDecodeTables {
    // entry: (len << 16) | symbol for codes <= 12 bits; 0xFFFF0000 | node index for longer codes; 0 = invalid
    unsigned lut[4096];
    short child[512][2];       // tree for the slow path: >= 0 node index, -(sym+1)-1 ... see host builder
    int nodes;
};

__global__ void huff_decode_kernel(const uint8_t *__restrict__ in, size_t n, unsigned long long start_bit,
                                   const HuffDecodeTables *tab, uint8_t *out, size_t out_cap, unsigned long long *out_count, int *err,
                                   const unsigned *spec_ok) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (spec_ok && *spec_ok) return;                 // the parallel decode verified: nothing to do
    const unsigned long long total = (unsigned long long)n * 8ull;
    unsigned long long pos = start_bit;
    size_t o = 0;
    while (pos < total) {
        // 32-bit window at pos (zero beyond the end: BitStream.cpp:17-20)
        const unsigned long long b = pos >> 3;
        unsigned long long win = 0;
#pragma unroll
        for (int i = 0; i < 6; i++) win = (win << 8) | ((b + i < n) ? (unsigned long long)__ldg(in + b + i) : 0ull);
        const unsigned w32 = (unsigned)(win >> (16 - (pos & 7)));     // 32 bits starting at pos
        unsigned e = tab->lut[w32 >> 20];
        unsigned sym, len;
        if ((e >> 16) != 0xFFFFu) {
            if (e == 0) { atomicExch(err, IE_EFORMAT); break; }
            sym = e & 0xffffu;
            len = e >> 16;
        } else {
            int node = (int)(e & 0xffffu);
            len = 12;
            sym = 0;
            bool found = false;
            while (len < 32) {
                const int bit = (w32 >> (31 - len)) & 1;
                const int c = tab->child[node][bit];
                len++;
                if (c == -1) { break; }
                if (c <= -2) { sym = (unsigned)(-c - 2); found = true; break; }
                node = c;
            }
            if (!found) { atomicExch(err, IE_EFORMAT); break; }
        }
        if (o >= out_cap) { atomicExch(err, IE_ENOSPC); break; }
        out[o++] = (uint8_t)sym;
        pos = min(pos + len, total);              // a final code may run past the end: zero bits, no advance
    }
    *out_count = o;
}

// ---------------------------------------------------------------------------------------------------------
// Parallel decode: speculate + verify (same scheme as the block parser, parse.cu).  The code stream is cut into groups of
// kHuffGroupBits bits.  Huffman codes re-synchronise quickly, so every group starts kHuffLead bits early at an arbitrary
// bit and ASSUMES it is on a codeword boundary when it reaches its first bit; the assumption is then verified exactly
// (group g's entry offset must equal group g-1's exit offset; group 0 starts at the true first code).  Mismatching groups
// are re-walked from their predecessor's exit for a few rounds; if any mismatch remains, spec_ok stays 0 and the serial
// kernel above decodes instead.  Symbol counts per group are scanned to place every group's symbols in the output.
// ---------------------------------------------------------------------------------------------------------
constexpr int kHuffGroupBits = 8192;
constexpr int kHuffLead = 2048;
constexpr int kHuffRounds = 3;
constexpr unsigned kHuffDead = 0xFFFFFFFFu;

struct HuffParParams {
    const uint8_t *in;                 // 4-byte aligned
    unsigned long long total;          // bits
    unsigned long long start_bit;      // first code
    const HuffDecodeTables *tab;
    unsigned ngroups;
    uint2 *entry;                      // [ngroups] (entry offset, first symbol index)
    uint2 *exit_;                      // [ngroups] (exit offset | dead, symbols started in the group)
    unsigned *flags;                   // [1] spec_ok
    uint8_t *out;
    size_t out_cap;
    unsigned long long *out_count;
    int *err;
};

// length of the code at absolute bit p (p < total) and its symbol; len 0 = no such code (garbage phase or malformed)
__device__ __forceinline__ unsigned huff_code_at(const HuffParParams &p, unsigned long long pos, unsigned &sym) {
    const unsigned *wp = reinterpret_cast<const unsigned *>(p.in) + (pos >> 5);
    const unsigned long long nwords = (p.total + 31) >> 5;
    const unsigned w0 = __byte_perm(__ldg(wp), 0, 0x0123);
    const unsigned w1 = ((pos >> 5) + 1 < nwords) ? __byte_perm(__ldg(wp + 1), 0, 0x0123) : 0u;
    unsigned w32 = __funnelshift_l(w1, w0, (unsigned)(pos & 31));
    if (pos + 32 > p.total) w32 &= ~((p.total - pos >= 32) ? 0u : (0xFFFFFFFFu >> (unsigned)(p.total - pos)));   // zero bits past the end
    const unsigned e = __ldg(&p.tab->lut[w32 >> 20]);
    if ((e >> 16) != 0xFFFFu) { sym = e & 0xffffu; return e >> 16; }        // e == 0 -> len 0
    int node = (int)(e & 0xffffu);
    unsigned len = 12;
    while (len < 32) {
        const int c = p.tab->child[node][(w32 >> (31 - len)) & 1];
        len++;
        if (c == -1) return 0u;
        if (c <= -2) { sym = (unsigned)(-c - 2); return len; }
        node = c;
    }
    return 0u;
}

// decodes from absolute bit pos to the end of the group; optionally writes the symbols
__device__ __forceinline__ uint2 huff_walk_group(const HuffParParams &p, unsigned long long pos, unsigned long long g_end, uint8_t *dst, size_t dst_cap) {
    unsigned cnt = 0;
    while (true) {
        if (pos >= p.total) return make_uint2(kHuffDead, cnt);
        if (pos >= g_end) return make_uint2((unsigned)(pos - g_end), cnt);
        unsigned sym = 0;
        const unsigned len = huff_code_at(p, pos, sym);
        if (len == 0) return make_uint2(kHuffDead - 1, cnt);               // malformed on this phase
        if (dst) { if (cnt < dst_cap) dst[cnt] = (uint8_t)sym; }
        cnt++;
        pos += len;                                                        // a final code may run past the end (zero bits)
    }
}

__global__ void __launch_bounds__(64) huff_spec_walk(const HuffParParams p) {
    const unsigned g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g == 0) p.flags[0] = 0;
    if (g >= p.ngroups) return;
    const unsigned long long g_start = p.start_bit + (unsigned long long)g * kHuffGroupBits, g_end = g_start + kHuffGroupBits;
    unsigned entry;
    unsigned long long pos = (g == 0) ? p.start_bit : g_start - kHuffLead;
    if (g_start >= p.total) entry = kHuffDead;
    else {
        while (pos < g_start) {
            unsigned sym;
            const unsigned len = huff_code_at(p, pos, sym);
            pos += len ? len : 1u;
        }
        entry = (pos >= p.total) ? kHuffDead : (unsigned)(pos - g_start);
    }
    p.entry[g] = make_uint2(entry, 0u);
    p.exit_[g] = (entry == kHuffDead) ? make_uint2(kHuffDead, 0u) : huff_walk_group(p, pos, g_end, nullptr, 0);
}

__global__ void __launch_bounds__(64) huff_spec_repair(const HuffParParams p) {
    const unsigned g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p.ngroups || g == 0) return;
    if (g >= 2 && p.exit_[g - 2].x != p.entry[g - 1].x) return;            // only adopt the exit of a consistent predecessor
    const unsigned want = p.exit_[g - 1].x;
    if (want == p.entry[g].x) return;
    const unsigned long long g_start = p.start_bit + (unsigned long long)g * kHuffGroupBits, g_end = g_start + kHuffGroupBits;
    p.entry[g] = make_uint2(want, 0u);
    p.exit_[g] = (want >= kHuffDead - 1) ? make_uint2(kHuffDead, 0u) : huff_walk_group(p, g_start + want, g_end, nullptr, 0);
}

__global__ void __launch_bounds__(1024) huff_spec_finish(const HuffParParams p) {
    __shared__ unsigned s_bad;
    __shared__ unsigned long long s_sum[1024];
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    const unsigned per = (p.ngroups + 1023) / 1024;
    const unsigned g0 = threadIdx.x * per, g1 = min(g0 + per, p.ngroups);
    unsigned long long sum = 0;
    unsigned bad = 0;
    for (unsigned g = g0; g < g1; g++) {
        if (g > 0 && p.exit_[g - 1].x != p.entry[g].x) bad = 1;
        if (p.exit_[g].x == kHuffDead - 1) bad = 1;                        // malformed code on the (supposedly) true chain
        sum += p.exit_[g].y;
    }
    if (bad) atomicOr(&s_bad, 1u);
    s_sum[threadIdx.x] = sum;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        const unsigned long long v = (threadIdx.x >= (unsigned)d) ? s_sum[threadIdx.x - d] : 0ull;
        __syncthreads();
        s_sum[threadIdx.x] += v;
        __syncthreads();
    }
    if (s_bad) return;                                                    // spec_ok stays 0: the serial kernel decodes
    unsigned long long base = s_sum[threadIdx.x] - sum;
    for (unsigned g = g0; g < g1; g++) {
        p.entry[g].y = (unsigned)base;                                     // first symbol index of the group (< 2^32 checked by the host)
        base += p.exit_[g].y;
    }
    if (threadIdx.x == 1023) *p.out_count = s_sum[1023];
    if (threadIdx.x == 0) p.flags[0] = 1;
}

__global__ void __launch_bounds__(64) huff_spec_emit(const HuffParParams p) {
    const unsigned g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p.ngroups || !p.flags[0]) return;
    const uint2 en = p.entry[g];
    if (en.x == kHuffDead) return;
    const unsigned long long g_start = p.start_bit + (unsigned long long)g * kHuffGroupBits, g_end = g_start + kHuffGroupBits;
    if ((size_t)en.y + p.exit_[g].y > p.out_cap) { atomicExch(p.err, IE_ENOSPC); return; }
    huff_walk_group(p, g_start + en.x, g_end, p.out + en.y, p.out_cap - en.y);
}

// ---------------------------------------------------------------------------------------------------------
// host side: tree / dictionary exactly as the reference builds them
// ---------------------------------------------------------------------------------------------------------
namespace {
struct HNode { uint8_t data; size_t freq; HNode *left, *right; };
struct HNodeCmp { bool operator()(const HNode *a, const HNode *b) const { return a->freq > b->freq; } };   // Huffman.hpp:70-74
struct HCode { uint32_t word, len; };
using HDict = std::pmr::unordered_map<uint8_t, HCode>;

void walk(const HNode *n, uint32_t word, uint32_t len, HDict &dict) {                                    // Huffman.cpp:79-104
    if (!n) return;
    if (!n->left && !n->right) { dict[n->data] = HCode{word, len}; return; }
    walk(n->left, word << 1, len + 1, dict);
    walk(n->right, (word << 1) | 1u, len + 1, dict);
}

struct HostBitWriter {
    std::vector<uint8_t> buf;
    size_t pos = 0;
    void put(unsigned len, uint32_t v) {                        // BitStream.cpp:73-77, MSB first; up to a byte per step
        if (((pos + len + 7) >> 3) > buf.size()) buf.resize(((pos + len + 7) >> 3) + 1024, 0);
        while (len) {
            const unsigned room = 8u - (unsigned)(pos & 7), take = len < room ? len : room;
            const unsigned bits = (v >> (len - take)) & ((1u << take) - 1u);
            buf[pos >> 3] |= uint8_t(bits << (room - take));
            pos += take;
            len -= take;
        }
    }
};
}  // namespace

// hist/first -> codes + dictionary header bits.  Returns IE_EINVAL when a code would exceed 32 bits (the reference's
// uint32 code words overflow there, Huffman.cpp:86-88).
// The reference's tie-breaking is whatever libstdc++'s unordered_map iteration order, priority_queue heap order and (unstable)
// std::sort produce when fed in its order, so the containers and the sequence of operations on them are exactly
// the reference's; only their storage comes from one stack arena (polymorphic allocators change neither hashing nor bucket
// growth) and the tree nodes from a fixed pool -- 114 -> 26 us per call on the bench stream's histogram (tests/host/huffdict_check.cu compares
// this function with the plain-allocator transcription it replaced on random histograms).
static int build_dictionary(const unsigned *hist, const unsigned long long *first, HuffCodes &codes, HostBitWriter &hdr) {
    alignas(16) unsigned char arena[64 * 1024];
    std::pmr::monotonic_buffer_resource pool(arena, sizeof arena);       // spills to the heap if ever exhausted
    std::pmr::vector<int> syms(&pool);
    syms.reserve(256);
    for (int i = 0; i < 256; i++) if (hist[i]) syms.push_back(i);
    if (syms.empty()) { set_error("empty input for the Huffman stage"); return IE_EINVAL; }
    std::sort(syms.begin(), syms.end(), [&](int a, int b) { return first[a] < first[b]; });
    std::pmr::unordered_map<uint8_t, uint32_t> freqs(&pool);               // Huffman.cpp:237-243 (insertion = first occurrence)
    for (int s : syms) freqs[(uint8_t)s] = hist[s];
    HNode nodes[511];                                                      // 256 leaves + 255 inner nodes at most
    int nn = 0;
    std::priority_queue<HNode *, std::pmr::vector<HNode *>, HNodeCmp> pq{HNodeCmp(), std::pmr::vector<HNode *>(&pool)};   // Huffman.cpp:246-251
    for (const auto &pr : freqs) { nodes[nn] = HNode{pr.first, pr.second, nullptr, nullptr}; pq.push(&nodes[nn++]); }
    while (pq.size() > 1) {                                                // Huffman.cpp:253-260
        HNode *l = pq.top(); pq.pop();
        HNode *r = pq.top(); pq.pop();
        nodes[nn] = HNode{0xFF, l->freq + r->freq, l, r};
        pq.push(&nodes[nn++]);
    }
    HNode *root = pq.top();
    HDict dict(&pool);
    walk(root, 0, 0, dict);                                                // Huffman.cpp:266
    std::pmr::vector<std::pair<uint8_t, HCode>> sorted(dict.begin(), dict.end(), &pool);   // Huffman.cpp:269
    std::sort(sorted.begin(), sorted.end(),                                 // Huffman.cpp:272 (unstable, len descending)
              [](const std::pair<uint8_t, HCode> &a, const std::pair<uint8_t, HCode> &b) { return a.second.len > b.second.len; });
    std::pmr::unordered_map<uint32_t, uint32_t> bit_freqs(&pool);
    for (const auto &w : sorted) bit_freqs[w.second.len]++;
    uint32_t seq_len = 0, bit_len = 0;
    for (const auto &w : sorted) {                                         // Huffman.cpp:298-309
        if (w.second.len > 32) { set_error("Huffman code longer than 32 bits (undefined in the reference)"); return IE_EINVAL; }
        if (seq_len == 0) {
            bit_len = w.second.len;
            seq_len = bit_freqs[bit_len];
            hdr.put(8, 0x80u | (seq_len & 0x7Fu));                         // Huffman.cpp:36-46
            hdr.put(4, bit_len & 0xFu);
        }
        hdr.put(8, w.first);
        hdr.put(bit_len, w.second.word);
        seq_len--;
    }
    hdr.put(1, 0);                                                         // Huffman.cpp:311
    memset(&codes, 0, sizeof codes);
    for (const auto &pr : dict) { codes.word[pr.first] = pr.second.word; codes.len[pr.first] = (unsigned char)pr.second.len; }
    return IE_OK;
}

static int ensure_scratch(ie_session *s, size_t need) { return session_reserve(&s->d_scratch, &s->scratch_cap, need); }

// hist[256] (u32) and first[256] (u64) of n bytes, on the device; d_first_cta: 256 u32 of scratch (variant 1 only)
static int launch_byte_histogram(const uint8_t *d_in, size_t n, unsigned *d_hist, unsigned long long *d_first, unsigned *d_first_cta,
                                 int sm_count, cudaStream_t st) {
    IE_CUDA(cudaMemsetAsync(d_hist, 0, 256 * 4, st));
    if (g_huffman_variant.load() == 0 || !d_first_cta || ((uintptr_t)d_in % 16)) {
        IE_CUDA(cudaMemsetAsync(d_first, 0xff, 256 * 8, st));
        const int grid = (int)std::min<size_t>((n / 16 + 255) / 256 + 1, (size_t)sm_count * 8);
        byte_hist_kernel<<<grid, 256, 0, st>>>(d_in, n, d_hist, d_first);
        count_launch();
    } else {
        IE_CUDA(cudaMemsetAsync(d_first_cta, 0xff, 256 * 4, st));
        const size_t want = (size_t)sm_count * 8;
        size_t span = ((n + want - 1) / want + 15) / 16 * 16;
        span = std::max<size_t>(span, 4096);
        const unsigned grid = (unsigned)((n + span - 1) / span);
        byte_hist_span_kernel<<<grid, kHist2Threads, 0, st>>>(d_in, n, span, d_hist, d_first_cta);
        byte_first_kernel<<<32, 256, 0, st>>>(d_in, n, span, d_first_cta, d_first);
        count_launch(2);
    }
    IE_CUDA(cudaGetLastError());
    return IE_OK;
}

// pinned staging of the stage: [0, 2048) dictionary header | HuffCodes | first bit (u64) | pad | first[256] u64 + hist[256] u32 as
// they come back from the device | stream bits (u64) + error flag
constexpr size_t kHuffPinnedRecv = (2048 + sizeof(HuffCodes) + 16 + 63) / 64 * 64;
constexpr size_t kHuffPinned = kHuffPinnedRecv + 256 * 8 + 256 * 4 + 64;      // ... + status of the host callback (at + 16)
static int ensure_huff_pinned(ie_session *s) {
    if (!s->h_huff) IE_CUDA(cudaMallocHost(&s->h_huff, kHuffPinned));
    return IE_OK;
}

// Dictionary from a (global) histogram, then the scan-pack of `n` bytes behind it (write_dict) or from bit 0 (a later shard of
// a multi-GPU stream).  Leaves the stream's bit count in s->d_counter[0].  The scratch must be ensured by the caller.
static int huffman_pack_dev(ie_session *s, const uint8_t *d_in, size_t n, const unsigned *hist, const unsigned long long *first,
                            int write_dict, uint8_t *d_out, size_t out_cap, cudaStream_t st) {
    HuffCodes *d_codes = reinterpret_cast<HuffCodes *>(s->d_scratch + 256 * 8 + 256 * 4);
    // dictionary header, codes and the stream's first bit travel from pinned memory the session owns: the copies are truly
    // asynchronous and nothing on this frame has to outlive them (no stream synchronisation in here)
    IE_TRY(ensure_huff_pinned(s));
    HuffCodes &codes = *reinterpret_cast<HuffCodes *>(s->h_huff + 2048);
    unsigned long long &hb = *reinterpret_cast<unsigned long long *>(s->h_huff + 2048 + sizeof(HuffCodes));
    HostBitWriter hdr;
    IE_TRY(build_dictionary(hist, first, codes, hdr));
    if (!write_dict) { hdr.pos = 0; hdr.buf.assign(16, 0); }
    const size_t hdr_bytes16 = std::max<size_t>(16, (hdr.pos + 127) / 128 * 16);
    hdr.buf.resize(hdr_bytes16, 0);
    if (hdr_bytes16 > 2048) { set_error("Huffman dictionary header larger than 2 KiB"); return IE_EINVAL; }
    if (out_cap < hdr_bytes16 + 16) { set_error("output buffer too small"); return IE_ENOSPC; }
    memcpy(s->h_huff, hdr.buf.data(), hdr_bytes16);
    IE_CUDA(cudaMemcpyAsync(d_out, s->h_huff, hdr_bytes16, cudaMemcpyHostToDevice, st));
    IE_CUDA(cudaMemcpyAsync(d_codes, &codes, sizeof codes, cudaMemcpyHostToDevice, st));
    const bool v0 = g_huffman_variant.load() == 0;
    // (the scan arrays -- the stream's bit counter among them -- may be re-allocated here: before the counter is set)
    IE_TRY(session_ensure_scan(s, 1, (unsigned)((n + (v0 ? kHuffTileBytes : kPackTileBytes) - 1) / (v0 ? kHuffTileBytes : kPackTileBytes))));
    hb = hdr.pos;
    IE_CUDA(cudaMemcpyAsync(s->d_counter, &hb, sizeof hb, cudaMemcpyHostToDevice, st));
    if (v0) {
        const unsigned ntiles = (unsigned)((n + kHuffTileBytes - 1) / kHuffTileBytes);
        HuffEncodeParams p;
        p.in = d_in; p.n = n; p.ntiles = ntiles; p.codes = d_codes; p.out = d_out; p.out_cap = out_cap;
        p.bit_counter = s->d_counter; p.err = s->d_err; p.scan = s->scan_state();
        huff_encode_kernel<<<ntiles, kThreads, 0, st>>>(p);
        count_launch();
    } else {
        const unsigned ntiles = (unsigned)((n + kPackTileBytes - 1) / kPackTileBytes);
        IE_TRY(session_reserve(&s->d_tile_meta, &s->tile_meta_cap, (size_t)ntiles * (sizeof(unsigned long long) + sizeof(unsigned)) + 64));
        HuffPackParams p;
        p.in = d_in; p.n = n; p.ntiles = ntiles; p.codes = d_codes; p.out = d_out; p.out_cap = out_cap;
        p.tile_off = reinterpret_cast<unsigned long long *>(s->d_tile_meta);
        p.tile_bits = reinterpret_cast<unsigned *>(s->d_tile_meta + (size_t)ntiles * sizeof(unsigned long long));
        p.bit_counter = s->d_counter; p.err = s->d_err; p.scan = s->scan_state();
        huff_bits_kernel<<<std::min<unsigned>(ntiles, (unsigned)s->dev->sm_count * 8u), kPackThreads, 0, st>>>(p);
        huff_scan_kernel<<<1, 1024, 0, st>>>(p);
        huff_pack_kernel<<<ntiles, kPackThreads, 0, st>>>(p);
        count_launch(3);
    }
    IE_CUDA(cudaGetLastError());
    return IE_OK;
}


// ---- the whole stage without a host synchronisation -------------------------------------------------------------------
// histogram kernels -> copy of hist / first occurrences to pinned memory -> HOST CALLBACK in stream order (cudaLaunchHostFunc):
// the reference's dictionary build (build_dictionary, 26 us) writing header, codes and first bit into pinned memory ->
// copies of those to the device -> bits / scan / pack kernels -> the revert rule decided on the device.  Nothing here waits on
// the host side: the caller's thread returns at once, two sessions on two streams overlap one call's dictionary build with the
// other's kernels.  One call in flight per session (the pinned staging is the session's).
constexpr size_t kHuffScratch = 256 * 8 + 256 * 4 + sizeof(HuffCodes) + 64 + 256 * 4 + 64;
struct HuffAsyncCtx { uint8_t *h_huff; int write_dict; };

static void CUDART_CB huff_build_cb(void *ud) {
    HuffAsyncCtx *c = static_cast<HuffAsyncCtx *>(ud);
    const unsigned long long *first = reinterpret_cast<const unsigned long long *>(c->h_huff + kHuffPinnedRecv);
    const unsigned *hist = reinterpret_cast<const unsigned *>(c->h_huff + kHuffPinnedRecv + 256 * 8);
    HuffCodes &codes = *reinterpret_cast<HuffCodes *>(c->h_huff + 2048);
    unsigned long long &hb = *reinterpret_cast<unsigned long long *>(c->h_huff + 2048 + sizeof(HuffCodes));
    int *status = reinterpret_cast<int *>(c->h_huff + kHuffPinnedRecv + 256 * 8 + 256 * 4 + 16);
    HostBitWriter hdr;
    const int rc = build_dictionary(hist, first, codes, hdr);
    if (!c->write_dict) { hdr.pos = 0; hdr.buf.assign(16, 0); }           // a later shard of a multi-GPU stream: codes only
    memset(c->h_huff, 0, 2048);
    if (rc != IE_OK || (hdr.pos + 7) / 8 > 2048) {
        memset(&codes, 0, sizeof codes);
        hb = 0;
        *status = IE_EINVAL;
        return;
    }
    memcpy(c->h_huff, hdr.buf.data(), std::min<size_t>(hdr.buf.size(), (hdr.pos + 7) / 8));
    hb = hdr.pos;
    *status = 0;
}

// a failed dictionary build (callback) becomes the session's device-side error flag
__global__ void huff_status_kernel(const int *h_status, int *err) {
    const int st = *h_status;
    if (st != 0 && err) atomicExch(err, st);
}

// Huffman.cpp:329-341 on the device: if the coded stream is longer than the input, the output is a '0' bit + the input bytes.
// final_bytes = what the caller's file holds; the coded stream's bit counter is left alone (every CTA reads it).
__global__ void __launch_bounds__(256) huff_revert_kernel(const uint8_t *__restrict__ in, size_t n, uint8_t *out, size_t out_cap,
                                                          const unsigned long long *bit_counter, unsigned long long *final_bytes, int *err) {
    const unsigned long long coded = (*bit_counter + 7) / 8;
    const bool revert = (unsigned long long)n < coded;
    if (blockIdx.x == 0 && threadIdx.x == 0) *final_bytes = revert ? (unsigned long long)n + 1 : coded;
    if (!revert) return;
    if (out_cap < (n + 1 + 3) / 4 * 4) { if (blockIdx.x == 0 && threadIdx.x == 0 && err) atomicExch(err, IE_ENOSPC); return; }
    const size_t nwords = (n * 8 + 1 + 31) / 32;
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += (size_t)gridDim.x * blockDim.x) {
        unsigned long long window = 0;
        for (int i = -1; i < 4; i++) {
            const long long bi = (long long)w * 4 + i;
            const unsigned byte = (bi >= 0 && (size_t)bi < n) ? in[bi] : 0u;
            window = (window << 8) | byte;
        }
        const unsigned v = (unsigned)((window >> 1) & 0xffffffffull);
        reinterpret_cast<unsigned *>(out)[w] = __byte_perm(v, 0, 0x0123);
    }
}

// d_hist_in / d_first_in != NULL: the (global) histogram and first occurrences are given, on the device (a shard of a multi-GPU
// stream): no histogram kernels, no revert rule (the caller decides it over all shards), the shard's bit count goes to d_out_bits.
static int huffman_encode_async(ie_session *s, const uint8_t *d_in, size_t n, uint8_t *d_out, size_t out_cap, unsigned long long *d_final_bytes,
                                cudaStream_t st, const unsigned *d_hist_in = nullptr, const unsigned long long *d_first_in = nullptr,
                                int write_dict = 1, unsigned long long *d_out_bits = nullptr) {
    if (out_cap < 2048 + 32) { set_error("output buffer too small"); return IE_ENOSPC; }
    IE_TRY(ensure_scratch(s, kHuffScratch));
    unsigned long long *d_first = reinterpret_cast<unsigned long long *>(s->d_scratch);
    unsigned *d_hist = reinterpret_cast<unsigned *>(s->d_scratch + 256 * 8);
    HuffCodes *d_codes = reinterpret_cast<HuffCodes *>(s->d_scratch + 256 * 8 + 256 * 4);
    unsigned *d_first_cta = reinterpret_cast<unsigned *>(s->d_scratch + 256 * 8 + 256 * 4 + sizeof(HuffCodes) + 64);
    IE_TRY(session_ensure_err(s));
    IE_TRY(ensure_huff_pinned(s));
    if (!s->huff_ctx) {
        s->huff_ctx = malloc(sizeof(HuffAsyncCtx));
        if (!s->huff_ctx) { set_error("out of memory"); return IE_ECUDA; }
    }
    static_cast<HuffAsyncCtx *>(s->huff_ctx)->h_huff = s->h_huff;
    static_cast<HuffAsyncCtx *>(s->huff_ctx)->write_dict = write_dict;
    // every allocation before the first launch: nothing below may re-allocate what an enqueued operation uses
    const unsigned ntiles = (unsigned)((n + kPackTileBytes - 1) / kPackTileBytes);
    IE_TRY(session_ensure_scan(s, 1, ntiles));
    IE_TRY(session_reserve(&s->d_tile_meta, &s->tile_meta_cap, (size_t)ntiles * (sizeof(unsigned long long) + sizeof(unsigned)) + 64));

    if (d_hist_in) {
        IE_CUDA(cudaMemcpyAsync(s->h_huff + kHuffPinnedRecv, d_first_in, 256 * 8, cudaMemcpyDeviceToHost, st));
        IE_CUDA(cudaMemcpyAsync(s->h_huff + kHuffPinnedRecv + 256 * 8, d_hist_in, 256 * 4, cudaMemcpyDeviceToHost, st));
    } else {
        IE_TRY(launch_byte_histogram(d_in, n, d_hist, d_first, d_first_cta, s->dev->sm_count, st));
        IE_CUDA(cudaMemcpyAsync(s->h_huff + kHuffPinnedRecv, d_first, 256 * 8 + 256 * 4, cudaMemcpyDeviceToHost, st));
    }
    IE_CUDA(cudaLaunchHostFunc(st, huff_build_cb, s->huff_ctx));
    HuffCodes *h_codes = reinterpret_cast<HuffCodes *>(s->h_huff + 2048);
    unsigned long long *h_hb = reinterpret_cast<unsigned long long *>(s->h_huff + 2048 + sizeof(HuffCodes));
    const int *h_status = reinterpret_cast<const int *>(s->h_huff + kHuffPinnedRecv + 256 * 8 + 256 * 4 + 16);
    IE_CUDA(cudaMemcpyAsync(d_out, s->h_huff, 2048, cudaMemcpyHostToDevice, st));               // header, zero padded
    IE_CUDA(cudaMemcpyAsync(d_codes, h_codes, sizeof(HuffCodes), cudaMemcpyHostToDevice, st));
    IE_CUDA(cudaMemcpyAsync(s->d_counter, h_hb, sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
    huff_status_kernel<<<1, 1, 0, st>>>(h_status, s->d_err);
    HuffPackParams p;
    p.in = d_in; p.n = n; p.ntiles = ntiles; p.codes = d_codes; p.out = d_out; p.out_cap = out_cap;
    p.tile_off = reinterpret_cast<unsigned long long *>(s->d_tile_meta);
    p.tile_bits = reinterpret_cast<unsigned *>(s->d_tile_meta + (size_t)ntiles * sizeof(unsigned long long));
    p.bit_counter = s->d_counter; p.err = s->d_err; p.scan = s->scan_state();
    huff_bits_kernel<<<std::min<unsigned>(ntiles, (unsigned)s->dev->sm_count * 8u), kPackThreads, 0, st>>>(p);
    huff_scan_kernel<<<1, 1024, 0, st>>>(p);
    huff_pack_kernel<<<ntiles, kPackThreads, 0, st>>>(p);
    if (d_out_bits) {
        IE_CUDA(cudaMemcpyAsync(d_out_bits, s->d_counter, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
    } else {
        huff_revert_kernel<<<(unsigned)std::min<size_t>(((n * 8 + 32) / 32 + 255) / 256, (size_t)s->dev->sm_count * 8), 256, 0, st>>>(
            d_in, n, d_out, out_cap, s->d_counter, d_final_bytes, s->d_err);
    }
    count_launch(5);
    IE_CUDA(cudaGetLastError());
    return IE_OK;
}

// the host builds the dictionary between two synchronisations of the caller's thread (huffman_variant 0: the round-1 kernels)
static int huffman_encode_sync(ie_session *s, const uint8_t *d_in, size_t n, uint8_t *d_out, size_t out_cap, size_t *out_bytes, cudaStream_t st) {
    IE_TRY(ensure_scratch(s, kHuffScratch));
    unsigned long long *d_first = reinterpret_cast<unsigned long long *>(s->d_scratch);
    unsigned *d_hist = reinterpret_cast<unsigned *>(s->d_scratch + 256 * 8);
    unsigned *d_first_cta = reinterpret_cast<unsigned *>(s->d_scratch + 256 * 8 + 256 * 4 + sizeof(HuffCodes) + 64);
    IE_TRY(session_ensure_err(s));
    IE_TRY(launch_byte_histogram(d_in, n, d_hist, d_first, d_first_cta, s->dev->sm_count, st));
    IE_TRY(ensure_huff_pinned(s));
    const unsigned long long *first = reinterpret_cast<const unsigned long long *>(s->h_huff + kHuffPinnedRecv);
    const unsigned *hist = reinterpret_cast<const unsigned *>(s->h_huff + kHuffPinnedRecv + 256 * 8);
    IE_CUDA(cudaMemcpyAsync(s->h_huff + kHuffPinnedRecv, d_first, 256 * 8 + 256 * 4, cudaMemcpyDeviceToHost, st));
    IE_CUDA(cudaStreamSynchronize(st));
    IE_TRY(huffman_pack_dev(s, d_in, n, hist, first, 1, d_out, out_cap, st));
    // the stream's size and the error flag come back together
    unsigned long long *h_bits = reinterpret_cast<unsigned long long *>(s->h_huff + kHuffPinnedRecv + 256 * 8 + 256 * 4);
    int *h_err = reinterpret_cast<int *>(h_bits + 1);
    IE_CUDA(cudaMemcpyAsync(h_bits, s->d_counter, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    IE_CUDA(cudaMemcpyAsync(h_err, s->d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    IE_CUDA(cudaStreamSynchronize(st));
    if (*h_err != 0) {
        const int e = *h_err;
        IE_CUDA(cudaMemsetAsync(s->d_err, 0, sizeof(int), st));
        set_error(e == IE_ENOSPC ? "output buffer too small" : "device-side error");
        return e;
    }
    size_t total = (size_t)((*h_bits + 7) / 8);
    if (n < total) {                                                       // Huffman.cpp:329-341
        total = n + 1;
        if (out_cap < (total + 3) / 4 * 4) { set_error("output buffer too small"); return IE_ENOSPC; }
        const size_t nwords = (n * 8 + 1 + 31) / 32;
        shift_copy_kernel<<<(unsigned)std::min<size_t>((nwords + 255) / 256, 65535), 256, 0, st>>>(d_in, n, d_out, 1);
        count_launch();
        IE_CUDA(cudaGetLastError());
    }
    *out_bytes = total;
    return IE_OK;
}

}  // namespace ie

using namespace ie;

extern "C" {

int ie_byte_histogram_dev(const uint8_t *d_in, size_t n, uint32_t *hist, uint64_t *first_pos, void *stream) {
    if (!d_in || !hist || !first_pos) { set_error("NULL argument"); return IE_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    DeviceState *dev;
    IE_TRY(get_device_state(&dev));
    uint8_t *d = nullptr;
    IE_CUDA(cudaMalloc(&d, 256 * 4 + 256 * 8 + 256 * 4));
    unsigned *d_hist = reinterpret_cast<unsigned *>(d + 256 * 8);
    unsigned long long *d_first = reinterpret_cast<unsigned long long *>(d);
    {
        const int rc = launch_byte_histogram(d_in, n, d_hist, d_first, reinterpret_cast<unsigned *>(d + 256 * 8 + 256 * 4), dev->sm_count, st);
        if (rc != IE_OK) { cudaFree(d); return rc; }
    }
    cudaMemcpyAsync(hist, d_hist, 256 * 4, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(first_pos, d_first, 256 * 8, cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(d);
    if (e != cudaSuccess) return cuda_fail(e, "byte histogram", __FILE__, __LINE__);
    return IE_OK;
}

int ie_huffman_encode_async_dev(ie_session *s, const uint8_t *d_in, size_t n, uint8_t *d_out, size_t out_cap, uint64_t *d_out_bytes,
                                void *stream) {
    if (!s || !d_in || !d_out || !d_out_bytes || n == 0) { set_error("NULL/empty argument"); return IE_EINVAL; }
    if ((uintptr_t)d_out % 16) { set_error("stream buffers must be 16-byte aligned"); return IE_EINVAL; }
    return huffman_encode_async(s, d_in, n, d_out, out_cap, reinterpret_cast<unsigned long long *>(d_out_bytes), (cudaStream_t)stream);
}

int ie_huffman_encode_dev(ie_session *s, const uint8_t *d_in, size_t n, uint8_t *d_out, size_t out_cap, size_t *out_bytes,
                          void *stream) {
    if (!s || !d_in || !d_out || !out_bytes || n == 0) { set_error("NULL/empty argument"); return IE_EINVAL; }
    if ((uintptr_t)d_out % 16) { set_error("stream buffers must be 16-byte aligned"); return IE_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    // synchronous entry point: the host builds the dictionary between two synchronisations of the caller's thread (0.196 ms on
    // the 27.7 MB stream; through the stream-ordered callback of ie_huffman_encode_async_dev one isolated call takes 0.25 ms,
    // but two sessions overlap to 0.144 ms per call)
    return huffman_encode_sync(s, d_in, n, d_out, out_cap, out_bytes, st);
}

int ie_byte_histogram_async_dev(ie_session *s, const uint8_t *d_in, size_t n, uint32_t *d_hist, uint64_t *d_first, void *stream) {
    if (!s || !d_in || !d_hist || !d_first || n == 0) { set_error("NULL/empty argument"); return IE_EINVAL; }
    IE_TRY(ensure_scratch(s, kHuffScratch));
    unsigned *d_first_cta = reinterpret_cast<unsigned *>(s->d_scratch + 256 * 8 + 256 * 4 + sizeof(HuffCodes) + 64);
    return launch_byte_histogram(d_in, n, d_hist, reinterpret_cast<unsigned long long *>(d_first), d_first_cta, s->dev->sm_count, (cudaStream_t)stream);
}

int ie_huffman_encode_shard_async_dev(ie_session *s, const uint8_t *d_in, size_t n, const uint32_t *d_hist, const uint64_t *d_first,
                                      int write_dictionary, uint8_t *d_out, size_t out_cap, uint64_t *d_out_bits, void *stream) {
    if (!s || !d_in || !d_out || !d_hist || !d_first || !d_out_bits || n == 0) { set_error("NULL/empty argument"); return IE_EINVAL; }
    if ((uintptr_t)d_out % 16) { set_error("stream buffers must be 16-byte aligned"); return IE_EINVAL; }
    return huffman_encode_async(s, d_in, n, d_out, out_cap, nullptr, (cudaStream_t)stream, d_hist, reinterpret_cast<const unsigned long long *>(d_first),
                                write_dictionary, reinterpret_cast<unsigned long long *>(d_out_bits));
}

int ie_huffman_encode_shard_dev(ie_session *s, const uint8_t *d_in, size_t n, const uint32_t *hist, const uint64_t *first_pos,
                                int write_dictionary, uint8_t *d_out, size_t out_cap, uint64_t *d_out_bits, void *stream) {
    if (!s || !d_in || !d_out || !hist || !first_pos || n == 0) { set_error("NULL/empty argument"); return IE_EINVAL; }
    if ((uintptr_t)d_out % 16) { set_error("stream buffers must be 16-byte aligned"); return IE_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    IE_TRY(ensure_scratch(s, 256 * 8 + 256 * 4 + sizeof(HuffCodes) + 64 + 256 * 4));
    IE_TRY(session_ensure_err(s));
    unsigned h[256];
    unsigned long long f[256];
    for (int i = 0; i < 256; i++) { h[i] = hist[i]; f[i] = first_pos[i]; }
    IE_TRY(huffman_pack_dev(s, d_in, n, h, f, write_dictionary, d_out, out_cap, st));
    if (d_out_bits) IE_CUDA(cudaMemcpyAsync(d_out_bits, s->d_counter, sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
    return read_err_flag(s, st);
}

int ie_stream_shift_dev(const uint8_t *d_in, const uint64_t *d_params, uint8_t *d_out, size_t out_cap, void *stream) {
    if (!d_in || !d_params || !d_out) { set_error("NULL argument"); return IE_EINVAL; }
    if ((uintptr_t)d_out % 16) { set_error("stream buffers must be 16-byte aligned"); return IE_EINVAL; }
    DeviceState *dev;
    IE_TRY(get_device_state(&dev));
    stream_shift_kernel<<<dev->sm_count * 8, 256, 0, (cudaStream_t)stream>>>(d_in, reinterpret_cast<const unsigned long long *>(d_params),
                                                                            d_out, out_cap);
    count_launch();
    IE_CUDA(cudaGetLastError());
    return IE_OK;
}

int ie_huffman_decode_dev(ie_session *s, const uint8_t *d_in, size_t n, uint8_t *d_out, size_t out_cap, size_t *out_bytes,
                          uint64_t *start_bit, void *stream) {
    if (!s || !d_in || !d_out || !out_bytes || !start_bit || n == 0) { set_error("NULL/empty argument"); return IE_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    IE_TRY(session_ensure_err(s));
    // dictionary header: at most 256 * (8 + 15) + 16 * 12 + 1 bits < 1 KiB (4-bit length field); parse it on the host
    std::vector<uint8_t> head(std::min<size_t>(n, 1024));
    IE_CUDA(cudaMemcpyAsync(head.data(), d_in, head.size(), cudaMemcpyDeviceToHost, st));
    IE_CUDA(cudaStreamSynchronize(st));
    size_t pos = 0;
    auto get_bit = [&]() -> uint32_t { if ((pos >> 3) >= head.size()) return 0; uint32_t b = (head[pos >> 3] >> (7 - (pos & 7))) & 1u; pos++; return b; };
    auto get = [&](unsigned l) { uint32_t v = 0; for (unsigned i = 0; i < l; i++) v |= get_bit() << (l - i - 1); return v; };
    std::vector<HuffDecodeTables> tabv(1);
    HuffDecodeTables &t = tabv[0];
    memset(&t, 0, sizeof t);
    for (auto &c : t.child) c[0] = c[1] = -1;
    t.nodes = 1;                                                           // node 0 = root (Huffman.cpp:123)
    bool any = false;
    while (get_bit()) {                                                    // Huffman.cpp:57-65, 128-142
        uint32_t seq = get(7), bl = get(4);
        while (seq--) {
            const uint32_t key = get(8), word = get(bl);
            if (bl == 0) { set_error("zero-length Huffman code (undefined in the reference decoder)"); return IE_EFORMAT; }
            int cur = 0;                                                   // Huffman.cpp:153-180
            for (int b = (int)bl - 1; b >= 0; b--) {
                const int dir = (word >> b) & 1;
                if (b == 0) { t.child[cur][dir] = (short)(-(int)key - 2); }
                else {
                    if (t.child[cur][dir] < 0) {
                        if (t.nodes >= 512) { set_error("Huffman dictionary too large"); return IE_EFORMAT; }
                        t.child[cur][dir] = (short)t.nodes++;
                    }
                    cur = t.child[cur][dir];
                }
            }
            any = true;
        }
    }
    if (!any) {                                                            // pass-through (Huffman.cpp:361-371)
        if (out_cap < n) { set_error("output buffer too small"); return IE_ENOSPC; }
        IE_CUDA(cudaMemcpyAsync(d_out, d_in, n, cudaMemcpyDeviceToDevice, st));
        *out_bytes = n;
        *start_bit = pos;
        return IE_OK;
    }
    // primary 12-bit table by walking the tree for every prefix
    for (unsigned pre = 0; pre < 4096; pre++) {
        int cur = 0;
        unsigned entry = 0;
        for (int l = 0; l < 12; l++) {
            const int dir = (pre >> (11 - l)) & 1;
            const int c = t.child[cur][dir];
            if (c == -1) { entry = 0; cur = -1; break; }
            if (c <= -2) { entry = ((unsigned)(l + 1) << 16) | (unsigned)(-c - 2); cur = -1; break; }
            cur = c;
        }
        if (cur >= 0) entry = 0xFFFF0000u | (unsigned)cur;
        t.lut[pre] = entry;
    }
    const size_t tab_bytes = (sizeof(HuffDecodeTables) + 63) / 64 * 64;
    IE_TRY(ensure_scratch(s, tab_bytes + 64));
    HuffDecodeTables *d_tab = reinterpret_cast<HuffDecodeTables *>(s->d_scratch);
    unsigned long long *d_count = reinterpret_cast<unsigned long long *>(s->d_scratch + tab_bytes);
    IE_CUDA(cudaMemcpyAsync(d_tab, &t, sizeof t, cudaMemcpyHostToDevice, st));
    {
        if ((uintptr_t)d_in % 4) { set_error("Huffman stream must be 4-byte aligned"); return IE_EINVAL; }
        HuffParParams hp;
        hp.in = d_in; hp.total = (unsigned long long)n * 8ull; hp.start_bit = (unsigned long long)pos; hp.tab = d_tab;
        hp.ngroups = (unsigned)((hp.total - std::min<unsigned long long>(hp.total, hp.start_bit) + kHuffGroupBits - 1) / kHuffGroupBits + 1);
        const size_t need_par = (size_t)hp.ngroups * 2 * sizeof(uint2) + 64;
        IE_TRY(session_reserve(&s->d_parse, &s->parse_cap, need_par));
        hp.entry = reinterpret_cast<uint2 *>(s->d_parse);
        hp.exit_ = hp.entry + hp.ngroups;
        hp.flags = reinterpret_cast<unsigned *>(hp.exit_ + hp.ngroups);
        hp.out = d_out; hp.out_cap = out_cap; hp.out_count = d_count; hp.err = s->d_err;
        const unsigned gb = (hp.ngroups + 63) / 64;
        huff_spec_walk<<<gb, 64, 0, st>>>(hp);
        for (int r = 0; r < kHuffRounds; r++) huff_spec_repair<<<gb, 64, 0, st>>>(hp);
        huff_spec_finish<<<1, 1024, 0, st>>>(hp);
        huff_spec_emit<<<gb, 64, 0, st>>>(hp);
        huff_decode_kernel<<<1, 32, 0, st>>>(d_in, n, (unsigned long long)pos, d_tab, d_out, out_cap, d_count, s->d_err, hp.flags);
        count_launch(4 + kHuffRounds);
    }
    count_launch();
    IE_CUDA(cudaGetLastError());
    IE_CUDA(cudaMemcpyAsync(s->h_pinned, d_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    IE_TRY(read_err_flag(s, st));
    *out_bytes = (size_t)s->h_pinned[0];
    *start_bit = 0;
    return IE_OK;
}

}  // extern "C"
