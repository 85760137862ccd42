// Staging of a span of the bit stream in shared memory (shared by the parser and the block decoder).
#pragma once
#include "common.cuh"

namespace ie {

// ---- stream staging: the bits a CTA walks are first copied to shared memory with 16-byte cp.async (coalesced, no
// registers, zero fill past the end), because a walk is a chain of dependent reads: ~80 of them per group, each a round
// trip to L2/HBM when done on global memory (measured: 127 us for the walk kernel on 27.7 MB), ~30 cycles in shared memory.
__device__ __forceinline__ void cp_async16_zfill(void *smem_dst, const void *gsrc, unsigned src_bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory"); }

struct StagedStream {
    const unsigned *w;                 // shared memory, raw little-endian words of the stream
    unsigned long long base;           // absolute bit of w[0]
    unsigned total_rel;                // stream end relative to base (clamped)
};

// stage the words holding bits [first_bit, last_bit) (+ the 64-bit header window) of the stream; returns the view.
// Positions inside the view are 32-bit offsets from `base` (a CTA's range is < 2^20 bits).
__device__ __forceinline__ StagedStream stage_stream(unsigned *s_w, unsigned cap_words, const uint8_t *enc, unsigned long long total,
                                                     unsigned long long first_bit, unsigned long long last_bit) {
    StagedStream st;
    st.w = s_w;
    const unsigned long long w0 = (first_bit >> 7) << 2;                   // 16-byte granules
    st.base = w0 * 32;
    st.total_rel = (unsigned)min(total - min(total, st.base), 0x7FFFFFFFull);
    const unsigned long long nbytes = ((total + 31) >> 5) << 2;            // readable bytes (whole words, as block_bits_at)
    const unsigned long long b0 = w0 * 4;
    unsigned long long b1 = ((last_bit + 64 + 127) >> 7) << 4;
    if (b1 > b0 + (unsigned long long)cap_words * 4) b1 = b0 + (unsigned long long)cap_words * 4;
    const unsigned n16 = (unsigned)((b1 - b0) >> 4);
    for (unsigned i = threadIdx.x; i < n16; i += blockDim.x) {
        const unsigned long long b = b0 + (unsigned long long)i * 16;
        const unsigned have = (b >= nbytes) ? 0u : (unsigned)min(16ull, nbytes - b);
        cp_async16_zfill(s_w + i * 4, enc + (have ? b : 0ull), have);
    }
    cp_async_wait_all();
    __syncthreads();
    // bits past the end of the stream read as 0 (BitStream.cpp:17-20): whole words beyond it were zero-filled above, the
    // word that holds the end is trimmed here, so readers need no end-of-stream test of their own
    if (threadIdx.x == 0 && (st.total_rel & 31u) && (st.total_rel >> 5) < n16 * 4u) {
        unsigned *wp = s_w + (st.total_rel >> 5);
        *wp = __byte_perm(__byte_perm(*wp, 0, 0x0123) & ~(0xFFFFFFFFu >> (st.total_rel & 31u)), 0, 0x0123);
    }
    __syncthreads();
    return st;
}

}  // namespace ie
