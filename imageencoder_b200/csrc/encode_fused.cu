// Fused stream encoder: pixels -> final stream in ONE persistent kernel (ie_set_option("encode_variant", 8)).
//
// Replaces ImageEncoder.cpp:121-138 (the parallel DCT loop + the strictly serial streamEncoded loop) like the two-kernel path
// of encode_image.cu (encode_tiles_kernel + tile_copyout_fast_kernel), with the three things its ncu captures blamed removed
// (profiles/r1_encode_v9_summary.md: 24 % of all warp time in CTA barriers, 17 % behind the exact queue alone; an 18 % copy-out
// kernel that only exists because the packed tiles round-trip through 137 MB of scratch):
//
//   * WARP-AUTONOMOUS tiles.  One CTA per SM, NPW producer warps (24 at 8x8); a producer warp owns "warp-tiles" of 32 (8x8) or
//     128 (4x4) consecutive blocks, lane per block (per 4 blocks), and never meets a CTA-wide barrier.  Per warp: one staging
//     area for the quantised coefficients and a small ring of packed tile images -- 8.6 KiB, which is what lets 24 warps share
//     an SM (the first version of this kernel, triple-buffered staging + an exact-queue helper warp, fitted 14 and was slower
//     than the two-kernel path: profiles/r2_fused_helper_design_summary.md).
//   * guard-band coefficients (transform_fast.cuh) are resolved by the lane that owns the block, all lanes of the warp that
//     have one at the same time (short binary64 chain first, the reference's 2*N*N-step chain for the true ties, exact.cuh).
//   * FUSED COPY-OUT.  The stream position of a tile needs the bit totals of every earlier tile.  Tiles are dealt in rounds
//     (round r: CTA c takes the NPW consecutive warp-tiles of "CTA-tile" r*G + c); a helper warp adds up the published
//     CTA-tile totals (G epoch-tagged u64 per round, L2) and hands the producers their base.  A producer packs its tile at
//     tile-local alignment into its image ring and goes on with the next tile; an image leaves for the stream (funnel-shifted
//     to its final alignment, 128-bit stores) as soon as its base is known -- up to three tiles later, so nobody waits in
//     practice (the waits are real spins, correctness does not depend on timing).  No scratch round trip, no second kernel.
//     Chunks two warp-tiles share: shared-memory hand-off inside a CTA-tile, the fence-free records of pack.cuh between
//     CTA-tiles.
#include "encode_image.cuh"
#include "transform_fast.cuh"
#include "exact.cuh"

namespace ie {

std::atomic<int> g_fused_debug{0};

namespace {

constexpr size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

template <int N>
struct Fused {
    static constexpr int BPL = (N == 8) ? 1 : 4;          // blocks per lane
    static constexpr int TB = 32 * BPL;                   // blocks per warp-tile
    static constexpr int NN = N * N;
    static constexpr int STRIDE = NN + 2;                 // halfwords per block in the staging area (odd word count: bank spread)
    static constexpr int NSEG = NN / 8;
    static constexpr int NPW = (N == 8) ? 24 : 22;        // producer warps per CTA
    static constexpr int THREADS = (NPW + 1) * 32;        // + the prefix helper
    static constexpr int QMAX = 3;                        // packed tiles a warp may hold back while their base is unknown
    static constexpr int MAX_TILE_WORDS = (TB * (4 + 16 + 16 * NN) + 31) / 32;
    static constexpr int IMG_PAD_FRONT = 4, IMG_PAD_BACK = 8;              // zero words around an image (unconditional funnel loads)
    static constexpr int RING_WORDS = ((IMG_PAD_FRONT + MAX_TILE_WORDS + 3 + IMG_PAD_BACK + 40 /* stream prefix */) + 31) / 32 * 32;
    static constexpr int RINGR = 16;                      // rounds a CTA-level slot stays valid (> 2 * QMAX + 2)
    static constexpr size_t COEF_BYTES = (size_t)TB * STRIDE * 2;
    // per producer warp
    static constexpr size_t W_COEF = 0;
    static constexpr size_t W_RING = align16(W_COEF + COEF_BYTES);
    static constexpr size_t W_BYTES = align16(W_RING + (size_t)RING_WORDS * 4);
    // CTA level, behind the NPW warp regions
    static constexpr size_t C_TOT = (size_t)NPW * W_BYTES;                 // u32 [RINGR][32]: tag << 24 | warp-tile bits
    static constexpr size_t C_BASE = C_TOT + (size_t)RINGR * 32 * 4;       // u64 [RINGR]
    static constexpr size_t C_BTAG = C_BASE + RINGR * 8;                   // u32 [RINGR]
    static constexpr size_t C_MISC = C_BTAG + RINGR * 4;                   // u32 [4]: CTA id
    static constexpr size_t C_HAND = align16(C_MISC + 16);                 // [RINGR][NPW - 1] hand-off slots of 5 words
    static constexpr size_t SMEM = align16(C_HAND + (size_t)RINGR * (NPW - 1) * 20);
    // staging slot of block lb: lanes must hit different banks when they store pair words (lane stride = BPL blocks)
    __device__ static __forceinline__ int slot(int lb) { return (BPL == 1) ? lb : ((lb & (BPL - 1)) * 32 + lb / BPL); }
};
static_assert(Fused<8>::SMEM <= 232448 && Fused<4>::SMEM <= 232448, "shared memory budget of one CTA per SM");
static_assert(Fused<8>::THREADS <= 1024 && Fused<4>::THREADS <= 1024, "CTA size");

struct FusedParams {
    unsigned n_wtiles;                // warp-tiles of the image
    unsigned n_ctatiles;              // groups of NPW warp-tiles
    unsigned long long *agg;          // [n_ctatiles] epoch-tagged CTA-tile totals (ScanState::tile_state)
    TileBoundary *bnd;                // [n_ctatiles]
    unsigned *ticket;                 // [2]: dynamic CTA id, CTAs finished (both return to 0 when the grid ends)
    unsigned epoch;
    int append;                       // 1: the stream continues at *bit_counter (no prefix)
    int debug;                        // timing experiments only (ie_set_option("fused_debug")): 1 = made-up bases (WRONG output)
};

// flags in shared memory: volatile accesses ordered by CTA-scope fences
__device__ __forceinline__ unsigned ld_flag(const unsigned *p) { return *reinterpret_cast<const volatile unsigned *>(p); }
__device__ __forceinline__ void st_flag_release(unsigned *p, unsigned v) {
    __threadfence_block();
    *reinterpret_cast<volatile unsigned *>(p) = v;
}

__device__ __forceinline__ unsigned warp_sum(unsigned v) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
__device__ __forceinline__ unsigned long long warp_sum64(unsigned long long v) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// RLE info of one staged block from scratch (after a guard-band patch): bit 31 | prevnz << 23 | lastnz << 16 | orbits
template <int NN>
__device__ __noinline__ unsigned stats_full(const short *cf) {
    int lastnz = 0, prevnz = 0;
    unsigned orbits = 0;
    for (int k = 0; k < NN; k++) {
        const int q = cf[k];
        if (q != 0) {
            orbits |= (unsigned)(q ^ (q >> 31));
            prevnz = (k < NN - 1) ? (k + 1) : prevnz;
            lastnz = k + 1;
        }
    }
    return 0x80000000u | ((unsigned)prevnz << 23) | ((unsigned)lastnz << 16) | (orbits & 0xffffu);
}


// RLE info of a block in the form the quantiser leaves it: non-empty 8-coefficient zigzag segments << 16 | orbits
// (bit_length(orbits) + 1 = widest bits_needed of the block, utils.hpp:226-243).  A guard-band patch of coefficient k
// (q_old -> q_new) is folded in here; only when the patched coefficient alone carried the widest value is the block rescanned.
template <int NN>
__device__ __forceinline__ unsigned stat_after_patch(unsigned stat, const short *cf, int k, int q_old, int q_new) {
    if (stat >> 31) return stats_full<NN>(cf);
    const unsigned bo = (unsigned)(q_old ^ (q_old >> 31)), bn = (unsigned)(q_new ^ (q_new >> 31));
    unsigned orbits = stat & 0xffffu;
    if (__clz(bn) > __clz(bo) && __clz(bo) == __clz(orbits)) return stats_full<NN>(cf);   // q_old held the widest value's top bit, q_new does not
    orbits |= bn;
    const int s = k >> 3;
    const unsigned *cw = reinterpret_cast<const unsigned *>(cf) + s * 4;
    const unsigned any = cw[0] | cw[1] | cw[2] | cw[3];
    unsigned segmask = (stat >> 16) & 0xffu;
    segmask = any ? (segmask | (1u << s)) : (segmask & ~(1u << s));
    return (segmask << 16) | orbits;
}

// one block's fields, MSB first, into the image at bit `pos` (Block.cpp:371-413): whole words with plain stores, the <= 2
// words shared with the neighbouring blocks with shared-memory atomicOr
__device__ __forceinline__ void pack_block(unsigned *img, unsigned pos, int w, int len, const unsigned *cw, int use_rle) {
    const unsigned bo = pos & 31u;
    unsigned *ow = img + (pos >> 5);
    const unsigned mask = (1u << w) - 1u;                                             // w <= 16
    unsigned long long acc = use_rle ? ((((unsigned long long)w & 15ull) << w) | (unsigned long long)len) : ((unsigned long long)w & 15ull);
    int nacc = (int)bo + 4 + (use_rle ? w : 0);
    const int nfull = len >> 1;
    const int s2 = 2 * w;
    int j = 0;
    bool odd_left = (len & 1) != 0;
    while (nacc < 32 && j < nfull) {
        const unsigned x2 = cw[j++];
        acc = (acc << s2) | (((x2 & mask) << w) | ((x2 >> 16) & mask));
        nacc += s2;
    }
    if (nacc < 32 && odd_left) { acc = (acc << w) | (cw[nfull] & mask); nacc += w; odd_left = false; }
    if (nacc >= 32) {
        const unsigned word = (unsigned)(acc >> (nacc - 32));
        if (bo) atomicOr(ow, word); else *ow = word;
        ow++;
        nacc -= 32;
        for (; j < nfull; j++) {
            const unsigned x2 = cw[j];
            acc = (acc << s2) | (((x2 & mask) << w) | ((x2 >> 16) & mask));
            nacc += s2;
            if (nacc >= 32) { nacc -= 32; *ow++ = (unsigned)(acc >> nacc); }
        }
        if (odd_left) {
            acc = (acc << w) | (cw[nfull] & mask);
            nacc += w;
            if (nacc >= 32) { nacc -= 32; *ow++ = (unsigned)(acc >> nacc); }
        }
    }
    if (nacc > 0) atomicOr(ow, (unsigned)(acc << (32 - nacc)));                        // tail shared with the next block
}

// hand-off of a 128-bit chunk two warp-tiles share (pack.cuh, TileBoundary): the second arriver stores the word
__device__ __forceinline__ void handoff_chunk(TileBoundary *bd, uint4 *dst, const uint4 v) {
    unsigned *dw = reinterpret_cast<unsigned *>(dst);
    const unsigned mine[4] = {v.x, v.y, v.z, v.w};
    unsigned long long old[4];
#pragma unroll
    for (int i = 0; i < 4; i++) old[i] = atomicAdd(&bd->w[i], (1ull << 32) | (unsigned long long)mine[i]);   // four round trips in flight
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if ((old[i] >> 32) == 1ull) {
            dw[i] = (unsigned)old[i] | mine[i];
            atomicExch(&bd->w[i], 0ull);
        }
    }
}


// the same between two warps of one CTA (all but one boundary of a CTA-tile): both sides OR their bits into a zeroed slot in
// shared memory (their bits are disjoint); the second arriver stores the chunk and leaves the slot zeroed
struct SmemHandoff {
    unsigned w[4];
    unsigned flag;
};
__device__ __forceinline__ void handoff_chunk_smem(SmemHandoff *h, uint4 *dst, const uint4 v) {
    atomicOr(&h->w[0], v.x); atomicOr(&h->w[1], v.y); atomicOr(&h->w[2], v.z); atomicOr(&h->w[3], v.w);
    __threadfence_block();
    if (atomicAdd(&h->flag, 1u) == 1u) {
        __threadfence_block();
        uint4 o;
        o.x = atomicExch(&h->w[0], 0u); o.y = atomicExch(&h->w[1], 0u); o.z = atomicExch(&h->w[2], 0u); o.w = atomicExch(&h->w[3], 0u);
        *dst = o;
        __threadfence_block();
        h->flag = 0u;            // next use: RINGR rounds later
    }
}

// a packed tile waiting in the ring for its base
struct Pending {
    unsigned start;              // first word of the entry in the ring (IMG_PAD_FRONT zero words, then the image)
    unsigned nbits;              // bits of the image (for the stream's first tile: prefix + header + blocks)
    unsigned round;              // its round (CTA-tile = round * G + c)
    unsigned words;              // ring words the entry occupies
};

// ---- copy-out: the image's bits [0, nbits) become stream bits [P, P + nbits) -------------------------------------------------
template <int N>
__device__ __forceinline__ void fused_copyout(const EncodeParams &p, const FusedParams &f, const unsigned *ring, const Pending &e,
                                              unsigned long long P, unsigned g, unsigned ct, int warp, bool merge_head,
                                              SmemHandoff *hand, int lane) {
    using F = Fused<N>;
    const bool last_tile = (g + 1 == f.n_wtiles);
    const unsigned g0 = (unsigned)(P & 127ull);
    const unsigned long long c0 = P >> 7;
    const unsigned nchunks = (g0 + e.nbits + 127u) / 128u;
    const bool head_shared = g0 != 0;
    const bool tail_shared = ((g0 + e.nbits) & 127u) != 0 && !last_tile;
    const unsigned *img = ring + e.start + F::IMG_PAD_FRONT;
    uint4 *out4 = reinterpret_cast<uint4 *>(p.out);
    const unsigned long long cap_chunks = p.out_cap / 16ull;
    for (unsigned c = lane; c < nchunks; c += 32) {
        const int ls = (int)(c * 128u) - (int)g0;                 // image bit of the chunk's first bit (>= -127)
        const int wi = ls >> 5;                                   // floor; >= -4: the zero words in front of the image
        const unsigned sh = (unsigned)ls & 31u;
        const unsigned w0 = img[wi], w1 = img[wi + 1], w2 = img[wi + 2], w3 = img[wi + 3], w4 = img[wi + 4];
        uint4 v;
        v.x = __byte_perm(__funnelshift_l(w1, w0, sh), 0, 0x0123);
        v.y = __byte_perm(__funnelshift_l(w2, w1, sh), 0, 0x0123);
        v.z = __byte_perm(__funnelshift_l(w3, w2, sh), 0, 0x0123);
        v.w = __byte_perm(__funnelshift_l(w4, w3, sh), 0, 0x0123);
        if (c0 + c >= cap_chunks) { if (p.err) atomicExch(p.err, IE_ENOSPC); continue; }
        uint4 *dst = out4 + c0 + c;
        if (c == 0 && head_shared) {
            if (merge_head) {                                      // append: merge with what the earlier launch left here
                const uint4 o = *dst;
                v.x |= o.x; v.y |= o.y; v.z |= o.z; v.w |= o.w;
                *dst = v;
            } else if (warp > 0) {
                handoff_chunk_smem(&hand[warp - 1], dst, v);       // shared with the previous warp of this CTA
            } else {
                handoff_chunk(&f.bnd[ct - 1], dst, v);             // shared with the previous CTA-tile
            }
        } else if (c + 1 == nchunks && tail_shared) {
            if (warp + 1 < F::NPW) handoff_chunk_smem(&hand[warp], dst, v);
            else handoff_chunk(&f.bnd[ct], dst, v);
        } else {
            *dst = v;
        }
    }
    if (last_tile && lane == 0) {
        p.bit_counter[0] = P + e.nbits;
        if (p.out_bits) p.out_bits[0] = P + e.nbits;
    }
}

template <int N>
__global__ void __launch_bounds__(Fused<N>::THREADS, 1) encode_fused_kernel(const EncodeParams p, const FusedParams f) {
    using F = Fused<N>;
    constexpr int NN = F::NN, BPL = F::BPL, STRIDE = F::STRIDE, NSEG = F::NSEG, NPW = F::NPW, RINGR = F::RINGR;
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned *s_tot = reinterpret_cast<unsigned *>(smem + F::C_TOT);
    unsigned long long *s_base = reinterpret_cast<unsigned long long *>(smem + F::C_BASE);
    unsigned *s_btag = reinterpret_cast<unsigned *>(smem + F::C_BTAG);
    unsigned *s_misc = reinterpret_cast<unsigned *>(smem + F::C_MISC);
    SmemHandoff *s_hand = reinterpret_cast<SmemHandoff *>(smem + F::C_HAND);
    static_assert(sizeof(SmemHandoff) == 20, "hand-off slot size");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // dynamic CTA id: every CTA with a smaller id is running (a spinning CTA only ever waits for smaller ids)
    if (threadIdx.x == 0) s_misc[0] = atomicAdd(&f.ticket[0], 1u);
    for (int i = threadIdx.x; i < RINGR * 32; i += blockDim.x) s_tot[i] = 0;
    for (int i = threadIdx.x; i < RINGR; i += blockDim.x) s_btag[i] = 0;
    for (int i = threadIdx.x; i < RINGR * (NPW - 1) * 5; i += blockDim.x) reinterpret_cast<unsigned *>(s_hand)[i] = 0;
    __syncthreads();                                                       // the only CTA-wide barrier before the end
    const unsigned c = s_misc[0], Gc = gridDim.x;
    const unsigned rounds = (f.n_ctatiles > c) ? (f.n_ctatiles - c + Gc - 1) / Gc : 0u;
    const unsigned long long ep = (unsigned long long)(f.epoch & 0xFFFFFFu);

    if (warp < NPW) {
        // ================================================= producer warps =================================================
        unsigned char *wbase = smem + (size_t)warp * F::W_BYTES;
        short *coef = reinterpret_cast<short *>(wbase + F::W_COEF);
        unsigned *ring = reinterpret_cast<unsigned *>(wbase + F::W_RING);
        ExactCtx ex;
        ex.src = p.src; ex.ref = nullptr; ex.res_coord = nullptr; ex.tab = p.tab; ex.pitch = p.pitch; ex.bx = p.bx; ex.mbx = 0;
        // FIFO of packed tiles waiting for their base (warp-uniform registers)
        Pending q0{0, 0, 0, 0}, q1{0, 0, 0, 0}, q2{0, 0, 0, 0};
        int nq = 0;
        unsigned head = 0;                                                 // next free word of the ring (entries are contiguous)

        // copies the oldest waiting tile out if its base is known (or waits for it); returns false if it is not known yet
        auto drain_one = [&](bool blocking) -> bool {
            const unsigned r = q0.round;
            const unsigned g = (r * Gc + c) * NPW + warp;
            unsigned long long P;
            if (f.debug & 1) {
                P = (unsigned long long)g * 6784ull + 1024ull;
            } else {
                for (;;) {
                    bool ok = ld_flag(&s_btag[r % RINGR]) == r + 1;
                    unsigned t = 0;
                    if (lane < warp) { t = ld_flag(&s_tot[(r % RINGR) * 32 + lane]); ok = ok && (t >> 24) == ((r + 1) & 0xffu); }
                    if (__all_sync(0xffffffffu, ok)) {
                        __threadfence_block();
                        P = s_base[r % RINGR] + warp_sum(t & 0xffffffu);
                        break;
                    }
                    if (!blocking) return false;
                    __nanosleep(40);
                }
            }
            const bool first_tile = (g == 0);
            if (first_tile && !f.append) P = 0;                            // its image starts with the stream's prefix
            fused_copyout<N>(p, f, ring, q0, P, g, r * Gc + c, warp, first_tile && f.append, s_hand + (r % RINGR) * (NPW - 1), lane);
            q0 = q1; q1 = q2; nq--;
            return true;
        };

        for (unsigned it = 0; it < rounds; it++) {
            // the producer warps start every tile together: 24 warps spread over a 5000-instruction program miss the
            // instruction cache most of the time (measured: 3.9 no-instruction stall cycles per issue without this)
            if (!(f.debug & 4)) asm volatile("bar.sync 1, %0;" ::"n"(NPW * 32) : "memory");
            const unsigned g = (it * Gc + c) * NPW + warp;
            const unsigned first_blk = g * F::TB;
            const int nblk = (g < f.n_wtiles) ? (int)min((unsigned)F::TB, p.nblocks - first_blk) : 0;
            if (it + 1 < rounds) {
                // next tile's pixels on their way into L2 while this one is transformed: one lane per 128-byte line of the
                // tile's pixel rows (a hint only)
                const unsigned gn = ((it + 1) * Gc + c) * NPW + warp;
                constexpr int LPR = F::TB * N / 128;
                const unsigned nb = gn * F::TB + (unsigned)(lane % LPR) * (128u / N);
                if (gn < f.n_wtiles && lane < N * LPR && nb < p.nblocks) {
                    const unsigned byi = nb / p.bx, bxi = nb - byi * p.bx;
                    const uint8_t *row = p.src + (size_t)(byi * N + lane / LPR) * p.pitch + (size_t)bxi * N;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(row));
                }
            }
            // ---- phase 1: load, transform, quantise (the arithmetic of encode_tiles_kernel's variant 2, transform_fast.cuh) ----
            unsigned stat[BPL];
            unsigned long long near[BPL];
#pragma unroll
            for (int r = 0; r < BPL; r++) {
                const int lb = lane * BPL + r;
                stat[r] = 0; near[r] = 0;
                if (lb >= nblk) continue;
                short *cf = coef + (size_t)F::slot(lb) * STRIDE;
                const unsigned gb = first_blk + lb;
                const unsigned byi = gb / p.bx, bxi = gb - byi * p.bx;
                unsigned raw[N][N / 4];
#pragma unroll
                for (int y = 0; y < N; y++) {
                    const uint8_t *row = p.src + (size_t)(byi * N + y) * p.pitch + (size_t)bxi * N;
                    if (N == 8) {
                        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(row));
                        raw[y][0] = v.x; raw[y][N / 4 - 1] = v.y;
                    } else {
                        raw[y][0] = __ldg(reinterpret_cast<const unsigned *>(row));
                    }
                }
                float2 x2[NN / 2], y2[NN / 2];
#pragma unroll
                for (int r2 = 0; r2 < N / 2; r2++)
#pragma unroll
                    for (int k = 0; k < N; k++) {
                        // bytes -> floats by planting them in the mantissa of 2^23, then one packed subtraction of 2^23 + 128 (exact)
                        const float a = __uint_as_float(__byte_perm(raw[2 * r2][k >> 2], 0x4B000000u, 0x7650u | (unsigned)(k & 3)));
                        const float b = __uint_as_float(__byte_perm(raw[2 * r2 + 1][k >> 2], 0x4B000000u, 0x7650u | (unsigned)(k & 3)));
                        x2[r2 * N + k] = lean::add2(make_float2(a, b), make_float2(-8388736.0f, -8388736.0f));
                    }
                lean::fdct2d_packed<N>(x2, y2);
                unsigned nlo, nhi, orseg[NSEG], orbits;
                lean::quantise_block_packed<N>(y2, p.fq, p.dc_den2, p.dc_rcp, reinterpret_cast<unsigned *>(cf), nlo, nhi, orseg, orbits);
                near[r] = ((unsigned long long)nhi << 32) | nlo;
                unsigned segmask = 0;
#pragma unroll
                for (int s = 0; s < NSEG; s++) segmask |= orseg[s] ? (1u << s) : 0u;
                stat[r] = (segmask << 16) | (orbits & 0xffffu);
            }
            // ---- guard-band coefficients: the owning lanes evaluate them in the reference's precision, one per lane at a time ----
#pragma unroll
            for (int r = 0; r < BPL; r++) {
                while (__ballot_sync(0xffffffffu, near[r] != 0)) {         // warp-uniform
                    const bool have = near[r] != 0;
                    const int uv = have ? (__ffsll((long long)near[r]) - 1) : 0;
                    near[r] &= near[r] - 1;
                    const unsigned gb = first_blk + (unsigned)(lane * BPL + r);
                    const double m_uv = p.quant.m[uv];
                    int q = 0;
                    bool undecided = false;
                    if (have) undecided = !fast64_coefficient<N, false>(ex, gb, uv, m_uv, q);
                    if (__ballot_sync(0xffffffffu, undecided)) {
                        if (undecided) q = exact_coefficient<N, false>(ex, gb, uv, m_uv);
                    }
                    if (have) {
                        short *cf = coef + (size_t)F::slot(lane * BPL + r) * STRIDE;
                        const int k = p.tab->izz[uv];
                        const int q_old = cf[k];
                        if (q_old != (int)(short)q) {
                            cf[k] = (short)q;
                            stat[r] = stat_after_patch<NN>(stat[r], cf, k, q_old, (int)(short)q);
                        }
                    }
                }
            }
            // ---- RLE info -> bit_len, length, bits (Block.cpp:185-232, 371-413) ------------------------------------------------
            unsigned wl[BPL], bits[BPL], lane_bits = 0;
#pragma unroll
            for (int r = 0; r < BPL; r++) {
                const int lb = lane * BPL + r;
                wl[r] = 0x10000u; bits[r] = 0;
                if (lb >= nblk) continue;
                const unsigned st = stat[r];
                const short *cf = coef + (size_t)F::slot(lb) * STRIDE;
                int lastnz = 0, prevnz = 0;
                const unsigned orbits = st & 0xffffu;
                if (st >> 31) {
                    lastnz = (int)((st >> 16) & 0x7f);
                    prevnz = (int)((st >> 23) & 0x7f);
                } else {
                    const unsigned segmask = (st >> 16) & 0xffu;
                    if (segmask) {
                        const int lastseg = 31 - __clz(segmask);
                        const unsigned *cw = reinterpret_cast<const unsigned *>(cf) + lastseg * 4;
                        unsigned nz = 0;
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const unsigned w2 = cw[j];
                            nz |= (w2 & 0xffffu) ? (1u << (2 * j)) : 0u;
                            nz |= (w2 >> 16) ? (2u << (2 * j)) : 0u;
                        }
                        lastnz = lastseg * 8 + (32 - __clz(nz));
                        if (lastnz == NN) {                                // rare: the RLE quirk needs the previous non-zero
                            for (int k = 0; k < NN - 1; k++) if (cf[k] != 0) prevnz = k + 1;
                        }
                    }
                }
                // Block.cpp:214-219, 231: data_bits = max(max bits_needed(nz), ffs(data)), data = last non-zero index + 1
                int w = lastnz ? (33 - __clz(orbits)) : 0;
                w = max(w, dev_ffs((unsigned)lastnz));
                int len = lastnz;
                if (p.use_rle) {
                    if (lastnz == NN && prevnz != NN - 1) len = prevnz;    // Block.cpp:388-390
                } else {
                    len = NN;                                              // Block.cpp:396
                }
                wl[r] = (unsigned)w | ((unsigned)len << 8);
                bits[r] = 4u + (p.use_rle ? (unsigned)w : 0u) + (unsigned)len * (unsigned)w;
                lane_bits += bits[r];
            }
            unsigned inc = lane_bits;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const unsigned o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
            const unsigned lane_off = inc - lane_bits;
            const unsigned T = __shfl_sync(0xffffffffu, inc, 31);
            if (lane == 0) st_flag_release(&s_tot[(it % RINGR) * 32 + warp], (((it + 1) & 0xffu) << 24) | T);
            if (g < f.n_wtiles) {
                // ---- pack at tile-local alignment into the ring (the stream's first tile carries the prefix in front) ------------
                const bool with_prefix = (g == 0) && !f.append;
                const unsigned pre = with_prefix ? p.prefix_first + p.hdr.bits : 0u;
                const unsigned nbits = pre + T;
                const unsigned nw4 = ((nbits + 31) / 32 + 3) & ~3u;
                const unsigned need = F::IMG_PAD_FRONT + nw4 + F::IMG_PAD_BACK;
                // room in the ring: entries are contiguous; the live ones are [q0.start, head) (possibly wrapped once)
                for (;;) {
                    if (nq == 0) { head = 0; break; }
                    if (nq < F::QMAX) {
                        if (head >= q0.start) {                            // not wrapped
                            if (head + need <= (unsigned)F::RING_WORDS) break;
                            if (need <= q0.start) { head = 0; break; }
                        } else if (head + need <= q0.start) break;
                    }
                    drain_one(true);
                }
                const unsigned start = head;
                head += need;
                unsigned *img = ring + start + F::IMG_PAD_FRONT;
                uint4 *e4 = reinterpret_cast<uint4 *>(ring + start);
                for (unsigned i = lane; i < need / 4; i += 32) e4[i] = make_uint4(0u, 0u, 0u, 0u);
                __syncwarp();
                if (with_prefix) {
                    const unsigned hw = (pre + 31) / 32;
                    for (unsigned i = lane; i < hw; i += 32) {
                        const long long hb = (long long)i * 32 - (long long)p.prefix_first;
                        const int sh = (int)(((hb % 32) + 32) % 32);
                        const long long wi = (hb - sh) / 32;
                        const unsigned hi = (wi >= 0 && wi < kHdrWordsMax) ? p.hdr.words[wi] : 0u;
                        const unsigned lo = (wi + 1 >= 0 && wi + 1 < kHdrWordsMax) ? p.hdr.words[wi + 1] : 0u;
                        img[i] = sh ? ((hi << sh) | (lo >> (32 - sh))) : hi;
                    }
                    __syncwarp();
                }
                unsigned pos = pre + lane_off;
#pragma unroll
                for (int r = 0; r < BPL; r++) {
                    if (wl[r] & 0x10000u) break;
                    const unsigned *cw = reinterpret_cast<const unsigned *>(coef + (size_t)F::slot(lane * BPL + r) * STRIDE);
                    pack_block(img, pos, (int)(wl[r] & 0xffu), (int)((wl[r] >> 8) & 0xffu), cw, p.use_rle);
                    pos += bits[r];
                }
                __syncwarp();
                const Pending e{start, nbits, it, need};
                if (nq == 0) q0 = e; else if (nq == 1) q1 = e; else q2 = e;
                nq++;
            }
            // whatever has its base by now leaves for the stream
            while (nq > 0 && drain_one(false)) {}
        }
        while (nq > 0) drain_one(true);
    } else {
        // ================================================= prefix helper ==================================================
        // start of the blocks: behind the prefix this launch writes, or where the previous launch stopped.  (Read before this
        // CTA publishes anything: the counter is only rewritten once every CTA-tile total is known.)
        const unsigned long long start = f.append ? p.bit_counter[0] : (unsigned long long)(p.prefix_first + p.hdr.bits);
        unsigned long long base = start;
        // two independent jobs, polled in turn (neither blocks the other): publish this CTA's total of round `rp` as soon as
        // its producers have posted theirs; hand the producers the base of round `rb` as soon as every earlier CTA-tile total
        // is there (the G totals are fetched with all loads in flight, missing ones are fetched again)
        constexpr int KMAX = 8;                                            // 32 * KMAX >= gridDim.x
        unsigned rp = 0, rb = 0;
        unsigned long long have = 0;                                       // lanes' partial sums of the totals already seen for rb
        unsigned seen = 0;                                                 // bit k: this lane's k-th total of round rb is in `have`
        while (rp < rounds || rb < rounds) {
            bool progress = false;
            if (rp < rounds) {
                unsigned v = 0;
                bool ok = true;
                if (lane < NPW) {
                    v = ld_flag(&s_tot[(rp % RINGR) * 32 + lane]);
                    ok = (v >> 24) == ((rp + 1) & 0xffu);
                }
                if (__all_sync(0xffffffffu, ok)) {
                    const unsigned A = warp_sum(v & 0xffffffu);
                    if (lane == 0) st_relaxed_u64(&f.agg[rp * Gc + c], make_state(f.epoch, kFlagAggregate, (unsigned long long)A));
                    rp++;
                    progress = true;
                }
            }
            if (rb < rounds) {
                // every CTA-tile between this CTA's previous one (inclusive) and this one (exclusive)
                const unsigned ct = rb * Gc + c;
                const unsigned i0 = (rb == 0) ? 0u : ct - Gc;
                unsigned long long sv[KMAX];
#pragma unroll
                for (int k = 0; k < KMAX; k++) {
                    const unsigned i = i0 + (unsigned)lane + 32u * k;
                    sv[k] = (i < ct && !((seen >> k) & 1u)) ? ld_relaxed_u64(&f.agg[i]) : 0ull;
                }
                bool all = true;
#pragma unroll
                for (int k = 0; k < KMAX; k++) {
                    const unsigned i = i0 + (unsigned)lane + 32u * k;
                    if (i < ct && !((seen >> k) & 1u)) {
                        if ((sv[k] >> 40) == ep && ((sv[k] >> kValueBits) & 3ull) != 0ull) { have += sv[k] & kValueMask; seen |= 1u << k; }
                        else all = false;
                    }
                }
                if (__all_sync(0xffffffffu, all)) {
                    base += warp_sum64(have);
                    if (lane == 0) {
                        s_base[rb % RINGR] = base;
                        st_flag_release(&s_btag[rb % RINGR], rb + 1);
                    }
                    have = 0; seen = 0;
                    rb++;
                    progress = true;
                }
            }
            if (!progress) __nanosleep(40);
        }
    }

    // the grid's last CTA returns the ticket words to 0 for the next launch
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&f.ticket[1], 1u) == gridDim.x - 1) { f.ticket[0] = 0; f.ticket[1] = 0; __threadfence(); }
    }
}

}  // namespace

// grid size (one CTA per SM) and scan-array sizes the fused kernel needs for an image of `nblocks` blocks
void encode_fused_sizes(int N, unsigned nblocks, unsigned &n_wtiles, unsigned &n_ctatiles) {
    const unsigned TB = (N == 8) ? Fused<8>::TB : Fused<4>::TB, NPW = (N == 8) ? Fused<8>::NPW : Fused<4>::NPW;
    n_wtiles = (nblocks + TB - 1) / TB;
    n_ctatiles = (n_wtiles + NPW - 1) / NPW;
}

template <int N>
static int launch_fused_cfg(const EncodeParams &p, const FusedParams &f, int sm_count, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        IE_CUDA(cudaFuncSetAttribute(encode_fused_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Fused<N>::SMEM));
        configured = true;
    }
    const unsigned grid = std::min<unsigned>((unsigned)sm_count, f.n_ctatiles);
    encode_fused_kernel<N><<<grid, Fused<N>::THREADS, Fused<N>::SMEM, stream>>>(p, f);
    count_launch();
    IE_CUDA(cudaGetLastError());
    return IE_OK;
}

// p.scan.tile_state and p.scan.bnd must hold >= n_ctatiles entries, p.scan.ticket >= 2 words (all zero at rest)
int launch_encode_fused(int N, const EncodeParams &p, int append, int sm_count, cudaStream_t stream) {
    FusedParams f;
    encode_fused_sizes(N, p.nblocks, f.n_wtiles, f.n_ctatiles);
    f.agg = p.scan.tile_state;
    f.bnd = p.scan.bnd;
    f.ticket = p.scan.ticket;
    f.epoch = p.scan.epoch;
    f.append = append;
    f.debug = g_fused_debug.load();
    if (sm_count > 256) { set_error("more than 256 SMs: raise KMAX of the fused encoder's prefix helper"); return IE_EINVAL; }
    if (N == 8) return launch_fused_cfg<8>(p, f, sm_count, stream);
    if (N == 4) return launch_fused_cfg<4>(p, f, sm_count, stream);
    set_error("block size must be 4 or 8");
    return IE_EINVAL;
}

}  // namespace ie
