// Fused stream encoder: pixels -> final stream in ONE persistent kernel (ie_set_option("encode_variant", 8)).
//
// Replaces ImageEncoder.cpp:121-138 (the parallel DCT loop + the strictly serial streamEncoded loop) like the two-kernel path
// of encode_image.cu (encode_tiles_kernel + tile_copyout_fast_kernel), with the three things its ncu captures blamed removed
// (profiles/r1_encode_v9_summary.md: 24 % of all warp time in CTA barriers, 17 % behind the exact queue alone; an 18 % copy-out
// kernel that only exists because the packed tiles round-trip through 137 MB of scratch):
//
//   * WARP-AUTONOMOUS tiles.  One CTA per SM, NPW producer warps; a producer warp owns "warp-tiles" of 32 (8x8) or 128 (4x4)
//     consecutive blocks, lane per block (per 4 blocks), and never meets a CTA-wide barrier: everything it shares with other
//     warps goes through release/acquire flags in shared memory that are set one or two tiles before they are needed.
//   * the EXACT QUEUE HAS ITS OWN WARP.  Guard-band coefficients (transform_fast.cuh) are posted to a helper warp that
//     gathers the entries of all producers, evaluates them 32 at a time (short binary64 chain first, the reference's
//     2*N*N-step chain for the true ties, exact.cuh) and patches the staged coefficients -- while the producers transform
//     their next tile.  Staging is triple-buffered per warp.
//   * FUSED COPY-OUT.  The stream position of a tile needs the bit totals of every earlier tile.  Tiles are dealt in rounds
//     (round r: CTA c takes the NPW consecutive warp-tiles of "CTA-tile" r*G + c); a second helper warp adds up the
//     published CTA-tile totals (G epoch-tagged u64 per round, L2) and hands the producers their base.  A producer packs
//     tile t two iterations after it transformed it, straight at its FINAL alignment in a 2 KiB shared-memory image, and
//     stores the 128-bit chunks to the stream; by then every total it needs was published long ago, so nobody waits in
//     practice (the waits are real spins, correctness does not depend on the timing).  No scratch round trip, no second
//     kernel.  Chunks shared by two warp-tiles go through the fence-free hand-off records of pack.cuh.
//
// Per warp-tile iteration `it`:  phase1(it) load / transform / quantise -> coef[it % 3], guard-band entries -> helper
//                                mid(it-1)  patches done? -> RLE info, bit_len, length, bit count -> total posted
//                                late(it-2) base known?   -> pack at final alignment -> chunks to the stream
#include "encode_image.cuh"
#include "transform_fast.cuh"
#include "exact.cuh"

namespace ie {

namespace {

constexpr size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

template <int N>
struct Fused {
    static constexpr int BPL = (N == 8) ? 1 : 4;          // blocks per lane
    static constexpr int TB = 32 * BPL;                   // blocks per warp-tile
    static constexpr int NN = N * N;
    static constexpr int STRIDE = NN + 2;                 // halfwords per block in the staging area (odd word count: bank spread)
    static constexpr int NSEG = NN / 8;
    static constexpr int NPW = (N == 8) ? 14 : 12;        // producer warps per CTA
    static constexpr int NBUF = 3;                        // staging buffers per producer warp
    static constexpr int QCAP = 16;                       // guard-band entries per tile handed to the helper (more: done in place)
    static constexpr int IMG_WORDS = 512;                 // per-warp image of stream bits (2 KiB)
    static constexpr int IMG_BITS = IMG_WORDS * 32;
    static constexpr int RING = 8;                        // rounds a CTA-level slot stays valid
    static constexpr int THREADS = (NPW + 2) * 32;
    static constexpr size_t COEF_BYTES = (size_t)TB * STRIDE * 2;
    // per producer warp
    static constexpr size_t W_COEF = 0;
    static constexpr size_t W_STAT = W_COEF + NBUF * COEF_BYTES;          // u32 [NBUF][TB]: RLE info, later bit_len | length << 8
    static constexpr size_t W_ENT = W_STAT + (size_t)NBUF * TB * 4;        // u32 [NBUF][QCAP]
    static constexpr size_t W_IMG = align16(W_ENT + (size_t)NBUF * QCAP * 4);
    static constexpr size_t W_BYTES = align16(W_IMG + (size_t)(IMG_WORDS + 4) * 4);
    // CTA level, behind the NPW warp regions
    static constexpr size_t C_TOT = (size_t)NPW * W_BYTES;                 // u32 [RING][16]: tag << 24 | warp-tile bits
    static constexpr size_t C_BASE = C_TOT + RING * 16 * 4;                // u64 [RING]
    static constexpr size_t C_BTAG = C_BASE + RING * 8;                    // u32 [RING]
    static constexpr size_t C_PEND = C_BTAG + RING * 4;                    // u32 [NPW * NBUF]: entries waiting for the helper
    static constexpr size_t C_FIRST = C_PEND + (size_t)NPW * NBUF * 4;     // u32 [NPW * NBUF]: first block of the tile in that buffer
    static constexpr size_t C_MISC = C_FIRST + (size_t)NPW * NBUF * 4;     // u32 [4]: CTA id, producers done
    // chunks two warp-tiles of the same CTA-tile share: [RING][NPW - 1] x { tail of the left tile, head of the right tile, flag }
    static constexpr size_t C_HAND = align16(C_MISC + 16);
    static constexpr size_t HAND_BYTES = 48;
    static constexpr size_t SMEM = align16(C_HAND + (size_t)RING * (NPW - 1) * HAND_BYTES);
    // staging slot of block lb: lanes must hit different banks when they store pair words (lane stride = BPL blocks)
    __device__ static __forceinline__ int slot(int lb) { return (BPL == 1) ? lb : ((lb & (BPL - 1)) * 32 + lb / BPL); }
};
static_assert(Fused<8>::SMEM <= 232448 && Fused<4>::SMEM <= 232448, "shared memory budget of one CTA per SM");

struct FusedParams {
    unsigned n_wtiles;                // warp-tiles of the image
    unsigned n_ctatiles;              // groups of NPW warp-tiles
    unsigned long long *agg;          // [n_ctatiles] epoch-tagged CTA-tile totals (ScanState::tile_state)
    TileBoundary *bnd;                // [n_wtiles]
    unsigned *ticket;                 // [2]: dynamic CTA id, CTAs finished (both return to 0 when the grid ends)
    unsigned epoch;
    int append;                       // 1: the stream continues at *bit_counter (no prefix)
    int debug;                        // timing experiments only (ie_set_option("fused_debug")): 1 = do not wait for the base
                                      // (tiles land at made-up positions: WRONG output), 2 = do not wait for the exact helper
};

__device__ __forceinline__ unsigned ld_acq_cta(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.cta.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_rel_cta(unsigned *p, unsigned v) { asm volatile("st.release.cta.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

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

// one guard-band coefficient, in place (queue overflow: adversarial content only)
template <int N>
__device__ __noinline__ bool resolve_in_place(const ExactCtx ex, unsigned gb, int uv, double m_uv, short *cf) {
    int q;
    if (!fast64_coefficient<N, false>(ex, gb, uv, m_uv, q)) q = exact_coefficient<N, false>(ex, gb, uv, m_uv);
    const int k = ex.tab->izz[uv];
    if (cf[k] != (short)q) { cf[k] = (short)q; return true; }
    return false;
}

// ---- phase 1: load, transform, quantise (the arithmetic of encode_tiles_kernel's variant 2, transform_fast.cuh) ----------
template <int N>
__device__ __forceinline__ void fused_phase1(const EncodeParams &p, const ExactCtx &ex, unsigned first_blk, int nblk, short *coef,
                                             unsigned *stat, unsigned *ent, unsigned *pend, unsigned *firstw, int lane) {
    using F = Fused<N>;
    constexpr int NN = F::NN, BPL = F::BPL, STRIDE = F::STRIDE, NSEG = F::NSEG;
    unsigned nq = 0;                                     // entries posted so far (warp-uniform)
#pragma unroll 1
    for (int r = 0; r < BPL; r++) {
        const int lb = lane * BPL + r;
        const bool active = lb < nblk;
        unsigned long long near = 0;
        short *cf = coef + (size_t)F::slot(lb) * STRIDE;
        const unsigned gb = first_blk + lb;
        if (active) {
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
            near = ((unsigned long long)nhi << 32) | nlo;
            unsigned segmask = 0;
#pragma unroll
            for (int s = 0; s < NSEG; s++) segmask |= orseg[s] ? (1u << s) : 0u;
            stat[F::slot(lb)] = (segmask << 16) | (orbits & 0xffffu);
        }
        // guard-band coefficients -> the helper warp's queue (warp-uniform branch; almost every tile has a few)
        if (__ballot_sync(0xffffffffu, near != 0)) {
            const unsigned n = (unsigned)__popcll(near);
            unsigned inc = n;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const unsigned o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
            unsigned pos = nq + inc - n;
            nq += __shfl_sync(0xffffffffu, inc, 31);
            bool dirty = false;
            while (near) {
                const int uv = __ffsll((long long)near) - 1;
                near &= near - 1;
                if (pos < (unsigned)F::QCAP) ent[pos] = ((unsigned)lb << 8) | (unsigned)uv;
                else dirty |= resolve_in_place<N>(ex, gb, uv, p.quant.m[uv], cf);
                pos++;
            }
            if (dirty) stat[F::slot(lb)] = stats_full<NN>(cf);
        }
    }
    __syncwarp();
    if (lane == 0) {
        *firstw = first_blk;
        st_rel_cta(pend, min(nq, (unsigned)F::QCAP));
    }
}

// ---- mid: RLE info -> bit_len, length, bits (Block.cpp:185-232, 371-413) ---------------------------------------------------
template <int N>
__device__ __forceinline__ unsigned fused_mid(const EncodeParams &p, int nblk, const short *coef, unsigned *stat, int lane) {
    using F = Fused<N>;
    constexpr int NN = F::NN, BPL = F::BPL, STRIDE = F::STRIDE;
    unsigned lane_bits = 0;
#pragma unroll 1
    for (int r = 0; r < BPL; r++) {
        const int lb = lane * BPL + r;
        const int s = F::slot(lb);
        if (lb >= nblk) { stat[s] = 0x10000u; continue; }                  // no block here: 0 bits
        const unsigned st = stat[s];
        const short *cf = coef + (size_t)s * STRIDE;
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
                if (lastnz == NN) {                                        // rare: the RLE quirk needs the previous non-zero
                    for (int k = 0; k < NN - 1; k++) if (cf[k] != 0) prevnz = k + 1;
                }
            }
        }
        // Block.cpp:214-219, 231: data_bits = max(max bits_needed(nz), ffs(data)), data = last non-zero index + 1
        int w = lastnz ? (33 - __clz(orbits)) : 0;
        w = max(w, dev_ffs((unsigned)lastnz));
        int len = lastnz;
        if (p.use_rle) {
            if (lastnz == NN && prevnz != NN - 1) len = prevnz;            // Block.cpp:388-390
        } else {
            len = NN;                                                      // Block.cpp:396
        }
        stat[s] = (unsigned)w | ((unsigned)len << 8);
        lane_bits += 4u + (p.use_rle ? (unsigned)w : 0u) + (unsigned)len * (unsigned)w;
    }
    return lane_bits;
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

// the same between two warps of one CTA (13 of 14 boundaries): both sides leave their half in shared memory, the second
// arriver (shared-memory atomic, no L2 round trip) merges and stores the chunk
struct SmemHandoff {
    uint4 half[2];               // [0] tail of the left tile, [1] head of the right tile
    unsigned flag;
    unsigned pad[3];
};
__device__ __forceinline__ void handoff_chunk_smem(SmemHandoff *h, int side, uint4 *dst, uint4 v) {
    h->half[side] = v;
    __threadfence_block();
    if (atomicAdd(&h->flag, 1u) == 1u) {
        __threadfence_block();
        const uint4 o = h->half[side ^ 1];
        v.x |= o.x; v.y |= o.y; v.z |= o.z; v.w |= o.w;
        *dst = v;
        h->flag = 0u;            // next use: RING rounds later
    }
}

// ---- late: pack at the final alignment, chunks to the stream ------------------------------------------------------------------
// G = stream bit of the tile's first bit.  The tile is cut into segments of whole lanes whose bits fit the image (one segment
// unless the content is noise-like); a chunk two segments share is carried over inside the image.
template <int N>
__device__ __forceinline__ void fused_late(const EncodeParams &p, const FusedParams &f, unsigned g, unsigned ct, int warp, SmemHandoff *hand,
                                           unsigned long long G, const short *coef, const unsigned *stat, unsigned *img, int lane) {
    using F = Fused<N>;
    constexpr int BPL = F::BPL, STRIDE = F::STRIDE;
    unsigned wl[BPL], bits[BPL], lane_bits = 0;
#pragma unroll
    for (int r = 0; r < BPL; r++) {
        wl[r] = stat[F::slot(lane * BPL + r)];
        const unsigned w = wl[r] & 0xffu, len = (wl[r] >> 8) & 0xffu;
        bits[r] = (wl[r] & 0x10000u) ? 0u : 4u + (p.use_rle ? w : 0u) + len * w;
        lane_bits += bits[r];
    }
    unsigned inc = lane_bits;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
    const unsigned lane_off = inc - lane_bits;
    const unsigned T = __shfl_sync(0xffffffffu, inc, 31);
    const bool first_tile = (g == 0), last_tile = (g + 1 == f.n_wtiles);
    uint4 *img4 = reinterpret_cast<uint4 *>(img);
    uint4 *out4 = reinterpret_cast<uint4 *>(p.out);
    const unsigned long long cap_chunks = p.out_cap / 16ull;

    int a = 0;                                   // first lane of the segment
    unsigned carry_words = 0;                    // words of the image that already hold stream bits
    unsigned long long O;                        // stream bit of image bit 0 (chunk aligned)
    bool head_pending;                           // the segment's first chunk still has to be merged with foreign bits
    if (first_tile && !f.append) {
        // the stream's prefix (prefix_first zero bits + header, ImageEncoder.cpp:84-94) is the start of the first image
        O = 0;
        const unsigned total = p.prefix_first + p.hdr.bits;               // == G
        carry_words = (total + 31) / 32;
        for (unsigned i = lane; i < carry_words; i += 32) {
            const long long hb = (long long)i * 32 - (long long)p.prefix_first;
            const int sh = (int)(((hb % 32) + 32) % 32);
            const long long wi = (hb - sh) / 32;
            const unsigned hi = (wi >= 0 && wi < kHdrWordsMax) ? p.hdr.words[wi] : 0u;
            const unsigned lo = (wi + 1 >= 0 && wi + 1 < kHdrWordsMax) ? p.hdr.words[wi + 1] : 0u;
            img[i] = sh ? ((hi << sh) | (lo >> (32 - sh))) : hi;
        }
        head_pending = false;
    } else {
        O = G & ~127ull;
        head_pending = (G & 127ull) != 0;
    }
    while (a < 32) {
        const unsigned lo_a = __shfl_sync(0xffffffffu, lane_off, a);
        const unsigned long long S0 = G + lo_a;
        const unsigned room = (unsigned)F::IMG_BITS - (unsigned)(S0 - O);
        // lanes a..b-1: their bits end inside the image (monotone in the lane index; lane a always fits: a lane holds at
        // most 1104 bits and at most one chunk + the prefix lie in front of it)
        const unsigned fit = __ballot_sync(0xffffffffu, lane >= a && (lane_off + lane_bits - lo_a) <= room);
        const int b = 32 - __clz(fit);
        const unsigned long long S1 = G + ((b < 32) ? __shfl_sync(0xffffffffu, lane_off, b & 31) : T);
        const unsigned nbits = (unsigned)(S1 - O);
        const unsigned nchunks = (nbits + 127) / 128;
        // zero everything behind the words that already hold stream bits (header / carried chunk)
        for (unsigned i = (carry_words + 3) / 4 + lane; i < nchunks; i += 32) img4[i] = make_uint4(0u, 0u, 0u, 0u);
        if ((carry_words & 3u) && lane < 4 && lane >= (int)(carry_words & 3u)) img[(carry_words & ~3u) + lane] = 0u;
        __syncwarp();
        if (lane >= a && lane < b) {
            unsigned pos = (unsigned)(S0 - O) + (lane_off - lo_a);
#pragma unroll
            for (int r = 0; r < BPL; r++) {
                if (wl[r] & 0x10000u) break;
                const int w = (int)(wl[r] & 0xffu), len = (int)((wl[r] >> 8) & 0xffu);
                const unsigned *cw = reinterpret_cast<const unsigned *>(coef + (size_t)F::slot(lane * BPL + r) * STRIDE);
                pack_block(img, pos, w, len, cw, p.use_rle);
                pos += bits[r];
            }
        }
        __syncwarp();
        const bool more = b < 32;
        const bool tail_partial = (S1 & 127ull) != 0;
        const unsigned long long c0 = O / 128ull;
        for (unsigned c = lane; c < nchunks; c += 32) {
            const uint4 wv = img4[c];
            uint4 v;
            v.x = __byte_perm(wv.x, 0, 0x0123); v.y = __byte_perm(wv.y, 0, 0x0123);
            v.z = __byte_perm(wv.z, 0, 0x0123); v.w = __byte_perm(wv.w, 0, 0x0123);
            const bool is_tail = (c + 1 == nchunks) && tail_partial;
            if (is_tail && more) continue;                                 // carried into the next segment's image
            if (c0 + c >= cap_chunks) { if (p.err) atomicExch(p.err, IE_ENOSPC); continue; }
            uint4 *dst = out4 + c0 + c;
            if (c == 0 && head_pending) {
                if (first_tile) {                                          // append: merge with what the earlier launch left here
                    const uint4 o = *dst;
                    v.x |= o.x; v.y |= o.y; v.z |= o.z; v.w |= o.w;
                    *dst = v;
                } else if (warp > 0) {
                    handoff_chunk_smem(&hand[warp - 1], 1, dst, v);      // shared with the previous warp of this CTA
                } else {
                    handoff_chunk(&f.bnd[ct - 1], dst, v);               // shared with the previous CTA-tile
                }
            } else if (is_tail && !last_tile) {
                if (warp + 1 < F::NPW) handoff_chunk_smem(&hand[warp], 0, dst, v);
                else handoff_chunk(&f.bnd[ct], dst, v);
            } else {
                *dst = v;
            }
        }
        head_pending = false;
        if (more) {
            // the partially filled last chunk becomes the first chunk of the next segment's image
            __syncwarp();
            unsigned keep = 0;
            if (tail_partial && lane < 4) keep = img[(nchunks - 1) * 4 + lane];
            __syncwarp();
            if (tail_partial && lane < 4) img[lane] = keep;
            carry_words = tail_partial ? 4u : 0u;
            O = S1 & ~127ull;
        }
        a = b;
    }
    if (last_tile && lane == 0) {
        p.bit_counter[0] = G + T;
        if (p.out_bits) p.out_bits[0] = G + T;
    }
}

template <int N>
__global__ void __launch_bounds__(Fused<N>::THREADS, 1) encode_fused_kernel(const EncodeParams p, const FusedParams f) {
    using F = Fused<N>;
    constexpr int NPW = F::NPW, NBUF = F::NBUF, RING = F::RING;
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned *s_tot = reinterpret_cast<unsigned *>(smem + F::C_TOT);
    unsigned long long *s_base = reinterpret_cast<unsigned long long *>(smem + F::C_BASE);
    unsigned *s_btag = reinterpret_cast<unsigned *>(smem + F::C_BTAG);
    unsigned *s_pend = reinterpret_cast<unsigned *>(smem + F::C_PEND);
    unsigned *s_first = reinterpret_cast<unsigned *>(smem + F::C_FIRST);
    unsigned *s_misc = reinterpret_cast<unsigned *>(smem + F::C_MISC);
    SmemHandoff *s_hand = reinterpret_cast<SmemHandoff *>(smem + F::C_HAND);
    static_assert(sizeof(SmemHandoff) == F::HAND_BYTES, "hand-off slot size");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // dynamic CTA id: every CTA with a smaller id is running (a spinning CTA only ever waits for smaller ids)
    if (threadIdx.x == 0) { s_misc[0] = atomicAdd(&f.ticket[0], 1u); s_misc[1] = 0; }
    for (int i = threadIdx.x; i < RING * 16; i += blockDim.x) s_tot[i] = 0;
    for (int i = threadIdx.x; i < RING; i += blockDim.x) s_btag[i] = 0;
    for (int i = threadIdx.x; i < NPW * NBUF; i += blockDim.x) s_pend[i] = 0;
    for (int i = threadIdx.x; i < RING * (NPW - 1); i += blockDim.x) s_hand[i].flag = 0;
    __syncthreads();                                                       // the only CTA-wide barrier of the kernel
    const unsigned c = s_misc[0], Gc = gridDim.x;
    const unsigned rounds = (f.n_ctatiles > c) ? (f.n_ctatiles - c + Gc - 1) / Gc : 0u;
    const unsigned long long ep = (unsigned long long)(f.epoch & 0xFFFFFFu);

    ExactCtx ex;
    ex.src = p.src; ex.ref = nullptr; ex.res_coord = nullptr; ex.tab = p.tab; ex.pitch = p.pitch; ex.bx = p.bx; ex.mbx = 0;

    if (warp < NPW) {
        // ================================================= producer warps =================================================
        unsigned char *wbase = smem + (size_t)warp * F::W_BYTES;
        unsigned *img = reinterpret_cast<unsigned *>(wbase + F::W_IMG);
        for (unsigned it = 0; it < rounds + 2; it++) {
            if (it + 1 < rounds) {
                // next tile's pixels on their way into L2 while this one is transformed: one lane per 128-byte line of the
                // tile's pixel rows (a hint only)
                const unsigned gn = ((it + 1) * Gc + c) * NPW + warp;
                constexpr int LPR = F::TB * N / 128;                       // 128-byte lines per pixel row of a warp-tile
                const unsigned nb = gn * F::TB + (unsigned)(lane % LPR) * (128u / N);
                if (gn < f.n_wtiles && lane < N * LPR && nb < p.nblocks) {
                    const unsigned byi = nb / p.bx, bxi = nb - byi * p.bx;
                    const uint8_t *row = p.src + (size_t)(byi * N + lane / LPR) * p.pitch + (size_t)bxi * N;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(row));
                }
            }
            if (it < rounds) {
                const unsigned g = (it * Gc + c) * NPW + warp;
                const int buf = (int)(it % NBUF);
                const unsigned first_blk = g * F::TB;
                const int nblk = (g < f.n_wtiles) ? (int)min((unsigned)F::TB, p.nblocks - first_blk) : 0;
                fused_phase1<N>(p, ex, first_blk, nblk, reinterpret_cast<short *>(wbase + F::W_COEF + buf * F::COEF_BYTES),
                                reinterpret_cast<unsigned *>(wbase + F::W_STAT) + buf * F::TB,
                                reinterpret_cast<unsigned *>(wbase + F::W_ENT) + buf * F::QCAP, &s_pend[warp * NBUF + buf],
                                &s_first[warp * NBUF + buf], lane);
            }
            if (it >= 1 && it - 1 < rounds) {
                const unsigned r = it - 1;
                const unsigned g = (r * Gc + c) * NPW + warp;
                const int buf = (int)(r % NBUF);
                const int nblk = (g < f.n_wtiles) ? (int)min((unsigned)F::TB, p.nblocks - g * F::TB) : 0;
                if (!(f.debug & 2)) while (ld_acq_cta(&s_pend[warp * NBUF + buf]) != 0) __nanosleep(40);      // the helper has patched this tile
                const unsigned lane_bits = fused_mid<N>(p, nblk, reinterpret_cast<const short *>(wbase + F::W_COEF + buf * F::COEF_BYTES),
                                                        reinterpret_cast<unsigned *>(wbase + F::W_STAT) + buf * F::TB, lane);
                const unsigned T = warp_sum(lane_bits);
                if (lane == 0) st_rel_cta(&s_tot[(r % RING) * 16 + warp], (((r + 1) & 0xffu) << 24) | T);
            }
            if (it >= 2 && it - 2 < rounds) {
                const unsigned r = it - 2;
                const unsigned g = (r * Gc + c) * NPW + warp;
                const int buf = (int)(r % NBUF);
                if (g < f.n_wtiles) {
                    if (!(f.debug & 1)) while (ld_acq_cta(&s_btag[r % RING]) != r + 1) __nanosleep(40);
                    unsigned long long G = (f.debug & 1) ? (unsigned long long)g * 6784ull + 1024ull : s_base[r % RING];
                    unsigned t = 0;
                    if (lane < warp && !(f.debug & 1)) {
                        while (((t = ld_acq_cta(&s_tot[(r % RING) * 16 + lane])) >> 24) != ((r + 1) & 0xffu)) __nanosleep(32);
                        t &= 0xffffffu;
                    }
                    G += warp_sum(t);
                    fused_late<N>(p, f, g, r * Gc + c, warp, s_hand + (r % RING) * (NPW - 1), G, reinterpret_cast<const short *>(wbase + F::W_COEF + buf * F::COEF_BYTES),
                                  reinterpret_cast<const unsigned *>(wbase + F::W_STAT) + buf * F::TB, img, lane);
                }
            }
        }
        __syncwarp();
        if (lane == 0) atomicAdd(&s_misc[1], 1u);
    } else if (warp == NPW) {
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
                    v = ld_acq_cta(&s_tot[(rp % RING) * 16 + lane]);
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
                        s_base[rb % RING] = base;
                        st_rel_cta(&s_btag[rb % RING], rb + 1);
                    }
                    have = 0; seen = 0;
                    rb++;
                    progress = true;
                }
            }
            if (!progress) __nanosleep(40);
        }
    } else {
        // ================================================= exact helper ===================================================
        // gathers the guard-band entries the producers posted (whole slots, up to 32 entries per pass), evaluates them in the
        // reference's precision and patches the staged coefficients; a patched block gets its RLE info recomputed
        constexpr int NSLOT = NPW * NBUF;
        static_assert(NSLOT <= 64, "two poll words per lane");
        for (;;) {
            const unsigned n0 = (lane < NSLOT) ? ld_acq_cta(&s_pend[lane]) : 0u;
            const unsigned n1 = (lane + 32 < NSLOT) ? ld_acq_cta(&s_pend[lane + 32]) : 0u;
            const unsigned m0 = __ballot_sync(0xffffffffu, n0 != 0), m1 = __ballot_sync(0xffffffffu, n1 != 0);
            if (!(m0 | m1)) {
                if (ld_acq_cta(&s_misc[1]) == (unsigned)NPW) break;        // every producer is done: nothing can arrive any more
                __nanosleep(200);
                continue;
            }
            unsigned count = 0, sel0 = 0, sel1 = 0;
            int my_slot = -1, my_j = 0;
            for (int half = 0; half < 2; half++) {
                unsigned m = half ? m1 : m0;
                while (m) {
                    const int sl = __ffs((int)m) - 1;
                    m &= m - 1;
                    const unsigned n = __shfl_sync(0xffffffffu, half ? n1 : n0, sl);
                    if (count + n > 32u) continue;
                    if ((unsigned)lane >= count && (unsigned)lane < count + n) { my_slot = sl + 32 * half; my_j = lane - (int)count; }
                    count += n;
                    if (half) sel1 |= 1u << sl; else sel0 |= 1u << sl;
                }
            }
            bool dirty = false;
            short *cf = nullptr;
            unsigned *stat_word = nullptr;
            if (my_slot >= 0) {
                const int w = my_slot / NBUF, buf = my_slot % NBUF;
                unsigned char *wb = smem + (size_t)w * F::W_BYTES;
                const unsigned e = (reinterpret_cast<const unsigned *>(wb + F::W_ENT) + buf * F::QCAP)[my_j];
                const int lb = (int)(e >> 8), uv = (int)(e & 0xff);
                const unsigned gb = s_first[my_slot] + (unsigned)lb;
                cf = reinterpret_cast<short *>(wb + F::W_COEF + buf * F::COEF_BYTES) + (size_t)F::slot(lb) * F::STRIDE;
                stat_word = reinterpret_cast<unsigned *>(wb + F::W_STAT) + buf * F::TB + F::slot(lb);
                int q;
                if (!fast64_coefficient<N, false>(ex, gb, uv, p.quant.m[uv], q)) q = exact_coefficient<N, false>(ex, gb, uv, p.quant.m[uv]);
                const int k = p.tab->izz[uv];
                if (cf[k] != (short)q) { cf[k] = (short)q; dirty = true; }
            }
            __syncwarp();                        // every patch of this pass is in place before a block's RLE info is recomputed
            if (dirty) *stat_word = stats_full<F::NN>(cf);
            __syncwarp();
            if (lane < 32 && ((sel0 >> lane) & 1u)) st_rel_cta(&s_pend[lane], 0u);
            if ((sel1 >> lane) & 1u) st_rel_cta(&s_pend[lane + 32], 0u);
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

std::atomic<int> g_fused_debug{0};

bool encode_fused_eligible(const EncodeParams &p, unsigned images) {
    return images == 1 && !p.bits_only && p.phase == 0 && p.ref == nullptr;
}

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

// p.scan.tile_state must hold >= n_ctatiles entries, p.scan.bnd >= n_wtiles, p.scan.ticket >= 2 words (all zero at rest)
int launch_encode_fused(int N, const EncodeParams &p, int append, int sm_count, cudaStream_t stream) {
    FusedParams f;
    encode_fused_sizes(N, p.nblocks, f.n_wtiles, f.n_ctatiles);
    f.agg = p.scan.tile_state;
    f.bnd = p.scan.bnd;
    f.ticket = p.scan.ticket;
    f.epoch = p.scan.epoch;
    f.append = append;
    f.debug = g_fused_debug.load();
    if (N == 8) return launch_fused_cfg<8>(p, f, sm_count, stream);
    if (N == 4) return launch_fused_cfg<4>(p, f, sm_count, stream);
    set_error("block size must be 4 or 8");
    return IE_EINVAL;
}

}  // namespace ie
