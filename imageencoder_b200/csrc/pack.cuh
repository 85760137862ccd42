// Tile-level bitstream append: CTA-wide exclusive scan of per-block bit counts, decoupled look-back across
// tiles for the global bit offset, and an output-chunk-centric gather that writes the variable-length fields
// MSB-first straight into the stream with 128-bit stores.  Replaces the reference's strictly serial
// Block::streamEncoded -> BitStreamWriter::put -> put_bit chain (Block.cpp:371-413, BitStream.cpp:61-77,
// ImageEncoder.cpp:135-138).
//
// Stream contract (all "append" kernels share it):
//   * a stream buffer is addressed in 128-bit chunks; bit 0 is the MSB of byte 0.
//   * a launch appends bits [B, B+T): B is read from a device-resident counter (+ a launch constant);
//     the chunk holding bit B is merged (OR) with what an EARLIER launch left there, interior chunks are
//     written whole, the last chunk is written zero padded (the reference's pad bits are 0: utils.hpp:443-446).
//   * inside a launch, a chunk shared by two neighbouring tiles is combined through a small hand-off record
//     (TileBoundary); the second arriver writes the chunk.  No atomics on the stream itself, no pre-zeroing.
#pragma once
#include "common.cuh"

namespace ie {

constexpr int kThreads = 128;            // small CTAs: several tiles per SM overlap each other's barriers and look-back waits
constexpr int kChunkBits = 128;

// status word of the decoupled look-back: [epoch:24][flag:2][value:38]
constexpr unsigned long long kFlagAggregate = 1ull, kFlagPrefix = 2ull;
constexpr int kValueBits = 38;
constexpr unsigned long long kValueMask = (1ull << kValueBits) - 1;

// Hand-off record for the 128-bit chunk two neighbouring tiles share: per 32-bit word one 64-bit cell that both tiles
// atomicAdd ((1 << 32) | their bits) to -- the bits of the two tiles are disjoint, so add == or; whoever finds the other's
// arrival mark already there owns the complete word, stores it and clears the cell.  Data and arrival travel in the same
// atomic, so no fence is needed.
struct TileBoundary {
    unsigned long long w[4];
};

struct ScanState {
    unsigned long long *tile_state;   // [max_tiles]
    TileBoundary *bnd;                // [max_tiles]
    unsigned *ticket;                 // dynamic tile id
    unsigned epoch;                   // differs from the previous launch on these arrays
};

#ifdef __CUDACC__

__device__ __forceinline__ unsigned long long make_state(unsigned epoch, unsigned long long flag, unsigned long long v) {
    return ((unsigned long long)(epoch & 0xFFFFFFu) << 40) | (flag << kValueBits) | (v & kValueMask);
}

// CTA-wide exclusive scan of one value per thread.  Returns the exclusive prefix; *total = CTA sum (all threads).
__device__ __forceinline__ unsigned cta_exclusive_scan(unsigned v, unsigned *s_warp /*[kThreads/32 + 1]*/, unsigned *total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        unsigned ws = (lane < kThreads / 32) ? s_warp[lane] : 0u;
        unsigned wi = ws;
#pragma unroll
        for (int d = 1; d < kThreads / 32; d <<= 1) {
            unsigned o = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += o;
        }
        if (lane < kThreads / 32) s_warp[lane] = wi - ws;
        if (lane == kThreads / 32 - 1) s_warp[kThreads / 32] = wi;
    }
    __syncthreads();
    *total = s_warp[kThreads / 32];
    return s_warp[wid] + inc - v;
}

// Decoupled look-back (single-pass chained scan), split in two so that a tile can publish its aggregate as soon as it
// knows it and resolve its prefix only when it needs it (after it has packed its bits locally):
//   tile_publish_aggregate : one thread; tile 0 publishes its inclusive prefix (base + total) directly.
//   tile_resolve_prefix    : warp 0 sums predecessor aggregates back to the nearest published prefix, 8 states per lane
//                            per round (256 predecessors in flight: one L2 round trip covers a whole wave of CTAs), then
//                            publishes this tile's inclusive prefix.  Returns the exclusive prefix to every thread.
// Tiles get their ids from a ticket, so every predecessor is already running (forward progress).
__device__ __forceinline__ void tile_publish_aggregate(const ScanState &st, unsigned tile, unsigned long long total_incl_base) {
    st_relaxed_u64(&st.tile_state[tile], make_state(st.epoch, tile == 0 ? kFlagPrefix : kFlagAggregate, total_incl_base));
}

// Warp-level look-back, nearest predecessor first: lanes examine predecessors tile-1-lane (32 at a time).  The walk stops
// at the nearest predecessor that has published an inclusive prefix; it only ever WAITS for predecessors that are nearer
// than that one (they must at least have published their aggregate).  Must be called by all 32 lanes of one warp.
// Publishes this tile's inclusive prefix and returns the exclusive prefix (in every lane).
__device__ __forceinline__ unsigned long long tile_lookback_warp(const ScanState &st, unsigned tile, unsigned long long total) {
    constexpr int R = 8;                                   // windows fetched per round trip (256 predecessors in flight)
    const int lane = threadIdx.x & 31;
    unsigned long long excl = 0;
    if (tile != 0) {
        const unsigned long long ep = (unsigned long long)(st.epoch & 0xFFFFFFu);
        int base = (int)tile - 1;
        bool done = false;
        while (!done) {
            unsigned long long sv[R];
#pragma unroll
            for (int k = 0; k < R; k++) {
                const int idx = base - 32 * k - lane;
                sv[k] = (idx >= 0) ? ld_relaxed_u64(&st.tile_state[idx]) : 0ull;
            }
#pragma unroll
            for (int k = 0; k < R; k++) {
                if (done) break;
                const int idx = base - 32 * k - lane;
                unsigned long long s = sv[k];
                unsigned pm;
                while (true) {
                    const bool valid = (idx < 0) || ((s >> 40) == ep && ((s >> kValueBits) & 3ull) != 0ull);
                    const bool is_prefix = (idx >= 0) && valid && (((s >> kValueBits) & 3ull) == kFlagPrefix);
                    pm = __ballot_sync(0xffffffffu, is_prefix);
                    const unsigned inval = __ballot_sync(0xffffffffu, !valid);
                    // this window is settled as soon as nothing nearer than its nearest prefix is still missing
                    // (or, without a prefix here, every entry is there); only then do we wait -- and only on this window
                    const unsigned nearer = pm ? ((1u << (__ffs(pm) - 1)) - 1u) : 0xffffffffu;
                    if ((inval & nearer) == 0u) break;
                    if (idx >= 0) s = ld_relaxed_u64(&st.tile_state[idx]);
                }
                const int first = pm ? (__ffs(pm) - 1) : 31;
                unsigned long long c = (idx >= 0 && lane <= first) ? (s & kValueMask) : 0ull;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) c += __shfl_down_sync(0xffffffffu, c, d);
                c = __shfl_sync(0xffffffffu, c, 0);
                excl += c;
                if (pm || base - 32 * (k + 1) < 0) done = true;
            }
            base -= 32 * R;
        }
        if (lane == 0) st_relaxed_u64(&st.tile_state[tile], make_state(st.epoch, kFlagPrefix, excl + total));
    }
    return excl;
}

__device__ __forceinline__ unsigned long long tile_resolve_prefix(const ScanState &st, unsigned tile, unsigned long long total,
                                                                  unsigned long long *s_bcast) {
    if (threadIdx.x < 32) {
        const unsigned long long excl = tile_lookback_warp(st, tile, total);
        if (threadIdx.x == 0) *s_bcast = excl;
    }
    __syncthreads();
    return *s_bcast;
}

// Both steps back to back (kernels that have nothing to overlap with the wait).
__device__ __forceinline__ unsigned long long tile_lookback(const ScanState &st, unsigned tile, unsigned long long total,
                                                            unsigned long long *s_bcast) {
    if (threadIdx.x == 0) tile_publish_aggregate(st, tile, total);
    return tile_resolve_prefix(st, tile, total, s_bcast);
}

// Tile whose bits were already assembled in shared memory, starting at bit 0 of `words` (32 stream bits per word, MSB
// first, zero beyond the tile's last bit).  The copy-out re-aligns them to the chunk grid of the global stream.
struct SmemStreamTile {
    const unsigned *words;
    unsigned nwords;             // words that hold tile bits
};
__device__ __forceinline__ uint4 gather_chunk(const SmemStreamTile &t, long long ls) {
    // chunk bits [ls, ls + 128) of the tile-local image; ls may be negative (leading bits of the tile's first chunk)
    const long long wi = ls >> 5;                          // floor
    const unsigned sh = (unsigned)(ls & 31);
    unsigned w[5];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const long long i = wi + k;
        w[k] = (i >= 0 && i < (long long)t.nwords) ? t.words[i] : 0u;
    }
    uint4 o;
    o.x = __byte_perm(__funnelshift_l(w[1], w[0], sh), 0, 0x0123);
    o.y = __byte_perm(__funnelshift_l(w[2], w[1], sh), 0, 0x0123);
    o.z = __byte_perm(__funnelshift_l(w[3], w[2], sh), 0, 0x0123);
    o.w = __byte_perm(__funnelshift_l(w[4], w[3], sh), 0, 0x0123);
    return o;
}

// Writes the tile's bits [G, G+T) into `out`.  first_tile/last_tile refer to the launch.  See the contract above.
template <int THREADS, class Tile>
__device__ __forceinline__ void tile_write_chunks_n(const Tile &t, const ScanState &st, unsigned tile, bool first_tile, bool last_tile,
                                                    unsigned long long G, unsigned T, uint8_t *out, unsigned long long out_cap_bytes,
                                                    int *d_err) {
    if (T == 0) return;
    const unsigned long long c0 = G / kChunkBits, c1 = (G + T - 1) / kChunkBits;
    const bool head_shared = (G % kChunkBits) != 0;
    const bool tail_shared = ((G + T) % kChunkBits) != 0 && !last_tile;
    for (unsigned long long c = c0 + threadIdx.x; c <= c1; c += THREADS) {
        const long long ls = (long long)(c * kChunkBits) - (long long)G;
        uint4 v = gather_chunk(t, ls);
        if ((c + 1) * 16ull > out_cap_bytes) { if (d_err) atomicExch(d_err, IE_ENOSPC); continue; }
        uint4 *dst = reinterpret_cast<uint4 *>(out) + c;
        const bool is_head = (c == c0) && head_shared;
        const bool is_tail = (c == c1) && tail_shared;
        if (is_head && first_tile) {
            // merge with what an earlier launch (header copy / previous frame) left in this chunk
            const uint4 o = *dst;
            v.x |= o.x; v.y |= o.y; v.z |= o.z; v.w |= o.w;
        } else if (is_head || is_tail) {
            // chunk shared with the neighbouring tile of this launch (every non-last tile has >= 128 bits, so a chunk is
            // never head and tail at once)
            TileBoundary *bd = &st.bnd[is_head ? tile - 1 : tile];
            unsigned *dw = reinterpret_cast<unsigned *>(dst);
            const unsigned mine[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const unsigned long long old = atomicAdd(&bd->w[i], (1ull << 32) | (unsigned long long)mine[i]);
                if ((old >> 32) == 1ull) {                       // second arriver for this word: it is complete
                    dw[i] = (unsigned)old | mine[i];
                    atomicExch(&bd->w[i], 0ull);                 // clean for the next launch
                }
            }
            continue;
        }
        *dst = v;
    }
}

template <class Tile>
__device__ __forceinline__ void tile_write_chunks(const Tile &t, const ScanState &st, unsigned tile, bool first_tile, bool last_tile,
                                                  unsigned long long G, unsigned T, uint8_t *out, unsigned long long out_cap_bytes,
                                                  int *d_err) {
    tile_write_chunks_n<kThreads>(t, st, tile, first_tile, last_tile, G, T, out, out_cap_bytes, d_err);
}

#endif  // __CUDACC__
}  // namespace ie
