// Parallel discovery of block offsets in the marker-less stream.
//
// A block's size is in its own header (Block.cpp:443-444: bit_len(4) [length(bit_len)] length x bit_len), so block k's
// first bit depends on every earlier block -- the reference reads the stream strictly in sequence (ImageDecoder.cpp:89-92).
// Here the chain is resolved with transfer functions over spans of the stream, measured from the first block's bit B0:
//   * a block that starts before a span boundary ends at most E = 4 + 16 + 16 N^2 bits after it, so the chain can enter
//     a span only at offsets [0, E).  A span's transfer function maps every possible entry offset to (offset at which the
//     chain enters the next span, number of blocks started inside).
//   * parse_group_tables : one CTA per group (kSubsPerGroup sub-spans, short ones first).  All E hypothetical chains are
//                          advanced sub-span by sub-span; at every sub-span boundary identical chains are merged (only one
//                          walk per DISTINCT entry offset), and real streams merge to a single chain almost at once.
//   * parse_super_tables : composes kSuper group tables for all E entries in parallel (E independent lookup chains).
//   * parse_top_walk     : one thread walks the super-groups from the true entry (offset 0).
//   * parse_down_super   : one thread per super-group hands every group its true entry and first block index.
//   * parse_emit_offsets : one thread per group walks its true chain and writes block_off[].
// Exact for any stream (no statistical self-synchronisation assumption).  Reads past the end of the stream give 0 bits and
// do not advance (BitStream.cpp:17-20): a chain that reaches the end is DEAD and every remaining block starts at `total`.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>

#include "decode_image.cuh"
#include "stage.cuh"

namespace ie {

// 1 = the speculative parse's verdict is ignored and the exact transfer-function path produces the offsets (tests)
std::atomic<int> g_parse_variant{0};
// speculative grid of 4x4 streams: 0 = 2048-bit groups with a 4096-bit lead-in, 1 = 1024-bit groups with a 2048-bit lead-in
std::atomic<int> g_parse_grid4{-1};             // -1 = by kind of stream (g4_bits)

// Sub-span boundaries inside a group: short spans first, so that the E hypothetical chains of a group are merged after
// a block or two instead of each being walked through kilobits of stream.
constexpr int kSubsPerGroup = 9;
__constant__ int c_sub_bounds[kSubsPerGroup + 1] = {0, 256, 512, 1024, 1536, 2048, 3072, 4096, 6144, 8192};
constexpr int kGroupBits = 8192;                         // 1 KiB of stream per group
constexpr int kSuper = 128;                              // groups per super-group (128 KiB of stream)
constexpr unsigned kDead = 0xFFFFu;

struct ParseParams {
    const uint8_t *enc;
    const unsigned long long *enc_bits;     // device: stream size in bits
    const unsigned long long *start;        // device: cursor / first-block bit (skip_bits is added)
    unsigned skip_bits;
    int NN, use_rle, E;
    unsigned ngroups, nsuper, nblocks;
    uint2 *group_table;                     // [ngroups][E] (exit offset | kDead, blocks)
    uint2 *super_table;                     // [nsuper][E]
    uint2 *super_entry;                     // [nsuper]  (entry offset | kDead, first block index)
    uint2 *group_entry;                     // [ngroups]
    unsigned long long *block_off;          // [nblocks + 1]
    unsigned long long *cursor_out;         // device, may be NULL: receives the bit after the last block (parse_commit_cursor)
    unsigned long long *cursor_next;        // scratch: where the emit kernels leave it (cursor_out aliases `start`, which they read)
    int *err;
    // speculative path
    unsigned nspec;                         // groups of the speculative grid
    uint2 *spec_entry;                      // [nspec] (entry offset | kDead, -)
    uint2 *spec_exit;                       // [nspec] (exit offset | kDead, blocks started in the group)
    unsigned *spec_flags;                   // [0] CTA ticket of parse_spec_check, [1] spec_ok, [2] unused, [3] first inconsistent group
    unsigned *walk_base;                    // [nwalk] block count per walk CTA, then (in place) its exclusive scan
    int force_exact;                        // ie_set_option("parse_variant", 1): ignore the speculation, take the exact path
    // sharded decode of one stream (ie_decode_image_shard_*): a rank's walk launch covers the CTAs from walk_cta0 on; the
    // emit kernel only has to produce block_off[emit_lo .. emit_hi] (emit_hi = 0: everything)
    unsigned walk_cta0, emit_lo, emit_hi;
};

__device__ __forceinline__ unsigned parse_read_bits(const uint8_t *__restrict__ s, unsigned long long total_bits, unsigned long long p, int n) {
    if (n == 0) return 0u;
    const unsigned long long nbytes = (total_bits + 7) >> 3, b = p >> 3;
    unsigned v = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) v = (v << 8) | ((b + i < nbytes) ? (unsigned)__ldg(s + b + i) : 0u);
    return (v >> (32 - (int)(p & 7) - n)) & ((1u << n) - 1u);
}

// size in bits of the block whose header is at absolute bit p (p < total); kBadBlock if the length field exceeds N*N.
// bit_len (4 bits) and the length field (<= 15 bits) fit one 32-bit window at any bit alignment (7 + 4 + 15 <= 32).
// A hypothetical chain that starts at a wrong offset reads garbage headers; one with length > N*N cannot be the real
// chain of a valid stream (Block.cpp:185-232 never produces it) and is dropped (DEAD).  If the REAL chain contains such
// a block the stream is malformed: parse_emit_offsets reports IE_EFORMAT (the reference indexes out of bounds there).
constexpr unsigned kBadBlock = 0xFFFFFFFFu;
__device__ __forceinline__ unsigned block_bits_at(const uint8_t *s, unsigned long long total, unsigned long long p, int NN, int rle) {
    // 32 stream bits starting at p from two aligned words (the stream buffer is 4-byte aligned and padded; bits past
    // `total` are forced to zero, BitStream.cpp:17-20)
    const unsigned *wp = reinterpret_cast<const unsigned *>(s) + (p >> 5);
    const unsigned long long nwords = (total + 31) >> 5;
    const unsigned w0 = __byte_perm(__ldg(wp), 0, 0x0123);
    const unsigned w1 = ((p >> 5) + 1 < nwords) ? __byte_perm(__ldg(wp + 1), 0, 0x0123) : 0u;
    unsigned v = __funnelshift_l(w1, w0, (unsigned)(p & 31));
    if (p + 32 > total) v &= ~((total - p >= 32) ? 0u : (0xFFFFFFFFu >> (unsigned)(total - p)));
    const unsigned w = v >> 28;
    unsigned len = (unsigned)NN;
    if (rle) len = w ? ((v << 4) >> (32 - w)) : 0u;
    if (len > (unsigned)NN) return kBadBlock;
    return 4u + (rle ? w : 0u) + len * w;
}

// The exact path as ONE kernel (parse_exact_kernel below): its five phases used to be five launches, each returning at once in
// the normal case (speculation verified) -- 10 us of launch gaps per image for nothing.  Phases are separated by a grid barrier
// (every CTA is resident: the grid is sized to the SM count).
__device__ __forceinline__ void parse_group_tables(const ParseParams &p) {
    extern __shared__ unsigned s_parse[];
    const int E = p.E;
    unsigned *s_cnt = s_parse;                                            // [E]
    unsigned *s_memo = s_cnt + E;                                         // [E] exit << 16 | blocks
    unsigned short *s_cur = reinterpret_cast<unsigned short *>(s_memo + E);   // [E]
    unsigned char *s_need = reinterpret_cast<unsigned char *>(s_cur + E + (E & 1));   // [E]
    const unsigned long long total = *p.enc_bits;
    const unsigned long long B0 = *p.start + p.skip_bits;
    for (unsigned g = blockIdx.x; g < p.ngroups; g += gridDim.x) {
    const unsigned long long g_start = B0 + (unsigned long long)g * kGroupBits;
    __syncthreads();
    for (int e = threadIdx.x; e < E; e += blockDim.x) { s_cur[e] = (unsigned short)e; s_cnt[e] = 0; }
    __syncthreads();
    for (int sub = 0; sub < kSubsPerGroup; sub++) {
        const unsigned long long s_start = g_start + (unsigned long long)c_sub_bounds[sub];
        for (int e = threadIdx.x; e < E; e += blockDim.x) { s_need[e] = 0; }
        __syncthreads();
        for (int e = threadIdx.x; e < E; e += blockDim.x) { const unsigned c = s_cur[e]; if (c != kDead) s_need[c] = 1; }
        __syncthreads();
        for (int t = threadIdx.x; t < E; t += blockDim.x) {
            if (!s_need[t]) continue;
            unsigned long long pos = s_start + (unsigned)t;
            const unsigned long long s_end = g_start + (unsigned long long)c_sub_bounds[sub + 1];
            unsigned cnt = 0;
            unsigned res;
            while (true) {
                if (pos >= total) { res = (kDead << 16) | cnt; break; }
                if (pos >= s_end) { res = ((unsigned)(pos - s_end) << 16) | cnt; break; }
                const unsigned bits = block_bits_at(p.enc, total, pos, p.NN, p.use_rle);
                if (bits == kBadBlock) { res = (kDead << 16) | cnt; break; }
                pos += bits;
                cnt++;
            }
            s_memo[t] = res;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < E; e += blockDim.x) {
            const unsigned c = s_cur[e];
            if (c != kDead) { const unsigned m = s_memo[c]; s_cur[e] = (unsigned short)(m >> 16); s_cnt[e] += m & 0xFFFFu; }
        }
        __syncthreads();
    }
    for (int e = threadIdx.x; e < E; e += blockDim.x) p.group_table[(size_t)g * E + e] = make_uint2(s_cur[e], s_cnt[e]);
    }
}

__device__ __forceinline__ void parse_super_tables(const ParseParams &p) {
    for (unsigned sg = blockIdx.x; sg < p.nsuper; sg += gridDim.x) {
        const unsigned g0 = sg * kSuper, g1 = min(g0 + kSuper, p.ngroups);
        for (int e = threadIdx.x; e < p.E; e += blockDim.x) {
            unsigned cur = (unsigned)e, cnt = 0;
            for (unsigned g = g0; g < g1 && cur != kDead; g++) {
                const uint2 t = __ldcg(p.group_table + (size_t)g * p.E + cur);
                cur = t.x;
                cnt += t.y;
            }
            p.super_table[(size_t)sg * p.E + e] = make_uint2(cur, cnt);
        }
    }
}

__device__ __forceinline__ void parse_top_walk(const ParseParams &p) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned cur = 0, cnt = 0;
    for (unsigned sg = 0; sg < p.nsuper; sg++) {
        p.super_entry[sg] = make_uint2(cur, cnt);
        if (cur != kDead) { const uint2 t = __ldcg(p.super_table + (size_t)sg * p.E + cur); cur = t.x; cnt += t.y; }
    }
}

__device__ __forceinline__ void parse_down_super(const ParseParams &p) {
    for (unsigned sg = blockIdx.x * blockDim.x + threadIdx.x; sg < p.nsuper; sg += gridDim.x * blockDim.x) {
        const uint2 se = __ldcg(p.super_entry + sg);
        unsigned cur = se.x, cnt = se.y;
        const unsigned g0 = sg * kSuper, g1 = min(g0 + kSuper, p.ngroups);
        for (unsigned g = g0; g < g1; g++) {
            p.group_entry[g] = make_uint2(cur, cnt);
            if (cur != kDead) { const uint2 t = __ldcg(p.group_table + (size_t)g * p.E + cur); cur = t.x; cnt += t.y; }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Speculative fast path.  Real streams re-synchronise: a walk started at an arbitrary bit lands on the true chain after
// a few kilobits (measured: all 1044 entry offsets of a group merge into ONE chain within 3072 bits).  So every group
// starts a lead-in (4096 or 8192 bits) early at an arbitrary position and *assumes* it is on the true chain when it reaches its first bit.
// The assumption is then VERIFIED exactly: group g's speculative entry must equal group g-1's exit; group 0 starts at the
// true first block, so if all neighbours agree every group is on the true chain (induction).  Groups that disagree are
// re-walked from their predecessor's exit (a few rounds); if disagreement remains, spec_ok stays 0 and the exact
// transfer-function kernels above do the work instead (they return immediately when spec_ok is 1).
// ---------------------------------------------------------------------------------------------------------
// lead-in before a group's first bit (measured on 8x8 streams: 4096 bits leave 3 % of the groups unsynchronised, 8192 none)
constexpr int kSpecRounds = 3;            // parallel repair rounds inside a CTA (a serial pass settles what they leave)

// One step of a walk on the staged copy, `rel` < lim relative to the view's base.  Returns the bits consumed and the number
// of blocks they hold: an all-zero block is 4 zero bits (bit_len 0, no length/values), so a run of zero nibbles is a run of
// blocks and is taken in one step (up to 8, never starting a block at or past `lim`).  0 blocks = invalid length field.
__device__ __forceinline__ unsigned staged_step(const StagedStream &st, unsigned rel, unsigned lim, int NN, int rle, unsigned &blocks) {
    // branch-free: the lanes of a warp are at unrelated places of the stream, every data-dependent branch would serialise them
    const unsigned i = rel >> 5;
    const unsigned w0 = __byte_perm(st.w[i], 0, 0x0123);
    const unsigned w1 = __byte_perm(st.w[i + 1], 0, 0x0123);                // zero past the end of the stream (stage_stream)
    const unsigned v = __funnelshift_l(w1, w0, rel & 31u);
    const unsigned w = v >> 28;
    const unsigned zk = min(min((unsigned)__clz((int)v) >> 2, 8u), (lim - rel + 3u) >> 2);    // zero nibbles = all-zero blocks
    const unsigned lenf = __funnelshift_rc(v << 4, 0u, 32u - w);                               // (v << 4) >> (32 - w), 0 for w = 0
    const unsigned len = rle ? lenf : (unsigned)NN;
    const unsigned one = 4u + (rle ? w : 0u) + len * w;
    blocks = (w == 0) ? zk : ((len > (unsigned)NN) ? 0u : 1u);
    return (w == 0) ? 4u * zk : one;
}

// walks the chain from `rel` to the group's end (both relative to the view): (exit offset | kDead, blocks started)
__device__ __forceinline__ uint2 walk_group_staged(const StagedStream &st, int NN, int rle, unsigned rel, unsigned g_end_rel) {
    unsigned cnt = 0;
    const unsigned lim = min(st.total_rel, g_end_rel);
    while (rel < lim) {
        unsigned nb;
        const unsigned bits = staged_step(st, rel, lim, NN, rle, nb);
        if (nb == 0) return make_uint2(kDead, cnt);
        rel += bits;
        cnt += nb;
    }
    if (rel >= st.total_rel) return make_uint2(kDead, cnt);
    return make_uint2(rel - g_end_rel, cnt);
}

// the same on global memory (seam repairs)
__device__ __forceinline__ uint2 walk_group(const ParseParams &p, unsigned long long total, unsigned long long pos, unsigned long long g_end) {
    unsigned cnt = 0;
    while (true) {
        if (pos >= total) return make_uint2(kDead, cnt);
        if (pos >= g_end) return make_uint2((unsigned)(pos - g_end), cnt);
        const unsigned bits = block_bits_at(p.enc, total, pos, p.NN, p.use_rle);
        if (bits == kBadBlock) return make_uint2(kDead, cnt);
        pos += bits;
        cnt++;
    }
}

// Speculative grid: GB bits per group (one thread), TH groups per CTA, lead-in LEAD bits.  8x8 streams: 8192-bit groups,
// 64 per CTA, lead-in 8192 (4096 leaves 3 % of the groups unsynchronised).  4x4 streams (blocks <= 276 bits, ~50 typical):
// 2048-bit groups, 256 per CTA, lead-in 4096 -- a thread's serial chain is what the kernels' time is made of.
template <int GB, int TH> struct SpecCfg {
    static constexpr unsigned kLead = (GB == 8192) ? 8192u : (GB == 2048 ? 4096u : (GB == 1024 ? 2048u : 1024u));
    static constexpr unsigned kStageWords = (TH * GB + kLead) / 32 + 16;             // groups + lead-in + window/alignment slack
};

// One CTA = TH consecutive groups, one thread each, on a staged copy of their bits.  After the walks the CTA checks
// its own neighbours and re-walks the groups whose entry disagrees with their predecessor's exit (kSpecRounds rounds);
// the first group of every CTA is checked against the previous CTA by parse_spec_boundary.
template <int GB, int TH>
__global__ void __launch_bounds__(TH) parse_spec_walk(const ParseParams p) {
    pdl_wait();
    extern __shared__ __align__(16) unsigned s_stage[];
    __shared__ unsigned s_entry[TH], s_exit[TH];
    const unsigned cta = blockIdx.x + p.walk_cta0;
    const unsigned g = cta * TH + threadIdx.x;
    if (blockIdx.x == 0 && threadIdx.x == 0) { p.spec_flags[0] = 0; p.spec_flags[1] = 0; p.spec_flags[2] = 0; p.spec_flags[3] = 0xFFFFFFFFu; *p.cursor_next = ~0ull; }
    const unsigned long long total = *p.enc_bits;
    const unsigned long long B0 = *p.start + p.skip_bits;
    constexpr unsigned lead = SpecCfg<GB, TH>::kLead;
    const unsigned long long c_start = B0 + (unsigned long long)cta * TH * GB;
    const unsigned long long c_first = (c_start < B0 + lead) ? B0 : c_start - lead;
    const unsigned long long c_end = min(total, c_start + (unsigned long long)TH * GB);
    if (c_first >= total) {                                 // uniform: nothing of the stream in this CTA's range
        if (g < p.nspec) { p.spec_entry[g] = make_uint2(kDead, 0u); p.spec_exit[g] = make_uint2(kDead, 0u); }
        return;
    }
    const StagedStream st = stage_stream(s_stage, SpecCfg<GB, TH>::kStageWords, p.enc, total, c_first, c_end);
    const unsigned long long g_start = B0 + (unsigned long long)g * GB;
    const unsigned g_rel = (unsigned)(g_start - st.base), g_end_rel = g_rel + GB;     // CTA-local: fits 32 bits
    unsigned entry;
    uint2 ex = make_uint2(kDead, 0u);
    if (g >= p.nspec || g_start >= total) { entry = kDead; }
    else {
        unsigned rel = (g == 0 || g_start < B0 + lead) ? (unsigned)(B0 - st.base) : g_rel - lead;
        const unsigned lim = min(g_rel, st.total_rel);
        while (rel < lim) {                                // lead-in on an arbitrary phase; invalid headers slide by a bit
            unsigned nb;
            const unsigned bits = staged_step(st, rel, lim, p.NN, p.use_rle, nb);
            rel += nb ? bits : 1u;
        }
        entry = (rel >= st.total_rel) ? kDead : rel - g_rel;
        if (entry != kDead) ex = walk_group_staged(st, p.NN, p.use_rle, rel, g_end_rel);
    }
    s_entry[threadIdx.x] = entry;
    s_exit[threadIdx.x] = ex.x;
    __syncthreads();
    for (int r = 0; r < kSpecRounds; r++) {
        // only trust a predecessor that is itself consistent with ITS predecessor: a group whose lead-in failed to
        // synchronise walks garbage; adopting its exit would push the error forward round after round
        const int t = (int)threadIdx.x;
        bool redo = false;
        unsigned want = 0;
        // (the CTA's first group has no predecessor here, so nothing vouches for it: its successor is left to
        // parse_spec_boundary -- adopting the exit of an unsynchronised first group used to cascade through the CTA, one group
        // per round, and leave the boundary kernel a long serial re-walk)
        if (t >= 2 && g < p.nspec && s_exit[t - 2] == s_entry[t - 1]) {
            want = s_exit[t - 1];
            redo = (want != s_entry[t]);
        }
        if (!__syncthreads_or(redo)) break;
        if (redo) {
            entry = want;
            ex = (want == kDead || want >= (unsigned)GB) ? make_uint2(kDead, 0u)
                                                               : walk_group_staged(st, p.NN, p.use_rle, g_rel + want, g_end_rel);
        }
        __syncthreads();
        if (redo) { s_entry[t] = entry; s_exit[t] = ex.x; }
        __syncthreads();
    }
    // What the rounds left (runs of unsynchronised groups, groups behind one that nothing vouches for): one thread goes through
    // the CTA's seams in order and re-walks every group whose entry disagrees with its predecessor's exit.  If the CTA's first
    // group is itself wrong this follows its false chain -- until that chain meets the true one, as chains do after a block
    // or two; parse_spec_boundary then repairs exactly those groups.  A dead predecessor (a false chain that ran into an invalid
    // length field) is never followed: the group keeps what its own lead-in found.
    __shared__ unsigned s_cnt[TH];
    __shared__ unsigned char s_fix[TH];
    s_fix[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int t = 1; t < TH; t++) {
            const unsigned gt = cta * TH + (unsigned)t;
            const unsigned long long gt_start = B0 + (unsigned long long)gt * GB;
            if (gt >= p.nspec || gt_start >= total) break;
            const unsigned want = s_exit[t - 1];
            if (want == s_entry[t] || want == kDead || want >= (unsigned)GB) continue;
            const unsigned gr = (unsigned)(gt_start - st.base);
            const uint2 e2 = walk_group_staged(st, p.NN, p.use_rle, gr + want, gr + GB);
            s_entry[t] = want; s_exit[t] = e2.x; s_cnt[t] = e2.y; s_fix[t] = 1;
        }
    }
    __syncthreads();
    if (s_fix[threadIdx.x]) { entry = s_entry[threadIdx.x]; ex = make_uint2(s_exit[threadIdx.x], s_cnt[threadIdx.x]); }
    if (g < p.nspec) { p.spec_entry[g] = make_uint2(entry, 0u); p.spec_exit[g] = ex; }
}

// CTA seams: thread k checks the first group of CTA k against the last group of CTA k-1 and, on a mismatch, re-walks
// groups (on global memory: rare) until the chain agrees again or its CTA ends.  What it cannot settle is caught by
// parse_spec_finish (-> exact kernels).
template <int GB, int TH>
__global__ void __launch_bounds__(TH) parse_spec_boundary(const ParseParams p) {
    pdl_wait();
    const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned g = k * TH;
    if (k == 0 || g >= p.nspec) return;
    const unsigned long long total = *p.enc_bits;
    const unsigned long long B0 = *p.start + p.skip_bits;
    unsigned want = p.spec_exit[g - 1].x;
    for (unsigned n = 0; n < (unsigned)TH && g < p.nspec; n++, g++) {
        if (p.spec_entry[g].x == want) {
            // the CTA's first seam agrees: its second one is this kernel's too (parse_spec_walk does not touch it); from the
            // third group on an agreeing seam ends the repair
            if (n >= 1) break;
            want = p.spec_exit[g].x;
            continue;
        }
        const unsigned long long g_start = B0 + (unsigned long long)g * GB, g_end = g_start + GB;
        const uint2 ex = (want == kDead || want >= (unsigned)GB) ? make_uint2(kDead, 0u)
                                                                       : walk_group(p, total, g_start + want, g_end);
        p.spec_entry[g] = make_uint2(want, 0u);
        p.spec_exit[g] = ex;
        want = ex.x;
    }
}

// Final exact verification + block counts.  One thread per group: the seam in front of it must agree (first disagreeing
// group -> spec_flags[3]); every CTA adds up the blocks of its TH groups.  The last CTA to finish (ticket) turns
// the per-CTA counts into their exclusive scan and decides spec_ok.  Only the groups that hold the stream's nblocks blocks
// have to be consistent: behind the last block the chain runs into whatever follows (pad bits, the next frame's motion
// vectors), where the speculative walks may legitimately disagree.
template <int GB, int TH>
__global__ void __launch_bounds__(TH) parse_spec_check(const ParseParams p) {
    pdl_wait();
    __shared__ unsigned s_part[TH];
    __shared__ unsigned s_last;
    const unsigned g = blockIdx.x * TH + threadIdx.x;
    unsigned cnt = 0;
    if (g < p.nspec) {
        cnt = p.spec_exit[g].y;
        if (g > 0 && p.spec_exit[g - 1].x != p.spec_entry[g].x) atomicMin(&p.spec_flags[3], g);
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned sum = 0;
        for (int w = 0; w < TH / 32; w++) sum += s_part[w];
        p.walk_base[blockIdx.x] = sum;
        __threadfence();
        s_last = (atomicAdd(&p.spec_flags[0], 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // exclusive scan of walk_base[0 .. gridDim.x), TH values at a time (read past L1: other CTAs wrote them)
    unsigned run = 0;
    for (unsigned c0 = 0; c0 < gridDim.x; c0 += TH) {
        const unsigned i = c0 + threadIdx.x;
        const unsigned v = (i < gridDim.x) ? __ldcg(p.walk_base + i) : 0u;
        unsigned inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned o = __shfl_up_sync(0xffffffffu, inc, d); if ((int)(threadIdx.x & 31) >= d) inc += o; }
        __syncthreads();
        if ((threadIdx.x & 31) == 31) s_part[threadIdx.x >> 5] = inc;
        __syncthreads();
        unsigned before = run;
        for (unsigned w = 0; w < (threadIdx.x >> 5); w++) before += s_part[w];
        if (i < gridDim.x) p.walk_base[i] = before + inc - v;
        for (int w = 0; w < TH / 32; w++) run += s_part[w];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned firstbad = __ldcg(&p.spec_flags[3]);
        unsigned ok = 1;
        if (firstbad < p.nspec) {
            // blocks that start before the first unverified group; if that does not cover the stream's blocks, give up
            __threadfence();
            unsigned base = p.walk_base[firstbad / TH];
            for (unsigned gg = firstbad / TH * TH; gg < firstbad; gg++) base += p.spec_exit[gg].y;
            if (base < p.nblocks) ok = 0;
        }
        p.spec_flags[1] = p.force_exact ? 0u : ok;
    }
}

// block_off[] from the verified speculative walk.  One CTA = TH groups on a staged copy: every thread walks its group's TRUE
// chain from its entry; the first block index of a group = scanned per-CTA counts + the counts of the CTA's earlier groups.
template <int GB, int TH>
__global__ void __launch_bounds__(TH) parse_spec_emit(const ParseParams p) {
    pdl_wait();
    extern __shared__ __align__(16) unsigned s_stage[];
    __shared__ unsigned s_wsum[TH / 32];
    if (!p.spec_flags[1]) return;                            // uniform: the exact kernels produce the offsets
    const unsigned g = blockIdx.x * TH + threadIdx.x;
    const unsigned long long total = *p.enc_bits;
    const unsigned long long B0 = *p.start + p.skip_bits;
    if (B0 >= total) {
        // nothing of the stream is left for this span (a truncated file): every block starts and ends at `total` and reads
        // as zero bits (BitStream.cpp:17-20)
        if (blockIdx.x == 0) {
            for (unsigned i = threadIdx.x; i <= p.nblocks; i += TH) p.block_off[i] = total;
            if (threadIdx.x == 0) *p.cursor_next = total;
        }
        return;
    }
    const unsigned long long c_start = B0 + (unsigned long long)blockIdx.x * TH * GB;
    if (c_start >= total) return;                            // uniform
    if (p.emit_hi) {
        // sharded decode: a CTA none of whose blocks this rank decodes has nothing to stage (walk_base = the exclusive scan
        // of the per-CTA block counts, final since parse_spec_check)
        const unsigned lo_c = p.walk_base[blockIdx.x];
        const unsigned hi_c = (blockIdx.x + 1 < gridDim.x) ? p.walk_base[blockIdx.x + 1] : 0xFFFFFFFFu;
        if (lo_c > p.emit_hi || hi_c < p.emit_lo) return;   // uniform
    }
    const StagedStream st = stage_stream(s_stage, SpecCfg<GB, TH>::kStageWords, p.enc, total, c_start,
                                         min(total, c_start + (unsigned long long)TH * GB));
    const unsigned cnt = (g < p.nspec) ? p.spec_exit[g].y : 0u;
    unsigned inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned o = __shfl_up_sync(0xffffffffu, inc, d); if ((int)(threadIdx.x & 31) >= d) inc += o; }
    if ((threadIdx.x & 31) == 31) s_wsum[threadIdx.x >> 5] = inc;
    __syncthreads();
    unsigned idx = p.walk_base[blockIdx.x] + inc - cnt;
    for (unsigned w = 0; w < (threadIdx.x >> 5); w++) idx += s_wsum[w];
    if (g >= p.nspec || g >= p.spec_flags[3]) return;        // unverified tail: holds no block of this stream
    const unsigned entry = p.spec_entry[g].x;
    if (entry == kDead) return;
    if (p.emit_hi && (idx > p.emit_hi || idx + cnt < p.emit_lo)) return;      // sharded decode: none of this rank's blocks
    const unsigned first_idx = idx;
    const unsigned long long g_start = B0 + (unsigned long long)g * GB;
    const unsigned g_end_rel = (unsigned)(g_start - st.base) + GB;
    unsigned rel = (unsigned)(g_start - st.base) + entry;
    const unsigned lim = min(g_end_rel, st.total_rel);
    while (rel < lim && idx < p.nblocks) {
        unsigned nb;
        const unsigned bits = staged_step(st, rel, lim, p.NN, p.use_rle, nb);
        if (nb == 0) { if (p.err) atomicExch(p.err, IE_EFORMAT); rel = st.total_rel; break; }             // malformed stream
        for (unsigned j = 0; j < nb && idx < p.nblocks; j++) p.block_off[idx++] = st.base + rel + 4u * j;  // nb > 1: zero blocks
        rel = min(rel + bits, st.total_rel);
    }
    const unsigned long long pos = st.base + rel;
    if (idx <= p.nblocks && pos >= total) {
        // the chain reached the end of the stream in this group: every remaining block starts (and ends) at `total`
        for (; idx < p.nblocks; idx++) p.block_off[idx] = total;
        p.block_off[p.nblocks] = total;
        *p.cursor_next = total;
    } else if (idx == p.nblocks && first_idx < p.nblocks) {
        p.block_off[idx] = pos;                 // this thread emitted the last block
        *p.cursor_next = pos;
    }
}

// block_off[] from the exact path's group entries (only when the speculative parse did not verify)
__device__ __forceinline__ void parse_emit_offsets(const ParseParams &p) {
    const unsigned long long total = *p.enc_bits;
    const unsigned long long B0 = *p.start + p.skip_bits;
    for (unsigned g = blockIdx.x * blockDim.x + threadIdx.x; g < p.ngroups; g += gridDim.x * blockDim.x) {
        const uint2 ge = __ldcg(p.group_entry + g);
        if (ge.x == kDead) continue;
        const unsigned long long g_start = B0 + (unsigned long long)g * kGroupBits, g_end = g_start + kGroupBits;
        unsigned long long pos = g_start + ge.x;
        unsigned idx = ge.y;
        while (pos < g_end && pos < total && idx < p.nblocks) {
            const unsigned bits = block_bits_at(p.enc, total, pos, p.NN, p.use_rle);
            if (bits == kBadBlock) { if (p.err) atomicExch(p.err, IE_EFORMAT); pos = total; break; }   // malformed stream
            p.block_off[idx++] = pos;
            pos = min(pos + bits, total);
        }
        if (idx <= p.nblocks && pos >= total) {
            for (; idx < p.nblocks; idx++) p.block_off[idx] = total;
            p.block_off[p.nblocks] = total;
            *p.cursor_next = total;
        } else if (idx == p.nblocks && ge.y < p.nblocks) {
            p.block_off[idx] = pos;                 // this thread emitted the last block
            *p.cursor_next = pos;
        }
    }
}

// grid barrier of parse_exact_kernel: spec_flags[2] counts arrivals (zeroed by parse_spec_walk at the start of every parse)
__device__ __forceinline__ void parse_grid_barrier(const ParseParams &p, unsigned phase) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(&p.spec_flags[2], 1u);
        while (ld_relaxed_u32(&p.spec_flags[2]) < phase * gridDim.x) __nanosleep(100);
        __threadfence();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(64) parse_exact_kernel(const ParseParams p) {
    pdl_wait();
    if (p.spec_flags[1]) return;                            // the speculative parse verified (every CTA sees the same flag)
    parse_group_tables(p);
    parse_grid_barrier(p, 1);
    parse_super_tables(p);
    parse_grid_barrier(p, 2);
    parse_top_walk(p);
    parse_grid_barrier(p, 3);
    parse_down_super(p);
    parse_grid_barrier(p, 4);
    parse_emit_offsets(p);
}

__global__ void parse_commit_cursor(const ParseParams p) {
    pdl_wait();
    const unsigned long long next = *p.cursor_next;
    if (next != ~0ull) *p.cursor_out = next;
}

static void parse_sizes(size_t span_bits, int N, unsigned &E, unsigned &ngroups, unsigned &nsuper) {
    E = 4 + 16 + 16 * N * N;
    ngroups = (unsigned)((span_bits + kGroupBits - 1) / kGroupBits + 1);
    nsuper = (ngroups + kSuper - 1) / kSuper;
}

// 4x4 streams: images walk 2048-bit groups; video frames (cursor mode: one short parse per frame, pure latency) finer ones
static int g4_bits(bool video) { const int v = g_parse_grid4.load(); return v == 2 ? 512 : v == 1 ? 1024 : v == 0 ? 2048 : (video ? 1024 : 2048); }
static unsigned spec_groups(size_t span_bits, int N, bool video = true) { const size_t gb = (N == 8) ? 8192 : (size_t)g4_bits(video); return (unsigned)((span_bits + gb - 1) / gb + 1); }

size_t parse_scratch_bytes(size_t enc_bytes, int N) {
    unsigned E, ng, ns;
    parse_sizes(enc_bytes * 8, N, E, ng, ns);
    const size_t nspec = (N == 8) ? spec_groups(enc_bytes * 8, N) : (enc_bytes * 8 + 511) / 512 + 1;      // the finest grid any setting uses
    // (+ slack for the padded grid of a sharded decode: up to 64 parts x one walk CTA)
    return ((size_t)ng * E + (size_t)ns * E + ns + ng + 8) * sizeof(uint2) + nspec * 2 * sizeof(uint2) + (nspec / 64 + 16 + 64 * 256) * sizeof(unsigned) + 256;
}

template <int GB, int TH>
static int launch_spec(const ParseParams &p, cudaStream_t stream) {
    const unsigned nwalk = (p.nspec + TH - 1) / TH;
    const size_t stage_bytes = (size_t)SpecCfg<GB, TH>::kStageWords * sizeof(unsigned);
    IE_CUDA(cudaFuncSetAttribute(parse_spec_walk<GB, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes));
    IE_CUDA(cudaFuncSetAttribute(parse_spec_emit<GB, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes));
    IE_CUDA(launch_pdl(parse_spec_walk<GB, TH>, dim3(nwalk), dim3(TH), stage_bytes, stream, p));
    IE_CUDA(launch_pdl(parse_spec_boundary<GB, TH>, dim3((nwalk + 63) / 64), dim3(64), 0, stream, p));
    IE_CUDA(launch_pdl(parse_spec_check<GB, TH>, dim3(nwalk), dim3(TH), 0, stream, p));
    IE_CUDA(launch_pdl(parse_spec_emit<GB, TH>, dim3(nwalk), dim3(TH), stage_bytes, stream, p));
    IE_CUDA(cudaGetLastError());
    return IE_OK;
}

// Fills d.block_off[0..nblocks] for one stream (and advances d.cursor when the last block lies inside the span).
// `span_bits`: how far past the first block the chain is followed (the whole stream for images; a per-frame budget for
// video, where the caller retries with a larger span if the cursor did not move).  `scratch`: parse_scratch_bytes(bytes
// of the whole stream) bytes, 16-aligned.
static void parse_setup(const DecodeParams &d, size_t span_bits, uint8_t *scratch, ParseParams &p, unsigned &E) {
    memset(&p, 0, sizeof p);
    parse_sizes(span_bits, d.N, E, p.ngroups, p.nsuper);
    p.enc = d.enc; p.enc_bits = d.enc_bits;
    p.start = d.cursor ? d.cursor : d.start_bit;
    p.skip_bits = d.cursor ? d.skip_bits : 0;
    p.NN = d.N * d.N; p.use_rle = d.use_rle; p.E = (int)E;
    p.nblocks = d.nblocks;
    uint2 *s = reinterpret_cast<uint2 *>(scratch);
    p.group_table = s; s += (size_t)p.ngroups * E;
    p.super_table = s; s += (size_t)p.nsuper * E;
    p.super_entry = s; s += p.nsuper;
    p.group_entry = s; s += p.ngroups;
    p.nspec = spec_groups(span_bits, d.N, d.cursor != nullptr);
    p.spec_entry = s; s += p.nspec;
    p.spec_exit = s; s += p.nspec;
    p.spec_flags = reinterpret_cast<unsigned *>(s);
    p.cursor_next = reinterpret_cast<unsigned long long *>(p.spec_flags + 8);
    p.walk_base = p.spec_flags + 12;
    p.block_off = d.block_off;
    p.cursor_out = d.cursor;
    p.err = d.err;
    p.force_exact = g_parse_variant.load() == 1;
}

int launch_parallel_parse(const DecodeParams &d, size_t span_bits, uint8_t *scratch, cudaStream_t stream) {
    ParseParams p;
    unsigned E;
    parse_setup(d, span_bits, scratch, p, E);
    const size_t smem = (size_t)E * (4 + 4 + 2 + 1) + 16;
    static const bool dbg = getenv("IE_DEBUG_SYNC") != nullptr;
#define IE_DBG_STEP(name) do { if (dbg) { cudaError_t e_ = cudaStreamSynchronize(stream); if (e_ != cudaSuccess) { fprintf(stderr, "[ie] %s failed: %s\n", name, cudaGetErrorString(e_)); return cuda_fail(e_, name, __FILE__, __LINE__); } } } while (0)
    IE_DBG_STEP("before parse");
    if ((uintptr_t)d.enc % 16) { set_error("encoded stream must be 16-byte aligned on the device"); return IE_EINVAL; }
    if (d.N == 8) IE_TRY((launch_spec<8192, 64>(p, stream)));
    else if (g4_bits(d.cursor != nullptr) == 512) IE_TRY((launch_spec<512, 256>(p, stream)));
    else if (g4_bits(d.cursor != nullptr) == 1024) IE_TRY((launch_spec<1024, 256>(p, stream)));
    else IE_TRY((launch_spec<2048, 256>(p, stream)));
    IE_DBG_STEP("parse_spec");
    {
        // one launch for the whole exact path; every CTA must be resident for its grid barriers: 2 CTAs of 64 threads per SM
        static int sms = 0;
        if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
        static bool configured = false;
        if (!configured) { IE_CUDA(cudaFuncSetAttribute(parse_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1024)); configured = true; }
        IE_CUDA(launch_pdl(parse_exact_kernel, dim3(2u * (unsigned)sms), dim3(64), smem, stream, p));
        IE_DBG_STEP("parse_exact_kernel");
    }
    if (p.cursor_out) { IE_CUDA(launch_pdl(parse_commit_cursor, dim3(1), dim3(1), 0, stream, p)); count_launch(); }
#undef IE_DBG_STEP
    count_launch(5);
    IE_CUDA(cudaGetLastError());
    return IE_OK;
}


// ---------------------------------------------------------------------------------------------------------
// Sharded decode of ONE image stream (ImageDecoder.cpp:88-112): every rank holds the stream; rank `part` walks only
// its share of the speculative grid (the walk is the expensive step of the parse), the ranks all-gather the per-group results
// (16 bytes per group: the caller's collective, in place on d_spec), and every rank then verifies the seams, counts, and
// emits the offsets of its own block rows only.  d_spec = [entry: parts x chunk][exit: parts x chunk].
// ---------------------------------------------------------------------------------------------------------
ShardedParseGeom sharded_parse_geom(size_t enc_bytes, int N, unsigned parts) {
    ShardedParseGeom g;
    const size_t gb = (N == 8) ? 8192 : 2048, th = (N == 8) ? 64 : 256;
    const size_t nspec = (enc_bytes * 8 + gb - 1) / gb + 1;
    const size_t nwalk = (nspec + th - 1) / th;
    g.ctas_per_part = (unsigned)((nwalk + parts - 1) / parts);
    g.groups_per_part = g.ctas_per_part * (unsigned)th;
    g.chunk_bytes = (size_t)g.groups_per_part * sizeof(uint2);
    g.spec_bytes = 2 * (size_t)parts * g.chunk_bytes;
    return g;
}

__global__ void fill_off_kernel(unsigned long long *off, unsigned n, const unsigned long long *value) {
    pdl_wait();
    const unsigned long long v = *value;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) off[i] = v;
}

template <int GB, int TH>
static int launch_walk_part(ParseParams &p, const ShardedParseGeom &g, unsigned part, cudaStream_t stream) {
    const size_t stage_bytes = (size_t)SpecCfg<GB, TH>::kStageWords * sizeof(unsigned);
    IE_CUDA(cudaFuncSetAttribute(parse_spec_walk<GB, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes));
    p.walk_cta0 = part * g.ctas_per_part;
    IE_CUDA(launch_pdl(parse_spec_walk<GB, TH>, dim3(g.ctas_per_part), dim3(TH), stage_bytes, stream, p));
    count_launch();
    IE_CUDA(cudaGetLastError());
    return IE_OK;
}

template <int GB, int TH>
static int launch_finish(const ParseParams &p, cudaStream_t stream) {
    const unsigned nwalk = (p.nspec + TH - 1) / TH;
    const size_t stage_bytes = (size_t)SpecCfg<GB, TH>::kStageWords * sizeof(unsigned);
    IE_CUDA(cudaFuncSetAttribute(parse_spec_emit<GB, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes));
    IE_CUDA(launch_pdl(parse_spec_boundary<GB, TH>, dim3((nwalk + 63) / 64), dim3(64), 0, stream, p));
    IE_CUDA(launch_pdl(parse_spec_check<GB, TH>, dim3(nwalk), dim3(TH), 0, stream, p));
    IE_CUDA(launch_pdl(parse_spec_emit<GB, TH>, dim3(nwalk), dim3(TH), stage_bytes, stream, p));
    count_launch(3);
    IE_CUDA(cudaGetLastError());
    return IE_OK;
}

static void sharded_setup(const DecodeParams &d, size_t span_bits, uint8_t *scratch, uint8_t *d_spec, unsigned parts, ParseParams &p, unsigned &E,
                          ShardedParseGeom &g) {
    parse_setup(d, span_bits, scratch, p, E);
    g = sharded_parse_geom(span_bits / 8, d.N, parts);
    p.nspec = parts * g.groups_per_part;
    p.spec_entry = reinterpret_cast<uint2 *>(d_spec);
    p.spec_exit = reinterpret_cast<uint2 *>(d_spec + (size_t)parts * g.chunk_bytes);
}

int launch_parse_walk_part(const DecodeParams &d, size_t span_bits, uint8_t *scratch, uint8_t *d_spec, unsigned part, unsigned parts,
                           cudaStream_t stream) {
    ParseParams p;
    unsigned E;
    ShardedParseGeom g;
    sharded_setup(d, span_bits, scratch, d_spec, parts, p, E, g);
    if ((uintptr_t)d.enc % 16 || (uintptr_t)d_spec % 16) { set_error("stream and spec buffers must be 16-byte aligned on the device"); return IE_EINVAL; }
    if (d.N == 8) return launch_walk_part<8192, 64>(p, g, part, stream);
    return launch_walk_part<2048, 256>(p, g, part, stream);
}

// after the all-gather of d_spec: seams, counts, block_off[lo .. hi] (hi inclusive), exact path if the speculation failed
int launch_parse_finish_range(const DecodeParams &d, size_t span_bits, uint8_t *scratch, uint8_t *d_spec, unsigned parts, unsigned lo, unsigned hi,
                              cudaStream_t stream) {
    ParseParams p;
    unsigned E;
    ShardedParseGeom g;
    sharded_setup(d, span_bits, scratch, d_spec, parts, p, E, g);
    p.emit_lo = lo; p.emit_hi = hi;
    // blocks the chain never reaches (a truncated stream) start and end at the end of the stream (BitStream.cpp:17-20)
    IE_CUDA(launch_pdl(fill_off_kernel, dim3(64), dim3(256), 0, stream, p.block_off + lo, hi - lo + 1, p.enc_bits));
    count_launch();
    if (d.N == 8) IE_TRY((launch_finish<8192, 64>(p, stream)));
    else IE_TRY((launch_finish<2048, 256>(p, stream)));
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    IE_CUDA(cudaFuncSetAttribute(parse_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1024));
    const size_t smem = (size_t)E * (4 + 4 + 2 + 1) + 16;
    IE_CUDA(launch_pdl(parse_exact_kernel, dim3(2u * (unsigned)sms), dim3(64), smem, stream, p));
    count_launch();
    IE_CUDA(cudaGetLastError());
    return IE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Whole-stream parse of a video (Frame.cpp:47-127, VideoDecoder.cpp:33-62).  The reference reads the frames strictly in
// sequence: frame f+1 starts where frame f's last block ends, and a P-frame's blocks start a fixed number of bits (the
// motion-vector section, Block.cpp:415-423) after that.  The per-frame path above follows that chain with one parse (six short
// launches) per frame.  Here:
//   1. parse_spec_walk over the WHOLE stream (one launch): every 1024-bit group gets a speculative entry, exit and block
//      count.  Inside a frame's block section the speculation is right as soon as the walkers have synchronised; across a
//      motion-vector section it is garbage -- which is harmless, because nothing below trusts a group before an exact walk has
//      arrived at precisely its entry.
//   2. vparse_scan_kernel: prefix sums over the groups of (blocks started, bad seams), one launch.
//   3. vparse_chain_kernel (one CTA, the only sequential part: a few microseconds per frame): from a frame's true first
//      bit it walks until its position at a group boundary equals that group's speculative entry (from there on the
//      speculative chain IS the true chain, the walk being deterministic), jumps over all groups whose seams are verified
//      (prefix sums: a cooperative search for the group that holds the frame's last block, a check that no bad seam lies in
//      between), walks inside that last group to the exact end of the frame.  What it finds is a list of pieces per frame:
//      HEAD pieces (block starts found by its own walk, kept in a pool) and SPEC pieces (runs of verified groups).
//   4. vparse_emit_kernel: block_off[] of frame k of every GOP at once from the pieces (parallel over groups).
// Anything unusual (stream ends inside a frame, invalid length field on the true chain, no re-synchronisation within the
// limits) clears the ok flag and the caller takes the per-frame path, which implements the reference's behaviour for
// truncated and damaged streams.
// ---------------------------------------------------------------------------------------------------------
constexpr int kVG = 1024;              // bits per group of the whole-stream grid
constexpr int kVTH = 256;              // groups per walk / emit CTA
constexpr unsigned kVHeadGroups = 16;  // groups staged per window of a head walk
constexpr unsigned kVMaxWindows = 64;  // head walk gives up after this many windows without meeting the speculative chain
constexpr unsigned kVChainThreads = 512;

__device__ __forceinline__ bool vseam_bad(const uint2 *spec_entry, const uint2 *spec_exit, unsigned g) {
    if (g == 0) return false;
    const unsigned e = spec_entry[g].x;
    return e == kDead || spec_exit[g - 1].x != e;
}

// pq[g] = (blocks started in groups < g, bad seams with index <= g), g = 0 .. nspec (pq[nspec].y repeats the last).
// One launch: per-CTA totals, grid barrier, every CTA adds up the totals in front of it and rescans its own span.
__global__ void __launch_bounds__(256) vparse_scan_kernel(const ParseParams p, uint2 *pq, uint2 *partial) {
    pdl_wait();
    __shared__ unsigned s_a[8], s_b[8];
    __shared__ unsigned s_base_a, s_base_b;
    const unsigned n = p.nspec;
    const unsigned per = ((n + gridDim.x - 1) / gridDim.x + 255u) / 256u * 256u;
    const unsigned g0 = blockIdx.x * per, g1 = min(n, g0 + per);
    unsigned a = 0, b = 0;
    for (unsigned g = g0 + threadIdx.x; g < g1; g += 256) { a += p.spec_exit[g].y; b += vseam_bad(p.spec_entry, p.spec_exit, g) ? 1u : 0u; }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, d); b += __shfl_xor_sync(0xffffffffu, b, d); }
    if ((threadIdx.x & 31) == 0) { s_a[threadIdx.x >> 5] = a; s_b[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned ta = 0, tb = 0;
        for (int w = 0; w < 8; w++) { ta += s_a[w]; tb += s_b[w]; }
        partial[blockIdx.x] = make_uint2(ta, tb);
    }
    parse_grid_barrier(p, 1);
    a = 0; b = 0;
    for (unsigned c = threadIdx.x; c < blockIdx.x; c += 256) { const uint2 t = __ldcg(partial + c); a += t.x; b += t.y; }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, d); b += __shfl_xor_sync(0xffffffffu, b, d); }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { s_a[threadIdx.x >> 5] = a; s_b[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned ta = 0, tb = 0;
        for (int w = 0; w < 8; w++) { ta += s_a[w]; tb += s_b[w]; }
        s_base_a = ta; s_base_b = tb;
    }
    __syncthreads();
    unsigned run_a = s_base_a, run_b = s_base_b;
    for (unsigned c0 = g0; c0 < g1; c0 += 256) {
        const unsigned g = c0 + threadIdx.x;
        const unsigned va = (g < g1) ? p.spec_exit[g].y : 0u;
        const unsigned vb = (g < g1 && vseam_bad(p.spec_entry, p.spec_exit, g)) ? 1u : 0u;
        unsigned ia = va, ib = vb;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned oa = __shfl_up_sync(0xffffffffu, ia, d), ob = __shfl_up_sync(0xffffffffu, ib, d);
            if ((int)(threadIdx.x & 31) >= d) { ia += oa; ib += ob; }
        }
        __syncthreads();
        if ((threadIdx.x & 31) == 31) { s_a[threadIdx.x >> 5] = ia; s_b[threadIdx.x >> 5] = ib; }
        __syncthreads();
        unsigned ba = run_a, bb = run_b;
        for (unsigned w = 0; w < (threadIdx.x >> 5); w++) { ba += s_a[w]; bb += s_b[w]; }
        if (g < g1) pq[g] = make_uint2(ba + ia - va, bb + ib);         // exclusive block count, inclusive bad-seam count
        for (int w = 0; w < 8; w++) { run_a += s_a[w]; run_b += s_b[w]; }
        if (g + 1 == n) pq[n] = make_uint2(ba + ia, bb + ib);
    }
    if (n == 0 && blockIdx.x == 0 && threadIdx.x == 0) pq[0] = make_uint2(0u, 0u);
}

// largest g in [lo, hi] for which pred(g) holds; pred is monotone (true ... true false ... false) and pred(lo) is true.
// All threads of the CTA call it with the same arguments and get the same answer.
template <typename Pred>
__device__ __forceinline__ unsigned vsearch_last_true(unsigned lo, unsigned hi, Pred pred) {
    while (hi > lo) {
        const unsigned n = hi - lo;
        const unsigned step = (n + kVChainThreads - 1) / kVChainThreads;
        const unsigned long long idx = (unsigned long long)lo + (unsigned long long)(threadIdx.x + 1) * step;
        const int ok = (idx <= hi) && pred((unsigned)idx);
        const unsigned k = (unsigned)__syncthreads_count(ok);
        const unsigned nlo = lo + k * step;
        hi = min(hi, nlo + step - 1);
        lo = nlo;
    }
    return lo;
}

struct VChainShared {
    unsigned long long pos;
    unsigned cnt, pool_n, merged, m, fail;
    uint2 pqm;
};

// Frames [f0, f1) of the clip; the chain's state (next bit, pool fill, failure) travels from launch to launch in v.state, so that
// the reconstruction of a batch of GOPs (second stream) can start while the chain works on the next batch.  After a failure the
// remaining frames get empty records that point at the end of the stream (the reconstruction kernels of the batches already
// enqueued stay well-defined; the caller discards their output and decodes frame by frame).
__global__ void __launch_bounds__(kVChainThreads) vparse_chain_kernel(const ParseParams p, const VideoParse v, unsigned f0, unsigned f1) {
    pdl_wait();
    constexpr unsigned kMaxBlock = 4 + 16 + 16 * 16;
    constexpr unsigned kWinWords = (kVHeadGroups * kVG + kVG + kMaxBlock) / 32 + 32;
    __shared__ __align__(16) unsigned s_win[kWinWords];
    __shared__ unsigned s_ent[kVHeadGroups];
    __shared__ uint2 s_pq[kVHeadGroups];
    __shared__ VChainShared sh;
    const unsigned long long total = *p.enc_bits;
    const unsigned long long B0 = *p.start;
    const unsigned B = v.nblocks;
    const unsigned max_fg = (unsigned)(((unsigned long long)B * kMaxBlock) / kVG) + 4u;
    unsigned long long pos = (f0 == 0) ? B0 : v.state[0];
    unsigned pool_n = (f0 == 0) ? 0u : (unsigned)v.state[1];
    bool fail = (f0 == 0) ? ((B0 >= total) && v.frames != 0) : (v.state[2] != 0);
    __syncthreads();                                        // everyone has read the state before thread 0 rewrites it
    unsigned f = f0;
    for (; f < f1 && !fail; f++) {
        const bool is_p = (f % v.gop) != 0;
        if (is_p) pos += v.mv_bits;
        if (pos >= total) { fail = true; break; }
        VFrameRec *rec = v.rec + f;
        if (threadIdx.x == 0) rec->first = pos;
        unsigned cnt = 0, npieces = 0;
        unsigned long long E = pos;
        while (cnt < B && !fail) {
            // ---- HEAD piece: exact walk from `pos` until a group boundary is met at that group's speculative entry
            const unsigned pool_start = pool_n, head_base = cnt;
            bool merged = false;
            unsigned m = 0;
            for (unsigned win = 0; win < kVMaxWindows && !merged && cnt < B && !fail; win++) {
                const unsigned long long g = (pos - B0) / kVG;
                if (g + 1 >= p.nspec) { fail = true; break; }
                const unsigned long long w_end = B0 + (g + 1 + kVHeadGroups) * (unsigned long long)kVG;
                __syncthreads();
                const StagedStream st = stage_stream(s_win, kWinWords, p.enc, total, pos, min(total, w_end) + kMaxBlock);
                if (threadIdx.x < kVHeadGroups) {
                    const unsigned long long gg = g + 1 + threadIdx.x;
                    s_ent[threadIdx.x] = (gg < p.nspec) ? p.spec_entry[gg].x : kDead;
                } else if (threadIdx.x < 2 * kVHeadGroups) {                                   // the prefix sums of the same groups:
                    const unsigned long long gg = g + 1 + (threadIdx.x - kVHeadGroups);          // one of them is the merge group's
                    s_pq[threadIdx.x - kVHeadGroups] = (gg <= p.nspec) ? __ldcg(v.pq + gg) : make_uint2(0u, 0u);
                }
                __syncthreads();
                if (threadIdx.x == 0) {
                    unsigned rel = (unsigned)(pos - st.base);
                    unsigned c = cnt, pn = pool_n, mg = 0, mm = 0, fl = 0;
                    for (unsigned j = 0; j < kVHeadGroups && !fl; j++) {
                        const unsigned long long gend = B0 + (g + 1 + j) * (unsigned long long)kVG;
                        const unsigned gend_rel = (unsigned)(gend - st.base);
                        const unsigned lim = min(gend_rel, st.total_rel);
                        while (rel < lim && c < B) {
                            unsigned nb;
                            const unsigned bits = staged_step(st, rel, lim, p.NN, p.use_rle, nb);
                            if (nb == 0) { fl = 1; break; }                                  // invalid length field on the true chain
                            const unsigned take = min(nb, B - c);
                            if (pn + take > v.pool_cap) { fl = 1; break; }
                            for (unsigned q = 0; q < take; q++) v.pool[pn + q] = st.base + rel + 4u * q;
                            pn += take; c += take;
                            rel += (take < nb) ? 4u * take : bits;
                        }
                        if (fl || c >= B) break;
                        if (rel >= st.total_rel) { fl = 1; break; }                         // the stream ends inside the frame
                        if (rel - gend_rel == s_ent[j]) { mg = 1; mm = (unsigned)(g + 1 + j); break; }
                    }
                    sh.pos = st.base + rel; sh.cnt = c; sh.pool_n = pn; sh.merged = mg; sh.m = mm; sh.fail = fl;
                    if (mg) sh.pqm = s_pq[mm - (unsigned)(g + 1)];
                }
                __syncthreads();
                pos = sh.pos; cnt = sh.cnt; pool_n = sh.pool_n; merged = sh.merged != 0; m = sh.m; fail = sh.fail != 0;
            }
            if (fail) break;
            if (pool_n > pool_start) {
                if (npieces >= kVMaxPieces) { fail = true; break; }
                if (threadIdx.x == 0) rec->piece[npieces] = VFramePiece{0u, pool_start, pool_n - pool_start, head_base, 0u};
                npieces++;
            }
            if (cnt >= B) { E = pos; break; }
            if (!merged) { fail = true; break; }
            // ---- SPEC piece: groups m .. glast, all seams verified
            const unsigned R = B - cnt;
            const uint2 pqm = sh.pqm;                                                     // staged with the merge group's entry
            const unsigned hi = (unsigned)min((unsigned long long)p.nspec - 1ull, (unsigned long long)m + max_fg);
            const uint2 *pq = v.pq;
            unsigned gs = vsearch_last_true(m, hi, [&](unsigned g) { return (__ldcg(pq + g).x - pqm.x) < R; });
            const uint2 pqs = __ldcg(pq + gs), pqs1 = __ldcg(pq + gs + 1);               // gs + 1 <= nspec: pq has nspec + 1 entries
            const unsigned es = p.spec_entry[gs].x;
            if (npieces >= kVMaxPieces) { fail = true; break; }
            if (pqs.y != pqm.y) {
                // a bad seam before the frame's last group: take the verified groups, go on with an exact walk from there
                const unsigned glast = vsearch_last_true(m, gs, [&](unsigned g) { return __ldcg(pq + g).y == pqm.y; });
                const unsigned xe = p.spec_exit[glast].x;
                if (xe == kDead) { fail = true; break; }
                if (threadIdx.x == 0) rec->piece[npieces] = VFramePiece{1u, m, glast, cnt, pqm.x};
                npieces++;
                cnt += __ldcg(pq + glast + 1).x - pqm.x;
                pos = B0 + (unsigned long long)(glast + 1) * kVG + xe;
                continue;
            }
            const unsigned r = R - (pqs.x - pqm.x);                                        // blocks to take in group gs (>= 1)
            if (r > pqs1.x - pqs.x) { fail = true; break; }                                // the stream ends inside the frame
            if (es == kDead) { fail = true; break; }
            if (threadIdx.x == 0) rec->piece[npieces] = VFramePiece{1u, m, gs, cnt, pqm.x};
            npieces++;
            const unsigned long long gpos = B0 + (unsigned long long)gs * kVG + es;
            __syncthreads();
            const StagedStream st = stage_stream(s_win, kWinWords, p.enc, total, gpos, min(total, gpos + kVG + kMaxBlock) + kMaxBlock);
            if (threadIdx.x == 0) {
                unsigned rel = (unsigned)(gpos - st.base), left = r, fl = 0;
                while (left) {
                    if (rel >= st.total_rel) { fl = 1; break; }
                    unsigned nb;
                    const unsigned bits = staged_step(st, rel, st.total_rel, p.NN, p.use_rle, nb);
                    if (nb == 0) { fl = 1; break; }
                    const unsigned take = min(nb, left);
                    rel += (take < nb) ? 4u * take : bits;
                    left -= take;
                }
                if (rel > st.total_rel) fl = 1;
                sh.pos = st.base + rel; sh.fail = fl;
            }
            __syncthreads();
            fail = sh.fail != 0;
            E = sh.pos; pos = E; cnt = B;
        }
        if (fail) break;
        if (threadIdx.x == 0) { rec->end = E; rec->npieces = npieces; }
        pos = E;
    }
    if (fail) {
        // f = the frame that failed (or the first one of this launch): it and everything behind it in this launch decodes nothing
        for (unsigned ff = f0 + threadIdx.x; ff < f1; ff += kVChainThreads) {
            if (ff < f) continue;
            VFrameRec *rec = v.rec + ff;
            rec->first = total; rec->end = total; rec->npieces = 0;
        }
    }
    if (threadIdx.x == 0) {
        v.state[0] = pos; v.state[1] = pool_n; v.state[2] = fail ? 1ull : 0ull;
        v.result[0] = fail ? 0u : 1u; v.result[1] = pool_n;
    }
}

// block_off[img][0 .. nblocks] of frame first_frame + img * frame_step (img = blockIdx.y) from the frame's pieces
__global__ void __launch_bounds__(kVTH) vparse_emit_kernel(const ParseParams p, const VideoParse v, unsigned first_frame, unsigned frame_step,
                                                           unsigned long long *block_off) {
    pdl_wait();
    extern __shared__ __align__(16) unsigned s_stage[];
    const unsigned f = first_frame + blockIdx.y * frame_step;
    const VFrameRec *rec = v.rec + f;
    unsigned long long *off = block_off + (size_t)blockIdx.y * (v.nblocks + 1);
    const unsigned long long total = *p.enc_bits;
    const unsigned long long B0 = *p.start;
    const unsigned np = min(rec->npieces, kVMaxPieces);
    if (blockIdx.x == 0 && threadIdx.x == 0) off[v.nblocks] = rec->end;
    if (np == 0) {                                           // a frame the chain did not reach (failed parse): every block at the frame's end
        for (unsigned j = blockIdx.x * kVTH + threadIdx.x; j < v.nblocks; j += gridDim.x * kVTH) off[j] = rec->end;
        return;
    }
    for (unsigned i = 0; i < np; i++) {
        const VFramePiece pc = rec->piece[i];
        if (pc.kind == 0) {
            for (unsigned j = blockIdx.x * kVTH + threadIdx.x; j < pc.b; j += gridDim.x * kVTH) off[pc.idx_base + j] = v.pool[pc.a + j];
            continue;
        }
        for (unsigned g0 = pc.a + blockIdx.x * kVTH; g0 <= pc.b; g0 += gridDim.x * kVTH) {      // uniform
            const unsigned long long c_start = B0 + (unsigned long long)g0 * kVG;
            __syncthreads();
            const StagedStream st = stage_stream(s_stage, SpecCfg<kVG, kVTH>::kStageWords, p.enc, total, c_start,
                                                 min(total, c_start + (unsigned long long)kVTH * kVG));
            const unsigned g = g0 + threadIdx.x;
            if (g > pc.b) continue;
            const unsigned entry = p.spec_entry[g].x;
            if (entry == kDead) continue;
            unsigned idx = pc.idx_base + (v.pq[g].x - pc.pm);
            const unsigned g_rel = (unsigned)(c_start + (unsigned long long)threadIdx.x * kVG - st.base);
            unsigned rel = g_rel + entry;
            const unsigned lim = min(g_rel + kVG, st.total_rel);
            while (rel < lim && idx < v.nblocks) {
                unsigned nb;
                const unsigned bits = staged_step(st, rel, lim, p.NN, p.use_rle, nb);
                if (nb == 0) { if (p.err) atomicExch(p.err, IE_EFORMAT); break; }
                for (unsigned j = 0; j < nb && idx < v.nblocks; j++) off[idx++] = st.base + rel + 4u * j;
                rel += bits;
            }
        }
    }
}

static void video_parse_layout(const VideoParseSizes &z, uint8_t *scratch, ParseParams &p, VideoParse &v) {
    uint2 *s = reinterpret_cast<uint2 *>(scratch);
    p.spec_entry = s; s += z.nspec;
    p.spec_exit = s; s += z.nspec;
    v.pq = s; s += z.nspec + 1;
    v.partial = s; s += z.scan_ctas;
    p.spec_flags = reinterpret_cast<unsigned *>(s);
    p.cursor_next = reinterpret_cast<unsigned long long *>(p.spec_flags + 8);
    v.result = p.spec_flags + 12;
    p.walk_base = p.spec_flags + 16;
    v.state = reinterpret_cast<unsigned long long *>(p.spec_flags + 24);      // 3 x u64
    uint8_t *b = reinterpret_cast<uint8_t *>(p.spec_flags + 32);
    v.pool = reinterpret_cast<unsigned long long *>(b); b += z.pool_cap * sizeof(unsigned long long);
    v.rec = reinterpret_cast<VFrameRec *>(b);
    v.pool_cap = (unsigned)z.pool_cap;
}

VideoParseSizes video_parse_sizes(size_t enc_bytes, unsigned frames, int sm_count) {
    VideoParseSizes z;
    z.nspec = (enc_bytes * 8 + kVG - 1) / kVG + 1;
    z.scan_ctas = (unsigned)std::max(1, sm_count);
    z.pool_cap = (size_t)frames * 4096 + 65536;
    z.bytes = (2 * z.nspec + z.nspec + 1 + z.scan_ctas) * sizeof(uint2) + 32 * sizeof(unsigned) + z.pool_cap * sizeof(unsigned long long) +
              (size_t)frames * sizeof(VFrameRec) + 256;
    return z;
}

// steps 1-2 (walk + prefix sums); the chain is launched batch by batch with launch_video_chain
int launch_video_parse(const uint8_t *d_enc, const unsigned long long *d_enc_bits, const unsigned long long *d_start, int use_rle,
                       unsigned nblocks, unsigned mv_bits, unsigned frames, unsigned gop, const VideoParseSizes &z, uint8_t *scratch,
                       VideoParse &v, ParseParamsOpaque &popaque, cudaStream_t stream) {
    static_assert(sizeof(ParseParamsOpaque) >= sizeof(ParseParams), "opaque storage too small");
    ParseParams &p = *reinterpret_cast<ParseParams *>(&popaque);
    memset(&p, 0, sizeof p);
    memset(&v, 0, sizeof v);
    p.enc = d_enc; p.enc_bits = d_enc_bits; p.start = d_start; p.skip_bits = 0;
    p.NN = 16; p.use_rle = use_rle; p.E = 4 + 16 + 16 * 16;
    p.nblocks = nblocks;
    p.nspec = (unsigned)z.nspec;
    v.nblocks = nblocks; v.mv_bits = mv_bits; v.frames = frames; v.gop = gop;
    video_parse_layout(z, scratch, p, v);
    if ((uintptr_t)d_enc % 16) { set_error("encoded stream must be 16-byte aligned on the device"); return IE_EINVAL; }
    const unsigned nwalk = (p.nspec + kVTH - 1) / kVTH;
    const size_t stage_bytes = (size_t)SpecCfg<kVG, kVTH>::kStageWords * sizeof(unsigned);
    IE_CUDA(cudaFuncSetAttribute(parse_spec_walk<kVG, kVTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes));
    IE_CUDA(launch_pdl(parse_spec_walk<kVG, kVTH>, dim3(nwalk), dim3(kVTH), stage_bytes, stream, p));
    IE_CUDA(launch_pdl(vparse_scan_kernel, dim3(z.scan_ctas), dim3(256), 0, stream, p, v.pq, v.partial));
    count_launch(2);
    IE_CUDA(cudaGetLastError());
    return IE_OK;
}

// step 3 for frames [f0, f1); launches must be made in frame order on one stream.  After the last one v.result[0] says whether
// the frame records are valid.
int launch_video_chain(const VideoParse &v, const ParseParamsOpaque &popaque, unsigned f0, unsigned f1, cudaStream_t stream) {
    const ParseParams &p = *reinterpret_cast<const ParseParams *>(&popaque);
    IE_CUDA(launch_pdl(vparse_chain_kernel, dim3(1), dim3(kVChainThreads), 0, stream, p, v, f0, f1));
    count_launch();
    IE_CUDA(cudaGetLastError());
    return IE_OK;
}

int launch_video_emit(const VideoParse &v, const ParseParamsOpaque &popaque, unsigned first_frame, unsigned frame_step, unsigned nimg,
                      unsigned long long *block_off, cudaStream_t stream) {
    const ParseParams &p = *reinterpret_cast<const ParseParams *>(&popaque);
    const size_t stage_bytes = (size_t)SpecCfg<kVG, kVTH>::kStageWords * sizeof(unsigned);
    static bool configured = false;
    if (!configured) { IE_CUDA(cudaFuncSetAttribute(vparse_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes)); configured = true; }
    IE_CUDA(launch_pdl(vparse_emit_kernel, dim3(48, nimg), dim3(kVTH), stage_bytes, stream, p, v, first_frame, frame_step, block_off));
    count_launch();
    IE_CUDA(cudaGetLastError());
    return IE_OK;
}

}  // namespace ie
