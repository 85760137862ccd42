// Video path: MacroBlock motion estimation, mvec pack, P-frame residual encode + in-place reconstruction, decode.
//
//   me_search_kernel   warp per 16x16 MacroBlock: the reference's fixed 2-D-log search (Block.cpp:267-339) over the
//                      pattern of algo.cpp:90-139 -- levels merange/2, /4, ..., 1; 9 candidates per level in MER_SIGNS
//                      order around the current best; candidate coordinate clamped into the frame (ImageBase.cpp:253-254);
//                      for p>0 skip when the clamped coordinate is the block's own (Block.cpp:297-301); `<=` so the
//                      later candidate wins ties (Block.cpp:306); the UNCLAMPED offset is stored (Block.cpp:333-334).
//   mvec_pack2_kernel  all motion vectors of the frame, MVEC_BIT_SIZE bits each, x then y (Block.cpp:415-423)
//   P-frame blocks     encode_tiles_kernel<4,4,PF=true> (encode_image.cu)
//   mc_copy_kernel     decoder: out[MB] = ref[clamp(MB + mv)] (Block.cpp:481-496)
// Frames are strictly sequential inside a GOP (each P-frame searches the reconstruction of its predecessor,
// Frame.cpp:210-242); frames append to one stream through the device-resident bit counter (pack.cuh contract).
#include <atomic>
#include <cstring>
#include <vector>

#include "api_internal.cuh"

namespace ie {

__constant__ int c_mer_sx[9] = {0, +1, +1, 0, -1, -1, -1, 0, +1};      // algo.cpp:90-100
__constant__ int c_mer_sy[9] = {0, 0, +1, +1, +1, 0, -1, -1, -1};

struct MEParams {
    const uint8_t *cur;
    const uint8_t *ref;
    int W, H, mx, nmb, merange;
    short *mv;           // [nmb][2] unclamped offsets
    short *res_coord;    // [nmb][2] clamped pixel coordinate of the best block (residual source)
    short *copy_coord;   // [nmb][2] clamp(MB + mv)  (Frame.cpp:218-220)
    // blockIdx.y = GOP of the batch: frame pointers advance by frame_stride bytes, the three arrays by mv_stride shorts
    size_t frame_stride, mv_stride;
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__global__ void __launch_bounds__(256) me_search_kernel(MEParams p) {
#define IE_ME_REDUX 0
#include "me_search_body.inc"
#undef IE_ME_REDUX
}
// ie_set_option("me_variant", 1) (experimental, not yet run on a B200): SAD partial sums reduced with REDUX
__global__ void __launch_bounds__(256) me_search_redux_kernel(MEParams p) {
#define IE_ME_REDUX 1
#include "me_search_body.inc"
#undef IE_ME_REDUX
}

// me_variant 2 (the default): eight lanes per MacroBlock (two pixel rows each: li and li + 8), four horizontally adjacent MacroBlocks per warp
// on one shared search window.  The selection logic, the reduction and the address arithmetic of a level cost the same per
// warp whatever the number of MacroBlocks in it, so four per warp cut the instructions per MacroBlock by ~3x; the window of
// four neighbours (46 rows x 100 pixels) is staged once instead of four times (46 x 52 each).  Same candidates, same order,
// same tie rule as me_search_kernel (Block.cpp:267-339).
__device__ __forceinline__ unsigned sad4_acc(unsigned a, unsigned b, unsigned c) {
    unsigned d;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// One level of the search on the staged window when no candidate of any lane is clamped and the step is 8, 4, 2 or 1: the
// three candidates of a row (x - S, x, x + S) lie in one span of 16 + 2 S bytes, which is read once, aligned to the first
// candidate with one funnel shift per word, and then used at compile-time offsets (shared-memory reads per level: 6 rows of
// 9 / 7 / 6 / 6 words instead of 18 rows of 5 -- the kernel is bound by shared-memory bandwidth).
template <int S>
__device__ __forceinline__ void me_level_span(const unsigned *win, int row_words, int rb0, int rb1, int rb2, int xw0, unsigned sh, const uint4 &c0,
                                              const uint4 &c1, unsigned (&d)[9]) {
    constexpr int W = (16 + 2 * S + 3 + 3) / 4;          // words that hold bytes a .. a + 16 + 2 S, a = 0 .. 3
    unsigned acc[3][3];
#pragma unroll
    for (int jy = 0; jy < 3; jy++)
#pragma unroll
        for (int jx = 0; jx < 3; jx++) acc[jy][jx] = 0;
#pragma unroll
    for (int jy = 0; jy < 3; jy++) {
        const int rb = (jy == 0) ? rb0 : (jy == 1 ? rb1 : rb2);
#pragma unroll
        for (int rr = 0; rr < 2; rr++) {
            const unsigned *wp = win + rb + rr * 8 * row_words + xw0;
            unsigned w[W], n[W - 1];
#pragma unroll
            for (int k = 0; k < W; k++) w[k] = wp[k];
#pragma unroll
            for (int k = 0; k < W - 1; k++) n[k] = __funnelshift_r(w[k], w[k + 1], sh);
            const unsigned cr[4] = {rr ? c1.x : c0.x, rr ? c1.y : c0.y, rr ? c1.z : c0.z, rr ? c1.w : c0.w};
#pragma unroll
            for (int jx = 0; jx < 3; jx++) {
                constexpr int dummy = 0; (void)dummy;
                const int o = jx * S, ow = o >> 2, os = (o & 3) * 8;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const unsigned r = os ? __funnelshift_r(n[ow + k], n[(ow + k + 1 < W - 1) ? ow + k + 1 : W - 2], (unsigned)os) : n[ow + k];
                    acc[jy][jx] = sad4_acc(cr[k], r, acc[jy][jx]);
                }
            }
        }
    }
    constexpr int kSX[9] = {1, 2, 2, 1, 0, 0, 0, 1, 2}, kSY[9] = {1, 1, 2, 2, 2, 1, 0, 0, 0};
#pragma unroll
    for (int q = 0; q < 9; q++) d[q] = acc[kSY[q]][kSX[q]];
}
__global__ void __launch_bounds__(256) me_search8_kernel(MEParams p) {
    constexpr int kWinRows = 46, kWinWords = 27;        // 27: lanes 0..7 of a MacroBlock (rows li, li + 8) and its neighbours spread over the banks
    __shared__ unsigned s_win[8][kWinRows * kWinWords];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gx = (p.mx + 3) >> 2;                                     // groups of four MacroBlocks per MacroBlock row
    const int grp = blockIdx.x * 8 + warp;
    if (grp >= gx * (p.nmb / p.mx)) return;
    p.cur += (size_t)blockIdx.y * p.frame_stride; p.ref += (size_t)blockIdx.y * p.frame_stride;
    p.mv += (size_t)blockIdx.y * p.mv_stride; p.res_coord += (size_t)blockIdx.y * p.mv_stride; p.copy_coord += (size_t)blockIdx.y * p.mv_stride;
    const int gy = grp / gx, g0x = (grp - gy * gx) * 4;
    const int sub = lane >> 3, li = lane & 7;
    const bool live = g0x + sub < p.mx;
    const int mbi = min(g0x + sub, p.mx - 1);                           // lanes of a missing MacroBlock shadow the row's last one
    const int mbx = mbi * kMB, mby = gy * kMB;
    const int mb = gy * p.mx + mbi;
    const uint4 c0 = *reinterpret_cast<const uint4 *>(p.cur + (size_t)(mby + li) * p.W + mbx);          // rows li and li + 8
    const uint4 c1 = *reinterpret_cast<const uint4 *>(p.cur + (size_t)(mby + li + 8) * p.W + mbx);
    const bool staged = p.merange <= 16;
    const int wy0 = max(mby - 15, 0), wxa = max(g0x * kMB - 15, 0) & ~3;
    unsigned *win = s_win[warp];
    if (staged) {
        constexpr int kIters = (kWinRows * kWinWords + 31) / 32, kHalf = (kIters + 1) / 2;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            unsigned v[kHalf];
#pragma unroll
            for (int i = 0; i < kHalf; i++) {                          // half of the loads in flight before the first store
                const int idx = lane + 32 * (h * kHalf + i);
                const int r = idx / kWinWords, k = idx - r * kWinWords;
                const int y = wy0 + r, x = wxa + 4 * k;
                v[i] = (idx < kWinRows * kWinWords && k < 25 && y < p.H && x + 4 <= p.W) ? __ldg(reinterpret_cast<const unsigned *>(p.ref + (size_t)y * p.W + x)) : 0u;
            }
#pragma unroll
            for (int i = 0; i < kHalf; i++) {
                const int idx = lane + 32 * (h * kHalf + i);
                if (idx < kWinRows * kWinWords) win[idx] = v[i];
            }
        }
        __syncwarp();
    }
    int best_x = 0, best_y = 0;
    int bcx = 0, bcy = 0;                                               // Block.cpp:273: the initial block is the one at pixel (0,0)
    unsigned best_d = 0xffffffffu;
    for (int step = p.merange / 2; step > 0; step >>= 1) {              // algo.cpp:129,138
        int cx3[3], cy3[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            cx3[j] = clampi((int)(short)(best_x + (j - 1) * step + mbx), 0, p.W - kMB);
            cy3[j] = clampi((int)(short)(best_y + (j - 1) * step + mby), 0, p.H - kMB);
        }
        constexpr int kSX[9] = {1, 2, 2, 1, 0, 0, 0, 1, 2}, kSY[9] = {1, 1, 2, 2, 2, 1, 0, 0, 0};      // MER_SIGNS, algo.cpp:90-100
        unsigned d[9];
        // span path: every lane's three x (and y) candidates are exactly step apart (nothing clamped)
        const bool regular = staged && (step == 8 || step == 4 || step == 2 || step == 1) && cx3[1] - cx3[0] == step && cx3[2] - cx3[1] == step &&
                             cy3[1] - cy3[0] == step && cy3[2] - cy3[1] == step;
        if (__all_sync(0xffffffffu, regular)) {
            const int bx = cx3[0] - wxa;
            const int xw0 = bx >> 2;
            const unsigned sh = (unsigned)(bx & 3) * 8;
            const int rb0 = (cy3[0] - wy0 + li) * kWinWords, rb1 = (cy3[1] - wy0 + li) * kWinWords, rb2 = (cy3[2] - wy0 + li) * kWinWords;
            if (step == 8) me_level_span<8>(win, kWinWords, rb0, rb1, rb2, xw0, sh, c0, c1, d);
            else if (step == 4) me_level_span<4>(win, kWinWords, rb0, rb1, rb2, xw0, sh, c0, c1, d);
            else if (step == 2) me_level_span<2>(win, kWinWords, rb0, rb1, rb2, xw0, sh, c0, c1, d);
            else me_level_span<1>(win, kWinWords, rb0, rb1, rb2, xw0, sh, c0, c1, d);
        } else if (staged) {
            int xw[3], rb[3];
            unsigned xs[3];
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const int bx = cx3[j] - wxa;
                xw[j] = bx >> 2; xs[j] = (unsigned)(bx & 3) * 8;
                rb[j] = (cy3[j] - wy0 + li) * kWinWords;
            }
#pragma unroll
            for (int q = 0; q < 9; q++) {
                const unsigned *wp = win + rb[kSY[q]] + xw[kSX[q]];
                const unsigned sh = xs[kSX[q]];
                unsigned a[5], b[5];
#pragma unroll
                for (int k = 0; k < 5; k++) { a[k] = wp[k]; b[k] = wp[8 * kWinWords + k]; }     // word 4 is inside the window row (<= 24)
                unsigned acc = 0;
                acc = sad4_acc(c0.x, __funnelshift_r(a[0], a[1], sh), acc);
                acc = sad4_acc(c0.y, __funnelshift_r(a[1], a[2], sh), acc);
                acc = sad4_acc(c0.z, __funnelshift_r(a[2], a[3], sh), acc);
                acc = sad4_acc(c0.w, __funnelshift_r(a[3], a[4], sh), acc);
                acc = sad4_acc(c1.x, __funnelshift_r(b[0], b[1], sh), acc);
                acc = sad4_acc(c1.y, __funnelshift_r(b[1], b[2], sh), acc);
                acc = sad4_acc(c1.z, __funnelshift_r(b[2], b[3], sh), acc);
                acc = sad4_acc(c1.w, __funnelshift_r(b[3], b[4], sh), acc);
                d[q] = acc;                                                               // Block.cpp:241-254 (this lane's 32 pixels)
            }
        } else {
#pragma unroll
            for (int q = 0; q < 9; q++) {
                const uint8_t *rp = p.ref + (size_t)(cy3[kSY[q]] + li) * p.W + cx3[kSX[q]];
                const uintptr_t ad = (uintptr_t)rp;
                const unsigned *wp = reinterpret_cast<const unsigned *>(ad & ~(uintptr_t)3);
                const unsigned sh = (unsigned)(ad & 3) * 8;
                const unsigned *wq = wp + 8 * (p.W >> 2);
                unsigned a[5], b[5];
#pragma unroll
                for (int k = 0; k < 4; k++) { a[k] = __ldg(wp + k); b[k] = __ldg(wq + k); }
                a[4] = sh ? __ldg(wp + 4) : 0u; b[4] = sh ? __ldg(wq + 4) : 0u;
                unsigned acc = 0;
                acc = sad4_acc(c0.x, __funnelshift_r(a[0], a[1], sh), acc);
                acc = sad4_acc(c0.y, __funnelshift_r(a[1], a[2], sh), acc);
                acc = sad4_acc(c0.z, __funnelshift_r(a[2], a[3], sh), acc);
                acc = sad4_acc(c0.w, __funnelshift_r(a[3], a[4], sh), acc);
                acc = sad4_acc(c1.x, __funnelshift_r(b[0], b[1], sh), acc);
                acc = sad4_acc(c1.y, __funnelshift_r(b[1], b[2], sh), acc);
                acc = sad4_acc(c1.z, __funnelshift_r(b[2], b[3], sh), acc);
                acc = sad4_acc(c1.w, __funnelshift_r(b[3], b[4], sh), acc);
                d[q] = acc;
            }
        }
        // a MacroBlock's SAD is <= 256 * 255 < 2^16: two partial sums per register, reduced over the MacroBlock's eight lanes
        unsigned pk[5] = {d[0] | (d[1] << 16), d[2] | (d[3] << 16), d[4] | (d[5] << 16), d[6] | (d[7] << 16), d[8]};
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
#pragma unroll
            for (int k = 0; k < 5; k++) pk[k] += __shfl_xor_sync(0xffffffffu, pk[k], o);
        }
        int bq = -1, ncx = 0, ncy = 0;
        unsigned nd = best_d;
#pragma unroll
        for (int q = 0; q < 9; q++) {
            const int cpx = cx3[kSX[q]], cpy = cy3[kSY[q]];
            if (q > 0 && cpx == mbx && cpy == mby) continue;                   // Block.cpp:297-301
            const unsigned dq = (q == 8) ? pk[4] : ((q & 1) ? (pk[q >> 1] >> 16) : (pk[q >> 1] & 0xffffu));
            if (dq <= nd) { bq = q; nd = dq; ncx = cpx; ncy = cpy; }           // Block.cpp:306
        }
        // Block.cpp:318-321: no candidate at all -- cannot happen (candidate 0 is never skipped and d <= best_d holds for it:
        // the centre of a level is the previous level's best); lanes must stay together for the shuffles, so no early exit
        if (bq >= 0) { best_x += c_mer_sx[bq] * step; best_y += c_mer_sy[bq] * step; best_d = nd; bcx = ncx; bcy = ncy; }
    }
    if (li == 0 && live) {
        p.mv[2 * mb] = (short)best_x;
        p.mv[2 * mb + 1] = (short)best_y;
        p.res_coord[2 * mb] = (short)bcx;
        p.res_coord[2 * mb + 1] = (short)bcy;
        p.copy_coord[2 * mb] = (short)clampi((int)(short)(mbx + best_x), 0, p.W - kMB);
        p.copy_coord[2 * mb + 1] = (short)clampi((int)(short)(mby + best_y), 0, p.H - kMB);
    }
}
std::atomic<int> g_me_variant{2};
// 2 = the GOPs of a batch in two halves on two streams (default), 1 = all on the caller's stream
std::atomic<int> g_video_encode_streams{2};

// fixed-width fields: field i = low `bits` bits of val[i]
struct FixedFieldTile {
    const short *val;
    unsigned nfields;
    unsigned bits;
};

__device__ __forceinline__ uint4 gather_chunk(const FixedFieldTile &t, long long ls) {
    unsigned ow[4] = {0, 0, 0, 0};
    const unsigned long long total = (unsigned long long)t.nfields * t.bits;
#pragma unroll
    for (int wd = 0; wd < 4; wd++) {
        unsigned word = 0;
        // bits [ls + 32 wd, +32): walk the fields that overlap
        long long b0 = ls + 32 * wd;
        for (int done = 0; done < 32;) {
            const long long pos = b0 + done;
            if (pos < 0) { done += (int)min((long long)(32 - done), -pos); continue; }
            if ((unsigned long long)pos >= total) break;
            const unsigned f = (unsigned)(pos / t.bits), fo = (unsigned)(pos % t.bits);
            const unsigned take = min(t.bits - fo, (unsigned)(32 - done));
            const unsigned v = ((unsigned)(int)t.val[f] & ((1u << t.bits) - 1u)) >> (t.bits - fo - take) & ((1u << take) - 1u);
            word |= v << (32 - done - take);
            done += (int)take;
        }
        ow[wd] = __byte_perm(word, 0, 0x0123);
    }
    return make_uint4(ow[0], ow[1], ow[2], ow[3]);
}

// All motion vectors of the frame, MVEC_BIT_SIZE bits each, x then y (Block.cpp:415-423): a thread per 128-bit chunk of the
// motion-vector section over as many CTAs as it takes (the round-1 kernel walked the section with one CTA per GOP: 43 us per
// frame slot on config 5).  The chunk that holds the section's first bit already holds
// the end of the previous frame: its owner merges (stream order: that frame's copy-out has completed); bits past the section
// in the last chunk are zero, the tile copy-out that follows merges into it.  The last CTA of a GOP to finish advances the
// GOP's bit counter (ticket; every CTA has read the counter before it takes one) and resets the ticket.
__global__ void __launch_bounds__(256) mvec_pack2_kernel(const short *mv, size_t mv_stride, unsigned nfields, unsigned bits, uint8_t *out,
                                                         size_t out_stride, size_t out_cap, unsigned long long *bit_counter, unsigned *ticket, int *err) {
    mv += (size_t)blockIdx.y * mv_stride; out += (size_t)blockIdx.y * out_stride; bit_counter += blockIdx.y; ticket += blockIdx.y;
    const FixedFieldTile t{mv, nfields, bits};
    const unsigned long long G = *bit_counter;
    const unsigned T = nfields * bits;
    if (T != 0) {
        const unsigned long long c0 = G / kChunkBits, c1 = (G + T - 1) / kChunkBits;
        const unsigned long long c = c0 + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
        if (c <= c1) {
            if ((c + 1) * 16ull > out_cap) { if (err) atomicExch(err, IE_ENOSPC); }
            else {
                uint4 v = gather_chunk(t, (long long)(c * kChunkBits) - (long long)G);
                uint4 *dst = reinterpret_cast<uint4 *>(out) + c;
                if (c == c0 && (G % kChunkBits) != 0) { const uint4 o = *dst; v.x |= o.x; v.y |= o.y; v.z |= o.z; v.w |= o.w; }
                *dst = v;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) { *bit_counter = G + T; *ticket = 0u; }
    }
}

struct MCParams {
    const uint8_t *enc;
    unsigned long long enc_bits;
    const unsigned long long *cursor;   // frame's first bit
    unsigned mvbits;
    const uint8_t *ref;
    uint8_t *cur;
    int W, H, mx, nmb;
};

__device__ __forceinline__ unsigned read_bits_dev(const uint8_t *s, unsigned long long total_bits, unsigned long long p, int n) {
    if (n == 0) return 0u;
    const unsigned long long nbytes = (total_bits + 7) >> 3, b = p >> 3;
    unsigned v = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) v = (v << 8) | ((b + i < nbytes) ? (unsigned)__ldg(s + b + i) : 0u);
    return (v >> (32 - (int)(p & 7) - n)) & ((1u << n) - 1u);
}

__global__ void __launch_bounds__(256) mc_copy_kernel(const MCParams p) {
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mb = blockIdx.x * 8 + warp;
    if (mb >= p.nmb) return;
    const int mbx = (mb % p.mx) * kMB, mby = (mb / p.mx) * kMB;
    const unsigned long long base = *p.cursor + (unsigned long long)mb * 2 * p.mvbits;
    const int sh = 16 - (int)p.mvbits;
    const int vx = (int)(short)(unsigned short)(read_bits_dev(p.enc, p.enc_bits, min(base, p.enc_bits), p.mvbits) << sh) >> sh;   // Block.cpp:484-485
    const int vy = (int)(short)(unsigned short)(read_bits_dev(p.enc, p.enc_bits, min(base + p.mvbits, p.enc_bits), p.mvbits) << sh) >> sh;
    const int cx = clampi((int)(short)(mbx + vx), 0, p.W - kMB), cy = clampi((int)(short)(mby + vy), 0, p.H - kMB);
    const int row = lane >> 1, half = lane & 1;
    const uint8_t *rp = p.ref + (size_t)(cy + row) * p.W + cx + half * 8;
    uint8_t *dp = p.cur + (size_t)(mby + row) * p.W + mbx + half * 8;
    unsigned lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { lo |= (unsigned)__ldg(rp + k) << (8 * k); hi |= (unsigned)__ldg(rp + 4 + k) << (8 * k); }
    *reinterpret_cast<uint2 *>(dp) = make_uint2(lo, hi);
}

// the same for frame `first_frame + blockIdx.y * gop` of a whole-stream decode: the frame's motion vectors end where its first
// block starts (VFrameRec::first), the reference frame is the one before it in the output
struct MCBatchParams {
    const uint8_t *enc;
    unsigned long long enc_bits;
    const VFrameRec *rec;
    unsigned first_frame, gop, mvbits;
    uint8_t *out;
    size_t fsz;
    int W, H, mx, nmb;
};
__global__ void __launch_bounds__(256) mc_copy_batch_kernel(const MCBatchParams p) {
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mb = blockIdx.x * 8 + warp;
    if (mb >= p.nmb) return;
    const unsigned f = p.first_frame + blockIdx.y * p.gop;
    const uint8_t *ref = p.out + (size_t)(f - 1) * p.fsz;
    uint8_t *cur = p.out + (size_t)f * p.fsz;
    const int mbx = (mb % p.mx) * kMB, mby = (mb / p.mx) * kMB;
    const unsigned long long base = p.rec[f].first - (unsigned long long)p.nmb * 2 * p.mvbits + (unsigned long long)mb * 2 * p.mvbits;
    const int sh = 16 - (int)p.mvbits;
    const int vx = (int)(short)(unsigned short)(read_bits_dev(p.enc, p.enc_bits, min(base, p.enc_bits), p.mvbits) << sh) >> sh;   // Block.cpp:484-485
    const int vy = (int)(short)(unsigned short)(read_bits_dev(p.enc, p.enc_bits, min(base + p.mvbits, p.enc_bits), p.mvbits) << sh) >> sh;
    const int cx = clampi((int)(short)(mbx + vx), 0, p.W - kMB), cy = clampi((int)(short)(mby + vy), 0, p.H - kMB);
    const int row = lane >> 1, half = lane & 1;
    const uint8_t *rp = ref + (size_t)(cy + row) * p.W + cx + half * 8;
    uint8_t *dp = cur + (size_t)(mby + row) * p.W + mbx + half * 8;
    // 8 source bytes at any alignment from three aligned words
    const uintptr_t a = reinterpret_cast<uintptr_t>(rp);
    const unsigned *wp = reinterpret_cast<const unsigned *>(a & ~(uintptr_t)3);
    const unsigned sel = 0x3210u + 0x1111u * (unsigned)(a & 3);
    const unsigned w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = (a & 3) ? __ldg(wp + 2) : 0u;
    *reinterpret_cast<uint2 *>(dp) = make_uint2(__byte_perm(w0, w1, sel), __byte_perm(w1, w2, sel));
}

// the same without the copy: clamp(MB + mv) of every MacroBlock, for the block decoder that reads its prediction from the
// reference frame itself (DecodeParams::mc_coord)
__global__ void __launch_bounds__(256) mc_coord_batch_kernel(const MCBatchParams p, short *coord) {
    pdl_wait();
    const int mb = blockIdx.x * 256 + threadIdx.x;
    if (mb >= p.nmb) return;
    const unsigned f = p.first_frame + blockIdx.y * p.gop;
    const int mbx = (mb % p.mx) * kMB, mby = (mb / p.mx) * kMB;
    const unsigned long long base = p.rec[f].first - (unsigned long long)p.nmb * 2 * p.mvbits + (unsigned long long)mb * 2 * p.mvbits;
    const int sh = 16 - (int)p.mvbits;
    const int vx = (int)(short)(unsigned short)(read_bits_dev(p.enc, p.enc_bits, min(base, p.enc_bits), p.mvbits) << sh) >> sh;   // Block.cpp:484-485
    const int vy = (int)(short)(unsigned short)(read_bits_dev(p.enc, p.enc_bits, min(base + p.mvbits, p.enc_bits), p.mvbits) << sh) >> sh;
    short *c = coord + ((size_t)blockIdx.y * p.nmb + mb) * 2;
    c[0] = (short)clampi((int)(short)(mbx + vx), 0, p.W - kMB);
    c[1] = (short)clampi((int)(short)(mby + vy), 0, p.H - kMB);
}

__global__ void fill_uv_kernel(uint8_t *yuv, size_t ysz, size_t fsz, unsigned frames) {
    const size_t uv = fsz - ysz;
    const size_t total = uv * frames;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total / 16; i += (size_t)gridDim.x * blockDim.x) {
        const size_t byte = i * 16;
        const size_t f = byte / uv, o = byte % uv;
        *reinterpret_cast<uint4 *>(yuv + f * fsz + ysz + o) = make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);
    }
}


static unsigned host_bits_needed(int v) {                    // utils.hpp:226-243
    const short value = (short)v;
    unsigned bits = 1;
    while ((short)((short)((value & ((1 << bits) - 1)) << (16 - bits)) >> (16 - bits)) != value) bits++;
    return bits;
}

static int check_video_dims(uint32_t W, uint32_t H) {
    IE_TRY(check_dims(W, H, 4));
    if (W % kMB || H % kMB) {
        set_error("video width/height must be multiples of 16: micro blocks outside a MacroBlock are never coded (Block.cpp:373-375)");
        return IE_EINVAL;
    }
    return IE_OK;
}

// scratch layout inside s->d_scratch for video: mv | res_coord | copy_coord (shorts), cursor (u64)
// ---------------------------------------------------------------------------------------------------------
// GOP batches.  GOPs are independent (VideoBase.hpp:32, VideoBase.cpp:105-118), so frame k of every GOP of a batch is
// encoded by ONE launch (blockIdx.y = GOP) into that GOP's own stream buffer; after the last frame the GOP streams are
// appended, in order, to the video stream (one offsets kernel + one copy kernel).
// ---------------------------------------------------------------------------------------------------------
// off[j] = first bit of GOP j in the video stream; chunks two GOPs share are zeroed so that both can OR their part in.
__global__ void gop_offsets_kernel(const unsigned long long *gop_bits, unsigned ngops, unsigned long long *stream_bits,
                                   unsigned long long *off, uint8_t *out, size_t out_cap) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned long long pos = *stream_bits;
    for (unsigned j = 0; j < ngops; j++) {
        off[j] = pos;
        if (j > 0 && pos % kChunkBits && (pos / kChunkBits + 1) * 16 <= out_cap)
            reinterpret_cast<uint4 *>(out)[pos / kChunkBits] = make_uint4(0u, 0u, 0u, 0u);
        pos += gop_bits[j];
    }
    *stream_bits = pos;
}

// GOP stream j (bit 0 of its buffer, gop_bits[j] bits, zero padded to a chunk) -> bits [off[j], off[j] + gop_bits[j]) of out.
// Thread per 128-bit chunk of the output; chunks shared with a neighbour (or with what earlier launches wrote: the header,
// the previous batch) are merged with atomicOr, the rest are plain 128-bit stores.
__global__ void __launch_bounds__(256) gop_append_kernel(const uint8_t *gop_streams, size_t gop_stride, const unsigned long long *gop_bits,
                                                         const unsigned long long *off, unsigned ngops, uint8_t *out, size_t out_cap, int *err) {
    const unsigned j = blockIdx.y;
    const unsigned long long G = off[j], T = gop_bits[j];
    if (T == 0) return;
    const unsigned *w = reinterpret_cast<const unsigned *>(gop_streams + (size_t)j * gop_stride);
    const long long nwords = (long long)((T + 31) / 32);
    const unsigned long long c0 = G / kChunkBits, c1 = (G + T - 1) / kChunkBits;
    for (unsigned long long c = c0 + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; c <= c1; c += (unsigned long long)gridDim.x * blockDim.x) {
        const long long ls = (long long)(c * kChunkBits) - (long long)G;      // first GOP bit of this chunk (may be < 0)
        const long long wi = ls >> 5;
        const unsigned sh = (unsigned)(ls & 31);
        unsigned x[5];
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const long long i = wi + k;
            x[k] = (i >= 0 && i < nwords) ? __byte_perm(__ldg(w + i), 0, 0x0123) : 0u;
        }
        // the last word may hold pad bits beyond T: they are zero (the encoder zero-pads its last chunk)
        uint4 v;
        v.x = __byte_perm(__funnelshift_l(x[1], x[0], sh), 0, 0x0123);
        v.y = __byte_perm(__funnelshift_l(x[2], x[1], sh), 0, 0x0123);
        v.z = __byte_perm(__funnelshift_l(x[3], x[2], sh), 0, 0x0123);
        v.w = __byte_perm(__funnelshift_l(x[4], x[3], sh), 0, 0x0123);
        if ((c + 1) * 16ull > out_cap) { if (err) atomicExch(err, IE_ENOSPC); continue; }
        const bool shared = (c == c0 && (G % kChunkBits) != 0) || (c == c1 && ((G + T) % kChunkBits) != 0 && j + 1 < ngops);
        if (shared) {
            unsigned *d = reinterpret_cast<unsigned *>(out) + c * 4;
            atomicOr(d + 0, v.x); atomicOr(d + 1, v.y); atomicOr(d + 2, v.z); atomicOr(d + 3, v.w);
        } else {
            reinterpret_cast<uint4 *>(out)[c] = v;
        }
    }
}

// Copy / compute pipeline of the host entry point ie_encode_video: the device entry point encodes the clip in batches of at
// most max_batch GOPs and calls before_batch (wait for the upload of the batch's frames) / after_batch (the stream up to here
// is final: read the cursor back, let the download start) around each.
struct VideoEncodeHooks {
    uint32_t max_batch;
    uint32_t used_batch;      // out: the GOPs per batch the device entry point settled on (<= max_batch)
    int (*before_batch)(void *ctx, uint32_t g0, uint32_t nb, cudaStream_t st);
    int (*after_batch)(void *ctx, uint32_t index, const unsigned long long *d_cursor, cudaStream_t st);
    void *ctx;
};

// ie_decode_video's side of the same idea: once a GOP batch is reconstructed (chroma filled), its frames can go down.
struct VideoDecodeHooks {
    int (*after_batch)(void *ctx, uint32_t index, size_t byte0, size_t byte1, cudaStream_t st);
    void *ctx;
    bool whole_done;          // out: the whole-stream path produced the frames (the early copies are valid)
};

struct VideoScratch { short *mv, *res, *copy; unsigned long long *cursor, *gop_off; unsigned *ticket; };
// motion vectors / residual / copy coordinates of `ngops` frames in flight ([gop][nmb][2] each), the stream's bit cursor and
// the GOP offsets of a batch
static int video_scratch(ie_session *s, size_t nmb, size_t ngops, VideoScratch &v) {
    const size_t sec = (nmb * 2 * sizeof(short) * ngops + 63) / 64 * 64;
    IE_TRY(session_reserve(&s->d_scratch, &s->scratch_cap, 3 * sec + 64 + ngops * (sizeof(unsigned long long) + sizeof(unsigned))));
    v.mv = reinterpret_cast<short *>(s->d_scratch);
    v.res = reinterpret_cast<short *>(s->d_scratch + sec);
    v.copy = reinterpret_cast<short *>(s->d_scratch + 2 * sec);
    v.cursor = reinterpret_cast<unsigned long long *>(s->d_scratch + 3 * sec);
    v.gop_off = v.cursor + 8;
    v.ticket = reinterpret_cast<unsigned *>(v.gop_off + ngops);
    return IE_OK;
}

}  // namespace ie

using namespace ie;

extern "C" {

int ie_encode_video_dev(ie_session *s, uint8_t *d_yuv, size_t yuv_bytes, uint32_t W, uint32_t H, const uint16_t *quant, int use_rle,
                        uint32_t gop, uint32_t merange, int lead_bit, uint8_t *d_out, size_t out_cap, uint64_t *d_out_bits,
                        int16_t *d_mvecs, void *stream) {
    if (!s || !d_yuv || !d_out) { set_error("NULL argument"); return IE_EINVAL; }
    IE_TRY(check_video_dims(W, H));
    IE_TRY(check_quant(quant, 4));
    if ((uintptr_t)d_out % 16) { set_error("stream buffers must be 16-byte aligned"); return IE_EINVAL; }
    if (merange > 32767) { set_error("merange must fit 15 bits"); return IE_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    if (gop < 1) gop = 1;                                                      // VideoBase.cpp:34
    const size_t ysz = (size_t)W * H, fsz = ysz + ysz / 2;
    const uint32_t frames = (uint32_t)(yuv_bytes / fsz);                      // VideoBase.cpp:39-40
    if (frames > 32767) { set_error("more than 32767 frames do not fit the 15-bit header field"); return IE_EINVAL; }
    const unsigned nblocks = (W / 4) * (H / 4), nmb = (W / kMB) * (H / kMB);
    const unsigned TB = encode_tile_blocks(4), tiles = (nblocks + TB - 1) / TB;
    const unsigned mvbits = host_bits_needed((int)(short)merange);           // VideoBase.cpp:42
    // GOP batches: at most kMaxGopBatch GOPs (and ~2 GiB of per-GOP stream buffers) in flight
    const uint32_t ngops = (frames + gop - 1) / gop;
    const size_t frame_cap_bits = (size_t)nblocks * (4 + 16 + 16 * 16) + (size_t)nmb * 2 * mvbits;
    const size_t gop_cap = ((frame_cap_bits * std::min<size_t>(gop, std::max<uint32_t>(frames, 1)) + 127) / 128 + 2) * 16;
    constexpr uint32_t kMaxGopBatch = 64;
    const uint32_t mem_batch = (uint32_t)std::max<size_t>(1, ((size_t)2 << 30) / gop_cap);
    uint32_t batch = std::max<uint32_t>(1, std::min(std::min(ngops, kMaxGopBatch), mem_batch));
    VideoEncodeHooks *hooks = static_cast<VideoEncodeHooks *>(s->video_hooks);
    if (hooks) { batch = std::max<uint32_t>(1, std::min(batch, hooks->max_batch)); hooks->used_batch = batch; }
    IE_TRY(session_ensure_scan(s, batch, tiles));
    IE_TRY(session_ensure_err(s));
    VideoScratch vs;
    IE_TRY(video_scratch(s, nmb, batch, vs));
    IE_TRY(session_reserve(&s->d_tmp, &s->d_tmp_cap, (size_t)batch * gop_cap));          // the GOP streams of a batch
    IE_CUDA(cudaMemsetAsync(vs.ticket, 0, batch * sizeof(unsigned), st));

    HeaderParam hdr, nohdr;
    memset(&nohdr, 0, sizeof nohdr);
    if (s->video_no_header) hdr = nohdr;         // a later GOP shard of a clip: its stream starts with its first frame
    else IE_TRY(build_header(hdr, 4, quant, use_rle, W, H, lead_bit, 1, s->header_frames ? s->header_frames : frames, gop, merange));
    if (out_cap < 256) { set_error("output buffer too small"); return IE_ENOSPC; }
    IE_TRY(launch_stream_init(d_out, 0, 1, hdr, 0, vs.cursor, st));

    EncodeParams p;
    memset(&p, 0, sizeof p);
    p.pitch = W; p.bx = W / 4; p.nblocks = nblocks; p.tiles_per_image = tiles; p.use_rle = use_rle ? 1 : 0;
    make_quant(p.quant, quant, 4);
    FastQuant fq_i, fq_p;
    make_fast_quant(fq_i, quant, 4, 128.0);      // I-frames: pixel - 128
    make_fast_quant(fq_p, quant, 4, 383.0);      // P-frames: (pixel - ref) - 128 in [-383, 127]
    p.dc_den2 = 8 * (int)quant[0]; p.dc_rcp = 1.0f / (float)p.dc_den2;
    p.tab = s->dev->d_t4;
    p.out = s->d_tmp; p.out_stride = gop_cap; p.out_cap = gop_cap; p.bit_counter = s->d_counter; p.err = s->d_err;
    p.mbx = W / kMB;
    p.img_stride = (size_t)gop * fsz;            // image i of a launch = the same frame index of GOP i
    p.coord_stride = (size_t)nmb * 2;
    { float k2[kMaxNN]; make_k2(k2, quant, 4); for (int i = 0; i < 16; i++) p.k2[i] = k2[i]; }
    p.slot_bytes = encode_tile_slot_bytes(4);
    const size_t ntot = (size_t)batch * tiles;
    IE_TRY(session_reserve(&s->d_tile_scratch, &s->tile_scratch_cap, ntot * p.slot_bytes));
    IE_TRY(session_reserve(&s->d_tile_meta, &s->tile_meta_cap, ntot * (sizeof(unsigned long long) + sizeof(unsigned)) + 64));
    p.tile_scratch = s->d_tile_scratch;
    p.bit_base = reinterpret_cast<unsigned long long *>(s->d_tile_meta);
    p.tile_bits = reinterpret_cast<unsigned *>(s->d_tile_meta + ntot * sizeof(unsigned long long));
    // One frame slot of `act` GOPs starting at GOP `gb` of the batch (every per-GOP array is addressed from its element gb), on
    // stream `sx`.  Two halves of a batch run on two streams: the P-frame tiles of one half (issue-bound) next to the motion
    // search of the other (shared-memory / XU bound), and one half's tail next to the other's next kernel.
    auto launch_slot = [&](uint32_t g0, uint32_t gb, uint32_t act, uint32_t k, cudaStream_t sx) -> int {
        EncodeParams q = p;
        const uint32_t f = (g0 + gb) * gop + k;                                          // frame k of the half's first GOP
        uint8_t *cur = d_yuv + (size_t)f * fsz;
        q.src = cur;
        q.out = s->d_tmp + (size_t)gb * gop_cap;
        q.bit_counter = s->d_counter + gb;
        q.tile_scratch = p.tile_scratch + (size_t)gb * tiles * p.slot_bytes;
        q.bit_base = p.bit_base + (size_t)gb * tiles;
        q.tile_bits = p.tile_bits + (size_t)gb * tiles;
        q.scan = s->scan_state();
        q.scan.tile_state += (size_t)gb * tiles; q.scan.bnd += (size_t)gb * tiles; q.scan.ticket += gb;
        short *mv = vs.mv + (size_t)gb * p.coord_stride, *res = vs.res + (size_t)gb * p.coord_stride, *cpy = vs.copy + (size_t)gb * p.coord_stride;
        if (k == 0) {                                                                    // VideoBase.hpp:32, Frame.cpp:130-159
            q.fq = fq_i;
            IE_TRY(launch_encode_tiles(4, q, act, sx));
            if (d_mvecs) IE_CUDA(cudaMemset2DAsync(d_mvecs + (size_t)f * nmb * 2, (size_t)gop * nmb * 2 * sizeof(short), 0,
                                                   nmb * 2 * sizeof(short), act, sx));
            return IE_OK;
        }
        MEParams me;
        me.cur = cur; me.ref = cur - fsz; me.W = (int)W; me.H = (int)H; me.mx = (int)(W / kMB); me.nmb = (int)nmb;
        me.merange = (int)merange; me.mv = mv; me.res_coord = res; me.copy_coord = cpy;
        me.frame_stride = p.img_stride; me.mv_stride = p.coord_stride;
        if (g_me_variant.load() == 2) me_search8_kernel<<<dim3((((W / kMB + 3) / 4) * (H / kMB) + 7) / 8, act), 256, 0, sx>>>(me);
        else if (g_me_variant.load() == 1) me_search_redux_kernel<<<dim3((nmb + 7) / 8, act), 256, 0, sx>>>(me);
        else me_search_kernel<<<dim3((nmb + 7) / 8, act), 256, 0, sx>>>(me);
        count_launch();
        {
            const unsigned mv_chunks = (nmb * 2 * mvbits + 127) / 128 + 1;
            mvec_pack2_kernel<<<dim3((mv_chunks + 255) / 256, act), 256, 0, sx>>>(mv, p.coord_stride, nmb * 2, mvbits, q.out, gop_cap, gop_cap,
                                                                                 q.bit_counter, vs.ticket + gb, s->d_err);
        }
        count_launch();
        IE_CUDA(cudaGetLastError());
        if (d_mvecs) IE_CUDA(cudaMemcpy2DAsync(d_mvecs + (size_t)f * nmb * 2, (size_t)gop * nmb * 2 * sizeof(short), mv,
                                               nmb * 2 * sizeof(short), nmb * 2 * sizeof(short), act, cudaMemcpyDeviceToDevice, sx));
        q.fq = fq_p;
        q.ref = me.ref; q.res_coord = res; q.copy_coord = cpy; q.cur_rw = cur;
        return launch_pframe_tiles(q, act, sx);
    };
    const bool two_streams = g_video_encode_streams.load() == 2;
    cudaStream_t st2 = st;
    if (two_streams) { IE_TRY(session_ensure_pipeline(s)); st2 = s->stream_aux; }
    for (uint32_t g0 = 0; g0 < ngops; g0 += batch) {
        const uint32_t nb = std::min(batch, ngops - g0);
        if (hooks) IE_TRY(hooks->before_batch(hooks->ctx, g0, nb, st));
        IE_TRY(launch_stream_init(s->d_tmp, gop_cap, nb, nohdr, 0, s->d_counter, st));   // empty GOP streams, counters = 0
        // halves: GOPs [0, nA) on the caller's stream, [nA, nb) on the second one (fork after the initialisation, join before
        // the GOP streams are appended)
        const uint32_t nA = (two_streams && nb >= 2) ? (nb + 1) / 2 : nb;
        if (nA < nb) {
            IE_CUDA(cudaEventRecord(s->ev_aux_fork, st));
            IE_CUDA(cudaStreamWaitEvent(st2, s->ev_aux_fork, 0));
        }
        const uint32_t last_len = frames - (g0 + nb - 1) * gop;                          // frames of the batch's last GOP (>= 1)
        for (uint32_t k = 0; k < gop; k++) {
            // GOPs of this batch that have a frame k (only the clip's last GOP can be short)
            const uint32_t act = (k < last_len) ? nb : nb - 1;
            if (act == 0) break;
            const uint32_t actA = std::min(act, nA), actB = act - actA;
            IE_TRY(launch_slot(g0, 0, actA, k, st));
            if (actB) IE_TRY(launch_slot(g0, nA, actB, k, st2));
        }
        if (nA < nb) {
            IE_CUDA(cudaEventRecord(s->ev_aux_join, st2));
            IE_CUDA(cudaStreamWaitEvent(st, s->ev_aux_join, 0));
        }
        // the batch's GOP streams, in order, onto the video stream
        gop_offsets_kernel<<<1, 32, 0, st>>>(s->d_counter, nb, vs.cursor, vs.gop_off, d_out, out_cap);
        gop_append_kernel<<<dim3(s->dev->sm_count * 2, nb), 256, 0, st>>>(s->d_tmp, gop_cap, s->d_counter, vs.gop_off, nb, d_out, out_cap, s->d_err);
        count_launch(2);
        IE_CUDA(cudaGetLastError());
        if (hooks) IE_TRY(hooks->after_batch(hooks->ctx, g0 / batch, vs.cursor, st));
    }
    if (frames == 0) { /* header only */ }
    if (d_out_bits) IE_CUDA(cudaMemcpyAsync(d_out_bits, vs.cursor, sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
    return IE_OK;
}

int ie_session_set_video_shard(ie_session *s, uint32_t total_frames, int write_header) {
    if (!s || total_frames > 32767) { set_error("bad argument"); return IE_EINVAL; }
    s->header_frames = total_frames;
    s->video_no_header = write_header ? 0 : 1;
    return IE_OK;
}

int ie_encode_video(const uint8_t *yuv, size_t yuv_bytes, uint32_t W, uint32_t H, const uint16_t *quant, int use_rle, uint32_t gop,
                    uint32_t merange, int huffman, uint8_t *out, size_t out_cap, size_t *out_bytes) {
    if (!yuv || !out || !out_bytes) { set_error("NULL argument"); return IE_EINVAL; }
    IE_TRY(check_video_dims(W, H));
    if (gop < 1) gop = 1;                                                      // VideoBase.cpp:34 (as ie_encode_video_dev)
    const size_t fsz = (size_t)W * H * 3 / 2;
    const uint32_t frames = (uint32_t)(yuv_bytes / fsz);
    SessionLease lease;
    IE_TRY(lease.acquire(2, W, H, 4, frames));
    ie_session *s = lease.get();
    const size_t cap = ie_max_encoded_bytes(W, H, 4, std::max(1u, frames));
    IE_TRY(session_reserve(&s->d_in, &s->d_in_cap, std::max<size_t>(yuv_bytes, 16)));
    const size_t cap16 = (cap + 15) / 16 * 16;
    IE_TRY(session_reserve(&s->d_out, &s->d_out_cap, cap16 + 64));
    uint64_t *d_total = reinterpret_cast<uint64_t *>(s->d_out + cap16);           // the stream's bit count, behind the stream
    cudaStream_t st = s->stream;
    IE_TRY(session_ensure_pipeline(s));
    // Copy / compute pipeline: the clip goes up in batches of GOPs on stream_in, a batch is encoded as soon as it has arrived,
    // and -- without Huffman -- the part of the stream that a batch completed goes down on stream_out while later batches are
    // still on their way up / being encoded.  PCIe runs in both directions at once.
    const uint32_t ngops = (frames + gop - 1) / gop;
    const uint32_t want = std::min<uint32_t>(std::max<uint32_t>(ngops, 1), 8);                // batches
    const uint32_t gpb = std::max<uint32_t>(1, (ngops + want - 1) / want);                   // GOPs per batch
    const uint32_t nbatches = ngops ? (ngops + gpb - 1) / gpb : 0;
    struct Ctx { ie_session *s; uint32_t gpb; } ctx{s, gpb};
    VideoEncodeHooks hooks;
    hooks.max_batch = gpb;
    hooks.ctx = &ctx;
    hooks.used_batch = 0;
    hooks.before_batch = [](void *c, uint32_t g0, uint32_t nb, cudaStream_t stx) -> int {
        Ctx *x = static_cast<Ctx *>(c);
        // uploads complete in order: the event of the upload batch that holds this batch's last GOP covers all of it
        IE_CUDA(cudaStreamWaitEvent(stx, x->s->ev_in[((g0 + nb - 1) / x->gpb) % ie_session::kMaxStripes], 0));
        return IE_OK;
    };
    hooks.after_batch = [](void *c, uint32_t index, const unsigned long long *d_cursor, cudaStream_t stx) -> int {
        Ctx *x = static_cast<Ctx *>(c);
        IE_CUDA(cudaMemcpyAsync(&x->s->h_pinned[16 + index % 32], d_cursor, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stx));
        IE_CUDA(cudaEventRecord(x->s->ev_done[index % ie_session::kMaxStripes], stx));
        return IE_OK;
    };
    for (uint32_t b = 0; b < nbatches; b++) {
        const size_t f0 = (size_t)b * gpb * gop, f1 = std::min<size_t>(frames, (size_t)(b + 1) * gpb * gop);
        IE_CUDA(cudaMemcpyAsync(s->d_in + f0 * fsz, yuv + f0 * fsz, (f1 - f0) * fsz, cudaMemcpyHostToDevice, s->stream_in));
        IE_CUDA(cudaEventRecord(s->ev_in[b % ie_session::kMaxStripes], s->stream_in));
    }
    s->video_hooks = (nbatches > 0 && nbatches <= 32) ? &hooks : nullptr;
    if (!s->video_hooks && frames) IE_CUDA(cudaStreamWaitEvent(st, s->ev_in[(nbatches - 1) % ie_session::kMaxStripes], 0));
    const int rc = ie_encode_video_dev(s, s->d_in, yuv_bytes, W, H, quant, use_rle, gop, merange, huffman ? 0 : 1, s->d_out, cap16,
                                       d_total, nullptr, st);
    // the stream can go down batch by batch only if the device entry point encoded exactly the upload batches
    const bool piped = s->video_hooks != nullptr && hooks.used_batch == gpb;
    s->video_hooks = nullptr;
    if (rc != IE_OK) { cudaStreamSynchronize(s->stream_in); cudaStreamSynchronize(st); return rc; }
    size_t sent = 0;                                                               // bytes of the stream already on their way down
    if (piped && !huffman) {
        for (uint32_t b = 0; b + 1 < nbatches; b++) {
            IE_CUDA(cudaEventSynchronize(s->ev_done[b % ie_session::kMaxStripes]));
            // whole 128-bit chunks below the batch's last bit are final (the chunk that holds it is shared with the next batch)
            const size_t safe = (size_t)(s->h_pinned[16 + b % 32] / 128) * 16;
            if (safe > out_cap) break;                                             // reported below, once the size is known
            if (safe > sent) {
                IE_CUDA(cudaMemcpyAsync(out + sent, s->d_out + sent, safe - sent, cudaMemcpyDeviceToHost, s->stream_out));
                sent = safe;
            }
        }
    }
    IE_CUDA(cudaMemcpyAsync(s->h_pinned, d_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    {
        const int erc = read_err_flag(s, st);
        if (erc != IE_OK) { cudaStreamSynchronize(s->stream_out); return erc; }     // no copy into the caller's buffer outlives the call
    }
    size_t bytes = (size_t)((s->h_pinned[0] + 7) / 8);
    const uint8_t *d_result = s->d_out;
    if (huffman) {
        IE_TRY(session_reserve(&s->d_tmp, &s->d_tmp_cap, bytes + 4096 + 32));
        size_t hb = 0;
        IE_TRY(ie_huffman_encode_dev(s, s->d_out, bytes, s->d_tmp, s->d_tmp_cap, &hb, st));
        bytes = hb;
        d_result = s->d_tmp;
        sent = 0;
    }
    *out_bytes = bytes;
    if (bytes > out_cap) { cudaStreamSynchronize(s->stream_out); set_error("output buffer too small"); return IE_ENOSPC; }
    if (bytes > sent) IE_CUDA(cudaMemcpyAsync(out + sent, d_result + sent, bytes - sent, cudaMemcpyDeviceToHost, st));
    IE_CUDA(cudaStreamSynchronize(st));
    IE_CUDA(cudaStreamSynchronize(s->stream_out));
    return IE_OK;
}

}  // extern "C"

namespace ie {
// 1 = whole-stream parse + frame k of every GOP per launch (default), 0 = frame by frame (also the fallback for truncated and
// damaged streams)
std::atomic<int> g_video_decode_variant{1};
std::atomic<int> g_video_decode_batches{4};      // GOP batches of the whole-stream decode (chain of batch b+1 next to the reconstruction of batch b)
std::atomic<uint64_t> g_stat_video_whole{0}, g_stat_video_frames{0};

// Returns IE_OK with done = false when the stream is not a plain well-formed one (the caller then decodes frame by frame).
static int decode_video_whole(ie_session *s, const uint8_t *d_enc, size_t enc_bytes, const ParsedHeader &h, int motioncomp, uint8_t *d_out,
                              cudaStream_t st, bool &done) {
    done = false;
    const uint32_t W = h.W, H = h.H, frames = h.frames, gop = h.gop;
    const size_t ysz = (size_t)W * H, fsz = ysz + ysz / 2;
    const unsigned nblocks = (W / 4) * (H / 4), nmb = (W / kMB) * (H / kMB);
    const unsigned mvbits = host_bits_needed((int)(short)h.merange);
    const uint32_t ngops = (frames + gop - 1) / gop;
    if (frames == 0 || !s->h_pinned) return IE_OK;
    if ((unsigned long long)enc_bytes * 8ull >= (1ull << 41)) return IE_OK;          // group indices are 32 bits
    const VideoParseSizes z = video_parse_sizes(enc_bytes, frames, s->dev->sm_count);
    IE_TRY(session_reserve(&s->d_parse, &s->parse_cap, z.bytes));
    // block offsets of one frame slot (frame k of every GOP), then the stream's size once per GOP and the first frame's bit
    const uint32_t nbatches = std::min<uint32_t>((uint32_t)g_video_decode_batches.load(), ngops), gpb = (ngops + nbatches - 1) / nbatches;      // GOP batches (below)
    const size_t batch_frames = (size_t)gpb * gop;
    const size_t need_off = (batch_frames * (nblocks + 1) + ngops + 2) * sizeof(unsigned long long) + 64;
    if (s->block_off_cap < need_off) {
        if (s->d_block_off) IE_CUDA(cudaFree(s->d_block_off));
        s->d_block_off = nullptr; s->block_off_cap = 0;
        IE_CUDA(cudaMalloc(&s->d_block_off, need_off));
        s->block_off_cap = need_off;
    }
    unsigned long long *d_totals = s->d_block_off + batch_frames * (nblocks + 1);
    std::vector<unsigned long long> consts(ngops + 1, (unsigned long long)enc_bytes * 8ull);
    consts[ngops] = (unsigned long long)h.end_bit;
    IE_CUDA(cudaMemcpyAsync(d_totals, consts.data(), consts.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
    VideoScratch vs;
    IE_TRY(video_scratch(s, nmb, ngops, vs));                 // vs.copy: clamped prediction coordinates of a frame slot
    VideoParse v;
    ParseParamsOpaque po;
    IE_TRY(launch_video_parse(d_enc, d_totals, d_totals + ngops, h.use_rle, nblocks, nmb * 2 * mvbits, frames, gop, z, s->d_parse, v, po, st));

    DecodeParams p;
    memset(&p, 0, sizeof p);
    p.enc = d_enc; p.enc_stride = 0; p.enc_bits = d_totals; p.start_bit = d_totals; p.block_off = s->d_block_off; p.nblocks = nblocks;
    p.bx = W / 4; p.N = 4; p.use_rle = h.use_rle; make_quant(p.quant, h.quant, 4); make_k2(p.k2, h.quant, 4); p.tab = s->dev->d_t4; p.pitch = W; p.err = s->d_err;
    p.out_stride = (size_t)gop * fsz;
    // The frame chain is sequential (one CTA, ~9 us per frame); the reconstruction is not.  GOPs go in up to four batches (tools/ab_video_batches.py: 1 / 2 / 3 / 4 / 5 / 8 / 20 batches = 4.58 / 3.92 / 3.84 / 3.84 / 3.92 / 4.16 / 5.74 ms): the
    // chain of batch b runs on the caller's stream, the reconstruction of batch b on a second stream as soon as the chain has
    // passed it, i.e. next to the chain of batch b + 1.  Everything is enqueued before the one synchronisation below; if the
    // chain gives up (truncated / damaged stream) the batches already enqueued decode empty records (harmless) and the caller
    // decodes frame by frame over their output.
    IE_TRY(session_ensure_pipeline(s));
    VideoDecodeHooks *dhooks = static_cast<VideoDecodeHooks *>(s->video_dec_hooks);
    cudaStream_t st2 = s->stream_aux;
    cudaEvent_t ev_join = nullptr;
    IE_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    int rc = IE_OK;
    for (uint32_t g0 = 0; g0 < ngops && rc == IE_OK; g0 += gpb) {
        const uint32_t g1 = std::min(ngops, g0 + gpb);
        const uint32_t f0 = g0 * gop, f1 = std::min(frames, g1 * gop);
        rc = launch_video_chain(v, po, f0, f1, st);
        cudaEvent_t ev = s->ev_in[(g0 / gpb) % ie_session::kMaxStripes];
        if (rc == IE_OK && cudaEventRecord(ev, st) != cudaSuccess) rc = IE_ECUDA;
        if (rc == IE_OK && cudaStreamWaitEvent(st2, ev, 0) != cudaSuccess) rc = IE_ECUDA;
        // the offsets of every frame of the batch in one launch (frame f0 + i -> array i); a frame slot then reads every gop-th
        // (without motion compensation only the I-frames are decoded: array i = the first frame of the batch's GOP i)
        if (rc == IE_OK) rc = motioncomp ? launch_video_emit(v, po, f0, 1, f1 - f0, s->d_block_off, st2)
                                         : launch_video_emit(v, po, f0, gop, (f1 - f0 + gop - 1) / gop, s->d_block_off, st2);
        p.block_off_stride = motioncomp ? (size_t)gop * (nblocks + 1) : 0;
        for (uint32_t k = 0; k < gop && f0 + k < f1 && rc == IE_OK; k++) {
            const uint32_t nimg = (f1 - f0 - k + gop - 1) / gop;                  // GOPs of the batch that have a frame k
            p.out = d_out + (size_t)(f0 + k) * fsz;
            if (k > 0) {
                MCBatchParams mc;
                mc.enc = d_enc; mc.enc_bits = consts[0]; mc.rec = v.rec; mc.first_frame = f0 + k; mc.gop = gop; mc.mvbits = mvbits; mc.out = d_out; mc.fsz = fsz;
                mc.W = (int)W; mc.H = (int)H; mc.mx = (int)(W / kMB); mc.nmb = (int)nmb;
                if (!motioncomp) {                                                // Frame.cpp:107-117: the motion-compensated copy only
                    if (launch_pdl(mc_copy_batch_kernel, dim3((nmb + 7) / 8, nimg), dim3(256), 0, st2, mc) != cudaSuccess) { rc = IE_ECUDA; break; }
                    count_launch();
                    continue;
                }
                // with residuals the copy is folded into the block decoder: only the clamped coordinates are prepared here
                if (launch_pdl(mc_coord_batch_kernel, dim3((nmb + 255) / 256, nimg), dim3(256), 0, st2, mc, vs.copy) != cudaSuccess) { rc = IE_ECUDA; break; }
                count_launch();
            }
            p.add_mode = k > 0 ? 1 : 0;
            p.mc_coord = k > 0 ? vs.copy : nullptr; p.ref_delta = fsz; p.mbx = W / kMB; p.mc_stride = nmb;
            p.block_off = s->d_block_off + (size_t)k * (nblocks + 1);
            rc = launch_decode_blocks(p, nimg, st2);
        }
        if (dhooks && rc == IE_OK) {
            fill_uv_kernel<<<256, 256, 0, st2>>>(d_out + (size_t)f0 * fsz, ysz, fsz, f1 - f0);       // Frame.cpp:122-124
            count_launch();
            rc = dhooks->after_batch(dhooks->ctx, g0 / gpb, (size_t)f0 * fsz, (size_t)f1 * fsz, st2);
        }
    }
    // join: the caller's stream continues after the second stream's last kernel
    if (cudaEventRecord(ev_join, st2) != cudaSuccess || cudaStreamWaitEvent(st, ev_join, 0) != cudaSuccess) rc = (rc == IE_OK) ? IE_ECUDA : rc;
    cudaError_t ce = cudaMemcpyAsync(&s->h_pinned[8], v.result, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);                // consts is read by the copy at the top until here
    cudaEventDestroy(ev_join);
    if (rc != IE_OK) { if (rc == IE_ECUDA) set_error("CUDA error in the whole-stream video decode"); return rc; }
    IE_CUDA(ce);
    if ((unsigned)(s->h_pinned[8] & 0xffffffffu) != 1u) {
        // a device-side error flag raised by the batches that decoded empty records must not outlive this attempt
        IE_CUDA(cudaMemsetAsync(s->d_err, 0, sizeof(int), st));
        return IE_OK;
    }
    if (!dhooks) {
        fill_uv_kernel<<<256, 256, 0, st>>>(d_out, ysz, fsz, frames);               // Frame.cpp:122-124
        count_launch();
    }
    IE_CUDA(cudaGetLastError());
    done = true;
    return IE_OK;
}
}  // namespace ie

extern "C" {

int ie_decode_video_dev(ie_session *s, const uint8_t *d_enc, size_t enc_bytes, uint64_t start_bit, int motioncomp, uint8_t *d_out,
                        size_t out_cap, uint32_t *Wo, uint32_t *Ho, uint32_t *Fo, void *stream) {
    if (!s || !d_enc || !d_out) { set_error("NULL argument"); return IE_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t hb[192];
    const size_t first = (size_t)(start_bit / 8);
    if (first >= enc_bytes) { set_error("start_bit beyond the stream"); return IE_EFORMAT; }
    const size_t n = std::min(sizeof hb, enc_bytes - first);
    IE_CUDA(cudaMemcpyAsync(hb, d_enc + first, n, cudaMemcpyDeviceToHost, st));
    IE_CUDA(cudaStreamSynchronize(st));
    ParsedHeader h;
    parse_header(hb, n, (size_t)(start_bit % 8), 4, h, 1);
    h.end_bit += first * 8;
    if (Wo) *Wo = h.W;
    if (Ho) *Ho = h.H;
    if (Fo) *Fo = h.frames;
    IE_TRY(check_video_dims(h.W, h.H));
    if (h.gop < 1) { set_error("gop 0 in the stream header"); return IE_EFORMAT; }
    const uint32_t W = h.W, H = h.H, frames = h.frames;
    const size_t ysz = (size_t)W * H, fsz = ysz + ysz / 2;
    if (fsz * frames > out_cap) { set_error("decoded video does not fit the output buffer"); return IE_ENOSPC; }
    const unsigned nblocks = (W / 4) * (H / 4), nmb = (W / kMB) * (H / kMB);
    const unsigned mvbits = host_bits_needed((int)(short)h.merange);
    IE_TRY(session_ensure_err(s));
    if (g_video_decode_variant.load() == 1) {
        bool done = false;
        IE_TRY(decode_video_whole(s, d_enc, enc_bytes, h, motioncomp, d_out, st, done));
        if (done) {
            g_stat_video_whole.fetch_add(1);
            if (s->video_dec_hooks) static_cast<VideoDecodeHooks *>(s->video_dec_hooks)->whole_done = true;
            return IE_OK;
        }
    }
    g_stat_video_frames.fetch_add(1);
    VideoScratch vs;
    IE_TRY(video_scratch(s, nmb, 1, vs));
    const size_t need_off = ((size_t)nblocks + 1) * sizeof(unsigned long long) + 64;
    if (s->block_off_cap < need_off) {
        if (s->d_block_off) IE_CUDA(cudaFree(s->d_block_off));
        IE_CUDA(cudaMalloc(&s->d_block_off, need_off));
        s->block_off_cap = need_off;
    }
    unsigned long long consts[2] = {(unsigned long long)enc_bytes * 8ull, (unsigned long long)h.end_bit};
    unsigned long long *d_consts = s->d_block_off + (nblocks + 1);
    IE_CUDA(cudaMemcpyAsync(d_consts, consts, sizeof consts, cudaMemcpyHostToDevice, st));
    IE_CUDA(cudaMemcpyAsync(vs.cursor, &consts[1], sizeof(unsigned long long), cudaMemcpyHostToDevice, st));

    DecodeParams p;
    memset(&p, 0, sizeof p);
    p.enc = d_enc; p.enc_bits = d_consts; p.start_bit = d_consts + 1; p.block_off = s->d_block_off; p.nblocks = nblocks;
    p.bx = W / 4; p.N = 4; p.use_rle = h.use_rle; make_quant(p.quant, h.quant, 4); make_k2(p.k2, h.quant, 4); p.tab = s->dev->d_t4; p.pitch = W; p.err = s->d_err;
    p.cursor = vs.cursor;
    // The chain of a frame is followed over a span of the stream that covers the largest possible frame (every block at its
    // maximum size): the frame's end is not known in advance, and a read-back per frame to size the span would serialise the
    // host with the device.  What the walkers find behind the frame's last block is ignored (parse.cu); the cursor stays on
    // the device.
    const size_t span = (size_t)nblocks * (4 + 16 + 16 * 16) + ((size_t)1 << 16);
    IE_TRY(session_reserve(&s->d_parse, &s->parse_cap, parse_scratch_bytes(std::max(enc_bytes, span / 8 + 1), 4)));
    for (uint32_t f = 0; f < frames; f++) {
        uint8_t *cur = d_out + (size_t)f * fsz;
        const bool is_i = (f % h.gop) == 0;
        p.out = cur;
        if (is_i) {
            p.skip_bits = 0; p.add_mode = 0;
        } else {
            MCParams mc;
            mc.enc = d_enc; mc.enc_bits = consts[0]; mc.cursor = vs.cursor; mc.mvbits = mvbits; mc.ref = d_out + (size_t)(f - 1) * fsz;
            mc.cur = cur; mc.W = (int)W; mc.H = (int)H; mc.mx = (int)(W / kMB); mc.nmb = (int)nmb;
            IE_CUDA(launch_pdl(mc_copy_kernel, dim3((nmb + 7) / 8), dim3(256), 0, st, mc));
            count_launch();
            IE_CUDA(cudaGetLastError());
            p.skip_bits = nmb * 2 * mvbits; p.add_mode = 1;
        }
        IE_TRY(launch_parallel_parse(p, span, s->d_parse, st));              // advances the cursor past this frame
        if (is_i || motioncomp) IE_TRY(launch_decode_blocks(p, 1, st));      // Frame.cpp:107-117
    }
    if (frames) {
        fill_uv_kernel<<<256, 256, 0, st>>>(d_out, ysz, fsz, frames);         // Frame.cpp:122-124 (W*H/2 is a multiple of 16)
        count_launch();
        IE_CUDA(cudaGetLastError());
    }
    return IE_OK;
}

int ie_decode_video(const uint8_t *enc, size_t enc_bytes, int motioncomp, uint8_t *yuv_out, size_t yuv_cap, size_t *yuv_bytes,
                    uint32_t *Wo, uint32_t *Ho, uint32_t *Fo) {
    if (!enc || !yuv_out || enc_bytes == 0) { set_error("NULL/empty argument"); return IE_EINVAL; }
    SessionLease lease;
    IE_TRY(lease.acquire(3, 0, 0, 4, 0));
    ie_session *s = lease.get();
    cudaStream_t st = s->stream;
    IE_TRY(session_reserve(&s->d_in, &s->d_in_cap, enc_bytes + 16));
    IE_CUDA(cudaMemcpyAsync(s->d_in, enc, enc_bytes, cudaMemcpyHostToDevice, st));
    const uint8_t *d_plain = s->d_in;
    size_t plain_bytes = enc_bytes;
    uint64_t start_bit = 1;
    if (enc[0] & 0x80) {
        IE_TRY(session_reserve(&s->d_tmp, &s->d_tmp_cap, enc_bytes * 8 + 64));
        IE_TRY(ie_huffman_decode_dev(s, s->d_in, enc_bytes, s->d_tmp, s->d_tmp_cap, &plain_bytes, &start_bit, st));
        d_plain = s->d_tmp;
    }
    uint8_t hb[192];
    const size_t n = std::min(sizeof hb, plain_bytes);
    IE_CUDA(cudaMemcpyAsync(hb, d_plain, n, cudaMemcpyDeviceToHost, st));
    IE_CUDA(cudaStreamSynchronize(st));
    ParsedHeader h;
    parse_header(hb, n, (size_t)start_bit, 4, h, 1);
    if (Wo) *Wo = h.W;
    if (Ho) *Ho = h.H;
    if (Fo) *Fo = h.frames;
    IE_TRY(check_video_dims(h.W, h.H));
    const size_t total = (size_t)h.W * h.H * 3 / 2 * h.frames;
    if (yuv_bytes) *yuv_bytes = total;
    if (total > yuv_cap) { set_error("yuv_out too small"); return IE_ENOSPC; }
    IE_TRY(session_reserve(&s->d_out, &s->d_out_cap, std::max<size_t>(total, 16)));
    uint32_t w, hh, ff;
    // Copy / compute pipeline: the frames of a reconstructed GOP batch go down on stream_out while later batches are still
    // being decoded.  If the whole-stream decode gives up (damaged stream: frame-by-frame fallback over the same buffer), the
    // early copies are discarded and everything goes down again at the end.
    IE_TRY(session_ensure_pipeline(s));
    struct Ctx { ie_session *s; uint8_t *host; size_t sent; bool contiguous; } ctx{s, yuv_out, 0, true};
    VideoDecodeHooks hooks;
    hooks.ctx = &ctx;
    hooks.whole_done = false;
    hooks.after_batch = [](void *c, uint32_t index, size_t b0, size_t b1, cudaStream_t stx) -> int {
        Ctx *x = static_cast<Ctx *>(c);
        if (!x->contiguous || b0 != x->sent || index >= (uint32_t)ie_session::kMaxStripes) { x->contiguous = false; return IE_OK; }
        IE_CUDA(cudaEventRecord(x->s->ev_done[index], stx));
        IE_CUDA(cudaStreamWaitEvent(x->s->stream_out, x->s->ev_done[index], 0));
        IE_CUDA(cudaMemcpyAsync(x->host + b0, x->s->d_out + b0, b1 - b0, cudaMemcpyDeviceToHost, x->s->stream_out));
        x->sent = b1;
        return IE_OK;
    };
    // early copies only into page-locked memory: a device-to-host copy into pageable memory blocks the calling thread until it has
    // completed, which would stall the enqueueing of the later batches behind the reconstruction of the earlier ones
    cudaPointerAttributes pa;
    const bool pinned = cudaPointerGetAttributes(&pa, yuv_out) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    cudaGetLastError();
    s->video_dec_hooks = pinned ? &hooks : nullptr;
    const int rc = ie_decode_video_dev(s, d_plain, plain_bytes, start_bit, motioncomp, s->d_out, s->d_out_cap, &w, &hh, &ff, st);
    s->video_dec_hooks = nullptr;
    const bool early = rc == IE_OK && hooks.whole_done && ctx.contiguous && ctx.sent == total;
    if (rc != IE_OK) { cudaStreamSynchronize(s->stream_out); cudaStreamSynchronize(st); return rc; }
    const int erc = read_err_flag(s, st);
    if (erc != IE_OK) { cudaStreamSynchronize(s->stream_out); return erc; }
    IE_CUDA(cudaStreamSynchronize(s->stream_out));
    if (!early) {
        IE_CUDA(cudaMemcpyAsync(yuv_out, s->d_out, total, cudaMemcpyDeviceToHost, st));
        IE_CUDA(cudaStreamSynchronize(st));
    }
    return IE_OK;
}

}  // extern "C"
