// Image-block encode: one CTA per tile of consecutive blocks (raster order, ImageBase.cpp:187-199):
//   encode_tiles_kernel:  pixels -> (-128) -> forward DCT -> quant divide/round -> zigzag -> RLE info   (lane per block)
//                         -> CTA scan of block bit counts -> block-centric pack into a shared-memory image of the tile's
//                         bits -> the tile's scratch slot + its bit count                      (no inter-CTA dependency)
//   tile_copyout_kernel:  per group of tiles: sum of the earlier tiles' bit counts, then the tile images re-aligned into
//                         the stream                                                          (thread per 128-bit chunk)
// HBM traffic: W*H bytes in, the stream out, plus one L2-resident round trip of the stream through the tile scratch.
// Replaces ImageEncoder.cpp:121-138 (parallel DCT loop + the strictly serial streamEncoded loop), Frame.cpp:141-158
// (I-frames) and, with PF, Frame.cpp:160-244 (P-frame blocks).
//
// Template switches: N block size, BPL blocks per lane, PF P-frame mode (residual in, reconstruction out),
// FAST = FP32 factorised transform + guard band + exact fallback (transform_fast.cuh); FAST=false evaluates every
// coefficient in the reference's exact order (kept for cross-checking: ie_set_option("exact_transform", 1)).
#include "encode_image.cuh"
#include "transform_fast.cuh"
#include "exact.cuh"

namespace ie {

// inverse zigzag (raster index -> zigzag position), algo.cpp:68-87, for compile-time use in the unrolled fast path
__device__ constexpr unsigned char kZigzagInv4[16] = {0, 1, 5, 6, 2, 4, 7, 12, 3, 8, 11, 13, 9, 10, 14, 15};
__device__ constexpr unsigned char kZigzagInv8[64] = {
    0,  1,  5,  6,  14, 15, 27, 28, 2,  4,  7,  13, 16, 26, 29, 42, 3,  8,  12, 17, 25, 30, 41, 43, 9,  11, 18, 24, 31, 40, 44, 53,
    10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38, 46, 51, 55, 60, 21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63};

// RLE info from the staged zigzag coefficients (slow generic path; used after a guard-band patch and by FAST=false)
template <int NN>
__device__ __forceinline__ void block_stats_from_staging(const short *cf, int &lastnz, int &prevnz, unsigned &orbits) {
    lastnz = 0; prevnz = 0; orbits = 0;
#pragma unroll 4
    for (int k = 0; k < NN; k++) {
        const int q = cf[k];
        if (q != 0) {
            orbits |= (unsigned)(q ^ (q >> 31));
            prevnz = (k < NN - 1) ? (k + 1) : prevnz;
            lastnz = k + 1;
        }
    }
}

// The pack loop of encode_tiles_kernel's phase 3 (same statements) writing to any address space: used by the reduced-staging
// instantiations (encode_variant 3 / 4) when a tile image is larger than their shared-memory staging area and is packed straight
// into the tile's slot of the global scratch buffer instead.  Every lane writes its blocks' fields MSB-first into the image of
// the tile's bits at `outw` (tile-local alignment, zero-filled by the caller): whole words with plain stores, the <= 2 words a
// block shares with its neighbours with atomicOr.  (The kernel keeps its own inline copy of the loop for the shared-memory
// case so that the default instantiations' SASS stays what was measured.)
template <int NN, int BPL>
__device__ __forceinline__ void pack_tile_blocks(unsigned *outw, const short *s_coef, const unsigned *s_off, const unsigned char *s_w,
                                                 const unsigned char *s_len, int nblk, int use_rle) {
    constexpr int STRIDE = NN + 2;
#pragma unroll 1
    for (int r = 0; r < BPL; r++) {
        const int lb = threadIdx.x * BPL + r;
        if (lb >= nblk) break;
        const int w = s_w[lb], len = s_len[lb];
        const unsigned pos = s_off[lb];
        const unsigned bo = pos & 31u;
        unsigned *ow = outw + (pos >> 5);
        const unsigned *cw = reinterpret_cast<const unsigned *>(s_coef + (size_t)lb * STRIDE);
        const unsigned mask = (1u << w) - 1u;                                             // w <= 16
        // header: bit_len (low 4 bits survive, Block.cpp:381) and, with RLE, the length field (Block.cpp:393)
        unsigned long long acc = use_rle ? ((((unsigned long long)w & 15ull) << w) | (unsigned long long)len) : ((unsigned long long)w & 15ull);
        int nacc = (int)bo + 4 + (use_rle ? w : 0);
        const int nfull = len >> 1;
        const int s2 = 2 * w;
        int j = 0;
        bool odd_left = (len & 1) != 0;
        // stage A: up to the first completed word -- the only one that can hold bits of the previous block(s)
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
            // stage B: whole words that belong to this block alone
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
}

// four bytes at any alignment from the two aligned words that hold them (the second is not touched when the address is aligned)
__device__ __forceinline__ unsigned load_u8x4_unaligned(const uint8_t *p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const unsigned *wp = reinterpret_cast<const unsigned *>(a & ~(uintptr_t)3);
    const unsigned k = (unsigned)(a & 3);
    const unsigned w0 = __ldg(wp), w1 = k ? __ldg(wp + 1) : 0u;
    return __byte_perm(w0, w1, 0x3210u + 0x1111u * k);
}

constexpr int kQueueCap = 128;        // guard-band fallback entries per tile handled by the CTA-wide queue

// VAR: 0 scalar transform, 1 lean quantise, 2 packed transform + lean quantise (default); experimental, not yet run on a B200:
// 3 / 4 = 2 with a reduced staging area for the tile image (kSmallOutChunks chunks instead of the worst case; larger tile
// images are packed in global memory) and 7 / 8 CTAs per SM instead of 6; 5 = 2 with the short-chain binary64 evaluation in
// front of the exact queue ("fast64", transform_fast.cuh); 6 / 7 = 3 / 4 with fast64.
constexpr int kSmallOutChunks = 512;
constexpr bool var_small_out(int VAR) { return VAR == 3 || VAR == 4 || VAR == 6 || VAR == 7; }
constexpr bool var_fast64(int VAR) { return VAR >= 5; }
// 9 = 2 with the guard-band coefficients evaluated by the lanes that own them, all lanes of a warp that have one at the same
// time (short binary64 chain, then the exact chain for the ties): no CTA-wide queue, no barrier behind it
constexpr bool var_inwarp(int VAR) { return VAR == 9; }
constexpr int encode_min_ctas(int N, bool PF, bool FAST, int VAR) {
    return (N == 8 && !FAST) ? 1 : (FAST && !PF ? ((VAR == 3 || VAR == 6) ? 7 : (VAR == 4 || VAR == 7) ? 8 : 6) : 2);
}

template <int N, int BPL, bool PF, bool FAST, int VAR = 0>
__global__ void __launch_bounds__(kThreads, encode_min_ctas(N, PF, FAST, VAR)) encode_tiles_kernel(const EncodeParams p) {
    constexpr int NN = N * N;
    constexpr int TB = kThreads * BPL;            // blocks per tile
    constexpr int STRIDE = NN + 2;                // halfwords per block in the staging area: NN/2 + 1 words (odd -> bank spread)
    constexpr int NSEG = NN / 8;
    constexpr int MAXCHUNKS = (TB * (4 + 16 + 16 * NN) + 127) / 128 + 2;
    constexpr bool SMALL_OUT = var_small_out(VAR);
    constexpr bool FAST64 = var_fast64(VAR);
    constexpr bool INWARP = var_inwarp(VAR) && FAST && !PF;
    constexpr int OUTCHUNKS = SMALL_OUT ? kSmallOutChunks : MAXCHUNKS;     // staging area for the tile image, in 128-bit chunks
    extern __shared__ __align__(16) unsigned char smem[];
    short *s_coef = reinterpret_cast<short *>(smem);
    unsigned *s_off = reinterpret_cast<unsigned *>(smem + (size_t)TB * STRIDE * sizeof(short));
    uint4 *s_out = reinterpret_cast<uint4 *>(smem + (((size_t)TB * STRIDE * sizeof(short) + (TB + 1) * sizeof(unsigned) + 15) & ~(size_t)15));
    unsigned *s_queue = reinterpret_cast<unsigned *>(s_out + OUTCHUNKS);
    unsigned char *s_w = reinterpret_cast<unsigned char *>(s_queue + kQueueCap + TB + (kQueueCap + TB) / 2);
    unsigned char *s_len = s_w + TB;
    unsigned char *s_dirty = s_len + TB;
    unsigned *s_stats = reinterpret_cast<unsigned *>(s_queue + kQueueCap);          // [TB] packed RLE info of patched blocks
    unsigned short *s_dlist = reinterpret_cast<unsigned short *>(s_stats + TB);      // [kQueueCap + TB] patched blocks (duplicates allowed)
    __shared__ unsigned s_warp[kThreads / 32 + 1];
    __shared__ unsigned s_qn, s_nd;

    const unsigned img = blockIdx.y;
    ScanState st = p.scan;
    st.tile_state += (size_t)img * p.tiles_per_image;
    st.bnd += (size_t)img * p.tiles_per_image;
    st.ticket += img;

    // Tile id = blockIdx.x: CTAs of a 1-D grid row are dispatched in increasing index order (the same assumption CUB's
    // single-pass scan makes), so every predecessor of a spinning tile is resident or done.  (A ticket atomic costs a
    // global round trip at the start of every tile.)
    if (threadIdx.x == 0) { s_qn = 0; s_nd = 0; }
    __syncthreads();
    const unsigned tile = blockIdx.x;
    const unsigned ntiles = p.tiles_per_image;
    const unsigned first_blk = tile * TB;
    const int nblk = min((unsigned)TB, p.nblocks - first_blk);

    const uint8_t *src = p.src + (size_t)img * p.img_stride;
    // P-frames of several GOPs in one launch: image i = the frame of GOP i (same strides for the reference frame, which is
    // the frame before it, and for the frame being rebuilt in place)
    const uint8_t *ref_i = PF ? p.ref + (size_t)img * p.img_stride : nullptr;
    uint8_t *cur_rw_i = PF ? p.cur_rw + (size_t)img * p.img_stride : nullptr;
    const short *res_coord_i = PF ? p.res_coord + (size_t)img * p.coord_stride : nullptr;
    const short *copy_coord_i = PF ? p.copy_coord + (size_t)img * p.coord_stride : nullptr;
    const BlockTables *tab = p.tab;
    ExactCtx ex;
    ex.src = src; ex.ref = ref_i; ex.res_coord = res_coord_i; ex.tab = tab; ex.pitch = p.pitch; ex.bx = p.bx; ex.mbx = p.mbx;

    // ---- phase 1: transform + quantise, lane per block ---------------------------------------------------
    unsigned r_orbits[BPL];
    unsigned r_orseg[BPL][NSEG];
    unsigned long long r_near[INWARP ? BPL : 1];
#pragma unroll
    for (int r = 0; r < BPL; r++) {
        const int lb = threadIdx.x * BPL + r;
        if (INWARP) r_near[r] = 0;
        r_orbits[r] = 0;
#pragma unroll
        for (int s = 0; s < NSEG; s++) r_orseg[r][s] = 0;
        if (lb >= nblk) continue;
        s_dirty[lb] = 0;
        const unsigned gb = first_blk + lb;
        const unsigned byi = gb / p.bx, bxi = gb - byi * p.bx;
        short *cf = s_coef + (size_t)lb * STRIDE;
        int rx = 0, ry = 0;
        if (PF) {
            const unsigned mb = (byi >> 2) * p.mbx + (bxi >> 2);
            rx = res_coord_i[2 * mb] + (int)(bxi & 3) * 4;
            ry = res_coord_i[2 * mb + 1] + (int)(byi & 3) * 4;
        }
        if (FAST) {
            unsigned long long near = 0;
            if (VAR >= 2) {
            // variant 2: the whole block is loaded first, rows 2r and 2r+1 are converted and transformed as f32x2 pairs
            unsigned raw[N][N / 4];
            unsigned rraw[PF ? N : 1];          // P-frames (N = 4): the four reference pixels of each row of the residual's source block
#pragma unroll
            for (int y = 0; y < N; y++) {
                const uint8_t *row = src + (size_t)(byi * N + y) * p.pitch + (size_t)bxi * N;
                if (N == 8) {
                    const uint2 v = __ldg(reinterpret_cast<const uint2 *>(row));
                    raw[y][0] = v.x; raw[y][N / 4 - 1] = v.y;
                } else {
                    raw[y][0] = PF ? *reinterpret_cast<const unsigned *>(row) : __ldg(reinterpret_cast<const unsigned *>(row));
                }
                if (PF) rraw[PF ? y : 0] = load_u8x4_unaligned(ref_i + (size_t)(ry + y) * p.pitch + rx);
            }
            float2 x2[NN / 2], y2[NN / 2];
#pragma unroll
            for (int r2 = 0; r2 < N / 2; r2++)
#pragma unroll
                for (int k = 0; k < N; k++) {
                    // bytes -> floats by planting them in the mantissa of 2^23, then one packed subtraction of 2^23 + 128 (exact)
                    const float a = __uint_as_float(__byte_perm(raw[2 * r2][k >> 2], 0x4B000000u, 0x7650u | (unsigned)(k & 3)));
                    const float b = __uint_as_float(__byte_perm(raw[2 * r2 + 1][k >> 2], 0x4B000000u, 0x7650u | (unsigned)(k & 3)));
                    if (!PF) {
                        x2[r2 * N + k] = lean::add2(make_float2(a, b), make_float2(-8388736.0f, -8388736.0f));
                    } else {
                        // residual (Block.cpp:262) and the -128 of the second pass: (2^23 + c) - (2^23 + r) - 128, all exact
                        const float ra = __uint_as_float(__byte_perm(rraw[PF ? 2 * r2 : 0], 0x4B000000u, 0x7650u | (unsigned)(k & 3)));
                        const float rb = __uint_as_float(__byte_perm(rraw[PF ? 2 * r2 + 1 : 0], 0x4B000000u, 0x7650u | (unsigned)(k & 3)));
                        x2[r2 * N + k] = lean::add2(lean::sub2(make_float2(a, b), make_float2(ra, rb)), make_float2(-128.0f, -128.0f));
                    }
                }
            lean::fdct2d_packed<N>(x2, y2);
            unsigned nlo, nhi;
            lean::quantise_block_packed<N>(y2, p.fq, p.dc_den2, p.dc_rcp, reinterpret_cast<unsigned *>(cf), nlo, nhi, r_orseg[r], r_orbits[r]);
            near = ((unsigned long long)nhi << 32) | nlo;
            } else {
            float x[NN];
#pragma unroll
            for (int y = 0; y < N; y++) {
                const uint8_t *row = src + (size_t)(byi * N + y) * p.pitch + (size_t)bxi * N;
                unsigned raw[N / 4];
                if (N == 8) {
                    const uint2 v = __ldg(reinterpret_cast<const uint2 *>(row));
                    raw[0] = v.x; raw[N / 4 - 1] = v.y;
                } else {
                    raw[0] = PF ? *reinterpret_cast<const unsigned *>(row) : __ldg(reinterpret_cast<const unsigned *>(row));
                }
#pragma unroll
                for (int k = 0; k < N; k++) {
                    // byte -> float by planting it in the mantissa of 2^23, then subtracting 2^23 + 128 (exact)
                    float f = __uint_as_float(__byte_perm(raw[k >> 2], 0x4B000000u, 0x7650u | (unsigned)(k & 3))) - 8388736.0f;
                    if (PF) f -= (float)(int)__ldg(ref_i + (size_t)(ry + y) * p.pitch + rx + k);      // exact small integers
                    x[y * N + k] = f;
                }
            }
            fdct2d_fast<N>(x);
            if (VAR == 1) {
                // variant 1 (transform_fast.cuh, lean::): raster pairs quantised with packed f32x2 operations, zigzag pairs
                // stored as words, max bits_needed from a packed running max / min -- same values, fewer instructions
                float2 y2[NN / 2];
#pragma unroll
                for (int i = 0; i < NN / 2; i++) y2[i] = make_float2(x[2 * i], x[2 * i + 1]);
                unsigned nlo, nhi;
                lean::quantise_block_packed<N>(y2, p.fq, p.dc_den2, p.dc_rcp, reinterpret_cast<unsigned *>(cf), nlo, nhi, r_orseg[r], r_orbits[r]);
                near = ((unsigned long long)nhi << 32) | nlo;
            } else
#pragma unroll
            for (int uv = 0; uv < NN; uv++) {
                int q;
                if (uv == 0) {
                    // DC: the sum of the samples is an exact integer in FP32 and so is the reference's value
                    // (cos(0) = 1, C(0)^2 = 0.25): q = round_half_away(S / (4 Q00)) in integer arithmetic
                    const int S = (int)x[0];
                    const int n = abs(S);
                    const int D2 = p.dc_den2;                              // 2 * 4 * Q00
                    const int num = 2 * n + (D2 >> 1);
                    int qq = (int)((float)num * p.dc_rcp);                  // floor estimate, corrected below
                    const int rem = num - qq * D2;
                    qq += (rem >= D2) ? 1 : 0;
                    qq -= (rem < 0) ? 1 : 0;
                    q = (S < 0) ? -qq : qq;
                } else {
                    const float rr = fmaf(x[uv], p.fq.k[uv], kMagic);      // rn(q~) in the mantissa
                    const float rf = rr - kMagic;
                    const float d = fmaf(x[uv], p.fq.k[uv], -rf);
                    if (fabsf(d) >= p.fq.thr[uv]) near |= 1ull << uv;       // inside the guard band -> exact recompute
                    q = __float_as_int(rr) - kMagicBits;
                }
                const int k = (N == 8) ? kZigzagInv8[uv] : kZigzagInv4[uv];
                cf[k] = (short)q;
                r_orseg[r][k >> 3] |= (unsigned)q;                          // non-zero detection per zigzag segment
                r_orbits[r] |= (unsigned)(q ^ (q >> 31));                   // bits_needed of the widest value (-1 -> 0)
            }
            }
            if (INWARP) {
                r_near[INWARP ? r : 0] = near;
            } else if (near) {
                // hand the guard-band coefficients to the CTA-wide queue (filled lanes instead of one lane per warp)
                const int n = __popcll(near);
                unsigned slot = atomicAdd(&s_qn, (unsigned)n);
                while (near) {
                    const int uv = __ffsll((long long)near) - 1;
                    near &= near - 1;
                    if (slot < (unsigned)kQueueCap) {
                        s_queue[slot++] = ((unsigned)lb << 8) | (unsigned)uv;
                    } else {                                                // queue full (adversarial input): do it here
                        const int q = exact_coefficient<N, PF>(ex, gb, uv, p.quant.m[uv]);
                        const int k = tab->izz[uv];
                        if (cf[k] != (short)q) { cf[k] = (short)q; if (!s_dirty[lb]) { s_dirty[lb] = 1; s_dlist[atomicAdd(&s_nd, 1u)] = (unsigned short)lb; } }
                    }
                }
            }
        } else {
            double x[NN];
#pragma unroll
            for (int ij = 0; ij < NN; ij++) x[ij] = exact_sample<N, PF>(ex, src, byi, bxi, rx, ry, ij);
#pragma unroll 1
            for (int uv = 0; uv < NN; uv++) {
                const double e = fdct_coef_exact<NN>(tab->fw + uv * NN, x, tab->cc[uv]);
                const double qd = round_half_away(__ddiv_rn(e, p.quant.m[uv]));               // Block.cpp:152
                cf[tab->izz[uv]] = (short)__double2int_rz(qd);                                // Block.cpp:205: int16_t(double)
            }
        }
    }
    if (INWARP) {
        // ---- phase 1b (variant 9): the owning lanes evaluate their guard-band coefficients, one per lane at a time, and fold
        // the result into the RLE info they hold in registers; nothing is shared with another lane, so no barrier follows
#pragma unroll
        for (int r = 0; r < BPL; r++) {
            while (__ballot_sync(0xffffffffu, r_near[INWARP ? r : 0] != 0)) {            // warp-uniform
                unsigned long long &nr = r_near[INWARP ? r : 0];
                const bool have = nr != 0;
                const int uv = have ? (__ffsll((long long)nr) - 1) : 0;
                nr &= nr - 1;
                const int lb = threadIdx.x * BPL + r;
                const unsigned gb = first_blk + lb;
                const double m_uv = p.quant.m[uv];
                int q = 0;
                bool undecided = false;
                if (have) undecided = !fast64_coefficient<N, PF>(ex, gb, uv, m_uv, q);
                if (__ballot_sync(0xffffffffu, undecided)) {
                    if (undecided) q = exact_coefficient<N, PF>(ex, gb, uv, m_uv);
                }
                if (have) {
                    short *cf = s_coef + (size_t)lb * STRIDE;
                    const int k = tab->izz[uv];
                    const int q_old = cf[k];
                    q = (int)(short)q;
                    if (q_old != q) {
                        cf[k] = (short)q;
                        const unsigned bo = (unsigned)(q_old ^ (q_old >> 31)), bn = (unsigned)(q ^ (q >> 31));
                        if (s_dirty[lb] || (__clz(bn) > __clz(bo) && __clz(bo) == __clz(r_orbits[r]))) {
                            // the patched coefficient alone carried the widest value: rescan the block
                            int lastnz, prevnz;
                            unsigned ob;
                            block_stats_from_staging<NN>(cf, lastnz, prevnz, ob);
                            s_stats[lb] = (unsigned)lastnz | ((unsigned)prevnz << 8) | (ob << 16);
                            s_dirty[lb] = 1;
                        } else {
                            r_orbits[r] |= bn;
                            const unsigned *cw = reinterpret_cast<const unsigned *>(cf) + (k >> 3) * 4;
                            const unsigned any = cw[0] | cw[1] | cw[2] | cw[3];
#pragma unroll
                            for (int s2 = 0; s2 < NSEG; s2++) if (s2 == (k >> 3)) r_orseg[r][s2] = any;
                        }
                    }
                }
            }
        }
    } else {
        __syncthreads();
    }

    // ---- phase 1b: exact recomputation of the queued guard-band coefficients, one per thread -------------------
    if (FAST && !INWARP) {
        const unsigned qn = min(s_qn, (unsigned)kQueueCap);
        if (qn) {                                                              // uniform
            for (unsigned e = threadIdx.x; e < qn; e += kThreads) {
                const unsigned ent = s_queue[e];
                const int lb = (int)(ent >> 8), uv = (int)(ent & 0xff);
                int q;
                if (FAST64) {
                    if (!fast64_coefficient<N, PF>(ex, first_blk + lb, uv, p.quant.m[uv], q))
                        q = exact_coefficient<N, PF>(ex, first_blk + lb, uv, p.quant.m[uv]);
                } else {
                    q = exact_coefficient<N, PF>(ex, first_blk + lb, uv, p.quant.m[uv]);
                }
                short *cf = s_coef + (size_t)lb * STRIDE;
                const int k = tab->izz[uv];
                if (cf[k] != (short)q) {
                    cf[k] = (short)q;
                    // first patch of this block: put it on the list of blocks whose RLE info must be recomputed.
                    // (two threads patching the same block may both list it; recomputing twice is harmless)
                    if (!s_dirty[lb]) { s_dirty[lb] = 1; s_dlist[atomicAdd(&s_nd, 1u)] = (unsigned short)lb; }
                }
            }
            __syncthreads();
            const unsigned nd = s_nd;
            // RLE info of the patched blocks, NN/8 lanes per block (8 coefficients each), combined with shuffles
            constexpr int LPB = NN / 8;
            for (unsigned e0 = 0; e0 < nd; e0 += kThreads / LPB) {              // uniform
                const unsigned e = e0 + threadIdx.x / LPB;
                const int part = (int)(threadIdx.x % LPB);
                const bool act = e < nd;
                const int lb = act ? (int)s_dlist[e] : 0;
                const unsigned *cw = reinterpret_cast<const unsigned *>(s_coef + (size_t)lb * STRIDE + part * 8);
                unsigned m8 = 0, ob = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const unsigned w2 = cw[j];
                    const int q0 = (int)(short)(w2 & 0xffffu), q1 = (int)w2 >> 16;
                    if (q0 != 0) { m8 |= 1u << (2 * j); ob |= (unsigned)(q0 ^ (q0 >> 31)); }
                    if (q1 != 0) { m8 |= 1u << (2 * j + 1); ob |= (unsigned)(q1 ^ (q1 >> 31)); }
                }
                unsigned long long mask = (unsigned long long)m8 << (8 * part);
#pragma unroll
                for (int d = 1; d < LPB; d <<= 1) {
                    mask |= __shfl_xor_sync(0xffffffffu, mask, d);
                    ob |= __shfl_xor_sync(0xffffffffu, ob, d);
                }
                if (act && part == 0) {
                    const unsigned long long m2 = mask & ~(1ull << (NN - 1));
                    const unsigned lastnz = mask ? 64u - (unsigned)__clzll((long long)mask) : 0u;
                    const unsigned prevnz = m2 ? 64u - (unsigned)__clzll((long long)m2) : 0u;
                    s_stats[lb] = lastnz | (prevnz << 8) | (ob << 16);
                }
            }
            __syncthreads();
        }
    }

    // ---- phase 1c: RLE info, P-frame reconstruction, header unit + coefficient pairs ----------------------------
    unsigned my_bits = 0;
#pragma unroll 1
    for (int r = 0; r < BPL; r++) {
        const int lb = threadIdx.x * BPL + r;
        if (lb >= nblk) { if (lb < TB) s_off[lb] = 0; continue; }
        short *cf = s_coef + (size_t)lb * STRIDE;
        int lastnz = 0, prevnz = 0;      // (zigzag index + 1) of the last / of the last non-final-position non-zero
        unsigned orbits = 0;
        if (!FAST) {
            block_stats_from_staging<NN>(cf, lastnz, prevnz, orbits);
        } else if (s_dirty[lb]) {
            const unsigned sp = s_stats[lb];
            lastnz = (int)(sp & 0xff); prevnz = (int)((sp >> 8) & 0xff); orbits = sp >> 16;
        } else {
            orbits = r_orbits[r];
            int lastseg = -1;
#pragma unroll
            for (int s = 0; s < NSEG; s++) if (r_orseg[r][s]) lastseg = s;
            if (lastseg >= 0) {
                // last non-zero inside the last non-empty 8-coefficient zigzag segment
#pragma unroll
                for (int j = 0; j < 8; j++) if (cf[lastseg * 8 + j] != 0) lastnz = lastseg * 8 + j + 1;
                if (lastnz == NN) {          // rare: the RLE quirk needs the previous non-zero as well
                    for (int k = 0; k < NN - 1; k++) if (cf[k] != 0) prevnz = k + 1;
                }
            }
        }
        // Block.cpp:214-219, 231: data_bits = max(max bits_needed(nz), ffs(data)), data = last non-zero index + 1
        int w = lastnz ? (33 - __clz(orbits)) : 0;
        w = max(w, dev_ffs((unsigned)lastnz));
        int len = lastnz;
        if (p.use_rle) {
            // Block.cpp:388-390: a full-length block whose last entry has leading zeroes loses that entry
            if (lastnz == NN && prevnz != NN - 1) len = prevnz;
        } else {
            len = NN;                                                                     // Block.cpp:396
        }
        if (PF) {
            // ImageBase.cpp:303 + Frame.cpp:218-242 + Block.cpp:110-119: decode what was just encoded and rebuild the
            // frame in place: cur = (u8)clamp(double(ref[copy block]) + (IDCT(coef*Q) + 128))
            const unsigned gb = first_blk + lb;
            const unsigned byi = gb / p.bx, bxi = gb - byi * p.bx;
            const unsigned mb = (byi >> 2) * p.mbx + (bxi >> 2);
            const int kx = copy_coord_i[2 * mb] + (int)(bxi & 3) * 4, ky = copy_coord_i[2 * mb + 1] + (int)(byi & 3) * 4;
            if (FAST && N == 4 && VAR >= 2) {
                // the same in packed f32x2 operations (lean::idct2d_packed, transform_fast.cuh); the prediction is added as
                // x + (ref + 128) -- one rounding instead of two, inside the same bound; a pixel within the bound of an integer
                // boundary takes the exact path below, so the result does not depend on how the fast value was rounded
                float2 X2[NN / 2], P2[NN / 2];
                float S = 0.f;
                unsigned nzmask = 0;
#pragma unroll
                for (int r2 = 0; r2 < N / 2; r2++)
#pragma unroll
                    for (int v = 0; v < N; v++) {
                        const int ua = (2 * r2) * N + v, ub = (2 * r2 + 1) * N + v;
                        const int ca = cf[kZigzagInv4[ua]], cb = cf[kZigzagInv4[ub]];
                        if (ca != 0) nzmask |= 1u << ua;
                        if (cb != 0) nzmask |= 1u << ub;
                        const float da = (float)ca * p.k2[ua], db = (float)cb * p.k2[ub];
                        X2[r2 * N + v] = make_float2(da, db);
                        S += fabsf(da) + fabsf(db);
                    }
                lean::idct2d_packed<4>(X2, P2);
                const float delta = (18.f * S + 2.f * (S + 383.f)) * 5.9604645e-8f * 1.0001f + 2e-6f;      // decode_image.cu: <= 14 u S
                const float hi_thr = (delta < 0.49f) ? 0.5f - delta : 0.f;
                unsigned outw[4];
                unsigned rpxw[4];
                unsigned unsure = 0;
#pragma unroll
                for (int y = 0; y < 4; y++) {
                    rpxw[y] = load_u8x4_unaligned(ref_i + (size_t)(ky + y) * p.pitch + kx);
                    unsigned fl[4];
#pragma unroll
                    for (int c = 0; c < 2; c++) {
                        const float ra = __uint_as_float(__byte_perm(rpxw[y], 0x4B000000u, 0x7650u | (unsigned)(2 * c)));
                        const float rb = __uint_as_float(__byte_perm(rpxw[y], 0x4B000000u, 0x7650u | (unsigned)(2 * c + 1)));
                        const float2 addend = lean::add2(make_float2(ra, rb), make_float2(-8388480.0f, -8388480.0f));      // ref + 128, exact
                        unsigned un = 0;
                        lean::pixel_pair_add<1u, 2u>(P2[y * 2 + c], addend, hi_thr, fl[2 * c], fl[2 * c + 1], un);
                        unsure |= un << (y * 4 + 2 * c);
                    }
                    outw[y] = __byte_perm(__byte_perm(fl[0], fl[1], 0x0040), __byte_perm(fl[2], fl[3], 0x0040), 0x5410);
                }
                while (unsure) {
                    const int ij = __ffs((int)unsure) - 1;
                    unsure &= unsure - 1;
                    const double acc = exact_inverse_pixel<NN, 4>(nzmask, ij, cf, tab, p.quant);
                    const int yy = ij >> 2, kk = ij & 3;
                    unsigned rrow = 0, orow = 0;
#pragma unroll
                    for (int r4 = 0; r4 < 4; r4++) if (r4 == yy) { rrow = rpxw[r4]; orow = outw[r4]; }
                    const double rpx = (double)(int)((rrow >> (8 * kk)) & 0xffu);
                    const unsigned px = clamp_trunc_u8(__dadd_rn(rpx, __dadd_rn(acc, 128.0)));           // Frame.cpp:218-242
                    orow = (orow & ~(0xffu << (8 * kk))) | (px << (8 * kk));
#pragma unroll
                    for (int r4 = 0; r4 < 4; r4++) if (r4 == yy) outw[r4] = orow;
                }
#pragma unroll
                for (int y = 0; y < 4; y++)
                    *reinterpret_cast<unsigned *>(cur_rw_i + (size_t)(byi * N + y) * p.pitch + (size_t)bxi * N) = outw[y];
            } else if (FAST && N == 4) {
                // fast reconstruction as in decode_blocks_fast_kernel<4, ADD>: FP32 inverse transform, and every pixel whose
                // value lies within the block's error bound of an integer boundary is recomputed in the reference's exact
                // order over the non-zero coefficients (transform bound: decode_image.cu)
                float xf[NN];
                float S = 0.f;
                unsigned nzmask = 0;
#pragma unroll
                for (int uv = 0; uv < NN; uv++) {
                    const int c = cf[kZigzagInv4[uv]];
                    if (c != 0) nzmask |= 1u << uv;
                    const float d = (float)c * p.k2[uv];
                    xf[uv] = d;
                    S += fabsf(d);
                }
                idct2d_fast<4>(xf);
                const float delta = (18.f * S + 2.f * (S + 383.f)) * 5.9604645e-8f * 1.0001f + 2e-6f;      // decode_image.cu: <= 14 u S
                const float hi_thr = (delta < 0.49f) ? 0.5f - delta : 0.f;
                unsigned outw[4];
                unsigned rpxw[4];                                     // the four reference pixels of each row, one per byte
                unsigned unsure = 0;
#pragma unroll
                for (int y = 0; y < 4; y++) {
                    unsigned fl[4];
                    rpxw[y] = 0;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const unsigned rp = __ldg(ref_i + (size_t)(ky + y) * p.pitch + kx + k);
                        rpxw[y] |= rp << (8 * k);
                        const float v = xf[y * 4 + k] + 128.f + (float)rp;
                        const float u = fminf(fmaxf(v, 0.5f), 255.5f);
                        const float fm = __fadd_rd(u, 8388608.0f);
                        const float frac = u - (fm - 8388608.0f);
                        if (fabsf(frac - 0.5f) >= hi_thr) unsure |= 1u << (y * 4 + k);
                        fl[k] = __float_as_uint(fm);
                    }
                    outw[y] = __byte_perm(__byte_perm(fl[0], fl[1], 0x0040), __byte_perm(fl[2], fl[3], 0x0040), 0x5410);
                }
                while (unsure) {
                    const int ij = __ffs((int)unsure) - 1;
                    unsure &= unsure - 1;
                    const double acc = exact_inverse_pixel<NN, 4>(nzmask, ij, cf, tab, p.quant);
                    const int yy = ij >> 2, kk = ij & 3;
                    unsigned rrow = 0, orow = 0;
#pragma unroll
                    for (int r4 = 0; r4 < 4; r4++) if (r4 == yy) { rrow = rpxw[r4]; orow = outw[r4]; }
                    const double rpx = (double)(int)((rrow >> (8 * kk)) & 0xffu);
                    const unsigned px = clamp_trunc_u8(__dadd_rn(rpx, __dadd_rn(acc, 128.0)));           // Frame.cpp:218-242
                    orow = (orow & ~(0xffu << (8 * kk))) | (px << (8 * kk));
#pragma unroll
                    for (int r4 = 0; r4 < 4; r4++) if (r4 == yy) outw[r4] = orow;
                }
#pragma unroll
                for (int y = 0; y < 4; y++)
                    *reinterpret_cast<unsigned *>(cur_rw_i + (size_t)(byi * N + y) * p.pitch + (size_t)bxi * N) = outw[y];
            } else {
            double X[NN];
#pragma unroll
            for (int i = 0; i < NN; i++) X[i] = 0.0;
#pragma unroll 1
            for (int uv = 0; uv < NN; uv++) {
                const int c = cf[tab->izz[uv]];
                if (c != 0) {
                    const double d = __dmul_rn((double)c, p.quant.m[uv]);
                    const double *t = tab->inv + uv * NN;
#pragma unroll
                    for (int ij = 0; ij < NN; ij++) X[ij] = __dadd_rn(X[ij], __dmul_rn(__ldg(t + ij), d));
                }
            }
#pragma unroll
            for (int y = 0; y < N; y++) {
                unsigned outw = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const double rpx = (double)(int)__ldg(ref_i + (size_t)(ky + y) * p.pitch + kx + k);
                    const unsigned v = clamp_trunc_u8(__dadd_rn(rpx, __dadd_rn(X[y * N + k], 128.0)));
                    outw |= v << (8 * k);
                }
                *reinterpret_cast<unsigned *>(cur_rw_i + (size_t)(byi * N + y) * p.pitch + (size_t)bxi * N) = outw;
            }
            }
        }
        s_w[lb] = (unsigned char)w;
        s_len[lb] = (unsigned char)len;
        const unsigned bits = 4u + (p.use_rle ? w : 0) + (unsigned)len * (unsigned)w;
        s_off[lb] = bits;
        my_bits += bits;
    }
    __syncthreads();

    // ---- phase 2: offsets ----------------------------------------------------------------------------------
    unsigned T;
    unsigned excl = cta_exclusive_scan(my_bits, s_warp, &T);
#pragma unroll
    for (int r = 0; r < BPL; r++) {
        const int lb = threadIdx.x * BPL + r;
        const unsigned b = s_off[lb];
        s_off[lb] = excl;
        excl += b;
    }
    if (threadIdx.x == kThreads - 1) s_off[TB] = T;
    const bool last_tile = (tile == ntiles - 1);
    (void)last_tile;

    // ---- phase 3: pack ---------------------------------------------------------------------------------------
    // Block-centric: every lane writes its block's fields MSB-first into a shared-memory image of the tile's bits (tile
    // local alignment), whole words with plain stores, the <= 2 words it shares with its neighbours with shared-memory
    // atomicOr.  Then the tile's stream offset is resolved (look-back) and the image goes out re-aligned to the chunk
    // grid of the global stream with coalesced 128-bit stores.
    if (SMALL_OUT && !p.bits_only && (T + 127) / 128 > (unsigned)OUTCHUNKS) {                 // uniform
        // reduced staging area and a tile image that does not fit it (noise-like content, quantisers near 1): the image is
        // packed straight into the tile's slot of the scratch buffer (global atomics instead of shared ones; rare)
        uint4 *slot = reinterpret_cast<uint4 *>(p.tile_scratch + ((size_t)img * ntiles + tile) * p.slot_bytes);
        for (unsigned c = threadIdx.x; c < (T + 127) / 128; c += kThreads) slot[c] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
        pack_tile_blocks<NN, BPL>(reinterpret_cast<unsigned *>(slot), s_coef, s_off, s_w, s_len, nblk, p.use_rle);
    } else
    if (!p.bits_only) {
        unsigned *s_outw = reinterpret_cast<unsigned *>(s_out);
        const unsigned nwords = (T + 31) / 32;
        for (unsigned c = threadIdx.x; c < (nwords + 3) / 4; c += kThreads) s_out[c] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
#pragma unroll 1
        for (int r = 0; r < BPL; r++) {
            const int lb = threadIdx.x * BPL + r;
            if (lb >= nblk) break;
            const int w = s_w[lb], len = s_len[lb];
            const unsigned pos = s_off[lb];
            const unsigned bo = pos & 31u;
            unsigned *ow = s_outw + (pos >> 5);
            const unsigned *cw = reinterpret_cast<const unsigned *>(s_coef + (size_t)lb * STRIDE);
            const unsigned mask = (1u << w) - 1u;                                             // w <= 16
            // header: bit_len (low 4 bits survive, Block.cpp:381) and, with RLE, the length field (Block.cpp:393)
            unsigned long long acc = p.use_rle ? ((((unsigned long long)w & 15ull) << w) | (unsigned long long)len) : ((unsigned long long)w & 15ull);
            int nacc = (int)bo + 4 + (p.use_rle ? w : 0);
            const int nfull = len >> 1;
            const int s2 = 2 * w;
            int j = 0;
            bool odd_left = (len & 1) != 0;
            // stage A: up to the first completed word -- the only one that can hold bits of the previous block(s)
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
                // stage B: whole words that belong to this block alone
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
        __syncthreads();
        // the tile's image goes to its slot of the scratch buffer; tile_copyout_kernel re-aligns it into the stream once the
        // exclusive scan of the tile totals is known (no inter-CTA dependency in this kernel: no look-back wait)
        uint4 *slot = reinterpret_cast<uint4 *>(p.tile_scratch + ((size_t)img * ntiles + tile) * p.slot_bytes);
        for (unsigned c = threadIdx.x; c < (nwords + 3) / 4; c += kThreads) slot[c] = s_out[c];
    }
    if (threadIdx.x == 0) {
        p.tile_bits[(size_t)img * ntiles + tile] = T;
        if (p.bits_only) atomicAdd(p.bit_counter + img, (unsigned long long)T);
        else if (tile == 0 && !p.write_prefix) p.bit_base[img] = p.bit_counter[img];   // the copy-out kernel of this launch starts here
    }
    if (p.write_prefix && tile == 0 && !p.bits_only) {
        // the stream's first chunks: zero bits up to prefix_first, then the header, zero padded (the copy-out kernel merges
        // the first tile's bits into the last of them)
        const unsigned total = p.prefix_first + p.hdr.bits;
        const unsigned nw = ((total + 127) / 128) * 4;
        unsigned *o = reinterpret_cast<unsigned *>(p.out + (size_t)img * p.out_stride);
        for (unsigned i = threadIdx.x; i < nw; i += kThreads) {
            const long long hb = (long long)i * 32 - (long long)p.prefix_first;       // header bit of this word's first bit
            const int sh = (int)(((hb % 32) + 32) % 32);
            const long long wi = (hb - sh) / 32;                                      // floor division
            const unsigned hi = (wi >= 0 && wi < kHdrWordsMax) ? p.hdr.words[wi] : 0u;
            const unsigned lo = (wi + 1 >= 0 && wi + 1 < kHdrWordsMax) ? p.hdr.words[wi + 1] : 0u;
            const unsigned v = sh ? ((hi << sh) | (lo >> (32 - sh))) : hi;
            o[i] = __byte_perm(v, 0, 0x0123);
        }
        if (threadIdx.x == 0) p.bit_base[img] = total;
    }
}

// Tile image in global scratch (tile-local alignment, 32 stream bits per word, MSB first, zero beyond the last bit)
struct GlobalStreamTile {
    const unsigned *words;
    unsigned nwords;
};
__device__ __forceinline__ uint4 gather_chunk(const GlobalStreamTile &t, long long ls) {
    const long long wi = ls >> 5;
    const unsigned sh = (unsigned)(ls & 31);
    unsigned w[5];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const long long i = wi + k;
        w[k] = (i >= 0 && i < (long long)t.nwords) ? __ldg(t.words + i) : 0u;
    }
    uint4 o;
    o.x = __byte_perm(__funnelshift_l(w[1], w[0], sh), 0, 0x0123);
    o.y = __byte_perm(__funnelshift_l(w[2], w[1], sh), 0, 0x0123);
    o.z = __byte_perm(__funnelshift_l(w[3], w[2], sh), 0, 0x0123);
    o.w = __byte_perm(__funnelshift_l(w[4], w[3], sh), 0, 0x0123);
    return o;
}

// A group of consecutive tile images seen as one piece of the stream: off[j] = first bit of tile j relative to the group.
struct GroupStreamTiles {
    const uint8_t *slots;          // slot of the group's first tile
    size_t slot_bytes;
    const unsigned *off;           // shared memory, n + 1 entries
    unsigned n;
};
__device__ __forceinline__ uint4 gather_chunk(const GroupStreamTiles &g, long long ls) {
    const unsigned lo = (unsigned)max(ls, 0ll);
    unsigned j = 0;
#pragma unroll
    for (unsigned step = 16; step >= 1; step >>= 1)
        if (j + step < g.n && g.off[j + step] <= lo) j += step;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    for (;;) {
        GlobalStreamTile t;
        t.words = reinterpret_cast<const unsigned *>(g.slots + (size_t)j * g.slot_bytes);
        t.nwords = (g.off[j + 1] - g.off[j] + 31) / 32;
        const uint4 a = gather_chunk(t, ls - (long long)g.off[j]);
        v.x |= a.x; v.y |= a.y; v.z |= a.z; v.w |= a.w;
        if (j + 1 >= g.n || ls + (long long)kChunkBits <= (long long)g.off[j + 1]) break;     // the chunk ends inside tile j
        j++;
    }
    return v;
}

// Re-aligns the tile images into the stream.  One CTA per group of kTilesPerGroup consecutive tiles: it sums the bit totals of
// every earlier tile of the stream (<= 32 KiB of L2 reads, all of them final: the tile kernel has completed) and then writes
// the group's chunks, a thread per 128-bit chunk.  No scan kernel, no inter-CTA wait; only the chunks two groups share go
// through the hand-off records.
constexpr unsigned kTilesPerGroup = 8;
__global__ void __launch_bounds__(kThreads) tile_copyout_kernel(const EncodeParams p) {
    __shared__ unsigned long long s_part[kThreads / 32];
    __shared__ unsigned s_goff[kTilesPerGroup + 1];
    pdl_wait();                                              // everything the preceding kernel wrote is visible after this
    const unsigned img = blockIdx.y, ntiles = p.tiles_per_image;
    const unsigned t0 = blockIdx.x * kTilesPerGroup, t1 = min(t0 + kTilesPerGroup, ntiles);
    const unsigned *tb = p.tile_bits + (size_t)img * ntiles;
    unsigned long long sum = 0;
    for (unsigned i = threadIdx.x; i < t0; i += kThreads) sum += tb[i];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    if ((threadIdx.x & 31u) == 0) s_part[threadIdx.x >> 5] = sum;
    if (threadIdx.x < 32) {                                    // kTilesPerGroup <= 32: one warp scans the group's totals
        const unsigned t = t0 + threadIdx.x;
        const unsigned v = (t < t1) ? tb[t] : 0u;
        unsigned inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned o = __shfl_up_sync(0xffffffffu, inc, d); if ((int)threadIdx.x >= d) inc += o; }
        if (threadIdx.x < kTilesPerGroup) s_goff[threadIdx.x + 1] = inc;
        if (threadIdx.x == 0) s_goff[0] = 0;
    }
    __syncthreads();
    unsigned long long G = p.bit_base[img];
#pragma unroll
    for (int w = 0; w < kThreads / 32; w++) G += s_part[w];
    ScanState st = p.scan;
    st.bnd += (size_t)img * ntiles;
    GroupStreamTiles g;
    g.slots = p.tile_scratch + ((size_t)img * ntiles + t0) * p.slot_bytes;
    g.slot_bytes = p.slot_bytes;
    g.off = s_goff;
    g.n = t1 - t0;
    const unsigned T = s_goff[g.n];
    tile_write_chunks(g, st, blockIdx.x, t0 == 0, t1 == ntiles, G, T, p.out + (size_t)img * p.out_stride, p.out_cap, p.err);
    if (t1 == ntiles && threadIdx.x == 0) {
        p.bit_counter[img] = G + T;
        if (p.out_bits) p.out_bits[img] = G + T;
    }
}

// Copy-out variant 1 (ie_set_option("copyout_variant", 1)): same result as tile_copyout_kernel, with a short path for the
// chunks that lie inside one tile image and inside the group (all but ~1 %): 32-bit positions, a 3-step search over the
// group's <= 8 tile offsets, five unconditional word loads (a word past the tile's last one is only ever read when the
// funnel shift ignores it; the slot has two chunks of slack), funnel shifts, one 128-bit store.  Every other chunk -- first
// and last of the group, chunks that span two tiles -- takes the generic path of tile_write_chunks, chunk by chunk.
// U = chunks per thread in flight (copyout_variant 1: U = 1, 2: U = 4).
template <int U, unsigned TPG>
__global__ void __launch_bounds__(kThreads) tile_copyout_fast_kernel(const EncodeParams p) {
    __shared__ unsigned long long s_part[kThreads / 32];
    __shared__ unsigned s_goff[TPG + 1];
    pdl_wait();
    const unsigned img = blockIdx.y, ntiles = p.tiles_per_image;
    const unsigned t0 = blockIdx.x * TPG, t1 = min(t0 + TPG, ntiles);
    const unsigned *tb = p.tile_bits + (size_t)img * ntiles;
    unsigned long long sum = 0;
    if ((reinterpret_cast<size_t>(tb) & 15) == 0) {
        // earlier tiles' totals: 128-bit loads (t0 is a multiple of 4)
        const uint4 *tb4 = reinterpret_cast<const uint4 *>(tb);
        for (unsigned i = threadIdx.x; i < t0 / 4; i += kThreads) {
            const uint4 v = __ldg(tb4 + i);
            sum += (unsigned long long)v.x + v.y + v.z + v.w;
        }
    } else {
        for (unsigned i = threadIdx.x; i < t0; i += kThreads) sum += tb[i];
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    if ((threadIdx.x & 31u) == 0) s_part[threadIdx.x >> 5] = sum;
    if (threadIdx.x < 32) {
        const unsigned t = t0 + threadIdx.x;
        const unsigned v = (t < t1) ? tb[t] : 0u;
        unsigned inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned o = __shfl_up_sync(0xffffffffu, inc, d); if ((int)threadIdx.x >= d) inc += o; }
        if (threadIdx.x < TPG) s_goff[threadIdx.x + 1] = inc;      // entries past the group's last tile repeat its end
        if (threadIdx.x == 0) s_goff[0] = 0;
    }
    __syncthreads();
    unsigned long long G = p.bit_base[img];
#pragma unroll
    for (int w = 0; w < kThreads / 32; w++) G += s_part[w];
    ScanState st = p.scan;
    st.bnd += (size_t)img * ntiles;
    GroupStreamTiles g;
    g.slots = p.tile_scratch + ((size_t)img * ntiles + t0) * p.slot_bytes;
    g.slot_bytes = p.slot_bytes;
    g.off = s_goff;
    g.n = t1 - t0;
    const unsigned T = s_goff[g.n];
    uint8_t *out = p.out + (size_t)img * p.out_stride;
    if (T != 0) {
        const unsigned long long c0 = G / kChunkBits, c1 = (G + T - 1) / kChunkBits;
        const unsigned nchunks = (unsigned)(c1 - c0) + 1u;
        const int g0 = (int)(G % kChunkBits);
        const bool first_group = (t0 == 0), last_group = (t1 == ntiles);
        const bool head_shared = g0 != 0;
        const bool tail_shared = ((G + T) % kChunkBits) != 0 && !last_group;
        const unsigned slot_words = (unsigned)(p.slot_bytes / 4);
        const unsigned *slot0 = reinterpret_cast<const unsigned *>(g.slots);
        uint4 *dst0 = reinterpret_cast<uint4 *>(out) + c0;
        const unsigned long long cap_chunks = p.out_cap / 16ull;
        // U chunks per thread in flight: the positions of all U are resolved and their loads issued before the first is
        // shifted and stored (a chunk's loads depend on its search, so one chunk at a time is one L2 round trip per chunk)
        for (unsigned kb = threadIdx.x; kb < nchunks; kb += kThreads * U) {
            unsigned w[U][5], sh[U];
            bool fast[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const unsigned k = kb + (unsigned)u * kThreads;
                const int ls = (int)(k * (unsigned)kChunkBits) - g0;          // group-local bit of the chunk's first bit
                fast[u] = (k < nchunks) && (k != 0) && (k + 1 != nchunks) && (c0 + k < cap_chunks);
                unsigned j = 0;
                const unsigned lo = (unsigned)ls;                             // k != 0: ls > 0
                if (fast[u]) {
                    // s_goff is non-decreasing and has TPG + 1 valid entries (the tail repeats the group's end)
#pragma unroll
                    for (unsigned step = TPG / 2; step >= 1; step >>= 1)
                        if (s_goff[j + step] <= lo) j += step;
                    fast[u] = (j < g.n) && (lo + (unsigned)kChunkBits <= s_goff[j + 1]);   // the chunk ends inside tile j
                }
                if (fast[u]) {
                    const unsigned tl = lo - s_goff[j];                       // tile-local bit
                    sh[u] = tl & 31u;
                    const unsigned *wp = slot0 + (size_t)j * slot_words + (tl >> 5);
#pragma unroll
                    for (int i = 0; i < 5; i++) w[u][i] = __ldg(wp + i);
                }
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const unsigned k = kb + (unsigned)u * kThreads;
                if (k >= nchunks) break;
                if (fast[u]) {
                    uint4 o;
                    o.x = __byte_perm(__funnelshift_l(w[u][1], w[u][0], sh[u]), 0, 0x0123);
                    o.y = __byte_perm(__funnelshift_l(w[u][2], w[u][1], sh[u]), 0, 0x0123);
                    o.z = __byte_perm(__funnelshift_l(w[u][3], w[u][2], sh[u]), 0, 0x0123);
                    o.w = __byte_perm(__funnelshift_l(w[u][4], w[u][3], sh[u]), 0, 0x0123);
                    dst0[k] = o;
                    continue;
                }
                // generic path (tile_write_chunks, one chunk)
                const int ls = (int)(k * (unsigned)kChunkBits) - g0;
                const unsigned long long c = c0 + k;
                uint4 v = gather_chunk(g, (long long)ls);
                if ((c + 1) * 16ull > p.out_cap) { if (p.err) atomicExch(p.err, IE_ENOSPC); continue; }
                uint4 *dst = reinterpret_cast<uint4 *>(out) + c;
                const bool is_head = (k == 0) && head_shared;
                const bool is_tail = (k + 1 == nchunks) && tail_shared;
                if (is_head && first_group) {
                    const uint4 o = *dst;
                    v.x |= o.x; v.y |= o.y; v.z |= o.z; v.w |= o.w;
                } else if (is_head || is_tail) {
                    TileBoundary *bd = &st.bnd[is_head ? blockIdx.x - 1 : blockIdx.x];
                    unsigned *dw = reinterpret_cast<unsigned *>(dst);
                    const unsigned mine[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const unsigned long long old = atomicAdd(&bd->w[i], (1ull << 32) | (unsigned long long)mine[i]);
                        if ((old >> 32) == 1ull) {
                            dw[i] = (unsigned)old | mine[i];
                            atomicExch(&bd->w[i], 0ull);
                        }
                    }
                    continue;
                }
                *dst = v;
            }
        }
    }
    if (t1 == ntiles && threadIdx.x == 0) {
        p.bit_counter[img] = G + T;
        if (p.out_bits) p.out_bits[img] = G + T;
    }
}

// Copy-out by tile and by 32-bit word (copyout_variant 3).  The chunk-centric kernels above search, for every 128-bit chunk of the
// stream, the tile it comes from; here a warp takes a whole tile image: its shift against the stream's word grid is one constant,
// consecutive lanes read consecutive words of the image and write consecutive words of the stream (two loads, one funnel shift, one
// byte swap, one store per word -- about a quarter of the instructions per byte of the chunk-centric path).
// Ownership: a stream word belongs to the tile that holds its FIRST bit.  The owner of a word that ends in the next tile ORs in that
// tile's first bits, read from its image in the scratch buffer (complete: the tile kernel has finished) -- so neither inside a group
// nor between groups is there a hand-off, an atomic or a pre-zeroed stream.  (A full tile holds >= 512 bits, so a word never spans
// three tiles; only the stream's last tile can be shorter than a word, and nothing follows it.)  The word that the stream's first
// tile shares with what precedes it (header / earlier appends: zero padded, pack.cuh) is read, merged and written by that tile; the
// stream's last tile pads its last 128-bit chunk with zeroes as the append contract wants.
#ifndef IE_WORDS_U
#define IE_WORDS_U 8
#endif
template <unsigned TPG, unsigned NT>
__global__ void __launch_bounds__(NT) tile_copyout_words_kernel(const EncodeParams p) {
    __shared__ unsigned long long s_part[NT / 32];
    __shared__ unsigned s_goff[TPG + 1];
    pdl_wait();
    const unsigned img = blockIdx.y, ntiles = p.tiles_per_image;
    const unsigned t0 = blockIdx.x * TPG, t1 = min(t0 + TPG, ntiles);
    const unsigned *tb = p.tile_bits + (size_t)img * ntiles;
    unsigned long long sum = 0;
    if ((reinterpret_cast<size_t>(tb) & 15) == 0) {
        // earlier tiles' totals (t0 is a multiple of 4), four 128-bit loads per thread in flight: the groups at the end of a large
        // image read 32 KiB here, and a load per round trip made their CTAs the kernel's critical path
        const uint4 *tb4 = reinterpret_cast<const uint4 *>(tb);
        const unsigned n4 = t0 / 4;
        unsigned i = threadIdx.x;
        for (; i + 3 * NT < n4; i += 4 * NT) {
            const uint4 a = __ldg(tb4 + i), b = __ldg(tb4 + i + NT), c = __ldg(tb4 + i + 2 * NT), d = __ldg(tb4 + i + 3 * NT);
            sum += ((unsigned long long)a.x + a.y + a.z + a.w) + ((unsigned long long)b.x + b.y + b.z + b.w)
                 + ((unsigned long long)c.x + c.y + c.z + c.w) + ((unsigned long long)d.x + d.y + d.z + d.w);
        }
        for (; i < n4; i += NT) {
            const uint4 v = __ldg(tb4 + i);
            sum += (unsigned long long)v.x + v.y + v.z + v.w;
        }
    } else {
        for (unsigned i = threadIdx.x; i < t0; i += NT) sum += tb[i];
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    if ((threadIdx.x & 31u) == 0) s_part[threadIdx.x >> 5] = sum;
    if (threadIdx.x < 32) {
        const unsigned t = t0 + threadIdx.x;
        const unsigned v = (t < t1) ? tb[t] : 0u;
        unsigned inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned o = __shfl_up_sync(0xffffffffu, inc, d); if ((int)threadIdx.x >= d) inc += o; }
        if (threadIdx.x < TPG) s_goff[threadIdx.x + 1] = inc;
        if (threadIdx.x == 0) s_goff[0] = 0;
    }
    __syncthreads();
    unsigned long long G = p.bit_base[img];
#pragma unroll
    for (int w = 0; w < NT / 32; w++) G += s_part[w];
    const unsigned n = t1 - t0;
    const unsigned slot_words = (unsigned)(p.slot_bytes / 4);
    const unsigned *slot0 = reinterpret_cast<const unsigned *>(p.tile_scratch + ((size_t)img * ntiles + t0) * p.slot_bytes);
    unsigned *out = reinterpret_cast<unsigned *>(p.out + (size_t)img * p.out_stride);
    const unsigned long long cap_words = (p.out_cap / 16ull) * 4ull;       // whole chunks only, as the chunk-centric kernels
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    bool overflow = false;
    for (unsigned j = warp; j < n; j += NT / 32) {
        const unsigned Tj = s_goff[j + 1] - s_goff[j];
        if (Tj == 0) continue;
        const unsigned long long B0 = G + s_goff[j], B1 = B0 + Tj;         // the tile's bits in the stream
        const unsigned long long q0 = (B0 + 31) >> 5, q1 = (B1 + 31) >> 5; // words whose first bit lies in the tile
        const unsigned sh = (unsigned)((32u - (unsigned)(B0 & 31u)) & 31u);// image bit of word q0's first bit
        const unsigned nw = (Tj + 31) >> 5;                                // words of the image (zero beyond the last bit)
        const unsigned *I = slot0 + (size_t)j * slot_words;
        const bool has_next = (t0 + j + 1) < ntiles;
        const unsigned tail = (unsigned)(B1 & 31u);                        // bits of the tile in its last, shared word (0: none shared)
        if (lane == 0 && t0 + j == 0 && (B0 & 31u) != 0) {
            // the word the stream's first tile shares with the header / the earlier appends
            const unsigned long long q = B0 >> 5;
            if (q < cap_words) out[q] |= __byte_perm(__ldg(I) >> (unsigned)(B0 & 31u), 0, 0x0123);
            else overflow = true;
        }
        const unsigned cnt = (unsigned)(q1 - q0);
        const bool fits = q1 <= cap_words;                                 // uniform
        unsigned *ow = out + q0;
        unsigned k0 = 0;
        // whole rounds of 32 * U words well inside the image: no bounds, no neighbour (32-bit indices throughout)
        if (fits) {
            for (; k0 + 32u * IE_WORDS_U < cnt && k0 + 32u * IE_WORDS_U < nw; k0 += 32u * IE_WORDS_U) {
                unsigned lo[IE_WORDS_U], hi[IE_WORDS_U];
                const unsigned *Ik = I + k0 + lane;
#pragma unroll
                for (int u = 0; u < IE_WORDS_U; u++) { lo[u] = __ldg(Ik + 32 * u); hi[u] = __ldg(Ik + 32 * u + 1); }
#pragma unroll
                for (int u = 0; u < IE_WORDS_U; u++) ow[k0 + lane + 32u * u] = __byte_perm(__funnelshift_l(hi[u], lo[u], sh), 0, 0x0123);
            }
        }
        for (; k0 < cnt; k0 += 32u * IE_WORDS_U) {
            unsigned lo[IE_WORDS_U], hi[IE_WORDS_U];
#pragma unroll
            for (int u = 0; u < IE_WORDS_U; u++) {
                const unsigned k = k0 + (unsigned)u * 32u + lane;          // word q0 + k <- image words k, k + 1
                lo[u] = (k < nw && k < cnt) ? __ldg(I + k) : 0u;
                hi[u] = (k + 1 < nw && k < cnt) ? __ldg(I + k + 1) : 0u;
            }
#pragma unroll
            for (int u = 0; u < IE_WORDS_U; u++) {
                const unsigned k = k0 + (unsigned)u * 32u + lane;
                if (k >= cnt) break;
                unsigned v = __funnelshift_l(hi[u], lo[u], sh);
                if (k + 1 == cnt && tail != 0 && has_next) v |= __ldg(I + slot_words) >> tail;      // first bits of the next tile
                if (fits || q0 + k < cap_words) ow[k] = __byte_perm(v, 0, 0x0123);
                else overflow = true;
            }
        }
        if (!has_next && lane < 4) {
            // end of the stream: the rest of its last chunk is zero
            const unsigned long long q = q1 + lane, qe = ((B1 + 127) >> 7) << 2;
            if (q < qe) { if (q < cap_words) out[q] = 0u; else overflow = true; }
        }
    }
    if (overflow && p.err) atomicExch(p.err, IE_ENOSPC);
    if (t1 == ntiles && threadIdx.x == 0) {
        const unsigned long long total = G + s_goff[n];
        p.bit_counter[img] = total;
        if (p.out_bits) p.out_bits[img] = total;
    }
}

// copy-out kernel: 2 = short path for interior chunks, four chunks per thread in flight (the default: 0.1125 -> 0.1068 ms on
// config 2 with tile-kernel variant 2, profiles/r1_ab_copyout_v9.log), 1 = short path, one chunk at a time, 0 = the generic
// kernel of versions v7/v8.  Eight chunks in flight and groups of 4 or 16 tiles measured slower
// (profiles/r1_ab_copyout_groups_v9.log).
std::atomic<int> g_copyout_variant{3};
#ifndef IE_WORDS_THREADS
#define IE_WORDS_THREADS 128
#endif
constexpr unsigned kWordsThreads = IE_WORDS_THREADS;

int launch_tile_copyout(const EncodeParams &p, unsigned images, cudaStream_t stream) {
    const int cv = g_copyout_variant.load();
    dim3 cgrid((p.tiles_per_image + kTilesPerGroup - 1) / kTilesPerGroup, images);
    // programmatic dependent launch: the copy-out grid is set up while the tile kernel drains and waits on the device
    // (griddepcontrol.wait) for its results instead of being launched after the tile kernel has completed
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = cgrid; cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (cv == 1) IE_CUDA(cudaLaunchKernelEx(&cfg, tile_copyout_fast_kernel<1, kTilesPerGroup>, p));
    else if (cv == 2) IE_CUDA(cudaLaunchKernelEx(&cfg, tile_copyout_fast_kernel<4, kTilesPerGroup>, p));
    else if (cv == 3) { cfg.blockDim = dim3(kWordsThreads); IE_CUDA(cudaLaunchKernelEx(&cfg, tile_copyout_words_kernel<kTilesPerGroup, kWordsThreads>, p)); }
    else IE_CUDA(cudaLaunchKernelEx(&cfg, tile_copyout_kernel, p));
    count_launch();
    IE_CUDA(cudaGetLastError());
    return IE_OK;
}

__global__ void __launch_bounds__(256) tile_totals_kernel(const unsigned *__restrict__ tile_bits, unsigned ntiles, unsigned long long add,
                                                          unsigned long long *d_total) {
    pdl_wait();
    __shared__ unsigned long long s_part[8];
    const unsigned *tb = tile_bits + (size_t)blockIdx.x * ntiles;
    unsigned long long sum = 0;
    for (unsigned i = threadIdx.x; i < ntiles; i += 256) sum += tb[i];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    if ((threadIdx.x & 31u) == 0) s_part[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = add;
        for (int w = 0; w < 8; w++) t += s_part[w];
        d_total[blockIdx.x] = t;
    }
}

int launch_tile_totals(const unsigned *tile_bits, unsigned ntiles, unsigned images, unsigned long long add, unsigned long long *d_total,
                       cudaStream_t stream) {
    IE_CUDA(launch_pdl(tile_totals_kernel, dim3(images), dim3(256), 0, stream, tile_bits, ntiles, add, d_total));
    count_launch();
    IE_CUDA(cudaGetLastError());
    return IE_OK;
}

std::atomic<int> g_encode_pad_smem{0};
template <int N, int BPL, bool PF, bool FAST, int VAR = 0>
static int launch_cfg(const EncodeParams &p, unsigned images, cudaStream_t stream) {
    constexpr int TB = kThreads * BPL;
    constexpr int STRIDE = N * N + 2;
    constexpr int MAXCHUNKS = (TB * (4 + 16 + 16 * N * N) + 127) / 128 + 2;
    constexpr int OUTCHUNKS = var_small_out(VAR) ? kSmallOutChunks : MAXCHUNKS;
    size_t smem = (((size_t)TB * STRIDE * sizeof(short) + (TB + 1) * sizeof(unsigned) + 15) & ~(size_t)15) + (size_t)OUTCHUNKS * 16 +
                  (kQueueCap + TB + (kQueueCap + TB) / 2) * sizeof(unsigned) + 3 * TB;
    static size_t configured = 0;
    smem += (size_t)g_encode_pad_smem.load();     // occupancy experiments only (ie_set_option("encode_pad_smem", bytes)); 0 by default
    if (configured != smem) {
        IE_CUDA(cudaFuncSetAttribute(encode_tiles_kernel<N, BPL, PF, FAST, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid(p.tiles_per_image, images);
    encode_tiles_kernel<N, BPL, PF, FAST, VAR><<<grid, kThreads, smem, stream>>>(p);
    count_launch();
    if (!p.bits_only && p.phase == 0) return launch_tile_copyout(p, images, stream);
    IE_CUDA(cudaGetLastError());
    return IE_OK;
}

unsigned encode_tile_blocks(int N) { return N == 8 ? kThreads * 1 : kThreads * 4; }
size_t encode_tile_slot_bytes(int N) {
    const size_t tb = encode_tile_blocks(N);
    return ((tb * (4 + 16 + 16 * (size_t)N * N) + 127) / 128 + 2) * 16;
}

std::atomic<int> g_exact_transform{0};
// instantiation of the tile kernel for image blocks / I-frames: 2 = packed f32x2 transform + lean quantise bookkeeping
// (transform_fast.cuh, lean::; the default since it measured 0.1298 -> 0.1119 ms on config 2, profiles/r1_ab_variants_v9.log),
// 1 = lean quantise only, 0 = the scalar kernel of versions v0..v8 (kept for A/B and cross-checks)
std::atomic<int> g_encode_variant{2};

int launch_encode_tiles(int N, const EncodeParams &p, unsigned images, cudaStream_t stream) {
    const bool exact = g_exact_transform.load() != 0;
    int var = exact ? 0 : g_encode_variant.load();
    if (var == 8) var = 2;          // batches, split encodes and bit-count passes of the fused variant use the default tile kernel
    if (var == 1) {
        if (N == 8) return launch_cfg<8, 1, false, true, 1>(p, images, stream);
        if (N == 4) return launch_cfg<4, 4, false, true, 1>(p, images, stream);
    } else if (var == 2) {
        if (N == 8) return launch_cfg<8, 1, false, true, 2>(p, images, stream);
        if (N == 4) return launch_cfg<4, 4, false, true, 2>(p, images, stream);
    } else if (var == 3) {
        if (N == 8) return launch_cfg<8, 1, false, true, 3>(p, images, stream);
        if (N == 4) return launch_cfg<4, 4, false, true, 3>(p, images, stream);
    } else if (var == 4) {
        if (N == 8) return launch_cfg<8, 1, false, true, 4>(p, images, stream);
        if (N == 4) return launch_cfg<4, 4, false, true, 4>(p, images, stream);
    }
    if (var == 9) { if (N == 8) return launch_cfg<8, 1, false, true, 9>(p, images, stream); if (N == 4) return launch_cfg<4, 4, false, true, 9>(p, images, stream); }
    if (var == 5) { if (N == 8) return launch_cfg<8, 1, false, true, 5>(p, images, stream); if (N == 4) return launch_cfg<4, 4, false, true, 5>(p, images, stream); }
    if (var == 6) { if (N == 8) return launch_cfg<8, 1, false, true, 6>(p, images, stream); if (N == 4) return launch_cfg<4, 4, false, true, 6>(p, images, stream); }
    if (var == 7) { if (N == 8) return launch_cfg<8, 1, false, true, 7>(p, images, stream); if (N == 4) return launch_cfg<4, 4, false, true, 7>(p, images, stream); }
    if (N == 8) return exact ? launch_cfg<8, 1, false, false>(p, images, stream) : launch_cfg<8, 1, false, true>(p, images, stream);
    if (N == 4) return exact ? launch_cfg<4, 4, false, false>(p, images, stream) : launch_cfg<4, 4, false, true>(p, images, stream);
    set_error("block size must be 4 or 8");
    return IE_EINVAL;
}

// P-frame tiles: 2 = residual, transform, quantisation and reconstruction in packed f32x2 operations, reference rows read as
// words (default); 0 = the scalar kernel it replaced
std::atomic<int> g_pframe_variant{2};
int launch_pframe_tiles(const EncodeParams &p, unsigned images, cudaStream_t stream) {
    if (g_exact_transform.load()) return launch_cfg<4, 4, true, false>(p, images, stream);
    if (g_pframe_variant.load() == 2) return launch_cfg<4, 4, true, true, 2>(p, images, stream);
    return launch_cfg<4, 4, true, true>(p, images, stream);
}

}  // namespace ie
