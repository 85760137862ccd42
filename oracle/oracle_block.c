/*
 * TEST INFRASTRUCTURE ONLY -- the CPU oracle for the block-transform hot path.
 *
 * A plain-C, single-threaded restatement of what ThenTech/ImageEncoder computes on the path
 *   Block<4>/Block<8>/MacroBlock  ->  BitStream  ->  (Huffman: see oracle_huffman.cpp)  ->  video P-frames.
 * Every function cites the reference file:line it restates.  Nothing in the product (imageencoder_b200/,
 * include/, the CLIs) may include, link, import or execute this file; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg use it, and only as the checker.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py compares this file byte-for-byte against the
 * compiled reference itself (oracle/_ref/, built by oracle/build_ref.sh from /root/reference) on every
 * bin/ex*.raw sample, on 8x8 variants and on synthetic video; tests/golden/ holds the committed hashes.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math  (NO -march=native: an FMA changes the bitstream)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAXN 8
#define ORC_MAXNN 64
#define ORC_MB 16

/* ------------------------------------------------------------------------------------------------
 * Bit I/O : BitStream.cpp:14-77.  MSB-first, one buffer, reads past the end give 0 (BitStream.cpp:17-20),
 * the writer's buffer is zero-initialised (utils.hpp:443-446) so pad bits are 0.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { uint8_t *buf; size_t cap; size_t pos; int overflow; } orc_bw;
typedef struct { const uint8_t *buf; size_t size; size_t pos; } orc_br;

static void bw_put(orc_bw *w, unsigned len, uint32_t value) {          /* BitStream.cpp:61-77 */
    for (unsigned p = 0; p < len; p++) {
        unsigned bit = (value >> (len - 1 - p)) & 1u;
        size_t byte = w->pos >> 3;
        if (byte >= w->cap) { w->overflow = 1; w->pos++; continue; }
        if (bit) w->buf[byte] |= (uint8_t)(1u << (7 - (w->pos & 7)));
        else     w->buf[byte] &= (uint8_t)~(1u << (7 - (w->pos & 7)));
        w->pos++;
    }
}
static uint32_t br_get(orc_br *r, unsigned len) {                       /* BitStream.cpp:14-40 */
    uint32_t v = 0;
    for (unsigned i = 0; i < len; i++) {
        size_t byte = r->pos >> 3;
        uint32_t bit = 0;
        if (byte < r->size) { bit = (r->buf[byte] >> (7 - (r->pos & 7))) & 1u; r->pos++; }
        /* past the end: returns 0 and does NOT advance (BitStream.cpp:17-20) */
        v |= bit << (len - i - 1);
    }
    return v;
}

/* utils.hpp:210-216 with the -mlzcnt semantics ffs(0) = 0 */
static unsigned orc_ffs(uint32_t v) { unsigned n = 0; while (v) { n++; v >>= 1; } return n; }
/* utils.hpp:226-243 : minimal two's complement width of an int16 (0 -> 1, -1 -> 1, 1 -> 2, 127 -> 8, -128 -> 8) */
static unsigned orc_bits_needed(int16_t value) {
    unsigned bits = 1;
    while ((int16_t)((int16_t)((value & ((1 << bits) - 1)) << (16 - bits)) >> (16 - bits)) != value) bits++;
    return bits;
}
/* utils.hpp:265-269 */
static int16_t orc_shift_signed(uint32_t value, unsigned src_bits) {
    unsigned sh = 16 - src_bits;
    return (int16_t)((int16_t)(uint16_t)(value << sh) >> sh);
}
size_t orc_round_to_byte(size_t bits) { return (bits + (8u - (bits % 8u)) % 8u) / 8u; }   /* utils.hpp:253-255 */

unsigned orc_bits_needed_pub(int v) { return orc_bits_needed((int16_t)v); }
unsigned orc_ffs_pub(uint32_t v) { return orc_ffs(v); }

/* ------------------------------------------------------------------------------------------------
 * Tables : algo.cpp:68-87 (zigzag), algo.cpp:294-297 (C), algo.cpp:312,318-319 (cos arguments)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int N;
    double cs[ORC_MAXNN];   /* cs[i*N+u] = cos(((2i+1)*u)*f), f = M_PI_2/N */
    double cc[ORC_MAXNN];   /* cc[u*N+v] = C(u)*C(v)                       */
    uint8_t zz[ORC_MAXNN];  /* zz[k] = y*N+x of the k-th zigzag position   */
} orc_tab;

static int zz_less(int N, int a, int b) {                               /* algo.cpp:33-37, 73-83 */
    int ax = a % N, ay = a / N, bx = b % N, by = b / N;
    int ag = ax + ay, bg = bx + by;
    int ac = (((int8_t)(ax - ay)) & 1) ? ay : ax;
    int bc = (((int8_t)(bx - by)) & 1) ? by : bx;
    return (ag == bg) ? (ac < bc) : (ag < bg);
}
static void orc_tab_init(orc_tab *t, int N) {
    t->N = N;
    const double f = M_PI_2 / (double)N;
    for (int i = 0; i < N; i++)
        for (int u = 0; u < N; u++)
            t->cs[i * N + u] = cos((double)(2.0 * i + 1.0) * (double)u * f);
    for (int u = 0; u < N; u++)
        for (int v = 0; v < N; v++) {
            double cu = (u == 0) ? 0.5 : M_SQRT1_2, cv = (v == 0) ? 0.5 : M_SQRT1_2;
            t->cc[u * N + v] = cu * cv;
        }
    /* keys are unique (within a group all elements use the same one of x|y), so any sort gives std::sort's result */
    int NN = N * N;
    for (int i = 0; i < NN; i++) t->zz[i] = (uint8_t)i;
    for (int i = 1; i < NN; i++) {
        uint8_t k = t->zz[i]; int j = i - 1;
        while (j >= 0 && zz_less(N, k, t->zz[j])) { t->zz[j + 1] = t->zz[j]; j--; }
        t->zz[j + 1] = k;
    }
}
int orc_zigzag(int N, uint8_t *zz) {
    if (N != 4 && N != 8) return -1;
    orc_tab t; orc_tab_init(&t, N); memcpy(zz, t.zz, (size_t)N * N); return 0;
}
/* host tables as the device code needs them (doubles, glibc cos) */
int orc_cos_table(int N, double *cs, double *cc) {
    if (N != 4 && N != 8) return -1;
    orc_tab t; orc_tab_init(&t, N);
    memcpy(cs, t.cs, sizeof(double) * N * N); memcpy(cc, t.cc, sizeof(double) * N * N); return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Forward: Block.cpp:138-153 + algo.cpp:309-331.   x[] holds pixel (or residual) values; -128 is applied here.
 *   t = +0.0; for i: for j: t += (cs[i][u]*cs[j][v]) * x[i][j];  e = t * (C(u)*C(v));  q = round(e / m)
 * ---------------------------------------------------------------------------------------------- */
static void orc_fdct_quant(const orc_tab *t, const double *x_in, const double *m, double *q) {
    const int N = t->N;
    double x[ORC_MAXNN];
    for (int i = 0; i < N * N; i++) x[i] = x_in[i] + (-128.0);          /* Block.cpp:141-143 */
    for (int u = 0; u < N; u++)
        for (int v = 0; v < N; v++) {
            double acc = 0.0;
            for (int i = 0; i < N; i++)
                for (int j = 0; j < N; j++)
                    acc += t->cs[i * N + u] * t->cs[j * N + v] * x[i * N + j];   /* algo.cpp:318-320 */
            acc *= t->cc[u * N + v];                                             /* algo.cpp:325 */
            q[u * N + v] = round(acc / m[u * N + v]);                            /* Block.cpp:152 */
        }
}
/* Inverse: Block.cpp:162-177 + algo.cpp:343-363.  out = IDCT(c*m) + 128 (doubles, not yet clamped) */
static void orc_dequant_idct(const orc_tab *t, const double *c, const double *m, double *out) {
    const int N = t->N;
    double d[ORC_MAXNN], tmp[ORC_MAXNN];
    for (int i = 0; i < N * N; i++) { d[i] = c[i] * m[i]; tmp[i] = 0.0; }       /* Block.cpp:165-168 */
    for (int u = 0; u < N; u++)
        for (int v = 0; v < N; v++)
            for (int i = 0; i < N; i++)
                for (int j = 0; j < N; j++)
                    tmp[i * N + j] += t->cc[u * N + v] * t->cs[i * N + u] * t->cs[j * N + v] * d[u * N + v]; /* algo.cpp:352-355 */
    for (int i = 0; i < N * N; i++) out[i] = tmp[i] + 128.0;                    /* Block.cpp:173-175 */
}
static uint8_t orc_clamp_u8(double v) {                                          /* Block.cpp:103 : truncation */
    if (v < 0.0) v = 0.0;
    if (v > 255.0) v = 255.0;
    return (uint8_t)v;
}

/* ------------------------------------------------------------------------------------------------
 * RLE info: Block.cpp:185-232.   zzc[k] = int16 coefficient at zigzag position k.
 *   data      = index of last non-zero + 1
 *   data_bits = max(max bits_needed(non-zero), ffs(data))
 *   last_zeroes = zeroes in front of the last non-zero (rle_Data->back()->zeroes), 0 when no non-zero
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int16_t zzc[ORC_MAXNN]; unsigned data_bits; int data; int last_zeroes; } orc_rle;

static void orc_make_rle(const orc_tab *t, const double *q, orc_rle *r) {
    const int NN = t->N * t->N;
    int zeroes = 0;
    r->data_bits = 0; r->data = 0; r->last_zeroes = 0;
    for (int k = 0; k < NN; k++) {
        int16_t data = (int16_t)q[t->zz[k]];                                     /* Block.cpp:205 */
        r->zzc[k] = data;
        if (data == 0) { zeroes++; }
        else {
            unsigned b = orc_bits_needed(data);
            if (b > r->data_bits) r->data_bits = b;
            r->data += 1 + zeroes;
            r->last_zeroes = zeroes;
            zeroes = 0;
        }
    }
    unsigned f = orc_ffs((uint32_t)r->data);                                     /* Block.cpp:231 */
    if (f > r->data_bits) r->data_bits = f;
}
/* Block.cpp:371-413.  Returns the number of coefficients written; *len_field = value of the length field. */
static int orc_stream_block(const orc_tab *t, const orc_rle *r, int use_rle, orc_bw *w, int *len_field) {
    const int NN = t->N * t->N;
    const unsigned bit_len = r->data_bits;
    int length = r->data;
    bw_put(w, 4, bit_len);                                                       /* only the low 4 bits survive */
    if (use_rle) {
        if (length == NN && r->last_zeroes) length -= r->last_zeroes + 1;        /* Block.cpp:388-390 */
        bw_put(w, bit_len, (uint32_t)length);
    } else {
        length = NN;
    }
    if (len_field) *len_field = length;
    /* the RLE list replayed with explicit zeroes == the first `length` zigzag coefficients; without RLE the
       tail is padded with zeroes up to N*N (Block.cpp:401-412) */
    for (int k = 0; k < length; k++) bw_put(w, bit_len, (uint32_t)(int32_t)r->zzc[k]);
    return length;
}
/* Block.cpp:441-472 */
static void orc_load_block(const orc_tab *t, orc_br *rd, int use_rle, double *c) {
    const int NN = t->N * t->N;
    unsigned bit_len = br_get(rd, 4);
    unsigned length = use_rle ? br_get(rd, bit_len) : (unsigned)NN;
    for (int i = 0; i < NN; i++) c[i] = 0.0;
    for (unsigned i = 0; i < length; i++) {
        int16_t v = orc_shift_signed(br_get(rd, bit_len), bit_len);
        if (i < (unsigned)NN) c[t->zz[i]] = (double)v;     /* the reference indexes out of bounds for length > N*N */
    }
}

/* ------------------------------------------------------------------------------------------------
 * Header: ImageEncoder.cpp:84-94, MatrixReader.cpp:144-158,181-190, VideoEncoder.cpp:60-73
 * ---------------------------------------------------------------------------------------------- */
static void orc_write_header(orc_bw *w, int N, const uint16_t *quant, int use_rle, int W, int H, int lead_bit) {
    if (lead_bit) bw_put(w, 1, 0);                                               /* '0': no Huffman (ImageEncoder.cpp:84-86) */
    unsigned qb = 0;
    for (int i = 0; i < N * N; i++) { unsigned f = orc_ffs(quant[i]); if (f > qb) qb = f; }
    bw_put(w, 5, qb);
    for (int i = 0; i < N * N; i++) bw_put(w, qb, quant[i]);
    bw_put(w, 1, use_rle ? 1u : 0u);
    bw_put(w, 15, (uint32_t)W);
    bw_put(w, 15, (uint32_t)H);
}
static void orc_read_header(orc_br *r, int N, uint16_t *quant, int *use_rle, int *W, int *H) {  /* MatrixReader.cpp:45-57, ImageBase.cpp:122-128 */
    unsigned qb = br_get(r, 5);
    for (int i = 0; i < N * N; i++) quant[i] = (uint16_t)br_get(r, qb);
    *use_rle = (int)br_get(r, 1);
    *W = (int)br_get(r, 15);
    *H = (int)br_get(r, 15);
}

/* ------------------------------------------------------------------------------------------------
 * Image encode: ImageEncoder.cpp:52-175 (without the Huffman stage), block order ImageBase.cpp:187-199.
 * Returns the number of bits written (the file is orc_round_to_byte(bits) bytes), or -1.
 * Optional per-stage outputs (may be NULL): coef_zz[nblocks*N*N] int16 zigzag-ordered quantised coefficients,
 * bitlen[nblocks], lenfield[nblocks].
 * ---------------------------------------------------------------------------------------------- */
long long orc_image_encode(const uint8_t *raw, int W, int H, int N, const uint16_t *quant, int use_rle,
                           int lead_bit, uint8_t *out, size_t out_cap,
                           int16_t *coef_zz, uint8_t *bitlen, uint8_t *lenfield) {
    if ((N != 4 && N != 8) || W % N || H % N || W <= 0 || H <= 0) return -1;
    orc_tab t; orc_tab_init(&t, N);
    double m[ORC_MAXNN];
    for (int i = 0; i < N * N; i++) m[i] = (double)quant[i];                     /* MatrixReader.cpp:128 */
    memset(out, 0, out_cap);
    orc_bw w = { out, out_cap, 0, 0 };
    orc_write_header(&w, N, quant, use_rle, W, H, lead_bit);
    const int bx = W / N, by = H / N;
    size_t blk = 0;
    for (int yb = 0; yb < by; yb++)
        for (int xb = 0; xb < bx; xb++, blk++) {
            double x[ORC_MAXNN], q[ORC_MAXNN];
            for (int y = 0; y < N; y++)
                for (int xx = 0; xx < N; xx++)
                    x[y * N + xx] = (double)raw[(size_t)(yb * N + y) * W + xb * N + xx];   /* Block.cpp:52-54 */
            orc_fdct_quant(&t, x, m, q);
            orc_rle r; orc_make_rle(&t, q, &r);
            int lf = 0;
            orc_stream_block(&t, &r, use_rle, &w, &lf);
            if (coef_zz) memcpy(coef_zz + blk * N * N, r.zzc, sizeof(int16_t) * N * N);
            if (bitlen) bitlen[blk] = (uint8_t)r.data_bits;
            if (lenfield) lenfield[blk] = (uint8_t)lf;
        }
    return w.overflow ? -2 : (long long)w.pos;
}

/* Image decode of a plain (already Huffman-decoded) stream starting at bit `start_bit`:
 * ImageBase.cpp:118-128 + ImageDecoder.cpp:55-122.  Returns 0, or <0 on error; out must hold W*H bytes. */
int orc_image_decode(const uint8_t *enc, size_t enc_bytes, size_t start_bit, int N,
                     uint8_t *out, size_t out_cap, int *W_out, int *H_out, int16_t *coef_zz) {
    if (N != 4 && N != 8) return -1;
    orc_tab t; orc_tab_init(&t, N);
    orc_br r = { enc, enc_bytes, start_bit };
    uint16_t quant[ORC_MAXNN]; int use_rle, W, H;
    orc_read_header(&r, N, quant, &use_rle, &W, &H);
    *W_out = W; *H_out = H;
    if (W % N || H % N) return -2;
    if ((size_t)W * H > out_cap) return -3;
    double m[ORC_MAXNN];
    for (int i = 0; i < N * N; i++) m[i] = (double)quant[i];
    const int bx = W / N, by = H / N;
    size_t blk = 0;
    for (int yb = 0; yb < by; yb++)
        for (int xb = 0; xb < bx; xb++, blk++) {
            double c[ORC_MAXNN], px[ORC_MAXNN];
            orc_load_block(&t, &r, use_rle, c);
            if (coef_zz) for (int k = 0; k < N * N; k++) coef_zz[blk * N * N + k] = (int16_t)c[t.zz[k]];
            orc_dequant_idct(&t, c, m, px);
            for (int y = 0; y < N; y++)
                for (int xx = 0; xx < N; xx++)
                    out[(size_t)(yb * N + y) * W + xb * N + xx] = orc_clamp_u8(px[y * N + xx]);   /* Block.cpp:99-107 */
        }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Video.  BlockSize is 4 (ImageBase.cpp:266-306 only works for 4), MacroBlock 16 (Block.hpp:14).
 * ---------------------------------------------------------------------------------------------- */
static const int MER_SIGNS[9][2] = {                                             /* algo.cpp:90-100 (x, y) */
    {0, 0}, {+1, 0}, {+1, +1}, {0, +1}, {-1, +1}, {-1, 0}, {-1, -1}, {0, -1}, {+1, -1}
};
static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

static size_t orc_sad16(const uint8_t *cur, const uint8_t *ref, int W) {         /* Block.cpp:241-254 */
    size_t d = 0;
    for (int y = 0; y < ORC_MB; y++)
        for (int x = 0; x < ORC_MB; x++)
            d += (size_t)abs((int)cur[y * W + x] - (int)ref[y * W + x]);
    return d;
}
/* Block.cpp:267-339 + algo.cpp:119-139 + ImageBase.cpp:243-264.
 * mv = unclamped offset; (cx,cy) = clamped pixel coordinate of the block the residual is taken from. */
static void orc_motion_search(const uint8_t *cur, const uint8_t *ref, int W, int H, int mbx, int mby, int merange,
                              int *mvx, int *mvy, int *cx, int *cy) {
    int best_x = 0, best_y = 0;
    int bcx = clampi(0, 0, W - ORC_MB), bcy = clampi(0, 0, H - ORC_MB);          /* Block.cpp:273: block at pixel (0,0) */
    size_t best_d = (size_t)-1;
    for (int step = merange / 2; step > 0; step /= 2) {                          /* algo.cpp:129,138 */
        int have = 0, nx = 0, ny = 0, ncx = 0, ncy = 0;
        size_t nd = best_d;
        for (int p = 0; p < 9; p++) {
            int ox = best_x + MER_SIGNS[p][0] * step, oy = best_y + MER_SIGNS[p][1] * step;
            int px = clampi((int16_t)(ox + mbx), 0, W - ORC_MB), py = clampi((int16_t)(oy + mby), 0, H - ORC_MB);
            if (p > 0 && px == mbx && py == mby) continue;                        /* Block.cpp:297-301 */
            size_t d = orc_sad16(cur + (size_t)mby * W + mbx, ref + (size_t)py * W + px, W);
            if (d <= nd) { have = 1; nx = ox; ny = oy; nd = d; ncx = px; ncy = py; }   /* Block.cpp:306 */
        }
        if (!have) break;                                                         /* Block.cpp:318-321 (dead) */
        best_x = nx; best_y = ny; best_d = nd; bcx = ncx; bcy = ncy;
    }
    *mvx = best_x; *mvy = best_y; *cx = bcx; *cy = bcy;
}

/* VideoEncoder.cpp:22-111 + Frame.cpp:129-247 (without the Huffman stage).  yuv is modified in place exactly as the
 * reference modifies its raw buffer (P-frame reconstruction), so on return it holds the encoder-side reconstruction.
 * mvecs (optional): int16 pairs (x,y) per macroblock per frame (zero for I-frames).  Returns bits or <0. */
long long orc_video_encode(uint8_t *yuv, size_t yuv_bytes, int W, int H, const uint16_t *quant, int use_rle,
                           int gop, int merange, int lead_bit, uint8_t *out, size_t out_cap, int16_t *mvecs) {
    const int N = 4;
    if (W % ORC_MB || H % ORC_MB || W <= 0 || H <= 0) return -1;
    orc_tab t; orc_tab_init(&t, N);
    double m[16];
    for (int i = 0; i < 16; i++) m[i] = (double)quant[i];
    if (gop < 1) gop = 1;                                                         /* VideoBase.cpp:34 */
    const size_t ysz = (size_t)W * H, fsz = ysz + ysz / 2;
    const size_t frames = yuv_bytes / fsz;                                        /* VideoBase.cpp:39-40 */
    const unsigned mvbits = orc_bits_needed((int16_t)merange);                    /* VideoBase.cpp:42 */
    memset(out, 0, out_cap);
    orc_bw w = { out, out_cap, 0, 0 };
    orc_write_header(&w, N, quant, use_rle, W, H, lead_bit);
    bw_put(&w, 15, (uint32_t)frames); bw_put(&w, 15, (uint32_t)gop); bw_put(&w, 15, (uint32_t)merange);
    const int bx = W / N, by = H / N, mx = W / ORC_MB, my = H / ORC_MB;
    orc_rle *rles = (orc_rle *)malloc(sizeof(orc_rle) * (size_t)bx * by);
    double *rec = (double *)malloc(sizeof(double) * ysz);
    int16_t *mv = (int16_t *)malloc(sizeof(int16_t) * 2 * (size_t)mx * my);
    int *cc = (int *)malloc(sizeof(int) * 2 * (size_t)mx * my);
    for (size_t f = 0; f < frames; f++) {
        uint8_t *cur = yuv + f * fsz;
        if (f % (size_t)gop == 0) {                                               /* VideoBase.hpp:32; Frame.cpp:130-159 */
            for (int yb = 0; yb < by; yb++)
                for (int xb = 0; xb < bx; xb++) {
                    double x[16], q[16];
                    for (int y = 0; y < 4; y++) for (int xx = 0; xx < 4; xx++) x[y * 4 + xx] = (double)cur[(size_t)(yb * 4 + y) * W + xb * 4 + xx];
                    orc_fdct_quant(&t, x, m, q);
                    orc_rle r; orc_make_rle(&t, q, &r);
                    orc_stream_block(&t, &r, use_rle, &w, NULL);
                }
            if (mvecs) memset(mvecs + f * 2 * (size_t)mx * my, 0, sizeof(int16_t) * 2 * (size_t)mx * my);
            continue;
        }
        const uint8_t *ref = yuv + (f - 1) * fsz;                                 /* VideoBase.cpp:105-118 */
        /* pass 1: search + residual transform for every macroblock (Frame.cpp:181-192 / 205-228) */
        for (int mby = 0; mby < my; mby++)
            for (int mbx = 0; mbx < mx; mbx++) {
                int idx = mby * mx + mbx, vx, vy, cx, cy;
                orc_motion_search(cur, ref, W, H, mbx * ORC_MB, mby * ORC_MB, merange, &vx, &vy, &cx, &cy);
                mv[2 * idx] = (int16_t)vx; mv[2 * idx + 1] = (int16_t)vy;
                /* block copied into cur afterwards: getCoordAfterMotion + clamp (Frame.cpp:218-223) */
                cc[2 * idx] = clampi((int16_t)(mbx * ORC_MB + vx), 0, W - ORC_MB);
                cc[2 * idx + 1] = clampi((int16_t)(mby * ORC_MB + vy), 0, H - ORC_MB);
                for (int sy = 0; sy < 4; sy++)
                    for (int sx = 0; sx < 4; sx++) {                              /* ImageBase.cpp:281-305 */
                        double x[16], q[16], px[16];
                        for (int y = 0; y < 4; y++)
                            for (int xx = 0; xx < 4; xx++) {
                                int gy = mby * ORC_MB + sy * 4 + y, gx = mbx * ORC_MB + sx * 4 + xx;
                                int ry = cy + sy * 4 + y, rx = cx + sx * 4 + xx;
                                x[y * 4 + xx] = (double)cur[(size_t)gy * W + gx] - (double)ref[(size_t)ry * W + rx];   /* Block.cpp:262 */
                            }
                        orc_fdct_quant(&t, x, m, q);                              /* -128 applied to the residual too */
                        size_t bidx = (size_t)(mby * 4 + sy) * bx + (mbx * 4 + sx);
                        orc_make_rle(&t, q, &rles[bidx]);
                        orc_dequant_idct(&t, q, m, px);                           /* ImageBase.cpp:303 */
                        for (int y = 0; y < 4; y++)
                            for (int xx = 0; xx < 4; xx++)
                                rec[(size_t)(mby * ORC_MB + sy * 4 + y) * W + mbx * ORC_MB + sx * 4 + xx] = px[y * 4 + xx];
                    }
            }
        /* pass 2: cur[MB] = ref[clamped MB+mv]  (Frame.cpp:218-225), mvecs to the stream (Block.cpp:415-423) */
        for (int mby = 0; mby < my; mby++)
            for (int mbx = 0; mbx < mx; mbx++) {
                int idx = mby * mx + mbx;
                for (int y = 0; y < ORC_MB; y++)
                    memcpy(cur + (size_t)(mby * ORC_MB + y) * W + mbx * ORC_MB,
                           ref + (size_t)(cc[2 * idx + 1] + y) * W + cc[2 * idx], ORC_MB);
                bw_put(&w, mvbits, (uint32_t)(int32_t)mv[2 * idx]);
                bw_put(&w, mvbits, (uint32_t)(int32_t)mv[2 * idx + 1]);
            }
        if (mvecs) memcpy(mvecs + f * 2 * (size_t)mx * my, mv, sizeof(int16_t) * 2 * (size_t)mx * my);
        /* pass 3: reconstruct + stream every micro block (Frame.cpp:234-242, Block.cpp:110-119) */
        for (int yb = 0; yb < by; yb++)
            for (int xb = 0; xb < bx; xb++) {
                for (int y = 0; y < 4; y++)
                    for (int xx = 0; xx < 4; xx++) {
                        size_t o = (size_t)(yb * 4 + y) * W + xb * 4 + xx;
                        cur[o] = orc_clamp_u8((double)cur[o] + rec[o]);
                    }
                orc_stream_block(&t, &rles[(size_t)yb * bx + xb], use_rle, &w, NULL);
            }
    }
    free(rles); free(rec); free(mv); free(cc);
    return w.overflow ? -2 : (long long)w.pos;
}

/* VideoBase.cpp:45-85 + VideoDecoder.cpp:33-62 + Frame.cpp:47-127 + Block.cpp:481-496, plain stream at start_bit.
 * out receives frames * (W*H*3/2) bytes (Y then 0x80 fill).  Returns 0 or <0. */
int orc_video_decode(const uint8_t *enc, size_t enc_bytes, size_t start_bit, int motioncomp,
                     uint8_t *out, size_t out_cap, int *W_out, int *H_out, int *frames_out, int *gop_out, int *mer_out) {
    const int N = 4;
    orc_tab t; orc_tab_init(&t, N);
    orc_br r = { enc, enc_bytes, start_bit };
    uint16_t quant[16]; int use_rle, W, H;
    orc_read_header(&r, N, quant, &use_rle, &W, &H);
    int frames = (int)br_get(&r, 15), gop = (int)br_get(&r, 15), merange = (int)br_get(&r, 15);
    *W_out = W; *H_out = H; *frames_out = frames; *gop_out = gop; *mer_out = merange;
    if (W % ORC_MB || H % ORC_MB || W <= 0 || H <= 0 || gop < 1) return -2;
    const size_t ysz = (size_t)W * H, fsz = ysz + ysz / 2;
    if (fsz * (size_t)frames > out_cap) return -3;
    const unsigned mvbits = orc_bits_needed((int16_t)merange);
    double m[16];
    for (int i = 0; i < 16; i++) m[i] = (double)quant[i];
    const int bx = W / N, by = H / N, mx = W / ORC_MB, my = H / ORC_MB;
    for (int f = 0; f < frames; f++) {
        uint8_t *cur = out + (size_t)f * fsz;
        const int is_i = (f % gop == 0);
        if (!is_i) {
            const uint8_t *ref = out + (size_t)(f - 1) * fsz;
            for (int mby = 0; mby < my; mby++)
                for (int mbx = 0; mbx < mx; mbx++) {                               /* Block.cpp:481-496 */
                    int vx = orc_shift_signed(br_get(&r, mvbits), mvbits), vy = orc_shift_signed(br_get(&r, mvbits), mvbits);
                    int cx = clampi((int16_t)(mbx * ORC_MB + vx), 0, W - ORC_MB), cy = clampi((int16_t)(mby * ORC_MB + vy), 0, H - ORC_MB);
                    for (int y = 0; y < ORC_MB; y++)
                        memcpy(cur + (size_t)(mby * ORC_MB + y) * W + mbx * ORC_MB, ref + (size_t)(cy + y) * W + cx, ORC_MB);
                }
        }
        for (int yb = 0; yb < by; yb++)
            for (int xb = 0; xb < bx; xb++) {
                double c[16], px[16];
                orc_load_block(&t, &r, use_rle, c);
                if (!is_i && !motioncomp) continue;                                /* Frame.cpp:110-116 */
                orc_dequant_idct(&t, c, m, px);
                for (int y = 0; y < 4; y++)
                    for (int xx = 0; xx < 4; xx++) {
                        size_t o = (size_t)(yb * 4 + y) * W + xb * 4 + xx;
                        cur[o] = is_i ? orc_clamp_u8(px[y * 4 + xx]) : orc_clamp_u8((double)cur[o] + px[y * 4 + xx]);
                    }
            }
        memset(cur + ysz, 0x80, ysz / 2);                                          /* Frame.cpp:122-124 */
    }
    return 0;
}

/* header sizes, for tests */
int orc_header_bits(int N, const uint16_t *quant, int lead_bit, int video) {
    unsigned qb = 0;
    for (int i = 0; i < N * N; i++) { unsigned f = orc_ffs(quant[i]); if (f > qb) qb = f; }
    return (lead_bit ? 1 : 0) + 5 + (int)qb * N * N + 1 + 15 + 15 + (video ? 45 : 0);
}
