// CPU check of the encode-variant arithmetic (imageencoder_b200/csrc/transform_fast.cuh, namespace ie::lean).
// Test infrastructure, not product code.  Compiled with nvcc as HOST code: every device-only instruction used by the variants
// (FFMA2/FADD2/FMUL2, PRMT, VIMNMX3.S16x2, the predicated OR) has a bit-identical host shim in the header, so this runs the
// same C++ the kernels run, lane by lane, against a transcription of the default path of encode_tiles_kernel
// (encode_image.cu, "phase 1": quantise loop) on millions of blocks:
//   * variant 1 (scalar transform + lean::quantise_block_packed): staged zigzag coefficients, guard-band mask, max
//     bits_needed, segment non-zero flags
//   * variant 2 (lean::fdct2d_packed + quantise_block_packed): additionally the packed transform, bit for bit
//   * decode variant 1 (lean::idct2d_packed + lean::pixel_stage): inverse transform bit for bit, output words, `unsure` mask
//   * exact-queue variant "fast64" (lean::row_dot64 + lean::decide64): every coefficient it decides equals the reference's
//     exact-order chain (algo.cpp:309-331, Block.cpp:152), the quotients differ by < 1e-9, true ties are left undecided
// and the integer DC rounding against round_half_away(S / (4 Q00)) in exact integer arithmetic.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <random>
#include <vector>

#include "../../imageencoder_b200/csrc/transform_fast.cuh"

using namespace ie;

static const unsigned char kZigzagInv4[16] = {0, 1, 5, 6, 2, 4, 7, 12, 3, 8, 11, 13, 9, 10, 14, 15};
static const unsigned char kZigzagInv8[64] = {
    0,  1,  5,  6,  14, 15, 27, 28, 2,  4,  7,  13, 16, 26, 29, 42, 3,  8,  12, 17, 25, 30, 41, 43, 9,  11, 18, 24, 31, 40, 44, 53,
    10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38, 46, 51, 55, 60, 21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63};

// core.cu: make_fast_quant
static void make_fq(FastQuant &fq, const uint16_t *quant, int N, double max_abs_sample) {
    memset(&fq, 0, sizeof fq);
    for (int u = 0; u < N; u++)
        for (int v = 0; v < N; v++) {
            const double cu = (u == 0) ? 0.5 : M_SQRT1_2, cv = (v == 0) ? 0.5 : M_SQRT1_2;
            const double k = cu * cv / (double)quant[u * N + v];
            const double delta = 1.25 * (768.0 + 64.0) * ldexp(1.0, -24) * max_abs_sample * k + 1e-6;
            fq.k[u * N + v] = (float)k;
            double thr = 0.5 - delta;
            if (thr < 0.0) thr = 0.0;
            fq.thr[u * N + v] = nextafterf((float)thr, 0.0f);
        }
    for (int i = N * N; i < kMaxNN; i++) { fq.k[i] = 0.f; fq.thr[i] = 1.f; }
}

static int bitlen(unsigned v) { int n = 0; while (v) { n++; v >>= 1; } return n; }

struct Ref {
    short cf[64];
    unsigned long long near;
    unsigned orbits;
    unsigned orseg[8];
};

// transcription of the default path (encode_image.cu, quantise loop of encode_tiles_kernel<N, BPL, PF, FAST=true, VAR=0>)
template <int N>
static void ref_quantise(const float *y, const FastQuant &fq, int dc_den2, float dc_rcp, Ref &r) {
    constexpr int NN = N * N;
    constexpr float kMagic = 12582912.0f;
    constexpr int kMagicBits = 0x4B400000;
    r.near = 0; r.orbits = 0;
    for (int s = 0; s < 8; s++) r.orseg[s] = 0;
    for (int uv = 0; uv < NN; uv++) {
        int q;
        if (uv == 0) {
            const int S = (int)y[0];
            const int n = abs(S);
            const int D2 = dc_den2;
            const int num = 2 * n + (D2 >> 1);
            int qq = (int)((float)num * dc_rcp);
            const int rem = num - qq * D2;
            qq += (rem >= D2) ? 1 : 0;
            qq -= (rem < 0) ? 1 : 0;
            q = (S < 0) ? -qq : qq;
        } else {
            const float rr = fmaf(y[uv], fq.k[uv], kMagic);
            const float rf = rr - kMagic;
            const float d = fmaf(y[uv], fq.k[uv], -rf);
            if (fabsf(d) >= fq.thr[uv]) r.near |= 1ull << uv;
            unsigned b; memcpy(&b, &rr, 4);
            q = (int)b - kMagicBits;
        }
        const int k = (N == 8) ? kZigzagInv8[uv] : kZigzagInv4[uv];
        r.cf[k] = (short)q;
        r.orseg[k >> 3] |= (unsigned)q;
        r.orbits |= (unsigned)(q ^ (q >> 31));
    }
}

static long long g_fail = 0;
#define CHECK(cond, ...) do { if (!(cond)) { if (g_fail < 20) { printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } g_fail++; } } while (0)

template <int N>
static void check_block(const uint8_t *px, const uint16_t *quant, const FastQuant &fq, long long id) {
    constexpr int NN = N * N;
    float x[NN], y[NN];
    for (int i = 0; i < NN; i++) x[i] = (float)px[i] - 128.0f;
    memcpy(y, x, sizeof y);
    fdct2d_fast<N>(y);
    const int dc_den2 = 8 * (int)quant[0];
    const float dc_rcp = 1.0f / (float)dc_den2;
    Ref r;
    ref_quantise<N>(y, fq, dc_den2, dc_rcp, r);

    // DC against exact integer arithmetic: round_half_away(S / (4 Q00))
    {
        long long S = 0;
        for (int i = 0; i < NN; i++) S += (int)px[i] - 128;
        const long long D = 4LL * quant[0], a = S < 0 ? -S : S;
        long long qq = (2 * a + D) / (2 * D);
        if (S < 0) qq = -qq;
        CHECK(r.cf[0] == (short)qq, "block %lld: DC %d, exact integer rounding gives %lld", id, r.cf[0], qq);
    }

    // ---- variant 1
    {
        float2 y2[NN / 2];
        for (int i = 0; i < NN / 2; i++) y2[i] = make_float2(y[2 * i], y[2 * i + 1]);
        unsigned cfw[NN / 2], nlo, nhi, orseg[8] = {0}, orbits;
        lean::quantise_block_packed<N>(y2, fq, dc_den2, dc_rcp, cfw, nlo, nhi, orseg, orbits);
        const unsigned long long near = ((unsigned long long)nhi << 32) | nlo;
        CHECK(near == r.near, "block %lld v1: near mask %llx != %llx", id, near, r.near);
        for (int k = 0; k < NN; k++) {
            const short got = (short)((cfw[k >> 1] >> (16 * (k & 1))) & 0xffffu);
            CHECK(got == r.cf[k], "block %lld v1: zigzag %d: %d != %d", id, k, got, r.cf[k]);
        }
        CHECK(bitlen(orbits) == bitlen(r.orbits), "block %lld v1: orbits %x vs %x", id, orbits, r.orbits);
        for (int s = 0; s < NN / 8; s++) CHECK((orseg[s] != 0) == (r.orseg[s] != 0), "block %lld v1: segment %d flag", id, s);
    }
    // ---- variant 2
    {
        float2 x2[NN / 2], y2[NN / 2];
        for (int rr = 0; rr < N / 2; rr++)
            for (int k = 0; k < N; k++) x2[rr * N + k] = make_float2(x[(2 * rr) * N + k], x[(2 * rr + 1) * N + k]);
        lean::fdct2d_packed<N>(x2, y2);
        for (int u = 0; u < N; u++)
            for (int c = 0; c < N / 2; c++) {
                const float2 v = y2[u * (N / 2) + c];
                CHECK(memcmp(&v.x, &y[u * N + 2 * c], 4) == 0 && memcmp(&v.y, &y[u * N + 2 * c + 1], 4) == 0,
                      "block %lld v2: transform output (%d,%d) differs: %a %a vs %a %a", id, u, 2 * c, v.x, v.y, y[u * N + 2 * c], y[u * N + 2 * c + 1]);
            }
        unsigned cfw[NN / 2], nlo, nhi, orseg[8] = {0}, orbits;
        lean::quantise_block_packed<N>(y2, fq, dc_den2, dc_rcp, cfw, nlo, nhi, orseg, orbits);
        const unsigned long long near = ((unsigned long long)nhi << 32) | nlo;
        CHECK(near == r.near, "block %lld v2: near mask %llx != %llx", id, near, r.near);
        for (int k = 0; k < NN; k++) {
            const short got = (short)((cfw[k >> 1] >> (16 * (k & 1))) & 0xffffu);
            CHECK(got == r.cf[k], "block %lld v2: zigzag %d: %d != %d", id, k, got, r.cf[k]);
        }
        CHECK(bitlen(orbits) == bitlen(r.orbits), "block %lld v2: orbits %x vs %x", id, orbits, r.orbits);
        for (int s = 0; s < NN / 8; s++) CHECK((orseg[s] != 0) == (r.orseg[s] != 0), "block %lld v2: segment %d flag", id, s);
    }
}

template <int N>
static long long run(long long nblocks, unsigned seed) {
    constexpr int NN = N * N;
    std::mt19937 rng(seed);
    static const unsigned char zz4[16] = {0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15};
    unsigned char zz[64];
    for (int uv = 0; uv < NN; uv++) zz[(N == 8) ? kZigzagInv8[uv] : kZigzagInv4[uv]] = (unsigned char)uv;
    if (N == 4) for (int i = 0; i < 16; i++) if (zz[i] != zz4[i]) { printf("zigzag table mismatch\n"); return 1; }
    for (int i = 0; i < NN; i++) {
        const int t = (N == 8) ? lean::ZZ8::t[i] : lean::ZZ4::t[i];
        if (t != zz[i]) { printf("lean::ZZ%d::t[%d] = %d, expected %d\n", N, i, t, zz[i]); return 1; }
    }
    long long done = 0;
    for (int qm = 0; qm < 6; qm++) {
        uint16_t quant[64];
        for (int i = 0; i < NN; i++) {
            const int u = i / N, v = i % N;
            switch (qm) {
                case 0: quant[i] = 1; break;                                        // largest coefficients
                case 1: quant[i] = (uint16_t)(2 + 3 * (u + v)); break;              // JPEG-like ramp
                case 2: quant[i] = (uint16_t)(16 + 11 * u * v + 7 * (u + v)); break;
                case 3: quant[i] = 255; break;
                case 4: quant[i] = (uint16_t)(1 + (rng() % 64)); break;
                default: quant[i] = (uint16_t)((i == 0) ? 3 : 2); break;            // odd DC divisor, many half-way cases
            }
        }
        FastQuant fq;
        make_fq(fq, quant, N, 128.0);
        for (long long b = 0; b < nblocks; b++) {
            uint8_t px[NN];
            const int mode = (int)(rng() % 6);
            const int base = (int)(rng() % 256), amp = (int)(rng() % 64);
            const int gx = (int)(rng() % 17) - 8, gy = (int)(rng() % 17) - 8;
            for (int i = 0; i < NN; i++) {
                const int yy = i / N, xx = i % N;
                int v;
                switch (mode) {
                    case 0: v = (int)(rng() % 256); break;                                         // noise
                    case 1: v = base + gx * xx + gy * yy + (int)(rng() % (amp + 1)) - amp / 2; break;   // gradient + noise
                    case 2: v = base + ((int)(rng() % 3) - 1); break;                              // nearly flat
                    case 3: v = ((xx + yy) & 1) ? 255 : 0; break;                                  // checkerboard (extreme AC)
                    case 4: v = (rng() & 1) ? 255 : 0; break;                                      // random extremes
                    default: v = base; break;                                                      // constant (all AC zero)
                }
                px[i] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
            }
            check_block<N>(px, quant, fq, done);
            done++;
        }
    }
    return 0;
}


// ---------------------------------------------------------------------------------------------------------------------
// decode variant 1: lean::idct2d_packed + lean::pixel_stage against a transcription of decode_blocks_fast_kernel<N, false>
// (decode_image.cu: "fast inverse transform" and the pixel loop)
// ---------------------------------------------------------------------------------------------------------------------
template <int N>
static void check_decode_block(const short *cf /*zigzag order*/, const float *k2 /*raster*/, long long id) {
    constexpr int NN = N * N;
    // ---- transcription of the default path
    float x[NN];
    float S = 0.f;
    for (int uv = 0; uv < NN; uv++) {
        const int k = (N == 8) ? kZigzagInv8[uv] : kZigzagInv4[uv];
        const int c = cf[k];
        const float d = (float)c * k2[uv];
        x[uv] = d;
        S += fabsf(d);
    }
    float xin[NN];
    memcpy(xin, x, sizeof x);
    idct2d_fast<N>(x);
    const float delta = (32.f * S + 2.f * (S + 383.f)) * 5.9604645e-8f * 1.0001f + 2e-6f;
    const float hi_thr = (delta < 0.49f) ? 0.5f - delta : 0.f;
    unsigned outw[N * (N / 4)];
    unsigned long long unsure = 0;
    for (int y = 0; y < N; y++)
        for (int q4 = 0; q4 < N / 4; q4++) {
            unsigned fl[4];
            for (int b = 0; b < 4; b++) {
                const int ij = y * N + q4 * 4 + b;
                const float v = x[ij] + 128.f;
                const float u = fminf(fmaxf(v, 0.5f), 255.5f);
                const float fm = 8388608.0f + floorf(u);                       // __fadd_rd(u, 2^23) for u in [0.5, 255.5]
                const float frac = u - (fm - 8388608.0f);
                if (fabsf(frac - 0.5f) >= hi_thr) unsure |= 1ull << ij;
                memcpy(&fl[b], &fm, 4);
            }
            outw[y * (N / 4) + q4] = (fl[0] & 0xffu) | ((fl[1] & 0xffu) << 8) | ((fl[2] & 0xffu) << 16) | ((fl[3] & 0xffu) << 24);
        }
    // ---- variant
    float2 x2[NN / 2], p2[NN / 2];
    for (int r = 0; r < N / 2; r++)
        for (int v = 0; v < N; v++) x2[r * N + v] = make_float2(xin[(2 * r) * N + v], xin[(2 * r + 1) * N + v]);
    lean::idct2d_packed<N>(x2, p2);
    for (int i = 0; i < N; i++)
        for (int c = 0; c < N / 2; c++) {
            const float2 v = p2[i * (N / 2) + c];
            CHECK(memcmp(&v.x, &x[i * N + 2 * c], 4) == 0 && memcmp(&v.y, &x[i * N + 2 * c + 1], 4) == 0,
                  "decode block %lld: inverse transform output (%d,%d) differs: %a %a vs %a %a", id, i, 2 * c, v.x, v.y, x[i * N + 2 * c],
                  x[i * N + 2 * c + 1]);
        }
    unsigned ow[N * (N / 4)], ulo, uhi;
    lean::pixel_stage<N>(p2, hi_thr, ow, ulo, uhi);
    const unsigned long long un = ((unsigned long long)uhi << 32) | ulo;
    CHECK(un == unsure, "decode block %lld: unsure mask %llx != %llx", id, un, unsure);
    for (int i = 0; i < N * (N / 4); i++) CHECK(ow[i] == outw[i], "decode block %lld: output word %d: %08x != %08x", id, i, ow[i], outw[i]);
}

template <int N>
static void run_decode(long long nblocks, unsigned seed) {
    constexpr int NN = N * N;
    std::mt19937 rng(seed);
    for (int qm = 0; qm < 4; qm++) {
        float k2[64];
        for (int i = 0; i < NN; i++) {
            const int u = i / N, v = i % N;
            const double cu = (u == 0) ? 0.5 : M_SQRT1_2, cv = (v == 0) ? 0.5 : M_SQRT1_2;
            const int Q = (qm == 0) ? 1 : (qm == 1) ? 2 + 3 * (u + v) : (qm == 2) ? 16 + 11 * u * v + 7 * (u + v) : 255;
            k2[i] = (float)((double)Q * (cu * cv));
        }
        for (long long b = 0; b < nblocks; b++) {
            short cf[64] = {0};
            const int mode = (int)(rng() % 5);
            const int len = (mode == 0) ? 0 : (mode == 1) ? 1 : (int)(rng() % (NN + 1));
            const int amp = (mode == 4) ? 32767 : (mode == 3) ? 600 : 24;
            for (int k = 0; k < len; k++) {
                if (mode == 2 && (rng() % 3)) continue;                           // sparse
                cf[k] = (short)((int)(rng() % (2 * amp + 1)) - amp);
            }
            if (len && (rng() & 1)) cf[0] = (short)((int)(rng() % 1200) - 600);   // a DC that lands the pixels inside [0, 255]
            check_decode_block<N>(cf, k2, b);
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// exact-queue variant "fast64": lean::row_dot64 + lean::decide64 against the reference's exact-order chain
// (transcription of exact_coefficient, encode_image.cu; this file is compiled without FMA contraction)
// ---------------------------------------------------------------------------------------------------------------------
static long long g_f64_decided = 0, g_f64_undecided = 0;
static double g_f64_maxdev = 0.0;

template <int N>
static void run_fast64(long long nblocks, unsigned seed) {
    constexpr int NN = N * N;
    std::mt19937 rng(seed);
    double cs[NN], cc[NN];
    const double f = M_PI_2 / double(N);
    for (int i = 0; i < N; i++)
        for (int u = 0; u < N; u++) cs[i * N + u] = std::cos(double(2.0 * i + 1.0) * double(u) * f);
    for (int u = 0; u < N; u++)
        for (int v = 0; v < N; v++) cc[u * N + v] = ((u == 0) ? 0.5 : M_SQRT1_2) * ((v == 0) ? 0.5 : M_SQRT1_2);
    for (long long blk = 0; blk < nblocks; blk++) {
        int x[NN];
        const int mode = (int)(rng() % 7);
        const int base = (int)(rng() % 256) - 128;
        for (int i = 0; i < NN; i++) {
            switch (mode) {
            case 0: x[i] = (int)(rng() % 256) - 128; break;                               // pixels - 128
            case 1: x[i] = (int)(rng() % 511) - 255 - 128; break;                         // P-frame residual - 128
            case 2: x[i] = (rng() & 1) ? 8 : -8; break;                                   // two levels: ties at the rational positions
            case 3: x[i] = (int)(rng() % 4) * 8 - 16; break;                              // four levels
            case 4: x[i] = (rng() & 1) ? 127 : -383; break;                               // saturated
            case 5: x[i] = base; break;                                                   // flat
            default: x[i] = base + (int)(rng() % 5) - 2; break;                           // nearly flat
            }
        }
        const int mq = (int)(rng() % 4);
        for (int uv = 0; uv < NN; uv++) {
            const int u = uv / N, v = uv % N;
            const double m = (mq == 0) ? 1.0 : (mq == 1) ? 2.0 : (mq == 2) ? 16.0 : (double)(1 + rng() % 255);
            // the reference's chain
            double acc = 0.0;
            for (int i = 0; i < N; i++)
                for (int j = 0; j < N; j++) {
                    volatile double ab = cs[i * N + u] * cs[j * N + v];
                    volatile double t = ab * (double)x[i * N + j];
                    volatile double sum = acc + t;
                    acc = sum;
                }
            volatile double e = acc * cc[uv];
            volatile double tq = e / m;
            const double a0 = fabs(tq);
            double r0 = floor(a0);
            if (a0 - r0 >= 0.5) r0 += 1.0;                                                 // std::round: half away from zero
            const int q_ref = (int)(short)(int)copysign(r0, tq);
            // fast64
            double a[N], b[N];
            for (int y = 0; y < N; y++) { a[y] = cs[y * N + u]; b[y] = cs[y * N + v]; }
            double acc2 = 0.0;
            for (int y = 0; y < N; y++) acc2 = fma(a[y], lean::row_dot64<N>(b, x + y * N), acc2);
            int q = 0x7fffffff;
            const bool decided = lean::decide64(acc2, cc[uv], m, q);
            const double dev = fabs(acc2 * cc[uv] / m - tq);
            if (dev > g_f64_maxdev) g_f64_maxdev = dev;
            CHECK(dev < 1e-9, "fast64 block %lld uv %d: quotient deviates by %g", blk, uv, dev);
            if (decided) {
                g_f64_decided++;
                CHECK(q == q_ref, "fast64 block %lld uv %d (mode %d, m %g): decided %d, reference %d (t = %.17g)", blk, uv, mode, m, q, q_ref,
                      (double)tq);
            } else {
                g_f64_undecided++;
                const double fr = a0 - floor(a0);
                CHECK(fabs(fr - 0.5) < 2e-8, "fast64 block %lld uv %d: undecided although t = %.17g is clear of a boundary", blk, uv, (double)tq);
            }
        }
    }
}

int main(int argc, char **argv) {
    const long long n = (argc > 1) ? atoll(argv[1]) : 200000;
    if (run<8>(n, 12345u) || run<4>(n, 54321u)) return 2;
    run_decode<8>(n, 777u);
    run_decode<4>(n, 888u);
    run_fast64<8>(n / 8 + 1, 4242u);
    run_fast64<4>(n / 2 + 1, 2424u);
    if (g_fail) { printf("lean_check: %lld failures\n", g_fail); return 1; }
    printf("lean_check: ok (%lld encode blocks per block size with 6 quant matrices, encode variants 1 and 2; %lld decode blocks per "
           "block size with 4 matrices, decode variant 1; fast64: %lld coefficients decided, %lld left to the exact chain, max "
           "quotient deviation %.3g)\n", n * 6, n * 4, g_f64_decided, g_f64_undecided, g_f64_maxdev);
    return 0;
}
