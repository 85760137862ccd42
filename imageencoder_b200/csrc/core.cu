// Host-side core: error plumbing, arithmetic tables, header builder, per-device state.
#include "common.cuh"
#include "transform_fast.cuh"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "cos_tables.inc"

namespace ie {

extern std::atomic<int> g_exact_transform;   // encode_image.cu
extern std::atomic<int> g_encode_variant;
extern std::atomic<int> g_copyout_variant;
extern std::atomic<int> g_fused_debug;
extern std::atomic<int> g_encode_pad_smem;
extern std::atomic<int> g_parse_variant;       // parse.cu
extern std::atomic<int> g_huffman_variant;     // huffman.cu
extern std::atomic<int> g_parse_grid4;         // parse.cu
extern std::atomic<int> g_decode_variant;      // decode_image.cu
extern std::atomic<int> g_me_variant;          // api_video.cu
extern std::atomic<int> g_video_decode_variant;
extern std::atomic<int> g_video_decode_batches;
extern std::atomic<int> g_video_encode_streams;
extern std::atomic<int> g_pframe_variant;       // encode_image.cu
extern std::atomic<uint64_t> g_stat_video_whole, g_stat_video_frames;
static thread_local std::string t_error;
std::atomic<uint64_t> g_launches{0};

void set_error(const std::string &msg) { t_error = msg; }

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    t_error = buf;
    cudaGetLastError();   // clear the sticky-less error state
    return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorNoKernelImageForDevice) ? IE_ENODEVICE
                                                                                                                 : IE_ECUDA;
}

// ---------------------------------------------------------------------------------------------------------
// tables
// ---------------------------------------------------------------------------------------------------------
static void build_tables(BlockTables &t, int N, const double *cs) {
    memset(&t, 0, sizeof t);
    const int NN = N * N;
    for (int i = 0; i < NN; i++) t.cs[i] = cs[i];
    for (int u = 0; u < N; u++)
        for (int v = 0; v < N; v++) {
            const double cu = (u == 0) ? 0.5 : M_SQRT1_2, cv = (v == 0) ? 0.5 : M_SQRT1_2;     // algo.cpp:294-297
            t.cc[u * N + v] = cu * cv;
        }
    for (int u = 0; u < N; u++)
        for (int v = 0; v < N; v++)
            for (int i = 0; i < N; i++)
                for (int j = 0; j < N; j++) {
                    const int uv = u * N + v, ij = i * N + j;
                    // volatile: each product must be rounded to binary64 on its own (no contraction, no reassociation)
                    volatile double a = cs[i * N + u] * cs[j * N + v];                           // algo.cpp:318-319
                    t.fw[uv * NN + ij] = a;
                    volatile double b = t.cc[uv] * cs[i * N + u];                                // algo.cpp:352-353
                    volatile double c = b * cs[j * N + v];                                       // algo.cpp:354
                    t.inv[uv * NN + ij] = c;
                }
    // zigzag (algo.cpp:68-87): order positions by (x+y, (x-y)&1 ? y : x); the keys are unique
    int order[kMaxNN];
    for (int i = 0; i < NN; i++) order[i] = i;
    auto key = [N](int idx) {
        const int x = idx % N, y = idx / N;
        return (x + y) * 64 + ((((int8_t)(x - y)) & 1) ? y : x);
    };
    for (int i = 1; i < NN; i++) {
        const int k = order[i];
        int j = i - 1;
        while (j >= 0 && key(order[j]) > key(k)) { order[j + 1] = order[j]; j--; }
        order[j + 1] = k;
    }
    for (int k = 0; k < NN; k++) { t.zz[k] = (uint8_t)order[k]; t.izz[order[k]] = (uint8_t)k; }
}

const HostTables &host_tables() {
    static HostTables ht;
    static std::once_flag once;
    std::call_once(once, [] {
        build_tables(ht.t4, 4, kCos4);
        build_tables(ht.t8, 8, kCos8);
        // cross-check the frozen cosines against the running libm (informational only; the frozen values win)
        for (int N : {4, 8}) {
            const double *cs = (N == 4) ? kCos4 : kCos8;
            const double f = M_PI_2 / double(N);
            for (int i = 0; i < N; i++)
                for (int u = 0; u < N; u++)
                    if (std::cos(double(2.0 * i + 1.0) * double(u) * f) != cs[i * N + u]) {
                        fprintf(stderr, "[imageencoder_b200] note: host libm cos differs from the frozen reference table "
                                        "(N=%d i=%d u=%d); using the frozen table\n", N, i, u);
                        return;
                    }
        }
    });
    return ht;
}

// ---------------------------------------------------------------------------------------------------------
// header
// ---------------------------------------------------------------------------------------------------------
namespace {
struct HostBits {
    uint32_t *w;
    uint32_t pos = 0;
    void put(unsigned len, uint32_t v) {                        // BitStream.cpp:73-77, MSB first
        for (unsigned p = 0; p < len; p++) {
            if ((v >> (len - 1 - p)) & 1u) w[pos >> 5] |= 1u << (31 - (pos & 31));
            pos++;
        }
    }
};
unsigned host_ffs(uint32_t v) { unsigned n = 0; while (v) { n++; v >>= 1; } return n; }   // utils.hpp:210-216, ffs(0)=0
}  // namespace

int build_header(HeaderParam &h, int N, const uint16_t *quant, int use_rle, uint32_t W, uint32_t H, int lead_bit,
                 int video, uint32_t frames, uint32_t gop, uint32_t merange) {
    memset(&h, 0, sizeof h);
    HostBits b{h.words};
    if (lead_bit) b.put(1, 0);                                   // ImageEncoder.cpp:84-86
    unsigned qb = 0;
    for (int i = 0; i < N * N; i++) qb = std::max(qb, host_ffs(quant[i]));   // MatrixReader.cpp:181-190
    b.put(5, qb);                                                // MatrixReader.cpp:150
    for (int i = 0; i < N * N; i++) b.put(qb, quant[i]);         // MatrixReader.cpp:151-155
    b.put(1, use_rle ? 1u : 0u);                                 // ImageEncoder.cpp:92
    b.put(15, W);                                                // ImageEncoder.cpp:93-94
    b.put(15, H);
    if (video) { b.put(15, frames); b.put(15, gop); b.put(15, merange); }   // VideoEncoder.cpp:71-73
    h.bits = b.pos;
    return IE_OK;
}

// fast-path constants: k = C(u)C(v)/Q, thr = 0.5 - delta with delta = 1.25 * (768 + 64) * 2^-24 * X * k  (transform_fast.cuh)
void make_fast_quant(FastQuant &fq, const uint16_t *quant, int N, double max_abs_sample) {
    memset(&fq, 0, sizeof fq);
    for (int u = 0; u < N; u++)
        for (int v = 0; v < N; v++) {
            const double cu = (u == 0) ? 0.5 : M_SQRT1_2, cv = (v == 0) ? 0.5 : M_SQRT1_2;
            const double k = cu * cv / (double)quant[u * N + v];
            const double delta = 1.25 * (768.0 + 64.0) * ldexp(1.0, -24) * max_abs_sample * k + 1e-6;
            fq.k[u * N + v] = (float)k;
            double thr = 0.5 - delta;
            if (thr < 0.0) thr = 0.0;                       // absurd quant/geometry: every coefficient takes the exact path
            fq.thr[u * N + v] = nextafterf((float)thr, 0.0f);
        }
    for (int i = N * N; i < kMaxNN; i++) { fq.k[i] = 0.f; fq.thr[i] = 1.f; }
}

// ---------------------------------------------------------------------------------------------------------
// per-device state
// ---------------------------------------------------------------------------------------------------------
static std::mutex g_mu;
static DeviceState g_dev[16];

int get_device_state(DeviceState **out) {
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice", __FILE__, __LINE__) == IE_ECUDA ? IE_ENODEVICE : IE_ENODEVICE;
    if (dev < 0 || dev >= 16) { set_error("device index out of range"); return IE_ENODEVICE; }
    std::lock_guard<std::mutex> lk(g_mu);
    DeviceState &s = g_dev[dev];
    if (s.device != dev) {
        cudaDeviceProp prop;
        e = cudaGetDeviceProperties(&prop, dev);
        if (e != cudaSuccess) { cuda_fail(e, "cudaGetDeviceProperties", __FILE__, __LINE__); return IE_ENODEVICE; }
        if (prop.major != 10) {
            char buf[160];
            snprintf(buf, sizeof buf, "device %d is sm_%d%d; this library carries sm_100a code only (no fallback path)", dev,
                     prop.major, prop.minor);
            set_error(buf);
            return IE_ENODEVICE;
        }
        const HostTables &ht = host_tables();
        IE_CUDA(cudaMalloc(&s.d_t4, sizeof(BlockTables)));
        IE_CUDA(cudaMalloc(&s.d_t8, sizeof(BlockTables)));
        IE_CUDA(cudaMemcpy(s.d_t4, &ht.t4, sizeof(BlockTables), cudaMemcpyHostToDevice));
        IE_CUDA(cudaMemcpy(s.d_t8, &ht.t8, sizeof(BlockTables), cudaMemcpyHostToDevice));
        IE_CUDA(cudaDeviceSynchronize());   // pageable H2D may still be in flight; other streams do not wait for stream 0
        s.sm_count = prop.multiProcessorCount;
        s.device = dev;
    }
    *out = &s;
    return IE_OK;
}

void drop_cached_sessions();      // api_image.cu: the host entry points' session cache
}  // namespace ie

extern "C" {

int ie_init(int device) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { ie::cuda_fail(e, "cudaSetDevice", __FILE__, __LINE__); return IE_ENODEVICE; }
    ie::DeviceState *s;
    return ie::get_device_state(&s);
}

void ie_shutdown(void) {
    // the cached sessions of the host entry points hold device scratch and point at the tables freed below
    ie::drop_cached_sessions();
    std::lock_guard<std::mutex> lk(ie::g_mu);
    int cur = -1;
    cudaGetDevice(&cur);
    for (auto &s : ie::g_dev) {
        if (s.device >= 0) {
            cudaSetDevice(s.device);
            cudaFree(s.d_t4);
            cudaFree(s.d_t8);
            s = ie::DeviceState{};
        }
    }
    if (cur >= 0) cudaSetDevice(cur);
}

const char *ie_last_error(void) { return ie::t_error.c_str(); }
const char *ie_version(void) { return "imageencoder_b200 0.1 (sm_100a)"; }
uint64_t ie_kernel_launch_count(void) { return ie::g_launches.load(); }
uint64_t ie_stat(const char *name) {
    if (name && !strcmp(name, "video_decode_whole_stream")) return ie::g_stat_video_whole.load();
    if (name && !strcmp(name, "video_decode_frame_by_frame")) return ie::g_stat_video_frames.load();
    return 0;
}

int ie_set_option(const char *name, int value) {
    if (name && !strcmp(name, "exact_transform")) { ie::g_exact_transform.store(value); return IE_OK; }
    if (name && !strcmp(name, "encode_variant")) {
        if (value < 0 || value > 9) {
            ie::set_error("encode_variant: 0 (scalar kernel), 1 (lean quantise), 2 (1 + packed f32x2 transform, the default), "
                          "3 / 4 (2 with a reduced staging area and 7 / 8 CTAs per SM), 5 (2 with the short-chain binary64 pre-check of "
                          "the exact queue), 6 / 7 (3 / 4 with it), 8 (fused persistent stream kernel, encode_fused.cu; other launch shapes fall back to 2), 9 (2 with the guard-band "
                          "coefficients evaluated by their owning lanes, no exact queue)");
            return IE_EINVAL;
        }
        ie::g_encode_variant.store(value);
        return IE_OK;
    }
    if (name && !strcmp(name, "encode_pad_smem")) {     // occupancy experiments: extra dynamic shared memory per tile-kernel CTA
        if (value < 0 || value > 150000) { ie::set_error("encode_pad_smem: 0 .. 150000 bytes"); return IE_EINVAL; }
        ie::g_encode_pad_smem.store(value);
        return IE_OK;
    }
    if (name && !strcmp(name, "fused_debug")) { ie::g_fused_debug.store(value); return IE_OK; }      // timing experiments, wrong output
    if (name && !strcmp(name, "parse_grid4")) {
        if (value < -1 || value > 2) { ie::set_error("parse_grid4: -1 (default: images 0, video frames 1), 0 (2048-bit groups, 4096-bit lead-in), 1 (1024 / 2048) or 2 (512 / 1024)"); return IE_EINVAL; }
        ie::g_parse_grid4.store(value);
        return IE_OK;
    }
    if (name && !strcmp(name, "huffman_variant")) {
        if (value < 0 || value > 1) { ie::set_error("huffman_variant: 1 (span histogram + bits / scan / pack kernels, default) or 0 (the round-1 kernels)"); return IE_EINVAL; }
        ie::g_huffman_variant.store(value);
        return IE_OK;
    }
    if (name && !strcmp(name, "parse_variant")) {
        if (value < 0 || value > 1) { ie::set_error("parse_variant: 0 (speculate + verify, exact path as fallback; default) or 1 (always the exact path)"); return IE_EINVAL; }
        ie::g_parse_variant.store(value);
        return IE_OK;
    }
    if (name && !strcmp(name, "copyout_variant")) {
        if (value < 0 || value > 3) { ie::set_error("copyout_variant: 0 (generic kernel), 1 (short path for interior chunks), 2 (1 + four chunks in flight, the default) or 3 (a warp per tile image, word by word)"); return IE_EINVAL; }
        ie::g_copyout_variant.store(value);
        return IE_OK;
    }
    if (name && !strcmp(name, "decode_variant")) {
        if (value < 0 || value > 1) { ie::set_error("decode_variant: 0 (default) or 1 (packed f32x2 inverse transform, experimental)"); return IE_EINVAL; }
        ie::g_decode_variant.store(value);
        return IE_OK;
    }
    if (name && !strcmp(name, "pframe_variant")) {
        if (value != 0 && value != 2) { ie::set_error("pframe_variant: 2 (packed f32x2 P-frame tiles, default) or 0 (scalar)"); return IE_EINVAL; }
        ie::g_pframe_variant.store(value);
        return IE_OK;
    }
    if (name && !strcmp(name, "video_encode_streams")) {
        if (value < 1 || value > 2) { ie::set_error("video_encode_streams: 2 (the GOPs of a batch in two halves on two streams, default) or 1"); return IE_EINVAL; }
        ie::g_video_encode_streams.store(value);
        return IE_OK;
    }
    if (name && !strcmp(name, "video_decode_batches")) {
        if (value < 1 || value > 32) { ie::set_error("video_decode_batches: 1 .. 32 GOP batches (default 4)"); return IE_EINVAL; }
        ie::g_video_decode_batches.store(value);
        return IE_OK;
    }
    if (name && !strcmp(name, "video_decode_variant")) {
        if (value < 0 || value > 1) { ie::set_error("video_decode_variant: 1 (whole-stream parse, frame k of every GOP per launch; default) or 0 (frame by frame)"); return IE_EINVAL; }
        ie::g_video_decode_variant.store(value);
        return IE_OK;
    }
    if (name && !strcmp(name, "me_variant")) {
        if (value < 0 || value > 2) { ie::set_error("me_variant: 2 (eight lanes per MacroBlock, four MacroBlocks per warp; default), 0 (warp per MacroBlock) or 1 (0 with REDUX reductions)"); return IE_EINVAL; }
        ie::g_me_variant.store(value);
        return IE_OK;
    }
    ie::set_error("unknown option");
    return IE_EINVAL;
}

size_t ie_max_encoded_bytes(uint32_t width, uint32_t height, uint32_t block, uint32_t frames) {
    if (block != 4 && block != 8) return 0;
    const size_t nblk = (size_t)(width / block) * (height / block);
    const size_t per_block = 4 + 16 + 16 * (size_t)block * block;                 // Block.cpp:346-354
    size_t bits = 1100 + 45 + (size_t)frames * (nblk * per_block + (size_t)(width / 16) * (height / 16) * 32);
    size_t bytes = (bits + 7) / 8;
    bytes += 1024;                                                                // Huffman dictionary + revert byte
    return (bytes + 15) / 16 * 16 + 16;
}

}  // extern "C"
