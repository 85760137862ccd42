// C-ABI: sessions and the image entry points (host-buffer and device-resident).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "api_internal.cuh"

namespace ie {

int session_reserve(uint8_t **p, size_t *cap, size_t need) {
    if (*cap >= need && *p) return IE_OK;
    if (*p) IE_CUDA(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    need = (need + 255) / 256 * 256;
    IE_CUDA(cudaMalloc(p, need));
    *cap = need;
    return IE_OK;
}

int session_ensure_scan(ie_session *s, unsigned images, unsigned tiles) {
    if (s->d_tile_state && s->images >= images && s->max_tiles >= tiles) return IE_OK;
    if (s->d_tile_state) { cudaFree(s->d_tile_state); cudaFree(s->d_bnd); cudaFree(s->d_ticket); cudaFree(s->d_counter); }
    images = std::max(images, s->images);
    tiles = std::max(tiles, s->max_tiles);
    const size_t n = (size_t)images * tiles;
    IE_CUDA(cudaMalloc(&s->d_tile_state, n * sizeof(unsigned long long)));
    IE_CUDA(cudaMalloc(&s->d_bnd, n * sizeof(TileBoundary)));
    IE_CUDA(cudaMalloc(&s->d_ticket, (images + 2) * sizeof(unsigned)));      // + the fused encoder's CTA id / finished-CTA words
    IE_CUDA(cudaMalloc(&s->d_counter, images * sizeof(unsigned long long)));
    IE_CUDA(cudaMemset(s->d_tile_state, 0, n * sizeof(unsigned long long)));
    IE_CUDA(cudaMemset(s->d_bnd, 0, n * sizeof(TileBoundary)));
    IE_CUDA(cudaMemset(s->d_ticket, 0, (images + 2) * sizeof(unsigned)));
    IE_CUDA(cudaMemset(s->d_counter, 0, images * sizeof(unsigned long long)));
    // cudaMemset on device memory is asynchronous on the legacy default stream, which the sessions' non-blocking streams
    // do not wait for: make the zeroes visible before any kernel can touch these arrays
    IE_CUDA(cudaDeviceSynchronize());
    s->images = images;
    s->max_tiles = tiles;
    s->epoch = 0;
    return IE_OK;
}

int session_ensure_err(ie_session *s) {
    if (s->d_err) return IE_OK;
    IE_CUDA(cudaMalloc(&s->d_err, sizeof(int)));
    IE_CUDA(cudaMemset(s->d_err, 0, sizeof(int)));
    IE_CUDA(cudaDeviceSynchronize());        // see session_ensure_scan
    return IE_OK;
}

int check_quant(const uint16_t *quant, int N) {
    if (!quant) { set_error("quant matrix is NULL"); return IE_EINVAL; }
    for (int i = 0; i < N * N; i++)
        if (quant[i] == 0) { set_error("quant matrix entry 0 (division by zero in the reference, MatrixReader.cpp:104)"); return IE_EINVAL; }
    return IE_OK;
}

int check_dims(uint32_t W, uint32_t H, uint32_t N) {
    if (N != 4 && N != 8) { set_error("block size must be 4 or 8"); return IE_EINVAL; }
    if (W == 0 || H == 0 || W > 32767 || H > 32767) { set_error("width/height must be in 1..32767 (DIM_BITS=15, ImageBase.hpp:76)"); return IE_EINVAL; }
    if (W % N || H % N) { set_error("width/height must be multiples of the block size (ImageEncoder.cpp:26-27)"); return IE_EINVAL; }
    return IE_OK;
}

// Writes the stream prefix (zero bits up to first_bit, then the header) into whole chunks and sets the bit counters.
__global__ void stream_init_kernel(uint8_t *out, size_t out_stride, unsigned images, HeaderParam hdr, unsigned first_bit,
                                   unsigned long long *counter) {
    const unsigned img = blockIdx.x;
    if (img >= images) return;
    const unsigned total = first_bit + hdr.bits;
    const unsigned nwords = ((total + 127) / 128) * 4;
    unsigned *o = reinterpret_cast<unsigned *>(out + (size_t)img * out_stride);
    for (unsigned i = threadIdx.x; i < nwords; i += blockDim.x) {
        // word i of the stream = header bits [32 i - first_bit, +32)
        const long long hb = (long long)i * 32 - (long long)first_bit;
        unsigned v = 0;
        const int sh = (int)(((hb % 32) + 32) % 32);
        const long long wi = (hb - sh) / 32;               // floor division
        const unsigned hi = (wi >= 0 && wi < kHdrWordsMax) ? hdr.words[wi] : 0u;
        const unsigned lo = (wi + 1 >= 0 && wi + 1 < kHdrWordsMax) ? hdr.words[wi + 1] : 0u;
        v = sh ? ((hi << sh) | (lo >> (32 - sh))) : hi;
        o[i] = __byte_perm(v, 0, 0x0123);
    }
    if (threadIdx.x == 0) counter[img] = total;
}

// The same for a shard of a multi-GPU stream: the shard's first bit in the global stream is the sum of the totals of the
// shards in front of it (a device array, straight from the all-gather); the local buffer starts at the 128-bit chunk that
// holds that bit, so the prefix is (first % 128) zero bits, then the header (rank 0 only).
__global__ void stream_init_shard_kernel(uint8_t *out, HeaderParam hdr, const unsigned long long *shard_totals, unsigned shard_index,
                                         unsigned long long *counter, unsigned long long *bit_base, unsigned long long *first_out) {
    pdl_wait();
    __shared__ unsigned s_fb;
    if (threadIdx.x == 0) {
        unsigned long long first = 0;
        for (unsigned i = 0; i < shard_index; i++) first += shard_totals[i];
        s_fb = (unsigned)(first % 128);
        if (first_out) *first_out = first;
    }
    __syncthreads();
    const unsigned first_bit = s_fb;
    const unsigned total = first_bit + hdr.bits;
    const unsigned nwords = ((total + 127) / 128) * 4;
    unsigned *o = reinterpret_cast<unsigned *>(out);
    for (unsigned i = threadIdx.x; i < max(nwords, 4u); i += blockDim.x) {
        const long long hb = (long long)i * 32 - (long long)first_bit;
        const int sh = (int)(((hb % 32) + 32) % 32);
        const long long wi = (hb - sh) / 32;               // floor division
        const unsigned hi = (wi >= 0 && wi < kHdrWordsMax) ? hdr.words[wi] : 0u;
        const unsigned lo = (wi + 1 >= 0 && wi + 1 < kHdrWordsMax) ? hdr.words[wi + 1] : 0u;
        const unsigned v = sh ? ((hi << sh) | (lo >> (32 - sh))) : hi;
        o[i] = __byte_perm(v, 0, 0x0123);
    }
    if (threadIdx.x == 0) { counter[0] = total; bit_base[0] = total; }
}

int launch_stream_init(uint8_t *out, size_t out_stride, unsigned images, const HeaderParam &hdr, unsigned first_bit,
                       unsigned long long *counter, cudaStream_t stream) {
    stream_init_kernel<<<images, 64, 0, stream>>>(out, out_stride, images, hdr, first_bit, counter);
    count_launch();
    IE_CUDA(cudaGetLastError());
    return IE_OK;
}

void make_k2(float *k2, const uint16_t *quant, int N) {
    for (int u = 0; u < N; u++)
        for (int v = 0; v < N; v++) {
            const double cu = (u == 0) ? 0.5 : 0.70710678118654752440, cv = (v == 0) ? 0.5 : 0.70710678118654752440;
            k2[u * N + v] = (float)(cu * cv * (double)quant[u * N + v]);
        }
    for (int i = N * N; i < kMaxNN; i++) k2[i] = 0.f;
}

int make_quant(QuantParam &q, const uint16_t *quant, int N) {
    memset(&q, 0, sizeof q);
    for (int i = 0; i < N * N; i++) q.m[i] = (double)quant[i];          // MatrixReader.cpp:128,195-198
    for (int i = N * N; i < kMaxNN; i++) q.m[i] = 1.0;
    return IE_OK;
}

// Encode `images` equally sized images that are resident on the device.
int encode_images_dev(ie_session *s, const uint8_t *d_raw, size_t img_stride, unsigned images, uint32_t W, uint32_t H, int N,
                      const uint16_t *quant, int use_rle, int lead_bit, int write_header, unsigned first_bit, int bits_only,
                      uint8_t *d_out, size_t out_stride, size_t out_cap, cudaStream_t stream, int append, uint32_t header_H, int split,
                      uint64_t *d_out_bits) {
    // split: only the tile kernel runs (tile images stay in the session's scratch), the parameters and the header are kept in
    // the session for ie_encode_image_end_dev
    // append: the streams continue at their device-resident bit counters (no prefix is written); header_H: the height the
    // header announces when this call encodes only the first stripe of a taller image
    IE_TRY(check_dims(W, H, N));
    IE_TRY(check_quant(quant, N));
    if (first_bit >= 128) { set_error("first_bit must be < 128"); return IE_EINVAL; }
    if (!bits_only && ((uintptr_t)d_out % 16 || out_stride % 16)) { set_error("stream buffers must be 16-byte aligned"); return IE_EINVAL; }
    const unsigned nblocks = (W / N) * (H / N);
    const unsigned TB = encode_tile_blocks(N);
    const unsigned tiles = (nblocks + TB - 1) / TB;
    // the fused stream kernel (encode_fused.cu): one image, one launch, no scratch round trip
    const bool fused = g_encode_variant.load() == 8 && !g_exact_transform.load() && images == 1 && !bits_only && !split;
    unsigned n_wtiles = 0, n_ctatiles = 0;
    if (fused) encode_fused_sizes(N, nblocks, n_wtiles, n_ctatiles);
    IE_TRY(session_ensure_scan(s, images, std::max(tiles, n_wtiles)));
    IE_TRY(session_ensure_err(s));

    HeaderParam hdr;
    memset(&hdr, 0, sizeof hdr);
    if (write_header && !append)
        IE_TRY(build_header(hdr, N, quant, use_rle, W, header_H ? header_H : (s->header_height ? s->header_height : H), lead_bit, 0, 0, 0, 0));
    if (append || split) {
        if (bits_only) { set_error("append/split and bits_only exclude each other"); return IE_EINVAL; }
    } else if (!bits_only) {
        const size_t need = ((size_t)first_bit + hdr.bits + 127) / 128 * 16;
        if (out_cap < need) { set_error("output buffer too small for the header"); return IE_ENOSPC; }
        // (the prefix is written by tile 0 of the tile kernel: no launch of its own)
    } else {
        IE_CUDA(cudaMemsetAsync(s->d_counter, 0, images * sizeof(unsigned long long), stream));
    }

    EncodeParams p;
    memset(&p, 0, sizeof p);
    p.src = d_raw; p.pitch = W; p.img_stride = img_stride;
    p.bx = W / N; p.nblocks = nblocks; p.tiles_per_image = tiles;
    p.use_rle = use_rle ? 1 : 0; p.bits_only = bits_only;
    make_quant(p.quant, quant, N);
    make_fast_quant(p.fq, quant, N, 128.0);
    p.dc_den2 = 8 * (int)quant[0]; p.dc_rcp = 1.0f / (float)p.dc_den2;
    p.tab = (N == 8) ? s->dev->d_t8 : s->dev->d_t4;
    p.out = d_out; p.out_stride = out_stride; p.out_cap = out_cap;
    p.bit_counter = s->d_counter; p.err = s->d_err;
    p.write_prefix = (!append && !split && !bits_only) ? 1 : 0;
    p.prefix_first = first_bit;
    p.hdr = hdr;
    p.out_bits = reinterpret_cast<unsigned long long *>(d_out_bits);
    p.scan = s->scan_state();
    if (fused) {
        p.tiles_per_image = n_wtiles;
        p.scan.ticket = s->d_ticket + s->images;
        return launch_encode_fused(N, p, append ? 1 : 0, s->dev->sm_count, stream);
    }
    {
        const size_t ntot = (size_t)images * tiles;
        p.slot_bytes = encode_tile_slot_bytes(N);
        if (!bits_only) IE_TRY(session_reserve(&s->d_tile_scratch, &s->tile_scratch_cap, ntot * p.slot_bytes));
        IE_TRY(session_reserve(&s->d_tile_meta, &s->tile_meta_cap, ntot * (sizeof(unsigned long long) + sizeof(unsigned)) + 64));
        p.tile_scratch = s->d_tile_scratch;
        p.bit_base = reinterpret_cast<unsigned long long *>(s->d_tile_meta);
        p.tile_bits = reinterpret_cast<unsigned *>(s->d_tile_meta + ntot * sizeof(unsigned long long));
    }
    // scan arrays are indexed [image][tile] with stride tiles_per_image (the allocation is at least that large)
    if (split) {
        p.phase = 1;
        IE_TRY(launch_encode_tiles(N, p, images, stream));
        p.phase = 2;
        s->split_params = p;
        s->split_hdr = hdr;
        s->split_pending = true;
        return IE_OK;
    }
    return launch_encode_tiles(N, p, images, stream);
}

int session_ensure_pipeline(ie_session *s) {
    if (s->stream_in) return IE_OK;
    IE_CUDA(cudaStreamCreateWithFlags(&s->stream_in, cudaStreamNonBlocking));
    IE_CUDA(cudaStreamCreateWithFlags(&s->stream_out, cudaStreamNonBlocking));
    IE_CUDA(cudaStreamCreateWithFlags(&s->stream_aux, cudaStreamNonBlocking));
    IE_CUDA(cudaEventCreateWithFlags(&s->ev_aux_fork, cudaEventDisableTiming));
    IE_CUDA(cudaEventCreateWithFlags(&s->ev_aux_join, cudaEventDisableTiming));
    for (int i = 0; i < ie_session::kMaxStripes; i++) {
        IE_CUDA(cudaEventCreateWithFlags(&s->ev_in[i], cudaEventDisableTiming));
        IE_CUDA(cudaEventCreateWithFlags(&s->ev_done[i], cudaEventDisableTiming));
    }
    return IE_OK;
}

int read_err_flag(ie_session *s, cudaStream_t stream) {
    if (!s->d_err) return IE_OK;
    int e = 0;
    IE_CUDA(cudaMemcpyAsync(&e, s->d_err, sizeof(int), cudaMemcpyDeviceToHost, stream));
    IE_CUDA(cudaStreamSynchronize(stream));
    if (e != 0) {
        IE_CUDA(cudaMemsetAsync(s->d_err, 0, sizeof(int), stream));
        set_error(e == IE_ENOSPC ? "output buffer too small" : (e == IE_EFORMAT ? "malformed stream" : "device-side error"));
        return e;
    }
    return IE_OK;
}

// ---- header parsing on the host (MatrixReader.cpp:45-57, ImageBase.cpp:122-128, VideoBase.cpp:72-83) ---------------
namespace {
struct HostReader {
    const uint8_t *b;
    size_t n;
    size_t pos;
    uint32_t get(unsigned l) {
        uint32_t v = 0;
        for (unsigned i = 0; i < l; i++) {
            uint32_t bit = 0;
            if ((pos >> 3) < n) { bit = (b[pos >> 3] >> (7 - (pos & 7))) & 1u; pos++; }
            v |= bit << (l - i - 1);
        }
        return v;
    }
};
}  // namespace

int parse_header(const uint8_t *bytes, size_t n, size_t start_bit, int N, ParsedHeader &h, int video) {
    HostReader r{bytes, n, start_bit};
    const unsigned qb = r.get(5);
    for (int i = 0; i < N * N; i++) h.quant[i] = (uint16_t)r.get(qb);
    h.use_rle = (int)r.get(1);
    h.W = r.get(15);
    h.H = r.get(15);
    h.frames = h.gop = h.merange = 0;
    if (video) { h.frames = r.get(15); h.gop = r.get(15); h.merange = r.get(15); }
    h.end_bit = r.pos;
    return IE_OK;
}

int decode_image_dev(ie_session *s, const uint8_t *d_enc, size_t enc_bytes, size_t start_bit, int N, const ParsedHeader &h,
                     uint8_t *d_out, size_t out_cap, cudaStream_t stream, uint8_t *host_out) {
    const uint32_t W = h.W, H = h.H;
    IE_TRY(check_dims(W, H, N));
    if ((uintptr_t)d_enc % 16) { set_error("encoded stream must be 16-byte aligned (and readable up to its size rounded up to 4)"); return IE_EINVAL; }
    if ((size_t)W * H > out_cap) { set_error("decoded image does not fit the output buffer"); return IE_ENOSPC; }
    for (int i = 0; i < N * N; i++)
        if (h.quant[i] == 0) { /* a zero entry decodes to zero coefficients * 0: allowed */ }
    const unsigned nblocks = (W / N) * (H / N);
    IE_TRY(session_ensure_scan(s, 1, 1));
    IE_TRY(session_ensure_err(s));
    const size_t need_off = ((size_t)nblocks + 1) * sizeof(unsigned long long) + 64;
    if (s->block_off_cap < need_off) {
        if (s->d_block_off) IE_CUDA(cudaFree(s->d_block_off));
        IE_CUDA(cudaMalloc(&s->d_block_off, need_off));
        s->block_off_cap = need_off;
    }
    // two u64 of launch constants live at the end of the offsets array: [stream bits, first block bit]
    unsigned long long consts[2] = {(unsigned long long)enc_bytes * 8ull, (unsigned long long)h.end_bit};
    (void)start_bit;
    unsigned long long *d_consts = s->d_block_off + (nblocks + 1);
    IE_CUDA(cudaMemcpyAsync(d_consts, consts, sizeof consts, cudaMemcpyHostToDevice, stream));
    DecodeParams p;
    memset(&p, 0, sizeof p);
    p.enc = d_enc; p.enc_stride = 0; p.enc_bits = d_consts; p.start_bit = d_consts + 1;
    p.block_off = s->d_block_off; p.nblocks = nblocks; p.bx = W / N; p.N = N; p.use_rle = h.use_rle;
    make_quant(p.quant, h.quant, N);
    make_k2(p.k2, h.quant, N);
    p.tab = (N == 8) ? s->dev->d_t8 : s->dev->d_t4;
    p.out = d_out; p.out_stride = 0; p.pitch = W; p.err = s->d_err;
    IE_TRY(session_reserve(&s->d_parse, &s->parse_cap, parse_scratch_bytes(enc_bytes, N)));
    IE_TRY(launch_parallel_parse(p, enc_bytes * 8, s->d_parse, stream));
    if (!host_out) return launch_decode_blocks(p, 1, stream);
    // host entry point: stripes of whole block rows (>= 4 MiB of pixels); stripe i's pixels travel to the host on stream_out
    // while stripe i + 1 is decoded -- the read-back (2.4x the stream's size on config 2) is what the call's time is made of
    IE_TRY(session_ensure_pipeline(s));
    const uint32_t block_rows = H / N;
    const size_t npx = (size_t)W * H;
    uint32_t stripes = (uint32_t)std::min<size_t>(ie_session::kMaxStripes, std::max<size_t>(1, npx / ((size_t)4 << 20)));
    stripes = std::min(stripes, block_rows);
    const uint32_t rows_per = (block_rows + stripes - 1) / stripes;
    stripes = (block_rows + rows_per - 1) / rows_per;
    for (uint32_t i = 0; i < stripes; i++) {
        const uint32_t r0 = i * rows_per, r1 = std::min(block_rows, r0 + rows_per);
        p.block_base = r0 * (W / N); p.block_end = r1 * (W / N);
        IE_TRY(launch_decode_blocks(p, 1, stream));
        IE_CUDA(cudaEventRecord(s->ev_done[i], stream));
        IE_CUDA(cudaStreamWaitEvent(s->stream_out, s->ev_done[i], 0));
        const size_t o = (size_t)r0 * N * W, n = (size_t)(r1 - r0) * N * W;
        IE_CUDA(cudaMemcpyAsync(host_out + o, d_out + o, n, cudaMemcpyDeviceToHost, s->stream_out));
    }
    return IE_OK;
}

// ---- cached sessions for the host-buffer entry points -----------------------------------------------------------
// A session owns scratch sized for its shape (tile scratch, staging buffers: hundreds of MB for large images), so sessions
// are cached -- and LEASED: a host call holds its session exclusively from entry to return (SessionLease), so two threads
// encoding images of the same shape never share staging buffers, streams or scan state.  A second concurrent call of a
// shape gets a second session.  The cache is bounded: when it is full the least recently used IDLE session is destroyed
// (a leased one never is).
static std::mutex g_smu;
struct CachedSession {
    std::tuple<int, int, uint32_t, uint32_t, uint32_t, uint32_t> key;
    ie_session *s;
    unsigned long long last_use;
    bool busy;
};
static std::vector<CachedSession> g_sessions;
static unsigned long long g_session_clock = 0;
constexpr size_t kMaxCachedSessions = 8;

int SessionLease::acquire(int kind, uint32_t W, uint32_t H, uint32_t N, uint32_t frames) {
    release();
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); set_error("no CUDA device (this library has no CPU path)"); return IE_ENODEVICE; }
    DeviceState *ds = nullptr;
    IE_TRY(get_device_state(&ds));                 // (re-)initialises the device tables, e.g. after ie_shutdown()
    const auto key = std::make_tuple(dev, kind, W, H, N, frames);
    {
        std::lock_guard<std::mutex> lk(g_smu);
        for (auto &c : g_sessions)
            if (!c.busy && c.key == key) { c.busy = true; c.last_use = ++g_session_clock; s_ = c.s; s_->dev = ds; return IE_OK; }
        // make room: destroy idle sessions, least recently used first
        while (g_sessions.size() >= kMaxCachedSessions) {
            int lru = -1;
            for (size_t i = 0; i < g_sessions.size(); i++)
                if (!g_sessions[i].busy && (lru < 0 || g_sessions[i].last_use < g_sessions[(size_t)lru].last_use)) lru = (int)i;
            if (lru < 0) break;                    // every cached session is leased: grow past the bound for now
            ie_session_destroy(g_sessions[(size_t)lru].s);
            g_sessions.erase(g_sessions.begin() + lru);
        }
    }
    ie_session *s = nullptr;
    IE_TRY(ie_session_create(&s, kind, W, H, N, frames));       // outside the lock: allocations can take milliseconds
    std::lock_guard<std::mutex> lk(g_smu);
    g_sessions.push_back(CachedSession{key, s, ++g_session_clock, true});
    s_ = s;
    return IE_OK;
}

void SessionLease::release() {
    if (!s_) return;
    std::lock_guard<std::mutex> lk(g_smu);
    for (auto &c : g_sessions)
        if (c.s == s_) { c.busy = false; c.last_use = ++g_session_clock; }
    s_ = nullptr;
}

void drop_cached_sessions() {
    std::lock_guard<std::mutex> lk(g_smu);
    for (auto it = g_sessions.begin(); it != g_sessions.end();) {
        if (it->busy) { ++it; continue; }          // a call still in flight keeps its session
        ie_session_destroy(it->s);
        it = g_sessions.erase(it);
    }
}

}  // namespace ie

using namespace ie;

extern "C" {

int ie_session_create(ie_session **out, int kind, uint32_t W, uint32_t H, uint32_t N, uint32_t frames) {
    if (!out) return IE_EINVAL;
    DeviceState *dev = nullptr;
    IE_TRY(get_device_state(&dev));
    ie_session *s = new ie_session();
    s->kind = kind; s->W = W; s->H = H; s->N = N; s->frames = frames ? frames : 1;
    s->dev = dev; s->device = dev->device;
    cudaError_t e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMallocHost(&s->h_pinned, 64 * sizeof(unsigned long long));
    if (e != cudaSuccess) { delete s; return cuda_fail(e, "session create", __FILE__, __LINE__); }
    *out = s;
    return IE_OK;
}

void ie_session_destroy(ie_session *s) {
    if (!s) return;
    for (int k = 0; k < ie_session::kDecodeWorkers; k++) {
        if (s->workers[k]) ie_session_destroy(s->workers[k]);
        if (s->ev_join[k]) cudaEventDestroy(s->ev_join[k]);
    }
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    cudaFree(s->d_tile_state); cudaFree(s->d_bnd); cudaFree(s->d_ticket); cudaFree(s->d_counter); cudaFree(s->d_err);
    cudaFree(s->d_block_off); cudaFree(s->d_parse); cudaFree(s->d_tile_scratch); cudaFree(s->d_tile_meta); cudaFree(s->d_scratch); cudaFree(s->d_in); cudaFree(s->d_out); cudaFree(s->d_tmp);
    if (s->h_pinned) cudaFreeHost(s->h_pinned);
    if (s->h_huff) cudaFreeHost(s->h_huff);
    if (s->huff_ctx) free(s->huff_ctx);
    if (s->stream) cudaStreamDestroy(s->stream);
    if (s->stream_in) {
        cudaStreamDestroy(s->stream_in); cudaStreamDestroy(s->stream_out);
        if (s->stream_aux) cudaStreamDestroy(s->stream_aux);
        if (s->ev_aux_fork) cudaEventDestroy(s->ev_aux_fork);
        if (s->ev_aux_join) cudaEventDestroy(s->ev_aux_join);
        for (int i = 0; i < ie_session::kMaxStripes; i++) { cudaEventDestroy(s->ev_in[i]); cudaEventDestroy(s->ev_done[i]); }
    }
    delete s;
}

int ie_encode_image_dev(ie_session *s, const uint8_t *d_raw, uint32_t W, uint32_t H, const uint16_t *quant, int use_rle,
                        int lead_bit, int write_header, uint64_t first_bit, uint8_t *d_out, size_t out_cap,
                        uint64_t *d_out_bits, void *stream) {
    if (!s || !d_raw || !d_out) { set_error("NULL argument"); return IE_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    return encode_images_dev(s, d_raw, 0, 1, W, H, (int)s->N, quant, use_rle, lead_bit, write_header, (unsigned)first_bit, 0,
                             d_out, 0, out_cap, st, 0, 0, 0, d_out_bits);
}

int ie_encode_images_dev(ie_session *s, const uint8_t *d_raws, size_t raw_stride, uint32_t count, uint32_t W, uint32_t H,
                         const uint16_t *quant, int use_rle, int lead_bit, uint8_t *d_out, size_t out_stride, uint64_t *d_out_bits,
                         void *stream) {
    if (!s || !d_raws || !d_out || count == 0) { set_error("NULL/empty argument"); return IE_EINVAL; }
    if (raw_stride < (size_t)W * H) { set_error("raw_stride smaller than an image"); return IE_EINVAL; }
    return encode_images_dev(s, d_raws, raw_stride, count, W, H, (int)s->N, quant, use_rle, lead_bit, 1, 0, 0, d_out, out_stride, out_stride,
                             (cudaStream_t)stream, 0, 0, 0, d_out_bits);
}

int ie_encode_image_begin_dev(ie_session *s, const uint8_t *d_raw, uint32_t W, uint32_t H, const uint16_t *quant, int use_rle,
                              int lead_bit, int write_header, uint64_t *d_total_bits, void *stream) {
    if (!s || !d_raw || !d_total_bits) { set_error("NULL argument"); return IE_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    IE_TRY(encode_images_dev(s, d_raw, 0, 1, W, H, (int)s->N, quant, use_rle, lead_bit, write_header, 0, 0, nullptr, 0, 0, st, 0, 0, 1));
    return launch_tile_totals(s->split_params.tile_bits, s->split_params.tiles_per_image, 1, s->split_hdr.bits,
                              reinterpret_cast<unsigned long long *>(d_total_bits), st);
}

int ie_encode_image_end_dev(ie_session *s, const uint64_t *d_shard_totals, uint32_t shard_index, uint8_t *d_out, size_t out_cap,
                            uint64_t *d_out_bits, uint64_t *d_first_bit, void *stream) {
    if (!s || !d_shard_totals || !d_out) { set_error("NULL argument"); return IE_EINVAL; }
    if (!s->split_pending) { set_error("ie_encode_image_end_dev without ie_encode_image_begin_dev"); return IE_EINVAL; }
    if ((uintptr_t)d_out % 16) { set_error("stream buffers must be 16-byte aligned"); return IE_EINVAL; }
    if (out_cap < ((size_t)128 + s->split_hdr.bits + 127) / 128 * 16) { set_error("output buffer too small for the header"); return IE_ENOSPC; }
    cudaStream_t st = (cudaStream_t)stream;
    EncodeParams p = s->split_params;
    s->split_pending = false;
    p.out = d_out; p.out_stride = 0; p.out_cap = out_cap;
    IE_CUDA(launch_pdl(stream_init_shard_kernel, dim3(1), dim3(64), 0, st, d_out, s->split_hdr,
                       reinterpret_cast<const unsigned long long *>(d_shard_totals), (unsigned)shard_index, s->d_counter, p.bit_base,
                       reinterpret_cast<unsigned long long *>(d_first_bit)));
    count_launch();
    IE_TRY(launch_tile_copyout(p, 1, st));
    if (d_out_bits) IE_CUDA(cudaMemcpyAsync(d_out_bits, s->d_counter, sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
    return IE_OK;
}

int ie_session_set_header_height(ie_session *s, uint32_t full_height) {
    if (!s || full_height > 32767) { set_error("bad argument"); return IE_EINVAL; }
    s->header_height = full_height;
    return IE_OK;
}

int ie_image_bits_dev(ie_session *s, const uint8_t *d_raw, uint32_t W, uint32_t H, const uint16_t *quant, int use_rle,
                      uint64_t *d_total_bits, void *stream) {
    if (!s || !d_raw || !d_total_bits) { set_error("NULL argument"); return IE_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    IE_TRY(encode_images_dev(s, d_raw, 0, 1, W, H, (int)s->N, quant, use_rle, 0, 0, 0, 1, nullptr, 0, 0, st));
    IE_CUDA(cudaMemcpyAsync(d_total_bits, s->d_counter, sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
    return IE_OK;
}

int ie_encode_image(const uint8_t *raw, uint32_t W, uint32_t H, uint32_t N, const uint16_t *quant, int use_rle, int huffman,
                    uint8_t *out, size_t out_cap, size_t *out_bytes) {
    if (!raw || !out || !out_bytes) { set_error("NULL argument"); return IE_EINVAL; }
    IE_TRY(check_dims(W, H, N));
    IE_TRY(check_quant(quant, (int)N));
    SessionLease lease;
    IE_TRY(lease.acquire(0, W, H, N, 1));
    ie_session *s = lease.get();
    const size_t npx = (size_t)W * H;
    const size_t cap = ie_max_encoded_bytes(W, H, N, 1);
    IE_TRY(session_reserve(&s->d_in, &s->d_in_cap, npx));
    IE_TRY(session_reserve(&s->d_out, &s->d_out_cap, cap));
    IE_TRY(session_ensure_pipeline(s));
    cudaStream_t st = s->stream;

    // Stripes of whole block rows (>= 4 MiB of pixels each, at most kMaxStripes): stripe i is copied in on stream_in, encoded
    // on `st` as an append to the stream (device-resident bit counter), and every 128-bit chunk it completed leaves on
    // stream_out while the following stripes are still on their way in.  PCIe runs in both directions at once; what is left
    // after the last pixel arrived is one stripe of encoding and its share of the output.
    const uint32_t block_rows = H / N;
    uint32_t stripes = (uint32_t)std::min<size_t>(ie_session::kMaxStripes, std::max<size_t>(1, npx / ((size_t)4 << 20)));
    stripes = std::min(stripes, block_rows);
    const uint32_t rows_per = (block_rows + stripes - 1) / stripes;          // block rows per stripe
    stripes = (block_rows + rows_per - 1) / rows_per;
    for (uint32_t i = 0; i < stripes; i++) {
        const uint32_t r0 = i * rows_per, r1 = std::min(block_rows, r0 + rows_per);
        const size_t off = (size_t)r0 * N * W, len = (size_t)(r1 - r0) * N * W;
        IE_CUDA(cudaMemcpyAsync(s->d_in + off, raw + off, len, cudaMemcpyHostToDevice, s->stream_in));
        IE_CUDA(cudaEventRecord(s->ev_in[i], s->stream_in));
        IE_CUDA(cudaStreamWaitEvent(st, s->ev_in[i], 0));
        IE_TRY(encode_images_dev(s, s->d_in + off, 0, 1, W, (r1 - r0) * N, (int)N, quant, use_rle, huffman ? 0 : 1, 1, 0, 0, s->d_out, 0,
                                 s->d_out_cap, st, i > 0, H));
        IE_CUDA(cudaMemcpyAsync(s->h_pinned + 1 + i, s->d_counter, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        IE_CUDA(cudaEventRecord(s->ev_done[i], st));
    }
    size_t bytes = 0, sent = 0;
    bool fits = true;
    for (uint32_t i = 0; i < stripes; i++) {
        IE_CUDA(cudaEventSynchronize(s->ev_done[i]));
        const unsigned long long bits = s->h_pinned[1 + i];
        const bool last = (i + 1 == stripes);
        bytes = (size_t)((bits + 7) / 8);                                  // util::round_to_byte, ImageBase.cpp:316
        if (huffman) continue;
        const size_t fin = last ? bytes : (size_t)(bits / 128) * 16;      // chunks no later launch will touch
        if (fin > out_cap) { fits = false; continue; }
        if (fin > sent) {
            IE_CUDA(cudaMemcpyAsync(out + sent, s->d_out + sent, fin - sent, cudaMemcpyDeviceToHost, s->stream_out));
            sent = fin;
        }
    }
    IE_TRY(read_err_flag(s, st));                              // synchronises
    if (huffman) {
        IE_TRY(session_reserve(&s->d_tmp, &s->d_tmp_cap, bytes + 4096 + 32));
        size_t hb = 0;
        IE_TRY(ie_huffman_encode_dev(s, s->d_out, bytes, s->d_tmp, s->d_tmp_cap, &hb, st));
        bytes = hb;
        fits = bytes <= out_cap;
        if (fits) IE_CUDA(cudaMemcpyAsync(out, s->d_tmp, bytes, cudaMemcpyDeviceToHost, st));
        IE_CUDA(cudaStreamSynchronize(st));
    }
    IE_CUDA(cudaStreamSynchronize(s->stream_out));
    *out_bytes = bytes;
    if (!fits) { set_error("output buffer too small"); return IE_ENOSPC; }
    return IE_OK;
}

int ie_encode_images(const uint8_t *raws, uint32_t count, uint32_t W, uint32_t H, uint32_t N, const uint16_t *quant, int use_rle,
                     int huffman, uint8_t *out, size_t out_stride, size_t *out_bytes) {
    if (!raws || !out || !out_bytes || count == 0) { set_error("NULL/empty argument"); return IE_EINVAL; }
    IE_TRY(check_dims(W, H, N));
    IE_TRY(check_quant(quant, (int)N));
    const size_t npx = (size_t)W * H;
    const size_t slot = ie_max_encoded_bytes(W, H, N, 1);
    if (!huffman) {
        // Three-stage pipeline over chunks of ~128 MiB of pixels (double-buffered staging): chunk c+1 is copied in on stream_in
        // while chunk c is encoded on the session's stream (one launch of each kernel for the whole chunk) and the streams of
        // chunk c-1 leave on stream_out -- PCIe busy in both directions, the kernels hidden behind it.  The only host waits
        // are for a chunk's bit counts (the copies out are as long as the streams, not as long as their worst case).
        const uint32_t nc = (uint32_t)std::max<size_t>(1, std::min<size_t>(std::min<size_t>(count, 32), ((size_t)128 << 20) / npx));
        SessionLease lease;
        IE_TRY(lease.acquire(0, W, H, N, nc));
        ie_session *s = lease.get();
        IE_TRY(session_ensure_pipeline(s));
        IE_TRY(session_reserve(&s->d_in, &s->d_in_cap, 2 * npx * nc));
        IE_TRY(session_reserve(&s->d_out, &s->d_out_cap, 2 * slot * nc));
        IE_TRY(session_reserve(&s->d_tmp, &s->d_tmp_cap, 2 * 32 * sizeof(unsigned long long)));
        cudaStream_t st = s->stream;
        unsigned long long *d_bits = reinterpret_cast<unsigned long long *>(s->d_tmp);
        cudaEvent_t *ev_in = s->ev_in, *ev_done = s->ev_done, *ev_out = s->ev_in + 2;       // [2] each: staged, encoded, copied out
        const uint32_t nchunks = (count + nc - 1) / nc;
        int rc = IE_OK;
        for (uint32_t c = 0; c <= nchunks && rc == IE_OK; c++) {
            const int b = (int)(c & 1);
            if (c < nchunks) {
                const uint32_t first = c * nc, n = std::min(nc, count - first);
                if (c >= 2) IE_CUDA(cudaStreamWaitEvent(s->stream_in, ev_done[b], 0));        // chunk c-2 no longer reads this d_in half
                IE_CUDA(cudaMemcpyAsync(s->d_in + (size_t)b * npx * nc, raws + (size_t)first * npx, npx * n, cudaMemcpyHostToDevice, s->stream_in));
                IE_CUDA(cudaEventRecord(ev_in[b], s->stream_in));
                IE_CUDA(cudaStreamWaitEvent(st, ev_in[b], 0));
                if (c >= 2) IE_CUDA(cudaStreamWaitEvent(st, ev_out[b], 0));                   // chunk c-2's streams have left this d_out half
                rc = encode_images_dev(s, s->d_in + (size_t)b * npx * nc, npx, n, W, H, (int)N, quant, use_rle, 1, 1, 0, 0,
                                       s->d_out + (size_t)b * slot * nc, slot, slot, st, 0, 0, 0, reinterpret_cast<uint64_t *>(d_bits + b * 32));
                if (rc != IE_OK) break;
                IE_CUDA(cudaMemcpyAsync(s->h_pinned + b * 32, d_bits + b * 32, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
                IE_CUDA(cudaEventRecord(ev_done[b], st));
            }
            if (c >= 1) {
                const int pb = b ^ 1;
                const uint32_t first = (c - 1) * nc, n = std::min(nc, count - first);
                IE_CUDA(cudaEventSynchronize(ev_done[pb]));
                IE_CUDA(cudaStreamWaitEvent(s->stream_out, ev_done[pb], 0));
                for (uint32_t i = 0; i < n; i++) {
                    const size_t bytes = (size_t)((s->h_pinned[pb * 32 + i] + 7) / 8);
                    out_bytes[first + i] = bytes;
                    if (bytes > out_stride) { set_error("output slot too small"); rc = IE_ENOSPC; break; }
                    IE_CUDA(cudaMemcpyAsync(out + (size_t)(first + i) * out_stride, s->d_out + (size_t)pb * slot * nc + (size_t)i * slot, bytes,
                                            cudaMemcpyDeviceToHost, s->stream_out));
                }
                IE_CUDA(cudaEventRecord(ev_out[pb], s->stream_out));
            }
        }
        IE_CUDA(cudaStreamSynchronize(s->stream_in));
        IE_CUDA(cudaStreamSynchronize(st));
        IE_CUDA(cudaStreamSynchronize(s->stream_out));
        if (rc != IE_OK) return rc;
        return read_err_flag(s, st);
    }
    // Huffman-coded batch: the Huffman stage synchronises per image (host-built dictionary), one image after the other
    const uint32_t sub = (uint32_t)std::max<size_t>(1, std::min<size_t>(count, ((size_t)2 << 30) / (npx + slot)));
    SessionLease lease;
    IE_TRY(lease.acquire(0, W, H, N, sub));
    ie_session *s = lease.get();
    IE_TRY(session_reserve(&s->d_in, &s->d_in_cap, npx * sub));
    IE_TRY(session_reserve(&s->d_out, &s->d_out_cap, slot * sub));
    cudaStream_t st = s->stream;
    std::vector<unsigned long long> bits(sub);
    for (uint32_t first = 0; first < count; first += sub) {
        const uint32_t n = std::min(sub, count - first);
        IE_CUDA(cudaMemcpyAsync(s->d_in, raws + (size_t)first * npx, npx * n, cudaMemcpyHostToDevice, st));
        IE_TRY(encode_images_dev(s, s->d_in, npx, n, W, H, (int)N, quant, use_rle, 0, 1, 0, 0, s->d_out, slot, slot, st));
        IE_CUDA(cudaMemcpyAsync(bits.data(), s->d_counter, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        IE_TRY(read_err_flag(s, st));
        for (uint32_t i = 0; i < n; i++) {
            size_t bytes = (size_t)((bits[i] + 7) / 8);
            IE_TRY(session_reserve(&s->d_tmp, &s->d_tmp_cap, slot + 4096));
            size_t hb = 0;
            IE_TRY(ie_huffman_encode_dev(s, s->d_out + (size_t)i * slot, bytes, s->d_tmp, s->d_tmp_cap, &hb, st));
            out_bytes[first + i] = hb;
            if (hb > out_stride) { set_error("output slot too small"); return IE_ENOSPC; }
            IE_CUDA(cudaMemcpyAsync(out + (size_t)(first + i) * out_stride, s->d_tmp, hb, cudaMemcpyDeviceToHost, st));
            IE_CUDA(cudaStreamSynchronize(st));     // d_tmp is reused by the next image
        }
    }
    return IE_OK;
}

int ie_decode_images_dev(ie_session *s, const uint8_t *d_encs, size_t enc_stride, const size_t *enc_bytes, uint32_t count,
                         uint64_t start_bit, uint8_t *d_raws_out, size_t raw_stride, uint32_t *W, uint32_t *H, void *stream);

int ie_decode_images(const uint8_t *encs, size_t enc_stride, const size_t *enc_bytes, uint32_t count, uint32_t N, uint8_t *raws_out,
                     size_t raw_stride, uint32_t *W, uint32_t *H) {
    if (!encs || !enc_bytes || !raws_out || count == 0) { set_error("NULL/empty argument"); return IE_EINVAL; }
    if (N != 4 && N != 8) { set_error("block size must be 4 or 8"); return IE_EINVAL; }
    // streams of a batch are independent.  Plain streams go through the concurrent device batch in sub-batches;
    // Huffman-coded ones (first bit 1, Huffman.cpp:361-371) through the single-image path.
    bool plain = true;
    size_t maxb = 0;
    for (uint32_t i = 0; i < count; i++) { plain = plain && !(encs[(size_t)i * enc_stride] & 0x80); maxb = std::max(maxb, enc_bytes[i]); }
    if (!plain || maxb > enc_stride) {
        for (uint32_t i = 0; i < count; i++) {
            uint32_t w = 0, h = 0;
            IE_TRY(ie_decode_image(encs + (size_t)i * enc_stride, enc_bytes[i], N, raws_out + (size_t)i * raw_stride, raw_stride, &w, &h));
            if (W) *W = w;
            if (H) *H = h;
        }
        return IE_OK;
    }
    // Headers are parsed from the host copy of the streams (no read-back).  Three-stage pipeline over chunks of streams,
    // double-buffered: chunk c+1 goes in on stream_in while chunk c is decoded (streams of a chunk spread over the worker
    // sessions' streams, so the latency-bound parse kernels of one overlap the block decode of another) and the pixels of chunk
    // c-1 leave on stream_out.
    std::vector<ParsedHeader> hdrs(count);
    size_t maxpx = 0;
    for (uint32_t i = 0; i < count; i++) {
        parse_header(encs + (size_t)i * enc_stride, enc_bytes[i], 1, (int)N, hdrs[i], 0);
        IE_TRY(check_dims(hdrs[i].W, hdrs[i].H, N));
        const size_t px = (size_t)hdrs[i].W * hdrs[i].H;
        if (px > raw_stride) { set_error("raw_stride smaller than a decoded image"); return IE_ENOSPC; }
        maxpx = std::max(maxpx, px);
    }
    if (W) *W = hdrs[count - 1].W;
    if (H) *H = hdrs[count - 1].H;
    SessionLease lease;
    IE_TRY(lease.acquire(1, 0, 0, N, 2));
    ie_session *s = lease.get();
    IE_TRY(session_ensure_pipeline(s));
    cudaStream_t st = s->stream;
    const size_t es = (maxb + 16 + 15) / 16 * 16;                                   // device stride of a stream
    const size_t rs = (maxpx + 15) / 16 * 16;
    const uint32_t nc = (uint32_t)std::max<size_t>(1, std::min<size_t>(std::min<size_t>(count, 32), ((size_t)128 << 20) / rs));
    IE_TRY(session_reserve(&s->d_in, &s->d_in_cap, 2 * es * nc));
    IE_TRY(session_reserve(&s->d_out, &s->d_out_cap, 2 * rs * nc));
    const int nw = (int)std::min<uint32_t>(nc, ie_session::kDecodeWorkers);
    for (int k = 0; k < nw; k++) {
        if (!s->workers[k]) IE_TRY(ie_session_create(&s->workers[k], 1, 0, 0, N, 1));
        if (!s->ev_join[k]) IE_CUDA(cudaEventCreateWithFlags(&s->ev_join[k], cudaEventDisableTiming));
    }
    cudaEvent_t *ev_in = s->ev_in, *ev_done = s->ev_done, *ev_out = s->ev_in + 2;
    const uint32_t nchunks = (count + nc - 1) / nc;
    int rc = IE_OK;
    for (uint32_t c = 0; c <= nchunks && rc == IE_OK; c++) {
        const int b = (int)(c & 1);
        if (c < nchunks) {
            const uint32_t first = c * nc, n = std::min(nc, count - first);
            uint8_t *d_in = s->d_in + (size_t)b * es * nc, *d_out = s->d_out + (size_t)b * rs * nc;
            if (c >= 2) IE_CUDA(cudaStreamWaitEvent(s->stream_in, ev_done[b], 0));
            IE_CUDA(cudaMemcpy2DAsync(d_in, es, encs + (size_t)first * enc_stride, enc_stride, std::min(es, enc_stride), n, cudaMemcpyHostToDevice,
                                      s->stream_in));
            IE_CUDA(cudaEventRecord(ev_in[b], s->stream_in));
            for (int k = 0; k < nw; k++) {
                IE_CUDA(cudaStreamWaitEvent(s->workers[k]->stream, ev_in[b], 0));
                if (c >= 2) IE_CUDA(cudaStreamWaitEvent(s->workers[k]->stream, ev_out[b], 0));
            }
            for (uint32_t i = 0; i < n && rc == IE_OK; i++) {
                ie_session *w = s->workers[i % nw];
                rc = decode_image_dev(w, d_in + (size_t)i * es, enc_bytes[first + i], 1, (int)N, hdrs[first + i], d_out + (size_t)i * rs, rs, w->stream);
            }
            for (int k = 0; k < nw; k++) {                       // join on the session's stream
                IE_CUDA(cudaEventRecord(s->ev_join[k], s->workers[k]->stream));
                IE_CUDA(cudaStreamWaitEvent(st, s->ev_join[k], 0));
            }
            IE_CUDA(cudaEventRecord(ev_done[b], st));
        }
        if (c >= 1 && rc == IE_OK) {
            const int pb = b ^ 1;
            const uint32_t first = (c - 1) * nc, n = std::min(nc, count - first);
            IE_CUDA(cudaStreamWaitEvent(s->stream_out, ev_done[pb], 0));
            for (uint32_t i = 0; i < n; i++)
                IE_CUDA(cudaMemcpyAsync(raws_out + (size_t)(first + i) * raw_stride, s->d_out + (size_t)pb * rs * nc + (size_t)i * rs,
                                        (size_t)hdrs[first + i].W * hdrs[first + i].H, cudaMemcpyDeviceToHost, s->stream_out));
            IE_CUDA(cudaEventRecord(ev_out[pb], s->stream_out));
        }
    }
    IE_CUDA(cudaStreamSynchronize(s->stream_in));
    for (int k = 0; k < nw; k++) IE_CUDA(cudaStreamSynchronize(s->workers[k]->stream));
    IE_CUDA(cudaStreamSynchronize(st));
    IE_CUDA(cudaStreamSynchronize(s->stream_out));
    if (rc != IE_OK) return rc;
    for (int k = 0; k < nw; k++) IE_TRY(read_err_flag(s->workers[k], s->workers[k]->stream));
    return IE_OK;
}

int ie_decode_images_dev(ie_session *s, const uint8_t *d_encs, size_t enc_stride, const size_t *enc_bytes, uint32_t count,
                         uint64_t start_bit, uint8_t *d_raws_out, size_t raw_stride, uint32_t *W, uint32_t *H, void *stream) {
    if (!s || !d_encs || !enc_bytes || !d_raws_out || count == 0) { set_error("NULL/empty argument"); return IE_EINVAL; }
    if (enc_stride % 16) { set_error("enc_stride must be a multiple of 16"); return IE_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    const int N = (int)s->N;
    // all headers with one strided copy and one synchronisation
    constexpr size_t kHdr = 160;
    const size_t first = (size_t)(start_bit / 8);
    std::vector<uint8_t> hb((size_t)count * kHdr, 0);
    for (uint32_t i = 0; i < count; i++)
        if (first >= enc_bytes[i] || enc_bytes[i] > enc_stride) { set_error("start_bit beyond a stream / stream longer than enc_stride"); return IE_EFORMAT; }
    // kHdr bytes of every row (bounded by the stride, which the caller owns in full); each header is then parsed from the
    // bytes its own stream holds -- a short stream in the batch does not truncate its neighbours' headers
    const size_t n = std::min(kHdr, enc_stride - first);
    IE_CUDA(cudaMemcpy2DAsync(hb.data(), kHdr, d_encs + first, enc_stride, n, count, cudaMemcpyDeviceToHost, st));
    IE_CUDA(cudaStreamSynchronize(st));
    if (!s->ev_fork) IE_CUDA(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
    const int nw = (int)std::min<uint32_t>(count, ie_session::kDecodeWorkers);
    for (int k = 0; k < nw; k++) {
        if (!s->workers[k]) IE_TRY(ie_session_create(&s->workers[k], 1, 0, 0, (uint32_t)N, 1));
        if (!s->ev_join[k]) IE_CUDA(cudaEventCreateWithFlags(&s->ev_join[k], cudaEventDisableTiming));
    }
    IE_CUDA(cudaEventRecord(s->ev_fork, st));
    for (int k = 0; k < nw; k++) IE_CUDA(cudaStreamWaitEvent(s->workers[k]->stream, s->ev_fork, 0));
    int rc = IE_OK;
    for (uint32_t i = 0; i < count && rc == IE_OK; i++) {
        ie_session *w = s->workers[i % nw];
        ParsedHeader h;
        parse_header(hb.data() + (size_t)i * kHdr, std::min(n, enc_bytes[i] - first), (size_t)(start_bit % 8), N, h, 0);
        h.end_bit += first * 8;
        if (W) W[i] = h.W;
        if (H) H[i] = h.H;
        if ((size_t)h.W * h.H > raw_stride) { set_error("raw_stride smaller than a decoded image"); rc = IE_ENOSPC; break; }
        rc = decode_image_dev(w, d_encs + (size_t)i * enc_stride, enc_bytes[i], (size_t)start_bit, N, h, d_raws_out + (size_t)i * raw_stride,
                              raw_stride, w->stream);
    }
    for (int k = 0; k < nw; k++) {                           // join, also on errors: the caller's stream stays ordered
        cudaEventRecord(s->ev_join[k], s->workers[k]->stream);
        cudaStreamWaitEvent(st, s->ev_join[k], 0);
    }
    return rc;
}

int ie_decode_image_dev(ie_session *s, const uint8_t *d_enc, size_t enc_bytes, uint64_t start_bit, uint8_t *d_raw_out,
                        size_t raw_cap, uint32_t *W, uint32_t *H, void *stream) {
    if (!s || !d_enc || !d_raw_out) { set_error("NULL argument"); return IE_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    const int N = (int)s->N;
    // the header (<= 134 bytes) is parsed on the host
    uint8_t hb[160];
    const size_t first = (size_t)(start_bit / 8);
    if (first >= enc_bytes) { set_error("start_bit beyond the stream"); return IE_EFORMAT; }
    const size_t n = std::min(sizeof hb, enc_bytes - first);
    IE_CUDA(cudaMemcpyAsync(hb, d_enc + first, n, cudaMemcpyDeviceToHost, st));
    IE_CUDA(cudaStreamSynchronize(st));
    ParsedHeader h;
    parse_header(hb, n, (size_t)(start_bit % 8), N, h, 0);
    h.end_bit += first * 8;
    if (W) *W = h.W;
    if (H) *H = h.H;
    return decode_image_dev(s, d_enc, enc_bytes, (size_t)start_bit, N, h, d_raw_out, raw_cap, st);
}

int ie_parse_image_header(const uint8_t *bytes, size_t nbytes, uint64_t start_bit, uint32_t N, ie_image_header *out) {
    if (!bytes || !out || nbytes == 0) { set_error("NULL/empty argument"); return IE_EINVAL; }
    if (N != 4 && N != 8) { set_error("block size must be 4 or 8"); return IE_EINVAL; }
    ParsedHeader h;
    parse_header(bytes, nbytes, (size_t)start_bit, (int)N, h, 0);
    memset(out, 0, sizeof *out);
    out->block = N; out->width = h.W; out->height = h.H; out->use_rle = (uint32_t)h.use_rle;
    out->first_block_bit = (uint64_t)h.end_bit;
    for (uint32_t i = 0; i < N * N; i++) out->quant[i] = h.quant[i];
    return IE_OK;
}

int ie_decode_image_with_header_dev(ie_session *s, const ie_image_header *hdr, const uint8_t *d_enc, size_t enc_bytes, uint8_t *d_raw_out,
                                    size_t raw_cap, void *stream) {
    if (!s || !hdr || !d_enc || !d_raw_out) { set_error("NULL argument"); return IE_EINVAL; }
    if (hdr->block != s->N) { set_error("header block size differs from the session's"); return IE_EINVAL; }
    ParsedHeader h;
    memset(&h, 0, sizeof h);
    for (uint32_t i = 0; i < hdr->block * hdr->block; i++) h.quant[i] = hdr->quant[i];
    h.use_rle = (int)hdr->use_rle; h.W = hdr->width; h.H = hdr->height; h.end_bit = (size_t)hdr->first_block_bit;
    return decode_image_dev(s, d_enc, enc_bytes, 0, (int)s->N, h, d_raw_out, raw_cap, (cudaStream_t)stream);
}


// ---- sharded decode of one stream: every rank holds the stream, walks 1 / parts of it, decodes its own block rows ----------
namespace ie {
static int shard_decode_params(ie_session *s, const ie_image_header *hdr, const uint8_t *d_enc, size_t enc_bytes, DecodeParams &p, cudaStream_t st) {
    const int N = (int)s->N;
    const uint32_t W = hdr->width, H = hdr->height;
    IE_TRY(check_dims(W, H, N));
    if ((uintptr_t)d_enc % 16) { set_error("encoded stream must be 16-byte aligned (and readable up to its size rounded up to 4)"); return IE_EINVAL; }
    const unsigned nblocks = (W / N) * (H / N);
    IE_TRY(session_ensure_scan(s, 1, 1));
    IE_TRY(session_ensure_err(s));
    const size_t need_off = ((size_t)nblocks + 1) * sizeof(unsigned long long) + 64;
    if (s->block_off_cap < need_off) {
        if (s->d_block_off) IE_CUDA(cudaFree(s->d_block_off));
        s->d_block_off = nullptr; s->block_off_cap = 0;
        IE_CUDA(cudaMalloc(&s->d_block_off, need_off));
        s->block_off_cap = need_off;
    }
    if (!s->h_pinned) { set_error("session without pinned staging"); return IE_EINVAL; }
    unsigned long long *consts = s->h_pinned + 16;                    // pinned: the copy below is asynchronous
    consts[0] = (unsigned long long)enc_bytes * 8ull; consts[1] = (unsigned long long)hdr->first_block_bit;
    unsigned long long *d_consts = s->d_block_off + (nblocks + 1);
    IE_CUDA(cudaMemcpyAsync(d_consts, consts, 2 * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
    memset(&p, 0, sizeof p);
    p.enc = d_enc; p.enc_stride = 0; p.enc_bits = d_consts; p.start_bit = d_consts + 1;
    p.block_off = s->d_block_off; p.nblocks = nblocks; p.bx = W / N; p.N = N; p.use_rle = (int)hdr->use_rle;
    make_quant(p.quant, hdr->quant, N);
    make_k2(p.k2, hdr->quant, N);
    p.tab = (N == 8) ? s->dev->d_t8 : s->dev->d_t4;
    p.pitch = W; p.err = s->d_err;
    IE_TRY(session_reserve(&s->d_parse, &s->parse_cap, parse_scratch_bytes(enc_bytes, N)));
    return IE_OK;
}
}  // namespace ie

size_t ie_decode_shard_spec_bytes(size_t enc_bytes, uint32_t N, uint32_t parts, size_t *chunk_bytes) {
    if ((N != 4 && N != 8) || parts == 0) return 0;
    const ShardedParseGeom g = sharded_parse_geom(enc_bytes, (int)N, parts);
    if (chunk_bytes) *chunk_bytes = g.chunk_bytes;
    return g.spec_bytes;
}

int ie_decode_image_shard_begin_dev(ie_session *s, const ie_image_header *hdr, const uint8_t *d_enc, size_t enc_bytes, uint32_t part,
                                    uint32_t parts, uint8_t *d_spec, void *stream) {
    if (!s || !hdr || !d_enc || !d_spec) { set_error("NULL argument"); return IE_EINVAL; }
    if (hdr->block != s->N) { set_error("header block size differs from the session's"); return IE_EINVAL; }
    if (parts == 0 || parts > 64 || part >= parts) { set_error("part / parts out of range (1 .. 64 parts)"); return IE_EINVAL; }
    DecodeParams p;
    IE_TRY(shard_decode_params(s, hdr, d_enc, enc_bytes, p, (cudaStream_t)stream));
    return launch_parse_walk_part(p, enc_bytes * 8, s->d_parse, d_spec, part, parts, (cudaStream_t)stream);
}

int ie_decode_image_shard_end_dev(ie_session *s, const ie_image_header *hdr, const uint8_t *d_enc, size_t enc_bytes, uint32_t parts,
                                  uint8_t *d_spec, uint32_t block_row0, uint32_t block_row1, uint8_t *d_rows_out, size_t out_cap, void *stream) {
    if (!s || !hdr || !d_enc || !d_spec || !d_rows_out) { set_error("NULL argument"); return IE_EINVAL; }
    if (hdr->block != s->N) { set_error("header block size differs from the session's"); return IE_EINVAL; }
    if (parts == 0 || parts > 64) { set_error("parts out of range (1 .. 64)"); return IE_EINVAL; }
    const uint32_t N = s->N;
    if (block_row0 >= block_row1 || block_row1 > hdr->height / N) { set_error("block rows out of range"); return IE_EINVAL; }
    if ((size_t)(block_row1 - block_row0) * N * hdr->width > out_cap) { set_error("decoded rows do not fit the output buffer"); return IE_ENOSPC; }
    DecodeParams p;
    cudaStream_t st = (cudaStream_t)stream;
    IE_TRY(shard_decode_params(s, hdr, d_enc, enc_bytes, p, st));
    p.block_base = block_row0 * (hdr->width / N); p.block_end = block_row1 * (hdr->width / N);
    IE_TRY(launch_parse_finish_range(p, enc_bytes * 8, s->d_parse, d_spec, parts, p.block_base, p.block_end, st));
    // the kernels address pixels by absolute block row: the band's buffer starts at row block_row0 * N
    p.out = d_rows_out - (size_t)block_row0 * N * hdr->width;
    return launch_decode_blocks(p, 1, st);
}

int ie_decode_image(const uint8_t *enc, size_t enc_bytes, uint32_t N, uint8_t *raw_out, size_t raw_cap, uint32_t *W, uint32_t *H) {
    if (!enc || !raw_out || enc_bytes == 0) { set_error("NULL/empty argument"); return IE_EINVAL; }
    if (N != 4 && N != 8) { set_error("block size must be 4 or 8"); return IE_EINVAL; }
    SessionLease lease;
    IE_TRY(lease.acquire(1, 0, 0, N, 1));
    ie_session *s = lease.get();
    cudaStream_t st = s->stream;
    IE_TRY(session_reserve(&s->d_in, &s->d_in_cap, enc_bytes + 16));
    IE_CUDA(cudaMemcpyAsync(s->d_in, enc, enc_bytes, cudaMemcpyHostToDevice, st));
    const uint8_t *d_plain = s->d_in;
    size_t plain_bytes = enc_bytes;
    uint64_t start_bit = 1;                                         // '0' = no Huffman table (Huffman.cpp:361-371)
    if (enc[0] & 0x80) {
        size_t cap = enc_bytes * 8 + 64;                            // a code is at least 1 bit per byte
        IE_TRY(session_reserve(&s->d_tmp, &s->d_tmp_cap, cap));
        IE_TRY(ie_huffman_decode_dev(s, s->d_in, enc_bytes, s->d_tmp, s->d_tmp_cap, &plain_bytes, &start_bit, st));
        d_plain = s->d_tmp;
    }
    uint32_t w = 0, h = 0;
    // decode straight into a device buffer sized after the header is known
    uint8_t hb[160];
    const size_t n = std::min(sizeof hb, plain_bytes);
    ParsedHeader ph;
    if (d_plain == s->d_in) {
        parse_header(enc, n, (size_t)start_bit, (int)N, ph, 0);     // plain stream: the header is in the caller's buffer
    } else {
        IE_CUDA(cudaMemcpyAsync(hb, d_plain, n, cudaMemcpyDeviceToHost, st));
        IE_CUDA(cudaStreamSynchronize(st));
        parse_header(hb, n, (size_t)start_bit, (int)N, ph, 0);
    }
    w = ph.W; h = ph.H;
    if (W) *W = w;
    if (H) *H = h;
    IE_TRY(check_dims(w, h, N));
    if ((size_t)w * h > raw_cap) { set_error("raw_out too small"); return IE_ENOSPC; }
    IE_TRY(session_reserve(&s->d_out, &s->d_out_cap, (size_t)w * h));
    IE_TRY(decode_image_dev(s, d_plain, plain_bytes, (size_t)start_bit, (int)N, ph, s->d_out, s->d_out_cap, st, raw_out));
    const int rc = read_err_flag(s, st);                           // synchronises st: every stripe has been decoded
    IE_CUDA(cudaStreamSynchronize(s->stream_out));                 // ... and has arrived
    return rc;
}

}  // extern "C"
