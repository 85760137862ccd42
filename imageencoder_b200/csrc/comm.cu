// Multi-GPU exchange for sharded streams, behind the C-ABI (SURVEY 8e, 8b "ie_init(n_gpus) owning the communicator").
//
// One ie_comm per rank (one process per GPU, or several GPUs in one process).  A rank owns a MAILBOX in its HBM that its
// peers write with plain stores over NVLink (peer-mapped memory: CUDA IPC between processes, peer access inside one):
//   * shard bit totals: the only exchange of a block-row sharded encode (ImageEncoder.cpp:121-138 made parallel).  A
//     one-CTA kernel between the tile kernel and the copy-out stores this rank's u64 into every peer's mailbox and spins
//     until the peers' values are there -- no NCCL launch on the critical path (the all-gather it replaces cost 16-25 us of
//     a 110 us step).  Every word carries its epoch, so no fence and no reset is needed.
//   * stitch: the ONE output stream of a sharded encode (the reference's writer yields one buffer, ImageBase.cpp:315-336)
//     is assembled on the root GPU by the ranks themselves: every rank stores its 128-bit chunks straight into the root's
//     buffer at their final position; the chunk two neighbouring shards share travels through the right neighbour's mailbox
//     and is written once, complete -- no atomics on the stream, no pre-zeroing, no host round trip.
#include <unistd.h>

#include <cstring>
#include <vector>

#include "api_internal.cuh"

namespace ie {

constexpr int kMaxRanks = 16;
constexpr int kTailSlotWords = 4;                  // u64 words per tail slot: payload (2) + tag (1) + pad

// mailbox layout (u64 words): [2][kMaxRanks] totals (by epoch parity), then [2] tail slots of kTailSlotWords
constexpr size_t kMailboxWords = 2 * kMaxRanks + 2 * kTailSlotWords;

struct CommBlob {
    int pid, device, rank, has_stitch;
    void *mailbox, *stitch;
    size_t stitch_bytes;
    cudaIpcMemHandle_t mailbox_h, stitch_h;
};

struct CommDev {                                   // what the kernels need (passed by value)
    unsigned long long *peer_box[kMaxRanks];       // peer-mapped mailboxes (own one included)
    unsigned long long *box;                       // own mailbox
    int rank, world;
    unsigned epoch;
};

__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

constexpr long long kSpinLimit = 4000000000ll;     // ~2 s of SM clocks: a peer that never arrives is an error, not a hang

// thread r: my total -> rank r's mailbox; then: rank r's total <- my mailbox.  totals[world] for ie_encode_image_end_dev.
__global__ void __launch_bounds__(32) shard_exchange_kernel(const CommDev c, const unsigned long long *d_total, unsigned long long *totals,
                                                           int *err) {
    pdl_wait();
    const int r = threadIdx.x;
    if (r >= c.world) return;
    const unsigned long long tag = (unsigned long long)(c.epoch & 0xffffu) << 48;
    const unsigned slot = (c.epoch & 1u) * kMaxRanks;
    st_sys_u64(c.peer_box[r] + slot + c.rank, tag | (*d_total & 0xffffffffffffull));
    const long long t0 = clock64();
    unsigned long long v;
    while (((v = ld_sys_u64(c.box + slot + r)) >> 48) != (c.epoch & 0xffffu)) {
        if (clock64() - t0 > kSpinLimit) { if (err) atomicExch(err, IE_ECUDA); v = 0; break; }
        __nanosleep(100);
    }
    totals[r] = v & 0xffffffffffffull;
}

// The three one-CTA steps between the tile kernel and the copy-out of a shard, in ONE launch (they were tile_totals_kernel,
// shard_exchange_kernel and stream_init_shard_kernel: two launch gaps of a 70 us step at 8 GPUs): this shard's bit total from its
// tile totals (+ the header on rank 0) -> every peer's mailbox; the peers' totals <- this rank's mailbox; the shard's first bit;
// the stream prefix (first % 128 zero bits, then the header on rank 0) and the bit counters the copy-out kernel starts from.
__global__ void __launch_bounds__(256) shard_totals_exchange_init_kernel(const CommDev c, const unsigned *__restrict__ tile_bits, unsigned ntiles,
                                                                        unsigned long long add, unsigned long long *totals, uint8_t *out,
                                                                        HeaderParam hdr, unsigned long long *counter, unsigned long long *bit_base,
                                                                        unsigned long long *first_out, int *err) {
    pdl_wait();
    __shared__ unsigned long long s_part[8];
    __shared__ unsigned long long s_tot[kMaxRanks];
    unsigned long long sum = 0;
    for (unsigned i = threadIdx.x; i < ntiles; i += 256) sum += tile_bits[i];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    if ((threadIdx.x & 31u) == 0) s_part[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x < (unsigned)c.world) {
        unsigned long long mine = add;
        for (int w = 0; w < 8; w++) mine += s_part[w];
        const int r = threadIdx.x;
        const unsigned long long tag = (unsigned long long)(c.epoch & 0xffffu) << 48;
        const unsigned slot = (c.epoch & 1u) * kMaxRanks;
        st_sys_u64(c.peer_box[r] + slot + c.rank, tag | (mine & 0xffffffffffffull));
        const long long t0 = clock64();
        unsigned long long v;
        while (((v = ld_sys_u64(c.box + slot + r)) >> 48) != (c.epoch & 0xffffu)) {
            if (clock64() - t0 > kSpinLimit) { if (err) atomicExch(err, IE_ECUDA); v = 0; break; }
            __nanosleep(100);
        }
        v &= 0xffffffffffffull;
        totals[r] = v;
        s_tot[r] = v;
    }
    __syncthreads();
    unsigned long long first = 0;
    for (int i = 0; i < c.rank; i++) first += s_tot[i];
    const unsigned first_bit = (unsigned)(first % 128);
    const unsigned total = first_bit + hdr.bits;
    const unsigned nwords = ((total + 127) / 128) * 4;
    unsigned *o = reinterpret_cast<unsigned *>(out);
    for (unsigned i = threadIdx.x; i < max(nwords, 4u); i += 256) {
        const long long hb = (long long)i * 32 - (long long)first_bit;
        const int sh = (int)(((hb % 32) + 32) % 32);
        const long long wi = (hb - sh) / 32;               // floor division
        const unsigned hi = (wi >= 0 && wi < kHdrWordsMax) ? hdr.words[wi] : 0u;
        const unsigned lo = (wi + 1 >= 0 && wi + 1 < kHdrWordsMax) ? hdr.words[wi + 1] : 0u;
        const unsigned v = sh ? ((hi << sh) | (lo >> (32 - sh))) : hi;
        o[i] = __byte_perm(v, 0, 0x0123);
    }
    if (threadIdx.x == 0) {
        counter[0] = total;
        bit_base[0] = total;
        if (first_out) *first_out = first;
    }
}

// This rank's chunks of the global stream -> the root's buffer.  d_shard: the rank's bytes from the chunk that holds its first
// bit (what ie_encode_image_end_dev leaves); *d_bits = (first % 128) + shard bits; *d_first = first bit in the global stream.
__global__ void __launch_bounds__(256) stitch_kernel(const CommDev c, const uint4 *__restrict__ d_shard, const unsigned long long *d_bits,
                                                     const unsigned long long *d_first, uint4 *root, size_t root_chunks, int *err) {
    const unsigned long long bits = *d_bits, first = *d_first;
    const unsigned long long c0 = first >> 7;
    const unsigned long long nchunks = (bits + 127) >> 7;
    const bool head_shared = (first & 127ull) != 0 && c.rank > 0;
    const bool tail_shared = (bits & 127ull) != 0 && c.rank + 1 < c.world;
    const unsigned long long tag = (unsigned long long)(c.epoch & 0xffffu) << 48 | 1ull;
    const unsigned tslot = 2 * kMaxRanks + (c.epoch & 1u) * kTailSlotWords;
    for (unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; k < nchunks; k += (unsigned long long)gridDim.x * blockDim.x) {
        uint4 v = d_shard[k];
        if (k == 0 && head_shared) {
            // the left neighbour's bits of this chunk arrive in my mailbox (payload first, then the tag)
            const long long t0 = clock64();
            while (ld_sys_u64(c.box + tslot + 2) != tag) {
                if (clock64() - t0 > kSpinLimit) { if (err) atomicExch(err, IE_ECUDA); break; }
                __nanosleep(100);
            }
            __threadfence_system();
            const unsigned long long lo = ld_sys_u64(c.box + tslot), hi = ld_sys_u64(c.box + tslot + 1);
            v.x |= (unsigned)lo; v.y |= (unsigned)(lo >> 32); v.z |= (unsigned)hi; v.w |= (unsigned)(hi >> 32);
        }
        if (k + 1 == nchunks && tail_shared) {
            // my bits of the chunk I share with the right neighbour: it completes and writes the chunk
            unsigned long long *dst = c.peer_box[c.rank + 1] + tslot;
            st_sys_u64(dst, (unsigned long long)v.x | ((unsigned long long)v.y << 32));
            st_sys_u64(dst + 1, (unsigned long long)v.z | ((unsigned long long)v.w << 32));
            __threadfence_system();
            st_sys_u64(dst + 2, tag);
            continue;
        }
        if (c0 + k >= root_chunks) { if (err) atomicExch(err, IE_ENOSPC); continue; }
        root[c0 + k] = v;
    }
}

}  // namespace ie

using namespace ie;

struct ie_comm {
    int rank = 0, world = 1, device = -1;
    unsigned epoch = 0;
    unsigned long long *box = nullptr;                       // own mailbox (device)
    unsigned long long *peer_box[kMaxRanks] = {};
    bool opened_box[kMaxRanks] = {};
    uint8_t *stitch = nullptr;                               // root only
    size_t stitch_bytes = 0;
    uint8_t *root_stitch = nullptr;                          // mapped pointer to the root's buffer (root: == stitch)
    size_t root_stitch_bytes = 0;
    bool opened_stitch = false;
    unsigned long long *d_totals = nullptr;                  // [2][kMaxRanks + 1] by epoch parity: the exchanged totals, this rank's own total
    unsigned long long *totals_of_epoch() const { return d_totals + (epoch & 1u) * (kMaxRanks + 1); }
    bool connected = false;

    CommDev dev() const {
        CommDev d;
        for (int i = 0; i < kMaxRanks; i++) d.peer_box[i] = peer_box[i];
        d.box = box; d.rank = rank; d.world = world; d.epoch = epoch;
        return d;
    }
};

extern "C" {

int ie_comm_create(ie_comm **out, int rank, int world, size_t stitch_bytes) {
    if (!out || rank < 0 || world < 1 || rank >= world) { set_error("bad rank / world"); return IE_EINVAL; }
    if (world > kMaxRanks) { set_error("at most 16 ranks (one box)"); return IE_EINVAL; }
    DeviceState *ds = nullptr;
    IE_TRY(get_device_state(&ds));
    ie_comm *c = new ie_comm();
    c->rank = rank; c->world = world; c->device = ds->device;
    IE_CUDA(cudaMalloc(&c->box, kMailboxWords * sizeof(unsigned long long)));
    IE_CUDA(cudaMemset(c->box, 0, kMailboxWords * sizeof(unsigned long long)));
    IE_CUDA(cudaMalloc(&c->d_totals, 2 * (kMaxRanks + 1) * sizeof(unsigned long long)));
    IE_CUDA(cudaMemset(c->d_totals, 0, 2 * (kMaxRanks + 1) * sizeof(unsigned long long)));
    if (rank == 0 && stitch_bytes) {
        c->stitch_bytes = (stitch_bytes + 15) / 16 * 16 + 16;
        IE_CUDA(cudaMalloc(&c->stitch, c->stitch_bytes));
    }
    IE_CUDA(cudaDeviceSynchronize());
    c->peer_box[rank] = c->box;
    if (world == 1) { c->root_stitch = c->stitch; c->root_stitch_bytes = c->stitch_bytes; c->connected = true; }
    *out = c;
    return IE_OK;
}

size_t ie_comm_handle_bytes(void) { return sizeof(CommBlob); }

int ie_comm_export(ie_comm *c, void *blob_out) {
    if (!c || !blob_out) { set_error("NULL argument"); return IE_EINVAL; }
    CommBlob b;
    memset(&b, 0, sizeof b);
    b.pid = (int)getpid(); b.device = c->device; b.rank = c->rank;
    b.mailbox = c->box; b.stitch = c->stitch; b.stitch_bytes = c->stitch_bytes; b.has_stitch = c->stitch != nullptr;
    IE_CUDA(cudaIpcGetMemHandle(&b.mailbox_h, c->box));
    if (c->stitch) IE_CUDA(cudaIpcGetMemHandle(&b.stitch_h, c->stitch));
    memcpy(blob_out, &b, sizeof b);
    return IE_OK;
}

int ie_comm_connect(ie_comm *c, const void *blobs) {
    if (!c || !blobs) { set_error("NULL argument"); return IE_EINVAL; }
    const CommBlob *bs = reinterpret_cast<const CommBlob *>(blobs);
    const int me = (int)getpid();
    for (int r = 0; r < c->world; r++) {
        CommBlob b;
        memcpy(&b, &bs[r], sizeof b);
        if (b.rank != r) { set_error("handle blobs must be in rank order"); return IE_EINVAL; }
        if (r == c->rank) {
            if (r == 0) { c->root_stitch = c->stitch; c->root_stitch_bytes = c->stitch_bytes; }
            continue;
        }
        if (b.pid == me) {
            // several GPUs in one process: plain peer access
            int can = 0;
            IE_CUDA(cudaDeviceCanAccessPeer(&can, c->device, b.device));
            if (!can) { set_error("no peer access between the GPUs of this communicator"); return IE_ENODEVICE; }
            cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
            cudaGetLastError();
            c->peer_box[r] = reinterpret_cast<unsigned long long *>(b.mailbox);
            if (r == 0 && b.has_stitch) { c->root_stitch = reinterpret_cast<uint8_t *>(b.stitch); c->root_stitch_bytes = b.stitch_bytes; }
        } else {
            void *p = nullptr;
            IE_CUDA(cudaIpcOpenMemHandle(&p, b.mailbox_h, cudaIpcMemLazyEnablePeerAccess));
            c->peer_box[r] = reinterpret_cast<unsigned long long *>(p);
            c->opened_box[r] = true;
            if (r == 0 && b.has_stitch) {
                IE_CUDA(cudaIpcOpenMemHandle(&p, b.stitch_h, cudaIpcMemLazyEnablePeerAccess));
                c->root_stitch = reinterpret_cast<uint8_t *>(p);
                c->root_stitch_bytes = b.stitch_bytes;
                c->opened_stitch = true;
            }
        }
    }
    c->connected = true;
    return IE_OK;
}

void ie_comm_destroy(ie_comm *c) {
    if (!c) return;
    for (int r = 0; r < kMaxRanks; r++)
        if (c->opened_box[r]) cudaIpcCloseMemHandle(c->peer_box[r]);
    if (c->opened_stitch) cudaIpcCloseMemHandle(c->root_stitch);
    cudaFree(c->box); cudaFree(c->d_totals); cudaFree(c->stitch);
    delete c;
}

int ie_comm_exchange_totals_dev(ie_comm *c, const uint64_t *d_total, uint64_t *d_totals_out, void *stream) {
    if (!c || !d_total) { set_error("NULL argument"); return IE_EINVAL; }
    if (!c->connected) { set_error("ie_comm_connect has not run"); return IE_EINVAL; }
    c->epoch = (c->epoch + 1) & 0xffffu;
    if (c->epoch == 0) c->epoch = 2;                         // 0 is the value of a fresh mailbox; keep the parity sequence
    unsigned long long *totals = d_totals_out ? reinterpret_cast<unsigned long long *>(d_totals_out) : c->totals_of_epoch();
    IE_CUDA(launch_pdl(shard_exchange_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, c->dev(),
                       reinterpret_cast<const unsigned long long *>(d_total), totals, (int *)nullptr));
    count_launch();
    return IE_OK;
}

int ie_encode_image_shard_dev(ie_session *s, ie_comm *c, const uint8_t *d_raw, uint32_t W, uint32_t H_shard, uint32_t H_total,
                              const uint16_t *quant, int use_rle, int lead_bit, uint8_t *d_out, size_t out_cap,
                              uint64_t *d_out_bits, uint64_t *d_first_bit, void *stream) {
    if (!s || !c || !d_raw || !d_out) { set_error("NULL argument"); return IE_EINVAL; }
    if (H_shard == 0) { set_error("every rank needs at least one block row"); return IE_EINVAL; }
    if (!c->connected) { set_error("ie_comm_connect has not run"); return IE_EINVAL; }
    if ((uintptr_t)d_out % 16) { set_error("stream buffers must be 16-byte aligned"); return IE_EINVAL; }
    IE_TRY(ie_session_set_header_height(s, H_total));
    cudaStream_t st = (cudaStream_t)stream;
    // tile kernel: the shard's packed tiles stay in the session's scratch (split encode, as ie_encode_image_begin_dev)
    IE_TRY(encode_images_dev(s, d_raw, 0, 1, W, H_shard, (int)s->N, quant, use_rle, lead_bit, c->rank == 0, 0, 0, nullptr, 0, 0, st, 0, 0, 1));
    if (out_cap < ((size_t)128 + s->split_hdr.bits + 127) / 128 * 16) { set_error("output buffer too small for the header"); return IE_ENOSPC; }
    // Two calls may be in flight at once (a caller alternating two sessions / streams so that one shard's copy-out overlaps
    // the next tile kernel): mailbox slots and the totals arrays go by the parity of the call's epoch.
    c->epoch = (c->epoch + 1) & 0xffffu;
    if (c->epoch == 0) c->epoch = 2;                         // 0 is the value of a fresh mailbox; keep the parity sequence
    EncodeParams p = s->split_params;
    s->split_pending = false;
    p.out = d_out; p.out_stride = 0; p.out_cap = out_cap;
    p.out_bits = reinterpret_cast<unsigned long long *>(d_out_bits);          // written by the copy-out kernel itself
    IE_CUDA(launch_pdl(shard_totals_exchange_init_kernel, dim3(1), dim3(256), 0, st, c->dev(), (const unsigned *)p.tile_bits, p.tiles_per_image,
                       (unsigned long long)s->split_hdr.bits, c->totals_of_epoch(), d_out, s->split_hdr, s->d_counter, p.bit_base,
                       reinterpret_cast<unsigned long long *>(d_first_bit), s->d_err));
    count_launch();
    return launch_tile_copyout(p, 1, st);
}

int ie_comm_stitch_dev(ie_comm *c, const uint8_t *d_shard, const uint64_t *d_bits, const uint64_t *d_first_bit, void *stream) {
    if (!c || !d_shard || !d_bits || !d_first_bit) { set_error("NULL argument"); return IE_EINVAL; }
    if (!c->connected || !c->root_stitch) { set_error("no stitch buffer: create rank 0's communicator with stitch_bytes > 0"); return IE_EINVAL; }
    if ((uintptr_t)d_shard % 16) { set_error("shard buffer must be 16-byte aligned"); return IE_EINVAL; }
    // (uses the epoch of the preceding exchange: one stitch per exchange)
    stitch_kernel<<<dim3(296), dim3(256), 0, (cudaStream_t)stream>>>(c->dev(), reinterpret_cast<const uint4 *>(d_shard),
                                                                  reinterpret_cast<const unsigned long long *>(d_bits),
                                                                  reinterpret_cast<const unsigned long long *>(d_first_bit),
                                                                  reinterpret_cast<uint4 *>(c->root_stitch), c->root_stitch_bytes / 16, nullptr);
    count_launch();
    IE_CUDA(cudaGetLastError());
    return IE_OK;
}

int ie_comm_totals_dev(ie_comm *c, const uint64_t **d_totals) {
    if (!c || !d_totals) { set_error("NULL argument"); return IE_EINVAL; }
    *d_totals = reinterpret_cast<const uint64_t *>(c->totals_of_epoch());
    return IE_OK;
}

int ie_comm_copy_totals(ie_comm *c, uint64_t *dst, void *stream) {
    if (!c || !dst) { set_error("NULL argument"); return IE_EINVAL; }
    IE_CUDA(cudaMemcpyAsync(dst, c->totals_of_epoch(), (size_t)c->world * sizeof(uint64_t), cudaMemcpyDefault, (cudaStream_t)stream));
    return IE_OK;
}

int ie_comm_stitched_stream(ie_comm *c, uint8_t **d_stream, size_t *capacity) {
    if (!c || !d_stream) { set_error("NULL argument"); return IE_EINVAL; }
    if (c->rank != 0 || !c->stitch) { set_error("only rank 0 holds the stitched stream"); return IE_EINVAL; }
    *d_stream = c->stitch;
    if (capacity) *capacity = c->stitch_bytes;
    return IE_OK;
}

int ie_comm_copy_stitched(ie_comm *c, void *dst, size_t nbytes, void *stream) {
    if (!c || !dst) { set_error("NULL argument"); return IE_EINVAL; }
    if (c->rank != 0 || !c->stitch) { set_error("only rank 0 holds the stitched stream"); return IE_EINVAL; }
    if (nbytes > c->stitch_bytes) { set_error("more bytes than the stitch buffer holds"); return IE_EINVAL; }
    IE_CUDA(cudaMemcpyAsync(dst, c->stitch, nbytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return IE_OK;
}

}  // extern "C"
