// ie_session: scratch that one problem shape needs (look-back states, hand-off records, counters, staging).
#pragma once
#include "common.cuh"
#include "pack.cuh"
#include "encode_image.cuh"

struct ie_session {
    int kind = 0;
    uint32_t W = 0, H = 0, N = 0, frames = 1;
    int device = -1;
    uint32_t header_height = 0;                   // 0: use the height of the call
    // GOP shard of a longer clip (multi-GPU): frame count announced by the header (0: the frames of the call) and whether
    // this shard writes the header at all (the first shard does)
    uint32_t header_frames = 0;
    int video_no_header = 0;
    ie::DeviceState *dev = nullptr;

    // scan scratch, sized for `images` streams of `max_tiles` tiles (zeroed once; self-cleaning afterwards)
    unsigned images = 1;
    unsigned max_tiles = 0;
    unsigned long long *d_tile_state = nullptr;
    ie::TileBoundary *d_bnd = nullptr;
    unsigned *d_ticket = nullptr;
    unsigned long long *d_counter = nullptr;      // [images] stream bit counters
    int *d_err = nullptr;
    unsigned epoch = 0;

    // tile scratch of the encoder (packed tile images, per-tile bit totals and stream offsets)
    uint8_t *d_tile_scratch = nullptr; size_t tile_scratch_cap = 0;
    uint8_t *d_tile_meta = nullptr;    size_t tile_meta_cap = 0;

    // batch decode (ie_decode_images_dev): worker sessions with their own streams, so that the latency-bound parse kernels
    // of one stream overlap the block decode of another; fork/join events keep the caller's stream semantics
    static constexpr int kDecodeWorkers = 4;
    ie_session *workers[kDecodeWorkers] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[kDecodeWorkers] = {};

    // split encode (ie_encode_image_begin_dev / _end_dev): what `begin` left for `end`
    bool split_pending = false;
    ie::EncodeParams split_params;
    ie::HeaderParam split_hdr;

    // decode scratch
    unsigned long long *d_block_off = nullptr;    // [images * nblocks (+1)]
    size_t block_off_cap = 0;

    uint8_t *d_parse = nullptr;                   // transfer-function tables of the parallel parser
    size_t parse_cap = 0;

    // generic device scratch (Huffman stage, video)
    uint8_t *d_scratch = nullptr;
    size_t scratch_cap = 0;

    // staging for the host-buffer entry points
    uint8_t *d_in = nullptr;  size_t d_in_cap = 0;
    uint8_t *d_out = nullptr; size_t d_out_cap = 0;
    uint8_t *d_tmp = nullptr; size_t d_tmp_cap = 0;
    unsigned long long *h_pinned = nullptr;       // small pinned read-back area (64 u64)
    uint8_t *h_huff = nullptr;                    // pinned staging of the Huffman stage: dictionary header, codes, first bit
    void *huff_ctx = nullptr;                     // argument block of the stage's host callback (huffman.cu), malloc'ed
    void *video_hooks = nullptr;                  // ie_encode_video's copy / compute pipeline (api_video.cu), set during one call
    void *video_dec_hooks = nullptr;              // ie_decode_video's (frames of a finished GOP batch go down next to later batches)
    cudaStream_t stream = nullptr;                // used by the host-buffer entry points

    // copy/compute pipeline of the host-buffer image entry points: pixels come in by stripes on stream_in, the encoded
    // bytes leave on stream_out while later stripes are still being copied in / encoded
    static constexpr int kMaxStripes = 32;
    cudaStream_t stream_in = nullptr, stream_out = nullptr;
    cudaEvent_t ev_in[kMaxStripes] = {}, ev_done[kMaxStripes] = {};
    // second compute stream of the video entry points (half of a batch's GOPs / the reconstruction next to the frame chain):
    // never the copy streams, or kernels would queue behind every copy already enqueued there
    cudaStream_t stream_aux = nullptr;
    cudaEvent_t ev_aux_fork = nullptr, ev_aux_join = nullptr;

    ie::ScanState scan_state() {
        ie::ScanState st;
        st.tile_state = d_tile_state;
        st.bnd = d_bnd;
        st.ticket = d_ticket;
        epoch = (epoch + 1) & 0xFFFFFFu;
        if (epoch == 0) epoch = 1;                // 0 is the "never written" value of a fresh array
        st.epoch = epoch;
        return st;
    }
};

namespace ie {
int session_reserve(uint8_t **p, size_t *cap, size_t need);
int session_ensure_scan(ie_session *s, unsigned images, unsigned tiles);
int session_ensure_err(ie_session *s);
}  // namespace ie
