#pragma once
#include "common.cuh"

namespace ie {

struct DecodeParams {
    const uint8_t *enc;                     // device, plain stream(s)
    size_t enc_stride;                      // bytes between the streams of a batch
    const unsigned long long *enc_bits;     // [images] device: size of each stream in bits (bytes * 8)
    const unsigned long long *start_bit;    // [images] device: first block's bit (just after the header)
    unsigned long long *block_off;          // [images][nblocks + 1] device
    unsigned nblocks;                       // per image
    unsigned bx;
    int N;
    int use_rle;
    QuantParam quant;
    float k2[kMaxNN];                       // float(Q[u][v] * C(u)C(v)): dequantisation folded into the fast inverse transform
    const BlockTables *tab;
    uint8_t *out;                           // device, decoded pixels
    size_t out_stride;
    size_t pitch;
    int *err;
    // video (Frame.cpp:47-127): the frame's first bit comes from a device-resident cursor that the parse kernel advances
    unsigned long long *cursor;             // device; NULL for images
    unsigned skip_bits;                     // bits between the cursor and the first block (the mvec section of a P-frame)
    int add_mode;                           // 1: pixel = (u8)clamp(cur + (X + 128))   (Block.cpp:110-119, P-frames)
    // launch_decode_blocks decodes blocks [block_base, block_end) of every image; block_end = 0 means nblocks (stripes of the
    // host entry point: a stripe's pixels leave for the host while the next stripe is decoded)
    unsigned block_base, block_end;
    // add_mode with the motion-compensated copy folded in (whole-stream video decode): the prediction of a block is read from
    // the reference frame at the MacroBlock's clamped coordinate instead of from the output (where mc_copy_kernel would have
    // put it).  mc_coord: [images][nmb][2] shorts (x, y), NULL = predict from the output; the reference frame of image i
    // starts ref_delta bytes before its own pixels; mbx = MacroBlocks per row.
    const short *mc_coord;
    size_t ref_delta;
    unsigned mbx, mc_stride;
    size_t block_off_stride;                // entries between the offset arrays of consecutive images; 0 = nblocks + 1
};

// ---- whole-stream video parse (parse.cu) ----
constexpr unsigned kVMaxPieces = 8;
struct VFramePiece {
    unsigned kind;          // 0 HEAD: pool[a .. a + b) are the starts of blocks idx_base ..; 1 SPEC: groups a .. b (inclusive)
    unsigned a, b;
    unsigned idx_base;      // index (within the frame) of the piece's first block
    unsigned pm;            // SPEC: pq[a].x
};
struct VFrameRec {
    unsigned long long first;      // first bit of the frame's first block (a P-frame's motion vectors end here)
    unsigned long long end;        // first bit after the frame's last block
    unsigned npieces;
    VFramePiece piece[kVMaxPieces];
};
struct VideoParse {
    unsigned nblocks, mv_bits, frames, gop;
    uint2 *pq, *partial;
    unsigned long long *pool;
    unsigned pool_cap;
    VFrameRec *rec;
    unsigned *result;              // [0] 1 = the records are valid, [1] pool entries used
    unsigned long long *state;     // chain state between launches: next bit, pool entries used, failed
};
struct VideoParseSizes { size_t nspec, pool_cap, bytes; unsigned scan_ctas; };
struct ParseParamsOpaque { unsigned long long raw[32]; };
VideoParseSizes video_parse_sizes(size_t enc_bytes, unsigned frames, int sm_count);
int launch_video_parse(const uint8_t *d_enc, const unsigned long long *d_enc_bits, const unsigned long long *d_start, int use_rle,
                       unsigned nblocks, unsigned mv_bits, unsigned frames, unsigned gop, const VideoParseSizes &z, uint8_t *scratch,
                       VideoParse &v, ParseParamsOpaque &popaque, cudaStream_t stream);
int launch_video_chain(const VideoParse &v, const ParseParamsOpaque &popaque, unsigned f0, unsigned f1, cudaStream_t stream);
int launch_video_emit(const VideoParse &v, const ParseParamsOpaque &popaque, unsigned first_frame, unsigned frame_step, unsigned nimg,
                      unsigned long long *block_off, cudaStream_t stream);

// ---- sharded decode of one image stream (parse.cu) ----
struct ShardedParseGeom { unsigned ctas_per_part, groups_per_part; size_t chunk_bytes, spec_bytes; };
ShardedParseGeom sharded_parse_geom(size_t enc_bytes, int N, unsigned parts);
int launch_parse_walk_part(const DecodeParams &d, size_t span_bits, uint8_t *scratch, uint8_t *d_spec, unsigned part, unsigned parts,
                           cudaStream_t stream);
int launch_parse_finish_range(const DecodeParams &d, size_t span_bits, uint8_t *scratch, uint8_t *d_spec, unsigned parts, unsigned lo, unsigned hi,
                              cudaStream_t stream);

int launch_parse_blocks(const DecodeParams &p, unsigned images, cudaStream_t stream);
int launch_decode_blocks(const DecodeParams &p, unsigned images, cudaStream_t stream);
size_t parse_scratch_bytes(size_t enc_bytes, int N);
int launch_parallel_parse(const DecodeParams &d, size_t span_bits, uint8_t *scratch, cudaStream_t stream);

}  // namespace ie
