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
};

int launch_parse_blocks(const DecodeParams &p, unsigned images, cudaStream_t stream);
int launch_decode_blocks(const DecodeParams &p, unsigned images, cudaStream_t stream);
size_t parse_scratch_bytes(size_t enc_bytes, int N);
int launch_parallel_parse(const DecodeParams &d, size_t span_bits, uint8_t *scratch, cudaStream_t stream);

}  // namespace ie
