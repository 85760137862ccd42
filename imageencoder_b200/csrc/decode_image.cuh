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
    const BlockTables *tab;
    uint8_t *out;                           // device, decoded pixels
    size_t out_stride;
    size_t pitch;
    int *err;
};

int launch_parse_blocks(const DecodeParams &p, unsigned images, cudaStream_t stream);
int launch_decode_blocks(const DecodeParams &p, unsigned images, cudaStream_t stream);

}  // namespace ie
