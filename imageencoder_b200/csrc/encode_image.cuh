#pragma once
#include "common.cuh"
#include "pack.cuh"
#include "transform.cuh"

namespace ie {

struct EncodeParams {
    const uint8_t *src;               // device, image(s) row-major u8
    size_t pitch;                     // bytes per pixel row (== W)
    size_t img_stride;                // bytes between images of a batch
    unsigned bx;                      // blocks per block-row
    unsigned nblocks;                 // blocks per image
    unsigned tiles_per_image;
    int use_rle;
    int bits_only;                    // 1: only the bit totals (first pass of a sharded encode)
    QuantParam quant;
    const BlockTables *tab;
    uint8_t *out;                     // device stream buffer(s), 16-byte aligned
    size_t out_stride;                // bytes between the streams of a batch
    size_t out_cap;                   // bytes per stream
    unsigned long long *bit_counter;  // [images] in: first free bit of the stream, out: one past the last written bit
    int *err;                         // device error flag
    ScanState scan;
};

unsigned encode_tile_blocks(int N);
int launch_encode_tiles(int N, const EncodeParams &p, unsigned images, cudaStream_t stream);

}  // namespace ie
