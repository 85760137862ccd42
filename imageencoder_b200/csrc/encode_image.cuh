#pragma once
#include "common.cuh"
#include "pack.cuh"
#include "transform.cuh"
#include "transform_fast.cuh"
#include <atomic>

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
    FastQuant fq;                     // fast-path constants + guard-band thresholds (transform_fast.cuh)
    int dc_den2;                      // 2 * 4 * Q[0][0]: exact integer DC rounding
    float dc_rcp;                     // 1 / dc_den2
    const BlockTables *tab;
    uint8_t *out;                     // device stream buffer(s), 16-byte aligned
    size_t out_stride;                // bytes between the streams of a batch
    size_t out_cap;                   // bytes per stream
    unsigned long long *bit_counter;  // [images] in: first free bit of the stream, out: one past the last written bit
    int *err;                         // device error flag
    ScanState scan;                   // only the tile-boundary hand-off records are used by the image path
    uint8_t *tile_scratch;            // [images * tiles] slots of slot_bytes: packed tile images (tile-local alignment)
    size_t slot_bytes;
    int phase;                        // 0: tile kernel + copy-out; 1: tile kernel only (tile images stay in scratch); 2: copy-out only
    // stream prefix written by the tile kernel itself (tile 0 of every image) instead of a separate launch: `prefix_first`
    // zero bits, then the header; the stream's blocks start behind it
    int write_prefix;
    unsigned prefix_first;
    HeaderParam hdr;
    unsigned long long *out_bits;     // optional [images]: receives the streams' final bit counts (copy-out kernel)
    unsigned *tile_bits;              // [images * tiles] bits per tile
    unsigned long long *bit_base;     // [images] bit position each stream had when this launch started
    // P-frame mode (Frame.cpp:160-244): src is the CURRENT frame (read, then overwritten with the reconstruction),
    // ref the previous frame as the encoder left it; per-MacroBlock pixel coordinates of the block the residual is
    // taken from (res_coord, Block.cpp:337) and of the block copied in (copy_coord, Frame.cpp:218-225).
    const uint8_t *ref;
    const short *res_coord;           // [nMB][2] (x, y)
    const short *copy_coord;          // [nMB][2]
    uint8_t *cur_rw;                  // == src, writable
    unsigned mbx;                     // MacroBlocks per row
    size_t coord_stride;              // shorts between the coordinate arrays of consecutive images (GOP batch)
    float k2[16];                     // P-frame reconstruction, fast path: C(u)C(v) * Q[u][v] (4x4)
};

unsigned encode_tile_blocks(int N);
size_t encode_tile_slot_bytes(int N);      // bytes of scratch per tile
extern std::atomic<int> g_exact_transform;
extern std::atomic<int> g_encode_variant;
extern std::atomic<int> g_copyout_variant;
extern std::atomic<int> g_fused_debug;
extern std::atomic<int> g_encode_pad_smem;
// max_abs_sample: 128 for pixels - 128, 383 for P-frame residuals - 128
void make_fast_quant(FastQuant &fq, const uint16_t *quant, int N, double max_abs_sample);
int launch_encode_tiles(int N, const EncodeParams &p, unsigned images, cudaStream_t stream);
// fused stream kernel (encode_fused.cu): sizes of the scan arrays it needs, and the launch (one image, whole stream)
void encode_fused_sizes(int N, unsigned nblocks, unsigned &n_wtiles, unsigned &n_ctatiles);
int launch_encode_fused(int N, const EncodeParams &p, int append, int sm_count, cudaStream_t stream);
int launch_pframe_tiles(const EncodeParams &p, unsigned images, cudaStream_t stream);
int launch_tile_copyout(const EncodeParams &p, unsigned images, cudaStream_t stream);          // phase 2 of a split encode
int launch_tile_totals(const unsigned *tile_bits, unsigned ntiles, unsigned images, unsigned long long add, unsigned long long *d_total,
                       cudaStream_t stream);                                                   // d_total[img] = add + sum of the image's tile bits

}  // namespace ie
