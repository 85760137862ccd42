/*
 * imageencoder_b200 -- C-ABI of the B200-native block-transform hot path of ThenTech/ImageEncoder.
 *
 * The reference has no FFI; the path sits behind four C++ classes.  Each entry point below names
 * the reference interface it replaces; the thin C++ classes in imageencoder_b200/csrc/host/ keep the
 * reference's class names / ctor signatures and call these functions.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types.
 *   - return 0 on success, a negative IE_E* code on failure; ie_last_error() gives a thread-local message.
 *     Nothing here calls exit()/assert() (the reference does: ImageBase.cpp:24-27, ImageEncoder.cpp:26-28).
 *   - host entry points take HOST buffers and include the H2D/D2H copies; `_dev` entry points take DEVICE
 *     buffers + a cudaStream_t (passed as void*) and are what the HBM-resident benchmark times.
 *   - block size (4|8) and Huffman on/off are compile-time in the reference (Block.hpp:13, makefile:13)
 *     and run-time arguments here.
 *   - there is NO CPU fallback: every call fails with IE_ENODEVICE when no sm_100 device is usable.
 *   - threads: the HOST entry points (ie_encode_image, ie_decode_image, ie_encode_images, ie_decode_images,
 *     ie_encode_video, ie_decode_video) may be called concurrently from any number of threads; each call holds its own
 *     leased scratch session from entry to return (the reference is not re-entrant: file-static LUTs Block.cpp:30-31,
 *     static Frame::MVEC_BIT_SIZE Frame.cpp:6).  A `_dev` entry point works on the ie_session the caller passes: one
 *     session, one call at a time (use one session per thread / stream).
 *   - malformed streams: reads past the end of the stream give zero bits (BitStream.cpp:17-20) and decode like the
 *     reference; a block whose length field exceeds block*block -- where the reference indexes its zigzag table out of
 *     bounds, Block.cpp:460-465 -- fails with IE_EFORMAT.
 */
#ifndef IMAGEENCODER_B200_H
#define IMAGEENCODER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IE_OK            0
#define IE_EINVAL       -1   /* bad argument (dims not multiple of block size, > 32767, quant entry 0, ...) */
#define IE_ENODEVICE    -2   /* no CUDA device / wrong architecture / CUDA error at init */
#define IE_ECUDA        -3   /* CUDA runtime error during the call */
#define IE_ENOSPC       -4   /* caller-provided output buffer too small */
#define IE_EFORMAT      -5   /* malformed encoded stream */
#define IE_ENOMEM       -6

/* ---- lifetime ------------------------------------------------------------------------------------------- */
/* Select `device`, create the per-device state (LUTs: zigzag algo.cpp:68-87, cos/C tables algo.cpp:294-297,312,
 * MER pattern algo.cpp:90-139).  Idempotent per device; thread-safe. */
int ie_init(int device);
/* Frees the per-device tables and every idle cached session of the host entry points.  A later call re-initialises. */
void ie_shutdown(void);
const char *ie_last_error(void);
/* "imageencoder_b200 x.y (sm_100a)" */
const char *ie_version(void);

/* ---- sizing helpers --------------------------------------------------------------------------------------- */
/* Upper bound of the encoded size in bytes (header + blocks * (4 + 16 + 16*N*N) bits, Block.cpp:346-354
 * streamSize(), + Huffman dictionary slack). */
size_t ie_max_encoded_bytes(uint32_t width, uint32_t height, uint32_t block, uint32_t frames);

/* ---- images: replaces dc::ImageEncoder::process (ImageEncoder.cpp:52-175) and dc::ImageDecoder
 *      (ImageBase.cpp:98-129 + ImageDecoder.cpp:55-122) ------------------------------------------------------- */
/* raw: width*height bytes, row-major 8-bit grayscale.  quant: block*block u16, row-major (MatrixReader.cpp:65-134).
 * huffman != 0 produces what the reference's ENABLE_HUFFMAN build writes.  out receives the exact .enc bytes. */
int ie_encode_image(const uint8_t *raw, uint32_t width, uint32_t height, uint32_t block,
                    const uint16_t *quant, int use_rle, int huffman,
                    uint8_t *out, size_t out_cap, size_t *out_bytes);
/* enc: the .enc bytes.  The stream does not record the block size (compile-time in the reference) -> `block`.
 * raw_out receives width x height bytes; the width and height outputs are set even when IE_ENOSPC is returned. */
int ie_decode_image(const uint8_t *enc, size_t enc_bytes, uint32_t block,
                    uint8_t *raw_out, size_t raw_cap, uint32_t *width, uint32_t *height);
/* Batch of `count` equally sized images (BASELINE config 4).  raws: count*width*height contiguous bytes;
 * out: count slots of out_stride bytes; out_bytes[count]. */
int ie_encode_images(const uint8_t *raws, uint32_t count, uint32_t width, uint32_t height, uint32_t block,
                     const uint16_t *quant, int use_rle, int huffman,
                     uint8_t *out, size_t out_stride, size_t *out_bytes);
int ie_decode_images(const uint8_t *encs, size_t enc_stride, const size_t *enc_bytes, uint32_t count, uint32_t block,
                     uint8_t *raws_out, size_t raw_stride, uint32_t *width, uint32_t *height);

/* ---- video: replaces dc::VideoEncoder::process (VideoEncoder.cpp:22-111 + Frame.cpp:129-247) and
 *      dc::VideoDecoder (VideoBase.cpp:45-85 + VideoDecoder.cpp:33-62 + Frame.cpp:47-127) -------------------- */
/* yuv420: frames * (w*h*3/2) bytes, planar; only Y is coded.  Block size is 4, MacroBlock 16 (Block.hpp:13-14).
 * recon_out (optional, may be NULL): receives the encoder-side reconstruction the reference leaves in its raw buffer.
 * Both calls run as copy / compute pipelines: the clip goes up in batches of GOPs while earlier batches are encoded and the
 * finished part of the stream comes down; the decoder sends the frames of a reconstructed GOP batch down next to the decode of
 * the following ones.  Pass page-locked buffers (cudaHostAlloc / cudaHostRegister) for full PCIe speed in both directions; with
 * pageable buffers the calls are correct but the copies serialise (and the decoder's early copies are switched off). */
int ie_encode_video(const uint8_t *yuv420, size_t yuv_bytes, uint32_t width, uint32_t height,
                    const uint16_t *quant, int use_rle, uint32_t gop, uint32_t merange, int huffman,
                    uint8_t *out, size_t out_cap, size_t *out_bytes);
int ie_decode_video(const uint8_t *enc, size_t enc_bytes, int motioncompensation,
                    uint8_t *yuv_out, size_t yuv_cap, size_t *yuv_bytes,
                    uint32_t *width, uint32_t *height, uint32_t *frames);

/* ---- device-resident variants (what bench.py's HBM-resident `value` times) --------------------------------- */
typedef struct ie_session ie_session;   /* owns scratch (tile states, staging) sized for one problem shape */

/* kind: 0 image encode, 1 image decode, 2 video encode, 3 video decode.  frames = 1 for images (or the batch size). */
int ie_session_create(ie_session **s, int kind, uint32_t width, uint32_t height, uint32_t block, uint32_t frames);
void ie_session_destroy(ie_session *s);

/* d_raw/d_out are DEVICE pointers.  Asynchronous on `stream`; *d_out_bits (DEVICE u64, may be NULL) receives the
 * stream length in bits.  `first_bit` lets a shard write into a larger stream: the shard's first block lands at
 * that bit of d_out (multi-GPU stitching); header is written only when write_header != 0.
 * Huffman is a separate stage (ie_huffman_encode_dev). */
int ie_encode_image_dev(ie_session *s, const uint8_t *d_raw, uint32_t width, uint32_t height,
                        const uint16_t *quant, int use_rle, int lead_bit, int write_header, uint64_t first_bit,
                        uint8_t *d_out, size_t out_cap, uint64_t *d_out_bits, void *stream);
/* Batch of equally sized device-resident images (BASELINE config 4), one launch of each kernel for all of them: image i is
 * at d_raws + i * raw_stride, its stream goes to d_out + i * out_stride (out_stride >= ie_max_encoded_bytes of one image,
 * multiple of 16), d_out_bits[i] (device, optional) receives its bit count. */
int ie_encode_images_dev(ie_session *s, const uint8_t *d_raws, size_t raw_stride, uint32_t count, uint32_t width,
                         uint32_t height, const uint16_t *quant, int use_rle, int lead_bit, uint8_t *d_out,
                         size_t out_stride, uint64_t *d_out_bits, void *stream);
/* Split encode for a shard of a multi-GPU stream, so that the ONE collective of a sharded encode sits between
 * the two kernels and no re-alignment pass is needed:
 *   begin: transform/quantise/pack the shard's blocks into the session's tile scratch; *d_total_bits (device) = header bits
 *          (if write_header) + block bits of this shard                      -> all-gather of one u64 per rank
 *   end:   d_shard_totals (device, one u64 per shard, from the all-gather); this shard starts at bit
 *          first = sum(d_shard_totals[0 .. shard_index)) of the global stream.  d_out receives this shard's bytes of the
 *          global stream starting at byte (first / 128) * 16 (leading bits of that chunk zero); *d_out_bits (device,
 *          optional) = (first % 128) + this shard's bits; *d_first_bit (device, optional) = first.
 * Replaces ie_encode_image_dev + ie_stream_shift_dev. */
int ie_encode_image_begin_dev(ie_session *s, const uint8_t *d_raw, uint32_t width, uint32_t height,
                              const uint16_t *quant, int use_rle, int lead_bit, int write_header,
                              uint64_t *d_total_bits, void *stream);
int ie_encode_image_end_dev(ie_session *s, const uint64_t *d_shard_totals, uint32_t shard_index,
                            uint8_t *d_out, size_t out_cap, uint64_t *d_out_bits, uint64_t *d_first_bit, void *stream);
/* Height written into the header by the next ie_encode_image_dev calls on this session (0 = the height passed to
 * the call).  A block-row shard of a larger image writes the FULL image height (ImageEncoder.cpp:93-94). */
int ie_session_set_header_height(ie_session *s, uint32_t full_height);
/* Only the per-block bit lengths (first pass of a sharded encode): *d_total_bits = sum over the shard's blocks. */
int ie_image_bits_dev(ie_session *s, const uint8_t *d_raw, uint32_t width, uint32_t height,
                      const uint16_t *quant, int use_rle, uint64_t *d_total_bits, void *stream);
/* Decode of a plain (not Huffman-coded) device-resident stream whose header starts at bit `start_bit`.
 * d_enc must be 16-byte aligned (cudaMalloc'ed buffers are) and readable up to enc_bytes rounded up to 4. */
int ie_decode_image_dev(ie_session *s, const uint8_t *d_enc, size_t enc_bytes, uint64_t start_bit,
                        uint8_t *d_raw_out, size_t raw_cap, uint32_t *width, uint32_t *height, void *stream);
/* The same without the header read-back (and the stream synchronisation it costs): a caller that holds the first bytes of
 * the plain stream on the host -- the file reader does (ImageBase.cpp:98-129) -- parses the header once
 * (ie_parse_image_header: MatrixReader.cpp:45-57, ImageBase.cpp:122-128; start_bit = 1 behind the '0' "no Huffman" bit)
 * and every decode of that stream is then fully asynchronous on `stream`. */
typedef struct ie_image_header {
    uint32_t block, width, height, use_rle;
    uint64_t first_block_bit;          /* bit of the stream at which the first block starts */
    uint16_t quant[64];
} ie_image_header;
int ie_parse_image_header(const uint8_t *bytes, size_t nbytes, uint64_t start_bit, uint32_t block, ie_image_header *out);
int ie_decode_image_with_header_dev(ie_session *s, const ie_image_header *hdr, const uint8_t *d_enc, size_t enc_bytes,
                                    uint8_t *d_raw_out, size_t raw_cap, void *stream);
/* Sharded decode of ONE plain stream over several GPUs (replaces the serial block loop of
 * ImageDecoder.cpp:88-112 for an image too large for one GPU's share of the time).  Every rank holds the whole stream.
 *   begin: rank `part` of `parts` walks its share of the stream's speculative parse grid and leaves the per-group results in
 *          its chunk of d_spec (ie_decode_shard_spec_bytes bytes: [entry: parts x chunk_bytes][exit: parts x chunk_bytes];
 *          part r's chunks are at r * chunk_bytes of either half);
 *   the caller all-gathers both halves of d_spec in place (its own collective: NCCL all-gather, peer copies, ...);
 *   end:   verifies the seams of the gathered grid, emits the offsets of the blocks of block rows [block_row0, block_row1)
 *          and decodes them into d_rows_out (pitch = width: the band's pixels only).  If the speculation does not verify
 *          the exact parse runs on every rank (whole stream) -- same result, no scaling.
 * Both calls are asynchronous on `stream`; hdr from ie_parse_image_header. */
size_t ie_decode_shard_spec_bytes(size_t enc_bytes, uint32_t block, uint32_t parts, size_t *chunk_bytes);
int ie_decode_image_shard_begin_dev(ie_session *s, const ie_image_header *hdr, const uint8_t *d_enc, size_t enc_bytes,
                                    uint32_t part, uint32_t parts, uint8_t *d_spec, void *stream);
int ie_decode_image_shard_end_dev(ie_session *s, const ie_image_header *hdr, const uint8_t *d_enc, size_t enc_bytes,
                                  uint32_t parts, uint8_t *d_spec, uint32_t block_row0, uint32_t block_row1,
                                  uint8_t *d_rows_out, size_t out_cap, void *stream);
/* Byte-wise Huffman stage over a device-resident, byte-rounded plain stream (Huffman.cpp:232-344).
 * Synchronises `stream` once (the 256-entry tree is built on the host exactly as the reference does). */
/* Batch of device-resident plain streams (stream i at d_encs + i * enc_stride, enc_bytes[i] bytes; enc_stride a multiple
 * of 16), decoded concurrently on worker streams that fork from / join `stream`; image i goes to d_raws_out + i *
 * raw_stride; W[i], H[i] (host, optional) from the headers.  Errors found on the device (malformed stream) are reported by
 * the workers' error flags at the next synchronising call. */
int ie_decode_images_dev(ie_session *s, const uint8_t *d_encs, size_t enc_stride, const size_t *enc_bytes, uint32_t count,
                         uint64_t start_bit, uint8_t *d_raws_out, size_t raw_stride, uint32_t *W, uint32_t *H,
                         void *stream);
int ie_huffman_encode_dev(ie_session *s, const uint8_t *d_in, size_t in_bytes,
                          uint8_t *d_out, size_t out_cap, size_t *out_bytes, void *stream);
/* The same stage without any host synchronisation: the reference's dictionary build runs as a host callback in stream
 * order (cudaLaunchHostFunc) between the histogram kernels and the pack kernels, the revert rule (Huffman.cpp:329-341) is
 * decided on the device, the file's size in bytes is left in *d_out_bytes (device).  The call returns at once; errors (output
 * too small, a code longer than 32 bits) surface as the session's device-side error at the next synchronising call.  One call
 * in flight per session; out_cap >= 2080. */
int ie_huffman_encode_async_dev(ie_session *s, const uint8_t *d_in, size_t in_bytes, uint8_t *d_out, size_t out_cap,
                                uint64_t *d_out_bytes, void *stream);
int ie_huffman_decode_dev(ie_session *s, const uint8_t *d_in, size_t in_bytes,
                          uint8_t *d_out, size_t out_cap, size_t *out_bytes, uint64_t *start_bit, void *stream);
/* Huffman stage of one shard of a multi-GPU stream: the dictionary is built from the GLOBAL histogram and
 * first-occurrence positions (host arrays: the ranks' ie_byte_histogram_dev results, summed / min-reduced with the shard's
 * byte offset added), so every rank derives the same codes; `write_dictionary` (rank 0) puts the dictionary header in
 * front.  The shard's code bits start at bit 0 of d_out (after the dictionary on rank 0); *d_out_bits (device) = bits
 * written.  The "no gain" revert rule (Huffman.cpp:329-341) needs the global total and is the caller's decision. */
int ie_huffman_encode_shard_dev(ie_session *s, const uint8_t *d_in, size_t in_bytes, const uint32_t *hist,
                                const uint64_t *first_pos, int write_dictionary, uint8_t *d_out, size_t out_cap,
                                uint64_t *d_out_bits, void *stream);
/* Device histogram + first-occurrence positions of a byte stream (Huffman.cpp:236-243); hist[256] u32,
 * first_pos[256] u64 (UINT64_MAX = absent).  HOST outputs; synchronises `stream`. */
int ie_byte_histogram_dev(const uint8_t *d_in, size_t in_bytes, uint32_t *hist, uint64_t *first_pos, void *stream);
/* The same two steps with everything on the device and no host synchronisation: histogram and first occurrences into device
 * arrays (absent symbols: first = UINT64_MAX), so that the caller's collective (sum of the histograms, minimum of the offset
 * first occurrences) runs on them directly; the shard is then coded from the reduced device arrays -- the dictionary is built by
 * a stream-ordered host callback, as in ie_huffman_encode_async_dev -- and its bit count is left in *d_out_bits (device).
 * The revert rule (Huffman.cpp:329-341) is the caller's: it needs every shard's count.  out_cap >= 2080. */
int ie_byte_histogram_async_dev(ie_session *s, const uint8_t *d_in, size_t in_bytes, uint32_t *d_hist, uint64_t *d_first, void *stream);
int ie_huffman_encode_shard_async_dev(ie_session *s, const uint8_t *d_in, size_t in_bytes, const uint32_t *d_hist,
                                      const uint64_t *d_first, int write_dictionary, uint8_t *d_out, size_t out_cap,
                                      uint64_t *d_out_bits, void *stream);

int ie_encode_video_dev(ie_session *s, uint8_t *d_yuv420, size_t yuv_bytes, uint32_t width, uint32_t height,
                        const uint16_t *quant, int use_rle, uint32_t gop, uint32_t merange, int lead_bit,
                        uint8_t *d_out, size_t out_cap, uint64_t *d_out_bits, int16_t *d_mvecs, void *stream);
/* GOP shard of a longer clip (multi-GPU): the next ie_encode_video_dev calls on this session announce
 * `total_frames` in the header (0 = the frames of the call) and write the header only if write_header != 0 (the shard that
 * holds frame 0).  Shards start on GOP boundaries, so they are independent; their streams concatenate bit-contiguously
 * (all-gather of the bit totals + ie_stream_shift_dev, as for image shards). */
int ie_session_set_video_shard(ie_session *s, uint32_t total_frames, int write_header);
int ie_decode_video_dev(ie_session *s, const uint8_t *d_enc, size_t enc_bytes, uint64_t start_bit, int motioncompensation,
                        uint8_t *d_yuv_out, size_t yuv_cap, uint32_t *width, uint32_t *height, uint32_t *frames,
                        void *stream);

/* Multi-GPU stitch: re-align a shard stream (bit 0 of d_in, d_params[0] bits long) to the 128-bit chunk
 * grid of the global stream in which it starts at bit d_params[1]; d_params is a DEVICE array of two u64 so that the
 * offsets can come straight from an NCCL all-gather + scan without a host round trip. */
int ie_stream_shift_dev(const uint8_t *d_in, const uint64_t *d_params, uint8_t *d_out, size_t out_cap, void *stream);

/* ---- multi-GPU exchange: one ie_comm per rank -- one process per GPU, or several GPUs in one process.
 * The reference is single-process, single-device (ImageEncoder.cpp:121-138); what these replace is the hand-over between its
 * parallel block loop and its serial writer loop when the block rows of one image live on different GPUs.  A rank owns a
 * mailbox in its HBM that the peers write over NVLink (CUDA IPC between processes, peer access inside one); nothing here
 * needs NCCL, MPI or torch.  Set-up is three calls because the processes have to swap handles through whatever plumbing
 * the host has (torch.distributed.all_gather_object, MPI_Allgather, a file ...):
 *   ie_comm_create        on every rank (after ie_init on its device); stitch_bytes > 0 on rank 0 allocates the buffer the
 *                         single output stream is assembled in
 *   ie_comm_export        this rank's handle blob (ie_comm_handle_bytes() bytes)
 *   ie_comm_connect       `world` blobs in rank order (all-gathered by the caller)
 * All calls below are collective: every rank makes them, in the same order. */
typedef struct ie_comm ie_comm;
int ie_comm_create(ie_comm **c, int rank, int world, size_t stitch_bytes);
size_t ie_comm_handle_bytes(void);
int ie_comm_export(ie_comm *c, void *blob_out);
int ie_comm_connect(ie_comm *c, const void *blobs);
void ie_comm_destroy(ie_comm *c);
/* All-gather of one u64 per rank through the mailboxes (a one-CTA kernel: P2P stores, then a spin on this rank's mailbox):
 * d_totals_out[world] (device; NULL: kept inside the communicator, see ie_comm_totals_dev). */
int ie_comm_exchange_totals_dev(ie_comm *c, const uint64_t *d_total, uint64_t *d_totals_out, void *stream);
int ie_comm_totals_dev(ie_comm *c, const uint64_t **d_totals);
/* the totals of the last exchange -> dst[world] (host or device memory), asynchronous on `stream` */
int ie_comm_copy_totals(ie_comm *c, uint64_t *dst, void *stream);
/* Block-row shard of ONE image stream (BASELINE config 3): ie_encode_image_begin_dev -> mailbox exchange of the shard totals
 * -> ie_encode_image_end_dev.  This rank's rows [its share of height_total] are at d_raw; d_out receives its bytes of the
 * global stream from the chunk that holds its first bit; *d_out_bits = (first % 128) + shard bits, *d_first_bit = first
 * (device).  Rank 0 writes the header, announcing height_total.  Every rank needs at least one block row. */
int ie_encode_image_shard_dev(ie_session *s, ie_comm *c, const uint8_t *d_raw, uint32_t width, uint32_t height_shard,
                              uint32_t height_total, const uint16_t *quant, int use_rle, int lead_bit, uint8_t *d_out,
                              size_t out_cap, uint64_t *d_out_bits, uint64_t *d_first_bit, void *stream);
/* Device-side stitch: every rank stores its chunks into rank 0's stitch buffer over NVLink at their final position; the
 * chunk two neighbouring shards share is completed by the right neighbour (it receives the left one's bits through its
 * mailbox).  d_shard / d_bits / d_first_bit: what ie_encode_image_shard_dev (or ie_stream_shift_dev) left.  The stream's
 * length is the sum of the exchanged totals. */
int ie_comm_stitch_dev(ie_comm *c, const uint8_t *d_shard, const uint64_t *d_bits, const uint64_t *d_first_bit, void *stream);
/* rank 0: the buffer the stitch wrote (device pointer, valid until ie_comm_destroy) */
int ie_comm_stitched_stream(ie_comm *c, uint8_t **d_stream, size_t *capacity);
/* rank 0: the first nbytes of the stitched stream -> dst (host or device memory), asynchronous on `stream` */
int ie_comm_copy_stitched(ie_comm *c, void *dst, size_t nbytes, void *stream);

/* Diagnostics switches (process-wide; every setting produces identical streams, tests cross-check them):
 *   "exact_transform" = 1  the encoders evaluate every coefficient in the reference's exact binary64 order instead of the
 *                          guarded FP32 fast path;
 *   "encode_variant"  = 0 | 1 | 2 (default) | 3..7  instantiation of the tile kernel: 2 = packed f32x2 transform + quantise,
 *                          1 = packed quantise only, 0 = the scalar kernel they replaced (A/B timing, cross-checks);
 *                          experimental (arithmetic checked on the CPU, not yet timed): 3 / 4 = 2 with an 8 KiB staging area
 *                          for the tile image and 7 / 8 CTAs per SM, 5 = 2 with a short-chain binary64 pre-check in front
 *                          of the exact queue, 6 / 7 = 3 / 4 with that pre-check;
 *   "copyout_variant" = 0 | 1 | 2 | 3 (default)  copy-out kernel: 3 = a warp per tile image, word by word (no search, no
 *                          hand-off between groups), 2 = chunk-centric with a short path for interior chunks and four chunks
 *                          per thread in flight, 1 = short path one chunk at a time, 0 = the generic kernel;
 *   "decode_variant"  = 0 | 1 (default)  block-decode kernel of images and I-frames: 1 = inverse transform and pixel
 *                          stage in packed f32x2 operations, 0 = the scalar kernel it replaced;
 *   "video_decode_variant" = 0 | 1 (default)  video decode: 1 = one speculative parse over the whole stream, a short
 *                          sequential frame chain, then frame k of every GOP per launch (streams that end inside a frame or
 *                          hold an invalid length field fall back to 0); 0 = frame by frame;
 *   "video_encode_streams" = 1 | 2 (default)  video encode: the GOPs of a batch in two halves on two streams (one half's P-frame
 *                          tiles next to the other half's motion search) or all on the caller's stream;
 *   "video_decode_batches" = 1 .. 32 (default 4)  GOP batches of the whole-stream video decode: the sequential frame chain of
 *                          batch b + 1 runs next to the reconstruction of batch b (second stream);
 *   "me_variant"      = 0 | 1 | 2 (default)  motion-search kernel: 2 = eight lanes per MacroBlock, four MacroBlocks per warp
 *                          on one shared search window; 0 = a warp per MacroBlock (round 1); 1 = 0 with the SAD partial sums
 *                          reduced by warp-wide integer reductions (REDUX);
 *   "pframe_variant"  = 0 | 2 (default)  P-frame tiles: 2 = residual, transform, quantisation and reconstruction in packed
 *                          f32x2 operations, reference rows read as words; 0 = the scalar kernel it replaced;
 *   "encode_pad_smem" = bytes (default 0)  extra dynamic shared memory per CTA of the tile kernel: lowers its occupancy for
 *                          measurements (tools/ab_quick.py --pads), never useful otherwise.
 * Returns IE_EINVAL for an unknown name or an out-of-range value. */
int ie_set_option(const char *name, int value);

/* Number of kernels this library has launched since load (bench.py's `gpu_launches`). */
uint64_t ie_kernel_launch_count(void);

/* Counters since load, for tests that must know which path ran: "video_decode_whole_stream" (video decodes that took the
 * whole-stream parse), "video_decode_frame_by_frame" (decodes that took, or fell back to, the per-frame path).  Unknown
 * names give 0. */
uint64_t ie_stat(const char *name);

#ifdef __cplusplus
}
#endif
#endif /* IMAGEENCODER_B200_H */
