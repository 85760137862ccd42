// Declarations shared by the api_*.cu translation units.
#pragma once
#include "common.cuh"
#include "session.cuh"
#include "encode_image.cuh"
#include "decode_image.cuh"

namespace ie {

struct ParsedHeader {
    uint16_t quant[kMaxNN];
    int use_rle;
    uint32_t W, H, frames, gop, merange;
    size_t end_bit;              // first bit after the header
};

int check_quant(const uint16_t *quant, int N);
int check_dims(uint32_t W, uint32_t H, uint32_t N);
int make_quant(QuantParam &q, const uint16_t *quant, int N);
void make_k2(float *k2, const uint16_t *quant, int N);
int parse_header(const uint8_t *bytes, size_t n, size_t start_bit, int N, ParsedHeader &h, int video);
int launch_stream_init(uint8_t *out, size_t out_stride, unsigned images, const HeaderParam &hdr, unsigned first_bit,
                       unsigned long long *counter, cudaStream_t stream);
int encode_images_dev(ie_session *s, const uint8_t *d_raw, size_t img_stride, unsigned images, uint32_t W, uint32_t H, int N,
                      const uint16_t *quant, int use_rle, int lead_bit, int write_header, unsigned first_bit, int bits_only,
                      uint8_t *d_out, size_t out_stride, size_t out_cap, cudaStream_t stream, int append = 0, uint32_t header_H = 0, int split = 0,
                      uint64_t *d_out_bits = nullptr);
int session_ensure_pipeline(ie_session *s);
// host_out != NULL: the decoded pixels are also copied to that (host) buffer, stripe by stripe on the session's stream_out
int decode_image_dev(ie_session *s, const uint8_t *d_enc, size_t enc_bytes, size_t start_bit, int N, const ParsedHeader &h,
                     uint8_t *d_out, size_t out_cap, cudaStream_t stream, uint8_t *host_out = nullptr);
int read_err_flag(ie_session *s, cudaStream_t stream);
// Exclusive use of a cached session for the duration of one host-buffer call (api_image.cu).
class SessionLease {
  public:
    SessionLease() = default;
    SessionLease(const SessionLease &) = delete;
    SessionLease &operator=(const SessionLease &) = delete;
    ~SessionLease() { release(); }
    int acquire(int kind, uint32_t W, uint32_t H, uint32_t N, uint32_t frames);
    void release();
    ie_session *get() const { return s_; }

  private:
    ie_session *s_ = nullptr;
};
void drop_cached_sessions();

}  // namespace ie
