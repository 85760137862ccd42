// C-ABI: video entry points (placeholder until video.cu lands in this round).
#include "api_internal.cuh"
using namespace ie;
extern "C" {
int ie_encode_video(const uint8_t *, size_t, uint32_t, uint32_t, const uint16_t *, int, uint32_t, uint32_t, int, uint8_t *, size_t, size_t *) {
    set_error("video path not built yet"); return IE_EINVAL;
}
int ie_decode_video(const uint8_t *, size_t, int, uint8_t *, size_t, size_t *, uint32_t *, uint32_t *, uint32_t *) {
    set_error("video path not built yet"); return IE_EINVAL;
}
int ie_encode_video_dev(ie_session *, uint8_t *, size_t, uint32_t, uint32_t, const uint16_t *, int, uint32_t, uint32_t, int, uint8_t *, size_t, uint64_t *, int16_t *, void *) {
    set_error("video path not built yet"); return IE_EINVAL;
}
int ie_decode_video_dev(ie_session *, const uint8_t *, size_t, uint64_t, int, uint8_t *, size_t, uint32_t *, uint32_t *, uint32_t *, void *) {
    set_error("video path not built yet"); return IE_EINVAL;
}
}
