"""Device-resident calls (torch tensors in HBM, torch's current stream).  torch is plumbing only: it owns the memory
and the stream; the work is done by the C-ABI `_dev` entry points."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import check, lib


def _dp(t: torch.Tensor):
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Session:
    """ie_session wrapper (scratch for one problem shape)."""
    IMAGE_ENCODE, IMAGE_DECODE, VIDEO_ENCODE, VIDEO_DECODE = 0, 1, 2, 3

    def __init__(self, kind: int, width: int, height: int, block: int, frames: int = 1):
        self.h = C.c_void_p()
        check(lib().ie_session_create(C.byref(self.h), kind, width, height, block, frames))
        self.width, self.height, self.block, self.frames = width, height, block, frames

    def close(self):
        if self.h:
            lib().ie_session_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _q(quant):
    q = np.ascontiguousarray(np.asarray(quant, dtype=np.uint16)).reshape(-1)
    return q, q.ctypes.data_as(C.POINTER(C.c_uint16))


def encode_image_dev(s: Session, d_raw: torch.Tensor, quant, rle: bool, d_out: torch.Tensor, d_bits: torch.Tensor | None = None,
                     lead_bit: bool = True, write_header: bool = True, first_bit: int = 0, width=None, height=None) -> None:
    """Asynchronous on torch's current stream.  d_raw u8 [H*W], d_out u8 (16-byte aligned), d_bits u64/int64 [1]."""
    q, qp = _q(quant)
    check(lib().ie_encode_image_dev(s.h, _dp(d_raw), width or s.width, height or s.height, qp, int(rle), int(lead_bit),
                                    int(write_header), first_bit, _dp(d_out), d_out.numel(),
                                    _dp(d_bits) if d_bits is not None else None, _stream()))


def encode_images_dev(s: Session, d_raws: torch.Tensor, count: int, quant, rle: bool, d_out: torch.Tensor, out_stride: int,
                      d_bits: torch.Tensor | None = None, lead_bit: bool = True, width=None, height=None) -> None:
    """Batch of `count` equally sized images (contiguous in d_raws) in one launch of each kernel; stream i at d_out[i*out_stride:]."""
    q, qp = _q(quant)
    w, h = width or s.width, height or s.height
    check(lib().ie_encode_images_dev(s.h, _dp(d_raws), w * h, count, w, h, qp, int(rle), int(lead_bit), _dp(d_out), out_stride,
                                     _dp(d_bits) if d_bits is not None else None, _stream()))


def encode_image_begin_dev(s: Session, d_raw: torch.Tensor, quant, rle: bool, d_total_bits: torch.Tensor, lead_bit: bool = True,
                           write_header: bool = True, width=None, height=None) -> None:
    """First half of a sharded encode: blocks -> tile scratch; d_total_bits[0] = header bits + block bits of this shard."""
    q, qp = _q(quant)
    check(lib().ie_encode_image_begin_dev(s.h, _dp(d_raw), width or s.width, height or s.height, qp, int(rle), int(lead_bit),
                                          int(write_header), _dp(d_total_bits), _stream()))


def encode_image_end_dev(s: Session, d_shard_totals: torch.Tensor, shard_index: int, d_out: torch.Tensor,
                         d_out_bits: torch.Tensor | None = None, d_first_bit: torch.Tensor | None = None) -> None:
    """Second half: the shard's bytes of the global stream, from byte (first // 128) * 16, into d_out."""
    check(lib().ie_encode_image_end_dev(s.h, _dp(d_shard_totals), shard_index, _dp(d_out), d_out.numel(),
                                        _dp(d_out_bits) if d_out_bits is not None else None,
                                        _dp(d_first_bit) if d_first_bit is not None else None, _stream()))


def image_bits_dev(s: Session, d_raw: torch.Tensor, quant, rle: bool, d_bits: torch.Tensor, width=None, height=None) -> None:
    q, qp = _q(quant)
    check(lib().ie_image_bits_dev(s.h, _dp(d_raw), width or s.width, height or s.height, qp, int(rle), _dp(d_bits), _stream()))


def decode_image_dev(s: Session, d_enc: torch.Tensor, enc_bytes: int, d_raw_out: torch.Tensor, start_bit: int = 1):
    w, h = C.c_uint32(0), C.c_uint32(0)
    check(lib().ie_decode_image_dev(s.h, _dp(d_enc), enc_bytes, start_bit, _dp(d_raw_out), d_raw_out.numel(), C.byref(w),
                                    C.byref(h), _stream()))
    return w.value, h.value


class ImageHeader(C.Structure):
    """ie_image_header (include/imageencoder_b200.h)"""
    _fields_ = [("block", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32), ("use_rle", C.c_uint32),
                ("first_block_bit", C.c_uint64), ("quant", C.c_uint16 * 64)]


def parse_image_header(first_bytes: bytes, block: int, start_bit: int = 1) -> ImageHeader:
    h = ImageHeader()
    buf = (C.c_uint8 * len(first_bytes)).from_buffer_copy(first_bytes)
    check(lib().ie_parse_image_header(buf, len(first_bytes), start_bit, block, C.byref(h)))
    return h


def decode_image_with_header_dev(s: Session, hdr: ImageHeader, d_enc: torch.Tensor, enc_bytes: int, d_raw_out: torch.Tensor) -> None:
    """fully asynchronous on torch's current stream (no header read-back)"""
    check(lib().ie_decode_image_with_header_dev(s.h, C.byref(hdr), _dp(d_enc), enc_bytes, _dp(d_raw_out), d_raw_out.numel(), _stream()))


def decode_shard_spec_bytes(enc_bytes: int, block: int, parts: int) -> tuple[int, int]:
    """(bytes of the spec buffer of a sharded decode, bytes of one part's chunk in either half of it)"""
    chunk = C.c_size_t(0)
    total = int(lib().ie_decode_shard_spec_bytes(enc_bytes, block, parts, C.byref(chunk)))
    return total, int(chunk.value)


def decode_image_shard_begin_dev(s: Session, hdr: ImageHeader, d_enc: torch.Tensor, enc_bytes: int, part: int, parts: int,
                                 d_spec: torch.Tensor) -> None:
    check(lib().ie_decode_image_shard_begin_dev(s.h, C.byref(hdr), _dp(d_enc), enc_bytes, part, parts, _dp(d_spec), _stream()))


def decode_image_shard_end_dev(s: Session, hdr: ImageHeader, d_enc: torch.Tensor, enc_bytes: int, parts: int, d_spec: torch.Tensor,
                               block_row0: int, block_row1: int, d_rows_out: torch.Tensor) -> None:
    check(lib().ie_decode_image_shard_end_dev(s.h, C.byref(hdr), _dp(d_enc), enc_bytes, parts, _dp(d_spec), block_row0, block_row1,
                                              _dp(d_rows_out), d_rows_out.numel(), _stream()))


def decode_images_dev(s: Session, d_encs: torch.Tensor, enc_stride: int, enc_bytes, d_raws_out: torch.Tensor, raw_stride: int,
                      start_bit: int = 1):
    """Batch decode of device-resident plain streams on the session's worker streams; returns the (W, H) lists."""
    n = len(enc_bytes)
    eb = (C.c_size_t * n)(*[int(b) for b in enc_bytes])
    w, h = (C.c_uint32 * n)(), (C.c_uint32 * n)()
    check(lib().ie_decode_images_dev(s.h, _dp(d_encs), enc_stride, eb, n, start_bit, _dp(d_raws_out), raw_stride, w, h, _stream()))
    return list(w), list(h)


def encode_video_dev(s: Session, d_yuv: torch.Tensor, width: int, height: int, quant, rle: bool, gop: int, merange: int,
                     d_out: torch.Tensor, d_bits: torch.Tensor | None = None, lead_bit: bool = True, d_mvecs: torch.Tensor | None = None) -> None:
    """``dc::VideoEncoder::process`` on device-resident YUV420 frames (the Y planes are rebuilt in place, Frame.cpp:218-242).
    Asynchronous on torch's current stream."""
    q, qp = _q(quant)
    check(lib().ie_encode_video_dev(s.h, _dp(d_yuv), d_yuv.numel(), width, height, qp, int(rle), gop, merange, int(lead_bit),
                                    _dp(d_out), d_out.numel(), _dp(d_bits) if d_bits is not None else None,
                                    _dp(d_mvecs) if d_mvecs is not None else None, _stream()))


def decode_video_dev(s: Session, d_enc: torch.Tensor, enc_bytes: int, d_yuv_out: torch.Tensor, motioncompensation: bool = True,
                     start_bit: int = 1):
    """``dc::VideoDecoder`` on a device-resident plain stream; returns (width, height, frames)."""
    w, h, f = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
    check(lib().ie_decode_video_dev(s.h, _dp(d_enc), enc_bytes, start_bit, int(motioncompensation), _dp(d_yuv_out),
                                    d_yuv_out.numel(), C.byref(w), C.byref(h), C.byref(f), _stream()))
    return w.value, h.value, f.value


def huffman_encode_dev(s: Session, d_in: torch.Tensor, in_bytes: int, d_out: torch.Tensor) -> int:
    n = C.c_size_t(0)
    check(lib().ie_huffman_encode_dev(s.h, _dp(d_in), in_bytes, _dp(d_out), d_out.numel(), C.byref(n), _stream()))
    return n.value


def huffman_encode_async_dev(s: Session, d_in: torch.Tensor, in_bytes: int, d_out: torch.Tensor, d_out_bytes: torch.Tensor) -> None:
    """the stage without a host synchronisation (host callback in stream order); d_out_bytes: int64 device tensor [1]"""
    check(lib().ie_huffman_encode_async_dev(s.h, _dp(d_in), in_bytes, _dp(d_out), d_out.numel(), _dp(d_out_bytes), _stream()))


def byte_histogram_async_dev(s: Session, d_in: torch.Tensor, n: int, d_hist: torch.Tensor, d_first: torch.Tensor) -> None:
    """histogram (int32 tensor [256]) and first occurrences (int64 tensor [256], -1 = absent) on the device, no synchronisation"""
    check(lib().ie_byte_histogram_async_dev(s.h, _dp(d_in), n, _dp(d_hist), _dp(d_first), _stream()))


def huffman_encode_shard_async_dev(s: Session, d_in: torch.Tensor, n: int, d_hist: torch.Tensor, d_first: torch.Tensor, write_dictionary: bool,
                                   d_out: torch.Tensor, d_out_bits: torch.Tensor) -> None:
    check(lib().ie_huffman_encode_shard_async_dev(s.h, _dp(d_in), n, _dp(d_hist), _dp(d_first), int(bool(write_dictionary)), _dp(d_out),
                                                  d_out.numel(), _dp(d_out_bits), _stream()))


def huffman_encode_shard_dev(s: Session, d_in: torch.Tensor, in_bytes: int, hist, first_pos, write_dictionary: bool,
                             d_out: torch.Tensor, d_out_bits: torch.Tensor) -> None:
    """Huffman stage of one shard with the GLOBAL histogram / first-occurrence positions (host arrays)."""
    h = np.ascontiguousarray(hist, dtype=np.uint32)
    f = np.ascontiguousarray(first_pos, dtype=np.uint64)
    check(lib().ie_huffman_encode_shard_dev(s.h, _dp(d_in), in_bytes, h.ctypes.data_as(C.POINTER(C.c_uint32)),
                                            f.ctypes.data_as(C.POINTER(C.c_uint64)), int(write_dictionary), _dp(d_out),
                                            d_out.numel(), _dp(d_out_bits), _stream()))


def stream_shift_dev(d_in: torch.Tensor, d_params: torch.Tensor, d_out: torch.Tensor) -> None:
    """d_out bit (params[1] % 128 + i) = d_in bit i for i < params[0]  (re-alignment of a shard stream)."""
    check(lib().ie_stream_shift_dev(_dp(d_in), _dp(d_params), _dp(d_out), d_out.numel(), _stream()))


def byte_histogram_dev(d_in: torch.Tensor, n: int):
    hist = np.zeros(256, dtype=np.uint32)
    first = np.zeros(256, dtype=np.uint64)
    check(lib().ie_byte_histogram_dev(_dp(d_in), n, hist.ctypes.data_as(C.POINTER(C.c_uint32)),
                                      first.ctypes.data_as(C.POINTER(C.c_uint64)), _stream()))
    return hist, first
