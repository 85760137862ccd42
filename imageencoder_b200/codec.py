"""Python mirror of the reference's entry points, above the C-ABI.

``ImageEncoder / ImageDecoder / VideoEncoder / VideoDecoder`` keep the reference's constructor arguments and its
``process()`` / ``saveResult()`` protocol (ImageEncoder.hpp:16-22, ImageDecoder.hpp:15-19, VideoEncoder.hpp:13-15,
VideoDecoder.hpp:15-17) so the parity tests read like the reference's own usage in main.cpp:105-150.  All compute goes
through ``libimageencoder_b200.so``; nothing here computes a pixel or a bit on the CPU.
"""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

import numpy as np

from . import _lib
from ._lib import check, lib


def _lexical_cast_u16(item: str) -> int:
    """``util::lexical_cast<uint16_t>`` (utils.hpp:293-305): formatted stream extraction -- hexadecimal if the text starts with
    0x / 0X, leading whitespace and a sign accepted (a negative value wraps), whatever follows the number ignored, failure if
    there is no number or its magnitude does not fit 16 bits."""
    hexmode = item[:2].upper() == "0X"
    m = re.match(r"\s*([+-]?)(?:0[xX])?([0-9a-fA-F]+)" if hexmode else r"\s*([+-]?)([0-9]+)", item)
    if not m:
        raise ValueError(f"[MatrixReader] cannot cast '{item}' to uint16")
    v = int(m.group(2), 16 if hexmode else 10)
    if v > 65535:
        raise ValueError(f"[MatrixReader] '{item}' does not fit uint16")
    return (-v) & 0xFFFF if m.group(1) == "-" else v


def read_matrix(path) -> np.ndarray:
    """``MatrixReader<>::read`` (MatrixReader.cpp:65-134), the same way ``csrc/host/codec.cpp`` mirrors it: lines split at newlines,
    trimmed, runs of spaces collapsed, items separated by single spaces only (a tab is not a separator), each item through
    ``lexical_cast<uint16_t>``; a blank line is a row without columns (an error).  The block size is the number of rows (4 or 8)."""
    text = Path(path).read_bytes().decode("latin-1")
    lines = text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()                                   # std::getline yields no empty line after a final newline
    rows = []
    for line in lines:
        line = re.sub(" +", " ", line.strip(" \t\n\v\f\r"))
        rows.append([_lexical_cast_u16(item) for item in line.split(" ")] if line else [])
    n = len(rows)
    if n not in (4, 8) or any(len(r) != n for r in rows):
        raise ValueError(f"[MatrixReader] expected a 4x4 or 8x8 matrix in {path}")
    return np.array(rows, dtype=np.uint16)


def _quant(quant, block=None) -> np.ndarray:
    q = np.ascontiguousarray(np.asarray(quant, dtype=np.uint16)).reshape(-1)
    if block is None:
        block = {16: 4, 64: 8}.get(q.size)
    if block not in (4, 8) or q.size != block * block:
        raise ValueError("quant matrix must be 4x4 or 8x8 and match the block size")
    return q


def _ptr(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


def _u16(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_uint16))


# ------------------------------------------------------------------------------------------------------------------
# functional forms (host buffers in, host buffers out; H2D/D2H inside)
# ------------------------------------------------------------------------------------------------------------------
def encode_image(raw, width: int, height: int, quant, rle: bool = True, huffman: bool = False, block: int | None = None,
                 out: np.ndarray | None = None) -> bytes | int:
    """``dc::ImageEncoder::process`` + ``saveResult``: returns the exact ``.enc`` bytes (or the size if `out` is given)."""
    q = _quant(quant, block)
    block = 4 if q.size == 16 else 8
    a = np.ascontiguousarray(raw, dtype=np.uint8).reshape(-1)
    if a.size != width * height:
        raise ValueError("raw size must be width*height (ImageEncoder.cpp:28)")
    own = out is None
    if own:
        out = np.empty(lib().ie_max_encoded_bytes(width, height, block, 1), dtype=np.uint8)
    n = C.c_size_t(0)
    check(lib().ie_encode_image(_ptr(a), width, height, block, _u16(q), int(bool(rle)), int(bool(huffman)), _ptr(out),
                                out.size, C.byref(n)))
    return out[: n.value].tobytes() if own else n.value


def decode_image(enc, block: int = 4, out: np.ndarray | None = None) -> np.ndarray:
    """``dc::ImageDecoder``: returns the decoded image as an (H, W) uint8 array."""
    e = np.frombuffer(enc, dtype=np.uint8) if not isinstance(enc, np.ndarray) else np.ascontiguousarray(enc, dtype=np.uint8)
    w = C.c_uint32(0)
    h = C.c_uint32(0)
    buf = out if out is not None else np.empty(1 << 16, dtype=np.uint8)
    rc = lib().ie_decode_image(_ptr(e), e.size, block, _ptr(buf), buf.size, C.byref(w), C.byref(h))
    if rc == _lib.IE_ENOSPC and out is None and w.value and h.value:
        buf = np.empty(w.value * h.value, dtype=np.uint8)
        rc = lib().ie_decode_image(_ptr(e), e.size, block, _ptr(buf), buf.size, C.byref(w), C.byref(h))
    check(rc)
    return buf[: w.value * h.value].reshape(h.value, w.value)


def encode_images(raws, width: int, height: int, quant, rle: bool = True, huffman: bool = False, block: int | None = None) -> list:
    """Batch of equally sized images (BASELINE config 4) through ``ie_encode_images``: every image is what
    ``dc::ImageEncoder`` writes for it.  `raws`: (count, height, width) uint8.  Returns a list of ``.enc`` byte strings."""
    q = _quant(quant, block)
    block = 4 if q.size == 16 else 8
    a = np.ascontiguousarray(raws, dtype=np.uint8)
    npx = width * height
    if a.size % npx:
        raise ValueError("raws must hold a whole number of width*height images")
    count = a.size // npx
    a = a.reshape(-1)
    slot = int(lib().ie_max_encoded_bytes(width, height, block, 1))
    out = np.empty(count * slot, dtype=np.uint8)
    sizes = (C.c_size_t * count)()
    check(lib().ie_encode_images(_ptr(a), count, width, height, block, _u16(q), int(bool(rle)), int(bool(huffman)), _ptr(out),
                                 slot, sizes))
    return [out[i * slot: i * slot + sizes[i]].tobytes() for i in range(count)]


def decode_images(encs, block: int = 4) -> list:
    """Batch decode through ``ie_decode_images``: `encs` is a list of ``.enc`` byte strings of equally sized images."""
    count = len(encs)
    stride = (max(len(e) for e in encs) + 16 + 15) // 16 * 16
    buf = np.zeros(count * stride, dtype=np.uint8)
    sizes = (C.c_size_t * count)()
    for i, e in enumerate(encs):
        buf[i * stride: i * stride + len(e)] = np.frombuffer(e, dtype=np.uint8)
        sizes[i] = len(e)
    # the size of the images is in the first stream's header: ask for it with a zero-capacity call of the single-image entry
    w, h = C.c_uint32(0), C.c_uint32(0)
    probe = np.empty(1, dtype=np.uint8)
    e0 = np.frombuffer(encs[0], dtype=np.uint8)
    rc = lib().ie_decode_image(_ptr(e0), e0.size, block, _ptr(probe), 0, C.byref(w), C.byref(h))
    if rc not in (_lib.IE_OK, _lib.IE_ENOSPC):
        check(rc)
    npx = w.value * h.value
    raws = np.empty(count * npx, dtype=np.uint8)
    check(lib().ie_decode_images(_ptr(buf), stride, sizes, count, block, _ptr(raws), npx, C.byref(w), C.byref(h)))
    return [raws[i * npx: (i + 1) * npx].reshape(h.value, w.value) for i in range(count)]


def encode_video(yuv, width: int, height: int, quant, rle: bool = True, gop: int = 4, merange: int = 16,
                 huffman: bool = False) -> bytes:
    """``dc::VideoEncoder::process`` + ``saveResult`` on a YUV420 planar buffer (only Y is coded)."""
    q = _quant(quant, 4)
    a = np.ascontiguousarray(yuv, dtype=np.uint8).reshape(-1)
    frames = a.size // (width * height * 3 // 2)
    out = np.empty(lib().ie_max_encoded_bytes(width, height, 4, max(1, frames)), dtype=np.uint8)
    n = C.c_size_t(0)
    check(lib().ie_encode_video(_ptr(a), a.size, width, height, _u16(q), int(bool(rle)), gop, merange, int(bool(huffman)),
                                _ptr(out), out.size, C.byref(n)))
    return out[: n.value].tobytes()


def decode_video(enc, motioncompensation: bool = True):
    """``dc::VideoDecoder``: returns (yuv420 bytes as uint8 array, width, height, frames)."""
    e = np.frombuffer(enc, dtype=np.uint8) if not isinstance(enc, np.ndarray) else np.ascontiguousarray(enc, dtype=np.uint8)
    w, h, f = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
    n = C.c_size_t(0)
    buf = np.empty(1 << 16, dtype=np.uint8)
    rc = lib().ie_decode_video(_ptr(e), e.size, int(bool(motioncompensation)), _ptr(buf), buf.size, C.byref(n), C.byref(w),
                               C.byref(h), C.byref(f))
    if rc == _lib.IE_ENOSPC and n.value:
        buf = np.empty(n.value, dtype=np.uint8)
        rc = lib().ie_decode_video(_ptr(e), e.size, int(bool(motioncompensation)), _ptr(buf), buf.size, C.byref(n),
                                   C.byref(w), C.byref(h), C.byref(f))
    check(rc)
    return buf[: n.value], w.value, h.value, f.value


# ------------------------------------------------------------------------------------------------------------------
# the reference's classes
# ------------------------------------------------------------------------------------------------------------------
class ImageEncoder:
    """``dc::ImageEncoder(rawfile, encfile, width, height, rle, quant_m)`` (ImageEncoder.hpp:16-18).

    `huffman` selects what the reference decides at compile time with -DENABLE_HUFFMAN (makefile:13)."""

    def __init__(self, source_file, dest_file, width: int, height: int, use_rle: bool, quant_m, huffman: bool = False):
        self.dest_file = str(dest_file)
        self.width, self.height, self.use_rle, self.huffman = int(width), int(height), bool(use_rle), bool(huffman)
        self.quant = np.asarray(quant_m, dtype=np.uint16)
        self.raw = np.fromfile(source_file, dtype=np.uint8)        # ImageBase.cpp:19-30
        block = 4 if self.quant.size == 16 else 8
        if self.width % block or self.height % block or self.raw.size != self.width * self.height:
            raise ValueError("width/height must be multiples of the block size and match the file size "
                             "(ImageEncoder.cpp:26-28)")
        self.result: bytes | None = None

    def process(self) -> bool:
        self.result = encode_image(self.raw, self.width, self.height, self.quant, self.use_rle, self.huffman)
        return True

    def saveResult(self) -> None:                                   # ImageBase.cpp:315-336
        Path(self.dest_file).write_bytes(self.result)


class ImageDecoder:
    """``dc::ImageDecoder(encfile, decfile)`` (ImageDecoder.hpp:15).  `block` is compile-time in the reference."""

    def __init__(self, source_file, dest_file, block: int = 4):
        self.dest_file = str(dest_file)
        self.enc = np.fromfile(source_file, dtype=np.uint8)
        self.block = block
        self.result: np.ndarray | None = None

    def process(self) -> bool:
        self.result = decode_image(self.enc, self.block)
        self.height, self.width = self.result.shape
        return True

    def saveResult(self) -> None:
        self.result.tofile(self.dest_file)


class VideoEncoder:
    """``dc::VideoEncoder(rawfile, encfile, width, height, rle, quant_m, gop, merange)`` (VideoEncoder.hpp:13-15)."""

    def __init__(self, source_file, dest_file, width, height, use_rle, quant_m, gop, merange, huffman: bool = False):
        self.dest_file = str(dest_file)
        self.width, self.height, self.use_rle = int(width), int(height), bool(use_rle)
        self.gop, self.merange, self.huffman = max(1, int(gop)), int(merange), bool(huffman)   # VideoBase.cpp:34
        self.quant = np.asarray(quant_m, dtype=np.uint16)
        self.raw = np.fromfile(source_file, dtype=np.uint8)
        if self.raw.size % (self.width * self.height * 3 // 2):
            raise ValueError("file size must be a multiple of the YUV420 frame size (VideoEncoder.cpp:16)")
        self.result: bytes | None = None

    def process(self) -> bool:
        self.result = encode_video(self.raw, self.width, self.height, self.quant, self.use_rle, self.gop, self.merange,
                                   self.huffman)
        return True

    def saveResult(self) -> None:
        Path(self.dest_file).write_bytes(self.result)


class VideoDecoder:
    """``dc::VideoDecoder(encfile, decfile, motioncomp)`` (VideoDecoder.hpp:15-17)."""

    def __init__(self, source_file, dest_file, motioncomp: bool = True):
        self.dest_file = str(dest_file)
        self.enc = np.fromfile(source_file, dtype=np.uint8)
        self.motioncomp = bool(motioncomp)
        self.result = None

    def process(self) -> bool:
        self.result, self.width, self.height, self.frames = decode_video(self.enc, self.motioncomp)
        return True

    def saveResult(self) -> None:
        self.result.tofile(self.dest_file)
