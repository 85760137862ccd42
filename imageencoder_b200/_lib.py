"""ctypes loader for libimageencoder_b200.so.  Fails loudly: no fallback of any kind."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
_LIB_NAME = "libimageencoder_b200.so"

IE_OK, IE_EINVAL, IE_ENODEVICE, IE_ECUDA, IE_ENOSPC, IE_EFORMAT, IE_ENOMEM = 0, -1, -2, -3, -4, -5, -6

_u8p = C.POINTER(C.c_uint8)
_u16p = C.POINTER(C.c_uint16)
_i16p = C.POINTER(C.c_int16)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_szp = C.POINTER(C.c_size_t)
_vp = C.c_void_p

# name -> (restype, argtypes).  Every symbol include/imageencoder_b200.h declares must be listed here
# (tests/test_capi_symbols.py checks both directions).
SIGNATURES = {
    "ie_init": (C.c_int, [C.c_int]),
    "ie_shutdown": (None, []),
    "ie_last_error": (C.c_char_p, []),
    "ie_version": (C.c_char_p, []),
    "ie_max_encoded_bytes": (C.c_size_t, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]),
    "ie_encode_image": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.c_uint32, _u16p, C.c_int, C.c_int, _vp, C.c_size_t, _szp]),
    "ie_decode_image": (C.c_int, [_vp, C.c_size_t, C.c_uint32, _vp, C.c_size_t, _u32p, _u32p]),
    "ie_encode_images": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _u16p, C.c_int, C.c_int, _vp,
                                   C.c_size_t, _szp]),
    "ie_decode_images": (C.c_int, [_vp, C.c_size_t, _szp, C.c_uint32, C.c_uint32, _vp, C.c_size_t, _u32p, _u32p]),
    "ie_encode_video": (C.c_int, [_vp, C.c_size_t, C.c_uint32, C.c_uint32, _u16p, C.c_int, C.c_uint32, C.c_uint32, C.c_int,
                                  _vp, C.c_size_t, _szp]),
    "ie_decode_video": (C.c_int, [_vp, C.c_size_t, C.c_int, _vp, C.c_size_t, _szp, _u32p, _u32p, _u32p]),
    "ie_session_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]),
    "ie_session_destroy": (None, [_vp]),
    "ie_encode_image_dev": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, _u16p, C.c_int, C.c_int, C.c_int, C.c_uint64, _vp,
                                      C.c_size_t, _vp, _vp]),
    "ie_encode_images_dev": (C.c_int, [_vp, _vp, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, _u16p, C.c_int, C.c_int, _vp, C.c_size_t,
                                       _vp, _vp]),
    "ie_encode_image_begin_dev": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, _u16p, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "ie_encode_image_end_dev": (C.c_int, [_vp, _vp, C.c_uint32, _vp, C.c_size_t, _vp, _vp, _vp]),
    "ie_session_set_header_height": (C.c_int, [_vp, C.c_uint32]),
    "ie_image_bits_dev": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, _u16p, C.c_int, _vp, _vp]),
    "ie_decode_image_dev": (C.c_int, [_vp, _vp, C.c_size_t, C.c_uint64, _vp, C.c_size_t, _u32p, _u32p, _vp]),
    "ie_parse_image_header": (C.c_int, [_vp, C.c_size_t, C.c_uint64, C.c_uint32, _vp]),
    "ie_decode_image_with_header_dev": (C.c_int, [_vp, _vp, _vp, C.c_size_t, _vp, C.c_size_t, _vp]),
    "ie_decode_shard_spec_bytes": (C.c_size_t, [C.c_size_t, C.c_uint32, C.c_uint32, _szp]),
    "ie_decode_image_shard_begin_dev": (C.c_int, [_vp, _vp, _vp, C.c_size_t, C.c_uint32, C.c_uint32, _vp, _vp]),
    "ie_decode_image_shard_end_dev": (C.c_int, [_vp, _vp, _vp, C.c_size_t, C.c_uint32, _vp, C.c_uint32, C.c_uint32, _vp, C.c_size_t, _vp]),
    "ie_decode_images_dev": (C.c_int, [_vp, _vp, C.c_size_t, _szp, C.c_uint32, C.c_uint64, _vp, C.c_size_t, _u32p, _u32p, _vp]),
    "ie_huffman_encode_dev": (C.c_int, [_vp, _vp, C.c_size_t, _vp, C.c_size_t, _szp, _vp]),
    "ie_byte_histogram_async_dev": (C.c_int, [_vp, _vp, C.c_size_t, _vp, _vp, _vp]),
    "ie_huffman_encode_shard_async_dev": (C.c_int, [_vp, _vp, C.c_size_t, _vp, _vp, C.c_int, _vp, C.c_size_t, _vp, _vp]),
    "ie_huffman_encode_async_dev": (C.c_int, [_vp, _vp, C.c_size_t, _vp, C.c_size_t, _vp, _vp]),
    "ie_huffman_decode_dev": (C.c_int, [_vp, _vp, C.c_size_t, _vp, C.c_size_t, _szp, _u64p, _vp]),
    "ie_huffman_encode_shard_dev": (C.c_int, [_vp, _vp, C.c_size_t, _u32p, _u64p, C.c_int, _vp, C.c_size_t, _vp, _vp]),
    "ie_byte_histogram_dev": (C.c_int, [_vp, C.c_size_t, _u32p, _u64p, _vp]),
    "ie_encode_video_dev": (C.c_int, [_vp, _vp, C.c_size_t, C.c_uint32, C.c_uint32, _u16p, C.c_int, C.c_uint32, C.c_uint32,
                                      C.c_int, _vp, C.c_size_t, _vp, _vp, _vp]),
    "ie_session_set_video_shard": (C.c_int, [_vp, C.c_uint32, C.c_int]),
    "ie_decode_video_dev": (C.c_int, [_vp, _vp, C.c_size_t, C.c_uint64, C.c_int, _vp, C.c_size_t, _u32p, _u32p, _u32p, _vp]),
    "ie_stream_shift_dev": (C.c_int, [_vp, _vp, _vp, C.c_size_t, _vp]),
    "ie_comm_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_size_t]),
    "ie_comm_handle_bytes": (C.c_size_t, []),
    "ie_comm_export": (C.c_int, [_vp, _vp]),
    "ie_comm_connect": (C.c_int, [_vp, _vp]),
    "ie_comm_destroy": (None, [_vp]),
    "ie_comm_exchange_totals_dev": (C.c_int, [_vp, _vp, _vp, _vp]),
    "ie_comm_totals_dev": (C.c_int, [_vp, C.POINTER(_vp)]),
    "ie_comm_copy_totals": (C.c_int, [_vp, _vp, _vp]),
    "ie_encode_image_shard_dev": (C.c_int, [_vp, _vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, _u16p, C.c_int, C.c_int, _vp,
                                            C.c_size_t, _vp, _vp, _vp]),
    "ie_comm_stitch_dev": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "ie_comm_stitched_stream": (C.c_int, [_vp, C.POINTER(_vp), _szp]),
    "ie_comm_copy_stitched": (C.c_int, [_vp, _vp, C.c_size_t, _vp]),
    "ie_set_option": (C.c_int, [C.c_char_p, C.c_int]),
    "ie_kernel_launch_count": (C.c_uint64, []),
    "ie_stat": (C.c_uint64, [C.c_char_p]),
}


class IEError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"imageencoder_b200 error {code}: {msg}")
        self.code = code


def lib_path() -> Path:
    return Path(os.environ.get("IMAGEENCODER_B200_LIB", _HERE / _LIB_NAME))


_lib = None


def lib() -> C.CDLL:
    """Loads the C-ABI library.  Raises (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        p = lib_path()
        if not p.exists():
            raise ImportError(f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              f"(make -C imageencoder_b200/csrc).  There is no CPU fallback.")
        L = C.CDLL(str(p))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)      # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != IE_OK:
        raise IEError(rc, lib().ie_last_error().decode("utf-8", "replace"))


def launch_count() -> int:
    return int(lib().ie_kernel_launch_count())
