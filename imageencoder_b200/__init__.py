"""imageencoder_b200 -- B200-native block-transform hot path of ThenTech/ImageEncoder.

The product is the C-ABI shared library ``libimageencoder_b200.so`` (hand-written sm_100a CUDA + C++ host code,
``include/imageencoder_b200.h``).  This package is the thin Python face used by the tests, ``bench.py`` and the
multi-GPU launcher: ctypes bindings, torch only for device memory / streams / ``torch.distributed`` plumbing.

There is no CPU fallback: importing works anywhere, but every compute call raises if the library or a B200 is missing.
"""
from ._lib import IEError, lib, lib_path, launch_count  # noqa: F401
from .codec import (  # noqa: F401
    ImageDecoder,
    ImageEncoder,
    VideoDecoder,
    VideoEncoder,
    decode_image,
    decode_images,
    decode_video,
    encode_image,
    encode_images,
    encode_video,
    read_matrix,
)

__version__ = "0.1.0"
