"""The BASELINE.json configurations at their FULL sizes against the UNMODIFIED reference: tests/golden/golden_configs.json
holds the sha256 of what oracle/_ref (the reference compiled by oracle/build_ref.sh; ImageEncoder.cpp:52-175,
ImageDecoder.cpp:55-122, VideoEncoder.cpp:22-111, VideoDecoder.cpp:33-62) wrote for the seeded synthetic inputs
(tests/golden/make_golden_configs.py, run in the build container).  Everything here goes through the C-ABI host entry
points; nothing is compared with another output of this repository."""
import hashlib
import json

import numpy as np
import pytest

from conftest import GOLDEN, INPUTS

pytestmark = pytest.mark.gpu
sha = lambda b: hashlib.sha256(bytes(b)).hexdigest()


@pytest.fixture(scope="module")
def golden():
    return json.loads((GOLDEN / "golden_configs.json").read_text())


def _image_config(gpu, oracle_mod, e, kinds):
    from imageencoder_b200.synth import synth_image
    W, H, N = e["W"], e["H"], e["block"]
    q = oracle_mod.read_matrix(INPUTS / e["matrix"])
    img = synth_image(W, H, e["seed"])
    assert sha(img) == e["input_sha256"], "synthetic input differs from the one the reference encoded"
    for kind in kinds:
        g = e[kind]
        enc = gpu.encode_image(img, W, H, q, True, kind == "huff")
        assert len(enc) == g["enc_bytes"], (kind, len(enc), g["enc_bytes"])
        assert sha(enc) == g["enc_sha256"], f"{kind}: encoded file differs from the reference's"
        if g["dec_sha256"] is not None:
            dec = gpu.decode_image(enc, N)
            assert dec.shape == (H, W)
            assert sha(dec.tobytes()) == g["dec_sha256"], f"{kind}: decoded pixels differ from the reference's"


def test_config2_8192_8x8(gpu, oracle_mod, golden):
    """BASELINE configs[1]: 8192x8192, 8x8, matrix8_1, RLE on, Huffman off"""
    _image_config(gpu, oracle_mod, golden["C2|8192x8192|matrix8_1|seed1234"], ("plain",))


def test_config3_16384_8x8_huffman_on_and_off(gpu, oracle_mod, golden):
    """BASELINE configs[2]: 16384x16384, 8x8, matrix8_2, RLE + Huffman (and the plain stream it is made from).  The input's
    Huffman code stays below 16 bits / 128 symbols per length, so the reference can decode its own file."""
    e = golden["C3|16384x16384|matrix8_2|seed1235"]
    assert e["huff"]["dictionary_overflows"] is False and e["huff"]["reverted"] is False
    _image_config(gpu, oracle_mod, e, ("plain", "huff"))


@pytest.mark.parametrize("seed", [2000, 2001, 2002])
def test_config4_4096_4x4(gpu, oracle_mod, golden, seed):
    """BASELINE configs[3]: images of the 1024 x 4096x4096 batch (4x4, matrix4_2), encode + decode round trip"""
    _image_config(gpu, oracle_mod, golden[f"C4|4096x4096|matrix4_2|seed{seed}"], ("plain",))


def test_config4_batch_entry_points(gpu, oracle_mod, golden):
    """the same three images through ie_encode_images / ie_decode_images (one launch sequence for the batch)"""
    from imageencoder_b200.synth import synth_image
    es = [golden[f"C4|4096x4096|matrix4_2|seed{s}"] for s in (2000, 2001, 2002)]
    q = oracle_mod.read_matrix(INPUTS / "matrix4_2.txt")
    imgs = np.stack([synth_image(4096, 4096, e["seed"]) for e in es])
    encs = gpu.encode_images(imgs, 4096, 4096, q, True, False)
    for enc, e in zip(encs, es):
        assert sha(enc) == e["plain"]["enc_sha256"]
    decs = gpu.decode_images(encs, 4)
    for dec, e in zip(decs, es):
        assert sha(np.asarray(dec).tobytes()) == e["plain"]["dec_sha256"]


def test_config5_video_240_frames(gpu, oracle_mod, golden):
    """BASELINE configs[4]: 1920x1088, 240 frames, GOP 12, merange 16, matrix.txt; decode with and without motion compensation"""
    from imageencoder_b200.synth import synth_video
    e = golden["C5|1920x1088x240|gop12|mer16|matrix|seed4000"]
    q = oracle_mod.read_matrix(INPUTS / e["matrix"])
    yuv = synth_video(e["W"], e["H"], e["frames"], e["seed"])
    assert sha(yuv) == e["input_sha256"]
    enc = gpu.encode_video(yuv, e["W"], e["H"], q, True, e["gop"], e["merange"], False)
    assert len(enc) == e["enc_bytes"]
    assert sha(enc) == e["enc_sha256"], "encoded clip differs from the reference's"
    for mc in (True, False):
        dec = gpu.decode_video(enc, mc)[0]
        assert sha(np.asarray(dec).tobytes()) == e[f"dec_mc{int(mc)}_sha256"], f"decoded clip (motioncompensation={int(mc)}) differs"
