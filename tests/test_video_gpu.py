"""Parity of the CUDA video path (C-ABI) against the CPU oracle: byte-identical stream, identical motion vectors
(observable through the stream), identical decoded frames with and without motion compensation."""
import hashlib
import json

import numpy as np
import pytest

from conftest import GOLDEN, INPUTS

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("W,H,F,gop,mer", [(64, 48, 7, 4, 16), (128, 96, 9, 3, 8), (176, 144, 6, 6, 32),
                                            (64, 64, 5, 1, 16), (64, 48, 6, 7, 2), (64, 48, 5, 4, 1), (96, 64, 6, 4, 10),
                                            (64, 48, 70, 1, 4), (64, 48, 131, 2, 8)])      # more GOPs than one launch batch (64)
@pytest.mark.parametrize("huffman", [False, True])
def test_video_matches_oracle(gpu, oracle_mod, W, H, F, gop, mer, huffman):
    from imageencoder_b200.synth import synth_video
    q = oracle_mod.read_matrix(INPUTS / "matrix.txt")
    yuv = synth_video(W, H, F, 4000)
    got = gpu.encode_video(yuv, W, H, q, True, gop, mer, huffman)
    want = oracle_mod.video_encode(yuv, W, H, q, True, gop, mer, huffman)
    assert got == want, f"stream differs ({len(got)} vs {len(want)} bytes)"
    key = f"synth4000|{W}x{H}x{F}|gop{gop}|mer{mer}|{'huff' if huffman else 'plain'}"
    gold = json.loads((GOLDEN / "golden.json").read_text())["video"].get(key)
    if gold:
        assert hashlib.sha256(got).hexdigest() == gold["enc_sha256"]
    if huffman:
        import ctypes as C
        # only decodable when the dictionary header did not overflow (Huffman.cpp:39-42)
        plain = oracle_mod.video_encode(yuv, W, H, q, True, gop, mer, False)
        if gold and gold["dec_mc1_sha256"] is None:
            return
    L = gpu.lib()
    try:
        for variant in (1, 0):              # whole-stream parse (default) and the frame-by-frame path it falls back to
            assert L.ie_set_option(b"video_decode_variant", variant) == 0
            for mc in (True, False):
                before = L.ie_stat(b"video_decode_whole_stream")
                dec, w, h, f = gpu.decode_video(got, mc)
                if variant == 1:        # a well-formed stream must take the whole-stream path, not silently fall back
                    assert L.ie_stat(b"video_decode_whole_stream") == before + 1
                odec = oracle_mod.video_decode(want, mc)[0]
                assert (w, h, f) == (W, H, F)
                assert np.array_equal(dec, odec), f"decoded frames differ (motioncompensation={mc}, variant={variant})"
                if gold:
                    assert hashlib.sha256(dec.tobytes()).hexdigest() == gold["dec_mc1_sha256" if mc else "dec_mc0_sha256"]
    finally:
        L.ie_set_option(b"video_decode_variant", 1)


def test_video_rle_off_and_other_matrix(gpu, oracle_mod):
    from imageencoder_b200.synth import synth_video
    q = oracle_mod.read_matrix(INPUTS / "matrix4_2.txt")
    yuv = synth_video(80, 48, 5, 4001)
    for rle in (True, False):
        got = gpu.encode_video(yuv, 80, 48, q, rle, 3, 16, False)
        assert got == oracle_mod.video_encode(yuv, 80, 48, q, rle, 3, 16, False)
        assert np.array_equal(gpu.decode_video(got, True)[0], oracle_mod.video_decode(got, True)[0])


def test_video_scene_cut_and_saturation(gpu, oracle_mod):
    """frames with nothing in common (large residuals, DC down to -766) and saturated pixels."""
    rng = np.random.default_rng(5)
    W, H, F = 64, 48, 6
    fsz = W * H * 3 // 2
    yuv = np.full(F * fsz, 0x80, np.uint8)
    for t in range(F):
        y = rng.integers(0, 256, (H, W)).astype(np.uint8) if t % 2 else np.full((H, W), 255 * (t // 2 % 2), np.uint8)
        yuv[t * fsz: t * fsz + W * H] = y.reshape(-1)
    q = oracle_mod.read_matrix(INPUTS / "matrix.txt")
    got = gpu.encode_video(yuv, W, H, q, True, 6, 16, False)
    assert got == oracle_mod.video_encode(yuv, W, H, q, True, 6, 16, False)
    assert np.array_equal(gpu.decode_video(got, True)[0], oracle_mod.video_decode(got, True)[0])


@pytest.mark.parametrize("keep", [0.9, 0.45])
def test_truncated_video_decodes_like_the_reference(gpu, oracle_mod, keep):
    """a cut file: fields past the end read as zero bits (BitStream.cpp:17-20), later frames decode from an exhausted stream"""
    from imageencoder_b200.synth import synth_video
    W, H, F = 96, 64, 7
    q = oracle_mod.read_matrix(INPUTS / "matrix.txt")
    yuv = synth_video(W, H, F, 4000)
    enc = oracle_mod.video_encode(yuv, W, H, q, True, 3, 8, False)
    cut = enc[: int(len(enc) * keep)]
    for mc in (True, False):
        want = oracle_mod.video_decode(cut, mc)[0]
        got, w, h, f = gpu.decode_video(cut, mc)
        assert (w, h, f) == (W, H, F)
        assert np.array_equal(got, want), f"motioncompensation={mc}"


def test_static_and_flashing_clips(gpu, oracle_mod):
    """identical frames (zero motion, residual = the -128 offset alone) and frames alternating black/white (residual +-255)"""
    W, H, F = 64, 48, 6
    q = oracle_mod.read_matrix(INPUTS / "matrix.txt")
    fsz = W * H * 3 // 2
    rng = np.random.default_rng(9)
    base = rng.integers(0, 256, W * H).astype(np.uint8)
    for name in ("static", "flash"):
        yuv = np.full(F * fsz, 0x80, np.uint8)
        for t in range(F):
            yuv[t * fsz: t * fsz + W * H] = base if name == "static" else (255 if t % 2 else 0)
        got = gpu.encode_video(yuv, W, H, q, True, 4, 16, False)
        want = oracle_mod.video_encode(yuv, W, H, q, True, 4, 16, False)
        assert got == want, name
        dec, w, h, f = gpu.decode_video(got, True)
        assert np.array_equal(dec, oracle_mod.video_decode(want, True)[0]), name


def test_video_errors(gpu):
    from imageencoder_b200 import IEError
    with pytest.raises(IEError):
        gpu.encode_video(np.zeros(40 * 40 * 3 // 2, np.uint8), 40, 40, np.full(16, 2))   # not a multiple of 16


def test_motion_vectors_and_reconstruction_direct(gpu, oracle_mod):
    """Motion vectors (Block.cpp:267-339) and the encoder-side reconstruction (Frame.cpp:218-242) compared directly, not only
    through the stream: ie_encode_video_dev's optional mvec output [frames][MacroBlocks][2] and the Y planes it rebuilds in place."""
    import torch
    from imageencoder_b200 import device
    from imageencoder_b200.synth import synth_video
    W, H, F, gop, mer = 96, 64, 7, 3, 16
    q = oracle_mod.read_matrix(INPUTS / "matrix.txt")
    yuv = synth_video(W, H, F, 4100)
    want, mv, rec = oracle_mod.video_encode(yuv, W, H, q, True, gop, mer, False, stages=True)
    nmb = (W // 16) * (H // 16)
    fsz = W * H * 3 // 2
    d_yuv = torch.from_numpy(np.array(yuv, dtype=np.uint8).reshape(-1).copy()).cuda()
    cap = int(gpu.lib().ie_max_encoded_bytes(W, H, 4, F)) + 4096
    d_out = torch.zeros(cap, dtype=torch.uint8, device="cuda")
    d_bits = torch.zeros(1, dtype=torch.int64, device="cuda")
    d_mv = torch.full((F * nmb * 2,), 77, dtype=torch.int16, device="cuda")
    sess = device.Session(device.Session.VIDEO_ENCODE, W, H, 4, F)
    device.encode_video_dev(sess, d_yuv, W, H, q, True, gop, mer, d_out, d_bits, lead_bit=True, d_mvecs=d_mv)
    torch.cuda.synchronize()
    nbytes = (int(d_bits.item()) + 7) // 8
    assert d_out[:nbytes].cpu().numpy().tobytes() == want
    got_mv = d_mv.cpu().numpy().reshape(F, nmb, 2)
    assert np.array_equal(got_mv, np.asarray(mv).reshape(F, nmb, 2)), "motion vectors differ from the oracle's"
    got = d_yuv.cpu().numpy().reshape(F, fsz)[:, : W * H]
    exp = np.asarray(rec, dtype=np.uint8).reshape(F, fsz)[:, : W * H]
    assert np.array_equal(got, exp), "encoder-side reconstruction (Y planes) differs from the oracle's"
