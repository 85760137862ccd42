"""Parity of the CUDA image path (through the C-ABI) against the CPU oracle: byte-identical .enc, identical pixels."""
import hashlib
import json

import numpy as np
import pytest

from conftest import GOLDEN, INPUTS

pytestmark = pytest.mark.gpu

SAMPLES = {"ex0": (8, 8), "ex1": (936, 936), "ex2": (512, 512), "ex3": (400, 400), "ex4": (4096, 912), "ex6": (512, 256)}


def _mat(oracle_mod, name):
    return oracle_mod.read_matrix(INPUTS / name)


def _cases():
    from imageencoder_b200.synth import synth_image
    rng = np.random.default_rng(7)
    yield "synth", synth_image(256, 192, 11)
    yield "synth_flat", synth_image(256, 192, 12, flat=True)
    yield "all128", np.full((64, 64), 128, np.uint8)
    yield "all0", np.zeros((64, 64), np.uint8)
    yield "all255", np.full((64, 64), 255, np.uint8)
    yield "checker", ((np.indices((64, 64)).sum(0) & 1) * 255).astype(np.uint8)
    yield "noise", rng.integers(0, 256, (128, 96)).astype(np.uint8)
    yield "stripes", np.tile(np.array([0, 255, 255, 0, 0, 0, 255, 255], np.uint8), (72, 10))
    yield "tiny", rng.integers(0, 256, (8, 8)).astype(np.uint8)


@pytest.mark.parametrize("name", list(SAMPLES))
@pytest.mark.parametrize("huffman", [False, True])
def test_reference_samples_4x4(gpu, oracle_mod, name, huffman):
    W, H = SAMPLES[name]
    raw = np.fromfile(INPUTS / f"{name}.raw", dtype=np.uint8)
    q = _mat(oracle_mod, "matrix.txt")
    got = gpu.encode_image(raw, W, H, q, True, huffman)
    want = oracle_mod.image_encode(raw, W, H, 4, q, True, huffman)
    assert len(got) == len(want)
    assert got == want
    gold = json.loads((GOLDEN / "golden.json").read_text())["images"][f"{name}|matrix.txt|rle1|{'huff' if huffman else 'plain'}"]
    assert hashlib.sha256(got).hexdigest() == gold["enc_sha256"]
    dec = gpu.decode_image(got, 4)
    assert hashlib.sha256(dec.tobytes()).hexdigest() == gold["dec_sha256"]


@pytest.mark.parametrize("matrix", ["matrix.txt", "matrix4_2.txt", "matrix8_1.txt", "matrix8_2.txt"])
@pytest.mark.parametrize("rle", [True, False])
@pytest.mark.parametrize("huffman", [False, True])
def test_encode_decode_matches_oracle(gpu, oracle_mod, matrix, rle, huffman):
    q = _mat(oracle_mod, matrix)
    N = q.shape[0]
    for name, img in _cases():
        H, W = img.shape
        got = gpu.encode_image(img, W, H, q, rle, huffman)
        want = oracle_mod.image_encode(img, W, H, N, q, rle, huffman)
        assert got == want, f"{name}: encoded stream differs (len {len(got)} vs {len(want)})"
        if huffman and oracle_mod.huffman_header_overflows(oracle_mod.image_encode_plain(img, W, H, N, q, rle, lead_bit=False)[0]):
            continue          # the reference cannot decode its own output here (Huffman.cpp:39-42); encoder parity only
        dec = gpu.decode_image(got, N)
        odec, _, _ = oracle_mod.image_decode(want, N)
        assert np.array_equal(dec, odec), f"{name}: decoded pixels differ"


def test_quant_one_widest_fields(gpu, oracle_mod):
    """quant = 1 everywhere gives the widest coefficients (up to 13 bits at 8x8)."""
    rng = np.random.default_rng(3)
    img = rng.integers(0, 2, (64, 64)).astype(np.uint8) * 255
    for N in (4, 8):
        q = np.ones((N, N), np.uint16)
        got = gpu.encode_image(img, 64, 64, q, True, False)
        assert got == oracle_mod.image_encode(img, 64, 64, N, q, True, False)
        assert np.array_equal(gpu.decode_image(got, N), oracle_mod.image_decode(got, N)[0])


def test_large_synthetic_8x8(gpu, oracle_mod):
    from imageencoder_b200.synth import synth_image
    img = synth_image(1024, 1024, 1234)
    q = _mat(oracle_mod, "matrix8_1.txt")
    got = gpu.encode_image(img, 1024, 1024, q, True, False)
    want = oracle_mod.image_encode(img, 1024, 1024, 8, q, True, False)
    assert got == want
    assert np.array_equal(gpu.decode_image(got, 8), oracle_mod.image_decode(want, 8)[0])


@pytest.mark.parametrize("N,W,H,huffman", [(8, 4096, 2056, False), (4, 2052, 4100, False), (8, 2048, 4096, True)])
def test_striped_host_pipeline_matches_oracle(gpu, oracle_mod, N, W, H, huffman):
    """ie_encode_image copies/encodes/returns images above 8 MiB in stripes of block rows (copy-compute pipeline): the bytes
    must be those of a one-launch encode, i.e. the oracle's."""
    from imageencoder_b200.synth import synth_image
    img = synth_image(W, H, 77)
    q = _mat(oracle_mod, "matrix8_1.txt" if N == 8 else "matrix.txt")
    got = gpu.encode_image(img, W, H, q, True, huffman)
    want = oracle_mod.image_encode(img, W, H, N, q, True, huffman)
    assert hashlib.sha256(got).hexdigest() == hashlib.sha256(want).hexdigest()
    if not huffman:
        assert np.array_equal(gpu.decode_image(got, N), np.asarray(oracle_mod.image_decode(want, N)[0]))


@pytest.mark.parametrize("N,W,H", [(8, 32760, 8), (8, 8, 32760), (4, 32764, 4), (4, 4, 32764), (8, 8, 8), (4, 4, 4), (8, 1000, 24)])
def test_extreme_shapes(gpu, oracle_mod, N, W, H):
    """largest widths/heights the 15-bit header fields hold (ImageBase.hpp:76), single-block images, tiles straddling rows"""
    rng = np.random.default_rng(W * 7 + H)
    img = rng.integers(0, 256, (H, W)).astype(np.uint8)
    q = _mat(oracle_mod, "matrix8_1.txt" if N == 8 else "matrix.txt")
    for huffman in (False, True):
        got = gpu.encode_image(img, W, H, q, True, huffman)
        want = oracle_mod.image_encode(img, W, H, N, q, True, huffman)
        assert got == want, f"huffman={huffman}"
    dec = gpu.decode_image(oracle_mod.image_encode(img, W, H, N, q, True, False), N)
    assert np.array_equal(dec, np.asarray(oracle_mod.image_decode(oracle_mod.image_encode(img, W, H, N, q, True, False), N)[0]))


@pytest.mark.parametrize("N", [4, 8])
@pytest.mark.parametrize("keep", [0.97, 0.5, 0.1])
def test_truncated_stream_decodes_like_the_reference(gpu, oracle_mod, N, keep):
    """reads past the end of the stream return 0 bits (BitStream.cpp:17-20): a truncated file still decodes, identically"""
    from imageencoder_b200.synth import synth_image
    W, H = 256, 192
    img = synth_image(W, H, 31)
    q = _mat(oracle_mod, "matrix8_1.txt" if N == 8 else "matrix.txt")
    enc = oracle_mod.image_encode(img, W, H, N, q, True, False)
    cut = enc[: max(200, int(len(enc) * keep))]
    want = np.asarray(oracle_mod.image_decode(cut, N)[0])
    got = gpu.decode_image(cut, N)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("N", [4, 8])
@pytest.mark.parametrize("keep,extra", [(0.8, 0), (1.0, 37)])
def test_truncated_or_padded_huffman_file_decodes_like_the_reference(gpu, oracle_mod, N, keep, extra):
    """Huffman-coded file cut inside the payload / followed by stray bytes: the reference decodes symbols until the input is
    exhausted (Huffman.cpp:376-383) and then reads the blocks from whatever that produced"""
    from imageencoder_b200.synth import synth_image
    W, H = 256, 192
    img = synth_image(W, H, 32)
    q = _mat(oracle_mod, "matrix8_1.txt" if N == 8 else "matrix.txt")
    plain = oracle_mod.image_encode_plain(img, W, H, N, q, True, lead_bit=False)[0]
    if oracle_mod.huffman_header_overflows(plain):
        pytest.skip("the reference cannot decode its own dictionary header here (Huffman.cpp:39-42)")
    enc = oracle_mod.image_encode(img, W, H, N, q, True, True)
    data = enc[: int(len(enc) * keep)] + bytes((i * 73 + 5) & 0xFF for i in range(extra))
    want = np.asarray(oracle_mod.image_decode(data, N)[0])
    got = gpu.decode_image(data, N)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("kind", ["zeros", "white", "mid", "stripes", "checker", "ramp"])
def test_degenerate_images(gpu, oracle_mod, kind):
    """constant and maximally busy images: all-zero blocks (bit_len 0), DC-only blocks, the widest AC fields, RLE on and off"""
    W, H = 192, 160
    yy, xx = np.mgrid[0:H, 0:W]
    img = {"zeros": np.zeros((H, W)), "white": np.full((H, W), 255), "mid": np.full((H, W), 128), "stripes": (xx % 2) * 255,
           "checker": ((xx + yy) % 2) * 255, "ramp": (xx + yy) % 256}[kind].astype(np.uint8)
    for mat in ("matrix.txt", "matrix8_1.txt"):
        q = _mat(oracle_mod, mat)
        N = q.shape[0]
        for rle in (True, False):
            for huffman in (False, True):
                got = gpu.encode_image(img, W, H, q, rle, huffman)
                want = oracle_mod.image_encode(img, W, H, N, q, rle, huffman)
                assert got == want, f"{kind} {mat} rle={rle} huffman={huffman}"
            plain = oracle_mod.image_encode(img, W, H, N, q, rle, False)
            assert np.array_equal(gpu.decode_image(plain, N), np.asarray(oracle_mod.image_decode(plain, N)[0])), f"{kind} {mat} rle={rle}"


def test_errors(gpu):
    from imageencoder_b200 import IEError
    with pytest.raises(IEError):
        gpu.encode_image(np.zeros(60, np.uint8), 10, 6, np.full(16, 2))          # not a multiple of the block size
    with pytest.raises(IEError):
        gpu.encode_image(np.zeros(64, np.uint8), 8, 8, np.zeros(16))             # quant entry 0


def test_host_entry_points_are_thread_safe(gpu, oracle_mod):
    """include/imageencoder_b200.h: the host entry points may be called from several threads.  Two threads encode and decode
    images of the SAME shape at the same time (the case where a shared cached session would hand both the same staging
    buffers and streams); every result must be the oracle's."""
    import threading
    from imageencoder_b200.synth import synth_image
    q = _mat(oracle_mod, "matrix8_1.txt")
    W, H = 1024, 512
    imgs = [synth_image(W, H, 500 + i) for i in range(4)]
    want = [oracle_mod.image_encode(im, W, H, 8, q, True, False) for im in imgs]
    want_dec = [oracle_mod.image_decode(w, 8)[0] for w in want]
    errors = []

    def worker(tid):
        try:
            for rep in range(12):
                i = (tid + rep) % len(imgs)
                enc = gpu.encode_image(imgs[i], W, H, q, True, False)
                if enc != want[i]:
                    errors.append(f"thread {tid} rep {rep}: stream differs")
                    return
                if not np.array_equal(gpu.decode_image(enc, 8), want_dec[i]):
                    errors.append(f"thread {tid} rep {rep}: pixels differ")
                    return
        except Exception as e:      # noqa: BLE001
            errors.append(f"thread {tid}: {e!r}")

    ts = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors
    gpu.lib().ie_shutdown()          # drops the idle cached sessions; the next call re-initialises the device tables
    assert gpu.encode_image(imgs[0], W, H, q, True, False) == want[0]


def test_batch_decode_with_one_short_stream(gpu, oracle_mod):
    """ie_decode_images: a stream cut inside its header must not truncate the headers of the other streams of the batch"""
    from imageencoder_b200.synth import synth_image
    q = _mat(oracle_mod, "matrix8_1.txt")
    W, H = 256, 128
    imgs = [synth_image(W, H, 600 + i) for i in range(3)]
    encs = [oracle_mod.image_encode(im, W, H, 8, q, True, False) for im in imgs]
    good = gpu.decode_images(encs, 8)
    for d, e in zip(good, encs):
        assert np.array_equal(d, oracle_mod.image_decode(e, 8)[0])
    # the header of an 8x8 stream with 8-bit quantiser entries is 549 bits = 69 bytes: cut the middle stream to 70 bytes
    cut = [encs[0], encs[1][:70], encs[2]]
    got = gpu.decode_images(cut, 8)
    for d, e in zip(got, cut):
        assert np.array_equal(d, oracle_mod.image_decode(e, 8)[0])


def test_decode_with_header_parsed_on_the_host(gpu, oracle_mod):
    """ie_parse_image_header + ie_decode_image_with_header_dev: no header read-back, same pixels"""
    import torch
    from imageencoder_b200 import device
    from imageencoder_b200.synth import synth_image
    for mat, (W, H) in (("matrix8_1.txt", (512, 384)), ("matrix4_2.txt", (256, 128))):
        q = _mat(oracle_mod, mat)
        N = q.shape[0]
        img = synth_image(W, H, 700)
        for rle in (True, False):
            enc = oracle_mod.image_encode(img, W, H, N, q, rle, False)
            hdr = device.parse_image_header(enc[:160], N)
            assert (hdr.width, hdr.height, hdr.block, hdr.use_rle) == (W, H, N, int(rle))
            assert list(hdr.quant[: N * N]) == [int(v) for v in np.asarray(q).reshape(-1)]
            d_enc = torch.zeros(len(enc) + 32, dtype=torch.uint8, device="cuda")
            d_enc[: len(enc)] = torch.frombuffer(bytearray(enc), dtype=torch.uint8).cuda()
            d_raw = torch.zeros(W * H, dtype=torch.uint8, device="cuda")
            sess = device.Session(device.Session.IMAGE_DECODE, W, H, N)
            device.decode_image_with_header_dev(sess, hdr, d_enc, len(enc), d_raw)
            torch.cuda.synchronize()
            assert np.array_equal(d_raw.cpu().numpy().reshape(H, W), oracle_mod.image_decode(enc, N)[0])
