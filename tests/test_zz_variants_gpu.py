"""Instantiations of the tile kernel (ie_set_option("encode_variant", 0..7)): byte-identical streams required.

Variant 2 (packed f32x2 transform + lean quantise bookkeeping, imageencoder_b200/csrc/transform_fast.cuh) is the default
kernel; 0 is the scalar kernel it replaced and 1 the intermediate step, both kept for A/B timing (tools/ab_quick.py) and as
cross-checks.  Each variant runs in its own process against the CPU oracle and against variant 0
(tests/_variant_worker.py); the same arithmetic is also run on the CPU (tests/test_lean_variant_cpu.py).
First confirmed on a B200 in profiles/r1_variant_parity_v9.log; the variants 3..7, decode_variant 1 and me_variant 1
all passed their first B200 run at the end of round 1 (GPUTEST_r01.json) and are strict since."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.parametrize("variant", [1, 2])
def test_variant_streams_identical(variant):
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "_variant_worker.py"), str(variant)], capture_output=True, text=True,
                       timeout=900)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("variant", [3, 4, 5, 6, 7, 8, 9])
def test_experimental_variant_streams_identical(variant):
    """the noise / quantiser-1 cases of the worker overflow the reduced staging area and take the global-memory pack; the
    checker / two-level cases are tie-heavy (exact queue); 8 = the fused persistent stream kernel (encode_fused.cu)"""
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "_variant_worker.py"), str(variant)], capture_output=True, text=True,
                       timeout=180)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_decode_variant_pixels_identical():
    """ie_set_option("decode_variant", 1): packed f32x2 inverse transform + pixel stage (decode_blocks_lean_kernel); own process
    so that a fault in the experimental kernel cannot poison this process's CUDA context."""
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "_variant_worker.py"), "dec1"], capture_output=True, text=True,
                       timeout=180)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_me_variant_streams_identical():
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "_variant_worker.py"), "me1"], capture_output=True, text=True,
                       timeout=180)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_damaged_streams_decode_like_the_oracle():
    """300 damaged image streams (bit flips, truncation, trailing garbage) through the default decode path, in their own process"""
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "_variant_worker.py"), "corrupt"], capture_output=True, text=True,
                       timeout=300)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_copyout_variant_streams_identical(gpu, oracle_mod):
    """ie_set_option("copyout_variant", 0|1|2|3): 3 (a warp per tile image, word by word) is the default kernel, 2 (chunk-centric, four in flight) and
    0 the generic one it replaced; none may change a byte."""
    import numpy as np
    from conftest import INPUTS
    from imageencoder_b200.synth import synth_image
    L = gpu.lib()
    try:
        for mat in ("matrix8_1.txt", "matrix4_2.txt"):
            q = oracle_mod.read_matrix(INPUTS / mat)
            n = q.shape[0]
            for img in (synth_image(1024, 768, 31), synth_image(512, 384, 32, flat=True), np.full((64, 64), 128, np.uint8)):
                h, w = img.shape
                want = oracle_mod.image_encode(img, w, h, n, q, True, False)
                for cv in (0, 1, 2, 3):
                    assert L.ie_set_option(b"copyout_variant", cv) == 0
                    assert gpu.encode_image(img, w, h, q, True, False) == want, (mat, img.shape, cv)
    finally:
        L.ie_set_option(b"copyout_variant", 3)


def test_exact_parse_path_matches_oracle(gpu, oracle_mod):
    """ie_set_option("parse_variant", 1): the speculative parse's verdict is ignored, the exact transfer-function path
    (parse_exact_kernel: five phases behind grid barriers in one launch) finds the block offsets.  Images, a video (cursor mode,
    one parse per frame) and truncated streams."""
    import numpy as np
    from conftest import INPUTS
    from imageencoder_b200.synth import synth_image, synth_video
    L = gpu.lib()
    assert L.ie_set_option(b"parse_variant", 1) == 0
    try:
        for mat in ("matrix8_1.txt", "matrix8_2.txt", "matrix4_2.txt"):
            q = oracle_mod.read_matrix(INPUTS / mat)
            n = q.shape[0]
            for img in (synth_image(1024, 768, 41), synth_image(512, 384, 42, flat=True), np.full((64, 64), 128, np.uint8),
                        np.random.default_rng(9).integers(0, 256, (128, 96)).astype(np.uint8)):
                h, w = img.shape
                for rle in (True, False):
                    enc = oracle_mod.image_encode(img, w, h, n, q, rle, False)
                    want = oracle_mod.image_decode(enc, n)[0]
                    assert np.array_equal(gpu.decode_image(enc, n), want), (mat, img.shape, rle)
                    cut = enc[: max(160, len(enc) * 2 // 3)]            # the header stays whole
                    assert np.array_equal(gpu.decode_image(cut, n), oracle_mod.image_decode(cut, n)[0]), (mat, img.shape, rle, "cut")
        q = oracle_mod.read_matrix(INPUTS / "matrix8_1.txt")
        big = synth_image(4096, 2048, 1234)
        enc = gpu.encode_image(big, 4096, 2048, q, True, False)
        L.ie_set_option(b"parse_variant", 0)
        want = gpu.decode_image(enc, 8)
        L.ie_set_option(b"parse_variant", 1)
        assert np.array_equal(gpu.decode_image(enc, 8), want)
        q4 = oracle_mod.read_matrix(INPUTS / "matrix.txt")
        yuv = synth_video(64, 48, 5)
        venc = oracle_mod.video_encode(yuv, 64, 48, q4, True, 3, 16, False)
        assert np.array_equal(gpu.decode_video(venc, True)[0], oracle_mod.video_decode(venc, True)[0])
    finally:
        L.ie_set_option(b"parse_variant", 0)
