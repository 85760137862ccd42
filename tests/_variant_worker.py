"""Runs in its own process (tests/test_zz_variants_gpu.py): encodes with ie_set_option("encode_variant", V) and compares the
stream with the default kernel's and with the CPU oracle's; `dec<V>` does the same for ie_set_option("decode_variant", V)
and the decoded pixels.  Exit code 0 = identical."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main(variant: int) -> int:
    import imageencoder_b200 as ie
    import oracle
    from imageencoder_b200 import _lib
    from imageencoder_b200.synth import synth_image, synth_video

    L = ie.lib()
    _lib.check(L.ie_init(0))
    inputs = ROOT / "tests" / "golden" / "inputs"
    rng = np.random.default_rng(5)
    bad = 0
    for mat in ("matrix8_1.txt", "matrix8_2.txt", "matrix.txt", "matrix4_2.txt"):
        q = ie.read_matrix(inputs / mat)
        n = q.shape[0]
        cases = [("synth", synth_image(512, 384, 21)), ("flat", synth_image(512, 384, 22, flat=True)),
                 ("noise", rng.integers(0, 256, (128, 96)).astype(np.uint8)),
                 ("checker", ((np.indices((64, 64)).sum(0) & 1) * 255).astype(np.uint8)),
                 ("all128", np.full((64, 64), 128, np.uint8)), ("extremes", (rng.integers(0, 2, (64, 128)) * 255).astype(np.uint8))]
        for qq, qname in ((q, mat), (np.ones_like(q), "ones")):
            for name, img in cases:
                h, w = img.shape
                for rle in (True, False):
                    _lib.check(L.ie_set_option(b"encode_variant", 0))
                    base = ie.encode_image(img, w, h, qq, rle, False)
                    _lib.check(L.ie_set_option(b"encode_variant", variant))
                    got = ie.encode_image(img, w, h, qq, rle, False)
                    want = oracle.image_encode(img, w, h, n, qq, rle, False)
                    if got != base or got != want:
                        print(f"variant {variant}: {qname} {name} rle={rle}: differs (default==oracle: {base == want})")
                        bad += 1
    # larger image: the variant against the default kernel only (the oracle would take too long)
    q = ie.read_matrix(inputs / "matrix8_1.txt")
    img = synth_image(4096, 2048, 1234)
    _lib.check(L.ie_set_option(b"encode_variant", 0))
    base = ie.encode_image(img, 4096, 2048, q, True, False)
    _lib.check(L.ie_set_option(b"encode_variant", variant))
    if ie.encode_image(img, 4096, 2048, q, True, False) != base:
        print(f"variant {variant}: 4096x2048 differs from the default kernel")
        bad += 1
    # video I-frames go through the same kernel (4x4)
    q = ie.read_matrix(inputs / "matrix.txt")
    yuv = synth_video(64, 48, 5)
    if ie.encode_video(yuv, 64, 48, q, True, 3, 16, False) != oracle.video_encode(yuv, 64, 48, q, True, 3, 16, False):
        print(f"variant {variant}: video stream differs")
        bad += 1
    _lib.check(L.ie_set_option(b"encode_variant", 2))
    print(f"variant {variant}: {'ok' if not bad else f'{bad} mismatches'}")
    return 1 if bad else 0


def main_decode(variant: int) -> int:
    import imageencoder_b200 as ie
    import oracle
    from imageencoder_b200 import _lib
    from imageencoder_b200.synth import synth_image, synth_video

    L = ie.lib()
    _lib.check(L.ie_init(0))
    inputs = ROOT / "tests" / "golden" / "inputs"
    rng = np.random.default_rng(6)
    bad = 0

    def both(stream, n):
        _lib.check(L.ie_set_option(b"decode_variant", 0))
        base = ie.decode_image(stream, n)
        _lib.check(L.ie_set_option(b"decode_variant", variant))
        return base, ie.decode_image(stream, n)

    for mat in ("matrix8_1.txt", "matrix8_2.txt", "matrix.txt", "matrix4_2.txt"):
        q = ie.read_matrix(inputs / mat)
        n = q.shape[0]
        cases = [("synth", synth_image(512, 384, 21)), ("flat", synth_image(512, 384, 22, flat=True)),
                 ("noise", rng.integers(0, 256, (128, 96)).astype(np.uint8)),
                 ("checker", ((np.indices((64, 64)).sum(0) & 1) * 255).astype(np.uint8)),
                 ("all128", np.full((64, 64), 128, np.uint8)), ("extremes", (rng.integers(0, 2, (64, 128)) * 255).astype(np.uint8))]
        for qq, qname in ((q, mat), (np.ones_like(q), "ones"), (np.full_like(q, 255), "255s")):
            for name, img in cases:
                h, w = img.shape
                stream = oracle.image_encode(img, w, h, n, qq, True, False)
                want = oracle.image_decode(stream, n)[0]
                base, got = both(stream, n)
                if not (np.array_equal(got, base) and np.array_equal(got, want)):
                    print(f"decode variant {variant}: {qname} {name}: differs (default==oracle: {np.array_equal(base, want)})")
                    bad += 1
    # larger image: the variant against the default kernel only
    q = ie.read_matrix(inputs / "matrix8_1.txt")
    _lib.check(L.ie_set_option(b"decode_variant", 0))
    stream = ie.encode_image(synth_image(4096, 2048, 1234), 4096, 2048, q, True, False)
    base, got = both(stream, 8)
    if not np.array_equal(base, got):
        print(f"decode variant {variant}: 4096x2048 differs from the default kernel")
        bad += 1
    # video: I-frames take the variant, P-frames (add mode) stay on the default kernel
    q = ie.read_matrix(inputs / "matrix.txt")
    yuv = synth_video(64, 48, 5)
    enc = oracle.video_encode(yuv, 64, 48, q, True, 3, 16, False)
    if not np.array_equal(ie.decode_video(enc, True)[0], oracle.video_decode(enc, True)[0]):
        print(f"decode variant {variant}: decoded video differs")
        bad += 1
    _lib.check(L.ie_set_option(b"decode_variant", 1))
    print(f"decode variant {variant}: {'ok' if not bad else f'{bad} mismatches'}")
    return 1 if bad else 0


def main_me(variant: int) -> int:
    """ie_set_option("me_variant", V): every motion vector is a field of the P-frame stream, so stream equality covers them"""
    import imageencoder_b200 as ie
    import oracle
    from imageencoder_b200 import _lib
    from imageencoder_b200.synth import synth_video

    L = ie.lib()
    _lib.check(L.ie_init(0))
    q = ie.read_matrix(ROOT / "tests" / "golden" / "inputs" / "matrix.txt")
    bad = 0
    _lib.check(L.ie_set_option(b"me_variant", variant))
    for W, H, F, gop, mer in ((64, 48, 7, 4, 16), (128, 96, 9, 3, 8), (176, 144, 6, 6, 32), (64, 48, 6, 7, 2), (96, 64, 6, 4, 10),
                              (64, 48, 5, 4, 1)):
        yuv = synth_video(W, H, F, 4000)
        if ie.encode_video(yuv, W, H, q, True, gop, mer, False) != oracle.video_encode(yuv, W, H, q, True, gop, mer, False):
            print(f"me variant {variant}: {W}x{H}x{F} gop {gop} merange {mer}: stream differs")
            bad += 1
    _lib.check(L.ie_set_option(b"me_variant", 0))
    print(f"me variant {variant}: {'ok' if not bad else f'{bad} mismatches'}")
    return 1 if bad else 0


def corrupt_cases(count: int = 300, seed: int = 777):
    """(N, W, H, mode, damaged stream) -- mode 0 truncated, 1 bit flips in the body, 2 trailing garbage"""
    import oracle
    rng = np.random.default_rng(seed)
    for _ in range(count):
        N = int(rng.choice([4, 8]))
        W, H = N * int(rng.integers(1, 24)), N * int(rng.integers(1, 24))
        img = np.clip(rng.normal(128, rng.uniform(1, 80), (H, W)), 0, 255).astype(np.uint8)
        q = rng.integers(1, 64, (N, N)).astype(np.uint16)
        enc = bytearray(oracle.image_encode(img, W, H, N, q, bool(rng.integers(0, 2)), False))
        hdr = (1 + 5 + N * N * 8 + 1 + 30 + 7) // 8 + 1
        mode = int(rng.integers(0, 3))
        if mode == 0 and len(enc) > hdr + 1:
            enc = enc[: int(rng.integers(hdr, len(enc)))]
        elif mode == 1 and len(enc) > hdr + 1:
            for _k in range(int(rng.integers(1, 6))):
                enc[int(rng.integers(hdr, len(enc)))] ^= 1 << int(rng.integers(0, 8))
        else:
            enc = enc + bytes(rng.integers(0, 256, int(rng.integers(1, 40))).astype(np.uint8))
        yield N, W, H, mode, bytes(enc)


def main_corrupt() -> int:
    """Damaged image streams (bit flips in the body, truncation, trailing garbage; oracle/fuzz_vs_ref.py's generator with a fixed
    seed) through the default decode path.  One contract (against Block.cpp:441-472):
      * no block on the true chain carries a length field > N*N  ->  same pixels as the oracle (= the compiled reference,
        profiles/r1_oracle_fuzz.md), reads past the end giving zero bits (BitStream.cpp:17-20);
      * some block's length field exceeds N*N, where the reference indexes its zigzag table out of bounds (undefined
        behaviour, Block.cpp:460-465)  ->  IE_EFORMAT, nothing else."""
    import imageencoder_b200 as ie
    import oracle
    from imageencoder_b200 import _lib
    from oracle.fuzz_vs_ref import _has_overlong_block

    _lib.check(ie.lib().ie_init(0))
    only = int(sys.argv[2]) if len(sys.argv) > 2 else 0          # replay one case (1-based), e.g. under compute-sanitizer
    bad = n = n_over = 0
    for n, (N, W, H, mode, enc) in enumerate(corrupt_cases(), 1):
        if only and n != only:
            continue
        overlong = _has_overlong_block(enc, N)
        n_over += overlong
        try:
            got = ie.decode_image(enc, N)
        except ie.IEError as e:
            if not (overlong and e.code == _lib.IE_EFORMAT):
                print(f"corrupt case {n} (mode {mode}, {N}x{N}, {W}x{H}, overlong={overlong}): decoder raised {e}", flush=True)
                bad += 1
                if e.code == _lib.IE_ECUDA:
                    break                      # the CUDA context is gone: every later case would fail the same way
            continue
        if overlong:
            print(f"corrupt case {n} (mode {mode}, {N}x{N}, {W}x{H}): length field > N*N accepted, IE_EFORMAT expected")
            bad += 1
            continue
        want = oracle.image_decode(enc, N)[0]
        if got.shape != want.shape or not np.array_equal(got, want):
            print(f"corrupt case {n} (mode {mode}, {N}x{N}, {W}x{H}): pixels differ from the oracle's")
            bad += 1
    print(f"corrupt streams: {n} cases, {n_over} with a length field > N*N (IE_EFORMAT), "
          f"{'ok' if not bad else f'{bad} wrong'}")
    return 1 if bad else 0


if __name__ == "__main__":
    a = sys.argv[1]
    if a == "corrupt":
        sys.exit(main_corrupt())
    sys.exit(main_decode(int(a[3:])) if a.startswith("dec") else main_me(int(a[2:])) if a.startswith("me") else main(int(a)))
