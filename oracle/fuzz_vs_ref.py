#!/usr/bin/env python
"""Randomised comparison of the oracle (the restatement in this directory) with the compiled, unmodified reference
(oracle/_ref, built by build_ref.sh where /root/reference exists).  Test infrastructure, not product code.

    python oracle/fuzz_vs_ref.py image SECONDS [SEED]      random sizes / contents / quantisers / RLE / Huffman, both block sizes
    python oracle/fuzz_vs_ref.py video SECONDS [SEED]      random clips, gop 1..8, merange in {1..64, not only powers of two}
    python oracle/fuzz_vs_ref.py corrupt SECONDS [SEED]    decode of truncated / bit-flipped / padded image streams
    python oracle/fuzz_vs_ref.py trunc SECONDS [SEED]      decode of truncated / padded Huffman-coded image streams and of
                                                           truncated video streams (with and without motion compensation)

A case counts as a REAL mismatch only where the reference's behaviour is defined.  Two regimes are the reference's own
undefined behaviour and are reported separately ("avoid"):
  * Huffman revert path: 8*original_length + 1 bits written into original_length bytes (Huffman.cpp:332-338) -- the last
    byte of the file is heap memory (pad bits always, the data bit too when original_length % 16 == 8);
  * video streams longer than the raw Y planes: the output buffer is sized frames * W * H * 8 bits (VideoEncoder.cpp:35-53,
    Frame.cpp:23-29) and put_bit has no bounds check (BitStream.cpp:61-71) -- heap overflow, anything from a garbage last byte
    to corrupted later frames to a glibc abort.
  * (corrupt mode) a block whose length field exceeds N*N: the reference indexes its zigzag table out of bounds
    (Block.cpp:460-465) -- undefined behaviour.  The oracle reads and discards the surplus values (so that it can keep
    walking); the product rejects such a stream with IE_EFORMAT (tests/_variant_worker.py corrupt).
Prints one summary line; exit code 1 if there is a REAL mismatch."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import oracle  # noqa: E402


def image_case(rng):
    N = int(rng.choice([4, 8]))
    W = N * int(rng.integers(1, 20 if rng.integers(0, 4) else 64))
    H = N * int(rng.integers(1, 20 if rng.integers(0, 4) else 64))
    kind = int(rng.integers(0, 6))
    if kind == 0:
        img = rng.integers(0, 256, (H, W))
    elif kind == 1:
        img = rng.integers(0, 2, (H, W)) * int(rng.integers(1, 128)) + int(rng.integers(0, 128))      # two levels: ties
    elif kind == 2:
        img = np.full((H, W), int(rng.integers(0, 256)))
    elif kind == 3:
        yy, xx = np.mgrid[0:H, 0:W]
        img = (128 + 100 * np.sin(xx / rng.uniform(2, 30)) * np.cos(yy / rng.uniform(2, 30))).astype(int) + rng.integers(-3, 4, (H, W))
    elif kind == 4:
        img = rng.integers(0, 2, (H, W)) * 255
    else:
        img = np.clip(rng.normal(128, rng.uniform(1, 80), (H, W)), 0, 255)
    img = np.clip(img, 0, 255).astype(np.uint8)
    qk = int(rng.integers(0, 5))
    if qk == 0:
        q = rng.integers(1, 256, (N, N))
    elif qk == 1:
        q = np.ones((N, N), int)
    elif qk == 2:
        q = rng.integers(1, 8, (N, N))
    elif qk == 3:
        q = rng.integers(1, 65536, (N, N))
    else:
        q = 1 + np.add.outer(np.arange(N), np.arange(N)) * int(rng.integers(1, 20))
    return N, W, H, img, q.astype(np.uint16), bool(rng.integers(0, 2)), bool(rng.integers(0, 2))


def run_image(seconds, seed):
    rng = np.random.default_rng(seed)
    t0 = time.time()
    n = real = ub = crashed = 0
    while time.time() - t0 < seconds:
        N, W, H, img, q, rle, huff = image_case(rng)
        try:
            ref, _ = oracle.ref_image_encode(img, W, H, N, q, rle, huff, threads=1)
        except RuntimeError:
            crashed += 1                    # glibc abort after the revert path's heap overflow
            continue
        mine = oracle.image_encode(img, W, H, N, q, rle, huff)
        n += 1
        if ref != mine:
            reverted = huff and len(ref) == len(mine) and not (ref[0] & 0x80)
            if reverted and ref[:-1] == mine[:-1] and ((ref[-1] ^ mine[-1]) & 0x80 == 0 or (len(ref) - 1) % 16 == 8):
                ub += 1
            else:
                real += 1
                print("REAL image mismatch", dict(seed=seed, n=n, N=N, W=W, H=H, rle=rle, huff=huff), flush=True)
            continue
        if not huff:
            rdec, _ = oracle.ref_image_decode(ref, N, W, H, threads=1)
            if not np.array_equal(oracle.image_decode(ref, N)[0], rdec):
                real += 1
                print("REAL image decode mismatch", dict(seed=seed, n=n, N=N, W=W, H=H, rle=rle), flush=True)
    print(f"image fuzz seed {seed}: {n} cases, {real} REAL mismatches, {ub} differ only in the revert path's heap byte, "
          f"{crashed} reference aborts")
    return real


def video_case(rng):
    W = 16 * int(rng.integers(1, 7))
    H = 16 * int(rng.integers(1, 6))
    F = int(rng.integers(1, 8))
    fsz = W * H * 3 // 2
    kind = int(rng.integers(0, 5))
    big = rng.integers(0, 256, (H + 64, W + 64)).astype(np.uint8)
    if kind in (1, 3):
        yy, xx = np.mgrid[0:H + 64, 0:W + 64]
        big = np.clip(128 + 90 * np.sin(xx / rng.uniform(2, 12)) * np.cos(yy / rng.uniform(2, 12)) + rng.normal(0, 3, big.shape), 0, 255).astype(np.uint8)
    yuv = np.full((F, fsz), 0x80, np.uint8)
    x0, y0 = int(rng.integers(0, 32)), int(rng.integers(0, 32))
    for t in range(F):
        if kind == 2:
            fr = rng.integers(0, 256, (H, W))
        elif kind == 4:
            fr = np.full((H, W), int(rng.integers(0, 256)))
        else:                                  # a window moving over a larger picture (+ noise): real motion vectors
            x0 = int(np.clip(x0 + int(rng.integers(-6, 7)), 0, 63))
            y0 = int(np.clip(y0 + int(rng.integers(-6, 7)), 0, 63))
            fr = big[y0:y0 + H, x0:x0 + W].astype(int) + (rng.integers(-2, 3, (H, W)) if kind == 3 else 0)
        yuv[t, :W * H] = np.clip(fr, 0, 255).astype(np.uint8).reshape(-1)
    q = (rng.integers(1, 256, (4, 4)) if rng.integers(0, 2) else rng.integers(4, 40, (4, 4))).astype(np.uint16)
    gop = int(rng.integers(1, 9))
    mer = int(rng.choice([1, 2, 3, 4, 5, 7, 8, 12, 16, 31, 32, 64]))
    return W, H, F, yuv.reshape(-1), q, gop, mer, bool(rng.integers(0, 2))


def run_video(seconds, seed):
    rng = np.random.default_rng(seed)
    t0 = time.time()
    n = real = ub = crashed = 0
    while time.time() - t0 < seconds:
        W, H, F, yuv, q, gop, mer, rle = video_case(rng)
        mine = oracle.video_encode(yuv, W, H, q, rle, gop, mer, False)
        overflow = len(mine) * 8 > 211 + F * W * H * 8          # longer than the reference's output buffer: heap overflow there
        try:
            ref, _ = oracle.ref_video_encode(yuv, W, H, q, rle, gop, mer, False, threads=1)
        except RuntimeError:
            crashed += 1
            if not overflow:
                real += 1
                print("REAL: reference failed on a stream inside its buffer", dict(seed=seed, W=W, H=H, F=F, gop=gop, mer=mer), flush=True)
            continue
        n += 1
        if ref != mine:
            if overflow:
                ub += 1
            else:
                real += 1
                print("REAL video mismatch", dict(seed=seed, n=n, W=W, H=H, F=F, gop=gop, mer=mer, rle=rle), flush=True)
            continue
        for mc in (True, False):
            rdec, _ = oracle.ref_video_decode(ref, mc, threads=1)
            if not np.array_equal(np.asarray(oracle.video_decode(ref, mc)[0]).reshape(-1), rdec):
                real += 1
                print("REAL video decode mismatch", dict(seed=seed, n=n, W=W, H=H, F=F, gop=gop, mer=mer, rle=rle, mc=mc), flush=True)
    print(f"video fuzz seed {seed}: {n} cases, {real} REAL mismatches, {ub} differ where the stream overflows the reference's "
          f"output buffer, {crashed} reference aborts")
    return real


def _has_overlong_block(enc: bytes, N: int, lead_bit: bool = True) -> bool:
    """plain image stream: does any block carry a length field > N*N (the reference's out-of-bounds case)?"""
    bits = np.unpackbits(np.frombuffer(enc, np.uint8))
    pos = 0

    def get(n):
        nonlocal pos
        v = 0
        for _ in range(n):
            v = (v << 1) | (int(bits[pos]) if pos < len(bits) else 0)
            pos += 1
        return v

    if lead_bit:
        get(1)
    qb = get(5)
    for _ in range(N * N):
        get(qb)
    rle, W, H = get(1), get(15), get(15)
    for _ in range((W // N) * (H // N)):
        bl = get(4)
        ln = get(bl) if rle else N * N
        if ln > N * N:
            return True
        pos += ln * bl
        if pos > len(bits) + 64:
            break
    return False


def run_corrupt(seconds, seed):
    rng = np.random.default_rng(seed)
    t0 = time.time()
    n = real = ub = crashed = 0
    while time.time() - t0 < seconds:
        N = int(rng.choice([4, 8]))
        W, H = N * int(rng.integers(1, 12)), N * int(rng.integers(1, 12))
        img = np.clip(rng.normal(128, rng.uniform(1, 80), (H, W)), 0, 255).astype(np.uint8)
        q = rng.integers(1, 64, (N, N)).astype(np.uint16)
        enc = bytearray(oracle.image_encode(img, W, H, N, q, bool(rng.integers(0, 2)), False))
        hdr = (1 + 5 + N * N * 8 + 1 + 30 + 7) // 8 + 1            # the header stays intact: W and H pick the output size
        mode = int(rng.integers(0, 3))
        if mode == 0 and len(enc) > hdr + 1:
            enc = enc[: int(rng.integers(hdr, len(enc)))]           # truncated: reads past the end return 0 (BitStream.cpp:17-20)
        elif mode == 1 and len(enc) > hdr + 1:
            for _ in range(int(rng.integers(1, 6))):                # bit flips in the body
                enc[int(rng.integers(hdr, len(enc)))] ^= 1 << int(rng.integers(0, 8))
        else:
            enc = enc + bytes(rng.integers(0, 256, int(rng.integers(1, 40))).astype(np.uint8))      # trailing garbage
        enc = bytes(enc)
        try:
            rdec, _ = oracle.ref_image_decode(enc, N, W, H, threads=1)
        except Exception:
            crashed += 1
            continue
        n += 1
        if not np.array_equal(oracle.image_decode(enc, N)[0], rdec):
            if _has_overlong_block(enc, N):
                ub += 1
            else:
                real += 1
                print("REAL corrupt-stream decode mismatch", dict(seed=seed, n=n, N=N, W=W, H=H, mode=mode), flush=True)
    print(f"corrupt fuzz seed {seed}: {n} cases, {real} REAL mismatches, {ub} differ where a length field exceeds N*N, "
          f"{crashed} reference crashes")
    return real


def run_trunc(seconds, seed):
    rng = np.random.default_rng(seed)
    t0 = time.time()
    n = real = ub = crashed = 0
    while time.time() - t0 < seconds:
        if rng.integers(0, 2):
            N = int(rng.choice([4, 8]))
            W, H = N * int(rng.integers(2, 16)), N * int(rng.integers(2, 16))
            img = np.clip(rng.normal(128, rng.uniform(5, 60), (H, W)), 0, 255).astype(np.uint8)
            q = rng.integers(2, 64, (N, N)).astype(np.uint16)
            enc = oracle.image_encode(img, W, H, N, q, True, True)
            if not (enc[0] & 0x80):
                continue                                    # reverted: not a Huffman stream
            if oracle.huffman_header_overflows(oracle.image_encode_plain(img, W, H, N, q, True, lead_bit=False)[0]):
                continue                                    # the reference cannot decode its own output (Huffman.cpp:39-42)
            if rng.integers(0, 2):
                enc2 = enc[: int(rng.integers(len(enc) * 6 // 10, len(enc)))]
            else:
                enc2 = enc + bytes(rng.integers(0, 256, int(rng.integers(1, 40))).astype(np.uint8))
            try:
                rdec, _ = oracle.ref_image_decode(enc2, N, W, H, threads=1)
            except Exception:
                crashed += 1                                # missing tree branch: the reference dereferences nullptr (Huffman.cpp:197-200)
                continue
            n += 1
            try:
                same = np.array_equal(oracle.image_decode(enc2, N)[0], rdec)
            except ValueError:
                same = False
            if not same:
                if _has_overlong_block(oracle.huffman_decode(enc2)[0], N, lead_bit=False):
                    ub += 1                                 # the garbage block at the cut carries a length field > N*N
                else:
                    real += 1
                    print("REAL truncated Huffman decode mismatch", dict(seed=seed, n=n, N=N, W=W, H=H), flush=True)
        else:
            W, H, F = 16 * int(rng.integers(1, 5)), 16 * int(rng.integers(1, 4)), int(rng.integers(2, 7))
            fsz = W * H * 3 // 2
            yuv = np.full((F, fsz), 0x80, np.uint8)
            big = np.clip(rng.normal(128, 50, (H + 32, W + 32)), 0, 255).astype(np.uint8)
            for t in range(F):
                x0, y0 = int(rng.integers(0, 32)), int(rng.integers(0, 32))
                yuv[t, :W * H] = big[y0:y0 + H, x0:x0 + W].reshape(-1)
            q = rng.integers(8, 64, (4, 4)).astype(np.uint16)
            enc = oracle.video_encode(yuv.reshape(-1), W, H, q, True, int(rng.integers(1, 5)), int(rng.choice([4, 8, 16])), False)
            if len(enc) <= 30:
                continue
            enc2 = enc[: int(rng.integers(28, len(enc)))]
            mc = bool(rng.integers(0, 2))
            try:
                rdec, _ = oracle.ref_video_decode(enc2, mc, threads=1)
            except Exception:
                crashed += 1
                continue
            n += 1
            if not np.array_equal(np.asarray(oracle.video_decode(enc2, mc)[0]).reshape(-1), rdec):
                real += 1
                print("REAL truncated video decode mismatch", dict(seed=seed, n=n, W=W, H=H, F=F, mc=mc), flush=True)
    print(f"trunc fuzz seed {seed}: {n} cases, {real} REAL mismatches, {ub} differ where the block at the cut has a length field > N*N, "
          f"{crashed} reference crashes")
    return real


if __name__ == "__main__":
    mode, secs = sys.argv[1], float(sys.argv[2])
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    sys.exit(1 if {"image": run_image, "video": run_video, "corrupt": run_corrupt, "trunc": run_trunc}[mode](secs, seed) else 0)
