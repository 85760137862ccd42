#!/usr/bin/env python
"""Generates tests/golden/golden.json by running the UNMODIFIED reference (oracle/_ref, built by oracle/build_ref.sh from
/root/reference) on its own sample inputs and on seeded synthetic inputs.  Run in the build container only; the GPU box
reads the committed JSON.  Usage: python tests/golden/make_golden.py
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402
from imageencoder_b200.synth import synth_image, synth_video  # noqa: E402

REF_BIN = Path("/root/reference/bin")
INPUTS = ROOT / "tests" / "golden" / "inputs"
sha = lambda b: hashlib.sha256(bytes(b)).hexdigest()


def image_entry(raw, W, H, N, q, rle, huff):
    enc, _ = oracle.ref_image_encode(raw, W, H, N, q, rle, huff)
    if huff and oracle.huffman_header_overflows(oracle.image_encode_plain(raw, W, H, N, q, rle, lead_bit=False)[0]):
        # Huffman code lengths >= 16 (or a same-length group >= 128) overflow the dictionary header fields
        # (Huffman.cpp:39-42): the reference cannot decode its own output; encoder parity is still defined.
        dsha = None
    else:
        dec, _ = oracle.ref_image_decode(enc, N, W, H)
        dsha = sha(dec.tobytes())
    return {"W": W, "H": H, "block": N, "enc_bytes": len(enc), "enc_sha256": sha(enc), "dec_sha256": dsha}


def _plain_video(yuv, W, H, q, gop, mer):
    """the byte-rounded stream the Huffman stage sees (built without the leading bit, VideoEncoder.cpp:60-62)"""
    import ctypes as C
    buf = np.array(yuv, dtype=np.uint8).copy()
    qq = np.ascontiguousarray(q, dtype=np.uint16).reshape(-1)
    cap = buf.size * 3 + 1024
    o = np.zeros(cap, dtype=np.uint8)
    bits = oracle.lib().orc_video_encode(buf.ctypes.data_as(C.POINTER(C.c_uint8)), buf.size, W, H,
                                         qq.ctypes.data_as(C.POINTER(C.c_uint16)), 1, gop, mer, 0,
                                         o.ctypes.data_as(C.POINTER(C.c_uint8)), cap, None)
    return o[: (bits + 7) // 8].tobytes()


def main():
    out = {"how": "reference compiled with g++ -std=c++17 -O3 -mlzcnt -fopenmp (oracle/build_ref.sh)", "images": {}, "video": {}}
    samples = {"ex0": (8, 8), "ex1": (936, 936), "ex2": (512, 512), "ex3": (400, 400), "ex4": (4096, 912), "ex6": (512, 256)}
    mats = {m: oracle.read_matrix(INPUTS / m) for m in ("matrix.txt", "matrix4_2.txt", "matrix8_1.txt", "matrix8_2.txt")}
    for name, (W, H) in samples.items():
        raw = np.fromfile(REF_BIN / f"{name}.raw", dtype=np.uint8)
        out["images"][f"{name}|input"] = {"sha256": sha(raw), "bytes": int(raw.size)}
        for huff in (False, True):
            out["images"][f"{name}|matrix.txt|rle1|{'huff' if huff else 'plain'}"] = image_entry(raw, W, H, 4, mats["matrix.txt"], True, huff)
    raw = np.fromfile(REF_BIN / "ex2.raw", dtype=np.uint8)
    for m in ("matrix8_1.txt", "matrix8_2.txt"):
        for huff in (False, True):
            out["images"][f"ex2|{m}|rle1|{'huff' if huff else 'plain'}"] = image_entry(raw, 512, 512, 8, mats[m], True, huff)
    for seed, W, H, flat in ((1234, 1024, 1024, False), (1235, 512, 768, True), (2000, 256, 256, False)):
        img = synth_image(W, H, seed, flat=flat)
        out["images"][f"synth{seed}|input"] = {"sha256": sha(img), "W": W, "H": H, "flat": flat}
        for m in mats:
            N = mats[m].shape[0]
            for rle in (True, False):
                for huff in (False, True):
                    out["images"][f"synth{seed}|{m}|rle{int(rle)}|{'huff' if huff else 'plain'}"] = image_entry(img, W, H, N, mats[m], rle, huff)
    for (W, H, F, gop, mer) in ((64, 48, 7, 4, 16), (128, 96, 9, 3, 8), (176, 144, 6, 6, 32)):
        yuv = synth_video(W, H, F, 4000)
        for huff in (False, True):
            enc, _ = oracle.ref_video_encode(yuv, W, H, mats["matrix.txt"], True, gop, mer, huff)
            if huff and oracle.huffman_header_overflows(_plain_video(yuv, W, H, mats["matrix.txt"], gop, mer)):
                out["video"][f"synth4000|{W}x{H}x{F}|gop{gop}|mer{mer}|huff"] = {
                    "input_sha256": sha(yuv), "enc_bytes": len(enc), "enc_sha256": sha(enc), "dec_mc1_sha256": None, "dec_mc0_sha256": None}
                continue
            d1, _ = oracle.ref_video_decode(enc, True)
            d0, _ = oracle.ref_video_decode(enc, False)
            out["video"][f"synth4000|{W}x{H}x{F}|gop{gop}|mer{mer}|{'huff' if huff else 'plain'}"] = {
                "input_sha256": sha(yuv), "enc_bytes": len(enc), "enc_sha256": sha(enc), "dec_mc1_sha256": sha(d1), "dec_mc0_sha256": sha(d0)}
    (ROOT / "tests" / "golden" / "golden.json").write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")
    print("entries:", len(out["images"]), len(out["video"]))


if __name__ == "__main__":
    main()
