#!/usr/bin/env python
"""Generates tests/golden/golden_configs.json: the BASELINE.json configurations at their FULL sizes, run through the
UNMODIFIED reference (oracle/_ref, built by oracle/build_ref.sh from /root/reference; ImageEncoder.cpp:52-175,
ImageDecoder.cpp:55-122, VideoEncoder.cpp:22-111, VideoDecoder.cpp:33-62) on the seeded synthetic inputs of
imageencoder_b200/synth.py.  Build container only (minutes of CPU); the GPU box reads the committed JSON.

    python tests/golden/make_golden_configs.py [c2] [c3] [c4] [c5]        (default: all)

Every entry also records whether the C restatement (oracle/oracle_block.c, oracle_huffman.cpp) produced the same
bytes at that size ("oracle_equal"), so the restatement is pinned at config size too, not only on the small cases of
golden.json."""
import hashlib
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402
from imageencoder_b200.synth import synth_image, synth_video  # noqa: E402

INPUTS = ROOT / "tests" / "golden" / "inputs"
OUT = ROOT / "tests" / "golden" / "golden_configs.json"
sha = lambda b: hashlib.sha256(bytes(b)).hexdigest()


def image_config(name, W, H, seed, mat, huffs, out):
    q = oracle.read_matrix(INPUTS / mat)
    N = q.shape[0]
    img = synth_image(W, H, seed)
    e = {"W": W, "H": H, "block": N, "matrix": mat, "seed": seed, "rle": True, "input_sha256": sha(img)}
    plain_nolead, nbits = oracle.image_encode_plain(img, W, H, N, q, True, lead_bit=False)
    for huff in huffs:
        t0 = time.time()
        enc, res = oracle.ref_image_encode(img, W, H, N, q, True, huff)
        t_enc = time.time() - t0
        k = "huff" if huff else "plain"
        e[k] = {"enc_bytes": len(enc), "enc_sha256": sha(enc), "ref_process_ms": res["process_ms"][-1], "ref_threads": res["threads"]}
        mine = oracle.image_encode(img, W, H, N, q, True, huff)
        e[k]["oracle_equal"] = (mine == enc)
        if huff:
            # Huffman.cpp:39-42: a code length >= 16 or a same-length group >= 128 symbols overflows the dictionary
            # header; the reference then cannot decode its own output -- say which regime this input is in
            e[k]["dictionary_overflows"] = bool(oracle.huffman_header_overflows(plain_nolead))
            e[k]["reverted"] = bool((enc[0] & 0x80) == 0)
        if huff and e[k]["dictionary_overflows"]:
            e[k]["dec_sha256"] = None
        else:
            dec, dres = oracle.ref_image_decode(enc, N, W, H)
            e[k]["dec_sha256"] = sha(dec.tobytes())
            e[k]["ref_decode_ms"] = dres["process_ms"][-1]
            e[k]["oracle_dec_equal"] = bool(np.array_equal(oracle.image_decode(enc, N)[0].reshape(-1), np.asarray(dec).reshape(-1)))
        print(name, k, e[k], f"({t_enc:.0f} s)", flush=True)
    out[name] = e


def main():
    which = set(sys.argv[1:]) or {"c2", "c3", "c4", "c5"}
    out = json.loads(OUT.read_text()) if OUT.exists() else {}
    out["how"] = ("reference compiled with g++ -std=c++17 -O3 -mlzcnt -fopenmp (oracle/build_ref.sh), run by "
                  "tests/golden/make_golden_configs.py on imageencoder_b200.synth inputs")
    if "c2" in which:
        image_config("C2|8192x8192|matrix8_1|seed1234", 8192, 8192, 1234, "matrix8_1.txt", (False,), out)
        OUT.write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")
    if "c4" in which:
        for seed in (2000, 2001, 2002):
            image_config(f"C4|4096x4096|matrix4_2|seed{seed}", 4096, 4096, seed, "matrix4_2.txt", (False,), out)
        OUT.write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")
    if "c5" in which:
        W, H, F, gop, mer = 1920, 1088, 240, 12, 16
        q = oracle.read_matrix(INPUTS / "matrix.txt")
        yuv = synth_video(W, H, F, 4000)
        t0 = time.time()
        enc, res = oracle.ref_video_encode(yuv, W, H, q, True, gop, mer, False)
        e = {"W": W, "H": H, "frames": F, "gop": gop, "merange": mer, "matrix": "matrix.txt", "seed": 4000, "rle": True,
             "input_sha256": sha(yuv), "enc_bytes": len(enc), "enc_sha256": sha(enc),
             "ref_process_ms": res["process_ms"][-1], "ref_threads": res["threads"]}
        print("C5 enc", e, f"({time.time() - t0:.0f} s)", flush=True)
        for mc in (True, False):
            dec, dres = oracle.ref_video_decode(enc, mc)
            e[f"dec_mc{int(mc)}_sha256"] = sha(np.asarray(dec).tobytes())
            e[f"ref_decode_mc{int(mc)}_ms"] = dres["process_ms"][-1]
        e["oracle_equal"] = (oracle.video_encode(yuv, W, H, q, True, gop, mer, False) == enc)
        print("C5", e, flush=True)
        out["C5|1920x1088x240|gop12|mer16|matrix|seed4000"] = e
        OUT.write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")
    if "c3" in which:
        image_config("C3|16384x16384|matrix8_2|seed1235", 16384, 16384, 1235, "matrix8_2.txt", (False, True), out)
        OUT.write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")
    print("entries:", sorted(k for k in out if k != "how"))


if __name__ == "__main__":
    main()
