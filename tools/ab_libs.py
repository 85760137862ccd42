"""A/B of two builds of the library on one box (tools/build_variant.sh name -> .tmp_libs/name.so):
     python tools/ab_libs.py base new [more builds ...] [what ...]        what: dec2 (config-2 decode), enc2 (config-2 encode), video (config 5)
   Every build runs in its own process (IMAGEENCODER_B200_LIB); results are checked against the reference's sha256
   (tests/golden/golden_configs.json) and timed with CUDA events, device-resident.  Writes gpurun_out/ab_libs.json."""
import hashlib
import json
import os
import subprocess
import sys


def worker(what):
    sys.path.insert(0, '.')
    import numpy as np
    import torch
    import imageencoder_b200 as ie
    from imageencoder_b200 import _lib, device
    from imageencoder_b200.synth import synth_image, synth_video
    _lib.check(ie.lib().ie_init(0))
    gold = json.load(open('tests/golden/golden_configs.json'))
    out = {}

    def timed(fn, reps=20, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(3):
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) / reps)
        return best

    if "dec2" in what or "enc2" in what:
        size = 8192
        g = gold["C2|8192x8192|matrix8_1|seed1234"]
        q = ie.read_matrix('tests/golden/inputs/matrix8_1.txt')
        img = torch.from_numpy(synth_image(size, size, 1234)).cuda().reshape(-1)
        cap = int(ie.lib().ie_max_encoded_bytes(size, size, 8, 1))
        d_out = torch.zeros(cap, dtype=torch.uint8, device="cuda")
        d_bits = torch.zeros(1, dtype=torch.int64, device="cuda")
        se = device.Session(device.Session.IMAGE_ENCODE, size, size, 8)
        device.encode_image_dev(se, img, q, True, d_out, d_bits)
        torch.cuda.synchronize()
        n = (int(d_bits.item()) + 7) // 8
        out["enc2_sha_ok"] = hashlib.sha256(d_out[:n].cpu().numpy().tobytes()).hexdigest() == g["plain"]["enc_sha256"]
        if "enc2" in what:
            out["enc2_ms"] = timed(lambda: device.encode_image_dev(se, img, q, True, d_out, d_bits))
        if "dec2" in what:
            hdr = device.parse_image_header(d_out[:160].cpu().numpy().tobytes(), 8)
            sd = device.Session(device.Session.IMAGE_DECODE, 0, 0, 8)
            full = torch.empty(size * size, dtype=torch.uint8, device="cuda")
            out["dec2_ms"] = timed(lambda: device.decode_image_with_header_dev(sd, hdr, d_out, n, full))
            out["dec2_sha_ok"] = hashlib.sha256(full.cpu().numpy().tobytes()).hexdigest() == g["plain"]["dec_sha256"]
    if "video" in what:
        W, H, F = 1920, 1088, 240
        g = gold['C5|1920x1088x240|gop12|mer16|matrix|seed4000']
        q = ie.read_matrix('tests/golden/inputs/matrix.txt')
        d_yuv = torch.from_numpy(np.ascontiguousarray(synth_video(W, H, F, 4000))).cuda().reshape(-1)
        d_out = torch.empty(int(ie.lib().ie_max_encoded_bytes(W, H, 4, F)) + 4096, dtype=torch.uint8, device="cuda")
        d_bits = torch.zeros(1, dtype=torch.int64, device="cuda")
        sv = device.Session(2, W, H, 4, F)
        keep = d_yuv.clone()

        def enc():
            d_yuv.copy_(keep)          # the encoder rebuilds the P-frames in place (Frame.cpp:218-242)
            device.encode_video_dev(sv, d_yuv, W, H, q, True, 12, 16, d_out, d_bits, lead_bit=True)
        t_copy = timed(lambda: d_yuv.copy_(keep), reps=5)
        out["venc_ms"] = timed(enc, reps=5) - t_copy
        nb = (int(d_bits.item()) + 7) // 8
        out["venc_sha_ok"] = hashlib.sha256(d_out[:nb].cpu().numpy().tobytes()).hexdigest() == g["enc_sha256"]
        sd = device.Session(3, W, H, 4, F)
        d_dec = torch.empty(W * H * 3 // 2 * F, dtype=torch.uint8, device="cuda")
        out["vdec_ms"] = timed(lambda: device.decode_video_dev(sd, d_out, nb, d_dec, True), reps=5)
        out["vdec_sha_ok"] = hashlib.sha256(d_dec.cpu().numpy().tobytes()).hexdigest() == g["dec_mc1_sha256"]
    print("RESULT " + json.dumps(out), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "--worker":
        worker(sys.argv[2:])
    else:
        WHAT = ("dec2", "enc2", "video")
        names = [a for a in sys.argv[1:] if a not in WHAT]
        what = [a for a in sys.argv[1:] if a in WHAT] or ["dec2", "video"]
        res = {}
        for rnd in range(2):                       # base, new, base, new: the box warms up
            for nm in names:
                env = dict(os.environ, IMAGEENCODER_B200_LIB=os.path.abspath(f".tmp_libs/{nm}.so"))
                p = subprocess.run([sys.executable, __file__, "--worker"] + what, env=env, capture_output=True, text=True)
                line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
                r = json.loads(line[-1][7:]) if line else {"error": (p.stderr or p.stdout)[-800:]}
                res.setdefault(nm, []).append(r)
                print(nm, rnd, json.dumps(r), flush=True)
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(res, open("gpurun_out/ab_libs.json", "w"), indent=1)
