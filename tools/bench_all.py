#!/usr/bin/env python
"""Times the other rows of SURVEY 8 on one B200 (device-resident where an entry point exists, host API otherwise).
Not the driver's benchmark (that is bench.py); prints one JSON object for DESIGN.md / profiles/."""
import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import imageencoder_b200 as ie
from imageencoder_b200 import _lib, device
from imageencoder_b200.synth import synth_image, synth_video

INP = ROOT / "tests" / "golden" / "inputs"
_lib.check(ie.lib().ie_init(0))
L = ie.lib()
out = {}


def ev_time(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def wall(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); fn(); ts.append((time.perf_counter() - t) * 1e3)
    return float(np.median(ts))


size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
# ---- image encode / decode, 8x8, config 2 shape ------------------------------------------------------------
q8 = ie.read_matrix(INP / "matrix8_1.txt")
img = synth_image(size, size, 1234)
d_raw = torch.from_numpy(img).cuda().reshape(-1)
cap = int(L.ie_max_encoded_bytes(size, size, 8, 1))
d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
d_bits = torch.zeros(1, dtype=torch.int64, device="cuda")
s_enc = device.Session(0, size, size, 8)
ms = ev_time(lambda: device.encode_image_dev(s_enc, d_raw, q8, True, d_out, d_bits))
nbytes = (int(d_bits.item()) + 7) // 8
out["encode8_ms"] = ms; out["encode8_gpx_s"] = size * size / ms / 1e6; out["encoded_bytes"] = nbytes
s_dec = device.Session(1, size, size, 8)
d_dec = torch.empty(size * size, dtype=torch.uint8, device="cuda")
ms = ev_time(lambda: device.decode_image_dev(s_dec, d_out, nbytes, d_dec, 1), reps=3, warm=1)
out["decode8_ms"] = ms; out["decode8_gpx_s"] = size * size / ms / 1e6
# ---- Huffman stage on that stream ----------------------------------------------------------------------------
d_h = torch.empty(nbytes + 8192, dtype=torch.uint8, device="cuda")
hb = [0]
def huff():
    hb[0] = device.huffman_encode_dev(s_enc, d_out, nbytes, d_h)
ms = wall(huff)
out["huffman_encode_ms_wall"] = ms; out["huffman_bytes"] = hb[0]; out["huffman_GBps_in"] = nbytes / ms / 1e6
# ---- 4x4 batch, config 4 shape (8 images) --------------------------------------------------------------------
q4 = ie.read_matrix(INP / "matrix4_2.txt")
b = 4
imgs = np.stack([synth_image(4096, 4096, 2000 + i) for i in range(b)])
slot = int(L.ie_max_encoded_bytes(4096, 4096, 4, 1))
outb = np.zeros((b, slot), np.uint8); sizes = (C.c_size_t * b)()
qq = np.ascontiguousarray(q4, np.uint16).reshape(-1)
ms = wall(lambda: _lib.check(L.ie_encode_images(C.c_void_p(imgs.ctypes.data), b, 4096, 4096, 4, qq.ctypes.data_as(C.POINTER(C.c_uint16)), 1, 0,
                                                C.c_void_p(outb.ctypes.data), slot, sizes)))
out["batch4_encode_host_ms_per_image"] = ms / b; out["batch4_encode_host_gpx_s"] = b * 4096 * 4096 / ms / 1e6
# the same batch device-resident, one launch of each kernel for all images
cnt = 16
d_batch = torch.from_numpy(np.concatenate([imgs] * (cnt // b))).cuda().reshape(-1)
slot16 = (slot + 15) // 16 * 16
d_bout = torch.empty(cnt * slot16, dtype=torch.uint8, device="cuda")
d_bbits = torch.zeros(cnt, dtype=torch.int64, device="cuda")
sb = device.Session(0, 4096, 4096, 4)
ms = ev_time(lambda: device.encode_images_dev(sb, d_batch, cnt, q4, True, d_bout, slot16, d_bbits))
out["batch16_encode4_dev_ms_per_image"] = ms / cnt; out["batch16_encode4_dev_gpx_s"] = cnt * 4096 * 4096 / ms / 1e6
bsz = [(int(x) + 7) // 8 for x in d_bbits.cpu().tolist()]
d_bdec = torch.empty(cnt * 4096 * 4096, dtype=torch.uint8, device="cuda")
sbd = device.Session(1, 4096, 4096, 4)
ms = ev_time(lambda: device.decode_images_dev(sbd, d_bout, slot16, bsz, d_bdec, 4096 * 4096), reps=3, warm=1)
out["batch16_decode4_dev_ms_per_image"] = ms / cnt; out["batch16_decode4_dev_gpx_s"] = cnt * 4096 * 4096 / ms / 1e6
del d_batch, d_bout, d_bdec
d_raw4 = torch.from_numpy(imgs[0]).cuda().reshape(-1)
d_out4 = torch.empty(slot, dtype=torch.uint8, device="cuda")
s4 = device.Session(0, 4096, 4096, 4)
ms = ev_time(lambda: device.encode_image_dev(s4, d_raw4, q4, True, d_out4, d_bits))
out["encode4_ms"] = ms; out["encode4_gpx_s"] = 4096 * 4096 / ms / 1e6
nb4 = (int(d_bits.item()) + 7) // 8
s4d = device.Session(1, 4096, 4096, 4)
d_dec4 = torch.empty(4096 * 4096, dtype=torch.uint8, device="cuda")
ms = ev_time(lambda: device.decode_image_dev(s4d, d_out4, nb4, d_dec4, 1), reps=3, warm=1)
out["decode4_ms"] = ms; out["decode4_gpx_s"] = 4096 * 4096 / ms / 1e6
# ---- video, config 5 shape (24 frames) -------------------------------------------------------------------------
W, H, F = 1920, 1088, 24
yuv = synth_video(W, H, F, 4000)
qv = ie.read_matrix(INP / "matrix.txt")
enc = [None]
def venc():
    enc[0] = ie.encode_video(yuv, W, H, qv, True, 12, 16, False)
ms = wall(venc, reps=2, warm=1)
out["video_encode_host_ms_per_frame"] = ms / F; out["video_encode_host_gpx_s"] = W * H * F / ms / 1e6; out["video_bytes"] = len(enc[0])
ms = wall(lambda: ie.decode_video(enc[0], True), reps=2, warm=1)
out["video_decode_host_ms_per_frame"] = ms / F; out["video_decode_host_gpx_s"] = W * H * F / ms / 1e6
# device-resident video (frames and stream stay in HBM)
d_yuv0 = torch.from_numpy(np.ascontiguousarray(yuv)).cuda().reshape(-1)
d_yuv = d_yuv0.clone()
d_vout = torch.empty(int(L.ie_max_encoded_bytes(W, H, 4, F)) + 4096, dtype=torch.uint8, device="cuda")
sv = device.Session(2, W, H, 4, F)
def venc_dev():
    d_yuv.copy_(d_yuv0)                                   # the encoder rebuilds the frames in place
    device.encode_video_dev(sv, d_yuv, W, H, qv, True, 12, 16, d_vout, d_bits, lead_bit=True)
copy_ms = ev_time(lambda: d_yuv.copy_(d_yuv0))
ms = ev_time(venc_dev, reps=3, warm=1) - copy_ms
vb = (int(d_bits.item()) + 7) // 8
out["video_encode_dev_ms_per_frame"] = ms / F; out["video_encode_dev_gpx_s"] = W * H * F / ms / 1e6
svd = device.Session(3, W, H, 4, F)
d_vdec = torch.empty(W * H * 3 // 2 * F, dtype=torch.uint8, device="cuda")
ms = wall(lambda: (device.decode_video_dev(svd, d_vout, vb, d_vdec, True), torch.cuda.synchronize()), reps=3, warm=1)
out["video_decode_dev_ms_per_frame"] = ms / F; out["video_decode_dev_gpx_s"] = W * H * F / ms / 1e6
# ---- config 5 shape: 240 frames, GOP 12 -> 20 GOPs per launch ------------------------------------------------
del d_yuv, d_yuv0, d_vout, d_vdec
F5 = 240
yuv5 = synth_video(W, H, F5, 4000)
d_yuv0 = torch.from_numpy(np.ascontiguousarray(yuv5)).cuda().reshape(-1)
d_yuv = d_yuv0.clone()
d_vout = torch.empty(int(L.ie_max_encoded_bytes(W, H, 4, F5)) + 4096, dtype=torch.uint8, device="cuda")
sv5 = device.Session(2, W, H, 4, F5)
def venc5():
    d_yuv.copy_(d_yuv0)
    device.encode_video_dev(sv5, d_yuv, W, H, qv, True, 12, 16, d_vout, d_bits, lead_bit=True)
copy_ms = ev_time(lambda: d_yuv.copy_(d_yuv0))
ms = ev_time(venc5, reps=3, warm=1) - copy_ms
out["video240_encode_dev_ms_per_frame"] = ms / F5; out["video240_encode_dev_gpx_s"] = W * H * F5 / ms / 1e6
out["video240_bytes"] = (int(d_bits.item()) + 7) // 8
print(json.dumps(out, indent=1))
