"""A/B of the number of GOP batches of the whole-stream video decode (ie_set_option("video_decode_batches", n)) on config 5"""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import imageencoder_b200 as ie
from imageencoder_b200 import _lib, device
from imageencoder_b200.synth import synth_video
_lib.check(ie.lib().ie_init(0))
W, H, F = 1920, 1088, 240
q = ie.read_matrix('tests/golden/inputs/matrix.txt')
yuv = synth_video(W, H, F, 4000)
d_yuv = torch.from_numpy(np.ascontiguousarray(yuv)).cuda().reshape(-1)
d_out = torch.empty(int(ie.lib().ie_max_encoded_bytes(W, H, 4, F)) + 4096, dtype=torch.uint8, device="cuda")
d_bits = torch.zeros(1, dtype=torch.int64, device="cuda")
sv = device.Session(2, W, H, 4, F)
device.encode_video_dev(sv, d_yuv, W, H, q, True, 12, 16, d_out, d_bits, lead_bit=True)
torch.cuda.synchronize()
nb = (int(d_bits.item()) + 7) // 8
sd = device.Session(3, W, H, 4, F)
d_dec = torch.empty(W * H * 3 // 2 * F, dtype=torch.uint8, device="cuda")
ref = None
for nbat in (1, 2, 3, 4, 5, 6, 8, 10, 20, 5):
    _lib.check(ie.lib().ie_set_option(b"video_decode_batches", nbat))
    for _ in range(2):
        device.decode_video_dev(sd, d_out, nb, d_dec, True)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(5):
        device.decode_video_dev(sd, d_out, nb, d_dec, True)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t) / 5 * 1e3
    if ref is None:
        ref = d_dec.clone()
    print(f"batches {nbat:2d}: {ms:.3f} ms per decode, frames identical: {bool(torch.equal(ref, d_dec))}", flush=True)
_lib.check(ie.lib().ie_set_option(b"video_decode_batches", 4))
