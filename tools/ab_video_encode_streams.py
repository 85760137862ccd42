"""A/B of ie_set_option("video_encode_streams", 1 | 2) on config 5: ms per encode, streams compared"""
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
h_yuv = torch.from_numpy(np.ascontiguousarray(yuv)).reshape(-1)
d_yuv = h_yuv.cuda()
d_out = torch.empty(int(ie.lib().ie_max_encoded_bytes(W, H, 4, F)) + 4096, dtype=torch.uint8, device="cuda")
d_bits = torch.zeros(1, dtype=torch.int64, device="cuda")
sv = device.Session(2, W, H, 4, F)
ref = None
for ns in (1, 2, 1, 2):
    _lib.check(ie.lib().ie_set_option(b"video_encode_streams", ns))
    ts = []
    for rep in range(4):
        d_yuv.copy_(h_yuv)                      # the encoder rebuilds the P-frames in place
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        device.encode_video_dev(sv, d_yuv, W, H, q, True, 12, 16, d_out, d_bits, lead_bit=True)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    nb = (int(d_bits.item()) + 7) // 8
    cur = d_out[:nb].clone()
    if ref is None:
        ref = cur
    print(f"video_encode_streams {ns}: {min(ts[1:]):.3f} ms per encode (best of 3), {nb} bytes, identical: {bool(torch.equal(ref, cur))}", flush=True)
_lib.check(ie.lib().ie_set_option(b"video_encode_streams", 2))
