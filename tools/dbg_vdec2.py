import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import imageencoder_b200 as ie
from imageencoder_b200 import _lib, device
from imageencoder_b200.synth import synth_video
L = ie.lib(); _lib.check(L.ie_init(0))
W, H = 1920, 1088
q = ie.read_matrix('tests/golden/inputs/matrix.txt')
for F in (24, 48, 96):
    yuv = synth_video(W, H, F, 4000)
    d_yuv = torch.from_numpy(np.ascontiguousarray(yuv)).cuda().reshape(-1)
    d_out = torch.empty(int(L.ie_max_encoded_bytes(W, H, 4, F)) + 4096, dtype=torch.uint8, device="cuda")
    d_bits = torch.zeros(1, dtype=torch.int64, device="cuda")
    sv = device.Session(2, W, H, 4, F)
    device.encode_video_dev(sv, d_yuv, W, H, q, True, 12, 16, d_out, d_bits, lead_bit=True)
    torch.cuda.synchronize()
    nb = (int(d_bits.item()) + 7) // 8
    sd = device.Session(3, W, H, 4, F)
    d_dec = torch.empty(W * H * 3 // 2 * F, dtype=torch.uint8, device="cuda")
    for rep in range(3):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t = time.perf_counter()
        a.record()
        device.decode_video_dev(sd, d_out, nb, d_dec, True)
        t_launch = time.perf_counter() - t
        b.record()
        torch.cuda.synchronize()
        print(f"F={F} rep {rep}: wall {(time.perf_counter() - t) * 1e3:.2f} ms, host launch part {t_launch * 1e3:.2f} ms, device {a.elapsed_time(b):.2f} ms, bytes {nb}", flush=True)
    del sv, sd, d_dec, d_out, d_yuv
