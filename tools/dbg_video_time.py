"""Device-resident video encode of F frames (GOP 12) for an ncu launch list: python tools/dbg_video_time.py [F]"""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
import imageencoder_b200 as ie
from imageencoder_b200 import _lib, device
from imageencoder_b200.synth import synth_video
_lib.check(ie.lib().ie_init(0))
F = int(sys.argv[1]) if len(sys.argv) > 1 else 240
W, H = 1920, 1088
base = synth_video(W, H, 24, 4000)
yuv = np.tile(base, (F + 23) // 24)[: F * W * H * 3 // 2]
qv = ie.read_matrix('tests/golden/inputs/matrix.txt')
d0 = torch.from_numpy(yuv).cuda(); d = d0.clone()
d_out = torch.empty(int(ie.lib().ie_max_encoded_bytes(W, H, 4, F)) + 4096, dtype=torch.uint8, device="cuda")
d_bits = torch.zeros(1, dtype=torch.int64, device="cuda")
s = device.Session(2, W, H, 4, F)
for _ in range(2):
    d.copy_(d0)
    device.encode_video_dev(s, d, W, H, qv, True, 12, 16, d_out, d_bits)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
d.copy_(d0); a.record()
device.encode_video_dev(s, d, W, H, qv, True, 12, 16, d_out, d_bits)
b.record(); torch.cuda.synchronize()
print("frames", F, "ms", a.elapsed_time(b), "ms/frame", a.elapsed_time(b) / F, "bytes", (int(d_bits.item()) + 7) // 8)
