"""workload for ncu: device-resident video encode + decode of a 1920x1088 clip (24 frames, GOP 12)"""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
import imageencoder_b200 as ie
from imageencoder_b200 import _lib, device
from imageencoder_b200.synth import synth_video
_lib.check(ie.lib().ie_init(0))
W, H, F = 1920, 1088, int(sys.argv[1]) if len(sys.argv) > 1 else 24
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
for _ in range(2):
    device.decode_video_dev(sd, d_out, nb, d_dec, True)
torch.cuda.synchronize()
print("bytes", nb)
