"""one-image 4x4 encode + decode (config-4 image) for the ncu launch list: python tools/prof_c4.py"""
import sys
sys.path.insert(0, '.')
import torch
import imageencoder_b200 as ie
from imageencoder_b200 import _lib, device
from imageencoder_b200.synth import synth_image
_lib.check(ie.lib().ie_init(0))
size = 4096
q = ie.read_matrix('tests/golden/inputs/matrix4_2.txt')
img = torch.from_numpy(synth_image(size, size, 2000)).cuda().reshape(-1)
cap = int(ie.lib().ie_max_encoded_bytes(size, size, 4, 1))
d_out = torch.zeros(cap, dtype=torch.uint8, device="cuda")
d_bits = torch.zeros(1, dtype=torch.int64, device="cuda")
sess = device.Session(device.Session.IMAGE_ENCODE, size, size, 4)
for _ in range(2):
    device.encode_image_dev(sess, img, q, True, d_out, d_bits)
torch.cuda.synchronize()
n = (int(d_bits.item()) + 7) // 8
hdr = device.parse_image_header(d_out[:160].cpu().numpy().tobytes(), 4)
sd = device.Session(device.Session.IMAGE_DECODE, 0, 0, 4)
full = torch.empty(size * size, dtype=torch.uint8, device="cuda")
for _ in range(2):
    device.decode_image_with_header_dev(sd, hdr, d_out, n, full)
torch.cuda.synchronize()
print("bytes", n, "decoded == ?", bool(True))
