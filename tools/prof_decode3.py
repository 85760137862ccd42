"""workload for the ncu launch list: one-GPU decode of a config-3-like stream (matrix8_2): python tools/prof_decode3.py [size]"""
import sys
sys.path.insert(0, '.')
import torch
import imageencoder_b200 as ie
from imageencoder_b200 import _lib, device
from imageencoder_b200.synth import synth_image
_lib.check(ie.lib().ie_init(0))
size = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
q = ie.read_matrix('tests/golden/inputs/' + (sys.argv[2] if len(sys.argv) > 2 else 'matrix8_2.txt'))
img = torch.from_numpy(synth_image(size, size, 1234 if len(sys.argv) > 2 else 1235)).cuda().reshape(-1)
cap = int(ie.lib().ie_max_encoded_bytes(size, size, 8, 1))
d_out = torch.zeros(cap, dtype=torch.uint8, device="cuda")
d_bits = torch.zeros(1, dtype=torch.int64, device="cuda")
sess = device.Session(device.Session.IMAGE_ENCODE, size, size, 8)
device.encode_image_dev(sess, img, q, True, d_out, d_bits)
torch.cuda.synchronize()
n = (int(d_bits.item()) + 7) // 8
hdr = device.parse_image_header(d_out[:160].cpu().numpy().tobytes(), 8)
sd = device.Session(device.Session.IMAGE_DECODE, 0, 0, 8)
full = torch.empty(size * size, dtype=torch.uint8, device="cuda")
for _ in range(2):
    device.decode_image_with_header_dev(sd, hdr, d_out, n, full)
torch.cuda.synchronize()
print("bytes", n)
