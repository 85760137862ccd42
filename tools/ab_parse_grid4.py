"""A/B of the speculative grid of 4x4 image streams (ie_set_option("parse_grid4", 0 | 1 | 2)): decode of a config-4 image"""
import sys
sys.path.insert(0, '.')
import torch
import imageencoder_b200 as ie
from imageencoder_b200 import _lib, device
from imageencoder_b200.synth import synth_image
_lib.check(ie.lib().ie_init(0))
size = 4096
q = ie.read_matrix('tests/golden/inputs/matrix4_2.txt')
src = synth_image(size, size, 2000)
img = torch.from_numpy(src).cuda().reshape(-1)
cap = int(ie.lib().ie_max_encoded_bytes(size, size, 4, 1))
d_out = torch.zeros(cap, dtype=torch.uint8, device="cuda")
d_bits = torch.zeros(1, dtype=torch.int64, device="cuda")
sess = device.Session(device.Session.IMAGE_ENCODE, size, size, 4)
device.encode_image_dev(sess, img, q, True, d_out, d_bits)
torch.cuda.synchronize()
n = (int(d_bits.item()) + 7) // 8
hdr = device.parse_image_header(d_out[:160].cpu().numpy().tobytes(), 4)
sd = device.Session(device.Session.IMAGE_DECODE, 0, 0, 4)
full = torch.empty(size * size, dtype=torch.uint8, device="cuda")
ref = None
for grid in (0, 1, 2, 0, 1, 2):
    _lib.check(ie.lib().ie_set_option(b"parse_grid4", grid))
    for _ in range(3):
        device.decode_image_with_header_dev(sd, hdr, d_out, n, full)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        device.decode_image_with_header_dev(sd, hdr, d_out, n, full)
    e1.record()
    torch.cuda.synchronize()
    if ref is None:
        ref = full.clone()
    print(f"parse_grid4 {grid}: {e0.elapsed_time(e1) / 20:.4f} ms per decode, pixels identical: {bool(torch.equal(full, ref))}")
_lib.check(ie.lib().ie_set_option(b"parse_grid4", -1))
