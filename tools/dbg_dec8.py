import sys
sys.path.insert(0, '.')
import numpy as np, torch
import imageencoder_b200 as ie
from imageencoder_b200 import _lib, device
from imageencoder_b200.synth import synth_image
_lib.check(ie.lib().ie_init(0))
size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8
q = ie.read_matrix('tests/golden/inputs/' + ('matrix8_1.txt' if N == 8 else 'matrix4_2.txt'))
img = synth_image(size, size, 1234)
d_raw = torch.from_numpy(img).cuda().reshape(-1)
cap = int(ie.lib().ie_max_encoded_bytes(size, size, N, 1))
d_out = torch.empty(cap, dtype=torch.uint8, device="cuda"); d_bits = torch.zeros(1, dtype=torch.int64, device="cuda")
s_enc = device.Session(0, size, size, N)
device.encode_image_dev(s_enc, d_raw, q, True, d_out, d_bits); torch.cuda.synchronize()
nbytes = (int(d_bits.item()) + 7) // 8
s_dec = device.Session(1, size, size, N)
d_dec = torch.empty(size * size, dtype=torch.uint8, device="cuda")
for _ in range(3):
    device.decode_image_dev(s_dec, d_out, nbytes, d_dec, 1)
torch.cuda.synchronize()
print("ok", nbytes)
