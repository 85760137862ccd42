import sys, os
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
d_raw = [torch.from_numpy(np.roll(img, 8 * 37 * i, axis=0).copy()).cuda().reshape(-1) for i in range(4)]
cap = int(ie.lib().ie_max_encoded_bytes(size, size, N, 1))
d_out = [torch.empty(cap, dtype=torch.uint8, device="cuda") for _ in range(4)]
d_bits = torch.zeros(1, dtype=torch.int64, device="cuda")
s_enc = device.Session(0, size, size, N)
def run(n):
    for i in range(n):
        device.encode_image_dev(s_enc, d_raw[i % 4], q, True, d_out[i % 4], d_bits)
run(4); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); run(20); b.record(); torch.cuda.synchronize()
print(f"{size}x{size} N={N} ms/step", a.elapsed_time(b) / 20)
