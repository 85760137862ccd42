"""A/B of the tile-kernel variants on one B200 (device-resident, CUDA events):
     python tools/ab_encode_variants.py [size] [block]
   For each ie_set_option("encode_variant", V): the stream must equal variant 0's byte for byte; prints ms per encode."""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
import imageencoder_b200 as ie
from imageencoder_b200 import _lib, device
from imageencoder_b200.synth import synth_image

L = ie.lib()
_lib.check(L.ie_init(0))
size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8
q = ie.read_matrix('tests/golden/inputs/' + ('matrix8_1.txt' if N == 8 else 'matrix4_2.txt'))
img = synth_image(size, size, 1234)
d_raw = [torch.from_numpy(np.roll(img, 8 * 37 * i, axis=0).copy()).cuda().reshape(-1) for i in range(4)]
cap = int(L.ie_max_encoded_bytes(size, size, N, 1))
d_out = [torch.empty(cap, dtype=torch.uint8, device="cuda") for _ in range(4)]
d_bits = torch.zeros(1, dtype=torch.int64, device="cuda")
sess = device.Session(0, size, size, N)


def run(n):
    for i in range(n):
        device.encode_image_dev(sess, d_raw[i % 4], q, True, d_out[i % 4], d_bits)


ref = None
for rnd in range(2):                     # two rounds: the second one is the figure to read (clocks settled)
    for v in (0, 1, 2):
        _lib.check(L.ie_set_option(b"encode_variant", v))
        run(4); torch.cuda.synchronize()
        if v == 0 and ref is None:
            ref = [t.clone() for t in d_out]
        same = all(torch.equal(t, r) for t, r in zip(d_out, ref))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(40); b.record(); torch.cuda.synchronize()
        print(f"round {rnd} variant {v}: {size}x{size} N={N}  {a.elapsed_time(b) / 40:.4f} ms/encode  identical_to_default={same}")
_lib.check(L.ie_set_option(b"encode_variant", 2))
