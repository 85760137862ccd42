"""Huffman stage of the config-2 stream (27.7 MB): wall time of ie_huffman_encode_dev, and a workload for the ncu launch list"""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import imageencoder_b200 as ie
from imageencoder_b200 import _lib, device
from imageencoder_b200.synth import synth_image
_lib.check(ie.lib().ie_init(0))
W = H = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
q = ie.read_matrix('tests/golden/inputs/matrix8_1.txt')
img = torch.from_numpy(synth_image(W, H, 1234)).cuda().reshape(-1)
cap = int(ie.lib().ie_max_encoded_bytes(W, H, 8, 1))
d_plain = torch.zeros(cap, dtype=torch.uint8, device="cuda")
d_bits = torch.zeros(1, dtype=torch.int64, device="cuda")
sess = device.Session(device.Session.IMAGE_ENCODE, W, H, 8)
device.encode_image_dev(sess, img, q, True, d_plain, d_bits, lead_bit=False)
torch.cuda.synchronize()
n = (int(d_bits.item()) + 7) // 8
d_out = torch.zeros(cap + 4096, dtype=torch.uint8, device="cuda")
for _ in range(3):
    nb = device.huffman_encode_dev(sess, d_plain, n, d_out)
torch.cuda.synchronize()
t = time.perf_counter()
reps = 10
for _ in range(reps):
    nb = device.huffman_encode_dev(sess, d_plain, n, d_out)
torch.cuda.synchronize()
print(f"huffman encode of {n} bytes -> {nb} bytes: {(time.perf_counter() - t) / reps * 1e3:.3f} ms wall per call")
# the stage without host synchronisation, two sessions on two streams: one call's dictionary build (host callback) overlaps the
# other's kernels
sess2 = device.Session(device.Session.IMAGE_ENCODE, W, H, 8)
d_out2 = torch.zeros(cap + 4096, dtype=torch.uint8, device="cuda")
nbytes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(2)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
pairs = [(sess, d_out), (sess2, d_out2)]
def run(k):
    for i in range(k):
        j = i % 2
        with torch.cuda.stream(streams[j]):
            device.huffman_encode_async_dev(pairs[j][0], d_plain, n, pairs[j][1], nbytes[j])
run(4)
torch.cuda.synchronize()
t = time.perf_counter()
reps = 20
run(reps)
torch.cuda.synchronize()
print(f"async, two sessions / streams: {(time.perf_counter() - t) / reps * 1e3:.3f} ms per call; bytes {int(nbytes[0].item())} {int(nbytes[1].item())} "
      f"identical outputs: {bool(torch.equal(d_out[:nb], d_out2[:nb]))}")
with torch.cuda.stream(streams[0]):
    t = time.perf_counter()
    for _ in range(reps):
        device.huffman_encode_async_dev(sess, d_plain, n, d_out, nbytes[0])
    streams[0].synchronize()
print(f"async, one stream: {(time.perf_counter() - t) / reps * 1e3:.3f} ms per call")
