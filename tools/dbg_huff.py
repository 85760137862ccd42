import sys, time
sys.path.insert(0, '.')
import numpy as np
import imageencoder_b200 as ie
from imageencoder_b200 import _lib
from imageencoder_b200.synth import synth_image
_lib.check(ie.lib().ie_init(0))
size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
q = ie.read_matrix('tests/golden/inputs/matrix8_2.txt')
img = synth_image(size, size, 1235)
enc = ie.encode_image(img, size, size, q, True, True)
for _ in range(2):
    t = time.perf_counter(); dec = ie.decode_image(enc, 8); dt = time.perf_counter() - t
print("huffman decode+image decode host ms", dt * 1e3, len(enc))
t = time.perf_counter(); enc2 = ie.encode_image(img, size, size, q, True, True); print("encode+huffman host ms", (time.perf_counter() - t) * 1e3)
plain = ie.encode_image(img, size, size, q, True, False)
print("round trip equal to plain decode:", np.array_equal(dec, ie.decode_image(plain, 8)))
