import sys, hashlib
sys.path.insert(0, '.')
import numpy as np
import imageencoder_b200 as ie
from imageencoder_b200 import _lib
from imageencoder_b200.synth import synth_image
L = ie.lib(); _lib.check(L.ie_init(0))
for (W, H, mat, seed) in ((2048, 2048, 'matrix8_2.txt', 1235), (8192, 8192, 'matrix8_2.txt', 1235), (16384, 16384, 'matrix8_2.txt', 1235)):
    q = ie.read_matrix('tests/golden/inputs/' + mat)
    img = synth_image(W, H, seed)
    outs = []
    for v in (0, 1):
        _lib.check(L.ie_set_option(b"huffman_variant", v))
        enc = ie.encode_image(img, W, H, q, True, True)
        outs.append(enc)
        print(W, H, "variant", v, len(enc), hashlib.sha256(enc).hexdigest()[:16], flush=True)
    if outs[0] != outs[1]:
        a, b = np.frombuffer(outs[0], np.uint8), np.frombuffer(outs[1], np.uint8)
        n = min(len(a), len(b))
        d = np.nonzero(a[:n] != b[:n])[0]
        print("  first diff at byte", int(d[0]) if len(d) else None, "of", n, "; ndiff", len(d))
