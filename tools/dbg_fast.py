import sys, time
sys.path.insert(0, '.')
import numpy as np
import imageencoder_b200 as ie
from imageencoder_b200 import _lib
from imageencoder_b200.synth import synth_image
_lib.check(ie.lib().ie_init(0))
L = ie.lib()
case = sys.argv[1]
q8 = ie.read_matrix('tests/golden/inputs/matrix8_1.txt'); q4 = ie.read_matrix('tests/golden/inputs/matrix.txt')
rng = np.random.default_rng(9)
cases = {"synth8192": (lambda: synth_image(8192, 8192, 1234), q8), "synth4096_4x4": (lambda: synth_image(4096, 4096, 2000), q4),
         "ties8": (lambda: (rng.integers(0, 2, (1024, 1024)) * 16 + 120).astype(np.uint8), q8),
         "ties4": (lambda: (rng.integers(0, 4, (1024, 1024)) * 8 + 112).astype(np.uint8), q4),
         "noise8": (lambda: rng.integers(0, 256, (1024, 1024)).astype(np.uint8), np.ones((8, 8), np.uint16)),
         "small8": (lambda: synth_image(1024, 1024, 1234), q8)}
img, q = cases[case][0](), cases[case][1]
H, W = img.shape
t = time.time(); fast = ie.encode_image(img, W, H, q, True, False); print(case, "fast ok", len(fast), time.time() - t, flush=True)
L.ie_set_option(b"exact_transform", 1)
t = time.time(); exact = ie.encode_image(img, W, H, q, True, False); print(case, "exact ok", len(exact), time.time() - t, fast == exact, flush=True)
