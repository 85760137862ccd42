import sys
sys.path.insert(0, '.')
import numpy as np
import imageencoder_b200 as ie
from imageencoder_b200 import _lib
from imageencoder_b200.synth import synth_image
_lib.check(ie.lib().ie_init(0))
q = ie.read_matrix('tests/golden/inputs/matrix.txt')
img = synth_image(512, 256, 5)
enc = ie.encode_image(img, 512, 256, q, True, False)
dec = ie.decode_image(enc, 4)
import oracle
print("ok", np.array_equal(dec, oracle.image_decode(enc, 4)[0]))
