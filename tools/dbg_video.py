import sys, time
sys.path.insert(0, '.')
import numpy as np
import imageencoder_b200 as ie
from imageencoder_b200 import _lib
from imageencoder_b200.synth import synth_video
_lib.check(ie.lib().ie_init(0))
W, H, F = 1920, 1088, 13
yuv = synth_video(W, H, F, 4000)
q = ie.read_matrix('tests/golden/inputs/matrix.txt')
for _ in range(2):
    t = time.perf_counter(); enc = ie.encode_video(yuv, W, H, q, True, 12, 16, False); dt = time.perf_counter() - t
print("video encode ms/frame", dt * 1e3 / F, len(enc))
for _ in range(2):
    t = time.perf_counter(); dec = ie.decode_video(enc, True); dt = time.perf_counter() - t
print("video decode ms/frame", dt * 1e3 / F)
