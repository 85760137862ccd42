"""Torch-free workload for ncu: a few device-resident encodes of the bench image (config 2) with the library's defaults.
     python tools/prof_encode.py [steps]        (ncu: -k regex:'encode_tiles|tile_copyout' -s 4 -c 4)"""
import ctypes as C
import sys

sys.path.insert(0, '.')
import numpy as np

import imageencoder_b200 as ie
from imageencoder_b200 import _lib
from imageencoder_b200.synth import synth_image

L = ie.lib()
_lib.check(L.ie_init(0))
rt = C.CDLL("libcudart.so.12")
size, N = 8192, 8
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
if len(sys.argv) > 2:                                  # python tools/prof_encode.py 4 8  -> encode_variant 8
    _lib.check(L.ie_set_option(b"encode_variant", int(sys.argv[2])))
q = np.ascontiguousarray(ie.read_matrix('tests/golden/inputs/matrix8_1.txt'), dtype=np.uint16).reshape(-1)
qp = q.ctypes.data_as(C.POINTER(C.c_uint16))
img = np.ascontiguousarray(synth_image(size, size, 1234))
cap = int(L.ie_max_encoded_bytes(size, size, N, 1))
d_raw, d_out, d_bits = C.c_void_p(), C.c_void_p(), C.c_void_p()
assert rt.cudaMalloc(C.byref(d_raw), C.c_size_t(size * size)) == 0
assert rt.cudaMalloc(C.byref(d_out), C.c_size_t(cap)) == 0
assert rt.cudaMalloc(C.byref(d_bits), C.c_size_t(8)) == 0
assert rt.cudaMemcpy(d_raw, img.ctypes.data_as(C.c_void_p), C.c_size_t(img.nbytes), 1) == 0
sess = C.c_void_p()
_lib.check(L.ie_session_create(C.byref(sess), 0, size, size, N, 1))
for _ in range(steps):
    _lib.check(L.ie_encode_image_dev(sess, d_raw, size, size, qp, 1, 1, 1, 0, d_out, C.c_size_t(cap), d_bits, None))
assert rt.cudaDeviceSynchronize() == 0
bits = np.zeros(1, np.uint64)
assert rt.cudaMemcpy(bits.ctypes.data_as(C.c_void_p), d_bits, C.c_size_t(8), 2) == 0
print("encoded bytes", (int(bits[0]) + 7) // 8, "launches", ie.launch_count())
