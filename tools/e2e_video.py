"""End to end through the host entry points ie_encode_video / ie_decode_video on config 5 with pinned host buffers (the copies are
inside the timed region); the stream's and the frames' sha256 are compared with the reference's (tests/golden/golden_configs.json)"""
import ctypes as C
import hashlib
import json
import sys
import time
sys.path.insert(0, '.')
import numpy as np
import torch
import imageencoder_b200 as ie
from imageencoder_b200 import _lib
from imageencoder_b200.synth import synth_video
L = ie.lib()
_lib.check(L.ie_init(0))
W, H, F = 1920, 1088, 240
q = np.ascontiguousarray(ie.read_matrix('tests/golden/inputs/matrix.txt'), dtype=np.uint16).reshape(-1)
qp = q.ctypes.data_as(C.POINTER(C.c_uint16))
g = json.load(open('tests/golden/golden_configs.json'))['C5|1920x1088x240|gop12|mer16|matrix|seed4000']
yuv = torch.from_numpy(np.ascontiguousarray(synth_video(W, H, F, 4000)).reshape(-1)).pin_memory()
cap = int(L.ie_max_encoded_bytes(W, H, 4, F))
enc = torch.empty(cap, dtype=torch.uint8).pin_memory()
dec = torch.empty(yuv.numel(), dtype=torch.uint8).pin_memory()
nb = C.c_size_t(0)
res = {}
for rep in range(4):
    t = time.perf_counter()
    _lib.check(L.ie_encode_video(C.c_void_p(yuv.data_ptr()), yuv.numel(), W, H, qp, 1, 12, 16, 0, C.c_void_p(enc.data_ptr()), cap, C.byref(nb)))
    te = time.perf_counter() - t
    nby = C.c_size_t(0)
    w, h, f = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
    t = time.perf_counter()
    _lib.check(L.ie_decode_video(C.c_void_p(enc.data_ptr()), nb.value, 1, C.c_void_p(dec.data_ptr()), dec.numel(), C.byref(nby), C.byref(w), C.byref(h), C.byref(f)))
    td = time.perf_counter() - t
    res = {"encode_ms": te * 1e3, "decode_ms": td * 1e3}
    print(f"rep {rep}: ie_encode_video {te * 1e3:.2f} ms, ie_decode_video {td * 1e3:.2f} ms", flush=True)
px = W * H * F
res.update({"encode_gpx_s_e2e": px / res["encode_ms"] / 1e6, "decode_gpx_s_e2e": px / res["decode_ms"] / 1e6,
            "h2d_bytes_encode": yuv.numel(), "d2h_bytes_encode": nb.value,
            "enc_sha_ok": hashlib.sha256(enc[:nb.value].numpy().tobytes()).hexdigest() == g["enc_sha256"],
            "dec_sha_ok": hashlib.sha256(dec.numpy().tobytes()).hexdigest() == g["dec_mc1_sha256"]})
print(json.dumps(res))
