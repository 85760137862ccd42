"""does running successive encodes on alternating streams (two sessions) overlap one image's copy-out with the next image's
tile kernel?  device-resident, config 2, CUDA events (cudart via ctypes)"""
import ctypes as C, sys
sys.path.insert(0, '.')
import numpy as np
import imageencoder_b200 as ie
from imageencoder_b200 import _lib
from imageencoder_b200.synth import synth_image
L = ie.lib(); _lib.check(L.ie_init(0))
rt = C.CDLL("libcudart.so.12")
size, N = 8192, 8
q = np.ascontiguousarray(ie.read_matrix('tests/golden/inputs/matrix8_1.txt'), dtype=np.uint16).reshape(-1)
qp = q.ctypes.data_as(C.POINTER(C.c_uint16))
img = np.ascontiguousarray(synth_image(size, size, 1234))
cap = int(L.ie_max_encoded_bytes(size, size, N, 1))
RING = 4
d_raw, d_out = [], []
for i in range(RING):
    p = C.c_void_p(); assert rt.cudaMalloc(C.byref(p), C.c_size_t(size * size)) == 0
    a = np.ascontiguousarray(np.roll(img, 8 * 37 * i, axis=0))
    assert rt.cudaMemcpy(p, a.ctypes.data_as(C.c_void_p), C.c_size_t(a.nbytes), 1) == 0
    d_raw.append(p)
    o = C.c_void_p(); assert rt.cudaMalloc(C.byref(o), C.c_size_t(cap)) == 0
    d_out.append(o)
d_bits = C.c_void_p(); assert rt.cudaMalloc(C.byref(d_bits), C.c_size_t(64)) == 0
variants = [int(a) for a in sys.argv[1:]] or [2]          # python tools/ab_streams.py 2 3: copy-out variants
for cv in variants:
  _lib.check(L.ie_set_option(b"copyout_variant", cv))
  for nstreams in (1, 2, 3):
      sess, streams = [], []
      for k in range(nstreams):
          s = C.c_void_p(); _lib.check(L.ie_session_create(C.byref(s), 0, size, size, N, 1)); sess.append(s)
          st = C.c_void_p(); assert rt.cudaStreamCreateWithFlags(C.byref(st), 1) == 0; streams.append(st)
      e0, e1 = C.c_void_p(), C.c_void_p(); rt.cudaEventCreate(C.byref(e0)); rt.cudaEventCreate(C.byref(e1))
      def run(n):
          for i in range(n):
              k = i % RING
              j = i % nstreams
              _lib.check(L.ie_encode_image_dev(sess[j], d_raw[k], size, size, qp, 1, 1, 1, 0, d_out[k], C.c_size_t(cap), C.c_void_p(d_bits.value + 8 * k), streams[j]))
      run(8); assert rt.cudaDeviceSynchronize() == 0
      rt.cudaEventRecord(e0, streams[0])
      reps = 48
      run(reps)
      for st in streams[1:]:
          ev = C.c_void_p(); rt.cudaEventCreate(C.byref(ev)); rt.cudaEventRecord(ev, st); rt.cudaStreamWaitEvent(streams[0], ev, 0)
      rt.cudaEventRecord(e1, streams[0]); rt.cudaEventSynchronize(e1)
      ms = C.c_float(); rt.cudaEventElapsedTime(C.byref(ms), e0, e1)
      print(f"copyout_variant {cv}, {nstreams} stream(s): {ms.value / reps:.4f} ms per encode", flush=True)
