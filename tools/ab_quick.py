"""Torch-free A/B of the tile-kernel variants (starts in seconds on a fresh box):
     python tools/ab_quick.py [size]
   1. small images, every variant through the host API: streams must equal the default kernel's AND the CPU oracle's;
   2. size x size (default 8192, config 2), device-resident through ie_encode_image_dev on the legacy default stream, timed with
      CUDA events (cudart via ctypes): ms per encode of each variant, streams compared with the default kernel's.
   Writes gpurun_out/ab_quick.json."""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, '.')
import numpy as np

import imageencoder_b200 as ie
from imageencoder_b200 import _lib
from imageencoder_b200.synth import synth_image

t_start = time.time()
L = ie.lib()
_lib.check(L.ie_init(0))
res = {"parity_small": {}, "timing": {}}
COMBOS = [(0, 0), (2, 0), (2, 1), (2, 2), (3, 2), (4, 2), (5, 2), (6, 2), (7, 2), (8, 2)]          # (encode_variant, copyout_variant)
if "--combos" in sys.argv:                       # e.g. --combos 2:2,8:2
    i = sys.argv.index("--combos")
    COMBOS = [tuple(int(x) for x in c.split(":")) for c in sys.argv[i + 1].split(",")]
    del sys.argv[i:i + 2]

PADS = [0]
if "--pads" in sys.argv:                         # occupancy experiment: extra dynamic shared memory per tile-kernel CTA
    i = sys.argv.index("--pads")
    PADS = [int(x) for x in sys.argv[i + 1].split(",")]
    del sys.argv[i:i + 2]

# ---- 1. parity on small images against the oracle
import oracle
for mat in ("matrix8_1.txt", "matrix4_2.txt"):
    q = ie.read_matrix('tests/golden/inputs/' + mat)
    n = q.shape[0]
    for name, img in (("synth", synth_image(512, 384, 21)), ("flat", synth_image(512, 384, 22, flat=True)),
                      ("noise", np.random.default_rng(3).integers(0, 256, (128, 96)).astype(np.uint8))):
        h, w = img.shape
        want = oracle.image_encode(img, w, h, n, q, True, False)
        for v, cv in COMBOS:
            _lib.check(L.ie_set_option(b"encode_variant", v))
            _lib.check(L.ie_set_option(b"copyout_variant", cv))
            got = ie.encode_image(img, w, h, q, True, False)
            res["parity_small"][f"{mat}/{name}/v{v}c{cv}"] = (got == want)
_lib.check(L.ie_set_option(b"encode_variant", 2))
_lib.check(L.ie_set_option(b"copyout_variant", 3))
if not all(res["parity_small"].values()):
    print("PARITY FAILURE on small images")
print("parity (small, vs oracle): all identical =", all(res["parity_small"].values()), flush=True)

# ---- 2. device-resident timing
rt = None
for cand in ("libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so.12", "/usr/local/cuda/targets/x86_64-linux/lib/libcudart.so.12"):
    try:
        rt = C.CDLL(cand)
        break
    except OSError:
        continue
if rt is None:
    import glob
    c = glob.glob(os.path.join(sys.prefix, "lib/python*/site-packages/nvidia/cuda_runtime/lib/libcudart.so.12"))
    rt = C.CDLL(c[0])


def ck(e, what):
    if e != 0:
        raise SystemExit(f"{what}: cuda error {e}")


size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
N = 8
q = ie.read_matrix('tests/golden/inputs/matrix8_1.txt')
qa = np.ascontiguousarray(q, dtype=np.uint16).reshape(-1)
qp = qa.ctypes.data_as(C.POINTER(C.c_uint16))
img = synth_image(size, size, 1234)
RING = 4
cap = int(L.ie_max_encoded_bytes(size, size, N, 1))
d_raw, d_out = [], []
for i in range(RING):
    p = C.c_void_p()
    ck(rt.cudaMalloc(C.byref(p), C.c_size_t(size * size)), "cudaMalloc")
    a = np.ascontiguousarray(np.roll(img, 8 * 37 * i, axis=0))
    ck(rt.cudaMemcpy(p, a.ctypes.data_as(C.c_void_p), C.c_size_t(a.nbytes), 1), "H2D")
    d_raw.append(p)
    o = C.c_void_p()
    ck(rt.cudaMalloc(C.byref(o), C.c_size_t(cap)), "cudaMalloc")
    ck(rt.cudaMemset(o, 0, C.c_size_t(cap)), "memset")
    d_out.append(o)
d_bits = C.c_void_p()
ck(rt.cudaMalloc(C.byref(d_bits), C.c_size_t(8 * RING)), "cudaMalloc")
sess = C.c_void_p()
_lib.check(L.ie_session_create(C.byref(sess), 0, size, size, N, 1))
e0, e1 = C.c_void_p(), C.c_void_p()
ck(rt.cudaEventCreate(C.byref(e0)), "event")
ck(rt.cudaEventCreate(C.byref(e1)), "event")


def run(n):
    for i in range(n):
        k = i % RING
        _lib.check(L.ie_encode_image_dev(sess, d_raw[k], size, size, qp, 1, 1, 1, 0, d_out[k], C.c_size_t(cap),
                                         C.c_void_p(d_bits.value + 8 * k), None))


def fetch():
    bits = np.zeros(RING, np.uint64)
    ck(rt.cudaMemcpy(bits.ctypes.data_as(C.c_void_p), d_bits, C.c_size_t(8 * RING), 2), "D2H")
    outs = []
    for k in range(RING):
        nb = int((int(bits[k]) + 7) // 8)
        b = np.empty(nb, np.uint8)
        ck(rt.cudaMemcpy(b.ctypes.data_as(C.c_void_p), d_out[k], C.c_size_t(nb), 2), "D2H")
        outs.append(b.tobytes())
    return outs


ref = None
for pad in PADS[1:] + PADS[:1]:
    _lib.check(L.ie_set_option(b"encode_pad_smem", pad))
    if len(PADS) == 1:
        break
    _lib.check(L.ie_set_option(b"encode_variant", 2))
    _lib.check(L.ie_set_option(b"copyout_variant", 3))
    run(RING)
    ck(rt.cudaDeviceSynchronize(), "sync")
    reps = 40
    ck(rt.cudaEventRecord(e0, None), "record")
    run(reps)
    ck(rt.cudaEventRecord(e1, None), "record")
    ck(rt.cudaEventSynchronize(e1), "sync")
    ms = C.c_float()
    ck(rt.cudaEventElapsedTime(C.byref(ms), e0, e1), "elapsed")
    res["timing"][f"pad{pad}"] = ms.value / reps
    print(f"encode_pad_smem {pad}: {ms.value / reps:.4f} ms/encode", flush=True)
for rnd in range(2):
    for v, cv in COMBOS:
        _lib.check(L.ie_set_option(b"encode_variant", v))
        _lib.check(L.ie_set_option(b"copyout_variant", cv))
        run(RING)
        ck(rt.cudaDeviceSynchronize(), "sync")
        outs = fetch()
        if ref is None:
            ref = outs
        same = outs == ref
        reps = 40
        ck(rt.cudaEventRecord(e0, None), "record")
        run(reps)
        ck(rt.cudaEventRecord(e1, None), "record")
        ck(rt.cudaEventSynchronize(e1), "sync")
        ms = C.c_float()
        ck(rt.cudaEventElapsedTime(C.byref(ms), e0, e1), "elapsed")
        res["timing"][f"round{rnd}/v{v}c{cv}"] = {"ms_per_encode": ms.value / reps, "identical_to_default": same, "bytes": len(outs[0])}
        print(f"round {rnd} encode_variant {v} copyout_variant {cv}: {size}x{size}  {ms.value / reps:.4f} ms/encode  identical_to_default={same}",
              flush=True)
_lib.check(L.ie_set_option(b"encode_variant", 2))
_lib.check(L.ie_set_option(b"copyout_variant", 3))
res["wall_s"] = time.time() - t_start
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/ab_quick.json", "w") as f:
    json.dump(res, f, indent=1)
print("wall", res["wall_s"])
