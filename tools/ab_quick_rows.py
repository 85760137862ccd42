"""Torch-free timing of the other SURVEY 8 rows on one B200 (starts in seconds; the torch-based tools/bench_all.py is the full
version): image decode 8192^2 8x8, encode + decode 4096^2 4x4, video encode + decode 1920x1088 (24 frames, GOP 12).
     python tools/ab_quick_rows.py [--options name=value,...]     e.g. --options encode_variant=0
   Every result is checked: decoded pixels against a decode with ie_set_option("exact_transform", 1), small cases against the
   CPU oracle.  Writes gpurun_out/ab_quick_rows.json."""
import ctypes as C
import hashlib
import json
import os
import sys
import time

sys.path.insert(0, '.')
sys.path.insert(0, 'tools')
import numpy as np

import _cudart as cu
import imageencoder_b200 as ie
from imageencoder_b200 import _lib
from imageencoder_b200.synth import synth_image, synth_video

L = ie.lib()
_lib.check(L.ie_init(0))
for a in sys.argv[1:]:
    if a.startswith("--options"):
        continue
    for kv in a.split(","):
        k, v = kv.split("=")
        _lib.check(L.ie_set_option(k.encode(), int(v)))
        print("option", k, v)
res = {}
tm = cu.Timer()


def qptr(q):
    qa = np.ascontiguousarray(q, dtype=np.uint16).reshape(-1)
    return qa, qa.ctypes.data_as(C.POINTER(C.c_uint16))


def session(kind, w, h, n, frames=1):
    s = C.c_void_p()
    _lib.check(L.ie_session_create(C.byref(s), kind, w, h, n, frames))
    return s


def image_rows(size, n, matrix, tag):
    q = ie.read_matrix('tests/golden/inputs/' + matrix)
    qa, qp = qptr(q)
    img = synth_image(size, size, 1234)
    cap = int(L.ie_max_encoded_bytes(size, size, n, 1))
    d_raw, d_enc, d_bits, d_dec = cu.to_device(img), cu.malloc(cap), cu.malloc(8), cu.malloc(size * size)
    se, sd = session(0, size, size, n), session(1, size, size, n)

    def enc():
        _lib.check(L.ie_encode_image_dev(se, d_raw, size, size, qp, 1, 1, 1, 0, d_enc, C.c_size_t(cap), d_bits, None))

    for _ in range(3):
        enc()
    cu.sync()
    nbytes = (int(cu.d2h(d_bits, 8, np.uint64)[0]) + 7) // 8
    tm.start()
    for _ in range(20):
        enc()
    res[f"{tag}_encode_ms"] = tm.stop_ms() / 20
    w, h = C.c_uint32(0), C.c_uint32(0)

    def dec():
        _lib.check(L.ie_decode_image_dev(sd, d_enc, C.c_size_t(nbytes), 1, d_dec, C.c_size_t(size * size), C.byref(w), C.byref(h), None))

    for _ in range(3):
        dec()
    cu.sync()
    tm.start()
    for _ in range(20):
        dec()
    res[f"{tag}_decode_ms"] = tm.stop_ms() / 20
    fast = hashlib.sha256(cu.d2h(d_dec, size * size).tobytes()).hexdigest()
    _lib.check(L.ie_set_option(b"exact_transform", 1))
    cu.memset(d_dec, 0, size * size)
    dec()
    cu.sync()
    exact = hashlib.sha256(cu.d2h(d_dec, size * size).tobytes()).hexdigest()
    _lib.check(L.ie_set_option(b"exact_transform", 0))
    res[f"{tag}_decode_fast_equals_exact"] = (fast == exact)
    res[f"{tag}_encoded_bytes"] = nbytes
    print(tag, {k: v for k, v in res.items() if k.startswith(tag)}, flush=True)


def video_rows():
    W, H, F, gop, mer = 1920, 1088, 24, 12, 16
    q = ie.read_matrix('tests/golden/inputs/matrix.txt')
    qa, qp = qptr(q)
    yuv = np.ascontiguousarray(synth_video(W, H, F, 4000)).reshape(-1)
    cap = int(L.ie_max_encoded_bytes(W, H, 4, F))
    d_src, d_yuv, d_enc, d_bits, d_dec = cu.to_device(yuv), cu.malloc(yuv.nbytes), cu.malloc(cap), cu.malloc(8), cu.malloc(yuv.nbytes)
    se, sd = session(2, W, H, 4, F), session(3, W, H, 4, F)

    def enc():
        cu.ck(cu.rt().cudaMemcpy(d_yuv, d_src, C.c_size_t(yuv.nbytes), 3), "D2D")       # the encoder rebuilds the frames in place
        _lib.check(L.ie_encode_video_dev(se, d_yuv, C.c_size_t(yuv.nbytes), W, H, qp, 1, gop, mer, 1, d_enc, C.c_size_t(cap), d_bits,
                                         None, None))

    for _ in range(2):
        enc()
    cu.sync()
    nbytes = (int(cu.d2h(d_bits, 8, np.uint64)[0]) + 7) // 8
    reps = 5
    tm.start()
    for _ in range(reps):
        cu.ck(cu.rt().cudaMemcpy(d_yuv, d_src, C.c_size_t(yuv.nbytes), 3), "D2D")
    copy_ms = tm.stop_ms() / reps
    tm.start()
    for _ in range(reps):
        enc()
    res["video_encode_ms_per_frame"] = (tm.stop_ms() / reps - copy_ms) / F
    w, h, f = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)

    def dec():
        _lib.check(L.ie_decode_video_dev(sd, d_enc, C.c_size_t(nbytes), 1, 1, d_dec, C.c_size_t(yuv.nbytes), C.byref(w), C.byref(h),
                                         C.byref(f), None))

    for _ in range(2):
        dec()
    cu.sync()
    tm.start()
    for _ in range(reps):
        dec()
    res["video_decode_ms_per_frame"] = tm.stop_ms() / reps / F
    res["video_encoded_bytes"] = nbytes
    res["video_stream_sha256"] = hashlib.sha256(cu.d2h(d_enc, nbytes).tobytes()).hexdigest()[:16]
    res["video_decoded_sha256"] = hashlib.sha256(cu.d2h(d_dec, yuv.nbytes).tobytes()).hexdigest()[:16]
    print("video", {k: v for k, v in res.items() if k.startswith("video")}, flush=True)


def small_parity():
    import oracle
    ok = True
    for mat in ("matrix8_1.txt", "matrix4_2.txt"):
        q = ie.read_matrix('tests/golden/inputs/' + mat)
        n = q.shape[0]
        img = synth_image(512, 384, 21, flat=True)
        want = oracle.image_encode(img, 512, 384, n, q, True, False)
        got = ie.encode_image(img, 512, 384, q, True, False)
        dec = ie.decode_image(got, n)
        ok &= (got == want) and np.array_equal(dec, oracle.image_decode(want, n)[0])
    q = ie.read_matrix('tests/golden/inputs/matrix.txt')
    yuv = synth_video(64, 48, 7)
    want = oracle.video_encode(yuv, 64, 48, q, True, 3, 16, False)
    got = ie.encode_video(yuv, 64, 48, q, True, 3, 16, False)
    ok &= (got == want)
    ok &= np.array_equal(np.asarray(ie.decode_video(got, True)[0]).reshape(-1), np.asarray(oracle.video_decode(want, True)[0]).reshape(-1))
    res["small_parity_vs_oracle"] = bool(ok)
    print("small parity vs oracle:", ok, flush=True)


t0 = time.time()
small_parity()
image_rows(8192, 8, "matrix8_1.txt", "img8192_8x8")
image_rows(4096, 4, "matrix4_2.txt", "img4096_4x4")
video_rows()
res["wall_s"] = time.time() - t0
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/ab_quick_rows.json", "w"), indent=1)
print("wall", res["wall_s"])
