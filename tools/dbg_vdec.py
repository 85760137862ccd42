import sys, time, hashlib
sys.path.insert(0, '.')
import numpy as np, torch
import imageencoder_b200 as ie
from imageencoder_b200 import _lib, device
from imageencoder_b200.synth import synth_video, synth_image
L = ie.lib(); _lib.check(L.ie_init(0))
W, H, F = 1920, 1088, 48
q = ie.read_matrix('tests/golden/inputs/matrix.txt')
yuv = synth_video(W, H, F, 4000)
d_yuv = torch.from_numpy(np.ascontiguousarray(yuv)).cuda().reshape(-1)
d_out = torch.empty(int(L.ie_max_encoded_bytes(W, H, 4, F)) + 4096, dtype=torch.uint8, device="cuda")
d_bits = torch.zeros(1, dtype=torch.int64, device="cuda")
sv = device.Session(2, W, H, 4, F)
device.encode_video_dev(sv, d_yuv, W, H, q, True, 12, 16, d_out, d_bits, lead_bit=True)
torch.cuda.synchronize()
nb = (int(d_bits.item()) + 7) // 8
sd = device.Session(3, W, H, 4, F)
d_dec = torch.empty(W * H * 3 // 2 * F, dtype=torch.uint8, device="cuda")
ref = None
for grid4 in (0, 1, 2, -1, 0, 1, 2, -1):
    _lib.check(L.ie_set_option(b"parse_grid4", grid4))
    device.decode_video_dev(sd, d_out, nb, d_dec, True); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(3):
        device.decode_video_dev(sd, d_out, nb, d_dec, True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    h = hashlib.sha256(d_dec.cpu().numpy().tobytes()).hexdigest()[:12]
    ref = ref or h
    print(f"parse_grid4 {grid4}: {dt / 3 / F * 1e3:.4f} ms per frame, same pixels: {h == ref}", flush=True)
# 4x4 image decode too
q4 = ie.read_matrix('tests/golden/inputs/matrix4_2.txt')
img = synth_image(4096, 4096, 2000)
enc = ie.encode_image(img, 4096, 4096, q4, True, False)
d_enc = torch.zeros(len(enc) + 64, dtype=torch.uint8, device="cuda"); d_enc[:len(enc)] = torch.frombuffer(bytearray(enc), dtype=torch.uint8).cuda()
d_raw = torch.zeros(4096 * 4096, dtype=torch.uint8, device="cuda")
s4 = device.Session(1, 4096, 4096, 4)
hdr = device.parse_image_header(enc[:160], 4)
for grid4 in (0, 1, 0, 1):
    _lib.check(L.ie_set_option(b"parse_grid4", grid4))
    device.decode_image_with_header_dev(s4, hdr, d_enc, len(enc), d_raw); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        device.decode_image_with_header_dev(s4, hdr, d_enc, len(enc), d_raw)
    b.record(); torch.cuda.synchronize()
    ok = np.array_equal(d_raw.cpu().numpy().reshape(4096, 4096), ie.decode_image(enc, 4)) 
    print(f"4096^2 4x4 decode parse_grid4 {grid4}: {a.elapsed_time(b) / 10:.4f} ms, ok {ok}", flush=True)
