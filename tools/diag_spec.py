"""How well does the speculative parse grid synchronise on a stream?  python tools/diag_spec.py [config 2|3] [size]
   Encodes the config's synthetic image, runs the walk of the sharded decode with one part (the caller owns the per-group
   results), and counts the seams that disagree."""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
import imageencoder_b200 as ie
from imageencoder_b200 import _lib, device
from imageencoder_b200.synth import synth_image
_lib.check(ie.lib().ie_init(0))
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
size = int(sys.argv[2]) if len(sys.argv) > 2 else (8192 if cfg == 2 else 16384)
mat, seed = (("matrix8_1.txt", 1234) if cfg == 2 else ("matrix8_2.txt", 1235))
q = ie.read_matrix('tests/golden/inputs/' + mat)
img = torch.from_numpy(synth_image(size, size, seed)).cuda().reshape(-1)
cap = int(ie.lib().ie_max_encoded_bytes(size, size, 8, 1))
d_out = torch.zeros(cap, dtype=torch.uint8, device="cuda")
d_bits = torch.zeros(1, dtype=torch.int64, device="cuda")
sess = device.Session(device.Session.IMAGE_ENCODE, size, size, 8)
device.encode_image_dev(sess, img, q, True, d_out, d_bits)
torch.cuda.synchronize()
n = (int(d_bits.item()) + 7) // 8
print("stream bytes", n, "bytes/px", n / size / size)
hdr = device.parse_image_header(d_out[:160].cpu().numpy().tobytes(), 8)
total, chunk = device.decode_shard_spec_bytes(n, 8, 1)
d_spec = torch.zeros(total, dtype=torch.uint8, device="cuda")
sd = device.Session(device.Session.IMAGE_DECODE, 0, 0, 8)
device.decode_image_shard_begin_dev(sd, hdr, d_out, n, 0, 1, d_spec)
torch.cuda.synchronize()
a = d_spec.cpu().numpy().view(np.uint32).reshape(2, -1, 2)
entry, exit_, cnt = a[0, :, 0], a[1, :, 0], a[1, :, 1]
ng = (n * 8 - int(hdr.first_block_bit) + 8191) // 8192
bad = np.nonzero(entry[1:ng] != exit_[:ng - 1])[0] + 1
print("groups", ng, "bad seams", len(bad), "first", bad[:20], "blocks/group", cnt[:ng].mean())
if len(bad):
    runs = np.split(bad, np.nonzero(np.diff(bad) > 1)[0] + 1)
    print("runs of consecutive bad seams:", len(runs), "longest", max(len(r) for r in runs), "at CTA seam (g % 64 == 0):", int(sum(1 for b in bad if b % 64 == 0)))
full = torch.empty(size * size, dtype=torch.uint8, device="cuda")
for _ in range(2):
    device.decode_image_with_header_dev(sd, hdr, d_out, n, full)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    device.decode_image_with_header_dev(sd, hdr, d_out, n, full)
e1.record()
torch.cuda.synchronize()
print("one-GPU decode ms", e0.elapsed_time(e1) / 5, "decoded == source image" , bool(torch.equal(full, img)) if False else "")
if len(bad) and size <= 4096 and "--chain" in sys.argv:
    # the true chain on the host
    s = d_out[:n + 8].cpu().numpy()
    bits = np.unpackbits(s)
    pos = int(hdr.first_block_bit)
    B0 = pos
    tot = n * 8
    true_entry = {}
    g = 0
    true_entry[0] = 0
    nblk = (size // 8) ** 2
    k = 0
    def rd(p, l):
        v = 0
        for b in bits[p:p + l]:
            v = (v << 1) | int(b)
        return v
    while k < nblk and pos < tot:
        gg = (pos - B0) // 8192
        if gg not in true_entry:
            true_entry[gg] = pos - B0 - gg * 8192
        w = rd(pos, 4)
        ln = rd(pos + 4, w) if w else 0
        pos += 4 + w + ln * w
        k += 1
    print("blocks walked", k, "end", pos, "of", tot)
    for b in bad[:12]:
        lo = max(0, b - 3)
        print("seam", b, [(int(x), int(entry[x]), int(exit_[x]), true_entry.get(int(x))) for x in range(lo, b + 3)])
