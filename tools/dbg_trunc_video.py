import sys
sys.path.insert(0, '.')
import numpy as np
import imageencoder_b200 as ie
import oracle
from imageencoder_b200.synth import synth_video
W, H, F = 96, 64, 7
keep = float(sys.argv[1]) if len(sys.argv) > 1 else 0.45
q = np.array(oracle.read_matrix('tests/golden/inputs/matrix.txt'), np.uint16)
yuv = synth_video(W, H, F, 4000)
enc = oracle.video_encode(yuv, W, H, q, True, 3, 8, False)
cut = enc[: int(len(enc) * keep)]
want = np.asarray(oracle.video_decode(cut, True)[0])
got, w, h, f = ie.decode_video(cut, True)
fsz = W * H * 3 // 2
print("enc", len(enc), "cut", len(cut), "bits", len(cut) * 8)
for fr in range(F):
    a = got[fr * fsz: fr * fsz + W * H].reshape(H, W); b = want[fr * fsz: fr * fsz + W * H].reshape(H, W)
    d = np.argwhere(a != b)
    blocks = sorted(set((int(y) // 4) * (W // 4) + int(x) // 4 for y, x in d))
    mbs = sorted(set((int(y) // 16) * (W // 16) + int(x) // 16 for y, x in d))
    print("frame", fr, "ndiff", len(d), "blocks", blocks[:8], len(blocks), "MBs", mbs[:8], len(mbs))
    if len(d):
        y, x = d[0]
        print("  first diff at", y, x, "got", a[y, x], "want", b[y, x], " got row", a[y, x:x+8], "want row", b[y, x:x+8])
