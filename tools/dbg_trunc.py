import sys
sys.path.insert(0, '.')
import numpy as np
import imageencoder_b200 as ie
import oracle
from imageencoder_b200.synth import synth_image
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
keeps = [float(a) for a in sys.argv[2:]] or [0.97]
W, H = 256, 192
img = synth_image(W, H, 31)
q = np.array(oracle.read_matrix('tests/golden/inputs/' + ('matrix8_1.txt' if N == 8 else 'matrix.txt')), np.uint16)
enc = oracle.image_encode(img, W, H, N, q, True, False)
for keep in keeps:
    cut = enc[: max(200, int(len(enc) * keep))]
    want = np.asarray(oracle.image_decode(cut, N)[0])
    got = ie.decode_image(cut, N)
    d = np.argwhere(got != want)
    print("len", len(enc), "cut", len(cut), "ndiff", len(d))
    if len(d):
        ys, xs = d[:, 0], d[:, 1]
        blocks = sorted(set((int(y) // N) * (W // N) + int(x) // N for y, x in d))
        print("differing blocks", blocks[:10], "...", len(blocks))
        # which block holds the cut
        _, _, coef, bl, lf = oracle.image_encode_plain(img, W, H, N, q, True, True, stages=True)
        per = 4 + bl.astype(np.int64) + lf.astype(np.int64) * bl.astype(np.int64)
        hdr = oracle.header_bits(N, q, True)
        ends = hdr + np.cumsum(per)
        print("first block ending after the cut:", int(np.searchsorted(ends, len(cut) * 8)), "of", len(per))
        b = blocks[0]
        by, bx = divmod(b, W // N)
        print("got\n", got[by*N:(by+1)*N, bx*N:(bx+1)*N], "\nwant\n", want[by*N:(by+1)*N, bx*N:(bx+1)*N])
