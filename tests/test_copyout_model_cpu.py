"""CPU model of the copy-out by tile and word (imageencoder_b200/csrc/encode_image.cu, tile_copyout_words_kernel): the ownership
rule -- a stream word belongs to the tile that holds its first bit, its owner ORs in the first bits of the next tile, the stream's
first tile merges into the word it shares with the header, the last tile zero-pads its 128-bit chunk -- restated in numpy and checked
against plain bit concatenation on the edge cases the kernel's comment names (a last tile shorter than a word, tiles ending on word
boundaries, a header ending inside / on a word).  The GPU tests compare the kernel itself with the oracle byte for byte; this model
pins the rule the kernel implements, so a change of the rule shows up without a GPU."""
import numpy as np
import pytest


def _image(bits):
    """tile image: 32-bit words, MSB first, zero beyond the last bit"""
    n = (len(bits) + 31) // 32
    padded = np.zeros(n * 32, np.uint8)
    padded[:len(bits)] = bits
    return [int("".join(map(str, padded[32 * i: 32 * i + 32])), 2) for i in range(n)]


def _copyout_words(header_bits, tiles):
    """the kernel's per-tile work in program order of any schedule (the writes of different tiles never touch the same word,
    except tile 0's read-modify-write of the header word, which only needs the header to be there first)"""
    G0 = len(header_bits)
    total = G0 + sum(len(t) for t in tiles)
    out = [None] * (((total + 127) // 128) * 4 + 4)
    for i, w in enumerate(_image(header_bits)):           # what the tile kernel's prefix writer / an earlier append left
        out[i] = w
    for i in range((G0 + 31) // 32, ((G0 + 127) // 128) * 4):
        out[i] = 0                                          # the header's chunk is zero padded (append contract)
    images = [_image(t) for t in tiles]
    B = G0
    for j, t in enumerate(tiles):
        B0, B1 = B, B + len(t)
        B = B1
        if not len(t):
            continue
        I, nw = images[j], (len(t) + 31) // 32
        q0, q1 = (B0 + 31) >> 5, (B1 + 31) >> 5
        sh = (32 - (B0 & 31)) & 31
        has_next = j + 1 < len(tiles)
        tail = B1 & 31
        if j == 0 and (B0 & 31):
            out[B0 >> 5] |= I[0] >> (B0 & 31)
        for k in range(q1 - q0):
            lo = I[k] if k < nw else 0
            hi = I[k + 1] if k + 1 < nw else 0
            v = ((lo << sh) | (hi >> (32 - sh))) & 0xFFFFFFFF if sh else lo
            if k + 1 == q1 - q0 and tail and has_next:
                v |= images[j + 1][0] >> tail
            assert out[q0 + k] is None or out[q0 + k] == 0, "a word written twice"
            out[q0 + k] = v
        if not has_next:
            for q in range(q1, ((B1 + 127) >> 7) << 2):
                out[q] = 0
    return out, total


@pytest.mark.parametrize("hdr,lens", [
    (0, [512, 700, 33]), (7, [512, 1024, 5]), (32, [512, 544, 31]), (37, [545, 512, 1]), (128, [640, 640]),
    (100, [513] * 9 + [3]), (1, [2048, 4099, 2048, 17]), (95, [600]), (64, [512, 512, 512]), (31, [544, 33]),
])
def test_words_rule_equals_bit_concatenation(hdr, lens):
    rng = np.random.default_rng(hdr * 1000 + len(lens))
    header = rng.integers(0, 2, hdr, dtype=np.uint8)
    tiles = [rng.integers(0, 2, n, dtype=np.uint8) for n in lens]
    out, total = _copyout_words(header, tiles)
    want = np.concatenate([header] + tiles)
    nwords = ((total + 127) // 128) * 4
    got = np.zeros(nwords * 32, np.uint8)
    for i in range(nwords):
        assert out[i] is not None, f"word {i} never written"
        got[32 * i: 32 * i + 32] = [(out[i] >> (31 - b)) & 1 for b in range(32)]
    assert np.array_equal(got[:total], want)
    assert not got[total:].any(), "the last chunk is not zero padded"
