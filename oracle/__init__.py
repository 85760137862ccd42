"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the block-transform hot path.

Python face of ``oracle/liboracle.so`` (oracle_block.c + oracle_huffman.cpp, the plain restatement of
ThenTech/ImageEncoder's Block / BitStream / Huffman / P-frame path) and of ``oracle/_ref/`` (the UNMODIFIED
reference compiled by ``oracle/build_ref.sh``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this package, and only as the checker / reported baseline.  The product (``imageencoder_b200``) never
imports it and has no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "liboracle.so"
REF_DIR = HERE / "_ref"

_u8p = C.POINTER(C.c_uint8)
_u16p = C.POINTER(C.c_uint16)
_i16p = C.POINTER(C.c_int16)
_u32p = C.POINTER(C.c_uint32)


def build(ref: bool = True) -> None:
    """Compile the oracle (and the reference when /root/reference exists)."""
    subprocess.run(["make", "-s", "-C", str(HERE), "liboracle.so"], check=True)
    if ref:
        subprocess.run(["make", "-s", "-C", str(HERE), "ref"], check=True)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            build(ref=False)
        L = C.CDLL(str(LIB_PATH))
        L.orc_image_encode.restype = C.c_longlong
        L.orc_image_encode.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u16p, C.c_int, C.c_int, _u8p, C.c_size_t,
                                       _i16p, _u8p, _u8p]
        L.orc_image_decode.restype = C.c_int
        L.orc_image_decode.argtypes = [_u8p, C.c_size_t, C.c_size_t, C.c_int, _u8p, C.c_size_t,
                                       C.POINTER(C.c_int), C.POINTER(C.c_int), _i16p]
        L.orc_video_encode.restype = C.c_longlong
        L.orc_video_encode.argtypes = [_u8p, C.c_size_t, C.c_int, C.c_int, _u16p, C.c_int, C.c_int, C.c_int, C.c_int,
                                       _u8p, C.c_size_t, _i16p]
        L.orc_video_decode.restype = C.c_int
        L.orc_video_decode.argtypes = [_u8p, C.c_size_t, C.c_size_t, C.c_int, _u8p, C.c_size_t] + [C.POINTER(C.c_int)] * 5
        L.orc_huffman_encode.restype = C.c_longlong
        L.orc_huffman_encode.argtypes = [_u8p, C.c_size_t, _u8p, C.c_size_t, _u32p, _u32p, C.POINTER(C.c_int)]
        L.orc_huffman_decode.restype = C.c_longlong
        L.orc_huffman_decode.argtypes = [_u8p, C.c_size_t, _u8p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.orc_zigzag.argtypes = [C.c_int, _u8p]
        L.orc_cos_table.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_header_bits.argtypes = [C.c_int, _u16p, C.c_int, C.c_int]
        L.orc_bits_needed_pub.argtypes = [C.c_int]
        L.orc_bits_needed_pub.restype = C.c_uint
        L.orc_ffs_pub.argtypes = [C.c_uint32]
        L.orc_ffs_pub.restype = C.c_uint
        L.orc_round_to_byte.argtypes = [C.c_size_t]
        L.orc_round_to_byte.restype = C.c_size_t
        _lib = L
    return _lib


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(t)


def _quant(q, N) -> np.ndarray:
    q = np.ascontiguousarray(np.asarray(q, dtype=np.uint16).reshape(-1))
    assert q.size == N * N, f"quant matrix must have {N * N} entries"
    return q


def read_matrix(path) -> np.ndarray:
    """MatrixReader.cpp:65-134: N rows of N whitespace-separated u16."""
    rows = [ln.split() for ln in Path(path).read_text().splitlines() if ln.strip()]
    return np.array([[int(v) for v in r] for r in rows], dtype=np.uint16)


def zigzag(N: int) -> np.ndarray:
    z = np.zeros(N * N, dtype=np.uint8)
    lib().orc_zigzag(N, _p(z, _u8p))
    return z


def bits_needed(v: int) -> int:
    return int(lib().orc_bits_needed_pub(int(v)))


def ffs(v: int) -> int:
    return int(lib().orc_ffs_pub(int(v)))


def header_bits(N, quant, lead_bit=True, video=False) -> int:
    q = _quant(quant, N)
    return int(lib().orc_header_bits(N, _p(q, _u16p), int(lead_bit), int(video)))


def image_encode_plain(raw: np.ndarray, W: int, H: int, N: int, quant, rle: bool = True, lead_bit: bool = True,
                       stages: bool = False):
    """Plain (pre-Huffman) stream.  Returns (bytes, nbits[, coef_zz, bitlen, lenfield])."""
    raw = np.ascontiguousarray(raw, dtype=np.uint8).reshape(-1)
    assert raw.size == W * H
    q = _quant(quant, N)
    nblk = (W // N) * (H // N)
    cap = (600 + nblk * (4 + 16 + 16 * N * N)) // 8 + 16
    out = np.zeros(cap, dtype=np.uint8)
    if stages:
        coef = np.zeros(nblk * N * N, dtype=np.int16)
        bl = np.zeros(nblk, dtype=np.uint8)
        lf = np.zeros(nblk, dtype=np.uint8)
        args = (_p(coef, _i16p), _p(bl, _u8p), _p(lf, _u8p))
    else:
        args = (None, None, None)
    bits = lib().orc_image_encode(_p(raw, _u8p), W, H, N, _p(q, _u16p), int(rle), int(lead_bit), _p(out, _u8p), cap, *args)
    if bits < 0:
        raise ValueError(f"orc_image_encode failed: {bits}")
    data = out[: (bits + 7) // 8].tobytes()
    if stages:
        return data, int(bits), coef.reshape(nblk, N * N), bl, lf
    return data, int(bits)


def huffman_encode(data: bytes, with_dict: bool = False):
    a = np.frombuffer(data, dtype=np.uint8)
    cap = len(data) + 4096
    out = np.zeros(cap, dtype=np.uint8)
    lens = np.zeros(256, dtype=np.uint32)
    words = np.zeros(256, dtype=np.uint32)
    rev = C.c_int(0)
    n = lib().orc_huffman_encode(_p(a, _u8p), a.size, _p(out, _u8p), cap, _p(lens, _u32p), _p(words, _u32p), C.byref(rev))
    if n < 0:
        raise ValueError(f"orc_huffman_encode failed: {n}")
    res = out[:n].tobytes()
    if with_dict:
        return res, lens, words, bool(rev.value)
    return res


def huffman_header_overflows(plain: bytes) -> bool:
    """True when the dictionary header cannot represent the code (Huffman.cpp:39-42: length & 0xF, group & 0x7F):
    the reference then cannot decode its own output, so only encoder parity is defined."""
    _, lens, _, rev = huffman_encode(plain, with_dict=True)
    if rev:
        return False
    l = lens[lens != 0xFFFFFFFF]
    return bool(l.max() >= 16 or np.bincount(l).max() >= 128)


def huffman_decode(data: bytes):
    """Returns (bytes, start_bit)."""
    a = np.frombuffer(data, dtype=np.uint8)
    cap = max(64, len(data) * 9)
    while True:
        out = np.zeros(cap, dtype=np.uint8)
        sb = C.c_size_t(0)
        n = lib().orc_huffman_decode(_p(a, _u8p), a.size, _p(out, _u8p), cap, C.byref(sb))
        if n == -1:
            cap *= 2
            continue
        if n < 0:
            raise ValueError(f"orc_huffman_decode failed: {n}")
        return out[:n].tobytes(), int(sb.value)


def image_encode(raw, W, H, N, quant, rle=True, huffman=False) -> bytes:
    """What `encoder <conf>` writes (ImageEncoder.cpp:52-175 + ImageBase.cpp:315-336)."""
    data, _ = image_encode_plain(raw, W, H, N, quant, rle, lead_bit=not huffman)
    return huffman_encode(data) if huffman else data


def image_decode(enc: bytes, N: int, stages: bool = False):
    """What `decoder <conf>` writes.  Returns (raw u8 array HxW, W, H)."""
    plain, sb = huffman_decode(enc)
    a = np.frombuffer(plain, dtype=np.uint8)
    W = C.c_int(0)
    H = C.c_int(0)
    cap = 32767 * 32767 if False else 1 << 20
    while True:
        out = np.zeros(cap, dtype=np.uint8)
        coef = np.zeros(cap, dtype=np.int16) if stages else None
        rc = lib().orc_image_decode(_p(a, _u8p), a.size, sb, N, _p(out, _u8p), cap, C.byref(W), C.byref(H),
                                    _p(coef, _i16p) if stages else None)
        if rc == -3:
            cap = W.value * H.value
            continue
        if rc < 0:
            raise ValueError(f"orc_image_decode failed: {rc}")
        break
    img = out[: W.value * H.value].reshape(H.value, W.value).copy()
    if stages:
        return img, W.value, H.value, coef[: W.value * H.value].reshape(-1, N * N).copy()
    return img, W.value, H.value


def video_encode(yuv: np.ndarray, W, H, quant, rle=True, gop=4, merange=16, huffman=False, stages=False):
    """Returns encoded bytes (and, with stages, the mvecs [frames, MBs, 2] and the encoder-side reconstruction)."""
    buf = np.array(yuv, dtype=np.uint8).reshape(-1).copy()
    q = _quant(quant, 4)
    fsz = W * H * 3 // 2
    frames = buf.size // fsz
    nmb = (W // 16) * (H // 16)
    cap = 64 + buf.size * 3 + 1024
    out = np.zeros(cap, dtype=np.uint8)
    mv = np.zeros(max(1, frames * nmb * 2), dtype=np.int16)
    bits = lib().orc_video_encode(_p(buf, _u8p), buf.size, W, H, _p(q, _u16p), int(rle), gop, merange, int(not huffman),
                                  _p(out, _u8p), cap, _p(mv, _i16p))
    if bits < 0:
        raise ValueError(f"orc_video_encode failed: {bits}")
    data = out[: (bits + 7) // 8].tobytes()
    if huffman:
        data = huffman_encode(data)
    if stages:
        return data, mv[: frames * nmb * 2].reshape(frames, nmb, 2), buf
    return data


def video_decode(enc: bytes, motioncomp: bool = True):
    """Returns (yuv bytes array, W, H, frames, gop, merange)."""
    plain, sb = huffman_decode(enc)
    a = np.frombuffer(plain, dtype=np.uint8)
    vals = [C.c_int(0) for _ in range(5)]
    cap = 1 << 20
    while True:
        out = np.zeros(cap, dtype=np.uint8)
        rc = lib().orc_video_decode(_p(a, _u8p), a.size, sb, int(motioncomp), _p(out, _u8p), cap, *[C.byref(v) for v in vals])
        W, H, F, G, M = [v.value for v in vals]
        if rc == -3:
            cap = W * H * 3 // 2 * F
            continue
        if rc < 0:
            raise ValueError(f"orc_video_decode failed: {rc}")
        return out[: W * H * 3 // 2 * F].copy(), W, H, F, G, M


# --------------------------------------------------------------------------------------------------------------
# The compiled reference (oracle/_ref), run as a subprocess.
# --------------------------------------------------------------------------------------------------------------
def ref_available(N: int = 4, huffman: bool = False) -> bool:
    return (REF_DIR / f"ref_n{N}_{'huff' if huffman else 'plain'}").exists()


def _ref_bin(N, huffman) -> str:
    p = REF_DIR / f"ref_n{N}_{'huff' if huffman else 'plain'}"
    if not p.exists():
        raise FileNotFoundError(f"{p} missing: run oracle/build_ref.sh where /root/reference exists")
    return str(p)


def _run_ref(cmd, threads=None, timeout=3600):
    env = dict(os.environ)
    if threads is not None:
        env["OMP_NUM_THREADS"] = str(threads)
    import resource

    def _limits():          # a corrupt header makes the reference allocate frames*W*H bytes: cap it instead of the box
        resource.setrlimit(resource.RLIMIT_AS, (48 << 30, 48 << 30))

    r = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=env, text=True, preexec_fn=_limits,
                       timeout=timeout)
    for line in r.stderr.splitlines():
        if line.startswith("@@RESULT "):
            return json.loads(line[len("@@RESULT "):])
    raise RuntimeError(f"reference harness failed rc={r.returncode}: {r.stderr[-2000:]}")


def _write_matrix(path, quant, N):
    q = np.asarray(quant, dtype=np.uint16).reshape(N, N)
    Path(path).write_text("\n".join(" ".join(str(int(v)) for v in row) for row in q))


def ref_image_encode(raw, W, H, N, quant, rle=True, huffman=False, threads=None, reps=1, workdir=None):
    """Runs the real reference encoder.  Returns (encoded bytes, result dict with timings)."""
    with tempfile.TemporaryDirectory(dir=workdir) as d:
        np.ascontiguousarray(raw, dtype=np.uint8).tofile(f"{d}/in.raw")
        _write_matrix(f"{d}/q.txt", quant, N)
        res = _run_ref([_ref_bin(N, huffman), "enc", f"{d}/in.raw", f"{d}/out.enc", str(W), str(H), str(int(rle)),
                        f"{d}/q.txt", str(reps)], threads)
        return Path(f"{d}/out.enc").read_bytes(), res


def ref_image_decode(enc: bytes, N, W, H, threads=None, reps=1, workdir=None):
    with tempfile.TemporaryDirectory(dir=workdir) as d:
        Path(f"{d}/in.enc").write_bytes(enc)
        res = _run_ref([_ref_bin(N, False), "dec", f"{d}/in.enc", f"{d}/out.raw", str(reps)], threads)
        return np.fromfile(f"{d}/out.raw", dtype=np.uint8).reshape(H, W), res


def ref_video_encode(yuv, W, H, quant, rle=True, gop=4, merange=16, huffman=False, threads=None, reps=1, workdir=None):
    with tempfile.TemporaryDirectory(dir=workdir) as d:
        np.ascontiguousarray(yuv, dtype=np.uint8).tofile(f"{d}/in.yuv")
        _write_matrix(f"{d}/q.txt", quant, 4)
        res = _run_ref([_ref_bin(4, huffman), "venc", f"{d}/in.yuv", f"{d}/out.enc", str(W), str(H), str(int(rle)),
                        f"{d}/q.txt", str(gop), str(merange), str(reps)], threads)
        return Path(f"{d}/out.enc").read_bytes(), res


def ref_video_decode(enc: bytes, motioncomp=True, threads=None, reps=1, workdir=None):
    with tempfile.TemporaryDirectory(dir=workdir) as d:
        Path(f"{d}/in.enc").write_bytes(enc)
        res = _run_ref([_ref_bin(4, False), "vdec", f"{d}/in.enc", f"{d}/out.yuv", str(int(motioncomp)), str(reps)], threads)
        return np.fromfile(f"{d}/out.yuv", dtype=np.uint8), res
