"""Empirical check of the fast inverse transform's error bound (decode_image.cu: |X_fast - X_real| <= 14 u S claimed, delta
uses a multiple of u S): emulates the FP32 operation DAG of idct8_inplace / idct4_inplace (transform_fast.cuh) in numpy and
compares with binary64 on random, sparse, saturated and sign-aligned coefficient blocks.  Prints max |err| / (u S)."""
import numpy as np
f32, f64 = np.float32, np.float64
u = 2.0 ** -24

def fma(a, b, c):
    return (a.astype(f64) * f64(b) + c.astype(f64)).astype(f32)
def mul(a, b):
    return (a.astype(f64) * f64(b)).astype(f32)

def consts(N):
    k = np.arange(16)
    c = np.cos(k * np.pi / 16).astype(f32)
    return c

C = consts(8)
C1, C2, C3, C4, C5, C6, C7 = [f32(C[i]) for i in range(1, 8)]

def idct8(v):          # v: (..., 8) float32, transform along the last axis
    y0, y1, y2, y3, y4, y5, y6, y7 = [v[..., i] for i in range(8)]
    t4 = mul(y4, C4)
    a, b = y0 + t4, y0 - t4
    p = fma(y2, C2, mul(y6, C6)); q = fma(y2, C6, -mul(y6, C2))
    e0, e3, e1, e2 = a + p, a - p, b + q, b - q
    o0 = fma(y1, C1, fma(y3, C3, fma(y5, C5, mul(y7, C7))))
    o1 = fma(y1, C3, fma(y3, -C7, fma(y5, -C1, mul(y7, -C5))))
    o2 = fma(y1, C5, fma(y3, -C1, fma(y5, C7, mul(y7, C3))))
    o3 = fma(y1, C7, fma(y3, -C5, fma(y5, C3, mul(y7, -C1))))
    return np.stack([e0 + o0, e1 + o1, e2 + o2, e3 + o3, e3 - o3, e2 - o2, e1 - o1, e0 - o0], axis=-1)

def idct4(v):
    y0, y1, y2, y3 = [v[..., i] for i in range(4)]
    t2 = mul(y2, C4)
    a, b = y0 + t2, y0 - t2
    p = fma(y1, C2, mul(y3, C6)); q = fma(y1, C6, -mul(y3, C2))
    return np.stack([a + p, b + q, b - q, a - p], axis=-1)

def run(N, blocks):
    x = blocks.astype(f32)                                  # x[u][v]
    one = idct8 if N == 8 else idct4
    t = one(x)                                              # rows: along v
    t = np.swapaxes(one(np.swapaxes(t, -1, -2)), -1, -2)    # columns: along u
    i = np.arange(N)
    cs = np.cos((2 * i[:, None] + 1) * np.arange(N)[None, :] * np.pi / (2 * N))      # cs[i][u], binary64
    exact = np.einsum('iu,buv,jv->bij', cs, x.astype(f64), cs)
    S = np.abs(x.astype(f64)).sum(axis=(1, 2))
    ratio = np.abs(t.astype(f64) - exact).max(axis=(1, 2)) / (u * np.maximum(S, 1e-30))
    return ratio.max()

rng = np.random.default_rng(1)
for N in (8, 4):
    M = 200000
    worst = 0.0
    for name, gen in (("gauss", lambda: rng.normal(0, 100, (M, N, N))),
                      ("sparse", lambda: rng.normal(0, 300, (M, N, N)) * (rng.random((M, N, N)) < 0.2)),
                      ("one", lambda: np.eye(N * N)[rng.integers(0, N * N, M)].reshape(M, N, N) * rng.normal(0, 1000, (M, 1, 1))),
                      ("aligned", lambda: np.abs(rng.normal(0, 100, (M, N, N))) * np.sign(np.cos((2 * rng.integers(0, N) + 1) * np.arange(N)[:, None] * np.pi / (2 * N)) * np.cos((2 * rng.integers(0, N) + 1) * np.arange(N)[None, :] * np.pi / (2 * N)) + 1e-9)),
                      ("ints", lambda: rng.integers(-2000, 2000, (M, N, N)).astype(f64))):
        r = run(N, gen())
        worst = max(worst, r)
        print(f"N={N} {name:8s} max |err| / (u S) = {r:.3f}")
    print(f"N={N}: worst {worst:.3f}  (bound claimed 14, delta uses 32)")
