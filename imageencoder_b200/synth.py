"""Deterministic synthetic inputs for the parity tests and bench.py.

Integer-only after three tiny float64 look-up tables are quantised to fixed point, so the same seed gives the same
bytes on every machine (no dependence on SIMD libm variants): smooth low-frequency content + texture + bounded noise,
roughly natural-image statistics for the block transform, plus an optional flat rectangle (all-zero quantised blocks,
SURVEY 0.4).  Rows are generated in independent 256-row stripes seeded by (seed, stripe), so a rank can generate only
its own shard of a large image.
"""
from __future__ import annotations

import math

import numpy as np

_STRIPE = 256
_FP = 14


def _lut(period: int, fn) -> np.ndarray:
    return np.array([int(round(fn(2.0 * math.pi * k / period) * (1 << _FP))) for k in range(period)], dtype=np.int64)


_SX = _lut(257, math.sin)
_CY = _lut(193, math.cos)
_S61 = _lut(61, math.sin)


def synth_rows(width: int, y0: int, y1: int, seed: int, noise: int = 6, x0: int = 0) -> np.ndarray:
    """Rows [y0, y1) of the infinite synthetic image, columns [x0, x0+width).  uint8 array (y1-y0, width)."""
    out = np.empty((y1 - y0, width), dtype=np.uint8)
    xs = np.arange(x0, x0 + width, dtype=np.int64)
    sx = _SX[xs % 257]
    y = y0
    while y < y1:
        stripe = y // _STRIPE
        ye = min(y1, (stripe + 1) * _STRIPE)
        ys = np.arange(y, ye, dtype=np.int64)
        smooth = 56 * (sx[None, :] * _CY[ys % 193][:, None])                      # 2*_FP fractional bits
        tex = 28 * (_S61[(xs[None, :] + 2 * ys[:, None]) % 61] << _FP)
        v = (128 << (2 * _FP)) + smooth + tex
        px = (v + (1 << (2 * _FP - 1))) >> (2 * _FP)
        if noise:
            # sum of three uniforms on [-noise, noise]: sigma ~ 1.08*noise, drawn for the whole stripe so that any
            # row range of the stripe sees the same values
            rng = np.random.default_rng([seed, stripe])
            n = rng.integers(-noise, noise + 1, size=(3, _STRIPE, x0 + width), dtype=np.int8).astype(np.int64).sum(axis=0)
            px = px + n[y - stripe * _STRIPE: ye - stripe * _STRIPE, x0:x0 + width]
        out[y - y0: ye - y0] = np.clip(px, 0, 255).astype(np.uint8)
        y = ye
    return out


def synth_image(width: int, height: int, seed: int, flat: bool = False, noise: int = 6) -> np.ndarray:
    img = synth_rows(width, 0, height, seed, noise)
    if flat:
        img[height // 4: height // 4 + height // 2, width // 4: width // 4 + width // 2] = 128
    return img


def synth_video(width: int, height: int, frames: int, seed: int = 4000) -> np.ndarray:
    """YUV420 planar clip: frame t is the (height x width) crop at (40+t, 40+2t) of one synthetic image plus small
    per-frame noise; UV = 0x80.  Returns a flat uint8 array of frames*width*height*3/2 bytes."""
    big = synth_rows(width + 2 * frames + 80, 0, height + frames + 80, seed, noise=6).astype(np.int16)
    fsz = width * height * 3 // 2
    out = np.full(frames * fsz, 0x80, dtype=np.uint8)
    for t in range(frames):
        rng = np.random.default_rng([seed, 1000003 + t])
        n = rng.integers(-2, 3, size=(2, height, width), dtype=np.int8).astype(np.int16).sum(axis=0)
        y = np.clip(big[40 + t: 40 + t + height, 40 + 2 * t: 40 + 2 * t + width] + n, 0, 255).astype(np.uint8)
        out[t * fsz: t * fsz + width * height] = y.reshape(-1)
    return out
