"""Seeded synthetic eye textures: the bench and the parity tests draw from here.

(U) uniform random bytes -- adversarial for the direction logic;
(N) "natural-like": low-frequency sinusoids + hard-edged rectangles/discs + a black border band (the
    hidden-area mesh of a VR frame), which exercises edges, flats and RCAS's 0*inf path.
"""
from __future__ import annotations

import numpy as np


def uniform_rgba8(w: int, h: int, seed: int = 0) -> np.ndarray:
    img = np.random.default_rng(seed).integers(0, 256, (h, w, 4), dtype=np.uint8)
    img[..., 3] = 255
    return img


def natural_f32(w: int, h: int, seed: int = 1) -> np.ndarray:
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.zeros((h, w, 3), np.float32)
    for c in range(3):
        for _ in range(6):
            fx, fy = rng.uniform(0.5, 6.0, 2) * 2 * np.pi / np.array([w, h], np.float32)
            img[..., c] += rng.uniform(0.05, 0.2) * np.sin(fx * xx + fy * yy + rng.uniform(0, 6.28))
        img[..., c] += 0.5
    for _ in range(32):
        x0, y0 = rng.integers(0, w), rng.integers(0, h)
        col = rng.uniform(0, 1, 3).astype(np.float32)
        if rng.random() < 0.5:
            ww, hh = rng.integers(2, max(3, w // 4)), rng.integers(2, max(3, h // 4))
            img[y0:y0 + hh, x0:x0 + ww] = col
        else:
            r = rng.integers(2, max(3, min(w, h) // 6))
            img[(xx - x0) ** 2 + (yy - y0) ** 2 <= r * r] = col
    bx, by = max(1, int(0.02 * w)), max(1, int(0.02 * h))
    img[:by] = 0; img[-by:] = 0; img[:, :bx] = 0; img[:, -bx:] = 0
    return np.clip(img, 0.0, 1.0)


def natural_rgba8(w: int, h: int, seed: int = 1) -> np.ndarray:
    rgb = natural_f32(w, h, seed)
    out = np.empty((h, w, 4), np.uint8)
    out[..., :3] = (rgb * 255.0 + 0.5).astype(np.uint8)
    out[..., 3] = 255
    return out


def natural_rgba16f(w: int, h: int, seed: int = 1, peak: float = 4.0) -> np.ndarray:
    rgb = natural_f32(w, h, seed) * peak
    out = np.ones((h, w, 4), np.float16)
    out[..., :3] = rgb.astype(np.float16)
    return out


def stereo_pair(kind: str, w: int, h: int, seed: int = 0):
    """(left, right): the right eye is the left eye rolled by 16 px so the eyes differ."""
    gen = {"uniform": uniform_rgba8, "natural": natural_rgba8, "natural16f": natural_rgba16f}[kind]
    left = gen(w, h, seed)
    return left, np.ascontiguousarray(np.roll(left, 16, axis=1))
