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


def textured_rgba8(w: int, h: int, seed: int = 1, amp: float = 0.04, cell: int = 96) -> np.ndarray:
    """(T) the natural scene with texel-scale detail (2x2-box-filtered noise, +-amp) on a checkerboard of cell x cell
    regions: about a third of the texels carry a NIS edge (natural: 1-2 %, uniform: 84 %), so roughly a third of the
    32-pixel output segments are edge-free -- content between the two extremes for the data-dependent NIS paths."""
    rng = np.random.default_rng(seed + 1000)
    n = rng.uniform(-1, 1, (h + 1, w + 1)).astype(np.float32)
    n = (n[:-1, :-1] + n[1:, :-1] + n[:-1, 1:] + n[1:, 1:]) * 0.5
    yy, xx = np.mgrid[0:h, 0:w]
    mask = (((yy // cell) + (xx // cell)) % 2 == 0).astype(np.float32)
    img = np.clip(natural_f32(w, h, seed) + (amp * n * mask)[..., None], 0.0, 1.0)
    out = np.empty((h, w, 4), np.uint8)
    out[..., :3] = (img * 255.0 + 0.5).astype(np.uint8)
    out[..., 3] = 255
    return out


def natural_rgba16f(w: int, h: int, seed: int = 1, peak: float = 4.0) -> np.ndarray:
    rgb = natural_f32(w, h, seed) * peak
    out = np.ones((h, w, 4), np.float16)
    out[..., :3] = rgb.astype(np.float16)
    return out


def stereo_pair(kind: str, w: int, h: int, seed: int = 0):
    """(left, right): the right eye is the left eye rolled by 16 px so the eyes differ."""
    gen = {"uniform": uniform_rgba8, "natural": natural_rgba8, "textured": textured_rgba8, "natural16f": natural_rgba16f}[kind]
    left = gen(w, h, seed)
    return left, np.ascontiguousarray(np.roll(left, 16, axis=1))


def pack_rgb10a2(rgb10: np.ndarray, a2: np.ndarray) -> np.ndarray:
    """(H, W, 3) values 0..1023 and (H, W) values 0..3 -> (H, W, 4) uint8 = the bytes of the little-endian
    R10G10B10A2 word per texel (R bits 0-9, G 10-19, B 20-29, A 30-31)."""
    w = (rgb10[..., 0].astype(np.uint32) | (rgb10[..., 1].astype(np.uint32) << 10) |
         (rgb10[..., 2].astype(np.uint32) << 20) | (a2.astype(np.uint32) << 30))
    return np.ascontiguousarray(w.astype("<u4")).view(np.uint8).reshape(w.shape[0], w.shape[1], 4)


def unpack_rgb10a2(img: np.ndarray) -> np.ndarray:
    """(H, W, 4) uint8 bytes of R10G10B10A2 words -> (H, W, 4) int32 channel codes (r, g, b, a)."""
    w = np.ascontiguousarray(img).view("<u4").reshape(img.shape[0], img.shape[1]).astype(np.uint32)
    return np.stack([w & 1023, (w >> 10) & 1023, (w >> 20) & 1023, w >> 30], axis=-1).astype(np.int32)


def natural_rgb10a2(width: int, height: int, seed: int) -> np.ndarray:
    """The natural_rgba16f scene quantised to R10G10B10A2_UNORM (as bytes, see pack_rgb10a2)."""
    f = np.clip(natural_rgba16f(width, height, seed).astype(np.float32), 0.0, 1.0)
    rng = np.random.default_rng(seed + 77)
    rgb = np.floor(f[..., :3] * 1023.0 + 0.5).astype(np.uint32)
    return pack_rgb10a2(rgb, rng.integers(0, 4, size=(height, width)))


def uniform_rgb10a2(width: int, height: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return pack_rgb10a2(rng.integers(0, 1024, size=(height, width, 3)), rng.integers(0, 4, size=(height, width)))
