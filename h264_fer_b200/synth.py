"""Deterministic synthetic YUV 4:2:0 clips and the Y4M container the reference parses.

The reference's Y4M reader only understands ` W<w>` / ` H<h>` header tokens and parameter-less ``FRAME``
lines (reference fileIO.cpp:195-252) and centre-crops to multiples of 16 (fileIO.cpp:242-243,290-339).

Content: a (twice 5x5 box-)blurred random 8-px-grid texture panning 2 px/frame right and 1 px/frame down, one 32x32 inverted
square moving 5 px/frame, N(0,1) noise per frame, luma clipped to [16,235] — which keeps every 8x8 window sum
away from 0 and from >= 16,203, the two inputs for which the reference's sum-sorted index is undefined
behaviour (moestimation.cpp:153-158,477-480).
"""
from __future__ import annotations

import numpy as np


def _box5(a: np.ndarray) -> np.ndarray:
    """Separable 5x5 box blur with edge replication, integer arithmetic."""
    p = np.pad(a.astype(np.int32), ((2, 2), (2, 2)), mode="edge")
    h = p[:, 0:-4] + p[:, 1:-3] + p[:, 2:-2] + p[:, 3:-1] + p[:, 4:]
    v = h[0:-4] + h[1:-3] + h[2:-2] + h[3:-1] + h[4:]
    return (v + 12) // 25


class SynthClip:
    """Frame generator: ``clip.frame(t) -> (Y, Cb, Cr)`` uint8 arrays of the *input* size (before crop)."""

    def __init__(self, width: int, height: int, seed: int, pan=(2, 1), noise=1.0, square=True, contrast=1.0):
        self.w, self.h, self.seed = int(width), int(height), int(seed)
        self.pan, self.noise, self.square, self.contrast = pan, float(noise), bool(square), float(contrast)
        rng = np.random.default_rng(seed)
        gh, gw = self.h // 8 + 40, self.w // 8 + 40
        grid = rng.integers(40, 200, size=(gh, gw), dtype=np.int32)
        self.canvas = _box5(_box5(np.kron(grid, np.ones((8, 8), dtype=np.int32))))

    def frame(self, t: int):
        w, h = self.w, self.h
        ch, cw = self.canvas.shape
        oy, ox = 20 + self.pan[1] * t, 20 + self.pan[0] * t
        rows = (np.arange(h) + oy) % ch
        cols = (np.arange(w) + ox) % cw
        y = self.canvas[np.ix_(rows, cols)].astype(np.float64)
        if self.contrast != 1.0:
            y = 128.0 + (y - 128.0) * self.contrast
        # moving inverted square
        sx = (16 + 5 * t) % max(1, w - 32)
        sy = (h // 3 + 3 * t) % max(1, h - 32)
        if self.square:
            y[sy:sy + 32, sx:sx + 32] = 255.0 - y[sy:sy + 32, sx:sx + 32]
        if self.noise > 0:
            rng = np.random.default_rng([self.seed, 7919, t])
            y = y + self.noise * rng.standard_normal(size=y.shape)
        yi = np.clip(np.rint(y), 16, 235).astype(np.uint8)
        y2 = yi[0::2, 0::2].astype(np.int32)
        cb = np.clip(128 + (y2 - 128) // 4, 16, 240).astype(np.uint8)
        cr = np.clip(128 - (y2 - 128) // 4, 16, 240).astype(np.uint8)
        return yi, cb, cr


def y4m_header(width: int, height: int) -> bytes:
    return ("YUV4MPEG2 W%d H%d F30:1 Ip A1:1 C420jpeg\n" % (width, height)).encode()


def write_y4m(path: str, width: int, height: int, seed: int, frames: int, **kw) -> None:
    clip = SynthClip(width, height, seed, **kw)
    with open(path, "wb") as f:
        f.write(y4m_header(width, height))
        for t in range(frames):
            y, cb, cr = clip.frame(t)
            f.write(b"FRAME\n")
            f.write(y.tobytes()); f.write(cb.tobytes()); f.write(cr.tobytes())


def crop16(plane: np.ndarray, chroma: bool = False) -> np.ndarray:
    """Centre-crop exactly as reference fileIO.cpp:242-243,290-339 (luma offsets halved for chroma)."""
    h, w = plane.shape
    if chroma:
        fh, fw = (2 * h) & ~15, (2 * w) & ~15
        top, left = ((2 * h - fh) >> 1) >> 1, ((2 * w - fw) >> 1) >> 1
        return np.ascontiguousarray(plane[top:top + fh // 2, left:left + fw // 2])
    fh, fw = h & ~15, w & ~15
    top, left = (h - fh) >> 1, (w - fw) >> 1
    return np.ascontiguousarray(plane[top:top + fh, left:left + fw])
