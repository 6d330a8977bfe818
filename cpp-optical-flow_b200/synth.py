"""Deterministic synthetic textured frame pairs with a known translational flow.

BASELINE.json configs 2-5 ("synthetic textured ... frame pair with known translational flow").
The texture is a sum of 48 plane waves, so frame1 is the *analytic* shift of frame0 (no
interpolation), and every wave is separable, sin(kx x + ky y + p) = sin(kx x + p) cos(ky y) +
cos(kx x + p) sin(ky y), which turns evaluation into one (H x 96) @ (96 x W) product - seconds
even at 16384^2.  NumPy only (no cv2), so the GPU path and the CPU oracle see identical bytes.
"""
from __future__ import annotations

import numpy as np

N_WAVES = 48


def _waves(seed: int):
    rng = np.random.default_rng(seed)
    lam = np.exp(rng.uniform(np.log(6.0), np.log(96.0), N_WAVES))   # wavelength, px
    theta = rng.uniform(0.0, 2.0 * np.pi, N_WAVES)
    phase = rng.uniform(0.0, 2.0 * np.pi, N_WAVES)
    amp = np.sqrt(lam)
    kx = 2.0 * np.pi * np.cos(theta) / lam
    ky = 2.0 * np.pi * np.sin(theta) / lam
    return kx, ky, phase, amp


def _texture(h: int, w: int, kx, ky, phase, amp, dx: float, dy: float, y0: int = 0) -> np.ndarray:
    x = np.arange(w, dtype=np.float64) - dx
    y = np.arange(y0, y0 + h, dtype=np.float64) - dy
    ax = np.outer(kx, x) + phase[:, None]                 # (48, W)
    ay = np.outer(y, ky)                                  # (H, 48)
    X = np.concatenate([np.sin(ax), np.cos(ax)], axis=0)  # (96, W)
    Y = np.concatenate([np.cos(ay) * amp, np.sin(ay) * amp], axis=1)  # (H, 96)
    return Y @ X


def frame_pair(height: int, width: int, seed: int = 1234, shift=(1.0, 0.5),
               y0: int = 0, full_height: int | None = None):
    """Return (frame0, frame1) as uint8 H x W.  frame1(x, y) = frame0(x - dx, y - dy).

    `y0`/`full_height` generate rows [y0, y0+height) of a taller image (row-slab ranks build
    only their own rows); the normalisation constant is then the analytic bound sum(amp), which
    does not depend on which rows are looked at."""
    kx, ky, phase, amp = _waves(seed)
    norm = float(np.sum(amp)) * 0.35                      # fixed, slab-independent contrast
    out = []
    for (dx, dy) in ((0.0, 0.0), shift):
        t = _texture(height, width, kx, ky, phase, amp, dx, dy, y0)
        out.append(np.clip(np.rint(127.5 + 110.0 * t / norm), 0, 255).astype(np.uint8))
    return out[0], out[1]


def video_pair(index: int, height: int = 1080, width: int = 1920):
    """Pair `index` of the synthetic video batch (BASELINE config 4)."""
    dx = 0.25 * (1 + index % 8)
    dy = -1.0 + 0.25 * (index % 9)
    return frame_pair(height, width, seed=1000 + index, shift=(dx, dy))
