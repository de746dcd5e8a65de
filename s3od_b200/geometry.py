"""Host-side coefficient tables for the preprocess / postprocess kernels (vectorised numpy).

The kernels do the per-pixel arithmetic; the per-row / per-column coefficients are computed here once per image
shape, with exactly the arithmetic of the libraries the reference calls:
  * `linear_tables`  - cv2.resize INTER_LINEAR (predictor.py:81): 11-bit fixed-point taps, OpenCV's float/double mix;
  * `aa_tables`      - torch F.interpolate(..., antialias=True) (predictor.py:118-123): ATen
                       `_compute_indices_min_size_weights_aa` triangle-filter weights for scalar_t = float.
"""
from typing import Tuple

import numpy as np


def resize_mode(h: int, w: int, new_h: int, new_w: int) -> int:
    """0 = identity, 1 = exact 2x (OpenCV switches INTER_LINEAR to the 2x2 INTER_AREA box filter), 2 = general."""
    if (h, w) == (new_h, new_w):
        return 0
    if h == 2 * new_h and w == 2 * new_w:
        return 1
    return 2


def linear_tables(dst: int, src: int, vertical: bool) -> np.ndarray:
    """int32 [4, dst] = (i0, i1, c0, c1).  Horizontal taps clamp the index AND zero the fraction at the borders;
    vertical taps keep the fraction and only clamp the two row indices (cv::resize HResizeLinear / VResizeLinear)."""
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * (src / dst) - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if not vertical:
        lo = s < 0
        s = np.where(lo, 0, s)
        f = np.where(lo, np.float32(0), f)
        hi = s >= src - 1
        s = np.where(hi, src - 1, s)
        f = np.where(hi, np.float32(0), f)
    i0 = np.clip(s, 0, src - 1)
    i1 = np.clip(s + 1, 0, src - 1)
    c0 = np.rint(((np.float32(1.0) - f) * np.float32(2048)).astype(np.float32)).astype(np.int64)
    c1 = np.rint((f * np.float32(2048)).astype(np.float32)).astype(np.int64)
    return np.stack([i0, i1, c0, c1]).astype(np.int32)


def aa_tables(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray]:
    """(start int32 [out], weights float32 [out, k]) of the antialiased bilinear resize, zero-padded to k taps."""
    f32, f64 = np.float32, np.float64
    scale = f32(in_size) / f32(out_size)
    support = scale if scale >= 1.0 else f32(1.0)
    invscale = f32(f64(1.0) / f64(scale)) if scale >= 1.0 else f32(1.0)
    i = np.arange(out_size, dtype=f64)
    center = (f64(scale) * (i + 0.5)).astype(f32)
    xmin = np.maximum(np.trunc((center - support).astype(f32).astype(f64) + 0.5).astype(np.int64), 0)
    xmax = np.minimum(np.trunc((center + support).astype(f32).astype(f64) + 0.5).astype(np.int64), in_size)
    size = xmax - xmin
    k = int(size.max())
    w = np.zeros((out_size, k), f32)
    total = np.zeros(out_size, f32)
    for j in range(k):
        x = (((f32(j) + xmin.astype(f32)).astype(f32) - center).astype(f32).astype(f64) + 0.5) * f64(invscale)
        x = np.abs(x.astype(f32))
        wj = np.where(x < 1.0, f32(1.0) - x, f32(0.0)).astype(f32)
        wj = np.where(j < size, wj, f32(0.0)).astype(f32)
        w[:, j] = wj
        total = (total + wj).astype(f32)
    w = (w / total[:, None]).astype(f32)
    return xmin.astype(np.int32), np.ascontiguousarray(w)
