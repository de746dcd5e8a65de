"""CPU restatement of the pre- and post-processing around the model (TEST INFRASTRUCTURE - see oracle/__init__.py).

Follows:
  * get_pad_info / remove_padding            /root/reference/src/s3od/utils.py:6-29, 32-37
  * BackgroundRemoval._preprocess            /root/reference/src/s3od/predictor.py:79-94
  * post-processing in remove_background     /root/reference/src/s3od/predictor.py:113-132
Third-party arithmetic restated here (absent from /root/reference, so restated from the published
algorithms and checked against the installed libraries in tests/test_oracle_prepost.py):
  * cv2.resize INTER_LINEAR on uint8 (OpenCV 4.13.0 installed; uv.lock:2875 pins 4.12.0.88):
    11-bit fixed-point separable bilinear (INTER_RESIZE_COEF_BITS=11), vertical pass
    ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2 >> 2; exact 2x down-scale takes the INTER_AREA
    2x2 fast path (a+b+c+d+2)>>2.
  * torch F.interpolate(bilinear, align_corners=False, antialias=True) (torch 2.11 installed;
    ATen UpSampleKernel `_compute_indices_min_size_weights_aa`): separable triangle filter with
    support = max(scale, 1), weights normalised per output index.
"""
import math
from typing import Any, Dict, Tuple

import numpy as np

MEAN = np.array([0.485, 0.456, 0.406])
STD = np.array([0.229, 0.224, 0.225])


def get_pad_info(h: int, w: int, image_size: int = 1024) -> Dict[str, Any]:
    """utils.py:6-29 - long side -> image_size, int() truncation, symmetric floor padding."""
    aspect = w / h
    if aspect > 1:
        new_w = image_size
        new_h = int(new_w / aspect)
        return dict(height_pad=(image_size - new_h) // 2, width_pad=0, original_size=(h, w), resized_size=(new_h, new_w))
    new_h = image_size
    new_w = int(new_h * aspect)
    return dict(height_pad=0, width_pad=(image_size - new_w) // 2, original_size=(h, w), resized_size=(new_h, new_w))


def _linear_coeffs(dst: int, src: int, vertical: bool):
    """OpenCV resize INTER_LINEAR index/coefficient tables (cv::resize, `fx = (dx+0.5)*scale-0.5`).

    Horizontal: at the borders the tap index is clamped AND the fraction is zeroed.
    Vertical  : the fraction is kept and only the two row indices are clamped (both rows then alias the
                border row and each still goes through its own >>16 truncation).
    Returns (i0, i1, c0, c1): tap indices and 11-bit coefficients."""
    scale = src / dst
    i0 = np.zeros(dst, np.int64)
    i1 = np.zeros(dst, np.int64)
    c0 = np.zeros(dst, np.int64)
    c1 = np.zeros(dst, np.int64)
    for d in range(dst):
        f = np.float32((d + 0.5) * scale - 0.5)     # double scale, float fx as in OpenCV
        s = int(math.floor(f))
        f = np.float32(f - np.float32(s))
        if not vertical:
            if s < 0:
                s, f = 0, np.float32(0)
            if s >= src - 1:
                s, f = src - 1, np.float32(0)
        i0[d] = min(max(s, 0), src - 1)
        i1[d] = min(max(s + 1, 0), src - 1)
        # saturate_cast<short>(x * INTER_RESIZE_COEF_SCALE) == cvRound (round-half-even)
        c0[d] = int(np.rint(np.float32((np.float32(1.0) - f) * np.float32(2048))))
        c1[d] = int(np.rint(np.float32(f * np.float32(2048))))
    return i0, i1, c0, c1


def cv_resize_linear_u8(img: np.ndarray, new_h: int, new_w: int) -> np.ndarray:
    """cv2.resize(img, (new_w, new_h)) with the default INTER_LINEAR, uint8 HWC."""
    h, w = img.shape[:2]
    if (new_h, new_w) == (h, w):
        return img.copy()
    if h == 2 * new_h and w == 2 * new_w:
        s = img.astype(np.int32)
        return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    x0, x1, xa0, xa1 = _linear_coeffs(new_w, w, vertical=False)
    y0, y1, yb0, yb1 = _linear_coeffs(new_h, h, vertical=True)
    s = img.astype(np.int64)
    rows = s[:, x0] * xa0[None, :, None] + s[:, x1] * xa1[None, :, None]          # (h, new_w, 3) horizontal pass
    r0, r1 = rows[y0], rows[y1]
    out = (((yb0[:, None, None] * (r0 >> 4)) >> 16) + ((yb1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def letterbox_u8(img: np.ndarray, image_size: int) -> Tuple[np.ndarray, Dict[str, Any]]:
    """predictor.py:79-89: resize + paste on a black canvas.  Raises like the reference for odd padding."""
    pad = get_pad_info(img.shape[0], img.shape[1], image_size)
    new_h, new_w = pad["resized_size"]
    resized = cv_resize_linear_u8(img, new_h, new_w)
    canvas = np.zeros((image_size, image_size, 3), np.uint8)
    hp, wp = pad["height_pad"], pad["width_pad"]
    if hp > 0:
        if image_size - 2 * hp != new_h:
            raise ValueError(f"could not broadcast input array from shape {resized.shape} into shape "
                             f"{(image_size - 2 * hp, image_size, 3)}")
        canvas[hp:image_size - hp, :] = resized
    elif wp > 0:
        if image_size - 2 * wp != new_w:
            raise ValueError(f"could not broadcast input array from shape {resized.shape} into shape "
                             f"{(image_size, image_size - 2 * wp, 3)}")
        canvas[:, wp:image_size - wp] = resized
    else:
        canvas = resized
    return canvas, pad


def preprocess(img: np.ndarray, image_size: int = 1024):
    """predictor.py:79-94 -> float32 (1,3,S,S) numpy + pad_info.  Normalise in float64 then cast, like the reference."""
    canvas, pad = letterbox_u8(img, image_size)
    norm = (canvas.astype(np.float32) / 255.0 - MEAN) / STD
    return np.ascontiguousarray(norm.transpose(2, 0, 1)[None]).astype(np.float32), pad


def _aa_weights(in_size: int, out_size: int):
    """ATen `_compute_indices_min_size_weights_aa` (UpSampleKernel.cpp) for the triangle filter, align_corners=False.

    Restated with the same mixed float/double arithmetic the C++ performs for scalar_t=float:
    scale, support, invscale and center are fp32; `x + 0.5` promotes to double; int - float stays float.
    """
    f32, f64 = np.float32, np.float64
    scale = f32(in_size) / f32(out_size)
    support = scale if scale >= 1.0 else f32(1.0)
    invscale = f32(f64(1.0) / f64(scale)) if scale >= 1.0 else f32(1.0)
    starts, weights = [], []
    for i in range(out_size):
        center = f32(f64(scale) * (i + 0.5))
        xmin = max(int(f64(f32(center - support)) + 0.5), 0)
        xmax = min(int(f64(f32(center + support)) + 0.5), in_size)
        ws = []
        for j in range(xmax - xmin):
            x = f32((f64(f32(f32(j + xmin) - center)) + 0.5) * f64(invscale))
            x = abs(x)
            ws.append(f32(1.0) - x if x < 1.0 else f32(0.0))
        ws = np.array(ws, f32)
        total = f32(0.0)
        for wv in ws:
            total = f32(total + wv)
        ws = (ws / total).astype(f32)
        starts.append(xmin)
        weights.append(ws)
    return starts, weights


def aa_resize(planes: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """F.interpolate(x, size, mode='bilinear', align_corners=False, antialias=True) on (C,h,w) fp32."""
    c, h, w = planes.shape
    xs, xw = _aa_weights(w, out_w)
    ys, yw = _aa_weights(h, out_h)
    tmp = np.empty((c, h, out_w), np.float32)
    for i in range(out_w):
        tmp[:, :, i] = (planes[:, :, xs[i]:xs[i] + len(xw[i])] * xw[i][None, None, :]).sum(-1, dtype=np.float32)
    out = np.empty((c, out_h, out_w), np.float32)
    for i in range(out_h):
        out[:, i, :] = (tmp[:, ys[i]:ys[i] + len(yw[i]), :] * yw[i][None, :, None]).sum(1, dtype=np.float32)
    return out


def postprocess(mask_logits: np.ndarray, iou_logits: np.ndarray, pad: Dict[str, Any], image: np.ndarray):
    """predictor.py:113-132 for one image.  mask_logits (K,S,S) fp32, iou_logits (K,), image (H,W,3) u8."""
    probs = (1.0 / (1.0 + np.exp(-mask_logits.astype(np.float32)))).astype(np.float32)
    ious = (1.0 / (1.0 + np.exp(-iou_logits.astype(np.float32)))).astype(np.float32)
    S = probs.shape[-1]
    hp, wp = pad["height_pad"], pad["width_pad"]
    if hp > 0:
        probs = probs[:, hp:S - hp, :]
    if wp > 0:
        probs = probs[:, :, wp:S - wp]
    H, W = pad["original_size"]
    all_masks = aa_resize(np.ascontiguousarray(probs), H, W)
    best = int(ious.argmax())
    predicted = all_masks[best]
    alpha = (predicted * 255).astype(np.uint8)          # truncation, predictor.py:130
    rgba = np.dstack([image, alpha])
    return dict(predicted_mask=predicted, all_masks=all_masks, all_ious=ious, rgba=rgba, best_idx=best)
