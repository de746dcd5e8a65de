"""CPU oracle of the training-side `SODPredictor.predict` (TEST INFRASTRUCTURE - see oracle/__init__.py).

Follows /root/reference/synth_sod/src/synth_sod/model_training/predictor.py:330-477 line by line with the restated
pieces of oracle/model.py and oracle/prepost.py.

PARITY UNPINNED for the front end: the reference module cannot be imported here (`albumentations`, `hydra` and the FLUX
pipeline it imports at module level are not installed, and `albumentations` is not in /opt/wheelhouse), so its three
transforms are restated from albumentations' published functional code (1.4.x `_func_max_size` / `py3round`,
`PadIfNeeded.update_params`, `normalize`):
  * LongestMaxSize: scale = max_size / float(max(w, h)); if scale != 1.0: (new_h, new_w) = round-half-even(dim * scale),
    cv2.resize(..., interpolation=cv2.INTER_LINEAR)            [predictor.py:347]
  * PadIfNeeded(position=center, BORDER_CONSTANT, fill 0): top = int(pad / 2.0), bottom = pad - top   [predictor.py:348-353]
  * Normalize: float32  img -= mean * 255 ; img *= 1 / (std * 255)                                    [predictor.py:355]
Everything behind the transform (model forward, sigmoid, remove_padding, antialiased resize, argmax, threshold) is the
arithmetic pinned by tests/golden for `remove_background`.
"""
from typing import Any, Dict

import numpy as np
import torch

from . import model as _model
from . import prepost as _pp


def get_pad_info(h: int, w: int, image_size: int) -> Dict[str, Any]:
    """predictor.py:374-398"""
    aspect_ratio = w / h
    if aspect_ratio > 1:
        new_w = image_size
        new_h = int(new_w / aspect_ratio)
        return {"height_pad": (image_size - new_h) // 2, "width_pad": 0, "original_size": (h, w), "resized_size": (new_h, new_w)}
    new_h = image_size
    new_w = int(new_h * aspect_ratio)
    return {"height_pad": 0, "width_pad": (image_size - new_w) // 2, "original_size": (h, w), "resized_size": (new_h, new_w)}


def transform(image: np.ndarray, image_size: int) -> np.ndarray:
    """predictor.py:345-356 -> float32 (S, S, 3)."""
    h, w = image.shape[:2]
    scale = image_size / float(max(w, h))
    if scale != 1.0:
        new_h, new_w = (int(round(dim * scale)) for dim in (h, w))
        image = _pp.cv_resize_linear_u8(image, new_h, new_w)
    rows, cols = image.shape[:2]
    top = int((image_size - rows) / 2.0) if rows < image_size else 0
    left = int((image_size - cols) / 2.0) if cols < image_size else 0
    canvas = np.zeros((max(image_size, rows), max(image_size, cols), 3), np.uint8)
    canvas[top:top + rows, left:left + cols] = image
    f32 = np.float32
    mean = np.array([0.485, 0.456, 0.406], dtype=f32) * f32(255.0)
    den = np.reciprocal(np.array([0.229, 0.224, 0.225], dtype=f32) * f32(255.0), dtype=f32)
    img = canvas.astype(f32)
    img -= mean
    img *= den
    return img


def predict(sd, image: np.ndarray, arch, image_size: int = 840, threshold: float = 0.5) -> Dict[str, Any]:
    """predictor.py:408-477; returns the PredictionResult fields as a dict (all_masks / all_ious None for one output)."""
    h, w = image.shape[:2]
    pad_info = get_pad_info(h, w, image_size)
    x = torch.from_numpy(np.ascontiguousarray(transform(image, image_size).transpose(2, 0, 1)[None]))
    out = _model.forward(sd, x, arch)
    probs = torch.sigmoid(out["pred_masks"])[0].numpy()
    ious = torch.sigmoid(out["pred_iou"])[0].numpy()
    hp, wp = pad_info["height_pad"], pad_info["width_pad"]
    if hp > 0:
        probs = probs[:, hp:-hp, :]
    if wp > 0:
        probs = probs[:, :, wp:-wp]
    resized = _pp.aa_resize(np.ascontiguousarray(probs), h, w)
    if resized.shape[0] == 1:
        soft = resized[0]
        return dict(binary_mask=(soft > threshold).astype(np.float32), soft_mask=soft, all_masks=None, all_ious=None)
    best = int(ious.argmax())
    soft = resized[best]
    return dict(binary_mask=(soft > threshold).astype(np.float32), soft_mask=soft,
                all_masks=(resized > threshold).astype(np.float32), all_ious=ious, best_idx=best)
