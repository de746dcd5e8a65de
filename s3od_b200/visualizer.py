"""Drop-in for `s3od.visualizer` (/root/reference/src/s3od/visualizer.py:8-48) and for the mask-ambiguity check of the demo
(/root/reference/demo/app.py:38-56), computed by the CUDA library.

Same names, arguments and return types as the reference: `visualize_removal(image, result, background_color)` and
`visualize_all_masks(image, result)` return PIL images; `compute_mask_iou` / `is_ambiguous` return a float / bool.
`image` and the mask fields of `result` may be numpy arrays (copied to the GPU) or CUDA tensors already on the device
(the `*_device` functions below take and return CUDA tensors and never touch the host).  Outputs are bit-identical to
the reference's numpy arithmetic.  There is no CPU fallback.
"""
from typing import Tuple, Union

import numpy as np
import torch
from PIL import Image

from .engine import _check, _stream_ptr, load_library

ArrayLike = Union[np.ndarray, torch.Tensor]


def _dev(t: ArrayLike, dtype: torch.dtype, device=None) -> torch.Tensor:
    if isinstance(t, torch.Tensor):
        if not t.is_cuda:
            t = t.to(device or "cuda")
        return t.to(dtype).contiguous()
    return torch.from_numpy(np.ascontiguousarray(t)).to(device or "cuda").to(dtype).contiguous()


def _image_u8(image) -> ArrayLike:
    if isinstance(image, Image.Image):
        return np.array(image.convert("RGB"))                    # visualizer.py:13-14
    return image


def composite_device(d_image: torch.Tensor, d_mask: torch.Tensor, background_color: Tuple[int, int, int]) -> torch.Tensor:
    """(H,W,3) u8, (H,W) f32 CUDA tensors -> (H,W,3) u8 CUDA tensor: trunc(mask * image + (1 - mask) * background)."""
    lib = load_library()
    h, w = d_mask.shape
    if tuple(d_image.shape) != (h, w, 3):
        raise ValueError(f"image {tuple(d_image.shape)} does not match mask {(h, w)}")
    out = torch.empty((h, w, 3), dtype=torch.uint8, device=d_image.device)
    r, g, b = (int(c) for c in background_color)
    with torch.cuda.device(d_image.device):
        _check(lib, lib.s3od_vis_composite(d_image.data_ptr(), d_mask.data_ptr(), out.data_ptr(), h, w, r, g, b,
                                           _stream_ptr(d_image.device)), "s3od_vis_composite")
    return out


def mask_grid_device(d_image: torch.Tensor, d_masks: torch.Tensor) -> torch.Tensor:
    """(H,W,3) u8, (K,H,W) f32 -> (ceil(K/4)*H, min(K,4)*W, 3) u8 grid of trunc(mask_k * image) cells."""
    lib = load_library()
    k, h, w = d_masks.shape
    if tuple(d_image.shape) != (h, w, 3):
        raise ValueError(f"image {tuple(d_image.shape)} does not match masks {(k, h, w)}")
    gw = min(k, 4)
    gh = (k + gw - 1) // gw
    out = torch.zeros((h * gh, w * gw, 3), dtype=torch.uint8, device=d_image.device)      # visualizer.py:38
    with torch.cuda.device(d_image.device):
        _check(lib, lib.s3od_vis_mask_grid(d_image.data_ptr(), d_masks.data_ptr(), k, out.data_ptr(), h, w,
                                           _stream_ptr(d_image.device)), "s3od_vis_mask_grid")
    return out


def pair_counts_device(d_masks: torch.Tensor) -> torch.Tensor:
    """(K,H,W) f32, K <= 4 -> int64 CUDA tensor [K(K-1)/2, 2]: (intersection, union) pixel counts of the 0.5-thresholded pairs."""
    lib = load_library()
    k, h, w = d_masks.shape
    counts = torch.zeros((max(k * (k - 1) // 2, 1), 2), dtype=torch.int64, device=d_masks.device)
    if k >= 2:
        with torch.cuda.device(d_masks.device):
            _check(lib, lib.s3od_mask_pair_counts(d_masks.data_ptr(), k, h, w, counts.data_ptr(), _stream_ptr(d_masks.device)),
                   "s3od_mask_pair_counts")
    return counts[:k * (k - 1) // 2]


# ---------------------------------------------------------------------------------------------------------------------
# the reference's surface
def visualize_removal(image, result, background_color: Tuple[int, int, int] = (0, 255, 0)) -> Image.Image:
    d_img = _dev(_image_u8(image), torch.uint8)
    d_mask = _dev(result.predicted_mask, torch.float32, d_img.device)
    return Image.fromarray(composite_device(d_img, d_mask, background_color).cpu().numpy())


def visualize_all_masks(image, result) -> Image.Image:
    d_img = _dev(_image_u8(image), torch.uint8)
    d_masks = _dev(result.all_masks, torch.float32, d_img.device)
    return Image.fromarray(mask_grid_device(d_img, d_masks).cpu().numpy())


def compute_mask_iou(mask1: ArrayLike, mask2: ArrayLike) -> float:
    d = torch.stack([_dev(mask1, torch.float32), _dev(mask2, torch.float32)])
    inter, union = (int(v) for v in pair_counts_device(d)[0].cpu())
    return inter / (union + 1e-6)                                 # app.py:42, float64 like numpy


def is_ambiguous(all_masks: ArrayLike, threshold: float = 0.8) -> bool:
    if len(all_masks) < 2:                                        # app.py:47-48
        return False
    d = _dev(all_masks, torch.float32)
    if d.shape[0] > 4:
        raise ValueError("is_ambiguous supports up to 4 masks on the device path")
    for inter, union in pair_counts_device(d).cpu().tolist():
        if inter / (union + 1e-6) < threshold:
            return True
    return False
