"""Letterbox geometry - host mirror of /root/reference/src/s3od/utils.py (same names, same arithmetic)."""
from typing import Any, Dict

import numpy as np


def get_pad_info(image: np.ndarray, image_size: int = 1024) -> Dict[str, Any]:
    """utils.py:6-29: long side -> image_size, `int()` truncation of the short side, symmetric floor padding."""
    h, w = image.shape[:2]
    aspect_ratio = w / h
    if aspect_ratio > 1:
        new_w = image_size
        new_h = int(new_w / aspect_ratio)
        return {"height_pad": (image_size - new_h) // 2, "width_pad": 0, "original_size": (h, w), "resized_size": (new_h, new_w)}
    new_h = image_size
    new_w = int(new_h * aspect_ratio)
    return {"height_pad": 0, "width_pad": (image_size - new_w) // 2, "original_size": (h, w), "resized_size": (new_h, new_w)}


def remove_padding(masks, pad_info: Dict[str, Any]):
    """utils.py:32-37 on a (K, S, S) array or tensor."""
    if pad_info["height_pad"] > 0:
        masks = masks[:, pad_info["height_pad"]:-pad_info["height_pad"], :]
    if pad_info["width_pad"] > 0:
        masks = masks[:, :, pad_info["width_pad"]:-pad_info["width_pad"]]
    return masks


def check_padding(pad_info: Dict[str, Any], image_size: int) -> None:
    """The reference pastes the resized image into `padded[pad:-pad]`; when (image_size - resized) is odd the slice
    is one pixel short and numpy raises ValueError (predictor.py:83-89).  Same error, same condition."""
    new_h, new_w = pad_info["resized_size"]
    hp, wp = pad_info["height_pad"], pad_info["width_pad"]
    if hp > 0 and image_size - 2 * hp != new_h:
        raise ValueError(f"could not broadcast input array from shape ({new_h},{new_w},3) into shape "
                         f"({image_size - 2 * hp},{image_size},3)")
    if hp == 0 and wp > 0 and image_size - 2 * wp != new_w:
        raise ValueError(f"could not broadcast input array from shape ({new_h},{new_w},3) into shape "
                         f"({image_size},{image_size - 2 * wp},3)")
    if hp == 0 and wp == 0 and (new_h != image_size or new_w != image_size):
        # One pixel short on the short side (e.g. a 1025 x 1024 source): both pads are 0, so the reference takes
        # `padded = resized` (predictor.py:88-89) and runs the network on a NON-square (S-1) x S input - a 63 x 64 patch grid
        # and a 1008-row mask.  The launch plan here is built for the square S x S canvas only; rather than return a mask
        # that differs from the reference's, refuse the geometry explicitly.
        raise ValueError(f"unsupported geometry: the resized image is {new_h}x{new_w} with zero padding, which the reference "
                         f"feeds to the network as a non-square input; s3od_b200 only runs the square {image_size}x{image_size} canvas")
