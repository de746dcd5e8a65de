"""CPU restatement of the reference visualisation helpers (TEST INFRASTRUCTURE - see oracle/__init__.py).

Follows /root/reference/src/s3od/visualizer.py:8-48 (visualize_removal, visualize_all_masks) and
/root/reference/demo/app.py:38-56 (compute_mask_iou, is_ambiguous) in plain numpy; returns arrays instead of PIL images.
"""
import numpy as np


def visualize_removal(image: np.ndarray, predicted_mask: np.ndarray, background_color=(0, 255, 0)) -> np.ndarray:
    mask = predicted_mask[..., None]                                   # visualizer.py:17
    background = np.full_like(image, background_color, dtype=np.uint8)  # :19
    return (mask * image + (1 - mask) * background).astype(np.uint8)    # :21 (float32 arithmetic, truncation)


def visualize_all_masks(image: np.ndarray, all_masks: np.ndarray) -> np.ndarray:
    h, w = image.shape[:2]
    num_masks = len(all_masks)
    grid_width = min(num_masks, 4)                                      # :35
    grid_height = (num_masks + grid_width - 1) // grid_width
    grid = np.zeros((h * grid_height, w * grid_width, 3), dtype=np.uint8)
    for idx, mask in enumerate(all_masks):
        row, col = idx // grid_width, idx % grid_width
        grid[row * h:(row + 1) * h, col * w:(col + 1) * w] = (mask[..., None] * image).astype(np.uint8)   # :44-45
    return grid


def compute_mask_iou(mask1, mask2):
    intersection = np.logical_and(mask1 > 0.5, mask2 > 0.5).sum()       # app.py:40
    union = np.logical_or(mask1 > 0.5, mask2 > 0.5).sum()
    return intersection / (union + 1e-6)


def is_ambiguous(all_masks, threshold=0.8):
    if len(all_masks) < 2:                                              # app.py:47
        return False
    for i in range(len(all_masks)):
        for j in range(i + 1, len(all_masks)):
            if compute_mask_iou(all_masks[i], all_masks[j]) < threshold:
                return True
    return False
