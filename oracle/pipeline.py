"""End-to-end CPU oracle of BackgroundRemoval.remove_background (TEST INFRASTRUCTURE - see oracle/__init__.py).

Follows /root/reference/src/s3od/predictor.py:96-139 with the restated pieces of oracle/model.py and
oracle/prepost.py.  Returns plain numpy arrays (the RGBA image as an (H,W,4) uint8 array).
"""
import numpy as np
import torch

from . import model as _model
from . import prepost as _pp


def remove_background(sd, image: np.ndarray, arch, image_size: int = 1024):
    x, pad = _pp.preprocess(image, image_size)
    out = _model.forward(sd, torch.from_numpy(x), arch)
    res = _pp.postprocess(out["pred_masks"][0].numpy(), out["pred_iou"][0].numpy(), pad, image)
    res["pred_masks"] = out["pred_masks"][0].numpy()
    res["pred_iou"] = out["pred_iou"][0].numpy()
    return res
