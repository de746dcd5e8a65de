"""Drop-in for the training-side `SODPredictor` on the B200-native path.

Reference: /root/reference/synth_sod/src/synth_sod/model_training/predictor.py:22-41 (`PredictionResult`) and :330-477
(`SODPredictor`).  Same constructor arguments, `get_pad_info` / `remove_padding` / `predict` signatures and result
fields.  The math is the `remove_background` path with a different front end:

  * `A.LongestMaxSize(image_size)`: scale = image_size / max(h, w); when scale != 1 the image is resized with
    cv2.INTER_LINEAR to `round(dim * scale)` (Python round, half to even) - NOT the `int()` truncation of `get_pad_info`;
  * `A.PadIfNeeded(image_size, image_size, BORDER_CONSTANT, fill=0)`: centred, top / left = floor, the extra pixel goes
    to the bottom / right (no odd-padding error here, unlike `BackgroundRemoval`);
  * `A.Normalize`: float32 `(v - mean * 255) * (1 / (std * 255))`;
  * the model sees an image_size^2 canvas; its stride-16 patch convolution ignores the last `image_size % 16` rows and
    columns, so the masks come out at `S_eff = 16 * (image_size // 16)` (832 for the default 840) while `remove_padding`
    still slices with the image_size-based padding of `get_pad_info` - the reference's geometry, kept as is;
  * `binary_mask = soft > threshold`, `all_masks` are the binary masks, single-output models return no all_masks / all_ious.

albumentations is not installed in the build container, so its three transforms are restated from their published
functional code (see oracle/sod_predictor.py: "parity unpinned" for that front end; everything behind it is the pinned
`remove_background` arithmetic).  No CPU fallback: the model object raises without the CUDA library / an sm_100 GPU.
"""
from dataclasses import dataclass, replace
from typing import Optional

import numpy as np
import torch

from .arch import ARCHS
from .engine import B200DPTSegmentation


@dataclass
class PredictionResult:
    """Standard prediction result structure (predictor.py:22-41)."""
    binary_mask: np.ndarray                     # thresholded binary mask [H, W]
    soft_mask: np.ndarray                       # probability mask [H, W]
    all_masks: Optional[np.ndarray] = None      # all binary masks [N, H, W] if multiple predictions
    all_ious: Optional[np.ndarray] = None       # all IoU scores [N] if multiple predictions

    @property
    def has_multiple_masks(self) -> bool:
        return self.all_masks is not None

    @property
    def num_masks(self) -> int:
        return len(self.all_masks) if self.has_multiple_masks else 1


def transform_placement(h: int, w: int, image_size: int) -> dict:
    """Where `LongestMaxSize` + centred `PadIfNeeded` put an (h, w) image on the image_size^2 canvas."""
    scale = image_size / float(max(w, h))
    new_h, new_w = (h, w) if scale == 1.0 else (int(round(h * scale)), int(round(w * scale)))
    top = int((image_size - new_h) / 2.0) if new_h < image_size else 0
    left = int((image_size - new_w) / 2.0) if new_w < image_size else 0
    return {"height_pad": top, "width_pad": left, "original_size": (h, w), "resized_size": (new_h, new_w)}


def _strip_model_prefix(key: str) -> str:
    """`k.lstrip("model").lstrip(".")` of predictor.py:365: the character strip stops at the first '.', so for the
    Lightning keys `model.<name>` it removes exactly the leading `model.`."""
    return key.lstrip("model").lstrip(".")


class SODPredictor:
    def __init__(self, checkpoint_path: str, image_size: int = 840, device: str = "cuda", encoder_name: Optional[str] = None,
                 num_outputs: Optional[int] = None, max_batch: int = 1):
        self.device = device
        if not str(device).startswith("cuda"):
            raise RuntimeError("s3od_b200.SODPredictor runs on CUDA (sm_100a) devices only; there is no CPU fallback")
        self.image_size = int(image_size)
        self._s_eff = 16 * (self.image_size // 16)              # what the stride-16 patch embedding keeps
        if self._s_eff < 32 or self._s_eff % 32 != 0:
            raise ValueError(f"image_size {image_size} gives an odd {self._s_eff // 16}-patch grid; the DPT head needs an even one")
        self.model = self._load_checkpoint(checkpoint_path, encoder_name, num_outputs, max_batch)
        self.model.to(device)
        self.model.eval()

    def _load_checkpoint(self, checkpoint_path: str, encoder_name, num_outputs, max_batch) -> B200DPTSegmentation:
        """predictor.py:358-372.  A Lightning checkpoint ({'state_dict': {'model.…'}, 'hyper_parameters': {'config': …}}),
        the exported {'state_dict'} of src/s3od, or a pickled nn.Module."""
        checkpoint = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
        if isinstance(checkpoint, dict) and "state_dict" in checkpoint:
            sd = checkpoint["state_dict"]
            if any(k.startswith("model.") for k in sd):
                sd = {_strip_model_prefix(k): v for k, v in sd.items() if k.startswith("model.")}
            cfg = (checkpoint.get("hyper_parameters") or {}).get("config")
            model_cfg = getattr(cfg, "model", None) if cfg is not None else None
            if model_cfg is None and isinstance(cfg, dict):
                model_cfg = cfg.get("model")
            if model_cfg is not None:
                get = model_cfg.get if hasattr(model_cfg, "get") else (lambda k, d=None: getattr(model_cfg, k, d))
                encoder_name = encoder_name or get("encoder_name")
                num_outputs = num_outputs or get("num_outputs")
        elif hasattr(checkpoint, "state_dict"):
            sd = checkpoint.state_dict()
        else:
            raise ValueError(f"{checkpoint_path}: neither a checkpoint with a 'state_dict' nor a pickled module")
        if encoder_name is None:                                 # ViT-B / ViT-L by the width of the class token
            width = sd["encoder.embeddings.cls_token"].shape[-1]
            encoder_name = "dinov3_large" if width == 1024 else "dinov3_base"
        if num_outputs is None:
            num_outputs = sum(1 for k in sd if k.startswith("seg_head.mask_head.mask_heads.") and k.endswith(".2.weight"))
        arch = ARCHS[encoder_name]
        if arch.num_outputs != num_outputs:
            arch = replace(arch, num_outputs=int(num_outputs))
        return B200DPTSegmentation(sd, arch, self._s_eff, self.device, max_batch=max_batch, normalisation="albumentations")

    def get_pad_info(self, image: np.ndarray) -> dict:
        """predictor.py:374-398: the same `int()` geometry as s3od.utils.get_pad_info, on self.image_size."""
        h, w = image.shape[:2]
        aspect_ratio = w / h
        if aspect_ratio > 1:
            new_w = self.image_size
            new_h = int(new_w / aspect_ratio)
            return {"height_pad": (self.image_size - new_h) // 2, "width_pad": 0, "original_size": (h, w),
                    "resized_size": (new_h, new_w)}
        new_h = self.image_size
        new_w = int(new_h * aspect_ratio)
        return {"height_pad": 0, "width_pad": (self.image_size - new_w) // 2, "original_size": (h, w),
                "resized_size": (new_h, new_w)}

    def remove_padding(self, masks, pad_info: dict):
        """predictor.py:400-406 on a (N, S, S) array or tensor."""
        if pad_info["height_pad"] > 0:
            masks = masks[:, pad_info["height_pad"]:-pad_info["height_pad"], :]
        if pad_info["width_pad"] > 0:
            masks = masks[:, :, pad_info["width_pad"]:-pad_info["width_pad"]]
        return masks

    @torch.no_grad()
    def predict(self, image: np.ndarray, threshold: float = 0.5) -> PredictionResult:
        """predictor.py:408-477: RGB uint8 [H, W, 3] -> PredictionResult."""
        if not (isinstance(image, np.ndarray) and image.dtype == np.uint8 and image.ndim == 3 and image.shape[2] == 3):
            raise ValueError(f"expected an RGB uint8 array of shape (H, W, 3), got {getattr(image, 'shape', None)}")
        model = self.model
        pad_info = self.get_pad_info(image)
        if self._s_eff - 2 * pad_info["height_pad"] < 1 or self._s_eff - 2 * pad_info["width_pad"] < 1:
            raise ValueError("image too elongated: nothing is left of the mask after remove_padding")
        d_img = torch.from_numpy(np.ascontiguousarray(image)).to(model.device)
        model.preprocess([d_img], placements=[transform_placement(image.shape[0], image.shape[1], self.image_size)])
        with torch.cuda.device(model.dev_index):
            out = model._forward_staged(1)
        outs, ious, best = model.postprocess(out["pred_masks"], out["pred_iou"], [d_img], [pad_info])
        all_soft = outs[0][0]                                    # (N, H, W) soft masks at the source size
        all_binary = model.threshold(all_soft, threshold)
        if model.K == 1:
            soft = all_soft[0].cpu().numpy()
            return PredictionResult(binary_mask=all_binary[0].cpu().numpy(), soft_mask=soft)
        ious_np = ious[0].cpu().numpy()
        best_idx = int(best[0])
        all_binary_np = all_binary.cpu().numpy()
        return PredictionResult(binary_mask=all_binary_np[best_idx], soft_mask=all_soft[best_idx].cpu().numpy(),
                                all_masks=all_binary_np, all_ious=ious_np)
