"""Architecture constants for the S3OD DPT segmentation models.

Mirrors what the reference derives at construction time:
  * ViT hyper-parameters: /root/reference/src/s3od/dinov3_config/config.json:7-31
  * tap layers            : /root/reference/src/s3od/model.py:36-40
  * DPT head widths       : /root/reference/src/s3od/model.py:44 (out_channels), :67-74 of predictor.py (features=256)
"""
from dataclasses import dataclass, field
from typing import Tuple


@dataclass(frozen=True)
class ArchSpec:
    name: str
    hidden: int
    layers: int
    heads: int
    mlp: int
    taps: Tuple[int, int, int, int]          # indices into hidden_states (0 = embeddings)
    num_outputs: int = 3
    features: int = 256
    out_channels: Tuple[int, int, int, int] = (256, 512, 1024, 1024)
    n_prefix: int = 5                         # cls + 4 register tokens
    patch: int = 16
    head_dim: int = 64
    ln_eps: float = 1e-5
    bn_eps: float = 1e-5
    rope_theta: float = 100.0
    inter_features: int = 32                  # MultiMaskHead inter_features (model.py:181)

    @property
    def layers_needed(self) -> int:
        """hidden_states[k] is the output of layer k-1, so only the first taps[-1] layers matter."""
        return self.taps[-1]


VITB = ArchSpec("dinov3_base", hidden=768, layers=12, heads=12, mlp=3072, taps=(2, 5, 8, 11), num_outputs=3)
# ViT-L/16 values of the DINOv3 release (config not in the reference tree) - dinol.yaml uses num_outputs=1.
VITL = ArchSpec("dinov3_large", hidden=1024, layers=24, heads=16, mlp=4096, taps=(4, 11, 17, 23), num_outputs=1)

ARCHS = {"dinov3_base": VITB, "dinob": VITB, "dinov3_large": VITL, "dinol": VITL}
