"""Seeded synthetic checkpoints and images.

Checkpoints and datasets are not available offline, so parity and throughput are measured on
random weights of the named architecture.  Vanilla random init is useless for that (SURVEY F12:
sigmoid masks land in [0.48, 0.52] and BN / LayerScale / biases are identities), so every tensor is
drawn here from a seeded distribution that makes each folded term non-trivial and gives mask logits
a standard deviation of a few units.

The key set and shapes are exactly those of the reference checkpoint format
(`{'state_dict': OrderedDict}`; /root/reference/scripts/export_model.py:97-109, read back at
/root/reference/src/s3od/predictor.py:65,76; full key list in SURVEY 8b).
"""
from collections import OrderedDict
import math

import numpy as np
import torch

from .arch import ArchSpec, VITB


def _randn(gen, shape, std=1.0, mean=0.0):
    return torch.randn(shape, generator=gen, dtype=torch.float32) * std + mean


def _rand(gen, shape, lo, hi):
    return torch.rand(shape, generator=gen, dtype=torch.float32) * (hi - lo) + lo


def synth_state_dict(arch: ArchSpec = VITB, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic fp32 state_dict with the reference's key names (encoder.model.layer.N.* layout)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1234567 + seed)
    D, I, F = arch.hidden, arch.mlp, arch.features
    sd = OrderedDict()

    def lin(prefix, out_f, in_f, bias=True, gain=1.0, bias_std=0.1):
        sd[prefix + ".weight"] = _randn(g, (out_f, in_f), gain / math.sqrt(in_f))
        if bias:
            sd[prefix + ".bias"] = _randn(g, (out_f,), bias_std)

    def conv(prefix, out_c, in_c, k, bias=True, gain=1.0, bias_std=0.1, transposed=False):
        shape = (in_c, out_c, k, k) if transposed else (out_c, in_c, k, k)
        sd[prefix + ".weight"] = _randn(g, shape, gain / math.sqrt(in_c * k * k))
        if bias:
            sd[prefix + ".bias"] = _randn(g, (out_c,), bias_std)

    e = "encoder.embeddings."
    sd[e + "cls_token"] = _randn(g, (1, 1, D), 1.0)
    sd[e + "mask_token"] = torch.zeros(1, 1, D)
    sd[e + "register_tokens"] = _randn(g, (1, 4, D), 1.0)
    # pixels are ~N(0,1) after normalisation: unit-gain patch embedding keeps tokens O(1)
    sd[e + "patch_embeddings.weight"] = _randn(g, (D, 3, arch.patch, arch.patch), 1.0 / math.sqrt(3 * arch.patch ** 2))
    sd[e + "patch_embeddings.bias"] = _randn(g, (D,), 0.2)

    for i in range(arch.layers):
        p = f"encoder.model.layer.{i}."
        sd[p + "norm1.weight"] = _randn(g, (D,), 0.2, 1.0)
        sd[p + "norm1.bias"] = _randn(g, (D,), 0.1)
        # gain 2 on q/k makes the softmax peaky enough that attention is not a plain average
        sd[p + "attention.k_proj.weight"] = _randn(g, (D, D), 2.0 / math.sqrt(D))
        lin(p + "attention.v_proj", D, D)
        lin(p + "attention.q_proj", D, D, gain=2.0)
        lin(p + "attention.o_proj", D, D)
        sd[p + "layer_scale1.lambda1"] = _rand(g, (D,), 0.2, 1.0)
        sd[p + "norm2.weight"] = _randn(g, (D,), 0.2, 1.0)
        sd[p + "norm2.bias"] = _randn(g, (D,), 0.1)
        lin(p + "mlp.up_proj", I, D, gain=1.4)
        lin(p + "mlp.down_proj", D, I, gain=1.4)
        sd[p + "layer_scale2.lambda1"] = _rand(g, (D,), 0.2, 1.0)
    sd["encoder.norm.weight"] = _randn(g, (D,), 0.2, 1.0)
    sd["encoder.norm.bias"] = _randn(g, (D,), 0.1)

    h = "seg_head."
    oc = arch.out_channels
    for i, c in enumerate(oc):
        # taps are un-normed residual streams whose rms grows with depth; 0.5 gain keeps the head O(1)
        conv(h + f"projects.{i}", c, D, 1, gain=0.5)
    conv(h + "resize_layers.0", oc[0], oc[0], 4, transposed=True)
    conv(h + "resize_layers.1", oc[1], oc[1], 2, transposed=True)
    conv(h + "resize_layers.3", oc[3], oc[3], 3)
    for i, c in enumerate(oc):
        conv(h + f"scratch.layer{i + 1}_rn", F, c, 3, bias=False)
    for r in (1, 2, 3, 4):
        p = h + f"scratch.refinenet{r}."
        conv(p + "out_conv", F, F, 1)
        for u in (1, 2):
            q = p + f"resConfUnit{u}."
            conv(q + "conv1", F, F, 3, gain=1.4)
            conv(q + "conv2", F, F, 3, gain=0.7)
            for b in (1, 2):
                sd[q + f"bn{b}.weight"] = _randn(g, (F,), 0.2, 1.0)
                sd[q + f"bn{b}.bias"] = _randn(g, (F,), 0.1)
                sd[q + f"bn{b}.running_mean"] = _randn(g, (F,), 0.2)
                sd[q + f"bn{b}.running_var"] = _rand(g, (F,), 0.5, 1.5)
                sd[q + f"bn{b}.num_batches_tracked"] = torch.tensor(100, dtype=torch.int64)
    m = h + "mask_head."
    conv(m + "output_conv1", F // 2, F, 3)
    conv(m + "upsample_2x.0", 2 * arch.inter_features, F // 2, 4, transposed=True, gain=2.8)  # 4 of 16 taps hit
    conv(m + "upsample_2x.2", 2 * arch.inter_features, 2 * arch.inter_features, 3, gain=1.4)
    for k in range(arch.num_outputs):
        conv(m + f"mask_heads.{k}.0", arch.inter_features, 2 * arch.inter_features, 3, gain=1.4)
        conv(m + f"mask_heads.{k}.2", 1, arch.inter_features, 1, gain=1.0, bias_std=0.5)
        # the 1x1 sees post-ReLU (positive) features: zero-sum weights keep the logits centred on 0
        sd[m + f"mask_heads.{k}.2.weight"] -= sd[m + f"mask_heads.{k}.2.weight"].mean()
    lin(h + "classifier_head.2", 64, F, gain=1.0)
    lin(h + "classifier_head.4", arch.num_outputs, 64, gain=1.0, bias_std=0.3)
    sd[h + "classifier_head.4.weight"] -= sd[h + "classifier_head.4.weight"].mean(dim=1, keepdim=True)
    return sd


def save_checkpoint(path, arch: ArchSpec = VITB, seed: int = 0):
    """Write the reference's inference checkpoint format (scripts/export_model.py:97-109)."""
    torch.save({"state_dict": synth_state_dict(arch, seed)}, path)
    return path


def synth_image(h: int, w: int, seed: int = 0) -> np.ndarray:
    """Seeded uint8 RGB test image: smooth blobs + noise so masks have structure (seeded rng)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.zeros((h, w, 3), np.float32)
    for c in range(3):
        acc = np.zeros((h, w), np.float32)
        for _ in range(4):
            cy, cx = rng.uniform(0, h), rng.uniform(0, w)
            s = rng.uniform(0.08, 0.35) * max(h, w)
            acc += rng.uniform(0.3, 1.0) * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * s * s))
        img[..., c] = acc / max(acc.max(), 1e-6)
    img = img * 200.0 + rng.uniform(0, 55, size=(h, w, 3)).astype(np.float32)
    return np.clip(img, 0, 255).astype(np.uint8)


def synth_noise_image(h: int, w: int, seed: int = 0) -> np.ndarray:
    """Uniform uint8 noise (the reference tests' fixture style, tests/conftest.py:39-54)."""
    return np.random.default_rng(seed).integers(0, 256, size=(h, w, 3), dtype=np.uint8)
