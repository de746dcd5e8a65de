"""Checkpoint ingest: reference `state_dict` -> packed tensors for the CUDA library.

Input is the reference checkpoint format (`{'state_dict': ...}`, /root/reference/src/s3od/predictor.py:65,76;
key list in SURVEY 8b).  Both `encoder.model.layer.N.*` (transformers 5.x) and `encoder.layer.N.*` prefixes are
accepted.  Output: dict name -> contiguous CPU tensor (bf16 for GEMM operands, fp32 for vectors):

  * every GEMM B operand is K-major `[N, K]` bf16 (what TMA + tcgen05 read);
  * q/k/v projections are fused into one `[3D, D]` matrix (k bias = 0, config.json key_bias=false);
  * convolution weights are tap-major `[Cout, (ky*3+kx)*Cin + ci]` to match NHWC implicit GEMM;
  * eval-mode BatchNorm of the residual units is folded into the preceding conv (model.py:334-345);
  * the k == stride transposed convs become one `[k*k*Cout, Cin]` GEMM (depth-to-space in the epilogue);
  * the k4 s2 p1 transposed conv is split into its four 2x2 sub-pixel phases;
  * the `num_outputs` mask heads are merged into one `[32*K, 9*64]` conv + a `[K, 32]` block of 1x1 weights;
  * RoPE cos/sin tables (HF:168-200) and the preprocess normalisation LUT (predictor.py:91) are precomputed.
"""
import math
from typing import Dict

import numpy as np
import torch

from .arch import ArchSpec


def _enc_prefix(sd) -> str:
    return "encoder.model.layer." if any(k.startswith("encoder.model.layer.") for k in sd) else "encoder.layer."


def _bf16(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).contiguous()


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.float32).contiguous()


def conv_to_gemm(w: torch.Tensor) -> torch.Tensor:
    """(Cout, Cin, kh, kw) -> [Cout, (ky*kw + kx)*Cin + ci]  (tap-major, channels fastest: NHWC implicit GEMM)."""
    co, ci, kh, kw = w.shape
    return w.permute(0, 2, 3, 1).reshape(co, kh * kw * ci)


def fold_bn(w, b, gamma, beta, mean, var, eps):
    """conv followed by eval BatchNorm == conv with scaled weights: y = (conv(x)+b - mean) * gamma/sqrt(var+eps) + beta."""
    s = gamma / torch.sqrt(var + eps)
    return w * s[:, None, None, None], (b - mean) * s + beta


def rope_tables(gh: int, gw: int, head_dim: int = 64, theta: float = 100.0):
    """cos/sin (P, head_dim/2) fp32: eval path of DINOv3ViTRopePositionEmbedding (HF:168-200); the reference tiles the
    32 angles twice, so only the first half is stored."""
    inv_freq = 1.0 / theta ** torch.arange(0, 1, 4 / head_dim, dtype=torch.float32)
    ch = torch.arange(0.5, gh, dtype=torch.float32) / gh
    cw = torch.arange(0.5, gw, dtype=torch.float32) / gw
    coords = torch.stack(torch.meshgrid(ch, cw, indexing="ij"), dim=-1).flatten(0, 1)
    coords = 2.0 * coords - 1.0
    angles = (2 * math.pi * coords[:, :, None] * inv_freq[None, None, :]).flatten(1, 2)
    return torch.cos(angles).contiguous(), torch.sin(angles).contiguous()


def normalisation_lut() -> torch.Tensor:
    """bf16[3*256]: (v/255 - mean)/std computed like predictor.py:91 (float32 /255, float64 mean/std, cast to fp32)."""
    mean = np.array([0.485, 0.456, 0.406])
    std = np.array([0.229, 0.224, 0.225])
    v = np.arange(256, dtype=np.uint8).astype(np.float32) / 255.0
    lut = ((v[None, :] - mean[:, None]) / std[:, None]).astype(np.float32)
    return torch.from_numpy(lut).reshape(-1).to(torch.bfloat16).contiguous()


def normalisation_affine() -> torch.Tensor:
    """fp32[6] = (a_c, b_c): (v/255 - mean)/std ~= fma(v, a_c, b_c).  Only a hint: s3od_finalize uses it if it reproduces
    every bf16 entry of `normalisation_lut` exactly, otherwise the kernel keeps the table."""
    mean = np.array([0.485, 0.456, 0.406])
    std = np.array([0.229, 0.224, 0.225])
    return torch.from_numpy(np.concatenate([1.0 / (255.0 * std), -mean / std]).astype(np.float32)).contiguous()


# ConvTranspose2d(k4, s2, p1): output row 2i+a gets input rows / kernel rows  a=0: (i, kh=1), (i-1, kh=3);
# a=1: (i+1, kh=0), (i, kh=2).  engine.cu::geom_convt_phase uses the same (offset, tap) order.
_CT_K = {0: (1, 3), 1: (0, 2)}


# Row-streaming form of the same transposed conv: output column 2j+b takes kernel columns (kw, input column shift)
# b=0: (1, 0), (3, -1);  b=1: (0, +1), (2, 0);  all four kernel rows are stacked (output row 2i - 1 + kh).
_CT_ROW_KW = {0: (1, 3), 1: (0, 2)}


def convt_rows_weights(wt: torch.Tensor) -> torch.Tensor:
    """(Cin, Cout, 4, 4) -> [((b*2 + t)*4 + kh)*Cout + co, Cin]"""
    blocks = []
    for b in (0, 1):
        for kw in _CT_ROW_KW[b]:
            for kh in range(4):
                blocks.append(wt[:, :, kh, kw].t())                                   # (Cout, Cin)
    return torch.cat(blocks, dim=0)


def albumentations_lut() -> torch.Tensor:
    """bf16[3*256] of `A.Normalize(mean, std)` on uint8 input, the front end of the training-side SODPredictor
    (synth_sod/.../predictor.py:355): float32 throughout - `img.astype(f32)`, `img -= mean * 255`, `img *= 1 / (std * 255)`
    with mean / std float32 arrays.  albumentations is not installed here, so this arithmetic is restated from its
    published functional `normalize` (parity unpinned, see oracle/sod_predictor.py)."""
    f32 = np.float32
    mean = np.array([0.485, 0.456, 0.406], dtype=f32) * f32(255.0)
    den = np.reciprocal(np.array([0.229, 0.224, 0.225], dtype=f32) * f32(255.0), dtype=f32)
    v = np.arange(256, dtype=np.uint8).astype(f32)
    lut = ((v[None, :] - mean[:, None]).astype(f32) * den[:, None]).astype(f32)
    return torch.from_numpy(lut).reshape(-1).to(torch.bfloat16).contiguous()


def albumentations_affine() -> torch.Tensor:
    """fp32[6] hint for the table above: (v - m) * d ~= fma(v, d, -m * d); only used if it reproduces the table."""
    f32 = np.float32
    mean = np.array([0.485, 0.456, 0.406], dtype=f32) * f32(255.0)
    den = np.reciprocal(np.array([0.229, 0.224, 0.225], dtype=f32) * f32(255.0), dtype=f32)
    return torch.from_numpy(np.concatenate([den, (-mean.astype(np.float64) * den.astype(np.float64)).astype(f32)])).contiguous()


def expected_shapes(arch: ArchSpec, prefix: str = "encoder.model.layer.") -> Dict[str, tuple]:
    """Key -> shape of the reference checkpoint for `arch` (SURVEY 8b weight contract; what `DPTSegmentation(...)` builds at
    /root/reference/src/s3od/predictor.py:67-74).  `num_batches_tracked` buffers are listed with shape ()."""
    D, I, F, K, ps = arch.hidden, arch.mlp, arch.features, arch.num_outputs, arch.patch
    oc, inter = arch.out_channels, arch.inter_features
    sh: Dict[str, tuple] = {}
    e = "encoder.embeddings."
    sh[e + "cls_token"] = (1, 1, D)
    sh[e + "mask_token"] = (1, 1, D)
    sh[e + "register_tokens"] = (1, arch.n_prefix - 1, D)
    sh[e + "patch_embeddings.weight"] = (D, 3, ps, ps)
    sh[e + "patch_embeddings.bias"] = (D,)
    for i in range(arch.layers):
        p = f"{prefix}{i}."
        for n in ("norm1", "norm2"):
            sh[p + n + ".weight"] = sh[p + n + ".bias"] = (D,)
        for n in ("q_proj", "v_proj", "o_proj"):
            sh[p + f"attention.{n}.weight"] = (D, D)
            sh[p + f"attention.{n}.bias"] = (D,)
        sh[p + "attention.k_proj.weight"] = (D, D)
        sh[p + "layer_scale1.lambda1"] = sh[p + "layer_scale2.lambda1"] = (D,)
        sh[p + "mlp.up_proj.weight"], sh[p + "mlp.up_proj.bias"] = (I, D), (I,)
        sh[p + "mlp.down_proj.weight"], sh[p + "mlp.down_proj.bias"] = (D, I), (D,)
    sh["encoder.norm.weight"] = sh["encoder.norm.bias"] = (D,)
    h = "seg_head."
    for j, c in enumerate(oc):
        sh[h + f"projects.{j}.weight"], sh[h + f"projects.{j}.bias"] = (c, D, 1, 1), (c,)
        sh[h + f"scratch.layer{j + 1}_rn.weight"] = (F, c, 3, 3)
    sh[h + "resize_layers.0.weight"], sh[h + "resize_layers.0.bias"] = (oc[0], oc[0], 4, 4), (oc[0],)
    sh[h + "resize_layers.1.weight"], sh[h + "resize_layers.1.bias"] = (oc[1], oc[1], 2, 2), (oc[1],)
    sh[h + "resize_layers.3.weight"], sh[h + "resize_layers.3.bias"] = (oc[3], oc[3], 3, 3), (oc[3],)
    for r in range(1, 5):
        p = h + f"scratch.refinenet{r}."
        sh[p + "out_conv.weight"], sh[p + "out_conv.bias"] = (F, F, 1, 1), (F,)
        for u in (1, 2):
            for cidx in (1, 2):
                q = p + f"resConfUnit{u}."
                sh[q + f"conv{cidx}.weight"], sh[q + f"conv{cidx}.bias"] = (F, F, 3, 3), (F,)
                for n in ("weight", "bias", "running_mean", "running_var"):
                    sh[q + f"bn{cidx}.{n}"] = (F,)
                sh[q + f"bn{cidx}.num_batches_tracked"] = ()
    m = h + "mask_head."
    sh[m + "output_conv1.weight"], sh[m + "output_conv1.bias"] = (F // 2, F, 3, 3), (F // 2,)
    sh[m + "upsample_2x.0.weight"], sh[m + "upsample_2x.0.bias"] = (F // 2, 2 * inter, 4, 4), (2 * inter,)
    sh[m + "upsample_2x.2.weight"], sh[m + "upsample_2x.2.bias"] = (2 * inter, 2 * inter, 3, 3), (2 * inter,)
    for k in range(K):
        sh[m + f"mask_heads.{k}.0.weight"], sh[m + f"mask_heads.{k}.0.bias"] = (inter, 2 * inter, 3, 3), (inter,)
        sh[m + f"mask_heads.{k}.2.weight"], sh[m + f"mask_heads.{k}.2.bias"] = (1, inter, 1, 1), (1,)
    c = h + "classifier_head."
    sh[c + "2.weight"], sh[c + "2.bias"] = (64, F), (64,)
    sh[c + "4.weight"], sh[c + "4.bias"] = (K, 64), (K,)
    return sh


# present in some exports, never in the parameter set the path reads: the non-persistent RoPE buffer and k_proj.bias
# tensors of checkpoints written with key_bias=true (config.json:12 says false; a present bias is honoured below)
_OPTIONAL_SUFFIXES = ("rope_embeddings.inv_freq", "attention.k_proj.bias", "num_batches_tracked")


def validate_state_dict(sd: Dict[str, torch.Tensor], arch: ArchSpec) -> None:
    """Strict ingest, like `model.load_state_dict(state_dict)` at /root/reference/src/s3od/predictor.py:76: missing keys,
    unexpected keys and size mismatches raise RuntimeError (the type torch raises there) before anything reaches the GPU."""
    want = expected_shapes(arch, _enc_prefix(sd))
    missing = [k for k in want if k not in sd and not k.endswith(_OPTIONAL_SUFFIXES)]
    unexpected = [k for k in sd if k not in want and not k.endswith(_OPTIONAL_SUFFIXES)]
    mismatched = [f"{k}: checkpoint {tuple(sd[k].shape)} vs model {want[k]}" for k in want
                  if k in sd and torch.is_tensor(sd[k]) and tuple(sd[k].shape) != want[k]]
    if missing or unexpected or mismatched:
        parts = []
        if missing:
            parts.append("Missing key(s) in state_dict: " + ", ".join(repr(k) for k in missing[:8]) + (" ..." if len(missing) > 8 else ""))
        if unexpected:
            parts.append("Unexpected key(s) in state_dict: " + ", ".join(repr(k) for k in unexpected[:8]) + (" ..." if len(unexpected) > 8 else ""))
        if mismatched:
            parts.append("size mismatch for " + "; ".join(mismatched[:8]) + (" ..." if len(mismatched) > 8 else ""))
        raise RuntimeError(f"Error(s) in loading state_dict for DPTSegmentation ({arch.name}, num_outputs={arch.num_outputs}):\n\t"
                           + "\n\t".join(parts))


def pack_weights(sd: Dict[str, torch.Tensor], arch: ArchSpec, image_size: int, normalisation: str = "s3od") -> Dict[str, torch.Tensor]:
    validate_state_dict(sd, arch)
    D, I, K = arch.hidden, arch.mlp, arch.num_outputs
    out: Dict[str, torch.Tensor] = {}
    e = "encoder.embeddings."
    out["patch.w"] = _bf16(sd[e + "patch_embeddings.weight"].reshape(D, -1))          # k = c*256 + ky*16 + kx
    out["patch.b"] = _f32(sd[e + "patch_embeddings.bias"])
    out["prefix"] = _f32(torch.cat([sd[e + "cls_token"].reshape(1, D), sd[e + "register_tokens"].reshape(-1, D)], 0))
    g = image_size // arch.patch
    out["rope.cos"], out["rope.sin"] = rope_tables(g, g, arch.head_dim, arch.rope_theta)
    if normalisation == "s3od":                     # BackgroundRemoval._preprocess (src/s3od/predictor.py:91)
        out["pre.lut"] = normalisation_lut()
        out["pre.affine"] = normalisation_affine()  # optional fast form; the library verifies it against pre.lut
    elif normalisation == "albumentations":         # SODPredictor's A.Normalize (synth_sod/.../predictor.py:355)
        out["pre.lut"] = albumentations_lut()
        out["pre.affine"] = albumentations_affine()
    else:
        raise ValueError(f"unknown normalisation {normalisation!r}")

    pre = _enc_prefix(sd)
    for l in range(arch.layers_needed):                                               # layer 12 / final norm are dead (F3)
        p, o = f"{pre}{l}.", f"enc.{l}."
        out[o + "ln1.w"], out[o + "ln1.b"] = _f32(sd[p + "norm1.weight"]), _f32(sd[p + "norm1.bias"])
        out[o + "ln2.w"], out[o + "ln2.b"] = _f32(sd[p + "norm2.weight"]), _f32(sd[p + "norm2.bias"])
        a = p + "attention."
        kb = sd.get(a + "k_proj.bias", torch.zeros(D))
        out[o + "qkv.w"] = _bf16(torch.cat([sd[a + "q_proj.weight"], sd[a + "k_proj.weight"], sd[a + "v_proj.weight"]], 0))
        out[o + "qkv.b"] = _f32(torch.cat([sd[a + "q_proj.bias"], kb, sd[a + "v_proj.bias"]], 0))
        out[o + "o.w"], out[o + "o.b"] = _bf16(sd[a + "o_proj.weight"]), _f32(sd[a + "o_proj.bias"])
        out[o + "ls1"], out[o + "ls2"] = _f32(sd[p + "layer_scale1.lambda1"]), _f32(sd[p + "layer_scale2.lambda1"])
        out[o + "up.w"], out[o + "up.b"] = _bf16(sd[p + "mlp.up_proj.weight"]), _f32(sd[p + "mlp.up_proj.bias"])
        out[o + "down.w"], out[o + "down.b"] = _bf16(sd[p + "mlp.down_proj.weight"]), _f32(sd[p + "mlp.down_proj.bias"])

    h = "seg_head."
    for j in range(4):
        w = sd[h + f"projects.{j}.weight"]
        out[f"head.proj{j}.w"] = _bf16(w.reshape(w.shape[0], -1))
        out[f"head.proj{j}.b"] = _f32(sd[h + f"projects.{j}.bias"])
    for j, k in ((0, 4), (1, 2)):
        w = sd[h + f"resize_layers.{j}.weight"]                                       # (Cin, Cout, k, k)
        ci, co = w.shape[0], w.shape[1]
        out[f"head.rs{j}.w"] = _bf16(w.permute(2, 3, 1, 0).reshape(k * k * co, ci))   # row = (a*k + b)*Cout + co
        out[f"head.rs{j}.b"] = _f32(sd[h + f"resize_layers.{j}.bias"])
    out["head.rs3.w"] = _bf16(conv_to_gemm(sd[h + "resize_layers.3.weight"]))
    out["head.rs3.b"] = _f32(sd[h + "resize_layers.3.bias"])
    s = h + "scratch."
    for j in range(1, 5):
        out[f"head.rn{j}.w"] = _bf16(conv_to_gemm(sd[s + f"layer{j}_rn.weight"]))
    for r in range(1, 5):
        p = s + f"refinenet{r}."
        w = sd[p + "out_conv.weight"]
        out[f"head.ref{r}.out.w"] = _bf16(w.reshape(w.shape[0], -1))
        out[f"head.ref{r}.out.b"] = _f32(sd[p + "out_conv.bias"])
        for u in ((2,) if r == 4 else (1, 2)):                                        # refinenet4.resConfUnit1 is unused (F8)
            q = p + f"resConfUnit{u}."
            for cidx in (1, 2):
                bn = q + f"bn{cidx}."
                w, b = sd[q + f"conv{cidx}.weight"], sd[q + f"conv{cidx}.bias"]
                if bn + "weight" in sd:
                    w, b = fold_bn(w, b, sd[bn + "weight"], sd[bn + "bias"], sd[bn + "running_mean"], sd[bn + "running_var"],
                                   arch.bn_eps)
                out[f"head.ref{r}.rcu{u}.c{cidx}.w"] = _bf16(conv_to_gemm(w))
                out[f"head.ref{r}.rcu{u}.c{cidx}.b"] = _f32(b)
    m = h + "mask_head."
    out["head.mh.c1.w"] = _bf16(conv_to_gemm(sd[m + "output_conv1.weight"]))
    out["head.mh.c1.b"] = _f32(sd[m + "output_conv1.bias"])
    wt = sd[m + "upsample_2x.0.weight"]                                               # (Cin=128, Cout=64, 4, 4)
    phases = []
    for a in (0, 1):
        for b in (0, 1):
            taps = [wt[:, :, kh, kw].t() for kh in _CT_K[a] for kw in _CT_K[b]]       # each (Cout, Cin); tap = r*2 + c
            phases.append(torch.cat(taps, dim=1))                                     # (Cout, 4*Cin)
    out["head.mh.up.w"] = _bf16(torch.cat(phases, dim=0))                             # row = (a*2+b)*Cout + co
    out["head.mh.up.wr"] = _bf16(convt_rows_weights(wt))                              # row-streaming kernel (conv_rows.cuh)
    out["head.mh.up.b"] = _f32(sd[m + "upsample_2x.0.bias"])
    out["head.mh.c2.w"] = _bf16(conv_to_gemm(sd[m + "upsample_2x.2.weight"]))
    out["head.mh.c2.b"] = _f32(sd[m + "upsample_2x.2.bias"])
    out["head.mh.heads.w"] = _bf16(torch.cat([conv_to_gemm(sd[m + f"mask_heads.{k}.0.weight"]) for k in range(K)], 0))
    out["head.mh.heads.b"] = _f32(torch.cat([sd[m + f"mask_heads.{k}.0.bias"] for k in range(K)], 0))
    out["head.mh.heads.w2"] = _f32(torch.cat([sd[m + f"mask_heads.{k}.2.weight"].reshape(1, -1) for k in range(K)], 0))
    out["head.mh.heads.b2"] = _f32(torch.cat([sd[m + f"mask_heads.{k}.2.bias"].reshape(1) for k in range(K)], 0))
    c = h + "classifier_head."
    out["head.cls.w1"], out["head.cls.b1"] = _f32(sd[c + "2.weight"]), _f32(sd[c + "2.bias"])
    out["head.cls.w2"], out["head.cls.b2"] = _f32(sd[c + "4.weight"]), _f32(sd[c + "4.bias"])
    return out
