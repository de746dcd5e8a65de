"""fp32 CPU restatement of DPTSegmentation.forward (TEST INFRASTRUCTURE - see oracle/__init__.py).

Follows, function by function:
  * DINOv3ViTEmbeddings.forward          HF modeling_dinov3_vit.py:75-92
  * get_patches_center_coordinates       HF:95-121
  * DINOv3ViTRopePositionEmbedding.forward (eval)  HF:168-200
  * rotate_half / apply_rotary_pos_emb   HF:203-207, 238-268
  * DINOv3ViTAttention.forward           HF:294-334   (softmax(q k^T / 8) v, k_proj has no bias)
  * DINOv3ViTMLP / LayerScale / Layer    HF:385-386, 342-343, 424-450
  * extract_intermediate_features        /root/reference/src/s3od/model.py:62-86
  * process_encoder_features             model.py:193-211
  * ResidualConvUnit / FeatureFusionBlock model.py:334-345, 383-405
  * MultiMaskHead.forward                model.py:455-467
  * DPTSegmentationHead.forward          model.py:217-238
(HF = transformers 5.5.0 `models/dinov3_vit/modeling_dinov3_vit.py`, the installed third-party
backbone; the reference pins 4.57.1 in uv.lock:4870.)

Everything operates on a flat `state_dict` with the reference's key names.
"""
import math
from typing import Dict, List

import torch
import torch.nn.functional as F


def _enc_prefix(sd) -> str:
    # transformers 5.x nests the blocks under `encoder.model.layer`, older releases under `encoder.layer`
    return "encoder.model.layer." if any(k.startswith("encoder.model.layer.") for k in sd) else "encoder.layer."


def rope_tables(gh: int, gw: int, head_dim: int = 64, theta: float = 100.0):
    """cos/sin tables (gh*gw, head_dim) in fp32; eval-mode path of HF:168-200."""
    inv_freq = 1.0 / theta ** torch.arange(0, 1, 4 / head_dim, dtype=torch.float32)
    ch = torch.arange(0.5, gh, dtype=torch.float32) / gh
    cw = torch.arange(0.5, gw, dtype=torch.float32) / gw
    coords = torch.stack(torch.meshgrid(ch, cw, indexing="ij"), dim=-1).flatten(0, 1)
    coords = 2.0 * coords - 1.0
    angles = 2 * math.pi * coords[:, :, None] * inv_freq[None, None, :]
    angles = angles.flatten(1, 2).tile(2)
    return torch.cos(angles), torch.sin(angles)


def _rot_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def embed(sd, x):
    """(B,3,H,W) fp32 -> (B, 5+P, D)."""
    e = "encoder.embeddings."
    t = F.conv2d(x, sd[e + "patch_embeddings.weight"], sd[e + "patch_embeddings.bias"], stride=16)
    t = t.flatten(2).transpose(1, 2)
    B = x.shape[0]
    return torch.cat([sd[e + "cls_token"].expand(B, -1, -1), sd[e + "register_tokens"].expand(B, -1, -1), t], dim=1)


def attention(sd, p, xn, cos, sin, heads, stages=None, sdpa=False):
    B, N, D = xn.shape
    hd = D // heads
    q = F.linear(xn, sd[p + "q_proj.weight"], sd[p + "q_proj.bias"])
    k = F.linear(xn, sd[p + "k_proj.weight"], sd.get(p + "k_proj.bias"))
    v = F.linear(xn, sd[p + "v_proj.weight"], sd[p + "v_proj.bias"])
    q = q.view(B, N, heads, hd).transpose(1, 2)
    k = k.view(B, N, heads, hd).transpose(1, 2)
    v = v.view(B, N, heads, hd).transpose(1, 2)
    npre = N - cos.shape[0]
    qp, kp = q[:, :, npre:], k[:, :, npre:]
    q = torch.cat((q[:, :, :npre], qp * cos + _rot_half(qp) * sin), dim=2)
    k = torch.cat((k[:, :, :npre], kp * cos + _rot_half(kp) * sin), dim=2)
    if stages is not None:
        stages["q_rope"], stages["k_rope"], stages["v"] = q, k, v
    if sdpa:
        # the reference's own dispatch (HF:316-329 with attn_implementation "sdpa"): flash / memory-efficient kernels on a GPU.
        # Used by the same-GPU PyTorch baseline of bench.py and by the error-envelope test; the default below is the plain
        # matmul-softmax form the CPU golden checks were pinned with.
        o = F.scaled_dot_product_attention(q, k, v, scale=hd ** -0.5)
    else:
        s = torch.matmul(q, k.transpose(2, 3)) * (hd ** -0.5)
        o = torch.matmul(torch.softmax(s, dim=-1), v)
    o = o.transpose(1, 2).reshape(B, N, D)
    if stages is not None:
        stages["attn_ctx"] = o
    return F.linear(o, sd[p + "o_proj.weight"], sd[p + "o_proj.bias"])


def encoder_layer(sd, p, x, cos, sin, heads, eps, stages=None, sdpa=False):
    xn = F.layer_norm(x, (x.shape[-1],), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps)
    if stages is not None:
        stages["ln1"] = xn
    x = x + attention(sd, p + "attention.", xn, cos, sin, heads, stages, sdpa) * sd[p + "layer_scale1.lambda1"]
    if stages is not None:
        stages["x_attn"] = x
    xn = F.layer_norm(x, (x.shape[-1],), sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps)
    hmid = F.gelu(F.linear(xn, sd[p + "mlp.up_proj.weight"], sd[p + "mlp.up_proj.bias"]))  # exact erf GELU
    if stages is not None:
        stages["mlp_hidden"] = hmid
    y = F.linear(hmid, sd[p + "mlp.down_proj.weight"], sd[p + "mlp.down_proj.bias"])
    return x + y * sd[p + "layer_scale2.lambda1"]


def encoder_taps(sd, x, arch, stages=None, sdpa=False) -> List[torch.Tensor]:
    """hidden_states[taps][:, 5:]  - un-normed residual streams (model.py:72-84)."""
    gh, gw = x.shape[-2] // arch.patch, x.shape[-1] // arch.patch
    cos, sin = rope_tables(gh, gw, arch.head_dim, arch.rope_theta)
    cos, sin = cos.to(x.device), sin.to(x.device)          # the tests also run this restatement with torch on a GPU
    pre = _enc_prefix(sd)
    h = embed(sd, x)
    if stages is not None:
        stages["embed"] = h
    taps = []
    for i in range(arch.layers_needed):
        st = stages if (stages is not None and i == 0) else None
        h = encoder_layer(sd, f"{pre}{i}.", h, cos, sin, arch.heads, arch.ln_eps, st, sdpa)
        if stages is not None and i == 0:
            stages["layer0"] = h
        if (i + 1) in arch.taps:
            taps.append(h[:, arch.n_prefix:])
    return taps


def _bn(sd, p, x, eps, train=False):
    if train:       # nn.BatchNorm2d in train mode: batch statistics (the running buffers' update does not enter the output)
        return F.batch_norm(x, None, None, sd[p + "weight"], sd[p + "bias"], True, 0.0, eps)
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"], False, 0.0, eps)


def rcu(sd, p, x, eps, train=False):
    o = F.conv2d(F.relu(x), sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1)
    o = _bn(sd, p + "bn1.", o, eps, train)
    o = F.conv2d(F.relu(o), sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1)
    o = _bn(sd, p + "bn2.", o, eps, train)
    return o + x


def fusion(sd, p, x0, skip, size, eps, train=False):
    out = x0
    if skip is not None:
        out = out + rcu(sd, p + "resConfUnit1.", skip, eps, train)
    out = rcu(sd, p + "resConfUnit2.", out, eps, train)
    if size is None:
        out = F.interpolate(out, scale_factor=2, mode="bilinear", align_corners=False)
    else:
        out = F.interpolate(out, size=size, mode="bilinear", align_corners=False)
    return F.conv2d(out, sd[p + "out_conv.weight"], sd[p + "out_conv.bias"])


def head_forward(sd, taps, gh, gw, arch, stages=None, train=False) -> Dict[str, torch.Tensor]:
    """train=True: BatchNorm with batch statistics, as the reference's training step runs the head (lightning_module.py:234-245)."""
    h = "seg_head."
    B = taps[0].shape[0]
    feats = []
    for i, t in enumerate(taps):
        x = t.permute(0, 2, 1).reshape(B, t.shape[-1], gh, gw)
        x = F.conv2d(x, sd[h + f"projects.{i}.weight"], sd[h + f"projects.{i}.bias"])
        if i == 0:
            x = F.conv_transpose2d(x, sd[h + "resize_layers.0.weight"], sd[h + "resize_layers.0.bias"], stride=4)
        elif i == 1:
            x = F.conv_transpose2d(x, sd[h + "resize_layers.1.weight"], sd[h + "resize_layers.1.bias"], stride=2)
        elif i == 3:
            x = F.conv2d(x, sd[h + "resize_layers.3.weight"], sd[h + "resize_layers.3.bias"], stride=2, padding=1)
        feats.append(x)
    s = h + "scratch."
    l = [F.conv2d(feats[i], sd[s + f"layer{i + 1}_rn.weight"], None, padding=1) for i in range(4)]
    eps = arch.bn_eps
    p4 = fusion(sd, s + "refinenet4.", l[3], None, l[2].shape[2:], eps, train)
    p3 = fusion(sd, s + "refinenet3.", p4, l[2], l[1].shape[2:], eps, train)
    p2 = fusion(sd, s + "refinenet2.", p3, l[1], l[0].shape[2:], eps, train)
    p1 = fusion(sd, s + "refinenet1.", p2, l[0], None, eps, train)
    pooled = p1.mean(dim=(2, 3))
    c = h + "classifier_head."
    iou = F.linear(F.relu(F.linear(pooled, sd[c + "2.weight"], sd[c + "2.bias"])), sd[c + "4.weight"], sd[c + "4.bias"])
    m = h + "mask_head."
    f = F.conv2d(p1, sd[m + "output_conv1.weight"], sd[m + "output_conv1.bias"], padding=1)
    f = F.relu(F.conv_transpose2d(f, sd[m + "upsample_2x.0.weight"], sd[m + "upsample_2x.0.bias"], stride=2, padding=1))
    f = F.relu(F.conv2d(f, sd[m + "upsample_2x.2.weight"], sd[m + "upsample_2x.2.bias"], padding=1))
    # model.py:458-460 resizes `f` to (16*gh, 16*gw) with antialias; f already has that size, so it is the identity
    assert f.shape[2:] == (gh * arch.patch, gw * arch.patch)
    masks = []
    for k in range(arch.num_outputs):
        t = F.relu(F.conv2d(f, sd[m + f"mask_heads.{k}.0.weight"], sd[m + f"mask_heads.{k}.0.bias"], padding=1))
        masks.append(F.conv2d(t, sd[m + f"mask_heads.{k}.2.weight"], sd[m + f"mask_heads.{k}.2.bias"]))
    if stages is not None:
        stages.update(l1=l[0], l2=l[1], l3=l[2], l4=l[3], p4=p4, p3=p3, p2=p2, p1=p1, feat=f)
    return {"pred_masks": torch.cat(masks, dim=1), "pred_iou": iou, "features": p1}


@torch.no_grad()
def forward(sd, x, arch, stages=None, sdpa=False) -> Dict[str, torch.Tensor]:
    """DPTSegmentation.forward (model.py:99-106) on fp32 tensors (CPU, or a GPU for the same-device baseline)."""
    x = x.float()
    taps = encoder_taps(sd, x, arch, stages, sdpa)
    if stages is not None:
        for i, t in enumerate(taps):
            stages[f"tap{i}"] = t
    return head_forward(sd, taps, x.shape[-2] // arch.patch, x.shape[-1] // arch.patch, arch, stages)
