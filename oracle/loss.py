"""fp32 restatement of the reference training loss and optimiser step (TEST INFRASTRUCTURE - see oracle/__init__.py).

Follows /root/reference/synth_sod/src/synth_sod/model_training/loss.py with the `focal_iou.yaml` configuration
(config/loss/focal_iou.yaml:1-27: focal 20, IoU 1.0, MSE on the IoU head 0.05, full_mask_lambda 0.1, decay_rate 0.2):
  * FocalLoss.forward                      loss.py:126-143   (alpha 0.25, gamma 2, reduction none)
  * IoULoss.forward                        loss.py:79-99     (smooth 1e-6, reduction none)
  * MaskLossHandler.compute_iou            loss.py:155-164
  * MaskLossHandler.compute_multi_mask_losses   loss.py:190-233
  * LossModule.forward                     loss.py:242-275   (num_masks == 3 branch + the aux MSE component)
and the optimiser of lightning_module.py:183-193 (AdamW, encoder lr, head lr x 10, weight_decay 0.05, betas (0.9, 0.999),
eps 1e-8).

SURVEY F10: `LossComponent.from_dict` never sets `add_sigmoid`, so it is True for every component - FocalLoss receives
ALREADY-SIGMOIDED masks and applies binary_cross_entropy_with_logits to them again.  Reproduced here on purpose.

Everything is differentiable torch code, so `torch.autograd` supplies the reference gradients for the CUDA backward kernels.
"""
import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

FOCAL_WEIGHT, IOU_WEIGHT, MSE_WEIGHT = 20.0, 1.0, 0.05
FULL_MASK_LAMBDA, DECAY_RATE = 0.1, 0.2
ALPHA, GAMMA, SMOOTH = 0.25, 2.0, 1e-6


def focal_loss(pred, target):
    """loss.py:134-137 on whatever it is given (the caller passes sigmoid(pred_masks), F10)."""
    bce = F.binary_cross_entropy_with_logits(pred, target, reduction="none")
    pt = torch.exp(-bce)
    return ALPHA * (1 - pt) ** GAMMA * bce


def iou_loss(pred, target):
    """loss.py:86-93: per-row soft IoU with union = sum(p) + sum(t) - intersection."""
    pred = pred.reshape(pred.size(0), -1)
    target = target.reshape(target.size(0), -1)
    inter = (pred * target).sum(dim=1)
    union = pred.sum(dim=1) + target.sum(dim=1) - inter
    return 1 - (inter + SMOOTH) / (union + SMOOTH)


def compute_iou(pred, target):
    """loss.py:155-164: selection IoU with union = sum(t^2) + sum(p^2) - intersection; inputs (B*K, 1, H, W)."""
    def op_sum(x):
        return x.reshape(x.shape[0], x.shape[1], -1).sum(2)
    inter = op_sum(target * pred)
    union = op_sum(target ** 2) + op_sum(pred ** 2) - inter
    return ((inter + SMOOTH) / (union + SMOOTH)).mean(dim=1)


def loss_module(pred_masks, pred_iou, target_masks, epoch: int) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """LossModule.forward (loss.py:242-275) for num_masks > 1 with the focal_iou.yaml components."""
    B, K = pred_masks.shape[:2]
    target_expanded = target_masks.unsqueeze(1).expand(-1, K, -1, -1)
    exp_decay = FULL_MASK_LAMBDA * math.exp(-DECAY_RATE * epoch)
    pred_sigmoid = torch.sigmoid(pred_masks)
    pred_flat = pred_sigmoid.contiguous().reshape(B * K, 1, *pred_masks.shape[2:])
    gt_flat = target_expanded.contiguous().reshape(B * K, 1, *target_masks.shape[1:])
    with torch.no_grad():
        ious = compute_iou(pred_flat, gt_flat).reshape(B, K)
    best = ious.argmax(dim=1)
    total = torch.tensor(0.0, device=pred_masks.device)
    parts = {"best_iou": ious.max(dim=1)[0].mean(), "gt_ious": ious, "best_indices": best}
    for name, weight, fn in (("focal_loss", FOCAL_WEIGHT, focal_loss), ("iou_loss", IOU_WEIGHT, iou_loss)):
        all_losses = fn(pred_flat, gt_flat)
        if all_losses.dim() == 4:
            all_losses = all_losses.mean(dim=(1, 2, 3))
        all_losses = all_losses.reshape(B, K)
        best_loss = all_losses.gather(1, best.unsqueeze(1)).mean()
        total = total + weight * (best_loss + all_losses.mean() * exp_decay)
        parts[name + "_best"] = best_loss
        parts[name + "_full"] = all_losses
    mse = F.mse_loss(torch.sigmoid(pred_iou), ious)
    total = total + MSE_WEIGHT * mse
    parts["mse_ious_loss"] = mse
    return total, parts


def adamw_step(p, g, m, v, step: int, lr: float, wd: float = 0.05, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8):
    """One torch.optim.AdamW update (decoupled weight decay, bias correction; the single-tensor path of torch/optim/adamw.py)
    written out on plain tensors; returns (p, m, v) after step number `step` (1-based)."""
    p = p * (1 - lr * wd)
    m = m * b1 + g * (1 - b1)
    v = v * b2 + g * g * (1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    return p - (lr / bc1) * (m / denom), m, v


def single_mask_loss(pred_masks, target_masks) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """MaskLossHandler.compute_single_mask_loss (loss.py:166-188), the num_masks == 1 branch of LossModule.forward (:252-257):
    only the mask components, each reduced to its mean, no best-mask selection and no auxiliary MSE."""
    pred = torch.sigmoid(pred_masks.squeeze(1))
    parts = {"focal_loss": focal_loss(pred, target_masks).mean(), "iou_loss": iou_loss(pred, target_masks).mean()}
    return FOCAL_WEIGHT * parts["focal_loss"] + IOU_WEIGHT * parts["iou_loss"], parts
