"""Generate tests/golden/loss.npz from the REAL reference loss (TEST INFRASTRUCTURE - see oracle/__init__.py).

Run in the build container only:   python -m oracle.make_golden_loss
Loads /root/reference/synth_sod/src/synth_sod/model_training/loss.py (unmodified) by path - `hydra` is not installed, so an
empty stub module satisfies its import and the LossComponents are built directly with the focal_iou.yaml values
(config/loss/focal_iou.yaml:1-27), exactly as LossComponent.from_dict would (add_sigmoid stays at its default True, F10).
Records inputs, the total loss, every loss part and the autograd gradients w.r.t. pred_masks / pred_iou; also one
torch.optim.AdamW step (lightning_module.py:183-193 hyper-parameters) for the fused optimiser kernel."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/synth_sod/src/synth_sod/model_training/loss.py"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_reference_loss():
    sys.modules.setdefault("hydra", types.ModuleType("hydra"))
    spec = importlib.util.spec_from_file_location("ref_loss", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    return ref


def build_loss_module(ref):
    comps = [ref.LossComponent("focal_loss", 20, "masks", "pred_masks", ref.FocalLoss(reduction="none")),
             ref.LossComponent("iou_loss", 1.0, "masks", "pred_masks", ref.IoULoss(smooth=1e-6, reduction="none")),
             ref.LossComponent("mse_ious_loss", 0.05, "gt_ious", "pred_iou", torch.nn.MSELoss())]
    lm = ref.LossModule.__new__(ref.LossModule)
    torch.nn.Module.__init__(lm)
    lm.components = comps
    lm.mask_handler = ref.MaskLossHandler(comps, full_mask_lambda=0.1, decay_rate=0.2)
    return lm


def cases():
    g = torch.Generator().manual_seed(11)
    out = []
    for name, B, K, H, W, epoch in (("b2_32", 2, 3, 32, 32, 0), ("b3_48x40", 3, 3, 48, 40, 7), ("b1_64", 1, 3, 64, 64, 30)):
        yy, xx = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing="ij")
        masks = torch.stack([((((yy - (0.3 + 0.2 * b) * H) / (0.3 * H)) ** 2 + ((xx - 0.5 * W) / (0.28 * W)) ** 2) < 1).float()
                             for b in range(B)])
        masks = (masks * (0.85 + 0.15 * torch.rand(B, H, W, generator=g))).clamp(0, 1)          # soft targets in [0, 1]
        logits = 3.0 * torch.randn(B, K, H, W, generator=g) + 2.0 * (masks.unsqueeze(1) - 0.5) * torch.tensor([3.0, 1.0, -1.0]).view(1, K, 1, 1)
        iou_logits = torch.randn(B, K, generator=g)
        out.append((name, logits, iou_logits, masks, epoch))
    return out


def main():
    ref = load_reference_loss()
    lm = build_loss_module(ref)
    rec, names = {}, []
    for name, logits, iou_logits, masks, epoch in cases():
        z = logits.clone().requires_grad_(True)
        q = iou_logits.clone().requires_grad_(True)
        loss, parts = lm({"pred_masks": z, "pred_iou": q}, {"masks": masks}, epoch)
        loss.backward()
        names.append(name)
        rec[name + "_logits"], rec[name + "_iou_logits"], rec[name + "_masks"] = logits.numpy(), iou_logits.numpy(), masks.numpy()
        rec[name + "_epoch"] = np.int64(epoch)
        rec[name + "_loss"] = np.float64(loss.item())
        rec[name + "_grad_logits"], rec[name + "_grad_iou"] = z.grad.numpy(), q.grad.numpy()
        for k in ("best_iou", "gt_ious", "focal_loss_best", "focal_loss_full", "iou_loss_best", "iou_loss_full", "mse_ious_loss"):
            rec[name + "_" + k] = np.float64(parts[k].item())              # LossModule.forward returns the means (loss.py:274-275)
        print(name, float(loss), {k: float(v) for k, v in parts.items()})
    # num_masks == 1 branch (LossModule.forward -> compute_single_mask_loss, loss.py:252-257 / 166-188): dinol.yaml's configuration
    name, logits, _, masks, epoch = cases()[1]
    z1 = logits[:, :1].clone().requires_grad_(True)
    loss1, parts1 = lm({"pred_masks": z1}, {"masks": masks}, epoch)
    loss1.backward()
    rec["single_logits"], rec["single_masks"] = logits[:, :1].numpy(), masks.numpy()
    rec["single_loss"], rec["single_grad_logits"] = np.float64(loss1.item()), z1.grad.numpy()
    rec["single_focal_loss"], rec["single_iou_loss"] = np.float64(parts1["focal_loss"].item()), np.float64(parts1["iou_loss"].item())
    print("single", float(loss1.detach()), {k: float(v.detach()) for k, v in parts1.items()})
    # one AdamW step with the reference's hyper-parameters on a small tensor pair (encoder group lr, head group lr x 10)
    g = torch.Generator().manual_seed(3)
    for tag, lr in (("enc", 1e-5), ("head", 1e-4)):
        p = torch.nn.Parameter(torch.randn(1000, generator=g))
        opt = torch.optim.AdamW([p], lr=lr, weight_decay=0.05, betas=(0.9, 0.999), eps=1e-8)
        rec[f"adamw_{tag}_p0"] = p.detach().clone().numpy()
        grads = []
        for step in range(3):
            p.grad = torch.randn(1000, generator=g) * (0.1 if step else 1.0)
            grads.append(p.grad.clone().numpy())
            opt.step()
            rec[f"adamw_{tag}_p{step + 1}"] = p.detach().clone().numpy()
        rec[f"adamw_{tag}_grads"] = np.stack(grads)
    rec["names"] = np.array(names)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "loss.npz"), **rec)


if __name__ == "__main__":
    main()
