"""Generate tests/golden/*.npz from the REAL reference (TEST INFRASTRUCTURE - see oracle/__init__.py).

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python -m oracle.make_golden            # writes tests/golden/{model_s64,model_s128,pipe_*,full_s1024}.npz

It imports `s3od.model.DPTSegmentation` and `s3od.BackgroundRemoval` from /root/reference/src
(unmodified), loads the seeded synthetic checkpoint from `s3od_b200.synth`, runs them on seeded
inputs on CPU fp32 and stores the outputs.  The oracle (oracle/model.py, oracle/prepost.py) and the
CUDA path are both tested against these files.
"""
import os
import sys
import tempfile

import numpy as np
import torch

REF_SRC = "/root/reference/src"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")

# (name, image_size, (H, W) of the source image, image seed)
PIPE_CASES = [
    ("pipe_s128_120x160", 128, (120, 160), 11),      # landscape, down-scale, height padding 16
    ("pipe_s128_160x120", 128, (160, 120), 12),      # portrait, width padding 16
    ("pipe_s64_128x128", 64, (128, 128), 13),        # exact 2x integer box down-scale, 2x mask up-sample
    ("pipe_s128_50x50", 128, (50, 50), 14),          # up-scale in, AA down-scale out
    ("pipe_s128_128x128", 128, (128, 128), 15),      # identity resize
]


def main():
    sys.path.insert(0, REF_SRC)
    sys.path.insert(0, ROOT)
    from s3od.model import DPTSegmentation            # the reference, unmodified
    from s3od.predictor import BackgroundRemoval
    from s3od_b200.arch import VITB
    from s3od_b200.synth import synth_state_dict, synth_image, save_checkpoint
    from oracle.prepost import preprocess

    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    sd = synth_state_dict(VITB, 0)
    model = DPTSegmentation(num_classes=1, num_outputs=3, encoder_name="dinov3_base", features=256,
                            use_bn=True, use_clstoken=False)
    model.load_state_dict(sd)
    model.eval()

    # ---- model-level fixtures: reference forward on the letterboxed synthetic image ----------------
    for S, B in ((64, 2), (128, 1)):
        xs = [preprocess(synth_image(S, S, seed=100 + S + b), S)[0] for b in range(B)]
        x = torch.from_numpy(np.concatenate(xs, 0))
        enc = model.encoder(pixel_values=x, output_hidden_states=True, return_dict=True)
        out = model(x)
        rec = dict(x=x.numpy(), pred_masks=out["pred_masks"].numpy(), pred_iou=out["pred_iou"].numpy(),
                   features_sub=out["features"][:, ::8, ::2, ::2].numpy(),
                   embed=enc.hidden_states[0].numpy(), layer0=enc.hidden_states[1].numpy())
        for i, t in enumerate(VITB.taps):
            rec[f"tap{i}"] = enc.hidden_states[t][:, 5:].numpy()
        np.savez(os.path.join(OUT, f"model_s{S}.npz"), **rec)
        print("wrote model", S, {k: v.shape for k, v in rec.items()})

    # ---- pipeline fixtures: BackgroundRemoval.remove_background end to end -------------------------
    ckpt = os.path.join(tempfile.gettempdir(), "s3od_synth_vitb_seed0.pt")
    save_checkpoint(ckpt, VITB, 0)
    for name, S, (H, W), seed in PIPE_CASES:
        br = BackgroundRemoval(model_id=ckpt, image_size=S, device="cpu")
        img = synth_image(H, W, seed=seed)
        res = br.remove_background(img)
        rgba = np.array(res.rgba_image)
        np.savez(os.path.join(OUT, name + ".npz"), image=img, all_masks=res.all_masks, all_ious=res.all_ious,
                 predicted_mask=res.predicted_mask, alpha=rgba[..., 3], image_size=np.int64(S))
        print("wrote", name, res.all_masks.shape, res.all_ious)

    # the reference's odd-padding failure: record the exception type for the parity test
    br = BackgroundRemoval(model_id=ckpt, image_size=128, device="cpu")
    try:
        br.remove_background(synth_image(300, 200, seed=1))
        err = "none"
    except Exception as e:                                    # noqa: BLE001 - recording whatever the reference raises
        err = type(e).__name__
    np.savez(os.path.join(OUT, "pipe_odd_padding.npz"), error=np.array(err))
    print("odd padding ->", err)

    # ---- full-size fixture (default image_size=1024): sub-sampled outputs --------------------------
    br = BackgroundRemoval(model_id=ckpt, image_size=1024, device="cpu")
    img = synth_image(1024, 1024, seed=7)
    x, _ = br._preprocess(img)
    out = br.model(x)
    res = br.remove_background(img)
    np.savez(os.path.join(OUT, "full_s1024.npz"),
             pred_iou=out["pred_iou"].numpy(), pred_masks_sub=out["pred_masks"][:, :, 5::16, 3::16].numpy(),
             all_ious=res.all_ious, all_masks_sub=res.all_masks[:, 5::16, 3::16],
             alpha_sub=np.array(res.rgba_image)[5::16, 3::16, 3], image_seed=np.int64(7))
    print("wrote full_s1024", out["pred_iou"].numpy())


if __name__ == "__main__":
    main()
