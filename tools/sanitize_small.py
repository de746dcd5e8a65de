"""Small end-to-end run for compute-sanitizer (memcheck): S=64 model, a few shapes, kernel-level ops at small sizes."""
import numpy as np
import torch
from s3od_b200 import BackgroundRemoval
from s3od_b200.arch import VITB
from s3od_b200.synth import save_checkpoint, synth_image

ck = "/tmp/san_ck.pt"
save_checkpoint(ck, VITB, 0)
br = BackgroundRemoval(model_id=ck, image_size=64, device="cuda:0", max_batch=3, micro_batch=2)
imgs = [synth_image(128, 128, 1), synth_image(48, 64, 2), synth_image(37, 91, 3)]
try:
    res = br.remove_background_batch(imgs[:2] + [synth_image(64, 64, 4)])
    print("ok", [r.all_masks.shape for r in res])
except Exception as e:  # noqa: BLE001
    print("exception", e)
torch.cuda.synchronize()
