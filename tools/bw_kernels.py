"""Bandwidth of the pre / post-processing kernels alone (CUDA events, batch 32):  python -m tools.bw_kernels
Algorithmic bytes per image: preprocess = source u8 + bf16 model input; postprocess = fp32 logits of the cropped
region + source RGB + fp32 all_masks + RGBA.  Shapes: BASELINE.json configs[1] / [2] and the reference's fixtures
(/root/reference/tests/conftest.py:39-54: 480 x 640, 100 x 100, 2000 x 2000; tests/test_inference_package.py: 400 x 800)."""
import sys

import torch

from s3od_b200.arch import VITB
from s3od_b200.engine import B200DPTSegmentation
from s3od_b200.synth import synth_noise_image, synth_state_dict

B, S = 32, 1024
sd = synth_state_dict(VITB, 0)
m = B200DPTSegmentation(sd, VITB, S, "cuda:0", max_batch=B, micro_batch=16)
peak = 6552.0
KERNEL = {(1024, 1024): "copy / identity", (2048, 2048): "exact 2x box / exact 2x", (2000, 2000): "fixed-point bilinear / tile (x1.95)",
          (480, 640): "fixed-point bilinear (up) / per-pixel antialias (down x0.625)", (100, 100): "bilinear (up) / per-pixel antialias (down x0.098)",
          (400, 800): "bilinear (up) / per-pixel antialias (down x0.78)", (1536, 1536): "bilinear / tile (x1.5)"}
for (h, w) in ((1024, 1024), (2048, 2048), (2000, 2000), (1536, 1536), (480, 640), (400, 800), (100, 100)):
    imgs = [torch.from_numpy(synth_noise_image(h, w, seed=i)).cuda() for i in range(B)]
    masks = torch.randn(B, 3, S, S, device="cuda") * 3
    ious = torch.randn(B, 3, device="cuda")
    pads = m.preprocess(imgs)
    m.postprocess(masks, ious, imgs, pads, slot=0)
    torch.cuda.synchronize()

    def timed(fn, n=10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(60_000_000)          # ~30 ms of GPU work: the host enqueues all n calls behind it (kernel time, not host time)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    t_pre = timed(lambda: m.preprocess(imgs))
    t_post = timed(lambda: m.postprocess(masks, ious, imgs, pads, slot=0))
    ch, cw = S - 2 * pads[0]["height_pad"], S - 2 * pads[0]["width_pad"]
    b_pre = B * (h * w * 3 + 3 * S * S * 2)
    b_post = B * (3 * ch * cw * 4 + h * w * 3 + 3 * h * w * 4 + h * w * 4)
    print(f"source {h}x{w} [{KERNEL[(h, w)]}]: preprocess {t_pre * 1e3:.0f} us = {b_pre / t_pre / 1e6:.0f} GB/s ({b_pre / t_pre / 1e6 / peak:.0%} of {peak:.0f}); "
          f"postprocess {t_post * 1e3:.0f} us = {b_post / t_post / 1e6:.0f} GB/s ({b_post / t_post / 1e6 / peak:.0%})", flush=True)
m.close()
