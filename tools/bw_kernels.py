"""Bandwidth of the pre / post-processing kernels alone (CUDA events, batch 32):  python -m tools.bw_kernels
Algorithmic bytes per image: preprocess = source u8 + bf16 model input; postprocess = fp32 logits + source RGB
+ fp32 all_masks + RGBA."""
import sys

import numpy as np
import torch

from s3od_b200.arch import VITB
from s3od_b200.engine import B200DPTSegmentation
from s3od_b200.synth import synth_noise_image, synth_state_dict

B, S = 32, 1024
sd = synth_state_dict(VITB, 0)
m = B200DPTSegmentation(sd, VITB, S, "cuda:0", max_batch=B, micro_batch=16)
peak = 6552.0
for src in (1024, 2048):
    imgs = [torch.from_numpy(synth_noise_image(src, src, seed=i)).cuda() for i in range(B)]
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
    b_pre = B * (src * src * 3 + 3 * S * S * 2)
    b_post = B * (3 * S * S * 4 + src * src * 3 + 3 * src * src * 4 + src * src * 4)
    print(f"source {src}: preprocess {t_pre * 1e3:.0f} us = {b_pre / t_pre / 1e6:.0f} GB/s ({b_pre / t_pre / 1e6 / peak:.0%} of {peak:.0f}); "
          f"postprocess {t_post * 1e3:.0f} us = {b_post / t_post / 1e6:.0f} GB/s ({b_post / t_post / 1e6 / peak:.0%})", flush=True)
m.close()
