"""Pre / post-process kernels under compute-sanitizer (memcheck): every kernel variant at small sizes, ragged tiles.
    compute-sanitizer --tool memcheck python -m tools.sanitize_prepost"""
import numpy as np
import torch

from s3od_b200.arch import VITB
from s3od_b200.engine import B200DPTSegmentation
from s3od_b200.synth import synth_noise_image, synth_state_dict

sd = synth_state_dict(VITB, 0)
for S in (64, 128):
    m = B200DPTSegmentation(sd, VITB, S, "cuda:0", max_batch=4, micro_batch=2)
    # (H, W, pad_h, pad_w): identity, exact 2x, other up-sampling ratios (ragged 512-wide tiles), down-scaling, odd widths
    cases = [(S, S, 0, 0), (2 * S, 2 * S, 0, 0), (3 * S + 4, 5 * S + 8, 0, 0), (1100, 604, 0, 8), (S // 2, S // 2, 0, 0), (37, 91, 0, 0),
             (S - 32, S, 16, 0), (2 * S - 64, 2 * S, 16, 0)]
    for H, W, hp, wp in cases:
        img = torch.from_numpy(synth_noise_image(H, W, seed=H + W)).cuda()
        logits = torch.randn(1, 3, S, S, device="cuda") * 3
        ious = torch.randn(1, 3, device="cuda")
        pad = dict(height_pad=hp, width_pad=wp, original_size=(H, W), resized_size=(S - 2 * hp, S - 2 * wp))
        outs, _, _ = m.postprocess(logits, ious, [img], [pad])
        torch.cuda.synchronize()
        assert bool(torch.isfinite(outs[0][0]).all())
    for h, w in [(S, S), (2 * S, 2 * S), (50, 70), (3 * S, 2 * S), (S // 2, S)]:
        try:
            m.preprocess([torch.from_numpy(synth_noise_image(h, w, seed=h)).cuda()])
        except ValueError:
            pass                                   # odd padding, like the reference
        torch.cuda.synchronize()
    m.close()
print("sanitize_prepost done")
