"""Host-side profile of BackgroundRemoval.remove_background_batch (cProfile) on the GPU box."""
import cProfile
import io
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from s3od_b200 import BackgroundRemoval
from s3od_b200.arch import VITB
from s3od_b200.synth import save_checkpoint, synth_noise_image

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
MB = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ck = "/tmp/prof_ck.pt"
save_checkpoint(ck, VITB, 0)
br = BackgroundRemoval(model_id=ck, image_size=1024, device="cuda:0", max_batch=B, micro_batch=MB)
imgs = []
for i in range(B):
    t = torch.empty((1024, 1024, 3), dtype=torch.uint8, pin_memory=True)
    t.numpy()[...] = synth_noise_image(1024, 1024, seed=i)
    imgs.append(t.numpy())
for i in range(8):
    t1 = time.perf_counter()
    res = br.remove_background_batch(imgs)
    torch.cuda.synchronize()
    print(f"call {i}: {1e3 * (time.perf_counter() - t1):.1f} ms")
t0 = time.perf_counter()
for _ in range(3):
    res = br.remove_background_batch(imgs)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
print(f"e2e: {dt * 1e3:.1f} ms per batch of {B} -> {B / dt:.1f} img/s")
pr = cProfile.Profile()
pr.enable()
res = br.remove_background_batch(imgs)
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22)
print(s.getvalue()[:3500])
