"""Experiment: two micro-batches in flight on two streams (two contexts) against one stream.  python -m tools.two_stream"""
import torch

from s3od_b200.arch import VITB
from s3od_b200.engine import B200DPTSegmentation
from s3od_b200.synth import synth_noise_image, synth_state_dict

B, S, MB = 32, 1024, 16
sd = synth_state_dict(VITB, 0)
imgs = [torch.from_numpy(synth_noise_image(S, S, seed=i)).cuda() for i in range(B)]


def timed(fn, n=6, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


one = B200DPTSegmentation(sd, VITB, S, "cuda:0", max_batch=B, micro_batch=MB)
ms = timed(lambda: one.run_u8(imgs, slot=0))
print(f"one stream, batch {B}, micro-batch {MB}: {ms:.2f} ms = {B / ms * 1e3:.1f} images/s", flush=True)
one.close()
del one
for mb in (16, 8):
    half = B // 2
    ms_ = [B200DPTSegmentation(sd, VITB, S, "cuda:0", max_batch=half, micro_batch=mb) for _ in range(2)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    main = torch.cuda.current_stream()

    def step():
        for k in range(2):
            streams[k].wait_stream(main)
            with torch.cuda.stream(streams[k]):
                ms_[k].run_u8(imgs[k * half:(k + 1) * half], slot=0)
        for k in range(2):
            main.wait_stream(streams[k])
    t = timed(step)
    print(f"two streams, 2 x batch {half}, micro-batch {mb}: {t:.2f} ms = {B / t * 1e3:.1f} images/s", flush=True)
    for m in ms_:
        m.close()
    del ms_
