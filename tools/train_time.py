import time, torch, sys
sys.path.insert(0, "/root/repo")
from s3od_b200.arch import VITB
from s3od_b200.synth import synth_state_dict
from s3od_b200.training_head import TrainStep
sd = synth_state_dict(VITB, 0)
for S, B in ((224, 8), (1024, 1), (1024, 2), (1024, 4), (1024, 8)):
    ts = TrainStep(sd, VITB, S, "cuda:0", lr=1e-5)
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(B, 3, S, S, device="cuda", generator=g)
    m = (torch.rand(B, S, S, device="cuda", generator=g) > 0.5).float()
    ts.step(x, m); torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    t0 = time.perf_counter()
    for _ in range(2):
        loss, _ = ts.step(x, m)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 2
    print(f"S={S} B={B}: {dt*1e3:.0f} ms per step = {B/dt:.2f} images/s, loss {float(loss):.4f}, peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)
    del ts
    torch.cuda.empty_cache()
