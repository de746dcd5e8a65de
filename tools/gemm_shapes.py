import sys, torch
sys.path.insert(0, "/root/repo")
from s3od_b200.arch import VITB
from s3od_b200.synth import synth_state_dict
from s3od_b200 import training_head as th
orig = th._Ops.gemm
seen = []
def gemm(self, a, b, M, N, K):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = orig(self, a, b, M, N, K)
    e1.record(); torch.cuda.synchronize()
    seen.append((M, N, K, e0.elapsed_time(e1)))
    return out
th._Ops.gemm = gemm
ts = th.TrainStep(synth_state_dict(VITB, 0), VITB, 1024, "cuda:0")
x = torch.randn(4, 3, 1024, 1024, device="cuda"); m = (torch.rand(4, 1024, 1024, device="cuda") > 0.5).float()
ts.step(x, m); seen.clear(); ts.step(x, m)
for s in sorted(seen, key=lambda t: -t[3])[:14]: print("head gemm M=%d N=%d K=%d  %.3f ms  %.0f TFLOP/s" % (*s, 2e-9 * s[0] * s[1] * s[2] / s[3]))
