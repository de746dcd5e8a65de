"""Two training steps (dinob, 1024^2, batch 4 unless given) for ncu captures:  python tools/train_one.py [S] [B]"""
import sys
import torch
sys.path.insert(0, "/root/repo")
from s3od_b200.arch import VITB
from s3od_b200.synth import synth_state_dict
from s3od_b200.training_head import TrainStep
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ts = TrainStep(synth_state_dict(VITB, 0), VITB, S, "cuda:0")
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(B, 3, S, S, device="cuda", generator=g)
m = (torch.rand(B, S, S, device="cuda", generator=g) > 0.5).float()
for _ in range(2):
    loss, _ = ts.step(x, m)
torch.cuda.synchronize()
print("loss", float(loss))
