"""Launch one kernel family a few times (for ncu captures):  python -m tools.run_op attn|gemm_o|conv|bench_small"""
import sys
import torch
from s3od_b200.engine import load_library

lib = load_library()
st = torch.cuda.current_stream().cuda_stream
which = sys.argv[1]
if which == "attn":
    B, H, ntok = 8, 12, 4101
    q = torch.randn(B * H, ntok, 64, device="cuda").bfloat16()
    k = torch.randn(B * H, ntok, 64, device="cuda").bfloat16()
    v = torch.randn(B * H, ntok, 64, device="cuda").bfloat16()
    out = torch.empty(B * ntok, H * 64, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        assert lib.s3od_op_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, H, ntok, st) == 0
elif which == "conv":
    B, h, w, cin, cout = 4, 256, 256, 256, 256
    x = torch.randn(B, h, w, cin, device="cuda").bfloat16()
    wp = torch.randn(cout, 9 * cin, device="cuda").bfloat16()
    y = torch.empty(B, h, w, cout, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        assert lib.s3od_op_conv3x3(x.data_ptr(), wp.data_ptr(), None, y.data_ptr(), B, h, w, cin, cout, 0, st) == 0
torch.cuda.synchronize()
print("ok")
