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
if which == "attn" and __import__("os").environ.get("S3OD_ATTN_TRACE"):
    import ctypes, numpy as np
    buf = np.zeros((64, 8), np.int64)
    lib.s3od_debug_attn_trace.argtypes = [ctypes.c_void_p]
    assert lib.s3od_debug_attn_trace(buf.ctypes.data) == 0
    t = buf[:33].astype(np.float64)
    names = ["S seen", "pass1 done", "PV(j-1) seen", "S released", "P published", "MMA: S(j+1) issued", "MMA: P seen", "MMA: PV issued"]
    base = t[:, 0:1]
    print("per-tile stamps relative to 'S seen' (cycles), tiles 5..30 averaged:")
    for i, n in enumerate(names):
        print(f"  {n:22s} {np.mean((t[5:30, i] - base[5:30, 0])):9.0f}")
    print("  tile period           ", np.mean(np.diff(t[5:30, 0])))
