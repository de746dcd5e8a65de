"""Which kernels run into the 1000 W cap?  Loops one op for ~3 s each and samples nvidia-smi (clocks.sm, power.draw) meanwhile.
    python -m tools.power_probe"""
import subprocess
import threading
import time

import torch

from s3od_b200.engine import load_library

lib = load_library()
st = torch.cuda.current_stream().cuda_stream


def sample(stop, rows):
    p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"],
                         stdout=subprocess.PIPE, text=True)
    while not stop.is_set():
        line = p.stdout.readline()
        if line:
            try:
                c, w = [float(x) for x in line.split(",")]
                rows.append((c, w))
            except ValueError:
                pass
    p.terminate()


def run(name, fn, flops_per_call, seconds=3.0):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    stop, rows = threading.Event(), []
    th = threading.Thread(target=sample, args=(stop, rows))
    th.start()
    time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t0 = time.perf_counter()
    e0.record()
    while time.perf_counter() - t0 < seconds:
        for _ in range(10):
            fn()
        n += 10
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    stop.set()
    th.join()
    rows = rows[3:] or rows
    clk = sorted(r[0] for r in rows)[len(rows) // 2] if rows else float("nan")
    pw = sorted(r[1] for r in rows)[len(rows) // 2] if rows else float("nan")
    print(f"{name:34s} {ms:8.3f} ms  {flops_per_call / ms / 1e9:8.1f} TFLOP/s   median {clk:6.0f} MHz  {pw:6.0f} W   ({len(rows)} samples)", flush=True)


B, H, ntok = 32, 12, 4101
q = (torch.randn(B * H, ntok, 64, device="cuda") * 0.3).bfloat16()
k = torch.randn(B * H, ntok, 64, device="cuda").bfloat16()
v = torch.randn(B * H, ntok, 64, device="cuda").bfloat16()
out = torch.empty(B * ntok, H * 64, device="cuda", dtype=torch.bfloat16)
run("attention (32 img, 12 heads)", lambda: lib.s3od_op_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, H, ntok, st),
    4.0 * B * H * ntok * ntok * 64)
M, N, K = 32 * 4101, 3072, 768
a = torch.randn(M, K, device="cuda").bfloat16()
b = torch.randn(N, K, device="cuda").bfloat16()
c = torch.empty(M, N, device="cuda")
run("GEMM 131232 x 3072 x 768 (fp32 out)", lambda: lib.s3od_op_gemm_f32(a.data_ptr(), b.data_ptr(), c.data_ptr(), M, N, K, st), 2.0 * M * N * K)
x = torch.randn(16, 256, 256, 256, device="cuda").bfloat16()
wp = torch.randn(256, 9 * 256, device="cuda").bfloat16()
y = torch.empty(16, 256, 256, 256, device="cuda", dtype=torch.bfloat16)
run("conv3x3 256->256 @256^2 x16", lambda: lib.s3od_op_conv3x3(x.data_ptr(), wp.data_ptr(), None, y.data_ptr(), 16, 256, 256, 256, 256, 0, st),
    2.0 * 16 * 256 * 256 * 256 * 9 * 256)
xa = torch.randn(8192, 8192, device="cuda").bfloat16()
run("torch.matmul bf16 8192^3 (cuBLAS)", lambda: torch.matmul(xa, xa), 2.0 * 8192 ** 3)
