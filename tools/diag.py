"""GPU bring-up diagnostics: run one kernel family against a torch CUDA fp32 reference and print error metrics.

    python -m tools.diag gemm|ln|attn|conv|model|pipe|all

Each family prints a line per case; nothing asserts, so a single gpurun call shows every number.  (The gated
parity tests live in tests/; this is the tool used to find out *where* a mismatch starts.)
"""
import os
import sys
import time
import ctypes

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from s3od_b200.engine import load_library, B200DPTSegmentation  # noqa: E402
from s3od_b200.arch import VITB  # noqa: E402


def _st():
    return torch.cuda.current_stream().cuda_stream


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-12)), float((a - b).abs().max())


def diag_gemm():
    lib = load_library()
    g = torch.Generator(device="cuda").manual_seed(0)
    for (M, N, K) in [(128, 128, 64), (128, 128, 256), (256, 256, 128), (300, 256, 768), (4101, 768, 768), (8202, 3072, 768)]:
        a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
        b = torch.randn(N, K, device="cuda", generator=g).bfloat16()
        c = torch.full((M, N), float("nan"), device="cuda")
        rc = lib.s3od_op_gemm_f32(a.data_ptr(), b.data_ptr(), c.data_ptr(), M, N, K, _st())
        torch.cuda.synchronize()
        ref = a.float() @ b.float().t()
        r, m = rel(c, ref)
        print(f"gemm M={M} N={N} K={K} rc={rc} rel={r:.3e} maxabs={m:.3e} nan={int(torch.isnan(c).sum())}", flush=True)
        if r > 1e-2 and M <= 256:
            d = (c - ref).abs()
            bad = (d > 1e-2 * ref.abs().max()).nonzero()
            print("   first bad idx", bad[:8].tolist(), "c", c[0, :4].tolist(), "ref", ref[0, :4].tolist())
    # timing of the big one
    M, N, K = 32808, 3072, 768
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    b = torch.randn(N, K, device="cuda", generator=g).bfloat16()
    c = torch.empty(M, N, device="cuda")
    for _ in range(3):
        lib.s3od_op_gemm_f32(a.data_ptr(), b.data_ptr(), c.data_ptr(), M, N, K, _st())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        lib.s3od_op_gemm_f32(a.data_ptr(), b.data_ptr(), c.data_ptr(), M, N, K, _st())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"gemm {M}x{N}x{K}: {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s (fp32-store epilogue)", flush=True)


def diag_ln():
    lib = load_library()
    for (M, D) in [(37, 768), (4101, 768), (4101, 1024)]:
        x = torch.randn(M, D, device="cuda") * 3 + 0.5
        w = torch.randn(D, device="cuda")
        b = torch.randn(D, device="cuda")
        y = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
        rc = lib.s3od_op_layernorm(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), M, D, 1e-5, _st())
        torch.cuda.synchronize()
        r, m = rel(y, F.layer_norm(x, (D,), w, b, 1e-5))
        print(f"layernorm M={M} D={D} rc={rc} rel={r:.3e} maxabs={m:.3e}", flush=True)


def diag_attn():
    lib = load_library()
    g = torch.Generator(device="cuda").manual_seed(1)
    for (B, H, ntok) in [(1, 1, 21), (1, 2, 128), (2, 3, 261), (1, 12, 4101)]:
        q = (torch.randn(B * H, ntok, 64, device="cuda", generator=g) * 1.5).bfloat16()
        k = (torch.randn(B * H, ntok, 64, device="cuda", generator=g) * 1.5).bfloat16()
        v = torch.randn(B * H, ntok, 64, device="cuda", generator=g).bfloat16()
        qs = (q.float() * (0.125 * 1.4426950408889634)).bfloat16()
        out = torch.full((B * ntok, H * 64), float("nan"), device="cuda", dtype=torch.bfloat16)
        rc = lib.s3od_op_attention(qs.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, H, ntok, _st())
        torch.cuda.synchronize()
        s = (qs.float() @ k.float().transpose(1, 2)) * 0.6931471805599453
        ref = torch.softmax(s, -1) @ v.float()
        ref = ref.reshape(B, H, ntok, 64).permute(0, 2, 1, 3).reshape(B * ntok, H * 64)
        r, m = rel(out, ref)
        print(f"attention B={B} H={H} ntok={ntok} rc={rc} rel={r:.3e} maxabs={m:.3e} nan={int(torch.isnan(out.float()).sum())}", flush=True)
    B, H, ntok = 8, 12, 4101
    q = torch.randn(B * H, ntok, 64, device="cuda").bfloat16()
    k = torch.randn(B * H, ntok, 64, device="cuda").bfloat16()
    v = torch.randn(B * H, ntok, 64, device="cuda").bfloat16()
    out = torch.empty(B * ntok, H * 64, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        lib.s3od_op_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, H, ntok, _st())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        lib.s3od_op_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, H, ntok, _st())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"attention B=8 H=12 ntok=4101: {ms:.3f} ms  {4 * B * H * ntok * ntok * 64 / ms / 1e9:.1f} TFLOP/s", flush=True)


def diag_conv():
    lib = load_library()
    g = torch.Generator(device="cuda").manual_seed(2)
    for (B, h, w, cin, cout) in [(1, 8, 16, 64, 256), (1, 4, 4, 256, 256), (2, 32, 32, 256, 256), (1, 20, 24, 128, 256), (1, 64, 64, 1024, 256)]:
        x = torch.randn(B, h, w, cin, device="cuda", generator=g).bfloat16()
        wt = (torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (3 * cin ** 0.5)).bfloat16()
        bias = torch.randn(cout, device="cuda", generator=g)
        wp = wt.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()
        y = torch.full((B, h, w, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
        rc = lib.s3od_op_conv3x3(x.data_ptr(), wp.data_ptr(), bias.data_ptr(), y.data_ptr(), B, h, w, cin, cout, 0, _st())
        torch.cuda.synchronize()
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, padding=1).permute(0, 2, 3, 1)
        r, m = rel(y, ref)
        print(f"conv3x3 B={B} {h}x{w} cin={cin} cout={cout} rc={rc} rel={r:.3e} maxabs={m:.3e} nan={int(torch.isnan(y.float()).sum())}", flush=True)
    B, h, w, cin, cout = 4, 256, 256, 256, 256
    x = torch.randn(B, h, w, cin, device="cuda").bfloat16()
    wp = torch.randn(cout, 9 * cin, device="cuda").bfloat16()
    y = torch.empty(B, h, w, cout, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        lib.s3od_op_conv3x3(x.data_ptr(), wp.data_ptr(), None, y.data_ptr(), B, h, w, cin, cout, 0, _st())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        lib.s3od_op_conv3x3(x.data_ptr(), wp.data_ptr(), None, y.data_ptr(), B, h, w, cin, cout, 0, _st())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"conv3x3 B=4 256x256 256->256: {ms:.3f} ms  {2 * B * h * w * cout * 9 * cin / ms / 1e9:.1f} TFLOP/s", flush=True)


def diag_convrows():
    lib = load_library()
    g = torch.Generator(device="cuda").manual_seed(3)
    for mode in (0,):
        for (B, h, w) in [(1, 1, 128), (1, 5, 128), (2, 70, 256), (1, 130, 384)]:
            x = torch.randn(B, h, w, 64, device="cuda", generator=g).bfloat16()
            wt = (torch.randn(64, 64, 3, 3, device="cuda", generator=g) / 24).bfloat16()
            bias = torch.randn(64, device="cuda", generator=g)
            wp = wt.permute(0, 2, 3, 1).reshape(64, 9 * 64).contiguous()
            y = torch.full((B, h, w, 64), float("nan"), device="cuda", dtype=torch.bfloat16)
            rc = lib.s3od_op_conv3x3_rows(x.data_ptr(), wp.data_ptr(), bias.data_ptr(), y.data_ptr(), B, h, w, 1, _st())
            torch.cuda.synchronize()
            ref = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, padding=1)).permute(0, 2, 3, 1)
            r, m = rel(y, ref)
            print(f"conv_rows base_offset={mode} B={B} {h}x{w} rc={rc} rel={r:.3e} maxabs={m:.3e} nan={int(torch.isnan(y.float()).sum())}", flush=True)
    B, h, w = 16, 1024, 1024
    x = torch.randn(B, h, w, 64, device="cuda").bfloat16()
    wp = torch.randn(64, 9 * 64, device="cuda").bfloat16()
    y = torch.empty(B, h, w, 64, device="cuda", dtype=torch.bfloat16)
    for mode in (1,):
        for _ in range(2):
            lib.s3od_op_conv3x3_rows(x.data_ptr(), wp.data_ptr(), None, y.data_ptr(), B, h, w, 1, _st())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            lib.s3od_op_conv3x3_rows(x.data_ptr(), wp.data_ptr(), None, y.data_ptr(), B, h, w, 1, _st())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"conv_rows B=16 1024x1024 64->64: {ms:.3f} ms  {2 * B * h * w * 64 * 9 * 64 / ms / 1e9:.1f} TFLOP/s", flush=True)


def diag_convt():
    from s3od_b200.weights import convt_rows_weights
    lib = load_library()
    g = torch.Generator(device="cuda").manual_seed(4)
    for (B, h, w) in [(1, 1, 128), (1, 3, 128), (2, 40, 256), (1, 70, 128)]:
        x = torch.randn(B, h, w, 128, device="cuda", generator=g).bfloat16()
        wt = (torch.randn(128, 64, 4, 4, device="cuda", generator=g) / 32).bfloat16()
        bias = torch.randn(64, device="cuda", generator=g)
        wr = convt_rows_weights(wt).contiguous()
        y = torch.full((B, 2 * h, 2 * w, 64), float("nan"), device="cuda", dtype=torch.bfloat16)
        rc = lib.s3od_op_convt_rows(x.data_ptr(), wr.data_ptr(), bias.data_ptr(), y.data_ptr(), B, h, w, 1, _st())
        torch.cuda.synchronize()
        ref = F.relu(F.conv_transpose2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, stride=2, padding=1)).permute(0, 2, 3, 1)
        r, m = rel(y, ref)
        print(f"convt_rows B={B} {h}x{w} rc={rc} rel={r:.3e} maxabs={m:.3e} nan={int(torch.isnan(y.float()).sum())}", flush=True)
    B, h, w = 16, 512, 512
    x = torch.randn(B, h, w, 128, device="cuda").bfloat16()
    wr = torch.randn(1024, 128, device="cuda").bfloat16()
    y = torch.empty(B, 2 * h, 2 * w, 64, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        lib.s3od_op_convt_rows(x.data_ptr(), wr.data_ptr(), None, y.data_ptr(), B, h, w, 1, _st())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        lib.s3od_op_convt_rows(x.data_ptr(), wr.data_ptr(), None, y.data_ptr(), B, h, w, 1, _st())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"convt_rows B=16 512x512 128->64 (both phases): {ms:.3f} ms  {2 * B * h * w * 128 * 16 * 64 / ms / 1e9:.1f} TFLOP/s", flush=True)


def diag_model():
    from s3od_b200.synth import synth_state_dict
    from oracle import model as om
    sd = synth_state_dict(VITB, 0)
    for S in (64, 128):
        gold = np.load(os.path.join(ROOT, "tests", "golden", f"model_s{S}.npz"))
        x = torch.from_numpy(gold["x"])
        B = x.shape[0]
        m = B200DPTSegmentation(sd, VITB, S, "cuda:0", max_batch=B, micro_batch=B)
        out = m(x.cuda())
        torch.cuda.synchronize()
        st = {}
        ref = om.forward(sd, x, VITB, st)
        g_ = S // 16
        P, ntok = g_ * g_, g_ * g_ + 5
        for j in range(4):
            t = m.stage(f"tap{j}", torch.bfloat16, (B, P, 768)).float().cpu()
            print(f"S={S} tap{j}: rel/max", rel(t, torch.from_numpy(gold[f'tap{j}'])), flush=True)
        for name, key, C, R in (("l1", "l1", 256, 4 * g_), ("l2", "l2", 256, 2 * g_), ("l3", "l3", 256, g_), ("l4", "l4", 256, g_ // 2),
                                ("p4", "p4", 256, g_), ("p3", "p3", 256, 2 * g_), ("p2", "p2", 256, 4 * g_), ("p1", "p1", 256, 8 * g_),
                                ("feat", "feat", 64, S)):
            t = m.stage(name, torch.bfloat16, (B, R, R, C)).float().cpu().permute(0, 3, 1, 2)
            print(f"S={S} {name}: rel/max", rel(t, st[key]), flush=True)
        pm, pi = out["pred_masks"].cpu(), out["pred_iou"].cpu()
        gm, gi = torch.from_numpy(gold["pred_masks"]), torch.from_numpy(gold["pred_iou"])
        print(f"S={S} pred_masks rel/max", rel(pm, gm), "sigmoid maxabs", float((torch.sigmoid(pm) - torch.sigmoid(gm)).abs().max()),
              "mean", float((torch.sigmoid(pm) - torch.sigmoid(gm)).abs().mean()), flush=True)
        a, b = pm > 0, gm > 0
        print(f"S={S} thresholded IoU", float((a & b).sum()) / max(1.0, float((a | b).sum())), "pred_iou", pi.tolist(), "gold", gi.tolist(), flush=True)
        m.close()


def diag_pipe():
    from s3od_b200.synth import save_checkpoint
    from s3od_b200 import BackgroundRemoval
    ck = "/tmp/s3od_synth_vitb_seed0.pt"
    if not os.path.exists(ck):
        save_checkpoint(ck, VITB, 0)
    cache = {}
    for name in ["pipe_s128_120x160", "pipe_s128_160x120", "pipe_s64_128x128", "pipe_s128_50x50", "pipe_s128_128x128"]:
        gold = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        S = int(gold["image_size"])
        if S not in cache:
            cache[S] = BackgroundRemoval(model_id=ck, image_size=S, device="cuda:0")
        res = cache[S].remove_background(gold["image"])
        d = np.abs(res.all_masks - gold["all_masks"])
        da = np.abs(np.array(res.rgba_image)[..., 3].astype(int) - gold["alpha"].astype(int))
        print(f"{name}: all_masks maxabs={d.max():.4f} mean={d.mean():.5f} ious={res.all_ious.tolist()} gold={gold['all_ious'].tolist()} "
              f"alpha maxdiff={da.max()} rgb_equal={np.array_equal(np.array(res.rgba_image)[..., :3], gold['image'])}", flush=True)


FAMILIES = {"gemm": diag_gemm, "ln": diag_ln, "attn": diag_attn, "conv": diag_conv, "model": diag_model, "pipe": diag_pipe}

def diag_step():
    """Time preprocess / forward / postprocess of one batch separately (CUDA events), with and without per-op events."""
    from s3od_b200.synth import synth_state_dict, synth_noise_image
    sd = synth_state_dict(VITB, 0)
    B, S = 16, 1024
    m = B200DPTSegmentation(sd, VITB, S, "cuda:0", max_batch=B, micro_batch=8)
    imgs = [torch.from_numpy(synth_noise_image(S, S, seed=i)).cuda() for i in range(B)]
    for prof in (False, True, False):
        m.profile_enable(prof)
        for it in range(3):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            t0 = time.perf_counter()
            ev[0].record()
            pads = m.preprocess(imgs)
            ev[1].record()
            out = m._forward_staged(B)
            ev[2].record()
            t1 = time.perf_counter()
            res = m.postprocess(out["pred_masks"], out["pred_iou"], imgs, pads)
            ev[3].record()
            t2 = time.perf_counter()
            torch.cuda.synchronize()
            t3 = time.perf_counter()
            print(f"profile={prof} it={it}: pre {ev[0].elapsed_time(ev[1]):.2f} ms, fwd {ev[1].elapsed_time(ev[2]):.2f} ms, "
                  f"post {ev[2].elapsed_time(ev[3]):.2f} ms | host: launch fwd {1e3 * (t1 - t0):.2f} ms, post {1e3 * (t2 - t1):.2f} ms, "
                  f"sync {1e3 * (t3 - t2):.2f} ms", flush=True)
        if prof:
            rows = m.profile_read()
            print("   per-op sum", sum(r[3] for r in rows) / 3, "ms per iteration")
    m.close()


FAMILIES["step"] = diag_step
FAMILIES["convrows"] = diag_convrows
FAMILIES["convt"] = diag_convt


def diag_stall():
    """Find intermittent host stalls: time torch.empty and the library calls separately over many iterations."""
    from s3od_b200.synth import synth_state_dict, synth_noise_image
    from s3od_b200.engine import _check, _stream_ptr
    sd = synth_state_dict(VITB, 0)
    B, S = 16, 1024
    m = B200DPTSegmentation(sd, VITB, S, "cuda:0", max_batch=B, micro_batch=8)
    imgs = [torch.from_numpy(synth_noise_image(S, S, seed=i)).cuda() for i in range(B)]
    pads = m.preprocess(imgs)
    keep = None
    for it in range(12):
        t0 = time.perf_counter()
        masks = torch.empty((B, 3, S, S), dtype=torch.float32, device="cuda")
        ious = torch.empty((B, 3), dtype=torch.float32, device="cuda")
        t1 = time.perf_counter()
        _check(m.lib, m.lib.s3od_forward(m._ctx, B, masks.data_ptr(), ious.data_ptr(), _stream_ptr(m.device)), "fwd")
        t2 = time.perf_counter()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        keep = (masks, ious) if it % 2 == 0 else keep
        print(f"it={it}: alloc {1e3 * (t1 - t0):.2f} ms, forward call {1e3 * (t2 - t1):.2f} ms, sync {1e3 * (t3 - t2):.2f} ms", flush=True)
    m.close()


FAMILIES["stall"] = diag_stall


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    torch.manual_seed(0)
    print("device:", torch.cuda.get_device_name(0), "lib:", load_library().s3od_version().decode(), flush=True)
    for name, fn in FAMILIES.items():
        if which in ("all", name):
            t = time.time()
            try:
                fn()
            except Exception as e:  # noqa: BLE001 - report and continue with the next family
                print(f"[{name}] EXCEPTION {type(e).__name__}: {e}", flush=True)
            print(f"[{name}] done in {time.time() - t:.1f}s", flush=True)
