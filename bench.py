"""Benchmark of the S3OD `remove_background` hot path (dinob) - see the contract in DESIGN.md section "Measurement".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--batch B] [--source 1024|2048]

A step = one pass of the hot path over one batch of synthetic uint8 images per GPU:
preprocess kernel -> ViT-B + DPT head forward -> postprocess kernel, all inputs resident in HBM (`value`), and the
same through the public API with host buffers (`e2e`).  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from s3od_b200 import sharder                       # noqa: E402
from s3od_b200.arch import VITB, VITL               # noqa: E402
from s3od_b200.synth import save_checkpoint, synth_noise_image, synth_state_dict   # noqa: E402

METRIC = "images/sec (dinob, device-timed)"
UNIT = "images/s"
TRAIN_BATCH = 4            # images per GPU and step of the training legs (train_batch_size of the reference's model_training/config/dataset/synth.yaml:6)
# algorithmic FLOPs of the needed layers, dinob @ 1024^2 = 2276.8 GF as written; the plan runs the four
# fusion-block out_convs (1x1) BEFORE the 2x interpolation (engine.cu, "out_conv commuted"), which removes 34.2 GF of them
# ("if used, subtract 34.2 GF") - the work actually executed is what the roofline is computed from.
OUT_CONV_COMMUTE_GFLOP = 34.2
GFLOP_PER_IMAGE = 2276.8 - OUT_CONV_COMMUTE_GFLOP          # 2242.6
GFLOP_PER_IMAGE_VITL = 4958.1 - OUT_CONV_COMMUTE_GFLOP     # ViT-L @ 1024^2, one mask


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], tflops_burst=p["bf16_tflops"], tflops_sustained=p["bf16_tflops_sustained"], source="measured")
    return dict(hbm_gbs=6650.0, tflops_burst=1590.0, tflops_sustained=1400.0, source="fallback")


def newest_traffic(kernel: str, micro_batch: int, model: str, image_size: int):
    """dram bytes per launch of `kernel` from the NEWEST committed ncu --set full summary (profiles/r*_traffic.json, highest
    round tag first) whose capture matches this run's micro-batch / model / image size; (None, None) if there is none."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")), reverse=True):
        try:
            with open(path) as f:
                t = json.load(f).get(kernel)
            if t and t["micro_batch"] == micro_batch and t["model"] == model and t["image_size"] == image_size:
                return t["dram_bytes_per_launch"], os.path.basename(path)
        except (OSError, ValueError, KeyError, TypeError):
            continue
    return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, watts = [], None, set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                watts.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "power_w": float(np.median(watts)) if watts else None}       # the sustained step sits at the board's power cap


def kernel_family(label: str) -> str:
    if label in ("preprocess", "postprocess"):
        return label
    if label.endswith("attention"):
        return "attention"
    if label.endswith((".ln1", ".ln2", "final_residual")):
        return "layernorm"
    if label.startswith("enc.") or label == "patch_embed":
        return "encoder_gemm"
    if label.endswith("upsample") or label in ("prefix_tokens", "head.iou"):
        return "head_bandwidth"
    if label in ("head.mh.c2", "head.mh.heads") or label.startswith("head.mh.up"):
        return "head_conv_smallN"
    return "head_conv"


# algorithmic GFLOP per image of each family (SURVEY 8(d) breakdown, dinob @ 1024^2)
FAMILY_GFLOP = {
    "attention": 11 * 51.67,
    "encoder_gemm": 4.83 + 11 * (19.35 + 38.70),
    "head_conv": 17.7 + 36.5 + 140.1 + 2.95 + 21.47 + 85.90 + 343.60 + 154.62 - OUT_CONV_COMMUTE_GFLOP,
    "head_conv_smallN": 68.72 + 77.31 + 116.17,
}


def config_dict(args):
    """The workload both arms (--impl b200 / reference) are quoted on: BASELINE.json configs[1] unless overridden."""
    B, S, src = args.batch, args.image_size, args.source
    return {"workload": f"{args.model} inference, batch {B} synthetic {src}x{src} uint8 images per GPU, image_size {S} "
                        "(preprocess + backbone + mask decoder + IoU head + postprocess), seeded random weights",
            "batch_per_gpu": B, "image_size": S, "source": src, "micro_batch": min(args.micro_batch, B),
            "cache": "inputs + activations per step exceed the 126 MB L2 by >10x (no flush needed)"}


def make_images(B, src, first_seed, dev):
    """Seeded synthetic uint8 images (reference fixture style, tests/conftest.py:39-54): ordinary PAGEABLE numpy arrays for
    the public API - what /root/reference/src/s3od/predictor.py:96-106 receives - and device copies for the device-timed leg."""
    np_imgs = [synth_noise_image(src, src, seed=first_seed + i) for i in range(B)]
    d_imgs = [torch.from_numpy(a).to(dev) for a in np_imgs]
    return np_imgs, d_imgs


def time_device(model, d_imgs, steps, warmup, dev, sampler=None, profile=False):
    """`steps` passes of preprocess -> forward -> postprocess over the device-resident batch, CUDA events on the launch stream."""
    for _ in range(warmup):
        model.run_u8(d_imgs, slot=0)                 # reusable output buffers: no allocator traffic in the timed loop
    torch.cuda.synchronize(dev)
    if profile:
        model.profile_enable(True)
    sharder.barrier()
    torch.cuda.synchronize(dev)
    if sampler is not None:
        sampler.start()
    launches0 = model.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = model.run_u8(d_imgs, slot=0)
    e1.record()
    torch.cuda.synchronize(dev)
    sharder.barrier()
    ms_local = e0.elapsed_time(e1)
    launches = model.launch_count() - launches0
    prof = model.profile_read() if profile else None
    if profile:
        model.profile_enable(False)
    clocks = sampler.stop() if sampler is not None else None
    return sharder.max_over_ranks(ms_local, device=dev), launches, prof, clocks, out


def time_e2e(br, np_imgs, steps, warmup, dev):
    """The same metric through the public API: pageable host uint8 in (pinned on ingest through the predictor's staging ring),
    host results out; H2D and D2H inside the timed region, wall clock around the calls."""
    res = None
    for _ in range(warmup):
        res = br.remove_background_batch(np_imgs)      # same binding pattern as the timed loop (two result sets alive)
    sharder.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(steps):
        res = br.remove_background_batch(np_imgs)
        _ = float(res[0].all_ious.sum())               # host read of the step's result
    torch.cuda.synchronize(dev)
    t_local = time.perf_counter() - t0
    return sharder.max_over_ranks(t_local, device=dev), res


def verify_batch_against_single(model, d_imgs, out_batch, picks):
    """After the timed loop: images `picks` of the benchmarked batch are run ALONE (batch 1) on the same context and must give
    the same bits - mask logits, all_masks, RGBA (the IoU logits' pooled mean is a two-stage sum: equal to 2e-5)."""
    out, outs, ious, best = out_batch
    keep = {i: (out["pred_masks"][i].clone(), out["pred_iou"][i].clone(), outs[i][0].clone(), outs[i][1].clone(), int(best[i]))
            for i in picks}
    ok = True
    for i in picks:
        o1, outs1, _, best1 = model.run_u8([d_imgs[i]], slot=1)
        lm, li, am, rg, bi = keep[i]
        ok = ok and bool(torch.equal(o1["pred_masks"][0], lm)) and bool(torch.equal(outs1[0][0], am)) and bool(torch.equal(outs1[0][1], rg))
        ok = ok and float((o1["pred_iou"][0] - li).abs().max()) <= 2e-5 and int(best1[0]) == bi
    return ok


def latency_batch1(br, model, np_img, d_img, dev, iters=20):
    """The reference API is one image per call: device-timed and end-to-end latency of a single image."""
    for _ in range(3):
        model.run_u8([d_img], slot=1)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        model.run_u8([d_img], slot=1)
    e1.record()
    torch.cuda.synchronize(dev)
    dev_ms = e0.elapsed_time(e1) / iters
    for _ in range(2):
        br.remove_background(np_img)
    t0 = time.perf_counter()
    for _ in range(iters):
        r = br.remove_background(np_img)
        _ = float(r.all_ious.sum())
    return dev_ms, (time.perf_counter() - t0) / iters * 1e3


def gpu_baseline(S, batch, dev):
    """SURVEY 8(d) last row: the reference's own GPU path on the SAME B200 - the oracle restatement of DPTSegmentation.forward
    run by PyTorch (cuBLAS / cuDNN / SDPA flash, channels_last, cudnn.benchmark) in fp32 (TF32 off), under
    torch.autocast(bfloat16) and with bf16 weights; model forward only (no pre / post-process), CUDA-event timed.
    This - not the CPU number - is the bar the hand-written kernels have to beat."""
    from oracle import model as om
    sd = synth_state_dict(VITB, 0)
    cl = torch.channels_last
    sd_dev = {k: (v.to(dev).contiguous(memory_format=cl) if v.dim() == 4 else v.to(dev)) for k, v in sd.items()}
    x = torch.randn(batch, 3, S, S, device=dev)
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.benchmark = True
    res = {"batch": batch, "sdpa": True, "channels_last": True, "cudnn_benchmark": True, "unit": UNIT,
           "what": "oracle restatement of DPTSegmentation.forward (model.py:99-106) run by torch on this GPU, model forward only"}

    def run(tag, tf32, autocast, weights):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        xx = x.to(weights.get("_dtype", torch.float32))
        try:
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                for _ in range(2):
                    om.forward(weights, xx, VITB, sdpa=True) if weights is sd_dev else _forward_lowp(om, weights, xx)
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                n = 3
                for _ in range(n):
                    om.forward(weights, xx, VITB, sdpa=True) if weights is sd_dev else _forward_lowp(om, weights, xx)
                e1.record()
                torch.cuda.synchronize(dev)
            res[tag] = round(n * batch / (e0.elapsed_time(e1) / 1e3), 2)
        except Exception as e:  # noqa: BLE001 - a baseline that cannot run is reported, not fatal
            res[tag] = None
            res[tag + "_error"] = repr(e)[:200]

    try:
        run("fp32", False, False, sd_dev)
        run("tf32", True, False, sd_dev)
        run("bf16_autocast", False, True, sd_dev)
        sd_bf = {k: (v.to(torch.bfloat16) if v.is_floating_point() else v) for k, v in sd_dev.items()}
        sd_bf["_dtype"] = torch.bfloat16
        run("bf16_weights", False, True, sd_bf)          # + autocast: the fp32 RoPE tables / LayerNorm outputs meet bf16 operands
        # the training step of config 4 the same way: oracle arithmetic (train-mode head) + loss restatement under autocast(bf16),
        # torch.autograd, torch.optim.AdamW with the reference's two groups; batch TRAIN_BATCH at 1024^2
        try:
            from oracle import loss as ol
            sd_t = {k: v.detach().clone() for k, v in sd_dev.items()}
            names = [k for k, v in sd_t.items() if v.is_floating_point() and "running" not in k]
            for k in names:
                sd_t[k].requires_grad_(True)
            opt = torch.optim.AdamW([{"params": [sd_t[k] for k in names if k.startswith("encoder.")], "lr": 1e-5},
                                     {"params": [sd_t[k] for k in names if k.startswith("seg_head.")], "lr": 1e-4}], weight_decay=0.05)
            xt = torch.randn(TRAIN_BATCH, 3, S, S, device=dev)
            mt = (torch.rand(TRAIN_BATCH, S, S, device=dev) > 0.5).float()

            def train_step():
                opt.zero_grad(set_to_none=True)
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    taps = om.encoder_taps(sd_t, xt, VITB, None, True)
                    out = om.head_forward(sd_t, taps, S // 16, S // 16, VITB, train=True)
                loss, _ = ol.loss_module(out["pred_masks"].float(), out["pred_iou"].float(), mt, 0)
                loss.backward()
                opt.step()
            for _ in range(2):
                train_step()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                train_step()
            e1.record()
            torch.cuda.synchronize(dev)
            res["train_step_bf16_autocast"] = round(3 * TRAIN_BATCH / (e0.elapsed_time(e1) / 1e3), 2)
            res["train_step_batch"] = TRAIN_BATCH
        except Exception as e:  # noqa: BLE001
            res["train_step_bf16_autocast"] = None
            res["train_step_error"] = repr(e)[:200]
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark = old
    return res


def _forward_lowp(om, sd_bf, x):
    """The oracle arithmetic with bf16 weights and activations (`model.to(torch.bfloat16)` in reference terms)."""
    sd = {k: v for k, v in sd_bf.items() if k != "_dtype"}
    x = x.to(torch.bfloat16)
    taps = om.encoder_taps(sd, x, VITB, None, True)
    return om.head_forward(sd, taps, x.shape[-2] // 16, x.shape[-1] // 16, VITB)


def extra_leg(model_name, S, src, B, micro_batch, local_rank, rank, world, dev, steps, e2e_steps, br=None):
    """Short device-timed + end-to-end measurement of another BASELINE.json configuration (reported under `extra`, the headline
    stays configs[1]): cfg3 = 2048^2 sources with full-resolution mask up-sampling + RGBA composite, cfg5 = ViT-L / one mask."""
    from s3od_b200 import BackgroundRemoval
    arch = VITL if model_name == "dinol" else VITB
    own = br is None
    if own:
        ckpt = os.path.join("/tmp", f"s3od_synth_{model_name}_seed0_{os.getpid()}.pt")
        save_checkpoint(ckpt, arch, 0)
        br = BackgroundRemoval(model_id=ckpt, image_size=S, device=f"cuda:{local_rank}", max_batch=B, micro_batch=micro_batch,
                               encoder_name="dinov3_large" if model_name == "dinol" else "dinov3_base", num_outputs=arch.num_outputs)
        os.remove(ckpt)
    b0, _ = sharder.shard_range(B * world, rank, world)
    np_imgs, d_imgs = make_images(B, src, 10_000 + b0, dev)
    ms, launches, _, _, out = time_device(br.model, d_imgs, steps, 3, dev)
    del out
    t_e2e, res = time_e2e(br, np_imgs, e2e_steps, 2, dev)      # 2 warm-ups: both alternating sets of pinned result buffers exist
    del res, d_imgs
    K = arch.num_outputs
    ent = {"workload": f"{model_name} inference bf16, batch {B} synthetic {src}x{src} uint8 images per GPU, image_size {S}",
           "value": round(world * B * steps / (ms / 1e3), 2), "unit": UNIT, "steps": steps, "ms_per_step": round(ms / steps, 3),
           "e2e": {"value": round(world * B * e2e_steps / t_e2e, 2), "unit": UNIT, "steps": e2e_steps,
                   "h2d_bytes_per_step": B * src * src * 3, "d2h_bytes_per_step": B * (K * src * src * 4 + src * src * 4 + K * 4 + 4)},
           "gpu_launches": int(launches), "n_gpus": world}
    if S == 1024:
        gf = GFLOP_PER_IMAGE_VITL if model_name == "dinol" else GFLOP_PER_IMAGE
        ent["whole_step_tflops_per_gpu"] = round(gf * ent["value"] / world / 1e3, 1)
    if own:
        br.close()
    return ent


def training_slice_leg(rank, world, dev):
    """BASELINE.json configs[3] - the parts of the training step that exist (DESIGN.md section 6 says what does not): the loss
    forward + backward kernels at config/dataset/synth.yaml's shape (batch 4 per GPU, 1024^2, 3 masks), and the data-parallel
    exchange step over the 107.8 M gradients (431 MB fp32): (A) bucketed NCCL all-reduce + fused AdamW, (B) ONE fused kernel per
    rank over CUDA-IPC peer memory (reduce its slice over NVLink, AdamW, broadcast fp32 + bf16 parameters).  At N > 1 this is the
    one place the repository moves data between GPUs."""
    from s3od_b200.training import FusedAdamW, FusedDataParallelAdamW, GradientAllReduce, LossModule, ParameterLayout
    ent = {"what": "training step (config 4): loss fwd+bwd, gradient exchange + AdamW, encoder block, whole step", "n_gpus": world}
    try:
        lm = LossModule()
        B, S = 4, 1024
        g = torch.Generator(device=dev).manual_seed(7 + rank)
        z = torch.randn(B, 3, S, S, device=dev, generator=g) * 3
        q = torch.randn(B, 3, device=dev, generator=g)
        t = (torch.rand(B, S, S, device=dev, generator=g) > 0.5).float()
        reps = 8                                            # back-to-back calls: the GPU queue hides the host-side launch cost
        for _ in range(3):
            lm.forward_backward({"pred_masks": z, "pred_iou": q}, {"masks": t}, 0)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            loss, _, _, _ = lm.forward_backward({"pred_masks": z, "pred_iou": q}, {"masks": t}, 0)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / reps
        nbytes = B * S * S * 4 * (4 + 4 + 3)               # pass 1 reads z (3 planes) + t; pass 3 reads them again and writes dz (3 planes)
        ent["loss_fwd_bwd"] = {"batch": B, "image_size": S, "ms": round(ms, 4), "hbm_gbs": round(nbytes / ms / 1e6, 1),
                               "hbm_frac": round(nbytes / ms / 1e6 / load_peaks()["hbm_gbs"], 3), "loss": round(float(loss), 5)}
        lay = ParameterLayout(VITB)
        gen = torch.Generator(device=dev).manual_seed(1)
        p0 = torch.randn(lay.total, device=dev, generator=gen) * 0.05
        grads = torch.randn(lay.total, device=dev, generator=torch.Generator(device=dev).manual_seed(100 + rank))
        pa, ga = p0.clone(), torch.empty_like(p0)
        opt_a = FusedAdamW(lay, pa, lr=1e-5, bf16_copy=True)
        red = GradientAllReduce(lay, ga)

        def timed(fn, prep, n=5):
            best = None
            for _ in range(n):
                prep()
                sharder.barrier()
                torch.cuda.synchronize(dev)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize(dev)
                tm = sharder.max_over_ranks(a.elapsed_time(b), device=dev)
                best = tm if best is None else min(best, tm)
            return best

        def step_a():
            for sgm in lay.segments:
                red.mark_ready(sgm.name)
            red.finish()
            opt_a.step(ga, grad_scale=1.0 / world)

        def prep_a():
            ga.copy_(grads)
            red.reset()
        ta = timed(step_a, prep_a)
        opt_b = FusedDataParallelAdamW(lay, dev, lr=1e-5)
        opt_b.param().copy_(p0)
        tb = timed(opt_b.step, lambda: opt_b.grad().copy_(grads))
        ent["exchange_and_adamw"] = {
            "params_with_grad": lay.numel_with_grad(), "grad_bytes": lay.total * 4, "buckets": lay.num_buckets,
            "nccl_allreduce_plus_adamw_ms": round(ta, 3), "fused_p2p_kernel_ms": round(tb, 3), "speedup": round(ta / tb, 2),
            # each link direction of a GPU carries its pushes (6 B / parameter of its slice per peer) plus the gradient slices its
            # peers read from it (4 B): the roofline of the fused step is that volume over the measured 770 GB/s peer bandwidth
            "link_bytes_per_direction": opt_b.link_bytes_per_step(),
            "link_gbs": round(opt_b.link_bytes_per_step() / tb / 1e6, 1) if world > 1 else None,
            "frac_of_link_roofline_770": round(opt_b.link_bytes_per_step() / 770e6 / tb, 3) if world > 1 else None,
            "adamw_hbm_gbs_n1": round(lay.total * 30 / ta / 1e6, 1) if world == 1 else None}
        opt_b.close()
        if world > 1:
            opt_c = FusedDataParallelAdamW(lay, dev, lr=1e-5, replicate_fp32=False)       # fp32 masters sharded, bf16 copy pushed
            opt_c.param().copy_(p0)
            tc = timed(opt_c.step, lambda: opt_c.grad().copy_(grads))
            ent["exchange_and_adamw"]["fused_p2p_sharded_masters_ms"] = round(tc, 3)
            ent["exchange_and_adamw"]["sharded_frac_of_link_roofline_770"] = round(opt_c.link_bytes_per_step() / 770e6 / tc, 3)
            opt_c.close()
        # one encoder block forward + backward at the configuration's resolution (first, unfused form: see DESIGN.md section 6)
        from s3od_b200.training import EncoderBlockStep
        blk = EncoderBlockStep(synth_state_dict(VITB, 0), "encoder.model.layer.3.", VITB, 1024, dev)
        Bb = 2
        xb = torch.randn(Bb, 4101, 768, device=dev, generator=gen)
        gb = torch.randn(Bb, 4101, 768, device=dev, generator=gen)
        blk.forward(xb)
        blk.backward(gb)
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        blk.forward(xb)
        blk.backward(gb)
        b.record()
        torch.cuda.synchronize(dev)
        tms = a.elapsed_time(b)
        ent["encoder_block_fwd_bwd"] = {"batch": Bb, "image_size": 1024, "ms": round(tms, 2),
                                        "tflops": round(3 * 109.7 * Bb / tms, 1),
                                        "note": "109.7 GFLOP forward per image and block (SURVEY 8d), backward counted as 2x"}
        del blk, xb, gb
        torch.cuda.empty_cache()
        # the whole optimisation step of config 4 (train-mode forward, loss, backward, gradient exchange over the process group,
        # fused AdamW) at config/dataset/synth.yaml's batch of four 1024^2 images per GPU: a REAL step with a
        # REAL collective, so the N = 1, 2, 4, 8 records show how the exchange scales
        from s3od_b200.training_head import TrainStep
        ts = TrainStep(synth_state_dict(VITB, 0), VITB, 1024, dev, lr=1e-5)
        g2 = torch.Generator(device=dev).manual_seed(11 + rank)
        xi = torch.randn(TRAIN_BATCH, 3, 1024, 1024, device=dev, generator=g2)
        mi = (torch.rand(TRAIN_BATCH, 1024, 1024, device=dev, generator=g2) > 0.5).float()
        ts.step(xi, mi)
        torch.cuda.synchronize(dev)
        sharder.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        nst = 2
        for _ in range(nst):
            lossv, _ = ts.step(xi, mi)
        b.record()
        torch.cuda.synchronize(dev)
        tstep = sharder.max_over_ranks(a.elapsed_time(b), device=dev) / nst
        ent["train_step"] = {"batch_per_gpu": TRAIN_BATCH, "image_size": 1024, "ms_per_step": round(tstep, 1),
                             "images_per_s": round(world * TRAIN_BATCH * 1e3 / tstep, 2),
                             "loss": round(float(lossv), 4), "allreduce_buckets": ts.layout.num_buckets,
                             "what": "TrainStep.step: encoder + head forward (train mode), loss fwd+bwd, full backward, bucketed all-reduce, fused AdamW"}
        del ts
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001 - an extra leg must not take the headline number down with it
        ent["error"] = repr(e)[:300]
    return ent


def run_b200(args, rank, local_rank, world):
    from s3od_b200 import BackgroundRemoval
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    B, S, src = args.batch, args.image_size, args.source
    arch = VITL if args.model == "dinol" else VITB
    ckpt = os.path.join("/tmp", f"s3od_synth_{args.model}_seed0_{os.getpid()}.pt")
    save_checkpoint(ckpt, arch, 0)
    br = BackgroundRemoval(model_id=ckpt, image_size=S, device=f"cuda:{local_rank}", max_batch=B, micro_batch=args.micro_batch,
                           encoder_name="dinov3_large" if args.model == "dinol" else "dinov3_base", num_outputs=arch.num_outputs)
    os.remove(ckpt)
    model = br.model
    # distinct images per rank / slot; device copies resident in HBM for `value`, pageable host arrays for `e2e`
    b0, _ = sharder.shard_range(B * world, rank, world)
    np_imgs, d_imgs = make_images(B, src, b0, dev)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_total, launches, prof, clocks, out = time_device(model, d_imgs, args.steps, args.warmup, dev, sampler, profile=True)
    # ---- the benchmarked batch is checked, not only timed: two of its images alone give the same bits
    verified = {"batch_vs_single_bitwise": verify_batch_against_single(model, d_imgs, out, sorted({0, B - 1}))}
    del out

    # ---- end to end through the public API (H2D + D2H inside the timed region)
    e2e_steps = max(1, args.e2e_steps)
    t_e2e, res = time_e2e(br, np_imgs, e2e_steps, args.warmup, dev)
    h2d = B * src * src * 3
    K = arch.num_outputs
    d2h = B * (K * src * src * 4 + src * src * 4 + K * 4 + 4)
    api_first = [(res[i].all_masks.copy(), int(res[i].all_ious.argmax()), res[i].all_ious.copy()) for i in range(min(2, B))]
    del res
    lat_dev_ms, lat_e2e_ms = latency_batch1(br, model, np_imgs[0], d_imgs[0], dev)

    extra = {}
    if args.extras and args.model == "dinob" and S == 1024 and src == 1024:
        # BASELINE.json configs[2] shape on the same context, configs[4] on its own ViT-L context (after this one is freed)
        extra["cfg3_source2048"] = extra_leg("dinob", S, 2048, B, args.micro_batch, local_rank, rank, world, dev, 5, 3, br=br)
    del d_imgs
    br.close()
    torch.cuda.empty_cache()
    if args.extras and args.model == "dinob" and S == 1024 and src == 1024:
        extra["cfg5_dinol"] = extra_leg("dinol", S, 1024, B, args.micro_batch, local_rank, rank, world, dev, 3, 2)
        extra["cfg4_training_slice"] = training_slice_leg(rank, world, dev)

    if rank != 0:
        return None
    ms_per_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total / 1e3)
    # ---- per-family device time (CUDA events on the launch stream, recorded inside the timed region)
    if args.dump_profile:
        with open(args.dump_profile, "w") as f:
            f.write("label\tlaunches\timages\ttotal_ms\tms_per_launch\n")
            for label, n, imgs, ms in prof:
                f.write(f"{label}\t{n}\t{imgs}\t{ms:.4f}\t{ms / max(n, 1):.4f}\n")
    peaks = load_peaks()
    fam_ms, fam_n = {}, {}
    for label, n, imgs, ms in prof:
        f = kernel_family(label)
        fam_ms[f] = fam_ms.get(f, 0.0) + ms
        fam_n[f] = fam_n.get(f, 0) + n
    total_ms = sum(fam_ms.values())
    images = B * args.steps
    families = {}
    K = arch.num_outputs
    # algorithmic bytes per image of the two bandwidth kernels either side of the network
    family_bytes = {"preprocess": src * src * 3 + 3 * S * S * 2,
                    "postprocess": K * S * S * 4 + src * src * 3 + K * src * src * 4 + src * src * 4}
    if args.model == "dinob":
        ntok, g = (S // 16) ** 2 + 5, S // 16
        # residual add + LayerNorm: x fp32 read + write, dx bf16 read, y bf16 write = 12 B per element, 2 per layer, 11 layers
        # (the first has no dx; the four tap copies add 2 B per patch element)
        family_bytes["layernorm"] = ntok * 768 * (22 * 12 - 2) + 4 * (ntok - 5) * 768 * 2
        # bilinear x2 of the four fusion levels, 256 channels bf16: read 1 + write 4 per input pixel (+ the tiny IoU head / prefix kernels)
        family_bytes["head_bandwidth"] = sum((g * 2 ** l // 2) ** 2 for l in range(4)) * 256 * 2 * 5
    for f, ms in sorted(fam_ms.items(), key=lambda kv: -kv[1]):
        if ms <= 0.0:
            continue
        ent = {"share": round(ms / total_ms, 4), "ms_per_image": round(ms / images, 5), "launches": fam_n[f]}
        if f in FAMILY_GFLOP and S == 1024 and args.model == "dinob":
            ent["tflops"] = round(FAMILY_GFLOP[f] * images / ms, 1)
        if f in family_bytes:
            gbs = family_bytes[f] * images / ms / 1e6
            ent["hbm_gbs"] = round(gbs, 1)
            ent["hbm_frac"] = round(gbs / peaks["hbm_gbs"], 4)
        families[f] = ent
    dom = max((f for f in fam_ms if f in FAMILY_GFLOP), key=lambda f: fam_ms[f])
    flops_known = S == 1024 and args.model == "dinob"
    gflop_img = GFLOP_PER_IMAGE if args.model == "dinob" else GFLOP_PER_IMAGE_VITL
    achieved = FAMILY_GFLOP[dom] * images / fam_ms[dom] if flops_known else None
    # dram bytes per launch of the dominant kernel from the committed ncu --set full capture (same command, same micro-batch)
    traffic, traffic_src = newest_traffic(dom, model.micro_batch, args.model, S)
    roofline = {"kernel": dom, "bound": "tensor", "achieved": round(achieved, 1) if achieved else None,
                "peak": peaks["tflops_sustained"], "peak_source": peaks["source"] + " (sustained: timed inside a long step)",
                "unit": "TFLOP/s", "frac": round(achieved / peaks["tflops_sustained"], 4) if achieved else None,
                "traffic": traffic, "traffic_source": traffic_src, "avg_launch_ms": round(fam_ms[dom] / fam_n[dom], 4),
                "whole_step_tflops": round(gflop_img * value / world / 1e3, 1) if S == 1024 else None,
                "whole_step_frac": round(gflop_img * value / world / 1e3 / peaks["tflops_sustained"], 4) if S == 1024 else None}
    if dom == "attention" and achieved and clocks and clocks.get("sm_mhz"):
        # the attention kernel is bound by the SFU (one MUFU.EX2 per score = per 256 tensor FLOP at head_dim 64; 16 per clock
        # and SM at every precision, tools/lab/pipe_rate.cu) - reported next to the tensor fraction the contract asks for
        exps_per_s = achieved * 1e12 / 256.0
        sfu_peak = 16.0 * torch.cuda.get_device_properties(dev).multi_processor_count * clocks["sm_mhz"] * 1e6
        roofline["sfu"] = {"achieved_gexp_s": round(exps_per_s / 1e9, 1), "peak_gexp_s": round(sfu_peak / 1e9, 1),
                           "frac": round(exps_per_s / sfu_peak, 4), "note": "16 MUFU.EX2 / clk / SM at the sampled SM clock"}
    line = {
        "metric": METRIC.replace("dinob", args.model), "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": config_dict(args),
        "clocks": clocks,
        "e2e": {"value": round(world * B * e2e_steps / t_e2e, 3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "api": "BackgroundRemoval.remove_background_batch (pageable host uint8 arrays in, pinned on "
                                           "ingest through a staging ring; host results out)"},
        "gpu_launches": int(launches),
        "latency_b1_ms": round(lat_dev_ms, 3),
        "latency_b1": {"device_ms": round(lat_dev_ms, 3), "e2e_ms": round(lat_e2e_ms, 3), "images_per_s_device": round(1e3 / lat_dev_ms, 1),
                       "api": "BackgroundRemoval.remove_background (one image per call, the reference API's shape)"},
        "verified": verified,
        "roofline": roofline,
        "kernels": families,
    }
    if extra:
        line["extra"] = extra
    line["_api_first"] = api_first
    return line


def cpu_baseline(sample_images: int, S: int, src: int):
    """The CPU oracle (port of the reference path) on the host cores: full remove_background per image."""
    from oracle.pipeline import remove_background
    sd = synth_state_dict(VITB, 0)
    torch.set_num_threads(os.cpu_count() or 1)
    remove_background(sd, synth_noise_image(64, 64, seed=0), VITB, 64)          # thread-pool / allocator warm-up, tiny
    t0 = time.perf_counter()
    outs = []
    for i in range(sample_images):
        outs.append(remove_background(sd, synth_noise_image(src, src, seed=i), VITB, S))
    dt = time.perf_counter() - t0
    return {"value": round(sample_images / dt, 4), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{sample_images} image(s) {src}x{src}, image_size {S}, full remove_background on the CPU oracle "
                      f"(torch {torch.__version__} fp32), {dt:.1f} s"}, outs


def check_against_oracle(api_first, oracle_outs):
    """The images the CPU baseline just processed are images 0.. of rank 0's benchmarked batch: compare the public-API results
    of the timed e2e loop with the oracle's (tolerances of tests/test_gpu_parity.py)."""
    n = min(len(api_first), len(oracle_outs))
    if n == 0:
        return None
    mx = mean = 0.0
    idx_ok = True
    for (am, best, ious), ref in zip(api_first[:n], oracle_outs[:n]):
        d = np.abs(am - ref["all_masks"])
        mx, mean = max(mx, float(d.max())), max(mean, float(d.mean()))
        idx_ok = idx_ok and best == int(ref["best_idx"]) and float(np.abs(ious - ref["all_ious"]).max()) <= 1e-2
    return {"images": n, "all_masks_max_abs": round(mx, 5), "all_masks_mean_abs": round(mean, 6), "best_idx_and_ious_match": idx_ok,
            "ok": bool(mx <= 5e-2 and mean <= 6e-3 and idx_ok)}


def run_reference(args, rank, world):
    """--impl reference: the reference path's own CPU implementation on the host cores (the oracle port: the reference is
    Python under /root/reference and cannot travel to the GPU box), all host threads, same config / metric / unit as the
    b200 arm; one step = a 1-image sample of the batch (about 3 s of CPU work)."""
    if rank != 0:
        return None
    from oracle.pipeline import remove_background
    sd = synth_state_dict(VITB, 0)
    torch.set_num_threads(os.cpu_count() or 1)
    S, src = args.image_size, args.source
    imgs = [synth_noise_image(src, src, seed=i) for i in range(2)]
    for i in range(args.warmup):
        remove_background(sd, imgs[i % 2], VITB, S)
    t0 = time.perf_counter()
    for i in range(args.steps):
        remove_background(sd, imgs[i % 2], VITB, S)       # one step = a 1-image sample of the batch
    dt = time.perf_counter() - t0
    v = round(args.steps / dt, 4)
    sample = f"1 image {src}x{src} per step (bounded sample of the batch-{args.batch} workload), oracle port of the reference CPU path"
    return {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 2), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(args),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU per step (BASELINE.json configs[1])")
    ap.add_argument("--image-size", type=int, default=1024)
    ap.add_argument("--source", type=int, default=1024, help="source image side (2048 = configs[2] shape)")
    ap.add_argument("--micro-batch", type=int, default=32)
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-extras", dest="extras", action="store_false", help="skip the short cfg3 (2048^2 sources) / cfg5 (ViT-L) legs")
    ap.add_argument("--no-gpu-baseline", dest="gpu_baseline", action="store_false", help="skip the PyTorch-on-this-GPU baseline leg")
    ap.add_argument("--gpu-baseline-batch", type=int, default=8)
    ap.add_argument("--config", default="infer", choices=["infer", "train"], help="train = the optimisation step of BASELINE.json configs[3] as the headline")
    ap.add_argument("--no-numa-bind", dest="numa_bind", action="store_false", help="do not pin the rank to the GPU's local CPU cores")
    ap.add_argument("--model", default="dinob", choices=["dinob", "dinol"], help="dinol = ViT-L backbone, one mask (BASELINE.json configs[4])")
    ap.add_argument("--dump-profile", default=None, help="write the per-kernel CUDA-event table (label, launches, images, ms) here")
    ap.add_argument("--cpu-sample", type=int, default=2, help="images timed on the CPU oracle (0 = skip)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        line = run_reference(args, rank, world)
        if line is not None:
            print(json.dumps(line), flush=True)
        return
    if args.warmup < 3:
        args.warmup = 3
    rank, local_rank, world = sharder.init_from_env("nccl")
    if args.config == "train":
        # BASELINE.json configs[3] as the headline of this invocation: the whole optimisation step, weak scaling
        dev = torch.device("cuda", local_rank)
        torch.cuda.set_device(dev)
        ent = training_slice_leg(rank, world, dev)
        if rank == 0:
            tsr = ent.get("train_step") or {}
            print(json.dumps({"metric": "images/sec (dinob training step: forward + loss + backward + gradient all-reduce + AdamW)",
                              "value": tsr.get("images_per_s"), "unit": UNIT, "n_gpus": world, "steps": 2, "warmup": 1,
                              "ms_per_step": tsr.get("ms_per_step"), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                              "dtype": "bf16 operands, fp32 accumulation / master weights", "data": "synthetic",
                              "config": {"workload": f"dinob training step, {TRAIN_BATCH} synthetic 1024x1024 images per GPU, image_size 1024, seeded random weights",
                                         "batch_per_gpu": TRAIN_BATCH, "image_size": 1024}, "detail": ent}), flush=True)
        if world > 1:
            torch.distributed.barrier()
            torch.distributed.destroy_process_group()
        return
    # N > 1: host staging / pinned result buffers stay on the GPU's socket (at N = 1 the CPU-baseline leg wants every core)
    numa_bound = sharder.bind_to_gpu_numa(local_rank) if (args.numa_bind and world > 1) else False
    line = run_b200(args, rank, local_rank, world)
    if rank == 0:
        line["numa_bound"] = bool(numa_bound)           # (not part of `config`: both arms quote the same config dict)
    if rank == 0:
        api_first = line.pop("_api_first")
        if world == 1 and args.gpu_baseline and args.model == "dinob":
            line["gpu_baseline"] = gpu_baseline(args.image_size, args.gpu_baseline_batch, torch.device("cuda", local_rank))
            ours = line["value"]
            best = max((line["gpu_baseline"].get(k) or 0.0) for k in ("fp32", "tf32", "bf16_autocast", "bf16_weights"))
            line["gpu_baseline"]["speedup_over_best_torch"] = round(ours / best, 2) if best > 0 else None
            line["gpu_baseline"]["note"] = ("PRIMARY baseline: same GPU, same arithmetic, PyTorch library kernels; the CPU figure "
                                            "below is the reference's CPU path and is secondary")
        if world == 1 and args.cpu_sample > 0:
            line["cpu_baseline"], oracle_outs = cpu_baseline(args.cpu_sample, args.image_size, args.source)
            line["cpu_baseline"]["note"] = "secondary (reported, not the bar): CPU path of the reference on the host cores"
            line["verified"]["e2e_vs_cpu_oracle"] = check_against_oracle(api_first, oracle_outs)
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
