"""Multi-GPU check + timing of the data-parallel exchange step (run with torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/ddp_fused_check.py

  A  baseline: bucketed NCCL all-reduce of the 431 MB flat gradient buffer (GradientAllReduce) + FusedAdamW (two kernels);
  B  fused:    s3od_ddp_fused_adamw_step over CUDA-IPC peer-mapped buffers (one kernel per rank between two tiny barriers).
Both start from the same parameters and the same per-rank gradients; the resulting parameters must agree to one fp32 ulp of
the update (the all-reduce sums in a ring order, the fused kernel in rank order), and every replica of B must be bit-identical.
Times are CUDA events on the launch stream, max over ranks."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from s3od_b200 import sharder                                                             # noqa: E402
from s3od_b200.arch import VITB                                                           # noqa: E402
from s3od_b200.training import FusedAdamW, FusedDataParallelAdamW, GradientAllReduce, ParameterLayout   # noqa: E402


def main():
    rank, local_rank, world = sharder.init_from_env("nccl")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lay = ParameterLayout(VITB)
    gen = torch.Generator(device=dev).manual_seed(1)
    p0 = torch.randn(lay.total, device=dev, generator=gen) * 0.05            # same seed on every rank: identical replicas
    gen_r = torch.Generator(device=dev).manual_seed(100 + rank)
    grads = [torch.randn(lay.total, device=dev, generator=gen_r) for _ in range(2)]

    # ---- A: NCCL all-reduce per bucket + fused AdamW
    pa = p0.clone()
    ga = torch.empty_like(pa)
    opt_a = FusedAdamW(lay, pa, lr=1e-5, bf16_copy=True)
    red = GradientAllReduce(lay, ga)
    times_a = []
    for it in range(6):
        ga.copy_(grads[it % 2])
        red.reset()
        sharder.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in lay.segments:
            red.mark_ready(s.name)
        red.finish()
        opt_a.step(ga, grad_scale=1.0 / world)
        e1.record()
        torch.cuda.synchronize(dev)
        times_a.append(sharder.max_over_ranks(e0.elapsed_time(e1), device=dev))

    # ---- B: fused exchange + AdamW over peer memory
    opt_b = FusedDataParallelAdamW(lay, dev, lr=1e-5)
    opt_b.param().copy_(p0)
    times_b = []
    for it in range(6):
        opt_b.grad().copy_(grads[it % 2])
        sharder.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        opt_b.step()
        e1.record()
        torch.cuda.synchronize(dev)
        times_b.append(sharder.max_over_ranks(e0.elapsed_time(e1), device=dev))
    diff = float((opt_b.param() - pa).abs().max())
    bf_ok = bool(torch.equal(opt_b.param_bf16().float(), opt_b.param().bfloat16().float()))
    # every replica of B holds the same bits: compare checksums across ranks
    chk = opt_b.param().double().sum().reshape(1)
    allc = [torch.zeros_like(chk) for _ in range(world)]
    if world > 1:
        dist.all_gather(allc, chk)
    same = all(float(c) == float(allc[0]) for c in allc) if world > 1 else True
    if rank == 0:
        ta, tb = min(times_a[2:]), min(times_b[2:])
        nbytes = lay.total * 4
        print(json.dumps({"world": world, "params_with_grad": lay.numel_with_grad(), "grad_bytes": nbytes, "buckets": lay.num_buckets,
                          "nccl_allreduce_plus_adamw_ms": round(ta, 3), "fused_p2p_step_ms": round(tb, 3), "speedup": round(ta / tb, 2),
                          "fused_link_bytes_per_gpu": opt_b.link_bytes_per_step(),
                          "fused_link_gbs_per_direction": round(opt_b.link_bytes_per_step() / tb / 1e6, 1) if world > 1 else None,
                          "frac_of_link_roofline_770": round(opt_b.link_bytes_per_step() / 770e6 / tb, 3) if world > 1 else None,
                          "max_abs_diff_vs_allreduce_path": diff, "bf16_copy_consistent": bf_ok, "replicas_bit_identical": same,
                          "ok": bool(diff <= 1e-6 and bf_ok and same)}), flush=True)
    opt_b.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
