"""Where the time of the fused exchange + AdamW step goes (torchrun, one rank per GPU): the two barriers alone, the kernel alone
(ranks synchronised before, no barriers inside the timed region), and the whole step."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from s3od_b200 import sharder                                                             # noqa: E402
from s3od_b200.arch import VITB                                                           # noqa: E402
from s3od_b200.training import FusedDataParallelAdamW, ParameterLayout                    # noqa: E402


def timed(fn, dev, n=8):
    best = 1e9
    for _ in range(n):
        sharder.barrier()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize(dev)
        best = min(best, sharder.max_over_ranks(a.elapsed_time(b), device=dev))
    return best


def main():
    rank, local_rank, world = sharder.init_from_env("nccl")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lay = ParameterLayout(VITB)
    for rep in (True, False):
        opt = FusedDataParallelAdamW(lay, dev, lr=1e-5, replicate_fp32=rep)
        opt.param().normal_(0, 0.05)
        opt.grad().normal_()
        t_bar = timed(lambda: (opt._barrier(), opt._barrier()), dev)
        t_step = timed(opt.step, dev)
        saved = opt._barrier
        opt._barrier = lambda: None
        t_kernel = timed(opt.step, dev)
        opt._barrier = saved
        if rank == 0:
            link = opt.link_bytes_per_step()             # bytes per link direction: pushes + the slices the peers read
            print(json.dumps({"world": world, "replicate_fp32": rep, "two_barriers_ms": round(t_bar, 3), "kernel_only_ms": round(t_kernel, 3),
                              "step_ms": round(t_step, 3), "link_bytes_per_direction": link, "kernel_link_gbs": round(link / t_kernel / 1e6, 1),
                              "link_roofline_ms_at_770": round(link / 770e6, 3), "frac_of_link_roofline": round(link / 770e6 / t_kernel, 3)}), flush=True)
        opt.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
