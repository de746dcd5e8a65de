"""Aggregate device-to-host bandwidth of the box (torchrun, one rank per GPU): every rank copies a 2 GiB device buffer into pinned
host memory 6 times; prints per-rank and total GB/s.  Tells whether the end-to-end 2048^2 number at 8 GPUs is platform-bound."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from s3od_b200 import sharder  # noqa: E402

rank, local_rank, world = sharder.init_from_env("nccl")
dev = torch.device("cuda", local_rank)
torch.cuda.set_device(dev)
n = 2 << 30
d = torch.empty(n, dtype=torch.uint8, device=dev)
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h.copy_(d, non_blocking=True)
torch.cuda.synchronize(dev)
sharder.barrier()
t0 = time.perf_counter()
for _ in range(6):
    h.copy_(d, non_blocking=True)
torch.cuda.synchronize(dev)
dt = time.perf_counter() - t0
sharder.barrier()
gbs = 6 * n / dt / 1e9
tot = sharder.sum_over_ranks(gbs, device=dev)
tmax = sharder.max_over_ranks(dt, device=dev)
# pinned allocation cost (what a call that needs NEW result buffers pays)
t1 = time.perf_counter()
h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
t_alloc = time.perf_counter() - t1
if rank == 0:
    print(f"world {world}: D2H per rank {gbs:.1f} GB/s, sum over ranks {tot:.1f} GB/s (slowest rank {6 * n / tmax / 1e9:.1f} GB/s); "
          f"pinning 2 GiB of fresh host memory: {t_alloc * 1e3:.0f} ms", flush=True)
if world > 1:
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
