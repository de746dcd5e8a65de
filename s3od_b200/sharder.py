"""Batch sharding across the GPUs of one box: images are independent, so the inference path has no collective.

One process per GPU (torchrun); every rank takes a contiguous slice of the batch, runs the full pipeline on its
own replica of the weights, and only timing is reduced (max over ranks).  SURVEY 8(e).
"""
import os
from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of n_items for `rank`; the first n_items % world ranks get one extra item."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world: {rank}/{world}")
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def init_from_env(backend: str = "nccl"):
    """(rank, local_rank, world) from the torchrun environment; initialises the process group when world > 1."""
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device="cpu") -> float:
    """Max of a scalar over all ranks (timing is reported as the slowest rank)."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device="cpu") -> float:
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def bind_to_gpu_numa(device_index: int) -> bool:
    """Pin the calling thread to the CPU cores NVML reports as local to GPU `device_index` (same NUMA node / PCIe root).
    The end-to-end path moves 0.5 - 2 GB of results per step and GPU through pinned host buffers and stages pageable inputs with
    host memcpys; with eight ranks on a two-socket box, unpinned ranks allocate and copy across the socket interconnect.
    Best effort: returns False (and changes nothing) when NVML or the affinity call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = device_index
        if visible:
            ids = [v.strip() for v in visible.split(",") if v.strip()]
            if device_index < len(ids) and ids[device_index].isdigit():
                phys = int(ids[device_index])
        handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:  # noqa: BLE001 - an optimisation only
        return False
