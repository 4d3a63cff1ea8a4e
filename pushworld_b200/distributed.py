"""Multi-GPU plumbing: environments are independent, so the batch is split contiguously over
ranks (one process per GPU) and the only collective is ONE all-reduce (sum) of the small int64
counter vector a rollout produces.  ``torch.distributed`` with the NCCL
backend on GPUs; the same code runs over gloo with CPU tensors in the CPU test-suite."""

from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of ``total`` items owned by ``rank``; sizes differ by <= 1."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError("bad rank / world_size")
    base, extra = divmod(total, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def init_from_env(backend: str = "nccl") -> Tuple[int, int, int]:
    """Initialises the default process group from torchrun's environment variables; returns
    (rank, local_rank, world_size).  A single process without RANK set is (0, 0, 1)."""
    if "RANK" not in os.environ or int(os.environ.get("WORLD_SIZE", "1")) == 1:
        return 0, int(os.environ.get("LOCAL_RANK", "0")), 1
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    if backend == "nccl":
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        kwargs = {}
        if backend == "nccl":
            kwargs["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, local, world


def allreduce_counters(counters: torch.Tensor) -> torch.Tensor:
    """In-place sum over ranks of the rollout counter vector (no-op without a process group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return counters


def max_over_ranks(value: float, device=None) -> float:
    """Max of a python float over ranks (used for device-timed durations)."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier() -> None:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
