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


def shard_puzzle_ids(num_envs: int, num_puzzles: int, puzzle_ids, rank: int, world_size: int):
    """The slice of a GLOBAL batch one rank owns: ``(ids, env_index, lo, hi)``.

    ``puzzle_ids`` None: global environment e plays puzzle ``e % num_puzzles`` and the batch is
    grouped by puzzle (stable sort) BEFORE it is cut, so a rank's slice covers a contiguous range
    of puzzles; ``env_index[k]`` is the global environment number of local slot k.  Otherwise
    ``puzzle_ids`` (length ``num_envs``) is cut as given.  Concatenating the slices of all ranks
    gives back the global layout (tests/test_distributed_cpu.py)."""
    lo, hi = shard_range(num_envs, rank, world_size)
    if puzzle_ids is None:
        ids = torch.arange(num_envs, dtype=torch.int64) % num_puzzles
        order = torch.argsort(ids, stable=True)
        return ids[order][lo:hi], order[lo:hi], lo, hi
    ids = torch.as_tensor(puzzle_ids, dtype=torch.int64).cpu()
    if ids.numel() != num_envs:
        raise ValueError("puzzle_ids must have num_envs entries")
    return ids[lo:hi], torch.arange(lo, hi, dtype=torch.int64), lo, hi


def rank_and_world() -> Tuple[int, int]:
    """(rank, world_size) of the default process group; (0, 1) without one."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


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


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Restricts this process to the CPUs of the NUMA node its GPU hangs off, so that pinned host
    buffers allocated afterwards (first touch) are local to that GPU's PCIe root: with one process
    per GPU every rank then streams through its own memory controller instead of all ranks
    sharing node 0.  Best effort: returns what it found / did and never raises."""
    info = {"numa_node": None, "cpus": None, "bound": False}
    try:
        import torch.cuda as tc

        bus = tc.get_device_properties(device_index).pci_bus_id
        dom = tc.get_device_properties(device_index).pci_domain_id
        dev = tc.get_device_properties(device_index).pci_device_id
        bdf = f"{dom:04x}:{bus:02x}:{dev:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        info["cpus"] = len(allowed)
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["bound"] = True
    except Exception as exc:  # no sysfs / no permission: keep the default placement
        info["error"] = repr(exc)
    return info


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
