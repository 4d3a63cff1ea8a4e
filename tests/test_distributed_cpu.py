"""Host-side multi-GPU logic on CPU: world_size-2 gloo process group (no GPU needed).  The data
path has no collective (environments are independent); the only exchange is the all-reduce of
the int64 counter vector and the max-over-ranks of the timed duration."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pushworld_b200 import distributed


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 262144, 10**7 + 3):
        for world in (1, 2, 3, 8):
            spans = [distributed.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        distributed.shard_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total_envs, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, _, w = distributed.init_from_env("gloo")
    assert (r, w) == (rank, world)
    lo, hi = distributed.shard_range(total_envs, rank, world)
    steps = 17
    # what a rank's rollout would have counted: env-steps, terminated, moved, goals
    counters = torch.tensor([(hi - lo) * steps, rank + 1, 10 * (rank + 1), 0], dtype=torch.int64)
    distributed.allreduce_counters(counters)
    slowest = distributed.max_over_ranks(1.0 + rank)
    distributed.barrier()
    assert distributed.rank_and_world() == (rank, world)
    # the sharded vector environment's layout: every rank's slice of the global batch
    mine = []
    for ids in (None, [(7 * e) % 5 for e in range(total_envs)]):
        got_ids, env_index, slo, shi = distributed.shard_puzzle_ids(total_envs, 5, ids, rank, world)
        assert (slo, shi) == (lo, hi) and len(got_ids) == hi - lo == len(env_index)
        mine.append((got_ids.tolist(), env_index.tolist()))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        for case, ids in enumerate((None, [(7 * e) % 5 for e in range(total_envs)])):
            all_ids = sum((g[case][0] for g in gathered), [])
            all_index = sum((g[case][1] for g in gathered), [])
            assert sorted(all_index) == list(range(total_envs))  # every environment exactly once
            want = [e % 5 for e in range(total_envs)] if ids is None else ids
            assert all(all_ids[k] == want[all_index[k]] for k in range(total_envs))
            if ids is None:
                assert all_ids == sorted(all_ids)  # grouped by puzzle before the cut
    if rank == 0:
        out.put((counters.tolist(), slowest))
    dist.destroy_process_group()


def test_two_rank_gloo_counter_allreduce():
    world, total = 2, 1001
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    counters, slowest = out.get()
    assert counters == [total * 17, 3, 30, 0]
    assert slowest == 2.0


def test_single_process_is_a_noop():
    env = {k: os.environ.pop(k, None) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    try:
        assert distributed.init_from_env("gloo") == (0, 0, 1)
        t = torch.tensor([5, 0, 0, 0])
        assert distributed.allreduce_counters(t).tolist() == [5, 0, 0, 0]
        assert distributed.max_over_ranks(3.5) == 3.5
    finally:
        for k, v in env.items():
            if v is not None:
                os.environ[k] = v
