"""``PushWorldVectorEnv``: N PushWorld environments stepped together on one GPU.

The batched counterpart of the reference's single-environment wrappers: the same reward /
terminated / truncated semantics as ``PushWorldEnv.step`` (python3/src/pushworld/gym_env.py:
188-226) and the same padded float32 observation (utils/env_utils.py:44-91), for N
environments per call.  As in the reference there is no implicit reset: stepping after
``terminated`` keeps working (test_gym_env.py:89-102) unless ``autoreset=True``.

All tensors live on the GPU; ``rollout_host`` is the host-buffer entry point (pinned host
memory in, pinned host memory out, copies pipelined against the kernel on separate streams).
"""

from __future__ import annotations

import os
import random
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import data as _data
from .engine import Engine, StepOutput
from .utils.filesystem import iter_files_with_extension

PUZZLE_EXTENSION = ".pwp"


@dataclass
class RolloutOutput:
    reward: Optional[torch.Tensor]  # float32 [T, N]
    terminated: Optional[torch.Tensor]  # uint8 [T, N]
    dgoals: Optional[torch.Tensor] = None
    truncated: Optional[torch.Tensor] = None
    return_sum: Optional[torch.Tensor] = None  # float32 [N]
    counters: Optional[torch.Tensor] = None  # int64 [4]
    record: Optional[torch.Tensor] = None  # uint8 [T, N] packed records (rollout(packed=True))
    global_counters: Optional[torch.Tensor] = None  # int64 [4]: `counters` summed over all ranks (shard=True)


def load_puzzle_texts(puzzle_path: str) -> List[str]:
    """Reads every ``.pwp`` under ``puzzle_path`` in the reference's discovery order
    (``os.walk``, utils/filesystem.py:21-61)."""
    texts = []
    for path in iter_files_with_extension(puzzle_path, PUZZLE_EXTENSION):
        with open(path, "r") as f:
            texts.append(f.read())
    return texts


class PushWorldVectorEnv:
    """N environments over a set of puzzles.

    Args:
        puzzles: a path (file or directory, as ``PushWorldEnv``'s ``puzzle_path``) or a list of
            ``.pwp`` texts.
        num_envs: N.
        puzzle_ids: per-environment puzzle index (sequence / tensor of length N), an int (all
            environments share one puzzle -> the shared-memory table fast path), or None:
            environment e gets puzzle ``e % num_puzzles`` and environments are then GROUPED by
            puzzle (stable sort) so that a warp works on one puzzle; ``env_index`` maps slots
            back to the round-robin numbering.
        max_steps, border_width, pixels_per_cell, standard_padding: as ``PushWorldEnv``
            (gym_env.py:57-64).
        observation: None (no rendering), "float32" (the reference observation) or "uint8".
        reuse_step_outputs: ``step`` returns the same output tensors on every call (see ``step``).
        shard: one process per GPU (``torch.distributed``): ``num_envs`` / ``puzzle_ids`` describe
            the GLOBAL batch and this process owns the contiguous slice
            ``distributed.shard_range(num_envs, rank, world_size)`` of it (``shard_lo``, ``shard_hi``;
            ``num_envs`` becomes the local count, ``global_num_envs`` the total).  Environments
            never interact, so nothing is exchanged except ONE all-reduce of the int64[4] counter
            vector per ``rollout`` (``RolloutOutput.global_counters``), SURVEY.md 8e.
    """

    def __init__(
        self,
        puzzles: Union[str, Sequence[str]],
        num_envs: int,
        puzzle_ids: Union[None, int, Sequence[int], torch.Tensor] = None,
        max_steps: Optional[int] = None,
        border_width: int = 2,
        pixels_per_cell: int = 20,
        standard_padding: bool = False,
        observation: Optional[str] = None,
        device: Union[int, str, torch.device, None] = None,
        autoreset: bool = False,
        reuse_step_outputs: bool = False,
        shard: bool = False,
        check_rollout_actions: bool = False,
    ) -> None:
        texts = load_puzzle_texts(puzzles) if isinstance(puzzles, str) else list(puzzles)
        if len(texts) == 0:
            raise ValueError(f"No PushWorld puzzles found in: {puzzles}")
        if border_width < 1:
            raise ValueError("border_width must be >= 1")
        if pixels_per_cell < 3:
            raise ValueError("pixels_per_cell must be >= 3")
        if observation not in (None, "float32", "uint8"):
            raise ValueError("observation must be None, 'float32' or 'uint8'")
        self.engine = Engine(texts, device=device)
        self.device = self.engine.device
        self.global_num_envs = int(num_envs)
        self.sharded = bool(shard)
        if shard:
            from . import distributed as _dist

            rank, world = _dist.rank_and_world()
            self.shard_lo, self.shard_hi = _dist.shard_range(self.global_num_envs, rank, world)
        else:
            self.shard_lo, self.shard_hi = 0, self.global_num_envs
        self.num_envs = self.shard_hi - self.shard_lo
        self.max_steps = max_steps
        self.border_width = border_width
        self.pixels_per_cell = pixels_per_cell
        self.observation = observation
        self.autoreset = autoreset
        # per-step loops: bind the native call's arguments once and write every step's outputs into
        # the same tensors (Engine.prepare_step: ~2x fewer microseconds per step() call)
        self.reuse_step_outputs = reuse_step_outputs
        # rollouts count out-of-range actions as no-ops; True = poll after every rollout (this
        # synchronises) and raise the reference's ValueError like step() does
        self.check_rollout_actions = check_rollout_actions
        self._status_dirty = False  # a rollout may have left counts in the engine's status block

        max_h, max_w = self.engine.max_height, self.engine.max_width
        if standard_padding:  # gym_env.py:86-105
            std_h, std_w = _data.max_puzzle_dimensions()
            if std_h < max_h or std_w < max_w:
                raise ValueError(
                    "`standard_padding` is True, but the benchmark's maximum puzzle size is "
                    "smaller than the puzzle(s) in the given `puzzle_path`."
                )
            max_h, max_w = std_h, std_w
        self.pad_cells = (max_h, max_w)

        n = self.num_envs
        lo, hi = self.shard_lo, self.shard_hi
        if isinstance(puzzle_ids, int):
            self._uniform: Optional[int] = puzzle_ids
            self.puzzle_id = None
            self.env_index = torch.arange(lo, hi, device=self.device)
        else:
            self._uniform = None
            from . import distributed as _dist

            rank, world = _dist.rank_and_world() if shard else (0, 1)
            ids, env_index, _, _ = _dist.shard_puzzle_ids(self.global_num_envs, self.engine.num_puzzles,
                                                          puzzle_ids, rank, world)
            self.env_index = env_index.to(self.device)
            if n and (ids.min() < 0 or ids.max() >= self.engine.num_puzzles):
                raise ValueError("puzzle id out of range")
            self.puzzle_id = ids.to(torch.int32).to(self.device)
        self.xy = self.engine.new_state(n)
        self.steps = torch.zeros(n, dtype=torch.int32, device=self.device)
        self._is_reset = False
        self._rng = random.Random(123)  # gym_env.py:109

    # ------------------------------------------------------------------ helpers
    @property
    def _sel(self):
        return self._uniform if self._uniform is not None else self.puzzle_id

    @property
    def observation_shape(self) -> Tuple[int, int, int]:
        return (self.pad_cells[0] * self.pixels_per_cell, self.pad_cells[1] * self.pixels_per_cell, 3)

    def _observe(self, out: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
        if self.observation is None:
            return None
        return self.engine.render(self.xy, self._sel, dtype=self.observation,
                                  pixels_per_cell=self.pixels_per_cell,
                                  border_width=self.border_width, pad_cells=self.pad_cells, out=out)

    def _require_reset(self) -> None:
        if not self._is_reset:
            raise RuntimeError("reset() must be called before step() can be called.")

    def states(self) -> List[Tuple[Tuple[int, int], ...]]:
        """Current states as python tuples (host copy; for inspection and tests)."""
        host = self.xy.cpu().numpy()
        if self._uniform is not None:
            counts = [self.engine.parsed[self._uniform].num_objects] * self.num_envs
        else:
            per_puzzle = [p.num_objects for p in self.engine.parsed]
            counts = [per_puzzle[i] for i in self.puzzle_id.cpu().tolist()]
        return [
            tuple((int(host[0, j, e]), int(host[1, j, e])) for j in range(counts[e]))
            for e in range(self.num_envs)
        ]

    # ------------------------------------------------------------------ gym-like surface
    def reset(self, seed: Optional[int] = None, mask: Optional[torch.Tensor] = None):
        """All (or the masked) environments go back to their puzzle's initial state."""
        if seed is not None:
            self._rng = random.Random(seed)
        self.engine.reset(self.xy, self._sel, steps=self.steps, mask=mask)
        self._is_reset = True
        return self._observe()

    def step(self, actions: torch.Tensor, check_actions: bool = True):
        """-> (observation or None, reward f32 [N], terminated u8 [N], truncated u8 [N], info)

        With ``reuse_step_outputs=True`` (constructor) and no observations the returned tensors
        are REUSED by the next ``step`` call (clone them to keep them)."""
        self._require_reset()
        actions = self._as_actions(actions, (self.num_envs,))
        if check_actions and self._status_dirty:
            # invalid actions an earlier (unchecked) rollout counted must not fail THIS step
            self.engine.clear_status()
            self._status_dirty = False
        if self.observation is None and not self.reuse_step_outputs:
            out = self.engine.step(self.xy, self._sel, actions, steps=self.steps,
                                   max_steps=self.max_steps)
            obs = None
        elif self.observation is None:
            stepper = getattr(self, "_stepper", None)
            if stepper is None:  # arguments bound once (Engine.prepare_step); outputs are reused
                stepper = self._stepper = self.engine.prepare_step(
                    self.xy, self._sel, steps=self.steps, max_steps=self.max_steps)
            out = stepper(actions)
            obs = None
        else:
            obs, out = self.engine.step_render(
                self.xy, self._sel, actions, dtype=self.observation,
                pixels_per_cell=self.pixels_per_cell, border_width=self.border_width,
                pad_cells=self.pad_cells, steps=self.steps, max_steps=self.max_steps)
        if check_actions and self.engine.poll_invalid_actions():
            raise ValueError("The provided action is not in the action space.")  # gym_env.py:195
        return obs, out.reward, out.terminated, out.truncated, {"dgoals": out.dgoals}

    def capture_step(self) -> "GraphedStep":
        """One ``step`` (plus the observation render, if enabled) captured as a CUDA graph: the
        returned object owns a static ``actions`` tensor (uint8 [N]; write the policy's actions into
        it in place) and static output tensors; ``replay()`` enqueues the whole step with ONE launch
        call.  Invalid actions are counted, not raised (poll ``engine.poll_invalid_actions()``)."""
        self._require_reset()
        return GraphedStep(self)

    def _as_actions(self, actions, shape) -> torch.Tensor:
        if not isinstance(actions, torch.Tensor):
            actions = torch.as_tensor(np.asarray(actions))
        if actions.dtype != torch.uint8:
            if actions.numel() and (actions.min() < 0 or actions.max() > 255):
                raise ValueError("The provided action is not in the action space.")
            actions = actions.to(torch.uint8)
        actions = actions.to(self.device).contiguous()
        if tuple(actions.shape) != tuple(shape):
            raise ValueError(f"actions must have shape {tuple(shape)}")
        return actions

    def rollout(self, actions: torch.Tensor, want_reward: bool = True, want_terminated: bool = True,
                want_dgoals: bool = False, want_truncated: bool = False, want_return: bool = False,
                out: Optional[RolloutOutput] = None, packed: bool = False) -> RolloutOutput:
        """T fused steps in one launch; ``actions`` uint8 ``[T, N]`` on the device.  With only
        reward + terminated requested and no ``max_steps`` / ``autoreset`` the engine runs its
        lean rollout loop (every optional output compiled out)."""
        self._require_reset()
        if actions.dim() != 2:
            raise ValueError("actions must have shape [T, N]")
        actions = self._as_actions(actions, (actions.shape[0], self.num_envs))
        t, n = actions.shape
        dev = self.device
        if out is None and packed:
            out = RolloutOutput(reward=None, terminated=None,
                                record=torch.empty((t, n), dtype=torch.uint8, device=dev),
                                return_sum=torch.zeros(n, dtype=torch.float32, device=dev) if want_return else None,
                                counters=torch.zeros(4, dtype=torch.int64, device=dev))
        if out is None:
            out = RolloutOutput(
                reward=torch.empty((t, n), dtype=torch.float32, device=dev) if want_reward else None,
                terminated=torch.empty((t, n), dtype=torch.uint8, device=dev) if want_terminated else None,
                dgoals=torch.empty((t, n), dtype=torch.int8, device=dev) if want_dgoals else None,
                truncated=torch.empty((t, n), dtype=torch.uint8, device=dev) if want_truncated else None,
                return_sum=torch.zeros(n, dtype=torch.float32, device=dev) if want_return else None,
                counters=torch.zeros(4, dtype=torch.int64, device=dev),
            )
        track_steps = self.max_steps is not None or self.autoreset
        if out.record is not None:
            self.engine.rollout_packed(self.xy, self._sel, actions, out.record,
                                       steps=self.steps if track_steps else None,
                                       max_steps=self.max_steps, return_sum=out.return_sum,
                                       counters=out.counters, autoreset=self.autoreset)
        else:
            self.engine.rollout(self.xy, self._sel, actions, reward=out.reward, terminated=out.terminated,
                                dgoals=out.dgoals, truncated=out.truncated,
                                steps=self.steps if track_steps else None,
                                max_steps=self.max_steps, return_sum=out.return_sum,
                                counters=out.counters, autoreset=self.autoreset)
        if not track_steps:
            self.steps += actions.shape[0]
        self._poll_after_rollout()
        if self.sharded and out.counters is not None:
            # the path's single collective: one all-reduce (sum) of int64[4] per rollout
            from . import distributed as _dist

            if out.global_counters is None:
                out.global_counters = torch.empty_like(out.counters)
            out.global_counters.copy_(out.counters)
            _dist.allreduce_counters(out.global_counters)
        return out

    def _poll_after_rollout(self) -> None:
        """Rollouts treat out-of-range actions as counted no-ops (the launch is asynchronous); with
        ``check_rollout_actions`` set the count is polled here (synchronises) and raises the
        reference's ValueError (gym_env.py:195-196).  Either way the count never leaks into a
        later ``step`` call: ``step`` only looks at actions counted by its own launch."""
        if self.check_rollout_actions:
            if self.engine.poll_invalid_actions():
                raise ValueError("The provided action is not in the action space.")
        else:
            self._status_dirty = True

    # ------------------------------------------------------------------ host-buffer entry point
    def rollout_host(self, actions: torch.Tensor, reward: torch.Tensor, terminated: torch.Tensor,
                     chunk_steps: int = 64) -> None:
        """``actions`` (uint8 [T, N]), ``reward`` (float32 [T, N]) and ``terminated`` (uint8
        [T, N]) are PINNED HOST tensors.  The T steps are cut into chunks; chunk k's
        host->device copy, chunk k-1's kernel and chunk k-2's device->host copies run
        concurrently on three streams (both PCIe directions stay busy).  Blocks until the
        outputs are in host memory."""
        self._require_reset()
        t, n = actions.shape
        assert n == self.num_envs and reward.shape == (t, n) and terminated.shape == (t, n)
        assert actions.is_pinned() and reward.is_pinned() and terminated.is_pinned()
        pipe = self._host_pipe(chunk_steps)
        pipe.run(self, actions, reward, terminated)
        self._poll_after_rollout()

    def rollout_host_packed(self, actions: torch.Tensor, record: torch.Tensor,
                            chunk_steps: int = 64) -> None:
        """``rollout_host`` with ONE byte out per environment step: ``record`` (uint8 [T, N],
        pinned host memory) receives the packed per-step records of ``pw_rollout_packed``
        (delta_goals, terminated, truncated); ``engine.record_to_outputs(record)`` expands them
        to the float32 rewards / flags ``rollout_host`` returns, bit for bit.  PCIe traffic is
        1 byte in + 1 byte out per environment step instead of 1 + 5."""
        self._require_reset()
        t, n = actions.shape
        assert n == self.num_envs and record.shape == (t, n) and record.dtype == torch.uint8
        assert actions.is_pinned() and record.is_pinned()
        pipe = self._host_pipe(chunk_steps)
        pipe.run(self, actions, None, None, record=record)
        self._poll_after_rollout()

    def rollout_host_returns(self, actions: torch.Tensor, returns: torch.Tensor,
                             n_terminated: torch.Tensor, chunk_steps: int = 64) -> None:
        """Host-buffer rollout that brings back only per-environment SUMMARIES: ``returns``
        (float32 [N], sum of the T rewards) and ``n_terminated`` (int32 [N], number of steps with
        terminated set) -- pinned host tensors, like ``actions`` (uint8 [T, N]).  The per-step
        reward / terminated arrays stay on the device (``self.last_reward`` / ``self.last_terminated``
        hold the last chunk), so the PCIe traffic is 1 byte per env-step in and 8 bytes per
        environment out."""
        self._require_reset()
        t, n = actions.shape
        assert n == self.num_envs and returns.shape == (n,) and n_terminated.shape == (n,)
        assert actions.is_pinned() and returns.is_pinned() and n_terminated.is_pinned()
        pipe = self._host_pipe(chunk_steps)
        pipe.run_returns(self, actions, returns, n_terminated)
        self._poll_after_rollout()

    def _host_pipe(self, chunk_steps: int):
        pipe = getattr(self, "_pipe", None)
        if pipe is None or pipe.chunk != chunk_steps:
            pipe = _HostPipe(self.device, self.num_envs, chunk_steps)
            self._pipe = pipe
        return pipe


class GraphedStep:
    """A vector-environment step as a replayable CUDA graph (see ``PushWorldVectorEnv.capture_step``)."""

    def __init__(self, env: "PushWorldVectorEnv") -> None:
        dev, n = env.device, env.num_envs
        self.env = env
        self.actions = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.out = StepOutput(
            reward=torch.empty(n, dtype=torch.float32, device=dev),
            dgoals=torch.empty(n, dtype=torch.int8, device=dev),
            terminated=torch.empty(n, dtype=torch.uint8, device=dev),
            truncated=torch.empty(n, dtype=torch.uint8, device=dev),
        )
        self.observation = None
        if env.observation is not None:
            tdtype = torch.float32 if env.observation == "float32" else torch.uint8
            self.observation = torch.empty((n,) + env.observation_shape, dtype=tdtype, device=dev)
        self._stream = torch.cuda.Stream(dev)
        saved = (env.xy.clone(), env.steps.clone())
        self._stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self._stream):
            self._enqueue()  # warm-up outside the capture
        self._stream.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self._stream):
            self._enqueue()
        self._stream.synchronize()
        env.xy.copy_(saved[0])  # the warm-up stepped the environments: put them back
        env.steps.copy_(saved[1])
        torch.cuda.current_stream(dev).synchronize()

    def _enqueue(self) -> None:
        env = self.env
        if self.observation is None:
            env.engine.step(env.xy, env._sel, self.actions, steps=env.steps, max_steps=env.max_steps, out=self.out)
        else:
            env.engine.step_render(env.xy, env._sel, self.actions, dtype=env.observation,
                                   pixels_per_cell=env.pixels_per_cell, border_width=env.border_width,
                                   pad_cells=env.pad_cells, steps=env.steps, max_steps=env.max_steps,
                                   out=self.observation, step_out=self.out)

    def replay(self):
        """-> (observation or None, reward, terminated, truncated): static tensors, overwritten by
        the next replay.  Runs on the current stream."""
        self.graph.replay()
        return self.observation, self.out.reward, self.out.terminated, self.out.truncated


class _HostPipe:
    """Three-stream software pipeline for ``rollout_host`` with double-buffered device staging."""

    DEPTH = 3

    def __init__(self, device, n: int, chunk: int):
        self.chunk = chunk
        self.s_in = torch.cuda.Stream(device)
        self.s_run = torch.cuda.Stream(device)
        self.s_out = torch.cuda.Stream(device)
        self._mk = lambda dt: [torch.empty((chunk, n), dtype=dt, device=device) for _ in range(self.DEPTH)]
        self.d_act = self._mk(torch.uint8)
        self._d_rew = self._d_term = self._d_rec = None  # allocated by the first run that needs them
        self.ev_in = [torch.cuda.Event() for _ in range(self.DEPTH)]
        self.ev_run = [torch.cuda.Event() for _ in range(self.DEPTH)]
        self.ev_out = [torch.cuda.Event() for _ in range(self.DEPTH)]

    @property
    def d_rew(self):
        if self._d_rew is None:
            self._d_rew = self._mk(torch.float32)
        return self._d_rew

    @property
    def d_term(self):
        if self._d_term is None:
            self._d_term = self._mk(torch.uint8)
        return self._d_term

    @property
    def d_rec(self):
        if self._d_rec is None:
            self._d_rec = self._mk(torch.uint8)
        return self._d_rec

    def run(self, env: "PushWorldVectorEnv", actions, reward, terminated, record=None) -> None:
        t = actions.shape[0]
        cur = torch.cuda.current_stream(env.device)
        for s in (self.s_in, self.s_run, self.s_out):
            s.wait_stream(cur)
        n_chunks = (t + self.chunk - 1) // self.chunk
        for k in range(n_chunks):
            b = k % self.DEPTH
            lo, hi = k * self.chunk, min(t, (k + 1) * self.chunk)
            m = hi - lo
            with torch.cuda.stream(self.s_in):
                if k >= self.DEPTH:
                    self.s_in.wait_event(self.ev_run[b])  # kernel k-DEPTH done with d_act[b]
                self.d_act[b][:m].copy_(actions[lo:hi], non_blocking=True)
                self.ev_in[b].record(self.s_in)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(self.ev_in[b])
                if k >= self.DEPTH:
                    self.s_run.wait_event(self.ev_out[b])  # outputs of chunk k-DEPTH copied out
                track = env.max_steps is not None or env.autoreset
                if record is not None:
                    env.engine.rollout_packed(env.xy, env._sel, self.d_act[b][:m], self.d_rec[b][:m],
                                              steps=env.steps if track else None,
                                              max_steps=env.max_steps, autoreset=env.autoreset)
                else:
                    env.engine.rollout(env.xy, env._sel, self.d_act[b][:m], reward=self.d_rew[b][:m],
                                       terminated=self.d_term[b][:m], steps=env.steps if track else None,
                                       max_steps=env.max_steps, autoreset=env.autoreset)
                if not track:
                    env.steps += m
                self.ev_run[b].record(self.s_run)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_run[b])
                if record is not None:
                    record[lo:hi].copy_(self.d_rec[b][:m], non_blocking=True)
                else:
                    reward[lo:hi].copy_(self.d_rew[b][:m], non_blocking=True)
                    terminated[lo:hi].copy_(self.d_term[b][:m], non_blocking=True)
                self.ev_out[b].record(self.s_out)
        cur.wait_stream(self.s_out)
        cur.wait_stream(self.s_run)
        cur.synchronize()

    def run_returns(self, env: "PushWorldVectorEnv", actions, returns, n_terminated) -> None:
        t, n = actions.shape
        dev = env.device
        cur = torch.cuda.current_stream(dev)
        for s in (self.s_in, self.s_run):
            s.wait_stream(cur)
        if not hasattr(self, "d_ret"):
            self.d_ret = torch.zeros(n, dtype=torch.float32, device=dev)
            self.d_nterm = torch.zeros(n, dtype=torch.int32, device=dev)
        n_chunks = (t + self.chunk - 1) // self.chunk
        with torch.cuda.stream(self.s_run):
            self.d_ret.zero_()
            self.d_nterm.zero_()
        for k in range(n_chunks):
            b = k % self.DEPTH
            lo, hi = k * self.chunk, min(t, (k + 1) * self.chunk)
            m = hi - lo
            with torch.cuda.stream(self.s_in):
                if k >= self.DEPTH:
                    self.s_in.wait_event(self.ev_run[b])
                self.d_act[b][:m].copy_(actions[lo:hi], non_blocking=True)
                self.ev_in[b].record(self.s_in)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(self.ev_in[b])
                track = env.max_steps is not None or env.autoreset
                env.engine.rollout(env.xy, env._sel, self.d_act[b][:m], reward=self.d_rew[b][:m],
                                   terminated=self.d_term[b][:m], steps=env.steps if track else None,
                                   max_steps=env.max_steps, return_sum=self.d_ret,
                                   autoreset=env.autoreset)
                # per-environment count of terminated steps (a column sum; device-side plumbing)
                self.d_nterm += self.d_term[b][:m].sum(dim=0, dtype=torch.int32)
                if not track:
                    env.steps += m
                self.ev_run[b].record(self.s_run)
        with torch.cuda.stream(self.s_run):
            returns.copy_(self.d_ret, non_blocking=True)
            n_terminated.copy_(self.d_nterm, non_blocking=True)
        env.last_reward, env.last_terminated = self.d_rew[(n_chunks - 1) % self.DEPTH], self.d_term[(n_chunks - 1) % self.DEPTH]
        cur.wait_stream(self.s_run)
        cur.synchronize()
