"""``Engine``: PyTorch-facing wrapper over the C ABI (``include/pushworld_b200.h``).

PyTorch is used for device memory and streams only; every operation below is one call into
``libpushworld_b200.so`` that enqueues hand-written sm_100a kernels on the current stream.
State layout: ``xy`` int32 tensor of shape ``[2, n_max, N]`` (x planes then y planes,
structure of arrays; object order as in python3/src/pushworld/puzzle.py:170-257).
"""

from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import List, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib, tables

PuzzleSel = Union[int, torch.Tensor]


@dataclass
class StepOutput:
    reward: torch.Tensor  # float32 [N] or [T, N]
    dgoals: torch.Tensor  # int8
    terminated: torch.Tensor  # uint8
    truncated: torch.Tensor  # uint8
    moved: Optional[torch.Tensor] = None  # uint32 as int32 bit pattern, pw_step only


def draw_puzzle(seed: int, env: int, episode: int, num_puzzles: int) -> int:
    """Host mirror of the device generator used by pw_rollout_episodes (include/pushworld_b200.h)."""
    m = (1 << 64) - 1
    h = (seed ^ (env * 0x9E3779B97F4A7C15) ^ (episode * 0xD1B54A32D192ED03)) & m
    h = (h + 0x9E3779B97F4A7C15) & m
    h = ((h ^ (h >> 30)) * 0xBF58476D1CE4E5B9) & m
    h = ((h ^ (h >> 27)) * 0x94D049BB133111EB) & m
    h ^= h >> 31
    return ((h >> 32) * num_puzzles) >> 32


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _max_steps_arg(max_steps: Optional[int]) -> int:
    """The C ABI's ``max_steps`` (<= 0 = no limit).  The reference computes ``truncated = steps >=
    max_steps`` after incrementing ``steps`` (gym_env.py:201,223), so any limit <= 1 truncates
    every step: ``max_steps=0`` (or a negative value) is passed on as 1, not as "no limit"."""
    return 0 if max_steps is None else max(1, int(max_steps))


_RECORD_LUT = None


def record_to_outputs(record):
    """Expands packed per-step records (``pw_rollout_packed``: uint8, bits 0..5 = delta_goals + 32,
    bit 6 = truncated, bit 7 = terminated) into what ``PushWorldEnv.step`` returns
    (gym_env.py:210-226): ``(reward float32, terminated uint8, truncated uint8, dgoals int8)``.
    ``reward`` is bit-identical to the float32 rewards ``pw_rollout`` writes:
    ``float32(10.0)`` if terminated else ``float32(float64(delta_goals) - 0.01)``.
    Accepts a torch tensor (any device) or a numpy array; returns the same kind."""
    global _RECORD_LUT
    if _RECORD_LUT is None:
        codes = np.arange(256)
        dg = (codes & 0x3F).astype(np.float64) - 32.0
        _RECORD_LUT = np.where(codes & 0x80, np.float32(10.0), (dg - 0.01).astype(np.float32)).astype(np.float32)
    if isinstance(record, torch.Tensor):
        lut = torch.from_numpy(_RECORD_LUT).to(record.device)
        reward = lut[record.long()]
        return (reward, (record >> 7), (record >> 6) & 1, ((record & 0x3F).to(torch.int16) - 32).to(torch.int8))
    record = np.asarray(record, dtype=np.uint8)
    return (_RECORD_LUT[record], record >> 7, (record >> 6) & 1, ((record & 0x3F).astype(np.int16) - 32).astype(np.int8))


class Engine:
    """Immutable set of compiled puzzles resident on one CUDA device."""

    def __init__(self, puzzle_texts: Sequence[str], device: Union[int, str, torch.device, None] = None):
        if len(puzzle_texts) == 0:
            raise ValueError("an Engine needs at least one puzzle")
        if not torch.cuda.is_available():
            raise _lib.PushWorldNativeError(
                _lib.PW_E_CUDA, "no CUDA device: pushworld_b200 has no CPU fallback"
            )
        self._lib = _lib.load()
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.parsed: List[tables.ParsedPuzzle] = [tables.parse_puzzle(t) for t in puzzle_texts]
        blobs = [_lib.compile_puzzle_text(t) for t in puzzle_texts]  # native compiler
        offsets, pos = [], 0
        for b in blobs:
            offsets.append(pos // 8)
            pos += len(b)
        self._blob = np.frombuffer(b"".join(blobs), dtype=np.uint8).copy()
        self._offsets = np.array(offsets, dtype=np.uint32)
        handle = ctypes.c_void_p()
        _lib.check(
            self._lib.pw_engine_create(
                self._blob.ctypes.data, self._blob.nbytes, self._offsets.ctypes.data,
                len(blobs), self.device.index, ctypes.byref(handle),
            )
        )
        self._h = handle
        self.num_puzzles = len(blobs)
        self.n_max = max(p.num_objects for p in self.parsed)
        self.max_width = max(p.width for p in self.parsed)
        self.max_height = max(p.height for p in self.parsed)
        self.bad_puzzle_ids = 0  # out-of-range ids seen by polls that were asked about something else

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.pw_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _sel(self, puzzle: PuzzleSel, n: int):
        if isinstance(puzzle, torch.Tensor):
            self._chk(puzzle, torch.int32, (n,), "puzzle_id")
            return puzzle.data_ptr(), 0
        pid = int(puzzle)
        if not 0 <= pid < self.num_puzzles:
            raise ValueError(f"puzzle index {pid} out of range")
        return None, pid

    def _chk(self, t: torch.Tensor, dtype, shape, name: str) -> None:
        if t.device != self.device:
            raise ValueError(f"{name} must live on {self.device}, got {t.device}")
        if t.dtype != dtype:
            raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name} must have shape {tuple(shape)}, got {tuple(t.shape)}")
        if not t.is_contiguous():
            raise ValueError(f"{name} must be contiguous")

    def _chk_state(self, xy: torch.Tensor) -> int:
        if xy.dim() != 3 or xy.shape[0] != 2 or xy.shape[1] != self.n_max:
            raise ValueError(f"state must have shape [2, {self.n_max}, N], got {tuple(xy.shape)}")
        self._chk(xy, torch.int32, xy.shape, "state")
        return xy.shape[2]

    def new_state(self, n_envs: int) -> torch.Tensor:
        return torch.zeros((2, self.n_max, n_envs), dtype=torch.int32, device=self.device)

    def state_from_tuples(self, states: Sequence[Sequence[Sequence[int]]]) -> torch.Tensor:
        """[N] python states ((x, y), ...) -> device ``xy`` tensor (zero padded to n_max)."""
        n = len(states)
        host = np.zeros((2, self.n_max, n), dtype=np.int32)
        for e, s in enumerate(states):
            for j, (x, y) in enumerate(s):
                host[0, j, e] = x
                host[1, j, e] = y
        return torch.from_numpy(host).to(self.device)

    # ------------------------------------------------------------------ operations
    def reset(self, xy: torch.Tensor, puzzle: PuzzleSel, steps: Optional[torch.Tensor] = None,
              mask: Optional[torch.Tensor] = None) -> None:
        n = self._chk_state(xy)
        pid_ptr, pid = self._sel(puzzle, n)
        if steps is not None:
            self._chk(steps, torch.int32, (n,), "steps")
        if mask is not None:
            self._chk(mask, torch.uint8, (n,), "mask")
        _lib.check(
            self._lib.pw_reset(self._h, n, pid_ptr, pid, xy[0].data_ptr(), xy[1].data_ptr(), n,
                               _ptr(steps), _ptr(mask), self._stream())
        )

    def step(self, xy: torch.Tensor, puzzle: PuzzleSel, action: torch.Tensor,
             steps: Optional[torch.Tensor] = None, max_steps: Optional[int] = None,
             want_moved: bool = False, out: Optional[StepOutput] = None) -> StepOutput:
        n = self._chk_state(xy)
        pid_ptr, pid = self._sel(puzzle, n)
        self._chk(action, torch.uint8, (n,), "action")
        if steps is not None:
            self._chk(steps, torch.int32, (n,), "steps")
        if out is None:
            out = StepOutput(
                reward=torch.empty(n, dtype=torch.float32, device=self.device),
                dgoals=torch.empty(n, dtype=torch.int8, device=self.device),
                terminated=torch.empty(n, dtype=torch.uint8, device=self.device),
                truncated=torch.empty(n, dtype=torch.uint8, device=self.device),
                moved=torch.empty(n, dtype=torch.int32, device=self.device) if want_moved else None,
            )
        _lib.check(
            self._lib.pw_step(self._h, n, pid_ptr, pid, xy[0].data_ptr(), xy[1].data_ptr(), n,
                              action.data_ptr(), _ptr(out.reward), _ptr(out.dgoals),
                              _ptr(out.terminated), _ptr(out.truncated), _ptr(steps),
                              _max_steps_arg(max_steps), _ptr(out.moved), self._stream())
        )
        return out

    def prepare_step(self, xy: torch.Tensor, puzzle: PuzzleSel, out: Optional[StepOutput] = None,
                     steps: Optional[torch.Tensor] = None, max_steps: Optional[int] = None,
                     want_moved: bool = False):
        """``step`` with everything but the action bound once: returns ``f(action) -> StepOutput``.

        A per-step loop pays for argument checks, tensor views and pointer conversions on every
        ``step`` call (about 27 us, against 8 us for the kernel); the returned function only checks
        the action tensor and makes the native call.  ``xy``, ``steps`` and the output tensors must
        stay alive and unmoved for as long as the function is used."""
        n = self._chk_state(xy)
        pid_ptr, pid = self._sel(puzzle, n)
        if steps is not None:
            self._chk(steps, torch.int32, (n,), "steps")
        if out is None:
            out = StepOutput(
                reward=torch.empty(n, dtype=torch.float32, device=self.device),
                dgoals=torch.empty(n, dtype=torch.int8, device=self.device),
                terminated=torch.empty(n, dtype=torch.uint8, device=self.device),
                truncated=torch.empty(n, dtype=torch.uint8, device=self.device),
                moved=torch.empty(n, dtype=torch.int32, device=self.device) if want_moved else None,
            )
        keep = (xy, puzzle, steps, out)  # the pointers below must not outlive their tensors
        fn, handle, device = self._lib.pw_step, self._h, self.device
        head = (handle, n, pid_ptr, pid, xy[0].data_ptr(), xy[1].data_ptr(), n)
        tail = (_ptr(out.reward), _ptr(out.dgoals), _ptr(out.terminated), _ptr(out.truncated),
                _ptr(steps), _max_steps_arg(max_steps), _ptr(out.moved))
        current_stream = torch.cuda.current_stream
        shape = (n,)

        def run(action: torch.Tensor) -> StepOutput:
            if (action.dtype is not torch.uint8 or action.device != device or tuple(action.shape) != shape
                    or not action.is_contiguous()):
                self._chk(action, torch.uint8, shape, "action")
            rc = fn(*head, action.data_ptr(), *tail, current_stream(device).cuda_stream)
            if rc:
                _lib.check(rc)
            return out

        run.keep = keep
        return run

    def rollout(self, xy: torch.Tensor, puzzle: PuzzleSel, actions: torch.Tensor,
                reward: Optional[torch.Tensor] = None, terminated: Optional[torch.Tensor] = None,
                dgoals: Optional[torch.Tensor] = None, truncated: Optional[torch.Tensor] = None,
                steps: Optional[torch.Tensor] = None, max_steps: Optional[int] = None,
                return_sum: Optional[torch.Tensor] = None, counters: Optional[torch.Tensor] = None,
                autoreset: bool = False) -> None:
        """T fused steps; ``actions`` is uint8 ``[T, N]``; every output is optional and, when
        given, a preallocated ``[T, N]`` (``return_sum`` ``[N]``, ``counters`` int64 ``[4]``)."""
        n = self._chk_state(xy)
        pid_ptr, pid = self._sel(puzzle, n)
        if actions.dim() != 2 or actions.shape[1] != n:
            raise ValueError(f"actions must have shape [T, {n}]")
        t = actions.shape[0]
        self._chk(actions, torch.uint8, (t, n), "actions")
        for name, ten, dt in (("reward", reward, torch.float32), ("terminated", terminated, torch.uint8),
                              ("dgoals", dgoals, torch.int8), ("truncated", truncated, torch.uint8)):
            if ten is not None:
                self._chk(ten, dt, (t, n), name)
        if steps is not None:
            self._chk(steps, torch.int32, (n,), "steps")
        if return_sum is not None:
            self._chk(return_sum, torch.float32, (n,), "return_sum")
        if counters is not None:
            self._chk(counters, torch.int64, (4,), "counters")
        _lib.check(
            self._lib.pw_rollout(self._h, n, t, pid_ptr, pid, xy[0].data_ptr(), xy[1].data_ptr(), n,
                                 actions.data_ptr(), _ptr(reward), _ptr(dgoals), _ptr(terminated),
                                 _ptr(truncated), _ptr(steps), _max_steps_arg(max_steps), _ptr(return_sum),
                                 _ptr(counters), _lib.PW_FLAG_AUTORESET if autoreset else 0,
                                 self._stream())
        )

    def rollout_packed(self, xy: torch.Tensor, puzzle: PuzzleSel, actions: torch.Tensor,
                       record: torch.Tensor, steps: Optional[torch.Tensor] = None,
                       max_steps: Optional[int] = None, return_sum: Optional[torch.Tensor] = None,
                       counters: Optional[torch.Tensor] = None, autoreset: bool = False,
                       episode: Optional[torch.Tensor] = None, seed: Optional[int] = None) -> None:
        """``rollout`` writing ONE byte per environment step (``record`` uint8 ``[T, N]``, see
        ``record_to_outputs``) instead of reward / terminated / truncated / dgoals arrays
        (pw_rollout_packed).  With ``episode`` (int32 [N]) and ``seed`` the episode handling of
        ``rollout_episodes`` is done on the device (``puzzle`` must then be the int32 [N] tensor)."""
        n = self._chk_state(xy)
        pid_ptr, pid = self._sel(puzzle, n)
        if actions.dim() != 2 or actions.shape[1] != n:
            raise ValueError(f"actions must have shape [T, {n}]")
        t = actions.shape[0]
        self._chk(actions, torch.uint8, (t, n), "actions")
        self._chk(record, torch.uint8, (t, n), "record")
        if steps is not None:
            self._chk(steps, torch.int32, (n,), "steps")
        if return_sum is not None:
            self._chk(return_sum, torch.float32, (n,), "return_sum")
        if counters is not None:
            self._chk(counters, torch.int64, (4,), "counters")
        flags = _lib.PW_FLAG_AUTORESET if autoreset else 0
        if episode is not None:
            if pid_ptr is None or steps is None or seed is None:
                raise ValueError("episode handling needs a puzzle_id tensor, steps and a seed")
            self._chk(episode, torch.int32, (n,), "episode")
            flags |= _lib.PW_FLAG_RESAMPLE
        _lib.check(
            self._lib.pw_rollout_packed(self._h, n, t, pid_ptr, pid, xy[0].data_ptr(), xy[1].data_ptr(), n,
                                        actions.data_ptr(), record.data_ptr(), _ptr(steps),
                                        _max_steps_arg(max_steps), _ptr(return_sum), _ptr(counters),
                                        flags, int(seed or 0) & (2**64 - 1), _ptr(episode),
                                        self._stream())
        )

    def rollout_episodes(self, xy: torch.Tensor, puzzle_id: torch.Tensor, actions: torch.Tensor,
                         steps: torch.Tensor, episode: torch.Tensor, seed: int,
                         max_steps: Optional[int] = None, reward: Optional[torch.Tensor] = None,
                         terminated: Optional[torch.Tensor] = None, dgoals: Optional[torch.Tensor] = None,
                         truncated: Optional[torch.Tensor] = None,
                         return_sum: Optional[torch.Tensor] = None,
                         counters: Optional[torch.Tensor] = None) -> None:
        """T fused steps with on-device episode handling (pw_rollout_episodes): a terminated or
        truncated environment draws a new puzzle (see ``draw_puzzle``) and restarts; ``puzzle_id``
        (int32 [N]), ``steps`` (int32 [N]) and ``episode`` (int32 [N], used as uint32) are updated
        in place."""
        n = self._chk_state(xy)
        self._chk(puzzle_id, torch.int32, (n,), "puzzle_id")
        t = actions.shape[0]
        self._chk(actions, torch.uint8, (t, n), "actions")
        self._chk(steps, torch.int32, (n,), "steps")
        self._chk(episode, torch.int32, (n,), "episode")
        for name, ten, dt in (("reward", reward, torch.float32), ("terminated", terminated, torch.uint8),
                              ("dgoals", dgoals, torch.int8), ("truncated", truncated, torch.uint8)):
            if ten is not None:
                self._chk(ten, dt, (t, n), name)
        if return_sum is not None:
            self._chk(return_sum, torch.float32, (n,), "return_sum")
        if counters is not None:
            self._chk(counters, torch.int64, (4,), "counters")
        _lib.check(
            self._lib.pw_rollout_episodes(self._h, n, t, puzzle_id.data_ptr(), xy[0].data_ptr(),
                                          xy[1].data_ptr(), n, actions.data_ptr(), _ptr(reward),
                                          _ptr(dgoals), _ptr(terminated), _ptr(truncated),
                                          steps.data_ptr(), _max_steps_arg(max_steps), _ptr(return_sum),
                                          _ptr(counters), int(seed) & (2**64 - 1), episode.data_ptr(),
                                          self._stream())
        )

    def goals(self, xy: torch.Tensor, puzzle: PuzzleSel):
        """(achieved int32 [N], is_goal uint8 [N]) of the current states
        (count_achieved_goals / is_goal_state, puzzle.py:396-411)."""
        n = self._chk_state(xy)
        pid_ptr, pid = self._sel(puzzle, n)
        achieved = torch.empty(n, dtype=torch.int32, device=self.device)
        is_goal = torch.empty(n, dtype=torch.uint8, device=self.device)
        _lib.check(
            self._lib.pw_goals(self._h, n, pid_ptr, pid, xy[0].data_ptr(), xy[1].data_ptr(), n,
                               achieved.data_ptr(), is_goal.data_ptr(), self._stream())
        )
        return achieved, is_goal

    def expand(self, xy: torch.Tensor, puzzle: PuzzleSel, xy_out: Optional[torch.Tensor] = None,
               moved: Optional[torch.Tensor] = None, goal: Optional[torch.Tensor] = None):
        """4 successors per state: returns (xy_out [2, n_max, 4N], moved int32 [4N], goal uint8
        [4N]); successor ``a * N + e`` is action ``a`` applied to state ``e``."""
        n = self._chk_state(xy)
        pid_ptr, pid = self._sel(puzzle, n)
        if xy_out is None:
            xy_out = torch.empty((2, self.n_max, 4 * n), dtype=torch.int32, device=self.device)
        if moved is None:
            moved = torch.empty(4 * n, dtype=torch.int32, device=self.device)
        if goal is None:
            goal = torch.empty(4 * n, dtype=torch.uint8, device=self.device)
        self._chk(xy_out, torch.int32, (2, self.n_max, 4 * n), "xy_out")
        self._chk(moved, torch.int32, (4 * n,), "moved")
        self._chk(goal, torch.uint8, (4 * n,), "goal")
        _lib.check(
            self._lib.pw_expand(self._h, n, pid_ptr, pid, xy[0].data_ptr(), xy[1].data_ptr(), n,
                                xy_out[0].data_ptr(), xy_out[1].data_ptr(), 4 * n,
                                moved.data_ptr(), goal.data_ptr(), self._stream())
        )
        return xy_out, moved, goal

    def expand_indexed(self, arena: torch.Tensor, puzzle: int, index: torch.Tensor, out_base: int,
                       moved: torch.Tensor, goal: torch.Tensor) -> None:
        """Expands the states ``arena[:, :, index]`` of a state arena ``[2, n_max, capacity]`` and
        writes the ``4 * len(index)`` successors back into the arena at columns
        ``out_base + a * len(index) + i`` (pw_expand_indexed)."""
        cap = self._chk_state(arena)
        n = index.numel()
        self._chk(index, torch.int64, (n,), "index")
        self._chk(moved, torch.int32, (4 * n,), "moved")
        self._chk(goal, torch.uint8, (4 * n,), "goal")
        if out_base < 0 or out_base + 4 * n > cap:
            raise ValueError("successors do not fit the arena")
        x_out = arena[0].data_ptr() + 4 * out_base
        y_out = arena[1].data_ptr() + 4 * out_base
        _lib.check(
            self._lib.pw_expand_indexed(self._h, n, int(puzzle), index.data_ptr(), arena[0].data_ptr(),
                                        arena[1].data_ptr(), cap, x_out, y_out, cap,
                                        moved.data_ptr(), goal.data_ptr(), self._stream())
        )

    def visited_insert(self, arena: torch.Tensor, puzzle: int, first_id: int, count: int,
                       table: torch.Tensor, is_new: torch.Tensor,
                       valid: Optional[torch.Tensor] = None) -> None:
        """Exact duplicate detection (pw_visited_insert): ``is_new[i] = 1`` iff state
        ``first_id + i`` of the arena was not yet in ``table`` (int64 ``[capacity]``, power of
        two, zero-initialised)."""
        cap = self._chk_state(arena)
        self._chk(table, torch.int64, (table.numel(),), "table")
        self._chk(is_new, torch.uint8, (count,), "is_new")
        if valid is not None:
            self._chk(valid, torch.uint8, (count,), "valid")
        _lib.check(
            self._lib.pw_visited_insert(self._h, int(puzzle), arena[0].data_ptr(), arena[1].data_ptr(),
                                        cap, int(first_id), int(count), _ptr(valid), table.data_ptr(),
                                        table.numel(), is_new.data_ptr(), self._stream())
        )

    def novelty_table(self, puzzle: int) -> torch.Tensor:
        """A zeroed ``seen`` table (int32 words) for ``novelty`` on this puzzle."""
        words = ctypes.c_uint64(0)
        _lib.check(self._lib.pw_novelty_table_words(self._h, int(puzzle), ctypes.byref(words)))
        return torch.zeros(int(words.value), dtype=torch.int32, device=self.device)

    def novelty(self, arena: torch.Tensor, puzzle: int, first_id: int, count: int,
                moved: torch.Tensor, seen: torch.Tensor, out: Optional[torch.Tensor] = None,
                valid: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Novelty (1, 2 or 3) of ``count`` arena states given their moved-object masks, judged
        against ``seen`` as it was before the call; the states are then inserted (pw_novelty)."""
        cap = self._chk_state(arena)
        self._chk(moved, torch.int32, (count,), "moved")
        self._chk(seen, torch.int32, (seen.numel(),), "seen")
        if out is None:
            out = torch.empty(count, dtype=torch.uint8, device=self.device)
        self._chk(out, torch.uint8, (count,), "novelty")
        if valid is not None:
            self._chk(valid, torch.uint8, (count,), "valid")
        _lib.check(
            self._lib.pw_novelty(self._h, int(puzzle), arena[0].data_ptr(), arena[1].data_ptr(), cap,
                                 int(first_id), int(count), moved.data_ptr(), _ptr(valid),
                                 seen.data_ptr(), out.data_ptr(), self._stream())
        )
        return out

    def poll_status(self):
        """(invalid actions, visited-table overflows) since the last poll (synchronises)."""
        status = (ctypes.c_uint64 * 4)()
        _lib.check(self._lib.pw_engine_poll(self._h, status, self._stream()))
        self.bad_puzzle_ids += int(status[2])
        return int(status[0]), int(status[1])

    def poll_bad_puzzle_ids(self) -> int:
        """Number of per-environment puzzle ids outside ``[0, num_puzzles)`` the kernels have seen
        since the last call (such environments are run on puzzle 0 instead of reading the tables
        out of bounds).  Synchronises; also clears the other status words."""
        status = (ctypes.c_uint64 * 4)()
        _lib.check(self._lib.pw_engine_poll(self._h, status, self._stream()))
        count, self.bad_puzzle_ids = self.bad_puzzle_ids + int(status[2]), 0
        return count

    def _render_out(self, n, dtype, ppc, pad_cells, out):
        tdtype = torch.float32 if dtype == "float32" else torch.uint8
        shape = (n, pad_cells[0] * ppc, pad_cells[1] * ppc, 3)
        if out is None:
            out = torch.empty(shape, dtype=tdtype, device=self.device)
        self._chk(out, tdtype, shape, "out")
        return out, (_lib.PW_DTYPE_F32 if dtype == "float32" else _lib.PW_DTYPE_U8)

    def _pad_cells(self, puzzle: PuzzleSel, pad_cells):
        if pad_cells is not None:
            return int(pad_cells[0]), int(pad_cells[1])
        if isinstance(puzzle, torch.Tensor):
            return self.max_height, self.max_width
        p = self.parsed[int(puzzle)]
        return p.height, p.width

    def render(self, xy: torch.Tensor, puzzle: PuzzleSel, dtype: str = "uint8",
               pixels_per_cell: int = 20, border_width: int = 2, pad_cells=None,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``[N, pad_h*ppc, pad_w*ppc, 3]`` images; ``pad_cells=(rows, cols)`` (default: the
        puzzle's own size = PushWorldPuzzle.render)."""
        if dtype not in ("uint8", "float32"):
            raise ValueError("dtype must be 'uint8' or 'float32'")
        n = self._chk_state(xy)
        pid_ptr, pid = self._sel(puzzle, n)
        pad = self._pad_cells(puzzle, pad_cells)
        out, code = self._render_out(n, dtype, pixels_per_cell, pad, out)
        _lib.check(
            self._lib.pw_render(self._h, n, pid_ptr, pid, xy[0].data_ptr(), xy[1].data_ptr(), n,
                                out.data_ptr(), code, pixels_per_cell, border_width, pad[0], pad[1],
                                self._stream())
        )
        return out

    def step_render(self, xy: torch.Tensor, puzzle: PuzzleSel, action: torch.Tensor,
                    dtype: str = "float32", pixels_per_cell: int = 20, border_width: int = 2,
                    pad_cells=None, steps: Optional[torch.Tensor] = None,
                    max_steps: Optional[int] = None, out: Optional[torch.Tensor] = None,
                    step_out: Optional[StepOutput] = None):
        """``step`` + ``render`` of the new state in one native call (two launches on one stream)."""
        n = self._chk_state(xy)
        pid_ptr, pid = self._sel(puzzle, n)
        self._chk(action, torch.uint8, (n,), "action")
        pad = self._pad_cells(puzzle, pad_cells)
        out, code = self._render_out(n, dtype, pixels_per_cell, pad, out)
        if step_out is None:
            step_out = StepOutput(
                reward=torch.empty(n, dtype=torch.float32, device=self.device),
                dgoals=torch.empty(n, dtype=torch.int8, device=self.device),
                terminated=torch.empty(n, dtype=torch.uint8, device=self.device),
                truncated=torch.empty(n, dtype=torch.uint8, device=self.device),
            )
        _lib.check(
            self._lib.pw_step_render(self._h, n, pid_ptr, pid, xy[0].data_ptr(), xy[1].data_ptr(),
                                     n, action.data_ptr(), _ptr(step_out.reward),
                                     _ptr(step_out.dgoals), _ptr(step_out.terminated),
                                     _ptr(step_out.truncated), _ptr(steps), _max_steps_arg(max_steps),
                                     out.data_ptr(), code, pixels_per_cell, border_width, pad[0],
                                     pad[1], self._stream())
        )
        return out, step_out

    def clear_status(self) -> None:
        """Drops the invalid-action / overflow counts of earlier launches (asynchronous)."""
        _lib.check(self._lib.pw_engine_clear_status(self._h, self._stream()))

    def poll_invalid_actions(self) -> int:
        """Number of out-of-range actions seen since the last poll (synchronises)."""
        status = (ctypes.c_uint64 * 4)()
        _lib.check(self._lib.pw_engine_poll(self._h, status, self._stream()))
        self.bad_puzzle_ids += int(status[2])
        return int(status[0])
