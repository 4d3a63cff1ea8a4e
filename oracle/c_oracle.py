"""ctypes binding of oracle/pw_oracle.c -- TEST INFRASTRUCTURE (see pushworld_oracle.py)."""

import ctypes
import os
import subprocess

import numpy as np

from pushworld_oracle import OraclePuzzle

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libpw_oracle.so")
_LIB = None


def build(force=False):
    """gcc build of the C restatement; serialised by a file lock and renamed into place so that
    concurrent callers (several test or bench processes) never load a half-written library."""
    import fcntl

    src = os.path.join(HERE, "pw_oracle.c")
    stale = lambda: not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src)
    if not force and not stale():
        return SO
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if force or stale():
                tmp = f"libpw_oracle.so.tmp.{os.getpid()}"
                subprocess.check_call(["make", "-s", "-B", "-C", HERE, f"OUT={tmp}", "libpw_oracle.so"])
                os.replace(os.path.join(HERE, tmp), SO)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return SO


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.pwo_create.restype = ctypes.c_void_p
        _LIB.pwo_step.restype = ctypes.c_uint64
        _LIB.pwo_step.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        _LIB.pwo_achieved.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        _LIB.pwo_free.argtypes = [ctypes.c_void_p]
        _LIB.pwo_rollout.restype = ctypes.c_int64
        _LIB.pwo_rollout.argtypes = [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
        ]
        _LIB.pwo_expand.restype = None
        _LIB.pwo_expand.argtypes = [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_int,
        ]
    return _LIB


def _ints(values):
    return np.ascontiguousarray(np.array(values, dtype=np.int32).reshape(-1))


class COraclePuzzle:
    """A puzzle parsed by the Python oracle and stepped by the C restatement."""

    def __init__(self, text: str):
        self.py = OraclePuzzle(text)
        p = self.py
        ncells = _ints([len(s) for s in p.shapes])
        cells = _ints([c for s in p.shapes for c in sorted(s)])
        walls = _ints(sorted(p.walls))
        awalls = _ints(sorted(p.agent_walls))
        ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        self.n = p.num_objects
        self._h = lib().pwo_create(
            p.width, p.height, p.num_objects, p.num_goals, ptr(ncells), ptr(cells),
            ptr(_ints(p.initial_state)), ptr(_ints(p.goal_state) if p.num_goals else _ints([0])),
            len(p.walls), ptr(walls), len(p.agent_walls), ptr(awalls),
        )
        assert self._h, "pwo_create failed"

    def __del__(self):
        if getattr(self, "_h", None):
            lib().pwo_free(self._h)
            self._h = None

    def step(self, state, action):
        """-> (next_state tuple, moved mask)"""
        xy = _ints(state)
        moved = lib().pwo_step(self._h, xy.ctypes.data_as(ctypes.c_void_p), int(action))
        return tuple((int(xy[2 * k]), int(xy[2 * k + 1])) for k in range(self.n)), int(moved)

    def achieved(self, state):
        xy = _ints(state)
        return lib().pwo_achieved(self._h, xy.ctypes.data_as(ctypes.c_void_p))

    def rollout(self, x, y, actions, reward=None, terminated=None, threads=0):
        """x, y: int32 [n, N] (modified in place); actions uint8 [T, N]."""
        n_envs = x.shape[1]
        assert x.flags.c_contiguous and y.flags.c_contiguous and actions.flags.c_contiguous
        ptr = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
        return lib().pwo_rollout(self._h, n_envs, actions.shape[0], ptr(x), ptr(y), n_envs,
                                 ptr(actions), ptr(reward), ptr(terminated), threads)


    def expand(self, x, y, threads=0):
        """x, y: int32 [n, N] -> (x_out [n, 4N], y_out [n, 4N], moved uint32 [4N], goal uint8 [4N]);
        successor a * N + e = action a applied to state e."""
        n_states = x.shape[1]
        assert x.flags.c_contiguous and y.flags.c_contiguous and x.shape[0] == self.n
        x_out = np.empty((self.n, 4 * n_states), np.int32)
        y_out = np.empty((self.n, 4 * n_states), np.int32)
        moved = np.empty(4 * n_states, np.uint32)
        goal = np.empty(4 * n_states, np.uint8)
        ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        lib().pwo_expand(self._h, n_states, ptr(x), ptr(y), n_states, ptr(x_out), ptr(y_out),
                         4 * n_states, ptr(moved), ptr(goal), threads)
        return x_out, y_out, moved, goal


def max_threads():
    return lib().pwo_max_threads()
