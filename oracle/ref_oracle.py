"""ctypes binding of oracle/_ref/libpw_ref.so -- the REFERENCE's own C++ simulator
(cpp/src/pushworld_puzzle.cc) and novelty heuristic (cpp/src/heuristics/novelty.cc), compiled from
the reference sources by `make -C oracle _ref` (see oracle/Makefile, oracle/ref_wrap.cc).
TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
--impl reference legs may import this.

The C++ twin orders objects differently from the python reference (the parity target):
python: "a", the movables of the goal ids sorted DESCENDING, then the other movables in file order
(python3/src/pushworld/puzzle.py:178-179,227-237); C++: "a", the movables of the goal ids
ASCENDING, then the other movables ascending (cpp/src/pushworld_puzzle.cc:274-315, std::map order).
`RefPuzzle` converts: every method takes and returns states in the PYTHON order.

The built library travels to the GPU box; the reference sources do not, and are not needed there.
"""

import ctypes
import os
import subprocess
import tempfile

import numpy as np

from pushworld_oracle import OraclePuzzle

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_ref", "libpw_ref.so")
REFERENCE = os.environ.get("PUSHWORLD_REFERENCE", "/root/reference")
_LIB = None


def can_build() -> bool:
    return os.path.exists(os.path.join(REFERENCE, "cpp", "src", "pushworld_puzzle.cc"))


def build(force=False):
    """Compiles the reference's C++ sources (in place under /root/reference, build container only)
    into oracle/_ref/libpw_ref.so; serialised by a file lock, renamed into place."""
    import fcntl

    if os.path.exists(SO) and not force:
        return SO
    if not can_build():
        raise RuntimeError(f"{SO} is missing and the reference sources are not available to build it")
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    with open(os.path.join(HERE, ".build_ref.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if force or not os.path.exists(SO):
                tmp = f"_ref/libpw_ref.so.tmp.{os.getpid()}"
                subprocess.check_call(["make", "-s", "-B", "-C", HERE, f"REF_OUT={tmp}",
                                       f"REFERENCE={REFERENCE}", "_ref"])
                os.replace(os.path.join(HERE, tmp), SO)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return SO


def available() -> bool:
    return os.path.exists(SO) or can_build()


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        vp, i32, i64, u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint64
        _LIB.pwr_error.restype = ctypes.c_char_p
        _LIB.pwr_load.restype = vp
        _LIB.pwr_load.argtypes = [ctypes.c_char_p]
        _LIB.pwr_free.argtypes = [vp]
        _LIB.pwr_num_objects.argtypes = [vp]
        _LIB.pwr_num_goals.argtypes = [vp]
        _LIB.pwr_initial_state.argtypes = [vp, vp]
        _LIB.pwr_goal.argtypes = [vp, vp]
        _LIB.pwr_step.restype = u64
        _LIB.pwr_step.argtypes = [vp, vp, i32]
        _LIB.pwr_satisfies_goal.argtypes = [vp, vp]
        _LIB.pwr_rollout.restype = i64
        _LIB.pwr_rollout.argtypes = [vp, i64, i32, vp, vp, i64, vp, vp, vp, i32]
        _LIB.pwr_novelty_create.restype = vp
        _LIB.pwr_novelty_create.argtypes = [i32]
        _LIB.pwr_novelty_free.argtypes = [vp]
        _LIB.pwr_novelty.argtypes = [vp, vp, i32, u64]
    return _LIB


def cpp_object_order(py_names, goal_names):
    """Names in the C++ twin's state order, given the python order's names."""
    movable_of_goal = ["m" + g[1:] for g in sorted(goal_names)]  # std::map: ascending
    rest = sorted(n for n in py_names if n != "a" and n not in movable_of_goal)
    return ["a"] + movable_of_goal + rest


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class RefPuzzle:
    """One puzzle loaded by the reference's C++ `PushWorldPuzzle(filename)`; states are exchanged
    in the python reference's object order."""

    def __init__(self, text: str):
        self.py = OraclePuzzle(text)  # parsing only: names and the python order
        with tempfile.NamedTemporaryFile("w", suffix=".pwp", delete=False) as f:
            f.write(text)
            path = f.name
        try:
            self._h = lib().pwr_load(path.encode())
        finally:
            os.unlink(path)
        if not self._h:
            raise ValueError((lib().pwr_error() or b"").decode())
        self.n = lib().pwr_num_objects(self._h)
        self.g = lib().pwr_num_goals(self._h)
        assert self.n == self.py.num_objects and self.g == self.py.num_goals
        cpp_names = cpp_object_order(self.py.object_names, self.py.goal_names)
        # cpp_of_py[k] = index in the C++ state of python object k
        self.cpp_of_py = np.array([cpp_names.index(name) for name in self.py.object_names], dtype=np.int64)
        self.py_of_cpp = np.argsort(self.cpp_of_py)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().pwr_free(self._h)
            self._h = None

    def _to_cpp(self, state):
        xy = np.zeros(2 * self.n, np.int32)
        for k, (x, y) in enumerate(state):
            xy[2 * self.cpp_of_py[k]] = x
            xy[2 * self.cpp_of_py[k] + 1] = y
        return xy

    def _to_py(self, xy):
        return tuple((int(xy[2 * self.cpp_of_py[k]]), int(xy[2 * self.cpp_of_py[k] + 1])) for k in range(self.n))

    @property
    def initial_state(self):
        xy = np.zeros(2 * self.n, np.int32)
        lib().pwr_initial_state(self._h, _ptr(xy))
        return self._to_py(xy)

    @property
    def goal_state(self):
        """Goal positions indexed like the python reference's goal_state (goal k <-> object k+1)."""
        xy = np.zeros(2 * max(self.g, 1), np.int32)
        lib().pwr_goal(self._h, _ptr(xy))
        # C++ goal k belongs to C++ object k+1
        out = [None] * self.g
        for kc in range(self.g):
            k_py = int(self.py_of_cpp[kc + 1]) - 1
            out[k_py] = (int(xy[2 * kc]), int(xy[2 * kc + 1]))
        return tuple(out)

    def step(self, state, action):
        """getNextState -> (next state, moved-object mask), both in the python object order."""
        xy = self._to_cpp(state)
        moved_cpp = int(lib().pwr_step(self._h, _ptr(xy), int(action)))
        moved = 0
        for k in range(self.n):
            if (moved_cpp >> int(self.cpp_of_py[k])) & 1:
                moved |= 1 << k
        return self._to_py(xy), moved

    def satisfies_goal(self, state) -> bool:
        return bool(lib().pwr_satisfies_goal(self._h, _ptr(self._to_cpp(state))))

    def rollout(self, x, y, actions, reward=None, terminated=None, threads=0):
        """x, y: int32 [n, N] in the python object order (modified in place); actions uint8 [T, N].
        Every OpenMP thread steps its own copy of the C++ puzzle (getNextState is not re-entrant)."""
        n_envs = x.shape[1]
        cx = np.ascontiguousarray(x[self.py_of_cpp])
        cy = np.ascontiguousarray(y[self.py_of_cpp])
        assert actions.flags.c_contiguous
        ptr = lambda a: None if a is None else _ptr(a)
        done = lib().pwr_rollout(self._h, n_envs, actions.shape[0], _ptr(cx), _ptr(cy), n_envs,
                                 _ptr(actions), ptr(reward), ptr(terminated), threads)
        x[:] = cx[self.cpp_of_py]
        y[:] = cy[self.cpp_of_py]
        return done


class RefNovelty:
    """The reference's NoveltyHeuristic (novelty.cc:30-77).  Its value does not depend on how the
    objects are numbered (sets per object / per unordered object pair), so states are passed in the
    python order as they are."""

    def __init__(self, state_size: int):
        self.n = state_size
        self._h = lib().pwr_novelty_create(state_size)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().pwr_novelty_free(self._h)
            self._h = None

    def evaluate(self, state, moved_mask: int) -> int:
        xy = np.array([c for pos in state for c in pos], dtype=np.int32)
        return int(lib().pwr_novelty(self._h, _ptr(xy), self.n, int(moved_mask)))
