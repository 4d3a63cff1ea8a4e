"""Builds and binds tests/hostcheck/hostcheck.cpp (g++), the CPU compile of the kernels'
transition rule.  Test infrastructure only."""

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        out_dir = os.path.join(HERE, "_build")
        os.makedirs(out_dir, exist_ok=True)
        so = os.path.join(out_dir, "libhostcheck.so")
        src = os.path.join(HERE, "hostcheck.cpp")
        csrc = os.path.join(REPO, "pushworld_b200", "csrc")
        deps = [src] + [os.path.join(csrc, f) for f in ("pw_core.cuh", "pw_render_core.cuh", "pw_board.cuh", "pw_fast.cuh")]
        if not os.path.exists(so) or os.path.getmtime(so) < max(map(os.path.getmtime, deps)):
            tmp = f"{so}.tmp.{os.getpid()}"  # renamed into place: safe with parallel test processes
            subprocess.check_call(
                [
                    "g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas",
                    "-x", "c++", "-I", os.path.join(REPO, "pushworld_b200", "csrc"),
                    src, "-o", tmp,
                ]
            )
            os.replace(tmp, so)
        _LIB = ctypes.CDLL(so)
        _LIB.hc_walk.restype = ctypes.c_int
    return _LIB


def walk(blob: bytes, state, actions):
    """Returns (states [T, n, 2], moved [T], achieved [T]) after each action."""
    buf = np.frombuffer(blob, dtype=np.uint64).copy()
    n = len(state)
    x = np.array([p[0] for p in state], dtype=np.int32)
    y = np.array([p[1] for p in state], dtype=np.int32)
    acts = np.asarray(actions, dtype=np.uint8)
    T = len(acts)
    xs = np.zeros((T, n), np.int32)
    ys = np.zeros((T, n), np.int32)
    moved = np.zeros(T, np.uint32)
    ach = np.zeros(T, np.int32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    got_n = lib().hc_walk(p(buf), p(x), p(y), p(acts), T, p(xs), p(ys), p(moved), p(ach))
    assert got_n == n
    return np.stack([xs, ys], axis=-1), moved, ach


def walk_tab(blob: bytes, state, actions):
    """`walk` through the collision bit tables (pw_fast.cuh); also returns the table size in bytes."""
    buf = np.frombuffer(blob, dtype=np.uint64).copy()
    n = len(state)
    x = np.array([p[0] for p in state], dtype=np.int32)
    y = np.array([p[1] for p in state], dtype=np.int32)
    acts = np.asarray(actions, dtype=np.uint8)
    T = len(acts)
    xs = np.zeros((T, n), np.int32)
    ys = np.zeros((T, n), np.int32)
    moved = np.zeros(T, np.uint32)
    ach = np.zeros(T, np.int32)
    size = ctypes.c_int(0)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    got_n = lib().hc_walk_tab(p(buf), p(x), p(y), p(acts), T, p(xs), p(ys), p(moved), p(ach), ctypes.byref(size))
    assert got_n == n
    return np.stack([xs, ys], axis=-1), moved, ach, size.value * 8


def board_walk(blob: bytes, state, actions, word_bits: int = 64):
    """Bitboard rule; returns None when the puzzle/state is not board-eligible."""
    buf = np.frombuffer(blob, dtype=np.uint64).copy()
    n = len(state)
    x = np.array([p[0] for p in state], dtype=np.int32)
    y = np.array([p[1] for p in state], dtype=np.int32)
    acts = np.asarray(actions, dtype=np.uint8)
    T = len(acts)
    xs = np.zeros((T, n), np.int32)
    ys = np.zeros((T, n), np.int32)
    moved = np.zeros(T, np.uint32)
    ach = np.zeros(T, np.int32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib().hc_board_walk(p(buf), word_bits, p(x), p(y), p(acts), T, p(xs), p(ys), p(moved), p(ach))
    if rc < 0:
        return None
    return np.stack([xs, ys], axis=-1), moved, ach


def render_u8(blob: bytes, state, ppc: int, bw: int):
    from pushworld_b200 import tables

    buf = np.frombuffer(blob, dtype=np.uint64).copy()
    words = buf.view(np.uint32)
    W, H = int(words[1] & 0xFFFF), int(words[1] >> 16)
    x = np.array([p[0] for p in state], dtype=np.int32)
    y = np.array([p[1] for p in state], dtype=np.int32)
    pal = np.zeros((16, 3), np.uint8)
    pal[: len(tables.PALETTE)] = np.array(tables.PALETTE, np.uint8)
    out = np.zeros((H * ppc, W * ppc, 3), np.uint8)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib().hc_render_u8(p(buf), p(x), p(y), ppc, bw, p(pal), p(out))
    return out
