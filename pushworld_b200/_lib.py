"""ctypes binding of ``libpushworld_b200.so`` (the C ABI in ``include/pushworld_b200.h``).

There is NO fallback: if the library is missing or a call fails this module raises.  The
library is built in-tree by ``pushworld_b200.build`` (nvcc, sm_100a).
"""

from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_int, c_int32, c_int64, c_size_t, c_uint32, c_uint64, c_void_p

from . import build as _build

PW_OK = 0
PW_E_INVALID, PW_E_CUDA, PW_E_FORMAT, PW_E_LIMIT, PW_E_NOMEM = -1, -2, -3, -4, -5
PW_DTYPE_U8, PW_DTYPE_F32 = 0, 1
PW_FLAG_AUTORESET = 1
PW_FLAG_RESAMPLE = 2
ABI_VERSION = 1

# every symbol include/pushworld_b200.h declares, with its argument types
SIGNATURES = {
    "pw_abi_version": (c_int, []),
    "pw_last_error": (c_char_p, []),
    "pw_compile_puzzle": (c_int, [c_char_p, c_size_t, c_void_p, c_size_t, POINTER(c_size_t)]),
    "pw_engine_create": (
        c_int,
        [c_void_p, c_size_t, c_void_p, c_int, c_int, POINTER(c_void_p)],
    ),
    "pw_engine_destroy": (c_int, [c_void_p]),
    "pw_engine_info": (
        c_int,
        [c_void_p, POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_int)],
    ),
    "pw_engine_poll": (c_int, [c_void_p, POINTER(c_uint64), c_void_p]),
    "pw_engine_clear_status": (c_int, [c_void_p, c_void_p]),
    "pw_reset": (
        c_int,
        [c_void_p, c_int64, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p],
    ),
    "pw_step": (
        c_int,
        [c_void_p, c_int64, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
         c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p],
    ),
    "pw_rollout": (
        c_int,
        [c_void_p, c_int64, c_int32, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_void_p,
         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_uint32,
         c_void_p],
    ),
    "pw_rollout_episodes": (
        c_int,
        [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
         c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_uint64, c_void_p,
         c_void_p],
    ),
    "pw_rollout_packed": (
        c_int,
        [c_void_p, c_int64, c_int32, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_void_p,
         c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_uint32, c_uint64, c_void_p, c_void_p],
    ),
    "pw_goals": (
        c_int,
        [c_void_p, c_int64, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p],
    ),
    "pw_expand": (
        c_int,
        [c_void_p, c_int64, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
         c_int64, c_void_p, c_void_p, c_void_p],
    ),
    "pw_expand_indexed": (
        c_int,
        [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
         c_int64, c_void_p, c_void_p, c_void_p],
    ),
    "pw_visited_insert": (
        c_int,
        [c_void_p, c_int, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p,
         c_int64, c_void_p, c_void_p],
    ),
    "pw_novelty_table_words": (c_int, [c_void_p, c_int, POINTER(c_uint64)]),
    "pw_novelty": (
        c_int,
        [c_void_p, c_int, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p,
         c_void_p, c_void_p, c_void_p],
    ),
    "pw_render": (
        c_int,
        [c_void_p, c_int64, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int,
         c_int, c_int, c_int, c_void_p],
    ),
    "pw_step_render": (
        c_int,
        [c_void_p, c_int64, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
         c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_int, c_int, c_int, c_int,
         c_int, c_void_p],
    ),
}

_lib = None


class PushWorldNativeError(RuntimeError):
    """The CUDA library is missing or a C-ABI call failed."""

    def __init__(self, code: int, message: str):
        super().__init__(f"pushworld_b200 native error {code}: {message}")
        self.code = code
        self.message = message


def library_path() -> str:
    # PUSHWORLD_B200_LIB points at an alternative build of the same ABI (kernel experiments)
    return os.environ.get("PUSHWORLD_B200_LIB") or _build.LIB_PATH


def load(build_if_missing: bool = True) -> ctypes.CDLL:
    """Loads (building first if the sources are newer) the native library."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if build_if_missing and path == _build.LIB_PATH:
        try:
            _build.build()
        except Exception as exc:  # no nvcc on this machine: a prebuilt .so may still be there
            if not os.path.exists(path):
                raise PushWorldNativeError(PW_E_INVALID, f"cannot build {path}: {exc}") from exc
    if not os.path.exists(path):
        raise PushWorldNativeError(
            PW_E_INVALID, f"{path} not found: run `python -m pushworld_b200.build` (needs nvcc)"
        )
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError = the .so does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.pw_abi_version() != ABI_VERSION:
        raise PushWorldNativeError(PW_E_INVALID, "ABI version mismatch; rebuild the library")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != PW_OK:
        msg = load().pw_last_error()
        raise PushWorldNativeError(rc, msg.decode("utf-8", "replace") if msg else "")


def compile_puzzle_text(text: str) -> bytes:
    """Native ``.pwp`` text -> table blob (``pw_compile_puzzle``); raises the same exception
    types as the reference parser (puzzle.py:141-157,230-232)."""
    lib = load()
    raw = text.encode("utf-8")
    size = c_size_t(0)
    rc = lib.pw_compile_puzzle(raw, len(raw), None, 0, ctypes.byref(size))
    if rc == PW_OK:
        buf = ctypes.create_string_buffer(size.value)
        rc = lib.pw_compile_puzzle(raw, len(raw), buf, size.value, ctypes.byref(size))
        if rc == PW_OK:
            return buf.raw[: size.value]
    msg = (lib.pw_last_error() or b"").decode("utf-8", "replace")
    if rc == PW_E_FORMAT and msg.startswith("Goal has no associated movable"):
        raise AssertionError(msg)
    if rc in (PW_E_FORMAT, PW_E_LIMIT):
        raise ValueError(msg)
    raise PushWorldNativeError(rc, msg)
