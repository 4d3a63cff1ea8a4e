"""Builds ``pushworld_b200/lib/libpushworld_b200.so`` with nvcc for sm_100a (in-tree, so the
library travels with the repo snapshot; nvcc cross-compiles without a GPU)."""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libpushworld_b200.so")

NVCC_FLAGS = [
    "-std=c++17",
    "-O3",
    "-lineinfo",
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-Xcompiler",
    "-fPIC,-fvisibility=hidden",
]
# one object per translation unit, compiled in parallel and linked into one shared library; no
# relocatable device code: every kernel file is self-contained (device tables that several files
# need are uploaded once per file)
SOURCES = ["pw_kernels.cu", "pw_occ.cu", "pw_table.cu", "pw_render.cu", "pw_abi.cu", "pw_compile.cpp"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the pushworld_b200 CUDA library cannot be built")


STAMP_PATH = LIB_PATH + ".sha256"


def _sources_digest() -> str:
    """sha256 over the native sources and the public header (content, not timestamps: a checkout
    or a copy of the tree must not trigger a rebuild on every rank of a multi-GPU launch)."""
    import hashlib

    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    paths = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC))
    paths.append(os.path.join(os.path.dirname(PKG), "include", "pushworld_b200.h"))
    for path in paths:
        if os.path.isfile(path):
            h.update(os.path.basename(path).encode())
            with open(path, "rb") as f:
                h.update(f.read())
    return h.hexdigest()


def _stale() -> bool:
    if not os.path.exists(LIB_PATH) or not os.path.exists(STAMP_PATH):
        return True
    with open(STAMP_PATH) as f:
        return f.read().strip() != _sources_digest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compiles the CUDA library if it is missing or older than its sources.

    Safe to call from several processes at once (one rank per GPU under torchrun): an exclusive
    file lock serialises the builders, staleness is re-checked under the lock, and the compiler
    writes to a private temporary file that is renamed into place, so no process can ever load a
    partially written library."""
    if not force and not _stale():
        return LIB_PATH
    import fcntl

    os.makedirs(LIB_DIR, exist_ok=True)
    with open(os.path.join(LIB_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():  # another process built it while this one waited
                return LIB_PATH
            tmp = f"{LIB_PATH}.tmp.{os.getpid()}"
            obj_dir = os.path.join(LIB_DIR, f"obj.{os.getpid()}")
            os.makedirs(obj_dir, exist_ok=True)

            def compile_one(src):
                obj = os.path.join(obj_dir, os.path.splitext(src)[0] + ".o")
                cmd = [_nvcc(), *NVCC_FLAGS]
                if verbose:
                    cmd += ["-Xptxas", "-v"]
                cmd += ["-c", os.path.join(CSRC, src), "-o", obj]
                proc = subprocess.run(cmd, capture_output=True, text=True)
                return obj, cmd, proc

            try:
                from concurrent.futures import ThreadPoolExecutor

                with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
                    results = list(pool.map(compile_one, SOURCES))
                log = ""
                for obj, cmd, proc in results:
                    if proc.returncode != 0:
                        raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{proc.stdout}\n{proc.stderr}")
                    log += proc.stderr
                cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a",
                       *[obj for obj, _, _ in results], "-o", tmp]
                proc = subprocess.run(cmd, capture_output=True, text=True)
                if proc.returncode != 0:
                    raise RuntimeError(f"link failed:\n{' '.join(cmd)}\n{proc.stdout}\n{proc.stderr}")
                proc.stderr = log + proc.stderr
            finally:
                shutil.rmtree(obj_dir, ignore_errors=True)
                if os.path.exists(tmp) and os.path.getsize(tmp) == 0:
                    os.remove(tmp)
            os.replace(tmp, LIB_PATH)
            with open(STAMP_PATH + f".tmp.{os.getpid()}", "w") as f:
                f.write(_sources_digest() + "\n")
            os.replace(STAMP_PATH + f".tmp.{os.getpid()}", STAMP_PATH)
            if verbose:
                print(proc.stderr, file=sys.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
