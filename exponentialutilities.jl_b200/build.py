"""In-tree build of libb200krylov.so (sm_100a only) with nvcc.

The shared library is the product; it is built next to this file so that it travels with the
repository snapshot to the GPU box (it is git-ignored, not gpurun-ignored).  Staleness is decided by a
content hash of the sources stored beside the library (file mtimes do not survive the snapshot copy),
and concurrent builders (several ranks importing at once) are serialised with a file lock; the
library is replaced atomically.
"""
from __future__ import annotations

import fcntl
import hashlib
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb200krylov.so")
HASHFILE = LIB + ".srchash"
SOURCES = ["b200krylov.cu"]


def _deps():
    """Every source/header under csrc/ plus the public header (hashing all of them means a new header can
    never be forgotten)."""
    files = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".hpp", ".h")))
    return files + [os.path.join("..", "..", "include", "b200krylov.h")]


NVCC_FLAGS = [
    "-shared", "-Xcompiler", "-fPIC", "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libb200krylov.so")


def source_hash() -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for d in _deps():
        with open(os.path.join(CSRC, d), "rb") as f:
            h.update(d.encode())
            h.update(f.read())
    return h.hexdigest()


def is_stale() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(HASHFILE):
        return True
    try:
        return open(HASHFILE).read().strip() != source_hash()
    except OSError:
        return True


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the library if it is missing or its sources changed; returns its path."""
    if not force and not is_stale():
        return LIB
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():  # another process built it while we waited
                return LIB
            tmp = LIB + f".tmp{os.getpid()}"
            cmd = [_nvcc(), *NVCC_FLAGS, "-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
            os.replace(tmp, LIB)
            with open(HASHFILE, "w") as f:
                f.write(source_hash())
            if verbose:
                print(res.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
