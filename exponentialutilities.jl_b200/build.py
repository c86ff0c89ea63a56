"""In-tree build of libb200krylov.so (sm_100a only) with nvcc.

The shared library is the product; it is built next to this file so that it travels with the
repository snapshot to the GPU box (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb200krylov.so")
SOURCES = ["b200krylov.cu"]
DEPS = ["b200krylov.cu", "krylov_kernel.cuh", "aux_kernels.cuh", "ptx.cuh", "smallmat.hpp",
        os.path.join("..", "..", "include", "b200krylov.h")]

NVCC_FLAGS = [
    "-shared", "-Xcompiler", "-fPIC", "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libb200krylov.so")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the library if it is missing or older than its sources; returns its path."""
    if not force and not is_stale():
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
