"""ctypes wrapper of oracle/cpu_krylov.c: the "best-effort CPU" baseline (OpenMP CSR mat-vec + fused modified
Gram-Schmidt on all cores).

THIS IS TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/oracle.py).  The factorisation runs in C; the
small dense phase (exp(tH) e1) reuses oracle.oracle's restatement of expv! (src/krylov_phiv.jl:200-247).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libcpukrylov.so")
_lib = None


def build(force: bool = False) -> str:
    """gcc -O3 -fopenmp (oracle/Makefile).  -march=native: the library is rebuilt on the box it runs on if the
    source is newer or the file is missing (bench.py calls this), never shipped across CPU generations blindly."""
    src = os.path.join(HERE, "cpu_krylov.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-s", "-B", "-C", HERE, "libcpukrylov.so"], check=True)
    return LIB


def load():
    global _lib
    if _lib is None:
        try:
            lib = C.CDLL(build())
        except OSError:  # built on another CPU generation (-march=native): rebuild here
            lib = C.CDLL(build(force=True))
        lib.cpuk_threads.restype = C.c_int
        lib.cpuk_set_threads.argtypes = [C.c_int]
        lib.cpuk_arnoldi.restype = C.c_int
        lib.cpuk_arnoldi.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                                     C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_double),
                                     C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.cpuk_project.argtypes = [C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        _lib = lib
    return _lib


def threads() -> int:
    return load().cpuk_threads()


def set_threads(nt: int):
    load().cpuk_set_threads(int(nt))


class Workspace:
    """Basis storage reused across calls (a user would hold a KrylovSubspace the same way)."""

    def __init__(self, n, m):
        self.n, self.m = n, m
        self.V = np.zeros((m + 1, n))           # row k = basis vector k (column-major n x (m+1))
        self.H = np.zeros((m, m + 1))           # row j = column j of the (m+1) x m Hessenberg matrix
        self.w = np.zeros(n)


def arnoldi(A, b, m=30, tol=1e-7, iop=0, ishermitian_=False, ws: Workspace | None = None):
    """Returns (V (n x (m_out+1) view), H ((m_out+1) x m_out), beta, m_out, breakdown)."""
    lib = load()
    A = A.tocsr()
    n = A.shape[0]
    rp = np.ascontiguousarray(A.indptr, dtype=np.int32)
    ci = np.ascontiguousarray(A.indices, dtype=np.int32)
    va = np.ascontiguousarray(A.data, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    ws = ws or Workspace(n, m)
    beta, mo, bd = C.c_double(), C.c_int(), C.c_int()
    lib.cpuk_arnoldi(n, rp.ctypes.data, ci.ctypes.data, va.ctypes.data, b.ctypes.data, m, tol, iop,
                     1 if ishermitian_ else 0, ws.V.ctypes.data, ws.H.ctypes.data, C.byref(beta), C.byref(mo),
                     C.byref(bd))
    return ws.V, ws.H.T, beta.value, mo.value, bool(bd.value)


def expv(t, A, b, m=30, tol=1e-7, iop=0, ishermitian_=False, ws: Workspace | None = None):
    """expv(t, A, b) with the factorisation and the projection in C/OpenMP."""
    from . import oracle as O
    lib = load()
    n = A.shape[0]
    ws = ws or Workspace(n, m)
    V, H, beta, mo, _ = arnoldi(A, b, m=m, tol=tol, iop=iop, ishermitian_=ishermitian_, ws=ws)
    if beta == 0.0:
        return np.zeros(n)
    y = np.ascontiguousarray(O.expv_small(t, np.array(H[:mo, :mo])))
    lib.cpuk_project(n, mo, V.ctypes.data, y.ctypes.data, beta, ws.w.ctypes.data)
    return ws.w.copy()
