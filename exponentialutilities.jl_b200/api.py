"""Host-side mirror of the ExponentialUtilities.jl Krylov API over the C ABI.

Same names, argument meaning, defaults and error behaviour as the reference (Julia's ``f!`` is
spelled ``f_`` here):

    KrylovSubspace, arnoldi, arnoldi_, lanczos_      src/arnoldi.jl
    expv, expv_, phiv, phiv_                         src/krylov_phiv.jl
    kiops                                            src/kiops.jl
    exponential_, phiv_dense                         src/exp_baseexp.jl, src/phi.jl

PyTorch is used only for device memory and streams.  All arithmetic happens in
libb200krylov.so; if the library or a CUDA device is missing every call raises (no CPU path).
Vectors may be float64 CUDA tensors (results are CUDA tensors) or NumPy arrays (copied to the
device, results returned as NumPy arrays).
"""
from __future__ import annotations

import ctypes as C
import math
import weakref

import numpy as np

from . import _lib
from ._lib import ArgumentError, DimensionMismatch, KiopsOpts, KrylovOpts, TimestepOpts

try:  # torch is the device-memory / stream plumbing
    import torch
except Exception as _e:  # pragma: no cover
    torch = None
    _torch_err = _e


def _need_torch():
    if torch is None:
        raise RuntimeError(f"PyTorch is required for device memory: {_torch_err}")
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: the B200 Krylov engine has no CPU fallback")


def _round_up(x, a):
    return (x + a - 1) // a * a


# ------------------------------------------------------------------------------------------
# engine (handle)
# ------------------------------------------------------------------------------------------
class Engine:
    """One library handle on one device.  Calls run on torch's current stream of that device."""

    def __init__(self, device=None):
        _need_torch()
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else
                                   (device.index if isinstance(device, torch.device) else int(device)))
        torch.cuda.init()
        with torch.cuda.device(self.device):
            torch.zeros(1, device=self.device)  # make sure the primary context exists
            h = C.c_void_p()
            st = self.lib.b200k_create(C.byref(h), self.device.index,
                                       C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        _lib.check(st)
        self.handle = h
        self._finalizer = weakref.finalize(self, self.lib.b200k_destroy, h)

    def bind_stream(self):
        s = torch.cuda.current_stream(self.device).cuda_stream
        self.lib.b200k_set_stream(self.handle, C.c_void_p(s))

    def check(self, st):
        _lib.check(st, self.handle)

    def synchronize(self):
        self.check(self.lib.b200k_synchronize(self.handle))

    def device_info(self):
        sm, team, launches = C.c_int(), C.c_int(), C.c_int64()
        self.check(self.lib.b200k_device_info(self.handle, C.byref(sm), C.byref(team), C.byref(launches)))
        return {"sm_count": sm.value, "max_team": team.value, "launches": launches.value}

    def set_timing(self, enabled: bool):
        self.check(self.lib.b200k_set_timing(self.handle, 1 if enabled else 0))

    def set_flag(self, name: str, value: bool):
        """'force_ldg' or 'host_smallexp' (include/b200krylov.h: B200K_FLAG_*)."""
        flag = {"force_ldg": 1, "host_smallexp": 2, "l2hint": 3, "no_xl": 4, "no_mv": 5, "sym_pade": 6, "no_lz1": 7}[name]
        self.check(self.lib.b200k_set_flag(self.handle, flag, int(value)))

    def last_kernel(self):
        """'ldg' (krylov_persistent_kernel), 'tma' (krylov_tma_kernel), 'tma_xl' (its short-window instance), 'z'
        (complex LDG kernel) or 'tma_z' (complex kernel on the TMA ring) for the last factorisation."""
        w = C.c_int()
        self.check(self.lib.b200k_last_kernel(self.handle, C.byref(w)))
        return {1: "ldg", 2: "tma", 3: "z", 4: "tma_xl", 5: "tma_mv", 6: "tma_xl1", 7: "tma_z"}.get(w.value, "none")

    def last_timing(self):
        a, b = C.c_float(), C.c_float()
        self.check(self.lib.b200k_last_timing(self.handle, C.byref(a), C.byref(b)))
        return {"krylov_ms": a.value, "project_ms": b.value}


_engines = {}


def get_engine(device=None) -> Engine:
    _need_torch()
    idx = torch.cuda.current_device() if device is None else (
        device.index if isinstance(device, torch.device) else int(device))
    if idx is None:
        idx = torch.cuda.current_device()
    if idx not in _engines:
        _engines[idx] = Engine(idx)
    return _engines[idx]


# ------------------------------------------------------------------------------------------
# operators (docs/src/interfaces.md: size / eltype / mul! / ishermitian / opnorm)
# ------------------------------------------------------------------------------------------
class Operator:
    def __init__(self, engine: Engine, ptr, keep, is_complex=False, src=None):
        self.engine = engine
        self.ptr = ptr
        self._keep = keep
        self.is_complex = bool(is_complex)
        self._src = src          # host matrix the operator was ingested from (needed to promote it to complex)
        self._as_complex = None
        n, nnz, kind, herm, nrm = C.c_int64(), C.c_int64(), C.c_int(), C.c_int(), C.c_double()
        _lib.check(engine.lib.b200k_op_info(ptr, C.byref(n), C.byref(nnz), C.byref(kind), C.byref(herm),
                                            C.byref(nrm)))
        self.n, self.nnz, self.kind = n.value, nnz.value, kind.value
        self.ishermitian, self.opnorm_inf = bool(herm.value), nrm.value
        self.shape = (self.n, self.n)
        self.dtype = np.complex128 if self.is_complex else np.float64
        self._finalizer = weakref.finalize(self, engine.lib.b200k_op_destroy, ptr)

    def as_complex(self):
        """The same operator with ComplexF64 entries (promote_type(eltype(A), eltype(b)) when b is complex)."""
        if self.is_complex:
            return self
        if self._as_complex is None:
            if self._src is None:
                raise _lib.UnsupportedError("cannot promote a device-ingested operator to complex: pass a complex matrix")
            src = self._src
            self._as_complex = operator(src.astype(np.complex128), self.engine)
        return self._as_complex

    def mul(self, x):
        """mul!(y, A, x)"""
        if self.is_complex:
            raise _lib.UnsupportedError("mul! is not exposed for complex operators")
        xd, was_np = _to_device(x, self.engine)
        if xd.numel() != self.n:
            raise DimensionMismatch("length(x) != size(A, 2)")
        y = torch.empty_like(xd)
        self.engine.bind_stream()
        self.engine.check(self.engine.lib.b200k_op_apply(self.engine.handle, self.ptr, C.c_void_p(xd.data_ptr()),
                                                         C.c_void_p(y.data_ptr())))
        return _from_device(y, was_np)


def operator(A, engine: Engine | None = None) -> Operator:
    """Ingest a matrix: scipy.sparse (any format), 2-D NumPy array, 2-D float64 CUDA tensor, or a
    (rowptr, colind, val) triple of CUDA tensors / NumPy arrays describing 0-based CSR."""
    if isinstance(A, Operator):
        return A
    eng = engine or get_engine()
    lib = eng.lib
    eng.bind_stream()
    ptr = C.c_void_p()
    try:
        import scipy.sparse as sp
    except Exception:  # pragma: no cover
        sp = None
    if sp is not None and sp.issparse(A):
        if A.shape[0] != A.shape[1]:
            raise DimensionMismatch("operator must be square")
        A = A.tocsr()
        if not A.has_canonical_format:
            A = A.copy()
            A.sum_duplicates()
        rp = np.ascontiguousarray(A.indptr, dtype=np.int32)
        ci = np.ascontiguousarray(A.indices, dtype=np.int32)
        if np.iscomplexobj(A.data):
            va = np.ascontiguousarray(A.data, dtype=np.complex128)
            eng.check(lib.b200k_op_csr_create_z(eng.handle, A.shape[0], va.size, rp.ctypes.data, ci.ctypes.data,
                                                va.ctypes.data, 0, 1, C.byref(ptr)))
            return Operator(eng, ptr, None, is_complex=True, src=A)
        va = np.ascontiguousarray(A.data, dtype=np.float64)
        eng.check(lib.b200k_op_csr_create(eng.handle, A.shape[0], va.size, rp.ctypes.data, ci.ctypes.data,
                                          va.ctypes.data, 0, 1, C.byref(ptr)))
        return Operator(eng, ptr, None, src=A)
    if isinstance(A, tuple) and len(A) == 3:
        rp, ci, va = A
        if torch is not None and isinstance(va, torch.Tensor):
            rp = rp.to(torch.int32).contiguous()
            ci = ci.to(torch.int32).contiguous()
            va = va.to(torch.float64).contiguous()
            eng.check(lib.b200k_op_csr_create(eng.handle, rp.numel() - 1, va.numel(), C.c_void_p(rp.data_ptr()),
                                              C.c_void_p(ci.data_ptr()), C.c_void_p(va.data_ptr()), 0, 0,
                                              C.byref(ptr)))
            eng.synchronize()
            return Operator(eng, ptr, None)
        rp = np.ascontiguousarray(rp, dtype=np.int32)
        ci = np.ascontiguousarray(ci, dtype=np.int32)
        va = np.ascontiguousarray(va, dtype=np.float64)
        eng.check(lib.b200k_op_csr_create(eng.handle, rp.size - 1, va.size, rp.ctypes.data, ci.ctypes.data,
                                          va.ctypes.data, 0, 1, C.byref(ptr)))
        return Operator(eng, ptr, None)
    if torch is not None and isinstance(A, torch.Tensor) and A.is_complex():
        if A.dim() != 2 or A.shape[0] != A.shape[1]:
            raise DimensionMismatch("operator must be square")
        n = A.shape[0]
        At = A.to(device=eng.device, dtype=torch.complex128).t().contiguous()  # row i = column i of A
        eng.check(lib.b200k_op_dense_create_z(eng.handle, n, C.c_void_p(At.data_ptr()), n, 0, C.byref(ptr)))
        return Operator(eng, ptr, At, is_complex=True)
    if torch is not None and isinstance(A, torch.Tensor):
        if A.dim() != 2 or A.shape[0] != A.shape[1]:
            raise DimensionMismatch("operator must be square")
        n = A.shape[0]
        ld = _round_up(n, 2)
        At = torch.zeros((n, ld), dtype=torch.float64, device=eng.device)  # row i = column i of A
        At[:, :n] = A.to(device=eng.device, dtype=torch.float64).t()
        eng.check(lib.b200k_op_dense_create(eng.handle, n, C.c_void_p(At.data_ptr()), ld, 0, C.byref(ptr)))
        return Operator(eng, ptr, At)
    if np.iscomplexobj(A):
        A = np.asarray(A, dtype=np.complex128)
        if A.ndim != 2 or A.shape[0] != A.shape[1]:
            raise DimensionMismatch("operator must be square")
        Af = np.asfortranarray(A)
        eng.check(lib.b200k_op_dense_create_z(eng.handle, A.shape[0], Af.ctypes.data, A.shape[0], 1, C.byref(ptr)))
        return Operator(eng, ptr, None, is_complex=True, src=A)
    A = np.asarray(A, dtype=np.float64)
    if A.ndim != 2 or A.shape[0] != A.shape[1]:
        raise DimensionMismatch("operator must be square")
    Af = np.asfortranarray(A)
    eng.check(lib.b200k_op_dense_create(eng.handle, A.shape[0], Af.ctypes.data, A.shape[0], 1, C.byref(ptr)))
    return Operator(eng, ptr, None, src=A)


def _is_complex_value(x):
    if torch is not None and isinstance(x, torch.Tensor):
        return x.is_complex()
    return np.iscomplexobj(x)


def _to_device(x, eng: Engine):
    if torch is not None and isinstance(x, torch.Tensor):
        if x.dtype not in (torch.float64, torch.complex128):
            raise ArgumentError("only Float64 / ComplexF64 vectors are supported")
        if not x.is_cuda:
            return x.to(eng.device).contiguous(), False
        return x.contiguous(), False
    a = np.ascontiguousarray(np.asarray(x, dtype=np.complex128 if np.iscomplexobj(x) else np.float64))
    return torch.from_numpy(a).to(eng.device), True


def _from_device(x, was_np):
    return x.cpu().numpy() if was_np else x


# ------------------------------------------------------------------------------------------
# KrylovSubspace (src/arnoldi.jl:50-93)
# ------------------------------------------------------------------------------------------
class KrylovSubspace:
    """m, maxiter, augmented, beta, wasbreakdown, V, H -- V on the device, H on the host.

    V is (n + augmented) x (maxiter + 1) column-major; it is held as a (maxiter + 1, ldv) torch tensor
    (one basis vector per row, ldv padded to a multiple of 16) and exposed transposed through ``.V``.
    """

    def __init__(self, n, maxiter=30, augmented=0, engine: Engine | None = None, dtype=np.float64):
        self.engine = engine or get_engine()
        self.n = int(n)
        self.m = int(maxiter)
        self.maxiter = int(maxiter)
        self.augmented = int(augmented)
        self.beta = 0.0
        self.wasbreakdown = False
        self.is_complex = np.dtype(dtype) == np.complex128
        if self.is_complex and self.augmented:
            raise _lib.UnsupportedError("augmented Krylov subspaces are real only")
        self.nrows = self.n + self.augmented
        self.ldv = _round_up(self.nrows, 16)
        self._tdtype = torch.complex128 if self.is_complex else torch.float64
        self.Vt = torch.zeros((self.maxiter + 1, self.ldv), dtype=self._tdtype, device=self.engine.device)
        self.H = np.zeros((self.maxiter + 1, self.maxiter + (1 if self.augmented else 0)), order="F",
                          dtype=np.complex128 if self.is_complex else np.float64)

    @property
    def V(self):
        return self.Vt[:, : self.nrows].t()

    def getV(self):
        return self.Vt[: self.m + 1, : self.nrows].t()

    def getH(self):
        return self.H[: self.m + 1, : self.m + (1 if self.augmented else 0)]

    def resize(self, maxiter):
        """Base.resize! (src/arnoldi.jl:80-93): contents are preserved only for augmented subspaces."""
        isaug = self.augmented != 0
        Vt = torch.zeros((maxiter + 1, self.ldv), dtype=self._tdtype, device=self.engine.device)
        H = np.zeros((maxiter + 1, maxiter + (1 if isaug else 0)), order="F", dtype=self.H.dtype)
        if isaug:
            Vt[: self.Vt.shape[0]] = self.Vt
            H[: self.H.shape[0], : self.H.shape[1]] = self.H
        self.Vt, self.H = Vt, H
        self.m = self.maxiter = int(maxiter)
        return self


def _resolve_op(A, engine=None):
    if isinstance(A, tuple) and len(A) == 2:  # augmented (A, B)
        return operator(A[0], engine), A[1]
    return operator(A, engine), None


def arnoldi_(Ks: KrylovSubspace, A, b, *, tol=1.0e-7, m=None, ishermitian=None, opnorm=None, iop=0, init=0,
             t=float("nan"), mu=float("nan"), l=-1):
    """arnoldi!(Ks, A, b; tol, m, ishermitian, opnorm, iop, init, t, mu, l) -- src/arnoldi.jl:345-377.

    ``A`` may be ``(A, B)`` and ``b`` may be ``(w, w_aug)`` for the augmented form used by kiops
    (``w`` is n x numSteps, ``l`` the 1-based column)."""
    eng = Ks.engine
    op, B = _resolve_op(A, eng)
    if Ks.is_complex:
        return _arnoldi_z(Ks, op, b, tol=tol, m=m, ishermitian=ishermitian, iop=iop, init=init)
    if op.is_complex or (not isinstance(b, tuple) and _is_complex_value(b)):
        raise ArgumentError("complex operator or vector needs a complex KrylovSubspace (dtype=np.complex128)")
    if m is None:
        m = min(Ks.maxiter, op.n)
    herm = op.ishermitian if ishermitian is None else bool(ishermitian)
    Ks.wasbreakdown = False
    if m > Ks.maxiter:
        Ks.resize(m)
    else:
        Ks.m = m
    # checkdims (src/arnoldi.jl:207-220)
    if isinstance(b, tuple):
        bw, b_aug = b
        bd, _ = _to_device(bw, eng)
        if bd.dim() == 2:
            if bd.shape[1] != 1 and bd.numel() != op.n:
                raise DimensionMismatch("length(b') != size(A, 1)")
            bd = bd[:, l - 1].contiguous() if bd.shape[1] > 1 else bd.reshape(-1)
        p = int(np.asarray(b_aug).shape[0])
    else:
        bd, _ = _to_device(b, eng)
        bd = bd.reshape(-1)
        p = 0
    if not (bd.numel() == op.n == Ks.nrows - p):
        raise DimensionMismatch(f"length(b) [{bd.numel()}] == size(A,1) [{op.n}] == size(V,1)-p [{Ks.nrows - p}] "
                                "doesn't hold")
    opts = KrylovOpts()
    eng.lib.b200k_krylov_opts_default(C.byref(opts))
    opts.m, opts.tol, opts.iop, opts.hermitian, opts.init = int(m), float(tol), int(iop), int(herm), int(init)
    keep = None
    if p > 0:
        Bd, _ = _to_device(B, eng)
        ldb = _round_up(op.n, 2)
        keep = torch.zeros((p, ldb), dtype=torch.float64, device=eng.device)
        keep[:, : op.n] = Bd.reshape(op.n, p).t()
        opts.p, opts.B, opts.ldb, opts.t, opts.mu = p, keep.data_ptr(), ldb, float(t), float(mu)
        if init == 0:  # the reference also overwrites the caller's w_aug (src/arnoldi.jl:259-266)
            ba = np.asarray(b_aug)
            for k in range(1, p + 1):
                ba[k - 1] = mu if k == p else t ** (p - k) / math.factorial(p - k) * mu
    beta = C.c_double(Ks.beta)
    m_out, brk = C.c_int(), C.c_int()
    eng.bind_stream()
    st = eng.lib.b200k_arnoldi(eng.handle, op.ptr, C.c_void_p(bd.data_ptr()), C.byref(opts),
                               C.c_void_p(Ks.Vt.data_ptr()), Ks.ldv, Ks.maxiter,
                               Ks.H.ctypes.data_as(_lib.c_double_p), Ks.H.shape[0], C.byref(beta), C.byref(m_out),
                               C.byref(brk))
    eng.check(st)
    Ks.beta = beta.value
    Ks.m = m_out.value
    Ks.wasbreakdown = bool(brk.value)
    return Ks


def _arnoldi_z(Ks, op, b, *, tol, m, ishermitian, iop, init):
    """arnoldi!/lanczos! on a ComplexF64 basis (b200k_arnoldi_z)."""
    eng = Ks.engine
    op = op.as_complex()
    if m is None:
        m = min(Ks.maxiter, op.n)
    herm = op.ishermitian if ishermitian is None else bool(ishermitian)
    Ks.wasbreakdown = False
    if m > Ks.maxiter:
        Ks.resize(m)
    else:
        Ks.m = m
    bd, _ = _to_device(b, eng)
    bd = bd.reshape(-1).to(torch.complex128).contiguous()
    if not (bd.numel() == op.n == Ks.nrows):
        raise DimensionMismatch(f"length(b) [{bd.numel()}] == size(A,1) [{op.n}] == size(V,1) [{Ks.nrows}] doesn't hold")
    opts = KrylovOpts()
    eng.lib.b200k_krylov_opts_default(C.byref(opts))
    opts.m, opts.tol, opts.iop, opts.hermitian, opts.init = int(m), float(tol), int(iop), int(herm), int(init)
    beta = C.c_double(Ks.beta)
    m_out, brk = C.c_int(), C.c_int()
    eng.bind_stream()
    st = eng.lib.b200k_arnoldi_z(eng.handle, op.ptr, C.c_void_p(bd.data_ptr()), C.byref(opts),
                                 C.c_void_p(Ks.Vt.data_ptr()), Ks.ldv, Ks.maxiter, C.c_void_p(Ks.H.ctypes.data),
                                 Ks.H.shape[0], C.byref(beta), C.byref(m_out), C.byref(brk))
    eng.check(st)
    Ks.beta, Ks.m, Ks.wasbreakdown = beta.value, m_out.value, bool(brk.value)
    return Ks


def lanczos_(Ks: KrylovSubspace, A, b, **kw):
    """lanczos!(Ks, A, b; tol, m, init, t, mu, l) -- src/arnoldi.jl:456-490."""
    kw.pop("ishermitian", None)
    return arnoldi_(Ks, A, b, ishermitian=True, **kw)


def arnoldi(A, b, *, m=None, ishermitian=None, **kw):
    """arnoldi(A, b; m = min(30, size(A,1)), ishermitian = ishermitian(A), kw...) -- src/arnoldi.jl:161-180."""
    op = operator(A)
    n = int(np.prod(b.shape))
    if m is None:
        m = min(30, op.n)
    cplx = op.is_complex or _is_complex_value(b)  # T = promote_type(eltype(A), eltype(b))
    Ks = KrylovSubspace(n, m, 0, engine=op.engine, dtype=np.complex128 if cplx else np.float64)
    return arnoldi_(Ks, op, b, m=m, ishermitian=ishermitian, **kw)


# ------------------------------------------------------------------------------------------
# expv / phiv (src/krylov_phiv.jl)
# ------------------------------------------------------------------------------------------
class ExpvCache:
    """ExpvCache{T}(maxiter) -- src/krylov_phiv.jl:45-77.  Same state and growth rules as the reference: a flat ``mem``
    of maxiter^2 elements that ``get_cache(m)`` views as the m x m working copy of H (``resize`` doubles it on demand,
    :69-77), ``expcol`` for the first column of the reduced exponential, and the size-keyed store of exponential!
    workspaces (``expcache``: at most 64 entries, oldest evicted first, :447-470) -- here an entry records that the
    library's own workspace for that size is warm (the Pade scratch itself lives in the handle).  ``expv_`` runs the
    small dense phase IN this memory when a cache is passed."""

    def __init__(self, maxiter: int, dtype=np.float64):
        self.dtype = np.dtype(dtype)
        self.mem = np.empty(int(maxiter) ** 2, dtype=self.dtype)
        self.expcol = np.empty(int(maxiter), dtype=self.dtype)
        self.expcache = []  # [(n, workspace token)]

    @property
    def maxiter(self):
        return int(math.isqrt(self.mem.size))

    def resize(self, maxiter: int):
        """Base.resize!(C::ExpvCache, maxiter) -- krylov_phiv.jl:69-73 (note the factor 2)."""
        self.mem = np.empty(int(maxiter) ** 2 * 2, dtype=self.dtype)
        if self.expcol.size < maxiter:
            self.expcol = np.empty(int(maxiter), dtype=self.dtype)
        return self

    def get_cache(self, m: int):
        """get_cache(C, m) -- krylov_phiv.jl:74-77: the m x m (column-major) view, grown on demand."""
        if m * m > self.mem.size:
            self.resize(m)
        return self.mem[: m * m].reshape((m, m), order="F")

    def get_expcache(self, n: int):
        """get_expcache!(C, A, ExpMethodHigham2005Base()) -- krylov_phiv.jl:447-470: FIFO store bounded at 64 sizes."""
        for nc, work in self.expcache:
            if nc == n:
                return work
        if len(self.expcache) >= 64:
            self.expcache.pop(0)
        work = {"n": n, "uses": 0}
        self.expcache.append((n, work))
        return work


class PhivCache:
    """PhivCache(w, maxiter, p) -- src/krylov_phiv.jl:404-428, 471-504: one flat buffer of
    maxiter + maxiter^2 + (maxiter + p)^2 + maxiter (p + 1) elements that ``get_caches(m, p)`` splits into
    (e, Hcopy, C1, C2) and doubles on demand; ``coeffs`` and ``ts1`` scratch of phiv_timestep!.  ``phiv_`` evaluates
    phiv_dense! in this memory when a cache is passed."""

    def __init__(self, w, maxiter: int, p: int):
        dt = np.complex128 if (_is_complex_value(w) if w is not None else False) else np.float64
        self.dtype = np.dtype(dt)
        maxiter, p = int(maxiter), int(p)
        self.mem = np.empty(self._numelems(maxiter, p), dtype=self.dtype)
        self.expcache = []
        self.coeffs = np.ones(max(p, 1), dtype=self.dtype)
        self.ts1 = np.empty(1)
        self.useview = not (torch is not None and isinstance(w, torch.Tensor) and w.is_cuda)  # krylov_phiv.jl:420

    @staticmethod
    def _numelems(m, p):
        return m + m * m + (m + p) ** 2 + m * (p + 1)

    def resize(self, maxiter: int, p: int):
        self.mem = np.empty(self._numelems(int(maxiter), int(p)) * 2, dtype=self.dtype)
        return self

    def get_caches(self, m: int, p: int):
        """get_caches(C, m, p) -- krylov_phiv.jl:479-504."""
        if self._numelems(m, p) > self.mem.size:
            self.resize(m, p)
        e = self.mem[:m]
        off = m
        Hcopy = self.mem[off: off + m * m].reshape((m, m), order="F")
        off += m * m
        C1 = self.mem[off: off + (m + p) ** 2].reshape((m + p, m + p), order="F")
        off += (m + p) ** 2
        C2 = self.mem[off: off + m * (p + 1)].reshape((m, p + 1), order="F")
        return e, Hcopy, C1, C2

    get_expcache = ExpvCache.get_expcache


def _expv_with_cache(w, t, Ks, cache):
    """expv!(w, t, Ks; cache) for a real subspace and real t with the small dense phase in the cache's memory
    (krylov_phiv.jl:214-244): Hcopy = get_cache(cache, m) <- H[1:m, 1:m]; expcol <- exp(t Hcopy) e1; w = beta V expcol."""
    eng = Ks.engine
    m = Ks.m
    if Ks.beta == 0.0:
        w.zero_()
        return w
    Hcopy = cache.get_cache(m)
    Hcopy[:, :] = Ks.H[:m, :m]
    cache.get_expcache(m)["uses"] += 1
    if cache.expcol.size < m:
        cache.expcol = np.empty(m, dtype=cache.dtype)
    y = cache.expcol[:m]
    _lib.check(eng.lib.b200k_expv_small(m, Hcopy.ctypes.data_as(_lib.c_double_p), m, float(t),
                                        y.ctypes.data_as(_lib.c_double_p), None))
    eng.bind_stream()
    eng.check(eng.lib.b200k_project(eng.handle, C.c_void_p(Ks.Vt.data_ptr()), Ks.ldv, Ks.nrows, m, Ks.beta,
                                    y.ctypes.data_as(_lib.c_double_p), m, 1, C.c_void_p(w.data_ptr()), Ks.nrows))
    return w


def expv_(w, t, Ks: KrylovSubspace, *, cache=None):
    """expv!(w, t, Ks) -- src/krylov_phiv.jl:200-247.  ``w``: float64 CUDA tensor of length size(V,1)."""
    if cache is not None and not isinstance(cache, ExpvCache):
        raise ArgumentError("Cache must be an ExpvCache")  # krylov_phiv.jl:221
    eng = Ks.engine
    t_c = isinstance(t, (complex, np.complexfloating))
    if w.numel() != Ks.nrows:
        raise DimensionMismatch("Dimension mismatch")
    if Ks.is_complex:  # expv!(w::Complex, t, Ks{Complex}) -- krylov_phiv.jl:200-280
        if not w.is_complex():
            raise ArgumentError("a complex Krylov subspace needs a complex output vector")
        eng.bind_stream()
        st = eng.lib.b200k_expv_ks_z(eng.handle, float(np.real(t)), float(np.imag(t)), C.c_void_p(Ks.Vt.data_ptr()),
                                     Ks.ldv, Ks.nrows, C.c_void_p(Ks.H.ctypes.data), Ks.H.shape[0], Ks.m, Ks.beta,
                                     C.c_void_p(w.data_ptr()))
        eng.check(st)
        return w
    if t_c:  # real subspace, complex t: y is complex, w = beta V (re y) + i beta V (im y)
        if not w.is_complex():
            raise ArgumentError("complex t needs a complex output vector")
        m = Ks.m
        if Ks.beta == 0.0:
            w.zero_()
            return w
        Hc = np.array(Ks.H[:m, :m], dtype=np.complex128, order="F")
        y = np.zeros(m, dtype=np.complex128)
        _lib.check(eng.lib.b200k_expv_small_z(m, C.c_void_p(Hc.ctypes.data), m, float(np.real(t)), float(np.imag(t)),
                                              C.c_void_p(y.ctypes.data), None))
        Y = np.asfortranarray(np.stack([y.real, y.imag], 1))
        ld = _round_up(Ks.nrows, 2)
        Wt = torch.empty((2, ld), dtype=torch.float64, device=eng.device)
        eng.bind_stream()
        eng.check(eng.lib.b200k_project(eng.handle, C.c_void_p(Ks.Vt.data_ptr()), Ks.ldv, Ks.nrows, m, Ks.beta,
                                        Y.ctypes.data_as(_lib.c_double_p), m, 2, C.c_void_p(Wt.data_ptr()), ld))
        w.copy_(torch.complex(Wt[0, : Ks.nrows], Wt[1, : Ks.nrows]))
        return w
    if w.is_complex():  # real subspace and real t with a ComplexF64 output vector: promote the real result
        wr = torch.empty(Ks.nrows, dtype=torch.float64, device=eng.device)
        expv_(wr, t, Ks)
        w.copy_(wr.to(torch.complex128))
        return w
    if w.dtype != torch.float64:
        raise ArgumentError("expv! needs a Float64 (or ComplexF64) output vector")
    if cache is not None and cache.dtype == np.float64:
        return _expv_with_cache(w, t, Ks, cache)
    eng.bind_stream()
    st = eng.lib.b200k_expv_ks(eng.handle, float(t), C.c_void_p(Ks.Vt.data_ptr()), Ks.ldv, Ks.nrows,
                               Ks.H.ctypes.data_as(_lib.c_double_p), Ks.H.shape[0], Ks.m, Ks.beta,
                               C.c_void_p(w.data_ptr()))
    eng.check(st)
    return w


def expv(t, A, b=None, *, mode="happy_breakdown", m=None, tol=1.0e-7, ishermitian=None, iop=0, opnorm=None,
         cache=None, expmethod=None, rtol=None, return_m=False):
    """expv(t, A, b; m, tol, ishermitian, iop, ...) or expv(t, Ks) -- src/krylov_phiv.jl:125-168."""
    t_c = isinstance(t, (complex, np.complexfloating))
    if cache is not None and not isinstance(cache, ExpvCache):
        raise ArgumentError("Cache must be an ExpvCache")  # expv forwards `cache` to expv! (krylov_phiv.jl:135-144, 221)
    if isinstance(A, KrylovSubspace):
        Ks = A
        wdt = torch.complex128 if (Ks.is_complex or t_c) else torch.float64
        w = torch.empty(Ks.nrows, dtype=wdt, device=Ks.engine.device)
        return expv_(w, t, Ks, cache=cache)
    if mode not in ("happy_breakdown", "error_estimate"):
        raise ArgumentError(f"Unknown Krylov iteration termination mode, {mode}")
    op = operator(A)
    eng = op.engine
    bd, was_np = _to_device(b, eng)
    bd = bd.reshape(-1)
    if bd.numel() != op.n:
        raise DimensionMismatch("length(b) != size(A, 1)")
    if m is None:
        m = min(30, op.n)
    if op.is_complex or bd.is_complex():  # ComplexF64 basis
        if mode == "error_estimate":
            raise _lib.UnsupportedError("mode=:error_estimate is real-only in the B200 engine")
        opz = op.as_complex()
        bz = bd.to(torch.complex128).contiguous()
        opts = KrylovOpts()
        eng.lib.b200k_krylov_opts_default(C.byref(opts))
        opts.m, opts.tol, opts.iop = int(m), float(tol), int(iop)
        opts.hermitian = -1 if ishermitian is None else int(bool(ishermitian))
        w = torch.empty_like(bz)
        eng.bind_stream()
        eng.check(eng.lib.b200k_expv_z(eng.handle, opz.ptr, float(np.real(t)), float(np.imag(t)),
                                       C.c_void_p(bz.data_ptr()), C.byref(opts), C.c_void_p(w.data_ptr()), None, None))
        return _from_device(w, was_np)
    if t_c:  # real operator and vector, complex t: real Krylov subspace, complex projection coefficients
        Ks = arnoldi(op, bd, m=m, tol=tol, ishermitian=ishermitian, iop=iop)
        w = torch.empty(Ks.nrows, dtype=torch.complex128, device=eng.device)
        return _from_device(expv_(w, t, Ks), was_np)
    if mode == "error_estimate":  # _expv_ee (src/krylov_phiv.jl:145-160): atol = tol, rtol = sqrt(tol)
        herm = op.ishermitian if ishermitian is None else bool(ishermitian)
        if not herm:
            raise _lib.UnsupportedError("Error estimation not yet available for non-Hermitian matrices.")
        w = torch.empty_like(bd)
        mo = C.c_int()
        eng.bind_stream()
        st = eng.lib.b200k_expv_ee(eng.handle, op.ptr, float(t), C.c_void_p(bd.data_ptr()), int(m), float(tol),
                                   float(math.sqrt(tol) if rtol is None else rtol), C.c_void_p(w.data_ptr()),
                                   C.byref(mo))
        eng.check(st)
        out = _from_device(w, was_np)
        return (out, mo.value) if return_m else out
    opts = KrylovOpts()
    eng.lib.b200k_krylov_opts_default(C.byref(opts))
    opts.m, opts.tol, opts.iop = int(m), float(tol), int(iop)
    opts.hermitian = -1 if ishermitian is None else int(bool(ishermitian))
    w = torch.empty_like(bd)
    eng.bind_stream()
    st = eng.lib.b200k_expv(eng.handle, op.ptr, float(t), C.c_void_p(bd.data_ptr()), C.byref(opts),
                            C.c_void_p(w.data_ptr()), None, None, None)
    eng.check(st)
    if was_np:  # host result: complete the call here, which also surfaces a deferred SingularException of the device-side
        eng.synchronize()  # small exponential (CUDA-tensor callers stay asynchronous and get it at their next synchronize())
    return _from_device(w, was_np)


def expv_host(t, op: Operator, b_host, w_host, *, m=30, tol=1.0e-7, ishermitian=None, iop=0):
    """End-to-end call with HOST (ideally pinned) float64 tensors: H2D(b) -> expv -> D2H(w), synchronous."""
    eng = op.engine
    opts = KrylovOpts()
    eng.lib.b200k_krylov_opts_default(C.byref(opts))
    opts.m, opts.tol, opts.iop = int(min(m, op.n)), float(tol), int(iop)
    opts.hermitian = -1 if ishermitian is None else int(bool(ishermitian))
    eng.bind_stream()
    st = eng.lib.b200k_expv_host(eng.handle, op.ptr, float(t), C.c_void_p(b_host.data_ptr()), C.byref(opts),
                                 C.c_void_p(w_host.data_ptr()), None, None)
    eng.check(st)
    return w_host


def expv_host_async(t, op: Operator, b_host, w_host, *, m=30, tol=1.0e-7, ishermitian=None, iop=0, engine=None):
    """expv_host without the final synchronisation, on ``engine`` (default: the operator's) and torch's current
    stream: H2D(b), the kernels and D2H(w) are queued and the call returns; ``engine.synchronize()`` completes it.
    Two engines on two streams keep two requests in flight so that copies overlap kernels."""
    eng = engine or op.engine
    opts = KrylovOpts()
    eng.lib.b200k_krylov_opts_default(C.byref(opts))
    opts.m, opts.tol, opts.iop = int(min(m, op.n)), float(tol), int(iop)
    opts.hermitian = -1 if ishermitian is None else int(bool(ishermitian))
    eng.bind_stream()
    eng.check(eng.lib.b200k_expv_host_async(eng.handle, op.ptr, float(t), C.c_void_p(b_host.data_ptr()), C.byref(opts),
                                            C.c_void_p(w_host.data_ptr())))
    return w_host


def phiv_(w, t, Ks: KrylovSubspace, k, *, cache=None, correct=False, errest=False):
    """phiv!(w, t, Ks, k; correct, errest) -- src/krylov_phiv.jl:607-653.

    ``w``: float64 CUDA tensor holding the column-major nrows x (k+1) result, i.e. of shape (k+1, nrows)."""
    if cache is not None and not isinstance(cache, PhivCache):
        raise ArgumentError("Cache must be a PhivCache")  # krylov_phiv.jl:630
    eng = Ks.engine
    if w.dim() != 2 or w.shape[1] != Ks.nrows or w.shape[0] != k + 1:
        raise DimensionMismatch("Dimension mismatch")
    if Ks.is_complex or isinstance(t, (complex, np.complexfloating)):
        return _phiv_z(w, t, Ks, k, correct=correct, errest=errest)
    if w.dtype != torch.float64 or not w.is_cuda or w.stride(1) != 1:
        raise ArgumentError("phiv! on a real Krylov subspace needs a float64 CUDA output of shape (k+1, nrows)")
    if cache is not None and cache.dtype == np.float64 and Ks.beta != 0.0:
        return _phiv_with_cache(w, t, Ks, k, cache, correct, errest)
    err = C.c_double()
    eng.bind_stream()
    st = eng.lib.b200k_phiv_ks(eng.handle, float(t), C.c_void_p(Ks.Vt.data_ptr()), Ks.ldv, Ks.nrows,
                               Ks.H.ctypes.data_as(_lib.c_double_p), Ks.H.shape[0], Ks.m, Ks.beta, int(k),
                               1 if correct else 0, C.c_void_p(w.data_ptr()), w.stride(0), C.byref(err))
    eng.check(st)
    return (w, err.value) if errest else w


def _phiv_with_cache(w, t, Ks, k, cache, correct, errest):
    """_phiv! with the small dense phase in the PhivCache's memory (krylov_phiv.jl:632-652): (e, Hcopy, C1, C2) =
    get_caches(cache, m, k); Hcopy <- t H[1:m, :]; C2 <- phiv_dense!(Hcopy, e, k); w = beta V C2 (+ correction)."""
    eng = Ks.engine
    m = Ks.m
    e, Hcopy, C1, C2 = cache.get_caches(m, k)
    Hcopy[:, :] = float(t) * Ks.H[:m, :m]
    e[:] = 0.0
    e[0] = 1.0
    cache.get_expcache(m + k)["uses"] += 1
    _lib.check(eng.lib.b200k_phiv_dense(m, Hcopy.ctypes.data_as(_lib.c_double_p), m, e.ctypes.data_as(_lib.c_double_p),
                                        int(k), C2.ctypes.data_as(_lib.c_double_p), m))
    hlast = float(Ks.H[m, m - 1])  # H[end, end] of the (m+1) x m view
    err = abs(Ks.beta * hlast * float(t) * C2[m - 1, k])
    mm = m + 1 if correct else m
    Y = np.zeros((mm, k + 1), order="F")
    Y[:m, :] = C2
    if correct:  # w[:, i] += beta h t C2[end, i+1] v_{m+1}: one more basis column with that coefficient / beta
        Y[m, :k] = hlast * float(t) * C2[m - 1, 1:]
    eng.bind_stream()
    eng.check(eng.lib.b200k_project(eng.handle, C.c_void_p(Ks.Vt.data_ptr()), Ks.ldv, Ks.nrows, mm, Ks.beta,
                                    Y.ctypes.data_as(_lib.c_double_p), mm, k + 1, C.c_void_p(w.data_ptr()), w.stride(0)))
    return (w, err) if errest else w


def _phiv_z(w, t, Ks, k, *, correct, errest):
    """_phiv! on a ComplexF64 subspace (b200k_phiv_ks_z); a real subspace with complex t is not offered."""
    eng = Ks.engine
    if not Ks.is_complex:
        raise _lib.UnsupportedError("phiv! with complex t on a real Krylov subspace: build the subspace with a complex b")
    if not w.is_complex() or not w.is_cuda or w.stride(1) != 1:
        raise ArgumentError("a complex Krylov subspace needs a ComplexF64 CUDA output of shape (k+1, nrows)")
    err = C.c_double()
    eng.bind_stream()
    st = eng.lib.b200k_phiv_ks_z(eng.handle, float(np.real(t)), float(np.imag(t)), C.c_void_p(Ks.Vt.data_ptr()), Ks.ldv,
                                 Ks.nrows, C.c_void_p(Ks.H.ctypes.data), Ks.H.shape[0], Ks.m, Ks.beta, int(k),
                                 1 if correct else 0, C.c_void_p(w.data_ptr()), w.stride(0), C.byref(err))
    eng.check(st)
    return (w, err.value) if errest else w


def phiv(t, A, b=None, k=None, *, cache=None, correct=False, errest=False, m=None, tol=1.0e-7, ishermitian=None,
         iop=0, opnorm=None):
    """phiv(t, A, b, k; ...) or phiv(t, Ks, k; ...) -- src/krylov_phiv.jl:563-575.

    Returns the n x (k+1) matrix [phi_0(tA) b ... phi_k(tA) b] (NumPy for NumPy input, otherwise a
    transposed view of a (k+1, n) CUDA tensor)."""
    if isinstance(A, KrylovSubspace):
        Ks, k = A, b
        was_np = False
    else:
        op = operator(A)
        was_np = not (torch is not None and isinstance(b, torch.Tensor))
        Ks = arnoldi(op, b, m=m, tol=tol, ishermitian=ishermitian, iop=iop)
    cplx = Ks.is_complex or isinstance(t, (complex, np.complexfloating))
    w = torch.empty((k + 1, Ks.nrows), dtype=torch.complex128 if cplx else torch.float64, device=Ks.engine.device)
    res = phiv_(w, t, Ks, k, cache=cache, correct=correct, errest=errest)
    wt = (res[0] if errest else res).t()
    out = wt.cpu().numpy() if was_np else wt
    return (out, res[1]) if errest else out


def expv_batched(ts, A, B, *, m=30, tol=1.0e-7, ishermitian=None, iop=0, out=None):
    """nb independent expv(t_i, A, B[:, i]) on a shared operator in one launch.  B: n x nb.
    ``out``: optional (nb, round_up(n, 2)) float64 CUDA tensor that receives the results (row i = w_i)."""
    op = operator(A)
    eng = op.engine
    ts = np.ascontiguousarray(np.asarray(ts, dtype=np.float64).reshape(-1))
    if (torch is not None and isinstance(B, torch.Tensor) and B.is_cuda and B.dtype == torch.float64 and B.dim() == 2
            and B.stride(0) == 1):
        Bd, was_np = B, False  # column-major view: keep the layout (no .contiguous())
    else:
        Bd, was_np = _to_device(B, eng)
    if Bd.dim() != 2 or Bd.shape[0] != op.n or Bd.shape[1] != ts.size:
        raise DimensionMismatch("B must be n x nb with nb == length(ts)")
    nb = ts.size
    if (not was_np and Bd.stride(0) == 1 and (nb == 1 or (Bd.stride(1) >= op.n and Bd.stride(1) % 2 == 0))
            and Bd.data_ptr() % 16 == 0):
        # already column-major with an even leading dimension (Julia's layout, e.g. `Bt.t()` of an (nb, ld) tensor):
        # handed to the library as it is
        Bt, ld = Bd, (Bd.stride(1) if nb > 1 else _round_up(op.n, 2))
    else:
        ld = _round_up(op.n, 2)
        Bt = torch.zeros((nb, ld), dtype=torch.float64, device=eng.device)
        Bt[:, : op.n] = Bd.t()
    ldw = _round_up(op.n, 2)
    Wt = torch.empty((nb, ldw), dtype=torch.float64, device=eng.device) if out is None else out
    if Wt.shape != (nb, ldw) or Wt.dtype != torch.float64 or not Wt.is_contiguous():
        raise DimensionMismatch("out must be a contiguous float64 (nb, round_up(n, 2)) CUDA tensor")
    opts = KrylovOpts()
    eng.lib.b200k_krylov_opts_default(C.byref(opts))
    opts.m, opts.tol, opts.iop = int(min(m, op.n)), float(tol), int(iop)
    opts.hermitian = -1 if ishermitian is None else int(bool(ishermitian))
    eng.bind_stream()
    st = eng.lib.b200k_expv_batched(eng.handle, op.ptr, nb, ts.ctypes.data_as(_lib.c_double_p),
                                    C.c_void_p(Bt.data_ptr()), ld, C.byref(opts), C.c_void_p(Wt.data_ptr()), ldw,
                                    None, None)
    eng.check(st)
    if was_np:
        eng.synchronize()  # (surfaces a deferred SingularException of the device-side small exponentials)
    W = Wt[:, : op.n].t()
    return W.cpu().numpy() if was_np else W


# ------------------------------------------------------------------------------------------
# kiops (src/kiops.jl:57-281)
# ------------------------------------------------------------------------------------------
def kiops(tau_out, A, u, *, mmin=10, mmax=128, m=None, tol=1.0e-7, opnorm=None, iop=2, ishermitian=None,
          task1=False, _normU=None, return_device=False):
    """kiops(tau_out, A, u; mmin, mmax, m, tol, opnorm, iop, ishermitian, task1) -> (w, stats).

    ``w`` is n x numSteps (host NumPy array, as the reference returns a host ``zeros(n, numSteps)``,
    src/kiops.jl:89) and ``stats = (steps, rejected, krylov_steps, exps, m)``."""
    op = operator(A)
    eng = op.engine
    tau_arr = np.asarray(tau_out, dtype=np.float64)
    tau_is_row = 1 if (tau_arr.ndim == 2 and tau_arr.shape[0] == 1) else 0
    numSteps = tau_arr.shape[1] if tau_arr.ndim == 2 else 1
    tau_flat = np.ascontiguousarray(tau_arr.reshape(-1))
    ud, _ = _to_device(u, eng)
    if ud.dim() == 1:
        ud = ud.reshape(-1, 1)
    n, ppo = ud.shape
    if n != op.n:
        raise DimensionMismatch("size(u, 1) != size(A, 1)")
    ld = _round_up(n, 2)
    Ut = torch.zeros((ppo, ld), dtype=torch.float64, device=eng.device)
    Ut[:, :n] = ud.t()
    Wt = torch.zeros((numSteps, ld), dtype=torch.float64, device=eng.device)
    ko = KiopsOpts()
    eng.lib.b200k_kiops_opts_default(C.byref(ko))
    ko.mmin, ko.mmax, ko.tol, ko.iop = int(mmin), int(mmax), float(tol), int(iop)
    ko.m = int(min(mmin, mmax) if m is None else m)
    ko.hermitian = -1 if ishermitian is None else int(bool(ishermitian))
    ko.task1 = 1 if task1 else 0
    if _normU is not None:  # row-sharded callers supply the global norm(u[:, 2:end], 1)
        ko.normU = float(_normU)
    stats = (C.c_int64 * 5)()
    eng.bind_stream()
    st = eng.lib.b200k_kiops(eng.handle, op.ptr, tau_flat.size, tau_flat.ctypes.data_as(_lib.c_double_p),
                             tau_is_row, C.c_void_p(Ut.data_ptr()), ld, ppo, C.byref(ko),
                             C.c_void_p(Wt.data_ptr()), ld, stats)
    eng.check(st)
    w = Wt[:, :n].t()
    if not return_device:  # the reference returns a host matrix (src/kiops.jl:89)
        w = w.cpu().numpy()
    return w, tuple(int(s) for s in stats)


# ------------------------------------------------------------------------------------------
# phiv_timestep / expv_timestep (src/krylov_phiv_adaptive.jl)
# ------------------------------------------------------------------------------------------
def phiv_timestep(ts, A, B, *, tau=0.0, m=None, tol=1.0e-7, opnorm=None, iop=0, correct=False, adaptive=False,
                  delta=1.2, ishermitian=None, gamma=0.8, NA=0, return_steps=False):
    """phiv_timestep(ts, A, B; tau, m, tol, opnorm, iop, correct, adaptive, delta, ishermitian, gamma, NA)
    -- src/krylov_phiv_adaptive.jl:116-453.  B is n x (p+1) (a vector gives expv_timestep); ``ts`` a scalar or a
    list of output times; ``opnorm`` None (Arnoldi estimate), a scalar, or a callable (A, inf) -> bound."""
    op = operator(A)
    eng = op.engine
    if op.is_complex or _is_complex_value(B):
        nco = 1 if (not hasattr(B, "shape") or len(B.shape) == 1) else int(B.shape[1])
        if nco != 1:
            raise _lib.UnsupportedError("ComplexF64 phiv_timestep with p > 0 is not implemented (expv_timestep is)")
        return _expv_timestep_z(ts, op, B, tau=tau, m=m, tol=tol, opnorm=opnorm, iop=iop, correct=correct,
                                adaptive=adaptive, delta=delta, ishermitian=ishermitian, gamma=gamma, NA=NA,
                                return_steps=return_steps)
    scalar_t = np.isscalar(ts)
    tsa = np.ascontiguousarray(np.atleast_1d(np.asarray(ts, dtype=np.float64)))
    Bd, was_np = _to_device(B, eng)
    if Bd.dim() == 1:
        Bd = Bd.reshape(-1, 1)
    n, ncoef = Bd.shape
    if n != op.n:
        raise DimensionMismatch("Dimension mismatch")
    ld = _round_up(n, 2)
    Bt = torch.zeros((ncoef, ld), dtype=torch.float64, device=eng.device)
    Bt[:, :n] = Bd.t()
    Ut = torch.zeros((tsa.size, ld), dtype=torch.float64, device=eng.device)
    to = TimestepOpts()
    eng.lib.b200k_timestep_opts_default(C.byref(to))
    to.tau, to.tol, to.iop = float(tau), float(tol), int(iop)
    to.m = int(min(10, op.n) if m is None else m)
    if opnorm is not None:
        to.opnorm = float(opnorm(A, np.inf)) if callable(opnorm) else float(opnorm)
    to.correct, to.adaptive = int(bool(correct)), int(bool(adaptive))
    to.delta, to.gamma, to.NA = float(delta), float(gamma), int(NA)
    to.hermitian = -1 if ishermitian is None else int(bool(ishermitian))
    nsteps = C.c_int()
    eng.bind_stream()
    st = eng.lib.b200k_phiv_timestep(eng.handle, op.ptr, tsa.size, tsa.ctypes.data_as(_lib.c_double_p),
                                     C.c_void_p(Bt.data_ptr()), ld, ncoef, C.byref(to), C.c_void_p(Ut.data_ptr()), ld,
                                     C.byref(nsteps))
    eng.check(st)
    eng.synchronize()
    Uo = Ut[:, :n].t()
    out = Uo[:, 0] if scalar_t else Uo
    out = out.cpu().numpy() if was_np else out
    return (out, nsteps.value) if return_steps else out


def _ts_flops(m, tau, n, p, NA, iop, Hnorm, maxtau):
    """_phiv_timestep_estimate_flops -- src/krylov_phiv_adaptive.jl:482-501."""
    iop = m if iop == 0 else iop
    flops = 2 * (p - 1) * (NA + n) + (2 * p + 1) * n + 2 * m * NA + sum(3 * min(i, iop) for i in range(1, m + 1))
    MH = 44 / 3 + 2 * math.ceil(max(0.0, math.log2(Hnorm / 5.37)))
    return (flops + round(MH * (m + p) ** 3)) * int(math.ceil(maxtau / tau))


def _ts_adapt(m, tau, eps, m_old, tau_old, eps_old, q, kappa, gamma, omega, maxtau, n, p, NA, iop, Hnorm):
    """_phiv_timestep_adapt (Niesen-Wright, Algorithm 4) -- src/krylov_phiv_adaptive.jl:455-481."""
    if tau_old > tau:
        q = math.log(tau / tau_old) / math.log(eps / eps_old) - 1
    tau_new = min(max(tau * (gamma / omega) ** (1 / (q + 1)), tau / 5), 2 * tau, maxtau)
    if m_old < m:
        kappa = (eps / eps_old) ** (1 / (m_old - m))
    m_new = m + math.ceil(math.log(omega / gamma) / math.log(kappa))
    m_new = min(max(m_new, (3 * m) // 4, 1), int(math.ceil(4 * m / 3)))
    if _ts_flops(m, tau_new, n, p, NA, iop, Hnorm, maxtau) < _ts_flops(m_new, tau, n, p, NA, iop, Hnorm, maxtau):
        m_new = m
    else:
        tau_new = tau
    return m_new, tau_new, q, kappa


def _expv_timestep_z(ts, op, b, *, tau, m, tol, opnorm, iop, correct, adaptive, delta, ishermitian, gamma, NA,
                     return_steps):
    """expv_timestep for ComplexF64 operators / vectors (the reference's own GPU test, test/gpu/gputests.jl:41-58).
    The Niesen-Wright controller of phiv_timestep! (src/krylov_phiv_adaptive.jl:260-453) with p = 0, on the host as in
    the reference; every length-n operation runs in the library (b200k_arnoldi_z, b200k_phiv_ks_z)."""
    eng = op.engine
    opz = op.as_complex()
    scalar_t = np.isscalar(ts)
    tsa = np.sort(np.atleast_1d(np.asarray(ts, dtype=np.float64)))
    bd, was_np = _to_device(b, eng)
    bd = bd.reshape(-1).to(torch.complex128).contiguous()
    n = opz.n
    if bd.numel() != n:
        raise DimensionMismatch("Dimension mismatch")
    m = int(min(10, n) if m is None else m)
    herm = opz.ishermitian if ishermitian is None else bool(ishermitian)
    abstol = opn = None
    if opnorm is not None:
        opn = float(opnorm(op, np.inf)) if callable(opnorm) else float(opnorm)
        abstol = tol * opn
    b0norm = float(bd.abs().max().item())

    def tau_guess(mm):
        return 10 / opn * (abstol * ((mm + 1) / math.e) ** (mm + 1) * math.sqrt(2 * math.pi * (mm + 1)) /
                           (4 * opn * b0norm)) ** (1 / mm)

    if opn is not None and tau == 0:
        tau = tau_guess(m)
    tend = float(tsa[-1])
    seed = opn is None and tau == 0
    if seed:
        tau = tend
    if adaptive:
        if herm:
            iop = 2
        if NA == 0:
            NA = opz.nnz
    Ut = torch.zeros((tsa.size, n), dtype=torch.complex128, device=eng.device)
    u = bd.clone()
    Ks = KrylovSubspace(n, m, engine=eng, dtype=np.complex128)
    w = torch.empty((2, n), dtype=torch.complex128, device=eng.device)
    ws = torch.empty((2, n), dtype=torch.complex128, device=eng.device)
    t, snapshot, nsteps = 0.0, 1, 0
    while t < tend:
        if t + tau > tend:
            tau = tend - t
        arnoldi_(Ks, opz, u, tol=tol, m=m, iop=iop)
        if abstol is None:
            opn = float(np.linalg.norm(Ks.getH(), 1))
            abstol = tol * opn
            if seed:
                tau = min(tend - t, gamma * tau_guess(m))
        if Ks.wasbreakdown:
            tau = tend - t
        _, eps = phiv_(w, tau, Ks, 1, correct=correct, errest=True)
        if adaptive:
            omega = (tend / tau) * (eps / abstol)
            eps_old, m_old, tau_old = eps, m, tau
            q, kappa = m / 4, 2.0
            maxtau = tend - t
            guard = 0
            while omega > delta and guard < 200:
                guard += 1
                m_new, tau_new, q, kappa = _ts_adapt(m, tau, eps, m_old, tau_old, eps_old, q, kappa, gamma, omega, maxtau,
                                                     n, 0, NA, iop, float(np.linalg.norm(Ks.getH(), 1)))
                m, m_old = m_new, m
                tau, tau_old = tau_new, tau
                arnoldi_(Ks, opz, u, tol=tol, m=m, iop=iop)
                _, eps_new = phiv_(w, tau, Ks, 1, correct=correct, errest=True)
                eps, eps_old = eps_new, eps
                omega = (tend / tau) * (eps / abstol)
        while snapshot <= tsa.size and t + tau >= tsa[snapshot - 1]:
            phiv_(ws, float(tsa[snapshot - 1] - t), Ks, 1, correct=correct)
            Ut[snapshot - 1].copy_(ws[0])
            snapshot += 1
        u = w[0].clone()  # u = tau^0 * P[:, end-1]: the phi_0 column
        t += tau
        nsteps += 1
    out = Ut[0] if scalar_t else Ut.t()
    out = out.cpu().numpy() if was_np else out
    return (out, nsteps) if return_steps else out


def expv_timestep(ts, A, b, **kw):
    """expv_timestep(ts, A, b; ...) -- src/krylov_phiv_adaptive.jl:57-114 (phiv_timestep with p = 0)."""
    return phiv_timestep(ts, A, b, **kw)


# ------------------------------------------------------------------------------------------
# small dense (host) -- src/exp_baseexp.jl, src/phi.jl
# ------------------------------------------------------------------------------------------
def exponential_(A):
    """exponential!(A, ExpMethodHigham2005Base()) on a host matrix (returns a new F-ordered array)."""
    lib = _lib.load()
    Af = np.array(A, dtype=np.float64, order="F", copy=True)
    if Af.ndim != 2 or Af.shape[0] != Af.shape[1]:
        raise DimensionMismatch("matrix is not square")
    _lib.check(lib.b200k_exponential(Af.shape[0], Af.ctypes.data_as(_lib.c_double_p), max(Af.shape[0], 1)))
    return Af


def exponential_batched_(A, engine: Engine | None = None):
    """exponential! of a batch of small matrices on the device, in place.  ``A``: float64 CUDA tensor of
    shape (nbatch, n, n) holding each matrix in COLUMN-major order (i.e. A[b] is the transpose of the matrix as
    torch prints it), n <= 48; NumPy input of shape (nbatch, n, n) (ordinary row-major matrices) is converted,
    processed on the device and returned as NumPy."""
    eng = engine or get_engine()
    was_np = not (torch is not None and isinstance(A, torch.Tensor))
    if was_np:
        An = np.asarray(A, dtype=np.float64)
        if An.ndim != 3 or An.shape[1] != An.shape[2]:
            raise DimensionMismatch("expected (nbatch, n, n)")
        At = torch.from_numpy(np.ascontiguousarray(An.transpose(0, 2, 1))).to(eng.device)
    else:
        At = A
        if At.dim() != 3 or At.shape[1] != At.shape[2] or At.dtype != torch.float64 or not At.is_contiguous():
            raise DimensionMismatch("expected a contiguous float64 (nbatch, n, n) tensor")
    nb, n, _ = At.shape
    eng.bind_stream()
    eng.check(eng.lib.b200k_exponential_batched(eng.handle, int(nb), int(n), C.c_void_p(At.data_ptr()), int(n),
                                                int(n) * int(n)))
    return At.cpu().numpy().transpose(0, 2, 1).copy() if was_np else At


def phiv_dense(A, v, k):
    """phiv_dense(A, v, k) -- src/phi.jl:63-66, 84-115 (host)."""
    lib = _lib.load()
    Af = np.array(A, dtype=np.float64, order="F", copy=True)
    v = np.ascontiguousarray(v, dtype=np.float64)
    m = v.shape[0]
    if Af.shape != (m, m):
        raise DimensionMismatch("Dimension mismatch")
    w = np.zeros((m, k + 1), order="F")
    _lib.check(lib.b200k_phiv_dense(m, Af.ctypes.data_as(_lib.c_double_p), m, v.ctypes.data_as(_lib.c_double_p),
                                    int(k), w.ctypes.data_as(_lib.c_double_p), m))
    return w


def expv_small(H, t):
    """y = exp(t*H) e1 exactly as the small dense phase of expv! does it (host); returns (y, branch)
    with branch 1 = SymTridiagonal eigen, 0 = Higham-2005 Pade (src/krylov_phiv.jl:223-244)."""
    lib = _lib.load()
    Hf = np.array(H, dtype=np.float64, order="F", copy=True)
    m = Hf.shape[0]
    y = np.zeros(m)
    br = C.c_int()
    _lib.check(lib.b200k_expv_small(m, Hf.ctypes.data_as(_lib.c_double_p), m, float(t),
                                    y.ctypes.data_as(_lib.c_double_p), C.byref(br)))
    return y, br.value
