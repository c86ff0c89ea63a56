"""CPU oracle: a NumPy/SciPy restatement of the ExponentialUtilities.jl Krylov hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  The product (``exponentialutilities.jl_b200``) never does.

What it restates (all paths relative to /root/reference, v1.35.0 @ 6474ddc):

* ``KrylovSubspace``            src/arnoldi.jl:50-93
* ``arnoldi`` / ``arnoldi!``    src/arnoldi.jl:161-180, 345-377   (sequential *modified* Gram-Schmidt,
                                IOP window, absolute breakdown test ``beta < tol``, ``init`` continuation)
* ``applyA!`` (plain/augmented) src/arnoldi.jl:183-205
* ``firststep!`` (both)         src/arnoldi.jl:230-279
* ``lanczos!``                  src/arnoldi.jl:388-490             (loop 1:m ignores ``init``; mirror copy :488)
* ``expv`` / ``expv!``          src/krylov_phiv.jl:125-247         (exactly-symmetric H -> SymTridiagonal eigen,
                                else Higham-2005 Pade on t*H)
* ``phiv`` / ``_phiv!``         src/krylov_phiv.jl:563-653
* ``phiv_dense!``               src/phi.jl:84-115                  (Sidje augmented matrix)
* ``exponential!(Higham2005Base)`` src/exp_baseexp.jl:65-161       (thresholds, coefficient tuples, the generic
                                even/odd power loop, s = ceil(log2(nA/5.4)), balance/unbalance)
* ``kiops``                     src/kiops.jl:57-326                (incl. the integer-division quirks at :210,
                                ``numSteps = size(tau_out, 2)`` and ``krystep == 0``)

Third-party arithmetic that is NOT under /root/reference and is restated through LAPACK/BLAS here:
PureGebal (compat "1") balance!/unbalance!  -> LAPACK dgebal job 'B' (scipy.linalg.matrix_balance);
LinearSolve (compat "5") LU solve           -> LAPACK dgesv (numpy.linalg.solve);
LinearAlgebra eigen!(SymTridiagonal)        -> LAPACK dstemr (scipy.linalg.eigh_tridiagonal);
LinearAlgebra dot/axpy!/norm                -> OpenBLAS ddot/daxpy/dnrm2 (scipy.linalg.blas);
LinearAlgebra.exp (kiops.jl:156,307)        -> the same Higham-2005 restatement.
All of these act on <= 129x129 matrices or are exactly specified BLAS-1 operations.

PARITY PIN.  Julia is not installed in the build container and the reference ships no golden
vectors for this path, so the oracle cannot be pinned to reference
*outputs*.  It is pinned instead to every analytic assertion the reference's own tests make for
this path (tests/test_oracle_reference_tests.py restates test/basictests.jl:515-574, 650-664,
731-754, 786-816, 952-974) and to scipy.linalg.expm / expm_multiply on converged cases.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp
from scipy.linalg import blas as _blas

__all__ = [
    "KrylovSubspace", "arnoldi", "arnoldi_", "lanczos_", "expv", "expv_ks", "phiv", "phiv_ks",
    "phiv_dense", "exponential_higham2005base", "kiops", "ishermitian",
]


# --------------------------------------------------------------------------------------
# operator helpers (the reference's operator interface: size / eltype / mul! / ishermitian)
# --------------------------------------------------------------------------------------
def ishermitian(A) -> bool:
    """LinearAlgebra.ishermitian(A) for dense ndarrays and scipy sparse matrices."""
    if sp.issparse(A):
        return (A != A.conj().T).nnz == 0
    A = np.asarray(A)
    return A.shape[0] == A.shape[1] and np.array_equal(A, A.conj().T)


def _mul(A, x):
    """mul!(y, A, x) -- src/arnoldi.jl:185."""
    if hasattr(A, "matvec") and not sp.issparse(A) and not isinstance(A, np.ndarray):
        return np.asarray(A.matvec(x))
    return np.asarray(A @ x).reshape(-1)


def _dot(x, y):
    if np.iscomplexobj(x) or np.iscomplexobj(y):
        return np.vdot(x, y)
    return float(_blas.ddot(x, y))


def _axpy(a, x, y):
    """axpy!(a, x, y): y += a*x in place (OpenBLAS daxpy for real contiguous columns)."""
    if (not np.iscomplexobj(x) and not np.iscomplexobj(y) and not np.iscomplexobj(a)
            and x.dtype == np.float64 and y.dtype == np.float64):
        out = _blas.daxpy(x, y, a=a)
        if out is not y and not np.shares_memory(out, y):
            y[...] = out
    else:
        y += a * x


def _nrm2(x):
    if np.iscomplexobj(x):
        return float(_blas.dznrm2(x))
    return float(_blas.dnrm2(x))


# --------------------------------------------------------------------------------------
# KrylovSubspace  (src/arnoldi.jl:50-93)
# --------------------------------------------------------------------------------------
class KrylovSubspace:
    """State of a Krylov factorisation: m, maxiter, augmented, beta, wasbreakdown, V, H.

    V is (n + augmented) x (maxiter + 1) column-major; H is (maxiter+1) x (maxiter + [aug != 0]).
    """

    def __init__(self, n, maxiter=30, augmented=0, dtype=np.float64, hdtype=None):
        self.m = maxiter
        self.maxiter = maxiter
        self.augmented = int(augmented)
        self.beta = 0.0
        self.wasbreakdown = False
        self.V = np.zeros((n + self.augmented, maxiter + 1), dtype=dtype, order="F")
        self.H = np.zeros((maxiter + 1, maxiter + (1 if augmented else 0)),
                          dtype=hdtype or dtype, order="F")

    def getV(self):
        return self.V[:, : self.m + 1]

    def getH(self):
        return self.H[: self.m + 1, : self.m + (1 if self.augmented else 0)]

    def resize(self, maxiter):
        """Base.resize! -- src/arnoldi.jl:80-93 (contents preserved only when augmented)."""
        isaug = self.augmented != 0
        V = np.zeros((self.V.shape[0], maxiter + 1), dtype=self.V.dtype, order="F")
        H = np.zeros((maxiter + 1, maxiter + (1 if isaug else 0)), dtype=self.H.dtype, order="F")
        if isaug:
            V[:, : self.V.shape[1]] = self.V
            H[: self.H.shape[0], : self.H.shape[1]] = self.H
        self.V, self.H = V, H
        self.m = self.maxiter = maxiter
        return self


# --------------------------------------------------------------------------------------
# Arnoldi / Lanczos  (src/arnoldi.jl)
# --------------------------------------------------------------------------------------
def _applyA(A, V, j, n, p):
    """applyA! -- src/arnoldi.jl:183-205; j is 0-based: V[:, j+1] = A_aug V[:, j]."""
    if isinstance(A, tuple):
        A0, B = A
        V[:n, j + 1] = _mul(A0, V[:n, j])
        V[:n, j + 1] += B @ V[n:n + p, j]
        V[n:n + p - 1, j + 1] = V[n + 1:n + p, j]
        V[n + p - 1, j + 1] = 0.0
    else:
        V[:, j + 1] = _mul(A, V[:, j])


def _checkdims(A, b, V):
    """src/arnoldi.jl:207-220."""
    if isinstance(b, tuple):
        bp, b_aug = b
        n, p = bp.shape[0], b_aug.shape[0]
        A0 = A[0]
    else:
        n, p = V.shape[0], 0
        A0, bp, b_aug = A, b, None
    if not (bp.shape[0] == A0.shape[0] == A0.shape[1] == V.shape[0] - p):
        raise ValueError("DimensionMismatch")
    return bp, b_aug, n, p


def _firststep(Ks, V, H, b):
    """src/arnoldi.jl:230-250."""
    H[...] = 0
    Ks.beta = _nrm2(np.ascontiguousarray(b))
    if Ks.beta != 0:
        V[:, 0] = b * (1.0 / Ks.beta)


def _firststep_aug(Ks, V, H, b, b_aug, t, mu, l):
    """src/arnoldi.jl:257-279 (l is 1-based column of b as in the reference)."""
    n, p = b.shape[0], b_aug.shape[0]
    for k in range(1, p + 1):
        if k == p:
            b_aug[k - 1] = mu
        else:
            i = p - k
            b_aug[k - 1] = t ** i / math.factorial(i) * mu
    H[...] = 0
    bl = b[:, l - 1]
    Ks.beta = beta = math.sqrt(_dot(bl, bl) + _dot(b_aug, b_aug))
    if beta != 0:
        V[:n, 0] = bl / beta
        V[n:n + p, 0] = b_aug / beta


def _coeff(alpha, hreal):
    return alpha.real if hreal and np.iscomplexobj(alpha) else alpha


def _arnoldi_step(j, iop, A, V, H, n, p):
    """arnoldi_step! -- src/arnoldi.jl:289-308.  j is 1-based as in the reference."""
    _applyA(A, V, j - 1, n, p)
    y = V[:, j]
    hreal = not np.iscomplexobj(H)
    for i in range(max(1, j - iop + 1), j + 1):
        alpha = _coeff(_dot(V[:, i - 1], y), hreal)
        H[i - 1, j - 1] = alpha
        _axpy(-alpha, V[:, i - 1], y)
    beta = _nrm2(y)
    H[j, j - 1] = beta
    with np.errstate(all="ignore"):
        y /= beta
    return beta


def arnoldi_(Ks, A, b, *, tol=1e-7, m=None, ishermitian_=None, opnorm=None, iop=0, init=0,
             t=float("nan"), mu=float("nan"), l=-1):
    """arnoldi! -- src/arnoldi.jl:345-377."""
    A0 = A[0] if isinstance(A, tuple) else A
    if m is None:
        m = min(Ks.maxiter, A0.shape[0])
    if ishermitian_ is None:
        ishermitian_ = ishermitian(A0)
    Ks.wasbreakdown = False
    if ishermitian_:
        return lanczos_(Ks, A, b, tol=tol, m=m, init=init, t=t, mu=mu, l=l)
    if m > Ks.maxiter:
        Ks.resize(m)
    else:
        Ks.m = m
    V, H = Ks.getV(), Ks.getH()
    bp, b_aug, n, p = _checkdims(A, b, V)
    if init == 0:
        if isinstance(A, tuple):
            _firststep_aug(Ks, V, H, bp, b_aug, t, mu, l)
        else:
            _firststep(Ks, V, H, b)
        init = 1
    if Ks.beta == 0:
        return Ks
    if iop == 0:
        iop = m
    for j in range(init, m + 1):
        beta = _arnoldi_step(j, iop, A, V, H, n, p)
        if beta < tol:
            Ks.m = j
            Ks.wasbreakdown = True
            break
    return Ks


def _lanczos_step(j, A, V, H, n, p):
    """lanczos_step! -- src/arnoldi.jl:388-403 (u = diag(H), v = subdiag(H)); j is 1-based."""
    _applyA(A, V, j - 1, n, p)
    x, y = V[:, j - 1], V[:, j]
    alpha = _dot(x, y)
    alpha = alpha.real if np.iscomplexobj(alpha) else alpha
    H[j - 1, j - 1] = alpha
    _axpy(-alpha, x, y)
    if j > 1:
        _axpy(-H[j - 1, j - 2], V[:, j - 2], y)
    beta = _nrm2(y)
    H[j, j - 1] = beta
    with np.errstate(all="ignore"):
        y /= beta
    return beta


def lanczos_(Ks, A, b, *, tol=1e-7, m=None, opnorm=None, init=0, t=float("nan"),
             mu=float("nan"), l=-1):
    """lanczos! -- src/arnoldi.jl:456-490."""
    A0 = A[0] if isinstance(A, tuple) else A
    if m is None:
        m = min(Ks.maxiter, A0.shape[0])
    Ks.wasbreakdown = False
    if m > Ks.maxiter:
        Ks.resize(m)
    else:
        Ks.m = m
    V, H = Ks.getV(), Ks.getH()
    bp, b_aug, n, p = _checkdims(A, b, V)
    if init == 0:
        if isinstance(A, tuple):
            _firststep_aug(Ks, V, H, bp, b_aug, t, mu, l)
        else:
            _firststep(Ks, V, H, b)
        init = 1
    if Ks.beta == 0:
        return Ks
    for j in range(1, m + 1):  # NB: ignores init (src/arnoldi.jl:480)
        if tol > _lanczos_step(j, A, V, H, n, p):
            Ks.m = j
            Ks.wasbreakdown = True
            break
    # copyto!(@diagview(H, 1), v[1:end-1])  (src/arnoldi.jl:488); v = subdiagonal of the H view
    nsub = min(H.shape[0] - 1, H.shape[1])
    for i in range(nsub - 1):
        H[i, i + 1] = H[i + 1, i]
    return Ks


def arnoldi(A, b, *, m=None, ishermitian_=None, **kw):
    """arnoldi(A, b; m, ishermitian, kw...) -- src/arnoldi.jl:161-180."""
    n = b.shape[0]
    if m is None:
        m = min(30, A.shape[0])
    if ishermitian_ is None:
        ishermitian_ = ishermitian(A)
    T = np.result_type(A.dtype, b.dtype)
    U = np.empty(0, T).real.dtype if ishermitian_ else T
    Ks = KrylovSubspace(n, m, 0, dtype=T, hdtype=U)
    return arnoldi_(Ks, A, b, m=m, ishermitian_=ishermitian_, **kw)


# --------------------------------------------------------------------------------------
# small dense exponential  (src/exp_baseexp.jl)
# --------------------------------------------------------------------------------------
_PADE_C3 = (120.0, 60.0, 12.0, 1.0)
_PADE_C5 = (30240.0, 15120.0, 3360.0, 420.0, 30.0, 1.0)
_PADE_C7 = (17297280.0, 8648640.0, 1995840.0, 277200.0, 25200.0, 1512.0, 56.0, 1.0)
_PADE_C9 = (17643225600.0, 8821612800.0, 2075673600.0, 302702400.0, 30270240.0, 2162160.0,
            110880.0, 3960.0, 90.0, 1.0)
_PADE_C13 = (64764752532480000.0, 32382376266240000.0, 7771770303897600.0, 1187353796428800.0,
             129060195264000.0, 10559470521600.0, 670442572800.0, 33522128640.0, 1323241920.0,
             40840800.0, 960960.0, 16380.0, 182.0, 1.0)


def _pade_evaluate(A, C):
    """_pade_evaluate! -- src/exp_baseexp.jl:84-105 (generic even/odd power loop)."""
    n = A.shape[0]
    N = len(C)
    A2 = A @ A
    P = np.eye(n, dtype=A.dtype)
    U = C[1] * P
    V = C[0] * P
    for k in range(1, N // 2):
        k2 = 2 * k
        P = P @ A2
        U = U + C[k2 + 1] * P
        V = V + C[k2] * P
    U = A @ U
    X = V + U
    temp = V - U
    try:
        return np.linalg.solve(temp, X)  # _pade_linsolve! :44-59
    except np.linalg.LinAlgError as e:  # SingularException(0)
        raise np.linalg.LinAlgError("SingularException(0)") from e


def exponential_higham2005base(A):
    """exponential!(A, ExpMethodHigham2005Base()) -- src/exp_baseexp.jl:112-161."""
    A = np.array(A, dtype=np.result_type(A.dtype, np.float64), order="F", copy=True)
    n = A.shape[0]
    if n == 0:
        return A
    # PureGebal.balance!  (gebal 'B': permute + power-of-two scaling)
    Ab, T = sla.matrix_balance(A, permute=True, scale=True, separate=False)
    nA = np.linalg.norm(Ab, 1)
    if nA <= 2.1:
        if nA > 0.95:
            X = _pade_evaluate(Ab, _PADE_C9)
        elif nA > 0.25:
            X = _pade_evaluate(Ab, _PADE_C7)
        elif nA > 0.015:
            X = _pade_evaluate(Ab, _PADE_C5)
        else:
            X = _pade_evaluate(Ab, _PADE_C3)
    else:
        s = math.log2(nA / 5.4)
        si = 0
        if s > 0:
            si = math.ceil(s)
            Ab = Ab / (2.0 ** si)
        X = _pade_evaluate(Ab, _PADE_C13)
        if s > 0:
            for _ in range(si):
                X = X @ X
    # PureGebal.unbalance!: X <- T X T^{-1}; T is a permuted power-of-two diagonal, so exact.
    Tinv = T.T.copy()
    nz = Tinv != 0
    Tinv[nz] = 1.0 / Tinv[nz]
    return T @ X @ Tinv


# --------------------------------------------------------------------------------------
# expv / phiv  (src/krylov_phiv.jl, src/phi.jl)
# --------------------------------------------------------------------------------------
def expv_small(t, Hcopy):
    """The small dense phase of expv! (src/krylov_phiv.jl:223-244): exp(t*H) e1 for the m x m block Hcopy."""
    if np.array_equal(Hcopy, Hcopy.conj().T):
        lam, Z = sla.eigh_tridiagonal(np.real(np.diag(Hcopy)).copy(),
                                      np.real(np.diag(Hcopy, -1)).copy(), lapack_driver="stemr")
        return Z @ (np.exp(t * lam) * Z[0, :])
    return exponential_higham2005base(t * Hcopy)[:, 0]


def expv_ks(t, Ks):
    """expv!(w, t, Ks) -- src/krylov_phiv.jl:200-280 (real and complex t)."""
    m, beta, V, H = Ks.m, Ks.beta, Ks.getV(), Ks.getH()
    n = V.shape[0]
    wtype = np.result_type(V.dtype, np.asarray(t).dtype)
    if beta == 0:
        return np.zeros(n, dtype=wtype)
    Hcopy = np.array(H[:m, :m], copy=True)
    Vm = V[:, :m]
    expHe = expv_small(t, Hcopy)
    return beta * (Vm @ expHe)


def expv(t, A, b, *, m=None, **kw):
    """expv(t, A, b; m, tol, ishermitian, iop, ...) happy-breakdown mode -- src/krylov_phiv.jl:125-144."""
    Ks = arnoldi(A, b, m=m, **kw)
    return expv_ks(t, Ks)


def phiv_dense(A, v, k):
    """phiv_dense! -- src/phi.jl:84-115."""
    m = v.shape[0]
    T = np.result_type(A.dtype, v.dtype)
    C = np.zeros((m + k, m + k), dtype=T)
    C[:m, :m] = A
    C[:m, m] = v
    for i in range(m + 1, m + k):  # 1-based i in (m+1):(m+k-1): cache[i, i+1] = 1
        C[i - 1, i] = 1
    P = exponential_higham2005base(C)
    w = np.zeros((m, k + 1), dtype=T)
    w[:, 0] = P[:m, :m] @ v
    for i in range(1, k + 1):
        w[:, i] = P[:m, m + i - 1]
    return w


def phiv_ks(t, Ks, k, *, correct=False, errest=False):
    """_phiv! -- src/krylov_phiv.jl:620-653."""
    m, beta, V, H = Ks.m, Ks.beta, Ks.getV(), Ks.getH()
    Hcopy = t * np.array(H[:m, :m], copy=True)
    e = np.zeros(m, dtype=Hcopy.dtype)
    e[0] = 1
    C2 = phiv_dense(Hcopy, e, k)
    w = beta * (V[:, :m] @ C2)
    if correct:
        betah = beta * H[-1, -1] * t
        vlast = V[:, -1]
        for i in range(1, k + 1):
            w[:, i - 1] += betah * C2[-1, i] * vlast
    err = abs(beta * H[-1, -1] * t * C2[-1, -1])
    return (w, err) if errest else w


def phiv(t, A, b, k, *, correct=False, errest=False, **kw):
    """phiv(t, A, b, k; ...) -- src/krylov_phiv.jl:563-570."""
    Ks = arnoldi(A, b, **kw)
    return phiv_ks(t, Ks, k, correct=correct, errest=errest)


# --------------------------------------------------------------------------------------
# kiops  (src/kiops.jl)
# --------------------------------------------------------------------------------------
def _opnorm_inf(A):
    if sp.issparse(A):
        return abs(A).sum(axis=1).max()
    return np.linalg.norm(A, np.inf)


def _cld(a, b):
    return -((-a) // b)


def kiops(tau_out, A, u, *, mmin=10, mmax=128, m=None, tol=1e-7, opnorm=None, iop=2,
          ishermitian_=None, task1=False):
    """kiops -- src/kiops.jl:57-281.  Returns (w [n x numSteps], stats 5-tuple).

    ``tau_out`` may be a scalar or a 1-D/2-D array; as in the reference ``numSteps = size(tau_out, 2)``
    (a scalar or a column vector therefore gives one output column).
    """
    tau_arr = np.atleast_1d(np.asarray(tau_out, dtype=float))
    numSteps = tau_arr.shape[1] if tau_arr.ndim == 2 else 1
    tau_flat = tau_arr.reshape(-1)
    u = np.asarray(u, dtype=float)
    if u.ndim == 1:
        u = u[:, None]
    if m is None:
        m = min(mmin, mmax)
    if ishermitian_ is None:
        ishermitian_ = ishermitian(A)
    n, ppo = u.shape
    p = ppo - 1
    if p == 0:
        p = 1
        u = np.hstack([u, np.zeros_like(u)])
    Ks = KrylovSubspace(n, m, p)
    step = krystep = ireject = reject = exps = 0
    sgn = np.sign(tau_flat[-1])
    tau_now = 0.0
    tau_end = abs(tau_flat[-1])
    j = 0
    w = np.zeros((n, numSteps), order="F")
    w_aug = np.zeros(p)
    w[:, 0] = u[:, 0]
    normU = np.abs(u[:, 1:]).sum()  # norm(view, 1) on a matrix view = vector 1-norm of entries
    if ppo > 1 and normU > 0:
        ex = math.ceil(math.log2(normU))
        nu, mu = 2.0 ** (-ex), 2.0 ** ex
    else:
        nu, mu = 1.0, 1.0
    u_flip = nu * u[:, :0:-1]
    tau = tau_end
    if tau_end > 1:
        gamma, gamma_mmax = 0.2, 0.1
    else:
        gamma, gamma_mmax = 0.9, 0.6
    delta = 1.4
    oldm, oldtau, omega = -1, float("nan"), float("nan")
    orderold = kestold = True
    order, kest = 0.0, 2
    l = 1
    while tau_now < tau_end:
        oldj = Ks.m
        arnoldi_(Ks, (A, u_flip), (w, w_aug), ishermitian_=ishermitian_, iop=iop, init=j,
                 t=tau_now, mu=mu, l=l, m=m)
        V, H = Ks.V, Ks.H
        j = Ks.m
        happy = j < oldj
        beta = Ks.beta
        H[0, j] = 1
        nrm = H[j, j - 1]
        H[j, j - 1] = 0
        F = exponential_higham2005base(sgn * tau * H[: j + 1, : j + 1])
        exps += 1
        H[j, j - 1] = nrm
        if happy:
            omega = 0
            tau_new = min(tau_end - (tau_now + tau), tau)
            m_new = m
            happy = False
        else:
            err = abs(beta * nrm * F[j - 1, j])
            oldomega = omega
            omega = tau_end * err / (tau * tol)
            if m == oldm and tau != oldtau and ireject >= 1:
                order = max(1, math.log(omega / oldomega) / math.log(tau / oldtau))
                orderold = False
            elif orderold or ireject == 0:
                orderold = True
                order = j / 4
            else:
                orderold = True
            if m != oldm and tau == oldtau and ireject >= 1:
                kest = max(1.1, (omega / oldomega) ** (1 / (oldm - m)))
                kestold = False
            elif kestold or ireject == 0:
                kestold = True
                kest = 2
            else:
                kestold = True
            if omega > delta:
                remaining_time = tau_end - tau_now
            else:
                remaining_time = tau_end - (tau_now + tau)
            same_tau = min(remaining_time, tau)
            tau_opt = tau * (gamma / omega) ** (1 / order)
            tau_opt = min(remaining_time, max(tau / 5, min(5 * tau, tau_opt)))
            m_opt = math.ceil(j + math.log(omega / gamma) / math.log(kest))
            # quirk kept verbatim: `3 ÷ 4 * m` == 0 and `cld(4, 3) * m` == 2m  (src/kiops.jl:210)
            m_opt = max(mmin, min(mmax, max(3 // 4 * m, min(m_opt, _cld(4, 3) * m))))
            if j == mmax:
                if omega > delta:
                    m_new = j
                    tau_new = tau * (gamma_mmax / omega) ** (1 / order)
                    tau_new = min(tau_end - tau_now, max(tau / 5, tau_new))
                else:
                    tau_new = tau_opt
                    m_new = m
            else:
                m_new = m_opt
                tau_new = same_tau
        if omega <= delta:
            # kiops_update_solution! -- src/kiops.jl:283-326
            reject += ireject
            step += 1
            blownTs = 0
            nextT = tau_now + tau
            for k in range(l, numSteps + 1):
                if abs(tau_flat[k - 1]) < abs(nextT):
                    blownTs += 1
            if blownTs != 0:
                w[:, l + blownTs - 1] = w[:, l - 1]
                for k in range(blownTs):
                    tauPhantom = tau_flat[l + k - 1] - tau_now
                    F2 = exponential_higham2005base(np.sign(tau_flat[-1]) * tauPhantom * H[:j, :j])
                    w[:, l + k - 1] = beta * (V[:n, :j] @ F2[:j, 0])
                l += blownTs
            w[:, l - 1] = beta * (V[:n, :j] @ F[:j, 0])
            tau_now = tau_now + tau
            j = 0
            ireject = 0
        else:
            ireject += 1
            H[0, j] = 0
        oldtau, tau = tau, tau_new
        oldm, m = m, m_new
    if tau_flat[0] != 1 and task1:
        if tau_flat.size == 1:
            w[:, l - 1] = w[:, l - 1] * (1 / tau_flat[l - 1]) ** p
        # the multi-output branch is marked FIXME in the reference (src/kiops.jl:255-274); not restated
    return w, (step, reject, krystep, exps, m)


# --------------------------------------------------------------------------------------
# phiv_timestep / expv_timestep  (src/krylov_phiv_adaptive.jl:260-501)
# --------------------------------------------------------------------------------------
def _timestep_flops(m, tau, n, p, NA, iop, Hnorm, maxtau):
    """_phiv_timestep_estimate_flops -- src/krylov_phiv_adaptive.jl:482-501."""
    flops_W = 2 * (p - 1) * (NA + n)
    flops_u = (2 * p + 1) * n
    if iop == 0:
        iop = m
    flops_matvec = 2 * m * NA
    flops_vecvec = sum(3 * min(i, iop) for i in range(1, m + 1))
    MH = 44 / 3 + 2 * math.ceil(max(0.0, math.log2(Hnorm / 5.37)))
    flops_phiv = round(MH * (m + p) ** 3)
    return (flops_W + flops_u + flops_matvec + flops_vecvec + flops_phiv) * int(math.ceil(maxtau / tau))


def _timestep_adapt(m, tau, epsilon, m_old, tau_old, epsilon_old, q, kappa, gamma, omega, maxtau, n, p, NA, iop,
                    Hnorm):
    """_phiv_timestep_adapt -- src/krylov_phiv_adaptive.jl:455-481."""
    if tau_old > tau:
        q = math.log(tau / tau_old) / math.log(epsilon / epsilon_old) - 1
    tau_new = tau * (gamma / omega) ** (1 / (q + 1))
    tau_new = min(max(tau_new, tau / 5), 2 * tau, maxtau)
    if m_old < m:
        kappa = (epsilon / epsilon_old) ** (1 / (m_old - m))
    m_new = m + math.ceil(math.log(omega / gamma) / math.log(kappa))
    m_new = min(max(m_new, (3 * m) // 4, 1), int(math.ceil(4 * m / 3)))
    cost_tau = _timestep_flops(m, tau_new, n, p, NA, iop, Hnorm, maxtau)
    cost_m = _timestep_flops(m_new, tau, n, p, NA, iop, Hnorm, maxtau)
    if cost_tau < cost_m:
        m_new = m
    else:
        tau_new = tau
    return m_new, tau_new, q, kappa


def phiv_timestep(ts, A, B, *, tau=0.0, m=None, tol=1e-7, opnorm=None, iop=0, correct=False, adaptive=False,
                  delta=1.2, ishermitian_=None, gamma=0.8, NA=0, return_steps=False):
    """phiv_timestep!(U, ts, A, B; ...) -- src/krylov_phiv_adaptive.jl:260-453.

    u(t) = phi_0(tA) b_0 + t phi_1(tA) b_1 + ... + t^p phi_p(tA) b_p at the times ``ts``; B is n x (p+1)
    (or a vector for p = 0, which is expv_timestep).  Returns U (n x len(ts)), or a vector for a scalar ``ts``."""
    scalar_t = np.isscalar(ts)
    ts = np.sort(np.atleast_1d(np.asarray(ts, dtype=float)))
    # T = promote_type(eltype(A), eltype(B)) (krylov_phiv_adaptive.jl:116-131): ComplexF64 operators / vectors are fine
    cplx = np.iscomplexobj(B) or np.iscomplexobj(A if not sp.issparse(A) else A.data)
    dt = np.complex128 if cplx else np.float64
    B = np.asarray(B, dtype=dt)
    Bm = B.reshape(B.shape[0], -1)
    n = A.shape[0]
    if m is None:
        m = min(10, n)
    if ishermitian_ is None:
        ishermitian_ = ishermitian(A)
    arnoldi_scale = opnorm is None
    abstol = None
    opn = None
    if not arnoldi_scale:
        opn = opnorm if np.isscalar(opnorm) else opnorm(A, np.inf)
        abstol = tol * opn
        if tau == 0:
            b0norm = np.abs(Bm[:, 0]).max()
            tau = 10 / opn * (abstol * ((m + 1) / math.e) ** (m + 1) * math.sqrt(2 * math.pi * (m + 1)) /
                              (4 * opn * b0norm)) ** (1 / m)
    tend = ts[-1]
    seed_arnoldi_tau = arnoldi_scale and tau == 0
    if seed_arnoldi_tau:
        tau = tend
    p = Bm.shape[1] - 1
    U = np.zeros((n, ts.size), order="F", dtype=dt)
    u = Bm[:, 0].copy()
    W = np.zeros((n, p + 1), order="F", dtype=dt)
    Ks = KrylovSubspace(n, m, dtype=dt, hdtype=(np.float64 if (cplx and ishermitian_) else dt))
    coeffs = np.ones(max(p, 0))
    if adaptive:
        if ishermitian_:
            iop = 2
        if NA == 0:
            NA = A.nnz if sp.issparse(A) else int(np.count_nonzero(A))
    t = 0.0
    snapshot = 1
    nsteps = 0
    while t < tend:
        if t + tau > tend:
            tau = tend - t
        W[:, 0] = u
        for l in range(1, p):
            coeffs[l] = coeffs[l - 1] * t / l
        for j in range(1, p + 1):
            W[:, j] = _mul(A, W[:, j - 1])
            for l in range(0, p - j + 1):
                W[:, j] += coeffs[l] * Bm[:, j + l]
        arnoldi_(Ks, A, W[:, p].copy(), tol=tol, m=m, iop=iop)
        if abstol is None:
            opn = np.linalg.norm(Ks.getH(), 1)
            abstol = tol * opn
            if seed_arnoldi_tau:
                b0norm = np.abs(Bm[:, 0]).max()
                tau = min(tend - t, gamma * 10 / opn * (abstol * ((m + 1) / math.e) ** (m + 1) *
                                                      math.sqrt(2 * math.pi * (m + 1)) / (4 * opn * b0norm)) ** (1 / m))
        if Ks.wasbreakdown:
            tau = tend - t
        P, epsilon = phiv_ks(tau, Ks, p + 1, correct=correct, errest=True)
        if adaptive:
            omega = (tend / tau) * (epsilon / abstol)
            epsilon_old, m_old, tau_old = epsilon, m, tau
            q, kappa = m / 4, 2.0
            maxtau = tend - t
            while omega > delta:
                m_new, tau_new, q, kappa = _timestep_adapt(m, tau, epsilon, m_old, tau_old, epsilon_old, q, kappa,
                                                           gamma, omega, maxtau, n, p, NA, iop,
                                                           np.linalg.norm(Ks.getH(), 1))
                m, m_old = m_new, m
                tau, tau_old = tau_new, tau
                arnoldi_(Ks, A, W[:, p].copy(), tol=tol, m=m, iop=iop)
                P, epsilon_new = phiv_ks(tau, Ks, p + 1, correct=correct, errest=True)
                epsilon, epsilon_old = epsilon_new, epsilon
                omega = (tend / tau) * (epsilon / abstol)
        u = tau ** p * P[:, -2]
        for l in range(1, p):
            coeffs[l] = coeffs[l - 1] * tau / l
        for j in range(0, p):
            u = u + coeffs[j] * W[:, j]
        while snapshot <= ts.size and t + tau >= ts[snapshot - 1]:
            tau_s = ts[snapshot - 1] - t
            Ps = phiv_ks(tau_s, Ks, p + 1, correct=correct)
            us = tau_s ** p * Ps[:, -2]
            for l in range(1, p):
                coeffs[l] = coeffs[l - 1] * tau_s / l
            for j in range(0, p):
                us = us + coeffs[j] * W[:, j]
            U[:, snapshot - 1] = us
            snapshot += 1
        t += tau
        nsteps += 1
    out = U[:, 0] if scalar_t else U
    return (out, nsteps) if return_steps else out


def expv_timestep(ts, A, b, **kw):
    """expv_timestep(ts, A, b; ...) -- src/krylov_phiv_adaptive.jl:57-114 (phiv_timestep with p = 0)."""
    return phiv_timestep(ts, A, np.asarray(b).reshape(-1), **kw)


# --------------------------------------------------------------------------------------
# expv(...; mode = :error_estimate)  (src/krylov_phiv_error_estimate.jl:149-207, src/krylov_phiv.jl:145-160)
# --------------------------------------------------------------------------------------
def expv_ee(t, A, b, *, m=None, tol=1e-7, rtol=None, return_m=False):
    """_expv_ee + expv!(w, t, A, b, Ks, cache; atol = tol, rtol = sqrt(tol)): Lanczos stopped by Saad's estimate
    sigma_j = beta_j * beta * |(exp(t T_j) e_1)_j| < atol + rtol * beta.  Hermitian A only."""
    if not ishermitian(A):
        raise ValueError("Error estimation not yet available for non-Hermitian matrices.")
    n = A.shape[0]
    if m is None:
        m = min(30, n)
    if rtol is None:
        rtol = math.sqrt(tol)
    Ks = KrylovSubspace(n, m)
    V, H = Ks.getV(), Ks.getH()
    Ks.beta = _nrm2(np.ascontiguousarray(b))
    if Ks.beta == 0:
        return (np.zeros(n), 0) if return_m else np.zeros(n)
    V[:, 0] = b / Ks.beta
    eps = tol + rtol * Ks.beta
    v = None
    for j in range(1, m + 1):
        _lanczos_step(j, A, V, H, -1, -1)
        alpha = np.diag(H)[:j].copy()
        beta = np.diag(H, -1)[:j].copy()
        lam, Z = sla.eigh_tridiagonal(alpha, beta[: j - 1], lapack_driver="stemr") if j > 1 else (alpha, np.ones((1, 1)))
        v = Z @ (np.exp(t * lam) * Z[0, :])
        sigma = beta[j - 1] * Ks.beta * abs(v[j - 1])
        if sigma < eps:
            Ks.m = j
            break
    w = Ks.beta * (V[:, : Ks.m] @ v[: Ks.m])
    return (w, Ks.m) if return_m else w
