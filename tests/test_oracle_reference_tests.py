"""Pin the CPU oracle to the reference's own analytic assertions for the hot path
(/root/reference/test/basictests.jl; Julia RNG streams are irrelevant, NumPy seeds are used)."""
import numpy as np
import pytest
import scipy.linalg as sla
import scipy.sparse as sp

from conftest import laplacian2d, relerr

SQRT_EPS = np.sqrt(np.finfo(float).eps)  # isapprox default rtol


def phis_dense(A, b, K):
    """[phi_0(A) b, ..., phi_K(A) b] through scipy.linalg.expm of the Sidje block matrix (independent of oracle)."""
    n = A.shape[0]
    M = np.zeros((n + K, n + K))
    M[:n, :n] = A
    M[:n, n] = b
    for i in range(n, n + K - 1):
        M[i, i + 1] = 1
    E = sla.expm(M)
    return np.stack([E[:n, :n] @ b] + [E[:n, n + i] for i in range(K)], 1)


def test_arnoldi_and_krylov(oracle):
    """test/basictests.jl:515-574 "Arnoldi & Krylov"."""
    O = oracle
    rng = np.random.default_rng(0)
    n, m, K = 20, 5, 4
    A = rng.standard_normal((n, n))
    t = 1e-2
    b = rng.standard_normal(n)
    direct = sla.expm(t * A) @ b
    assert relerr(O.expv(t, A, b, m=m), direct) < SQRT_EPS                     # :524
    assert relerr(O.kiops(t, A, b)[0][:, 0], direct) < SQRT_EPS                # :526
    W = phis_dense(t * A, b, K)
    Ks = O.arnoldi(A, b, m=m)
    assert relerr(O.phiv_ks(t, Ks, K), W) < SQRT_EPS                           # :527-534
    U = np.stack([b * (1 / t) ** i for i in range(K)], 1)
    assert relerr(O.kiops(t, A, U)[0][:, 0], W[:, :K].sum(1)) < SQRT_EPS       # :535-536
    v = rng.standard_normal(n)
    v /= np.linalg.norm(v)
    P = np.outer(v, v)
    assert O.arnoldi(P, b).m == 2                                              # :544-547 happy breakdown
    z = np.zeros(n)
    assert np.linalg.norm(O.expv(t, P, z, m=m)) == 0.0                         # :550-553
    S = rng.standard_normal((n, n))
    S = S + S.T
    Sp = S + 1e-10 * rng.standard_normal((n, n))
    w = O.expv(t, S, b, m=m)
    assert relerr(O.expv(t, Sp, b, m=m), w) < SQRT_EPS                         # :556-562
    assert relerr(O.kiops(t, S, b, m=m)[0][:, 0], w) < SQRT_EPS
    assert np.linalg.norm(O.expv(t, S, z, m=m)) == 0.0                         # :565-566
    n = 30
    T3 = np.diag(np.ones(n - 1), -1) + np.diag(30 * np.ones(n)) + np.diag(np.ones(n - 1), 1)
    t = 0.1
    Q = O.phiv(t, T3, np.ones(n), 10)
    ref = np.linalg.solve(t * T3, (sla.expm(t * T3) - np.eye(n)) @ np.ones(n))
    assert relerr(Q[:, 1], ref) < SQRT_EPS                                     # :569-573


def test_hermitian_arnoldi_vs_lanczos_H(oracle):
    """test/basictests.jl:731-754: arnoldi! and lanczos! give the same H (real symmetric restatement)."""
    O = oracle
    n, m = 100, 15
    rng = np.random.default_rng(3)
    d = rng.standard_normal(n)
    e = rng.standard_normal(n - 1)
    A = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    b = rng.standard_normal(n)
    Ka = O.arnoldi(A, b, m=m, ishermitian_=False)
    Kl = O.arnoldi(A, b, m=m, ishermitian_=True)
    assert np.abs(Ka.getH() - Kl.getH()).max() < 1e-12


def test_matrix_free_interface(oracle):
    """test/basictests.jl:786-816: expv(...; m = n) reproduces exp(tA)b to 1e-12 on a small operator."""
    O = oracle
    n = 10
    rng = np.random.default_rng(5)
    A = rng.standard_normal((n, n)) / 4
    b = rng.standard_normal(n)
    w = O.expv(0.3, A, b, m=n)
    assert np.abs(w - sla.expm(0.3 * A) @ b).max() < 1e-12


def test_small_exp_all_pade_branches(oracle):
    """test/basictests.jl:952-974: every Pade branch (C13, C9, C7, C5, C3), n = 40, relerr < 1e-11."""
    rng = np.random.default_rng(7)
    for scale in (3.0, 1.5, 0.5, 0.1, 0.005):
        A = rng.standard_normal((40, 40))
        A *= scale / np.linalg.norm(A, 1)
        assert relerr(oracle.exponential_higham2005base(A), sla.expm(A)) < 1e-11


def test_sparse_laplacian_converged_vs_expm_multiply(oracle):
    """Converged Krylov on the C2-style operator agrees with scipy's independent expm_multiply."""
    from scipy.sparse.linalg import expm_multiply
    A = laplacian2d(30, 20)
    b = np.random.default_rng(0).standard_normal(600)
    ref = expm_multiply(A.tocsc(), b)
    assert relerr(oracle.expv(1.0, A, b, m=30), ref) < 1e-12                    # Lanczos dispatch
    assert relerr(oracle.expv(1.0, A, b, m=30, ishermitian_=False), ref) < 1e-12


def test_kiops_quirks_and_stats(oracle):
    """stats = (steps, rejected, krystep == 0, exps, m) and the Hermitian/IOP agreement (basictests.jl:560-562)."""
    A = laplacian2d(12, 10)
    u = np.random.default_rng(4).standard_normal((120, 2))
    w1, s1 = oracle.kiops(1.0, A, u, ishermitian_=True)
    w2, s2 = oracle.kiops(1.0, A, u, ishermitian_=False)
    assert s1[2] == 0 and s2[2] == 0
    assert relerr(w1, w2) < 1e-6
    W = phis_dense(1.0 * A.toarray(), u[:, 0], 1)  # exp(A) u0 + phi_1(A) u1
    M = phis_dense(1.0 * A.toarray(), u[:, 1], 2)
    assert relerr(w2[:, 0], W[:, 0] + M[:, 1]) < 1e-6


def test_adaptive_krylov_timestepping(oracle):
    """test/basictests.jl:666-691 "Adaptive Krylov" and :693-729 (Arnoldi-estimated tolerance scale)."""
    n, K, t, tol = 100, 4, 5.0, 1e-7
    A = sp.diags([np.ones(n - 1), -2 * np.ones(n), np.ones(n - 1)], [-1, 0, 1]).tocsr()
    B = np.random.default_rng(14).standard_normal((n, K + 1))
    Ad = A.toarray()

    def exact(tt):
        return sum(tt ** i * phis_dense(tt * Ad, B[:, i], max(i, 1))[:, i] for i in range(K + 1))

    U = oracle.phiv_timestep([t / 2, t], A, B, adaptive=True, tol=tol)
    assert relerr(U[:, 0], exact(t / 2)) < tol and relerr(U[:, 1], exact(t)) < tol          # :680-682
    ue = sla.expm(t * Ad) @ B[:, 0]
    opn = abs(A).sum(axis=1).max()
    for on in (lambda A_, p_: abs(A_).sum(axis=1).max(), opn):                                # :685-690
        assert relerr(oracle.expv_timestep(t, A, B[:, 0], adaptive=True, tol=tol, opnorm=on), ue) < tol
    assert relerr(oracle.expv_timestep(t, A, B[:, 0], adaptive=True, tol=tol), ue) < 1e-5    # :716-717


def test_error_estimate_lanczos_mode(oracle):
    """test/basictests.jl:756-784 "Alternative Lanczos expv Interface": n = 300, m = 30, atol = rtol = 1e-10."""
    n, m = 300, 30
    rng = np.random.default_rng(8)
    d = rng.standard_normal(n)
    e = rng.standard_normal(n - 1)
    A = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    b = rng.standard_normal(n)
    t = 0.1
    w, mm = oracle.expv_ee(t, A, b, m=m, tol=1e-10, rtol=1e-10, return_m=True)
    assert relerr(w, sla.expm(t * A) @ b) < 1e-9 and 1 <= mm <= m
    assert np.linalg.norm(oracle.expv_ee(t, A, np.zeros(n), m=m)) == 0.0          # :783


def test_complex_value_testset(oracle):
    """test/basictests.jl:650-664 "Complex Value": Hermitian / general, complex / real A, complex / real b, real /
    imaginary / complex t, n = 20, m = 10: exp(t A) b ~ expv(t, A, b; m) at isapprox's default rtol."""
    n, m = 20, 10
    rng = np.random.default_rng(21)

    def herm(M):
        return (M + M.conj().T) / 2

    Az = rng.random((n, n)) + 1j * rng.random((n, n))
    Ar = rng.random((n, n))
    for A in (herm(Az), herm(Ar), Az, Ar):
        for b in (rng.random(n) + 1j * rng.random(n), rng.random(n)):
            for t in (1e-2, 1e-2j, 1e-2 + 1e-2j):
                w = oracle.expv(t, A, b, m=m)
                assert relerr(w, sla.expm(t * A) @ b) < SQRT_EPS


def test_gpu_testset_inputs_on_the_oracle(oracle):
    """test/gpu/gputests.jl:41-58, 62-77: the inputs of the reference's own GPU test (sparse ComplexF64 strictly upper
    triangular + sparse perturbation, n = 1000; a 4 x 4 complex Arnoldi with imaginary dt) pin the oracle to
    exp(tA)b -- the GPU tests then compare the CUDA path with the oracle on the same kind of input."""
    n = 1000
    rng = np.random.default_rng(22)

    def sprand_c(density):
        M = sp.random(n, n, density=density, random_state=rng, format="csr")
        M.data = M.data + 1j * rng.random(M.nnz)
        return M

    A = (sp.triu(sprand_c(10 / n), 1) + sprand_c(1 / n)).tocsr()
    b = rng.random(n) + 1j * rng.random(n)
    w = oracle.expv(0.1, A, b)
    assert relerr(w, sla.expm(0.1 * A.toarray()) @ b) < SQRT_EPS
    A4 = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))
    v0 = rng.standard_normal(4) + 1j * rng.standard_normal(4)
    Ks = oracle.arnoldi(A4, v0, tol=1e-7, ishermitian_=False)
    assert relerr(oracle.expv_ks(0.01j, Ks), sla.expm(0.01j * A4) @ v0) < SQRT_EPS



# ---- the C/OpenMP restatement (oracle/cpu_krylov.c, the "best-effort CPU" baseline) against the NumPy port ----------
def test_c_openmp_restatement_matches_numpy_port():
    from conftest import convdiff2d, laplacian2d, relerr
    from oracle import cpu_fast as F
    from oracle import oracle as O
    rng = np.random.default_rng(21)
    for A, herm, iop in ((laplacian2d(60, 45), True, 0), (laplacian2d(60, 45), False, 0), (convdiff2d(50, 40), False, 0),
                         (convdiff2d(50, 40), False, 2)):
        b = rng.standard_normal(A.shape[0])
        V, H, beta, mo, bd = F.arnoldi(A, b, m=25, iop=iop, ishermitian_=herm)
        Ko = O.arnoldi(A, b, m=25, iop=iop, ishermitian_=herm)
        assert mo == Ko.m and bd == Ko.wasbreakdown and abs(beta - Ko.beta) <= 1e-14 * Ko.beta
        assert np.abs(H[:26, :25] - Ko.getH()).max() < 1e-12
        assert relerr(V[:26].T, Ko.getV()) < 1e-10
        assert relerr(F.expv(0.7, A, b, m=25, iop=iop, ishermitian_=herm), O.expv(0.7, A, b, m=25, iop=iop, ishermitian_=herm)) < 1e-12
    # happy breakdown and the zero vector
    import scipy.sparse as sp
    v = rng.standard_normal(30)
    v /= np.linalg.norm(v)
    Pm = sp.csr_matrix(np.outer(v, v))
    V, H, beta, mo, bd = F.arnoldi(Pm, rng.standard_normal(30), m=10)
    assert mo == 2 and bd
    assert np.all(F.expv(1.0, Pm, np.zeros(30), m=10) == 0.0)
