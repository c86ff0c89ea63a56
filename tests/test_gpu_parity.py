"""GPU parity tests: the CUDA path (through the C ABI, via the Python host mirror) against the CPU oracle
on identical seeded inputs.  Tolerance: relative 2-norm error <= 1e-10 (fp64), the bar of BASELINE.json.
Run on the B200 box with `pytest -m gpu`."""
import numpy as np
import pytest
import scipy.linalg as sla
import scipy.sparse as sp

from conftest import convdiff2d, laplacian2d, relerr

pytestmark = pytest.mark.gpu
RTOL = 1e-10


@pytest.fixture(scope="module")
def gpu(eu):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (no CPU fallback exists)")
    eng = eu.get_engine()
    assert eng.device_info()["sm_count"] > 0
    return eu


# ---- reference test "Arnoldi & Krylov" (test/basictests.jl:515-574) on the GPU path --------------------
def test_reference_arnoldi_krylov_suite(gpu, oracle):
    eu, O = gpu, oracle
    rng = np.random.default_rng(0)
    n, m, K = 20, 5, 4
    A = rng.standard_normal((n, n))
    t = 1e-2
    b = rng.standard_normal(n)
    direct = sla.expm(t * A) @ b
    w = eu.expv(t, A, b, m=m)
    assert relerr(w, direct) < 1.5e-8                       # the reference's own assertion
    assert relerr(w, O.expv(t, A, b, m=m)) < RTOL           # parity with the restated reference
    wk, st = eu.kiops(t, A, b)
    wo, so = O.kiops(t, A, b)
    assert relerr(wk, direct[:, None]) < 1.5e-8 and relerr(wk, wo) < RTOL and st == so
    Ks = eu.arnoldi(A, b, m=m)
    Ko = O.arnoldi(A, b, m=m)
    W = eu.phiv(t, Ks, K).cpu().numpy()
    assert relerr(W, O.phiv_ks(t, Ko, K)) < RTOL
    # happy breakdown: A = v v' is idempotent -> Ks.m == 2
    v = rng.standard_normal(n)
    v /= np.linalg.norm(v)
    Ks = eu.arnoldi(np.outer(v, v), b)
    assert Ks.m == 2 and Ks.wasbreakdown
    # zero input -> exactly zero output, Arnoldi and Lanczos
    z = np.zeros(n)
    assert np.linalg.norm(eu.expv(t, np.outer(v, v), z, m=m, ishermitian=False)) == 0.0
    S = rng.standard_normal((n, n))
    S = S + S.T
    assert np.linalg.norm(eu.expv(t, S, z, m=m)) == 0.0
    # Arnoldi vs Lanczos vs kiops agree on a Hermitian matrix and a 1e-10 perturbation
    Sp = S + 1e-10 * rng.standard_normal((n, n))
    w = eu.expv(t, S, b, m=m)
    assert relerr(eu.expv(t, Sp, b, m=m), w) < 1.5e-8
    assert relerr(eu.kiops(t, S, b, m=m)[0][:, 0], w) < 1.5e-8
    assert relerr(w, O.expv(t, S, b, m=m)) < RTOL
    # tridiagonal phiv
    n = 30
    T3 = np.diag(np.ones(n - 1), -1) + np.diag(30 * np.ones(n)) + np.diag(np.ones(n - 1), 1)
    Q = eu.phiv(0.1, T3, np.ones(n), 10)
    assert relerr(Q[:, 1], np.linalg.solve(0.1 * T3, (sla.expm(0.1 * T3) - np.eye(n)) @ np.ones(n))) < 1.5e-8


def test_kiops_multi_column_matches_oracle(gpu, oracle):
    rng = np.random.default_rng(5)
    n, t = 20, 1e-2
    A = rng.standard_normal((n, n))
    b = rng.standard_normal(n)
    U = np.stack([b * (1 / t) ** i for i in range(4)], 1)
    w, st = gpu.kiops(t, A, U)
    wo, so = oracle.kiops(t, A, U)
    assert st == so
    # u has columns of magnitude 1e6: the result is conditioned at ~1e-10 relative to |u|, compare absolutely
    assert np.linalg.norm(w - wo) / np.abs(U).max() < 1e-12


# ---- BASELINE configs --------------------------------------------------------------------------------------
def test_C1_dense_512(gpu, oracle):
    n = 512
    b = np.random.default_rng(1).standard_normal(n)
    for scale in (1 / np.sqrt(n), 1.0):  # the unscaled variant exercises 4 squarings
        A = np.random.default_rng(0).standard_normal((n, n)) * scale
        assert relerr(gpu.expv(1.0, A, b, m=30), oracle.expv(1.0, A, b, m=30)) < RTOL
    import torch
    At = torch.from_numpy(A).cuda()
    bt = torch.from_numpy(b).cuda()
    w = gpu.expv(1.0, At, bt, m=30)  # device-resident operator and vector
    assert w.is_cuda and relerr(w.cpu().numpy(), oracle.expv(1.0, A, b, m=30)) < RTOL


@pytest.mark.parametrize("nx,ny", [(40, 50), (37, 41), (300, 400), (1000, 1000)])
def test_C2_laplacian_arnoldi_and_lanczos(gpu, oracle, nx, ny):
    A = laplacian2d(nx, ny)
    b = np.random.default_rng(0).standard_normal(nx * ny)
    op = gpu.operator(A)
    assert op.ishermitian and abs(op.opnorm_inf - 8.0) < 1e-14
    w_l = gpu.expv(1.0, op, b, m=30)                       # default dispatch -> Lanczos
    w_a = gpu.expv(1.0, op, b, m=30, ishermitian=False)    # full Arnoldi: the fused CGS kernel
    assert relerr(w_l, oracle.expv(1.0, A, b, m=30)) < RTOL
    assert relerr(w_a, oracle.expv(1.0, A, b, m=30, ishermitian_=False)) < RTOL


def test_C2_nonsymmetric_convection_diffusion(gpu, oracle):
    A = convdiff2d(300, 400)
    b = np.random.default_rng(0).standard_normal(120000)
    op = gpu.operator(A)
    assert not op.ishermitian
    assert relerr(gpu.expv(1.0, op, b, m=30), oracle.expv(1.0, A, b, m=30)) < RTOL
    assert relerr(gpu.expv(1.0, op, b, m=30, iop=2), oracle.expv(1.0, A, b, m=30, iop=2)) < RTOL


def test_C3_phiv_dense(gpu, oracle):
    n = 2048  # C3 is n = 16384; the oracle's dense mat-vecs keep this test at a size that runs in seconds
    A = np.random.default_rng(2).standard_normal((n, n)) / np.sqrt(n) * 4
    b = np.random.default_rng(3).standard_normal(n)
    W, e = gpu.phiv(1.0, A, b, 4, m=30, correct=True, errest=True)
    Wo, eo = oracle.phiv(1.0, A, b, 4, m=30, correct=True, errest=True)
    assert relerr(W, Wo) < RTOL and abs(e - eo) <= 1e-10 * abs(eo)
    W = gpu.phiv(1.0, A, b, 4, m=30)
    assert relerr(W, oracle.phiv(1.0, A, b, 4, m=30)) < RTOL


def test_C4_kiops_laplacian(gpu, oracle):
    A = laplacian2d(250, 400)
    u = np.stack([np.random.default_rng(4).standard_normal(100000),
                  np.random.default_rng(5).standard_normal(100000)], 1)
    op = gpu.operator(A)
    for herm in (True, False):  # Lanczos-in-kiops (reference default) and IOP-2
        w, st = gpu.kiops(1.0, op, u, ishermitian=herm)
        wo, so = oracle.kiops(1.0, A, u, ishermitian_=herm)
        assert st == so, (st, so)
        assert relerr(w, wo) < RTOL


def test_C5_batched_expv(gpu, oracle):
    A = laplacian2d(250, 400)
    n, nb = 100000, 24
    B = np.random.default_rng(6).standard_normal((n, nb))
    ts = np.random.default_rng(7).uniform(0.1, 1.0, nb)
    op = gpu.operator(A)
    for herm in (True, False):
        W = gpu.expv_batched(ts, op, B, m=30, ishermitian=herm)
        for i in (0, 7, nb - 1):
            assert relerr(W[:, i], oracle.expv(ts[i], A, B[:, i], m=30, ishermitian_=herm)) < RTOL
    # zero columns inside a batch give exactly zero
    B[:, 3] = 0.0
    W = gpu.expv_batched(ts, op, B, m=30)
    assert np.linalg.norm(W[:, 3]) == 0.0
    assert relerr(W[:, 4], oracle.expv(ts[4], A, B[:, 4], m=30)) < RTOL


# ---- factorisation-level parity, continuation, layouts ---------------------------------------------------
def test_krylov_factorisation_H_V(gpu, oracle):
    A = convdiff2d(40, 50)
    b = np.random.default_rng(3).standard_normal(2000)
    for iop in (0, 2, 5):
        Ks = gpu.arnoldi(A, b, m=20, iop=iop)
        Ko = oracle.arnoldi(A, b, m=20, iop=iop)
        assert Ks.m == Ko.m and Ks.beta == pytest.approx(Ko.beta, rel=1e-14)
        assert np.abs(Ks.getH() - Ko.getH()).max() < 1e-11
        assert np.abs(Ks.getV().cpu().numpy() - Ko.getV()).max() < 1e-11
    L = laplacian2d(40, 50)
    Ks = gpu.arnoldi(L, b, m=20)
    Ko = oracle.arnoldi(L, b, m=20)
    H = Ks.getH()
    assert np.array_equal(H[:20, :20], H[:20, :20].T)  # mirrored exactly -> the eigen branch is taken
    assert np.abs(H - Ko.getH()).max() < 1e-11
    # the basis is orthonormal and satisfies the Arnoldi relation A V_m = V_{m+1} H
    Ks = gpu.arnoldi(A, b, m=20)
    V = Ks.getV().cpu().numpy()
    assert np.abs(V.T @ V - np.eye(21)).max() < 1e-10
    assert np.abs(A @ V[:, :20] - V @ Ks.getH()).max() < 1e-10


def test_arnoldi_continuation_init(gpu, oracle):
    """arnoldi!(...; init = j) resumes an existing factorisation (src/arnoldi.jl:360-368)."""
    A = convdiff2d(30, 30)
    b = np.random.default_rng(9).standard_normal(900)
    Ks = gpu.KrylovSubspace(900, 20)
    gpu.arnoldi_(Ks, A, b, m=10, ishermitian=False)
    H10 = Ks.H.copy()
    gpu.arnoldi_(Ks, A, b, m=20, ishermitian=False, init=10)
    Ko = oracle.KrylovSubspace(900, 20)
    oracle.arnoldi_(Ko, A, b, m=20, ishermitian_=False)
    assert np.abs(Ks.getH() - Ko.getH()).max() < 1e-11
    assert np.abs(H10[:10, :9] - Ks.H[:10, :9]).max() == 0.0
    w = gpu.expv(0.7, Ks)
    assert relerr(w.cpu().numpy(), oracle.expv_ks(0.7, Ko)) < RTOL


def test_dimension_errors(gpu):
    A = laplacian2d(10, 10)
    with pytest.raises(gpu.DimensionMismatch):
        gpu.expv(1.0, A, np.ones(99))
    Ks = gpu.KrylovSubspace(101, 10)
    with pytest.raises(gpu.DimensionMismatch):
        gpu.arnoldi_(Ks, A, np.ones(100))
    with pytest.raises(gpu.ArgumentError):
        gpu.expv(1.0, A, np.ones(100), mode="nonsense")


def test_operator_mul_and_long_rows(gpu, oracle):
    """mul! parity, and an operator with long rows (warp-per-row mat-vec path) through expv."""
    rng = np.random.default_rng(21)
    n = 3000
    A = sp.random(n, n, density=0.05, random_state=3, format="csr") - 6 * sp.identity(n)
    A = A.tocsr()
    x = rng.standard_normal(n)
    op = gpu.operator(A)
    assert relerr(op.mul(x), A @ x) < 1e-13
    assert relerr(gpu.expv(0.5, op, x, m=25), oracle.expv(0.5, A, x, m=25)) < RTOL
    D = rng.standard_normal((300, 300))
    assert relerr(gpu.operator(D).mul(x[:300]), D @ x[:300]) < 1e-13


def test_run_to_run_determinism(gpu):
    A = convdiff2d(200, 200)
    b = np.random.default_rng(0).standard_normal(40000)
    op = gpu.operator(A)
    w1 = gpu.expv(1.0, op, b, m=30)
    w2 = gpu.expv(1.0, op, b, m=30)
    assert np.array_equal(w1, w2)


def test_full_size_properties_C2(gpu):
    """Size-independent properties at the full C2 size: linearity in b and the semigroup identity
    exp(tA) exp(sA) b = exp((t+s)A) b (converged Krylov, so it holds to solver accuracy)."""
    import torch
    A = laplacian2d(1000, 1000)
    op = gpu.operator(A)
    g = torch.Generator(device="cuda").manual_seed(0)
    b1 = torch.randn(10 ** 6, dtype=torch.float64, device="cuda", generator=g)
    b2 = torch.randn(10 ** 6, dtype=torch.float64, device="cuda", generator=g)
    for herm in (True, False):
        w1 = gpu.expv(0.5, op, b1, m=30, ishermitian=herm)
        w2 = gpu.expv(0.5, op, b2, m=30, ishermitian=herm)
        w12 = gpu.expv(0.5, op, 2.0 * b1 - 3.0 * b2, m=30, ishermitian=herm)
        assert float(torch.linalg.norm(w12 - (2.0 * w1 - 3.0 * w2)) / torch.linalg.norm(w12)) < 1e-9
        ww = gpu.expv(0.25, op, gpu.expv(0.25, op, b1, m=30, ishermitian=herm), m=30, ishermitian=herm)
        assert float(torch.linalg.norm(ww - w1) / torch.linalg.norm(w1)) < 1e-9


def test_kernel_variants_agree(gpu, oracle):
    """LDG kernel vs TMA-ring kernel, and host vs device small exponential, on the same problems."""
    eng = gpu.get_engine()
    A = convdiff2d(120, 100)
    L = laplacian2d(120, 100)
    b = np.random.default_rng(0).standard_normal(12000)
    ref_a = oracle.expv(0.8, A, b, m=30)
    ref_l = oracle.expv(0.8, L, b, m=30)
    results = {}
    try:
        for ldg in (False, True):
            for hostexp in (False, True):
                eng.set_flag("force_ldg", ldg)
                eng.set_flag("host_smallexp", hostexp)
                wa = gpu.expv(0.8, A, b, m=30)
                assert eng.last_kernel() == ("ldg" if ldg else "tma")
                wl = gpu.expv(0.8, L, b, m=30)
                assert relerr(wa, ref_a) < RTOL and relerr(wl, ref_l) < RTOL
                results[(ldg, hostexp)] = (wa, wl)
    finally:
        eng.set_flag("force_ldg", False)
        eng.set_flag("host_smallexp", False)
    base = results[(False, False)]
    for k, v in results.items():
        assert relerr(v[0], base[0]) < 1e-12 and relerr(v[1], base[1]) < 1e-12, k
    # the two Krylov kernels use the same reduction order: identical factorisations
    eng.set_flag("force_ldg", True)
    K1 = gpu.arnoldi(A, b, m=20)
    eng.set_flag("force_ldg", False)
    K2 = gpu.arnoldi(A, b, m=20)
    assert np.abs(K1.getH() - K2.getH()).max() < 1e-12


def test_short_window_instance_xl(gpu, oracle):
    """The short-window (XL) instance of the TMA-ring kernel -- resident basis vector in shared memory, inner product
    fused into the mat-vec, lazily stored basis columns, packet all-reduces -- against the general instance and
    the oracle: Lanczos, IOP-2, IOP-6 (window wider than one packet set), continuation, kiops, a batch."""
    eng = gpu.get_engine()
    A = convdiff2d(130, 90)
    L = laplacian2d(130, 90)
    n = 130 * 90
    b = np.random.default_rng(5).standard_normal(n)
    opA, opL = gpu.operator(A), gpu.operator(L)
    res = {}
    try:
        # (False: default = XL instance, two-reduction Lanczos step; "lz1": its one-reduction step (the default of
        #  row-sharded operators); True: general instance)
        for no_xl in (False, "lz1", True):
            eng.set_flag("no_xl", no_xl is True)
            eng.set_flag("no_lz1", 2 if no_xl == "lz1" else 0)
            want = "tma" if no_xl is True else "tma_xl"
            r = {}
            Ks = gpu.arnoldi(opL, b, m=30)  # Hermitian -> lanczos!
            assert eng.last_kernel() == ("tma_xl1" if no_xl == "lz1" else want)
            r["lan_H"], r["lan_V"], r["lan_beta"] = Ks.getH().copy(), Ks.getV().cpu().numpy().copy(), Ks.beta
            r["lan_w"] = gpu.expv(0.8, opL, b, m=30)
            for q in (2, 6):
                Ks = gpu.arnoldi(opA, b, m=30, iop=q)
                assert eng.last_kernel() == want
                r[f"iop{q}_H"], r[f"iop{q}_V"] = Ks.getH().copy(), Ks.getV().cpu().numpy().copy()
                r[f"iop{q}_w"] = gpu.expv(0.8, opA, b, m=30, iop=q)
            Ks = gpu.KrylovSubspace(n, 20)
            gpu.arnoldi_(Ks, opA, b, m=10, iop=3)
            gpu.arnoldi_(Ks, opA, b, m=20, iop=3, init=10)
            r["cont_H"], r["cont_V"] = Ks.getH().copy(), Ks.getV().cpu().numpy().copy()
            u = np.stack([b, 0.3 * b[::-1]], 1)
            for herm in (True, False):
                w, st = gpu.kiops(1.0, opL if herm else opA, u, ishermitian=herm)
                r[f"kiops{int(herm)}"], r[f"kiops{int(herm)}_st"] = w, st
            Bm = np.random.default_rng(6).standard_normal((n, 5))
            r["batch"] = gpu.expv_batched([0.2, 0.4, 0.6, 0.8, 1.0], opL, Bm, m=25)
            res[no_xl] = r
    finally:
        eng.set_flag("no_xl", False)
        eng.set_flag("no_lz1", False)
    g = res[True]
    for variant in (False, "lz1"):
        x = res[variant]
        for k in x:
            if k.endswith("_st"):
                assert x[k] == g[k], (variant, k)
            elif k == "lan_beta":
                assert abs(x[k] - g[k]) <= 1e-14 * abs(g[k])
            else:
                assert relerr(x[k], g[k]) < 1e-11, (variant, k, relerr(x[k], g[k]))
    x = res[False]
    assert relerr(x["lan_w"], oracle.expv(0.8, L, b, m=30)) < RTOL
    assert relerr(x["iop2_w"], oracle.expv(0.8, A, b, m=30, iop=2)) < RTOL
    assert relerr(x["iop6_w"], oracle.expv(0.8, A, b, m=30, iop=6)) < RTOL
    Ko = oracle.arnoldi(A, b, m=20, iop=3)
    assert np.abs(x["cont_H"] - Ko.getH()).max() < 1e-10
    # orthonormality of neighbouring Lanczos vectors and the three-term relation A V_m = V_{m+1} H
    V, H = x["lan_V"], x["lan_H"]
    assert np.abs(V[:, :5].T @ V[:, :5] - np.eye(5)).max() < 1e-12
    assert relerr(L @ V[:, :30], V @ H) < 1e-12


def test_batched_multivector_lanczos(gpu, oracle):
    """The lock-step multi-vector Lanczos kernel of batched expv (four problems per team, krylov_kernel_mv.cuh) against
    per-problem teams and the oracle: group padding (nb = 9), a zero start vector, mixed times, different m."""
    eng = gpu.get_engine()
    L = laplacian2d(90, 70)
    n, nb = 6300, 9
    rng = np.random.default_rng(31)
    B = rng.standard_normal((n, nb))
    B[:, 5] = 0.0
    ts = rng.uniform(0.1, 1.0, nb)
    op = gpu.operator(L)
    out = {}
    try:
        for no_mv in (False, True):
            eng.set_flag("no_mv", 1 if no_mv else 2)  # 2: always (the cost model prefers per-problem teams for 9 problems)
            for m in (30, 7):
                out[(no_mv, m)] = gpu.expv_batched(ts, op, B, m=m)
                assert eng.last_kernel() == ("tma_xl" if no_mv else "tma_mv")
    finally:
        eng.set_flag("no_mv", 0)
    # problems of one group that break down at different steps (Krylov dimensions 3, 2, 0, 3, 1) while others run on
    import scipy.sparse as sp
    nd = 6000
    d = np.array([1.0, 2.0, 3.0])[np.arange(nd) % 3]
    Dg = sp.diags(d).tocsr()
    Bd = rng.standard_normal((nd, 6))
    Bd[d == 3.0, 1] = 0.0
    Bd[:, 2] = 0.0
    Bd[d != 2.0, 4] = 0.0
    td = [0.1, 0.2, 0.3, 0.4, 0.5, 0.6]
    try:
        eng.set_flag("no_mv", 2)
        Wd = gpu.expv_batched(td, Dg, Bd, m=30)
        assert eng.last_kernel() == "tma_mv"
    finally:
        eng.set_flag("no_mv", 0)
    for i, t in enumerate(td):
        ref = np.exp(t * d) * Bd[:, i]
        assert np.linalg.norm(Wd[:, i] - ref) <= 1e-9 * max(np.linalg.norm(ref), 1e-300), i
    for m in (30, 7):
        W, G = out[(False, m)], out[(True, m)]
        assert np.linalg.norm(W[:, 5]) == 0.0
        for i in range(nb):
            if i != 5:
                assert relerr(W[:, i], G[:, i]) < 1e-11, (m, i)
        for i in (0, 4, 8):
            assert relerr(W[:, i], oracle.expv(ts[i], L, B[:, i], m=m)) < RTOL


def test_device_small_exp_branches(gpu, oracle):
    """Fused expv (device Pade) across the Pade orders: scale t so that ||tH|| hits C3..C13 and several squarings."""
    A = convdiff2d(60, 50)
    b = np.random.default_rng(2).standard_normal(3000)
    op = gpu.operator(A)
    for t in (1e-4, 2e-3, 2e-2, 0.1, 0.25, 1.0, 6.0):
        assert relerr(gpu.expv(t, op, b, m=20), oracle.expv(t, A, b, m=20)) < RTOL, t
    # m above the device limit (48) takes the host path transparently
    assert relerr(gpu.expv(0.5, op, b, m=60), oracle.expv(0.5, A, b, m=60)) < RTOL
    # badly scaled operator: D A D^-1 exercises the balancing sweeps of the device kernel
    import scipy.sparse as sp
    d = 2.0 ** np.random.default_rng(3).integers(-6, 7, 3000)
    As = (sp.diags(d) @ A @ sp.diags(1.0 / d)).tocsr()
    assert relerr(gpu.expv(0.3, As, b, m=20), oracle.expv(0.3, As, b, m=20)) < 1e-8


def test_phiv_timestep_and_expv_timestep(gpu, oracle):
    """phiv_timestep! / expv_timestep! (src/krylov_phiv_adaptive.jl): the reference's own assertions
    (test/basictests.jl:666-691) and step-for-step parity with the oracle's controller."""
    n, K, t, tol = 100, 4, 5.0, 1e-7
    A = sp.diags([np.ones(n - 1), -2 * np.ones(n), np.ones(n - 1)], [-1, 0, 1]).tocsr()
    B = np.random.default_rng(14).standard_normal((n, K + 1))
    U, ns = gpu.phiv_timestep([t / 2, t], A, B, adaptive=True, tol=tol, return_steps=True)
    Uo, nso = oracle.phiv_timestep([t / 2, t], A, B, adaptive=True, tol=tol, return_steps=True)
    assert ns == nso and relerr(U, Uo) < 1e-8
    ue = sla.expm(t * A.toarray()) @ B[:, 0]
    opn = abs(A).sum(axis=1).max()
    for on in (None, opn, lambda A_, p_: abs(A_).sum(axis=1).max()):
        u, ns = gpu.expv_timestep(t, A, B[:, 0], adaptive=True, tol=tol, opnorm=on, return_steps=True)
        uo, nso = oracle.expv_timestep(t, A, B[:, 0], adaptive=True, tol=tol, opnorm=on, return_steps=True)
        assert ns == nso and relerr(u, uo) < 1e-8
        assert relerr(u, ue) < (1e-5 if on is None else tol)
    # larger non-symmetric operator, non-adaptive and adaptive, several snapshots, correct=True
    A2 = convdiff2d(60, 50)
    B2 = np.random.default_rng(15).standard_normal((3000, 3))
    for kw in ({}, {"adaptive": True}, {"adaptive": True, "correct": True, "m": 15}):
        U, ns = gpu.phiv_timestep([0.3, 0.7, 1.0], A2, B2, tol=1e-8, return_steps=True, **kw)
        Uo, nso = oracle.phiv_timestep([0.3, 0.7, 1.0], A2, B2, tol=1e-8, return_steps=True, **kw)
        assert ns == nso, (kw, ns, nso)
        assert relerr(U, Uo) < 1e-8, kw


def test_error_estimate_mode(gpu, oracle):
    """expv(...; mode = :error_estimate): same stopping index and result as the restated reference."""
    A = laplacian2d(60, 50)
    b = np.random.default_rng(9).standard_normal(3000)
    for t, tol in ((0.05, 1e-7), (0.5, 1e-7), (1.0, 1e-10)):
        w, mm = gpu.expv(t, A, b, mode="error_estimate", m=30, tol=tol, return_m=True)
        wo, mo = oracle.expv_ee(t, A, b, m=30, tol=tol, return_m=True)
        assert mm == mo and relerr(w, wo) < RTOL, (t, tol, mm, mo)
    assert gpu.expv(0.05, A, b, mode="error_estimate", return_m=True)[1] < 30      # it does stop early
    assert np.linalg.norm(gpu.expv(0.5, A, np.zeros(3000), mode="error_estimate")) == 0.0
    with pytest.raises(gpu.UnsupportedError):
        gpu.expv(0.5, convdiff2d(60, 50), b, mode="error_estimate")


def test_edge_layouts_and_breakdown_in_stream_mode(gpu, oracle):
    """Odd sizes (LDG kernel), tiny CSR operators, empty rows, and a happy breakdown while the TMA producer is
    running ahead (CSR stream mode)."""
    rng = np.random.default_rng(31)
    eng = gpu.get_engine()
    # dense with odd n -> unaligned columns -> LDG kernel
    D = rng.standard_normal((301, 301)) / 10
    b = rng.standard_normal(301)
    assert relerr(gpu.expv(1.0, D, b, m=30), oracle.expv(1.0, D, b, m=30)) < RTOL
    assert eng.last_kernel() == "ldg"
    # CSR with n smaller than one CTA slice, and one with empty rows
    T = sp.diags([np.ones(9), -2 * np.ones(10), np.ones(9)], [-1, 0, 1]).tocsr()
    b10 = rng.standard_normal(10)
    assert relerr(gpu.expv(0.5, T, b10, m=10), oracle.expv(0.5, T, b10, m=10)) < RTOL
    E = sp.random(400, 400, density=0.01, random_state=5, format="lil")
    E[7, :] = 0
    E[123, :] = 0
    E = (E.tocsr() - 2 * sp.identity(400)).tolil()
    E[7, :] = 0      # truly empty rows (no diagonal either)
    E[123, :] = 0
    E = E.tocsr()
    E.eliminate_zeros()
    b400 = rng.standard_normal(400)
    assert relerr(gpu.expv(0.7, E, b400, m=25), oracle.expv(0.7, E, b400, m=25)) < RTOL
    # breakdown in CSR stream mode: a diagonal operator with 3 distinct eigenvalues has a 3-dimensional Krylov space
    n = 6000
    d = np.array([1.0, 2.0, 3.0])[np.arange(n) % 3]
    Dg = sp.diags(d).tocsr()
    bn = rng.standard_normal(n)
    for herm in (True, False):
        Ks = gpu.arnoldi(Dg, bn, m=30, ishermitian=herm)
        Ko = oracle.arnoldi(Dg, bn, m=30, ishermitian_=herm)
        assert eng.last_kernel() == ("tma_xl" if herm else "tma")
        assert Ks.m == Ko.m == 3 and Ks.wasbreakdown
        w = gpu.expv(0.9, Dg, bn, m=30, ishermitian=herm)
        assert relerr(w, np.exp(0.9 * d) * bn) < 1e-9
    # a batch in which some problems break down and some do not
    B = rng.standard_normal((n, 5))
    W = gpu.expv_batched([0.1, 0.2, 0.3, 0.4, 0.5], Dg, B, m=30)
    for i, t in enumerate([0.1, 0.2, 0.3, 0.4, 0.5]):
        assert relerr(W[:, i], np.exp(t * d) * B[:, i]) < 1e-9
    # large Krylov dimension (m = 100), full Arnoldi
    A = convdiff2d(50, 40)
    b2 = rng.standard_normal(2000)
    assert relerr(gpu.expv(2.0, A, b2, m=100), oracle.expv(2.0, A, b2, m=100)) < 1e-9


def test_batched_device_exponential(gpu, oracle):
    """b200k_exponential_batched: many small matrices on the device, every Pade branch, against
    the oracle's exponential! and scipy.linalg.expm (test/basictests.jl:952-974 style)."""
    rng = np.random.default_rng(41)
    mats = []
    for scale in (3.0, 1.5, 0.5, 0.1, 0.005, 40.0, 0.0):
        for n in (1, 2, 7, 34, 48):
            A = rng.standard_normal((n, n))
            nrm = np.linalg.norm(A, 1)
            mats.append((n, A * (scale / nrm if nrm > 0 else 0.0)))
    for n in (1, 2, 7, 34, 48):
        batch = np.stack([A for (k, A) in mats if k == n])
        E = gpu.exponential_batched_(batch)
        for b in range(batch.shape[0]):
            ref = oracle.exponential_higham2005base(batch[b])
            assert relerr(E[b], ref) < 1e-11, (n, b)
            assert relerr(E[b], sla.expm(batch[b])) < 1e-10, (n, b)
    # badly scaled matrices exercise the balancing sweeps
    A = rng.standard_normal((12, 12))
    D = np.diag(2.0 ** rng.integers(-15, 15, 12))
    Bs = np.stack([D @ A @ np.linalg.inv(D), A])
    E = gpu.exponential_batched_(Bs)
    assert relerr(E[0], oracle.exponential_higham2005base(Bs[0])) < 1e-9
    with pytest.raises(gpu.UnsupportedError):
        gpu.exponential_batched_(np.zeros((2, 49, 49)))


def test_complex_values(gpu, oracle):
    """test/basictests.jl:650-664 "Complex Value": {Hermitian complex, Hermitian real, general complex, general real}
    x {complex b, real b} x t in {real, imaginary, complex}, n = 20, m = 10 -- against dense exp and the oracle."""
    rng = np.random.default_rng(3)
    n, m = 20, 10
    X = rng.random((n, n)) + 1j * rng.random((n, n))
    mats = {"herm_c": X + X.conj().T, "herm_r": X.real + X.real.T, "gen_c": X, "gen_r": X.real}
    for name, A in mats.items():
        for b in (rng.random(n) + 1j * rng.random(n), rng.random(n)):
            for t in (1e-2, 1e-2j, 1e-2 + 1e-2j):
                w = gpu.expv(t, A, b, m=m)
                assert relerr_c(w, sla.expm(t * A) @ b) < 1.5e-8, (name, b.dtype, t)
                assert relerr_c(w, oracle.expv(t, A, b, m=m)) < RTOL, (name, b.dtype, t)


def relerr_c(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def test_complex_sparse_and_factorisation(gpu, oracle):
    """The reference's GPU test shape (test/gpu/gputests.jl:41-58): complex sparse operator, expv vs the CPU path;
    plus H / V parity of the complex Arnoldi and Hermitian-Lanczos factorisations and a Schroedinger-type case."""
    rng = np.random.default_rng(5)
    n = 1000
    A = (sp.random(n, n, density=0.01, random_state=1) + 1j * sp.random(n, n, density=0.01, random_state=2)).tocsr()
    A = (sp.triu(A) + 0.1 * sp.identity(n)).tocsr()
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    assert relerr_c(gpu.expv(0.3, A, b, m=30), oracle.expv(0.3, A, b, m=30)) < RTOL
    assert relerr_c(gpu.expv(0.3 - 0.2j, A, b, m=30), oracle.expv(0.3 - 0.2j, A, b, m=30)) < RTOL
    Ks = gpu.arnoldi(A, b, m=15)
    Ko = oracle.arnoldi(A, b, m=15)
    assert Ks.is_complex and Ks.m == Ko.m and abs(Ks.beta - Ko.beta) < 1e-12
    assert np.abs(Ks.getH() - Ko.getH()).max() < 1e-11
    assert np.abs(Ks.getV().cpu().numpy() - Ko.getV()).max() < 1e-11
    # Hermitian complex operator: -i * (Laplacian + complex Hermitian coupling), Lanczos with real coefficients
    L = laplacian2d(40, 30)
    Cc = sp.diags([1j * np.ones(1199), -1j * np.ones(1199)], [1, -1])
    Hm = (L + 0.5 * Cc).tocsr()
    psi = rng.standard_normal(1200) + 1j * rng.standard_normal(1200)
    opH = gpu.operator(Hm)
    assert opH.is_complex and opH.ishermitian
    Ks = gpu.arnoldi(opH, psi, m=20)
    Ko = oracle.arnoldi(Hm, psi, m=20)
    assert np.abs(Ks.getH().imag).max() == 0.0 and np.abs(Ks.getH().real - Ko.getH()).max() < 1e-11
    w = gpu.expv(-0.2j, opH, psi, m=30)   # unitary propagation
    assert relerr_c(w, oracle.expv(-0.2j, Hm, psi, m=30)) < RTOL
    assert abs(np.linalg.norm(w) - np.linalg.norm(psi)) < 1e-8 * np.linalg.norm(psi)
    # zero vector and happy breakdown
    assert np.linalg.norm(gpu.expv(0.1, A, np.zeros(n, dtype=complex), m=10)) == 0.0
    v = rng.standard_normal(20) + 1j * rng.standard_normal(20)
    v /= np.linalg.norm(v)
    Kb = gpu.arnoldi(np.outer(v, v.conj()), rng.standard_normal(20) + 0j)
    assert Kb.m == 2 and Kb.wasbreakdown


def test_abi_status_codes(gpu):
    """Error behaviour at the C ABI itself (status codes + last_error), bypassing the Python argument checks."""
    import ctypes as C
    import torch
    from importlib import import_module
    L = import_module("eu_b200._lib")
    lib, eng = gpu.load(), gpu.get_engine()
    op = gpu.operator(laplacian2d(10, 10))
    opts = L.KrylovOpts()
    lib.b200k_krylov_opts_default(C.byref(opts))
    assert (opts.m, opts.tol, opts.iop, opts.hermitian, opts.init, opts.p) == (30, 1e-7, 0, -1, 0, 0)
    V = torch.zeros((11, 112), dtype=torch.float64, device="cuda")
    H = np.zeros((11, 10), order="F")
    b = torch.ones(100, dtype=torch.float64, device="cuda")
    beta, mo, bd = C.c_double(), C.c_int(), C.c_int()

    def call(m, ldv, maxiter, ldh):
        opts.m = m
        return lib.b200k_arnoldi(eng.handle, op.ptr, C.c_void_p(b.data_ptr()), C.byref(opts), C.c_void_p(V.data_ptr()),
                                 ldv, maxiter, H.ctypes.data_as(L.c_double_p), ldh, C.byref(beta), C.byref(mo), C.byref(bd))

    assert call(10, 112, 10, 11) == L.OK and mo.value == 10
    assert call(11, 112, 10, 11) == L.EDIM and b"maxiter" in lib.b200k_last_error(eng.handle)      # resize! is the host's job
    assert call(10, 99, 10, 11) == L.EDIM                                                           # size(V,1) < size(A,1)
    assert call(10, 112, 10, 5) == L.EDIM                                                           # H too small
    assert call(0, 112, 10, 11) == L.EARG
    assert lib.b200k_arnoldi(eng.handle, op.ptr, None, C.byref(opts), None, 0, 0, None, 0, None, None, None) == L.EARG
    opts.m, opts.p = 10, 17
    assert call(10, 128, 10, 11) == L.EUNSUPPORTED                                                  # p > 16
    opts.p = 0
    # a complex operator through a real entry point (and vice versa) is an argument error
    opz = gpu.operator(laplacian2d(10, 10).astype(np.complex128))
    assert lib.b200k_arnoldi(eng.handle, opz.ptr, C.c_void_p(b.data_ptr()), C.byref(opts), C.c_void_p(V.data_ptr()), 112,
                             10, H.ctypes.data_as(L.c_double_p), 11, C.byref(beta), C.byref(mo), C.byref(bd)) == L.EARG
    ko = L.KiopsOpts()
    lib.b200k_kiops_opts_default(C.byref(ko))
    assert (ko.mmin, ko.mmax, ko.m, ko.tol, ko.iop) == (10, 128, 10, 1e-7, 2)
    # kiops with several output times is a DimensionMismatch in the reference (checkdims) and here
    tau = np.array([0.5, 1.0])
    W = torch.zeros((2, 100), dtype=torch.float64, device="cuda")
    stats = (C.c_int64 * 5)()
    assert lib.b200k_kiops(eng.handle, op.ptr, 2, tau.ctypes.data_as(L.c_double_p), 1, C.c_void_p(b.data_ptr()), 100, 1,
                           C.byref(ko), C.c_void_p(W.data_ptr()), 100, stats) == L.EDIM
    assert lib.b200k_status_string(L.EDIM) == b"dimension mismatch"
