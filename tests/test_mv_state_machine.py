"""CPU model of the lock-step multi-vector Lanczos kernel (csrc/krylov_kernel_mv.cuh): the per-problem state machine
(dead / breakdown / padding problems inside one group, folded beta*v_{j-1} term, basis columns stored one step late,
epilogue column) restated in NumPy and checked against the oracle's lanczos! per problem.  It documents the algorithm
the kernel implements and guards its logic when the kernel is tuned; the GPU test
(test_gpu_parity.py::test_batched_multivector_lanczos) checks the kernel itself."""
import numpy as np
import scipy.sparse as sp

from conftest import laplacian2d

KV = 4


def mv_group_model(A, B, m, tol=1e-7):
    """One group of KV problems (columns of B; all-zero columns allowed; missing columns = padding) advanced in lock
    step exactly as mv_consumer does.  Returns per problem (V, H, beta0, m_out, breakdown)."""
    n, nvalid = B.shape
    valid = [v < nvalid for v in range(KV)]
    xin = np.zeros((n, KV))
    for v in range(KV):
        if valid[v]:
            xin[:, v] = B[:, v]
    ws = np.zeros((n, KV))                      # contents irrelevant before the first fold
    V = [np.full((n, m + 1), np.nan) for _ in range(KV)]
    H = [np.zeros((m + 1, m)) for _ in range(KV)]
    beta = np.sqrt((xin * xin).sum(0))
    dead = [(not valid[v]) or beta[v] == 0.0 for v in range(KV)]
    run = [not d for d in dead]
    xscale = np.array([0.0 if dead[v] else 1.0 / beta[v] for v in range(KV)])
    vscale = xscale.copy()
    beta0 = beta.copy()
    beta_prev = np.zeros(KV)
    xscale_prev = np.zeros(KV)
    m_out = [m] * KV
    brk = [0] * KV
    jlast = 0
    j = 1
    while j <= m and not all(dead):
        jlast = j
        jc = j - 1
        foldc = beta_prev * xscale_prev
        # mat-vec of the KV problems, inner product with v_j before the fold (arnoldi.jl:396-399)
        wv = (A @ xin) * xscale
        alpha = (xin * wv).sum(0) * xscale
        if j > 1:
            wv = wv - foldc * ws                # ws still holds the unnormalised v_{j-1}
        ws = wv
        for v in range(KV):
            if not dead[v]:
                H[v][jc, jc] = alpha[v]
        ws = ws - (alpha * xscale) * xin        # update
        nrm2 = (ws * ws).sum(0)
        for v in range(KV):                     # column jc goes to V one step late
            if run[v] and jc <= m_out[v]:
                V[v][:, jc] = xin[:, v] * vscale[v]
        for v in range(KV):
            bt = np.sqrt(nrm2[v])
            if not dead[v]:
                H[v][jc + 1, jc] = bt
                beta_prev[v], xscale_prev[v] = bt, xscale[v]
                with np.errstate(divide="ignore"):
                    vscale[v] = 1.0 / bt
                beta[v] = bt
                if bt < tol:
                    m_out[v], brk[v], dead[v] = j, 1, True
                    xscale[v] = beta_prev[v] = xscale_prev[v] = 0.0
                else:
                    xscale[v] = vscale[v]
            else:
                beta_prev[v] = xscale_prev[v] = 0.0
        xin, ws = ws, xin                       # swap
        j += 1
    for v in range(KV):                         # epilogue
        if run[v] and jlast > 0 and (m_out[v] if dead[v] else jlast) == jlast:
            with np.errstate(divide="ignore", invalid="ignore"):
                V[v][:, jlast] = xin[:, v] / beta[v]
    out = []
    for v in range(nvalid):
        Hm = H[v].copy()
        for i in range(m - 1):                  # lanczos! mirrors the sub-diagonal (arnoldi.jl:488)
            Hm[i, i + 1] = Hm[i + 1, i]
        out.append((V[v], Hm, beta0[v], m_out[v], brk[v]))
    return out


def check_against_oracle(oracle, A, B, m):
    res = mv_group_model(A, B, m)
    for v, (V, H, b0, mo, bd) in enumerate(res):
        Ks = oracle.KrylovSubspace(A.shape[0], m)
        oracle.lanczos_(Ks, A, B[:, v], m=m)
        assert abs(b0 - Ks.beta) <= 1e-14 * max(Ks.beta, 1.0)
        if Ks.beta == 0:
            assert (mo, bd) == (m, 0)
            continue
        assert mo == Ks.m and bool(bd) == Ks.wasbreakdown, (v, mo, Ks.m)
        Ho, Vo = Ks.getH(), Ks.getV()
        assert np.abs(H[: mo + 1, :mo] - Ho[: mo + 1, :mo]).max() < 1e-10
        # the columns expv uses (1..m); the last one may be huge after a breakdown and is compared relatively
        assert np.abs(V[:, :mo] - Vo[:, :mo]).max() < 1e-8
        assert not np.isnan(V[:, : mo + 1]).any() or Ks.wasbreakdown


def test_model_matches_lanczos_on_a_laplacian(oracle):
    A = laplacian2d(24, 17)
    rng = np.random.default_rng(3)
    check_against_oracle(oracle, A, rng.standard_normal((A.shape[0], 4)), 30)
    B = rng.standard_normal((A.shape[0], 3))       # padding slot + a zero start vector
    B[:, 1] = 0.0
    check_against_oracle(oracle, A, B, 12)


def test_model_breakdowns_inside_one_group(oracle):
    nd = 300
    d = np.array([1.0, 2.0, 3.0])[np.arange(nd) % 3]
    Dg = sp.diags(d).tocsr()
    rng = np.random.default_rng(4)
    B = rng.standard_normal((nd, 4))
    B[d == 3.0, 1] = 0.0      # Krylov dimension 2
    B[:, 2] = 0.0             # zero vector
    B[d != 2.0, 3] = 0.0      # Krylov dimension 1
    check_against_oracle(oracle, Dg, B, 30)
    res = mv_group_model(Dg, B, 30)
    assert [r[3] for r in res] == [3, 2, 30, 1] and [r[4] for r in res] == [1, 1, 0, 1]
