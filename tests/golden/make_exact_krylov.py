"""Exact-arithmetic Krylov approximants (mpmath, 60 significant digits) -> tests/golden/exact_krylov.json.

Why: the oracle (oracle/oracle.py) is our own restatement of the reference and Julia cannot run in the build
container, so nothing ties it to reference OUTPUTS.  This script pins the ALGORITHM independently of the oracle
author's floating-point code: in exact arithmetic the Krylov approximant

        w_m = beta * V_m * exp(t * H_m) * e_1,      [phi_0 .. phi_k](t H_m) e_1 projected the same way,

is uniquely determined by (A, b, m, t) -- it does not depend on the orthogonalisation scheme (MGS in the reference,
src/arnoldi.jl:289-308; classical Gram-Schmidt in the CUDA kernel), on the Pade / eigen branch of the small
exponential (src/krylov_phiv.jl:223-244, src/exp_baseexp.jl:112-161) or on any summation order.  Both the oracle and
the CUDA path must therefore reproduce these numbers to rounding, INCLUDING cases where w_m is far from exp(tA) b
(non-converged: the comparison against a dense exp would say nothing there).

Cases
  * the reference's only seed-free fixture for this path: test/basictests.jl:859-871 (`mkA(n)`, `b = [1/i]`, m = 30,
    `expv!(w, 0.1, Ks)`, `phiv!(w, 0.1, Ks, 3)`), n = 64 and n = 200.  tests/golden/check_with_julia.jl evaluates the
    same calls with the real package so that a maintainer with Julia can compare against the committed values;
  * non-converged: unscaled randn(100,100) (NumPy seed 31), m = 10, t = 0.5 -- far from exp(tA)b;
  * symmetric (Lanczos / eigen branch in the reference): 1-D Laplacian n = 150, b = [1/i], m = 12, t = 2.0.

The small phi functions are evaluated through the augmented-matrix identity the reference itself uses
(src/phi.jl:84-115) with mpmath.expm at 60 digits.  Run from the repository root:
    python tests/golden/make_exact_krylov.py
"""
import json
import os

import mpmath as mp
import numpy as np

mp.mp.dps = 60
HERE = os.path.dirname(os.path.abspath(__file__))


def mkA(n):
    """test/basictests.jl:859-862"""
    return np.array([[-2.0 if i == j else 0.1 / (1 + abs(i - j)) * (1.0 if i < j else 0.5) for j in range(1, n + 1)]
                     for i in range(1, n + 1)])


def exact_krylov(A, b, m, t, k):
    """Arnoldi in 60-digit arithmetic; returns the n x (k+1) matrix [beta V_m phi_i(t H_m) e_1] and (H, beta)."""
    n = A.shape[0]
    Am = mp.matrix(A.tolist())
    V = [mp.matrix([mp.mpf(float(x)) for x in b])]
    beta = mp.sqrt(sum(x * x for x in V[0]))
    V[0] = V[0] / beta
    H = mp.zeros(m + 1, m)
    for j in range(m):
        w = Am * V[j]
        for i in range(j + 1):  # (the scheme does not matter at 60 digits; two passes for good measure)
            h = sum(V[i][r] * w[r] for r in range(n))
            H[i, j] += h
            w = w - h * V[i]
        for i in range(j + 1):
            h = sum(V[i][r] * w[r] for r in range(n))
            H[i, j] += h
            w = w - h * V[i]
        H[j + 1, j] = mp.sqrt(sum(x * x for x in w))
        V.append(w / H[j + 1, j])
    # [phi_0(tH) e1 ... phi_k(tH) e1] via exp of the augmented matrix (src/phi.jl:84-115)
    N = m + max(k, 1)
    C = mp.zeros(N, N)
    for i in range(m):
        for j in range(m):
            C[i, j] = t * H[i, j]
    C[0, m] = 1
    for i in range(m, N - 1):
        C[i, i + 1] = 1
    E = mp.expm(C)
    cols = [[E[i, 0] for i in range(m)]] + [[E[i, m + c - 1] for i in range(m)] for c in range(1, k + 1)]
    W = np.zeros((n, k + 1))
    for c, y in enumerate(cols):
        for r in range(n):
            W[r, c] = float(beta * sum(V[i][r] * y[i] for i in range(m)))
    Hf = np.array([[float(H[i, j]) for j in range(m)] for i in range(m + 1)])
    return W, Hf, float(beta)


def main():
    out = {"_doc": "exact-arithmetic Krylov approximants, see tests/golden/make_exact_krylov.py", "cases": {}}
    specs = []
    for n in (64, 200):
        specs.append((f"mkA_{n}", mkA(n), np.array([1.0 / i for i in range(1, n + 1)]), 30, 0.1, 3, False))
    rng = np.random.default_rng(31)
    specs.append(("randn100_m10_nonconverged", rng.standard_normal((100, 100)), rng.standard_normal(100), 10, 0.5, 2, False))
    n = 150
    L = (np.diag(-2.0 * np.ones(n)) + np.diag(np.ones(n - 1), 1) + np.diag(np.ones(n - 1), -1))
    specs.append(("lap1d_150_m12_symmetric", L, np.array([1.0 / i for i in range(1, n + 1)]), 12, 2.0, 2, True))
    for name, A, b, m, t, k, sym in specs:
        W, H, beta = exact_krylov(A, b, m, t, k)
        import scipy.linalg as sla
        dense = sla.expm(t * A) @ b
        out["cases"][name] = {
            "n": int(A.shape[0]), "m": m, "t": t, "k": k, "symmetric": sym, "beta": beta,
            "W": W.tolist(), "H": H.tolist(),
            "rel_dist_to_dense_exp": float(np.linalg.norm(W[:, 0] - dense) / np.linalg.norm(dense)),
        }
        if name.startswith("mkA_"):  # plain CSV for tests/golden/check_with_julia.jl (readdlm)
            np.savetxt(os.path.join(HERE, f"exact_{name}_W.csv"), W, delimiter=",", fmt="%.17e")
        print(name, "distance of the Krylov approximant from exp(tA)b:", out["cases"][name]["rel_dist_to_dense_exp"])
    with open(os.path.join(HERE, "exact_krylov.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
