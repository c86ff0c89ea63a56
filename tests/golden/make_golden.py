"""Regenerates tests/golden/krylov_golden.npz from the CPU oracle (oracle/oracle.py).

The reference is Julia and cannot run in the build container, and it ships no golden vectors for this path, so
these fixtures pin the ORACLE (against accidental change) and give the GPU tests fixed expected outputs; they do
not pin the oracle to reference outputs.  Run from the repository root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import convdiff2d, laplacian2d  # noqa: E402
from oracle import oracle as O  # noqa: E402


def cases():
    rng = np.random.default_rng(2026)
    out = {}
    L = laplacian2d(12, 10)
    C = convdiff2d(12, 10)
    b = rng.standard_normal(120)
    out["b"] = b
    out["expv_lanczos_t1"] = O.expv(1.0, L, b, m=20)
    out["expv_arnoldi_t1"] = O.expv(1.0, C, b, m=20)
    out["expv_iop2_t05"] = O.expv(0.5, C, b, m=20, iop=2)
    Ks = O.arnoldi(C, b, m=12)
    out["arnoldi_H"] = np.array(Ks.getH())
    out["arnoldi_beta"] = np.array([Ks.beta])
    out["phiv_k3_correct"] = O.phiv(0.7, C, b, 3, m=12, correct=True)
    u = rng.standard_normal((120, 3))
    out["u"] = u
    w, st = O.kiops(0.8, C, u)
    out["kiops_w"] = w
    out["kiops_stats"] = np.array(st)
    w, st = O.kiops(0.8, L, u[:, :2], ishermitian_=True)
    out["kiops_herm_w"] = w
    out["kiops_herm_stats"] = np.array(st)
    U, ns = O.phiv_timestep([0.4, 1.0], C, u, adaptive=True, tol=1e-8, return_steps=True)
    out["timestep_U"] = U
    out["timestep_nsteps"] = np.array([ns])
    w, mm = O.expv_ee(0.3, L, b, m=20, tol=1e-9, return_m=True)
    out["ee_w"] = w
    out["ee_m"] = np.array([mm])
    D = rng.standard_normal((40, 40)) / 3
    out["dense_A"] = D
    out["dense_b"] = rng.standard_normal(40)
    out["dense_expv"] = O.expv(1.0, D, out["dense_b"], m=25)
    H = np.triu(rng.standard_normal((9, 9)), -1)
    out["small_H"] = H
    out["small_exp"] = O.exponential_higham2005base(0.9 * H)
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "krylov_golden.npz"), **cases())
    print("written")
