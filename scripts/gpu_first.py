"""First-light script: a handful of parity checks with verbose output (run under gpurun)."""
import sys, os, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, scipy.sparse as sp, torch
import eu_b200 as eu
from oracle import oracle as O
from conftest import laplacian2d, convdiff2d, relerr

def run(name, fn):
    t0 = time.time()
    try:
        out = fn()
        print(f"[{name}] {out}  ({time.time()-t0:.2f}s)", flush=True)
    except Exception as e:
        print(f"[{name}] FAILED: {type(e).__name__}: {e}", flush=True)
        traceback.print_exc()

rng = np.random.default_rng(0)
small = len(sys.argv) > 1 and sys.argv[1] == "small"

def t_dense_small():
    n, m = 20, 5
    A = rng.standard_normal((n, n)); b = rng.standard_normal(n); t = 1e-2
    return relerr(eu.expv(t, A, b, m=m), O.expv(t, A, b, m=m))
run("dense n=20 arnoldi", t_dense_small)

def t_dense_512():
    n = 512
    A = np.random.default_rng(0).standard_normal((n, n)) / np.sqrt(n); b = np.random.default_rng(1).standard_normal(n)
    return relerr(eu.expv(1.0, A, b, m=30), O.expv(1.0, A, b, m=30))
run("C1 dense 512 arnoldi", t_dense_512)

def t_csr(nx, ny, herm, nonsym=False):
    A = convdiff2d(nx, ny) if nonsym else laplacian2d(nx, ny)
    b = np.random.default_rng(0).standard_normal(nx * ny)
    w = eu.expv(1.0, A, b, m=30, ishermitian=herm)
    wo = O.expv(1.0, A, b, m=30, ishermitian_=herm)
    return relerr(w, wo)
run("csr 40x50 lanczos", lambda: t_csr(40, 50, True))
run("csr 40x50 arnoldi", lambda: t_csr(40, 50, False))
run("csr 40x50 convdiff arnoldi", lambda: t_csr(40, 50, None, True))
run("csr 37x41 (odd n) arnoldi", lambda: t_csr(37, 41, False))
if not small:
    run("csr 300x400 arnoldi", lambda: t_csr(300, 400, False))
    run("csr 300x400 lanczos", lambda: t_csr(300, 400, True))
    run("csr 1000x1000 arnoldi (C2)", lambda: t_csr(1000, 1000, False))
    run("csr 1000x1000 lanczos (C2)", lambda: t_csr(1000, 1000, True))

def t_H():
    A = convdiff2d(40, 50); b = np.random.default_rng(3).standard_normal(2000)
    Ks = eu.arnoldi(A, b, m=20); Ko = O.arnoldi(A, b, m=20)
    return (Ks.m, Ko.m, abs(Ks.beta - Ko.beta), np.abs(Ks.getH() - Ko.getH()).max(),
            np.abs(Ks.getV().cpu().numpy() - Ko.getV()).max())
run("H/V parity convdiff m=20", t_H)

def t_phiv():
    A = convdiff2d(40, 50); b = np.random.default_rng(3).standard_normal(2000)
    W, e = eu.phiv(0.5, A, b, 4, m=20, correct=True, errest=True)
    Wo, eo = O.phiv(0.5, A, b, 4, m=20, correct=True, errest=True)
    return relerr(W, Wo), abs(e - eo)
run("phiv k=4 correct", t_phiv)

def t_break():
    n = 20; v = rng.standard_normal(n); v /= np.linalg.norm(v); b = rng.standard_normal(n)
    Ks = eu.arnoldi(np.outer(v, v), b)
    z = eu.expv(1e-2, np.outer(v, v), np.zeros(n), m=5)
    return Ks.m, Ks.wasbreakdown, float(np.linalg.norm(z))
run("breakdown + zero", t_break)

def t_kiops():
    n = 20; A = np.random.default_rng(5).standard_normal((n, n)); b = np.random.default_rng(6).standard_normal(n); t = 1e-2
    w, st = eu.kiops(t, A, b); wo, so = O.kiops(t, A, b)
    U = np.stack([b * (1 / t) ** i for i in range(4)], 1)
    w3, st3 = eu.kiops(t, A, U); wo3, so3 = O.kiops(t, A, U)
    return relerr(w, wo), st, so, relerr(w3, wo3), st3, so3
run("kiops small", t_kiops)

def t_kiops_lap():
    A = laplacian2d(40, 50); u = np.random.default_rng(4).standard_normal((2000, 2))
    out = []
    for herm in (True, False):
        w, st = eu.kiops(1.0, A, u, ishermitian=herm); wo, so = O.kiops(1.0, A, u, ishermitian_=herm)
        out.append((relerr(w, wo), st, so))
    return out
run("kiops laplacian p=1", t_kiops_lap)

def t_batched():
    A = laplacian2d(40, 50); n = 2000; nb = 37
    B = np.random.default_rng(6).standard_normal((n, nb)); ts = np.random.default_rng(7).uniform(0.1, 1, nb)
    out = []
    for herm in (True, False):
        W = eu.expv_batched(ts, A, B, m=30, ishermitian=herm)
        err = max(relerr(W[:, i], O.expv(ts[i], A, B[:, i], m=30, ishermitian_=herm)) for i in range(nb))
        out.append(err)
    return out
run("batched expv", t_batched)
print(eu.get_engine().device_info())
