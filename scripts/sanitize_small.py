"""Small-n run of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):
    compute-sanitizer --tool racecheck python scripts/sanitize_small.py
Sizes are tiny (cooperative persistent kernels run ~100x slower under the sanitizer); every result is still checked
against the CPU oracle so that a tool-induced timing change cannot hide a wrong answer."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import eu_b200 as eu
from conftest import laplacian2d, convdiff2d, relerr
from oracle import oracle as O

which = set(sys.argv[1:]) or {"arnoldi", "xl", "iop", "mv", "dense", "reorth", "z", "kiops", "phiv", "ldg"}
rng = np.random.default_rng(0)
L = laplacian2d(96, 80); C = convdiff2d(96, 80); n = L.shape[0]
b = rng.standard_normal(n)
eng = eu.get_engine()
def chk(name, a, r):
    e = relerr(a, r); print(f"{name}: relerr {e:.2e} kernel {eng.last_kernel()}", flush=True); assert e < 1e-10, name
if "arnoldi" in which: chk("arnoldi tma", eu.expv(1.0, C, b, m=12), O.expv(1.0, C, b, m=12))
if "xl" in which: chk("lanczos xl", eu.expv(1.0, L, b, m=12), O.expv(1.0, L, b, m=12))
if "iop" in which: chk("iop2 xl", eu.expv(1.0, C, b, m=12, iop=2), O.expv(1.0, C, b, m=12, iop=2))
if "mv" in which:
    B = rng.standard_normal((n, 6)); ts = rng.uniform(0.1, 1, 6)
    for flag in (0, 2):
        eng.set_flag("no_mv", flag)
        W = eu.expv_batched(ts, L, B, m=10)
        chk(f"batched lanczos mv={flag}", W[:, 4], O.expv(ts[4], L, B[:, 4], m=10))
    eng.set_flag("no_mv", 0)
    W = eu.expv_batched(ts, C, B, m=10)
    chk("batched arnoldi", W[:, 5], O.expv(ts[5], C, B[:, 5], m=10))
if "dense" in which:
    A = rng.standard_normal((256, 256)) / 16; bd = rng.standard_normal(256)
    chk("dense tma", eu.expv(1.0, A, bd, m=10), O.expv(1.0, A, bd, m=10))
if "reorth" in which:
    import scipy.sparse as sp
    d = np.concatenate([-np.ones(1000), -2 * np.ones(500), -1e3 * np.ones(500)]) + 1e-8 * rng.standard_normal(2000)
    D = sp.diags(d).tocsr(); bb = rng.standard_normal(2000)
    chk("reorth csr", eu.expv(0.01, D, bb, m=20, ishermitian=False), O.expv(0.01, D, bb, m=20, ishermitian_=False))
if "z" in which:
    Az = (L.astype(np.complex128) + 0.3j * C).tocsr(); bz = b + 1j * rng.standard_normal(n)
    chk("complex arnoldi", eu.expv(0.5, Az, bz, m=10), O.expv(0.5, Az, bz, m=10))
    import scipy.sparse as sp
    Hz = (L + 0.5 * sp.diags([1j * np.ones(n - 1), -1j * np.ones(n - 1)], [1, -1])).tocsr()
    chk("complex hermitian lanczos", eu.expv(-0.3j, Hz, bz, m=10), O.expv(-0.3j, Hz, bz, m=10))
    dz = np.concatenate([-np.ones(1000), -2 * np.ones(500), -1e3 * np.ones(500)]) + 1e-8 * rng.standard_normal(2000)
    Dz = (sp.diags(dz).astype(np.complex128) + 1e-3j * sp.diags(rng.standard_normal(2000))).tocsr()
    bzz = rng.standard_normal(2000) + 1j * rng.standard_normal(2000)
    chk("complex reorth hand-over", eu.expv(0.01, Dz, bzz, m=20), O.expv(0.01, Dz, bzz, m=20))
    eng.set_flag("force_ldg", 1)
    chk("complex arnoldi ldg", eu.expv(0.5, Az, bz, m=10), O.expv(0.5, Az, bz, m=10))
    eng.set_flag("force_ldg", 0)
if "kiops" in which:
    u = rng.standard_normal((n, 2))
    w, st = eu.kiops(0.5, C, u); wo, so = O.kiops(0.5, C, u)
    chk("kiops", w, wo); assert st == so
if "phiv" in which:
    chk("phiv", eu.phiv(0.5, C, b, 3, m=10, correct=True), O.phiv(0.5, C, b, 3, m=10, correct=True))
if "ldg" in which:
    Lo = laplacian2d(31, 33); bo = rng.standard_normal(31 * 33)  # odd n -> LDG kernel
    chk("ldg odd n", eu.expv(1.0, Lo, bo, m=10, ishermitian=False), O.expv(1.0, Lo, bo, m=10, ishermitian_=False))
torch.cuda.synchronize()
print("SANITIZE_SMALL_OK")
