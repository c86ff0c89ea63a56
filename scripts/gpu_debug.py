import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, scipy.sparse as sp
import eu_b200 as eu
from oracle import oracle as O
from conftest import laplacian2d, relerr
rng = np.random.default_rng(21)
def case(name, A, m=25, t=0.5):
    n = A.shape[0]
    x = np.random.default_rng(1).standard_normal(n)
    op = eu.operator(A)
    Ks = eu.arnoldi(op, x, m=m, ishermitian=False); Ko = O.arnoldi(A, x, m=m, ishermitian_=False)
    dH = np.abs(Ks.getH() - Ko.getH()).max(axis=0)
    print(name, "n", n, "relerr", relerr(eu.expv(t, op, x, m=m, ishermitian=False), O.expv(t, A, x, m=m, ishermitian_=False)),
          "first bad H col", int(np.argmax(dH > 1e-9)) if (dH > 1e-9).any() else -1, "maxdH", dH.max())
case("lap 60x50 stream, zero-row CTAs", laplacian2d(60, 50))
case("dense-rows n=2368 warp, no zero-row CTAs", (sp.random(2368, 2368, density=0.05, random_state=3, format='csr') - 6 * sp.identity(2368)).tocsr())
case("dense-rows n=3000 warp, zero-row CTAs", (sp.random(3000, 3000, density=0.05, random_state=3, format='csr') - 6 * sp.identity(3000)).tocsr())
case("dense-rows n=3000 scaled 0.01", (0.01 * sp.random(3000, 3000, density=0.05, random_state=3, format='csr') - 6 * sp.identity(3000)).tocsr())
case("dense-rows n=3000 m=5", (sp.random(3000, 3000, density=0.05, random_state=3, format='csr') - 6 * sp.identity(3000)).tocsr(), m=5)
case("dense matrix n=3000", rng.standard_normal((3000, 3000)) / 50)
