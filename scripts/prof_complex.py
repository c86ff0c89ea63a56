"""ComplexF64 expv at the headline size for profiler captures: three general (Arnoldi) launches of krylov_tma_z_kernel,
then three Hermitian (Lanczos) ones.
    ncu --set full --clock-control none --import-source on -k regex:krylov_tma_z_kernel -s 1 -c 1 -o out python scripts/prof_complex.py
(-s 4 for the Hermitian case)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, scipy.sparse as sp
import eu_b200 as eu
from conftest import laplacian2d

n = 10**6
L = laplacian2d(1000, 1000)
Az = (L.astype(np.complex128) + sp.diags([0.3j * np.ones(n - 1), 0.3j * np.ones(n - 1)], [1, -1])).tocsr()
Hs = (-1.0 * L).astype(np.complex128).tocsr()
rng = np.random.default_rng(12)
psi = torch.from_numpy(rng.standard_normal(n) + 1j * rng.standard_normal(n)).cuda()
for M, t in ((Az, 0.5), (Hs, -0.5j)):
    op = eu.operator(M)
    for _ in range(3):
        eu.expv(t, op, psi, m=30)
    torch.cuda.synchronize()
    print(eu.get_engine().last_kernel(), flush=True)
