import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch, eu_b200 as eu
from conftest import laplacian2d
A = laplacian2d(1000, 1000); op = eu.operator(A); eng = op.engine
b = torch.randn(10**6, dtype=torch.float64, device="cuda")
out = {}
for hint in (0, 1, -1, 0, -1):
    eng.set_flag("l2hint", hint)
    for herm in (False, True):
        f = lambda: eu.expv(1.0, op, b, m=30, ishermitian=herm)
        for _ in range(3): f()
        eng.set_timing(True); ks = []
        for _ in range(10):
            f(); torch.cuda.synchronize(); ks.append(eng.last_timing()["krylov_ms"])
        eng.set_timing(False)
        print(f"l2hint={hint} {'lanczos' if herm else 'arnoldi'} kernel_ms {np.mean(ks):.4f} (min {np.min(ks):.4f})", flush=True)
