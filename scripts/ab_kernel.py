"""A/B timing of the persistent Krylov kernel with different builds of the library (B200K_LIB=path):
    B200K_LIB=scripts/ab/r1.so python scripts/ab_kernel.py [arnoldi] [lanczos]
Uses only entry points that exist in every build since round 1; prints kernel ms (library events) and ms per expv."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import eu_b200 as eu
from conftest import laplacian2d

which = sys.argv[1:] or ["arnoldi", "lanczos"]
A = laplacian2d(1000, 1000); n = 10**6
op = eu.operator(A); eng = eu.get_engine()
b = torch.from_numpy(np.random.default_rng(0).standard_normal(n)).cuda()
out = {"lib": os.environ.get("B200K_LIB", "current")}
for path in which:
    herm = path == "lanczos"
    f = lambda: eu.expv(1.0, op, b, m=30, ishermitian=herm)
    for _ in range(5): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30): f()
    e1.record(); torch.cuda.synchronize()
    eng.set_timing(True); ks = []
    for _ in range(10):
        f(); torch.cuda.synchronize(); ks.append(eng.last_timing()["krylov_ms"])
    eng.set_timing(False)
    out[path] = {"ms_per_expv": e0.elapsed_time(e1) / 30, "kernel_ms": float(np.mean(ks)), "kernel_ms_min": float(np.min(ks))}
print(json.dumps(out))
