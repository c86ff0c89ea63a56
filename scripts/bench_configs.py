"""Timings of the other BASELINE configs on one GPU (C3 dense phiv, C5 batched share, C4 kiops single GPU).
Usage: python scripts/bench_configs.py [c3] [c5] [c4] [c2var]"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import eu_b200 as eu
from conftest import laplacian2d, convdiff2d

which = set(sys.argv[1:]) or {"c3", "c5", "c4", "c2var"}
try:  # the measured HBM peak of this pool (driver-written); fallback = the profiling guide's figure
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0
eng = eu.get_engine()

def timeit(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def kernel_ms(fn, reps=5):
    eng.set_timing(True); ks = []
    for _ in range(reps):
        fn(); torch.cuda.synchronize(); ks.append(eng.last_timing()["krylov_ms"])
    eng.set_timing(False)
    return float(np.mean(ks))

out = {}
if "c2var" in which:
    n = 10**6
    A = convdiff2d(1000, 1000); op = eu.operator(A)
    b = torch.randn(n, dtype=torch.float64, device="cuda")
    f = lambda: eu.expv(1.0, op, b, m=30)
    ms = timeit(f); k = kernel_ms(f)
    S_A = 12 * A.nnz + 4 * (n + 1); B = 30 * (S_A + 16 * n) + 8 * n * 30 * 31 + 16 * n
    out["c2_convdiff_arnoldi"] = {"ms": ms, "kernel_ms": k, "gbs": B / k / 1e6, "frac": B / k / 1e6 / PEAK, "kernel": eng.last_kernel()}
    f = lambda: eu.expv(1.0, op, b, m=30, iop=2)
    ms = timeit(f); k = kernel_ms(f)
    out["c2_convdiff_iop2"] = {"ms": ms, "kernel_ms": k}
if "c3" in which:
    n = 16384
    g = torch.Generator(device="cuda").manual_seed(2)
    A = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g) / 128
    b = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    op = eu.operator(A); del A
    Ks = eu.KrylovSubspace(n, 30)
    W = torch.empty((5, n), dtype=torch.float64, device="cuda")
    def f():
        eu.arnoldi_(Ks, op, b, m=30, ishermitian=False)
        eu.phiv_(W, 1.0, Ks, 4)
    ms = timeit(f, reps=5, warm=2); k = kernel_ms(f, 3)
    B = 30 * (8 * n * n + 16 * n) + 8 * n * 30 * 31 + 16 * n
    out["c3_phiv_dense16384"] = {"ms": ms, "phiv_per_s": 1e3 / ms, "kernel_ms": k, "gbs": B / k / 1e6, "frac": B / k / 1e6 / PEAK, "kernel": eng.last_kernel()}
if "c5" in which:
    A = laplacian2d(250, 400); n = 100000; nb = 128
    op = eu.operator(A)
    B_ = torch.randn(n, nb, dtype=torch.float64, device="cuda")
    ts = np.random.default_rng(7).uniform(0.1, 1.0, nb)
    for herm, name in ((False, "arnoldi"), (True, "lanczos")):
        f = lambda: eu.expv_batched(ts, op, B_, m=30, ishermitian=herm)
        ms = timeit(f, reps=5, warm=2); k = kernel_ms(f, 3)
        S_A = 12 * A.nnz + 4 * (n + 1)
        # per GPU the operator counts ONCE per Krylov step (shared by the batch), vectors + projection per problem
        per = (30 * 16 * n + 8 * n * 30 * 31 + 16 * n) if not herm else (30 * 24 * n + 16 * n)
        per += 8 * n * 30 + 8 * n
        by = nb * per + 30 * S_A
        out[f"c5_batched128_{name}"] = {"ms": ms, "expv_per_s_per_gpu": nb * 1e3 / ms, "kernel_ms": k,
                                        "bytes_per_gpu": by, "gbs": by / ms / 1e6, "frac": by / ms / 1e6 / PEAK,
                                        "kernel": eng.last_kernel()}
if "c4" in which:
    A = laplacian2d(2500, 4000); n = 10**7
    op = eu.operator(A)
    u = torch.stack([torch.randn(n, dtype=torch.float64, device="cuda"), torch.randn(n, dtype=torch.float64, device="cuda")], 1)
    for herm in (True, False):
        t0 = time.time(); w, st = eu.kiops(1.0, op, u, ishermitian=herm, return_device=True); torch.cuda.synchronize(); dt = time.time() - t0
        t0 = time.time(); w, st = eu.kiops(1.0, op, u, ishermitian=herm, return_device=True); torch.cuda.synchronize(); dt = time.time() - t0
        out[f"c4_kiops_1gpu_herm{int(herm)}"] = {"s": dt, "stats": st}
if "ts" in which:
    # phiv_timestep on the C2 operator, p = 2, adaptive, next to the CPU oracle on the same input
    from oracle import oracle as O
    A = laplacian2d(1000, 1000); n = 10**6
    op = eu.operator(A)
    Bh = np.random.default_rng(11).standard_normal((n, 3))
    Bd = torch.from_numpy(Bh).cuda()
    f = lambda: eu.phiv_timestep([0.5, 1.0], op, Bd, adaptive=True, tol=1e-7, return_steps=True)
    U, ns = f(); ms = timeit(lambda: f(), reps=5, warm=1)
    t0 = time.time(); Uo, nso = O.phiv_timestep([0.5, 1.0], A, Bh, adaptive=True, tol=1e-7, return_steps=True); cpu_s = time.time() - t0
    out["ts_phiv_timestep_p2_adaptive"] = {"ms": ms, "internal_steps": ns, "cpu_oracle_s": cpu_s, "cpu_steps": nso,
                                          "relerr_vs_oracle": float(np.linalg.norm(U.cpu().numpy() - Uo) / np.linalg.norm(Uo))}
    b = torch.from_numpy(Bh[:, 0].copy()).cuda()
    g = lambda: eu.expv(1.0, op, b, mode="error_estimate", m=30, tol=1e-7, return_m=True)
    w, mm = g(); ms = timeit(lambda: g(), reps=10)
    wo, mo = O.expv_ee(1.0, A, Bh[:, 0], m=30, tol=1e-7, return_m=True)
    out["ee_expv_error_estimate"] = {"ms": ms, "m_stop": mm, "m_stop_oracle": mo,
                                    "relerr_vs_oracle": float(np.linalg.norm(w.cpu().numpy() - wo) / np.linalg.norm(wo))}
if "z" in which:
    # Schroedinger-type propagation exp(-i t H) psi, H = -Laplacian (complex Hermitian path -> Lanczos)
    # and a general complex operator (Arnoldi), n = 1e6, m = 30, next to the CPU oracle
    from oracle import oracle as O
    import scipy.sparse as sp
    L = laplacian2d(1000, 1000); n = 10**6
    Hs = (-1.0 * L).astype(np.complex128).tocsr()
    Cz = sp.diags([0.3j * np.ones(n - 1), 0.3j * np.ones(n - 1)], [1, -1])
    Az = (L.astype(np.complex128) + Cz).tocsr()          # complex symmetric, not Hermitian -> Arnoldi
    rng = np.random.default_rng(12)
    psi_h = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    psi = torch.from_numpy(psi_h).cuda()
    for name, M, t in (("z_hermitian_lanczos", Hs, -0.5j), ("z_general_arnoldi", Az, 0.5)):
        op = eu.operator(M)
        f = lambda: eu.expv(t, op, psi, m=30)
        w = f(); ms = timeit(f, reps=10); k = kernel_ms(f, 5)
        t0 = time.time(); wo = O.expv(t, M, psi_h, m=30); cpu_s = time.time() - t0
        S_A = 20 * M.nnz + 4 * (n + 1)
        B = 30 * (S_A + 48 * n) + 32 * n if op.ishermitian else 30 * (S_A + 32 * n) + 16 * n * 30 * 31 + 32 * n
        out[name] = {"ms": ms, "kernel_ms": k, "alg_gbs": B / k / 1e6, "cpu_oracle_s": cpu_s,
                     "relerr_vs_oracle": float(np.linalg.norm(w.cpu().numpy() - wo) / np.linalg.norm(wo)), "kernel": eng.last_kernel()}
print(json.dumps(out, indent=1))
