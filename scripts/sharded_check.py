"""Row-sharded expv / kiops across the GPUs of one node: parity against the CPU oracle, then timings.
Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
             scripts/sharded_check.py [parity] [c2] [c4]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
import eu_b200 as eu
from conftest import laplacian2d, convdiff2d, relerr

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
which = set(sys.argv[1:]) or {"parity"}
P = eu.parallel

def log(*a):
    if rank == 0: print(*a, flush=True)

def gather_rows(x_local, ranges):
    """all ranks' row blocks -> full vector on every rank (test plumbing)"""
    mx = max(r[1] for r in ranges)  # equal-size padded blocks (the last rank's block may be longer)
    pad = torch.zeros(mx, dtype=torch.float64, device="cuda")
    pad[: x_local.numel()] = x_local.reshape(-1)
    outs = [torch.empty(mx, dtype=torch.float64, device="cuda") for _ in ranges]
    dist.all_gather(outs, pad)
    return torch.cat([o[: r[1]] for o, r in zip(outs, ranges)]).cpu().numpy()

def make(A, herm):
    n = A.shape[0]
    ranges = P.row_partition(n, world)
    r0, nl = ranges[rank]
    sop = P.ShardedOperator(P.local_block(A, r0, nl), r0, n, ishermitian=herm)
    return sop, ranges, r0, nl

if "parity" in which:
    from oracle import oracle as O
    for name, A, herm in (("laplacian 200x240", laplacian2d(200, 240), True), ("convdiff 200x240", convdiff2d(200, 240), False)):
        n = A.shape[0]
        sop, ranges, r0, nl = make(A, herm)
        b = np.random.default_rng(0).standard_normal(n)
        bl = torch.from_numpy(b[r0:r0 + nl]).cuda()
        for h in ([True, False] if herm else [False]):
            w = eu.expv(1.0, sop.op, bl, m=30, ishermitian=h)
            wf = gather_rows(w, ranges)
            if rank == 0:
                print(f"[{name}] expv herm={h} relerr {relerr(wf, O.expv(1.0, A, b, m=30, ishermitian_=h)):.3e}", flush=True)
        # arnoldi continuation + phiv on the sharded basis
        Ks = eu.KrylovSubspace(nl, 20, engine=sop.engine)
        eu.arnoldi_(Ks, sop.op, bl, m=10, ishermitian=False)
        eu.arnoldi_(Ks, sop.op, bl, m=20, ishermitian=False, init=10)
        Ko = O.KrylovSubspace(n, 20); O.arnoldi_(Ko, A, b, m=20, ishermitian_=False)
        Wl = eu.phiv(0.5, Ks, 3)
        Wf = np.stack([gather_rows(Wl[:, c].contiguous(), ranges) for c in range(4)], 1)
        if rank == 0:
            print(f"[{name}] continuation H err {np.abs(Ks.getH() - Ko.getH()).max():.2e}  phiv relerr {relerr(Wf, O.phiv_ks(0.5, Ko, 3)):.3e}", flush=True)
        # kiops (augmented operator, replicated tail rows)
        u = np.random.default_rng(4).standard_normal((n, 2))
        for h in ([True, False] if herm else [False]):
            wl, st = P.kiops_sharded(1.0, sop, torch.from_numpy(u[r0:r0 + nl]).cuda(), ishermitian=h)
            wf = gather_rows(torch.from_numpy(np.ascontiguousarray(wl[:, 0])).cuda(), ranges)
            if rank == 0:
                wo, so = O.kiops(1.0, A, u, ishermitian_=h)
                print(f"[{name}] kiops herm={h} relerr {relerr(wf, wo[:, 0]):.3e} stats {st} vs {so}", flush=True)
        dist.barrier(); sop.close()

if "dense" in which:
    # dense operator, row blocks: x is all-gathered inside the kernel every step
    from oracle import oracle as O
    n = 2048
    rng = np.random.default_rng(2)
    A = rng.standard_normal((n, n)) / np.sqrt(n) * 4
    Asym = (A + A.T) / 2
    b = np.random.default_rng(3).standard_normal(n)
    ranges = P.row_partition(n, world)
    starts = [r[0] for r in ranges] + [n]
    r0, nl = ranges[rank]
    bl = torch.from_numpy(b[r0:r0 + nl]).cuda()
    for name, M, herm in (("dense randn", A, False), ("dense symmetric", Asym, True)):
        sop = P.ShardedDenseOperator(M[r0:r0 + nl], starts, ishermitian=herm)
        for h in ([True, False] if herm else [False]):
            w = eu.expv(1.0, sop.op, bl, m=30, ishermitian=h)
            wf = gather_rows(w, ranges)
            if rank == 0:
                print(f"[{name} n={n}] expv herm={h} relerr {relerr(wf, O.expv(1.0, M, b, m=30, ishermitian_=h)):.3e} kernel {sop.engine.last_kernel()}", flush=True)
        Ks = eu.KrylovSubspace(nl, 30, engine=sop.engine)
        eu.arnoldi_(Ks, sop.op, bl, m=30, ishermitian=False)
        Ko = O.arnoldi(M, b, m=30, ishermitian_=False)
        Wl = eu.phiv(1.0, Ks, 4, correct=True)
        Wf = np.stack([gather_rows(Wl[:, c].contiguous(), ranges) for c in range(5)], 1)
        if rank == 0:
            print(f"[{name}] H err {np.abs(Ks.getH() - Ko.getH()).max():.2e}  phiv relerr {relerr(Wf, O.phiv_ks(1.0, Ko, 4, correct=True)):.3e}", flush=True)
        u = np.random.default_rng(4).standard_normal((n, 2))
        wl, st = P.kiops_sharded(0.5, sop, torch.from_numpy(u[r0:r0 + nl]).cuda(), ishermitian=herm)
        wf = gather_rows(torch.from_numpy(np.ascontiguousarray(wl[:, 0])).cuda(), ranges)
        if rank == 0:
            wo, so = O.kiops(0.5, M, u, ishermitian_=herm)
            print(f"[{name}] kiops herm={herm} relerr {relerr(wf, wo[:, 0]):.3e} stats {st} vs {so}", flush=True)
        dist.barrier(); sop.close()

if "reorth" in which:
    # ill-conditioned Krylov sequence on a row-sharded operator: the SAFE instance (second Gram-Schmidt pass) with the
    # two-level packet all-reduces, CSR and dense row blocks
    from oracle import oracle as O
    import scipy.sparse as sp
    rng = np.random.default_rng(9)
    n = 2048
    d = np.concatenate([-np.ones(1024), -2 * np.ones(512), -1e3 * np.ones(512)]) + 1e-8 * rng.standard_normal(n)
    D = sp.diags(d).tocsr()
    D.indices = D.indices.astype(np.int32); D.indptr = D.indptr.astype(np.int32)
    b = rng.standard_normal(n)
    sop, ranges, r0, nl = make(D, False)
    bl = torch.from_numpy(b[r0:r0 + nl]).cuda()
    w = eu.expv(0.01, sop.op, bl, m=30, ishermitian=False)
    wf = gather_rows(w, ranges)
    if rank == 0:
        print(f"[clustered diag, CSR sharded] expv relerr {relerr(wf, O.expv(0.01, D, b, m=30, ishermitian_=False)):.3e}", flush=True)
    dist.barrier(); sop.close()
    starts = [r[0] for r in ranges] + [n]
    Dd = D.toarray() + 1e-9 * np.random.default_rng(3).standard_normal((n, n))
    sopd = P.ShardedDenseOperator(Dd[r0:r0 + nl], starts, ishermitian=False)
    w = eu.expv(0.01, sopd.op, bl, m=30, ishermitian=False)
    wf = gather_rows(w, ranges)
    if rank == 0:
        print(f"[clustered diag, dense sharded] expv relerr {relerr(wf, O.expv(0.01, Dd, b, m=30, ishermitian_=False)):.3e}", flush=True)
    dist.barrier(); sopd.close()

def timed(fn, reps, warm=3):
    for _ in range(warm): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

if "c2" in which:
    A = laplacian2d(1000, 1000); n = 10**6
    sop, ranges, r0, nl = make(A, True)
    bl = torch.randn(nl, dtype=torch.float64, device="cuda")
    res = {}
    for h, name in ((False, "arnoldi"), (True, "lanczos")):
        f = lambda: eu.expv(1.0, sop.op, bl, m=30, ishermitian=h)
        ms = timed(f, 20)
        sop.engine.set_timing(True); ks = []
        for _ in range(5):
            f(); torch.cuda.synchronize(); ks.append(sop.engine.last_timing()["krylov_ms"])
        sop.engine.set_timing(False)
        res[name] = {"ms_per_expv": ms, "expv_per_s": 1e3 / ms, "kernel_ms_rank0": float(np.mean(ks)),
                     "us_per_krylov_step": float(np.mean(ks)) * 1e3 / 30}
    log(json.dumps({"c2_row_sharded": res, "n_gpus": world}))
    dist.barrier(); sop.close()

if "c3" in which:
    # BASELINE config 3 (phiv K = 4, dense n = 16384, m = 30) with the operator row-sharded over the ranks
    n = 16384
    ranges = P.row_partition(n, world)
    starts = [r[0] for r in ranges] + [n]
    r0, nl = ranges[rank]
    g = torch.Generator(device="cuda").manual_seed(2)   # every rank generates the same matrix and keeps its rows
    Afull = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g) / 128
    bfull = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    sop = P.ShardedDenseOperator(Afull[r0:r0 + nl], starts, ishermitian=False)
    bl = bfull[r0:r0 + nl].contiguous()
    Ks = eu.KrylovSubspace(nl, 30, engine=sop.engine)
    W = torch.empty((5, nl), dtype=torch.float64, device="cuda")
    def f3():
        eu.arnoldi_(Ks, sop.op, bl, m=30, ishermitian=False)
        eu.phiv_(W, 1.0, Ks, 4)
    ms = timed(f3, 5, 2)
    sop.engine.set_timing(True); ks = []
    for _ in range(3):
        f3(); torch.cuda.synchronize(); ks.append(sop.engine.last_timing()["krylov_ms"])
    sop.engine.set_timing(False)
    res = {"ms_per_phiv": ms, "phiv_per_s": 1e3 / ms, "kernel_ms_rank0": float(np.mean(ks)), "kernel": sop.engine.last_kernel()}
    if world <= 2:  # parity at full size against one GPU holding the whole matrix (rank 0 only has room for it at small world)
        pass
    wf = gather_rows(W[0].contiguous(), ranges)
    if rank == 0:
        op1 = eu.operator(Afull)
        W1 = eu.phiv(1.0, op1, bfull, 4, m=30)
        res["relerr_vs_one_gpu"] = relerr(wf, W1[:, 0].cpu().numpy())
    log(json.dumps({"c3_dense_row_sharded": res, "n_gpus": world}))
    del Afull
    dist.barrier(); sop.close()

if "c4" in which:
    ny4 = int(os.environ.get("C4_NY", 4000))  # (C4_NY = 500 * world emulates the 8-GPU per-GPU load on fewer GPUs)
    A = laplacian2d(2500, ny4); n = 2500 * ny4
    sop, ranges, r0, nl = make(A, True)
    ul = torch.randn(nl, 2, dtype=torch.float64, device="cuda")
    res = {}
    for h in (True, False):
        P.kiops_sharded(1.0, sop, ul, ishermitian=h, return_device=True)
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3): w, st = P.kiops_sharded(1.0, sop, ul, ishermitian=h, return_device=True)
        torch.cuda.synchronize(); dist.barrier(); dt = (time.perf_counter() - t0) / 3
        res[f"herm{int(h)}"] = {"s_per_solve": dt, "stats": st}
    # per-Krylov-step cost at the C4 size (120 MB Lanczos / 150 MB IOP-2 per step per GPU at 8 GPUs)
    bl = torch.randn(nl, dtype=torch.float64, device="cuda")
    for name, kw in (("lanczos", dict(ishermitian=True)), ("iop2", dict(ishermitian=False, iop=2))):
        f = lambda: eu.expv(1.0, sop.op, bl, m=30, **kw)
        ms = timed(f, 10)
        sop.engine.set_timing(True); ks = []
        for _ in range(5):
            f(); torch.cuda.synchronize(); ks.append(sop.engine.last_timing()["krylov_ms"])
        sop.engine.set_timing(False)
        nnz_l = A.nnz / world; n_l = n / world
        step_bytes = (12 * nnz_l + 4 * n_l) + (24 * n_l if name == "lanczos" else 16 * n_l + 32 * n_l)
        us = float(np.mean(ks)) * 1e3 / 30
        res[f"expv_m30_{name}"] = {"ms_per_expv": ms, "us_per_krylov_step": us, "kernel": sop.engine.last_kernel(),
                                   "alg_gbs_per_gpu": step_bytes / us / 1e3}
    log(json.dumps({"c4_kiops_row_sharded": res, "n_gpus": world}))
    dist.barrier(); sop.close()
dist.destroy_process_group()
