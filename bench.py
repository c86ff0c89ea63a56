#!/usr/bin/env python
"""bench.py -- expv/s on BASELINE.json configs[1]: CSR 5-point Laplacian n = 10^6, m = 30, fp64.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--path arnoldi|lanczos] [--also LIST]

One "step" = one expv(t, A, v): the persistent Krylov kernel (30 fused SpMV + Gram-Schmidt steps), the
small dense exp(tH)e1, and the projection w = beta V y.

* own arm (default):  value   = expv/s with b and w resident in HBM (CUDA events on the launching stream)
                      e2e     = the same call through the C ABI with pinned HOST vectors
                                (b200k_expv_host: H2D(b) + expv + D2H(w) inside the timed region)
                      roofline = algorithmic bytes of the persistent kernel / its measured duration
                      cpu_baseline = the oracle (CPU restatement of the reference) on a bounded sample,
                                plus the best-effort C/OpenMP restatement (oracle/cpu_krylov.c)
* --impl reference:   times the CPU restatement of the reference path (oracle/, Julia is not installed)
                      on the host cores for the same config.
* N > 1 (torchrun):   headline: every rank runs its own independent (t_i, v_i) on the shared operator
                      (embarrassingly parallel replicas, no data-path collective) -> weak scaling.

After the headline the same run measures, under "also" (each with an in-run parity check against the CPU oracle):
    lanczos   the reference's default dispatch for this symmetric operator, with its own roofline (rank 0)
    complex   SURVEY 8(f)-2: ComplexF64 operators of the headline size, general (Arnoldi) and Hermitian (Lanczos)
    c5        BASELINE configs[4]: 1024 independent (t_i, v_i), Laplacian 250x400 (n = 1e5), split over the ranks
    c2s       N >= 2: configs[1] as ONE problem row-sharded over the N GPUs (in-kernel NVLink halo + all-reduce)
    c3s       N >= 2: configs[2] (phiv K = 4, dense n = 16384) with the dense operator in row blocks over the N GPUs
    c4        N == 8: configs[3]: kiops, Laplacian 2500x4000 (n = 1e7), row-sharded over the 8 GPUs
`--also lanczos,complex,c5,c2s,c3s,c4,none` overrides the default selection.

The headline path is full Arnoldi (ishermitian=false), the north-star kernel; --path lanczos times the
reference's default dispatch for this (symmetric) operator as the headline instead.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NX = NY = 1000
M = 30
T = 1.0


def laplacian2d(nx, ny):
    import scipy.sparse as sp
    ex, ey = np.ones(nx), np.ones(ny)
    Tx = sp.diags([ex[:-1], -2 * ex, ex[:-1]], [-1, 0, 1])
    Ty = sp.diags([ey[:-1], -2 * ey, ey[:-1]], [-1, 0, 1])
    A = (sp.kron(sp.identity(ny), Tx) + sp.kron(Ty, sp.identity(nx))).tocsr()
    A.indices = A.indices.astype(np.int32)
    A.indptr = A.indptr.astype(np.int32)
    return A


def laplacian2d_rows(nx, ny, r0, r1):
    """Rows [r0, r1) of laplacian2d(nx, ny) as (indptr int32, indices int64 (global columns), data) without building
    the whole matrix (a rank of the row-sharded runs only needs its own block)."""
    i = np.arange(r0, r1, dtype=np.int64)
    x, y = i % nx, i // nx
    cols = np.stack([i - nx, i - 1, i, i + 1, i + nx], 1)
    mask = np.stack([y > 0, x > 0, np.ones_like(i, dtype=bool), x < nx - 1, y < ny - 1], 1)
    vals = np.broadcast_to(np.array([1.0, 1.0, -4.0, 1.0, 1.0]), cols.shape)
    indptr = np.zeros(i.size + 1, dtype=np.int64)
    np.cumsum(mask.sum(1), out=indptr[1:])
    return indptr.astype(np.int32), np.ascontiguousarray(cols[mask]), np.ascontiguousarray(vals[mask])


def algorithmic_bytes(n, nnz, m, path):
    """SURVEY.md 8(d) / BASELINE.md 3: bytes of one factorisation and of the projection."""
    S_A = 12 * nnz + 4 * (n + 1)
    if path == "arnoldi":
        fact = m * (S_A + 16 * n) + 8 * n * m * (m + 1) + 16 * n
    else:
        fact = m * (S_A + 24 * n) + 16 * n
    proj = 8 * n * m + 8 * n
    return fact, proj


def batch_bytes_per_gpu(n, nnz, m, path, nb_local):
    """per GPU the operator is counted ONCE per Krylov step (shared by the problems of the
    batch); the vector traffic and the projection are per problem (104.9 GB Arnoldi / 12.8 GB Lanczos for 128)."""
    S_A = 12 * nnz + 4 * (n + 1)
    if path == "arnoldi":
        per = m * 16 * n + 8 * n * m * (m + 1) + 16 * n
    else:
        per = m * 24 * n + 16 * n
    per += 8 * n * m + 8 * n
    return nb_local * per + m * S_A


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
            except ValueError:
                continue
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------------
# CPU side: the oracle (reference-faithful port) and the best-effort C/OpenMP restatement
# ------------------------------------------------------------------------------------------------------------
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


try:
    ORIG_AFFINITY = os.sched_getaffinity(0)
except Exception:
    ORIG_AFFINITY = None


class all_host_threads:
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline must still get every host core
    (the BLAS-1 calls of the port are threaded by OpenBLAS, exactly as in the Julia package).  The GPU arm narrows the
    process to the socket next to its GPU (parallel.bind_host_near_gpu); the CPU legs run on the original mask."""

    def __enter__(self):
        self.ctx = None
        self.narrow = None
        try:
            if ORIG_AFFINITY is not None and os.sched_getaffinity(0) != ORIG_AFFINITY:
                self.narrow = os.sched_getaffinity(0)
                os.sched_setaffinity(0, ORIG_AFFINITY)
        except Exception:
            self.narrow = None
        try:
            from threadpoolctl import threadpool_limits
            self.ctx = threadpool_limits(limits=host_threads())
            self.ctx.__enter__()
        except Exception:
            self.ctx = None
        return self

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)
        if self.narrow is not None:
            try:
                os.sched_setaffinity(0, self.narrow)
            except Exception:
                pass


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([d.get("num_threads", 1) for d in threadpool_info()] + [1])
    except Exception:
        return host_threads()


def cpu_oracle_run(A, b, path, reps):
    """Time the CPU restatement of the reference path (oracle/) on this host.  Returns (seconds per expv, threads)."""
    from oracle import oracle as O
    herm = path == "lanczos"
    with all_host_threads():
        O.expv(T, A, b, m=M, ishermitian_=herm)  # warm-up (page faults, BLAS thread start)
        thr = blas_threads()
        t0 = time.perf_counter()
        for _ in range(reps):
            O.expv(T, A, b, m=M, ishermitian_=herm)
        return (time.perf_counter() - t0) / reps, thr


def cpu_best_effort_run(A, b, path, reps):
    """The C/OpenMP restatement (OpenMP CSR mat-vec + fused MGS on all cores).  Returns
    (seconds per expv, threads, relative error against the oracle) or None if it cannot be built on this box."""
    try:
        from oracle import cpu_fast as F
        from oracle import oracle as O
        F.set_threads(host_threads())
        herm = path == "lanczos"
        ws = F.Workspace(A.shape[0], M)
        w = F.expv(T, A, b, m=M, ishermitian_=herm, ws=ws)
        t0 = time.perf_counter()
        for _ in range(reps):
            w = F.expv(T, A, b, m=M, ishermitian_=herm, ws=ws)
        sec = (time.perf_counter() - t0) / reps
        with all_host_threads():
            wo = O.expv(T, A, b, m=M, ishermitian_=herm)
        return sec, F.threads(), float(np.linalg.norm(w - wo) / np.linalg.norm(wo))
    except Exception as e:  # no compiler on the box: report why, never fail the bench for a secondary baseline
        return None, 0, repr(e)


def relerr(a, b):
    nb = np.linalg.norm(b)
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / (nb if nb > 0 else 1.0))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--path", default="arnoldi", choices=["arnoldi", "lanczos"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-reps", type=int, default=8)
    ap.add_argument("--also", default="auto", help="comma list of lanczos,c5,c2s,c3s,c4 or none / auto")
    ap.add_argument("--pipe", type=int, default=4, help="requests in flight in the pipelined end-to-end measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n = NX * NY
    workload = (f"expv(t=1, A, v): CSR 5-pt Laplacian {NX}x{NY} (n=1e6, nnz=4996000), m={M}, fp64, "
                f"{'full Arnoldi (ishermitian=false)' if args.path == 'arnoldi' else 'Lanczos (reference default dispatch)'}")
    config = {"workload": workload, "n": n, "m": M, "path": args.path,
              "l2": "inputs exceed L2: operator 64 MB + basis 248 MB streamed per expv vs 126 MB L2",
              "replicas": "one independent (t_i, v_i) per rank on a shared operator"}

    # ---------------------------------------------------------------- reference arm (CPU) ----------
    if args.impl == "reference":
        if rank != 0:
            return 0
        A = laplacian2d(NX, NY)
        b = np.random.default_rng(0).standard_normal(n)
        reps = max(1, args.steps)
        for _ in range(max(0, args.warmup - 1)):
            cpu_oracle_run(A, b, args.path, 1)
        sec, cores = cpu_oracle_run(A, b, args.path, reps)
        val = 1.0 / sec
        bsec, bthr, berr = cpu_best_effort_run(A, b, args.path, max(3, min(reps, 10)))
        line = {
            "impl": "reference", "metric": "expv/s", "value": val, "unit": "expv/s", "n_gpus": args.gpus,
            "steps": reps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": val, "unit": "expv/s", "cores": cores, "kind": "port",
                             "sample": f"{reps} full expv on the host CPU: NumPy/SciPy restatement of the "
                                       "reference (serial CSR mat-vec, OpenBLAS ddot/daxpy/dnrm2 MGS with "
                                       f"{cores} BLAS threads set explicitly); "
                                       "Julia is not installed, so this is a port, not the package"},
            "cpu_best_effort": {"value": (1.0 / bsec) if bsec else None, "unit": "expv/s", "cores": bthr,
                                "kind": "port-openmp", "rel_err_vs_port": berr,
                                "sample": "C/OpenMP restatement (oracle/cpu_krylov.c): OpenMP CSR mat-vec + fused "
                                          "modified Gram-Schmidt on all cores -- faster than what the Julia package "
                                          "does (its sparse mul! is serial); reported next to the faithful port"},
            "e2e": {"value": val, "unit": "expv/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "host_cores": os.cpu_count(),
        }
        print(json.dumps(line))
        return 0

    # ---------------------------------------------------------------- own arm (B200) ---------------
    import torch
    import eu_b200 as eu

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 engine has no CPU path)")
    torch.cuda.set_device(local_rank)
    from eu_b200 import parallel as eu_parallel
    host_binding = eu_parallel.bind_host_near_gpu(local_rank)  # (before any pinned allocation)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    eng = eu.get_engine(local_rank)

    if args.also == "auto":
        also_set = {"lanczos", "c5", "complex"} | ({"c2s", "c3s"} if world >= 2 else set()) | ({"c4"} if world == 8 else set())
    else:
        also_set = {s for s in args.also.split(",") if s and s != "none"}

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)"

    A = laplacian2d(NX, NY)
    nnz = A.nnz
    op = eu.operator(A)
    herm = args.path == "lanczos"
    b_host_np = np.random.default_rng(rank).standard_normal(n)  # rank r owns its own start vector
    b_dev = torch.from_numpy(b_host_np).to(dev)
    b_pin = torch.from_numpy(b_host_np).pin_memory()
    w_pin = torch.empty(n, dtype=torch.float64).pin_memory()
    t_rank = T

    def step_resident():
        return eu.expv(t_rank, op, b_dev, m=M, ishermitian=herm)

    def step_e2e():
        return eu.expv_host(t_rank, op, b_pin, w_pin, m=M, ishermitian=herm)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return float(x)
        tt = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    nonlocal_l = [0, 0]

    def timed(fn, steps, warmup):
        """ms for `steps` calls: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks."""
        for _ in range(warmup):
            fn()
        barrier()
        nonlocal_l[0] = eng.device_info()["launches"]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        nonlocal_l[1] = eng.device_info()["launches"]
        return max_over_ranks(e0.elapsed_time(e1))

    def kernel_times(fn, reps, lockstep=False):
        """(mean krylov-kernel ms, mean small-exp + projection ms) from the library's own events, this rank.
        lockstep (row-sharded sections: one kernel spans all ranks): the ranks line up before every repetition and the
        median is taken -- a rank that enters the kernel early waits inside it for its peers, which is launch skew, not
        kernel time."""
        eng.set_timing(True)
        kms, pms = [], []
        for _ in range(reps):
            if lockstep:
                barrier()
            fn()
            torch.cuda.synchronize()
            tm = eng.last_timing()
            kms.append(tm["krylov_ms"])
            pms.append(tm["project_ms"])
        eng.set_timing(False)
        if lockstep:
            return float(np.median(kms)), float(np.median(pms))
        return float(np.mean(kms)), float(np.mean(pms))

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total = timed(step_resident, args.steps, args.warmup)
    launches = nonlocal_l[1] - nonlocal_l[0]
    ms_e2e_total = timed(step_e2e, args.steps, args.warmup)

    # e2e with NPIPE requests in flight: one handle / stream each, so the PCIe copies of one request overlap the kernels
    # of the others (b200k_expv_host_async).  Every step still does its own H2D(b) and D2H(w).  Four, not two: the
    # persistent kernel of request i+1 may grab the SMs before the small tail kernels (small exp, projection) of request
    # i; with two in flight the host then learns late that request i is done and the H2D of request i+2 lands on an
    # idle GPU (same box, 36 steps: 578 / 588 / 613 / 620 expv/s with 2 / 3 / 4 / 6 in flight, resident 635).
    NPIPE = max(1, args.pipe)
    eng2 = [eng] + [eu.Engine(local_rank) for _ in range(NPIPE - 1)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(NPIPE)]
    w_pins = [w_pin] + [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(NPIPE - 1)]
    b_pins = [b_pin] + [torch.from_numpy(b_host_np).pin_memory() for _ in range(NPIPE - 1)]

    def run_pipelined(steps):
        for i in range(steps):
            k = i % NPIPE
            with torch.cuda.stream(streams[k]):
                if i >= NPIPE:
                    eng2[k].synchronize()  # the result of request i - NPIPE has landed: its buffers are free again
                eu.expv_host_async(t_rank, op, b_pins[k], w_pins[k], m=M, ishermitian=herm, engine=eng2[k])
        for k in range(NPIPE):
            with torch.cuda.stream(streams[k]):
                eng2[k].synchronize()

    run_pipelined(max(args.warmup, 2 * NPIPE))
    barrier()
    t0 = time.perf_counter()
    run_pipelined(args.steps)
    torch.cuda.synchronize()
    ms_e2e_pipe_total = max_over_ranks((time.perf_counter() - t0) * 1e3)
    barrier()
    pipe_check = max(relerr(w_pins[k].numpy(), w_pins[0].numpy()) for k in range(1, NPIPE))  # all computed the same expv
    clocks = sampler.stop() if rank == 0 else None

    # kernel-only duration of the persistent Krylov kernel (events on the launching stream, in the library)
    k_ms, p_ms = kernel_times(step_resident, max(5, min(args.steps, 20)))
    headline_kernel = eng.last_kernel()

    ms_per_step = ms_total / args.steps
    value = world * args.steps / (ms_total * 1e-3)
    e2e_sync_value = world * args.steps / (ms_e2e_total * 1e-3)
    e2e_value = world * args.steps / (ms_e2e_pipe_total * 1e-3)
    fact_bytes, proj_bytes = algorithmic_bytes(n, nnz, M, args.path)
    achieved = fact_bytes / (k_ms * 1e-3) / 1e9
    traffic = None
    try:  # DRAM bytes per launch from the committed ncu --set full capture, if present
        prof = json.load(open(os.path.join(ROOT, "profiles", f"krylov_kernel_{args.path}_traffic.json")))
        traffic = prof.get("dram_bytes_per_launch")
    except Exception:
        pass

    kname = {"ldg": "krylov_persistent_kernel", "tma": "krylov_tma_kernel", "tma_xl": "krylov_tma_kernel<XL>",
             "tma_xl1": "krylov_tma_kernel<XL, one-reduction Lanczos>",
             "tma_mv": "krylov_mv_kernel"}
    line = {
        "metric": "expv/s", "value": value, "unit": "expv/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
        "e2e": {"value": e2e_value, "unit": "expv/s", "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 8 * n,
                "ms_per_step": ms_e2e_pipe_total / args.steps,
                "mode": "%d requests in flight (b200k_expv_host_async, one handle / stream each): every step copies its "
                        "own 8 MB b from pinned host memory and its own 8 MB w back; the copies of one request overlap "
                        "the kernels of the others; host wall clock around the whole loop incl. the final synchronise" % NPIPE,
                "requests_in_flight": NPIPE,
                "one_call_at_a_time": {"value": e2e_sync_value, "ms_per_step": ms_e2e_total / args.steps,
                                       "note": "synchronous b200k_expv_host: H2D -> kernels -> D2H serialised per call "
                                               "(CUDA events)"},
                "pipelined_results_agree": pipe_check,
                "host_binding": host_binding,
                "note": "through the C ABI with HOST buffers; operator resident (uploaded once at ingestion); the rank's "
                        "host threads and pinned buffers sit on the NUMA node of its GPU (host_binding)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": kname.get(headline_kernel, "?"), "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "traffic_source": "profiles/krylov_kernel_%s_traffic.json (ncu --set full, committed; not "
                                       "re-measured in this run)" % args.path,
                     "algorithmic_bytes_per_launch": fact_bytes, "kernel_ms": k_ms, "peak_source": peak_src,
                     "smallexp_plus_project_ms": p_ms, "project_bytes": proj_bytes,
                     "note": "smallexp_plus_project_ms spans small_exp_kernel AND project_kernel (one event pair); "
                             "the projection alone streams project_bytes"},
        "hbm_gbs_whole_expv": (fact_bytes + proj_bytes) / (ms_per_step * 1e-3) / 1e9,
        "clocks": clocks,
    }
    also = {}

    # ------------------------------------------------------------------------------------------------------
    # also.lanczos: the other Krylov path on the same operator (rank-local, same timing protocol)
    # ------------------------------------------------------------------------------------------------------
    other = "lanczos" if args.path == "arnoldi" else "arnoldi"
    if "lanczos" in also_set:
        def step_other():
            return eu.expv(t_rank, op, b_dev, m=M, ishermitian=(other == "lanczos"))

        ms_o = timed(step_other, 20, 3) / 20
        ko_ms, po_ms = kernel_times(step_other, 10)
        fo, pro = algorithmic_bytes(n, nnz, M, other)
        sec = {"path": other, "expv_per_s_one_gpu": 1e3 / ms_o, "ms_per_expv": ms_o, "kernel": kname.get(eng.last_kernel(), "?"),
               "kernel_ms": ko_ms, "smallexp_plus_project_ms": po_ms,
               "roofline": {"bound": "hbm", "achieved": fo / (ko_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                            "frac": fo / (ko_ms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": fo},
               "whole_expv": {"achieved": (fo + pro) / (ms_o * 1e-3) / 1e9, "frac": (fo + pro) / (ms_o * 1e-3) / 1e9 / peak,
                              "bytes": fo + pro, "bar_0.70_expv_per_s": 0.70 * peak * 1e9 / (fo + pro)},
               "note": "same operator through the other path; Lanczos is what the reference dispatches to by default "
                       "for this symmetric operator (src/arnoldi.jl:355-356)"}
        if rank == 0 and not args.no_cpu_baseline:
            from oracle import oracle as O
            with all_host_threads():
                w_ref = O.expv(T, A, b_host_np, m=M, ishermitian_=(other == "lanczos"))
            sec["parity_rel_err_vs_oracle"] = relerr(step_other().cpu().numpy(), w_ref)
        also[f"{other}_expv_per_s_one_gpu"] = sec["expv_per_s_one_gpu"]  # (round-1 key, kept)
        also[other] = sec

    # ------------------------------------------------------------------------------------------------------
    # also.complex: SURVEY 8(f)-2 -- ComplexF64 operators of the headline size (n = 1e6, m = 30): a general complex one
    # (Arnoldi) and a Hermitian one (Schroedinger-type propagation exp(-i t H) psi, Lanczos with real coefficients)
    # ------------------------------------------------------------------------------------------------------
    if "complex" in also_set:
        import scipy.sparse as sp
        Hs = (-1.0 * A).astype(np.complex128).tocsr()
        Az = (A.astype(np.complex128) + sp.diags([0.3j * np.ones(n - 1), 0.3j * np.ones(n - 1)], [1, -1])).tocsr()
        rngz = np.random.default_rng(12 + rank)
        psi_h = rngz.standard_normal(n) + 1j * rngz.standard_normal(n)
        psi = torch.from_numpy(psi_h).to(dev)
        secz = {"n": n, "m": M, "byte_model": "BASELINE.md 3 with 16-byte vector elements and 20 bytes per stored entry"}
        for name, Mz, tz in (("general_arnoldi", Az, 0.5), ("hermitian_lanczos", Hs, -0.5j)):
            opz = eu.operator(Mz)

            def stepz():
                return eu.expv(tz, opz, psi, m=M)

            ms_z = timed(stepz, 10, 3) / 10
            kz, pz = kernel_times(stepz, 5)
            S_Az = 20 * Mz.nnz + 4 * (n + 1)
            Bz = (30 * (S_Az + 48 * n) + 32 * n) if opz.ishermitian else (30 * (S_Az + 32 * n) + 16 * n * 30 * 31 + 32 * n)
            ez = {"ms_per_expv": ms_z, "expv_per_s_one_gpu": 1e3 / ms_z, "kernel": eng.last_kernel(), "kernel_ms": kz,
                  "roofline": {"bound": "hbm", "achieved": Bz / (kz * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                               "frac": Bz / (kz * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": Bz}}
            if rank == 0 and not args.no_cpu_baseline:
                from oracle import oracle as O
                with all_host_threads():
                    t0 = time.perf_counter()
                    wz = O.expv(tz, Mz, psi_h, m=M)
                    ez["cpu_oracle_s"] = time.perf_counter() - t0
                ez["parity_rel_err_vs_oracle"] = relerr(stepz().cpu().numpy(), wz)
            secz[name] = ez
            del opz
        also["complex"] = secz

    # ------------------------------------------------------------------------------------------------------
    # also.c5: BASELINE configs[4] -- 1024 independent (t_i, v_i) on a shared Laplacian 250x400, split over the ranks
    # ------------------------------------------------------------------------------------------------------
    if "c5" in also_set:
        NB = 1024
        A5 = laplacian2d(250, 400)
        n5, nnz5 = A5.shape[0], A5.nnz
        lo, hi = eu.parallel.shard_batch(NB, rank, world)
        nbl = hi - lo
        ts_all = np.random.default_rng(7).uniform(0.1, 1.0, NB)
        op5 = eu.operator(A5)
        # V0 = randn(n, 1024) seed 6, generated per column so that every rank can make exactly its own share
        cols = {}

        def v0(i):
            if i not in cols:
                cols[i] = np.random.default_rng([6, i]).standard_normal(n5)
            return cols[i]

        g = torch.Generator(device=dev).manual_seed(600 + rank)
        Bt = torch.randn((nbl, n5), dtype=torch.float64, device=dev, generator=g)
        rng_pick = np.random.default_rng(1000 + rank)
        ncheck = max(2, -(-16 // world))
        pick = sorted(rng_pick.choice(nbl, size=min(ncheck, nbl), replace=False).tolist())
        for q in pick:  # the checked columns carry the seeded vectors the oracle also gets
            Bt[q] = torch.from_numpy(v0(lo + q)).to(dev)
        Wt = torch.empty((nbl, n5), dtype=torch.float64, device=dev)
        sec5 = {"problems": NB, "problems_per_rank": nbl, "n": n5, "nnz": nnz5, "m": M,
                "byte_model": "BASELINE.md 3: operator once per Krylov step per GPU + per-problem vector traffic + projection"}
        for path5 in ("arnoldi", "lanczos"):
            h5 = path5 == "lanczos"

            def step5():
                return eu.expv_batched(ts_all[lo:hi], op5, Bt.t(), m=M, ishermitian=h5, out=Wt)

            ms5 = timed(step5, 5, 2) / 5
            k5, p5 = kernel_times(step5, 3)
            by = batch_bytes_per_gpu(n5, nnz5, M, path5, nbl)
            err5 = 0.0
            if not args.no_cpu_baseline:
                from oracle import oracle as O
                W = step5()
                with all_host_threads():
                    for q in pick:
                        wo = O.expv(float(ts_all[lo + q]), A5, v0(lo + q), m=M, ishermitian_=h5)
                        err5 = max(err5, relerr(W[:, q].cpu().numpy(), wo))
                err5 = max_over_ranks(err5)
            agg = NB / (ms5 * 1e-3)
            sec5[path5] = {"expv_per_s": agg, "expv_per_s_per_gpu": agg / world, "ms_per_batch": ms5,
                           "kernel": kname.get(eng.last_kernel(), "?"), "kernel_ms_rank0": k5,
                           "smallexp_plus_project_ms_rank0": p5, "bytes_per_gpu": by,
                           "achieved_gbs_per_gpu": by / (ms5 * 1e-3) / 1e9, "frac": by / (ms5 * 1e-3) / 1e9 / peak,
                           "bar_0.70_expv_per_s": 0.70 * peak * 1e9 / by * nbl * world,
                           "parity_rel_err_vs_oracle_max": err5 if not args.no_cpu_baseline else None,
                           "parity_columns_checked": len(pick) * world}
        also["c5_batched_1024"] = sec5
        del Bt, Wt, op5

    # ------------------------------------------------------------------------------------------------------
    # row-sharded sections (ONE problem over the N GPUs; halo exchange + all-reduces inside the persistent kernel)
    # ------------------------------------------------------------------------------------------------------
    P = eu.parallel

    def gather_rows(x_local, ranges):
        """Row blocks of all ranks -> the full vector on every rank (NCCL all_gather of equal-size padded blocks)."""
        mx = max(r[1] for r in ranges)
        pad = torch.zeros(mx, dtype=torch.float64, device=dev)
        pad[: x_local.numel()] = x_local.reshape(-1)
        outs = [torch.empty(mx, dtype=torch.float64, device=dev) for _ in ranges]
        dist.all_gather(outs, pad)
        return torch.cat([o[: r[1]] for o, r in zip(outs, ranges)]).cpu().numpy()

    if "c2s" in also_set and world >= 2:
        ranges = P.row_partition(n, world)
        r0, nl = ranges[rank]
        sop = P.ShardedOperator(laplacian2d_rows(NX, NY, r0, r0 + nl), r0, n, ishermitian=True)
        b2 = np.random.default_rng(0).standard_normal(n)
        bl = torch.from_numpy(b2[r0:r0 + nl]).to(dev)
        sec2 = {"n": n, "rows_per_rank": nl, "halo_entries_rank0": sop.nhalo,
                "one_gpu_resident_ms_per_expv": {args.path: ms_per_step}}
        if other in also:
            sec2["one_gpu_resident_ms_per_expv"][other] = also[other]["ms_per_expv"]
        for path2 in ("arnoldi", "lanczos"):
            h2 = path2 == "lanczos"

            def step2():
                return eu.expv(T, sop.op, bl, m=M, ishermitian=h2)

            ms2 = timed(step2, 20, 3) / 20
            k2, _ = kernel_times(step2, 7, lockstep=True)
            k2 = max_over_ranks(k2)
            wf = gather_rows(step2(), ranges)
            r = {"ms_per_expv": ms2, "expv_per_s": 1e3 / ms2, "kernel": kname.get(eng.last_kernel(), "?"),
                 "kernel_ms_max_over_ranks": k2, "us_per_krylov_step": k2 * 1e3 / M}
            one = sec2["one_gpu_resident_ms_per_expv"].get(path2)
            if one:
                r["speedup_vs_one_gpu"] = one / ms2
                r["strong_scaling_efficiency"] = one / ms2 / world
            if rank == 0 and not args.no_cpu_baseline:
                from oracle import oracle as O
                with all_host_threads():
                    r["parity_rel_err_vs_oracle"] = relerr(wf, O.expv(T, A, b2, m=M, ishermitian_=h2))
            sec2[path2] = r
        also["c2_row_sharded"] = sec2
        barrier()
        sop.close()

    # ------------------------------------------------------------------------------------------------------
    # also.c3_dense_row_sharded: BASELINE configs[2] (phiv K = 4, dense n = 16384, m = 30) with the operator split into
    # row blocks over the N GPUs: the all-gather of x happens inside the persistent kernel
    # ------------------------------------------------------------------------------------------------------
    if "c3s" in also_set and world >= 2:
        n3 = 16384
        ranges = P.row_partition(n3, world)
        starts = [r[0] for r in ranges] + [n3]
        r0, nl = ranges[rank]
        g3 = torch.Generator(device=dev).manual_seed(2)  # every rank generates the same matrix and keeps its rows
        A3 = torch.randn(n3, n3, dtype=torch.float64, device=dev, generator=g3) / 128
        b3 = torch.randn(n3, dtype=torch.float64, device=dev, generator=g3)
        sop = P.ShardedDenseOperator(A3[r0:r0 + nl], starts, ishermitian=False)
        bl = b3[r0:r0 + nl].contiguous()
        Ks3 = eu.KrylovSubspace(nl, M, engine=eng)
        W3 = torch.empty((5, nl), dtype=torch.float64, device=dev)

        def step3():
            eu.arnoldi_(Ks3, sop.op, bl, m=M, ishermitian=False)
            eu.phiv_(W3, 1.0, Ks3, 4)

        ms3 = timed(step3, 5, 2) / 5
        by3 = M * (8 * n3 * n3 + 16 * n3) + 8 * n3 * M * (M + 1) + 16 * n3 + 8 * n3 * M + 8 * n3 * 5  # BASELINE.md 3 (whole operator)
        sec3 = {"n": n3, "k": 4, "m": M, "rows_per_rank": nl, "ms_per_phiv": ms3, "phiv_per_s": 1e3 / ms3,
                "kernel": kname.get(eng.last_kernel(), "?"), "bytes_all_gpus": by3,
                "achieved_gbs_per_gpu": by3 / world / (ms3 * 1e-3) / 1e9, "frac": by3 / world / (ms3 * 1e-3) / 1e9 / peak,
                "x_allgather": "in-kernel peer stores: %d doubles pushed per rank and step" % (nl * (world - 1))}
        cols3 = [gather_rows(W3[c].contiguous(), ranges) for c in range(5)]
        if rank == 0:
            op1 = eu.operator(A3)  # the whole matrix on one GPU: same call, same library
            W1 = eu.phiv(1.0, op1, b3, 4, m=M)
            sec3["rel_err_vs_one_gpu"] = relerr(np.stack(cols3, 1), W1.cpu().numpy())
            del op1, W1
            if not args.no_cpu_baseline:
                from oracle import oracle as O
                with all_host_threads():
                    Wo = O.phiv(1.0, A3.cpu().numpy(), b3.cpu().numpy(), 4, m=M)
                sec3["parity_rel_err_vs_oracle"] = relerr(np.stack(cols3, 1), Wo)
        also["c3_dense_row_sharded"] = sec3
        del A3
        barrier()
        sop.close()

    if "c4" in also_set and world >= 2:
        NX4, NY4 = 2500, 4000
        n4, nnz4 = NX4 * NY4, 49_987_000
        ranges = P.row_partition(n4, world)
        r0, nl = ranges[rank]
        sop = P.ShardedOperator(laplacian2d_rows(NX4, NY4, r0, r0 + nl), r0, n4, ishermitian=True)
        u4 = np.stack([np.random.default_rng(4).standard_normal(n4), np.random.default_rng(5).standard_normal(n4)], 1)
        ul = torch.from_numpy(np.ascontiguousarray(u4[r0:r0 + nl])).to(dev)
        sec4 = {"n": n4, "nnz": nnz4, "rows_per_rank": nl, "tau_out": 1.0, "kiops": "defaults (tol 1e-7, mmin 10, mmax 128, iop 2)"}
        A4 = None
        if rank == 0 and not args.no_cpu_baseline:
            A4 = laplacian2d(NX4, NY4)
        for h4, name in ((True, "kiops_hermitian_lanczos"), (False, "kiops_iop2")):
            def solve():
                return P.kiops_sharded(1.0, sop, ul, ishermitian=h4, return_device=True)

            solve()
            barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                w4, st4 = solve()
            barrier()
            dt = max_over_ranks((time.perf_counter() - t0) / 3)
            wf = gather_rows(w4[:, 0].contiguous(), ranges)
            r = {"ms_per_solve": dt * 1e3, "stats": list(st4),
                 "timing": "host wall clock around 3 solves (kiops is a host controller with device syncs), max over ranks"}
            if A4 is not None:
                from oracle import oracle as O
                with all_host_threads():
                    wo, so = O.kiops(1.0, A4, u4, ishermitian_=h4)
                r["parity_rel_err_vs_oracle"] = relerr(wf, wo[:, 0])
                r["stats_oracle"] = list(so)
                r["stats_equal"] = tuple(st4) == tuple(so)
            sec4[name] = r
        # per-Krylov-step cost at the C4 size (120 MB Lanczos / 150 MB IOP-2 per step per GPU at 8 GPUs)
        bl = torch.from_numpy(np.ascontiguousarray(u4[r0:r0 + nl, 0])).to(dev)
        for name, kw in (("expv_m30_lanczos", dict(ishermitian=True)), ("expv_m30_iop2", dict(ishermitian=False, iop=2))):
            def step4():
                return eu.expv(1.0, sop.op, bl, m=M, **kw)

            ms4 = timed(step4, 10, 3) / 10
            k4, _ = kernel_times(step4, 7, lockstep=True)
            k4 = max_over_ranks(k4)
            nnz_l, n_l = nnz4 / world, n4 / world
            step_bytes = (12 * nnz_l + 4 * n_l) + (24 * n_l if "lanczos" in name else 16 * n_l + 32 * n_l)
            us = k4 * 1e3 / M
            sec4[name] = {"ms_per_expv": ms4, "us_per_krylov_step": us, "kernel": kname.get(eng.last_kernel(), "?"),
                          "bytes_per_step_per_gpu": step_bytes, "achieved_gbs_per_gpu": step_bytes / us / 1e3,
                          "frac": step_bytes / us / 1e3 / peak, "us_per_step_at_100pct": step_bytes / peak / 1e3}
        also["c4_kiops_row_sharded"] = sec4
        barrier()
        sop.close()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    line["also"] = also
    if not args.no_cpu_baseline:
        reps = args.cpu_reps
        sec, thr = cpu_oracle_run(A, b_host_np, args.path, reps)
        line["cpu_baseline"] = {
            "value": 1.0 / sec, "unit": "expv/s", "cores": thr, "kind": "port",
            "sample": f"{reps} full expv of the same workload on the host CPU (oracle/: serial CSR mat-vec + "
                      f"OpenBLAS BLAS-1 modified Gram-Schmidt with {thr} threads, as the reference does)",
            "host_cores": os.cpu_count()}
        bsec, bthr, berr = cpu_best_effort_run(A, b_host_np, args.path, reps)
        line["cpu_best_effort"] = {
            "value": (1.0 / bsec) if bsec else None, "unit": "expv/s", "cores": bthr, "kind": "port-openmp",
            "rel_err_vs_port": berr,
            "sample": f"{reps} full expv, C/OpenMP restatement (oracle/cpu_krylov.c): OpenMP CSR mat-vec + fused MGS on "
                      "all cores (BASELINE.md 4.2)"}
        from oracle import oracle as O
        with all_host_threads():
            w_ref = O.expv(T, A, b_host_np, m=M, ishermitian_=herm)
        w_gpu = step_resident().cpu().numpy()
        line["parity_rel_err_vs_oracle"] = relerr(w_gpu, w_ref)
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
