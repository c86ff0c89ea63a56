#!/usr/bin/env python
"""bench.py -- expv/s on BASELINE.json configs[1]: CSR 5-point Laplacian n = 10^6, m = 30, fp64.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--path arnoldi|lanczos]

One "step" = one expv(t, A, v): the persistent Krylov kernel (30 fused SpMV + Gram-Schmidt steps), the
small dense exp(tH)e1, and the projection w = beta V y.

* own arm (default):  value   = expv/s with b and w resident in HBM (CUDA events on the launching stream)
                      e2e     = the same call through the C ABI with pinned HOST vectors
                                (b200k_expv_host: H2D(b) + expv + D2H(w) inside the timed region)
                      roofline = algorithmic bytes of the persistent kernel / its measured duration
                      cpu_baseline = the oracle (CPU restatement of the reference) on a bounded sample
* --impl reference:   times the CPU restatement of the reference path (oracle/, Julia is not installed)
                      on the host cores for the same config.
* N > 1 (torchrun):   every rank runs its own independent (t_i, v_i) on the shared operator
                      (embarrassingly parallel replicas, no data-path collective) -> weak scaling.

The headline path is full Arnoldi (ishermitian=false), the north-star kernel; --path lanczos times the
reference's default dispatch for this (symmetric) operator.
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


def algorithmic_bytes(n, nnz, m, path):
    """SURVEY.md 8(d) / BASELINE.md 3: bytes of one factorisation and of the projection."""
    S_A = 12 * nnz + 4 * (n + 1)
    if path == "arnoldi":
        fact = m * (S_A + 16 * n) + 8 * n * m * (m + 1) + 16 * n
    else:
        fact = m * (S_A + 24 * n) + 16 * n
    proj = 8 * n * m + 8 * n
    return fact, proj


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


def cpu_oracle_run(A, b, path, reps):
    """Time the CPU restatement of the reference path (oracle/) on this host.  Returns seconds per expv."""
    from oracle import oracle as O
    herm = path == "lanczos"
    O.expv(T, A, b, m=M, ishermitian_=herm)  # warm-up (page faults, BLAS thread start)
    t0 = time.perf_counter()
    for _ in range(reps):
        O.expv(T, A, b, m=M, ishermitian_=herm)
    return (time.perf_counter() - t0) / reps


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([d.get("num_threads", 1) for d in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--path", default="arnoldi", choices=["arnoldi", "lanczos"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-reps", type=int, default=8)
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
        sec = cpu_oracle_run(A, b, args.path, reps)
        val = 1.0 / sec
        cores = blas_threads()
        line = {
            "impl": "reference", "metric": "expv/s", "value": val, "unit": "expv/s", "n_gpus": args.gpus,
            "steps": reps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": val, "unit": "expv/s", "cores": cores, "kind": "port",
                             "sample": f"{reps} full expv on the host CPU: NumPy/SciPy restatement of the "
                                       "reference (serial CSR mat-vec, OpenBLAS ddot/daxpy/dnrm2 MGS); "
                                       "Julia is not installed, so this is a port, not the package"},
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
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    eng = eu.get_engine(local_rank)

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

    nonlocal_l = [0, 0]

    def timed(fn, steps, warmup):
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
        ms = e0.elapsed_time(e1)
        if dist is not None:
            tt = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total = timed(step_resident, args.steps, args.warmup)
    launches = nonlocal_l[1] - nonlocal_l[0]
    ms_e2e_total = timed(step_e2e, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None

    # kernel-only duration of the persistent Krylov kernel (events on the launching stream, in the library)
    eng.set_timing(True)
    kms, pms = [], []
    for _ in range(max(5, min(args.steps, 20))):
        step_resident()
        torch.cuda.synchronize()
        tm = eng.last_timing()
        kms.append(tm["krylov_ms"])
        pms.append(tm["project_ms"])
    eng.set_timing(False)
    k_ms = float(np.mean(kms))
    p_ms = float(np.mean(pms))

    # parity guard: the timed result must match the CPU oracle (skipped with --no-cpu-baseline)
    ms_per_step = ms_total / args.steps
    value = world * args.steps / (ms_total * 1e-3)
    e2e_value = world * args.steps / (ms_e2e_total * 1e-3)
    fact_bytes, proj_bytes = algorithmic_bytes(n, nnz, M, args.path)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)"
    achieved = fact_bytes / (k_ms * 1e-3) / 1e9
    traffic = None
    try:  # DRAM bytes per launch from the committed ncu --set full capture, if present
        prof = json.load(open(os.path.join(ROOT, "profiles", f"krylov_kernel_{args.path}_traffic.json")))
        traffic = prof.get("dram_bytes_per_launch")
    except Exception:
        pass

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    line = {
        "metric": "expv/s", "value": value, "unit": "expv/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
        "e2e": {"value": e2e_value, "unit": "expv/s", "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 8 * n,
                "ms_per_step": ms_e2e_total / args.steps,
                "note": "b200k_expv_host through the C ABI; operator resident (uploaded once at ingestion)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": {"ldg": "krylov_persistent_kernel", "tma": "krylov_tma_kernel", "tma_xl": "krylov_tma_kernel<XL>"}.get(eng.last_kernel(), "?"), "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "algorithmic_bytes_per_launch": fact_bytes, "kernel_ms": k_ms, "peak_source": peak_src,
                     "project_kernel_ms": p_ms, "project_bytes": proj_bytes,
                     "project_gbs": proj_bytes / (p_ms * 1e-3) / 1e9 if p_ms > 0 else None},
        "hbm_gbs_whole_expv": (fact_bytes + proj_bytes) / (ms_per_step * 1e-3) / 1e9,
        "clocks": clocks,
    }
    # secondary figure: the other Krylov path on the same operator (a few steps, same timing protocol)
    other = "lanczos" if args.path == "arnoldi" else "arnoldi"

    def step_other():
        return eu.expv(t_rank, op, b_dev, m=M, ishermitian=(other == "lanczos"))

    for _ in range(3):
        step_other()
    torch.cuda.synchronize()
    eo0, eo1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eo0.record()
    for _ in range(10):
        step_other()
    eo1.record()
    torch.cuda.synchronize()
    line["also"] = {f"{other}_expv_per_s_one_gpu": 10.0 / (eo0.elapsed_time(eo1) * 1e-3),
                    "note": "same operator through the other path; "
                            "Lanczos is what the reference dispatches to by default for this symmetric operator"}
    if not args.no_cpu_baseline:
        reps = args.cpu_reps
        sec = cpu_oracle_run(A, b_host_np, args.path, reps)
        line["cpu_baseline"] = {
            "value": 1.0 / sec, "unit": "expv/s", "cores": blas_threads(), "kind": "port",
            "sample": f"{reps} full expv of the same workload on the host CPU (oracle/: serial CSR mat-vec + "
                      "OpenBLAS BLAS-1 modified Gram-Schmidt, as the reference does)",
            "host_cores": os.cpu_count()}
        from oracle import oracle as O
        w_ref = O.expv(T, A, b_host_np, m=M, ishermitian_=herm)
        w_gpu = step_resident().cpu().numpy()
        line["parity_rel_err_vs_oracle"] = float(np.linalg.norm(w_gpu - w_ref) / np.linalg.norm(w_ref))
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
