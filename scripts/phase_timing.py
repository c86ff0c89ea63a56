"""Per-phase cycle breakdown of krylov_tma_kernel (profiling build, -DB200K_PHASE_TIMING).

Build (here, no GPU needed):   python scripts/phase_timing.py build
Run (GPU box):                 B200K_LIB=exponentialutilities.jl_b200/libb200krylov_prof.so \
                               python scripts/phase_timing.py lanczos|arnoldi|c5l|c5a

Marks per step (clock64 of thread 0 of each CTA): 0 step start, 1 after mat-vec, 2 after inner products,
3 after team reduction #1, 4 after update, 5 after team reduction #2, 6 after normalise + publish.
"""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
PROF_LIB = os.path.join(ROOT, "exponentialutilities.jl_b200", "libb200krylov_prof.so")


def build(extra=()):
    """`build` -> libb200krylov_prof.so; `build NAME -DMACRO...` -> libb200krylov_prof_NAME.so (experiments)."""
    import eu_b200 as eu
    b = eu.build
    out = PROF_LIB if not extra else PROF_LIB.replace("_prof.so", f"_prof_{extra[0]}.so")
    cmd = [b._nvcc(), *b.NVCC_FLAGS, "-DB200K_PHASE_TIMING", *extra[1:], "-o", out, os.path.join(b.CSRC, "b200krylov.cu")]
    subprocess.run(cmd, check=True)
    print(out)


def main(which):
    import numpy as np
    import torch
    import eu_b200 as eu
    from conftest import laplacian2d

    lib = eu.load()
    lib.b200k_debug_phase_ts.restype = C.c_int
    lib.b200k_debug_phase_ts.argtypes = [C.c_void_p, C.c_longlong]
    CT, ST, MK = 160, 64, 16
    m = 30
    if which in ("lanczos", "arnoldi"):
        A = laplacian2d(1000, 1000)
        op = eu.operator(A)
        b = torch.randn(10**6, dtype=torch.float64, device="cuda")
        f = lambda: eu.expv(1.0, op, b, m=m, ishermitian=(which == "lanczos"))
        nct = 148
    elif which == "c3":  # dense phiv operator of BASELINE config 3 (general instance, tensor-map tiles)
        n3 = 16384
        g3 = torch.Generator(device="cuda").manual_seed(2)
        A3 = torch.randn(n3, n3, dtype=torch.float64, device="cuda", generator=g3) / 128
        b3 = torch.randn(n3, dtype=torch.float64, device="cuda", generator=g3)
        op = eu.operator(A3)
        del A3
        Ks3 = eu.KrylovSubspace(n3, m)
        f = lambda: eu.arnoldi_(Ks3, op, b3, m=m, ishermitian=False)
        nct = 148
    else:
        A = laplacian2d(250, 400)
        op = eu.operator(A)
        B_ = torch.randn(100000, 128, dtype=torch.float64, device="cuda")
        ts = np.random.default_rng(7).uniform(0.1, 1.0, 128)
        f = lambda: eu.expv_batched(ts, op, B_, m=m, ishermitian=(which == "c5l"))
        nct = 144
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    buf = np.zeros(CT * ST * MK, dtype=np.int64)
    assert lib.b200k_debug_phase_ts(buf.ctypes.data, buf.size) == 0
    allc = buf.reshape(CT, ST, MK)
    used = np.nonzero(allc[:, 5, 0] > 0)[0]  # CTAs that ran (geometry-dependent for batches)
    full = allc[used][:, 1:m + 1, :].astype(np.float64)
    out_n = int(used.size)
    ts_ = full[:, :, :7]
    names = ["matvec", "dots", "reduce1", "update", "reduce2", "normalise"]
    d = np.diff(ts_, axis=2)                       # [cta, step, phase]
    step_total = ts_[:, 1:, 0] - ts_[:, :-1, 0]    # start-to-start
    out = {"which": which, "ctas": out_n, "cycles_per_step_mean": float(step_total[:, 3:].mean()), "phases": {}}
    for k, nm in enumerate(names):
        x = d[:, 3:, k]
        out["phases"][nm] = {"mean": float(x.mean()), "min_cta_mean": float(x.mean(1).min()),
                             "max_cta_mean": float(x.mean(1).max())}
    if full[:, 3:, 7].min() > 0:  # XL instance: sub-phases of the norm reduction
        out["reduce2_parts"] = {"blocksum_fence_publish": float((full[:, 3:, 7] - full[:, 3:, 4]).mean()),
                                "lazy_store_proxyfence": float((full[:, 3:, 8] - full[:, 3:, 7]).mean()),
                                "collect": float((full[:, 3:, 5] - full[:, 3:, 8]).mean())}
    if full[:, 3:, 10].min() > 0:  # XL mat-vec: waiting for the ring vs computing (warps 0 and 15), producer
        out["matvec_parts"] = {"warp0_wait": float(full[:, 3:, 9].mean()), "warp0_compute": float(full[:, 3:, 10].mean()),
                               "warp15_wait": float(full[:, 3:, 11].mean()), "warp15_compute": float(full[:, 3:, 12].mean()),
                               "producer_wait_empty": float(full[:, 3:, 13].mean()),
                               "producer_chunks_total": float(full[:, 3:, 14].mean())}
    if hasattr(lib, "b200k_debug_se_ts") and which in ("lanczos", "arnoldi"):
        se = np.zeros(16, dtype=np.int64)
        lib.b200k_debug_se_ts.argtypes = [C.c_void_p]
        lib.b200k_debug_se_ts(se.ctypes.data)
        names_se = ["load_H", "balance", "norm_scale", "A2", "pade_powers", "A_times_U_and_split", "lu", "backsub",
                    "squaring", "unbalance_store"]
        out["small_exp_cycles"] = {nm: int(se[i + 1] - se[i]) for i, nm in enumerate(names_se)}
        out["small_exp_cycles"]["total"] = int(se[10] - se[0])
        out["small_exp_lu_step5"] = {"pivot_and_barrier": int(se[13] - se[11]), "eliminate": int(se[14] - se[13]),
                                     "barrier": int(se[15] - se[14])}
    out["gap_between_steps"] = float((ts_[:, 1:, 0] - ts_[:, :-1, 6])[:, 3:].mean())
    out["by_step_matvec_mean"] = [float(v) for v in d[:, :, 0].mean(0)]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    else:
        for w in sys.argv[1:]:
            main(w)
