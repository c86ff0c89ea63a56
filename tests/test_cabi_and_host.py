"""CPU-only checks of the product boundary: the shared library loads, exports every symbol that
include/b200krylov.h declares, fails loudly without a GPU, and its host-side small dense functions
(the part of the path that runs on the host by design) match the oracle."""
import os
import re
import subprocess

import numpy as np
import pytest
import scipy.linalg as sla

from conftest import ROOT, relerr


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "b200krylov.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200k_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(eu):
    lib = eu.load()
    syms = header_symbols()
    assert len(syms) >= 25
    out = subprocess.run(["nm", "-D", "--defined-only", eu.lib_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\sT\s+(b200k_\w+)", out))
    for s in syms:
        assert s in exported, f"{s} declared in include/b200krylov.h but not exported"
        assert hasattr(lib, s)
    from importlib import import_module
    protos = import_module("eu_b200._lib").PROTOTYPES
    assert set(protos) == set(syms), set(protos) ^ set(syms)


def test_plain_c_consumer_compiles_links_and_runs(eu, tmp_path):
    """include/b200krylov.h is C (not just ctypes-) clean: a C99 translation unit compiled with gcc -pedantic links
    against the shared library and calls the host-side entry points; struct sizes agree with b200k_sizeof."""
    lib = eu.load()
    exe = str(tmp_path / "consumer")
    src = os.path.join(ROOT, "tests", "cconsumer", "consumer.c")
    libdir = os.path.dirname(eu.lib_path())
    cc = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src,
                         "-o", exe, "-L", libdir, "-l:libb200krylov.so", "-lm", f"-Wl,-rpath,{libdir}"],
                        capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    run = subprocess.run([exe], capture_output=True, text=True)
    assert run.returncode == 0, (run.returncode, run.stdout, run.stderr)
    assert "consumer ok" in run.stdout
    import ctypes as C
    from importlib import import_module
    L = import_module("eu_b200._lib")
    assert lib.b200k_sizeof(1) == C.sizeof(L.KrylovOpts)
    assert lib.b200k_sizeof(2) == C.sizeof(L.KiopsOpts)
    assert lib.b200k_sizeof(3) == C.sizeof(L.TimestepOpts)


def test_no_cpu_fallback(eu):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        eu.expv(1.0, np.eye(4), np.ones(4))
    # the raw ABI also refuses: no device -> B200K_ECUDA
    import ctypes as C
    h = C.c_void_p()
    assert eu.load().b200k_create(C.byref(h), 0, None) == 4


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "exponentialutilities.jl_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                assert "oracle" not in open(os.path.join(dp, f)).read().lower().replace("checked against the cpu oracle", ""), f


def test_exponential_matches_oracle_all_branches(eu, oracle):
    rng = np.random.default_rng(11)
    for scale in (300.0, 40.0, 3.0, 1.5, 0.5, 0.1, 0.005, 0.0):
        A = rng.standard_normal((34, 34))
        A *= scale / np.linalg.norm(A, 1)
        E = eu.exponential_(A)
        assert relerr(E, oracle.exponential_higham2005base(A)) < 1e-12
        if scale <= 40:
            assert relerr(E, sla.expm(A)) < 1e-11


def test_exponential_badly_scaled_uses_balancing(eu, oracle):
    rng = np.random.default_rng(12)
    A = rng.standard_normal((12, 12))
    D = np.diag(2.0 ** rng.integers(-20, 20, 12))
    B = D @ A @ np.linalg.inv(D)
    assert relerr(eu.exponential_(B), oracle.exponential_higham2005base(B)) < 1e-10
    # upper Hessenberg with an isolated eigenvalue exercises the permutation phase
    Hh = np.triu(rng.standard_normal((9, 9)), -1)
    Hh[5, 4] = 0.0
    Hh[8, :8] = 0.0
    assert relerr(eu.exponential_(Hh), sla.expm(Hh)) < 1e-12


def test_expv_small_branches(eu, oracle):
    rng = np.random.default_rng(13)
    m = 30
    d = rng.standard_normal(m) - 4
    e = np.abs(rng.standard_normal(m - 1)) + 0.5
    Tm = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    y, br = eu.expv_small(Tm, 0.7)
    assert br == 1
    assert relerr(y, sla.expm(0.7 * Tm)[:, 0]) < 1e-13
    Hn = np.triu(rng.standard_normal((m, m)), -1)
    y, br = eu.expv_small(Hn, 0.2)
    assert br == 0
    assert relerr(y, oracle.exponential_higham2005base(0.2 * Hn)[:, 0]) < 1e-13
    # 1 x 1 and 2 x 2 edge cases
    y, _ = eu.expv_small(np.array([[-3.0]]), 2.0)
    assert abs(y[0] - np.exp(-6.0)) < 1e-15
    y, br = eu.expv_small(np.array([[1.0, 2.0], [2.0, -1.0]]), 0.5)
    assert br == 1 and relerr(y, sla.expm(0.5 * np.array([[1.0, 2.0], [2.0, -1.0]]))[:, 0]) < 1e-14


def test_phiv_dense_matches_oracle(eu, oracle):
    rng = np.random.default_rng(14)
    for m, k in ((5, 1), (10, 4), (30, 4), (20, 10)):
        H = np.triu(rng.standard_normal((m, m)), -1)
        v = rng.standard_normal(m)
        assert relerr(eu.phiv_dense(H, v, k), oracle.phiv_dense(H, v, k)) < 1e-12


def test_error_mapping(eu):
    with pytest.raises(eu.DimensionMismatch):
        eu.exponential_(np.zeros((3, 4)))
    with pytest.raises(eu.DimensionMismatch):
        eu.phiv_dense(np.zeros((3, 3)), np.zeros(4), 2)
    import eu_b200._lib as L
    assert L.load().b200k_status_string(3).decode().startswith("singular")


def test_cache_mirrors_and_their_error(eu):
    """expv!/phiv! reject a cache of the wrong type with ArgumentError (src/krylov_phiv.jl:221, 630); the check comes
    before anything touches the device."""
    # ExpvCache: maxiter^2 elements, get_cache views m x m, resize! doubles (src/krylov_phiv.jl:45-77)
    c = eu.ExpvCache(30)
    assert c.mem.size == 900 and c.expcol.size == 30 and c.maxiter == 30
    v = c.get_cache(20)
    assert v.shape == (20, 20) and v.flags.f_contiguous and np.shares_memory(v, c.mem)
    assert c.get_cache(40).shape == (40, 40) and c.mem.size == 40 * 40 * 2 and c.expcol.size == 40  # grown on demand
    for n in range(70):  # size-keyed exponential! workspaces: FIFO store bounded at 64 (krylov_phiv.jl:447-470)
        c.get_expcache(n)
    assert len(c.expcache) == 64 and c.expcache[0][0] == 6 and c.get_expcache(69) is c.expcache[-1][1]
    # PhivCache: m + m^2 + (m+p)^2 + m(p+1) elements split by get_caches (krylov_phiv.jl:404-428, 479-504)
    pc = eu.PhivCache(None, 30, 4)
    assert pc.mem.size == 30 + 900 + 34 * 34 + 30 * 5 and pc.coeffs.size == 4 and pc.useview
    e, Hc, C1, C2 = pc.get_caches(10, 3)
    assert e.shape == (10,) and Hc.shape == (10, 10) and C1.shape == (13, 13) and C2.shape == (10, 4)
    assert all(np.shares_memory(x, pc.mem) for x in (e, Hc, C1, C2))
    pc.get_caches(50, 4)
    assert pc.mem.size == 2 * (50 + 2500 + 54 * 54 + 50 * 5)
    with pytest.raises(eu.ArgumentError):
        eu.expv_(None, 1.0, None, cache=eu.PhivCache(None, 30, 4))
    with pytest.raises(eu.ArgumentError):
        eu.phiv_(None, 1.0, None, 2, cache=object())


def test_runtime_flags_match_header(eu):
    """The B200K_FLAG_* constants of include/b200krylov.h are the ones the host mirror passes to b200k_set_flag."""
    import inspect
    txt = open(os.path.join(ROOT, "include", "b200krylov.h")).read()
    flags = {m.group(1).lower(): int(m.group(2)) for m in re.finditer(r"#define\s+B200K_FLAG_(\w+)\s+(\d+)", txt)}
    assert flags == {"force_ldg": 1, "host_smallexp": 2, "l2hint": 3, "no_xl": 4, "no_mv": 5, "sym_pade": 6, "no_lz1": 7}
    src = inspect.getsource(eu.api.Engine.set_flag)
    for name, val in flags.items():
        assert f'"{name}": {val}' in src, name



def test_bench_contract_on_cpu():
    """bench.py: the reference arm runs without a GPU and prints ONE JSON line with the contract's keys; the B200 arm
    refuses to run without a CUDA device (no CPU fallback)."""
    import json
    import subprocess
    import sys
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-reps", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "expv/s" and line["unit"] == "expv/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["value"] > 0
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0 == line["e2e"]["d2h_bytes_per_step"]
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == line["value"]
    assert "workload" in line["config"] and line["dtype"] == "f64"
    if not torch.cuda.is_available():
        r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1", "--warmup", "0"],
                           capture_output=True, text=True, timeout=600, cwd=root)
        assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
