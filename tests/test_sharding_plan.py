"""Host-side logic of the multi-GPU paths, on CPU: row partition, halo / send-list planning, and a
world_size-2 gloo run that performs the halo exchange the kernel does (each rank stores its boundary rows at
the planned positions of the peer's gather buffer) and checks the local mat-vec against the global one."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, convdiff2d, laplacian2d


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_row_partition_and_batch_shards(eu):
    P = eu.parallel
    for n, w in ((10 ** 6, 8), (10 ** 7, 8), (1920, 3), (100, 2)):
        parts = P.row_partition(n, w)
        assert parts[0][0] == 0 and sum(nl for _, nl in parts) == n
        assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        assert all(r0 % 16 == 0 for r0, _ in parts)
    seen = []
    for r in range(8):
        lo, hi = P.shard_batch(1024, r, 8)
        assert hi - lo == 128
        seen += list(range(lo, hi))
    assert seen == list(range(1024))
    assert [P.shard_batch(10, r, 4) for r in range(4)] == [(0, 2), (2, 5), (5, 7), (7, 10)]


def test_halo_plan_single_process(eu):
    P = eu.parallel
    A = convdiff2d(24, 20)
    n = A.shape[0]
    parts = P.row_partition(n, 3)
    blocks, halos = [], []
    for r0, nl in parts:
        ip, ix, d = P.local_block(A, r0, nl)
        halo, loc = P.plan_halo(ix, r0, nl)
        assert np.all(np.diff(halo) > 0) and loc.min() >= 0 and loc.max() < nl + len(halo)
        blocks.append((ip, loc, d))
        halos.append(halo)
    x = np.random.default_rng(0).standard_normal(n)
    xbufs = [np.concatenate([x[r0:r0 + nl], np.full(len(h), np.nan)]) for (r0, nl), h in zip(parts, halos)]
    for r, (r0, nl) in enumerate(parts):
        rows, peers, pos = P.plan_sends(halos, parts, r)
        assert np.all(np.diff(rows) >= 0)
        for row, q, ps in zip(rows, peers, pos):
            xbufs[q][ps] = x[r0 + row]
    y = A @ x
    for (r0, nl), (ip, loc, d), xb in zip(parts, blocks, xbufs):
        assert not np.isnan(xb).any()  # every halo slot was filled by exactly the planned sends
        yl = np.add.reduceat(d * xb[loc], ip[:-1])
        assert np.abs(yl - y[r0:r0 + nl]).max() < 1e-12


def _gloo_worker(rank, world, port, root):
    import sys
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import eu_b200 as eu
    from conftest import laplacian2d
    P = eu.parallel
    A = laplacian2d(32, 40)
    n = A.shape[0]
    parts = P.row_partition(n, world)
    r0, nl = parts[rank]
    ip, ix, d = P.local_block(A, r0, nl)
    halo, loc = P.plan_halo(ix, r0, nl)
    gathered = [None] * world
    dist.all_gather_object(gathered, (r0, nl, halo))
    halos = [g[2] for g in gathered]
    rows, peers, pos = P.plan_sends(halos, [(g[0], g[1]) for g in gathered], rank)
    x = np.random.default_rng(1).standard_normal(n)
    xb = np.concatenate([x[r0:r0 + nl], np.full(len(halo), np.nan)])
    # "push": every rank publishes (peer, position, value); the peer stores them -- what the kernel does over NVLink
    msgs = [None] * world
    dist.all_gather_object(msgs, (peers, pos, x[r0 + rows]))
    for pr, ps, vals in msgs:
        sel = pr == rank
        xb[ps[sel]] = vals[sel]
    assert not np.isnan(xb).any()
    yl = np.add.reduceat(d * xb[loc], ip[:-1])
    err = float(np.abs(yl - (A @ x)[r0:r0 + nl]).max())
    # the replicated reductions: every rank sums the partials in rank order -> bitwise identical everywhere
    partial = torch.tensor([float(np.dot(yl, yl))], dtype=torch.float64)
    allp = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(allp, partial)
    total = sum(float(p) for p in allp)
    ok = err < 1e-12 and abs(total - float(np.dot(A @ x, A @ x))) < 1e-9 * total
    res = torch.tensor([1 if ok else 0])
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if int(res) != 1:
        raise SystemExit(1)


def test_halo_exchange_world2_gloo():
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, ROOT)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]


def _dense_worker(rank, world, port, n, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import eu_b200 as eu
    P = eu.parallel
    rng = np.random.default_rng(5)
    A = rng.standard_normal((n, n))
    x = rng.standard_normal(n)
    parts = P.row_partition(n, world, align=2)
    starts = np.array([p[0] for p in parts] + [n])
    r0, nl = parts[rank]
    blk = P.dense_block_in_gather_order(A[r0:r0 + nl], starts, rank)
    # the "all-gather of x" the kernel performs with peer stores: every rank writes its rows into every peer's buffer
    # at the planned positions (here: one all_gather, then each rank scatters into its own gather order)
    mx = max(p[1] for p in parts)
    mine = torch.zeros(mx, dtype=torch.float64)
    mine[:nl] = torch.from_numpy(x[r0:r0 + nl])
    outs = [torch.zeros(mx, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(outs, mine)
    xbuf = np.full(n, np.nan)
    for q, (q0, nq) in enumerate(parts):
        xbuf[P.dense_gather_position(np.arange(q0, q0 + nq), starts, rank)] = outs[q][:nq].numpy()
    y = blk @ xbuf
    ok = bool(np.allclose(y, (A @ x)[r0:r0 + nl], rtol=1e-13, atol=1e-13)) and not np.isnan(xbuf).any()
    t = torch.tensor([1 if ok else 0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put(int(t.item()))
    dist.destroy_process_group()


def test_dense_row_sharding_gather_order_gloo_world2(eu):
    """Dense row blocks: the gather-order permutation of the block's columns and the all-gather of x,
    world size 2 on CPU."""
    P = eu.parallel
    starts = np.array([0, 6, 10, 16])
    for q in range(3):
        pos = P.dense_gather_position(np.arange(16), starts, q)
        assert sorted(pos.tolist()) == list(range(16))                      # a permutation
        q0, q1 = starts[q], starts[q + 1]
        assert pos[q0:q1].tolist() == list(range(q1 - q0))                  # own entries first
        rest = np.concatenate([pos[:q0], pos[q1:]])
        assert rest.tolist() == list(range(q1 - q0, 16))                    # the others in ascending global order
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dense_worker, args=(r, 2, port, 64, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs)
    assert ret.get(timeout=5) == 1


def test_host_binding_helpers(eu, monkeypatch):
    """parallel.bind_host_near_gpu: sysfs cpulist parsing; unknown NUMA node (no GPU here, or a single-socket host that
    reports -1) is a recorded no-op that leaves the affinity mask alone; a known node narrows the mask to its CPUs."""
    import os
    P = eu.parallel
    assert P._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert P._parse_cpulist("") == set() and P._parse_cpulist("5") == {5}
    before = os.sched_getaffinity(0)
    rec = P.bind_host_near_gpu(0)
    assert rec == {"numa_node": None, "bound": False} and os.sched_getaffinity(0) == before
    one = {sorted(before)[0]}
    monkeypatch.setattr(P, "gpu_numa_node", lambda d: (1, one | {10**6}))
    try:
        rec = P.bind_host_near_gpu(0)
        assert rec["bound"] and rec["numa_node"] == 1 and rec["cpus"] == 1 and os.sched_getaffinity(0) == one
        monkeypatch.setattr(P, "gpu_numa_node", lambda d: (1, {10**6}))   # no overlap with the allowed CPUs: untouched
        assert not P.bind_host_near_gpu(0)["bound"] and os.sched_getaffinity(0) == one
    finally:
        os.sched_setaffinity(0, before)
