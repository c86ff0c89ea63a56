"""Row sharding of one Krylov problem across the GPUs of a node (one process per GPU).

Host-side plumbing only: partition rows, work out which remote entries each rank gathers (halo) and which of
its own rows other ranks gather (send list), exchange the 64-byte CUDA IPC handles with torch.distributed, and
hand everything to the C ABI (b200k_comm_*, b200k_op_csr_create_sharded).  The per-step halo exchange and the
all-reduce of the inner products happen inside the persistent kernel over peer-mapped memory; there is no
collective call on the data path.  The planning functions are pure NumPy and are tested on CPU (gloo).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(device_index: int):
    """(NUMA node, its CPUs) of the PCIe root the GPU hangs off, from sysfs; (None, set()) when the platform does not
    say (single-socket hosts report -1)."""
    import torch
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read())
        if node < 0:
            return None, set()
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            return node, _parse_cpulist(f.read())
    except Exception:
        return None, set()


def bind_host_near_gpu(device_index: int):
    """One process per GPU: run this rank's host threads on the socket its GPU is attached to, BEFORE the pinned
    staging buffers are allocated (they are placed on the allocating thread's node), so that every H2D / D2H copy of the
    end-to-end path crosses one PCIe root and no socket interconnect.  Without it the ranks whose GPU sits on the other
    socket copy through remote memory and set the max-over-ranks time of a multi-GPU job.  Returns a small record for
    the bench line; a no-op (and says so) when the node is unknown or the affinity mask cannot be narrowed."""
    import os
    node, cpus = gpu_numa_node(device_index)
    rec = {"numa_node": node, "bound": False}
    if node is None:
        return rec
    try:
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            rec["bound"] = True
            rec["cpus"] = len(allowed)
    except Exception:
        pass
    return rec


def row_partition(n: int, world: int, align: int = 16):
    """Contiguous, `align`-row-aligned blocks: returns [(row0, nloc)] for every rank."""
    base = (n // world) // align * align
    out, r0 = [], 0
    for r in range(world):
        nloc = base if r < world - 1 else n - r0
        out.append((r0, nloc))
        r0 += nloc
    return out


def shard_batch(nb: int, rank: int, world: int):
    """Independent problems i = lo..hi-1 of a batch that `rank` owns (embarrassingly parallel replicas)."""
    lo = nb * rank // world
    hi = nb * (rank + 1) // world
    return lo, hi


def local_block(A, row0: int, nloc: int):
    """Rows [row0, row0+nloc) of a scipy.sparse matrix as CSR arrays (global column indices)."""
    B = A.tocsr()[row0:row0 + nloc]
    return (np.ascontiguousarray(B.indptr, dtype=np.int32), np.ascontiguousarray(B.indices, dtype=np.int64),
            np.ascontiguousarray(B.data, dtype=np.float64))


def plan_halo(indices: np.ndarray, row0: int, nloc: int):
    """Sorted unique remote columns this rank gathers, and the column indices rewritten to local gather
    positions: own columns -> [0, nloc), remote column halo[k] -> nloc + k."""
    indices = np.asarray(indices, dtype=np.int64)
    remote = (indices < row0) | (indices >= row0 + nloc)
    halo = np.unique(indices[remote])
    loc = np.empty(indices.shape, dtype=np.int32)
    loc[~remote] = (indices[~remote] - row0).astype(np.int32)
    loc[remote] = (nloc + np.searchsorted(halo, indices[remote])).astype(np.int32)
    return halo, loc


def plan_sends(all_halos, all_ranges, rank: int):
    """Send list of `rank`: for every own row that rank q gathers, (local row, q, position in q's gather buffer).
    Sorted by local row (the kernel hands contiguous ranges of it to the CTAs that own those rows)."""
    row0, nloc = all_ranges[rank]
    rows, peers, pos = [], [], []
    for q, halo in enumerate(all_halos):
        if q == rank or len(halo) == 0:
            continue
        halo = np.asarray(halo, dtype=np.int64)
        mask = (halo >= row0) & (halo < row0 + nloc)
        idx = np.nonzero(mask)[0]
        rows.append(halo[idx] - row0)
        peers.append(np.full(idx.size, q, dtype=np.int64))
        pos.append(all_ranges[q][1] + idx)
    if not rows:
        z = np.zeros(0, dtype=np.int32)
        return z, z.copy(), z.copy()
    rows, peers, pos = np.concatenate(rows), np.concatenate(peers), np.concatenate(pos)
    order = np.lexsort((peers, rows))
    return rows[order].astype(np.int32), peers[order].astype(np.int32), pos[order].astype(np.int32)


class ShardedOperator:
    """An `Operator` whose rows are split across the ranks of a torch.distributed group.

    ``op`` behaves like any other operator for arnoldi_/expv/phiv/kiops, but every vector handed to those calls
    is this rank's row block, and all ranks must make the same calls in the same order (the kernels of all
    GPUs synchronise with each other inside every Krylov step)."""

    def __init__(self, A_block, row0: int, n_global: int, *, ishermitian: bool, group=None, engine=None):
        import torch
        import torch.distributed as dist
        from .api import Operator, get_engine

        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        eng = engine or get_engine()
        self.engine = eng
        indptr, indices, data = A_block
        nloc = indptr.size - 1
        self.row0, self.nloc, self.n_global = int(row0), int(nloc), int(n_global)
        halo, loc = plan_halo(indices, row0, nloc)
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (int(row0), int(nloc), halo), group=group)
        ranges = [(g[0], g[1]) for g in gathered]
        halos = [g[2] for g in gathered]
        srow, speer, spos = plan_sends(halos, ranges, self.rank)
        xlen = max(r[1] + len(hh) for r, hh in zip(ranges, halos)) + 16
        lib = eng.lib
        eng.bind_stream()
        handle = (C.c_ubyte * 64)()
        comm = C.c_void_p()
        eng.check(lib.b200k_comm_create(eng.handle, self.rank, self.world, xlen, handle, C.byref(comm)))
        hs = [None] * self.world
        dist.all_gather_object(hs, bytes(handle), group=group)
        allh = (C.c_ubyte * (64 * self.world)).from_buffer_copy(b"".join(hs))
        eng.check(lib.b200k_comm_connect(comm, allh))
        dist.barrier(group=group)
        self.comm = comm
        ptr = C.c_void_p()
        eng.check(lib.b200k_op_csr_create_sharded(
            eng.handle, comm, nloc, len(halo), data.size, indptr.ctypes.data, loc.ctypes.data, data.ctypes.data,
            0, 1, 1 if ishermitian else 0, srow.size, srow.ctypes.data, speer.ctypes.data, spos.ctypes.data,
            C.byref(ptr)))
        self.op = Operator(eng, ptr, (comm,))
        self.nhalo = len(halo)
        self.nsend = int(srow.size)

    def close(self):
        if self.comm:
            self.op = None
            self.engine.lib.b200k_comm_destroy(self.comm)
            self.comm = None


def dense_gather_position(g, row_starts, q: int):
    """Position of global row/column ``g`` in rank ``q``'s gather buffer for a row-sharded DENSE operator: the rank's own
    entries come first, every other row follows in ascending global order (the layout b200k_op_dense_create_sharded
    permutes the block's columns into and builds its send list for)."""
    g = np.asarray(g, dtype=np.int64)
    q0, q1 = int(row_starts[q]), int(row_starts[q + 1])
    nq = q1 - q0
    return np.where((g >= q0) & (g < q1), g - q0, np.where(g < q0, nq + g, g))


def dense_block_in_gather_order(A_block, row_starts, q: int):
    """Columns of rank q's row block permuted into its gather order (host mirror of what the library does at ingestion)."""
    n = int(row_starts[-1])
    pos = dense_gather_position(np.arange(n), row_starts, q)
    out = np.empty_like(A_block)
    out[:, pos] = A_block
    return out


class ShardedDenseOperator:
    """A dense operator whose rows are split across the ranks (BASELINE config 3 beyond one GPU).  ``A_block`` is this
    rank's (nloc x n) row block (NumPy array or CUDA tensor, columns in global order); every rank passes the same
    ``row_starts`` (length world + 1, even offsets, e.g. from ``row_partition``).  ``op`` is used like any operator
    with this rank's row block of every vector; per Krylov step every rank pushes its block of the new vector to all
    peers inside the persistent kernel (the all-gather of x) -- no collective call on the data path."""

    def __init__(self, A_block, row_starts, *, ishermitian: bool, group=None, engine=None):
        import torch
        import torch.distributed as dist
        from .api import Operator, get_engine

        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        eng = engine or get_engine()
        self.engine = eng
        rs = np.ascontiguousarray(np.asarray(row_starts, dtype=np.int64))
        n = int(rs[-1])
        self.row0, self.nloc, self.n_global = int(rs[self.rank]), int(rs[self.rank + 1] - rs[self.rank]), n
        lib = eng.lib
        eng.bind_stream()
        handle = (C.c_ubyte * 64)()
        comm = C.c_void_p()
        eng.check(lib.b200k_comm_create(eng.handle, self.rank, self.world, n + 32, handle, C.byref(comm)))
        hs = [None] * self.world
        dist.all_gather_object(hs, bytes(handle), group=group)
        allh = (C.c_ubyte * (64 * self.world)).from_buffer_copy(b"".join(hs))
        eng.check(lib.b200k_comm_connect(comm, allh))
        dist.barrier(group=group)
        self.comm = comm
        ptr = C.c_void_p()
        if isinstance(A_block, torch.Tensor):
            if A_block.shape != (self.nloc, n):
                raise _lib.DimensionMismatch("A_block must be nloc x n")
            At = A_block.to(device=eng.device, dtype=torch.float64).t().contiguous()  # row j = column j of the block
            eng.check(lib.b200k_op_dense_create_sharded(eng.handle, comm, n, rs.ctypes.data, C.c_void_p(At.data_ptr()),
                                                        self.nloc, 0, 1 if ishermitian else 0, C.byref(ptr)))
            eng.synchronize()
            del At
        else:
            Af = np.asfortranarray(np.asarray(A_block, dtype=np.float64))
            if Af.shape != (self.nloc, n):
                raise _lib.DimensionMismatch("A_block must be nloc x n")
            eng.check(lib.b200k_op_dense_create_sharded(eng.handle, comm, n, rs.ctypes.data, Af.ctypes.data, self.nloc, 1,
                                                        1 if ishermitian else 0, C.byref(ptr)))
        self.op = Operator(eng, ptr, (comm,))

    def close(self):
        if self.comm:
            self.op = None
            self.engine.lib.b200k_comm_destroy(self.comm)
            self.comm = None


def kiops_sharded(tau_out, sop: ShardedOperator, u_local, **kw):
    """kiops on a row-sharded operator: u_local is this rank's row block of u (nloc x (p+1)); returns this rank's
    rows of w and the (replicated) stats."""
    import torch
    import torch.distributed as dist
    from .api import kiops
    ul = torch.as_tensor(u_local)
    if ul.dim() == 1:
        ul = ul.reshape(-1, 1)
    normU = None
    if ul.shape[1] > 1:
        t = ul[:, 1:].abs().sum().to(dtype=torch.float64, device=sop.engine.device).reshape(1)
        dist.all_reduce(t, group=sop.group)
        normU = float(t.item())
    return kiops(tau_out, sop.op, u_local, _normU=normU, **kw)  # pass return_device=True to keep w on the GPU
