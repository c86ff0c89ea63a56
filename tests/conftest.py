import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests/` on a box without a CUDA device skips the gpu-marked tests instead of failing them.
    (On a GPU box nothing is skipped: there the product must load its CUDA library or fail loudly.)"""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device on this box (GPU parity tests run with -m gpu on a B200)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def laplacian2d(nx, ny):
    """2-D 5-point Laplacian (Dirichlet, stencil -4 / +1) on an nx x ny grid, CSR int32."""
    import scipy.sparse as sp
    ex, ey = np.ones(nx), np.ones(ny)
    Tx = sp.diags([ex[:-1], -2 * ex, ex[:-1]], [-1, 0, 1])
    Ty = sp.diags([ey[:-1], -2 * ey, ey[:-1]], [-1, 0, 1])
    A = (sp.kron(sp.identity(ny), Tx) + sp.kron(Ty, sp.identity(nx))).tocsr()
    A.indices = A.indices.astype(np.int32)
    A.indptr = A.indptr.astype(np.int32)
    return A


def convdiff2d(nx, ny, c=0.3):
    """Non-symmetric convection-diffusion: Laplacian + c * (0.5 on the +1 diagonal, -0.5 on the -1)."""
    import scipy.sparse as sp
    n = nx * ny
    A = laplacian2d(nx, ny)
    Cm = sp.diags([-0.5 * np.ones(n - 1), 0.5 * np.ones(n - 1)], [-1, 1])
    return (A + c * Cm).tocsr()


def relerr(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    if not (np.iscomplexobj(a) or np.iscomplexobj(b)):
        a = a.astype(float)
        b = b.astype(float)
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)


@pytest.fixture(scope="session")
def eu():
    import eu_b200
    return eu_b200


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    return O
