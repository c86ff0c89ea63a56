"""exponentialutilities.jl_b200 -- B200-native Krylov expmv/phiv engine behind the
ExponentialUtilities.jl API surface (arnoldi!/lanczos! + expv/phiv + kiops).

The directory name contains a dot, so import it through the repository-root shim::

    import eu_b200 as eu
    w = eu.expv(1.0, A, b, m=30)

The product is ``libb200krylov.so`` (hand-written sm_100a CUDA behind the C ABI in
``include/b200krylov.h``); this package is the thin ctypes host layer mirroring the reference's
function names, keywords and error behaviour.
"""
from . import _lib, build, parallel  # noqa: F401
from ._lib import (ArgumentError, DimensionMismatch, SingularException, UnsupportedError,  # noqa: F401
                   lib_path, load)
from .api import (Engine, ExpvCache, KrylovSubspace, Operator, PhivCache, arnoldi, arnoldi_, expv, expv_, expv_batched,  # noqa: F401
                  expv_host, expv_host_async, expv_small, expv_timestep, exponential_, exponential_batched_, get_engine, kiops, lanczos_, operator, phiv, phiv_, phiv_dense, phiv_timestep)

__all__ = [
    "Engine", "KrylovSubspace", "Operator", "ExpvCache", "PhivCache", "arnoldi", "arnoldi_", "lanczos_", "expv", "expv_", "expv_batched",
    "expv_host", "expv_host_async", "expv_small", "expv_timestep", "phiv_timestep", "phiv", "phiv_", "kiops", "exponential_", "exponential_batched_", "phiv_dense", "operator", "get_engine",
    "DimensionMismatch", "ArgumentError", "SingularException", "UnsupportedError", "load", "lib_path",
]
