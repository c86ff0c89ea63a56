"""ctypes binding of include/b200krylov.h (the C ABI is the product boundary)."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)
c_float_p = C.POINTER(C.c_float)

# status codes (include/b200krylov.h)
OK, EDIM, EARG, ESINGULAR, ECUDA, ECOMM, EUNSUPPORTED, ENOMEM = range(8)


class KrylovOpts(C.Structure):
    _fields_ = [
        ("m", C.c_int), ("tol", C.c_double), ("iop", C.c_int), ("hermitian", C.c_int),
        ("init", C.c_int), ("p", C.c_int), ("B", C.c_void_p), ("ldb", C.c_int64),
        ("t", C.c_double), ("mu", C.c_double),
    ]


class KiopsOpts(C.Structure):
    _fields_ = [
        ("mmin", C.c_int), ("mmax", C.c_int), ("m", C.c_int), ("tol", C.c_double),
        ("iop", C.c_int), ("hermitian", C.c_int), ("task1", C.c_int), ("opnorm", C.c_double),
        ("normU", C.c_double),
    ]


class TimestepOpts(C.Structure):
    _fields_ = [
        ("tau", C.c_double), ("m", C.c_int), ("tol", C.c_double), ("opnorm", C.c_double), ("iop", C.c_int),
        ("correct", C.c_int), ("adaptive", C.c_int), ("delta", C.c_double), ("hermitian", C.c_int),
        ("gamma", C.c_double), ("NA", C.c_int64),
    ]


# Every symbol include/b200krylov.h declares: (restype, argtypes)
PROTOTYPES = {
    "b200k_version": (C.c_int, []),
    "b200k_sizeof": (C.c_int, [C.c_int]),
    "b200k_status_string": (C.c_char_p, [C.c_int]),
    "b200k_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_void_p]),
    "b200k_destroy": (C.c_int, [C.c_void_p]),
    "b200k_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b200k_synchronize": (C.c_int, [C.c_void_p]),
    "b200k_last_error": (C.c_char_p, [C.c_void_p]),
    "b200k_device_info": (C.c_int, [C.c_void_p, c_int_p, c_int_p, c_int64_p]),
    "b200k_op_csr_create": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "b200k_op_dense_create": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int,
                                        C.POINTER(C.c_void_p)]),
    "b200k_op_destroy": (C.c_int, [C.c_void_p]),
    "b200k_op_info": (C.c_int, [C.c_void_p, c_int64_p, c_int64_p, c_int_p, c_int_p, c_double_p]),
    "b200k_op_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200k_krylov_opts_default": (None, [C.POINTER(KrylovOpts)]),
    "b200k_kiops_opts_default": (None, [C.POINTER(KiopsOpts)]),
    "b200k_arnoldi": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(KrylovOpts), C.c_void_p,
                                C.c_int64, C.c_int, c_double_p, C.c_int, c_double_p, c_int_p, c_int_p]),
    "b200k_expv_ks": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_int64, C.c_int64, c_double_p,
                                C.c_int, C.c_int, C.c_double, C.c_void_p]),
    "b200k_phiv_ks": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_int64, C.c_int64, c_double_p,
                                C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_int64,
                                c_double_p]),
    "b200k_expv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.POINTER(KrylovOpts),
                             C.c_void_p, c_int_p, c_int_p, c_double_p]),
    "b200k_expv_ee": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_int, C.c_double, C.c_double,
                                C.c_void_p, c_int_p]),
    "b200k_expv_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.POINTER(KrylovOpts),
                                  C.c_void_p, c_int_p, c_int_p]),
    "b200k_expv_host_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.POINTER(KrylovOpts),
                                        C.c_void_p]),
    "b200k_phiv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_int,
                             C.POINTER(KrylovOpts), C.c_int, C.c_void_p, C.c_int64, c_double_p, c_int_p,
                             c_int_p]),
    "b200k_expv_batched": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, c_double_p, C.c_void_p, C.c_int64,
                                     C.POINTER(KrylovOpts), C.c_void_p, C.c_int64, c_int_p, c_int_p]),
    "b200k_kiops": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, c_double_p, C.c_int, C.c_void_p, C.c_int64,
                              C.c_int, C.POINTER(KiopsOpts), C.c_void_p, C.c_int64, c_int64_p]),
    "b200k_op_csr_create_z": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "b200k_op_dense_create_z": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int,
                                          C.POINTER(C.c_void_p)]),
    "b200k_arnoldi_z": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(KrylovOpts), C.c_void_p,
                                  C.c_int64, C.c_int, C.c_void_p, C.c_int, c_double_p, c_int_p, c_int_p]),
    "b200k_expv_ks_z": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                  C.c_int, C.c_int, C.c_double, C.c_void_p]),
    "b200k_phiv_ks_z": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                  C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_int64, c_double_p]),
    "b200k_expv_z": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.POINTER(KrylovOpts),
                               C.c_void_p, c_int_p, c_int_p]),
    "b200k_expv_small_z": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p, c_int_p]),
    "b200k_project": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_double, c_double_p, C.c_int,
                                C.c_int, C.c_void_p, C.c_int64]),
    "b200k_timestep_opts_default": (None, [C.POINTER(TimestepOpts)]),
    "b200k_phiv_timestep": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, c_double_p, C.c_void_p, C.c_int64, C.c_int,
                                      C.POINTER(TimestepOpts), C.c_void_p, C.c_int64, c_int_p]),
    "b200k_comm_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.POINTER(C.c_void_p)]),
    "b200k_comm_connect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b200k_comm_destroy": (C.c_int, [C.c_void_p]),
    "b200k_op_csr_create_sharded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "b200k_op_dense_create_sharded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                                C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "b200k_exponential_batched": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int64]),
    "b200k_exponential": (C.c_int, [C.c_int, c_double_p, C.c_int]),
    "b200k_expv_small": (C.c_int, [C.c_int, c_double_p, C.c_int, C.c_double, c_double_p, c_int_p]),
    "b200k_phiv_dense": (C.c_int, [C.c_int, c_double_p, C.c_int, c_double_p, C.c_int, c_double_p, C.c_int]),
    "b200k_last_timing": (C.c_int, [C.c_void_p, c_float_p, c_float_p]),
    "b200k_set_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "b200k_last_kernel": (C.c_int, [C.c_void_p, c_int_p]),
    "b200k_set_flag": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
}

_lib = None


def lib_path() -> str:
    return _build.LIB


def load():
    """Load libb200krylov.so (building it first if the sources are newer).  Fails loudly: there is
    no Python / CPU fallback for the product path."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("B200K_LIB")  # A/B measurements against an older build of the library
    if not path:
        path = _build.LIB
        if _build.is_stale():
            path = _build.build()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run __graft_entry__.build() (nvcc, sm_100a)")
    lib = C.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        if os.environ.get("B200K_LIB") and not hasattr(lib, name):
            continue  # older build used for an A/B measurement
        fn = getattr(lib, name)  # AttributeError if the ABI drifted
        fn.restype = res
        fn.argtypes = args
    if hasattr(lib, "b200k_sizeof"):  # the binding's struct definitions must match the library's (ABI guard)
        for which, cls in ((1, KrylovOpts), (2, KiopsOpts), (3, TimestepOpts)):
            if lib.b200k_sizeof(which) != C.sizeof(cls):
                raise RuntimeError(f"ABI mismatch: {cls.__name__} is {C.sizeof(cls)} bytes here, "
                                   f"{lib.b200k_sizeof(which)} in {path}")
    _lib = lib
    return lib


class DimensionMismatch(ValueError):
    """Mirror of Julia's DimensionMismatch / the "Dimension mismatch" asserts."""


class ArgumentError(ValueError):
    """Mirror of Julia's ArgumentError."""


class SingularException(ArithmeticError):
    """Mirror of LinearAlgebra.SingularException."""


class UnsupportedError(NotImplementedError):
    pass


def check(status: int, handle=None):
    if status == OK:
        return
    lib = load()
    msg = lib.b200k_status_string(status).decode()
    if handle:
        detail = lib.b200k_last_error(handle).decode()
        if detail:
            msg = f"{msg}: {detail}"
    if status == EDIM:
        raise DimensionMismatch(msg)
    if status == EARG:
        raise ArgumentError(msg)
    if status == ESINGULAR:
        raise SingularException(msg)
    if status == EUNSUPPORTED:
        raise UnsupportedError(msg)
    if status == ENOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)
