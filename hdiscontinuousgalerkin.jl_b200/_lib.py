"""ctypes binding of libhdg_b200.so (include/hdg_b200.h).

The library is built in-tree (csrc/Makefile -> lib/libhdg_b200.so) and holds the hand-written
sm_100a CUDA kernels.  There is no CPU fallback: if the shared library is missing, or no CUDA
device is present, every compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HDG_B200_LIB", os.path.join(_HERE, "lib", "libhdg_b200.so"))   # override: kernel-variant experiments

HDG_OK = 0
STATUS_NAMES = {
    0: "HDG_OK", 1: "HDG_ERR_INVALID", 2: "HDG_ERR_BAD_GEOMETRY", 3: "HDG_ERR_UNSUPPORTED_RULE",
    4: "HDG_ERR_SINGULAR_LOCAL", 5: "HDG_ERR_CUDA", 6: "HDG_ERR_NCCL", 7: "HDG_ERR_NOT_CONVERGED",
    8: "HDG_ERR_NOT_BOUNDARY",
}


class HDGError(RuntimeError):
    """Base class; `.status` holds the hdg_status code."""

    def __init__(self, status, msg):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {msg}")
        self.status = status


class BadGeometryError(HDGError, ValueError):      # ArgumentError, src/ScalarFunctionSpaces.jl:110
    pass


class UnsupportedRuleError(HDGError, ValueError):  # ArgumentError, src/quadrature.jl:24
    pass


class SingularLocalError(HDGError, ArithmeticError):  # LAPACK SingularException
    pass


class NotBoundaryError(HDGError, AssertionError):  # @assert src/boundary.jl:22
    pass


class NotConvergedError(HDGError):
    pass


_ERR_CLASS = {2: BadGeometryError, 3: UnsupportedRuleError, 4: SingularLocalError, 7: NotConvergedError,
              8: NotBoundaryError}


class Params(C.Structure):
    _fields_ = [("order", C.c_int32), ("quad_degree", C.c_int32), ("tau", C.c_double),
                ("source_id", C.c_int32), ("device", C.c_int32), ("local_solver", C.c_int32),
                ("reserved", C.c_int32)]


class Sizes(C.Structure):
    _fields_ = [("ncell", C.c_int64), ("nnode", C.c_int64), ("nface", C.c_int64), ("nbface", C.c_int64),
                ("n", C.c_int32), ("nt", C.c_int32), ("m", C.c_int32), ("t", C.c_int32),
                ("nq", C.c_int32), ("nfq", C.c_int32), ("ndof", C.c_int64), ("nnz", C.c_int64)]


class SolveInfo(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("converged", C.c_int32), ("relres", C.c_double),
                ("bnorm", C.c_double), ("solve_ms", C.c_double)]


_P = C.c_void_p
_I64P = C.POINTER(C.c_int64)
_F64P = C.POINTER(C.c_double)

# name -> (restype, argtypes); every symbol include/hdg_b200.h declares
SIGNATURES = {
    "hdg_create": (C.c_int, [C.POINTER(Params), C.POINTER(_P)]),
    "hdg_destroy": (None, [_P]),
    "hdg_last_error": (C.c_char_p, [_P]),
    "hdg_version": (C.c_char_p, []),
    "hdg_set_mesh": (C.c_int, [_P, _I64P, C.c_int64, _F64P, C.c_int64, _I64P, C.c_int64, _I64P, C.c_int64]),
    "hdg_set_rectangle_mesh": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double]),
    "hdg_number_faces": (C.c_int, [_P, _I64P, C.c_int64, _F64P, C.c_int64, _I64P, _I64P, C.c_int64, _I64P]),
    "hdg_order_cells": (C.c_int, [_P, _I64P, C.c_int64, _F64P, C.c_int64, _I64P]),
    "hdg_perturb_nodes": (C.c_int, [_P, C.c_double, C.c_uint64]),
    "hdg_get_sizes": (C.c_int, [_P, C.POINTER(Sizes)]),
    "hdg_get_mesh": (C.c_int, [_P, _I64P, _F64P, _I64P, _I64P]),
    "hdg_get_table": (C.c_int, [_P, C.c_char_p, _F64P, _I64P]),
    "hdg_ref_table": (C.c_int, [C.c_int32, C.c_int32, C.c_char_p, _F64P, _I64P]),
    "hdg_set_source_values": (C.c_int, [_P, _F64P]),
    "hdg_assemble": (C.c_int, [_P]),
    "hdg_apply_dirichlet": (C.c_int, [_P, _F64P]),
    "hdg_solve": (C.c_int, [_P, C.c_double, C.c_int32, C.POINTER(SolveInfo)]),
    "hdg_set_preconditioner": (C.c_int, [_P, C.c_int32]),
    "hdg_recover": (C.c_int, [_P]),
    "hdg_errornorm": (C.c_int, [_P, C.c_int32, _F64P]),
    "hdg_basis_value": (C.c_int, [C.c_int32, C.c_int32, _F64P, _F64P, _F64P]),
    "hdg_errornorm_values": (C.c_int, [_P, _F64P, _F64P]),
    "hdg_set_dirichlet_faces": (C.c_int, [_P, _I64P, C.c_int64]),
    "hdg_assemble_async": (C.c_int, [_P]),
    "hdg_sync": (C.c_int, [_P]),
    "hdg_stream": (C.c_uint64, [_P]),
    "hdg_get_pattern": (C.c_int, [_P, _I64P, _I64P]),
    "hdg_get_values": (C.c_int, [_P, _F64P]),
    "hdg_get_rhs": (C.c_int, [_P, _F64P]),
    "hdg_get_trace": (C.c_int, [_P, _F64P]),
    "hdg_set_trace": (C.c_int, [_P, _F64P]),
    "hdg_get_meandiag": (C.c_int, [_P, _F64P]),
    "hdg_get_local": (C.c_int, [_P, C.c_int64, _F64P, _F64P]),
    "hdg_get_condensed": (C.c_int, [_P, C.c_int64, _F64P, _F64P]),
    "hdg_get_mvalues": (C.c_int, [_P, _F64P, _F64P, _F64P]),
    "hdg_nodal_avg": (C.c_int, [_P, _F64P]),
    "hdg_comm_unique_id": (C.c_int, [C.POINTER(C.c_uint8)]),
    "hdg_comm_init": (C.c_int, [_P, C.c_int32, C.c_int32, C.POINTER(C.c_uint8)]),
    "hdg_get_partition": (C.c_int, [_P, _I64P]),
    "hdg_get_ghost_cells": (C.c_int, [_P, _I64P]),
    "hdg_comm_pingpong": (C.c_int, [_P, C.c_int32, _F64P]),
    "hdg_last_phase_ms": (C.c_int, [_P, C.c_char_p, _F64P]),
    "hdg_measure_fp64_peak": (C.c_int, [_P, _F64P]),
    "hdg_mg_trace": (C.c_int32, [_P, _F64P, C.c_int32]),
    "hdg_cg_setup": (C.c_int, [_P, C.c_int32, _I64P]),
    "hdg_cg_get_sizes": (C.c_int, [_P, _I64P]),
    "hdg_cg_get_dofhandler": (C.c_int, [_P, _I64P, _I64P, _I64P]),
    "hdg_cg_assemble": (C.c_int, [_P]),
    "hdg_cg_apply_dirichlet": (C.c_int, [_P]),
    "hdg_cg_solve": (C.c_int, [_P, C.c_double, C.c_int32, C.POINTER(SolveInfo)]),
    "hdg_cg_get_system": (C.c_int, [_P, _F64P, _F64P, _F64P]),
    "hdg_cg_errornorm": (C.c_int, [_P, _F64P]),
    "hdg_cg_get_meandiag": (C.c_int, [_P, _F64P]),
    "hdg_launch_count": (C.c_int64, [_P]),
}

_lib = None


def load():
    """Load libhdg_b200.so (once).  Raises if it has not been built - no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HDGError(5, f"{LIB_PATH} not found: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
                          "(or __graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status, ctx=None):
    if status == HDG_OK:
        return
    msg = load().hdg_last_error(ctx)
    msg = msg.decode() if msg else ""
    raise _ERR_CLASS.get(status, HDGError)(status, msg)


def f64p(a):
    return a.ctypes.data_as(_F64P) if a is not None else None


def i64p(a):
    return a.ctypes.data_as(_I64P) if a is not None else None
