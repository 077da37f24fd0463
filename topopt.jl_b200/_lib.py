"""ctypes binding of libtopopt_cuda.so -- a 1:1 image of include/topopt_cuda.h.

There is no CPU fallback: if the shared library is missing this module raises, and
``topopt_create`` fails with TOPOPT_ERR_NO_DEVICE when no GPU is visible.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtopopt_cuda.so")

OK = 0
ERR_INVALID, ERR_CUDA, ERR_NCCL, ERR_NONFINITE, ERR_NO_DEVICE, ERR_MISMATCH = -1, -2, -3, -4, -5, -6
PENALTY_POWER, PENALTY_RATIONAL, PENALTY_SINH = 0, 1, 2
OP_MATRIX_FREE, OP_ASSEMBLED = 0, 1
PRECOND_NONE, PRECOND_JACOBI, PRECOND_MULTIGRID = 0, 1, 2
CRITERIA_DEFAULT, CRITERIA_ENERGY = 0, 1
CG_REFERENCE, CG_SINGLE_PASS = 0, 1
PHYSICS_ELASTICITY, PHYSICS_HEAT = 0, 1
FILTER_FORWARD, FILTER_TRANSPOSE = 0, 1

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int64)


class Desc(C.Structure):
    _fields_ = [
        ("dim", C.c_int32),
        ("ncomp", C.c_int32),
        ("nels", C.c_int64 * 3),
        ("sizes", C.c_double * 3),
        ("Ke", c_dp),
        ("prescribed_dofs", c_ip),
        ("n_prescribed", C.c_int64),
        ("fixedload", c_dp),
        ("cellvolumes", c_dp),
        ("cell_dofs", c_ip),
        ("fixed_diag", C.c_double),
        ("device", C.c_int32),
        ("rank", C.c_int32),
        ("world", C.c_int32),
        ("nccl_unique_id", C.c_void_p),
    ]


class CGOpts(C.Structure):
    _fields_ = [
        ("abstol", C.c_double),
        ("reltol", C.c_double),
        ("maxiter", C.c_int32),
        ("op", C.c_int32),
        ("precond", C.c_int32),
        ("criteria", C.c_int32),
        ("check_every", C.c_int32),
        ("variant", C.c_int32),
        ("warm_start", C.c_int32),
        ("refresh_precond", C.c_int32),
        ("mg_degree", C.c_int32),
        ("reserved0", C.c_int32),
        ("mg_ratio", C.c_double),
    ]


class CGResult(C.Structure):
    _fields_ = [
        ("iters", C.c_int32),
        ("converged", C.c_int32),
        ("residual", C.c_double),
        ("tol", C.c_double),
        ("solve_ms", C.c_double),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("ndof", C.c_int64),
        ("nel", C.c_int64),
        ("nnodes", C.c_int64),
        ("nnz", C.c_int64),
        ("ndof_local", C.c_int64),
        ("nel_local", C.c_int64),
        ("kernel_launches", C.c_int64),
        ("cg_iterations", C.c_int64),
        ("h2d_bytes", C.c_int64),
        ("d2h_bytes", C.c_int64),
        ("last_apply_ms", C.c_double),
        ("last_solve_ms", C.c_double),
        ("last_sens_ms", C.c_double),
        ("last_filter_ms", C.c_double),
    ]


# every symbol include/topopt_cuda.h declares: name -> (restype, argtypes)
VP = C.c_void_p
SIGNATURES = {
    "topopt_version": (C.c_char_p, []),
    "topopt_last_error": (C.c_char_p, [VP]),
    "topopt_sizes": (C.c_int, [C.c_int32, C.c_int32, c_ip, c_ip, c_ip, c_ip, c_ip]),
    "topopt_node_dofs": (C.c_int, [C.c_int32, C.c_int32, c_ip, c_ip]),
    "topopt_cell_dofs": (C.c_int, [C.c_int32, C.c_int32, c_ip, c_ip]),
    "topopt_cells": (C.c_int, [C.c_int32, c_ip, c_ip]),
    "topopt_csc_pattern": (C.c_int, [C.c_int32, C.c_int32, c_ip, c_ip, c_ip]),
    "topopt_element_matrix": (C.c_int, [C.c_int32, C.c_int32, c_dp, C.c_double, C.c_double, C.c_int32, c_dp]),
    "topopt_nccl_unique_id": (C.c_int, [VP]),
    "topopt_ipc_export": (C.c_int, [VP, VP, c_ip]),
    "topopt_ipc_import": (C.c_int, [VP, VP, C.c_int64]),
    "topopt_create": (C.c_int, [C.POINTER(Desc), C.POINTER(VP)]),
    "topopt_destroy": (C.c_int, [VP]),
    "topopt_get_stats": (C.c_int, [VP, C.POINTER(Stats)]),
    "topopt_reset_stats": (C.c_int, [VP]),
    "topopt_set_density": (C.c_int, [VP, VP, C.c_int32, C.c_double, C.c_double, C.c_int32]),
    "topopt_set_projection": (C.c_int, [VP, C.c_int32, C.c_double]),
    "topopt_set_stiffness": (C.c_int, [VP, VP, VP]),
    "topopt_get_stiffness": (C.c_int, [VP, VP, VP]),
    "topopt_apply": (C.c_int, [VP, VP, VP]),
    "topopt_apply_ex": (C.c_int, [VP, VP, VP, C.c_int32, c_dp]),
    "topopt_assemble": (C.c_int, [VP, VP, VP]),
    "topopt_spmv": (C.c_int, [VP, VP, VP]),
    "topopt_set_jacobi": (C.c_int, [VP, VP]),
    "topopt_solve": (C.c_int, [VP, VP, VP, C.POINTER(CGOpts), C.POINTER(CGResult)]),
    "topopt_compliance": (C.c_int, [VP, VP, c_dp, VP, VP]),
    "topopt_bilinear_sens": (C.c_int, [VP, VP, VP, VP, VP]),
    "topopt_swap_solution_lambda": (C.c_int, [VP]),
    "topopt_dot": (C.c_int, [VP, VP, VP, c_dp]),
    "topopt_quadratic_form": (C.c_int, [VP, VP, c_dp]),
    "topopt_filter_create": (C.c_int, [VP, C.c_double, C.POINTER(VP)]),
    "topopt_filter_apply": (C.c_int, [VP, VP, VP, C.c_int32]),
    "topopt_filter_destroy": (C.c_int, [VP]),
    "topopt_simp_eval": (
        C.c_int,
        [VP, VP, C.c_int32, VP, C.c_int32, C.c_double, C.c_double, C.POINTER(CGOpts), c_dp, VP, C.POINTER(CGResult)],
    ),
    "topopt_oc_update": (C.c_int, [VP, VP, VP, VP, C.c_double, C.c_double, C.c_double, C.c_double, VP, c_dp, C.POINTER(C.c_int32)]),
    "topopt_get_design": (C.c_int, [VP, VP]),
    "topopt_time_kernel": (C.c_int, [VP, VP, C.c_int32, C.c_int32, c_dp]),
}

_lib = None


class TopOptCUDAError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libtopopt_cuda error {code}: {msg}")
        self.code = code


def load():
    """dlopen the in-tree library and declare every prototype.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python topopt.jl_b200/build.py` "
            "(or __graft_entry__.build()); there is no CPU fallback"
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, handle=None):
    if rc != OK:
        msg = load().topopt_last_error(handle)
        msg = msg.decode() if msg else ""
        if rc == ERR_INVALID:
            raise ValueError(msg)  # the reference throws ArgumentError
        if rc == ERR_NONFINITE:
            raise FloatingPointError(msg)  # the reference throws DomainError
        raise TopOptCUDAError(rc, msg)


def ptr(a):
    """Pointer for the ABI: None -> NULL, numpy array -> host pointer, torch tensor -> data_ptr()."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.dtype == np.float64 and a.flags.c_contiguous, "need contiguous float64"
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        assert str(a.dtype) == "torch.float64" and a.is_contiguous(), "need contiguous float64 tensor"
        return a.data_ptr()
    raise TypeError(f"cannot pass {type(a)} through the C ABI")


def i64(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(c_ip)


def nels3(nels):
    arr = (C.c_int64 * 3)(*([int(n) for n in nels] + [1] * (3 - len(nels))))
    return arr
