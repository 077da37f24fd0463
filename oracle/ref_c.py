"""CPU ORACLE (C port loader) -- TEST / BASELINE INFRASTRUCTURE ONLY.

Compiles oracle/topopt_ref.c with gcc and exposes it through ctypes.  Two builds:
serial (faithful to the single-threaded reference) and OpenMP (all host cores)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "topopt_ref.c")
OUT = os.path.join(HERE, "_build")


def build(openmp=False, native=False, force=False):
    """gcc -O3 [-march=native | -mavx2 -mfma] [-fopenmp] -shared -> oracle/_build/libtopopt_ref*.so"""
    os.makedirs(OUT, exist_ok=True)
    name = "libtopopt_ref" + ("_omp" if openmp else "") + ("_native" if native else "") + ".so"
    path = os.path.join(OUT, name)
    if not force and os.path.exists(path) and os.path.getmtime(path) >= os.path.getmtime(SRC):
        return path
    arch = ["-march=native"] if native else ["-mavx2", "-mfma"]
    cmd = ["gcc", "-O3", "-std=c11", "-fPIC", "-shared", *arch] + (["-fopenmp"] if openmp else []) + ["-o", path, SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return path


def load(openmp=False, native=False):
    try:
        path = build(openmp=openmp, native=native)
    except Exception:
        if not native:
            raise
        path = build(openmp=openmp, native=False)
    lib = C.CDLL(path)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int64)
    lib.ref_create.restype = C.c_void_p
    lib.ref_create.argtypes = [C.c_int, C.c_int, ip, dp, ip, C.c_int64]
    lib.ref_destroy.argtypes = [C.c_void_p]
    lib.ref_threads.restype = C.c_int
    lib.ref_ndof.restype = C.c_int64
    lib.ref_ndof.argtypes = [C.c_void_p]
    lib.ref_cell_dofs.argtypes = [C.c_void_p, ip]
    lib.ref_set_density.argtypes = [C.c_void_p, dp, C.c_double, C.c_double]
    lib.ref_mul.argtypes = [C.c_void_p, dp, dp]
    lib.ref_cg.restype = C.c_int
    lib.ref_cg.argtypes = [C.c_void_p, dp, dp, C.c_double, C.c_double, C.c_int, dp]
    lib.ref_compliance.restype = C.c_double
    lib.ref_compliance.argtypes = [C.c_void_p, dp, dp, C.c_double, C.c_double, dp, dp]
    lib.ref_filter.argtypes = [C.c_void_p, dp, C.c_double, dp, dp]
    return lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class RefProblem:
    """Thin object wrapper over the C port."""

    def __init__(self, dim, ncomp, nels, Ke, prescribed_1based, sizes=None, openmp=False, native=False):
        self.lib = load(openmp=openmp, native=native)
        self.dim, self.ncomp = dim, ncomp
        self.nels = tuple(int(n) for n in nels)
        self.sizes = np.array(list(sizes if sizes is not None else (1.0,) * dim) + [1.0] * (3 - dim), dtype=np.float64)
        nel3 = np.array(list(self.nels) + [1] * (3 - dim), dtype=np.int64)
        Kc = np.ascontiguousarray(np.asarray(Ke, dtype=np.float64).T)
        pres = np.ascontiguousarray(prescribed_1based, dtype=np.int64)
        self.h = self.lib.ref_create(dim, ncomp, nel3.ctypes.data_as(C.POINTER(C.c_int64)), _dp(Kc), pres.ctypes.data_as(C.POINTER(C.c_int64)), pres.shape[0])
        self.ndof = int(self.lib.ref_ndof(self.h))
        self.nel = int(np.prod(self.nels))
        self.threads = int(self.lib.ref_threads())

    def close(self):
        if self.h:
            self.lib.ref_destroy(self.h)
            self.h = None

    def cell_dofs(self):
        ks = self.ncomp * 2**self.dim
        out = np.empty((self.nel, ks), dtype=np.int64)
        self.lib.ref_cell_dofs(self.h, out.ctypes.data_as(C.POINTER(C.c_int64)))
        return out.T

    def set_density(self, rho, p, xmin):
        rho = np.ascontiguousarray(rho, dtype=np.float64)
        self.lib.ref_set_density(self.h, _dp(rho), p, xmin)

    def mul(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        self.lib.ref_mul(self.h, _dp(x), _dp(y))
        return y

    def cg(self, b, abstol=1e-7, reltol=None, maxiter=700):
        reltol = float(np.sqrt(np.finfo(float).eps)) if reltol is None else reltol
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty_like(b)
        res = C.c_double()
        it = self.lib.ref_cg(self.h, _dp(b), _dp(x), abstol, reltol, maxiter, C.byref(res))
        return x, int(it), res.value

    def compliance(self, u, rho, p, xmin):
        u = np.ascontiguousarray(u, dtype=np.float64)
        rho = np.ascontiguousarray(rho, dtype=np.float64)
        cell, grad = np.empty(self.nel), np.empty(self.nel)
        obj = self.lib.ref_compliance(self.h, _dp(u), _dp(rho), p, xmin, _dp(cell), _dp(grad))
        return float(obj), cell, grad

    def filter(self, rmin, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        self.lib.ref_filter(self.h, _dp(self.sizes), rmin, _dp(x), _dp(y))
        return y


def point_load_cantilever_inputs(nels, sizes=None, E=1.0, nu=0.3, force=1.0):
    """Inputs of the C port for PointLoadCantilever built from the ORACLE only (no product code):
    (Ke, prescribed dofs 1-based sorted, fixedload).  Same construction as
    topopt_oracle.PointLoadCantilever (problem_types.jl:162-223) without the ragged Metadata tables,
    so it stays cheap at 12.8 M dofs."""
    import topopt_oracle as o

    g = o.Grid(nels, sizes)
    block = o.ferrite_node_blocks(g.cells, g.nnodes)
    ncomp = g.dim
    fixed_nodes = np.nonzero(g.left())[0]
    prescribed = np.unique((ncomp * block[fixed_nodes][:, None] + np.arange(ncomp)[None, :]).ravel()) + 1
    fnode = np.nonzero(g.right() & g.middley())[0][0]
    f = np.zeros(ncomp * g.nnodes)
    f[ncomp * block[fnode] + 1] = -force
    return o.element_stiffness(g.dim, g.sizes, E, nu), prescribed.astype(np.int64), f
