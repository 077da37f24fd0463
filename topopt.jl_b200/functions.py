"""Host-side mirror of src/Functions for the compliance path: ComplianceFun
(compliance.jl:16-76), ThermalComplianceFun (thermal_compliance.jl:45-210), VolumeFun
(volume.jl:47-80).  The arithmetic runs in libtopopt_cuda; these classes only keep the fields
the reference's rrules and BESO read (``cell_comp``, ``grad``)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .fea import PseudoDensities, _x


class ComplianceFun:
    def __init__(self, solver):
        self.solver = solver
        self.problem = solver.problem
        self.cell_comp = np.zeros(solver.problem.nel)
        self.grad = np.zeros(solver.problem.nel)
        self.comp = 0.0

    def __call__(self, x):
        """comp(x): solver.vars .= x; solver(); compute_compliance (compliance.jl:58-70)."""
        s = self.solver
        s.vars = _x(x)
        s(download=True)
        obj = C.c_double()
        s._check(s._lib.topopt_compliance(s.handle, None, C.byref(obj), _lib.ptr(self.cell_comp), _lib.ptr(self.grad)))
        self.comp = obj.value
        return self.comp

    def value_and_grad(self, x):
        """What Zygote.pullback over the reference's rrule (compliance.jl:72-76) yields."""
        v = self(x)
        return v, self.grad.copy()


class ThermalComplianceFun:
    def __init__(self, solver):
        self.solver = solver
        self.problem = solver.problem
        self.cell_comp = np.zeros(solver.problem.nel)
        self.grad = np.zeros(solver.problem.nel)
        self.comp = 0.0
        self.adjoint_iters = 0

    def __call__(self, x):
        """J = Q'T, adjoint solve K lam = -Q_cond, grad_e = dE_e lam_e' Ke T_e
        (thermal_compliance.jl:116-158)."""
        s = self.solver
        s.vars = _x(x)
        s(download=True)
        obj = C.c_double()
        s._check(s._lib.topopt_dot(s.handle, None, None, C.byref(obj)))  # dot(fixedload, T)
        s._check(s._lib.topopt_swap_solution_lambda(s.handle))  # T -> lambda slot
        rhs = -np.asarray(self.problem.fixedload, dtype=np.float64)  # prescribed entries are zeroed by the library
        opts = s.cg_opts()
        res = _lib.CGResult()
        s._check(s._lib.topopt_solve(s.handle, _lib.ptr(rhs), None, C.byref(opts), C.byref(res)))
        self.adjoint_iters = res.iters
        # resident: u-slot = lambda, lambda-slot = T; the bilinear form is symmetric
        s._check(s._lib.topopt_bilinear_sens(s.handle, None, None, _lib.ptr(self.cell_comp), _lib.ptr(self.grad)))
        s._check(s._lib.topopt_swap_solution_lambda(s.handle))  # leave T as the resident solution
        self.comp = obj.value
        return self.comp

    def value_and_grad(self, x):
        v = self(x)
        return v, self.grad.copy()


class VolumeFun:
    """dot(x, cellvolumes)/total, gradient cellvolumes/total (volume.jl:59-80).  O(nel) host work;
    SURVEY 8a15 leaves it outside the GPU path."""

    def __init__(self, problem_or_solver):
        problem = getattr(problem_or_solver, "problem", problem_or_solver)
        self.problem = problem
        self.cellvolumes = problem.cellvolumes
        self.total_volume = float(problem.cellvolumes.sum())
        self.grad = problem.cellvolumes / self.total_volume

    def __call__(self, x):
        return float(np.dot(_x(x), self.cellvolumes)) / self.total_volume

    def value_and_grad(self, x):
        return self(x), self.grad.copy()
