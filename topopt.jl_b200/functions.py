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


class _FieldFun:
    """Shared body of DisplacementFun / TemperatureFun: value = the solved field, pullback = one adjoint solve
    with the cotangent as right-hand side and the bilinear element form."""

    _physics = None

    def __init__(self, solver, maxfevals=10**8):
        if self._physics is not None and getattr(solver.problem, "physics", self._physics) != self._physics:
            kind = "StiffnessTopOptProblem" if self._physics == _lib.PHYSICS_ELASTICITY else "HeatTransferTopOptProblem"
            raise ValueError(f"{type(self).__name__} can only be used with {kind}")  # ArgumentError in the reference
        self.solver = solver
        self.problem = solver.problem
        self.u = np.zeros(solver.problem.ndof)
        self.dudx_tmp = np.zeros(solver.problem.nel)
        self.fevals = 0
        self.maxfevals = maxfevals

    def __call__(self, x):
        """solver.vars .= x; solver(); copy(solver.u)  (displacement.jl:70-84, temperature.jl:71-83)"""
        s = self.solver
        self.fevals += 1
        s.vars = _x(x)
        s(download=True)
        self.u = s.u.copy()
        return self.u.copy()

    def pullback(self, delta):
        """(du/dx)' delta: K lam = delta (apply_zero! on the rhs, solvers_api.jl:364-367), then
        out_e = -dE_e u_e' Ke lam_e  (displacement.jl:97-121, temperature.jl:92-119).  Both fields vanish on
        the prescribed dofs (homogeneous Dirichlet), so bcmatrix(Ke) and rawmatrix(Ke) give the same form."""
        s = self.solver
        lam = s(assemble_f=False, rhs=np.ascontiguousarray(delta, dtype=np.float64), lhs=np.zeros(self.problem.ndof))
        # device: resident solution = lam -> move it to the lambda slot, upload the forward field as u
        s._check(s._lib.topopt_swap_solution_lambda(s.handle))
        s._check(s._lib.topopt_bilinear_sens(s.handle, None, _lib.ptr(self.u), None, _lib.ptr(self.dudx_tmp)))
        self.lam = lam
        return -self.dudx_tmp

    def value_and_pullback(self, x):
        return self(x), self.pullback


class DisplacementFun(_FieldFun):
    """src/Functions/displacement.jl:28-121"""

    _physics = _lib.PHYSICS_ELASTICITY


class TemperatureFun(_FieldFun):
    """src/Functions/temperature.jl:30-119"""

    _physics = _lib.PHYSICS_HEAT
