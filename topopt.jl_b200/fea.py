"""Host-side mirror of src/FEA: ``FEASolver(Solver, problem; ...)`` returning a
``GenericFEASolver`` whose call runs the CG solve on the GPU through the C ABI.

Reference: src/FEA/solvers_api.jl:44-66 (solver types), :86-123 (GenericFEASolver fields),
:285-374 (call operator), :468-574,599-603 (FEASolver factory and defaults),
src/FEA/convergence_criteria.jl:13-45, src/Utilities/penalties.jl:30-54.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


# ---- penalties (src/Utilities/penalties.jl) ------------------------------------------------
class PowerPenaltyFun:
    kind = _lib.PENALTY_POWER

    def __init__(self, p):
        self.p = float(p)


class RationalPenaltyFun(PowerPenaltyFun):
    kind = _lib.PENALTY_RATIONAL


class SinhPenaltyFun(PowerPenaltyFun):
    kind = _lib.PENALTY_SINH


class HeavisideProjectionFun:
    """1 - exp(-beta x) + x exp(-beta)  (penalties.jl:83-86)"""

    code = 1

    def __init__(self, beta):
        self.beta = float(beta)


class SigmoidProjectionFun(HeavisideProjectionFun):
    """1 / (1 + exp((beta+1)(0.5 - x)))  (penalties.jl:93-96)"""

    code = 2


class ProjectedPenaltyFun:
    """penalty(proj(x)), default projection HeavisideProjectionFun(10)  (penalties.jl:62-70)"""

    def __init__(self, penalty, proj=None):
        self.penalty = penalty
        self.proj = proj if proj is not None else HeavisideProjectionFun(10.0)
        self.kind = penalty.kind

    @property
    def p(self):  # @forward_property ProjectedPenaltyFun penalty
        return self.penalty.p

    @p.setter
    def p(self, v):
        self.penalty.p = float(v)


# ---- convergence criteria (src/FEA/convergence_criteria.jl) -----------------------------------
class DefaultCriteria:
    code = _lib.CRITERIA_DEFAULT


class EnergyCriteria:
    code = _lib.CRITERIA_ENERGY


# ---- solver types selectable through FEASolver(...) (solvers_api.jl:44-66) ---------------------
class AbstractLinearSolver:
    pass


class CUDAMatrixFreeSolver(AbstractLinearSolver):
    """GPU counterpart of CGMatrixFreeSolver."""

    op = _lib.OP_MATRIX_FREE


class CUDAAssemblySolver(AbstractLinearSolver):
    """GPU counterpart of CGAssemblySolver (element-gathered CSR + SpMV)."""

    op = _lib.OP_ASSEMBLED


class PseudoDensities:
    """src/TopOpt.jl:30-76 value type; only ``.x`` matters at this boundary."""

    def __init__(self, x):
        self.x = np.ascontiguousarray(x, dtype=np.float64)


def _x(v):
    return v.x if isinstance(v, PseudoDensities) else np.ascontiguousarray(v, dtype=np.float64)


class GenericFEASolver:
    """Same public fields as the reference struct (solvers_api.jl:100-122) that the objective
    functions and filters read: problem, u, lhs, rhs, vars, penalty, xmin, cg_max_iter, abstol,
    preconditioner, conv.  The device handle replaces globalinfo/elementinfo."""

    def __init__(self, solver_type, problem, xmin, penalty, abstol, cg_max_iter, preconditioner, conv, reltol, device, comm, check_every,
                 cg_variant=0, warm_start=False, refresh_preconditioner=False):
        self.solver_type = solver_type
        self.problem = problem
        self.xmin = float(xmin)
        self.penalty = penalty
        self.prev_penalty = penalty
        self.abstol = float(abstol)
        self.reltol = float(reltol)
        self.cg_max_iter = int(cg_max_iter)
        self.preconditioner = preconditioner
        self.preconditioner_initialized = False
        self.conv = conv
        self.check_every = int(check_every)
        # extensions over the reference (all off by default): see include/topopt_cuda.h topopt_cg_opts
        self.cg_variant = int(cg_variant)
        self.warm_start = bool(warm_start)
        self.refresh_preconditioner = bool(refresh_preconditioner)
        self._solved_once = False
        self.u = np.zeros(problem.ndof)
        self.lhs = np.zeros(problem.ndof)
        self.rhs = np.zeros(problem.ndof)
        self.vars = np.ones(problem.nel)
        self.last_result = None
        self._lib = _lib.load()
        self._handle = C.c_void_p()
        rank, world, uid = (0, 1, None) if comm is None else (comm.rank, comm.world, comm.nccl_id)
        self.rank, self.world = rank, world
        d = _lib.Desc()
        d.dim, d.ncomp = problem.dim, problem.ncomp
        d.nels = _lib.nels3(problem.nels)
        d.sizes = (C.c_double * 3)(*([*problem.sizes] + [1.0] * (3 - problem.dim)))
        Ke = np.ascontiguousarray(problem.Ke.T, dtype=np.float64)  # column-major for the ABI
        pres, pres_p = _lib.i64(problem.prescribed_dofs)
        fl = np.ascontiguousarray(problem.fixedload, dtype=np.float64)
        d.Ke = Ke.ctypes.data_as(_lib.c_dp)
        d.prescribed_dofs = pres_p
        d.n_prescribed = pres.shape[0]
        d.fixedload = fl.ctypes.data_as(_lib.c_dp)
        d.cellvolumes = None
        d.cell_dofs = None
        d.fixed_diag = 0.0
        d.device = int(device)
        d.rank, d.world = rank, world
        self._uid = uid
        d.nccl_unique_id = C.cast(C.c_char_p(uid), C.c_void_p) if uid is not None else None
        _lib.check(self._lib.topopt_create(C.byref(d), C.byref(self._handle)))
        if comm is not None and world > 1 and getattr(comm, "peer_memory", True):
            # peer-memory fast path: exchange cudaIpc handles (collective over all ranks)
            buf = C.create_string_buffer(1024)
            n = C.c_int64()
            self._check(self._lib.topopt_ipc_export(self._handle, C.cast(buf, C.c_void_p), C.byref(n)))
            blobs = comm.all_gather_bytes(buf.raw[: n.value])
            ok = self._lib.topopt_ipc_import(self._handle, C.cast(C.c_char_p(blobs), C.c_void_p), n.value) == _lib.OK
            # the peer path is all-or-nothing: if one rank cannot map its neighbours (no P2P, ranks on different
            # nodes) every rank falls back to the NCCL halo exchange / all-reduce instead of hanging in a collective
            if not comm.all_agree(ok):
                self._check(self._lib.topopt_ipc_import(self._handle, None, 0))

    # -- lifetime ------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            self._lib.topopt_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._handle

    def _check(self, rc):
        _lib.check(rc, self._handle)

    # -- options -------------------------------------------------------------------------------
    def cg_opts(self, **over):
        o = _lib.CGOpts()
        o.abstol = over.get("abstol", self.abstol)
        o.reltol = over.get("reltol", self.reltol)
        o.maxiter = over.get("maxiter", self.cg_max_iter)
        o.op = self.solver_type.op
        if self.preconditioner in (None, "identity"):
            o.precond = _lib.PRECOND_NONE
        elif self.preconditioner == "multigrid":  # extension: geometric multigrid V-cycle, rebuilt when the stiffness changes
            o.precond = _lib.PRECOND_MULTIGRID
            o.mg_degree = int(over.get("mg_degree", getattr(self, "mg_degree", 0)))
            o.mg_ratio = float(over.get("mg_ratio", getattr(self, "mg_ratio", 0.0)))
        else:
            o.precond = _lib.PRECOND_JACOBI
        o.criteria = self.conv.code
        o.check_every = over.get("check_every", self.check_every)
        o.variant = over.get("variant", self.cg_variant)
        o.warm_start = 1 if over.get("warm_start", self.warm_start and self._solved_once) else 0
        o.refresh_precond = 1 if over.get("refresh_precond", self.refresh_preconditioner) else 0
        return o

    def set_density(self, x=None):
        """solver.vars .= x and the penalised stiffness on the device (compliance.jl:65)."""
        if x is not None:
            self.vars = _x(x)
        self.sync_projection()
        self._check(
            self._lib.topopt_set_density(self._handle, _lib.ptr(self.vars), self.penalty.kind, self.penalty.p, self.xmin, 1)
        )

    def sync_projection(self):
        """Tell the library which projection (if any) precedes the penalty."""
        proj = getattr(self.penalty, "proj", None)
        self._check(self._lib.topopt_set_projection(self._handle, proj.code if proj else 0, proj.beta if proj else 0.0))

    # -- the call operator (solvers_api.jl:285-374) -----------------------------------------------
    def __call__(self, assemble_f=True, rhs=None, lhs=None, download=True):
        """solver(): upload vars, CG solve, write solver.u (or ``lhs`` for a caller-supplied rhs).
        A 2-D ``rhs`` solves every column with a zero initial guess (solvers_api.jl:294-350)."""
        self.set_density()
        if not self.preconditioner_initialized and self.preconditioner not in (None, "identity", "multigrid"):
            # UpdatePreconditioner! runs once per solver lifetime (solvers_api.jl:187-192)
            self._check(self._lib.topopt_set_jacobi(self._handle, None))
            self.preconditioner_initialized = True
        default_rhs = assemble_f and rhs is None
        opts = self.cg_opts() if default_rhs else self.cg_opts(warm_start=False)  # a warm start only makes sense for the same load
        res = _lib.CGResult()
        if rhs is not None and np.ndim(rhs) == 2:
            rhs = np.asarray(rhs, dtype=np.float64)
            out = np.zeros_like(rhs) if lhs is None else lhs
            for j in range(rhs.shape[1]):
                col = np.ascontiguousarray(rhs[:, j])
                sol = np.zeros(self.problem.ndof)
                self._check(self._lib.topopt_solve(self._handle, _lib.ptr(col), _lib.ptr(sol), C.byref(opts), C.byref(res)))
                out[:, j] = sol
            self.last_result = res
            return out
        if assemble_f and rhs is None:
            target = self.u
            self._check(
                self._lib.topopt_solve(self._handle, None, _lib.ptr(target) if download else None, C.byref(opts), C.byref(res))
            )
            self._solved_once = True
        else:
            b = np.ascontiguousarray(self.rhs if rhs is None else rhs, dtype=np.float64)
            target = self.lhs if lhs is None else lhs
            self._check(self._lib.topopt_solve(self._handle, _lib.ptr(b), _lib.ptr(target), C.byref(opts), C.byref(res)))
        self.last_result = res
        return target

    # -- operator-level access used by the parity tests -------------------------------------------
    def mul(self, x):
        """mul!(y, MatrixFreeOperator, x) (matrix_free_operator.jl:66-105)."""
        y = np.empty(self.problem.ndof)
        self._check(self._lib.topopt_apply(self._handle, _lib.ptr(np.ascontiguousarray(x, dtype=np.float64)), _lib.ptr(y)))
        return y

    def mul_ex(self, x, kernel):
        """topopt_apply_ex: K x through one specific kernel generation; returns (y, x.y, y.y)."""
        y = np.empty(self.problem.ndof)
        dots = (C.c_double * 2)()
        self._check(self._lib.topopt_apply_ex(self._handle, _lib.ptr(np.ascontiguousarray(x, dtype=np.float64)), _lib.ptr(y), kernel, dots))
        return y, dots[0], dots[1]

    def assemble(self):
        """assemble! + apply! (assemble.jl:28-90): returns (nzval in CSC order, f)."""
        md = self.problem.metadata
        nz = np.empty(md.nnz)
        f = np.empty(md.ndof)
        self._check(self._lib.topopt_assemble(self._handle, _lib.ptr(nz), _lib.ptr(f)))
        return nz, f

    def spmv(self, x):
        y = np.empty(self.problem.ndof)
        self._check(self._lib.topopt_spmv(self._handle, _lib.ptr(np.ascontiguousarray(x, dtype=np.float64)), _lib.ptr(y)))
        return y

    def stats(self):
        s = _lib.Stats()
        self._check(self._lib.topopt_get_stats(self._handle, C.byref(s)))
        return s

    def reset_stats(self):
        self._check(self._lib.topopt_reset_stats(self._handle))

    def time_kernel(self, which, reps, filt=None):
        ms = C.c_double()
        self._check(self._lib.topopt_time_kernel(self._handle, filt.handle if filt is not None else None, which, reps, C.byref(ms)))
        return ms.value


def getcompliance(solver, u=None):
    """FEA.getcompliance (src/FEA/FEA.jl:40): u' K u, evaluated on the device with the current stiffness
    (``u`` None = the device-resident solution).  Equals dot(u, f) only for a fully converged default solve."""
    out = C.c_double()
    up = None if u is None else _lib.ptr(np.ascontiguousarray(u, dtype=np.float64))
    solver._check(solver._lib.topopt_quadratic_form(solver.handle, up, C.byref(out)))
    return out.value


def FEASolver(Solver, problem, *, xmin=1e-3, penalty=None, abstol=1e-7, cg_max_iter=700, preconditioner=None,
              conv=None, reltol=None, device=0, comm=None, check_every=0, cg_variant=0, warm_start=False,
              refresh_preconditioner=False):
    """FEASolver(Solver, problem; xmin, penalty, abstol, cg_max_iter, preconditioner, conv)
    with the reference's defaults (solvers_api.jl:468-489).  ``reltol`` is IterativeSolvers'
    sqrt(eps) unless overridden (the reference cannot override it)."""
    if not (isinstance(Solver, type) and issubclass(Solver, AbstractLinearSolver)):
        raise TypeError("Solver must be a subtype of AbstractLinearSolver")
    if not hasattr(Solver, "op"):
        raise TypeError(f"{Solver.__name__} has no GPU operator; use CUDAMatrixFreeSolver or CUDAAssemblySolver")
    return GenericFEASolver(
        Solver,
        problem,
        xmin,
        penalty if penalty is not None else PowerPenaltyFun(1.0),
        abstol,
        cg_max_iter,
        preconditioner,
        conv if conv is not None else DefaultCriteria(),
        reltol if reltol is not None else float(np.sqrt(np.finfo(np.float64).eps)),
        device,
        comm,
        check_every,
        cg_variant,
        warm_start,
        refresh_preconditioner,
    )
