"""SIMP driver used by the parity tests and the benchmark: one evaluation is
x -> filter -> penalise -> CG solve -> compliance + sensitivity -> filter pullback
(docs/tutorials/simp.qmd:86-114).  The design update is a host-side optimality-criteria step
(the reference drives MMA from the third-party NonconvexMMA package)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def simp_eval(solver, filt, x, grad_out=None):
    """Fused device-resident evaluation through topopt_simp_eval.  ``x``/``grad_out`` may be numpy
    arrays (host), torch CUDA tensors (device) or None (resident design / gradient stays on the
    device).  Returns (objective, CGResult)."""
    opts = solver.cg_opts()
    res = _lib.CGResult()
    obj = C.c_double()
    kind = 0 if filt is None else filt.kind
    solver.sync_projection()
    solver._check(
        solver._lib.topopt_simp_eval(
            solver.handle, None if filt is None else filt.handle, kind, _lib.ptr(x), solver.penalty.kind, solver.penalty.p,
            solver.xmin, C.byref(opts), C.byref(obj), _lib.ptr(grad_out), C.byref(res),
        )
    )
    solver._solved_once = True
    return obj.value, res


def oc_update(x, dc, dv, volfrac, move=0.2, eta=0.5, xlo=0.0):
    """Optimality-criteria update with bisection on the volume multiplier."""
    l1, l2 = 0.0, 1e9
    dc = np.minimum(dc, 0.0)
    xn = x
    while (l2 - l1) / (l1 + l2) > 1e-12 and l2 > 1e-40:
        lmid = 0.5 * (l1 + l2)
        xn = np.maximum(xlo, np.maximum(x - move, np.minimum(1.0, np.minimum(x + move, x * (-dc / dv / lmid) ** eta))))
        if float(xn @ dv) > volfrac:
            l1 = lmid
        else:
            l2 = lmid
    return xn


def simp_loop(solver, filt, volfrac, iters=10, x0=None):
    """iters x (simp_eval + OC).  Returns (x, objective history)."""
    prob = solver.problem
    x = np.full(prob.nel, float(volfrac)) if x0 is None else np.array(x0, dtype=np.float64)
    dv = prob.cellvolumes / prob.cellvolumes.sum()
    dvf = filt.pullback(dv) if (filt is not None and filt.kind == 1) else dv
    g = np.empty(prob.nel)
    hist = []
    for _ in range(iters):
        obj, _res = simp_eval(solver, filt, x, g)
        hist.append(obj)
        x = oc_update(x, g, dvf, volfrac)
    return x, hist


def oc_update_device(solver, volfrac, dv=None, x=None, dc=None, x_out=None, move=0.2, eta=0.5, xlo=0.0):
    """topopt_oc_update: the same update as ``oc_update`` on the device-resident design and gradient.
    Returns (||xn - x||_2, bisection steps)."""
    change = C.c_double()
    nb = C.c_int32()
    solver._check(
        solver._lib.topopt_oc_update(solver.handle, _lib.ptr(x), _lib.ptr(dc), _lib.ptr(dv), float(volfrac), float(move), float(eta), float(xlo),
                                     _lib.ptr(x_out), C.byref(change), C.byref(nb))
    )
    return change.value, nb.value


def simp_loop_device(solver, filt, volfrac, iters=10, x0=None):
    """Device-resident SIMP loop: the design, the state and the gradient never leave the GPU between iterations
    (per iteration 8 bytes of objective and 8 bytes per bisection step come back).  Returns (x, objective history)."""
    prob = solver.problem
    x = np.full(prob.nel, float(volfrac)) if x0 is None else np.array(x0, dtype=np.float64)
    dv = prob.cellvolumes / prob.cellvolumes.sum()
    dvf = filt.pullback(dv) if (filt is not None and filt.kind == 1) else dv
    hist = []
    for k in range(iters):
        obj, _res = simp_eval(solver, filt, x if k == 0 else None, None)
        hist.append(obj)
        oc_update_device(solver, volfrac, dv=np.ascontiguousarray(dvf) if k == 0 else None)
    out = np.empty(prob.nel)
    solver._check(solver._lib.topopt_get_design(solver.handle, _lib.ptr(out)))
    return out, hist
