"""Host-side mirror of src/CheqFilters: DensityFilterFun (density_filter.jl:14-47) and
SensFilterFun (sens_filter.jl:13-70), applied on the GPU as a two-stage stencil instead of the
reference's explicit sparse Jacobian."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .fea import PseudoDensities, _x


class _Filter:
    def __init__(self, solver, rmin):
        self.solver = solver
        self.rmin = float(rmin)
        self._f = C.c_void_p()
        _lib.check(solver._lib.topopt_filter_create(solver.handle, self.rmin, C.byref(self._f)), solver.handle)

    @property
    def handle(self):
        return self._f

    def _apply(self, x, mode):
        x = _x(x)
        y = np.empty_like(x)
        _lib.check(self.solver._lib.topopt_filter_apply(self._f, _lib.ptr(x), _lib.ptr(y), mode), self.solver.handle)
        return y

    def close(self):
        if self._f is not None and self._f.value and self.solver.handle.value:
            self.solver._lib.topopt_filter_destroy(self._f)
        self._f = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DensityFilterFun(_Filter):
    kind = 1

    def __call__(self, x):
        """out = jacobian * x (density_filter.jl:35-40)"""
        return self._apply(x, _lib.FILTER_FORWARD)

    def pullback(self, delta):
        """jacobian' * delta (density_filter.jl:41-47)"""
        return self._apply(delta, _lib.FILTER_TRANSPOSE)


class SensFilterFun(_Filter):
    kind = 2

    def __call__(self, x):
        """identity forward (sens_filter.jl:49-51)"""
        return _x(x).copy()

    def pullback(self, delta):
        """nodal smoothing of the cotangent: the forward two-stage map (sens_filter.jl:52-110)"""
        return self._apply(delta, _lib.FILTER_FORWARD)
