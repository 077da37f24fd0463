"""Safety net for what ships: secondary kernel instantiations that the BASELINE
configurations do not reach, the DomainError path, run-to-run determinism under stress on multi-tile grids,
and compute-sanitizer memcheck over one small solve through the ring-staged kernel."""
import os
import subprocess
import sys

import numpy as np
import pytest

import topopt_oracle as o

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300))


def heat3d(t, nels):
    """3-D HeatTree (scalar hex8: k_apply<3,1>, k_sens<3,1>, k_diag<3,1>): the host mirror only builds the 2-D
    heat problems, so the 3-D one is assembled here from the oracle's data on the generic structured problem."""
    oprob = o.HeatTree(nels)
    prob = t.problems.StructuredProblem(nels, None, 1)
    prob.physics = t._lib.PHYSICS_HEAT
    prob.Ke = oprob.Ke.copy()
    prob.prescribed_dofs = np.sort(oprob.prescribed) + 1
    prob.fixedload = oprob.fixedload.copy()
    return prob, oprob


def test_scalar_hex8_heat_kernels(lib):
    t = lib
    prob, oprob = heat3d(t, (9, 6, 8))
    s = t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), abstol=1e-12, reltol=1e-14, cg_max_iter=20000)
    rho = np.random.default_rng(3).uniform(0.2, 1.0, prob.nel)
    E = o.get_rho(rho, 3.0, 1e-3)
    s.set_density(rho)
    for zero_fixed in (True, False):
        x = np.random.default_rng(4).standard_normal(prob.ndof)
        if zero_fixed:
            x[oprob.prescribed] = 0.0
        assert rel(s.mul(x), o.matfree_mul(oprob, E, x)) < 1e-12
    tc = t.ThermalComplianceFun(s)
    J, g = tc.value_and_grad(rho)
    Jo, _, go, _, _ = o.thermal_compliance(oprob, rho, 3.0, 1e-3, abstol=1e-13, reltol=1e-15, maxiter=20000)
    assert abs(J - Jo) / abs(Jo) < 1e-8 and rel(g, go) < 1e-8
    # Jacobi-preconditioned solve: k_diag<3,1>
    sj = t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), abstol=1e-12, reltol=1e-14, cg_max_iter=20000, preconditioner="jacobi")
    sj.vars = rho
    u = sj().copy()
    assert rel(u, o.solve_direct(oprob, E)) < 1e-8
    s.close()
    sj.close()


@pytest.mark.parametrize("nels,sizes,rmin,env", [
    ((10, 6, 8), None, 1.2, {}),                                 # R = 1: k_filter_stencil3<1,*>
    ((12, 8, 10), None, 2.7, {}),                                # R = 3: k_filter_stencil3<3,*>
    ((12, 8, 10), None, 3.6, {}),                                # R = 4: generic shared-memory tiled stencil
    ((9, 6, 8), (1.0, 0.5, 2.0), 1.7, {}),                       # anisotropic window: generic tiled stencil
    ((10, 6, 8), None, 2.0, {"TOPOPT_FILTER_GENERIC": "1"}),     # cubic window forced through the generic tiled stencil
    ((10, 6, 8), None, 2.0, {"TOPOPT_FILTER_UNTILED": "1", "TOPOPT_FILTER_GENERIC": "1"}),  # untiled gather kernels
    ((14, 10), None, 3.3, {"TOPOPT_FILTER_UNTILED": "1"}),       # 2-D untiled
])
def test_filter_kernel_variants(lib, nels, sizes, rmin, env):
    t = lib
    args = (nels,) if sizes is None else (nels, sizes)
    prob, oprob = t.PointLoadCantilever(*args), o.PointLoadCantilever(*args)
    os.environ.update(env)
    try:
        s = t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0))
        F, S = t.DensityFilterFun(s, rmin), t.SensFilterFun(s, rmin)
        Fo, So = o.DensityFilter(oprob, rmin), o.SensFilter(oprob, rmin)
        x = np.random.default_rng(7).uniform(0.2, 1.0, prob.nel)
        d = np.random.default_rng(8).standard_normal(prob.nel)
        assert rel(F(x), Fo(x)) < 1e-13
        assert rel(F.pullback(d), Fo.pullback(d)) < 1e-12
        assert rel(S.pullback(d), So.pullback(d)) < 1e-12
        assert np.max(np.abs(F(np.full(prob.nel, 0.37)) - 0.37)) < 1e-14
        F.close(); S.close(); s.close()
    finally:
        for k in env:
            os.environ.pop(k, None)


def test_nonfinite_is_a_domain_error(lib):
    """TOPOPT_ERR_NONFINITE <-> DomainError (src/FEA/convergence_criteria.jl:34-41): a NaN element matrix is
    refused at construction, a NaN stiffness surfaces from the CG loop (default and energy criteria)."""
    import ctypes as C

    t = lib
    prob = t.PointLoadCantilever((8, 4, 4))
    bad = t.PointLoadCantilever((8, 4, 4))
    bad.Ke = bad.Ke.copy()
    bad.Ke[3, 5] = np.nan
    with pytest.raises(FloatingPointError):
        t.FEASolver(t.CUDAMatrixFreeSolver, bad, penalty=t.PowerPenaltyFun(3.0))
    for conv in (t.DefaultCriteria(), t.EnergyCriteria()):
        s = t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), conv=conv, cg_max_iter=50)
        E = np.full(prob.nel, 0.5)
        E[7] = np.nan
        s._check(s._lib.topopt_set_stiffness(s.handle, E.ctypes.data, None))
        opts, res = s.cg_opts(), t._lib.CGResult()
        u = np.zeros(prob.ndof)
        with pytest.raises(FloatingPointError):
            s._check(s._lib.topopt_solve(s.handle, None, u.ctypes.data, C.byref(opts), C.byref(res)))
        s.close()


@pytest.mark.parametrize("kernel", [2, 3, 4])
def test_hundredfold_bitwise_repeat_on_multi_tile_grid(lib, kernel):
    """No atomics on the data path and fixed summation orders: 100 launches of the neighbour-synchronised one-row,
    two-row and ring-staged kernels on a >= 6-tile, >= 96-plane grid give bit-identical vectors and dot products."""
    t = lib
    nels = (64, 46, 98)
    prob = t.PointLoadCantilever(nels)
    s = t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0))
    s.set_density(np.random.default_rng(1).uniform(0.2, 1.0, prob.nel))
    x = np.random.default_rng(2).standard_normal(prob.ndof)
    x[prob.prescribed_dofs - 1] = 0.0
    y0, a0, b0 = s.mul_ex(x, kernel)
    for _ in range(100):
        y, a, b = s.mul_ex(x, kernel)
        assert a == a0 and b == b0 and np.array_equal(y, y0)
    s.close()


@pytest.mark.skipif(os.environ.get("TOPOPT_SKIP_SANITIZER") == "1", reason="sanitizer disabled")
def test_memcheck_small_solve():
    """compute-sanitizer --tool memcheck over one small SIMP evaluation through the ring-staged kernel (bulk copies,
    mbarriers, per-warp rings) and the single-pass CG: no invalid accesses, no leaks of device errors."""
    import shutil

    cs = shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"
    if not os.path.exists(cs):
        pytest.skip("compute-sanitizer not installed")
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import numpy as np, topopt_jl_b200 as t\n"
        "p = t.PointLoadCantilever((33, 24, 14))\n"
        "s = t.FEASolver(t.CUDAMatrixFreeSolver, p, penalty=t.PowerPenaltyFun(3.0), cg_max_iter=12, cg_variant=1)\n"
        "F = t.DensityFilterFun(s, 2.0)\n"
        "g = np.empty(p.nel)\n"
        "obj, res = t.simp_eval(s, F, np.full(p.nel, 0.4), g)\n"
        "y, a, b = s.mul_ex(np.zeros(p.ndof), 4)\n"
        "print('OK', obj, res.iters)\n" % ROOT
    )
    out = subprocess.run([cs, "--tool", "memcheck", "--error-exitcode", "9", "--print-limit", "5", sys.executable, "-c", code],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "OK" in out.stdout and "ERROR SUMMARY: 0 errors" in out.stdout, out.stdout[-2000:]
