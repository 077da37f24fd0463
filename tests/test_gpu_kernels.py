"""GPU parity tests of the kernel generations that ship for the large 3-D configurations, driven one by
one through the C ABI (topopt_apply_ex) against the CPU oracle, plus the CG variants built on them.

* every hex8 K.u generation (dense gather, modal one-row, modal two-row, ring-staged) on grids that are
  tall enough for the two-row / ring kernels to be the ones the library selects (>= 96 node planes),
  span several tiles in x and y, have ragged last tiles and odd / even node counts (bulk-copy alignment);
* the dot products each kernel fuses for CG;
* CG iterate parity with the reference recurrence THROUGH the selected kernels, and the single-pass
  recurrence against the reference recurrence;
* BASELINE config 4 (256x128x128, 12.8 M dofs) at full size against the OpenMP C port of the
  reference CPU path: K.u <= 1e-12, ten CG iterates <= 1e-10.

Reference: src/FEA/matrix_free_operator.jl:66-105, IterativeSolvers 0.9 cg!.
"""
import ctypes as C
import os

import numpy as np
import pytest

import topopt_oracle as o

pytestmark = pytest.mark.gpu

RTOL_OP = 1e-12
KERNELS = {"dense": 1, "one_row": 2, "two_row": 3, "ring": 4}


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300))


def rand_rho(n, seed=42):
    return np.random.default_rng(seed).uniform(0.2, 1.0, n)


def masked_x(oprob, seed):
    x = np.random.default_rng(seed).standard_normal(oprob.ndof)
    x[oprob.prescribed] = 0.0
    return x


def make(t, nels, sizes=None, **kw):
    args = (nels,) if sizes is None else (nels, sizes)
    prob, oprob = t.PointLoadCantilever(*args), o.PointLoadCantilever(*args)
    prob.Ke = oprob.Ke.copy()
    kw.setdefault("penalty", t.PowerPenaltyFun(3.0))
    return prob, oprob, t.FEASolver(t.CUDAMatrixFreeSolver, prob, **kw)


# nown >= 96 selects the two-row kernel in topopt_apply and nown >= 24 the ring kernel inside CG
TALL_GRIDS = [
    (8, 4, 100),      # one tile, thin and tall
    (33, 24, 100),    # 2 x 2 tiles, ragged, NX even / NY odd
    (64, 46, 98),     # 3 x 3 tiles (ring: 31-column, 23-row tiles), NX odd
    (30, 22, 26),     # exactly one ring tile of owned nodes + 1
    (62, 22, 26),     # tile boundaries fall on the last node column / row
    (61, 46, 24),     # NX even (row alignment alternates differently), last ring tile one column wide
]


@pytest.mark.parametrize("nels", TALL_GRIDS)
@pytest.mark.parametrize("kernel", list(KERNELS))
def test_hex8_kernel_generations_match_reference_operator(lib, nels, kernel):
    t = lib
    prob, oprob, s = make(t, nels)
    rho = rand_rho(prob.nel, 5)
    E = o.get_rho(rho, 3.0, 1e-3)
    s.set_density(rho)
    x = masked_x(oprob, 3)
    yref = o.matfree_mul(oprob, E, x)
    y, xy, yy = s.mul_ex(x, KERNELS[kernel])
    assert rel(y, yref) < RTOL_OP
    assert abs(xy - float(x @ yref)) <= 1e-11 * float(np.abs(x) @ np.abs(yref))
    if kernel == "ring":
        assert abs(yy - float(yref @ yref)) <= 1e-11 * float(yref @ yref)
    # bitwise run-to-run reproducibility (fixed summation order, no atomics on the data path)
    y2, xy2, yy2 = s.mul_ex(x, KERNELS[kernel])
    assert np.array_equal(y, y2) and xy == xy2 and yy == yy2
    s.close()


def test_ring_kernel_noncubic_cells_and_default_selection(lib):
    """Non-cubic cells change every modal coefficient; topopt_apply (kernel 0) on >= 96 planes is the
    two-row kernel and must agree with the ring kernel on a premasked vector to rounding."""
    t = lib
    prob, oprob, s = make(t, (20, 10, 98), (1.0, 0.5, 2.0))
    rho = rand_rho(prob.nel, 8)
    s.set_density(rho)
    E = o.get_rho(rho, 3.0, 1e-3)
    x = masked_x(oprob, 13)
    yref = o.matfree_mul(oprob, E, x)
    for k in (0, 3, 4):
        y, _, _ = s.mul_ex(x, k)
        assert rel(y, yref) < RTOL_OP
    # arbitrary (unmasked) x through the default path: bcmatrix column masking + meandiag rows
    xr = np.random.default_rng(2).standard_normal(oprob.ndof)
    assert rel(s.mul(xr), o.matfree_mul(oprob, E, xr)) < RTOL_OP
    s.close()


@pytest.mark.parametrize("nels", [(12, 8, 30), (33, 24, 100)])
def test_cg_iterates_through_shipped_kernels(lib, nels):
    """IterativeSolvers' recurrence through the kernels the library selects on tall grids (ring-staged
    K.u with the fused p.Ap): same iterates as the oracle's cg! for 1 / 5 / 20 iterations."""
    t = lib
    prob, oprob, s0 = make(t, nels)
    s0.close()
    rho = rand_rho(prob.nel)
    E = o.get_rho(rho, 3.0, 1e-3)
    for maxiter in (1, 5, 20):
        s = t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), cg_max_iter=maxiter, abstol=0.0, reltol=0.0)
        s.vars = rho
        u = s().copy()
        uref, it, res = o.solve_matfree(oprob, E, abstol=0.0, reltol=0.0, maxiter=maxiter)
        assert s.last_result.iters == it == maxiter
        assert abs(s.last_result.residual - res) <= 1e-9 * res
        assert rel(u, uref) < (1e-12 if maxiter <= 5 else 1e-7)
        s.close()


@pytest.mark.parametrize("nels", [(12, 8, 30), (40, 26, 50)])
def test_single_pass_cg_matches_reference_recurrence(lib, nels):
    """TOPOPT_CG_SINGLE_PASS: beta predicted from alpha^2 Ap.Ap - r.r, one fused vector pass.  Same
    Krylov method: iterates agree with the reference recurrence to rounding for short runs, the converged
    solution / compliance / sensitivities to the solve tolerance, iteration counts within a few."""
    t = lib
    prob, oprob, s0 = make(t, nels)
    s0.close()
    rho = rand_rho(prob.nel, 21)
    mk = lambda variant, **kw: t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), cg_variant=variant, **kw)
    for maxiter in (1, 5, 20):
        a, b = mk(0, cg_max_iter=maxiter, abstol=0.0, reltol=0.0), mk(1, cg_max_iter=maxiter, abstol=0.0, reltol=0.0)
        a.vars = rho
        b.vars = rho
        ua, ub = a().copy(), b().copy()
        assert a.last_result.iters == b.last_result.iters == maxiter
        assert rel(ub, ua) < (1e-12 if maxiter <= 5 else 1e-8)
        assert abs(a.last_result.residual - b.last_result.residual) <= 1e-8 * a.last_result.residual
        a.close()
        b.close()
    a, b = mk(0, abstol=1e-10, reltol=0.0, cg_max_iter=20000), mk(1, abstol=1e-10, reltol=0.0, cg_max_iter=20000)
    ca, cb = t.ComplianceFun(a), t.ComplianceFun(b)
    va, ga = ca.value_and_grad(rho)
    vb, gb = cb.value_and_grad(rho)
    assert a.last_result.converged == 1 and b.last_result.converged == 1
    assert abs(a.last_result.iters - b.last_result.iters) <= max(3, a.last_result.iters // 20)
    assert abs(va - vb) / va < 1e-8 and rel(gb, ga) < 1e-8 and rel(b.u, a.u) < 1e-7
    if oprob.nel < 3000:
        uref = o.solve_direct(oprob, o.get_rho(rho, 3.0, 1e-3))
        obj, _, g = o.compliance(oprob, uref, rho, 3.0, 1e-3)
        assert abs(vb - obj) / obj < 1e-8 and rel(gb, g) < 1e-8
    a.close()
    b.close()


@pytest.mark.parametrize("nels", [(12, 8, 30), (40, 26, 50), (70, 44, 24)])
def test_one_kernel_cg_iteration_matches_two_kernel_single_pass(lib, nels, monkeypatch):
    """kxu_hex8_cgfused.cuh: x, r, p of iteration k-1 are formed while the planes of K.p_k are staged (one launch
    per iteration, ping-pong p / r / Ap).  Same recurrence as k_apply_hex8_ring + k_update_xrp: iterates agree to
    rounding (only the summation order of the dot products differs), iteration counts and residuals match, the
    converged solution meets the oracle, and a second solve on the same handle (graph replay, parity restart) agrees."""
    t = lib
    prob, oprob, s0 = make(t, nels)
    s0.close()
    rho = rand_rho(prob.nel, 33)

    def mk(fused, **kw):
        monkeypatch.setenv("TOPOPT_CG_FUSED", "1" if fused else "0")
        return t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), cg_variant=1, **kw)

    for maxiter in (1, 2, 5, 20, 61):
        a, b = mk(False, cg_max_iter=maxiter, abstol=0.0, reltol=0.0), mk(True, cg_max_iter=maxiter, abstol=0.0, reltol=0.0)
        a.vars = rho
        b.vars = rho
        ua, ub = a().copy(), b().copy()
        la, lb = a.stats().kernel_launches, b.stats().kernel_launches
        assert a.last_result.iters == b.last_result.iters == maxiter
        assert rel(ub, ua) < (1e-12 if maxiter <= 5 else 1e-8), maxiter
        assert abs(a.last_result.residual - b.last_result.residual) <= 1e-8 * a.last_result.residual
        assert lb < la or maxiter == 1  # one launch per iteration instead of two
        ub2 = b().copy()  # same handle again: the iteration parity restarts from the first buffer set
        assert np.array_equal(ub2, ub)
        a.close()
        b.close()
    a, b = mk(False, abstol=1e-10, reltol=0.0, cg_max_iter=20000), mk(True, abstol=1e-10, reltol=0.0, cg_max_iter=20000)
    ca, cb = t.ComplianceFun(a), t.ComplianceFun(b)
    va, ga = ca.value_and_grad(rho)
    vb, gb = cb.value_and_grad(rho)
    assert a.last_result.converged == 1 and b.last_result.converged == 1
    assert abs(a.last_result.iters - b.last_result.iters) <= max(3, a.last_result.iters // 20)
    assert abs(va - vb) / va < 1e-8 and rel(gb, ga) < 1e-8 and rel(b.u, a.u) < 1e-7
    if oprob.nel < 3000:
        uref = o.solve_direct(oprob, o.get_rho(rho, 3.0, 1e-3))
        obj, _, g = o.compliance(oprob, uref, rho, 3.0, 1e-3)
        assert abs(vb - obj) / obj < 1e-8 and rel(gb, g) < 1e-8
    a.close()
    b.close()


@pytest.mark.parametrize("tma", ["0", "1"])
@pytest.mark.parametrize("nels,grid", [((70, 44, 24), "0"), ((70, 44, 24), "7"), ((70, 44, 24), "37"), ((40, 26, 50), "29"),
                                       ((70, 44, 24), "3"), ((70, 44, 24), "11"), ((40, 26, 50), "5"), ((256, 128, 128), "0")])
def test_one_kernel_cg_iteration_dense_rhs(lib, nels, grid, tma, monkeypatch):
    """A DENSE right-hand side makes every owned dof, every tile halo and every segment boundary of the one-kernel
    iteration live from the first iteration on (a point load only reaches them after many iterations).  Odd CTA counts
    (TOPOPT_CG_FUSED_GRID) put segment boundaries inside tile columns and across them; both staging variants (row-wise
    bulk copies, tensor-map boxes).  Against the two-kernel path:
    iterates <= 1e-12, residual norms <= 1e-11; the 3-iterate also against the oracle / the C port."""
    t = lib
    if nels[0] == 256 and os.environ.get("TOPOPT_SKIP_FULL_SIZE") == "1":
        pytest.skip("full-size config 4 disabled")
    prob = t.PointLoadCantilever(nels)
    rho = rand_rho(prob.nel, 44)
    b = np.random.default_rng(45).standard_normal(prob.ndof)
    b[prob.prescribed_dofs - 1] = 0.0

    def mk(fused, maxiter):
        monkeypatch.setenv("TOPOPT_CG_FUSED", "1" if fused else "0")
        monkeypatch.setenv("TOPOPT_CG_FUSED_TMA", tma)  # 1: planes staged by tensor-map copies (kxu_hex8_cgtma.cuh)
        if grid != "0":
            monkeypatch.setenv("TOPOPT_CG_FUSED_GRID", grid)
        return t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), cg_variant=1, cg_max_iter=maxiter, abstol=0.0, reltol=0.0)

    for maxiter in (1, 3, 12, 40):
        a, f = mk(False, maxiter), mk(True, maxiter)
        a.set_density(rho)
        f.set_density(rho)
        ua, uf = a(rhs=b).copy(), f(rhs=b).copy()
        assert a.last_result.iters == f.last_result.iters == maxiter
        assert rel(uf, ua) < 1e-12, (maxiter, rel(uf, ua))
        assert abs(a.last_result.residual - f.last_result.residual) <= 1e-11 * a.last_result.residual, maxiter
        if maxiter == 3 and prob.nel < 100000:
            oprob = o.PointLoadCantilever(nels)
            uref, it, res = o.solve_matfree(oprob, o.get_rho(rho, 3.0, 1e-3), abstol=0.0, reltol=0.0, maxiter=3, rhs=b)
            assert rel(uf, uref) < 1e-12 and abs(f.last_result.residual - res) <= 1e-11 * res
        a.close()
        f.close()


@pytest.mark.parametrize("nels", [(32, 16, 16), (48, 16, 32)])
def test_multigrid_preconditioned_cg(lib, nels):
    """Opt-in geometric-multigrid V-cycle preconditioner (the reference only has the stale Jacobi of
    solvers_api.jl:187-203): same solution as plain CG to the solve tolerance, in far fewer iterations, for a
    high-contrast SIMP field; the hierarchy follows a change of the densities."""
    t = lib
    prob = t.PointLoadCantilever(nels)
    rho = rand_rho(prob.nel, 5) ** 2  # down to ~0.04 -> E contrast ~1e4 with p = 3
    mk = lambda pre: t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), xmin=1e-6, abstol=1e-9, reltol=0.0,
                                 cg_max_iter=50000, preconditioner=pre)
    a, b = mk(None), mk("multigrid")
    a.vars = rho
    b.vars = rho
    ua, ub = a().copy(), b().copy()
    assert a.last_result.converged == 1 and b.last_result.converged == 1
    assert b.last_result.residual <= 1e-9
    assert rel(ub, ua) < 1e-6
    assert b.last_result.iters * 8 < a.last_result.iters, (a.last_result.iters, b.last_result.iters)
    ca, cb = t.ComplianceFun(a), t.ComplianceFun(b)
    rho2 = np.clip(rho + 0.2, 0.0, 1.0)
    va, ga = ca.value_and_grad(rho2)
    vb, gb = cb.value_and_grad(rho2)  # stiffness changed: coarse operators, diagonals and bounds are rebuilt
    assert b.last_result.converged == 1 and abs(va - vb) / va < 1e-7 and rel(gb, ga) < 1e-6
    a.close()
    b.close()


@pytest.mark.parametrize("kind,nels", [("cantilever", (40, 12)), ("cantilever", (10, 6, 8)), ("heat", (24, 24))])
def test_persistent_cg_for_small_grids(lib, kind, nels, monkeypatch):
    """k_cg_persistent: batches of CG iterations of the reference recurrence in ONE cooperative launch (three device-wide
    barriers per iteration) on small grids.  Same iterates as the one-launch-per-kernel path and the oracle for
    1 / 5 / 20 / 61 iterations (61 crosses a batch boundary), same stopping iteration, far fewer launches."""
    t = lib
    prob = t.PointLoadCantilever(nels) if kind == "cantilever" else t.HeatTree(nels)
    oprob = o.PointLoadCantilever(nels) if kind == "cantilever" else o.HeatTree(nels)
    prob.Ke = oprob.Ke.copy()
    rho = rand_rho(prob.nel, 8)
    E = o.get_rho(rho, 3.0, 1e-3)

    def mk(persist, **kw):
        monkeypatch.setenv("TOPOPT_CG_PERSIST", "1" if persist else "0")
        return t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), cg_variant=0, **kw)

    for maxiter in (1, 5, 20, 61):
        a, b = mk(False, cg_max_iter=maxiter, abstol=0.0, reltol=0.0), mk(True, cg_max_iter=maxiter, abstol=0.0, reltol=0.0)
        a.vars = rho
        b.vars = rho
        ua, ub = a().copy(), b().copy()
        assert a.last_result.iters == b.last_result.iters == maxiter
        # rounding differences grow exponentially with the iteration count on these small high-contrast grids (both GPU
        # paths leave the oracle together: 1e-16 / 3e-14 / 1e-10 / 1e-4 after 5 / 20 / 40 / 61 iterations on the 40x12 grid,
        # tools/r02_probe_persist.py), so the long run only checks the batch boundary loosely
        tol = 1e-12 if maxiter <= 5 else (1e-9 if maxiter <= 20 else 5e-3)
        assert rel(ub, ua) < tol, (maxiter, rel(ub, ua))
        assert abs(a.last_result.residual - b.last_result.residual) <= 10 * tol * a.last_result.residual
        assert b.stats().kernel_launches < a.stats().kernel_launches or maxiter == 1
        uo, it, res = o.solve_matfree(oprob, E, abstol=0.0, reltol=0.0, maxiter=maxiter)
        assert rel(ub, uo) < tol
        a.close()
        b.close()
    a, b = mk(False, abstol=1e-10, reltol=0.0, cg_max_iter=20000), mk(True, abstol=1e-10, reltol=0.0, cg_max_iter=20000)
    a.vars = rho
    b.vars = rho
    ua, ub = a().copy(), b().copy()
    assert a.last_result.converged == 1 and b.last_result.converged == 1
    assert abs(a.last_result.iters - b.last_result.iters) <= 2 and rel(ub, ua) < 1e-8
    assert rel(ub, o.solve_direct(oprob, E)) < 1e-7
    a.close()
    b.close()


def test_warm_start_and_refreshed_jacobi(lib):
    """Opt-in extensions over the reference (solvers_api.jl:187-203): warm start from the resident solution
    and a Jacobi preconditioner rebuilt from the current stiffness reach the same solution in fewer
    iterations; the defaults (zero start, stale preconditioner) are unchanged."""
    t = lib
    prob, oprob, s = make(t, (16, 8, 30), abstol=1e-10, reltol=0.0, cg_max_iter=20000)
    rho1, rho2 = rand_rho(prob.nel, 1), rand_rho(prob.nel, 1) * 0.98 + 0.01
    s.vars = rho1
    s()
    s.vars = rho2
    u_cold = s().copy()
    it_cold = s.last_result.iters
    w = t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), abstol=1e-10, reltol=0.0, cg_max_iter=20000, warm_start=True)
    w.vars = rho1
    w()
    w.vars = rho2
    u_warm = w().copy()
    assert w.last_result.converged == 1 and w.last_result.iters < it_cold
    assert rel(u_warm, u_cold) < 1e-7
    uref = o.solve_direct(oprob, o.get_rho(rho2, 3.0, 1e-3))
    assert rel(u_warm, uref) < 1e-7
    # refreshed Jacobi: same solution as the oracle's direct solve, fewer iterations than the stale one
    j0 = t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), abstol=1e-10, reltol=0.0, cg_max_iter=20000, preconditioner="jacobi")
    j1 = t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), abstol=1e-10, reltol=0.0, cg_max_iter=20000, preconditioner="jacobi",
                     refresh_preconditioner=True)
    for j in (j0, j1):
        j.vars = np.full(prob.nel, 0.5)
        j()  # the reference builds the preconditioner here, once
        j.vars = rho2
        j()
        assert j.last_result.converged == 1
        assert rel(j.u, uref) < 1e-7
    assert j1.last_result.iters <= j0.last_result.iters
    for q in (s, w, j0, j1):
        q.close()


def test_graph_cache_survives_solution_swap(lib):
    """the CUDA-graph cache key must cover every buffer baked into the captured launches.
    solve(rhs) -> swap -> solve(rhs) with several graph batches per solve must still match the oracle."""
    t = lib
    prob, oprob, s = make(t, (10, 4, 6), abstol=1e-12, reltol=1e-14, cg_max_iter=5000, check_every=5)
    rho = rand_rho(prob.nel, 4)
    E = o.get_rho(rho, 3.0, 1e-3)
    s.vars = rho
    rhs = np.random.default_rng(9).standard_normal(prob.ndof)
    uref = o.solve_direct(oprob, E, rhs=rhs)
    for k in range(3):
        out = np.zeros(prob.ndof)
        s(rhs=rhs, lhs=out)
        assert s.last_result.converged == 1 and s.last_result.iters > 10
        assert rel(out, uref) < 1e-8
        s._check(s._lib.topopt_swap_solution_lambda(s.handle))
    s.close()


@pytest.mark.skipif(os.environ.get("TOPOPT_SKIP_FULL_SIZE") == "1", reason="full-size config 4 disabled")
def test_config4_full_size_against_c_port(lib):
    """BASELINE config 4 at full size (256x128x128 hex8, 12 830 211 dofs), the OpenMP C port of the
    reference CPU path as the checker: K.u through every shipped kernel <= 1e-12, ten CG iterates of
    both recurrences <= 1e-10 (relative to max|u|), residual norms <= 1e-9."""
    import ref_c

    t = lib
    nels = (256, 128, 128)
    prob = t.PointLoadCantilever(nels)
    R = ref_c.RefProblem(3, 3, nels, prob.Ke, prob.prescribed_dofs, openmp=True, native=True)
    nel = prob.nel
    # SURVEY 8d density field (B): deterministic, cheap to build at 4.2 M elements
    rho = 0.2 + 0.8 * ((np.arange(nel, dtype=np.int64) * 2654435761 % 1000003) / 1000003.0)
    R.set_density(rho, 3.0, 1e-3)
    s = t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), cg_max_iter=10, abstol=0.0, reltol=0.0)
    s.set_density(rho)
    x = np.random.default_rng(7).standard_normal(prob.ndof)
    x[prob.prescribed_dofs - 1] = 0.0
    yref = R.mul(x)
    for k in (0, 2, 3, 4):
        y, xy, _ = s.mul_ex(x, k)
        assert rel(y, yref) < RTOL_OP, f"kernel {k}"
        assert abs(xy - float(x @ yref)) <= 1e-10 * float(np.abs(x) @ np.abs(yref))
    b = prob.fixedload.copy()
    b[prob.prescribed_dofs - 1] = 0.0
    uref, it, res = R.cg(b, abstol=0.0, reltol=0.0, maxiter=10)
    assert it == 10
    for variant in (0, 1):
        s.cg_variant = variant
        s.vars = rho
        u = s().copy()
        assert s.last_result.iters == 10
        assert rel(u, uref) < 1e-10
        assert abs(s.last_result.residual - res) <= 1e-9 * res
    R.close()
    s.close()


@pytest.mark.skipif(os.environ.get("TOPOPT_SKIP_FULL_SIZE") == "1", reason="full-size config 4 disabled")
def test_config4_converged_solution_verified_by_c_port(lib):
    """North-star acceptance at full size: the multigrid-preconditioned solve of config 4 (12.8 M dofs) to |r| <= 1e-8 is
    checked a posteriori with the C port of the reference operator: |f - K_port u| <= 2e-8 (the port's own matrix-free
    mul!, matrix_free_operator.jl:66-105), compliance u'K_port u = f'u to 1e-9, and the compliance of the plain-CG
    solve at the same tolerance agrees to 1e-8."""
    import ref_c

    t = lib
    nels = (256, 128, 128)
    prob = t.PointLoadCantilever(nels)
    R = ref_c.RefProblem(3, 3, nels, prob.Ke, prob.prescribed_dofs, openmp=True, native=True)
    rho = 0.2 + 0.8 * ((np.arange(prob.nel, dtype=np.int64) * 2654435761 % 1000003) / 1000003.0)
    R.set_density(rho, 3.0, 1e-6)
    b = prob.fixedload.copy()
    b[prob.prescribed_dofs - 1] = 0.0
    mk = lambda pre, mi: t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), xmin=1e-6, abstol=1e-8, reltol=0.0,
                                     cg_max_iter=mi, preconditioner=pre, cg_variant=1)
    s = mk("multigrid", 500)
    s.vars = rho
    u = s().copy()
    assert s.last_result.converged == 1 and s.last_result.iters < 100
    r = b - R.mul(u)
    r[prob.prescribed_dofs - 1] = 0.0
    assert np.linalg.norm(r) <= 2e-8
    c_mg = float(b @ u)
    assert abs(float(u @ R.mul(u)) - c_mg) <= 1e-9 * c_mg
    s.close()
    s = mk(None, 30000)
    s.vars = rho
    u2 = s().copy()
    assert s.last_result.converged == 1
    assert abs(float(b @ u2) - c_mg) <= 1e-8 * c_mg
    R.close()
    s.close()


def test_plain_c_harness_drives_the_abi(lib, tmp_path):
    """The boundary is a C ABI: a plain C99 program (tests/c/harness.c) creates the handle, sets the density,
    solves, evaluates compliance / sensitivities / u'Ku, filters and runs the fused SIMP evaluation -- on every
    visible device (up to two, one handle each in the same process: per-device kernel attributes) -- and its
    numbers must match the oracle."""
    import subprocess

    import torch

    t = lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "topopt.jl_b200")
    exe = tmp_path / "harness"
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-O1", "-I", os.path.join(root, "include"),
           os.path.join(root, "tests", "c", "harness.c"), "-o", str(exe), "-L", libdir, "-ltopopt_cuda", f"-Wl,-rpath,{libdir}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    ndev = min(2, torch.cuda.device_count())
    out = subprocess.run([str(exe), str(ndev)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    nels = (12, 4, 28)
    oprob = o.PointLoadCantilever(nels)
    ne = oprob.nel
    e = np.arange(ne, dtype=np.uint64)
    rho = 0.2 + 0.8 * ((e * np.uint64(2654435761)) % np.uint64(1000)).astype(np.float64) / 1000.0
    E = o.get_rho(rho, 3.0, 1e-3)
    uref = o.solve_direct(oprob, E)
    obj, _, g = o.compliance(oprob, uref, rho, 3.0, 1e-3)
    Fo = o.DensityFilter(oprob, 2.0)
    xf = Fo(rho)
    uf = o.solve_direct(oprob, o.get_rho(xf, 3.0, 1e-3))
    obj_f, _, _ = o.compliance(oprob, uf, xf, 3.0, 1e-3)
    lines = [ln.split() for ln in out.stdout.strip().splitlines()]
    assert len(lines) == ndev
    for ln in lines:
        vals = [float(v) for v in ln[1:]]
        assert abs(vals[0] - obj) / obj < 1e-8
        assert abs(vals[1] - obj_f) / obj_f < 1e-8
        assert abs(vals[2] - float(g @ (1 + np.arange(ne) % 7))) <= 1e-8 * float(np.abs(g) @ (1 + np.arange(ne) % 7))
        assert abs(vals[3] - float(xf @ (1 + np.arange(ne) % 5))) <= 1e-12 * float(xf @ (1 + np.arange(ne) % 5))
        assert abs(vals[5] - obj) / obj < 1e-8  # u'Ku == sum E_e u_e'Ke u_e


@pytest.mark.parametrize("kind", ["displacement", "temperature"])
def test_displacement_and_temperature_fun_adjoints(lib, kind):
    """DisplacementFun / TemperatureFun (displacement.jl:97-121, temperature.jl:92-119): the value is
    the solved field; the pullback of a cotangent delta is one adjoint solve K lam = delta (arbitrary right-hand side,
    apply_zero!) and out_e = -dE_e u_e' Ke lam_e.  Checked against the oracle and against a finite difference."""
    t = lib
    if kind == "displacement":
        nels = (10, 4, 6)
        prob, oprob = t.PointLoadCantilever(nels), o.PointLoadCantilever(nels)
        Fun = t.DisplacementFun
    else:
        nels = (12, 9)
        prob, oprob = t.HeatTree(nels), o.HeatTree(nels)
        Fun = t.TemperatureFun
    prob.Ke = oprob.Ke.copy()
    s = t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), abstol=1e-13, reltol=1e-15, cg_max_iter=50000)
    f = Fun(s)
    rho = rand_rho(prob.nel, 3)
    u = f(rho)
    E, dE = o.get_rho_drho(rho, 3.0, 1e-3)
    uref = o.solve_direct(oprob, E)
    assert rel(u, uref) < 1e-8
    delta = np.random.default_rng(8).standard_normal(prob.ndof)
    g = f.pullback(delta)
    lam = o.solve_direct(oprob, E, rhs=delta)
    gref = -dE * o.element_energy(oprob, uref, lam)
    assert rel(g, gref) < 1e-8
    # directional finite difference of delta . u(rho)
    d = np.random.default_rng(9).standard_normal(prob.nel)
    h = 1e-6
    up = o.solve_direct(oprob, o.get_rho(rho + h * d, 3.0, 1e-3))
    um = o.solve_direct(oprob, o.get_rho(rho - h * d, 3.0, 1e-3))
    fd = float(delta @ (up - um)) / (2 * h)
    free = np.ones(prob.ndof, dtype=bool)
    free[oprob.prescribed] = False
    fd_free = float(delta[free] @ (up - um)[free]) / (2 * h)
    assert abs(float(g @ d) - fd_free) <= 1e-5 * max(abs(fd_free), 1e-12)
    # wrong physics is an ArgumentError in the reference
    Other = t.TemperatureFun if kind == "displacement" else t.DisplacementFun
    with pytest.raises(ValueError):
        Other(s)
    s.close()


@pytest.mark.parametrize("nels,filt", [((16, 8), "density"), ((10, 4, 6), "density"), ((16, 8), "sens")])
def test_device_resident_simp_loop(lib, nels, filt):
    """the optimality-criteria update (bisection on the volume multiplier as device reductions) keeps the
    design on the GPU between iterations.  Ten iterations must reproduce the host-side loop (same update rule) and the
    oracle's loop to the north-star design tolerance."""
    t = lib
    prob, oprob = t.PointLoadCantilever(nels), o.PointLoadCantilever(nels)
    prob.Ke = oprob.Ke.copy()
    mk = lambda: t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), abstol=1e-11, reltol=1e-14, cg_max_iter=20000, xmin=1e-3)
    s1, s2 = mk(), mk()
    Fc = {"density": t.DensityFilterFun, "sens": t.SensFilterFun}[filt]
    F1, F2 = Fc(s1, 2.0), Fc(s2, 2.0)
    xh, hh = t.simp_loop(s1, F1, 0.4, iters=10)
    xd, hd = t.simp_loop_device(s2, F2, 0.4, iters=10)
    assert np.max(np.abs(xd - xh)) < 1e-9 and rel(hd, hh) < 1e-10
    xo, ho = o.simp_loop(oprob, 2.0, 0.4, p=3.0, xmin=1e-3, iters=10, filt=filt, abstol=1e-11, maxiter=20000)
    assert np.max(np.abs(xd - xo)) < 1e-6 and rel(hd, ho) < 1e-8
    assert abs(float(xd @ (prob.cellvolumes / prob.cellvolumes.sum())) - 0.4) < 1e-9 or filt == "density"
    # VTK export straight from the device-resident design (VTK.jl:33-62): same cells as from the host copy
    assert np.array_equal(t.device_design(s2), xd)
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        (fa,) = t.save_mesh(os.path.join(d, "dev"), prob, solver=s2, nodal=True)
        (fb,) = t.save_mesh(os.path.join(d, "host"), prob, xd, nodal=s2.u)
        assert open(fa).read() == open(fb).read() and f'NumberOfCells="{int((xd >= 0.5).sum())}"' in open(fa).read()
    # one update against the host rule on arbitrary inputs
    x = np.random.default_rng(1).uniform(0.05, 1.0, prob.nel)
    dc = -np.random.default_rng(2).uniform(0.0, 3.0, prob.nel)
    dv = np.random.default_rng(3).uniform(0.5, 1.5, prob.nel) / prob.nel
    out = np.empty(prob.nel)
    change, nb = t.oc_update_device(s2, 0.45, dv=dv, x=x, dc=dc, x_out=out)
    ref = t.oc_update(x, dc, dv, 0.45)
    assert np.max(np.abs(out - ref)) < 1e-10 and nb > 30
    assert abs(change - float(np.linalg.norm(ref - x))) < 1e-9
    for q in (F1, F2):
        q.close()
    s1.close(); s2.close()
