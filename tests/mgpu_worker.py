"""Worker for the multi-GPU parity test (launched with torchrun, one rank per GPU): the z-slab
sharded solve / compliance / sensitivity / filter must agree with the CPU oracle on every rank."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np

import topopt_jl_b200 as t
import topopt_oracle as o
from topopt_jl_b200 import distributed as D


def rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def main():
    comm, local = D.init_from_env()
    assert comm is not None, "run under torchrun with >= 2 ranks"
    for nels in ((10, 4, 8), (6, 4, 16), (16, 10)):
        prob, oprob = t.PointLoadCantilever(nels), o.PointLoadCantilever(nels)
        prob.Ke = oprob.Ke.copy()
        s = t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), abstol=1e-12, reltol=1e-14,
                        cg_max_iter=20000, device=local, comm=comm)
        rho = np.random.default_rng(42).uniform(0.2, 1.0, prob.nel)
        E = o.get_rho(rho, 3.0, 1e-3)
        s.set_density(rho)
        x = np.random.default_rng(1).standard_normal(prob.ndof)
        assert rel(s.mul(x), o.matfree_mul(oprob, E, x)) < 1e-12
        comp = t.ComplianceFun(s)
        val, grad = comp.value_and_grad(rho)
        uref = o.solve_direct(oprob, E)
        obj, cc, g = o.compliance(oprob, uref, rho, 3.0, 1e-3)
        assert s.last_result.converged == 1
        assert abs(val - obj) / obj < 1e-8 and rel(grad, g) < 1e-8 and rel(s.u, uref) < 1e-7
        # few-iteration recurrence parity across the slab boundary
        s2 = t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), abstol=0.0, reltol=0.0,
                         cg_max_iter=10, device=local, comm=comm)
        s2.vars = rho
        u10 = s2().copy()
        uo, it, res = o.solve_matfree(oprob, E, abstol=0.0, reltol=0.0, maxiter=10)
        assert s2.last_result.iters == 10 and rel(u10, uo) < 1e-9
        F = t.DensityFilterFun(s, 2.0)
        gx = np.empty(prob.nel)
        objf, _ = t.simp_eval(s, F, rho, gx)
        Fo = o.DensityFilter(oprob, 2.0)
        xf = Fo(rho)
        uf = o.solve_direct(oprob, o.get_rho(xf, 3.0, 1e-3))
        of, _, gf = o.compliance(oprob, uf, xf, 3.0, 1e-3)
        assert abs(objf - of) / of < 1e-8 and rel(gx, Fo.pullback(gf)) < 1e-8
        F.close(); s.close(); s2.close()
    # thick slabs (>= 100 node planes per rank): the kernels selected for the large configurations, with their
    # peer-memory halo reads -- ring-staged (bulk copies straight from the neighbours' memory) and, with the ring
    # kernel disabled, the two-row shuffle kernel; both CG recurrences
    nels = (8, 4, 100 * comm.world)
    prob, oprob = t.PointLoadCantilever(nels), o.PointLoadCantilever(nels)
    prob.Ke = oprob.Ke.copy()
    rho = np.random.default_rng(7).uniform(0.2, 1.0, prob.nel)
    E = o.get_rho(rho, 3.0, 1e-3)
    uref = o.solve_direct(oprob, E)
    obj, cc, g = o.compliance(oprob, uref, rho, 3.0, 1e-3)
    uo, it, res = o.solve_matfree(oprob, E, abstol=0.0, reltol=0.0, maxiter=10)
    for env in ({}, {"TOPOPT_KXU_RING": "0"}):
        os.environ.update(env)
        for variant in (0, 1):
            s = t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), abstol=1e-12, reltol=1e-14,
                            cg_max_iter=50000, device=local, comm=comm, cg_variant=variant)
            comp = t.ComplianceFun(s)
            val, grad = comp.value_and_grad(rho)
            assert s.last_result.converged == 1, (env, variant)
            assert abs(val - obj) / obj < 1e-8 and rel(grad, g) < 1e-8 and rel(s.u, uref) < 1e-7, (env, variant, abs(val - obj) / obj)
            s2 = t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), abstol=0.0, reltol=0.0,
                             cg_max_iter=10, device=local, comm=comm, cg_variant=variant)
            s2.vars = rho
            u10 = s2().copy()
            assert s2.last_result.iters == 10 and rel(u10, uo) < 1e-9, (env, variant, rel(u10, uo))
            s.close(); s2.close()
        for k in env:
            os.environ.pop(k, None)
    # one-kernel CG iteration across ranks (ghost planes of p, r, Ap of both parities read from the neighbours): dense
    # right-hand side, multi-tile grid, against the two-kernel path and the oracle
    nels = (40, 26, 30 * comm.world)
    prob, oprob = t.PointLoadCantilever(nels), o.PointLoadCantilever(nels)
    prob.Ke = oprob.Ke.copy()
    rho = np.random.default_rng(9).uniform(0.2, 1.0, prob.nel)
    b = np.random.default_rng(10).standard_normal(prob.ndof)
    b[prob.prescribed_dofs - 1] = 0.0
    for maxiter in (1, 3, 12, 40):
        sols = []
        for fused in ("1", "0"):
            os.environ["TOPOPT_CG_FUSED_MGPU"] = fused
            s = t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), abstol=0.0, reltol=0.0, cg_max_iter=maxiter,
                            device=local, comm=comm, cg_variant=1)
            s.set_density(rho)
            sols.append((s(rhs=b).copy(), s.last_result.residual, s.stats().kernel_launches))
            assert s.last_result.iters == maxiter
            s.close()
        os.environ.pop("TOPOPT_CG_FUSED_MGPU", None)
        (uf, rf, lf), (ua, ra, la) = sols
        assert rel(uf, ua) < (1e-12 if maxiter <= 12 else 1e-9) and abs(rf - ra) <= 1e-11 * ra, (maxiter, rel(uf, ua), rf, ra)
        assert lf < la or maxiter == 1, (lf, la)  # the fused path really ran
        if maxiter == 3:
            uo3, it, res = o.solve_matfree(oprob, o.get_rho(rho, 3.0, 1e-3), abstol=0.0, reltol=0.0, maxiter=3, rhs=b)
            assert rel(uf, uo3) < 1e-12 and abs(rf - res) <= 1e-11 * res
    import torch.distributed as dist

    dist.barrier()
    if comm.rank == 0:
        print(f"MGPU_OK world={comm.world}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
