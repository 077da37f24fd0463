import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import topopt_jl_b200 as t
import topopt_oracle as o
rel = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
nels = (40, 12)
prob, oprob = t.PointLoadCantilever(nels), o.PointLoadCantilever(nels)
prob.Ke = oprob.Ke.copy()
rho = np.random.default_rng(8).uniform(0.2, 1.0, prob.nel)
E = o.get_rho(rho, 3.0, 1e-3)
for maxiter in (1, 5, 10, 20, 40, 50, 51, 61, 100):
    sols = []
    for persist in ("0", "1"):
        os.environ["TOPOPT_CG_PERSIST"] = persist
        s = t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), cg_variant=0, cg_max_iter=maxiter, abstol=0.0, reltol=0.0)
        s.vars = rho
        sols.append((s().copy(), s.last_result.residual))
        s.close()
    uo, it, res = o.solve_matfree(oprob, E, abstol=0.0, reltol=0.0, maxiter=maxiter)
    (ua, ra), (ub, rb) = sols
    print(maxiter, "kernel-vs-oracle %.2e persist-vs-oracle %.2e kernel-vs-persist %.2e  res %.6e %.6e %.6e" % (rel(ua, uo), rel(ub, uo), rel(ua, ub), ra, rb, res), flush=True)
