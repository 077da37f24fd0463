"""Small end-to-end run of every kernel class for compute-sanitizer (memcheck / racecheck)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import topopt_jl_b200 as t

for mk, solver_t in ((lambda: t.PointLoadCantilever((34, 16, 4)), t.CUDAMatrixFreeSolver), (lambda: t.HalfMBB((20, 12)), t.CUDAAssemblySolver),
                     (lambda: t.HeatTree((16, 12)), t.CUDAMatrixFreeSolver)):
    prob = mk()
    s = t.FEASolver(solver_t, prob, penalty=t.PowerPenaltyFun(3.0), abstol=1e-9, cg_max_iter=200)
    rho = np.random.default_rng(0).uniform(0.2, 1.0, prob.nel)
    F = t.DensityFilterFun(s, 2.0)
    g = np.empty(prob.nel)
    if prob.ncomp == 1:
        tc = t.ThermalComplianceFun(s)
        print("thermal", tc(F(rho)))
    else:
        obj, res = t.simp_eval(s, F, rho, g)
        print("simp_eval", obj, res.iters)
    print("pullback", float(F.pullback(g if prob.ncomp > 1 else rho).sum()))
    F.close()
    s.close()
print("SANITIZE_RUN_OK")
