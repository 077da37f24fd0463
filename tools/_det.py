import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "oracle"))
import numpy as np
import topopt_jl_b200 as t
import topopt_oracle as o
for nels in ((60, 20, 20), (33, 8, 4), (64, 30, 18)):
    prob, oprob = t.PointLoadCantilever(nels), o.PointLoadCantilever(nels)
    prob.Ke = oprob.Ke.copy()
    for ty in (4, 6, 8, 12):
        os.environ["TOPOPT_KXU_TY"] = str(ty); os.environ["TOPOPT_KXU_ZC"] = "7"
        s = t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), xmin=1e-6)
        rho = np.random.default_rng(1).uniform(0.2, 1, prob.nel)
        s.set_density(rho)
        x = np.random.default_rng(0).standard_normal(prob.ndof)
        y = s.mul(x); yr = o.matfree_mul(oprob, o.get_rho(rho, 3.0, 1e-6), x)
        print(nels, ty, "rel err", float(np.max(np.abs(y - yr)) / np.max(np.abs(yr))))
        s.close()
