"""Multigrid-PCG vs plain CG at one grid: iterations and device time to a tolerance.
usage: r02_probe_mg.py nx,ny,nz abstol "deg:ratio" ..."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import topopt_jl_b200 as t

nels = tuple(int(v) for v in sys.argv[1].split(","))
abstol = float(sys.argv[2])
prob = t.PointLoadCantilever(nels)
fields = {
    "uniform 0.3": np.full(prob.nel, 0.3),
    "random [0.05,1]^1": np.random.default_rng(0).uniform(0.05, 1.0, prob.nel),
    "0/1 blocks": (np.add.outer(np.add.outer(np.arange(nels[2]) // 8, np.arange(nels[1]) // 8), np.arange(nels[0]) // 8) % 2).astype(float).ravel() * 0.999 + 0.001,
}
for name, rho in fields.items():
    s = t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), xmin=1e-6, abstol=abstol, reltol=0.0, cg_max_iter=30000, cg_variant=1)
    s.vars = rho
    s(download=False)
    r = s.last_result
    print(f"{name:20s} plain CG        iters {r.iters:6d} conv {r.converged} res {r.residual:.2e} solve {r.solve_ms:9.2f} ms", flush=True)
    s.close()
    for cfg in sys.argv[3:]:
        deg, ratio = cfg.split(":")
        s = t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), xmin=1e-6, abstol=abstol, reltol=0.0, cg_max_iter=2000,
                        preconditioner="multigrid")
        s.mg_degree, s.mg_ratio = int(deg), float(ratio)
        s.vars = rho
        t0 = time.perf_counter()
        s(download=False)
        w = time.perf_counter() - t0
        r = s.last_result
        s(download=False)  # hierarchy already built: solve only
        r2 = s.last_result
        print(f"{name:20s} MG deg {deg} ratio {ratio:>4s} iters {r.iters:6d} conv {r.converged} res {r.residual:.2e} solve {r.solve_ms:9.2f} ms "
              f"(wall incl. setup {w * 1e3:8.1f} ms; resolve {r2.solve_ms:8.2f} ms)", flush=True)
        s.close()
