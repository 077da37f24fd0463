"""Launch each kernel class of the hot path a few times on the config-4 grid so that ncu can
capture them (used only for profiling; never for a benchmark number)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import topopt_jl_b200 as t

nels = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "256,128,128").split(","))
which = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "0,1,2,3").split(",")]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
prob = t.PointLoadCantilever(nels) if len(nels) == 3 else t.HalfMBB(nels)
Solver = t.CUDAAssemblySolver if any(w in (4, 5, 6) for w in which) else t.CUDAMatrixFreeSolver
s = t.FEASolver(Solver, prob, penalty=t.PowerPenaltyFun(3.0), xmin=1e-6)
F = t.DensityFilterFun(s, 2.0)
s.set_density(np.full(prob.nel, 0.3))
for w in which:
    print(w, s.time_kernel(w, reps, F), "ms")
