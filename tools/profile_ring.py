"""ncu target: a few launches of one K.u kernel class at config 4 (which = 7 previous generation, 0/8 ring)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import topopt_jl_b200 as t

nels = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "256,128,128").split(","))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 8
prob = t.PointLoadCantilever(nels)
s = t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), xmin=1e-6)
s.set_density(np.random.default_rng(0).uniform(0.2, 1.0, prob.nel))
print(which, s.time_kernel(which, 3))
s.close()
