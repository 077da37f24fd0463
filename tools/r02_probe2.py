"""Quick K.u probe: which in argv[2:] at grid argv[1]."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import topopt_jl_b200 as t

nels = tuple(int(v) for v in sys.argv[1].split(","))
whiches = [int(v) for v in sys.argv[2:]] or [7, 0, 8]
prob = t.PointLoadCantilever(nels)
s = t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), xmin=1e-6)
s.set_density(np.random.default_rng(0).uniform(0.2, 1.0, prob.nel))
bytes_kxu = 16 * prob.ndof + 8 * prob.nel
for w in whiches:
    s.time_kernel(w, 3)
    ms = min(s.time_kernel(w, 30) for _ in range(3))
    print(f"which={w} {ms * 1e3:9.2f} us" + (f" {bytes_kxu / ms / 1e6:8.1f} GB/s alg" if w in (0, 7, 8) else ""), flush=True)
s.close()
