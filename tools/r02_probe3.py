"""K.u ring-kernel sweep over thread rows per CTA (env TOPOPT_KXU_RING) at one grid."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import topopt_jl_b200 as t

nels = tuple(int(v) for v in sys.argv[1].split(","))
prob = t.PointLoadCantilever(nels)
rho = np.random.default_rng(0).uniform(0.2, 1.0, prob.nel)
bytes_kxu = 16 * prob.ndof + 8 * prob.nel
for tyt in sys.argv[2:]:
    os.environ["TOPOPT_KXU_RING"] = tyt
    s = t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), xmin=1e-6)
    s.set_density(rho)
    for w in (8,):
        s.time_kernel(w, 3)
        ms = min(s.time_kernel(w, 30) for _ in range(3))
        print(f"TYT={tyt} which={w} {ms * 1e3:9.2f} us {bytes_kxu / ms / 1e6:8.1f} GB/s alg", flush=True)
    s.close()
