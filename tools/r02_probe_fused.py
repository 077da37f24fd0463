"""CG iteration time (topopt_time_kernel class 9: single-pass CG on a dense rhs), one-kernel vs two-kernel iteration.
Each argv[2:] is "K=V,K=V" (environment for that run)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import topopt_jl_b200 as t

nels = tuple(int(v) for v in sys.argv[1].split(","))
prob = t.PointLoadCantilever(nels)
rho = np.random.default_rng(0).uniform(0.2, 1.0, prob.nel)
for cfg in sys.argv[2:]:
    kv = dict(p.split("=") for p in cfg.split(",") if p)
    for k, v in kv.items():
        os.environ[k] = v
    s = t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), xmin=1e-6, check_every=100)
    s.set_density(rho)
    s.time_kernel(9, 20)
    ms = min(s.time_kernel(9, 200) for _ in range(3))
    fused = os.environ.get("TOPOPT_CG_FUSED", "1") != "0"
    nbytes = (64 if fused else 72) * prob.ndof + 8 * prob.nel
    print(f"{cfg:40s} {ms * 1e3:9.2f} us/iteration {nbytes / ms / 1e6:8.1f} GB/s", flush=True)
    s.close()
    for k in kv:
        os.environ.pop(k, None)
