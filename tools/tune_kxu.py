"""Sweep the tile parameters of the modal hex8 K.u kernel (profiling helper, not a benchmark)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import topopt_jl_b200 as t

nels = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "256,128,128").split(","))
tys = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "8,12,16").split(",")]
nsyncs = [int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "0,1").split(",")]
prob = t.PointLoadCantilever(nels)
bytes_ = 16 * prob.ndof + 8 * prob.nel
for ty in tys:
    for ns in nsyncs:
        os.environ["TOPOPT_KXU_TY"] = str(ty)
        os.environ["TOPOPT_KXU_NSYNC"] = str(ns)
        s = t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), xmin=1e-6)
        s.set_density(np.full(prob.nel, 0.3))
        s.time_kernel(0, 3)
        ms = s.time_kernel(0, 20)
        cg = s.time_kernel(1, 20)
        print(f"TY={ty:2d} NSYNC={ns}  K.u {ms*1e3:7.1f} us  {bytes_/ms/1e6:7.1f} GB/s  {prob.ndof/ms/1e6:6.1f} GDOF/s | CG iteration {cg*1e3:7.1f} us", flush=True)
        s.close()
