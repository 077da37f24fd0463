"""Sweep variants of the modal hex8 K.u kernel (profiling helper, not a benchmark).
usage: tune_kxu.py NELS "TY list" "2ROW list"   (2ROW = 0 selects the one-row kernel with TY rows)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import topopt_jl_b200 as t

nels = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "256,128,128").split(","))
tys = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "16").split(",")]
rows2 = [int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "0,6,8,10").split(",")]
prob = t.PointLoadCantilever(nels)
bytes_ = 16 * prob.ndof + 8 * prob.nel
for r2 in rows2:
    for ty in (tys if r2 == 0 else [0]):
        os.environ["TOPOPT_KXU_TY"] = str(ty)
        os.environ["TOPOPT_KXU_2ROW"] = str(r2)
        s = t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), xmin=1e-6)
        s.set_density(np.full(prob.nel, 0.3))
        s.time_kernel(0, 3)
        s.time_kernel(1, 3)
        ms = s.time_kernel(0, 20)
        cg = s.time_kernel(1, 20)
        print(f"2ROW={r2:2d} TY={ty:2d}  K.u {ms*1e3:7.1f} us  {bytes_/ms/1e6:7.1f} GB/s  {prob.ndof/ms/1e6:6.1f} GDOF/s | CG iteration {cg*1e3:7.1f} us", flush=True)
        s.close()
