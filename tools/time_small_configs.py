"""Kernel timings of BASELINE configs 1, 2, 3, 5 (latency / L2-resident cases; profiling helper)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import topopt_jl_b200 as t

cases = [
    ("cfg1 cantilever 160x40 quad4 matrix-free", t.PointLoadCantilever((160, 40)), t.CUDAMatrixFreeSolver, [0, 1, 2, 3]),
    ("cfg2 HalfMBB 600x200 quad4 assembled", t.HalfMBB((600, 200)), t.CUDAAssemblySolver, [0, 4, 5, 6, 2, 3]),
    ("cfg3 cantilever 60x20x20 hex8 matrix-free", t.PointLoadCantilever((60, 20, 20)), t.CUDAMatrixFreeSolver, [0, 1, 2, 3]),
    ("cfg5 HeatTree 1024x1024 quad4 scalar", t.HeatTree((1024, 1024)), t.CUDAMatrixFreeSolver, [0, 1, 2, 3]),
]
names = {0: "K.u matrix-free", 1: "CG iteration (matrix-free)", 2: "sensitivity", 3: "filter forward", 4: "SpMV", 5: "assembly", 6: "CG iteration (assembled)"}
for title, prob, S, which in cases:
    s = t.FEASolver(S, prob, penalty=t.PowerPenaltyFun(3.0))
    F = t.DensityFilterFun(s, 2.0)
    s.set_density(np.full(prob.nel, 0.5))
    md = prob.metadata
    print(f"## {title}: ndof={md.ndof} nel={md.nel} nnz={md.nnz}")
    for w in which:
        s.time_kernel(w, 5, F)
        ms = s.time_kernel(w, 50, F)
        extra = ""
        if w == 0:
            extra = f"  {(16 * md.ndof + 8 * md.nel) / ms / 1e6:8.1f} GB/s algorithmic"
        if w == 4:
            extra = f"  {(12 * md.nnz + 20 * md.ndof) / ms / 1e6:8.1f} GB/s algorithmic"
        print(f"  {names[w]:28s} {ms * 1e3:8.2f} us{extra}")
    F.close()
    s.close()
