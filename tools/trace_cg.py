"""Per-phase device timing of the CG iteration (TOPOPT_TRACE=1), single- or multi-GPU."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["TOPOPT_TRACE"] = "1"
import numpy as np
import topopt_jl_b200 as t
from topopt_jl_b200 import distributed as D
comm, local = D.init_from_env()
import torch
torch.cuda.set_device(local)
prob = t.PointLoadCantilever((256, 128, 128))
s = t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), xmin=1e-6, device=local, comm=comm)
s.set_density(np.full(prob.nel, 0.3))
for _ in range(3):
    print("rank", 0 if comm is None else comm.rank, "cg iteration ms", s.time_kernel(1, 40), flush=True)
if comm is not None:
    torch.distributed.barrier(); torch.distributed.destroy_process_group()
