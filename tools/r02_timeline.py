"""Device-side timeline of two consecutive one-kernel CG iterations (needs a library built with -DTOPOPT_TIMELINE):
per CTA start / first plane ready / compute end / return from the reduction, per rank.  Run under torchrun for ranks > 1."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import topopt_jl_b200 as t
from topopt_jl_b200 import distributed as D

comm, local = D.init_from_env() if "RANK" in os.environ else (None, 0)
nels = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "256,128,128").split(","))
prob = t.PointLoadCantilever(nels)
s = t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), xmin=1e-6, device=local, comm=comm, cg_variant=1, check_every=100,
                abstol=0.0, reltol=0.0, cg_max_iter=450)
s.vars = np.full(prob.nel, 0.3)
s(download=False)  # batches of 100 iterations: 300 and 301 are graph-replayed launches
ms = s.last_result.solve_ms / s.last_result.iters
buf = np.zeros((2, 512, 4), dtype=np.uint64)
fn = s._lib.topopt_debug_timeline
fn.argtypes = [C.c_void_p, C.c_void_p]
assert fn(s.handle, buf.ctypes.data) == 0
rank = 0 if comm is None else comm.rank
for slot in range(2):
    b = buf[slot]
    used = b[:, 0] > 0
    n = int(used.sum())
    t0 = b[used, 0].min()
    ev = (b[used].astype(np.int64) - np.int64(t0)) / 1e3  # us
    q = lambda a: f"min {a.min():7.1f} med {np.median(a):7.1f} max {a.max():7.1f}"
    print(f"[rank {rank}] iteration {300 + slot}: {n} CTAs | start {q(ev[:, 0])} | first plane ready {q(ev[:, 1] - ev[:, 0])} after start | "
          f"compute end {q(ev[:, 2])} | returned {q(ev[:, 3])} | reduction+all-reduce tail {ev[:, 3].max() - ev[:, 2].max():6.1f} us")
    if slot == 1:
        gap = (buf[1][used, 0].min().astype(np.int64) - buf[0][buf[0][:, 0] > 0, 3].max().astype(np.int64)) / 1e3
        print(f"[rank {rank}] launch gap (last return of iteration 300 -> first CTA start of 301): {gap:6.1f} us; in-loop {ms * 1e3:7.1f} us / iteration")
s.close()
if comm is not None:
    import torch.distributed as dist

    dist.barrier()
    dist.destroy_process_group()
