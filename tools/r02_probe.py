"""Round-2 probe: K.u kernel generations and CG iteration variants at a given grid (CUDA events on the
library's stream, dense vectors).  usage: python tools/r02_probe.py [nx,ny,nz] [reps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import topopt_jl_b200 as t

nels = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "256,128,128").split(","))
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
prob = t.PointLoadCantilever(nels)
rho = np.random.default_rng(0).uniform(0.2, 1.0, prob.nel)
bytes_kxu = 16 * prob.ndof + 8 * prob.nel
out = {"nels": nels, "ndof": prob.ndof}


def run(tag, env, whiches):
    for k, v in env.items():
        os.environ[k] = v
    s = t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(3.0), xmin=1e-6)
    s.set_density(rho)
    for w in whiches:
        try:
            s.time_kernel(w, 3)
            ms = min(s.time_kernel(w, reps) for _ in range(3))
        except Exception as e:  # noqa: BLE001
            ms = None
            print(tag, w, "failed:", e, flush=True)
        out[f"{tag}:{w}"] = ms
        if ms:
            extra = f" {bytes_kxu / ms / 1e6:8.1f} GB/s alg" if w in (0, 7, 8) else ""
            print(f"{tag:28s} which={w} {ms * 1e3:9.2f} us{extra}", flush=True)
    s.close()
    for k in env:
        os.environ.pop(k, None)


run("default", {}, [7, 0, 8, 1, 9])
for tyt in (8, 10):
    run(f"ring_tyt{tyt}", {"TOPOPT_KXU_RING": str(tyt)}, [8])
run("waves2", {"TOPOPT_KXU_WAVES": "2"}, [8])
print(json.dumps(out))
