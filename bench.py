#!/usr/bin/env python
"""Benchmark of the SIMP inner loop on B200 (BASELINE.json metric).

A step is one SIMP evaluation on the 3-D PointLoadCantilever 256x128x128 hex8 grid
(12 830 211 dofs): density filter -> penalisation -> matrix-free CG solve with the reference's
default settings (abstol 1e-7, reltol sqrt(eps), <= 700 iterations, zero initial guess) ->
compliance + sensitivity -> filter pullback.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

Prints one JSON line (rank 0).  `value` is device-resident (design vector already in HBM, gradient
left in HBM); `e2e` is the same step through the public API with pinned HOST buffers (design up,
objective + gradient down, every step).  `roofline` is the K.u kernel, timed live with CUDA events
on the library's stream.  `cpu_baseline` times the C port of the reference's CPU path
(oracle/topopt_ref.c) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

DEFAULT_NELS = (256, 128, 128)
VOLFRAC, PENAL, XMIN, RMIN = 0.3, 3.0, 1e-6, 2.0  # docs/tutorials/simp.qmd:63-65


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region: one streaming
    `nvidia-smi -lms 100` process started just before the region and stopped right after it."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index = index
        self.lines = []
        self._p = None
        self._t = None

    def _pump(self):
        try:
            for ln in self._p.stdout:
                self.lines.append(ln)
        except Exception:
            pass

    def __enter__(self):
        try:
            self._p = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            self._t = threading.Thread(target=self._pump, daemon=True)
            self._t.start()
            time.sleep(0.25)  # first sample lands before the timed region starts
        except Exception:
            self._p = None
        return self

    def __exit__(self, *a):
        if self._p is not None:
            time.sleep(0.12)
            self._p.terminate()
            try:
                self._p.wait(timeout=5)
            except Exception:
                self._p.kill()
            if self._t is not None:
                self._t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            s = [c.strip() for c in ln.split(",")]
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
                for n, v in zip(names, s[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(world):
    """dram__bytes_read.sum + dram__bytes_write.sum per K.u launch from the committed ncu capture
    (profiles/ncu_traffic.json); only recorded for the single-GPU configuration-4 kernel."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            d = json.load(f)
        return d.get(f"n{world}", {}).get("kxu_dram_bytes_per_launch")
    except Exception:
        return None


def workload_name(nels):
    return f"3D PointLoadCantilever {'x'.join(map(str, nels))} hex8 matrix-free CG, volfrac {VOLFRAC}, p={PENAL:g}, xmin={XMIN:g}, DensityFilter rmin={RMIN:g}"


# ------------------------------------------------------------------------------------------------
class CpuPort:
    """The C port of the reference's CPU path (oracle/topopt_ref.c) set up once for a grid."""

    def __init__(self, nels, openmp):
        import ref_c
        import topopt_jl_b200 as t

        self.nels = nels
        self.prob = t.PointLoadCantilever(nels)
        self.R = ref_c.RefProblem(3, 3, nels, self.prob.Ke, self.prob.prescribed_dofs, openmp=openmp, native=True)
        self.b = self.prob.fixedload.copy()
        self.b[self.prob.prescribed_dofs - 1] = 0.0

    def sample(self, cg_iters, maxiter):
        """Times a bounded sample: 1 filter pass, cg_iters CG iterations, 1 compliance/sensitivity,
        2 K.u applications; the SIMP step is extrapolated to `maxiter` CG iterations."""
        R, prob = self.R, self.prob
        rho = np.full(prob.nel, VOLFRAC)
        t0 = time.perf_counter()
        xf = R.filter(RMIN, rho)
        t_filter = time.perf_counter() - t0
        R.set_density(xf, PENAL, XMIN)
        t0 = time.perf_counter()
        u, it, _ = R.cg(self.b, abstol=0.0, reltol=0.0, maxiter=cg_iters)
        t_cg = (time.perf_counter() - t0) / max(it, 1)
        t0 = time.perf_counter()
        R.compliance(u, xf, PENAL, XMIN)
        t_sens = time.perf_counter() - t0
        t0 = time.perf_counter()
        for _ in range(2):
            R.mul(u)
        t_mul = (time.perf_counter() - t0) / 2
        step = maxiter * t_cg + t_sens + 2 * t_filter
        return {
            "step_s": step, "t_cg_iter_s": t_cg, "t_mul_s": t_mul, "t_sens_s": t_sens, "t_filter_s": t_filter, "threads": R.threads,
            "sample": f"{cg_iters} of {maxiter} CG iterations + 1 compliance/sensitivity + 1 filter pass at {'x'.join(map(str, self.nels))}, SIMP step extrapolated as {maxiter}*t_cg + t_sens + 2*t_filter",
            "kxu_gdofs": prob.ndof / t_mul / 1e9,
        }

    def close(self):
        self.R.close()


def cpu_port_sample(nels, cg_iters, openmp, maxiter):
    port = CpuPort(nels, openmp)
    try:
        return port.sample(cg_iters, maxiter)
    finally:
        port.close()


def run_reference(args, nels):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    info = None
    port = CpuPort(nels, openmp=True)
    for _ in range(min(args.warmup, 1)):
        port.sample(1, args.maxiter)
    for _ in range(args.steps):
        info = port.sample(args.ref_cg_iters, args.maxiter)
        vals.append(info["step_s"])
    port.close()
    step = float(np.mean(vals))
    v = 1.0 / step
    line = {
        "impl": "reference", "metric": "simp_iterations_per_sec", "value": v, "unit": "it/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(nels), "cg": {"abstol": 1e-7, "maxiter": args.maxiter}},
        "cpu_baseline": {"value": v, "unit": "it/s", "cores": info["threads"], "kind": "port", "sample": info["sample"],
                         "note": "C port of the TopOpt.jl CPU path (Julia is not installed in this image), OpenMP on all host threads"},
        "e2e": {"value": v, "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "kxu_gdofs": info["kxu_gdofs"],
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
def run_native(args, nels):
    import torch

    import topopt_jl_b200 as t
    from topopt_jl_b200 import distributed as D

    comm, local = D.init_from_env()
    rank = comm.rank if comm else 0
    world = comm.world if comm else 1
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libtopopt_cuda has no CPU fallback")
    torch.cuda.set_device(local)
    dist = torch.distributed if world > 1 else None

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    prob = t.PointLoadCantilever(nels)
    solver = t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(PENAL), xmin=XMIN, abstol=1e-7,
                         cg_max_iter=args.maxiter, device=local, comm=comm)
    filt = t.DensityFilterFun(solver, RMIN)
    x_host = torch.full((prob.nel,), VOLFRAC, dtype=torch.float64).pin_memory()
    g_host = torch.empty(prob.nel, dtype=torch.float64).pin_memory()
    x_dev = x_host.cuda()
    g_dev = torch.empty_like(x_dev)
    torch.cuda.synchronize()

    # ---- device-resident steps -----------------------------------------------------------------
    for _ in range(args.warmup):
        obj, res = t.simp_eval(solver, filt, x_dev, g_dev)
    solver.reset_stats()
    barrier()
    with ClockSampler(local) as clk:
        t0 = time.perf_counter()
        dev_ms = 0.0
        iters = 0
        for _ in range(args.steps):
            obj, res = t.simp_eval(solver, filt, x_dev, g_dev)
            dev_ms += res.solve_ms
            iters += res.iters
        barrier()
        wall = time.perf_counter() - t0
    st = solver.stats()
    launches = int(st.kernel_launches)
    tt = torch.tensor([wall], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    wall = float(tt.item())
    value = args.steps / wall

    # ---- end to end: pinned host buffers through the public API ---------------------------------
    for _ in range(min(args.warmup, 1)):
        t.simp_eval(solver, filt, x_host.numpy(), g_host.numpy())
    solver.reset_stats()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        obj_e2e, _ = t.simp_eval(solver, filt, x_host.numpy(), g_host.numpy())
    barrier()
    wall_e2e = time.perf_counter() - t0
    st2 = solver.stats()
    tt = torch.tensor([wall_e2e], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    wall_e2e = float(tt.item())

    # ---- dominant kernel: K.u, CUDA events on the library's stream ------------------------------
    kxu_ms = solver.time_kernel(0, args.kernel_reps)
    cg_ms = solver.time_kernel(1, args.kernel_reps)
    sens_ms = solver.time_kernel(2, 5)
    filt_ms = solver.time_kernel(3, 5, filt)
    tk = torch.tensor([kxu_ms, cg_ms], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(tk, op=dist.ReduceOp.MAX)
    kxu_ms, cg_ms = float(tk[0].item()), float(tk[1].item())
    peak, peak_src = measured_peaks()
    nown = nels[2] // world + (1 if world == 1 else 0)
    kxu_name = ("k_apply_hex8_modal2<12> (matrix-free hex8 K.u, two node rows per thread)" if nown >= 96
                else "k_apply_hex8_modal<16> (matrix-free hex8 K.u)")
    kxu_bytes_total = 16 * prob.ndof + 8 * prob.nel  # read x, write y, read E_e
    kxu_bytes_launch = kxu_bytes_total / world       # one launch per rank over its slab
    achieved = kxu_bytes_launch / (kxu_ms * 1e-3) / 1e9
    cg_bytes = (80 * prob.ndof + 8 * prob.nel) / world

    if rank == 0:
        line = {
            "metric": "simp_iterations_per_sec", "value": value, "unit": "it/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": workload_name(nels), "ndof": prob.ndof, "nel": prob.nel,
                "cg": {"abstol": 1e-7, "reltol": "sqrt(eps)", "maxiter": args.maxiter, "iters_per_step": iters / args.steps,
                       "converged": bool(res.converged), "residual": res.residual},
                "parallelism": f"z-slab x{world}" if world > 1 else "single GPU",
                "l2": "inputs larger than L2 (K.u touches 239 MB, one CG iteration 1.06 GB per pass)",
                "objective": obj,
            },
            "clocks": clk.summary(),
            "e2e": {"value": args.steps / wall_e2e, "unit": "it/s", "h2d_bytes_per_step": int(st2.h2d_bytes // args.steps),
                    "d2h_bytes_per_step": int(st2.d2h_bytes // args.steps) + 8, "objective": obj_e2e},
            "gpu_launches": launches,
            "roofline": {"kernel": kxu_name, "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(world) if nels == DEFAULT_NELS else None,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": kxu_bytes_launch, "ms_per_launch": kxu_ms},
            "kxu_gdofs": prob.ndof / (kxu_ms * 1e-3) / 1e9,
            "cg_iteration": {"ms": cg_ms, "it_per_s": 1e3 / cg_ms, "achieved_gbs": cg_bytes / (cg_ms * 1e-3) / 1e9,
                             "frac_of_hbm": cg_bytes / (cg_ms * 1e-3) / 1e9 / peak},
            "kernels_ms": {"kxu": kxu_ms, "cg_iteration": cg_ms, "sensitivity": sens_ms, "filter_forward": filt_ms},
            "solve_ms_per_step": dev_ms / args.steps,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                cb = cpu_port_sample(nels, args.cpu_cg_iters, False, args.maxiter)
                line["cpu_baseline"] = {"value": 1.0 / cb["step_s"], "unit": "it/s", "cores": cb["threads"], "kind": "port", "sample": cb["sample"],
                                        "kxu_gdofs": cb["kxu_gdofs"], "t_cg_iter_s": cb["t_cg_iter_s"],
                                        "note": "single-threaded C port of the TopOpt.jl CPU path (the reference hot path has no threading; Julia is not installed)"}
            except Exception as e:  # the baseline must never take the GPU number down with it
                line["cpu_baseline"] = {"value": None, "unit": "it/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
        emit(line)
    filt.close()
    solver.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict):
    """The one JSON line goes to the real stdout; everything else (NCCL banners, warnings) was
    rerouted to stderr so that stdout carries exactly one line."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--nels", type=str, default=",".join(map(str, DEFAULT_NELS)))
    ap.add_argument("--maxiter", type=int, default=700)
    ap.add_argument("--kernel-reps", type=int, default=50)
    ap.add_argument("--cpu-cg-iters", type=int, default=8)
    ap.add_argument("--ref-cg-iters", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    nels = tuple(int(v) for v in args.nels.split(","))
    if args.impl == "reference":
        run_reference(args, nels)
    else:
        run_native(args, nels)


if __name__ == "__main__":
    main()
