#!/usr/bin/env python
"""Benchmark of the SIMP inner loop on B200 (BASELINE.json metric).

A step is one SIMP evaluation on the 3-D PointLoadCantilever 256x128x128 hex8 grid
(12 830 211 dofs): density filter -> penalisation -> matrix-free CG solve with the reference's
default settings (abstol 1e-7, reltol sqrt(eps), <= 700 iterations, zero initial guess) ->
compliance + sensitivity -> filter pullback.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--config 1|2|3|4|5]

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


def ncu_traffic(world, key="kxu_dram_bytes_per_launch"):
    """dram__bytes_read.sum + dram__bytes_write.sum per K.u launch from the committed ncu capture
    (profiles/ncu_traffic.json); only recorded for the single-GPU configuration-4 kernel."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            d = json.load(f)
        return d.get(f"n{world}", {}).get(key)
    except Exception:
        return None


def workload_name(nels):
    return f"3D PointLoadCantilever {'x'.join(map(str, nels))} hex8 matrix-free CG, volfrac {VOLFRAC}, p={PENAL:g}, xmin={XMIN:g}, DensityFilter rmin={RMIN:g}"


# ------------------------------------------------------------------------------------------------
def host_threads():
    """All host cores, regardless of the OMP_NUM_THREADS=1 that torchrun exports to its workers."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class CpuPort:
    """The C port of the reference's CPU path (oracle/topopt_ref.c) set up once for a grid.  Its inputs come
    from the oracle alone (oracle/ref_c.point_load_cantilever_inputs): this arm never touches the product."""

    def __init__(self, nels, openmp):
        if openmp:
            os.environ["OMP_NUM_THREADS"] = str(host_threads())  # before the OpenMP runtime starts
        import ref_c

        self.nels = nels
        Ke, pres, f = ref_c.point_load_cantilever_inputs(nels)
        self.ndof, self.nel = f.shape[0], int(np.prod(nels))
        self.R = ref_c.RefProblem(3, 3, nels, Ke, pres, openmp=openmp, native=True)
        self.b = f.copy()
        self.b[pres - 1] = 0.0

    def sample(self, cg_iters, maxiter):
        """Times a bounded sample: 1 filter pass, cg_iters CG iterations, 1 compliance/sensitivity,
        2 K.u applications; the SIMP step is extrapolated to `maxiter` CG iterations."""
        R = self.R
        rho = np.full(self.nel, VOLFRAC)
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
            "kxu_gdofs": self.ndof / t_mul / 1e9,
        }

    def full_step(self, maxiter):
        """One COMPLETE SIMP evaluation (nothing extrapolated): filter, penalise, CG with the reference's
        defaults, compliance + sensitivity, filter pullback (the port's filter is symmetric in cost)."""
        R = self.R
        rho = np.full(self.nel, VOLFRAC)
        t0 = time.perf_counter()
        xf = R.filter(RMIN, rho)
        R.set_density(xf, PENAL, XMIN)
        u, it, res = R.cg(self.b, abstol=1e-7, maxiter=maxiter)
        obj, _, grad = R.compliance(u, xf, PENAL, XMIN)
        R.filter(RMIN, grad)
        return {"step_s": time.perf_counter() - t0, "cg_iters": it, "residual": res, "objective": obj, "threads": R.threads}

    def close(self):
        self.R.close()


def cpu_port_sample(nels, cg_iters, openmp, maxiter):
    port = CpuPort(nels, openmp)
    try:
        return port.sample(cg_iters, maxiter)
    finally:
        port.close()


CFG3_NELS = (60, 20, 20)


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
    # beside the extrapolated figure: config 3 (60x20x20) run in full, nothing extrapolated
    p3 = CpuPort(CFG3_NELS, openmp=True)
    p3.full_step(args.maxiter)
    full3 = p3.full_step(args.maxiter)
    p3.close()
    line = {
        "impl": "reference", "metric": "simp_iterations_per_sec", "value": v, "unit": "it/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(nels), "cg": {"abstol": 1e-7, "maxiter": args.maxiter}},
        "cpu_baseline": {"value": v, "unit": "it/s", "cores": info["threads"], "kind": "port", "sample": info["sample"],
                         "note": "C port of the TopOpt.jl CPU path (Julia is not installed in this image), OpenMP on all host threads; "
                                 "inputs built from oracle/ only"},
        "e2e": {"value": v, "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "kxu_gdofs": info["kxu_gdofs"],
        "host_threads": info["threads"],
        "config3_full_step": {"workload": workload_name(CFG3_NELS), "it_per_s": 1.0 / full3["step_s"], "ms_per_step": full3["step_s"] * 1e3,
                              "cg_iters": full3["cg_iters"], "objective": full3["objective"], "threads": full3["threads"],
                              "note": "complete SIMP evaluation, not extrapolated"},
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

    # ---- parity self-check against the oracle, before anything is timed (exits non-zero on failure) ----
    parity = parity_check(t, comm, local, world)
    if rank == 0 and not parity["ok"]:
        emit({"metric": "simp_iterations_per_sec", "value": None, "n_gpus": world, "parity_check": parity, "error": "parity check failed"})
    if not parity["ok"]:
        raise SystemExit(3)

    prob = t.PointLoadCantilever(nels)
    solver = t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(PENAL), xmin=XMIN, abstol=1e-7,
                         cg_max_iter=args.maxiter, device=local, comm=comm, cg_variant=args.cg_variant)
    filt = t.DensityFilterFun(solver, RMIN)
    x_host = torch.full((prob.nel,), VOLFRAC, dtype=torch.float64).pin_memory()
    g_host = torch.empty(prob.nel, dtype=torch.float64).pin_memory()
    x_dev = x_host.cuda()
    g_dev = torch.empty_like(x_dev)
    torch.cuda.synchronize()

    def max_over_ranks(v):
        tt = torch.tensor([v], dtype=torch.float64, device="cuda")
        if dist:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def timed_steps(xin, gout, nsteps):
        """nsteps SIMP evaluations bracketed by barrier + synchronize; wall = max over ranks."""
        solver.reset_stats()
        barrier()
        t0 = time.perf_counter()
        dev_ms, iters, obj, res = 0.0, 0, None, None
        for _ in range(nsteps):
            obj, res = t.simp_eval(solver, filt, xin, gout)
            dev_ms += res.solve_ms
            iters += res.iters
        barrier()
        wall = max_over_ranks(time.perf_counter() - t0)
        return wall, dev_ms, iters, obj, res, solver.stats()

    # ---- device-resident steps -----------------------------------------------------------------
    for _ in range(args.warmup):
        t.simp_eval(solver, filt, x_dev, g_dev)
    with ClockSampler(local) as clk:
        wall, dev_ms, iters, obj, res, st = timed_steps(x_dev, g_dev, args.steps)
    launches = int(st.kernel_launches)
    value = args.steps / wall

    # ---- end to end: pinned host buffers through the public API ---------------------------------
    for _ in range(min(args.warmup, 1)):
        t.simp_eval(solver, filt, x_host.numpy(), g_host.numpy())
    wall_e2e, _, _, obj_e2e, _, st2 = timed_steps(x_host.numpy(), g_host.numpy(), args.steps)

    # ---- the reference's own scalar recurrence (three vector passes), for comparison -----------------
    ref_rec = None
    if args.cg_variant != 0:
        solver.cg_variant = 0
        t.simp_eval(solver, filt, x_dev, g_dev)
        w0, d0, i0, o0, r0, _ = timed_steps(x_dev, g_dev, 1)
        ref_rec = {"value": 1.0 / w0, "unit": "it/s", "ms_per_step": w0 * 1e3, "iters": i0, "objective": o0, "residual": r0.residual,
                   "objective_rel_diff_vs_headline": abs(o0 - obj) / abs(o0)}
        solver.cg_variant = args.cg_variant

    # ---- SURVEY 8d converged run: abstol 1e-10, iteration cap 20000, zero initial guess -------------------
    conv = None
    if not args.no_converged_run:
        solver.abstol, keep_max = 1e-10, solver.cg_max_iter
        solver.cg_max_iter = 20000
        barrier()
        t0 = time.perf_counter()
        oc, rc = t.simp_eval(solver, filt, x_dev, g_dev)
        barrier()
        wc = max_over_ranks(time.perf_counter() - t0)
        conv = {"abstol": 1e-10, "maxiter": 20000, "iters": rc.iters, "converged": bool(rc.converged), "residual": rc.residual,
                "solve_ms": rc.solve_ms, "step_s": wc, "objective": oc, "cg_it_per_s": rc.iters / (rc.solve_ms * 1e-3)}
        solver.abstol, solver.cg_max_iter = 1e-7, keep_max

    # ---- opt-in extension: the same converged step with the geometric-multigrid preconditioner --------
    mg = None
    if not args.no_converged_run and world == 1:
        solver.abstol, keep_max, solver.preconditioner = 1e-10, solver.cg_max_iter, "multigrid"
        solver.cg_max_iter = 2000
        try:
            t.simp_eval(solver, filt, x_dev, g_dev)  # builds the hierarchy
            barrier()
            t0 = time.perf_counter()
            om, rm = t.simp_eval(solver, filt, x_dev, g_dev)  # stiffness re-uploaded: coarse operators and bounds are rebuilt
            barrier()
            wm = time.perf_counter() - t0
            mg = {"abstol": 1e-10, "iters": rm.iters, "converged": bool(rm.converged), "residual": rm.residual, "solve_ms": rm.solve_ms,
                  "step_s": wm, "it_per_s": 1.0 / wm, "objective": om,
                  "objective_rel_diff_vs_converged_run": (abs(om - conv["objective"]) / abs(conv["objective"])) if conv else None,
                  "note": "V-cycle (Chebyshev degree 2 smoother, rediscretised coarse operators, dense coarsest solve) preconditioned CG; "
                          "step_s includes the rebuild of the hierarchy for the re-uploaded stiffness"}
        except Exception as e:
            mg = {"failed": str(e)}
        solver.abstol, solver.cg_max_iter, solver.preconditioner = 1e-7, keep_max, None

    # ---- dominant kernel: K.u as the CG loop launches it, CUDA events on the library's stream -------------
    single = args.cg_variant == 1 and world == 1 or (args.cg_variant == 1 and comm is not None and getattr(comm, "peer_memory", True))
    nown = nels[2] // world + (1 if world == 1 else 0)
    ring = nown >= int(os.environ.get("TOPOPT_KXU_RING_MIN", "12")) and os.environ.get("TOPOPT_KXU_RING", "1") != "0"
    kxu_which = 8 if (single and ring) else 0
    kxu_ms = solver.time_kernel(kxu_which, args.kernel_reps)
    cg_ms = solver.time_kernel(9 if single and ring else 1, max(args.kernel_reps, 200))
    # single-pass recurrence on hex8 slabs: the whole CG iteration is ONE kernel (kxu_hex8_cgfused.cuh), on one GPU and -- over
    # peer memory -- across ranks; it is the dominant kernel of the step
    fused = (single and ring and os.environ.get("TOPOPT_CG_FUSED", "1") != "0"
             and (world == 1 or os.environ.get("TOPOPT_CG_FUSED_MGPU", "1") != "0"))
    cg_ref_ms = solver.time_kernel(1, args.kernel_reps)
    kxu_prev_ms = solver.time_kernel(7, args.kernel_reps)
    sens_ms = solver.time_kernel(2, 5)
    filt_ms = solver.time_kernel(3, 5, filt)
    kxu_ms, cg_ms, cg_ref_ms = max_over_ranks(kxu_ms), max_over_ranks(cg_ms), max_over_ranks(cg_ref_ms)
    peak, peak_src = measured_peaks()
    if ring:
        kxu_name = "k_apply_hex8_ring (matrix-free hex8 K.u: bulk-copy producer warp + compute warps, fused p.Ap" + (" and Ap.Ap)" if kxu_which == 8 else ")")
    else:
        kxu_name = ("k_apply_hex8_modal2<12> (matrix-free hex8 K.u, two node rows per thread)" if nown >= 96
                    else "k_apply_hex8_modal<16> (matrix-free hex8 K.u)")
    kxu_bytes_total = 16 * prob.ndof + 8 * prob.nel  # read x, write y, read E_e
    kxu_bytes_launch = kxu_bytes_total / world       # one launch per rank over its slab
    achieved = kxu_bytes_launch / (kxu_ms * 1e-3) / 1e9
    cg_bytes = ((64 if fused else (72 if single and ring else 88)) * prob.ndof + 8 * prob.nel) / world  # bytes one iteration moves
    kxu_alone = {"kernel": kxu_name, "algorithmic_bytes_per_launch": kxu_bytes_launch, "ms_per_launch": kxu_ms, "achieved": achieved,
                 "frac": achieved / peak, "traffic": ncu_traffic(world) if nels == DEFAULT_NELS else None,
                 "previous_generation_ms": kxu_prev_ms,
                 "timed_on": "dense direction vector (zero on prescribed dofs), fused dot products; launched once per solve when the iteration is one kernel"}
    if fused:
        # read p, r, Ap, x and write p, r, Ap, x (64 B per dof) + E_e (8 B per element): DESIGN.md section 4
        roof_name = ("k_cg_fused_hex8 (one CG iteration per launch: x += a p, r -= a Ap, p = r + b p applied while the planes are staged, "
                     "matrix-free hex8 K.p, fused p.Ap / Ap.Ap / r.r and scalar step)")
        roof_bytes = (64 * prob.ndof + 8 * prob.nel) / world  # one launch per rank over its slab
        roof_ms = cg_ms
        roof_traffic = ncu_traffic(world, "cg_fused_dram_bytes_per_launch") if nels == DEFAULT_NELS else None
        roof_timed = ("inside the CG loop on a dense right-hand side (topopt_time_kernel class 9: solve time / iterations, graph-replayed launches"
                      + ("; includes the in-kernel all-reduce wait for the slowest rank)" if world > 1 else ")"))
    else:
        roof_name, roof_bytes, roof_ms, roof_traffic = kxu_name, kxu_bytes_launch, kxu_ms, kxu_alone["traffic"]
        roof_timed = "dense direction vector (zero on prescribed dofs), fused dot products, the variant the CG loop launches"
    roof_achieved = roof_bytes / (roof_ms * 1e-3) / 1e9

    # ---- BASELINE config 3 (60x20x20) in full, the un-extrapolated twin of the reference arm's figure ----------
    cfg3 = None
    if world == 1:
        p3 = t.PointLoadCantilever(CFG3_NELS)
        s3 = t.FEASolver(t.CUDAMatrixFreeSolver, p3, penalty=t.PowerPenaltyFun(PENAL), xmin=XMIN, abstol=1e-7, cg_max_iter=args.maxiter,
                         device=local, cg_variant=0)
        f3 = t.DensityFilterFun(s3, RMIN)
        x3, g3 = np.full(p3.nel, VOLFRAC), np.empty(p3.nel)
        for _ in range(3):
            t.simp_eval(s3, f3, x3, g3)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            o3, r3 = t.simp_eval(s3, f3, x3, g3)
        torch.cuda.synchronize()
        w3 = (time.perf_counter() - t0) / 5
        cfg3 = {"workload": workload_name(CFG3_NELS), "it_per_s": 1.0 / w3, "ms_per_step": w3 * 1e3, "cg_iters": r3.iters, "objective": o3,
                "note": "complete SIMP evaluation through the public API with host buffers, not extrapolated"}
        f3.close()
        s3.close()

    if rank == 0:
        line = {
            "metric": "simp_iterations_per_sec", "value": value, "unit": "it/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": workload_name(nels), "ndof": prob.ndof, "nel": prob.nel,
                "cg": {"abstol": 1e-7, "reltol": "sqrt(eps)", "maxiter": args.maxiter, "iters_per_step": iters / args.steps,
                       "converged": bool(res.converged), "residual": res.residual,
                       "recurrence": ("single-pass CG (beta predicted from alpha^2 Ap.Ap - r.r; x, r, p updated in one pass; iterates agree with "
                                      "IterativeSolvers' cg! to rounding, see reference_recurrence)" if single and ring else "IterativeSolvers cg! recurrence"),
                       "kernels_per_iteration": 1 if fused else (2 if single and ring else 3)},
                "parallelism": f"z-slab x{world}" if world > 1 else "single GPU",
                "l2": "inputs larger than L2 (K.u touches 239 MB, one CG iteration 0.86-1.06 GB)",
                "objective": obj,
            },
            "clocks": clk.summary(),
            "e2e": {"value": args.steps / wall_e2e, "unit": "it/s", "h2d_bytes_per_step": int(st2.h2d_bytes // args.steps),
                    "d2h_bytes_per_step": int(st2.d2h_bytes // args.steps) + 8, "objective": obj_e2e},
            "gpu_launches": launches,
            "parity_check": parity,
            "roofline": {"kernel": roof_name, "bound": "hbm", "achieved": roof_achieved, "peak": peak,
                         "unit": "GB/s", "frac": roof_achieved / peak, "traffic": roof_traffic,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": roof_bytes, "ms_per_launch": roof_ms,
                         "timed_on": roof_timed},
            "kxu_alone": kxu_alone,
            "kxu_gdofs": prob.ndof / (kxu_ms * 1e-3) / 1e9,
            "cg_iteration": {"ms": cg_ms, "it_per_s": 1e3 / cg_ms, "achieved_gbs": cg_bytes / (cg_ms * 1e-3) / 1e9,
                             "frac_of_hbm": cg_bytes / (cg_ms * 1e-3) / 1e9 / peak,
                             "reference_recurrence_ms": cg_ref_ms,
                             "bytes_moved_per_iteration": cg_bytes},
            "kernels_ms": {"kxu": kxu_ms, "cg_iteration": cg_ms, "sensitivity": sens_ms, "filter_forward": filt_ms},
            "solve_ms_per_step": dev_ms / args.steps,
            "reference_recurrence": ref_rec,
            "converged_run": conv,
            "multigrid_run": mg,
            "config3_full_step": cfg3,
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


def run_small_config(args):
    """BASELINE configs 1, 2, 3, 5 (single GPU; latency / L2-resident cases): one JSON line each with the same keys
    as the headline.  A step is one objective + gradient evaluation through the public API with host buffers."""
    import torch

    import topopt_jl_b200 as t

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libtopopt_cuda has no CPU fallback")
    c = args.config
    if c == 1:
        name, prob, S, xmin, vf = "2D PointLoadCantilever 160x40 quad4 matrix-free CG, DensityFilter rmin=2, p=3", t.PointLoadCantilever((160, 40)), t.CUDAMatrixFreeSolver, 1e-3, 0.5
    elif c == 2:
        name, prob, S, xmin, vf = "2D HalfMBB 600x200 quad4 assembled CSR + SpMV CG, DensityFilter rmin=2, p=3", t.HalfMBB((600, 200)), t.CUDAAssemblySolver, 1e-3, 0.5
    elif c == 3:
        name, prob, S, xmin, vf = workload_name(CFG3_NELS), t.PointLoadCantilever(CFG3_NELS), t.CUDAMatrixFreeSolver, XMIN, VOLFRAC
    else:
        name, prob, S, xmin, vf = "2D HeatConductionProblem 1024x1024 quad4 thermal compliance, SensFilter rmin=2, p=3", t.HeatTree((1024, 1024)), t.CUDAMatrixFreeSolver, 1e-3, 0.4
    s = t.FEASolver(S, prob, penalty=t.PowerPenaltyFun(PENAL), xmin=xmin, abstol=1e-7, cg_max_iter=args.maxiter, cg_variant=args.cg_variant)
    x, g = np.full(prob.nel, vf), np.empty(prob.nel)
    if c == 5:
        F = t.SensFilterFun(s, RMIN)
        tc = t.ThermalComplianceFun(s)

        def step():
            J, gr = tc.value_and_grad(x)
            return J, F.pullback(gr), s.last_result
    else:
        F = t.DensityFilterFun(s, RMIN)

        def step():
            o_, r_ = t.simp_eval(s, F, x, g)
            return o_, g, r_
    for _ in range(args.warmup):
        step()
    s.reset_stats()
    torch.cuda.synchronize()
    with ClockSampler(0) as clk:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            obj, _, res = step()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
    st = s.stats()
    md = prob.metadata
    peak, peak_src = measured_peaks()
    assembled = c == 2
    kxu_ms = s.time_kernel(4 if assembled else 0, args.kernel_reps, F)
    one_kernel = c == 3 and args.cg_variant == 1 and os.environ.get("TOPOPT_CG_FUSED", "1") != "0"  # hex8, >= 12 node planes
    cg_ms = s.time_kernel(6 if assembled else (9 if one_kernel else 1), max(args.kernel_reps, 200) if one_kernel else args.kernel_reps, F)
    kbytes = (12 * md.nnz + 20 * md.ndof) if assembled else (16 * md.ndof + 8 * md.nel)
    line = {
        "metric": "simp_iterations_per_sec", "value": args.steps / wall, "unit": "it/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "baseline_config": c, "ndof": md.ndof, "nel": md.nel,
                   "cg": {"abstol": 1e-7, "maxiter": args.maxiter, "iters_last_solve": res.iters, "converged": bool(res.converged)},
                   "l2": "working set fits in L2: latency / launch-bound case, timed back to back (no flush)", "objective": obj},
        "clocks": clk.summary(),
        "e2e": {"value": args.steps / wall, "unit": "it/s", "h2d_bytes_per_step": int(st.h2d_bytes // args.steps), "d2h_bytes_per_step": int(st.d2h_bytes // args.steps) + 8},
        "gpu_launches": int(st.kernel_launches),
        "roofline": {"kernel": "k_spmv (assembled CSR)" if assembled else "k_apply (matrix-free K.u)", "bound": "hbm", "achieved": kbytes / (kxu_ms * 1e-3) / 1e9,
                     "peak": peak, "unit": "GB/s", "frac": kbytes / (kxu_ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": kbytes, "ms_per_launch": kxu_ms},
        "cg_iteration": {"ms": cg_ms, "it_per_s": 1e3 / cg_ms, "kernels_per_iteration": 1 if one_kernel else 3},
        "cpu_baseline": {"value": None, "unit": "it/s", "cores": 0, "kind": "port", "sample": "not timed for the secondary configs (see --config 4 and config3_full_step)"},
    }
    emit(line)
    F.close()
    s.close()


def parity_check(t, comm, local, world):
    """Small slab grids through the public API on THIS run's rank layout against the CPU oracle: K.u, compliance
    and the filtered gradient (1e-12 / 1e-8 / 1e-8).  Grid 1 has 3 node planes per rank (thin-slab kernels),
    grid 2 has >= 100 per rank (the kernels selected for thick slabs, peer-memory halo reads included)."""
    import topopt_oracle as o

    out = {"ok": True, "grids": []}
    for nels in ((12, 6, 3 * world + (1 if world == 1 else 0)), (8, 4, 100 * world)):
        nels = (nels[0], nels[1], nels[2] + (nels[2] % 2))
        prob, oprob = t.PointLoadCantilever(nels), o.PointLoadCantilever(nels)
        prob.Ke = oprob.Ke.copy()
        s = t.FEASolver(t.CUDAMatrixFreeSolver, prob, penalty=t.PowerPenaltyFun(PENAL), xmin=1e-3, abstol=1e-11, reltol=1e-14,
                        cg_max_iter=50000, device=local, comm=comm, cg_variant=1)
        F = t.DensityFilterFun(s, RMIN)
        x = np.random.default_rng(5).uniform(0.2, 1.0, prob.nel)
        g = np.empty(prob.nel)
        obj, res = t.simp_eval(s, F, x, g)
        Fo = o.DensityFilter(oprob, RMIN)
        xf = Fo(x)
        E = o.get_rho(xf, PENAL, 1e-3)
        u = o.solve_direct(oprob, E)
        oo, _, go = o.compliance(oprob, u, xf, PENAL, 1e-3)
        go = Fo.pullback(go)
        v = np.random.default_rng(6).standard_normal(prob.ndof)
        v[oprob.prescribed] = 0.0
        y = s.mul(v)
        yref = o.matfree_mul(oprob, E, v)
        r = {"nels": list(nels), "obj_rel": abs(obj - oo) / abs(oo), "grad_rel": float(np.max(np.abs(g - go)) / np.max(np.abs(go))),
             "mul_rel": float(np.max(np.abs(y - yref)) / np.max(np.abs(yref))), "cg_iters": res.iters, "converged": bool(res.converged)}
        r["ok"] = bool(r["obj_rel"] < 1e-8 and r["grad_rel"] < 1e-8 and r["mul_rel"] < 1e-12 and res.converged)
        out["ok"] = out["ok"] and r["ok"]
        out["grids"].append(r)
        F.close()
        s.close()
    return out


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
    ap.add_argument("--no-converged-run", action="store_true")
    ap.add_argument("--cg-variant", type=int, default=1, help="0 = IterativeSolvers recurrence, 1 = single-pass recurrence")
    ap.add_argument("--config", type=int, default=4, choices=[1, 2, 3, 4, 5], help="BASELINE.json config (4 = the headline)")
    args = ap.parse_args()
    nels = tuple(int(v) for v in args.nels.split(","))
    if args.impl == "reference":
        run_reference(args, nels)
    elif args.config != 4:
        run_small_config(args)
    else:
        run_native(args, nels)


if __name__ == "__main__":
    main()
