"""bench.py contract on CPU: the reference arm (`--impl reference`) prints exactly one JSON line on
stdout with the keys the driver reads (tiny grid so it runs in seconds without a GPU)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--nels", "12,4,4", "--steps", "2", "--warmup", "1",
           "--ref-cg-iters", "3"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "simp_iterations_per_sec" and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["vs_baseline"] is None and "workload" in d["config"]


def test_native_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback: on a machine without a CUDA device the native arm must fail loudly instead of
    silently timing something else."""
    import torch

    if torch.cuda.is_available():
        import pytest

        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--nels", "8,4,4", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode != 0
    assert out.stdout.strip() == ""  # no JSON line is fabricated
    assert "CUDA" in out.stderr or "cuda" in out.stderr


def test_reference_arm_is_independent_of_the_product_and_of_torchrun_thread_limits():
    """the reference arm builds its inputs from oracle/ only (the product package and its
    CUDA library are never imported) and uses every host core even when the launcher exports
    OMP_NUM_THREADS=1 (torchrun does); it also reports one COMPLETE config-3 step next to the extrapolated one."""
    code = (
        "import sys, json, io, os\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "sys.argv = ['bench.py', '--impl', 'reference', '--nels', '12,4,4', '--steps', '1', '--warmup', '1', '--ref-cg-iters', '2']\n"
        "import bench\n"
        "bench.main()\n"
        "bad = [m for m in sys.modules if m.startswith('topopt_jl_b200') or m == 'torch']\n"
        "maps = open('/proc/self/maps').read()\n"
        "sys.stderr.write('LOADED=' + json.dumps(bad) + ' CUDA_SO=' + str('libtopopt_cuda' in maps) + '\\n')\n"
    )
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "LOADED=[] CUDA_SO=False" in out.stderr, out.stderr[-500:]
    d = json.loads([ln for ln in out.stdout.splitlines() if ln.strip()][0])
    ncpu = len(os.sched_getaffinity(0))
    assert d["cpu_baseline"]["cores"] == ncpu == d["host_threads"]
    full = d["config3_full_step"]
    assert full["cg_iters"] > 0 and full["it_per_s"] > 0 and "60x20x20" in full["workload"]
