"""N>1 host logic on CPU: world_size 2 over gloo (no GPU needed)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_ranges(lib):
    from topopt_jl_b200.distributed import slab_ranges

    for nl in (2, 7, 128):
        for w in (1, 2, 4, 8):
            if nl < w:
                continue
            r = slab_ranges(nl, w)
            layers = [e for (e0, e1), _ in r for e in range(e0, e1)]
            planes = [k for _, (k0, k1) in r for k in range(k0, k1)]
            assert layers == list(range(nl)) and planes == list(range(nl + 1))
            assert all(e1 > e0 for (e0, e1), _ in r)
    assert slab_ranges(128, 8)[3] == ((48, 64), (48, 64)) and slab_ranges(128, 8)[7] == ((112, 128), (112, 129))


def test_two_ranks_gloo():
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "gloo_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and "GLOO_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
