"""Build recipe for libtopopt_cuda.so (in-tree, sm_100a only)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libtopopt_cuda.so")
SOURCES = [os.path.join(HERE, "csrc", "api.cu"), os.path.join(HERE, "csrc", "host_mesh.cpp")]
DEPS = SOURCES + [
    os.path.join(HERE, "csrc", "kernels.cuh"),
    os.path.join(HERE, "csrc", "kxu_hex8.cuh"),
    os.path.join(HERE, "csrc", "kxu_hex8_2row.cuh"),
    os.path.join(HERE, "csrc", "kxu_hex8_ring.cuh"),
    os.path.join(HERE, "csrc", "kxu_hex8_cgfused.cuh"),
    os.path.join(HERE, "csrc", "kxu_hex8_cgtma.cuh"),
    os.path.join(HERE, "csrc", "mg_solve.inl"),
    os.path.join(HERE, "csrc", "multigrid.cuh"),
    os.path.join(HERE, "csrc", "common.h"),
    os.path.join(ROOT, "include", "topopt_cuda.h"),
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> topopt.jl_b200/libtopopt_cuda.so"""
    if not force and not is_stale():
        return LIB
    cmd = [
        nvcc_path(),
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-lineinfo", "-O3", "-std=c++17",
        "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
        "-I", os.path.join(ROOT, "include"),
        "-o", LIB,
    ] + SOURCES + ["-ldl"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
