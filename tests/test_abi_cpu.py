"""CPU-side checks of the C ABI: the shared library loads without a GPU, exports every symbol
include/topopt_cuda.h declares, its host entry points (numbering, connectivity, CSC pattern,
element matrices) agree BIT-EXACTLY with the oracle, and it refuses to compute without a device.
Also validates the oracle's C port (the cpu_baseline) against the numpy oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import topopt_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

GRIDS = [((2, 2), 2), ((7, 4), 2), ((160, 40), 2), ((5, 3), 1), ((3, 2, 2), 3), ((6, 4, 5), 3), ((4, 3, 2), 1), ((1, 1), 2), ((1, 1, 1), 3)]


def test_library_exports_every_declared_symbol(lib):
    from topopt_jl_b200 import _lib

    L = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "topopt_cuda.h")).read()
    declared = set(re.findall(r"\b(topopt_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 28
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/topopt_cuda.h but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes binding and header disagree"
    assert b"sm_100a" in L.topopt_version()


@pytest.mark.parametrize("nels,ncomp", GRIDS)
def test_numbering_bit_exact(lib, nels, ncomp):
    t = lib
    md = t.Metadata(len(nels), ncomp, nels)
    g = o.Grid(nels)
    omd = o.Metadata(g, ncomp)
    assert (md.nnodes, md.nel, md.ndof) == (g.nnodes, g.nel, omd.ndof)
    assert np.array_equal(md.cells, g.cells + 1)
    assert np.array_equal(md.node_dofs, omd.node_dofs + 1)
    assert np.array_equal(md.cell_dofs, omd.cell_dofs + 1)


@pytest.mark.parametrize("nels,ncomp", [g for g in GRIDS if np.prod(g[0]) < 2000])
def test_csc_pattern_bit_exact(lib, nels, ncomp):
    t = lib
    md = t.Metadata(len(nels), ncomp, nels)
    colptr, rowval = md.csc_pattern()
    g = o.Grid(nels)
    prob = o.Problem(g, o.Metadata(g, ncomp), np.eye(ncomp * 2 ** len(nels)), [], np.zeros(ncomp * g.nnodes), "x")
    cp, rv = o.csc_pattern(prob)
    assert md.nnz == rv.shape[0]
    assert np.array_equal(colptr - 1, cp) and np.array_equal(rowval - 1, rv)


def test_library_matches_reference_goldens_directly(lib):
    """The library's host entry points against the reference's golden integers themselves
    (tests/golden/reference_goldens.json), not only through the oracle."""
    import json

    t = lib
    G = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_goldens.json")))
    md = t.Metadata(2, 2, (2, 2))
    assert md.cells.tolist() == G["halfmbb_2x2"]["cells"]
    assert md.node_dofs.tolist() == G["halfmbb_2x2"]["node_dofs"]
    assert md.cell_dofs.tolist() == G["halfmbb_2x2"]["cell_dofs"]
    assert t.PointLoadCantilever((160, 40)).force_dof == G["force_dofs"]["PointLoadCantilever_160x40"]
    assert t.HalfMBB((60, 20)).force_dof == G["force_dofs"]["HalfMBB_60x20"]
    assert len(t.HeatTree(tuple(G["heat_tree"]["nels"])).prescribed_dofs) == G["heat_tree"]["n_prescribed"]


def test_config_sizes(lib):
    """SURVEY 8 config table (ndof, nnz) from the closed forms in the library."""
    t = lib
    assert (t.Metadata(2, 2, (600, 200)).ndof, t.Metadata(2, 2, (600, 200)).nnz) == (241602, 4329604)
    assert (t.Metadata(3, 3, (60, 20, 20)).ndof, t.Metadata(3, 3, (60, 20, 20)).nnz) == (80703, 6061509)
    m4 = t.Metadata(3, 3, (256, 128, 128))
    assert (m4.nel, m4.nnodes, m4.ndof, m4.nnz) == (4194304, 4276737, 12830211, 1025865225)
    assert t.Metadata(2, 1, (1024, 1024)).ndof == 1050625


def test_element_matrices_match_oracle(lib):
    t = lib
    from topopt_jl_b200 import _lib

    for dim, sizes in ((2, (1.0, 1.0)), (2, (0.5, 2.0)), (3, (1.0, 1.0, 1.0)), (3, (1.0, 0.5, 2.0))):
        K = t.element_matrix(dim, _lib.PHYSICS_ELASTICITY, sizes, 1.3, 0.3)
        assert np.max(np.abs(K - o.element_stiffness(dim, sizes, 1.3, 0.3))) < 1e-14
        assert np.array_equal(K, K.T)
        H = t.element_matrix(dim, _lib.PHYSICS_HEAT, sizes, 2.0)
        assert np.max(np.abs(H - o.element_conductivity(dim, sizes, 2.0))) < 1e-14


def test_problem_mirror_matches_oracle(lib):
    t = lib
    for mk, mko in [
        (lambda: t.PointLoadCantilever((160, 40)), lambda: o.PointLoadCantilever((160, 40))),
        (lambda: t.HalfMBB((60, 20)), lambda: o.HalfMBB((60, 20))),
        (lambda: t.PointLoadCantilever((6, 4, 2), (1.0, 0.5, 2.0)), lambda: o.PointLoadCantilever((6, 4, 2), (1.0, 0.5, 2.0))),
        (lambda: t.HeatTree((9, 7)), lambda: o.HeatTree((9, 7))),
    ]:
        p, q = mk(), mko()
        assert np.array_equal(p.prescribed_dofs - 1, q.prescribed)
        assert np.array_equal(p.fixedload, q.fixedload)
        assert np.allclose(p.cellvolumes, q.cellvolumes)
        if hasattr(q, "force_dof"):
            assert p.force_dof == q.force_dof + 1
    assert t.PointLoadCantilever((160, 40)).force_dof == 161 * 21 * 2  # problems.jl:20
    assert t.HalfMBB((60, 20)).force_dof == (61 * 20 + 2) * 2  # problems.jl:63
    with pytest.raises(ValueError):
        t.PointLoadCantilever((4, 3))
    with pytest.raises(ValueError):
        t.HeatConductionProblem((4, 4), Tleft=1.0)


def test_invalid_arguments_and_no_cpu_fallback(lib):
    t = lib
    from topopt_jl_b200 import _lib

    L = _lib.load()
    out = (C.c_int64 * 4)()
    assert L.topopt_sizes(4, 4, _lib.nels3((2, 2)), None, None, None, None) == _lib.ERR_INVALID
    assert L.topopt_sizes(2, 3, _lib.nels3((2, 2)), None, None, None, None) == _lib.ERR_INVALID
    assert L.topopt_sizes(2, 2, _lib.nels3((0, 2)), None, None, None, None) == _lib.ERR_INVALID
    assert b"nels" in L.topopt_last_error(None)
    import torch

    if not torch.cuda.is_available():
        # the product path must fail loudly, not fall back to the CPU
        with pytest.raises(_lib.TopOptCUDAError) as e:
            t.FEASolver(t.CUDAMatrixFreeSolver, t.HalfMBB((2, 2)))
        assert e.value.code == _lib.ERR_NO_DEVICE
    with pytest.raises(TypeError):
        t.FEASolver(object, t.HalfMBB((2, 2)))


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "topopt.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "topopt_oracle" not in src and "ref_c" not in src and "oracle/" not in src, f


@pytest.mark.parametrize("openmp", [False, True])
def test_c_port_matches_numpy_oracle(openmp):
    import ref_c

    for nels in ((10, 4, 6), (9, 6)):
        dim = len(nels)
        p = o.PointLoadCantilever(nels)
        R = ref_c.RefProblem(dim, dim, nels, p.Ke, p.prescribed + 1, openmp=openmp)
        assert np.array_equal(R.cell_dofs(), p.metadata.cell_dofs + 1)
        rng = np.random.default_rng(5)
        rho = rng.uniform(0.2, 1, p.nel)
        E = o.get_rho(rho, 3.0, 1e-3)
        R.set_density(rho, 3.0, 1e-3)
        x = rng.standard_normal(p.ndof)
        yref = o.matfree_mul(p, E, x)
        assert np.max(np.abs(R.mul(x) - yref)) < 1e-13 * np.max(np.abs(yref))
        b = p.fixedload.copy()
        b[p.prescribed] = 0
        u, it, res = R.cg(b, abstol=0.0, reltol=0.0, maxiter=10)
        uo, ito, reso = o.solve_matfree(p, E, abstol=0.0, reltol=0.0, maxiter=10)
        assert it == ito and abs(res - reso) < 1e-11 * reso and np.max(np.abs(u - uo)) < 1e-10 * np.max(np.abs(uo))
        obj, c, g = R.compliance(uo, rho, 3.0, 1e-3)
        oo, co, go = o.compliance(p, uo, rho, 3.0, 1e-3)
        assert abs(obj - oo) < 1e-12 * abs(oo) and np.max(np.abs(g - go)) < 1e-12 * np.max(np.abs(go))
        assert np.max(np.abs(R.filter(2.0, rho) - o.SensFilter(p, 2.0).pullback(rho))) < 1e-14
        R.close()


def test_modal_pattern_table_covers_brick_element_matrices():
    """The fast hex8 kernel hard-codes which entries of Khat = T' Ke T / 64 can be non-zero
    (modal_pattern() in csrc/kxu_hex8.cuh).  Parse that table and check it against the oracle's Ke
    for cubes, bricks and several Poisson ratios: every non-zero of Khat must be in the table."""
    src = open(os.path.join(ROOT, "topopt.jl_b200", "csrc", "kxu_hex8.cuh")).read()
    body = src[src.index("static const ModalEntry p[45]"):src.index("return p;")]
    pat = [(int(a), int(b)) for a, b in re.findall(r"\{(\d+),\s*(\d+)\}", body)]
    assert len(pat) == 45 and len(set(pat)) == 45
    corners = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
    modes = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (0, 1, 1), (1, 0, 1), (1, 1, 1)]
    H = np.array([[np.prod([(2 * c[d] - 1) ** m[d] for d in range(3)]) for m in modes] for c in corners], dtype=float)
    T = np.kron(H, np.eye(3))
    allowed = np.zeros((24, 24), dtype=bool)
    for r, c in pat:
        allowed[r, c] = True
    assert np.array_equal(allowed, allowed.T)
    for sizes in ((1.0, 1.0, 1.0), (1.0, 0.5, 2.0), (0.3, 0.3, 1.7)):
        for nu in (0.0, 0.3, 0.45):
            Ke = o.element_stiffness(3, sizes, 2.5, nu)
            Kh = T.T @ Ke @ T / 64.0
            assert np.max(np.abs(Kh[~allowed])) < 1e-12 * np.max(np.abs(Kh))
            assert np.max(np.abs(T @ Kh @ T.T - Ke)) < 1e-12 * np.max(np.abs(Ke))  # Ke = T Khat T'
    # an anisotropic / distorted matrix must NOT pass the check (the library then uses the dense kernel)
    rng = np.random.default_rng(0)
    A = rng.standard_normal((24, 24))
    Kh = T.T @ (A @ A.T) @ T / 64.0
    assert np.max(np.abs(Kh[~allowed])) > 1e-3 * np.max(np.abs(Kh))


def test_host_oc_update_matches_oracle(lib):
    """The design update is host code on both sides (MMA is third-party in the reference); the
    package's optimality-criteria step must be the oracle's, otherwise the 10-iteration SIMP parity
    test would compare two different optimisers."""
    t = lib
    rng = np.random.default_rng(12)
    for n in (50, 1000):
        x = rng.uniform(0.05, 1.0, n)
        dc = -rng.uniform(0.0, 5.0, n)
        dv = np.full(n, 1.0 / n)
        for vf in (0.3, 0.5):
            a = t.oc_update(x, dc, dv, vf)
            b = o.oc_update(x, dc, dv, vf)
            assert np.array_equal(a, b)
            assert a.min() >= 0.0 and a.max() <= 1.0 and np.max(np.abs(a - x)) <= 0.2 + 1e-15  # bounds and move limit


C_PROGRAM = r"""
#include "topopt_cuda.h"
#include <stdio.h>
int main(void) {
  int64_t nels[3] = {4, 2, 2}, nn, ne, nd, nz;
  int rc = topopt_sizes(3, 3, nels, &nn, &ne, &nd, &nz);
  printf("%d %lld %lld %lld %lld\n", rc, (long long)nn, (long long)ne, (long long)nd, (long long)nz);
  return rc;
}
"""


def test_header_is_plain_c_and_links(tmp_path):
    """include/topopt_cuda.h is the drop-in boundary: it must compile as pedantic C99 (no C++ or
    torch types in the signatures) and the library must link from a C program."""
    import subprocess

    src = tmp_path / "abi.c"
    src.write_text(C_PROGRAM)
    exe = tmp_path / "abi"
    libdir = os.path.join(ROOT, "topopt.jl_b200")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
           "-L", libdir, "-ltopopt_cuda", f"-Wl,-rpath,{libdir}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert out == ["0", "45", "16", "135", "5733"]


def test_save_mesh_vtu_like_the_reference(tmp_path):
    """save_mesh (src/TopOptProblems/IO/VTK.jl:33-62): cells with density >= 0.5 only, every node, Ferrite local node
    order; a wrong-length density vector raises like the reference's ArgumentError.  Host-only (no GPU)."""
    import base64
    import re

    import topopt_jl_b200 as t

    prob = t.PointLoadCantilever((4, 2, 2))
    rho = np.linspace(0.0, 1.0, prob.nel)
    (path,) = t.save_mesh(str(tmp_path / "design"), prob, rho)
    xml = open(path).read()
    assert f'NumberOfPoints="{prob.nnodes}"' in xml and f'NumberOfCells="{int((rho >= 0.5).sum())}"' in xml
    blob = re.search(r'Name="connectivity" format="binary">([^<]*)<', xml).group(1)
    raw = base64.b64decode(blob)
    conn = np.frombuffer(raw[4:], dtype=np.int64).reshape(-1, 8)
    assert np.array_equal(conn, prob.metadata.cells[rho >= 0.5] - 1)
    dens = np.frombuffer(base64.b64decode(re.search(r'Name="density" format="binary">([^<]*)<', xml).group(1))[4:], dtype=np.float64)
    assert np.array_equal(dens, rho[rho >= 0.5])
    u = np.arange(prob.ndof, dtype=np.float64)
    (path2,) = t.save_mesh(str(tmp_path / "with_u.vtu"), prob, rho, nodal=u)
    disp = np.frombuffer(base64.b64decode(re.search(r'Name="displacement"[^>]*>([^<]*)<', open(path2).read()).group(1))[4:], dtype=np.float64)
    assert np.array_equal(disp.reshape(-1, 3), u[prob.metadata.node_dofs.T - 1])
    with pytest.raises(ValueError):
        t.save_mesh(str(tmp_path / "bad"), prob, rho[:-1])


def test_struct_layouts_match_the_ctypes_and_julia_mirrors(tmp_path):
    """The ABI structs are mirrored field by field in topopt.jl_b200/_lib.py (ctypes) and julia/TopOptCUDA.jl: sizes and
    the offsets of the fields added this round must agree with what a C compiler lays out for include/topopt_cuda.h."""
    import ctypes as C
    import re
    import subprocess

    import topopt_jl_b200 as t

    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stddef.h>\n#include <stdio.h>\n#include "topopt_cuda.h"\n'
        'int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(topopt_cg_opts), offsetof(topopt_cg_opts, variant), '
        'offsetof(topopt_cg_opts, mg_degree), offsetof(topopt_cg_opts, mg_ratio), sizeof(topopt_cg_result), sizeof(topopt_stats), '
        'sizeof(topopt_desc)); return 0; }\n'
    )
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    L = t._lib
    want = [C.sizeof(L.CGOpts), L.CGOpts.variant.offset, L.CGOpts.mg_degree.offset, L.CGOpts.mg_ratio.offset, C.sizeof(L.CGResult),
            C.sizeof(L.Stats), C.sizeof(L.Desc)]
    assert got == want, (got, want)
    # the Julia mirror lists the same fields in the same order (isbits struct: same layout rules as C for these types)
    jl = open(os.path.join(ROOT, "julia", "TopOptCUDA.jl")).read()
    body = re.search(r"struct CGOpts.*?\nend", jl, re.S).group(0)
    fields = re.findall(r"(\w+)::(?:Float64|Int32)", body)
    assert fields == [name for name, _ in L.CGOpts._fields_], fields
