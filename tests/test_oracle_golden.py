"""Pins the CPU oracle (oracle/topopt_oracle.py) against every golden the reference's own tests
hold for this path, and re-runs the reference tests' floating-point *properties* on it.
Golden integers: test/topopt_problems/metadata.jl:53-132, problems.jl:20,63.
There are no golden floating-point vectors in the reference: "parity unpinned" for
floats, which are therefore covered by properties only."""
import json
import os

import numpy as np
import pytest

import topopt_oracle as o


GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_goldens.json")))


def test_golden_metadata_halfmbb_2x2():
    """test/topopt_problems/metadata.jl:53-132 (HalfMBB((2,2),(1.,1.),1.,0.3,1.))"""
    G = GOLD["halfmbb_2x2"]
    p = o.HalfMBB((2, 2))
    assert [list(x) for x in p.grid.nodes] == G["coords"]
    assert (p.grid.cells + 1).tolist() == G["cells"]
    assert (p.metadata.node_dofs + 1).tolist() == G["node_dofs"]
    assert (p.metadata.cell_dofs + 1).tolist() == G["cell_dofs"]
    md = p.metadata
    for d in range(1, md.ndof + 1):  # dof_cells is the inverse of cell_dofs (metadata.jl:97-104)
        for c, l in md.dof_cells_1based(d):
            assert md.cell_dofs[l - 1, c - 1] + 1 == d
    for n, lst in G["node_cells"].items():
        assert md.node_cells_1based(int(n)) == [tuple(x) for x in lst]


def test_golden_force_dofs():
    """test/topopt_problems/problems.jl:20 and :63"""
    assert o.PointLoadCantilever((160, 40)).force_dof + 1 == GOLD["force_dofs"]["PointLoadCantilever_160x40"] == 161 * 21 * 2
    assert o.HalfMBB((60, 20)).force_dof + 1 == GOLD["force_dofs"]["HalfMBB_60x20"] == (61 * 20 + 2) * 2


def test_numbering_loop_equals_vectorised():
    for nels in ((5, 3), (4, 3, 2), (7, 2, 5)):
        g = o.Grid(nels)
        assert np.array_equal(o.ferrite_node_blocks(g.cells, g.nnodes), o.ferrite_node_blocks_loop(g.cells, g.nnodes))


def test_closed_form_numbering_2d():
    """SURVEY Appendix B closed form for the 2-D Ferrite numbering."""
    nx, ny = 6, 5
    md = o.Metadata(o.Grid((nx, ny)), 1)
    W = nx + 1
    for j in range(ny + 1):
        for i in range(nx + 1):
            if j <= 1:
                b = {(0, 0): 0, (1, 0): 1, (1, 1): 2, (0, 1): 3}.get((i, j), 4 + 2 * (i - 2) + j)
            else:
                b = W * j + (1 if i == 0 else 0 if i == 1 else i)
            assert md.node_block[i + W * j] == b


def test_element_matrices_known_answers():
    """SURVEY Appendix E known answers (plane strain in 2-D; finding 9) + symmetry / PSD / rigid modes
    (problems.jl:534-558,573-594 properties)."""
    K3 = o.element_stiffness(3, (1, 1, 1))
    assert abs(K3[0, 0] - 0.235042735042735) < 1e-14 and abs(np.trace(K3) - 5.641025641025641) < 1e-13
    K2 = o.element_stiffness(2, (1, 1))
    assert abs(K2[0, 0] - 0.5769230769230769) < 1e-14 and abs(K2[0, 1] - 0.2403846153846154) < 1e-14
    assert abs(np.trace(K2) - 4.615384615384615) < 1e-13
    for K, nc in ((K3, 3), (K2, 2), (o.element_conductivity(2, (1, 1)), 1), (o.element_conductivity(3, (1, 0.5, 2)), 1)):
        assert np.array_equal(K, K.T)
        w = np.linalg.eigvalsh(K)
        assert w.min() > -1e-13
        for c in range(nc):  # rigid translations carry no energy
            t = np.zeros(K.shape[0])
            t[c::nc] = 1.0
            assert np.max(np.abs(K @ t)) < 1e-14
    # 2-point Gauss is already exact on bricks (Appendix A.2)
    assert np.max(np.abs(o.element_stiffness(3, (1, 0.5, 2), quad_order=2) - o.element_stiffness(3, (1, 0.5, 2)))) < 1e-14


def test_heat_tree_load_and_bcs():
    """problems.jl:456-464 (prescribed-dof count) and the consistent nodal flux load."""
    p = o.HeatTree((8, 6))
    assert len(p.prescribed) == 9
    assert abs(p.fixedload.sum() - 8.0) < 1e-14  # q * top length
    top = np.nonzero(p.grid.top())[0]
    assert np.allclose(np.sort(p.fixedload[p.metadata.node_dofs[0, top]]), [0.5, 0.5] + [1.0] * 7)


@pytest.mark.parametrize("mk", [lambda: o.PointLoadCantilever((8, 4)), lambda: o.PointLoadCantilever((6, 2, 2)), lambda: o.HeatTree((5, 4))])
def test_matrix_free_equals_assembled_on_free_rows(mk):
    """test/FEA/misc.jl:608-660: the two operators agree except on prescribed rows (different meandiag)."""
    p = mk()
    rng = np.random.default_rng(0)
    E = o.get_rho(rng.uniform(0.2, 1, p.nel), 3.0, 1e-3)
    x = rng.standard_normal(p.ndof)
    x[p.prescribed] = 0
    cp, rv, nz, f = o.assemble(p, E)
    y1, y2 = o.matfree_mul(p, E, x), o.csc_mul(cp, rv, nz, x)
    free = ~p.fixed_mask
    assert np.max(np.abs(y1 - y2)[free]) < 1e-13 * np.max(np.abs(y1))
    # structural zeros kept, symmetric
    import scipy.sparse as sp

    K = sp.csc_matrix((nz, rv, cp), shape=(p.ndof, p.ndof))
    assert abs(K - K.T).max() == 0.0


def test_cg_equals_direct_and_pcg():
    """test/FEA/solvers.jl:32-59: CG-assembled and matrix-free u ~ direct u (rtol 1e-4), rho = 0.5"""
    p = o.PointLoadCantilever((10, 4, 4))
    E = o.get_rho(np.full(p.nel, 0.5), 1.0, 1e-3)
    ud = o.solve_direct(p, E)
    for solve in (o.solve_matfree, o.solve_assembled):
        u, it, res = solve(p, E)
        assert np.max(np.abs(u - ud)) < 1e-4 * np.max(np.abs(ud))
        assert np.all(u[p.prescribed] == 0)
    cp, rv, nz, _ = o.assemble(p, E)
    import scipy.sparse as sp

    D = sp.csc_matrix((nz, rv, cp)).diagonal()
    u, it_pcg, _ = o.solve_assembled(p, E, precond_diag=D)
    assert np.max(np.abs(u - ud)) < 1e-4 * np.max(np.abs(ud))
    u, it_e, _ = o.solve_matfree(p, E, criteria="energy", abstol=1e-8)
    assert np.max(np.abs(u - ud)) < 1e-3 * np.max(np.abs(ud))


def test_compliance_gradient_fd():
    """test/Functions/test_common_fns.jl:32-49: 2x2 HalfMBB, p in {1,2,3}, xmin = 0.01, central FD <= 1e-5"""
    p = o.HalfMBB((2, 2))
    rng = np.random.default_rng(1)
    for pen in (1.0, 2.0, 3.0):
        x = rng.uniform(0.1, 1.0, p.nel)

        def comp(r):
            u = o.solve_direct(p, o.get_rho(r, pen, 0.01))
            return o.compliance(p, u, r, pen, 0.01)

        val, _, g = comp(x)
        for e in range(p.nel):
            h = 1e-5
            xp, xm = x.copy(), x.copy()
            xp[e] += h
            xm[e] -= h
            fd = (comp(xp)[0] - comp(xm)[0]) / (2 * h)
            assert abs(fd - g[e]) <= 1e-5 * max(1.0, abs(g[e]))
        # energy balance: sum E_e c_e == f'u
        u = o.solve_direct(p, o.get_rho(x, pen, 0.01))
        assert abs(val - p.fixedload @ u) < 1e-10 * abs(val)


def test_thermal_compliance_properties():
    """test/Functions/test_thermal_compliance.jl:22-140: J == dot(fixedload,u); lambda == -T for
    homogeneous BCs; gradient vs FD; J = 0 with no source."""
    p = o.HeatTree((6, 5))
    x = np.clip(np.random.default_rng(2).uniform(0, 1, p.nel), 0.2, 1.0)
    kw = dict(abstol=1e-13, reltol=1e-14, maxiter=5000)
    J, c, g, T, lam = o.thermal_compliance(p, x, 3.0, 1e-3, **kw)
    assert abs(J - p.fixedload @ T) <= 1e-10 * abs(J)
    assert np.max(np.abs(lam + T)) < 1e-9 * np.max(np.abs(T))
    for e in (0, 7, 29):
        h = 1e-6
        xp, xm = x.copy(), x.copy()
        xp[e] += h
        xm[e] -= h
        fd = (o.thermal_compliance(p, xp, 3.0, 1e-3, **kw)[0] - o.thermal_compliance(p, xm, 3.0, 1e-3, **kw)[0]) / (2 * h)
        assert abs(fd - g[e]) <= 1e-5 * abs(g[e])
    p.fixedload[:] = 0
    assert o.thermal_compliance(p, x, 3.0, 1e-3, **kw)[0] == 0.0


@pytest.mark.parametrize("mk,rmin", [(lambda: o.HalfMBB((6, 4)), 2.0), (lambda: o.PointLoadCantilever((5, 2, 4)), 1.5), (lambda: o.HalfMBB((4, 3), (1.0, 0.5)), 1.3)])
def test_filter_closed_form_equals_reference_bfs(mk, rmin):
    """The literal BFS of CheqFilters.jl:66-118 (duplicates kept) equals the closed form with
    multiplicity m_n = |cells(n)|."""
    p = mk()
    M1, M2 = o.filter_matrices(p, rmin)
    M1b, M2b = o.filter_matrices_bfs(p, rmin)
    assert abs(M1 - M1b).max() < 1e-15 and abs(M2 - M2b).max() < 1e-15
    nodes, w = o.neighbour_info_bfs(p, rmin)
    mult = np.diff(p.metadata.node_cells_offsets)
    for i in (0, p.nel // 2, p.nel - 1):  # every in-radius node is pushed once per adjacent cell
        ids, counts = np.unique(nodes[i], return_counts=True)
        assert np.array_equal(counts, mult[ids])


def test_filter_properties():
    """test/CheqFilters/test_filters.jl:126-244: uniform fields invariant; grad == jacobian' * delta."""
    p = o.PointLoadCantilever((8, 4))
    F, S = o.DensityFilter(p, 2.0), o.SensFilter(p, 2.0)
    assert np.max(np.abs(F(np.full(p.nel, 0.3)) - 0.3)) < 1e-15
    assert np.max(np.abs(S.pullback(np.full(p.nel, -1.7)) + 1.7)) < 1e-15
    rng = np.random.default_rng(3)
    x, d = rng.uniform(0, 1, p.nel), rng.standard_normal(p.nel)
    J = (F.M2 @ F.M1).toarray()
    assert np.allclose(F.pullback(d), J.T @ d, rtol=1e-10, atol=0)
    assert np.array_equal(S(x), x) and not np.allclose(S.pullback(d), d)
    with pytest.raises(ValueError):  # too-small rmin -> ArgumentError (test_filters.jl:71-73)
        o.DensityFilter(p, 0.2)


def test_penalty_formulas():
    """test/Utilities/test_penalties.jl: get_rho / get_rho_drho for both preference settings."""
    x = np.linspace(0.05, 1, 7)
    for first in (True, False):
        r, dr = o.get_rho_drho(x, 3.0, 1e-3, first)
        assert np.allclose(r, o.get_rho(x, 3.0, 1e-3, first))
        h = 1e-7
        fd = (o.get_rho(x + h, 3.0, 1e-3, first) - o.get_rho(x - h, 3.0, 1e-3, first)) / (2 * h)
        assert np.allclose(dr, fd, rtol=1e-6)
    assert np.allclose(o.get_rho_drho(x, 3.0, 1e-3)[1], (1 - 1e-3) * 3 * x**2)  # SURVEY finding 6


def test_projected_penalty_formulas():
    """ProjectedPenaltyFun = penalty(proj(x)) (penalties.jl:62-69): derivative vs central FD, both
    projections, both preference orders; beta -> 0 Heaviside tends to the identity."""
    x = np.linspace(0.05, 0.95, 9)
    for proj in (o.heaviside_projection, o.sigmoid_projection):
        for first in (True, False):
            r, dr = o.get_rho_drho_projected(x, 3.0, 1e-3, proj, 4.0, first)
            h = 1e-6
            rp = o.get_rho_drho_projected(x + h, 3.0, 1e-3, proj, 4.0, first)[0]
            rm = o.get_rho_drho_projected(x - h, 3.0, 1e-3, proj, 4.0, first)[0]
            assert np.allclose(dr, (rp - rm) / (2 * h), rtol=1e-6)
    y, dy = o.heaviside_projection(x, 1e-9)
    assert np.allclose(y, x, atol=1e-8) and np.allclose(dy, 1.0, atol=1e-8)
    assert abs(o.heaviside_projection(np.array([1.0]), 10.0)[0][0] - 1.0) < 1e-15  # proj(1) = 1
