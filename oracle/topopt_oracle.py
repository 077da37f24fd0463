"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

A numpy/scipy restatement of the TopOpt.jl (v0.14.0) linear-elastic / thermal SIMP
inner loop, used only by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` leg of ``bench.py`` as the checker for ``libtopopt_cuda``.
Nothing under ``topopt.jl_b200/`` imports this file.

Pinning status
--------------
* Integer data (grid connectivity, Ferrite DOF numbering, cell_dofs, node_dofs,
  node_cells, force dofs) is pinned against the reference's own golden integers
  (``test/topopt_problems/metadata.jl:53-132``, ``problems.jl:20,63``) by
  ``tests/test_oracle_golden.py``.
* Floating point (Ke, compliance, sensitivities, filters) is **parity unpinned**:
  the reference has no golden floating-point vectors for this path and Julia is
  not installed here, so the reference cannot be run.  It is pinned only through
  the reference tests' own *properties* (FD gradient checks, CG == direct solve,
  uniform-field filter invariance, J'Δ identity, J == Q.T), which the same test
  file re-runs on this restatement.

Third-party arithmetic restated from the published algorithms (not vendored under
/root/reference): Ferrite.jl 1.x (generate_grid, DofHandler close!, allocate_matrix,
assemble!, apply!), IterativeSolvers.jl 0.9 (cg!), Preconditioners.jl 0.6
(DiagonalPreconditioner).  Call sites are cited per function.

All indices are 0-based internally; ``*_1based`` helpers exist for the goldens.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

# --------------------------------------------------------------------------------------
# Grid  (Ferrite.generate_grid via src/TopOptProblems/grids.jl:80-92)
# --------------------------------------------------------------------------------------


class Grid:
    """Structured quad4/hex8 grid; nodes and cells x-fastest (Ferrite generate_grid)."""

    def __init__(self, nels, sizes=None):
        self.nels = tuple(int(n) for n in nels)
        self.dim = len(self.nels)
        self.sizes = tuple(float(s) for s in (sizes if sizes is not None else (1.0,) * self.dim))
        nn = tuple(n + 1 for n in self.nels)
        self.nnodes_per_dim = nn
        axes = [np.linspace(0.0, self.nels[d] * self.sizes[d], nn[d]) for d in range(self.dim)]
        if self.dim == 2:
            Y, X = np.meshgrid(axes[1], axes[0], indexing="ij")
            self.nodes = np.stack([X.ravel(), Y.ravel()], axis=1)
            j, i = np.meshgrid(np.arange(self.nels[1]), np.arange(self.nels[0]), indexing="ij")
            i = i.ravel()
            j = j.ravel()
            W = nn[0]
            n00 = i + W * j
            self.cells = np.stack([n00, n00 + 1, n00 + 1 + W, n00 + W], axis=1).astype(np.int64)
        else:
            Z, Y, X = np.meshgrid(axes[2], axes[1], axes[0], indexing="ij")
            self.nodes = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
            k, j, i = np.meshgrid(
                np.arange(self.nels[2]), np.arange(self.nels[1]), np.arange(self.nels[0]), indexing="ij"
            )
            i, j, k = i.ravel(), j.ravel(), k.ravel()
            W = nn[0]
            P = nn[0] * nn[1]
            n0 = i + W * j + P * k
            bot = [n0, n0 + 1, n0 + 1 + W, n0 + W]
            self.cells = np.stack(bot + [b + P for b in bot], axis=1).astype(np.int64)
        self.nnodes = self.nodes.shape[0]
        self.nel = self.cells.shape[0]
        self.corners = (np.zeros(self.dim), np.array([self.nels[d] * self.sizes[d] for d in range(self.dim)]))

    # node-set predicates, src/TopOptProblems/grids.jl:111-125 (isapprox)
    def left(self):
        return np.isclose(self.nodes[:, 0], self.corners[0][0])

    def right(self):
        return np.isclose(self.nodes[:, 0], self.corners[1][0])

    def bottom(self):
        return np.isclose(self.nodes[:, 1], self.corners[0][1])

    def top(self):
        return np.isclose(self.nodes[:, 1], self.corners[1][1])

    def middley(self):
        return np.isclose(self.nodes[:, 1], 0.5 * (self.corners[0][1] + self.corners[1][1]))


# --------------------------------------------------------------------------------------
# Ferrite DofHandler numbering + TopOpt Metadata (src/TopOptProblems/metadata.jl:27-145)
# --------------------------------------------------------------------------------------


def ferrite_node_blocks(cells, nnodes):
    """Ferrite ``close!(dh)``: cells ascending, local nodes ascending, a node gets the next
    block of ``ncomp`` dofs the first time it is seen.  Returns block index per node."""
    flat = cells.ravel()
    _, first = np.unique(flat, return_index=True)  # first occurrence of node id 0..nnodes-1
    assert first.shape[0] == nnodes
    order = np.argsort(first, kind="stable")  # nodes in order of first visit
    block = np.empty(nnodes, dtype=np.int64)
    block[order] = np.arange(nnodes)
    return block


def ferrite_node_blocks_loop(cells, nnodes):
    """Literal loop version of the above (small meshes; used to cross-check)."""
    block = -np.ones(nnodes, dtype=np.int64)
    nxt = 0
    for c in range(cells.shape[0]):
        for n in cells[c]:
            if block[n] < 0:
                block[n] = nxt
                nxt += 1
    return block


class Metadata:
    """cell_dofs (Kesize x nel), dof_cells (ragged), node_cells (ragged), node_dofs (ncomp x nnodes)."""

    def __init__(self, grid: Grid, ncomp: int):
        self.ncomp = ncomp
        block = ferrite_node_blocks(grid.cells, grid.nnodes)
        self.node_block = block
        # node_dofs[c, n]   (metadata.jl:116-145)
        self.node_dofs = (ncomp * block[None, :] + np.arange(ncomp)[:, None]).astype(np.int64)
        N = grid.cells.shape[1]
        # cell_dofs[ncomp*a + c, e]  (metadata.jl:40-50; Ferrite celldofs: node-major, comp-minor)
        cd = ncomp * block[grid.cells][:, :, None] + np.arange(ncomp)[None, None, :]
        self.cell_dofs = cd.reshape(grid.nel, N * ncomp).T.copy()
        self.ndof = ncomp * grid.nnodes
        self.kesize = N * ncomp
        # dof_cells: for each dof, (cell, local) pairs in ascending cell order (metadata.jl:64-76)
        flat_dof = self.cell_dofs.T.ravel()  # cell-major, local-minor == push order
        order = np.argsort(flat_dof, kind="stable")
        self.dof_cells_offsets = np.zeros(self.ndof + 1, dtype=np.int64)
        np.add.at(self.dof_cells_offsets, flat_dof + 1, 1)
        self.dof_cells_offsets = np.cumsum(self.dof_cells_offsets)
        self.dof_cells_cell = (order // self.kesize).astype(np.int64)
        self.dof_cells_local = (order % self.kesize).astype(np.int64)
        # node_cells: (cell, local node) pairs, ascending cell order (metadata.jl:99-109)
        flat_node = grid.cells.ravel()
        order = np.argsort(flat_node, kind="stable")
        self.node_cells_offsets = np.zeros(grid.nnodes + 1, dtype=np.int64)
        np.add.at(self.node_cells_offsets, flat_node + 1, 1)
        self.node_cells_offsets = np.cumsum(self.node_cells_offsets)
        self.node_cells_cell = (order // N).astype(np.int64)
        self.node_cells_local = (order % N).astype(np.int64)

    def node_cells_1based(self, n1):
        r = slice(self.node_cells_offsets[n1 - 1], self.node_cells_offsets[n1])
        return [(int(c) + 1, int(l) + 1) for c, l in zip(self.node_cells_cell[r], self.node_cells_local[r])]

    def dof_cells_1based(self, d1):
        r = slice(self.dof_cells_offsets[d1 - 1], self.dof_cells_offsets[d1])
        return [(int(c) + 1, int(l) + 1) for c, l in zip(self.dof_cells_cell[r], self.dof_cells_local[r])]


# --------------------------------------------------------------------------------------
# Element matrices (src/TopOptProblems/matrices_and_vectors.jl:63-177, 413-496)
# --------------------------------------------------------------------------------------

_REF_NODES_2D = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], dtype=float)
_REF_NODES_3D = np.array(
    [[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], dtype=float
)


def _shape_gradients(dim, xi):
    """d phi_a / d xi_d at reference point xi for order-1 Lagrange quad/hex."""
    ref = _REF_NODES_2D if dim == 2 else _REF_NODES_3D
    N = ref.shape[0]
    g = np.zeros((N, dim))
    for a in range(N):
        for d in range(dim):
            v = ref[a, d] / 2.0
            for o in range(dim):
                if o != d:
                    v *= (1.0 + ref[a, o] * xi[o]) / 2.0
            g[a, d] = v
    return g


def _quadrature(dim, order):
    x, w = np.polynomial.legendre.leggauss(order)
    pts, wts = [], []
    if dim == 2:
        for j in range(order):
            for i in range(order):
                pts.append((x[i], x[j]))
                wts.append(w[i] * w[j])
    else:
        for k in range(order):
            for j in range(order):
                for i in range(order):
                    pts.append((x[i], x[j], x[k]))
                    wts.append(w[i] * w[j] * w[k])
    return np.array(pts), np.array(wts)


def element_stiffness(dim, sizes, E=1.0, nu=0.3, quad_order=4):
    """Elasticity Ke = int B'CB, C from 3-D Lame constants for any dim (plane strain in 2-D),
    matrices_and_vectors.jl:83-87,148-166; Gauss order 4 (solvers_api.jl:432-444);
    stored as Symmetric(upper triangle)."""
    lam = E * nu / ((1 + nu) * (1 - 2 * nu))
    mu = E / (2 * (1 + nu))
    N = 4 if dim == 2 else 8
    pts, wts = _quadrature(dim, quad_order)
    h = np.asarray(sizes, dtype=float)
    detJ = np.prod(h / 2.0)
    Ke = np.zeros((dim * N, dim * N))
    for q in range(len(wts)):
        dO = detJ * wts[q]
        g = _shape_gradients(dim, pts[q]) / (h / 2.0)[None, :]
        for b in range(N):
            for a in range(N):
                ga, gb = g[a], g[b]
                # dotdot(ga, C, gb)[j,k] = ga_i C_ijkl gb_l
                blk = lam * np.outer(ga, gb) + mu * (np.outer(gb, ga) + np.dot(ga, gb) * np.eye(dim))
                Ke[dim * a : dim * a + dim, dim * b : dim * b + dim] += blk * dO
    U = np.triu(Ke)
    return U + np.triu(Ke, 1).T


def element_conductivity(dim, sizes, k=1.0, quad_order=4):
    """Heat Ke[a,b] = int k grad(phi_a).grad(phi_b), matrices_and_vectors.jl:455-496."""
    N = 4 if dim == 2 else 8
    pts, wts = _quadrature(dim, quad_order)
    h = np.asarray(sizes, dtype=float)
    detJ = np.prod(h / 2.0)
    Ke = np.zeros((N, N))
    for q in range(len(wts)):
        dO = detJ * wts[q]
        g = _shape_gradients(dim, pts[q]) / (h / 2.0)[None, :]
        Ke += k * (g @ g.T) * dO
    U = np.triu(Ke)
    return U + np.triu(Ke, 1).T


# --------------------------------------------------------------------------------------
# Problems (src/TopOptProblems/problem_types.jl:162-231, 334-402, 926-1088)
# --------------------------------------------------------------------------------------


class Problem:
    """grid + metadata + prescribed dofs + fixedload + Ke + cellvolumes."""

    def __init__(self, grid, md, Ke, prescribed, fixedload, physics):
        self.grid = grid
        self.ncomp = md.ncomp
        self.physics = physics
        self.metadata = md
        self.Ke = Ke
        self.prescribed = np.unique(np.asarray(prescribed, dtype=np.int64))  # sorted (ch.prescribed_dofs)
        self.fixedload = fixedload
        self.ndof = self.metadata.ndof
        self.nel = grid.nel
        self.cellvolumes = np.full(grid.nel, float(np.prod(grid.sizes)))  # elementinfo.jl:180-192
        fixed = np.zeros(self.ndof, dtype=bool)
        fixed[self.prescribed] = True
        self.fixed_mask = fixed
        # per-element BC mask (elementmatrix.jl:64-109): mask[l,e] False where local dof is prescribed
        self.cell_mask = ~fixed[self.metadata.cell_dofs]
        # matrix-free fixed-dof diagonal = sum_e tr(Ke)  (solvers_api.jl:526-527)
        self.meandiag_mf = float(np.trace(Ke)) * grid.nel


def PointLoadCantilever(nels, sizes=None, E=1.0, nu=0.3, force=1.0):
    g = Grid(nels, sizes)
    if g.nels[1] % 2 or (g.dim == 3 and g.nels[2] % 2):
        raise ValueError("Grid does not have an even number of elements along the y and/or z axes.")
    md = Metadata(g, g.dim)
    fixed_nodes = np.nonzero(g.left())[0]
    prescribed = md.node_dofs[:, fixed_nodes].ravel()
    fnode = np.nonzero(g.right() & g.middley())[0][0]  # lowest node id of "down_force" (problem_types.jl:223)
    f = np.zeros(md.ndof)
    force_dof = md.node_dofs[1, fnode]
    f[force_dof] = -force  # getcloaddict, problem_types.jl:398-402
    p = Problem(g, md, element_stiffness(g.dim, g.sizes, E, nu), prescribed, f, "elasticity")
    p.force_dof = int(force_dof)
    return p


def HalfMBB(nels, sizes=None, E=1.0, nu=0.3, force=1.0):
    g = Grid(nels, sizes)
    md = Metadata(g, g.dim)
    u1 = np.nonzero(g.left())[0]
    u2 = np.nonzero(g.bottom() & g.right())[0]
    prescribed = np.concatenate([md.node_dofs[0, u1], md.node_dofs[1, u2]])
    fnode = np.nonzero(g.top() & g.left())[0][0]
    f = np.zeros(md.ndof)
    force_dof = md.node_dofs[1, fnode]
    f[force_dof] = -force
    p = Problem(g, md, element_stiffness(g.dim, g.sizes, E, nu), prescribed, f, "elasticity")
    p.force_dof = int(force_dof)
    return p


def HeatTree(nels, sizes=None, k=1.0, q=1.0):
    """HeatConductionProblem with Tbottom=0 and heat flux q on the "top" facet set
    (problem_types.jl:1074-1088).  Consistent nodal load: each top-edge (2-D) / top-face (3-D)
    node receives q * (facet area) / (nodes per facet) per adjacent facet
    (matrices_and_vectors.jl:288-339 + assemble.jl:145-160)."""
    g = Grid(nels, sizes)
    md = Metadata(g, 1)
    bnodes = np.nonzero(g.bottom())[0]
    prescribed = md.node_dofs[0, bnodes]
    f = np.zeros(md.ndof)
    # cells whose top facet lies on y = ymax, ascending cell order (update_f!, assemble.jl:151-160)
    if g.dim == 2:
        top_cells = np.nonzero(np.isclose(g.nodes[g.cells[:, 2], 1], g.corners[1][1]))[0]
        loc = (2, 3)
        area = g.sizes[0]
    else:
        # Ferrite's 3-D "top" facet set is z = zmax (facet 6, local nodes 5..8) while TopOpt's
        # "bottom_boundary" node set is y = ymin (grids.jl:113); both are restated literally.
        top_cells = np.nonzero(np.isclose(g.nodes[g.cells[:, 4], 2], g.corners[1][2]))[0]
        loc = (4, 5, 6, 7)
        area = g.sizes[0] * g.sizes[1]
    share = q * area / len(loc)
    for l in loc:  # noqa: E741
        np.add.at(f, md.node_dofs[0, g.cells[top_cells, l]], share)
    return Problem(g, md, element_conductivity(g.dim, g.sizes, k), prescribed, f, "heat")


# --------------------------------------------------------------------------------------
# Penalty / interpolation (src/Utilities/penalties.jl:30-33,113-130; utils.jl:77)
# --------------------------------------------------------------------------------------


def get_rho(x, p, xmin, penalty_before_interpolation=True):
    if penalty_before_interpolation:
        return x**p * (1 - xmin) + xmin
    return (x * (1 - xmin) + xmin) ** p


def get_rho_drho(x, p, xmin, penalty_before_interpolation=True):
    if penalty_before_interpolation:
        return x**p * (1 - xmin) + xmin, (1 - xmin) * p * x ** (p - 1)
    d = x * (1 - xmin) + xmin
    return d**p, p * d ** (p - 1) * (1 - xmin)


def heaviside_projection(x, beta):
    """HeavisideProjectionFun (penalties.jl:83-86) and its derivative."""
    return 1 - np.exp(-beta * x) + x * np.exp(-beta), beta * np.exp(-beta * x) + np.exp(-beta)


def sigmoid_projection(x, beta):
    """SigmoidProjectionFun (penalties.jl:93-96) and its derivative."""
    e = np.exp((beta + 1) * (0.5 - x))
    return 1 / (1 + e), (beta + 1) * e / (1 + e) ** 2


def get_rho_drho_projected(x, p, xmin, proj, beta, penalty_before_interpolation=True):
    """ProjectedPenaltyFun with a power penalty: penalty(proj(.)) inside get_rho (penalties.jl:62-69,113-130)."""
    a = x if penalty_before_interpolation else x * (1 - xmin) + xmin
    pa, dpa = proj(a, beta)
    f, df = pa**p, p * pa ** (p - 1) * dpa
    if penalty_before_interpolation:
        return f * (1 - xmin) + xmin, (1 - xmin) * df
    return f, df * (1 - xmin)


# --------------------------------------------------------------------------------------
# Matrix-free operator (src/FEA/matrix_free_operator.jl:66-105)
# --------------------------------------------------------------------------------------


def _ragged_sum(offsets, values_idx, flat_values, n):
    """y[d] = sum over slots of flat_values[values_idx[slot]] in slot order (sequential, like the
    reference's `yi += ...` loop)."""
    y = np.zeros(n)
    counts = np.diff(offsets)
    for s in range(int(counts.max())):
        sel = np.nonzero(counts > s)[0]
        y[sel] += flat_values[values_idx[offsets[sel] + s]]
    return y


def matfree_mul(prob: Problem, E, x):
    """y = K(rho) x exactly as MatrixFreeOperator mul!: pass 1 per element
    xes = E_e*(bcmatrix(Ke)*x_e), pass 2 fixed dofs y=meandiag*x, free dofs ragged gather."""
    md = prob.metadata
    xe = x[md.cell_dofs] * prob.cell_mask  # masked columns
    acc = np.zeros_like(xe)
    for j in range(md.kesize):  # column-ordered accumulation
        acc += prob.Ke[:, j][:, None] * xe[j][None, :]
    xes = (acc * prob.cell_mask) * E[None, :]
    flat = xes.T.ravel()  # index = cell*kesize + local
    idx = md.dof_cells_cell * md.kesize + md.dof_cells_local
    y = _ragged_sum(md.dof_cells_offsets, idx, flat, md.ndof)
    y[prob.prescribed] = prob.meandiag_mf * x[prob.prescribed]
    return y


# --------------------------------------------------------------------------------------
# Assembled path (Ferrite allocate_matrix / assemble! / apply!, src/TopOptProblems/assemble.jl:28-90)
# --------------------------------------------------------------------------------------


def csc_pattern(prob: Problem):
    """Sparsity of allocate_matrix(dh): all dof pairs sharing a cell, both triangles, sorted rows.
    Returns (colptr, rowval) 0-based, int64."""
    md = prob.metadata
    ks = md.kesize
    rows = np.repeat(md.cell_dofs.T, ks, axis=1).ravel()  # (e, a, b) -> row = dofs[a]
    cols = np.tile(md.cell_dofs.T, (1, ks)).ravel()  # col = dofs[b]
    key = np.unique(cols * md.ndof + rows)
    colidx = key // md.ndof
    rowval = (key % md.ndof).astype(np.int64)
    colptr = np.zeros(md.ndof + 1, dtype=np.int64)
    np.add.at(colptr, colidx + 1, 1)
    return np.cumsum(colptr), rowval


def assemble(prob: Problem, E, apply_bc=True):
    """K.nzval as Ferrite assemble! would produce: per entry, terms fl(E_e*Ke[a,b]) summed in
    ascending cell order; then apply! (homogeneous): m = mean|diag K|, zero prescribed rows+cols,
    K[d,d]=m, f[d]=0.  Returns (colptr, rowval, nzval, f)."""
    md = prob.metadata
    ks = md.kesize
    colptr, rowval = csc_pattern(prob)
    nnz = rowval.shape[0]
    rows = np.repeat(md.cell_dofs.T, ks, axis=1).ravel()
    cols = np.tile(md.cell_dofs.T, (1, ks)).ravel()
    vals = (E[:, None, None] * prob.Ke[None, :, :]).reshape(-1)  # [e, a, b]
    key = cols * md.ndof + rows
    order = np.argsort(key, kind="stable")  # stable -> ascending cell within an entry
    skey = key[order]
    svals = vals[order]
    start = np.nonzero(np.concatenate([[True], skey[1:] != skey[:-1]]))[0]
    offsets = np.concatenate([start, [skey.shape[0]]])
    assert start.shape[0] == nnz
    nzval = _ragged_sum(offsets, np.arange(skey.shape[0]), svals, nnz)
    f = prob.fixedload.copy()
    if apply_bc:
        K = sp.csc_matrix((nzval, rowval, colptr), shape=(md.ndof, md.ndof))
        d = K.diagonal()
        m = 0.0
        for v in np.abs(d):  # sequential sum like Ferrite meandiag
            m += v
        m /= md.ndof
        colidx = np.repeat(np.arange(md.ndof), np.diff(colptr))
        kill = prob.fixed_mask[rowval] | prob.fixed_mask[colidx]
        nzval = np.where(kill, 0.0, nzval)
        diag = (rowval == colidx) & prob.fixed_mask[rowval]
        nzval[diag] = m
        f[prob.prescribed] = 0.0
    return colptr, rowval, nzval, f


def csc_mul(colptr, rowval, nzval, x):
    n = colptr.shape[0] - 1
    return sp.csc_matrix((nzval, rowval, colptr), shape=(n, n)) @ x


# --------------------------------------------------------------------------------------
# CG / PCG (IterativeSolvers.jl 0.9 cg!, called at src/FEA/solvers_api.jl:196-219)
# --------------------------------------------------------------------------------------


def cg(A_mul, b, abstol=1e-7, reltol=None, maxiter=700, precond_diag=None, criteria="default", history=None):
    """x0 = 0 (cg_solve! zero-fills lhs), r = b, tol = max(reltol*||r0||, abstol),
    reltol = sqrt(eps).  Returns (x, iters, residual).  `criteria="energy"` restates
    src/FEA/convergence_criteria.jl:26-45."""
    if reltol is None:
        reltol = np.sqrt(np.finfo(float).eps)
    x = np.zeros_like(b)
    u = np.zeros_like(b)
    r = b.copy()
    res = float(np.linalg.norm(r))
    tol = max(reltol * res, abstol)
    prev = 1.0
    rho = 1.0
    it = 0
    energy = 0.0

    def converged():
        nonlocal energy
        if criteria == "default":
            return res <= tol
        xtr = float(x @ r)
        xAx = float(b @ x) - xtr
        change = xAx - energy
        if np.isnan(change) or np.isnan(xAx):
            raise FloatingPointError("EnergyCriteria: NaN detected")
        if xAx < 0:
            raise FloatingPointError("EnergyCriteria: expected non-negative xAx")
        with np.errstate(divide="ignore", invalid="ignore"):
            ok = (np.abs(change) / xAx <= tol) and (np.abs(xtr) / xAx <= tol)
        energy = xAx
        return bool(ok)

    while it < maxiter and not converged():
        if precond_diag is None:
            beta = res**2 / prev**2
            u = r + beta * u
            c = A_mul(u)
            alpha = res**2 / float(u @ c)
        else:
            c = r / precond_diag
            rho_prev = rho
            rho = float(c @ r)
            beta = rho / rho_prev
            u = c + beta * u
            c = A_mul(u)
            alpha = rho / float(u @ c)
        x = x + alpha * u
        r = r - alpha * c
        prev = res
        res = float(np.linalg.norm(r))
        it += 1
        if history is not None:
            history.append(res)
    return x, it, res


def solve_matfree(prob, E, rhs=None, **kw):
    b = prob.fixedload.copy() if rhs is None else rhs.copy()
    b[prob.prescribed] = 0.0  # apply!(f) / apply_zero!(rhs): solvers_api.jl:364-367, assemble.jl:87
    return cg(lambda v: matfree_mul(prob, E, v), b, **kw)


def solve_assembled(prob, E, rhs=None, **kw):
    colptr, rowval, nzval, f = assemble(prob, E)
    b = f if rhs is None else rhs.copy()
    b[prob.prescribed] = 0.0
    n = prob.ndof
    K = sp.csc_matrix((nzval, rowval, colptr), shape=(n, n))
    return cg(lambda v: K @ v, b, **kw)


def solve_direct(prob, E, rhs=None):
    colptr, rowval, nzval, f = assemble(prob, E)
    b = f if rhs is None else rhs.copy()
    b[prob.prescribed] = 0.0
    n = prob.ndof
    K = sp.csc_matrix((nzval, rowval, colptr), shape=(n, n))
    return spla.spsolve(K, b)


# --------------------------------------------------------------------------------------
# Compliance + sensitivity (src/Functions/compute_element_energy.jl:18-38, compliance.jl:58-76)
# --------------------------------------------------------------------------------------


def element_energy(prob, u, v=None):
    """c_e = v_e' Ke u_e with the raw (unmasked) Ke; accumulation order w outer, v inner."""
    md = prob.metadata
    ue = u[md.cell_dofs]
    ve = ue if v is None else v[md.cell_dofs]
    c = np.zeros(prob.nel)
    for w in range(md.kesize):
        for vv in range(md.kesize):
            c += ve[vv] * prob.Ke[vv, w] * ue[w]
    return c


def compliance(prob, u, rho, p, xmin):
    """returns (obj, cell_comp, grad): obj = sum E_e c_e (sequential), grad_e = -dE_e c_e."""
    c = element_energy(prob, u)
    E, dE = get_rho_drho(rho, p, xmin)
    obj = 0.0
    for t in E * c:
        obj += t
    return obj, c, -dE * c


def thermal_compliance(prob, rho, p, xmin, **cgkw):
    """src/Functions/thermal_compliance.jl:116-158: J = Q.T, adjoint solve K lam = -Q_cond,
    grad_e = dE_e * lam_e' Ke T_e."""
    E, dE = get_rho_drho(rho, p, xmin)
    T, it1, _ = solve_matfree(prob, E, **cgkw)
    obj = float(prob.fixedload @ T)
    rhs = -prob.fixedload.copy()
    rhs[prob.prescribed] = 0.0
    lam, it2, _ = solve_matfree(prob, E, rhs=rhs, **cgkw)
    c = element_energy(prob, T, lam)
    return obj, c, dE * c, T, lam


def volume(prob, x):
    """src/Functions/volume.jl:59-67"""
    return float(x @ prob.cellvolumes) / float(prob.cellvolumes.sum())


# --------------------------------------------------------------------------------------
# Filters (src/CheqFilters/CheqFilters.jl:66-118, density_filter.jl:35-121, sens_filter.jl:52-110)
# --------------------------------------------------------------------------------------


def neighbour_info_bfs(prob, rmin):
    """Literal restatement of get_neighbour_info (BFS, duplicates kept).  Small meshes only."""
    g = prob.grid
    md = prob.metadata
    all_nodes, all_w = [], []
    for cell in range(g.nel):
        center = g.nodes[g.cells[cell]].mean(axis=0)
        queue = [cell]
        visited = {cell}
        nn, ww = [], []
        while queue:
            cid = queue.pop(0)
            for n in g.cells[cid]:
                dist = float(np.linalg.norm(g.nodes[n] - center))
                if dist < rmin:
                    nn.append(int(n))
                    ww.append(max(rmin - dist, 0.0))
                    r = slice(md.node_cells_offsets[n], md.node_cells_offsets[n + 1])
                    for c in md.node_cells_cell[r]:
                        if int(c) not in visited:
                            queue.append(int(c))
                            visited.add(int(c))
        all_nodes.append(nn)
        all_w.append(ww)
    return all_nodes, all_w


def filter_M1(prob):
    """cell->node volume-weighted mean (density_filter.jl:62-73, sens_filter.jl:72-96)."""
    g = prob.grid
    md = prob.metadata
    nc = np.diff(md.node_cells_offsets)
    rows = np.repeat(np.arange(g.nnodes), nc)
    cols = md.node_cells_cell
    vals = prob.cellvolumes[cols]
    M1 = sp.csr_matrix((vals, (rows, cols)), shape=(g.nnodes, g.nel))
    return (sp.diags(1.0 / np.asarray(M1.sum(axis=1)).ravel()) @ M1).tocsr()


def filter_matrices(prob, rmin):
    """Closed form of the two filter stages:
    M1[n,c] = V_c / sum_{c' in cells(n)} V_c'   (nnodes x nel)
    M2[i,n] = m_n (rmin-d_in) / sum_n m_n (rmin-d_in),  d_in < rmin strictly, m_n = |cells(n)|.
    Returns scipy CSR (M1, M2)."""
    g = prob.grid
    md = prob.metadata
    nc = np.diff(md.node_cells_offsets)
    M1 = filter_M1(prob)
    # M2 via offsets on the structured grid
    h = np.array(g.sizes)
    reach = [int(np.ceil(rmin / h[d] + 0.5)) + 1 for d in range(g.dim)]
    nn = g.nnodes_per_dim
    cent = g.nodes[g.cells].mean(axis=1)
    if g.dim == 2:
        ci = np.arange(g.nel) % g.nels[0]
        cj = np.arange(g.nel) // g.nels[0]
        cidx = [ci, cj]
    else:
        ci = np.arange(g.nel) % g.nels[0]
        cj = (np.arange(g.nel) // g.nels[0]) % g.nels[1]
        ck = np.arange(g.nel) // (g.nels[0] * g.nels[1])
        cidx = [ci, cj, ck]
    R, C, V = [], [], []
    ranges = [range(-reach[d] + 1, reach[d] + 1) for d in range(g.dim)]
    import itertools

    for off in itertools.product(*ranges):
        idx = [cidx[d] + off[d] for d in range(g.dim)]
        ok = np.ones(g.nel, dtype=bool)
        for d in range(g.dim):
            ok &= (idx[d] >= 0) & (idx[d] < nn[d])
        e = np.nonzero(ok)[0]
        if g.dim == 2:
            n = idx[0][e] + nn[0] * idx[1][e]
        else:
            n = idx[0][e] + nn[0] * (idx[1][e] + nn[1] * idx[2][e])
        dist = np.linalg.norm(g.nodes[n] - cent[e], axis=1)
        inr = dist < rmin
        e, n, dist = e[inr], n[inr], dist[inr]
        R.append(e)
        C.append(n)
        V.append(nc[n] * (rmin - dist))
    R, C, V = np.concatenate(R), np.concatenate(C), np.concatenate(V)
    M2 = sp.csr_matrix((V, (R, C)), shape=(g.nel, g.nnodes))
    s = np.asarray(M2.sum(axis=1)).ravel()
    if np.all(s == 0):
        raise ValueError("DensityFilterFun: no neighbouring nodes were found within the filter radius `rmin`")
    s[s == 0] = 1.0
    M2 = sp.diags(1.0 / s) @ M2
    return M1.tocsr(), M2.tocsr()


def filter_matrices_bfs(prob, rmin):
    """Same matrices from the literal BFS lists (density_filter.jl:49-108)."""
    g = prob.grid
    M1 = filter_M1(prob)
    nodes, weights = neighbour_info_bfs(prob, rmin)
    R, C, V = [], [], []
    for i in range(g.nel):
        if not nodes[i]:
            continue
        s = sum(weights[i])
        for n, w in zip(nodes[i], weights[i]):
            R.append(i)
            C.append(n)
            V.append(w / s)
    M2 = sp.csr_matrix((V, (R, C)), shape=(g.nel, g.nnodes))  # duplicates summed like sparse()
    return M1, M2


class DensityFilter:
    """forward J x, pullback J' d with J = M2 M1 (density_filter.jl:35-47)."""

    def __init__(self, prob, rmin):
        self.M1, self.M2 = filter_matrices(prob, rmin)

    def __call__(self, x):
        return self.M2 @ (self.M1 @ x)

    def pullback(self, d):
        return self.M1.T @ (self.M2.T @ d)


class SensFilter:
    """forward identity; pullback = forward two-stage map applied to the cotangent
    (sens_filter.jl:52-110)."""

    def __init__(self, prob, rmin):
        self.M1, self.M2 = filter_matrices(prob, rmin)

    def __call__(self, x):
        return x

    def pullback(self, d):
        return self.M2 @ (self.M1 @ d)


# --------------------------------------------------------------------------------------
# Host-side optimality-criteria update shared by the parity runs (MMA is third-party;
# "10-iteration design <= 1e-6 with a shared host-side OC update").
# --------------------------------------------------------------------------------------


def oc_update(x, dc, dv, volfrac, move=0.2, eta=0.5, xlo=0.0):
    l1, l2 = 0.0, 1e9
    dc = np.minimum(dc, 0.0)
    while (l2 - l1) / (l1 + l2) > 1e-12 and l2 > 1e-40:
        lmid = 0.5 * (l1 + l2)
        xn = np.maximum(
            xlo, np.maximum(x - move, np.minimum(1.0, np.minimum(x + move, x * (-dc / dv / lmid) ** eta)))
        )
        if float(xn @ dv) > volfrac:
            l1 = lmid
        else:
            l2 = lmid
    return xn


def simp_loop(prob, rmin, volfrac, p=3.0, xmin=1e-3, iters=10, filt="density", abstol=1e-10, maxiter=20000, solver="matfree", reltol=1e-14):
    """x -> filter -> solve -> compliance/sens -> filter pullback -> OC.  Returns (x, history)."""
    F = {"density": DensityFilter, "sens": SensFilter}[filt](prob, rmin)
    x = np.full(prob.nel, volfrac)
    dv = prob.cellvolumes / prob.cellvolumes.sum()
    hist = []
    solve = solve_matfree if solver == "matfree" else solve_assembled
    for _ in range(iters):
        xf = F(x)
        E = get_rho(xf, p, xmin)
        if prob.physics == "heat":
            obj, _, g, _, _ = thermal_compliance(prob, xf, p, xmin, abstol=abstol, maxiter=maxiter, reltol=reltol)
        else:
            u, _, _ = solve(prob, E, abstol=abstol, maxiter=maxiter, reltol=reltol)
            obj, _, g = compliance(prob, u, xf, p, xmin)
        dc = F.pullback(g)
        dvf = F.pullback(dv) if filt == "density" else dv
        hist.append(obj)
        x = oc_update(x, dc, dvf, volfrac)
    return x, hist
