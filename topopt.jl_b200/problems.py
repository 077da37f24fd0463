"""Host-side mirror of the reference's structured-grid problem types
(src/TopOptProblems/problem_types.jl:162-231 PointLoadCantilever, :334-402 HalfMBB,
:926-1088 HeatConductionProblem / HeatTree).  Setup only: numbering, node sets, loads and the
element matrix all come from libtopopt_cuda's host entry points; nothing here touches a GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _sizes(dim, ncomp, nels):
    lib = _lib.load()
    out = [C.c_int64() for _ in range(4)]
    _lib.check(lib.topopt_sizes(dim, ncomp, _lib.nels3(nels), *[C.byref(o) for o in out]))
    return tuple(int(o.value) for o in out)  # nnodes, nel, ndof, nnz


class Metadata:
    """metadata.node_dofs / cell_dofs (src/TopOptProblems/metadata.jl:19-35), 1-based, lazily built."""

    def __init__(self, dim, ncomp, nels):
        self.dim, self.ncomp, self.nels = dim, ncomp, tuple(nels)
        self.nnodes, self.nel, self.ndof, self.nnz = _sizes(dim, ncomp, nels)
        self._node_dofs = self._cell_dofs = self._cells = None

    @property
    def node_dofs(self):
        if self._node_dofs is None:
            out = np.empty((self.nnodes, self.ncomp), dtype=np.int64)
            _lib.check(_lib.load().topopt_node_dofs(self.dim, self.ncomp, _lib.nels3(self.nels), out.ctypes.data_as(_lib.c_ip)))
            self._node_dofs = out.T  # ncomp x nnodes view, like the reference's Matrix
        return self._node_dofs

    @property
    def cell_dofs(self):
        if self._cell_dofs is None:
            ks = self.ncomp * 2**self.dim
            out = np.empty((self.nel, ks), dtype=np.int64)
            _lib.check(_lib.load().topopt_cell_dofs(self.dim, self.ncomp, _lib.nels3(self.nels), out.ctypes.data_as(_lib.c_ip)))
            self._cell_dofs = out.T
        return self._cell_dofs

    @property
    def cells(self):
        if self._cells is None:
            out = np.empty((self.nel, 2**self.dim), dtype=np.int64)
            _lib.check(_lib.load().topopt_cells(self.dim, _lib.nels3(self.nels), out.ctypes.data_as(_lib.c_ip)))
            self._cells = out
        return self._cells

    def csc_pattern(self):
        colptr = np.empty(self.ndof + 1, dtype=np.int64)
        rowval = np.empty(self.nnz, dtype=np.int64)
        _lib.check(
            _lib.load().topopt_csc_pattern(
                self.dim, self.ncomp, _lib.nels3(self.nels), colptr.ctypes.data_as(_lib.c_ip), rowval.ctypes.data_as(_lib.c_ip)
            )
        )
        return colptr, rowval


def element_matrix(dim, physics, sizes, a, b=0.0, quad_order=4):
    """Kes[1] of the reference (matrices_and_vectors.jl:63-177 / :413-496), Gauss order 4 by default
    (solvers_api.jl:432-444)."""
    ncomp = 1 if physics == _lib.PHYSICS_HEAT else dim
    ks = ncomp * 2**dim
    Ke = np.empty((ks, ks), dtype=np.float64)
    sz = (C.c_double * 3)(*([float(s) for s in sizes] + [1.0] * (3 - len(sizes))))
    _lib.check(_lib.load().topopt_element_matrix(dim, physics, sz, float(a), float(b), int(quad_order), Ke.ctypes.data_as(_lib.c_dp)))
    return Ke  # symmetric, so row/column-major agree


class StructuredProblem:
    """Common data of a problem on a RectilinearGrid (src/TopOptProblems/grids.jl:66-106)."""

    physics = _lib.PHYSICS_ELASTICITY

    def __init__(self, nels, sizes, ncomp):
        self.nels = tuple(int(n) for n in nels)
        self.dim = len(self.nels)
        if self.dim not in (2, 3):
            raise ValueError("nels must have 2 or 3 entries")
        self.sizes = tuple(float(s) for s in (sizes if sizes is not None else (1.0,) * self.dim))
        self.ncomp = ncomp
        self.metadata = Metadata(self.dim, ncomp, self.nels)
        self.nnodes, self.nel, self.ndof = self.metadata.nnodes, self.metadata.nel, self.metadata.ndof
        self.prescribed_dofs = np.zeros(0, dtype=np.int64)  # 1-based, sorted (ch.prescribed_dofs)
        self.fixedload = np.zeros(self.ndof)
        self.cellvolumes = np.full(self.nel, float(np.prod(self.sizes)))
        self.Ke = None

    # lexicographic node ids (0-based) of the node sets used by the reference's predicates
    def _node_index_grids(self):
        nn = [n + 1 for n in self.nels]
        idx = np.arange(self.nnodes)
        i = idx % nn[0]
        j = (idx // nn[0]) % nn[1]
        k = idx // (nn[0] * nn[1]) if self.dim == 3 else np.zeros_like(idx)
        return i, j, k

    def nodeset(self, name):
        i, j, k = self._node_index_grids()
        nx, ny = self.nels[0], self.nels[1]
        if name == "left":
            m = i == 0
        elif name == "right":
            m = i == nx
        elif name == "bottom":
            m = j == 0
        elif name == "top":
            m = j == ny
        elif name == "middley":
            m = (2 * j == ny) if ny % 2 == 0 else np.zeros_like(i, dtype=bool)
        else:
            raise KeyError(name)
        return m


class PointLoadCantilever(StructuredProblem):
    """problem_types.jl:162-231: left face fully fixed, point load -force in y at the first node
    with x = xmax, y = ymid."""

    def __init__(self, nels, sizes=None, E=1.0, nu=0.3, force=1.0):
        super().__init__(nels, sizes, len(nels))
        if self.nels[1] % 2 or (self.dim == 3 and self.nels[2] % 2):
            raise ValueError("Grid does not have an even number of elements along the y and/or z axes.")
        self.E, self.nu, self.force = float(E), float(nu), float(force)
        nd = self.metadata.node_dofs
        self.prescribed_dofs = np.sort(nd[:, np.nonzero(self.nodeset("left"))[0]].ravel())
        fnode = np.nonzero(self.nodeset("right") & self.nodeset("middley"))[0][0]
        self.force_dof = int(nd[1, fnode])
        self.fixedload[self.force_dof - 1] = -self.force
        self.Ke = element_matrix(self.dim, _lib.PHYSICS_ELASTICITY, self.sizes, self.E, self.nu)


class HalfMBB(StructuredProblem):
    """problem_types.jl:334-402: left nodes fixed in x, bottom-right node fixed in y, load at top-left."""

    def __init__(self, nels, sizes=None, E=1.0, nu=0.3, force=1.0):
        super().__init__(nels, sizes, len(nels))
        self.E, self.nu, self.force = float(E), float(nu), float(force)
        nd = self.metadata.node_dofs
        u1 = nd[0, np.nonzero(self.nodeset("left"))[0]]
        u2 = nd[1, np.nonzero(self.nodeset("bottom") & self.nodeset("right"))[0]]
        self.prescribed_dofs = np.sort(np.concatenate([u1, u2]))
        fnode = np.nonzero(self.nodeset("top") & self.nodeset("left"))[0][0]
        self.force_dof = int(nd[1, fnode])
        self.fixedload[self.force_dof - 1] = -self.force
        self.Ke = element_matrix(self.dim, _lib.PHYSICS_ELASTICITY, self.sizes, self.E, self.nu)


class HeatConductionProblem(StructuredProblem):
    """problem_types.jl:926-1052 (2-D): scalar temperature field, homogeneous Dirichlet sides,
    boundary heat flux per facet set, point sources.  Non-zero prescribed temperatures are rejected
    exactly like CGMatrixFreeSolver does (solvers_api.jl:515-525)."""

    physics = _lib.PHYSICS_HEAT

    def __init__(self, nels, sizes=None, k=1.0, Tleft=0.0, Tright=0.0, Ttop=None, Tbottom=None, heatflux=None, cload=None):
        super().__init__(nels, sizes, 1)
        if self.dim != 2:
            raise ValueError("HeatConductionProblem: only 2-D grids are supported")
        self.k = float(k)
        nd = self.metadata.node_dofs
        pres = []
        for name, val in (("left", Tleft), ("right", Tright), ("top", Ttop), ("bottom", Tbottom)):
            if val is None:
                continue
            if val != 0:
                raise ValueError(
                    "CGMatrixFreeSolver does not yet support inhomogeneous Dirichlet BCs (nonzero prescribed values)"
                )
            pres.append(nd[0, np.nonzero(self.nodeset(name))[0]])
        self.prescribed_dofs = np.unique(np.concatenate(pres)) if pres else np.zeros(0, dtype=np.int64)
        # consistent nodal flux load: q*h/2 to both nodes of every boundary facet
        # (matrices_and_vectors.jl:288-339, assemble.jl:145-160)
        nx, ny = self.nels
        W = nx + 1
        for name, q in (heatflux or {}).items():
            if name == "top":
                a = np.arange(nx) + W * ny
                b, h = a + 1, self.sizes[0]
            elif name == "bottom":
                a = np.arange(nx)
                b, h = a + 1, self.sizes[0]
            elif name == "left":
                a = W * np.arange(ny)
                b, h = a + W, self.sizes[1]
            elif name == "right":
                a = W * np.arange(ny) + nx
                b, h = a + W, self.sizes[1]
            else:
                raise KeyError(name)
            np.add.at(self.fixedload, nd[0, a] - 1, q * h / 2.0)
            np.add.at(self.fixedload, nd[0, b] - 1, q * h / 2.0)
        for node, val in (cload or {}).items():  # node ids are 1-based like the reference's dict keys
            self.fixedload[nd[0, node - 1] - 1] += val
        self.Ke = element_matrix(self.dim, _lib.PHYSICS_HEAT, self.sizes, self.k)


def HeatTree(nels, sizes=None, k=1.0, q=1.0):
    """problem_types.jl:1074-1088: flux q through the top edge, bottom edge held at T = 0."""
    return HeatConductionProblem(nels, sizes, k, Tleft=None, Tright=None, Ttop=None, Tbottom=0.0, heatflux={"top": q})
