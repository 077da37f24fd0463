"""save_mesh: VTK (.vtu) export of a design, mirroring src/TopOptProblems/IO/VTK.jl:15-62,70-110 -- the cells whose
density is >= 0.5, all node coordinates, and optionally a nodal field (temperature or displacement).  The design and the
solution can come straight from the device-resident buffers of a solver (topopt_get_design / the last solve)."""
from __future__ import annotations

import base64
import os

import numpy as np

from . import _lib

_VTK_CELL = {2: 9, 3: 12}  # VTK_QUAD, VTK_HEXAHEDRON: same local node order as Ferrite's Quadrilateral / Hexahedron


def _b64(a):
    raw = np.ascontiguousarray(a).tobytes()
    return base64.b64encode(np.uint32(len(raw)).tobytes() + raw).decode("ascii")


def device_design(solver):
    """The design vector as it stands in the solver's HBM buffer (after topopt_simp_eval / topopt_oc_update)."""
    out = np.empty(solver.problem.nel)
    solver._check(solver._lib.topopt_get_design(solver.handle, _lib.ptr(out)))
    return out


def save_mesh(filename, problem, vars=None, nodal=None, solver=None, nodal_name=None):
    """save_mesh(filename, problem[, vars][, temperature]) (VTK.jl:15-62).

    vars   : density vector of length nel; None = all ones (VTK.jl:15-19) or, with ``solver``, the device-resident design
    nodal  : optional nodal field in Ferrite dof order (ndof values: temperature for heat problems, displacement
             components for elasticity), written as point data; with ``solver`` and nodal=True the solver's solution u
    Returns the list of files written, like WriteVTK.vtk_save."""
    if hasattr(vars, "vars") and solver is None:  # save_mesh(filename, problem, solver) (VTK.jl:20-22)
        solver, vars = vars, None
    md = problem.metadata
    nel, nn, dim = problem.nel, problem.nnodes, problem.dim
    if vars is None:
        rho = device_design(solver) if (solver is not None and getattr(solver, "_solved_once", False)) else (
            np.asarray(solver.vars, dtype=np.float64) if solver is not None else np.ones(nel))
    else:
        rho = np.asarray(vars, dtype=np.float64)
    if rho.shape != (nel,):
        raise ValueError(f"Density vector ρ must have length equal to number of cells ({nel})")  # ArgumentError, VTK.jl:44-50
    if nodal is True:
        nodal = solver.u
    keep = np.nonzero(rho >= 0.5)[0]
    cells = md.cells[keep] - 1  # 0-based node ids, Ferrite local order
    i, j, k = problem._node_index_grids()
    pts = np.zeros((nn, 3))
    pts[:, 0] = i * problem.sizes[0]
    pts[:, 1] = j * problem.sizes[1]
    if dim == 3:
        pts[:, 2] = k * problem.sizes[2]
    nper = 2**dim
    if not filename.endswith(".vtu"):
        filename += ".vtu"
    point_data = ""
    if nodal is not None:
        nodal = np.asarray(nodal, dtype=np.float64)
        nc = md.ncomp
        if nodal.shape != (nn * nc,):
            raise ValueError(f"nodal field must have length ndof ({nn * nc})")
        field = nodal[md.node_dofs.T - 1]  # nnodes x ncomp, from Ferrite dof order
        if nc == 2:
            field = np.concatenate([field, np.zeros((nn, 1))], axis=1)  # VTK vectors have three components
        name = nodal_name or ("temperature" if nc == 1 else "displacement")
        point_data = (f'<PointData>\n<DataArray type="Float64" Name="{name}" NumberOfComponents="{field.shape[1]}" format="binary">'
                      f"{_b64(field)}</DataArray>\n</PointData>\n")
    xml = (
        '<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" version="0.1" byte_order="LittleEndian" header_type="UInt32">\n'
        f'<UnstructuredGrid>\n<Piece NumberOfPoints="{nn}" NumberOfCells="{len(keep)}">\n'
        f'<Points>\n<DataArray type="Float64" NumberOfComponents="3" format="binary">{_b64(pts)}</DataArray>\n</Points>\n'
        f'<Cells>\n<DataArray type="Int64" Name="connectivity" format="binary">{_b64(cells.astype(np.int64))}</DataArray>\n'
        f'<DataArray type="Int64" Name="offsets" format="binary">{_b64(np.arange(1, len(keep) + 1, dtype=np.int64) * nper)}</DataArray>\n'
        f'<DataArray type="UInt8" Name="types" format="binary">{_b64(np.full(len(keep), _VTK_CELL[dim], dtype=np.uint8))}</DataArray>\n</Cells>\n'
        f'<CellData>\n<DataArray type="Float64" Name="density" format="binary">{_b64(rho[keep])}</DataArray>\n</CellData>\n'
        f"{point_data}</Piece>\n</UnstructuredGrid>\n</VTKFile>\n"
    )
    with open(filename, "w") as f:
        f.write(xml)
    return [os.path.abspath(filename)]
