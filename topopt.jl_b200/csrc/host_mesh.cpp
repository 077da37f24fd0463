// Host-side structured-grid bookkeeping for libtopopt_cuda: Ferrite-compatible numbering,
// connectivity, CSC sparsity pattern and element matrices.  Pure C++ (no CUDA calls), so these
// entry points work on a machine without a GPU.
//
// What this replaces in the reference (JuliaTopOpt/TopOpt.jl v0.14.0):
//   grid.cells                         Ferrite.generate_grid via src/TopOptProblems/grids.jl:80-92
//   metadata.node_dofs / cell_dofs     src/TopOptProblems/metadata.jl:40-50,116-145
//   allocate_matrix(dh) pattern        src/TopOptProblems/matrices_and_vectors.jl:57
//   Kes[1]                             src/TopOptProblems/matrices_and_vectors.jl:63-177,413-496
#include "common.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace topopt {

thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

bool check_dims(int32_t dim, int32_t ncomp, const int64_t* nels) {
  if (dim != 2 && dim != 3) {
    set_error("dim must be 2 or 3");
    return false;
  }
  if (ncomp != 1 && ncomp != dim) {
    set_error("ncomp must be 1 (scalar field) or dim (displacement field)");
    return false;
  }
  if (!nels) {
    set_error("nels is NULL");
    return false;
  }
  for (int d = 0; d < dim; ++d)
    if (nels[d] < 1 || nels[d] > (int64_t(1) << 20)) {
      set_error("nels entries must be in [1, 2^20]");
      return false;
    }
  return true;
}

GridDims make_dims(int32_t dim, const int64_t* nels) {
  GridDims g;
  g.dim = dim;
  g.nx = nels[0];
  g.ny = nels[1];
  g.nz = dim == 3 ? nels[2] : 1;
  g.NX = g.nx + 1;
  g.NY = g.ny + 1;
  g.NZ = dim == 3 ? g.nz + 1 : 1;
  g.nnodes = g.NX * g.NY * g.NZ;
  g.nel = g.nx * g.ny * g.nz;
  return g;
}

// Local corner -> Ferrite local node number.  Quadrilateral: (0,0),(1,0),(1,1),(0,1);
// Hexahedron: that ring at z=0 then at z=1.
static inline int corner_to_local(int ox, int oy, int oz) { return ((ox ^ oy) | (oy << 1)) + 4 * oz; }

void cell_nodes(const GridDims& g, int64_t e, int64_t* out) {
  const int64_t i = e % g.nx, j = (e / g.nx) % g.ny, k = e / (g.nx * g.ny);
  const int nz_corners = g.dim == 3 ? 2 : 1;
  for (int oz = 0; oz < nz_corners; ++oz)
    for (int oy = 0; oy < 2; ++oy)
      for (int ox = 0; ox < 2; ++ox)
        out[corner_to_local(ox, oy, oz)] = (i + ox) + g.NX * ((j + oy) + g.NY * (k + oz));
}

// Ferrite close!(dh): visit cells in ascending order and, inside a cell, local nodes in ascending
// order; a node receives the next block of ncomp consecutive dofs the first time it is met.
std::vector<int64_t> ferrite_node_blocks(const GridDims& g) {
  std::vector<int64_t> block(g.nnodes, -1);
  int64_t next = 0;
  const int nen = g.dim == 3 ? 8 : 4;
  int64_t nodes[8];
  for (int64_t e = 0; e < g.nel; ++e) {
    cell_nodes(g, e, nodes);
    for (int a = 0; a < nen; ++a)
      if (block[nodes[a]] < 0) block[nodes[a]] = next++;
  }
  return block;
}

// number of structural non-zeros of allocate_matrix(dh)
int64_t pattern_nnz(const GridDims& g, int ncomp) {
  auto span = [](int64_t N) { return 3 * N - 2; };  // sum over nodes of in-range neighbours along an axis
  int64_t pairs = span(g.NX) * span(g.NY) * (g.dim == 3 ? span(g.NZ) : 1);
  return pairs * ncomp * ncomp;
}

// Gauss-Legendre nodes/weights on [-1,1] by Newton iteration on P_n.
static void gauss_legendre(int n, std::vector<double>& x, std::vector<double>& w) {
  x.resize(n);
  w.resize(n);
  for (int i = 0; i < n; ++i) {
    double z = std::cos(M_PI * (i + 0.75) / (n + 0.5));
    double pp = 0;
    for (int it = 0; it < 100; ++it) {
      double p1 = 1.0, p2 = 0.0;
      for (int j = 0; j < n; ++j) {
        double p3 = p2;
        p2 = p1;
        p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1.0);
      }
      pp = n * (z * p1 - p2) / (z * z - 1.0);
      double dz = p1 / pp;
      z -= dz;
      if (std::fabs(dz) < 1e-16) break;
    }
    x[n - 1 - i] = z;
    w[n - 1 - i] = 2.0 / ((1.0 - z * z) * pp * pp);
  }
}

}  // namespace topopt

using namespace topopt;

extern "C" {

const char* topopt_version(void) { return "libtopopt_cuda 0.1.0 (sm_100a)"; }

int topopt_sizes(int32_t dim, int32_t ncomp, const int64_t* nels, int64_t* nnodes, int64_t* nel,
                 int64_t* ndof, int64_t* nnz) {
  if (!check_dims(dim, ncomp, nels)) return TOPOPT_ERR_INVALID;
  GridDims g = make_dims(dim, nels);
  if (nnodes) *nnodes = g.nnodes;
  if (nel) *nel = g.nel;
  if (ndof) *ndof = g.nnodes * ncomp;
  if (nnz) *nnz = pattern_nnz(g, ncomp);
  return TOPOPT_OK;
}

int topopt_cells(int32_t dim, const int64_t* nels, int64_t* cells) {
  if (!check_dims(dim, 1, nels) || !cells) {
    if (!cells) set_error("cells is NULL");
    return TOPOPT_ERR_INVALID;
  }
  GridDims g = make_dims(dim, nels);
  const int nen = dim == 3 ? 8 : 4;
  int64_t nodes[8];
  for (int64_t e = 0; e < g.nel; ++e) {
    cell_nodes(g, e, nodes);
    for (int a = 0; a < nen; ++a) cells[e * nen + a] = nodes[a] + 1;
  }
  return TOPOPT_OK;
}

int topopt_node_dofs(int32_t dim, int32_t ncomp, const int64_t* nels, int64_t* node_dofs) {
  if (!check_dims(dim, ncomp, nels) || !node_dofs) {
    if (!node_dofs) set_error("node_dofs is NULL");
    return TOPOPT_ERR_INVALID;
  }
  GridDims g = make_dims(dim, nels);
  std::vector<int64_t> block = ferrite_node_blocks(g);
  for (int64_t n = 0; n < g.nnodes; ++n)
    for (int c = 0; c < ncomp; ++c) node_dofs[n * ncomp + c] = block[n] * ncomp + c + 1;
  return TOPOPT_OK;
}

int topopt_cell_dofs(int32_t dim, int32_t ncomp, const int64_t* nels, int64_t* cell_dofs) {
  if (!check_dims(dim, ncomp, nels) || !cell_dofs) {
    if (!cell_dofs) set_error("cell_dofs is NULL");
    return TOPOPT_ERR_INVALID;
  }
  GridDims g = make_dims(dim, nels);
  std::vector<int64_t> block = ferrite_node_blocks(g);
  const int nen = dim == 3 ? 8 : 4;
  const int ks = nen * ncomp;
  int64_t nodes[8];
  for (int64_t e = 0; e < g.nel; ++e) {
    cell_nodes(g, e, nodes);
    for (int a = 0; a < nen; ++a)
      for (int c = 0; c < ncomp; ++c) cell_dofs[e * ks + a * ncomp + c] = block[nodes[a]] * ncomp + c + 1;
  }
  return TOPOPT_OK;
}

int topopt_csc_pattern(int32_t dim, int32_t ncomp, const int64_t* nels, int64_t* colptr,
                       int64_t* rowval) {
  if (!check_dims(dim, ncomp, nels) || !colptr || !rowval) {
    if (!colptr || !rowval) set_error("colptr/rowval is NULL");
    return TOPOPT_ERR_INVALID;
  }
  GridDims g = make_dims(dim, nels);
  std::vector<int64_t> block = ferrite_node_blocks(g);
  std::vector<int64_t> node_of_block(g.nnodes);
  for (int64_t n = 0; n < g.nnodes; ++n) node_of_block[block[n]] = n;
  int64_t pos = 0;
  std::vector<int64_t> rows;
  rows.reserve(81);
  colptr[0] = 1;
  for (int64_t b = 0; b < g.nnodes; ++b) {
    const int64_t n = node_of_block[b];
    const int64_t i = n % g.NX, j = (n / g.NX) % g.NY, k = n / (g.NX * g.NY);
    rows.clear();
    for (int64_t dk = (g.dim == 3 ? -1 : 0); dk <= (g.dim == 3 ? 1 : 0); ++dk)
      for (int64_t dj = -1; dj <= 1; ++dj)
        for (int64_t di = -1; di <= 1; ++di) {
          const int64_t ii = i + di, jj = j + dj, kk = k + dk;
          if (ii < 0 || ii >= g.NX || jj < 0 || jj >= g.NY || kk < 0 || kk >= g.NZ) continue;
          const int64_t m = ii + g.NX * (jj + g.NY * kk);
          for (int c = 0; c < ncomp; ++c) rows.push_back(block[m] * ncomp + c + 1);
        }
    std::sort(rows.begin(), rows.end());
    for (int c = 0; c < ncomp; ++c) {
      std::memcpy(rowval + pos, rows.data(), rows.size() * sizeof(int64_t));
      pos += (int64_t)rows.size();
      colptr[b * ncomp + c + 1] = pos + 1;
    }
  }
  return TOPOPT_OK;
}

int topopt_element_matrix(int32_t dim, int32_t physics, const double* sizes, double a, double b,
                          int32_t quad_order, double* Ke) {
  if ((dim != 2 && dim != 3) || !sizes || !Ke || quad_order < 1 || quad_order > 16) {
    set_error("topopt_element_matrix: invalid argument");
    return TOPOPT_ERR_INVALID;
  }
  if (physics != TOPOPT_PHYSICS_ELASTICITY && physics != TOPOPT_PHYSICS_HEAT) {
    set_error("topopt_element_matrix: unknown physics");
    return TOPOPT_ERR_INVALID;
  }
  const int nen = dim == 3 ? 8 : 4;
  const int ncomp = physics == TOPOPT_PHYSICS_HEAT ? 1 : dim;
  const int ks = nen * ncomp;
  std::vector<double> gx, gw;
  gauss_legendre(quad_order, gx, gw);
  // reference corner signs in Ferrite local order
  double sgn[8][3];
  for (int oz = 0; oz < (dim == 3 ? 2 : 1); ++oz)
    for (int oy = 0; oy < 2; ++oy)
      for (int ox = 0; ox < 2; ++ox) {
        const int l = corner_to_local(ox, oy, oz);
        sgn[l][0] = ox ? 1.0 : -1.0;
        sgn[l][1] = oy ? 1.0 : -1.0;
        sgn[l][2] = oz ? 1.0 : -1.0;
      }
  double detJ = 1.0;
  for (int d = 0; d < dim; ++d) detJ *= sizes[d] / 2.0;
  const double lam = a * b / ((1.0 + b) * (1.0 - 2.0 * b));
  const double mu = a / (2.0 * (1.0 + b));
  std::vector<double> K(ks * ks, 0.0);
  const int nq = quad_order;
  for (int qz = 0; qz < (dim == 3 ? nq : 1); ++qz)
    for (int qy = 0; qy < nq; ++qy)
      for (int qx = 0; qx < nq; ++qx) {
        const double xi[3] = {gx[qx], gx[qy], dim == 3 ? gx[qz] : 0.0};
        const double dO = detJ * gw[qx] * gw[qy] * (dim == 3 ? gw[qz] : 1.0);
        double grad[8][3];
        for (int n = 0; n < nen; ++n)
          for (int d = 0; d < dim; ++d) {
            double v = sgn[n][d] / 2.0;
            for (int o = 0; o < dim; ++o)
              if (o != d) v *= (1.0 + sgn[n][o] * xi[o]) / 2.0;
            grad[n][d] = v / (sizes[d] / 2.0);
          }
        for (int nb = 0; nb < nen; ++nb)
          for (int na = 0; na < nen; ++na) {
            double dotab = 0.0;
            for (int d = 0; d < dim; ++d) dotab += grad[na][d] * grad[nb][d];
            if (physics == TOPOPT_PHYSICS_HEAT) {
              K[na + ks * nb] += a * dotab * dO;
            } else {
              for (int d2 = 0; d2 < dim; ++d2)
                for (int d1 = 0; d1 < dim; ++d1) {
                  double v = lam * grad[na][d1] * grad[nb][d2] + mu * grad[nb][d1] * grad[na][d2];
                  if (d1 == d2) v += mu * dotab;
                  K[(dim * na + d1) + ks * (dim * nb + d2)] += v * dO;
                }
            }
          }
      }
  // Symmetric(Ke_0): the upper triangle defines the matrix
  for (int c = 0; c < ks; ++c)
    for (int r = 0; r < ks; ++r) Ke[r + ks * c] = r <= c ? K[r + ks * c] : K[c + ks * r];
  return TOPOPT_OK;
}

}  // extern "C"
