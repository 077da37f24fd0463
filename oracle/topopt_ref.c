/* CPU ORACLE (C port) -- TEST / BASELINE INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * A plain-C restatement of the TopOpt.jl CPU hot path with the reference's own data structures
 * and loop structure, used (a) by tests/ as a second checker next to oracle/topopt_oracle.py and
 * (b) by bench.py's cpu_baseline / --impl reference legs, where it is the thing timed on the
 * host cores.  Julia is not installed in this image, so the reference itself cannot run:
 * "parity unpinned" for floating point (see oracle/topopt_oracle.py header); integer tables are
 * pinned by the reference's golden tests through the numpy oracle, against which this file is
 * checked in tests/test_oracle_golden.py.
 *
 * Restated (file:line in JuliaTopOpt/TopOpt.jl v0.14.0):
 *   metadata.cell_dofs / dof_cells            src/TopOptProblems/metadata.jl:40-76
 *   ElementMatrix mask (bcmatrix)             src/TopOptProblems/elementmatrix.jl:33-41,64-109
 *   MatrixFreeOperator mul!                   src/FEA/matrix_free_operator.jl:66-105
 *   cg! (IterativeSolvers 0.9, CGIterable)    call site src/FEA/solvers_api.jl:196-219
 *   compute_element_energy                    src/Functions/compute_element_energy.jl:18-38
 *   sens-filter style two-stage filter        src/CheqFilters/sens_filter.jl:72-110, CheqFilters.jl:66-118
 * Deviation kept for memory: one shared Ke instead of one 24x24 copy per element (19 GB at
 * 256x128x128); the per-element BC mask is kept.  Single-threaded by default like the reference;
 * OpenMP (-fopenmp) parallelises the element and dof loops for the "all host cores" figure.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  int dim, ncomp, ks, nen;
  int64_t nx, ny, nz, NX, NY, NZ, nnodes, nel, ndof;
  double Ke[24 * 24];
  int64_t* cell_dofs;        /* ks x nel, 0-based */
  int64_t* dof_off;          /* ndof + 1 */
  int64_t* dof_cell;         /* (cell, local) pairs in ascending cell order */
  int32_t* dof_local;
  unsigned char* mask;       /* ks x nel, 1 = free */
  unsigned char* fixed;      /* ndof */
  double* xes;               /* ks x nel scratch */
  double* E;                 /* nel */
  double meandiag;
  double *r, *u, *c;         /* CG state */
} ref_t;

static int corner_local(int dim, int ox, int oy, int oz) { return dim == 3 ? (((ox ^ oy) | (oy << 1)) + 4 * oz) : ((ox ^ oy) | (oy << 1)); }

int ref_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

ref_t* ref_create(int dim, int ncomp, const int64_t* nels, const double* Ke, const int64_t* prescribed1, int64_t npres) {
  ref_t* R = (ref_t*)calloc(1, sizeof(ref_t));
  R->dim = dim; R->ncomp = ncomp; R->nen = dim == 3 ? 8 : 4; R->ks = R->nen * ncomp;
  R->nx = nels[0]; R->ny = nels[1]; R->nz = dim == 3 ? nels[2] : 1;
  R->NX = R->nx + 1; R->NY = R->ny + 1; R->NZ = dim == 3 ? R->nz + 1 : 1;
  R->nnodes = R->NX * R->NY * R->NZ; R->nel = R->nx * R->ny * R->nz; R->ndof = R->nnodes * ncomp;
  memcpy(R->Ke, Ke, sizeof(double) * R->ks * R->ks);
  const int ks = R->ks;
  /* Ferrite numbering: first visit in cell order */
  int64_t* block = (int64_t*)malloc(sizeof(int64_t) * R->nnodes);
  for (int64_t n = 0; n < R->nnodes; ++n) block[n] = -1;
  R->cell_dofs = (int64_t*)malloc(sizeof(int64_t) * ks * R->nel);
  int64_t next = 0;
  for (int64_t e = 0; e < R->nel; ++e) {
    int64_t i = e % R->nx, j = (e / R->nx) % R->ny, k = e / (R->nx * R->ny);
    int64_t nodes[8];
    for (int oz = 0; oz < (dim == 3 ? 2 : 1); ++oz)
      for (int oy = 0; oy < 2; ++oy)
        for (int ox = 0; ox < 2; ++ox) nodes[corner_local(dim, ox, oy, oz)] = (i + ox) + R->NX * ((j + oy) + R->NY * (k + oz));
    for (int a = 0; a < R->nen; ++a) {
      if (block[nodes[a]] < 0) block[nodes[a]] = next++;
      for (int c = 0; c < ncomp; ++c) R->cell_dofs[e * ks + a * ncomp + c] = block[nodes[a]] * ncomp + c;
    }
  }
  free(block);
  /* dof_cells ragged array (push order = ascending cell, then local) */
  R->dof_off = (int64_t*)calloc(R->ndof + 1, sizeof(int64_t));
  for (int64_t t = 0; t < ks * R->nel; ++t) R->dof_off[R->cell_dofs[t] + 1]++;
  for (int64_t d = 0; d < R->ndof; ++d) R->dof_off[d + 1] += R->dof_off[d];
  R->dof_cell = (int64_t*)malloc(sizeof(int64_t) * ks * R->nel);
  R->dof_local = (int32_t*)malloc(sizeof(int32_t) * ks * R->nel);
  int64_t* fill = (int64_t*)malloc(sizeof(int64_t) * R->ndof);
  memcpy(fill, R->dof_off, sizeof(int64_t) * R->ndof);
  for (int64_t e = 0; e < R->nel; ++e)
    for (int l = 0; l < ks; ++l) {
      int64_t d = R->cell_dofs[e * ks + l];
      R->dof_cell[fill[d]] = e;
      R->dof_local[fill[d]] = l;
      fill[d]++;
    }
  free(fill);
  R->fixed = (unsigned char*)calloc(R->ndof, 1);
  for (int64_t p = 0; p < npres; ++p) R->fixed[prescribed1[p] - 1] = 1;
  R->mask = (unsigned char*)malloc((size_t)ks * R->nel);
  for (int64_t t = 0; t < ks * R->nel; ++t) R->mask[t] = !R->fixed[R->cell_dofs[t]];
  R->xes = (double*)malloc(sizeof(double) * ks * R->nel);
  R->E = (double*)malloc(sizeof(double) * R->nel);
  for (int64_t e = 0; e < R->nel; ++e) R->E[e] = 1.0;
  double tr = 0;
  for (int l = 0; l < ks; ++l) tr += Ke[l * (ks + 1)];
  R->meandiag = tr * (double)R->nel; /* solvers_api.jl:526-527 */
  R->r = (double*)malloc(sizeof(double) * R->ndof);
  R->u = (double*)malloc(sizeof(double) * R->ndof);
  R->c = (double*)malloc(sizeof(double) * R->ndof);
  return R;
}

void ref_destroy(ref_t* R) {
  if (!R) return;
  free(R->cell_dofs); free(R->dof_off); free(R->dof_cell); free(R->dof_local); free(R->mask); free(R->fixed);
  free(R->xes); free(R->E); free(R->r); free(R->u); free(R->c); free(R);
}

int64_t ref_ndof(const ref_t* R) { return R->ndof; }
int64_t ref_nel(const ref_t* R) { return R->nel; }
void ref_cell_dofs(const ref_t* R, int64_t* out1) { for (int64_t t = 0; t < R->ks * R->nel; ++t) out1[t] = R->cell_dofs[t] + 1; }

/* E_e = rho^p (1-xmin) + xmin  (penalties.jl:113-119) */
void ref_set_density(ref_t* R, const double* rho, double p, double xmin) {
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < R->nel; ++e) R->E[e] = pow(rho[e], p) * (1.0 - xmin) + xmin;
}

/* mul!(y, A::MatrixFreeOperator, x) */
void ref_mul(ref_t* R, const double* x, double* y) {
  const int ks = R->ks;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < R->nel; ++e) {
    double xe[24], out[24];
    const int64_t* cd = R->cell_dofs + e * ks;
    const unsigned char* m = R->mask + e * ks;
    for (int l = 0; l < ks; ++l) { xe[l] = x[cd[l]]; out[l] = 0.0; }
    for (int j = 0; j < ks; ++j) { /* bcmatrix(Ke) * xe, column-ordered */
      if (!m[j]) continue;
      const double xj = xe[j];
      const double* col = R->Ke + j * ks;
      for (int i = 0; i < ks; ++i) out[i] += col[i] * xj;
    }
    const double px = R->E[e];
    double* dst = R->xes + e * ks;
    for (int i = 0; i < ks; ++i) dst[i] = m[i] ? px * out[i] : 0.0;
  }
#pragma omp parallel for schedule(static)
  for (int64_t d = 0; d < R->ndof; ++d) {
    if (R->fixed[d]) { y[d] = R->meandiag * x[d]; continue; }
    double yi = 0.0;
    for (int64_t s = R->dof_off[d]; s < R->dof_off[d + 1]; ++s) yi += R->xes[R->dof_cell[s] * ks + R->dof_local[s]];
    y[d] = yi;
  }
}

static double dotp(const double* a, const double* b, int64_t n) {
  double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

/* IterativeSolvers cg! with x0 = 0 (cg_solve! zero-fills lhs), Identity preconditioner.
 * b must already be zero on prescribed dofs.  Returns the iteration count. */
int ref_cg(ref_t* R, const double* b, double* x, double abstol, double reltol, int maxiter, double* residual) {
  const int64_t n = R->ndof;
  double *r = R->r, *u = R->u, *c = R->c;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) { x[i] = 0.0; u[i] = 0.0; r[i] = b[i]; }
  double res = sqrt(dotp(r, r, n)), prev = 1.0;
  const double tol = fmax(reltol * res, abstol);
  int it = 0;
  while (it < maxiter && res > tol) {
    const double beta = res * res / (prev * prev);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) u[i] = r[i] + beta * u[i];
    ref_mul(R, u, c);
    const double alpha = res * res / dotp(u, c, n);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) { x[i] += alpha * u[i]; r[i] -= alpha * c[i]; }
    prev = res;
    res = sqrt(dotp(r, r, n));
    ++it;
  }
  if (residual) *residual = res;
  return it;
}

/* compute_element_energy: cell_e = u_e' Ke u_e (raw Ke), grad_e = -dE_e cell_e, returns sum E_e cell_e */
double ref_compliance(ref_t* R, const double* u, const double* rho, double p, double xmin, double* cell, double* grad) {
  const int ks = R->ks;
  double obj = 0.0;
#pragma omp parallel for reduction(+ : obj) schedule(static)
  for (int64_t e = 0; e < R->nel; ++e) {
    const int64_t* cd = R->cell_dofs + e * ks;
    double ce = 0.0;
    for (int w = 0; w < ks; ++w)
      for (int v = 0; v < ks; ++v) ce += u[cd[v]] * R->Ke[v + ks * w] * u[cd[w]];
    const double Ee = pow(rho[e], p) * (1.0 - xmin) + xmin;
    const double dEe = (1.0 - xmin) * p * pow(rho[e], p - 1.0);
    cell[e] = ce;
    grad[e] = -dEe * ce;
    obj += Ee * ce;
  }
  return obj;
}

/* two-stage filter, forward map (cell -> node volume-weighted mean -> cell, duplicate-weighted
 * cone with strict d < rmin); uniform cells. */
void ref_filter(ref_t* R, const double* sizes, double rmin, const double* x, double* y) {
  double* nodal = (double*)malloc(sizeof(double) * R->nnodes);
  const int dim = R->dim;
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < R->nnodes; ++n) {
    int64_t i = n % R->NX, j = (n / R->NX) % R->NY, k = n / (R->NX * R->NY);
    double s = 0.0, w = 0.0;
    for (int dk = (dim == 3 ? -1 : 0); dk <= 0; ++dk)
      for (int dj = -1; dj <= 0; ++dj)
        for (int di = -1; di <= 0; ++di) {
          int64_t ci = i + di, cj = j + dj, ck = k + dk;
          if (ci < 0 || ci >= R->nx || cj < 0 || cj >= R->ny || ck < 0 || ck >= R->nz) continue;
          s += x[ci + R->nx * (cj + R->ny * ck)];
          w += 1.0;
        }
    nodal[n] = s / w;
  }
  int Rr[3] = {0, 0, 0};
  for (int a = 0; a < dim; ++a) Rr[a] = (int)ceil(rmin / sizes[a] + 0.5 + 1e-9) - 1;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < R->nel; ++e) {
    int64_t i = e % R->nx, j = (e / R->nx) % R->ny, k = e / (R->nx * R->ny);
    double num = 0.0, den = 0.0;
    for (int oz = (dim == 3 ? 1 - Rr[2] : 0); oz <= (dim == 3 ? Rr[2] : 0); ++oz)
      for (int oy = 1 - Rr[1]; oy <= Rr[1]; ++oy)
        for (int ox = 1 - Rr[0]; ox <= Rr[0]; ++ox) {
          int64_t ni = i + ox, nj = j + oy, nk = k + oz;
          if (ni < 0 || ni >= R->NX || nj < 0 || nj >= R->NY || nk < 0 || nk >= R->NZ) continue;
          double dx = (ox - 0.5) * sizes[0], dy = (oy - 0.5) * sizes[1], dz = dim == 3 ? (oz - 0.5) * sizes[2] : 0.0;
          double dist = sqrt(dx * dx + dy * dy + dz * dz);
          if (!(dist < rmin)) continue;
          int m = ((ni > 0 && ni < R->NX - 1) ? 2 : 1) * ((nj > 0 && nj < R->NY - 1) ? 2 : 1) * ((dim == 3 && nk > 0 && nk < R->NZ - 1) ? 2 : 1);
          double w = m * (rmin - dist);
          num += w * nodal[ni + R->NX * (nj + R->NY * nk)];
          den += w;
        }
    y[e] = num / den;
  }
  free(nodal);
}
