/* Plain-C driver of the drop-in boundary (include/topopt_cuda.h), the way a ccall / cgo / JNI binding uses it:
 * host helpers -> create -> set_density -> solve -> compliance -> filter -> fused SIMP evaluation -> destroy,
 * once per visible device (at most two) to cover handles on different devices in one process.
 * Prints one line per device: "dev obj obj_fused grad_checksum filt_checksum iters quad_form".
 * Problem: PointLoadCantilever nx x ny x nz (problem_types.jl:162-223): face x=0 clamped, unit load -y at the
 * first node with x = xmax, y = ymid. */
#include "topopt_cuda.h"
#include <stdio.h>
#include <stdlib.h>

#define CHECK(call)                                                                        \
  do {                                                                                     \
    int rc_ = (call);                                                                      \
    if (rc_ != TOPOPT_OK) {                                                                \
      fprintf(stderr, "%s failed: %d %s\n", #call, rc_, topopt_last_error(NULL));          \
      return 10;                                                                           \
    }                                                                                      \
  } while (0)

int main(int argc, char** argv) {
  int64_t nels[3] = {12, 4, 28};
  int ndev = argc > 1 ? atoi(argv[1]) : 1;
  int64_t nn, ne, nd, nz;
  int dev;
  CHECK(topopt_sizes(3, 3, nels, &nn, &ne, &nd, &nz));
  {
    const int64_t NX = nels[0] + 1, NY = nels[1] + 1, NZ = nels[2] + 1;
    double sizes[3] = {1.0, 1.0, 1.0}, Ke[24 * 24];
    int64_t* node_dofs = (int64_t*)malloc(sizeof(int64_t) * 3 * (size_t)nn);
    int64_t* pres = (int64_t*)malloc(sizeof(int64_t) * 3 * (size_t)(NY * NZ));
    double* f = (double*)calloc((size_t)nd, sizeof(double));
    double* rho = (double*)malloc(sizeof(double) * (size_t)ne);
    double* u = (double*)malloc(sizeof(double) * (size_t)nd);
    double* grad = (double*)malloc(sizeof(double) * (size_t)ne);
    double* xf = (double*)malloc(sizeof(double) * (size_t)ne);
    double* gx = (double*)malloc(sizeof(double) * (size_t)ne);
    int64_t np = 0, j, k, e;
    CHECK(topopt_element_matrix(3, TOPOPT_PHYSICS_ELASTICITY, sizes, 1.0, 0.3, 4, Ke));
    CHECK(topopt_node_dofs(3, 3, nels, node_dofs));
    for (k = 0; k < NZ; ++k)
      for (j = 0; j < NY; ++j) {
        const int64_t n = NX * (j + NY * k); /* i = 0 */
        int c;
        for (c = 0; c < 3; ++c) pres[np++] = node_dofs[3 * n + c];
      }
    {
      const int64_t n = (NX - 1) + NX * (NY / 2); /* x = xmax, y = ymid, z = 0 */
      f[node_dofs[3 * n + 1] - 1] = -1.0;
    }
    for (e = 0; e < ne; ++e) rho[e] = 0.2 + 0.8 * (double)((e * 2654435761u) % 1000) / 1000.0;
    for (dev = 0; dev < ndev; ++dev) {
      topopt_desc d;
      topopt_handle* h = NULL;
      topopt_filter* flt = NULL;
      topopt_cg_opts o;
      topopt_cg_result r;
      double obj = 0, obj2 = 0, gs = 0, fs = 0, q = 0;
      d.dim = 3; d.ncomp = 3;
      d.nels[0] = nels[0]; d.nels[1] = nels[1]; d.nels[2] = nels[2];
      d.sizes[0] = d.sizes[1] = d.sizes[2] = 1.0;
      d.Ke = Ke; d.prescribed_dofs = pres; d.n_prescribed = np; d.fixedload = f;
      d.cellvolumes = NULL; d.cell_dofs = NULL; d.fixed_diag = 0.0;
      d.device = dev; d.rank = 0; d.world = 1; d.nccl_unique_id = NULL;
      CHECK(topopt_create(&d, &h));
      o.abstol = 1e-11; o.reltol = 1e-14; o.maxiter = 20000; o.op = TOPOPT_OP_MATRIX_FREE; o.precond = TOPOPT_PRECOND_NONE;
      o.criteria = TOPOPT_CRITERIA_DEFAULT; o.check_every = 0; o.variant = dev == 0 ? TOPOPT_CG_REFERENCE : TOPOPT_CG_SINGLE_PASS;
      o.warm_start = 0; o.refresh_precond = 0;
      CHECK(topopt_set_density(h, rho, TOPOPT_PENALTY_POWER, 3.0, 1e-3, 1));
      CHECK(topopt_solve(h, NULL, u, &o, &r));
      CHECK(topopt_compliance(h, NULL, &obj, NULL, grad));
      CHECK(topopt_quadratic_form(h, NULL, &q));
      for (e = 0; e < ne; ++e) gs += grad[e] * (double)(1 + e % 7);
      CHECK(topopt_filter_create(h, 2.0, &flt));
      CHECK(topopt_filter_apply(flt, rho, xf, TOPOPT_FILTER_FORWARD));
      for (e = 0; e < ne; ++e) fs += xf[e] * (double)(1 + e % 5);
      CHECK(topopt_simp_eval(h, flt, 1, rho, TOPOPT_PENALTY_POWER, 3.0, 1e-3, &o, &obj2, gx, &r));
      printf("%d %.15e %.15e %.15e %.15e %d %.15e\n", dev, obj, obj2, gs, fs, (int)r.iters, q);
      CHECK(topopt_filter_destroy(flt));
      CHECK(topopt_destroy(h));
    }
    free(node_dofs); free(pres); free(f); free(rho); free(u); free(grad); free(xf); free(gx);
  }
  return 0;
}
