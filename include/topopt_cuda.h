/*
 * libtopopt_cuda -- C ABI of the B200 (sm_100a) SIMP inner loop behind TopOpt.jl's solver API.
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes, and is what the Julia
 * glue binds with `ccall`.  Reference citations are file:line in
 * JuliaTopOpt/TopOpt.jl v0.14.0.
 *
 * Conventions
 *  - All vectors crossing the ABI are fp64 in the REFERENCE's ordering: dof vectors in Ferrite
 *    DOF order (length ndof), element vectors in Ferrite cell order (length nel).  Indices
 *    crossing the ABI (prescribed dofs, cell_dofs, CSC pattern) are 1-based Int64 like Julia's.
 *  - A pointer may be host memory (pageable or pinned) or device memory of the handle's GPU;
 *    copies use cudaMemcpyDefault.  A NULL *input* means "use the device-resident value from
 *    the previous call"; a NULL *output* means "leave the result resident on the device".
 *  - Return value: TOPOPT_OK (0) or a negative topopt_status; the message is available from
 *    topopt_last_error().  CG non-convergence is NOT an error (the reference is silent:
 *    src/FEA/solvers_api.jl:203); it is reported in topopt_cg_result.
 *  - One handle = one caller thread at a time (the reference is single-task and mutates the
 *    solver in place: src/Functions/compliance.jl:65-66).  Calls are synchronous on return.
 *  - Multi-GPU: one process per GPU (rank), structured grid split into slabs along the last
 *    axis; every rank passes the same full-length arrays and receives full-length results.
 */
#ifndef TOPOPT_CUDA_H
#define TOPOPT_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct topopt_handle topopt_handle;
typedef struct topopt_filter topopt_filter;

typedef enum {
  TOPOPT_OK = 0,
  TOPOPT_ERR_INVALID = -1,    /* ArgumentError on the Julia side */
  TOPOPT_ERR_CUDA = -2,
  TOPOPT_ERR_NCCL = -3,
  TOPOPT_ERR_NONFINITE = -4,  /* DomainError (src/FEA/convergence_criteria.jl:34-41) */
  TOPOPT_ERR_NO_DEVICE = -5,
  TOPOPT_ERR_MISMATCH = -6    /* caller's cell_dofs disagree with the internal numbering */
} topopt_status;

/* penalty kinds: src/Utilities/penalties.jl:30-54 */
enum { TOPOPT_PENALTY_POWER = 0, TOPOPT_PENALTY_RATIONAL = 1, TOPOPT_PENALTY_SINH = 2 };
/* operator used by topopt_solve: CGMatrixFreeSolver / CGAssemblySolver (solvers_api.jl:58,66) */
enum { TOPOPT_OP_MATRIX_FREE = 0, TOPOPT_OP_ASSEMBLED = 1 };
/* preconditioner: identity or DiagonalPreconditioner (Preconditioners.jl, solvers_api.jl:187-192) */
enum { TOPOPT_PRECOND_NONE = 0, TOPOPT_PRECOND_JACOBI = 1, TOPOPT_PRECOND_MULTIGRID = 2 };
/* CG scalar recurrence.  REFERENCE reproduces IterativeSolvers 0.9 cg! iterate by iterate
 * (three vector passes per iteration).  SINGLE_PASS is the same Krylov method with beta predicted
 * from |r - alpha Ap|^2 = alpha^2 Ap.Ap - r.r, which lets x, r and p be updated in one pass; its
 * iterates agree with the reference's to rounding (identity / default criteria only, otherwise
 * the library falls back to REFERENCE). */
enum { TOPOPT_CG_REFERENCE = 0, TOPOPT_CG_SINGLE_PASS = 1 };
/* convergence criteria: src/FEA/convergence_criteria.jl:13-45 */
enum { TOPOPT_CRITERIA_DEFAULT = 0, TOPOPT_CRITERIA_ENERGY = 1 };
/* physics for topopt_element_matrix */
enum { TOPOPT_PHYSICS_ELASTICITY = 0, TOPOPT_PHYSICS_HEAT = 1 };
/* filter application modes */
enum {
  TOPOPT_FILTER_FORWARD = 0,   /* J x          DensityFilterFun call, density_filter.jl:35-40;
                                  also the SensFilterFun pullback map, sens_filter.jl:52-70     */
  TOPOPT_FILTER_TRANSPOSE = 1  /* J' d         DensityFilterFun rrule, density_filter.jl:41-47 */
};

/* Problem description.  Replaces what GenericFEASolver holds on the CPU
 * (src/FEA/solvers_api.jl:86-123): elementinfo.Kes[1], metadata.cell_dofs,
 * ch.prescribed_dofs, elementinfo.fixedload, elementinfo.cellvolumes, meandiag. */
typedef struct {
  int32_t dim;                    /* 2 (quad4) or 3 (hex8)                                      */
  int32_t ncomp;                  /* dofs per node: dim (elasticity) or 1 (heat)                 */
  int64_t nels[3];                /* elements per axis (nels[2] ignored when dim == 2)           */
  double sizes[3];                /* element edge lengths (filter geometry, cell volumes)        */
  const double* Ke;               /* Kesize x Kesize, column-major like SMatrix; Kesize=ncomp*2^dim */
  const int64_t* prescribed_dofs; /* 1-based Ferrite dofs (ch.prescribed_dofs), any order        */
  int64_t n_prescribed;
  const double* fixedload;        /* ndof, Ferrite order (elementinfo.fixedload); may be NULL=0  */
  const double* cellvolumes;      /* nel or NULL (= prod(sizes)); must be uniform                */
  const int64_t* cell_dofs;       /* optional Kesize x nel 1-based cross-check of the numbering  */
  double fixed_diag;              /* y[d] = fixed_diag*x[d] on prescribed rows; <=0 selects the
                                     reference's sum_e tr(Ke) (solvers_api.jl:526-527)           */
  int32_t device;                 /* CUDA device ordinal                                         */
  int32_t rank;                   /* slab rank in [0, world)                                     */
  int32_t world;                  /* number of ranks (1 = single GPU)                            */
  const void* nccl_unique_id;     /* 128-byte ncclUniqueId shared by all ranks (world > 1)       */
} topopt_desc;

typedef struct {
  double abstol;          /* solver.abstol (default 1e-7, solvers_api.jl:489)                    */
  double reltol;          /* IterativeSolvers default sqrt(eps); tol = max(reltol*|r0|, abstol)  */
  int32_t maxiter;        /* solver.cg_max_iter (default 700, solvers_api.jl:478)                */
  int32_t op;             /* TOPOPT_OP_*                                                         */
  int32_t precond;        /* TOPOPT_PRECOND_*                                                    */
  int32_t criteria;       /* TOPOPT_CRITERIA_*                                                   */
  int32_t check_every;    /* host polls the device convergence flag every N iterations (0=auto) */
  int32_t variant;        /* TOPOPT_CG_*: scalar recurrence (0 = the reference's)                */
  int32_t warm_start;     /* 0 = zero initial guess like the reference (solvers_api.jl:203);
                             1 = start from the device-resident solution of the previous solve   */
  int32_t refresh_precond;/* 0 = preconditioner built once per solver like the reference
                             (solvers_api.jl:187-192); 1 = rebuild it from the current stiffness */
  int32_t mg_degree;      /* TOPOPT_PRECOND_MULTIGRID: Chebyshev smoother degree (0 = default 2)  */
  int32_t reserved0;
  double mg_ratio;        /* ... smoothed spectrum [lambda_max/ratio, lambda_max] (0 = default 6) */
} topopt_cg_opts;

typedef struct {
  int32_t iters;
  int32_t converged;
  double residual;        /* final ||r||_2                                                       */
  double tol;             /* tolerance that was applied                                          */
  double solve_ms;        /* device time of the CG loop (CUDA events)                            */
} topopt_cg_result;

typedef struct {
  int64_t ndof, nel, nnodes, nnz;
  int64_t ndof_local, nel_local;  /* owned by this rank                                          */
  int64_t kernel_launches;        /* launches of this library's kernels since create/reset       */
  int64_t cg_iterations;          /* total CG iterations since create/reset                      */
  int64_t h2d_bytes, d2h_bytes;   /* bytes copied through the ABI since create/reset             */
  double last_apply_ms, last_solve_ms, last_sens_ms, last_filter_ms;
} topopt_stats;

/* ---- library / host-only helpers (no GPU needed) -------------------------------------- */
const char* topopt_version(void);
/* message of the last failing call on this thread (h may be NULL) */
const char* topopt_last_error(const topopt_handle* h);

/* sizes of the structured problem */
int topopt_sizes(int32_t dim, int32_t ncomp, const int64_t* nels, int64_t* nnodes, int64_t* nel,
                 int64_t* ndof, int64_t* nnz);
/* Ferrite DofHandler numbering (src/TopOptProblems/metadata.jl:116-145): ncomp x nnodes, 1-based */
int topopt_node_dofs(int32_t dim, int32_t ncomp, const int64_t* nels, int64_t* node_dofs);
/* metadata.cell_dofs (metadata.jl:40-50): Kesize x nel column-major, 1-based */
int topopt_cell_dofs(int32_t dim, int32_t ncomp, const int64_t* nels, int64_t* cell_dofs);
/* grid.cells connectivity (Ferrite generate_grid): 2^dim x nel column-major, 1-based node ids */
int topopt_cells(int32_t dim, const int64_t* nels, int64_t* cells);
/* sparsity of allocate_matrix(dh) (src/TopOptProblems/matrices_and_vectors.jl:57):
 * CSC colptr (ndof+1) and rowval (nnz), 1-based, rows sorted; symmetric so also valid CSR */
int topopt_csc_pattern(int32_t dim, int32_t ncomp, const int64_t* nels, int64_t* colptr,
                       int64_t* rowval);
/* element matrix by Gauss quadrature (matrices_and_vectors.jl:63-177 elasticity, plane strain
 * in 2-D; :413-496 heat).  a = E (elasticity) or k (heat); b = nu (elasticity).  Column-major. */
int topopt_element_matrix(int32_t dim, int32_t physics, const double* sizes, double a, double b,
                          int32_t quad_order, double* Ke);

/* ncclGetUniqueId into a 128-byte buffer; rank 0 calls it and ships the bytes to the other
 * ranks (any transport), every rank then passes them in topopt_desc.nccl_unique_id. */
int topopt_nccl_unique_id(void* out128);

/* Peer-memory fast path for world > 1 (all ranks on one NVSwitch node, one process per GPU):
 * every rank exports a blob (cudaIpc handles of its communication block and its six CG vectors, 452 bytes;
 * pass a buffer of >= 1024), the host all-gathers the blobs (any transport) and every rank imports all of
 * them.  Afterwards the CG loop all-reduces its scalars inside its kernels and reads halo planes directly
 * from the neighbours' memory over NVLink; without this call NCCL is used.  The path is all-or-nothing: if
 * the import fails on ANY rank (no P2P), every rank must call topopt_ipc_import(h, NULL, 0), which unmaps
 * whatever was mapped and returns the handle to the NCCL path. */
int topopt_ipc_export(topopt_handle* h, void* out, int64_t* nbytes);
int topopt_ipc_import(topopt_handle* h, const void* all_blobs, int64_t nbytes_each);

/* ---- handle --------------------------------------------------------------------------- */
/* FEASolver(Solver, problem; ...) (src/FEA/solvers_api.jl:468-574, 599-603) */
int topopt_create(const topopt_desc* desc, topopt_handle** out);
int topopt_destroy(topopt_handle* h);
int topopt_get_stats(topopt_handle* h, topopt_stats* out);
int topopt_reset_stats(topopt_handle* h);

/* solver.vars .= x; penalised stiffness E_e and dE_e
 * (penalties.jl:113-130; density utils.jl:77).  order: 1 = penalty then interpolation (default
 * preference, LocalPreferences.toml:2), 0 = interpolation then penalty. */
int topopt_set_density(topopt_handle* h, const double* rho, int32_t penalty_kind, double p,
                       double xmin, int32_t penalty_before_interpolation);
/* ProjectedPenaltyFun (penalties.jl:62-69): the penalty of every following topopt_set_density /
 * topopt_simp_eval is applied to proj(x).  proj_kind 0 = none, 1 = HeavisideProjectionFun(beta)
 * `1 - exp(-beta x) + x exp(-beta)`, 2 = SigmoidProjectionFun(beta) (penalties.jl:77-96). */
int topopt_set_projection(topopt_handle* h, int32_t proj_kind, double beta);
/* direct control of E_e / dE_e (nel each; dE may be NULL) */
int topopt_set_stiffness(topopt_handle* h, const double* E, const double* dE);
/* read back E_e, dE_e (nel each; either may be NULL) */
int topopt_get_stiffness(topopt_handle* h, double* E, double* dE);

/* mul!(y, MatrixFreeOperator, x)  (src/FEA/matrix_free_operator.jl:66-105) */
int topopt_apply(topopt_handle* h, const double* x, double* y);

/* The same product through ONE kernel generation, with the dot products that kernel fuses for the
 * CG loop: kernel 0 = what topopt_apply selects, 1 = dense node-centric gather, 2 = hex8 modal
 * one-row, 3 = hex8 modal two-row, 4 = hex8 ring-staged (requires x == 0 on prescribed dofs, as
 * every CG direction is).  dots (2 doubles, may be NULL): x.y and, for kernel 4, y.y. */
int topopt_apply_ex(topopt_handle* h, const double* x, double* y, int32_t kernel, double* dots);

/* assemble!(globalinfo, ...) + apply!  (src/TopOptProblems/assemble.jl:28-90).  nzval (nnz, in
 * the order of topopt_csc_pattern) and f (ndof) may be NULL. */
int topopt_assemble(topopt_handle* h, double* nzval, double* f);
/* mul!(y, MatrixOperator, x)  (src/FEA/matrix_free_operator.jl:16) on the assembled matrix */
int topopt_spmv(topopt_handle* h, const double* x, double* y);

/* Jacobi preconditioner: diag = NULL computes diag K(rho) from the current stiffness
 * (UpdatePreconditioner!, solvers_api.jl:187-192), else uploads the caller's diagonal. */
int topopt_set_jacobi(topopt_handle* h, const double* diag);

/* solver() : cg_solve! with zero initial guess (solvers_api.jl:177-282, IterativeSolvers cg!).
 * rhs NULL = fixedload with prescribed entries zeroed (assemble.jl:51,87); a caller rhs gets
 * apply_zero! (solvers_api.jl:364-367).  u may be NULL (stays resident for topopt_compliance). */
int topopt_solve(topopt_handle* h, const double* rhs, double* u, const topopt_cg_opts* opts,
                 topopt_cg_result* result);

/* ComplianceFun call (src/Functions/compliance.jl:58-70, compute_element_energy.jl:18-38):
 * cell_comp_e = u_e' Ke u_e, grad_e = -dE_e cell_comp_e, obj = sum E_e cell_comp_e. */
int topopt_compliance(topopt_handle* h, const double* u, double* obj, double* cell_comp,
                      double* grad);
/* thermal/adjoint form (src/Functions/thermal_compliance.jl:144-155):
 * cell_e = lambda_e' Ke u_e, grad_e = dE_e cell_e. */
int topopt_bilinear_sens(topopt_handle* h, const double* lambda, const double* u,
                         double* cell_out, double* grad);
/* swap the device-resident solution with the device-resident lambda vector, so that a forward
 * solve, a swap and an adjoint solve (solve_adjoint!, thermal_compliance.jl:169-210) leave both
 * fields resident for topopt_bilinear_sens(h, NULL, NULL, ...). */
int topopt_swap_solution_lambda(topopt_handle* h);
/* dot(a, b) over dofs on the device (thermal J = dot(fixedload, u), thermal_compliance.jl:127;
 * getcompliance).  a NULL = fixedload, b NULL = resident u. */
int topopt_dot(topopt_handle* h, const double* a, const double* b, double* out);

/* FEA.getcompliance (src/FEA/FEA.jl:40): u' K u with the current stiffness, evaluated on the device.
 * u NULL = the device-resident solution. */
int topopt_quadratic_form(topopt_handle* h, const double* u, double* out);

/* ---- filters (src/CheqFilters) -------------------------------------------------------- */
/* FilterMetadata + getJacobian (CheqFilters.jl:66-118, density_filter.jl:49-108) without
 * materialising J: two-stage cell->node->cell stencil with the reference's duplicate weights. */
int topopt_filter_create(topopt_handle* h, double rmin, topopt_filter** out);
int topopt_filter_apply(topopt_filter* f, const double* x, double* y, int32_t mode);
int topopt_filter_destroy(topopt_filter* f);

/* ---- fused device-resident SIMP evaluation -------------------------------------------- */
/* x (nel, design) -> filter -> penalise -> solve -> compliance + sensitivity -> filter pullback.
 * filter_kind: 0 none, 1 density (forward J x, pullback J' g), 2 sensitivity (identity forward,
 * pullback J g).  x NULL = resident design.  grad_x (nel) may be NULL. */
int topopt_simp_eval(topopt_handle* h, topopt_filter* f, int32_t filter_kind, const double* x,
                     int32_t penalty_kind, double p, double xmin, const topopt_cg_opts* opts,
                     double* obj, double* grad_x, topopt_cg_result* result);

/* ---- on-device design update (device-resident SIMP loop) ---------------------------------- */
/* Optimality-criteria update with bisection on the volume multiplier: xn = clamp(x (-dc / (dv lambda))^eta) within
 * [max(xlo, x - move), min(1, x + move)], lambda such that sum xn dv = volfrac.  Every bisection step is one fused
 * kernel over the device-resident design; only 8 bytes per step cross PCIe.  x NULL = resident design (the one the last
 * topopt_simp_eval / topopt_oc_update left), dc NULL = the design gradient the last topopt_simp_eval left on the device,
 * dv NULL = resident volume gradient (must be passed once).  x_out (nel) may be NULL: the new design stays resident and
 * is what topopt_simp_eval(x = NULL) evaluates next.  change_l2 = ||xn - x||_2.  Replaces the host-side update that the
 * reference delegates to its optimiser packages (cf. the threshold bisection of src/Algorithms/beso.jl:154-173). */
int topopt_oc_update(topopt_handle* h, const double* x, const double* dc, const double* dv, double volfrac,
                     double move, double eta, double xlo, double* x_out, double* change_l2,
                     int32_t* nbisect);

/* the device-resident design vector (nel) */
int topopt_get_design(topopt_handle* h, double* x);

/* ---- measurement ---------------------------------------------------------------------- */
/* time `reps` back-to-back launches of one kernel class with CUDA events on the library's
 * stream; which: 0 = K.u as the CG loop launches it (dense direction vector, fused p.Ap),
 * 1 = CG iteration (matrix-free, reference recurrence), 2 = sensitivity, 3 = filter forward,
 * 4 = SpMV, 5 = assembly, 6 = CG iteration (assembled), 7 = K.u, previous-generation shuffle
 * kernels, 8 = K.u ring-staged with p.Ap and Ap.Ap, 9 = CG iteration (single-pass recurrence).
 * ms = average per rep. */
int topopt_time_kernel(topopt_handle* h, topopt_filter* f, int32_t which, int32_t reps,
                       double* ms);

#ifdef __cplusplus
}
#endif
#endif /* TOPOPT_CUDA_H */
