// Host side of the geometric-multigrid preconditioned CG (TOPOPT_PRECOND_MULTIGRID; kernels in multigrid.cuh).
// Included by api.cu after the K.u launchers.  Opt-in extension: the reference zero-starts every solve and
// builds its preconditioner once (src/FEA/solvers_api.jl:187-203); this path rebuilds the hierarchy whenever the
// stiffness changed and converges the large configurations in tens of iterations instead of thousands.
//
//   V-cycle(l): Chebyshev(m) pre-smoothing from x = 0 -> residual -> full-weighting restriction -> V-cycle(l + 1) ->
//               trilinear prolongation -> Chebyshev(m) post-smoothing;  coarsest level: dense inverse.
//   The smoother is a fixed polynomial in D^-1 K and the same one before and after, so the cycle is a symmetric
//   positive definite operator and ordinary PCG applies.
//   Coarse operators are the same matrix-free K.u kernels on the coarsened grid with E_c = 2 * mean(E of the 8 children)
//   (a hex8 stiffness scales linearly with the edge length), prescribed where the coincident fine node is.

constexpr int kMGStopDofs = 400;      // stop coarsening below this many dofs ...
constexpr int kMGMaxCoarse = 2600;    // ... and refuse a coarsest level that is still larger than this (dense inverse on the host)

static void mg_free(MGHierarchy* m) {
  if (!m) return;
  for (MGLevel& L : m->lv) {
    for (double* q : {L.x, L.r, L.d, L.t, L.D})
      if (q) cudaFree(q);
    if (L.owns) {
      if (L.E) cudaFree(L.E);
      if (L.b) cudaFree(L.b);
      if (L.fixed) cudaFree(L.fixed);
    }
  }
  if (m->d_Ainv) cudaFree(m->d_Ainv);
  if (m->d_loc) cudaFree(m->d_loc);
  delete m;
}

struct LevelScope {  // K.u launchers act on a coarse level while this is alive
  topopt_handle* h;
  LevelScope(topopt_handle* hh, const MGLevel* L, bool coarse) : h(hh) { h->lvl = coarse ? L : nullptr; }
  ~LevelScope() { h->lvl = nullptr; }
};

static int mg_apply_level(topopt_handle* h, int l, const double* x, double* y) {
  LevelScope scope(h, &h->mg->lv[l], l > 0);
  return launch_apply<false>(h, x, y, FIN_NONE);
}

static int mg_host_scalar(topopt_handle* h, double* out) {
  CUDA_TRY(h, cudaMemcpyAsync(h->h_st, h->d_st, 8 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  *out = h->h_st->sums[0];
  return TOPOPT_OK;
}

static int mg_dot(topopt_handle* h, const MGLevel& L, const double* a, const double* b, double* out) {
  LAUNCH(h, k_dot, grid_for(L.nown_dofs), L.off, L.nown_dofs, a, b, h->d_partials, h->d_st);
  TRY(check_launch(h, "k_dot"));
  return mg_host_scalar(h, out);
}

static int mg_build(topopt_handle* h) {
  if (h->mg) return TOPOPT_OK;
  if (h->dim != 3) return fail(h, TOPOPT_ERR_INVALID, "multigrid preconditioner: 3-D structured grids only");
  if (h->world != 1) return fail(h, TOPOPT_ERR_INVALID, "multigrid preconditioner: single-GPU only");
  MGHierarchy* m = new MGHierarchy();
  auto bail = [&](int rc) {
    mg_free(m);
    return rc;
  };
  auto alloc_vecs = [&](MGLevel& L) -> int {
    for (double** q : {&L.x, &L.r, &L.d, &L.t, &L.D}) TRY(dev_alloc(h, q, (size_t)L.nloc_dofs));
    return TOPOPT_OK;
  };
  {
    MGLevel L;
    L.g = h->g;
    L.nloc_nodes = h->nloc_nodes;
    L.nloc_dofs = h->nloc_dofs;
    L.off = h->off;
    L.nown_dofs = h->nown_dofs;
    L.nloc_el = h->nloc_el;
    L.E = h->d_E;
    L.fixed = h->d_fixed;
    L.owns = false;
    m->lv.push_back(L);
    if (alloc_vecs(m->lv.back()) != TOPOPT_OK) return bail(TOPOPT_ERR_CUDA);
  }
  while (true) {
    const Geo gf = m->lv.back().g;
    const long long dofs = (long long)gf.S * gf.NPg * h->nc;
    if (dofs <= kMGStopDofs || (gf.nx & 1) || (gf.ny & 1) || (gf.NLg & 1)) break;
    MGLevel L;
    Geo& g = L.g;
    g.nx = gf.nx / 2;
    g.ny = gf.ny / 2;
    g.NX = g.nx + 1;
    g.NY = g.ny + 1;
    g.S = g.NX * g.NY;
    g.SE = g.nx * g.ny;
    g.NLg = gf.NLg / 2;
    g.NPg = g.NLg + 1;
    g.p0 = -1;
    g.nlay = g.NLg;
    g.nown = g.NPg;
    L.nloc_nodes = (int64_t)g.S * (g.nown + 2);
    L.nloc_dofs = L.nloc_nodes * h->nc;
    L.off = (int64_t)g.S * h->nc;
    L.nown_dofs = L.off * g.nown;
    L.nloc_el = (int64_t)g.SE * (g.nown + 1);
    L.owns = true;
    m->lv.push_back(L);
    MGLevel& B = m->lv.back();
    if (dev_alloc(h, &B.E, (size_t)B.nloc_el) != TOPOPT_OK || dev_alloc(h, &B.b, (size_t)B.nloc_dofs) != TOPOPT_OK ||
        dev_alloc(h, &B.fixed, (size_t)B.nloc_nodes) != TOPOPT_OK || alloc_vecs(B) != TOPOPT_OK)
      return bail(TOPOPT_ERR_CUDA);
    const MGLevel& F = m->lv[m->lv.size() - 2];
    k_mg_coarsen_flags<<<grid_for((long long)g.S * g.NPg, kWideGrid), kBlock, 0, h->stream>>>(F.g, g, F.fixed, B.fixed);
    h->stats.kernel_launches += 1;
  }
  const MGLevel& C = m->lv.back();
  m->ncoarse = (int)((long long)C.g.S * C.g.NPg * h->nc);
  if (m->lv.size() < 2 || m->ncoarse > kMGMaxCoarse) {
    const std::string msg = "multigrid preconditioner: the grid cannot be halved far enough (coarsest level would have " +
                            std::to_string(m->ncoarse) + " dofs); use element counts with more factors of two";
    mg_free(m);
    return fail(h, TOPOPT_ERR_INVALID, msg);
  }
  if (dev_alloc(h, &m->d_Ainv, (size_t)m->ncoarse * m->ncoarse) != TOPOPT_OK || dev_alloc(h, &m->d_loc, (size_t)m->ncoarse) != TOPOPT_OK)
    return bail(TOPOPT_ERR_CUDA);
  {
    std::vector<int> loc(m->ncoarse);
    for (int dof = 0; dof < m->ncoarse; ++dof) loc[dof] = (dof / h->nc + C.g.S) * h->nc + dof % h->nc;  // one ghost plane below
    if (cudaMemcpyAsync(m->d_loc, loc.data(), sizeof(int) * m->ncoarse, cudaMemcpyHostToDevice, h->stream) != cudaSuccess ||
        cudaStreamSynchronize(h->stream) != cudaSuccess)
      return bail(fail(h, TOPOPT_ERR_CUDA, "multigrid: upload of the coarse index map failed"));
  }
  if (check_launch(h, "k_mg_coarsen_flags") != TOPOPT_OK) return bail(TOPOPT_ERR_CUDA);
  h->mg = m;
  h->mg_dirty = true;
  return TOPOPT_OK;
}

// dense inverse of the coarsest operator (host): A = sum_e E_e Ke with identity rows / columns on prescribed dofs
static int mg_coarse_inverse(topopt_handle* h) {
  MGHierarchy* m = h->mg;
  const MGLevel& C = m->lv.back();
  const int n = m->ncoarse, nc = h->nc, ks = h->ks;
  std::vector<double> Ec((size_t)C.nloc_el);
  std::vector<unsigned char> fl((size_t)C.nloc_nodes);
  CUDA_TRY(h, cudaMemcpyAsync(Ec.data(), C.E, sizeof(double) * Ec.size(), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(fl.data(), C.fixed, fl.size(), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  const int64_t nels[3] = {C.g.nx, C.g.ny, C.g.NLg};
  const GridDims gd = make_dims(3, nels);
  std::vector<double> A((size_t)n * n, 0.0);
  int64_t nodes[8];
  for (int64_t e = 0; e < gd.nel; ++e) {
    const double Ee = Ec[(size_t)(e + C.g.SE)];  // one ghost layer below
    cell_nodes(gd, e, nodes);
    for (int a = 0; a < 8; ++a)
      for (int ca = 0; ca < nc; ++ca) {
        const int64_t row = nodes[a] * nc + ca;
        for (int b = 0; b < 8; ++b)
          for (int cb = 0; cb < nc; ++cb) A[(size_t)row * n + nodes[b] * nc + cb] += Ee * h->Ke[(a * nc + ca) + ks * (b * nc + cb)];
      }
  }
  std::vector<char> fixed(n, 0);
  double dmean = 0.0;
  for (int d = 0; d < n; ++d) {
    fixed[d] = (fl[(size_t)(d / nc + C.g.S)] >> (d % nc)) & 1;
    dmean += A[(size_t)d * n + d];
  }
  dmean /= n;
  for (int d = 0; d < n; ++d) {
    if (fixed[d]) {
      for (int k = 0; k < n; ++k) A[(size_t)d * n + k] = A[(size_t)k * n + d] = 0.0;
      A[(size_t)d * n + d] = 1.0;
    } else {
      A[(size_t)d * n + d] += 1e-12 * dmean;  // coincident-node restriction of the supports may leave a rigid mode
    }
  }
  // Cholesky A = L L' (lower triangle in place), then A^-1 = L^-T L^-1 column by column
  for (int j = 0; j < n; ++j) {
    double s = A[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) s -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
    if (!(s > 0.0)) return fail(h, TOPOPT_ERR_INVALID, "multigrid: the coarsest operator is not positive definite (supports too sparse?)");
    const double ljj = std::sqrt(s);
    A[(size_t)j * n + j] = ljj;
    for (int i = j + 1; i < n; ++i) {
      double t = A[(size_t)i * n + j];
      const double *ri = &A[(size_t)i * n], *rj = &A[(size_t)j * n];
      for (int k = 0; k < j; ++k) t -= ri[k] * rj[k];
      A[(size_t)i * n + j] = t / ljj;
    }
  }
  std::vector<double> Ainv((size_t)n * n, 0.0), col(n);
  for (int c = 0; c < n; ++c) {
    if (fixed[c]) continue;  // prescribed dofs: zero row and column, the coarse correction leaves them alone
    for (int i = 0; i < n; ++i) {  // L y = e_c
      double t = i == c ? 1.0 : 0.0;
      if (i < c) {
        col[i] = 0.0;
        continue;
      }
      const double* ri = &A[(size_t)i * n];
      for (int k = c; k < i; ++k) t -= ri[k] * col[k];
      col[i] = t / ri[i];
    }
    for (int i = n - 1; i >= 0; --i) {  // L' x = y
      double t = col[i];
      for (int k = i + 1; k < n; ++k) t -= A[(size_t)k * n + i] * col[k];
      col[i] = t / A[(size_t)i * n + i];
    }
    for (int i = 0; i < n; ++i) Ainv[(size_t)i * n + c] = fixed[i] ? 0.0 : col[i];
  }
  CUDA_TRY(h, cudaMemcpyAsync(m->d_Ainv, Ainv.data(), sizeof(double) * Ainv.size(), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return TOPOPT_OK;
}

// coarse stiffness, diagonals, smoother bounds, coarsest inverse: after every change of the fine stiffness
static int mg_refresh(topopt_handle* h) {
  MGHierarchy* m = h->mg;
  const int nl = (int)m->lv.size();
  for (int l = 1; l < nl; ++l) {
    const MGLevel &F = m->lv[l - 1], &C = m->lv[l];
    k_mg_coarsen_E<<<grid_for((long long)C.g.SE * C.g.NLg, kWideGrid), kBlock, 0, h->stream>>>(F.g, C.g, F.E, C.E, 2.0);
    h->stats.kernel_launches += 1;
  }
  for (int l = 0; l < nl; ++l) {
    const MGLevel& L = m->lv[l];
#define CALL(DIM_, NC_) LAUNCH(h, (k_diag<DIM_, NC_>), grid_for((long long)L.g.S * L.g.nown, kWideGrid), L.g, L.D, L.E, L.fixed, h->fixed_diag)
    DISPATCH(h, CALL);
#undef CALL
  }
  TRY(check_launch(h, "multigrid setup"));
  // lambda_max(D^-1 K) per smoothed level by power iteration (v in d, K v in t)
  for (int l = 0; l + 1 < nl; ++l) {
    MGLevel& L = m->lv[l];
    LAUNCH(h, k_fill_dense, grid_for(L.nown_dofs), L.off, L.nown_dofs, L.d);
#define CALL(DIM_, NC_) LAUNCH(h, (k_apply_zero<NC_>), grid_for(L.nloc_nodes, kWideGrid), (long long)L.nloc_nodes, L.fixed, L.d)
    DISPATCH(h, CALL);
#undef CALL
    double nrm2 = 0.0, lam = 1.0;
    TRY(mg_dot(h, L, L.d, L.d, &nrm2));
    const int its = 15;
    for (int k = 0; k < its; ++k) {
      TRY(mg_apply_level(h, l, L.d, L.t));
      LAUNCH(h, k_scale_inv, grid_for(L.nown_dofs), L.off, L.nown_dofs, L.t, L.D, 1.0 / std::sqrt(nrm2), L.d);  // d = D^-1 K v / |v|
      double n2 = 0.0;
      TRY(mg_dot(h, L, L.d, L.d, &n2));
      if (!(n2 > 0.0) || !std::isfinite(n2)) return fail(h, TOPOPT_ERR_NONFINITE, "multigrid: power iteration broke down");
      lam = std::sqrt(n2);  // |D^-1 K v| / |v| with |v| = 1 after the scaling
      nrm2 = n2;
    }
    L.lmax = 1.2 * lam;  // power iteration approaches lambda_max from below
  }
  TRY(mg_coarse_inverse(h));
  h->mg_dirty = false;
  return TOPOPT_OK;
}

// x = S(b) from a zero initial guess (ZERO) or improving x; optionally leaves r = b - K x
static int mg_smooth(topopt_handle* h, int l, const double* b, bool zero_start, bool want_resid) {
  MGHierarchy* m = h->mg;
  MGLevel& L = m->lv[l];
  const int deg = std::max(1, m->degree);
  const double lmax = L.lmax, lmin = lmax / m->ratio;
  const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
  const int grid = grid_for(L.nown_dofs);
  const long long off = L.off, n = L.nown_dofs;
  if (zero_start) {
    LAUNCH(h, (k_cheb_first<false>), grid, off, n, b, (const double*)nullptr, L.r, L.D, 1.0 / theta, L.d);
  } else {
    TRY(mg_apply_level(h, l, L.x, L.t));
    LAUNCH(h, (k_cheb_first<true>), grid, off, n, b, L.t, L.r, L.D, 1.0 / theta, L.d);
  }
  double rho = 1.0 / sigma;
  const double* rin = zero_start ? b : L.r;
  bool xzero = zero_start;
  for (int k = 1; k < deg; ++k) {
    const double rho1 = 1.0 / (2.0 * sigma - rho);
    TRY(mg_apply_level(h, l, L.d, L.t));
    if (xzero)
      LAUNCH(h, (k_cheb_step<true>), grid, off, n, L.x, rin, L.r, L.d, L.t, L.D, rho1 * rho, 2.0 * rho1 / delta);
    else
      LAUNCH(h, (k_cheb_step<false>), grid, off, n, L.x, rin, L.r, L.d, L.t, L.D, rho1 * rho, 2.0 * rho1 / delta);
    rho = rho1;
    rin = L.r;
    xzero = false;
  }
  if (want_resid) TRY(mg_apply_level(h, l, L.d, L.t));
  if (xzero) {
    if (want_resid)
      LAUNCH(h, (k_cheb_last<true, true>), grid, off, n, L.x, L.d, rin, L.r, L.t);
    else
      LAUNCH(h, (k_cheb_last<true, false>), grid, off, n, L.x, L.d, rin, L.r, L.t);
  } else {
    if (want_resid)
      LAUNCH(h, (k_cheb_last<false, true>), grid, off, n, L.x, L.d, rin, L.r, L.t);
    else
      LAUNCH(h, (k_cheb_last<false, false>), grid, off, n, L.x, L.d, rin, L.r, L.t);
  }
  return TOPOPT_OK;
}

static int mg_vcycle(topopt_handle* h, int l, const double* b) {
  MGHierarchy* m = h->mg;
  const int nl = (int)m->lv.size();
  MGLevel& L = m->lv[l];
  if (l == nl - 1) {
    const int n = m->ncoarse;
    k_mg_dense_solve<<<(n * 32 + 127) / 128, 128, 0, h->stream>>>(n, m->d_Ainv, b, L.x, m->d_loc);
    h->stats.kernel_launches += 1;
    return TOPOPT_OK;
  }
  MGLevel& C = m->lv[l + 1];
  TRY(mg_smooth(h, l, b, true, true));
  const int gc = grid_for((long long)C.g.S * C.g.NPg, kWideGrid), gf = grid_for((long long)L.g.S * L.g.NPg, kWideGrid);
  if (h->nc == 3)
    k_mg_restrict<3><<<gc, kBlock, 0, h->stream>>>(L.g, C.g, L.r, C.b, C.fixed);
  else
    k_mg_restrict<1><<<gc, kBlock, 0, h->stream>>>(L.g, C.g, L.r, C.b, C.fixed);
  h->stats.kernel_launches += 1;
  TRY(mg_vcycle(h, l + 1, C.b));
  if (h->nc == 3)
    k_mg_prolong_add<3><<<gf, kBlock, 0, h->stream>>>(L.g, C.g, C.x, L.x, L.fixed);
  else
    k_mg_prolong_add<1><<<gf, kBlock, 0, h->stream>>>(L.g, C.g, C.x, L.x, L.fixed);
  h->stats.kernel_launches += 1;
  return mg_smooth(h, l, b, false, false);
}

// Preconditioned CG with one V-cycle per iteration.  Scalars travel to the host (three small reads per iteration next to
// a millisecond of V-cycle); tolerances and the iteration count mean what they mean in cg_solve.
static int mg_pcg_solve(topopt_handle* h, const double* b, const topopt_cg_opts* o, topopt_cg_result* res) {
  if (o->op == TOPOPT_OP_ASSEMBLED) return fail(h, TOPOPT_ERR_INVALID, "multigrid preconditioner: matrix-free operator only");
  if (o->criteria != TOPOPT_CRITERIA_DEFAULT) return fail(h, TOPOPT_ERR_INVALID, "multigrid preconditioner: default convergence criteria only");
  TRY(mg_build(h));
  MGHierarchy* m = h->mg;
  m->degree = o->mg_degree > 0 ? o->mg_degree : 2;
  m->ratio = o->mg_ratio > 1.0 ? o->mg_ratio : 6.0;
  if (h->mg_dirty || o->refresh_precond) TRY(mg_refresh(h));
  MGLevel& L0 = m->lv[0];
  CGState& s = *h->h_st;
  std::memset(&s, 0, sizeof(CGState));
  s.world = 1;
  CUDA_TRY(h, cudaMemcpyAsync(h->d_st, &s, sizeof(CGState), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
  const int vgrid = grid_for(h->nown_dofs);
  const long long off = h->off, n = h->nown_dofs;
  if (o->warm_start) {
    TRY(launch_apply<false>(h, h->d_u, h->d_Ap, FIN_NONE));
    LAUNCH(h, k_axpby, vgrid, off, n, 1.0, b, -1.0, h->d_Ap, h->d_r);
  } else {
    CUDA_TRY(h, cudaMemsetAsync(h->d_u + off, 0, sizeof(double) * n, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_r + off, b + off, sizeof(double) * n, cudaMemcpyDeviceToDevice, h->stream));
  }
  double rr = 0.0;
  TRY(mg_dot(h, L0, h->d_r, h->d_r, &rr));
  double resn = std::sqrt(rr);
  const double tol = std::max(o->reltol * resn, o->abstol);
  int iters = 0;
  bool conv = resn <= tol, bad = !std::isfinite(resn);
  double rz_old = 1.0;
  int restarts = 0;
  bool fresh = true;  // next direction is the preconditioned residual itself
  while (!conv && !bad && iters < o->maxiter) {
    TRY(mg_vcycle(h, 0, h->d_r));
    m->cycles += 1;
    double rz = 0.0;
    TRY(mg_dot(h, L0, h->d_r, L0.x, &rz));
    const double beta = fresh ? 0.0 : rz / rz_old;
    fresh = false;
    LAUNCH(h, k_axpby, vgrid, off, n, 1.0, L0.x, beta, h->d_p, h->d_p);  // p = z + beta p
    TRY(launch_cg_apply<1>(h, false, FIN_NONE));                         // Ap = K p, sums[0] = p.Ap
    double pAp = 0.0;
    TRY(mg_host_scalar(h, &pAp));
    if (!(pAp > 0.0) || !(rz > 0.0)) {
      bad = !std::isfinite(pAp) || !std::isfinite(rz);
      if (bad || restarts >= 4) break;
      // r.M^-1 r <= 0: the cycle was not positive definite, i.e. a smoother bound was too low for this stiffness.
      // x and r are still consistent: raise the bounds and restart the recurrence from the current iterate.
      for (MGLevel& L : m->lv) L.lmax *= 1.3;
      restarts += 1;
      fresh = true;
      continue;
    }
    const double alpha = rz / pAp;
    LAUNCH(h, k_mg_update_xr, vgrid, off, n, h->d_u, h->d_r, h->d_p, h->d_Ap, alpha, h->d_partials, h->d_st);
    TRY(check_launch(h, "k_mg_update_xr"));
    TRY(mg_host_scalar(h, &rr));
    resn = std::sqrt(rr);
    rz_old = rz;
    iters += 1;
    conv = resn <= tol;
    bad = !std::isfinite(resn);
  }
  CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
  TRY(exchange_halo(h, h->d_u));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  float ms = 0;
  CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->stats.cg_iterations += iters;
  h->stats.last_solve_ms = ms;
  if (res) {
    res->iters = iters;
    res->converged = conv ? 1 : 0;
    res->residual = resn;
    res->tol = tol;
    res->solve_ms = ms;
  }
  if (bad) return fail(h, TOPOPT_ERR_NONFINITE, "multigrid PCG: NaN in the residual");
  return TOPOPT_OK;
}
