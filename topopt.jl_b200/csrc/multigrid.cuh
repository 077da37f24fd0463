// Geometric multigrid V-cycle on the structured grid, used as a preconditioner of CG (the reference
// zero-starts every solve and never refreshes its preconditioner, solvers_api.jl:187-203; this is the opt-in
// alternative).  Levels are the grid coarsened by 2 per axis; coarse operators are rediscretised with the mean
// modulus of the 8 children (times 2 per level: Ke of a brick scales with its edge length), applied matrix-free
// by the same K.u kernels; transfers are trilinear interpolation and its transpose; the smoother is a Chebyshev
// polynomial in D^-1 K, a fixed symmetric operator, so plain PCG applies; the coarsest level is solved exactly with a
// dense inverse built on the host.
#pragma once
#include "kernels.cuh"

namespace topopt {

// node (i, j, k) of a level stored with one ghost plane below: local node index
__device__ __forceinline__ long long mg_node(const Geo& g, int i, int j, int k) { return (long long)(k + 1) * g.S + (long long)j * g.NX + i; }

// Ec(I, J, K) = scale * mean of the 8 children of the fine level (element layers are stored with one ghost layer below)
__global__ void k_mg_coarsen_E(Geo gf, Geo gc, const double* __restrict__ Ef, double* __restrict__ Ec, double scale) {
  const long long ne = (long long)gc.SE * gc.NLg;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < ne; e += (long long)gridDim.x * blockDim.x) {
    const int I = (int)(e % gc.nx), J = (int)((e / gc.nx) % gc.ny), K = (int)(e / gc.SE);
    double acc = 0.0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int i = 2 * I + (c & 1), j = 2 * J + ((c >> 1) & 1), k = 2 * K + (c >> 2);
      acc += Ef[(long long)(k + 1) * gf.SE + (long long)j * gf.nx + i];
    }
    Ec[(long long)(K + 1) * gc.SE + (long long)J * gc.nx + I] = scale * 0.125 * acc;
  }
}

// a coarse node is prescribed where the fine node it coincides with is
__global__ void k_mg_coarsen_flags(Geo gf, Geo gc, const unsigned char* __restrict__ ff, unsigned char* __restrict__ fc) {
  const long long nn = (long long)gc.S * gc.NPg;
  for (long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x; n < nn; n += (long long)gridDim.x * blockDim.x) {
    const int I = (int)(n % gc.NX), J = (int)((n / gc.NX) % gc.NY), K = (int)(n / gc.S);
    fc[mg_node(gc, I, J, K)] = ff[mg_node(gf, 2 * I, 2 * J, 2 * K)];
  }
}

// restriction = transpose of trilinear interpolation: bc(I) = sum_{d in {-1,0,1}^3} 2^-(|dx|+|dy|+|dz|) rf(2I + d)
template <int NC>
__global__ void k_mg_restrict(Geo gf, Geo gc, const double* __restrict__ rf, double* __restrict__ bc,
                              const unsigned char* __restrict__ fc) {
  const long long nn = (long long)gc.S * gc.NPg;
  for (long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x; n < nn; n += (long long)gridDim.x * blockDim.x) {
    const int I = (int)(n % gc.NX), J = (int)((n / gc.NX) % gc.NY), K = (int)(n / gc.S);
    double acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] = 0.0;
    for (int dk = -1; dk <= 1; ++dk) {
      const int k = 2 * K + dk;
      if (k < 0 || k >= gf.NPg) continue;
      for (int dj = -1; dj <= 1; ++dj) {
        const int j = 2 * J + dj;
        if (j < 0 || j >= gf.NY) continue;
        for (int di = -1; di <= 1; ++di) {
          const int i = 2 * I + di;
          if (i < 0 || i >= gf.NX) continue;
          const double w = (di ? 0.5 : 1.0) * (dj ? 0.5 : 1.0) * (dk ? 0.5 : 1.0);
          const long long f = mg_node(gf, i, j, k) * NC;
#pragma unroll
          for (int c = 0; c < NC; ++c) acc[c] = fma(w, rf[f + c], acc[c]);
        }
      }
    }
    const long long o = mg_node(gc, I, J, K);
    const unsigned char fl = fc[o];
#pragma unroll
    for (int c = 0; c < NC; ++c) bc[o * NC + c] = (fl & (1 << c)) ? 0.0 : acc[c];
  }
}

// xf += P xc (trilinear interpolation), prescribed fine dofs stay zero
template <int NC>
__global__ void k_mg_prolong_add(Geo gf, Geo gc, const double* __restrict__ xc, double* __restrict__ xf,
                                 const unsigned char* __restrict__ ff) {
  const long long nn = (long long)gf.S * gf.NPg;
  for (long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x; n < nn; n += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(n % gf.NX), j = (int)((n / gf.NX) % gf.NY), k = (int)(n / gf.S);
    const int I0 = i >> 1, J0 = j >> 1, K0 = k >> 1;
    const int ni = i & 1, nj = j & 1, nk = k & 1;  // odd index: average of the two coarse neighbours on that axis
    double acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] = 0.0;
    for (int a = 0; a <= nk; ++a)
      for (int b = 0; b <= nj; ++b)
        for (int d = 0; d <= ni; ++d) {
          const long long o = mg_node(gc, I0 + d, J0 + b, K0 + a) * NC;
#pragma unroll
          for (int c = 0; c < NC; ++c) acc[c] += xc[o + c];
        }
    const double w = (ni ? 0.5 : 1.0) * (nj ? 0.5 : 1.0) * (nk ? 0.5 : 1.0);
    const long long f = mg_node(gf, i, j, k);
    const unsigned char fl = ff[f];
#pragma unroll
    for (int c = 0; c < NC; ++c)
      if (!(fl & (1 << c))) xf[f * NC + c] = fma(w, acc[c], xf[f * NC + c]);
  }
}

// Chebyshev smoother pieces on owned dofs [off, off + n)  (Saad, Iterative Methods, Alg. 12.1 with D^-1 K):
//   first:  [r = b - Kx ;]  d = (r / D) / theta
//   step :  x (+)= d ; r = rin - Kd ; d = c1 d + c2 (r / D)          (rin may alias r; rin = b on the first step)
//   last :  x (+)= d ; [r = rin - Kd]
template <bool RESID>
__global__ void __launch_bounds__(kBlock) k_cheb_first(long long off, long long n, const double* b, const double* Kx, double* r,
                                                       const double* __restrict__ D, double inv_theta, double* __restrict__ d) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const long long k = off + t;
    double rk = b[k];
    if (RESID) {
      rk -= Kx[k];
      r[k] = rk;
    }
    d[k] = inv_theta * (rk / D[k]);
  }
}
template <bool ZERO_X>
__global__ void __launch_bounds__(kBlock) k_cheb_step(long long off, long long n, double* __restrict__ x, const double* rin, double* r,
                                                      double* __restrict__ d, const double* __restrict__ Kd,
                                                      const double* __restrict__ D, double c1, double c2) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const long long k = off + t;
    const double dk = d[k];
    const double rk = rin[k] - Kd[k];
    x[k] = ZERO_X ? dk : x[k] + dk;
    r[k] = rk;
    d[k] = fma(c1, dk, c2 * (rk / D[k]));
  }
}
template <bool ZERO_X, bool RESID>
__global__ void __launch_bounds__(kBlock) k_cheb_last(long long off, long long n, double* __restrict__ x, const double* __restrict__ d,
                                                      const double* rin, double* r, const double* __restrict__ Kd) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const long long k = off + t;
    x[k] = ZERO_X ? d[k] : x[k] + d[k];
    if (RESID) r[k] = rin[k] - Kd[k];
  }
}

// outer PCG: x += alpha p ; r -= alpha Ap ; sum r.r -> st->sums[0]
__global__ void __launch_bounds__(kBlock) k_mg_update_xr(long long off, long long n, double* __restrict__ x, double* __restrict__ r,
                                                         const double* __restrict__ p, const double* __restrict__ Ap, double alpha,
                                                         double* partials, CGState* st) {
  __shared__ double sm[32];
  double v[1] = {0.0};
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const long long k = off + t;
    x[k] = fma(alpha, p[k], x[k]);
    const double rk = fma(-alpha, Ap[k], r[k]);
    r[k] = rk;
    v[0] = fma(rk, rk, v[0]);
  }
  block_partials_finish<1>(v, partials, st, FIN_PLAIN, sm);
}

// coarsest level: x = Ainv b (dense, n <= ~1500), one warp per row
__global__ void k_mg_dense_solve(int n, const double* __restrict__ Ainv, const double* __restrict__ b, double* __restrict__ x,
                                 const int* __restrict__ loc) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  double acc = 0.0;
  for (int c = lane; c < n; c += 32) acc = fma(Ainv[(long long)row * n + c], b[loc[c]], acc);
  acc = warp_sum(acc);
  if (lane == 0) x[loc[row]] = acc;
}

// z = r / D ; sums: r.z  (power iteration / diagnostics)
__global__ void __launch_bounds__(kBlock) k_scale_inv(long long off, long long n, const double* __restrict__ r,
                                                      const double* __restrict__ D, double s, double* __restrict__ z) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
    z[off + t] = s * (r[off + t] / D[off + t]);
}

}  // namespace topopt
