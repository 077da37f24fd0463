// Matrix-free K.u for hex8 elasticity, "modal" form.
//
// On a brick element with an isotropic material the 24x24 Ke becomes very sparse in the
// tensor-product Haar basis {1,x,y,z,xy,yz,xz,xyz} (x) {ux,uy,uz}:  Khat = T' Ke T / 64 with
// T = H (x) I3 (H = 8x8 Hadamard over the corners) has 45 non-zeros in a fixed pattern, so
//     f_e = E_e * H ( Khat ( H' u_e ) )
// costs ~200 fp64 instructions instead of 576 DFMA.  The pattern is verified numerically on the
// host for the Ke the caller passes; any other Ke falls back to the dense gather kernel.
//
// Work decomposition: a persistent CTA of 32 x TY threads (one per SM) loads a 32 x TY patch of node
// columns, evaluates the 31 x (TY-1) element columns inside it and owns the 30 x (TY-2) interior node
// columns; it marches along the slab axis over an equal share of the linearised (tile, plane) space.
// Thread (tx,ty) owns node column (i0-1+tx, j0-1+ty) and the element column whose lower corner is that
// node.  Per step it loads ONE new node (3 doubles, prefetched two planes ahead), forms the z stage of
// the Hadamard transform once per node column and shares it (x+1 by warp shuffle, y+1 row through
// shared memory), evaluates its element in registers, and the corner forces are reduced onto node
// columns before the inverse z stage (x by shuffle, y by one shared-memory hop); z is a register
// carry.  Rows synchronise only with their two neighbours.  No atomics, fixed summation order.
#pragma once
#include "kernels.cuh"

namespace topopt {

__constant__ double cKh[48];  // the 45 modal coefficients, order documented in modal_pattern()

// (row, col) pairs of the modal matrix in the order the kernel consumes them.
// modal dof index = 3*mode + comp, modes: 0:1 1:x 2:y 3:z 4:xy 5:yz 6:xz 7:xyz
struct ModalEntry { int row, col; };
inline const ModalEntry* modal_pattern() {
  static const ModalEntry p[45] = {
      // normal block {x:x, y:y, z:z}
      {3, 3}, {3, 7}, {3, 11}, {7, 3}, {7, 7}, {7, 11}, {11, 3}, {11, 7}, {11, 11},
      // shear pairs {x:y, y:x}, {x:z, z:x}, {y:z, z:y}
      {4, 4}, {4, 6}, {6, 4}, {6, 6},
      {5, 5}, {5, 9}, {9, 5}, {9, 9},
      {8, 8}, {8, 10}, {10, 8}, {10, 10},
      // bilinear pairs {xy:x, yz:z}, {xy:y, xz:z}, {yz:y, xz:x}
      {12, 12}, {12, 17}, {17, 12}, {17, 17},
      {13, 13}, {13, 20}, {20, 13}, {20, 20},
      {16, 16}, {16, 18}, {18, 16}, {18, 18},
      // bilinear triple {xy:z, yz:x, xz:y}
      {14, 14}, {14, 15}, {14, 19}, {15, 14}, {15, 15}, {15, 19}, {19, 14}, {19, 15}, {19, 19},
      // trilinear diagonal
      {21, 21}, {22, 22}, {23, 23}};
  return p;
}

// FUSEP: x is formed on the fly as p_new = r + beta * p_old (the CG direction update,
// IterativeSolvers' `u .= r .+ beta .* u`), written to pnew on owned nodes, so that the separate
// vector pass disappears.  p_old / p_new are distinct buffers (neighbouring CTAs read halo values
// of p_old while the owner writes p_new).
template <int TY, bool DOT, bool FUSEP, bool PEER, bool NSYNC>
__global__ void __launch_bounds__(32 * TY, (TY <= 4 ? 4 : (TY <= 8 ? 2 : 1)))
    k_apply_hex8_modal(Geo g, const double* __restrict__ x, double* __restrict__ y, const double* __restrict__ E,
                       const unsigned char* __restrict__ fixed, double fixed_diag, int tilesX, int tilesY,
                       double* partials, CGState* st, int fin, const double* __restrict__ rvec, double* __restrict__ pnew,
                       const double* __restrict__ xlo, const double* __restrict__ xhi) {
  extern __shared__ double smem_dyn[];
  // two parity buffers, each [12][TY][32]: rows 0-5 = S/D of the next plane, rows 6-11 = partial node sums
  double(*sX)[12][TY][32] = reinterpret_cast<double(*)[12][TY][32]>(smem_dyn);
  __shared__ double sm[32];
  // NSYNC: rows synchronise only with their two neighbours (a per-row "published step" counter in
  // shared memory) instead of a CTA-wide barrier, so warps drift apart and the exchange phase of
  // one row overlaps the fp64 phase of another.  (Pairwise named hardware barriers were tried:
  // the two-sided rendezvous couples the rows too tightly and measured 20 % slower at TY = 16.)
  __shared__ int sflag[TY];
  int it = 0;  // published-step counter, monotonic across segments
  if ((DOT || FUSEP) && st->done) return;
  if (NSYNC) {
    if (threadIdx.x < TY) sflag[threadIdx.x] = 0;
  }
  const double beta = FUSEP ? st->beta : 0.0;
  const int tid = threadIdx.x;
  const int tx = tid & 31, ty = tid >> 5;
  const unsigned FULL = 0xffffffffu;
  const long long xs = (long long)g.S * 3;
  double dot = 0.0;
  // Persistent CTAs: the (tile, plane) space is linearised and cut into gridDim.x equal ranges,
  // so every CTA marches the same number of planes (no wave quantisation, and only one redundant
  // priming layer per segment).
  const long long units = (long long)tilesX * tilesY * g.nown;
  long long u0 = units * blockIdx.x / gridDim.x;
  const long long u1 = units * (blockIdx.x + 1) / gridDim.x;
  while (u0 < u1) {
  const int tile = (int)(u0 / g.nown);
  const int zoff = (int)(u0 % g.nown);
  const int zlen = (int)min((long long)(g.nown - zoff), u1 - u0);
  u0 += zlen;
  const int bx = tile % tilesX;
  const int by = tile / tilesX;
  const int in = bx * 30 - 1 + tx;          // node / element column
  const int jn = by * (TY - 2) - 1 + ty;
  const int z0 = 1 + zoff;                  // first owned local plane of this segment
  const int z1 = z0 + zlen;                 // one past the last
  __syncthreads();                          // shared buffers are reused across segments
  const bool node_ok = in >= 0 && in < g.NX && jn >= 0 && jn < g.NY;
  const bool own = node_ok && tx >= 1 && tx <= 30 && ty >= 1 && ty <= TY - 2;
  const bool elem_ok = in >= 0 && in < g.nx && jn >= 0 && jn < g.ny && tx < 31 && ty < TY - 1;
  const long long ncol = node_ok ? (long long)jn * g.NX + in : 0;
  const long long ecol = elem_ok ? (long long)jn * g.nx + in : 0;
  long long noff = ((long long)(z0 - 1) * g.S + ncol) * 3;  // own node, plane z0-1
  const unsigned char* fp = fixed + (long long)(z0 - 1) * g.S + ncol;
  const double* Ep = E + (long long)(z0 - 1) * g.SE + ecol;

  // Multi-GPU: the ghost planes (local 0 and nown+1) are read straight from the slab neighbours'
  // memory (xlo = their top owned plane, xhi = their bottom owned plane); CTAs that touch them
  // first wait for the neighbours' "direction vector final" flag of this iteration.
  if (PEER && (xlo != nullptr || xhi != nullptr)) {
    const bool need_lo = xlo != nullptr && z0 == 1, need_hi = xhi != nullptr && z1 == g.nown + 1;
    if ((need_lo || need_hi) && tid == 0) {
      PeerComm* pc = st->peer;
      const unsigned long long want = pc->halo_seq;  // k_signal_halo ran just before this kernel
      volatile unsigned long long* f = pc->block[pc->rank]->halo_flag;
      long long spins = 0;
      while ((need_lo && f[0] < want) || (need_hi && f[1] < want)) {
        if (++spins > kSpinLimit) {
          pc->timeout = 1;
          break;
        }
      }
      __threadfence_system();
    }
    __syncthreads();
  }
  auto load_node = [&](long long off, double (&v)[3]) {
    const double* src = x + off;
    bool remote = false;  // CTA-uniform: ghost planes live in the neighbours' memory
    if (PEER) {
      if (xlo != nullptr && off < xs) {
        src = xlo + ncol * 3;
        remote = true;
      }
      if (xhi != nullptr && off >= xs * (g.nown + 1)) {
        src = xhi + ncol * 3;
        remote = true;
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      // peer planes change between launches and are not L1-coherent: read them through L2
      double a = node_ok ? (remote ? __ldcg(src + c) : src[c]) : 0.0;
      if (FUSEP) a = fma(beta, a, node_ok ? rvec[off + c] : 0.0);
      v[c] = a;
    }
  };

  // own node column: masked bottom value, raw bottom value (row mask / dot), flag
  double vb[3], xo[3], carry[3] = {0.0, 0.0, 0.0};
  unsigned char fo = node_ok ? fp[0] : 0;
  load_node(noff, xo);
#pragma unroll
  for (int c = 0; c < 3; ++c) vb[c] = (fo & (1 << c)) ? 0.0 : xo[c];
  // plane z0 -> S/D of the first layer published to buffer 0
  double rt[3];
  unsigned char ftf = node_ok ? fp[g.S] : 0;
  load_node(noff + xs, rt);
  double vt[3], S0[3], D0[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    vt[c] = (ftf & (1 << c)) ? 0.0 : rt[c];
    S0[c] = vt[c] + vb[c];
    D0[c] = vt[c] - vb[c];
    sX[0][c][ty][tx] = S0[c];
    sX[0][3 + c][ty][tx] = D0[c];
  }
  double Ee;
  {
    const int gl = z0 - 1 + g.p0;
    Ee = (elem_ok && gl >= 0 && gl < g.NLg) ? Ep[0] : 0.0;
  }
  // software prefetch of plane z0+1 / layer z0
  double rn[3] = {0.0, 0.0, 0.0};
  unsigned char fn = 0;
  double En = 0.0;
  noff += xs;
  fp += g.S;
  Ep += g.SE;
  if (z0 < z1) {
    fn = node_ok ? fp[g.S] : 0;
    load_node(noff + xs, rn);
    const int gl = z0 + g.p0;
    En = (elem_ok && gl >= 0 && gl < g.NLg) ? Ep[0] : 0.0;
  }
  __syncthreads();

  int par = 0;
  for (int ll = z0 - 1; ll < z1; ++ll, par ^= 1) {  // element layer ll touches planes ll (bottom), ll+1 (top)
    // gather S/D of the three neighbour columns (x+1 by shuffle, y+1 row from shared memory)
    double S[4][3], D[4][3];  // columns: 0 own, 1 x+1, 2 y+1, 3 x+1,y+1
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      S[0][c] = S0[c];
      D[0][c] = D0[c];
      S[1][c] = __shfl_down_sync(FULL, S0[c], 1);
      D[1][c] = __shfl_down_sync(FULL, D0[c], 1);
      S[2][c] = sX[par][c][ty + 1 < TY ? ty + 1 : ty][tx];
      D[2][c] = sX[par][3 + c][ty + 1 < TY ? ty + 1 : ty][tx];
      S[3][c] = __shfl_down_sync(FULL, S[2][c], 1);
      D[3][c] = __shfl_down_sync(FULL, D[2][c], 1);
    }
    // x, y stages -> modal coefficients scaled by E_e (the constant mode is never needed)
    double X[3], Y[3], Z[3], XY[3], YZ[3], XZ[3], XYZ[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double ss0 = S[1][c] + S[0][c], ds0 = S[1][c] - S[0][c], ss1 = S[3][c] + S[2][c], ds1 = S[3][c] - S[2][c];
      const double sd0 = D[1][c] + D[0][c], dd0 = D[1][c] - D[0][c], sd1 = D[3][c] + D[2][c], dd1 = D[3][c] - D[2][c];
      X[c] = Ee * (ds1 + ds0);
      Y[c] = Ee * (ss1 - ss0);
      Z[c] = Ee * (sd1 + sd0);
      XY[c] = Ee * (ds1 - ds0);
      YZ[c] = Ee * (sd1 - sd0);
      XZ[c] = Ee * (dd1 + dd0);
      XYZ[c] = Ee * (dd1 - dd0);
    }
    double vX[3], vY[3], vZ[3], vXY[3], vYZ[3], vXZ[3], vXYZ[3];
    vX[0] = fma(cKh[2], Z[2], fma(cKh[1], Y[1], cKh[0] * X[0]));
    vY[1] = fma(cKh[5], Z[2], fma(cKh[4], Y[1], cKh[3] * X[0]));
    vZ[2] = fma(cKh[8], Z[2], fma(cKh[7], Y[1], cKh[6] * X[0]));
    vX[1] = fma(cKh[10], Y[0], cKh[9] * X[1]);
    vY[0] = fma(cKh[12], Y[0], cKh[11] * X[1]);
    vX[2] = fma(cKh[14], Z[0], cKh[13] * X[2]);
    vZ[0] = fma(cKh[16], Z[0], cKh[15] * X[2]);
    vY[2] = fma(cKh[18], Z[1], cKh[17] * Y[2]);
    vZ[1] = fma(cKh[20], Z[1], cKh[19] * Y[2]);
    vXY[0] = fma(cKh[22], YZ[2], cKh[21] * XY[0]);
    vYZ[2] = fma(cKh[24], YZ[2], cKh[23] * XY[0]);
    vXY[1] = fma(cKh[26], XZ[2], cKh[25] * XY[1]);
    vXZ[2] = fma(cKh[28], XZ[2], cKh[27] * XY[1]);
    vYZ[1] = fma(cKh[30], XZ[0], cKh[29] * YZ[1]);
    vXZ[0] = fma(cKh[32], XZ[0], cKh[31] * YZ[1]);
    vXY[2] = fma(cKh[35], XZ[1], fma(cKh[34], YZ[0], cKh[33] * XY[2]));
    vYZ[0] = fma(cKh[38], XZ[1], fma(cKh[37], YZ[0], cKh[36] * XY[2]));
    vXZ[1] = fma(cKh[41], XZ[1], fma(cKh[40], YZ[0], cKh[39] * XY[2]));
    vXYZ[0] = cKh[42] * XYZ[0];
    vXYZ[1] = cKh[43] * XYZ[1];
    vXYZ[2] = cKh[44] * XYZ[2];
    // inverse Hadamard in y and x; the four corner columns are reduced onto node columns
    // BEFORE the z stage (x by shuffle, y through shared memory), so the z stage runs once per node
    double H0[3], H1[3];  // node-column sums for mz = 0, 1
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double g10a = vX[c] + vXY[c], g10b = vX[c] - vXY[c];      // (mx=1,mz=0): oy=1, oy=0
      const double g01a = vZ[c] + vYZ[c], g01b = vZ[c] - vYZ[c];      // (mx=0,mz=1)
      const double g11a = vXZ[c] + vXYZ[c], g11b = vXZ[c] - vXYZ[c];  // (mx=1,mz=1)
      const double h110 = vY[c] + g10a, h010 = vY[c] - g10a;          // h[ox][oy][mz]
      const double h100 = g10b - vY[c], h000 = -(vY[c] + g10b);
      const double h111 = g01a + g11a, h011 = g01a - g11a;
      const double h101 = g01b + g11b, h001 = g01b - g11b;
      H0[c] = h000 + __shfl_up_sync(FULL, h100, 1);
      H1[c] = h001 + __shfl_up_sync(FULL, h101, 1);
      sX[par][6 + c][ty][tx] = h010 + __shfl_up_sync(FULL, h110, 1);
      sX[par][9 + c][ty][tx] = h011 + __shfl_up_sync(FULL, h111, 1);
    }
    // publish S/D of the next layer (plane ll+2) into the other parity buffer, then ONE barrier
    const double xo_next[3] = {rt[0], rt[1], rt[2]};
    const unsigned char fo_next = ftf;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      vb[c] = vt[c];
      rt[c] = rn[c];
    }
    ftf = fn;
    Ee = En;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      vt[c] = (ftf & (1 << c)) ? 0.0 : rt[c];
      S0[c] = vt[c] + vb[c];
      D0[c] = vt[c] - vb[c];
      sX[par ^ 1][c][ty][tx] = S0[c];
      sX[par ^ 1][3 + c][ty][tx] = D0[c];
    }
    // prefetch plane ll+3 / layer ll+2
    noff += xs;
    fp += g.S;
    Ep += g.SE;
    if (ll + 2 < z1) {
      fn = node_ok ? fp[g.S] : 0;
      load_node(noff + xs, rn);
      const int gl = ll + 2 + g.p0;
      En = (elem_ok && gl >= 0 && gl < g.NLg) ? Ep[0] : 0.0;
    }
    if (NSYNC) {
      // Pairwise rendezvous on hardware named barriers (id k couples rows k-1 and k, 64 threads).
      // Even rows meet their lower neighbour first, odd rows their upper one: the pairs of each
      // round are disjoint, so there is no chain through the CTA and no deadlock.
      // publish "row ty finished step it", then wait until both neighbours have published it.
      // The loop exit is a warp vote, i.e. provably uniform: the shuffles below stay convergent.
      ++it;
      __syncwarp();
      if (tx == 0) {
        __threadfence_block();
        *(volatile int*)&sflag[ty] = it;
      }
      const int lo = ty > 0 ? ty - 1 : 0, hi = ty + 1 < TY ? ty + 1 : TY - 1;
      int spins = 0;
      bool ready;
      do {
        ready = (*(volatile int*)&sflag[lo] >= it && *(volatile int*)&sflag[hi] >= it) || ++spins > (1 << 24);
      } while (!__all_sync(FULL, ready));
      __threadfence_block();
    } else {
      __syncthreads();
    }
    if (ty >= 1) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        H0[c] += sX[par][6 + c][ty - 1][tx];
        H1[c] += sX[par][9 + c][ty - 1][tx];
      }
    }
    if (own && ll >= z0) {
      const long long yo = ((long long)ll * g.S + ncol) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        double v = carry[c] + (H0[c] - H1[c]);  // bottom plane of this layer + top of the previous one
        if (fo & (1 << c)) v = fixed_diag * xo[c];
        y[yo + c] = v;
        if (FUSEP) pnew[yo + c] = xo[c];
        if (DOT) dot = fma(xo[c], v, dot);
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      carry[c] = H0[c] + H1[c];
      xo[c] = xo_next[c];
    }
    fo = fo_next;
  }
  }  // segments
  if (DOT) {
    const double v[1] = {dot};
    block_partials_finish<1>(v, partials, st, fin, sm);
  }
}

// ---- compliance / sensitivity in the modal basis -------------------------------------------------
// u_e' Ke v_e = (T'u_e)' Khat (T'v_e): 45 products instead of 576 (compute_element_energy.jl:18-38,
// thermal_compliance.jl:144-155 for the bilinear form).
__device__ __forceinline__ void hex8_modal7(const double (&b)[4], const double (&t)[4], double (&m)[7]) {
  // columns 0:(0,0) 1:(1,0) 2:(0,1) 3:(1,1); b = bottom plane, t = top plane
  const double s0 = t[0] + b[0], d0 = t[0] - b[0], s1 = t[1] + b[1], d1 = t[1] - b[1];
  const double s2 = t[2] + b[2], d2 = t[2] - b[2], s3 = t[3] + b[3], d3 = t[3] - b[3];
  const double ss0 = s1 + s0, ds0 = s1 - s0, ss1 = s3 + s2, ds1 = s3 - s2;
  const double sd0 = d1 + d0, dd0 = d1 - d0, sd1 = d3 + d2, dd1 = d3 - d2;
  m[0] = ds1 + ds0;  // x
  m[1] = ss1 - ss0;  // y
  m[2] = sd1 + sd0;  // z
  m[3] = ds1 - ds0;  // xy
  m[4] = sd1 - sd0;  // yz
  m[5] = dd1 + dd0;  // xz
  m[6] = dd1 - dd0;  // xyz
}

template <bool BILINEAR>
__global__ void __launch_bounds__(kBlock) k_sens_hex8_modal(Geo g, const double* __restrict__ u, const double* __restrict__ v,
                                                            const double* __restrict__ E, const double* __restrict__ dE,
                                                            double* __restrict__ cell, double* __restrict__ grad, double gsign,
                                                            double* partials, CGState* st) {
  __shared__ double sm[32];
  const long long nel = (long long)g.SE * g.nlay;
  double obj[1] = {0.0};
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < nel; t += (long long)gridDim.x * blockDim.x) {
    const int inl = (int)(t % g.SE);
    const int ll = (int)(t / g.SE) + 1;
    const int i = inl % g.nx, j = inl / g.nx;
    const long long n00 = ((long long)ll * g.S + (long long)j * g.NX + i) * 3;
    const long long dy = (long long)g.NX * 3, dz = (long long)g.S * 3;
    double mu[7][3], mv[7][3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double b[4] = {u[n00 + c], u[n00 + 3 + c], u[n00 + dy + c], u[n00 + dy + 3 + c]};
      const double tp[4] = {u[n00 + dz + c], u[n00 + dz + 3 + c], u[n00 + dz + dy + c], u[n00 + dz + dy + 3 + c]};
      double m[7];
      hex8_modal7(b, tp, m);
#pragma unroll
      for (int k = 0; k < 7; ++k) mu[k][c] = m[k];
      if (BILINEAR) {
        const double b2[4] = {v[n00 + c], v[n00 + 3 + c], v[n00 + dy + c], v[n00 + dy + 3 + c]};
        const double t2[4] = {v[n00 + dz + c], v[n00 + dz + 3 + c], v[n00 + dz + dy + c], v[n00 + dz + dy + 3 + c]};
        hex8_modal7(b2, t2, m);
      }
#pragma unroll
      for (int k = 0; k < 7; ++k) mv[k][c] = m[k];
    }
    double ce = 0.0;
    ce = fma(cKh[0] * mv[0][0], mu[0][0], ce);
    ce = fma(cKh[1] * mv[0][0], mu[1][1], ce);
    ce = fma(cKh[2] * mv[0][0], mu[2][2], ce);
    ce = fma(cKh[3] * mv[1][1], mu[0][0], ce);
    ce = fma(cKh[4] * mv[1][1], mu[1][1], ce);
    ce = fma(cKh[5] * mv[1][1], mu[2][2], ce);
    ce = fma(cKh[6] * mv[2][2], mu[0][0], ce);
    ce = fma(cKh[7] * mv[2][2], mu[1][1], ce);
    ce = fma(cKh[8] * mv[2][2], mu[2][2], ce);
    ce = fma(cKh[9] * mv[0][1], mu[0][1], ce);
    ce = fma(cKh[10] * mv[0][1], mu[1][0], ce);
    ce = fma(cKh[11] * mv[1][0], mu[0][1], ce);
    ce = fma(cKh[12] * mv[1][0], mu[1][0], ce);
    ce = fma(cKh[13] * mv[0][2], mu[0][2], ce);
    ce = fma(cKh[14] * mv[0][2], mu[2][0], ce);
    ce = fma(cKh[15] * mv[2][0], mu[0][2], ce);
    ce = fma(cKh[16] * mv[2][0], mu[2][0], ce);
    ce = fma(cKh[17] * mv[1][2], mu[1][2], ce);
    ce = fma(cKh[18] * mv[1][2], mu[2][1], ce);
    ce = fma(cKh[19] * mv[2][1], mu[1][2], ce);
    ce = fma(cKh[20] * mv[2][1], mu[2][1], ce);
    ce = fma(cKh[21] * mv[3][0], mu[3][0], ce);
    ce = fma(cKh[22] * mv[3][0], mu[4][2], ce);
    ce = fma(cKh[23] * mv[4][2], mu[3][0], ce);
    ce = fma(cKh[24] * mv[4][2], mu[4][2], ce);
    ce = fma(cKh[25] * mv[3][1], mu[3][1], ce);
    ce = fma(cKh[26] * mv[3][1], mu[5][2], ce);
    ce = fma(cKh[27] * mv[5][2], mu[3][1], ce);
    ce = fma(cKh[28] * mv[5][2], mu[5][2], ce);
    ce = fma(cKh[29] * mv[4][1], mu[4][1], ce);
    ce = fma(cKh[30] * mv[4][1], mu[5][0], ce);
    ce = fma(cKh[31] * mv[5][0], mu[4][1], ce);
    ce = fma(cKh[32] * mv[5][0], mu[5][0], ce);
    ce = fma(cKh[33] * mv[3][2], mu[3][2], ce);
    ce = fma(cKh[34] * mv[3][2], mu[4][0], ce);
    ce = fma(cKh[35] * mv[3][2], mu[5][1], ce);
    ce = fma(cKh[36] * mv[4][0], mu[3][2], ce);
    ce = fma(cKh[37] * mv[4][0], mu[4][0], ce);
    ce = fma(cKh[38] * mv[4][0], mu[5][1], ce);
    ce = fma(cKh[39] * mv[5][1], mu[3][2], ce);
    ce = fma(cKh[40] * mv[5][1], mu[4][0], ce);
    ce = fma(cKh[41] * mv[5][1], mu[5][1], ce);
    ce = fma(cKh[42] * mv[6][0], mu[6][0], ce);
    ce = fma(cKh[43] * mv[6][1], mu[6][1], ce);
    ce = fma(cKh[44] * mv[6][2], mu[6][2], ce);
    const long long le = (long long)ll * g.SE + inl;
    if (cell) cell[le] = ce;
    if (grad) grad[le] = gsign * dE[le] * ce;
    obj[0] = fma(E[le], ce, obj[0]);
  }
  block_partials_finish<1>(obj, partials, st, FIN_PLAIN, sm);
}

}  // namespace topopt
