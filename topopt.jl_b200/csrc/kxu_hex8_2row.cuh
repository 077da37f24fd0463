// Matrix-free K.u for hex8 elasticity, modal form, TWO node rows / element rows per thread.
//
// Same mathematics and data flow as kxu_hex8.cuh (read that header first); the difference is the
// register blocking along y: thread (tx,ty) owns node rows A = 2*ty and B = 2*ty+1 of the CTA's
// patch and the two element rows above them.  Row B is shared by both of the thread's elements
// without any exchange, so per element the kernel issues ~30 % fewer instructions (one x-stage
// per node row instead of two, 36 instead of 48 shuffles, half the shared-memory hops and half the
// per-thread bookkeeping) at the same fp64 work.
#pragma once
#include "kxu_hex8.cuh"

namespace topopt {

struct RowX {  // x stage of the forward Hadamard for one node row: (x+1) +/- own, for S and D
  double ss[3], ds[3], sd[3], dd[3];
};

__device__ __forceinline__ void hex8_xstage(const double (&S0)[3], const double (&D0)[3], RowX& o) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double S1 = __shfl_down_sync(0xffffffffu, S0[c], 1);
    const double D1 = __shfl_down_sync(0xffffffffu, D0[c], 1);
    o.ss[c] = S1 + S0[c];
    o.ds[c] = S1 - S0[c];
    o.sd[c] = D1 + D0[c];
    o.dd[c] = D1 - D0[c];
  }
}

// One element between node rows L (oy = 0) and U (oy = 1).  Corner forces before the z stage:
//   n_lo / s_lo : own column (ox = 0) / x+1 column (ox = 1) of row L,  [mz][comp]
//   n_hi / s_hi : the same for row U.
// ADD_LO: row L already holds the contribution of the element below it.
template <bool ADD_LO>
__device__ __forceinline__ void hex8_element(const RowX& L, const RowX& U, double Ee, double (&n_lo)[2][3],
                                             double (&s_lo)[2][3], double (&n_hi)[2][3], double (&s_hi)[2][3]) {
  double X[3], Y[3], Z[3], XY[3], YZ[3], XZ[3], XYZ[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    X[c] = Ee * (U.ds[c] + L.ds[c]);
    Y[c] = Ee * (U.ss[c] - L.ss[c]);
    Z[c] = Ee * (U.sd[c] + L.sd[c]);
    XY[c] = Ee * (U.ds[c] - L.ds[c]);
    YZ[c] = Ee * (U.sd[c] - L.sd[c]);
    XZ[c] = Ee * (U.dd[c] + L.dd[c]);
    XYZ[c] = Ee * (U.dd[c] - L.dd[c]);
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
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double g10a = vX[c] + vXY[c], g10b = vX[c] - vXY[c];      // (mx=1,mz=0): oy=1, oy=0
    const double g01a = vZ[c] + vYZ[c], g01b = vZ[c] - vYZ[c];      // (mx=0,mz=1)
    const double g11a = vXZ[c] + vXYZ[c], g11b = vXZ[c] - vXYZ[c];  // (mx=1,mz=1)
    const double h110 = vY[c] + g10a, h010 = vY[c] - g10a;          // h[ox][oy][mz]
    const double h100 = g10b - vY[c], h000 = -(vY[c] + g10b);
    const double h111 = g01a + g11a, h011 = g01a - g11a;
    const double h101 = g01b + g11b, h001 = g01b - g11b;
    if (ADD_LO) {
      n_lo[0][c] += h000;
      n_lo[1][c] += h001;
      s_lo[0][c] += h100;
      s_lo[1][c] += h101;
    } else {
      n_lo[0][c] = h000;
      n_lo[1][c] = h001;
      s_lo[0][c] = h100;
      s_lo[1][c] = h101;
    }
    n_hi[0][c] = h010;
    n_hi[1][c] = h011;
    s_hi[0][c] = h110;
    s_hi[1][c] = h111;
  }
}

template <int TYT, bool DOT, bool PEER>
__global__ void __launch_bounds__(32 * TYT, 1)
    k_apply_hex8_modal2(Geo g, const double* __restrict__ x, double* __restrict__ y, const double* __restrict__ E,
                        const unsigned char* __restrict__ fixed, double fixed_diag, int tilesX, int tilesY, double* partials,
                        CGState* st, int fin, const double* __restrict__ xlo, const double* __restrict__ xhi) {
  extern __shared__ double smem_dyn[];
  // two parity buffers [12][TYT][32]: rows 0-5 = S/D of this thread's row A for the next layer (read by
  // the thread row below as its row C), rows 6-11 = partial node sums of row C (read by the row above)
  double(*sX)[12][TYT][32] = reinterpret_cast<double(*)[12][TYT][32]>(smem_dyn);
  __shared__ double sm[32];
  __shared__ int sflag[TYT];
  int it = 0;
  if (DOT && st->done) return;
  if (threadIdx.x < TYT) sflag[threadIdx.x] = 0;
  const int tid = threadIdx.x;
  const int tx = tid & 31, ty = tid >> 5;
  const unsigned FULL = 0xffffffffu;
  constexpr int ROWS = 2 * TYT - 2;  // owned node rows per tile
  double dot = 0.0;
  const long long units = (long long)tilesX * tilesY * g.nown;
  long long u0 = units * blockIdx.x / gridDim.x;
  const long long u1 = units * (blockIdx.x + 1) / gridDim.x;
  while (u0 < u1) {
    const int tile = (int)(u0 / g.nown);
    const int zoff = (int)(u0 % g.nown);
    const int zlen = (int)min((long long)(g.nown - zoff), u1 - u0);
    u0 += zlen;
    const int bx = tile % tilesX, by = tile / tilesX;
    const int in = bx * 30 - 1 + tx;
    const int jn[2] = {by * ROWS - 1 + 2 * ty, by * ROWS + 2 * ty};  // node rows A, B
    const int z0 = 1 + zoff, z1 = z0 + zlen;
    __syncthreads();
    bool node_ok[2], own[2], elem_ok[2];
    long long ncol[2], ecol[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      node_ok[r] = in >= 0 && in < g.NX && jn[r] >= 0 && jn[r] < g.NY;
      elem_ok[r] = in >= 0 && in < g.nx && jn[r] >= 0 && jn[r] < g.ny && tx < 31;
      ncol[r] = node_ok[r] ? (long long)jn[r] * g.NX + in : 0;
      ecol[r] = elem_ok[r] ? (long long)jn[r] * g.nx + in : 0;
    }
    elem_ok[1] = elem_ok[1] && ty < TYT - 1;  // element B needs row C of the thread row above
    own[0] = node_ok[0] && tx >= 1 && tx <= 30 && ty >= 1;
    own[1] = node_ok[1] && tx >= 1 && tx <= 30 && ty <= TYT - 2;

    if (PEER && (xlo != nullptr || xhi != nullptr)) {
      const bool need_lo = xlo != nullptr && z0 == 1, need_hi = xhi != nullptr && z1 == g.nown + 1;
      if ((need_lo || need_hi) && tid == 0) {
        PeerComm* pc = st->peer;
        const unsigned long long want = pc->halo_seq;
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
    // raw node value of row r at local plane lp (ghost planes come from the neighbours' memory)
    auto load_node = [&](int r, int lp, double (&v)[3]) {
      const long long off = ((long long)lp * g.S + ncol[r]) * 3;
      const double* src = x + off;
      bool remote = false;
      if (PEER) {
        if (xlo != nullptr && lp == 0) {
          src = xlo + ncol[r] * 3;
          remote = true;
        }
        if (xhi != nullptr && lp == g.nown + 1) {
          src = xhi + ncol[r] * 3;
          remote = true;
        }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = node_ok[r] ? (remote ? __ldcg(src + c) : src[c]) : 0.0;
    };
    auto load_flag = [&](int r, int lp) -> unsigned char { return node_ok[r] ? fixed[(long long)lp * g.S + ncol[r]] : 0; };
    auto load_E = [&](int r, int ll) -> double {
      const int gl = ll + g.p0;
      return (elem_ok[r] && gl >= 0 && gl < g.NLg) ? E[(long long)ll * g.SE + ecol[r]] : 0.0;
    };

    // per row: vt = masked value of the current top plane (next step's bottom), S0/D0 = z stage of
    // the current layer, carry = top-plane force of the previous layer, rn/fn = prefetched next plane
    double vt[2][3], S0[2][3], D0[2][3], carry[2][3], rn[2][3], Ee[2], En[2];
    unsigned char fb[2], ft[2], fn[2];  // flags of the bottom / top / prefetched plane
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      double vb[3], rt[3];
      load_node(r, z0 - 1, vb);
      load_node(r, z0, rt);
      fb[r] = load_flag(r, z0 - 1);
      ft[r] = load_flag(r, z0);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double b = (fb[r] & (1 << c)) ? 0.0 : vb[c];
        vt[r][c] = (ft[r] & (1 << c)) ? 0.0 : rt[c];
        S0[r][c] = vt[r][c] + b;
        D0[r][c] = vt[r][c] - b;
        carry[r][c] = 0.0;
      }
      Ee[r] = load_E(r, z0 - 1);
      load_node(r, z0 + 1, rn[r]);  // z0 + 1 <= nown + 1: inside the allocation
      fn[r] = load_flag(r, z0 + 1);
      En[r] = load_E(r, z0);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      sX[0][c][ty][tx] = S0[0][c];
      sX[0][3 + c][ty][tx] = D0[0][c];
    }
    __syncthreads();

    int par = 0;
    for (int ll = z0 - 1; ll < z1; ++ll, par ^= 1) {
      // ---- x stage of rows A, B (own) and C (row A of the thread row above)
      RowX xa, xb, xc;
      hex8_xstage(S0[0], D0[0], xa);
      hex8_xstage(S0[1], D0[1], xb);
      {
        double SC[3], DC[3];
        const int tyc = ty + 1 < TYT ? ty + 1 : ty;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          SC[c] = sX[par][c][tyc][tx];
          DC[c] = sX[par][3 + c][tyc][tx];
        }
        hex8_xstage(SC, DC, xc);
      }
      // ---- the two elements; corner forces collected per node row
      double nA[2][3], sA[2][3], nB[2][3], sB[2][3], nC[2][3], sC[2][3];
      hex8_element<false>(xa, xb, Ee[0], nA, sA, nB, sB);
      hex8_element<true>(xb, xc, Ee[1], nB, sB, nC, sC);
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          nA[m][c] += __shfl_up_sync(FULL, sA[m][c], 1);
          nB[m][c] += __shfl_up_sync(FULL, sB[m][c], 1);
          sX[par][6 + 3 * m + c][ty][tx] = nC[m][c] + __shfl_up_sync(FULL, sC[m][c], 1);
        }
      // ---- advance the planes: z stage of the next layer, publish row A, prefetch
      unsigned char fcur[2];  // flags of the plane being stored (bottom of this layer)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        fcur[r] = fb[r];
        fb[r] = ft[r];
        ft[r] = fn[r];
        Ee[r] = En[r];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const double b = vt[r][c];
          vt[r][c] = (ft[r] & (1 << c)) ? 0.0 : rn[r][c];
          S0[r][c] = vt[r][c] + b;
          D0[r][c] = vt[r][c] - b;
        }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        sX[par ^ 1][c][ty][tx] = S0[0][c];
        sX[par ^ 1][3 + c][ty][tx] = D0[0][c];
      }
      if (ll + 2 < z1) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          load_node(r, ll + 3, rn[r]);
          fn[r] = load_flag(r, ll + 3);
          En[r] = load_E(r, ll + 2);
        }
      }
      // raw x of the plane being stored (p.Ap and prescribed rows): issued here so that the L1
      // latency overlaps the neighbour wait; few registers are live at this point
      double xd[2][3];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const bool need = own[r] && ll >= z0 && (DOT || fcur[r]);
        const double* xp = x + ((long long)ll * g.S + ncol[r]) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) xd[r][c] = need ? xp[c] : 0.0;
      }
      // ---- neighbour-only synchronisation (see kxu_hex8.cuh)
      ++it;
      __syncwarp();
      if (tx == 0) {
        __threadfence_block();
        *(volatile int*)&sflag[ty] = it;
      }
      {
        const int lo = ty > 0 ? ty - 1 : 0, hi = ty + 1 < TYT ? ty + 1 : TYT - 1;
        int spins = 0;
        bool ready;
        do {
          ready = (*(volatile int*)&sflag[lo] >= it && *(volatile int*)&sflag[hi] >= it) || ++spins > (1 << 24);
        } while (!__all_sync(FULL, ready));
        __threadfence_block();
      }
      if (ty >= 1) {
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
          for (int c = 0; c < 3; ++c) nA[m][c] += sX[par][6 + 3 * m + c][ty - 1][tx];
      }
      // ---- inverse z stage per node row, store the bottom plane of this layer
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        double(&n)[2][3] = r == 0 ? nA : nB;
        if (own[r] && ll >= z0) {
          const long long yo = ((long long)ll * g.S + ncol[r]) * 3;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            double v = carry[r][c] + (n[0][c] - n[1][c]);
            if (fcur[r] & (1 << c)) v = fixed_diag * xd[r][c];  // prescribed row: meandiag * x (raw value)
            y[yo + c] = v;
            if (DOT) dot = fma(xd[r][c], v, dot);
          }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) carry[r][c] = n[0][c] + n[1][c];
      }
    }
  }  // segments
  if (DOT) {
    const double v[1] = {dot};
    block_partials_finish<1>(v, partials, st, fin, sm);
  }
}

}  // namespace topopt
