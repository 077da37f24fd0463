// Matrix-free K.u for hex8 elasticity, modal form, node planes staged in a shared-memory ring by the
// bulk-copy engine (cp.async.bulk + mbarrier, SASS UBLKCP).
//
// Same mathematics as kxu_hex8_2row.cuh (two node rows / two elements per thread, z-marching persistent
// CTAs); what changes is how the data moves:
//   * One lane per warp issues 800-byte bulk copies of whole tile rows (33 nodes x 24 B) of the next
//     node planes into a ring of NST plane buffers; completion is tracked by one mbarrier per stage.
//     No per-thread global loads, no prefetch registers, no address arithmetic in the loop.
//   * Every thread reads the RAW values of the six nodes it needs of the new plane (own and x+1 column
//     of its rows A, B and of row C of the thread row above) straight from the ring, so the forward
//     transform needs no shuffles, no shared-memory publication and no synchronisation; the x stage
//     is done first on the raw plane, the z stage against the previous plane's x-staged values kept in
//     registers.  Lane 31's element is valid too (the ring row holds 33 columns): a tile owns 31 of 32
//     columns and 2*TYT-1 of 2*TYT rows.
//   * The penalised modulus E_e is applied when the element's corner forces are accumulated onto the
//     x edges (one DMUL/DFMA per value, the accumulation is free), the inverse x stage runs once per
//     x edge after the two elements sharing it were added, and the trilinear modes are folded into
//     the first inverse butterfly.
//   * The tail of a step (y reduction through shared memory, inverse z stage, store, dot products) is
//     deferred to the start of the next step: by then the neighbouring rows have long published their
//     partial sums, so the neighbour wait does not spin and the tail's latencies overlap the next
//     step's loads.
// Precondition: x is zero on prescribed dofs (true for every CG direction: b, r and p are zero there),
// so bcmatrix column masking is a no-op and only the row rule y[d] = fixed_diag * x[d] is applied.
#pragma once
#include "kxu_hex8_2row.cuh"

namespace topopt {

constexpr int kRingPitch = 848;  // bytes per ring row: 32 (clip shift) + 8 (alignment lead) + 33 * 24, rounded to 16

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// global -> shared bulk copy (16-byte aligned source, destination and size), completes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// z stage of one node row from the x-staged new plane (a = x+1 + own, e = x+1 - own) and the x-staged
// previous plane (pa, pe); the new plane then becomes the previous one.
__device__ __forceinline__ void hex8_zstage(const double (&own)[3], const double (&right)[3], double (&pa)[3], double (&pe)[3],
                                            RowX& o) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double a = right[c] + own[c], e = right[c] - own[c];
    o.ss[c] = a + pa[c];
    o.sd[c] = a - pa[c];
    o.ds[c] = e + pe[c];
    o.dd[c] = e - pe[c];
    pa[c] = a;
    pe[c] = e;
  }
}

// Element between node rows L (oy = 0) and U (oy = 1): forward y stage, modal matrix, inverse y stage.
// wL / wU [mx][mz][comp]: UNSCALED corner-force combinations still to be multiplied by E_e, summed over
// the two elements sharing the x edge and sent through the inverse x stage
//   own column (ox = 0): W[0] - W[1],   x+1 column (ox = 1): W[0] + W[1].
__device__ __forceinline__ void hex8_core(const RowX& L, const RowX& U, double (&wL)[2][2][3], double (&wU)[2][2][3]) {
  double X[3], Y[3], Z[3], XY[3], YZ[3], XZ[3], XYZ[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    X[c] = U.ds[c] + L.ds[c];
    Y[c] = U.ss[c] - L.ss[c];
    Z[c] = U.sd[c] + L.sd[c];
    XY[c] = U.ds[c] - L.ds[c];
    YZ[c] = U.sd[c] - L.sd[c];
    XZ[c] = U.dd[c] + L.dd[c];
    XYZ[c] = U.dd[c] - L.dd[c];
  }
  double vX[3], vY[3], vZ[3], vXY[3], vYZ[3], vXZ[3];
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
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    wU[0][0][c] = vY[c];  // (mx, mz) = (0, 0): +/- vY
    wL[0][0][c] = -vY[c];
    wU[1][0][c] = vX[c] + vXY[c];  // (1, 0)
    wL[1][0][c] = vX[c] - vXY[c];
    wU[0][1][c] = vZ[c] + vYZ[c];  // (0, 1)
    wL[0][1][c] = vZ[c] - vYZ[c];
    wU[1][1][c] = fma(cKh[42 + c], XYZ[c], vXZ[c]);  // (1, 1): trilinear mode folded in
    wL[1][1][c] = fma(-cKh[42 + c], XYZ[c], vXZ[c]);
  }
}

// DOT: 0 = none, 1 = sum x.y, 2 = sum x.y and sum y.y (single-pass CG)
template <int TYT, int NST, int DOT, bool PEER>
__global__ void __launch_bounds__(32 * TYT, 1)
    k_apply_hex8_ring(Geo g, const double* __restrict__ x, double* __restrict__ y, const double* __restrict__ E,
                      const unsigned char* __restrict__ fixed, double fixed_diag, int tilesX, int tilesY, double* partials,
                      CGState* st, int fin, const double* __restrict__ xlo, const double* __restrict__ xhi) {
  constexpr int RR = 2 * TYT + 1;        // ring rows per plane (rows A, B of every thread row + row C of the last)
  constexpr int STAGE = RR * kRingPitch; // bytes per plane
  constexpr int OWNR = 2 * TYT - 1;      // node rows a tile owns
  constexpr int NS = DOT == 2 ? 2 : 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char* ring = smem_raw;
  double(*yb)[6][TYT][32] = reinterpret_cast<double(*)[6][TYT][32]>(smem_raw + NST * STAGE);  // [2] parity buffers
  __shared__ uint64_t full[NST];
  __shared__ int sflag[TYT];
  __shared__ double sm[32];
  if (DOT && st->done) return;
  const int tid = threadIdx.x;
  const int tx = tid & 31, ty = tid >> 5;
  const unsigned FULL = 0xffffffffu;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NST; ++s) mbar_init(&full[s], TYT);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < TYT) sflag[tid] = 0;
  // stale ring contents are read for out-of-domain nodes (their elements carry E = 0): keep them finite
  for (int i = tid; i < NST * STAGE / 16; i += 32 * TYT) reinterpret_cast<uint4*>(ring)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();

  int it = 0;             // published-step counter (monotonic across segments)
  unsigned fbase = 0;     // fills issued before this segment: fill f -> stage f % NST, parity (f / NST) & 1
  double dots[NS];
#pragma unroll
  for (int k = 0; k < NS; ++k) dots[k] = 0.0;

  const long long units = (long long)tilesX * tilesY * g.nown;
  long long u0 = units * blockIdx.x / gridDim.x;
  const long long u1 = units * (blockIdx.x + 1) / gridDim.x;
  while (u0 < u1) {
    const int tile = (int)(u0 / g.nown);
    const int zoff = (int)(u0 % g.nown);
    const int zlen = (int)min((long long)(g.nown - zoff), u1 - u0);
    u0 += zlen;
    const int bx = tile % tilesX, by = tile / tilesX;
    const int c0 = bx * 31 - 1, r0 = by * OWNR - 1;  // node column of lane 0 / node row of ring row 0
    const int z0 = 1 + zoff, z1 = z0 + zlen;         // owned local planes [z0, z1); planes z0-1 .. z1 are read
    const int first = z0 - 1, last = z1;
    const int c_lo = max(c0, 0), c_hi = min(c0 + 33, g.NX), cnt = c_hi - c_lo;
    const int shift = c_lo - c0;              // 1 on the left-edge tile (column -1 does not exist)
    const int dst_off = shift ? 32 : 0;       // keeps lane 0's (unused) slot inside the row buffer
    __syncthreads();                          // the previous segment's ring / yb reads are complete

    const int col = c0 + tx;
    int rrow[3];          // node rows of A, B, C
    unsigned rowpar[3];   // alignment parity of the row's first copied node
    int base_off[3];      // byte offset of this lane's node in the ring row (before the alignment lead)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      rrow[k] = r0 + 2 * ty + k;
      rowpar[k] = (unsigned)(((long long)rrow[k] * g.NX + c_lo) & 1) << 3;
      base_off[k] = (2 * ty + k) * kRingPitch + dst_off - 24 * shift + 24 * tx;
    }
    bool node_ok[2], own[2], el_ok[2];
    long long ncol[2], ecol[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      node_ok[k] = col >= 0 && col < g.NX && rrow[k] >= 0 && rrow[k] < g.NY;
      el_ok[k] = col >= 0 && col < g.nx && rrow[k] >= 0 && rrow[k] < g.ny;
      ncol[k] = node_ok[k] ? (long long)rrow[k] * g.NX + col : 0;
      ecol[k] = el_ok[k] ? (long long)rrow[k] * g.nx + col : 0;
    }
    own[0] = node_ok[0] && tx >= 1 && ty >= 1;  // row A of thread row 0 lacks the tile below
    own[1] = node_ok[1] && tx >= 1;             // lane 0 lacks the element column to its left

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
        asm volatile("fence.proxy.async;" ::: "memory");
      }
      __syncthreads();
    }
    auto plane_ptr = [&](int P) -> const char* {
      if (PEER) {
        if (xlo != nullptr && P == 0) return reinterpret_cast<const char*>(xlo);
        if (xhi != nullptr && P == g.nown + 1) return reinterpret_cast<const char*>(xhi);
      }
      return reinterpret_cast<const char*>(x + (long long)P * g.S * 3);
    };
    // lane 0 of every warp copies its rows A, B (and C for the last thread row) of plane P into stage f % NST
    auto issue_fill = [&](int P, unsigned f) {
      const int stg = (int)(f % NST);
      const char* pp = plane_ptr(P);
      unsigned char* sbase = ring + stg * STAGE;
      const char* src[3];
      uint32_t nb[3], total = 0;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        nb[k] = 0;
        src[k] = pp;
        const bool want = (k < 2 || ty == TYT - 1) && rrow[k] >= 0 && rrow[k] < g.NY;
        if (want) {
          const char* a = pp + 24ll * ((long long)rrow[k] * g.NX + c_lo);
          const uint32_t lead = (uint32_t)(reinterpret_cast<uintptr_t>(a) & 15);
          src[k] = a - lead;
          nb[k] = (lead + 24u * (uint32_t)cnt + 15u) & ~15u;
          total += nb[k];
        }
      }
      mbar_arrive_expect_tx(&full[stg], total);
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (nb[k]) bulk_g2s(sbase + (2 * ty + k) * kRingPitch + dst_off, src[k], nb[k], &full[stg]);
    };
    // raw values of this lane's node and its x+1 neighbour in rows A, B, C of plane P
    auto read_plane = [&](int P, double (&ow)[3][3], double (&rt)[3][3]) {
      const unsigned f = fbase + (unsigned)(P - first);
      const int stg = (int)(f % NST);
      const uint32_t par = (f / NST) & 1u;
      while (!mbar_try_wait(&full[stg], par)) {
      }
      const unsigned pl = (unsigned)(reinterpret_cast<uintptr_t>(plane_ptr(P)) & 8);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double* q = reinterpret_cast<const double*>(ring + stg * STAGE + base_off[k] + (int)(pl ^ rowpar[k]));
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          ow[k][c] = q[c];
          rt[k][c] = q[3 + c];
        }
      }
    };
    auto load_E = [&](int k, int ll) -> double {
      const int gl = ll + g.p0;
      return (el_ok[k] && gl >= 0 && gl < g.NLg) ? E[(long long)ll * g.SE + ecol[k]] : 0.0;
    };

    // ---- prologue: fill the ring, x stage of the first plane
    const int nplanes = zlen + 2;
    if (tx == 0) {
      for (int f = 0; f < NST && f < nplanes; ++f) issue_fill(first + f, fbase + (unsigned)f);
    }
    __syncwarp();
    double pa[3][3], pe[3][3];  // x-staged previous plane, rows A, B, C
    {
      double ow[3][3], rt[3][3];
      read_plane(first, ow, rt);
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          pa[k][c] = rt[k][c] + ow[k][c];
          pe[k][c] = rt[k][c] - ow[k][c];
        }
    }
    double carry[2][3];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) carry[r][c] = 0.0;
    double En[2] = {load_E(0, first), load_E(1, first)};

    // deferred tail state: x-reduced corner forces of rows A, B of the previous step
    double nA[2][3], nB[2][3];
    int par = 0;

    // tail of step L: y reduction, inverse z stage, store of plane L, refill of plane L's ring stage
    auto tail = [&](int L, int parL) {
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
          for (int c = 0; c < 3; ++c) nA[m][c] += yb[parL][3 * m + c][ty - 1][tx];
      }
      if (L >= z0) {
        // raw x of the plane being stored (prescribed rows, dot products): still in the ring
        const unsigned f = fbase + (unsigned)(L - first);
        const int stg = (int)(f % NST);
        const unsigned pl = (unsigned)(reinterpret_cast<uintptr_t>(plane_ptr(L)) & 8);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          double(&n)[2][3] = r == 0 ? nA : nB;
          const double* q = reinterpret_cast<const double*>(ring + stg * STAGE + base_off[r] + (int)(pl ^ rowpar[r]));
          const unsigned char fl = node_ok[r] ? fixed[(long long)L * g.S + ncol[r]] : 0;
          if (own[r]) {
            const long long yo = ((long long)L * g.S + ncol[r]) * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const double xo = q[c];
              double v = carry[r][c] + (n[0][c] - n[1][c]);
              if (fl & (1 << c)) v = fixed_diag * xo;  // prescribed row: meandiag * x
              y[yo + c] = v;
              if (DOT >= 1) dots[0] = fma(xo, v, dots[0]);
              if (DOT == 2) dots[1] = fma(v, v, dots[1]);
            }
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        carry[0][c] = nA[0][c] + nA[1][c];
        carry[1][c] = nB[0][c] + nB[1][c];
      }
      // plane L is no longer needed by this thread row nor (as its row C) by the row below: refill its stage
      if (L + NST <= last) {
        if (tx == 0) issue_fill(L + NST, fbase + (unsigned)(L + NST - first));
        __syncwarp();
      }
    };

    for (int ll = first; ll < z1; ++ll, par ^= 1) {  // element layer ll: planes ll (bottom), ll + 1 (top)
      double ow[3][3], rt[3][3];
      read_plane(ll + 1, ow, rt);
      const double Ee[2] = {En[0], En[1]};
      if (ll + 1 < z1) {
        En[0] = load_E(0, ll + 1);
        En[1] = load_E(1, ll + 1);
      }
      if (ll > first) tail(ll - 1, par ^ 1);
      RowX xa, xb, xc;
      hex8_zstage(ow[0], rt[0], pa[0], pe[0], xa);
      hex8_zstage(ow[1], rt[1], pa[1], pe[1], xb);
      hex8_zstage(ow[2], rt[2], pa[2], pe[2], xc);
      // ---- the two elements; E-scaled accumulation onto the x edges of rows A, B, C
      double WA[2][2][3], WB[2][2][3], WC[2][2][3];
      {
        double wL[2][2][3], wU[2][2][3];
        hex8_core(xa, xb, wL, wU);
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              WA[a][m][c] = Ee[0] * wL[a][m][c];
              WB[a][m][c] = Ee[0] * wU[a][m][c];
            }
      }
      {
        double wL[2][2][3], wU[2][2][3];
        hex8_core(xb, xc, wL, wU);
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              WB[a][m][c] = fma(Ee[1], wL[a][m][c], WB[a][m][c]);
              WC[a][m][c] = Ee[1] * wU[a][m][c];
            }
      }
      // ---- inverse x stage per x edge, reduction over x by shuffle; row C goes to the thread row above
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          nA[m][c] = (WA[0][m][c] - WA[1][m][c]) + __shfl_up_sync(FULL, WA[0][m][c] + WA[1][m][c], 1);
          nB[m][c] = (WB[0][m][c] - WB[1][m][c]) + __shfl_up_sync(FULL, WB[0][m][c] + WB[1][m][c], 1);
          yb[par][3 * m + c][ty][tx] = (WC[0][m][c] - WC[1][m][c]) + __shfl_up_sync(FULL, WC[0][m][c] + WC[1][m][c], 1);
        }
      ++it;
      __syncwarp();
      if (tx == 0) {
        __threadfence_block();
        *(volatile int*)&sflag[ty] = it;
      }
    }
    tail(z1 - 1, par ^ 1);
    fbase += (unsigned)nplanes;
  }  // segments
  if (DOT) block_partials_finish<NS>(dots, partials, st, fin, sm);
}

}  // namespace topopt
