// One CG iteration in one kernel, node planes staged by TENSOR-MAP bulk copies (cp.async.bulk.tensor.2d, SASS UTMALDG).
//
// Same iteration as kxu_hex8_cgfused.cuh (read that header first: recurrence, combine step, ping-pong buffers, x slots,
// halo signalling).  What changes is the staging.  The row-wise bulk copies of that kernel cost ~50 cycles EACH in the
// SM's copy engine regardless of their 800 bytes (issuing every copy as two halves took the kernel from 210 to 308 us;
// a second producer warp gained 4 %), and a plane needs 63 of them.  Here a plane of a tile is SIX requests:
//
//   * a dof vector in local layout is a 2-D array [R = plane * NY + row][3 NX] of doubles whose row pitch 24 NX is not
//     a multiple of 16 bytes when NX is odd (257 at config 4), which a tensor map does not accept.  So each vector gets
//     TWO maps, one over its even rows R and one over its odd rows (pitch 48 NX); the odd map starts 8 bytes early when
//     NX is odd, with its x coordinate shifted by one, so that base and pitch are both 16-byte aligned;
//   * the 2 TYT + 1 node rows of a tile plane are the boxes [100 x (TYT + 1)] at (3 c0, ceil(R0 / 2)) of the even map and
//     (3 c0 + shift, floor(R0 / 2)) of the odd map (x rounded down to an even element: a box must start on a 16-byte
//     boundary of its row; the box is 100 wide for the 99 values); tile row t lands in sub-box (R0 + t) & 1 at row t >> 1.  Columns
//     outside the domain are zero-filled by the copy engine, rows outside wrap into the neighbouring plane (finite
//     values that only reach elements whose modulus is zero);
//   * the ring of plane stages is shared by the CTA (no per-warp copies of rows): row C of a thread row IS row A of the
//     thread row above, the in-place combine needs no second store, and shared memory drops from 222 to 190 KB.
// One elected lane issues everything: per plane 2 copies of p when the ring stage is free (all thread rows released it)
// and 4 copies of r / Ap when the staging slot is free (all thread rows combined its previous plane).
#pragma once
#include <cuda.h>

#include "kxu_hex8_cgfused.cuh"

namespace topopt {

__host__ __device__ constexpr int tma_sub_bytes(int tyt) { return ((tyt + 1) * 800 + 127) / 128 * 128; }  // one sub-box, 128-byte aligned
__host__ __device__ constexpr size_t hex8_cgtma_smem(int tyt, int nst) {
  return (size_t)(nst + 4) * 2 * tma_sub_bytes(tyt) + (size_t)tyt * 2 * 3 * 32 * 8 + sizeof(double) * 2 * 6 * (tyt + 1) * 32;
}

// tensor maps in global memory: [buffer][even, odd]; buffers 0..5 = p0, p1, r0, r1, Ap0, Ap1 of this rank,
// 6..11 the lower neighbour's, 12..17 the upper neighbour's (multi-GPU)
struct CGTmaArgs {
  const CUtensorMap* maps;
  int shift;      // x offset of the odd-row maps (1 when NX is odd)
  int lo_plane;   // lower neighbour's top owned plane in ITS local numbering (-1: no neighbour)
  int hi_plane;   // upper neighbour's first owned plane (1), -1: none
};

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
               : "memory");
}

template <int TYT, int NST, bool CUBE, bool PEER>
__global__ void __launch_bounds__(32 * (TYT + 1), 1)
    k_cg_tma_hex8(Geo g, CGFusedVecs vec, CGTmaArgs ta, const double* __restrict__ E, const unsigned char* __restrict__ fixed,
                  double fixed_diag, int tilesX, int tilesY, double* partials, CGState* st, int fin) {
  static_assert(NST % 2 == 0, "the staging area is two deep: its slot is the ring stage modulo 2");
  constexpr int SUB = tma_sub_bytes(TYT);       // bytes of one sub-box (even or odd rows of a tile plane)
  constexpr int PLANE = 2 * SUB;                // one vector, one tile plane
  constexpr int BOX = (TYT + 1) * 800;          // bytes one tensor copy delivers
  constexpr int S0 = NST * PLANE;               // staging: [slot][r, Ap]
  constexpr int XS0 = S0 + 4 * PLANE;           // x_{k-1} of the next plane: [thread row][row A, B][j][lane]
  constexpr int YB0 = XS0 + TYT * 2 * 3 * 32 * 8;
  constexpr int OWNR = 2 * TYT - 1;
  extern __shared__ __align__(128) unsigned char smem_tma[];
  double(*yb)[6][TYT + 1][32] = reinterpret_cast<double(*)[6][TYT + 1][32]>(smem_tma + YB0);
  // full: the plane's p and r / Ap have landed (two arrivals with byte counts); empty: ring stage released by every
  // thread row (tail); sempty: staging slot released by every thread row (combine)
  __shared__ uint64_t full[NST], empty[NST], sempty[2];
  __shared__ int sflag[TYT], cflag[TYT];
  __shared__ double zero3[4];
  __shared__ double sm[32];
  if (st->done) return;
  const int parity = st->iters & 1;
  const double alpha = st->alpha, beta = st->beta;
  double* __restrict__ pout = vec.p[parity ^ 1];
  double* __restrict__ rout = vec.r[parity ^ 1];
  double* __restrict__ y = vec.ap[parity ^ 1];
  double* __restrict__ xsol = vec.x;
  const int tid = threadIdx.x;
  const int tx = tid & 31, wid = tid >> 5;
  const bool producer = wid == TYT;
  const int ty = producer ? 0 : TYT - 1 - wid;
  const unsigned FULL = 0xffffffffu;
  if (tid < NST) {
    mbar_init(&full[tid], 2);
    mbar_init(&empty[tid], TYT);
    if (tid < 2) mbar_init(&sempty[tid], TYT);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < TYT) sflag[tid] = cflag[tid] = 0;
  if (tid < 4) zero3[tid] = 0.0;
  for (int i = tid; i < YB0 / 16; i += (int)blockDim.x) reinterpret_cast<uint4*>(smem_tma)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < 2 * 6 * (TYT + 1) * 32; i += (int)blockDim.x) (&yb[0][0][0][0])[i] = 0.0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();

  double dots[3] = {0.0, 0.0, 0.0};  // p.Ap, Ap.Ap, r.r
  const long long units = (long long)tilesX * tilesY * g.nown;
  long long u0 = units * blockIdx.x / gridDim.x;
  const long long u1 = units * (blockIdx.x + 1) / gridDim.x;

  // row of the [R][3 NX] view at which plane P of this tile starts, and which maps hold it
  //   (own planes: buffers `parity` of each pair; ghost planes of a slab neighbour: its buffers, its plane numbering)
  auto plane_src = [&](int P, int rfirst, int& mapbase) -> int {
    int Pb = P;
    mapbase = 0;
    if (PEER) {
      if (ta.lo_plane >= 0 && P == 0) {
        Pb = ta.lo_plane;
        mapbase = 12;
      } else if (ta.hi_plane >= 0 && P == g.nown + 1) {
        Pb = ta.hi_plane;
        mapbase = 24;
      }
    }
    return Pb * g.NY + rfirst;
  };

  if (producer) {
    // =========================== producer: one elected lane ===========================
    if (tx == 0) {
      struct Stream {
        long long su0;
        int P, last, c0, rfirst;
        unsigned f;
        bool done;
      } sp{u0, 1, 0, 0, 0, 0u, false}, ss{u0, 1, 0, 0, 0, 0u, false};
      auto next_segment = [&](Stream& S) {
        if (S.su0 >= u1) {
          S.done = true;
          return;
        }
        const int tile = (int)(S.su0 / g.nown);
        const int zoff = (int)(S.su0 % g.nown);
        const int zlen = (int)min((long long)(g.nown - zoff), u1 - S.su0);
        S.su0 += zlen;
        const int bx = tile % tilesX, by = tile / tilesX;
        S.c0 = bx * 31 - 1;
        S.rfirst = by * OWNR - 1;
        S.P = zoff;
        S.last = zoff + zlen + 1;
      };
      long long halo_spins = 0;
      auto halo_ready = [&](int P) -> bool {  // PEER: the neighbour's previous kernel has finished
        if (!PEER) return true;
        const bool glo = ta.lo_plane >= 0 && P == 0, ghi = ta.hi_plane >= 0 && P == g.nown + 1;
        if (!glo && !ghi) return true;
        PeerComm* pc = st->peer;
        const volatile unsigned long long* f = pc->block[pc->rank]->halo_flag;
        if (f[glo ? 0 : 1] < pc->halo_seq) {
          if (++halo_spins > kSpinLimit / 8) {
            pc->timeout = 1;  // a dead peer must not hang the GPU: proceed, the solve reports the error
            return true;
          }
          return false;
        }
        __threadfence_system();
        asm volatile("fence.proxy.async;" ::: "memory");
        return true;
      };
      // two sub-boxes of vector `v` (0 = p, 1 = r, 2 = Ap) of plane S.P into dst
      auto issue = [&](const Stream& S, int v, unsigned char* dst, uint64_t* bar) {
        int mapbase;
        const int R0 = plane_src(S.P, S.rfirst, mapbase);
        const CUtensorMap* m = ta.maps + mapbase + 2 * (2 * v + parity);
        // the x coordinate of a box must be a multiple of 16 bytes (odd coordinates raise "illegal instruction"): start
        // at the even element below, the consumers skip the extra leading double
        tma_load_2d(dst, m, (3 * S.c0) & ~1, (R0 + 1) >> 1, bar);                     // even rows
        tma_load_2d(dst + SUB, m + 1, (3 * S.c0 + ta.shift) & ~1, R0 >> 1, bar);      // odd rows
      };
      while (!(sp.done && ss.done)) {
        bool progressed = false;
        if (!sp.done && sp.P > sp.last) next_segment(sp);
        if (!sp.done) {
          const int s = (int)(sp.f % NST);
          if ((sp.f < NST || mbar_test_wait(&empty[s], ((sp.f / NST) + 1u) & 1u)) && halo_ready(sp.P)) {
            mbar_arrive_expect_tx(&full[s], 2 * BOX);
            issue(sp, 0, smem_tma + s * PLANE, &full[s]);
            sp.P += 1;
            sp.f += 1;
            progressed = true;
          }
        }
        if (!ss.done && ss.P > ss.last) next_segment(ss);
        if (!ss.done) {
          const int s = (int)(ss.f % NST), slot = (int)(ss.f & 1u);
          if ((ss.f < 2 || mbar_test_wait(&sempty[slot], ((ss.f >> 1) + 1u) & 1u)) && halo_ready(ss.P)) {
            mbar_arrive_expect_tx(&full[s], 4 * BOX);
            issue(ss, 1, smem_tma + S0 + slot * 2 * PLANE, &full[s]);
            issue(ss, 2, smem_tma + S0 + slot * 2 * PLANE + PLANE, &full[s]);
            ss.P += 1;
            ss.f += 1;
            progressed = true;
          }
        }
#ifndef TOPOPT_TMA_NOSLEEP
        if (!progressed) __nanosleep(64);
#endif
      }
    }
    __syncwarp();
  } else {
    // =========================== compute warps ===========================
    int it = 0;
    int cc = 0;  // planes combined so far (monotonic across segments; equal on all thread rows)
    int sf = 0;
    unsigned phase = 0;
    auto next_stage = [](int s) { return s + 1 == NST ? 0 : s + 1; };
    while (u0 < u1) {
      const int tile = (int)(u0 / g.nown);
      const int zoff = (int)(u0 % g.nown);
      const int zlen = (int)min((long long)(g.nown - zoff), u1 - u0);
      u0 += zlen;
      const int bx = tile % tilesX, by = tile / tilesX;
      const int rfirst = by * OWNR - 1;
      const int c0 = bx * 31 - 1, r0 = rfirst + 2 * ty;
      const int z0 = 1 + zoff, z1 = z0 + zlen;
      const int first = z0 - 1;
      asm volatile("bar.sync 1, %0;" ::"n"(32 * TYT) : "memory");

      const int col = c0 + tx;
      bool node_ok[2], own[2], el_ok[2];
      int ncol[2], ecol[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int r = r0 + k;
        node_ok[k] = col >= 0 && col < g.NX && r >= 0 && r < g.NY;
        el_ok[k] = col >= 0 && col < g.nx && r >= 0 && r < g.ny;
        ncol[k] = node_ok[k] ? r * g.NX + col : 0;
        ecol[k] = el_ok[k] ? r * g.nx + col : 0;
      }
      own[0] = node_ok[0] && tx >= 1 && ty >= 1;
      own[1] = node_ok[1] && tx >= 1;
      // Offsets of this thread row's node rows A, B, C inside a staged plane: tile row t sits in sub-box (R0 + t) & 1 at
      // row t >> 1, with t = 2 ty, 2 ty + 1, 2 ty + 2
      auto plane_e0 = [&](int P) -> int {
        int mapbase;
        return plane_src(P, rfirst, mapbase) & 1;
      };
      const int lead[2] = {8 * ((3 * c0) & 1), 8 * ((3 * c0 + ta.shift) & 1)};  // boxes start at even x (see the producer)
      auto row_off = [&](int e0, int k) -> int {  // k = 0, 1, 2: rows A, B, C
        const int q = k == 1 ? (e0 ^ 1) : e0;
        return q * SUB + (ty + (k == 2 ? 1 : 0)) * 800 + lead[q];
      };
      // ---- combine bookkeeping: lane tx handles the flat values v = tx + 32 j (v = 3 node + comp) of a 33-node row
      unsigned vmask = 0, omask = 0;  // bit j: value in the domain / value of a node this tile owns (columns 1..31)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int v = tx + 32 * j, n = v / 3, cj = c0 + n;
        const bool ok = v < 99 && cj >= 0 && cj < g.NX;
        if (ok) vmask |= 1u << j;
        if (ok && n >= 1 && n <= 31) omask |= 1u << j;
      }
      bool crow_ok[3], crow_own[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        crow_ok[k] = r0 + k >= 0 && r0 + k < g.NY;
        crow_own[k] = crow_ok[k] && (k == 1 || (k == 0 && ty >= 1));
      }
      const long long ystep = (long long)g.S * 3;
      const long long grow[2] = {((long long)r0 * g.NX + c0) * 3 + tx, ((long long)(r0 + 1) * g.NX + c0) * 3 + tx};

      auto wait_stage = [&](int s, bool hint) {
        const uint32_t par = (phase >> s) & 1u;
        if (!__all_sync(FULL, hint)) {
          bool landed;
          do {
            landed = mbar_try_wait(&full[s], par);
          } while (!__all_sync(FULL, landed));
        }
        phase ^= 1u << s;
      };
      auto test_stage = [&](int s) -> bool { return mbar_test_wait(&full[s], (phase >> s) & 1u); };
      auto release_stage = [&](int s) {
        __syncwarp();
        if (tx == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the stage was written through the generic proxy
          mbar_arrive(&empty[s]);
        }
      };
      auto prefetch_l1 = [](const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); };

      double(*xs)[3][32] = reinterpret_cast<double(*)[3][32]>(smem_tma + XS0 + ty * (2 * 3 * 32 * 8));
      auto issue_x = [&](int P) {
        const bool wr = P >= z0 && P < z1;
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            if (wr && crow_own[k] && ((omask >> j) & 1u))
              asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(&xs[k][j][tx])),
                           "l"(xsol + ((long long)P * ystep + grow[k] + 32 * j))
                           : "memory");
          }
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
      // Combine plane P (ring stage s): p <- (r - alpha Ap) + beta p in place; owned nodes of owned planes also get
      // r, p and x written to global memory.  The top thread row also combines the tile's last row (its row C).
      auto combine = [&](int P, int s) {
        const bool wr = P >= z0 && P < z1;
        const int e0 = plane_e0(P);
        unsigned char* pb = smem_tma + s * PLANE + 8 * tx;
        const unsigned char* rb = smem_tma + S0 + (s & 1) * 2 * PLANE + 8 * tx;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          if (k == 2 && ty != TYT - 1) continue;
          if (!crow_ok[k]) continue;  // warp-uniform
          const int ro = row_off(e0, k);
          double* qp = reinterpret_cast<double*>(pb + ro);
          const double* qr = reinterpret_cast<const double*>(rb + ro);
          const double* qa = reinterpret_cast<const double*>(rb + PLANE + ro);
          const bool wo = k < 2 && wr && crow_own[k];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (!((vmask >> j) & 1u)) continue;
            const double po = qp[32 * j];
            const double rn = fma(-alpha, qa[32 * j], qr[32 * j]);
            const double pn = fma(beta, po, rn);
            qp[32 * j] = pn;
            if (wo && ((omask >> j) & 1u)) {
              const long long gi = (long long)P * ystep + grow[k < 2 ? k : 0] + 32 * j;
              rout[gi] = rn;
              pout[gi] = pn;
              xsol[gi] = fma(alpha, po, xs[k < 2 ? k : 0][j < 3 ? j : 0][tx]);
              dots[2] = fma(rn, rn, dots[2]);
            }
          }
        }
        ++cc;
        __syncwarp();
        if (tx == 0) {
          mbar_arrive(&sempty[s & 1]);  // r / Ap of this plane are consumed by this thread row
          __threadfence_block();
          *(volatile int*)&cflag[ty] = cc;
        }
      };
      // row C of every plane combined so far (= row A of the thread row above) has been combined
      auto wait_rowC = [&]() {
        if (ty + 1 < TYT) {
          bool ready = *(volatile int*)&cflag[ty + 1] >= cc;
          if (!__all_sync(FULL, ready)) {
            int spins = 0;
            do {
              ready = *(volatile int*)&cflag[ty + 1] >= cc || ++spins > (1 << 24);
            } while (!__all_sync(FULL, ready));
          }
          __threadfence_block();
        }
      };

      int s_tail = sf, s_old = sf, s_new = next_stage(sf);
      int e_tail = 0, e_old = plane_e0(first);
      wait_stage(s_old, false);
      combine(first, s_old);  // not an owned plane: no x
      issue_x(first + 1);
      bool hint_new = false;
      double carry[2][3], nA[2][3], nB[2][3];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) carry[r][c] = nA[r][c] = nB[r][c] = 0.0;
      double* yp[2] = {y + ((long long)(first - 1) * g.S + ncol[0]) * 3, y + ((long long)(first - 1) * g.S + ncol[1]) * 3};
      const double* Ep[2] = {E + (long long)first * g.SE + ecol[0], E + (long long)first * g.SE + ecol[1]};
      const unsigned char* fp[2] = {fixed + (long long)(first - 1) * g.S + ncol[0], fixed + (long long)(first - 1) * g.S + ncol[1]};
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (el_ok[k]) prefetch_l1(Ep[k]);
      int par = 0, lo_seen = it;

      auto tail = [&](int L, int parL, const unsigned char (&flraw)[2]) {
        {
          const int lo = ty > 0 ? ty - 1 : 0;
          if (!__all_sync(FULL, lo_seen >= it)) {
            int spins = 0;
            bool ready;
            do {
              ready = *(volatile int*)&sflag[lo] >= it || ++spins > (1 << 24);
            } while (!__all_sync(FULL, ready));
          }
          __threadfence_block();
        }
        double lowc[2][3], xo[2][3];
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
          for (int c = 0; c < 3; ++c) lowc[m][c] = yb[parL][3 * m + c][ty][tx];
        const bool store = L >= z0;
        bool st_ok[2];
        unsigned char fl[2];
        {
          const unsigned char* sb = smem_tma + s_tail * PLANE + 24 * tx;
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            st_ok[r] = store && own[r];
            const double* q = st_ok[r] ? reinterpret_cast<const double*>(sb + row_off(e_tail, r)) : zero3;
#pragma unroll
            for (int c = 0; c < 3; ++c) xo[r][c] = q[c];
            fl[r] = st_ok[r] ? flraw[r] : (unsigned char)7;
          }
        }
        if (L >= first) release_stage(s_tail);
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
          for (int c = 0; c < 3; ++c) nA[m][c] += lowc[m][c];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          double(&n)[2][3] = r == 0 ? nA : nB;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            double v = carry[r][c] + (n[0][c] - n[1][c]);
            carry[r][c] = n[0][c] + n[1][c];
            if (fl[r] & (1 << c)) v = fixed_diag * xo[r][c];
            if (st_ok[r]) yp[r][c] = v;
            dots[0] = fma(xo[r][c], v, dots[0]);
            dots[1] = fma(v, v, dots[1]);
          }
          yp[r] += ystep;
        }
      };

      for (int ll = first; ll < z1; ++ll, par ^= 1) {
        double Ee[2];
        unsigned char flraw[2];
        {
          const int gl = ll + g.p0;
          const bool lay = gl >= 0 && gl < g.NLg;
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            Ee[k] = (el_ok[k] && lay) ? Ep[k][0] : 0.0;
            flraw[k] = (node_ok[k] && ll - 1 >= z0) ? fp[k][0] : (unsigned char)0;
            Ep[k] += g.SE;
            fp[k] += g.S;
            if (el_ok[k] && ll + 1 < z1) prefetch_l1(Ep[k]);
            if (node_ok[k]) prefetch_l1(fp[k]);
          }
        }
        const int e_new = plane_e0(ll + 1);
        wait_stage(s_new, hint_new);
        hint_new = test_stage(next_stage(s_new));
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        combine(ll + 1, s_new);
        issue_x(ll + 2);
        tail(ll - 1, par ^ 1, flraw);
        wait_rowC();
        RowX xr[3];
        {
          const unsigned char* bb = smem_tma + s_old * PLANE + 24 * tx;
          const unsigned char* bt = smem_tma + s_new * PLANE + 24 * tx;
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const double* qb = reinterpret_cast<const double*>(bb + row_off(e_old, k));
            const double* qt = reinterpret_cast<const double*>(bt + row_off(e_new, k));
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const double ab = qb[3 + c] + qb[c], eb = qb[3 + c] - qb[c];
              const double at = qt[3 + c] + qt[c], et = qt[3 + c] - qt[c];
              xr[k].ss[c] = at + ab;
              xr[k].sd[c] = at - ab;
              xr[k].ds[c] = et + eb;
              xr[k].dd[c] = et - eb;
            }
          }
        }
        double WA[2][2][3], WB[2][2][3], WC[2][2][3];
        {
          double wL[2][2][3], wU[2][2][3];
          hex8_core_t<CUBE>(xr[0], xr[1], wL, wU);
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
          hex8_core_t<CUBE>(xr[1], xr[2], wL, wU);
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
        // The y buffer written now (one of two) was read by the thread row above in its tail of the previous step; it
        // has finished that step, because its combine of this step's top plane (wait_rowC above) comes after it.
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            nA[m][c] = (WA[0][m][c] - WA[1][m][c]) + __shfl_up_sync(FULL, WA[0][m][c] + WA[1][m][c], 1);
            nB[m][c] = (WB[0][m][c] - WB[1][m][c]) + __shfl_up_sync(FULL, WB[0][m][c] + WB[1][m][c], 1);
            yb[par][3 * m + c][ty + 1][tx] = (WC[0][m][c] - WC[1][m][c]) + __shfl_up_sync(FULL, WC[0][m][c] + WC[1][m][c], 1);
          }
        ++it;
        __syncwarp();
        if (tx == 0) {
          __threadfence_block();
          *(volatile int*)&sflag[ty] = it;
        }
        lo_seen = *(volatile int*)&sflag[ty > 0 ? ty - 1 : 0];
        s_tail = s_old;
        s_old = s_new;
        s_new = next_stage(s_new);
        e_tail = e_old;
        e_old = e_new;
      }
      {
        unsigned char flraw[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) flraw[k] = node_ok[k] ? fp[k][0] : (unsigned char)0;
        tail(z1 - 1, par ^ 1, flraw);
      }
      release_stage(s_old);
      sf = s_new;
    }  // segments
  }
  block_partials_finish<3>(dots, partials, st, fin, sm, PEER);
}

}  // namespace topopt
