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

// Cubic cells (hx = hy = hz) with an isotropic material: the 45 modal coefficients collapse to 8 distinct values
//   normal block  [a b b; b a b; b b a]      shear pairs  c [1 1; 1 1]      bilinear pairs  [p q; q p]
//   bilinear triple  [e f f; f e f; f f e]   trilinear diagonal  g
// (verified numerically at topopt_create).  cKc = {a - b, b, c, p, q, e - f, f, g}: 30 instead of 42 fp64 instructions
// for the modal product and few enough constants to stay in uniform registers for the whole kernel.
__constant__ double cKc[8];

template <bool CUBE>
__device__ __forceinline__ void hex8_core_t(const RowX& L, const RowX& U, double (&wL)[2][2][3], double (&wU)[2][2][3]) {
  if constexpr (!CUBE) {
    hex8_core(L, U, wL, wU);
  } else {
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
  const double amb = cKc[0], b = cKc[1], cs = cKc[2], pp = cKc[3], qq = cKc[4], emf = cKc[5], ff = cKc[6], gg = cKc[7];
  double vX[3], vY[3], vZ[3], vXY[3], vYZ[3], vXZ[3];
  {  // normal block on (X.x, Y.y, Z.z)
    const double t = b * ((X[0] + Y[1]) + Z[2]);
    vX[0] = fma(amb, X[0], t);
    vY[1] = fma(amb, Y[1], t);
    vZ[2] = fma(amb, Z[2], t);
  }
  vX[1] = vY[0] = cs * (X[1] + Y[0]);  // shear pairs {X.y, Y.x}, {X.z, Z.x}, {Y.z, Z.y}
  vX[2] = vZ[0] = cs * (X[2] + Z[0]);
  vY[2] = vZ[1] = cs * (Y[2] + Z[1]);
  vXY[0] = fma(qq, YZ[2], pp * XY[0]);  // bilinear pairs {XY.x, YZ.z}, {XY.y, XZ.z}, {YZ.y, XZ.x}
  vYZ[2] = fma(pp, YZ[2], qq * XY[0]);
  vXY[1] = fma(qq, XZ[2], pp * XY[1]);
  vXZ[2] = fma(pp, XZ[2], qq * XY[1]);
  vYZ[1] = fma(qq, XZ[0], pp * YZ[1]);
  vXZ[0] = fma(pp, XZ[0], qq * YZ[1]);
  {  // bilinear triple on (XY.z, YZ.x, XZ.y)
    const double t = ff * ((XY[2] + YZ[0]) + XZ[1]);
    vXY[2] = fma(emf, XY[2], t);
    vYZ[0] = fma(emf, YZ[0], t);
    vXZ[1] = fma(emf, XZ[1], t);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    wU[0][0][c] = vY[c];
    wL[0][0][c] = -vY[c];
    wU[1][0][c] = vX[c] + vXY[c];
    wL[1][0][c] = vX[c] - vXY[c];
    wU[0][1][c] = vZ[c] + vYZ[c];
    wL[0][1][c] = vZ[c] - vYZ[c];
    wU[1][1][c] = fma(gg, XYZ[c], vXZ[c]);
    wL[1][1][c] = fma(-gg, XYZ[c], vXZ[c]);
  }
  }
}

// One stage of a warp's private ring: node rows A, B, C of a plane (E_e and the prescribed-dof flags, 10 bytes per
// thread and step, come through ordinary loads issued one step ahead).
constexpr int kRingStage = 3 * kRingPitch;
constexpr int kRingJobs = 3;  // bulk copies per compute warp and plane

__host__ __device__ constexpr size_t hex8_ring_smem(int tyt, int nst) {
  return (size_t)tyt * nst * kRingStage + sizeof(double) * 3 * 6 * (tyt + 1) * 32;
}

__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Warp-specialised: TYT compute warps (thread rows) + ONE producer warp that issues every bulk copy of the
// CTA (the TMA producer / consumer pipeline: full[w][s] = data landed, empty[w][s] = stage released).
// DOT: 0 = none, 1 = sum x.y, 2 = sum x.y and sum y.y (single-pass CG)
template <int TYT, int NST, int DOT, bool PEER, bool CUBE>
__global__ void __launch_bounds__(32 * (TYT + 1), (TYT <= 5 ? 2 : 1))
    k_apply_hex8_ring(Geo g, const double* __restrict__ x, double* __restrict__ y, const double* __restrict__ E,
                      const unsigned char* __restrict__ fixed, double fixed_diag, int tilesX, int tilesY, double* partials,
                      CGState* st, int fin, const double* __restrict__ xlo, const double* __restrict__ xhi) {
  constexpr int WRING = NST * kRingStage;  // bytes of one warp's private ring
  constexpr int OWNR = 2 * TYT - 1;        // node rows a tile owns
  constexpr int NS = DOT == 2 ? 2 : 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // y exchange: [step % 3][value][thread row + 1][lane]; row 0 stays zero (the bottom thread row has no lower neighbour)
  double(*yb)[6][TYT + 1][32] = reinterpret_cast<double(*)[6][TYT + 1][32]>(smem_raw + TYT * WRING);
  __shared__ uint64_t full[TYT][NST], empty[TYT][NST];
  __shared__ int sflag[TYT];
  __shared__ double zero3[4];  // what non-owned lanes read as their x: keeps the tail branch-free
  __shared__ double sm[32];
  if (DOT && st->done) return;
  const int tid = threadIdx.x;
  const int tx = tid & 31, wid = tid >> 5;
  const bool producer = wid == TYT;
  // the warp scheduler favours high warp ids: the bottom thread row, which every other row depends on
  // through the y reduction chain, gets the highest compute warp id (the producer sits above it)
  const int ty = producer ? 0 : TYT - 1 - wid;
  const unsigned FULL = 0xffffffffu;
  if (tid < TYT * NST) {
    mbar_init(&full[tid / NST][tid % NST], kRingJobs);
    mbar_init(&empty[tid / NST][tid % NST], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < TYT) sflag[tid] = 0;
  if (tid < 4) zero3[tid] = 0.0;
  // stale ring contents are read for out-of-domain nodes (their elements get E = 0): keep them finite
  for (int i = tid; i < TYT * WRING / 16; i += 32 * (TYT + 1)) reinterpret_cast<uint4*>(smem_raw)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < 3 * 6 * (TYT + 1) * 32; i += 32 * (TYT + 1)) (&yb[0][0][0][0])[i] = 0.0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();

  double dots[NS];
#pragma unroll
  for (int k = 0; k < NS; ++k) dots[k] = 0.0;
  const long long units = (long long)tilesX * tilesY * g.nown;
  long long u0 = units * blockIdx.x / gridDim.x;
  const long long u1 = units * (blockIdx.x + 1) / gridDim.x;

  if (producer) {
    // =========================== producer warp ===========================
    // Every lane is an independent state machine for one (compute warp, node row) job: it walks the same segment list
    // as the compute warps and refills a ring stage as soon as THAT warp has released it, so the rows never wait for
    // each other through the producer.
    constexpr int NJ = kRingJobs * TYT, JPL = (NJ + 31) / 32;
    struct Job {
      long long su0;      // first unit of the next segment
      long long idx;      // byte offset of the row inside a plane
      int P, last;        // next plane to fill, last plane of the current segment
      int dst, w, k;      // ring offset, compute warp, row
      unsigned fcount;    // fills issued so far
      uint32_t bytes;
      bool valid, done, ok;
    } job[JPL];
#pragma unroll
    for (int q = 0; q < JPL; ++q) {
      const int j = tx + 32 * q;
      job[q].valid = j < NJ;
      job[q].done = !job[q].valid;
      job[q].w = j / kRingJobs;
      job[q].k = j % kRingJobs;
      job[q].su0 = u0;
      job[q].P = 1;
      job[q].last = 0;  // "segment exhausted": the first pass sets up the first segment
      job[q].fcount = 0;
      job[q].idx = 0;
      job[q].dst = 0;
      job[q].bytes = 0;
      job[q].ok = false;
    }
    long long halo_spins = 0;
    bool alldone;
    do {
      bool progressed = false;
      alldone = true;
#pragma unroll
      for (int q = 0; q < JPL; ++q) {
        Job& J = job[q];
        if (J.done) continue;
        if (J.P > J.last) {  // next segment
          if (J.su0 >= u1) {
            J.done = true;
            continue;
          }
          const int tile = (int)(J.su0 / g.nown);
          const int zoff = (int)(J.su0 % g.nown);
          const int zlen = (int)min((long long)(g.nown - zoff), u1 - J.su0);
          J.su0 += zlen;
          const int bx = tile % tilesX, by = tile / tilesX;
          const int c0 = bx * 31 - 1;
          const int c_lo = max(c0, 0), cnt = min(c0 + 33, g.NX) - c_lo;
          const int frow = by * OWNR - 1 + 2 * J.w + J.k;
          J.P = zoff;  // planes zoff .. zoff + zlen + 1
          J.last = zoff + zlen + 1;
          J.ok = frow >= 0 && frow < g.NY;
          J.idx = 24ll * ((long long)frow * g.NX + c_lo);
          J.dst = J.w * WRING + J.k * kRingPitch + (c_lo != c0 ? 32 : 0);
          J.bytes = 24u * (uint32_t)cnt;
        }
        alldone = false;
        const int sidx = (int)(J.fcount % NST);
        bool ready = J.fcount < NST || mbar_test_wait(&empty[J.w][sidx], ((J.fcount / NST) + 1u) & 1u);
        const char* pp = reinterpret_cast<const char*>(x + (long long)J.P * g.S * 3);
        if (PEER) {
          // ghost planes come straight from the slab neighbours' memory, once their "direction vector final" flag is up
          const bool glo = xlo != nullptr && J.P == 0, ghi = xhi != nullptr && J.P == g.nown + 1;
          if (glo || ghi) {
            PeerComm* pc = st->peer;
            const volatile unsigned long long* f = pc->block[pc->rank]->halo_flag;
            if (f[glo ? 0 : 1] < pc->halo_seq) {
              if (++halo_spins > kSpinLimit / 8) {
                pc->timeout = 1;  // a dead peer must not hang the GPU: proceed, the solve reports the error
              } else {
                ready = false;
              }
            }
            pp = reinterpret_cast<const char*>(glo ? xlo : xhi);
            if (ready) {
              __threadfence_system();
              asm volatile("fence.proxy.async;" ::: "memory");
            }
          }
        }
        if (ready) {
          uint64_t* bar = &full[J.w][sidx];
          if (J.ok) {
            const char* a = pp + J.idx;
            const uint32_t lead = (uint32_t)(reinterpret_cast<uintptr_t>(a) & 15);
            const uint32_t nb = (lead + J.bytes + 15u) & ~15u;
            mbar_arrive_expect_tx(bar, nb);
            bulk_g2s(smem_raw + J.dst + sidx * kRingStage, a - lead, nb, bar);
          } else {
            mbar_arrive(bar);
          }
          J.P += 1;
          J.fcount += 1;
          progressed = true;
        }
      }
      if (!__any_sync(FULL, progressed)) __nanosleep(100);
    } while (!__all_sync(FULL, alldone));
  } else {
    // =========================== compute warps ===========================
    unsigned char* ring = smem_raw + ty * WRING;  // this warp's ring
    int it = 0;          // published-step counter (monotonic across segments)
    int sf = 0;          // ring stage of the next plane to be consumed
    unsigned phase = 0;  // bit s: parity to wait for on full[ty][s]
    auto next_stage = [](int s) { return s + 1 == NST ? 0 : s + 1; };
    while (u0 < u1) {
      const int tile = (int)(u0 / g.nown);
      const int zoff = (int)(u0 % g.nown);
      const int zlen = (int)min((long long)(g.nown - zoff), u1 - u0);
      u0 += zlen;
      const int bx = tile % tilesX, by = tile / tilesX;
      const int c0 = bx * 31 - 1, r0 = by * OWNR - 1 + 2 * ty;  // node column of lane 0 / node row of this warp's row A
      const int z0 = 1 + zoff, z1 = z0 + zlen;                  // owned local planes [z0, z1); planes z0-1 .. z1 are read
      const int first = z0 - 1;
      const int c_lo = max(c0, 0);
      const int shift = c_lo - c0;  // 1 on the left-edge tile (column -1 does not exist)
      asm volatile("bar.sync 1, %0;" ::"n"(32 * TYT) : "memory");  // compute warps only: the previous segment's yb reads are complete

      const int col = c0 + tx;
      const int lane_off = (shift ? 32 - 24 : 0) + 24 * tx;  // byte offset of this lane's node inside a ring row (before the lead)
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
      own[0] = node_ok[0] && tx >= 1 && ty >= 1;  // row A of thread row 0 lacks the tile below
      own[1] = node_ok[1] && tx >= 1;             // lane 0 lacks the element column to its left
      // alignment leads (address & 15) of the copied rows; the plane-dependent part is added per step
      const unsigned leadN0 = (unsigned)(((long long)r0 * g.NX + c_lo) & 1) << 3;  // rows A, C
      const unsigned leadNB = leadN0 ^ ((unsigned)(g.NX & 1) << 3);                // row B

      auto plane_lead = [&](int P) -> unsigned {
        const char* pp = reinterpret_cast<const char*>(x + (long long)P * g.S * 3);
        if (PEER) {
          if (xlo != nullptr && P == 0) pp = reinterpret_cast<const char*>(xlo);
          if (xhi != nullptr && P == g.nown + 1) pp = reinterpret_cast<const char*>(xhi);
        }
        return (unsigned)(reinterpret_cast<uintptr_t>(pp) & 8);
      };
      // `hint`: result of a test issued one step earlier (its latency is hidden behind that step's work)
      auto wait_stage = [&](int s, bool hint) {
        const uint32_t par = (phase >> s) & 1u;
        if (!__all_sync(FULL, hint)) {
          bool landed;
          do {
            landed = mbar_try_wait(&full[ty][s], par);
          } while (!__all_sync(FULL, landed));
        }
        phase ^= 1u << s;
      };
      auto test_stage = [&](int s) -> bool { return mbar_test_wait(&full[ty][s], (phase >> s) & 1u); };
      auto release_stage = [&](int s) {
        __syncwarp();
        if (tx == 0) mbar_arrive(&empty[ty][s]);
      };
      auto prefetch_l1 = [](const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); };

      int s_tail = sf, s_old = sf, s_new = next_stage(sf);  // stages of planes ll - 1 (tail), ll, ll + 1 at step ll
      wait_stage(s_old, false);  // the first plane only serves as the bottom plane of the first layer
      bool hint_new = false;
      double carry[2][3], nA[2][3], nB[2][3];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) carry[r][c] = nA[r][c] = nB[r][c] = 0.0;
      double* yp[2] = {y + ((long long)(first - 1) * g.S + ncol[0]) * 3, y + ((long long)(first - 1) * g.S + ncol[1]) * 3};
      const long long ystep = (long long)g.S * 3;
      // E_e of the layer and the prescribed-dof flags of the plane being stored: 10 bytes per thread and step, pulled
      // into L1 one step ahead (prefetch.global.L1 holds no register) and loaded where they are used
      const double* Ep[2] = {E + (long long)first * g.SE + ecol[0], E + (long long)first * g.SE + ecol[1]};
      const unsigned char* fp[2] = {fixed + (long long)(first - 1) * g.S + ncol[0], fixed + (long long)(first - 1) * g.S + ncol[1]};
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (el_ok[k]) prefetch_l1(Ep[k]);
      int par = 0, hi_seen = 0, lo_seen = it;

      // Tail of layer L (deferred to the start of the next step): y reduction with the partial sums of the thread row
      // below, inverse z stage, store of plane L, dot products, release of plane L's ring stage.  It only depends on the
      // thread row BELOW, so the rows form a one-directional chain.  Branch-free: lanes that do not own their node
      // read zeros as x and are treated as fully prescribed, so they store nothing and add 0 to the dot products.
      // Called with L = first - 1 before the first step: a no-op (carry is overwritten before it is used).
      auto tail = [&](int L, int parL, const unsigned char (&flraw)[2]) {
        {
          const int lo = ty > 0 ? ty - 1 : 0;
          if (!__all_sync(FULL, lo_seen >= it)) {  // lo_seen: read right after this row's own publication, a step ago
            int spins = 0;
            bool ready;
            do {
              ready = *(volatile int*)&sflag[lo] >= it || ++spins > (1 << 24);
            } while (!__all_sync(FULL, ready));
          }
          __threadfence_block();
          // the row above must have consumed the y buffer this step will overwrite (written three steps ago):
          // true once it has published step it - 2; read here, early, tested right before the stores
          hi_seen = *(volatile int*)&sflag[ty + 1 < TYT ? ty + 1 : ty];
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
          const unsigned pl = plane_lead(L);
          const unsigned char* sb = ring + s_tail * kRingStage + lane_off;
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            st_ok[r] = store && own[r];
            const double* q = st_ok[r] ? reinterpret_cast<const double*>(sb + r * kRingPitch + (int)(pl ^ (r == 1 ? leadNB : leadN0))) : zero3;
#pragma unroll
            for (int c = 0; c < 3; ++c) xo[r][c] = q[c];
            fl[r] = st_ok[r] ? flraw[r] : (unsigned char)7;
          }
        }
        if (L >= first) release_stage(s_tail);  // plane L has been read for the last time: hand its stage back
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
            if (fl[r] & (1 << c)) v = fixed_diag * xo[r][c];  // prescribed row: meandiag * x
            if (st_ok[r]) yp[r][c] = v;
            if (DOT >= 1) dots[0] = fma(xo[r][c], v, dots[0]);
            if (DOT == 2) dots[1] = fma(v, v, dots[1]);
          }
          yp[r] += ystep;
        }
      };

      for (int ll = first; ll < z1; ++ll, par = par == 2 ? 0 : par + 1) {  // element layer ll: planes ll (bottom), ll + 1 (top)
        // E_e of this layer and the flags of the plane the tail stores: L1 hits (prefetched one step ago), issued
        // before the waits so that their latency is hidden; then the prefetch for the next step
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
        wait_stage(s_new, hint_new);
        hint_new = test_stage(next_stage(s_new));  // next step's plane: usually landed already
        tail(ll - 1, par == 0 ? 2 : par - 1, flraw);
        // ---- forward x and z stages of rows A, B, C from the raw planes ll and ll + 1
        RowX xr[3];
        {
          const unsigned plb = plane_lead(ll), plt = plane_lead(ll + 1);
          const unsigned char* bb = ring + s_old * kRingStage + lane_off;
          const unsigned char* bt = ring + s_new * kRingStage + lane_off;
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const unsigned lk = k == 1 ? leadNB : leadN0;
            const double* qb = reinterpret_cast<const double*>(bb + k * kRingPitch + (int)(plb ^ lk));
            const double* qt = reinterpret_cast<const double*>(bt + k * kRingPitch + (int)(plt ^ lk));
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
        // ---- the two elements; E-scaled accumulation onto the x edges of rows A, B, C
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
        // ---- inverse x stage per x edge, reduction over x by shuffle; row C goes to the thread row above.
        // The y buffer written now (one of three) was last read by the row above in its tail of step ll - 3, which
        // precedes its publication of step ll - 2: normally known from the flag value read at the top of the step.
        if (!__all_sync(FULL, hi_seen >= it - 1)) {
          const int hi = ty + 1 < TYT ? ty + 1 : ty;
          int spins = 0;
          bool ready;
          do {
            ready = *(volatile int*)&sflag[hi] >= it - 1 || ++spins > (1 << 24);
          } while (!__all_sync(FULL, ready));
        }
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
        lo_seen = *(volatile int*)&sflag[ty > 0 ? ty - 1 : 0];  // consumed by the next tail
        s_tail = s_old;
        s_old = s_new;
        s_new = next_stage(s_new);
      }
      {
        unsigned char flraw[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) flraw[k] = node_ok[k] ? fp[k][0] : (unsigned char)0;
        tail(z1 - 1, par == 0 ? 2 : par - 1, flraw);
      }
      release_stage(s_old);  // the top plane of the last layer was only read as a "new" plane
      sf = s_new;
    }  // segments
  }
  // fin & 0x100: also tell the slab neighbours that this rank's product is final (first K.p of the one-kernel CG loop)
  if (DOT) block_partials_finish<NS>(dots, partials, st, fin & 0xff, sm, PEER && (fin & 0x100) != 0);
}

}  // namespace topopt
