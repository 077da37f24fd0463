// One CG iteration in ONE kernel: the vector updates of iteration k-1 are applied while the node planes of the
// direction vector are staged for the K.u of iteration k (hex8 elasticity, modal form, single GPU).
//
//   r_k = r_{k-1} - alpha_{k-1} Ap_{k-1}      p_k = r_k + beta_{k-1} p_{k-1}      x_k = x_{k-1} + alpha_{k-1} p_{k-1}
//   Ap_k = K p_k                              sums: p_k.Ap_k, Ap_k.Ap_k, r_k.r_k  -> alpha_k, beta_k (cg_finalize FIN_FUSED)
//
// Same recurrence as k_update_xrp followed by k_apply_hex8_ring (single-pass CG: beta is predicted from
// alpha^2 Ap.Ap - r.r, the true r.r of every residual is accumulated so the prediction error does not build up),
// but the 56 bytes per dof of the vector pass and the 16 bytes per dof of K.u are no longer two passes that run one
// after the other: the K.u arithmetic (fp64-pipe bound) overlaps the vector traffic (HBM bound), and p is never
// re-read.  Per iteration the kernel moves 64 bytes per dof + 8 per element.
//
// Data flow (see kxu_hex8_ring.cuh for the K.u part, which is unchanged):
//   * the producer warp bulk-copies the rows of p_{k-1} into each compute warp's private ring AND the matching rows
//     of r_{k-1} and Ap_{k-1} into a two-deep staging area (cp.async.bulk, one mbarrier per warp and stage);
//   * when a plane has landed, the compute warp combines its two node rows in place (p_k replaces p_{k-1} in the ring),
//     writes r_k, p_k and x_k of the nodes the tile owns to global memory and accumulates r_k.r_k; x_{k-1} travels
//     through cp.async (LDGSTS) into a per-lane slot, issued one step ahead right after the slot was consumed (loaded
//     where it is used it exposed a DRAM latency per step, 22 % of all stall samples; a bulk-copy slot filled by the
//     producer made the plane's mbarrier wait for that same latency instead);
//   * row C of a warp's ring (= row A of the thread row above) is written by that thread row after its combine instead
//     of being copied and combined twice; a per-row counter in shared memory tells the lower row when it is there;
//   * p, r and Ap are ping-pong buffers selected by the parity of the iteration counter, because the tiles' halo rows,
//     columns and planes are read (and combined redundantly) by neighbouring tiles while the owner already writes the
//     new values.  x is only touched by its owner and is updated in place.
// Reference: IterativeSolvers cg! as driven by src/FEA/iterative_solver.jl (solvers_api.jl:177-220); the recurrence
// is SURVEY App. A.3's with the scalar step moved to the device.
#pragma once
#include "kxu_hex8_ring.cuh"

namespace topopt {

constexpr int kFusedProducers = 2;  // producer warps: every bulk copy costs ~50 cycles of its issuing warp (one lane at a time
                                    // through the uniform datapath), 63 copies per plane -- a single producer is the bottleneck
constexpr int kFusedStageRows = 8;  // staging rows per compute warp: 2 stages x {r, Ap} x rows {A, B}

__host__ __device__ constexpr size_t hex8_fused_smem(int tyt, int nst) {
  return (size_t)tyt * nst * kRingStage + (size_t)tyt * kFusedStageRows * kRingPitch + 4 * kRingPitch +
         sizeof(double) * 2 * 3 * 32 * tyt + sizeof(double) * 2 * 6 * (tyt + 1) * 32;
}

#ifdef TOPOPT_TIMELINE  // diagnostic build: %globaltimer stamps of two consecutive iterations (tools/r02_timeline.py)
__device__ unsigned long long g_tl[2][512][4];
__device__ __forceinline__ unsigned long long tl_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TL_STAMP(ev)                                                              \
  do {                                                                            \
    if (tl_slot >= 0 && tl_slot < 2 && blockIdx.x < 512) g_tl[tl_slot][blockIdx.x][ev] = tl_now(); \
  } while (0)
#else
#define TL_STAMP(ev) \
  do {               \
  } while (0)
#endif

struct CGFusedVecs {
  double* p[2];
  double* r[2];
  double* ap[2];
  double* x;
  // multi-GPU (PEER): the slab neighbours' GHOST planes of the same six buffers, [p, r, Ap][parity], mapped over cudaIpc;
  // wlo = the lower neighbour's top ghost plane (receives this rank's plane 1), whi = the upper neighbour's bottom ghost
  // plane (receives this rank's plane nown).  NULL at the ends of the slab stack.
  double* wlo[3][2];
  double* whi[3][2];
};

// PEER: a rank PUSHES the boundary planes of p_k, r_k, Ap_k it forms into its neighbours' ghost planes (posted NVLink
// stores next to the local ones) and signals from its last CTA; the next kernel stages its ghost planes from LOCAL memory
// once the neighbours' signal of the previous kernel is in.  (Pulling them with bulk copies from the neighbours' memory
// cost ~0.4 us per 800-byte request: 25 us on the critical CTAs of a 2-rank run, profiles/r02_timeline_n2_remote_reads.log.)
// The ping-pong parity keeps a neighbour that is one kernel ahead from overwriting ghost planes that are still being read:
// it cannot start iteration k + 1 before this rank has posted its sums of iteration k, which it does after its last read.
template <int TYT, int NST, bool CUBE, bool PEER>
__global__ void __launch_bounds__(32 * (TYT + kFusedProducers), 1)
    k_cg_fused_hex8(Geo g, CGFusedVecs vec, const double* __restrict__ E, const unsigned char* __restrict__ fixed, double fixed_diag,
                    int tilesX, int tilesY, double* partials, CGState* st, int fin) {
  static_assert(NST % 2 == 0, "the staging area is two deep: its slot is the ring stage modulo 2");
  constexpr int WRING = NST * kRingStage;                 // bytes of one warp's private p ring
  constexpr int WSTG = kFusedStageRows * kRingPitch;      // bytes of one warp's staging area: [stage & 1][r, Ap][row A, B]
  constexpr int STG0 = TYT * WRING;                       // first staging area
  constexpr int TOP0 = STG0 + TYT * WSTG;                 // row C of the top thread row: [stage & 1][r, Ap]
  constexpr int XS0 = TOP0 + 4 * kRingPitch;              // x_{k-1} of the next plane: [thread row][row A, B][j][lane]
  constexpr int YB0 = XS0 + TYT * 2 * 3 * 32 * 8;
  constexpr int OWNR = 2 * TYT - 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double(*yb)[6][TYT + 1][32] = reinterpret_cast<double(*)[6][TYT + 1][32]>(smem_raw + YB0);
  // full: a plane's rows have landed; empty: ring stage released (tail); sempty: staging slot released (combine).  The
  // staging area is only two deep, so its slots are handed back separately: at a segment boundary two ring stages are
  // released back to back while the first planes of the next segment still wait to be combined.
  __shared__ uint64_t full[TYT][NST], empty[TYT][NST], sempty[TYT][2];
  __shared__ int sflag[TYT], cflag[TYT];
  __shared__ double zero3[4];
  __shared__ double sm[32];
  const int tid = threadIdx.x;
  const int tx = tid & 31, wid = tid >> 5;
  const bool producer = wid >= TYT;
  const int ty = producer ? 0 : TYT - 1 - wid;
  const unsigned FULL = 0xffffffffu;
  if (tid < TYT * NST) {
    const int w = tid / NST;
    mbar_init(&full[w][tid % NST], w == TYT - 1 ? 9 : 6);  // bulk copies per plane: {p, r, Ap} x rows {A, B} (+ C on top)
    mbar_init(&empty[w][tid % NST], 1);
    if (tid % NST < 2) mbar_init(&sempty[w][tid % NST], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < TYT) sflag[tid] = cflag[tid] = 0;
  if (tid < 4) zero3[tid] = 0.0;
  for (int i = tid; i < YB0 / 16; i += (int)blockDim.x) reinterpret_cast<uint4*>(smem_raw)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < 2 * 6 * (TYT + 1) * 32; i += (int)blockDim.x) (&yb[0][0][0][0])[i] = 0.0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  // Programmatic dependent launch: everything above (220 KB of shared memory zeroed, barriers initialised) ran while the
  // previous iteration's kernel was still finishing its reduction / all-reduce; nothing it wrote has been read yet.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (st->done) return;
  const int parity = st->iters & 1;
#ifdef TOPOPT_TIMELINE
  const int tl_slot = st->iters - 300;
#endif
  const double alpha = st->alpha, beta = st->beta;
  const double* __restrict__ pin = vec.p[parity];
  const double* __restrict__ rin = vec.r[parity];
  const double* __restrict__ apin = vec.ap[parity];
  double* __restrict__ pout = vec.p[parity ^ 1];
  double* __restrict__ rout = vec.r[parity ^ 1];
  double* __restrict__ y = vec.ap[parity ^ 1];
  double* __restrict__ xsol = vec.x;

  double dots[3] = {0.0, 0.0, 0.0};  // p.Ap, Ap.Ap, r.r
  const long long units = (long long)tilesX * tilesY * g.nown;
  long long u0 = units * blockIdx.x / gridDim.x;
  const long long u1 = units * (blockIdx.x + 1) / gridDim.x;
  if (tid == 0) TL_STAMP(0);

  if (producer) {
    // =========================== producer warp ===========================
    // job j < 6 TYT: thread row j / 6, vector (j % 6) / 2 in {p, r, Ap}, row (j % 2) in {A, B}; the last three jobs are
    // row C of the top thread row (its upper neighbour belongs to another tile).
    // the jobs are dealt round-robin to the producer warps
    constexpr int NJ = 6 * TYT + 3, JPL = (NJ + 32 * kFusedProducers - 1) / (32 * kFusedProducers);
    const int pw = wid - TYT;
    struct Job {
      long long su0;
      long long idx;
      const double* src;
      int P, last;
      int dst, dst_stride, w, k, clip;
      unsigned fcount;
      uint32_t bytes;
      bool valid, done, ok, stg;
    } job[JPL];
#pragma unroll
    for (int q = 0; q < JPL; ++q) {
      const int j = pw + kFusedProducers * (tx + 32 * q);
      Job& J = job[q];
      J.valid = j < NJ;
      J.done = !J.valid;
      int v;
      if (j < 6 * TYT) {
        J.w = j / 6;
        v = (j % 6) / 2;
        J.k = j % 2;
      } else {
        J.w = TYT - 1;
        v = j - 6 * TYT;
        J.k = 2;
      }
      J.src = v == 0 ? pin : (v == 1 ? rin : apin);
      J.stg = v != 0;
      // destination of stage 0 and the distance between stages (the staging area reuses slot `stage & 1`)
      if (v == 0) {
        J.dst = J.w * WRING + J.k * kRingPitch;
        J.dst_stride = kRingStage;
      } else if (J.k < 2) {
        J.dst = STG0 + J.w * WSTG + (v - 1) * 2 * kRingPitch + J.k * kRingPitch;
        J.dst_stride = 4 * kRingPitch;
      } else {
        J.dst = TOP0 + (v - 1) * kRingPitch;
        J.dst_stride = 2 * kRingPitch;
      }
      J.su0 = u0;
      J.P = 1;
      J.last = 0;
      J.fcount = 0;
      J.idx = 0;
      J.bytes = 0;
      J.ok = false;
      J.clip = 0;
    }
    if (PEER && st->peer != nullptr) {
      // The ghost planes of this kernel's inputs were pushed into local memory by the neighbours' previous kernel: wait for
      // their signal ONCE, before the first copy.  (It was raised before they posted the sums this rank's previous kernel
      // waited for, so this does not spin; PEER code inside the copy loop cost 30 us per launch.)
      PeerComm* pc = st->peer;
      const volatile unsigned long long* f = pc->block[pc->rank]->halo_flag;
      const bool need_lo = vec.wlo[0][0] != nullptr, need_hi = vec.whi[0][0] != nullptr;
      long long spins = 0;
      while ((need_lo && f[0] < pc->halo_seq) || (need_hi && f[1] < pc->halo_seq)) {
        if (++spins > kSpinLimit / 8) {
          pc->timeout = 1;  // a dead peer must not hang the GPU: proceed, the solve reports the error
          break;
        }
      }
      // No fence: the neighbour ordered its plane stores before its flag store (system-scope fence on its side), both land
      // in THIS GPU's L2, and the bulk copies below are issued after the loop has seen the flag and read through L2.  (A
      // system-scope fence here cost every CTA ~4 us before its first copy.)
    }
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
          J.P = zoff;
          J.last = zoff + zlen + 1;
          J.ok = frow >= 0 && frow < g.NY;
          J.idx = 24ll * ((long long)frow * g.NX + c_lo);
          J.clip = c_lo != c0 ? 32 : 0;
          J.bytes = 24u * (uint32_t)cnt;
        }
        alldone = false;
        const int sidx = (int)(J.fcount % NST);
        const bool ready = J.stg ? (J.fcount < 2 || mbar_test_wait(&sempty[J.w][J.fcount & 1u], ((J.fcount >> 1) + 1u) & 1u))
                           : (J.fcount < NST || mbar_test_wait(&empty[J.w][sidx], ((J.fcount / NST) + 1u) & 1u));
        const char* plane = reinterpret_cast<const char*>(J.src + (long long)J.P * g.S * 3);
        if (ready) {
          uint64_t* bar = &full[J.w][sidx];
          if (J.ok) {
            const char* a = plane + J.idx;
            const uint32_t lead = (uint32_t)(reinterpret_cast<uintptr_t>(a) & 15);
            const uint32_t nb = (lead + J.bytes + 15u) & ~15u;
            const int slot = J.dst_stride == kRingStage ? sidx : (sidx & 1);
            mbar_arrive_expect_tx(bar, nb);
#ifdef TOPOPT_EXPERIMENT_SPLIT_COPIES  // diagnostic: twice the bulk-copy requests for the same bytes
            {
              const uint32_t h1 = (nb / 2) & ~15u;
              bulk_g2s(smem_raw + J.dst + slot * J.dst_stride + J.clip, a - lead, h1, bar);
              bulk_g2s(smem_raw + J.dst + slot * J.dst_stride + J.clip + h1, a - lead + h1, nb - h1, bar);
            }
#else
            bulk_g2s(smem_raw + J.dst + slot * J.dst_stride + J.clip, a - lead, nb, bar);
#endif
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
    unsigned char* ring = smem_raw + ty * WRING;
    const unsigned char* stg = smem_raw + STG0 + ty * WSTG;
    int it = 0;
    int cc = 0;          // planes combined so far (monotonic across segments; equal on all thread rows)
    int sf = 0;
    unsigned phase = 0;
    auto next_stage = [](int s) { return s + 1 == NST ? 0 : s + 1; };
    while (u0 < u1) {
      const int tile = (int)(u0 / g.nown);
      const int zoff = (int)(u0 % g.nown);
      const int zlen = (int)min((long long)(g.nown - zoff), u1 - u0);
      u0 += zlen;
      const int bx = tile % tilesX, by = tile / tilesX;
      const int c0 = bx * 31 - 1, r0 = by * OWNR - 1 + 2 * ty;
      const int z0 = 1 + zoff, z1 = z0 + zlen;
      const int first = z0 - 1;
      const int c_lo = max(c0, 0);
      const int shift = c_lo - c0;
      asm volatile("bar.sync 1, %0;" ::"n"(32 * TYT) : "memory");

      const int col = c0 + tx;
      const int lane_off = (shift ? 32 - 24 : 0) + 24 * tx;
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
      const unsigned leadN0 = (unsigned)(((long long)r0 * g.NX + c_lo) & 1) << 3;
      const unsigned leadNB = leadN0 ^ ((unsigned)(g.NX & 1) << 3);
      auto plane_lead = [&](int P) -> unsigned {
        return (unsigned)(reinterpret_cast<uintptr_t>(pin + (long long)P * g.S * 3) & 8);
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
      const int voff = (shift ? 8 : 0) + 8 * tx;  // byte offset of value j = 0 inside a ring row (before the lead)
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
            landed = mbar_try_wait(&full[ty][s], par);
          } while (!__all_sync(FULL, landed));
        }
        phase ^= 1u << s;
      };
      auto test_stage = [&](int s) -> bool { return mbar_test_wait(&full[ty][s], (phase >> s) & 1u); };
      auto release_stage = [&](int s) {
        __syncwarp();
        if (tx == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // this warp wrote the stage through the generic proxy
          mbar_arrive(&empty[ty][s]);
        }
      };
      auto prefetch_l1 = [](const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); };

      // x_{k-1} of the nodes this thread writes in plane P (flat values j = 0..2; j = 3 is the halo column): asynchronous
      // copies into the lane's own slots, waited for at the top of the next step
      double(*xs)[3][32] = reinterpret_cast<double(*)[3][32]>(smem_raw + XS0 + ty * (2 * 3 * 32 * 8));
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
      // r, p and x written to global memory.
      auto combine = [&](int P, int s) {
        const bool wr = P >= z0 && P < z1;
        const unsigned pl = plane_lead(P);
        const unsigned char* sr = stg + (s & 1) * (4 * kRingPitch);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (!crow_ok[k]) continue;  // warp-uniform
          const int lo = (int)(pl ^ (k == 1 ? leadNB : leadN0)) + voff;
          double* qp = reinterpret_cast<double*>(ring + s * kRingStage + k * kRingPitch + lo);
          const double* qr = reinterpret_cast<const double*>(sr + k * kRingPitch + lo);
          const double* qa = reinterpret_cast<const double*>(sr + (2 + k) * kRingPitch + lo);
          // row A is row C of the thread row below
          double* qc = reinterpret_cast<double*>(smem_raw + (ty > 0 ? ty - 1 : 0) * WRING + s * kRingStage + 2 * kRingPitch + lo);
          const bool wo = wr && crow_own[k];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (!((vmask >> j) & 1u)) continue;
            const double po = qp[32 * j];
            const double rn = fma(-alpha, qa[32 * j], qr[32 * j]);
            const double pn = fma(beta, po, rn);
            qp[32 * j] = pn;
            if (k == 0 && ty > 0) qc[32 * j] = pn;
            if (wo && ((omask >> j) & 1u)) {
              const long long gi = (long long)P * ystep + grow[k] + 32 * j;
              rout[gi] = rn;
              pout[gi] = pn;
              xsol[gi] = fma(alpha, po, xs[k][j < 3 ? j : 0][tx]);
              dots[2] = fma(rn, rn, dots[2]);

            }
          }
        }
        if (ty == TYT - 1 && crow_ok[2]) {  // the top thread row combines its own row C (never owned by this tile)
          const int lo = (int)(pl ^ leadN0) + voff;
          double* qp = reinterpret_cast<double*>(ring + s * kRingStage + 2 * kRingPitch + lo);
          const double* qr = reinterpret_cast<const double*>(smem_raw + TOP0 + (s & 1) * (2 * kRingPitch) + lo);
          const double* qa = reinterpret_cast<const double*>(smem_raw + TOP0 + (s & 1) * (2 * kRingPitch) + kRingPitch + lo);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (!((vmask >> j) & 1u)) continue;
            const double rn = fma(-alpha, qa[32 * j], qr[32 * j]);
            qp[32 * j] = fma(beta, qp[32 * j], rn);
          }
        }
        ++cc;
        __syncwarp();
        if (tx == 0) {
          mbar_arrive(&sempty[ty][s & 1]);  // r / Ap of this plane are consumed
          __threadfence_block();
          *(volatile int*)&cflag[ty] = cc;
        }
      };
      // row C of every plane combined so far has been written by the thread row above
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
      wait_stage(s_old, false);
#ifdef TOPOPT_TIMELINE
      if (ty == 0 && tx == 0 && it == 0) TL_STAMP(1);
#endif
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
        wait_stage(s_new, hint_new);
        hint_new = test_stage(next_stage(s_new));
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        combine(ll + 1, s_new);
        issue_x(ll + 2);
        tail(ll - 1, par ^ 1, flraw);
        wait_rowC();
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
  if (PEER) {
    // Push pass: the boundary planes (1 and nown) of p_k, r_k, Ap_k that this CTA's segments own go to the neighbours'
    // ghost planes -- re-read from this rank's memory (L2-hot, written by this CTA) and stored over NVLink by all threads.
    // Kept out of the plane loop: pushes inside the combine / tail cost every step registers and predicates.
    __syncthreads();
    long long su = units * blockIdx.x / gridDim.x;
    const long long ystep = (long long)g.S * 3;
    while (su < u1) {
      const int tile = (int)(su / g.nown);
      const int zoff = (int)(su % g.nown);
      const int zlen = (int)min((long long)(g.nown - zoff), u1 - su);
      su += zlen;
      const int bx = tile % tilesX, by = tile / tilesX;
      const int rlo = by * OWNR, rhi = min(rlo + OWNR, g.NY), clo = bx * 31, chi = min(clo + 31, g.NX);
      const int ncv = (chi - clo) * 3, nrows = rhi - rlo;
#pragma unroll
      for (int side = 0; side < 2; ++side) {
        const int P = side == 0 ? 1 : g.nown;
        const bool mine = side == 0 ? zoff == 0 : zoff + zlen == g.nown;
        double* dp = side == 0 ? (parity ? vec.wlo[0][0] : vec.wlo[0][1]) : (parity ? vec.whi[0][0] : vec.whi[0][1]);
        double* dr = side == 0 ? (parity ? vec.wlo[1][0] : vec.wlo[1][1]) : (parity ? vec.whi[1][0] : vec.whi[1][1]);
        double* da = side == 0 ? (parity ? vec.wlo[2][0] : vec.wlo[2][1]) : (parity ? vec.whi[2][0] : vec.whi[2][1]);
        if (!mine || dp == nullptr || ncv <= 0 || nrows <= 0) continue;
        for (int idx = tid; idx < nrows * ncv; idx += (int)blockDim.x) {
          const long long off = ((long long)(rlo + idx / ncv) * g.NX + clo) * 3 + idx % ncv;
          dp[off] = __ldcg(pout + P * ystep + off);
          dr[off] = __ldcg(rout + P * ystep + off);
          da[off] = __ldcg(y + P * ystep + off);
        }
      }
    }
  }
  if (tid == 0) TL_STAMP(2);
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // the next iteration's CTAs may start their prologue
  block_partials_finish<3>(dots, partials, st, fin, sm, PEER);
  if (tid == 0) TL_STAMP(3);
}

}  // namespace topopt
