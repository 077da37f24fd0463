// Device kernels of libtopopt_cuda (fp64, sm_100a).  All kernels are atomic-free on the data
// path and deterministic: reductions go through fixed-order per-block partials that the last
// block to finish sums in a fixed order.
//
// Data layout (per rank): nodes lexicographic (x fastest), dofs interleaved per node
// (ncomp*node + c); the grid is cut into slabs along the last axis.  A rank stores its owned
// node planes 1..nown plus one ghost plane on each side (local planes 0 and nown+1) and its
// owned element layers 1..nlay plus one ghost layer below (local layer 0).  Local layer l
// touches local node planes l and l+1.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace topopt {

struct Geo {
  int NX, NY;      // nodes per in-plane axis (NY = 1 in 2-D)
  int nx, ny;      // elements per in-plane axis (ny = 1 in 2-D)
  int S, SE;       // nodes per plane, elements per layer
  int NPg, NLg;    // global number of node planes / element layers along the slab axis
  int p0;          // global index of local plane 0 ( = global index of local layer 0 ); may be -1
  int nown;        // owned node planes (local 1..nown)
  int nlay;        // owned element layers (local 1..nlay)
};

constexpr int kMaxRanks = 8;

// Peer-memory communication block (one per rank, in a cudaIpc-shared allocation that every other
// rank of the node maps).  Used by the in-kernel one-shot allreduce of the CG scalars and by the
// halo-ready handshake of the K.u kernel: plain stores / loads over NVLink, no NCCL on the
// per-iteration path.
struct PeerBlock {
  // [parity][source rank][scalar] = (value, sequence number): every scalar travels with its own
  // sequence tag in ONE aligned 16-byte store, so the receiver needs neither a fence nor a flag
  double2 slots[4][kMaxRanks][8];           // four-deep: a deferred value stays valid until its collector has run
  unsigned long long halo_flag[2];          // [0]: written by the rank below, [1]: by the rank above
};

struct PeerComm {
  PeerBlock* block[kMaxRanks];  // block[r] = rank r's PeerBlock as mapped into this process
  unsigned long long seq;       // allreduce sequence number (device-maintained, never reset)
  unsigned long long halo_seq;  // halo handshake sequence number
  int rank, world;
  int timeout;                  // set when a spin-wait gave up (peer died): reported as an NCCL-class error
};

struct CGState {
  double sums[8];   // this rank's reduction results
  double gsums[8];  // after the allreduce (aliases sums when world == 1)
  double res, prev_res, rho, rho_prev, alpha, beta, tol, abstol, reltol, energy, pAp;
  int iters, done, converged, maxiter, criteria, precond, nonfinite, world;
  int variant;      // 0 = IterativeSolvers recurrence, 1 = single-pass recurrence (FIN_PAP2 / FIN_RR2)
  int rr_pending;   // multi-GPU single-pass: this rank's r.r is posted to the peers but not collected yet
  unsigned long long rr_seq;
  unsigned int counter, counter2;
  PeerComm* peer;   // non-null: reductions are all-reduced inside the kernel over peer memory
};

constexpr long long kSpinLimit = 400000000LL;  // ~1 s of polling before a wait is abandoned

constexpr int kMaxKe = 24;
__constant__ double cKe[kMaxKe * kMaxKe];  // shared element matrix, column-major

constexpr int kBlock = 256;
constexpr int kReduceBlocks = 148 * 4;  // persistent grid for vector kernels (multiple of the SM count)

// ---- reductions ---------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// sum over the block; result valid in thread 0.  sm must hold 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* sm) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sm[w] = v;
  __syncthreads();
  if (w == 0) {
    v = (lane < (blockDim.x >> 5)) ? sm[lane] : 0.0;
    v = warp_sum(v);
  }
  return v;
}

enum { FIN_NONE = 0, FIN_INIT = 1, FIN_PAP = 2, FIN_RR = 3, FIN_PLAIN = 4, FIN_PAP2 = 5, FIN_RR2 = 6, FIN_FUSED = 7 };

// scalar recurrences of IterativeSolvers.cg! (CGIterable / PCGIterable iterate) and the
// convergence tests of src/FEA/convergence_criteria.jl:26-45, evaluated on the device so the
// host never has to synchronise inside the loop.
__device__ inline void cg_finalize(CGState* st, int which) {
  const double* s = st->gsums;
  if (which == FIN_PAP) {
    st->pAp = s[0];
    st->alpha = st->precond ? st->rho / s[0] : (st->res * st->res) / s[0];
    return;
  }
  if (which == FIN_PAP2) {
    // Single-pass variant: s = {p.Ap, Ap.Ap}.  With r.Ap = p.Ap (A-orthogonality of the directions)
    // |r - alpha Ap|^2 = alpha^2 Ap.Ap - r.r, so beta is known BEFORE the residual is updated and
    // x += alpha p, r -= alpha Ap, p = r + beta p become one pass over the vectors.  st->rho is the TRUE
    // r.r accumulated by that pass (FIN_RR2), so the prediction error does not accumulate.
    st->pAp = s[0];
    const double rho = st->rho;
    const double alpha = rho / s[0];
    double rho_next = fma(alpha * alpha, s[1], -rho);
    if (!(rho_next > 0.0)) rho_next = 0.0;  // cancellation at a (near-)exact solve: restart with steepest descent
    st->alpha = alpha;
    st->beta = rho_next / rho;
    return;
  }
  if (which == FIN_FUSED) {
    // One-kernel iteration (kxu_hex8_cgfused.cuh): s = {p.Ap, Ap.Ap, r.r} of the residual, direction and product the
    // kernel has just formed.  First the end of the iteration that produced r (FIN_RR2), then, unless it converged,
    // the scalars of the next one (FIN_PAP2).
    st->prev_res = st->res;
    st->res = sqrt(s[2]);
    st->iters += 1;
    st->rho = s[2];
    const bool conv = st->res <= st->tol;
    if (isnan(st->res)) st->nonfinite = 1;
    st->converged = conv ? 1 : 0;
    st->done = (conv || st->iters >= st->maxiter || st->nonfinite) ? 1 : 0;
    st->pAp = s[0];
    const double rho = s[2];
    const double alpha = rho / s[0];
    double rho_next = fma(alpha * alpha, s[1], -rho);
    if (!(rho_next > 0.0)) rho_next = 0.0;
    st->alpha = alpha;
    st->beta = rho_next / rho;
    return;
  }
  if (which == FIN_INIT) {
    st->res = sqrt(s[0]);
    st->tol = fmax(st->reltol * st->res, st->abstol);
    st->prev_res = 1.0;
    st->rho_prev = 1.0;
    st->energy = 0.0;
    st->iters = 0;
  } else {  // FIN_RR, FIN_RR2
    st->prev_res = st->res;
    st->res = sqrt(s[0]);
    st->iters += 1;
  }
  if (st->variant == 1) {
    st->rho = s[0];
  } else if (st->precond) {
    if (which == FIN_RR) st->rho_prev = st->rho;
    st->rho = s[1];
    st->beta = st->rho / st->rho_prev;
  } else {
    st->beta = (st->res * st->res) / (st->prev_res * st->prev_res);
  }
  bool conv;
  if (st->criteria == 0) {
    conv = st->res <= st->tol;
  } else {
    const double xtr = s[2];
    const double xAx = s[3] - xtr;
    const double change = xAx - st->energy;
    if (isnan(change) || isnan(xAx) || xAx < 0.0) st->nonfinite = 1;
    conv = (fabs(change) / xAx <= st->tol) && (fabs(xtr) / xAx <= st->tol);
    st->energy = xAx;
  }
  if (isnan(st->res)) st->nonfinite = 1;
  st->converged = conv ? 1 : 0;
  st->done = (conv || st->iters >= st->maxiter || st->nonfinite) ? 1 : 0;
}

// "this rank's direction vector is final for this iteration": told to both slab neighbours, which read
// its boundary planes over NVLink inside their K.u kernel
__device__ __forceinline__ void signal_halo_flags(CGState* st) {
  PeerComm* pc = st->peer;
  const unsigned long long seq = pc->halo_seq + 1;
  __threadfence_system();
  if (pc->rank + 1 < pc->world) *(volatile unsigned long long*)&pc->block[pc->rank + 1]->halo_flag[0] = seq;
  if (pc->rank > 0) *(volatile unsigned long long*)&pc->block[pc->rank - 1]->halo_flag[1] = seq;
  pc->halo_seq = seq;
}

// Each block deposits NS partial sums; the last block to arrive adds all partials in a fixed
// order, publishes them and (single-GPU) runs the scalar step.
template <int NS>
__device__ __forceinline__ void block_partials_finish(const double (&v)[NS], double* partials,
                                                      CGState* st, int which, double* sm, bool signal_halo = false) {
  __shared__ bool is_last;
  double r[NS];
#pragma unroll
  for (int k = 0; k < NS; ++k) r[k] = block_sum(v[k], sm);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NS; ++k) partials[blockIdx.x * NS + k] = r[k];
    // signal_halo: this kernel may have written into the neighbours' memory (pushed boundary planes): system scope
    if (signal_halo) __threadfence_system(); else __threadfence();
    const unsigned int t = atomicInc(&st->counter, gridDim.x - 1);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // every block's stores are visible now: the vector this kernel wrote may be read by the neighbours
  if (signal_halo && threadIdx.x == 0 && st->peer != nullptr) signal_halo_flags(st);
  double a[NS];
#pragma unroll
  for (int k = 0; k < NS; ++k) a[k] = 0.0;
  for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
#pragma unroll
    for (int k = 0; k < NS; ++k) a[k] += __ldcg(&partials[b * NS + k]);
  }
#pragma unroll
  for (int k = 0; k < NS; ++k) a[k] = block_sum(a[k], sm);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NS; ++k) st->sums[k] = a[k];
    if (st->world == 1 && which != FIN_NONE && which != FIN_PLAIN) {
#pragma unroll
      for (int k = 0; k < NS; ++k) st->gsums[k] = a[k];
      cg_finalize(st, which);
    }
  }
  PeerComm* pc = st->peer;
  if (st->world > 1 && pc != nullptr && which != FIN_NONE && which != FIN_PLAIN) {
    // one-shot allreduce over peer memory: every rank stores its partial sums into every rank's
    // slot array, then waits for all contributions and adds them in rank order (bitwise identical
    // on all ranks).  Slots are four-deep in the sequence number.
    // Single-pass recurrence: the r.r of the fused vector pass (FIN_RR2) is only POSTED here; it is
    // collected together with p.Ap / Ap.Ap in the next K.u kernel (FIN_PAP2) or by k_cg_flush, so an
    // iteration has ONE rendezvous instead of two and the r.r transfer overlaps the next K.u.
    __shared__ double sh_sums[NS];
    __shared__ double sh_recv[kMaxRanks][NS];
    __shared__ double sh_rr[kMaxRanks];
    const unsigned long long seq = pc->seq + 1;
    const int par = (int)(seq & 3);
    const bool defer = which == FIN_RR2;
    const bool collect_rr = which == FIN_PAP2 && st->rr_pending != 0;
    const unsigned long long rr_seq = st->rr_seq;
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 0; k < NS; ++k) sh_sums[k] = a[k];
    }
    __syncthreads();
    const double tag = __longlong_as_double((long long)seq);
    const int nposts = pc->world * NS;
    if ((int)threadIdx.x < nposts) {
      // thread (r, k): store scalar k into rank r's slot, then poll rank r's scalar k in my slots
      const int r = threadIdx.x / NS, k = threadIdx.x % NS;
      double2* dst = &pc->block[r]->slots[par][pc->rank][defer ? 4 + k : k];
      asm volatile("st.volatile.global.v2.f64 [%0], {%1, %2};" ::"l"(dst), "d"(sh_sums[k]), "d"(tag) : "memory");
      if (!defer) {
        const double2* src = &pc->block[pc->rank]->slots[par][r][k];
        double val, got;
        long long spins = 0;
        while (true) {
          asm volatile("ld.volatile.global.v2.f64 {%0, %1}, [%2];" : "=d"(val), "=d"(got) : "l"(src) : "memory");
          if (__double_as_longlong(got) == (long long)seq) break;
          if (++spins > kSpinLimit) {
            pc->timeout = 1;
            val = 0.0;
            break;
          }
        }
        sh_recv[r][k] = val;
      }
    } else if (collect_rr && (int)threadIdx.x < nposts + pc->world) {
      const int r = threadIdx.x - nposts;
      const double2* src = &pc->block[pc->rank]->slots[(int)(rr_seq & 3)][r][4];
      double val, got;
      long long spins = 0;
      while (true) {
        asm volatile("ld.volatile.global.v2.f64 {%0, %1}, [%2];" : "=d"(val), "=d"(got) : "l"(src) : "memory");
        if (__double_as_longlong(got) == (long long)rr_seq) break;
        if (++spins > kSpinLimit) {
          pc->timeout = 1;
          val = 0.0;
          break;
        }
      }
      sh_rr[r] = val;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      pc->seq = seq;
      if (defer) {
        st->rr_seq = seq;
        st->rr_pending = 1;
      } else {
        if (collect_rr) {
          double g = 0.0;
          for (int r = 0; r < pc->world; ++r) g += sh_rr[r];
          st->gsums[0] = g;
          st->rr_pending = 0;
          cg_finalize(st, FIN_RR2);
        }
#pragma unroll
        for (int k = 0; k < NS; ++k) {
          double g = 0.0;
          for (int r = 0; r < pc->world; ++r) g += sh_recv[r][k];  // rank order: identical on all ranks
          st->gsums[k] = g;
        }
        if (pc->timeout) st->nonfinite = 2;
        if (!(collect_rr && st->done)) cg_finalize(st, which);
      }
    }
  }
}

// collects a posted-but-pending r.r (see block_partials_finish): run after the last iteration of a batch
__global__ void k_cg_flush(CGState* st) {
  __shared__ double sh_rr[kMaxRanks];
  PeerComm* pc = st->peer;
  if (pc == nullptr || !st->rr_pending) return;
  const unsigned long long rr_seq = st->rr_seq;
  if ((int)threadIdx.x < pc->world) {
    const double2* src = &pc->block[pc->rank]->slots[(int)(rr_seq & 3)][threadIdx.x][4];
    double val, got;
    long long spins = 0;
    while (true) {
      asm volatile("ld.volatile.global.v2.f64 {%0, %1}, [%2];" : "=d"(val), "=d"(got) : "l"(src) : "memory");
      if (__double_as_longlong(got) == (long long)rr_seq) break;
      if (++spins > kSpinLimit) {
        pc->timeout = 1;
        val = 0.0;
        break;
      }
    }
    sh_rr[threadIdx.x] = val;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double g = 0.0;
    for (int r = 0; r < pc->world; ++r) g += sh_rr[r];
    st->gsums[0] = g;
    st->rr_pending = 0;
    if (pc->timeout) st->nonfinite = 2;
    cg_finalize(st, FIN_RR2);
  }
}

__global__ void k_finalize(CGState* st, int which) {
  if (which == FIN_INIT || !st->done) cg_finalize(st, which);
}

// ---- layout conversion at the ABI edge -----------------------------------------------------
// local (lexicographic slab, ghosts included) <- full Ferrite-ordered vector
template <int NC>
__global__ void k_gather_dofs(Geo g, const int* __restrict__ block_of_node, const double* __restrict__ full,
                              double* __restrict__ loc) {
  const long long nloc = (long long)g.S * (g.nown + 2);
  for (long long ln = blockIdx.x * (long long)blockDim.x + threadIdx.x; ln < nloc;
       ln += (long long)gridDim.x * blockDim.x) {
    const int lp = (int)(ln / g.S);
    const int gp = lp + g.p0;
    if (gp < 0 || gp >= g.NPg) {
#pragma unroll
      for (int c = 0; c < NC; ++c) loc[ln * NC + c] = 0.0;
      continue;
    }
    const long long b = block_of_node[ln];
#pragma unroll
    for (int c = 0; c < NC; ++c) loc[ln * NC + c] = full[b * NC + c];
  }
}

// full Ferrite-ordered vector <- owned part of a local vector
template <int NC>
__global__ void k_scatter_dofs(Geo g, const int* __restrict__ block_of_node, const double* __restrict__ loc,
                               double* __restrict__ full) {
  const long long nown = (long long)g.S * g.nown;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < nown;
       t += (long long)gridDim.x * blockDim.x) {
    const long long ln = t + g.S;
    const long long b = block_of_node[ln];
#pragma unroll
    for (int c = 0; c < NC; ++c) full[b * NC + c] = loc[ln * NC + c];
  }
}

// ---- penalisation (src/Utilities/penalties.jl:30-54,113-130; utils.jl:77) ------------------
__global__ void k_penalize(long long n, const double* __restrict__ rho, double* __restrict__ E,
                           double* __restrict__ dE, int kind, double p, double xmin, int pen_first, int proj,
                           double beta) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
       e += (long long)gridDim.x * blockDim.x) {
    const double x = rho[e];
    double a = pen_first ? x : __dadd_rn(__dmul_rn(x, 1.0 - xmin), xmin);  // argument of the penalty
    // ProjectedPenaltyFun: penalty(proj(a)) (penalties.jl:62-69,77-96); da = d proj / d a
    double da = 1.0;
    if (proj == 1) {  // Heaviside: 1 - exp(-beta a) + a exp(-beta)
      const double eb = exp(-beta), ea = exp(-beta * a);
      da = beta * ea + eb;
      a = 1.0 - ea + a * eb;
    } else if (proj == 2) {  // sigmoid: 1 / (1 + exp((beta+1)(0.5 - a)))
      const double ex = exp((beta + 1.0) * (0.5 - a));
      da = (beta + 1.0) * ex / ((1.0 + ex) * (1.0 + ex));
      a = 1.0 / (1.0 + ex);
    }
    double f, df;
    if (kind == 0) {  // x^p
      f = pow(a, p);
      df = p * pow(a, p - 1.0);
    } else if (kind == 1) {  // x / (1 + p (1 - x))
      const double d = 1.0 + p * (1.0 - a);
      f = a / d;
      df = (1.0 + p) / (d * d);
    } else {  // sinh(p x) / sinh(p)
      const double s = sinh(p);
      f = sinh(p * a) / s;
      df = p * cosh(p * a) / s;
    }
    df *= da;
    if (pen_first) {
      // density(var, xmin) = var*(1-xmin) + xmin, two roundings like the Julia expression (no FMA)
      E[e] = __dadd_rn(__dmul_rn(f, 1.0 - xmin), xmin);
      dE[e] = (1.0 - xmin) * df;
    } else {
      E[e] = f;
      dE[e] = df * (1.0 - xmin);
    }
  }
}

// ---- matrix-free operator -------------------------------------------------------------------
template <int DIM>
__host__ __device__ constexpr int corner_local(int ox, int oy, int op) {
  // Ferrite local node of corner (ox, oy, op); in 2-D the slab axis (op) is y.
  return DIM == 3 ? (((ox ^ oy) | (oy << 1)) + 4 * op) : ((ox ^ op) | (op << 1));
}

// One node of y = K(E) x, node-centric gather (matrix_free_operator.jl:66-105): for each of the <= 2^DIM adjacent
// elements (ascending cell id) accumulate row(Ke_bc) . x_e, scale by E_e, then add the element contributions in
// ascending cell order.  Prescribed rows return fixed_diag * x.  t = owned node index; returns the local node index.
// x carries no __restrict__ here: k_cg_persistent writes the vector it passes between its phases, so its loads must stay
// on the coherent path (k_apply's own parameter is const __restrict__, which still gives it the read-only path).
template <int DIM, int NC>
__device__ __forceinline__ long long apply_node(const Geo& g, const double* x, const double* __restrict__ E,
                                                const unsigned char* __restrict__ fixed, double fixed_diag, long long t,
                                                double (&yo)[NC], double (&xo)[NC]) {
  constexpr int NQ = 1 << DIM;
  constexpr int KS = NQ * NC;
  constexpr int NYR = DIM == 3 ? 3 : 1;
  constexpr int NEY = DIM == 3 ? 2 : 1;
  const int inpl = (int)(t % g.S);
  const int lp = (int)(t / g.S) + 1;
  const int i = inpl % g.NX;
  const int j = DIM == 3 ? inpl / g.NX : 0;
  const int gk = lp + g.p0;  // global plane
  double Eq[NQ];
  double acc[NQ][NC];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int ex = q & 1, ey = DIM == 3 ? (q >> 1) & 1 : 0, ep = DIM == 3 ? (q >> 2) & 1 : (q >> 1) & 1;
    const int ei = i - 1 + ex, ej = j - 1 + ey, el = lp - 1 + ep, egl = gk - 1 + ep;
    bool ok = ei >= 0 && ei < g.nx && egl >= 0 && egl < g.NLg;
    if (DIM == 3) ok = ok && ej >= 0 && ej < g.ny;
    Eq[q] = ok ? E[(long long)el * g.SE + (DIM == 3 ? ej * g.nx : 0) + ei] : 0.0;
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[q][c] = 0.0;
  }
  unsigned char fo = 0;
#pragma unroll
  for (int dp = 0; dp < 3; ++dp) {
#pragma unroll
    for (int dy = 0; dy < NYR; ++dy) {
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int ii = i - 1 + dx, jj = DIM == 3 ? j - 1 + dy : 0, pp = lp - 1 + dp, gp = gk - 1 + dp;
        bool ok = ii >= 0 && ii < g.NX && gp >= 0 && gp < g.NPg;
        if (DIM == 3) ok = ok && jj >= 0 && jj < g.NY;
        double xb[NC];
        unsigned char fb = 0;
        if (ok) {
          const long long ln = (long long)pp * g.S + (long long)jj * g.NX + ii;
          fb = fixed[ln];
#pragma unroll
          for (int c = 0; c < NC; ++c) xb[c] = x[ln * NC + c];
        } else {
#pragma unroll
          for (int c = 0; c < NC; ++c) xb[c] = 0.0;
        }
        if (dp == 1 && dx == 1 && (DIM == 2 || dy == 1)) {
          fo = fb;
#pragma unroll
          for (int c = 0; c < NC; ++c) xo[c] = xb[c];
        }
#pragma unroll
        for (int c = 0; c < NC; ++c)
          if (fb & (1 << c)) xb[c] = 0.0;  // bcmatrix: constrained columns are zero
#pragma unroll
        for (int ep = 0; ep < 2; ++ep) {
          const int op = dp - ep;
          if (op < 0 || op > 1) continue;
#pragma unroll
          for (int ey = 0; ey < NEY; ++ey) {
            const int oy = DIM == 3 ? dy - ey : 0;
            if (oy < 0 || oy > 1) continue;
#pragma unroll
            for (int ex = 0; ex < 2; ++ex) {
              const int ox = dx - ex;
              if (ox < 0 || ox > 1) continue;
              const int q = DIM == 3 ? ex + 2 * ey + 4 * ep : ex + 2 * ep;
              const int aloc = corner_local<DIM>(1 - ex, 1 - ey, 1 - ep);
              const int bloc = corner_local<DIM>(ox, oy, op);
#pragma unroll
              for (int c = 0; c < NC; ++c)
#pragma unroll
                for (int c2 = 0; c2 < NC; ++c2)
                  acc[q][c] = fma(cKe[(NC * aloc + c) + KS * (NC * bloc + c2)], xb[c2], acc[q][c]);
            }
          }
        }
      }
    }
  }
  const long long ln = (long long)lp * g.S + inpl;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    double v = 0.0;
#pragma unroll
    for (int q = 0; q < NQ; ++q) v += Eq[q] * acc[q][c];
    if (fo & (1 << c)) v = fixed_diag * xo[c];
    yo[c] = v;
  }
  return ln;
}

// y = K(E) x on the owned nodes.  DOT: also reduce sum_owned x.y into partials and run the CG scalar step `fin`.
template <int DIM, int NC, bool DOT>
__global__ void __launch_bounds__(kBlock) k_apply(Geo g, const double* __restrict__ x, double* __restrict__ y,
                                                  const double* __restrict__ E,
                                                  const unsigned char* __restrict__ fixed, double fixed_diag,
                                                  double* partials, CGState* st, int fin) {
  __shared__ double sm[32];
  if (DOT && st->done) return;
  const long long nown = (long long)g.S * g.nown;
  double dot = 0.0;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < nown;
       t += (long long)gridDim.x * blockDim.x) {
    double v[NC], xo[NC];
    const long long ln = apply_node<DIM, NC>(g, x, E, fixed, fixed_diag, t, v, xo);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      y[ln * NC + c] = v[c];
      if (DOT) dot = fma(xo[c], v[c], dot);
    }
  }
  if (DOT) {
    const double v[1] = {dot};
    block_partials_finish<1>(v, partials, st, fin, sm);
  }
}

// diagonal of K(E) on owned dofs (Jacobi preconditioner)
template <int DIM, int NC>
__global__ void k_diag(Geo g, double* __restrict__ d, const double* __restrict__ E,
                       const unsigned char* __restrict__ fixed, double fixed_diag) {
  constexpr int NQ = 1 << DIM;
  constexpr int KS = NQ * NC;
  const long long nown = (long long)g.S * g.nown;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < nown;
       t += (long long)gridDim.x * blockDim.x) {
    const int inpl = (int)(t % g.S);
    const int lp = (int)(t / g.S) + 1;
    const int i = inpl % g.NX;
    const int j = DIM == 3 ? inpl / g.NX : 0;
    const int gk = lp + g.p0;
    const long long ln = (long long)lp * g.S + inpl;
    const unsigned char fo = fixed[ln];
    double acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] = 0.0;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int ex = q & 1, ey = DIM == 3 ? (q >> 1) & 1 : 0, ep = DIM == 3 ? (q >> 2) & 1 : (q >> 1) & 1;
      const int ei = i - 1 + ex, ej = j - 1 + ey, el = lp - 1 + ep, egl = gk - 1 + ep;
      bool ok = ei >= 0 && ei < g.nx && egl >= 0 && egl < g.NLg;
      if (DIM == 3) ok = ok && ej >= 0 && ej < g.ny;
      if (!ok) continue;
      const double Ee = E[(long long)el * g.SE + (DIM == 3 ? ej * g.nx : 0) + ei];
      const int aloc = corner_local<DIM>(1 - ex, 1 - ey, 1 - ep);
#pragma unroll
      for (int c = 0; c < NC; ++c) acc[c] += Ee * cKe[(NC * aloc + c) * (KS + 1)];
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) d[ln * NC + c] = (fo & (1 << c)) ? fixed_diag : acc[c];
  }
}

// deterministic dense pseudo-random values in (-1, 1) on owned dofs (kernel timing on a live-like vector)
__global__ void __launch_bounds__(kBlock) k_fill_dense(long long off, long long n, double* __restrict__ v) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    unsigned long long z = (unsigned long long)t * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    v[off + t] = (double)(long long)(z >> 11) * (1.0 / 4503599627370496.0) - 1.0;
  }
}

// out = a u + b v on owned dofs
__global__ void __launch_bounds__(kBlock) k_axpby(long long off, long long n, double a, const double* __restrict__ u, double b,
                                                  const double* __restrict__ v, double* __restrict__ out) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
    out[off + t] = fma(a, u[off + t], b * v[off + t]);
}

// ---- CG vector kernels (IterativeSolvers cg!, see cg_finalize) --------------------------------
// r = b, x = 0, p = 0 on owned dofs; sums: r.r [, r.(r/D)]
__global__ void __launch_bounds__(kBlock) k_cg_init(long long off, long long n, const double* __restrict__ b,
                                                    double* __restrict__ x, double* __restrict__ r,
                                                    double* __restrict__ p, const double* __restrict__ D,
                                                    double* partials, CGState* st, int p_is_r, int signal_halo) {
  __shared__ double sm[32];
  double v[2] = {0.0, 0.0};
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n;
       t += (long long)gridDim.x * blockDim.x) {
    const long long k = off + t;
    const double bv = b[k];
    r[k] = bv;
    if (x) x[k] = 0.0;  // x == nullptr: warm start, b already holds b - K x0
    p[k] = p_is_r ? bv : 0.0;  // single-pass variant: the first direction is formed here
    v[0] = fma(bv, bv, v[0]);
    if (D) v[1] = fma(bv, bv / D[k], v[1]);
  }
  block_partials_finish<2>(v, partials, st, FIN_INIT, sm, signal_halo != 0);
}

// p = z + beta p   (z = r, or r/D when preconditioned).  4 independent elements per thread and
// trip keep enough loads in flight to approach the HBM roofline.
__global__ void __launch_bounds__(kBlock) k_update_p(long long off, long long n, const double* __restrict__ r,
                                                     double* __restrict__ p, const double* __restrict__ D,
                                                     CGState* __restrict__ st, int signal_halo) {
  if (st->done) return;
  const double beta = st->beta;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  r += off;
  p += off;
  if (D) D += off;
  for (; t + 3 * stride < n; t += 4 * stride) {
    double rv[4], pv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      rv[k] = r[t + k * stride];
      pv[k] = p[t + k * stride];
    }
    if (D) {
#pragma unroll
      for (int k = 0; k < 4; ++k) rv[k] = rv[k] / D[t + k * stride];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) p[t + k * stride] = fma(beta, pv[k], rv[k]);
  }
  for (; t < n; t += stride) {
    const double z = D ? r[t] / D[t] : r[t];
    p[t] = fma(beta, p[t], z);
  }
  if (signal_halo) {
    // the last CTA to finish tells both slab neighbours that this rank's p is final for this
    // iteration (they read its boundary planes over NVLink inside their K.u kernel)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned int tk = atomicInc(&st->counter, gridDim.x - 1);
      if (tk == gridDim.x - 1) signal_halo_flags(st);
    }
  }
}

// x += alpha p ; r -= alpha Ap ; sums: r.r, r.(r/D), x.r, b.x
template <bool ENERGY>
__global__ void __launch_bounds__(kBlock) k_update_xr(long long off, long long n, double* __restrict__ x,
                                                      double* __restrict__ r, const double* __restrict__ p,
                                                      const double* __restrict__ Ap, const double* __restrict__ D,
                                                      const double* __restrict__ b, double* partials,
                                                      CGState* st) {
  __shared__ double sm[32];
  if (st->done) return;
  const double alpha = st->alpha;
  double v[4] = {0.0, 0.0, 0.0, 0.0};
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  x += off;
  r += off;
  p += off;
  Ap += off;
  if (D) D += off;
  if (ENERGY) b += off;
  auto body = [&](long long k, double xk, double rk, double pk, double apk) {
    const double xv = fma(alpha, pk, xk);
    const double rv = fma(-alpha, apk, rk);
    x[k] = xv;
    r[k] = rv;
    v[0] = fma(rv, rv, v[0]);
    if (D) v[1] = fma(rv, rv / D[k], v[1]);
    if (ENERGY) {
      v[2] = fma(xv, rv, v[2]);
      v[3] = fma(b[k], xv, v[3]);
    }
  };
  for (; t + 3 * stride < n; t += 4 * stride) {
    double xk[4], rk[4], pk[4], ak[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      xk[k] = x[t + k * stride];
      rk[k] = r[t + k * stride];
      pk[k] = p[t + k * stride];
      ak[k] = Ap[t + k * stride];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) body(t + k * stride, xk[k], rk[k], pk[k], ak[k]);
  }
  for (; t < n; t += stride) body(t, x[t], r[t], p[t], Ap[t]);
  block_partials_finish<4>(v, partials, st, FIN_RR, sm);
}

// Single-pass variant (cg_finalize FIN_PAP2 / FIN_RR2): x += alpha p ; r -= alpha Ap ; p = r + beta p ;
// sums: r.r of the NEW residual.  56 bytes per dof instead of 72 for the two separate passes.
__global__ void __launch_bounds__(kBlock) k_update_xrp(long long off, long long n, double* __restrict__ x,
                                                       double* __restrict__ r, double* __restrict__ p,
                                                       const double* __restrict__ Ap, double* partials, CGState* st,
                                                       long long halo_n) {
  __shared__ double sm[32];
  if (st->done) return;
  const double alpha = st->alpha, beta = st->beta;
  double v[1] = {0.0};
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long t0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  x += off;
  r += off;
  p += off;
  Ap += off;
  auto body = [&](long long k, double xk, double rk, double pk, double apk) {
    const double rv = fma(-alpha, apk, rk);
    x[k] = fma(alpha, pk, xk);
    r[k] = rv;
    p[k] = fma(beta, pk, rv);
    v[0] = fma(rv, rv, v[0]);
  };
  auto range = [&](long long lo, long long hi) {
    long long t = lo + t0;
    for (; t + 3 * stride < hi; t += 4 * stride) {
      double xk[4], rk[4], pk[4], ak[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        xk[k] = x[t + k * stride];
        rk[k] = r[t + k * stride];
        pk[k] = p[t + k * stride];
        ak[k] = Ap[t + k * stride];
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) body(t + k * stride, xk[k], rk[k], pk[k], ak[k]);
    }
    for (; t < hi; t += stride) body(t, x[t], r[t], p[t], Ap[t]);
  };
  if (halo_n > 0 && 2 * halo_n < n) {
    // multi-GPU: the two boundary planes first; the last block to finish them tells the slab neighbours that this
    // rank's direction vector is final there, long before their next K.u kernel asks for it
    range(0, halo_n);
    range(n - halo_n, n);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned int tk = atomicInc(&st->counter2, gridDim.x - 1);
      if (tk == gridDim.x - 1) signal_halo_flags(st);
    }
    range(halo_n, n - halo_n);
    block_partials_finish<1>(v, partials, st, FIN_RR2, sm, false);
  } else {
    range(0, n);
    block_partials_finish<1>(v, partials, st, FIN_RR2, sm, halo_n > 0);
  }
}

// ---- persistent CG for small grids (launch-latency bound: 3 launches + 2 device-wide reductions per iteration cost
// ~27 us at configs 1 / 3 for < 2 MB of traffic).  One cooperative launch runs `niter` iterations of IterativeSolvers'
// recurrence (identity preconditioner, default criteria) with three device-wide barriers per iteration; every block sums
// the per-block partials in the same fixed order and runs cg_finalize on its own copy of the scalar state, so no block
// waits for another one's scalars.  All blocks must be co-resident (cudaLaunchCooperativeKernel).
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int nblocks, unsigned int& epoch) {
  __syncthreads();
  epoch += 1;
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    const unsigned int target = epoch * nblocks;
    while (*(volatile unsigned int*)counter < target) {
    }
    __threadfence();
  }
  __syncthreads();
}

template <int DIM, int NC>
__global__ void __launch_bounds__(kBlock) k_cg_persistent(Geo g, double* __restrict__ x, double* __restrict__ r, double* __restrict__ p,
                                                          double* __restrict__ Ap, const double* __restrict__ E,
                                                          const unsigned char* __restrict__ fixed, double fixed_diag, double* partials,
                                                          CGState* st, unsigned int* barrier_counter, int niter) {
  __shared__ double sm[32];
  __shared__ CGState ls;  // this block's copy of the scalar state (identical on every block)
  __shared__ double total;
  const long long off = (long long)g.S * NC, n = (long long)g.S * g.nown * NC, nnodes = (long long)g.S * g.nown;
  const long long tid0 = blockIdx.x * (long long)blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  if (threadIdx.x == 0) ls = *st;
  __syncthreads();
  unsigned int epoch = 0;
  double* part[2] = {partials, partials + gridDim.x};
  // sum of the per-block partials in a fixed order (every block computes the same bits)
  auto reduce_all = [&](const double* pp) -> double {
    double a = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) a += __ldcg(&pp[b]);
    a = block_sum(a, sm);
    if (threadIdx.x == 0) total = a;
    __syncthreads();
    return total;
  };
  for (int it = 0; it < niter; ++it) {
    if (ls.done) break;
    const double beta = ls.beta;
    for (long long t = tid0; t < n; t += stride) p[off + t] = fma(beta, p[off + t], r[off + t]);  // k_update_p
    grid_barrier(barrier_counter, gridDim.x, epoch);
    double dot = 0.0;
    for (long long t = tid0; t < nnodes; t += stride) {  // k_apply<DIM, NC, true>
      double v[NC], xo[NC];
      const long long ln = apply_node<DIM, NC>(g, p, E, fixed, fixed_diag, t, v, xo);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        Ap[ln * NC + c] = v[c];
        dot = fma(xo[c], v[c], dot);
      }
    }
    dot = block_sum(dot, sm);
    if (threadIdx.x == 0) part[0][blockIdx.x] = dot;
    grid_barrier(barrier_counter, gridDim.x, epoch);
    const double pAp = reduce_all(part[0]);
    if (threadIdx.x == 0) {
      ls.gsums[0] = ls.sums[0] = pAp;
      cg_finalize(&ls, FIN_PAP);
    }
    __syncthreads();
    const double alpha = ls.alpha;
    double rr = 0.0;
    for (long long t = tid0; t < n; t += stride) {  // k_update_xr<false>
      const long long k = off + t;
      x[k] = fma(alpha, p[k], x[k]);
      const double rv = fma(-alpha, Ap[k], r[k]);
      r[k] = rv;
      rr = fma(rv, rv, rr);
    }
    rr = block_sum(rr, sm);
    if (threadIdx.x == 0) part[1][blockIdx.x] = rr;
    grid_barrier(barrier_counter, gridDim.x, epoch);
    const double rrs = reduce_all(part[1]);
    if (threadIdx.x == 0) {
      ls.gsums[0] = ls.sums[0] = rrs;
      cg_finalize(&ls, FIN_RR);
    }
    __syncthreads();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *st = ls;
}

// plain dot over owned dofs -> st->sums[0]
__global__ void __launch_bounds__(kBlock) k_dot(long long off, long long n, const double* __restrict__ a,
                                                const double* __restrict__ b, double* partials, CGState* st) {
  __shared__ double sm[32];
  double v[1] = {0.0};
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n;
       t += (long long)gridDim.x * blockDim.x)
    v[0] = fma(a[off + t], b[off + t], v[0]);
  block_partials_finish<1>(v, partials, st, FIN_PLAIN, sm);
}

// zero prescribed dofs of a local vector using the node flags (apply_zero!)
template <int NC>
__global__ void k_apply_zero(long long nnodes_loc, const unsigned char* __restrict__ fixed, double* __restrict__ v) {
  for (long long ln = blockIdx.x * (long long)blockDim.x + threadIdx.x; ln < nnodes_loc;
       ln += (long long)gridDim.x * blockDim.x) {
    const unsigned char f = fixed[ln];
    if (!f) continue;
#pragma unroll
    for (int c = 0; c < NC; ++c)
      if (f & (1 << c)) v[ln * NC + c] = 0.0;
  }
}

// ---- compliance / sensitivity (compute_element_energy.jl:18-38, thermal_compliance.jl:144-155)
// c_e = v_e' Ke u_e with the raw Ke (v = u for compliance); grad_e = gsign * dE_e c_e;
// sums[0] = sum E_e c_e.
template <int DIM, int NC>
__global__ void __launch_bounds__(kBlock) k_sens(Geo g, const double* __restrict__ u, const double* __restrict__ v,
                                                 const double* __restrict__ E, const double* __restrict__ dE,
                                                 double* __restrict__ cell, double* __restrict__ grad, double gsign,
                                                 double* partials, CGState* st) {
  constexpr int NQ = 1 << DIM;
  constexpr int KS = NQ * NC;
  __shared__ double sm[32];
  const long long nel = (long long)g.SE * g.nlay;
  double obj[1] = {0.0};
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < nel;
       t += (long long)gridDim.x * blockDim.x) {
    const int inl = (int)(t % g.SE);
    const int ll = (int)(t / g.SE) + 1;
    const int i = inl % g.nx;
    const int j = DIM == 3 ? inl / g.nx : 0;
    double ue[KS], ve[KS];
#pragma unroll
    for (int a = 0; a < NQ; ++a) {
      // inverse of corner_local: local node a -> corner offsets
      const int az = DIM == 3 ? a >> 2 : 0;
      const int a4 = a & 3;
      const int o2 = a4 >> 1;          // second in-ring axis
      const int ox = (a4 & 1) ^ o2;    // x
      const int oy = DIM == 3 ? o2 : 0;
      const int op = DIM == 3 ? az : o2;
      const long long ln = (long long)(ll + op) * g.S + (long long)(j + oy) * g.NX + (i + ox);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        ue[a * NC + c] = u[ln * NC + c];
        ve[a * NC + c] = v[ln * NC + c];
      }
    }
    double ce = 0.0;
#pragma unroll
    for (int w = 0; w < KS; ++w) {
#pragma unroll
      for (int r = 0; r < KS; ++r) ce = fma(ve[r] * cKe[r + KS * w], ue[w], ce);
    }
    const long long le = (long long)ll * g.SE + inl;
    if (cell) cell[le] = ce;
    if (grad) grad[le] = gsign * dE[le] * ce;
    obj[0] = fma(E[le], ce, obj[0]);
  }
  block_partials_finish<1>(obj, partials, st, FIN_PLAIN, sm);
}

// ---- filter (CheqFilters.jl:66-118, density_filter.jl:49-108, sens_filter.jl:72-110) ---------
// On a uniform grid the reference's two stages collapse to
//   s_n = sum_{c in cells(n)} x_c                       (cell -> node; the 1/m_n of the mean cancels
//   y_i = sum_n w(n-i) s_n / den_i                       against the duplicate multiplicity m_n)
//   den_i = sum_n m_n w(n-i),  w = max(rmin - dist, 0) for dist < rmin, m_n = |cells(n)|.
struct FilterGeo {
  int nx, ny, nz;  // cells
  int NX, NY, NZ;  // nodes
  int R[3];        // node offsets o in [1-R, R] per axis (R[2] = 0 -> single plane in 2-D)
  int dim;
};

__global__ void k_filter_c2n(FilterGeo f, const double* __restrict__ x, double* __restrict__ s) {
  const long long nn = (long long)f.NX * f.NY * f.NZ;
  for (long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x; n < nn;
       n += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(n % f.NX), j = (int)((n / f.NX) % f.NY), k = (int)(n / ((long long)f.NX * f.NY));
    double acc = 0.0;
    for (int dk = (f.dim == 3 ? -1 : 0); dk <= 0; ++dk)
      for (int dj = -1; dj <= 0; ++dj)
        for (int di = -1; di <= 0; ++di) {
          const int ci = i + di, cj = j + dj, ck = k + dk;
          if (ci < 0 || ci >= f.nx || cj < 0 || cj >= f.ny || ck < 0 || ck >= f.nz) continue;
          acc += x[ci + (long long)f.nx * (cj + (long long)f.ny * ck)];
        }
    s[n] = acc;
  }
}

// MODE 0: y_i = sum_n w s_n / den_i ; MODE 1: den_i = sum_n m_n w (setup)
template <int MODE>
__global__ void k_filter_n2c(FilterGeo f, const double* __restrict__ wtab, const double* __restrict__ s,
                             const double* __restrict__ den, double* __restrict__ y) {
  const long long ne = (long long)f.nx * f.ny * f.nz;
  const int wx = 2 * f.R[0], wy = 2 * f.R[1], wz = f.dim == 3 ? 2 * f.R[2] : 1;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < ne;
       e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e % f.nx), j = (int)((e / f.nx) % f.ny), k = (int)(e / ((long long)f.nx * f.ny));
    double acc = 0.0;
    for (int tz = 0; tz < wz; ++tz) {
      const int nk = f.dim == 3 ? k + 1 - f.R[2] + tz : 0;
      if (nk < 0 || nk >= f.NZ) continue;
      const int mz = (f.dim == 3 && nk > 0 && nk < f.NZ - 1) ? 2 : 1;
      for (int ty = 0; ty < wy; ++ty) {
        const int nj = j + 1 - f.R[1] + ty;
        if (nj < 0 || nj >= f.NY) continue;
        const int my = (nj > 0 && nj < f.NY - 1) ? 2 : 1;
        for (int tx = 0; tx < wx; ++tx) {
          const int ni = i + 1 - f.R[0] + tx;
          if (ni < 0 || ni >= f.NX) continue;
          const double w = wtab[tx + wx * (ty + wy * tz)];
          if (MODE == 0) {
            acc = fma(w, s[ni + (long long)f.NX * (nj + (long long)f.NY * nk)], acc);
          } else {
            const int mx = (ni > 0 && ni < f.NX - 1) ? 2 : 1;
            acc = fma(w, (double)(mx * my * mz), acc);
          }
        }
      }
    }
    y[e] = MODE == 0 ? acc / den[e] : acc;
  }
}

// Shared-memory tiled radius stencil used by both filter directions:
//   forward   (TRANSPOSE = false): out = cells, in = nodal sums s;  y_i = sum_t w[t] s[i + 1 - R + t] / den_i
//   transpose (TRANSPOSE = true):  out = nodes, in = cells (d/den); t_n = sum_t w[t] (d/den)[n - R + t]
// (the cone table is symmetric, w[t] = w[2R-1-t], so both directions index it with t).  A CTA owns a
// BX x BY x BZ block of outputs and stages the (B + 2R - 1)^dim input tile once.
template <bool TRANSPOSE>
__global__ void __launch_bounds__(kBlock) k_filter_stencil(FilterGeo f, int BY, int BZ, const double* __restrict__ wtab,
                                                           const double* __restrict__ in, const double* __restrict__ den,
                                                           double* __restrict__ out) {
  extern __shared__ double tile[];
  constexpr int BX = 32;
  const int ox_n = TRANSPOSE ? f.NX : f.nx, oy_n = TRANSPOSE ? f.NY : f.ny, oz_n = TRANSPOSE ? f.NZ : f.nz;  // outputs
  const int ix_n = TRANSPOSE ? f.nx : f.NX, iy_n = TRANSPOSE ? f.ny : f.NY, iz_n = TRANSPOSE ? f.nz : f.NZ;  // inputs
  const int wx = 2 * f.R[0], wy = 2 * f.R[1], wz = f.dim == 3 ? 2 * f.R[2] : 1;
  const int WX = BX + wx - 1, WY = BY + wy - 1, WZ = BZ + wz - 1;
  double* wsm = tile + (size_t)WX * WY * WZ;  // cone table copy
  const int tilesX = (ox_n + BX - 1) / BX, tilesY = (oy_n + BY - 1) / BY;
  int b = blockIdx.x;
  const int bx = b % tilesX;
  b /= tilesX;
  const int by = b % tilesY, bz = b / tilesY;
  const int o0x = bx * BX, o0y = by * BY, o0z = bz * BZ;
  const int basex = (TRANSPOSE ? -f.R[0] : 1 - f.R[0]), basey = (TRANSPOSE ? -f.R[1] : 1 - f.R[1]);
  const int basez = f.dim == 3 ? (TRANSPOSE ? -f.R[2] : 1 - f.R[2]) : 0;
  const int nt = WX * WY * WZ;
  for (int k = threadIdx.x; k < nt; k += blockDim.x) {
    const int tx = k % WX, ty = (k / WX) % WY, tz = k / (WX * WY);
    const int gx = o0x + basex + tx, gy = o0y + basey + ty, gz = o0z + basez + tz;
    double v = 0.0;
    if (gx >= 0 && gx < ix_n && gy >= 0 && gy < iy_n && gz >= 0 && gz < iz_n) {
      const long long gi = gx + (long long)ix_n * (gy + (long long)iy_n * gz);
      v = TRANSPOSE ? in[gi] / den[gi] : in[gi];
    }
    tile[k] = v;
  }
  for (int k = threadIdx.x; k < wx * wy * wz; k += blockDim.x) wsm[k] = wtab[k];
  __syncthreads();
  const int lx = threadIdx.x % BX, ly = (threadIdx.x / BX) % BY, lz = threadIdx.x / (BX * BY);
  const int ox = o0x + lx, oy = o0y + ly, oz = o0z + lz;
  if (ox >= ox_n || oy >= oy_n || oz >= oz_n) return;
  double acc = 0.0;
  for (int tz = 0; tz < wz; ++tz)
    for (int ty = 0; ty < wy; ++ty) {
      const double* row = tile + ((size_t)(lz + tz) * WY + (ly + ty)) * WX + lx;
      const double* wr = wsm + (tz * wy + ty) * wx;
      for (int tx = 0; tx < wx; ++tx) acc = fma(wr[tx], row[tx], acc);
    }
  const long long oi = ox + (long long)ox_n * (oy + (long long)oy_n * oz);
  out[oi] = TRANSPOSE ? acc : acc / den[oi];
}

// Specialisation for cubic windows (R equal on all axes, 3-D): the cone table sits in constant
// memory and every loop is unrolled, so each weight is an immediate constant-bank operand of its
// DFMA; a thread owns a z-column of 4 outputs and reuses every staged value for up to 4 of them
// (28 shared-memory reads per output instead of 64 for R = 2).
__constant__ double cW[216];

template <int R, bool TRANSPOSE>
__global__ void __launch_bounds__(128) k_filter_stencil3(FilterGeo f, const double* __restrict__ in,
                                                          const double* __restrict__ den, double* __restrict__ out) {
  constexpr int BX = 32, BY = 4, BZ = 4, W = 2 * R;
  constexpr int WX = BX + W - 1, WY = BY + W - 1, WZ = BZ + W - 1;
  __shared__ double tile[WZ][WY][WX];
  const int ox_n = TRANSPOSE ? f.NX : f.nx, oy_n = TRANSPOSE ? f.NY : f.ny, oz_n = TRANSPOSE ? f.NZ : f.nz;
  const int ix_n = TRANSPOSE ? f.nx : f.NX, iy_n = TRANSPOSE ? f.ny : f.NY, iz_n = TRANSPOSE ? f.nz : f.NZ;
  const int tilesX = (ox_n + BX - 1) / BX, tilesY = (oy_n + BY - 1) / BY;
  int b = blockIdx.x;
  const int bx = b % tilesX;
  b /= tilesX;
  const int by = b % tilesY, bz = b / tilesY;
  const int o0x = bx * BX, o0y = by * BY, o0z = bz * BZ;
  constexpr int base = TRANSPOSE ? -R : 1 - R;
  for (int k = threadIdx.x; k < WX * WY * WZ; k += blockDim.x) {
    const int tx = k % WX, ty = (k / WX) % WY, tz = k / (WX * WY);
    const int gx = o0x + base + tx, gy = o0y + base + ty, gz = o0z + base + tz;
    double v = 0.0;
    if (gx >= 0 && gx < ix_n && gy >= 0 && gy < iy_n && gz >= 0 && gz < iz_n) {
      const long long gi = gx + (long long)ix_n * (gy + (long long)iy_n * gz);
      v = TRANSPOSE ? in[gi] / den[gi] : in[gi];
    }
    tile[tz][ty][tx] = v;
  }
  __syncthreads();
  const int lx = threadIdx.x % BX, ly = threadIdx.x / BX;
  const int ox = o0x + lx, oy = o0y + ly;
  if (ox >= ox_n || oy >= oy_n) return;
  double acc[BZ];
#pragma unroll
  for (int k = 0; k < BZ; ++k) acc[k] = 0.0;
#pragma unroll
  for (int tz = 0; tz < WZ; ++tz)
#pragma unroll
    for (int ty = 0; ty < W; ++ty)
#pragma unroll
      for (int tx = 0; tx < W; ++tx) {
        const double v = tile[tz][ly + ty][lx + tx];
#pragma unroll
        for (int k = 0; k < BZ; ++k)
          if (tz - k >= 0 && tz - k < W) acc[k] = fma(cW[((tz - k) * W + ty) * W + tx], v, acc[k]);
      }
#pragma unroll
  for (int k = 0; k < BZ; ++k) {
    const int oz = o0z + k;
    if (oz < oz_n) {
      const long long oi = ox + (long long)ox_n * (oy + (long long)oy_n * oz);
      out[oi] = TRANSPOSE ? acc[k] : acc[k] / den[oi];
    }
  }
}

// transpose, stage 2': t_n = sum_i w(n-i) d_i / den_i   (gather over cells around node n)
__global__ void k_filter_c2n_T(FilterGeo f, const double* __restrict__ wtab, const double* __restrict__ d,
                               const double* __restrict__ den, double* __restrict__ t) {
  const long long nn = (long long)f.NX * f.NY * f.NZ;
  const int wx = 2 * f.R[0], wy = 2 * f.R[1], wz = f.dim == 3 ? 2 * f.R[2] : 1;
  for (long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x; n < nn;
       n += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(n % f.NX), j = (int)((n / f.NX) % f.NY), k = (int)(n / ((long long)f.NX * f.NY));
    double acc = 0.0;
    // node n = cell + 1 - R + t  ->  cell = n - 1 + R - t
    for (int tz = 0; tz < wz; ++tz) {
      const int ck = f.dim == 3 ? k - 1 + f.R[2] - tz : 0;
      if (ck < 0 || ck >= f.nz) continue;
      for (int ty = 0; ty < wy; ++ty) {
        const int cj = j - 1 + f.R[1] - ty;
        if (cj < 0 || cj >= f.ny) continue;
        for (int tx = 0; tx < wx; ++tx) {
          const int ci = i - 1 + f.R[0] - tx;
          if (ci < 0 || ci >= f.nx) continue;
          const long long e = ci + (long long)f.nx * (cj + (long long)f.ny * ck);
          acc = fma(wtab[tx + wx * (ty + wy * tz)], d[e] / den[e], acc);
        }
      }
    }
    t[n] = acc;
  }
}

// transpose, stage 1': y_c = sum_{n in c} t_n
__global__ void k_filter_n2c_T(FilterGeo f, const double* __restrict__ t, double* __restrict__ y) {
  const long long ne = (long long)f.nx * f.ny * f.nz;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < ne;
       e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e % f.nx), j = (int)((e / f.nx) % f.ny), k = (int)(e / ((long long)f.nx * f.ny));
    double acc = 0.0;
    for (int dk = 0; dk <= (f.dim == 3 ? 1 : 0); ++dk)
      for (int dj = 0; dj <= 1; ++dj)
        for (int di = 0; di <= 1; ++di) acc += t[(i + di) + (long long)f.NX * ((j + dj) + (long long)f.NY * (k + dk))];
    y[e] = acc;
  }
}

// ---- optimality-criteria design update on the device ------------------------------
// xn_e = max(xlo, max(x_e - move, min(1, min(x_e + move, x_e (-dc_e / (dv_e lambda))^eta))))  with dc clamped to <= 0.
__device__ __forceinline__ double oc_candidate(double x, double dc, double dv, double lam, double move, double eta, double xlo) {
  const double dcm = fmin(dc, 0.0);
  const double b = x * pow(-dcm / dv / lam, eta);
  return fmax(xlo, fmax(x - move, fmin(1.0, fmin(x + move, b))));
}
// APPLY = false: sums[0] = sum xn_e dv_e for the trial multiplier (one bisection step);
// APPLY = true : writes xn and reduces max |xn - x| (as a sum of per-block maxima is not a max: see below)
template <bool APPLY>
__global__ void __launch_bounds__(kBlock) k_oc(long long n, const double* __restrict__ x, const double* __restrict__ dc,
                                               const double* __restrict__ dv, double lam, double move, double eta, double xlo,
                                               double* __restrict__ xout, double* partials, CGState* st) {
  __shared__ double sm[32];
  double v[2] = {0.0, 0.0};
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const double xe = x[e];
    const double xn = oc_candidate(xe, dc[e], dv[e], lam, move, eta, xlo);
    v[0] = fma(xn, dv[e], v[0]);
    if (APPLY) {
      xout[e] = xn;
      const double d = fabs(xn - xe);
      v[1] = fma(d, d, v[1]);  // sum of squared changes (the max is taken on the host copy when asked for)
    }
  }
  block_partials_finish<2>(v, partials, st, FIN_PLAIN, sm);
}

// ---- assembled path (assemble.jl:28-90; Ferrite assemble!/apply!) -----------------------------
// Internal CSR in lexicographic dof order.  Row (n,c) holds, for every in-range neighbour node m
// (z,y,x ascending) and component c2, K[(n,c),(m,c2)].  nbr_start[n] = number of (node,neighbour)
// pairs before node n, so row (n,c) starts at NC*NC*nbr_start[n] + c*NC*cnt(n).
template <int DIM>
__device__ __forceinline__ int nbr_count(const Geo& g, int i, int j, int k) {
  const int cx = 1 + (i > 0) + (i < g.NX - 1);
  const int cp = 1 + (k > 0) + (k < g.NPg - 1);
  if (DIM == 2) return cx * cp;
  const int cy = 1 + (j > 0) + (j < g.NY - 1);
  return cx * cy * cp;
}

// one thread per (node, neighbour offset): value = sum over shared elements (ascending cell id)
// of fl(E_e * Ke[a,b]) -- the exact arithmetic of Ferrite's assemble!(assembler, dofs, px*Ke).
template <int DIM, int NC>
__global__ void k_assemble(Geo g, const long long* __restrict__ nbr_start, const double* __restrict__ E,
                           double* __restrict__ nz, int* __restrict__ col) {
  constexpr int NQ = 1 << DIM;
  constexpr int KS = NQ * NC;
  constexpr int NN = DIM == 3 ? 27 : 9;
  const long long total = (long long)g.S * g.NPg * NN;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long n = t / NN;
    const int o = (int)(t % NN);
    const int dx = o % 3, dy = DIM == 3 ? (o / 3) % 3 : 1, dp = DIM == 3 ? o / 9 : o / 3;
    const int inpl = (int)(n % g.S), k = (int)(n / g.S);
    const int i = inpl % g.NX, j = DIM == 3 ? inpl / g.NX : 0;
    const int ii = i - 1 + dx, jj = DIM == 3 ? j - 1 + dy : 0, kk = k - 1 + dp;
    bool ok = ii >= 0 && ii < g.NX && kk >= 0 && kk < g.NPg;
    if (DIM == 3) ok = ok && jj >= 0 && jj < g.NY;
    if (!ok) continue;
    // slot of this neighbour among the valid ones (z,y,x ascending)
    int slot = 0;
    for (int o2 = 0; o2 < o; ++o2) {
      const int ex = o2 % 3, ey = DIM == 3 ? (o2 / 3) % 3 : 1, ez = DIM == 3 ? o2 / 9 : o2 / 3;
      const int a = i - 1 + ex, b = DIM == 3 ? j - 1 + ey : 0, c = k - 1 + ez;
      bool v = a >= 0 && a < g.NX && c >= 0 && c < g.NPg;
      if (DIM == 3) v = v && b >= 0 && b < g.NY;
      slot += v ? 1 : 0;
    }
    const int cnt = nbr_count<DIM>(g, i, j, k);
    double val[NC][NC];
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int c2 = 0; c2 < NC; ++c2) val[c][c2] = 0.0;
    for (int q = 0; q < NQ; ++q) {  // ascending cell id
      const int ex = q & 1, ey = DIM == 3 ? (q >> 1) & 1 : 0, ep = DIM == 3 ? (q >> 2) & 1 : (q >> 1) & 1;
      const int ox = dx - ex, oy = DIM == 3 ? dy - ey : 0, op = dp - ep;
      if (ox < 0 || ox > 1 || oy < 0 || oy > 1 || op < 0 || op > 1) continue;
      const int ei = i - 1 + ex, ej = j - 1 + ey, el = k - 1 + ep;
      bool eok = ei >= 0 && ei < g.nx && el >= 0 && el < g.NLg;
      if (DIM == 3) eok = eok && ej >= 0 && ej < g.ny;
      if (!eok) continue;
      // single-GPU path: local layer index = global layer + 1 (ghost layer 0)
      const double Ee = E[(long long)(el + 1) * g.SE + (DIM == 3 ? ej * g.nx : 0) + ei];
      const int aloc = corner_local<DIM>(1 - ex, 1 - ey, 1 - ep);
      const int bloc = corner_local<DIM>(ox, oy, op);
#pragma unroll
      for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int c2 = 0; c2 < NC; ++c2)
          val[c][c2] = __dadd_rn(val[c][c2], __dmul_rn(Ee, cKe[(NC * aloc + c) + KS * (NC * bloc + c2)]));
    }
    const long long m = (long long)kk * g.S + (long long)jj * g.NX + ii;
    const long long base = (long long)NC * NC * nbr_start[n];
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int c2 = 0; c2 < NC; ++c2) {
        const long long pos = base + (long long)c * NC * cnt + slot * NC + c2;
        nz[pos] = val[c][c2];
        col[pos] = (int)(m * NC + c2);
      }
  }
}

// sum |K_dd| over all rows -> st->sums[0]  (Ferrite meandiag numerator)
template <int DIM, int NC>
__global__ void __launch_bounds__(kBlock) k_csr_absdiag(Geo g, const long long* __restrict__ nbr_start,
                                                        const double* __restrict__ nz, double* __restrict__ diag,
                                                        double* partials, CGState* st) {
  __shared__ double sm[32];
  const long long nn = (long long)g.S * g.NPg;
  double v[1] = {0.0};
  for (long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x; n < nn;
       n += (long long)gridDim.x * blockDim.x) {
    const int inpl = (int)(n % g.S), k = (int)(n / g.S);
    const int i = inpl % g.NX, j = DIM == 3 ? inpl / g.NX : 0;
    const int cnt = nbr_count<DIM>(g, i, j, k);
    // slot of the node itself = number of valid neighbours that precede it
    const int bx = (i > 0), by = DIM == 3 ? (j > 0) : 0, bp = (k > 0);
    const int cx = 1 + (i > 0) + (i < g.NX - 1);
    const int cy = DIM == 3 ? 1 + (j > 0) + (j < g.NY - 1) : 1;
    const int self = bp * cx * cy + by * cx + bx;
    const long long base = (long long)NC * NC * nbr_start[n];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const double dv = nz[base + (long long)c * NC * cnt + self * NC + c];
      if (diag) diag[n * NC + c] = dv;
      v[0] += fabs(dv);
    }
  }
  block_partials_finish<1>(v, partials, st, FIN_PLAIN, sm);
}

// Ferrite apply!(K, f, ch) for homogeneous constraints: zero prescribed rows and columns, put
// m = mean|diag K| on their diagonal.  `fixed` here is indexed by global lexicographic node.
template <int NC>
__global__ void k_csr_apply_bc(long long nrows, const int* __restrict__ rowptr, const int* __restrict__ col,
                               double* __restrict__ nz, const unsigned char* __restrict__ fixed, const CGState* st,
                               double inv_n) {
  const double m = st->sums[0] * inv_n;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < nrows;
       r += (long long)gridDim.x * blockDim.x) {
    const bool rf = (fixed[r / NC] >> (r % NC)) & 1;
    for (int p = rowptr[r]; p < rowptr[r + 1]; ++p) {
      const int c = col[p];
      const bool cf = (fixed[c / NC] >> (c % NC)) & 1;
      if (rf || cf) nz[p] = (c == r) ? m : 0.0;
    }
  }
}

// y = K x, CSR, LANES lanes cooperate on one row; optional fused dot x.y
template <int LANES, bool DOT>
__global__ void __launch_bounds__(kBlock) k_spmv(long long nrows, const int* __restrict__ rowptr,
                                                 const int* __restrict__ col, const double* __restrict__ nz,
                                                 const double* __restrict__ x, double* __restrict__ y,
                                                 double* partials, CGState* st, int fin) {
  __shared__ double sm[32];
  if (DOT && st->done) return;
  const int lane = threadIdx.x % LANES;
  const long long rows_per_pass = (long long)gridDim.x * (blockDim.x / LANES);
  double dot = 0.0;
  for (long long r0 = (long long)blockIdx.x * (blockDim.x / LANES); r0 < nrows; r0 += rows_per_pass) {
    const long long r = r0 + threadIdx.x / LANES;
    double acc = 0.0;
    if (r < nrows) {
      const int b = rowptr[r], e = rowptr[r + 1];
      for (int p = b + lane; p < e; p += LANES) acc = fma(nz[p], x[col[p]], acc);
    }
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o, LANES);
    if (lane == 0 && r < nrows) {
      y[r] = acc;
      if (DOT) dot = fma(x[r], acc, dot);
    }
  }
  if (DOT) {
    const double v[1] = {dot};
    block_partials_finish<1>(v, partials, st, fin, sm);
  }
}

}  // namespace topopt
