// C ABI of libtopopt_cuda (see include/topopt_cuda.h).  Host orchestration of the kernels in
// kernels.cuh: handle lifetime, layout conversion at the ABI edge, the CG loop, the slab halo
// exchange and reductions over NCCL.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.h"
#include "kernels.cuh"
#include "kxu_hex8.cuh"
#include "kxu_hex8_2row.cuh"
#include "kxu_hex8_ring.cuh"
#include "kxu_hex8_cgfused.cuh"
#include "kxu_hex8_cgtma.cuh"
#include "multigrid.cuh"

using namespace topopt;

// ---- NCCL, loaded lazily so that the library also loads where NCCL is absent -----------------
namespace {
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load(std::string& err) {
    if (lib) return true;
    lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) {
      err = std::string("cannot load libnccl.so.2: ") + dlerror();
      return false;
    }
#define TOPOPT_SYM(name) \
  name = reinterpret_cast<decltype(name)>(dlsym(lib, "nccl" #name)); \
  if (!name) { err = "libnccl lacks nccl" #name; return false; }
    TOPOPT_SYM(GetUniqueId)
    TOPOPT_SYM(CommInitRank)
    TOPOPT_SYM(CommDestroy)
    TOPOPT_SYM(AllReduce)
    TOPOPT_SYM(Send)
    TOPOPT_SYM(Recv)
    TOPOPT_SYM(GroupStart)
    TOPOPT_SYM(GroupEnd)
    TOPOPT_SYM(GetErrorString)
#undef TOPOPT_SYM
    return true;
  }
};
NcclApi g_nccl;
}  // namespace

// one level of the geometric multigrid hierarchy (multigrid.cuh); level 0 aliases the handle's own grid
struct MGLevel {
  Geo g{};
  int64_t nloc_nodes = 0, nloc_dofs = 0, off = 0, nown_dofs = 0, nloc_el = 0;
  double *E = nullptr, *D = nullptr, *x = nullptr, *b = nullptr, *r = nullptr, *d = nullptr, *t = nullptr;
  unsigned char* fixed = nullptr;
  double lmax = 2.0;  // estimate of lambda_max(D^-1 K)
  bool owns = false;  // E / fixed / b allocated by the hierarchy (levels >= 1)
};
struct MGHierarchy {
  std::vector<MGLevel> lv;
  double* d_Ainv = nullptr;  // dense inverse of the coarsest operator
  int* d_loc = nullptr;      // dense dof -> local vector index on the coarsest level
  int ncoarse = 0;
  int degree = 2;            // Chebyshev smoother degree
  double ratio = 6.0;        // smoothed part of the spectrum: [lambda_max / ratio, lambda_max]
  long long cycles = 0;
};

struct topopt_handle {
  uint64_t id = 0;
  int dim = 0, nc = 0, ks = 0;
  GridDims gd{};
  Geo g{};
  int device = 0, rank = 0, world = 1;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  ncclComm_t comm = nullptr;
  int64_t ndof = 0, nel = 0, nnodes = 0, nnz = 0;
  int64_t nloc_nodes = 0, nloc_dofs = 0, off = 0, nown_dofs = 0;  // dof vectors
  int64_t nloc_el = 0, eoff = 0, nown_el = 0;                     // element vectors
  int64_t plane_dofs = 0;
  double Ke[kMaxKe * kMaxKe];
  double Kh[48];          // modal coefficients (hex8 elasticity fast path)
  bool modal_ok = false;  // Ke has the brick/isotropic modal sparsity pattern
  bool modal_cube = false;  // ... and the 8-value structure of cubic cells (kxu_hex8_ring.cuh: cKc)
  double Kc[8] = {0};
  int kxu_ty = 16, kxu_waves = 1, kxu_nsync = 1, kxu_2row = 1;  // 2row: 0 = one-row kernel, 1 = auto, else thread rows
  int kxu_ring = 1;       // ring-staged kernel for premasked inputs (CG directions): 0 = off, 1 = auto, else thread rows
  int kxu_cube = 1;       // ring kernel: use the cubic-cell specialisation of the modal matrix when it applies
  int kxu_ring_min = 12;  // fewest owned node planes per rank for which the ring kernel is selected
  int kxu_grid = 0;       // ring kernel: explicit CTA count (0 = 148 x waves)
  int cg_fused_grid = 0;  // one-kernel iteration: explicit CTA count (0 = aligned_grid())
  int cg_fused_pdl = 1;  // one-kernel iteration: programmatic dependent launch (prologue overlaps the previous kernel's tail)
  int cg_persist = 1;  // small single-GPU grids: whole batches of CG iterations in one cooperative kernel (k_cg_persistent)
  unsigned int* d_barrier = nullptr;
  bool cg_fused_single_only = false;  // TOPOPT_CG_FUSED_MGPU=0: ranks of a multi-GPU run keep the two-kernel iteration
  int cg_fused = 1;       // single GPU, single-pass recurrence: one kernel per CG iteration (kxu_hex8_cgfused.cuh); 0 = off, else thread rows
  int cg_variant_env = -1;  // TOPOPT_CG_VARIANT overrides topopt_cg_opts.variant (diagnostics)
  double fixed_diag = 0.0, cellvol = 1.0;
  double sizes[3] = {1, 1, 1};
  // device buffers
  int* d_block = nullptr;
  unsigned char* d_fixed = nullptr;
  double *d_b = nullptr, *d_fload = nullptr, *d_u = nullptr, *d_r = nullptr, *d_p = nullptr, *d_p2 = nullptr, *d_Ap = nullptr;
  double *d_r2 = nullptr, *d_Ap2 = nullptr;  // ping-pong partners of d_r / d_Ap (fused CG iteration kernel)
  // peer-memory communication (cudaIpc): in-kernel allreduce + direct halo reads
  PeerBlock* d_peerblock = nullptr;  // this rank's block (IPC-exported)
  PeerComm* d_peercomm = nullptr;    // device copy of the mapping table
  void* peer_mapped[3 * kMaxRanks] = {nullptr};
  // the neighbours' p, p2, r, r2, Ap, Ap2 (mapped over cudaIpc) for the one-kernel CG iteration; [0] duplicates peer_p_*
  const double* peer_lo[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  const double* peer_hi[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  void* peer_mapped_vec[2][6] = {{nullptr}};
  bool peer_fused_ready = false;
  CUtensorMap* d_tmaps = nullptr;  // tensor maps of the CG vectors (kxu_hex8_cgtma.cuh), built at the first one-kernel iteration
  int tmaps_rows = 0;              // thread rows the maps' boxes were built for
  int cg_fused_tma = 0;            // 1: stage planes with tensor-map copies (kxu_hex8_cgtma.cuh; 219 us vs 202 us at config 4), 0: row-wise bulk copies
  const double* peer_p_lo = nullptr;   // lower neighbour's d_p (mapped)
  const double* peer_p_hi = nullptr;   // upper neighbour's d_p (mapped)
  bool peer_ready = false;
  int nown_lower = 0;
  // CUDA-graph cache for one batch of CG iterations (same kernel arguments every batch)
  cudaGraphExec_t cg_graph = nullptr;
  unsigned long long cg_graph_key = 0;
  long long cg_graph_launches = 0;
  bool use_graphs = true;
  bool no_fuse = true;  // fusing p = r + beta p into K.u measured slower (LSU-bound kernel); opt in with TOPOPT_FUSE_P=1
  double *d_D = nullptr, *d_rhs = nullptr, *d_lam = nullptr, *d_tmp = nullptr;
  double *d_E = nullptr, *d_dE = nullptr, *d_rho = nullptr, *d_cell = nullptr, *d_grad = nullptr;
  double *d_full_dof = nullptr, *d_full_el = nullptr, *d_design = nullptr, *d_xf = nullptr, *d_gfull = nullptr;
  double *d_dv = nullptr, *d_dc = nullptr, *d_xnew = nullptr;  // on-device design update
  const double* d_lastgrad = nullptr;                           // gradient w.r.t. the design left by topopt_simp_eval
  MGHierarchy* mg = nullptr;   // geometric multigrid preconditioner (opt-in)
  bool mg_dirty = true;        // coarse operators / smoother bounds must be rebuilt (stiffness changed)
  const MGLevel* lvl = nullptr;  // K.u launchers act on this level instead of the handle's grid
  double* d_partials = nullptr;
  CGState* d_st = nullptr;
  CGState* h_st = nullptr;  // pinned
  bool have_jacobi = false;
  int proj_kind = 0;       // ProjectedPenaltyFun: 0 none, 1 Heaviside, 2 sigmoid
  double proj_beta = 0.0;
  // assembled path
  long long* d_nbr_start = nullptr;
  int* d_rowptr = nullptr;
  int* d_col = nullptr;
  double* d_nz = nullptr;
  double* d_fasm = nullptr;
  bool pattern_ready = false, assembled = false, stiffness_dirty = true;
  std::vector<int64_t> export_map;  // CSC position (topopt_csc_pattern order) -> internal CSR position
  std::vector<int64_t> node_of_block, block_of_node;
  topopt_stats stats{};
  std::string err;
};

struct topopt_filter {
  topopt_handle* h = nullptr;
  FilterGeo fg{};
  double rmin = 0;
  double* d_w = nullptr;
  double* d_den = nullptr;
  double* d_nodal = nullptr;
  double* d_in = nullptr;
  double* d_out = nullptr;
  bool attr_set = false;
  std::vector<double> h_w;  // host copy of the cone table
};

namespace {

int fail(topopt_handle* h, int code, const std::string& msg) {
  set_error(msg);
  if (h) h->err = msg;
  return code;
}

#define CUDA_TRY(h, expr)                                                                          \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(h, TOPOPT_ERR_CUDA, std::string(#expr ": ") + cudaGetErrorString(e_));           \
  } while (0)

#define NCCL_TRY(h, expr)                                                                          \
  do {                                                                                             \
    ncclResult_t r_ = (expr);                                                                      \
    if (r_ != ncclSuccess)                                                                         \
      return fail(h, TOPOPT_ERR_NCCL, std::string(#expr ": ") + g_nccl.GetErrorString(r_));        \
  } while (0)

#define TRY(expr)                 \
  do {                            \
    int rc_ = (expr);             \
    if (rc_ != TOPOPT_OK) return rc_; \
  } while (0)

#define DISPATCH(h, CALL)                    \
  do {                                       \
    if ((h)->dim == 2 && (h)->nc == 2) {     \
      CALL(2, 2);                            \
    } else if ((h)->dim == 2) {              \
      CALL(2, 1);                            \
    } else if ((h)->nc == 3) {               \
      CALL(3, 3);                            \
    } else {                                 \
      CALL(3, 1);                            \
    }                                        \
  } while (0)

inline int grid_for(long long n, int cap = kReduceBlocks) {
  long long b = (n + kBlock - 1) / kBlock;
  if (b < 1) b = 1;
  return (int)std::min<long long>(b, cap);
}
constexpr int kWideGrid = 148 * 16;

#define LAUNCH(h, kernel, grid, ...)                              \
  do {                                                            \
    kernel<<<(grid), kBlock, 0, (h)->stream>>>(__VA_ARGS__);      \
    (h)->stats.kernel_launches += 1;                              \
  } while (0)

int check_launch(topopt_handle* h, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(h, TOPOPT_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  return TOPOPT_OK;
}

// + 64 bytes: the ring-staged kernel's 16-byte aligned bulk copies may read up to 15 bytes past a row end
template <typename T>
int dev_alloc(topopt_handle* h, T** p, size_t n) {
  const size_t bytes = std::max<size_t>(n, 1) * sizeof(T) + 64;
  CUDA_TRY(h, cudaMalloc((void**)p, bytes));
  CUDA_TRY(h, cudaMemsetAsync(*p, 0, bytes, h->stream));
  return TOPOPT_OK;
}

// The shared element matrix lives in __constant__ memory (one copy per device).  Handles that
// share a device re-upload it when ownership changes; such handles must not run concurrently.
std::mutex g_const_mutex;
uint64_t g_const_owner[64] = {0};
std::atomic<uint64_t> g_next_id{1};

int use_device(topopt_handle* h) {
  CUDA_TRY(h, cudaSetDevice(h->device));
  std::lock_guard<std::mutex> lock(g_const_mutex);
  if (g_const_owner[h->device & 63] != h->id) {
    CUDA_TRY(h, cudaDeviceSynchronize());
    CUDA_TRY(h, cudaMemcpyToSymbol(cKe, h->Ke, sizeof(double) * h->ks * h->ks, 0, cudaMemcpyHostToDevice));
    if (h->modal_ok) CUDA_TRY(h, cudaMemcpyToSymbol(cKh, h->Kh, sizeof(double) * 48, 0, cudaMemcpyHostToDevice));
    if (h->modal_cube) CUDA_TRY(h, cudaMemcpyToSymbol(cKc, h->Kc, sizeof(double) * 8, 0, cudaMemcpyHostToDevice));
    g_const_owner[h->device & 63] = h->id;
  }
  return TOPOPT_OK;
}

// ---- collectives ----------------------------------------------------------------------------
int allreduce_sums(topopt_handle* h, int n) {
  if (h->world == 1) return TOPOPT_OK;
  NCCL_TRY(h, g_nccl.AllReduce(h->d_st->sums, h->d_st->gsums, n, ncclDouble, ncclSum, h->comm, h->stream));
  return TOPOPT_OK;
}

// one node plane to each slab neighbour (ghost planes 0 and nown+1)
int exchange_halo(topopt_handle* h, double* v) {
  if (h->world == 1) return TOPOPT_OK;
  const size_t ps = (size_t)h->plane_dofs;
  NCCL_TRY(h, g_nccl.GroupStart());
  if (h->rank + 1 < h->world) {
    NCCL_TRY(h, g_nccl.Send(v + ps * h->g.nown, ps, ncclDouble, h->rank + 1, h->comm, h->stream));
    NCCL_TRY(h, g_nccl.Recv(v + ps * (h->g.nown + 1), ps, ncclDouble, h->rank + 1, h->comm, h->stream));
  }
  if (h->rank > 0) {
    NCCL_TRY(h, g_nccl.Send(v + ps, ps, ncclDouble, h->rank - 1, h->comm, h->stream));
    NCCL_TRY(h, g_nccl.Recv(v, ps, ncclDouble, h->rank - 1, h->comm, h->stream));
  }
  NCCL_TRY(h, g_nccl.GroupEnd());
  return TOPOPT_OK;
}

// ---- ABI edge: Ferrite-ordered full vectors <-> local slabs -----------------------------------
int upload_dofs(topopt_handle* h, const double* src, double* dst_local) {
  CUDA_TRY(h, cudaMemcpyAsync(h->d_full_dof, src, sizeof(double) * h->ndof, cudaMemcpyDefault, h->stream));
  h->stats.h2d_bytes += sizeof(double) * h->ndof;
#define CALL(D, C) LAUNCH(h, (k_gather_dofs<C>), grid_for(h->nloc_nodes, kWideGrid), h->g, h->d_block, h->d_full_dof, dst_local)
  DISPATCH(h, CALL);
#undef CALL
  return check_launch(h, "k_gather_dofs");
}

int download_dofs(topopt_handle* h, const double* src_local, double* dst) {
  if (h->world > 1) CUDA_TRY(h, cudaMemsetAsync(h->d_full_dof, 0, sizeof(double) * h->ndof, h->stream));
#define CALL(D, C) LAUNCH(h, (k_scatter_dofs<C>), grid_for(h->nloc_nodes, kWideGrid), h->g, h->d_block, src_local, h->d_full_dof)
  DISPATCH(h, CALL);
#undef CALL
  TRY(check_launch(h, "k_scatter_dofs"));
  if (h->world > 1)
    NCCL_TRY(h, g_nccl.AllReduce(h->d_full_dof, h->d_full_dof, h->ndof, ncclDouble, ncclSum, h->comm, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(dst, h->d_full_dof, sizeof(double) * h->ndof, cudaMemcpyDefault, h->stream));
  h->stats.d2h_bytes += sizeof(double) * h->ndof;
  return TOPOPT_OK;
}

// element vectors: Ferrite cell order == lexicographic order, so a slab is a contiguous range
int upload_elems(topopt_handle* h, const double* src_full, double* dst_local, bool count = true) {
  const Geo& g = h->g;
  const long long lo = std::max(g.p0, 0), hi = std::min<long long>(g.p0 + g.nown + 1, g.NLg);  // global layers
  if (hi > lo) {
    CUDA_TRY(h, cudaMemcpyAsync(dst_local + (lo - g.p0) * g.SE, src_full + lo * g.SE,
                                sizeof(double) * (hi - lo) * g.SE, cudaMemcpyDefault, h->stream));
    if (count) h->stats.h2d_bytes += sizeof(double) * (hi - lo) * g.SE;
  }
  return TOPOPT_OK;
}

// owned layers of a local element vector -> full vector at dst_dev_or_host
int gather_elems_device(topopt_handle* h, const double* src_local, double** full_out) {
  if (h->world == 1) {
    *full_out = const_cast<double*>(src_local) + h->eoff;
    return TOPOPT_OK;
  }
  const Geo& g = h->g;
  CUDA_TRY(h, cudaMemsetAsync(h->d_full_el, 0, sizeof(double) * h->nel, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_full_el + (long long)(g.p0 + 1) * g.SE, src_local + h->eoff,
                              sizeof(double) * h->nown_el, cudaMemcpyDeviceToDevice, h->stream));
  NCCL_TRY(h, g_nccl.AllReduce(h->d_full_el, h->d_full_el, h->nel, ncclDouble, ncclSum, h->comm, h->stream));
  *full_out = h->d_full_el;
  return TOPOPT_OK;
}

int download_elems(topopt_handle* h, const double* src_local, double* dst) {
  double* full = nullptr;
  TRY(gather_elems_device(h, src_local, &full));
  CUDA_TRY(h, cudaMemcpyAsync(dst, full, sizeof(double) * h->nel, cudaMemcpyDefault, h->stream));
  h->stats.d2h_bytes += sizeof(double) * h->nel;
  return TOPOPT_OK;
}

int sync(topopt_handle* h) {
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return TOPOPT_OK;
}

// ---- operator application -----------------------------------------------------------------
constexpr int kMaxPartialBlocks = 16384;

inline const Geo& cur_geo(const topopt_handle* h) { return h->lvl ? h->lvl->g : h->g; }
inline const double* cur_E(const topopt_handle* h) { return h->lvl ? h->lvl->E : h->d_E; }
inline const unsigned char* cur_fixed(const topopt_handle* h) { return h->lvl ? h->lvl->fixed : h->d_fixed; }

// cudaFuncSetAttribute is per device: one bit per device ordinal for every kernel instantiation
template <typename K>
int ensure_dyn_smem(topopt_handle* h, K kernel, size_t smem, std::atomic<unsigned long long>& mask) {
  const unsigned long long bit = 1ull << (h->device & 63);
  if (mask.load(std::memory_order_acquire) & bit) return TOPOPT_OK;
  CUDA_TRY(h, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mask.fetch_or(bit, std::memory_order_release);
  return TOPOPT_OK;
}

template <int TY, bool DOT, bool FUSEP, bool PEER, bool NSYNC>
int launch_hex8_modal(topopt_handle* h, const double* x, double* y, int fin, const double* r, double* pnew) {
  const Geo& g = cur_geo(h);
  const int tilesX = (g.NX + 29) / 30, tilesY = (g.NY + TY - 3) / (TY - 2);
  // persistent grid: resident CTAs per SM (register-limited) x 148 SMs, capped by the work
  const int per_sm = TY <= 4 ? 4 : (TY <= 8 ? 2 : 1);
  int grid = 148 * per_sm * std::max(1, h->kxu_waves);
  const long long units = (long long)tilesX * tilesY * g.nown;
  if (grid > units) grid = (int)units;
  if (grid > kMaxPartialBlocks) grid = kMaxPartialBlocks;
  const size_t smem = sizeof(double) * 2 * 12 * TY * 32;
  static std::atomic<unsigned long long> attr_mask{0};  // per template instantiation, one bit per device
  TRY(ensure_dyn_smem(h, k_apply_hex8_modal<TY, DOT, FUSEP, PEER, NSYNC>, smem, attr_mask));
  const double* xlo = nullptr;
  const double* xhi = nullptr;
  if (PEER) {  // neighbours' boundary planes of the same buffer (d_p), read over NVLink
    if (h->peer_p_lo) xlo = h->peer_p_lo + (size_t)h->plane_dofs * h->nown_lower;
    if (h->peer_p_hi) xhi = h->peer_p_hi + (size_t)h->plane_dofs;
  }
  k_apply_hex8_modal<TY, DOT, FUSEP, PEER, NSYNC><<<grid, 32 * TY, smem, h->stream>>>(g, x, y, cur_E(h), cur_fixed(h), h->fixed_diag, tilesX,
                                                                              tilesY, h->d_partials, h->d_st, fin, r, pnew, xlo, xhi);
  h->stats.kernel_launches += 1;
  return check_launch(h, "k_apply_hex8_modal");
}

template <int TYT, bool DOT, bool PEER>
int launch_hex8_modal2(topopt_handle* h, const double* x, double* y, int fin) {
  const Geo& g = cur_geo(h);
  constexpr int ROWS = 2 * TYT - 2;
  const int tilesX = (g.NX + 29) / 30, tilesY = (g.NY + ROWS - 1) / ROWS;
  int grid = 148 * std::max(1, h->kxu_waves);
  const long long units = (long long)tilesX * tilesY * g.nown;
  if (grid > units) grid = (int)units;
  const size_t smem = sizeof(double) * 2 * 12 * TYT * 32;
  static std::atomic<unsigned long long> attr_mask{0};
  TRY(ensure_dyn_smem(h, k_apply_hex8_modal2<TYT, DOT, PEER>, smem, attr_mask));
  const double* xlo = nullptr;
  const double* xhi = nullptr;
  if (PEER) {
    if (h->peer_p_lo) xlo = h->peer_p_lo + (size_t)h->plane_dofs * h->nown_lower;
    if (h->peer_p_hi) xhi = h->peer_p_hi + (size_t)h->plane_dofs;
  }
  k_apply_hex8_modal2<TYT, DOT, PEER><<<grid, 32 * TYT, smem, h->stream>>>(g, x, y, cur_E(h), cur_fixed(h), h->fixed_diag, tilesX, tilesY,
                                                                          h->d_partials, h->d_st, fin, xlo, xhi);
  h->stats.kernel_launches += 1;
  return check_launch(h, "k_apply_hex8_modal2");
}

// Ring-staged kernel (kxu_hex8_ring.cuh).  x must be zero on prescribed dofs.
template <int TYT, int DOT, bool PEER>
int launch_hex8_ring_t(topopt_handle* h, const double* x, double* y, int fin) {
  constexpr int NST = 4;
  const Geo& g = cur_geo(h);
  constexpr int OWNR = 2 * TYT - 1;
  const int tilesX = (g.NX + 30) / 31, tilesY = (g.NY + OWNR - 1) / OWNR;
  int grid = 148 * (TYT <= 5 ? 2 : 1) * std::max(1, h->kxu_waves);
  if (h->kxu_grid > 0) grid = h->kxu_grid;
  const long long units = (long long)tilesX * tilesY * g.nown;
  if (grid > units) grid = (int)units;
  const size_t smem = hex8_ring_smem(TYT, NST);  // per-warp rings + y exchange
  const double* xlo = nullptr;
  const double* xhi = nullptr;
  if (PEER) {
    if (h->peer_p_lo) xlo = h->peer_p_lo + (size_t)h->plane_dofs * h->nown_lower;
    if (h->peer_p_hi) xhi = h->peer_p_hi + (size_t)h->plane_dofs;
  }
  if (h->modal_cube && h->kxu_cube) {
    static std::atomic<unsigned long long> attr_mask{0};
    TRY(ensure_dyn_smem(h, k_apply_hex8_ring<TYT, NST, DOT, PEER, true>, smem, attr_mask));
    k_apply_hex8_ring<TYT, NST, DOT, PEER, true><<<grid, 32 * (TYT + 1), smem, h->stream>>>(g, x, y, cur_E(h), cur_fixed(h), h->fixed_diag, tilesX, tilesY,
                                                                                           h->d_partials, h->d_st, fin, xlo, xhi);
  } else {
    static std::atomic<unsigned long long> attr_mask{0};
    TRY(ensure_dyn_smem(h, k_apply_hex8_ring<TYT, NST, DOT, PEER, false>, smem, attr_mask));
    k_apply_hex8_ring<TYT, NST, DOT, PEER, false><<<grid, 32 * (TYT + 1), smem, h->stream>>>(g, x, y, cur_E(h), cur_fixed(h), h->fixed_diag, tilesX, tilesY,
                                                                                            h->d_partials, h->d_st, fin, xlo, xhi);
  }
  h->stats.kernel_launches += 1;
  return check_launch(h, "k_apply_hex8_ring");
}

// compute warps (thread rows) per CTA of the ring kernel: fewest node rows processed for this grid (a tile of
// 2*TYT rows owns 2*TYT-1).  With the producer warp, 11 compute warps is the most that fits 156 registers.
inline int ring_rows(const topopt_handle* h) {
  if (h->kxu_ring > 1) return h->kxu_ring;
  long long best = -1;
  int tyt = 10;
  const int cand[3] = {11, 10, 8};
  for (int k = 0; k < 3; ++k) {
    const int own = 2 * cand[k] - 1;
    const long long rows = (long long)((cur_geo(h).NY + own - 1) / own) * 2 * cand[k];
    if (best < 0 || rows < best) {
      best = rows;
      tyt = cand[k];
    }
  }
  return tyt;
}

inline bool use_ring(const topopt_handle* h) {
  return h->kxu_ring != 0 && h->dim == 3 && h->nc == 3 && h->modal_ok && cur_geo(h).nown >= h->kxu_ring_min;
}

template <int DOT, bool PEER>
int launch_hex8_ring(topopt_handle* h, const double* x, double* y, int fin) {
  switch (ring_rows(h)) {
    case 5: return launch_hex8_ring_t<5, DOT, PEER>(h, x, y, fin);
    case 8: return launch_hex8_ring_t<8, DOT, PEER>(h, x, y, fin);
    case 11: return launch_hex8_ring_t<11, DOT, PEER>(h, x, y, fin);
    default: return launch_hex8_ring_t<10, DOT, PEER>(h, x, y, fin);
  }
}

// ---- one CG iteration per launch (kxu_hex8_cgfused.cuh) ----
// CTAs for a z-marching kernel whose units (tile column x plane) are split evenly over the grid: k segments per tile
// column, so that no CTA straddles two columns (a new column costs two halo planes and a pipeline refill), with
// tiles * k close below a multiple of the SM count (full waves; the block scheduler evens out slow SMs, which a single
// static wave cannot).  Measured at config 4 (63 columns x 129 planes): 148 CTAs 246 us, 444 CTAs 214 us, 441 = 63 x 7 207 us.
inline int aligned_grid(int tiles, int planes) {
  double best = -1.0;
  int grid = tiles;
  for (int k = 1; k <= 16 && (k == 1 || 4 * k <= planes); ++k) {
    const long long ctas = (long long)tiles * k;
    const double fill = (double)ctas / (double)(((ctas + 147) / 148) * 148);
    const double seg = (double)planes / k;
    const double eff = fill * seg / (seg + 4.0) * (ctas >= 2 * 148 ? 1.0 : 0.93);  // one or two waves: no dynamic balancing
    if (eff > best) {
      best = eff;
      grid = (int)ctas;
    }
  }
  return grid;
}

// thread rows per CTA: the variant that wastes the fewest node rows among those whose shared memory fits
inline int fused_rows(const topopt_handle* h) {
  if (h->cg_fused > 1) return h->cg_fused;
  long long best = -1;
  int tyt = 10;
  const int cand[2] = {10, 8};
  for (int k = 0; k < 2; ++k) {
    const int own = 2 * cand[k] - 1;
    const long long rows = (long long)((h->g.NY + own - 1) / own) * 2 * cand[k];
    if (best < 0 || rows < best) {
      best = rows;
      tyt = cand[k];
    }
  }
  return tyt;
}

template <int TYT, bool PEER>
int launch_cg_fused_t(topopt_handle* h, int fin) {
  constexpr int NST = 4;
  const Geo& g = h->g;
  constexpr int OWNR = 2 * TYT - 1;
  const int tilesX = (g.NX + 30) / 31, tilesY = (g.NY + OWNR - 1) / OWNR;
  const long long units = (long long)tilesX * tilesY * g.nown;
  int grid = h->cg_fused_grid > 0 ? h->cg_fused_grid : aligned_grid(tilesX * tilesY, g.nown);
  if (units < grid) grid = (int)units;
  const size_t smem = hex8_fused_smem(TYT, NST);
  CGFusedVecs v;
  v.p[0] = h->d_p;
  v.p[1] = h->d_p2;
  v.r[0] = h->d_r;
  v.r[1] = h->d_r2;
  v.ap[0] = h->d_Ap;
  v.ap[1] = h->d_Ap2;
  v.x = h->d_u;
  for (int k = 0; k < 3; ++k)
    for (int par = 0; par < 2; ++par) {
      double* lo = PEER ? const_cast<double*>(h->peer_lo[2 * k + par]) : nullptr;
      double* hi = PEER ? const_cast<double*>(h->peer_hi[2 * k + par]) : nullptr;
      v.wlo[k][par] = lo ? lo + (size_t)h->plane_dofs * (h->nown_lower + 1) : nullptr;  // the lower neighbour's top ghost plane
      v.whi[k][par] = hi;                                                               // the upper neighbour's bottom ghost plane
    }
  // programmatic dependent launch: this kernel's CTAs may run their prologue while the previous iteration's kernel finishes
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(32 * (TYT + kFusedProducers));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = h->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = h->cg_fused_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const double* dE = h->d_E;
  const unsigned char* dfix = h->d_fixed;
  cudaError_t le;
  if (h->modal_cube && h->kxu_cube) {
    static std::atomic<unsigned long long> attr_mask{0};
    TRY(ensure_dyn_smem(h, k_cg_fused_hex8<TYT, NST, true, PEER>, smem, attr_mask));
    le = cudaLaunchKernelEx(&cfg, k_cg_fused_hex8<TYT, NST, true, PEER>, g, v, dE, dfix, h->fixed_diag, tilesX, tilesY, h->d_partials, h->d_st, fin);
  } else {
    static std::atomic<unsigned long long> attr_mask{0};
    TRY(ensure_dyn_smem(h, k_cg_fused_hex8<TYT, NST, false, PEER>, smem, attr_mask));
    le = cudaLaunchKernelEx(&cfg, k_cg_fused_hex8<TYT, NST, false, PEER>, g, v, dE, dfix, h->fixed_diag, tilesX, tilesY, h->d_partials, h->d_st, fin);
  }
  if (le != cudaSuccess) return fail(h, TOPOPT_ERR_CUDA, std::string("k_cg_fused_hex8 launch: ") + cudaGetErrorString(le));
  h->stats.kernel_launches += 1;
  return check_launch(h, "k_cg_fused_hex8");
}

// ---- tensor maps of the six CG vectors (and the slab neighbours') for kxu_hex8_cgtma.cuh ----
typedef CUresult (*PFN_tmap_encode_tiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int ensure_tmaps(topopt_handle* h, int tyt) {
  if (h->d_tmaps && h->tmaps_rows == tyt) return TOPOPT_OK;
  static PFN_tmap_encode_tiled encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || fn == nullptr ||
        qres != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return fail(h, TOPOPT_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    }
    encode = reinterpret_cast<PFN_tmap_encode_tiled>(fn);
  }
  const Geo& g = h->g;
  const int shift = (g.NX & 1) ? 1 : 0;
  const long long row_doubles = (long long)g.NX * 3;
  std::vector<CUtensorMap> maps(36);
  std::memset(maps.data(), 0, sizeof(CUtensorMap) * maps.size());
  // even / odd rows of the [R = plane * NY + row][3 NX] view of one vector with `planes` node planes
  auto build = [&](const double* base, long long planes, CUtensorMap* out) -> int {
    const long long Rtot = planes * g.NY;
    for (int q = 0; q < 2; ++q) {
      const long long rows = q == 0 ? (Rtot + 1) / 2 : Rtot / 2;
      const double* b = q == 0 ? base : base + row_doubles - shift;
      const cuuint64_t gdim[2] = {(cuuint64_t)(row_doubles + (q == 1 ? shift : 0)), (cuuint64_t)std::max<long long>(rows, 1)};
      const cuuint64_t gstride[1] = {(cuuint64_t)(2 * row_doubles * 8)};
      const cuuint32_t box[2] = {100u, (cuuint32_t)(tyt + 1)};
      const cuuint32_t estr[2] = {1u, 1u};
      const CUresult r = encode(&out[q], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(b), gdim, gstride, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(h, TOPOPT_ERR_CUDA, "cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
    }
    return TOPOPT_OK;
  };
  const double* own[6] = {h->d_p, h->d_p2, h->d_r, h->d_r2, h->d_Ap, h->d_Ap2};
  for (int b = 0; b < 6; ++b) {
    TRY(build(own[b], g.nown + 2, &maps[2 * b]));
    // neighbours: only their boundary plane is read; their buffers hold at least nown_lower + 2 / 3 planes
    if (h->peer_lo[b]) TRY(build(h->peer_lo[b], h->nown_lower + 2, &maps[12 + 2 * b]));
    if (h->peer_hi[b]) TRY(build(h->peer_hi[b], 3, &maps[24 + 2 * b]));
  }
  if (!h->d_tmaps) CUDA_TRY(h, cudaMalloc((void**)&h->d_tmaps, sizeof(CUtensorMap) * maps.size()));
  CUDA_TRY(h, cudaMemcpy(h->d_tmaps, maps.data(), sizeof(CUtensorMap) * maps.size(), cudaMemcpyHostToDevice));
  h->tmaps_rows = tyt;
  return TOPOPT_OK;
}

#ifndef TOPOPT_TMA_NST
#define TOPOPT_TMA_NST 4
#endif
template <int TYT, bool PEER>
int launch_cg_tma_t(topopt_handle* h, int fin) {
  constexpr int NST = TOPOPT_TMA_NST;
  const Geo& g = h->g;
  constexpr int OWNR = 2 * TYT - 1;
  const int tilesX = (g.NX + 30) / 31, tilesY = (g.NY + OWNR - 1) / OWNR;
  const long long units = (long long)tilesX * tilesY * g.nown;
  int grid = h->cg_fused_grid > 0 ? h->cg_fused_grid : aligned_grid(tilesX * tilesY, g.nown);
  if (units < grid) grid = (int)units;
  const size_t smem = hex8_cgtma_smem(TYT, NST);
  TRY(ensure_tmaps(h, TYT));
  CGFusedVecs v;
  std::memset(&v, 0, sizeof(v));
  v.p[0] = h->d_p;
  v.p[1] = h->d_p2;
  v.r[0] = h->d_r;
  v.r[1] = h->d_r2;
  v.ap[0] = h->d_Ap;
  v.ap[1] = h->d_Ap2;
  v.x = h->d_u;
  CGTmaArgs ta;
  ta.maps = h->d_tmaps;
  ta.shift = (g.NX & 1) ? 1 : 0;
  ta.lo_plane = (PEER && h->peer_lo[0]) ? h->nown_lower : -1;
  ta.hi_plane = (PEER && h->peer_hi[0]) ? 1 : -1;
  if (h->modal_cube && h->kxu_cube) {
    static std::atomic<unsigned long long> attr_mask{0};
    TRY(ensure_dyn_smem(h, k_cg_tma_hex8<TYT, NST, true, PEER>, smem, attr_mask));
    k_cg_tma_hex8<TYT, NST, true, PEER><<<grid, 32 * (TYT + 1), smem, h->stream>>>(g, v, ta, h->d_E, h->d_fixed, h->fixed_diag, tilesX,
                                                                                 tilesY, h->d_partials, h->d_st, fin);
  } else {
    static std::atomic<unsigned long long> attr_mask{0};
    TRY(ensure_dyn_smem(h, k_cg_tma_hex8<TYT, NST, false, PEER>, smem, attr_mask));
    k_cg_tma_hex8<TYT, NST, false, PEER><<<grid, 32 * (TYT + 1), smem, h->stream>>>(g, v, ta, h->d_E, h->d_fixed, h->fixed_diag, tilesX,
                                                                                  tilesY, h->d_partials, h->d_st, fin);
  }
  h->stats.kernel_launches += 1;
  return check_launch(h, "k_cg_tma_hex8");
}

int launch_cg_fused(topopt_handle* h, bool peer, int fin) {
  const int rows = fused_rows(h);
  const bool force_peer_code = getenv("TOPOPT_CG_FUSED_FORCE_PEER_CODE") != nullptr;  // diagnostic: PEER instantiation on one GPU
  if (force_peer_code && !h->cg_fused_tma) peer = true;
  // the tensor-map variant is single-GPU (its multi-GPU instantiation pulls ghost planes and was never run across ranks)
  if (h->cg_fused_tma && !peer) return rows == 8 ? launch_cg_tma_t<8, false>(h, fin) : launch_cg_tma_t<10, false>(h, fin);
  if (peer) return rows == 8 ? launch_cg_fused_t<8, true>(h, fin) : launch_cg_fused_t<10, true>(h, fin);
  return rows == 8 ? launch_cg_fused_t<8, false>(h, fin) : launch_cg_fused_t<10, false>(h, fin);
}

template <bool DOT, bool FUSEP, bool PEER = false>
int launch_hex8(topopt_handle* h, const double* x, double* y, int fin, const double* r, double* pnew) {
  // Two node rows per thread win on thick slabs (single GPU: +6 %); on thin slabs (>= 4 ranks at
  // config 4) the persistent CTAs' segments get short and the one-row kernel measured 8-15 % faster.
  if (h->kxu_2row && !FUSEP && (h->kxu_2row != 1 || cur_geo(h).nown >= 96)) {
    int tyt = h->kxu_2row;
    if (tyt == 1) {  // auto: thread rows per CTA that waste the fewest node rows for this grid
      double best = -1.0;
      const int cand[3] = {12, 10, 8};
      const double speed[3] = {1.0, 0.97, 0.93};  // measured relative efficiency of the variants
      for (int k = 0; k < 3; ++k) {
        const int rows = 2 * cand[k] - 2;
        const int tiles = (cur_geo(h).NY + rows - 1) / rows;
        const double eff = speed[k] * (double)cur_geo(h).NY / ((double)tiles * (rows + 2));
        if (eff > best) {
          best = eff;
          tyt = cand[k];
        }
      }
    }
    if (tyt == 10) return launch_hex8_modal2<10, DOT, PEER>(h, x, y, fin);
    if (tyt == 14) return launch_hex8_modal2<14, DOT, PEER>(h, x, y, fin);
    if (tyt == 8) return launch_hex8_modal2<8, DOT, PEER>(h, x, y, fin);
    return launch_hex8_modal2<12, DOT, PEER>(h, x, y, fin);
  }
  if (h->kxu_nsync) {
    if (h->kxu_ty == 8) return launch_hex8_modal<8, DOT, FUSEP, PEER, true>(h, x, y, fin, r, pnew);
    if (h->kxu_ty == 16) return launch_hex8_modal<16, DOT, FUSEP, PEER, true>(h, x, y, fin, r, pnew);
    if (h->kxu_ty == 14) return launch_hex8_modal<14, DOT, FUSEP, PEER, true>(h, x, y, fin, r, pnew);
    return launch_hex8_modal<12, DOT, FUSEP, PEER, true>(h, x, y, fin, r, pnew);
  }
  if (h->kxu_ty == 8) return launch_hex8_modal<8, DOT, FUSEP, PEER, false>(h, x, y, fin, r, pnew);
  if (h->kxu_ty == 16) return launch_hex8_modal<16, DOT, FUSEP, PEER, false>(h, x, y, fin, r, pnew);
  return launch_hex8_modal<12, DOT, FUSEP, PEER, false>(h, x, y, fin, r, pnew);
}

template <bool DOT>
int launch_apply(topopt_handle* h, const double* x, double* y, int fin) {
  if (h->dim == 3 && h->nc == 3 && h->modal_ok) return launch_hex8<DOT, false>(h, x, y, fin, nullptr, nullptr);
  const int grid = grid_for((long long)cur_geo(h).S * cur_geo(h).nown, DOT ? kReduceBlocks : kWideGrid);
#define CALL(D, C) \
  LAUNCH(h, (k_apply<D, C, DOT>), grid, cur_geo(h), x, y, cur_E(h), cur_fixed(h), h->fixed_diag, h->d_partials, h->d_st, fin)
  DISPATCH(h, CALL);
#undef CALL
  return check_launch(h, "k_apply");
}

template <bool DOT>
int launch_spmv(topopt_handle* h, const double* x, double* y, int fin) {
  const long long nrows = h->ndof;
  const int lanes = (h->dim == 3 && h->nc == 3) ? 32 : 8;
  const int grid = grid_for(nrows * lanes, DOT ? kReduceBlocks : kWideGrid);
  if (lanes == 32)
    LAUNCH(h, (k_spmv<32, DOT>), grid, nrows, h->d_rowptr, h->d_col, h->d_nz, x + h->off, y + h->off, h->d_partials,
           h->d_st, fin);
  else
    LAUNCH(h, (k_spmv<8, DOT>), grid, nrows, h->d_rowptr, h->d_col, h->d_nz, x + h->off, y + h->off, h->d_partials,
           h->d_st, fin);
  return check_launch(h, "k_spmv");
}

int build_pattern(topopt_handle* h);
int do_assemble(topopt_handle* h);

// K.u inside the CG loop: the direction vector is zero on prescribed dofs, so the ring-staged kernel
// applies; DOT = 1 reduces p.Ap, DOT = 2 also Ap.Ap (single-pass recurrence).
template <int DOT>
int launch_cg_apply(topopt_handle* h, bool peer_halo, int fin) {
  const bool hex8 = h->dim == 3 && h->nc == 3 && h->modal_ok;
  if (hex8 && use_ring(h)) {
    if (peer_halo) return launch_hex8_ring<DOT, true>(h, h->d_p, h->d_Ap, fin);
    TRY(exchange_halo(h, h->d_p));
    return launch_hex8_ring<DOT, false>(h, h->d_p, h->d_Ap, fin);
  }
  if (DOT == 2) return fail(h, TOPOPT_ERR_INVALID, "single-pass CG needs the ring-staged hex8 kernel");
  if (peer_halo) return launch_hex8<true, false, true>(h, h->d_p, h->d_Ap, fin, nullptr, nullptr);
  TRY(exchange_halo(h, h->d_p));
  return launch_apply<true>(h, h->d_p, h->d_Ap, fin);
}

// ---- persistent CG for small grids (kernels.cuh: k_cg_persistent) ----
template <int DIM, int NC>
int launch_cg_persistent_t(topopt_handle* h, int niter) {
  static std::atomic<int> blocks_per_sm{0};
  int nb = blocks_per_sm.load();
  if (nb == 0) {
    CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_cg_persistent<DIM, NC>, kBlock, 0));
    blocks_per_sm.store(nb);
  }
  if (nb < 1) return TOPOPT_ERR_CUDA;
  int sms = 0;
  CUDA_TRY(h, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
  const long long nnodes = (long long)h->g.S * h->g.nown;
  const int grid = (int)std::max<long long>(1, std::min<long long>((long long)nb * sms, (nnodes + kBlock - 1) / kBlock));
  if (2 * grid > kMaxPartialBlocks) return TOPOPT_ERR_CUDA;
  if (!h->d_barrier) CUDA_TRY(h, cudaMalloc((void**)&h->d_barrier, sizeof(unsigned int)));
  CUDA_TRY(h, cudaMemsetAsync(h->d_barrier, 0, sizeof(unsigned int), h->stream));
  Geo g = h->g;
  double *x = h->d_u, *r = h->d_r, *p = h->d_p, *Ap = h->d_Ap, *partials = h->d_partials;
  const double* E = h->d_E;
  const unsigned char* fixed = h->d_fixed;
  double fd = h->fixed_diag;
  CGState* st = h->d_st;
  unsigned int* bar = h->d_barrier;
  void* args[] = {&g, &x, &r, &p, &Ap, &E, &fixed, &fd, &partials, &st, &bar, &niter};
  const cudaError_t e = cudaLaunchCooperativeKernel((const void*)k_cg_persistent<DIM, NC>, dim3(grid), dim3(kBlock), args, 0, h->stream);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return TOPOPT_ERR_CUDA;  // the caller falls back to one launch per kernel
  }
  h->stats.kernel_launches += 1;
  return TOPOPT_OK;
}

int launch_cg_persistent(topopt_handle* h, int niter) {
  if (h->dim == 2 && h->nc == 2) return launch_cg_persistent_t<2, 2>(h, niter);
  if (h->dim == 2) return launch_cg_persistent_t<2, 1>(h, niter);
  if (h->nc == 3) return launch_cg_persistent_t<3, 3>(h, niter);
  return launch_cg_persistent_t<3, 1>(h, niter);
}

#include "mg_solve.inl"

// IterativeSolvers cg!: see cg_finalize() for the scalar recurrences.
// b (local layout) must already be zero on prescribed dofs.
int cg_solve(topopt_handle* h, const double* b, const topopt_cg_opts* o, topopt_cg_result* res, bool ignore_convergence = false,
             int fixed_iters = 0) {
  if (o->precond == TOPOPT_PRECOND_MULTIGRID) return mg_pcg_solve(h, b, o, res);
  const bool assembled = o->op == TOPOPT_OP_ASSEMBLED;
  if (assembled) {
    if (h->world > 1) return fail(h, TOPOPT_ERR_INVALID, "the assembled operator is single-GPU only");
    if (!h->assembled || h->stiffness_dirty) TRY(do_assemble(h));
  }
  const bool pre = o->precond == TOPOPT_PRECOND_JACOBI;
  if (pre && o->refresh_precond) {
#define CALL(D, C) LAUNCH(h, (k_diag<D, C>), grid_for((long long)h->g.S * h->g.nown, kWideGrid), h->g, h->d_D, h->d_E, h->d_fixed, h->fixed_diag)
    DISPATCH(h, CALL);
#undef CALL
    TRY(check_launch(h, "k_diag"));
    h->have_jacobi = true;
  }
  if (pre && !h->have_jacobi) return fail(h, TOPOPT_ERR_INVALID, "Jacobi preconditioner requested but topopt_set_jacobi was not called");
  const bool energy = o->criteria == TOPOPT_CRITERIA_ENERGY;
  const bool peer = h->world > 1 && h->peer_ready;                       // in-kernel allreduce over peer memory
  const bool hex8 = !assembled && h->dim == 3 && h->nc == 3 && h->modal_ok;
  const bool peer_halo = peer && hex8;  // + direct halo reads
  // single-pass recurrence: identity preconditioner, default criteria, ring-staged K.u, device-side scalars
  int want_variant = h->cg_variant_env >= 0 ? h->cg_variant_env : o->variant;
  const bool single = want_variant == TOPOPT_CG_SINGLE_PASS && !pre && !energy && hex8 && use_ring(h) && (h->world == 1 || peer);
  // ... and on one GPU the whole iteration is one kernel (vector updates applied while the planes are staged)
  const bool fused = single && h->cg_fused != 0 && (h->world == 1 || (peer_halo && h->peer_fused_ready && !h->cg_fused_single_only));
  // small grids, the reference's recurrence: batches of iterations in one cooperative kernel (launch latency, not bytes,
  // is what an iteration costs there).  2-D grids and 3-D grids of at most ~200 k nodes; the large hex8 grids have
  // the ring-staged / one-kernel paths above.
  bool persist = h->cg_persist != 0 && h->world == 1 && !assembled && !pre && !energy && !single &&
                 (h->dim == 2 || (long long)h->g.S * h->g.nown <= 200000);
  CGState& s = *h->h_st;
  std::memset(&s, 0, sizeof(CGState));
  s.abstol = ignore_convergence ? -1.0 : o->abstol;
  s.reltol = ignore_convergence ? 0.0 : o->reltol;
  s.maxiter = ignore_convergence ? fixed_iters : o->maxiter;
  s.criteria = ignore_convergence ? 0 : o->criteria;
  s.precond = pre ? 1 : 0;
  s.world = h->world;
  s.variant = single ? 1 : 0;
  s.peer = peer ? h->d_peercomm : nullptr;
  CUDA_TRY(h, cudaMemcpyAsync(h->d_st, &s, sizeof(CGState), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
  const double* D = pre ? h->d_D : nullptr;
  // vector kernels: 4 elements per thread and trip; small problems get a proportionally small grid
  const int vgrid = (int)std::min<long long>(kReduceBlocks, std::max<long long>(1, (h->nown_dofs + 4 * kBlock - 1) / (4 * kBlock)));
  const double* b0 = b;
  if (o->warm_start) {
    // r0 = b - K u_prev (u_prev is zero on prescribed dofs: it came out of a CG solve); the ghost planes of
    // d_u were refreshed at the end of that solve
    if (assembled) {
      TRY(launch_spmv<false>(h, h->d_u, h->d_Ap, FIN_NONE));
    } else {
      TRY(launch_apply<false>(h, h->d_u, h->d_Ap, FIN_NONE));
    }
    LAUNCH(h, k_axpby, vgrid, h->off, h->nown_dofs, 1.0, b, -1.0, h->d_Ap, h->d_tmp);
    b0 = h->d_tmp;
  }
  LAUNCH(h, k_cg_init, vgrid, h->off, h->nown_dofs, b0, o->warm_start ? (double*)nullptr : h->d_u, h->d_r, h->d_p, D, h->d_partials,
         h->d_st, single ? 1 : 0, (single && peer_halo) ? 1 : 0);
  TRY(check_launch(h, "k_cg_init"));
  if (h->world > 1 && !peer) {
    TRY(allreduce_sums(h, 2));
    k_finalize<<<1, 1, 0, h->stream>>>(h->d_st, FIN_INIT);
    h->stats.kernel_launches += 1;
  }
  // Ap_0 = K p_0, alpha_0, beta_0: every later iteration is one launch (multi-GPU: and tell the neighbours Ap_0 is final)
  if (fused) {
    TRY(launch_cg_apply<2>(h, peer_halo, FIN_PAP2 | (peer_halo ? 0x100 : 0)));
    if (peer_halo) {
      // the one-kernel iteration stages its ghost planes from local memory (the neighbours push them from then on)
      TRY(exchange_halo(h, h->d_p));
      TRY(exchange_halo(h, h->d_r));
      TRY(exchange_halo(h, h->d_Ap));
    }
  }
  int batch = o->check_every > 0 ? o->check_every : (h->ndof > 2000000 ? 25 : 50);
  int issued = 0;
  const int maxiter = s.maxiter;
  while (true) {
    CUDA_TRY(h, cudaMemcpyAsync(h->h_st, h->d_st, sizeof(CGState), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (h->h_st->done || issued >= maxiter) break;
    const int n = std::min(batch, maxiter - issued);
    const bool fusep = !single && !assembled && !pre && h->world == 1 && hex8 && !h->no_fuse && !use_ring(h);
    // TOPOPT_TRACE=1: per-phase device times of the first batch (diagnostic, perturbs the run)
    static const bool trace = getenv("TOPOPT_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    auto mark = [&]() {
      if (!trace || issued > 0) return;
      cudaEvent_t e;
      cudaEventCreate(&e);
      cudaEventRecord(e, h->stream);
      tev.push_back(e);
    };
    int launches_per_iter = 0;
    auto one_iteration = [&]() -> int {
      launches_per_iter = 0;
      mark();
      if (fused) {  // r, p, x of the previous step while staging ; K.u ; p.Ap, Ap.Ap, r.r -> alpha, beta, convergence
        TRY(launch_cg_fused(h, peer_halo, FIN_FUSED));
        launches_per_iter = 1;
        return TOPOPT_OK;
      }
      if (single) {  // K.u (+ p.Ap, Ap.Ap -> alpha, beta) ; x, r, p in one pass (+ r.r)
        TRY(launch_cg_apply<2>(h, peer_halo, FIN_PAP2));
        mark();
        LAUNCH(h, k_update_xrp, vgrid, h->off, h->nown_dofs, h->d_u, h->d_r, h->d_p, h->d_Ap, h->d_partials, h->d_st,
               peer_halo ? (long long)h->plane_dofs : 0ll);
        launches_per_iter = 2;
        return TOPOPT_OK;
      }
      if (fusep) {  // p_new = r + beta p_old formed inside the K.u kernel
        TRY((launch_hex8<true, true>(h, h->d_p, h->d_Ap, FIN_PAP, h->d_r, h->d_p2)));
        std::swap(h->d_p, h->d_p2);
      } else {
        LAUNCH(h, k_update_p, vgrid, h->off, h->nown_dofs, h->d_r, h->d_p, D, h->d_st, peer_halo ? 1 : 0);
        if (assembled) {
          TRY(launch_spmv<true>(h, h->d_p, h->d_Ap, FIN_PAP));
        } else {
          mark();
          if (hex8) {
            TRY(launch_cg_apply<1>(h, peer_halo, FIN_PAP));
          } else {
            TRY(exchange_halo(h, h->d_p));
            TRY(launch_apply<true>(h, h->d_p, h->d_Ap, FIN_PAP));
          }
        }
      }
      if (h->world > 1 && !peer) {
        TRY(allreduce_sums(h, 1));
        k_finalize<<<1, 1, 0, h->stream>>>(h->d_st, FIN_PAP);
        h->stats.kernel_launches += 1;
      }
      mark();
      if (energy)
        LAUNCH(h, (k_update_xr<true>), vgrid, h->off, h->nown_dofs, h->d_u, h->d_r, h->d_p, h->d_Ap, D, b,
               h->d_partials, h->d_st);
      else
        LAUNCH(h, (k_update_xr<false>), vgrid, h->off, h->nown_dofs, h->d_u, h->d_r, h->d_p, h->d_Ap, D, b,
               h->d_partials, h->d_st);
      if (h->world > 1 && !peer) {
        TRY(allreduce_sums(h, 4));
        k_finalize<<<1, 1, 0, h->stream>>>(h->d_st, FIN_RR);
        h->stats.kernel_launches += 1;
      }
      launches_per_iter = 3;
      return TOPOPT_OK;
    };
    // multi-GPU single-pass recurrence: the last iteration's r.r is posted but not collected yet
    auto flush_batch = [&]() {
      if (single && peer && !fused) {
        k_cg_flush<<<1, 32, 0, h->stream>>>(h->d_st);
        h->stats.kernel_launches += 1;
      }
    };
    // Replay a captured graph of n iterations when nothing in the batch depends on host state:
    // single GPU or the peer-memory path (no NCCL calls inside), no fused pointer swap, no tracing.
    if (persist && !trace) {
      const int rc = launch_cg_persistent(h, n);
      if (rc == TOPOPT_OK) {
        TRY(check_launch(h, "k_cg_persistent"));
        issued += n;
        continue;
      }
      persist = false;  // cooperative launch not available here: one launch per kernel from now on
    }
    const bool graphable = h->use_graphs && !trace && !fusep && (h->world == 1 || (peer && (peer_halo || assembled))) && issued > 0;
    if (graphable) {
      // every pointer and flag baked into the captured launches is part of the key
      unsigned long long key = 1469598103934665603ULL;
      auto mix = [&](unsigned long long v) { key = (key ^ v) * 1099511628211ULL; };
      for (const void* q : {(const void*)b, (const void*)D, (const void*)h->d_p, (const void*)h->d_u, (const void*)h->d_r,
                            (const void*)h->d_Ap, (const void*)h->d_E, (const void*)h->d_st, (const void*)h->d_p2})
        mix((unsigned long long)(uintptr_t)q);
      mix((unsigned long long)n);
      mix((unsigned long long)(energy ? 1 : 0) | (assembled ? 2 : 0) | (peer_halo ? 4 : 0) | (peer ? 8 : 0) | (single ? 16 : 0) |
          (use_ring(h) ? 32 : 0) | (fused ? 64 : 0) | (h->cg_fused_tma ? 128 : 0) | ((unsigned long long)ring_rows(h) << 8) | ((unsigned long long)(fused ? fused_rows(h) : 0) << 16));
      if (h->cg_graph == nullptr || h->cg_graph_key != key) {
        if (h->cg_graph) {
          cudaGraphExecDestroy(h->cg_graph);
          h->cg_graph = nullptr;
        }
        cudaGraph_t graph = nullptr;
        CUDA_TRY(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        int rc = TOPOPT_OK;
        for (int it = 0; it < n && rc == TOPOPT_OK; ++it) rc = one_iteration();
        if (rc == TOPOPT_OK) flush_batch();
        cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
        if (rc != TOPOPT_OK || ce != cudaSuccess || graph == nullptr) {
          if (graph) cudaGraphDestroy(graph);
          cudaGetLastError();
          h->use_graphs = false;  // fall back to plain launches from now on
          for (int it = 0; it < n; ++it) TRY(one_iteration());
          flush_batch();
        } else {
          ce = cudaGraphInstantiate(&h->cg_graph, graph, 0);
          cudaGraphDestroy(graph);
          if (ce != cudaSuccess) {
            h->cg_graph = nullptr;
            h->use_graphs = false;
            cudaGetLastError();
            for (int it = 0; it < n; ++it) TRY(one_iteration());
            flush_batch();
          } else {
            h->cg_graph_key = key;
            h->cg_graph_launches = (long long)n * launches_per_iter;
            CUDA_TRY(h, cudaGraphLaunch(h->cg_graph, h->stream));
          }
        }
      } else {
        CUDA_TRY(h, cudaGraphLaunch(h->cg_graph, h->stream));
        h->stats.kernel_launches += h->cg_graph_launches;
      }
    } else {
      for (int it = 0; it < n; ++it) TRY(one_iteration());
      flush_batch();
    }
    if (trace && issued == 0 && !tev.empty()) {
      mark();
      cudaStreamSynchronize(h->stream);
      const int per = (int)((tev.size() - 1) / n);
      std::vector<double> acc(per, 0.0);
      for (int it = 1; it < n; ++it)  // skip the first iteration
        for (int k = 0; k < per; ++k) {
          float ms = 0;
          cudaEventElapsedTime(&ms, tev[it * per + k], tev[it * per + k + 1]);
          acc[k] += ms;
        }
      fprintf(stderr, "[topopt trace rank %d] phases/iter (us):", h->rank);
      for (int k = 0; k < per; ++k) fprintf(stderr, " %.1f", 1e3 * acc[k] / std::max(1, n - 1));
      fprintf(stderr, "\n");
      for (cudaEvent_t e : tev) cudaEventDestroy(e);
    }
    TRY(check_launch(h, "cg iteration"));
    issued += n;
  }
  CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
  // ghost planes of the solution are needed by the sensitivity kernel
  TRY(exchange_halo(h, h->d_u));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  float ms = 0;
  CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->stats.cg_iterations += h->h_st->iters;
  h->stats.last_solve_ms = ms;
  if (res) {
    res->iters = h->h_st->iters;
    res->converged = h->h_st->converged;
    res->residual = h->h_st->res;
    res->tol = h->h_st->tol;
    res->solve_ms = ms;
  }
  if (h->h_st->nonfinite == 2)
    return fail(h, TOPOPT_ERR_NCCL, "peer-memory wait timed out: a neighbouring rank stopped participating");
  if (h->h_st->nonfinite)
    return fail(h, TOPOPT_ERR_NONFINITE, "CG: NaN or negative energy detected (EnergyCriteria / residual)");
  return TOPOPT_OK;
}

int run_sens(topopt_handle* h, const double* u, const double* v, double gsign, double* obj_out) {
  CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
  if (h->dim == 3 && h->nc == 3 && h->modal_ok) {
    if (u == v)
      LAUNCH(h, (k_sens_hex8_modal<false>), kWideGrid, h->g, u, v, h->d_E, h->d_dE, h->d_cell, h->d_grad, gsign, h->d_partials, h->d_st);
    else
      LAUNCH(h, (k_sens_hex8_modal<true>), kWideGrid, h->g, u, v, h->d_E, h->d_dE, h->d_cell, h->d_grad, gsign, h->d_partials, h->d_st);
  } else {
#define CALL(D, C) \
  LAUNCH(h, (k_sens<D, C>), kReduceBlocks, h->g, u, v, h->d_E, h->d_dE, h->d_cell, h->d_grad, gsign, h->d_partials, h->d_st)
    DISPATCH(h, CALL);
#undef CALL
  }
  TRY(check_launch(h, "k_sens"));
  if (h->world > 1)
    NCCL_TRY(h, g_nccl.AllReduce(h->d_st->sums, h->d_st->sums, 1, ncclDouble, ncclSum, h->comm, h->stream));
  CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(h->h_st, h->d_st, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  float ms = 0;
  CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->stats.last_sens_ms = ms;
  if (obj_out) *obj_out = h->h_st->sums[0];
  return TOPOPT_OK;
}

int penalize(topopt_handle* h, int kind, double p, double xmin, int pen_first) {
  if (kind < 0 || kind > 2) return fail(h, TOPOPT_ERR_INVALID, "unknown penalty kind");
  LAUNCH(h, k_penalize, grid_for(h->nloc_el, kWideGrid), (long long)h->nloc_el, h->d_rho, h->d_E, h->d_dE, kind, p, xmin,
         pen_first, h->proj_kind, h->proj_beta);
  h->stiffness_dirty = true;
  h->mg_dirty = true;
  return check_launch(h, "k_penalize");
}

// ---- filter internals ---------------------------------------------------------------------
int filter_run(topopt_filter* f, const double* in_dev, double* out_dev, int mode) {
  topopt_handle* h = f->h;
  const FilterGeo& fg = f->fg;
  const long long nn = (long long)fg.NX * fg.NY * fg.NZ;
  const long long ne = (long long)fg.nx * fg.ny * fg.nz;
  // shared-memory tiled stencil when the tile fits (it always does for rmin of a few elements)
  const int BY = fg.dim == 3 ? 4 : 8, BZ = fg.dim == 3 ? 2 : 1;
  const int wx = 2 * fg.R[0], wy = 2 * fg.R[1], wz = fg.dim == 3 ? 2 * fg.R[2] : 1;
  const size_t smem = sizeof(double) * ((size_t)(32 + wx - 1) * (BY + wy - 1) * (BZ + wz - 1) + (size_t)wx * wy * wz);
  const bool tiled = smem <= 160 * 1024 && !getenv("TOPOPT_FILTER_UNTILED");
  if (tiled && !f->attr_set) {
    CUDA_TRY(h, cudaFuncSetAttribute(k_filter_stencil<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(h, cudaFuncSetAttribute(k_filter_stencil<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    f->attr_set = true;
  }
  auto blocks = [&](int nxo, int nyo, int nzo) { return ((nxo + 31) / 32) * ((nyo + BY - 1) / BY) * ((nzo + BZ - 1) / BZ); };
  // cubic 3-D windows with R <= 3: fully unrolled kernel with the cone table in constant memory
  const int Rc = (fg.dim == 3 && fg.R[0] == fg.R[1] && fg.R[1] == fg.R[2] && fg.R[0] <= 3 && !getenv("TOPOPT_FILTER_GENERIC")) ? fg.R[0] : 0;
  if (Rc) {
    CUDA_TRY(h, cudaMemcpyToSymbolAsync(cW, f->h_w.data(), sizeof(double) * f->h_w.size(), 0, cudaMemcpyHostToDevice, h->stream));
    auto blocks3 = [&](int nxo, int nyo, int nzo) { return ((nxo + 31) / 32) * ((nyo + 3) / 4) * ((nzo + 3) / 4); };
    const bool fwd = mode == TOPOPT_FILTER_FORWARD;
    if (fwd) LAUNCH(h, k_filter_c2n, grid_for(nn, kWideGrid), fg, in_dev, f->d_nodal);
    const int nb = fwd ? blocks3(fg.nx, fg.ny, fg.nz) : blocks3(fg.NX, fg.NY, fg.NZ);
    const double* src = fwd ? f->d_nodal : in_dev;
    double* dst = fwd ? out_dev : f->d_nodal;
#define STENCIL3(RR)                                                                                  \
  if (fwd)                                                                                            \
    k_filter_stencil3<RR, false><<<nb, 128, 0, h->stream>>>(fg, src, f->d_den, dst);                  \
  else                                                                                                \
    k_filter_stencil3<RR, true><<<nb, 128, 0, h->stream>>>(fg, src, f->d_den, dst);
    if (Rc == 1) { STENCIL3(1) } else if (Rc == 2) { STENCIL3(2) } else { STENCIL3(3) }
#undef STENCIL3
    h->stats.kernel_launches += 1;
    if (!fwd) LAUNCH(h, k_filter_n2c_T, grid_for(ne, kWideGrid), fg, f->d_nodal, out_dev);
    return check_launch(h, "filter");
  }
  if (mode == TOPOPT_FILTER_FORWARD) {
    LAUNCH(h, k_filter_c2n, grid_for(nn, kWideGrid), fg, in_dev, f->d_nodal);
    if (tiled) {
      k_filter_stencil<false><<<blocks(fg.nx, fg.ny, fg.nz), kBlock, smem, h->stream>>>(fg, BY, BZ, f->d_w, f->d_nodal, f->d_den, out_dev);
      h->stats.kernel_launches += 1;
    } else {
      LAUNCH(h, (k_filter_n2c<0>), grid_for(ne, kWideGrid), fg, f->d_w, f->d_nodal, f->d_den, out_dev);
    }
  } else {
    if (tiled) {
      k_filter_stencil<true><<<blocks(fg.NX, fg.NY, fg.NZ), kBlock, smem, h->stream>>>(fg, BY, BZ, f->d_w, in_dev, f->d_den, f->d_nodal);
      h->stats.kernel_launches += 1;
    } else {
      LAUNCH(h, k_filter_c2n_T, grid_for(nn, kWideGrid), fg, f->d_w, in_dev, f->d_den, f->d_nodal);
    }
    LAUNCH(h, k_filter_n2c_T, grid_for(ne, kWideGrid), fg, f->d_nodal, out_dev);
  }
  return check_launch(h, "filter");
}

}  // namespace

// =================================================================================================
extern "C" {

const char* topopt_last_error(const topopt_handle* h) {
  if (h && !h->err.empty()) return h->err.c_str();
  return g_last_error.c_str();
}

int topopt_nccl_unique_id(void* out128) {
  if (!out128) return fail(nullptr, TOPOPT_ERR_INVALID, "topopt_nccl_unique_id: NULL argument");
  std::string err;
  if (!g_nccl.load(err)) return fail(nullptr, TOPOPT_ERR_NCCL, err);
  ncclUniqueId id;
  ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != ncclSuccess) return fail(nullptr, TOPOPT_ERR_NCCL, std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  std::memcpy(out128, &id, sizeof(id));
  return TOPOPT_OK;
}

// ---- peer-memory setup (world > 1, one process per GPU on one NVSwitch node) ---------------------
// export: [cudaIpcMemHandle_t of the PeerBlock][cudaIpcMemHandle_t of d_p][int32 nown] = 132 bytes
int topopt_ipc_export(topopt_handle* h, void* out, int64_t* nbytes) {
  if (!h || !out || !nbytes) return fail(h, TOPOPT_ERR_INVALID, "topopt_ipc_export: NULL argument");
  TRY(use_device(h));
  if (!h->d_peerblock) {
    CUDA_TRY(h, cudaMalloc((void**)&h->d_peerblock, sizeof(PeerBlock)));
    CUDA_TRY(h, cudaMemset(h->d_peerblock, 0, sizeof(PeerBlock)));
  }
  // blob: PeerBlock handle, handles of p, p2, r, r2, Ap, Ap2, owned plane count
  cudaIpcMemHandle_t hb;
  CUDA_TRY(h, cudaIpcGetMemHandle(&hb, h->d_peerblock));
  char* o = static_cast<char*>(out);
  std::memcpy(o, &hb, sizeof(hb));
  double* vecs[6] = {h->d_p, h->d_p2, h->d_r, h->d_r2, h->d_Ap, h->d_Ap2};
  for (int k = 0; k < 6; ++k) {
    cudaIpcMemHandle_t hv;
    CUDA_TRY(h, cudaIpcGetMemHandle(&hv, vecs[k]));
    std::memcpy(o + (1 + k) * sizeof(hb), &hv, sizeof(hv));
  }
  const int32_t nown = h->g.nown;
  std::memcpy(o + 7 * sizeof(hb), &nown, sizeof(nown));
  *nbytes = 7 * sizeof(hb) + sizeof(nown);
  return TOPOPT_OK;
}

// import: the export blobs of all ranks, concatenated in rank order (every rank calls this)
// unmap everything a previous import opened (idempotent)
static void ipc_close_all(topopt_handle* h) {
  for (void*& m : h->peer_mapped)
    if (m) {
      cudaIpcCloseMemHandle(m);
      m = nullptr;
    }
  for (int side = 0; side < 2; ++side)
    for (void*& m : h->peer_mapped_vec[side])
      if (m) {
        cudaIpcCloseMemHandle(m);
        m = nullptr;
      }
  for (int k = 0; k < 6; ++k) h->peer_lo[k] = h->peer_hi[k] = nullptr;
  h->peer_p_lo = h->peer_p_hi = nullptr;
  h->peer_ready = h->peer_fused_ready = false;
  cudaGetLastError();
}

int topopt_ipc_import(topopt_handle* h, const void* all, int64_t nbytes_each) {
  if (!h) return fail(h, TOPOPT_ERR_INVALID, "topopt_ipc_import: NULL handle");
  if (!all) {
    // all == NULL: give the peer-memory path up (some rank could not map its neighbours): every rank then uses NCCL
    TRY(use_device(h));
    ipc_close_all(h);
    return TOPOPT_OK;
  }
  if (h->world < 2 || h->world > kMaxRanks) return fail(h, TOPOPT_ERR_INVALID, "topopt_ipc_import: needs 2..8 ranks");
  if (nbytes_each != (int64_t)(7 * sizeof(cudaIpcMemHandle_t) + sizeof(int32_t)) || !h->d_peerblock)
    return fail(h, TOPOPT_ERR_INVALID, "topopt_ipc_import: call topopt_ipc_export first / bad blob size");
  TRY(use_device(h));
  ipc_close_all(h);  // a second import must not leak the first one's mappings
  PeerComm pc;
  std::memset(&pc, 0, sizeof(pc));
  pc.rank = h->rank;
  pc.world = h->world;
  const char* a = static_cast<const char*>(all);
  for (int r = 0; r < h->world; ++r) {
    const char* blob = a + (size_t)r * nbytes_each;
    cudaIpcMemHandle_t hb, hp;
    int32_t nown = 0;
    std::memcpy(&hb, blob, sizeof(hb));
    std::memcpy(&hp, blob + sizeof(hb), sizeof(hp));
    std::memcpy(&nown, blob + 7 * sizeof(hb), sizeof(nown));
    if (r == h->rank) {
      pc.block[r] = h->d_peerblock;
      continue;
    }
    void* m = nullptr;
    CUDA_TRY(h, cudaIpcOpenMemHandle(&m, hb, cudaIpcMemLazyEnablePeerAccess));
    h->peer_mapped[r] = m;
    pc.block[r] = static_cast<PeerBlock*>(m);
    if (r == h->rank - 1 || r == h->rank + 1) {
      void* mp = nullptr;
      CUDA_TRY(h, cudaIpcOpenMemHandle(&mp, hp, cudaIpcMemLazyEnablePeerAccess));
      h->peer_mapped[kMaxRanks + r] = mp;
      const int side = r == h->rank - 1 ? 0 : 1;
      if (side == 0) {
        h->peer_p_lo = static_cast<const double*>(mp);
        h->nown_lower = nown;
      } else {
        h->peer_p_hi = static_cast<const double*>(mp);
      }
      (side == 0 ? h->peer_lo : h->peer_hi)[0] = static_cast<const double*>(mp);
      for (int k = 1; k < 6; ++k) {
        cudaIpcMemHandle_t hv;
        std::memcpy(&hv, blob + (1 + k) * sizeof(hb), sizeof(hv));
        void* mv = nullptr;
        CUDA_TRY(h, cudaIpcOpenMemHandle(&mv, hv, cudaIpcMemLazyEnablePeerAccess));
        h->peer_mapped_vec[side][k] = mv;
        (side == 0 ? h->peer_lo : h->peer_hi)[k] = static_cast<const double*>(mv);
      }
    }
  }
  if (!h->d_peercomm) CUDA_TRY(h, cudaMalloc((void**)&h->d_peercomm, sizeof(PeerComm)));
  CUDA_TRY(h, cudaMemcpy(h->d_peercomm, &pc, sizeof(pc), cudaMemcpyHostToDevice));
  h->peer_ready = !getenv("TOPOPT_NO_PEER");
  h->peer_fused_ready = h->peer_ready;
  return TOPOPT_OK;
}

int topopt_create(const topopt_desc* d, topopt_handle** out) {
  if (!d || !out) return fail(nullptr, TOPOPT_ERR_INVALID, "topopt_create: NULL argument");
  *out = nullptr;
  if (!check_dims(d->dim, d->ncomp, d->nels)) return TOPOPT_ERR_INVALID;
  if (!d->Ke) return fail(nullptr, TOPOPT_ERR_INVALID, "topopt_create: Ke is NULL");
  if (d->world < 1 || d->rank < 0 || d->rank >= d->world)
    return fail(nullptr, TOPOPT_ERR_INVALID, "topopt_create: bad rank/world");
  if (d->world > 1 && !d->nccl_unique_id)
    return fail(nullptr, TOPOPT_ERR_INVALID, "topopt_create: world > 1 needs nccl_unique_id");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(nullptr, TOPOPT_ERR_NO_DEVICE,
                "topopt_create: no CUDA device available (libtopopt_cuda has no CPU fallback)");
  }
  if (d->device < 0 || d->device >= ndev) return fail(nullptr, TOPOPT_ERR_INVALID, "topopt_create: bad device ordinal");

  topopt_handle* h = new topopt_handle();
  h->id = g_next_id.fetch_add(1);
  h->dim = d->dim;
  h->nc = d->ncomp;
  h->ks = (1 << d->dim) * d->ncomp;
  h->gd = make_dims(d->dim, d->nels);
  h->device = d->device;
  h->rank = d->rank;
  h->world = d->world;
  const GridDims& gd = h->gd;
  h->nnodes = gd.nnodes;
  h->nel = gd.nel;
  h->ndof = gd.nnodes * h->nc;
  h->nnz = pattern_nnz(gd, h->nc);
  if (h->ndof >= (int64_t(1) << 31)) {
    delete h;
    return fail(nullptr, TOPOPT_ERR_INVALID, "topopt_create: more than 2^31 dofs are not supported");
  }
  for (int a = 0; a < 3; ++a) h->sizes[a] = a < d->dim ? d->sizes[a] : 1.0;
  for (int a = 0; a < d->dim; ++a)
    if (!(h->sizes[a] > 0)) {
      delete h;
      return fail(nullptr, TOPOPT_ERR_INVALID, "topopt_create: sizes must be positive");
    }
  h->cellvol = 1.0;
  for (int a = 0; a < d->dim; ++a) h->cellvol *= h->sizes[a];
  if (d->cellvolumes) {
    for (int64_t e = 0; e < h->nel; ++e)
      if (std::fabs(d->cellvolumes[e] - d->cellvolumes[0]) > 1e-12 * std::fabs(d->cellvolumes[0])) {
        delete h;
        return fail(nullptr, TOPOPT_ERR_INVALID, "topopt_create: non-uniform cellvolumes are not supported on the structured grid");
      }
    h->cellvol = d->cellvolumes[0];
  }
  std::memcpy(h->Ke, d->Ke, sizeof(double) * h->ks * h->ks);
  for (int r = 0; r < h->ks; ++r)
    for (int c = 0; c < h->ks; ++c)
      if (!std::isfinite(h->Ke[r + h->ks * c])) {
        delete h;
        return fail(nullptr, TOPOPT_ERR_NONFINITE, "topopt_create: Ke has non-finite entries");
      }
  double tr = 0.0;
  for (int r = 0; r < h->ks; ++r) tr += h->Ke[r * (h->ks + 1)];
  h->fixed_diag = d->fixed_diag > 0 ? d->fixed_diag : tr * (double)h->nel;  // solvers_api.jl:526-527

  // hex8 elasticity fast path: Khat = T' Ke T / 64 must have the 45-entry brick pattern
  if (h->dim == 3 && h->nc == 3 && !getenv("TOPOPT_KXU_DENSE")) {
    static const int corner[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
    static const int mode[8][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 0}, {0, 1, 1}, {1, 0, 1}, {1, 1, 1}};
    double H[8][8];
    for (int a = 0; a < 8; ++a)
      for (int m = 0; m < 8; ++m) {
        double v = 1.0;
        for (int dd = 0; dd < 3; ++dd)
          if (mode[m][dd]) v *= corner[a][dd] ? 1.0 : -1.0;
        H[a][m] = v;
      }
    std::vector<double> Kh(24 * 24, 0.0), tmp(24 * 24, 0.0);
    for (int r = 0; r < 24; ++r)  // tmp = Ke * T
      for (int m = 0; m < 8; ++m)
        for (int c = 0; c < 3; ++c) {
          double acc = 0.0;
          for (int a = 0; a < 8; ++a) acc += h->Ke[r + 24 * (3 * a + c)] * H[a][m];
          tmp[r + 24 * (3 * m + c)] = acc;
        }
    double kmax = 0.0;
    for (int m = 0; m < 8; ++m)  // Kh = T' * tmp / 64
      for (int c = 0; c < 3; ++c)
        for (int col = 0; col < 24; ++col) {
          double acc = 0.0;
          for (int a = 0; a < 8; ++a) acc += H[a][m] * tmp[(3 * a + c) + 24 * col];
          Kh[(3 * m + c) + 24 * col] = acc / 64.0;
          kmax = std::max(kmax, std::fabs(acc / 64.0));
        }
    std::vector<char> in_pattern(24 * 24, 0);
    const ModalEntry* pat = modal_pattern();
    for (int k = 0; k < 45; ++k) {
      in_pattern[pat[k].row + 24 * pat[k].col] = 1;
      h->Kh[k] = Kh[pat[k].row + 24 * pat[k].col];
    }
    bool ok = kmax > 0.0;
    for (int k = 0; k < 24 * 24 && ok; ++k)
      if (!in_pattern[k] && std::fabs(Kh[k]) > 1e-11 * kmax) ok = false;
    h->modal_ok = ok;
    if (ok) {
      // cubic cells: [a b b; ...] normal block, c [1 1; 1 1] shear pairs, [p q; q p] bilinear pairs, [e f f; ...] triple, g
      const double* K = h->Kh;
      auto eq = [&](double u, double v) { return std::fabs(u - v) <= 1e-13 * kmax; };
      bool cube = eq(K[0], K[4]) && eq(K[0], K[8]);
      for (int k : {1, 2, 3, 5, 6, 7}) cube = cube && eq(K[k], K[1]);
      for (int k = 9; k < 21; ++k) cube = cube && eq(K[k], K[9]);
      for (int blk = 0; blk < 3; ++blk) {
        const double* B = K + 21 + 4 * blk;
        cube = cube && eq(B[0], K[21]) && eq(B[3], K[21]) && eq(B[1], K[22]) && eq(B[2], K[22]);
      }
      cube = cube && eq(K[33], K[37]) && eq(K[33], K[41]);
      for (int k : {34, 35, 36, 38, 39, 40}) cube = cube && eq(K[k], K[34]);
      cube = cube && eq(K[42], K[43]) && eq(K[42], K[44]);
      h->modal_cube = cube;
      if (cube) {
        const double v[8] = {K[0] - K[1], K[1], K[9], K[21], K[22], K[33] - K[34], K[34], K[42]};
        std::memcpy(h->Kc, v, sizeof(v));
      }
    }
    if (const char* e = getenv("TOPOPT_KXU_TY")) h->kxu_ty = atoi(e);
    if (getenv("TOPOPT_FUSE_P")) h->no_fuse = false;
    if (getenv("TOPOPT_NO_GRAPH")) h->use_graphs = false;
    if (const char* e = getenv("TOPOPT_KXU_WAVES")) h->kxu_waves = atoi(e);
    if (const char* e = getenv("TOPOPT_KXU_NSYNC")) h->kxu_nsync = atoi(e);
    if (const char* e = getenv("TOPOPT_KXU_2ROW")) h->kxu_2row = atoi(e);
    if (const char* e = getenv("TOPOPT_KXU_RING")) h->kxu_ring = atoi(e);
    if (const char* e = getenv("TOPOPT_KXU_RING_MIN")) h->kxu_ring_min = atoi(e);
    if (const char* e = getenv("TOPOPT_KXU_CUBE")) h->kxu_cube = atoi(e);
    if (const char* e = getenv("TOPOPT_CG_FUSED")) h->cg_fused = atoi(e);
    if (const char* e = getenv("TOPOPT_CG_FUSED_GRID")) h->cg_fused_grid = std::max(1, atoi(e));
    if (const char* e = getenv("TOPOPT_KXU_GRID")) h->kxu_grid = std::max(1, atoi(e));
    if (const char* e = getenv("TOPOPT_CG_FUSED_MGPU")) h->cg_fused_single_only = atoi(e) == 0;
    if (const char* e = getenv("TOPOPT_CG_FUSED_TMA")) h->cg_fused_tma = atoi(e);
  }
  if (const char* e = getenv("TOPOPT_CG_VARIANT")) h->cg_variant_env = atoi(e);
  if (const char* e = getenv("TOPOPT_CG_PERSIST")) h->cg_persist = atoi(e);
  if (const char* e = getenv("TOPOPT_CG_FUSED_PDL")) h->cg_fused_pdl = atoi(e);

  // slab partition along the last axis
  const int NLg = (int)(h->dim == 3 ? gd.nz : gd.ny);
  const int NPg = NLg + 1;
  if (NLg < h->world) {
    delete h;
    return fail(nullptr, TOPOPT_ERR_INVALID, "topopt_create: fewer element layers than ranks");
  }
  const int e0 = (int)((int64_t)h->rank * NLg / h->world), e1 = (int)((int64_t)(h->rank + 1) * NLg / h->world);
  Geo& g = h->g;
  g.NX = (int)gd.NX;
  g.NY = h->dim == 3 ? (int)gd.NY : 1;
  g.nx = (int)gd.nx;
  g.ny = h->dim == 3 ? (int)gd.ny : 1;
  g.S = g.NX * g.NY;
  g.SE = g.nx * g.ny;
  g.NPg = NPg;
  g.NLg = NLg;
  g.p0 = e0 - 1;
  g.nlay = e1 - e0;
  g.nown = (h->rank == h->world - 1) ? (NPg - e0) : (e1 - e0);
  h->nloc_nodes = (int64_t)g.S * (g.nown + 2);
  h->nloc_dofs = h->nloc_nodes * h->nc;
  h->plane_dofs = (int64_t)g.S * h->nc;
  h->off = h->plane_dofs;
  h->nown_dofs = h->plane_dofs * g.nown;
  h->nloc_el = (int64_t)g.SE * (g.nown + 1);
  h->eoff = g.SE;
  h->nown_el = (int64_t)g.SE * g.nlay;

  // numbering
  h->block_of_node = ferrite_node_blocks(gd);
  h->node_of_block.resize(gd.nnodes);
  for (int64_t n = 0; n < gd.nnodes; ++n) h->node_of_block[h->block_of_node[n]] = n;
  if (d->cell_dofs) {
    const int nen = 1 << h->dim;
    int64_t nodes[8];
    for (int64_t e = 0; e < gd.nel; ++e) {
      cell_nodes(gd, e, nodes);
      for (int a = 0; a < nen; ++a)
        for (int c = 0; c < h->nc; ++c)
          if (d->cell_dofs[e * h->ks + a * h->nc + c] != h->block_of_node[nodes[a]] * h->nc + c + 1) {
            delete h;
            return fail(nullptr, TOPOPT_ERR_MISMATCH,
                        "topopt_create: caller's cell_dofs differ from the structured-grid Ferrite numbering");
          }
    }
  }
  std::vector<unsigned char> flags(gd.nnodes, 0);
  for (int64_t k = 0; k < d->n_prescribed; ++k) {
    const int64_t dof = d->prescribed_dofs[k] - 1;
    if (dof < 0 || dof >= h->ndof) {
      delete h;
      return fail(nullptr, TOPOPT_ERR_INVALID, "topopt_create: prescribed dof out of range");
    }
    flags[h->node_of_block[dof / h->nc]] |= (unsigned char)(1u << (dof % h->nc));
  }

  auto bail = [&](int rc) {
    topopt_destroy(h);
    return rc;
  };
#define CTRY(expr)                         \
  do {                                     \
    int rc__ = (expr);                     \
    if (rc__ != TOPOPT_OK) return bail(rc__); \
  } while (0)
#define CCUDA(expr)                                                                              \
  do {                                                                                           \
    cudaError_t e_ = (expr);                                                                     \
    if (e_ != cudaSuccess) {                                                                     \
      fail(nullptr, TOPOPT_ERR_CUDA, std::string(#expr ": ") + cudaGetErrorString(e_));          \
      return bail(TOPOPT_ERR_CUDA);                                                              \
    }                                                                                            \
  } while (0)

  CCUDA(cudaSetDevice(h->device));
  CCUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CCUDA(cudaEventCreate(&h->ev0));
  CCUDA(cudaEventCreate(&h->ev1));
  CCUDA(cudaMallocHost((void**)&h->h_st, sizeof(CGState)));
  CTRY(dev_alloc(h, &h->d_st, 1));
  CTRY(dev_alloc(h, &h->d_partials, (size_t)std::max(kWideGrid * 4, kMaxPartialBlocks)));
  CTRY(dev_alloc(h, &h->d_block, h->nloc_nodes));
  CTRY(dev_alloc(h, &h->d_fixed, h->nloc_nodes));
  for (double** v : {&h->d_b, &h->d_fload, &h->d_u, &h->d_r, &h->d_p, &h->d_p2, &h->d_Ap, &h->d_r2, &h->d_Ap2, &h->d_D, &h->d_rhs, &h->d_lam, &h->d_tmp})
    CTRY(dev_alloc(h, v, h->nloc_dofs));
  for (double** v : {&h->d_E, &h->d_dE, &h->d_rho, &h->d_cell, &h->d_grad}) CTRY(dev_alloc(h, v, h->nloc_el));
  CTRY(dev_alloc(h, &h->d_full_dof, h->ndof));
  for (double** v : {&h->d_full_el, &h->d_design, &h->d_xf, &h->d_gfull}) CTRY(dev_alloc(h, v, h->nel));

  {  // local numbering and flags
    std::vector<int> lblock(h->nloc_nodes, 0);
    std::vector<unsigned char> lfixed(h->nloc_nodes, 0);
    for (int lp = 0; lp < g.nown + 2; ++lp) {
      const int gp = lp + g.p0;
      if (gp < 0 || gp >= NPg) continue;
      for (int64_t q = 0; q < g.S; ++q) {
        lblock[(int64_t)lp * g.S + q] = (int)h->block_of_node[(int64_t)gp * g.S + q];
        lfixed[(int64_t)lp * g.S + q] = flags[(int64_t)gp * g.S + q];
      }
    }
    CCUDA(cudaMemcpyAsync(h->d_block, lblock.data(), sizeof(int) * h->nloc_nodes, cudaMemcpyHostToDevice, h->stream));
    CCUDA(cudaMemcpyAsync(h->d_fixed, lfixed.data(), h->nloc_nodes, cudaMemcpyHostToDevice, h->stream));
    CCUDA(cudaStreamSynchronize(h->stream));
  }
  if (h->world > 1) {
    std::string err;
    if (!g_nccl.load(err)) {
      fail(nullptr, TOPOPT_ERR_NCCL, err);
      return bail(TOPOPT_ERR_NCCL);
    }
    ncclUniqueId id;
    std::memcpy(&id, d->nccl_unique_id, sizeof(id));
    ncclResult_t r = g_nccl.CommInitRank(&h->comm, h->world, id, h->rank);
    if (r != ncclSuccess) {
      fail(nullptr, TOPOPT_ERR_NCCL, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r));
      return bail(TOPOPT_ERR_NCCL);
    }
  }
  CTRY(use_device(h));
  if (d->fixedload) {
    CTRY(upload_dofs(h, d->fixedload, h->d_fload));
    CCUDA(cudaMemcpyAsync(h->d_b, h->d_fload, sizeof(double) * h->nloc_dofs, cudaMemcpyDeviceToDevice, h->stream));
#define CALL(D, C) LAUNCH(h, (k_apply_zero<C>), grid_for(h->nloc_nodes, kWideGrid), (long long)h->nloc_nodes, h->d_fixed, h->d_b)
    DISPATCH(h, CALL);
#undef CALL
  }
  // E = 1 until a density is set (solver.vars = ones, solvers_api.jl:501 with penalty p = 1)
  {
    std::vector<double> ones(h->nloc_el, 1.0);
    CCUDA(cudaMemcpyAsync(h->d_rho, ones.data(), sizeof(double) * h->nloc_el, cudaMemcpyHostToDevice, h->stream));
    CCUDA(cudaMemcpyAsync(h->d_E, ones.data(), sizeof(double) * h->nloc_el, cudaMemcpyHostToDevice, h->stream));
    CCUDA(cudaMemcpyAsync(h->d_dE, ones.data(), sizeof(double) * h->nloc_el, cudaMemcpyHostToDevice, h->stream));
    CCUDA(cudaStreamSynchronize(h->stream));
  }
  CTRY(check_launch(h, "topopt_create"));
  CCUDA(cudaStreamSynchronize(h->stream));
  std::memset(&h->stats, 0, sizeof(h->stats));
#undef CTRY
#undef CCUDA
  *out = h;
  return TOPOPT_OK;
}

int topopt_destroy(topopt_handle* h) {
  if (!h) return TOPOPT_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->cg_graph) cudaGraphExecDestroy(h->cg_graph);
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  mg_free(h->mg);
  h->mg = nullptr;
  if (h->d_tmaps) cudaFree(h->d_tmaps);
  if (h->d_barrier) cudaFree(h->d_barrier);
  ipc_close_all(h);
  if (h->d_peerblock) cudaFree(h->d_peerblock);
  if (h->d_peercomm) cudaFree(h->d_peercomm);
  void* ptrs[] = {h->d_block, h->d_fixed, h->d_b, h->d_fload, h->d_u, h->d_r, h->d_p, h->d_p2, h->d_Ap, h->d_r2, h->d_Ap2, h->d_D, h->d_rhs, h->d_lam,
                  h->d_tmp, h->d_E, h->d_dE, h->d_rho, h->d_cell, h->d_grad, h->d_full_dof, h->d_full_el, h->d_design,
                  h->d_xf, h->d_gfull, h->d_partials, h->d_st, h->d_nbr_start, h->d_rowptr, h->d_col, h->d_nz, h->d_fasm,
                  h->d_dv, h->d_dc, h->d_xnew};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (h->h_st) cudaFreeHost(h->h_st);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return TOPOPT_OK;
}

int topopt_get_stats(topopt_handle* h, topopt_stats* out) {
  if (!h || !out) return fail(h, TOPOPT_ERR_INVALID, "topopt_get_stats: NULL argument");
  h->stats.ndof = h->ndof;
  h->stats.nel = h->nel;
  h->stats.nnodes = h->nnodes;
  h->stats.nnz = h->nnz;
  h->stats.ndof_local = h->nown_dofs;
  h->stats.nel_local = h->nown_el;
  *out = h->stats;
  return TOPOPT_OK;
}

int topopt_reset_stats(topopt_handle* h) {
  if (!h) return fail(h, TOPOPT_ERR_INVALID, "topopt_reset_stats: NULL handle");
  std::memset(&h->stats, 0, sizeof(h->stats));
  return TOPOPT_OK;
}

int topopt_set_density(topopt_handle* h, const double* rho, int32_t kind, double p, double xmin, int32_t pen_first) {
  if (!h) return fail(h, TOPOPT_ERR_INVALID, "topopt_set_density: NULL handle");
  TRY(use_device(h));
  if (rho) TRY(upload_elems(h, rho, h->d_rho));
  TRY(penalize(h, kind, p, xmin, pen_first));
  return sync(h);
}

int topopt_set_projection(topopt_handle* h, int32_t proj_kind, double beta) {
  if (!h) return fail(h, TOPOPT_ERR_INVALID, "topopt_set_projection: NULL handle");
  if (proj_kind < 0 || proj_kind > 2 || !std::isfinite(beta)) return fail(h, TOPOPT_ERR_INVALID, "topopt_set_projection: bad projection");
  h->proj_kind = proj_kind;
  h->proj_beta = beta;
  return TOPOPT_OK;
}

int topopt_set_stiffness(topopt_handle* h, const double* E, const double* dE) {
  if (!h || !E) return fail(h, TOPOPT_ERR_INVALID, "topopt_set_stiffness: NULL argument");
  TRY(use_device(h));
  TRY(upload_elems(h, E, h->d_E));
  if (dE) TRY(upload_elems(h, dE, h->d_dE));
  h->stiffness_dirty = true;
  h->mg_dirty = true;
  return sync(h);
}

int topopt_get_stiffness(topopt_handle* h, double* E, double* dE) {
  if (!h) return fail(h, TOPOPT_ERR_INVALID, "topopt_get_stiffness: NULL handle");
  TRY(use_device(h));
  if (E) TRY(download_elems(h, h->d_E, E));
  if (E && dE) TRY(sync(h));
  if (dE) TRY(download_elems(h, h->d_dE, dE));
  return sync(h);
}

int topopt_apply(topopt_handle* h, const double* x, double* y) {
  if (!h) return fail(h, TOPOPT_ERR_INVALID, "topopt_apply: NULL handle");
  TRY(use_device(h));
  if (x) TRY(upload_dofs(h, x, h->d_p));
  CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
  TRY(launch_apply<false>(h, h->d_p, h->d_Ap, FIN_NONE));
  CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
  if (y) TRY(download_dofs(h, h->d_Ap, y));
  TRY(sync(h));
  float ms = 0;
  CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->stats.last_apply_ms = ms;
  return TOPOPT_OK;
}

// topopt_apply through one specific kernel generation, with the fused dot products the CG loop uses
int topopt_apply_ex(topopt_handle* h, const double* x, double* y, int32_t kernel, double* dots) {
  if (!h) return fail(h, TOPOPT_ERR_INVALID, "topopt_apply_ex: NULL handle");
  TRY(use_device(h));
  const bool hex8 = h->dim == 3 && h->nc == 3 && h->modal_ok;
  if (kernel < 0 || kernel > 4) return fail(h, TOPOPT_ERR_INVALID, "topopt_apply_ex: unknown kernel");
  if (kernel >= 2 && !hex8) return fail(h, TOPOPT_ERR_INVALID, "topopt_apply_ex: hex8 modal kernels need a brick/isotropic hex8 Ke");
  if (x) TRY(upload_dofs(h, x, h->d_p));
  TRY(exchange_halo(h, h->d_p));
  CGState zero;
  std::memset(&zero, 0, sizeof(zero));
  zero.world = 1;  // the reduction stays rank-local; summed below
  CUDA_TRY(h, cudaMemcpyAsync(h->d_st, &zero, sizeof(CGState), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
  switch (kernel) {
    case 0: TRY(launch_apply<true>(h, h->d_p, h->d_Ap, FIN_NONE)); break;
    case 1: {  // dense node-centric gather
      const int grid = grid_for((long long)h->g.S * h->g.nown, kReduceBlocks);
#define CALL(D, C) LAUNCH(h, (k_apply<D, C, true>), grid, h->g, h->d_p, h->d_Ap, h->d_E, h->d_fixed, h->fixed_diag, h->d_partials, h->d_st, FIN_NONE)
      DISPATCH(h, CALL);
#undef CALL
      TRY(check_launch(h, "k_apply"));
      break;
    }
    case 2: TRY((launch_hex8_modal<16, true, false, false, true>(h, h->d_p, h->d_Ap, FIN_NONE, nullptr, nullptr))); break;
    case 3: TRY((launch_hex8_modal2<12, true, false>(h, h->d_p, h->d_Ap, FIN_NONE))); break;
    case 4: TRY((launch_hex8_ring<2, false>(h, h->d_p, h->d_Ap, FIN_NONE))); break;
  }
  CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
  if (h->world > 1)
    NCCL_TRY(h, g_nccl.AllReduce(h->d_st->sums, h->d_st->sums, 2, ncclDouble, ncclSum, h->comm, h->stream));
  if (y) TRY(download_dofs(h, h->d_Ap, y));
  CUDA_TRY(h, cudaMemcpyAsync(h->h_st, h->d_st, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  TRY(sync(h));
  if (dots) {
    dots[0] = h->h_st->sums[0];
    dots[1] = kernel == 4 ? h->h_st->sums[1] : 0.0;
  }
  float ms = 0;
  CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->stats.last_apply_ms = ms;
  return TOPOPT_OK;
}

int topopt_set_jacobi(topopt_handle* h, const double* diag) {
  if (!h) return fail(h, TOPOPT_ERR_INVALID, "topopt_set_jacobi: NULL handle");
  TRY(use_device(h));
  if (diag) {
    TRY(upload_dofs(h, diag, h->d_D));
  } else {
#define CALL(D, C) LAUNCH(h, (k_diag<D, C>), grid_for((long long)h->g.S * h->g.nown, kWideGrid), h->g, h->d_D, h->d_E, h->d_fixed, h->fixed_diag)
    DISPATCH(h, CALL);
#undef CALL
    TRY(check_launch(h, "k_diag"));
  }
  h->have_jacobi = true;
  return sync(h);
}

int topopt_solve(topopt_handle* h, const double* rhs, double* u, const topopt_cg_opts* opts, topopt_cg_result* result) {
  if (!h || !opts) return fail(h, TOPOPT_ERR_INVALID, "topopt_solve: NULL argument");
  if (opts->maxiter < 0 || !(opts->abstol >= 0) || !(opts->reltol >= 0))
    return fail(h, TOPOPT_ERR_INVALID, "topopt_solve: bad tolerances / maxiter");
  TRY(use_device(h));
  const double* b = h->d_b;
  if (rhs) {
    TRY(upload_dofs(h, rhs, h->d_rhs));
#define CALL(D, C) LAUNCH(h, (k_apply_zero<C>), grid_for(h->nloc_nodes, kWideGrid), (long long)h->nloc_nodes, h->d_fixed, h->d_rhs)
    DISPATCH(h, CALL);  // apply_zero!(rhs, ch), solvers_api.jl:364-367
#undef CALL
    b = h->d_rhs;
  }
  TRY(cg_solve(h, b, opts, result));
  if (u) {
    TRY(download_dofs(h, h->d_u, u));
    TRY(sync(h));
  }
  return TOPOPT_OK;
}

int topopt_compliance(topopt_handle* h, const double* u, double* obj, double* cell_comp, double* grad) {
  if (!h) return fail(h, TOPOPT_ERR_INVALID, "topopt_compliance: NULL handle");
  TRY(use_device(h));
  if (u) TRY(upload_dofs(h, u, h->d_u));
  TRY(run_sens(h, h->d_u, h->d_u, -1.0, obj));
  if (cell_comp) {
    TRY(download_elems(h, h->d_cell, cell_comp));
    TRY(sync(h));
  }
  if (grad) TRY(download_elems(h, h->d_grad, grad));
  return sync(h);
}

int topopt_bilinear_sens(topopt_handle* h, const double* lambda, const double* u, double* cell_out, double* grad) {
  if (!h) return fail(h, TOPOPT_ERR_INVALID, "topopt_bilinear_sens: NULL handle");
  TRY(use_device(h));
  if (u) TRY(upload_dofs(h, u, h->d_u));
  if (lambda) TRY(upload_dofs(h, lambda, h->d_lam));
  TRY(run_sens(h, h->d_u, h->d_lam, 1.0, nullptr));
  if (cell_out) {
    TRY(download_elems(h, h->d_cell, cell_out));
    TRY(sync(h));
  }
  if (grad) TRY(download_elems(h, h->d_grad, grad));
  return sync(h);
}

int topopt_dot(topopt_handle* h, const double* a, const double* b, double* out) {
  if (!h || !out) return fail(h, TOPOPT_ERR_INVALID, "topopt_dot: NULL argument");
  TRY(use_device(h));
  const double* da = h->d_fload;
  const double* db = h->d_u;
  if (a) {
    TRY(upload_dofs(h, a, h->d_tmp));
    da = h->d_tmp;
  }
  if (b) {
    TRY(upload_dofs(h, b, h->d_rhs));
    db = h->d_rhs;
  }
  LAUNCH(h, k_dot, kReduceBlocks, h->off, h->nown_dofs, da, db, h->d_partials, h->d_st);
  TRY(check_launch(h, "k_dot"));
  if (h->world > 1)
    NCCL_TRY(h, g_nccl.AllReduce(h->d_st->sums, h->d_st->sums, 1, ncclDouble, ncclSum, h->comm, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(h->h_st, h->d_st, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  TRY(sync(h));
  *out = h->h_st->sums[0];
  return TOPOPT_OK;
}

// u' K u with the current stiffness (FEA.getcompliance, src/FEA/FEA.jl:40); u NULL = resident solution
int topopt_quadratic_form(topopt_handle* h, const double* u, double* out) {
  if (!h || !out) return fail(h, TOPOPT_ERR_INVALID, "topopt_quadratic_form: NULL argument");
  TRY(use_device(h));
  if (u) {
    TRY(upload_dofs(h, u, h->d_p));
  } else {
    CUDA_TRY(h, cudaMemcpyAsync(h->d_p, h->d_u, sizeof(double) * h->nloc_dofs, cudaMemcpyDeviceToDevice, h->stream));
    TRY(exchange_halo(h, h->d_p));
  }
  CGState zero;
  std::memset(&zero, 0, sizeof(zero));
  zero.world = 1;
  CUDA_TRY(h, cudaMemcpyAsync(h->d_st, &zero, sizeof(CGState), cudaMemcpyHostToDevice, h->stream));
  if (h->assembled && !h->stiffness_dirty && h->world == 1) {
    TRY(launch_spmv<true>(h, h->d_p, h->d_Ap, FIN_NONE));
  } else {
    TRY(launch_apply<true>(h, h->d_p, h->d_Ap, FIN_NONE));
  }
  if (h->world > 1)
    NCCL_TRY(h, g_nccl.AllReduce(h->d_st->sums, h->d_st->sums, 1, ncclDouble, ncclSum, h->comm, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(h->h_st, h->d_st, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  TRY(sync(h));
  *out = h->h_st->sums[0];
  return TOPOPT_OK;
}

// swap the resident solution with the resident lambda vector: solve (T), swap, solve adjoint
// (lambda), then topopt_bilinear_sens(h, NULL, NULL, ...) evaluates T_e' Ke lambda_e.
int topopt_swap_solution_lambda(topopt_handle* h) {
  if (!h) return fail(h, TOPOPT_ERR_INVALID, "topopt_swap_solution_lambda: NULL handle");
  TRY(sync(h));
  std::swap(h->d_u, h->d_lam);
  return TOPOPT_OK;
}

// ---- assembled path ---------------------------------------------------------------------------
}  // extern "C"

namespace {

int build_pattern(topopt_handle* h) {
  if (h->pattern_ready) return TOPOPT_OK;
  if (h->world > 1) return fail(h, TOPOPT_ERR_INVALID, "the assembled operator is single-GPU only");
  if (h->nnz >= (int64_t(1) << 31)) return fail(h, TOPOPT_ERR_INVALID, "assembled pattern exceeds 2^31 non-zeros; use the matrix-free operator");
  const GridDims& gd = h->gd;
  const int nc = h->nc;
  std::vector<long long> nbr_start(gd.nnodes + 1, 0);
  std::vector<int> rowptr(h->ndof + 1, 0);
  for (int64_t n = 0; n < gd.nnodes; ++n) {
    const int64_t i = n % gd.NX, j = (n / gd.NX) % gd.NY, k = n / (gd.NX * gd.NY);
    const int cx = 1 + (i > 0) + (i < gd.NX - 1);
    const int cy = 1 + (j > 0) + (j < gd.NY - 1);
    const int cz = gd.dim == 3 ? 1 + (k > 0) + (k < gd.NZ - 1) : 1;
    const int cnt = cx * cy * cz;
    nbr_start[n + 1] = nbr_start[n] + cnt;
    for (int c = 0; c < nc; ++c) rowptr[n * nc + c + 1] = (int)((long long)nc * nc * nbr_start[n] + (long long)(c + 1) * nc * cnt);
  }
  TRY(dev_alloc(h, &h->d_nbr_start, gd.nnodes + 1));
  TRY(dev_alloc(h, &h->d_rowptr, h->ndof + 1));
  TRY(dev_alloc(h, &h->d_col, h->nnz));
  TRY(dev_alloc(h, &h->d_nz, h->nnz));
  TRY(dev_alloc(h, &h->d_fasm, h->ndof));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_nbr_start, nbr_start.data(), sizeof(long long) * (gd.nnodes + 1), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_rowptr, rowptr.data(), sizeof(int) * (h->ndof + 1), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  h->pattern_ready = true;
  return TOPOPT_OK;
}

// CSC position (Ferrite order, as topopt_csc_pattern emits it) -> internal CSR position
void build_export_map(topopt_handle* h) {
  if (!h->export_map.empty()) return;
  const GridDims& gd = h->gd;
  const int nc = h->nc;
  h->export_map.resize(h->nnz);
  std::vector<long long> nbr_start(gd.nnodes + 1, 0);
  auto count = [&](int64_t n) {
    const int64_t i = n % gd.NX, j = (n / gd.NX) % gd.NY, k = n / (gd.NX * gd.NY);
    const int cx = 1 + (i > 0) + (i < gd.NX - 1);
    const int cy = 1 + (j > 0) + (j < gd.NY - 1);
    const int cz = gd.dim == 3 ? 1 + (k > 0) + (k < gd.NZ - 1) : 1;
    return cx * cy * cz;
  };
  for (int64_t n = 0; n < gd.nnodes; ++n) nbr_start[n + 1] = nbr_start[n] + count(n);
  struct Ent {
    int64_t fdof;
    int64_t pos;
  };
  std::vector<Ent> ents;
  int64_t out = 0;
  for (int64_t b = 0; b < gd.nnodes; ++b) {  // Ferrite column block
    const int64_t m = h->node_of_block[b];   // column node (lexicographic)
    const int64_t i = m % gd.NX, j = (m / gd.NX) % gd.NY, k = m / (gd.NX * gd.NY);
    for (int c2 = 0; c2 < nc; ++c2) {
      ents.clear();
      // rows of column (m,c2): all neighbour nodes n of m and components c; K[(n,c),(m,c2)]
      for (int64_t dk = (gd.dim == 3 ? -1 : 0); dk <= (gd.dim == 3 ? 1 : 0); ++dk)
        for (int64_t dj = -1; dj <= 1; ++dj)
          for (int64_t di = -1; di <= 1; ++di) {
            const int64_t ii = i + di, jj = j + dj, kk = k + dk;
            if (ii < 0 || ii >= gd.NX || jj < 0 || jj >= gd.NY || kk < 0 || kk >= gd.NZ) continue;
            const int64_t n = ii + gd.NX * (jj + gd.NY * kk);
            // slot of m in row-node n's neighbour list: neighbours of n that precede m
            int slot = 0;
            for (int64_t ek = (gd.dim == 3 ? -1 : 0); ek <= (gd.dim == 3 ? 1 : 0); ++ek)
              for (int64_t ej = -1; ej <= 1; ++ej)
                for (int64_t ei = -1; ei <= 1; ++ei) {
                  const int64_t a = ii + ei, bb = jj + ej, cc = kk + ek;
                  if (a < 0 || a >= gd.NX || bb < 0 || bb >= gd.NY || cc < 0 || cc >= gd.NZ) continue;
                  const int64_t q = a + gd.NX * (bb + gd.NY * cc);
                  if (q < m) ++slot;
                }
            const int cnt = count(n);
            for (int c = 0; c < nc; ++c) {
              const int64_t pos = (int64_t)nc * nc * nbr_start[n] + (int64_t)c * nc * cnt + (int64_t)slot * nc + c2;
              ents.push_back({h->block_of_node[n] * nc + c, pos});
            }
          }
      std::sort(ents.begin(), ents.end(), [](const Ent& a, const Ent& b2) { return a.fdof < b2.fdof; });
      for (const Ent& e : ents) h->export_map[out++] = e.pos;
    }
  }
}

int do_assemble(topopt_handle* h) {
  TRY(build_pattern(h));
  const long long total = (long long)h->nnodes * (h->dim == 3 ? 27 : 9);
#define CALL(D, C) LAUNCH(h, (k_assemble<D, C>), grid_for(total, kWideGrid), h->g, h->d_nbr_start, h->d_E, h->d_nz, h->d_col)
  DISPATCH(h, CALL);
#undef CALL
  TRY(check_launch(h, "k_assemble"));
#define CALL(D, C) LAUNCH(h, (k_csr_absdiag<D, C>), kReduceBlocks, h->g, h->d_nbr_start, h->d_nz, (double*)nullptr, h->d_partials, h->d_st)
  DISPATCH(h, CALL);
#undef CALL
  TRY(check_launch(h, "k_csr_absdiag"));
  const unsigned char* gfixed = h->d_fixed + h->g.S;  // single GPU: local plane 1 == global plane 0
  if (h->nc == 1)
    LAUNCH(h, (k_csr_apply_bc<1>), grid_for(h->ndof, kWideGrid), (long long)h->ndof, h->d_rowptr, h->d_col, h->d_nz, gfixed, h->d_st, 1.0 / (double)h->ndof);
  else if (h->nc == 2)
    LAUNCH(h, (k_csr_apply_bc<2>), grid_for(h->ndof, kWideGrid), (long long)h->ndof, h->d_rowptr, h->d_col, h->d_nz, gfixed, h->d_st, 1.0 / (double)h->ndof);
  else
    LAUNCH(h, (k_csr_apply_bc<3>), grid_for(h->ndof, kWideGrid), (long long)h->ndof, h->d_rowptr, h->d_col, h->d_nz, gfixed, h->d_st, 1.0 / (double)h->ndof);
  TRY(check_launch(h, "k_csr_apply_bc"));
  h->assembled = true;
  h->stiffness_dirty = false;
  return TOPOPT_OK;
}

}  // namespace

extern "C" {

int topopt_assemble(topopt_handle* h, double* nzval, double* f) {
  if (!h) return fail(h, TOPOPT_ERR_INVALID, "topopt_assemble: NULL handle");
  TRY(use_device(h));
  TRY(do_assemble(h));
  if (nzval) {
    build_export_map(h);
    std::vector<double> internal(h->nnz), outv(h->nnz);
    CUDA_TRY(h, cudaMemcpyAsync(internal.data(), h->d_nz, sizeof(double) * h->nnz, cudaMemcpyDeviceToHost, h->stream));
    TRY(sync(h));
    for (int64_t k = 0; k < h->nnz; ++k) outv[k] = internal[h->export_map[k]];
    CUDA_TRY(h, cudaMemcpyAsync(nzval, outv.data(), sizeof(double) * h->nnz, cudaMemcpyDefault, h->stream));
    TRY(sync(h));
    h->stats.d2h_bytes += sizeof(double) * h->nnz;
  }
  if (f) TRY(download_dofs(h, h->d_b, f));  // f = fixedload with prescribed entries zeroed (assemble.jl:51,87)
  return sync(h);
}

int topopt_spmv(topopt_handle* h, const double* x, double* y) {
  if (!h) return fail(h, TOPOPT_ERR_INVALID, "topopt_spmv: NULL handle");
  TRY(use_device(h));
  if (!h->assembled || h->stiffness_dirty) TRY(do_assemble(h));
  if (x) TRY(upload_dofs(h, x, h->d_p));
  TRY(launch_spmv<false>(h, h->d_p, h->d_Ap, FIN_NONE));
  if (y) TRY(download_dofs(h, h->d_Ap, y));
  return sync(h);
}

// ---- filters ------------------------------------------------------------------------------------
int topopt_filter_create(topopt_handle* h, double rmin, topopt_filter** out) {
  if (!h || !out) return fail(h, TOPOPT_ERR_INVALID, "topopt_filter_create: NULL argument");
  *out = nullptr;
  if (!(rmin > 0) || !std::isfinite(rmin)) return fail(h, TOPOPT_ERR_INVALID, "topopt_filter_create: rmin must be positive");
  TRY(use_device(h));
  topopt_filter* f = new topopt_filter();
  f->h = h;
  f->rmin = rmin;
  FilterGeo& fg = f->fg;
  const GridDims& gd = h->gd;
  fg.dim = h->dim;
  fg.nx = (int)gd.nx;
  fg.ny = (int)gd.ny;
  fg.nz = (int)gd.nz;
  fg.NX = (int)gd.NX;
  fg.NY = (int)gd.NY;
  fg.NZ = (int)gd.NZ;
  for (int a = 0; a < 3; ++a) fg.R[a] = 0;
  for (int a = 0; a < h->dim; ++a) {
    fg.R[a] = (int)std::ceil(rmin / h->sizes[a] + 0.5 + 1e-9) - 1;
    if (fg.R[a] < 1) fg.R[a] = 1;
    if (fg.R[a] > 64) {
      delete f;
      return fail(h, TOPOPT_ERR_INVALID, "topopt_filter_create: rmin spans more than 64 elements");
    }
  }
  const int wx = 2 * fg.R[0], wy = 2 * fg.R[1], wz = h->dim == 3 ? 2 * fg.R[2] : 1;
  std::vector<double> w((size_t)wx * wy * wz, 0.0);
  double wsum = 0.0;
  for (int tz = 0; tz < wz; ++tz)
    for (int ty = 0; ty < wy; ++ty)
      for (int tx = 0; tx < wx; ++tx) {
        // node offset o = 1 - R + t from the cell index; distance to the centroid per axis (o - 1/2) h
        const double ddx = (0.5 - fg.R[0] + tx) * h->sizes[0];
        const double ddy = (0.5 - fg.R[1] + ty) * h->sizes[1];
        const double ddz = h->dim == 3 ? (0.5 - fg.R[2] + tz) * h->sizes[2] : 0.0;
        const double dist = std::sqrt(ddx * ddx + ddy * ddy + ddz * ddz);
        const double wt = dist < rmin ? rmin - dist : 0.0;  // strict '<' (CheqFilters.jl:99)
        w[tx + (size_t)wx * (ty + (size_t)wy * tz)] = wt;
        wsum += wt;
      }
  f->h_w = w;
  if (wsum == 0.0) {
    delete f;
    return fail(h, TOPOPT_ERR_INVALID,
                "DensityFilterFun: no neighbouring nodes were found within the filter radius `rmin` for any element; "
                "increase `rmin` (mesh-coordinate units)");  // density_filter.jl:98-105
  }
  auto bail = [&](int rc) {
    topopt_filter_destroy(f);
    return rc;
  };
  int rc;
  if ((rc = dev_alloc(h, &f->d_w, w.size())) != TOPOPT_OK) return bail(rc);
  if ((rc = dev_alloc(h, &f->d_den, h->nel)) != TOPOPT_OK) return bail(rc);
  if ((rc = dev_alloc(h, &f->d_nodal, h->nnodes)) != TOPOPT_OK) return bail(rc);
  if ((rc = dev_alloc(h, &f->d_in, h->nel)) != TOPOPT_OK) return bail(rc);
  if ((rc = dev_alloc(h, &f->d_out, h->nel)) != TOPOPT_OK) return bail(rc);
  if (cudaMemcpyAsync(f->d_w, w.data(), sizeof(double) * w.size(), cudaMemcpyHostToDevice, h->stream) != cudaSuccess)
    return bail(fail(h, TOPOPT_ERR_CUDA, "filter weight upload failed"));
  LAUNCH(h, (k_filter_n2c<1>), grid_for(h->nel, kWideGrid), fg, f->d_w, (const double*)nullptr, (const double*)nullptr, f->d_den);
  if ((rc = check_launch(h, "k_filter_n2c<1>")) != TOPOPT_OK) return bail(rc);
  if ((rc = sync(h)) != TOPOPT_OK) return bail(rc);
  *out = f;
  return TOPOPT_OK;
}

int topopt_filter_apply(topopt_filter* f, const double* x, double* y, int32_t mode) {
  if (!f) return fail(nullptr, TOPOPT_ERR_INVALID, "topopt_filter_apply: NULL filter");
  topopt_handle* h = f->h;
  if (mode != TOPOPT_FILTER_FORWARD && mode != TOPOPT_FILTER_TRANSPOSE) return fail(h, TOPOPT_ERR_INVALID, "topopt_filter_apply: bad mode");
  TRY(use_device(h));
  if (x) {
    CUDA_TRY(h, cudaMemcpyAsync(f->d_in, x, sizeof(double) * h->nel, cudaMemcpyDefault, h->stream));
    h->stats.h2d_bytes += sizeof(double) * h->nel;
  }
  CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
  TRY(filter_run(f, f->d_in, f->d_out, mode));
  CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
  if (y) {
    CUDA_TRY(h, cudaMemcpyAsync(y, f->d_out, sizeof(double) * h->nel, cudaMemcpyDefault, h->stream));
    h->stats.d2h_bytes += sizeof(double) * h->nel;
  }
  TRY(sync(h));
  float ms = 0;
  CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->stats.last_filter_ms = ms;
  return TOPOPT_OK;
}

int topopt_filter_destroy(topopt_filter* f) {
  if (!f) return TOPOPT_OK;
  cudaSetDevice(f->h->device);
  cudaStreamSynchronize(f->h->stream);
  for (void* p : {(void*)f->d_w, (void*)f->d_den, (void*)f->d_nodal, (void*)f->d_in, (void*)f->d_out})
    if (p) cudaFree(p);
  delete f;
  return TOPOPT_OK;
}

// ---- fused SIMP evaluation --------------------------------------------------------------------
int topopt_simp_eval(topopt_handle* h, topopt_filter* f, int32_t filter_kind, const double* x, int32_t penalty_kind, double p,
                     double xmin, const topopt_cg_opts* opts, double* obj, double* grad_x, topopt_cg_result* result) {
  if (!h || !opts) return fail(h, TOPOPT_ERR_INVALID, "topopt_simp_eval: NULL argument");
  if (filter_kind < 0 || filter_kind > 2 || (filter_kind != 0 && (!f || f->h != h)))
    return fail(h, TOPOPT_ERR_INVALID, "topopt_simp_eval: bad filter");
  TRY(use_device(h));
  if (x) {
    CUDA_TRY(h, cudaMemcpyAsync(h->d_design, x, sizeof(double) * h->nel, cudaMemcpyDefault, h->stream));
    h->stats.h2d_bytes += sizeof(double) * h->nel;
  }
  const double* xf = h->d_design;
  if (filter_kind == 1) {
    TRY(filter_run(f, h->d_design, h->d_xf, TOPOPT_FILTER_FORWARD));
    xf = h->d_xf;
  }
  TRY(upload_elems(h, xf, h->d_rho, false));
  TRY(penalize(h, penalty_kind, p, xmin, 1));
  TRY(cg_solve(h, h->d_b, opts, result));
  TRY(run_sens(h, h->d_u, h->d_u, -1.0, obj));
  double* gfull = nullptr;
  TRY(gather_elems_device(h, h->d_grad, &gfull));
  const double* gout = gfull;
  if (filter_kind != 0) {
    TRY(filter_run(f, gfull, h->d_gfull, filter_kind == 1 ? TOPOPT_FILTER_TRANSPOSE : TOPOPT_FILTER_FORWARD));
    gout = h->d_gfull;
  }
  h->d_lastgrad = gout;
  if (grad_x) {
    CUDA_TRY(h, cudaMemcpyAsync(grad_x, gout, sizeof(double) * h->nel, cudaMemcpyDefault, h->stream));
    h->stats.d2h_bytes += sizeof(double) * h->nel;
  }
  return sync(h);
}

int topopt_get_design(topopt_handle* h, double* x) {
  if (!h || !x) return fail(h, TOPOPT_ERR_INVALID, "topopt_get_design: NULL argument");
  TRY(use_device(h));
  CUDA_TRY(h, cudaMemcpyAsync(x, h->d_design, sizeof(double) * h->nel, cudaMemcpyDefault, h->stream));
  h->stats.d2h_bytes += sizeof(double) * h->nel;
  return sync(h);
}

// Optimality-criteria update with bisection on the volume multiplier, entirely on the device: every bisection step is
// one fused candidate + volume reduction over the resident design; only the 8-byte volume crosses PCIe.
int topopt_oc_update(topopt_handle* h, const double* x, const double* dc, const double* dv, double volfrac, double move, double eta,
                     double xlo, double* x_out, double* change_l2, int32_t* nbisect) {
  if (!h) return fail(h, TOPOPT_ERR_INVALID, "topopt_oc_update: NULL handle");
  if (!(volfrac > 0) || !(move > 0) || !(eta > 0)) return fail(h, TOPOPT_ERR_INVALID, "topopt_oc_update: volfrac, move and eta must be positive");
  TRY(use_device(h));
  const size_t bytes = sizeof(double) * h->nel;
  if (!h->d_dv) {
    TRY(dev_alloc(h, &h->d_dv, h->nel));
    TRY(dev_alloc(h, &h->d_dc, h->nel));
    TRY(dev_alloc(h, &h->d_xnew, h->nel));
    if (!dv) return fail(h, TOPOPT_ERR_INVALID, "topopt_oc_update: the first call must pass dv (volume gradient)");
  }
  if (x) {
    CUDA_TRY(h, cudaMemcpyAsync(h->d_design, x, bytes, cudaMemcpyDefault, h->stream));
    h->stats.h2d_bytes += bytes;
  }
  if (dv) {
    CUDA_TRY(h, cudaMemcpyAsync(h->d_dv, dv, bytes, cudaMemcpyDefault, h->stream));
    h->stats.h2d_bytes += bytes;
  }
  const double* g = h->d_lastgrad;
  if (dc) {
    CUDA_TRY(h, cudaMemcpyAsync(h->d_dc, dc, bytes, cudaMemcpyDefault, h->stream));
    h->stats.h2d_bytes += bytes;
    g = h->d_dc;
  }
  if (!g) return fail(h, TOPOPT_ERR_INVALID, "topopt_oc_update: no gradient (pass dc or call topopt_simp_eval first)");
  double l1 = 0.0, l2 = 1e9, lam = 0.0;
  int steps = 0;
  while ((l2 - l1) / (l1 + l2) > 1e-12 && l2 > 1e-40) {
    lam = 0.5 * (l1 + l2);
    LAUNCH(h, (k_oc<false>), kReduceBlocks, (long long)h->nel, h->d_design, g, h->d_dv, lam, move, eta, xlo, (double*)nullptr, h->d_partials, h->d_st);
    CUDA_TRY(h, cudaMemcpyAsync(h->h_st, h->d_st, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    TRY(sync(h));
    if (h->h_st->sums[0] > volfrac) {
      l1 = lam;
    } else {
      l2 = lam;
    }
    ++steps;
  }
  // the candidate of the LAST trial multiplier is the new design (the rule of the host-side loop it replaces)
  LAUNCH(h, (k_oc<true>), kReduceBlocks, (long long)h->nel, h->d_design, g, h->d_dv, lam, move, eta, xlo, h->d_xnew, h->d_partials, h->d_st);
  TRY(check_launch(h, "k_oc"));
  CUDA_TRY(h, cudaMemcpyAsync(h->h_st, h->d_st, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  std::swap(h->d_design, h->d_xnew);
  if (x_out) {
    CUDA_TRY(h, cudaMemcpyAsync(x_out, h->d_design, bytes, cudaMemcpyDefault, h->stream));
    h->stats.d2h_bytes += bytes;
  }
  TRY(sync(h));
  if (change_l2) *change_l2 = std::sqrt(h->h_st->sums[1]);
  if (nbisect) *nbisect = steps;
  return TOPOPT_OK;
}

// ---- measurement ------------------------------------------------------------------------------
#ifdef TOPOPT_TIMELINE
int topopt_debug_timeline(topopt_handle* h, unsigned long long* out) {
  TRY(use_device(h));
  CUDA_TRY(h, cudaMemcpyFromSymbol(out, g_tl, sizeof(unsigned long long) * 2 * 512 * 4));
  return TOPOPT_OK;
}
#endif

int topopt_time_kernel(topopt_handle* h, topopt_filter* f, int32_t which, int32_t reps, double* ms_out) {
  if (!h || !ms_out || reps < 1) return fail(h, TOPOPT_ERR_INVALID, "topopt_time_kernel: bad argument");
  TRY(use_device(h));
  topopt_cg_opts o{};
  o.abstol = 0;
  o.reltol = 0;
  o.maxiter = reps;
  o.check_every = reps;
  float ms = 0;
  switch (which) {
    case 0:
    case 7:
    case 8: {
      // The K.u launch of the CG loop, as the loop issues it: dense direction vector (zero on prescribed
      // dofs), fused dot product(s), no scalar step.  0 = the kernel the default CG selects, 7 = previous
      // generation (two-row / one-row shuffle kernels), 8 = ring-staged kernel with both dot products.
      CGState zero;
      std::memset(&zero, 0, sizeof(zero));
      zero.world = 1;
      CUDA_TRY(h, cudaMemcpyAsync(h->d_st, &zero, sizeof(CGState), cudaMemcpyHostToDevice, h->stream));
      LAUNCH(h, k_fill_dense, kReduceBlocks, h->off, h->nown_dofs, h->d_p);
#define CALL(D, C) LAUNCH(h, (k_apply_zero<C>), grid_for(h->nloc_nodes, kWideGrid), (long long)h->nloc_nodes, h->d_fixed, h->d_p)
      DISPATCH(h, CALL);
#undef CALL
      TRY(exchange_halo(h, h->d_p));
      const bool hex8 = h->dim == 3 && h->nc == 3 && h->modal_ok;
      const int saved_ring = h->kxu_ring;
      if (which == 7) h->kxu_ring = 0;
      auto once = [&]() -> int {
        if (which == 8 && !(hex8 && use_ring(h))) return fail(h, TOPOPT_ERR_INVALID, "topopt_time_kernel: ring kernel not applicable");
        if (hex8 && use_ring(h)) return which == 8 ? launch_hex8_ring<2, false>(h, h->d_p, h->d_Ap, FIN_NONE) : launch_hex8_ring<1, false>(h, h->d_p, h->d_Ap, FIN_NONE);
        return launch_apply<true>(h, h->d_p, h->d_Ap, FIN_NONE);
      };
      int rc = once();  // warm-up (attribute setup, instruction cache)
      CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
      for (int r = 0; r < reps && rc == TOPOPT_OK; ++r) rc = once();
      CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
      h->kxu_ring = saved_ring;
      TRY(rc);
      break;
    }
    case 1:
    case 6:
    case 9: {
      // `reps` CG iterations on a dense right-hand side (every vector is live from the first iteration)
      o.op = which == 6 ? TOPOPT_OP_ASSEMBLED : TOPOPT_OP_MATRIX_FREE;
      o.variant = which == 9 ? TOPOPT_CG_SINGLE_PASS : TOPOPT_CG_REFERENCE;
      LAUNCH(h, k_fill_dense, kReduceBlocks, h->off, h->nown_dofs, h->d_rhs);
#define CALL(D, C) LAUNCH(h, (k_apply_zero<C>), grid_for(h->nloc_nodes, kWideGrid), (long long)h->nloc_nodes, h->d_fixed, h->d_rhs)
      DISPATCH(h, CALL);
#undef CALL
      const int saved_variant = h->cg_variant_env;
      h->cg_variant_env = -1;
      topopt_cg_result r{};
      const int rc = cg_solve(h, h->d_rhs, &o, &r, true, reps);
      h->cg_variant_env = saved_variant;
      TRY(rc);
      *ms_out = r.solve_ms / std::max(1, r.iters);
      return TOPOPT_OK;
    }
    case 2:
      CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
      for (int r = 0; r < reps; ++r) TRY(run_sens(h, h->d_u, h->d_u, -1.0, nullptr));
      *ms_out = h->stats.last_sens_ms;
      return TOPOPT_OK;
    case 3:
      if (!f) return fail(h, TOPOPT_ERR_INVALID, "topopt_time_kernel: filter needed");
      CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
      for (int r = 0; r < reps; ++r) TRY(filter_run(f, h->d_design, h->d_xf, TOPOPT_FILTER_FORWARD));
      CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
      break;
    case 4:
      if (!h->assembled || h->stiffness_dirty) TRY(do_assemble(h));
      CUDA_TRY(h, cudaMemcpyAsync(h->d_p, h->d_b, sizeof(double) * h->nloc_dofs, cudaMemcpyDeviceToDevice, h->stream));
      CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
      for (int r = 0; r < reps; ++r) TRY(launch_spmv<false>(h, h->d_p, h->d_Ap, FIN_NONE));
      CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
      break;
    case 5:
      TRY(do_assemble(h));
      CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
      for (int r = 0; r < reps; ++r) TRY(do_assemble(h));
      CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
      break;
    default:
      return fail(h, TOPOPT_ERR_INVALID, "topopt_time_kernel: unknown kernel class");
  }
  TRY(check_launch(h, "topopt_time_kernel"));
  TRY(sync(h));
  CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  *ms_out = ms / reps;
  return TOPOPT_OK;
}

}  // extern "C"
