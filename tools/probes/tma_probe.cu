// Stand-alone check of the tensor-map staging used by kxu_hex8_cgtma.cuh: even/odd row maps over a [R][3 NX] fp64 array
// with an odd NX, boxes at negative coordinates, map passed by global pointer or as a __grid_constant__ parameter.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>
typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <bool PARAM>
__global__ void k(const __grid_constant__ CUtensorMap pm, const CUtensorMap* gm, int x, int y, double* out, int n) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const CUtensorMap* m = PARAM ? &pm : gm;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(n * 8) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s32(sm)),
                 "l"(m), "r"(x), "r"(y), "r"(s32(&bar))
                 : "memory");
  }
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(&bar)), "r"(0) : "memory");
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = reinterpret_cast<double*>(sm)[i];
}
int run(int NX, CUtensorMapDataType dt, CUtensorMapL2promotion l2, int BX, int BR, int X0, int Y0) {
  const int NY = 9, planes = 5, rows = planes * NY, rd = NX * 3;
  const int esz = dt == CU_TENSOR_MAP_DATA_TYPE_FLOAT32 ? 4 : 8;
  printf("---- NX %d dtype %d l2 %d box %d x %d at %d %d\n", NX, (int)dt, (int)l2, BX, BR, X0, Y0);
  std::vector<double> h((size_t)rows * rd);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (double)i;
  double *d, *out;
  cudaMalloc(&d, h.size() * 8 + 256);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  cudaMalloc(&out, BX * BR * 8);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  PFN enc = (PFN)fn;
  CUtensorMap maps[2];
  for (int par = 0; par < 2; ++par) {
    const cuuint64_t gdim[2] = {(cuuint64_t)((rd + par) * 8 / esz), (cuuint64_t)(par ? rows / 2 : (rows + 1) / 2)};
    const cuuint64_t gs[1] = {(cuuint64_t)(2 * rd * 8)};
    const cuuint32_t box[2] = {(cuuint32_t)BX, (cuuint32_t)BR}, es[2] = {1, 1};
    CUresult r = enc(&maps[par], dt, 2, par ? (void*)(d + rd - 1) : (void*)d, gdim, gs, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d -> %d\n", par, (int)r);
  }
  CUtensorMap* dm;
  cudaMalloc(&dm, sizeof(maps));
  cudaMemcpy(dm, maps, sizeof(maps), cudaMemcpyHostToDevice);
  std::vector<double> o(BX * BR);
  for (int variant = 0; variant < 2; ++variant)
    for (int par = 0; par < 2; ++par) {
      const int x = X0, y = Y0;  // the same (even) x for both maps; the odd map's columns are shifted by one
      if (variant == 0) k<true><<<1, 128, BX * BR * 8>>>(maps[par], dm + par, x, y, out, BX * BR * esz / 8);
      else k<false><<<1, 128, BX * BR * 8>>>(maps[par], dm + par, x, y, out, BX * BR * esz / 8);
      cudaError_t e = cudaDeviceSynchronize();
      printf("variant %s map %d: %s\n", variant == 0 ? "param" : "global", par, cudaGetErrorString(e));
      if (e != cudaSuccess) return 1;
      cudaMemcpy(o.data(), out, o.size() * 8, cudaMemcpyDeviceToHost);
      // expected: box row j = array row R = 2 (y + j) + par, columns -3 .. 96 (zero outside [0, rd))
      int bad = 0;
      for (int j = 0; j < BR; ++j)
        for (int c = 0; c < BX && esz == 8; ++c) {
          const int col = c + x - par, R = 2 * (y + j) + par;
          const double want = (col >= 0 && col < rd && R < rows) ? (double)((size_t)R * rd + col) : 0.0;
          if (o[j * BX + c] != want && bad++ < 5) printf("  mismatch row %d col %d: got %g want %g\n", j, c, o[j * BX + c], want);
        }
      printf("  mismatches: %d\n", bad);
    }
  return 0;
}
int main(int argc, char** argv) {
  const int NX = argc > 1 ? atoi(argv[1]) : 13;
  const int dt = argc > 2 ? atoi(argv[2]) : (int)CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
  const int l2 = argc > 3 ? atoi(argv[3]) : (int)CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  const int BX = argc > 4 ? atoi(argv[4]) : 100, BR = argc > 5 ? atoi(argv[5]) : 6, X0 = argc > 6 ? atoi(argv[6]) : -3, Y0 = argc > 7 ? atoi(argv[7]) : 2;
  return run(NX, (CUtensorMapDataType)dt, (CUtensorMapL2promotion)l2, BX, BR, X0, Y0);
}
