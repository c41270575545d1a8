// Dev probe: which 4-D / strided TMA tile loads does B200 accept?  nvcc -arch=sm_100a tma_probe.cu -o tma_probe -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include "../mpi4dl_b200/csrc/tc_common.cuh"
using namespace spc::tc;

__global__ void probe4(const __grid_constant__ CUtensorMap m, int c0, int c1, int c2, int c3, __nv_bfloat16* out, int n) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 16384);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) { mbar_arrive_expect_tx(bar, n * 2); tma_load_4d(sm, &m, bar, c0, c1, c2, c3); }
  mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = reinterpret_cast<__nv_bfloat16*>(sm)[i];
}

typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  int only = argc > 1 ? atoi(argv[1]) : -1;
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  Enc enc = (Enc)fp;
  const int W = 128, H = 8, C = 64, N = 1;
  size_t n = (size_t)W * H * C * N;
  __nv_bfloat16* h = (__nv_bfloat16*)malloc(n * 2);
  for (size_t i = 0; i < n; ++i) h[i] = __float2bfloat16((float)(i % 251));
  __nv_bfloat16 *d, *o;
  cudaMalloc(&d, n * 2); cudaMalloc(&o, 65536);
  cudaMemcpy(d, h, n * 2, cudaMemcpyHostToDevice);
  struct Cfg { const char* name; cuuint32_t box[4]; cuuint32_t es[4]; int c[4]; CUtensorMapSwizzle sw; int Cdim; };
  Cfg cfgs[] = {
      {"4d box{64,1,64,1} sw128 coords 0", {64, 1, 64, 1}, {1, 1, 1, 1}, {0, 0, 0, 0}, CU_TENSOR_MAP_SWIZZLE_128B, 64},
      {"4d box{64,1,64,1} sw128 coords (-3,2,0,0)", {64, 1, 64, 1}, {1, 1, 1, 1}, {-3, 2, 0, 0}, CU_TENSOR_MAP_SWIZZLE_128B, 64},
      {"4d box{64,1,64,1} sw128 coords (3,2,0,0)", {64, 1, 64, 1}, {1, 1, 1, 1}, {3, 2, 0, 0}, CU_TENSOR_MAP_SWIZZLE_128B, 64},
      {"4d box{64,1,64,1} sw128 coords (-8,-1,0,0)", {64, 1, 64, 1}, {1, 1, 1, 1}, {-8, -1, 0, 0}, CU_TENSOR_MAP_SWIZZLE_128B, 64},
      {"4d box{64,1,64,1} sw128 coords (8,2,0,0)", {64, 1, 64, 1}, {1, 1, 1, 1}, {8, 2, 0, 0}, CU_TENSOR_MAP_SWIZZLE_128B, 64},
      {"4d box{64,1,64,1} NONE coords (3,2,0,0)", {64, 1, 64, 1}, {1, 1, 1, 1}, {3, 2, 0, 0}, CU_TENSOR_MAP_SWIZZLE_NONE, 64},
      {"4d box{64,1,64,1} NONE coords (-3,2,0,0)", {64, 1, 64, 1}, {1, 1, 1, 1}, {-3, 2, 0, 0}, CU_TENSOR_MAP_SWIZZLE_NONE, 64},
      {"4d box{64,1,64,1} sw128 C=52", {64, 1, 64, 1}, {1, 1, 1, 1}, {0, -1, 0, 0}, CU_TENSOR_MAP_SWIZZLE_128B, 52},
      {"4d box{64,1,64,1} none", {64, 1, 64, 1}, {1, 1, 1, 1}, {0, 0, 0, 0}, CU_TENSOR_MAP_SWIZZLE_NONE, 64},
      {"4d box{128,1,64,1} estride{2,1,1,1} sw128", {128, 1, 64, 1}, {2, 1, 1, 1}, {0, 0, 0, 0}, CU_TENSOR_MAP_SWIZZLE_128B, 64},
      {"4d box{128,1,64,1} estride{2,1,1,1} sw128 coord 1", {128, 1, 64, 1}, {2, 1, 1, 1}, {1, 2, 0, 0}, CU_TENSOR_MAP_SWIZZLE_128B, 64},
      {"4d box{64,1,64,1} estride{2,1,1,1} sw64", {64, 1, 64, 1}, {2, 1, 1, 1}, {0, 0, 0, 0}, CU_TENSOR_MAP_SWIZZLE_64B, 64},
  };
  int idx = -1;
  for (auto& c : cfgs) {
    ++idx;
    if (only >= 0 && idx != only) continue;
    CUtensorMap m;
    cuuint64_t gd[4] = {W, H, (cuuint64_t)c.Cdim, N};
    cuuint64_t gs[3] = {W * 2, (cuuint64_t)W * H * 2, (cuuint64_t)W * H * C * 2};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, gd, gs, c.box, c.es, CU_TENSOR_MAP_INTERLEAVE_NONE, c.sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%-55s encode FAILED %d\n", c.name, (int)r); continue; }
    int nel = 1;
    for (int i = 0; i < 4; ++i) nel *= (c.box[i] + c.es[i] - 1) / c.es[i];
    cudaMemset(o, 0, 65536);
    probe4<<<1, 128, 16384 + 64>>>(m, c.c[0], c.c[1], c.c[2], c.c[3], o, nel);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-55s launch FAILED: %s\n", c.name, cudaGetErrorString(e)); return 1; }
    __nv_bfloat16 res[8];
    cudaMemcpy(res, o, 16, cudaMemcpyDeviceToHost);
    // expected first elements (row c1, col c0..): value = index % 251 (no swizzle effect on the first 16B chunk of row 0)
    printf("%-55s ok nel=%d first:", c.name, nel);
    for (int i = 0; i < 8; ++i) printf(" %g", __bfloat162float(res[i]));
    printf("\n");
  }
  return 0;
}
