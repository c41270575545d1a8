// common.cuh -- shared device helpers for libspconv (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/spconv.h"

namespace spc {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);   // bookkeeping for spc_launch_count()

#define SPC_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      spc::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr,                \
                     cudaGetErrorString(_e));                                             \
      return SPC_ECUDA;                                                                   \
    }                                                                                     \
  } while (0)

#define SPC_REQUIRE(cond, ...)                                                            \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      spc::set_error(__VA_ARGS__);                                                        \
      return SPC_EINVAL;                                                                  \
    }                                                                                     \
  } while (0)

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// A tile plus its received halo strips, addressed in UNPADDED tile coordinates: h in
// [-hh, H+hh), w in [-hw, W+hw).  Anything not covered by the tile or a non-NULL strip reads
// as 0 -- this is ZeroPad2d + copy_halo_exchange_values (reference spatial.py:1020,405-413)
// without ever materialising the padded tensor.
struct TileView {
  const void* x;
  const void* strip[9];
  int N, C, H, W;
  int hh, hw;  // halo strip extents
};

template <typename T>
__device__ __forceinline__ float tile_load(const TileView& v, int n, int c, int h, int w) {
  const bool hin = (unsigned)h < (unsigned)v.H;
  const bool win = (unsigned)w < (unsigned)v.W;
  if (hin && win) {
    if (v.x == nullptr) return 0.f;  // halo-only view (used for linear boundary corrections)
    return to_f32<T>(reinterpret_cast<const T*>(v.x)[(((size_t)n * v.C + c) * v.H + h) * v.W + w]);
  }
  const int dr = h < 0 ? 0 : (hin ? 1 : 2);
  const int dc = w < 0 ? 0 : (win ? 1 : 2);
  const T* s = reinterpret_cast<const T*>(v.strip[dr * 3 + dc]);
  if (s == nullptr) return 0.f;
  const int sh = (dr == 1) ? v.H : v.hh;
  const int sw = (dc == 1) ? v.W : v.hw;
  const int y = (dr == 0) ? h + v.hh : (dr == 2 ? h - v.H : h);
  const int xx = (dc == 0) ? w + v.hw : (dc == 2 ? w - v.W : w);
  if ((unsigned)y >= (unsigned)sh || (unsigned)xx >= (unsigned)sw) return 0.f;
  return to_f32<T>(s[(((size_t)n * v.C + c) * sh + y) * sw + xx]);
}

// Address of element (n,c,h,w) of the tile-or-halo view, or nullptr where the view reads as zero.
// Splitting the address computation from the load lets callers batch several independent loads.
template <typename T>
__device__ __forceinline__ const T* tile_ptr(const TileView& v, int n, int c, int h, int w) {
  const bool hin = (unsigned)h < (unsigned)v.H;
  const bool win = (unsigned)w < (unsigned)v.W;
  if (hin && win) {
    if (v.x == nullptr) return nullptr;
    return reinterpret_cast<const T*>(v.x) + (((size_t)n * v.C + c) * v.H + h) * v.W + w;
  }
  const int dr = h < 0 ? 0 : (hin ? 1 : 2);
  const int dc = w < 0 ? 0 : (win ? 1 : 2);
  const T* s = reinterpret_cast<const T*>(v.strip[dr * 3 + dc]);
  if (s == nullptr) return nullptr;
  const int sh = (dr == 1) ? v.H : v.hh;
  const int sw = (dc == 1) ? v.W : v.hw;
  const int y = (dr == 0) ? h + v.hh : (dr == 2 ? h - v.H : h);
  const int xx = (dc == 0) ? w + v.hw : (dc == 2 ? w - v.W : w);
  if ((unsigned)y >= (unsigned)sh || (unsigned)xx >= (unsigned)sw) return nullptr;
  return s + (((size_t)n * v.C + c) * sh + y) * sw + xx;
}

inline TileView make_view(const void* x, const spc_halo* halo, int N, int C, int H, int W, int hh, int hw) {
  TileView v;
  v.x = x;
  for (int i = 0; i < 9; ++i) v.strip[i] = halo ? halo->strip[i] : nullptr;
  v.strip[4] = nullptr;
  v.N = N; v.C = C; v.H = H; v.W = W; v.hh = hh; v.hw = hw;
  return v;
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t dtype_size(int dt) { return dt == SPC_BF16 ? 2 : 4; }

// ---- direct (CUDA-core) kernels: conv_direct.cu ----------------------------------------
// Generalised correlation: y[n,k, oy0+i*oys, ox0+j*oxs] = bias[k] + sum_{c,r,s}
//   w[w_off + k*wKs + c*wCs + r*wRs + s*wSs] * in(n, c, i*sh + r - pt, j*sw + s - pl)
// The affine weight view lets fprop, dgrad (transposed + flipped filter) and the stride-2
// dgrad parity classes share one kernel without repacking weights.
struct DirectConvParams {
  TileView in;
  const void* w;
  const void* bias;
  void* y;
  int K, R, S, sh, sw, pt, pl;
  int Ho, Wo;                       // logical output extent of this launch
  int YH, YW, oy0, ox0, oys, oxs;   // physical output tensor and mapping
  long long w_off, wKs, wCs, wRs, wSs;
  int vert;                         // 1: CTA tile is 128 rows x 8 cols (thin column strips) instead of 8 x 128
};
int launch_conv_direct(const DirectConvParams& p, int dtype, cudaStream_t st);

struct DirectWgradParams {
  TileView in;        // x + halos
  const void* dy;     // [N][K][Ho][Wo]
  float* dw;          // [K][C][R][S] fp32, accumulated with atomics (zeroed by caller)
  int K, R, S, sh, sw, ph, pw, Ho, Wo;
  int ry0, rx0, rH, rW;   // output sub-rectangle to reduce over (rH == 0: the whole output)
};
int launch_wgrad_direct(const DirectWgradParams& p, int dtype, cudaStream_t st);
// dw += contribution of the halo pixels only (boundary output pixels x taps that fall outside the tile)
int launch_wgrad_halo(const DirectWgradParams& p, int dtype, cudaStream_t st);
int launch_bias_grad(const void* dy, float* db, int N, int K, int HW, int dtype, int accumulate, cudaStream_t st);

// ---- boundary patches (halo.cu) ---------------------------------------------------------------
void* boundary_scratch(size_t bytes);
int launch_patch_gather(const TileView& v, void* P, int Hp, int Wp, int h0, int w0, int dtype, cudaStream_t st);
int launch_patch_gather_dy(const void* dy, void* G, int NK, int Ho, int Wo, int Hp, int Wp, int y0, int x0, int rh, int rw,
                           int ph, int pw, int dtype, cudaStream_t st);
int launch_patch_scatter(const void* O, void* y, int NK, int Ho, int Wo, int Hp, int Wp, int y0, int x0, int rh, int rw,
                         int ph, int pw, int dtype, cudaStream_t st);

// ---- halo fix-up as a small GEMM over the boundary outputs only (halo.cu kernels, api.cu orchestration) ----------
// The boundary outputs of a tile (those whose window reaches a received strip) are listed as <= 4 disjoint output
// rectangles; P_b = their pixel count (x N), padded to a multiple of 64.
struct BoundaryRects {
  int n;                       // rectangles in use
  int y0[4], y1[4], x0[4], x1[4];
  int start[5];                // prefix sums of the pixel counts PER IMAGE
  int N, per_image, total;     // images, pixels per image, N * per_image
  int padded;                  // total rounded up to 64
};
// V[(c,r,s)][p] = the pixel the tap (r,s) of boundary output p reads in channel c, through `view`: tile + strips for
// fprop (full windows), the halo-only view for wgrad (0 where the tap stays inside the tile); bf16
int launch_halo_im2col(const TileView& halo_only, const BoundaryRects& b, int R, int S, int sh, int sw, int ph, int pw, void* V,
                       cudaStream_t st);
// G[k][p] = dy[n][k][i][j] at the boundary outputs;  y[n][k][i][j] = O[k][p]
int launch_boundary_gather(const void* dy, const BoundaryRects& b, int K, int Ho, int Wo, void* G, cudaStream_t st);
int launch_boundary_scatter(const void* O, const BoundaryRects& b, int K, int Ho, int Wo, void* y, cudaStream_t st);
size_t tc_pw_workspace_bytes(int M, int Cin);
int tc_pw_fwd(const void* w, int ld, int M, int Cin, const void* x, const void* bias, void* y, int P, void* ws, size_t ws_bytes,
              cudaStream_t st);
int tc_pw_wgrad(const void* x, const void* dy, float* dw, int K, int C, int P, cudaStream_t st);

// ---- tcgen05 pointwise GEMM path: gemm_tc.cu ---------------------------------------------
bool tc_supported(const spc_conv_desc* d, int op);
size_t tc_workspace_bytes(const spc_conv_desc* d, int op);
int tc_conv_fwd(const spc_conv_desc* d, const void* x, const void* w, const void* bias, void* y,
                void* ws, size_t ws_bytes, cudaStream_t st);
int tc_conv_dgrad(const spc_conv_desc* d, const void* dy, const void* w, void* dx, void* ws,
                  size_t ws_bytes, cudaStream_t st);
int tc_conv_wgrad(const spc_conv_desc* d, const void* x, const void* dy, float* dw, int accumulate,
                  void* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace spc
