// api.cu -- C-ABI entry points for convolution (include/spconv.h) and algorithm dispatch.
//
// Dispatch: shapes the tcgen05 implicit-GEMM path supports (gemm_tc.cu) run there over the
// whole tile with zero padding; if the tile has neighbours, the thin output strips whose
// receptive field reaches into a halo are then recomputed by the direct kernel, which reads the
// received strips in place.  Interior compute therefore never waits on the halo exchange --
// the overlap the reference left as dead code (spatial.py:415-866).  Everything else runs
// entirely on the direct kernel.
#include <stdlib.h>

#include "common.cuh"

namespace spc {
namespace {

int validate(const spc_conv_desc* d) {
  SPC_REQUIRE(d != nullptr, "conv: null descriptor");
  SPC_REQUIRE(d->N >= 0 && d->C > 0 && d->K > 0 && d->H > 0 && d->W > 0, "conv: bad shape N=%d C=%d K=%d H=%d W=%d",
              d->N, d->C, d->K, d->H, d->W);
  SPC_REQUIRE(d->R >= 1 && d->S >= 1 && d->stride_h >= 1 && d->stride_w >= 1, "conv: bad filter/stride");
  // reference spatial.py:115-121: halo_len = (k-1)/2 must equal the conv padding ("same")
  SPC_REQUIRE(d->pad_h == (d->R - 1) / 2 && d->pad_w == (d->S - 1) / 2,
              "conv: Spatial not supported yet for this configuration (pad (%d,%d) != ((R-1)/2,(S-1)/2) for %dx%d)",
              d->pad_h, d->pad_w, d->R, d->S);
  SPC_REQUIRE(d->dtype == SPC_F32 || d->dtype == SPC_BF16, "conv: bad dtype %d", d->dtype);
  SPC_REQUIRE(d->pad_h <= d->H && d->pad_w <= d->W, "conv: halo larger than the tile");
  return SPC_OK;
}

bool has_halo(const spc_halo* h) {
  if (!h) return false;
  for (int i = 0; i < 9; ++i)
    if (i != 4 && h->strip[i]) return true;
  return false;
}

DirectConvParams fwd_params(const spc_conv_desc* d, const void* x, const spc_halo* halo, const void* w,
                            const void* bias, void* y) {
  DirectConvParams p{};
  p.in = make_view(x, halo, d->N, d->C, d->H, d->W, d->pad_h, d->pad_w);
  p.w = w; p.bias = bias; p.y = y;
  p.K = d->K; p.R = d->R; p.S = d->S; p.sh = d->stride_h; p.sw = d->stride_w;
  p.pt = d->pad_h; p.pl = d->pad_w;
  spc_conv_out_shape(d, &p.Ho, &p.Wo);
  p.YH = p.Ho; p.YW = p.Wo; p.oy0 = 0; p.ox0 = 0; p.oys = 1; p.oxs = 1;
  p.w_off = 0; p.wKs = (long long)d->C * d->R * d->S; p.wCs = (long long)d->R * d->S; p.wRs = d->S; p.wSs = 1;
  return p;
}

inline size_t al256(size_t b) { return (b + 255) & ~(size_t)255; }

// Boundary rect through the tcgen05 path: gather a 64-column-aligned patch of tile+halo around the
// rect, run the SAME fast convolution on that small image, scatter the rect back.  Stride 1 only.
static bool patch_ok(const spc_conv_desc* d, int op) {
  if (d->dtype != SPC_BF16 || d->stride_h != 1 || d->stride_w != 1 || d->algo == SPC_ALGO_DIRECT) return false;
  spc_conv_desc q = *d;
  q.H = 64; q.W = 64; q.N = d->N;
  return tc_supported(&q, op);
}

static int patch_fwd_rect(const spc_conv_desc* d, const void* x, const spc_halo* halo, const void* w, const void* bias,
                          void* y, int y0, int y1, int x0, int x1, cudaStream_t st) {
  if (y1 <= y0 || x1 <= x0) return SPC_OK;
  const int rh = y1 - y0, rw = x1 - x0, ph = d->pad_h, pw = d->pad_w;
  spc_conv_desc q = *d;
  q.H = rh + 2 * ph;
  q.W = ((rw + 2 * pw) + 63) & ~63;
  q.algo = SPC_ALGO_TCGEN05;
  const size_t esz = dtype_size(d->dtype);
  const size_t pbytes = al256((size_t)d->N * d->C * q.H * q.W * esz), obytes = al256((size_t)d->N * d->K * q.H * q.W * esz);
  const size_t wsb = tc_workspace_bytes(&q, 0);
  char* base = (char*)boundary_scratch(pbytes + obytes + wsb + 1024);
  SPC_REQUIRE(base != nullptr, "boundary scratch allocation failed");
  void* P = base; void* O = base + pbytes; void* ws = base + pbytes + obytes;
  TileView v = make_view(x, halo, d->N, d->C, d->H, d->W, ph, pw);
  int rc = launch_patch_gather(v, P, q.H, q.W, y0 - ph, x0 - pw, d->dtype, st);
  if (rc) return rc;
  rc = tc_conv_fwd(&q, P, w, bias, O, ws, wsb, st);
  if (rc) return rc;
  int Ho, Wo;
  spc_conv_out_shape(d, &Ho, &Wo);
  return launch_patch_scatter(O, y, d->N * d->K, Ho, Wo, q.H, q.W, y0, x0, rh, rw, ph, pw, d->dtype, st);
}

// dw += (halo pixels only) x (dy restricted to the rect), through the tcgen05 wgrad on a patch
static int patch_wgrad_rect(const spc_conv_desc* d, const spc_halo* halo, const void* dy, float* dw, int y0, int y1,
                            int x0, int x1, cudaStream_t st) {
  if (y1 <= y0 || x1 <= x0) return SPC_OK;
  const int rh = y1 - y0, rw = x1 - x0, ph = d->pad_h, pw = d->pad_w;
  spc_conv_desc q = *d;
  q.H = rh + 2 * ph;
  q.W = ((rw + 2 * pw) + 63) & ~63;
  q.algo = SPC_ALGO_TCGEN05;
  const size_t esz = dtype_size(d->dtype);
  const size_t pbytes = al256((size_t)d->N * d->C * q.H * q.W * esz), gbytes = al256((size_t)d->N * d->K * q.H * q.W * esz);
  const size_t wsb = tc_workspace_bytes(&q, 2);
  char* base = (char*)boundary_scratch(pbytes + gbytes + wsb + 1024);
  SPC_REQUIRE(base != nullptr, "boundary scratch allocation failed");
  void* P = base; void* G = base + pbytes; void* ws = base + pbytes + gbytes;
  TileView v = make_view(nullptr, halo, d->N, d->C, d->H, d->W, ph, pw);   // halo pixels only
  int rc = launch_patch_gather(v, P, q.H, q.W, y0 - ph, x0 - pw, d->dtype, st);
  if (rc) return rc;
  int Ho, Wo;
  spc_conv_out_shape(d, &Ho, &Wo);
  rc = launch_patch_gather_dy(dy, G, d->N * d->K, Ho, Wo, q.H, q.W, y0, x0, rh, rw, ph, pw, d->dtype, st);
  if (rc) return rc;
  return tc_conv_wgrad(&q, P, G, dw, 1, ws, wsb, st);
}

// ---- halo fix-up of the tcgen05 paths: a small GEMM over the boundary outputs only ------------------------------
// After the interior pass ran the whole tile with zero padding, only the outputs whose window reaches a received
// strip are wrong (P_b of them: a few rows / columns).  fprop: V[(c,r,s)][p] = im2col of tile + strips over those
// P_b outputs, then ONE pointwise GEMM  O[K][P_b] = w[K][C*R*S] * V (+bias)  on the tcgen05 kernel -- the filter tensor
// IS that matrix -- and a scatter that overwrites them.  wgrad: the halo pixels' share is linear, so with V taken
// from the HALO-ONLY view  dW[K][C*R*S] += dY_b[K][P_b] * V^T  (pw_wgrad_kernel accumulating straight into dw).
// Round 1 gathered a 64-column-aligned patch around every boundary rectangle and re-ran the convolution on it
// (3 launches per rectangle, up to 4 rectangles; strided layers fell to the direct kernel on thin strips): +0.18 ms
// per 1x7 fprop and +0.59 ms per wgrad on the 1024x128 tiles of an 8-GPU run (profiles/r2_halo_cost_n8.txt) -- more
// than the interior pass itself; now +0.07 / +0.08 ms (profiles/r2_halo_cost_n8_v2.txt).
static bool boundary_rects(const spc_conv_desc* d, const spc_halo* halo, int Ho, int Wo, BoundaryRects* b) {
  const int top = min(Ho, ceil_div(d->pad_h, d->stride_h));
  int bot0 = ceil_div(d->H + d->pad_h - d->R + 1, d->stride_h);      // first output row touching the bottom halo
  bot0 = max(top, min(Ho, bot0));
  const int left = min(Wo, ceil_div(d->pad_w, d->stride_w));
  int right0 = ceil_div(d->W + d->pad_w - d->S + 1, d->stride_w);
  right0 = max(left, min(Wo, right0));
  const bool any_top = halo->strip[0] || halo->strip[1] || halo->strip[2];
  const bool any_bot = halo->strip[6] || halo->strip[7] || halo->strip[8];
  const bool any_left = halo->strip[0] || halo->strip[3] || halo->strip[6];
  const bool any_right = halo->strip[2] || halo->strip[5] || halo->strip[8];
  const int sy0 = any_top ? top : 0, sy1 = any_bot ? bot0 : Ho;       // rows not already covered by the bands
  int n = 0, acc = 0;
  auto add = [&](int y0, int y1, int x0, int x1) {
    if (y1 <= y0 || x1 <= x0) return;
    b->y0[n] = y0; b->y1[n] = y1; b->x0[n] = x0; b->x1[n] = x1;
    b->start[n] = acc;
    acc += (y1 - y0) * (x1 - x0);
    ++n;
  };
  if (any_top) add(0, top, 0, Wo);
  if (any_bot) add(bot0, Ho, 0, Wo);
  if (any_left) add(sy0, sy1, 0, left);
  if (any_right) add(sy0, sy1, right0, Wo);
  for (int i = n; i < 5; ++i) b->start[i] = acc;
  for (int i = n; i < 4; ++i) { b->y0[i] = b->y1[i] = b->x0[i] = 0; b->x1[i] = 1; }
  b->n = n; b->N = d->N; b->per_image = acc; b->total = acc * d->N;
  b->padded = (b->total + 63) & ~63;
  return b->total > 0;
}

static int boundary_fwd_tc(const spc_conv_desc* d, const void* x, const spc_halo* halo, const void* w, const void* bias,
                           void* y, cudaStream_t st) {
  int Ho, Wo;
  spc_conv_out_shape(d, &Ho, &Wo);
  BoundaryRects b;
  if (!boundary_rects(d, halo, Ho, Wo, &b)) return SPC_OK;
  const int CT = d->C * d->R * d->S;
  const size_t vbytes = al256((size_t)CT * b.padded * 2), obytes = al256((size_t)d->K * b.padded * 2);
  const size_t wsb = tc_pw_workspace_bytes(d->K, CT);
  char* base = (char*)boundary_scratch(vbytes + obytes + wsb + 2048);
  SPC_REQUIRE(base != nullptr, "boundary scratch allocation failed");
  void* V = base; void* O = base + vbytes; void* ws = base + vbytes + obytes;
  // fprop gathers the FULL windows (tile + strips) of the boundary outputs and overwrites them: the result is rounded
  // to bf16 once, like every other output (adding a halo-only correction to the already rounded interior value
  // rounds twice -- 335 of 4.4e8 stem outputs left the parity tolerance that way)
  TileView v = make_view(x, halo, d->N, d->C, d->H, d->W, d->pad_h, d->pad_w);
  int rc = launch_halo_im2col(v, b, d->R, d->S, d->stride_h, d->stride_w, d->pad_h, d->pad_w, V, st);
  if (rc) return rc;
  rc = tc_pw_fwd(w, CT, d->K, CT, V, bias, O, b.padded, ws, wsb, st);
  if (rc) return rc;
  return launch_boundary_scatter(O, b, d->K, Ho, Wo, y, st);
}

static int boundary_wgrad_tc(const spc_conv_desc* d, const spc_halo* halo, const void* dy, float* dw, cudaStream_t st) {
  int Ho, Wo;
  spc_conv_out_shape(d, &Ho, &Wo);
  BoundaryRects b;
  if (!boundary_rects(d, halo, Ho, Wo, &b)) return SPC_OK;
  const int CT = d->C * d->R * d->S;
  const size_t vbytes = al256((size_t)CT * b.padded * 2), gbytes = al256((size_t)d->K * b.padded * 2);
  char* base = (char*)boundary_scratch(vbytes + gbytes + 2048);
  SPC_REQUIRE(base != nullptr, "boundary scratch allocation failed");
  void* V = base; void* G = base + vbytes;
  TileView v = make_view(nullptr, halo, d->N, d->C, d->H, d->W, d->pad_h, d->pad_w);
  int rc = launch_halo_im2col(v, b, d->R, d->S, d->stride_h, d->stride_w, d->pad_h, d->pad_w, V, st);
  if (rc) return rc;
  rc = launch_boundary_gather(dy, b, d->K, Ho, Wo, G, st);
  if (rc) return rc;
  return tc_pw_wgrad(V, G, dw, d->K, CT, b.padded, st);   // dw[k][(c,r,s)] += sum_p G[k][p] * V[(c,r,s)][p]
}

// Launch the direct kernel on the output sub-rectangle [y0,y1) x [x0,x1).
int fwd_rect(DirectConvParams p, int dtype, int y0, int y1, int x0, int x1, cudaStream_t st) {
  if (y1 <= y0 || x1 <= x0) return SPC_OK;
  p.oy0 = y0; p.ox0 = x0;
  p.pt -= y0 * p.sh; p.pl -= x0 * p.sw;
  p.Ho = y1 - y0; p.Wo = x1 - x0;
  p.vert = (p.Wo <= 16 && p.Ho > p.Wo) ? 1 : 0;   // thin column strip: 128-row x 8-col CTA tiles
  return launch_conv_direct(p, dtype, st);
}

}  // namespace
}  // namespace spc

using namespace spc;

extern "C" {

void spc_conv_out_shape(const spc_conv_desc* d, int* Ho, int* Wo) {
  if (Ho) *Ho = (d->H + 2 * d->pad_h - d->R) / d->stride_h + 1;
  if (Wo) *Wo = (d->W + 2 * d->pad_w - d->S) / d->stride_w + 1;
}

int spc_conv_uses_tcgen05(const spc_conv_desc* d, int op) {
  if (!d || d->algo == SPC_ALGO_DIRECT) return 0;
  return tc_supported(d, op) ? 1 : 0;
}

size_t spc_conv_workspace_bytes(const spc_conv_desc* d, int op) {
  if (!d) return 0;
  return spc_conv_uses_tcgen05(d, op) ? tc_workspace_bytes(d, op) : 0;
}

// Boundary strips: output rows / columns whose window reaches outside the tile, recomputed from
// tile + received halo strips (direct kernel).  Valid after ANY interior pass that used zero padding.
static int fwd_boundary(const spc_conv_desc* d, DirectConvParams p, const spc_halo* halo, cudaStream_t st) {
  if (!has_halo(halo)) return SPC_OK;
  int rc;
  const int Ho = p.Ho, Wo = p.Wo;
  const int top = min(Ho, ceil_div(d->pad_h, d->stride_h));
  int bot0 = ceil_div(d->H + d->pad_h - d->R + 1, d->stride_h);  // first row touching the bottom halo
  bot0 = max(top, min(Ho, bot0));
  const int left = min(Wo, ceil_div(d->pad_w, d->stride_w));
  int right0 = ceil_div(d->W + d->pad_w - d->S + 1, d->stride_w);
  right0 = max(left, min(Wo, right0));
  const bool any_top = halo->strip[0] || halo->strip[1] || halo->strip[2];
  const bool any_bot = halo->strip[6] || halo->strip[7] || halo->strip[8];
  const bool any_left = halo->strip[0] || halo->strip[3] || halo->strip[6];
  const bool any_right = halo->strip[2] || halo->strip[5] || halo->strip[8];
  const int sy0 = any_top ? top : 0, sy1 = any_bot ? bot0 : Ho;   // rows not already redone by the bands
  if (d->dtype == SPC_BF16 && d->algo != SPC_ALGO_DIRECT && !getenv("SPC_BOUNDARY_V1"))
    return boundary_fwd_tc(d, p.in.x, halo, p.w, p.bias, p.y, st);
  if (patch_ok(d, 0)) {
    const void* x = p.in.x; const void* w = p.w; const void* bias = p.bias; void* y = p.y;
    if (any_top && (rc = patch_fwd_rect(d, x, halo, w, bias, y, 0, top, 0, Wo, st))) return rc;
    if (any_bot && (rc = patch_fwd_rect(d, x, halo, w, bias, y, bot0, Ho, 0, Wo, st))) return rc;
    if (any_left && (rc = patch_fwd_rect(d, x, halo, w, bias, y, sy0, sy1, 0, left, st))) return rc;
    if (any_right && (rc = patch_fwd_rect(d, x, halo, w, bias, y, sy0, sy1, right0, Wo, st))) return rc;
    return SPC_OK;
  }
  if (any_top && (rc = fwd_rect(p, d->dtype, 0, top, 0, Wo, st))) return rc;
  if (any_bot && (rc = fwd_rect(p, d->dtype, bot0, Ho, 0, Wo, st))) return rc;
  if (any_left && (rc = fwd_rect(p, d->dtype, sy0, sy1, 0, left, st))) return rc;
  if (any_right && (rc = fwd_rect(p, d->dtype, sy0, sy1, right0, Wo, st))) return rc;
  return SPC_OK;
}

static int fwd_interior(const spc_conv_desc* d, const void* x, const void* w, const void* bias, void* y, void* workspace,
                        size_t workspace_bytes, cudaStream_t st) {
  const bool tc = spc_conv_uses_tcgen05(d, 0);
  if (d->algo == SPC_ALGO_TCGEN05 && !tc) {
    set_error("conv_fwd: SPC_ALGO_TCGEN05 requested but the shape is not supported by the tcgen05 path");
    return SPC_EUNSUPPORTED;
  }
  if (tc) return tc_conv_fwd(d, x, w, bias, y, workspace, workspace_bytes, st);
  return launch_conv_direct(fwd_params(d, x, nullptr, w, bias, y), d->dtype, st);
}

int spc_conv2d_fwd(const spc_conv_desc* d, const void* x, const spc_halo* halo, const void* w, const void* bias,
                   void* y, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = validate(d);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (d->N == 0) return SPC_OK;
  SPC_REQUIRE(x && w && y, "conv_fwd: null tensor pointer");
  if (!spc_conv_uses_tcgen05(d, 0) && d->algo != SPC_ALGO_TCGEN05)   // direct kernel reads tile + strips in one pass
    return launch_conv_direct(fwd_params(d, x, halo, w, bias, y), d->dtype, st);
  rc = fwd_interior(d, x, w, bias, y, workspace, workspace_bytes, st);
  if (rc) return rc;
  return fwd_boundary(d, fwd_params(d, x, halo, w, bias, y), halo, st);
}

int spc_conv2d_fwd_interior(const spc_conv_desc* d, const void* x, const void* w, const void* bias, void* y,
                            void* workspace, size_t workspace_bytes, void* stream) {
  int rc = validate(d);
  if (rc) return rc;
  if (d->N == 0) return SPC_OK;
  SPC_REQUIRE(x && w && y, "conv_fwd_interior: null tensor pointer");
  return fwd_interior(d, x, w, bias, y, workspace, workspace_bytes, (cudaStream_t)stream);
}

int spc_conv2d_fwd_boundary(const spc_conv_desc* d, const void* x, const spc_halo* halo, const void* w,
                            const void* bias, void* y, void* stream) {
  int rc = validate(d);
  if (rc) return rc;
  if (d->N == 0) return SPC_OK;
  SPC_REQUIRE(x && w && y, "conv_fwd_boundary: null tensor pointer");
  return fwd_boundary(d, fwd_params(d, x, halo, w, bias, y), halo, (cudaStream_t)stream);
}

int spc_conv2d_dgrad(const spc_conv_desc* d, const void* dy, const void* w, void* dx, void* workspace,
                     size_t workspace_bytes, void* stream) {
  int rc = validate(d);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (d->N == 0) return SPC_OK;
  SPC_REQUIRE(dy && w && dx, "conv_dgrad: null tensor pointer");
  const bool tc = spc_conv_uses_tcgen05(d, 1);
  if (d->algo == SPC_ALGO_TCGEN05 && !tc) {
    set_error("conv_dgrad: SPC_ALGO_TCGEN05 requested but the shape is not supported by the tcgen05 path");
    return SPC_EUNSUPPORTED;
  }
  if (tc) return tc_conv_dgrad(d, dy, w, dx, workspace, workspace_bytes, st);

  int Ho, Wo;
  spc_conv_out_shape(d, &Ho, &Wo);
  const int sh = d->stride_h, sw = d->stride_w, R = d->R, S = d->S;
  // Decompose by output parity class (a,b): each class is a stride-1 correlation of dy with a
  // flipped sub-filter (taps r = r_a + sh*t), written with output stride (sh,sw).
  bool need_zero = false;
  for (int a = 0; a < sh; ++a) if ((a + d->pad_h) % sh >= R) need_zero = true;
  for (int b = 0; b < sw; ++b) if ((b + d->pad_w) % sw >= S) need_zero = true;
  // rows/cols of dx beyond the reach of any output window are also zero; simplest: clear first
  if (need_zero || (Ho - 1) * sh + R - d->pad_h < d->H || (Wo - 1) * sw + S - d->pad_w < d->W)
    SPC_CHECK_CUDA(cudaMemsetAsync(dx, 0, (size_t)d->N * d->C * d->H * d->W * dtype_size(d->dtype), st));
  for (int a = 0; a < sh; ++a) {
    const int ra = (a + d->pad_h) % sh;
    if (ra >= R) continue;
    const int Ta = (R - ra + sh - 1) / sh;
    const int qa = (a + d->pad_h - ra) / sh;
    for (int b = 0; b < sw; ++b) {
      const int sb = (b + d->pad_w) % sw;
      if (sb >= S) continue;
      const int Tb = (S - sb + sw - 1) / sw;
      const int qb = (b + d->pad_w - sb) / sw;
      DirectConvParams p{};
      p.in = make_view(dy, nullptr, d->N, d->K, Ho, Wo, 0, 0);
      p.w = w; p.bias = nullptr; p.y = dx;
      p.K = d->C; p.R = Ta; p.S = Tb; p.sh = 1; p.sw = 1;
      p.pt = (Ta - 1) - qa; p.pl = (Tb - 1) - qb;
      p.Ho = (d->H - a + sh - 1) / sh; p.Wo = (d->W - b + sw - 1) / sw;
      p.YH = d->H; p.YW = d->W; p.oy0 = a; p.ox0 = b; p.oys = sh; p.oxs = sw;
      p.w_off = (long long)(ra + sh * (Ta - 1)) * S + (sb + sw * (Tb - 1));
      p.wKs = (long long)R * S;               // output channel of this launch = c
      p.wCs = (long long)d->C * R * S;        // input channel of this launch = k
      p.wRs = -(long long)sh * S; p.wSs = -(long long)sw;
      rc = launch_conv_direct(p, d->dtype, st);
      if (rc) return rc;
    }
  }
  return SPC_OK;
}

int spc_conv2d_wgrad(const spc_conv_desc* d, const void* x, const spc_halo* halo, const void* dy, float* dw,
                     float* db, int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = validate(d);
  if (rc) return rc;
  SPC_REQUIRE(dw && (d->N == 0 || (x && dy)), "conv_wgrad: null tensor pointer");
  cudaStream_t st = (cudaStream_t)stream;
  int Ho, Wo;
  spc_conv_out_shape(d, &Ho, &Wo);
  const size_t wn = (size_t)d->K * d->C * d->R * d->S;
  if (!accumulate) SPC_CHECK_CUDA(cudaMemsetAsync(dw, 0, wn * sizeof(float), st));
  if (d->N == 0) {
    if (db && !accumulate) SPC_CHECK_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * d->K, st));
    return SPC_OK;
  }
  const bool tc = spc_conv_uses_tcgen05(d, 2);
  if (d->algo == SPC_ALGO_TCGEN05 && !tc) {
    set_error("conv_wgrad: SPC_ALGO_TCGEN05 requested but the shape is not supported by the tcgen05 path");
    return SPC_EUNSUPPORTED;
  }
  DirectWgradParams p{};
  p.dy = dy; p.dw = dw;
  p.K = d->K; p.R = d->R; p.S = d->S; p.sh = d->stride_h; p.sw = d->stride_w; p.ph = d->pad_h; p.pw = d->pad_w;
  p.Ho = Ho; p.Wo = Wo;
  if (tc) {
    rc = tc_conv_wgrad(d, x, dy, dw, /*accumulate=*/1, workspace, workspace_bytes, st);
    if (rc) return rc;
    if (has_halo(halo)) {
      // add the halo pixels' contribution: the direct kernel over a view that holds ONLY the
      // strips (interior reads as zero) -- exact by linearity -- restricted to the output strips
      // whose windows reach outside the tile.
      if (d->dtype == SPC_BF16 && !getenv("SPC_BOUNDARY_V1")) {
        rc = boundary_wgrad_tc(d, halo, dy, dw, st);
        if (rc) return rc;
      } else if (patch_ok(d, 2)) {
        const int top = min(Ho, d->pad_h), bot0 = max(top, min(Ho, d->H + d->pad_h - d->R + 1));
        const int left = min(Wo, d->pad_w), right0 = max(left, min(Wo, d->W + d->pad_w - d->S + 1));
        const bool any_top = halo->strip[0] || halo->strip[1] || halo->strip[2];
        const bool any_bot = halo->strip[6] || halo->strip[7] || halo->strip[8];
        const bool any_left = halo->strip[0] || halo->strip[3] || halo->strip[6];
        const bool any_right = halo->strip[2] || halo->strip[5] || halo->strip[8];
        const int sy0 = any_top ? top : 0, sy1 = any_bot ? bot0 : Ho;
        if (any_top && (rc = patch_wgrad_rect(d, halo, dy, dw, 0, top, 0, Wo, st))) return rc;
        if (any_bot && (rc = patch_wgrad_rect(d, halo, dy, dw, bot0, Ho, 0, Wo, st))) return rc;
        if (any_left && (rc = patch_wgrad_rect(d, halo, dy, dw, sy0, sy1, 0, left, st))) return rc;
        if (any_right && (rc = patch_wgrad_rect(d, halo, dy, dw, sy0, sy1, right0, Wo, st))) return rc;
      } else {
        p.in = make_view(nullptr, halo, d->N, d->C, d->H, d->W, d->pad_h, d->pad_w);
        rc = launch_wgrad_halo(p, d->dtype, st);
        if (rc) return rc;
      }
    }
  } else {
    p.in = make_view(x, halo, d->N, d->C, d->H, d->W, d->pad_h, d->pad_w);
    rc = launch_wgrad_direct(p, d->dtype, st);
    if (rc) return rc;
  }
  if (db) return launch_bias_grad(dy, db, d->N, d->K, Ho * Wo, d->dtype, accumulate, st);
  return SPC_OK;
}

}  // extern "C"
