/* ORACLE -- TEST INFRASTRUCTURE ONLY (not part of the product).
 *
 * Plain-C restatement of the arithmetic on the reference's hot path, used as (a) a second,
 * torch-free checker next to oracle/spatial_oracle.py and (b) the "port" CPU baseline timed by
 * bench.py (cpu_baseline / --impl reference).  It computes what the reference computes on a
 * tile AFTER its halo exchange: `nn.Conv2d(padding=0)` / `nn.AvgPool2d` / `nn.MaxPool2d` on
 * the explicitly padded tile (reference src/torchgems/spatial.py:1027, :1478-1498), and the
 * autograd of those calls.  The arithmetic lives in PyTorch/cuDNN/oneDNN (not vendored by the
 * reference, pinned only by prose "PyTorch 1.12.1 or 1.13.1", README.md:109); this file restates
 * the published definition (cross-correlation, fp32 accumulate in double).
 *
 * Pinned by tests/test_oracle_golden.py against golden vectors from the unmodified reference.
 * Layout: NCHW contiguous float32.  OpenMP over (n, k) planes.
 */
#include <math.h>
#include <stddef.h>
#include <string.h>

#define IDX4(n, c, h, w, C, H, W) ((((size_t)(n) * (C) + (c)) * (H) + (h)) * (W) + (w))

/* y[n,k,i,j] = b[k] + sum_{c,r,s} w[k,c,r,s] * xp[n,c,i*sh+r,j*sw+s]      (spatial.py:1027) */
void ref_conv2d_fwd(const float* xp, const float* w, const float* b, float* y, int N, int C, int Hp, int Wp,
                    int K, int R, int S, int sh, int sw) {
  const int Ho = (Hp - R) / sh + 1, Wo = (Wp - S) / sw + 1;
#pragma omp parallel for collapse(2) schedule(static)
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      float* yp = y + IDX4(n, k, 0, 0, K, Ho, Wo);
      for (int i = 0; i < Ho * Wo; ++i) yp[i] = b ? b[k] : 0.f;
      for (int c = 0; c < C; ++c)
        for (int r = 0; r < R; ++r)
          for (int s = 0; s < S; ++s) {
            const float wv = w[(((size_t)k * C + c) * R + r) * S + s];
            for (int i = 0; i < Ho; ++i) {
              const float* xr = xp + IDX4(n, c, i * sh + r, s, C, Hp, Wp);
              float* yr = yp + (size_t)i * Wo;
              for (int j = 0; j < Wo; ++j) yr[j] += wv * xr[(size_t)j * sw];
            }
          }
    }
}

/* dxp = autograd of ref_conv2d_fwd w.r.t. the padded tile (caller crops = ZeroPad2d backward) */
void ref_conv2d_dgrad(const float* gy, const float* w, float* dxp, int N, int C, int Hp, int Wp, int K, int R, int S,
                      int sh, int sw) {
  const int Ho = (Hp - R) / sh + 1, Wo = (Wp - S) / sw + 1;
  memset(dxp, 0, sizeof(float) * (size_t)N * C * Hp * Wp);
#pragma omp parallel for collapse(2) schedule(static)
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < C; ++c)
      for (int k = 0; k < K; ++k)
        for (int r = 0; r < R; ++r)
          for (int s = 0; s < S; ++s) {
            const float wv = w[(((size_t)k * C + c) * R + r) * S + s];
            for (int i = 0; i < Ho; ++i) {
              const float* gr = gy + IDX4(n, k, i, 0, K, Ho, Wo);
              float* dr = dxp + IDX4(n, c, i * sh + r, s, C, Hp, Wp);
              for (int j = 0; j < Wo; ++j) dr[(size_t)j * sw] += wv * gr[j];
            }
          }
}

/* dw[k,c,r,s] = sum_{n,i,j} gy[n,k,i,j] * xp[n,c,i*sh+r,j*sw+s];  db[k] = sum gy[n,k,:,:] */
void ref_conv2d_wgrad(const float* xp, const float* gy, float* dw, float* db, int N, int C, int Hp, int Wp, int K,
                      int R, int S, int sh, int sw) {
  const int Ho = (Hp - R) / sh + 1, Wo = (Wp - S) / sw + 1;
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 0; k < K; ++k)
    for (int c = 0; c < C; ++c)
      for (int r = 0; r < R; ++r)
        for (int s = 0; s < S; ++s) {
          double acc = 0.0;
          for (int n = 0; n < N; ++n)
            for (int i = 0; i < Ho; ++i) {
              const float* gr = gy + IDX4(n, k, i, 0, K, Ho, Wo);
              const float* xr = xp + IDX4(n, c, i * sh + r, s, C, Hp, Wp);
              double a = 0.0;
              for (int j = 0; j < Wo; ++j) a += (double)gr[j] * xr[(size_t)j * sw];
              acc += a;
            }
          dw[(((size_t)k * C + c) * R + r) * S + s] = (float)acc;
        }
  if (db)
    for (int k = 0; k < K; ++k) {
      double acc = 0.0;
      for (int n = 0; n < N; ++n) {
        const float* g = gy + IDX4(n, k, 0, 0, K, Ho, Wo);
        for (int i = 0; i < Ho * Wo; ++i) acc += g[i];
      }
      db[k] = (float)acc;
    }
}

/* mode 0 = max, 1 = avg; padding=0 on the padded tile (spatial.py:1478-1498) */
void ref_pool2d_fwd(const float* xp, float* y, int N, int C, int Hp, int Wp, int k, int stride, int mode) {
  const int Ho = (Hp - k) / stride + 1, Wo = (Wp - k) / stride + 1;
#pragma omp parallel for schedule(static)
  for (int nc = 0; nc < N * C; ++nc)
    for (int i = 0; i < Ho; ++i)
      for (int j = 0; j < Wo; ++j) {
        float m = -INFINITY;
        double s = 0.0;
        for (int a = 0; a < k; ++a)
          for (int b = 0; b < k; ++b) {
            const float v = xp[((size_t)nc * Hp + i * stride + a) * Wp + j * stride + b];
            if (v > m) m = v;
            s += v;
          }
        y[((size_t)nc * Ho + i) * Wo + j] = mode == 0 ? m : (float)(s / (k * k));
      }
}

/* dxp of ref_pool2d_fwd; max routes to the first maximal element in row-major window order */
void ref_pool2d_bwd(const float* xp, const float* gy, float* dxp, int N, int C, int Hp, int Wp, int k, int stride,
                    int mode) {
  const int Ho = (Hp - k) / stride + 1, Wo = (Wp - k) / stride + 1;
  memset(dxp, 0, sizeof(float) * (size_t)N * C * Hp * Wp);
#pragma omp parallel for schedule(static)
  for (int nc = 0; nc < N * C; ++nc)
    for (int i = 0; i < Ho; ++i)
      for (int j = 0; j < Wo; ++j) {
        const float g = gy[((size_t)nc * Ho + i) * Wo + j];
        if (mode == 1) {
          for (int a = 0; a < k; ++a)
            for (int b = 0; b < k; ++b) dxp[((size_t)nc * Hp + i * stride + a) * Wp + j * stride + b] += g / (k * k);
        } else {
          float m = -INFINITY;
          int ba = 0, bb = 0;
          for (int a = 0; a < k; ++a)
            for (int b = 0; b < k; ++b) {
              const float v = xp[((size_t)nc * Hp + i * stride + a) * Wp + j * stride + b];
              if (v > m) { m = v; ba = a; bb = b; }
            }
          dxp[((size_t)nc * Hp + i * stride + ba) * Wp + j * stride + bb] += g;
        }
      }
}
