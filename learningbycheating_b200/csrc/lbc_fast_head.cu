// lbc_fast_head.cu -- fused waypoint heads of the bf16 path (image.py:54-60,82-84; common.py:136-152):
// 4 x [BatchNorm2d(64) -> Conv2d(64,5,1)+bias -> SpatialSoftmax] over the SAME decoder output h.
// The four BNs share batch statistics, so BN and the 1x1 convs fold into one 64->20 affine map per pixel:
//   logit[kj] = sum_c A[kj][c] * h[c] + b'[kj],  A = W_k[j,c]*gamma_kc*rstd_c,  b' = bias + sum_c W (beta - gamma*mean*rstd)
// forward = 1 pass over h; backward = 1 pass for the 20x64 moment matrix S and 1 pass producing d(h) with the decoder's
// final ReLU mask fused in (SURVEY.md 9.1 closed forms).
#include "lbc_fast.h"

#ifndef LBC_HOST_EMU
#include <cuda_bf16.h>
#endif

namespace lbc {
namespace fast {

#ifndef LBC_HOST_EMU
struct k_head_fold;
struct k_head_coef;

__device__ __forceinline__ void unpack8h(const uint4& v, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
static int sm_count3() {
  static int n = [] {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v > 0 ? v : 148;
  }();
  return n;
}

// fold[0..1280) = A[kj][c]; fold[1280..1300) = b'[kj]
bool head_fold(ref::HeadParams hp, float* fold, lbc_stream_t s) {
  par_for<k_head_fold>(s, 20, [=] __device__(int64_t kj) {
    int k = (int)kj / 5, j = (int)kj % 5;
    float b = hp.bias[k][j];
    for (int c = 0; c < 64; ++c) {
      float w = hp.w[k][j * 64 + c];
      float sc = hp.gamma[k][c] * hp.rstd[k][c];
      fold[kj * 64 + c] = w * sc;
      b += w * (hp.beta[k][c] - hp.mean[k][c] * sc);
    }
    fold[1280 + kj] = b;
  });
  return true;
}

// A thread owns FOUR pixels and reads the folded coefficients as float4, so one LDS.128 feeds 16 FMAs (a one-pixel-per-
// thread version issued one broadcast LDS per FMA, 1280 per pixel, and was shared-memory-issue bound: 116 us for a 126 MB
// read at B = 256).
__global__ void __launch_bounds__(128) head_logits4_kernel(const uint4* __restrict__ h, const float* __restrict__ fold,
                                                           float* __restrict__ logits, int64_t npix, int HW) {
  __shared__ __align__(16) float A[1300];
  for (int i = threadIdx.x; i < 1300; i += 128) A[i] = fold[i];
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * 512 + threadIdx.x;
  float acc[4][20];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int kj = 0; kj < 20; ++kj) acc[i][kj] = A[1280 + kj];
#pragma unroll 1
  for (int v = 0; v < 8; ++v) {
    float f[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t pix = base + 128 * i;
      if (pix < npix) {
        unpack8h(__ldg(h + pix * 8 + v), f[i]);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) f[i][e] = 0.f;
      }
    }
#pragma unroll
    for (int kj = 0; kj < 20; ++kj) {
      const float4 a0 = *reinterpret_cast<const float4*>(&A[kj * 64 + v * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&A[kj * 64 + v * 8 + 4]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float a = acc[i][kj];
        a += a0.x * f[i][0];
        a += a0.y * f[i][1];
        a += a0.z * f[i][2];
        a += a0.w * f[i][3];
        a += a1.x * f[i][4];
        a += a1.y * f[i][5];
        a += a1.z * f[i][6];
        a += a1.w * f[i][7];
        acc[i][kj] = a;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t pix = base + 128 * i;
    if (pix >= npix) continue;
    const int64_t n = pix / HW;
    const int p = (int)(pix - n * HW);
#pragma unroll
    for (int kj = 0; kj < 20; ++kj) logits[(n * 20 + kj) * HW + p] = acc[i][kj];
  }
}

// one block per (n, kj) row; HW <= 128*32
__global__ void __launch_bounds__(128) head_softmax_kernel(const float* __restrict__ logits, float* __restrict__ rowmax,
                                                           float* __restrict__ rowsum, float* __restrict__ preds, int H, int W) {
  const int HW = H * W;
  const float* l = logits + (int64_t)blockIdx.x * HW;
  float v[32];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    int p = i * 128 + threadIdx.x;
    v[i] = p < HW ? l[p] : -INFINITY;
    m = fmaxf(m, v[i]);
  }
  __shared__ float red[4][4];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = m;
  __syncthreads();
  m = fmaxf(fmaxf(red[0][0], red[0][1]), fmaxf(red[0][2], red[0][3]));
  __syncthreads();
  float se = 0.f, sx = 0.f, sy = 0.f;
  const float kx = W > 1 ? 2.0f / (float)(W - 1) : 0.f, ky = H > 1 ? 2.0f / (float)(H - 1) : 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    int p = i * 128 + threadIdx.x;
    if (p < HW) {
      float e = __expf(v[i] - m);
      int hh = p / W, ww = p - hh * W;
      se += e;
      sx += e * (-1.f + kx * (float)ww);
      sy += e * (-1.f + ky * (float)hh);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    se += __shfl_xor_sync(0xffffffffu, se, o);
    sx += __shfl_xor_sync(0xffffffffu, sx, o);
    sy += __shfl_xor_sync(0xffffffffu, sy, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[1][threadIdx.x >> 5] = se;
    red[2][threadIdx.x >> 5] = sx;
    red[3][threadIdx.x >> 5] = sy;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    se = red[1][0] + red[1][1] + red[1][2] + red[1][3];
    sx = red[2][0] + red[2][1] + red[2][2] + red[2][3];
    sy = red[3][0] + red[3][1] + red[3][2] + red[3][3];
    rowmax[blockIdx.x] = m;
    rowsum[blockIdx.x] = se;
    preds[blockIdx.x * 2] = sx / se;
    preds[blockIdx.x * 2 + 1] = sy / se;
  }
}

// SpatialSoftmax alone (common.py:136-152) over [N*20] rows of H*W logits
bool head_softmax_f32(const float* logits, float* rowmax, float* rowsum, float* preds, int N, int H, int W, lbc_stream_t s) {
  if (!enabled() || H * W > 4096) return false;
  head_softmax_kernel<<<N * 20, 128, 0, s>>>(logits, rowmax, rowsum, preds, H, W);
  LBC_LAUNCHED("head_softmax_kernel");
  LBC_CUDA(cudaGetLastError());
  return true;
}
bool head_forward_bf16(const bf16* h, ref::HeadParams hp, float* fold, float* logits, float* rowmax, float* rowsum,
                       float* preds, int N, int H, int W, lbc_stream_t s) {
  if (!enabled()) return false;
  const int HW = H * W;
  if (HW > 4096) return false;
  head_fold(hp, fold, s);
  const int64_t npix = (int64_t)N * HW;
  head_logits4_kernel<<<(unsigned)((npix + 511) / 512), 128, 0, s>>>((const uint4*)h, fold, logits, npix, HW);
  LBC_LAUNCHED("head_logits4_kernel");
  head_softmax_kernel<<<N * 20, 128, 0, s>>>(logits, rowmax, rowsum, preds, H, W);
  LBC_LAUNCHED("head_softmax_kernel");
  LBC_CUDA(cudaGetLastError());
  return true;
}

// coef[0..64) = c0, coef[64..128) = c1 :  dh = A^T dl - c0 - hhat*c1
bool head_coef(ref::HeadParams hp, ref::HeadGrads hg, float* coef, float invM, lbc_stream_t s) {
  par_for<k_head_coef>(s, 64, [=] __device__(int64_t c) {
    float a = 0.f, b = 0.f;
    for (int k = 0; k < 4; ++k) {
      a += hp.gamma[k][c] * hg.dbeta[k][c];
      b += hp.gamma[k][c] * hg.dgamma[k][c];
    }
    coef[c] = hp.rstd[0][c] * a * invM;
    coef[64 + c] = hp.rstd[0][c] * b * invM;
  });
  return true;
}
// d(pre-ReLU deconv output)[pix][c] = (h>0) * (sum_kj A[kj][c] dl[kj] - c0[c] - hhat[c]*c1[c]);
// four pixels per thread, coefficients read as float4 over channels (one LDS.128 per 16 FMAs).
__global__ void __launch_bounds__(128) head_dh4_kernel(const float* __restrict__ dlogits, const uint4* __restrict__ h,
                                                       const float* __restrict__ fold, const float* __restrict__ coef,
                                                       const float* __restrict__ mean, const float* __restrict__ rstd,
                                                       uint4* __restrict__ dh, int64_t npix, int HW) {
  __shared__ __align__(16) float A[1280];
  __shared__ __align__(16) float cf[256];  // c0, c1, mean, rstd
  for (int i = threadIdx.x; i < 1280; i += 128) A[i] = fold[i];
  if (threadIdx.x < 64) {
    cf[threadIdx.x] = coef[threadIdx.x];
    cf[64 + threadIdx.x] = coef[64 + threadIdx.x];
    cf[128 + threadIdx.x] = mean[threadIdx.x];
    cf[192 + threadIdx.x] = rstd[threadIdx.x];
  }
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * 512 + threadIdx.x;
  float d[4][20];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t pix = base + 128 * i;
    const bool ok = pix < npix;
    const int64_t n = ok ? pix / HW : 0;
    const int p = ok ? (int)(pix - n * HW) : 0;
#pragma unroll
    for (int kj = 0; kj < 20; ++kj) d[i][kj] = ok ? dlogits[(n * 20 + kj) * HW + p] : 0.f;
  }
#pragma unroll 1
  for (int v = 0; v < 8; ++v) {
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[i][e] = 0.f;
#pragma unroll
    for (int kj = 0; kj < 20; ++kj) {
      const float4 a0 = *reinterpret_cast<const float4*>(&A[kj * 64 + v * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&A[kj * 64 + v * 8 + 4]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float dk = d[i][kj];
        acc[i][0] += a0.x * dk;
        acc[i][1] += a0.y * dk;
        acc[i][2] += a0.z * dk;
        acc[i][3] += a0.w * dk;
        acc[i][4] += a1.x * dk;
        acc[i][5] += a1.y * dk;
        acc[i][6] += a1.z * dk;
        acc[i][7] += a1.w * dk;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t pix = base + 128 * i;
      if (pix >= npix) continue;
      float f[8], o[8];
      unpack8h(__ldg(h + pix * 8 + v), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = v * 8 + e;
        const float xh = (f[e] - cf[128 + c]) * cf[192 + c];
        const float a = acc[i][e] - cf[c] - xh * cf[64 + c];
        o[e] = f[e] > 0.f ? a : 0.f;
      }
      uint4 w;
      __nv_bfloat162* hb = reinterpret_cast<__nv_bfloat162*>(&w);
#pragma unroll
      for (int q = 0; q < 4; ++q) hb[q] = __floats2bfloat162_rn(o[2 * q], o[2 * q + 1]);
      dh[pix * 8 + v] = w;
    }
  }
}

// dlogit = softmax * ((x - Ex) gx + (y - Ey) gy), g = d_preds + onehot * d_pred: one block per (n, kj) row, the
// row constants read once, coalesced 4-byte traffic.  (The one-thread-per-element version recomputed the row constants and
// the grid coordinates with integer divisions per element: 167 us for a 157 MB read + write at B = 256.)
__global__ void __launch_bounds__(128) head_dlogits_kernel(const float* __restrict__ logits, const float* __restrict__ rowmax,
                                                           const float* __restrict__ rowsum, const float* __restrict__ preds,
                                                           const float* __restrict__ onehot, const float* __restrict__ d_pred,
                                                           const float* __restrict__ d_preds, float* __restrict__ dlogits, int H,
                                                           int W) {
  const int HW = H * W;
  const int r = blockIdx.x;
  const int kj = r % 20, b = r / 20;
  const int k = kj / 5, j = kj % 5;
  float gx = 0.f, gy = 0.f;
  if (d_preds) {
    gx += d_preds[r * 2];
    gy += d_preds[r * 2 + 1];
  }
  if (d_pred) {
    const float oh = onehot[b * 4 + k];
    gx += oh * d_pred[(b * 5 + j) * 2];
    gy += oh * d_pred[(b * 5 + j) * 2 + 1];
  }
  const float m = rowmax[r], inv = 1.0f / rowsum[r], ex = preds[r * 2], ey = preds[r * 2 + 1];
  const float kx = W > 1 ? 2.0f / (float)(W - 1) : 0.f, ky = H > 1 ? 2.0f / (float)(H - 1) : 0.f;
  const float* l = logits + (int64_t)r * HW;
  float* d = dlogits + (int64_t)r * HW;
  for (int p = threadIdx.x; p < HW; p += 128) {
    const int hh = p / W, ww = p - hh * W;
    const float wgt = __expf(l[p] - m) * inv;
    d[p] = wgt * ((-1.f + kx * (float)ww - ex) * gx + (-1.f + ky * (float)hh - ey) * gy);
  }
}
bool head_dlogits_f32(const float* logits, const float* rowmax, const float* rowsum, const float* preds, const float* onehot,
                      const float* d_pred, const float* d_preds, float* dlogits, int N, int H, int W, lbc_stream_t s) {
  if (!enabled()) return false;
  head_dlogits_kernel<<<N * 20, 128, 0, s>>>(logits, rowmax, rowsum, preds, onehot, d_pred, d_preds, dlogits, H, W);
  LBC_LAUNCHED("head_dlogits_kernel");
  LBC_CUDA(cudaGetLastError());
  return true;
}

// S[kj][c] += sum_pix dl[n,kj,pix]*hhat[n,pix,c] (c<64), S[kj][64] += sum dl   (double atomics, S pre-zeroed):
// moment matrix with 4 channels x 5 kj x 4 pixels per inner step (9 LDS.128 per 80 FMAs).  Thread = (channel quad 0..15,
// kj group 0..3, pixel quarter 0..3); the four pixel quarters are combined in shared memory before the double atomics.
__global__ void __launch_bounds__(256) head_s4_kernel(const float* __restrict__ dlogits, const uint4* __restrict__ h,
                                                      const float* __restrict__ mean, const float* __restrict__ rstd, double* S,
                                                      int N, int HW) {
  __shared__ __align__(16) float smem_f[128 * 68 + 20 * 128];   // 45 KB: hh | dl ; `part` reuses the hh space at the end
  float(*hh)[68] = reinterpret_cast<float(*)[68]>(smem_f);
  float(*dl)[128] = reinterpret_cast<float(*)[128]>(smem_f + 128 * 68);
  float(*part)[20][65] = reinterpret_cast<float(*)[20][65]>(smem_f);   // 4 x 20 x 65 floats = 20.8 KB <= 34.8 KB
  const int t = threadIdx.x;
  const int cq = t & 15, grp = (t >> 4) & 3, pq = t >> 6;
  float acc[5][4];
  float acc0[5];
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    acc0[j] = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
  }
  const int tiles_per_img = (HW + 127) / 128;
  const int ntiles = N * tiles_per_img;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int n = tile / tiles_per_img;
    const int p0 = (tile - n * tiles_per_img) * 128;
    const int np = min(128, HW - p0);
    __syncthreads();
    for (int i = t; i < 20 * 128; i += 256) {
      int kj = i >> 7, p = i & 127;
      dl[kj][p] = p < np ? dlogits[((int64_t)n * 20 + kj) * HW + p0 + p] : 0.f;
    }
    for (int i = t; i < 128 * 8; i += 256) {
      int p = i >> 3, v = i & 7;
      float f[8];
      if (p < np) {
        unpack8h(__ldg(h + ((int64_t)n * HW + p0 + p) * 8 + v), f);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) hh[p][v * 8 + e] = p < np ? (f[e] - mean[v * 8 + e]) * rstd[v * 8 + e] : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int pp = 0; pp < 32; pp += 4) {
      const int p = pq * 32 + pp;
      float4 x[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) x[i] = *reinterpret_cast<const float4*>(&hh[p + i][cq * 4]);
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const float4 d4 = *reinterpret_cast<const float4*>(&dl[grp * 5 + j][p]);
        const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[j][0] += dv[i] * x[i].x;
          acc[j][1] += dv[i] * x[i].y;
          acc[j][2] += dv[i] * x[i].z;
          acc[j][3] += dv[i] * x[i].w;
        }
        if (cq == 0) acc0[j] += (dv[0] + dv[1]) + (dv[2] + dv[3]);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 5; ++j) {
#pragma unroll
    for (int e = 0; e < 4; ++e) part[pq][grp * 5 + j][cq * 4 + e] = acc[j][e];
    if (cq == 0) part[pq][grp * 5 + j][64] = acc0[j];
  }
  __syncthreads();
  for (int i = t; i < 20 * 65; i += 256) {
    const int kj = i / 65, c = i - kj * 65;
    const float v = (part[0][kj][c] + part[1][kj][c]) + (part[2][kj][c] + part[3][kj][c]);
    atomicAdd(&S[kj * 65 + c], (double)v);
  }
}

// S: [20][65] doubles (zeroed here).  Produces S, then the caller computes parameter grads (ref::head_param_grads),
// then head_backward_dh_bf16 writes the masked d(h).
bool head_backward_s_bf16(const float* dlogits, const bf16* h, const float* mean, const float* rstd, double* S, int N, int HW,
                          lbc_stream_t s) {
  if (!enabled()) return false;
  LBC_CUDA(cudaMemsetAsync(S, 0, sizeof(double) * 20 * 65, s));
  int grid = sm_count3() * 2;
  int ntiles = N * ((HW + 127) / 128);
  if (grid > ntiles) grid = ntiles;
  head_s4_kernel<<<grid, 256, 0, s>>>(dlogits, (const uint4*)h, mean, rstd, S, N, HW);
  LBC_LAUNCHED("head_s4_kernel");
  LBC_CUDA(cudaGetLastError());
  return true;
}
bool head_backward_dh_bf16(const float* dlogits, const bf16* h, ref::HeadParams hp, ref::HeadGrads hg, const float* fold,
                           float* coef, bf16* dh, int N, int HW, lbc_stream_t s) {
  if (!enabled()) return false;
  const int64_t npix = (int64_t)N * HW;
  head_coef(hp, hg, coef, 1.0f / (float)npix, s);
  head_dh4_kernel<<<(unsigned)((npix + 511) / 512), 128, 0, s>>>(dlogits, (const uint4*)h, fold, coef, hp.mean[0], hp.rstd[0],
                                                               (uint4*)dh, npix, HW);
  LBC_LAUNCHED("head_dh4_kernel");
  LBC_CUDA(cudaGetLastError());
  return true;
}

#else
bool head_softmax_f32(const float*, float*, float*, float*, int, int, int, lbc_stream_t) { return false; }
bool head_forward_bf16(const bf16*, ref::HeadParams, float*, float*, float*, float*, float*, int, int, int, lbc_stream_t) {
  return false;
}
bool head_backward_s_bf16(const float*, const bf16*, const float*, const float*, double*, int, int, lbc_stream_t) { return false; }
bool head_dlogits_f32(const float*, const float*, const float*, const float*, const float*, const float*, const float*, float*, int,
                      int, int, lbc_stream_t) { return false; }
bool head_backward_dh_bf16(const float*, const bf16*, ref::HeadParams, ref::HeadGrads, const float*, float*, bf16*, int, int,
                           lbc_stream_t) { return false; }
#endif

}  // namespace fast
}  // namespace lbc
