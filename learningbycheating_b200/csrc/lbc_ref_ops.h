// lbc_ref_ops.h -- correctness-first CUDA kernels for every op of the LBC image-agent step.
//
// NHWC activations (storage type T = float or bf16, fp32 math), conv weights packed as
// [Co][KH][KW][Ci].  One logical thread per output element, deterministic reductions through
// per-chunk partials (double accumulation).  These kernels are (a) the fp32 parity path that is
// compared with the CPU oracle at <=1e-3 and (b) the on-device checker for the tcgen05 / fused
// kernels in lbc_fast_*.cu.  Reference semantics cited per op (paths under /root/reference).
#pragma once
#include "lbc_common.h"

namespace lbc {
namespace ref {

// tags (kernel names in ncu)
struct k_input_nhwc; struct k_conv_fwd; struct k_conv_dgrad; struct k_conv_wgrad_part; struct k_reduce_part;
struct k_bn_sum_part; struct k_bn_var_part; struct k_bn_finalize; struct k_bn_apply; struct k_maxpool_fwd;
struct k_maxpool_bwd; struct k_relu_mask; struct k_add; struct k_bn_bwd_part; struct k_bn_bwd_final;
struct k_bn_bwd_apply; struct k_concat_speed; struct k_slice; struct k_colsum_part; struct k_head_logits;
struct k_head_softmax; struct k_head_select; struct k_head_dlogits; struct k_head_s_part; struct k_head_param_grads;
struct k_head_dh; struct k_pack_w; struct k_adam; struct k_l1_loss; struct k_phase0_target; struct k_phase1_fwd;
struct k_phase1_bwd; struct k_cast; struct k_speed_stats; struct k_scale; struct k_negate;
inline void negate_into(lbc_stream_t s, const float* src, float* dst, int64_t n) {
  par_for<k_negate>(s, n, [=] LBC_LAMBDA(int64_t i) { dst[i] = -src[i]; });
}

// accumulator type: double for the fp32 parity path (the CPU oracle's blocked GEMMs / double-accumulating
// reductions are far more accurate than a sequential fp32 sum), float for bf16 storage
template <class T> struct Acc { typedef float type; };
template <> struct Acc<float> { typedef double type; };

// ---------------------------------------------------------------------------------------------
// image NCHW fp32 -> NHWC T, optional (x-mean)/std   (bird_view/models/common.py:108-109)
template <class T>
void input_to_nhwc(lbc_stream_t s, const float* img, T* out, int N, int C, int H, int W, int Cpad,
                   bool normalize, float m0, float m1, float m2, float s0, float s1, float s2) {
  int64_t n = (int64_t)N * H * W * Cpad;
  par_for<k_input_nhwc>(s, n, [=] LBC_LAMBDA(int64_t i) {
    int c = (int)(i % Cpad);
    int64_t p = i / Cpad;
    int w = (int)(p % W);
    int64_t q = p / W;
    int h = (int)(q % H);
    int b = (int)(q / H);
    float v = 0.f;
    if (c < C) {
      v = img[(((int64_t)b * C + c) * H + h) * W + w];
      if (normalize) {
        float mean = c == 0 ? m0 : (c == 1 ? m1 : m2);
        float sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
        v = (v - mean) / sd;
      }
    }
    stf(out, i, v);
  });
}

// torchvision ToTensor on the device: u8 ([B,C,H,W] layout 0 or [B,H,W,C] layout 1) -> fp32 [B,C,H,W] in [0,1]
struct k_u8_to_f32;
inline void u8_to_f32_nchw(lbc_stream_t s, const uint8_t* src, float* dst, int N, int C, int H, int W, int layout) {
  int64_t n = (int64_t)N * C * H * W;
  par_for<k_u8_to_f32>(s, n, [=] LBC_LAMBDA(int64_t i) {
    int64_t j = i;
    if (layout == 1) {
      int w = (int)(i % W);
      int64_t t = i / W;
      int h = (int)(t % H);
      t /= H;
      int c = (int)(t % C);
      int b = (int)(t / C);
      j = (((int64_t)b * H + h) * W + w) * C + c;
    }
    dst[i] = (float)src[j] / 255.0f;
  });
}

// ---------------------------------------------------------------------------------------------
// y[n,oh,ow,co] = sum x[n,oh*s-p+kh,ow*s-p+kw,ci] * w[co,kh,kw,ci]  (+bias[co]) (relu)
// nn.Conv2d semantics, bird_view/models/resnet.py:15-22,102-103
template <class T>
void conv_fwd(lbc_stream_t s, const T* x, const T* w, const float* bias, bool relu, T* y, int N, int H,
              int W, int Ci, int Co, int K, int stride, int pad, int OH, int OW) {
  int64_t n = (int64_t)N * OH * OW * Co;
  par_for<k_conv_fwd>(s, n, [=] LBC_LAMBDA(int64_t i) {
    int co = (int)(i % Co);
    int64_t p = i / Co;
    int ow = (int)(p % OW);
    int64_t q = p / OW;
    int oh = (int)(q % OH);
    int b = (int)(q / OH);
    typedef typename Acc<T>::type acc_t;
    acc_t acc = bias ? (acc_t)bias[co] : (acc_t)0;
    for (int kh = 0; kh < K; ++kh) {
      int ih = oh * stride - pad + kh;
      if (ih < 0 || ih >= H) continue;
      for (int kw = 0; kw < K; ++kw) {
        int iw = ow * stride - pad + kw;
        if (iw < 0 || iw >= W) continue;
        const T* xp = x + (((int64_t)b * H + ih) * W + iw) * Ci;
        const T* wp = w + (((int64_t)co * K + kh) * K + kw) * Ci;
        acc_t a0 = 0, a1 = 0;
        int ci = 0;
        for (; ci + 1 < Ci; ci += 2) {
          a0 += (acc_t)ldf(xp, ci) * (acc_t)ldf(wp, ci);
          a1 += (acc_t)ldf(xp, ci + 1) * (acc_t)ldf(wp, ci + 1);
        }
        if (ci < Ci) a0 += (acc_t)ldf(xp, ci) * (acc_t)ldf(wp, ci);
        acc += a0 + a1;
      }
    }
    float r = (float)acc;
    if (relu) r = r > 0.f ? r : 0.f;
    stf(y, i, r);
  });
}

// dx[n,ih,iw,ci] = sum_{kh,kw,co} dy[n,oh,ow,co] * w[co,kh,kw,ci],  oh=(ih+p-kh)/s when divisible.
// Also nn.ConvTranspose2d(k3,s2,p1,op1) forward (image.py:39-46): x_deconv plays dy, the output
// plays dx, bias indexed by ci, optional ReLU.
template <class T>
void conv_dgrad(lbc_stream_t s, const T* dy, const T* w, T* dx, int N, int H, int W, int Ci, int Co, int K,
                int stride, int pad, int OH, int OW, const float* bias_ci, bool relu, bool accumulate) {
  int64_t n = (int64_t)N * H * W * Ci;
  par_for<k_conv_dgrad>(s, n, [=] LBC_LAMBDA(int64_t i) {
    int ci = (int)(i % Ci);
    int64_t p = i / Ci;
    int iw = (int)(p % W);
    int64_t q = p / W;
    int ih = (int)(q % H);
    int b = (int)(q / H);
    typedef typename Acc<T>::type acc_t;
    acc_t acc = bias_ci ? (acc_t)bias_ci[ci] : (acc_t)0;
    for (int kh = 0; kh < K; ++kh) {
      int t = ih + pad - kh;
      if (t < 0 || (t % stride) != 0) continue;
      int oh = t / stride;
      if (oh >= OH) continue;
      for (int kw = 0; kw < K; ++kw) {
        int u = iw + pad - kw;
        if (u < 0 || (u % stride) != 0) continue;
        int ow = u / stride;
        if (ow >= OW) continue;
        const T* dyp = dy + (((int64_t)b * OH + oh) * OW + ow) * Co;
        const T* wp = w + ((int64_t)kh * K + kw) * Ci + ci;
        int64_t wstride = (int64_t)K * K * Ci;
        acc_t a0 = 0;
        for (int co = 0; co < Co; ++co) a0 += (acc_t)ldf(dyp, co) * (acc_t)ldf(wp, (int64_t)co * wstride);
        acc += a0;
      }
    }
    if (accumulate) acc += (acc_t)ldf(dx, i);
    float r = (float)acc;
    if (relu) r = r > 0.f ? r : 0.f;
    stf(dx, i, r);
  });
}

// dw[co][ci][kh][kw] (the reference's nn.Conv2d layout; the same index formula is the
// nn.ConvTranspose2d layout [C_in][C_out][kh][kw] when the conv roles are swapped)
//   = sum_{n,oh,ow} dy[n,oh,ow,co] * x[n,oh*s-p+kh,ow*s-p+kw,ci].   ws: >= P*Co*K*K*Ci floats.
template <class T>
void conv_wgrad(lbc_stream_t s, const T* x, const T* dy, float* dw_ref, int N, int H, int W, int Ci, int Co,
                int K, int stride, int pad, int OH, int OW, float* ws, int64_t ws_floats) {
  int64_t wsize = (int64_t)Co * K * K * Ci;
  int64_t rows = (int64_t)N * OH;
  int64_t P = (1 << 19) / wsize;
  if (P < 1) P = 1;
  if (P > rows) P = rows;
  if (P * wsize > ws_floats) P = ws_floats / wsize;
  LBC_CHECK(P >= 1, "conv_wgrad workspace too small");
  int64_t rows_per = cdiv(rows, P);
  P = cdiv(rows, rows_per);
  par_for<k_conv_wgrad_part>(s, P * wsize, [=] LBC_LAMBDA(int64_t i) {
    int ci = (int)(i % Ci);
    int64_t t = i / Ci;
    int kw = (int)(t % K);
    t /= K;
    int kh = (int)(t % K);
    t /= K;
    int co = (int)(t % Co);
    int64_t chunk = t / Co;
    int64_t r0 = chunk * rows_per, r1 = r0 + rows_per;
    if (r1 > rows) r1 = rows;
    typedef typename Acc<T>::type acc_t;
    acc_t acc = 0;
    for (int64_t r = r0; r < r1; ++r) {
      int b = (int)(r / OH), oh = (int)(r % OH);
      int ih = oh * stride - pad + kh;
      if (ih < 0 || ih >= H) continue;
      const T* dyp = dy + (((int64_t)b * OH + oh) * OW) * Co + co;
      const T* xp = x + (((int64_t)b * H + ih) * W) * Ci + ci;
      acc_t a = 0;
      for (int ow = 0; ow < OW; ++ow) {
        int iw = ow * stride - pad + kw;
        if (iw < 0 || iw >= W) continue;
        a += (acc_t)ldf(dyp, (int64_t)ow * Co) * (acc_t)ldf(xp, (int64_t)iw * Ci);
      }
      acc += a;
    }
    ws[i] = (float)acc;
  });
  par_for<k_reduce_part>(s, wsize, [=] LBC_LAMBDA(int64_t i) {
    double a = 0.0;
    for (int64_t c = 0; c < P; ++c) a += (double)ws[c * wsize + i];
    int ci = (int)(i % Ci);
    int64_t t = i / Ci;
    int kw = (int)(t % K);
    t /= K;
    int kh = (int)(t % K);
    int co = (int)(t / K);
    dw_ref[(((int64_t)co * Ci + ci) * K + kh) * K + kw] = (float)a;
  });
}

// reference layout [Co][Ci][K][K] fp32 -> packed [Co][K][K][Ci] T
template <class T>
void pack_weight(lbc_stream_t s, const float* w_ref, T* w_packed, int Co, int Ci, int K) {
  int64_t n = (int64_t)Co * K * K * Ci;
  par_for<k_pack_w>(s, n, [=] LBC_LAMBDA(int64_t i) {
    int ci = (int)(i % Ci);
    int64_t t = i / Ci;
    int kw = (int)(t % K);
    t /= K;
    int kh = (int)(t % K);
    int co = (int)(t / K);
    stf(w_packed, i, w_ref[(((int64_t)co * Ci + ci) * K + kh) * K + kw]);
  });
}

// reference layout [Co][Ci][K][K] fp32 -> transposed pack [Ci][K][K][Co] T
template <class T>
void pack_weight_t(lbc_stream_t s, const float* w_ref, T* w_packed, int Co, int Ci, int K) {
  int64_t n = (int64_t)Co * K * K * Ci;
  par_for<k_pack_w>(s, n, [=] LBC_LAMBDA(int64_t i) {
    int co = (int)(i % Co);
    int64_t t = i / Co;
    int kw = (int)(t % K);
    t /= K;
    int kh = (int)(t % K);
    int ci = (int)(t / K);
    stf(w_packed, i, w_ref[(((int64_t)co * Ci + ci) * K + kh) * K + kw]);
  });
}

// block-entry combined pack: wcomb[ci][k], k < Co: centre tap of the 3x3 conv W1[k][ci][1][1]; k >= Co: Wd[k-Co][ci]
template <class T>
void pack_weight_comb(lbc_stream_t s, const float* w1_ref, const float* wd_ref, T* wcomb, int Co, int Ci) {
  int64_t n = (int64_t)Ci * 2 * Co;
  par_for<k_pack_w>(s, n, [=] LBC_LAMBDA(int64_t i) {
    int k = (int)(i % (2 * Co));
    int ci = (int)(i / (2 * Co));
    float v = k < Co ? w1_ref[(((int64_t)k * Ci + ci) * 3 + 1) * 3 + 1] : wd_ref[(int64_t)(k - Co) * Ci + ci];
    stf(wcomb, i, v);
  });
}

// stem 7x7 weights [64][C][7][7] -> GEMM rows [64][Kp], k = (kh*7+kw)*C + c, zero padded (the layout of pack_all type 3)
inline void pack_stem_weight(lbc_stream_t s, const float* w_ref, float* wp, int C, int Kp) {
  par_for<k_pack_w>(s, (int64_t)64 * Kp, [=] LBC_LAMBDA(int64_t i) {
    int k = (int)(i % Kp);
    int co = (int)(i / Kp);
    float v = 0.f;
    if (k < 49 * C) {
      int tap = k / C, c = k - tap * C;
      v = w_ref[((int64_t)co * C + c) * 49 + tap];
    }
    wp[i] = v;
  });
}

// all weight packs of a network in ONE launch (table-driven; replaces ~230 tiny launches per step)
//   type 0: [Co][Ci][K][K] -> [Co][K][K][Ci]      type 1: -> [Ci][K][K][Co] (transposed)
//   type 2: block-entry combined [Ci][2Co]        type 3: stem [64][C][7][7] -> [64][Kp] (k = tap*C + c, zero padded)
struct PackEntry {
  int64_t src_off, src2_off;
  void* dst;
  int type, Co, Ci, K, aux;
  int64_t n;
};
struct k_pack_all;
template <class T>
void pack_all(lbc_stream_t s, const float* P, const PackEntry* table, int n_entries) {
  const int64_t LANES = 32768;
  par_for<k_pack_all>(s, (int64_t)n_entries * LANES, [=] LBC_LAMBDA(int64_t t) {
    const PackEntry e = table[t / LANES];
    const float* src = P + e.src_off;
    T* dst = (T*)e.dst;
    for (int64_t i = t % LANES; i < e.n; i += LANES) {
      float v;
      if (e.type == 0) {
        int ci = (int)(i % e.Ci);
        int64_t r = i / e.Ci;
        int kw = (int)(r % e.K);
        r /= e.K;
        int kh = (int)(r % e.K);
        int co = (int)(r / e.K);
        v = src[(((int64_t)co * e.Ci + ci) * e.K + kh) * e.K + kw];
      } else if (e.type == 1) {
        int co = (int)(i % e.Co);
        int64_t r = i / e.Co;
        int kw = (int)(r % e.K);
        r /= e.K;
        int kh = (int)(r % e.K);
        int ci = (int)(r / e.K);
        v = src[(((int64_t)co * e.Ci + ci) * e.K + kh) * e.K + kw];
      } else if (e.type == 2) {
        int k = (int)(i % (2 * e.Co));
        int ci = (int)(i / (2 * e.Co));
        v = k < e.Co ? src[(((int64_t)k * e.Ci + ci) * 3 + 1) * 3 + 1] : (P + e.src2_off)[(int64_t)(k - e.Co) * e.Ci + ci];
      } else {
        int Kp = e.aux, C = e.Ci;
        int k = (int)(i % Kp);
        int co = (int)(i / Kp);
        v = 0.f;
        if (k < 49 * C) {
          int tap = k / C, c = k - tap * C;
          v = src[((int64_t)co * C + c) * 49 + tap];
        }
      }
      stf(dst, i, v);
    }
  });
}

// Same table, walked by (co, ci) PAIRS for the conv layouts: a lane reads the K*K contiguous taps of one (co, ci) pair
// and scatters them into the K*K rows of the packed operand.  type 0: consecutive lanes = consecutive ci -> the source
// is one contiguous run and every store row is coalesced; type 1: consecutive lanes = consecutive co -> 36-byte source
// chunks (1.8x sector amplification instead of 8x for the element walk) and coalesced stores.  No div/mod per element.
struct k_pack_all_pairs;
template <class T>
void pack_all_pairs(lbc_stream_t s, const float* P, const PackEntry* table, int n_entries) {
  const int64_t LANES = 32768;
  par_for<k_pack_all_pairs>(s, (int64_t)n_entries * LANES, [=] LBC_LAMBDA(int64_t t) {
    const PackEntry e = table[t / LANES];
    const float* src = P + e.src_off;
    T* dst = (T*)e.dst;
    if (e.type == 0 || e.type == 1) {
      const int KK = e.K * e.K;
      const int64_t pairs = (int64_t)e.Co * e.Ci;
      for (int64_t q = t % LANES; q < pairs; q += LANES) {
        int co, ci;
        if (e.type == 0) {
          ci = (int)(q % e.Ci);
          co = (int)(q / e.Ci);
        } else {
          co = (int)(q % e.Co);
          ci = (int)(q / e.Co);
        }
        const float* sp = src + ((int64_t)co * e.Ci + ci) * KK;
        if (e.type == 0) {
          T* dp = dst + (int64_t)co * KK * e.Ci + ci;
          for (int tap = 0; tap < KK; ++tap) stf(dp, (int64_t)tap * e.Ci, sp[tap]);
        } else {
          T* dp = dst + (int64_t)ci * KK * e.Co + co;
          for (int tap = 0; tap < KK; ++tap) stf(dp, (int64_t)tap * e.Co, sp[tap]);
        }
      }
      return;
    }
    for (int64_t i = t % LANES; i < e.n; i += LANES) {
      float v;
      if (e.type == 2) {
        int k = (int)(i % (2 * e.Co));
        int ci = (int)(i / (2 * e.Co));
        v = k < e.Co ? src[(((int64_t)k * e.Ci + ci) * 3 + 1) * 3 + 1] : (P + e.src2_off)[(int64_t)(k - e.Co) * e.Ci + ci];
      } else {
        int Kp = e.aux, C = e.Ci;
        int k = (int)(i % Kp);
        int co = (int)(i / Kp);
        v = 0.f;
        if (k < 49 * C) {
          int tap = k / C, c = k - tap * C;
          v = src[((int64_t)co * C + c) * 49 + tap];
        }
      }
      stf(dst, i, v);
    }
  });
}

// ---------------------------------------------------------------------------------------------
// BatchNorm2d, train mode (torch BN as constructed at resnet.py:104, image.py:38,56)
// column statistics over M rows of C channels.  ws: >= 2*P*C doubles.
inline int64_t bn_chunks(int64_t M, int C) {
  int64_t P = (1 << 16) / C;
  if (P < 1) P = 1;
  if (P > M) P = M;
  return P;
}
template <class T>
void bn_stats(lbc_stream_t s, const T* x, int64_t M, int C, float* mean, float* var_biased, double* ws) {
  int64_t P = bn_chunks(M, C);
  int64_t per = cdiv(M, P);
  P = cdiv(M, per);
  double* part = ws;
  double* dmean = ws + P * C;  // C doubles
  par_for<k_bn_sum_part>(s, P * C, [=] LBC_LAMBDA(int64_t i) {
    int c = (int)(i % C);
    int64_t ch = i / C;
    int64_t r0 = ch * per, r1 = r0 + per;
    if (r1 > M) r1 = M;
    double a = 0.0;
    for (int64_t r = r0; r < r1; ++r) a += (double)ldf(x, r * C + c);
    part[i] = a;
  });
  par_for<k_reduce_part>(s, C, [=] LBC_LAMBDA(int64_t c) {
    double a = 0.0;
    for (int64_t ch = 0; ch < P; ++ch) a += part[ch * C + c];
    a /= (double)M;
    dmean[c] = a;
    mean[c] = (float)a;
  });
  par_for<k_bn_var_part>(s, P * C, [=] LBC_LAMBDA(int64_t i) {
    int c = (int)(i % C);
    int64_t ch = i / C;
    int64_t r0 = ch * per, r1 = r0 + per;
    if (r1 > M) r1 = M;
    double mu = dmean[c];
    double a = 0.0;
    for (int64_t r = r0; r < r1; ++r) {
      double d = (double)ldf(x, r * C + c) - mu;
      a += d * d;
    }
    part[i] = a;
  });
  par_for<k_reduce_part>(s, C, [=] LBC_LAMBDA(int64_t c) {
    double a = 0.0;
    for (int64_t ch = 0; ch < P; ++ch) a += part[ch * C + c];
    var_biased[c] = (float)(a / (double)M);
  });
}

// rstd = 1/sqrt(var+eps); running stats: rm <- (1-mom) rm + mom*mean ; rv <- (1-mom) rv + mom*var*M/(M-1)
inline void bn_finalize(lbc_stream_t s, const float* mean, const float* var_biased, int C, int64_t M, float eps,
                        float momentum, float* rstd, float* running_mean, float* running_var) {
  par_for<k_bn_finalize>(s, C, [=] LBC_LAMBDA(int64_t c) {
    float v = var_biased[c];
    rstd[c] = 1.0f / sqrtf(v + eps);
    if (running_mean) {
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean[c];
      float unb = v * ((float)M / (float)(M > 1 ? M - 1 : 1));
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * unb;
    }
  });
}
// same from (sum, sum of squares) accumulated by the fast statistics kernel
inline void bn_finalize_sums(lbc_stream_t s, const float* sums, int C, int64_t M, float eps, float momentum, float* mean,
                             float* rstd, float* running_mean, float* running_var, float* negshift = nullptr) {
  par_for<k_bn_finalize>(s, C, [=] LBC_LAMBDA(int64_t c) {
    double m = (double)sums[c] / (double)M;
    double var = (double)sums[C + c] / (double)M - m * m;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)m;
    rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    double true_mean = negshift ? m - (double)negshift[c] : m;   // statistics were taken on (x + negshift)
    if (negshift) negshift[c] = -(float)true_mean;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)true_mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(var * ((double)M / (double)(M > 1 ? M - 1 : 1)));
  });
}
// bits[i] = (act[8i+j] > 0) in bit j: the mask format bn_apply_kernel emits on the bf16 path (tests build it from an activation)
struct k_mask_bits;
template <class T>
void pack_mask_bits(lbc_stream_t s, const T* act, uint8_t* bits, int64_t n8) {
  par_for<k_mask_bits>(s, n8, [=] LBC_LAMBDA(int64_t i) {
    unsigned b = 0;
    for (int j = 0; j < 8; ++j) b |= (unsigned)(ldf(act, i * 8 + j) > 0.f) << j;
    bits[i] = (uint8_t)b;
  });
}
struct k_rstd_to_var;
inline void rstd_to_var(lbc_stream_t s, const float* rstd, float* var_biased, int C, float eps) {
  par_for<k_rstd_to_var>(s, C, [=] LBC_LAMBDA(int64_t c) { var_biased[c] = 1.0f / (rstd[c] * rstd[c]) - eps; });
}
// eval mode: mean = running_mean, rstd from running_var
inline void bn_eval_stats(lbc_stream_t s, const float* running_mean, const float* running_var, int C, float eps,
                          float* mean, float* rstd, const float* negshift = nullptr) {
  par_for<k_bn_finalize>(s, C, [=] LBC_LAMBDA(int64_t c) {
    mean[c] = running_mean[c] + (negshift ? negshift[c] : 0.f);
    rstd[c] = 1.0f / sqrtf(running_var[c] + eps);
  });
}

// y = gamma*(x-mean)*rstd + beta (+residual) (relu)
template <class T>
void bn_apply(lbc_stream_t s, const T* x, const float* mean, const float* rstd, const float* gamma,
              const float* beta, const T* residual, bool relu, T* y, int64_t M, int C) {
  par_for<k_bn_apply>(s, M * C, [=] LBC_LAMBDA(int64_t i) {
    int c = (int)(i % C);
    float v = (ldf(x, i) - mean[c]) * rstd[c] * gamma[c] + beta[c];
    if (residual) v += ldf(residual, i);
    if (relu) v = v > 0.f ? v : 0.f;
    stf(y, i, v);
  });
}

// BN backward: dbeta = sum dy, dgamma = sum dy*xhat, dx = gamma*rstd*(dy - dbeta/M - xhat*dgamma/M)
// ws: >= 2*P*C doubles
template <class T>
void bn_bwd(lbc_stream_t s, const T* dy, const T* x, const float* mean, const float* rstd, const float* gamma,
            float* dgamma, float* dbeta, T* dx, int64_t M, int C, double* ws) {
  int64_t P = bn_chunks(M, C);
  int64_t per = cdiv(M, P);
  P = cdiv(M, per);
  double* p0 = ws;
  double* p1 = ws + P * C;
  par_for<k_bn_bwd_part>(s, P * C, [=] LBC_LAMBDA(int64_t i) {
    int c = (int)(i % C);
    int64_t ch = i / C;
    int64_t r0 = ch * per, r1 = r0 + per;
    if (r1 > M) r1 = M;
    double a = 0.0, b = 0.0;
    float mu = mean[c], rs = rstd[c];
    for (int64_t r = r0; r < r1; ++r) {
      float g = ldf(dy, r * C + c);
      float xh = (ldf(x, r * C + c) - mu) * rs;
      a += (double)g;
      b += (double)g * (double)xh;
    }
    p0[i] = a;
    p1[i] = b;
  });
  par_for<k_bn_bwd_final>(s, C, [=] LBC_LAMBDA(int64_t c) {
    double a = 0.0, b = 0.0;
    for (int64_t ch = 0; ch < P; ++ch) {
      a += p0[ch * C + c];
      b += p1[ch * C + c];
    }
    dbeta[c] = (float)a;
    dgamma[c] = (float)b;
  });
  if (dx) {
    float invM = 1.0f / (float)M;
    par_for<k_bn_bwd_apply>(s, M * C, [=] LBC_LAMBDA(int64_t i) {
      int c = (int)(i % C);
      float xh = (ldf(x, i) - mean[c]) * rstd[c];
      float v = gamma[c] * rstd[c] * (ldf(dy, i) - dbeta[c] * invM - xh * dgamma[c] * invM);
      stf(dx, i, v);
    });
  }
}

// ---------------------------------------------------------------------------------------------
// MaxPool2d(3,2,1)  (resnet.py:106); idx = kh*3+kw of the first maximum
template <class T>
void maxpool_fwd(lbc_stream_t s, const T* x, T* y, uint8_t* idx, int N, int H, int W, int C, int OH, int OW) {
  int64_t n = (int64_t)N * OH * OW * C;
  par_for<k_maxpool_fwd>(s, n, [=] LBC_LAMBDA(int64_t i) {
    int c = (int)(i % C);
    int64_t p = i / C;
    int ow = (int)(p % OW);
    int64_t q = p / OW;
    int oh = (int)(q % OH);
    int b = (int)(q / OH);
    float best = -INFINITY;
    int bi = 0;
    for (int kh = 0; kh < 3; ++kh) {
      int ih = oh * 2 - 1 + kh;
      if (ih < 0 || ih >= H) continue;
      for (int kw = 0; kw < 3; ++kw) {
        int iw = ow * 2 - 1 + kw;
        if (iw < 0 || iw >= W) continue;
        float v = ldf(x, (((int64_t)b * H + ih) * W + iw) * C + c);
        if (v > best) {
          best = v;
          bi = kh * 3 + kw;
        }
      }
    }
    stf(y, i, best);
    if (idx) idx[i] = (uint8_t)bi;
  });
}
template <class T>
void maxpool_bwd(lbc_stream_t s, const T* dy, const uint8_t* idx, T* dx, int N, int H, int W, int C, int OH,
                 int OW) {
  int64_t n = (int64_t)N * H * W * C;
  par_for<k_maxpool_bwd>(s, n, [=] LBC_LAMBDA(int64_t i) {
    int c = (int)(i % C);
    int64_t p = i / C;
    int iw = (int)(p % W);
    int64_t q = p / W;
    int ih = (int)(q % H);
    int b = (int)(q / H);
    float acc = 0.f;
    for (int kh = 0; kh < 3; ++kh) {
      int t = ih + 1 - kh;
      if (t < 0 || (t & 1)) continue;
      int oh = t >> 1;
      if (oh >= OH) continue;
      for (int kw = 0; kw < 3; ++kw) {
        int u = iw + 1 - kw;
        if (u < 0 || (u & 1)) continue;
        int ow = u >> 1;
        if (ow >= OW) continue;
        int64_t o = (((int64_t)b * OH + oh) * OW + ow) * C + c;
        if (idx[o] == kh * 3 + kw) acc += ldf(dy, o);
      }
    }
    stf(dx, i, acc);
  });
}

// ---------------------------------------------------------------------------------------------
// elementwise helpers
template <class T>
void relu_mask_inplace(lbc_stream_t s, T* g, const T* act, int64_t n) {  // g *= (act > 0)   nn.ReLU(True) bwd
  par_for<k_relu_mask>(s, n, [=] LBC_LAMBDA(int64_t i) {
    if (!(ldf(act, i) > 0.f)) stf(g, i, 0.f);
  });
}
template <class T>
void add_inplace(lbc_stream_t s, T* dst, const T* src, int64_t n) {
  par_for<k_add>(s, n, [=] LBC_LAMBDA(int64_t i) { stf(dst, i, ldf(dst, i) + ldf(src, i)); });
}
// dst += g * (act > 0)
template <class T>
void add_masked_inplace(lbc_stream_t s, T* dst, const T* g, const T* act, int64_t n) {
  par_for<k_add>(s, n, [=] LBC_LAMBDA(int64_t i) {
    if (ldf(act, i) > 0.f) stf(dst, i, ldf(dst, i) + ldf(g, i));
  });
}
// late fusion of speed, image.py:77-79: out[m, 0:Ct] = trunk[m,:], out[m, Ct:Ct+Cs] = speed[n]
template <class T>
void concat_speed(lbc_stream_t s, const T* trunk, const float* speed, T* out, int N, int HW, int Ct, int Cs) {
  int C = Ct + Cs;
  int64_t n = (int64_t)N * HW * C;
  par_for<k_concat_speed>(s, n, [=] LBC_LAMBDA(int64_t i) {
    int c = (int)(i % C);
    int64_t m = i / C;
    float v = c < Ct ? ldf(trunk, m * Ct + c) : speed[m / HW];
    stf(out, i, v);
  });
}
template <class T>
void slice_channels(lbc_stream_t s, const T* src, T* dst, int64_t M, int Csrc, int Cdst) {
  par_for<k_slice>(s, M * Cdst, [=] LBC_LAMBDA(int64_t i) {
    int c = (int)(i % Cdst);
    int64_t m = i / Cdst;
    stf(dst, i, ldf(src, m * Csrc + c));
  });
}
// column sums (deconv bias grad): out[c] = sum_m dy[m,c].  ws >= P*C doubles
template <class T>
void colsum(lbc_stream_t s, const T* dy, int64_t M, int C, float* out, double* ws) {
  int64_t P = bn_chunks(M, C);
  int64_t per = cdiv(M, P);
  P = cdiv(M, per);
  par_for<k_colsum_part>(s, P * C, [=] LBC_LAMBDA(int64_t i) {
    int c = (int)(i % C);
    int64_t ch = i / C;
    int64_t r0 = ch * per, r1 = r0 + per;
    if (r1 > M) r1 = M;
    double a = 0.0;
    for (int64_t r = r0; r < r1; ++r) a += (double)ldf(dy, r * C + c);
    ws[i] = a;
  });
  par_for<k_reduce_part>(s, C, [=] LBC_LAMBDA(int64_t c) {
    double a = 0.0;
    for (int64_t ch = 0; ch < P; ++ch) a += ws[ch * C + c];
    out[c] = (float)a;
  });
}

// ---------------------------------------------------------------------------------------------
// Waypoint heads, image.py:54-60,82-84: 4x [BN(64) -> conv1x1(64->5)+bias -> SpatialSoftmax], stack, select.
// h: [N,HW,C] post-ReLU decoder output; mean/rstd: shared batch stats of h (the four BNs see the same input).
// hp: per head k, packed params: gamma[4][C], beta[4][C], wgt[4][5][C], bias[4][5]
struct HeadParams {
  const float* gamma[4];
  const float* beta[4];
  const float* w[4];     // [5][C]
  const float* bias[4];  // [5]
  const float* mean[4];  // per-head BN statistics (train: all four alias the shared batch stats)
  const float* rstd[4];
};
template <class T>
void head_logits(lbc_stream_t s, const T* h, HeadParams hp, float* logits, int N, int HW, int C) {
  // logits layout [N][20][HW]
  int64_t n = (int64_t)N * 20 * HW;
  par_for<k_head_logits>(s, n, [=] LBC_LAMBDA(int64_t i) {
    int pix = (int)(i % HW);
    int64_t t = i / HW;
    int kj = (int)(t % 20);
    int b = (int)(t / 20);
    int k = kj / 5, j = kj % 5;
    const T* hpix = h + ((int64_t)b * HW + pix) * C;
    const float* g = hp.gamma[k];
    const float* be = hp.beta[k];
    const float* w = hp.w[k] + j * C;
    const float* mean = hp.mean[k];
    const float* rstd = hp.rstd[k];
    float acc = hp.bias[k][j];
    for (int c = 0; c < C; ++c) {
      float z = (ldf(hpix, c) - mean[c]) * rstd[c] * g[c] + be[c];
      acc += w[c] * z;
    }
    logits[i] = acc;
  });
}
// SpatialSoftmax (common.py:136-152): softmax over HW, expectation of pos_x (along W) and pos_y (along H)
inline void head_softmax(lbc_stream_t s, const float* logits, float* rowmax, float* rowsum, float* preds, int N,
                         int H, int W) {
  int HW = H * W;
  par_for<k_head_softmax>(s, (int64_t)N * 20, [=] LBC_LAMBDA(int64_t r) {
    const float* l = logits + r * HW;
    float m = -INFINITY;
    for (int p = 0; p < HW; ++p) m = l[p] > m ? l[p] : m;
    double se = 0.0, sx = 0.0, sy = 0.0;
    for (int p = 0; p < HW; ++p) {
      float e = expf(l[p] - m);
      int hh = p / W, ww = p % W;
      float px = W > 1 ? (float)(-1.0 + (double)ww * (2.0 / (double)(W - 1))) : -1.f;
      float py = H > 1 ? (float)(-1.0 + (double)hh * (2.0 / (double)(H - 1))) : -1.f;
      se += (double)e;
      sx += (double)e * (double)px;
      sy += (double)e * (double)py;
    }
    rowmax[r] = m;
    rowsum[r] = (float)se;
    preds[r * 2 + 0] = (float)(sx / se);
    preds[r * 2 + 1] = (float)(sy / se);
  });
}
// select_branch (common.py:29-35): pred[n,j,:] = sum_k onehot[n,k]*preds[n,k,j,:]
inline void head_select(lbc_stream_t s, const float* preds, const float* onehot, float* pred, int N) {
  par_for<k_head_select>(s, (int64_t)N * 10, [=] LBC_LAMBDA(int64_t i) {
    int e = (int)(i % 10);
    int b = (int)(i / 10);
    float a = 0.f;
    for (int k = 0; k < 4; ++k) a += onehot[b * 4 + k] * preds[((int64_t)b * 4 + k) * 10 + e];
    pred[i] = a;
  });
}
// dlogit = softmax * ((x-Ex) gx + (y-Ey) gy), g = d_preds + onehot*d_pred   
inline void head_dlogits(lbc_stream_t s, const float* logits, const float* rowmax, const float* rowsum,
                         const float* preds, const float* onehot, const float* d_pred, const float* d_preds,
                         float* dlogits, int N, int H, int W) {
  int HW = H * W;
  par_for<k_head_dlogits>(s, (int64_t)N * 20 * HW, [=] LBC_LAMBDA(int64_t i) {
    int p = (int)(i % HW);
    int64_t r = i / HW;
    int kj = (int)(r % 20);
    int b = (int)(r / 20);
    int k = kj / 5, j = kj % 5;
    float gx = 0.f, gy = 0.f;
    if (d_preds) {
      gx += d_preds[r * 2];
      gy += d_preds[r * 2 + 1];
    }
    if (d_pred) {
      float oh = onehot[b * 4 + k];
      gx += oh * d_pred[((int64_t)b * 5 + j) * 2];
      gy += oh * d_pred[((int64_t)b * 5 + j) * 2 + 1];
    }
    int hh = p / W, ww = p % W;
    float px = W > 1 ? (float)(-1.0 + (double)ww * (2.0 / (double)(W - 1))) : -1.f;
    float py = H > 1 ? (float)(-1.0 + (double)hh * (2.0 / (double)(H - 1))) : -1.f;
    float wgt = expf(logits[i] - rowmax[r]) / rowsum[r];
    dlogits[i] = wgt * ((px - preds[r * 2]) * gx + (py - preds[r * 2 + 1]) * gy);
  });
}
// S1[kj][c] = sum_{n,pix} dlogit*hhat_c ; S0[kj] = sum dlogit.  ws >= P*20*(C+1) doubles. out S: [20][C+1] (last = S0)
template <class T>
void head_s(lbc_stream_t s, const float* dlogits, const T* h, const float* mean, const float* rstd, double* S,
            int N, int HW, int C, double* ws) {
  int C1 = C + 1;
  int64_t P = N;  // one chunk per image
  par_for<k_head_s_part>(s, P * 20 * C1, [=] LBC_LAMBDA(int64_t i) {
    int c = (int)(i % C1);
    int64_t t = i / C1;
    int kj = (int)(t % 20);
    int b = (int)(t / 20);
    const float* dl = dlogits + ((int64_t)b * 20 + kj) * HW;
    double a = 0.0;
    if (c == C) {
      for (int p = 0; p < HW; ++p) a += (double)dl[p];
    } else {
      float mu = mean[c], rs = rstd[c];
      for (int p = 0; p < HW; ++p) a += (double)dl[p] * (double)((ldf(h, ((int64_t)b * HW + p) * C + c) - mu) * rs);
    }
    ws[i] = a;
  });
  par_for<k_reduce_part>(s, 20 * C1, [=] LBC_LAMBDA(int64_t i) {
    double a = 0.0;
    for (int64_t b = 0; b < P; ++b) a += ws[b * 20 * C1 + i];
    S[i] = a;
  });
}
struct HeadGrads {
  float* dgamma[4];
  float* dbeta[4];
  float* dw[4];
  float* dbias[4];
};
// dW_k[j,c] = gamma_kc S1 + beta_kc S0 ; dbias = S0 ; dgamma_kc = sum_j W S1 ; dbeta_kc = sum_j W S0
inline void head_param_grads(lbc_stream_t s, const double* S, HeadParams hp, HeadGrads hg, int C) {
  int C1 = C + 1;
  par_for<k_head_param_grads>(s, 4 * C1, [=] LBC_LAMBDA(int64_t i) {
    int c = (int)(i % C1);
    int k = (int)(i / C1);
    if (c == C) {
      for (int j = 0; j < 5; ++j) hg.dbias[k][j] = (float)S[(k * 5 + j) * C1 + C];
      return;
    }
    double dg = 0.0, db = 0.0;
    for (int j = 0; j < 5; ++j) {
      double s1 = S[(k * 5 + j) * C1 + c], s0 = S[(k * 5 + j) * C1 + C];
      double w = hp.w[k][j * C + c];
      hg.dw[k][j * C + c] = (float)(hp.gamma[k][c] * s1 + hp.beta[k][c] * s0);
      dg += w * s1;
      db += w * s0;
    }
    hg.dgamma[k][c] = (float)dg;
    hg.dbeta[k][c] = (float)db;
  });
}
// dh[n,pix,c] = rstd_c sum_k gamma_kc (dz_kc - dbeta_kc/M - hhat_c dgamma_kc/M), dz_kc = sum_j W_k[j,c] dlogit[n,kj,pix]
template <class T>
void head_dh(lbc_stream_t s, const float* dlogits, const T* h, const float* mean, const float* rstd, HeadParams hp,
             HeadGrads hg, T* dh, int N, int HW, int C) {
  float invM = 1.0f / (float)((int64_t)N * HW);
  par_for<k_head_dh>(s, (int64_t)N * HW * C, [=] LBC_LAMBDA(int64_t i) {
    int c = (int)(i % C);
    int64_t m = i / C;
    int p = (int)(m % HW);
    int b = (int)(m / HW);
    float xh = (ldf(h, i) - mean[c]) * rstd[c];
    float acc = 0.f;
    for (int k = 0; k < 4; ++k) {
      float dz = 0.f;
      for (int j = 0; j < 5; ++j) dz += hp.w[k][j * C + c] * dlogits[((int64_t)b * 20 + k * 5 + j) * HW + p];
      acc += hp.gamma[k][c] * (dz - hg.dbeta[k][c] * invM - xh * hg.dgamma[k][c] * invM);
    }
    stf(dh, i, acc * rstd[c]);
  });
}

// ---------------------------------------------------------------------------------------------
// losses (tiny tensors; one thread per sample)
// phase 0 target: training/train_image_phase0.py:54-79 (CoordConverter): teacher map coords [-1,1] ->
// image pixels, pinhole projection in float64, clipped to the image.
inline void phase0_target(lbc_stream_t s, const float* t_pred, float* target_px, int64_t count, float w, float h,
                          float fov_deg, float world_y, float fixed_offset) {
  double f = (double)w / (2.0 * tan((double)fov_deg * 3.14159265358979323846 / 360.0));
  par_for<k_phase0_target>(s, count, [=] LBC_LAMBDA(int64_t i) {
    const float CROP = 192.f, PPM = 5.f;
    float tx = (t_pred[i * 2] + 1.f) * CROP / 2.f;
    float ty = (t_pred[i * 2 + 1] + 1.f) * CROP / 2.f;
    ty = CROP - ty;
    tx -= CROP / 2.f;
    tx = tx / PPM;
    ty = ty / PPM;
    ty += fixed_offset;
    double X = (double)tx, Z = (double)ty;
    double u = f * X / Z + (double)w / 2.0;
    double v = f * (double)world_y / Z + (double)h / 2.0;
    u = u < 0.0 ? 0.0 : (u > (double)w ? (double)w : u);
    v = v < 0.0 ? 0.0 : (v > (double)h ? (double)h : v);
    target_px[i * 2] = (float)u;
    target_px[i * 2 + 1] = (float)v;
  });
}
// loss_b[n] = mean_D | a*sa + ta - (b*sb_xy + tb) | ; da = gout[n] * sign(.) * sa / D   (gout null -> 1/N)
// covers train_image_phase0.py:86-89, train_image_phase1.py:66-70, train_birdview.py:33-54 (l1)
inline void l1_loss(lbc_stream_t s, const float* a, const float* b, int N, int D, float sa, float ta, float sbx,
                    float sby, float tb, const float* gout, float* loss_b, float* da) {
  par_for<k_l1_loss>(s, N, [=] LBC_LAMBDA(int64_t n) {
    double acc = 0.0;
    float go = gout ? gout[n] : 1.0f / (float)N;
    for (int d = 0; d < D; ++d) {
      float av = a[n * D + d] * sa + ta;
      float bv = b[n * D + d] * ((d & 1) ? sby : sbx) + tb;
      float diff = av - bv;
      acc += (double)fabsf(diff);
      if (da) da[n * D + d] = (diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f)) * go * sa / (float)D;
    }
    if (loss_b) loss_b[n] = (float)(acc / (double)D);
  });
}
// phase-2 (DAgger) replay-buffer weight, training/phase2_utils.py:50-59 (get_weight): per sample
//   mean_j( (|learner - teacher| * [0.7, 0.3]).sum(xy) * 0.7^j ),  learner already in map coords [-1,1]
struct k_phase2_weight;
inline void phase2_weight(lbc_stream_t s, const float* learner, const float* teacher, float* weight, int N) {
  par_for<k_phase2_weight>(s, N, [=] LBC_LAMBDA(int64_t n) {
    float acc = 0.f, decay = 1.f;
    for (int j = 0; j < 5; ++j) {
      float dx = fabsf(learner[(n * 5 + j) * 2] - teacher[(n * 5 + j) * 2]) * 0.7f;
      float dy = fabsf(learner[(n * 5 + j) * 2 + 1] - teacher[(n * 5 + j) * 2 + 1]) * 0.3f;
      acc += (dx + dy) * decay;
      decay *= 0.7f;
    }
    weight[n] = acc / 5.f;
  });
}
// phase-1 CoordConverter (train_image_phase1.py:43-64): image [-1,1] -> map pixels, and its backward
inline float focal_px(float w, float fov_deg) {
  return (float)((double)w / (2.0 * tan((double)fov_deg * 3.14159265358979323846 / 360.0)));
}
inline void phase1_convert_fwd(lbc_stream_t s, const float* p, float* out, int64_t count, float w, float h,
                               float fov_deg, float world_y, float fixed_offset) {
  float f = focal_px(w, fov_deg);
  par_for<k_phase1_fwd>(s, count, [=] LBC_LAMBDA(int64_t i) {
    const float CROP = 192.f, PPM = 5.f;
    float cx = (p[i * 2] + 1.f) * w / 2.f, cy = (p[i * 2 + 1] + 1.f) * h / 2.f;
    float xt = (cx - w / 2.f) / f, yt = (cy - h / 2.f) / f;
    float wz = world_y / yt, wx = wz * xt;
    out[i * 2] = wx * PPM + CROP / 2.f;
    out[i * 2 + 1] = CROP - wz * PPM + fixed_offset * PPM;
  });
}
inline void phase1_convert_bwd(lbc_stream_t s, const float* p, const float* dout, float* dp, int64_t count, float w,
                               float h, float fov_deg, float world_y, float fixed_offset) {
  float f = focal_px(w, fov_deg);
  par_for<k_phase1_bwd>(s, count, [=] LBC_LAMBDA(int64_t i) {
    const float PPM = 5.f;
    float cx = (p[i * 2] + 1.f) * w / 2.f, cy = (p[i * 2 + 1] + 1.f) * h / 2.f;
    float xt = (cx - w / 2.f) / f, yt = (cy - h / 2.f) / f;
    float wz = world_y / yt;
    // out_x = PPM*wz*xt + c ; out_y = c' - PPM*wz
    float dox = dout[i * 2], doy = dout[i * 2 + 1];
    float d_xt = dox * PPM * wz;
    float d_wz = dox * PPM * xt - doy * PPM;
    float d_yt = d_wz * (-world_y / (yt * yt));
    dp[i * 2] = d_xt * (w / 2.f) / f;
    dp[i * 2 + 1] = d_yt * (h / 2.f) / f;
  });
}

// ---------------------------------------------------------------------------------------------
// torch.optim.Adam (torch 2.11 single-tensor semantics), flat range.
inline void adam(lbc_stream_t s, float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1,
                 float b2, float eps, int step, float grad_scale) {
  double bc1 = 1.0 - pow((double)b1, (double)step);
  double bc2 = 1.0 - pow((double)b2, (double)step);
  float step_size = (float)((double)lr / bc1);
  float bc2_sqrt = (float)sqrt(bc2);
  par_for<k_adam>(s, n, [=] LBC_LAMBDA(int64_t i) {
    float gi = g[i] * grad_scale;
    float mi = m[i] + (1.f - b1) * (gi - m[i]);
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
  });
}

template <class TS, class TD>
void cast(lbc_stream_t s, const TS* src, TD* dst, int64_t n) {
  par_for<k_cast>(s, n, [=] LBC_LAMBDA(int64_t i) { stf(dst, i, ldf(src, i)); });
}

}  // namespace ref
}  // namespace lbc
