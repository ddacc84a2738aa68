// lbc_bn_tail.h -- finalisation of train-mode BatchNorm statistics, shared by bn_finalize_kernel (lbc_fast_elem.cu) and the
// epilogues of the convolution kernels (lbc_fast_conv.cu).  Device code only (the CUDA build).
//
// The forward chain conv -> bn_finalize_kernel -> bn_apply_kernel made every BatchNorm wait for a one-wave kernel of a few
// blocks between two full-grid kernels: measured in the pipeline (all bn_finalize launches skipped) 0.65 ms of a 12.6 ms
// step for 48 launches, 13.6 us each -- two kernel boundaries plus a chain of dependent L2 round trips.  "Tail" mode moves
// the finalisation into the convolution itself: every CTA writes its partial row as before, fences, and takes a ticket; the
// CTA that draws the last ticket sums the <= 148 rows IN ROW ORDER (the result does not depend on which CTA is last, so the
// statistics stay bitwise reproducible) and runs the same per-channel arithmetic as bn_finalize_kernel.
#pragma once
#include "lbc_common.h"

#ifndef LBC_HOST_EMU
namespace lbc {
namespace fast {

struct BnFinalizeArgs {
  const float* partial;   // [P][2C]
  int P, C;
  double inv_m, unbias;   // 1 / M and M / (M - 1)
  const float* gamma;
  const float* beta;
  float eps, momentum;
  float* running_mean;
  float* running_var;
  float* saved_mean;
  float* saved_rstd;
  float* negshift;        // may be null
  float* scsh;            // out: [2C] = scale | shift
  float* sums;            // out (optional): [2C] raw column sums
};
// t0 = sum, t1 = sum of squares of channel c over the M stored (shifted) values
__device__ __forceinline__ void bn_finalize_channel(const BnFinalizeArgs& a, int c, float t0, float t1) {
  if (a.sums) {
    a.sums[c] = t0;
    a.sums[a.C + c] = t1;
  }
  // double only where it matters (sum / M and the E[x^2] - mean^2 cancellation), with multiplications by host-computed
  // reciprocals; 1/sqrt = float rsqrt + one Newton step in double (the software double division / square root sequences
  // made this kernel 3 us slower than the plain column sum it replaced)
  const double m = (double)t0 * a.inv_m;
  double var = (double)t1 * a.inv_m - m * m;
  if (var < 0.0) var = 0.0;
  const double v = var + (double)a.eps;
  double r = (double)rsqrtf((float)v);
  r = r * (1.5 - 0.5 * v * r * r);
  const float mean = (float)m;
  const float rstd = (float)r;
  a.saved_mean[c] = mean;
  a.saved_rstd[c] = rstd;
  const float true_mean = a.negshift ? (float)(m - (double)a.negshift[c]) : mean;
  if (a.negshift) a.negshift[c] = -true_mean;   // centring estimate for the next forward pass
  a.running_mean[c] = (1.f - a.momentum) * a.running_mean[c] + a.momentum * true_mean;
  const double unb = var * a.unbias;
  a.running_var[c] = (1.f - a.momentum) * a.running_var[c] + a.momentum * (float)unb;
  const float sc = a.gamma[c] * rstd;
  a.scsh[c] = sc;
  a.scsh[a.C + c] = a.beta[c] - mean * sc;
}

struct BnTail {
  int enabled;          // 0: the kernel only writes its partial rows
  unsigned* counter;    // device word, zero between launches (the last CTA resets it)
  BnFinalizeArgs a;     // a.partial / a.P are taken from the kernel (its statistics rows, one per CTA)
};
// Called by the 128 epilogue threads of a convolution CTA (named barrier 1) AFTER they stored this CTA's partial row.
// e = 0..127, flag = one shared-memory word.
__device__ __forceinline__ void bn_tail_run(const BnTail& t, const float* partial, int e, uint32_t* flag) {
  if (!t.enabled) return;
  __threadfence();                                   // this thread's part of the row, device-wide, before the ticket
  asm volatile("bar.sync 1, 128;" ::: "memory");
  if (e == 0) *flag = (atomicAdd(t.counter, 1u) == gridDim.x - 1) ? 1u : 0u;
  asm volatile("bar.sync 1, 128;" ::: "memory");
  if (*flag == 0u) return;
  __threadfence();
  const int C = t.a.C, P = (int)gridDim.x;
  const int64_t rs = (int64_t)2 * C;
  for (int c = e; c < C; c += 128) {
    const float* p0 = partial + c;
    float t0 = 0.f, t1 = 0.f;
    for (int r = 0; r < P; r += 32) {                // 64 independent L2 reads in flight, then added in row order
      float v0[32], v1[32];
#pragma unroll
      for (int u = 0; u < 32; ++u) {
        const bool ok = r + u < P;
        v0[u] = ok ? __ldcg(p0 + (int64_t)(r + u) * rs) : 0.f;
        v1[u] = ok ? __ldcg(p0 + (int64_t)(r + u) * rs + C) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 32; ++u) {
        t0 += v0[u];
        t1 += v1[u];
      }
    }
    bn_finalize_channel(t.a, c, t0, t1);
  }
  if (e == 0) *t.counter = 0u;
}

// host side (lbc_fast_conv.cu): the NEXT statistics-emitting convolution launch carries this tail (one shot)
void bn_tail_arm(const BnFinalizeArgs& a);
void bn_tail_disarm();
bool bn_tail_take(BnTail* out, int C);   // launchers: fills *out (enabled or not) for a kernel whose statistics cover C channels
bool bn_tail_fired();                    // true when the last armed tail was attached to a launch (reset by bn_tail_arm)

}  // namespace fast
}  // namespace lbc
#endif
