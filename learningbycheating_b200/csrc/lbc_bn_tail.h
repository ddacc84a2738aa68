// lbc_bn_tail.h -- finalisation of train-mode BatchNorm statistics, shared by bn_finalize_kernel (lbc_fast_elem.cu) and the
// epilogues of the convolution kernels (lbc_fast_conv.cu).  Device code only (the CUDA build).
//
// The forward chain conv -> bn_finalize_kernel -> bn_apply_kernel made every BatchNorm wait for a one-wave kernel of a few
// blocks between two full-grid kernels: measured in the pipeline (all bn_finalize launches skipped) 0.65 ms of a 12.6 ms
// step for 48 launches, 13.6 us each -- two kernel boundaries plus a chain of dependent L2 round trips.  "Tail" mode moves
// the finalisation into the convolution itself: every CTA writes its partial row as before, fences, and meets the others at
// a grid barrier; then the channels are dealt to the warps of ALL CTAs, each summing the <= 148 rows of its channel in a
// fixed order (the statistics stay bitwise reproducible) and running the same per-channel arithmetic as bn_finalize_kernel.
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
// per-channel inputs of the finalisation, loaded BEFORE the statistics are summed: behind the stores of the finalisation the
// compiler cannot hoist them (possible aliasing), and five dependent L2 round trips were a third of bn_finalize_kernel's time
struct BnChannelIn {
  float gamma, beta, running_mean, running_var, negshift;
};
__device__ __forceinline__ BnChannelIn bn_channel_in(const BnFinalizeArgs& a, int c) {
  BnChannelIn in;
  in.gamma = a.gamma[c];
  in.beta = a.beta[c];
  in.running_mean = a.running_mean[c];
  in.running_var = a.running_var[c];
  in.negshift = a.negshift ? a.negshift[c] : 0.f;
  return in;
}
// t0 = sum, t1 = sum of squares of channel c over the M stored (shifted) values
__device__ __forceinline__ void bn_finalize_channel(const BnFinalizeArgs& a, int c, float t0, float t1, const BnChannelIn& in) {
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
  const float true_mean = a.negshift ? (float)(m - (double)in.negshift) : mean;
  if (a.negshift) a.negshift[c] = -true_mean;   // centring estimate for the next forward pass
  a.running_mean[c] = (1.f - a.momentum) * in.running_mean + a.momentum * true_mean;
  const double unb = var * a.unbias;
  a.running_var[c] = (1.f - a.momentum) * in.running_var + a.momentum * (float)unb;
  const float sc = in.gamma * rstd;
  a.scsh[c] = sc;
  a.scsh[a.C + c] = in.beta - mean * sc;
}

struct BnTail {
  int enabled;          // 0: the kernel only writes its partial rows
  unsigned* counter;    // two device words: arrivals (zero between launches) | barrier generation
  BnFinalizeArgs a;     // a.partial / a.P are taken from the kernel (its statistics rows, one per CTA)
};
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Called by the 128 epilogue threads of a convolution CTA (named barrier 1) AFTER they stored this CTA's partial row.
// e = 0..127.  Two steps:
//   1. grid barrier (sense reversing, one spinning thread per CTA).  Legal here: the grid is persistent with at most one CTA
//      (or CTA pair) per SM, sized by the occupancy query, so every CTA is resident or becomes resident once foreign CTAs
//      drain; a PDL successor cannot take its place, it is only launched after ALL CTAs of this grid have started.
//   2. the channels are dealt round-robin to (CTA, warp): a warp sums the <= 160 rows of its channel -- lane l takes rows
//      l, l + 32, ... in order, then a fixed butterfly, so the result is reproducible -- and lane 0 finalises it.  One L2
//      round trip for the rows and the per-channel inputs together, whatever the channel count (the first version let the
//      CTA that drew the last ticket do everything: 15 us per convolution, slower than the kernel it replaced).
__device__ __forceinline__ void bn_tail_run(const BnTail& t, const float* partial, int e) {
  if (!t.enabled) return;
  __threadfence();                                   // this thread's part of the row, device-wide, before the arrival
  asm volatile("bar.sync 1, 128;" ::: "memory");
  if (e == 0) {
    unsigned* gen = t.counter + 1;
    const unsigned g0 = ld_acquire_gpu(gen);         // (before arriving: the generation cannot advance without this CTA)
    if (atomicAdd(t.counter, 1u) == gridDim.x - 1) {
      t.counter[0] = 0u;
      __threadfence();
      st_release_gpu(gen, g0 + 1u);
    } else {
      while (ld_acquire_gpu(gen) == g0) __nanosleep(32);
    }
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");
  const int C = t.a.C, P = (int)gridDim.x;
  const int64_t rs = (int64_t)2 * C;
  const int lane = e & 31;
  for (int c = (int)blockIdx.x * 4 + (e >> 5); c < C; c += 4 * P) {
    BnChannelIn in;
    if (lane == 0) in = bn_channel_in(t.a, c);
    const float* p0 = partial + c;
    float v0[5], v1[5];
#pragma unroll
    for (int u = 0; u < 5; ++u) {
      const int r = lane + 32 * u;
      const bool ok = r < P;
      v0[u] = ok ? __ldcg(p0 + (int64_t)r * rs) : 0.f;
      v1[u] = ok ? __ldcg(p0 + (int64_t)r * rs + C) : 0.f;
    }
    float t0 = 0.f, t1 = 0.f;
#pragma unroll
    for (int u = 0; u < 5; ++u) {
      t0 += v0[u];
      t1 += v1[u];
    }
    for (int r = lane + 160; r < P; r += 32) {   // (grids beyond 160 CTAs: none today)
      t0 += __ldcg(p0 + (int64_t)r * rs);
      t1 += __ldcg(p0 + (int64_t)r * rs + C);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      t0 += __shfl_xor_sync(0xffffffffu, t0, o);
      t1 += __shfl_xor_sync(0xffffffffu, t1, o);
    }
    if (lane == 0) bn_finalize_channel(t.a, c, t0, t1, in);
  }
}

// host side (lbc_fast_conv.cu): the NEXT statistics-emitting convolution launch carries this tail (one shot)
void bn_tail_arm(const BnFinalizeArgs& a);
void bn_tail_disarm();
bool bn_tail_take(BnTail* out, int C);   // launchers: fills *out (enabled or not) for a kernel whose statistics cover C channels
bool bn_tail_fired();                    // true when the last armed tail was attached to a launch (reset by bn_tail_arm)

}  // namespace fast
}  // namespace lbc
#endif
