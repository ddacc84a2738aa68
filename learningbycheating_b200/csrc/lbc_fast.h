// lbc_fast.h -- hooks from the graph executor into the sm_100a fast kernels (tcgen05 implicit-GEMM
// convolutions, fused BN / head kernels).  Each hook returns false when it does not handle the
// call (shape / dtype not covered, or host-emulation build) and the executor then runs the
// correctness-first kernel of lbc_ref_ops.h instead.
#pragma once
#include "lbc_common.h"
#include "lbc_net.h"

namespace lbc {
namespace fast {

// global switch (tests flip it to compare fast kernels against the correctness-first ones)
bool enabled();
void set_enabled(bool on);

bool conv_fwd_bf16(const ConvL& c, const bf16* x, bf16* y, int B, lbc_stream_t s);

template <class T>
inline bool conv_fwd(const ConvL& c, const T* x, T* y, int B, lbc_stream_t s) {
  (void)c; (void)x; (void)y; (void)B; (void)s;
  return false;
}
template <>
inline bool conv_fwd<bf16>(const ConvL& c, const bf16* x, bf16* y, int B, lbc_stream_t s) {
  if (!enabled()) return false;
  return conv_fwd_bf16(c, x, y, B, s);
}

bool conv_dgrad_bf16(const ConvL& c, const bf16* dy, bf16* dx, int B, const float* bias_ci, bool relu, lbc_stream_t s);

template <class T>
inline bool conv_dgrad(const ConvL& c, const T* dy, T* dx, int B, const float* bias_ci, bool relu, lbc_stream_t s) {
  (void)c; (void)dy; (void)dx; (void)B; (void)bias_ci; (void)relu; (void)s;
  return false;
}
template <>
inline bool conv_dgrad<bf16>(const ConvL& c, const bf16* dy, bf16* dx, int B, const float* bias_ci, bool relu,
                             lbc_stream_t s) {
  if (!enabled()) return false;
  return conv_dgrad_bf16(c, dy, dx, B, bias_ci, relu, s);
}

bool conv_wgrad_bf16(const ConvL& c, const bf16* x, const bf16* dy, float* dw_ref, int B, float* scratch,
                     int64_t scratch_floats, lbc_stream_t s);
template <class T>
inline bool conv_wgrad(const ConvL&, const T*, const T*, float*, int, float*, int64_t, lbc_stream_t) {
  return false;
}
template <>
inline bool conv_wgrad<bf16>(const ConvL& c, const bf16* x, const bf16* dy, float* dw_ref, int B, float* scratch,
                             int64_t scratch_floats, lbc_stream_t s) {
  if (!enabled()) return false;
  return conv_wgrad_bf16(c, x, dy, dw_ref, B, scratch, scratch_floats, s);
}

// ---- stem (7x7/s2, C_in = 3 or 7): explicit im2col (fused NCHW fp32 -> bf16 + RGB normalisation) feeding the
// same tcgen05 GEMM kernels as a 1x1 convolution over the [B,OH,OW,Kp] column tensor (Kp = 49*C padded to 64).
bool stem_im2col_bf16(const float* img, bf16* col, int B, int C, int H, int W, int OH, int OW, int Kp, bool normalize,
                      lbc_stream_t s);
bool stem_pack_weight_bf16(const float* w_ref, bf16* wp, int C, int Kp, lbc_stream_t s);      // [64][C][7][7] -> [64][Kp]
bool stem_unpack_wgrad(const float* dw_col, float* dw_ref, int C, int Kp, lbc_stream_t s);    // [64][Kp] -> [64][C][7][7]

}  // namespace fast
}  // namespace lbc
