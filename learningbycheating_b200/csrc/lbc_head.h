// lbc_head.h -- the four waypoint heads as one reusable op (image.py:54-60,82-84; common.py:136-152):
// 4 x [BatchNorm2d(64) -> Conv2d(64,5,1)+bias -> SpatialSoftmax] over the same decoder output, forward and backward.
// Used by the graph executor (lbc_net.cu) and by the single-op entry point lbc_op_head (lbc_capi.cu), so the parity
// test of the head kernels exercises exactly the call sequence the training step runs.
#pragma once
#include <type_traits>

#include "lbc_fast.h"
#include "lbc_ref_ops.h"

namespace lbc {

struct HeadCtx {
  int H = 0, W = 0;                                          // head feature map (40x96 student, 48x48 teacher)
  float *logits = nullptr, *dlogits = nullptr;               // [N][20][HW]
  float *rowmax = nullptr, *rowsum = nullptr, *preds = nullptr;   // [N*20], [N*20], [N][4][5][2]
  double* S = nullptr;                                       // [20][65] moment matrix of the backward pass
  float* fold = nullptr;                                     // >= 1300 + 128 floats: folded BN+1x1 map, dh coefficients
  float* bn_sums = nullptr;                                  // >= 128 floats
  double* ws_d = nullptr;                                    // >= max(N*20*65, bn_chunks*64) doubles (correctness-first path)
  const float *gamma[4], *beta[4], *w[4], *bias[4];          // parameters of head k
  float *mean[4], *rstd[4], *var0 = nullptr;                 // saved statistics (train: the four BNs share index 0's)
  float *rm[4], *rv[4];                                      // running buffers of head k
  bool fast_used = false, mask_fused = false;

  ref::HeadParams params(bool train) const {
    ref::HeadParams hp;
    for (int k = 0; k < 4; ++k) {
      hp.gamma[k] = gamma[k];
      hp.beta[k] = beta[k];
      hp.w[k] = w[k];
      hp.bias[k] = bias[k];
      hp.mean[k] = train ? mean[0] : mean[k];
      hp.rstd[k] = train ? rstd[0] : rstd[k];
    }
    return hp;
  }
};

// hfeat: [N][H*W][64] post-ReLU decoder output.  Leaves logits / rowmax / rowsum / preds in the context.
template <class T>
void head_forward(HeadCtx& h, const T* hfeat, int N, bool train, float eps, float momentum, lbc_stream_t s) {
  const int HW = h.H * h.W;
  const int64_t M = (int64_t)N * HW;
  bool stats_done = false;
  if (train && std::is_same<T, bf16>::value && fast::enabled()) {
    if (fast::bn_stats_bf16((const bf16*)hfeat, M, 64, h.bn_sums, s)) {
      for (int k = 0; k < 4; ++k)   // same statistics, four sets of running buffers (image.py:56)
        ref::bn_finalize_sums(s, h.bn_sums, 64, M, eps, momentum, h.mean[0], h.rstd[k], h.rm[k], h.rv[k]);
      stats_done = true;
    }
  }
  if (stats_done) {
  } else if (train) {
    ref::bn_stats<T>(s, hfeat, M, 64, h.mean[0], h.var0, h.ws_d);
    for (int k = 0; k < 4; ++k) ref::bn_finalize(s, h.mean[0], h.var0, 64, M, eps, momentum, h.rstd[k], h.rm[k], h.rv[k]);
  } else {
    for (int k = 0; k < 4; ++k) ref::bn_eval_stats(s, h.rm[k], h.rv[k], 64, eps, h.mean[k], h.rstd[k]);
  }
  h.fast_used = false;
  if (std::is_same<T, bf16>::value)
    h.fast_used = fast::head_forward_bf16((const bf16*)hfeat, h.params(train), h.fold, h.logits, h.rowmax, h.rowsum, h.preds, N,
                                          h.H, h.W, s);
  if (!h.fast_used) {
    ref::head_logits<T>(s, hfeat, h.params(train), h.logits, N, HW, 64);
    ref::head_softmax(s, h.logits, h.rowmax, h.rowsum, h.preds, N, h.H, h.W);
  }
}

// upstream gradients d_pred [N,5,2] / d_preds [N,4,5,2] (either may be null) -> parameter gradients hg and dh [N][HW][64].
// mask_fused (fast path): dh already carries the decoder's final ReLU mask (hfeat > 0).
template <class T>
void head_backward(HeadCtx& h, const T* hfeat, const float* onehot, const float* d_pred, const float* d_preds,
                   ref::HeadGrads hg, T* dh, int N, lbc_stream_t s) {
  const int HW = h.H * h.W;
  ref::HeadParams hp = h.params(true);
  if (!(h.fast_used && fast::head_dlogits_f32(h.logits, h.rowmax, h.rowsum, h.preds, onehot, d_pred, d_preds, h.dlogits, N, h.H,
                                              h.W, s)))
    ref::head_dlogits(s, h.logits, h.rowmax, h.rowsum, h.preds, onehot, d_pred, d_preds, h.dlogits, N, h.H, h.W);
  h.mask_fused = false;
  bool s_done = false;
  if (std::is_same<T, bf16>::value && h.fast_used)
    s_done = fast::head_backward_s_bf16(h.dlogits, (const bf16*)hfeat, h.mean[0], h.rstd[0], h.S, N, HW, s);
  if (!s_done) ref::head_s<T>(s, h.dlogits, hfeat, h.mean[0], h.rstd[0], h.S, N, HW, 64, h.ws_d);
  ref::head_param_grads(s, h.S, hp, hg, 64);
  if (s_done)
    h.mask_fused = fast::head_backward_dh_bf16(h.dlogits, (const bf16*)hfeat, hp, hg, h.fold, h.fold + 1300, (bf16*)dh, N, HW, s);
  if (!h.mask_fused) ref::head_dh<T>(s, h.dlogits, hfeat, h.mean[0], h.rstd[0], hp, hg, dh, N, HW, 64);
}

}  // namespace lbc
