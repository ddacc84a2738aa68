// lbc_capi.cu -- extern "C" surface declared in include/lbc_b200.h.
#include <mutex>

#include "../../include/lbc_b200.h"
#include "lbc_fast.h"
#include "lbc_net.h"
#include "lbc_ref_ops.h"

using namespace lbc;

struct lbc_net {
  std::unique_ptr<NetBase> impl;
};

static thread_local std::string g_err;

template <class F>
static int guarded(F f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  } catch (...) {
    g_err = "unknown error";
    return 2;
  }
}

static inline lbc_stream_t S(void* s) { return (lbc_stream_t)s; }

static void require_device() {
#ifndef LBC_HOST_EMU
  static std::once_flag once;
  static std::string err;
  std::call_once(once, [] {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
      err = std::string("liblbc_b200 needs a CUDA device (sm_100a); cudaGetDeviceCount: ") +
            (e == cudaSuccess ? "0 devices" : cudaGetErrorString(e));
      return;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, dev);
    if (p.major != 10) err = "liblbc_b200 is built for sm_100a only; found sm_" + std::to_string(p.major * 10 + p.minor);
  });
  if (!err.empty()) throw Error(err);
#endif
}

// ------------------------------------------------------------------ single-op entry points
namespace {
struct Tmp {
  std::vector<void*> v;
  template <class U>
  U* get(int64_t n) {
    void* p = dev_alloc(sizeof(U) * (size_t)n);
    v.push_back(p);
    return (U*)p;
  }
  ~Tmp() {
    for (void* p : v) dev_free(p);
  }
};
void sync_stream(lbc_stream_t s) {
#ifndef LBC_HOST_EMU
  LBC_CUDA(cudaStreamSynchronize(s));
#else
  (void)s;
#endif
}
ConvL make_conv(int H, int W, int Ci, int Co, int K, int stride, int pad) {
  ConvL c;
  c.Ci = Ci;
  c.Co = Co;
  c.K = K;
  c.stride = stride;
  c.pad = pad;
  c.H = H;
  c.W = W;
  c.OH = (H + 2 * pad - K) / stride + 1;
  c.OW = (W + 2 * pad - K) / stride + 1;
  return c;
}
}  // namespace

extern "C" {

const char* lbc_last_error(void) { return g_err.c_str(); }

int lbc_device_kind(void) {
#ifdef LBC_HOST_EMU
  return 0;
#else
  return 1;
#endif
}
const char* lbc_build_info(void) {
#ifdef LBC_HOST_EMU
  return "lbc host-emulation build (tests only)";
#else
  return "lbc_b200 CUDA build, sm_100a, " __DATE__ " " __TIME__;
#endif
}
int lbc_set_fast_kernels(int enabled) {
  fast::set_enabled((enabled & 1) != 0);
  fast::set_c64_variant((enabled & 2) == 0);   // bit 1 set: keep the generic tap-per-box kernel for 64->64 3x3 convs
  // variant switches (absent bits keep the LBC_PAIR default): 4 / 8 = CTA-pair (cta_group::2) conv GEMMs on / off,
  // 16 / 32 = row-of-taps weight-gradient kernel on / off
  int m = fast::pair_mode();
  if (enabled & 4) m |= 1;
  if (enabled & 8) m &= ~1;
  if (enabled & 16) m |= 2;
  if (enabled & 32) m &= ~2;
  if (enabled & 64) m |= 4;     // 64 / 128 = CTA-pair variant of the row-of-taps weight gradient on / off
  if (enabled & 128) m &= ~4;
  fast::set_pair_mode(m);
  if (enabled & 256) fast::set_experimental(fast::experimental() | 1);    // 256 / 512 = pair-walking weight pack on / off
  if (enabled & 512) fast::set_experimental(fast::experimental() & ~1);
  if (enabled & 1024) fast::set_experimental(fast::experimental() | 2);   // 1024 / 2048 = register-blocked head kernels on / off
  if (enabled & 2048) fast::set_experimental(fast::experimental() & ~2);
  if (enabled & 4096) fast::set_experimental(fast::experimental() | 4);   // 4096 / 8192 = capped par_for grids on / off
  if (enabled & 8192) fast::set_experimental(fast::experimental() & ~4);
  if (enabled & 16384) fast::set_experimental(fast::experimental() | 8);   // 16384 / 32768 = one-launch BatchNorm backward on / off
  if (enabled & 32768) fast::set_experimental(fast::experimental() & ~8);
  return 0;
}

long long lbc_kernel_launch_count(void) { return g_launches; }
int lbc_prof_enable(int on) {
  g_prof_on = on != 0;
  return 0;
}
int lbc_prof_reset(void) {
#ifndef LBC_HOST_EMU
  cudaDeviceSynchronize();
  for (auto& e : g_prof) {
    cudaEventDestroy(e.e0);
    cudaEventDestroy(e.e1);
  }
#endif
  g_prof.clear();
  return 0;
}
int lbc_prof_get(const char* category, double* total_ms, long long* count, double* flops, double* bytes) {
  return guarded([&] {
    double ms = 0, fl = 0, by = 0;
    long long n = 0;
#ifndef LBC_HOST_EMU
    LBC_CUDA(cudaDeviceSynchronize());
#endif
    for (auto& e : g_prof) {
      if (e.cat != category) continue;
#ifndef LBC_HOST_EMU
      float t = 0.f;
      LBC_CUDA(cudaEventElapsedTime(&t, e.e0, e.e1));
      ms += t;
#endif
      fl += e.flops;
      by += e.bytes;
      ++n;
    }
    *total_ms = ms;
    *count = n;
    *flops = fl;
    *bytes = by;
  });
}

int lbc_net_create(int kind, int precision, int max_batch, lbc_net_t** out) {
  return guarded([&] {
    require_device();
    LBC_CHECK(out, "null out pointer");
    lbc_net* n = new lbc_net;
    try {
      n->impl = make_net((NetKind)kind, (Precision)precision, max_batch);
    } catch (...) {
      delete n;
      throw;
    }
    *out = n;
  });
}
void lbc_net_destroy(lbc_net_t* net) { delete net; }
int lbc_net_num_params(const lbc_net_t* net) { return (int)net->impl->params.size(); }
int lbc_net_param_info(const lbc_net_t* net, int i, const char** name, int* ndim, int shape[4], int64_t* numel,
                       int64_t* offset, int* on_path) {
  return guarded([&] {
    LBC_CHECK(i >= 0 && i < (int)net->impl->params.size(), "param index out of range");
    const ParamInfo& p = net->impl->params[i];
    *name = p.name.c_str();
    *ndim = p.ndim;
    for (int d = 0; d < 4; ++d) shape[d] = p.shape[d];
    *numel = p.numel;
    *offset = p.offset;
    *on_path = p.on_path ? 1 : 0;
  });
}
int lbc_net_num_buffers(const lbc_net_t* net) { return (int)net->impl->buffers.size(); }
int lbc_net_buffer_info(const lbc_net_t* net, int i, const char** name, int64_t* numel, int64_t* offset) {
  return guarded([&] {
    LBC_CHECK(i >= 0 && i < (int)net->impl->buffers.size(), "buffer index out of range");
    const BufferInfo& b = net->impl->buffers[i];
    *name = b.name.c_str();
    *numel = b.numel;
    *offset = b.offset;
  });
}
int64_t lbc_net_total_params(const lbc_net_t* net) { return net->impl->n_params; }
int64_t lbc_net_total_buffers(const lbc_net_t* net) { return net->impl->n_buffers; }
int64_t lbc_net_workspace_bytes(const lbc_net_t* net) { return (int64_t)net->impl->workspace_bytes(); }
int lbc_net_bind(lbc_net_t* net, float* params, float* grads, float* buffers) {
  return guarded([&] {
    LBC_CHECK(params && buffers, "lbc_net_bind: null params/buffers");
    net->impl->bind(params, grads, buffers);
  });
}
int lbc_net_forward(lbc_net_t* net, const float* image, const float* speed, const float* command_onehot, int B,
                    int train, float* out_pred, float* out_preds, void* stream) {
  return guarded([&] {
    LBC_CHECK(image && speed && command_onehot, "lbc_net_forward: null input");
    net->impl->forward(image, speed, command_onehot, B, train != 0, out_pred, out_preds, S(stream));
  });
}
int lbc_net_forward_u8(lbc_net_t* net, const uint8_t* image_u8, int layout, const float* speed,
                       const float* command_onehot, int B, int train, float* out_pred, float* out_preds, void* stream) {
  return guarded([&] {
    LBC_CHECK(image_u8 && speed && command_onehot, "lbc_net_forward_u8: null input");
    net->impl->forward_u8(image_u8, layout, speed, command_onehot, B, train != 0, out_pred, out_preds, S(stream));
  });
}
int lbc_net_backward(lbc_net_t* net, const float* d_pred, const float* d_preds, void* stream) {
  return guarded([&] { net->impl->backward(d_pred, d_preds, S(stream)); });
}
int64_t lbc_net_read_tap(lbc_net_t* net, const char* name, float* out, int64_t capacity, void* stream) {
  int64_t n = -1;
  int rc = guarded([&] { n = net->impl->read_tap(name, out, capacity, S(stream)); });
  return rc == 0 ? n : -1;
}

int lbc_phase0_target(const float* teacher_pred, float* target_px, int64_t count, float w, float h, float fov_deg,
                      float world_y, float fixed_offset, void* stream) {
  return guarded([&] {
    require_device();
    ref::phase0_target(S(stream), teacher_pred, target_px, count, w, h, fov_deg, world_y, fixed_offset);
  });
}
int lbc_l1_loss(const float* a, const float* b, int N, int D, float sa, float ta, float sbx, float sby, float tb,
                const float* gout, float* loss_b, float* da, void* stream) {
  return guarded([&] {
    require_device();
    ref::l1_loss(S(stream), a, b, N, D, sa, ta, sbx, sby, tb, gout, loss_b, da);
  });
}
int lbc_phase1_convert_fwd(const float* p, float* out, int64_t count, float w, float h, float fov_deg,
                           float world_y, float fixed_offset, void* stream) {
  return guarded([&] {
    require_device();
    ref::phase1_convert_fwd(S(stream), p, out, count, w, h, fov_deg, world_y, fixed_offset);
  });
}
int lbc_phase1_convert_bwd(const float* p, const float* dout, float* dp, int64_t count, float w, float h,
                           float fov_deg, float world_y, float fixed_offset, void* stream) {
  return guarded([&] {
    require_device();
    ref::phase1_convert_bwd(S(stream), p, dout, dp, count, w, h, fov_deg, world_y, fixed_offset);
  });
}
int lbc_phase2_weight(const float* learner_map, const float* teacher_map, float* weight, int N, void* stream) {
  return guarded([&] {
    require_device();
    ref::phase2_weight(S(stream), learner_map, teacher_map, weight, N);
  });
}
int lbc_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                  float beta1, float beta2, float eps, int step, float grad_scale, void* stream) {
  return guarded([&] {
    require_device();
    LBC_CHECK(step >= 1, "adam step counter starts at 1");
    ref::adam(S(stream), params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step, grad_scale);
  });
}

// ------------------------------------------------------------------ single-op entry points

int lbc_op_conv_fwd(const float* x, const float* w_ref, float* y, int N, int H, int W, int Ci, int Co, int K,
                    int stride, int pad, int precision, void* stream) {
  return guarded([&] {
    require_device();
    lbc_stream_t s = S(stream);
    ConvL c = make_conv(H, W, Ci, Co, K, stride, pad);
    Tmp t;
    int64_t nx = (int64_t)N * H * W * Ci, ny = (int64_t)N * c.OH * c.OW * Co, nw = (int64_t)Co * K * K * Ci;
    if (precision == PREC_F32) {
      float* wp = t.get<float>(nw);
      ref::pack_weight<float>(s, w_ref, wp, Co, Ci, K);
      ref::conv_fwd<float>(s, x, wp, nullptr, false, y, N, H, W, Ci, Co, K, stride, pad, c.OH, c.OW);
    } else {
      bf16 *xb = t.get<bf16>(nx), *wb = t.get<bf16>(nw), *yb = t.get<bf16>(ny);
      ref::cast<float, bf16>(s, x, xb, nx);
      ref::pack_weight<bf16>(s, w_ref, wb, Co, Ci, K);
      c.wp = wb;
      if (!fast::conv_fwd<bf16>(c, xb, yb, N, s))
        ref::conv_fwd<bf16>(s, xb, wb, nullptr, false, yb, N, H, W, Ci, Co, K, stride, pad, c.OH, c.OW);
      ref::cast<bf16, float>(s, yb, y, ny);
    }
    sync_stream(s);
  });
}
int lbc_op_conv_dgrad(const float* dy, const float* w_ref, float* dx, int N, int H, int W, int Ci, int Co, int K,
                      int stride, int pad, int precision, void* stream) {
  return guarded([&] {
    require_device();
    lbc_stream_t s = S(stream);
    ConvL c = make_conv(H, W, Ci, Co, K, stride, pad);
    Tmp t;
    int64_t nx = (int64_t)N * H * W * Ci, ny = (int64_t)N * c.OH * c.OW * Co, nw = (int64_t)Co * K * K * Ci;
    if (precision == PREC_F32) {
      float* wp = t.get<float>(nw);
      ref::pack_weight<float>(s, w_ref, wp, Co, Ci, K);
      ref::conv_dgrad<float>(s, dy, wp, dx, N, H, W, Ci, Co, K, stride, pad, c.OH, c.OW, nullptr, false, false);
    } else {
      bf16 *xb = t.get<bf16>(nx), *wb = t.get<bf16>(nw), *yb = t.get<bf16>(ny);
      bf16* wtb = t.get<bf16>(nw);
      ref::cast<float, bf16>(s, dy, yb, ny);
      ref::pack_weight<bf16>(s, w_ref, wb, Co, Ci, K);
      ref::pack_weight_t<bf16>(s, w_ref, wtb, Co, Ci, K);
      c.wp = wb;
      c.wpt = wtb;
      if (!fast::conv_dgrad<bf16>(c, yb, xb, N, nullptr, false, s))
        ref::conv_dgrad<bf16>(s, yb, wb, xb, N, H, W, Ci, Co, K, stride, pad, c.OH, c.OW, nullptr, false, false);
      ref::cast<bf16, float>(s, xb, dx, nx);
    }
    sync_stream(s);
  });
}
int lbc_op_conv_wgrad(const float* x, const float* dy, float* dw_ref, int N, int H, int W, int Ci, int Co, int K,
                      int stride, int pad, int precision, void* stream) {
  return guarded([&] {
    require_device();
    lbc_stream_t s = S(stream);
    ConvL c = make_conv(H, W, Ci, Co, K, stride, pad);
    Tmp t;
    int64_t nx = (int64_t)N * H * W * Ci, ny = (int64_t)N * c.OH * c.OW * Co;
    int64_t wsn = 4 << 20;
    float* ws = t.get<float>(wsn);
    if (precision == PREC_F32) {
      ref::conv_wgrad<float>(s, x, dy, dw_ref, N, H, W, Ci, Co, K, stride, pad, c.OH, c.OW, ws, wsn);
    } else {
      bf16 *xb = t.get<bf16>(nx), *yb = t.get<bf16>(ny);
      ref::cast<float, bf16>(s, x, xb, nx);
      ref::cast<float, bf16>(s, dy, yb, ny);
      if (!fast::conv_wgrad<bf16>(c, xb, yb, dw_ref, N, ws, wsn, s))
        ref::conv_wgrad<bf16>(s, xb, yb, dw_ref, N, H, W, Ci, Co, K, stride, pad, c.OH, c.OW, ws, wsn);
    }
    sync_stream(s);
  });
}
int lbc_op_bn_train(const float* x, const float* gamma, const float* beta, const float* residual, int relu,
                    float* y, float* mean, float* var, int64_t M, int C, void* stream) {
  return guarded([&] {
    require_device();
    lbc_stream_t s = S(stream);
    Tmp t;
    double* ws = t.get<double>(1 << 20);
    float* rstd = t.get<float>(C);
    ref::bn_stats<float>(s, x, M, C, mean, var, ws);
    ref::bn_finalize(s, mean, var, C, M, 1e-5f, 0.1f, rstd, nullptr, nullptr);
    ref::bn_apply<float>(s, x, mean, rstd, gamma, beta, residual, relu != 0, y, M, C);
    sync_stream(s);
  });
}
int lbc_op_bn_bwd(const float* dy, const float* x, const float* gamma, float* dgamma, float* dbeta, float* dx,
                  int64_t M, int C, void* stream) {
  return guarded([&] {
    require_device();
    lbc_stream_t s = S(stream);
    Tmp t;
    double* ws = t.get<double>(1 << 20);
    float *mean = t.get<float>(C), *var = t.get<float>(C), *rstd = t.get<float>(C);
    ref::bn_stats<float>(s, x, M, C, mean, var, ws);
    ref::bn_finalize(s, mean, var, C, M, 1e-5f, 0.1f, rstd, nullptr, nullptr);
    ref::bn_bwd<float>(s, dy, x, mean, rstd, gamma, dgamma, dbeta, dx, M, C, ws);
    sync_stream(s);
  });
}
int lbc_op_maxpool(const float* x, float* y, const float* dy, float* dx, int N, int H, int W, int C, void* stream) {
  return guarded([&] {
    require_device();
    lbc_stream_t s = S(stream);
    int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
    Tmp t;
    uint8_t* idx = t.get<uint8_t>((int64_t)N * OH * OW * C);
    ref::maxpool_fwd<float>(s, x, y, idx, N, H, W, C, OH, OW);
    if (dy && dx) ref::maxpool_bwd<float>(s, dy, idx, dx, N, H, W, C, OH, OW);
    sync_stream(s);
  });
}
int lbc_op_spatial_softmax(const float* logits, float* out_xy, int rows, int H, int W, void* stream) {
  return guarded([&] {
    require_device();
    lbc_stream_t s = S(stream);
    LBC_CHECK(rows % 20 == 0, "rows must be a multiple of 20 (4 heads x 5 steps)");
    Tmp t;
    float *rmax = t.get<float>(rows), *rsum = t.get<float>(rows);
    ref::head_softmax(s, logits, rmax, rsum, out_xy, rows / 20, H, W);
    sync_stream(s);
  });
}

}  // extern "C"
