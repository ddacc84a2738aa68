// lbc_net.h -- native graph executor for ImagePolicyModelSS / BirdViewPolicyModelSS
// (bird_view/models/image.py:22-89, birdview.py:47-79): parameter table in the reference's
// state_dict order, activation plan in HBM (NHWC), forward (train / eval BN), backward.
//
// Parameters, gradients and BN running buffers are flat fp32 arrays OWNED BY THE CALLER (the
// Python nn.Module re-points its nn.Parameters at views of them, so torch.save / load_state_dict /
// torch.optim.Adam keep working).  Layouts inside the flat arrays are the reference's:
// Conv2d [Co][Ci][K][K], ConvTranspose2d [Ci][Co][K][K].
#pragma once
#include <memory>

#include "lbc_common.h"
#include "lbc_ref_ops.h"

namespace lbc {

enum NetKind { NET_IMAGE_RESNET34 = 0, NET_BIRDVIEW_RESNET18 = 1 };
// PREC_F32TC: fp32 NHWC storage and fp32 BN / softmax / loss as PREC_F32, but every convolution runs on the tcgen05 tensor
// cores with split-precision operands (hi + lo 16-bit planes, 3 MMAs per K block): parity-grade numerics at GEMM speed
enum Precision { PREC_F32 = 0, PREC_BF16 = 1, PREC_F32TC = 2 };

struct ParamInfo {
  std::string name;
  int ndim;
  int shape[4];
  int64_t numel;
  int64_t offset;  // into the flat parameter / gradient arrays
  bool on_path;    // false for conv.fc.* (never trained: grad is None in the reference)
};
struct BufferInfo {  // running_mean / running_var, in state_dict order
  std::string name;
  int64_t numel;
  int64_t offset;
};

struct ConvL {
  int Ci = 0, Co = 0, K = 0, stride = 1, pad = 0;  // conv roles (for a deconv: Co = deconv C_in)
  int H = 0, W = 0, OH = 0, OW = 0;                // conv-role input / output spatial size
  int64_t w_off = -1, b_off = -1;
  bool deconv = false;
  void* wp = nullptr;   // packed [Co][K][K][Ci]  (GEMM B operand of the forward conv)
  void* wpt = nullptr;  // packed [Ci][K][K][Co]  (GEMM B operand of the data gradient / deconv forward)
  void* wcomb = nullptr;  // block-entry 3x3/s2 conv only: [Ci][2*Co] = [centre tap of this conv | 1x1/s2 downsample]^T
  // PREC_F32TC: the same three packs split into fp16 [hi | lo] planes per tap slab (x kTcWeightScale), rebuilt every forward
  void* wp16 = nullptr;
  void* wpt16 = nullptr;
  void* wcomb16 = nullptr;
};
struct BNL {
  int C = 0;
  int64_t g_off = -1, b_off = -1, rm_off = -1, rv_off = -1;
  float *mean = nullptr, *var = nullptr, *rstd = nullptr;  // saved batch statistics
  // bf16 path: the conv feeding this BN stores (output + negshift[c]) so that the bf16 rounding acts on a centred
  // tensor; negshift = -(previous batch mean) (train) or -running_mean (eval).  See DESIGN.md "centring shift".
  float* negshift = nullptr;
};

class NetBase {
 public:
  virtual ~NetBase() { infer_release(); }
  NetKind kind;
  Precision prec;
  int max_batch = 0;
  int in_ch = 0, in_h = 0, in_w = 0, head_h = 0, head_w = 0;
  std::vector<ParamInfo> params;
  std::vector<BufferInfo> buffers;
  int64_t n_params = 0, n_buffers = 0;
  float *P = nullptr, *G = nullptr, *BUF = nullptr;  // bound flat arrays
  void bind(float* p, float* g, float* b) {
    if (p != P || b != BUF) infer_drop_graphs();   // captured launches carry the old parameter / buffer addresses
    P = p;
    G = g;
    BUF = b;
  }
  void infer_drop_graphs();
  virtual void forward(const float* image, const float* speed, const float* onehot, int B, bool train,
                       float* out_pred, float* out_preds, lbc_stream_t s) = 0;
  // uint8 frames ([B,C,H,W] or [B,H,W,C]): ToTensor's /255 on the device, then the float path
  virtual void forward_u8(const uint8_t* image, int layout, const float* speed, const float* onehot, int B, bool train,
                          float* out_pred, float* out_preds, lbc_stream_t s) = 0;
  virtual void backward(const float* d_pred, const float* d_preds, lbc_stream_t s) = 0;
  // ---- low-latency inference (ImageAgent.run_step is a B = 1 eval forward, image.py:124-196) ----
  // An eval forward is ~130 launches of a few microseconds each: at B = 1 the host-side launch cost IS the latency.  infer()
  // copies the inputs into engine-owned buffers, replays the forward as ONE CUDA graph (captured on first use per batch size /
  // input kind on an engine-owned stream; the caller's stream is joined with events) and copies the results out.  The
  // weight operands are packed OUTSIDE the graph, only when the caller says the parameters changed.
  virtual void repack(lbc_stream_t s) = 0;
  void infer(const float* image, const uint8_t* image_u8, int layout, const float* speed, const float* onehot, int B,
             bool weights_changed, float* out_pred, float* out_preds, lbc_stream_t s);
  bool skip_pack = false;       // set by infer() around forward(): the packs are current
  bool packs_current = false;   // cleared by every forward() that packs and by backward()
  bool infer_no_graph = false;  // a capture failed: infer() keeps running the eager call sequence
  struct InferGraph {
    int B = 0, kind = 0, layout = 0, variant = 0;
    void* exec = nullptr;
  };
  std::vector<InferGraph> infer_graphs;
  int infer_replays = 0;        // graph launches so far (tests assert that the graph path really ran)
  void *infer_stream = nullptr, *infer_ev_in = nullptr, *infer_ev_out = nullptr;
  float *inf_img = nullptr, *inf_speed = nullptr, *inf_onehot = nullptr, *inf_pred = nullptr, *inf_preds = nullptr;
  uint8_t* inf_img_u8 = nullptr;
  void infer_release();
  // debugging / parity taps: copy a named internal tensor out as fp32 NCHW
  virtual int64_t read_tap(const char* name, float* out, int64_t cap, lbc_stream_t s) = 0;
  virtual size_t workspace_bytes() const = 0;
  // gradient buckets in the order backward() completes them (0 = heads + decoder, then layer4, layer3, layer2, and
  // layer1 + stem last): each one contiguous [offset, offset + numel) range of the flat gradient array.  With events
  // enabled, backward() records one event per bucket on its stream right after the bucket's last gradient kernel, so a
  // second stream can start the bucket's all-reduce while the rest of backward still runs.
  struct GradBucket {
    int64_t offset = 0, numel = 0;
    void* event = nullptr;
  };
  std::vector<GradBucket> buckets;
  bool grad_events = false;
  void enable_grad_events(bool on);
  void stream_wait_bucket(int bucket, lbc_stream_t s);
};

std::unique_ptr<NetBase> make_net(NetKind kind, Precision prec, int max_batch);

}  // namespace lbc
