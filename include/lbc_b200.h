/* lbc_b200.h -- C ABI of the B200-native LearningByCheating image-agent training hot path.
 *
 * The reference (dotchen/LearningByCheating) has no FFI / plugin registry: its boundary for this
 * path is the Python nn.Module protocol.  This header is the native surface the
 * drop-in Python classes in learningbycheating_b200/ bind with ctypes; each entry point cites the
 * reference code it replaces (paths relative to the reference root).
 *
 * Conventions: every pointer is a DEVICE pointer (cudaMalloc'd, fp32 unless stated), tensors are
 * contiguous in the reference's own layouts (images NCHW, parameters as in state_dict()), `stream`
 * is a cudaStream_t passed as void* (NULL = legacy default stream).  All functions return 0 on
 * success and a non-zero code on failure; lbc_last_error() then returns a message.  Nothing here
 * falls back to the CPU: without a CUDA device the calls fail.
 */
#ifndef LBC_B200_H
#define LBC_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct lbc_net lbc_net_t;

#define LBC_NET_IMAGE_RESNET34 0    /* ImagePolicyModelSS('resnet34'),  bird_view/models/image.py:22-89   */
#define LBC_NET_BIRDVIEW_RESNET18 1 /* BirdViewPolicyModelSS('resnet18'), bird_view/models/birdview.py:47-79 */
#define LBC_PREC_F32 0  /* fp32 storage + fp32 math: parity mode (<=1e-3 vs the reference CPU path) */
#define LBC_PREC_BF16 1 /* bf16 NHWC activations / tcgen05 bf16 MMA, fp32 accumulation + statistics  */
/* fp32 NHWC storage, fp32 BN / softmax / loss / Adam; every convolution on the tcgen05 tensor cores with split-precision
 * operands (each fp32 operand = hi + lo 16-bit planes, 3 MMAs per K block, fp32 TMEM accumulation, fp32 epilogue):
 * parity-grade numerics (<=1e-3 vs the reference CPU path) at tensor-core speed */
#define LBC_PREC_F32TC 2

const char* lbc_last_error(void);
/* 1 = CUDA sm_100a build (the product), 0 = host-emulation build used only by the CPU unit tests */
int lbc_device_kind(void);
const char* lbc_build_info(void);
/* enable / disable the tcgen05 + fused kernels (tests compare them with the correctness-first kernels).
 * bit 0: fast kernels on; bit 1: generic tap-per-box kernel for the 64->64 3x3 convolutions;
 * variant switches (absent bits keep the LBC_PAIR default): 4 / 8 = CTA-pair (cta_group::2) conv GEMMs on / off,
 * 16 / 32 = row-of-taps weight gradient on / off, 64 / 128 = its CTA-pair variant on / off,
 * 256 / 512 = space-to-depth layout of the RGB stem operand on / off,
 * 1024 / 2048 = shared-row CTA-pair kernel (one activation box per kernel row) for the 3x3/s1 convolutions of the
 * 128-channel layers on / off, 4096 / 8192 = the same for the layers whose channel count is a multiple of 256,
 * 16384 / 32768 = weight gradient of the 64 -> 64 3x3 convolutions with all nine taps per CTA on / off,
 * 65536 / 131072 = the 64 -> 64 3x3/s1 convolutions on the shared-row CTA-pair kernel (64-wide N tiles) on / off. */
int lbc_set_fast_kernels(int enabled);
/* launch schedule of the bf16 throughput mode (a negative argument keeps the current value; the environment variables
 * LBC_WGRAD_OVERLAP / LBC_PDL set the initial ones).  wgrad_overlap: 0 = every kernel of backward on the caller's stream,
 * 1 = weight gradients on an engine-owned low-priority stream, 2 = and the dependency chain of backward on an engine-owned
 * high-priority stream (joined with the caller's stream by events on entry and exit).  pdl: 1 = every kernel is launched
 * with programmatic stream serialization (its prologue overlaps the tail of its predecessor).  Results do not depend on
 * either switch. */
int lbc_set_schedule(int wgrad_overlap, int pdl);
/* layout code of the padded stem operand lbc_op_stem's x4_out shows (bf16 path): 4 = [N][H+6][W+8][4], 8 = [N][H+6][W+8][8],
 * 16 = the 4-channel image with every 2x2 pixel block contiguous, [N][(H+6)/2][(W+8)/2][2][2][4]; 0 = column tensor */
int lbc_stem_layout(int C, int W, int normalize);

/* ---- instrumentation used by bench.py ---- */
/* number of kernels this library has launched so far (all streams) */
long long lbc_kernel_launch_count(void);
/* per-category CUDA-event timing of the engine's ops ("conv_fwd", "conv_dgrad", "conv_wgrad", "bn_fwd",
 * "bn_bwd", "pool", "elementwise", "head", "pack"): enable, run steps, then read totals.  flops / bytes are the
 * ALGORITHMIC figures of the timed launches. */
int lbc_prof_enable(int on);
int lbc_prof_reset(void);
int lbc_prof_get(const char* category, double* total_ms, long long* count, double* flops, double* bytes);

/* ---- network object: nn.Module construction (image.py:23-62, common.py:69-83, resnet.py:95-146) ---- */
int lbc_net_create(int kind, int precision, int max_batch, lbc_net_t** out);
void lbc_net_destroy(lbc_net_t* net);
/* parameter table in named_parameters() order; offsets index the flat fp32 parameter/gradient arrays */
int lbc_net_num_params(const lbc_net_t* net);
int lbc_net_param_info(const lbc_net_t* net, int i, const char** name, int* ndim, int shape[4], int64_t* numel,
                       int64_t* offset, int* on_path);
/* BatchNorm running_mean / running_var buffers; offsets index the flat fp32 buffer array */
int lbc_net_num_buffers(const lbc_net_t* net);
int lbc_net_buffer_info(const lbc_net_t* net, int i, const char** name, int64_t* numel, int64_t* offset);
int64_t lbc_net_total_params(const lbc_net_t* net);
int64_t lbc_net_total_buffers(const lbc_net_t* net);
int64_t lbc_net_workspace_bytes(const lbc_net_t* net);
/* bind caller-owned flat arrays (the nn.Parameters / .grad / BN buffers are views of them) */
int lbc_net_bind(lbc_net_t* net, float* params, float* grads, float* buffers);

/* nn.Module.__call__ -> ImagePolicyModelSS.forward (image.py:64-89) / BirdViewPolicyModelSS.forward
 * (birdview.py:61-79).  image: [B,C,H,W] fp32 NCHW (RGB in [0,1]; normalisation common.py:108 is applied
 * inside), speed [B], command_onehot [B,4] (train_utils.py:33-40).  train!=0: batch-statistics BN with
 * running-buffer update (net.train()); train==0: running statistics (net.eval()).
 * out_pred [B,5,2], out_preds [B,4,5,2] (either may be NULL). */
int lbc_net_forward(lbc_net_t* net, const float* image, const float* speed, const float* command_onehot, int B,
                    int train, float* out_pred, float* out_preds, void* stream);
/* Same, taking the camera frame (or bird's-eye masks) as uint8 exactly as stored by the data collector
 * (data_collector.py:234-252: rgb u8[160,384,3]; bird_view/utils/datasets/image_lmdb.py:133-135,190 applies
 * torchvision ToTensor = /255 on the host): layout 0 = [B,C,H,W] u8, 1 = [B,H,W,C] u8.  The /255 conversion is done
 * on the device, so the host->device copy is 4x smaller.  Results are bit-identical to
 * lbc_net_forward on float(image)/255. */
int lbc_net_forward_u8(lbc_net_t* net, const uint8_t* image_u8, int layout, const float* speed,
                       const float* command_onehot, int B, int train, float* out_pred, float* out_preds, void* stream);
/* loss.backward() through the network (train_image_phase0.py:184): upstream gradients wrt out_pred [B,5,2]
 * and/or out_preds [B,4,5,2] (NULL = none); writes every on-path parameter gradient into the bound
 * gradient array (overwrites, like backward() after zero_grad()). */
/* Low-latency eval forward (replaces the per-frame model call of ImageAgent.run_step, bird_view/models/image.py:124-196):
 * same results as lbc_net_forward(..., train = 0), replayed as one CUDA graph per (B, input kind) on an engine-owned stream
 * that is joined to `stream` with events.  Exactly one of image (fp32 [B,C,H,W]) / image_u8 (layout 0 [B,C,H,W], 1 [B,H,W,C])
 * is non-null.  weights_changed != 0: the parameters were modified since the previous lbc_net_infer call (the weight
 * operands are re-packed before the replay; training-mode calls in between re-pack on their own). */
int lbc_net_infer(lbc_net_t* net, const float* image, const uint8_t* image_u8, int layout, const float* speed,
                  const float* command_onehot, int B, int weights_changed, float* out_pred, float* out_preds, void* stream);
/* number of CUDA-graph replays lbc_net_infer has issued (0 on the host-emulation build and while a path is not capturable) */
int lbc_net_infer_replays(const lbc_net_t* net);
int lbc_net_backward(lbc_net_t* net, const float* d_pred, const float* d_preds, void* stream);
/* Data-parallel overlap (new: the reference is single-GPU; SURVEY.md 8(e) "one all-reduce, bucketed in reverse-execution
 * order and overlapped with backward").  The on-path gradients form lbc_net_num_grad_buckets() contiguous ranges of the flat
 * gradient array, numbered in the order backward completes them (0 = heads + decoder ... last = layer1 + stem; conv.fc.* is in
 * no bucket: it is never trained).  After lbc_net_enable_grad_events(net, 1), lbc_net_backward records one CUDA event per
 * bucket on its stream, and lbc_net_stream_wait_grads makes another stream wait for bucket k of the latest backward -- the
 * caller then launches that bucket's ncclAllReduce there while the rest of backward still runs. */
int lbc_net_num_grad_buckets(const lbc_net_t* net);
int lbc_net_grad_bucket(const lbc_net_t* net, int bucket, int64_t* offset, int64_t* numel);
int lbc_net_enable_grad_events(lbc_net_t* net, int on);
int lbc_net_stream_wait_grads(lbc_net_t* net, int bucket, void* stream);
/* copy an internal activation out as fp32 NCHW ("stem.raw", "stem.pool", "conv.layer1.0", ...,
 * "deconv.1|4|7", "logits"); returns element count or -1 */
int64_t lbc_net_read_tap(lbc_net_t* net, const char* name, float* out, int64_t capacity, void* stream);

/* ---- losses ---- */
/* CoordConverter.__call__ + _project_image_xy, training/train_image_phase0.py:54-79 (device-side, no
 * host round trip): teacher map coords [count,2] in [-1,1] -> image pixels, clipped to [0,w]x[0,h] */
int lbc_phase0_target(const float* teacher_pred, float* target_px, int64_t count, float w, float h, float fov_deg,
                      float world_y, float fixed_offset, void* stream);
/* per-sample L1: loss_b[n] = mean_d |a*sa+ta - (b*sb+tb)| with sb = sbx for even d, sby for odd d;
 * da[n,d] = gout[n]*sign(.)*sa/D (gout NULL -> 1/N, i.e. d(mean loss)).  Covers LocationLoss.forward of
 * train_image_phase0.py:86-89, train_image_phase1.py:66-70 and train_birdview.py:33-54 ('l1'). */
int lbc_l1_loss(const float* a, const float* b, int N, int D, float sa, float ta, float sbx, float sby, float tb,
                const float* gout, float* loss_b, float* da, void* stream);
/* CoordConverter.__call__, training/train_image_phase1.py:43-64, and its backward */
int lbc_phase1_convert_fwd(const float* p, float* out, int64_t count, float w, float h, float fov_deg,
                           float world_y, float fixed_offset, void* stream);
int lbc_phase1_convert_bwd(const float* p, const float* dout, float* dp, int64_t count, float w, float h,
                           float fov_deg, float world_y, float fixed_offset, void* stream);

/* get_weight, training/phase2_utils.py:50-59: per-sample replay-buffer weight of the DAgger stage;
 * learner_map / teacher_map [N,5,2] in map coords [-1,1] */
int lbc_phase2_weight(const float* learner_map, const float* teacher_map, float* weight, int N, void* stream);

/* ---- optimizer: torch.optim.Adam(lr).step() (train_image_phase0.py:231,185), one flat range ---- */
int lbc_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                  float beta1, float beta2, float eps, int step, float grad_scale, void* stream);

/* ---- launch trace (tests): which kernel family produced a result ---- */
/* on != 0: clear and start recording the kernel-family name of every launch; 0: stop */
int lbc_trace_enable(int on);
/* "name<TAB>count<LF>" per family (tcgen05 kernels by their own names, e.g. "conv_gemm_kernel<256,pair>"; the
 * correctness-first kernels as "... [with Tag = lbc::ref::k_conv_fwd]"); returns the full length, writes at most cap-1 chars */
int lbc_trace_dump(char* buf, int cap);

/* ---- single-op entry points (fp32 NHWC in/out; the parity tests of each kernel) ----
 * precision LBC_PREC_F32: correctness-first fp32 kernels; LBC_PREC_BF16: operands rounded to bf16 storage and run
 * through the kernels of the bf16 training step (tcgen05 GEMMs, fused BN / pool / head kernels).  Each op mirrors one
 * reference op: nn.Conv2d (resnet.py:15-22) and its two gradients, nn.ConvTranspose2d+bias+ReLU (image.py:39-46),
 * nn.BatchNorm2d train mode + residual + ReLU (resnet.py:41-53), nn.MaxPool2d (resnet.py:105), SpatialSoftmax
 * (common.py:136-152), the four heads (image.py:54-60), the stem (resnet.py:102,148 + common.py:101-109). */
int lbc_op_conv_fwd(const float* x, const float* w_ref, float* y, int N, int H, int W, int Ci, int Co, int K,
                    int stride, int pad, int precision, const float* bias_co, float* stats_out, void* stream);
int lbc_op_conv_dgrad(const float* dy, const float* w_ref, float* dx, int N, int H, int W, int Ci, int Co, int K,
                      int stride, int pad, int precision, const float* bias_ci, int relu, void* stream);
int lbc_op_block_dgrad_ds(const float* dy1, const float* dy_ds, const float* w1_ref, const float* wd_ref, float* dx, int N,
                          int H, int W, int Ci, int Co, int precision, void* stream);
int lbc_op_conv_wgrad(const float* x, const float* dy, float* dw_ref, int N, int H, int W, int Ci, int Co, int K,
                      int stride, int pad, int precision, void* stream);
int lbc_op_bn_train(const float* x, const float* gamma, const float* beta, const float* residual, int relu,
                    float* y, float* mean, float* var, int64_t M, int C, int precision, float* running_mean,
                    float* running_var, float* negshift, uint8_t* mask_bits_out, void* stream);
/* use_mask_bits (bf16 path): hand the mask to the kernels as one bit per element ([M][C/8] bytes, the format
 * lbc_op_bn_train's mask_bits_out has) instead of the activation tensor */
int lbc_op_bn_bwd(const float* dy, const float* x, const float* gamma, float* dgamma, float* dbeta, float* dx,
                  int64_t M, int C, int precision, const float* mask_act, const float* beta, int own_relu, int use_mask_bits,
                  void* stream);
int lbc_op_ew(float* dst, const float* src, const float* act, int64_t n, int mode, int precision, int use_mask_bits,
              void* stream);
/* dst += src * (act > 0), then BatchNorm backward of dst * (act_prev > 0) wrt x (batch statistics of x): the residual
 * blocks' d(out) chain, resnet.py:41-53 backward (bf16 path: add + reduce pass in one kernel) */
int lbc_op_resid_bn_bwd(float* dst, const float* src, const float* act, const float* x, const float* act_prev,
                        const float* gamma, float* dgamma, float* dbeta, float* dx, int64_t M, int C, int precision,
                        void* stream);
/* dst [M][Cd] = src [M][0:min(Cs, Cd)], channels Cs.. filled with fill[m / rows_per_fill]: the late fusion of the speed
 * (image.py:77-79: Cd = 640, Cs = 512, fill = speed) and, with Cd < Cs and fill = NULL, the slice of its backward */
int lbc_op_copy_channels(const float* src, float* dst, int64_t M, int Cd, int Cs, const float* fill, int rows_per_fill,
                         int precision, void* stream);
int lbc_op_maxpool(const float* x, float* y, const float* dy, float* dx, int N, int H, int W, int C, void* stream);
int lbc_op_bn_relu_maxpool(const float* x, const float* mean, const float* rstd, const float* gamma, const float* beta,
                           float* y, const float* dy, float* dx, int N, int H, int W, int C, int precision, void* stream);
int lbc_op_stem_tail(const float* x, const float* gamma, const float* beta, float* y, const float* dy, float* dgamma,
                     float* dbeta, float* dx, int N, int H, int W, int C, int precision, void* stream);
int lbc_op_spatial_softmax(const float* logits, float* out_xy, int rows, int H, int W, int precision, void* stream);
int lbc_op_head(const float* h, const float* gamma, const float* beta, const float* w, const float* bias, int N, int H, int W,
                float* running_mean, float* running_var, float* logits_out, float* preds_out, const float* onehot,
                const float* d_pred, const float* d_preds, float* dgamma, float* dbeta, float* dw, float* dbias, float* dh,
                int* dh_masked, int precision, void* stream);
int lbc_op_stem(const float* img, const uint8_t* img_u8, int layout, const float* w_ref, int normalize, int N, int C, int H,
                int W, float* x4_out, float* y, float* stats_out, const float* dy, float* dw, int precision, void* stream);

#ifdef __cplusplus
}
#endif
#endif
