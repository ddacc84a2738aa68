// lbc_fast.cu -- sm_100a fast kernels behind the hooks of lbc_fast.h.
#include "lbc_fast.h"

#include <map>

namespace lbc {
long long g_launches = 0;
bool g_trace_on = false;
static std::map<std::string, long long>& trace_map() {
  static std::map<std::string, long long> m;
  return m;
}
void trace_note(const char* name) { ++trace_map()[name]; }
void trace_reset() { trace_map().clear(); }
std::string trace_dump() {
  std::string out;
  for (auto& kv : trace_map()) out += kv.first + "\t" + std::to_string(kv.second) + "\n";
  return out;
}
bool g_prof_on = false;
int g_wgrad_overlap = [] {
  const char* e = getenv("LBC_WGRAD_OVERLAP");   // lbc_net.cu backward(): 0 serial, 1 side stream (default: 13.78 -> 12.94 ms per step on one box), 2 + high-priority chain (13.12)
  return e ? atoi(e) : 1;
}();
#ifndef LBC_HOST_EMU
int g_pdl = [] {
  const char* e = getenv("LBC_PDL");   // programmatic dependent launch of every kernel (lbc_common.h); 13.78 -> 13.16 ms per step alone
  return e ? atoi(e) : 1;
}();
#endif
std::vector<ProfEntry> g_prof;
namespace fast {

static bool g_enabled = true;
bool enabled() { return g_enabled; }
void set_enabled(bool on) { g_enabled = on; }

#ifdef LBC_HOST_EMU
bool stem_im2col_bf16(const float*, bf16*, int, int, int, int, int, int, int, bool, lbc_stream_t) { return false; }
bool stem_pack_weight_bf16(const float*, bf16*, int, int, lbc_stream_t) { return false; }
bool stem_pad4_bf16(const float*, bf16*, int, int, int, int, bool, lbc_stream_t) { return false; }
bool stem_pad4_u8_bf16(const uint8_t*, int, bf16*, int, int, int, int, bool, lbc_stream_t) { return false; }
bool stem_pack_w224_bf16(const float*, bf16*, int, lbc_stream_t, int) { return false; }
bool stem_unpack_wgrad(const float*, float*, int, int, lbc_stream_t) { return false; }
#else
struct k_stem_im2col;
struct k_stem_pack;
struct k_stem_unpack;

// One block = SEG (64, or 32 when the row is not a multiple of 64: the teacher's 96 columns) consecutive output positions of
// one output row.  Gather: thread (pos, c*7+kh) walks the 7 kw taps of
// one input row (lanes = consecutive positions -> 8-byte lane stride, L1-resident reuse); the [64][Kp] bf16 tile is
// staged in shared memory and written back as one contiguous 64*Kp*2-byte run with 16-byte stores.
__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ img, uint4* __restrict__ col, int C, int H,
                                                          int W, int OH, int OW, int Kp, int normalize, int SEG) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ uint16_t tile[];  // [SEG][Kp + 2]: odd word stride -> lanes (positions) hit distinct banks
  const int KS = Kp + 2;
  const int segs = OW / SEG;
  int blk = blockIdx.x;
  const int seg = blk % segs;
  blk /= segs;
  const int oh = blk % OH;
  const int b = blk / OH;
  const int ow0 = seg * SEG;
  const int KT = 49 * C;
  // zero the K padding
  for (int i = threadIdx.x; i < SEG * (Kp - KT); i += 256) {
    int pos = i / (Kp - KT), k = KT + i % (Kp - KT);
    tile[pos * KS + k] = 0;
  }
  const int pos = threadIdx.x % SEG;
  const int ow = ow0 + pos;
  for (int ck = threadIdx.x / SEG; ck < C * 7; ck += 256 / SEG) {
    const int c = ck / 7, kh = ck - c * 7;
    const int ih = oh * 2 - 3 + kh;
    const bool rowok = ih >= 0 && ih < H;
    const float* rowp = img + (((int64_t)b * C + c) * H + (rowok ? ih : 0)) * W;
    float mean = 0.f;
    if (normalize) {
      mean = c == 0 ? 0.485f : (c == 1 ? 0.456f : 0.406f);
    }
#pragma unroll
    for (int kw = 0; kw < 7; ++kw) {
      const int iw = ow * 2 - 3 + kw;
      float v = 0.f;
      if (rowok && iw >= 0 && iw < W) {
        v = __ldg(rowp + iw);
        if (normalize) v = (v - mean) / (c == 0 ? 0.229f : (c == 1 ? 0.224f : 0.225f));
      }
      tile[pos * KS + (kh * 7 + kw) * C + c] = float_to_bf16(v).v;
    }
  }

  __syncthreads();
  const int64_t base = (((int64_t)b * OH + oh) * OW + ow0) * Kp / 2;  // in 32-bit words
  const uint32_t* t32 = reinterpret_cast<const uint32_t*>(tile);
  uint32_t* out32 = reinterpret_cast<uint32_t*>(col);
  const int wpr = Kp / 2;
  for (int i = threadIdx.x; i < SEG * wpr; i += 256) {
    const int r = i / wpr, w = i - r * wpr;
    out32[base + i] = t32[r * (KS / 2) + w];
  }
}

bool stem_im2col_bf16(const float* img, bf16* col, int B, int C, int H, int W, int OH, int OW, int Kp, bool normalize,
                      lbc_stream_t s) {
  if (!enabled()) return false;
  if (OW % 32 == 0 && Kp % 8 == 0) {
    static bool configured = false;
    const int SEG = OW % 64 == 0 ? 64 : 32;
    const int smem = SEG * (Kp + 2) * 2;
    if (!configured) {
      LBC_CUDA(cudaFuncSetAttribute(stem_im2col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 514 * 2));
      configured = true;
    }
    if (smem <= 64 * 514 * 2) {
      { auto k_ = stem_im2col_kernel; LBC_LAUNCH(k_, dim3(B * OH * (OW / SEG)), dim3(256), smem, s, img, (uint4*)col, C, H, W, OH, OW, Kp, normalize ? 1 : 0, SEG); }
      LBC_LAUNCHED("stem_im2col_kernel");
      LBC_CUDA(cudaGetLastError());
      return true;
    }
  }
  const int G = Kp / 8;
  const int KT = 49 * C;
  int64_t n = (int64_t)B * OH * OW * G;
  par_for<k_stem_im2col>(s, n, [=] __device__(int64_t i) {
    int g = (int)(i % G);
    int64_t pix = i / G;
    int ow = (int)(pix % OW);
    int64_t t = pix / OW;
    int oh = (int)(t % OH);
    int b = (int)(t / OH);
    uint32_t pk[4];
#pragma unroll
    for (int e2 = 0; e2 < 4; ++e2) {
      float v[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        int k = g * 8 + e2 * 2 + h;
        float val = 0.f;
        if (k < KT) {
          int tap = k / C, c = k - tap * C;
          int kh = tap / 7, kw = tap - kh * 7;
          int ih = oh * 2 - 3 + kh, iw = ow * 2 - 3 + kw;
          if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
            val = __ldg(img + (((int64_t)b * C + c) * H + ih) * W + iw);
            if (normalize) {
              float mean = c == 0 ? 0.485f : (c == 1 ? 0.456f : 0.406f);
              float sd = c == 0 ? 0.229f : (c == 1 ? 0.224f : 0.225f);
              val = (val - mean) / sd;
            }
          }
        }
        v[h] = val;
      }
      pk[e2] = (uint32_t)float_to_bf16(v[0]).v | ((uint32_t)float_to_bf16(v[1]).v << 16);
    }
    reinterpret_cast<uint4*>(col)[i] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  });
  return true;
}
// ---- x4 writer, four padded pixels per thread (W % 4 == 0: a 4-pixel group is entirely image or entirely border) ----
// KIND 0: fp32 [B,C,H,W]; 1: uint8 [B,C,H,W]; 2: uint8 [B,H,W,C].  Same per-element expressions as the one-pixel writers
// below (bit-identical output); 32-byte stores, 16-byte / 4-byte vector loads, 4x fewer threads (round 1: 222 us for the
// 132 MB tensor at B = 256, bound by block dispatch of 16.6 M one-pixel threads).
// Four padded pixels of ROWS padded rows per thread.  ROWS = 1: [B][H+6][W+8][4], 32 contiguous bytes per thread.
// ROWS = 2 (space-to-depth operand, lbc_fast.h: stem_ch = 16): rows 2Y, 2Y+1 -> the two 2x2 blocks (Y, 2*gcol), (Y, 2*gcol+1)
// = 64 contiguous bytes per thread ([row parity][column parity][4 ch] per block).  (Writing one row per thread in that layout
// stored 16-byte pieces 32 bytes apart -- half sectors, completed by another warp: 117 us instead of 97 us per launch.)
template <int KIND, int ROWS>
__global__ void __launch_bounds__(256) stem_pad4_kernel(const void* __restrict__ img, uint4* __restrict__ x4, int B, int C, int H,
                                                        int W, int normalize) {
  pdl_wait();
  pdl_trigger();
  const int HP = H + 6, WG = (W + 8) / 4;
  const int64_t n = (int64_t)B * (HP / ROWS) * WG;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int gcol = (int)(i % WG);
    const int64_t t = i / WG;
    const int yrow = (int)(t % (HP / ROWS));
    const int b = (int)(t / (HP / ROWS));
    const int iw = gcol * 4 - 4;
    uint32_t w[ROWS][8];
#pragma unroll
    for (int rr = 0; rr < ROWS; ++rr) {
      const int ih = yrow * ROWS + rr - 3;
      float v[4][4];
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int c = 0; c < 4; ++c) v[p][c] = 0.f;
      if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
        for (int c = 0; c < C; ++c) {
          float x[4];
          if (KIND == 0) {
            const float4 q = __ldg(reinterpret_cast<const float4*>((const float*)img + (((int64_t)b * C + c) * H + ih) * W + iw));
            x[0] = q.x; x[1] = q.y; x[2] = q.z; x[3] = q.w;
          } else if (KIND == 1) {
            const uint32_t q = __ldg(reinterpret_cast<const uint32_t*>((const uint8_t*)img + (((int64_t)b * C + c) * H + ih) * W + iw));
#pragma unroll
            for (int p = 0; p < 4; ++p) x[p] = (float)((q >> (8 * p)) & 0xffu) / 255.0f;
          } else {
            const uint8_t* src = (const uint8_t*)img + (((int64_t)b * H + ih) * W + iw) * C + c;
#pragma unroll
            for (int p = 0; p < 4; ++p) x[p] = (float)__ldg(src + p * C) / 255.0f;
          }
          const float mean = c == 0 ? 0.485f : (c == 1 ? 0.456f : 0.406f);
          const float sd = c == 0 ? 0.229f : (c == 1 ? 0.224f : 0.225f);
#pragma unroll
          for (int p = 0; p < 4; ++p) v[p][c] = normalize ? (x[p] - mean) / sd : x[p];
        }
      }
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        w[rr][2 * p] = (uint32_t)float_to_bf16(v[p][0]).v | ((uint32_t)float_to_bf16(v[p][1]).v << 16);
        w[rr][2 * p + 1] = (uint32_t)float_to_bf16(v[p][2]).v | ((uint32_t)float_to_bf16(v[p][3]).v << 16);
      }
    }
    if (ROWS == 1) {
      x4[i * 2] = make_uint4(w[0][0], w[0][1], w[0][2], w[0][3]);
      x4[i * 2 + 1] = make_uint4(w[0][4], w[0][5], w[0][6], w[0][7]);
    } else {
      // block (Y, X) = 32 bytes = [row 2Y: px 2X, 2X+1][row 2Y+1: px 2X, 2X+1]; this thread: X = 2*gcol, 2*gcol + 1
      x4[i * 4] = make_uint4(w[0][0], w[0][1], w[0][2], w[0][3]);
      x4[i * 4 + 1] = make_uint4(w[ROWS - 1][0], w[ROWS - 1][1], w[ROWS - 1][2], w[ROWS - 1][3]);
      x4[i * 4 + 2] = make_uint4(w[0][4], w[0][5], w[0][6], w[0][7]);
      x4[i * 4 + 3] = make_uint4(w[ROWS - 1][4], w[ROWS - 1][5], w[ROWS - 1][6], w[ROWS - 1][7]);
    }
  }
}
template <int KIND>
static bool launch_stem_pad4(const void* img, bf16* x4, int B, int C, int H, int W, bool normalize, lbc_stream_t s) {
  const bool s2d = stem_ch(C, W, normalize) == 16;
  const int64_t n = (int64_t)B * ((H + 6) / (s2d ? 2 : 1)) * ((W + 8) / 4);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t blocks = (n + 255) / 256, cap = (int64_t)sms * 16;
  if (blocks > cap) blocks = cap;
  if (s2d)
    { auto k_ = stem_pad4_kernel<KIND, 2>; LBC_LAUNCH(k_, dim3((unsigned)blocks), dim3(256), 0, s, img, (uint4*)x4, B, C, H, W, normalize ? 1 : 0); }
  else
    { auto k_ = stem_pad4_kernel<KIND, 1>; LBC_LAUNCH(k_, dim3((unsigned)blocks), dim3(256), 0, s, img, (uint4*)x4, B, C, H, W, normalize ? 1 : 0); }
  LBC_LAUNCHED(KIND == 0 ? "stem_pad4_kernel<f32>" : "stem_pad4_kernel<u8>");
  LBC_CUDA(cudaGetLastError());
  return true;
}
// 8-channel padded image (C_in 5..8: the teacher's 7-channel bird's-eye view): one thread per padded pixel, one 16-byte store
template <bool U8>
__global__ void __launch_bounds__(256) stem_pad8_kernel(const void* __restrict__ img, int layout, uint4* __restrict__ x8, int B,
                                                        int C, int H, int W) {
  pdl_wait();
  pdl_trigger();
  const int HP = H + 6, WP = W + 8;
  const int64_t n = (int64_t)B * HP * WP;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int col = (int)(i % WP);
    const int64_t t = i / WP;
    const int row = (int)(t % HP);
    const int b = (int)(t / HP);
    const int ih = row - 3, iw = col - 4;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
      for (int c = 0; c < C; ++c) {
        if (U8) {
          const int64_t j = layout == 1 ? (((int64_t)b * H + ih) * W + iw) * C + c : (((int64_t)b * C + c) * H + ih) * W + iw;
          v[c] = (float)__ldg((const uint8_t*)img + j) / 255.0f;
        } else {
          v[c] = __ldg((const float*)img + (((int64_t)b * C + c) * H + ih) * W + iw);
        }
      }
    }
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) w[q] = (uint32_t)float_to_bf16(v[2 * q]).v | ((uint32_t)float_to_bf16(v[2 * q + 1]).v << 16);
    x8[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}
template <bool U8>
static bool launch_stem_pad8(const void* img, int layout, bf16* x8, int B, int C, int H, int W, lbc_stream_t s) {
  const int64_t n = (int64_t)B * (H + 6) * (W + 8);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t blocks = (n + 255) / 256, cap = (int64_t)sms * 16;
  if (blocks > cap) blocks = cap;
  { auto k_ = stem_pad8_kernel<U8>; LBC_LAUNCH(k_, dim3((unsigned)blocks), dim3(256), 0, s, img, layout, (uint4*)x8, B, C, H, W); }
  LBC_LAUNCHED(U8 ? "stem_pad8_kernel<u8>" : "stem_pad8_kernel<f32>");
  LBC_CUDA(cudaGetLastError());
  return true;
}
struct k_stem_pad4;
struct k_stem_w224;
// x4[b][ih+3][iw+4][c] = normalised pixel (c < C), zero elsewhere (borders, 4th channel)
bool stem_pad4_bf16(const float* img, bf16* x4, int B, int C, int H, int W, bool normalize, lbc_stream_t s) {
  if (enabled() && C > 4 && C <= 8 && !normalize) return launch_stem_pad8<false>(img, 0, x4, B, C, H, W, s);
  if (!enabled() || C > 4) return false;
  if (W % 4 == 0 && (C == 3 || !normalize)) return launch_stem_pad4<0>(img, x4, B, C, H, W, normalize, s);
  const int HP = H + 6, WP = W + 8;
  int64_t n = (int64_t)B * HP * WP;
  par_for<k_stem_pad4>(s, n, [=] __device__(int64_t i) {
    int col = (int)(i % WP);
    int64_t t = i / WP;
    int row = (int)(t % HP);
    int b = (int)(t / HP);
    int ih = row - 3, iw = col - 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
      for (int c = 0; c < C; ++c) {
        float x = __ldg(img + (((int64_t)b * C + c) * H + ih) * W + iw);
        if (normalize) {
          float mean = c == 0 ? 0.485f : (c == 1 ? 0.456f : 0.406f);
          float sd = c == 0 ? 0.229f : (c == 1 ? 0.224f : 0.225f);
          x = (x - mean) / sd;
        }
        v[c] = x;
      }
    }
    uint2 o;
    o.x = (uint32_t)float_to_bf16(v[0]).v | ((uint32_t)float_to_bf16(v[1]).v << 16);
    o.y = (uint32_t)float_to_bf16(v[2]).v | ((uint32_t)float_to_bf16(v[3]).v << 16);
    reinterpret_cast<uint2*>(x4)[i] = o;
  });
  return true;
}
struct k_stem_pad4_u8;
// same tensor straight from uint8 frames (layout 0 = [B,C,H,W], 1 = [B,H,W,C], the data collector's on-disk order):
// ToTensor's x/255 (torchvision, image_lmdb.py:133-135) and the normalisation evaluated with the expressions of the
// two-step path (u8_to_f32_nchw + stem_pad4_bf16), so the result is bit-identical while the 189 MB fp32 image
// (one write + one read per step at B = 256) never exists.
bool stem_pad4_u8_bf16(const uint8_t* img, int layout, bf16* x4, int B, int C, int H, int W, bool normalize, lbc_stream_t s) {
  if (enabled() && C > 4 && C <= 8 && !normalize) return launch_stem_pad8<true>(img, layout, x4, B, C, H, W, s);
  if (!enabled() || C > 4) return false;
  if (W % 4 == 0 && (C == 3 || !normalize))
    return layout == 1 ? launch_stem_pad4<2>(img, x4, B, C, H, W, normalize, s) : launch_stem_pad4<1>(img, x4, B, C, H, W, normalize, s);
  const int HP = H + 6, WP = W + 8;
  int64_t n = (int64_t)B * HP * WP;
  par_for<k_stem_pad4_u8>(s, n, [=] __device__(int64_t i) {
    int col = (int)(i % WP);
    int64_t t = i / WP;
    int row = (int)(t % HP);
    int b = (int)(t / HP);
    int ih = row - 3, iw = col - 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
      for (int c = 0; c < C; ++c) {
        const int64_t j = layout == 1 ? (((int64_t)b * H + ih) * W + iw) * C + c : (((int64_t)b * C + c) * H + ih) * W + iw;
        float x = (float)__ldg(img + j) / 255.0f;
        if (normalize) {
          float mean = c == 0 ? 0.485f : (c == 1 ? 0.456f : 0.406f);
          float sd = c == 0 ? 0.229f : (c == 1 ? 0.224f : 0.225f);
          x = (x - mean) / sd;
        }
        v[c] = x;
      }
    }
    uint2 o;
    o.x = (uint32_t)float_to_bf16(v[0]).v | ((uint32_t)float_to_bf16(v[1]).v << 16);
    o.y = (uint32_t)float_to_bf16(v[2]).v | ((uint32_t)float_to_bf16(v[3]).v << 16);
    reinterpret_cast<uint2*>(x4)[i] = o;
  });
  return true;
}
// w224[co][kh][kw'][c], c < CH = stem_ch(C): kw' = kw + 1 in 0..7 (kw' = 0 and c >= C are zero)
struct k_stem_w256;
bool stem_pack_w224_bf16(const float* w_ref, bf16* w224, int C, lbc_stream_t s, int CH) {
  if (!CH) CH = stem_ch(C);
  if (!CH) return false;
  if (CH == 16) {   // space-to-depth: K index = [row pair 4][column pair 4][row parity][column parity][4 ch]
    par_for<k_stem_w256>(s, (int64_t)64 * 256, [=] __device__(int64_t i) {
      const int e = (int)(i % 64);
      const int khp = (int)((i / 64) % 4);
      const int co = (int)(i / 256);
      const int xs = e >> 4, a = (e >> 3) & 1, b2 = (e >> 2) & 1, c = e & 3;
      const int kh = 2 * khp + a, kwp = 2 * xs + b2;
      float v = 0.f;
      if (kh < 7 && kwp >= 1 && c < C) v = w_ref[(((int64_t)co * C + c) * 7 + kh) * 7 + (kwp - 1)];
      w224[i] = float_to_bf16(v);
    });
    return true;
  }
  const int KR = 8 * CH;
  par_for<k_stem_w224>(s, (int64_t)64 * 7 * KR, [=] __device__(int64_t i) {
    int e = (int)(i % KR);
    int64_t t = i / KR;
    int kh = (int)(t % 7);
    int co = (int)(t / 7);
    int kwp = e / CH, c = e % CH;
    float v = 0.f;
    if (kwp >= 1 && c < C) v = w_ref[(((int64_t)co * C + c) * 7 + kh) * 7 + (kwp - 1)];
    w224[i] = float_to_bf16(v);
  });
  return true;
}
bool stem_pack_weight_bf16(const float* w_ref, bf16* wp, int C, int Kp, lbc_stream_t s) {
  par_for<k_stem_pack>(s, (int64_t)64 * Kp, [=] __device__(int64_t i) {
    int k = (int)(i % Kp);
    int co = (int)(i / Kp);
    float v = 0.f;
    if (k < 49 * C) {
      int tap = k / C, c = k - tap * C;
      v = w_ref[((int64_t)co * C + c) * 49 + tap];
    }
    wp[i] = float_to_bf16(v);
  });
  return true;
}
bool stem_unpack_wgrad(const float* dw_col, float* dw_ref, int C, int Kp, lbc_stream_t s) {
  par_for<k_stem_unpack>(s, (int64_t)64 * C * 49, [=] __device__(int64_t i) {
    int tap = (int)(i % 49);
    int64_t t = i / 49;
    int c = (int)(t % C);
    int co = (int)(t / C);
    dw_ref[i] = dw_col[(int64_t)co * Kp + tap * C + c];
  });
  return true;
}
#endif

}  // namespace fast
}  // namespace lbc
