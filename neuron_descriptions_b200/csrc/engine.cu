// MILAN engine: weight ingestion (BN folding, re-layout, bf16 hi/lo split), workspace, the encode / decode /
// rerank pipelines, and the C ABI declared in include/milan_b200.h.
#include "milan_b200.h"

#include "conv_gemm.h"
#include "decode_fused.h"
#include "decoder.h"
#include "encoder.h"

#include <cuda_fp16.h>
#include <nvtx3/nvToolsExt.h>  // header-only; ranges cost nothing unless a profiler injects the NVTX library

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

using namespace milan;

namespace {

thread_local char g_err[1024] = "";

int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return 1;
}

#define CU(expr)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (expr);                                                                             \
    if (e_ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define RC(expr)                                                                                         \
  do {                                                                                                   \
    int rc_ = (expr);                                                                                    \
    if (rc_ != 0)                                                                                        \
      return fail("%s failed with code %d (%s) (%s:%d)", #expr, rc_, cudaGetErrorString((cudaError_t)rc_), \
                  __FILE__, __LINE__);                                                                   \
  } while (0)

// NVTX range over a pipeline phase ("milan.encode chunk 3", ...): shows the chunk pipeline in nsys / ncu timelines
// (tracing).
struct NvtxRange {
  explicit NvtxRange(const char* fmt, int a = 0, int b = 0) {
    char name[96];
    snprintf(name, sizeof name, fmt, a, b);
    nvtxRangePushA(name);
  }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

// Entry points run on the engine's device and leave the caller's current device as they found it (a process may hold
// tensors on several GPUs; torch tracks its own notion of the current device).
struct DeviceGuard {
  int prev = -1;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int device) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != device) err = cudaSetDevice(device);
    else if (err == cudaSuccess) prev = -1;  // nothing to restore
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

struct HostTensor {
  std::vector<float> data;
  std::vector<int64_t> shape;
  int64_t numel() const {
    int64_t n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
};

uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return static_cast<uint16_t>(u >> 16);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}
float bf2f(uint16_t h) {
  uint32_t u = static_cast<uint32_t>(h) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

struct SplitMat {  // K-major [rows][cols] bf16 planes on the device
  __nv_bfloat16* hi = nullptr;
  __nv_bfloat16* lo = nullptr;
  int rows = 0, cols = 0;
};

struct ConvLayer {
  std::string name;
  int cin, cout, ksize, stride;
  int dual_cb = 0;  // > 0: conv3 + downsample fused (build_dual_1x1_params); cin = planes, dual_cb = block input depth
  SplitMat w;
  float* bias = nullptr;  // folded BN shift, padded to 128
};

struct Plan {
  ConvGemmParams p;
  int block_n = 128;
  int epilogue = EPI_BF16;
};

constexpr int kRes[4] = {56, 28, 14, 7};
constexpr int kPlanes[4] = {64, 128, 256, 512};
constexpr double kBnEps = 1e-5;
constexpr int kSpatialKeys = 49;  // SpatialConvEncoder: 7x7 positions of layer4 (src/milan/encoders.py:235)

// torchvision ResNet variants (Bottleneck v1.5 with the stride on the 3x3, or BasicBlock), indexed by
// MILAN_ENCODER_RESNET* (include/milan_b200.h).
struct EncoderArch {
  const char* name;
  bool bottleneck;
  int blocks[4];
};
constexpr EncoderArch kArchs[] = {
    {"resnet101", true, {3, 4, 23, 3}},
    {"resnet50", true, {3, 4, 6, 3}},
    {"resnet18", false, {2, 2, 2, 2}},
    {"resnet34", false, {3, 4, 6, 3}},
    {"alexnet", false, {0, 0, 0, 0}},  // torchvision alexnet `features`: its own pipeline (encode_alexnet)
};
// AlexNet pyramid (src/milan/encoders.py:328-334): retained modules features.{0,3,6,8,10}; because their ReLUs
// are in-place and nethook retains `output.detach()` (a view, nethook.py:226-235), the retained maps are the
// POST-ReLU activations. Geometry: (in resolution, Cin, Cout, ksize) per conv; features.0 runs as an im2col GEMM.
struct AlexConv { const char* name; int res, cin, cout, ksize; };
constexpr AlexConv kAlexConvs[5] = {{"features.0", 55, kAlexK0, 64, 1}, {"features.3", 27, 64, 192, 5},
                                    {"features.6", 13, 192, 384, 3}, {"features.8", 13, 384, 256, 3},
                                    {"features.10", 13, 256, 256, 3}};
constexpr int kAlexMaskSizes[3] = {55, 27, 13};
constexpr int kAlexMaskOffset[3] = {0, 3025, 3025 + 729};
constexpr int kAlexMaskStride = 3025 + 729 + 169;
constexpr int kNumArchs = sizeof(kArchs) / sizeof(kArchs[0]);

}  // namespace

struct MilanEngine {
  MilanConfig cfg{};
  int device = 0;
  int num_sms = 148;
  bool split = true;
  bool finalized = false;
  std::map<std::string, HostTensor> pending;
  std::vector<void*> allocs;

  // ---- encoder weights
  SplitMat stem_w;
  float* bn1_alpha = nullptr;
  float* bn1_beta = nullptr;
  float mean[3] = {0, 0, 0}, stdv[3] = {1, 1, 1};
  EncoderArch arch = kArchs[0];
  int expansion = 4;             // stage output channels = planes * expansion
  bool spatial = false;          // SpatialConvEncoder: mask the image, emit the layer4 map
  bool has_decoder = true;       // false: encoder-only engine (no decoder tensors were provided)
  bool alexnet = false;
  bool fuse_downsample = true;   // MILAN_FUSE_DOWNSAMPLE=0 keeps conv3 and downsample as separate kernels (A/B runs)
  __nv_bfloat16 *axA0[2] = {}, *axAct[5][2] = {}, *axPool[2][2] = {};  // alexnet: im2col, conv outputs, max-pools
  int enc_out_per_image = 0;     // floats milan_encode writes per image
  std::vector<ConvLayer> convs;  // convs after the stem, execution order (103 for resnet101)
  // ---- encoder workspace (hi/lo planes)
  __nv_bfloat16 *stemA[2] = {}, *c1raw[2] = {}, *bufX[2] = {}, *bufY[2] = {}, *bufT1[2] = {}, *bufT2[2] = {},
                *bufDS[2] = {};
  float* mask_wts = nullptr;
  void* ones_masks = nullptr;
  std::map<int, std::vector<Plan>> enc_plans;  // by n_images

  // ---- decoder weights
  SplitMat Wk, Winit, W1, W2, W3, L0, L1, Lout;
  float *bk = nullptr, *binit = nullptr, *b1 = nullptr, *b2 = nullptr, *b3 = nullptr, *bl0 = nullptr, *bl1 = nullptr,
        *blout = nullptr;
  float* w_o = nullptr;
  float b_o = 0.f;
  float* emb = nullptr;
  float* lm_emb = nullptr;
  int ldv = 0;
  // ---- fused beam step / LM rerank (decode_fused.h): gate-interleaved LSTM weights, the concatenated head
  SplitMat W2p, Whead, L0h, L1p;
  float *b2p = nullptr, *bhead = nullptr, *bl0p = nullptr, *bl1p = nullptr;
  float* lm_table = nullptr;  // [V][4 Hl]: LM embedding folded through weight_ih_l0
  int vocab_tiles = 0, q_tiles = 0, n_seg = 0;
  bool fused_decode = true;   // MILAN_FUSED_DECODE=0: the round-1 step (one kernel per op) for A/B runs
  bool fused_ready = false, fused_lm_ready = false;
  float2 *partials = nullptr, *lm_partials = nullptr;
  float* lm_tgt = nullptr;
  int* beam_counters = nullptr;
  __nv_bfloat16 *lm_h0[2][2] = {}, *lm_h1[2][2] = {};  // [ping-pong][hi|lo]
  // ---- decoder workspace
  int Rmax = 0, Bmax = 0;
  size_t FRcap = 0;
  __nv_bfloat16 *feat[2] = {}, *pooled[2] = {}, *Alstm[2] = {}, *hnew[2] = {}, *Alm0[2] = {}, *Alm1[2] = {},
                *lmh[2] = {}, *lmnew0[2] = {}, *lmnew1[2] = {};
  float *kh = nullptr, *init_pre = nullptr, *qg = nullptr, *gates = nullptr, *hnew_f32 = nullptr, *c = nullptr,
        *cnew = nullptr, *logits = nullptr, *logits_lm = nullptr, *cand_val = nullptr, *last_lp = nullptr,
        *next_lp = nullptr, *lm_c0 = nullptr, *lm_c1 = nullptr, *lm_c0n = nullptr, *lm_c1n = nullptr,
        *lm_scores = nullptr, *feat_enc = nullptr, *out_scores = nullptr, *attn_ws = nullptr, *attn_acc = nullptr, *greedy_scores = nullptr, *lm_h_f32 = nullptr;
  int *cand_cls = nullptr, *backptr = nullptr, *hist_tok = nullptr, *hist_bp = nullptr, *group_T = nullptr;
  int *d_done = nullptr, *lm_skip = nullptr;  // device early-exit flags (beam loop / LM positions)
  long long *tok_cur = nullptr, *tok_next = nullptr, *seqs = nullptr, *lm_inputs = nullptr, *out_tokens = nullptr;
  std::map<std::pair<int, long long>, Plan> gemm_plans;  // (which, M)
  int host_T = 0;
  int host_chunk = 0, host_groups_per_chunk = 0;  // geometry of the last describe() call
  bool profiling_append = false;                  // encode(): keep the conv events of earlier chunks
  // ---- host staging for milan_describe_host
  // two staging sets: the exemplars of chunk i+1 are copied (copy_stream) while chunk i is encoded / decoded
  uint8_t *d_img_stage[2] = {nullptr, nullptr}, *d_mask_stage[2] = {nullptr, nullptr};
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
  // decode of chunk i runs on its own (high-priority) stream under the encoder of chunk i+1: two feature buffers
  cudaStream_t dec_stream = nullptr;
  cudaEvent_t ev_feat_ready[2] = {nullptr, nullptr}, ev_feat_free[2] = {nullptr, nullptr}, ev_dec_done = nullptr;
  float* feat_enc2 = nullptr;  // second feature buffer (feat_enc is the first)
  bool overlap_decode = true;  // MILAN_OVERLAP_DECODE=0: encode and decode of a chunk back to back on one stream
  int describe(const uint8_t* images, const uint8_t* masks, bool host_inputs, int n_neurons, int k, int strategy, int mi,
               int length, int beam, int group_size, float temperature, cudaStream_t st);
  long long* d_tokens_all = nullptr;  // describe_host results of every chunk (one D2H at the end)
  float* d_scores_all = nullptr;
  int* d_steps_all = nullptr;
  size_t results_cap = 0;
  // ---- profiling
  bool profiling = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> conv_events;
  size_t conv_events_used = 0;
  float prof_conv_ms = 0, prof_enc_ms = 0, prof_dec_ms = 0;
  long long prof_conv_launches = 0;

  template <class T>
  int dalloc(T** out, size_t n) {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, n * sizeof(T) + 256);
    if (e != cudaSuccess) return fail("cudaMalloc(%zu bytes) failed: %s", n * sizeof(T), cudaGetErrorString(e));
    allocs.push_back(p);
    *out = static_cast<T*>(p);
    return 0;
  }
  int dalloc2(__nv_bfloat16* out[2], size_t n) {
    if (dalloc(&out[0], n)) return 1;
    if (split) {
      if (dalloc(&out[1], n)) return 1;
    } else {
      out[1] = nullptr;
    }
    return 0;
  }
  const HostTensor* get(const std::string& name) const {
    auto it = pending.find(name);
    return it == pending.end() ? nullptr : &it->second;
  }
  int upload_f32(float** out, const std::vector<float>& v, size_t padded = 0) {
    std::vector<float> tmp(v);
    if (padded > tmp.size()) tmp.resize(padded, 0.f);
    if (dalloc(out, tmp.size())) return 1;
    CU(cudaMemcpy(*out, tmp.data(), tmp.size() * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
  }
  // fp16 = true: IEEE half (hi, lo) pairs for the decoder / LM matrices; false: bf16 pairs for the encoder.
  int upload_split(SplitMat* m, const std::vector<float>& w, int rows, int cols, bool fp16 = false) {
    std::vector<uint16_t> hi(w.size()), lo(w.size());
    for (size_t i = 0; i < w.size(); ++i) {
      if (fp16) {
        const __half h = __float2half_rn(w[i]);
        const __half l = __float2half_rn(w[i] - __half2float(h));
        memcpy(&hi[i], &h, 2);
        memcpy(&lo[i], &l, 2);
      } else {
        hi[i] = f2bf(w[i]);
        lo[i] = f2bf(w[i] - bf2f(hi[i]));
      }
    }
    m->rows = rows;
    m->cols = cols;
    if (dalloc(&m->hi, w.size())) return 1;
    CU(cudaMemcpy(m->hi, hi.data(), hi.size() * 2, cudaMemcpyHostToDevice));
    if (split) {
      if (dalloc(&m->lo, w.size())) return 1;
      CU(cudaMemcpy(m->lo, lo.data(), lo.size() * 2, cudaMemcpyHostToDevice));
    }
    return 0;
  }

  int finalize_encoder();
  int finalize_encoder_alexnet();
  int build_alexnet_plans(int n, std::vector<Plan>** out);
  int encode_alexnet(const void* d_images, const void* d_masks, int n, int dtype, float* d_out, cudaStream_t st);
  int finalize_decoder();
  int alloc_workspace();
  int build_encoder_plans(int n, std::vector<Plan>** out);
  int encode(const void* d_images, const void* d_masks, int n, int dtype, float* d_out, cudaStream_t st);
  int run_conv(const Plan& pl, cudaStream_t st);
  int collect_conv_events(cudaStream_t st);

  int gemm(int which, long long M, const SplitMat& W, const float* bias, const __nv_bfloat16* a_hi,
           const __nv_bfloat16* a_lo, long long a_pitch, int K, float* out, long long ldc, cudaStream_t st,
           const int* skip = nullptr);
  int prepare_features(const float* d_features, int Bf, int n_keys, cudaStream_t st);
  int step_core(int R, int rpf, int n_keys, const float* d_features, const long long* d_tokens, float* attn_out,
                long long attn_pitch, cudaStream_t st, const int* skip = nullptr);
  int lm_reset(int M, cudaStream_t st);
  int lm_step_core(int M, const long long* d_tokens, bool to_new, cudaStream_t st, const int* skip = nullptr);
  int decode_greedy(const float* d_features, int B, int n_keys, int length, int mi, float temperature,
                    const long long* d_forced, long long* d_tokens_out, float* d_scores_out, float* d_pred_out,
                    float* d_attn_out, cudaStream_t st);
  int decode_beam(const float* d_features, int B, int n_keys, int length, int beam, int group_size, int rerank, int mi,
                  float temperature, long long* d_beam_tokens, float* d_beam_scores, int* d_group_steps,
                  long long* d_tokens_out, float* d_scores_out, float* d_lm_scores_out, cudaStream_t st);
  int lm_score_seqs(const long long* d_seqs, int M, int length, int beam, int group_size, cudaStream_t st);
  // fused path
  bool use_fused_beam(int n_keys, int mi, int beam) const;
  int fused_gemm(int which, long long M, const ConvGemmParams** out, int K, const SplitMat& W, const float* bias,
                 const __nv_bfloat16* a_hi, const __nv_bfloat16* a_lo, long long a_pitch,
                 const __nv_bfloat16* a1_hi = nullptr, const __nv_bfloat16* a1_lo = nullptr, long long a1_pitch = 0,
                 int K1 = 0);
  int run_fused(const ConvGemmParams& p, int epilogue, cudaStream_t st, const int* skip);
  int beam_steps_fused(const float* d_features, int B, int n_keys, int length, int beam, cudaStream_t st);
  int lm_score_seqs_fused(const long long* d_seqs, int M, int length, int beam, int group_size, cudaStream_t st);
};

// ============================================================================ weight ingestion
int MilanEngine::finalize_encoder() {
  const std::string pre = "encoder.encoder.model.";
  if (const HostTensor* m = get("encoder.mean")) {
    if (m->numel() != 3) return fail("encoder.mean must have 3 elements");
    for (int i = 0; i < 3; ++i) mean[i] = m->data[i];
  } else {
    return fail("missing tensor encoder.mean");
  }
  if (const HostTensor* s = get("encoder.std")) {
    for (int i = 0; i < 3; ++i) stdv[i] = s->data[i];
  } else {
    return fail("missing tensor encoder.std");
  }
  if (alexnet) return finalize_encoder_alexnet();
  // stem: [64][3][7][7] -> [64][256] in the window order of build_stem_params; no BN fold (raw conv1 is pooled).
  const HostTensor* w = get(pre + "conv1.weight");
  if (w == nullptr || w->numel() != 64 * 3 * 49) return fail("missing/invalid %sconv1.weight", pre.c_str());
  {
    std::vector<float> packed(64 * kStemKTotal, 0.f);
    pack_stem_weights(w->data.data(), packed.data());
    if (upload_split(&stem_w, packed, 64, kStemKTotal)) return 1;
  }
  auto bn_fold = [&](const std::string& bn, int c, std::vector<double>* scale, std::vector<double>* shift) -> int {
    const HostTensor *g = get(bn + ".weight"), *b = get(bn + ".bias"), *m = get(bn + ".running_mean"),
                     *v = get(bn + ".running_var");
    if (!g || !b || !m || !v) return fail("missing BN tensors for %s", bn.c_str());
    if (g->numel() != c) return fail("BN %s has %lld channels, expected %d", bn.c_str(), (long long)g->numel(), c);
    scale->resize(c);
    shift->resize(c);
    for (int i = 0; i < c; ++i) {
      // torch batch_norm (eval): invstd = 1/sqrt(var + eps); y = x*(w*invstd) + (b - mean*w*invstd)
      const double invstd = 1.0 / std::sqrt(static_cast<double>(v->data[i]) + kBnEps);
      (*scale)[i] = g->data[i] * invstd;
      (*shift)[i] = b->data[i] - m->data[i] * (*scale)[i];
    }
    return 0;
  };
  {
    std::vector<double> sc, sh;
    if (bn_fold(pre + "bn1", 64, &sc, &sh)) return 1;
    std::vector<float> a(64), b(64);
    for (int i = 0; i < 64; ++i) {
      a[i] = static_cast<float>(sc[i]);
      b[i] = static_cast<float>(sh[i]);
    }
    if (upload_f32(&bn1_alpha, a)) return 1;
    if (upload_f32(&bn1_beta, b)) return 1;
  }
  auto add_conv = [&](const std::string& conv, const std::string& bn, int cin, int cout, int ks, int stride) -> int {
    const HostTensor* cw = get(pre + conv + ".weight");
    if (cw == nullptr || cw->numel() != static_cast<int64_t>(cout) * cin * ks * ks)
      return fail("missing/invalid %s%s.weight", pre.c_str(), conv.c_str());
    std::vector<double> sc, sh;
    if (bn_fold(pre + bn, cout, &sc, &sh)) return 1;
    const int taps = ks * ks;
    std::vector<float> packed(static_cast<size_t>(cout) * taps * cin);
    for (int co = 0; co < cout; ++co)
      for (int ci = 0; ci < cin; ++ci)
        for (int t = 0; t < taps; ++t)
          packed[(static_cast<size_t>(co) * taps + t) * cin + ci] =
              static_cast<float>(cw->data[(static_cast<size_t>(co) * cin + ci) * taps + t] * sc[co]);
    ConvLayer L;
    L.name = conv;
    L.cin = cin; L.cout = cout; L.ksize = ks; L.stride = stride;
    if (upload_split(&L.w, packed, cout, taps * cin)) return 1;
    std::vector<float> bias(cout);
    for (int i = 0; i < cout; ++i) bias[i] = static_cast<float>(sh[i]);
    if (upload_f32(&L.bias, bias, (cout + 127) / 128 * 128)) return 1;
    convs.push_back(L);
    return 0;
  };
  // First bottleneck of every stage: relu(bn3(conv3(t)) + bn_ds(downsample(x))) as ONE GEMM over K = planes + inplanes
  // ([W3*s3 | Wds*sds], bias = shift3 + shift_ds): the downsample output never exists as a tensor.
  auto add_dual = [&](const std::string& blk, int planes, int inplanes, int stride) -> int {
    const int cout = planes * 4;
    const HostTensor* w3 = get(pre + blk + ".conv3.weight");
    const HostTensor* wd = get(pre + blk + ".downsample.0.weight");
    if (w3 == nullptr || w3->numel() != static_cast<int64_t>(cout) * planes)
      return fail("missing/invalid %s%s.conv3.weight", pre.c_str(), blk.c_str());
    if (wd == nullptr || wd->numel() != static_cast<int64_t>(cout) * inplanes)
      return fail("missing/invalid %s%s.downsample.0.weight", pre.c_str(), blk.c_str());
    std::vector<double> s3, h3, sd_, hd;
    if (bn_fold(pre + blk + ".bn3", cout, &s3, &h3)) return 1;
    if (bn_fold(pre + blk + ".downsample.1", cout, &sd_, &hd)) return 1;
    const int K = planes + inplanes;
    std::vector<float> packed(static_cast<size_t>(cout) * K);
    std::vector<float> bias(cout);
    for (int co = 0; co < cout; ++co) {
      for (int ci = 0; ci < planes; ++ci)
        packed[static_cast<size_t>(co) * K + ci] = static_cast<float>(w3->data[static_cast<size_t>(co) * planes + ci] * s3[co]);
      for (int ci = 0; ci < inplanes; ++ci)
        packed[static_cast<size_t>(co) * K + planes + ci] =
            static_cast<float>(wd->data[static_cast<size_t>(co) * inplanes + ci] * sd_[co]);
      bias[co] = static_cast<float>(h3[co] + hd[co]);
    }
    ConvLayer L;
    L.name = blk + ".conv3+downsample";
    L.cin = planes; L.cout = cout; L.ksize = 1; L.stride = stride; L.dual_cb = inplanes;
    if (upload_split(&L.w, packed, cout, K)) return 1;
    if (upload_f32(&L.bias, bias, (cout + 127) / 128 * 128)) return 1;
    convs.push_back(L);
    return 0;
  };
  int inplanes = 64;
  for (int li = 0; li < 4; ++li) {
    for (int bi = 0; bi < arch.blocks[li]; ++bi) {
      const int planes = kPlanes[li];
      const int stride = (bi == 0 && li > 0) ? 2 : 1;
      char buf[64];
      snprintf(buf, sizeof buf, "layer%d.%d", li + 1, bi);
      const std::string b(buf);
      if (arch.bottleneck && bi == 0 && fuse_downsample) {
        if (add_conv(b + ".conv1", b + ".bn1", inplanes, planes, 1, 1)) return 1;
        if (add_conv(b + ".conv2", b + ".bn2", planes, planes, 3, stride)) return 1;
        if (add_dual(b, planes, inplanes, stride)) return 1;
        inplanes = planes * expansion;
        continue;
      }
      if (arch.bottleneck) {
        if (add_conv(b + ".conv1", b + ".bn1", inplanes, planes, 1, 1)) return 1;
        if (add_conv(b + ".conv2", b + ".bn2", planes, planes, 3, stride)) return 1;
        if (add_conv(b + ".conv3", b + ".bn3", planes, planes * 4, 1, 1)) return 1;
      } else {
        if (add_conv(b + ".conv1", b + ".bn1", inplanes, planes, 3, stride)) return 1;
        if (add_conv(b + ".conv2", b + ".bn2", planes, planes, 3, 1)) return 1;
      }
      if (stride != 1 || inplanes != planes * expansion) {
        if (add_conv(b + ".downsample.0", b + ".downsample.1", inplanes, planes * expansion, 1, stride)) return 1;
      }
      inplanes = planes * expansion;
    }
  }
  return 0;
}

int MilanEngine::finalize_encoder_alexnet() {
  const std::string pre = "encoder.encoder.model.";
  for (const AlexConv& a : kAlexConvs) {
    const bool first = a.ksize == 1;  // features.0: [64][3][11][11], natural (c, r, s) order = im2col column order
    const int taps = first ? 1 : a.ksize * a.ksize;
    const int cin_w = first ? 3 * 11 * 11 : a.cin;
    const HostTensor* cw = get(pre + a.name + ".weight");
    const HostTensor* cb = get(pre + a.name + ".bias");
    if (cw == nullptr || cw->numel() != static_cast<int64_t>(a.cout) * cin_w * taps)
      return fail("missing/invalid %s%s.weight", pre.c_str(), a.name);
    if (cb == nullptr || cb->numel() != a.cout) return fail("missing/invalid %s%s.bias", pre.c_str(), a.name);
    std::vector<float> packed(static_cast<size_t>(a.cout) * taps * a.cin, 0.f);
    for (int co = 0; co < a.cout; ++co) {
      if (first) {
        for (int k = 0; k < cin_w; ++k) packed[static_cast<size_t>(co) * a.cin + k] = cw->data[static_cast<size_t>(co) * cin_w + k];
      } else {
        for (int ci = 0; ci < a.cin; ++ci)
          for (int t = 0; t < taps; ++t)
            packed[(static_cast<size_t>(co) * taps + t) * a.cin + ci] = cw->data[(static_cast<size_t>(co) * a.cin + ci) * taps + t];
      }
    }
    ConvLayer L;
    L.name = a.name;
    L.cin = a.cin; L.cout = a.cout; L.ksize = a.ksize; L.stride = 1;
    if (upload_split(&L.w, packed, a.cout, taps * a.cin)) return 1;
    if (upload_f32(&L.bias, cb->data, (a.cout + 127) / 128 * 128)) return 1;
    convs.push_back(L);
  }
  return 0;
}

int MilanEngine::build_alexnet_plans(int n, std::vector<Plan>** out) {
  auto it = enc_plans.find(n);
  if (it != enc_plans.end()) {
    *out = &it->second;
    return 0;
  }
  std::vector<Plan> plans;
  const int sp = split ? 1 : 0;
  for (int i = 0; i < 5; ++i) {
    const AlexConv& a = kAlexConvs[i];
    const ConvLayer& L = convs[i];
    __nv_bfloat16** in = i == 0 ? axA0 : (i == 1 ? axPool[0] : (i == 2 ? axPool[1] : axAct[i - 1]));
    Plan pl;
    ConvDesc d{n, a.res, a.res, L.cin, L.cout, L.ksize, 1};
    ConvIO io{};
    io.in_hi = in[0]; io.in_lo = in[1];
    io.w_hi = L.w.hi; io.w_lo = L.w.lo;
    io.bias = L.bias;
    io.out_hi = axAct[i][0]; io.out_lo = axAct[i][1];
    io.relu = 1;  // in-place ReLU reaches the retained tensor (see kAlexConvs)
    if (build_conv_params(&pl.p, d, io, sp, &pl.block_n)) return fail("plan %s: %s", L.name.c_str(), tmap_last_error());
    plans.push_back(pl);
  }
  auto ins = enc_plans.emplace(n, std::move(plans));
  *out = &ins.first->second;
  return 0;
}

// PyramidConvEncoder('alexnet').forward (src/milan/encoders.py:286-320 with the layer table of :328-334).
int MilanEngine::encode_alexnet(const void* d_images, const void* d_masks, int n, int dtype, float* d_out,
                                cudaStream_t st) {
  std::vector<Plan>* plans = nullptr;
  if (build_alexnet_plans(n, &plans)) return 1;
  const int F = cfg.feature_size, sp = split ? 1 : 0;
  if (profiling && !profiling_append) conv_events_used = 0;
  const void* masks = d_masks;
  int mask_dtype = dtype;
  if (masks == nullptr) {
    if (ones_masks == nullptr) {
      uint8_t* p = nullptr;
      if (dalloc(&p, static_cast<size_t>(cfg.max_images) * 224 * 224)) return 1;
      CU(cudaMemset(p, 1, static_cast<size_t>(cfg.max_images) * 224 * 224));
      ones_masks = p;
    }
    masks = ones_masks;
    mask_dtype = MILAN_DTYPE_U8;
  }
  for (int l = 0; l < 3; ++l)
    RC(launch_mask_resize(masks, mask_dtype, n, kAlexMaskSizes[l], mask_wts + kAlexMaskOffset[l], kAlexMaskStride, st));
  RC(launch_alexnet_im2col(d_images, dtype, n, axA0[0], axA0[1], mean, stdv, sp, st));
  int feat_off = 0;
  for (int i = 0; i < 5; ++i) {
    const AlexConv& a = kAlexConvs[i];
    if (run_conv((*plans)[i], st)) return 1;
    const int level = i < 2 ? i : 2;
    RC(launch_masked_pool(axAct[i][0], axAct[i][1], mask_wts + kAlexMaskOffset[level], kAlexMaskStride, n, a.res * a.res,
                          a.cout, d_out + feat_off, F, st));
    feat_off += a.cout;
    if (i < 2) RC(launch_maxpool3x3s2(axAct[i][0], axAct[i][1], n, a.res, a.res, a.cout, axPool[i][0], axPool[i][1], st));
  }
  return 0;
}

int MilanEngine::finalize_decoder() {
  const int V = cfg.vocab_size, E = cfg.embedding_size, H = cfg.hidden_size, A = cfg.attention_size,
            F = cfg.feature_size;
  ldv = (V + 3) / 4 * 4;
  auto need = [&](const char* name, int64_t numel) -> const HostTensor* {
    const HostTensor* t = get(name);
    if (t == nullptr) {
      fail("missing tensor %s", name);
      return nullptr;
    }
    if (t->numel() != numel) {
      fail("tensor %s has %lld elements, expected %lld", name, (long long)t->numel(), (long long)numel);
      return nullptr;
    }
    return t;
  };
  if ((E + F + H) % 64 || H % 64 || F % 64 || A % 64) return fail("decoder dims must be multiples of 64");
  const HostTensor *wq = need("attend.query_to_hidden.weight", (int64_t)A * H),
                   *bq = need("attend.query_to_hidden.bias", A),
                   *wk = need("attend.key_to_hidden.weight", (int64_t)A * F),
                   *bkk = need("attend.key_to_hidden.bias", A), *wo = need("attend.output.0.weight", A),
                   *bo = need("attend.output.0.bias", 1), *wg = need("feature_gate.0.weight", (int64_t)F * H),
                   *bg = need("feature_gate.0.bias", F), *wih = need("lstm.weight_ih", (int64_t)4 * H * (E + F)),
                   *whh = need("lstm.weight_hh", (int64_t)4 * H * H), *bih = need("lstm.bias_ih", 4 * H),
                   *bhh = need("lstm.bias_hh", 4 * H), *wout = need("output.1.weight", (int64_t)V * H),
                   *bout = need("output.1.bias", V), *wh = need("init_h.0.weight", (int64_t)H * F),
                   *bh = need("init_h.0.bias", H), *wc = need("init_c.0.weight", (int64_t)H * F),
                   *bc = need("init_c.0.bias", H), *em = need("embedding.weight", (int64_t)V * E);
  if (!wq || !bq || !wk || !bkk || !wo || !bo || !wg || !bg || !wih || !whh || !bih || !bhh || !wout || !bout ||
      !wh || !bh || !wc || !bc || !em)
    return 1;
  if (upload_split(&Wk, wk->data, A, F, true)) return 1;
  if (upload_f32(&bk, bkk->data, (A + 127) / 128 * 128)) return 1;
  {  // Winit = [W_h; W_c]
    std::vector<float> w(wh->data);
    w.insert(w.end(), wc->data.begin(), wc->data.end());
    std::vector<float> b(bh->data);
    b.insert(b.end(), bc->data.begin(), bc->data.end());
    if (upload_split(&Winit, w, 2 * H, F, true)) return 1;
    if (upload_f32(&binit, b, (2 * H + 127) / 128 * 128)) return 1;
  }
  {  // W1 = [W_q; W_g]
    std::vector<float> w(wq->data);
    w.insert(w.end(), wg->data.begin(), wg->data.end());
    std::vector<float> b(bq->data);
    b.insert(b.end(), bg->data.begin(), bg->data.end());
    if (upload_split(&W1, w, A + F, H, true)) return 1;
    if (upload_f32(&b1, b, (A + F + 127) / 128 * 128)) return 1;
  }
  {  // W2 = [W_ih | W_hh], bias = b_ih + b_hh
    const int K = E + F + H;
    std::vector<float> w(static_cast<size_t>(4) * H * K);
    for (int r = 0; r < 4 * H; ++r) {
      memcpy(&w[static_cast<size_t>(r) * K], &wih->data[static_cast<size_t>(r) * (E + F)], sizeof(float) * (E + F));
      memcpy(&w[static_cast<size_t>(r) * K + E + F], &whh->data[static_cast<size_t>(r) * H], sizeof(float) * H);
    }
    std::vector<float> b(4 * H);
    for (int i = 0; i < 4 * H; ++i) b[i] = bih->data[i] + bhh->data[i];
    if (upload_split(&W2, w, 4 * H, K, true)) return 1;
    if (upload_f32(&b2, b)) return 1;
    // fused step: rows interleaved so that row 4u + g is gate g (i, f, g, o) of hidden unit u (conv_gemm.h, EPI_LSTM)
    std::vector<float> wp(w.size()), bp(b.size());
    for (int u = 0; u < H; ++u)
      for (int g = 0; g < 4; ++g) {
        memcpy(&wp[static_cast<size_t>(4 * u + g) * K], &w[static_cast<size_t>(g * H + u) * K], sizeof(float) * K);
        bp[4 * u + g] = b[g * H + u];
      }
    if (upload_split(&W2p, wp, 4 * H, K, true)) return 1;
    if (upload_f32(&b2p, bp)) return 1;
  }
  {  // Whead = [W_out; 0 ...; W_q; 0 ...; W_g]: every linear applied to h', each section starting on a 128-row tile
    vocab_tiles = (V + 127) / 128;
    q_tiles = (A + 127) / 128;
    n_seg = 2 * vocab_tiles;
    const size_t rows = static_cast<size_t>(vocab_tiles + q_tiles) * 128 + F;
    const size_t padded = (rows + 127) / 128 * 128;
    std::vector<float> w(rows * H, 0.f), b(padded, 0.f);
    memcpy(&w[0], wout->data.data(), sizeof(float) * V * H);
    memcpy(&b[0], bout->data.data(), sizeof(float) * V);
    const size_t q0 = static_cast<size_t>(vocab_tiles) * 128, g0 = q0 + static_cast<size_t>(q_tiles) * 128;
    memcpy(&w[q0 * H], wq->data.data(), sizeof(float) * A * H);
    memcpy(&b[q0], bq->data.data(), sizeof(float) * A);
    memcpy(&w[g0 * H], wg->data.data(), sizeof(float) * F * H);
    memcpy(&b[g0], bg->data.data(), sizeof(float) * F);
    if (upload_split(&Whead, w, static_cast<int>(rows), H, true)) return 1;
    if (upload_f32(&bhead, b)) return 1;
  }
  if (upload_split(&W3, wout->data, V, H, true)) return 1;
  if (upload_f32(&b3, bout->data, (V + 127) / 128 * 128)) return 1;
  if (upload_f32(&w_o, wo->data)) return 1;
  b_o = bo->data[0];
  if (upload_f32(&emb, em->data)) return 1;

  if (cfg.has_lm) {
    const int El = cfg.lm_embedding_size, Hl = cfg.lm_hidden_size;
    if ((El + Hl) % 64 || Hl % 64) return fail("LM dims must be multiples of 64");
    const HostTensor *le = need("lm.embedding.weight", (int64_t)V * El),
                     *i0 = need("lm.lstm.weight_ih_l0", (int64_t)4 * Hl * El),
                     *h0 = need("lm.lstm.weight_hh_l0", (int64_t)4 * Hl * Hl),
                     *bi0 = need("lm.lstm.bias_ih_l0", 4 * Hl), *bh0 = need("lm.lstm.bias_hh_l0", 4 * Hl),
                     *i1 = need("lm.lstm.weight_ih_l1", (int64_t)4 * Hl * Hl),
                     *h1 = need("lm.lstm.weight_hh_l1", (int64_t)4 * Hl * Hl),
                     *bi1 = need("lm.lstm.bias_ih_l1", 4 * Hl), *bh1 = need("lm.lstm.bias_hh_l1", 4 * Hl),
                     *lo = need("lm.output.0.weight", (int64_t)V * Hl), *lb = need("lm.output.0.bias", V);
    if (!le || !i0 || !h0 || !bi0 || !bh0 || !i1 || !h1 || !bi1 || !bh1 || !lo || !lb) return 1;
    auto cat = [&](const HostTensor* a, int ka, const HostTensor* b, int kb, SplitMat* out) -> int {
      std::vector<float> w(static_cast<size_t>(4) * Hl * (ka + kb));
      for (int r = 0; r < 4 * Hl; ++r) {
        memcpy(&w[static_cast<size_t>(r) * (ka + kb)], &a->data[static_cast<size_t>(r) * ka], sizeof(float) * ka);
        memcpy(&w[static_cast<size_t>(r) * (ka + kb) + ka], &b->data[static_cast<size_t>(r) * kb], sizeof(float) * kb);
      }
      return upload_split(out, w, 4 * Hl, ka + kb, true);
    };
    if (cat(i0, El, h0, Hl, &L0)) return 1;
    if (cat(i1, Hl, h1, Hl, &L1)) return 1;
    std::vector<float> b0(4 * Hl), b1v(4 * Hl);
    for (int i = 0; i < 4 * Hl; ++i) {
      b0[i] = bi0->data[i] + bh0->data[i];
      b1v[i] = bi1->data[i] + bh1->data[i];
    }
    if (upload_f32(&bl0, b0)) return 1;
    if (upload_f32(&bl1, b1v)) return 1;
    if (upload_split(&Lout, lo->data, V, Hl, true)) return 1;
    if (upload_f32(&blout, lb->data, (V + 127) / 128 * 128)) return 1;
    if (upload_f32(&lm_emb, le->data)) return 1;
    {  // fused LM: gate-interleaved rows; layer 0 sees its input through lm_table (gathered add in the epilogue)
      std::vector<float> w0(static_cast<size_t>(4) * Hl * Hl), w1(static_cast<size_t>(4) * Hl * 2 * Hl), p0(4 * Hl),
          p1(4 * Hl);
      for (int u = 0; u < Hl; ++u)
        for (int g = 0; g < 4; ++g) {
          const size_t dst = 4 * u + g, src = static_cast<size_t>(g) * Hl + u;
          memcpy(&w0[dst * Hl], &h0->data[src * Hl], sizeof(float) * Hl);
          memcpy(&w1[dst * 2 * Hl], &i1->data[src * Hl], sizeof(float) * Hl);
          memcpy(&w1[dst * 2 * Hl + Hl], &h1->data[src * Hl], sizeof(float) * Hl);
          p0[dst] = b0[src];
          p1[dst] = b1v[src];
        }
      if (upload_split(&L0h, w0, 4 * Hl, Hl, true)) return 1;
      if (upload_split(&L1p, w1, 4 * Hl, 2 * Hl, true)) return 1;
      if (upload_f32(&bl0p, p0)) return 1;
      if (upload_f32(&bl1p, p1)) return 1;
      float* d_wih = nullptr;
      if (upload_f32(&d_wih, i0->data)) return 1;
      if (dalloc(&lm_table, static_cast<size_t>(V) * 4 * Hl)) return 1;
      RC(launch_lm_input_table(d_wih, lm_emb, V, El, Hl, lm_table, nullptr));
      CU(cudaDeviceSynchronize());
    }
  }
  return 0;
}

int MilanEngine::alloc_workspace() {
  const int V = cfg.vocab_size, E = cfg.embedding_size, H = cfg.hidden_size, A = cfg.attention_size,
            F = cfg.feature_size, Kk = cfg.max_keys, L = cfg.max_length;
  (void)V;
  if (cfg.has_encoder && alexnet) {
    const size_t n = cfg.max_images;
    if (dalloc2(axA0, n * 3025 * kAlexK0)) return 1;
    for (int i = 0; i < 5; ++i)
      if (dalloc2(axAct[i], n * kAlexConvs[i].res * kAlexConvs[i].res * kAlexConvs[i].cout)) return 1;
    if (dalloc2(axPool[0], n * 27 * 27 * 64)) return 1;
    if (dalloc2(axPool[1], n * 13 * 13 * 192)) return 1;
    if (dalloc(&mask_wts, n * kAlexMaskStride)) return 1;
    if (dalloc(&feat_enc, std::max(n * enc_out_per_image, static_cast<size_t>(cfg.max_neurons) * Kk * F))) return 1;
    if (dalloc(&feat_enc2, std::max(n * enc_out_per_image, static_cast<size_t>(cfg.max_neurons) * Kk * F))) return 1;
    for (int b = 0; b < 2; ++b) {
      if (dalloc(&d_img_stage[b], n * 3 * 224 * 224)) return 1;
      if (dalloc(&d_mask_stage[b], n * 224 * 224)) return 1;
    }
  } else if (cfg.has_encoder) {
    const size_t n = cfg.max_images;
    if (dalloc2(stemA, n * kStemPadH * kStemPadW * 4)) return 1;
    if (dalloc2(c1raw, n * 12544 * 64)) return 1;
    // largest tensors: stage outputs at 56x56 (64 * expansion channels), bottleneck conv1 of layer2.0 (128 @ 56x56)
    const size_t stage = 3136 * 64 * static_cast<size_t>(expansion);
    if (dalloc2(bufX, n * stage)) return 1;
    if (dalloc2(bufY, n * stage)) return 1;
    if (dalloc2(bufT1, n * 3136 * (arch.bottleneck ? 128 : 64))) return 1;
    if (arch.bottleneck && dalloc2(bufT2, n * 3136 * 64)) return 1;
    if (dalloc2(bufDS, n * stage)) return 1;
    if (dalloc(&mask_wts, n * kMaskPyramidSize)) return 1;
    if (dalloc(&feat_enc, std::max(n * enc_out_per_image, static_cast<size_t>(cfg.max_neurons) * Kk * F))) return 1;
    if (dalloc(&feat_enc2, std::max(n * enc_out_per_image, static_cast<size_t>(cfg.max_neurons) * Kk * F))) return 1;
    for (int b = 0; b < 2; ++b) {
      if (dalloc(&d_img_stage[b], n * 3 * 224 * 224)) return 1;
      if (dalloc(&d_mask_stage[b], n * 224 * 224)) return 1;
    }
  }
  if (!has_decoder) return 0;
  Bmax = cfg.max_neurons;
  Rmax = cfg.max_neurons * cfg.max_beam;
  FRcap = static_cast<size_t>(Rmax) * Kk;  // feature rows: milan_step may bring one feature set per row
  const size_t R = Rmax, B = Bmax;
  if (dalloc2(feat, FRcap * F)) return 1;
  if (dalloc(&kh, FRcap * A)) return 1;
  if (dalloc2(pooled, R * F)) return 1;
  if (dalloc(&init_pre, R * 2 * H)) return 1;
  if (dalloc2(Alstm, R * (E + F + H))) return 1;
  if (dalloc(&qg, R * (A + F))) return 1;
  if (dalloc(&attn_ws, R * std::max(Kk, 64))) return 1;
  if (Kk > 16 && dalloc(&attn_acc, R * F)) return 1;
  if (dalloc(&gates, R * 4 * std::max(H, cfg.lm_hidden_size))) return 1;
  if (dalloc2(hnew, R * H)) return 1;
  if (dalloc(&hnew_f32, R * H)) return 1;
  if (dalloc(&c, R * H)) return 1;
  if (dalloc(&cnew, R * H)) return 1;
  if (dalloc(&logits, R * ldv)) return 1;
  if (dalloc(&cand_val, R * cfg.max_beam)) return 1;
  if (dalloc(&cand_cls, R * cfg.max_beam)) return 1;
  if (dalloc(&last_lp, R)) return 1;
  if (dalloc(&next_lp, R)) return 1;
  if (dalloc(&backptr, R)) return 1;
  if (dalloc(&hist_tok, R * L)) return 1;
  if (dalloc(&hist_bp, R * L)) return 1;
  if (dalloc(&group_T, B + 1)) return 1;
  if (dalloc(&d_done, 4)) return 1;
  if (dalloc(&lm_skip, L + 1)) return 1;
  CU(cudaMemset(d_done, 0, 4 * sizeof(int)));
  CU(cudaMemset(lm_skip, 0, (L + 1) * sizeof(int)));
  if (dalloc(&partials, R * std::max(n_seg, 1))) return 1;
  if (dalloc(&beam_counters, 4)) return 1;
  CU(cudaMemset(beam_counters, 0, 4 * sizeof(int)));
  fused_ready = split && A % 128 == 0 && H % 64 == 0 && F % 2 == 0 && E % 2 == 0 &&
                beam_select_smem_bytes(cfg.max_beam, cfg.max_beam, V) <= 227 * 1024;
  if (dalloc(&tok_cur, R)) return 1;
  if (dalloc(&tok_next, R)) return 1;
  if (dalloc(&seqs, R * (L + 1))) return 1;
  if (dalloc(&out_tokens, B * L)) return 1;
  if (dalloc(&greedy_scores, R)) return 1;
  if (dalloc(&out_scores, R)) return 1;
  if (dalloc(&lm_scores, R)) return 1;
  if (cfg.has_lm) {
    const int El = cfg.lm_embedding_size, Hl = cfg.lm_hidden_size;
    if (dalloc(&logits_lm, R * ldv)) return 1;
    if (dalloc2(Alm0, R * (El + Hl))) return 1;
    if (dalloc2(Alm1, R * 2 * Hl)) return 1;
    if (dalloc2(lmh, R * Hl)) return 1;
    if (dalloc2(lmnew0, R * Hl)) return 1;
    if (dalloc2(lmnew1, R * Hl)) return 1;
    if (dalloc(&lm_c0, R * Hl)) return 1;
    if (dalloc(&lm_c1, R * Hl)) return 1;
    if (dalloc(&lm_c0n, R * Hl)) return 1;
    if (dalloc(&lm_c1n, R * Hl)) return 1;
    if (dalloc(&lm_inputs, R)) return 1;
    if (dalloc(&lm_h_f32, R * Hl * 2)) return 1;
    for (int pp = 0; pp < 2; ++pp) {
      if (dalloc2(lm_h0[pp], R * Hl)) return 1;
      if (dalloc2(lm_h1[pp], R * Hl)) return 1;
    }
    if (dalloc(&lm_partials, R * L * std::max(n_seg, 1))) return 1;
    if (dalloc(&lm_tgt, R * L)) return 1;
    fused_lm_ready = split && Hl % 64 == 0;
  }
  return 0;
}

// ============================================================================ encoder
int MilanEngine::run_conv(const Plan& pl, cudaStream_t st) {
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (profiling) {
    if (conv_events_used == conv_events.size()) {
      cudaEvent_t a, b;
      CU(cudaEventCreate(&a));
      CU(cudaEventCreate(&b));
      conv_events.emplace_back(a, b);
    }
    e0 = conv_events[conv_events_used].first;
    e1 = conv_events[conv_events_used].second;
    ++conv_events_used;
    CU(cudaEventRecord(e0, st));
  }
  RC(launch_conv_gemm(pl.p, pl.block_n, split ? 1 : 0, pl.epilogue, num_sms, st));
  if (profiling) CU(cudaEventRecord(e1, st));
  return 0;
}

int MilanEngine::collect_conv_events(cudaStream_t st) {
  if (!profiling) return 0;
  CU(cudaStreamSynchronize(st));
  float total = 0;
  for (size_t i = 0; i < conv_events_used; ++i) {
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, conv_events[i].first, conv_events[i].second));
    total += ms;
  }
  prof_conv_ms += total;
  prof_conv_launches += static_cast<long long>(conv_events_used);
  conv_events_used = 0;
  return 0;
}

int MilanEngine::build_encoder_plans(int n, std::vector<Plan>** out) {
  auto it = enc_plans.find(n);
  if (it != enc_plans.end()) {
    *out = &it->second;
    return 0;
  }
  std::vector<Plan> plans;
  const int sp = split ? 1 : 0;
  {  // stem: im2col-free implicit GEMM over the padded NHWC4 image -> raw conv1 (no bias, no ReLU)
    Plan pl;
    pl.block_n = 64;
    if (build_stem_params(&pl.p, n, stemA[0], stemA[1], stem_w.hi, stem_w.lo, c1raw[0], c1raw[1], sp))
      return fail("stem plan: %s", tmap_last_error());
    plans.push_back(pl);
  }
  __nv_bfloat16** x = bufX;
  __nv_bfloat16** y = bufY;
  size_t ci = 0;
  int res_in = 56;
  auto mk = [&](const ConvLayer& L, int res, __nv_bfloat16** in, __nv_bfloat16** outb, __nv_bfloat16** res_b,
                int relu) -> int {
    Plan pl;
    ConvDesc d{n, res, res, L.cin, L.cout, L.ksize, L.stride};
    ConvIO io{};
    io.in_hi = in[0]; io.in_lo = in[1];
    io.w_hi = L.w.hi; io.w_lo = L.w.lo;
    io.bias = L.bias;
    if (res_b != nullptr) { io.res_hi = res_b[0]; io.res_lo = res_b[1]; }
    io.out_hi = outb[0]; io.out_lo = outb[1];
    io.relu = relu;
    if (build_conv_params(&pl.p, d, io, sp, &pl.block_n)) return fail("plan %s: %s", L.name.c_str(), tmap_last_error());
    plans.push_back(pl);
    return 0;
  };
  for (int li = 0; li < 4; ++li) {
    for (int bi = 0; bi < arch.blocks[li]; ++bi) {
      const int main_convs = arch.bottleneck ? 3 : 2;
      const ConvLayer& strided = convs[ci + (arch.bottleneck ? 1 : 0)];
      const int res_out = res_in / strided.stride;
      const bool dual = arch.bottleneck && convs[ci + 2].dual_cb > 0;
      const bool has_ds = !dual && ci + main_convs < convs.size() &&
                          convs[ci + main_convs].name.find("downsample") != std::string::npos;
      __nv_bfloat16** identity = x;
      if (dual) {
        if (mk(convs[ci], res_in, x, bufT1, nullptr, 1)) return 1;
        if (mk(convs[ci + 1], res_in, bufT1, bufT2, nullptr, 1)) return 1;
        const ConvLayer& L = convs[ci + 2];
        Plan pl;
        DualConvIO io{};
        io.a_hi = bufT2[0]; io.a_lo = bufT2[1];
        io.b_hi = x[0]; io.b_lo = x[1];
        io.w_hi = L.w.hi; io.w_lo = L.w.lo;
        io.bias = L.bias;
        io.out_hi = y[0]; io.out_lo = y[1];
        io.relu = 1;
        if (build_dual_1x1_params(&pl.p, n, res_out, res_out, L.cin, L.dual_cb, L.cout, L.stride, io, sp, &pl.block_n))
          return fail("plan %s: %s", L.name.c_str(), tmap_last_error());
        plans.push_back(pl);
      } else if (arch.bottleneck) {
        if (mk(convs[ci], res_in, x, bufT1, nullptr, 1)) return 1;
        if (mk(convs[ci + 1], res_in, bufT1, bufT2, nullptr, 1)) return 1;
        if (has_ds) {
          if (mk(convs[ci + 3], res_in, x, bufDS, nullptr, 0)) return 1;
          identity = bufDS;
        }
        if (mk(convs[ci + 2], res_out, bufT2, y, identity, 1)) return 1;
      } else {
        // BasicBlock: relu(bn2(conv2(relu(bn1(conv1(x))))) + identity)
        if (mk(convs[ci], res_in, x, bufT1, nullptr, 1)) return 1;
        if (has_ds) {
          if (mk(convs[ci + 2], res_in, x, bufDS, nullptr, 0)) return 1;
          identity = bufDS;
        }
        if (mk(convs[ci + 1], res_out, bufT1, y, identity, 1)) return 1;
      }
      std::swap(x, y);
      ci += main_convs + (has_ds ? 1 : 0);
      res_in = res_out;
    }
  }
  auto ins = enc_plans.emplace(n, std::move(plans));
  *out = &ins.first->second;
  return 0;
}

int MilanEngine::encode(const void* d_images, const void* d_masks, int n, int dtype, float* d_out, cudaStream_t st) {
  if (!cfg.has_encoder) return fail("engine was created without an encoder");
  if (n <= 0) return 0;
  if (n > cfg.max_images) return fail("milan_encode: n_images %d exceeds max_images %d", n, cfg.max_images);
  if (reinterpret_cast<uintptr_t>(d_images) % 16 != 0) return fail("milan_encode: d_images must be 16-byte aligned");
  if (alexnet) return encode_alexnet(d_images, d_masks, n, dtype, d_out, st);
  std::vector<Plan>* plans = nullptr;
  if (build_encoder_plans(n, &plans)) return 1;
  const int F = cfg.feature_size;
  const int sp = split ? 1 : 0;
  if (profiling && !profiling_append) conv_events_used = 0;
  if (reinterpret_cast<uintptr_t>(d_masks) % 16 != 0 && spatial)
    return fail("milan_encode: d_masks must be 16-byte aligned for a spatial encoder");
  RC(launch_stem_pack(d_images, dtype, n, stemA[0], stemA[1], mean, stdv, sp, st, spatial ? d_masks : nullptr));
  if (!spatial) {
    if (d_masks == nullptr) {
      if (ones_masks == nullptr) {
        uint8_t* p = nullptr;
        if (dalloc(&p, static_cast<size_t>(cfg.max_images) * 224 * 224)) return 1;
        CU(cudaMemset(p, 1, static_cast<size_t>(cfg.max_images) * 224 * 224));
        ones_masks = p;
      }
      RC(launch_mask_pyramid(ones_masks, MILAN_DTYPE_U8, n, mask_wts, st));
    } else {
      RC(launch_mask_pyramid(d_masks, dtype, n, mask_wts, st));
    }
  }
  size_t pi = 0;
  if (run_conv((*plans)[pi++], st)) return 1;  // stem
  if (!spatial)
    RC(launch_masked_pool(c1raw[0], c1raw[1], mask_wts + kMaskLevelOffset[0], kMaskPyramidSize, n, 12544, 64, d_out, F, st));
  RC(launch_bn_relu_maxpool(c1raw[0], c1raw[1], bn1_alpha, bn1_beta, n, bufX[0], bufX[1], st));
  __nv_bfloat16** x = bufX;
  __nv_bfloat16** y = bufY;
  int feat_off = 64;
  for (int li = 0; li < 4; ++li) {
    for (int bi = 0; bi < arch.blocks[li]; ++bi) {
      const bool dual = arch.bottleneck && bi == 0 && fuse_downsample;
      const bool has_ds = !dual && (bi == 0) && (li > 0 || expansion != 1);
      const int nconv = (arch.bottleneck ? 3 : 2) + (has_ds ? 1 : 0);
      for (int j = 0; j < nconv; ++j)
        if (run_conv((*plans)[pi++], st)) return 1;
      std::swap(x, y);
    }
    const int C = kPlanes[li] * expansion, P = kRes[li] * kRes[li];
    if (!spatial) {
      RC(launch_masked_pool(x[0], x[1], mask_wts + kMaskLevelOffset[li + 1], kMaskPyramidSize, n, P, C, d_out + feat_off, F, st));
      feat_off += C;
    } else if (li == 3) {
      RC(launch_planes_to_f32(x[0], x[1], static_cast<long long>(n) * P * C, d_out, st));
    }
  }
  if (pi != plans->size()) return fail("internal: encoder plan count mismatch (%zu of %zu)", pi, plans->size());
  return 0;
}

// ============================================================================ decoder
enum GemmId { G_KH = 0, G_INIT, G_QG, G_LSTM, G_OUT, G_LM0, G_LM1, G_LMOUT };

int MilanEngine::gemm(int which, long long M, const SplitMat& W, const float* bias, const __nv_bfloat16* a_hi,
                      const __nv_bfloat16* a_lo, long long a_pitch, int K, float* out, long long ldc,
                      cudaStream_t st, const int* skip) {
  if (M <= 0) return 0;
  auto key = std::make_pair(which, M);
  auto it = gemm_plans.find(key);
  if (it == gemm_plans.end()) {
    Plan pl;
    pl.block_n = 128;
    pl.epilogue = EPI_F32;
    if (build_gemm_params(&pl.p, M, K, W.rows, a_hi, a_lo, a_pitch, W.hi, W.lo, bias, out, ldc, split ? 1 : 0, 1))
      return fail("gemm plan %d (M=%lld K=%d N=%d): %s", which, M, K, W.rows, tmap_last_error());
    it = gemm_plans.emplace(key, pl).first;
  }
  RC(launch_conv_gemm(it->second.p, 128, split ? 1 : 0, EPI_F32, num_sms, st, skip));
  return 0;
}

// features (Bf, n_keys, F) fp32 -> hi/lo planes and kh = W_k f + b_k (step-invariant, hoisted out of the loop).
int MilanEngine::prepare_features(const float* d_features, int Bf, int n_keys, cudaStream_t st) {
  const int F = cfg.feature_size, A = cfg.attention_size;
  const int rows = Bf * n_keys;
  RC(launch_split_rows(d_features, F, feat[0], feat[1], F, rows, F, st));
  if (gemm(G_KH, rows, Wk, bk, feat[0], feat[1], F, F, kh, A, st)) return 1;
  return 0;
}

// One Decoder.step on R rows whose recurrent state lives in the workspace (h as hi/lo in Alstm[:, E+F:], c in
// `c`). Leaves logits in `logits`, the new state in hnew / hnew_f32 / cnew.
int MilanEngine::step_core(int R, int rpf, int n_keys, const float* d_features, const long long* d_tokens,
                           float* attn_out, long long attn_pitch, cudaStream_t st, const int* skip) {
  const int V = cfg.vocab_size, E = cfg.embedding_size, H = cfg.hidden_size, A = cfg.attention_size,
            F = cfg.feature_size;
  const long long xp = E + F + H;
  // q and gate pre-activations: [R][A+F] = h W1^T + b1
  if (gemm(G_QG, R, W1, b1, Alstm[0] + E + F, split ? Alstm[1] + E + F : nullptr, xp, H, qg, A + F, st, skip)) return 1;
  AttendArgs aa{};
  aa.qg = qg; aa.qg_pitch = A + F;
  aa.kh = kh; aa.features = d_features; aa.w_o = w_o; aa.b_o = b_o;
  aa.embedding = emb; aa.tokens = d_tokens;
  aa.R = R; aa.rows_per_feature = rpf; aa.n_keys = n_keys; aa.A = A; aa.F = F; aa.E = E;
  aa.x_hi = Alstm[0]; aa.x_lo = Alstm[1]; aa.x_pitch = xp;
  aa.attn_out = attn_out; aa.attn_pitch = attn_pitch;
  aa.skip = skip;
  aa.attn_ws = attn_ws; aa.acc_ws = n_keys > 16 ? attn_acc : nullptr;
  RC(launch_attend(aa, st));
  if (gemm(G_LSTM, R, W2, b2, Alstm[0], Alstm[1], xp, static_cast<int>(xp), gates, 4 * H, st, skip)) return 1;
  LstmPointArgs la{};
  la.gates = gates; la.c_in = c; la.c_out = cnew; la.h_out = hnew_f32;
  la.h_hi[0] = hnew[0]; la.h_lo[0] = hnew[1]; la.h_pitch[0] = H;
  la.R = R; la.H = H; la.skip = skip;
  RC(launch_lstm_point(la, st));
  if (gemm(G_OUT, R, W3, b3, hnew[0], hnew[1], H, H, logits, ldv, st, skip)) return 1;
  (void)V;
  return 0;
}

int MilanEngine::lm_reset(int M, cudaStream_t st) {
  const int El = cfg.lm_embedding_size, Hl = cfg.lm_hidden_size;
  for (int pl = 0; pl < (split ? 2 : 1); ++pl) {
    CU(cudaMemsetAsync(Alm0[pl], 0, static_cast<size_t>(M) * (El + Hl) * 2, st));
    CU(cudaMemsetAsync(Alm1[pl], 0, static_cast<size_t>(M) * 2 * Hl * 2, st));
  }
  CU(cudaMemsetAsync(lm_c0, 0, static_cast<size_t>(M) * Hl * 4, st));
  CU(cudaMemsetAsync(lm_c1, 0, static_cast<size_t>(M) * Hl * 4, st));
  return 0;
}

// One LM step (embedding -> 2 LSTM layers -> vocab logits in logits_lm) on M rows; state in Alm0[:, El:],
// Alm1[:, Hl:], lm_c0, lm_c1. to_new: write the new state to lmnew0/1 + lm_c0n/1n instead (beam reorder follows).
int MilanEngine::lm_step_core(int M, const long long* d_tokens, bool to_new, cudaStream_t st, const int* skip) {
  const int El = cfg.lm_embedding_size, Hl = cfg.lm_hidden_size;
  const long long p0 = El + Hl, p1 = 2 * Hl;
  RC(launch_embed_rows(lm_emb, d_tokens, M, El, Alm0[0], Alm0[1], p0, st, skip));
  if (gemm(G_LM0, M, L0, bl0, Alm0[0], Alm0[1], p0, static_cast<int>(p0), gates, 4 * Hl, st, skip)) return 1;
  LstmPointArgs a0{};
  a0.gates = gates; a0.c_in = lm_c0; a0.c_out = to_new ? lm_c0n : lm_c0;
  a0.h_out = lm_h_f32;
  a0.h_hi[0] = Alm1[0]; a0.h_lo[0] = Alm1[1]; a0.h_pitch[0] = p1;  // input of layer 1, this step
  if (to_new) { a0.h_hi[1] = lmnew0[0]; a0.h_lo[1] = lmnew0[1]; a0.h_pitch[1] = Hl; }
  else        { a0.h_hi[1] = Alm0[0] + El; a0.h_lo[1] = split ? Alm0[1] + El : nullptr; a0.h_pitch[1] = p0; }
  a0.R = M; a0.H = Hl; a0.skip = skip;
  RC(launch_lstm_point(a0, st));
  if (gemm(G_LM1, M, L1, bl1, Alm1[0], Alm1[1], p1, static_cast<int>(p1), gates, 4 * Hl, st, skip)) return 1;
  LstmPointArgs a1{};
  a1.gates = gates; a1.c_in = lm_c1; a1.c_out = to_new ? lm_c1n : lm_c1;
  a1.h_out = lm_h_f32 + static_cast<size_t>(M) * Hl;
  a1.h_hi[0] = lmh[0]; a1.h_lo[0] = lmh[1]; a1.h_pitch[0] = Hl;  // input of the output projection
  if (to_new) { a1.h_hi[1] = lmnew1[0]; a1.h_lo[1] = lmnew1[1]; a1.h_pitch[1] = Hl; }
  else        { a1.h_hi[1] = Alm1[0] + Hl; a1.h_lo[1] = split ? Alm1[1] + Hl : nullptr; a1.h_pitch[1] = p1; }
  a1.R = M; a1.H = Hl; a1.skip = skip;
  RC(launch_lstm_point(a1, st));
  if (gemm(G_LMOUT, M, Lout, blout, lmh[0], lmh[1], Hl, Hl, logits_lm, ldv, st, skip)) return 1;
  return 0;
}

namespace {
// pooled mean -> [W_h; W_c] GEMM -> tanh; h as hi/lo into Alstm[:, E+F:], c into e->c.
int init_state_impl(MilanEngine* e, const float* d_features, int B, int n_keys, float* d_h, float* d_c,
                    cudaStream_t st) {
  const int F = e->cfg.feature_size, H = e->cfg.hidden_size, E = e->cfg.embedding_size;
  RC(launch_mean_keys(d_features, B, n_keys, F, e->pooled[0], e->pooled[1], st));
  if (e->gemm(G_INIT, B, e->Winit, e->binit, e->pooled[0], e->pooled[1], F, F, e->init_pre, 2 * H, st)) return 1;
  RC(launch_init_finish(e->init_pre, B, H, d_h, d_c != nullptr ? d_c : e->c, e->Alstm[0] + E + F,
                        e->split ? e->Alstm[1] + E + F : nullptr, E + F + H, st));
  return 0;
}
}  // namespace

int MilanEngine::decode_greedy(const float* d_features, int B, int n_keys, int length, int mi, float temperature,
                               const long long* d_forced, long long* d_tokens_out, float* d_scores_out,
                               float* d_pred_out, float* d_attn_out, cudaStream_t st) {
  const int V = cfg.vocab_size, H = cfg.hidden_size, E = cfg.embedding_size, F = cfg.feature_size;
  if (B > Rmax) return fail("decode_greedy: B=%d exceeds capacity %d", B, Rmax);
  if (length < 1 || length > cfg.max_length) return fail("length %d outside [1, max_length = %d]", length, cfg.max_length);
  if (static_cast<size_t>(B) * n_keys > FRcap) return fail("decode_greedy: too many feature rows");
  if (mi && !cfg.has_lm) return fail("cannot use MI decoding without an LM");
  if (prepare_features(d_features, B, n_keys, st)) return 1;
  if (init_state_impl(this, d_features, B, n_keys, nullptr, nullptr, st)) return 1;
  RC(launch_fill_i64(tok_cur, cfg.start_index, B, st));
  RC(launch_fill_f32(greedy_scores, 0.f, B, st));
  if (mi && lm_reset(B, st)) return 1;
  for (int t = 0; t < length; ++t) {
    float* attn = d_attn_out != nullptr ? d_attn_out + static_cast<size_t>(t) * n_keys : nullptr;
    if (step_core(B, 1, n_keys, d_features, tok_cur, attn, static_cast<long long>(length) * n_keys, st)) return 1;
    if (mi && lm_step_core(B, tok_cur, false, st)) return 1;
    RowArgs ra{};
    ra.logits = logits; ra.logits_lm = mi ? logits_lm : nullptr; ra.ld = ldv; ra.R = B; ra.V = V;
    ra.temperature = temperature;
    ra.next_tokens = tok_next; ra.scores = greedy_scores;
    ra.pred_out = d_pred_out != nullptr ? d_pred_out + static_cast<size_t>(t) * V : nullptr;
    ra.pred_pitch = static_cast<long long>(length) * V;
    ra.beam = 0;
    if (d_forced != nullptr) {
      // forced tokens for step t are column t of (B, length): gather into tok_next via strided copy
      CU(cudaMemcpy2DAsync(tok_next, sizeof(long long), d_forced + t, sizeof(long long) * length, sizeof(long long), B,
                           cudaMemcpyDeviceToDevice, st));
      ra.forced = tok_next;
    }
    RC(launch_row_logsoftmax(ra, st));
    CU(cudaMemcpy2DAsync(d_tokens_out + t, sizeof(long long) * length, tok_next, sizeof(long long), sizeof(long long), B,
                         cudaMemcpyDeviceToDevice, st));
    // new state -> current
    GatherArgs ga{};
    ga.backptr = nullptr; ga.R = B; ga.H = H;
    ga.src_hi = hnew[0]; ga.src_lo = hnew[1]; ga.src_pitch = H;
    ga.dst_hi = Alstm[0] + E + F; ga.dst_lo = split ? Alstm[1] + E + F : nullptr; ga.dst_pitch = E + F + H;
    ga.c_src = cnew; ga.c_dst = c;
    RC(launch_gather_state(ga, st));
    std::swap(tok_cur, tok_next);
  }
  CU(cudaMemcpyAsync(d_scores_out, greedy_scores, sizeof(float) * B, cudaMemcpyDeviceToDevice, st));
  return 0;
}

int MilanEngine::lm_score_seqs(const long long* d_seqs, int M, int length, int beam, int group_size, cudaStream_t st) {
  if (fused_decode && fused_lm_ready) return lm_score_seqs_fused(d_seqs, M, length, beam, group_size, st);
  const int V = cfg.vocab_size;
  if (lm_reset(M, st)) return 1;
  RC(launch_fill_f32(lm_scores, 0.f, M, st));
  const int groups = beam > 0 && group_size < (1 << 29) ? ((M / beam) + group_size - 1) / group_size : 1;
  RC(launch_lm_skip(group_T, groups, length, lm_skip, st));  // positions beyond every group's T are no-ops
  for (int t = 0; t < length; ++t) {
    const int* skip = lm_skip + t;
    RC(launch_lm_inputs(d_seqs, M, length, t, cfg.start_index, lm_inputs, st, skip));
    if (lm_step_core(M, lm_inputs, false, st, skip)) return 1;
    LmAccumArgs la{};
    la.logits = logits_lm; la.ld = ldv; la.M = M; la.V = V; la.length = length; la.t = t; la.beam = beam;
    la.group_size = group_size; la.seqs = d_seqs; la.group_T = group_T; la.stop_index = cfg.stop_index;
    la.lm_scores = lm_scores; la.skip = skip;
    RC(launch_lm_accumulate(la, st));
  }
  return 0;
}

int MilanEngine::decode_beam(const float* d_features, int B, int n_keys, int length, int beam, int group_size,
                             int rerank, int mi, float temperature, long long* d_beam_tokens, float* d_beam_scores,
                             int* d_group_steps, long long* d_tokens_out, float* d_scores_out, float* d_lm_scores_out,
                             cudaStream_t st) {
  const int V = cfg.vocab_size, H = cfg.hidden_size, E = cfg.embedding_size, F = cfg.feature_size;
  if (beam < 1 || beam > cfg.max_beam || beam > kMaxBeam) return fail("beam size %d unsupported (max %d)", beam, cfg.max_beam);
  if (beam > V) return fail("Target vocab size (%d) too small relative to per_node_beam_size (%d).", V, beam);
  if (B > Bmax) return fail("decode_beam: B=%d exceeds max_neurons %d", B, Bmax);
  if (static_cast<size_t>(B) * n_keys > FRcap) return fail("decode_beam: too many feature rows");
  if (length < 1 || length > cfg.max_length) return fail("length %d outside [1, max_length = %d]", length, cfg.max_length);
  if ((rerank || mi) && !cfg.has_lm) return fail("cannot use MI/rerank decoding without an LM");
  if (rerank && mi) return fail("cannot set `mi=` decoding when reranking");
  if (group_size <= 0) group_size = B;
  const int R = B * beam;
  if (prepare_features(d_features, B, n_keys, st)) return 1;
  if (init_state_impl(this, d_features, B, n_keys, nullptr, nullptr, st)) return 1;
  RC(launch_fill_i64(tok_cur, cfg.start_index, B, st));
  CU(cudaMemsetAsync(d_done, 0, sizeof(int), st));
  if (mi && lm_reset(R, st)) return 1;
  const int Hl = cfg.lm_hidden_size, El = cfg.lm_embedding_size;
  const bool fused = use_fused_beam(n_keys, mi, beam);
  if (fused && beam_steps_fused(d_features, B, n_keys, length, beam, st)) return 1;
  for (int t = 0; t < length && !fused; ++t) {
    const int rows = t == 0 ? B : R;
    const int rpf = t == 0 ? 1 : beam;
    if (step_core(rows, rpf, n_keys, d_features, tok_cur, nullptr, 0, st, d_done)) return 1;
    if (mi && lm_step_core(rows, tok_cur, true, st, d_done)) return 1;
    RowArgs ra{};
    ra.logits = logits; ra.logits_lm = mi ? logits_lm : nullptr; ra.ld = ldv; ra.R = rows; ra.V = V;
    ra.temperature = temperature;
    ra.beam = beam; ra.last_tokens = tok_cur; ra.last_lp = t == 0 ? nullptr : last_lp;
    ra.stop_index = cfg.stop_index; ra.cand_val = cand_val; ra.cand_cls = cand_cls; ra.skip = d_done;
    RC(launch_row_logsoftmax(ra, st));
    MergeArgs ma{};
    ma.cand_val = cand_val; ma.cand_cls = cand_cls; ma.n_neurons = B; ma.in_rows = rpf; ma.beam = beam;
    ma.next_tokens = tok_next; ma.next_lp = next_lp; ma.backptr = backptr;
    ma.hist_tok = hist_tok + static_cast<size_t>(t) * R; ma.hist_bp = hist_bp + static_cast<size_t>(t) * R;
    ma.skip = d_done; ma.cur_lp = last_lp; ma.stop_index = cfg.stop_index;
    RC(launch_beam_merge(ma, st));
    GatherArgs ga{};
    ga.backptr = backptr; ga.R = R; ga.H = H;
    ga.src_hi = hnew[0]; ga.src_lo = hnew[1]; ga.src_pitch = H;
    ga.dst_hi = Alstm[0] + E + F; ga.dst_lo = split ? Alstm[1] + E + F : nullptr; ga.dst_pitch = E + F + H;
    ga.c_src = cnew; ga.c_dst = c; ga.skip = d_done;
    RC(launch_gather_state(ga, st));
    if (mi) {  // the LM state follows the same backpointers (AllenNLPDecoderState h_lm / c_lm)
      GatherArgs g0{};
      g0.backptr = backptr; g0.R = R; g0.H = Hl;
      g0.src_hi = lmnew0[0]; g0.src_lo = lmnew0[1]; g0.src_pitch = Hl;
      g0.dst_hi = Alm0[0] + El; g0.dst_lo = split ? Alm0[1] + El : nullptr; g0.dst_pitch = El + Hl;
      g0.c_src = lm_c0n; g0.c_dst = lm_c0; g0.skip = d_done;
      RC(launch_gather_state(g0, st));
      GatherArgs g1{};
      g1.backptr = backptr; g1.R = R; g1.H = Hl;
      g1.src_hi = lmnew1[0]; g1.src_lo = lmnew1[1]; g1.src_pitch = Hl;
      g1.dst_hi = Alm1[0] + Hl; g1.dst_lo = split ? Alm1[1] + Hl : nullptr; g1.dst_pitch = 2 * Hl;
      g1.c_src = lm_c1n; g1.c_dst = lm_c1; g1.skip = d_done;
      RC(launch_gather_state(g1, st));
    }
    // allennlp: `if (last_predictions == end).all(): break` — evaluated on the device, later steps become no-ops
    RC(launch_check_done(tok_next, R, cfg.stop_index, d_done, st));
    std::swap(tok_cur, tok_next);
    std::swap(last_lp, next_lp);
  }
  BacktrackArgs ba{};
  ba.hist_tok = hist_tok; ba.hist_bp = hist_bp; ba.n_neurons = B; ba.beam = beam; ba.length = length;
  ba.group_size = group_size; ba.stop_index = cfg.stop_index; ba.seqs = seqs; ba.group_T = group_T;
  RC(launch_backtrack(ba, st));
  if (d_beam_tokens != nullptr)
    CU(cudaMemcpyAsync(d_beam_tokens, seqs, sizeof(long long) * R * length, cudaMemcpyDeviceToDevice, st));
  if (d_beam_scores != nullptr)
    CU(cudaMemcpyAsync(d_beam_scores, last_lp, sizeof(float) * R, cudaMemcpyDeviceToDevice, st));
  const int groups = (B + group_size - 1) / group_size;
  if (d_group_steps != nullptr)
    CU(cudaMemcpyAsync(d_group_steps, group_T, sizeof(int) * groups, cudaMemcpyDeviceToDevice, st));
  if (d_tokens_out != nullptr || d_scores_out != nullptr) {
    if (rerank) {
      if (lm_score_seqs(seqs, R, length, beam, group_size, st)) return 1;
      if (d_lm_scores_out != nullptr)
        CU(cudaMemcpyAsync(d_lm_scores_out, lm_scores, sizeof(float) * R, cudaMemcpyDeviceToDevice, st));
    }
    // rerank: argmax_j beam_lp[j] - temperature * lm[j] (decoders.py:495-512); 'beam': the sorted beam's first
    // entry (decoders.py:491-493), which is the same argmax with the LM term dropped.
    RerankArgs rr{};
    rr.beam_lp = last_lp; rr.lm_scores = rerank ? lm_scores : nullptr; rr.temperature = rerank ? temperature : 0.f;
    rr.n_neurons = B; rr.beam = beam; rr.length = length; rr.seqs = seqs;
    rr.out_tokens = d_tokens_out != nullptr ? d_tokens_out : out_tokens;
    rr.out_scores = d_scores_out != nullptr ? d_scores_out : out_scores; rr.out_index = nullptr;
    RC(launch_rerank_select(rr, st));
  }
  return 0;
}


// ============================================================================ fused beam step / LM rerank
enum FusedGemmId { F_HEAD0 = 100, F_LSTM, F_HEAD, F_LM0_A, F_LM0_B, F_LM1_A, F_LM1_B, F_LMOUT_A, F_LMOUT_B };

bool MilanEngine::use_fused_beam(int n_keys, int mi, int beam) const {
  return fused_decode && fused_ready && !mi && n_keys <= kFusedMaxKeys && beam <= kMaxBeam;
}

// Cached tensor maps / tiling of a flat GEMM whose epilogue is patched per launch (A = [a | a1] when K1 > 0).
int MilanEngine::fused_gemm(int which, long long M, const ConvGemmParams** out, int K, const SplitMat& W,
                            const float* bias, const __nv_bfloat16* a_hi, const __nv_bfloat16* a_lo, long long a_pitch,
                            const __nv_bfloat16* a1_hi, const __nv_bfloat16* a1_lo, long long a1_pitch, int K1) {
  auto key = std::make_pair(which, M);
  auto it = gemm_plans.find(key);
  if (it == gemm_plans.end()) {
    Plan pl;
    pl.block_n = 128;
    int rc;
    if (K1 > 0)
      rc = build_gemm2_params(&pl.p, M, K, K1, W.rows, a_hi, a_lo, a_pitch, a1_hi, a1_lo, a1_pitch, W.hi, W.lo, bias, 1, 1);
    else
      rc = build_gemm_params(&pl.p, M, K, W.rows, a_hi, a_lo, a_pitch, W.hi, W.lo, bias, nullptr, W.rows, 1, 1);
    if (rc) return fail("fused gemm plan %d (M=%lld K=%d+%d N=%d): %s", which, M, K, K1, W.rows, tmap_last_error());
    it = gemm_plans.emplace(key, pl).first;
  }
  *out = &it->second.p;
  return 0;
}

int MilanEngine::run_fused(const ConvGemmParams& p, int epilogue, cudaStream_t st, const int* skip) {
  RC(launch_conv_gemm(p, 128, 1, epilogue, num_sms, st, skip));
  return 0;
}

// The beam loop of decode_beam as three launches per step (decode_fused.h): LSTM GEMM, head GEMM, select + attend. On return tok_cur / last_lp hold the
// final beam, hist_tok / hist_bp the history; state lives in Alstm / hnew / c / cnew like the unfused path.
int MilanEngine::beam_steps_fused(const float* d_features, int B, int n_keys, int length, int beam, cudaStream_t st) {
  const int V = cfg.vocab_size, H = cfg.hidden_size, E = cfg.embedding_size, F = cfg.feature_size,
            A = cfg.attention_size;
  const int R = B * beam;
  const long long xp = E + F + H, qgp = A + F;
  const __nv_bfloat16* h0_hi = Alstm[0] + E + F;
  const __nv_bfloat16* h0_lo = Alstm[1] + E + F;
  {  // attention query and feature gate of the initial state (the head GEMM's q / gate sections only)
    const ConvGemmParams* base = nullptr;
    if (fused_gemm(F_HEAD0, B, &base, H, W1, b1, h0_hi, h0_lo, xp)) return 1;
    ConvGemmParams p = *base;
    p.fe.vocab_tiles = 0; p.fe.q_tiles = q_tiles;
    p.fe.q_out = qg; p.fe.q_pitch = qgp; p.fe.q_cols = A;
    p.fe.g_out = qg + A; p.fe.g_pitch = qgp; p.fe.gate_cols = F;
    if (run_fused(p, EPI_HEAD, st, nullptr)) return 1;
  }
  float* c_cur = c;
  float* c_nxt = cnew;
  // attention + gating + operand assembly of step t, from the tokens / backpointers / parent state as they are then
  auto attend_args = [&](int t, const long long* tokens) {
    AttendFusedArgs aa{};
    aa.q = qg; aa.q_pitch = qgp; aa.gate = qg + A; aa.gate_pitch = qgp;
    aa.src_row = t == 0 ? nullptr : backptr;
    aa.kh = kh; aa.features = d_features; aa.w_o = w_o; aa.b_o = b_o;
    aa.embedding = emb; aa.tokens = tokens;
    if (t > 0) { aa.h_src_hi = hnew[0]; aa.h_src_lo = hnew[1]; aa.h_src_pitch = H; }
    aa.R = t == 0 ? B : R; aa.rows_per_feature = t == 0 ? 1 : beam;
    aa.n_keys = n_keys; aa.A = A; aa.F = F; aa.E = E; aa.H = H;
    aa.x_hi = Alstm[0]; aa.x_lo = Alstm[1]; aa.x_pitch = xp;
    aa.attn_ws = attn_ws; aa.skip = d_done;
    return aa;
  };
  RC(launch_attend_fused(attend_args(0, tok_cur), st));
  for (int t = 0; t < length; ++t) {
    const int rows = t == 0 ? B : R;
    const int rpf = t == 0 ? 1 : beam;
    const int* parents = t == 0 ? nullptr : backptr;
    {
      const ConvGemmParams* base = nullptr;
      if (fused_gemm(F_LSTM, rows, &base, static_cast<int>(xp), W2p, b2p, Alstm[0], Alstm[1], xp)) return 1;
      ConvGemmParams p = *base;
      p.fe.hidden = H; p.fe.c_in = c_cur; p.fe.c_out = c_nxt; p.fe.src_row = parents;
      p.fe.h_hi = hnew[0]; p.fe.h_lo = hnew[1]; p.fe.h_pitch = H;
      if (run_fused(p, EPI_LSTM, st, d_done)) return 1;
    }
    {
      const ConvGemmParams* base = nullptr;
      if (fused_gemm(F_HEAD, rows, &base, H, Whead, bhead, hnew[0], hnew[1], H)) return 1;
      ConvGemmParams p = *base;
      p.fe.vocab_tiles = vocab_tiles; p.fe.q_tiles = q_tiles; p.fe.vocab = V;
      p.fe.logits = logits; p.fe.ld_logits = ldv; p.fe.partials = partials;
      p.fe.q_out = qg; p.fe.q_pitch = qgp; p.fe.q_cols = A;
      p.fe.g_out = qg + A; p.fe.g_pitch = qgp; p.fe.gate_cols = F;
      if (t == length - 1) {  // nothing reads the query / gate of the last step
        p.n_tiles = vocab_tiles;
        p.cout = std::min(p.cout, vocab_tiles * 128);
      }
      if (run_fused(p, EPI_HEAD, st, d_done)) return 1;
    }
    BeamSelectArgs sa{};
    sa.logits = logits; sa.ld = ldv; sa.partials = partials; sa.n_seg = n_seg; sa.V = V;
    sa.last_tokens = tok_cur; sa.last_lp = t == 0 ? nullptr : last_lp;
    sa.n_neurons = B; sa.in_rows = rpf; sa.beam = beam; sa.stop_index = cfg.stop_index;
    sa.cand_val = cand_val; sa.cand_cls = cand_cls;
    sa.next_tokens = tok_next; sa.next_lp = next_lp; sa.backptr = backptr;
    sa.hist_tok = hist_tok + static_cast<size_t>(t) * R; sa.hist_bp = hist_bp + static_cast<size_t>(t) * R;
    sa.cur_lp = last_lp; sa.done_flag = d_done; sa.counters = beam_counters;
    if (t + 1 < length) {
      // this step's selection + the next step's attention (which reads the beam the selection has just written)
      RC(launch_select_attend(sa, attend_args(t + 1, tok_next), st));
    } else {
      RC(launch_beam_select(sa, st));
    }
    std::swap(tok_cur, tok_next);
    std::swap(last_lp, next_lp);
    std::swap(c_cur, c_nxt);
  }
  return 0;
}

// lm_score_seqs with three launches per position: both LSTM cells finish in their GEMM's epilogue (layer 0 reads its
// input through lm_table), the vocabulary GEMM leaves softmax partials + the target logit, one read-out at the end.
int MilanEngine::lm_score_seqs_fused(const long long* d_seqs, int M, int length, int beam, int group_size,
                                     cudaStream_t st) {
  const int V = cfg.vocab_size, Hl = cfg.lm_hidden_size;
  const size_t plane = static_cast<size_t>(M) * Hl * sizeof(__nv_bfloat16);
  for (int hl = 0; hl < 2; ++hl) {
    CU(cudaMemsetAsync(lm_h0[0][hl], 0, plane, st));
    CU(cudaMemsetAsync(lm_h1[0][hl], 0, plane, st));
  }
  CU(cudaMemsetAsync(lm_c0, 0, static_cast<size_t>(M) * Hl * 4, st));
  CU(cudaMemsetAsync(lm_c1, 0, static_cast<size_t>(M) * Hl * 4, st));
  const int groups = beam > 0 && group_size < (1 << 29) ? ((M / beam) + group_size - 1) / group_size : 1;
  RC(launch_lm_skip(group_T, groups, length, lm_skip, st));  // positions beyond every group's T are no-ops
  for (int t = 0; t < length; ++t) {
    const int cur = t & 1, nxt = cur ^ 1;
    const int* skip = lm_skip + t;
    {
      const ConvGemmParams* base = nullptr;
      if (fused_gemm(cur ? F_LM0_B : F_LM0_A, M, &base, Hl, L0h, bl0p, lm_h0[cur][0], lm_h0[cur][1], Hl)) return 1;
      ConvGemmParams p = *base;
      p.fe.hidden = Hl; p.fe.c_in = lm_c0; p.fe.c_out = lm_c0;
      p.fe.h_hi = lm_h0[nxt][0]; p.fe.h_lo = lm_h0[nxt][1]; p.fe.h_pitch = Hl;
      p.fe.add_table = lm_table; p.fe.add_pitch = 4LL * Hl;
      p.fe.add_index = t == 0 ? nullptr : d_seqs + (t - 1);  // inputs = [<start>, seq...]
      p.fe.add_stride = length; p.fe.add_const = cfg.start_index;
      if (run_fused(p, EPI_LSTM, st, skip)) return 1;
    }
    {
      const ConvGemmParams* base = nullptr;
      if (fused_gemm(cur ? F_LM1_B : F_LM1_A, M, &base, Hl, L1p, bl1p, lm_h0[nxt][0], lm_h0[nxt][1], Hl, lm_h1[cur][0],
                     lm_h1[cur][1], Hl, Hl))
        return 1;
      ConvGemmParams p = *base;
      p.fe.hidden = Hl; p.fe.c_in = lm_c1; p.fe.c_out = lm_c1;
      p.fe.h_hi = lm_h1[nxt][0]; p.fe.h_lo = lm_h1[nxt][1]; p.fe.h_pitch = Hl;
      if (run_fused(p, EPI_LSTM, st, skip)) return 1;
    }
    {
      const ConvGemmParams* base = nullptr;
      if (fused_gemm(cur ? F_LMOUT_B : F_LMOUT_A, M, &base, Hl, Lout, blout, lm_h1[nxt][0], lm_h1[nxt][1], Hl)) return 1;
      ConvGemmParams p = *base;
      p.fe.vocab_tiles = vocab_tiles; p.fe.q_tiles = 0; p.fe.vocab = V;
      p.fe.partials = lm_partials + static_cast<size_t>(t) * M * n_seg;
      p.fe.target = d_seqs + t; p.fe.target_stride = length;
      p.fe.tgt_logit = lm_tgt + static_cast<size_t>(t) * M;
      if (run_fused(p, EPI_HEAD, st, skip)) return 1;
    }
  }
  LmFinalizeArgs fa{};
  fa.partials = lm_partials; fa.tgt_logit = lm_tgt; fa.n_seg = n_seg; fa.M = M; fa.length = length; fa.beam = beam;
  fa.group_size = group_size; fa.seqs = d_seqs; fa.group_T = group_T; fa.stop_index = cfg.stop_index;
  fa.lm_scores = lm_scores;
  RC(launch_lm_finalize(fa, st));
  return 0;
}

// ============================================================================ C ABI
extern "C" {

const char* milan_version(void) { return "milan_b200 0.1 (sm_100a: tcgen05 + TMA)"; }
const char* milan_last_error(void) { return g_err; }

int milan_engine_create(const MilanConfig* config, int device, MilanEngine** out) {
  g_err[0] = 0;
  if (config == nullptr || out == nullptr) return fail("null argument");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail("no CUDA device available (%s): the milan_b200 engine has no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= count) return fail("device %d out of range (%d devices)", device, count);
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail("device %s is sm_%d%d; this engine is built for sm_100a only", prop.name, prop.major, prop.minor);
  if (config->max_beam > kMaxBeam) return fail("max_beam %d exceeds %d", config->max_beam, kMaxBeam);
  if (config->encoder_arch < 0 || config->encoder_arch >= kNumArchs)
    return fail("encoder not supported: encoder_arch %d", config->encoder_arch);
  if (config->encoder_kind != MILAN_ENCODER_PYRAMID && config->encoder_kind != MILAN_ENCODER_SPATIAL)
    return fail("encoder not supported: encoder_kind %d", config->encoder_kind);
  const EncoderArch& arch = kArchs[config->encoder_arch];
  const int expansion = arch.bottleneck ? 4 : 1;
  const bool spatial = config->encoder_kind == MILAN_ENCODER_SPATIAL;
  const bool alexnet = config->encoder_arch == MILAN_ENCODER_ALEXNET;
  if (alexnet && spatial) return fail("encoder not supported: spatial alexnet");
  if (config->has_encoder) {
    const int want = alexnet ? 64 + 192 + 384 + 256 + 256
                             : (spatial ? 512 * expansion : 64 + (64 + 128 + 256 + 512) * expansion);
    if (config->feature_size != want)
      return fail("feature_size %d does not match the %s %s encoder (%d)", config->feature_size, arch.name,
                  spatial ? "spatial" : "pyramid", want);
  }
  auto* eng = new MilanEngine();
  eng->cfg = *config;
  eng->device = device;
  eng->num_sms = prop.multiProcessorCount;
  if (const char* e = getenv("MILAN_NUM_SMS")) {  // experiment knob: CTAs of the persistent conv / GEMM kernels (two engines can share a GPU)
    const int n = atoi(e);
    if (n >= 2 && n <= eng->num_sms) eng->num_sms = n & ~1;
  }
  eng->split = config->precision == MILAN_PRECISION_SPLIT;
  eng->arch = arch;
  eng->expansion = expansion;
  eng->spatial = spatial;
  eng->alexnet = alexnet;
  if (const char* env = getenv("MILAN_FUSE_DOWNSAMPLE")) eng->fuse_downsample = atoi(env) != 0;
  if (const char* env = getenv("MILAN_OVERLAP_DECODE")) eng->overlap_decode = atoi(env) != 0;
  if (const char* env = getenv("MILAN_FUSED_DECODE")) eng->fused_decode = atoi(env) != 0;
  eng->enc_out_per_image = (spatial ? kSpatialKeys : 1) * config->feature_size;
  *out = eng;
  return 0;
}

void milan_engine_destroy(MilanEngine* engine) {
  if (engine == nullptr) return;
  DeviceGuard device_guard(engine->device);
  cudaDeviceSynchronize();
  for (void* p : engine->allocs) cudaFree(p);
  for (auto& ev : engine->conv_events) {
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  if (engine->copy_stream != nullptr) {
    cudaStreamDestroy(engine->copy_stream);
    cudaStreamDestroy(engine->dec_stream);
    for (int b = 0; b < 2; ++b) {
      cudaEventDestroy(engine->ev_copied[b]);
      cudaEventDestroy(engine->ev_consumed[b]);
      cudaEventDestroy(engine->ev_feat_ready[b]);
      cudaEventDestroy(engine->ev_feat_free[b]);
    }
    cudaEventDestroy(engine->ev_dec_done);
  }
  delete engine;
}

int milan_engine_set_tensor(MilanEngine* engine, const char* name, const float* h_data, const int64_t* shape,
                            int32_t ndim) {
  g_err[0] = 0;
  if (engine == nullptr || name == nullptr || h_data == nullptr) return fail("null argument");
  if (engine->finalized) return fail("engine already finalized");
  HostTensor t;
  t.shape.assign(shape, shape + ndim);
  t.data.assign(h_data, h_data + t.numel());
  engine->pending[name] = std::move(t);
  return 0;
}

int milan_engine_finalize(MilanEngine* engine) {
  g_err[0] = 0;
  if (engine == nullptr) return fail("null engine");
  if (engine->finalized) return 0;
  DeviceGuard device_guard(engine->device);
  CU(device_guard.err);
  if (engine->cfg.has_encoder && engine->finalize_encoder()) return 1;
  // An engine given encoder tensors only (standalone Encoder use, src/milan/encoders.py) has no decoder half.
  engine->has_decoder = !engine->cfg.has_encoder || engine->get("lstm.weight_ih") != nullptr;
  if (engine->has_decoder && engine->finalize_decoder()) return 1;
  if (engine->alloc_workspace()) return 1;
  engine->pending.clear();
  engine->finalized = true;
  CU(cudaDeviceSynchronize());
  return 0;
}

#define CHECK_READY(e)                                             \
  g_err[0] = 0;                                                    \
  if ((e) == nullptr) return fail("null engine");                  \
  if (!(e)->finalized) return fail("engine not finalized");        \
  DeviceGuard device_guard_((e)->device);                          \
  CU(device_guard_.err);
#define CHECK_DECODER(e) \
  if (!(e)->has_decoder) return fail("engine was created without decoder weights (encoder-only)");

int milan_encode(MilanEngine* engine, const void* d_images, const void* d_masks, int32_t n_images, int32_t dtype,
                 float* d_features_out, void* stream) {
  CHECK_READY(engine);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int done = 0;
  const size_t img_stride = static_cast<size_t>(3) * 224 * 224 * (dtype == MILAN_DTYPE_U8 ? 1 : 4);
  const size_t msk_stride = static_cast<size_t>(224) * 224 * (dtype == MILAN_DTYPE_U8 ? 1 : 4);
  while (done < n_images) {
    const int n = std::min(n_images - done, engine->cfg.max_images);
    const uint8_t* img = static_cast<const uint8_t*>(d_images) + done * img_stride;
    const uint8_t* msk = d_masks ? static_cast<const uint8_t*>(d_masks) + done * msk_stride : nullptr;
    if (engine->encode(img, msk, n, dtype, d_features_out + static_cast<size_t>(done) * engine->enc_out_per_image, st))
      return 1;
    if (engine->collect_conv_events(st)) return 1;
    done += n;
  }
  return 0;
}

int milan_init_state(MilanEngine* engine, const float* d_features, int32_t B, int32_t n_keys, float* d_h,
                     float* d_c, void* stream) {
  CHECK_READY(engine);
  CHECK_DECODER(engine);
  if (B > engine->Rmax || static_cast<size_t>(B) * n_keys > engine->FRcap) return fail("init_state: B=%d exceeds capacity", B);
  return init_state_impl(engine, d_features, B, n_keys, d_h, d_c, static_cast<cudaStream_t>(stream));
}

int milan_step(MilanEngine* engine, const float* d_features, int32_t n_keys, const int64_t* d_tokens, float* d_h,
               float* d_c, float* d_h_lm, float* d_c_lm, int32_t R, int32_t rows_per_feature, float temperature,
               float* d_predictions_out, float* d_attentions_out, void* stream) {
  CHECK_READY(engine);
  CHECK_DECODER(engine);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MilanEngine* e = engine;
  const int V = e->cfg.vocab_size, H = e->cfg.hidden_size, E = e->cfg.embedding_size, F = e->cfg.feature_size;
  if (R > e->Rmax) return fail("milan_step: R=%d exceeds capacity %d", R, e->Rmax);
  if (rows_per_feature <= 0 || R % rows_per_feature) return fail("R must be a multiple of rows_per_feature");
  const int Bf = R / rows_per_feature;
  if (static_cast<size_t>(Bf) * n_keys > e->FRcap)
    return fail("milan_step: %d feature sets x %d keys exceed the workspace", Bf, n_keys);
  if ((d_h_lm == nullptr) != (d_c_lm == nullptr)) return fail("state must have both h_lm and c_lm or neither");
  if (d_h_lm != nullptr && !e->cfg.has_lm) return fail("state has h_lm or c_lm, but decoder has no lm");
  if (e->prepare_features(d_features, Bf, n_keys, st)) return 1;
  RC(launch_split_rows(d_h, H, e->Alstm[0] + E + F, e->split ? e->Alstm[1] + E + F : nullptr, E + F + H, R, H, st));
  CU(cudaMemcpyAsync(e->c, d_c, sizeof(float) * R * H, cudaMemcpyDeviceToDevice, st));
  if (e->step_core(R, rows_per_feature, n_keys, d_features, reinterpret_cast<const long long*>(d_tokens),
                   d_attentions_out, n_keys, st))
    return 1;
  const bool mi = d_h_lm != nullptr;
  if (mi) {
    const int El = e->cfg.lm_embedding_size, Hl = e->cfg.lm_hidden_size;
    RC(launch_split_rows(d_h_lm, Hl, e->Alm0[0] + El, e->split ? e->Alm0[1] + El : nullptr, El + Hl, R, Hl, st));
    RC(launch_split_rows(d_h_lm + static_cast<size_t>(R) * Hl, Hl, e->Alm1[0] + Hl, e->split ? e->Alm1[1] + Hl : nullptr,
                         2 * Hl, R, Hl, st));
    CU(cudaMemcpyAsync(e->lm_c0, d_c_lm, sizeof(float) * R * Hl, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(e->lm_c1, d_c_lm + static_cast<size_t>(R) * Hl, sizeof(float) * R * Hl, cudaMemcpyDeviceToDevice, st));
    if (e->lm_step_core(R, reinterpret_cast<const long long*>(d_tokens), false, st)) return 1;
    CU(cudaMemcpyAsync(d_h_lm, e->lm_h_f32, sizeof(float) * 2 * R * Hl, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(d_c_lm, e->lm_c0, sizeof(float) * R * Hl, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(d_c_lm + static_cast<size_t>(R) * Hl, e->lm_c1, sizeof(float) * R * Hl, cudaMemcpyDeviceToDevice, st));
  }
  RowArgs ra{};
  ra.logits = e->logits; ra.logits_lm = mi ? e->logits_lm : nullptr; ra.ld = e->ldv; ra.R = R; ra.V = V;
  ra.temperature = temperature; ra.pred_out = d_predictions_out; ra.pred_pitch = V; ra.beam = 0;
  RC(launch_row_logsoftmax(ra, st));
  CU(cudaMemcpyAsync(d_h, e->hnew_f32, sizeof(float) * R * H, cudaMemcpyDeviceToDevice, st));
  CU(cudaMemcpyAsync(d_c, e->cnew, sizeof(float) * R * H, cudaMemcpyDeviceToDevice, st));
  return 0;
}

int milan_decode_greedy(MilanEngine* engine, const float* d_features, int32_t B, int32_t n_keys, int32_t length,
                        int32_t mi, float temperature, const int64_t* d_forced, int64_t* d_tokens_out,
                        float* d_scores_out, float* d_predictions_out, float* d_attentions_out, void* stream) {
  CHECK_READY(engine);
  CHECK_DECODER(engine);
  return engine->decode_greedy(d_features, B, n_keys, length, mi, temperature,
                               reinterpret_cast<const long long*>(d_forced),
                               reinterpret_cast<long long*>(d_tokens_out), d_scores_out, d_predictions_out,
                               d_attentions_out, static_cast<cudaStream_t>(stream));
}

int milan_decode_beam(MilanEngine* engine, const float* d_features, int32_t B, int32_t n_keys, int32_t length,
                      int32_t beam, int32_t group_size, int32_t rerank, int32_t mi, float temperature,
                      int64_t* d_beam_tokens_out, float* d_beam_scores_out, int32_t* d_group_steps_out,
                      int64_t* d_tokens_out, float* d_scores_out, float* d_lm_scores_out, void* stream) {
  CHECK_READY(engine);
  CHECK_DECODER(engine);
  return engine->decode_beam(d_features, B, n_keys, length, beam, group_size, rerank, mi, temperature,
                             reinterpret_cast<long long*>(d_beam_tokens_out), d_beam_scores_out, d_group_steps_out,
                             reinterpret_cast<long long*>(d_tokens_out), d_scores_out, d_lm_scores_out,
                             static_cast<cudaStream_t>(stream));
}

int milan_lm_score(MilanEngine* engine, const int64_t* d_inputs, int32_t M, int32_t T1, float* d_out, void* stream) {
  CHECK_READY(engine);
  CHECK_DECODER(engine);
  MilanEngine* e = engine;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!e->cfg.has_lm) return fail("engine has no LM");
  if (M > e->Rmax) return fail("milan_lm_score: M=%d exceeds capacity %d", M, e->Rmax);
  const int length = T1 - 1;
  if (length < 1 || length > e->cfg.max_length) return fail("milan_lm_score: T1=%d unsupported", T1);
  // seqs = inputs[:, 1:]; the leading token is assumed to be <start> like the reference does (lms.py:66-70)
  CU(cudaMemcpy2DAsync(e->seqs, sizeof(long long) * length, d_inputs + 1, sizeof(long long) * T1,
                       sizeof(long long) * length, M, cudaMemcpyDeviceToDevice, st));
  // one group spanning every row, T = length
  e->host_T = length;
  CU(cudaMemcpyAsync(e->group_T, &e->host_T, sizeof(int), cudaMemcpyHostToDevice, st));
  if (e->lm_score_seqs(e->seqs, M, length, 1, 1 << 30, st)) return 1;
  CU(cudaMemcpyAsync(d_out, e->lm_scores, sizeof(float) * M, cudaMemcpyDeviceToDevice, st));
  return 0;
}

int milan_lm_logprobs(MilanEngine* engine, const int64_t* d_inputs, int32_t M, int32_t T, float* d_out, void* stream) {
  CHECK_READY(engine);
  CHECK_DECODER(engine);
  MilanEngine* e = engine;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!e->cfg.has_lm) return fail("engine has no LM");
  if (M < 1 || M > e->Rmax) return fail("milan_lm_logprobs: M=%d outside [1, capacity %d]", M, e->Rmax);
  if (T < 1) return fail("milan_lm_logprobs: T=%d", T);
  const int V = e->cfg.vocab_size;
  if (e->lm_reset(M, st)) return 1;
  for (int t = 0; t < T; ++t) {
    // column t of the (M, T) inputs -> the contiguous token vector the embedding kernel reads
    CU(cudaMemcpy2DAsync(e->tok_cur, sizeof(long long), d_inputs + t, sizeof(long long) * T, sizeof(long long), M,
                         cudaMemcpyDeviceToDevice, st));
    if (e->lm_step_core(M, e->tok_cur, false, st)) return 1;
    RowArgs ra{};
    ra.logits = e->logits_lm; ra.ld = e->ldv; ra.R = M; ra.V = V; ra.temperature = 0.0f;
    ra.pred_out = d_out + static_cast<long long>(t) * V; ra.pred_pitch = static_cast<long long>(T) * V;
    ra.stop_index = e->cfg.stop_index;
    RC(launch_row_logsoftmax(ra, st));
  }
  return 0;
}

}  // extern "C"

// Shared body of milan_describe_host / milan_describe_device: chunks of whole reference batches through
//   copy stream (host inputs only): exemplars of chunk c+1 -> staging set (c+1) % 2
//   `st`                          : encoder of chunk c -> feature buffer c % 2
//   decode stream (high priority) : beam search / rerank of chunk c-1, under the encoder of chunk c
// Results of every chunk land in d_tokens_all / d_scores_all / d_steps_all; the caller reads them after `st` (which
// is made to wait for the decode stream at the end).
int MilanEngine::describe(const uint8_t* images, const uint8_t* masks, bool host_inputs, int n_neurons, int k,
                          int strategy, int mi, int length, int beam, int group_size, float temperature,
                          cudaStream_t st) {
  MilanEngine* e = this;
  // the result buffers hold max_length ids per neuron and chunks write at done * length: reject what would overrun them
  if (length < 1 || length > cfg.max_length) return fail("length %d outside [1, max_length = %d]", length, cfg.max_length);
  if (strategy < 0 || strategy > 2) return fail("unknown strategy %d", strategy);
  if (strategy != 0 && (beam < 1 || beam > cfg.max_beam || beam > kMaxBeam))
    return fail("beam size %d unsupported (max %d)", beam, std::min(cfg.max_beam, kMaxBeam));
  const int n_keys = k * (spatial ? kSpatialKeys : 1);  // Decoder.encode: view(batch, -1, feature_size)
  // neurons per chunk: bounded by encoder image capacity and decoder capacity; whole reference groups only
  int chunk = std::min(cfg.max_images / k, cfg.max_neurons);
  if (chunk >= group_size) chunk = chunk / group_size * group_size;
  if (chunk < 1) return fail("max_images %d too small for k=%d", cfg.max_images, k);
  if (strategy != 0 && chunk % group_size != 0 && chunk < n_neurons)
    return fail("engine capacity (%d neurons/chunk) smaller than group_size %d", chunk, group_size);
  const size_t img_bytes = static_cast<size_t>(k) * 3 * 224 * 224, msk_bytes = static_cast<size_t>(k) * 224 * 224;
  // ---- lazily created pipeline resources
  if (copy_stream == nullptr) {
    CU(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    int lo_prio = 0, hi_prio = 0;
    CU(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
    CU(cudaStreamCreateWithPriority(&dec_stream, cudaStreamNonBlocking, hi_prio));
    for (int b = 0; b < 2; ++b) {
      CU(cudaEventCreateWithFlags(&ev_copied[b], cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&ev_consumed[b], cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&ev_feat_ready[b], cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&ev_feat_free[b], cudaEventDisableTiming));
    }
    CU(cudaEventCreateWithFlags(&ev_dec_done, cudaEventDisableTiming));
  }
  if (static_cast<size_t>(n_neurons) > results_cap) {
    const size_t cap = static_cast<size_t>(n_neurons);
    if (dalloc(&d_tokens_all, cap * cfg.max_length)) return 1;
    if (dalloc(&d_scores_all, cap)) return 1;
    if (dalloc(&d_steps_all, cap)) return 1;
    results_cap = cap;
  }
  const int n_chunks = (n_neurons + chunk - 1) / chunk;
  const int groups_per_chunk = (chunk + group_size - 1) / group_size;
  // with profiling on the chunks run back to back on `st`, so that the conv kernels' event-timed durations are not
  // stretched by decode kernels sharing the SMs
  const bool overlap = overlap_decode && n_chunks > 1 && !profiling;
  cudaStream_t ds = overlap ? dec_stream : st;
  float* feats[2] = {feat_enc, feat_enc2};
  if (profiling) conv_events_used = 0;
  // Exemplars of chunk c -> staging set c % 2 on the copy stream, once the encoder has consumed what the set held
  // (chunk c - 2). The host buffers are read asynchronously only if they are pinned; pageable memory still works.
  auto issue_copy = [&](int c) -> int {
    const int b = c & 1;
    const int done = c * chunk;
    const int nb = std::min(chunk, n_neurons - done);
    NvtxRange range("milan.h2d chunk %d (%d neurons)", c, nb);
    if (c >= 2) CU(cudaStreamWaitEvent(copy_stream, ev_consumed[b], 0));
    CU(cudaMemcpyAsync(d_img_stage[b], images + done * img_bytes, nb * img_bytes, cudaMemcpyHostToDevice, copy_stream));
    CU(cudaMemcpyAsync(d_mask_stage[b], masks + done * msk_bytes, nb * msk_bytes, cudaMemcpyHostToDevice, copy_stream));
    CU(cudaEventRecord(ev_copied[b], copy_stream));
    return 0;
  };
  if (host_inputs && issue_copy(0)) return 1;
  for (int c = 0; c < n_chunks; ++c) {
    const int b = c & 1;
    const int done = c * chunk;
    const int nb = std::min(chunk, n_neurons - done);
    const uint8_t* d_img = host_inputs ? d_img_stage[b] : images + done * img_bytes;
    const uint8_t* d_msk = host_inputs ? d_mask_stage[b] : masks + done * msk_bytes;
    if (host_inputs) {
      if (c + 1 < n_chunks && issue_copy(c + 1)) return 1;
      CU(cudaStreamWaitEvent(st, ev_copied[b], 0));
    }
    if (overlap && c >= 2) CU(cudaStreamWaitEvent(st, ev_feat_free[b], 0));  // decode of chunk c-2 read feats[b]
    profiling_append = profiling && c > 0;  // one conv-event list across the chunks of this call
    {
      NvtxRange range("milan.encode chunk %d (%d images)", c, nb * k);
      if (e->encode(d_img, d_msk, nb * k, MILAN_DTYPE_U8, feats[b], st)) return 1;
    }
    profiling_append = false;
    if (host_inputs) CU(cudaEventRecord(ev_consumed[b], st));
    if (overlap) {
      CU(cudaEventRecord(ev_feat_ready[b], st));
      CU(cudaStreamWaitEvent(ds, ev_feat_ready[b], 0));
    }
    long long* d_tok = d_tokens_all + static_cast<size_t>(done) * length;
    float* d_sc = d_scores_all + done;
    NvtxRange range("milan.decode chunk %d (strategy %d)", c, strategy);
    if (strategy == 0) {
      if (decode_greedy(feats[b], nb, n_keys, length, mi, temperature, nullptr, d_tok, d_sc, nullptr, nullptr, ds)) return 1;
    } else {
      if (decode_beam(feats[b], nb, n_keys, length, beam, group_size, strategy == 2, strategy == 1 ? mi : 0, temperature,
                      nullptr, nullptr, d_steps_all + c * groups_per_chunk, d_tok, d_sc, nullptr, ds))
        return 1;
    }
    if (overlap) CU(cudaEventRecord(ev_feat_free[b], ds));
  }
  if (overlap) {  // later work on `st` (the result copies) follows the last decode
    CU(cudaEventRecord(ev_dec_done, ds));
    CU(cudaStreamWaitEvent(st, ev_dec_done, 0));
  }
  host_chunk = chunk;
  host_groups_per_chunk = groups_per_chunk;
  return 0;
}

extern "C" {

int milan_describe_host(MilanEngine* engine, const uint8_t* h_images, const uint8_t* h_masks, int32_t n_neurons,
                        int32_t k, int32_t strategy, int32_t mi, int32_t length, int32_t beam, int32_t group_size,
                        float temperature, int64_t* h_tokens_out, float* h_scores_out, int32_t* h_steps_out,
                        void* stream) {
  CHECK_READY(engine);
  CHECK_DECODER(engine);
  MilanEngine* e = engine;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!e->cfg.has_encoder) return fail("engine was created without an encoder");
  const int n_keys = k * (e->spatial ? kSpatialKeys : 1);
  if (n_keys > e->cfg.max_keys) return fail("k=%d (%d keys) exceeds max_keys %d", k, n_keys, e->cfg.max_keys);
  if (group_size <= 0) group_size = 16;
  if (n_neurons <= 0) return 0;
  if (e->describe(h_images, h_masks, true, n_neurons, k, strategy, mi, length, beam, group_size, temperature, st)) return 1;
  // ---- one device->host read of everything, then a single synchronisation
  const int n_chunks = (n_neurons + e->host_chunk - 1) / e->host_chunk;
  std::vector<int> steps(static_cast<size_t>(n_chunks) * e->host_groups_per_chunk, length);
  CU(cudaMemcpyAsync(h_tokens_out, e->d_tokens_all, sizeof(long long) * static_cast<size_t>(n_neurons) * length,
                     cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(h_scores_out, e->d_scores_all, sizeof(float) * n_neurons, cudaMemcpyDeviceToHost, st));
  if (strategy != 0)
    CU(cudaMemcpyAsync(steps.data(), e->d_steps_all, sizeof(int) * steps.size(), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  for (int i = 0; i < n_neurons; ++i)
    h_steps_out[i] = strategy == 0 ? length
                                   : steps[(i / e->host_chunk) * e->host_groups_per_chunk + (i % e->host_chunk) / group_size];
  if (e->collect_conv_events(st)) return 1;
  return 0;
}

int milan_describe_device(MilanEngine* engine, const uint8_t* d_images, const uint8_t* d_masks, int32_t n_neurons,
                          int32_t k, int32_t strategy, int32_t mi, int32_t length, int32_t beam, int32_t group_size,
                          float temperature, int64_t* d_tokens_out, float* d_scores_out, void* stream) {
  CHECK_READY(engine);
  CHECK_DECODER(engine);
  MilanEngine* e = engine;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!e->cfg.has_encoder) return fail("engine was created without an encoder");
  const int n_keys = k * (e->spatial ? kSpatialKeys : 1);
  if (n_keys > e->cfg.max_keys) return fail("k=%d (%d keys) exceeds max_keys %d", k, n_keys, e->cfg.max_keys);
  if (group_size <= 0) group_size = 16;
  if (n_neurons <= 0) return 0;
  if (reinterpret_cast<uintptr_t>(d_images) % 16 != 0 || (static_cast<size_t>(k) * 3 * 224 * 224) % 16 != 0)
    return fail("milan_describe_device: d_images must be 16-byte aligned");
  if (e->describe(d_images, d_masks, false, n_neurons, k, strategy, mi, length, beam, group_size, temperature, st)) return 1;
  CU(cudaMemcpyAsync(d_tokens_out, e->d_tokens_all, sizeof(long long) * static_cast<size_t>(n_neurons) * length,
                     cudaMemcpyDeviceToDevice, st));
  CU(cudaMemcpyAsync(d_scores_out, e->d_scores_all, sizeof(float) * n_neurons, cudaMemcpyDeviceToDevice, st));
  if (e->profiling && e->collect_conv_events(st)) return 1;  // (synchronises `st`)
  return 0;
}

int64_t milan_launch_count(void) { return total_launch_count(); }

int milan_set_profiling(MilanEngine* engine, int32_t enabled) {
  if (engine == nullptr) return fail("null engine");
  engine->profiling = enabled != 0;
  engine->prof_conv_ms = engine->prof_enc_ms = engine->prof_dec_ms = 0;
  engine->prof_conv_launches = 0;
  return 0;
}

int milan_get_profile(MilanEngine* engine, float* conv_ms, float* encoder_ms, float* decoder_ms, int64_t* conv_launches) {
  if (engine == nullptr) return fail("null engine");
  if (conv_ms) *conv_ms = engine->prof_conv_ms;
  if (encoder_ms) *encoder_ms = engine->prof_enc_ms;
  if (decoder_ms) *decoder_ms = engine->prof_dec_ms;
  if (conv_launches) *conv_launches = engine->prof_conv_launches;
  return 0;
}

}  // extern "C"
