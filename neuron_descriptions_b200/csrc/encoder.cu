// Non-GEMM kernels of the pyramid encoder (reference: PyramidConvEncoder.forward, src/milan/encoders.py:286-320).
//   stem_pack        : normalise (encoders.py:294-295) + repack to padded NHWC4 bf16 hi/lo for the im2col-free stem
//   mask_pyramid     : bilinear (align_corners=False) mask downsample to the 5 retained resolutions +
//                      per-image sum-normalisation with the all-zero exception (encoders.py:303-314)
//   masked_pool      : weighted spatial sum of a retained NHWC map (encoders.py:317)
//   bn_relu_maxpool  : bn1 + ReLU + 3x3/2 max-pool after the stem (torchvision resnet forward)
// All are HBM/L2-bound streaming kernels: coalesced along channels, 128-bit where the layout allows.
#include "encoder.h"
#include "conv_gemm.h"
#include "ptx.cuh"

#include <cstdint>

namespace milan {

namespace {

constexpr int kImg = 224;
constexpr int kStemOut = 112;

template <typename T>
__device__ __forceinline__ float load_pixel(const T* p);
template <>
__device__ __forceinline__ float load_pixel<uint8_t>(const uint8_t* p) {
  // TopImagesDataset: images.float() * fp32(1/255)  (src/milannotations/datasets.py:191-197,
  // src/deps/netdissect/renormalize.py:118-139)
  return __fmul_rn(static_cast<float>(*p), 0.00392156862745098f);  // no FMA contraction with the mean subtract
}

template <typename T>
__device__ __forceinline__ void load_pixels4(const T* p, float (&v)[4]);
template <>
__device__ __forceinline__ void load_pixels4<uint8_t>(const uint8_t* p, float (&v)[4]) {
  const uchar4 q = *reinterpret_cast<const uchar4*>(p);
  const uint8_t b[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = load_pixel<uint8_t>(&b[j]);
}
template <>
__device__ __forceinline__ void load_pixels4<float>(const float* p, float (&v)[4]) {
  const float4 q = *reinterpret_cast<const float4*>(p);
  v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
}

template <typename T>
__device__ __forceinline__ void load_masks4(const T* p, float (&v)[4]);
template <>
__device__ __forceinline__ void load_masks4<uint8_t>(const uint8_t* p, float (&v)[4]) {
  const uchar4 q = *reinterpret_cast<const uchar4*>(p);
  v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
}
template <>
__device__ __forceinline__ void load_masks4<float>(const float* p, float (&v)[4]) {
  const float4 q = *reinterpret_cast<const float4*>(p);
  v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
}

// Normalise + repack NCHW images into the zero-padded NHWC4 bf16 layout whose raw rows the stem GEMM loads (its MMA
// descriptor forms the 7-tap windows): [n][232][232][4], pixel (ih, iw) at (ih + 3, iw + 4), channel 3 = 0.
// One thread per 4 consecutive padded pixels (the x padding of 4 keeps every group fully inside or fully outside
// the image): one 4-pixel load per channel, two 16-byte stores per plane.
template <typename T>
__global__ void __launch_bounds__(256) stem_pack_kernel(const T* __restrict__ images, int n_images,
                                                        __nv_bfloat16* __restrict__ p_hi,
                                                        __nv_bfloat16* __restrict__ p_lo, float3 mean, float3 stdv,
                                                        int split, const T* __restrict__ masks) {
  constexpr int PH = kStemPadH, PW4 = kStemPadW / 4;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(n_images) * PH * PW4;
  if (idx >= total) return;
  const int x4 = idx % PW4;
  const int y = (idx / PW4) % PH;
  const int n = idx / (PH * PW4);
  const int ih = y - 3, iw0 = 4 * x4 - 4;
  float v[3][4] = {};
  if (ih >= 0 && ih < kImg && iw0 >= 0 && iw0 < kImg) {
    const T* img = images + static_cast<long long>(n) * 3 * kImg * kImg + static_cast<long long>(ih) * kImg + iw0;
    const float m[3] = {mean.x, mean.y, mean.z};
    const float s[3] = {stdv.x, stdv.y, stdv.z};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      load_pixels4<T>(img + c * kImg * kImg, v[c]);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[c][j] = __fdiv_rn(__fsub_rn(v[c][j], m[c]), s[c]);
    }
    if (masks != nullptr) {  // SpatialConvEncoder: images * masks after normalisation (encoders.py:208-211)
      float mk[4];
      load_masks4<T>(masks + static_cast<long long>(n) * kImg * kImg + static_cast<long long>(ih) * kImg + iw0, mk);
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) v[c][j] = __fmul_rn(v[c][j], mk[j]);
    }
  }
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
  uint32_t hw[8], lw[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat16 h[3], l[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) split_bf16(v[c][j], h[c], l[c]);
    hw[2 * j] = pack_bf16x2(h[0], h[1]); hw[2 * j + 1] = pack_bf16x2(h[2], zero);
    lw[2 * j] = pack_bf16x2(l[0], l[1]); lw[2 * j + 1] = pack_bf16x2(l[2], zero);
  }
  uint4* dh = reinterpret_cast<uint4*>(p_hi + idx * 16);
  dh[0] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
  dh[1] = make_uint4(hw[4], hw[5], hw[6], hw[7]);
  if (split) {
    uint4* dl = reinterpret_cast<uint4*>(p_lo + idx * 16);
    dl[0] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    dl[1] = make_uint4(lw[4], lw[5], lw[6], lw[7]);
  }
}

template <typename T>
__device__ __forceinline__ float load_mask(const T* p) { return static_cast<float>(*p); }

__device__ __forceinline__ float block_reduce_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.0f;
  const int nw = blockDim.x >> 5;
  for (int i = 0; i < nw; ++i) t += red[i];
  return t;
}
__device__ __forceinline__ float block_reduce_max(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = red[0];
  const int nw = blockDim.x >> 5;
  for (int i = 1; i < nw; ++i) t = fmaxf(t, red[i]);
  return t;
}

// One CTA per image: all 5 levels.
template <typename T>
__global__ void __launch_bounds__(256) mask_pyramid_kernel(const T* __restrict__ masks, float* __restrict__ wts) {
  __shared__ float red[8];
  const int n = blockIdx.x;
  const T* m = masks + static_cast<long long>(n) * kImg * kImg;
  float* out = wts + static_cast<long long>(n) * kMaskPyramidSize;
  int off = 0;
  for (int level = 0; level < 5; ++level) {
    const int S = kStemOut >> level;  // 112, 56, 28, 14, 7
    const int scale = kImg / S;
    const int c0 = scale / 2 - 1;
    float local_sum = 0.0f, local_max = 0.0f;
    for (int p = threadIdx.x; p < S * S; p += blockDim.x) {
      const int i = p / S, j = p - i * S;
      const T* q = m + (i * scale + c0) * kImg + j * scale + c0;
      // upsample_bilinear2d: h0lambda*(w0lambda*p00 + w1lambda*p01) + h1lambda*(w0lambda*p10 + w1lambda*p11)
      const float top = 0.5f * load_mask(q) + 0.5f * load_mask(q + 1);
      const float bot = 0.5f * load_mask(q + kImg) + 0.5f * load_mask(q + kImg + 1);
      const float v = 0.5f * top + 0.5f * bot;
      out[off + p] = v;
      local_sum += v;
      local_max = fmaxf(local_max, fabsf(v));
    }
    const float total = block_reduce_sum(local_sum, red);
    const float amax = block_reduce_max(local_max, red);
    // valid = ~isclose(ms, 0).all(): |v| <= 1e-8 everywhere -> leave un-normalised (encoders.py:311-314)
    if (amax > 1e-8f) {
      for (int p = threadIdx.x; p < S * S; p += blockDim.x) out[off + p] = out[off + p] / total;
    }
    off += S * S;
    __syncthreads();
  }
}

// pooled[n][c] = sum_p w[n][p] * (hi + lo)[n][p][c]. CTA = (image, 64-channel group); each warp takes chunks of
// 32 consecutive pixels: one coalesced load of their weights, a ballot of the non-zero ones (masks are ~1-5 %
// dense), then per surviving pixel lane = 2 channels (one 4-byte load per plane, 128 B per warp).
__global__ void __launch_bounds__(256) masked_pool_kernel(const __nv_bfloat16* __restrict__ hi,
                                                          const __nv_bfloat16* __restrict__ lo,
                                                          const float* __restrict__ wts, int wts_stride, int P, int C,
                                                          float* __restrict__ out, int out_stride) {
  __shared__ float2 acc_s[8][32];
  const int n = blockIdx.x;
  const int cg = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* w = wts + static_cast<long long>(n) * wts_stride;
  const long long base = static_cast<long long>(n) * P * C + cg * 64 + lane * 2;
  float2 acc = make_float2(0.f, 0.f);
  for (int p0 = warp * 32; p0 < P; p0 += 8 * 32) {
    const float mine = (p0 + lane < P) ? __ldg(w + p0 + lane) : 0.0f;
    unsigned live = __ballot_sync(0xffffffffu, mine != 0.0f);
    while (live != 0u) {
      const int src = __ffs(live) - 1;
      live &= live - 1;
      const float wp = __shfl_sync(0xffffffffu, mine, src);
      const long long off = base + static_cast<long long>(p0 + src) * C;
      const uint32_t h = __ldg(reinterpret_cast<const uint32_t*>(hi + off));
      float x0 = bf16_lo_to_f32(h), x1 = bf16_hi_to_f32(h);
      if (lo != nullptr) {
        const uint32_t l = __ldg(reinterpret_cast<const uint32_t*>(lo + off));
        x0 += bf16_lo_to_f32(l);
        x1 += bf16_hi_to_f32(l);
      }
      acc.x = fmaf(wp, x0, acc.x);
      acc.y = fmaf(wp, x1, acc.y);
    }
  }
  acc_s[warp][lane] = acc;
  __syncthreads();
  if (warp == 0) {
    float2 t = acc_s[0][lane];
#pragma unroll
    for (int i = 1; i < 8; ++i) { t.x += acc_s[i][lane].x; t.y += acc_s[i][lane].y; }
    float* o = out + static_cast<long long>(n) * out_stride + cg * 64 + lane * 2;
    o[0] = t.x;
    o[1] = t.y;
  }
}

// y[n][oh][ow][c] = max_{3x3, stride 2, pad 1} relu(alpha[c] * x + beta[c]);  C = 64, thread = 8 channels
// (one 16-byte load per plane and window tap; 8 threads cover a pixel's 128-byte channel row). A CTA owns a
// 4 x 8 tile of output pixels so the input rows / columns shared by neighbouring windows are re-read from L1
// (9 x 17 input pixels per 32 outputs instead of 9 per output through L2).
__global__ void __launch_bounds__(256) bn_relu_maxpool_kernel(const __nv_bfloat16* __restrict__ x_hi,
                                                              const __nv_bfloat16* __restrict__ x_lo,
                                                              const float* __restrict__ alpha,
                                                              const float* __restrict__ beta, int n_images,
                                                              __nv_bfloat16* __restrict__ y_hi,
                                                              __nv_bfloat16* __restrict__ y_lo) {
  constexpr int C = 64, IN = 112, OUT = 56, TW = 8, TH = 4, TX = OUT / TW, TY = OUT / TH;
  const int cg = threadIdx.x & 7;
  const int ow = (blockIdx.x % TX) * TW + ((threadIdx.x >> 3) & 7);
  const int oh = ((blockIdx.x / TX) % TY) * TH + (threadIdx.x >> 6);
  const int n = blockIdx.x / (TX * TY);
  if (n >= n_images) return;
  const long long pix = (static_cast<long long>(n) * OUT + oh) * OUT + ow;
  float a[8], b[8], m[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a[j] = __ldg(alpha + cg * 8 + j);
    b[j] = __ldg(beta + cg * 8 + j);
    m[j] = 0.0f;  // relu output >= 0 and every window holds >= 1 valid pixel
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int ih = oh * 2 + r - 1;
    if (ih < 0 || ih >= IN) continue;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int iw = ow * 2 + s - 1;
      if (iw < 0 || iw >= IN) continue;
      const long long off = ((static_cast<long long>(n) * IN + ih) * IN + iw) * C + cg * 8;
      const uint4 h = __ldg(reinterpret_cast<const uint4*>(x_hi + off));
      float v[8] = {bf16_lo_to_f32(h.x), bf16_hi_to_f32(h.x), bf16_lo_to_f32(h.y), bf16_hi_to_f32(h.y),
                    bf16_lo_to_f32(h.z), bf16_hi_to_f32(h.z), bf16_lo_to_f32(h.w), bf16_hi_to_f32(h.w)};
      if (x_lo != nullptr) {
        const uint4 l = __ldg(reinterpret_cast<const uint4*>(x_lo + off));
        v[0] += bf16_lo_to_f32(l.x); v[1] += bf16_hi_to_f32(l.x); v[2] += bf16_lo_to_f32(l.y); v[3] += bf16_hi_to_f32(l.y);
        v[4] += bf16_lo_to_f32(l.z); v[5] += bf16_hi_to_f32(l.z); v[6] += bf16_lo_to_f32(l.w); v[7] += bf16_hi_to_f32(l.w);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], fmaf(v[j], a[j], b[j]));
    }
  }
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split_bf16x2(m[2 * j], m[2 * j + 1], hi[j], lo[j]);
  const long long o = pix * C + cg * 8;
  *reinterpret_cast<uint4*>(y_hi + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  if (y_lo != nullptr) *reinterpret_cast<uint4*>(y_lo + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// ---------------------------------------------------------------- AlexNet pyramid pieces
template <typename T>
__global__ void __launch_bounds__(256) alexnet_im2col_kernel(const T* __restrict__ images, int n_images,
                                                             __nv_bfloat16* __restrict__ a_hi,
                                                             __nv_bfloat16* __restrict__ a_lo, float3 mean,
                                                             float3 stdv, int split) {
  constexpr int O = 55, K2 = kAlexK0 / 2;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(n_images) * O * O * K2;
  if (idx >= total) return;
  const int k2 = idx % K2;
  const long long row = idx / K2;
  const int ow = row % O;
  const int oh = (row / O) % O;
  const int n = row / (O * O);
  const float m[3] = {mean.x, mean.y, mean.z};
  const float sd[3] = {stdv.x, stdv.y, stdv.z};
  float v[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int k = 2 * k2 + j;
    v[j] = 0.0f;
    if (k < 363) {
      const int c = k / 121, r = (k / 11) % 11, t = k % 11;
      const int ih = oh * 4 + r - 2, iw = ow * 4 + t - 2;
      if (ih >= 0 && ih < kImg && iw >= 0 && iw < kImg) {
        const T* px = images + ((static_cast<long long>(n) * 3 + c) * kImg + ih) * kImg + iw;
        float x;
        if (sizeof(T) == 1) x = __fmul_rn(static_cast<float>(*px), 0.00392156862745098f);
        else x = static_cast<float>(*px);
        v[j] = __fdiv_rn(__fsub_rn(x, m[c]), sd[c]);
      }
    }
  }
  uint32_t h, l;
  split_bf16x2(v[0], v[1], h, l);
  reinterpret_cast<uint32_t*>(a_hi)[idx] = h;
  if (split) reinterpret_cast<uint32_t*>(a_lo)[idx] = l;
}

// One CTA per image. Source index / weights follow ATen's area_pixel_compute_source_index for
// align_corners=False: src = max(0, scale * (dst + 0.5) - 0.5), scale = in / out (float).
template <typename T>
__global__ void __launch_bounds__(256) mask_resize_kernel(const T* __restrict__ masks, int S,
                                                          float* __restrict__ out, int out_stride) {
  __shared__ float red[8];
  const int n = blockIdx.x;
  const T* m = masks + static_cast<long long>(n) * kImg * kImg;
  float* o = out + static_cast<long long>(n) * out_stride;
  const float scale = static_cast<float>(kImg) / static_cast<float>(S);
  float local_sum = 0.0f, local_max = 0.0f;
  for (int p = threadIdx.x; p < S * S; p += blockDim.x) {
    const int i = p / S, j = p - i * S;
    float sy = __fsub_rn(__fmul_rn(scale, static_cast<float>(i) + 0.5f), 0.5f);
    float sx = __fsub_rn(__fmul_rn(scale, static_cast<float>(j) + 0.5f), 0.5f);
    sy = sy < 0.0f ? 0.0f : sy;
    sx = sx < 0.0f ? 0.0f : sx;
    const int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
    const int y1 = y0 + (y0 < kImg - 1 ? 1 : 0), x1 = x0 + (x0 < kImg - 1 ? 1 : 0);
    const float ly1 = sy - static_cast<float>(y0), lx1 = sx - static_cast<float>(x0);
    const float ly0 = 1.0f - ly1, lx0 = 1.0f - lx1;
    const float p00 = load_mask(m + y0 * kImg + x0), p01 = load_mask(m + y0 * kImg + x1);
    const float p10 = load_mask(m + y1 * kImg + x0), p11 = load_mask(m + y1 * kImg + x1);
    const float v = __fadd_rn(__fmul_rn(ly0, __fadd_rn(__fmul_rn(lx0, p00), __fmul_rn(lx1, p01))),
                              __fmul_rn(ly1, __fadd_rn(__fmul_rn(lx0, p10), __fmul_rn(lx1, p11))));
    o[p] = v;
    local_sum += v;
    local_max = fmaxf(local_max, fabsf(v));
  }
  const float total = block_reduce_sum(local_sum, red);
  const float amax = block_reduce_max(local_max, red);
  if (amax > 1e-8f) {
    for (int p = threadIdx.x; p < S * S; p += blockDim.x) o[p] = o[p] / total;
  }
}

__global__ void __launch_bounds__(256) maxpool3x3s2_kernel(const __nv_bfloat16* __restrict__ x_hi,
                                                           const __nv_bfloat16* __restrict__ x_lo, int n_images,
                                                           int H, int W, int C, __nv_bfloat16* __restrict__ y_hi,
                                                           __nv_bfloat16* __restrict__ y_lo) {
  const int Ho = (H - 3) / 2 + 1, Wo = (W - 3) / 2 + 1, G = C / 8;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(n_images) * Ho * Wo * G;
  if (idx >= total) return;
  const int cg = idx % G;
  const long long pix = idx / G;
  const int ow = pix % Wo;
  const int oh = (pix / Wo) % Ho;
  const int n = pix / (static_cast<long long>(Wo) * Ho);
  float m[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      const long long off = ((static_cast<long long>(n) * H + oh * 2 + r) * W + ow * 2 + t) * C + cg * 8;
      const uint4 h = __ldg(reinterpret_cast<const uint4*>(x_hi + off));
      float v[8] = {bf16_lo_to_f32(h.x), bf16_hi_to_f32(h.x), bf16_lo_to_f32(h.y), bf16_hi_to_f32(h.y),
                    bf16_lo_to_f32(h.z), bf16_hi_to_f32(h.z), bf16_lo_to_f32(h.w), bf16_hi_to_f32(h.w)};
      if (x_lo != nullptr) {
        const uint4 l = __ldg(reinterpret_cast<const uint4*>(x_lo + off));
        v[0] += bf16_lo_to_f32(l.x); v[1] += bf16_hi_to_f32(l.x); v[2] += bf16_lo_to_f32(l.y); v[3] += bf16_hi_to_f32(l.y);
        v[4] += bf16_lo_to_f32(l.z); v[5] += bf16_hi_to_f32(l.z); v[6] += bf16_lo_to_f32(l.w); v[7] += bf16_hi_to_f32(l.w);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], v[j]);
    }
  }
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split_bf16x2(m[2 * j], m[2 * j + 1], hi[j], lo[j]);
  const long long o = pix * C + cg * 8;
  *reinterpret_cast<uint4*>(y_hi + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  if (y_lo != nullptr) *reinterpret_cast<uint4*>(y_lo + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

__global__ void planes_to_f32_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                     long long n2, float* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n2) return;
  const uint32_t h = reinterpret_cast<const uint32_t*>(hi)[i];
  float x0 = bf16_lo_to_f32(h), x1 = bf16_hi_to_f32(h);
  if (lo != nullptr) {
    const uint32_t l = reinterpret_cast<const uint32_t*>(lo)[i];
    x0 += bf16_lo_to_f32(l);
    x1 += bf16_hi_to_f32(l);
  }
  reinterpret_cast<float2*>(out)[i] = make_float2(x0, x1);
}

}  // namespace

int launch_alexnet_im2col(const void* images, int dtype, int n_images, __nv_bfloat16* a_hi, __nv_bfloat16* a_lo,
                          const float mean[3], const float stdv[3], int split, cudaStream_t stream) {
  const long long total = static_cast<long long>(n_images) * 55 * 55 * (kAlexK0 / 2);
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  const float3 m = make_float3(mean[0], mean[1], mean[2]);
  const float3 s = make_float3(stdv[0], stdv[1], stdv[2]);
  if (dtype == 0)
    alexnet_im2col_kernel<uint8_t><<<blocks, 256, 0, stream>>>(static_cast<const uint8_t*>(images), n_images, a_hi,
                                                               a_lo, m, s, split);
  else
    alexnet_im2col_kernel<float><<<blocks, 256, 0, stream>>>(static_cast<const float*>(images), n_images, a_hi, a_lo,
                                                             m, s, split);
  note_launch();
  return static_cast<int>(cudaGetLastError());
}

int launch_mask_resize(const void* masks, int dtype, int n_images, int S, float* out, int out_stride,
                       cudaStream_t stream) {
  if (dtype == 0)
    mask_resize_kernel<uint8_t><<<n_images, 256, 0, stream>>>(static_cast<const uint8_t*>(masks), S, out, out_stride);
  else
    mask_resize_kernel<float><<<n_images, 256, 0, stream>>>(static_cast<const float*>(masks), S, out, out_stride);
  note_launch();
  return static_cast<int>(cudaGetLastError());
}

int launch_maxpool3x3s2(const __nv_bfloat16* x_hi, const __nv_bfloat16* x_lo, int n_images, int H, int W, int C,
                        __nv_bfloat16* y_hi, __nv_bfloat16* y_lo, cudaStream_t stream) {
  const int Ho = (H - 3) / 2 + 1, Wo = (W - 3) / 2 + 1;
  const long long total = static_cast<long long>(n_images) * Ho * Wo * (C / 8);
  maxpool3x3s2_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(x_hi, x_lo, n_images, H, W, C,
                                                                                      y_hi, y_lo);
  note_launch();
  return static_cast<int>(cudaGetLastError());
}

int launch_planes_to_f32(const __nv_bfloat16* hi, const __nv_bfloat16* lo, long long n, float* out,
                         cudaStream_t stream) {
  const long long n2 = n / 2;
  if (n2 == 0) return 0;
  planes_to_f32_kernel<<<static_cast<unsigned>((n2 + 255) / 256), 256, 0, stream>>>(hi, lo, n2, out);
  note_launch();
  return static_cast<int>(cudaGetLastError());
}

int launch_stem_pack(const void* images, int dtype, int n_images, __nv_bfloat16* p_hi, __nv_bfloat16* p_lo,
                     const float mean[3], const float stdv[3], int split, cudaStream_t stream, const void* masks) {
  const long long total = static_cast<long long>(n_images) * kStemPadH * (kStemPadW / 4);
  const int threads = 256;
  const long long blocks = (total + threads - 1) / threads;
  const float3 m = make_float3(mean[0], mean[1], mean[2]);
  const float3 s = make_float3(stdv[0], stdv[1], stdv[2]);
  if (dtype == 0) {
    stem_pack_kernel<uint8_t><<<static_cast<unsigned>(blocks), threads, 0, stream>>>(
        static_cast<const uint8_t*>(images), n_images, p_hi, p_lo, m, s, split, static_cast<const uint8_t*>(masks));
  } else {
    stem_pack_kernel<float><<<static_cast<unsigned>(blocks), threads, 0, stream>>>(
        static_cast<const float*>(images), n_images, p_hi, p_lo, m, s, split, static_cast<const float*>(masks));
  }
  note_launch();
  return static_cast<int>(cudaGetLastError());
}

int launch_mask_pyramid(const void* masks, int dtype, int n_images, float* wts, cudaStream_t stream) {
  if (dtype == 0) {
    mask_pyramid_kernel<uint8_t><<<n_images, 256, 0, stream>>>(static_cast<const uint8_t*>(masks), wts);
  } else {
    mask_pyramid_kernel<float><<<n_images, 256, 0, stream>>>(static_cast<const float*>(masks), wts);
  }
  note_launch();
  return static_cast<int>(cudaGetLastError());
}

int launch_masked_pool(const __nv_bfloat16* hi, const __nv_bfloat16* lo, const float* wts, int wts_stride,
                       int n_images, int P, int C, float* out, int out_stride, cudaStream_t stream) {
  dim3 grid(n_images, C / 64);
  masked_pool_kernel<<<grid, 256, 0, stream>>>(hi, lo, wts, wts_stride, P, C, out, out_stride);
  note_launch();
  return static_cast<int>(cudaGetLastError());
}

int launch_bn_relu_maxpool(const __nv_bfloat16* x_hi, const __nv_bfloat16* x_lo, const float* alpha,
                           const float* beta, int n_images, __nv_bfloat16* y_hi, __nv_bfloat16* y_lo,
                           cudaStream_t stream) {
  const int threads = 256;  // one CTA = 4 x 8 output pixels x 8 channel groups
  const long long blocks = static_cast<long long>(n_images) * (56 / 8) * (56 / 4);
  bn_relu_maxpool_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(x_hi, x_lo, alpha, beta, n_images,
                                                                               y_hi, y_lo);
  note_launch();
  return static_cast<int>(cudaGetLastError());
}

}  // namespace milan
