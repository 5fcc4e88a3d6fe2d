// Stage-1 exemplar statistics (reference: src/exemplars/compute.py:27-246 and the NetDissect pieces it runs,
// src/deps/netdissect/{runningstats,tally,imgviz,upsample}.py): HBM-bound scans over the activation maps of the
// network being described. The forward pass of that network is the caller's business (a library call).
//   tally_topk        per unit: spatial max of every image of the batch, merged into the running top-k
//                     (RunningTopK, runningstats.py:31-118; ties: the earlier dataset index wins)
//   tally_samples     exact regime of the quantile sketch (<= 8192 samples per unit, runningstats.py:295-386):
//                     every activation is kept, unit-major
//   tally_hist        beyond it: a deterministic 2^16-bin histogram of the order-preserving float key per unit
//                     (the reference switches to a randomised KLL sketch there; histograms add across GPUs)
//   quantile_exact    the reference's estimator on the kept samples (quantiles(), runningstats.py:557-580)
//   quantile_hist     the same read-out from the histogram, linear inside the selected bin
//   activation_masks  NetDissect's default upsampling grid + `> level` (imgviz.py:185-198, upsample.py:127-157)
#include "milan_b200.h"

#include "conv_gemm.h"  // note_launch

#include <cuda_runtime.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>

namespace {

__device__ __forceinline__ unsigned float_key(float x) {  // larger float <-> larger key
  const unsigned u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

constexpr int kTallyThreads = 256;
constexpr int kMaxTopK = 64;
constexpr int kMaxBatch = 1024;

// pooled[b][u] = max_p acts[b][u][p]: one group of TPR threads per (b, u) row (TPR = 8 / 32 / 256 by row length), so
// that short rows (7x7 maps) and long rows (112x112) both keep every lane loading; rows are contiguous in memory.
template <int TPR>
__global__ void __launch_bounds__(256) pooled_max_kernel(const float* __restrict__ acts, long long rows, int P,
                                                         float* __restrict__ pooled) {
  constexpr int kRowsPerCta = 256 / TPR;
  const long long row = static_cast<long long>(blockIdx.x) * kRowsPerCta + threadIdx.x / TPR;
  const int t = threadIdx.x % TPR;
  float m = -INFINITY;
  if (row < rows) {
    const float* src = acts + row * P;
    for (int p = t; p < P; p += TPR) m = fmaxf(m, __ldg(src + p));
  }
  if (TPR <= 32) {
#pragma unroll
    for (int o = TPR / 2; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (t == 0 && row < rows) pooled[row] = m;
  } else {
    __shared__ float red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
      if (row < rows) pooled[row] = m;
    }
  }
}

// One CTA per unit: merge the batch's pooled values (pooled: (B, U)) into the unit's sorted top-k.
__global__ void __launch_bounds__(kTallyThreads) tally_topk_kernel(const float* __restrict__ pooled, int B, int U,
                                                                   long long base_index, int k,
                                                                   float* __restrict__ top_vals,
                                                                   long long* __restrict__ top_ids) {
  __shared__ float val_s[kMaxBatch + kMaxTopK];
  __shared__ long long id_s[kMaxBatch + kMaxTopK];
  __shared__ float red_v[kTallyThreads / 32];
  __shared__ int red_i[kTallyThreads / 32];
  const int u = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = kTallyThreads / 32;
  for (int b = threadIdx.x; b < B; b += kTallyThreads) {
    val_s[b] = pooled[static_cast<long long>(b) * U + u];
    id_s[b] = base_index + b;
  }
  for (int j = threadIdx.x; j < k; j += kTallyThreads) {
    val_s[B + j] = top_vals[static_cast<long long>(u) * k + j];
    id_s[B + j] = top_ids[static_cast<long long>(u) * k + j];
  }
  __syncthreads();
  const int n = B + k;
  for (int r = 0; r < k; ++r) {
    // block arg-best: larger value, then smaller dataset index; empty slots (id < 0) lose
    float bv = -INFINITY;
    int bi = -1;
    for (int i = threadIdx.x; i < n; i += kTallyThreads) {
      if (id_s[i] < 0) continue;
      if (bi < 0 || val_s[i] > bv || (val_s[i] == bv && id_s[i] < id_s[bi])) { bv = val_s[i]; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && id_s[oi] < id_s[bi]))) { bv = ov; bi = oi; }
    }
    if (lane == 0) { red_v[warp] = bv; red_i[warp] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < nwarps; ++w) {
        const int oi = red_i[w];
        if (oi >= 0 && (bi < 0 || red_v[w] > bv || (red_v[w] == bv && id_s[oi] < id_s[bi]))) { bv = red_v[w]; bi = oi; }
      }
      top_vals[static_cast<long long>(u) * k + r] = bi >= 0 ? bv : -INFINITY;
      top_ids[static_cast<long long>(u) * k + r] = bi >= 0 ? id_s[bi] : -1;
      if (bi >= 0) id_s[bi] = -1;  // taken
    }
    __syncthreads();
  }
}

// dst[u][count + b*P + p] = acts[b][u][p]
__global__ void tally_samples_kernel(const float* __restrict__ acts, int U, int P, float* __restrict__ samples,
                                     long long capacity, long long count) {
  const int b = blockIdx.x, u = blockIdx.y;
  const float* src = acts + (static_cast<long long>(b) * U + u) * P;
  float* dst = samples + static_cast<long long>(u) * capacity + count + static_cast<long long>(b) * P;
  for (int p = threadIdx.x; p < P; p += blockDim.x) dst[p] = src[p];
}

// Histogram of the upper 16 key bits per unit. The 65536 bins of a unit (256 KB) do not fit in shared memory, but the
// activations of one unit live in a narrow band of them: bin = sign | exponent | 7 mantissa bits, so 64 binades around
// +-1.0 are 8192 bins each. A CTA owns one unit (and a slice of the batch), counts those two windows in shared memory
// (64 KB) and flushes the non-empty bins once; +-0 and anything outside the windows (|x| < 2^-32 or >= 2^32) go to
// global memory directly. Lanes of a warp that hit the same bin add once (post-ReLU maps are mostly one value).
constexpr int kHistWindow = 8192;
constexpr unsigned kHistPosBase = 0xBF80u - kHistWindow / 2;  // bin of +1.0 = 0xBF80
constexpr unsigned kHistNegBase = 0x407Fu - kHistWindow / 2;  // bin of -1.0 = 0x407F
__global__ void __launch_bounds__(256) tally_hist_kernel(const float* __restrict__ acts, int B, int U, int P,
                                                         int b_per_cta, unsigned* __restrict__ hist) {
  extern __shared__ unsigned win[];  // [2][kHistWindow]
  const int u = blockIdx.x;
  const int b0 = blockIdx.y * b_per_cta;
  const int b1 = min(B, b0 + b_per_cta);
  unsigned* h = hist + static_cast<long long>(u) * 65536;
  for (int i = threadIdx.x; i < 2 * kHistWindow; i += blockDim.x) win[i] = 0;
  __syncthreads();
  const long long total = static_cast<long long>(b1 - b0) * P;
  const long long rounds = (total + blockDim.x - 1) / blockDim.x;
  for (long long r = 0; r < rounds; ++r) {
    const long long i = r * blockDim.x + threadIdx.x;
    const bool live = i < total;
    unsigned bin = 0xFFFFFFFFu;
    if (live) {
      const int b = b0 + static_cast<int>(i / P);
      const int p = static_cast<int>(i - static_cast<long long>(b - b0) * P);
      bin = float_key(__ldg(acts + (static_cast<long long>(b) * U + u) * P + p)) >> 16;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, bin);
    if (live && (threadIdx.x & 31) == __ffs(peers) - 1) {
      const unsigned n = static_cast<unsigned>(__popc(peers));
      const unsigned dp = bin - kHistPosBase, dn = bin - kHistNegBase;  // unsigned: out of window -> huge
      if (dp < static_cast<unsigned>(kHistWindow)) atomicAdd(&win[dp], n);
      else if (dn < static_cast<unsigned>(kHistWindow)) atomicAdd(&win[kHistWindow + dn], n);
      else atomicAdd(&h[bin], n);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * kHistWindow; i += blockDim.x) {
    const unsigned n = win[i];
    if (n != 0) atomicAdd(&h[(i < kHistWindow ? kHistPosBase + i : kHistNegBase + (i - kHistWindow))], n);
  }
}

// The reference's read-out: numpy.interp(q, (cumsum(w) - w/2)/n, [min, sorted..., max]) with w = [0, 1...1, 0]
// (float32 abscissae, float64 interpolation). One CTA per unit, bitonic sort of n <= 8192 samples in shared memory.
__global__ void __launch_bounds__(1024) quantile_exact_kernel(const float* __restrict__ samples, long long capacity,
                                                             int n, float q, float* __restrict__ levels) {
  extern __shared__ float s[];
  const int u = blockIdx.x;
  int m = 1;
  while (m < n) m <<= 1;
  for (int i = threadIdx.x; i < m; i += blockDim.x) s[i] = i < n ? samples[static_cast<long long>(u) * capacity + i] : INFINITY;
  __syncthreads();
  for (int size = 2; size <= m; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const int j = i ^ stride;
        if (j > i) {
          const bool up = (i & size) == 0;
          const float a = s[i], b = s[j];
          if ((a > b) == up) { s[i] = b; s[j] = a; }
        }
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    const double qq = static_cast<double>(q);
    const float fn = static_cast<float>(n);
    auto xs = [&](int j) -> double {  // j = 0 .. n+1
      if (j == 0) return 0.0;
      if (j == n + 1) return static_cast<double>(fn / fn);
      return static_cast<double>((static_cast<float>(j) - 0.5f) / fn);
    };
    auto ys = [&](int j) -> double { return static_cast<double>(j == 0 ? s[0] : (j == n + 1 ? s[n - 1] : s[j - 1])); };
    double out;
    if (qq <= xs(0)) {
      out = ys(0);
    } else if (qq >= xs(n + 1)) {
      out = ys(n + 1);
    } else {
      int lo = 0, hi = n + 1;  // xs(lo) <= q < xs(hi)
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (xs(mid) <= qq) lo = mid; else hi = mid;
      }
      const double slope = (ys(lo + 1) - ys(lo)) / (xs(lo + 1) - xs(lo));
      out = slope * (qq - xs(lo)) + ys(lo);
    }
    levels[u] = static_cast<float>(out);
  }
}

// Same estimator from a histogram: the sample of (fractional) rank q*n - 0.5 lies in the bin where the cumulative
// count crosses it; values are taken uniformly spread inside the bin.
__global__ void quantile_hist_kernel(const unsigned* __restrict__ hist, long long n, float q, float* __restrict__ levels) {
  __shared__ unsigned long long part[256];
  const int u = blockIdx.x;
  const unsigned* h = hist + static_cast<long long>(u) * 65536;
  unsigned long long local = 0;
  for (int i = 0; i < 256; ++i) local += h[threadIdx.x * 256 + i];
  part[threadIdx.x] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    const double target = static_cast<double>(q) * static_cast<double>(n) - 0.5;
    unsigned long long cum = 0;
    int seg = 0;
    while (seg < 255 && static_cast<double>(cum + part[seg]) <= target) cum += part[seg++];
    int bin = seg * 256;
    while (bin < seg * 256 + 255 && static_cast<double>(cum + h[bin]) <= target) cum += h[bin++];
    const double inside = h[bin] > 0 ? (target - static_cast<double>(cum)) / static_cast<double>(h[bin]) : 0.0;
    const double lo = key_float(static_cast<unsigned>(bin) << 16);
    const double hi = key_float((static_cast<unsigned>(bin) << 16) | 0xFFFFu);
    const double a = fmin(lo, hi), b = fmax(lo, hi);
    levels[u] = static_cast<float>(a + fmin(fmax(inside, 0.0), 1.0) * (b - a));
  }
}

// masks[i][y][x] = grid_sample(maps[i], default NetDissect grid)(y, x) > levels[i]   (bilinear, zeros, align_corners)
__device__ __forceinline__ uint8_t mask_pixel(const float* __restrict__ m, float level, int H, int W, int y, int x,
                                              float oy, float cy, float ox, float cx) {
  // upsample_grid: g = (t - o) * (2 / (s * max(1, size - 1))) - 1; grid_sample: src = ((g + 1) / 2) * (size - 1)
  const float gy = __fsub_rn(__fmul_rn(__fsub_rn(static_cast<float>(y), oy), cy), 1.0f);
  const float gx = __fsub_rn(__fmul_rn(__fsub_rn(static_cast<float>(x), ox), cx), 1.0f);
  const float fy = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.0f), 2.0f), static_cast<float>(H - 1));
  const float fx = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.0f), 2.0f), static_cast<float>(W - 1));
  const float y0 = floorf(fy), x0 = floorf(fx);
  const float y1 = y0 + 1.0f, x1 = x0 + 1.0f;
  const float w_nw = __fmul_rn(x1 - fx, y1 - fy), w_ne = __fmul_rn(fx - x0, y1 - fy);
  const float w_sw = __fmul_rn(x1 - fx, fy - y0), w_se = __fmul_rn(fx - x0, fy - y0);
  const int iy0 = static_cast<int>(y0), ix0 = static_cast<int>(x0);
  auto at = [&](int yy, int xx) -> float { return (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(m + yy * W + xx) : 0.0f; };
  float v = 0.0f;
  v = __fadd_rn(v, __fmul_rn(at(iy0, ix0), w_nw));
  v = __fadd_rn(v, __fmul_rn(at(iy0, ix0 + 1), w_ne));
  v = __fadd_rn(v, __fmul_rn(at(iy0 + 1, ix0), w_sw));
  v = __fadd_rn(v, __fmul_rn(at(iy0 + 1, ix0 + 1), w_se));
  return v > level ? 1 : 0;
}
// One thread = 16 consecutive pixels of a row (one 16-byte store) when S % 16 == 0, else one pixel.
template <int PPT>
__global__ void activation_masks_kernel(const float* __restrict__ maps, const float* __restrict__ levels, int n, int H,
                                        int W, int S, float oy, float cy, float ox, float cx,
                                        uint8_t* __restrict__ masks) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int groups = S / PPT;
  if (idx >= static_cast<long long>(n) * S * groups) return;
  const int xg = idx % groups, y = (idx / groups) % S, i = idx / (static_cast<long long>(groups) * S);
  const float* m = maps + static_cast<long long>(i) * H * W;
  const float level = levels[i];
  alignas(16) uint8_t out[PPT];
#pragma unroll
  for (int j = 0; j < PPT; ++j) out[j] = mask_pixel(m, level, H, W, y, xg * PPT + j, oy, cy, ox, cx);
  uint8_t* dst = masks + (static_cast<long long>(i) * S + y) * S + xg * PPT;
  if (PPT == 16) *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(out);
  else dst[0] = out[0];
}

int done() {
  milan::note_launch();
  return static_cast<int>(cudaGetLastError());
}

// The stage-1 entry points take bare device pointers: run on the device that owns them (not whatever device is
// current in the calling thread) and restore the caller's device on the way out.
struct PointerDeviceGuard {
  int device = 0, prev = -1;
  cudaError_t err = cudaSuccess;
  explicit PointerDeviceGuard(const void* device_ptr) {
    cudaPointerAttributes attr{};
    err = cudaPointerGetAttributes(&attr, device_ptr);
    if (err != cudaSuccess) return;
    if (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged) { err = cudaErrorInvalidDevicePointer; return; }
    device = attr.device;
    err = cudaGetDevice(&prev);
    if (err != cudaSuccess) return;
    if (prev == device) prev = -1;
    else err = cudaSetDevice(device);
  }
  ~PointerDeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
  PointerDeviceGuard(const PointerDeviceGuard&) = delete;
  PointerDeviceGuard& operator=(const PointerDeviceGuard&) = delete;
};

}  // namespace

extern "C" {

int milan_tally_topk(const float* d_acts, int32_t B, int32_t U, int32_t P, int64_t base_index, int32_t k,
                     float* d_pooled_scratch, float* d_top_vals, int64_t* d_top_ids, void* stream) {
  if (B <= 0 || U <= 0) return 0;
  if (k < 1 || k > kMaxTopK || B > kMaxBatch || P < 1) return static_cast<int>(cudaErrorInvalidValue);
  PointerDeviceGuard guard(d_acts);
  if (guard.err != cudaSuccess) return static_cast<int>(guard.err);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* pooled = d_pooled_scratch;  // (B, U) spatial maxima
  const long long rows = static_cast<long long>(B) * U;
  if (pooled == nullptr) return static_cast<int>(cudaErrorInvalidValue);
  if (P <= 16) pooled_max_kernel<8><<<static_cast<unsigned>((rows + 31) / 32), 256, 0, st>>>(d_acts, rows, P, pooled);
  else if (P <= 1024) pooled_max_kernel<32><<<static_cast<unsigned>((rows + 7) / 8), 256, 0, st>>>(d_acts, rows, P, pooled);
  else pooled_max_kernel<256><<<static_cast<unsigned>(rows), 256, 0, st>>>(d_acts, rows, P, pooled);
  milan::note_launch();
  tally_topk_kernel<<<U, kTallyThreads, 0, st>>>(pooled, B, U, base_index, k, d_top_vals,
                                                 reinterpret_cast<long long*>(d_top_ids));
  return done();
}

int milan_tally_samples(const float* d_acts, int32_t B, int32_t U, int32_t P, float* d_samples, int64_t capacity,
                        int64_t count, void* stream) {
  if (B <= 0 || U <= 0) return 0;
  if (count + static_cast<int64_t>(B) * P > capacity) return static_cast<int>(cudaErrorInvalidValue);
  PointerDeviceGuard guard(d_acts);
  if (guard.err != cudaSuccess) return static_cast<int>(guard.err);
  tally_samples_kernel<<<dim3(B, U), 128, 0, static_cast<cudaStream_t>(stream)>>>(d_acts, U, P, d_samples, capacity, count);
  return done();
}

int milan_tally_hist(const float* d_acts, int32_t B, int32_t U, int32_t P, uint32_t* d_hist, void* stream) {
  if (B <= 0 || U <= 0) return 0;
  PointerDeviceGuard guard(d_acts);
  if (guard.err != cudaSuccess) return static_cast<int>(guard.err);
  constexpr int smem = 2 * kHistWindow * static_cast<int>(sizeof(unsigned));
  cudaError_t e = cudaFuncSetAttribute(tally_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return static_cast<int>(e);
  // one unit per CTA column; the batch is sliced so that the grid has a few CTAs per SM but every CTA still
  // amortises its 64 KB window flush over >= 64 K activations where the batch allows
  int sm_count = 148;
  cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, guard.device);
  int slices = (3 * sm_count + U - 1) / U;
  const long long per_image = P;
  const int max_slices = static_cast<int>(std::max<long long>(1, static_cast<long long>(B) * per_image / 65536));
  slices = std::max(1, std::min(std::min(slices, max_slices), B));
  const int b_per_cta = (B + slices - 1) / slices;
  slices = (B + b_per_cta - 1) / b_per_cta;
  tally_hist_kernel<<<dim3(U, slices), 256, smem, static_cast<cudaStream_t>(stream)>>>(d_acts, B, U, P, b_per_cta, d_hist);
  return done();
}

int milan_quantile_exact(const float* d_samples, int32_t U, int64_t capacity, int64_t n, float q, float* d_levels,
                         void* stream) {
  if (U <= 0) return 0;
  if (n < 1 || n > 8192 || n > capacity) return static_cast<int>(cudaErrorInvalidValue);
  PointerDeviceGuard guard(d_samples);
  if (guard.err != cudaSuccess) return static_cast<int>(guard.err);
  int m = 1;
  while (m < n) m <<= 1;
  quantile_exact_kernel<<<U, 1024, static_cast<size_t>(m) * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      d_samples, capacity, static_cast<int>(n), q, d_levels);
  return done();
}

int milan_quantile_hist(const uint32_t* d_hist, int32_t U, int64_t n, float q, float* d_levels, void* stream) {
  if (U <= 0) return 0;
  if (n < 1) return static_cast<int>(cudaErrorInvalidValue);
  PointerDeviceGuard guard(d_hist);
  if (guard.err != cudaSuccess) return static_cast<int>(guard.err);
  quantile_hist_kernel<<<U, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_hist, n, q, d_levels);
  return done();
}

int milan_activation_masks(const float* d_maps, const float* d_levels, int32_t n, int32_t H, int32_t W, int32_t S,
                           uint8_t* d_masks, void* stream) {
  if (n <= 0) return 0;
  if (H < 1 || W < 1 || S < 1) return static_cast<int>(cudaErrorInvalidValue);
  PointerDeviceGuard guard(d_maps);
  if (guard.err != cudaSuccess) return static_cast<int>(guard.err);
  // upsample_grid with scale_offset=None: scale = S / size, offset = 0.5 * scale - 0.5 (Python floats), then
  // (arange - offset) * (2 / (scale * max(1, size - 1))) - 1 on float32 tensors
  const double sy = static_cast<double>(S) / H, sx = static_cast<double>(S) / W;
  const float oy = static_cast<float>(0.5 * sy - 0.5), ox = static_cast<float>(0.5 * sx - 0.5);
  const float cy = static_cast<float>(2.0 / (sy * (H > 1 ? H - 1 : 1))), cx = static_cast<float>(2.0 / (sx * (W > 1 ? W - 1 : 1)));
  const bool wide = S % 16 == 0 && reinterpret_cast<uintptr_t>(d_masks) % 16 == 0;
  const long long total = static_cast<long long>(n) * S * (wide ? S / 16 : S);
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  if (wide)
    activation_masks_kernel<16><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_maps, d_levels, n, H, W, S, oy, cy,
                                                                                      ox, cx, d_masks);
  else
    activation_masks_kernel<1><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_maps, d_levels, n, H, W, S, oy, cy,
                                                                                     ox, cx, d_masks);
  return done();
}

}  // extern "C"
