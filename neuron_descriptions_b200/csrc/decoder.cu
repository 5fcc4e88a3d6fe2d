// Non-GEMM kernels of the decoder / beam search / LM rerank (see decoder.h). All arithmetic is fp32; operands
// that feed a tensor-core GEMM are emitted as (hi, lo) bf16 pairs.
#include "decoder.h"
#include "conv_gemm.h"
#include "ptx.cuh"

#include <cfloat>
#include <cmath>

namespace milan {

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// Decoder / LM GEMM operands are stored as IEEE fp16 (hi, lo) pairs: 22 mantissa bits, ~30x tighter than the
// bf16 pair the encoder uses; safe because every decoder operand is bounded (|x| << 65504). The planes keep the
// 16-bit `__nv_bfloat16*` pointer type of the shared GEMM plumbing; only the bit patterns differ.
__device__ __forceinline__ void store_split2(__nv_bfloat16* hi, __nv_bfloat16* lo, long long off, float v0,
                                             float v1) {
  uint32_t h, l;
  split_fp16x2(v0, v1, h, l);
  *reinterpret_cast<uint32_t*>(hi + off) = h;
  if (lo != nullptr) *reinterpret_cast<uint32_t*>(lo + off) = l;
}

// ------------------------------------------------------------------ split / mean / init
__global__ void split_rows_kernel(const float* __restrict__ src, long long src_pitch, __nv_bfloat16* dst_hi,
                                  __nv_bfloat16* dst_lo, long long dst_pitch, int M, int K2) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(M) * K2) return;
  const int k2 = idx % K2;
  const long long m = idx / K2;
  const float2 v = *reinterpret_cast<const float2*>(src + m * src_pitch + 2 * k2);
  store_split2(dst_hi, dst_lo, m * dst_pitch + 2 * k2, v.x, v.y);
}

__global__ void mean_keys_kernel(const float* __restrict__ f, int B, int n_keys, int F, __nv_bfloat16* dst_hi,
                                 __nv_bfloat16* dst_lo) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int F2 = F / 2;
  if (idx >= static_cast<long long>(B) * F2) return;
  const int j2 = idx % F2;
  const long long b = idx / F2;
  float s0 = 0.f, s1 = 0.f;
  for (int k = 0; k < n_keys; ++k) {
    const float2 v = *reinterpret_cast<const float2*>(f + (b * n_keys + k) * F + 2 * j2);
    s0 += v.x;
    s1 += v.y;
  }
  store_split2(dst_hi, dst_lo, b * F + 2 * j2, s0 / n_keys, s1 / n_keys);
}

__global__ void init_finish_kernel(const float* __restrict__ pre, int B, int H, float* h, float* c,
                                   __nv_bfloat16* h_hi, __nv_bfloat16* h_lo, long long h_pitch) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int H2 = H / 2;
  if (idx >= static_cast<long long>(B) * H2) return;
  const int j2 = idx % H2;
  const long long b = idx / H2;
  const float2 ph = *reinterpret_cast<const float2*>(pre + b * 2 * H + 2 * j2);
  const float2 pc = *reinterpret_cast<const float2*>(pre + b * 2 * H + H + 2 * j2);
  const float h0 = tanhf(ph.x), h1 = tanhf(ph.y);
  if (h != nullptr) *reinterpret_cast<float2*>(h + b * H + 2 * j2) = make_float2(h0, h1);
  *reinterpret_cast<float2*>(c + b * H + 2 * j2) = make_float2(tanhf(pc.x), tanhf(pc.y));
  if (h_hi != nullptr) store_split2(h_hi, h_lo, b * h_pitch + 2 * j2, h0, h1);
}

// ------------------------------------------------------------------ attention + gating + LSTM input assembly
// Kernel 1 (one CTA per row): attention scores + softmax over keys (Attention.forward, decoders.py:57-73) and the
// token embedding (decoders.py:618). Kernel 2 (one CTA per feature set x column slice): attenuate + gate
// (decoders.py:613-615); the feature tile is read once per neuron and reused by all of its beam rows.
__global__ void __launch_bounds__(256) attn_scores_kernel(const AttendArgs a, float* __restrict__ attn_ws) {
  if (a.skip != nullptr && *a.skip) return;
  extern __shared__ float sm[];
  float* q_s = sm;                 // [A]
  float* sc_s = sm + a.A;          // [n_keys]
  const int r = blockIdx.x;
  const int fidx = r / a.rows_per_feature;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* qrow = a.qg + static_cast<long long>(r) * a.qg_pitch;
  for (int i = threadIdx.x; i < a.A; i += blockDim.x) q_s[i] = qrow[i];
  __syncthreads();
  for (int k = warp; k < a.n_keys; k += 8) {
    const float* khr = a.kh + (static_cast<long long>(fidx) * a.n_keys + k) * a.A;
    float s = 0.f;
    for (int i = lane; i < a.A; i += 32) s += a.w_o[i] * tanhf(q_s[i] + khr[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) sc_s[k] = s + a.b_o;
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int k = 0; k < a.n_keys; ++k) mx = fmaxf(mx, sc_s[k]);
  float den = 0.f;
  for (int k = 0; k < a.n_keys; ++k) den += expf(sc_s[k] - mx);
  for (int k = threadIdx.x; k < a.n_keys; k += blockDim.x) {
    const float w = expf(sc_s[k] - mx) / den;
    attn_ws[static_cast<long long>(r) * a.n_keys + k] = w;
    if (a.attn_out != nullptr) a.attn_out[static_cast<long long>(r) * a.attn_pitch + k] = w;
  }
  const long long xoff = static_cast<long long>(r) * a.x_pitch;
  const float* erow = a.embedding + a.tokens[r] * a.E;
  for (int e2 = threadIdx.x; e2 < a.E / 2; e2 += blockDim.x) {
    const float2 v = *reinterpret_cast<const float2*>(erow + 2 * e2);
    store_split2(a.x_hi, a.x_lo, xoff + 2 * e2, v.x, v.y);
  }
}

constexpr int kApplyKeys = 16;  // keys held in registers per pass
__global__ void __launch_bounds__(256) attn_apply_kernel(const AttendArgs a, const float* __restrict__ attn_ws) {
  if (a.skip != nullptr && *a.skip) return;
  extern __shared__ float w_s[];  // [rows_per_feature][n_keys]
  const int fidx = blockIdx.x;
  const int j2 = blockIdx.y * blockDim.x + threadIdx.x;
  const int rpf = a.rows_per_feature;
  const long long row0 = static_cast<long long>(fidx) * rpf;
  for (int i = threadIdx.x; i < rpf * a.n_keys; i += blockDim.x) w_s[i] = attn_ws[row0 * a.n_keys + i];
  __syncthreads();
  if (j2 >= a.F / 2) return;
  const float* fb = a.features + static_cast<long long>(fidx) * a.n_keys * a.F + 2 * j2;
  for (int k0 = 0; k0 < a.n_keys; k0 += kApplyKeys) {
    float2 f[kApplyKeys];
#pragma unroll
    for (int k = 0; k < kApplyKeys; ++k)
      f[k] = (k0 + k < a.n_keys) ? __ldg(reinterpret_cast<const float2*>(fb + static_cast<long long>(k0 + k) * a.F))
                                 : make_float2(0.f, 0.f);
    if (a.n_keys <= kApplyKeys) {
      // common case (k = 15 exemplars): a single pass, results go straight to the LSTM operand
      for (int r = 0; r < rpf; ++r) {
        const float* w = w_s + r * a.n_keys;
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int k = 0; k < kApplyKeys; ++k) {
          if (k < a.n_keys) {
            s0 += w[k] * f[k].x;
            s1 += w[k] * f[k].y;
          }
        }
        const float2 g = *reinterpret_cast<const float2*>(a.qg + (row0 + r) * a.qg_pitch + a.A + 2 * j2);
        store_split2(a.x_hi, a.x_lo, (row0 + r) * a.x_pitch + a.E + 2 * j2, s0 * sigmoidf_(g.x), s1 * sigmoidf_(g.y));
      }
    } else {
      // many keys (e.g. a spatial encoder's k*49): accumulate per row across passes in the gate buffer
      for (int r = 0; r < rpf; ++r) {
        const float* w = w_s + r * a.n_keys + k0;
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int k = 0; k < kApplyKeys; ++k) {
          if (k0 + k < a.n_keys) {
            s0 += w[k] * f[k].x;
            s1 += w[k] * f[k].y;
          }
        }
        float* acc = const_cast<float*>(a.acc_ws) + (row0 + r) * a.F + 2 * j2;
        if (k0 > 0) { s0 += acc[0]; s1 += acc[1]; }
        if (k0 + kApplyKeys < a.n_keys) {
          acc[0] = s0; acc[1] = s1;
        } else {
          const float2 g = *reinterpret_cast<const float2*>(a.qg + (row0 + r) * a.qg_pitch + a.A + 2 * j2);
          store_split2(a.x_hi, a.x_lo, (row0 + r) * a.x_pitch + a.E + 2 * j2, s0 * sigmoidf_(g.x),
                       s1 * sigmoidf_(g.y));
        }
      }
    }
  }
}

// ------------------------------------------------------------------ LSTM pointwise
__global__ void lstm_point_kernel(const LstmPointArgs a) {
  if (a.skip != nullptr && *a.skip) return;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int H2 = a.H / 2;
  if (idx >= static_cast<long long>(a.R) * H2) return;
  const int j = 2 * (idx % H2);
  const long long r = idx / H2;
  const float* g = a.gates + r * 4 * a.H;
  const float2 gi = *reinterpret_cast<const float2*>(g + j);
  const float2 gf = *reinterpret_cast<const float2*>(g + a.H + j);
  const float2 gg = *reinterpret_cast<const float2*>(g + 2 * a.H + j);
  const float2 go = *reinterpret_cast<const float2*>(g + 3 * a.H + j);
  const float2 c = *reinterpret_cast<const float2*>(a.c_in + r * a.H + j);
  const float c0 = sigmoidf_(gf.x) * c.x + sigmoidf_(gi.x) * tanhf(gg.x);
  const float c1 = sigmoidf_(gf.y) * c.y + sigmoidf_(gi.y) * tanhf(gg.y);
  const float h0 = sigmoidf_(go.x) * tanhf(c0);
  const float h1 = sigmoidf_(go.y) * tanhf(c1);
  *reinterpret_cast<float2*>(a.c_out + r * a.H + j) = make_float2(c0, c1);
  if (a.h_out != nullptr) *reinterpret_cast<float2*>(a.h_out + r * a.H + j) = make_float2(h0, h1);
#pragma unroll
  for (int d = 0; d < 2; ++d)
    if (a.h_hi[d] != nullptr) store_split2(a.h_hi[d], a.h_lo[d], r * a.h_pitch[d] + j, h0, h1);
}

__global__ void embed_rows_kernel(const float* __restrict__ table, const long long* __restrict__ tokens, int M,
                                  int E2, __nv_bfloat16* dst_hi, __nv_bfloat16* dst_lo, long long dst_pitch,
                                  const int* __restrict__ skip) {
  if (skip != nullptr && *skip) return;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(M) * E2) return;
  const int e2 = idx % E2;
  const long long m = idx / E2;
  const float2 v = *reinterpret_cast<const float2*>(table + tokens[m] * (2 * E2) + 2 * e2);
  store_split2(dst_hi, dst_lo, m * dst_pitch + 2 * e2, v.x, v.y);
}

// ------------------------------------------------------------------ block argmax helper
struct ValIdx {
  float v;
  int i;
};
__device__ __forceinline__ bool better(const ValIdx& a, const ValIdx& b) {  // a strictly preferred over b
  return a.v > b.v || (a.v == b.v && a.i < b.i);
}
__device__ __forceinline__ ValIdx warp_argmax(ValIdx x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ValIdx y;
    y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
    y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
    if (better(y, x)) x = y;
  }
  return x;
}
// All threads get the block-wide best. red must hold blockDim/32 entries. Two barriers.
__device__ __forceinline__ ValIdx block_argmax(ValIdx x, ValIdx* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  x = warp_argmax(x);
  __syncthreads();
  if (lane == 0) red[warp] = x;
  __syncthreads();
  ValIdx best = red[0];
  const int nw = blockDim.x >> 5;
  for (int w = 1; w < nw; ++w)
    if (better(red[w], best)) best = red[w];
  return best;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = red[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) t = fmaxf(t, red[w]);
  return t;
}
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w];
  return t;
}

// order-preserving float -> uint key (larger float <-> larger key)
__device__ __forceinline__ unsigned float_key(float x) {
  const unsigned u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// log_softmax of one row into shared memory (decoders.py:621, :624-630). The row is read from global memory once
// (128-bit loads; rows are 16-byte aligned because ld % 4 == 0) and normalised in place; with `subtract` the
// row's log-softmax is scaled and subtracted from pred_s instead (MI decoding), streaming from global.
// Returns this thread's maximum of the values it wrote (used by the beam top-k prefilter).
__device__ __forceinline__ float row_log_softmax(const float* __restrict__ x, int V, float* pred_s, float* red,
                                                 float scale_sub, bool subtract) {
  float tmax = -INFINITY;
  if (!subtract) {
    float mx = -INFINITY;
    const int V4 = V >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (int v = threadIdx.x; v < V4; v += blockDim.x) {
      const float4 q = __ldg(x4 + v);
      *reinterpret_cast<float4*>(pred_s + 4 * v) = q;
      mx = fmaxf(fmaxf(mx, fmaxf(q.x, q.y)), fmaxf(q.z, q.w));
    }
    for (int v = 4 * V4 + threadIdx.x; v < V; v += blockDim.x) {
      const float q = x[v];
      pred_s[v] = q;
      mx = fmaxf(mx, q);
    }
    mx = block_max(mx, red);  // barriers inside also publish pred_s
    float s = 0.f;
    for (int v = threadIdx.x; v < V; v += blockDim.x) s += expf(pred_s[v] - mx);
    s = block_sum(s, red);
    const float lse = logf(s);
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
      const float y = (pred_s[v] - mx) - lse;
      pred_s[v] = y;
      tmax = fmaxf(tmax, y);
    }
  } else {
    float mx = -INFINITY;
    for (int v = threadIdx.x; v < V; v += blockDim.x) mx = fmaxf(mx, x[v]);
    mx = block_max(mx, red);
    float s = 0.f;
    for (int v = threadIdx.x; v < V; v += blockDim.x) s += expf(x[v] - mx);
    s = block_sum(s, red);
    const float lse = logf(s);
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
      const float y = pred_s[v] - scale_sub * ((x[v] - mx) - lse);
      pred_s[v] = y;
      tmax = fmaxf(tmax, y);
    }
  }
  __syncthreads();
  return tmax;
}

constexpr int kPrefilterCap = 1024;
__global__ void __launch_bounds__(256) row_kernel(const RowArgs a) {
  if (a.skip != nullptr && *a.skip) return;
  extern __shared__ float pred_s[];  // [V]
  __shared__ float red[8];
  __shared__ ValIdx redvi[8];
  const int r = blockIdx.x;
  float tmax = row_log_softmax(a.logits + static_cast<long long>(r) * a.ld, a.V, pred_s, red, 0.f, false);
  if (a.logits_lm != nullptr)
    tmax = row_log_softmax(a.logits_lm + static_cast<long long>(r) * a.ld, a.V, pred_s, red, a.temperature, true);
  if (a.pred_out != nullptr) {
    float* po = a.pred_out + static_cast<long long>(r) * a.pred_pitch;
    for (int v = threadIdx.x; v < a.V; v += blockDim.x) po[v] = pred_s[v];
  }
  if (a.beam == 0) {
    // greedy / forced (decoders.py:444-463): no stop handling, score accumulates every step
    long long next;
    if (a.forced != nullptr) {
      next = a.forced[r];
    } else {
      ValIdx b{-INFINITY, 0x7fffffff};
      for (int v = threadIdx.x; v < a.V; v += blockDim.x) {
        const ValIdx c{pred_s[v], v};
        if (better(c, b)) b = c;
      }
      b = block_argmax(b, redvi);
      next = b.i;
    }
    if (threadIdx.x == 0) {
      if (a.next_tokens != nullptr) a.next_tokens[r] = next;
      if (a.scores != nullptr) a.scores[r] += pred_s[next];
    }
    return;
  }
  // beam candidates
  const float lp = a.last_lp != nullptr ? a.last_lp[r] : 0.0f;
  float* cv = a.cand_val + static_cast<long long>(r) * a.beam;
  int* cc = a.cand_cls + static_cast<long long>(r) * a.beam;
  if (a.last_tokens[r] == a.stop_index) {
    // finished beam: only <stop> at cost 0 survives (log_probs_after_end); the other per-node candidates carry
    // min_value_of_dtype in allennlp and are never selected while >= beam finite candidates exist.
    for (int j = threadIdx.x; j < a.beam; j += blockDim.x) {
      cv[j] = j == 0 ? lp + 0.0f : -INFINITY;
      cc[j] = static_cast<int>(a.stop_index);
    }
    return;
  }
  // Exact top-`beam`, fast path: the beam-th largest of 64 thread-group maxima is a lower bound of the
  // beam-th largest value of the row (beam <= kMaxBeam = 64), so only the (typically ~beam) elements >= that bound can be selected;
  // they are compacted and ranked exactly (ties: lower class index first, like the radix path below).
  {
    __shared__ float gmax_s[64];
    __shared__ float tau_s;
    __shared__ int n_cand;
    __shared__ float pval[kPrefilterCap];
    __shared__ int pidx[kPrefilterCap];
    // maxima of 64 groups of 4 threads; the beam-th largest of them bounds the beam-th largest value from below
    float gmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 1));
    gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, 2));
    if ((threadIdx.x & 3) == 0) gmax_s[threadIdx.x >> 2] = gmax;
    if (threadIdx.x == 0) n_cand = 0;
    __syncthreads();
    if (threadIdx.x < 64) {
      const ValIdx me{gmax_s[threadIdx.x], static_cast<int>(threadIdx.x)};
      int rank = 0;
#pragma unroll 8
      for (int j = 0; j < 64; ++j) rank += better(ValIdx{gmax_s[j], j}, me) ? 1 : 0;
      if (rank == a.beam - 1) tau_s = me.v;
    }
    __syncthreads();
    const float tau = tau_s;
    for (int v = threadIdx.x; v < a.V; v += blockDim.x) {
      const float x = pred_s[v];
      if (x >= tau) {
        const int pos = atomicAdd(&n_cand, 1);
        if (pos < kPrefilterCap) { pval[pos] = x; pidx[pos] = v; }
      }
    }
    __syncthreads();
    const int nc = n_cand;
    if (nc <= kPrefilterCap) {  // uniform
      for (int i = threadIdx.x; i < nc; i += blockDim.x) {
        const ValIdx me{pval[i], pidx[i]};
        int rank = 0;
        for (int j = 0; j < nc; ++j) rank += better(ValIdx{pval[j], pidx[j]}, me) ? 1 : 0;
        if (rank < a.beam) {
          cv[rank] = me.v + lp;
          cc[rank] = me.i;
        }
      }
      return;
    }
  }
  // Fallback (more than kPrefilterCap values tie with / exceed the bound, e.g. masses of equal or -inf logits):
  // exact top-`beam` by MSB radix select on order-preserving keys (4 passes of 8 bits), then rank the survivors.
  __shared__ unsigned hist[256];
  __shared__ unsigned warp_tot[8];
  __shared__ unsigned sel_digit, sel_need;
  __shared__ int n_gt, n_eq;
  __shared__ float cval[kMaxBeam];
  __shared__ int cidx[kMaxBeam];
  __shared__ int eqidx[kMaxBeam];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned prefix = 0, need = static_cast<unsigned>(a.beam);
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    hist[threadIdx.x] = 0;
    __syncthreads();
    for (int v = threadIdx.x; v < a.V; v += blockDim.x) {
      const unsigned k = float_key(pred_s[v]);
      if (pass == 0 || (k >> (shift + 8)) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1u);
    }
    __syncthreads();
    // inclusive suffix sum over digits (thread t <-> digit t)
    const unsigned own = hist[threadIdx.x];
    unsigned suf = own;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_down_sync(0xffffffffu, suf, o);
      if (lane + o < 32) suf += t;
    }
    if (lane == 0) warp_tot[warp] = suf;
    __syncthreads();
    for (int w = warp + 1; w < 8; ++w) suf += warp_tot[w];
    const unsigned excl = suf - own;
    if (excl < need && suf >= need) {
      sel_digit = threadIdx.x;
      sel_need = need - excl;
    }
    __syncthreads();
    prefix = (prefix << 8) | sel_digit;
    need = sel_need;
  }
  // prefix = key of the beam-th largest value; `need` of the elements equal to it are taken (lowest indices first)
  if (threadIdx.x == 0) { n_gt = 0; n_eq = 0; }
  __syncthreads();
  for (int v = threadIdx.x; v < a.V; v += blockDim.x) {
    const float x = pred_s[v];
    const unsigned k = float_key(x);
    if (k > prefix) {
      const int pos = atomicAdd(&n_gt, 1);
      cval[pos] = x;
      cidx[pos] = v;
    } else if (k == prefix) {
      const int pos = atomicAdd(&n_eq, 1);
      if (pos < kMaxBeam) eqidx[pos] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int base = n_gt;
    const int take = static_cast<int>(need);
    if (n_eq <= kMaxBeam) {
      for (int i = 1; i < n_eq; ++i) {  // insertion sort by index (n_eq is almost always 1)
        const int x = eqidx[i];
        int j = i - 1;
        while (j >= 0 && eqidx[j] > x) { eqidx[j + 1] = eqidx[j]; --j; }
        eqidx[j + 1] = x;
      }
      for (int i = 0; i < take; ++i) { cidx[base + i] = eqidx[i]; cval[base + i] = pred_s[eqidx[i]]; }
    } else {  // many exact ties (e.g. -inf logits): lowest indices by a linear scan
      int got = 0;
      for (int v = 0; v < a.V && got < take; ++v)
        if (float_key(pred_s[v]) == prefix) { cidx[base + got] = v; cval[base + got] = pred_s[v]; ++got; }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < a.beam; i += blockDim.x) {
    const ValIdx me{cval[i], cidx[i]};
    int rank = 0;
    for (int j = 0; j < a.beam; ++j) rank += better(ValIdx{cval[j], cidx[j]}, me) ? 1 : 0;
    cv[rank] = me.v + lp;
    cc[rank] = me.i;
  }
}

// ------------------------------------------------------------------ beam merge: one warp per neuron
// The candidate lists of the neuron's source rows (each sorted descending) are staged in shared memory with
// coalesced loads; the beam-step merge then runs on shared memory and the results are written in parallel.
__global__ void __launch_bounds__(128) beam_merge_kernel(const MergeArgs a) {
  __shared__ float val_s[kMaxBeam * kMaxBeam];
  __shared__ int flat_s[kMaxBeam];
  __shared__ float best_s[kMaxBeam];
  const int nrn = blockIdx.x;
  const int lane = threadIdx.x & 31;
  if (a.skip != nullptr && *a.skip) {
    // every beam of every neuron has ended: the reference would have left its loop; keep the beams as they are
    for (int j = threadIdx.x; j < a.beam; j += blockDim.x) {
      const int out = nrn * a.beam + j;
      a.next_tokens[out] = a.stop_index;
      a.next_lp[out] = a.cur_lp[out];
      a.backptr[out] = out;
      a.hist_tok[out] = static_cast<int>(a.stop_index);
      a.hist_bp[out] = j;
    }
    return;
  }
  const int row_base = nrn * a.in_rows;
  const int n_cand = a.in_rows * a.beam;
  const float* cand = a.cand_val + static_cast<long long>(row_base) * a.beam;
#pragma unroll 4
  for (int i = threadIdx.x; i < n_cand; i += blockDim.x) val_s[i] = cand[i];
  __syncthreads();
  int ptr0 = 0, ptr1 = 0;  // heads of source rows lane and lane+32
  const int r0 = lane, r1 = lane + 32;
  for (int j = 0; threadIdx.x < 32 && j < a.beam; ++j) {
    ValIdx c0{-INFINITY, 0x7fffffff}, c1{-INFINITY, 0x7fffffff};
    if (r0 < a.in_rows && ptr0 < a.beam) c0 = ValIdx{val_s[r0 * a.beam + ptr0], r0 * a.beam + ptr0};
    if (r1 < a.in_rows && ptr1 < a.beam) c1 = ValIdx{val_s[r1 * a.beam + ptr1], r1 * a.beam + ptr1};
    ValIdx best = better(c1, c0) ? c1 : c0;
    best = warp_argmax(best);
    int flat = best.i;
    if (flat == 0x7fffffff) flat = 0;  // degenerate (fewer finite candidates than beams)
    const int src = flat / a.beam;
    if (src == r0) ++ptr0;
    if (src == r1) ++ptr1;
    if (lane == 0) {
      flat_s[j] = flat;
      best_s[j] = best.v;
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < a.beam; j += blockDim.x) {
    const int out = nrn * a.beam + j;
    const int flat = flat_s[j];
    const int src = flat / a.beam;
    const int cls = a.cand_cls[static_cast<long long>(row_base) * a.beam + flat];
    a.next_tokens[out] = cls;
    a.next_lp[out] = best_s[j];
    a.backptr[out] = row_base + src;
    a.hist_tok[out] = cls;
    a.hist_bp[out] = src;
  }
}

__global__ void gather_state_kernel(const GatherArgs a) {
  if (a.skip != nullptr && *a.skip) return;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int H8 = a.H / 8;
  if (idx >= static_cast<long long>(a.R) * H8) return;
  const int j = 8 * (idx % H8);
  const long long r = idx / H8;
  const long long s = a.backptr != nullptr ? a.backptr[r] : r;
  *reinterpret_cast<uint4*>(a.dst_hi + r * a.dst_pitch + j) =
      *reinterpret_cast<const uint4*>(a.src_hi + s * a.src_pitch + j);
  if (a.dst_lo != nullptr)
    *reinterpret_cast<uint4*>(a.dst_lo + r * a.dst_pitch + j) =
        *reinterpret_cast<const uint4*>(a.src_lo + s * a.src_pitch + j);
  if (a.c_dst != nullptr) {
    const float4* cs = reinterpret_cast<const float4*>(a.c_src + s * a.H + j);
    float4* cd = reinterpret_cast<float4*>(a.c_dst + r * a.H + j);
    cd[0] = cs[0];
    cd[1] = cs[1];
  }
}

// ------------------------------------------------------------------ backtrack + early-exit length
__global__ void backtrack_kernel(const BacktrackArgs a) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int rows = a.n_neurons * a.beam;
  if (idx >= rows) return;
  const int nrn = idx / a.beam;
  int p = idx % a.beam;
  for (int t = a.length - 1; t >= 0; --t) {
    const int row = nrn * a.beam + p;
    a.seqs[static_cast<long long>(idx) * a.length + t] = a.hist_tok[static_cast<long long>(t) * rows + row];
    p = a.hist_bp[static_cast<long long>(t) * rows + row];
  }
}
// One CTA per reference batch group: T = first step t >= 1 before which every beam of every neuron in the group
// had emitted <stop> (allennlp: `if (last_predictions == end).all(): break`), else `length`.
__global__ void group_T_kernel(const BacktrackArgs a) {
  const int g = blockIdx.x;
  const int rows = a.n_neurons * a.beam;
  const int lo = g * a.group_size * a.beam;
  int hi = lo + a.group_size * a.beam;
  if (hi > rows) hi = rows;
  int T = a.length;
  for (int t = 1; t < a.length; ++t) {
    int ok = 1;
    for (int row = lo + threadIdx.x; row < hi; row += blockDim.x)
      if (a.hist_tok[static_cast<long long>(t - 1) * rows + row] != a.stop_index) ok = 0;
    if (__syncthreads_and(ok)) {
      T = t;
      break;
    }
  }
  if (threadIdx.x == 0) a.group_T[g] = T;
}

// ------------------------------------------------------------------ LM scoring
__global__ void lm_inputs_kernel(const long long* __restrict__ seqs, int M, int length, int t, long long start,
                                 long long* inputs, const int* __restrict__ skip) {
  if (skip != nullptr && *skip) return;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  inputs[m] = t == 0 ? start : seqs[static_cast<long long>(m) * length + t - 1];
}

__global__ void __launch_bounds__(256) lm_accumulate_kernel(const LmAccumArgs a) {
  if (a.skip != nullptr && *a.skip) return;
  __shared__ float red[8];
  const int m = blockIdx.x;
  const int T = a.group_T[(m / a.beam) / a.group_size];
  const long long* seq = a.seqs + static_cast<long long>(m) * a.length;
  // inputs = [<start>, seq...]; target t (= seq[t]) is kept iff no <stop> among inputs[0..t-1], i.e. the first
  // <stop> in seq sits at index >= t-1: the token AFTER the first <stop> is still counted (lms.py:93-96).
  int first_stop = a.length + 8;
  for (int i = 0; i < T; ++i)
    if (seq[i] == a.stop_index) { first_stop = i; break; }
  const bool keep = a.t < T && a.t <= first_stop + 1;
  if (!keep) return;  // uniform across the block
  const float* x = a.logits + static_cast<long long>(m) * a.ld;
  // one pass over the row: this thread's elements stay in registers between the max and the sum (V <= 8192),
  // longer rows fall back to re-reading.
  constexpr int kKeep = 8;
  float4 q[kKeep];
  const int V4 = a.V >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < kKeep; ++i) {
    const int v = threadIdx.x + i * 256;
    q[i] = v < V4 ? __ldg(x4 + v) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    mx = fmaxf(fmaxf(mx, fmaxf(q[i].x, q[i].y)), fmaxf(q[i].z, q[i].w));
  }
  for (int v = threadIdx.x + kKeep * 256; v < V4; v += 256) {
    const float4 t = __ldg(x4 + v);
    mx = fmaxf(fmaxf(mx, fmaxf(t.x, t.y)), fmaxf(t.z, t.w));
  }
  for (int v = 4 * V4 + threadIdx.x; v < a.V; v += 256) mx = fmaxf(mx, x[v]);
  mx = block_max(mx, red);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kKeep; ++i)
    if (threadIdx.x + i * 256 < V4) s += (expf(q[i].x - mx) + expf(q[i].y - mx)) + (expf(q[i].z - mx) + expf(q[i].w - mx));
  for (int v = threadIdx.x + kKeep * 256; v < V4; v += 256) {
    const float4 t = __ldg(x4 + v);
    s += (expf(t.x - mx) + expf(t.y - mx)) + (expf(t.z - mx) + expf(t.w - mx));
  }
  for (int v = 4 * V4 + threadIdx.x; v < a.V; v += 256) s += expf(x[v] - mx);
  s = block_sum(s, red);
  if (threadIdx.x == 0) a.lm_scores[m] += (x[seq[a.t]] - mx) - logf(s);
}

__global__ void rerank_select_kernel(const RerankArgs a) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= a.n_neurons) return;
  float best = -INFINITY;
  int bi = 0;
  for (int j = 0; j < a.beam; ++j) {
    const float lm = a.lm_scores != nullptr ? a.lm_scores[n * a.beam + j] : 0.0f;
    const float s = a.beam_lp[n * a.beam + j] - a.temperature * lm;  // decoders.py:507
    if (s > best || j == 0) {
      if (j == 0 || s > best) { best = s; bi = j; }
    }
  }
  for (int t = 0; t < a.length; ++t)
    a.out_tokens[static_cast<long long>(n) * a.length + t] =
        a.seqs[(static_cast<long long>(n) * a.beam + bi) * a.length + t];
  a.out_scores[n] = best;
  if (a.out_index != nullptr) a.out_index[n] = bi;
}

__global__ void check_done_kernel(const long long* __restrict__ tokens, int n, long long stop, int* flag) {
  int ok = 1;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    if (tokens[i] != stop) ok = 0;
  ok = __syncthreads_and(ok);
  if (threadIdx.x == 0) *flag = ok;
}
__global__ void lm_skip_kernel(const int* __restrict__ group_T, int groups, int length, int* lm_skip) {
  int max_T = 0;
  for (int g = 0; g < groups; ++g) max_T = max(max_T, group_T[g]);
  for (int t = threadIdx.x; t < length; t += blockDim.x) lm_skip[t] = t >= max_T ? 1 : 0;
}
__global__ void fill_i64_kernel(long long* dst, long long v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = v;
}
__global__ void fill_f32_kernel(float* dst, float v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = v;
}

inline unsigned blocks_for(long long n, int threads) { return static_cast<unsigned>((n + threads - 1) / threads); }
inline int last_err() {
  note_launch();
  return static_cast<int>(cudaGetLastError());
}

}  // namespace

int launch_split_rows(const float* src, long long src_pitch, __nv_bfloat16* dst_hi, __nv_bfloat16* dst_lo,
                      long long dst_pitch, int M, int K, cudaStream_t stream) {
  const long long n = static_cast<long long>(M) * (K / 2);
  if (n == 0) return 0;
  split_rows_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(src, src_pitch, dst_hi, dst_lo, dst_pitch, M, K / 2);
  return last_err();
}
int launch_mean_keys(const float* features, int B, int n_keys, int F, __nv_bfloat16* dst_hi, __nv_bfloat16* dst_lo,
                     cudaStream_t stream) {
  const long long n = static_cast<long long>(B) * (F / 2);
  mean_keys_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(features, B, n_keys, F, dst_hi, dst_lo);
  return last_err();
}
int launch_init_finish(const float* pre, int B, int H, float* h, float* c, __nv_bfloat16* h_hi,
                       __nv_bfloat16* h_lo, long long h_pitch, cudaStream_t stream) {
  const long long n = static_cast<long long>(B) * (H / 2);
  init_finish_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(pre, B, H, h, c, h_hi, h_lo, h_pitch);
  return last_err();
}
int launch_attend(const AttendArgs& a, cudaStream_t stream) {
  if (a.R == 0) return 0;
  if (a.n_keys > kApplyKeys && a.acc_ws == nullptr) return static_cast<int>(cudaErrorInvalidValue);
  const size_t smem1 = (a.A + a.n_keys) * sizeof(float);
  attn_scores_kernel<<<a.R, 256, smem1, stream>>>(a, a.attn_ws);
  note_launch();
  const size_t smem2 = static_cast<size_t>(a.rows_per_feature) * a.n_keys * sizeof(float);
  static size_t configured2 = 0;
  if (smem2 > 48 * 1024 && smem2 > configured2) {  // many keys (spatial encoder: k * 49) x beam rows
    if (smem2 > 227 * 1024) return static_cast<int>(cudaErrorInvalidValue);
    cudaError_t e = cudaFuncSetAttribute(attn_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem2));
    if (e != cudaSuccess) return static_cast<int>(e);
    configured2 = smem2;
  }
  dim3 grid(a.R / a.rows_per_feature, (a.F / 2 + 255) / 256);
  attn_apply_kernel<<<grid, 256, smem2, stream>>>(a, a.attn_ws);
  return last_err();
}
int launch_lstm_point(const LstmPointArgs& a, cudaStream_t stream) {
  const long long n = static_cast<long long>(a.R) * (a.H / 2);
  if (n == 0) return 0;
  lstm_point_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(a);
  return last_err();
}
int launch_embed_rows(const float* table, const long long* tokens, int M, int E, __nv_bfloat16* dst_hi,
                      __nv_bfloat16* dst_lo, long long dst_pitch, cudaStream_t stream, const int* skip) {
  const long long n = static_cast<long long>(M) * (E / 2);
  if (n == 0) return 0;
  embed_rows_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(table, tokens, M, E / 2, dst_hi, dst_lo, dst_pitch, skip);
  return last_err();
}
int launch_row_logsoftmax(const RowArgs& a, cudaStream_t stream) {
  if (a.R == 0) return 0;
  if (a.beam > kMaxBeam) return static_cast<int>(cudaErrorInvalidValue);
  const size_t smem = static_cast<size_t>(a.V) * sizeof(float);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    configured = smem;
  }
  row_kernel<<<a.R, 256, smem, stream>>>(a);
  return last_err();
}
int launch_beam_merge(const MergeArgs& a, cudaStream_t stream) {
  if (a.in_rows > 64 || a.beam > kMaxBeam) return static_cast<int>(cudaErrorInvalidValue);
  beam_merge_kernel<<<a.n_neurons, 128, 0, stream>>>(a);
  return last_err();
}
int launch_gather_state(const GatherArgs& a, cudaStream_t stream) {
  const long long n = static_cast<long long>(a.R) * (a.H / 8);
  if (n == 0) return 0;
  gather_state_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(a);
  return last_err();
}
int launch_backtrack(const BacktrackArgs& a, cudaStream_t stream) {
  const int rows = a.n_neurons * a.beam;
  backtrack_kernel<<<blocks_for(rows, 128), 128, 0, stream>>>(a);
  note_launch();
  const int groups = (a.n_neurons + a.group_size - 1) / a.group_size;
  group_T_kernel<<<groups, 256, 0, stream>>>(a);
  return last_err();
}
int launch_lm_inputs(const long long* seqs, int M, int length, int t, long long start_index, long long* inputs,
                     cudaStream_t stream, const int* skip) {
  lm_inputs_kernel<<<blocks_for(M, 256), 256, 0, stream>>>(seqs, M, length, t, start_index, inputs, skip);
  return last_err();
}
int launch_lm_accumulate(const LmAccumArgs& a, cudaStream_t stream) {
  lm_accumulate_kernel<<<a.M, 256, 0, stream>>>(a);
  return last_err();
}
int launch_rerank_select(const RerankArgs& a, cudaStream_t stream) {
  rerank_select_kernel<<<blocks_for(a.n_neurons, 128), 128, 0, stream>>>(a);
  return last_err();
}
int launch_check_done(const long long* tokens, int n, long long stop_index, int* flag, cudaStream_t stream) {
  check_done_kernel<<<1, 256, 0, stream>>>(tokens, n, stop_index, flag);
  return last_err();
}
int launch_lm_skip(const int* group_T, int groups, int length, int* lm_skip, cudaStream_t stream) {
  lm_skip_kernel<<<1, 32, 0, stream>>>(group_T, groups, length, lm_skip);
  return last_err();
}
int launch_fill_i64(long long* dst, long long value, int n, cudaStream_t stream) {
  if (n == 0) return 0;
  fill_i64_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(dst, value, n);
  return last_err();
}
int launch_fill_f32(float* dst, float value, int n, cudaStream_t stream) {
  if (n == 0) return 0;
  fill_f32_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(dst, value, n);
  return last_err();
}

}  // namespace milan
