// Fused non-GEMM kernels of the beam step and the LM rerank (see decode_fused.h). fp32 throughout; operands that
// feed a tensor-core GEMM are emitted as IEEE fp16 (hi, lo) pairs like decoder.cu does.
#include "decode_fused.h"

#include "conv_gemm.h"  // note_launch
#include "decoder.h"    // kMaxBeam
#include "ptx.cuh"

#include <algorithm>
#include <cfloat>
#include <cmath>

namespace milan {

namespace {

__device__ __forceinline__ void store_split2(__nv_bfloat16* hi, __nv_bfloat16* lo, long long off, float v0,
                                             float v1) {
  uint32_t h, l;
  split_fp16x2(v0, v1, h, l);
  *reinterpret_cast<uint32_t*>(hi + off) = h;
  if (lo != nullptr) *reinterpret_cast<uint32_t*>(lo + off) = l;
}

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

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
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// k-th largest (k >= 1) of the keys a warp holds KMAX per lane (unused slots = 0, below every float_key): the answer is
// built bit by bit from the top, one warp-wide count per bit - 32 x (KMAX compares + one redux), no shared memory.
template <int KMAX>
__device__ __forceinline__ unsigned warp_kth_largest(const unsigned (&key)[KMAX], int k) {
  unsigned ans = 0;
#pragma unroll 1
  for (int bit = 31; bit >= 0; --bit) {
    const unsigned trial = ans | (1u << bit);
    int c = 0;
#pragma unroll
    for (int j = 0; j < KMAX; ++j) c += key[j] >= trial ? 1 : 0;
    if (__reduce_add_sync(0xffffffffu, c) >= k) ans = trial;
  }
  return ans;
}
__device__ __forceinline__ unsigned float_key(float x) {  // order-preserving float -> uint
  const unsigned u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float float_from_key(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// ------------------------------------------------------------------ attention + gating + operand assembly
// Cluster of kAttendCluster CTAs per feature set (neuron). Phase A: the neuron's rows are dealt to the (CTA, warp)
// pairs of the cluster, one row per warp, all in flight at once: attention weights (15 warp reductions over A),
// the token embedding, and the parent's h' copied into the operand row. Cluster barrier. Phase B: each CTA owns a
// column slice of the features, reads it ONCE into registers and produces that slice of `attenuated * gate` for
// every row of the neuron.
__device__ __forceinline__ void attend_phase(const AttendFusedArgs& a, float* sm) {
  float* q_s = sm;                                     // [8 warps][A]
  float* w_s = q_s + 8 * a.A;                          // [rpf][n_keys]
  int* src_s = reinterpret_cast<int*>(w_s + a.rows_per_feature * a.n_keys);  // [rpf]
  const int fidx = blockIdx.x / kAttendCluster;
  const int crank = blockIdx.x % kAttendCluster;
  const int rpf = a.rows_per_feature;
  const long long row0 = static_cast<long long>(fidx) * rpf;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  for (int rl = crank + kAttendCluster * warp; rl < rpf; rl += kAttendCluster * 8) {
    const long long r = row0 + rl;
    const long long src = a.src_row != nullptr ? a.src_row[r] : r;
    float* q = q_s + warp * a.A;
    const float* qrow = a.q + src * a.q_pitch;
    for (int i = lane; i < a.A; i += 32) q[i] = qrow[i];
    __syncwarp();
    float score = -INFINITY;  // lane k keeps the score of key k
    {
      // all keys of an attention column at once: n_keys independent tanh chains per lane keep the SFU pipe full (one
      // key at a time left every FADD waiting for its own ex2 + rcp). Per key the summation order is unchanged.
      float acc[kFusedMaxKeys];
#pragma unroll
      for (int k = 0; k < kFusedMaxKeys; ++k) acc[k] = 0.f;
      const float* khb = a.kh + static_cast<long long>(fidx) * a.n_keys * a.A;
      const int last = a.n_keys - 1;
      for (int i = lane; i < a.A; i += 32) {
        const float qi = q[i], wo = __ldg(a.w_o + i);
        float kv[kFusedMaxKeys];  // branch-free: slots past n_keys re-read the last key and are never used
#pragma unroll
        for (int k = 0; k < kFusedMaxKeys; ++k) kv[k] = __ldg(khb + static_cast<long long>(min(k, last)) * a.A + i);
#pragma unroll
        for (int k = 0; k < kFusedMaxKeys; ++k) acc[k] += wo * tanh_fast(qi + kv[k]);
      }
#pragma unroll
      for (int k = 0; k < kFusedMaxKeys; ++k) {
        if (k < a.n_keys) {
          const float sk = warp_sum(acc[k]);
          if (lane == k) score = sk + a.b_o;
        }
      }
    }
    const float mx = warp_max(score);
    const float e = lane < a.n_keys ? expf(score - mx) : 0.f;
    const float den = warp_sum(e);
    if (lane < a.n_keys) {
      const float w = e / den;
      a.attn_ws[r * a.n_keys + lane] = w;
      if (a.attn_out != nullptr) a.attn_out[r * a.attn_pitch + lane] = w;
    }
    const long long xoff = r * a.x_pitch;
    const float* erow = a.embedding + a.tokens[r] * a.E;
    for (int e2 = lane; e2 < a.E / 2; e2 += 32) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(erow + 2 * e2));
      store_split2(a.x_hi, a.x_lo, xoff + 2 * e2, v.x, v.y);
    }
    if (a.h_src_hi != nullptr) {  // the recurrent state follows the backpointer
      const long long dst = xoff + a.E + a.F, from = src * a.h_src_pitch;
      for (int j = lane; j < a.H / 8; j += 32) {
        *reinterpret_cast<uint4*>(a.x_hi + dst + 8 * j) = *reinterpret_cast<const uint4*>(a.h_src_hi + from + 8 * j);
        if (a.x_lo != nullptr)
          *reinterpret_cast<uint4*>(a.x_lo + dst + 8 * j) = *reinterpret_cast<const uint4*>(a.h_src_lo + from + 8 * j);
      }
    }
    __syncwarp();
  }
  __threadfence();
  cluster_sync_all();

  // phase B. attn_ws was written by other CTAs of this cluster during this kernel: plain (coherent) loads only.
  const volatile float* ws = a.attn_ws + row0 * a.n_keys;
  for (int i = threadIdx.x; i < rpf * a.n_keys; i += blockDim.x) w_s[i] = ws[i];
  for (int i = threadIdx.x; i < rpf; i += blockDim.x)
    src_s[i] = a.src_row != nullptr ? a.src_row[row0 + i] : static_cast<int>(row0 + i);
  __syncthreads();
  const int F2 = a.F / 2;
  const int per = (F2 + kAttendCluster - 1) / kAttendCluster;
  const int j_end = min(F2, (crank + 1) * per);
  const float* fb = a.features + static_cast<long long>(fidx) * a.n_keys * a.F;
  for (int j2 = crank * per + threadIdx.x; j2 < j_end; j2 += blockDim.x) {
    float2 f[kFusedMaxKeys];
#pragma unroll
    for (int k = 0; k < kFusedMaxKeys; ++k)
      f[k] = k < a.n_keys ? __ldg(reinterpret_cast<const float2*>(fb + static_cast<long long>(k) * a.F + 2 * j2))
                          : make_float2(0.f, 0.f);
#pragma unroll 8
    for (int rl = 0; rl < rpf; ++rl) {
      const float2 g = *reinterpret_cast<const float2*>(a.gate + static_cast<long long>(src_s[rl]) * a.gate_pitch + 2 * j2);
      const float* w = w_s + rl * a.n_keys;
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int k = 0; k < kFusedMaxKeys; ++k) {
        if (k < a.n_keys) {
          s0 += w[k] * f[k].x;
          s1 += w[k] * f[k].y;
        }
      }
      store_split2(a.x_hi, a.x_lo, (row0 + rl) * a.x_pitch + a.E + 2 * j2, s0 * g.x, s1 * g.y);
    }
  }
}

__global__ void __cluster_dims__(kAttendCluster, 1, 1) __launch_bounds__(256)
    attend_fused_kernel(const AttendFusedArgs a) {
  if (a.skip != nullptr && *a.skip) return;  // uniform over the grid
  extern __shared__ float sm[];
  attend_phase(a, sm);
}

// ------------------------------------------------------------------ per-neuron top-k + beam merge
// Exact top-`beam` of one row by MSB radix select (the fallback of decoder.cu's row_kernel, run here by the 256
// threads of a CTA on a row staged in shared memory): only reached when more than kSelectCandCap values pass the
// prefilter, e.g. masses of tied logits.
__device__ void exact_topk_256(const float* pred_s, int V, int beam, float lp, float* cv, int* cc, unsigned* hist,
                               unsigned* warp_tot, unsigned* sel, int* counts, float* cval, int* cidx, int* eqidx) {
  const int tid = threadIdx.x;  // < 256
  const int lane = tid & 31, warp = tid >> 5;
  unsigned prefix = 0, need = static_cast<unsigned>(beam);
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    hist[tid] = 0;
    named_bar_sync(1, 256);
    for (int v = tid; v < V; v += 256) {
      const unsigned k = float_key(pred_s[v]);
      if (pass == 0 || (k >> (shift + 8)) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1u);
    }
    named_bar_sync(1, 256);
    const unsigned own = hist[tid];
    unsigned suf = own;  // inclusive suffix sum over digits (thread t <-> digit t)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_down_sync(0xffffffffu, suf, o);
      if (lane + o < 32) suf += t;
    }
    if (lane == 0) warp_tot[warp] = suf;
    named_bar_sync(1, 256);
    for (int w = warp + 1; w < 8; ++w) suf += warp_tot[w];
    const unsigned excl = suf - own;
    if (excl < need && suf >= need) {
      sel[0] = tid;
      sel[1] = need - excl;
    }
    named_bar_sync(1, 256);
    prefix = (prefix << 8) | sel[0];
    need = sel[1];
    named_bar_sync(1, 256);
  }
  if (tid == 0) { counts[0] = 0; counts[1] = 0; }
  named_bar_sync(1, 256);
  for (int v = tid; v < V; v += 256) {
    const float x = pred_s[v];
    const unsigned k = float_key(x);
    if (k > prefix) {
      const int pos = atomicAdd(&counts[0], 1);
      cval[pos] = x;
      cidx[pos] = v;
    } else if (k == prefix) {
      const int pos = atomicAdd(&counts[1], 1);
      if (pos < kMaxBeam) eqidx[pos] = v;
    }
  }
  named_bar_sync(1, 256);
  if (tid == 0) {
    const int base = counts[0];
    const int take = static_cast<int>(need);
    if (counts[1] <= kMaxBeam) {
      for (int i = 1; i < counts[1]; ++i) {  // by index (almost always a single element)
        const int x = eqidx[i];
        int j = i - 1;
        while (j >= 0 && eqidx[j] > x) { eqidx[j + 1] = eqidx[j]; --j; }
        eqidx[j + 1] = x;
      }
      for (int i = 0; i < take; ++i) { cidx[base + i] = eqidx[i]; cval[base + i] = pred_s[eqidx[i]]; }
    } else {  // many exact ties: lowest indices by a linear scan
      int got = 0;
      for (int v = 0; v < V && got < take; ++v)
        if (float_key(pred_s[v]) == prefix) { cidx[base + got] = v; cval[base + got] = pred_s[v]; ++got; }
    }
  }
  named_bar_sync(1, 256);
  for (int i = tid; i < beam; i += 256) {
    const ValIdx me{cval[i], cidx[i]};
    int rank = 0;
    for (int j = 0; j < beam; ++j) rank += better(ValIdx{cval[j], cidx[j]}, me) ? 1 : 0;
    cv[rank] = me.v + lp;
    cc[rank] = me.i;
  }
  named_bar_sync(1, 256);
}

struct SelectSmem {  // fixed part of beam_select's shared memory
  float best[kMaxBeam];
  int flat[kMaxBeam];
  float row_max[8];  // per warp: the row it owns, when that row needs the exact fallback
  float row_lse[8];
  float row_lp[8];
  int row_fallback[8];
  unsigned hist[256];
  unsigned warp_tot[8];
  unsigned sel[2];
  int counts[2];
  float cval[kMaxBeam];
  int cidx[kMaxBeam];
  int eqidx[kMaxBeam];
};
constexpr int kWarpScratchFloats = 2 * kSelectCandCap + kSelectMaxGroups;  // cand values | cand ids | group maxima
constexpr int kSelectWarps = kSelectThreads / 32;

// Cluster of kSelectCluster CTAs per neuron, one source row per warp (in_rows <= 64 = 8 CTAs x 8 warps): finish the
// row's log-softmax from the GEMM's partials, prefilter, rank exactly, write the sorted top-`beam` list to global
// memory. Cluster barrier. CTA 0 of the cluster merges the lists into the next beam and publishes the early-exit flag.
// Every beam of every neuron has ended: the reference has left its loop; the beams stay as they are.
__device__ __forceinline__ void select_done_phase(const BeamSelectArgs& a) {
  const int nrn = blockIdx.x / kSelectCluster;
  if (blockIdx.x % kSelectCluster == 0) {
    for (int j = threadIdx.x; j < a.beam; j += blockDim.x) {
      const int out = nrn * a.beam + j;
      a.next_tokens[out] = a.stop_index;
      a.next_lp[out] = a.cur_lp[out];
      a.backptr[out] = out;
      a.hist_tok[out] = static_cast<int>(a.stop_index);
      a.hist_bp[out] = j;
    }
  }
}

// Rows phase (all CTAs of the cluster): each warp ranks its source row and writes the sorted top-`beam` list.
__device__ __forceinline__ void select_rows_phase(const BeamSelectArgs& a, uint8_t* smem_raw) {
  SelectSmem& S = *reinterpret_cast<SelectSmem*>(smem_raw);
  float* scratch = reinterpret_cast<float*>(smem_raw + sizeof(SelectSmem));  // per-warp scratch | staged row | lists
  const int nrn = blockIdx.x / kSelectCluster;
  const int crank = blockIdx.x % kSelectCluster;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < 8) S.row_fallback[threadIdx.x] = -1;
  __syncthreads();

  const int row_base = nrn * a.in_rows;
  const int V4 = static_cast<int>(a.ld >> 2);
  const int rl = crank + kSelectCluster * warp;  // this warp's source row
  if (rl < a.in_rows) {
    const long long r = row_base + rl;
    float* cv = a.cand_val + r * a.beam;
    int* cc = a.cand_cls + r * a.beam;
    const float lp = a.last_lp != nullptr ? a.last_lp[r] : 0.0f;
    if (a.last_tokens[r] == a.stop_index) {
      // finished beam: only <stop> at cost 0 survives (allennlp log_probs_after_end); the other per-node candidates
      // carry min_value_of_dtype there and are never selected while >= beam finite candidates exist
      for (int j = lane; j < a.beam; j += 32) {
        cv[j] = j == 0 ? lp + 0.0f : -INFINITY;
        cc[j] = static_cast<int>(a.stop_index);
      }
    } else {
      const float2* pr = a.partials + r * a.n_seg;
      float m = -INFINITY;
      for (int i = lane; i < a.n_seg; i += 32) m = fmaxf(m, pr[i].x);
      const float M = warp_max(m);
      float s = 0.f;
      for (int i = lane; i < a.n_seg; i += 32) {
        const float2 q = pr[i];
        if (q.y > 0.f) s += q.y * expf(q.x - M);
      }
      const float lse = logf(warp_sum(s));
      float* cand_v = scratch + warp * kWarpScratchFloats;
      int* cand_i = reinterpret_cast<int*>(cand_v + kSelectCandCap);
      float* gm = cand_v + 2 * kSelectCandCap;
      // The beam-th largest of G disjoint segment maxima bounds the beam-th largest value of the row from below.
      const int per = (a.n_seg + kSelectMaxGroups - 1) / kSelectMaxGroups;
      const int G = (a.n_seg + per - 1) / per;
      for (int g = lane; g < G; g += 32) {
        float x = -INFINITY;
        for (int i = g * per; i < min(a.n_seg, (g + 1) * per); ++i) x = fmaxf(x, pr[i].x);
        gm[g] = x;
      }
      __syncwarp();
      float tau = -INFINITY;
      if (G >= a.beam) {  // (G <= kSelectMaxGroups = 4 x 32)
        unsigned gkey[kSelectMaxGroups / 32];
#pragma unroll
        for (int j = 0; j < kSelectMaxGroups / 32; ++j) gkey[j] = lane + 32 * j < G ? float_key(gm[lane + 32 * j]) : 0u;
        const unsigned kth = warp_kth_largest(gkey, a.beam);
        tau = float_from_key(kth);
      }
      // compare in the log-softmax domain: distinct logits may round to equal log-probabilities, and ties are
      // ranked by class index
      const float y_tau = (tau - M) - lse;
      const float4* x4 = reinterpret_cast<const float4*>(a.logits + r * a.ld);
      int count = 0;
      constexpr int kBatch = 4;  // independent 16-byte loads in flight per lane
      for (int base = 0; base < V4; base += 32 * kBatch) {
        float4 q[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const int idx = base + u * 32 + lane;
          q[u] = idx < V4 ? x4[idx] : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const int idx = base + u * 32 + lane;
          const float e[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
          float y[4];
          bool take[4];
          bool any = false;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            y[c] = (e[c] - M) - lse;
            take[c] = idx < V4 && 4 * idx + c < a.V && y[c] >= y_tau;
            any |= take[c];
          }
          if (__any_sync(0xffffffffu, any)) {  // ~1.5 % of the values pass: most groups of 128 hold none
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const unsigned ballot = __ballot_sync(0xffffffffu, take[c]);
              if (take[c]) {
                const int pos = count + __popc(ballot & ((1u << lane) - 1u));
                if (pos < kSelectCandCap) { cand_v[pos] = y[c]; cand_i[pos] = 4 * idx + c; }
              }
              count += __popc(ballot);
            }
          }
        }
      }
      __syncwarp();
      if (count > kSelectCandCap) {  // warp-uniform: more ties than the prefilter holds -> exact fallback below
        if (lane == 0) {
          S.row_fallback[warp] = rl;
          S.row_max[warp] = M;
          S.row_lse[warp] = lse;
          S.row_lp[warp] = lp;
        }
      } else {
        for (int j = count + lane; j < a.beam; j += 32) {  // fewer candidates than beams (NaN rows): pad
          cv[j] = -INFINITY;
          cc[j] = 0;
        }
        // The prefilter passes 1.5-5x `beam` candidates. Rank only those at or above the beam-th largest VALUE
        // (found by a bitwise search over the keys, held 8 per lane), not every candidate against every other.
        int n_rank = count;
        if (count > a.beam) {
          constexpr int kPerLane = kSelectCandCap / 32;
          float v[kPerLane];
          int ix[kPerLane];
          unsigned key[kPerLane];
#pragma unroll
          for (int j = 0; j < kPerLane; ++j) {
            const int i = lane + 32 * j;
            v[j] = i < count ? cand_v[i] : 0.f;
            ix[j] = i < count ? cand_i[i] : 0;
            key[j] = i < count ? float_key(v[j]) : 0u;
          }
          const unsigned kth = warp_kth_largest(key, a.beam);
          __syncwarp();  // every lane holds its candidates in registers: the lists can be compacted in place
          int kept = 0;
#pragma unroll
          for (int j = 0; j < kPerLane; ++j) {
            const bool keep = lane + 32 * j < count && key[j] >= kth;
            const unsigned ballot = __ballot_sync(0xffffffffu, keep);
            if (keep) {
              const int pos = kept + __popc(ballot & ((1u << lane) - 1u));
              cand_v[pos] = v[j];
              cand_i[pos] = ix[j];
            }
            kept += __popc(ballot);
          }
          n_rank = kept;  // >= beam; more only when values tie at the threshold (ranked by class index below)
          __syncwarp();
        }
        for (int i = lane; i < n_rank; i += 32) {
          const ValIdx me{cand_v[i], cand_i[i]};
          int rank = 0;
          for (int j = 0; j < n_rank; ++j) rank += better(ValIdx{cand_v[j], cand_i[j]}, me) ? 1 : 0;
          if (rank < a.beam) {
            cv[rank] = me.v + lp;
            cc[rank] = me.i;
          }
        }
      }
    }
  }
  __syncthreads();
  // ---- rows the prefilter could not bound: exact radix select on the staged row, whole CTA, one row at a time
  for (int w = 0; w < kSelectWarps; ++w) {
    const int frl = S.row_fallback[w];
    if (frl < 0) continue;  // uniform (shared memory)
    float* pred_s = scratch;
    const long long r = row_base + frl;
    const float* x = a.logits + r * a.ld;
    const float M = S.row_max[w], lse = S.row_lse[w];
    for (int v = threadIdx.x; v < a.V; v += kSelectThreads) pred_s[v] = (x[v] - M) - lse;
    named_bar_sync(1, 256);
    exact_topk_256(pred_s, a.V, a.beam, S.row_lp[w], a.cand_val + r * a.beam, a.cand_cls + r * a.beam, S.hist,
                   S.warp_tot, S.sel, S.counts, S.cval, S.cidx, S.eqidx);
  }
}

// Merge phase (CTA 0 of the cluster, after a cluster barrier): the next beam = the `beam` best of the in_rows sorted
// lists. The lists were written by other CTAs of this cluster during this kernel: coherent loads only.
__device__ __forceinline__ void select_merge_phase(const BeamSelectArgs& a, uint8_t* smem_raw) {
  SelectSmem& S = *reinterpret_cast<SelectSmem*>(smem_raw);
  float* scratch = reinterpret_cast<float*>(smem_raw + sizeof(SelectSmem));
  const int nrn = blockIdx.x / kSelectCluster;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row_base = nrn * a.in_rows;
  float* val_s = scratch;
  int* cls_s = reinterpret_cast<int*>(val_s + a.in_rows * a.beam);
  const int n_cand = a.in_rows * a.beam;
  const volatile float* gv = a.cand_val + static_cast<long long>(row_base) * a.beam;
  const volatile int* gc = a.cand_cls + static_cast<long long>(row_base) * a.beam;
  for (int i = threadIdx.x; i < n_cand; i += blockDim.x) {
    val_s[i] = gv[i];
    cls_s[i] = gc[i];
  }
  __syncthreads();
  // The next beam = the `beam` best of n_cand candidates in (value desc, flat index asc) order - what a k-way merge of
  // the sorted lists yields. In parallel: the beam-th largest VALUE by a bitwise search over the keys (CTA-wide counts,
  // one barrier per bit), then only the candidates at or above it are ranked against each other. (The sequential
  // k-way merge by one warp took ~25 us of the step with the other seven CTAs of the cluster waiting; it stays as the
  // fallback for mass ties at the threshold, e.g. fewer finite candidates than beams.)
  constexpr int kPer = (kSelectCluster * kSelectWarps * kMaxBeam + kSelectThreads - 1) / kSelectThreads;
  constexpr int kWinCap = 128;
  unsigned* cnt = S.hist;          // [0, 32): candidates at or above the trial key of each bit; [33]: winners
  unsigned* win = S.hist + 64;     // [kWinCap] flat indices of the winners
  unsigned key[kPer];
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    const int i = threadIdx.x + kSelectThreads * j;
    key[j] = i < n_cand ? float_key(val_s[i]) : 0u;
  }
  if (threadIdx.x < 34) cnt[threadIdx.x] = 0;
  __syncthreads();
  unsigned kth = 0;
#pragma unroll 1
  for (int bit = 31; bit >= 0; --bit) {
    const unsigned trial = kth | (1u << bit);
    int c = 0;
#pragma unroll
    for (int j = 0; j < kPer; ++j) c += key[j] >= trial ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0 && c > 0) atomicAdd(&cnt[bit], static_cast<unsigned>(c));
    __syncthreads();
    if (cnt[bit] >= static_cast<unsigned>(a.beam)) kth = trial;
  }
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    const int i = threadIdx.x + kSelectThreads * j;
    if (i < n_cand && key[j] >= kth) {
      const unsigned pos = atomicAdd(&cnt[33], 1u);
      if (pos < kWinCap) win[pos] = static_cast<unsigned>(i);
    }
  }
  __syncthreads();
  const int n_win = static_cast<int>(cnt[33]);  // >= beam; uniform
  if (n_win <= kWinCap) {
    for (int w = threadIdx.x; w < n_win; w += kSelectThreads) {
      const int flat = static_cast<int>(win[w]);
      const ValIdx me{val_s[flat], flat};
      int rank = 0;
      for (int u = 0; u < n_win; ++u) {
        const int fu = static_cast<int>(win[u]);
        rank += better(ValIdx{val_s[fu], fu}, me) ? 1 : 0;
      }
      if (rank < a.beam) {
        S.flat[rank] = flat;
        S.best[rank] = me.v;
      }
    }
  } else if (warp == 0) {  // one warp, two list heads per lane
    int ptr0 = 0, ptr1 = 0;
    const int r0 = lane, r1 = lane + 32;
    for (int j = 0; j < a.beam; ++j) {
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
        S.flat[j] = flat;
        S.best[j] = best.v;
      }
    }
  }
  __syncthreads();
  int ended = 1;
  for (int j = threadIdx.x; j < a.beam; j += blockDim.x) {
    const int out = nrn * a.beam + j;
    const int flat = S.flat[j];
    const int src = flat / a.beam;
    const int cls = cls_s[flat];
    a.next_tokens[out] = cls;
    a.next_lp[out] = S.best[j];
    a.backptr[out] = row_base + src;
    a.hist_tok[out] = cls;
    a.hist_bp[out] = src;
    if (cls != a.stop_index) ended = 0;
  }
  // allennlp: `if (last_predictions == end).all(): break` — the last neuron to finish publishes it for the next step
  ended = __syncthreads_and(ended);
  if (threadIdx.x == 0) {
    if (ended) atomicAdd(&a.counters[0], 1);
    __threadfence();
    const int ticket = atomicAdd(&a.counters[1], 1);
    if (ticket == a.n_neurons - 1) {
      __threadfence();
      const int all = atomicAdd(&a.counters[0], 0);
      *a.done_flag = all == a.n_neurons ? 1 : 0;
      a.counters[0] = 0;
      a.counters[1] = 0;
    }
  }
}

__global__ void __cluster_dims__(kSelectCluster, 1, 1) __launch_bounds__(kSelectThreads)
    beam_select_kernel(const BeamSelectArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  if (*a.done_flag != 0) {
    select_done_phase(a);
    return;  // uniform over the grid: nobody reaches the cluster barrier
  }
  select_rows_phase(a, smem_raw);
  __threadfence();
  cluster_sync_all();
  if (blockIdx.x % kSelectCluster == 0) select_merge_phase(a, smem_raw);
}

// Step t's selection and step t + 1's attention as ONE launch: both are per-neuron work on the same cluster shape, and
// the attention of the new beam needs nothing but what this neuron's merge has just produced (tokens, backpointers)
// and what the head GEMM left in the PARENT rows (query, gate, h'). A beam step is then three launches - this kernel,
// the LSTM GEMM, the head GEMM - each boundary a grid-wide dependency (every output tile of a GEMM needs whole rows
// of the previous one).
// The early-exit flag is read ONCE, at kernel start (it was published by the previous step's launch): the merge of
// another cluster may publish it for the next step while this kernel runs, and a second read could differ between
// the CTAs of a cluster. The attention of a beam that has just ended everywhere is computed once for nothing.
static_assert(kSelectCluster == kAttendCluster && kSelectThreads == 256, "select + attend share one cluster shape");
// (4 CTAs per SM: 64 neurons x 8 CTAs must be resident at once - at 78 registers the 512 CTAs took two waves and the
// merged launch was slower than the two it replaces)
__global__ void __cluster_dims__(kSelectCluster, 1, 1) __launch_bounds__(kSelectThreads, 4)
    select_attend_kernel(const BeamSelectArgs s, const AttendFusedArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  if (*s.done_flag != 0) {
    select_done_phase(s);
    return;  // uniform over the grid
  }
  select_rows_phase(s, smem_raw);
  __threadfence();
  cluster_sync_all();
  if (blockIdx.x % kSelectCluster == 0) {
    select_merge_phase(s, smem_raw);
    __threadfence();  // next_tokens / backptr: read by every CTA of the cluster below
  }
  cluster_sync_all();
  attend_phase(a, reinterpret_cast<float*>(smem_raw));
}

// ------------------------------------------------------------------ LM rerank read-out
__global__ void __launch_bounds__(256) lm_finalize_kernel(const LmFinalizeArgs a) {
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (m >= a.M) return;
  const int T = a.group_T[(m / a.beam) / a.group_size];
  const long long* seq = a.seqs + static_cast<long long>(m) * a.length;
  // inputs = [<start>, seq...]; target t (= seq[t]) is kept iff no <stop> among inputs[0..t-1], i.e. the first <stop>
  // in seq sits at index >= t-1: the token AFTER the first <stop> is still counted (lms.py:93-96).
  int first_stop = a.length + 8;
  for (int i = 0; i < T; ++i)
    if (seq[i] == a.stop_index) { first_stop = i; break; }
  float score = 0.f;
  for (int t = 0; t < T && t <= first_stop + 1; ++t) {
    const float2* pr = a.partials + (static_cast<long long>(t) * a.M + m) * a.n_seg;
    float mx = -INFINITY;
    for (int i = lane; i < a.n_seg; i += 32) mx = fmaxf(mx, pr[i].x);
    mx = warp_max(mx);
    float s = 0.f;
    for (int i = lane; i < a.n_seg; i += 32) {
      const float2 q = pr[i];
      if (q.y > 0.f) s += q.y * expf(q.x - mx);
    }
    s = warp_sum(s);
    score += (a.tgt_logit[static_cast<long long>(t) * a.M + m] - mx) - logf(s);
  }
  if (lane == 0) a.lm_scores[m] = score;
}

__global__ void __launch_bounds__(256) lm_input_table_kernel(const float* __restrict__ w_ih,
                                                             const float* __restrict__ emb, int E, int H,
                                                             float* __restrict__ table) {
  extern __shared__ float e_s[];  // [E]
  const int v = blockIdx.x;
  for (int i = threadIdx.x; i < E; i += blockDim.x) e_s[i] = emb[static_cast<long long>(v) * E + i];
  __syncthreads();
  for (int col = threadIdx.x; col < 4 * H; col += blockDim.x) {
    const int u = col >> 2, g = col & 3;
    const float* w = w_ih + static_cast<long long>(g * H + u) * E;
    float acc = 0.f;
    for (int e = 0; e < E; ++e) acc = fmaf(w[e], e_s[e], acc);
    table[static_cast<long long>(v) * 4 * H + col] = acc;
  }
}

inline int last_err() {
  note_launch();
  return static_cast<int>(cudaGetLastError());
}

// cudaFuncSetAttribute is per device: remember what each device has been configured for
template <class K>
int ensure_dynamic_smem(K kernel, size_t bytes, size_t (&configured)[64]) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return static_cast<int>(e);
  if (dev < 0 || dev >= 64) return static_cast<int>(cudaErrorInvalidDevice);
  if (bytes > configured[dev]) {
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
    if (e != cudaSuccess) return static_cast<int>(e);
    configured[dev] = bytes;
  }
  return 0;
}

}  // namespace

int launch_attend_fused(const AttendFusedArgs& a, cudaStream_t stream) {
  if (a.R == 0) return 0;
  if (a.n_keys > kFusedMaxKeys || a.rows_per_feature < 1 || a.R % a.rows_per_feature != 0 || (a.F & 1) || (a.E & 1) ||
      (a.H & 7))
    return static_cast<int>(cudaErrorInvalidValue);
  const size_t smem = (8 * static_cast<size_t>(a.A) + static_cast<size_t>(a.rows_per_feature) * a.n_keys) * sizeof(float) +
                      a.rows_per_feature * sizeof(int);
  if (smem > 48 * 1024) return static_cast<int>(cudaErrorInvalidValue);
  const int sets = a.R / a.rows_per_feature;
  attend_fused_kernel<<<sets * kAttendCluster, 256, smem, stream>>>(a);
  return last_err();
}

size_t beam_select_smem_bytes(int in_rows, int beam, int V) {
  const size_t lists = static_cast<size_t>(in_rows) * beam * (sizeof(float) + sizeof(int));
  const size_t warps = static_cast<size_t>(kSelectWarps) * kWarpScratchFloats * sizeof(float);
  const size_t staged = static_cast<size_t>(V) * sizeof(float);
  return sizeof(SelectSmem) + std::max(lists, std::max(warps, staged)) + 16;
}

int launch_beam_select(const BeamSelectArgs& a, cudaStream_t stream) {
  if (a.n_neurons == 0) return 0;
  if (a.beam < 1 || a.beam > kMaxBeam || a.in_rows < 1 || a.in_rows > kSelectCluster * kSelectWarps || a.n_seg < 1 ||
      (a.ld & 3))
    return static_cast<int>(cudaErrorInvalidValue);
  const size_t smem = beam_select_smem_bytes(a.in_rows, a.beam, a.V);
  if (smem > 227 * 1024) return static_cast<int>(cudaErrorInvalidValue);
  static size_t configured[64] = {};
  if (int rc = ensure_dynamic_smem(beam_select_kernel, smem, configured)) return rc;
  beam_select_kernel<<<a.n_neurons * kSelectCluster, kSelectThreads, smem, stream>>>(a);
  return last_err();
}

int launch_select_attend(const BeamSelectArgs& s, const AttendFusedArgs& a, cudaStream_t stream) {
  if (s.n_neurons == 0) return 0;
  if (s.beam < 1 || s.beam > kMaxBeam || s.in_rows < 1 || s.in_rows > kSelectCluster * kSelectWarps || s.n_seg < 1 ||
      (s.ld & 3))
    return static_cast<int>(cudaErrorInvalidValue);
  if (a.n_keys > kFusedMaxKeys || a.rows_per_feature != s.beam || a.R != s.n_neurons * s.beam || (a.F & 1) ||
      (a.E & 1) || (a.H & 7) || a.tokens != s.next_tokens || a.src_row != s.backptr)
    return static_cast<int>(cudaErrorInvalidValue);
  const size_t attend = (8 * static_cast<size_t>(a.A) + static_cast<size_t>(a.rows_per_feature) * a.n_keys) * sizeof(float) +
                        a.rows_per_feature * sizeof(int);
  const size_t smem = std::max(beam_select_smem_bytes(s.in_rows, s.beam, s.V), attend);
  if (smem > 227 * 1024) return static_cast<int>(cudaErrorInvalidValue);
  static size_t configured[64] = {};
  if (int rc = ensure_dynamic_smem(select_attend_kernel, smem, configured)) return rc;
  select_attend_kernel<<<s.n_neurons * kSelectCluster, kSelectThreads, smem, stream>>>(s, a);
  return last_err();
}

int launch_lm_finalize(const LmFinalizeArgs& a, cudaStream_t stream) {
  if (a.M == 0) return 0;
  lm_finalize_kernel<<<(a.M + 7) / 8, 256, 0, stream>>>(a);
  return last_err();
}

int launch_lm_input_table(const float* w_ih, const float* emb, int V, int E, int H, float* table, cudaStream_t stream) {
  lm_input_table_kernel<<<V, 256, E * sizeof(float), stream>>>(w_ih, emb, E, H, table);
  return last_err();
}

}  // namespace milan
