// Launchers for the non-GEMM kernels of the attention-LSTM decoder, beam search and LM rerank (decoder.cu).
// Reference semantics: Decoder.init_state/step (src/milan/decoders.py:548-634), allennlp BeamSearch
// (call site decoders.py:467-484), LanguageModel.forward(reduce=True) (src/milan/lms.py:85-100),
// rerank (decoders.py:495-512).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace milan {

// dst_hi/lo[m][col0 + k] = split(src[m][k]); lo may be nullptr (fast mode).
int launch_split_rows(const float* src, long long src_pitch, __nv_bfloat16* dst_hi, __nv_bfloat16* dst_lo,
                      long long dst_pitch, int M, int K, cudaStream_t stream);

// pooled[b][:] = mean_k features[b][k][:]  -> hi/lo (decoders.py:564).
int launch_mean_keys(const float* features, int B, int n_keys, int F, __nv_bfloat16* dst_hi, __nv_bfloat16* dst_lo,
                     cudaStream_t stream);

// h = tanh(pre[:, :H]), c = tanh(pre[:, H:]); h also written as hi/lo at (h_hi, pitch). (decoders.py:565)
int launch_init_finish(const float* pre, int B, int H, float* h, float* c, __nv_bfloat16* h_hi,
                       __nv_bfloat16* h_lo, long long h_pitch, cudaStream_t stream);

struct AttendArgs {
  const float* qg;        // [R][qg_pitch]: cols [0,A) = W_q h + b_q, cols [A, A+F) = W_g h + b_g
  long long qg_pitch;
  const float* kh;        // [Bf*n_keys][A] = W_k f + b_k
  const float* features;  // [Bf][n_keys][F]
  const float* w_o;       // [A]
  float b_o;
  const float* embedding; // [V][E]
  const long long* tokens;  // [R]
  int R, rows_per_feature, n_keys, A, F, E;
  __nv_bfloat16* x_hi;    // [R][x_pitch]: cols [0,E) embedding, [E, E+F) gated features
  __nv_bfloat16* x_lo;
  long long x_pitch;
  float* attn_out;        // [R][attn_pitch] or nullptr
  long long attn_pitch;
  const int* skip;        // device flag: non-zero -> no-op (all beams ended); may be nullptr
  float* attn_ws;         // workspace [R][n_keys]
  const float* acc_ws;    // workspace [R][F], only needed when n_keys > 16
};
int launch_attend(const AttendArgs& a, cudaStream_t stream);

// LSTM pointwise on pre-activations gates[R][4H] (order i,f,g,o; biases already added):
// c_new = sig(f)*c + sig(i)*tanh(g); h_new = sig(o)*tanh(c_new). h_new is written as fp32 (optional) and as
// hi/lo to up to two destinations (pointer + pitch each).
struct LstmPointArgs {
  const float* gates;
  const float* c_in;
  float* c_out;
  float* h_out;  // optional fp32
  __nv_bfloat16* h_hi[2];
  __nv_bfloat16* h_lo[2];
  long long h_pitch[2];
  int R, H;
  const int* skip;
};
int launch_lstm_point(const LstmPointArgs& a, cudaStream_t stream);

// Embedding lookup -> hi/lo rows: dst[m][0:E] = split(table[tokens[m]]).
int launch_embed_rows(const float* table, const long long* tokens, int M, int E, __nv_bfloat16* dst_hi,
                      __nv_bfloat16* dst_lo, long long dst_pitch, cudaStream_t stream, const int* skip = nullptr);

constexpr int kMaxBeam = 64;

// Per-row log-softmax (+ optional MI: pred = logp - temperature * logp_lm) followed by one of:
//   greedy: next[r] = argmax, score[r] += pred[argmax]; predictions row optionally stored
//   beam  : finished rows (last token == stop) emit only (stop, last_lp); others their top-`beam` candidates,
//           cand_val = last_lp + pred, sorted descending (ties: lower class index first).
struct RowArgs {
  const float* logits;     // [R][ld]
  const float* logits_lm;  // [R][ld] or nullptr
  long long ld;
  int R, V;
  float temperature;
  // greedy
  long long* next_tokens;  // [R] or nullptr
  float* scores;           // [R] accumulated
  float* pred_out;         // predictions base for this step or nullptr; row r at pred_out + r*pred_pitch
  long long pred_pitch;
  const long long* forced; // [R] forced next tokens (forced decoding) or nullptr
  // beam
  int beam;                // 0 -> greedy mode
  const long long* last_tokens;  // [R] tokens fed into this step
  const float* last_lp;    // [R] or nullptr (treated as 0)
  long long stop_index;
  float* cand_val;         // [R][beam]
  int* cand_cls;           // [R][beam]
  const int* skip;
};
int launch_row_logsoftmax(const RowArgs& a, cudaStream_t stream);

// Per neuron: merge the sorted candidate lists of its `in_rows` source rows into the next beam (sorted desc).
struct MergeArgs {
  const float* cand_val;
  const int* cand_cls;
  int n_neurons, in_rows, beam;
  long long* next_tokens;  // [n_neurons*beam]
  float* next_lp;          // [n_neurons*beam]
  int* backptr;            // [n_neurons*beam] global source row
  int* hist_tok;           // this step's [n_neurons*beam]
  int* hist_bp;            // this step's [n_neurons*beam] (beam-local parent index)
  const int* skip;         // non-zero: every beam has ended -> emit <stop> / identity / unchanged scores
  const float* cur_lp;     // [n_neurons*beam] current scores (used when skipping)
  long long stop_index;
};
int launch_beam_merge(const MergeArgs& a, cudaStream_t stream);

// dst rows <- src rows[backptr] for the recurrent state (hi/lo bf16 planes and fp32 cell).
struct GatherArgs {
  const int* backptr;  // nullptr = identity
  int R, H;
  const __nv_bfloat16* src_hi; const __nv_bfloat16* src_lo; long long src_pitch;
  __nv_bfloat16* dst_hi; __nv_bfloat16* dst_lo; long long dst_pitch;
  const float* c_src; float* c_dst;  // [R][H] or nullptr
  const int* skip;
};
int launch_gather_state(const GatherArgs& a, cudaStream_t stream);

// Backtrack the beam history into sequences [n_neurons][beam][length] and compute, per reference batch group,
// the number of steps the reference would have taken before its early exit (decoders.py:483-484 + allennlp).
struct BacktrackArgs {
  const int* hist_tok;  // [length][n_neurons*beam]
  const int* hist_bp;   // [length][n_neurons*beam]
  int n_neurons, beam, length, group_size;
  long long stop_index;
  long long* seqs;      // [n_neurons*beam][length]
  int* group_T;         // [ceil(n_neurons/group_size)]
};
int launch_backtrack(const BacktrackArgs& a, cudaStream_t stream);

// LM scoring bookkeeping for position t: input token for the LM step and masked accumulation of the target
// log-prob with the reference's stop-mask off-by-one (lms.py:93-96).
int launch_lm_inputs(const long long* seqs, int M, int length, int t, long long start_index, long long* inputs,
                     cudaStream_t stream, const int* skip = nullptr);
struct LmAccumArgs {
  const float* logits;  // [M][ld]
  long long ld;
  int M, V, length, t, beam, group_size;
  const long long* seqs;  // [M][length]
  const int* group_T;     // per group of group_size neurons (M = n_neurons*beam rows)
  long long stop_index;
  float* lm_scores;       // [M] accumulated
  const int* skip;
};
int launch_lm_accumulate(const LmAccumArgs& a, cudaStream_t stream);

// scores = beam_lp - temperature * lm; argmax over the beam (first max); copy the chosen sequence.
struct RerankArgs {
  const float* beam_lp; const float* lm_scores; float temperature;
  int n_neurons, beam, length;
  const long long* seqs;   // [n_neurons*beam][length]
  long long* out_tokens;   // [n_neurons][length]
  float* out_scores;       // [n_neurons]
  int* out_index;          // [n_neurons] or nullptr
};
int launch_rerank_select(const RerankArgs& a, cudaStream_t stream);

// *flag = 1 iff every token equals stop (allennlp: `if (last_predictions == end).all(): break`).
int launch_check_done(const long long* tokens, int n, long long stop_index, int* flag, cudaStream_t stream);
// lm_skip[t] = (t >= max_g group_T[g]) for t < length: LM positions no group needs.
int launch_lm_skip(const int* group_T, int groups, int length, int* lm_skip, cudaStream_t stream);
int launch_fill_i64(long long* dst, long long value, int n, cudaStream_t stream);
int launch_fill_f32(float* dst, float value, int n, cudaStream_t stream);

}  // namespace milan
