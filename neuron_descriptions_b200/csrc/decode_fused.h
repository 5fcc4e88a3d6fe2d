// The fused beam step (BASELINE north_star: "one kernel per beam step" for the non-GEMM work) and LM rerank.
//
// One step of `Decoder.step` x allennlp `BeamSearch` (src/milan/decoders.py:576-634, call site :467-484) is four
// launches:
//   1. attend_fused      cluster of 8 CTAs per neuron: attention scores + softmax (Attention.forward,
//                        decoders.py:57-73), token embedding, attenuate + gate, and the parent-state gather that
//                        allennlp does by reordering every state tensor with the backpointers — all written straight
//                        into the LSTM GEMM's operand rows
//   2. LSTM GEMM         conv_gemm_kernel<EPI_LSTM>: gates -> c', h' in the epilogue
//   3. head GEMM         conv_gemm_kernel<EPI_HEAD>: [W_out; W_q; W_g] h' -> logits + softmax partials, next step's
//                        attention query and feature gate
//   4. beam_select       cluster of 8 CTAs per neuron, one source row per warp: log-softmax normaliser from the
//                        partials, exact per-row top-`beam`; then CTA 0 of the cluster does the `beam x beam` merge,
//                        backpointers / history, and allennlp's all-ended early-exit flag
// Steps 4 (of step t) and 1 (of step t + 1) run as ONE launch, select_attend: a beam step costs three launches.
// Everything that depends on the parent row only (query, gate, h', c') is produced in the parent's row order and read
// through the backpointers by the consumer, so no state tensor is ever reordered in memory.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace milan {

constexpr int kAttendCluster = 8;      // CTAs per neuron in attend_fused
constexpr int kFusedMaxKeys = 16;      // keys held in registers (k = 15 exemplars)
constexpr int kSelectThreads = 256;
constexpr int kSelectCluster = 8;      // CTAs per neuron in beam_select: one source row per warp
constexpr int kSelectCandCap = 256;    // candidates per row kept by the fast path
constexpr int kSelectMaxGroups = 128;  // segment maxima ranked per row

struct AttendFusedArgs {
  const float* q;            // [parent rows][q_pitch]: W_q h' + b_q
  long long q_pitch;
  const float* gate;         // [parent rows][gate_pitch]: sigmoid(W_g h' + b_g)
  long long gate_pitch;
  const int* src_row;        // [R] parent row of each row; nullptr = identity (first step, greedy)
  const float* kh;           // [Bf * n_keys][A] = W_k f + b_k
  const float* features;     // [Bf][n_keys][F]
  const float* w_o;          // [A]
  float b_o;
  const float* embedding;    // [V][E]
  const long long* tokens;   // [R]
  const __nv_bfloat16* h_src_hi;  // parent h' planes [parent rows][h_src_pitch], or nullptr (h already in x)
  const __nv_bfloat16* h_src_lo;
  long long h_src_pitch;
  int R, rows_per_feature, n_keys, A, F, E, H;
  __nv_bfloat16* x_hi;       // LSTM operand rows [R][x_pitch]: [0,E) embedding | [E,E+F) gated features | [E+F,E+F+H) h
  __nv_bfloat16* x_lo;
  long long x_pitch;
  float* attn_ws;            // [R][n_keys] softmax weights (exchanged between the CTAs of a cluster)
  float* attn_out;           // optional copy, row pitch attn_pitch
  long long attn_pitch;
  const int* skip;
};
int launch_attend_fused(const AttendFusedArgs& a, cudaStream_t stream);

struct BeamSelectArgs {
  const float* logits;       // [rows][ld] raw logits (head GEMM)
  long long ld;
  const float2* partials;    // [rows][n_seg] (max, sum exp(x - max)) per 64-column segment
  int n_seg;
  int V;
  const long long* last_tokens;  // [rows] tokens fed into this step
  const float* last_lp;          // [rows] or nullptr (first step: 0)
  int n_neurons, in_rows, beam;  // rows = n_neurons * in_rows; in_rows = 1 (first step) or beam
  long long stop_index;
  float* cand_val;           // workspace [rows][beam]: each row's sorted top-`beam` (score incl. last_lp)
  int* cand_cls;             // workspace [rows][beam]
  long long* next_tokens;    // [n_neurons * beam]
  float* next_lp;            // [n_neurons * beam]
  int* backptr;              // [n_neurons * beam] global parent row
  int* hist_tok;             // this step's [n_neurons * beam]
  int* hist_bp;              // this step's beam-local parent
  const float* cur_lp;       // [n_neurons * beam] scores carried over once every beam has ended
  int* done_flag;            // read at entry (non-zero: all beams ended earlier); rewritten by the last CTA
  int* counters;             // [2], zero before the first step; left zero by every launch
};
int launch_beam_select(const BeamSelectArgs& a, cudaStream_t stream);
// beam_select of step t and attend_fused of step t + 1 in one launch (a.tokens / a.src_row must be s.next_tokens /
// s.backptr; a.skip is ignored: the early-exit flag is read once, through s.done_flag).
int launch_select_attend(const BeamSelectArgs& s, const AttendFusedArgs& a, cudaStream_t stream);
size_t beam_select_smem_bytes(int in_rows, int beam, int V);

// LM rerank: lm_scores[m] = sum over kept positions t of log p(seq[m][t] | ...) from the per-position softmax
// partials and target logits the LM head GEMMs left behind (LanguageModel.forward(reduce=True), lms.py:85-100,
// with its stop-mask off-by-one: the token after the first <stop> still counts).
struct LmFinalizeArgs {
  const float2* partials;    // [length][M][n_seg]
  const float* tgt_logit;    // [length][M]
  int n_seg;
  int M, length, beam, group_size;
  const long long* seqs;     // [M][length]
  const int* group_T;        // reference early-exit length per group of group_size neurons
  long long stop_index;
  float* lm_scores;          // [M]
};
int launch_lm_finalize(const LmFinalizeArgs& a, cudaStream_t stream);

// table[v][4u + g] = sum_e w_ih[g * H + u][e] * emb[v][e]: the LM's first LSTM layer sees its input only through
// this product, so the embedding lookup + K = E slice of the GEMM become one gathered add in the epilogue.
int launch_lm_input_table(const float* w_ih, const float* emb, int V, int E, int H, float* table, cudaStream_t stream);

}  // namespace milan
