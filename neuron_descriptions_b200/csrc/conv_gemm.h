// Implicit-GEMM convolution / GEMM on the 5th-gen tensor cores (tcgen05 + TMEM), fed by TMA.
//
// One kernel covers every matrix product on the MILAN hot path:
//   * ResNet-101 1x1 / 3x3, stride 1 / 2 convolutions of the pyramid encoder
//     (reference: torchvision resnet101 called at src/milan/encoders.py:298),
//   * the 7x7 stem as a GEMM over an im2col matrix,
//   * the decoder / LM linear layers (src/milan/decoders.py:612-621, src/milan/lms.py:85-87).
//
// A (activations) is NHWC bf16, addressed through 4-D TMA boxes (c, w, h, n); each filter tap is the same box
// shifted by (dw, dh), with out-of-bounds pixels zero-filled by TMA (= conv padding). Stride-2 convs read one of
// four parity planes of the input (same memory, doubled strides) so every load stays a unit-stride box.
// B (weights) is [Cout][taps*Cin] bf16, K-major.
//
// Precision: in SPLIT mode every fp32 operand is carried as a (hi, lo) bf16 pair and each k-block issues
// hi*hi + lo*hi + hi*lo into the fp32 TMEM accumulator (~2^-16 relative operand error, i.e. fp32-class results
// on the bf16 tensor pipe). In FAST mode only the hi planes are used (plain bf16).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace milan {

constexpr int kMaxTaps = 25;  // up to 5x5 filters (alexnet features.3)
constexpr int kGemmBlockM = 128;
constexpr int kGemmBlockK = 64;

enum EpilogueKind : int {
  EPI_BF16 = 0,  // out_hi(/out_lo) bf16 NHWC, + bias, + residual, optional ReLU
  EPI_F32 = 1,   // out_f32 row-major [pixel][ldc], + bias
  EPI_LSTM = 2,  // LSTM cell in the epilogue: the GEMM's columns are gate pre-activations (FusedEpilogue)
  EPI_HEAD = 3,  // everything that reads h': vocabulary logits + softmax partials | attention query | feature gate
};

// Arguments of the fused decoder epilogues (flat GEMMs only: one output row per A row).
//
// EPI_LSTM (Decoder.step's LSTMCell, src/milan/decoders.py:619; LanguageModel's nn.LSTM, src/milan/lms.py:50-54):
//   the weight rows are interleaved so that column 4u + g is gate g (i, f, g, o) of hidden unit u; an epilogue
//   thread therefore owns all four gates of 8 units per 32-column chunk and finishes the cell in registers:
//   c' = sig(f) c + sig(i) tanh(g), h' = sig(o) tanh(c'). The pre-activation never reaches memory.
// EPI_HEAD (the three linears applied to h': output.1, attend.query_to_hidden, feature_gate.0,
//   decoders.py:612-621; LM: output.0, lms.py:55-56): column tiles [0, vocab_tiles) are vocabulary logits, the
//   next q_tiles are the attention query, the rest the feature gate. Each thread reduces its 64 logits of a row to
//   a (max, sum exp(x - max)) partial; log-softmax is finished by the consumer from 2 * vocab_tiles partials per row.
struct FusedEpilogue {
  // ---- EPI_LSTM
  int hidden;                // H (N = 4H)
  const float* c_in;         // [rows][H]; row src_row[r] when src_row != nullptr
  float* c_out;              // [rows][H] (may alias c_in only when src_row == nullptr)
  const int* src_row;        // beam backpointers: the parent row whose cell state this row continues
  __nv_bfloat16* h_hi;       // h' as fp16 (hi, lo) planes, row pitch h_pitch elements
  __nv_bfloat16* h_lo;
  long long h_pitch;
  float* h_f32;              // optional [rows][H]
  const float* add_table;    // optional: pre-activation += add_table[index(r)][col]  (an embedding folded through
  long long add_pitch;       //           W_ih: the LM's first layer), index(r) = add_index[r * add_stride] or add_const
  const long long* add_index;
  long long add_stride;
  long long add_const;
  // ---- EPI_HEAD
  int vocab_tiles, q_tiles;  // column tiles of 128; gate tiles = n_tiles - vocab_tiles - q_tiles
  int vocab;                 // valid vocabulary columns
  float* logits;             // [rows][ld_logits] or nullptr (not stored)
  long long ld_logits;
  float2* partials;          // [rows][2 * vocab_tiles]: (max, sum exp(x - max)) of each thread's 64 columns
  float* q_out;              // [rows][q_pitch]
  long long q_pitch;
  int q_cols;                // valid query columns (A)
  float* g_out;              // [rows][g_pitch] = sigmoid(gate pre-activation)
  long long g_pitch;
  int gate_cols;             // valid gate columns (F)
  const long long* target;   // optional: tgt_logit[r] = logit[r][target[r * target_stride]] (LM scoring)
  long long target_stride;
  float* tgt_logit;
};

struct alignas(64) ConvGemmParams {
  CUtensorMap tmap_a[2][4];  // [hi|lo][parity plane]
  CUtensorMap tmap_b[2];     // [hi|lo]
  CUtensorMap tmap_b_half[2];  // [hi|lo] boxes of block_n / 2 rows: what ONE CTA of a cta_group::2 pair loads (conv_gemm_pair.cu)
  int has_b_half;
  CUtensorMap tmap_out[2];   // [hi|lo] EPI_BF16: output tensor, box (64, box_w, box_h, box_n), TMA store
  CUtensorMap tmap_res[2];   // [hi|lo] EPI_BF16 + residual: same geometry as tmap_out, TMA load
  int stem_mode;             // 1: A is the raw-row map of the 7x7/2 stem, windows formed by the MMA descriptor (see build_stem_params)
  int has_res;               // residual add in the epilogue
  int block_k;               // k-block width in elements: 64 (default, 128B swizzle) or 32 (stem, 64B swizzle)
  int kb_per_chunk;          // k-blocks accumulated in TMEM before promotion to fp32 registers (0 = all)
  int fp16_operands;         // 1: A/B planes hold IEEE fp16 bit patterns (decoder GEMMs); 0: bf16 (encoder)
  int box_w, box_h, box_n;   // M tile = box_w*box_h*box_n (<=128) output pixels
  int tiles_w, tiles_h, tiles_n;
  int out_w, out_h, out_n;   // output extents (pixels / images)
  int n_tiles;               // ceil(cout / BLOCK_N)
  int cout;
  int cin;                   // channels per tap (multiple of 64)
  int num_taps;
  int8_t tap_plane[kMaxTaps], tap_dw[kMaxTaps], tap_dh[kMaxTaps];
  int16_t tap_cb[kMaxTaps];  // k-blocks of tap t (0 = cin / block_k): taps may read tensors of different depth
  uint32_t a_box_bytes;      // bytes one A box delivers
  const float* bias;         // [cout] or nullptr
  const __nv_bfloat16* res_hi;
  const __nv_bfloat16* res_lo;
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  float* out_f32;
  long long ldc;             // row pitch (elements) of the output / residual
  int relu;
  FusedEpilogue fe;          // EPI_LSTM / EPI_HEAD only
};

// block_n: 64 or 128. split: hi/lo planes (1) or hi only (0). Returns cudaError_t as int.
// EPI_BF16 kernels store through tmap_out (and read the residual through tmap_res when p.has_res).
// skip_flag (device, may be nullptr): when *skip_flag != 0 at kernel start the launch is a no-op — used by the
// beam-search loop to drop the GEMMs of steps after every beam has ended, without a host round trip.
int launch_conv_gemm(const ConvGemmParams& p, int block_n, int split, int epilogue, int num_sms,
                     cudaStream_t stream, const int* skip_flag = nullptr);
// CTA-pair variant (conv_gemm_pair.cu): BLOCK_N = 128, split mode, bf16 output, 64-wide k-blocks; p.has_res selects
// the residual form (cout must be a multiple of 128).
int launch_conv_gemm_pair(const ConvGemmParams& p, int num_sms, cudaStream_t stream, const int* skip_flag);
void count_conv_launch();
// Kernel launches performed by this library since process start (for bench.py's gpu_launches).
long long conv_gemm_launch_count();   // tcgen05 conv/GEMM kernel only
long long total_launch_count();       // every kernel of the library
void note_launch(int n = 1);

// Host helpers to build tensor maps (driver entry point resolved at runtime; no libcuda link dependency).
// 4-D bf16 map: dims (c, w, h, n), byte strides for w/h/n, box (64, bw, bh, bn), 128B swizzle, zero OOB fill.
int make_tmap_4d(CUtensorMap* out, const void* base, uint64_t c, uint64_t w, uint64_t h, uint64_t n,
                 uint64_t stride_w_bytes, uint64_t stride_h_bytes, uint64_t stride_n_bytes, uint32_t box_w,
                 uint32_t box_h, uint32_t box_n);
// 2-D bf16 map: dims (k, rows), row pitch bytes, box (64, box_rows).
int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t k, uint64_t rows, uint64_t pitch_bytes,
                 uint32_t box_rows, int block_k = kGemmBlockK);
// Generic bf16 map of `rank` dims (dims[0] innermost; strides_bytes[i] is the stride of dim i+1), 128B swizzle.
int make_tmap_nd(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                 const uint32_t* box, int swizzle_bytes = 128);
const char* tmap_last_error();

// Pick an M-tile box (bw, bh, bn) with bw*bh*bn <= 128 minimising the number of tiles for a WxHxN output.
void choose_box(int W, int H, int N, int* bw, int* bh, int* bn);

}  // namespace milan

namespace milan {

// Geometry of one convolution (pad = ksize/2, NHWC activations, weights [Cout][ksize*ksize*Cin]).
// ksize 1 with stride 1 is a plain GEMM over M = N*H*W rows (also used for the decoder linears with H=W=1).
struct ConvDesc {
  int N, H, W;   // input extents
  int Cin, Cout;
  int ksize;     // 1, 3 or 5 (pad = ksize/2)
  int stride;    // 1 or 2
};

struct ConvIO {
  const __nv_bfloat16* in_hi;
  const __nv_bfloat16* in_lo;   // nullptr in FAST mode
  const __nv_bfloat16* w_hi;    // [Cout][ksize*ksize*Cin]
  const __nv_bfloat16* w_lo;
  const float* bias;            // padded to a multiple of 128 floats, or nullptr
  const __nv_bfloat16* res_hi;  // same shape as the output, or nullptr
  const __nv_bfloat16* res_lo;
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  float* out_f32;               // EPI_F32 only
  long long ldc;                // output row pitch in elements (0 -> Cout)
  int relu;
};

// Fills `p` (tensor maps, tiling, taps) for the convolution; returns 0 on success. *block_n receives the
// N-tile width the kernel must be launched with.
int build_conv_params(ConvGemmParams* p, const ConvDesc& d, const ConvIO& io, int split, int* block_n);

// Two 1x1 convolutions summed in ONE accumulator: out = W_a * a + W_b * b(strided) + bias (+ReLU). This is the tail
// of the first bottleneck of a ResNet stage, relu(bn3(conv3(t)) + bn_ds(downsample(x))) (torchvision Bottleneck,
// called at src/milan/encoders.py:298): the downsample branch never round-trips HBM as a separate tensor.
// `a` is [N][Ho][Wo][Ca] (the 3x3 conv's output), `b` is [N][Ho*stride][Wo*stride][Cb] sampled at (stride*h,
// stride*w); w is [Cout][Ca + Cb] (both filters BN-folded, concatenated along K), bias the sum of both shifts.
struct DualConvIO {
  const __nv_bfloat16* a_hi; const __nv_bfloat16* a_lo;
  const __nv_bfloat16* b_hi; const __nv_bfloat16* b_lo;
  const __nv_bfloat16* w_hi; const __nv_bfloat16* w_lo;
  const float* bias;
  __nv_bfloat16* out_hi; __nv_bfloat16* out_lo;
  int relu;
};
int build_dual_1x1_params(ConvGemmParams* p, int N, int Ho, int Wo, int Ca, int Cb, int Cout, int stride,
                          const DualConvIO& io, int split, int* block_n);

// The 7x7 stride-2 stem as an implicit GEMM without im2col: the input is the zero-padded NHWC4 image
// [N][232][232][4] (pixel (ih, iw) at (ih + 3, iw + 4); channel 3 = 0) and the k-block of filter row r is the
// 64-byte window {8 pixels x 4 channels} of padded row 2*oh + r starting at pixel 2*ow. The windows are never
// materialised: a tile is 8 x 16 output pixels, TMA loads the raw 22-pixel (176 B) segments of the 16 padded rows a
// filter row touches through a 4-D map (row element, row parity, row pair, n; no swizzle), and the MMA's shared-memory
// descriptor reads them as overlapping windows (make_smem_desc_stem_rows in ptx.cuh). block_k = 32; window pixel 0
// carries zero weights. Weights: [64][7*32] (pack_stem_weights), 64-byte swizzle, resident in shared memory.
// Output: raw conv1 [N][112][112][64] through tmap_out. block_n is 64.
constexpr int kStemPadH = 232;
constexpr int kStemPadW = 232;
constexpr int kStemBlockK = 32;
constexpr int kStemKTotal = 7 * kStemBlockK;
int build_stem_params(ConvGemmParams* p, int N, const __nv_bfloat16* img_hi, const __nv_bfloat16* img_lo,
                      const __nv_bfloat16* w_hi, const __nv_bfloat16* w_lo, __nv_bfloat16* out_hi,
                      __nv_bfloat16* out_lo, int split);
// w: torchvision conv1.weight [64][3][7][7] -> packed [64][224] in the k order build_stem_params expects.
void pack_stem_weights(const float* w, float* packed);

// Plain GEMM out[M][ldc] (fp32) = A[M][K] * W[N][K]^T + bias, A given as bf16 hi/lo planes with row pitch
// a_pitch (elements), W contiguous [N][K]. K must be a multiple of 64. Launch with block_n = 128, EPI_F32.
int build_gemm_params(ConvGemmParams* p, long long M, int K, int N, const __nv_bfloat16* a_hi,
                      const __nv_bfloat16* a_lo, long long a_pitch, const __nv_bfloat16* w_hi,
                      const __nv_bfloat16* w_lo, const float* bias, float* out, long long ldc, int split,
                      int fp16_operands = 0);
// Same with the A rows given as TWO matrices side by side, A = [A0 | A1] (K = K0 + K1, both multiples of 64): the
// second LSTM layer of the LM reads [h of layer 0 | its own h] without either being copied next to the other.
int build_gemm2_params(ConvGemmParams* p, long long M, int K0, int K1, int N, const __nv_bfloat16* a0_hi,
                       const __nv_bfloat16* a0_lo, long long a0_pitch, const __nv_bfloat16* a1_hi,
                       const __nv_bfloat16* a1_lo, long long a1_pitch, const __nv_bfloat16* w_hi,
                       const __nv_bfloat16* w_lo, const float* bias, int split, int fp16_operands = 0);

}  // namespace milan
