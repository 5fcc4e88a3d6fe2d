/* milan_b200 — C ABI of the B200-native MILAN describe-neurons engine (libmilan_b200.so).
 *
 * The reference (evandez/neuron-descriptions) has no FFI: its boundary for this path is the Python object
 * protocol of `src/milan`. Each entry point below is what a binding for that path
 * would call, and cites the reference interface it replaces. The Python facade
 * (`neuron_descriptions_b200/milan/`) binds these with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; `milan_last_error()` returns a thread-local
 *     message for the last failure on the calling thread;
 *   - pointers named `d_*` are DEVICE pointers on the engine's GPU, `h_*` are HOST pointers; sizes are element
 *     counts unless stated; tensors are dense row-major with the shapes given;
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = default stream); calls are asynchronous on that
 *     stream unless they take host output pointers, in which case they synchronise the stream before returning;
 *   - token ids are int64 (the reference's torch.long), everything else float32 unless stated;
 *   - the caller owns every buffer it passes; the engine owns its weights and workspace.
 */
#ifndef MILAN_B200_H_
#define MILAN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct MilanEngine MilanEngine;

enum { MILAN_PRECISION_SPLIT = 0, /* bf16 hi/lo split operands, 3 MMAs per k-block: fp32-class results */
       MILAN_PRECISION_FAST = 1   /* plain bf16 operands */ };
enum { MILAN_DTYPE_U8 = 0, MILAN_DTYPE_F32 = 1 };
/* Backbone of the image encoder (torchvision graphs; src/milan/encoders.py:214-216,326-351). */
enum { MILAN_ENCODER_RESNET101 = 0, MILAN_ENCODER_RESNET50 = 1, MILAN_ENCODER_RESNET18 = 2,
       MILAN_ENCODER_RESNET34 = 3, MILAN_ENCODER_ALEXNET = 4 /* pyramid only, feature_size 1152 */ };
/* PyramidConvEncoder: masked spatial pooling of conv1 + layer1..4 -> one vector per image
 * (src/milan/encoders.py:286-320). SpatialConvEncoder: images * masks -> layer4 map -> 49 vectors per image
 * (src/milan/encoders.py:193-214). */
enum { MILAN_ENCODER_PYRAMID = 0, MILAN_ENCODER_SPATIAL = 1 };

/* Model dimensions (reference: Decoder.__init__, src/milan/decoders.py:233-323; LanguageModel.__init__,
 * src/milan/lms.py:20-56; PyramidConvEncoder('resnet101'), src/milan/encoders.py:251-284,346-350). */
typedef struct MilanConfig {
  int32_t vocab_size;        /* len(indexer): |vocab| + 4 specials */
  int32_t embedding_size;    /* 128 */
  int32_t hidden_size;       /* 512 */
  int32_t attention_size;    /* min(hidden, feature) = 512 */
  int32_t feature_size;      /* 3904 */
  int32_t start_index;       /* |vocab|     (src/utils/lang.py:242-245) */
  int32_t stop_index;        /* |vocab| + 1 (src/utils/lang.py:247-250) */
  int32_t has_encoder;       /* 1: ResNet-101 pyramid encoder weights will be provided */
  int32_t has_lm;            /* 1: 2-layer LSTM LM weights will be provided */
  int32_t lm_embedding_size; /* 128 */
  int32_t lm_hidden_size;    /* 512 */
  int32_t precision;         /* MILAN_PRECISION_* */
  int32_t max_images;        /* encoder micro-batch capacity in images (e.g. 16 neurons * 15) */
  int32_t max_neurons;       /* decoder capacity in neurons per call */
  int32_t max_beam;          /* <= 64 */
  int32_t max_keys;          /* exemplars per neuron, 15 */
  int32_t max_length;        /* decode length, 15 */
  int32_t encoder_arch;      /* MILAN_ENCODER_RESNET*; 0 = resnet101, the encoder of every shipped checkpoint */
  int32_t encoder_kind;      /* MILAN_ENCODER_PYRAMID (feature_size 3904 / 1024) or _SPATIAL (resnet18: 512) */
} MilanConfig;

const char* milan_version(void);
const char* milan_last_error(void);

/* Engine lifetime. Replaces: milan.pretrained() -> Decoder.load -> nn.Module tree
 * (src/milan/loaders.py:28-32, src/utils/serialize.py:221-269). One engine per (process, GPU). */
int milan_engine_create(const MilanConfig* config, int device, MilanEngine** out);
void milan_engine_destroy(MilanEngine* engine);

/* Hand the engine one tensor of a reference `Decoder.state_dict()` by its reference key name
 * (`encoder.encoder.model.layer3.4.conv2.weight`, `lstm.weight_ih`, `lm.lstm.weight_hh_l1`, ...; SURVEY.md
 * section 5). `h_data` is host fp32, `shape`/`ndim` its torch shape. Unknown names are ignored (the reference
 * loads with strict=False, serialize.py:250-251). */
int milan_engine_set_tensor(MilanEngine* engine, const char* name, const float* h_data, const int64_t* shape,
                            int32_t ndim);
/* Fold BN into the conv weights, re-lay-out and split every matrix for the tensor-core kernels, upload. */
int milan_engine_finalize(MilanEngine* engine);

/* PyramidConvEncoder.forward (src/milan/encoders.py:286-320) for n_images images.
 * d_images: (n,3,224,224) uint8 [0,255] or float32 [0,1]; d_masks: (n,1,224,224) uint8/float32, or NULL for
 * all-ones masks; d_features_out: (n, feature_size) for a pyramid encoder, (n, 49, feature_size) for a spatial one.
 * d_images (and d_masks of a spatial encoder) must be 16-byte aligned (vector loads). */
int milan_encode(MilanEngine* engine, const void* d_images, const void* d_masks, int32_t n_images, int32_t dtype,
                 float* d_features_out, void* stream);

/* Decoder.init_state (src/milan/decoders.py:548-574): d_features (B, n_keys, F) -> d_h, d_c (B, H). */
int milan_init_state(MilanEngine* engine, const float* d_features, int32_t B, int32_t n_keys, float* d_h,
                     float* d_c, void* stream);

/* Decoder.step (src/milan/decoders.py:576-634) on R rows. d_features holds R / rows_per_feature feature sets
 * (row r uses set r / rows_per_feature). d_h/d_c (R,H) are updated in place. If d_h_lm/d_c_lm (2,R,H_lm) are
 * non-NULL the LM is advanced too and predictions = log p - temperature * log p_lm (MI decoding).
 * d_predictions_out (R,V); d_attentions_out (R,n_keys) may be NULL. */
int milan_step(MilanEngine* engine, const float* d_features, int32_t n_keys, const int64_t* d_tokens, float* d_h,
               float* d_c, float* d_h_lm, float* d_c_lm, int32_t R, int32_t rows_per_feature, float temperature,
               float* d_predictions_out, float* d_attentions_out, void* stream);

/* Greedy / forced decoding loop of Decoder.forward (src/milan/decoders.py:430-463). mi != 0 -> MI decoding with
 * the LM. d_forced (B,length) int64 or NULL. Outputs: d_tokens_out (B,length) int64, d_scores_out (B),
 * d_predictions_out (B,length,V) or NULL, d_attentions_out (B,length,n_keys) or NULL. */
int milan_decode_greedy(MilanEngine* engine, const float* d_features, int32_t B, int32_t n_keys, int32_t length,
                        int32_t mi, float temperature, const int64_t* d_forced, int64_t* d_tokens_out,
                        float* d_scores_out, float* d_predictions_out, float* d_attentions_out, void* stream);

/* Beam search + optional LM rerank of Decoder.forward (src/milan/decoders.py:465-512; allennlp 2.10 BeamSearch
 * semantics). group_size = the reference DataLoader batch size (decoders.py:814): the
 * reference's early exit, and through it the LM mask, depends on which neurons share a batch.
 * mi != 0: MI beam decoding (the LM is advanced inside the loop and its state follows the beam backpointers,
 * decoders.py:170-175,192-196,624-630); incompatible with rerank (decoders.py:395-396).
 * Outputs: d_beam_tokens_out (B,beam,length) int64 (columns >= steps are <stop>), d_beam_scores_out (B,beam),
 * d_group_steps_out (ceil(B/group_size)) int32 = T the reference would return for each group,
 * and if rerank != 0: d_tokens_out (B,length), d_scores_out (B), d_lm_scores_out (B,beam) (may be NULL). */
int milan_decode_beam(MilanEngine* engine, const float* d_features, int32_t B, int32_t n_keys, int32_t length,
                      int32_t beam, int32_t group_size, int32_t rerank, int32_t mi, float temperature,
                      int64_t* d_beam_tokens_out, float* d_beam_scores_out, int32_t* d_group_steps_out,
                      int64_t* d_tokens_out, float* d_scores_out, float* d_lm_scores_out, void* stream);

/* LanguageModel.forward(inputs, reduce=True) (src/milan/lms.py:58-101): d_inputs (M, T1) int64 including the
 * leading <start>; d_out (M). */
int milan_lm_score(MilanEngine* engine, const int64_t* d_inputs, int32_t M, int32_t T1, float* d_out, void* stream);

/* LanguageModel.forward(inputs, reduce=False) (src/milan/lms.py:85-87): d_inputs (M, T) int64; d_out (M, T, V)
 * float32 = log-softmax of the LM's next-token distribution after each input token. */
int milan_lm_logprobs(MilanEngine* engine, const int64_t* d_inputs, int32_t M, int32_t T, float* d_out, void* stream);

/* End-to-end `Decoder.predict` equivalent on HOST buffers (src/milan/decoders.py:809-871 with
 * strategy='rerank' | 'beam' | 'greedy'): h_images (n,k,3,224,224) uint8, h_masks (n,k,1,224,224) uint8 (pinned
 * memory recommended). Chunks of whole reference batches flow through three streams: the exemplars of chunk i+1 are
 * copied while chunk i is encoded, and chunk i-1 is decoded (high-priority stream) under that encoder; results are
 * copied back once at the end.
 * strategy: 0 greedy (mi per `mi`), 1 beam, 2 rerank. Outputs (host): h_tokens_out (n,length) int64,
 * h_scores_out (n), h_steps_out (n) int32 = valid columns of each row (reference T of its group). */
int milan_describe_host(MilanEngine* engine, const uint8_t* h_images, const uint8_t* h_masks, int32_t n_neurons,
                        int32_t k, int32_t strategy, int32_t mi, int32_t length, int32_t beam, int32_t group_size,
                        float temperature, int64_t* h_tokens_out, float* h_scores_out, int32_t* h_steps_out,
                        void* stream);

/* The same pipeline with the exemplars already RESIDENT on the device: d_images (n,k,3,224,224) uint8, d_masks
 * (n,k,1,224,224) uint8 (16-byte aligned); outputs on the device: d_tokens_out (n,length) int64 (columns past a
 * group's early exit hold <stop>), d_scores_out (n). Asynchronous on `stream`; internally the decode of chunk i runs
 * on an engine-owned high-priority stream under the encoder of chunk i+1, and `stream` waits for it at the end. */
int milan_describe_device(MilanEngine* engine, const uint8_t* d_images, const uint8_t* d_masks, int32_t n_neurons,
                          int32_t k, int32_t strategy, int32_t mi, int32_t length, int32_t beam, int32_t group_size,
                          float temperature, int64_t* d_tokens_out, float* d_scores_out, void* stream);

/* ---- Stage-1 exemplar statistics (src/exemplars/compute.py:27-246). Stateless: the
 * caller owns every buffer and runs the described network itself; d_acts is its activation tensor (B, U, P) float32
 * (B images of the batch, U units, P = H*W positions).
 *   milan_tally_topk       RunningTopK.add (src/deps/netdissect/runningstats.py:58-94) on the spatial max of each
 *                          image: d_top_vals (U,k) descending / d_top_ids (U,k) dataset indices, initialised by the
 *                          caller to -inf / -1; dataset index of image b = base_index + b; k <= 64, B <= 1024
 *   milan_tally_samples    RunningQuantile.add in its exact regime (runningstats.py:343-386): appends the B*P
 *                          activations of every unit to d_samples (U, capacity) at column `count`
 *   milan_tally_hist       beyond that regime: d_hist (U, 65536) uint32 += histogram of the upper 16 bits of the
 *                          order-preserving float key (deterministic; histograms of several GPUs add)
 *   milan_quantile_exact   RunningQuantile.quantiles(q) (runningstats.py:557-580) on n <= 8192 kept samples
 *   milan_quantile_hist    the same estimator read from the histogram (linear inside the selected bin)
 *   milan_activation_masks ImageVisualizer.pytorch_mask (src/deps/netdissect/imgviz.py:185-198) with the default
 *                          grid of upsample.upsample_grid (upsample.py:127-157): d_maps (n,H,W) -> d_masks (n,S,S) of
 *                          0/1 bytes, mask = bilinear(zeros padding, align_corners) > d_levels[i] */
int milan_tally_topk(const float* d_acts, int32_t B, int32_t U, int32_t P, int64_t base_index, int32_t k,
                     float* d_pooled_scratch /* (B, U) floats */, float* d_top_vals, int64_t* d_top_ids, void* stream);
int milan_tally_samples(const float* d_acts, int32_t B, int32_t U, int32_t P, float* d_samples, int64_t capacity,
                        int64_t count, void* stream);
int milan_tally_hist(const float* d_acts, int32_t B, int32_t U, int32_t P, uint32_t* d_hist, void* stream);
int milan_quantile_exact(const float* d_samples, int32_t U, int64_t capacity, int64_t n, float q, float* d_levels,
                         void* stream);
int milan_quantile_hist(const uint32_t* d_hist, int32_t U, int64_t n, float q, float* d_levels, void* stream);
int milan_activation_masks(const float* d_maps, const float* d_levels, int32_t n, int32_t H, int32_t W, int32_t S,
                           uint8_t* d_masks, void* stream);

/* Counters for bench.py: kernels launched by this library since process start; device ms spent in the encoder
 * convolution kernels inside the last milan_describe_host / milan_encode call when profiling is enabled. */
int64_t milan_launch_count(void);
int milan_set_profiling(MilanEngine* engine, int32_t enabled);
int milan_get_profile(MilanEngine* engine, float* conv_ms, float* encoder_ms, float* decoder_ms, int64_t* conv_launches);

#ifdef __cplusplus
}
#endif
#endif /* MILAN_B200_H_ */
