/*
 * vcla.h -- C ABI of the B200-native VisualCLA multimodal forward path (libvcla.so).
 *
 * This is the drop-in boundary for ONE hot path of airaria/Visual-Chinese-LLaMA-Alpaca:
 *   image + prompt -> CLIP-ViT-L/14 -> post_layernorm -> 6-layer Resampler -> projector
 *   -> splice into the text embeddings -> LLaMA-7B prefill + KV-cache greedy decode.
 * The reference is pure Python and has no FFI of its own; each entry point
 * below names the reference code it replaces (paths relative to the reference repo root, `HF:` =
 * transformers 5.5.0).  The Python package `visualcla` (visual-chinese-llama-alpaca_b200/visualcla)
 * binds these with ctypes and exposes the reference's own API (VisualCLAModel.generate/.forward,
 * chat, get_model_and_tokenizer_and_processor).  See INTEGRATION.md for the binding stub.
 *
 * Conventions: plain pointers and sizes, no torch / C++ types; every function returns 0 on success,
 * non-zero on failure with a message retrievable through vcla_last_error() (thread-local); nothing
 * throws across the ABI.  "dev" pointers are CUDA device pointers on the context's device; the
 * library never frees caller memory.  One context per GPU, calls externally serialised.  Every call
 * that takes a `stream` only enqueues work on it (no host synchronisation) unless documented.
 */
#ifndef VCLA_H_
#define VCLA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vcla_ctx vcla_ctx;
typedef void* vcla_stream; /* cudaStream_t */

/* element types of caller buffers */
enum { VCLA_F32 = 0, VCLA_F16 = 1, VCLA_BF16 = 2 };

/* image layouts of the prompt (models/visualcla/modeling_visualcla.py:290-305 / :356-370) */
enum {
  VCLA_TEXT_ONLY = 0,       /* pixel_values=None: multimodal_embeds = input_embeds (:378-380) */
  VCLA_IMAGE_AT_HEAD = 1,   /* [e0,e1, img x nq, e2...]  (:291/:357); S = T + nq */
  VCLA_IMAGE_PLACEHOLDER = 2 /* image rows replace the nq <img_token> slots after <img> (:293-305); S = T */
};

/* Shapes of the path.  Mirrors VisualCLAConfig / VisualResamplerConfig / the HF CLIP + LLaMA configs
 * (models/visualcla/configuration_visualcla.py:10-39, modeling_visual_resampler.py:90-129). */
typedef struct {
  /* CLIP-ViT (HF:models/clip/modeling_clip.py) */
  int v_hidden, v_layers, v_heads, v_ffn, v_patch, v_image;
  float v_eps;
  /* Resampler (models/visualcla/modeling_visual_resampler.py) */
  int r_hidden, r_layers, r_heads, r_ffn, r_queries;
  float r_eps;
  /* LLaMA (HF:models/llama/modeling_llama.py) */
  int t_hidden, t_layers, t_heads, t_ffn, t_vocab;
  float t_eps, rope_theta;
  /* capacity of this context */
  int max_batch;           /* sequences resident at once (per GPU) */
  int max_seq;             /* prompt + generated tokens per sequence */
  int max_prefill_tokens;  /* max B*S of one prefill call */
  int page_tokens;         /* tokens per KV-cache page (default 64 if 0) */
} vcla_config;

const char* vcla_last_error(void);
const char* vcla_version(void);

/* ---- lifetime -------------------------------------------------------------------------------- */
/* Allocates the weight arena, paged KV cache and activation buffers on the current CUDA device.
 * Replaces model construction: VisualCLAModel.__init__ (modeling_visualcla.py:70-108). */
int vcla_create(const vcla_config* cfg, vcla_ctx** out);
void vcla_destroy(vcla_ctx* ctx);
int vcla_get_config(const vcla_ctx* ctx, vcla_config* out);
/* bytes of device memory held by the context (weights, kv, activations) */
int vcla_memory_bytes(const vcla_ctx* ctx, int64_t* weights, int64_t* kv, int64_t* activations);

/* ---- weights ---------------------------------------------------------------------------------- */
/* The logical tensors are addressed by the reference's own state-dict names
 * (VisualCLAModel.state_dict(): "vision_model.vision_model.*", "visual_resampler.*",
 * "image_projection_layer.*", "text_model.model.*", "text_model.lm_head.weight"; merged checkpoint
 * layout scripts/merge_llama_with_visualcla_lora.py:92-97, read back at modeling_visualcla.py:141-179). */
int vcla_weight_count(const vcla_ctx* ctx);
/* kind: 0 = matrix stored bf16, 1 = vector/table stored f32.  shape has up to 4 dims. */
int vcla_weight_info(const vcla_ctx* ctx, int index, const char** name, int64_t shape[4], int* ndim, int* kind);
/* Copy + repack one tensor into the arena (fused QKV, gate/up interleave, K padding).  `src` is a host
 * pointer (on_device = 0) or device pointer (on_device = 1) to the contiguous tensor in `dtype` holding `numel`
 * elements; the call fails (nothing is read) unless numel equals the element count of the named tensor, so a
 * checkpoint whose shape disagrees with the config is an error, never an out-of-bounds read.
 * Replaces from_merged_pretrained's loading (modeling_visualcla.py:120-181).  Synchronises the stream. */
int vcla_load_weight(vcla_ctx* ctx, const char* name, const void* src, int dtype, int64_t numel, int on_device,
                     vcla_stream stream);
/* Copy one logical tensor back to host: bf16 for kind 0, f32 for kind 1 (state_dict() equivalent). */
int vcla_read_weight(vcla_ctx* ctx, const char* name, void* dst_host, vcla_stream stream);
/* Deterministic synthetic weights: w = mean + std * IrwinHall4(hash(name, seed, index)), bit-identical
 * to oracle/visualcla_oracle.py:hash_normal_bf16 (std/mean per tensor as in oracle weight_specs). */
int vcla_init_synthetic(vcla_ctx* ctx, uint32_t seed, vcla_stream stream);

/* ---- the hot path ----------------------------------------------------------------------------- */
/* Drop all sequences (KV cache lengths -> 0, every KV page back on the free stack). */
int vcla_reset(vcla_ctx* ctx, vcla_stream stream);

/* Paged KV cache (the reference's DynamicCache, HF:cache_utils.py:88-120, re-designed): physical pages of `page_tokens`
 * tokens are handed to sequences on demand by a device-side allocator that runs inside the stream / the decode CUDA graph
 * (pages are assigned round-robin as sequences grow, so one sequence's pages are NOT contiguous; every kernel goes through
 * the page table).  vcla_decode_* fail with an error instead of running past max_seq.
 *   vcla_kv_geometry     pages per sequence (table row length), pages in the pool, tokens per page
 *   vcla_kv_read_pages   synchronous copy to the host of the page table (max_batch x pages_per_seq int32), the pages owned
 *                        per sequence (max_batch int32) and {free pages, exhausted flag} (2 int32); any pointer may be NULL
 *   vcla_kv_debug_shuffle  test hook: permute the order in which free pages are handed out (then resets the context) */
int vcla_kv_geometry(const vcla_ctx* ctx, int* pages_per_seq, int* total_pages, int* page_tokens);
int vcla_kv_read_pages(vcla_ctx* ctx, int32_t* table_host, int32_t* npages_host, int32_t* state_host);
int vcla_kv_debug_shuffle(vcla_ctx* ctx, uint32_t seed);

/* pixels (B,3,I,I) NCHW -> image embeddings (B, r_queries, t_hidden), kept inside the context for the
 * next vcla_prefill and optionally copied to `out_dev_f32`.  Replaces
 *   vision_model(pixel_values) -> post_layernorm -> visual_resampler -> image_projection_layer
 * (modeling_visualcla.py:283-288 / :349-354; HF:models/clip/modeling_clip.py:667-692;
 *  modeling_visual_resampler.py:609-737). */
int vcla_vision_encode(vcla_ctx* ctx, const void* pixels_dev, int pixel_dtype, int B, float* out_dev_f32, vcla_stream stream);

/* Prefill B prompts of T tokens each (equal length, unpadded; position_ids = arange).
 *   ids_dev        int64 (B,T) device
 *   image_mode     VCLA_TEXT_ONLY / VCLA_IMAGE_AT_HEAD / VCLA_IMAGE_PLACEHOLDER
 *   img_row_dev    int32 (B) device: first row of the image block inside each sequence (index of <img> + 1),
 *                  -1 = no image for this sample; ignored for TEXT_ONLY; for AT_HEAD pass NULL (row 2)
 *   left_pad_dev   int32 (B) device or NULL: number of LEFT padding tokens per sequence (attention_mask = [0]*p + [1]*(T-p),
 *                  what HF generate expects for batched prompts of different lengths): pad keys are masked, the KV cache
 *                  holds only the real tokens.  pos_from_mask != 0: RoPE positions count real tokens from 0
 *                  (HF generate: position_ids = attention_mask.cumsum(-1) - 1); 0: positions are arange(S) (plain forward(),
 *                  which the reference calls without position_ids, modeling_visualcla.py:321-328).  Not with IMAGE_AT_HEAD.
 *   logits_all_dev f32 (B,S,V) or NULL   -- VisualCLAModel.forward(...).logits (modeling_visualcla.py:321-328)
 *   last_logits_dev f32 (B,V) or NULL    -- logits of the last position (what generate() samples from)
 *   next_tok_dev   int32 (B) or NULL     -- argmax of last_logits (greedy)
 * Fills the KV cache; sequence length becomes S.  Replaces the splice (modeling_visualcla.py:290-312 /
 * :356-377) + LlamaForCausalLM prefill (HF:models/llama/modeling_llama.py:375-501). */
int vcla_prefill(vcla_ctx* ctx, const int64_t* ids_dev, int B, int T, int image_mode, const int32_t* img_row_dev,
                 const int32_t* left_pad_dev, int pos_from_mask, float* logits_all_dev, float* last_logits_dev,
                 int32_t* next_tok_dev, vcla_stream stream);

/* One greedy decode step for the B resident sequences: consumes tok_in_dev (int32 (B)), appends its K/V,
 * writes logits (f32 (B,V), optional) and the argmax (int32 (B)).  Captured into a CUDA graph on first use
 * (per distinct argument tuple) when use_graph != 0.  Replaces one iteration of
 * GenerationMixin._sample (HF:generation/utils.py:2743-2810) incl. DynamicCache.update (HF:cache_utils.py:119-120). */
int vcla_decode_step(vcla_ctx* ctx, const int32_t* tok_in_dev, int B, float* logits_dev, int32_t* tok_out_dev, int use_graph,
                     vcla_stream stream);

/* n_steps (1..64) greedy decode steps replayed as ONE CUDA graph; tok_inout_dev (int32 (B)) is consumed and rewritten in place by
 * every step and every chosen token is appended to the history (vcla_read_history).  Same arithmetic as n_steps calls of
 * vcla_decode_step; amortises the launch gap between steps. */
int vcla_decode_multi(vcla_ctx* ctx, int32_t* tok_inout_dev, int B, int n_steps, vcla_stream stream);

/* Tokens chosen so far: row 0 = the prefill's argmax, row s = decode step s.  Copies [n_steps, B] int32 to a DEVICE buffer
 * (async on `stream`): lets a greedy loop run as pure graph replays with no per-step host or torch work. */
int vcla_read_history(vcla_ctx* ctx, int32_t* dst_dev, int B, int n_steps, vcla_stream stream);

/* ---- device-side sampling ---------------------------------------------------------------------
 * chat()'s real default is sampling (models/visualcla/modeling_utils.py:36-47: temperature 0.5, top_k 40, top_p 0.9,
 * repetition_penalty 1.1, no_repeat_ngram_size 15).  With a sampler set, vcla_prefill / vcla_decode_step / vcla_decode_multi
 * replace the argmax by ONE fused kernel per step -- HF's processor chain in HF's order (RepetitionPenalty, NoRepeatNGram,
 * min_new_tokens EOS mask, Temperature, TopK, TopP; HF:generation/logits_process.py) + a Philox-keyed multinomial draw -- on the
 * device token history (with inputs_embeds HF's processors only see the new tokens), inside the captured CUDA graph: no logits
 * leave the device, no host work per token.  do_sample = 0 takes the argmax of the processed scores (greedy + penalties).
 * A sequence that emitted an EOS id keeps producing pad_token_id (sticky per-sequence flag, vcla_read_finished).
 * The parameters live in device memory: changing them does not re-capture graphs. */
typedef struct {
  int do_sample;
  float repetition_penalty;     /* 1 = off */
  int no_repeat_ngram_size;     /* 0 = off */
  float temperature;            /* 1 = off */
  int top_k;                    /* 1..1024, required when do_sample */
  float top_p;                  /* 1 = off */
  int min_new_tokens;
  int n_eos;                    /* <= 4 */
  int eos_token_id[4];
  int pad_token_id;
  uint64_t seed;                /* draw = Philox4x32-10(key = seed, counter = (step, sequence)) */
} vcla_sampler;
int vcla_sampler_supported(const vcla_ctx* ctx);   /* 1 when the vocabulary row fits one CTA's shared memory */
int vcla_set_sampler(vcla_ctx* ctx, const vcla_sampler* sampler_or_null, vcla_stream stream);   /* NULL: back to greedy argmax */
int vcla_read_finished(vcla_ctx* ctx, int32_t* dst_dev, int B, vcla_stream stream);
/* Operator-level entry (parity tests): the same kernel on caller logits (B,V) f32 and a token history [L][B] int32; writes the
 * chosen tokens (B) and, if not NULL, the processed scores (B,V) (-inf = filtered) exactly as HF's chain would return them.
 * Synchronises. */
int vcla_op_sample(const float* logits_dev, int B, int V, const int32_t* history_dev, int L, const vcla_sampler* sampler,
                   int32_t* tok_dev, float* scores_out_dev, vcla_stream stream);

/* ---- data parallel over the GPUs of one box (the reference has no DP of its own) ----------------
 * Requests are independent through the whole path, so each rank (one process + one context per GPU) runs a contiguous slice of
 * the batch and the ONLY exchange is one NCCL all-gather of the chosen token ids per decode step.  After vcla_nccl_init the
 * exchange is part of the path itself: vcla_prefill and every decode step (also inside the CUDA graphs of vcla_decode_multi)
 * all-gather `width` int32 slots per rank on a forked stream branch -- off the step's critical path, joined before the send
 * buffer is rewritten -- and append them to a device-side global history.
 *   vcla_nccl_unique_id   rank 0 creates the 128-byte ncclUniqueId and ships it to the other ranks by any means
 *   vcla_nccl_init        width = slots per rank (the largest shard's batch, <= 64); collective over all ranks
 *   vcla_allgather_tokens plain all-gather of n int32 per rank on `stream` (the non-graph building block)
 *   vcla_dp_set_active    the exchange is part of vcla_prefill / vcla_decode_* only while active (off after init): every rank must
 *                         then make the same sequence of calls; a rank-local generation runs with it off
 *   vcla_dp_exchange      one step's exchange without compute (a rank that holds no requests: global batch < world size)
 *   vcla_read_history_dp  [n_steps][world * width] int32 -> device buffer: row 0 = prefill argmax of every rank, row s = step s
 * NCCL is bound with dlopen("libnccl.so.2") at the first of these calls; single-GPU use never loads it. */
int vcla_nccl_unique_id(uint8_t* out128);
int vcla_nccl_init(vcla_ctx* ctx, const uint8_t* id128, int rank, int world, int width);
int vcla_allgather_tokens(vcla_ctx* ctx, const int32_t* local_dev, int n, int32_t* all_dev, vcla_stream stream);
int vcla_dp_set_active(vcla_ctx* ctx, int on);
int vcla_dp_exchange(vcla_ctx* ctx, vcla_stream stream);
int vcla_read_history_dp(vcla_ctx* ctx, int32_t* dst_dev, int n_steps, vcla_stream stream);

/* number of this library's kernels launched by the context since the last call with reset != 0 */
int64_t vcla_kernel_launches(vcla_ctx* ctx, int reset);

/* ---- introspection for parity tests ---------------------------------------------------------- */
/* Copy an internal fp32 activation to the host (synchronises): "vit_out" (B,tokens,v_hidden), "post_ln",
 * "resampler_out" (B,nq,r_hidden), "projector_out" (B,nq,t_hidden), "inputs_embeds" n/a after prefill. */
int vcla_read_stage(vcla_ctx* ctx, const char* stage, int B, float* dst_host, vcla_stream stream);

/* ---- operator-level entry points (kernel parity tests, micro-benchmarks) ---------------------- */
/* D = A[M,K] * W[N,K]^T on the tcgen05 path.  mode: 0 store bf16 (act: 0 none, 1 quick_gelu, 2 gelu),
 * 1 fp32 (accumulate flag), 2 SwiGLU (W rows interleaved [32 gate|32 up]), 3 swap-AB split-K partials
 * (out f32 [splits][M_b][N] with A = weights).  use_reference != 0 runs the naive CUDA-core kernel instead. */
int vcla_op_gemm(const void* A_dev_bf16, const void* W_dev_bf16, int M, int N, int K, int mode, int act, int accumulate,
                 const float* bias_dev, void* out_dev, int ldo, int splits, int tile_n, int use_reference, vcla_stream stream);
/* Decode GEMM with the split-K reduction inside a thread-block cluster (csrc/gemm_decode.cu): out[b, n] = sum_k W[n, k] X[b, k], W (M, K)
 * bf16 streamed once, X (B <= 32, K) bf16, `splits` CTAs per cluster (1..8).  mode 0: out_or_resid = out f32 (B, M) = rstd[b] * acc;
 * mode 1: out_or_resid = resid f32 (B, M) += acc, xw_or_h = bf16 (B, M) = resid * norm_w, ssq_out f32 (B, ceil(M / 128)) = per-tile sums of
 * squares; mode 2: W rows interleaved [32 gate | 32 up], xw_or_h = h bf16 (B, M / 2) = silu(rstd * g) * (rstd * u).
 * rstd[b] = rsqrt(sum_slots ssq_in[b][slot] * inv_dim + eps), 1 when ssq_in is NULL.  vcla_op_gemm_csk_clusters: co-resident clusters. */
int vcla_op_gemm_csk(const void* W_dev_bf16, const void* X_dev_bf16, int M, int B, int K, int splits, int mode, float* out_or_resid,
                     const float* norm_w, void* xw_or_h, float* ssq_out, const float* ssq_in, int ssq_slots, float inv_dim, float eps,
                     vcla_stream stream);
int vcla_op_gemm_csk_clusters(int B, int splits);
/* tuning hooks: read / override the CTAs-per-cluster of the five decode GEMM shapes {qkv, o, gate_up, down, lm_head} at batch B */
int vcla_debug_set_csk_splits(vcla_ctx* ctx, int B, int qkv, int o, int gate_up, int down, int lm_head);
int vcla_debug_get_csk_splits(vcla_ctx* ctx, int B, int* out5);
/* CTA-pair (tcgen05 cta_group::2, 256 x 256) tiles for the 256-wide prefill GEMMs: on by default; 0 selects the single-CTA 128 x 256 tile */
void vcla_set_gemm_two_cta(int on);
/* prefill attention kernel: 0 = the mma.sync kernel everywhere, 1 (default) = tcgen05 flash attention (QK^T / PV as UMMA, S and O in TMEM,
 * TMA operands) at head dim 128 (LLaMA prefill) and mma.sync at head dim 64 (ViT / Resampler), 2 = tcgen05 everywhere */
void vcla_set_attention_tc(int mode);
int vcla_op_attention(const void* q, int q_stride, const void* k0, const void* v0, int kv0_stride, int n0, const void* k1,
                      const void* v1, int kv1_stride, int n1, void* out, int o_stride, int B, int H, int Sq, int HD, float scale,
                      int causal, vcla_stream stream);
int vcla_op_layernorm(const float* x, int rows, int D, const float* w, const float* b, float eps, void* y_bf16, float* y_f32,
                      vcla_stream stream);
int vcla_op_rmsnorm(const float* x, int rows, int D, const float* w, float eps, void* y_bf16, vcla_stream stream);
/* Micro-benchmark of ONE decode weight-streaming GEMM shape (which: 0 fused QKV, 1 o_proj, 2 fused gate/up, 3 down_proj,
 * 4 lm_head) over every layer's distinct weights with batch B, timed with CUDA events on `stream`; returns the mean
 * microseconds per kernel launch and the algorithmic weight bytes one launch streams.  Synchronises. */
int vcla_bench_decode_gemm(vcla_ctx* ctx, int which, int B, int reps, float* avg_us, int64_t* weight_bytes, vcla_stream stream);
/* Timeline trace for profiles/: when enabled, CTA (0,0,0) of every kernel appends {tag, t_entry, t_dependency_resolved, t_exit}
 * (%globaltimer, ns).  Tags: 1 swap-AB GEMM, 2 GEMM, 3 prefill attention, 4 decode attention, 5 layernorm, 6 rmsnorm, 7 rope+cache,
 * 8 resid+rmsnorm, 9 silu*mul, 10/11 logits+argmax, 12 advance, 13 embed, 14 sampler.  vcla_trace_read synchronises and clears. */
int vcla_trace_enable(vcla_ctx* ctx, int max_events);
int vcla_trace_read(vcla_ctx* ctx, uint64_t* dst_host, int max_events, int* n_events);
/* enable/disable programmatic dependent launch for subsequently enqueued kernels (process-wide) */
void vcla_set_pdl(int on);

/* ---- image pre-processing (the step before the path) ------------------------ */
/* Replaces HF CLIPImageProcessor's PIL pipeline as the reference calls it for every request
 * (models/visualcla/modeling_utils.py:130 builds it, :150-152 / :187-189 call it):
 *   resize(shortest_edge = out_size, BICUBIC) -> center_crop(out_size) -> x * 1/255 -> (x - mean) / std
 * The resize is Pillow's 8-bit ImagingResample (antialiased separable bicubic, 22-bit taps, clip after each pass,
 * horizontal first) and the result is bit-identical to it.  No context needed: the caller owns every buffer.
 *
 * vcla_preprocess_workspace_bytes: device scratch one call needs for a (height, width) picture; -1 if unsupported
 *   (sides 1..32768, out_size 1..4096, resized long side <= 65536).  Host-only, no GPU needed.
 * vcla_resample_taps: Pillow's tap table of one axis (host-only): first[out], count[out], taps[out][ksize] with 22
 *   fractional bits.  Returns ksize; with all three pointers NULL only returns ksize.  -1 on error.
 * vcla_preprocess_image: rgb_dev = (height, width, 3) uint8 RGB in device memory; pixel_values_dev = (3, out_size,
 *   out_size) in `dtype` (VCLA_F32 / F16 / BF16, round-to-nearest from the float32 result); mean3/std3 = host floats.
 *   Builds the tap tables on the host, uploads them into the workspace and enqueues two kernels on `stream`.  The
 *   workspace must be 16-byte aligned and stay untouched until the stream has run them. */
int64_t vcla_preprocess_workspace_bytes(int height, int width, int out_size);
int vcla_resample_taps(int in_size, int out_size, int32_t* first, int32_t* count, int32_t* taps, int ksize_capacity);
int vcla_preprocess_image(const uint8_t* rgb_dev, int height, int width, int out_size, const float* mean3,
                          const float* std3, void* workspace_dev, int64_t workspace_bytes, void* pixel_values_dev,
                          int dtype, vcla_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* VCLA_H_ */
