// Host-side launch API of every CUDA kernel on the VisualCLA path (internal to libvcla.so).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace vcla {

typedef __nv_bfloat16 bf16;

void set_error(const char* fmt, ...);
const char* get_error();
int num_sms();
bool pdl_enabled();
void set_pdl(bool on);

#define VCLA_CUDA_OK(expr)                                                                   \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      vcla::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      (void)cudaGetLastError();                                                              \
      return -1;                                                                             \
    }                                                                                        \
  } while (0)

// ------------------------------------------------------------------------------------------
// tcgen05 GEMM:  D[M,N] = A[M,K] * B[N,K]^T   (both operands K-major bf16, fp32 accumulate in TMEM)
// ------------------------------------------------------------------------------------------
enum GemmMode {
  GEMM_STORE_BF16 = 0,   // out_bf16[orow, col] = act(acc + bias[col])
  GEMM_ADD_F32 = 1,      // out_f32[orow, col]  = (accumulate ? old : 0) + acc + bias[col] + rowtab[(row % period), col]
  GEMM_SWIGLU_BF16 = 2,  // B rows interleaved [32 gate | 32 up]: out_bf16[orow, col/2] = silu(g) * u
  GEMM_PARTIAL_F32 = 3,  // swap-AB split-K partials: ws[(split*ws_rows + col) * ldo + row] = acc   (row = A row)
};
enum Act { ACT_NONE = 0, ACT_QUICK_GELU = 1, ACT_GELU_ERF = 2 };
// Fused consumer of the split-K partials, run by the CTA that completes a tile ("last arriver", GEMM_PARTIAL_F32 only):
//   FIX_RESID : r = resid[b,n] + sum_s partial ; resid = r ; xw[b,n] = bf16(r * norm_w[n]) ; ssq[b, tile] = sum_n r^2 ;
//               the CTA that completes the LAST tile also writes rstd[b] = rsqrt(sum_tiles ssq / N + eps).
//               (RMSNorm with the per-row scale deferred to the next consumer: a row scalar commutes with the next GEMM.)
//   FIX_SWIGLU: g,u = rstd[b] * sum_s partial (rows interleaved [32 gate | 32 up]) ; h[b,j] = bf16(silu(g) * u)
enum FixMode { FIX_NONE = 0, FIX_RESID = 1, FIX_SWIGLU = 2 };
struct GemmFix {
  int mode = FIX_NONE;
  int32_t* tile_counters = nullptr;   // [m_tiles + 1] zero between launches; the last entry counts finished tiles
  float* resid = nullptr;             // [B, N_out] fp32 (FIX_RESID)
  const float* norm_w = nullptr;      // [N_out]
  bf16* xw_out = nullptr;             // [B, N_out]
  float* ssq = nullptr;               // [B, m_tiles] scratch
  float* rstd_out = nullptr;          // [B]
  float inv_dim = 0.f, eps = 0.f;
  const float* rstd_in = nullptr;     // [B] (FIX_SWIGLU)
  bf16* h_out = nullptr;              // [B, N_out/2]
};

// ---- prefill epilogue fusions (non-swap GEMMs) --------------------------------------------------------------------------
// Deferred RMSNorm: the A operand holds xw = bf16(resid * norm_w) (NOT normalised); the row scale rstd[row] =
// rsqrt(sum_slots ssq[row][slot] / dim + eps) commutes with the GEMM and is applied to the accumulator in the epilogue.
struct GemmRowScale {
  const float* ssq = nullptr;   // [M][slots] partial sums of squares of the fp32 residual rows (written by the producing GEMM)
  int slots = 0;
  float inv_dim = 0.f, eps = 0.f;
};
// GEMM_ADD_F32 + accumulate: besides resid += acc, emit the NEXT GEMM's operand and the row statistics:
//   xw[orow, col] = bf16(resid_new * norm_w[col]) ; ssq_out[orow][n_blk] = sum over this tile's columns of resid_new^2
struct GemmEmitNorm {
  const float* norm_w = nullptr;   // [N]
  bf16* xw = nullptr; int ldxw = 0;
  float* ssq_out = nullptr;        // [M][n_tiles]
};
// GEMM_STORE_BF16 on the fused QKV projection [M, 3T]: rotate q and k heads (HF rotate_half pairs d, d+64; fp32 tables
// [pos][64]) on the fp32 accumulator, store q|k|v rows (the prefill attention reads them) AND append k, v to the paged KV cache.
struct GemmRope {
  const float* cos = nullptr; const float* sin = nullptr;
  bf16* kv_pages = nullptr; const int32_t* page_table = nullptr; int pages_per_seq = 0, page_tokens = 0;
  int S = 0, T = 0, H = 0;                     // rows per sequence, hidden size (= H * 128), heads
  const int32_t* left_pad = nullptr; int pos_from_mask = 0;
};

struct GemmCall {
  const bf16* A = nullptr;   // [M, K], row pitch lda elements
  const bf16* B = nullptr;   // [N, K], row pitch ldb elements
  int M = 0, N = 0, K = 0;
  int lda = 0, ldb = 0;
  int mode = GEMM_STORE_BF16;
  void* out = nullptr;
  int ldo = 0;
  const float* bias = nullptr;
  int act = ACT_NONE;
  int accumulate = 0;
  const float* rowtab = nullptr;
  int rowtab_period = 1;
  // output row remap: orow = (row / rows_per_group) * group_stride + (row % rows_per_group) + row_offset
  int rows_per_group = 0;    // 0 = identity
  int group_stride = 0;
  int row_offset = 0;
  // split-K (GEMM_PARTIAL_F32 only)
  int splits = 1;
  int ws_rows = 0;           // padded batch rows in the partial workspace
  int weights_are_A = 0;     // cache-policy hint: A is the streamed-once operand (decode)
  int bn = 0;                // tile N override (0 = auto)
  int l2_prefetch_kb = 0;    // k-blocks of the weight operand each CTA prefetches into L2 while it waits for its dependency
  GemmFix fix;               // fused consumer of the split-K partials (decode)
  GemmRowScale rowscale;     // deferred RMSNorm scale of the A rows (STORE_BF16 / SWIGLU_BF16)
  GemmEmitNorm emit;         // ADD_F32 + accumulate: also write the next operand + row statistics
  GemmRope rope;             // STORE_BF16: RoPE + KV-cache append (needs BN = 256, N = 3T)
};
int gemm_tc(const GemmCall& c, cudaStream_t st);
void gemm_set_two_cta(int on);     // CTA-pair (cta_group::2) 256 x 256 tiles for the 256-wide prefill GEMMs (default on)
int gemm_pick_bn(int M, int N);   // tile width gemm_tc picks for a non-swap GEMM (= the number of ssq slots per row it emits: ceil(N / bn))
// correctness reference for the tests only (CUDA-core, one thread per output)
int gemm_naive(const GemmCall& c, cudaStream_t st);
int gemm_init();   // resolves cuTensorMapEncodeTiled, sets smem attributes

// ------------------------------------------------------------------------------------------
// decode GEMM with the split-K reduction inside a thread-block cluster (gemm_decode.cu)
//   out[b, n] = sum_k W[n, k] * X[b, k]   W [M = N_out, K] bf16 (streamed), X [B <= 32, K] bf16
// ------------------------------------------------------------------------------------------
enum CskMode {
  CSK_OUT_F32 = 0,   // out[b * ldo + n] = rstd[b] * acc
  CSK_RESID = 1,     // resid[b, n] += acc ; xw[b, n] = bf16(resid * norm_w[n]) ; ssq_out[b, n / 128] = sum over the tile of resid^2
  CSK_SWIGLU = 2,    // W rows interleaved [32 gate | 32 up]: h[b, j] = bf16(silu(rstd[b] * g) * (rstd[b] * u))
};
struct CskCall {
  const bf16* W = nullptr; const bf16* X = nullptr;
  int M = 0, B = 0, K = 0;
  int splits = 1;                       // CTAs per cluster = K slices (1..8, every slice non-empty)
  int mode = CSK_OUT_F32;
  float* out = nullptr; int ldo = 0;
  float* resid = nullptr; const float* norm_w = nullptr; bf16* xw = nullptr; float* ssq_out = nullptr;
  bf16* h = nullptr;
  const float* ssq_in = nullptr; int ssq_slots = 0; float inv_dim = 0.f, eps = 0.f;   // rstd[b] = rsqrt(sum_slots ssq_in[b][slot] * inv_dim + eps); null: 1
};
int gemm_csk(const CskCall& c, cudaStream_t st);
int gemm_csk_clusters(int B, int splits);   // clusters of `splits` CTAs that can be co-resident (occupancy query, cached)
int trace_set_gemm(void* buf, unsigned long long cap);
int trace_set_attention(void* buf, unsigned long long cap);
int trace_set_gemm_decode(void* buf, unsigned long long cap);
int trace_set_sampler(void* buf, unsigned long long cap);
int trace_set_elementwise(void* buf, unsigned long long cap);

// ------------------------------------------------------------------------------------------
// attention
// ------------------------------------------------------------------------------------------
struct AttnCall {
  const bf16* q = nullptr; int q_stride = 0;            // q[(b*Sq + i)*q_stride + h*HD + d]
  const bf16* k0 = nullptr; const bf16* v0 = nullptr; int kv0_stride = 0; int n0 = 0;   // segment 0: n0 rows / batch
  const bf16* k1 = nullptr; const bf16* v1 = nullptr; int kv1_stride = 0; int n1 = 0;   // segment 1 (optional)
  bf16* out = nullptr; int o_stride = 0;
  int B = 0, H = 0, Sq = 0, HD = 0;
  float scale = 1.f;
  int causal = 0;
  const int32_t* kv_start = nullptr;   // [B] first visible kv index per sequence (left padding); null: 0
};
int attention_prefill(const AttnCall& c, cudaStream_t st);      // dispatches to the tcgen05 kernel (attention_tc.cu) unless switched off
int attention_prefill_tc(const AttnCall& c, cudaStream_t st);   // tcgen05: QK^T and PV as UMMA, S / O in TMEM, Q / K / V by TMA
int attention_prefill_mma(const AttnCall& c, cudaStream_t st);  // mma.sync m16n8k16 fallback (attention.cu)
void attention_set_tc(int mode);                                 // 0: mma.sync everywhere, 1 (default): tcgen05 at head dim 128, 2: tcgen05 everywhere (VCLA_ATTN_TC)
int trace_set_attention_tc(void* buf, unsigned long long cap);

struct DecodeAttnCall {
  const float* qkv_partial = nullptr;  // [splits][ws_rows][3*T] fp32 split-K partials of the fused QKV projection
  int splits = 1, ws_rows = 0;
  bf16* kv_pages = nullptr;            // this layer: [pages][2][H][page_tokens][HD]
  const int32_t* page_table = nullptr; // [max_batch][pages_per_seq]
  int pages_per_seq = 0, page_tokens = 0;
  const int32_t* seq_len = nullptr;    // [B] tokens already in the cache (the new token is appended at this index)
  bf16* out = nullptr;                 // [ws_rows][T] attention output (bf16, GEMM operand of o_proj)
  float* scratch = nullptr;            // [B][H][kv_splits][HD+2]
  int32_t* counters = nullptr;         // [B][H]
  int B = 0, H = 0, HD = 0, kv_splits = 1;
  float scale = 1.f, rope_theta = 10000.f;
  const float* rstd = nullptr;         // [B] deferred RMSNorm scale of the QKV projection's input (null: 1)
  const float* rope_cos = nullptr;     // [max_pos][HD/2] fp32 tables owned by the context
  const float* rope_sin = nullptr;
  int persistent_mode = 1;             // VCLA_ATTN_PERSISTENT (read once per context)
  int persistent_grid = 0;             // VCLA_ATTN_PERSISTENT_GRID
};
int attention_decode(const DecodeAttnCall& c, cudaStream_t st);
int attention_init();          // sets the dynamic-smem attributes and reads the VCLA_ATTN_* switches once (call outside graph capture)

// ------------------------------------------------------------------------------------------
// normalisation / elementwise / data movement
// ------------------------------------------------------------------------------------------
// y_bf16 = LN(x) * w + b (fp32 statistics); optionally also writes the fp32 normalised row back (in place ok)
int layernorm(const float* x, int rows, int D, const float* w, const float* b, float eps, bf16* y_bf16, float* y_f32,
              cudaStream_t st);
int rmsnorm(const float* x, int rows, int D, const float* w, float eps, bf16* y_bf16, cudaStream_t st);
// head of the deferred-norm chain: xw = bf16(x * w) (not normalised), ssq[row][0] = sum x^2, ssq[row][1..slots) = 0
int prenorm_rows(const float* x, int rows, int D, const float* w, bf16* xw, float* ssq, int slots, cudaStream_t st);
// pixels (B,3,I,I) in f32/f16/bf16 -> im2col rows [B*g*g, Kpad] bf16 (k = c*P*P + ky*P + kx), zero padded
int im2col(const void* pixels, int dtype, int B, int image, int patch, int kpad, bf16* out, cudaStream_t st);
// hidden[b, 0, :] = cls + pos[0]
int vit_cls_rows(float* hidden, int B, int tokens, int D, const float* cls, const float* pos, cudaStream_t st);
int broadcast_rows(const float* src, int rows, int D, int B, float* dst_f32, bf16* dst_bf16, cudaStream_t st);
// text embedding gather into the fp32 residual stream: dst[b, dst_pos(t), :] = table[ids[b,t], :]
//   mode 0: dst_pos = t (text only / placeholder layout: image rows are overwritten afterwards by the projector GEMM)
//   mode 1: image at head: t<2 -> t ; t>=2 -> t + nq
int embed_tokens(const int64_t* ids, int B, int T, int S, int D, const bf16* table, int vocab, int mode, int nq,
                 float* dst, cudaStream_t st);
int embed_tokens_i32(const int32_t* ids, int B, int D, const bf16* table, int vocab, float* dst, cudaStream_t st);
// copy the projected image rows (B, nq, D) fp32 into the residual stream at per-sample row offsets
int scatter_image_rows(const float* img, int B, int nq, int D, const int32_t* row_start, int S, float* dst, cudaStream_t st);
// prefill: RoPE q,k in place in the fused qkv buffer [B*S, 3T] and append k,v to the paged cache
// left_pad[b] (nullable): rows s < left_pad[b] are padding (skipped); the cache index of row s is s - left_pad[b]; the RoPE
// position is s - left_pad[b] when pos_from_mask (HF generate) else s (plain forward without position_ids)
int rope_and_cache(bf16* qkv, int B, int S, int H, int HD, const float* rope_cos, const float* rope_sin, bf16* kv_pages,
                   const int32_t* page_table, int pages_per_seq, int page_tokens, const int32_t* left_pad, int pos_from_mask,
                   cudaStream_t st);
int gather_last_rows(const float* hidden, int B, int S, int D, float* dst, cudaStream_t st);

// decode consumers of split-K partials
// resid[b,:] += sum_s partial[s][b][:]  (partial may be null) ; xn = rmsnorm(resid) -> bf16
int dec_resid_norm(const float* partial, int splits, int ws_rows, float* resid, int B, int D, const float* w, float eps,
                   bf16* xn, cudaStream_t st);
// h[b, j] = silu(sum_s p[s][b][g(j)]) * (sum_s p[s][b][u(j)])  with the [32 gate | 32 up] interleave
int dec_silu_mul(const float* partial, int splits, int ws_rows, int B, int F, bf16* h, cudaStream_t st);
// logits[b, :] = sum_s partial[s][b][:V] ; tok[b] = argmax (first max wins, like torch.argmax)
// cand_val / cand_idx: [B][kArgmaxChunks] scratch owned by the context
constexpr int kArgmaxChunks = 32;
int dec_logits_argmax(const float* partial, int splits, int ws_rows, int ldp, int B, int V, float* logits, int ld_logits,
                      int32_t* tok, int32_t* history, const int32_t* step_idx, const float* rstd, float* cand_val,
                      int32_t* cand_idx, int32_t* dp_send, cudaStream_t st);
// stage 1 only: logits[b, :] = rstd[b] * sum_s partial (the sampler consumes them)
int dec_logits_reduce(const float* partial, int splits, int ws_rows, int ldp, int B, int V, float* logits, int ld_logits, const float* rstd,
                      float* cand_val, int32_t* cand_idx, cudaStream_t st);
// ---- device-side sampling (sampler.cu) ----------------------------------------------------------------------------------
struct SamplerParams {     // lives in device memory: graphs captured once serve every parameter set
  int do_sample;           // 0: argmax of the processed scores
  float rep_penalty;       // 1 = off
  int no_repeat_ngram;     // 0 = off
  float temperature;       // 1 = off
  int top_k;               // 1..1024 (required when do_sample)
  float top_p;             // 1 = off
  float one_minus_top_p;   // (float)(1.0 - (double)top_p): the constant HF compares the cumulative probabilities with
  int min_new_tokens, n_eos, pad_id;
  int eos[4];
  unsigned long long seed;
};
int sampler_supported(int V);
int sampler_init();
// history: [L][B] int32 with L = *step_idx; writes tok[b], history_out[L][b], dp_send[b]; finished[b] (nullable): sticky EOS flag
int dec_sample(const float* logits, int ld, int V, int B, const int32_t* history, const int32_t* step_idx, const SamplerParams* params_dev,
               int32_t* tok, int32_t* history_out, int32_t* dp_send, int32_t* finished, float* scores_out, cudaStream_t st);
int dp_unpack(const int32_t* recv, int n, int32_t* hist, int32_t* dp_step, cudaStream_t st);
// decode step entry: resid[b,:] = table[ids[b]] ; xw = bf16(resid * norm_w) ; rstd[b] = rsqrt(mean(resid^2) + eps) (nullable) ;
// ssq[b][0] = sum resid^2, ssq[b][1..slots) = 0 (nullable: head of the deferred-norm chain of the cluster split-K schedule)
int dec_embed(const int32_t* ids, int B, int D, const bf16* table, int vocab, float* resid, const float* norm_w, float eps,
              bf16* xw, float* rstd, float* ssq, int slots, cudaStream_t st);
// ---- device-side KV page allocator (stream-ordered, graph-capturable; one thread walks the <= 64 sequences, so the
//      assignment is deterministic) ------------------------------------------------------------------------------
// kv_state[0] = free pages, kv_state[1] = error flag (pool exhausted); kv_free = stack of free physical pages;
// kv_npages[b] = pages owned by sequence b; page_table[b][i] = i-th page of sequence b.
int kv_reset(int32_t* kv_free, const int32_t* kv_order, int32_t* kv_state, int32_t* kv_npages, int total_pages, int max_batch,
             cudaStream_t st);
// make sure sequence b owns pages for (S - left_pad[b]) tokens, b < B; pages are handed out round-robin over the sequences
int kv_reserve(int32_t* kv_free, int32_t* kv_state, int32_t* kv_npages, int32_t* page_table, int pages_per_seq, int page_tokens,
               int B, int S, const int32_t* left_pad, cudaStream_t st);
// seq_len[b] += by - left_pad[b] ; *step_idx += 1 ; then reserve the page the NEXT token of every sequence will be appended to
int advance_seq(int32_t* seq_len, int B, int by, const int32_t* left_pad, int32_t* step_idx, int32_t* kv_free, int32_t* kv_state,
                int32_t* kv_npages, int32_t* page_table, int pages_per_seq, int page_tokens, cudaStream_t st);
// fp32 RoPE tables [max_pos][head_dim/2], computed on the host the way HF does and uploaded to cos_dev / sin_dev
int rope_fill_tables(int max_pos, int head_dim, float theta, float* cos_dev, float* sin_dev);

// weights
int fill_hash_normal(bf16* dst_bf16, float* dst_f32, int64_t n, uint32_t seed, float mul, float offset, cudaStream_t st);
int convert_to_bf16(const void* src, int dtype, int64_t n, bf16* dst, cudaStream_t st);
int convert_to_f32(const void* src, int dtype, int64_t n, float* dst, cudaStream_t st);
// dst rows [r0, r0+rows) of a [*, ld] bf16 matrix <- src [rows, cols] (zero pad cols..ld)
int copy_rows_bf16(const bf16* src, int rows, int cols, bf16* dst, int ld, cudaStream_t st);
// interleave gate/up rows in blocks of 32: dst[(j/32)*64 + which*32 + j%32, :] = src[j, :]
int interleave_rows32(const bf16* src, int rows, int cols, int which, bf16* dst, cudaStream_t st);

}  // namespace vcla
