// Engine behind the C ABI (include/vcla.h): device arenas, weight registry / repacking, and the orchestration of
// the three phases of the path -- vision encode, prefill, decode step (CUDA-graph captured).
#include "../../include/vcla.h"
#include "kernels.h"

#include <dlfcn.h>
#include <math.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <list>
#include <map>
#include <string>
#include <vector>

using namespace vcla;

namespace {

enum SlotKind { SLOT_MAT = 0, SLOT_VEC = 1 };
enum SlotLayout { LAY_PLAIN = 0, LAY_INTERLEAVE32 = 1 };

struct Slot {
  std::string name;
  int64_t shape[4] = {0, 0, 0, 0};
  int ndim = 0;
  int kind = SLOT_MAT;
  int layout = LAY_PLAIN;
  int which = 0;           // interleave: 0 gate, 1 up
  int64_t rows = 0, cols = 0;  // 2D view of the logical tensor
  void* dst = nullptr;     // storage (bf16 for MAT, f32 for VEC) at the slot's first row
  int ld = 0;              // storage row pitch (elements) for MAT
  void* dst2 = nullptr;    // optional second copy (resampler k/v also live in the all-layer KV weight)
  double std = 0.0; float mean = 0.f;
};

struct VisionLayer {
  float *ln1_w, *ln1_b, *ln2_w, *ln2_b, *bqkv, *bo, *b1, *b2;
  bf16 *wqkv, *wo, *w1, *w2;
};
struct ResamplerLayer {
  bf16 *wqkv, *wo, *wi, *wo2;
  float *bqkv, *bo, *ln1_w, *ln1_b, *bi, *bo2, *ln2_w, *ln2_b;
};
struct TextLayer {
  float *ln1, *ln2;
  bf16 *wqkv, *wo, *wgu, *wd;
  bf16* kv;  // this layer's pages
};

struct GraphKey {
  int B; const void* tok_in; const void* logits; const void* tok_out; int n_steps = 1; int dp = 0; int samp = 0;
  bool operator<(const GraphKey& o) const {
    if (B != o.B) return B < o.B;
    if (tok_in != o.tok_in) return tok_in < o.tok_in;
    if (logits != o.logits) return logits < o.logits;
    if (n_steps != o.n_steps) return n_steps < o.n_steps;
    if (dp != o.dp) return dp < o.dp;
    if (samp != o.samp) return samp < o.samp;
    return tok_out < o.tok_out;
  }
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

uint32_t fnv1a32(const char* s) {
  uint32_t h = 0x811C9DC5u;
  for (; *s; ++s) { h ^= (uint8_t)*s; h *= 0x01000193u; }
  return h;
}

}  // namespace

struct vcla_ctx {
  vcla_config cfg;
  int v_tokens = 0, kpatch = 0, kpad = 0, hd_t = 0;
  // arenas
  uint8_t* w_arena = nullptr; size_t w_bytes = 0, w_off = 0;
  uint8_t* a_arena = nullptr; size_t a_bytes = 0, a_off = 0;
  bf16* kv_arena = nullptr; size_t kv_bytes = 0;
  std::vector<Slot> slots;
  std::map<std::string, int> slot_index;
  // vision weights
  bf16* patch_w = nullptr; float *cls = nullptr, *pos = nullptr, *pre_w = nullptr, *pre_b = nullptr, *post_w = nullptr, *post_b = nullptr;
  std::vector<VisionLayer> vl;
  // resampler
  float* rq = nullptr; bf16* r_wkv_all = nullptr; float* r_bkv_all = nullptr;
  std::vector<ResamplerLayer> rl;
  bf16* proj_w = nullptr; float* proj_b = nullptr;
  // text
  bf16 *embed = nullptr, *lm_head = nullptr; float* final_norm = nullptr;
  std::vector<TextLayer> tl;
  // kv cache
  int pages_per_seq = 0, page_tokens = 64, total_pages = 0;
  size_t kv_layer_elems = 0;
  int32_t *page_table = nullptr, *seq_len = nullptr, *img_row_default = nullptr;
  // device-side page allocator (elementwise.cu: kv_reset / kv_reserve / advance_seq): a stack of free physical pages, pages handed
  // to sequences round-robin as they grow (so a sequence's pages are NOT contiguous), all stream-ordered and graph-capturable
  int32_t *kv_free = nullptr, *kv_order = nullptr, *kv_state = nullptr, *kv_npages = nullptr;   // kv_state: [0] free count, [1] error flag
  // RoPE tables and argmax scratch are per context (another context may use another theta / device / stream)
  float *rope_cos = nullptr, *rope_sin = nullptr, *cand_val = nullptr; int32_t* cand_idx = nullptr;
  int attn_persistent_mode = 1, attn_persistent_grid = 0;
  // data parallel (vcla_nccl_init): per-step all-gather of the chosen tokens, captured inside the decode graph on a forked branch
  ncclComm_t comm = nullptr; int dp_rank = 0, dp_world = 1, dp_width = 0; bool dp_active = false;
  bool dp_on() const { return comm != nullptr && dp_active; }
  int32_t *dp_send = nullptr, *dp_recv = nullptr, *dp_hist = nullptr, *dp_step = nullptr;
  cudaStream_t dp_stream = nullptr; cudaEvent_t dp_fork = nullptr, dp_join = nullptr; bool dp_pending = false;
  // device-side sampling (vcla_set_sampler): replaces the argmax by the fused logits-processor chain + draw
  SamplerParams* samp_params = nullptr; float* samp_logits = nullptr; int32_t* finished = nullptr; bool samp_on = false;
  int64_t len_bound = 0;   // host-side upper bound of the cached tokens per sequence (prefill S + decode steps issued since)
  // vision activations
  bf16 *v_im2col = nullptr, *v_norm = nullptr, *v_qkv = nullptr, *v_attn = nullptr, *v_ffn = nullptr;
  float *v_hidden = nullptr, *v_postln_f32 = nullptr;
  float *r_hidden = nullptr, *img_embeds = nullptr;
  bf16 *r_hidden_bf16 = nullptr, *r_qkv = nullptr, *r_kvimg = nullptr, *r_ctx = nullptr, *r_ffn = nullptr;
  // prefill activations
  float* resid = nullptr; bf16 *xn = nullptr, *qkv = nullptr, *attn = nullptr, *hmid = nullptr;
  float* p_ssq = nullptr;    // [max_prefill_tokens][t_hidden / 64] row statistics of the deferred-RMSNorm prefill schedule
  int prefill_fused = 1;     // VCLA_PREFILL_FUSED=0: the 8-kernel/layer schedule with separate rmsnorm / rope_and_cache kernels
  // decode activations
  float* d_resid = nullptr; bf16 *d_xn = nullptr, *d_attn = nullptr, *d_h = nullptr;
  float *ws_qkv = nullptr, *ws_o = nullptr, *ws_gu = nullptr, *ws_d = nullptr, *ws_lm = nullptr;
  float* attn_scratch = nullptr; int32_t* attn_counters = nullptr;
  int32_t* d_tok = nullptr;
  int32_t *tok_hist = nullptr, *step_idx = nullptr;
  float *d_rstd = nullptr, *d_ssq = nullptr; int32_t *cnt_o = nullptr, *cnt_gu = nullptr, *cnt_d = nullptr;   // fused split-K consumers (decode)
  // decode schedule: 2 (default, batch <= 32) = cluster split-K GEMMs with fused consumers, 5 kernels / layer (gemm_decode.cu);
  // 0 = split-K partials in an L2 workspace + separate consumer kernels, 8 kernels / layer; 1 = VCLA_FUSED_DECODE (below).
  int decode_schedule = 2;
  int csk_qkv = 0, csk_o = 0, csk_gu = 0, csk_d = 0, csk_lm = 0;   // CTAs per cluster (= K splits), chosen per batch on first use
  int csk_batch = 0;
  int fused_decode = 0;   // VCLA_FUSED_DECODE=1: 5-kernel/layer schedule with in-GEMM split-K fixup. Correct (parity-tested) but measured slower
                          // on B200 (3.96 vs 3.32 ms/token at B=8): the fence->atomic->reload chain per tile outlasts a kernel boundary.   // tokens of every step since the last prefill, appended by the argmax kernel
  int sp_qkv = 1, sp_o = 1, sp_gu = 1, sp_d = 1, sp_lm = 1, kv_splits = 1;
  int l2_prefetch_kb = 0;    // decode GEMMs: weight k-blocks per CTA prefetched into L2 during the dependency wait (VCLA_L2_PREFETCH_KB).
                             // Measured harmful on B200 (0/8/16/24/48 -> 3142/3226/3349/3390/3437 us per step: the prefetch traffic delays
                             // the latency-critical consumer kernel in front of the GEMM), so it is off by default.
  // graphs
  std::map<GraphKey, cudaGraphExec_t> graphs;
  std::map<GraphKey, int64_t> graph_launches;
  std::list<GraphKey> graph_lru;          // most recently used first; bounded (kMaxGraphs) so a caller with ever-new buffers cannot grow it
  int64_t launches = 0;
  void* staging = nullptr; size_t staging_bytes = 0;
  void* trace_buf = nullptr; unsigned long long trace_cap = 0;
  cudaStream_t cap_stream = nullptr;   // graph capture is illegal on the legacy default stream torch uses
};

namespace {

template <typename T>
T* w_alloc(vcla_ctx* c, size_t n) {
  size_t bytes = align_up(n * sizeof(T), 256);
  if (c->w_arena == nullptr) { c->w_off += bytes; return nullptr; }   // sizing pass
  T* p = reinterpret_cast<T*>(c->w_arena + c->w_off);
  c->w_off += bytes;
  return p;
}
template <typename T>
T* a_alloc(vcla_ctx* c, size_t n) {
  size_t bytes = align_up(n * sizeof(T), 1024);
  if (c->a_arena == nullptr) { c->a_off += bytes; return nullptr; }
  T* p = reinterpret_cast<T*>(c->a_arena + c->a_off);
  c->a_off += bytes;
  return p;
}

void add_slot(vcla_ctx* c, const std::string& name, std::initializer_list<int64_t> shape, int kind, void* dst, int ld, double std_,
              float mean_, int layout = LAY_PLAIN, int which = 0, void* dst2 = nullptr) {
  if (c->w_arena == nullptr) return;  // sizing pass: no registry
  Slot s;
  s.name = name; s.kind = kind; s.layout = layout; s.which = which; s.dst = dst; s.ld = ld; s.dst2 = dst2; s.std = std_; s.mean = mean_;
  s.ndim = (int)shape.size();
  int i = 0; int64_t numel = 1;
  for (int64_t d : shape) { s.shape[i++] = d; numel *= d; }
  s.rows = shape.size() ? *shape.begin() : 1;
  s.cols = s.rows ? numel / s.rows : 0;
  if (kind == SLOT_VEC) { s.rows = 1; s.cols = numel; }
  c->slot_index[name] = (int)c->slots.size();
  c->slots.push_back(s);
}

// Lays out every weight in the arena.  Called twice: once with w_arena == nullptr to size, once to assign.
// std / mean per tensor follow oracle/visualcla_oracle.py:weight_specs exactly.
void layout_weights(vcla_ctx* c) {
  const vcla_config& g = c->cfg;
  c->w_off = 0;
  const int D = g.v_hidden, Fv = g.v_ffn;
  const std::string vp = "vision_model.vision_model.";
  c->cls = w_alloc<float>(c, D);
  add_slot(c, vp + "embeddings.class_embedding", {D}, SLOT_VEC, c->cls, 0, 1.0, 0.f);
  c->patch_w = w_alloc<bf16>(c, (size_t)D * c->kpad);
  add_slot(c, vp + "embeddings.patch_embedding.weight", {D, 3, g.v_patch, g.v_patch}, SLOT_MAT, c->patch_w, c->kpad, 1.0 / sqrt((double)c->kpatch), 0.f);
  c->pos = w_alloc<float>(c, (size_t)c->v_tokens * D);
  add_slot(c, vp + "embeddings.position_embedding.weight", {c->v_tokens, D}, SLOT_VEC, c->pos, 0, 0.5, 0.f);
  c->pre_w = w_alloc<float>(c, D); add_slot(c, vp + "pre_layrnorm.weight", {D}, SLOT_VEC, c->pre_w, 0, 0.1, 1.f);
  c->pre_b = w_alloc<float>(c, D); add_slot(c, vp + "pre_layrnorm.bias", {D}, SLOT_VEC, c->pre_b, 0, 0.1, 0.f);
  c->vl.resize(g.v_layers);
  for (int i = 0; i < g.v_layers; ++i) {
    VisionLayer& L = c->vl[i];
    const std::string lp = vp + "encoder.layers." + std::to_string(i) + ".";
    L.ln1_w = w_alloc<float>(c, D); add_slot(c, lp + "layer_norm1.weight", {D}, SLOT_VEC, L.ln1_w, 0, 0.1, 1.f);
    L.ln1_b = w_alloc<float>(c, D); add_slot(c, lp + "layer_norm1.bias", {D}, SLOT_VEC, L.ln1_b, 0, 0.1, 0.f);
    L.ln2_w = w_alloc<float>(c, D); add_slot(c, lp + "layer_norm2.weight", {D}, SLOT_VEC, L.ln2_w, 0, 0.1, 1.f);
    L.ln2_b = w_alloc<float>(c, D); add_slot(c, lp + "layer_norm2.bias", {D}, SLOT_VEC, L.ln2_b, 0, 0.1, 0.f);
    L.wqkv = w_alloc<bf16>(c, (size_t)3 * D * D);
    L.bqkv = w_alloc<float>(c, 3 * D);
    const char* pr[3] = {"q_proj", "k_proj", "v_proj"};
    for (int j = 0; j < 3; ++j) {
      add_slot(c, lp + "self_attn." + pr[j] + ".weight", {D, D}, SLOT_MAT, L.wqkv ? L.wqkv + (size_t)j * D * D : nullptr, D, 1.5 / sqrt((double)D), 0.f);
      add_slot(c, lp + "self_attn." + pr[j] + ".bias", {D}, SLOT_VEC, L.bqkv ? L.bqkv + j * D : nullptr, 0, 0.1, 0.f);
    }
    L.wo = w_alloc<bf16>(c, (size_t)D * D); add_slot(c, lp + "self_attn.out_proj.weight", {D, D}, SLOT_MAT, L.wo, D, 0.5 / sqrt((double)D), 0.f);
    L.bo = w_alloc<float>(c, D); add_slot(c, lp + "self_attn.out_proj.bias", {D}, SLOT_VEC, L.bo, 0, 0.05, 0.f);
    L.w1 = w_alloc<bf16>(c, (size_t)Fv * D); add_slot(c, lp + "mlp.fc1.weight", {Fv, D}, SLOT_MAT, L.w1, D, 1.0 / sqrt((double)D), 0.f);
    L.b1 = w_alloc<float>(c, Fv); add_slot(c, lp + "mlp.fc1.bias", {Fv}, SLOT_VEC, L.b1, 0, 0.1, 0.f);
    L.w2 = w_alloc<bf16>(c, (size_t)D * Fv); add_slot(c, lp + "mlp.fc2.weight", {D, Fv}, SLOT_MAT, L.w2, Fv, 0.5 / sqrt((double)Fv), 0.f);
    L.b2 = w_alloc<float>(c, D); add_slot(c, lp + "mlp.fc2.bias", {D}, SLOT_VEC, L.b2, 0, 0.05, 0.f);
  }
  c->post_w = w_alloc<float>(c, D); add_slot(c, vp + "post_layernorm.weight", {D}, SLOT_VEC, c->post_w, 0, 0.1, 1.f);
  c->post_b = w_alloc<float>(c, D); add_slot(c, vp + "post_layernorm.bias", {D}, SLOT_VEC, c->post_b, 0, 0.1, 0.f);

  const int R = g.r_hidden, Fr = g.r_ffn, Q = g.r_queries, RL = g.r_layers;
  const std::string rp = "visual_resampler.";
  c->rq = w_alloc<float>(c, (size_t)Q * R); add_slot(c, rp + "query_embeddding", {1, Q, R}, SLOT_VEC, c->rq, 0, 1.0, 0.f);
  c->r_wkv_all = w_alloc<bf16>(c, (size_t)RL * 2 * R * R);
  c->r_bkv_all = w_alloc<float>(c, (size_t)RL * 2 * R);
  c->rl.resize(RL);
  for (int i = 0; i < RL; ++i) {
    ResamplerLayer& L = c->rl[i];
    const std::string lp = rp + "encoder.layer." + std::to_string(i) + ".";
    L.wqkv = w_alloc<bf16>(c, (size_t)3 * R * R);
    L.bqkv = w_alloc<float>(c, 3 * R);
    const char* pr[3] = {"query", "key", "value"};
    for (int j = 0; j < 3; ++j) {
      // key/value additionally live in the all-layer [RL*2R, R] matrix used for the layer-invariant image rows
      void* d2w = (j > 0 && c->r_wkv_all) ? (void*)(c->r_wkv_all + ((size_t)i * 2 + (j - 1)) * R * R) : nullptr;
      void* d2b = (j > 0 && c->r_bkv_all) ? (void*)(c->r_bkv_all + ((size_t)i * 2 + (j - 1)) * R) : nullptr;
      add_slot(c, lp + "crossattention.self." + pr[j] + ".weight", {R, R}, SLOT_MAT, L.wqkv ? L.wqkv + (size_t)j * R * R : nullptr, R, 1.5 / sqrt((double)R), 0.f, LAY_PLAIN, 0, d2w);
      add_slot(c, lp + "crossattention.self." + pr[j] + ".bias", {R}, SLOT_VEC, L.bqkv ? L.bqkv + j * R : nullptr, 0, 0.1, 0.f, LAY_PLAIN, 0, d2b);
    }
    L.wo = w_alloc<bf16>(c, (size_t)R * R); add_slot(c, lp + "crossattention.output.dense.weight", {R, R}, SLOT_MAT, L.wo, R, 1.0 / sqrt((double)R), 0.f);
    L.bo = w_alloc<float>(c, R); add_slot(c, lp + "crossattention.output.dense.bias", {R}, SLOT_VEC, L.bo, 0, 0.05, 0.f);
    L.ln1_w = w_alloc<float>(c, R); add_slot(c, lp + "crossattention.output.LayerNorm.weight", {R}, SLOT_VEC, L.ln1_w, 0, 0.1, 1.f);
    L.ln1_b = w_alloc<float>(c, R); add_slot(c, lp + "crossattention.output.LayerNorm.bias", {R}, SLOT_VEC, L.ln1_b, 0, 0.1, 0.f);
    L.wi = w_alloc<bf16>(c, (size_t)Fr * R); add_slot(c, lp + "intermediate.dense.weight", {Fr, R}, SLOT_MAT, L.wi, R, 1.0 / sqrt((double)R), 0.f);
    L.bi = w_alloc<float>(c, Fr); add_slot(c, lp + "intermediate.dense.bias", {Fr}, SLOT_VEC, L.bi, 0, 0.1, 0.f);
    L.wo2 = w_alloc<bf16>(c, (size_t)R * Fr); add_slot(c, lp + "output.dense.weight", {R, Fr}, SLOT_MAT, L.wo2, Fr, 1.0 / sqrt((double)Fr), 0.f);
    L.bo2 = w_alloc<float>(c, R); add_slot(c, lp + "output.dense.bias", {R}, SLOT_VEC, L.bo2, 0, 0.05, 0.f);
    L.ln2_w = w_alloc<float>(c, R); add_slot(c, lp + "output.LayerNorm.weight", {R}, SLOT_VEC, L.ln2_w, 0, 0.1, 1.f);
    L.ln2_b = w_alloc<float>(c, R); add_slot(c, lp + "output.LayerNorm.bias", {R}, SLOT_VEC, L.ln2_b, 0, 0.1, 0.f);
  }
  const int T = g.t_hidden, Ft = g.t_ffn, V = g.t_vocab, TL = g.t_layers;
  c->proj_w = w_alloc<bf16>(c, (size_t)T * R); add_slot(c, "image_projection_layer.weight", {T, R}, SLOT_MAT, c->proj_w, R, 1.0 / sqrt((double)R), 0.f);
  c->proj_b = w_alloc<float>(c, T); add_slot(c, "image_projection_layer.bias", {T}, SLOT_VEC, c->proj_b, 0, 0.1, 0.f);

  const std::string tp = "text_model.model.";
  const double res_gain = 1.0 / sqrt(2.0 * (double)TL);
  c->embed = w_alloc<bf16>(c, (size_t)V * T); add_slot(c, tp + "embed_tokens.weight", {V, T}, SLOT_MAT, c->embed, T, 1.0, 0.f);
  c->tl.resize(TL);
  for (int i = 0; i < TL; ++i) {
    TextLayer& L = c->tl[i];
    const std::string lp = tp + "layers." + std::to_string(i) + ".";
    L.ln1 = w_alloc<float>(c, T); add_slot(c, lp + "input_layernorm.weight", {T}, SLOT_VEC, L.ln1, 0, 0.1, 1.f);
    L.ln2 = w_alloc<float>(c, T); add_slot(c, lp + "post_attention_layernorm.weight", {T}, SLOT_VEC, L.ln2, 0, 0.1, 1.f);
    L.wqkv = w_alloc<bf16>(c, (size_t)3 * T * T);
    add_slot(c, lp + "self_attn.q_proj.weight", {T, T}, SLOT_MAT, L.wqkv, T, 1.5 / sqrt((double)T), 0.f);
    add_slot(c, lp + "self_attn.k_proj.weight", {T, T}, SLOT_MAT, L.wqkv ? L.wqkv + (size_t)T * T : nullptr, T, 1.5 / sqrt((double)T), 0.f);
    add_slot(c, lp + "self_attn.v_proj.weight", {T, T}, SLOT_MAT, L.wqkv ? L.wqkv + (size_t)2 * T * T : nullptr, T, 1.0 / sqrt((double)T), 0.f);
    L.wo = w_alloc<bf16>(c, (size_t)T * T); add_slot(c, lp + "self_attn.o_proj.weight", {T, T}, SLOT_MAT, L.wo, T, res_gain * 2.0 / sqrt((double)T), 0.f);
    L.wgu = w_alloc<bf16>(c, (size_t)2 * Ft * T);
    add_slot(c, lp + "mlp.gate_proj.weight", {Ft, T}, SLOT_MAT, L.wgu, T, 1.0 / sqrt((double)T), 0.f, LAY_INTERLEAVE32, 0);
    add_slot(c, lp + "mlp.up_proj.weight", {Ft, T}, SLOT_MAT, L.wgu, T, 1.0 / sqrt((double)T), 0.f, LAY_INTERLEAVE32, 1);
    L.wd = w_alloc<bf16>(c, (size_t)T * Ft); add_slot(c, lp + "mlp.down_proj.weight", {T, Ft}, SLOT_MAT, L.wd, Ft, res_gain * 4.0 / sqrt((double)Ft), 0.f);
  }
  c->final_norm = w_alloc<float>(c, T); add_slot(c, tp + "norm.weight", {T}, SLOT_VEC, c->final_norm, 0, 0.1, 1.f);
  c->lm_head = w_alloc<bf16>(c, (size_t)V * T); add_slot(c, "text_model.lm_head.weight", {V, T}, SLOT_MAT, c->lm_head, T, 4.0 / sqrt((double)T), 0.f);
}

void layout_activations(vcla_ctx* c) {
  const vcla_config& g = c->cfg;
  c->a_off = 0;
  const size_t Bv = g.max_batch, VT = (size_t)Bv * c->v_tokens, D = g.v_hidden, gg = (size_t)(c->v_tokens - 1);
  c->v_im2col = a_alloc<bf16>(c, Bv * gg * c->kpad);
  c->v_hidden = a_alloc<float>(c, VT * D);
  c->v_norm = a_alloc<bf16>(c, VT * D);
  c->v_qkv = a_alloc<bf16>(c, VT * 3 * D);
  c->v_attn = a_alloc<bf16>(c, VT * D);
  c->v_ffn = a_alloc<bf16>(c, VT * g.v_ffn);
  c->v_postln_f32 = a_alloc<float>(c, VT * D);
  const size_t RQ = (size_t)Bv * g.r_queries, R = g.r_hidden;
  c->r_hidden = a_alloc<float>(c, RQ * R);
  c->r_hidden_bf16 = a_alloc<bf16>(c, RQ * R);
  c->r_qkv = a_alloc<bf16>(c, RQ * 3 * R);
  c->r_kvimg = a_alloc<bf16>(c, VT * (size_t)g.r_layers * 2 * R);
  c->r_ctx = a_alloc<bf16>(c, RQ * R);
  c->r_ffn = a_alloc<bf16>(c, RQ * g.r_ffn);
  c->img_embeds = a_alloc<float>(c, RQ * g.t_hidden);
  const size_t Tk = g.max_prefill_tokens, T = g.t_hidden, F = g.t_ffn;
  c->resid = a_alloc<float>(c, Tk * T);
  c->xn = a_alloc<bf16>(c, Tk * T);
  c->qkv = a_alloc<bf16>(c, Tk * 3 * T);
  c->attn = a_alloc<bf16>(c, Tk * T);
  c->hmid = a_alloc<bf16>(c, Tk * F);
  c->p_ssq = a_alloc<float>(c, Tk * ((T + 63) / 64));
  const size_t Bp = 64;  // decode operand rows (batch is processed in chunks of <= 64)
  c->d_resid = a_alloc<float>(c, Bp * T);
  c->d_xn = a_alloc<bf16>(c, Bp * T);
  c->d_attn = a_alloc<bf16>(c, Bp * T);
  c->d_h = a_alloc<bf16>(c, Bp * F);
  c->ws_qkv = a_alloc<float>(c, (size_t)c->sp_qkv * Bp * 3 * T);
  c->ws_o = a_alloc<float>(c, (size_t)c->sp_o * Bp * T);
  c->ws_gu = a_alloc<float>(c, (size_t)c->sp_gu * Bp * 2 * F);
  c->ws_d = a_alloc<float>(c, (size_t)c->sp_d * Bp * T);
  c->ws_lm = a_alloc<float>(c, (size_t)c->sp_lm * Bp * g.t_vocab);
  c->attn_scratch = a_alloc<float>(c, (size_t)Bp * g.t_heads * 8 * (128 + 2));   // up to 8 KV splits
  c->attn_counters = a_alloc<int32_t>(c, (size_t)Bp * g.t_heads);
  c->d_tok = a_alloc<int32_t>(c, Bp);
  c->tok_hist = a_alloc<int32_t>(c, (size_t)(g.max_seq + 2) * Bp);
  c->step_idx = a_alloc<int32_t>(c, 16);
  c->d_rstd = a_alloc<float>(c, Bp);
  c->d_ssq = a_alloc<float>(c, Bp * ((T + 127) / 128));
  c->cnt_o = a_alloc<int32_t>(c, (T + 127) / 128 + 1);
  c->cnt_d = a_alloc<int32_t>(c, (T + 127) / 128 + 1);
  c->cnt_gu = a_alloc<int32_t>(c, (2 * F + 127) / 128 + 1);
  c->page_table = a_alloc<int32_t>(c, (size_t)g.max_batch * c->pages_per_seq);
  c->seq_len = a_alloc<int32_t>(c, g.max_batch);
  c->img_row_default = a_alloc<int32_t>(c, g.max_batch);
  c->kv_free = a_alloc<int32_t>(c, (size_t)c->total_pages);
  c->kv_order = a_alloc<int32_t>(c, (size_t)c->total_pages);
  c->kv_state = a_alloc<int32_t>(c, 4);
  c->kv_npages = a_alloc<int32_t>(c, g.max_batch);
  c->rope_cos = a_alloc<float>(c, (size_t)(g.max_seq + 1) * 64);
  c->rope_sin = a_alloc<float>(c, (size_t)(g.max_seq + 1) * 64);
  c->cand_val = a_alloc<float>(c, Bp * kArgmaxChunks);
  c->cand_idx = a_alloc<int32_t>(c, Bp * kArgmaxChunks);
  c->samp_params = a_alloc<SamplerParams>(c, 1);
  c->samp_logits = a_alloc<float>(c, Bp * (size_t)g.t_vocab);
  c->finished = a_alloc<int32_t>(c, Bp);
}

// Split-K factor of a decode GEMM (row tiles of 128 x `splits` work units on 2 persistent CTAs per SM).  Measured on B200
// (profiles/r1_decode_step_trace_*.json): a thin last wave is latency-bound (one lone CTA streams ~80 GB/s), and a single
// wave of one-tile CTAs loses the epilogue/load overlap -> prefer >= ~2 waves with a last wave that is >= 60 % full.
int pick_splits(int n_out, int K) {
  const int tiles = (n_out + 127) / 128;
  const int kb = (K + 63) / 64;
  const double slots = 2.0 * num_sms();
  int best = 1;
  double best_score = 1e9;
  for (int want = 1; want <= 40; ++want) {
    if (kb / want < 4 && want > 1) break;               // at least 4 k-blocks per work unit
    const int per = (kb + want - 1) / want;
    const int s = (kb + per - 1) / per;                 // realisable: every split non-empty
    if (s != want) continue;
    const double w = tiles * (double)s / slots;
    const double f = w - floor(w);
    double score;
    if (w < 1.0) score = (1.0 - w) + 0.15;
    else if (f < 1e-9) score = 0.0;
    else if (f >= 0.6) score = (1.0 - f) * 0.3;
    else score = (0.6 - f) + 0.2;
    score += 0.01 * s;                                  // fewer partials for the consumers when otherwise equal
    if (score < best_score) { best_score = score; best = s; }
  }
  return best;
}

int count(vcla_ctx* c, int n = 1) { c->launches += n; return 0; }

}  // namespace

// ---- NCCL, bound at run time ------------------------------------------------------------------------------------
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
// NCCL is bound at run time (dlopen by soname: inside a torch process this resolves to the libnccl torch already loaded), so
// libvcla.so itself has no link-time dependency on it and single-GPU users never touch it.
static int nccl_api() {
  if (g_nccl.lib) return 0;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { set_error("NCCL not found (dlopen libnccl.so.2: %s)", dlerror()); return -1; }
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
  g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(h, "ncclAllGather");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.CommDestroy || !g_nccl.GetErrorString) {
    set_error("NCCL library lacks an expected symbol"); return -1;
  }
  g_nccl.lib = h;
  return 0;
}
#define VCLA_NCCL_OK(expr)                                                                                   \
  do {                                                                                                       \
    ncclResult_t _r = (expr);                                                                                \
    if (_r != ncclSuccess) { set_error("%s failed: %s", #expr, g_nccl.GetErrorString(_r)); return -1; }      \
  } while (0)


// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

const char* vcla_last_error(void) { return get_error(); }
const char* vcla_version(void) { return "vcla-b200 0.1 (sm_100a, tcgen05/TMA)"; }
void vcla_set_pdl(int on) { set_pdl(on != 0); }

int vcla_create(const vcla_config* cfg, vcla_ctx** out) {
  if (!cfg || !out) { set_error("vcla_create: null argument"); return -1; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("vcla_create: no CUDA device (this library has no CPU fallback)");
    return -1;
  }
  vcla_ctx* c = new vcla_ctx();
  c->cfg = *cfg;
  vcla_config& g = c->cfg;
  if (g.page_tokens <= 0) g.page_tokens = 64;
  auto bad = [&](const char* why) { set_error("vcla_create: %s", why); delete c; return -1; };
  if (g.t_hidden % g.t_heads || g.t_hidden / g.t_heads != 128) return bad("LLaMA head_dim must be 128");
  if (g.v_hidden % g.v_heads || g.v_hidden / g.v_heads != 64) return bad("ViT head_dim must be 64");
  if (g.r_hidden % g.r_heads || g.r_hidden / g.r_heads != 64) return bad("Resampler head_dim must be 64");
  if (g.t_ffn % 32) return bad("LLaMA ffn must be a multiple of 32");
  if (g.v_hidden % 8 || g.r_hidden % 8 || g.t_hidden % 8 || g.v_ffn % 8 || g.r_ffn % 8 || g.t_ffn % 8) return bad("hidden sizes must be multiples of 8");
  if (g.v_image % g.v_patch) return bad("image size must be a multiple of the patch size");
  if (g.max_batch < 1 || g.max_seq < 1 || g.max_prefill_tokens < 1) return bad("capacities must be positive");
  if (g.max_seq > 1 << 20) return bad("max_seq too large");
  if (g.page_tokens > 64 || g.page_tokens % 8) return bad("page_tokens must be a multiple of 8 and <= 64");
  c->v_tokens = (g.v_image / g.v_patch) * (g.v_image / g.v_patch) + 1;
  c->kpatch = 3 * g.v_patch * g.v_patch;
  c->kpad = (c->kpatch + 63) / 64 * 64;
  c->hd_t = 128;
  c->page_tokens = g.page_tokens;
  c->pages_per_seq = (g.max_seq + c->page_tokens - 1) / c->page_tokens;
  c->total_pages = c->pages_per_seq * g.max_batch;
  c->sp_qkv = pick_splits(3 * g.t_hidden, g.t_hidden);
  c->sp_o = pick_splits(g.t_hidden, g.t_hidden);
  c->sp_gu = pick_splits(2 * g.t_ffn, g.t_hidden);
  c->sp_d = pick_splits(g.t_hidden, g.t_ffn);
  c->sp_lm = pick_splits(g.t_vocab, g.t_hidden);
  if (const char* e = getenv("VCLA_L2_PREFETCH_KB")) c->l2_prefetch_kb = atoi(e);
  if (const char* e = getenv("VCLA_FUSED_DECODE")) c->fused_decode = atoi(e);
  if (const char* e = getenv("VCLA_PREFILL_FUSED")) c->prefill_fused = atoi(e);
  if (const char* e = getenv("VCLA_DECODE_SCHEDULE")) c->decode_schedule = !strcmp(e, "unfused") ? 0 : (!strcmp(e, "fix") ? 1 : 2);
  if (c->fused_decode) c->decode_schedule = 1;
  c->fused_decode = c->decode_schedule == 1;
  c->kv_splits = g.max_seq >= 1536 ? 4 : (g.max_seq >= 768 ? 2 : 1);   // context-driven minimum; raised per call for small batches
  if (const char* e = getenv("VCLA_KV_SPLITS")) { const int v = atoi(e); if (v >= 1 && v <= 8) c->kv_splits = v; }   // tuning override

  if (gemm_init()) { delete c; return -1; }
  // sizing passes
  layout_weights(c); c->w_bytes = c->w_off;
  layout_activations(c); c->a_bytes = c->a_off;
  c->kv_layer_elems = (size_t)c->total_pages * 2 * g.t_heads * c->page_tokens * 128;
  c->kv_bytes = c->kv_layer_elems * g.t_layers * sizeof(bf16);
  cudaError_t e;
  if ((e = cudaMalloc(&c->w_arena, c->w_bytes)) != cudaSuccess || (e = cudaMalloc(&c->a_arena, c->a_bytes)) != cudaSuccess ||
      (e = cudaMalloc(&c->kv_arena, c->kv_bytes)) != cudaSuccess) {
    set_error("vcla_create: cudaMalloc failed (%s): weights %.2f GB, activations %.2f GB, kv %.2f GB", cudaGetErrorString(e),
              c->w_bytes / 1e9, c->a_bytes / 1e9, c->kv_bytes / 1e9);
    vcla_destroy(c);
    return -1;
  }
  cudaMemset(c->w_arena, 0, c->w_bytes);
  cudaMemset(c->a_arena, 0, c->a_bytes);
  layout_weights(c);
  layout_activations(c);
  for (int i = 0; i < g.t_layers; ++i) c->tl[i].kv = c->kv_arena + (size_t)i * c->kv_layer_elems;
  // page allocation order: physical pages 0,1,2,... are handed out in this order after every reset (vcla_kv_debug_shuffle permutes it)
  std::vector<int32_t> order((size_t)c->total_pages);
  for (size_t i = 0; i < order.size(); ++i) order[i] = (int32_t)i;
  cudaMemcpy(c->kv_order, order.data(), order.size() * 4, cudaMemcpyHostToDevice);
  std::vector<int32_t> two(g.max_batch, 2);
  cudaMemcpy(c->img_row_default, two.data(), two.size() * 4, cudaMemcpyHostToDevice);
  if (const char* e = getenv("VCLA_ATTN_PERSISTENT")) c->attn_persistent_mode = atoi(e);
  if (const char* e = getenv("VCLA_ATTN_PERSISTENT_GRID")) c->attn_persistent_grid = atoi(e);
  if (rope_fill_tables(g.max_seq + 1, 128, g.rope_theta, c->rope_cos, c->rope_sin) || attention_init() || sampler_init() || vcla_reset(c, nullptr)) { vcla_destroy(c); return -1; }
  if (cudaDeviceSynchronize() != cudaSuccess) { set_error("vcla_create: device error %s", cudaGetErrorString(cudaGetLastError())); vcla_destroy(c); return -1; }
  *out = c;
  return 0;
}

void vcla_destroy(vcla_ctx* c) {
  if (!c) return;
  for (auto& kv : c->graphs) cudaGraphExecDestroy(kv.second);
  if (c->w_arena) cudaFree(c->w_arena);
  if (c->a_arena) cudaFree(c->a_arena);
  if (c->kv_arena) cudaFree(c->kv_arena);
  if (c->staging) cudaFree(c->staging);
  c->graphs.clear();
  if (c->comm) { cudaDeviceSynchronize(); g_nccl.CommDestroy(c->comm); c->comm = nullptr; }
  if (c->dp_send) cudaFree(c->dp_send);
  if (c->dp_stream) cudaStreamDestroy(c->dp_stream);
  if (c->dp_fork) cudaEventDestroy(c->dp_fork);
  if (c->dp_join) cudaEventDestroy(c->dp_join);
  if (c->cap_stream) cudaStreamDestroy(c->cap_stream);
  if (c->trace_buf) { trace_set_gemm(nullptr, 0); trace_set_attention_tc(nullptr, 0); trace_set_gemm_decode(nullptr, 0); trace_set_sampler(nullptr, 0); trace_set_attention(nullptr, 0); trace_set_elementwise(nullptr, 0); cudaFree(c->trace_buf); }
  delete c;
}

int vcla_get_config(const vcla_ctx* c, vcla_config* out) { if (!c || !out) return -1; *out = c->cfg; return 0; }
int vcla_memory_bytes(const vcla_ctx* c, int64_t* w, int64_t* kv, int64_t* a) {
  if (!c) return -1;
  if (w) *w = (int64_t)c->w_bytes;
  if (kv) *kv = (int64_t)c->kv_bytes;
  if (a) *a = (int64_t)c->a_bytes;
  return 0;
}
int64_t vcla_kernel_launches(vcla_ctx* c, int reset) { int64_t v = c->launches; if (reset) c->launches = 0; return v; }

int vcla_weight_count(const vcla_ctx* c) { return c ? (int)c->slots.size() : 0; }
int vcla_weight_info(const vcla_ctx* c, int index, const char** name, int64_t shape[4], int* ndim, int* kind) {
  if (!c || index < 0 || index >= (int)c->slots.size()) { set_error("vcla_weight_info: bad index"); return -1; }
  const Slot& s = c->slots[index];
  if (name) *name = s.name.c_str();
  if (shape) for (int i = 0; i < 4; ++i) shape[i] = s.shape[i];
  if (ndim) *ndim = s.ndim;
  if (kind) *kind = s.kind;
  return 0;
}

static int ensure_staging(vcla_ctx* c, size_t bytes) {
  if (c->staging_bytes >= bytes) return 0;
  if (c->staging) cudaFree(c->staging);
  c->staging = nullptr; c->staging_bytes = 0;
  VCLA_CUDA_OK(cudaMalloc(&c->staging, bytes));
  c->staging_bytes = bytes;
  return 0;
}

// place a contiguous bf16 [rows, cols] device matrix into a slot's storage
static int place_matrix(const Slot& s, const bf16* src, cudaStream_t st) {
  if (s.layout == LAY_INTERLEAVE32) return interleave_rows32(src, (int)s.rows, (int)s.cols, s.which, (bf16*)s.dst, st);
  if (copy_rows_bf16(src, (int)s.rows, (int)s.cols, (bf16*)s.dst, s.ld, st)) return -1;
  if (s.dst2) return copy_rows_bf16(src, (int)s.rows, (int)s.cols, (bf16*)s.dst2, (int)s.cols, st);
  return 0;
}

int vcla_load_weight(vcla_ctx* c, const char* name, const void* src, int dtype, int64_t numel, int on_device, vcla_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  auto it = c->slot_index.find(name);
  if (it == c->slot_index.end()) { set_error("vcla_load_weight: unknown tensor '%s'", name); return -1; }
  const Slot& s = c->slots[it->second];
  const size_t n = (size_t)s.rows * s.cols;
  if (src == nullptr || numel != (int64_t)n) {
    set_error("vcla_load_weight: '%s' holds %lld elements, the caller passed %lld (checkpoint / config shape mismatch)", name, (long long)n, (long long)numel);
    return -1;
  }
  const size_t esz = dtype == VCLA_F32 ? 4 : 2;
  if (dtype < 0 || dtype > 2) { set_error("vcla_load_weight: bad dtype"); return -1; }
  // staging: [raw source copy][bf16 contiguous]
  const size_t raw_bytes = align_up(n * esz, 256);
  if (ensure_staging(c, raw_bytes + n * 2 + 256)) return -1;
  const void* dsrc = src;
  if (!on_device) {
    VCLA_CUDA_OK(cudaMemcpyAsync(c->staging, src, n * esz, cudaMemcpyHostToDevice, st));
    dsrc = c->staging;
  }
  if (s.kind == SLOT_VEC) {
    if (convert_to_f32(dsrc, dtype, (int64_t)n, (float*)s.dst, st)) return -1;
    if (s.dst2 && convert_to_f32(dsrc, dtype, (int64_t)n, (float*)s.dst2, st)) return -1;
  } else {
    bf16* tmp = reinterpret_cast<bf16*>((uint8_t*)c->staging + raw_bytes);
    if (convert_to_bf16(dsrc, dtype, (int64_t)n, tmp, st)) return -1;
    if (place_matrix(s, tmp, st)) return -1;
  }
  VCLA_CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

int vcla_read_weight(vcla_ctx* c, const char* name, void* dst_host, vcla_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  auto it = c->slot_index.find(name);
  if (it == c->slot_index.end()) { set_error("vcla_read_weight: unknown tensor '%s'", name); return -1; }
  const Slot& s = c->slots[it->second];
  if (s.kind == SLOT_VEC) {
    VCLA_CUDA_OK(cudaMemcpyAsync(dst_host, s.dst, (size_t)s.cols * 4, cudaMemcpyDeviceToHost, st));
  } else if (s.layout == LAY_INTERLEAVE32) {
    for (int64_t j0 = 0; j0 < s.rows; j0 += 32) {
      const int64_t nr = (s.rows - j0) < 32 ? (s.rows - j0) : 32;
      const bf16* srcp = (const bf16*)s.dst + ((j0 / 32) * 64 + (int64_t)s.which * 32) * s.cols;
      VCLA_CUDA_OK(cudaMemcpyAsync((bf16*)dst_host + j0 * s.cols, srcp, (size_t)nr * s.cols * 2, cudaMemcpyDeviceToHost, st));
    }
  } else {
    VCLA_CUDA_OK(cudaMemcpy2DAsync(dst_host, (size_t)s.cols * 2, s.dst, (size_t)s.ld * 2, (size_t)s.cols * 2, (size_t)s.rows, cudaMemcpyDeviceToHost, st));
  }
  VCLA_CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

int vcla_init_synthetic(vcla_ctx* c, uint32_t seed, vcla_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  size_t max_n = 0;
  for (const Slot& s : c->slots) if (s.kind == SLOT_MAT) max_n = std::max(max_n, (size_t)s.rows * s.cols);
  if (ensure_staging(c, max_n * 2 + 256)) return -1;
  const double sigma = 65536.0 / sqrt(3.0);
  for (const Slot& s : c->slots) {
    const int64_t n = s.rows * s.cols;
    const uint32_t sd = fnv1a32(s.name.c_str()) ^ (uint32_t)(seed * 0x9E3779B1u);
    const float mul = (float)(s.std / sigma);
    if (s.kind == SLOT_VEC) {
      if (fill_hash_normal(nullptr, (float*)s.dst, n, sd, mul, s.mean, st)) return -1;
      if (s.dst2 && fill_hash_normal(nullptr, (float*)s.dst2, n, sd, mul, s.mean, st)) return -1;
    } else {
      bf16* tmp = (bf16*)c->staging;
      if (fill_hash_normal(tmp, nullptr, n, sd, mul, s.mean, st)) return -1;
      if (place_matrix(s, tmp, st)) return -1;
    }
  }
  VCLA_CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

int vcla_reset(vcla_ctx* c, vcla_stream stream) {
  VCLA_CUDA_OK(cudaMemsetAsync(c->seq_len, 0, (size_t)c->cfg.max_batch * 4, (cudaStream_t)stream));
  VCLA_CUDA_OK(cudaMemsetAsync(c->attn_counters, 0, (size_t)64 * c->cfg.t_heads * 4, (cudaStream_t)stream));
  VCLA_CUDA_OK(cudaMemsetAsync(c->step_idx, 0, 4, (cudaStream_t)stream));
  // every page back on the free stack, no sequence owns any
  if (kv_reset(c->kv_free, c->kv_order, c->kv_state, c->kv_npages, c->total_pages, c->cfg.max_batch, (cudaStream_t)stream)) return -1;
  if (c->dp_step) VCLA_CUDA_OK(cudaMemsetAsync(c->dp_step, 0, 4, (cudaStream_t)stream));
  VCLA_CUDA_OK(cudaMemsetAsync(c->finished, 0, 64 * 4, (cudaStream_t)stream));
  c->len_bound = 0;
  return 0;
}

int vcla_kv_debug_shuffle(vcla_ctx* c, uint32_t seed) {
  // test hook: permute the order in which physical pages are handed out (Fisher-Yates over an LCG); takes effect at the next reset
  std::vector<int32_t> order((size_t)c->total_pages);
  for (size_t i = 0; i < order.size(); ++i) order[i] = (int32_t)i;
  uint64_t x = 0x9E3779B97F4A7C15ull ^ seed;
  for (size_t i = order.size(); i > 1; --i) {
    x = x * 6364136223846793005ull + 1442695040888963407ull;
    std::swap(order[i - 1], order[(size_t)((x >> 33) % i)]);
  }
  VCLA_CUDA_OK(cudaDeviceSynchronize());
  VCLA_CUDA_OK(cudaMemcpy(c->kv_order, order.data(), order.size() * 4, cudaMemcpyHostToDevice));
  return vcla_reset(c, nullptr);
}

int vcla_kv_read_pages(vcla_ctx* c, int32_t* table_host, int32_t* npages_host, int32_t* state_host) {
  // synchronous copy of the page table [max_batch][pages_per_seq], the per-sequence page counts and {free pages, error flag}
  VCLA_CUDA_OK(cudaDeviceSynchronize());
  if (table_host) VCLA_CUDA_OK(cudaMemcpy(table_host, c->page_table, (size_t)c->cfg.max_batch * c->pages_per_seq * 4, cudaMemcpyDeviceToHost));
  if (npages_host) VCLA_CUDA_OK(cudaMemcpy(npages_host, c->kv_npages, (size_t)c->cfg.max_batch * 4, cudaMemcpyDeviceToHost));
  if (state_host) VCLA_CUDA_OK(cudaMemcpy(state_host, c->kv_state, 8, cudaMemcpyDeviceToHost));
  return 0;
}
int vcla_kv_geometry(const vcla_ctx* c, int* pages_per_seq, int* total_pages, int* page_tokens) {
  if (!c) return -1;
  if (pages_per_seq) *pages_per_seq = c->pages_per_seq;
  if (total_pages) *total_pages = c->total_pages;
  if (page_tokens) *page_tokens = c->page_tokens;
  return 0;
}

// -------------------------------------------------------------------------------------------------
// vision encode
// -------------------------------------------------------------------------------------------------
static int gemm_bf16(vcla_ctx* c, const bf16* A, int M, int K, int lda, const bf16* W, int N, int ldw, const float* bias, int act, bf16* out, int ldo, cudaStream_t st) {
  GemmCall g; g.A = A; g.B = W; g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldb = ldw; g.mode = GEMM_STORE_BF16; g.out = out; g.ldo = ldo; g.bias = bias; g.act = act;
  count(c); return gemm_tc(g, st);
}
static int gemm_f32(vcla_ctx* c, const bf16* A, int M, int K, int lda, const bf16* W, int N, int ldw, const float* bias, int accumulate, float* out, int ldo, cudaStream_t st) {
  GemmCall g; g.A = A; g.B = W; g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldb = ldw; g.mode = GEMM_ADD_F32; g.out = out; g.ldo = ldo; g.bias = bias; g.accumulate = accumulate;
  count(c); return gemm_tc(g, st);
}

int vcla_vision_encode(vcla_ctx* c, const void* pixels, int pixel_dtype, int B, float* out_dev, vcla_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const vcla_config& g = c->cfg;
  if (B < 1 || B > g.max_batch) { set_error("vision_encode: batch %d exceeds capacity %d", B, g.max_batch); return -1; }
  const int D = g.v_hidden, NT = c->v_tokens, NP = NT - 1, rows = B * NT;
  // patch embedding: im2col + GEMM; epilogue adds the position embedding and scatters to token rows 1..NP
  count(c); if (im2col(pixels, pixel_dtype, B, g.v_image, g.v_patch, c->kpad, c->v_im2col, st)) return -1;
  {
    GemmCall gc; gc.A = c->v_im2col; gc.B = c->patch_w; gc.M = B * NP; gc.N = D; gc.K = c->kpad; gc.lda = c->kpad; gc.ldb = c->kpad;
    gc.mode = GEMM_ADD_F32; gc.out = c->v_hidden; gc.ldo = D; gc.rowtab = c->pos + D; gc.rowtab_period = NP;
    gc.rows_per_group = NP; gc.group_stride = NT; gc.row_offset = 1;
    count(c); if (gemm_tc(gc, st)) return -1;
  }
  count(c); if (vit_cls_rows(c->v_hidden, B, NT, D, c->cls, c->pos, st)) return -1;
  count(c); if (layernorm(c->v_hidden, rows, D, c->pre_w, c->pre_b, g.v_eps, nullptr, c->v_hidden, st)) return -1;
  const float vscale = 1.0f / sqrtf(64.f);
  for (int i = 0; i < g.v_layers; ++i) {
    const VisionLayer& L = c->vl[i];
    count(c); if (layernorm(c->v_hidden, rows, D, L.ln1_w, L.ln1_b, g.v_eps, c->v_norm, nullptr, st)) return -1;
    if (gemm_bf16(c, c->v_norm, rows, D, D, L.wqkv, 3 * D, D, L.bqkv, ACT_NONE, c->v_qkv, 3 * D, st)) return -1;
    AttnCall a; a.q = c->v_qkv; a.q_stride = 3 * D; a.k0 = c->v_qkv + D; a.v0 = c->v_qkv + 2 * D; a.kv0_stride = 3 * D; a.n0 = NT;
    a.out = c->v_attn; a.o_stride = D; a.B = B; a.H = g.v_heads; a.Sq = NT; a.HD = 64; a.scale = vscale; a.causal = 0;
    count(c); if (attention_prefill(a, st)) return -1;
    if (gemm_f32(c, c->v_attn, rows, D, D, L.wo, D, D, L.bo, 1, c->v_hidden, D, st)) return -1;
    count(c); if (layernorm(c->v_hidden, rows, D, L.ln2_w, L.ln2_b, g.v_eps, c->v_norm, nullptr, st)) return -1;
    if (gemm_bf16(c, c->v_norm, rows, D, D, L.w1, g.v_ffn, D, L.b1, ACT_QUICK_GELU, c->v_ffn, g.v_ffn, st)) return -1;
    if (gemm_f32(c, c->v_ffn, rows, g.v_ffn, g.v_ffn, L.w2, D, g.v_ffn, L.b2, 1, c->v_hidden, D, st)) return -1;
  }
  // post_layernorm on ALL tokens (what the reference does, modeling_visualcla.py:284/350)
  count(c); if (layernorm(c->v_hidden, rows, D, c->post_w, c->post_b, g.v_eps, c->v_norm, c->v_postln_f32, st)) return -1;

  // ---- Resampler
  const int R = g.r_hidden, Q = g.r_queries, RL = g.r_layers, qrows = B * Q;
  count(c); if (broadcast_rows(c->rq, Q, R, B, c->r_hidden, c->r_hidden_bf16, st)) return -1;
  // K/V of the (layer-invariant) image rows for all layers in one GEMM
  if (gemm_bf16(c, c->v_norm, rows, R, R, c->r_wkv_all, RL * 2 * R, R, c->r_bkv_all, ACT_NONE, c->r_kvimg, RL * 2 * R, st)) return -1;
  const float rscale = 1.0f / sqrtf(64.f);
  for (int i = 0; i < RL; ++i) {
    const ResamplerLayer& L = c->rl[i];
    if (gemm_bf16(c, c->r_hidden_bf16, qrows, R, R, L.wqkv, 3 * R, R, L.bqkv, ACT_NONE, c->r_qkv, 3 * R, st)) return -1;
    AttnCall a; a.q = c->r_qkv; a.q_stride = 3 * R;
    a.k0 = c->r_qkv + R; a.v0 = c->r_qkv + 2 * R; a.kv0_stride = 3 * R; a.n0 = Q;          // the query rows themselves (:315 cat)
    a.k1 = c->r_kvimg + (size_t)i * 2 * R; a.v1 = a.k1 + R; a.kv1_stride = RL * 2 * R; a.n1 = NT;   // image rows
    a.out = c->r_ctx; a.o_stride = R; a.B = B; a.H = g.r_heads; a.Sq = Q; a.HD = 64; a.scale = rscale; a.causal = 0;
    count(c); if (attention_prefill(a, st)) return -1;
    if (gemm_f32(c, c->r_ctx, qrows, R, R, L.wo, R, R, L.bo, 1, c->r_hidden, R, st)) return -1;           // dense + residual
    count(c); if (layernorm(c->r_hidden, qrows, R, L.ln1_w, L.ln1_b, g.r_eps, c->r_hidden_bf16, c->r_hidden, st)) return -1;  // post-LN
    if (gemm_bf16(c, c->r_hidden_bf16, qrows, R, R, L.wi, g.r_ffn, R, L.bi, ACT_GELU_ERF, c->r_ffn, g.r_ffn, st)) return -1;
    if (gemm_f32(c, c->r_ffn, qrows, g.r_ffn, g.r_ffn, L.wo2, R, g.r_ffn, L.bo2, 1, c->r_hidden, R, st)) return -1;
    count(c); if (layernorm(c->r_hidden, qrows, R, L.ln2_w, L.ln2_b, g.r_eps, c->r_hidden_bf16, c->r_hidden, st)) return -1;
  }
  // projector -> fp32 image embeddings
  if (gemm_f32(c, c->r_hidden_bf16, qrows, R, R, c->proj_w, g.t_hidden, R, c->proj_b, 0, c->img_embeds, g.t_hidden, st)) return -1;
  if (out_dev) VCLA_CUDA_OK(cudaMemcpyAsync(out_dev, c->img_embeds, (size_t)qrows * g.t_hidden * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}

// -------------------------------------------------------------------------------------------------
// prefill
// -------------------------------------------------------------------------------------------------
static int swap_gemm(vcla_ctx* c, const bf16* W, int n_out, int K, const bf16* X, int B, int splits, float* ws, cudaStream_t st, const GemmFix* fix = nullptr) {
  GemmCall g; g.A = W; g.B = X; g.M = n_out; g.N = B; g.K = K; g.lda = K; g.ldb = K; g.mode = GEMM_PARTIAL_F32; g.out = ws; g.ldo = n_out;
  g.splits = splits; g.ws_rows = B; g.weights_are_A = 1; g.l2_prefetch_kb = c->l2_prefetch_kb;
  if (fix) g.fix = *fix;
  count(c); return gemm_tc(g, st);
}

// ---- data parallel token exchange -------------------------------------------------------------------------------
// One all-gather of this rank's `dp_width` token slots + append to the global history.  fork != 0: on the side stream, joined by
// the next dp_wait() (the exchange is off the step's critical path: only the caller, not the next step, consumes it).
static int dp_gather(vcla_ctx* c, cudaStream_t st, int fork) {
  cudaStream_t gs = st;
  if (fork) {
    VCLA_CUDA_OK(cudaEventRecord(c->dp_fork, st));
    VCLA_CUDA_OK(cudaStreamWaitEvent(c->dp_stream, c->dp_fork, 0));
    gs = c->dp_stream;
  }
  VCLA_NCCL_OK(g_nccl.AllGather(c->dp_send, c->dp_recv, (size_t)c->dp_width, ncclInt32, c->comm, gs));
  count(c); if (dp_unpack(c->dp_recv, c->dp_world * c->dp_width, c->dp_hist, c->dp_step, gs)) return -1;
  if (fork) { VCLA_CUDA_OK(cudaEventRecord(c->dp_join, gs)); c->dp_pending = true; }
  return 0;
}
static int dp_wait(vcla_ctx* c, cudaStream_t st) {
  if (!c->dp_pending) return 0;
  VCLA_CUDA_OK(cudaStreamWaitEvent(st, c->dp_join, 0));
  c->dp_pending = false;
  return 0;
}

// logits reduce + argmax (+ token exchange when data parallel)
static int logits_argmax(vcla_ctx* c, int B, float* logits, int32_t* tok, const float* rstd, int fork, cudaStream_t st, int lm_splits = 0) {
  const vcla_config& g = c->cfg;
  const int sp_lm = lm_splits > 0 ? lm_splits : c->sp_lm;     // 1: ws_lm already holds the reduced logits (cluster split-K lm_head)
  if (c->dp_on() && dp_wait(c, st)) return -1;            // the previous step's exchange must have read dp_send before it is rewritten
  if (c->samp_on) {
    // logits -> [repetition penalty, no-repeat-ngram, temperature, top-k, top-p, draw] in one kernel; raw logits stay available
    float* lg = logits ? logits : c->samp_logits;
    count(c, 2);
    if (dec_logits_reduce(c->ws_lm, sp_lm, B, g.t_vocab, B, g.t_vocab, lg, g.t_vocab, rstd, c->cand_val, c->cand_idx, st)) return -1;
    if (dec_sample(lg, g.t_vocab, g.t_vocab, B, c->tok_hist, c->step_idx, c->samp_params, tok, c->tok_hist, c->dp_on() ? c->dp_send : nullptr, c->finished,
                   nullptr, st)) return -1;
    if (c->dp_on()) return dp_gather(c, st, fork);
    return 0;
  }
  count(c, 2);
  if (dec_logits_argmax(c->ws_lm, sp_lm, B, g.t_vocab, B, g.t_vocab, logits, g.t_vocab, tok, c->tok_hist, c->step_idx, rstd, c->cand_val, c->cand_idx,
                        c->dp_on() ? c->dp_send : nullptr, st)) return -1;
  if (c->dp_on()) return dp_gather(c, st, fork);
  return 0;
}

static int lm_head_last(vcla_ctx* c, int B, float* logits_dev, int32_t* tok_dev, cudaStream_t st) {
  // d_resid[B, T] holds the hidden state of the positions to score
  const vcla_config& g = c->cfg;
  count(c); if (dec_resid_norm(nullptr, 0, B, c->d_resid, B, g.t_hidden, c->final_norm, g.t_eps, c->d_xn, st)) return -1;
  if (swap_gemm(c, c->lm_head, g.t_vocab, g.t_hidden, c->d_xn, B, c->sp_lm, c->ws_lm, st)) return -1;
  return logits_argmax(c, B, logits_dev, tok_dev ? tok_dev : c->d_tok, nullptr, 0, st);
}

int vcla_prefill(vcla_ctx* c, const int64_t* ids, int B, int T, int image_mode, const int32_t* img_row, const int32_t* left_pad,
                 int pos_from_mask, float* logits_all, float* last_logits, int32_t* next_tok, vcla_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const vcla_config& g = c->cfg;
  const int nq = g.r_queries, TH = g.t_hidden, F = g.t_ffn, H = g.t_heads;
  const int S = (image_mode == VCLA_IMAGE_AT_HEAD) ? T + nq : T;
  if (B < 1 || B > g.max_batch || B > 64) { set_error("prefill: batch %d exceeds capacity (max_batch %d, <= 64 per call)", B, g.max_batch); return -1; }
  if ((long)B * S > g.max_prefill_tokens) { set_error("prefill: %d x %d tokens exceed max_prefill_tokens %d", B, S, g.max_prefill_tokens); return -1; }
  if (S > g.max_seq) { set_error("prefill: sequence %d exceeds max_seq %d", S, g.max_seq); return -1; }
  if (image_mode == VCLA_IMAGE_AT_HEAD && T < 2) { set_error("prefill: image_at_head needs >= 2 text tokens"); return -1; }
  if (image_mode == VCLA_IMAGE_AT_HEAD && left_pad != nullptr) { set_error("prefill: left padding is not defined for the image_at_head layout"); return -1; }
  const int rows = B * S;
  if (vcla_reset(c, stream)) return -1;
  // pages for the prompt's tokens (real tokens only: left padding is never cached)
  count(c); if (kv_reserve(c->kv_free, c->kv_state, c->kv_npages, c->page_table, c->pages_per_seq, c->page_tokens, B, S, left_pad, st)) return -1;
  count(c); if (embed_tokens(ids, B, T, S, TH, c->embed, g.t_vocab, image_mode == VCLA_IMAGE_AT_HEAD ? 1 : 0, nq, c->resid, st)) return -1;
  if (image_mode != VCLA_TEXT_ONLY) {
    const int32_t* rs = (image_mode == VCLA_IMAGE_AT_HEAD || img_row == nullptr) ? c->img_row_default : img_row;
    count(c); if (scatter_image_rows(c->img_embeds, B, nq, TH, rs, S, c->resid, st)) return -1;
  }
  const float scale = 1.0f / sqrtf(128.f);
  if (c->prefill_fused) {
    // 5 kernels per layer: RMSNorm is deferred (operand = bf16(resid * norm_w); the row scale commutes with the GEMM and is applied in
    // the consuming GEMM's epilogue from the per-tile sums of squares the producing GEMM wrote), RoPE + KV-cache append run in the
    // QKV GEMM's epilogue on the fp32 accumulator, SwiGLU in the gate/up epilogue, the residual add in the O / down epilogues.
    const int slots = (TH + gemm_pick_bn(rows, TH) - 1) / gemm_pick_bn(rows, TH);
    GemmRowScale rsc; rsc.ssq = c->p_ssq; rsc.slots = slots; rsc.inv_dim = 1.0f / (float)TH; rsc.eps = g.t_eps;
    count(c); if (prenorm_rows(c->resid, rows, TH, c->tl[0].ln1, c->xn, c->p_ssq, slots, st)) return -1;
    for (int i = 0; i < g.t_layers; ++i) {
      const TextLayer& L = c->tl[i];
      {
        GemmCall gc; gc.A = c->xn; gc.B = L.wqkv; gc.M = rows; gc.N = 3 * TH; gc.K = TH; gc.lda = TH; gc.ldb = TH; gc.mode = GEMM_STORE_BF16; gc.out = c->qkv; gc.ldo = 3 * TH;
        gc.rowscale = rsc;
        gc.rope.cos = c->rope_cos; gc.rope.sin = c->rope_sin; gc.rope.kv_pages = L.kv; gc.rope.page_table = c->page_table; gc.rope.pages_per_seq = c->pages_per_seq;
        gc.rope.page_tokens = c->page_tokens; gc.rope.S = S; gc.rope.T = TH; gc.rope.H = H; gc.rope.left_pad = left_pad; gc.rope.pos_from_mask = pos_from_mask;
        count(c); if (gemm_tc(gc, st)) return -1;
      }
      AttnCall a; a.q = c->qkv; a.q_stride = 3 * TH; a.k0 = c->qkv + TH; a.v0 = c->qkv + 2 * TH; a.kv0_stride = 3 * TH; a.n0 = S;
      a.out = c->attn; a.o_stride = TH; a.B = B; a.H = H; a.Sq = S; a.HD = 128; a.scale = scale; a.causal = 1; a.kv_start = left_pad;
      count(c); if (attention_prefill(a, st)) return -1;
      {
        GemmCall gc; gc.A = c->attn; gc.B = L.wo; gc.M = rows; gc.N = TH; gc.K = TH; gc.lda = TH; gc.ldb = TH; gc.mode = GEMM_ADD_F32; gc.accumulate = 1; gc.out = c->resid; gc.ldo = TH;
        gc.emit.norm_w = L.ln2; gc.emit.xw = c->xn; gc.emit.ldxw = TH; gc.emit.ssq_out = c->p_ssq;
        count(c); if (gemm_tc(gc, st)) return -1;
      }
      {
        GemmCall gc; gc.A = c->xn; gc.B = L.wgu; gc.M = rows; gc.N = 2 * F; gc.K = TH; gc.lda = TH; gc.ldb = TH; gc.mode = GEMM_SWIGLU_BF16; gc.out = c->hmid; gc.ldo = F;
        gc.rowscale = rsc;
        count(c); if (gemm_tc(gc, st)) return -1;
      }
      {
        GemmCall gc; gc.A = c->hmid; gc.B = L.wd; gc.M = rows; gc.N = TH; gc.K = F; gc.lda = F; gc.ldb = F; gc.mode = GEMM_ADD_F32; gc.accumulate = 1; gc.out = c->resid; gc.ldo = TH;
        gc.emit.norm_w = (i + 1 < g.t_layers) ? c->tl[i + 1].ln1 : c->final_norm; gc.emit.xw = c->xn; gc.emit.ldxw = TH; gc.emit.ssq_out = c->p_ssq;
        count(c); if (gemm_tc(gc, st)) return -1;
      }
    }
  } else
  for (int i = 0; i < g.t_layers; ++i) {
    const TextLayer& L = c->tl[i];
    count(c); if (rmsnorm(c->resid, rows, TH, L.ln1, g.t_eps, c->xn, st)) return -1;
    if (gemm_bf16(c, c->xn, rows, TH, TH, L.wqkv, 3 * TH, TH, nullptr, ACT_NONE, c->qkv, 3 * TH, st)) return -1;
    count(c); if (rope_and_cache(c->qkv, B, S, H, 128, c->rope_cos, c->rope_sin, L.kv, c->page_table, c->pages_per_seq, c->page_tokens, left_pad, pos_from_mask, st)) return -1;
    AttnCall a; a.q = c->qkv; a.q_stride = 3 * TH; a.k0 = c->qkv + TH; a.v0 = c->qkv + 2 * TH; a.kv0_stride = 3 * TH; a.n0 = S;
    a.out = c->attn; a.o_stride = TH; a.B = B; a.H = H; a.Sq = S; a.HD = 128; a.scale = scale; a.causal = 1; a.kv_start = left_pad;
    count(c); if (attention_prefill(a, st)) return -1;
    if (gemm_f32(c, c->attn, rows, TH, TH, L.wo, TH, TH, nullptr, 1, c->resid, TH, st)) return -1;
    count(c); if (rmsnorm(c->resid, rows, TH, L.ln2, g.t_eps, c->xn, st)) return -1;
    {
      GemmCall gc; gc.A = c->xn; gc.B = L.wgu; gc.M = rows; gc.N = 2 * F; gc.K = TH; gc.lda = TH; gc.ldb = TH; gc.mode = GEMM_SWIGLU_BF16; gc.out = c->hmid; gc.ldo = F;
      count(c); if (gemm_tc(gc, st)) return -1;
    }
    if (gemm_f32(c, c->hmid, rows, F, F, L.wd, TH, F, nullptr, 1, c->resid, TH, st)) return -1;
  }
  if (logits_all) {
    count(c); if (rmsnorm(c->resid, rows, TH, c->final_norm, g.t_eps, c->xn, st)) return -1;
    if (gemm_f32(c, c->xn, rows, TH, TH, c->lm_head, g.t_vocab, TH, nullptr, 0, logits_all, g.t_vocab, st)) return -1;
  }
  count(c); if (gather_last_rows(c->resid, B, S, TH, c->d_resid, st)) return -1;
  if (lm_head_last(c, B, last_logits, next_tok, st)) return -1;
  // sequence lengths become S - pad; the page the first decoded token will be appended to is reserved here
  count(c); if (advance_seq(c->seq_len, B, S, left_pad, c->step_idx, c->kv_free, c->kv_state, c->kv_npages, c->page_table, c->pages_per_seq, c->page_tokens, st)) return -1;
  c->len_bound = S;
  return 0;
}

static int advance_and_reserve(vcla_ctx* c, int B, cudaStream_t st) {
  return advance_seq(c->seq_len, B, 1, nullptr, c->step_idx, c->kv_free, c->kv_state, c->kv_npages, c->page_table, c->pages_per_seq, c->page_tokens, st);
}

// -------------------------------------------------------------------------------------------------
// decode
// -------------------------------------------------------------------------------------------------
// Legacy schedule (8 kernels per layer: separate norm / silu consumers).  Kept for A/B measurements (VCLA_FUSED_DECODE=0).
static int decode_enqueue_unfused(vcla_ctx* c, const int32_t* tok_in, int B, float* logits, int32_t* tok_out, cudaStream_t st) {
  const vcla_config& g = c->cfg;
  const int TH = g.t_hidden, F = g.t_ffn, H = g.t_heads;
  count(c); if (embed_tokens_i32(tok_in, B, TH, c->embed, g.t_vocab, c->d_resid, st)) return -1;
  const float scale = 1.0f / sqrtf(128.f);
  for (int i = 0; i < g.t_layers; ++i) {
    const TextLayer& L = c->tl[i];
    count(c); if (dec_resid_norm(i == 0 ? nullptr : c->ws_d, c->sp_d, B, c->d_resid, B, TH, L.ln1, g.t_eps, c->d_xn, st)) return -1;
    if (swap_gemm(c, L.wqkv, 3 * TH, TH, c->d_xn, B, c->sp_qkv, c->ws_qkv, st)) return -1;
    DecodeAttnCall a; a.qkv_partial = c->ws_qkv; a.splits = c->sp_qkv; a.ws_rows = B; a.kv_pages = L.kv; a.page_table = c->page_table;
    a.pages_per_seq = c->pages_per_seq; a.page_tokens = c->page_tokens; a.seq_len = c->seq_len; a.out = c->d_attn; a.scratch = c->attn_scratch;
    a.counters = c->attn_counters; a.B = B; a.H = H; a.HD = 128; a.scale = scale; a.rope_theta = g.rope_theta;
    a.rope_cos = c->rope_cos; a.rope_sin = c->rope_sin; a.persistent_mode = c->attn_persistent_mode; a.persistent_grid = c->attn_persistent_grid;
    { int want = (num_sms() + B * H - 1) / (B * H); int ks = want > c->kv_splits ? want : c->kv_splits; a.kv_splits = ks > 8 ? 8 : ks; }
    count(c); if (attention_decode(a, st)) return -1;
    if (swap_gemm(c, L.wo, TH, TH, c->d_attn, B, c->sp_o, c->ws_o, st)) return -1;
    count(c); if (dec_resid_norm(c->ws_o, c->sp_o, B, c->d_resid, B, TH, L.ln2, g.t_eps, c->d_xn, st)) return -1;
    if (swap_gemm(c, L.wgu, 2 * F, TH, c->d_xn, B, c->sp_gu, c->ws_gu, st)) return -1;
    count(c); if (dec_silu_mul(c->ws_gu, c->sp_gu, B, B, F, c->d_h, st)) return -1;
    if (swap_gemm(c, L.wd, TH, F, c->d_h, B, c->sp_d, c->ws_d, st)) return -1;
  }
  count(c); if (dec_resid_norm(c->ws_d, c->sp_d, B, c->d_resid, B, TH, c->final_norm, g.t_eps, c->d_xn, st)) return -1;
  if (swap_gemm(c, c->lm_head, g.t_vocab, TH, c->d_xn, B, c->sp_lm, c->ws_lm, st)) return -1;
  if (logits_argmax(c, B, logits, tok_out, nullptr, 1, st)) return -1;
  count(c); if (advance_and_reserve(c, B, st)) return -1;
  return 0;
}

// ---- cluster split-K schedule (default for batch <= 32): 5 kernels per layer, no split-K workspace, no consumer kernels ----------
// CTAs per cluster for a [M, K] weight at batch B: the choice that keeps the largest share of the 2 x SMs CTA slots busy over whole
// rounds of cluster-tiles (clusters are gang-scheduled: floor(slots / S) of them are resident).
static int csk_pick(int M, int K, int B) {
  const int tiles = (M + 127) / 128, kb = (K + 63) / 64, bn = B <= 16 ? 16 : 32;
  int best = 1; double best_score = -1.0;
  for (int S = 1; S <= 8; ++S) {
    const int per = (kb + S - 1) / S;
    if ((kb + per - 1) / per != S) continue;                 // every K slice non-empty
    if (S > 1 && kb / S < 2) break;
    if (((B + S - 1) / S) * S > bn + 4) continue;            // reduce buffer columns
    int ncl = gemm_csk_clusters(B, S);
    if (ncl <= 0) continue;
    if (ncl > tiles) ncl = tiles;
    const int rounds = (tiles + ncl - 1) / ncl;
    double score = (double)tiles * S / ((double)rounds * 2.0 * num_sms());
    if (rounds >= 2) score += 0.02;                          // a second tile per CTA overlaps its loads with the first one's epilogue
    if (score > best_score) { best_score = score; best = S; }
  }
  return best;
}
static int csk_prepare(vcla_ctx* c, int B) {
  if (c->csk_batch == B) return 0;
  const vcla_config& g = c->cfg;
  int v[5] = {csk_pick(3 * g.t_hidden, g.t_hidden, B), csk_pick(g.t_hidden, g.t_hidden, B), csk_pick(2 * g.t_ffn, g.t_hidden, B),
              csk_pick(g.t_hidden, g.t_ffn, B), csk_pick(g.t_vocab, g.t_hidden, B)};
  if (const char* e = getenv("VCLA_CSK_SPLITS")) {            // tuning override: "qkv,o,gu,d,lm"
    int o[5];
    if (sscanf(e, "%d,%d,%d,%d,%d", &o[0], &o[1], &o[2], &o[3], &o[4]) == 5) for (int i = 0; i < 5; ++i) if (o[i] >= 1 && o[i] <= 8) v[i] = o[i];
  }
  c->csk_qkv = v[0]; c->csk_o = v[1]; c->csk_gu = v[2]; c->csk_d = v[3]; c->csk_lm = v[4];
  c->csk_batch = B;
  return 0;
}

static int decode_enqueue_csk(vcla_ctx* c, const int32_t* tok_in, int B, float* logits, int32_t* tok_out, cudaStream_t st) {
  const vcla_config& g = c->cfg;
  const int TH = g.t_hidden, F = g.t_ffn, H = g.t_heads;
  const int slots = (TH + 127) / 128;                            // per-row sum-of-squares slots = 128-row tiles of the o / down projections
  const float inv_dim = 1.0f / (float)TH;
  count(c); if (dec_embed(tok_in, B, TH, c->embed, g.t_vocab, c->d_resid, c->tl[0].ln1, g.t_eps, c->d_xn, nullptr, c->d_ssq, slots, st)) return -1;
  const float scale = 1.0f / sqrtf(128.f);
  auto base = [&](const bf16* W, const bf16* X, int M, int K, int splits, int mode) {
    CskCall k; k.W = W; k.X = X; k.M = M; k.B = B; k.K = K; k.splits = splits; k.mode = mode; k.inv_dim = inv_dim; k.eps = g.t_eps;
    return k;
  };
  for (int i = 0; i < g.t_layers; ++i) {
    const TextLayer& L = c->tl[i];
    { CskCall k = base(L.wqkv, c->d_xn, 3 * TH, TH, c->csk_qkv, CSK_OUT_F32); k.out = c->ws_qkv; k.ldo = 3 * TH; k.ssq_in = c->d_ssq; k.ssq_slots = slots;
      count(c); if (gemm_csk(k, st)) return -1; }
    DecodeAttnCall a; a.qkv_partial = c->ws_qkv; a.splits = 1; a.ws_rows = B; a.kv_pages = L.kv; a.page_table = c->page_table;
    a.pages_per_seq = c->pages_per_seq; a.page_tokens = c->page_tokens; a.seq_len = c->seq_len; a.out = c->d_attn; a.scratch = c->attn_scratch;
    a.counters = c->attn_counters; a.B = B; a.H = H; a.HD = 128; a.scale = scale; a.rope_theta = g.rope_theta;
    a.rope_cos = c->rope_cos; a.rope_sin = c->rope_sin; a.persistent_mode = c->attn_persistent_mode; a.persistent_grid = c->attn_persistent_grid;
    { int want = (num_sms() + B * H - 1) / (B * H); int ks = want > c->kv_splits ? want : c->kv_splits; a.kv_splits = ks > 8 ? 8 : ks; }
    count(c); if (attention_decode(a, st)) return -1;
    { CskCall k = base(L.wo, c->d_attn, TH, TH, c->csk_o, CSK_RESID); k.resid = c->d_resid; k.norm_w = L.ln2; k.xw = c->d_xn; k.ssq_out = c->d_ssq;
      count(c); if (gemm_csk(k, st)) return -1; }
    { CskCall k = base(L.wgu, c->d_xn, 2 * F, TH, c->csk_gu, CSK_SWIGLU); k.h = c->d_h; k.ssq_in = c->d_ssq; k.ssq_slots = slots;
      count(c); if (gemm_csk(k, st)) return -1; }
    { CskCall k = base(L.wd, c->d_h, TH, F, c->csk_d, CSK_RESID); k.resid = c->d_resid; k.norm_w = (i + 1 < g.t_layers) ? c->tl[i + 1].ln1 : c->final_norm;
      k.xw = c->d_xn; k.ssq_out = c->d_ssq;
      count(c); if (gemm_csk(k, st)) return -1; }
  }
  { CskCall k = base(c->lm_head, c->d_xn, g.t_vocab, TH, c->csk_lm, CSK_OUT_F32); k.out = c->ws_lm; k.ldo = g.t_vocab; k.ssq_in = c->d_ssq; k.ssq_slots = slots;
    count(c); if (gemm_csk(k, st)) return -1; }
  if (logits_argmax(c, B, logits, tok_out, nullptr, 1, st, 1)) return -1;
  count(c); if (advance_and_reserve(c, B, st)) return -1;
  return 0;
}

// Decode step, 5 kernels per layer:  QKV GEMM -> attention(+reduce, rstd, RoPE, append) -> O GEMM [+residual, norm weight, sum sq]
//   -> gate/up GEMM [+rstd, SiLU*mul] -> down GEMM [+residual, next norm weight, sum sq].  The bracketed consumers run inside
// the GEMM, in the CTA whose split-K partial completes a tile; RMSNorm's per-row scale is deferred to the next consumer.
static int decode_enqueue(vcla_ctx* c, const int32_t* tok_in, int B, float* logits, int32_t* tok_out, cudaStream_t st) {
  if (c->decode_schedule == 2 && B <= 32) return decode_enqueue_csk(c, tok_in, B, logits, tok_out, st);
  if (!c->fused_decode) return decode_enqueue_unfused(c, tok_in, B, logits, tok_out, st);
  const vcla_config& g = c->cfg;
  const int TH = g.t_hidden, F = g.t_ffn, H = g.t_heads;
  count(c); if (dec_embed(tok_in, B, TH, c->embed, g.t_vocab, c->d_resid, c->tl[0].ln1, g.t_eps, c->d_xn, c->d_rstd, nullptr, 0, st)) return -1;
  const float scale = 1.0f / sqrtf(128.f);
  GemmFix fr;   // residual + deferred norm
  fr.mode = FIX_RESID; fr.resid = c->d_resid; fr.xw_out = c->d_xn; fr.ssq = c->d_ssq; fr.rstd_out = c->d_rstd; fr.inv_dim = 1.0f / (float)TH; fr.eps = g.t_eps;
  GemmFix fs;   // SwiGLU
  fs.mode = FIX_SWIGLU; fs.tile_counters = c->cnt_gu; fs.rstd_in = c->d_rstd; fs.h_out = c->d_h;
  for (int i = 0; i < g.t_layers; ++i) {
    const TextLayer& L = c->tl[i];
    if (swap_gemm(c, L.wqkv, 3 * TH, TH, c->d_xn, B, c->sp_qkv, c->ws_qkv, st)) return -1;
    DecodeAttnCall a; a.qkv_partial = c->ws_qkv; a.splits = c->sp_qkv; a.ws_rows = B; a.kv_pages = L.kv; a.page_table = c->page_table;
    a.pages_per_seq = c->pages_per_seq; a.page_tokens = c->page_tokens; a.seq_len = c->seq_len; a.out = c->d_attn; a.scratch = c->attn_scratch;
    a.counters = c->attn_counters; a.B = B; a.H = H; a.HD = 128; a.scale = scale; a.rope_theta = g.rope_theta; a.rstd = c->d_rstd;
    a.rope_cos = c->rope_cos; a.rope_sin = c->rope_sin; a.persistent_mode = c->attn_persistent_mode; a.persistent_grid = c->attn_persistent_grid;
    // enough CTAs to cover the SMs for small batches; long contexts split so a CTA streams <= ~12 pages
    { int want = (num_sms() + B * H - 1) / (B * H); int ks = want > c->kv_splits ? want : c->kv_splits; a.kv_splits = ks > 8 ? 8 : ks; }
    count(c); if (attention_decode(a, st)) return -1;
    fr.tile_counters = c->cnt_o; fr.norm_w = L.ln2;
    if (swap_gemm(c, L.wo, TH, TH, c->d_attn, B, c->sp_o, c->ws_o, st, &fr)) return -1;
    if (swap_gemm(c, L.wgu, 2 * F, TH, c->d_xn, B, c->sp_gu, c->ws_gu, st, &fs)) return -1;
    fr.tile_counters = c->cnt_d; fr.norm_w = (i + 1 < g.t_layers) ? c->tl[i + 1].ln1 : c->final_norm;
    if (swap_gemm(c, L.wd, TH, F, c->d_h, B, c->sp_d, c->ws_d, st, &fr)) return -1;
  }
  if (swap_gemm(c, c->lm_head, g.t_vocab, TH, c->d_xn, B, c->sp_lm, c->ws_lm, st)) return -1;
  if (logits_argmax(c, B, logits, tok_out, c->d_rstd, 1, st)) return -1;
  count(c); if (advance_and_reserve(c, B, st)) return -1;
  return 0;
}

static int decode_graph(vcla_ctx* c, const int32_t* tok_in, int B, float* logits, int32_t* tok_out, int n_steps, cudaStream_t st) {
  GraphKey key{B, tok_in, logits, tok_out, n_steps, c->dp_on() ? 1 : 0, c->samp_on ? 1 : 0};
  auto it = c->graphs.find(key);
  if (it == c->graphs.end()) {
    const int64_t before = c->launches;
    cudaGraph_t graph = nullptr;
    if (!c->cap_stream) VCLA_CUDA_OK(cudaStreamCreateWithFlags(&c->cap_stream, cudaStreamNonBlocking));
    VCLA_CUDA_OK(cudaStreamBeginCapture(c->cap_stream, cudaStreamCaptureModeThreadLocal));
    int rc = 0;
    for (int i = 0; i < n_steps && rc == 0; ++i) rc = decode_enqueue(c, tok_in, B, logits, tok_out, c->cap_stream);
    if (rc == 0 && c->dp_on()) rc = dp_wait(c, c->cap_stream);      // a captured graph must join its forked exchange branch
    cudaError_t e = cudaStreamEndCapture(c->cap_stream, &graph);
    if (rc != 0) { if (graph) cudaGraphDestroy(graph); (void)cudaGetLastError(); return -1; }
    if (e != cudaSuccess) { set_error("decode: graph capture failed: %s", cudaGetErrorString(e)); return -1; }
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { set_error("decode: graph instantiate failed: %s", cudaGetErrorString(e)); return -1; }
    c->graph_launches[key] = c->launches - before;
    c->launches = before;  // capture enqueued nothing
    // bounded cache: a caller that keeps passing fresh buffers evicts the least recently used graph instead of growing forever.
    // (Evicting while an older launch of that graph may still be in flight needs the device to be idle first.)
    constexpr size_t kMaxGraphs = 24;
    if (c->graphs.size() >= kMaxGraphs) {
      const GraphKey victim = c->graph_lru.back();
      c->graph_lru.pop_back();
      VCLA_CUDA_OK(cudaDeviceSynchronize());
      cudaGraphExecDestroy(c->graphs[victim]);
      c->graphs.erase(victim);
      c->graph_launches.erase(victim);
    }
    c->graphs[key] = exec;
    c->graph_lru.push_front(key);
    it = c->graphs.find(key);
  } else {
    for (auto li = c->graph_lru.begin(); li != c->graph_lru.end(); ++li) {
      if (!(*li < key) && !(key < *li)) { c->graph_lru.erase(li); break; }
    }
    c->graph_lru.push_front(key);
  }
  VCLA_CUDA_OK(cudaGraphLaunch(it->second, st));
  c->launches += c->graph_launches[key];
  return 0;
}

// Every decode step appends one token per sequence: refuse the call instead of running past the context capacity (the kernels
// index the page table, the RoPE table and the token history by the sequence length).
static int decode_capacity(vcla_ctx* c, int n_steps) {
  if (c->len_bound <= 0) { set_error("decode: no prefilled sequences (call vcla_prefill first)"); return -1; }
  if (c->len_bound + n_steps > c->cfg.max_seq) {
    set_error("decode: %lld cached tokens + %d steps exceed the context capacity max_seq=%d", (long long)c->len_bound, n_steps, c->cfg.max_seq);
    return -1;
  }
  return 0;
}

int vcla_decode_step(vcla_ctx* c, const int32_t* tok_in, int B, float* logits, int32_t* tok_out, int use_graph, vcla_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (B < 1 || B > c->cfg.max_batch || B > 64) { set_error("decode: batch %d unsupported", B); return -1; }
  if (!tok_in || !tok_out) { set_error("decode: null token buffers"); return -1; }
  if (decode_capacity(c, 1)) return -1;
  if (c->decode_schedule == 2 && B <= 32 && csk_prepare(c, B)) return -1;      // occupancy queries: never inside a capture
  int rc = use_graph ? decode_graph(c, tok_in, B, logits, tok_out, 1, st) : decode_enqueue(c, tok_in, B, logits, tok_out, st);
  if (rc == 0 && !use_graph && c->dp_on()) rc = dp_wait(c, st);
  if (rc == 0) c->len_bound += 1;
  return rc;
}

int vcla_decode_multi(vcla_ctx* c, int32_t* tok_inout, int B, int n_steps, vcla_stream stream) {
  // n_steps greedy decode steps captured back to back in ONE CUDA graph (the token buffer is consumed and rewritten in place,
  // every chosen token is appended to the device-side history): amortises the gap between consecutive graph launches.
  if (B < 1 || B > c->cfg.max_batch || B > 64) { set_error("decode: batch %d unsupported", B); return -1; }
  if (!tok_inout || n_steps < 1 || n_steps > 64) { set_error("decode_multi: bad arguments"); return -1; }
  if (decode_capacity(c, n_steps)) return -1;
  if (c->decode_schedule == 2 && B <= 32 && csk_prepare(c, B)) return -1;
  const int rc = decode_graph(c, tok_inout, B, nullptr, tok_inout, n_steps, (cudaStream_t)stream);
  if (rc == 0) c->len_bound += n_steps;
  return rc;
}

// -------------------------------------------------------------------------------------------------
// device-side sampling
// -------------------------------------------------------------------------------------------------
static int sampler_to_params(const vcla_sampler* s, SamplerParams* p) {
  if (s->n_eos < 0 || s->n_eos > 4) { set_error("sampler: at most 4 eos ids"); return -1; }
  if (s->do_sample && (s->top_k < 1 || s->top_k > 1024)) { set_error("sampler: the device path needs 1 <= top_k <= 1024 (got %d)", s->top_k); return -1; }
  if (s->do_sample && !(s->temperature > 0.f)) { set_error("sampler: temperature must be > 0"); return -1; }
  if (!(s->repetition_penalty > 0.f)) { set_error("sampler: repetition_penalty must be > 0"); return -1; }
  if (s->top_p <= 0.f || s->top_p > 1.f) { set_error("sampler: top_p must be in (0, 1]"); return -1; }
  memset(p, 0, sizeof(*p));
  p->do_sample = s->do_sample ? 1 : 0;
  p->rep_penalty = s->repetition_penalty; p->no_repeat_ngram = s->no_repeat_ngram_size > 0 ? s->no_repeat_ngram_size : 0;
  p->temperature = s->temperature; p->top_k = s->top_k; p->top_p = s->top_p;
  p->one_minus_top_p = (float)(1.0 - (double)s->top_p);
  p->min_new_tokens = s->min_new_tokens; p->n_eos = s->n_eos; p->pad_id = s->pad_token_id;
  for (int i = 0; i < s->n_eos; ++i) p->eos[i] = s->eos_token_id[i];
  p->seed = s->seed;
  return 0;
}

int vcla_sampler_supported(const vcla_ctx* c) { return c ? sampler_supported(c->cfg.t_vocab) : 0; }

int vcla_set_sampler(vcla_ctx* c, const vcla_sampler* s, vcla_stream stream) {
  if (!c) return -1;
  if (s == nullptr) { c->samp_on = false; return 0; }
  if (!sampler_supported(c->cfg.t_vocab)) { set_error("sampler: vocabulary %d does not fit the device sampler", c->cfg.t_vocab); return -1; }
  SamplerParams p;
  if (sampler_to_params(s, &p)) return -1;
  // pageable source: the driver stages it before returning, so `p` may go out of scope
  VCLA_CUDA_OK(cudaMemcpyAsync(c->samp_params, &p, sizeof(p), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  c->samp_on = true;
  return 0;
}

int vcla_read_finished(vcla_ctx* c, int32_t* dst_dev, int B, vcla_stream stream) {
  if (!c || !dst_dev || B < 1 || B > 64) { set_error("vcla_read_finished: bad arguments"); return -1; }
  VCLA_CUDA_OK(cudaMemcpyAsync(dst_dev, c->finished, (size_t)B * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

int vcla_op_sample(const float* logits_dev, int B, int V, const int32_t* history_dev, int L, const vcla_sampler* s, int32_t* tok_dev,
                   float* scores_out_dev, vcla_stream stream) {
  if (!logits_dev || !s || B < 1 || V < 1 || L < 0 || (L > 0 && !history_dev)) { set_error("vcla_op_sample: bad arguments"); return -1; }
  SamplerParams p;
  if (sampler_to_params(s, &p) || sampler_init()) return -1;
  uint8_t* scratch = nullptr;
  VCLA_CUDA_OK(cudaMalloc(&scratch, sizeof(SamplerParams) + 16));
  cudaStream_t st = (cudaStream_t)stream;
  VCLA_CUDA_OK(cudaMemcpyAsync(scratch, &p, sizeof(p), cudaMemcpyHostToDevice, st));
  VCLA_CUDA_OK(cudaMemcpyAsync(scratch + sizeof(SamplerParams), &L, 4, cudaMemcpyHostToDevice, st));
  const int rc = dec_sample(logits_dev, V, V, B, history_dev, (const int32_t*)(scratch + sizeof(SamplerParams)), (const SamplerParams*)scratch, tok_dev, nullptr,
                            nullptr, nullptr, scores_out_dev, st);
  cudaStreamSynchronize(st);
  cudaFree(scratch);
  return rc;
}

// -------------------------------------------------------------------------------------------------
// data parallel: NCCL communicator owned by the context, token all-gather inside the decode graph
// -------------------------------------------------------------------------------------------------
int vcla_nccl_unique_id(uint8_t* out128) {
  if (!out128) { set_error("vcla_nccl_unique_id: null"); return -1; }
  if (nccl_api()) return -1;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  VCLA_NCCL_OK(g_nccl.GetUniqueId(&id));
  memcpy(out128, &id, 128);
  return 0;
}

int vcla_nccl_init(vcla_ctx* c, const uint8_t* id128, int rank, int world, int width) {
  if (!c || !id128 || world < 1 || rank < 0 || rank >= world || width < 1 || width > 64) { set_error("vcla_nccl_init: bad arguments"); return -1; }
  if (c->comm) { set_error("vcla_nccl_init: context already has a communicator"); return -1; }
  if (nccl_api()) return -1;
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  VCLA_NCCL_OK(g_nccl.CommInitRank(&c->comm, world, id, rank));
  c->dp_rank = rank; c->dp_world = world; c->dp_width = width;
  const size_t n_send = 64, n_recv = (size_t)world * 64, n_hist = (size_t)(c->cfg.max_seq + 2) * world * width;
  VCLA_CUDA_OK(cudaMalloc(&c->dp_send, (n_send + n_recv + n_hist + 4) * 4));
  VCLA_CUDA_OK(cudaMemset(c->dp_send, 0, (n_send + n_recv + n_hist + 4) * 4));
  c->dp_recv = c->dp_send + n_send; c->dp_hist = c->dp_recv + n_recv; c->dp_step = c->dp_hist + n_hist;
  VCLA_CUDA_OK(cudaStreamCreateWithFlags(&c->dp_stream, cudaStreamNonBlocking));
  VCLA_CUDA_OK(cudaEventCreateWithFlags(&c->dp_fork, cudaEventDisableTiming));
  VCLA_CUDA_OK(cudaEventCreateWithFlags(&c->dp_join, cudaEventDisableTiming));
  // graphs captured before the communicator existed do not contain the exchange
  VCLA_CUDA_OK(cudaDeviceSynchronize());
  for (auto& kv : c->graphs) cudaGraphExecDestroy(kv.second);
  c->graphs.clear(); c->graph_launches.clear(); c->graph_lru.clear();
  return 0;
}

int vcla_allgather_tokens(vcla_ctx* c, const int32_t* local_dev, int n, int32_t* all_dev, vcla_stream stream) {
  if (!c || !c->comm) { set_error("vcla_allgather_tokens: call vcla_nccl_init first"); return -1; }
  if (!local_dev || !all_dev || n < 1) { set_error("vcla_allgather_tokens: bad arguments"); return -1; }
  VCLA_NCCL_OK(g_nccl.AllGather(local_dev, all_dev, (size_t)n, ncclInt32, c->comm, (cudaStream_t)stream));
  return 0;
}

int vcla_dp_set_active(vcla_ctx* c, int on) {
  // the exchange is part of prefill / decode only while active (every rank of the communicator must then make the same calls);
  // a rank-local generate() on a context that owns a communicator runs with it off
  if (!c) return -1;
  if (on && !c->comm) { set_error("vcla_dp_set_active: call vcla_nccl_init first"); return -1; }
  c->dp_active = on != 0;
  return 0;
}

int vcla_dp_exchange(vcla_ctx* c, vcla_stream stream) {
  // the exchange of one step without any compute: for a rank that holds no requests (global batch < world size)
  if (!c || !c->comm) { set_error("vcla_dp_exchange: call vcla_nccl_init first"); return -1; }
  return dp_gather(c, (cudaStream_t)stream, 0);
}

int vcla_read_history_dp(vcla_ctx* c, int32_t* dst_dev, int n_steps, vcla_stream stream) {
  // [n_steps][world * width] int32: the tokens every rank chose at the prefill (row 0) and each decode step since
  if (!c || !c->comm) { set_error("vcla_read_history_dp: call vcla_nccl_init first"); return -1; }
  if (n_steps < 0 || n_steps > c->cfg.max_seq + 1) { set_error("vcla_read_history_dp: bad arguments"); return -1; }
  VCLA_CUDA_OK(cudaMemcpyAsync(dst_dev, c->dp_hist, (size_t)n_steps * c->dp_world * c->dp_width * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

// -------------------------------------------------------------------------------------------------
// introspection + operator-level entry points
// -------------------------------------------------------------------------------------------------
int vcla_read_stage(vcla_ctx* c, const char* stage, int B, float* dst, vcla_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const vcla_config& g = c->cfg;
  const float* src = nullptr; size_t n = 0;
  if (!strcmp(stage, "vit_out")) { src = c->v_hidden; n = (size_t)B * c->v_tokens * g.v_hidden; }
  else if (!strcmp(stage, "post_ln")) { src = c->v_postln_f32; n = (size_t)B * c->v_tokens * g.v_hidden; }
  else if (!strcmp(stage, "resampler_out")) { src = c->r_hidden; n = (size_t)B * g.r_queries * g.r_hidden; }
  else if (!strcmp(stage, "projector_out")) { src = c->img_embeds; n = (size_t)B * g.r_queries * g.t_hidden; }
  else { set_error("vcla_read_stage: unknown stage '%s'", stage); return -1; }
  VCLA_CUDA_OK(cudaMemcpyAsync(dst, src, n * 4, cudaMemcpyDeviceToHost, st));
  VCLA_CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

int vcla_bench_decode_gemm(vcla_ctx* c, int which, int B, int reps, float* avg_us, int64_t* weight_bytes, vcla_stream stream) {
  // Times one decode weight-streaming GEMM shape over all layers' (distinct) weights, reps times, with CUDA events on
  // the launching stream.  13 GB of weights >> 126 MB L2, so every launch streams from HBM.
  cudaStream_t st = (cudaStream_t)stream;
  const vcla_config& g = c->cfg;
  if (B < 1 || B > 64 || which < 0 || which > 4 || reps < 1) { set_error("bench_decode_gemm: bad arguments"); return -1; }
  const int TH = g.t_hidden, F = g.t_ffn;
  cudaEvent_t e0, e1;
  VCLA_CUDA_OK(cudaEventCreate(&e0));
  VCLA_CUDA_OK(cudaEventCreate(&e1));
  const bool csk = c->decode_schedule == 2 && B <= 32;
  if (csk && csk_prepare(c, B)) return -1;
  const int slots = (TH + 127) / 128;
  auto csk_one = [&](int w, const TextLayer* L) -> int {
    CskCall k; k.B = B; k.inv_dim = 1.0f / (float)TH; k.eps = g.t_eps;
    if (w == 0) { k.W = L->wqkv; k.X = c->d_xn; k.M = 3 * TH; k.K = TH; k.splits = c->csk_qkv; k.mode = CSK_OUT_F32; k.out = c->ws_qkv; k.ldo = 3 * TH; k.ssq_in = c->d_ssq; k.ssq_slots = slots; }
    if (w == 1) { k.W = L->wo; k.X = c->d_attn; k.M = TH; k.K = TH; k.splits = c->csk_o; k.mode = CSK_RESID; k.resid = c->d_resid; k.norm_w = L->ln2; k.xw = c->d_xn; k.ssq_out = c->d_ssq; }
    if (w == 2) { k.W = L->wgu; k.X = c->d_xn; k.M = 2 * F; k.K = TH; k.splits = c->csk_gu; k.mode = CSK_SWIGLU; k.h = c->d_h; k.ssq_in = c->d_ssq; k.ssq_slots = slots; }
    if (w == 3) { k.W = L->wd; k.X = c->d_h; k.M = TH; k.K = F; k.splits = c->csk_d; k.mode = CSK_RESID; k.resid = c->d_resid; k.norm_w = L->ln1; k.xw = c->d_xn; k.ssq_out = c->d_ssq; }
    if (w == 4) { k.W = c->lm_head; k.X = c->d_xn; k.M = g.t_vocab; k.K = TH; k.splits = c->csk_lm; k.mode = CSK_OUT_F32; k.out = c->ws_lm; k.ldo = g.t_vocab; k.ssq_in = c->d_ssq; k.ssq_slots = slots; }
    count(c); return gemm_csk(k, st);
  };
  auto run_all = [&]() -> int {
    if (csk) {
      if (which == 4) return csk_one(4, nullptr);
      for (int i = 0; i < g.t_layers; ++i) if (csk_one(which, &c->tl[i])) return -1;
      return 0;
    }
    if (which == 4) return swap_gemm(c, c->lm_head, g.t_vocab, TH, c->d_xn, B, c->sp_lm, c->ws_lm, st);
    for (int i = 0; i < g.t_layers; ++i) {
      const TextLayer& L = c->tl[i];
      int rc = 0;
      if (which == 0) rc = swap_gemm(c, L.wqkv, 3 * TH, TH, c->d_xn, B, c->sp_qkv, c->ws_qkv, st);
      if (which == 1) rc = swap_gemm(c, L.wo, TH, TH, c->d_attn, B, c->sp_o, c->ws_o, st);
      if (which == 2) rc = swap_gemm(c, L.wgu, 2 * F, TH, c->d_xn, B, c->sp_gu, c->ws_gu, st);
      if (which == 3) rc = swap_gemm(c, L.wd, TH, F, c->d_h, B, c->sp_d, c->ws_d, st);
      if (rc) return rc;
    }
    return 0;
  };
  if (run_all()) return -1;  // warm-up
  VCLA_CUDA_OK(cudaEventRecord(e0, st));
  for (int r = 0; r < reps; ++r) if (run_all()) return -1;
  VCLA_CUDA_OK(cudaEventRecord(e1, st));
  VCLA_CUDA_OK(cudaEventSynchronize(e1));
  float ms = 0.f;
  VCLA_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
  const int per = which == 4 ? 1 : g.t_layers;
  if (avg_us) *avg_us = ms * 1000.f / (float)(reps * per);
  if (weight_bytes) {
    const int64_t n[5] = {(int64_t)3 * TH * TH, (int64_t)TH * TH, (int64_t)2 * F * TH, (int64_t)TH * F, (int64_t)g.t_vocab * TH};
    *weight_bytes = n[which] * 2;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return 0;
}

int vcla_read_history(vcla_ctx* c, int32_t* dst_dev, int B, int n_steps, vcla_stream stream) {
  // tokens chosen by the prefill (step 0) and every decode step since, as a device [n_steps, B] int32 array
  if (n_steps < 0 || n_steps > c->cfg.max_seq + 1 || B < 1 || B > 64) { set_error("vcla_read_history: bad arguments"); return -1; }
  VCLA_CUDA_OK(cudaMemcpyAsync(dst_dev, c->tok_hist, (size_t)n_steps * B * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

int vcla_trace_enable(vcla_ctx* c, int max_events) {
  // installs (max_events > 0) or removes (0) the timeline buffer every kernel's CTA 0 appends to
  VCLA_CUDA_OK(cudaDeviceSynchronize());
  if (c->trace_buf) { trace_set_gemm(nullptr, 0); trace_set_attention_tc(nullptr, 0); trace_set_gemm_decode(nullptr, 0); trace_set_sampler(nullptr, 0); trace_set_attention(nullptr, 0); trace_set_elementwise(nullptr, 0); cudaFree(c->trace_buf); c->trace_buf = nullptr; c->trace_cap = 0; }
  if (max_events <= 0) return 0;
  const size_t bytes = 8 + (size_t)max_events * 32;
  VCLA_CUDA_OK(cudaMalloc(&c->trace_buf, bytes));
  VCLA_CUDA_OK(cudaMemset(c->trace_buf, 0, bytes));
  c->trace_cap = (unsigned long long)max_events;
  if (trace_set_gemm(c->trace_buf, c->trace_cap) || trace_set_attention_tc(c->trace_buf, c->trace_cap) || trace_set_gemm_decode(c->trace_buf, c->trace_cap) || trace_set_sampler(c->trace_buf, c->trace_cap) || trace_set_attention(c->trace_buf, c->trace_cap) || trace_set_elementwise(c->trace_buf, c->trace_cap)) {
    set_error("vcla_trace_enable: cudaMemcpyToSymbol failed");
    return -1;
  }
  return 0;
}
int vcla_trace_read(vcla_ctx* c, uint64_t* dst_host, int max_events, int* n_events) {
  // copies [tag, t_entry_ns, t_dep_ns, t_exit_ns] x n to the host and clears the buffer.  Synchronises.
  if (!c->trace_buf) { set_error("vcla_trace_read: tracing is off"); return -1; }
  VCLA_CUDA_OK(cudaDeviceSynchronize());
  unsigned long long cnt = 0;
  VCLA_CUDA_OK(cudaMemcpy(&cnt, c->trace_buf, 8, cudaMemcpyDeviceToHost));
  if (cnt > c->trace_cap) cnt = c->trace_cap;
  if (cnt > (unsigned long long)max_events) cnt = (unsigned long long)max_events;
  VCLA_CUDA_OK(cudaMemcpy(dst_host, (uint8_t*)c->trace_buf + 8, cnt * 32, cudaMemcpyDeviceToHost));
  VCLA_CUDA_OK(cudaMemset(c->trace_buf, 0, 8));
  if (n_events) *n_events = (int)cnt;
  return 0;
}

int vcla_op_gemm(const void* A, const void* W, int M, int N, int K, int mode, int act, int accumulate, const float* bias, void* out,
                 int ldo, int splits, int tile_n, int use_reference, vcla_stream stream) {
  GemmCall g;
  g.A = (const bf16*)A; g.B = (const bf16*)W; g.M = M; g.N = N; g.K = K; g.lda = K; g.ldb = K; g.mode = mode; g.act = act; g.accumulate = accumulate;
  g.bias = bias; g.out = out; g.ldo = ldo; g.splits = splits; g.ws_rows = N; g.bn = tile_n; g.weights_are_A = (mode == GEMM_PARTIAL_F32);
  return use_reference ? gemm_naive(g, (cudaStream_t)stream) : gemm_tc(g, (cudaStream_t)stream);
}
int vcla_op_gemm_csk(const void* W, const void* X, int M, int B, int K, int splits, int mode, float* out_or_resid, const float* norm_w, void* xw_or_h,
                     float* ssq_out, const float* ssq_in, int ssq_slots, float inv_dim, float eps, vcla_stream stream) {
  CskCall k; k.W = (const bf16*)W; k.X = (const bf16*)X; k.M = M; k.B = B; k.K = K; k.splits = splits; k.mode = mode;
  k.ssq_in = ssq_in; k.ssq_slots = ssq_slots; k.inv_dim = inv_dim; k.eps = eps;
  if (mode == CSK_OUT_F32) { k.out = out_or_resid; k.ldo = M; }
  else if (mode == CSK_RESID) { k.resid = out_or_resid; k.norm_w = norm_w; k.xw = (bf16*)xw_or_h; k.ssq_out = ssq_out; }
  else if (mode == CSK_SWIGLU) { k.h = (bf16*)xw_or_h; }
  else { set_error("vcla_op_gemm_csk: unknown mode %d", mode); return -1; }
  return gemm_csk(k, (cudaStream_t)stream);
}
int vcla_debug_set_csk_splits(vcla_ctx* c, int B, int qkv, int o, int gu, int d, int lm) {
  // tuning hook: CTAs per cluster of the five decode GEMM shapes at batch B (0 = keep the automatic choice); drops the captured graphs
  if (!c || B < 1 || B > 32) { set_error("vcla_debug_set_csk_splits: bad arguments"); return -1; }
  c->csk_batch = 0;
  if (csk_prepare(c, B)) return -1;
  int* dst[5] = {&c->csk_qkv, &c->csk_o, &c->csk_gu, &c->csk_d, &c->csk_lm};
  const int v[5] = {qkv, o, gu, d, lm};
  for (int i = 0; i < 5; ++i) if (v[i] >= 1 && v[i] <= 8) *dst[i] = v[i];
  VCLA_CUDA_OK(cudaDeviceSynchronize());
  for (auto& kv : c->graphs) cudaGraphExecDestroy(kv.second);
  c->graphs.clear(); c->graph_launches.clear(); c->graph_lru.clear();
  return 0;
}
int vcla_debug_get_csk_splits(vcla_ctx* c, int B, int* out5) {
  if (!c || !out5 || B < 1 || B > 32) return -1;
  if (c->csk_batch != B && csk_prepare(c, B)) return -1;
  out5[0] = c->csk_qkv; out5[1] = c->csk_o; out5[2] = c->csk_gu; out5[3] = c->csk_d; out5[4] = c->csk_lm;
  return 0;
}
int vcla_op_gemm_csk_clusters(int B, int splits) { return gemm_csk_clusters(B, splits); }
void vcla_set_gemm_two_cta(int on) { gemm_set_two_cta(on); }
void vcla_set_attention_tc(int on) { attention_set_tc(on); }
int vcla_op_attention(const void* q, int q_stride, const void* k0, const void* v0, int kv0_stride, int n0, const void* k1, const void* v1,
                      int kv1_stride, int n1, void* out, int o_stride, int B, int H, int Sq, int HD, float scale, int causal, vcla_stream stream) {
  AttnCall a; a.q = (const bf16*)q; a.q_stride = q_stride; a.k0 = (const bf16*)k0; a.v0 = (const bf16*)v0; a.kv0_stride = kv0_stride; a.n0 = n0;
  a.k1 = (const bf16*)k1; a.v1 = (const bf16*)v1; a.kv1_stride = kv1_stride; a.n1 = n1; a.out = (bf16*)out; a.o_stride = o_stride;
  a.B = B; a.H = H; a.Sq = Sq; a.HD = HD; a.scale = scale; a.causal = causal;
  return attention_prefill(a, (cudaStream_t)stream);
}
int vcla_op_layernorm(const float* x, int rows, int D, const float* w, const float* b, float eps, void* y_bf16, float* y_f32, vcla_stream stream) {
  return layernorm(x, rows, D, w, b, eps, (bf16*)y_bf16, y_f32, (cudaStream_t)stream);
}
int vcla_op_rmsnorm(const float* x, int rows, int D, const float* w, float eps, void* y_bf16, vcla_stream stream) {
  return rmsnorm(x, rows, D, w, eps, (bf16*)y_bf16, (cudaStream_t)stream);
}

}  // extern "C"
