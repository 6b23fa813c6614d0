// HBM-bound kernels of the VisualCLA path: normalisation, embedding / splice, RoPE + KV-cache append, the decode
// consumers of the split-K partial sums, argmax, synthetic-weight generation and checkpoint repacking.
// All are simple streaming kernels: 128-bit coalesced accesses where layouts allow, fp32 statistics.
#include "common.cuh"
#include "kernels.h"

#include <cooperative_groups.h>
#include <math.h>
#include <vector>

namespace vcla {

static inline cudaLaunchConfig_t make_cfg(dim3 grid, dim3 block, size_t smem, cudaStream_t st, cudaLaunchAttribute* attr) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  int na = 0;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  return cfg;
}
#define VCLA_LAUNCH(kernel, grid, block, smem, st, ...)                         \
  do {                                                                          \
    cudaLaunchAttribute _attr[1];                                               \
    cudaLaunchConfig_t _cfg = make_cfg(grid, block, smem, st, _attr);           \
    VCLA_CUDA_OK(cudaLaunchKernelEx(&_cfg, kernel, __VA_ARGS__));               \
  } while (0)

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  t = warp_sum(t);
  return t;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm / RMSNorm: one CTA per row, row cached in smem, two-pass fp32 statistics
// ------------------------------------------------------------------------------------------------
__global__ void layernorm_kernel(const float* x, int D, const float* __restrict__ w, const float* __restrict__ b,
                                 float eps, bf16* __restrict__ y_bf16, float* y_f32) {
  extern __shared__ float rowbuf[];
  __shared__ float red[32];
  TraceScope trace(5);
  pdl_launch_dependents();   // dependents may become resident early; they block in their own griddepcontrol.wait
  pdl_wait();
  trace.dep();
  const size_t row = blockIdx.x;
  const float* xr = x + row * D;
  float s = 0.f;
  for (int i = threadIdx.x * 4; i < D; i += blockDim.x * 4) {
    float4 v = *reinterpret_cast<const float4*>(xr + i);
    *reinterpret_cast<float4*>(rowbuf + i) = v;
    s += v.x + v.y + v.z + v.w;
  }
  const float mean = block_sum(s, red) / D;
  float q = 0.f;
  for (int i = threadIdx.x * 4; i < D; i += blockDim.x * 4) {
    float4 v = *reinterpret_cast<float4*>(rowbuf + i);
    float a = v.x - mean, bb = v.y - mean, cc = v.z - mean, dd = v.w - mean;
    q += a * a + bb * bb + cc * cc + dd * dd;
  }
  const float rstd = rsqrtf(block_sum(q, red) / D + eps);
  for (int i = threadIdx.x * 4; i < D; i += blockDim.x * 4) {
    float4 v = *reinterpret_cast<float4*>(rowbuf + i);
    float4 wv = *reinterpret_cast<const float4*>(w + i), bv = *reinterpret_cast<const float4*>(b + i);
    float o0 = (v.x - mean) * rstd * wv.x + bv.x, o1 = (v.y - mean) * rstd * wv.y + bv.y;
    float o2 = (v.z - mean) * rstd * wv.z + bv.z, o3 = (v.w - mean) * rstd * wv.w + bv.w;
    if (y_bf16) *reinterpret_cast<uint2*>(y_bf16 + row * D + i) = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
    if (y_f32) *reinterpret_cast<float4*>(y_f32 + row * D + i) = make_float4(o0, o1, o2, o3);
  }
}
int layernorm(const float* x, int rows, int D, const float* w, const float* b, float eps, bf16* y_bf16, float* y_f32, cudaStream_t st) {
  if (D % 4) { set_error("layernorm: D %% 4 != 0"); return -1; }
  if (rows == 0) return 0;
  VCLA_LAUNCH(layernorm_kernel, dim3(rows), dim3(D >= 2048 ? 256 : 128), (size_t)D * 4, st, x, D, w, b, eps, y_bf16, y_f32);
  return 0;
}

__global__ void rmsnorm_kernel(const float* __restrict__ x, int D, const float* __restrict__ w, float eps, bf16* __restrict__ y) {
  extern __shared__ float rowbuf[];
  __shared__ float red[32];
  TraceScope trace(6);
  pdl_launch_dependents();   // dependents may become resident early; they block in their own griddepcontrol.wait
  pdl_wait();
  trace.dep();
  const size_t row = blockIdx.x;
  const float* xr = x + row * D;
  float q = 0.f;
  for (int i = threadIdx.x * 4; i < D; i += blockDim.x * 4) {
    float4 v = *reinterpret_cast<const float4*>(xr + i);
    *reinterpret_cast<float4*>(rowbuf + i) = v;
    q += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  const float rstd = rsqrtf(block_sum(q, red) / D + eps);
  for (int i = threadIdx.x * 4; i < D; i += blockDim.x * 4) {
    float4 v = *reinterpret_cast<float4*>(rowbuf + i);
    float4 wv = *reinterpret_cast<const float4*>(w + i);
    *reinterpret_cast<uint2*>(y + row * D + i) =
        make_uint2(pack_bf16x2(v.x * rstd * wv.x, v.y * rstd * wv.y), pack_bf16x2(v.z * rstd * wv.z, v.w * rstd * wv.w));
  }
}
int rmsnorm(const float* x, int rows, int D, const float* w, float eps, bf16* y, cudaStream_t st) {
  if (D % 4) { set_error("rmsnorm: D %% 4 != 0"); return -1; }
  if (rows == 0) return 0;
  VCLA_LAUNCH(rmsnorm_kernel, dim3(rows), dim3(D >= 2048 ? 256 : 128), (size_t)D * 4, st, x, D, w, eps, y);
  return 0;
}

__global__ void prenorm_rows_kernel(const float* __restrict__ x, int D, const float* __restrict__ w, bf16* __restrict__ xw, float* __restrict__ ssq, int slots) {
  __shared__ float red[32];
  TraceScope trace(6);
  pdl_launch_dependents();
  pdl_wait();
  trace.dep();
  const size_t row = blockIdx.x;
  const float* xr = x + row * D;
  float q = 0.f;
  for (int i = threadIdx.x * 4; i < D; i += blockDim.x * 4) {
    const float4 v = *reinterpret_cast<const float4*>(xr + i);
    const float4 wv = *reinterpret_cast<const float4*>(w + i);
    q += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    *reinterpret_cast<uint2*>(xw + row * D + i) = make_uint2(pack_bf16x2(v.x * wv.x, v.y * wv.y), pack_bf16x2(v.z * wv.z, v.w * wv.w));
  }
  q = block_sum(q, red);
  for (int sidx = threadIdx.x; sidx < slots; sidx += blockDim.x) ssq[row * slots + sidx] = sidx == 0 ? q : 0.f;
}
int prenorm_rows(const float* x, int rows, int D, const float* w, bf16* xw, float* ssq, int slots, cudaStream_t st) {
  if (D % 4) { set_error("prenorm_rows: D %% 4 != 0"); return -1; }
  if (rows == 0) return 0;
  VCLA_LAUNCH(prenorm_rows_kernel, dim3(rows), dim3(D >= 2048 ? 256 : 128), 0, st, x, D, w, xw, ssq, slots);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// ViT front end
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }

template <typename T>
__global__ void im2col_kernel(const T* __restrict__ px, int B, int image, int patch, int kpad, bf16* __restrict__ out) {
  // one CTA per patch row; k = c*P*P + ky*P + kx  (Conv2d weight (D,3,P,P) flattened)
  const int g = image / patch;
  const int prow = blockIdx.x;  // b*g*g + py*g + px
  const int b = prow / (g * g), pp = prow % (g * g), py = pp / g, pxi = pp % g;
  const int kreal = 3 * patch * patch;
  for (int k = threadIdx.x; k < kpad; k += blockDim.x) {
    float v = 0.f;
    if (k < kreal) {
      int c = k / (patch * patch), r = k % (patch * patch), ky = r / patch, kx = r % patch;
      v = to_f32<T>(px[(((size_t)b * 3 + c) * image + (py * patch + ky)) * image + (pxi * patch + kx)]);
    }
    out[(size_t)prow * kpad + k] = __float2bfloat16(v);
  }
}
int im2col(const void* pixels, int dtype, int B, int image, int patch, int kpad, bf16* out, cudaStream_t st) {
  const int g = image / patch;
  dim3 grid(B * g * g), block(128);
  if (dtype == 0) im2col_kernel<float><<<grid, block, 0, st>>>((const float*)pixels, B, image, patch, kpad, out);
  else if (dtype == 1) im2col_kernel<__half><<<grid, block, 0, st>>>((const __half*)pixels, B, image, patch, kpad, out);
  else if (dtype == 2) im2col_kernel<bf16><<<grid, block, 0, st>>>((const bf16*)pixels, B, image, patch, kpad, out);
  else { set_error("im2col: unknown pixel dtype %d", dtype); return -1; }
  VCLA_CUDA_OK(cudaGetLastError());
  return 0;
}

__global__ void vit_cls_rows_kernel(float* hidden, int tokens, int D, const float* cls, const float* pos) {
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < D; i += blockDim.x) hidden[(size_t)b * tokens * D + i] = cls[i] + pos[i];
}
int vit_cls_rows(float* hidden, int B, int tokens, int D, const float* cls, const float* pos, cudaStream_t st) {
  vit_cls_rows_kernel<<<B, 256, 0, st>>>(hidden, tokens, D, cls, pos);
  VCLA_CUDA_OK(cudaGetLastError());
  return 0;
}

__global__ void broadcast_rows_kernel(const float* src, int rows, int D, float* dst_f32, bf16* dst_bf16) {
  const int r = blockIdx.x, b = blockIdx.y;
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    float v = src[(size_t)r * D + i];
    size_t o = ((size_t)b * rows + r) * D + i;
    if (dst_f32) dst_f32[o] = v;
    if (dst_bf16) dst_bf16[o] = __float2bfloat16(v);
  }
}
int broadcast_rows(const float* src, int rows, int D, int B, float* dst_f32, bf16* dst_bf16, cudaStream_t st) {
  broadcast_rows_kernel<<<dim3(rows, B), 256, 0, st>>>(src, rows, D, dst_f32, dst_bf16);
  VCLA_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// embedding gather / image splice
// ------------------------------------------------------------------------------------------------
__global__ void embed_tokens_kernel(const int64_t* __restrict__ ids, int T, int S, int D, const bf16* __restrict__ table,
                                    int vocab, int mode, int nq, float* __restrict__ dst) {
  const int t = blockIdx.x, b = blockIdx.y;
  long long id = ids[(size_t)b * T + t];
  if (id < 0 || id >= vocab) id = 0;
  const int pos = (mode == 1 && t >= 2) ? t + nq : t;
  const bf16* src = table + (size_t)id * D;
  float* d = dst + ((size_t)b * S + pos) * D;
  for (int i = threadIdx.x * 8; i < D; i += blockDim.x * 8) {
    uint4 v = *reinterpret_cast<const uint4*>(src + i);
    float2 a = unpack_bf16x2(v.x), bb = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z), e = unpack_bf16x2(v.w);
    *reinterpret_cast<float4*>(d + i) = make_float4(a.x, a.y, bb.x, bb.y);
    *reinterpret_cast<float4*>(d + i + 4) = make_float4(c.x, c.y, e.x, e.y);
  }
}
int embed_tokens(const int64_t* ids, int B, int T, int S, int D, const bf16* table, int vocab, int mode, int nq, float* dst, cudaStream_t st) {
  if (D % 8) { set_error("embed: D %% 8 != 0"); return -1; }
  VCLA_LAUNCH(embed_tokens_kernel, dim3(T, B), dim3(128), 0, st, ids, T, S, D, table, vocab, mode, nq, dst);
  return 0;
}
__global__ void embed_tokens_i32_kernel(const int32_t* __restrict__ ids, int D, const bf16* __restrict__ table, int vocab, float* __restrict__ dst) {
  TraceScope trace(13);
  pdl_launch_dependents();   // dependents may become resident early; they block in their own griddepcontrol.wait
  pdl_wait();
  trace.dep();
  const int b = blockIdx.x;
  int id = ids[b];
  if (id < 0 || id >= vocab) id = 0;
  const bf16* src = table + (size_t)id * D;
  float* d = dst + (size_t)b * D;
  for (int i = threadIdx.x * 8; i < D; i += blockDim.x * 8) {
    uint4 v = *reinterpret_cast<const uint4*>(src + i);
    float2 a = unpack_bf16x2(v.x), bb = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z), e = unpack_bf16x2(v.w);
    *reinterpret_cast<float4*>(d + i) = make_float4(a.x, a.y, bb.x, bb.y);
    *reinterpret_cast<float4*>(d + i + 4) = make_float4(c.x, c.y, e.x, e.y);
  }
}
int embed_tokens_i32(const int32_t* ids, int B, int D, const bf16* table, int vocab, float* dst, cudaStream_t st) {
  VCLA_LAUNCH(embed_tokens_i32_kernel, dim3(B), dim3(256), 0, st, ids, D, table, vocab, dst);
  return 0;
}

// decode step entry (one CTA per sequence): embedding row -> fp32 residual, bf16 GEMM operand pre-multiplied by the first
// layer's RMSNorm weight, and the row's deferred scale 1/rms
__global__ void dec_embed_kernel(const int32_t* __restrict__ ids, int D, const bf16* __restrict__ table, int vocab, float* __restrict__ resid,
                                 const float* __restrict__ norm_w, float eps, bf16* __restrict__ xw, float* __restrict__ rstd, float* __restrict__ ssq, int slots) {
  __shared__ float red[32];
  TraceScope trace(13);
  pdl_launch_dependents();   // dependents may become resident early; they block in their own griddepcontrol.wait
  pdl_wait();
  trace.dep();
  const int b = blockIdx.x;
  int id = ids[b];
  if (id < 0 || id >= vocab) id = 0;
  const bf16* src = table + (size_t)id * D;
  float q = 0.f;
  for (int i = threadIdx.x * 8; i < D; i += blockDim.x * 8) {
    const uint4 v = *reinterpret_cast<const uint4*>(src + i);
    const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
    float f[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 t = unpack_bf16x2(w4[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
    *reinterpret_cast<float4*>(resid + (size_t)b * D + i) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(resid + (size_t)b * D + i + 4) = make_float4(f[4], f[5], f[6], f[7]);
    const float4 w0 = *reinterpret_cast<const float4*>(norm_w + i), w1 = *reinterpret_cast<const float4*>(norm_w + i + 4);
    *reinterpret_cast<uint4*>(xw + (size_t)b * D + i) = make_uint4(pack_bf16x2(f[0] * w0.x, f[1] * w0.y), pack_bf16x2(f[2] * w0.z, f[3] * w0.w),
                                                                    pack_bf16x2(f[4] * w1.x, f[5] * w1.y), pack_bf16x2(f[6] * w1.z, f[7] * w1.w));
#pragma unroll
    for (int j = 0; j < 8; ++j) q += f[j] * f[j];
  }
  q = block_sum(q, red);
  if (threadIdx.x == 0 && rstd != nullptr) rstd[b] = rsqrtf(q / D + eps);
  // deferred-norm chain head: the row's sum of squares in slot 0, the other slots empty
  if (ssq != nullptr) for (int i = threadIdx.x; i < slots; i += blockDim.x) ssq[(size_t)b * slots + i] = i == 0 ? q : 0.f;
}
int dec_embed(const int32_t* ids, int B, int D, const bf16* table, int vocab, float* resid, const float* norm_w, float eps, bf16* xw, float* rstd,
              float* ssq, int slots, cudaStream_t st) {
  if (D % 8) { set_error("dec_embed: D %% 8 != 0"); return -1; }
  VCLA_LAUNCH(dec_embed_kernel, dim3(B), dim3(256), 0, st, ids, D, table, vocab, resid, norm_w, eps, xw, rstd, ssq, slots);
  return 0;
}

__global__ void scatter_image_rows_kernel(const float* __restrict__ img, int nq, int D, const int32_t* __restrict__ row_start, int S, float* __restrict__ dst) {
  const int q = blockIdx.x, b = blockIdx.y;
  const int rs = row_start[b];
  if (rs < 0) return;  // this sample carries no image
  const float4* s = reinterpret_cast<const float4*>(img + ((size_t)b * nq + q) * D);
  float4* d = reinterpret_cast<float4*>(dst + ((size_t)b * S + rs + q) * D);
  for (int i = threadIdx.x; i < D / 4; i += blockDim.x) d[i] = s[i];
}
int scatter_image_rows(const float* img, int B, int nq, int D, const int32_t* row_start, int S, float* dst, cudaStream_t st) {
  scatter_image_rows_kernel<<<dim3(nq, B), 256, 0, st>>>(img, nq, D, row_start, S, dst);
  VCLA_CUDA_OK(cudaGetLastError());
  return 0;
}

__global__ void gather_last_rows_kernel(const float* hidden, int S, int D, float* dst) {
  const int b = blockIdx.x;
  const float4* s = reinterpret_cast<const float4*>(hidden + ((size_t)b * S + (S - 1)) * D);
  float4* d = reinterpret_cast<float4*>(dst + (size_t)b * D);
  for (int i = threadIdx.x; i < D / 4; i += blockDim.x) d[i] = s[i];
}
int gather_last_rows(const float* hidden, int B, int S, int D, float* dst, cudaStream_t st) {
  gather_last_rows_kernel<<<B, 256, 0, st>>>(hidden, S, D, dst);
  VCLA_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// RoPE tables (fp32, computed once on the host the way HF does: HF:models/llama/modeling_llama.py:98-141)
// ------------------------------------------------------------------------------------------------
int rope_fill_tables(int max_pos, int head_dim, float theta, float* cos_dev, float* sin_dev) {
  const int half = head_dim / 2;
  std::vector<float> hc((size_t)max_pos * half), hs((size_t)max_pos * half);
  for (int i = 0; i < half; ++i) {
    const float inv = 1.0f / powf(theta, (float)(2 * i) / (float)head_dim);
    for (int p = 0; p < max_pos; ++p) {
      const float f = (float)p * inv;
      hc[(size_t)p * half + i] = (float)cos((double)f);
      hs[(size_t)p * half + i] = (float)sin((double)f);
    }
  }
  VCLA_CUDA_OK(cudaMemcpy(cos_dev, hc.data(), hc.size() * 4, cudaMemcpyHostToDevice));
  VCLA_CUDA_OK(cudaMemcpy(sin_dev, hs.data(), hs.size() * 4, cudaMemcpyHostToDevice));
  return 0;
}

// prefill: rotate q,k in place inside the fused [B*S, 3T] projection buffer; append k,v to the paged cache.
// One CTA per token; each thread owns 8 consecutive rotary pairs of one head: 128-bit loads/stores throughout.
__global__ void rope_and_cache_kernel(bf16* __restrict__ qkv, int S, int H, int HD, const float* __restrict__ rc, const float* __restrict__ rs,
                                      bf16* __restrict__ kv_pages, const int32_t* __restrict__ page_table, int pages_per_seq,
                                      int page_tokens, const int32_t* __restrict__ left_pad, int pos_from_mask) {
  TraceScope trace(7);
  pdl_launch_dependents();   // dependents may become resident early; they block in their own griddepcontrol.wait
  pdl_wait();
  trace.dep();
  const int s = blockIdx.x, b = blockIdx.y;
  const int T = H * HD, half = HD / 2, groups = half / 8;
  const int pad = left_pad ? left_pad[b] : 0;
  if (s < pad) return;                                 // padding row: not rotated, not cached, masked in attention
  const int cpos = s - pad;                            // index inside the (compact) KV cache
  const int pos = pos_from_mask ? cpos : s;            // rotary position
  bf16* row = qkv + ((size_t)b * S + s) * 3 * T;
  const int page = page_table[(size_t)b * pages_per_seq + cpos / page_tokens];
  const int slot = cpos % page_tokens;
  for (int idx = threadIdx.x; idx < H * groups; idx += blockDim.x) {
    const int h = idx / groups, i = (idx % groups) * 8;
    float cs[8], sn[8];
    *reinterpret_cast<float4*>(cs) = *reinterpret_cast<const float4*>(rc + (size_t)pos * half + i);
    *reinterpret_cast<float4*>(cs + 4) = *reinterpret_cast<const float4*>(rc + (size_t)pos * half + i + 4);
    *reinterpret_cast<float4*>(sn) = *reinterpret_cast<const float4*>(rs + (size_t)pos * half + i);
    *reinterpret_cast<float4*>(sn + 4) = *reinterpret_cast<const float4*>(rs + (size_t)pos * half + i + 4);
    bf16* kd = kv_pages + ((((size_t)page * 2 + 0) * H + h) * page_tokens + slot) * HD;
    bf16* vd = kv_pages + ((((size_t)page * 2 + 1) * H + h) * page_tokens + slot) * HD;
#pragma unroll
    for (int which = 0; which < 2; ++which) {   // 0: q, 1: k
      bf16* x = row + which * T + h * HD;
      const uint4 lo = *reinterpret_cast<const uint4*>(x + i), hi = *reinterpret_cast<const uint4*>(x + i + half);
      const uint32_t lw[4] = {lo.x, lo.y, lo.z, lo.w}, hw[4] = {hi.x, hi.y, hi.z, hi.w};
      uint32_t ol[4], oh[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a = unpack_bf16x2(lw[j]), c2 = unpack_bf16x2(hw[j]);
        ol[j] = pack_bf16x2(a.x * cs[2 * j] - c2.x * sn[2 * j], a.y * cs[2 * j + 1] - c2.y * sn[2 * j + 1]);
        oh[j] = pack_bf16x2(c2.x * cs[2 * j] + a.x * sn[2 * j], c2.y * cs[2 * j + 1] + a.y * sn[2 * j + 1]);
      }
      const uint4 vlo = make_uint4(ol[0], ol[1], ol[2], ol[3]), vhi = make_uint4(oh[0], oh[1], oh[2], oh[3]);
      *reinterpret_cast<uint4*>(x + i) = vlo;
      *reinterpret_cast<uint4*>(x + i + half) = vhi;
      if (which == 1) {
        *reinterpret_cast<uint4*>(kd + i) = vlo;
        *reinterpret_cast<uint4*>(kd + i + half) = vhi;
      }
    }
    const bf16* v = row + 2 * T + h * HD;
    *reinterpret_cast<uint4*>(vd + i) = *reinterpret_cast<const uint4*>(v + i);
    *reinterpret_cast<uint4*>(vd + i + half) = *reinterpret_cast<const uint4*>(v + i + half);
  }
}
int rope_and_cache(bf16* qkv, int B, int S, int H, int HD, const float* rope_cos, const float* rope_sin, bf16* kv_pages,
                   const int32_t* page_table, int pages_per_seq, int page_tokens, const int32_t* left_pad, int pos_from_mask, cudaStream_t st) {
  if (!rope_cos || !rope_sin) { set_error("rope table not initialised"); return -1; }
  VCLA_LAUNCH(rope_and_cache_kernel, dim3(S, B), dim3(256), 0, st, qkv, S, H, HD, rope_cos, rope_sin,
              kv_pages, page_table, pages_per_seq, page_tokens, left_pad, pos_from_mask);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// decode consumers of split-K partial sums
// ------------------------------------------------------------------------------------------------
// One 8-CTA thread-block cluster per batch row: each CTA reduces the split-K partials of D/8 columns into the fp32
// residual stream, the row's sum of squares is exchanged through distributed shared memory, then every CTA writes its
// slice of the normalised bf16 GEMM operand.  (One CTA per row was latency-bound: 12 us per launch, 65 launches/step.)
constexpr int kNormCluster = 8;
__global__ void __cluster_dims__(kNormCluster, 1, 1) __launch_bounds__(128)
dec_resid_norm_kernel(const float* __restrict__ partial, int splits, int ws_rows, float* __restrict__ resid, int D,
                      const float* __restrict__ w, float eps, bf16* __restrict__ xn) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float red[32];
  __shared__ float s_part;
  TraceScope trace(8);
  pdl_launch_dependents();   // dependents may become resident early; they block in their own griddepcontrol.wait
  pdl_wait();
  trace.dep();
  const int b = blockIdx.y;
  const int chunk = D / kNormCluster;
  const int c0 = blockIdx.x * chunk;
  constexpr int MAXV = 4;
  float4 vals[MAXV];
  float q = 0.f;
  int nv = 0;
  for (int i = threadIdx.x * 4; i < chunk; i += blockDim.x * 4, ++nv) {
    const int col = c0 + i;
    float4 v = *reinterpret_cast<const float4*>(resid + (size_t)b * D + col);
    if (partial) {
      // (issuing all <= 24 partial loads at once was measured slower in situ: 6.96 vs 5.40 us after o_proj)
#pragma unroll 4
      for (int s = 0; s < splits; ++s) {
        const float4 p = __ldcg(reinterpret_cast<const float4*>(partial + ((size_t)s * ws_rows + b) * D + col));
        v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
      }
      *reinterpret_cast<float4*>(resid + (size_t)b * D + col) = v;
    }
    vals[nv] = v;
    q += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  q = block_sum(q, red);
  if (threadIdx.x == 0) s_part = q;
  cluster.sync();
  float tot = 0.f;
#pragma unroll
  for (int r = 0; r < kNormCluster; ++r) tot += *cluster.map_shared_rank(&s_part, r);
  cluster.sync();            // nobody may exit while a peer still reads its shared memory
  const float rstd = rsqrtf(tot / D + eps);
  nv = 0;
  for (int i = threadIdx.x * 4; i < chunk; i += blockDim.x * 4, ++nv) {
    const int col = c0 + i;
    const float4 v = vals[nv];
    const float4 wv = *reinterpret_cast<const float4*>(w + col);
    *reinterpret_cast<uint2*>(xn + (size_t)b * D + col) =
        make_uint2(pack_bf16x2(v.x * rstd * wv.x, v.y * rstd * wv.y), pack_bf16x2(v.z * rstd * wv.z, v.w * rstd * wv.w));
  }
}
int dec_resid_norm(const float* partial, int splits, int ws_rows, float* resid, int B, int D, const float* w, float eps, bf16* xn, cudaStream_t st) {
  if (D % (kNormCluster * 4) != 0 || D / kNormCluster > 128 * 4 * 4) { set_error("dec_resid_norm: unsupported D %d", D); return -1; }
  VCLA_LAUNCH(dec_resid_norm_kernel, dim3(kNormCluster, B), dim3(128), 0, st, partial, splits, ws_rows, resid, D, w, eps, xn);
  return 0;
}

__global__ void dec_silu_mul_kernel(const float* __restrict__ partial, int splits, int ws_rows, int F, bf16* __restrict__ h) {
  TraceScope trace(9);
  pdl_launch_dependents();   // dependents may become resident early; they block in their own griddepcontrol.wait
  pdl_wait();
  trace.dep();
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= F) return;
  const int gi = (j >> 5) * 64 + (j & 31);
  float g = 0.f, u = 0.f;
  for (int s = 0; s < splits; ++s) {
    const float* row = partial + ((size_t)s * ws_rows + b) * (size_t)(2 * F);
    g += row[gi];
    u += row[gi + 32];
  }
  h[(size_t)b * F + j] = __float2bfloat16(g / (1.f + __expf(-g)) * u);
}
int dec_silu_mul(const float* partial, int splits, int ws_rows, int B, int F, bf16* h, cudaStream_t st) {
  VCLA_LAUNCH(dec_silu_mul_kernel, dim3((F + 255) / 256, B), dim3(256), 0, st, partial, splits, ws_rows, F, h);
  return 0;
}

// logits + argmax: stage 1 per (vocab chunk, b) -> candidate ; stage 2 per b
constexpr int kArgChunks = kArgmaxChunks;
__global__ void dec_logits_stage1(const float* __restrict__ partial, int splits, int ws_rows, int ldp, int V, float* __restrict__ logits,
                                  int ld_logits, float* __restrict__ cand_val, int* __restrict__ cand_idx, const float* __restrict__ rstd) {
  __shared__ float sv[32];
  __shared__ int si[32];
  TraceScope trace(10);
  pdl_launch_dependents();   // dependents may become resident early; they block in their own griddepcontrol.wait
  pdl_wait();
  trace.dep();
  const int b = blockIdx.y, ch = blockIdx.x;
  const float rs = rstd ? __ldcg(rstd + b) : 1.f;       // deferred RMSNorm scale of the final norm
  const int per = (V + kArgChunks - 1) / kArgChunks;
  const int v0 = ch * per, v1 = min(V, v0 + per);
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
    float x = 0.f;
    for (int s = 0; s < splits; ++s) x += __ldcg(partial + ((size_t)s * ws_rows + b) * (size_t)ldp + v);
    x *= rs;
    if (logits) logits[(size_t)b * ld_logits + v] = x;
    if (x > best) { best = x; bi = v; }   // strided order: smaller index kept on ties via the reduction below
  }
  // warp + block reduce, ties -> smallest index
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sv[warp] = best; si[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    best = lane < nw ? sv[lane] : -INFINITY;
    bi = lane < nw ? si[lane] : 0x7fffffff;
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) { cand_val[b * kArgChunks + ch] = best; cand_idx[b * kArgChunks + ch] = bi; }
  }
}
__global__ void dec_logits_stage2(const float* __restrict__ cand_val, const int* __restrict__ cand_idx, int32_t* __restrict__ tok,
                                  int32_t* __restrict__ history, const int32_t* __restrict__ step_idx, int32_t* __restrict__ dp_send) {
  TraceScope trace(11);
  pdl_launch_dependents();   // dependents may become resident early; they block in their own griddepcontrol.wait
  pdl_wait();
  trace.dep();
  const int b = blockIdx.x, lane = threadIdx.x;
  float best = cand_val[b * kArgChunks + lane];
  int bi = cand_idx[b * kArgChunks + lane];
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if (lane == 0) {
    tok[b] = bi;
    if (history) history[(size_t)(*step_idx) * gridDim.x + b] = bi;   // [step][B] log of every chosen token since the prefill
    if (dp_send) dp_send[b] = bi;                                      // send buffer of the per-step token all-gather (data parallel)
  }
}
int dec_logits_argmax(const float* partial, int splits, int ws_rows, int ldp, int B, int V, float* logits, int ld_logits, int32_t* tok,
                      int32_t* history, const int32_t* step_idx, const float* rstd, float* cand_val, int32_t* cand_idx, int32_t* dp_send, cudaStream_t st) {
  if (!cand_val || !cand_idx) { set_error("argmax scratch missing"); return -1; }
  VCLA_LAUNCH(dec_logits_stage1, dim3(kArgChunks, B), dim3(256), 0, st, partial, splits, ws_rows, ldp, V, logits, ld_logits, cand_val, (int*)cand_idx, rstd);
  VCLA_LAUNCH(dec_logits_stage2, dim3(B), dim3(32), 0, st, (const float*)cand_val, (const int*)cand_idx, tok, history, step_idx, dp_send);
  return 0;
}

int dec_logits_reduce(const float* partial, int splits, int ws_rows, int ldp, int B, int V, float* logits, int ld_logits, const float* rstd,
                      float* cand_val, int32_t* cand_idx, cudaStream_t st) {
  if (!logits || !cand_val || !cand_idx) { set_error("dec_logits_reduce: missing buffers"); return -1; }
  VCLA_LAUNCH(dec_logits_stage1, dim3(kArgChunks, B), dim3(256), 0, st, partial, splits, ws_rows, ldp, V, logits, ld_logits, cand_val, (int*)cand_idx, rstd);
  return 0;
}

// data parallel: append the all-gathered tokens of this step to the global history [step][world * width]
__global__ void dp_unpack_kernel(const int32_t* __restrict__ recv, int n, int32_t* __restrict__ hist, int32_t* __restrict__ dp_step) {
  const int s = *dp_step;
  for (int i = threadIdx.x; i < n; i += blockDim.x) hist[(size_t)s * n + i] = recv[i];
  __syncthreads();
  if (threadIdx.x == 0) *dp_step = s + 1;
}
int dp_unpack(const int32_t* recv, int n, int32_t* hist, int32_t* dp_step, cudaStream_t st) {
  dp_unpack_kernel<<<1, 128, 0, st>>>(recv, n, hist, dp_step);
  VCLA_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// device-side KV page allocator
// ------------------------------------------------------------------------------------------------
__global__ void kv_reset_kernel(int32_t* __restrict__ kv_free, const int32_t* __restrict__ kv_order, int32_t* __restrict__ kv_state,
                                int32_t* __restrict__ kv_npages, int total_pages, int max_batch) {
  // stack top is the END of kv_free: page kv_order[0] must be popped first
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total_pages; i += gridDim.x * blockDim.x) kv_free[total_pages - 1 - i] = kv_order[i];
  if (blockIdx.x == 0) {
    for (int b = threadIdx.x; b < max_batch; b += blockDim.x) kv_npages[b] = 0;
    if (threadIdx.x == 0) { kv_state[0] = total_pages; kv_state[1] = 0; }
  }
}
int kv_reset(int32_t* kv_free, const int32_t* kv_order, int32_t* kv_state, int32_t* kv_npages, int total_pages, int max_batch, cudaStream_t st) {
  int blocks = (total_pages + 255) / 256;
  if (blocks > 64) blocks = 64;
  if (blocks < 1) blocks = 1;
  kv_reset_kernel<<<blocks, 256, 0, st>>>(kv_free, kv_order, kv_state, kv_npages, total_pages, max_batch);
  VCLA_CUDA_OK(cudaGetLastError());
  return 0;
}

// One thread: for page rank p = 0,1,..: every sequence that needs a p-th page and does not own one yet pops the stack.
// need[b] = pages for `tokens[b]` tokens, clamped to the table row.
__device__ void kv_reserve_serial(int32_t* kv_free, int32_t* kv_state, int32_t* kv_npages, int32_t* page_table, int pages_per_seq,
                                  int page_tokens, int B, const int* s_tokens) {
  int top = kv_state[0];
  bool more = true;
  for (int p = 0; more && p < pages_per_seq; ++p) {
    more = false;
    for (int b = 0; b < B; ++b) {
      int need = (s_tokens[b] + page_tokens - 1) / page_tokens;
      if (need > pages_per_seq) need = pages_per_seq;
      if (need > p + 1) more = true;
      if (need > p && kv_npages[b] == p) {
        if (top <= 0) { kv_state[1] = 1; continue; }          // pool exhausted (unreachable behind the host-side capacity check)
        page_table[(size_t)b * pages_per_seq + p] = kv_free[--top];
        kv_npages[b] = p + 1;
      }
    }
  }
  kv_state[0] = top;
}
__global__ void kv_reserve_kernel(int32_t* kv_free, int32_t* kv_state, int32_t* kv_npages, int32_t* page_table, int pages_per_seq,
                                  int page_tokens, int B, int S, const int32_t* __restrict__ left_pad) {
  __shared__ int s_tokens[64];
  if (threadIdx.x < B) s_tokens[threadIdx.x] = S - (left_pad ? left_pad[threadIdx.x] : 0);
  __syncthreads();
  if (threadIdx.x == 0) kv_reserve_serial(kv_free, kv_state, kv_npages, page_table, pages_per_seq, page_tokens, B, s_tokens);
}
int kv_reserve(int32_t* kv_free, int32_t* kv_state, int32_t* kv_npages, int32_t* page_table, int pages_per_seq, int page_tokens, int B, int S,
               const int32_t* left_pad, cudaStream_t st) {
  if (B > 64) { set_error("kv_reserve: batch %d > 64", B); return -1; }
  kv_reserve_kernel<<<1, 64, 0, st>>>(kv_free, kv_state, kv_npages, page_table, pages_per_seq, page_tokens, B, S, left_pad);
  VCLA_CUDA_OK(cudaGetLastError());
  return 0;
}

__global__ void advance_seq_kernel(int32_t* seq_len, int B, int by, const int32_t* __restrict__ left_pad, int32_t* step_idx, int32_t* kv_free,
                                   int32_t* kv_state, int32_t* kv_npages, int32_t* page_table, int pages_per_seq, int page_tokens) {
  __shared__ int s_tokens[64];
  TraceScope trace(12);
  pdl_launch_dependents();   // dependents may become resident early; they block in their own griddepcontrol.wait
  pdl_wait();
  trace.dep();
  const int b = threadIdx.x;
  if (b < B) {
    const int L = seq_len[b] + by - (left_pad ? left_pad[b] : 0);
    seq_len[b] = L;
    s_tokens[b] = L + 1;                 // the next token of this sequence is appended at index L
  }
  if (step_idx != nullptr && b == 0) *step_idx += 1;
  __syncthreads();
  if (threadIdx.x == 0) {
    // fast path (every step but one in page_tokens): nobody crosses a page boundary
    bool any = false;
    for (int i = 0; i < B; ++i) {
      int need = (s_tokens[i] + page_tokens - 1) / page_tokens;
      if (need > pages_per_seq) need = pages_per_seq;
      any |= need > kv_npages[i];
    }
    if (any) kv_reserve_serial(kv_free, kv_state, kv_npages, page_table, pages_per_seq, page_tokens, B, s_tokens);
  }
}
int advance_seq(int32_t* seq_len, int B, int by, const int32_t* left_pad, int32_t* step_idx, int32_t* kv_free, int32_t* kv_state,
                int32_t* kv_npages, int32_t* page_table, int pages_per_seq, int page_tokens, cudaStream_t st) {
  if (B > 64) { set_error("advance_seq: batch %d > 64", B); return -1; }
  VCLA_LAUNCH(advance_seq_kernel, dim3(1), dim3(64), 0, st, seq_len, B, by, left_pad, step_idx, kv_free, kv_state, kv_npages, page_table,
              pages_per_seq, page_tokens);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// weights: synthetic generator (bit-identical to oracle.hash_normal_bf16) and checkpoint repacking
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}
__global__ void fill_hash_normal_kernel(bf16* dst_bf16, float* dst_f32, int64_t n, uint32_t seed, float mul, float offset) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t a = fmix32((uint32_t)i * 0x9E3779B1u + seed);
    const uint32_t b = fmix32(a ^ 0x7F4A7C15u);
    const int tot = (int)(a & 0xFFFFu) + (int)(a >> 16) + (int)(b & 0xFFFFu) + (int)(b >> 16) - 131070;
    const float v = __fadd_rn(__fmul_rn((float)tot, mul), offset);   // no FMA contraction: must match numpy bit for bit
    const bf16 r = __float2bfloat16(v);
    if (dst_bf16) dst_bf16[i] = r;
    if (dst_f32) dst_f32[i] = __bfloat162float(r);
  }
}
int fill_hash_normal(bf16* dst_bf16, float* dst_f32, int64_t n, uint32_t seed, float mul, float offset, cudaStream_t st) {
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  fill_hash_normal_kernel<<<(int)blocks, 256, 0, st>>>(dst_bf16, dst_f32, n, seed, mul, offset);
  VCLA_CUDA_OK(cudaGetLastError());
  return 0;
}

template <typename T, typename O>
__global__ void convert_kernel(const T* src, int64_t n, O* dst) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = to_f32<T>(src[i]);
    if constexpr (sizeof(O) == 2) dst[i] = __float2bfloat16(v);
    else dst[i] = v;
  }
}
template <typename O>
static int convert_any(const void* src, int dtype, int64_t n, O* dst, cudaStream_t st) {
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks == 0) return 0;
  if (dtype == 0) convert_kernel<float, O><<<(int)blocks, 256, 0, st>>>((const float*)src, n, dst);
  else if (dtype == 1) convert_kernel<__half, O><<<(int)blocks, 256, 0, st>>>((const __half*)src, n, dst);
  else if (dtype == 2) convert_kernel<bf16, O><<<(int)blocks, 256, 0, st>>>((const bf16*)src, n, dst);
  else { set_error("convert: unknown dtype %d", dtype); return -1; }
  VCLA_CUDA_OK(cudaGetLastError());
  return 0;
}
int convert_to_bf16(const void* src, int dtype, int64_t n, bf16* dst, cudaStream_t st) { return convert_any<bf16>(src, dtype, n, dst, st); }
int convert_to_f32(const void* src, int dtype, int64_t n, float* dst, cudaStream_t st) { return convert_any<float>(src, dtype, n, dst, st); }

__global__ void copy_rows_bf16_kernel(const bf16* src, int cols, bf16* dst, int ld) {
  const size_t r = blockIdx.x;
  for (int i = threadIdx.x; i < ld; i += blockDim.x) dst[r * ld + i] = (i < cols) ? src[r * cols + i] : __float2bfloat16(0.f);
}
int copy_rows_bf16(const bf16* src, int rows, int cols, bf16* dst, int ld, cudaStream_t st) {
  if (rows == 0) return 0;
  copy_rows_bf16_kernel<<<rows, 256, 0, st>>>(src, cols, dst, ld);
  VCLA_CUDA_OK(cudaGetLastError());
  return 0;
}
__global__ void interleave_rows32_kernel(const bf16* src, int cols, int which, bf16* dst) {
  const size_t j = blockIdx.x;
  const size_t r = (j >> 5) * 64 + (size_t)which * 32 + (j & 31);
  for (int i = threadIdx.x; i < cols; i += blockDim.x) dst[r * cols + i] = src[j * cols + i];
}
int interleave_rows32(const bf16* src, int rows, int cols, int which, bf16* dst, cudaStream_t st) {
  interleave_rows32_kernel<<<rows, 256, 0, st>>>(src, cols, which, dst);
  VCLA_CUDA_OK(cudaGetLastError());
  return 0;
}

VCLA_DEFINE_TRACE_SETTER(trace_set_elementwise)

}  // namespace vcla
