// tcgen05 GEMM for sm_100a:  D[M,N] = A[M,K] * B[N,K]^T, bf16 operands, fp32 accumulation in TMEM.
//
// One kernel template serves the whole path:
//   * prefill / ViT / Resampler / projector: A = activations (tokens are the 128-row UMMA M dimension),
//     B = nn.Linear weight [N_out, K]; fused epilogues: bias, quick_gelu / erf-gelu, fp32 residual add
//     with output-row remap (+ position-embedding table), SwiGLU.
//   * decode ("swap-AB"): A = weight [N_out, K] (streamed once from HBM through TMA), B = the few
//     activation rows [batch_pad, K]; split-K partials are written to an fp32 workspace that the
//     next (fused consumer) kernel reduces in a fixed order, so results are deterministic.
//
// Structure (per CTA, 192 threads): warp 0 lane 0 = TMA producer, warp 1 lane 0 = MMA issuer (+ TMEM
// alloc/dealloc by the whole warp), warps 2..5 = epilogue (TMEM lane quadrant = warp_idx % 4).
// smem ring of STAGES x {A 128x64, B BNx64} tiles in the 128B-swizzled K-major layout that both TMA
// and the UMMA shared-memory descriptors understand; two TMEM accumulator stages so the epilogue of
// tile i overlaps the MMAs of tile i+1.  Persistent: each CTA walks tiles blockIdx.x, +gridDim.x, ...
#include "common.cuh"
#include "kernels.h"

#include <mutex>
#include <stdarg.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

namespace vcla {

// ------------------------------------------------------------------------------------------------
// error / device info plumbing shared by all translation units
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }
int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}
static bool g_pdl = false;
bool pdl_enabled() { return g_pdl; }
void set_pdl(bool on) { g_pdl = on; }

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------
struct GemmParams {
  int M, N, K;
  int m_tiles, n_tiles, splits, kb_per_split, kb_total, total_tiles;
  int mode, act, accumulate;
  void* out;
  int ldo;
  const float* bias;
  const float* rowtab;
  int rowtab_period;
  int rows_per_group, group_stride, row_offset;
  int ws_rows;
  int l2_prefetch_kb;
  uint64_t policy_a, policy_b;
  GemmFix fix;
  GemmRowScale rowscale;
  GemmEmitNorm emit;
  GemmRope rope;
};

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kGemmThreads = 192;

template <int BN, int STAGES, int CTAS = 1>
struct GemmCfg {
  static constexpr int A_BYTES = kBlockM * kBlockK * 2;
  static constexpr int B_BYTES = (BN / CTAS) * kBlockK * 2;            // a CTA pair splits the B tile's N rows
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int FIX_OFF = BAR_OFF + 256;             // [4 warps][64] fp32 cross-warp scratch + 1 flag of the split-K fixup
  static constexpr int SMEM_BYTES = FIX_OFF + 1088 + 1024;  // barriers + tmem slot + fixup scratch, + slack for 1024 B alignment
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int CH = BN < 32 ? BN : 32;  // epilogue column chunk
  static_assert(STAGE_BYTES % 1024 == 0, "stage must keep 1024 B alignment for SWIZZLE_128B");
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N constraint for M=128");
};

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == ACT_QUICK_GELU) return x / (1.f + __expf(-1.702f * x));
  if (act == ACT_GELU_ERF) return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
  return x;
}

// CTAS = 2: a CTA PAIR (cluster (2,1,1), two SMs of one TPC) computes a 256 x BN tile with tcgen05.mma.cta_group::2: each CTA stages
// its own 128 A rows and HALF of the B tile, the leader (cluster rank 0) issues the MMAs for both, each CTA's TMEM receives its 128
// rows.  Per SM and k-block this halves the B bytes written to and read from shared memory, which is what bounds the 128 x 256
// single-CTA tile (A 4 KB + B 8 KB per 128-cycle MMA against a 128 B/cycle crossbar that also takes the TMA writes).
template <int BN, int STAGES, bool SWAP, int CTAS = 1>
__global__ void __launch_bounds__(kGemmThreads, SWAP ? 2 : 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using C = GemmCfg<BN, STAGES, CTAS>;
  static_assert(CTAS == 1 || (!SWAP && BN == 256), "the CTA-pair variant is the 256 x 256 prefill tile");
  constexpr bool TWO = CTAS == 2;
  const int crank = TWO ? (int)cluster_rank() : 0;
  const bool leader = crank == 0;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + C::BAR_OFF;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar0 + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + C::BAR_OFF + 8 * (2 * STAGES + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  TraceScope trace(SWAP ? 1 : 2);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4 * CTAS);      // the leader's MMA warp waits for the epilogue warps of BOTH CTAs
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) { if constexpr (TWO) tmem_alloc_2cta(tmem_slot, C::TMEM_COLS); else tmem_alloc(tmem_slot, C::TMEM_COLS); }
  tc_fence_before();
  __syncthreads();
  if constexpr (TWO) cluster_barrier();        // the peer's barriers exist before TMA / commits / arrives target them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // PDL: let the next kernel's CTAs become resident as soon as ours are (they only prefetch read-only weights and
  // then block in griddepcontrol.wait until this grid has completed), see the producer below.
  pdl_launch_dependents();
  const int tile0 = TWO ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;          // a pair walks the pair-tiles together
  const int tile_step = TWO ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  auto tile_coords = [&](int t, int& m_blk, int& n_blk, int& kb0, int& kb1, int& ks) {
    if constexpr (TWO) {
      const int m_pairs = (p.m_tiles + 1) >> 1;
      m_blk = (t % m_pairs) * 2 + crank;       // this CTA's 128 rows of the pair's 256
      n_blk = t / m_pairs;
      ks = 0; kb0 = 0; kb1 = p.kb_total;
      return;
    }
    m_blk = t % p.m_tiles;
    int r = t / p.m_tiles;
    n_blk = r % p.n_tiles;
    ks = r / p.n_tiles;
    kb0 = ks * p.kb_per_split;
    kb1 = min(p.kb_total, kb0 + p.kb_per_split);
  };

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      // The weight operand (A when swap-AB, else B) never depends on the previous kernel, so its tiles for the first
      // STAGES pipeline slots are requested BEFORE griddepcontrol.wait: weight streaming from HBM continues across
      // kernel boundaries.  The activation operand is loaded only after the dependency has resolved.
      int stage = 0;
      uint32_t phase = 0;
      bool dep_ready = false;
      int npend = 0;
      uint32_t pend_dst[STAGES], pend_bar[STAGES];
      int pend_c0[STAGES], pend_c1[STAGES];
      const CUtensorMap* act_map = SWAP ? &tmB : &tmA;
      const uint64_t act_policy = SWAP ? p.policy_b : p.policy_a;
      auto flush_pending = [&]() {
        pdl_wait();
        trace.dep();
        for (int i = 0; i < npend; ++i) {
          if constexpr (TWO) tma_load_2d_2cta(pend_dst[i], act_map, pend_c0[i], pend_c1[i], pend_bar[i], act_policy);
          else tma_load_2d(pend_dst[i], act_map, pend_c0[i], pend_c1[i], pend_bar[i], act_policy);
        }
        npend = 0;
        dep_ready = true;
      };
      for (int t = tile0; t < p.total_tiles; t += tile_step) {
        int m_blk, n_blk, kb0, kb1, ks;
        tile_coords(t, m_blk, n_blk, kb0, kb1, ks);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = base + stage * C::STAGE_BYTES;
          if constexpr (TWO) {
            // both CTAs' bytes are counted on the LEADER's full barrier (the only one the MMA thread waits on)
            if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * C::STAGE_BYTES);
            tma_load_2d_2cta(sa + C::A_BYTES, &tmB, kb * kBlockK, n_blk * BN + crank * (BN / 2), full_bar(stage), p.policy_b);   // weights (half)
            if (dep_ready) {
              tma_load_2d_2cta(sa, &tmA, kb * kBlockK, m_blk * kBlockM, full_bar(stage), p.policy_a);
            } else {
              pend_dst[npend] = sa; pend_bar[npend] = full_bar(stage); pend_c0[npend] = kb * kBlockK; pend_c1[npend] = m_blk * kBlockM; ++npend;
            }
            if (!dep_ready && npend == STAGES) flush_pending();
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            continue;
          }
          mbar_arrive_expect_tx(full_bar(stage), C::STAGE_BYTES);
          if constexpr (SWAP) {
            tma_load_2d(sa, &tmA, kb * kBlockK, m_blk * kBlockM, full_bar(stage), p.policy_a);           // weights
            if (dep_ready) {
              tma_load_2d(sa + C::A_BYTES, &tmB, kb * kBlockK, n_blk * BN, full_bar(stage), p.policy_b);
            } else {
              pend_dst[npend] = sa + C::A_BYTES; pend_bar[npend] = full_bar(stage); pend_c0[npend] = kb * kBlockK; pend_c1[npend] = n_blk * BN; ++npend;
            }
          } else {
            tma_load_2d(sa + C::A_BYTES, &tmB, kb * kBlockK, n_blk * BN, full_bar(stage), p.policy_b);   // weights
            if (dep_ready) {
              tma_load_2d(sa, &tmA, kb * kBlockK, m_blk * kBlockM, full_bar(stage), p.policy_a);
            } else {
              pend_dst[npend] = sa; pend_bar[npend] = full_bar(stage); pend_c0[npend] = kb * kBlockK; pend_c1[npend] = m_blk * kBlockM; ++npend;
            }
          }
          if (!dep_ready && npend == STAGES) {
            // The smem ring is full and the dependency is (probably) still unresolved: HBM would idle while the small
            // consumer kernel in front of us runs.  Pull this CTA's NEXT weight tiles into the 126 MB L2 so the main loop
            // streams them from L2 afterwards.
            if (p.l2_prefetch_kb > 0) {
              int budget = p.l2_prefetch_kb;
              int kbn = kb + 1;
              for (int t2 = t; t2 < p.total_tiles && budget > 0; t2 += gridDim.x) {
                int m2, n2, k0, k1, s2;
                tile_coords(t2, m2, n2, k0, k1, s2);
                for (int kk = (t2 == t ? kbn : k0); kk < k1 && budget > 0; ++kk, --budget) {
                  if constexpr (SWAP) tma_prefetch_l2_2d(&tmA, kk * kBlockK, m2 * kBlockM);
                  else tma_prefetch_l2_2d(&tmB, kk * kBlockK, n2 * BN);
                }
              }
            }
            flush_pending();
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
      if (!dep_ready) flush_pending();
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ===================== MMA issuer (CTA pair: the leader only) =====================
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM * CTAS, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t accphase = 0;
      for (int t = tile0; t < p.total_tiles; t += tile_step) {
        int m_blk, n_blk, kb0, kb1, ks;
        tile_coords(t, m_blk, n_blk, kb0, kb1, ks);
        mbar_wait(tempty_bar(acc), accphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = base + stage * C::STAGE_BYTES;
          const uint64_t adesc = make_desc_sw128(sa);
          const uint64_t bdesc = make_desc_sw128(sa + C::A_BYTES);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // advance 16 elements (32 B) along K inside the 128 B swizzle atom: +2 in the 16 B-unit address field
            if constexpr (TWO) umma_bf16_2cta(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else umma_bf16(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          // smem slot reusable once these MMAs have read it (pair: the slot of BOTH CTAs)
          if constexpr (TWO) umma_commit_2cta(empty_bar(stage), 3); else umma_commit(empty_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        // accumulator complete -> epilogue (pair: each CTA's epilogue reads its own TMEM)
        if constexpr (TWO) umma_commit_2cta(tfull_bar(acc), 3); else umma_commit(tfull_bar(acc));
        acc ^= 1;
        if (acc == 0) accphase ^= 1u;
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;                // TMEM lane quadrant this warp may access
    const int row_in_tile = q * 32 + lane;
    pdl_wait();                            // outputs / residual reads are ordered after the previous grid
    int acc = 0;
    uint32_t accphase = 0;
    for (int t = tile0; t < p.total_tiles; t += tile_step) {
      int m_blk, n_blk, kb0, kb1, ks;
      tile_coords(t, m_blk, n_blk, kb0, kb1, ks);
      mbar_wait(tfull_bar(acc), accphase);
      tc_fence_after();
      const int row = m_blk * kBlockM + row_in_tile;
      const bool row_ok = row < p.M;
      const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);

      if constexpr (SWAP) {
        // rows = weight rows (output features), columns = batch.  Coalesced across lanes.
        float* ws = reinterpret_cast<float*>(p.out);
#pragma unroll
        for (int c = 0; c < BN / C::CH; ++c) {
          uint32_t v[C::CH];
          if constexpr (C::CH == 32) tmem_ld_32x32(taddr0 + c * C::CH, v);
          else tmem_ld_32x16(taddr0 + c * C::CH, reinterpret_cast<uint32_t(&)[16]>(v));
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int i = 0; i < C::CH; ++i) {
              const int b = c * C::CH + i;
              if (b < p.ws_rows) ws[((size_t)ks * p.ws_rows + b) * (size_t)p.ldo + row] = __uint_as_float(v[i]);
            }
          }
        }
        if (p.fix.mode != FIX_NONE) {
          // hand the accumulator stage back first: the MMA warp runs ahead on the next tile while we fix this one up
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(acc));
          // ---- split-K fixup: the CTA whose partial completes the tile reduces all partials (fixed order) and applies
          //      the fused consumer.  Classic last-block pattern: fence, count, fence.
          volatile int* s_flag = reinterpret_cast<volatile int*>(base_ptr + C::FIX_OFF + 1024);
          float* s_red = reinterpret_cast<float*>(base_ptr + C::FIX_OFF);          // [4][64]
          __threadfence();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (warp == 2 && lane == 0) {
            const int old = atomicAdd(p.fix.tile_counters + m_blk, 1);
            *s_flag = (old == p.splits - 1) ? 1 : 0;
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (*s_flag) {
            __threadfence();
            const int nb = p.ws_rows;                 // batch rows
            if (p.fix.mode == FIX_RESID) {
              for (int b0 = 0; b0 < nb; b0 += 8) {    // 8 batch rows per pass; partial loads issued 4 splits x 8 rows at a time
                float r[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) r[j] = 0.f;
                if (row_ok) {
#pragma unroll
                  for (int j = 0; j < 8; ++j) if (b0 + j < nb) r[j] = p.fix.resid[(size_t)(b0 + j) * p.ldo + row];
                  for (int sp0 = 0; sp0 < p.splits; sp0 += 4) {
                    float t4[4][8];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                      for (int j = 0; j < 8; ++j)
                        t4[u][j] = (sp0 + u < p.splits && b0 + j < nb) ? __ldcg(ws + ((size_t)(sp0 + u) * nb + b0 + j) * (size_t)p.ldo + row) : 0.f;
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                      for (int j = 0; j < 8; ++j) r[j] += t4[u][j];      // fixed split order: deterministic
                  }
                  const float wn = __ldg(p.fix.norm_w + row);
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    if (b0 + j < nb) {
                      p.fix.resid[(size_t)(b0 + j) * p.ldo + row] = r[j];
                      p.fix.xw_out[(size_t)(b0 + j) * p.ldo + row] = __float2bfloat16(r[j] * wn);
                    }
                  }
                }
                // sum of squares over the tile's 128 rows, per batch row
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  float q2 = warp_sum(r[j] * r[j]);
                  if (lane == 0) s_red[q * 64 + b0 + j] = q2;
                }
              }
              asm volatile("bar.sync 1, 128;" ::: "memory");
              const int t = (warp - 2) * 32 + lane;     // 0..127
              if (t < nb) p.fix.ssq[(size_t)t * p.m_tiles + m_blk] = s_red[t] + s_red[64 + t] + s_red[128 + t] + s_red[192 + t];
              // second level: the CTA that finishes the last tile turns the per-tile sums into the per-row scale
              __threadfence();
              asm volatile("bar.sync 1, 128;" ::: "memory");
              if (warp == 2 && lane == 0) {
                const int old = atomicAdd(p.fix.tile_counters + p.m_tiles, 1);
                *s_flag = (old == p.m_tiles - 1) ? 2 : 1;
              }
              asm volatile("bar.sync 1, 128;" ::: "memory");
              if (*s_flag == 2) {
                __threadfence();
                if (t < nb) {
                  float ss = 0.f;
                  for (int t0 = 0; t0 < p.m_tiles; t0 += 8) {
                    float v8[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) v8[u] = (t0 + u < p.m_tiles) ? __ldcg(p.fix.ssq + (size_t)t * p.m_tiles + t0 + u) : 0.f;
#pragma unroll
                    for (int u = 0; u < 8; ++u) ss += v8[u];
                  }
                  p.fix.rstd_out[t] = rsqrtf(ss * p.fix.inv_dim + p.fix.eps);
                }
                if (warp == 2 && lane == 0) p.fix.tile_counters[p.m_tiles] = 0;
              }
            } else {  // FIX_SWIGLU: tile rows = [32 gate | 32 up | 32 gate | 32 up]
              const int blk = row_in_tile >> 6, within = row_in_tile & 63;
              const int gi = within & 31;                       // pair index inside the 64-row block
              const int grow = m_blk * kBlockM + blk * 64 + gi; // gate row ; up row = grow + 32
              const int jout = m_blk * 64 + blk * 32 + gi;      // output feature
              const bool pair_ok = (grow + 32) < p.M;
              const int half = (nb + 1) >> 1;
              const int bs = (within < 32) ? 0 : half, be = (within < 32) ? half : nb;   // the two partner lanes split the batch
              for (int b0 = bs; b0 < be; b0 += 4) {             // 4 batch rows x all splits in flight
                float g[4] = {0.f, 0.f, 0.f, 0.f}, u[4] = {0.f, 0.f, 0.f, 0.f};
                if (pair_ok) {
                  for (int sp = 0; sp < p.splits; ++sp) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      if (b0 + j < be) {
                        const float* pr = ws + ((size_t)sp * nb + b0 + j) * (size_t)p.ldo;
                        g[j] += __ldcg(pr + grow);
                        u[j] += __ldcg(pr + grow + 32);
                      }
                    }
                  }
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    if (b0 + j < be) {
                      const float rstd = __ldcg(p.fix.rstd_in + b0 + j);
                      const float gg = g[j] * rstd, uu = u[j] * rstd;
                      p.fix.h_out[(size_t)(b0 + j) * (p.ldo >> 1) + jout] = __float2bfloat16(gg / (1.f + __expf(-gg)) * uu);
                    }
                  }
                }
              }
            }
            __syncwarp();
            if (warp == 2 && lane == 0) p.fix.tile_counters[m_blk] = 0;   // ready for the next launch / graph replay
          }
          acc ^= 1;
          if (acc == 0) accphase ^= 1u;
          continue;
        }
      } else {
        int orow = row;
        if (p.rows_per_group > 0) orow = (row / p.rows_per_group) * p.group_stride + (row % p.rows_per_group) + p.row_offset;
        const int col_tile = n_blk * BN;
        // deferred RMSNorm: the operand rows were not normalised; their scale commutes with the GEMM and lands here
        float rs = 1.f;
        if (p.rowscale.ssq != nullptr && row_ok) {
          const float* sp = p.rowscale.ssq + (size_t)row * p.rowscale.slots;
          float ss = 0.f;
          for (int i = 0; i < p.rowscale.slots; ++i) ss += __ldg(sp + i);       // fixed order: deterministic
          rs = rsqrtf(ss * p.rowscale.inv_dim + p.rowscale.eps);
        }
        if (p.rope.cos != nullptr) {
          // ---- fused QKV epilogue: RoPE on q / k heads, store q|k|v rows, append k / v to the paged cache (BN = 256 = 2 heads)
          if constexpr (BN == 256) {
            const GemmRope& R = p.rope;
            const int bq = row_ok ? row / R.S : 0, sq = row_ok ? row % R.S : 0;
            const int pad = (R.left_pad != nullptr && row_ok) ? __ldg(R.left_pad + bq) : 0;
            const int cpos = sq - pad;                               // index inside the (compact) KV cache
            const bool cached = row_ok && cpos >= 0;                 // padding rows are neither rotated nor cached
            const int pos = R.pos_from_mask ? (cpos > 0 ? cpos : 0) : sq;
            int page = 0, slot = 0;
            if (cached) { page = __ldg(R.page_table + (size_t)bq * R.pages_per_seq + cpos / R.page_tokens); slot = cpos % R.page_tokens; }
            const float* ct = R.cos + (size_t)pos * 64;
            const float* stb = R.sin + (size_t)pos * 64;
            bf16* orow_ptr = reinterpret_cast<bf16*>(p.out) + (size_t)orow * p.ldo;
#pragma unroll 1
            for (int hh = 0; hh < 2; ++hh) {
              const int hcol = col_tile + hh * 128;                  // first column of this head inside [q | k | v]
              if (hcol >= p.N) break;
              const int region = hcol / R.T, head = (hcol % R.T) / 128;
#pragma unroll 1
              for (int half = 0; half < 2; ++half) {                 // dims [32 half, 32 half + 32) pair with [64 + 32 half, ...)
                uint32_t lo[32], hi[32];
                tmem_ld_32x32(taddr0 + hh * 128 + half * 32, lo);
                tmem_ld_32x32(taddr0 + hh * 128 + 64 + half * 32, hi);
                tmem_ld_wait();
                if (!row_ok) continue;
                uint32_t plo[16], phi[16];
                if (region < 2 && cached) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    const float4 c4 = __ldg(reinterpret_cast<const float4*>(ct + half * 32) + i);
                    const float4 s4 = __ldg(reinterpret_cast<const float4*>(stb + half * 32) + i);
                    const float cc[4] = {c4.x, c4.y, c4.z, c4.w}, sn[4] = {s4.x, s4.y, s4.z, s4.w};
                    float ol[4], oh[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      const float a = __uint_as_float(lo[4 * i + j]) * rs, b2 = __uint_as_float(hi[4 * i + j]) * rs;
                      ol[j] = a * cc[j] - b2 * sn[j];
                      oh[j] = b2 * cc[j] + a * sn[j];
                    }
                    plo[2 * i] = pack_bf16x2(ol[0], ol[1]); plo[2 * i + 1] = pack_bf16x2(ol[2], ol[3]);
                    phi[2 * i] = pack_bf16x2(oh[0], oh[1]); phi[2 * i + 1] = pack_bf16x2(oh[2], oh[3]);
                  }
                } else {
#pragma unroll
                  for (int i = 0; i < 16; ++i) {
                    plo[i] = pack_bf16x2(__uint_as_float(lo[2 * i]) * rs, __uint_as_float(lo[2 * i + 1]) * rs);
                    phi[i] = pack_bf16x2(__uint_as_float(hi[2 * i]) * rs, __uint_as_float(hi[2 * i + 1]) * rs);
                  }
                }
                uint4* d0 = reinterpret_cast<uint4*>(orow_ptr + hcol + half * 32);
                uint4* d1 = reinterpret_cast<uint4*>(orow_ptr + hcol + 64 + half * 32);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  d0[i] = make_uint4(plo[4 * i], plo[4 * i + 1], plo[4 * i + 2], plo[4 * i + 3]);
                  d1[i] = make_uint4(phi[4 * i], phi[4 * i + 1], phi[4 * i + 2], phi[4 * i + 3]);
                }
                if (region >= 1 && cached) {
                  bf16* cdst = R.kv_pages + ((((size_t)page * 2 + (region - 1)) * R.H + head) * R.page_tokens + slot) * 128;
                  uint4* c0 = reinterpret_cast<uint4*>(cdst + half * 32);
                  uint4* c1 = reinterpret_cast<uint4*>(cdst + 64 + half * 32);
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    c0[i] = make_uint4(plo[4 * i], plo[4 * i + 1], plo[4 * i + 2], plo[4 * i + 3]);
                    c1[i] = make_uint4(phi[4 * i], phi[4 * i + 1], phi[4 * i + 2], phi[4 * i + 3]);
                  }
                }
              }
            }
          }
        } else if (p.mode == GEMM_SWIGLU_BF16) {
          if constexpr (BN % 64 == 0) {
            bf16* out = reinterpret_cast<bf16*>(p.out);
#pragma unroll 1
            for (int c = 0; c < BN / 64; ++c) {
              uint32_t g[32], u[32];
              tmem_ld_32x32(taddr0 + c * 64, g);
              tmem_ld_32x32(taddr0 + c * 64 + 32, u);
              tmem_ld_wait();
              const int col0 = col_tile + c * 64;            // first gate column of this pair (in interleaved space)
              if (row_ok && col0 < p.N) {
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  float g0 = __uint_as_float(g[2 * i]) * rs, g1 = __uint_as_float(g[2 * i + 1]) * rs;
                  float u0 = __uint_as_float(u[2 * i]) * rs, u1 = __uint_as_float(u[2 * i + 1]) * rs;
                  float h0 = g0 / (1.f + __expf(-g0)) * u0;
                  float h1 = g1 / (1.f + __expf(-g1)) * u1;
                  pk[i] = pack_bf16x2(h0, h1);
                }
                uint4* dst = reinterpret_cast<uint4*>(out + (size_t)orow * p.ldo + col0 / 2);
#pragma unroll
                for (int i = 0; i < 4; ++i) dst[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
              }
            }
          }
        } else {
          float ssq_acc = 0.f;     // GemmEmitNorm: this tile's share of the row's sum of squares
#pragma unroll 1
          for (int c = 0; c < BN / C::CH; ++c) {
            uint32_t v[C::CH];
            if constexpr (C::CH == 32) tmem_ld_32x32(taddr0 + c * C::CH, v);
            else tmem_ld_32x16(taddr0 + c * C::CH, reinterpret_cast<uint32_t(&)[16]>(v));
            tmem_ld_wait();
            const int col0 = col_tile + c * C::CH;
            if (!row_ok || col0 >= p.N) continue;
            const bool full = (col0 + C::CH <= p.N);
            float x[C::CH];
#pragma unroll
            for (int i = 0; i < C::CH; ++i) x[i] = __uint_as_float(v[i]) * rs;
            if (p.bias != nullptr) {
              if (full) {
                const float4* bp = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
                for (int i = 0; i < C::CH / 4; ++i) {
                  float4 b4 = __ldg(bp + i);
                  x[4 * i] += b4.x; x[4 * i + 1] += b4.y; x[4 * i + 2] += b4.z; x[4 * i + 3] += b4.w;
                }
              } else {
#pragma unroll
                for (int i = 0; i < C::CH; ++i) if (col0 + i < p.N) x[i] += __ldg(p.bias + col0 + i);
              }
            }
            if (p.mode == GEMM_STORE_BF16) {
              if (p.act != ACT_NONE) {
#pragma unroll
                for (int i = 0; i < C::CH; ++i) x[i] = apply_act(x[i], p.act);
              }
              bf16* out = reinterpret_cast<bf16*>(p.out) + (size_t)orow * p.ldo + col0;
              if (full && (p.ldo & 7) == 0) {
                uint4* dst = reinterpret_cast<uint4*>(out);
#pragma unroll
                for (int i = 0; i < C::CH / 8; ++i)
                  dst[i] = make_uint4(pack_bf16x2(x[8 * i], x[8 * i + 1]), pack_bf16x2(x[8 * i + 2], x[8 * i + 3]),
                                      pack_bf16x2(x[8 * i + 4], x[8 * i + 5]), pack_bf16x2(x[8 * i + 6], x[8 * i + 7]));
              } else {
#pragma unroll
                for (int i = 0; i < C::CH; ++i) if (col0 + i < p.N) out[i] = __float2bfloat16(x[i]);
              }
            } else {  // GEMM_ADD_F32
              float* out = reinterpret_cast<float*>(p.out) + (size_t)orow * p.ldo + col0;
              const float* rt = p.rowtab ? p.rowtab + (size_t)(row % p.rowtab_period) * p.N + col0 : nullptr;
              if (full && (p.ldo & 3) == 0 && (p.N & 3) == 0) {
                // all loads first, then the arithmetic, then all stores: with the loads interleaved between stores to a second
                // (possibly aliasing) output the compiler has to serialise one L2 round trip per float4
                float4* __restrict__ dst = reinterpret_cast<float4*>(out);
                float4 oldv[C::CH / 4];
                if (p.accumulate) {
#pragma unroll
                  for (int i = 0; i < C::CH / 4; ++i) oldv[i] = dst[i];
                }
#pragma unroll
                for (int i = 0; i < C::CH / 4; ++i) {
                  if (rt) {
                    const float4 r4 = __ldg(reinterpret_cast<const float4*>(rt) + i);
                    x[4 * i] += r4.x; x[4 * i + 1] += r4.y; x[4 * i + 2] += r4.z; x[4 * i + 3] += r4.w;
                  }
                  if (p.accumulate) { x[4 * i] += oldv[i].x; x[4 * i + 1] += oldv[i].y; x[4 * i + 2] += oldv[i].z; x[4 * i + 3] += oldv[i].w; }
                }
#pragma unroll
                for (int i = 0; i < C::CH / 4; ++i) dst[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
                if (p.emit.xw != nullptr) {
                  uint2* __restrict__ xd = reinterpret_cast<uint2*>(p.emit.xw + (size_t)orow * p.emit.ldxw + col0);
                  const float4* __restrict__ wp = reinterpret_cast<const float4*>(p.emit.norm_w + col0);
#pragma unroll
                  for (int i = 0; i < C::CH / 4; ++i) {
                    const float4 w4 = __ldg(wp + i);
                    ssq_acc += x[4 * i] * x[4 * i] + x[4 * i + 1] * x[4 * i + 1] + x[4 * i + 2] * x[4 * i + 2] + x[4 * i + 3] * x[4 * i + 3];
                    xd[i] = make_uint2(pack_bf16x2(x[4 * i] * w4.x, x[4 * i + 1] * w4.y), pack_bf16x2(x[4 * i + 2] * w4.z, x[4 * i + 3] * w4.w));
                  }
                }
              } else {
#pragma unroll
                for (int i = 0; i < C::CH; ++i) {
                  if (col0 + i < p.N) {
                    float o = x[i];
                    if (rt) o += __ldg(rt + i);
                    if (p.accumulate) o += out[i];
                    out[i] = o;
                    if (p.emit.xw != nullptr) {
                      ssq_acc += o * o;
                      p.emit.xw[(size_t)orow * p.emit.ldxw + col0 + i] = __float2bfloat16(o * __ldg(p.emit.norm_w + col0 + i));
                    }
                  }
                }
              }
            }
          }
          if (p.emit.ssq_out != nullptr && row_ok) p.emit.ssq_out[(size_t)orow * p.n_tiles + n_blk] = ssq_acc;
        }
      }
      // release this accumulator stage back to the MMA warp (pair: the leader's)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (TWO && !leader) mbar_arrive_cluster(tempty_bar(acc), 0); else mbar_arrive(tempty_bar(acc)); }
      acc ^= 1;
      if (acc == 0) accphase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (TWO) cluster_barrier();        // the peer may still address this CTA's barriers / shared memory / TMEM
  trace.done();
  if (warp == 1) { if constexpr (TWO) tmem_dealloc_2cta(tmem_base, C::TMEM_COLS); else tmem_dealloc(tmem_base, C::TMEM_COLS); }
}

VCLA_DEFINE_TRACE_SETTER(trace_set_gemm)

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;
static std::once_flag g_gemm_once;
static int g_gemm_init_rc = 0;

template <int BN, int STAGES, bool SWAP, int CTAS = 1>
static int set_attr() {
  using C = GemmCfg<BN, STAGES, CTAS>;
  VCLA_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, SWAP, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  return 0;
}

// tile configurations (BN, STAGES): prefill 256x4 / 128x6 / 64x8 ; decode swap-AB 16x5 / 32x5 / 64x4 (two CTAs per SM)
static int gemm_init_impl() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    set_error("cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
    return -1;
  }
  g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
  if (set_attr<256, 4, false>()) return -1;
  if (set_attr<256, 6, false, 2>()) return -1;
  if (set_attr<128, 6, false>()) return -1;
  if (set_attr<64, 8, false>()) return -1;
  if (set_attr<16, 5, true>()) return -1;
  if (set_attr<32, 5, true>()) return -1;
  if (set_attr<64, 4, true>()) return -1;
  return 0;
}
int gemm_init() {
  std::call_once(g_gemm_once, [] { g_gemm_init_rc = gemm_init_impl(); });
  return g_gemm_init_rc;
}

static int make_tmap(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld * 2) % 16 != 0) {
    set_error("TMA operand must be 16 B aligned with a 16 B-multiple row pitch (ptr %p ld %llu)", ptr, (unsigned long long)ld);
    return -1;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBlockK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows %llu cols %llu ld %llu box %u", (int)r, (unsigned long long)rows,
              (unsigned long long)cols, (unsigned long long)ld, box_rows);
    return -1;
  }
  return 0;
}

template <int BN, int STAGES, bool SWAP, int CTAS = 1>
static int launch(const GemmCall& c, GemmParams p, cudaStream_t st) {
  using C = GemmCfg<BN, STAGES, CTAS>;
  CUtensorMap ta, tb;
  if (make_tmap(&ta, c.A, c.M, c.K, c.lda, kBlockM)) return -1;
  if (make_tmap(&tb, c.B, c.N, c.K, c.ldb, BN / CTAS)) return -1;
  p.n_tiles = (c.N + BN - 1) / BN;
  p.total_tiles = (CTAS == 2 ? (p.m_tiles + 1) / 2 : p.m_tiles) * p.n_tiles * p.splits;     // pair-tiles of 256 rows
  const int slots = CTAS == 2 ? num_sms() / 2 : num_sms() * (SWAP ? 2 : 1);
  const int grid = (p.total_tiles < slots ? p.total_tiles : slots) * CTAS;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int nattr = 0;
  if (CTAS == 2) {
    attr[nattr].id = cudaLaunchAttributeClusterDimension;
    attr[nattr].val.clusterDim.x = 2; attr[nattr].val.clusterDim.y = 1; attr[nattr].val.clusterDim.z = 1;
    ++nattr;
  }
  if (pdl_enabled()) {
    attr[nattr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[nattr].val.programmaticStreamSerializationAllowed = 1;
    ++nattr;
  }
  cfg.attrs = attr;
  cfg.numAttrs = nattr;
  if (getenv("VCLA_DEBUG")) {
    static bool once = false;
    if (!once) {
      once = true;
      int per_sm = -1;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gemm_tc_kernel<BN, STAGES, SWAP, CTAS>, kGemmThreads, C::SMEM_BYTES);
      fprintf(stderr, "[vcla] gemm_tc<%d,%d,%d,%d>: smem %d, occupancy query %d blocks/SM, grid %d\n", BN, STAGES, (int)SWAP, CTAS, C::SMEM_BYTES, per_sm, grid);
    }
  }
  VCLA_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, STAGES, SWAP, CTAS>, ta, tb, p));
  return 0;
}

// CTA-pair tiles (cta_group::2): on by default for the 256-wide prefill tile; VCLA_GEMM_2CTA=0 keeps the single-CTA 128 x 256 tile
static int g_two_cta = -1;
static bool two_cta_enabled() {
  if (g_two_cta < 0) { const char* e = getenv("VCLA_GEMM_2CTA"); g_two_cta = (e == nullptr) ? 1 : (atoi(e) != 0); }
  return g_two_cta != 0;
}
void gemm_set_two_cta(int on) { g_two_cta = on ? 1 : 0; }

// tile width of a non-swap GEMM with M rows and N output columns (the widest tile that still gives every SM work)
int gemm_pick_bn(int M, int N) {
  const long m_tiles = (M + kBlockM - 1) / kBlockM;
  const long tiles256 = m_tiles * ((N + 255) / 256);
  const long tiles128 = m_tiles * ((N + 127) / 128);
  if (tiles256 >= num_sms() || N >= 4096) return 256;
  if (tiles128 >= num_sms() / 2 || N > 64) return 128;
  return 64;
}

int gemm_tc(const GemmCall& c, cudaStream_t st) {
  if (gemm_init()) return -1;
  if (c.M <= 0 || c.N <= 0 || c.K <= 0) { set_error("gemm: empty problem"); return -1; }
  if (c.K % 8 != 0) { set_error("gemm: K (%d) must be a multiple of 8", c.K); return -1; }
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = c.M; p.N = c.N; p.K = c.K;
  p.m_tiles = (c.M + kBlockM - 1) / kBlockM;
  p.kb_total = (c.K + kBlockK - 1) / kBlockK;
  p.mode = c.mode; p.act = c.act; p.accumulate = c.accumulate;
  p.out = c.out; p.ldo = c.ldo; p.bias = c.bias;
  p.rowtab = c.rowtab; p.rowtab_period = c.rowtab_period > 0 ? c.rowtab_period : 1;
  p.rows_per_group = c.rows_per_group; p.group_stride = c.group_stride; p.row_offset = c.row_offset;
  p.ws_rows = c.ws_rows;
  p.l2_prefetch_kb = c.l2_prefetch_kb;
  p.fix = c.fix;
  p.rowscale = c.rowscale; p.emit = c.emit; p.rope = c.rope;
  p.policy_a = c.weights_are_A ? kEvictFirst : kEvictLast;
  p.policy_b = c.weights_are_A ? kEvictLast : kEvictNormal;

  if (c.mode == GEMM_PARTIAL_F32) {
    int splits = c.splits > 0 ? c.splits : 1;
    if (splits > p.kb_total) splits = p.kb_total;
    p.kb_per_split = (p.kb_total + splits - 1) / splits;
    p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;   // every split non-empty
    if (p.splits != splits) { set_error("gemm: split count %d not realisable for %d k-blocks (use %d)", splits, p.kb_total, p.splits); return -1; }
    if (c.N > 64 || c.ws_rows < c.N) { set_error("gemm: swap-AB batch rows %d (ws_rows %d) unsupported", c.N, c.ws_rows); return -1; }
    if (c.fix.mode != FIX_NONE && (c.fix.tile_counters == nullptr || c.ws_rows != c.N || c.ldo != c.M)) { set_error("gemm: split-K fixup needs tile counters, ws_rows == batch and ldo == rows"); return -1; }
    if (c.fix.mode == FIX_SWIGLU && (c.M % 64) != 0) { set_error("gemm: SwiGLU fixup needs rows %% 64 == 0"); return -1; }
    if (c.N <= 16) return launch<16, 5, true>(c, p, st);
    if (c.N <= 32) return launch<32, 5, true>(c, p, st);
    return launch<64, 4, true>(c, p, st);
  }
  p.splits = 1;
  p.kb_per_split = p.kb_total;
  if (c.mode == GEMM_SWIGLU_BF16 && (c.N % 64) != 0) { set_error("gemm: SwiGLU needs N %% 64 == 0 (N=%d)", c.N); return -1; }
  if (c.emit.xw != nullptr && !(c.mode == GEMM_ADD_F32 && c.accumulate && c.emit.norm_w && c.emit.ssq_out)) { set_error("gemm: emit-norm needs ADD_F32 + accumulate + norm_w + ssq_out"); return -1; }
  if (c.rope.cos != nullptr && !(c.mode == GEMM_STORE_BF16 && c.N == 3 * c.rope.T && c.rope.T % 128 == 0 && c.bias == nullptr && c.act == ACT_NONE && c.rows_per_group == 0 && (c.ldo % 8) == 0)) {
    set_error("gemm: the RoPE + KV-append epilogue needs the plain fused QKV projection (N = 3T, T %% 128 == 0)"); return -1;
  }
  int bn = c.bn;
  if (c.rope.cos != nullptr) bn = 256;    // two whole heads per tile
  if (bn == 0) bn = gemm_pick_bn(c.M, c.N);
  if (bn == 256 && two_cta_enabled() && p.m_tiles >= 2) return launch<256, 6, false, 2>(c, p, st);
  if (bn == 256) return launch<256, 4, false>(c, p, st);
  if (bn == 128) return launch<128, 6, false>(c, p, st);
  if (bn == 64) return launch<64, 8, false>(c, p, st);
  set_error("gemm: unsupported tile N %d", bn);
  return -1;
}

// ------------------------------------------------------------------------------------------------
// naive reference (tests only): one thread per output element, same epilogue semantics
// ------------------------------------------------------------------------------------------------
__global__ void gemm_naive_kernel(const bf16* A, const bf16* B, GemmParams p, int lda, int ldb) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = blockIdx.y;
  const int ncols = (p.mode == GEMM_SWIGLU_BF16) ? p.N / 2 : p.N;
  if (col >= ncols || row >= p.M) return;
  auto dot = [&](int n, int k0, int k1) {
    float s = 0.f;
    for (int k = k0; k < k1; ++k) s += __bfloat162float(A[(size_t)row * lda + k]) * __bfloat162float(B[(size_t)n * ldb + k]);
    return s;
  };
  int orow = row;
  if (p.rows_per_group > 0) orow = (row / p.rows_per_group) * p.group_stride + (row % p.rows_per_group) + p.row_offset;
  if (p.mode == GEMM_PARTIAL_F32) {
    float* ws = reinterpret_cast<float*>(p.out);
    for (int s = 0; s < p.splits; ++s) {
      int k0 = s * p.kb_per_split * kBlockK, k1 = min(p.K, (s + 1) * p.kb_per_split * kBlockK);
      ws[((size_t)s * p.ws_rows + col) * p.ldo + row] = dot(col, k0, k1);
    }
    return;
  }
  if (p.mode == GEMM_SWIGLU_BF16) {
    const int gcol = (col / 32) * 64 + (col % 32);
    float g = dot(gcol, 0, p.K), u = dot(gcol + 32, 0, p.K);
    reinterpret_cast<bf16*>(p.out)[(size_t)orow * p.ldo + col] = __float2bfloat16(g / (1.f + expf(-g)) * u);
    return;
  }
  float x = dot(col, 0, p.K);
  if (p.bias) x += p.bias[col];
  if (p.mode == GEMM_STORE_BF16) {
    if (p.act == ACT_QUICK_GELU) x = x / (1.f + expf(-1.702f * x));
    if (p.act == ACT_GELU_ERF) x = 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
    reinterpret_cast<bf16*>(p.out)[(size_t)orow * p.ldo + col] = __float2bfloat16(x);
  } else {
    float* o = reinterpret_cast<float*>(p.out) + (size_t)orow * p.ldo + col;
    if (p.rowtab) x += p.rowtab[(size_t)(row % p.rowtab_period) * p.N + col];
    if (p.accumulate) x += *o;
    *o = x;
  }
}

int gemm_naive(const GemmCall& c, cudaStream_t st) {
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = c.M; p.N = c.N; p.K = c.K;
  p.kb_total = (c.K + kBlockK - 1) / kBlockK;
  p.mode = c.mode; p.act = c.act; p.accumulate = c.accumulate;
  p.out = c.out; p.ldo = c.ldo; p.bias = c.bias;
  p.rowtab = c.rowtab; p.rowtab_period = c.rowtab_period > 0 ? c.rowtab_period : 1;
  p.rows_per_group = c.rows_per_group; p.group_stride = c.group_stride; p.row_offset = c.row_offset;
  p.ws_rows = c.ws_rows;
  p.splits = 1; p.kb_per_split = p.kb_total;
  if (c.mode == GEMM_PARTIAL_F32) {
    int splits = c.splits > 0 ? c.splits : 1;
    p.kb_per_split = (p.kb_total + splits - 1) / splits;
    p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  }
  const int ncols = (c.mode == GEMM_SWIGLU_BF16) ? c.N / 2 : c.N;
  dim3 grid((ncols + 127) / 128, c.M);
  gemm_naive_kernel<<<grid, 128, 0, st>>>(c.A, c.B, p, c.lda, c.ldb);
  VCLA_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace vcla
