// TEST INFRASTRUCTURE (built by tests/test_preprocess_core_cpu.py with g++, never part of libvcla.so).
// Replays the per-thread phase functions of csrc/preprocess_core.h -- the exact code the CUDA kernels in
// csrc/preprocess.cu execute -- block by block and thread by thread on the host, so the index arithmetic of the device
// path can be checked bit for bit against the Pillow/HF golden vectors in a container without a GPU.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "preprocess_core.h"

using namespace vcla_pp;

extern "C" int harness_ksize(int n_in, int n_out) { return axis_ksize(n_in, n_out); }

extern "C" void harness_taps(int n_in, int n_out, int32_t* first, int32_t* count, int32_t* taps) {
  build_axis(n_in, n_out, first, count, taps);
}

extern "C" void harness_geometry(int h, int w, int S, int32_t out[10]) {
  const Geometry g = geometry(h, w, S);
  std::vector<uint8_t> ws(layout(h, w, S).total);
  const float one[3] = {1, 1, 1}, zero[3] = {0, 0, 0};
  const Plan p = make_plan(nullptr, h, w, S, zero, one, ws.data(), ws.data());
  const int32_t v[10] = {g.rh, g.rw, g.top, g.left, g.kh, g.kv, p.row0, p.rows, p.col0, p.cols};
  memcpy(out, v, sizeof v);
}

// nthr = emulated block size (any value must give the same result).  pixel_values: (3, S, S) float32; u8: (S, S, 3) or NULL.
extern "C" int harness_preprocess(const uint8_t* rgb, int h, int w, int S, const float* mean, const float* stdv, int nthr,
                                  float* pixel_values, uint8_t* u8) {
  if (h < 1 || w < 1 || S < 1 || nthr < 1) return 1;
  std::vector<uint8_t> ws(layout(h, w, S).total);
  const Plan p = make_plan(rgb, h, w, S, mean, stdv, ws.data(), ws.data());
  if (p.row0 < 0 || p.row0 + p.rows > h || p.col0 < 0 || p.col0 + p.cols > w) return 2;
  std::vector<uint8_t> smem((size_t)(p.cols > S ? p.cols : S) * 3);
  for (int r = 0; r < p.rows; ++r) {                       // grid of the horizontal kernel
    memset(smem.data(), 0xCD, smem.size());                // poison: a read outside the staged span shows up as a mismatch
    for (int t = 0; t < nthr; ++t) hpass_stage(p, r, t, nthr, smem.data());
    /* __syncthreads() */
    for (int t = 0; t < nthr; ++t) hpass_filter(p, r, t, nthr, smem.data());
  }
  for (int yo = 0; yo < S; ++yo) {                         // grid of the vertical kernel
    memset(smem.data(), 0xCD, smem.size());
    for (int t = 0; t < nthr; ++t) vpass_filter(p, yo, t, nthr, smem.data());
    /* __syncthreads() */
    if (u8) memcpy(u8 + (size_t)yo * S * 3, smem.data(), (size_t)S * 3);
    for (int t = 0; t < nthr; ++t)
      vpass_store(p, yo, t, nthr, smem.data(), [&](size_t i, float v) { pixel_values[i] = v; });
  }
  return 0;
}
