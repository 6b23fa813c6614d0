// Image pre-processing core: the per-thread work of the two resampling kernels and the host-side
// plan (tap tables, geometry, byte->float table), written once so that preprocess.cu runs it on the device and
// tests/preprocess_harness.cpp replays the very same functions thread by thread on the host (test infrastructure; the
// product only ever launches the kernels).
//
// Replaces HF CLIPImageProcessor's PIL pipeline as the reference calls it (models/visualcla/modeling_utils.py:130,
// :150-152, :187-189): resize(shortest_edge=S, BICUBIC) -> center_crop(S) -> x/255 -> (x-mean)/std.  The resize is
// Pillow's ImagingResample for 8-bit pixels: separable, antialiased, taps quantised to 22 fractional bits, a rounding
// shift and an 8-bit clip after each pass, horizontal pass first.  Integer work: bit-exact against
// oracle/clip_preprocess_oracle.py.
#pragma once
#include <math.h>
#include <stdint.h>

#include <vector>

#ifdef __CUDACC__
#define VCLA_PP_HD __host__ __device__ __forceinline__
#else
#define VCLA_PP_HD inline
#endif

namespace vcla_pp {

constexpr int kPrecisionBits = 32 - 8 - 2;     // fractional bits of a tap; keeps 255 * sum|taps| inside int32

struct AxisTaps {        // one axis of the resize; the arrays live where the phases run (device, or host in the harness)
  const int32_t* first;  // [n_out] first source index of the output's window
  const int32_t* count;  // [n_out] taps in the window
  const int32_t* taps;   // [n_out][ksize]
  int ksize;
};

struct Plan {
  const uint8_t* src;    // RGB, (height, width, 3)
  int height, width;
  int out;               // S: side of the square result
  int top, left;         // crop origin inside the resized picture
  int row0, rows;        // source rows the S cropped output rows read:    [row0, row0 + rows)
  int col0, cols;        // source columns the S cropped output columns read: [col0, col0 + cols)
  AxisTaps h, v;         // indexed by coordinates of the RESIZED picture (left + xo, top + yo)
  uint8_t* tmp;          // horizontal-pass result, (rows, out, 3)
  const float* lut;      // (3, 256): byte -> normalised float, per channel
};

VCLA_PP_HD uint8_t clip8(int32_t acc) {
  const int32_t v = acc >> kPrecisionBits;      // arithmetic shift, like Pillow's clip8 lookup index
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// ---- horizontal pass: one block per needed source row ------------------------------------------------------------
// phase A: stage the row's needed byte span in shared memory (coalesced)
VCLA_PP_HD void hpass_stage(const Plan& p, int r, int tid, int nthr, uint8_t* row_smem) {
  const uint8_t* g = p.src + ((size_t)(p.row0 + r) * p.width + p.col0) * 3;
  for (int i = tid; i < p.cols * 3; i += nthr) row_smem[i] = g[i];
}
// phase B: every thread produces output bytes e = xo*3 + c of this row
VCLA_PP_HD void hpass_filter(const Plan& p, int r, int tid, int nthr, const uint8_t* row_smem) {
  for (int e = tid; e < p.out * 3; e += nthr) {
    const int xo = e / 3, c = e - xo * 3;
    const int X = p.left + xo;
    const int first = p.h.first[X] - p.col0, n = p.h.count[X];
    const int32_t* k = p.h.taps + (size_t)X * p.h.ksize;
    int32_t acc = 1 << (kPrecisionBits - 1);
    for (int j = 0; j < n; ++j) acc += (int32_t)row_smem[(first + j) * 3 + c] * k[j];
    p.tmp[((size_t)r * p.out + xo) * 3 + c] = clip8(acc);
  }
}

// ---- vertical pass + crop + normalise: one block per output row --------------------------------------------------
// phase A: filter down the column for each byte of the output row, keep the bytes in shared memory
VCLA_PP_HD void vpass_filter(const Plan& p, int yo, int tid, int nthr, uint8_t* out_smem) {
  const int Y = p.top + yo;
  const int first = p.v.first[Y] - p.row0, n = p.v.count[Y];
  const int32_t* k = p.v.taps + (size_t)Y * p.v.ksize;
  const int stride = p.out * 3;
  for (int e = tid; e < stride; e += nthr) {
    int32_t acc = 1 << (kPrecisionBits - 1);
    const uint8_t* col = p.tmp + (size_t)first * stride + e;
    for (int j = 0; j < n; ++j) acc += (int32_t)col[(size_t)j * stride] * k[j];
    out_smem[e] = clip8(acc);
  }
}
// phase B: planar, channel-major store through the byte->float table; `store(index, value)` writes the caller's dtype
template <class Store>
VCLA_PP_HD void vpass_store(const Plan& p, int yo, int tid, int nthr, const uint8_t* out_smem, Store store) {
  for (int i = tid; i < 3 * p.out; i += nthr) {
    const int c = i / p.out, xo = i - c * p.out;
    store(((size_t)c * p.out + yo) * p.out + xo, p.lut[c * 256 + out_smem[xo * 3 + c]]);
  }
}

// ---- host-side plan (plain host functions: never called from device code) ----------------------------------------
inline double bicubic(double x) {               // Keys kernel, a = -0.5, support 2 (Pillow: bicubic_filter)
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

inline int axis_ksize(int n_in, int n_out) {
  double fs = (double)n_in / n_out;
  if (fs < 1.0) fs = 1.0;
  return (int)ceil(2.0 * fs) * 2 + 1;
}

// Pillow precompute_coeffs + normalize_coeffs_8bpc for the whole-image box.  Arrays: first/count [n_out], taps [n_out*ksize].
inline void build_axis(int n_in, int n_out, int32_t* first, int32_t* count, int32_t* taps) {
  const double scale = (double)n_in / n_out;
  const double fs = scale < 1.0 ? 1.0 : scale;
  const double support = 2.0 * fs, inv = 1.0 / fs;
  const int ksize = axis_ksize(n_in, n_out);
  std::vector<double> w((size_t)ksize);
  for (int xx = 0; xx < n_out; ++xx) {
    const double center = (xx + 0.5) * scale;
    int lo = (int)(center - support + 0.5);
    if (lo < 0) lo = 0;
    int hi = (int)(center + support + 0.5);
    if (hi > n_in) hi = n_in;
    const int n = hi - lo;
    double ww = 0.0;
    for (int x = 0; x < n; ++x) {
      w[x] = bicubic((x + lo - center + 0.5) * inv);
      ww += w[x];
    }
    int32_t* k = taps + (size_t)xx * ksize;
    for (int x = 0; x < ksize; ++x) {
      if (x >= n) { k[x] = 0; continue; }
      double v = w[x];
      if (ww != 0.0) v /= ww;
      v *= (double)(1 << kPrecisionBits);
      k[x] = v < 0 ? (int)(-0.5 + v) : (int)(0.5 + v);
    }
    first[xx] = lo;
    count[xx] = n;
  }
}

struct Geometry {
  int rh, rw;            // resized picture (HF get_resize_output_image_size, shortest_edge, default_to_square=False)
  int top, left;         // HF center_crop box origin
  int kh, kv;            // taps per output, horizontal / vertical
};

inline Geometry geometry(int height, int width, int out) {
  Geometry g;
  const int shrt = width <= height ? width : height, lng = width <= height ? height : width;
  const int new_long = (int)((double)((int64_t)out * lng) / (double)shrt);       // python: int(size * long / short)
  g.rh = width <= height ? new_long : out;
  g.rw = width <= height ? out : new_long;
  g.top = (g.rh - out) / 2;
  g.left = (g.rw - out) / 2;
  g.kh = axis_ksize(width, g.rw);
  g.kv = axis_ksize(height, g.rh);
  return g;
}

// Layout of the caller's workspace: [int32 tables | float lut | uint8 tmp], every part 16-byte aligned.
struct Layout {
  size_t h_first, h_count, h_taps, v_first, v_count, v_taps, lut, tables_end, tmp, total;
};

inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

inline Layout layout(int height, int width, int out) {
  // tmp is sized for the worst case (every source row needed), so the size depends on the picture shape only
  const Geometry g = geometry(height, width, out);
  Layout L;
  size_t o = 0;
  L.h_first = o; o = align16(o + sizeof(int32_t) * (size_t)g.rw);
  L.h_count = o; o = align16(o + sizeof(int32_t) * (size_t)g.rw);
  L.h_taps = o;  o = align16(o + sizeof(int32_t) * (size_t)g.rw * g.kh);
  L.v_first = o; o = align16(o + sizeof(int32_t) * (size_t)g.rh);
  L.v_count = o; o = align16(o + sizeof(int32_t) * (size_t)g.rh);
  L.v_taps = o;  o = align16(o + sizeof(int32_t) * (size_t)g.rh * g.kv);
  L.lut = o;     o = align16(o + sizeof(float) * 3 * 256);
  L.tables_end = o;
  L.tmp = o;     o = align16(o + (size_t)height * out * 3);
  L.total = o;
  return L;
}

// Fills `host` (tables_end bytes) with both tap tables and the byte->float table, and a Plan whose table pointers are
// `base` + offsets (base = the device workspace, or `host` itself in the harness).
inline Plan make_plan(const uint8_t* src, int height, int width, int out, const float mean[3], const float stdv[3],
                      uint8_t* host, uint8_t* base) {
  const Geometry g = geometry(height, width, out);
  const Layout L = layout(height, width, out);
  int32_t* hf = (int32_t*)(host + L.h_first); int32_t* hc = (int32_t*)(host + L.h_count); int32_t* hk = (int32_t*)(host + L.h_taps);
  int32_t* vf = (int32_t*)(host + L.v_first); int32_t* vc = (int32_t*)(host + L.v_count); int32_t* vk = (int32_t*)(host + L.v_taps);
  build_axis(width, g.rw, hf, hc, hk);
  build_axis(height, g.rh, vf, vc, vk);
  float* lut = (float*)(host + L.lut);
  for (int c = 0; c < 3; ++c)
    for (int u = 0; u < 256; ++u) {
      const float x = (float)((double)u * (1.0 / 255.0));       // HF rescale: uint8 * (1/255) in double, cast to float32
      lut[c * 256 + u] = (x - mean[c]) / stdv[c];               // HF normalize: float32 arithmetic
    }
  Plan p;
  p.src = src; p.height = height; p.width = width; p.out = out;
  p.top = g.top; p.left = g.left;
  p.col0 = hf[g.left];
  p.cols = hf[g.left + out - 1] + hc[g.left + out - 1] - p.col0;            // `first` is non-decreasing
  p.row0 = vf[g.top];
  p.rows = vf[g.top + out - 1] + vc[g.top + out - 1] - p.row0;
  p.h.first = (const int32_t*)(base + L.h_first); p.h.count = (const int32_t*)(base + L.h_count);
  p.h.taps = (const int32_t*)(base + L.h_taps); p.h.ksize = g.kh;
  p.v.first = (const int32_t*)(base + L.v_first); p.v.count = (const int32_t*)(base + L.v_count);
  p.v.taps = (const int32_t*)(base + L.v_taps); p.v.ksize = g.kv;
  p.tmp = base + L.tmp;
  p.lut = (const float*)(base + L.lut);
  return p;
}

}  // namespace vcla_pp
