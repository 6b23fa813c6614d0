// Image pre-processing on the device: HF CLIPImageProcessor's PIL pipeline
//   resize(shortest_edge=S, BICUBIC) -> center_crop(S,S) -> x * 1/255 -> (x - mean) / std
// as the reference runs it on the host for every request (models/visualcla/modeling_utils.py:130, :150-152, :187-189),
// for ONE RGB uint8 picture already resident in HBM.  Two byte-streaming kernels (HBM/L2-bound integer work; nothing here
// is GEMM-shaped):
//   pp_hpass_kernel  one CTA per needed source row: the row's needed byte span is staged in shared memory with coalesced
//                    loads, then filtered horizontally for the S cropped columns only         -> tmp (rows, S, 3) u8
//   pp_vpass_kernel  one CTA per output row: vertical filter over tmp (coalesced along the row), crop, byte->float table,
//                    planar (3, S, S) store in the caller's dtype
// Only the cropped window is ever computed: a 1920x1080 picture reads 1080 x 1080 source pixels, not 1920 x 1080.
// The per-thread code lives in preprocess_core.h so tests/preprocess_harness.cpp can replay it on the host bit for bit.
#include <cuda_fp16.h>

#include <vector>

#include "kernels.h"
#include "preprocess_core.h"
#include "../../include/vcla.h"

namespace vcla {
namespace {

constexpr int kThreads = 256;
constexpr size_t kMaxRowSmem = 200 * 1024;

__global__ void __launch_bounds__(kThreads) pp_hpass_kernel(const vcla_pp::Plan p) {
  extern __shared__ uint8_t pp_row[];
  vcla_pp::hpass_stage(p, blockIdx.x, threadIdx.x, blockDim.x, pp_row);
  __syncthreads();
  vcla_pp::hpass_filter(p, blockIdx.x, threadIdx.x, blockDim.x, pp_row);
}

struct StoreF32 {
  float* d;
  __host__ __device__ void operator()(size_t i, float v) const { d[i] = v; }
};
struct StoreF16 {
  __half* d;
  __host__ __device__ void operator()(size_t i, float v) const { d[i] = __float2half_rn(v); }
};
struct StoreBF16 {
  __nv_bfloat16* d;
  __host__ __device__ void operator()(size_t i, float v) const { d[i] = __float2bfloat16_rn(v); }
};

template <class Store>
__global__ void __launch_bounds__(kThreads) pp_vpass_kernel(const vcla_pp::Plan p, Store store) {
  extern __shared__ uint8_t pp_out[];
  vcla_pp::vpass_filter(p, blockIdx.x, threadIdx.x, blockDim.x, pp_out);
  __syncthreads();
  vcla_pp::vpass_store(p, blockIdx.x, threadIdx.x, blockDim.x, pp_out, store);
}

// pictures up to 32768 px a side; the resized long side (out * long / short) is capped so extreme aspect ratios cannot ask
// for gigabyte tap tables
bool bad_shape(int height, int width, int out) {
  if (height < 1 || width < 1 || out < 1 || out > 4096 || height > 32768 || width > 32768) return true;
  const int64_t lng = height > width ? height : width, shrt = height > width ? width : height;
  return (int64_t)out * lng / shrt > 65536;
}

}  // namespace
}  // namespace vcla

extern "C" {

int64_t vcla_preprocess_workspace_bytes(int height, int width, int out_size) {
  if (vcla::bad_shape(height, width, out_size)) {
    vcla::set_error("preprocess: unsupported picture %dx%d -> %d", height, width, out_size);
    return -1;
  }
  return (int64_t)vcla_pp::layout(height, width, out_size).total;
}

int vcla_resample_taps(int in_size, int out_size, int32_t* first, int32_t* count, int32_t* taps, int ksize_capacity) {
  if (in_size < 1 || out_size < 1 || in_size > (1 << 24) || out_size > (1 << 24)) {
    vcla::set_error("resample_taps: bad sizes %d -> %d", in_size, out_size);
    return -1;
  }
  const int ks = vcla_pp::axis_ksize(in_size, out_size);
  if (!first && !count && !taps) return ks;
  if (!first || !count || !taps || ksize_capacity != ks) {
    vcla::set_error("resample_taps: need first/count/taps with ksize_capacity == %d (got %d)", ks, ksize_capacity);
    return -1;
  }
  vcla_pp::build_axis(in_size, out_size, first, count, taps);
  return ks;
}

int vcla_preprocess_image(const uint8_t* rgb_dev, int height, int width, int out_size, const float* mean3, const float* std3,
                          void* workspace_dev, int64_t workspace_bytes, void* pixel_values_dev, int dtype, vcla_stream stream) {
  using namespace vcla;
  if (bad_shape(height, width, out_size)) { set_error("preprocess: unsupported picture %dx%d -> %d", height, width, out_size); return -1; }
  if (!rgb_dev || !mean3 || !std3 || !workspace_dev || !pixel_values_dev) { set_error("preprocess: null argument"); return -1; }
  if (dtype != VCLA_F32 && dtype != VCLA_F16 && dtype != VCLA_BF16) { set_error("preprocess: bad dtype %d", dtype); return -1; }
  for (int c = 0; c < 3; ++c)
    if (!(std3[c] != 0.f)) { set_error("preprocess: std[%d] is zero", c); return -1; }
  if (((uintptr_t)workspace_dev & 15) != 0) { set_error("preprocess: workspace must be 16-byte aligned"); return -1; }
  const vcla_pp::Layout L = vcla_pp::layout(height, width, out_size);
  if (workspace_bytes < (int64_t)L.total) {
    set_error("preprocess: workspace %lld B < %lld B (vcla_preprocess_workspace_bytes)", (long long)workspace_bytes, (long long)L.total);
    return -1;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { (void)cudaGetLastError(); set_error("no CUDA device"); return -1; }

  std::vector<uint8_t> host(L.tables_end);
  const vcla_pp::Plan p = vcla_pp::make_plan(rgb_dev, height, width, out_size, mean3, std3, host.data(), (uint8_t*)workspace_dev);
  const size_t row_smem = (size_t)p.cols * 3, out_smem = (size_t)out_size * 3;
  if (row_smem > kMaxRowSmem) { set_error("preprocess: %d needed columns exceed the shared-memory row buffer", p.cols); return -1; }
  cudaStream_t st = (cudaStream_t)stream;
  // pageable source: the runtime stages the bytes before returning, so `host` may go out of scope afterwards
  VCLA_CUDA_OK(cudaMemcpyAsync(workspace_dev, host.data(), L.tables_end, cudaMemcpyHostToDevice, st));
  static bool attr_done = false;
  if (!attr_done) {
    VCLA_CUDA_OK(cudaFuncSetAttribute(pp_hpass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxRowSmem));
    attr_done = true;
  }
  pp_hpass_kernel<<<dim3(p.rows), dim3(kThreads), row_smem, st>>>(p);
  VCLA_CUDA_OK(cudaGetLastError());
  if (dtype == VCLA_F32) {
    pp_vpass_kernel<StoreF32><<<dim3(out_size), dim3(kThreads), out_smem, st>>>(p, StoreF32{(float*)pixel_values_dev});
  } else if (dtype == VCLA_F16) {
    pp_vpass_kernel<StoreF16><<<dim3(out_size), dim3(kThreads), out_smem, st>>>(p, StoreF16{(__half*)pixel_values_dev});
  } else {
    pp_vpass_kernel<StoreBF16><<<dim3(out_size), dim3(kThreads), out_smem, st>>>(p, StoreBF16{(__nv_bfloat16*)pixel_values_dev});
  }
  VCLA_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
