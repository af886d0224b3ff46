// image_io.cu — the byte-level ends of the inference path and of DSM tile preparation as batched device kernels
// (SURVEY.md §8f rank 3 and the normalisation half of rank 4).
//
// Reference code replaced (all of it per tile, on the host, through PIL / numpy):
//   mvp_gan/src/evaluate.py:21-33   Image 'L' -> Resize((512,512)) -> ToTensor (u8 / 255) ; mask = (mask > 0) ;
//                                   masked_img = image * mask
//   mvp_gan/src/evaluate.py:53-59   (output * 255).astype('uint8') -> Image.resize((500,500), BILINEAR)
//   utils/data_extraction.py:80-107 NaN-aware min-max normalisation of a DSM tile to uint8 and resize to 512x512
// The resize is Pillow's: libImaging/Resample.c — support-scaled triangle filter, coefficients normalised in double
// and converted to 22-bit fixed point, horizontal pass then vertical pass, each rounding to uint8. tg_resize_coeffs
// builds the same tables (host); the kernels do the same integer arithmetic, so results are BIT-EXACT with PIL
// (tests/golden/image_io.npz was produced by Pillow itself).
#include <math.h>

#include "tg_common.cuh"
#include "../../include/terragan_b200.h"

namespace tg {

constexpr int kPrecisionBits = 32 - 8 - 2;

// ---- u8 tile + u8 mask -> fp32 masked image and fp32 {0,1} mask (evaluate.py:28-33) ----
__global__ void u8_prepare_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ msk, long n,
                                  float* __restrict__ masked, float* __restrict__ mask_out) {
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float v = static_cast<float>(img[i]) / 255.f;      // ToTensor: .div(255) in fp32
    const float m = msk[i] > 0 ? 1.f : 0.f;                   // (mask > 0).float()
    masked[i] = v * m;
    mask_out[i] = m;
  }
}

// ---- one pass of Pillow's resample along the last (kAxis = 1) or the row (kAxis = 0) axis ----
// src: uint8, or fp32 quantised on the fly as (x * 255).astype(uint8) (evaluate.py:54-55) when kFloatSrc.
template <bool kFloatSrc>
__device__ __forceinline__ int load_px(const void* src, long i) {
  if (kFloatSrc) return static_cast<int>(static_cast<uint8_t>(reinterpret_cast<const float*>(src)[i] * 255.f));
  return reinterpret_cast<const uint8_t*>(src)[i];
}

template <bool kFloatSrc>
__global__ void resample_h_kernel(const void* __restrict__ src, int B, int H, int Win, int Wout,
                                  const int32_t* __restrict__ bounds, const int32_t* __restrict__ kk, int ksize,
                                  uint8_t* __restrict__ dst) {
  const long total = static_cast<long>(B) * H * Wout;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int xx = static_cast<int>(i % Wout);
    const long row = i / Wout;
    const int xmin = bounds[2 * xx], xmax = bounds[2 * xx + 1];
    int ss = 1 << (kPrecisionBits - 1);
    for (int x = 0; x < xmax; ++x) ss += load_px<kFloatSrc>(src, row * Win + xmin + x) * kk[xx * ksize + x];
    ss >>= kPrecisionBits;
    dst[i] = static_cast<uint8_t>(ss < 0 ? 0 : (ss > 255 ? 255 : ss));
  }
}

__global__ void resample_v_kernel(const uint8_t* __restrict__ src, int B, int Hin, int Hout, int W,
                                  const int32_t* __restrict__ bounds, const int32_t* __restrict__ kk, int ksize,
                                  uint8_t* __restrict__ dst) {
  const long total = static_cast<long>(B) * Hout * W;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % W);
    const int yy = static_cast<int>((i / W) % Hout);
    const long b = i / (static_cast<long>(W) * Hout);
    const int ymin = bounds[2 * yy], ymax = bounds[2 * yy + 1];
    int ss = 1 << (kPrecisionBits - 1);
    for (int y = 0; y < ymax; ++y) ss += src[(b * Hin + ymin + y) * W + x] * kk[yy * ksize + y];
    ss >>= kPrecisionBits;
    dst[i] = static_cast<uint8_t>(ss < 0 ? 0 : (ss > 255 ? 255 : ss));
  }
}

__global__ void quantize_u8_kernel(const float* __restrict__ src, long n, uint8_t* __restrict__ dst) {
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long>(gridDim.x) * blockDim.x)
    dst[i] = static_cast<uint8_t>(src[i] * 255.f);
}

// ---- DSM normalisation (data_extraction.py:80-103): per-tile NaN-aware min / max, then 255 (d - min) / (max - min) ----
__global__ void __launch_bounds__(256)
dsm_minmax_kernel(const double* __restrict__ data, long per_tile, double* __restrict__ mm /*[B][2]*/) {
  __shared__ double s_min[8], s_max[8];
  const double* d = data + static_cast<long>(blockIdx.x) * per_tile;
  double lo = INFINITY, hi = -INFINITY;
  for (long i = threadIdx.x; i < per_tile; i += 256) {
    const double v = d[i];
    if (v == v) { lo = fmin(lo, v); hi = fmax(hi, v); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { s_min[threadIdx.x >> 5] = lo; s_max[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 1; q < 8; ++q) { lo = fmin(lo, s_min[q]); hi = fmax(hi, s_max[q]); }
    mm[2 * blockIdx.x] = lo;
    mm[2 * blockIdx.x + 1] = hi;
  }
}

__global__ void dsm_normalize_kernel(const double* __restrict__ data, long per_tile, int B, const double* __restrict__ mm,
                                     uint8_t* __restrict__ out) {
  const long total = per_tile * B;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long b = i / per_tile;
    const double lo = mm[2 * b], hi = mm[2 * b + 1];
    const double v = data[i];
    uint8_t o = 0;
    if (v == v && hi > lo) o = static_cast<uint8_t>(255 * (v - lo) / (hi - lo));   // NaN -> 0; flat / empty tile -> 0
    out[i] = o;
  }
}

static int io_grid(long n) {
  long g = (n + 255) / 256;
  const long cap = static_cast<long>(num_sms() > 0 ? num_sms() : 148) * 8;
  if (g > cap) g = cap;
  return static_cast<int>(g < 1 ? 1 : g);
}

}  // namespace tg

extern "C" int tg_resize_ksize(int in_size, int out_size) {
  if (in_size <= 0 || out_size <= 0) return -1;
  double scale = static_cast<double>(in_size) / out_size;
  double filterscale = scale < 1.0 ? 1.0 : scale;
  return static_cast<int>(ceil(1.0 * filterscale)) * 2 + 1;
}

// HOST tables of Pillow's precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter:
// bounds[out_size][2] = (xmin, xmax count), kk[out_size][ksize] 22-bit fixed-point weights.
extern "C" int tg_resize_coeffs(int in_size, int out_size, int32_t* bounds, int32_t* kk, int ksize) {
  using namespace tg;
  TG_REQUIRE(bounds && kk && in_size > 0 && out_size > 0 && ksize == tg_resize_ksize(in_size, out_size),
             "tg_resize_coeffs: bad arguments");
  const double scale = static_cast<double>(in_size) / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * filterscale, ss = 1.0 / filterscale;
  double w[64];
  TG_REQUIRE(ksize <= 64, "tg_resize_coeffs: scale factor too large (ksize %d)", ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      double a = (x + xmin - center + 0.5) * ss;
      if (a < 0) a = -a;
      w[x] = a < 1.0 ? 1.0 - a : 0.0;
      ww += w[x];
    }
    for (int x = 0; x < xmax; ++x)
      if (ww != 0.0) w[x] /= ww;
    for (int x = xmax; x < ksize; ++x) w[x] = 0.0;
    for (int x = 0; x < ksize; ++x) {
      const double v = w[x] * (1 << kPrecisionBits);
      kk[xx * ksize + x] = w[x] < 0 ? static_cast<int32_t>(v - 0.5) : static_cast<int32_t>(v + 0.5);
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
  return 0;
}

extern "C" int tg_u8_prepare(const uint8_t* img, const uint8_t* mask, long n, float* masked, float* mask_out,
                             void* stream) {
  using namespace tg;
  TG_REQUIRE(img && mask && masked && mask_out && n > 0, "tg_u8_prepare: bad arguments");
  u8_prepare_kernel<<<io_grid(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(img, mask, n, masked, mask_out);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tg_resize_bilinear_u8(const void* src, int src_is_f32, int B, int Hin, int Win, int Hout, int Wout,
                                     const int32_t* bounds_w, const int32_t* kk_w, int ksize_w,
                                     const int32_t* bounds_h, const int32_t* kk_h, int ksize_h, uint8_t* tmp,
                                     uint8_t* dst, void* stream) {
  using namespace tg;
  TG_REQUIRE(src && dst && tmp && bounds_w && kk_w && bounds_h && kk_h && B > 0, "tg_resize_bilinear_u8: null pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // horizontal pass (always run: it also performs the fp32 -> uint8 quantisation), then vertical
  const long nh = static_cast<long>(B) * Hin * Wout;
  uint8_t* hdst = (Hin == Hout) ? dst : tmp;
  if (src_is_f32) resample_h_kernel<true><<<io_grid(nh), 256, 0, st>>>(src, B, Hin, Win, Wout, bounds_w, kk_w, ksize_w, hdst);
  else resample_h_kernel<false><<<io_grid(nh), 256, 0, st>>>(src, B, Hin, Win, Wout, bounds_w, kk_w, ksize_w, hdst);
  TG_CHECK_CUDA(cudaGetLastError());
  if (Hin != Hout) {
    resample_v_kernel<<<io_grid(static_cast<long>(B) * Hout * Wout), 256, 0, st>>>(tmp, B, Hin, Hout, Wout, bounds_h, kk_h,
                                                                                  ksize_h, dst);
    TG_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

extern "C" int tg_quantize_u8(const float* src, long n, uint8_t* dst, void* stream) {
  using namespace tg;
  TG_REQUIRE(src && dst && n > 0, "tg_quantize_u8: bad arguments");
  quantize_u8_kernel<<<io_grid(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, n, dst);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tg_dsm_normalize(const double* data, int B, int H, int W, double* minmax, uint8_t* out, void* stream) {
  using namespace tg;
  TG_REQUIRE(data && minmax && out && B > 0 && H > 0 && W > 0, "tg_dsm_normalize: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long per_tile = static_cast<long>(H) * W;
  dsm_minmax_kernel<<<B, 256, 0, st>>>(data, per_tile, minmax);
  TG_CHECK_CUDA(cudaGetLastError());
  dsm_normalize_kernel<<<io_grid(per_tile * B), 256, 0, st>>>(data, per_tile, B, minmax, out);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}
