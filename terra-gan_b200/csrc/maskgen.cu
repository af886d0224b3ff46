// maskgen.cu — the dense image operations of the reference's synthetic irregular-mask generator on the device
// (SURVEY.md §8f rank 4): /root/reference/random__annotation_mask_generator.py:33-148 builds each hole mask from
// scipy.ndimage calls on 500x500 float64 / bool arrays (binary_dilation / erosion / opening / closing with the default
// cross structuring element, gaussian_filter with sigma up to 30, thresholds and distance tests). The random DRAWS
// stay on the host in the reference's order (tg_b200/maskgen.py), so a seeded run reproduces the reference's masks
// bit for bit; everything per-pixel happens here, in float64 with scipy's operation order (no FMA contraction:
// explicit __dmul_rn / __dadd_rn), so thresholds fall on the same side.
#include "tg_common.cuh"
#include "../../include/terragan_b200.h"

namespace tg {

// scipy.ndimage binary_dilation / binary_erosion, structure = cross, border_value = 0, one iteration
__global__ void morph_cross_kernel(const uint8_t* __restrict__ in, int H, int W, int erode, uint8_t* __restrict__ out) {
  const int n = H * W;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int r = i / W, c = i % W;
    const int up = r > 0 ? in[i - W] != 0 : 0, dn = r < H - 1 ? in[i + W] != 0 : 0;
    const int lf = c > 0 ? in[i - 1] != 0 : 0, rt = c < W - 1 ? in[i + 1] != 0 : 0;
    const int me = in[i] != 0;
    out[i] = erode ? (me & up & dn & lf & rt) : (me | up | dn | lf | rt);
  }
}

// NI_EXTEND_REFLECT: d c b a | a b c d | d c b a
__device__ __forceinline__ int reflect_index(int p, int n) {
  while (p < 0 || p >= n) p = p < 0 ? -p - 1 : 2 * n - 1 - p;
  return p;
}

// scipy.ndimage.correlate1d with a symmetric odd kernel along `axis` (0: rows, 1: columns of the [H][W] array):
//   tmp = x[i] * w[center]; for jj = -radius .. -1: tmp += (x[i + jj] + x[i - jj]) * w[center + jj]
__global__ void gauss1d_f64_kernel(const double* __restrict__ in, int H, int W, int axis, const double* __restrict__ w,
                                   int radius, double* __restrict__ out) {
  const int n = H * W;
  const int len = axis == 0 ? H : W;
  const int stride = axis == 0 ? W : 1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int r = i / W, c = i % W;
    const int pos = axis == 0 ? r : c;
    const double* line = in + (axis == 0 ? c : r * W);
    double tmp = __dmul_rn(line[pos * stride], w[radius]);
    for (int jj = -radius; jj < 0; ++jj) {
      const double a = line[reflect_index(pos + jj, len) * stride];
      const double b = line[reflect_index(pos - jj, len) * stride];
      tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(a, b), w[radius + jj]));
    }
    out[i] = tmp;
  }
}

// mode 0: out = base > p0                                                   (edge approach, :75-76)
// mode 1: out |= sqrt(x^2 + y^2) <= p0 + field * p1                          (patch, :84-97; p0 radius, p1 amplitude)
// mode 2: out |= x^2 / p0^2 + y^2 / p1^2 <= 1                                (elliptical region, :113-117; p0 a, p1 b)
// mode 3: out |= (x^2 + y^2 <= p0^2) & (field > p1)                          (irregular region, :118-129)
// with x = col - cx, y = row - cy (np.ogrid[-cy:size-cy, -cx:size-cx])
__global__ void mask_shape_kernel(const double* __restrict__ field, int H, int W, int mode, int cx, int cy, double p0,
                                  double p1, uint8_t* __restrict__ out) {
  const int n = H * W;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const long x = (i % W) - cx, y = (i / W) - cy;
    bool on;
    if (mode == 0) {
      on = field[i] > p0;
    } else if (mode == 1) {
      const double dist = sqrt(static_cast<double>(x * x + y * y));
      on = dist <= __dadd_rn(p0, __dmul_rn(field[i], p1));
    } else if (mode == 2) {
      const long a2 = static_cast<long>(p0) * static_cast<long>(p0), b2 = static_cast<long>(p1) * static_cast<long>(p1);
      on = __dadd_rn(static_cast<double>(x * x) / static_cast<double>(a2), static_cast<double>(y * y) / static_cast<double>(b2)) <= 1.0;
    } else {
      const long m2 = static_cast<long>(p0) * static_cast<long>(p0);
      on = (x * x + y * y <= m2) && (field[i] > p1);
    }
    out[i] = mode == 0 ? (on ? 1 : 0) : (out[i] | (on ? 1 : 0));
  }
}

static int mg_grid(long n) {
  long g = (n + 255) / 256;
  const long cap = static_cast<long>(num_sms() > 0 ? num_sms() : 148) * 8;
  return static_cast<int>(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace tg

extern "C" int tg_morph_cross(const uint8_t* in, int H, int W, int erode, uint8_t* out, void* stream) {
  using namespace tg;
  TG_REQUIRE(in && out && in != out && H > 0 && W > 0, "tg_morph_cross: bad arguments (in-place is not supported)");
  morph_cross_kernel<<<mg_grid(static_cast<long>(H) * W), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, H, W, erode, out);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tg_gauss1d_f64(const double* in, int H, int W, int axis, const double* weights, int radius, double* out,
                              void* stream) {
  using namespace tg;
  TG_REQUIRE(in && out && in != out && weights && radius >= 0 && (axis == 0 || axis == 1), "tg_gauss1d_f64: bad arguments");
  gauss1d_f64_kernel<<<mg_grid(static_cast<long>(H) * W), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, H, W, axis,
                                                                                                         weights, radius, out);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tg_mask_shape(const double* field, int H, int W, int mode, int cx, int cy, double p0, double p1,
                             uint8_t* out, void* stream) {
  using namespace tg;
  TG_REQUIRE(out && mode >= 0 && mode <= 3 && (field || mode == 2), "tg_mask_shape: bad arguments");
  mask_shape_kernel<<<mg_grid(static_cast<long>(H) * W), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(field, H, W, mode, cx,
                                                                                                        cy, p0, p1, out);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}
