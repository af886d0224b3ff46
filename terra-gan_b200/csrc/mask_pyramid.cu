// mask_pyramid.cu — integer mask pyramid of the PConv U-Net.
//
// In the reference every PConv2d recomputes, in fp32 and twice per call, a 1->1 all-ones convolution
// of the single-channel mask (mvp_gan/src/models/pconv.py:33-40) to get the window count s, the
// updated mask [s > 0] and the ratio k^2/s; the decoder merges masks with
// max(nearest_up2(up_mask), skip_mask) (mvp_gan/src/models/generator.py:50-54,68-74).
// All of these depend only on the input mask, never on features, so they are computed here once per
// batch as uint8 integers (bit-exact by construction: counts <= 49 are exact in fp32 too) and
// shared by fprop, dgrad and wgrad. HBM-bound byte work: one thread per output pixel, coalesced.
#include <cstdlib>
#include "tg_common.cuh"
#include "../../include/terragan_b200.h"

namespace tg {

__global__ void mask_window_sum_kernel(const uint8_t* __restrict__ m, int B, int Hi, int Wi, int k,
                                       int s, int pad, int Ho, int Wo, uint8_t* __restrict__ sum,
                                       uint8_t* __restrict__ upd, uint8_t* __restrict__ upd_split) {
  const long total = static_cast<long>(B) * Ho * Wo;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int wo = static_cast<int>(i % Wo);
    const int ho = static_cast<int>((i / Wo) % Ho);
    const int b = static_cast<int>(i / (static_cast<long>(Wo) * Ho));
    const uint8_t* mb = m + static_cast<long>(b) * Hi * Wi;
    int acc = 0;
    for (int kh = 0; kh < k; ++kh) {
      const int h = ho * s + kh - pad;
      if (h < 0 || h >= Hi) continue;
      for (int kw = 0; kw < k; ++kw) {
        const int w = wo * s + kw - pad;
        if (w < 0 || w >= Wi) continue;
        acc += mb[h * Wi + w] != 0;
      }
    }
    if (sum) sum[i] = static_cast<uint8_t>(acc);
    const uint8_t u = acc > 0;
    if (upd) upd[i] = u;
    if (upd_split) {
      const int H2 = Ho >> 1, W2 = Wo >> 1;
      const long o = ((static_cast<long>(b) * 4 + 2 * (ho & 1) + (wo & 1)) * H2 + (ho >> 1)) * W2 + (wo >> 1);
      upd_split[o] = u;
    }
  }
}

// Fast path of the window sum for the shapes of the U-Net (odd k <= 7, pad = k/2, stride 1 or 2, widths divisible by 4):
// one thread = FOUR consecutive outputs of a row, arithmetic on bytes packed in 32-bit words (counts <= 49 never carry
// into the neighbouring byte). The k input rows are added word-wise (vertical sums of the aligned words i-1, i, i+1[, i+2]),
// the horizontal window is a sum of byte-shifted copies (funnel shifts across the word boundary), stride 2 keeps the even
// bytes. k = 7 / s = 2 reads 28 words per four outputs instead of 196 bytes; the generic kernel above took 147 us (enc1) and
// 205 us (dec1) per step for 17 MB of masks.
template <int K, int S>
__global__ void __launch_bounds__(256)
mask_window_sum4_kernel(const uint8_t* __restrict__ m, int B, int Hi, int Wi, int Ho, int Wo, uint8_t* __restrict__ sum,
                        uint8_t* __restrict__ upd, uint8_t* __restrict__ upd_split) {
  constexpr int P = K / 2;
  constexpr int NW = S;                      // stride-1 result words per thread (4 or 8 positions)
  const unsigned wq = static_cast<unsigned>(Wo) >> 2;
  const unsigned total = static_cast<unsigned>(B) * Ho * wq;
  const int wi_words = Wi >> 2;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned q = i % wq, r = i / wq;
    const int ho = static_cast<int>(r % Ho), b = static_cast<int>(r / Ho);
    const int iw0 = static_cast<int>(q) * NW - 1;
    const uint32_t* mb = reinterpret_cast<const uint32_t*>(m + static_cast<size_t>(b) * Hi * Wi);
    uint32_t V[NW + 2];
#pragma unroll
    for (int j = 0; j < NW + 2; ++j) V[j] = 0u;
#pragma unroll
    for (int kh = 0; kh < K; ++kh) {
      const int h = ho * S + kh - P;
      if (static_cast<unsigned>(h) < static_cast<unsigned>(Hi)) {
        const uint32_t* row = mb + static_cast<size_t>(h) * wi_words;
#pragma unroll
        for (int j = 0; j < NW + 2; ++j) {
          const int iw = iw0 + j;
          if (static_cast<unsigned>(iw) < static_cast<unsigned>(wi_words)) V[j] += __vcmpne4(__ldg(row + iw), 0u) & 0x01010101u;
        }
      }
    }
    uint32_t R[NW];
#pragma unroll
    for (int n = 0; n < NW; ++n) {
      const uint32_t l = V[n], c = V[n + 1], rr = V[n + 2];
      uint32_t acc = c;
#pragma unroll
      for (int d = 1; d <= P; ++d) {
        acc += __funnelshift_l(l, c, 8 * d);     // byte lane j <- position x - d
        acc += __funnelshift_r(c, rr, 8 * d);    // byte lane j <- position x + d
      }
      R[n] = acc;
    }
    const uint32_t out = S == 1 ? R[0] : __byte_perm(R[0], R[NW - 1], 0x6420);   // stride 2: the even positions
    const uint32_t u = __vcmpne4(out, 0u) & 0x01010101u;
    if (sum) reinterpret_cast<uint32_t*>(sum)[i] = out;
    if (upd) reinterpret_cast<uint32_t*>(upd)[i] = u;
    if (upd_split) {
      const int H2 = Ho >> 1, W2 = Wo >> 1;
      const size_t o = ((static_cast<size_t>(b) * 4 + 2 * (ho & 1)) * H2 + (ho >> 1)) * W2 + 2 * q;
      const size_t plane = static_cast<size_t>(H2) * W2;
      *reinterpret_cast<uint16_t*>(upd_split + o) = static_cast<uint16_t>((u & 0xffu) | ((u >> 8) & 0xff00u));              // wo even
      *reinterpret_cast<uint16_t*>(upd_split + o + plane) = static_cast<uint16_t>(((u >> 8) & 0xffu) | ((u >> 16) & 0xff00u));   // wo odd
    }
  }
}

__global__ void mask_split_kernel(const uint8_t* __restrict__ m, int B, int H, int W,
                                  uint8_t* __restrict__ out) {
  const long total = static_cast<long>(B) * H * W;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int w = static_cast<int>(i % W);
    const int h = static_cast<int>((i / W) % H);
    const int b = static_cast<int>(i / (static_cast<long>(W) * H));
    const int H2 = H >> 1, W2 = W >> 1;
    const long o = ((static_cast<long>(b) * 4 + 2 * (h & 1) + (w & 1)) * H2 + (h >> 1)) * W2 + (w >> 1);
    out[o] = m[i] != 0;
  }
}

__global__ void mask_merge_up_kernel(const uint8_t* __restrict__ up, const uint8_t* __restrict__ skip,
                                     int B, int H, int W, uint8_t* __restrict__ out) {
  const long total = static_cast<long>(B) * H * W;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int w = static_cast<int>(i % W);
    const int h = static_cast<int>((i / W) % H);
    const int b = static_cast<int>(i / (static_cast<long>(W) * H));
    const uint8_t u = up[(static_cast<long>(b) * (H >> 1) + (h >> 1)) * (W >> 1) + (w >> 1)];
    out[i] = (u | skip[i]) != 0;
  }
}

__global__ void mask_from_f32_kernel(const float* __restrict__ m, long n, uint8_t* __restrict__ out) {
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long>(gridDim.x) * blockDim.x)
    out[i] = m[i] > 0.f;
}
__global__ void mask_to_f32_kernel(const uint8_t* __restrict__ m, long n, float* __restrict__ out) {
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long>(gridDim.x) * blockDim.x)
    out[i] = m[i] ? 1.f : 0.f;
}

static int grid_for(long n, int block) {
  long g = (n + block - 1) / block;
  const long cap = static_cast<long>(num_sms() > 0 ? num_sms() : 148) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace tg

extern "C" int tg_mask_window_sum(const uint8_t* mask_in, int B, int Hi, int Wi, int k, int s, int pad,
                                  uint8_t* sum, uint8_t* upd, uint8_t* upd_split, uint8_t* in_split,
                                  void* stream) {
  using namespace tg;
  TG_REQUIRE(mask_in != nullptr && B > 0 && Hi > 0 && Wi > 0, "tg_mask_window_sum: bad input");
  TG_REQUIRE(k >= 1 && k <= 7 && s >= 1 && pad >= 0, "tg_mask_window_sum: bad window k=%d s=%d pad=%d", k, s, pad);
  const int Ho = (Hi + 2 * pad - k) / s + 1;
  const int Wo = (Wi + 2 * pad - k) / s + 1;
  TG_REQUIRE(Ho > 0 && Wo > 0, "tg_mask_window_sum: empty output");
  TG_REQUIRE(upd_split == nullptr || (Ho % 2 == 0 && Wo % 2 == 0),
             "tg_mask_window_sum: parity-split output needs even Ho, Wo (got %d x %d)", Ho, Wo);
  TG_REQUIRE(in_split == nullptr || (Hi % 2 == 0 && Wi % 2 == 0),
             "tg_mask_window_sum: parity-split input copy needs even Hi, Wi");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long total = static_cast<long>(B) * Ho * Wo;
  static const bool fast_on = [] { const char* e = getenv("TG_NO_MASK_FAST"); return !(e != nullptr && e[0] == '1'); }();
  auto al = [](const void* p, uintptr_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; };
  const bool fast = fast_on && (k == 3 || k == 5 || k == 7) && pad == k / 2 && (s == 1 || (s == 2 && k >= 3 && Hi % 2 == 0)) &&
                    Wi % 4 == 0 && Wo % 4 == 0 && Wo * s == Wi && Ho * s == Hi && al(mask_in, 4) && al(sum, 4) && al(upd, 4) &&
                    al(upd_split, 2) && total / 4 < (1L << 31);
#define TG_MASK_FAST(K, S)                                                                                              \
  mask_window_sum4_kernel<K, S><<<grid_for(total / 4, 256), 256, 0, st>>>(mask_in, B, Hi, Wi, Ho, Wo, sum, upd, upd_split)
  if (fast && k == 7 && s == 2) TG_MASK_FAST(7, 2);
  else if (fast && k == 5 && s == 2) TG_MASK_FAST(5, 2);
  else if (fast && k == 3 && s == 2) TG_MASK_FAST(3, 2);
  else if (fast && k == 3 && s == 1) TG_MASK_FAST(3, 1);
  else
    mask_window_sum_kernel<<<grid_for(total, 256), 256, 0, st>>>(mask_in, B, Hi, Wi, k, s, pad, Ho, Wo, sum, upd, upd_split);
#undef TG_MASK_FAST
  TG_CHECK_CUDA(cudaGetLastError());
  if (in_split != nullptr) {
    const long tin = static_cast<long>(B) * Hi * Wi;
    mask_split_kernel<<<grid_for(tin, 256), 256, 0, st>>>(mask_in, B, Hi, Wi, in_split);
    TG_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

extern "C" int tg_mask_merge_up(const uint8_t* up, const uint8_t* skip, int B, int H, int W, uint8_t* out,
                                void* stream) {
  using namespace tg;
  TG_REQUIRE(up && skip && out, "tg_mask_merge_up: null pointer");
  TG_REQUIRE(H % 2 == 0 && W % 2 == 0, "tg_mask_merge_up: H, W must be even");
  const long total = static_cast<long>(B) * H * W;
  mask_merge_up_kernel<<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(up, skip, B, H,
                                                                                                 W, out);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tg_mask_from_f32(const float* mask, long n, uint8_t* out, void* stream) {
  using namespace tg;
  TG_REQUIRE(mask && out && n > 0, "tg_mask_from_f32: bad arguments");
  mask_from_f32_kernel<<<grid_for(n, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(mask, n, out);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tg_mask_to_f32(const uint8_t* mask, long n, float* out, void* stream) {
  using namespace tg;
  TG_REQUIRE(mask && out && n > 0, "tg_mask_to_f32: bad arguments");
  mask_to_f32_kernel<<<grid_for(n, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(mask, n, out);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}
