// bn_kernels.cu — BatchNorm2d (train-mode statistics) + activation, forward and backward, as
// coalesced 16-byte-vectorised HBM-bound kernels over channels-last bf16 activations.
//
// Reference ops replaced: `self.bn(output)` + `self.activation(output)` in PConv2d.forward
// (mvp_gan/src/models/pconv.py:46-48; nn.BatchNorm2d eps 1e-5, momentum 0.1 + ReLU) and
// BatchNorm2d + LeakyReLU(0.2) in the Discriminator blocks (mvp_gan/src/models/discriminator.py:12-14),
// plus their autograd backward together with the `output * mask_ratio` factor (pconv.py:43).
//
// Forward:  the conv epilogue already produced per-CTA (sum, sumsq) partials of z; tg_bn_finalize
//           turns them into scale = gamma*invstd, shift = beta - mean*scale (and the running-stat
//           EMA); tg_bn_apply writes y = act(z*scale + shift) in the layouts the consumers need
//           (plain NHWC and/or parity-split, optionally multiplied by the layer's output mask so
//           the next PConv sees x*mask without another pass).
// Backward: tg_bn_bwd_reduce accumulates the five per-channel sums BN backward needs in one pass
//           over (g, z); tg_bn_bwd_finalize produces dgamma, dbeta, the conv-bias gradient and the
//           per-channel coefficients; tg_bn_bwd_apply writes gz = dL/d(conv+bias) (bf16), which
//           the dgrad / wgrad tensor-core kernels consume.
#include <stdlib.h>

#include "tg_common.cuh"
#include "../../include/terragan_b200.h"

namespace tg {

struct __align__(16) bf16x8 {
  __nv_bfloat162 v[4];
};

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) { vload8(p, f); }
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 raw;
  raw.x = pack_bf16x2(f[0], f[1]);
  raw.y = pack_bf16x2(f[2], f[3]);
  raw.z = pack_bf16x2(f[4], f[5]);
  raw.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = raw;
}
__device__ __forceinline__ void load8(const float* p, float (&f)[8]) { vload8(p, f); }
__device__ __forceinline__ void store8(float* p, const float (&f)[8]) { vstore8(p, f); }
__device__ __forceinline__ void ldg8f(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// pixel index of (b, h, w) in parity-split order [B][4][H/2][W/2]
__device__ __forceinline__ long split_index(int b, int h, int w, int H, int W) {
  return ((static_cast<long>(b) * 4 + 2 * (h & 1) + (w & 1)) * (H >> 1) + (h >> 1)) * (W >> 1) + (w >> 1);
}

// (b, h, w) of the pixels first, first + stride, first + 2 stride, ... without per-pixel integer divisions: the stride is
// decomposed once into (db, dh, dw) and added with carries. (The two runtime divisions per pixel of the parity-split
// index, plus 64-bit multiply-adds for three addresses, were ~100 of the ~280 instructions these issue-bound kernels
// spent per 8-channel vector.)
struct PixelWalk {
  int b, h, w, db, dh, dw, H, W;
  __device__ __forceinline__ PixelWalk(unsigned first, unsigned stride, int H_, int W_) : H(H_), W(W_) {
    const unsigned HW = static_cast<unsigned>(H_) * W_;
    b = static_cast<int>(first / HW);
    unsigned rem = first - static_cast<unsigned>(b) * HW;
    h = static_cast<int>(rem / W_);
    w = static_cast<int>(rem - static_cast<unsigned>(h) * W_);
    db = static_cast<int>(stride / HW);
    rem = stride - static_cast<unsigned>(db) * HW;
    dh = static_cast<int>(rem / W_);
    dw = static_cast<int>(rem - static_cast<unsigned>(dh) * W_);
  }
  __device__ __forceinline__ void next() {
    w += dw;
    int c = w >= W ? 1 : 0;
    w -= c ? W : 0;
    h += dh + c;
    c = h >= H ? 1 : 0;
    h -= c ? H : 0;
    b += db + c;
  }
  // parity-split pixel index, 32-bit (the callers guarantee fewer than 2^31 pixels)
  __device__ __forceinline__ unsigned split() const {
    const unsigned plane = static_cast<unsigned>(b) * 4u + 2u * (h & 1) + (w & 1);
    return (plane * static_cast<unsigned>(H >> 1) + static_cast<unsigned>(h >> 1)) * static_cast<unsigned>(W >> 1) +
           static_cast<unsigned>(w >> 1);
  }
};

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
// Block = 32 channels x 8 row lanes: the per-CTA partial rows are summed in fp64 (8-way parallel over
// rows, coalesced over channels), then the usual BN bookkeeping.
__global__ void __launch_bounds__(256)
bn_finalize_kernel(const float* __restrict__ partial, int rows, int C, double count,
                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float momentum,
                   float* __restrict__ running_mean, float* __restrict__ running_var, float* __restrict__ scale,
                   float* __restrict__ shift, float* __restrict__ mean_out, float* __restrict__ invstd_out) {
  __shared__ double s_sum[8][32][2];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  double s = 0.0, s2 = 0.0;
  if (c < C) {
    for (int r = rl; r < rows; r += 8) {
      s += partial[(static_cast<long>(r) * 2) * C + c];
      s2 += partial[(static_cast<long>(r) * 2 + 1) * C + c];
    }
  }
  s_sum[rl][cl][0] = s;
  s_sum[rl][cl][1] = s2;
  __syncthreads();
  if (rl != 0 || c >= C) return;
#pragma unroll
  for (int q = 1; q < 8; ++q) {
    s += s_sum[q][cl][0];
    s2 += s_sum[q][cl][1];
  }
  const double mean = s / count;
  double var = s2 / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  const float g = gamma ? gamma[c] : 1.f;
  const float b = beta ? beta[c] : 0.f;
  const float sc = g * invstd;
  scale[c] = sc;
  shift[c] = b - static_cast<float>(mean) * sc;
  mean_out[c] = static_cast<float>(mean);
  invstd_out[c] = invstd;
  if (running_mean) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(mean);
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
  }
}

// eval mode: scale/shift from the running statistics
__global__ void bn_eval_coeff_kernel(int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                                     const float* __restrict__ running_mean,
                                     const float* __restrict__ running_var, float eps,
                                     float* __restrict__ scale, float* __restrict__ shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float invstd = rsqrtf(running_var[c] + eps);
  // rsqrtf is approximate: refine once so eval-mode outputs match 1/sqrt to fp32 rounding
  const float v = running_var[c] + eps;
  const float is = invstd * (1.5f - 0.5f * v * invstd * invstd);
  const float sc = (gamma ? gamma[c] : 1.f) * is;
  scale[c] = sc;
  shift[c] = (beta ? beta[c] : 0.f) - running_mean[c] * sc;
}

// Channel-stationary: a thread owns one 8-channel vector (scale/shift live in registers) and strides
// over pixels; 256 threads = (C/8) channel vectors x 256/(C/8) pixel lanes; two pixels in flight per
// iteration; 32-bit pixel arithmetic.
template <typename T>
__global__ void __launch_bounds__(256)
bn_apply_kernel(const T* __restrict__ z, unsigned M, int C, int H, int W,
                const float* __restrict__ scale, const float* __restrict__ shift, int act, float slope,
                const uint8_t* __restrict__ code, T* __restrict__ y_nhwc,
                T* __restrict__ y_split, int mask_split) {
  const unsigned cv = C >> 3;
  const unsigned lanes = blockDim.x / cv;
  const unsigned c = (threadIdx.x % cv) << 3;
  const unsigned lane = threadIdx.x / cv;
  float sc[8], sh[8];
  ldg8f(scale + c, sc);
  ldg8f(shift + c, sh);
  const unsigned stride = gridDim.x * lanes;
  const unsigned first = blockIdx.x * lanes + lane;
  // the parity-split destination of consecutive pixels of the sequence: walked, not divided out per pixel
  PixelWalk walk(y_split ? first : 0u, y_split ? stride : 0u, H, W);
  for (unsigned p0 = first; p0 < M; p0 += 2 * stride) {
    const unsigned p1 = p0 + stride;
    const bool two = p1 < M;
    float v0[8], v1[8];
    load8(z + static_cast<size_t>(p0) * C + c, v0);
    if (two) load8(z + static_cast<size_t>(p1) * C + c, v1);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !two) break;
      float (&v)[8] = u ? v1 : v0;
      const unsigned p = u ? p1 : p0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = v[j] * sc[j] + sh[j];
        if (act == 1) t = fmaxf(t, 0.f);
        else if (act == 2) t = t > 0.f ? t : t * slope;
        v[j] = t;
      }
      if (y_nhwc) store8(y_nhwc + static_cast<size_t>(p) * C + c, v);
      if (y_split) {
        if (mask_split && code[p] == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = 0.f;
        }
        store8(y_split + static_cast<size_t>(walk.split()) * C + c, v);
        walk.next();
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
template <typename T>
struct GradSrcT {
  const T* ptr;              // null = absent
  long pix_stride;           // elements between consecutive pixels
  int chan_off;              // first channel of this layer's slice inside the source tensor
  int split;                 // 1: source pixels are in parity-split order
};

using GradSrc = GradSrcT<__nv_bfloat16>;

// ---- per-thread cp.async ring ---------------------------------------------------------------
// The backward kernels stream two or three bf16 tensors with one 16-byte vector per thread and pixel. Held in
// registers, the loads in flight are bounded by the register file (measured: 2.1 TB/s for the reduce). Here each
// thread keeps kRing pixels in flight through cp.async into its own shared-memory slots (no cross-thread
// sharing, so cp.async.wait_group is the only synchronisation) and the registers only hold the pixel in use.

// ring depth per kernel and storage type. Measured on dec1 (1.07 G elements, B = 64): 4 slots 1124 / 1137 us (reduce /
// apply), 8 / 6 slots 1183 / 1291 us — the larger ring costs more in occupancy than it gains in bytes in flight.
template <typename T> constexpr int ring_reduce() { return 4; }
template <typename T> constexpr int ring_apply() { return 4; }
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// one ring slot = the 8-element vector of one thread: 16 bytes of bf16 or 32 bytes of fp32
template <typename T>
__device__ __forceinline__ void cp_async_vec8(uint32_t saddr, const T* g) {
  cp_async16(saddr, g);
  if (sizeof(T) == 4) cp_async16(saddr + 16, reinterpret_cast<const uint8_t*>(g) + 16);
}
__device__ __forceinline__ void unpack8(const __nv_bfloat16* slot, float (&f)[8]) { vload8(slot, f); }
__device__ __forceinline__ void unpack8(const float* slot, float (&f)[8]) { vload8(slot, f); }
// Streams (g0 [+ g1], z, ratio) of pixels p = first, first + stride, ... < M to `body(p, g[8], z[8], r)`.
// ring: kRing * 3 * blockDim.x uint4 of shared memory.
// MODE: the launch's feature set as a compile-time constant, so that the per-pixel instruction stream only holds what
// the layer uses (the kernels are issue-bound; null-pointer tests, the second source and both activation variants were
// evaluated per pixel). Bits: 1 second gradient source, 2 parity-split source(s), 4 ratio code, 8 LeakyReLU (else ReLU).
// MODE < 0: everything decided at run time (any other combination, and the fp32 verification path).
constexpr int kBwdS1 = 1, kBwdSplit = 2, kBwdCode = 4, kBwdLeaky = 8;
template <typename T, int kRing, int MODE, typename Body>
__device__ __forceinline__ void stream_grad_z(const GradSrcT<T>& s0, const GradSrcT<T>& s1_in, const T* __restrict__ z,
                                              unsigned M, int C, int H, int W, int c, unsigned first, unsigned stride,
                                              const uint8_t* __restrict__ code_in, const float* __restrict__ lut,
                                              uint4* ring_raw, Body body) {
  constexpr unsigned kSlot = 8 * sizeof(T);     // bytes per ring slot
  T* ring = reinterpret_cast<T*>(ring_raw);
  GradSrcT<T> s1 = s1_in;
  if (MODE >= 0 && !(MODE & kBwdS1)) s1.ptr = nullptr;
  const uint8_t* code = (MODE >= 0 && !(MODE & kBwdCode)) ? nullptr : code_in;
  const bool need_split = MODE >= 0 ? (MODE & kBwdSplit) != 0 : (s0.split || (s1.ptr && s1.split));
  const unsigned nthr = blockDim.x, tid = threadIdx.x;
  const uint32_t ring_s = smem_u32(ring);
  uint8_t codes[kRing];
  // fills are issued for consecutive pixels of the sequence first, first + stride, ...: the parity-split position is walked
  // with carries; plain addresses are one 32 x 32 -> 64-bit multiply-add from the pixel index (persistent 64-bit pointers
  // per stream cost the apply kernel its third CTA per SM in registers)
  unsigned p_iss = first;
  const unsigned s0_ps = static_cast<unsigned>(s0.pix_stride), s1_ps = static_cast<unsigned>(s1.pix_stride);   // channel counts
  const T* const s0b = s0.ptr + s0.chan_off + c;
  const T* const s1b = s1.ptr ? s1.ptr + s1.chan_off + c : nullptr;
  const T* const zb = z + c;
  PixelWalk walk(need_split ? first : 0u, need_split ? stride : 0u, H, W);
  auto issue = [&](int d) {
    if (p_iss < M) {
      const unsigned ps = need_split ? walk.split() : 0u;
      cp_async_vec8(ring_s + ((d * 3 + 0) * nthr + tid) * kSlot, s0b + static_cast<size_t>(s0.split ? ps : p_iss) * s0_ps);
      if (s1.ptr)
        cp_async_vec8(ring_s + ((d * 3 + 1) * nthr + tid) * kSlot, s1b + static_cast<size_t>(s1.split ? ps : p_iss) * s1_ps);
      cp_async_vec8(ring_s + ((d * 3 + 2) * nthr + tid) * kSlot, zb + static_cast<size_t>(p_iss) * static_cast<unsigned>(C));
      codes[d] = code ? __ldg(code + p_iss) : static_cast<uint8_t>(0);
    }
    cp_async_commit();
    p_iss += stride;
    if (need_split) walk.next();
  };
#pragma unroll
  for (int d = 0; d < kRing; ++d) issue(d);
  for (unsigned base = first; base < M; base += kRing * stride) {
#pragma unroll
    for (int d = 0; d < kRing; ++d) {
      const unsigned p = base + d * stride;
      if (p < M) {
        cp_async_wait<kRing - 1>();
        float g[8], zz[8];
        unpack8(ring + ((d * 3 + 0) * nthr + tid) * 8, g);
        if (s1.ptr) {
          float t[8];
          unpack8(ring + ((d * 3 + 1) * nthr + tid) * 8, t);
#pragma unroll
          for (int j = 0; j < 8; ++j) g[j] += t[j];
        }
        unpack8(ring + ((d * 3 + 2) * nthr + tid) * 8, zz);
        const float r = code ? __ldg(lut + codes[d]) : 1.f;
        issue(d);      // the slot has been read into registers: refill it with the next pixel of the sequence
        body(p, g, zz, r);
      }
    }
  }
  cp_async_wait<0>();
}

// sums per channel over all pixels: [0] g', [1] g'*z, [2] r*g', [3] r*z, [4] r   (g' = g * act'(z*scale+shift))
// block = 256 threads = (C/8 channel vectors) x (256/(C/8) pixel lanes); partial[block][5][C]
template <typename T, int kR, int MODE, int kMinB>
__global__ void __launch_bounds__(256, kMinB)      // specialised bf16 launches: three CTAs per SM (80 registers instead of 110)
bn_bwd_reduce_kernel(GradSrcT<T> s0, GradSrcT<T> s1, const T* __restrict__ z, long M,
                                     int C, int H, int W, const float* __restrict__ scale,
                                     const float* __restrict__ shift, int act, float slope,
                                     const uint8_t* __restrict__ code, const float* __restrict__ lut,
                                     float* __restrict__ partial) {
  extern __shared__ float red[];  // [lanes][5*8] per channel vector -> reduced below
  const int cv = C >> 3;
  const int lanes = blockDim.x / cv;
  const int my_cv = threadIdx.x % cv;
  const int my_lane = threadIdx.x / cv;
  const int c = my_cv << 3;
  float acc[5][8];
#pragma unroll
  for (int k = 0; k < 5; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[k][j] = 0.f;
  float sc[8], sh[8];
  ldg8f(scale + c, sc);
  ldg8f(shift + c, sh);
  if (MODE >= 0) act = (MODE & kBwdLeaky) ? 2 : 1;
  stream_grad_z<T, kR, MODE>(s0, s1, z, static_cast<unsigned>(M), C, H, W, c, blockIdx.x * lanes + my_lane, gridDim.x * lanes, code, lut,
                reinterpret_cast<uint4*>(red), [&](unsigned, const float (&g)[8], const float (&zz)[8], float r) {
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const float pre = zz[j] * sc[j] + sh[j];
                    float gg = g[j];
                    if (act == 1) gg = pre > 0.f ? gg : 0.f;
                    else if (act == 2) gg = pre > 0.f ? gg : gg * slope;
                    acc[0][j] += gg;
                    acc[1][j] += gg * zz[j];
                    acc[2][j] += r * gg;
                    acc[3][j] += r * zz[j];
                  }
                  acc[4][0] += r;
                });
#pragma unroll
  for (int j = 1; j < 8; ++j) acc[4][j] = acc[4][0];
  __syncthreads();      // every thread is done with its ring slots: the buffer becomes the reduction scratch
  // reduce over pixel lanes through shared memory
  float* mine = red + (static_cast<long>(my_lane) * cv + my_cv) * 40;
#pragma unroll
  for (int k = 0; k < 5; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) mine[k * 8 + j] = acc[k][j];
  __syncthreads();
  for (int i = threadIdx.x; i < cv * 40; i += blockDim.x) {
    const int v = i / 40, kj = i % 40;
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += red[(static_cast<long>(l) * cv + v) * 40 + kj];
    const int k = kj / 8, j = kj % 8;
    partial[(static_cast<long>(blockIdx.x) * 5 + k) * C + v * 8 + j] = s;
  }
}

// coeff[0]=scale [1]=mean [2]=invstd [3]=c1 (= dbeta/M) [4]=c2 (= dgamma/M), each [C]
// Block = 8 channels x 32 row lanes: the ~600 partial rows are summed 32-way in parallel in fp64 and combined in a fixed
// order (with 8 row lanes the 23 launches of a step cost 0.57 ms of pure latency).
__global__ void __launch_bounds__(256)
bn_bwd_finalize_kernel(const float* __restrict__ partial, int rows, int C, double count,
                       const float* __restrict__ scale, const float* __restrict__ mean,
                       const float* __restrict__ invstd, float* __restrict__ coeff, float* __restrict__ dgamma,
                       float* __restrict__ dbeta, float* __restrict__ dbias, int accumulate, int batch_stats) {
  __shared__ double s_sum[32][8][5];
  const int cl = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int c = blockIdx.x * 8 + cl;
  double s[5] = {0, 0, 0, 0, 0};
  if (c < C) {
    for (int r = rl; r < rows; r += 32)
#pragma unroll
      for (int k = 0; k < 5; ++k) s[k] += partial[(static_cast<long>(r) * 5 + k) * C + c];
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) s_sum[rl][cl][k] = s[k];
  __syncthreads();
  if (rl != 0 || c >= C) return;
  for (int q = 1; q < 32; ++q)
#pragma unroll
    for (int k = 0; k < 5; ++k) s[k] += s_sum[q][cl][k];
  const double mu = mean[c], is = invstd[c], sc = scale[c];
  const double db = s[0];                          // dL/dbeta
  const double dg = is * (s[1] - mu * s[0]);       // dL/dgamma = sum g' * zhat
  // eval-mode BatchNorm (running statistics) has no dependence of mean/var on the batch
  const double c1 = batch_stats ? db / count : 0.0, c2 = batch_stats ? dg / count : 0.0;
  coeff[c] = static_cast<float>(sc);
  coeff[C + c] = static_cast<float>(mu);
  coeff[2 * C + c] = static_cast<float>(is);
  coeff[3 * C + c] = static_cast<float>(c1);
  coeff[4 * C + c] = static_cast<float>(c2);
  // conv bias: out = (conv + b) * r  ->  db = sum_p r * dL/dz,  dL/dz = scale * (g' - c1 - zhat*c2)
  const double dbi = sc * (s[2] - c1 * s[4] - c2 * is * (s[3] - mu * s[4]));
  if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + static_cast<float>(dg);
  if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + static_cast<float>(db);
  if (dbias) dbias[c] = (accumulate ? dbias[c] : 0.f) + static_cast<float>(dbi);
}

// gz[p][c] = r[p] * scale[c] * (g' - c1[c] - zhat * c2[c]) = r[p] * (A[c]*g' + Bz[c]*z + Cc[c])
// Channel-stationary like bn_apply_kernel: the five per-channel coefficients live in registers.
template <typename T, int kR, int MODE>
__global__ void __launch_bounds__(256, sizeof(T) == 2 ? 3 : 1)      // bf16: three CTAs per SM (the bytes in flight are what hides HBM latency)
bn_bwd_apply_kernel(GradSrcT<T> s0, GradSrcT<T> s1, const T* __restrict__ z, unsigned M, int C, int H, int W,
                    const float* __restrict__ shift, const float* __restrict__ coeff, int act, float slope,
                    const uint8_t* __restrict__ code, const float* __restrict__ lut,
                    T* __restrict__ gz) {
  const unsigned cv = C >> 3;
  const unsigned lanes = blockDim.x / cv;
  const unsigned c = (threadIdx.x % cv) << 3;
  const unsigned lane = threadIdx.x / cv;
  float sc[8], sh[8], bz[8], cc[8];
  {
    float mu[8], is[8], c1[8], c2[8];
    ldg8f(coeff + c, sc);
    ldg8f(shift + c, sh);
    ldg8f(coeff + C + c, mu);
    ldg8f(coeff + 2 * C + c, is);
    ldg8f(coeff + 3 * C + c, c1);
    ldg8f(coeff + 4 * C + c, c2);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      bz[j] = -sc[j] * c2[j] * is[j];
      cc[j] = sc[j] * (mu[j] * is[j] * c2[j] - c1[j]);
    }
  }
  extern __shared__ uint4 apply_ring[];
  if (MODE >= 0) act = (MODE & kBwdLeaky) ? 2 : 1;
  stream_grad_z<T, kR, MODE>(s0, s1, z, M, C, H, W, static_cast<int>(c), blockIdx.x * lanes + lane, gridDim.x * lanes, code, lut, apply_ring,
                [&](unsigned p, const float (&g)[8], const float (&zz)[8], float r) {
                  float o[8];
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const float pre = zz[j] * sc[j] + sh[j];
                    float gg = g[j];
                    if (act == 1) gg = pre > 0.f ? gg : 0.f;
                    else if (act == 2) gg = pre > 0.f ? gg : gg * slope;
                    o[j] = r * (sc[j] * gg + bz[j] * zz[j] + cc[j]);
                  }
                  store8(gz + static_cast<size_t>(p) * C + c, o);
                });
}

// compile-time feature set of a backward launch (see stream_grad_z); -1 = decided at run time
template <typename T>
static int bwd_mode(const tg_grad_src* g0, const tg_grad_src* g1, int act, const uint8_t* code) {
  static const bool off = [] { const char* e = getenv("TG_BN_BWD_GENERIC"); return e != nullptr && e[0] == '1'; }();
  if (off || sizeof(T) != 2 || (act != 1 && act != 2)) return -1;
  const bool has1 = g1 != nullptr && g1->ptr != nullptr;
  const bool split = g0->split != 0 || (has1 && g1->split != 0);
  // (which of the sources is parity-split stays a run-time flag per source; the bit only says "walk the split position")
  const int m = (has1 ? kBwdS1 : 0) | (split ? kBwdSplit : 0) | (code ? kBwdCode : 0) | (act == 2 ? kBwdLeaky : 0);
  switch (m) {
    case kBwdCode:                                   // decoder layers, enc7
    case kBwdCode | kBwdS1 | kBwdSplit:              // enc1..enc6: skip gradient + parity-split gradient of the next layer
    case kBwdLeaky:                                  // Discriminator model[8]
    case kBwdLeaky | kBwdSplit:                      // Discriminator model[2], model[5]
      return m;
    default:
      return -1;
  }
}

template <typename T>
static GradSrcT<T> to_src(const tg_grad_src* s) {
  GradSrcT<T> r;
  if (s != nullptr && s->ptr != nullptr) {
    r.ptr = reinterpret_cast<const T*>(s->ptr);
    r.pix_stride = s->pix_stride;
    r.chan_off = s->chan_off;
    r.split = s->split;
  } else {
    r.ptr = nullptr;
    r.pix_stride = 0;
    r.chan_off = 0;
    r.split = 0;
  }
  return r;
}

template <typename T>
static int bn_apply_impl(const void* z, int B, int H, int W, int C, const float* scale, const float* shift,
                         int act, float slope, const uint8_t* code, void* y_nhwc, void* y_split,
                         int mask_split, void* stream) {
  TG_REQUIRE(z && scale && shift && C % 8 == 0, "tg_bn_apply: bad arguments (C=%d)", C);
  TG_REQUIRE(y_nhwc || y_split, "tg_bn_apply: no output requested");
  TG_REQUIRE(!y_split || (H % 2 == 0 && W % 2 == 0), "tg_bn_apply: parity-split output needs even H, W");
  TG_REQUIRE(!(mask_split && y_split) || code, "tg_bn_apply: mask_split needs code");
  const long M = static_cast<long>(B) * H * W;
  TG_REQUIRE(C / 8 <= 256 && 256 % (C / 8) == 0 && M < (1L << 31), "tg_bn_apply: unsupported C=%d or too many pixels", C);
  bn_apply_kernel<T><<<wave_grid(bn_apply_kernel<T>, 256, 0, (M * (C / 8) + 511) / 512), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const T*>(z), static_cast<unsigned>(M), C, H, W, scale, shift, act, slope, code,
      reinterpret_cast<T*>(y_nhwc), reinterpret_cast<T*>(y_split), mask_split);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <typename T>
static int bn_bwd_reduce_impl(const tg_grad_src* g0, const tg_grad_src* g1, const void* z, int B, int H,
                              int W, int C, const float* scale, const float* shift, int act, float slope,
                              const uint8_t* code, const float* lut_dev, float* partial, int rows_cap,
                              int* rows_used, void* stream) {
  TG_REQUIRE(g0 && g0->ptr && z && scale && shift && partial && rows_used, "tg_bn_bwd_reduce: null pointer");
  TG_REQUIRE(C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0, "tg_bn_bwd_reduce: unsupported C=%d", C);
  TG_REQUIRE(!code || lut_dev, "tg_bn_bwd_reduce: code needs a device LUT");
  const long M = static_cast<long>(B) * H * W;
  const int cv = C / 8, lanes = 256 / cv;
  long grid = (M + lanes - 1) / lanes;
  // one wave: every CTA resident from the start with the same share of the pixels. The specialised bf16 kernels are built
  // for three CTAs per SM; TG_BN_REDUCE_OCC=1 selects the earlier build (no register cap, 4 x SMs CTAs) for A/B timing.
  static const bool occ3 = [] { const char* e = getenv("TG_BN_REDUCE_OCC"); return !(e != nullptr && e[0] == '1'); }();
  const int mode = bwd_mode<T>(g0, g1, act, code);
  const bool three = occ3 && sizeof(T) == 2 && mode >= 0;
  const long cap = static_cast<long>(num_sms()) * (three ? 3 : 4);
  if (grid > cap) grid = cap;
  if (grid > rows_cap) grid = rows_cap;
  TG_REQUIRE(grid >= 1, "tg_bn_bwd_reduce: rows_cap must be >= 1");
  *rows_used = static_cast<int>(grid);
  size_t smem = static_cast<size_t>(lanes) * cv * 40 * sizeof(float);
  constexpr int kR = ring_reduce<T>();
  const size_t ring_bytes = static_cast<size_t>(kR) * 3 * 256 * 8 * sizeof(T);
  if (smem < ring_bytes) smem = ring_bytes;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define TG_LAUNCH_REDUCE(MODE)                                                                                       \
  {                                                                                                                  \
    constexpr int kOcc = (sizeof(T) == 2 && (MODE) >= 0) ? 3 : 1;                                                    \
    if (three) {                                                                                                     \
      TG_SET_SMEM_ONCE((bn_bwd_reduce_kernel<T, kR, MODE, kOcc>), 100 * 1024);                                       \
      bn_bwd_reduce_kernel<T, kR, MODE, kOcc><<<static_cast<int>(grid), 256, smem, st>>>(to_src<T>(g0), to_src<T>(g1), \
          reinterpret_cast<const T*>(z), M, C, H, W, scale, shift, act, slope, code, lut_dev, partial);              \
    } else {                                                                                                         \
      TG_SET_SMEM_ONCE((bn_bwd_reduce_kernel<T, kR, MODE, 1>), 100 * 1024);                                          \
      bn_bwd_reduce_kernel<T, kR, MODE, 1><<<static_cast<int>(grid), 256, smem, st>>>(to_src<T>(g0), to_src<T>(g1),  \
          reinterpret_cast<const T*>(z), M, C, H, W, scale, shift, act, slope, code, lut_dev, partial);              \
    }                                                                                                                \
  }
  switch (mode) {
    case kBwdCode: TG_LAUNCH_REDUCE(kBwdCode) break;
    case kBwdCode | kBwdS1 | kBwdSplit: TG_LAUNCH_REDUCE(kBwdCode | kBwdS1 | kBwdSplit) break;
    case kBwdLeaky: TG_LAUNCH_REDUCE(kBwdLeaky) break;
    case kBwdLeaky | kBwdSplit: TG_LAUNCH_REDUCE(kBwdLeaky | kBwdSplit) break;
    default: TG_LAUNCH_REDUCE(-1) break;
  }
#undef TG_LAUNCH_REDUCE
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <typename T>
static int bn_bwd_apply_impl(const tg_grad_src* g0, const tg_grad_src* g1, const void* z, int B, int H, int W,
                             int C, const float* shift, const float* coeff, int act, float slope,
                             const uint8_t* code, const float* lut_dev, void* gz, void* stream) {
  TG_REQUIRE(g0 && g0->ptr && z && shift && coeff && gz, "tg_bn_bwd_apply: null pointer");
  TG_REQUIRE(C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0, "tg_bn_bwd_apply: unsupported C=%d", C);
  TG_REQUIRE(!code || lut_dev, "tg_bn_bwd_apply: code needs a device LUT");
  const long M = static_cast<long>(B) * H * W;
  constexpr int kR = ring_apply<T>();
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define TG_LAUNCH_APPLY(MODE)                                                                                        \
  {                                                                                                                  \
    TG_SET_SMEM_ONCE((bn_bwd_apply_kernel<T, kR, MODE>), 100 * 1024);                                                \
    bn_bwd_apply_kernel<T, kR, MODE><<<wave_grid(bn_bwd_apply_kernel<T, kR, MODE>, 256, kR * 3 * 256 * 8 * sizeof(T), \
                                                 (M * (C / 8) + 255) / 256), 256, kR * 3 * 256 * 8 * sizeof(T), st>>>(  \
        to_src<T>(g0), to_src<T>(g1), reinterpret_cast<const T*>(z), static_cast<unsigned>(M), C, H, W, shift, coeff, act,  \
        slope, code, lut_dev, reinterpret_cast<T*>(gz));                                                             \
  }
  switch (bwd_mode<T>(g0, g1, act, code)) {
    case kBwdCode: TG_LAUNCH_APPLY(kBwdCode) break;
    case kBwdCode | kBwdS1 | kBwdSplit: TG_LAUNCH_APPLY(kBwdCode | kBwdS1 | kBwdSplit) break;
    case kBwdLeaky: TG_LAUNCH_APPLY(kBwdLeaky) break;
    case kBwdLeaky | kBwdSplit: TG_LAUNCH_APPLY(kBwdLeaky | kBwdSplit) break;
    default: TG_LAUNCH_APPLY(-1) break;
  }
#undef TG_LAUNCH_APPLY
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace tg

extern "C" int tg_bn_finalize(const float* partial, int rows, int C, double count, const float* gamma,
                              const float* beta, float eps, float momentum, float* running_mean,
                              float* running_var, float* scale, float* shift, float* mean, float* invstd,
                              void* stream) {
  using namespace tg;
  TG_REQUIRE(partial && scale && shift && mean && invstd && rows > 0 && C > 0 && count > 0,
             "tg_bn_finalize: bad arguments");
  bn_finalize_kernel<<<(C + 31) / 32, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      partial, rows, C, count, gamma, beta, eps, momentum, running_mean, running_var, scale, shift, mean, invstd);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tg_bn_eval_coeff(int C, const float* gamma, const float* beta, const float* running_mean,
                                const float* running_var, float eps, float* scale, float* shift, void* stream) {
  using namespace tg;
  TG_REQUIRE(running_mean && running_var && scale && shift && C > 0, "tg_bn_eval_coeff: bad arguments");
  bn_eval_coeff_kernel<<<(C + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      C, gamma, beta, running_mean, running_var, eps, scale, shift);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tg_bn_apply(const void* z, int B, int H, int W, int C, const float* scale, const float* shift,
                           int act, float slope, const uint8_t* code, void* y_nhwc, void* y_split,
                           int mask_split, void* stream) {
  return tg::bn_apply_impl<__nv_bfloat16>(z, B, H, W, C, scale, shift, act, slope, code, y_nhwc, y_split, mask_split, stream);
}
extern "C" int tg_bn_apply_f32(const void* z, int B, int H, int W, int C, const float* scale, const float* shift,
                               int act, float slope, const uint8_t* code, void* y_nhwc, void* y_split,
                               int mask_split, void* stream) {
  return tg::bn_apply_impl<float>(z, B, H, W, C, scale, shift, act, slope, code, y_nhwc, y_split, mask_split, stream);
}

extern "C" int tg_bn_bwd_reduce(const tg_grad_src* g0, const tg_grad_src* g1, const void* z, int B, int H,
                                int W, int C, const float* scale, const float* shift, int act, float slope,
                                const uint8_t* code, const float* lut_dev, float* partial, int rows_cap,
                                int* rows_used, void* stream) {
  return tg::bn_bwd_reduce_impl<__nv_bfloat16>(g0, g1, z, B, H, W, C, scale, shift, act, slope, code, lut_dev, partial,
                                               rows_cap, rows_used, stream);
}
extern "C" int tg_bn_bwd_reduce_f32(const tg_grad_src* g0, const tg_grad_src* g1, const void* z, int B, int H,
                                    int W, int C, const float* scale, const float* shift, int act, float slope,
                                    const uint8_t* code, const float* lut_dev, float* partial, int rows_cap,
                                    int* rows_used, void* stream) {
  return tg::bn_bwd_reduce_impl<float>(g0, g1, z, B, H, W, C, scale, shift, act, slope, code, lut_dev, partial, rows_cap,
                                       rows_used, stream);
}

extern "C" int tg_bn_bwd_finalize(const float* partial, int rows, int C, double count, const float* scale,
                                  const float* mean, const float* invstd, float* coeff, float* dgamma,
                                  float* dbeta, float* dbias, int accumulate, int batch_stats, void* stream) {
  using namespace tg;
  TG_REQUIRE(partial && scale && mean && invstd && coeff && rows > 0, "tg_bn_bwd_finalize: bad arguments");
  bn_bwd_finalize_kernel<<<(C + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      partial, rows, C, count, scale, mean, invstd, coeff, dgamma, dbeta, dbias, accumulate, batch_stats);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tg_bn_bwd_apply(const tg_grad_src* g0, const tg_grad_src* g1, const void* z, int B, int H, int W,
                               int C, const float* shift, const float* coeff, int act, float slope,
                               const uint8_t* code, const float* lut_dev, void* gz, void* stream) {
  return tg::bn_bwd_apply_impl<__nv_bfloat16>(g0, g1, z, B, H, W, C, shift, coeff, act, slope, code, lut_dev, gz, stream);
}
extern "C" int tg_bn_bwd_apply_f32(const tg_grad_src* g0, const tg_grad_src* g1, const void* z, int B, int H, int W,
                                   int C, const float* shift, const float* coeff, int act, float slope,
                                   const uint8_t* code, const float* lut_dev, void* gz, void* stream) {
  return tg::bn_bwd_apply_impl<float>(g0, g1, z, B, H, W, C, shift, coeff, act, slope, code, lut_dev, gz, stream);
}
