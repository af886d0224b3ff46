// direct_conv.cu — the bandwidth-bound convolutions of the path (arithmetic intensity <= 46 FLOP/B,
// SURVEY.md §8a): one input channel (enc1 7x7/s2, Discriminator model[0] 4x4/s2, VGG conv0 with its
// three identical input channels folded into one) and one output channel (`final` 3x3 64->1 with
// sigmoid + composite, Discriminator model[11] 4x4 512->1), forward, data- and weight-gradient.
// These are CUDA-core kernels on purpose: K = 16..49 or N = 1 cannot feed a 128xN tensor-core tile.
//
// Reference ops replaced:
//   PConv2d enc1: input*mask, input_conv, mask ratio       mvp_gan/src/models/pconv.py:27-43
//   Discriminator model[0] (+LeakyReLU) and model[11]       mvp_gan/src/models/discriminator.py:17,22
//   VGG16 features[0] on input.repeat(1,3,1,1)              mvp_gan/src/utils/losses.py:79-89
//   final conv + sigmoid + `out*(1-mask) + x*mask`          mvp_gan/src/models/generator.py:56-62
// and the autograd backward of each.
#include <stdlib.h>

#include "tg_common.cuh"
#include "../../include/terragan_b200.h"
#include "thin_mma.cuh"

namespace tg {

__device__ __forceinline__ long dc_split_index(int b, int h, int w, int H, int W) {
  return ((static_cast<long>(b) * 4 + 2 * (h & 1) + (w & 1)) * (H >> 1) + (h >> 1)) * (W >> 1) + (w >> 1);
}

// ------------------------------------------------------------------------------------------------
// 1 -> 64 channels, forward.  One thread = one output pixel x 64 channels; block = 32x4 pixel tile.
// ------------------------------------------------------------------------------------------------
constexpr int kC1TW = 32, kC1TH = 4;

template <int K, int S, typename TO = __nv_bfloat16>
__global__ void __launch_bounds__(128, 4)
conv_c1_fwd_kernel(const float* __restrict__ x, const uint8_t* __restrict__ xmask, int B, int H, int W, int pad,
                   const float* __restrict__ wgt /*[64][K*K]*/, const float* __restrict__ bias, int Ho, int Wo,
                   const uint8_t* __restrict__ code, const float* __restrict__ lut, int act, float slope,
                   TO* __restrict__ out, int out_split, float* __restrict__ stats) {
  constexpr int IW = (kC1TW - 1) * S + K, IH = (kC1TH - 1) * S + K;
  __shared__ float s_w[K * K][64];
  __shared__ float s_in[IH][IW + 1];
  __shared__ float s_stats[4][2][64];
  __shared__ uint4 s_stage[4][32 * 8];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 64 * K * K; i += 128) s_w[i % (K * K)][i / (K * K)] = wgt[i];
  for (int i = tid; i < 4 * 2 * 64; i += 128) (&s_stats[0][0][0])[i] = 0.f;
  const int tiles_w = (Wo + kC1TW - 1) / kC1TW, tiles_h = (Ho + kC1TH - 1) / kC1TH;
  const long total_tiles = static_cast<long>(B) * tiles_h * tiles_w;
  const int tx = tid % kC1TW, ty = tid / kC1TW;

  for (long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int tw = static_cast<int>(tile % tiles_w);
    const int th = static_cast<int>((tile / tiles_w) % tiles_h);
    const int b = static_cast<int>(tile / (static_cast<long>(tiles_w) * tiles_h));
    const int ih0 = th * kC1TH * S - pad, iw0 = tw * kC1TW * S - pad;
    __syncthreads();  // previous tile's s_in consumers are done (also covers the s_w fill)
    for (int i = tid; i < IH * IW; i += 128) {
      const int r = i / IW, c = i % IW;
      const int h = ih0 + r, w = iw0 + c;
      float v = 0.f;
      if (h >= 0 && h < H && w >= 0 && w < W) {
        const long o = (static_cast<long>(b) * H + h) * W + w;
        v = x[o];
        if (xmask != nullptr && xmask[o] == 0) v = 0.f;
      }
      s_in[r][c] = v;
    }
    __syncthreads();
    float acc[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) acc[j] = 0.f;
#pragma unroll 1
    for (int kh = 0; kh < K; ++kh) {
#pragma unroll
      for (int kw = 0; kw < K; ++kw) {
        const float xv = s_in[ty * S + kh][tx * S + kw];
        const float4* wr = reinterpret_cast<const float4*>(&s_w[kh * K + kw][0]);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 wv = wr[j];
          acc[4 * j] += xv * wv.x;
          acc[4 * j + 1] += xv * wv.y;
          acc[4 * j + 2] += xv * wv.z;
          acc[4 * j + 3] += xv * wv.w;
        }
      }
    }
    const int ho = th * kC1TH + ty, wo = tw * kC1TW + tx;
    const bool valid = ho < Ho && wo < Wo;
    const long pix = (static_cast<long>(b) * Ho + ho) * Wo + wo;
    float rs = 1.f;
    if (code != nullptr && valid) rs = __ldg(lut + code[pix]);
#pragma unroll
    for (int j = 0; j < 64; ++j) acc[j] = valid ? (acc[j] + __ldg(bias + j)) * rs : 0.f;
    if (stats != nullptr) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float sm[32], sq[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          sm[j] = acc[half * 32 + j];
          sq[j] = sm[j] * sm[j];
        }
        const float a = warp_transpose_sum32(sm);
        const float q2 = warp_transpose_sum32(sq);
        s_stats[warp][0][half * 32 + lane] += a;
        s_stats[warp][1][half * 32 + lane] += q2;
      }
    }
    // A thread owns one pixel row (128 B): storing it directly would touch 32 different lines per
    // instruction. Rows go through a per-warp XOR-swizzled staging tile and are written back 4 rows
    // (512 contiguous bytes in the plain layout) per instruction.
    if constexpr (sizeof(TO) == 4) {
      // fp32 storage (verification path): each thread writes its own 256-byte pixel row
      if (valid) {
        const long opix = out_split ? dc_split_index(b, ho, wo, Ho, Wo) : pix;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float v[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float t = acc[8 * j + q];
            if (act == 1) t = fmaxf(t, 0.f);
            else if (act == 2) t = t > 0.f ? t : t * slope;
            v[q] = t;
          }
          vstore8(reinterpret_cast<float*>(out) + opix * 64 + 8 * j, v);
        }
      }
      continue;
    }
    uint4* st = &s_stage[warp][0];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float t = acc[8 * j + q];
        if (act == 1) t = fmaxf(t, 0.f);
        else if (act == 2) t = t > 0.f ? t : t * slope;
        v[q] = t;
      }
      st[lane * 8 + (j ^ (lane & 7))] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                                   pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
    __syncwarp();
    if (ho < Ho) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = i * 4 + (lane >> 3), chunk = lane & 7;
        const int wr = tw * kC1TW + row;            // tx == lane: row r of the warp is pixel (ho, tw*32 + r)
        if (wr < Wo) {
          const long opix = out_split ? dc_split_index(b, ho, wr, Ho, Wo)
                                      : (static_cast<long>(b) * Ho + ho) * Wo + wr;
          reinterpret_cast<uint4*>(out + opix * 64)[chunk] = st[row * 8 + (chunk ^ (row & 7))];
        }
      }
    }
    __syncwarp();
  }
  if (stats != nullptr) {
    __syncthreads();
    const int c = tid % 64, which = tid / 64;
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) s += s_stats[q][which][c];
    stats[(static_cast<long>(blockIdx.x) * 2 + which) * 64 + c] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// 1 -> 64 channels, weight (and bias) gradient. Warp = pixel stream, lane = output-channel pair.
// partial[block][64][K*K (+1 for bias)]
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 dc_ld2(const __nv_bfloat16* p) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}
__device__ __forceinline__ float2 dc_ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }

template <int K, int S, typename TG = __nv_bfloat16>
__global__ void __launch_bounds__(256, 2)
conv_c1_wgrad_kernel(const float* __restrict__ x, const uint8_t* __restrict__ xmask, int B, int H, int W, int pad,
                     const TG* __restrict__ g /*[B][Ho][Wo][64]*/, int Ho, int Wo, int g_split,
                     float* __restrict__ partial) {
  // Block = one 32x4 tile of output pixels at a time (persistent over tiles); the masked, zero-padded
  // input patch is staged in shared memory once, then warp w streams 16 of the 128 pixels: lane =
  // output-channel pair, per tap one broadcast LDS + 2 FMAs.
  constexpr int T = K * K;
  constexpr int IW = (kC1TW - 1) * S + K, IH = (kC1TH - 1) * S + K;
  __shared__ float s_red[64][T + 1];
  __shared__ float s_in[IH][IW + 1];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 64 * (T + 1); i += 256) (&s_red[0][0])[i] = 0.f;
  float acc[T + 1][2];
#pragma unroll
  for (int t = 0; t <= T; ++t) acc[t][0] = acc[t][1] = 0.f;
  const int tiles_w = (Wo + kC1TW - 1) / kC1TW, tiles_h = (Ho + kC1TH - 1) / kC1TH;
  const long total_tiles = static_cast<long>(B) * tiles_h * tiles_w;
  for (long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int tw = static_cast<int>(tile % tiles_w);
    const int th = static_cast<int>((tile / tiles_w) % tiles_h);
    const int b = static_cast<int>(tile / (static_cast<long>(tiles_w) * tiles_h));
    const int ih0 = th * kC1TH * S - pad, iw0 = tw * kC1TW * S - pad;
    // gradient rows of this warp's 16 pixels, fetched in batches of 4 one batch ahead of their use
    // (one dependent 128-byte load per pixel left the kernel latency-bound at ~0.4 TB/s)
    auto load_g = [&](int q) -> float2 {
      const int pt = warp * 16 + q;
      const int ho = th * kC1TH + pt / kC1TW, wo = tw * kC1TW + pt % kC1TW;
      if (ho >= Ho || wo >= Wo) return make_float2(0.f, 0.f);   // contributes nothing
      const long p = (static_cast<long>(b) * Ho + ho) * Wo + wo;
      const long gp = g_split ? dc_split_index(b, ho, wo, Ho, Wo) : p;
      return dc_ld2(g + gp * 64 + 2 * lane);
    };
    float2 gn[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) gn[j] = load_g(j);
    __syncthreads();
    for (int i = tid; i < IH * IW; i += 256) {
      const int r = i / IW, c = i % IW;
      const int h = ih0 + r, w = iw0 + c;
      float v = 0.f;
      if (h >= 0 && h < H && w >= 0 && w < W) {
        const long o = (static_cast<long>(b) * H + h) * W + w;
        v = x[o];
        if (xmask != nullptr && xmask[o] == 0) v = 0.f;
      }
      s_in[r][c] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int qb = 0; qb < 4; ++qb) {
      float2 gc[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) gc[j] = gn[j];
      if (qb < 3) {
#pragma unroll
        for (int j = 0; j < 4; ++j) gn[j] = load_g(qb * 4 + 4 + j);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int pt = warp * 16 + qb * 4 + j;
        const int tx = pt % kC1TW, ty = pt / kC1TW;
        const float2 gf = gc[j];
        acc[T][0] += gf.x;
        acc[T][1] += gf.y;
#pragma unroll
        for (int kh = 0; kh < K; ++kh) {
#pragma unroll
          for (int kw = 0; kw < K; ++kw) {
            const float xv = s_in[ty * S + kh][tx * S + kw];
            acc[kh * K + kw][0] += xv * gf.x;
            acc[kh * K + kw][1] += xv * gf.y;
          }
        }
      }
    }
  }
  // the eight warps add their sums one after the other: a fixed order, so the result is run-to-run reproducible
  // (shared-memory float atomics made it order-dependent)
  for (int wq = 0; wq < 8; ++wq) {
    __syncthreads();
    if (warp == wq) {
#pragma unroll
      for (int t = 0; t <= T; ++t) {
        s_red[2 * lane][t] += acc[t][0];
        s_red[2 * lane + 1][t] += acc[t][1];
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * (T + 1); i += 256)
    partial[static_cast<long>(blockIdx.x) * 64 * (T + 1) + i] = (&s_red[0][0])[i];
}

// dw[co][t] (+)= sum_rows partial[row][co][t];  db[co] (+)= partial[..][co][T]
__global__ void conv_c1_wgrad_reduce_kernel(const float* __restrict__ partial, int rows, int T,
                                            float* __restrict__ dw, float* __restrict__ db, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * (T + 1)) return;
  const int co = i / (T + 1), t = i % (T + 1);
  double s = 0.0;
  for (int r = 0; r < rows; ++r) s += partial[static_cast<long>(r) * 64 * (T + 1) + i];
  if (t < T) {
    float* d = dw + co * T + t;
    *d = (accumulate ? *d : 0.f) + static_cast<float>(s);
  } else if (db != nullptr) {
    db[co] = (accumulate ? db[co] : 0.f) + static_cast<float>(s);
  }
}

// ------------------------------------------------------------------------------------------------
// C -> 1 channel "gather" convolution with per-parity tap classes.
//   acc[b][oh][ow] = sum_{t in class} sum_c x[b][base + d_t][c] * w[cls][t][c]
// ncls = 1: base = (oh, ow). ncls = 4: cls = 2*(oh&1)+(ow&1), base = (oh>>1, ow>>1)  (this form is
// the data-gradient of a stride-2 1->C convolution: Discriminator model[0]).
// mode 0: out = acc + bias.   mode 1 (generator.py:56-62): sig = sigmoid(acc + bias) is saved and
// out = sig * (1 - mask) + xin * mask.
// 8 lanes per output pixel (8 channels each, 16-byte loads), 4 pixels per warp.
// ------------------------------------------------------------------------------------------------

template <typename TX>
__global__ void __launch_bounds__(256)
conv_to1_fwd_kernel(const TX* __restrict__ x, int x_split, int B, int H, int W, int C,
                    const float* __restrict__ wgt /*[ntaps][C]*/, To1Taps taps, const float* __restrict__ bias, int Ho, int Wo, int mode,
                    const uint8_t* __restrict__ mask, const float* __restrict__ xin, float* __restrict__ out,
                    float* __restrict__ sig_out) {
  const int sub = threadIdx.x & 7;
  const long gpix = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) >> 3;
  const long stride = (static_cast<long>(gridDim.x) * blockDim.x) >> 3;
  const long M = static_cast<long>(B) * Ho * Wo;
  const long iters = (M + stride - 1) / stride;
  for (long it = 0; it < iters; ++it) {
    const long p = gpix + it * stride;
    const bool live = p < M;
    float acc = 0.f;
    int ow = 0, oh = 0, b = 0;
    if (live) {
      ow = static_cast<int>(p % Wo);
      oh = static_cast<int>((p / Wo) % Ho);
      b = static_cast<int>(p / (static_cast<long>(Wo) * Ho));
      int cls = 0, bh = oh, bw = ow;
      if (taps.ncls == 4) {
        cls = 2 * (oh & 1) + (ow & 1);
        bh = oh >> 1;
        bw = ow >> 1;
      }
      const TX* xb = x + static_cast<long>(b) * H * W * C;
      for (int t = taps.begin[cls]; t < taps.begin[cls] + taps.count[cls]; ++t) {
        const int h = bh + taps.dh[t], w = bw + taps.dw[t];
        if (h < 0 || h >= H || w < 0 || w >= W) continue;
        const TX* xp = x_split ? x + dc_split_index(b, h, w, H, W) * C
                               : xb + (static_cast<long>(h) * W + w) * C;
        const float* wp = wgt + static_cast<long>(t) * C;
        for (int c = sub * 8; c < C; c += 64) {
          float xv[8];
          vload8(xp + c, xv);
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp + c));
          const float4 w1 = __ldg(reinterpret_cast<const float4*>(wp + c) + 1);
          acc += xv[0] * w0.x + xv[1] * w0.y + xv[2] * w0.z + xv[3] * w0.w + xv[4] * w1.x + xv[5] * w1.y + xv[6] * w1.z +
                 xv[7] * w1.w;
        }
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (live && sub == 0) {
      const float v = acc + (bias ? __ldg(bias) : 0.f);
      if (mode == 0) {
        out[p] = v;
      } else {
        const float s = sizeof(TX) == 4 ? 1.f / (1.f + expf(-v)) : 1.f / (1.f + __expf(-v));
        if (sig_out) sig_out[p] = s;
        const float m = mask[p] ? 1.f : 0.f;
        out[p] = s * (1.f - m) + xin[p] * m;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 3x3 / stride 1 / pad 1, C = 64 specialisation of the C->1 gather (final conv fwd + composite, VGG
// conv0 data gradient) and of its weight gradient. Block = 8x32 output tile; the 10x34-pixel halo
// tile (43.5 KB of bf16) is staged in shared memory with coalesced 16-byte loads so every input byte
// is read from HBM once; 8 lanes x 8 channels per pixel quad; each 16-byte LDS feeds up to 3 taps.
// ------------------------------------------------------------------------------------------------
constexpr int kT1H = 8, kT1W = 32;
struct Tap3x3 { int idx[9]; };   // idx[(dh+1)*3 + (dw+1)] = row of wgt for that offset

__device__ __forceinline__ void t1_load_halo(__nv_bfloat16 (*s_x)[kT1W + 2][64], const __nv_bfloat16* __restrict__ x,
                                             int b, int h0, int w0, int H, int W) {
  for (int i = threadIdx.x; i < (kT1H + 2) * (kT1W + 2) * 8; i += 256) {
    const int v = i & 7, pix = i >> 3;
    const int hr = pix / (kT1W + 2), hc = pix % (kT1W + 2);
    const int h = h0 - 1 + hr, w = w0 - 1 + hc;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (h >= 0 && h < H && w >= 0 && w < W)
      val = *reinterpret_cast<const uint4*>(x + ((static_cast<long>(b) * H + h) * W + w) * 64 + v * 8);
    *reinterpret_cast<uint4*>(&s_x[hr][hc][v * 8]) = val;
  }
}

__device__ __forceinline__ void t1_unpack(const uint4& raw, float (&f)[8]) {
  const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float2 t = __bfloat1622float2(hh[q]);
    f[2 * q] = t.x;
    f[2 * q + 1] = t.y;
  }
}

__global__ void __launch_bounds__(256)
conv3x3_c64_to1_kernel(const __nv_bfloat16* __restrict__ x, int B, int H, int W, const float* __restrict__ wgt,
                       Tap3x3 tp, const float* __restrict__ bias, int mode, const uint8_t* __restrict__ mask,
                       const float* __restrict__ xin, float* __restrict__ out, float* __restrict__ sig_out) {
  __shared__ __align__(16) __nv_bfloat16 s_x[kT1H + 2][kT1W + 2][64];
  const int sub = threadIdx.x & 7, grp = threadIdx.x >> 3;
  float wr[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) wr[t][j] = __ldg(wgt + tp.idx[t] * 64 + sub * 8 + j);
  const float bv = bias ? __ldg(bias) : 0.f;
  const int tiles_w = (W + kT1W - 1) / kT1W, tiles_h = (H + kT1H - 1) / kT1H;
  const long total_tiles = static_cast<long>(B) * tiles_h * tiles_w;
  for (long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int tw = static_cast<int>(tile % tiles_w);
    const int th = static_cast<int>((tile / tiles_w) % tiles_h);
    const int b = static_cast<int>(tile / (static_cast<long>(tiles_w) * tiles_h));
    const int h0 = th * kT1H, w0 = tw * kT1W;
    __syncthreads();
    t1_load_halo(s_x, x, b, h0, w0, H, W);
    __syncthreads();
#pragma unroll
    for (int ps = 0; ps < 2; ++ps) {
      const int q = ps * 32 + grp;
      const int row = q >> 3, col0 = (q & 7) * 4;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) {
          float v[8];
          t1_unpack(*reinterpret_cast<const uint4*>(&s_x[row + r][col0 + cc][sub * 8]), v);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int kw = cc - j;
            if (kw >= 0 && kw < 3) {
#pragma unroll
              for (int e = 0; e < 8; ++e) acc[j] += v[e] * wr[r * 3 + kw][e];
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 4);
        acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 2);
        acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 1);
      }
      const int h = h0 + row, w = w0 + col0;
      if (sub == 0 && h < H) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (w + j >= W) break;
          const long p = (static_cast<long>(b) * H + h) * W + w + j;
          const float val = acc[j] + bv;
          if (mode == 0) {
            out[p] = val;
          } else {
            const float sg = 1.f / (1.f + __expf(-val));
            if (sig_out) sig_out[p] = sg;
            const float m = mask[p] ? 1.f : 0.f;
            out[p] = sg * (1.f - m) + xin[p] * m;
          }
        }
      }
    }
  }
}

// dw[t][c] = sum_o g[o] * x[o + d_t][c], db = sum_o g[o]; partial[block][9][64], partial_b[block]
__global__ void __launch_bounds__(256)
conv3x3_c64_to1_wgrad_kernel(const __nv_bfloat16* __restrict__ x, int B, int H, int W, const float* __restrict__ g,
                             Tap3x3 tp, float* __restrict__ partial, float* __restrict__ partial_b) {
  __shared__ __align__(16) __nv_bfloat16 s_x[kT1H + 2][kT1W + 2][64];
  __shared__ float s_red[9][64];
  __shared__ float s_b;
  const int sub = threadIdx.x & 7, grp = threadIdx.x >> 3;
  for (int i = threadIdx.x; i < 9 * 64; i += 256) (&s_red[0][0])[i] = 0.f;
  if (threadIdx.x == 0) s_b = 0.f;
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
  float gsum = 0.f;
  const int tiles_w = (W + kT1W - 1) / kT1W, tiles_h = (H + kT1H - 1) / kT1H;
  const long total_tiles = static_cast<long>(B) * tiles_h * tiles_w;
  for (long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int tw = static_cast<int>(tile % tiles_w);
    const int th = static_cast<int>((tile / tiles_w) % tiles_h);
    const int b = static_cast<int>(tile / (static_cast<long>(tiles_w) * tiles_h));
    const int h0 = th * kT1H, w0 = tw * kT1W;
    __syncthreads();
    t1_load_halo(s_x, x, b, h0, w0, H, W);
    __syncthreads();
#pragma unroll
    for (int ps = 0; ps < 2; ++ps) {
      const int q = ps * 32 + grp;
      const int row = q >> 3, col0 = (q & 7) * 4;
      const int h = h0 + row, w = w0 + col0;
      float gv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        gv[j] = (h < H && w + j < W) ? __ldg(g + (static_cast<long>(b) * H + h) * W + w + j) : 0.f;
      if (sub == 0) gsum += gv[0] + gv[1] + gv[2] + gv[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) {
          float v[8];
          t1_unpack(*reinterpret_cast<const uint4*>(&s_x[row + r][col0 + cc][sub * 8]), v);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int kw = cc - j;
            if (kw >= 0 && kw < 3) {
#pragma unroll
              for (int e = 0; e < 8; ++e) acc[r * 3 + kw][e] += gv[j] * v[e];
            }
          }
        }
      }
    }
  }
  // the 32 pixel groups add their sums in a fixed order (deterministic; no float atomics)
  for (int gq = 0; gq < 32; ++gq) {
    __syncthreads();
    if (grp == gq) {
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int j = 0; j < 8; ++j) s_red[t][sub * 8 + j] += acc[t][j];
      if (sub == 0) s_b += gsum;
    }
  }
  __syncthreads();
  // s_red is indexed by spatial offset (dh+1)*3+(dw+1); emit in the caller's tap order
  for (int i = threadIdx.x; i < 9 * 64; i += 256) {
    const int o = i / 64, c = i % 64;
    partial[(static_cast<long>(blockIdx.x) * 9 + tp.idx[o]) * 64 + c] = s_red[o][c];
  }
  if (threadIdx.x == 0 && partial_b != nullptr) partial_b[blockIdx.x] = s_b;
}

// ------------------------------------------------------------------------------------------------
// Data gradient of a 4x4 / stride-2 / pad-1 convolution with ONE input channel and 64 output channels
// (Discriminator model[0], discriminator.py:17): out[b][h][w] = sum over the 2x2 taps of parity class
// (h&1, w&1) of <g[b][(h>>1)+dh][(w>>1)+dw][:], W[tap][:]>. g is bf16 [B][Hg][Wg][64], optionally in
// parity-split storage. Block = 8x32 tile of g pixels (16x64 outputs); the 10x34 halo tile of g is
// staged in shared memory once; 8 lanes x 8 channels per g pixel produce its four outputs.
// s_w[cls][pos][c]: weight of class cls at neighbourhood position pos = (dh+1)*3 + (dw+1), 0 if unused.
// ------------------------------------------------------------------------------------------------
constexpr int kS2H = 6;
struct S2Taps { unsigned used[4]; int row[4][9]; };   // used[cls] bitmask over pos; row = weight row or -1

__global__ void __launch_bounds__(256)
conv_s2_c64_to1_kernel(const __nv_bfloat16* __restrict__ g, int g_split, int B, int Hg, int Wg,
                       const float* __restrict__ wgt, S2Taps tp, float* __restrict__ out) {
  __shared__ __align__(16) __nv_bfloat16 s_x[kS2H + 2][kT1W + 2][64];
  __shared__ __align__(16) float s_w[4][9][64];
  const int sub = threadIdx.x & 7, grp = threadIdx.x >> 3;
  for (int i = threadIdx.x; i < 4 * 9 * 64; i += 256) {
    const int c = i & 63, pos = (i >> 6) % 9, cls = i / (9 * 64);
    const int r = tp.row[cls][pos];
    s_w[cls][pos][c] = r >= 0 ? __ldg(wgt + r * 64 + c) : 0.f;
  }
  const int tiles_w = (Wg + kT1W - 1) / kT1W, tiles_h = (Hg + kS2H - 1) / kS2H;
  const long total_tiles = static_cast<long>(B) * tiles_h * tiles_w;
  const int H = 2 * Hg, W = 2 * Wg;
  for (long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int tw = static_cast<int>(tile % tiles_w);
    const int th = static_cast<int>((tile / tiles_w) % tiles_h);
    const int b = static_cast<int>(tile / (static_cast<long>(tiles_w) * tiles_h));
    const int h0 = th * kS2H, w0 = tw * kT1W;
    __syncthreads();
    for (int i = threadIdx.x; i < (kS2H + 2) * (kT1W + 2) * 8; i += 256) {
      const int v = i & 7, pix = i >> 3;
      const int hr = pix / (kT1W + 2), hc = pix % (kT1W + 2);
      const int h = h0 - 1 + hr, w = w0 - 1 + hc;
      uint4 val = make_uint4(0u, 0u, 0u, 0u);
      if (h >= 0 && h < Hg && w >= 0 && w < Wg) {
        const long gp = g_split ? dc_split_index(b, h, w, Hg, Wg) : (static_cast<long>(b) * Hg + h) * Wg + w;
        val = *reinterpret_cast<const uint4*>(g + gp * 64 + v * 8);
      }
      *reinterpret_cast<uint4*>(&s_x[hr][hc][v * 8]) = val;
    }
    __syncthreads();
#pragma unroll 1
    for (int ps = 0; ps < (kS2H * kT1W) / 32; ++ps) {
      const int q = ps * 32 + grp;
      const int row = q / kT1W, col = q % kT1W;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int pos = 0; pos < 9; ++pos) {
        float v[8];
        t1_unpack(*reinterpret_cast<const uint4*>(&s_x[row + pos / 3][col + pos % 3][sub * 8]), v);
#pragma unroll
        for (int cls = 0; cls < 4; ++cls) {
          if ((tp.used[cls] >> pos) & 1u) {
            const float4 wa = *reinterpret_cast<const float4*>(&s_w[cls][pos][sub * 8]);
            const float4 wb = *reinterpret_cast<const float4*>(&s_w[cls][pos][sub * 8 + 4]);
            acc[cls] += v[0] * wa.x + v[1] * wa.y + v[2] * wa.z + v[3] * wa.w + v[4] * wb.x + v[5] * wb.y +
                        v[6] * wb.z + v[7] * wb.w;
          }
        }
      }
#pragma unroll
      for (int cls = 0; cls < 4; ++cls) {
        acc[cls] += __shfl_xor_sync(0xffffffffu, acc[cls], 4);
        acc[cls] += __shfl_xor_sync(0xffffffffu, acc[cls], 2);
        acc[cls] += __shfl_xor_sync(0xffffffffu, acc[cls], 1);
      }
      const int gh = h0 + row, gw = w0 + col;
      if (sub == 0 && gh < Hg && gw < Wg) {
        float* o = out + (static_cast<long>(b) * H + 2 * gh) * W + 2 * gw;
        *reinterpret_cast<float2*>(o) = make_float2(acc[0], acc[1]);
        *reinterpret_cast<float2*>(o + W) = make_float2(acc[2], acc[3]);
      }
    }
  }
}

// class (P,Q) = 2*P+Q with taps inside the 3x3 neighbourhood -> S2Taps; false if the pattern does not fit
static bool make_s2taps(S2Taps* tp, const int* cls_count, const int8_t* dh, const int8_t* dw) {
  int t = 0;
  for (int cls = 0; cls < 4; ++cls) {
    tp->used[cls] = 0;
    for (int i = 0; i < 9; ++i) tp->row[cls][i] = -1;
    for (int i = 0; i < cls_count[cls]; ++i, ++t) {
      if (dh[t] < -1 || dh[t] > 1 || dw[t] < -1 || dw[t] > 1) return false;
      const int pos = (dh[t] + 1) * 3 + dw[t] + 1;
      tp->used[cls] |= 1u << pos;
      tp->row[cls][pos] = t;
    }
  }
  return true;
}

static bool make_tap3x3(Tap3x3* tp, int ntaps, const int8_t* dh, const int8_t* dw) {
  if (ntaps != 9) return false;
  for (int i = 0; i < 9; ++i) tp->idx[i] = -1;
  for (int t = 0; t < 9; ++t) {
    if (dh[t] < -1 || dh[t] > 1 || dw[t] < -1 || dw[t] > 1) return false;
    tp->idx[(dh[t] + 1) * 3 + dw[t] + 1] = t;
  }
  for (int i = 0; i < 9; ++i)
    if (tp->idx[i] < 0) return false;
  return true;
}

// data gradient of a stride-1 C->1 conv: dx[b][h][w][c] = sum_t g[b][h - dh_t][w - dw_t] * w[t][c]
template <typename TO>
__global__ void __launch_bounds__(256)
conv_to1_bwd_data_kernel(const float* __restrict__ g, int B, int Ho, int Wo, const float* __restrict__ wgt, To1Taps taps,
                         int H, int W, int C, TO* __restrict__ dx) {
  const int cv = C >> 3;
  const long total = static_cast<long>(B) * H * W * cv;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long p = i / cv;
    const int c = static_cast<int>(i % cv) << 3;
    const int w = static_cast<int>(p % W);
    const int h = static_cast<int>((p / W) % H);
    const int b = static_cast<int>(p / (static_cast<long>(W) * H));
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int t = 0; t < taps.count[0]; ++t) {
      const int oh = h - taps.dh[t], ow = w - taps.dw[t];
      if (oh < 0 || oh >= Ho || ow < 0 || ow >= Wo) continue;
      const float gv = __ldg(g + (static_cast<long>(b) * Ho + oh) * Wo + ow);
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(wgt + static_cast<long>(t) * C + c));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(wgt + static_cast<long>(t) * C + c) + 1);
      acc[0] += gv * w0.x; acc[1] += gv * w0.y; acc[2] += gv * w0.z; acc[3] += gv * w0.w;
      acc[4] += gv * w1.x; acc[5] += gv * w1.y; acc[6] += gv * w1.z; acc[7] += gv * w1.w;
    }
    vstore8(dx + p * C + c, acc);
  }
}

// 3x3 / C = 64 specialisation: the 72 weights of a thread's 8-channel group live in registers
__global__ void __launch_bounds__(256)
conv3x3_c64_to1_bwd_data_kernel(const float* __restrict__ g, int B, int H, int W, const float* __restrict__ wgt,
                                Tap3x3 tp, __nv_bfloat16* __restrict__ dx) {
  const int sub = threadIdx.x & 7;
  float wr[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) wr[t][j] = __ldg(wgt + tp.idx[t] * 64 + sub * 8 + j);
  const unsigned total = static_cast<unsigned>(B) * H * W;          // pixels; host checks < 2^31
  const unsigned HW = static_cast<unsigned>(H) * W;
  const unsigned stride = (gridDim.x * blockDim.x) >> 3;
  constexpr int U = 4;                                               // pixels in flight per thread
  for (unsigned pbase = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; pbase < total; pbase += U * stride) {
    float gv[U][9];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned p = pbase + u * stride;
      const unsigned pc = p < total ? p : total - 1;
      const unsigned bimg = pc / HW, rem = pc - bimg * HW;
      const int h = static_cast<int>(rem / W);
      const int w = static_cast<int>(rem - h * W);
      const float* gb = g + static_cast<size_t>(bimg) * HW;
#pragma unroll
      for (int dh = -1; dh <= 1; ++dh) {
#pragma unroll
        for (int dw = -1; dw <= 1; ++dw) {
          const int oh = h - dh, ow = w - dw;
          const bool in = oh >= 0 && oh < H && ow >= 0 && ow < W;
          gv[u][(dh + 1) * 3 + dw + 1] = in ? __ldg(gb + oh * W + ow) : 0.f;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned p = pbase + u * stride;
      if (p >= total) break;
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += gv[u][t] * wr[t][j];
      *reinterpret_cast<uint4*>(dx + static_cast<size_t>(p) * 64 + sub * 8) =
          make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]),
                     pack_bf16x2(acc[6], acc[7]));
    }
  }
}

// weight gradient of a stride-1 C->1 conv: dw[t][c] = sum_o g[o] * x[o + d_t][c]; db = sum_o g[o].
// grid = (pixel blocks, C/64); 8 lanes per pixel; each lane keeps T x 8 accumulators.
template <int T, typename TX = __nv_bfloat16>
__global__ void __launch_bounds__(128)
conv_to1_wgrad_kernel(const TX* __restrict__ x, int B, int H, int W, int C, const float* __restrict__ g,
                      int Ho, int Wo, To1Taps taps, float* __restrict__ partial /*[gridDim.x][T][C]*/,
                      float* __restrict__ partial_b /*[gridDim.x]*/) {
  __shared__ float s_red[T][64];  // [tap][channel of the slab], accumulated with shared atomics
  __shared__ float s_b;
  const int sub = threadIdx.x & 7, grp = threadIdx.x >> 3;  // 16 pixel groups
  for (int i = threadIdx.x; i < T * 64; i += 128) (&s_red[0][0])[i] = 0.f;
  if (threadIdx.x == 0) s_b = 0.f;
  __syncthreads();
  const int c0 = blockIdx.y * 64 + sub * 8;
  float acc[T][8];
#pragma unroll
  for (int t = 0; t < T; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
  float gsum = 0.f;
  const long M = static_cast<long>(B) * Ho * Wo;
  for (long p = static_cast<long>(blockIdx.x) * 16 + grp; p < M; p += static_cast<long>(gridDim.x) * 16) {
    const int ow = static_cast<int>(p % Wo);
    const int oh = static_cast<int>((p / Wo) % Ho);
    const int b = static_cast<int>(p / (static_cast<long>(Wo) * Ho));
    const float gv = __ldg(g + p);
    gsum += gv;
    const TX* xb = x + static_cast<long>(b) * H * W * C + c0;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int h = oh + taps.dh[t], w = ow + taps.dw[t];
      if (h < 0 || h >= H || w < 0 || w >= W) continue;
      float f[8];
      vload8(xb + (static_cast<long>(h) * W + w) * C, f);
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[t][q] += gv * f[q];
    }
  }
  // the 16 pixel groups add their sums in a fixed order (deterministic; no float atomics)
  for (int gq = 0; gq < 16; ++gq) {
    __syncthreads();
    if (grp == gq) {
#pragma unroll
      for (int t = 0; t < T; ++t)
#pragma unroll
        for (int j = 0; j < 8; ++j) s_red[t][sub * 8 + j] += acc[t][j];
      if (sub == 0) s_b += gsum;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * 64; i += 128) {
    const int t = i / 64, c = i % 64;
    partial[(static_cast<long>(blockIdx.x) * T + t) * C + blockIdx.y * 64 + c] = s_red[t][c];
  }
  if (threadIdx.x == 0 && blockIdx.y == 0 && partial_b != nullptr) partial_b[blockIdx.x] = s_b;
}

// dw_out[c][perm[t]] (+)= sum_rows partial[row][t][c]   (PyTorch layout [1][C][kh][kw]);  db (+)= sum partial_b
// Block = 32 outputs x 8 row lanes: the rows are summed 8-way in parallel (one serial chain over several hundred
// rows was latency-bound: 160 us for 592 rows), then combined in a fixed order (deterministic).
__global__ void __launch_bounds__(256)
conv_to1_wgrad_reduce_kernel(const float* __restrict__ partial, const float* __restrict__ partial_b,
                             int rows, int T, int C, float* __restrict__ dw, float* __restrict__ db,
                             int accumulate) {
  __shared__ double s_sum[8][32];
  const int ol = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + ol;
  double s = 0.0;
  if (i < T * C)
    for (int r = rl; r < rows; r += 8) s += partial[static_cast<long>(r) * T * C + i];
  s_sum[rl][ol] = s;
  __syncthreads();
  if (rl == 0 && i < T * C) {
#pragma unroll
    for (int q = 1; q < 8; ++q) s += s_sum[q][ol];
    const int t = i / C, c = i % C;
    float* d = dw + static_cast<long>(c) * T + t;
    *d = (accumulate ? *d : 0.f) + static_cast<float>(s);
  }
  if (blockIdx.x == 0 && db != nullptr && partial_b != nullptr) {
    __syncthreads();
    double b = 0.0;
    for (int r = threadIdx.x; r < rows; r += 256) b += partial_b[r];
    // fixed-order block reduction
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
    if (ol == 0) s_sum[rl][0] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
#pragma unroll
      for (int q = 0; q < 8; ++q) tot += s_sum[q][0];
      db[0] = (accumulate ? db[0] : 0.f) + static_cast<float>(tot);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 512 -> 1 channels with a 4x4 window (Discriminator model[11], discriminator.py:22): forward, data- and weight-
// gradient. 1 GFLOP per call at batch 64 — far too small for a 128-row tensor-core tile to matter — but the generic
// kernels above re-read every activation row once per tap through L2 and spent 0.17 / 0.15 / 0.48 ms per launch
// (1.9 ms of the step). These read or write each activation exactly once:
//   forward   per-pixel tap dot products T[t][q] = <x[q][:], w[t][:]> (a warp owns 4 pixels, a lane 16 channels; the 64
//             sums of a warp are reduced by two transposing butterflies), then tg::to1_tapsum_launch does the shifted sum
//   dgrad     a thread keeps the 16 x 8 weights of its 8 channels in registers and walks over pixels
//   wgrad     a thread keeps the 16 x 8 accumulators of its 8 channels in registers; the four pixel lanes of a block and
//             then the blocks are summed in a fixed order (deterministic)
// ------------------------------------------------------------------------------------------------
constexpr int kWideC = 512, kWideT = 16;

__global__ void __launch_bounds__(256)
to1_wide_tapdot_kernel(const __nv_bfloat16* __restrict__ x, long total, const float* __restrict__ wgt /*[16][512]*/,
                       float* __restrict__ T /*[16][total]*/) {
  __shared__ float4 s_w[kWideT * 4 * 32];                 // [t][i][lane]: float4 i of the lane's 16 channels, conflict-free
  for (int i = threadIdx.x; i < kWideT * kWideC / 4; i += 256) {
    const int t = i / (kWideC / 4), c4 = i % (kWideC / 4);      // channels 4*c4 .. 4*c4+3
    const int lane = c4 / 4, q = c4 % 4;
    s_w[(t * 4 + q) * 32 + lane] = *reinterpret_cast<const float4*>(wgt + t * kWideC + 4 * c4);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long warp_id = (blockIdx.x * 256L + threadIdx.x) >> 5, n_warps = (gridDim.x * 256L) >> 5;
  for (long p0 = warp_id * 4; p0 < total; p0 += n_warps * 4) {
    float xv[4][16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (p0 + j < total) {
        float a[8], b[8];
        vload8(x + (p0 + j) * kWideC + lane * 16, a);
        vload8(x + (p0 + j) * kWideC + lane * 16 + 8, b);
#pragma unroll
        for (int e = 0; e < 8; ++e) { xv[j][e] = a[e]; xv[j][8 + e] = b[e]; }
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) xv[j][e] = 0.f;
      }
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float v[32];                                         // v[(t - 8 half) * 4 + pixel]
#pragma unroll
      for (int tt = 0; tt < 8; ++tt) {
        const int t = half * 8 + tt;
        float wv[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 w4 = s_w[(t * 4 + q) * 32 + lane];
          wv[4 * q] = w4.x; wv[4 * q + 1] = w4.y; wv[4 * q + 2] = w4.z; wv[4 * q + 3] = w4.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float acc = 0.f;
#pragma unroll
          for (int e = 0; e < 16; ++e) acc = fmaf(xv[j][e], wv[e], acc);
          v[tt * 4 + j] = acc;
        }
      }
      const float tot = warp_transpose_sum32(v);           // lane l now holds the warp total of v[l]
      const int t = half * 8 + (lane >> 2), j = lane & 3;
      if (p0 + j < total) T[static_cast<long>(t) * total + p0 + j] = tot;
    }
  }
}

// dx[q][c] = sum_t g[q - d_t] * w[t][c].  Block = 64 channel groups (8 channels each) x 4 pixel lanes; a thread produces 4
// pixels per iteration so that every pair of 16-byte weight reads from shared memory feeds 32 FMAs (with the 128
// weights of a thread held in registers the kernel ran one block per SM and was latency-bound: 155 us).
__global__ void __launch_bounds__(256)
to1_wide_bwd_data_kernel(const float* __restrict__ g, int B, int Ho, int Wo, const float* __restrict__ wgt, To1Taps taps,
                         int H, int W, __nv_bfloat16* __restrict__ dx) {
  __shared__ float4 s_w[kWideT][2][64];                     // [t][half][channel group]
  for (int i = threadIdx.x; i < kWideT * 128; i += 256) {
    const int t = i / 128, r = i % 128, cg = r >> 1, half = r & 1;
    s_w[t][half][cg] = __ldg(reinterpret_cast<const float4*>(wgt + t * kWideC + cg * 8) + half);
  }
  __syncthreads();
  const int cg = threadIdx.x & 63, pl = threadIdx.x >> 6;
  const unsigned total = static_cast<unsigned>(B) * H * W, HW = static_cast<unsigned>(H) * W;
  constexpr int U = 4;
  for (unsigned q0 = (blockIdx.x * 4 + pl) * U; q0 < total; q0 += gridDim.x * 4 * U) {
    float acc[U][8];
    const float* gp[U];
    int hh[U], ww[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned q = min(q0 + u, total - 1);
      const unsigned b = q / HW, rem = q - b * HW;
      hh[u] = static_cast<int>(rem / W);
      ww[u] = static_cast<int>(rem - static_cast<unsigned>(hh[u]) * W);
      gp[u] = g + static_cast<size_t>(b) * Ho * Wo;
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[u][e] = 0.f;
    }
#pragma unroll
    for (int t = 0; t < kWideT; ++t) {
      const float4 wa = s_w[t][0][cg], wb = s_w[t][1][cg];
      const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int oh = hh[u] - taps.dh[t], ow = ww[u] - taps.dw[t];
        const float gv = (oh >= 0 && oh < Ho && ow >= 0 && ow < Wo) ? __ldg(gp[u] + oh * Wo + ow) : 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[u][e] = fmaf(gv, wv[e], acc[u][e]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (q0 + u < total) vstore8(dx + static_cast<size_t>(q0 + u) * kWideC + cg * 8, acc[u]);
  }
}

// partial[block][t][c] = sum over the block's input pixels q of x[q][c] * g[q - d_t];  partial_b[block] = its share of sum g.
// Block = 64 channel groups x 2 tap halves x 2 pixel lanes: a thread keeps 8 taps x 8 channels of accumulators (with all
// 16 taps per thread: 169 registers, one block per SM, 185 us).
__global__ void __launch_bounds__(256, 2)
to1_wide_wgrad_kernel(const __nv_bfloat16* __restrict__ x, int B, int H, int W, const float* __restrict__ g, int Ho, int Wo,
                      To1Taps taps, float* __restrict__ partial, float* __restrict__ partial_b) {
  extern __shared__ float s_wide[];                         // dynamic: [kWideT][kWideC] sums of pixel lane 1 + 8 floats
  float* s_b = s_wide + kWideT * kWideC;
  const int cg = threadIdx.x & 63, th = (threadIdx.x >> 6) & 1, pl = threadIdx.x >> 7;
  float acc[8][8];
#pragma unroll
  for (int t = 0; t < 8; ++t)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[t][e] = 0.f;
  const unsigned total = static_cast<unsigned>(B) * H * W, HW = static_cast<unsigned>(H) * W;
  const unsigned per = (total + gridDim.x - 1) / gridDim.x;      // contiguous pixel range per block
  const unsigned q_begin = blockIdx.x * per, q_end = min(total, q_begin + per);
  for (unsigned q = q_begin + pl; q < q_end; q += 2) {
    const unsigned b = q / HW, rem = q - b * HW;
    const int h = static_cast<int>(rem / W), w = static_cast<int>(rem - static_cast<unsigned>(h) * W);
    const float* gb = g + static_cast<size_t>(b) * Ho * Wo;
    float xv[8];
    vload8(x + static_cast<size_t>(q) * kWideC + cg * 8, xv);
#pragma unroll
    for (int tt = 0; tt < 8; ++tt) {
      // both tap halves are unrolled with compile-time tap indices; `th` selects at run time (uniform per warp)
      const int dh = th ? taps.dh[8 + tt] : taps.dh[tt], dw = th ? taps.dw[8 + tt] : taps.dw[tt];
      const int oh = h - dh, ow = w - dw;
      const float gv = (oh >= 0 && oh < Ho && ow >= 0 && ow < Wo) ? __ldg(gb + oh * Wo + ow) : 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[tt][e] = fmaf(gv, xv[e], acc[tt][e]);
    }
  }
  if (pl == 1) {
#pragma unroll
    for (int tt = 0; tt < 8; ++tt)
#pragma unroll
      for (int e = 0; e < 8; ++e) s_wide[(th * 8 + tt) * kWideC + cg * 8 + e] = acc[tt][e];
  }
  // bias gradient: this block's contiguous share of sum g
  const unsigned gtot = static_cast<unsigned>(B) * Ho * Wo, gper = (gtot + gridDim.x - 1) / gridDim.x;
  float gs = 0.f;
  for (unsigned i = blockIdx.x * gper + threadIdx.x; i < min(gtot, (blockIdx.x + 1) * gper); i += 256) gs += __ldg(g + i);
  gs = warp_sum(gs);
  if ((threadIdx.x & 31) == 0) s_b[threadIdx.x >> 5] = gs;
  __syncthreads();
  if (pl == 0) {
#pragma unroll
    for (int tt = 0; tt < 8; ++tt)
#pragma unroll
      for (int e = 0; e < 8; ++e)
        partial[(static_cast<size_t>(blockIdx.x) * kWideT + th * 8 + tt) * kWideC + cg * 8 + e] =
            acc[tt][e] + s_wide[(th * 8 + tt) * kWideC + cg * 8 + e];
  }
  if (threadIdx.x == 0 && partial_b != nullptr) {
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) v += s_b[q];
    partial_b[blockIdx.x] = v;
  }
}

static bool wide_shape(int C, int ntaps, long pixels) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("TG_NO_WIDE_TO1"); off = (e && e[0] == '1') ? 1 : 0; }
  return off == 0 && C == kWideC && ntaps == kWideT && pixels < (1L << 31);
}

static int fill_taps(To1Taps* t, int ncls, const int* count, const int8_t* dh, const int8_t* dw) {
  int n = 0;
  t->ncls = ncls;
  for (int c = 0; c < 4; ++c) {
    t->begin[c] = n;
    t->count[c] = c < ncls ? count[c] : 0;
    n += t->count[c];
  }
  if (n > TG_MAX_TAPS) return -1;
  for (int i = 0; i < n; ++i) {
    t->dh[i] = dh[i];
    t->dw[i] = dw[i];
  }
  return n;
}

static int dc_grid(long n, int block, int per_sm) {
  long g = (n + block - 1) / block;
  const long cap = static_cast<long>(num_sms() > 0 ? num_sms() : 148) * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace tg

#define TG_C1_DISPATCH(K_, S_, CALL)            \
  if (k == K_ && s == S_) {                     \
    constexpr int K = K_, S = S_;               \
    CALL;                                       \
    handled = true;                             \
  }

extern "C" int tg_conv_c1_fwd(const float* x, const uint8_t* xmask, int B, int H, int W, int k, int s, int pad,
                              const float* wgt, const float* bias, const uint8_t* code, const float* lut_dev, int act,
                              float slope, void* out, int out_split, float* stats, int stats_rows_cap,
                              int* stats_rows_used, void* stream) {
  using namespace tg;
  TG_REQUIRE(x && wgt && bias && out, "tg_conv_c1_fwd: null pointer");
  TG_REQUIRE(!code || lut_dev, "tg_conv_c1_fwd: code needs a device LUT");
  const int Ho = (H + 2 * pad - k) / s + 1, Wo = (W + 2 * pad - k) / s + 1;
  TG_REQUIRE(!out_split || (Ho % 2 == 0 && Wo % 2 == 0), "tg_conv_c1_fwd: parity-split output needs even Ho, Wo");
  const long tiles = static_cast<long>(B) * ((Ho + kC1TH - 1) / kC1TH) * ((Wo + kC1TW - 1) / kC1TW);
  int grid = static_cast<int>(tiles < 4L * num_sms() ? tiles : 4L * num_sms());
  if (stats) {
    TG_REQUIRE(stats_rows_used != nullptr, "tg_conv_c1_fwd: stats_rows_used is null");
    if (grid > stats_rows_cap) grid = stats_rows_cap;
    TG_REQUIRE(grid >= 1, "tg_conv_c1_fwd: stats_rows_cap must be >= 1");
    *stats_rows_used = grid;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  if (thin_mma_enabled() && (k == 3 || k == 4 || k == 7) && static_cast<long>(B) * Ho * Wo < (1L << 31) &&
      (act != 2 || (slope >= 0.f && slope <= 1.f))) {
    RowGemmParams rp{};
    rp.src = x; rp.src_mask = xmask; rp.B = B; rp.H = H; rp.W = W; rp.Ho = Ho; rp.Wo = Wo;
    rp.S = s; rp.pad = pad; rp.flip = 0;
    rp.wgt = wgt; rp.w_sn = k * k; rp.w_st = 1;
    for (int t = 0; t < k * k; ++t) rp.perm[t] = static_cast<int8_t>(t);
    rp.bias = bias; rp.code = code; rp.lut = lut_dev; rp.act = act; rp.slope = slope;
    rp.out = o; rp.out_split = out_split; rp.stats = stats;
    rp.total = static_cast<unsigned>(static_cast<long>(B) * Ho * Wo);
    int used = 0;
    TG_REQUIRE(rowgemm_dispatch(k, rp, stats ? stats_rows_cap : 0, &used, st) == 0, "tg_conv_c1_fwd: launch failed");
    if (stats) *stats_rows_used = used;
    return 0;
  }
  bool handled = false;
  TG_C1_DISPATCH(7, 2, (conv_c1_fwd_kernel<K, S><<<grid, 128, 0, st>>>(x, xmask, B, H, W, pad, wgt, bias, Ho, Wo, code, lut_dev, act, slope, o, out_split, stats)))
  TG_C1_DISPATCH(4, 2, (conv_c1_fwd_kernel<K, S><<<grid, 128, 0, st>>>(x, xmask, B, H, W, pad, wgt, bias, Ho, Wo, code, lut_dev, act, slope, o, out_split, stats)))
  TG_C1_DISPATCH(3, 1, (conv_c1_fwd_kernel<K, S><<<grid, 128, 0, st>>>(x, xmask, B, H, W, pad, wgt, bias, Ho, Wo, code, lut_dev, act, slope, o, out_split, stats)))
  TG_REQUIRE(handled, "tg_conv_c1_fwd: unsupported window k=%d s=%d (supported: 7/2, 4/2, 3/1)", k, s);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tg_conv_c1_wgrad(const float* x, const uint8_t* xmask, int B, int H, int W, int k, int s, int pad,
                                const void* g, int g_split, float* partial, int rows_cap, float* dw, float* db,
                                int accumulate, void* stream) {
  using namespace tg;
  TG_REQUIRE(x && g && partial && dw, "tg_conv_c1_wgrad: null pointer");
  const int Ho = (H + 2 * pad - k) / s + 1, Wo = (W + 2 * pad - k) / s + 1;
  const long tiles = static_cast<long>(B) * ((Ho + kC1TH - 1) / kC1TH) * ((Wo + kC1TW - 1) / kC1TW);
  int grid = static_cast<int>(tiles < 2L * num_sms() ? tiles : 2L * num_sms());
  if (grid > rows_cap) grid = rows_cap;
  TG_REQUIRE(grid >= 1, "tg_conv_c1_wgrad: rows_cap must be >= 1");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const __nv_bfloat16* gg = reinterpret_cast<const __nv_bfloat16*>(g);
  if (thin_mma_enabled() && (k == 3 || k == 4 || k == 7) && static_cast<long>(B) * Ho * Wo < (1L << 31) &&
      (!g_split || (Ho % 2 == 0 && Wo % 2 == 0))) {
    const int T = k * k;
    TapWgradParams tp{};
    tp.src = x; tp.src_mask = xmask; tp.B = B; tp.H = H; tp.W = W; tp.Ho = Ho; tp.Wo = Wo;
    tp.S = s; tp.pad = pad; tp.flip = 0; tp.y_split = g_split;
    tp.total = static_cast<unsigned>(static_cast<long>(B) * Ho * Wo);
    tp.partial = partial; tp.partial_c = nullptr; tp.center = -1;
    const int TP = (T + 1) / 2 * 2;
    const int cap = static_cast<int>(static_cast<long>(rows_cap) * (T + 1) / (2 * TP + 1));   // same buffer, 2TP+1 rows per CTA
    TG_REQUIRE(cap >= 1, "tg_conv_c1_wgrad: rows_cap too small");
    int used = 0;
    TG_REQUIRE(tapwgrad_dispatch(k, tp, g, cap, &used, st) == 0, "tg_conv_c1_wgrad: launch failed");
    int8_t perm[49];
    for (int t = 0; t < T; ++t) perm[t] = static_cast<int8_t>(t);
    TG_REQUIRE(tapwgrad_reduce(k, partial, nullptr, used, dw, T, 1, perm, db, 1, accumulate, st) == 0,
               "tg_conv_c1_wgrad: reduce failed");
    return 0;
  }
  bool handled = false;
  TG_C1_DISPATCH(7, 2, (conv_c1_wgrad_kernel<K, S><<<grid, 256, 0, st>>>(x, xmask, B, H, W, pad, gg, Ho, Wo, g_split, partial)))
  TG_C1_DISPATCH(4, 2, (conv_c1_wgrad_kernel<K, S><<<grid, 256, 0, st>>>(x, xmask, B, H, W, pad, gg, Ho, Wo, g_split, partial)))
  TG_C1_DISPATCH(3, 1, (conv_c1_wgrad_kernel<K, S><<<grid, 256, 0, st>>>(x, xmask, B, H, W, pad, gg, Ho, Wo, g_split, partial)))
  TG_REQUIRE(handled, "tg_conv_c1_wgrad: unsupported window k=%d s=%d", k, s);
  TG_CHECK_CUDA(cudaGetLastError());
  const int T = k * k;
  conv_c1_wgrad_reduce_kernel<<<(64 * (T + 1) + 127) / 128, 128, 0, st>>>(partial, grid, T, dw, db, accumulate);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tg_conv_c1_wgrad_rows(void) { return tg::num_sms() * 2 + 8; }

extern "C" int tg_conv_to1_fwd(const void* x, int x_split, int B, int H, int W, int C, const float* wgt, int ncls,
                               const int* cls_count, const int8_t* tap_dh, const int8_t* tap_dw, const float* bias, int Ho,
                               int Wo, int mode, const uint8_t* mask, const float* xin, float* out, float* sig_out,
                               float* scratch, size_t scratch_floats, void* stream) {
  using namespace tg;
  TG_REQUIRE(x && wgt && out && cls_count && tap_dh && tap_dw, "tg_conv_to1_fwd: null pointer");
  TG_REQUIRE(C % 64 == 0, "tg_conv_to1_fwd: C=%d must be a multiple of 64", C);
  TG_REQUIRE(ncls == 1 || ncls == 4, "tg_conv_to1_fwd: ncls must be 1 or 4");
  TG_REQUIRE(mode == 0 || (mask && xin), "tg_conv_to1_fwd: composite mode needs mask and xin");
  To1Taps taps;
  TG_REQUIRE(fill_taps(&taps, ncls, cls_count, tap_dh, tap_dw) > 0, "tg_conv_to1_fwd: bad tap table");
  int ntaps_all = 0;
  for (int i = 0; i < ncls; ++i) ntaps_all += cls_count[i];
  if (thin_mma_enabled() && C == 64 && (scratch != nullptr || to1_fused_covers(x_split, H, W, taps, ntaps_all, Ho, Wo))) {
    const int ntaps = ntaps_all;
    const int rc = to1_fwd_mma(x, x_split, B, H, W, wgt, taps, ntaps, bias, Ho, Wo, mode, mask, xin, out, sig_out, scratch,
                               scratch_floats, reinterpret_cast<cudaStream_t>(stream));
    if (rc == 0) return 0;
    TG_REQUIRE(rc == -1, "tg_conv_to1_fwd: tensor-core path failed (%d)", rc);     // -1: shape not covered, fall through
  }
  if (ncls == 1 && !x_split && mode == 0 && wide_shape(C, cls_count[0], static_cast<long>(B) * H * W) && scratch != nullptr &&
      scratch_floats >= static_cast<size_t>(kWideT) * B * H * W) {
    const long total = static_cast<long>(B) * H * W;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    to1_wide_tapdot_kernel<<<dc_grid(total * 8, 256, 4), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), total, wgt, scratch);
    TG_CHECK_CUDA(cudaGetLastError());
    TG_REQUIRE(to1_tapsum_launch(scratch, total, 0, B, H, W, taps, bias, Ho, Wo, mode, mask, xin, out, sig_out, st) == 0,
               "tg_conv_to1_fwd: tap sum failed");
    return 0;
  }
  Tap3x3 tp;
  if (ncls == 1 && C == 64 && !x_split && Ho == H && Wo == W && make_tap3x3(&tp, cls_count[0], tap_dh, tap_dw)) {
    const long tiles = static_cast<long>(B) * ((H + kT1H - 1) / kT1H) * ((W + kT1W - 1) / kT1W);
    const int g3 = static_cast<int>(tiles < 4L * num_sms() ? tiles : 4L * num_sms());
    conv3x3_c64_to1_kernel<<<g3, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), B, H, W, wgt, tp, bias, mode, mask, xin, out, sig_out);
    TG_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  S2Taps s2;
  if (ncls == 4 && C == 64 && mode == 0 && bias == nullptr && Ho == 2 * H && Wo == 2 * W &&
      make_s2taps(&s2, cls_count, tap_dh, tap_dw)) {
    const long tiles = static_cast<long>(B) * ((H + kS2H - 1) / kS2H) * ((W + kT1W - 1) / kT1W);
    const int g3 = static_cast<int>(tiles < 4L * num_sms() ? tiles : 4L * num_sms());
    conv_s2_c64_to1_kernel<<<g3, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), x_split, B, H, W, wgt, s2, out);
    TG_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  const long M = static_cast<long>(B) * Ho * Wo;
  const int grid = dc_grid(M * 8, 256, 8);
  conv_to1_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), x_split, B, H, W, C, wgt, taps, bias, Ho, Wo, mode, mask, xin, out,
      sig_out);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tg_conv_to1_fwd_kernels(int x_split, int H, int W, int C, int ncls, const int* cls_count, const int8_t* tap_dh,
                                       const int8_t* tap_dw, int Ho, int Wo) {
  using namespace tg;
  if (cls_count == nullptr || tap_dh == nullptr || tap_dw == nullptr || (ncls != 1 && ncls != 4)) return 0;
  To1Taps taps;
  if (fill_taps(&taps, ncls, cls_count, tap_dh, tap_dw) <= 0) return 0;
  int ntaps = 0;
  for (int i = 0; i < ncls; ++i) ntaps += cls_count[i];
  return (C == 64 && to1_fused_covers(x_split, H, W, taps, ntaps, Ho, Wo)) ? 1 : 2;
}

extern "C" size_t tg_conv_to1_fwd_scratch_floats(int B, int H, int W, int C, int ntaps) {
  if (tg::wide_shape(C, ntaps, static_cast<long>(B) * H * W)) return static_cast<size_t>(ntaps) * B * H * W;
  if (C != 64 || ntaps < 1 || ntaps > 32) return 0;
  return static_cast<size_t>(ntaps) * B * H * W;
}

extern "C" int tg_conv_to1_bwd_data(const float* g, int B, int Ho, int Wo, const float* wgt, int ntaps,
                                    const int8_t* tap_dh, const int8_t* tap_dw, int H, int W, int C, void* dx,
                                    void* stream) {
  using namespace tg;
  TG_REQUIRE(g && wgt && dx && tap_dh && tap_dw && C % 8 == 0, "tg_conv_to1_bwd_data: bad arguments");
  To1Taps taps;
  TG_REQUIRE(fill_taps(&taps, 1, &ntaps, tap_dh, tap_dw) > 0, "tg_conv_to1_bwd_data: bad tap table");
  const long total = static_cast<long>(B) * H * W * (C / 8);
  if (wide_shape(C, ntaps, static_cast<long>(B) * H * W)) {
    const long px = static_cast<long>(B) * H * W;
    to1_wide_bwd_data_kernel<<<dc_grid(px, 16, 6), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        g, B, Ho, Wo, wgt, taps, H, W, reinterpret_cast<__nv_bfloat16*>(dx));
    TG_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  Tap3x3 tp;
  if (thin_mma_enabled() && C == 64 && Ho == H && Wo == W && make_tap3x3(&tp, ntaps, tap_dh, tap_dw) &&
      static_cast<long>(B) * H * W < (1L << 31)) {
    // dx[p][c] = sum_pos g[p - d_pos] * w[tap(pos)][c]: a "flipped" 3x3 1 -> 64 convolution of g
    RowGemmParams rp{};
    rp.src = g; rp.B = B; rp.H = H; rp.W = W; rp.Ho = H; rp.Wo = W;
    rp.S = 1; rp.pad = 1; rp.flip = 1;
    rp.wgt = wgt; rp.w_sn = 1; rp.w_st = 64;
    for (int t = 0; t < 9; ++t) rp.perm[t] = static_cast<int8_t>(tp.idx[t]);
    rp.out = reinterpret_cast<__nv_bfloat16*>(dx);
    rp.total = static_cast<unsigned>(static_cast<long>(B) * H * W);
    TG_REQUIRE(rowgemm_dispatch(3, rp, 0, nullptr, reinterpret_cast<cudaStream_t>(stream)) == 0,
               "tg_conv_to1_bwd_data: launch failed");
    return 0;
  }
  if (C == 64 && Ho == H && Wo == W && make_tap3x3(&tp, ntaps, tap_dh, tap_dw)) {
    conv3x3_c64_to1_bwd_data_kernel<<<dc_grid(total, 256, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        g, B, H, W, wgt, tp, reinterpret_cast<__nv_bfloat16*>(dx));
    TG_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  conv_to1_bwd_data_kernel<__nv_bfloat16><<<dc_grid(total, 256, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      g, B, Ho, Wo, wgt, taps, H, W, C, reinterpret_cast<__nv_bfloat16*>(dx));
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tg_conv_to1_wgrad(const void* x, int B, int H, int W, int C, const float* g, int Ho, int Wo, int ntaps,
                                 const int8_t* tap_dh, const int8_t* tap_dw, float* partial, float* partial_b,
                                 int rows_cap, float* dw, float* db, int accumulate, void* stream) {
  using namespace tg;
  TG_REQUIRE(x && g && partial && dw && tap_dh && tap_dw, "tg_conv_to1_wgrad: null pointer");
  TG_REQUIRE(C % 64 == 0, "tg_conv_to1_wgrad: C must be a multiple of 64");
  To1Taps taps;
  TG_REQUIRE(fill_taps(&taps, 1, &ntaps, tap_dh, tap_dw) > 0, "tg_conv_to1_wgrad: bad tap table");
  const long M = static_cast<long>(B) * Ho * Wo;
  int gx = dc_grid(M, 16, 4);
  const int slabs = C / 64;
  if (gx * slabs > 8 * num_sms()) gx = (8 * num_sms() + slabs - 1) / slabs;
  if (gx > rows_cap) gx = rows_cap;
  TG_REQUIRE(gx >= 1, "tg_conv_to1_wgrad: rows_cap must be >= 1");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const __nv_bfloat16* xx = reinterpret_cast<const __nv_bfloat16*>(x);
  if (wide_shape(C, ntaps, static_cast<long>(B) * H * W)) {
    int gw = 4 * num_sms();
    if (gw > rows_cap) gw = rows_cap;
    const long px = static_cast<long>(B) * H * W;
    if (gw > px) gw = static_cast<int>(px);
    constexpr int kSmem = kWideT * kWideC * 4 + 64;
    TG_SET_SMEM_ONCE((to1_wide_wgrad_kernel), kSmem);
    to1_wide_wgrad_kernel<<<gw, 256, kSmem, st>>>(xx, B, H, W, g, Ho, Wo, taps, partial, partial_b);
    TG_CHECK_CUDA(cudaGetLastError());
    conv_to1_wgrad_reduce_kernel<<<(ntaps * C + 31) / 32, 256, 0, st>>>(partial, partial_b, gw, ntaps, C, dw, db, accumulate);
    TG_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  Tap3x3 tp;
  if (thin_mma_enabled() && C == 64 && Ho == H && Wo == W && make_tap3x3(&tp, ntaps, tap_dh, tap_dw) &&
      static_cast<long>(B) * H * W < (1L << 31) && rows_cap * 9 / 21 >= 1) {
    // dw[pos][c] = sum_n g[n - d_pos] * x[n][c]: the "flipped" 3x3 taps of g against the activation tile
    TapWgradParams wp{};
    wp.src = g; wp.B = B; wp.H = H; wp.W = W; wp.Ho = H; wp.Wo = W;
    wp.S = 1; wp.pad = 1; wp.flip = 1; wp.y_split = 0;
    wp.total = static_cast<unsigned>(static_cast<long>(B) * H * W);
    wp.partial = partial; wp.partial_c = partial_b; wp.center = 4;
    int used = 0;
    TG_REQUIRE(tapwgrad_dispatch(3, wp, x, rows_cap * 9 / 21, &used, st) == 0, "tg_conv_to1_wgrad: launch failed");
    int8_t perm[9];
    for (int t = 0; t < 9; ++t) perm[t] = static_cast<int8_t>(tp.idx[t]);
    TG_REQUIRE(tapwgrad_reduce(3, partial, partial_b, used, dw, 9, 1, perm, db, partial_b ? 2 : 0, accumulate, st) == 0,
               "tg_conv_to1_wgrad: reduce failed");
    return 0;
  }
  if (C == 64 && Ho == H && Wo == W && make_tap3x3(&tp, ntaps, tap_dh, tap_dw)) {
    const long tiles = static_cast<long>(B) * ((H + kT1H - 1) / kT1H) * ((W + kT1W - 1) / kT1W);
    int g3 = static_cast<int>(tiles < 4L * num_sms() ? tiles : 4L * num_sms());
    if (g3 > rows_cap) g3 = rows_cap;
    conv3x3_c64_to1_wgrad_kernel<<<g3, 256, 0, st>>>(xx, B, H, W, g, tp, partial, partial_b);
    TG_CHECK_CUDA(cudaGetLastError());
    conv_to1_wgrad_reduce_kernel<<<(ntaps * C + 31) / 32, 256, 0, st>>>(partial, partial_b, g3, ntaps, C, dw, db,
                                                                         accumulate);
    TG_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  dim3 grid(gx, slabs);
  if (ntaps == 9) conv_to1_wgrad_kernel<9><<<grid, 128, 0, st>>>(xx, B, H, W, C, g, Ho, Wo, taps, partial, partial_b);
  else if (ntaps == 16) conv_to1_wgrad_kernel<16><<<grid, 128, 0, st>>>(xx, B, H, W, C, g, Ho, Wo, taps, partial, partial_b);
  else TG_REQUIRE(false, "tg_conv_to1_wgrad: unsupported tap count %d (supported: 9, 16)", ntaps);
  TG_CHECK_CUDA(cudaGetLastError());
  conv_to1_wgrad_reduce_kernel<<<(ntaps * C + 31) / 32, 256, 0, st>>>(partial, partial_b, gx, ntaps, C, dw, db,
                                                                       accumulate);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tg_conv_to1_wgrad_rows(void) { return tg::num_sms() * 8; }

// ------------------------------------------------------------------------------------------------
// fp32-storage twins (verification path): the generic CUDA-core kernels above with fp32 activations.
// ------------------------------------------------------------------------------------------------
extern "C" int tg_conv_c1_fwd_f32(const float* x, const uint8_t* xmask, int B, int H, int W, int k, int s, int pad,
                                  const float* wgt, const float* bias, const uint8_t* code, const float* lut_dev, int act,
                                  float slope, void* out, int out_split, float* stats, int stats_rows_cap,
                                  int* stats_rows_used, void* stream) {
  using namespace tg;
  TG_REQUIRE(x && wgt && bias && out, "tg_conv_c1_fwd_f32: null pointer");
  TG_REQUIRE(!code || lut_dev, "tg_conv_c1_fwd_f32: code needs a device LUT");
  const int Ho = (H + 2 * pad - k) / s + 1, Wo = (W + 2 * pad - k) / s + 1;
  TG_REQUIRE(!out_split || (Ho % 2 == 0 && Wo % 2 == 0), "tg_conv_c1_fwd_f32: parity-split output needs even Ho, Wo");
  const long tiles = static_cast<long>(B) * ((Ho + kC1TH - 1) / kC1TH) * ((Wo + kC1TW - 1) / kC1TW);
  int grid = static_cast<int>(tiles < 4L * num_sms() ? tiles : 4L * num_sms());
  if (stats) {
    TG_REQUIRE(stats_rows_used != nullptr, "tg_conv_c1_fwd_f32: stats_rows_used is null");
    if (grid > stats_rows_cap) grid = stats_rows_cap;
    TG_REQUIRE(grid >= 1, "tg_conv_c1_fwd_f32: stats_rows_cap must be >= 1");
    *stats_rows_used = grid;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* o = reinterpret_cast<float*>(out);
  bool handled = false;
  TG_C1_DISPATCH(7, 2, (conv_c1_fwd_kernel<K, S, float><<<grid, 128, 0, st>>>(x, xmask, B, H, W, pad, wgt, bias, Ho, Wo, code, lut_dev, act, slope, o, out_split, stats)))
  TG_C1_DISPATCH(4, 2, (conv_c1_fwd_kernel<K, S, float><<<grid, 128, 0, st>>>(x, xmask, B, H, W, pad, wgt, bias, Ho, Wo, code, lut_dev, act, slope, o, out_split, stats)))
  TG_C1_DISPATCH(3, 1, (conv_c1_fwd_kernel<K, S, float><<<grid, 128, 0, st>>>(x, xmask, B, H, W, pad, wgt, bias, Ho, Wo, code, lut_dev, act, slope, o, out_split, stats)))
  TG_REQUIRE(handled, "tg_conv_c1_fwd_f32: unsupported window k=%d s=%d (supported: 7/2, 4/2, 3/1)", k, s);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tg_conv_c1_wgrad_f32(const float* x, const uint8_t* xmask, int B, int H, int W, int k, int s, int pad,
                                    const void* g, int g_split, float* partial, int rows_cap, float* dw, float* db,
                                    int accumulate, void* stream) {
  using namespace tg;
  TG_REQUIRE(x && g && partial && dw, "tg_conv_c1_wgrad_f32: null pointer");
  const int Ho = (H + 2 * pad - k) / s + 1, Wo = (W + 2 * pad - k) / s + 1;
  const long tiles = static_cast<long>(B) * ((Ho + kC1TH - 1) / kC1TH) * ((Wo + kC1TW - 1) / kC1TW);
  int grid = static_cast<int>(tiles < 2L * num_sms() ? tiles : 2L * num_sms());
  if (grid > rows_cap) grid = rows_cap;
  TG_REQUIRE(grid >= 1, "tg_conv_c1_wgrad_f32: rows_cap must be >= 1");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const float* gg = reinterpret_cast<const float*>(g);
  bool handled = false;
  TG_C1_DISPATCH(7, 2, (conv_c1_wgrad_kernel<K, S, float><<<grid, 256, 0, st>>>(x, xmask, B, H, W, pad, gg, Ho, Wo, g_split, partial)))
  TG_C1_DISPATCH(4, 2, (conv_c1_wgrad_kernel<K, S, float><<<grid, 256, 0, st>>>(x, xmask, B, H, W, pad, gg, Ho, Wo, g_split, partial)))
  TG_C1_DISPATCH(3, 1, (conv_c1_wgrad_kernel<K, S, float><<<grid, 256, 0, st>>>(x, xmask, B, H, W, pad, gg, Ho, Wo, g_split, partial)))
  TG_REQUIRE(handled, "tg_conv_c1_wgrad_f32: unsupported window k=%d s=%d", k, s);
  TG_CHECK_CUDA(cudaGetLastError());
  const int T = k * k;
  conv_c1_wgrad_reduce_kernel<<<(64 * (T + 1) + 127) / 128, 128, 0, st>>>(partial, grid, T, dw, db, accumulate);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tg_conv_to1_fwd_f32(const void* x, int x_split, int B, int H, int W, int C, const float* wgt, int ncls,
                                   const int* cls_count, const int8_t* tap_dh, const int8_t* tap_dw, const float* bias,
                                   int Ho, int Wo, int mode, const uint8_t* mask, const float* xin, float* out,
                                   float* sig_out, void* stream) {
  using namespace tg;
  TG_REQUIRE(x && wgt && out && cls_count && tap_dh && tap_dw, "tg_conv_to1_fwd_f32: null pointer");
  TG_REQUIRE(C % 64 == 0, "tg_conv_to1_fwd_f32: C=%d must be a multiple of 64", C);
  TG_REQUIRE(ncls == 1 || ncls == 4, "tg_conv_to1_fwd_f32: ncls must be 1 or 4");
  TG_REQUIRE(mode == 0 || (mask && xin), "tg_conv_to1_fwd_f32: composite mode needs mask and xin");
  To1Taps taps;
  TG_REQUIRE(fill_taps(&taps, ncls, cls_count, tap_dh, tap_dw) > 0, "tg_conv_to1_fwd_f32: bad tap table");
  const long M = static_cast<long>(B) * Ho * Wo;
  conv_to1_fwd_kernel<float><<<dc_grid(M * 8, 256, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float*>(x), x_split, B, H, W, C, wgt, taps, bias, Ho, Wo, mode, mask, xin, out, sig_out);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tg_conv_to1_bwd_data_f32(const float* g, int B, int Ho, int Wo, const float* wgt, int ntaps,
                                        const int8_t* tap_dh, const int8_t* tap_dw, int H, int W, int C, void* dx,
                                        void* stream) {
  using namespace tg;
  TG_REQUIRE(g && wgt && dx && tap_dh && tap_dw && C % 8 == 0, "tg_conv_to1_bwd_data_f32: bad arguments");
  To1Taps taps;
  TG_REQUIRE(fill_taps(&taps, 1, &ntaps, tap_dh, tap_dw) > 0, "tg_conv_to1_bwd_data_f32: bad tap table");
  const long total = static_cast<long>(B) * H * W * (C / 8);
  conv_to1_bwd_data_kernel<float><<<dc_grid(total, 256, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      g, B, Ho, Wo, wgt, taps, H, W, C, reinterpret_cast<float*>(dx));
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int tg_conv_to1_wgrad_f32(const void* x, int B, int H, int W, int C, const float* g, int Ho, int Wo,
                                     int ntaps, const int8_t* tap_dh, const int8_t* tap_dw, float* partial,
                                     float* partial_b, int rows_cap, float* dw, float* db, int accumulate,
                                     void* stream) {
  using namespace tg;
  TG_REQUIRE(x && g && partial && dw && tap_dh && tap_dw, "tg_conv_to1_wgrad_f32: null pointer");
  TG_REQUIRE(C % 64 == 0, "tg_conv_to1_wgrad_f32: C must be a multiple of 64");
  To1Taps taps;
  TG_REQUIRE(fill_taps(&taps, 1, &ntaps, tap_dh, tap_dw) > 0, "tg_conv_to1_wgrad_f32: bad tap table");
  const long M = static_cast<long>(B) * Ho * Wo;
  int gx = dc_grid(M, 16, 4);
  const int slabs = C / 64;
  if (gx * slabs > 8 * num_sms()) gx = (8 * num_sms() + slabs - 1) / slabs;
  if (gx > rows_cap) gx = rows_cap;
  TG_REQUIRE(gx >= 1, "tg_conv_to1_wgrad_f32: rows_cap must be >= 1");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const float* xx = reinterpret_cast<const float*>(x);
  dim3 grid(gx, slabs);
  if (ntaps == 9) conv_to1_wgrad_kernel<9, float><<<grid, 128, 0, st>>>(xx, B, H, W, C, g, Ho, Wo, taps, partial, partial_b);
  else if (ntaps == 16) conv_to1_wgrad_kernel<16, float><<<grid, 128, 0, st>>>(xx, B, H, W, C, g, Ho, Wo, taps, partial, partial_b);
  else TG_REQUIRE(false, "tg_conv_to1_wgrad_f32: unsupported tap count %d (supported: 9, 16)", ntaps);
  TG_CHECK_CUDA(cudaGetLastError());
  conv_to1_wgrad_reduce_kernel<<<(ntaps * C + 31) / 32, 256, 0, st>>>(partial, partial_b, gx, ntaps, C, dw, db,
                                                                       accumulate);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}
