// conv_epilogue.cuh — epilogue of one 128-pixel x BN-channel accumulator tile, shared by the
// implicit-GEMM convolution kernels (conv_igemm.cu: per-tap TMA boxes; conv_halo.cu: halo-tile reuse).
// One thread = one output pixel (TMEM lane); columns are drained 32 at a time with tcgen05.ld.
#pragma once
#include "conv_igemm.cuh"
#include "tg_common.cuh"
#include <type_traits>

namespace tg {

// s_vec: [3][kVec] bias / scale / shift staged in shared memory; my_stats: this warp's [2][kVec] accumulators.
// stage: this warp's private 32 x (kSC*2)-byte staging tile. Each thread owns one pixel row, so a direct
// store would touch 32 different 128-byte lines per instruction (the LSU serialises them: measured 4x the
// MMA time on the N=64 layers). Rows are therefore written to shared memory (XOR-swizzled, conflict-free)
// and read back transposed so that every global store instruction writes full contiguous lines.
// MODE: compile-time feature set of the launch so that the per-tile instruction stream only holds what the layer
// uses (a warp's epilogue is one dependent chain; with every variant predicated at run time it was ~500
// instructions per 32x32 chunk and capped the N=64 layers). Bits: 1 ratio/mask code, 2 BN statistics,
// 4 activation-derivative gate, 8 affine (eval-mode BN), 16 per-channel vectors present (bias / scale / shift).
// MODE < 0: everything decided at run time.
constexpr int kEpiCode = 1, kEpiStats = 2, kEpiGate = 4, kEpiAffine = 8, kEpiVec = 16;
__host__ __device__ inline int conv_epilogue_mode(const void* code, const void* stats, const void* gate, const void* scale,
                                                  const void* shift, const void* bias) {
  const int m = (code ? kEpiCode : 0) | (stats ? kEpiStats : 0) | (gate ? kEpiGate : 0) |
                ((scale || shift) ? kEpiAffine : 0) | ((bias || scale || shift) ? kEpiVec : 0);
  switch (m) {
    case 0: case kEpiCode: case kEpiGate: case kEpiVec: case kEpiVec | kEpiCode: case kEpiVec | kEpiStats:
    case kEpiVec | kEpiCode | kEpiStats:
    case kEpiVec | kEpiAffine: case kEpiVec | kEpiCode | kEpiAffine:
      return m;
    default:
      return -1;
  }
}
// calls f(std::integral_constant<int, MODE>) for the specialised modes, MODE = -1 otherwise
template <typename F>
__device__ __forceinline__ void conv_epilogue_dispatch(int mode, F f) {
  switch (mode) {
    case 0: f(std::integral_constant<int, 0>{}); break;
    case kEpiCode: f(std::integral_constant<int, kEpiCode>{}); break;
    case kEpiGate: f(std::integral_constant<int, kEpiGate>{}); break;
    case kEpiVec: f(std::integral_constant<int, kEpiVec>{}); break;
    case kEpiVec | kEpiCode: f(std::integral_constant<int, kEpiVec | kEpiCode>{}); break;
    case kEpiVec | kEpiStats: f(std::integral_constant<int, kEpiVec | kEpiStats>{}); break;
    case kEpiVec | kEpiCode | kEpiStats: f(std::integral_constant<int, kEpiVec | kEpiCode | kEpiStats>{}); break;
    case kEpiVec | kEpiAffine: f(std::integral_constant<int, kEpiVec | kEpiAffine>{}); break;
    case kEpiVec | kEpiCode | kEpiAffine: f(std::integral_constant<int, kEpiVec | kEpiCode | kEpiAffine>{}); break;
    default: f(std::integral_constant<int, -1>{}); break;
  }
}

// Per-tile inputs of the epilogue that come from global memory (ratio / mask code, gate row). They only depend on
// the tile coordinates, so the callers fetch them BEFORE waiting for the accumulator: the loads then overlap the
// MMAs of the tile instead of sitting on the epilogue's dependent chain.
struct EpiPrefetch {
  float rs;
  uint4 g[4];
};
template <int BN, int MODE, bool kBox8 = false>
__device__ __forceinline__ void conv_epilogue_prefetch(const ConvKParams& p, int q, int lane, int nt, int sb, int tw, int th,
                                                       int tb, int hsel, int wt, int ht, int bt, EpiPrefetch& pre) {
  constexpr bool kGen = MODE < 0;
  const bool f_code = kGen ? (p.code != nullptr) : ((MODE & kEpiCode) != 0);
  const bool f_gate = kGen ? (p.gate != nullptr) : ((MODE & kEpiGate) != 0);
  const int w = tw * p.Wt + wt, h = th * p.Ht + ht, b = tb * p.Bt + bt;
  const bool valid = (w < p.Wo) && (h < p.Ho) && (b < p.B);
  const long pix = ((static_cast<long>(b) * p.Po + p.sub[sb].out_plane) * p.Ho + h) * p.Wo + w;
  pre.rs = 1.f;
  if (f_code && valid) pre.rs = p.lut[p.code[pix]];
  if (kBox8 && BN == 64 && f_gate) {
    // N = 64 halo kernels (8 x 16 pixel boxes, always inside the image): fetch the gate in the layout of the
    // transposed output store -- lane = (row i*8 + lane/4, 16-byte chunk lane%4) -- so that every load instruction
    // reads 8 rows x 64 contiguous bytes instead of 32 rows x 16 bytes
    const long tile_pix = ((static_cast<long>(tb) * p.Po + p.sub[sb].out_plane) * p.Ho + th * p.Ht) * p.Wo + tw * p.Wt;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rr = q * 32 + i * 8 + (lane >> 2);
      const long px = tile_pix + static_cast<long>(rr >> 3) * p.Wo + (rr & 7);
      const uint4* src = reinterpret_cast<const uint4*>(p.gate + px * p.Cout + nt * BN + hsel * 32 + (lane & 3) * 8);
      pre.g[i] = TG_DBG(p, 16) ? make_uint4(0x3f803f80u, 0x3f803f80u, 0xbf80bf80u, 0x3f803f80u) : *src;
    }
  }
}

template <int BN, int kVec, int kSC, int MODE, bool kBox8 = false>
__device__ __forceinline__ void conv_epilogue_tile(const ConvKParams& p, int q, int lane, int nt, int sb, int tw,
                                                   int th, int tb, uint32_t t_addr, const float* s_vec,
                                                   float* my_stats, bool has_vec, uint8_t* stage, int hsel,
                                                   int wt, int ht, int bt, float* racc = nullptr,
                                                   const EpiPrefetch* pre = nullptr) {
  static_assert(kSC == 32 || kSC == 64, "staging width is 32 or 64 columns");
  constexpr int kLPR = kSC / 8;           // 16-byte chunks (= lanes) per staged row
  constexpr int kRowBytes = kSC * 2;
  constexpr int kChunksPerStage = kSC / 32;
  // (wt, ht, bt): position of this thread's row (TMEM lane q*32 + lane) inside the pixel box. They depend on the
  // thread only, so the callers compute them once per kernel: the three runtime integer divisions sat at the head
  // of every tile's dependent chain (the epilogue is latency-bound per warp, not throughput-bound).
  __builtin_assume(__isShared(s_vec));
  __builtin_assume(__isShared(my_stats));
  __builtin_assume(__isShared(stage));
  constexpr bool kGen = MODE < 0;
  const bool f_code = kGen ? (p.code != nullptr) : ((MODE & kEpiCode) != 0);
  const bool f_stats = kGen ? (p.stats != nullptr) : ((MODE & kEpiStats) != 0);
  const bool f_gate = kGen ? (p.gate != nullptr) : ((MODE & kEpiGate) != 0);
  const bool f_vec = kGen ? has_vec : ((MODE & kEpiVec) != 0);
  const int w = tw * p.Wt + wt, h = th * p.Ht + ht, b = tb * p.Bt + bt;
  const bool valid = (w < p.Wo) && (h < p.Ho) && (b < p.B);
  const long pix =
      ((static_cast<long>(b) * p.Po + p.sub[sb].out_plane) * p.Ho + h) * p.Wo + w;
  float rs = 1.f;
  if (f_code && valid) rs = pre ? pre->rs : p.lut[p.code[pix]];
  const unsigned vmask = kBox8 ? 0xffffffffu : __ballot_sync(0xffffffffu, valid);
  const long box_pix = ((static_cast<long>(tb) * p.Po + p.sub[sb].out_plane) * p.Ho + th * p.Ht) * p.Wo + tw * p.Wt;
  const bool has_affine = kGen ? (p.scale != nullptr || p.shift != nullptr) : ((MODE & kEpiAffine) != 0);
  const int sw_w = (kSC == 64) ? (lane & 7) : ((lane >> 1) & 3);     // write-side swizzle key of this row
  const __nv_bfloat16* grow = f_gate ? p.gate + pix * p.Cout + nt * BN : nullptr;
  const bool relu_gate = f_gate && p.gate_slope == 0.f && p.act == 0;
  const bool gate_t = kBox8 && BN == 64 && kSC == 32 && f_gate && pre != nullptr;   // gate prefetched in store layout

  // two warps share each TMEM lane quarter: warp `hsel` takes every other kSC-column group
#pragma unroll 1
  for (int ch = 0; ch < BN / 32; ++ch) {
    if (((ch / kChunksPerStage) & 1) != hsel) continue;
    uint32_t raw[32];
    if (!TG_DBG(p, 8)) {
      tmem_ld_32x32(t_addr + ch * 32, raw);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) raw[j] = 0x3f800000u + lane;
    }
    const int n0 = nt * BN + ch * 32;
    float v[32];
    if (f_vec) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 bv = *reinterpret_cast<const float4*>(s_vec + n0 + j);
        v[j] = __uint_as_float(raw[j]) + bv.x;
        v[j + 1] = __uint_as_float(raw[j + 1]) + bv.y;
        v[j + 2] = __uint_as_float(raw[j + 2]) + bv.z;
        v[j + 3] = __uint_as_float(raw[j + 3]) + bv.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
    }
    // rows outside the image must contribute 0 to the statistics; elsewhere they are simply not stored
    if (f_code || f_stats) {
      const float rsv = valid ? rs : 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= rsv;
    }
    if (f_stats) {
      if (BN == 64 && racc != nullptr) {
        // one chunk per warp: per-thread running sums over all tiles (64 registers, independent FMAs); the caller
        // reduces them across lanes once at the end of the kernel instead of 62 dependent shuffles per tile
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          racc[j] += v[j];
          racc[32 + j] = fmaf(v[j], v[j], racc[32 + j]);
        }
      } else {
        float sq[32], sm[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          sm[j] = v[j];
          sq[j] = v[j] * v[j];
        }
        const float csum = warp_transpose_sum32(sm);
        const float csq = warp_transpose_sum32(sq);
        my_stats[n0 + lane] += csum;
        my_stats[kVec + n0 + lane] += csq;
      }
    }
    uint32_t packed[16];
    uint32_t gbits[16];
    if (f_gate && valid && !gate_t) {
      const uint4* gsrc = reinterpret_cast<const uint4*>(grow + ch * 32);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 t = TG_DBG(p, 16) ? make_uint4(0x3f803f80u, 0x3f803f80u, 0xbf80bf80u, 0x3f803f80u) : gsrc[j];
        gbits[4 * j] = t.x; gbits[4 * j + 1] = t.y; gbits[4 * j + 2] = t.z; gbits[4 * j + 3] = t.w;
      }
    }
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      float a = v[j], c = v[j + 1];
      if (has_affine) {
        const float2 sc2 = *reinterpret_cast<const float2*>(s_vec + kVec + n0 + j);
        const float2 sh2 = *reinterpret_cast<const float2*>(s_vec + 2 * kVec + n0 + j);
        a = a * sc2.x + sh2.x;
        c = c * sc2.y + sh2.y;
      }
      if (f_gate && valid && !relu_gate && !gate_t) {
        // derivative of LeakyReLU of the tensor this gradient flows into
        const float2 gv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gbits[j >> 1]));
        if (!(gv.x > 0.f)) a *= p.gate_slope;
        if (!(gv.y > 0.f)) c *= p.gate_slope;
      }
      if (p.act == 1) {
        a = fmaxf(a, 0.f);
        c = fmaxf(c, 0.f);
      } else if (p.act == 2) {
        a = a > 0.f ? a : a * p.slope;
        c = c > 0.f ? c : c * p.slope;
      }
      uint32_t pk = pack_bf16x2(a, c);
      if (f_gate && valid && relu_gate && !gate_t) {
        // ReLU derivative (slope 0, no activation after it): multiply the packed result by [gate > 0] in {1, 0};
        // exact, two packed instructions per element pair instead of unpack / compare / multiply per element
        const __nv_bfloat162 on = __hgt2(*reinterpret_cast<const __nv_bfloat162*>(&gbits[j >> 1]),
                                         __floats2bfloat162_rn(0.f, 0.f));
        const __nv_bfloat162 r2 = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&pk), on);
        pk = *reinterpret_cast<const uint32_t*>(&r2);
      }
      packed[j >> 1] = pk;
    }
    {
      const int cbase = (ch % kChunksPerStage) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(stage + lane * kRowBytes + (((cbase + j) ^ sw_w) << 4)) =
            make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
    }
    if ((ch % kChunksPerStage) == kChunksPerStage - 1 && !TG_DBG(p, 32)) {
      __syncwarp();
      const int col0 = nt * BN + (ch - (kChunksPerStage - 1)) * 32;   // first channel held by the staging tile
#pragma unroll
      for (int i = 0; i < kLPR; ++i) {
        const int row = i * (32 / kLPR) + lane / kLPR;
        const int chunk = lane % kLPR;
        const int key = (kSC == 64) ? (row & 7) : ((row >> 1) & 3);
        uint4 val = *reinterpret_cast<const uint4*>(stage + row * kRowBytes + ((chunk ^ key) << 4));
        if (gate_t) {
          // derivative of (Leaky)ReLU on the packed bf16 pairs: factor 1 where gate > 0, else the slope
          const uint4 g4 = pre->g[i & 3];
          const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
          const uint32_t gw[4] = {g4.x, g4.y, g4.z, g4.w};
          uint32_t vw[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const __nv_bfloat162 on = __hgt2(*reinterpret_cast<const __nv_bfloat162*>(&gw[e]), zero2);
            const __nv_bfloat162 x2 = *reinterpret_cast<const __nv_bfloat162*>(&vw[e]);
            __nv_bfloat162 r2;
            if (p.gate_slope == 0.f) {
              r2 = __hmul2(x2, on);
            } else {   // LeakyReLU: fp32 multiply by the slope, rounded once; per-half select through a bit mask
              const float2 xf = __bfloat1622float2(x2);
              const __nv_bfloat162 s2 = __floats2bfloat162_rn(xf.x * p.gate_slope, xf.y * p.gate_slope);
              const uint32_t m = __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&gw[e]), zero2);
              const uint32_t bits = (vw[e] & m) | (*reinterpret_cast<const uint32_t*>(&s2) & ~m);
              r2 = *reinterpret_cast<const __nv_bfloat162*>(&bits);
            }
            vw[e] = *reinterpret_cast<const uint32_t*>(&r2);
          }
          val = make_uint4(vw[0], vw[1], vw[2], vw[3]);
        }
        if (kBox8) {
          // 8 x 16 pixel boxes that always lie inside the image: the row's pixel follows from the tile origin
          // (no shuffles, no validity mask on the store path)
          const int rr = q * 32 + row;
          const long pix2 = box_pix + static_cast<long>(rr >> 3) * p.Wo + (rr & 7);
          if (!TG_DBG(p, 1) && !p.skip_out) *reinterpret_cast<uint4*>(p.out + pix2 * p.Cout + col0 + chunk * 8) = val;
        } else {
          // pixel index of the row this lane stores: held by lane `row` of the warp
          const unsigned lo = __shfl_sync(0xffffffffu, static_cast<unsigned>(pix & 0xffffffffu), row);
          const unsigned hi = __shfl_sync(0xffffffffu, static_cast<unsigned>(static_cast<unsigned long long>(pix) >> 32), row);
          if (((vmask >> row) & 1u) && !TG_DBG(p, 1)) {
            const long pix2 = static_cast<long>((static_cast<unsigned long long>(hi) << 32) | lo);
            *reinterpret_cast<uint4*>(p.out + pix2 * p.Cout + col0 + chunk * 8) = val;
          }
        }
      }
      if (kBox8 && p.pool_out != nullptr) {
        // fused nn.MaxPool2d(2, 2) (VGG16 features[4] / [9], losses.py:31-32): this warp's staging tile holds 4 tile rows x 8
        // columns of pixels (row = 8 * tile_row + col), i.e. 2 x 4 pooled pixels; every lane takes one (pooled pixel,
        // 16-byte channel chunk) and reduces its four source rows straight out of shared memory
        constexpr int kItems = 8 * kLPR;                     // pooled pixels x chunks held by the tile
#pragma unroll
        for (int it = 0; it < kItems / 32; ++it) {
          const int item = it * 32 + lane;
          const int pp = item / kLPR, chunk = item % kLPR;
          const int r00 = (pp >> 2) * 16 + (pp & 3) * 2;
          uint4 m4 = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
          for (int s4 = 0; s4 < 4; ++s4) {
            const int srow = r00 + (s4 & 1) + (s4 >> 1) * 8;
            const int key = (kSC == 64) ? (srow & 7) : ((srow >> 1) & 3);
            const uint4 v4 = *reinterpret_cast<const uint4*>(stage + srow * kRowBytes + ((chunk ^ key) << 4));
            if (s4 == 0) {
              m4 = v4;
            } else {
              const uint32_t a[4] = {m4.x, m4.y, m4.z, m4.w}, b[4] = {v4.x, v4.y, v4.z, v4.w};
              uint32_t o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const __nv_bfloat162 r2 = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a[e]),
                                                  *reinterpret_cast<const __nv_bfloat162*>(&b[e]));
                o[e] = *reinterpret_cast<const uint32_t*>(&r2);
              }
              m4 = make_uint4(o[0], o[1], o[2], o[3]);
            }
          }
          const int ph = (th * p.Ht >> 1) + 2 * q + (pp >> 2), pw = (tw * p.Wt >> 1) + (pp & 3);
          const long ppix = (static_cast<long>(tb) * (p.Ho >> 1) + ph) * (p.Wo >> 1) + pw;
          *reinterpret_cast<uint4*>(p.pool_out + ppix * p.Cout + col0 + chunk * 8) = m4;
        }
      }
      __syncwarp();
    }
  }
}

// fp32-storage epilogue of the verification path (kind::tf32 MMAs, fp32 activations): same arithmetic as
// conv_epilogue_tile with every feature decided at run time and plain per-row stores (it need not be fast; it has
// to be the same formula at fp32 precision). `p.out` / `p.gate` are fp32 tensors here.
template <int BN, int kVec>
__device__ __forceinline__ void conv_epilogue_tile_f32(const ConvKParams& p, int q, int lane, int nt, int sb, int tw,
                                                       int th, int tb, uint32_t t_addr, const float* s_vec,
                                                       float* my_stats, bool has_vec, int hsel, int wt, int ht,
                                                       int bt) {
  const int w = tw * p.Wt + wt, h = th * p.Ht + ht, b = tb * p.Bt + bt;
  const bool valid = (w < p.Wo) && (h < p.Ho) && (b < p.B);
  const long pix = ((static_cast<long>(b) * p.Po + p.sub[sb].out_plane) * p.Ho + h) * p.Wo + w;
  float rs = 1.f;
  if (p.code != nullptr && valid) rs = p.lut[p.code[pix]];
  const bool has_affine = p.scale != nullptr || p.shift != nullptr;
  float* out = reinterpret_cast<float*>(p.out);
  const float* gate = reinterpret_cast<const float*>(p.gate);
#pragma unroll 1
  for (int ch = 0; ch < BN / 32; ++ch) {
    if ((ch & 1) != hsel) continue;
    uint32_t raw[32];
    tmem_ld_32x32(t_addr + ch * 32, raw);
    tmem_ld_wait();
    const int n0 = nt * BN + ch * 32;
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
    if (p.addend != nullptr && valid) {
      const float* ar = p.addend + pix * p.Cout + n0;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 t = *reinterpret_cast<const float4*>(ar + j);
        v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
      }
    }
    if (has_vec) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += s_vec[n0 + j];
    }
    if (p.code != nullptr || p.stats != nullptr) {
      const float rsv = valid ? rs : 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= rsv;
    }
    if (p.stats != nullptr) {
      float sq[32], sm[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        sm[j] = v[j];
        sq[j] = v[j] * v[j];
      }
      const float csum = warp_transpose_sum32(sm);
      const float csq = warp_transpose_sum32(sq);
      my_stats[n0 + lane] += csum;
      my_stats[kVec + n0 + lane] += csq;
    }
    if (valid) {
      float* orow = out + pix * p.Cout + n0;
      const float* grow = gate ? gate + pix * p.Cout + n0 : nullptr;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float a = v[j + e];
          if (has_affine) a = a * s_vec[kVec + n0 + j + e] + s_vec[2 * kVec + n0 + j + e];
          if (grow != nullptr && !(grow[j + e] > 0.f)) a *= p.gate_slope;
          if (p.act == 1) a = fmaxf(a, 0.f);
          else if (p.act == 2) a = a > 0.f ? a : a * p.slope;
          o[e] = a;
        }
        *reinterpret_cast<float4*>(orow + j) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

}  // namespace tg
