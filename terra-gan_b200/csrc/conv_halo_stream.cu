// conv_halo_stream.cu — 3x3 / stride-1 implicit-GEMM convolution with halo-tile reuse for layers whose packed
// weights do NOT fit in shared memory (conv_halo.cu keeps them resident): dec2/dec3, VGG conv7 and the data
// gradients of the 128/256-channel 3x3 layers (reference pconv.py:30, losses.py:31-32).
//
// The generic kernel (conv_igemm.cu) re-fetches the 128-pixel A tile once per filter tap and the weight slab once
// per tile: 8 KB of operands per MMA at N = 128, ~125 B/clk/SM, which is L2-bound at ~45 % of the tensor peak
// (profiles/r01_*). Here
//   * the (8+2) x (16+2) input halo of a tile is loaded once per 64-channel block and serves all 9 taps
//     (descriptor offsets, as in conv_halo.cu), and
//   * the weight slab of one (tap, channel block) streams through a small ring and is used by TWO adjacent pixel
//     tiles (two accumulators) before it is released,
// i.e. (2 x 23 KB + 9 x 16 KB) per 72 MMAs = 2.6 KB per MMA at N = 128.
// warp 0: TMA producer, warp 1: MMA issue, warp 2: TMEM allocator, warps 4-11: epilogue (conv_epilogue.cuh).
#include "conv_epilogue.cuh"
#include "conv_igemm.cuh"
#include "tg_common.cuh"
#include "../../include/terragan_b200.h"

namespace tg {

constexpr int kHsW = 8, kHsH = 16;                        // output tile (pixels)
constexpr int kHsRows = (kHsW + 2) * (kHsH + 2);          // 180 halo pixels
constexpr int kHsHaloBytes = 23 * 1024;                   // 180 * 128 B rounded up to the swizzle period
constexpr int kHsHaloSlots = 4;                           // two (unit, channel block) steps in flight
constexpr int kHsVec = 128;

template <int BN>
struct HsCfg {
  static constexpr int kSlabBytes = BN * 128;
  static constexpr int kSlabSlots = BN == 128 ? 5 : 6;
  static constexpr int kSC = 32;
  static constexpr int kSmem = kHsHaloSlots * kHsHaloBytes + kSlabSlots * kSlabBytes + (4 * 2 * kHsVec + 3 * kHsVec) * 4 +
                               8 * 32 * kSC * 2 + 512 + 1024;
};

template <int BN>
__global__ void __launch_bounds__(384, 1)
conv_halo_stream_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const __grid_constant__ ConvKParams p) {
  using Cfg = HsCfg<BN>;
  constexpr int kSlots = Cfg::kSlabSlots, kSlab = Cfg::kSlabBytes, kSC = Cfg::kSC;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* s_a = smem;                                      // halo slots
  uint8_t* s_w = s_a + kHsHaloSlots * kHsHaloBytes;         // weight slab ring
  float* s_stats = reinterpret_cast<float*>(s_w + kSlots * kSlab);
  float* s_vec = s_stats + 4 * 2 * kHsVec;
  uint8_t* s_out = reinterpret_cast<uint8_t*>(s_vec + 3 * kHsVec);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_out + 8 * 32 * kSC * 2);
  uint64_t* hfull = bars;                                   // [4]
  uint64_t* hempty = bars + kHsHaloSlots;                   // [4]
  uint64_t* bfull = bars + 2 * kHsHaloSlots;                // [kSlots]
  uint64_t* bempty = bfull + kSlots;                        // [kSlots]
  uint64_t* tfull = bempty + kSlots;                        // [2]
  uint64_t* tempty = tfull + 2;                             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t kTmemCols = 4 * BN;                    // 2 sets x 2 tiles x BN columns

  for (int i = threadIdx.x; i < 4 * 2 * kHsVec; i += blockDim.x) s_stats[i] = 0.f;
  const bool has_vec = p.bias != nullptr || p.scale != nullptr || p.shift != nullptr;
  if (has_vec) {
    for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) {
      s_vec[i] = p.bias ? p.bias[i] : 0.f;
      s_vec[kHsVec + i] = p.scale ? p.scale[i] : 1.f;
      s_vec[2 * kHsVec + i] = p.shift ? p.shift[i] : 0.f;
    }
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kHsHaloSlots; ++i) {
      mbar_init(&hfull[i], 1);
      mbar_init(&hempty[i], 1);
    }
    for (int i = 0; i < kSlots; ++i) {
      mbar_init(&bfull[i], 1);
      mbar_init(&bempty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.tiles_b * p.tiles_h * p.tiles_w;
  const int units = (total_tiles + 1) >> 1;                 // a unit = tiles 2u and 2u+1 (neighbours along w)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t hcount = 0, bcount = 0;                      // halo / slab fills so far (slot = count % slots)
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int nt = (2 * u + 1 < total_tiles) ? 2 : 1;
        for (int cb = 0; cb < p.cin_blocks; ++cb) {
          for (int mt = 0; mt < 2; ++mt, ++hcount) {        // the slot of the missing tile of an odd tail is skipped
            if (mt >= nt) continue;
            const int tile = 2 * u + mt;
            const int tw = tile % p.tiles_w;
            const int th = (tile / p.tiles_w) % p.tiles_h;
            const int tb = tile / (p.tiles_w * p.tiles_h);
            const int slot = hcount % kHsHaloSlots;
            mbar_wait(&hempty[slot], ((hcount / kHsHaloSlots) & 1u) ^ 1u);
            mbar_arrive_expect_tx(&hfull[slot], kHsRows * 128);
            tma_load_5d(s_a + slot * kHsHaloBytes, &tmA, &hfull[slot], cb * 64, tw * kHsW - 1, th * kHsH - 1, 0, tb);
          }
          for (int t = 0; t < 9; ++t, ++bcount) {
            const int slot = bcount % kSlots;
            mbar_wait(&bempty[slot], ((bcount / kSlots) & 1u) ^ 1u);
            mbar_arrive_expect_tx(&bfull[slot], kSlab);
            tma_load_2d(s_w + slot * kSlab, &tmB, &bfull[slot], (t * p.cin_blocks + cb) * 64, 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issue: whole warp, warp-uniform addressing, one elected lane issues ==========
    constexpr uint32_t idesc = make_idesc_bf16(128, BN, false, false);
    const uint64_t da0 = make_smem_desc(smem_u32(s_a), 16, (kHsW + 2) * 128);
    const uint64_t db0 = make_smem_desc(smem_u32(s_w), 16, 1024);
    const uint32_t a_hi = static_cast<uint32_t>(da0 >> 32), b_hi = static_cast<uint32_t>(db0 >> 32);
    const uint32_t a_lo0 = static_cast<uint32_t>(da0), b_lo0 = static_cast<uint32_t>(db0);
    uint32_t tap_off[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) tap_off[t] = static_cast<uint32_t>((p.tap_dh[t] + 1) * (kHsW + 2) + (p.tap_dw[t] + 1)) * 8u;
    uint32_t hcount = 0, bcount = 0, ucount = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++ucount) {
      const int nt = (2 * u + 1 < total_tiles) ? 2 : 1;
      const int set = ucount & 1u;
      mbar_wait(&tempty[set], ((ucount >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d0 = tmem_base + set * 2 * BN;
      for (int cb = 0; cb < p.cin_blocks; ++cb, hcount += 2) {
        const int hs0 = hcount % kHsHaloSlots, hs1 = (hcount + 1) % kHsHaloSlots;
        mbar_wait(&hfull[hs0], (hcount / kHsHaloSlots) & 1u);
        if (nt == 2) mbar_wait(&hfull[hs1], ((hcount + 1) / kHsHaloSlots) & 1u);
        const uint32_t a0 = a_lo0 + static_cast<uint32_t>(hs0) * (kHsHaloBytes >> 4);
        const uint32_t a1 = a_lo0 + static_cast<uint32_t>(hs1) * (kHsHaloBytes >> 4);
#pragma unroll
        for (int t = 0; t < 9; ++t, ++bcount) {
          const int bs = bcount % kSlots;
          mbar_wait(&bfull[bs], (bcount / kSlots) & 1u);
          tc_fence_after();
          const uint32_t b_lo = b_lo0 + static_cast<uint32_t>(bs) * (kSlab >> 4);
          const uint32_t accum = (t == 0) ? static_cast<uint32_t>(cb != 0) : 1u;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_lh(d0, a0 + tap_off[t] + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc, k ? 1u : accum);
            if (nt == 2) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_lh(d0 + BN, a1 + tap_off[t] + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc, k ? 1u : accum);
            }
            umma_commit(&bempty[bs]);
          }
        }
        if (elect_one()) {
          umma_commit(&hempty[hs0]);
          if (nt == 2) umma_commit(&hempty[hs1]);
        }
      }
      if (elect_one()) umma_commit(&tfull[set]);
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = (warp - 4) & 3;
    const int hsel = (warp - 4) >> 2;
    float* my_stats = s_stats + q * (2 * kHsVec);
    const int row = q * 32 + lane;
    const int e_wt = row % kHsW, e_ht = row / kHsW, e_bt = 0;
    const int epi_mode = conv_epilogue_mode(p.code, p.stats, p.gate, p.scale, p.shift, p.bias);
    float racc[BN == 64 ? 64 : 1];
#pragma unroll
    for (int j = 0; j < (BN == 64 ? 64 : 1); ++j) racc[j] = 0.f;
    conv_epilogue_dispatch(epi_mode, [&](auto mode_tag) {
      constexpr int kMode = decltype(mode_tag)::value;
      uint32_t ucount = 0;
      TileWalk tk0(2 * blockIdx.x, 2 * gridDim.x, p.tiles_w, p.tiles_h);
      TileWalk tk1(2 * blockIdx.x + 1, 2 * gridDim.x, p.tiles_w, p.tiles_h);
      for (int u = blockIdx.x; u < units; u += gridDim.x, ++ucount, tk0.next(), tk1.next()) {
        const int nt = (2 * u + 1 < total_tiles) ? 2 : 1;
        const int set = ucount & 1u;
        const int tw0 = tk0.tw, th0 = tk0.th, tb0 = tk0.tb;
        const int tw1 = nt == 2 ? tk1.tw : tw0, th1 = nt == 2 ? tk1.th : th0, tb1 = nt == 2 ? tk1.tb : tb0;
        EpiPrefetch pre0, pre1;
        conv_epilogue_prefetch<BN, kMode, true>(p, q, lane, 0, 0, tw0, th0, tb0, hsel, e_wt, e_ht, e_bt, pre0);
        conv_epilogue_prefetch<BN, kMode, true>(p, q, lane, 0, 0, tw1, th1, tb1, hsel, e_wt, e_ht, e_bt, pre1);
        mbar_wait(&tfull[set], (ucount >> 1) & 1u);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + set * 2 * BN;
        conv_epilogue_tile<BN, kHsVec, kSC, kMode, true>(p, q, lane, 0, 0, tw0, th0, tb0, t_addr, s_vec, my_stats, has_vec,
                                                   s_out + (warp - 4) * (32 * kSC * 2), hsel, e_wt, e_ht, e_bt,
                                                   BN == 64 ? racc : nullptr, &pre0);
        if (nt == 2)
          conv_epilogue_tile<BN, kHsVec, kSC, kMode, true>(p, q, lane, 0, 0, tw1, th1, tb1, t_addr + BN, s_vec, my_stats, has_vec,
                                                     s_out + (warp - 4) * (32 * kSC * 2), hsel, e_wt, e_ht, e_bt,
                                                     BN == 64 ? racc : nullptr, &pre1);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[set]);
      }
    });
    if (BN == 64 && p.stats != nullptr) {
      // flush the per-thread running sums: this warp owns columns hsel*32 .. +31 of its lane quarter's rows
      float sm[32], sq[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        sm[j] = racc[j];
        sq[j] = racc[(BN == 64 ? 32 : 0) + j];
      }
      const float csum = warp_transpose_sum32(sm);
      const float csq = warp_transpose_sum32(sq);
      my_stats[hsel * 32 + lane] += csum;
      my_stats[kHsVec + hsel * 32 + lane] += csq;
    }
    if (p.stats != nullptr) {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int et = threadIdx.x - 128;
      float* dst = p.stats + static_cast<long>(blockIdx.x) * 2 * p.Cout;
      for (int c = et; c < p.Cout; c += 256) {
        float a = 0.f, s2 = 0.f;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          a += s_stats[qq * 2 * kHsVec + c];
          s2 += s_stats[qq * 2 * kHsVec + kHsVec + c];
        }
        dst[c] = a;
        dst[p.Cout + c] = s2;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc<kTmemCols>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// Stride-2 data gradient with N = 64 (the gradients of enc2 and of Discriminator model[2] w.r.t. their
// parity-split inputs): four sub-problems, one per input parity, each a handful of taps with offsets in {-1,0,1}
// on the SAME channels-last gradient tensor. One halo per 64-channel block serves all 16 / 25 taps; the weight
// slabs stream through the ring; the four sub-problems accumulate in four TMEM accumulators and the epilogue
// writes the four output planes. (The generic kernel runs them as four separate tile passes, re-fetching the
// A tile per tap: ~110 B/clk/SM of operands, 400-530 TFLOP/s.)
// ------------------------------------------------------------------------------------------------
constexpr int kS2Slots = 4;                                // ring slots of two 8 KB weight slabs (two taps)
constexpr int kS2HaloSlots = 3;
constexpr int kS2Smem = kS2HaloSlots * kHsHaloBytes + kS2Slots * 2 * 8192 + (4 * 2 * kHsVec + 3 * kHsVec) * 4 + 8 * 32 * 32 * 2 + 1024 + 1024;

__global__ void __launch_bounds__(384, 1)
conv_halo_s2dgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ ConvKParams p) {
  constexpr int BN = 64, kSC = 32, kSlab = 8192;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* s_a = smem;
  uint8_t* s_w = s_a + kS2HaloSlots * kHsHaloBytes;
  float* s_stats = reinterpret_cast<float*>(s_w + kS2Slots * 2 * kSlab);
  float* s_vec = s_stats + 4 * 2 * kHsVec;
  uint8_t* s_out = reinterpret_cast<uint8_t*>(s_vec + 3 * kHsVec);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_out + 8 * 32 * kSC * 2);
  uint64_t* hfull = bars;
  uint64_t* hempty = bars + kS2HaloSlots;
  uint64_t* bfull = bars + 2 * kS2HaloSlots;
  uint64_t* bempty = bfull + kS2Slots;
  uint64_t* tfull = bempty + kS2Slots;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint32_t* s_tapoff = tmem_slot + 2;                       // [kMaxTaps + 4] halo row offset of each tap, >> 4
  int* s_sub = reinterpret_cast<int*>(s_tapoff + kMaxTaps + 4);   // [4][2] tap_begin, tap_count
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < 4 * 2 * kHsVec; i += blockDim.x) s_stats[i] = 0.f;
  for (int i = threadIdx.x; i < kMaxTaps + 4; i += blockDim.x)
    s_tapoff[i] = i < kMaxTaps ? static_cast<uint32_t>((p.tap_dh[i] + 1) * (kHsW + 2) + (p.tap_dw[i] + 1)) * 8u : 0u;
  if (threadIdx.x < 4) {
    s_sub[2 * threadIdx.x] = p.sub[threadIdx.x].tap_begin;
    s_sub[2 * threadIdx.x + 1] = p.sub[threadIdx.x].tap_count;
  }
  const bool has_vec = p.bias != nullptr || p.scale != nullptr || p.shift != nullptr;
  if (has_vec) {
    for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) {
      s_vec[i] = p.bias ? p.bias[i] : 0.f;
      s_vec[kHsVec + i] = p.scale ? p.scale[i] : 1.f;
      s_vec[2 * kHsVec + i] = p.shift ? p.shift[i] : 0.f;
    }
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kS2HaloSlots; ++i) {
      mbar_init(&hfull[i], 1);
      mbar_init(&hempty[i], 1);
    }
    for (int i = 0; i < kS2Slots; ++i) {
      mbar_init(&bfull[i], 1);
      mbar_init(&bempty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int total_tiles = p.tiles_b * p.tiles_h * p.tiles_w;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t hcount = 0, bcount = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int tw = tile % p.tiles_w;
        const int th = (tile / p.tiles_w) % p.tiles_h;
        const int tb = tile / (p.tiles_w * p.tiles_h);
        for (int cb = 0; cb < p.cin_blocks; ++cb, ++hcount) {
          const int hs = hcount % kS2HaloSlots;
          mbar_wait(&hempty[hs], ((hcount / kS2HaloSlots) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&hfull[hs], kHsRows * 128);
          tma_load_5d(s_a + hs * kHsHaloBytes, &tmA, &hfull[hs], cb * 64, tw * kHsW - 1, th * kHsH - 1, 0, tb);
          for (int sb = 0; sb < 4; ++sb) {
            const ConvSubK sub = p.sub[sb];
            for (int t = 0; t < sub.tap_count; t += 2, ++bcount) {        // a ring slot holds the slabs of two taps
              const int n = min(2, sub.tap_count - t);
              const int slot = bcount % kS2Slots;
              mbar_wait(&bempty[slot], ((bcount / kS2Slots) & 1u) ^ 1u);
              mbar_arrive_expect_tx(&bfull[slot], n * kSlab);
              for (int j = 0; j < n; ++j)
                tma_load_2d(s_w + (slot * 2 + j) * kSlab, &tmB, &bfull[slot], sub.k_off + ((t + j) * p.cin_blocks + cb) * 64, 0);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, BN, false, false);
    const uint64_t da0 = make_smem_desc(smem_u32(s_a), 16, (kHsW + 2) * 128);
    const uint64_t db0 = make_smem_desc(smem_u32(s_w), 16, 1024);
    const uint32_t a_hi = static_cast<uint32_t>(da0 >> 32), b_hi = static_cast<uint32_t>(db0 >> 32);
    const uint32_t a_lo0 = static_cast<uint32_t>(da0), b_lo0 = static_cast<uint32_t>(db0);
    uint32_t hcount = 0, bcount = 0, ucount = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ucount) {
      const int set = ucount & 1u;
      mbar_wait(&tempty[set], ((ucount >> 1) & 1u) ^ 1u);
      tc_fence_after();
      for (int cb = 0; cb < p.cin_blocks; ++cb, ++hcount) {
        const int hs = hcount % kS2HaloSlots;
        mbar_wait(&hfull[hs], (hcount / kS2HaloSlots) & 1u);
        const uint32_t a0 = a_lo0 + static_cast<uint32_t>(hs) * (kHsHaloBytes >> 4);
        for (int sb = 0; sb < 4; ++sb) {
          const int tb0 = s_sub[2 * sb], tcnt = s_sub[2 * sb + 1];
          const uint32_t d = tmem_base + set * 4 * BN + sb * BN;
          uint32_t off0 = s_tapoff[tb0], off1 = s_tapoff[tb0 + 1];   // tap offsets come one iteration ahead of their use
          for (int t = 0; t < tcnt; t += 2, ++bcount) {
            const int bs = bcount % kS2Slots;
            const bool two = t + 1 < tcnt;
            const uint32_t al0 = a0 + off0, al1 = a0 + off1;
            off0 = s_tapoff[tb0 + t + 2];
            off1 = s_tapoff[tb0 + t + 3];
            mbar_wait(&bfull[bs], (bcount / kS2Slots) & 1u);
            tc_fence_after();
            const uint32_t b_lo = b_lo0 + static_cast<uint32_t>(bs) * (2 * kSlab >> 4);
            const uint32_t accum = static_cast<uint32_t>((cb | t) != 0);
            if (elect_one()) {
              if (!TG_DBG(p, 4)) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16_lh(d, al0 + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc, k ? 1u : accum);
                if (two) {
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_bf16_lh(d, al1 + 2 * k, a_hi, b_lo + (kSlab >> 4) + 2 * k, b_hi, idesc, 1u);
                }
              }
              umma_commit(&bempty[bs]);
            }
          }
        }
        if (elect_one()) umma_commit(&hempty[hs]);
      }
      if (elect_one()) umma_commit(&tfull[set]);
    }
  } else if (warp >= 4) {
    const int q = (warp - 4) & 3;
    const int hsel = (warp - 4) >> 2;
    float* my_stats = s_stats + q * (2 * kHsVec);
    const int row = q * 32 + lane;
    const int e_wt = row % kHsW, e_ht = row / kHsW, e_bt = 0;
    const int epi_mode = conv_epilogue_mode(p.code, nullptr, p.gate, p.scale, p.shift, p.bias);   // no statistics here
    conv_epilogue_dispatch(epi_mode, [&](auto mode_tag) {
      constexpr int kMode = decltype(mode_tag)::value;
      uint32_t ucount = 0;
      TileWalk tk(blockIdx.x, gridDim.x, p.tiles_w, p.tiles_h);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ucount, tk.next()) {
        const int tw = tk.tw, th = tk.th, tb = tk.tb;
        const int set = ucount & 1u;
        EpiPrefetch pre, nxt;
        conv_epilogue_prefetch<BN, kMode, true>(p, q, lane, 0, 0, tw, th, tb, hsel, e_wt, e_ht, e_bt, pre);
        mbar_wait(&tfull[set], (ucount >> 1) & 1u);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + set * 4 * BN;
        uint8_t* stg = s_out + (warp - 4) * (32 * kSC * 2);
#pragma unroll 1
        for (int sb = 0; sb < 4; ++sb) {
          // the next plane's ratio code / gate rows are requested before this plane's tile is drained
          if (sb < 3) conv_epilogue_prefetch<BN, kMode, true>(p, q, lane, 0, sb + 1, tw, th, tb, hsel, e_wt, e_ht, e_bt, nxt);
          if (!TG_DBG(p, 2))
            conv_epilogue_tile<BN, kHsVec, kSC, kMode, true>(p, q, lane, 0, sb, tw, th, tb, t_addr + sb * BN, s_vec, my_stats,
                                                             has_vec, stg, hsel, e_wt, e_ht, e_bt, nullptr, &pre);
          pre = nxt;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[set]);
      }
    });
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

// stride-2 data gradient, N = 64: P = 1 input, four output parity planes of the input's size, taps within +-1
bool conv_halo_s2dgrad_eligible(const tg_conv_args* a) {
  if (a->P != 1 || a->Po != 4 || a->num_sub != 4 || a->N != 64 || a->stats != nullptr) return false;
  if (a->Ho != a->H || a->Wo != a->W || a->Ho % kHsH || a->Wo % kHsW) return false;
  if (a->num_taps > TG_MAX_TAPS) return false;
  for (int t = 0; t < a->num_taps; ++t)
    if (a->tap_plane[t] != 0 || a->tap_dh[t] < -1 || a->tap_dh[t] > 1 || a->tap_dw[t] < -1 || a->tap_dw[t] > 1)
      return false;
  for (int s = 0; s < 4; ++s)
    if (a->sub[s].tap_count < 1) return false;
  return true;
}

int conv_halo_s2dgrad_launch(tg_conv_args* a, ConvKParams kp, cudaStream_t st) {
  kp.Bt = 1;
  kp.Ht = kHsH;
  kp.Wt = kHsW;
  kp.tiles_w = a->Wo / kHsW;
  kp.tiles_h = a->Ho / kHsH;
  kp.tiles_b = a->B;
  kp.n_tiles = 1;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[5] = {(uint64_t)a->C, (uint64_t)a->W, (uint64_t)a->H, 1, (uint64_t)a->B};
    uint64_t str[4] = {(uint64_t)a->C * 2, (uint64_t)a->C * 2 * a->W, (uint64_t)a->C * 2 * a->W * a->H,
                       (uint64_t)a->C * 2 * a->W * a->H};
    uint32_t box[5] = {64, kHsW + 2, kHsH + 2, 1, 1};
    if (make_tmap_bf16(&tmA, a->x, 5, dims, str, box) != 0) return -3;
  }
  {
    uint64_t dims[2] = {(uint64_t)a->Ktot, (uint64_t)a->N};
    uint64_t str[1] = {(uint64_t)a->Ktot * 2};
    uint32_t box[2] = {64, 64};
    if (make_tmap_bf16(&tmB, a->w, 2, dims, str, box) != 0) return -3;
  }
  const long total_tiles = (long)kp.tiles_b * kp.tiles_h * kp.tiles_w;
  const int sms = num_sms();
  const int grid = (int)(total_tiles < sms ? total_tiles : sms);
  a->stats_rows_used = 0;
  TG_SET_SMEM_ONCE((conv_halo_s2dgrad_kernel), kS2Smem);
  conv_halo_s2dgrad_kernel<<<grid, 384, kS2Smem, st>>>(tmA, tmB, kp);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int BN>
static int launch_halo_stream(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvKParams& kp, int grid, cudaStream_t st) {
  TG_SET_SMEM_ONCE((conv_halo_stream_kernel<BN>), HsCfg<BN>::kSmem);
  conv_halo_stream_kernel<BN><<<grid, 384, HsCfg<BN>::kSmem, st>>>(tmA, tmB, kp);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// 3x3 / stride-1 / pad-1 stencils with N = 64 or 128 whatever the size of the weights
bool conv_halo_stream_eligible(const tg_conv_args* a) {
  if (a->P != 1 || a->Po != 1 || a->num_sub != 1 || a->num_taps != 9) return false;
  if (a->N != 64 && a->N != 128) return false;
  if (a->Ho != a->H || a->Wo != a->W || a->Ho % kHsH || a->Wo % kHsW) return false;
  for (int t = 0; t < 9; ++t)
    if (a->tap_plane[t] != 0 || a->tap_dh[t] < -1 || a->tap_dh[t] > 1 || a->tap_dw[t] < -1 || a->tap_dw[t] > 1)
      return false;
  return true;
}

int conv_halo_stream_launch(tg_conv_args* a, ConvKParams kp, cudaStream_t st) {
  kp.Bt = 1;
  kp.Ht = kHsH;
  kp.Wt = kHsW;
  kp.tiles_w = a->Wo / kHsW;
  kp.tiles_h = a->Ho / kHsH;
  kp.tiles_b = a->B;
  kp.n_tiles = 1;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[5] = {(uint64_t)a->C, (uint64_t)a->W, (uint64_t)a->H, 1, (uint64_t)a->B};
    uint64_t str[4] = {(uint64_t)a->C * 2, (uint64_t)a->C * 2 * a->W, (uint64_t)a->C * 2 * a->W * a->H,
                       (uint64_t)a->C * 2 * a->W * a->H};
    uint32_t box[5] = {64, kHsW + 2, kHsH + 2, 1, 1};
    if (make_tmap_bf16(&tmA, a->x, 5, dims, str, box) != 0) return -3;
  }
  {
    uint64_t dims[2] = {(uint64_t)a->Ktot, (uint64_t)a->N};
    uint64_t str[1] = {(uint64_t)a->Ktot * 2};
    uint32_t box[2] = {64, (uint32_t)a->N};
    if (make_tmap_bf16(&tmB, a->w, 2, dims, str, box) != 0) return -3;
  }
  const long total_tiles = (long)kp.tiles_b * kp.tiles_h * kp.tiles_w;
  const long units = (total_tiles + 1) / 2;
  const int sms = num_sms();
  const int grid = (int)(units < sms ? units : sms);
  if (a->stats != nullptr) {
    TG_REQUIRE(a->stats_rows_cap >= grid, "tg_conv_igemm: stats_rows_cap %d < grid %d", a->stats_rows_cap, grid);
  }
  a->stats_rows_used = grid;
  if (a->N == 128) return launch_halo_stream<128>(tmA, tmB, kp, grid, st);
  return launch_halo_stream<64>(tmA, tmB, kp, grid, st);
}

}  // namespace tg
