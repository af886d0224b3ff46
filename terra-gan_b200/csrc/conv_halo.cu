// conv_halo.cu — 3x3 / stride-1 implicit-GEMM convolution with HALO-TILE REUSE and RESIDENT WEIGHTS.
//
// Same contract as conv_igemm.cu (fprop and dgrad of nn.Conv2d 3x3 s1 p1 on the path: dec1 / VGG conv2 /
// VGG conv5 and their data gradients — reference pconv.py:30, losses.py:31-32), specialised for the
// layers whose output width N is small (64 / 128): there the generic kernel re-fetches the A tile from
// L2 once per filter tap (9x) and the B slab once per K block, ~190 B/clk/SM of operand traffic against
// ~128 MMA clocks per block, and is L2-bound at ~30 % of the tensor peak (profiles/r01_*).
//
// Here one CTA tile is 8 (w) x 16 (h) output pixels. For each 64-channel block the (8+2) x (16+2) halo
// of the channels-last input is loaded ONCE by a single TMA box (zero-filled borders = conv padding)
// and all 9 taps are issued from it: for tap (dh, dw) the A descriptor simply starts at halo row
// (dh+1)*10 + (dw+1) with a stride of 10 rows (1280 B) between 8-pixel swizzle atoms. tcgen05's
// SWIZZLE_128B addressing is a function of the absolute shared-memory address bits (measured:
// tools/probe/umma_probe.cu), so TMA-written data can be consumed at any 128-byte row offset.
// The whole packed weight matrix (<= 144 KB) is loaded into shared memory once per CTA and stays
// resident, so the steady-state operand traffic is ~23 KB per 9*4 MMAs (~20 B/clk/SM).
#include "conv_epilogue.cuh"
#include "conv_igemm.cuh"
#include "tg_common.cuh"
#include "../../include/terragan_b200.h"

namespace tg {

constexpr int kHaloW = 8, kHaloH = 16;                       // output tile
constexpr int kHaloRows = (kHaloW + 2) * (kHaloH + 2);       // 180 halo pixels
constexpr int kHaloStageBytes = 23 * 1024;                   // 180 * 128 B rounded up to the swizzle period
constexpr int kHaloStages = 3;                               // upper bound; 2 when the weights are large
constexpr int kHaloVec = 128;                                // max output channels of this kernel

template <int BN>
struct HaloCfg {
  static constexpr int kStoreCols = 32;      // 32-column staging also for N = 128: the 64 -> 128 layer then fits with resident weights
};

struct HaloSmem {
  static int total(int w_bytes, int kStoreCols = 64, int stages = kHaloStages) {
    return w_bytes + stages * kHaloStageBytes + (4 * 2 * kHaloVec + 3 * kHaloVec) * 4 + 8 * 32 * kStoreCols * 2 + 256 + 1024;
  }
};

template <int BN>
__global__ void __launch_bounds__(384, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ ConvKParams p, int w_bytes, int n_stages) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* s_w = smem;                                       // [taps*cin_blocks] slabs of BN rows x 128 B
  uint8_t* s_a = smem + w_bytes;                             // kHaloStages halo tiles
  float* s_stats = reinterpret_cast<float*>(s_a + n_stages * kHaloStageBytes);
  float* s_vec = s_stats + 4 * 2 * kHaloVec;
  constexpr int kSC = HaloCfg<BN>::kStoreCols;
  uint8_t* s_out = reinterpret_cast<uint8_t*>(s_vec + 3 * kHaloVec);   // 8 warps x 32 rows x kSC*2 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_out + 8 * 32 * kSC * 2);
  uint64_t* full_bar = bars;                                 // [kHaloStages]
  uint64_t* empty_bar = bars + kHaloStages;                  // [kHaloStages]
  uint64_t* tfull_bar = bars + 2 * kHaloStages;              // [2]
  uint64_t* tempty_bar = bars + 2 * kHaloStages + 2;         // [2]
  uint64_t* w_bar = bars + 2 * kHaloStages + 4;              // weights resident
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kHaloStages + 5);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t kTmemCols = (2 * BN <= 128) ? 128 : 256;
  constexpr int kSlabBytes = BN * 128;

  for (int i = threadIdx.x; i < 4 * 2 * kHaloVec; i += blockDim.x) s_stats[i] = 0.f;
  const bool has_vec = p.bias != nullptr || p.scale != nullptr || p.shift != nullptr;
  if (has_vec) {
    for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) {
      s_vec[i] = p.bias ? p.bias[i] : 0.f;
      s_vec[kHaloVec + i] = p.scale ? p.scale[i] : 1.f;
      s_vec[2 * kHaloVec + i] = p.shift ? p.shift[i] : 0.f;
    }
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kHaloStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 8);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int taps = p.sub[0].tap_count;
  const int m_tiles = p.tiles_b * p.tiles_h * p.tiles_w;
  const int total_tiles = m_tiles * p.n_tiles;   // n_tiles == 1 for this kernel

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      // resident weights: one slab per (tap, channel block)
      mbar_arrive_expect_tx(w_bar, static_cast<uint32_t>(w_bytes));
      for (int t = 0; t < taps; ++t)
        for (int cb = 0; cb < p.cin_blocks; ++cb)
          tma_load_2d(s_w + (t * p.cin_blocks + cb) * kSlabBytes, &tmB, w_bar, (t * p.cin_blocks + cb) * 64, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int tw = tile % p.tiles_w;
        const int th = (tile / p.tiles_w) % p.tiles_h;
        const int tb = tile / (p.tiles_w * p.tiles_h);
        for (int cb = 0; cb < p.cin_blocks; ++cb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], kHaloRows * 128);
          tma_load_5d(s_a + stage * kHaloStageBytes, &tmA, &full_bar[stage], cb * 64, tw * kHaloW - 1,
                      th * kHaloH - 1, 0, tb);
          if (++stage == n_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the loop so that addresses and descriptors stay warp-uniform (uniform registers);
    // only the tcgen05 instructions themselves are issued by one elected lane. Under `if (lane == 0)` the
    // compiler wraps every MMA in an ELECT / R2UR.BROADCAST loop (~10 extra instructions per MMA).
    {
      constexpr uint32_t idesc = make_idesc_bf16(128, BN, false, false);
      mbar_wait(w_bar, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      // Descriptor words are fixed per operand; the loop below only adds byte offsets (>> 4) to the lo words:
      // tap (dh, dw) starts at halo row (dh+1)*10 + (dw+1), 8-pixel atoms are 10 halo rows (1280 B) apart.
      const uint64_t da0 = make_smem_desc(smem_u32(s_a), 16, (kHaloW + 2) * 128);
      const uint64_t db0 = make_smem_desc(smem_u32(s_w), 16, 1024);
      const uint32_t a_hi = static_cast<uint32_t>(da0 >> 32), b_hi = static_cast<uint32_t>(db0 >> 32);
      const uint32_t a_lo0 = static_cast<uint32_t>(da0), b_lo0 = static_cast<uint32_t>(db0);
      uint32_t tap_off[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) tap_off[t] = static_cast<uint32_t>((p.tap_dh[t] + 1) * (kHaloW + 2) + (p.tap_dw[t] + 1)) * 8u;
      const uint32_t slab16 = static_cast<uint32_t>(kSlabBytes >> 4);
      const uint32_t tap_step = static_cast<uint32_t>(p.cin_blocks) * slab16;
      const bool skip_mma = TG_DBG(p, 4);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int cb = 0; cb < p.cin_blocks; ++cb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + static_cast<uint32_t>(stage) * (kHaloStageBytes >> 4);
          uint32_t b_lo = b_lo0 + static_cast<uint32_t>(cb) * slab16;
          if (!skip_mma && elect_one()) {
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              const uint32_t al = a_lo + tap_off[t];
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_lh(d_tmem, al + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc, (t | k) ? 1u : static_cast<uint32_t>(cb != 0));
              b_lo += tap_step;
            }
          }
          if (elect_one()) umma_commit(&empty_bar[stage]);
          if (++stage == n_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma_commit(&tfull_bar[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = (warp - 4) & 3;
    const int hsel = (warp - 4) >> 2;
    float* my_stats = s_stats + q * (2 * kHaloVec);
    const int row = q * 32 + lane;                     // row of the tile = pixel in box order (8 wide, 16 high)
    const int e_wt = row % kHaloW, e_ht = row / kHaloW, e_bt = 0;
    const int epi_mode = conv_epilogue_mode(p.code, p.stats, p.gate, p.scale, p.shift, p.bias);
    float racc[BN == 64 ? 64 : 1];
#pragma unroll
    for (int j = 0; j < (BN == 64 ? 64 : 1); ++j) racc[j] = 0.f;
    conv_epilogue_dispatch(epi_mode, [&](auto mode_tag) {
      constexpr int kMode = decltype(mode_tag)::value;
      int acc = 0;
      uint32_t acc_phase = 0;
      TileWalk tk(blockIdx.x, gridDim.x, p.tiles_w, p.tiles_h);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, tk.next()) {
        const int tw = tk.tw, th = tk.th, tb = tk.tb;
        EpiPrefetch pre;
        conv_epilogue_prefetch<BN, kMode, true>(p, q, lane, 0, 0, tw, th, tb, hsel, e_wt, e_ht, e_bt, pre);
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
        if (!TG_DBG(p, 2))
          conv_epilogue_tile<BN, kHaloVec, kSC, kMode, true>(p, q, lane, 0, 0, tw, th, tb, t_addr, s_vec, my_stats, has_vec,
                                                       s_out + (warp - 4) * (32 * kSC * 2), hsel, e_wt, e_ht, e_bt,
                                                       BN == 64 ? racc : nullptr, &pre);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
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
      my_stats[kHaloVec + hsel * 32 + lane] += csq;
    }
    if (p.stats != nullptr) {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int et = threadIdx.x - 128;
      float* dst = p.stats + static_cast<long>(blockIdx.x) * 2 * p.Cout;
      for (int c = et; c < p.Cout; c += 256) {
        float a = 0.f, s2 = 0.f;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          a += s_stats[qq * 2 * kHaloVec + c];
          s2 += s_stats[qq * 2 * kHaloVec + kHaloVec + c];
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

template <int BN>
static int launch_halo(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvKParams& kp, int w_bytes, int grid,
                       cudaStream_t st) {
  const int stages = HaloSmem::total(w_bytes, HaloCfg<BN>::kStoreCols, 3) <= 227 * 1024 ? 3 : 2;
  const int smem = HaloSmem::total(w_bytes, HaloCfg<BN>::kStoreCols, stages);
  static int attr_set = 0;
  if (attr_set < smem) {
    TG_CHECK_CUDA(cudaFuncSetAttribute(conv_halo_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = smem;
  }
  conv_halo_kernel<BN><<<grid, 384, smem, st>>>(tmA, tmB, kp, w_bytes, stages);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Is this problem a 3x3 / stride-1 / pad-1 stencil the halo kernel can take?
bool conv_halo_eligible(const tg_conv_args* a) {
  if (a->P != 1 || a->Po != 1 || a->num_sub != 1 || a->num_taps != 9) return false;
  if (a->N != 64 && a->N != 128) return false;
  if (a->Ho != a->H || a->Wo != a->W || a->Ho % kHaloH || a->Wo % kHaloW) return false;
  for (int t = 0; t < 9; ++t)
    if (a->tap_plane[t] != 0 || a->tap_dh[t] < -1 || a->tap_dh[t] > 1 || a->tap_dw[t] < -1 || a->tap_dw[t] > 1)
      return false;
  const long w_bytes = static_cast<long>(a->N) * a->Ktot * 2;
  return HaloSmem::total(static_cast<int>(w_bytes), 32, 2) <= 227 * 1024;
}

int conv_halo_launch(tg_conv_args* a, ConvKParams kp, cudaStream_t st) {
  kp.Bt = 1;
  kp.Ht = kHaloH;
  kp.Wt = kHaloW;
  kp.tiles_w = a->Wo / kHaloW;
  kp.tiles_h = a->Ho / kHaloH;
  kp.tiles_b = a->B;
  kp.n_tiles = 1;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[5] = {(uint64_t)a->C, (uint64_t)a->W, (uint64_t)a->H, 1, (uint64_t)a->B};
    uint64_t str[4] = {(uint64_t)a->C * 2, (uint64_t)a->C * 2 * a->W, (uint64_t)a->C * 2 * a->W * a->H,
                       (uint64_t)a->C * 2 * a->W * a->H};
    uint32_t box[5] = {64, kHaloW + 2, kHaloH + 2, 1, 1};
    if (make_tmap_bf16(&tmA, a->x, 5, dims, str, box) != 0) return -3;
  }
  {
    uint64_t dims[2] = {(uint64_t)a->Ktot, (uint64_t)a->N};
    uint64_t str[1] = {(uint64_t)a->Ktot * 2};
    uint32_t box[2] = {64, (uint32_t)a->N};
    if (make_tmap_bf16(&tmB, a->w, 2, dims, str, box) != 0) return -3;
  }
  const int w_bytes = a->N * a->Ktot * 2;
  const long total_tiles = (long)kp.tiles_b * kp.tiles_h * kp.tiles_w;
  const int sms = num_sms();
  const int grid = (int)(total_tiles < sms ? total_tiles : sms);
  if (a->stats != nullptr) {
    TG_REQUIRE(a->stats_rows_cap >= grid, "tg_conv_igemm: stats_rows_cap %d < grid %d", a->stats_rows_cap, grid);
  }
  a->stats_rows_used = grid;
  if (a->N == 128) return launch_halo<128>(tmA, tmB, kp, w_bytes, grid, st);
  return launch_halo<64>(tmA, tmB, kp, w_bytes, grid, st);
}

}  // namespace tg
