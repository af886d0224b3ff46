// conv_igemm.cu — tcgen05/TMEM/TMA implicit-GEMM convolution (fprop + dgrad) for sm_100a.
//
// Replaces, on the TERRA-GAN hot path, nn.Conv2d forward and backward-data for
//   PConv2d.input_conv                    (reference mvp_gan/src/models/pconv.py:30)
//   Discriminator convs model[2,5,8]      (mvp_gan/src/models/discriminator.py:11)
//   VGG16 features[:16] convs             (mvp_gan/src/utils/losses.py:31-32,86-89)
// and fuses into the epilogue the PConv renormalisation `(conv + bias) * ratio`
// (pconv.py:38-43, ratio via a <=50-entry LUT indexed by the integer window count), the
// eval-mode BatchNorm affine (pconv.py:47), ReLU / LeakyReLU (pconv.py:48, discriminator.py:14)
// and the per-channel sum / sum-of-squares partials train-mode BatchNorm needs.
//
// GEMM view: M = output pixels (128 per tile = a Bt x Ht x Wt box of the output grid), N = output
// channels, K = taps x input channels. One K block = one tap x 64 channels: the A tile is ONE TMA
// box load of the channels-last input shifted by the tap offset (out-of-range rows/cols are
// zero-filled by TMA = conv zero padding), landing as 128 rows x 128 B in SWIZZLE_128B layout,
// which is exactly the K-major UMMA operand layout. Stride-2 convs read a parity-split input so
// every tap is still a dense shifted box. Accumulators live in TMEM (double buffered) so the
// epilogue of tile i overlaps the MMAs of tile i+1.
//
// Warp roles (256 threads): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator,
// warps 4..7 epilogue (TMEM lane quarter = warp_idx % 4).
#include <stdlib.h>

#include "conv_epilogue.cuh"
#include "conv_igemm.cuh"
#include "tg_common.cuh"
#include "../../include/terragan_b200.h"

namespace tg {

constexpr int kABytes = 128 * 128;  // 128 pixels x 64 bf16

// kMT: M-tiles (128 pixels each) that share every weight tile. The operand traffic per MMA — 16 KB of pixels + BN x 128 B
// of weights for four MMAs — is what bounds the N = 128 layers (128 B/clk/SM from L2, 50 % tensor pipe in ncu);
// two M-tiles per unit, each with its own accumulator pair, cut it to 96 B/clk.
// k2: CTA pair (cluster of 2, tcgen05.mma.cta_group::2): the pair computes a 256-pixel x BN tile; each CTA stages its own
// 128 pixels and HALF of the weight tile (BN/2 rows), the tensor cores of both SMs read the two halves across the pair.
// Per SM and K block that is 16 KB + 16 KB instead of 16 KB + 32 KB for four N = 256 MMAs (512 clocks): 64 instead of
// 96 B/clk of shared-memory operand reads and of L2 -> SM fill, and room for 5 stages instead of 3.
template <int BN, int kMT = 1, bool k2 = false>
struct ConvSmem {
  static constexpr int kStages = k2 ? 5 : (BN >= 256 || kMT > 1) ? 3 : 4;
  static constexpr int kBBytes = k2 ? BN * 64 : BN * 128;
  static constexpr int kStageBytes = kMT * kABytes + kBBytes;
  static constexpr int kTiles = kStages * kStageBytes;
  static constexpr int kStatsFloats = 4 * 2 * 512;  // per epilogue warp (sum, sumsq) x channel
  static constexpr int kVecFloats = 3 * 512;        // bias / scale / shift
  static constexpr int kStoreCols = (BN == 128 || BN == 192) ? 64 : 32;    // epilogue staging width
  static constexpr int kStageOutBytes = 8 * 32 * kStoreCols * 2;            // per-warp transposing tiles
  static constexpr int kTotal = kTiles + (kStatsFloats + kVecFloats) * 4 + kStageOutBytes + 256 + 1024;
};

// kF32: fp32 activations / weights in shared memory (TMA boxes of 32 channels = the same 128-byte rows),
// kind::tf32 MMAs, fp32 output -- the verification path.
template <int BN, bool kF32 = false, int kMT = 1, bool k2 = false>
__global__ void __launch_bounds__(384, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ ConvKParams p) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  using S = ConvSmem<BN, kMT, k2>;
  constexpr int kStages = S::kStages;
  constexpr int kAcc = 2 * kMT;                 // accumulators: two sets of kMT
  static_assert(kAcc * BN <= 512, "accumulators must fit the 512 TMEM columns");
  static_assert(!k2 || (kMT == 1 && !kF32), "the CTA-pair variant is bf16, one M-tile per CTA");
  // CTA pair: rank 0 (the leader) issues the MMAs for both SMs; its full / tempty barriers collect the arrivals of both CTAs
  const uint32_t cta_rank = k2 ? cluster_ctarank() : 0u;
  float* s_stats = reinterpret_cast<float*>(smem + S::kTiles);
  float* s_vec = s_stats + S::kStatsFloats;
  uint8_t* s_out = reinterpret_cast<uint8_t*>(s_vec + S::kVecFloats);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_out + S::kStageOutBytes);
  uint64_t* full_bar = bars;                    // [kStages]
  uint64_t* empty_bar = bars + kStages;         // [kStages]
  uint64_t* tfull_bar = bars + 2 * kStages;     // [kAcc]
  uint64_t* tempty_bar = bars + 2 * kStages + kAcc;  // [kAcc]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 2 * kAcc);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int kKB = kF32 ? 32 : 64;      // channels per K block (128 bytes)
  constexpr uint32_t kTmemCols = (kAcc * BN <= 32) ? 32 : (kAcc * BN <= 64) ? 64 : (kAcc * BN <= 128) ? 128
                                 : (kAcc * BN <= 256) ? 256 : 512;

  // ---- one-time setup ----
  for (int i = threadIdx.x; i < S::kStatsFloats; i += blockDim.x) s_stats[i] = 0.f;
  // per-channel epilogue vectors are staged once (channels <= 512 whenever any of them is used)
  const bool has_vec = p.bias != nullptr || p.scale != nullptr || p.shift != nullptr;
  if (has_vec) {
    for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) {
      s_vec[i] = p.bias ? p.bias[i] : 0.f;
      s_vec[512 + i] = p.scale ? p.scale[i] : 1.f;
      s_vec[1024 + i] = p.shift ? p.shift[i] : 0.f;
    }
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < kAcc; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], k2 ? 16 : 8);  // one arrive per epilogue warp (of both CTAs of a pair)
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (k2) tmem_alloc_2cta<kTmemCols>(tmem_slot);
    else tmem_alloc<kTmemCols>(tmem_slot);
  }
  tc_fence_before();
  if constexpr (k2) cluster_sync_all();      // the peer's barriers and TMEM exist before anything remote touches them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // a unit = kMT consecutive M-tiles (the host guarantees m_tiles % kMT == 0) x one N-tile of one sub-convolution
  const int m_tiles = p.tiles_b * p.tiles_h * p.tiles_w;
  constexpr int kUnitM = k2 ? 2 : kMT;        // M-tiles per unit (CTA pair: one per CTA)
  const int m_units = m_tiles / kUnitM;
  const int total_tiles = p.num_sub * m_units * p.n_tiles;
  const int first_tile = k2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int tile_step = k2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
        const int nt = tile % p.n_tiles;
        const int rest = tile / p.n_tiles;
        const int mu = rest % m_units;
        const int sb = rest / m_units;
        int w0[kMT], h0[kMT], b0[kMT];
#pragma unroll
        for (int j = 0; j < kMT; ++j) {
          const int mt = k2 ? mu * 2 + static_cast<int>(cta_rank) : mu * kMT + j;
          w0[j] = (mt % p.tiles_w) * p.Wt;
          h0[j] = ((mt / p.tiles_w) % p.tiles_h) * p.Ht;
          b0[j] = (mt / (p.tiles_w * p.tiles_h)) * p.Bt;
        }
        const ConvSubK sub = p.sub[sb];
        for (int t = 0; t < sub.tap_count; ++t) {
          const int tt = sub.tap_begin + t;
          const int dh = p.tap_dh[tt], dw = p.tap_dw[tt], pl = p.tap_plane[tt];
          for (int cb = 0; cb < p.cin_blocks; ++cb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * S::kStageBytes;
            uint8_t* sbm = sa + kMT * kABytes;
            if constexpr (k2) {
              // both CTAs' loads complete on the LEADER's barrier, which expects the bytes of the whole pair
              if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * S::kStageBytes);
              const uint32_t lead_bar = mapa_u32(smem_u32(&full_bar[stage]), 0);
              tma_load_5d_2cta(sa, &tmA, lead_bar, cb * kKB, w0[0] + dw, h0[0] + dh, pl, b0[0]);
              tma_load_2d_2cta(sbm, &tmB, lead_bar, sub.k_off + (t * p.cin_blocks + cb) * kKB,
                               nt * BN + static_cast<int>(cta_rank) * (BN / 2));
            } else {
            mbar_arrive_expect_tx(&full_bar[stage], S::kStageBytes);
#pragma unroll
            for (int j = 0; j < kMT; ++j)
              tma_load_5d(sa + j * kABytes, &tmA, &full_bar[stage], cb * kKB, w0[j] + dw, h0[j] + dh, pl, b0[j]);
            tma_load_2d(sbm, &tmB, &full_bar[stage], sub.k_off + (t * p.cin_blocks + cb) * kKB,
                        nt * BN);
            }
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
      if constexpr (k2) {
        // the leader's multicast commits arrive on THIS CTA's empty barriers: drain them before the CTA may exit
        for (int i = 0; i < kStages; ++i) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1 && cta_rank == 0) {
    // ===================== MMA issuer =====================
    // whole warp runs the loop (warp-uniform descriptors in uniform registers); one elected lane issues
    {
      constexpr uint32_t idesc = kF32 ? make_idesc_tf32(128, BN, false, false)
                                      : make_idesc_bf16(k2 ? 256 : 128, BN, false, false);
      const uint64_t d0 = make_smem_desc(smem_u32(smem), 16, 1024);     // A and B: K-major, 8-row atoms 1024 B apart
      const uint32_t d_lo0 = static_cast<uint32_t>(d0), d_hi = static_cast<uint32_t>(d0 >> 32);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;               // first accumulator of the current set (0 or kMT)
      uint32_t acc_phase = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
        const int sb = (tile / p.n_tiles) / m_units;
        const int kblocks = p.sub[sb].tap_count * p.cin_blocks;
#pragma unroll
        for (int j = 0; j < kMT; ++j) mbar_wait(&tempty_bar[acc + j], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          // descriptor lo word = base + stage offset (>> 4); +2 per 16 bf16 (32 B) along K inside the swizzle row
          const uint32_t a_lo = d_lo0 + static_cast<uint32_t>(stage) * (S::kStageBytes >> 4);
          const uint32_t b_lo = a_lo + (kMT * kABytes >> 4);
          if (elect_one()) {
#pragma unroll
            for (int j = 0; j < kMT; ++j) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (kF32) umma_tf32_lh(d_tmem + j * BN, a_lo + j * (kABytes >> 4) + 2 * k, d_hi, b_lo + 2 * k, d_hi, idesc, k ? 1u : static_cast<uint32_t>(kb != 0));
                else if (k2) umma_bf16_lh_2cta(d_tmem, a_lo + 2 * k, d_hi, b_lo + 2 * k, d_hi, idesc, k ? 1u : static_cast<uint32_t>(kb != 0));
                else umma_bf16_lh(d_tmem + j * BN, a_lo + j * (kABytes >> 4) + 2 * k, d_hi, b_lo + 2 * k, d_hi, idesc, k ? 1u : static_cast<uint32_t>(kb != 0));
              }
            }
            if constexpr (k2) umma_commit_2cta(&empty_bar[stage]);   // both CTAs' slots
            else umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < kMT; ++j) {  // accumulators complete -> epilogue (of both CTAs of a pair)
            if constexpr (k2) umma_commit_2cta(&tfull_bar[acc + j]);
            else umma_commit(&tfull_bar[acc + j]);
          }
        }
        acc += kMT;
        if (acc == kAcc) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = (warp - 4) & 3;   // TMEM lane quarter == warp_idx % 4
    const int hsel = (warp - 4) >> 2;  // which half of the column groups this warp drains
    float* my_stats = s_stats + q * (2 * 512);
    const int row = q * 32 + lane;                     // row of the tile = pixel in box order
    const int e_wt = row % p.Wt, e_ht = (row / p.Wt) % p.Ht, e_bt = row / (p.Wt * p.Ht);
    if constexpr (kF32) {
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
        const int nt = tile % p.n_tiles;
        const int rest = tile / p.n_tiles;
        const int sb = rest / m_units;
#pragma unroll 1
        for (int j = 0; j < kMT; ++j) {
          const int mt = (rest % m_units) * kMT + j;
          const int tw = mt % p.tiles_w;
          const int th = (mt / p.tiles_w) % p.tiles_h;
          const int tb = mt / (p.tiles_w * p.tiles_h);
          mbar_wait(&tfull_bar[acc], acc_phase);
          tc_fence_after();
          const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
          conv_epilogue_tile_f32<BN, 512>(p, q, lane, nt, sb, tw, th, tb, t_addr, s_vec, my_stats, has_vec, hsel, e_wt,
                                          e_ht, e_bt);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);
          if (++acc == kAcc) {
            acc = 0;
            acc_phase ^= 1;
          }
        }
      }
    } else {
    const int epi_mode = conv_epilogue_mode(p.code, p.stats, p.gate, p.scale, p.shift, p.bias);
    conv_epilogue_dispatch(epi_mode, [&](auto mode_tag) {
      constexpr int kMode = decltype(mode_tag)::value;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
        const int nt = tile % p.n_tiles;
        const int rest = tile / p.n_tiles;
        const int sb = rest / m_units;
#pragma unroll 1
        for (int j = 0; j < kMT; ++j) {
          const int mt = k2 ? (rest % m_units) * 2 + static_cast<int>(cta_rank) : (rest % m_units) * kMT + j;
          const int tw = mt % p.tiles_w;
          const int th = (mt / p.tiles_w) % p.tiles_h;
          const int tb = mt / (p.tiles_w * p.tiles_h);
          EpiPrefetch pre;
          conv_epilogue_prefetch<BN, kMode>(p, q, lane, nt, sb, tw, th, tb, hsel, e_wt, e_ht, e_bt, pre);
          mbar_wait(&tfull_bar[acc], acc_phase);
          tc_fence_after();
          const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
          conv_epilogue_tile<BN, 512, S::kStoreCols, kMode>(p, q, lane, nt, sb, tw, th, tb, t_addr, s_vec, my_stats, has_vec,
                                                            s_out + (warp - 4) * (32 * S::kStoreCols * 2), hsel, e_wt, e_ht,
                                                            e_bt, nullptr, &pre);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (k2) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));   // the leader's barrier
            else mbar_arrive(&tempty_bar[acc]);
          }
          if (++acc == kAcc) {
            acc = 0;
            acc_phase ^= 1;
          }
        }
      }
    });
    }
    // flush the per-CTA BatchNorm partials: one row per CTA, summed by tg_bn_finalize
    if (p.stats != nullptr) {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int et = threadIdx.x - 128;
      float* dst = p.stats + static_cast<long>(blockIdx.x) * 2 * p.Cout;
      for (int c = et; c < p.Cout; c += 256) {
        float a = 0.f, s2 = 0.f;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          a += s_stats[qq * 1024 + c];
          s2 += s_stats[qq * 1024 + 512 + c];
        }
        dst[c] = a;
        dst[p.Cout + c] = s2;
      }
    }
  }

  // ---- teardown ----
  tc_fence_before();
  if constexpr (k2) cluster_sync_all();      // neither CTA frees TMEM / exits while the pair's MMAs or arrivals may still touch it
  else __syncthreads();
  tc_fence_after();
  if (warp == 2) {
    if constexpr (k2) tmem_dealloc_2cta<kTmemCols>(tmem_base);
    else tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
bool conv_halo_eligible(const tg_conv_args* a);                       // conv_halo.cu
int conv_halo_launch(tg_conv_args* a, ConvKParams kp, cudaStream_t st);
bool conv_halo_stream_eligible(const tg_conv_args* a);                // conv_halo_stream.cu
int conv_halo_stream_launch(tg_conv_args* a, ConvKParams kp, cudaStream_t st);
bool conv_halo_s2dgrad_eligible(const tg_conv_args* a);
int conv_halo_s2dgrad_launch(tg_conv_args* a, ConvKParams kp, cudaStream_t st);

static bool halo_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TG_NO_HALO");
    v = (e != nullptr && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

static bool halo_stream_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TG_NO_HALO_STREAM");
    v = (e != nullptr && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

static void choose_box(int Ho, int Wo, int pixels, int* Bt, int* Ht, int* Wt) {
  int wt = 1;
  while (wt * 2 <= Wo && wt * 2 <= 16 && wt * 2 <= pixels) wt *= 2;
  int ht = 1;
  while (ht * 2 <= Ho && wt * ht * 2 <= pixels) ht *= 2;
  *Wt = wt;
  *Ht = ht;
  *Bt = pixels / (wt * ht);
}

static bool conv_wide_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TG_NO_CONV_WIDE");
    v = (e != nullptr && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

template <int BN, bool kF32 = false, int kMT = 1>
static int launch_conv(const tg_conv_args* a, const CUtensorMap& tmA, const CUtensorMap& tmB,
                       const ConvKParams& kp, int grid, cudaStream_t st) {
  using S = ConvSmem<BN, kMT>;
  TG_SET_SMEM_ONCE((conv_igemm_kernel<BN, kF32, kMT>), S::kTotal);
  conv_igemm_kernel<BN, kF32, kMT><<<grid, 384, S::kTotal, st>>>(tmA, tmB, kp);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static bool conv_pair_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TG_NO_CONV_PAIR");
    v = (e != nullptr && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}
// smallest number of 256-pixel units for which the CTA-pair kernel is chosen: one per pair of SMs, unless
// TG_CONV_PAIR_MIN overrides it (the parity tests force the pair kernel onto their small shapes with 1)
static long conv_pair_min_units(int sms) {
  static long v = -1;
  if (v < 0) {
    const char* e = getenv("TG_CONV_PAIR_MIN");
    v = (e != nullptr && atol(e) > 0) ? atol(e) : 0;
  }
  return v > 0 ? v : sms / 2;
}

// CTA-pair variant (cluster of 2, cta_group::2 MMAs); `grid` = 2 x the number of pairs
template <int BN>
static int launch_conv_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvKParams& kp, int grid, cudaStream_t st) {
  using S = ConvSmem<BN, 1, true>;
  TG_SET_SMEM_ONCE((conv_igemm_kernel<BN, false, 1, true>), S::kTotal);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(384);
  cfg.dynamicSmemBytes = S::kTotal;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  TG_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_igemm_kernel<BN, false, 1, true>, tmA, tmB, kp));
  return 0;
}

}  // namespace tg

extern "C" int tg_conv_pool_fusable(const tg_conv_args* a) {
  using namespace tg;
  if (a == nullptr || a->dtype != TG_DTYPE_BF16 || a->gate != nullptr || !halo_enabled()) return 0;
  return (conv_halo_eligible(a) || (halo_stream_enabled() && conv_halo_stream_eligible(a))) ? 1 : 0;
}

extern "C" int tg_conv_igemm(tg_conv_args* a, void* stream) {
  using namespace tg;
  TG_REQUIRE(a != nullptr, "tg_conv_igemm: null args");
  TG_REQUIRE(a->dtype == TG_DTYPE_BF16 || a->dtype == TG_DTYPE_F32, "tg_conv_igemm: bad dtype %d", a->dtype);
  const bool f32 = a->dtype == TG_DTYPE_F32;
  const int kb_elems = f32 ? 32 : 64;          // channels per 128-byte K block
  const int esz = f32 ? 4 : 2;
  TG_REQUIRE(a->C > 0 && a->C % kb_elems == 0, "tg_conv_igemm: C=%d must be a positive multiple of %d", a->C, kb_elems);
  const bool per_channel = a->bias || a->scale || a->shift || a->stats;
  TG_REQUIRE(a->N > 0 && a->N % 64 == 0 && a->N <= (per_channel ? 512 : 1024),
             "tg_conv_igemm: N=%d must be a multiple of 64 and <= %d", a->N, per_channel ? 512 : 1024);
  TG_REQUIRE(a->num_sub >= 1 && a->num_sub <= TG_MAX_SUB, "tg_conv_igemm: bad num_sub %d", a->num_sub);
  TG_REQUIRE(a->num_taps >= 1 && a->num_taps <= TG_MAX_TAPS, "tg_conv_igemm: bad num_taps %d", a->num_taps);
  TG_REQUIRE(a->code == nullptr || (a->lut != nullptr && a->lut_len >= 1 && a->lut_len <= TG_MAX_TAPS),
             "tg_conv_igemm: code given without a valid lut");
  TG_REQUIRE((reinterpret_cast<uintptr_t>(a->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->w) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(a->out) & 15) == 0,
             "tg_conv_igemm: x/w/out must be 16-byte aligned");
  TG_REQUIRE(a->Ktot % kb_elems == 0, "tg_conv_igemm: Ktot=%d must be a multiple of %d", a->Ktot, kb_elems);
  for (int s = 0; s < a->num_sub; ++s) {
    const tg_conv_sub& sb = a->sub[s];
    TG_REQUIRE(sb.tap_begin >= 0 && sb.tap_count >= 1 && sb.tap_begin + sb.tap_count <= a->num_taps,
               "tg_conv_igemm: sub %d taps out of range", s);
    TG_REQUIRE(sb.k_off >= 0 && sb.k_off + sb.tap_count * a->C <= a->Ktot,
               "tg_conv_igemm: sub %d weight slab out of range", s);
    TG_REQUIRE(sb.out_plane >= 0 && sb.out_plane < a->Po, "tg_conv_igemm: sub %d bad out_plane", s);
  }
  const int BN = (a->N % 256 == 0) ? 256 : (a->N == 192) ? 192 : (a->N % 128 == 0) ? 128 : 64;

  ConvKParams kp;
  memset(&kp, 0, sizeof(kp));
  choose_box(a->Ho, a->Wo, 128, &kp.Bt, &kp.Ht, &kp.Wt);
  kp.tiles_w = (a->Wo + kp.Wt - 1) / kp.Wt;
  kp.tiles_h = (a->Ho + kp.Ht - 1) / kp.Ht;
  kp.tiles_b = (a->B + kp.Bt - 1) / kp.Bt;
  kp.n_tiles = a->N / BN;
  kp.num_sub = a->num_sub;
  for (int s = 0; s < a->num_sub; ++s) {
    kp.sub[s].tap_begin = a->sub[s].tap_begin;
    kp.sub[s].tap_count = a->sub[s].tap_count;
    kp.sub[s].k_off = a->sub[s].k_off;
    kp.sub[s].out_plane = a->sub[s].out_plane;
  }
  for (int t = 0; t < a->num_taps; ++t) {
    TG_REQUIRE(a->tap_plane[t] >= 0 && a->tap_plane[t] < a->P, "tg_conv_igemm: tap %d bad plane", t);
    kp.tap_plane[t] = a->tap_plane[t];
    kp.tap_dh[t] = a->tap_dh[t];
    kp.tap_dw[t] = a->tap_dw[t];
  }
  kp.cin_blocks = a->C / kb_elems;
  kp.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  kp.B = a->B;
  kp.Ho = a->Ho;
  kp.Wo = a->Wo;
  kp.Po = a->Po;
  kp.Cout = a->N;
  kp.code = a->code;
  kp.bias = a->bias;
  kp.scale = a->scale;
  kp.shift = a->shift;
  if (a->code != nullptr)
    for (int i = 0; i < a->lut_len; ++i) kp.lut[i] = a->lut[i];
  kp.act = a->act;
  kp.slope = a->slope;
  kp.stats = a->stats;
  kp.gate = reinterpret_cast<const __nv_bfloat16*>(a->gate);
  kp.gate_slope = a->gate_slope;
  kp.addend = a->addend;
  TG_REQUIRE(a->addend == nullptr || f32, "tg_conv_igemm: addend is an fp32-path feature");
  kp.pool_out = reinterpret_cast<__nv_bfloat16*>(a->pool_out);
  kp.skip_out = a->skip_out;
  TG_REQUIRE(a->pool_out != nullptr || !a->skip_out, "tg_conv_igemm: skip_out needs pool_out");
  if (a->pool_out != nullptr) {
    // the fused 2x2 max-pool lives in the epilogue of the halo-reuse kernels (bf16, 3x3 / stride 1, N = 64 or 128, H % 16 == 0,
    // W % 8 == 0, no dgrad gate); tg_conv_pool_fusable() tells the caller in advance
    TG_REQUIRE(!f32 && a->gate == nullptr && halo_enabled() &&
                   (conv_halo_eligible(a) || (halo_stream_enabled() && conv_halo_stream_eligible(a))),
               "tg_conv_igemm: pool_out is not supported for this shape (see tg_conv_pool_fusable)");
  }
#ifdef TG_PERF_DEBUG
  {
    static const int dbg = [] { const char* e = getenv("TG_CONV_DEBUG"); return e ? atoi(e) : 0; }();
    kp.debug = dbg;
  }
#endif

  // small-N 3x3 stride-1 layers: halo-tile reuse + resident weights (conv_halo.cu); bf16 storage only
  if (!f32) {
    if (halo_enabled() && conv_halo_eligible(a)) return conv_halo_launch(a, kp, reinterpret_cast<cudaStream_t>(stream));
    if (halo_enabled() != 0 && halo_stream_enabled() && conv_halo_stream_eligible(a))
      return conv_halo_stream_launch(a, kp, reinterpret_cast<cudaStream_t>(stream));
    if (halo_enabled() != 0 && halo_stream_enabled() && conv_halo_s2dgrad_eligible(a))
      return conv_halo_s2dgrad_launch(a, kp, reinterpret_cast<cudaStream_t>(stream));
  }

  // A: channels-last activations as a 5-D tensor (C, W, H, P, B)
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[5] = {(uint64_t)a->C, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->P, (uint64_t)a->B};
    uint64_t str[4] = {(uint64_t)a->C * esz, (uint64_t)a->C * esz * a->W, (uint64_t)a->C * esz * a->W * a->H,
                       (uint64_t)a->C * esz * a->W * a->H * a->P};
    uint32_t box[5] = {(uint32_t)kb_elems, (uint32_t)kp.Wt, (uint32_t)kp.Ht, 1, (uint32_t)kp.Bt};
    if ((f32 ? make_tmap_f32 : make_tmap_bf16)(&tmA, a->x, 5, dims, str, box) != 0) return -3;
  }
  const long m_tiles_all = (long)kp.tiles_b * kp.tiles_h * kp.tiles_w;
  const int sms = num_sms();
  TG_REQUIRE(sms > 0, "tg_conv_igemm: no CUDA device");
  // N = 256 / 192 tiles on CTA pairs (cta_group::2) whenever there is at least one 256-pixel unit per pair of SMs
  const long pair_units = (long)kp.num_sub * (m_tiles_all / 2) * kp.n_tiles;
  static const bool pair192 = [] { const char* e = getenv("TG_CONV_PAIR_192"); return e != nullptr && e[0] == '1'; }();
  const bool pair = !f32 && (BN == 256 || (BN == 192 && pair192)) && conv_pair_enabled() && m_tiles_all % 2 == 0 &&
                    pair_units >= conv_pair_min_units(sms);
  {
    uint64_t dims[2] = {(uint64_t)a->Ktot, (uint64_t)a->N};
    uint64_t str[1] = {(uint64_t)a->Ktot * esz};
    uint32_t box[2] = {(uint32_t)kb_elems, (uint32_t)(pair ? BN / 2 : BN)};
    if ((f32 ? make_tmap_f32 : make_tmap_bf16)(&tmB, a->w, 2, dims, str, box) != 0) return -3;
  }
  if (pair) {
    const int grid2 = 2 * (int)(pair_units < sms / 2 ? pair_units : sms / 2);
    if (a->stats != nullptr) {
      TG_REQUIRE(a->stats_rows_cap >= grid2, "tg_conv_igemm: stats_rows_cap %d < grid %d", a->stats_rows_cap, grid2);
    }
    a->stats_rows_used = grid2;
    return BN == 256 ? launch_conv_pair<256>(tmA, tmB, kp, grid2, reinterpret_cast<cudaStream_t>(stream))
                     : launch_conv_pair<192>(tmA, tmB, kp, grid2, reinterpret_cast<cudaStream_t>(stream));
  }
  // two M-tiles per unit sharing the weight tile (N = 128 or 64): only when that still fills the machine
  const bool wide = !f32 && (BN == 128 || BN == 64) && conv_wide_enabled() && m_tiles_all % 2 == 0 &&
                    (long)kp.num_sub * (m_tiles_all / 2) * kp.n_tiles >= 2L * sms;
  const long total_tiles = (long)kp.num_sub * (wide ? m_tiles_all / 2 : m_tiles_all) * kp.n_tiles;
  int grid = (int)(total_tiles < sms ? total_tiles : sms);
  if (a->stats != nullptr) {
    TG_REQUIRE(a->stats_rows_cap >= grid, "tg_conv_igemm: stats_rows_cap %d < grid %d", a->stats_rows_cap, grid);
  }
  a->stats_rows_used = grid;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (f32) {
    if (BN == 256) return launch_conv<256, true>(a, tmA, tmB, kp, grid, st);
    if (BN == 192) return launch_conv<192, true>(a, tmA, tmB, kp, grid, st);
    if (BN == 128) return launch_conv<128, true>(a, tmA, tmB, kp, grid, st);
    return launch_conv<64, true>(a, tmA, tmB, kp, grid, st);
  }
  if (wide) return BN == 128 ? launch_conv<128, false, 2>(a, tmA, tmB, kp, grid, st)
                             : launch_conv<64, false, 2>(a, tmA, tmB, kp, grid, st);
  if (BN == 256) return launch_conv<256>(a, tmA, tmB, kp, grid, st);
  if (BN == 192) return launch_conv<192>(a, tmA, tmB, kp, grid, st);
  if (BN == 128) return launch_conv<128>(a, tmA, tmB, kp, grid, st);
  return launch_conv<64>(a, tmA, tmB, kp, grid, st);
}
