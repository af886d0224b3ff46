// wgrad_igemm.cu — tcgen05 implicit-GEMM weight gradient for sm_100a.
//
// Replaces autograd's convolution_backward(weight) for PConv2d.input_conv (reference
// mvp_gan/src/models/pconv.py:30) and the Discriminator convs (discriminator.py:11):
//   dW[co][ci][kh][kw] = sum_{b,ho,wo} g[b][ho][wo][co] * xm[b][(ho,wo) (+) tap(kh,kw)][ci]
// where xm is the masked layer input the forward pass consumed and g already carries the
// BatchNorm/ReLU/mask-ratio backward factors (pconv.py:43-48).
//
// GEMM view: D^T[(tap,ci)][co], M = 128 rows = two 64-row blocks of the (tap, ci-block) list,
// N = Cout tile, K = output pixels (64 per K block = a Bt x Ht x Wt pixel box). Both operands are
// channels-last tiles [pixel][64 channels] brought in by TMA boxes (the X tile shifted by the tap
// offset; zero fill = conv padding), i.e. MN-major SWIZZLE_128B UMMA operands: no transposes are
// ever materialised. The pixel axis is split over CTAs (split-K); each share is written to an fp32
// partial buffer and tg_wgrad_reduce sums the shares in a fixed order (deterministic) while
// scattering into PyTorch's [Cout][Cin][kh][kw] gradient layout.
#include <stdlib.h>

#include "conv_igemm.cuh"
#include "tg_common.cuh"
#include "../../include/terragan_b200.h"

namespace tg {

constexpr int kWStages = 4;
constexpr int kBoxBytes = 64 * 128;  // 64 pixels x 64 bf16

// kF32 (verification path): fp32 tiles [pixel][32 channels] (the same 128-byte rows), kind::tf32 MMAs with K = 8
// pixels each; M = 128 rows = FOUR 32-channel blocks of the (tap, channel block) list.
template <int BN, bool kF32 = false>
struct WgradSmem {
  static constexpr int kCB = kF32 ? 32 : 64;                 // channels per TMA box
  static constexpr int kABoxes = 128 / kCB;
  static constexpr int kStages = kF32 ? 3 : kWStages;
  static constexpr int kABytes = kABoxes * kBoxBytes;
  static constexpr int kBBytes = (BN / kCB) * kBoxBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTiles = kStages * kStageBytes;
  static constexpr int kTotal = kTiles + 256 + 1024;
};

template <int BN, bool kF32 = false>
__global__ void __launch_bounds__(256, 1)
wgrad_igemm_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG,
                   const __grid_constant__ WgradKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  using S = WgradSmem<BN, kF32>;
  constexpr int kWStages = S::kStages;     // shadows the namespace constant
  constexpr int kCB = S::kCB, kABoxes = S::kABoxes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kTiles);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kWStages;
  uint64_t* tfull_bar = bars + 2 * kWStages;
  uint64_t* tempty_bar = bars + 2 * kWStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmG);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kWStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int kboxes = p.tiles_b * p.tiles_h * p.tiles_w;
  const int per_split = (kboxes + p.splits - 1) / p.splits;
  const int total_units = p.m_tiles * p.n_tiles * p.splits;
  // unit -> (split, m tile, n tile), n fastest so CTAs sharing an X tile run together

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
        const int nt = u % p.n_tiles;
        const int mt = (u / p.n_tiles) % p.m_tiles;
        const int sp = u / (p.n_tiles * p.m_tiles);
        WgradBlk e[kABoxes];
#pragma unroll
        for (int j = 0; j < kABoxes; ++j)
          e[j] = p.blks[(kABoxes * mt + j < p.num_blk) ? kABoxes * mt + j : kABoxes * mt];
        const int kb_begin = sp * per_split;
        const int kb_end = min(kboxes, kb_begin + per_split);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          const int tw = kb % p.tiles_w;
          const int th = (kb / p.tiles_w) % p.tiles_h;
          const int tb = kb / (p.tiles_w * p.tiles_h);
          const int w0 = tw * p.Wt, h0 = th * p.Ht, b0 = tb * p.Bt;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * S::kStageBytes;
          uint8_t* sb = sa + S::kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], S::kStageBytes);
#pragma unroll
          for (int j = 0; j < kABoxes; ++j)
            tma_load_5d(sa + j * kBoxBytes, &tmX, &full_bar[stage], e[j].cb * kCB, w0 + e[j].dw, h0 + e[j].dh,
                        e[j].plane, b0);
#pragma unroll
          for (int j = 0; j < BN / kCB; ++j)
            tma_load_5d(sb + j * kBoxBytes, &tmG, &full_bar[stage], nt * BN + j * kCB, w0, h0, 0, b0);
          if (++stage == kWStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    {   // whole warp, warp-uniform addressing; one elected lane issues the tcgen05 instructions
      constexpr uint32_t idesc = kF32 ? make_idesc_tf32(128, BN, true, true) : make_idesc_bf16(128, BN, true, true);
      // both operands MN-major, same strides; fp32 tiles use 32-byte swizzle atoms over 4-row (512 B) groups
      const uint64_t d0 = kF32 ? make_smem_desc(smem_u32(smem), kBoxBytes, 512, 1) : make_smem_desc(smem_u32(smem), kBoxBytes, 1024);
      const uint32_t d_lo0 = static_cast<uint32_t>(d0), d_hi = static_cast<uint32_t>(d0 >> 32);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
        const int sp = u / (p.n_tiles * p.m_tiles);
        const int kb_begin = sp * per_split;
        const int kb_end = min(kboxes, kb_begin + per_split);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          // 16 pixels (= 16 rows of 128 B = two 8-row swizzle atoms) per MMA: +2048 B = +128 in the lo word
          const uint32_t a_lo = d_lo0 + static_cast<uint32_t>(stage) * (S::kStageBytes >> 4);
          const uint32_t b_lo = a_lo + (S::kABytes >> 4);
          if (elect_one()) {
            if (kF32) {   // 8 pixels (one 8-row swizzle atom, 1024 B) per MMA
#pragma unroll
              for (int k = 0; k < 8; ++k)
                umma_tf32_lh(d_tmem, a_lo + k * 64, d_hi, b_lo + k * 64, d_hi, idesc, k ? 1u : static_cast<uint32_t>(kb > kb_begin));
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_lh(d_tmem, a_lo + k * 128, d_hi, b_lo + k * 128, d_hi, idesc, k ? 1u : static_cast<uint32_t>(kb > kb_begin));
            }
            umma_commit(&empty_bar[stage]);
          }
          if (++stage == kWStages) {
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
    const int q = warp - 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
      const int nt = u % p.n_tiles;
      const int mt = (u / p.n_tiles) % p.m_tiles;
      const int sp = u / (p.n_tiles * p.m_tiles);
      const int r = q * 32 + lane;
      const int bi = kABoxes * mt + r / kCB;
      const bool valid = bi < p.num_blk;
      const int row = valid ? p.blks[bi].row + (r % kCB) : 0;
      float* dst = p.partial + (static_cast<long>(sp) * p.rows + row) * p.Cout + nt * BN;

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int ch = 0; ch < BN / 32; ++ch) {
        uint32_t raw[32];
        tmem_ld_32x32(t_addr + ch * 32, raw);
        tmem_ld_wait();
        if (valid) {
          uint4* d4 = reinterpret_cast<uint4*>(dst + ch * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            d4[j] = make_uint4(raw[4 * j], raw[4 * j + 1], raw[4 * j + 2], raw[4 * j + 3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc<kTmemCols>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// "Wide" variant (bf16): kMT M-tiles — 2 kMT (tap, channel-block) rows of dW^T — share every gradient tile.
// The kernel above moves one G tile per M-tile: at N = 256 that is 16 KB of X + 32 KB of G per four MMAs
// (94 B/clk/SM from L2, 67 % tensor pipe in ncu), at N = 128 16 + 16 KB per 256 clocks (128 B/clk, 41 %).
// Here a unit keeps kMT accumulators (kMT x BN = 512 TMEM columns: no double buffering, which a weight gradient does
// not need — its epilogue runs once per unit after hundreds of K blocks) and re-uses each G tile kMT times:
// N = 256: 32 + 32 KB per 8 MMAs (62 B/clk); N = 128: 64 + 16 KB per 16 MMAs (78 B/clk).
// ------------------------------------------------------------------------------------------------
template <int BN, int kMT>
struct WgradWideSmem {
  static constexpr int kABytes = kMT * 2 * kBoxBytes;
  static constexpr int kBBytes = (BN / 64) * kBoxBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (216 * 1024) / kStageBytes >= 4 ? 4 : (216 * 1024) / kStageBytes;
  static constexpr int kTiles = kStages * kStageBytes;
  static constexpr int kTotal = kTiles + 256 + 1024;
  static_assert(kStages >= 2, "at least two stages");
  static_assert(kMT * BN <= 512, "accumulators must fit the 512 TMEM columns");
};

template <int BN, int kMT>
__global__ void __launch_bounds__(256, 1)
wgrad_wide_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG,
                  const __grid_constant__ WgradKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  using S = WgradWideSmem<BN, kMT>;
  constexpr int kStages = S::kStages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kTiles);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tfull_bar = bars + 2 * kStages;
  uint64_t* tempty_bar = bars + 2 * kStages + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 2);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmG);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 4);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int kboxes = p.tiles_b * p.tiles_h * p.tiles_w;
  const int per_split = (kboxes + p.splits - 1) / p.splits;
  const int m_super = (p.m_tiles + kMT - 1) / kMT;
  const int total_units = m_super * p.n_tiles * p.splits;
  // unit -> (split, super M tile, n tile), n fastest so CTAs sharing X tiles run together

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
        const int nt = u % p.n_tiles;
        const int smt = (u / p.n_tiles) % m_super;
        const int sp = u / (p.n_tiles * m_super);
        WgradBlk e[2 * kMT];
#pragma unroll
        for (int j = 0; j < 2 * kMT; ++j) {
          const int bi = smt * 2 * kMT + j;
          e[j] = p.blks[bi < p.num_blk ? bi : p.num_blk - 1];      // rows past the table are loaded but never stored
        }
        const int kb_begin = sp * per_split;
        const int kb_end = min(kboxes, kb_begin + per_split);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          const int tw = kb % p.tiles_w;
          const int th = (kb / p.tiles_w) % p.tiles_h;
          const int tb = kb / (p.tiles_w * p.tiles_h);
          const int w0 = tw * p.Wt, h0 = th * p.Ht, b0 = tb * p.Bt;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * S::kStageBytes;
          uint8_t* sb = sa + S::kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], S::kStageBytes);
#pragma unroll
          for (int j = 0; j < 2 * kMT; ++j)
            tma_load_5d(sa + j * kBoxBytes, &tmX, &full_bar[stage], e[j].cb * 64, w0 + e[j].dw, h0 + e[j].dh, e[j].plane, b0);
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            tma_load_5d(sb + j * kBoxBytes, &tmG, &full_bar[stage], nt * BN + j * 64, w0, h0, 0, b0);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, BN, true, true);
    const uint64_t d0 = make_smem_desc(smem_u32(smem), kBoxBytes, 1024);   // both operands MN-major, same strides
    const uint32_t d_lo0 = static_cast<uint32_t>(d0), d_hi = static_cast<uint32_t>(d0 >> 32);
    int stage = 0;
    uint32_t phase = 0, ucount = 0;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++ucount) {
      const int sp = u / (p.n_tiles * m_super);
      const int kb_begin = sp * per_split;
      const int kb_end = min(kboxes, kb_begin + per_split);
      mbar_wait(tempty_bar, (ucount & 1u) ^ 1u);       // the previous unit's accumulators have been drained
      tc_fence_after();
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_lo = d_lo0 + static_cast<uint32_t>(stage) * (S::kStageBytes >> 4);
        const uint32_t b_lo = a_lo + (S::kABytes >> 4);
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < kMT; ++j) {
#pragma unroll
            for (int k = 0; k < 4; ++k)     // 16 pixels (two 8-row swizzle atoms, 2048 B) per MMA
              umma_bf16_lh(tmem_base + j * BN, a_lo + j * (2 * kBoxBytes >> 4) + k * 128, d_hi, b_lo + k * 128, d_hi, idesc,
                           k ? 1u : static_cast<uint32_t>(kb > kb_begin));
          }
          umma_commit(&empty_bar[stage]);
        }
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) umma_commit(tfull_bar);
    }
  } else if (warp >= 4) {
    const int q = warp - 4;
    uint32_t ucount = 0;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++ucount) {
      const int nt = u % p.n_tiles;
      const int smt = (u / p.n_tiles) % m_super;
      const int sp = u / (p.n_tiles * m_super);
      const int r = q * 32 + lane;
      mbar_wait(tfull_bar, ucount & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < kMT; ++j) {
        const int bi = (smt * kMT + j) * 2 + (r >> 6);
        const bool valid = bi < p.num_blk;
        const int row = valid ? p.blks[bi].row + (r & 63) : 0;
        float* dst = p.partial + (static_cast<long>(sp) * p.rows + row) * p.Cout + nt * BN;
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + j * BN;
#pragma unroll 1
        for (int ch = 0; ch < BN / 32; ++ch) {
          uint32_t raw[32];
          tmem_ld_32x32(t_addr + ch * 32, raw);
          tmem_ld_wait();
          if (valid) {
            uint4* d4 = reinterpret_cast<uint4*>(dst + ch * 32);
#pragma unroll
            for (int jj = 0; jj < 8; ++jj)
              d4[jj] = make_uint4(raw[4 * jj], raw[4 * jj + 1], raw[4 * jj + 2], raw[4 * jj + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar);
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

// partial[s][tap*C + c][n]  ->  dw[n][c][perm[tap]]  (kk = number of kernel positions)
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int taps, int C,
                                    int N, const int32_t* __restrict__ perm, float* __restrict__ dw,
                                    int accumulate) {
  // one thread per (row, n); threads along n are contiguous in `partial` (coalesced reads)
  const long total = static_cast<long>(taps) * C * N;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i % N);
    const long row = i / N;
    const int c = static_cast<int>(row % C);
    const int t = static_cast<int>(row / C);
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += partial[static_cast<long>(s) * total + i];
    const long o = (static_cast<long>(n) * C + c) * taps + perm[t];
    dw[o] = accumulate ? dw[o] + acc : acc;
  }
}

bool wgrad_halo_shape_ok(int Ho, int Wo, int num_taps, int N);       // wgrad_halo.cu
int wgrad_halo_splits(int B, int Ho, int Wo, int C, int sms);
bool wgrad_halo_eligible(const tg_wgrad_args* a);
int wgrad_halo_launch(tg_wgrad_args* a, cudaStream_t st);

static bool wgrad_halo_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TG_NO_HALO");
    v = (e != nullptr && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

static void choose_kbox(int Ho, int Wo, int* Bt, int* Ht, int* Wt) {
  int wt = 1;
  while (wt * 2 <= Wo && wt * 2 <= 16) wt *= 2;
  int ht = 1;
  while (ht * 2 <= Ho && wt * ht * 2 <= 64) ht *= 2;
  *Wt = wt;
  *Ht = ht;
  *Bt = 64 / (wt * ht);
}

// Split-K factor over the pixel boxes. The kernel is persistent (one CTA per SM, static round-robin over
// units = m tiles x n tiles x splits), so its time is waves x (K blocks per unit + a fixed per-unit cost), with
// waves = ceil(units / SMs). The round-1 rule "about 2 units per SM" produced 297 and 299 units on 148 SMs for dec3
// and enc2 -- a third, almost empty wave (535 / 525 TFLOP/s). Every candidate is now costed and the cheapest
// taken; ties go to the smaller split (fewer fp32 partial shares for tg_wgrad_reduce to read back).
static int wgrad_bn(int N, bool f32) {
  if (f32) return (N % 128 == 0) ? 128 : (N % 64 == 0) ? 64 : 32;
  return (N % 256 == 0) ? 256 : (N % 128 == 0) ? 128 : 64;
}

static int choose_splits(int B, int Ho, int Wo, int num_blk, int N, int sms, bool f32 = false, int mt_per_unit = 1) {
  int Bt, Ht, Wt;
  choose_kbox(Ho, Wo, &Bt, &Ht, &Wt);
  const long kboxes = (long)((B + Bt - 1) / Bt) * ((Ho + Ht - 1) / Ht) * ((Wo + Wt - 1) / Wt);
  const int BN = wgrad_bn(N, f32);
  const int per_m = (f32 ? 4 : 2) * mt_per_unit;     // (tap, channel-block) rows of dW^T per unit
  const long mn = (long)((num_blk + per_m - 1) / per_m) * (N / BN);
  long max_splits = (kboxes + 3) / 4;            // at least 4 K blocks per unit
  if (max_splits > 256) max_splits = 256;
  if (max_splits < 1) max_splits = 1;
  const double unit_fixed = 3.0;                 // accumulator drain + pipeline fill, in K-block times
  const double reduce_per_split = 0.75 * (double)BN / 128.0;   // partial write + read-back per share, same unit
  long best = 1;
  double best_cost = 1e30;
  for (long s = 1; s <= max_splits; ++s) {
    const long per = (kboxes + s - 1) / s;
    if ((kboxes + per - 1) / per != s) continue;  // would leave an empty trailing share
    const long waves = (mn * s + sms - 1) / sms;
    const double cost = (double)waves * ((double)per + unit_fixed) + reduce_per_split * (double)s / (double)waves;
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = s;
    }
  }
  return (int)best;
}

static bool wgrad_wide_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TG_NO_WGRAD_WIDE");
    v = (e != nullptr && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}
static int wide_mt(int BN) {      // M-tiles per unit: kMT * BN <= 512 TMEM columns
  if (BN == 256) return 2;
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TG_WGRAD_WIDE_MT");
    v = (e != nullptr && e[0] == '4') ? 4 : 2;      // 2 measured faster than 4 on every layer (more pipeline stages)
  }
  return v;
}

template <int BN, int kMT>
static int launch_wgrad_wide(const CUtensorMap& tmX, const CUtensorMap& tmG, const WgradKParams& kp, int grid,
                             cudaStream_t st) {
  using S = WgradWideSmem<BN, kMT>;
  TG_SET_SMEM_ONCE((wgrad_wide_kernel<BN, kMT>), S::kTotal);
  wgrad_wide_kernel<BN, kMT><<<grid, 256, S::kTotal, st>>>(tmX, tmG, kp);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int BN, bool kF32 = false>
static int launch_wgrad(const CUtensorMap& tmX, const CUtensorMap& tmG, const WgradKParams& kp,
                        int grid, cudaStream_t st) {
  using S = WgradSmem<BN, kF32>;
  TG_SET_SMEM_ONCE((wgrad_igemm_kernel<BN, kF32>), S::kTotal);
  wgrad_igemm_kernel<BN, kF32><<<grid, 256, S::kTotal, st>>>(tmX, tmG, kp);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace tg

extern "C" int64_t tg_wgrad_partial_floats_f32(int B, int Ho, int Wo, int num_taps, int C, int N) {
  const int sms = tg::num_sms() > 0 ? tg::num_sms() : 148;
  return (int64_t)tg::choose_splits(B, Ho, Wo, num_taps * (C / 32), N, sms, true) * num_taps * C * N;
}

extern "C" int64_t tg_wgrad_partial_floats(int B, int Ho, int Wo, int num_taps, int C, int N) {
  const int sms = tg::num_sms() > 0 ? tg::num_sms() : 148;
  int splits = tg::choose_splits(B, Ho, Wo, num_taps * (C / 64), N, sms);
  {
    const int ws = tg::choose_splits(B, Ho, Wo, num_taps * (C / 64), N, sms, false, tg::wide_mt(tg::wgrad_bn(N, false)));
    if (ws > splits) splits = ws;
  }
  if (tg::wgrad_halo_shape_ok(Ho, Wo, num_taps, N)) {
    const int hs = tg::wgrad_halo_splits(B, Ho, Wo, C, sms);
    if (hs > splits) splits = hs;
  }
  return (int64_t)splits * num_taps * C * N;
}

extern "C" int tg_wgrad_igemm(tg_wgrad_args* a, void* stream) {
  using namespace tg;
  TG_REQUIRE(a != nullptr, "tg_wgrad_igemm: null args");
  TG_REQUIRE(a->dtype == TG_DTYPE_BF16 || a->dtype == TG_DTYPE_F32, "tg_wgrad_igemm: bad dtype %d", a->dtype);
  const bool f32 = a->dtype == TG_DTYPE_F32;
  const int cb_elems = f32 ? 32 : 64, esz = f32 ? 4 : 2;
  TG_REQUIRE(a->C > 0 && a->C % cb_elems == 0, "tg_wgrad_igemm: C=%d must be a multiple of %d", a->C, cb_elems);
  TG_REQUIRE(a->N > 0 && a->N % cb_elems == 0, "tg_wgrad_igemm: N=%d must be a multiple of %d", a->N, cb_elems);
  TG_REQUIRE(a->num_taps >= 1 && a->num_taps <= TG_MAX_TAPS, "tg_wgrad_igemm: bad num_taps");
  const int sms = num_sms();
  TG_REQUIRE(sms > 0, "tg_wgrad_igemm: no CUDA device");
  if (!f32 && wgrad_halo_enabled() && wgrad_halo_eligible(a)) return wgrad_halo_launch(a, reinterpret_cast<cudaStream_t>(stream));
  const int BN = wgrad_bn(a->N, f32);
  const int num_blk = a->num_taps * (a->C / cb_elems);

  WgradKParams kp;
  memset(&kp, 0, sizeof(kp));
  choose_kbox(a->Ho, a->Wo, &kp.Bt, &kp.Ht, &kp.Wt);
  kp.tiles_w = (a->Wo + kp.Wt - 1) / kp.Wt;
  kp.tiles_h = (a->Ho + kp.Ht - 1) / kp.Ht;
  kp.tiles_b = (a->B + kp.Bt - 1) / kp.Bt;
  kp.n_tiles = a->N / BN;
  kp.num_blk = num_blk;
  kp.m_tiles = f32 ? (num_blk + 3) / 4 : (num_blk + 1) / 2;
  const bool wide = !f32 && wgrad_wide_enabled() && kp.m_tiles >= 2;
  const int kmt = wide ? wide_mt(BN) : 1;
  kp.splits = choose_splits(a->B, a->Ho, a->Wo, num_blk, a->N, sms, f32, kmt);
  kp.Cout = a->N;
  kp.rows = a->num_taps * a->C;
  kp.partial = a->partial;
  kp.blks = reinterpret_cast<const WgradBlk*>(a->blks);
  TG_REQUIRE(a->blks != nullptr && a->num_blk == num_blk, "tg_wgrad_igemm: blks table must hold num_taps*C/%d = %d entries", cb_elems, num_blk);
  TG_REQUIRE((int64_t)kp.splits * kp.rows * a->N <= a->partial_cap,
             "tg_wgrad_igemm: partial workspace too small (%lld floats needed)",
             (long long)((int64_t)kp.splits * kp.rows * a->N));
  a->splits = kp.splits;

  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

  CUtensorMap tmX, tmG;
  {
    uint64_t dims[5] = {(uint64_t)a->C, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->P, (uint64_t)a->B};
    uint64_t str[4] = {(uint64_t)a->C * esz, (uint64_t)a->C * esz * a->W, (uint64_t)a->C * esz * a->W * a->H,
                       (uint64_t)a->C * esz * a->W * a->H * a->P};
    uint32_t box[5] = {(uint32_t)cb_elems, (uint32_t)kp.Wt, (uint32_t)kp.Ht, 1, (uint32_t)kp.Bt};
    if ((f32 ? make_tmap_f32_atom32 : make_tmap_bf16)(&tmX, a->x, 5, dims, str, box) != 0) return -3;
  }
  {
    uint64_t dims[5] = {(uint64_t)a->N, (uint64_t)a->Wo, (uint64_t)a->Ho, 1, (uint64_t)a->B};
    uint64_t str[4] = {(uint64_t)a->N * esz, (uint64_t)a->N * esz * a->Wo, (uint64_t)a->N * esz * a->Wo * a->Ho,
                       (uint64_t)a->N * esz * a->Wo * a->Ho};
    uint32_t box[5] = {(uint32_t)cb_elems, (uint32_t)kp.Wt, (uint32_t)kp.Ht, 1, (uint32_t)kp.Bt};
    if ((f32 ? make_tmap_f32_atom32 : make_tmap_bf16)(&tmG, a->g, 5, dims, str, box) != 0) return -3;
  }
  const long total_units = (long)((kp.m_tiles + kmt - 1) / kmt) * kp.n_tiles * kp.splits;
  const int grid = (int)(total_units < sms ? total_units : sms);
  if (wide) {
    if (BN == 256) return launch_wgrad_wide<256, 2>(tmX, tmG, kp, grid, st);
    if (BN == 128) return kmt == 4 ? launch_wgrad_wide<128, 4>(tmX, tmG, kp, grid, st)
                                   : launch_wgrad_wide<128, 2>(tmX, tmG, kp, grid, st);
    return kmt == 4 ? launch_wgrad_wide<64, 4>(tmX, tmG, kp, grid, st) : launch_wgrad_wide<64, 2>(tmX, tmG, kp, grid, st);
  }
  if (f32) {
    if (BN == 128) return launch_wgrad<128, true>(tmX, tmG, kp, grid, st);
    if (BN == 64) return launch_wgrad<64, true>(tmX, tmG, kp, grid, st);
    return launch_wgrad<32, true>(tmX, tmG, kp, grid, st);
  }
  if (BN == 256) return launch_wgrad<256>(tmX, tmG, kp, grid, st);
  if (BN == 128) return launch_wgrad<128>(tmX, tmG, kp, grid, st);
  return launch_wgrad<64>(tmX, tmG, kp, grid, st);
}

extern "C" int tg_wgrad_reduce(const float* partial, int splits, int num_taps, int C, int N,
                               const int32_t* tap_perm_dev, float* dw, int accumulate, void* stream) {
  using namespace tg;
  TG_REQUIRE(partial && dw && tap_perm_dev, "tg_wgrad_reduce: null pointer");
  const long total = (long)num_taps * C * N;
  int grid = (int)((total + 255) / 256);
  const int cap = num_sms() * 16;
  if (grid > cap) grid = cap;
  wgrad_reduce_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      partial, splits, num_taps, C, N, tap_perm_dev, dw, accumulate);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}
