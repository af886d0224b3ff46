// wgrad_halo.cu — weight gradient of 3x3 / stride-1 convolutions with 64 output channels
// (dec1, dec2: reference pconv.py:30 via generator.py:27-28) with HALO-TILE REUSE.
//
// The generic kernel (wgrad_igemm.cu) loads, per 64-pixel K block and per tap pair, two shifted X boxes
// and the G box: ~24 KB of operands per 4 MMAs, L2-bound at ~300 TFLOP/s for N = 64 (profiles/r01_*).
// Here one work unit owns one 64-channel block of X and ALL nine taps: per 8x8-pixel K box the 10x10 halo
// of X (12.5 KB) and the G box (8 KB) are loaded once and nine shifted MN-major views of the halo feed five
// M=128 MMAs (tap pairs) per 16 pixels — the views only differ in their start row (address-based
// SWIZZLE_128B, measured by tools/probe/umma_probe.cu), the atom stride is the halo row (1280 B) and the
// leading-dimension offset between the two 64-channel halves of M is the distance between the two taps.
// Five accumulators (5 x 64 TMEM columns) live for the whole pixel range of the unit.
#include "conv_igemm.cuh"
#include "tg_common.cuh"
#include "../../include/terragan_b200.h"

namespace tg {

constexpr int kWhStages = 4;
constexpr int kWhXBytes = 13 * 1024;   // 100 halo pixels x 128 B, rounded up to the swizzle period
constexpr int kWhGBytes = 8 * 1024;    // 64 pixels x 128 B
constexpr int kWhStageBytes = kWhXBytes + kWhGBytes;
constexpr int kWhSmem = kWhStages * kWhStageBytes + 256 + 1024;

struct WgradHaloParams {
  int tiles_w, tiles_h, tiles_b;   // 8x8 pixel boxes
  int cin_blocks, splits, C, rows; // rows = 9 * C
  int row0[10];                    // halo row of tap t: (dh+1)*10 + (dw+1); entry 9 = dummy partner of tap 8
  float* partial;                  // [splits][9*C][64]
};

__global__ void __launch_bounds__(256, 1)
wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG,
                  const __grid_constant__ WgradHaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWhStages * kWhStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kWhStages;
  uint64_t* tfull_bar = bars + 2 * kWhStages;
  uint64_t* tempty_bar = bars + 2 * kWhStages + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWhStages + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmG);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kWhStages; ++i) {
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
  const int total_units = p.cin_blocks * p.splits;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
        const int cb = u % p.cin_blocks, sp = u / p.cin_blocks;
        const int kb_end = min(kboxes, (sp + 1) * per_split);
        for (int kb = sp * per_split; kb < kb_end; ++kb) {
          const int tw = kb % p.tiles_w, th = (kb / p.tiles_w) % p.tiles_h, tb = kb / (p.tiles_w * p.tiles_h);
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sx = smem + stage * kWhStageBytes;
          mbar_arrive_expect_tx(&full_bar[stage], 100 * 128 + kWhGBytes);
          tma_load_5d(sx, &tmX, &full_bar[stage], cb * 64, tw * 8 - 1, th * 8 - 1, 0, tb);
          tma_load_5d(sx + kWhXBytes, &tmG, &full_bar[stage], 0, tw * 8, th * 8, 0, tb);
          if (++stage == kWhStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    {   // whole warp, warp-uniform addressing; one elected lane issues the tcgen05 instructions
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, true, true);
      // Per tap pair j: A starts at halo row row0[2j] (+ 2*10 rows per 16-pixel step) with the second tap LBO =
      // (row0[2j+1] - row0[2j]) rows further. All descriptor words are formed once; the loop adds offsets >> 4.
      const uint32_t smem0 = smem_u32(smem);
      uint32_t a_off[5], a_hi[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const int r0 = p.row0[2 * j], r1 = p.row0[2 * j + 1];
        const uint64_t d = make_smem_desc(smem0, (r1 - r0) * 128, 1280);
        a_hi[j] = static_cast<uint32_t>(d >> 32);
        // lo word = start | LBO << 16: keep the LBO part here, the start address is added per stage
        a_off[j] = (static_cast<uint32_t>(d) - ((smem0 & 0x3FFFF) >> 4)) + static_cast<uint32_t>(r0) * 8u;
      }
      const uint64_t dg = make_smem_desc(smem0, 1024, 1024);
      const uint32_t g_hi = static_cast<uint32_t>(dg >> 32);
      const uint32_t x_lo0 = (smem0 & 0x3FFFF) >> 4;
      const uint32_t g_rel = (static_cast<uint32_t>(dg) - x_lo0) + (kWhXBytes >> 4);
      int stage = 0;
      uint32_t phase = 0, uphase = 0;
      for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
        const int sp = u / p.cin_blocks;
        const int kb_begin = sp * per_split, kb_end = min(kboxes, kb_begin + per_split);
        mbar_wait(tempty_bar, uphase ^ 1);
        tc_fence_after();
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t x_lo = x_lo0 + static_cast<uint32_t>(stage) * (kWhStageBytes >> 4);
          const uint32_t g_lo = x_lo + g_rel;
          const uint32_t accum = static_cast<uint32_t>(kb > kb_begin);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {        // 16 pixels = two rows of the 8x8 box per MMA
#pragma unroll
              for (int j = 0; j < 5; ++j)        // tap pairs (2j, 2j+1)
                umma_bf16_lh(tmem_base + j * 64, x_lo + a_off[j] + k * 160, a_hi[j], g_lo + k * 128, g_hi, idesc,
                             k ? 1u : accum);
            }
            umma_commit(&empty_bar[stage]);
          }
          if (++stage == kWhStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma_commit(tfull_bar);
        uphase ^= 1;
      }
    }
  } else if (warp >= 4) {
    const int q = warp - 4;
    uint32_t uphase = 0;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
      const int cb = u % p.cin_blocks, sp = u / p.cin_blocks;
      mbar_wait(tfull_bar, uphase);
      tc_fence_after();
      const int r = q * 32 + lane;
#pragma unroll 1
      for (int j = 0; j < 5; ++j) {
        const int tap = 2 * j + (r >> 6);
        const bool valid = tap < 9;
        float* dst = p.partial + (static_cast<long>(sp) * p.rows + (valid ? tap : 0) * p.C + cb * 64 + (r & 63)) * 64;
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + j * 64;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t raw[32];
          tmem_ld_32x32(t_addr + ch * 32, raw);
          tmem_ld_wait();
          if (valid) {
            uint4* d4 = reinterpret_cast<uint4*>(dst + ch * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i)
              d4[i] = make_uint4(raw[4 * i], raw[4 * i + 1], raw[4 * i + 2], raw[4 * i + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar);
      uphase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

bool wgrad_halo_shape_ok(int Ho, int Wo, int num_taps, int N) {
  return num_taps == 9 && N == 64 && Ho % 8 == 0 && Wo % 8 == 0;
}

int wgrad_halo_splits(int B, int Ho, int Wo, int C, int sms) {
  const long kboxes = (long)B * (Ho / 8) * (Wo / 8);
  const int cin_blocks = C / 64;
  long splits = sms / cin_blocks;               // units = cin_blocks * splits <= SM count: a single wave
  const long max_by_k = (kboxes + 15) / 16;
  if (splits > max_by_k) splits = max_by_k;
  if (splits < 1) splits = 1;
  long per = (kboxes + splits - 1) / splits;
  long eff = (kboxes + per - 1) / per;          // drop empty trailing shares
  return (int)eff;
}

bool wgrad_halo_eligible(const tg_wgrad_args* a) {
  if (!wgrad_halo_shape_ok(a->Ho, a->Wo, a->num_taps, a->N)) return false;
  if (a->P != 1 || a->H != a->Ho || a->W != a->Wo) return false;
  for (int t = 0; t < 9; ++t) {
    if (a->tap_plane[t] != 0 || a->tap_dh[t] < -1 || a->tap_dh[t] > 1 || a->tap_dw[t] < -1 || a->tap_dw[t] > 1)
      return false;
    // taps must come in raster order so that the partner of each pair lies at a higher address
    if (t > 0 && (a->tap_dh[t] * 3 + a->tap_dw[t]) <= (a->tap_dh[t - 1] * 3 + a->tap_dw[t - 1])) return false;
  }
  return true;
}

int wgrad_halo_launch(tg_wgrad_args* a, cudaStream_t st) {
  const int sms = num_sms();
  WgradHaloParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.tiles_w = a->Wo / 8;
  kp.tiles_h = a->Ho / 8;
  kp.tiles_b = a->B;
  kp.cin_blocks = a->C / 64;
  kp.splits = wgrad_halo_splits(a->B, a->Ho, a->Wo, a->C, sms);
  kp.C = a->C;
  kp.rows = 9 * a->C;
  for (int t = 0; t < 9; ++t) kp.row0[t] = (a->tap_dh[t] + 1) * 10 + (a->tap_dw[t] + 1);
  kp.row0[9] = kp.row0[8] + 1;   // dummy partner of the last tap (rows discarded)
  kp.partial = a->partial;
  TG_REQUIRE((int64_t)kp.splits * kp.rows * 64 <= a->partial_cap,
             "tg_wgrad_igemm: partial workspace too small (%lld floats needed)", (long long)kp.splits * kp.rows * 64);
  a->splits = kp.splits;
  CUtensorMap tmX, tmG;
  {
    uint64_t dims[5] = {(uint64_t)a->C, (uint64_t)a->W, (uint64_t)a->H, 1, (uint64_t)a->B};
    uint64_t str[4] = {(uint64_t)a->C * 2, (uint64_t)a->C * 2 * a->W, (uint64_t)a->C * 2 * a->W * a->H,
                       (uint64_t)a->C * 2 * a->W * a->H};
    uint32_t box[5] = {64, 10, 10, 1, 1};
    if (make_tmap_bf16(&tmX, a->x, 5, dims, str, box) != 0) return -3;
  }
  {
    uint64_t dims[5] = {64, (uint64_t)a->Wo, (uint64_t)a->Ho, 1, (uint64_t)a->B};
    uint64_t str[4] = {128, (uint64_t)128 * a->Wo, (uint64_t)128 * a->Wo * a->Ho, (uint64_t)128 * a->Wo * a->Ho};
    uint32_t box[5] = {64, 8, 8, 1, 1};
    if (make_tmap_bf16(&tmG, a->g, 5, dims, str, box) != 0) return -3;
  }
  TG_SET_SMEM_ONCE((wgrad_halo_kernel), kWhSmem);
  const int units = kp.cin_blocks * kp.splits;
  const int grid = units < sms ? units : sms;
  wgrad_halo_kernel<<<grid, 256, kWhSmem, st>>>(tmX, tmG, kp);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace tg
