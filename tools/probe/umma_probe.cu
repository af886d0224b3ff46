// umma_probe.cu — hardware probe: do tcgen05 SWIZZLE_128B descriptors accept operand tiles that start at an
// arbitrary 128-byte row of a TMA-written buffer (halo reuse), with an arbitrary stride between 8-row atoms?
// out[cfg][m][n] for K-major A:   sum_k X[row0 + (m/8)*(sbo/128) + m%8][k] * W[n][k]
// MN-major A (rows = K):          sum_k X[row0 + (k/8)*(sbo/128) + k%8][m] * Wt[...]
#include <cstdio>
#include <cstring>
#include "../../terra-gan_b200/csrc/tg_common.cuh"

using namespace tg;

struct ProbeCfg { int row0; int sbo; int base_off; int mn_major; };

__device__ __forceinline__ uint64_t desc_bo(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t bo) {
  uint64_t d = make_smem_desc(addr, lbo, sbo);
  d |= static_cast<uint64_t>(bo & 7) << 49;
  return d;
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, ProbeCfg cfg,
             float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;                 // 512 rows x 128 B = 64 KB
  uint8_t* sW = smem + 65536;         // 128 rows x 128 B = 16 KB (two 64-row blocks)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536 + 16384);
  uint64_t* mbar = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(mbar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<128>(slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 65536 + 16384);
    tma_load_2d(sX, &tmX, bar, 0, 0);
    tma_load_2d(sX + 32768, &tmX, bar, 0, 256);
    tma_load_2d(sW, &tmW, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t xa = smem_u32(sX) + cfg.row0 * 128;
    const uint32_t wa = smem_u32(sW);
    if (!cfg.mn_major) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, false, false);
      for (int k = 0; k < 4; ++k)
        umma_bf16(tm, desc_bo(xa, 16, cfg.sbo, cfg.base_off) + 2 * k, make_smem_desc(wa, 16, 1024) + 2 * k, idesc, k);
    } else {
      // A = X rows as K (MN-major, M = 64 channels... use M=128 by pairing the same block twice via LBO = 0)
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, true, false);
      for (int k = 0; k < 4; ++k) {
        const uint32_t a_k = xa + k * 2 * cfg.sbo;   // 16 K rows = 2 atoms
        umma_bf16(tm, desc_bo(a_k, 32768, cfg.sbo, cfg.base_off), make_smem_desc(wa, 16, 1024) + 2 * k, idesc, k);
      }
    }
    umma_commit(mbar);
  }
  __syncthreads();
  mbar_wait(mbar, 0);
  tc_fence_after();
  {
    const int r = warp * 32 + lane;
    for (int ch = 0; ch < 2; ++ch) {
      uint32_t v[32];
      tmem_ld_32x32(tm + (static_cast<uint32_t>(warp * 32) << 16) + ch * 32, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) out[r * 64 + ch * 32 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<128>(tm);
}

extern "C" int probe_run(const void* x /*[512][64] bf16*/, const void* w /*[128][64] bf16*/, int row0, int sbo,
                         int base_off, int mn_major, float* out /*[128][64]*/) {
  CUtensorMap tx, tw;
  uint64_t dx[2] = {64, 512}, sx[1] = {128};
  uint32_t bx[2] = {64, 256};
  if (make_tmap_bf16(&tx, x, 2, dx, sx, bx)) return -1;
  uint64_t dw[2] = {64, 128};
  uint32_t bw[2] = {64, 128};
  if (make_tmap_bf16(&tw, w, 2, dw, sx, bw)) return -1;
  ProbeCfg c{row0, sbo, base_off, mn_major};
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
  probe_kernel<<<1, 128, 100000>>>(tx, tw, c, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("probe error %s\n", cudaGetErrorString(e)); return -2; }
  return 0;
}
