// mma_rate_pair.cu — clocks per tcgen05.mma.cta_group::2 (M = 256 over a CTA pair, K = 16, bf16, operands in shared
// memory) as a function of N, next to the single-CTA M = 128 instruction (mma_rate.cu: 55 / 64 / 128 clk at N = 64 / 128 / 256).
// Per SM a pair MMA reads its own 128 A rows (4 KB) and HALF of the B rows from shared memory.
#include <cstdio>
#include "../../terra-gan_b200/csrc/tg_common.cuh"
using namespace tg;

template <int N>
__global__ void __launch_bounds__(128, 1) pair_rate_kernel(int iters, int nacc, long long* out, int a_off, int a_sbo) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + 98304);
  uint32_t* slot = reinterpret_cast<uint32_t*>(mbar + 1);
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < 98304 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(mbar, 1); fence_barrier_init(); }
  fence_proxy_async();
  if (warp == 1) tmem_alloc_2cta<512>(slot);
  tc_fence_before(); cluster_sync_all(); tc_fence_after();
  const uint32_t tm = *slot;
  if (threadIdx.x == 0 && rank == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(256, N, false, false);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 49152;
    // descriptor words precomputed (per-MMA 64-bit descriptor arithmetic makes the issuing thread the bound: 77 clk)
    const uint64_t da = make_smem_desc(a0 + a_off, 16, a_sbo), db = make_smem_desc(b0, 16, 1024);
    const uint32_t a_lo = static_cast<uint32_t>(da), a_hi = static_cast<uint32_t>(da >> 32);
    const uint32_t b_lo = static_cast<uint32_t>(db), b_hi = static_cast<uint32_t>(db >> 32);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tm + (nacc > 1 ? (it & 1) * N : 0);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_lh_2cta(d, a_lo + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc, 1);
    }
    umma_commit_2cta(mbar);
    mbar_wait(mbar, 0);
    long long t1 = clock64();
    out[0] = t1 - t0;
  }
  if (threadIdx.x == 0 && rank == 1) mbar_wait(mbar, 0);     // the multicast commit also lands here
  tc_fence_before(); cluster_sync_all();
  if (warp == 1) tmem_dealloc_2cta<512>(tm);
}

template <int N> static void run(int nacc, int a_off = 0, int a_sbo = 1024) {
  long long* d; cudaMalloc(&d, 8);
  const int iters = 2000;
  cudaFuncSetAttribute(pair_rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 100000;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, pair_rate_kernel<N>, iters, nacc, d, a_off, a_sbo);
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaGetLastError();
  printf("pair M=256 N=%3d nacc=%d A+%d sbo %d: %.1f clk per MMA (floor %d; single-CTA M=128 needs 2 MMAs for the same work)  %s\n",
         N, nacc, a_off, a_sbo, (double)h / (iters * 4), 128 * N / 256, cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  for (int nacc : {1, 2}) { run<64>(nacc); run<128>(nacc); run<256>(nacc); }
  run<64>(2, 1408, 1280); run<128>(2, 1408, 1280);      // halo-kernel addressing
  return 0;
}
