// mma_rate.cu — how many clocks does one tcgen05.mma (M=128, K=16, bf16, operands in shared memory) take as a
// function of N and of the number of independent accumulators the issue stream alternates between?
#include <cstdio>
#include "../../terra-gan_b200/csrc/tg_common.cuh"
using namespace tg;

template <int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, int nacc, long long* out, int a_off = 0, int a_sbo = 1024) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + 98304);
  uint32_t* slot = reinterpret_cast<uint32_t*>(mbar + 1);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 98304 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(mbar, 1); fence_barrier_init(); }
  fence_proxy_async();
  if (warp == 1) tmem_alloc<512>(slot);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = *slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, false, false);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 49152;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int acc = it % nacc;
      const uint64_t da = make_smem_desc(a0 + (it % 2) * 24576 + a_off, 16, a_sbo);
      const uint64_t db = make_smem_desc(b0 + (it % 3) * 8192, 16, 1024);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tm + acc * N, da + 2 * k, db + 2 * k, idesc, 1);
    }
    umma_commit(mbar);
    mbar_wait(mbar, 0);
    long long t1 = clock64();
    out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tm);
}

template <int N> static void run(int nacc) {
  long long* d; cudaMalloc(&d, 8);
  const int iters = 2000;
  cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
  rate_kernel<N><<<1, 128, 100000>>>(iters, nacc, d);
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  printf("N=%3d nacc=%d: %.1f clk per MMA (floor %d)  %s\n", N, nacc, (double)h / (iters * 4), 128 * N / 256, cudaGetErrorString(e));
  cudaFree(d);
}

// Two (or more) issuing warps, each with its own accumulator and operand buffers: is the ~78-clock spacing a
// property of the tensor pipe or of one thread's tcgen05.mma issue stream?
template <int N>
__global__ void __launch_bounds__(256, 1) rate_kernel_multi(int iters, int nissue, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + 196608);
  uint32_t* slot = reinterpret_cast<uint32_t*>(mbar + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 196608 / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&mbar[i], 1); fence_barrier_init(); }
  fence_proxy_async();
  if (warp == 7) tmem_alloc<512>(slot);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = *slot;
  long long t0 = clock64();
  if (warp < nissue && lane == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, false, false);
    const uint32_t a0 = smem_u32(smem) + warp * 49152, b0 = a0 + 16384;
    const uint64_t da = make_smem_desc(a0, 16, 1024);
    const uint64_t db = make_smem_desc(b0, 16, 1024);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tm + warp * 128, da + 2 * k, db + 2 * k, idesc, 1);
    }
    umma_commit(&mbar[warp]);
    mbar_wait(&mbar[warp], 0);
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  tc_fence_before(); __syncthreads();
  if (warp == 7) tmem_dealloc<512>(tm);
}
template <int N> static void run_multi(int nissue) {
  long long* d; cudaMalloc(&d, 8);
  const int iters = 2000;
  cudaFuncSetAttribute(rate_kernel_multi<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000);
  rate_kernel_multi<N><<<1, 256, 200000>>>(iters, nissue, d);
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  printf("N=%3d issuing warps=%d: %.1f clk per MMA (all issuers together)  %s\n", N, nissue, (double)h / (iters * 4 * nissue), cudaGetErrorString(e));
  cudaFree(d);
}
template <int N> static void run_off(int a_off, int a_sbo) {
  long long* d; cudaMalloc(&d, 8);
  const int iters = 2000;
  cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
  rate_kernel<N><<<1, 128, 100000>>>(iters, 2, d, a_off, a_sbo);
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  printf("N=%3d A start +%4d B, SBO %4d: %.1f clk per MMA  %s\n", N, a_off, a_sbo, (double)h / (iters * 4), cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  for (int ni : {1, 2, 3, 4}) { run_multi<64>(ni); }
  for (int ni : {1, 2, 4}) { run_multi<128>(ni); }
  run_multi<256>(1); run_multi<256>(2);

  // halo-kernel addressing: A tile starts at an arbitrary 128-byte row, 8-row atoms 1280 B apart
  for (int off : {0, 128, 256, 512, 1280, 1408}) { run_off<64>(off, 1024); run_off<64>(off, 1280); }
  for (int off : {0, 128, 1408}) { run_off<128>(off, 1024); run_off<128>(off, 1280); }
  for (int nacc : {1, 2, 4}) { run<64>(nacc); }
  for (int nacc : {1, 2}) { run<128>(nacc); }
  run<256>(1); run<256>(2);
  return 0;
}
