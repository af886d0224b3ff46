// adam_kernels.cu — fused multi-tensor Adam step that also refreshes the packed bf16 weight copies.
//
// The reference loops call torch.optim.Adam.step() on ~74 parameter tensors right after backward
// (mvp_gan/src/train.py:207,219; training/human_guided_trainer.py:153); the B200 engines then re-derive the bf16
// implicit-GEMM weight matrices (fprop [Cout][taps*Cin], dgrad [Cin][taps*Cout]) from the changed fp32 masters.
// Both packings are permutations of the weight tensor, so one HBM pass per parameter element does the Adam update
// (same arithmetic as torch.optim.Adam, amsgrad off, weight_decay 0) and scatters the rounded value into the two
// packed copies: 16 B read + 12 B written per element, + 8 B of index reads and 4 B of packed writes for conv weights.
#include "tg_common.cuh"
#include "../../include/terragan_b200.h"

namespace tg {

constexpr int kAdamGroup = 24;        // tensors per launch (the table travels in the kernel parameter space)
constexpr int kAdamChunk = 2048;      // elements per block

struct AdamGroup {
  tg_adam_tensor t[kAdamGroup];
  int blk_end[kAdamGroup];            // exclusive prefix sum of blocks per tensor
  int n;
};

__global__ void __launch_bounds__(256)
adam_repack_kernel(const __grid_constant__ AdamGroup grp, float step_size, float beta1, float beta2, float eps,
                   float inv_bc2_sqrt) {
  int ti = 0;
  while (ti < grp.n - 1 && static_cast<int>(blockIdx.x) >= grp.blk_end[ti]) ++ti;
  const tg_adam_tensor& t = grp.t[ti];
  const long first = static_cast<long>(blockIdx.x - (ti ? grp.blk_end[ti - 1] : 0)) * kAdamChunk;
  __nv_bfloat16* wf = reinterpret_cast<__nv_bfloat16*>(t.packed_fprop);
  __nv_bfloat16* wd = reinterpret_cast<__nv_bfloat16*>(t.packed_dgrad);
#pragma unroll
  for (int r = 0; r < kAdamChunk / 256; ++r) {
    const long i = first + r * 256 + threadIdx.x;
    if (i >= t.n) break;
    const float g = t.grad[i];
    float m = t.exp_avg[i], v = t.exp_avg_sq[i];
    m = m + (g - m) * (1.f - beta1);                       // exp_avg.lerp_(grad, 1 - beta1)
    v = v * beta2 + (1.f - beta2) * g * g;                 // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(v) * inv_bc2_sqrt + eps;
    const float pnew = t.param[i] - step_size * (m / denom);
    t.param[i] = pnew;
    t.exp_avg[i] = m;
    t.exp_avg_sq[i] = v;
    if (wf != nullptr) wf[t.dst_fprop[i]] = __float2bfloat16_rn(pnew);
    if (wd != nullptr) wd[t.dst_dgrad[i]] = __float2bfloat16_rn(pnew);
  }
}

}  // namespace tg

extern "C" int tg_adam_repack(const tg_adam_tensor* tensors, int n_tensors, float lr, float beta1, float beta2, float eps,
                              int step, void* stream) {
  using namespace tg;
  TG_REQUIRE(tensors != nullptr && n_tensors >= 0 && step >= 1, "tg_adam_repack: bad arguments");
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), step);
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
  const float step_size = static_cast<float>(lr / bc1);
  const float inv_bc2_sqrt = static_cast<float>(1.0 / sqrt(bc2));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  for (int base = 0; base < n_tensors; base += kAdamGroup) {
    AdamGroup grp;
    grp.n = n_tensors - base < kAdamGroup ? n_tensors - base : kAdamGroup;
    int blocks = 0;
    for (int i = 0; i < grp.n; ++i) {
      const tg_adam_tensor& t = tensors[base + i];
      TG_REQUIRE(t.param && t.grad && t.exp_avg && t.exp_avg_sq && t.n >= 0, "tg_adam_repack: tensor %d has a null pointer",
                 base + i);
      TG_REQUIRE((t.packed_fprop == nullptr) == (t.dst_fprop == nullptr) && (t.packed_dgrad == nullptr) == (t.dst_dgrad == nullptr),
                 "tg_adam_repack: tensor %d: packed copy without its scatter index", base + i);
      grp.t[i] = t;
      blocks += static_cast<int>((t.n + kAdamChunk - 1) / kAdamChunk);
      grp.blk_end[i] = blocks;
    }
    if (blocks == 0) continue;
    adam_repack_kernel<<<blocks, 256, 0, st>>>(grp, step_size, beta1, beta2, eps, inv_bc2_sqrt);
    TG_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}
