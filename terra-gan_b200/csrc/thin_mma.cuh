// thin_mma.cuh — parameter blocks of the thin-reduction tensor-core kernels (thin_mma.cu), shared with the
// C-ABI entry points in direct_conv.cu.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace tg {

struct RowGemmParams {
  const float* src;         // fp32 [B][H][W]
  const uint8_t* src_mask;  // optional u8 [B][H][W]; src is taken as 0 where the mask is 0
  int B, H, W, Ho, Wo;
  int S, pad, flip;         // source row of tap kh for output row ho: flip ? ho*S + pad - kh : ho*S - pad + kh
  const float* wgt;         // weight of (output channel n, tap t) = wgt[n * w_sn + perm[t] * w_st]
  int w_sn, w_st;
  int8_t perm[49];
  const float* bias;        // optional [64]
  const uint8_t* code;      // optional u8 [B][Ho][Wo] index into lut (partial-conv ratio)
  const float* lut;
  int act;
  float slope;
  __nv_bfloat16* out;       // [B][Ho][Wo][64] or parity-split [B][4][Ho/2][Wo/2][64]
  int out_split;
  float* stats;             // optional [grid][2][64] per-CTA sum / sum of squares
  unsigned total;           // B*Ho*Wo
  int debug;                // timing experiments only (env TG_THIN_DEBUG): 1 = skip the output store, 2 = skip the source loads
};

// k in {3, 4, 7}. grid_cap > 0 bounds the grid (= number of per-CTA stats rows); returns 0, or -1 if k is unsupported.
int rowgemm_dispatch(int k, const RowGemmParams& p, int grid_cap, int* grid_used, cudaStream_t st);
bool thin_mma_enabled();   // env TG_NO_THIN_MMA=1 selects the CUDA-core kernels (A/B timing only)

}  // namespace tg
