// thin_mma.cuh — parameter blocks of the thin-reduction tensor-core kernels (thin_mma.cu), shared with the
// C-ABI entry points in direct_conv.cu.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "../../include/terragan_b200.h"

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

// Weight gradient of the same thin convolutions:  D[tap][c] = sum over pixels of src_tap(pixel) * Y[pixel][c].
struct TapWgradParams {
  const float* src;         // fp32 [B][H][W]: the single-channel operand (image x, or the gradient g of a 64->1 conv)
  const uint8_t* src_mask;  // optional u8 [B][H][W]
  int B, H, W, Ho, Wo;      // Y pixel grid is [B][Ho][Wo]
  int S, pad, flip;         // same tap geometry as RowGemmParams
  int y_split;              // Y is stored parity-split ([B][4][Ho/2][Wo/2][64])
  unsigned total;           // B*Ho*Wo
  float* partial;           // [grid][2*k*k + 1][64]: rows t = hi part, k*k + t = lo part, 2*k*k = sum of Y (bias gradient)
  float* partial_c;         // optional [grid]: sum over pixels of the source value of tap `center`
  int center;
};
// Launches the accumulation kernel; *grid_used partial rows are written (<= grid_cap).
int tapwgrad_dispatch(int k, const TapWgradParams& p, const void* y, int grid_cap, int* grid_used, cudaStream_t st);
// out_w[c * w_sn + perm[t] * w_st] (+)= sum_r partial[r][t][c] + partial[r][T+t][c];
// bias_mode 1: out_b[c] (+)= sum_r partial[r][2T][c];  bias_mode 2: out_b[0] (+)= sum_r partial_c[r]
int tapwgrad_reduce(int k, const float* partial, const float* partial_c, int rows, float* out_w, int w_sn, int w_st,
                    const int8_t* perm, float* out_b, int bias_mode, int accumulate, cudaStream_t st);

// Tap table of a C -> 1 "gather" convolution (direct_conv.cu): classes of taps selected by output parity.
struct To1Taps {
  int ncls;
  int count[4];
  int begin[4];
  int8_t dh[TG_MAX_TAPS], dw[TG_MAX_TAPS];
};
// 64 -> 1 gather convolution in two steps: per-pixel tap dot products T[t][pixel] = <x[pixel][:], w[t][:]> on the
// tensor cores (x bf16 [pixels][64] in storage order), then the shifted sum + bias / sigmoid-composite epilogue.
// scratch: >= ntaps * B*H*W floats. Returns 0, or -1 when the shape is not covered (caller falls back).
int to1_fwd_mma(const void* x, int x_split, int B, int H, int W, const float* wgt, const To1Taps& taps, int ntaps,
                const float* bias, int Ho, int Wo, int mode, const uint8_t* mask, const float* xin, float* out,
                float* sig_out, float* scratch, size_t scratch_floats, cudaStream_t st);

// true when to1_fwd_mma runs this shape as ONE kernel without scratch (taps inside a 3x3 window, plain layout, same-size output)
bool to1_fused_covers(int x_split, int H, int W, const To1Taps& taps, int ntaps, int Ho, int Wo);

// k in {3, 4, 7}. grid_cap > 0 bounds the grid (= number of per-CTA stats rows); returns 0, or -1 if k is unsupported.
int rowgemm_dispatch(int k, const RowGemmParams& p, int grid_cap, int* grid_used, cudaStream_t st);
bool thin_mma_enabled();   // env TG_NO_THIN_MMA=1 selects the CUDA-core kernels (A/B timing only)

int to1_tapsum_launch(const float* T, long total_in, int x_split, int B, int H, int W, const To1Taps& taps,
                      const float* bias, int Ho, int Wo, int mode, const uint8_t* mask, const float* xin, float* out,
                      float* sig_out, cudaStream_t st);

}  // namespace tg
