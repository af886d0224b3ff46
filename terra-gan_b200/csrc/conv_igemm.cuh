// conv_igemm.cuh — kernel-parameter blocks of the tcgen05 implicit-GEMM convolution kernels.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace tg {

constexpr int kMaxTaps = 64;   // 7x7 = 49 is the largest window on the path (enc1; direct kernel)
constexpr int kMaxSub = 4;     // stride-2 dgrad: one sub-problem per input parity phase
constexpr int kMaxChan = 1024; // widest channel count on the path (dec5..7 inputs)

struct ConvSubK {
  int tap_begin;   // first entry of the tap table used by this sub-problem
  int tap_count;   // number of taps
  int k_off;       // column offset of this sub-problem's slab in the packed weight matrix
  int out_plane;   // output parity plane written by this sub-problem
};

// fprop / dgrad: out[pixel][n] = act(((sum_k A[pixel (+) tap][k] * W[n][k]) + bias[n]) * lut[code[pixel]]
//                                      * scale[n] + shift[n])
struct ConvKParams {
  int Bt, Ht, Wt;                 // pixel box of one M tile; Bt*Ht*Wt == 128
  int tiles_w, tiles_h, tiles_b;  // tile counts along each axis of the output grid
  int n_tiles;                    // Cout / BLOCK_N
  int num_sub;
  ConvSubK sub[kMaxSub];
  int8_t tap_plane[kMaxTaps];
  int8_t tap_dh[kMaxTaps];
  int8_t tap_dw[kMaxTaps];
  int cin_blocks;                 // input channels / 64
  __nv_bfloat16* out;             // [B][Po][Ho][Wo][Cout]
  int B, Ho, Wo, Po, Cout;
  const uint8_t* code;            // per output pixel, same pixel order as out; may be null
  const float* bias;              // [Cout] or null
  const float* scale;             // [Cout] or null (eval-mode folded BatchNorm)
  const float* shift;             // [Cout] or null
  float lut[kMaxTaps];            // code -> row scale (mask-ratio LUT or {0,1})
  int act;                        // 0 none, 1 ReLU, 2 LeakyReLU(slope)
  float slope;
  float* stats;                   // [gridDim.x][2][Cout] per-CTA (sum, sum of squares) or null
  const __nv_bfloat16* gate;      // same shape as out, or null: out *= (gate > 0 ? 1 : gate_slope)
  float gate_slope;
  int debug;                      // perf experiments only: 1 no stores, 2 no epilogue work, 4 no MMA
  const float* addend;            // fp32 path only: [out shape] added to the accumulator before the epilogue, or null
  __nv_bfloat16* pool_out;        // halo kernels only: [B][Ho/2][Wo/2][Cout] = 2x2 max-pool of the (activated) output, or null
  int skip_out;                   // with pool_out: do not store the full-resolution output at all
};

#ifdef __CUDACC__
// (tw, th, tb) of the tiles first, first + stride, ... without per-tile integer divisions (they sat on the
// epilogue warps' dependent chain once per tile)
struct TileWalk {
  int tw, th, tb, dw, dh, db, nw, nh;
  __device__ __forceinline__ TileWalk(int first, int stride, int tiles_w, int tiles_h) {
    nw = tiles_w;
    nh = tiles_h;
    tw = first % nw;
    int r = first / nw;
    th = r % nh;
    tb = r / nh;
    dw = stride % nw;
    r = stride / nw;
    dh = r % nh;
    db = r / nh;
  }
  __device__ __forceinline__ void next() {
    tw += dw;
    int c = tw >= nw ? 1 : 0;
    tw -= c ? nw : 0;
    th += dh + c;
    c = th >= nh ? 1 : 0;
    th -= c ? nh : 0;
    tb += db + c;
  }
};
#endif

struct WgradBlk {                 // one 64-row block of dW^T: (tap, 64-channel block of the input)
  int8_t plane, dh, dw, pad;
  int cb;                         // input channel block
  int row;                        // first row of this block in the packed [T*Cin][Cout] gradient
};

// wgrad: dWt[(tap, ci)][co] = sum_pixels X[pixel (+) tap][ci] * G[pixel][co]
struct WgradKParams {
  int Bt, Ht, Wt;                 // pixel box of one K block; Bt*Ht*Wt == 64
  int tiles_w, tiles_h, tiles_b;
  int n_tiles;                    // Cout / BLOCK_N
  int m_tiles;                    // ceil(num_blk / 2)
  int num_blk;
  int splits;                     // split-K factor over the pixel boxes
  int Cout;
  int rows;                       // T * Cin
  float* partial;                 // [splits][rows][Cout] fp32
  const WgradBlk* blks;           // [num_blk] in global memory
};

}  // namespace tg
