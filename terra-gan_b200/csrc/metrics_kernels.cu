// metrics_kernels.cu — the logging-interval image-quality metrics of the reference as ONE fused reduction
// (SURVEY.md §8f rank 2). Replaces, per call of ExperimentTracker.log_training_batch / the train loops'
// boundary logging (utils/experiment_tracking.py:196-231,678-695; mvp_gan/src/utils/metrics.py:12-46;
// mvp_gan/src/evaluation/metrics.py:79-133; called from train.py:229-266 every log_interval batches):
//   PSNR   = 20 log10(1 / sqrt(mse))                      mse = mean (p - t)^2
//   SSIM   = mean of the 11x11 avg_pool2d (stride 1, zero padding 5, divisor 121) SSIM map, C1 = 1e-4, C2 = 9e-4
//   L1, L2 = mean |p - t|, sqrt(mse)
//   calculate_boundary_quality: bd = clamp(maxpool3(m) - (1 - maxpool3(1 - m)), 0, 1);
//       boundary_mse = mean ((p - t) bd)^2 over ALL pixels, boundary_psnr = 10 log10(1 / (boundary_mse + 1e-6)),
//       boundary_gradient_diff = | (mean|dh p| + mean|dw p|) - (mean|dh t| + mean|dw t|) |, all three 0 if sum(bd) < 1e-6
// which the reference evaluates with ~40 ATen kernels and 8 .item() host syncs. Here: one pass over (pred, target,
// mask) — 12 B per pixel — with the 11x11 box sums done separably in shared memory, per-block partials, and a
// finalize kernel that leaves the nine numbers in a small device array (no host synchronisation).
#include "tg_common.cuh"
#include "../../include/terragan_b200.h"

namespace tg {

constexpr int kMT = 32;            // output tile (kMT x kMT pixels per block)
constexpr int kMR = 5;             // SSIM window radius (11 x 11)
constexpr int kMH = kMT + 2 * kMR; // halo tile edge
constexpr int kMetricSums = 10;    // per-block partial sums (see below)

// partial[block][kMetricSums]: 0 sum|d|, 1 sum d^2, 2 sum ssim, 3 sum (d*bd)^2, 4 sum bd, 5 sum|dh p|, 6 sum|dw p|,
//                              7 sum|dh t|, 8 sum|dw t|, 9 unused
__global__ void __launch_bounds__(256)
quality_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ target, const float* __restrict__ mask,
                       int B, int H, int W, double* __restrict__ partial) {
  __shared__ float s_p[kMH][kMH + 1];
  __shared__ float s_t[kMH][kMH + 1];
  __shared__ float s_h[5][kMH][kMT + 1];     // horizontal 11-sums of p, t, p^2, t^2, p t for every halo row
  __shared__ float s_m[kMT + 2][kMT + 3];    // mask tile with a 1-pixel halo (3x3 dilate / erode)
  __shared__ double s_red[8][kMetricSums];
  const int tid = threadIdx.x;
  const int tiles_w = (W + kMT - 1) / kMT, tiles_h = (H + kMT - 1) / kMT;
  const long total_tiles = static_cast<long>(B) * tiles_h * tiles_w;
  double acc[kMetricSums];
#pragma unroll
  for (int k = 0; k < kMetricSums; ++k) acc[k] = 0.0;

  for (long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int tw = static_cast<int>(tile % tiles_w);
    const int th = static_cast<int>((tile / tiles_w) % tiles_h);
    const int b = static_cast<int>(tile / (static_cast<long>(tiles_w) * tiles_h));
    const int h0 = th * kMT, w0 = tw * kMT;
    const float* pb = pred + static_cast<long>(b) * H * W;
    const float* tb = target + static_cast<long>(b) * H * W;
    const float* mb = mask ? mask + static_cast<long>(b) * H * W : nullptr;
    __syncthreads();
    for (int i = tid; i < kMH * kMH; i += 256) {
      const int r = i / kMH, c = i % kMH;
      const int h = h0 - kMR + r, w = w0 - kMR + c;
      const bool in = h >= 0 && h < H && w >= 0 && w < W;
      s_p[r][c] = in ? pb[static_cast<long>(h) * W + w] : 0.f;      // avg_pool2d zero padding
      s_t[r][c] = in ? tb[static_cast<long>(h) * W + w] : 0.f;
    }
    if (mb != nullptr) {
      for (int i = tid; i < (kMT + 2) * (kMT + 2); i += 256) {
        const int r = i / (kMT + 2), c = i % (kMT + 2);
        const int h = h0 - 1 + r, w = w0 - 1 + c;
        // max_pool2d pads with -inf: out-of-image neighbours never win. NaN marks "absent".
        s_m[r][c] = (h >= 0 && h < H && w >= 0 && w < W) ? mb[static_cast<long>(h) * W + w] : __int_as_float(0x7fc00000);
      }
    }
    __syncthreads();
    // horizontal pass: for every halo row, the 11-wide sums at the kMT output columns. One thread = 4 adjacent outputs
    // of one row: 14 loads and products feed 4 windows (the first summed in full, the next three by sliding:
    // + entering - leaving), instead of 11 loads per output
    for (int i = tid; i < kMH * (kMT / 4); i += 256) {
      const int r = i / (kMT / 4), c0 = (i % (kMT / 4)) * 4;
      float pv[14], tv[14];
#pragma unroll
      for (int k = 0; k < 14; ++k) { pv[k] = s_p[r][c0 + k]; tv[k] = s_t[r][c0 + k]; }
      float sp = 0.f, st = 0.f, spp = 0.f, stt = 0.f, spt = 0.f;
#pragma unroll
      for (int k = 0; k < 11; ++k) { sp += pv[k]; st += tv[k]; spp += pv[k] * pv[k]; stt += tv[k] * tv[k]; spt += pv[k] * tv[k]; }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j > 0) {
          const float pa = pv[j + 10], ta = tv[j + 10], pb = pv[j - 1], tb = tv[j - 1];
          sp += pa - pb; st += ta - tb; spp += pa * pa - pb * pb; stt += ta * ta - tb * tb; spt += pa * ta - pb * tb;
        }
        s_h[0][r][c0 + j] = sp; s_h[1][r][c0 + j] = st; s_h[2][r][c0 + j] = spp; s_h[3][r][c0 + j] = stt; s_h[4][r][c0 + j] = spt;
      }
    }
    __syncthreads();
    // vertical pass + every per-pixel term. One thread = 4 vertically adjacent pixels of one column (256 threads cover
    // the 32 x 32 tile): the column sums slide like the row sums above
    {
      const int c = tid & (kMT - 1), r0 = (tid / kMT) * 4;
      const int w = w0 + c;
      float sums[5];
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < 2 * kMR + 1; ++k) a += s_h[q][r0 + k][c];
        sums[q] = a;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = r0 + j;
        if (j > 0) {
#pragma unroll
          for (int q = 0; q < 5; ++q) sums[q] += s_h[q][r + 2 * kMR][c] - s_h[q][r - 1][c];
        }
        const int h = h0 + r;
        if (h >= H || w >= W) continue;
        const float inv = 1.f / 121.f;
        const float mu1 = sums[0] * inv, mu2 = sums[1] * inv;
        const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
        const float s1 = sums[2] * inv - mu1_sq, s2 = sums[3] * inv - mu2_sq, s12 = sums[4] * inv - mu12;
        const float C1 = 0.0001f, C2 = 0.0009f;
        const float ssim = ((2.f * mu12 + C1) * (2.f * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2));
        const float p = s_p[r + kMR][c + kMR], t = s_t[r + kMR][c + kMR];
        const float d = p - t;
        acc[0] += fabsf(d);
        acc[1] += static_cast<double>(d) * d;
        acc[2] += ssim;
        if (mb != nullptr) {
          float mx = -1e30f, mn = 1e30f;       // dilate = max over the 3x3 window, erode = min (= 1 - max(1 - m))
#pragma unroll
          for (int dr = 0; dr < 3; ++dr)
#pragma unroll
            for (int dc = 0; dc < 3; ++dc) {
              const float v = s_m[r + dr][c + dc];
              if (v == v) { mx = fmaxf(mx, v); mn = fminf(mn, v); }
            }
          float bd = mx - mn;                   // dilated - eroded
          bd = fminf(fmaxf(bd, 0.f), 1.f);
          const float e = d * bd;
          acc[3] += static_cast<double>(e) * e;
          acc[4] += bd;
        }
        if (h + 1 < H) {
          acc[5] += fabsf(s_p[r + kMR + 1][c + kMR] - p);
          acc[7] += fabsf(s_t[r + kMR + 1][c + kMR] - t);
        }
        if (w + 1 < W) {
          acc[6] += fabsf(s_p[r + kMR][c + kMR + 1] - p);
          acc[8] += fabsf(s_t[r + kMR][c + kMR + 1] - t);
        }
      }
    }
  }
  // block reduction (fixed order: deterministic)
#pragma unroll
  for (int k = 0; k < kMetricSums; ++k) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    acc[k] = v;
  }
  if ((tid & 31) == 0) {
#pragma unroll
    for (int k = 0; k < kMetricSums; ++k) s_red[tid >> 5][k] = acc[k];
  }
  __syncthreads();
  if (tid < kMetricSums) {
    double v = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) v += s_red[q][tid];
    partial[static_cast<long>(blockIdx.x) * kMetricSums + tid] = v;
  }
}

// out[0..8] = psnr, ssim, l1_distance, l2_distance, mse, boundary_mse, boundary_psnr, boundary_gradient_diff,
//             boundary pixel count
__global__ void quality_metrics_finalize_kernel(const double* __restrict__ partial, int rows, int B, int H, int W,
                                                int has_mask, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double s[kMetricSums];
  for (int k = 0; k < kMetricSums; ++k) s[k] = 0.0;
  for (int r = 0; r < rows; ++r)
    for (int k = 0; k < kMetricSums; ++k) s[k] += partial[static_cast<long>(r) * kMetricSums + k];
  const double n = static_cast<double>(B) * H * W;
  const float mse = static_cast<float>(s[1] / n);
  out[4] = mse;
  out[0] = mse == 0.f ? __int_as_float(0x7f800000) : 20.f * log10f(1.f / sqrtf(mse));
  out[1] = static_cast<float>(s[2] / n);
  out[2] = static_cast<float>(s[0] / n);
  out[3] = sqrtf(mse);
  out[5] = out[6] = out[7] = 0.f;
  out[8] = static_cast<float>(s[4]);
  if (has_mask && s[4] >= 1e-6) {
    const float bmse = static_cast<float>(s[3] / n);
    out[5] = bmse;
    out[6] = 10.f * log10f(1.f / (bmse + 1e-6f));
    const double nh = static_cast<double>(B) * (H - 1) * W, nw = static_cast<double>(B) * H * (W - 1);
    const float pd = static_cast<float>(s[5] / nh) + static_cast<float>(s[6] / nw);
    const float td = static_cast<float>(s[7] / nh) + static_cast<float>(s[8] / nw);
    out[7] = fabsf(pd - td);
  }
}

}  // namespace tg

extern "C" int tg_quality_metrics_rows(void) { return tg::num_sms() * 4; }

extern "C" int tg_quality_metrics(const float* pred, const float* target, const float* mask, int B, int H, int W,
                                  double* partial, int rows_cap, float* out, void* stream) {
  using namespace tg;
  TG_REQUIRE(pred && target && partial && out && B > 0 && H > 1 && W > 1, "tg_quality_metrics: bad arguments");
  const long tiles = static_cast<long>(B) * ((H + kMT - 1) / kMT) * ((W + kMT - 1) / kMT);
  long grid = tiles < 4L * num_sms() ? tiles : 4L * num_sms();
  if (grid > rows_cap) grid = rows_cap;
  TG_REQUIRE(grid >= 1, "tg_quality_metrics: rows_cap must be >= 1");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  quality_metrics_kernel<<<static_cast<int>(grid), 256, 0, st>>>(pred, target, mask, B, H, W, partial);
  TG_CHECK_CUDA(cudaGetLastError());
  quality_metrics_finalize_kernel<<<1, 32, 0, st>>>(partial, static_cast<int>(grid), B, H, W, mask != nullptr, out);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}
