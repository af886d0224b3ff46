// resample.cu — decoder input assembly and pooling as fused, vectorised bandwidth kernels.
//
// tg_upsample_concat replaces, for one decoder stage of PConvUNet.decode_step
// (mvp_gan/src/models/generator.py:66-76 and :50-55 for dec1):
//     interpolate(up, x2, bilinear, align_corners=False)  ++  skip   (channel concat, up first)
//     followed by the `input * mask` of the consuming PConv2d (pconv.py:27) with the merged mask
//     max(nearest_up2(up_mask), skip_mask), which tg_mask_merge_up precomputed.
// One pass: reads the low-res feature and the skip feature, writes the masked merged tensor the
// tensor-core conv consumes (5 ATen kernels in the reference). tg_upsample_concat_bwd is the
// transpose of the bilinear part (the skip part of the gradient is consumed in place).
//
// tg_maxpool2 / tg_maxpool2_bwd replace nn.MaxPool2d(2,2) of VGG16 features[4], [9]
// (mvp_gan/src/utils/losses.py:31-32) with the ReLU gate of the preceding layer folded into bwd.
#include "tg_common.cuh"
#include "../../include/terragan_b200.h"

namespace tg {

__device__ __forceinline__ void rs_load8(const __nv_bfloat16* p, float (&f)[8]) { vload8(p, f); }
__device__ __forceinline__ void rs_store8(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 raw;
  raw.x = pack_bf16x2(f[0], f[1]);
  raw.y = pack_bf16x2(f[2], f[3]);
  raw.z = pack_bf16x2(f[4], f[5]);
  raw.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = raw;
}

__device__ __forceinline__ void rs_load8(const float* p, float (&f)[8]) { vload8(p, f); }
__device__ __forceinline__ void rs_store8(float* p, const float (&f)[8]) { vstore8(p, f); }
// raw 8-element vector of the storage type (held in registers between load and use)
template <typename T> struct Raw8;
template <> struct Raw8<__nv_bfloat16> { uint4 v; };
template <> struct Raw8<float> { float4 a, b; };
__device__ __forceinline__ void raw_zero(Raw8<__nv_bfloat16>& r) { r.v = make_uint4(0u, 0u, 0u, 0u); }
__device__ __forceinline__ void raw_zero(Raw8<float>& r) { r.a = r.b = make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void raw_load(Raw8<__nv_bfloat16>& r, const __nv_bfloat16* p) { r.v = *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void raw_load(Raw8<float>& r, const float* p) {
  r.a = *reinterpret_cast<const float4*>(p);
  r.b = *(reinterpret_cast<const float4*>(p) + 1);
}
__device__ __forceinline__ void raw_store(__nv_bfloat16* p, const Raw8<__nv_bfloat16>& r) { *reinterpret_cast<uint4*>(p) = r.v; }
__device__ __forceinline__ void raw_store(float* p, const Raw8<float>& r) {
  *reinterpret_cast<float4*>(p) = r.a;
  *(reinterpret_cast<float4*>(p) + 1) = r.b;
}
__device__ __forceinline__ void raw_unpack(const Raw8<__nv_bfloat16>& r, float (&f)[8]) {
  unpack_bf16x2(r.v.x, f[0], f[1]);
  unpack_bf16x2(r.v.y, f[2], f[3]);
  unpack_bf16x2(r.v.z, f[4], f[5]);
  unpack_bf16x2(r.v.w, f[6], f[7]);
}
__device__ __forceinline__ void raw_unpack(const Raw8<float>& r, float (&f)[8]) {
  f[0] = r.a.x; f[1] = r.a.y; f[2] = r.a.z; f[3] = r.a.w;
  f[4] = r.b.x; f[5] = r.b.y; f[6] = r.b.z; f[7] = r.b.w;
}

// PyTorch's upsample_bilinear2d source index for scale 2, align_corners=False:
// src = max(0, (o + 0.5) / 2 - 0.5); i0 = floor(src); i1 = min(i0 + 1, n - 1); l1 = src - i0.
__device__ __forceinline__ void bilinear_src(int o, int n, int& i0, int& i1, float& l0, float& l1) {
  float src = (o + 0.5f) * 0.5f - 0.5f;
  src = src < 0.f ? 0.f : src;
  i0 = static_cast<int>(src);
  i1 = i0 + (i0 < n - 1 ? 1 : 0);
  l1 = src - i0;
  l0 = 1.f - l1;
}

// out[b][oh][ow][0:Cu] = bilinear_up2(up)[...] * mm ; out[...][Cu:Cu+Cs] = skip * mm
// One thread = one 8-channel vector of a 2x2 output block: the four outputs of source pixel (i, j)
// share its 3x3 neighbourhood (9 loads for 4 outputs instead of 16), index math is 32-bit and
// amortised over the block. Exact PyTorch weights: 0.75/0.25 with edge clamping.
// Two passes in one launch: first every (block, up-sampled channel vector) item, then every (block, skip channel vector)
// item. With one mixed item space (channel vector fastest) a warp held both kinds whenever (Cu + Cs) / 8 is not a
// multiple of 32 (dec2: 24) and executed the interpolation AND the copy path for every item; the skip part is a
// masked raw 16-byte copy (no unpack / repack).
// item = ((b * h + ii) * w + jj) * cv + channel vector; a thread's items are a constant stride apart, so the four
// digits are walked with carries instead of three integer divisions per item (they were ~60 of the ~130 instructions
// of a skip-part item and ~15 % of an up-sampled one; the kernel is issue-bound)
struct BlockWalk {
  int cvi, jj, ii;
  unsigned b;
  int d_cv, d_j, d_i;
  unsigned d_b;
  int cv, w, h;
  __device__ __forceinline__ BlockWalk(unsigned first, unsigned step, unsigned cv_, int h_, int w_) : cv(static_cast<int>(cv_)), w(w_), h(h_) {
    const unsigned hw = static_cast<unsigned>(h_) * w_;
    unsigned blk = first / cv_;
    cvi = static_cast<int>(first - blk * cv_);
    b = blk / hw;
    unsigned rem = blk - b * hw;
    ii = static_cast<int>(rem / w_);
    jj = static_cast<int>(rem - ii * w_);
    blk = step / cv_;
    d_cv = static_cast<int>(step - blk * cv_);
    d_b = blk / hw;
    rem = blk - d_b * hw;
    d_i = static_cast<int>(rem / w_);
    d_j = static_cast<int>(rem - d_i * w_);
  }
  __device__ __forceinline__ void advance() {
    cvi += d_cv;
    int cy = cvi >= cv ? 1 : 0;
    cvi -= cy ? cv : 0;
    jj += d_j + cy;
    cy = jj >= w ? 1 : 0;
    jj -= cy ? w : 0;
    ii += d_i + cy;
    cy = ii >= h ? 1 : 0;
    ii -= cy ? h : 0;
    b += d_b + cy;
  }
};
template <typename T>
__global__ void __launch_bounds__(256)
upsample_concat_kernel(const T* __restrict__ up, int B, int h, int w, int Cu,
                       const T* __restrict__ skip, int Cs, const uint8_t* __restrict__ mm,
                       T* __restrict__ out) {
  const int H = 2 * h, W = 2 * w, C = Cu + Cs;
  const unsigned hw = static_cast<unsigned>(h) * w;
  const unsigned i_first = blockIdx.x * blockDim.x + threadIdx.x, i_step = gridDim.x * blockDim.x;
  // ---- pass 1: the bilinearly up-sampled channels ----
  {
    const unsigned cv = Cu >> 3;
    const unsigned total = static_cast<unsigned>(B) * hw * cv;
    BlockWalk wk(i_first, i_step, cv, h, w);
    for (unsigned i = i_first; i < total; i += i_step, wk.advance()) {
      const int c = wk.cvi << 3, ii = wk.ii, jj = wk.jj;
      const unsigned b = wk.b;
      const size_t obase = (static_cast<size_t>(b) * H + 2 * ii) * W + 2 * jj;   // pixel (2i, 2j)
      const size_t opix[4] = {obase, obase + 1, obase + W, obase + W + 1};
      bool on[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) on[q] = (mm == nullptr) || (mm[opix[q]] != 0);
      float o[4][8];
      if (on[0] || on[1] || on[2] || on[3]) {
        const int im = ii > 0 ? ii - 1 : 0, ip = ii < h - 1 ? ii + 1 : h - 1;
        const int jm = jj > 0 ? jj - 1 : 0, jp = jj < w - 1 ? jj + 1 : w - 1;
        const T* base = up + static_cast<size_t>(b) * hw * Cu + c;
        float t[3][3][8];
        const int rr[3] = {im, ii, ip}, cc[3] = {jm, jj, jp};
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int q = 0; q < 3; ++q) rs_load8(base + (static_cast<size_t>(rr[r]) * w + cc[q]) * Cu, t[r][q]);
        // output row 2i uses rows (i-1, i) with (0.25, 0.75) [row 0: weight 1 on i], row 2i+1 uses (i, i+1)
        // with (0.75, 0.25) [last row: weight 1 on i]; same along columns. Clamped neighbours make the edge
        // cases fall out of the same formula: 0.25*x[i] + 0.75*x[i] = x[i].
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          // PyTorch evaluates lh0*(lw0*a00 + lw1*a01) + lh1*(lw0*a10 + lw1*a11) with (h0,h1) = (i-1,i) or (i,i+1)
          const float r0a = 0.25f * t[0][0][e] + 0.75f * t[0][1][e], r0b = 0.75f * t[0][1][e] + 0.25f * t[0][2][e];
          const float r1a = 0.25f * t[1][0][e] + 0.75f * t[1][1][e], r1b = 0.75f * t[1][1][e] + 0.25f * t[1][2][e];
          const float r2a = 0.25f * t[2][0][e] + 0.75f * t[2][1][e], r2b = 0.75f * t[2][1][e] + 0.25f * t[2][2][e];
          o[0][e] = 0.25f * r0a + 0.75f * r1a;
          o[1][e] = 0.25f * r0b + 0.75f * r1b;
          o[2][e] = 0.75f * r1a + 0.25f * r2a;
          o[3][e] = 0.75f * r1b + 0.25f * r2b;
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (!on[q]) {
#pragma unroll
          for (int e = 0; e < 8; ++e) o[q][e] = 0.f;
        }
        rs_store8(out + opix[q] * C + c, o[q]);
      }
    }
  }
  // ---- pass 2: the skip channels (masked copy) ----
  if (Cs > 0) {
    const unsigned cv = Cs >> 3;
    const unsigned total = static_cast<unsigned>(B) * hw * cv;
    BlockWalk wk(i_first, i_step, cv, h, w);
    for (unsigned i = i_first; i < total; i += i_step, wk.advance()) {
      const int c = wk.cvi << 3;
      const size_t obase = (static_cast<size_t>(wk.b) * H + 2 * wk.ii) * W + 2 * wk.jj;
      const size_t opix[4] = {obase, obase + 1, obase + W, obase + W + 1};
      Raw8<T> r[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        raw_zero(r[q]);
        if ((mm == nullptr) || (mm[opix[q]] != 0)) raw_load(r[q], skip + opix[q] * Cs + c);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) raw_store(out + opix[q] * C + Cu + c, r[q]);
    }
  }
}

// d_up[b][i][j][c] = sum over the (<= 4x4) output pixels whose bilinear footprint touches (i, j)
// of weight * d_merged[b][oh][ow][c]   (d_merged already carries the merged-mask factor)
// One thread = one 8-channel vector of a column strip of kUbSeg low-resolution rows. The transpose of the
// bilinear x2 stencil is separable: hr(oh) = sum_q ww[q] * g[oh][2j-1+q] is formed once per high-resolution row and
// shared by the two low-resolution rows that use it (a sliding window of 4 rows in registers), which halves the
// loads and bf16 unpacks per output of the 4x4-window gather this replaces (measured instruction-bound).
constexpr int kUbSeg = 16;
template <typename T>
__global__ void __launch_bounds__(256)
upsample_concat_bwd_kernel(const T* __restrict__ dm, int B, int h, int w, int Cu, int Ctot,
                           T* __restrict__ dup) {
  const int H = 2 * h, W = 2 * w;
  const unsigned cv = Cu >> 3;
  const unsigned segs = (static_cast<unsigned>(h) + kUbSeg - 1) / kUbSeg;
  const unsigned total = static_cast<unsigned>(B) * segs * w * cv;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = static_cast<int>(i % cv) << 3;
    unsigned rest = i / cv;
    const int jj = static_cast<int>(rest % w);
    rest /= w;
    const int seg = static_cast<int>(rest % segs);
    const int b = static_cast<int>(rest / segs);
    const T* base = dm + static_cast<size_t>(b) * H * W * Ctot + c;
    // source column j feeds output columns 2j-1, 2j, 2j+1, 2j+2 with weights 0.25, 0.75, 0.75, 0.25; at the borders
    // the clamped neighbour folds its weight onto the edge pixel (same along rows)
    float ww[4];
    ww[0] = jj > 0 ? 0.25f : 0.f;
    ww[1] = jj > 0 ? 0.75f : 1.f;
    ww[2] = jj < w - 1 ? 0.75f : 1.f;
    ww[3] = jj < w - 1 ? 0.25f : 0.f;
    // rows are fetched one loop iteration ahead of their use (the loop is otherwise one load-latency per output row)
    auto load_row = [&](int oh, Raw8<T> (&raw)[4]) {
#pragma unroll
      for (int q = 0; q < 4; ++q) raw_zero(raw[q]);
      if (oh < 0 || oh >= H) return;
      const T* rowp = base + static_cast<size_t>(oh) * W * Ctot;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int ow = 2 * jj - 1 + q;
        if (ww[q] != 0.f) raw_load(raw[q], rowp + static_cast<size_t>(ow) * Ctot);
      }
    };
    auto reduce_row = [&](const Raw8<T> (&raw)[4], float (&out)[8]) {
#pragma unroll
      for (int e = 0; e < 8; ++e) out[e] = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float t[8];
        raw_unpack(raw[q], t);
#pragma unroll
        for (int e = 0; e < 8; ++e) out[e] += ww[q] * t[e];
      }
    };
    const int i0 = seg * kUbSeg, i1 = min(h, i0 + kUbSeg);
    float r0[8], r1[8], r2[8], r3[8];
    Raw8<T> ra[4], rb[4];
    load_row(2 * i0 - 1, ra);
    load_row(2 * i0, rb);
    reduce_row(ra, r0);
    reduce_row(rb, r1);
    load_row(2 * i0 + 1, ra);
    load_row(2 * i0 + 2, rb);
    for (int ii = i0; ii < i1; ++ii) {
      Raw8<T> na[4], nb[4];
      load_row(ii + 1 < i1 ? 2 * ii + 3 : -1, na);
      load_row(ii + 1 < i1 ? 2 * ii + 4 : -1, nb);
      reduce_row(ra, r2);
      reduce_row(rb, r3);
      const float a0 = ii > 0 ? 0.25f : 0.f, a1 = ii > 0 ? 0.75f : 1.f;
      const float a2 = ii < h - 1 ? 0.75f : 1.f, a3 = ii < h - 1 ? 0.25f : 0.f;
      float acc[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = a0 * r0[e] + a1 * r1[e] + a2 * r2[e] + a3 * r3[e];
      rs_store8(dup + ((static_cast<size_t>(b) * h + ii) * w + jj) * Cu + c, acc);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        r0[e] = r2[e];
        r1[e] = r3[e];
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        ra[q] = na[q];
        rb[q] = nb[q];
      }
    }
  }
}

template <typename T>
__global__ void maxpool2_kernel(const T* __restrict__ x, int B, int H, int W, int C,
                                T* __restrict__ y) {
  const int h = H >> 1, w = W >> 1, cv = C >> 3;
  const long total = static_cast<long>(B) * h * w * cv;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long p = i / cv;
    const int c = static_cast<int>(i % cv) << 3;
    const int j = static_cast<int>(p % w);
    const int ii = static_cast<int>((p / w) % h);
    const int b = static_cast<int>(p / (static_cast<long>(w) * h));
    const T* base = x + ((static_cast<long>(b) * H + 2 * ii) * W + 2 * j) * C + c;
    float a[8], t[8];
    rs_load8(base, a);
    rs_load8(base + C, t);
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = fmaxf(a[q], t[q]);
    rs_load8(base + static_cast<long>(W) * C, t);
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = fmaxf(a[q], t[q]);
    rs_load8(base + static_cast<long>(W) * C + C, t);
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = fmaxf(a[q], t[q]);
    rs_store8(y + p * C + c, a);
  }
}

// gx[b][H][W][C]: the pooled gradient goes to the first maximum of each 2x2 window (scan order, as
// ATen's max_pool2d_with_indices does), times the ReLU gate [x > 0] of the layer that produced x.
template <typename T>
__global__ void maxpool2_bwd_kernel(const T* __restrict__ x, const T* __restrict__ gy,
                                    int B, int H, int W, int C, int relu_gate,
                                    T* __restrict__ gx) {
  const int h = H >> 1, w = W >> 1, cv = C >> 3;
  const long total = static_cast<long>(B) * h * w * cv;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long p = i / cv;
    const int c = static_cast<int>(i % cv) << 3;
    const int j = static_cast<int>(p % w);
    const int ii = static_cast<int>((p / w) % h);
    const int b = static_cast<int>(p / (static_cast<long>(w) * h));
    const long o00 = ((static_cast<long>(b) * H + 2 * ii) * W + 2 * j) * C + c;
    const long offs[4] = {o00, o00 + C, o00 + static_cast<long>(W) * C, o00 + static_cast<long>(W) * C + C};
    float v[4][8], g[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) rs_load8(x + offs[k], v[k]);
    rs_load8(gy + p * C + c, g);
    float o[4][8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      int best = 0;
      float m = v[0][q];
#pragma unroll
      for (int k = 1; k < 4; ++k)
        if (v[k][q] > m) { m = v[k][q]; best = k; }
      const float gg = (relu_gate && !(m > 0.f)) ? 0.f : g[q];
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k][q] = (k == best) ? gg : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) rs_store8(gx + offs[k], o[k]);
  }
}

template <typename T>
static int upsample_concat_impl(const void* up, int B, int h, int w, int Cu, const void* skip, int Cs,
                                const uint8_t* merged_mask, void* out, void* stream) {
  TG_REQUIRE(up && out && Cu > 0 && Cu % 8 == 0 && Cs >= 0 && Cs % 8 == 0, "tg_upsample_concat: bad arguments");
  TG_REQUIRE(Cs == 0 || skip, "tg_upsample_concat: skip missing");
  const long total = static_cast<long>(B) * h * w * ((Cu + Cs) / 8);     // one thread per 2x2 output block
  TG_REQUIRE(total < (1L << 31), "tg_upsample_concat: tensor too large for 32-bit indexing");
  upsample_concat_kernel<T><<<wave_grid(upsample_concat_kernel<T>, 256, 0, (total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const T*>(up), B, h, w, Cu, reinterpret_cast<const T*>(skip), Cs, merged_mask,
      reinterpret_cast<T*>(out));
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <typename T>
static int upsample_concat_bwd_impl(const void* d_merged, int B, int h, int w, int Cu, int Ctot, void* d_up,
                                    void* stream) {
  TG_REQUIRE(d_merged && d_up && Cu > 0 && Cu % 8 == 0 && Ctot >= Cu && Ctot % 8 == 0,
             "tg_upsample_concat_bwd: bad arguments");
  const long total = static_cast<long>(B) * ((h + kUbSeg - 1) / kUbSeg) * w * (Cu / 8);   // column strips
  TG_REQUIRE(total < (1L << 31), "tg_upsample_concat_bwd: tensor too large for 32-bit indexing");
  upsample_concat_bwd_kernel<T><<<wave_grid(upsample_concat_bwd_kernel<T>, 256, 0, (total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const T*>(d_merged), B, h, w, Cu, Ctot, reinterpret_cast<T*>(d_up));
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <typename T>
static int maxpool2_impl(const void* x, int B, int H, int W, int C, void* y, void* stream) {
  TG_REQUIRE(x && y && H % 2 == 0 && W % 2 == 0 && C % 8 == 0, "tg_maxpool2: bad arguments");
  const long total = static_cast<long>(B) * (H / 2) * (W / 2) * (C / 8);
  maxpool2_kernel<T><<<wave_grid(maxpool2_kernel<T>, 256, 0, (total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const T*>(x), B, H, W, C, reinterpret_cast<T*>(y));
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <typename T>
static int maxpool2_bwd_impl(const void* x, const void* gy, int B, int H, int W, int C, int relu_gate, void* gx,
                             void* stream) {
  TG_REQUIRE(x && gy && gx && H % 2 == 0 && W % 2 == 0 && C % 8 == 0, "tg_maxpool2_bwd: bad arguments");
  const long total = static_cast<long>(B) * (H / 2) * (W / 2) * (C / 8);
  maxpool2_bwd_kernel<T><<<wave_grid(maxpool2_bwd_kernel<T>, 256, 0, (total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const T*>(x), reinterpret_cast<const T*>(gy), B, H, W, C, relu_gate, reinterpret_cast<T*>(gx));
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace tg

#define TG_RS_TWINS(NAME, IMPL, PARAMS, ARGS)                                                   \
  extern "C" int NAME PARAMS { return tg::IMPL<__nv_bfloat16> ARGS; }                           \
  extern "C" int NAME##_f32 PARAMS { return tg::IMPL<float> ARGS; }

TG_RS_TWINS(tg_upsample_concat, upsample_concat_impl,
            (const void* up, int B, int h, int w, int Cu, const void* skip, int Cs, const uint8_t* merged_mask, void* out,
             void* stream),
            (up, B, h, w, Cu, skip, Cs, merged_mask, out, stream))
TG_RS_TWINS(tg_upsample_concat_bwd, upsample_concat_bwd_impl,
            (const void* d_merged, int B, int h, int w, int Cu, int Ctot, void* d_up, void* stream),
            (d_merged, B, h, w, Cu, Ctot, d_up, stream))
TG_RS_TWINS(tg_maxpool2, maxpool2_impl, (const void* x, int B, int H, int W, int C, void* y, void* stream),
            (x, B, H, W, C, y, stream))
TG_RS_TWINS(tg_maxpool2_bwd, maxpool2_bwd_impl,
            (const void* x, const void* gy, int B, int H, int W, int C, int relu_gate, void* gx, void* stream),
            (x, gy, B, H, W, C, relu_gate, gx, stream))
