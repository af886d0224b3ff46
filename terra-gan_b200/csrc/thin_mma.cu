// thin_mma.cu — the 1 <-> 64 channel convolutions of the path on the tensor cores.
//
// A k x k convolution with ONE input channel and 64 output channels (PConv enc1 7x7/s2, Discriminator
// model[0] 4x4/s2, VGG conv0 3x3 with its three identical input channels folded — reference pconv.py:27-43,
// discriminator.py:17, losses.py:79-89) and the data gradient of the 64 -> 1 `final` convolution
// (generator.py:56) are GEMMs with a tiny reduction dimension: [pixels] x [k*k] x [64]. On the CUDA cores they
// cost k*k*64 FMAs per pixel and ran at 1/4 of what the 128-byte-per-pixel output write allows
// (profiles/r01_*). Here every thread builds one im2col row of its pixel directly in shared memory in the
// canonical SWIZZLE_128B K-major layout, and one elected thread issues K/16 tcgen05.mma (M=128 pixels, N=64);
// the kernel is then bound by the bf16 output store (128 B per pixel).
//
// fp32 fidelity: the single-channel source and the weights are fp32 in the reference. Each is split into
// bf16 hi + lo parts and the GEMM sums src_hi*w_hi + src_lo*w_hi + src_hi*w_lo (3*k*k reduction
// elements, fp32 accumulate) — the dropped lo*lo term is ~2^-18 relative.
#include "thin_mma.cuh"
#include <cstdlib>
#include <cstring>
#include "tg_common.cuh"
#include "../../include/terragan_b200.h"

namespace tg {

template <int K>
struct RowGemmCfg {
  static constexpr int T = K * K;
  static constexpr int TP = (T + 1) / 2 * 2;             // taps padded to an even count: each section starts on a bf16x2 word
  static constexpr int KE = 3 * TP + 2;                  // src_hi*w_hi, src_lo*w_hi, src_hi*w_lo, 1*bias_hi, 1*bias_lo
  static constexpr int KP = (KE + 15) / 16 * 16;         // reduction length issued to the tensor core
  static constexpr int NKB = (KP + 63) / 64;             // 64-element (128-byte) swizzle blocks
  static constexpr int NBUF = NKB == 1 ? 2 : 1;          // A tile + accumulator double-buffered when A is small
  static constexpr int kABytes = NKB * 16384, kBBytes = NKB * 8192;
  // NBUF == 2: the bf16 output tile is staged in the A buffer its MMA has just finished reading
  static constexpr int kStageBytes = NBUF == 2 ? 0 : 4 * 4096;
  static constexpr int kSmem = NBUF * kABytes + kBBytes + kStageBytes + 4 * 2 * 64 * 4 + 64 + 1024;
};

__device__ __forceinline__ uint32_t split_hi_lo(float v) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(v, 0.f);                 // packed convert: not on the XU pipe
  const float hf = __uint_as_float(static_cast<uint32_t>(__bfloat16_as_ushort(h.x)) << 16);
  const __nv_bfloat162 l = __floats2bfloat162_rn(v - hf, 0.f);
  return static_cast<uint32_t>(__bfloat16_as_ushort(h.x)) | (static_cast<uint32_t>(__bfloat16_as_ushort(l.x)) << 16);
}

// Per CTA (128 threads, thread = pixel row = TMEM lane), software-pipelined over its tiles:
//   build A(i) from the prefetched source values -> MMA(i) issued -> prefetch source of tile i+1 ->
//   epilogue of tile i-1 (NBUF = 2) or i (NBUF = 1) while the loads / the MMA are in flight.
// The bias rides in the GEMM (two extra reduction elements 1 x bias_hi, 1 x bias_lo).
// The kernel is bound by instruction issue (ncu: 740-2500 instructions per warp and tile), so the per-tile integer work
// is kept small: pixel coordinates advance incrementally (no divisions in the loop), taps outside the image are read
// at clamped coordinates and zeroed with AND masks, the hi / lo split is done on bf16x2 pairs, ReLU on packed bf16x2.
template <int K, bool MASKED>
__global__ void __launch_bounds__(128)
rowgemm64_kernel(const __grid_constant__ RowGemmParams p, const __grid_constant__ CUtensorMap tm_out) {
  using Cfg = RowGemmCfg<K>;
  constexpr int T = Cfg::T, TP = Cfg::TP, KP = Cfg::KP, NBUF = Cfg::NBUF;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* s_a = smem;
  uint8_t* s_b = s_a + NBUF * Cfg::kABytes;
  uint8_t* s_stage = s_b + Cfg::kBBytes;                    // only when NBUF == 1
  float* s_stats = reinterpret_cast<float*>(s_stage + Cfg::kStageBytes);   // [4 warps][2][64]
  uint64_t* mbar = reinterpret_cast<uint64_t*>(s_stats + 4 * 2 * 64);      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time setup: B operand (weights hi/hi/lo, bias hi/lo), barriers, TMEM ----
  for (int i = tid; i < 64 * KP; i += 128) {
    const int n = i / KP, e = i % KP;
    unsigned short bits = 0;
    if (e < 3 * TP) {
      const int part = e / TP, t = e % TP;
      if (t < T) {
        const uint32_t hl = split_hi_lo(__ldg(p.wgt + n * p.w_sn + p.perm[t] * p.w_st));
        bits = static_cast<unsigned short>(part == 2 ? (hl >> 16) : (hl & 0xffffu));
      }
    } else if (e < 3 * TP + 2 && p.bias != nullptr) {
      const uint32_t hl = split_hi_lo(__ldg(p.bias + n));
      bits = static_cast<unsigned short>(e == 3 * TP ? (hl & 0xffffu) : (hl >> 16));
    }
    const int kb = e >> 6, ec = e & 63;
    uint8_t* dst = s_b + kb * 8192 + (n >> 3) * 1024 + (n & 7) * 128 + (((ec >> 3) ^ (n & 7)) << 4) + (ec & 7) * 2;
    *reinterpret_cast<unsigned short*>(dst) = bits;
  }
  for (int i = tid; i < 4 * 2 * 64; i += 128) s_stats[i] = 0.f;
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    fence_barrier_init();
    if (!p.out_split) tma_prefetch_desc(&tm_out);
  }
  if (warp == 0) tmem_alloc<64 * NBUF>(tmem_slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t a_addr = smem_u32(s_a), b_addr = smem_u32(s_b);
  constexpr uint32_t idesc = make_idesc_bf16(128, 64, false, false);
  float* my_stats = s_stats + warp * 2 * 64;
  const unsigned HoWo = static_cast<unsigned>(p.Ho) * p.Wo;
  const unsigned tiles = (p.total + 127u) / 128u;

  // this thread's pixel of a tile; consecutive tiles of a CTA are `step` pixels apart, decomposed once into (db, dh, dw)
  struct Geo { unsigned P; int b, ho, wo; };
  const unsigned stepP = 128u * gridDim.x;
  const int step_b = static_cast<int>(stepP / HoWo);
  const int step_h = static_cast<int>((stepP % HoWo) / p.Wo), step_w = static_cast<int>((stepP % HoWo) % p.Wo);
  auto geo_init = [&](unsigned tile) -> Geo {
    Geo g;
    g.P = tile * 128u + tid;
    g.b = static_cast<int>(g.P / HoWo);
    const unsigned rem = g.P - static_cast<unsigned>(g.b) * HoWo;
    g.ho = static_cast<int>(rem / p.Wo);
    g.wo = static_cast<int>(rem - static_cast<unsigned>(g.ho) * p.Wo);
    return g;
  };
  auto geo_step = [&](Geo& g) {
    g.P += stepP;
    g.wo += step_w;
    if (g.wo >= p.Wo) { g.wo -= p.Wo; ++g.ho; }
    g.ho += step_h;
    if (g.ho >= p.Ho) { g.ho -= p.Ho; ++g.b; }
    g.b += step_b;
  };
  const int dstep = p.flip ? -1 : 1;
  const int base_off = p.flip ? p.pad : -p.pad;

  // per-tile state that travels with the prefetched source values
  struct TileInfo { unsigned P, hbits, wbits, oidx; };

  // raw source values of this thread's pixel of a tile, read at clamped coordinates; `hbits` / `wbits` say which tap rows /
  // columns are inside the image (nothing may consume the loaded registers here: the loads stay in flight across the epilogue)
  auto load_src = [&](const Geo& g, float (&v)[T], uint8_t (&mk)[MASKED ? T : 1], TileInfo& ti) {
    const bool valid = g.P < p.total;
    const int b = valid ? g.b : 0, ho = valid ? g.ho : 0, wo = valid ? g.wo : 0;
    const int hb = ho * p.S + base_off, wb = wo * p.S + base_off;
    int hc[K], wc[K];
    unsigned hbits = 0, wbits = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int h = hb + dstep * k, w = wb + dstep * k;
      hbits |= (static_cast<unsigned>(h) < static_cast<unsigned>(p.H)) ? (1u << k) : 0u;
      wbits |= (static_cast<unsigned>(w) < static_cast<unsigned>(p.W)) ? (1u << k) : 0u;
      hc[k] = min(max(h, 0), p.H - 1) * p.W;
      wc[k] = min(max(w, 0), p.W - 1);
    }
    ti.P = g.P;
    ti.hbits = valid ? hbits : 0u;
    ti.wbits = wbits;
    ti.oidx = 0;
    if (p.out_split)
      ti.oidx = ((static_cast<unsigned>(b) * 4u + 2u * (ho & 1) + (wo & 1)) * (p.Ho >> 1) + (ho >> 1)) * (p.Wo >> 1) + (wo >> 1);
    const float* sb = p.src + static_cast<size_t>(b) * p.H * p.W;
    const uint8_t* mb = MASKED ? p.src_mask + static_cast<size_t>(b) * p.H * p.W : nullptr;
#pragma unroll
    for (int kh = 0; kh < K; ++kh) {
#pragma unroll
      for (int kw = 0; kw < K; ++kw) {
        const int off = hc[kh] + wc[kw];
        v[kh * K + kw] = __ldg(sb + off);
        if (MASKED) mk[kh * K + kw] = __ldg(mb + off);
      }
    }
  };

  // L2 prefetch of the source rows of a tile a few iterations ahead: its first touch is a DRAM read that queues
  // behind this kernel's own write stream (measured: 2x the kernel time when left on the critical path)
  auto prefetch_src = [&](const Geo& g) {
    if (g.P >= p.total) return;
    const int hb = g.ho * p.S + base_off;
    const int w = min(g.wo * p.S, p.W - 1);
    const size_t img = static_cast<size_t>(g.b) * p.H * p.W;
#pragma unroll
    for (int kh = 0; kh < K; ++kh) {
      const int h = hb + dstep * kh;
      if (static_cast<unsigned>(h) < static_cast<unsigned>(p.H)) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.src + img + h * p.W + w));
        if (MASKED) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.src_mask + img + h * p.W + w));
      }
    }
  };

  auto build_a = [&](uint8_t* a_tile, const float (&v)[T], const uint8_t (&mk)[MASKED ? T : 1], const TileInfo& ti) {
    uint32_t mh[K], mw[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      mh[k] = 0u - ((ti.hbits >> k) & 1u);
      mw[k] = 0u - ((ti.wbits >> k) & 1u);
    }
    uint32_t hiw[TP / 2], low[TP / 2];
#pragma unroll
    for (int q = 0; q < TP / 2; ++q) {
      float x[2];
#pragma unroll
      for (int z = 0; z < 2; ++z) {
        const int t = 2 * q + z;           // compile-time
        uint32_t bits = 0;
        if (t < T) {
          bits = __float_as_uint(v[t]) & mh[t / K] & mw[t % K];
          if (MASKED) bits = mk[t] != 0 ? bits : 0u;
        }
        x[z] = __uint_as_float(bits);
      }
      const uint32_t h = pack_bf16x2(x[0], x[1]);
      hiw[q] = h;
      low[q] = pack_bf16x2(x[0] - __uint_as_float(h << 16), x[1] - __uint_as_float(h & 0xffff0000u));
    }
    const uint32_t one2 = ti.P < p.total ? 0x3f803f80u : 0u;   // bf16 (1, 1): switches the bias (hi, lo) on for real pixels
    uint8_t* rowp = a_tile + (tid >> 3) * 1024 + (tid & 7) * 128;
#pragma unroll
    for (int j = 0; j < KP / 8; ++j) {
      uint32_t wd[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int w = j * 4 + q;           // compile-time word index: elements 2w, 2w + 1
        uint32_t x32 = 0;
        if (w < TP / 2) x32 = hiw[w];
        else if (w < TP) x32 = low[w - TP / 2];
        else if (w < 3 * TP / 2) x32 = hiw[w - TP];
        else if (w == 3 * TP / 2) x32 = one2;
        wd[q] = x32;
      }
      *reinterpret_cast<uint4*>(rowp + (j >> 3) * 16384 + ((((j & 7) ^ (tid & 7))) << 4)) =
          make_uint4(wd[0], wd[1], wd[2], wd[3]);
    }
  };

  // thread = pixel row = TMEM lane: ratio, stats, activation, bf16; the warp's 32 x 128 B sub-tile is staged in
  // SWIZZLE_128B order and leaves through one TMA tensor store (plain layout) or 4-rows-per-instruction stores
  auto epilogue = [&](unsigned tile, unsigned P, unsigned oidx, uint32_t t_acc, uint8_t* stage_tile) {
    const bool valid = P < p.total;
    uint4* st = reinterpret_cast<uint4*>(stage_tile) + warp * 256;
    float rs = 1.f;
    if (p.code != nullptr && valid) rs = __ldg(p.lut + p.code[P]);
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      uint32_t raw[32];
      tmem_ld_32x32(t_acc + (static_cast<uint32_t>(warp * 32) << 16) + ch * 32, raw);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
      if (p.code != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= rs;
      }
      if (p.stats != nullptr) {
        float sm[32], sq[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          sm[j] = v[j];
          sq[j] = v[j] * v[j];
        }
        const float csum = warp_transpose_sum32(sm);
        const float csq = warp_transpose_sum32(sq);
        my_stats[ch * 32 + lane] += csum;
        my_stats[64 + ch * 32 + lane] += csq;
      }
      if (p.act == 2) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], v[j] * p.slope);   // slope in [0, 1)
      }
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
      if (p.act == 1) {          // ReLU commutes with the rounding: max on the packed pairs
        const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const __nv_bfloat162 m = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&pk[j]), zero2);
          pk[j] = *reinterpret_cast<const uint32_t*>(&m);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        st[lane * 8 + ((ch * 4 + j) ^ (lane & 7))] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
    }
    if (!p.out_split) {
      fence_proxy_async();
      __syncwarp();
      if (lane == 0 && !TG_DBG(p, 1)) {
        tma_store_2d(&tm_out, st, 0, static_cast<int>(tile * 128u + warp * 32));   // rows >= total are clipped
        bulk_commit_group();
      }
    } else {
      const unsigned vmask = __ballot_sync(0xffffffffu, valid);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = i * 4 + (lane >> 3), chunk = lane & 7;
        const unsigned ridx = __shfl_sync(0xffffffffu, oidx, row);
        if ((vmask >> row) & 1u)
          reinterpret_cast<uint4*>(p.out + static_cast<size_t>(ridx) * 64)[chunk] = st[row * 8 + (chunk ^ (row & 7))];
      }
      __syncwarp();
    }
  };

  float v[T];
  uint8_t mk[MASKED ? T : 1];
  TileInfo cur{0u, 0u, 0u, 0u};
  Geo gl = geo_init(blockIdx.x);              // geometry of the tile whose source is loaded next
  Geo gp = geo_init(blockIdx.x + gridDim.x);  // geometry of the tile whose source rows are pulled into L2 next
  for (int d = 1; d < 4; ++d) {
    prefetch_src(gp);
    geo_step(gp);
  }
  if (blockIdx.x < tiles) load_src(gl, v, mk, cur);
  geo_step(gl);
  unsigned prev_tile = 0, prev_P = 0, prev_oidx = 0;
  int it = 0;
  for (unsigned tile = blockIdx.x;; tile += gridDim.x, ++it) {
    const bool have = tile < tiles;          // CTA-uniform
    const int buf = it % NBUF;
    if (NBUF == 2) {
      // this warp's rows of A[buf] were the staging area of tile it-2: its TMA store must have read them
      if (lane == 0) bulk_wait_read_all();
      __syncwarp();
    }
    if (have) build_a(s_a + buf * Cfg::kABytes, v, mk, cur);
    const unsigned this_P = cur.P, this_oidx = cur.oidx;
    fence_proxy_async();          // generic-proxy writes -> visible to the tensor core (async proxy)
    tc_fence_before();            // earlier tcgen05.ld of this accumulator are complete before it is overwritten
    __syncthreads();
    if (have && tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < KP / 16; ++ks) {
        const uint32_t off_a = buf * Cfg::kABytes + (ks >> 2) * 16384 + (ks & 3) * 32;
        const uint32_t off_b = (ks >> 2) * 8192 + (ks & 3) * 32;
        umma_bf16(tmem + buf * 64, make_smem_desc(a_addr + off_a, 16, 1024), make_smem_desc(b_addr + off_b, 16, 1024),
                  idesc, ks ? 1u : 0u);
      }
      umma_commit(&mbar[buf]);
    }
    if (have && tile + gridDim.x < tiles) load_src(gl, v, mk, cur);
    geo_step(gl);
    if (!TG_DBG(p, 4)) prefetch_src(gp);
    geo_step(gp);
    if (NBUF == 1) {
      if (have) {
        mbar_wait(&mbar[0], it & 1);
        tc_fence_after();
        if (lane == 0) bulk_wait_read_all();     // previous tile's store has left the staging area
        __syncwarp();
        epilogue(tile, this_P, this_oidx, tmem, s_stage);
      }
    } else {
      if (it > 0) {
        const int pb = (it - 1) % NBUF;
        mbar_wait(&mbar[pb], ((it - 1) / NBUF) & 1);
        tc_fence_after();
        epilogue(prev_tile, prev_P, prev_oidx, tmem + pb * 64, s_a + pb * Cfg::kABytes);
      }
      prev_tile = tile;
      prev_P = this_P;
      prev_oidx = this_oidx;
    }
    if (!have) break;
  }

  if (lane == 0) bulk_wait_read_all();
  tc_fence_before();
  __syncthreads();
  if (p.stats != nullptr) {
    const int c = tid & 63, which = tid >> 6;
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) s += s_stats[q * 128 + which * 64 + c];
    p.stats[(static_cast<size_t>(blockIdx.x) * 2 + which) * 64 + c] = s;
  }
  if (warp == 0) tmem_dealloc<64 * NBUF>(tmem);
}

template <int K, bool MASKED>
static int rowgemm_launch(const RowGemmParams& p, int grid_cap, int* grid_used, cudaStream_t st) {
  using Cfg = RowGemmCfg<K>;
  TG_SET_SMEM_ONCE((rowgemm64_kernel<K, MASKED>), Cfg::kSmem);
  CUtensorMap tm_out;
  memset(&tm_out, 0, sizeof(tm_out));
  if (!p.out_split) {
    const uint64_t dims[2] = {64, p.total};
    const uint64_t str[1] = {128};
    const uint32_t box[2] = {64, 32};
    if (make_tmap_bf16(&tm_out, p.out, 2, dims, str, box) != 0) return -3;
  }
  int per_sm = (220 * 1024) / Cfg::kSmem;
  if (per_sm > 512 / (64 * Cfg::NBUF)) per_sm = 512 / (64 * Cfg::NBUF);     // TMEM columns: a CTA beyond this waits in tmem_alloc
  if (per_sm > 5) per_sm = 5;
  const long tiles = (static_cast<long>(p.total) + 127) / 128;
  long grid = static_cast<long>(num_sms()) * per_sm;
  if (grid > tiles) grid = tiles;
  if (grid_cap > 0 && grid > grid_cap) grid = grid_cap;
  if (grid < 1) grid = 1;
  if (grid_used) *grid_used = static_cast<int>(grid);
  RowGemmParams q = p;
#ifdef TG_PERF_DEBUG
  {
    static const int dbg = [] { const char* e = getenv("TG_THIN_DEBUG"); return e ? atoi(e) : 0; }();
    q.debug = dbg;
  }
#endif
  rowgemm64_kernel<K, MASKED><<<static_cast<int>(grid), 128, Cfg::kSmem, st>>>(q, tm_out);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Weight gradient:  D[row][c] += sum over the 128 pixels of a tile of A[row][pixel] * Y[pixel][c]
//   A (thread-built): rows t / TP+t = bf16 hi / lo part of the tap-t source value of each pixel (TP = taps padded to an
//   even count), row 2TP = 1 (bias gradient). A is stored MN-major — [pixel][64 rows] 128-byte lines, SWIZZLE_128B, two
//   64-row blocks — so a thread writes the column of its pixel as a few 16-byte chunks instead of one 2-byte store per row.
//   The Y tile [128 pixels][64] arrives by TMA and is the MN-major B operand.
// One accumulator (128 lanes x 64 columns) lives in TMEM for the whole kernel and is read once at the end.
// warp 0: TMA producer, warp 1: MMA issuer, warp 2: TMEM allocator, then kGroups groups of 128 A builders taking tiles
// round-robin (a tile's source-load latency is otherwise exposed: ncu long-scoreboard stalls); group 0 reads out.
// ------------------------------------------------------------------------------------------------
template <int K>
struct TapWgradCfg {
  static constexpr int T = K * K;
  static constexpr int TP = (T + 1) / 2 * 2;
  static constexpr int NR = 2 * TP + 1;                     // accumulator rows in use
  static constexpr int kChunks = (NR + 7) / 8;              // 16-byte chunks (8 rows) a builder thread writes per pixel
  static constexpr int kGroups = K <= 4 ? 4 : 2;            // register budget: the 7x7 builder holds 49 values + 49 mask bytes
  static constexpr int kStages = K <= 4 ? 4 : 3;
  static constexpr int kThreads = 128 + kGroups * 128;
  static constexpr int kSmem = kStages * (32768 + 16384) + 256 + 1024;
  static_assert(NR <= 128, "tap rows must fit the 128 accumulator lanes");
};

template <int K, bool MASKED>
__global__ void __launch_bounds__(TapWgradCfg<K>::kThreads, 1)
tapwgrad_kernel(const __grid_constant__ TapWgradParams p, const __grid_constant__ CUtensorMap tm_y) {
  using Cfg = TapWgradCfg<K>;
  constexpr int T = Cfg::T, TP = Cfg::TP, NR = Cfg::NR, kStages = Cfg::kStages, kGroups = Cfg::kGroups;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* s_a = smem;                                   // [stage][2 row blocks][128 pixels][128 B]
  uint8_t* s_y = s_a + kStages * 32768;                  // [stage][128 pixels][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_y + kStages * 16384);
  uint64_t* full_y = bars;
  uint64_t* full_a = bars + kStages;
  uint64_t* empty = bars + 2 * kStages;
  uint64_t* done = bars + 3 * kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kStages + 1);
  float* s_red = reinterpret_cast<float*>(tmem_slot + 2);      // [4 * kGroups]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // rows >= NR of the A tiles are never written by the builders: zero once
  for (int i = tid; i < kStages * 32768 / 16; i += Cfg::kThreads) reinterpret_cast<uint4*>(s_a)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_y[i], 1);
      mbar_init(&full_a[i], 128);
      mbar_init(&empty[i], 1);
    }
    mbar_init(done, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tm_y);
  }
  if (warp == 2) tmem_alloc<64>(tmem_slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const unsigned tiles = (p.total + 127u) / 128u;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (unsigned tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_y[stage], 16384);
        tma_load_2d(s_y + stage * 16384, &tm_y, &full_y[stage], 0, static_cast<int>(tile * 128u));   // tail rows: zero fill
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    {   // whole warp (warp-uniform addressing); one elected lane issues
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, true, true);     // A and B (= Y) MN-major
      const uint64_t da0 = make_smem_desc(smem_u32(s_a), 16384, 1024), dy0 = make_smem_desc(smem_u32(s_y), 16384, 1024);
      const uint32_t a_lo0 = static_cast<uint32_t>(da0), a_hi = static_cast<uint32_t>(da0 >> 32);
      const uint32_t y_lo0 = static_cast<uint32_t>(dy0), y_hi = static_cast<uint32_t>(dy0 >> 32);
      int stage = 0;
      uint32_t phase = 0;
      bool first = true;
      for (unsigned tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        mbar_wait(&full_y[stage], phase);
        mbar_wait(&full_a[stage], phase);
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + static_cast<uint32_t>(stage) * (32768 >> 4);
        const uint32_t y_lo = y_lo0 + static_cast<uint32_t>(stage) * (16384 >> 4);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)       // 16 pixels (two 8-line swizzle atoms, 2048 B) per MMA
            umma_bf16_lh(tmem, a_lo + ks * 128, a_hi, y_lo + ks * 128, y_hi, idesc, ks ? 1u : (first ? 0u : 1u));
          umma_commit(&empty[stage]);
        }
        first = false;
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) umma_commit(done);
    }
  } else if (warp >= 4) {
    const int j = (tid - 128) & 127;             // pixel of the tile (= TMEM lane at read-out for group 0)
    const int grp = (tid - 128) >> 7;            // builder group: tiles i = grp, grp + kGroups, ...
    // pixel index -> (image, plane, row, column) of the Y grid as mixed-radix digits; consecutive tiles of a group are
    // `stepP` pixels apart, decomposed once, so the loop carries digits instead of dividing
    const int NP = p.y_split ? 4 : 1;
    const int Hd = p.y_split ? p.Ho >> 1 : p.Ho, Wd = p.y_split ? p.Wo >> 1 : p.Wo;
    struct Geo { unsigned P; int b, pl, h, w; };
    auto decompose = [&](unsigned P) -> Geo {
      Geo g;
      g.P = P;
      g.w = static_cast<int>(P % static_cast<unsigned>(Wd));
      P /= static_cast<unsigned>(Wd);
      g.h = static_cast<int>(P % static_cast<unsigned>(Hd));
      P /= static_cast<unsigned>(Hd);
      g.pl = static_cast<int>(P % static_cast<unsigned>(NP));
      g.b = static_cast<int>(P / static_cast<unsigned>(NP));
      return g;
    };
    const unsigned stepP = 128u * gridDim.x * kGroups;
    const Geo st = decompose(stepP);
    auto geo_step = [&](Geo& g) {
      g.P += stepP;
      g.w += st.w;
      if (g.w >= Wd) { g.w -= Wd; ++g.h; }
      g.h += st.h;
      if (g.h >= Hd) { g.h -= Hd; ++g.pl; }
      g.pl += st.pl;
      if (g.pl >= NP) { g.pl -= NP; ++g.b; }
      g.b += st.b;
    };
    const int dstep = p.flip ? -1 : 1;
    const int base_off = p.flip ? p.pad : -p.pad;
    // loads only, at clamped coordinates; nothing consumes the registers before the next build (they stay in flight)
    auto load_src = [&](const Geo& g, float (&v)[T], uint8_t (&mk)[MASKED ? T : 1], unsigned& hbits, unsigned& wbits) {
      const bool valid = g.P < p.total;
      const int b = valid ? g.b : 0;
      int ho = valid ? g.h : 0, wo = valid ? g.w : 0;
      if (p.y_split) {
        const int pl = valid ? g.pl : 0;
        ho = 2 * ho + (pl >> 1);
        wo = 2 * wo + (pl & 1);
      }
      const int hb = ho * p.S + base_off, wb = wo * p.S + base_off;
      int hc[K], wc[K];
      unsigned hb_ = 0, wb_ = 0;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int h = hb + dstep * k, w = wb + dstep * k;
        hb_ |= (static_cast<unsigned>(h) < static_cast<unsigned>(p.H)) ? (1u << k) : 0u;
        wb_ |= (static_cast<unsigned>(w) < static_cast<unsigned>(p.W)) ? (1u << k) : 0u;
        hc[k] = min(max(h, 0), p.H - 1) * p.W;
        wc[k] = min(max(w, 0), p.W - 1);
      }
      hbits = valid ? hb_ : 0u;
      wbits = wb_;
      const float* sb = p.src + static_cast<size_t>(b) * p.H * p.W;
      const uint8_t* mb = MASKED ? p.src_mask + static_cast<size_t>(b) * p.H * p.W : nullptr;
#pragma unroll
      for (int kh = 0; kh < K; ++kh) {
#pragma unroll
        for (int kw = 0; kw < K; ++kw) {
          const int off = hc[kh] + wc[kw];
          v[kh * K + kw] = __ldg(sb + off);
          if (MASKED) mk[kh * K + kw] = __ldg(mb + off);
        }
      }
    };

    float v[T];
    uint8_t mk[MASKED ? T : 1];
    unsigned hbits = 0, wbits = 0;
    float csum = 0.f;
    const unsigned my_tiles = blockIdx.x < tiles ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
    Geo g = decompose((blockIdx.x + static_cast<unsigned>(grp) * gridDim.x) * 128u + j);
    bool valid = g.P < p.total;
    if (static_cast<unsigned>(grp) < my_tiles) load_src(g, v, mk, hbits, wbits);
    // pixel j of the tile = line j of each 64-row block; its 16-byte chunk c (rows 8c .. 8c+7) sits at (c ^ (j & 7)) << 4
    uint8_t* const line = s_a + j * 128;
    for (unsigned i = grp; i < my_tiles; i += kGroups) {
      const int stage = static_cast<int>(i % kStages);
      const uint32_t phase = (i / kStages) & 1u;
      uint32_t mh[K], mw[K];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        mh[k] = 0u - ((hbits >> k) & 1u);
        mw[k] = 0u - ((wbits >> k) & 1u);
      }
      uint32_t hiw[TP / 2], low[TP / 2];
#pragma unroll
      for (int q = 0; q < TP / 2; ++q) {
        float x[2];
#pragma unroll
        for (int z = 0; z < 2; ++z) {
          const int t = 2 * q + z;           // compile-time
          uint32_t bits = 0;
          if (t < T) {
            bits = __float_as_uint(v[t]) & mh[t / K] & mw[t % K];
            if (MASKED) bits = mk[t] != 0 ? bits : 0u;
          }
          x[z] = __uint_as_float(bits);
          if (t < T && t == p.center) csum += x[z];
        }
        const uint32_t h = pack_bf16x2(x[0], x[1]);
        hiw[q] = h;
        low[q] = pack_bf16x2(x[0] - __uint_as_float(h << 16), x[1] - __uint_as_float(h & 0xffff0000u));
      }
      const uint32_t one = valid ? 0x3f80u : 0u;        // row 2TP = 1 for real pixels (bias gradient = sum of Y)
      mbar_wait(&empty[stage], phase ^ 1);
      uint8_t* a_st = line + stage * 32768;
#pragma unroll
      for (int c = 0; c < Cfg::kChunks; ++c) {
        uint32_t wd[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int w = c * 4 + q;           // compile-time word index: rows 2w, 2w + 1
          uint32_t x32 = 0;
          if (w < TP / 2) x32 = hiw[w];
          else if (w < TP) x32 = low[w - TP / 2];
          else if (w == TP) x32 = one;
          wd[q] = x32;
        }
        *reinterpret_cast<uint4*>(a_st + (c >> 3) * 16384 + (((c & 7) ^ (j & 7)) << 4)) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
      }
      fence_proxy_async();
      mbar_arrive(&full_a[stage]);
      if (i + kGroups < my_tiles) {
        geo_step(g);
        valid = g.P < p.total;
        load_src(g, v, mk, hbits, wbits);
      }
    }
    if (p.partial_c != nullptr) {
      const float ws = warp_sum(csum);
      if (lane == 0) s_red[warp - 4] = ws;
    }
    if (grp == 0) {
    // ---- read-out: lane j of the accumulator = row j ----
    mbar_wait(done, 0);
    tc_fence_after();
    float* dst = p.partial + (static_cast<size_t>(blockIdx.x) * NR + j) * 64;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      uint32_t raw[32];
      tmem_ld_32x32(tmem + (static_cast<uint32_t>((warp - 4) * 32) << 16) + ch * 32, raw);
      tmem_ld_wait();
      if (j < NR) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          reinterpret_cast<float4*>(dst + ch * 32)[q] = make_float4(__uint_as_float(raw[4 * q]), __uint_as_float(raw[4 * q + 1]),
                                                                   __uint_as_float(raw[4 * q + 2]), __uint_as_float(raw[4 * q + 3]));
      }
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0 && p.partial_c != nullptr) {
    float s = 0.f;
    for (int i = 0; i < 4 * kGroups; ++i) s += s_red[i];
    p.partial_c[blockIdx.x] = s;
  }
  if (warp == 2) tmem_dealloc<64>(tmem);
}

__global__ void tapwgrad_reduce_kernel(const float* __restrict__ partial, const float* __restrict__ partial_c, int rows, int T,
                                       float* __restrict__ out_w, int w_sn, int w_st, RowGemmParams perm_holder,
                                       float* __restrict__ out_b, int bias_mode, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int TP = (T + 1) / 2 * 2, NR = 2 * TP + 1;
  if (i < 64 * T) {
    const int t = i / 64, c = i % 64;
    double s = 0.0;
    for (int r = 0; r < rows; ++r) {
      const float* pr = partial + static_cast<size_t>(r) * NR * 64;
      s += static_cast<double>(pr[t * 64 + c]) + static_cast<double>(pr[(TP + t) * 64 + c]);
    }
    float* d = out_w + c * w_sn + perm_holder.perm[t] * w_st;
    *d = (accumulate ? *d : 0.f) + static_cast<float>(s);
  } else if (i < 64 * T + 64 && bias_mode == 1 && out_b != nullptr) {
    const int c = i - 64 * T;
    double s = 0.0;
    for (int r = 0; r < rows; ++r) s += partial[(static_cast<size_t>(r) * NR + 2 * TP) * 64 + c];
    out_b[c] = (accumulate ? out_b[c] : 0.f) + static_cast<float>(s);
  } else if (i == 64 * T + 64 && bias_mode == 2 && out_b != nullptr && partial_c != nullptr) {
    double s = 0.0;
    for (int r = 0; r < rows; ++r) s += partial_c[r];
    out_b[0] = (accumulate ? out_b[0] : 0.f) + static_cast<float>(s);
  }
}

template <int K, bool MASKED>
static int tapwgrad_launch(const TapWgradParams& p, const void* y, int grid_cap, int* grid_used, cudaStream_t st) {
  using Cfg = TapWgradCfg<K>;
  TG_SET_SMEM_ONCE((tapwgrad_kernel<K, MASKED>), Cfg::kSmem);
  CUtensorMap tm_y;
  const uint64_t dims[2] = {64, p.total};
  const uint64_t str[1] = {128};
  const uint32_t box[2] = {64, 128};
  if (make_tmap_bf16(&tm_y, y, 2, dims, str, box) != 0) return -3;
  const long tiles = (static_cast<long>(p.total) + 127) / 128;
  long grid = num_sms();
  if (grid > tiles) grid = tiles;
  if (grid_cap > 0 && grid > grid_cap) grid = grid_cap;
  if (grid < 1) grid = 1;
  if (grid_used) *grid_used = static_cast<int>(grid);
  tapwgrad_kernel<K, MASKED><<<static_cast<int>(grid), Cfg::kThreads, Cfg::kSmem, st>>>(p, tm_y);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int tapwgrad_dispatch(int k, const TapWgradParams& p, const void* y, int grid_cap, int* grid_used, cudaStream_t st) {
  const bool m = p.src_mask != nullptr;
  if (k == 3) return m ? tapwgrad_launch<3, true>(p, y, grid_cap, grid_used, st) : tapwgrad_launch<3, false>(p, y, grid_cap, grid_used, st);
  if (k == 4) return m ? tapwgrad_launch<4, true>(p, y, grid_cap, grid_used, st) : tapwgrad_launch<4, false>(p, y, grid_cap, grid_used, st);
  if (k == 7) return m ? tapwgrad_launch<7, true>(p, y, grid_cap, grid_used, st) : tapwgrad_launch<7, false>(p, y, grid_cap, grid_used, st);
  return -1;
}

int tapwgrad_reduce(int k, const float* partial, const float* partial_c, int rows, float* out_w, int w_sn, int w_st,
                    const int8_t* perm, float* out_b, int bias_mode, int accumulate, cudaStream_t st) {
  RowGemmParams holder{};
  for (int t = 0; t < k * k; ++t) holder.perm[t] = perm[t];
  const int n = 64 * k * k + 65;
  tapwgrad_reduce_kernel<<<(n + 127) / 128, 128, 0, st>>>(partial, partial_c, rows, k * k, out_w, w_sn, w_st, holder, out_b,
                                                          bias_mode, accumulate);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// 64 -> 1 gather convolution, step 1:  T[t][pixel] = sum_c w[t][c] * x[pixel][c]   (t < ntaps <= 32)
//   A = the tap weight vectors as rows 0..ntaps-1 of a 128-row K-major tile (hi and lo bf16 parts in two tiles,
//   accumulated into the same rows), B = 256 activation pixels x 64 channels per TMA box (storage order, so any
//   layout of x works), D = 128 lanes x 256 columns in TMEM, double-buffered. Only lanes < ntaps carry data:
//   the instruction costs the same 128 clocks either way and the kernel stays HBM-bound on reading x.
// warp 0: TMA producer, warp 1: MMA issuer, warp 2: TMEM allocator, warps 4 and 8 (TMEM lane quarter 0): read-out.
// ------------------------------------------------------------------------------------------------
constexpr int kTdStages = 4;
constexpr int kTdN = 256;
constexpr int kTdSmem = 2 * 16384 + kTdStages * kTdN * 128 + 256 + 1024;

__global__ void __launch_bounds__(384, 1)
tapdot_kernel(const __grid_constant__ CUtensorMap tm_x, const float* __restrict__ wgt /*[ntaps][64]*/, int ntaps,
              unsigned total, float* __restrict__ T) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* s_w = smem;                                    // [hi | lo][128 rows][128 B]
  uint8_t* s_x = s_w + 2 * 16384;                         // [stage][256 pixels][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_x + kTdStages * kTdN * 128);
  uint64_t* full = bars;                                  // [kTdStages]
  uint64_t* empty = bars + kTdStages;                     // [kTdStages]
  uint64_t* tfull = bars + 2 * kTdStages;                 // [2]
  uint64_t* tempty = bars + 2 * kTdStages + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kTdStages + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < 2 * 16384 / 16; i += 384) reinterpret_cast<uint4*>(s_w)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  for (int i = tid; i < ntaps * 64; i += 384) {
    const int m = i >> 6, c = i & 63;
    const uint32_t hl = split_hi_lo(__ldg(wgt + i));
    uint8_t* dst = s_w + (m >> 3) * 1024 + (m & 7) * 128 + (((c >> 3) ^ (m & 7)) << 4) + (c & 7) * 2;
    *reinterpret_cast<unsigned short*>(dst) = static_cast<unsigned short>(hl & 0xffffu);
    *reinterpret_cast<unsigned short*>(dst + 16384) = static_cast<unsigned short>(hl >> 16);
  }
  if (tid == 0) {
    for (int i = 0; i < kTdStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 1);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tm_x);
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const unsigned tiles = (total + kTdN - 1) / kTdN;
  const unsigned my_tiles = blockIdx.x < tiles ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;

  if (warp == 0) {
    if (lane == 0) {
      for (unsigned i = 0; i < my_tiles; ++i) {
        const int stage = static_cast<int>(i % kTdStages);
        mbar_wait(&empty[stage], ((i / kTdStages) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&full[stage], kTdN * 128);
        tma_load_2d(s_x + stage * kTdN * 128, &tm_x, &full[stage], 0,
                    static_cast<int>((blockIdx.x + i * gridDim.x) * kTdN));      // tail rows: zero fill
      }
    }
  } else if (warp == 1) {
    {   // whole warp (warp-uniform addressing); one elected lane issues
      constexpr uint32_t idesc = make_idesc_bf16(128, kTdN, false, false);
      const uint64_t dw0 = make_smem_desc(smem_u32(s_w), 16, 1024), dx0 = make_smem_desc(smem_u32(s_x), 16, 1024);
      const uint32_t w_lo0 = static_cast<uint32_t>(dw0), x_lo0 = static_cast<uint32_t>(dx0), d_hi = static_cast<uint32_t>(dw0 >> 32);
      for (unsigned i = 0; i < my_tiles; ++i) {
        const int stage = static_cast<int>(i % kTdStages);
        const int acc = static_cast<int>(i & 1u);
        mbar_wait(&tempty[acc], ((i >> 1) & 1u) ^ 1u);
        mbar_wait(&full[stage], (i / kTdStages) & 1u);
        tc_fence_after();
        const uint32_t x_lo = x_lo0 + static_cast<uint32_t>(stage) * (kTdN * 128 >> 4);
        if (elect_one()) {
#pragma unroll
          for (int part = 0; part < 2; ++part) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_bf16_lh(tmem + acc * kTdN, w_lo0 + part * (16384 >> 4) + ks * 2, d_hi, x_lo + ks * 2, d_hi, idesc,
                           (part | ks) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          umma_commit(&tfull[acc]);
        }
      }
    }
  } else if (warp == 4 || warp == 8) {
    // lane t of TMEM quarter 0 = tap t; warp 4 drains accumulator 0, warp 8 accumulator 1
    const int acc = warp == 4 ? 0 : 1;
    for (unsigned i = acc; i < my_tiles; i += 2) {
      mbar_wait(&tfull[acc], (i >> 1) & 1u);
      tc_fence_after();
      const unsigned p0 = (blockIdx.x + i * gridDim.x) * kTdN;
      float* trow = T + static_cast<size_t>(lane) * total + p0;
#pragma unroll 2
      for (int ch = 0; ch < kTdN / 32; ++ch) {
        uint32_t raw[32];
        tmem_ld_32x32(tmem + acc * kTdN + ch * 32, raw);
        tmem_ld_wait();
        if (lane < ntaps) {
          if (p0 + ch * 32 + 32 <= total) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              reinterpret_cast<float4*>(trow + ch * 32)[q] =
                  make_float4(__uint_as_float(raw[4 * q]), __uint_as_float(raw[4 * q + 1]), __uint_as_float(raw[4 * q + 2]),
                              __uint_as_float(raw[4 * q + 3]));
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q)
              if (p0 + ch * 32 + q < total) trow[ch * 32 + q] = __uint_as_float(raw[q]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem);
}

// step 2: out[b][oh][ow] = bias + sum over the taps t of the pixel's class of T[t][storage index of base + d_t]
__global__ void __launch_bounds__(256)
tapsum_kernel(const float* __restrict__ T, unsigned total_in, int x_split, int B, int H, int W, To1Taps taps,
              const float* __restrict__ bias, int Ho, int Wo, int mode, const uint8_t* __restrict__ mask,
              const float* __restrict__ xin, float* __restrict__ out, float* __restrict__ sig_out) {
  const unsigned M = static_cast<unsigned>(B) * Ho * Wo;
  const float bv = bias ? __ldg(bias) : 0.f;
  const unsigned HW = static_cast<unsigned>(H) * W, HoWo = static_cast<unsigned>(Ho) * Wo;
  for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < M; p += gridDim.x * blockDim.x) {
    const unsigned b = p / HoWo, rem = p - b * HoWo;
    const int oh = static_cast<int>(rem / Wo), ow = static_cast<int>(rem - static_cast<unsigned>(oh) * Wo);
    int cls = 0, bh = oh, bw = ow;
    if (taps.ncls == 4) {
      cls = 2 * (oh & 1) + (ow & 1);
      bh = oh >> 1;
      bw = ow >> 1;
    }
    float acc = 0.f;
    for (int t = taps.begin[cls]; t < taps.begin[cls] + taps.count[cls]; ++t) {
      const int h = bh + taps.dh[t], w = bw + taps.dw[t];
      if (h < 0 || h >= H || w < 0 || w >= W) continue;
      const unsigned sidx = x_split
          ? ((b * 4u + 2u * (h & 1) + (w & 1)) * (H >> 1) + (h >> 1)) * (W >> 1) + (w >> 1)
          : b * HW + static_cast<unsigned>(h) * W + w;
      acc += __ldg(T + static_cast<size_t>(t) * total_in + sidx);
    }
    const float v = acc + bv;
    if (mode == 0) {
      out[p] = v;
    } else {
      const float sg = 1.f / (1.f + __expf(-v));
      if (sig_out) sig_out[p] = sg;
      const float m = mask[p] ? 1.f : 0.f;
      out[p] = sg * (1.f - m) + xin[p] * m;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 64 -> 1 convolution with taps inside a 3x3 window, one kernel: a 16 x 16 pixel patch (interior 14 x 14 + halo) arrives as
// ONE TMA box (zero-filled outside the image = the zero padding), two MMAs (M = 128 patch pixels each, N = 32: the tap
// weight vectors as bf16 hi parts in columns 0..15 and lo parts in columns 16..31, K = 64 channels) leave
// T[pixel][tap] with pixel = TMEM lane; each epilogue thread owns one patch pixel: it adds hi + lo, parks its <= 16 tap
// products in shared memory, and after a barrier the interior pixels gather theirs at the tap offsets, add the bias and
// apply the sigmoid / composite epilogue (generator.py:56-62). Compared with tapdot + tapsum this drops the [taps][pixels]
// fp32 scratch round trip (72 B per pixel next to the 128 B read of x) and 3/4 of the MMA work (N = 32 instead of
// 128 mostly idle accumulator rows x 256 pixels).
// warp 0: TMA producer, warp 1: MMA issuer, warp 2: TMEM allocator, warps 4-7 / 8-11: two epilogue groups on alternate tiles.
// ------------------------------------------------------------------------------------------------
constexpr int kTfStages = 4;
constexpr int kTfPatch = 16, kTfIn = 14;                    // patch edge, interior edge
constexpr int kTfSmem = kTfStages * 32768 + 4096 + 4 * 16 * 256 * 4 + 256 + 1024;

struct To1FusedParams {
  int B, H, W, tiles_h, tiles_w, ntaps;
  int8_t dh[16], dw[16];
  const float* wgt;        // [ntaps][64]
  const float* bias;       // optional [1]
  int mode;                // 0: out = conv + bias; 1: out = sigmoid(conv + bias) * (1 - mask) + xin * mask
  const uint8_t* mask;
  const float* xin;
  float* out;
  float* sig_out;
};

__global__ void __launch_bounds__(384, 1)
to1_fused_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ To1FusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* s_x = smem;                                     // [stage][256 patch pixels][128 B]
  uint8_t* s_w = s_x + kTfStages * 32768;                  // [32 rows: hi taps | lo taps][128 B]
  float* s_t = reinterpret_cast<float*>(s_w + 4096);       // [2 groups][2][16 taps][256 pixels]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_t + 4 * 16 * 256);
  uint64_t* full = bars;                                   // [kTfStages]
  uint64_t* empty = bars + kTfStages;                      // [kTfStages]
  uint64_t* tfull = bars + 2 * kTfStages;                  // [2]
  uint64_t* tempty = bars + 2 * kTfStages + 2;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kTfStages + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < 4096 / 16; i += 384) reinterpret_cast<uint4*>(s_w)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  for (int i = tid; i < p.ntaps * 64; i += 384) {
    const int t = i >> 6, c = i & 63;
    const uint32_t hl = split_hi_lo(__ldg(p.wgt + i));
    const int n0 = t, n1 = 16 + t;
    *reinterpret_cast<unsigned short*>(s_w + (n0 >> 3) * 1024 + (n0 & 7) * 128 + (((c >> 3) ^ (n0 & 7)) << 4) + (c & 7) * 2) =
        static_cast<unsigned short>(hl & 0xffffu);
    *reinterpret_cast<unsigned short*>(s_w + (n1 >> 3) * 1024 + (n1 & 7) * 128 + (((c >> 3) ^ (n1 & 7)) << 4) + (c & 7) * 2) =
        static_cast<unsigned short>(hl >> 16);
  }
  if (tid == 0) {
    for (int i = 0; i < kTfStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tm_x);
  }
  if (warp == 2) tmem_alloc<128>(tmem_slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const unsigned per_img = static_cast<unsigned>(p.tiles_h) * p.tiles_w;
  const unsigned tiles = static_cast<unsigned>(p.B) * per_img;
  const unsigned my_tiles = blockIdx.x < tiles ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;

  if (warp == 0) {
    if (lane == 0) {
      for (unsigned i = 0; i < my_tiles; ++i) {
        const unsigned tile = blockIdx.x + i * gridDim.x;
        const int b = static_cast<int>(tile / per_img), r = static_cast<int>(tile % per_img);
        const int ty = r / p.tiles_w, tx = r % p.tiles_w;
        const int stage = static_cast<int>(i % kTfStages);
        mbar_wait(&empty[stage], ((i / kTfStages) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&full[stage], 32768);
        tma_load_4d(s_x + stage * 32768, &tm_x, &full[stage], 0, tx * kTfIn - 1, ty * kTfIn - 1, b);
      }
    }
  } else if (warp == 1) {
    {   // whole warp (warp-uniform addressing); one elected lane issues
      constexpr uint32_t idesc = make_idesc_bf16(128, 32, false, false);
      const uint64_t dx0 = make_smem_desc(smem_u32(s_x), 16, 1024), dw0 = make_smem_desc(smem_u32(s_w), 16, 1024);
      const uint32_t x_lo0 = static_cast<uint32_t>(dx0), w_lo0 = static_cast<uint32_t>(dw0), d_hi = static_cast<uint32_t>(dx0 >> 32);
      for (unsigned i = 0; i < my_tiles; ++i) {
        const int stage = static_cast<int>(i % kTfStages);
        const int acc = static_cast<int>(i & 1u);
        mbar_wait(&tempty[acc], ((i >> 1) & 1u) ^ 1u);
        mbar_wait(&full[stage], (i / kTfStages) & 1u);
        tc_fence_after();
        const uint32_t x_lo = x_lo0 + static_cast<uint32_t>(stage) * (32768 >> 4);
        if (elect_one()) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_bf16_lh(tmem + acc * 64 + half * 32, x_lo + half * (16384 >> 4) + ks * 2, d_hi, w_lo0 + ks * 2, d_hi, idesc,
                           ks ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          umma_commit(&tfull[acc]);
        }
      }
    }
  } else if (warp >= 4) {
    // two groups of four warps take alternate tiles (group g owns accumulator g), so that one tile's TMEM read, shared-memory
    // exchange and global epilogue traffic overlap the other's; thread = patch pixels e0 (rows 0-7) and e0 + 128 (rows 8-15)
    const int g = (warp - 4) >> 2, q = warp & 3;   // TMEM lane quarter = warp % 4
    const int e0 = q * 32 + lane;
    const float bv = p.bias ? __ldg(p.bias) : 0.f;
    int off[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) off[t] = t < p.ntaps ? t * 256 + p.dh[t] * kTfPatch + p.dw[t] : 0;
    float* const st_base = s_t + g * (2 * 16 * 256);
    unsigned n = 0;
    for (unsigned i = g; i < my_tiles; i += 2, ++n) {
      const unsigned tile = blockIdx.x + i * gridDim.x;
      const int b = static_cast<int>(tile / per_img), r = static_cast<int>(tile % per_img);
      const int ty = r / p.tiles_w, tx = r % p.tiles_w;
      float* st = st_base + (n & 1u) * (16 * 256);
      // output pixels of this thread and their composite operands, requested before the accumulator is waited for
      bool live[2];
      size_t o[2];
      uint32_t mk[2] = {0u, 0u};       // raw loads only: nothing may consume them before the gather (they stay in flight)
      float xi[2] = {0.f, 0.f};
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int e = e0 + h * 128, ph = e >> 4, pw = e & 15;
        const int oh = ty * kTfIn + ph - 1, ow = tx * kTfIn + pw - 1;
        live[h] = ph >= 1 && ph <= kTfIn && pw >= 1 && pw <= kTfIn && oh < p.H && ow < p.W;
        o[h] = (static_cast<size_t>(b) * p.H + (live[h] ? oh : 0)) * p.W + (live[h] ? ow : 0);
        if (p.mode != 0 && live[h]) {
          mk[h] = __ldg(p.mask + o[h]);
          xi[h] = __ldg(p.xin + o[h]);
        }
      }
      mbar_wait(&tfull[g], n & 1u);
      tc_fence_after();
      uint32_t raw0[32], raw1[32];
      tmem_ld_32x32(tmem + (static_cast<uint32_t>(q * 32) << 16) + g * 64, raw0);
      tmem_ld_32x32(tmem + (static_cast<uint32_t>(q * 32) << 16) + g * 64 + 32, raw1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[g]);
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        if (t < p.ntaps) {
          st[t * 256 + e0] = __uint_as_float(raw0[t]) + __uint_as_float(raw0[16 + t]);
          st[t * 256 + e0 + 128] = __uint_as_float(raw1[t]) + __uint_as_float(raw1[16 + t]);
        }
      }
      if (g == 0) asm volatile("bar.sync 1, 128;" ::: "memory");      // the tile's tap products are all in shared memory
      else asm volatile("bar.sync 2, 128;" ::: "memory");
      // pin the first use of the prefetched composite operands below the barrier (the compiler otherwise hoists the mask
      // test above the accumulator wait and stalls there on the load: ncu, 18 % of the samples)
      asm volatile("" : "+r"(mk[0]), "+r"(mk[1]), "+f"(xi[0]), "+f"(xi[1]));
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (!live[h]) continue;
        const int e = e0 + h * 128;
        float a = 0.f;
#pragma unroll
        for (int t = 0; t < 16; ++t)
          if (t < p.ntaps) a += st[off[t] + e];
        const float v = a + bv;
        if (p.mode == 0) {
          p.out[o[h]] = v;
        } else {
          const float sg = 1.f / (1.f + __expf(-v));
          if (p.sig_out) p.sig_out[o[h]] = sg;
          const float m = mk[h] ? 1.f : 0.f;
          p.out[o[h]] = sg * (1.f - m) + xi[h] * m;
        }
      }
      // no second barrier: the group's other buffer is used next, and nobody reaches this one again before the next bar.sync
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<128>(tmem);
}

static bool to1_fused_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("TG_NO_TO1_FUSED");
    on = (e && e[0] == '1') ? 0 : 1;
  }
  return on == 1;
}

bool to1_fused_covers(int x_split, int H, int W, const To1Taps& taps, int ntaps, int Ho, int Wo) {
  if (!thin_mma_enabled() || !to1_fused_enabled() || x_split || Ho != H || Wo != W) return false;
  if (taps.ncls != 1 || ntaps < 1 || ntaps > 16 || H < kTfPatch || W < kTfPatch) return false;
  for (int t = 0; t < ntaps; ++t) {
    const int dh = taps.dh[taps.begin[0] + t], dw = taps.dw[taps.begin[0] + t];
    if (dh < -1 || dh > 1 || dw < -1 || dw > 1) return false;
  }
  return true;
}

// returns 0 when launched, -1 when the shape is not covered
static int to1_fused_launch(const void* x, int B, int H, int W, const float* wgt, const To1Taps& taps, int ntaps, const float* bias,
                            int mode, const uint8_t* mask, const float* xin, float* out, float* sig_out, cudaStream_t st) {
  if (!to1_fused_covers(0, H, W, taps, ntaps, H, W)) return -1;
  To1FusedParams p{};
  for (int t = 0; t < ntaps; ++t) {
    p.dh[t] = taps.dh[taps.begin[0] + t];
    p.dw[t] = taps.dw[taps.begin[0] + t];
  }
  p.B = B; p.H = H; p.W = W; p.ntaps = ntaps;
  p.tiles_h = (H + kTfIn - 1) / kTfIn;
  p.tiles_w = (W + kTfIn - 1) / kTfIn;
  p.wgt = wgt; p.bias = bias; p.mode = mode; p.mask = mask; p.xin = xin; p.out = out; p.sig_out = sig_out;
  CUtensorMap tm_x;
  const uint64_t dims[4] = {64, static_cast<uint64_t>(W), static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
  const uint64_t str[3] = {128, static_cast<uint64_t>(W) * 128, static_cast<uint64_t>(W) * H * 128};
  const uint32_t box[4] = {64, kTfPatch, kTfPatch, 1};
  if (make_tmap_bf16(&tm_x, x, 4, dims, str, box) != 0) return -3;
  TG_SET_SMEM_ONCE((to1_fused_kernel), kTfSmem);
  const long tiles = static_cast<long>(B) * p.tiles_h * p.tiles_w;
  const int grid = static_cast<int>(tiles < num_sms() ? tiles : num_sms());
  to1_fused_kernel<<<grid, 384, kTfSmem, st>>>(tm_x, p);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int to1_fwd_mma(const void* x, int x_split, int B, int H, int W, const float* wgt, const To1Taps& taps, int ntaps,
                const float* bias, int Ho, int Wo, int mode, const uint8_t* mask, const float* xin, float* out,
                float* sig_out, float* scratch, size_t scratch_floats, cudaStream_t st) {
  const long total = static_cast<long>(B) * H * W;
  if (ntaps < 1 || ntaps > 32 || total >= (1L << 31) || static_cast<long>(B) * Ho * Wo >= (1L << 31)) return -1;
  if (to1_fused_covers(x_split, H, W, taps, ntaps, Ho, Wo)) {
    const int rc = to1_fused_launch(x, B, H, W, wgt, taps, ntaps, bias, mode, mask, xin, out, sig_out, st);
    if (rc != -1) return rc;
  }
  if (scratch == nullptr || scratch_floats < static_cast<size_t>(ntaps) * total) return -1;
  if (x_split && (H % 2 || W % 2)) return -1;
  if (total % 4 != 0) return -1;                       // 16-byte stores into the rows of T
  TG_SET_SMEM_ONCE((tapdot_kernel), kTdSmem);
  CUtensorMap tm_x;
  const uint64_t dims[2] = {64, static_cast<uint64_t>(total)};
  const uint64_t str[1] = {128};
  const uint32_t box[2] = {64, kTdN};
  if (make_tmap_bf16(&tm_x, x, 2, dims, str, box) != 0) return -3;
  const long tiles = (total + kTdN - 1) / kTdN;
  const int grid = static_cast<int>(tiles < num_sms() ? tiles : num_sms());
  tapdot_kernel<<<grid, 384, kTdSmem, st>>>(tm_x, wgt, ntaps, static_cast<unsigned>(total), scratch);
  TG_CHECK_CUDA(cudaGetLastError());
  const long M = static_cast<long>(B) * Ho * Wo;
  long g2 = (M + 255) / 256;
  if (g2 > 16L * num_sms()) g2 = 16L * num_sms();
  tapsum_kernel<<<static_cast<int>(g2), 256, 0, st>>>(scratch, static_cast<unsigned>(total), x_split, B, H, W, taps, bias,
                                                     Ho, Wo, mode, mask, xin, out, sig_out);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// step 2 alone, for callers that produced T themselves (direct_conv.cu: the 512 -> 1 convolution of D model[11])
int to1_tapsum_launch(const float* T, long total_in, int x_split, int B, int H, int W, const To1Taps& taps,
                      const float* bias, int Ho, int Wo, int mode, const uint8_t* mask, const float* xin, float* out,
                      float* sig_out, cudaStream_t st) {
  const long M = static_cast<long>(B) * Ho * Wo;
  if (total_in >= (1L << 31) || M >= (1L << 31)) return -1;
  long g2 = (M + 255) / 256;
  if (g2 > 16L * num_sms()) g2 = 16L * num_sms();
  tapsum_kernel<<<static_cast<int>(g2), 256, 0, st>>>(T, static_cast<unsigned>(total_in), x_split, B, H, W, taps, bias, Ho, Wo,
                                                     mode, mask, xin, out, sig_out);
  TG_CHECK_CUDA(cudaGetLastError());
  return 0;
}

bool thin_mma_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("TG_NO_THIN_MMA");
    on = (e && e[0] == '1') ? 0 : 1;
  }
  return on == 1;
}

// entry used by direct_conv.cu
int rowgemm_dispatch(int k, const RowGemmParams& p, int grid_cap, int* grid_used, cudaStream_t st) {
  const bool m = p.src_mask != nullptr;
  if (k == 3) return m ? rowgemm_launch<3, true>(p, grid_cap, grid_used, st) : rowgemm_launch<3, false>(p, grid_cap, grid_used, st);
  if (k == 4) return m ? rowgemm_launch<4, true>(p, grid_cap, grid_used, st) : rowgemm_launch<4, false>(p, grid_cap, grid_used, st);
  if (k == 7) return m ? rowgemm_launch<7, true>(p, grid_cap, grid_used, st) : rowgemm_launch<7, false>(p, grid_cap, grid_used, st);
  return -1;
}

}  // namespace tg
