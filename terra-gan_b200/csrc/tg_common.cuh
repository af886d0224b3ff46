// tg_common.cuh — sm_100a device-side primitives shared by the TERRA-GAN B200 kernels.
//
// Thin inline-PTX wrappers for the Blackwell execution model: mbarrier, TMA
// (cp.async.bulk.tensor), tcgen05 (TMEM alloc / mma / commit / ld) plus the
// shared-memory matrix descriptor and instruction descriptor encoders used by
// the implicit-GEMM convolution kernels (conv_igemm.cu, wgrad_igemm.cu).
//
// Nothing here exists in the reference (it ships no native code);
// the reference ops these kernels replace are cited at each kernel.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tg {

// ----------------------------------------------------------------------------------------------
// error plumbing (host)
// ----------------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);

#define TG_CHECK_CUDA(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      tg::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                   \
                         cudaGetErrorString(_e));                                        \
      return -2;                                                                         \
    }                                                                                    \
  } while (0)

#define TG_REQUIRE(cond, ...)                                                            \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      tg::set_last_error(__VA_ARGS__);                                                   \
      return -1;                                                                         \
    }                                                                                    \
  } while (0)

int num_sms();  // cached SM count of the current device

// Opt a kernel into > 48 KB of dynamic shared memory once per DEVICE (the attribute is per context; a
// process-wide flag was wrong for a process that drives more than one GPU).
#define TG_SET_SMEM_ONCE(kernel, bytes)                                                            \
  do {                                                                                             \
    static bool _done[64] = {};                                                                    \
    int _dev = 0;                                                                                  \
    TG_CHECK_CUDA(cudaGetDevice(&_dev));                                                           \
    if (!_done[_dev & 63]) {                                                                       \
      TG_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                         static_cast<int>(bytes)));                                \
      _done[_dev & 63] = true;                                                                     \
    }                                                                                              \
  } while (0)

// Grid of a grid-stride kernel as WHOLE waves: `resident` = CTAs of this kernel that fit on one SM (occupancy API, asked
// once per kernel and dynamic shared-memory size), grid = min(blocks_needed, resident * SMs). Every CTA is resident from
// the start and does the same share of the work. The fixed 8 x SMs grids this replaces ran 3 + 3 + 2 (or 6 + 2) CTAs
// per SM in turn for kernels with 3 (6) resident CTAs: a last, partly filled wave at a fraction of the HBM bandwidth.
// TG_WAVE_GRID=0 restores the old sizing (A/B timing). Returns <= 0 on error.
int wave_grid_lookup(const void* kernel, int block, size_t smem, int* resident);   // tg_runtime.cu (cache)
void wave_grid_store(const void* kernel, int block, size_t smem, int resident);
bool wave_grid_enabled();
template <typename K>
inline int wave_grid(K kernel, int block, size_t smem, long blocks_needed, int legacy_per_sm = 8) {
  const long sms = num_sms() > 0 ? num_sms() : 148;
  long per_sm = legacy_per_sm;
  if (wave_grid_enabled()) {
    int resident = 0;
    if (!wave_grid_lookup(reinterpret_cast<const void*>(kernel), block, smem, &resident)) {
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kernel, block, smem) != cudaSuccess || resident < 1)
        resident = legacy_per_sm;
      wave_grid_store(reinterpret_cast<const void*>(kernel), block, smem, resident);
    }
    per_sm = resident;
  }
  long g = blocks_needed < per_sm * sms ? blocks_needed : per_sm * sms;
  return static_cast<int>(g < 1 ? 1 : g);
}

// Perf-experiment switches (skip stores / loads / MMAs) used to live behind run-time `p.debug` bits that every
// epilogue tile evaluated; they are compiled out unless the library is built with -DTG_PERF_DEBUG.
#ifdef TG_PERF_DEBUG
#define TG_DBG(p, bit) (((p).debug & (bit)) != 0)
#else
#define TG_DBG(p, bit) false
#endif

// Encode a bf16 tiled tensor map (SWIZZLE_128B, zero OOB fill). dims/strides innermost first;
// strides_bytes has rank-1 entries (dims 1..rank-1). Returns 0 on success.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box);
// Same for fp32 elements (the TF32 verification path: fp32 storage, kind::tf32 MMAs).
int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box);
// fp32 with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B (32-byte chunks swizzled over 4 rows): the only shared-memory layout
// MN-major kind::tf32 operands can use (UMMA layout type SWIZZLE_128B_BASE32B) -- the fp32 weight-gradient kernel.
int make_tmap_f32_atom32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box);

// ----------------------------------------------------------------------------------------------
// device helpers
// ----------------------------------------------------------------------------------------------
#if defined(__CUDACC__)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() {
  uint32_t l;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
  return l;
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---- TMA ----
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
      "r"(c4)
      : "memory");
}

// shared -> global tensor store (bulk async group); the smem tile is in the tensor map's swizzle layout
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk groups of this thread have finished READING their shared-memory source
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// ---- tcgen05 / TMEM ----
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_slot)),
               "n"(kCols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(kCols));
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16 inputs with fp32 accumulate.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same, with the two shared-memory descriptors given as (lo, hi) 32-bit words. The issuing thread is a single
// dependent instruction stream: rebuilding 64-bit descriptors per MMA costs more clocks than an N <= 128 MMA
// takes (measured with tools/probe/mma_rate.cu: 78 clk/MMA with per-MMA descriptor arithmetic, 55 / 64 / 128 clk
// for N = 64 / 128 / 256 with precomputed words). Kernels therefore precompute the hi words and add byte offsets
// (>> 4) to the lo word, whose low 14 bits hold the start address.
__device__ __forceinline__ void umma_bf16_lh(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::tf32: fp32 words in shared memory, the tensor core reads the upper 19 bits (TF32), fp32 accumulate; K = 8
// per instruction (32 bytes along K, as for kind::f16). The verification path (fp32 storage) uses it.
__device__ __forceinline__ void umma_tf32_lh(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
          smem_u32(bar))
      : "memory");
}
// ---- CTA pair (cluster of 2, tcgen05 cta_group::2) ----
// One thread of the even CTA issues the MMAs for both SMs (M = 256: each CTA's 128 accumulator lanes live in its own
// TMEM, each CTA supplies its own A rows and half of the B rows from its own shared memory at the descriptor's offsets).
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {      // every thread of every CTA of the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_slot) {      // one warp of EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "n"(kCols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t addr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(kCols));
}
// TMA loads issued by either CTA of a pair that complete on an mbarrier given as a shared::cluster address (the leader's)
__device__ __forceinline__ void tma_load_2d_2cta(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d_2cta(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1,
                                                 int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_lh_2cta(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                  uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier at this shared-memory offset in BOTH CTAs of the pair once all prior MMAs of this thread completed
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives row (lane base + i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- descriptors ----
// Shared-memory matrix descriptor, SWIZZLE_128B (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=2.
// K-major tile (rows = M/N index, 128 B = 64 bf16 of K per row): SBO = 1024 (8-row atom), LBO unused.
// MN-major tile (rows = K index, 128 B = 64 bf16 of M/N per row): SBO = 1024 (8 k-rows), LBO =
// byte distance between consecutive 64-wide M/N blocks.
// layout_type: 2 = SWIZZLE_128B (16-byte chunks over 8 rows); 1 = SWIZZLE_128B_BASE32B (32-byte chunks over 4 rows:
// MN-major tf32 operands; SBO is then the distance between 4-row groups).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes, uint32_t layout_type = 2) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout_type) << 61;
  return d;
}
// Instruction descriptor for kind::f16, bf16 x bf16 -> fp32 (cute::UMMA::InstrDescriptor layout).
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t m, uint32_t n, bool a_mn_major,
                                                       bool b_mn_major) {
  return (1u << 4)                         // c_format = F32
         | (1u << 7)                       // a_format = BF16
         | (1u << 10)                      // b_format = BF16
         | ((a_mn_major ? 1u : 0u) << 15)  // a_major
         | ((b_mn_major ? 1u : 0u) << 16)  // b_major
         | ((n >> 3) << 17)                // n_dim
         | ((m >> 4) << 24);               // m_dim
}

// Same for kind::tf32 (a_format = b_format = 2).
__host__ __device__ constexpr uint32_t make_idesc_tf32(uint32_t m, uint32_t n, bool a_mn_major,
                                                       bool b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((n >> 3) << 17) | ((m >> 4) << 24);
}

// ---- 8-element vectors of the activation storage type (bf16: one 16-byte word; fp32: two) ----
// bf16 -> fp32 is exact and is just the 16 bits moved up: one shift for the low half, one AND for the high half of a
// packed pair (__bfloat1622float2 compiles to PRMT + shift for the high half: three instead of two operations per pair in
// kernels that are bound by instruction issue)
__device__ __forceinline__ void unpack_bf16x2(uint32_t w, float& lo, float& hi) {
  lo = __uint_as_float(w << 16);
  hi = __uint_as_float(w & 0xffff0000u);
}
__device__ __forceinline__ void vload8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 raw = *reinterpret_cast<const uint4*>(p);
  unpack_bf16x2(raw.x, f[0], f[1]);
  unpack_bf16x2(raw.y, f[2], f[3]);
  unpack_bf16x2(raw.z, f[4], f[5]);
  unpack_bf16x2(raw.w, f[6], f[7]);
}
__device__ __forceinline__ void vload8(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void vstore8(__nv_bfloat16* p, const float (&f)[8]) {
  __nv_bfloat162 v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(v);
}
__device__ __forceinline__ void vstore8(float* p, const float (&f)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *(reinterpret_cast<float4*>(p) + 1) = make_float4(f[4], f[5], f[6], f[7]);
}

// ---- misc math ----
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Transposing butterfly reduction: every lane holds 32 values v[0..31] (one per column); after the
// call lane j holds the sum over all 32 lanes of column j. 31 shuffles instead of 32*5.
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32]) {
  const uint32_t lane = lane_id();
#pragma unroll
  for (int step = 0; step < 5; ++step) {
    const int half = 16 >> step;  // 16, 8, 4, 2, 1
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      // lanes with the bit set keep the upper half of the columns, the others the lower half
      float send = upper ? v[i] : v[i + half];
      float keep = upper ? v[i + half] : v[i];
      float recv = __shfl_xor_sync(0xffffffffu, send, half);
      v[i] = keep + recv;
    }
  }
  return v[0];
}

#endif  // __CUDACC__

}  // namespace tg
