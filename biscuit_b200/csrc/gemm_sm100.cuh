// Dense bf16 GEMM with fused epilogue on the 5th-generation tensor cores (tcgen05 / TMEM / TMA).
//
//   D[M,N] = epilogue( A[M,K] * B[N,K]^T ),   A, B bf16 K-major, fp32 accumulation in TMEM
//   epilogue: v = acc * scale[n] + shift[n] (+ residual[m,n]) -> optional ReLU -> bf16
//
// It carries every GEMM-shaped op of the Xception-UQ hot path (K2, K3 pointwise, K4, and the
// dense layers of the MC-dropout head):
//   * pointwise 1x1 convolutions: A = NHWC activations viewed as [B*H*W, Cin], B = weights [Cout, Cin];
//     BatchNorm is the per-channel scale/shift, ReLU and the residual add ride in the epilogue;
//   * block1_conv2 (3x3 valid, 32->64) as an implicit GEMM: K-block kb is filter tap (kb/3, kb%3) and the A
//     tile of that tap is the SAME 2-D activation matrix shifted by (ky*W + kx) rows, so one TMA descriptor
//     serves all nine taps; rows whose (y, x) fall outside the valid output window are dropped in the epilogue;
//   * dense layers: scale = 1 (or 1/(1-p) after a dropout site), shift = bias.
//
// Structure (one CTA per SM, CTA pairs on a TPC, persistent over output tiles, warp specialised):
//   warp 0   : TMA producer  -- cp.async.bulk.tensor.2d into a STAGES-deep smem ring (128B swizzle)
//   warp 1   : TMEM allocator + tcgen05.mma issue (one elected lane of the converged warp), accumulators double-buffered
//              in TMEM (2 x 256 columns)
//   warps 2-9: epilogue      -- tcgen05.ld 32x32b.x32 -> registers -> scale/shift/residual/ReLU -> bf16 -> smem -> TMA store
// Three mbarrier pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace bq {

struct GemmParams {
  int M = 0;               // rows of D (conv mode: virtual rows over the INPUT grid)
  int N = 0;               // output channels
  int K = 0;               // reduction length (conv mode: 9 * Cin)
  int bn_box = 256;        // rows of the B TMA box == N-tile width (multiple of 16, <= 256)
  const float* scale = nullptr;   // [round_up(N,256)] per-channel scale, nullptr -> alpha
  const float* shift = nullptr;   // [round_up(N,256)] per-channel shift, nullptr -> 0
  float alpha = 1.0f;
  const __nv_bfloat16* residual = nullptr;   // [M, ldr] added before the ReLU
  int ldr = 0;
  __nv_bfloat16* out = nullptr;              // [M, ldc]
  int ldc = 0;
  int out_row_off = 0;     // TMA-store epilogues: rows are written at (row + out_row_off) of the output tensor map
  int relu = 0;
  // implicit 3x3 valid convolution (conv_mode = 1): virtual row m = img*in_hw + y*in_w + x
  int conv_mode = 0;
  int in_w = 0, in_hw = 0, out_w = 0, out_h = 0, out_hw = 0;
  // debug path only (SIMT kernel reads the operands directly)
  const __nv_bfloat16* a_ptr = nullptr;
  int lda = 0;
  long long a_rows = 0;
  const __nv_bfloat16* b_ptr = nullptr;
  int ldb = 0;
};

namespace sm100 {

constexpr int kBM = 128;        // UMMA M (cta_group::1)
constexpr int kBNMax = 256;     // UMMA N max
constexpr int kThreads = 192;   // 6 warps
constexpr int kAccStages = 2;

// Wait used by roles that idle for a long time (epilogue warps waiting for a whole item's accumulation): after the first
// hardware-suspended try the retry loop backs off with nanosleep.  ncu: without it the idle epilogue warps of the fused
// middle-flow kernel re-polled every ~60 cycles and issued 22 % of ALL instructions of the kernel.
__device__ __forceinline__ void mbar_wait_idle(uint32_t bar, uint32_t parity, uint32_t sleep_ns);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// try_wait with a suspend-time hint: the thread is parked by the hardware until the phase completes (or the hint
// expires) instead of spinning -- ncu showed ~40 % of all issued instructions of the fused kernels were barrier polling
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must abort the kernel (trap -> launch failure), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > 4000000u) {        // each failed try_wait parks the thread for up to the hint: this is seconds
      printf("bq: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}

__device__ __forceinline__ void mbar_wait_idle(uint32_t bar, uint32_t parity, uint32_t sleep_ns) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(sleep_ns);
    if (++spins > 40000000u) {
      printf("bq: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}

__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"((uint64_t)tmap), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)tmap) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t smem_slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_slot), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// One lane of a CONVERGED warp.  tcgen05.mma / tcgen05.commit take uniform-register operands: when they are issued from
// inside an `if (lane == 0)` region every operand lives in a vector register and ptxas wraps each instruction in an
// R2UR / ELECT / BRA.U.ANY loop (~25 instructions per MMA; ncu showed the single issuing thread, not L2, pacing the
// GEMMs).  Issuing from warp-uniform control flow under elect.sync keeps descriptors and loop counters in uniform
// registers: ~3 instructions per MMA.
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p));
  return p != 0;
}
// warp index the compiler can prove warp-uniform (the shuffle is the hint, as in cutlass::canonical_warp_idx_sync)
__device__ __forceinline__ int uniform_warp_idx() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): rows of SWIZZLE bytes, 8-row groups.
//   bits [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major) | [32,46) SBO>>4 | [46,48) version=1
//   | [61,64) layout (2 = SWIZZLE_128B, 4 = SWIZZLE_64B)
// ---- accumulator-fragment epilogue helpers (fused sepconv kernels: TMEM lane = output channel) ----
// 16 TMEM lanes x 256 bits: thread t gets (lane t/4, columns 2(t%4), 2(t%4)+1) in r0, r1 and (lane t/4 + 8, same columns) in
// r2, r3 -- the m16n8 accumulator-fragment layout, repeated along the columns for .x4 (registers 4i.. = columns 8i..).
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_16x256b_x1(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr)
               : "memory");
}
// four 8x8 b16 matrices, transposed on the way to / from shared memory: thread i supplies the address of row i % 8 of
// matrix i / 8; register k of thread t holds elements (row t/4, columns 2(t%4), 2(t%4)+1) of matrix k BEFORE the transpose
__device__ __forceinline__ void stmatrix_x4_trans(uint32_t addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
               : "memory");
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr)
               : "memory");
}
// lo += bf16(pair.lo), hi += bf16(pair.hi) in fp32, round to nearest: the mixed-precision add (FHADD.BF16 with an .H0 / .H1
// operand selector) reads the packed halves directly -- one instruction per element instead of unpack + add
__device__ __forceinline__ void add_bf16x2_to_f32(float& lo, float& hi, uint32_t pair) {
  asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tadd.rn.f32.bf16 %0, l, %0;\n\tadd.rn.f32.bf16 %1, h, %1;\n\t}"
      : "+f"(lo), "+f"(hi)
      : "r"(pair));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi, bool relu) {
  uint32_t r;
  if (relu) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

template <int SWIZZLE_BYTES>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  constexpr uint64_t layout = SWIZZLE_BYTES == 128 ? 2ull : (SWIZZLE_BYTES == 64 ? 4ull : 6ull);
  constexpr uint64_t sbo = (8 * SWIZZLE_BYTES) >> 4;
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}

// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, both K-major
__device__ __forceinline__ uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// TMA store: smem tile -> global (bulk async group)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tmap, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"((uint64_t)tmap), "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tmap, uint32_t smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"((uint64_t)tmap), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// all but the most recent bulk store have finished reading their smem source (double-buffered staging)
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

constexpr int kEpiCols = 64;                       // epilogue chunk: 128 rows x 64 bf16 columns = 16 KB
constexpr int kEpiBytes = kBM * kEpiCols * 2;

// =====================================================================================================
// 2-CTA variant (cta_group::2): a CTA pair on one TPC computes a 256 x N tile with ONE tcgen05.mma stream.
// Each CTA stages its own 128 rows of A and HALF of the B tile (N/2 rows), so operand traffic from L2 per
// CTA per k-block drops from 48 KB to 32 KB (the 1-CTA kernel is L2->SM bound: ncu shows 25 % tensor-pipe
// activity with neither L2 nor HBM saturated).  Protocol (after DeepGEMM's sm100 kernels):
//   full[s]   lives in the LEADER (cluster rank 0): 2 arrivals (leader expect_tx for both CTAs' bytes, peer
//             remote arrive); both CTAs' TMA loads complete_tx on the leader's barrier (.cta_group::2 form)
//   empty[s], tmem_full[a]  live in BOTH CTAs, signalled by multicast tcgen05.commit (mask 0b11)
//   tmem_empty[a] lives in the leader: 8 arrivals (4 epilogue warps x 2 CTAs, remote via mapa)
// Only the leader's warp 1 issues MMAs; each CTA runs its own TMA producer and its own epilogue on its
// 128 accumulator rows.
// =====================================================================================================
constexpr int k2Stages = 4;                          // operand ring depth with a residual operand; 5 without (smem permitting)
constexpr int k2EpiWarps = 8;                        // two warps per TMEM lane quadrant (32 columns of a chunk each)
constexpr int k2Threads = 64 + 32 * k2EpiWarps;      // warp 0 TMA, warp 1 MMA/TMEM, warps 2..9 epilogue
constexpr int k2MaxN = 2048;
constexpr int k2ResBufs = 3;                         // residual chunks in flight (prefetch distance 2)                         // per-channel scale/shift staged in smem for the whole N
template <int STAGES>
struct SmemPlan2 {
  static constexpr int kABytes = kBM * 64 * 2;                 // 16 KB
  static constexpr int kBBytes = (kBNMax / 2) * 64 * 2;        // 16 KB (this CTA's half of the first 256 columns)
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kOutOffset = STAGES * kStageBytes;
  static constexpr int kScaleOffset = kOutOffset + 2 * kEpiBytes;           // float scale[k2MaxN], shift[k2MaxN]
  static constexpr int kBarOffset = kScaleOffset + 2 * k2MaxN * 4;
  static constexpr int kResOffset = kBarOffset + 1024;                      // residual staging LAST: only requested when used
  static constexpr int kTotalNoRes = kResOffset + 1024;                     // 165 KB: leaves room for a co-resident depthwise block
  static constexpr int kTotal = kResOffset + k2ResBufs * kEpiBytes + 1024;  // + 1 KB slack for the 1024 B alignment
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(cta)
      : "memory");
}
// TMA load whose completion bytes are credited to the LEADER CTA's mbarrier (peer bit cleared)
__device__ __forceinline__ void tma_load_2d_2cta(uint32_t smem_dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"((uint64_t)tmap), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t smem_slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_slot), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}

template <int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(k2Threads, 1)
gemm_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_out,
                         const __grid_constant__ CUtensorMap tmap_res, const GemmParams p) {
  using Plan = SmemPlan2<STAGES>;
  constexpr int BLOCK_K = 64;
  constexpr int ACC = kAccStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + Plan::kBarOffset;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + kAccStages + a); };
  auto res_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 2 * kAccStages + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 2 * kAccStages + k2ResBufs);
  volatile uint32_t* tmem_slot_ptr =
      (volatile uint32_t*)(smem_gen + Plan::kBarOffset + 8 * (2 * STAGES + 2 * kAccStages + k2ResBufs));

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool is_leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int m_tiles = (p.M + 2 * kBM - 1) / (2 * kBM);              // 256-row pair tiles
  const int n_tiles = (p.N + p.bn_box - 1) / p.bn_box;
  const int total_tiles = m_tiles * n_tiles;
  const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
  const uint32_t b_box_bytes = (uint32_t)(p.bn_box / 2) * BLOCK_K * 2;
  const uint32_t stage_tx = 2u * (Plan::kABytes + b_box_bytes);     // both CTAs' bytes land on the leader's barrier

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_out);
    if (p.residual) tma_prefetch_desc(&tmap_res);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < kAccStages; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 2 * k2EpiWarps); }
    for (int b = 0; b < k2ResBufs; ++b) mbar_init(res_bar(b), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // per-channel epilogue constants -> smem (read back as broadcast LDS; global loads here stalled the epilogue)
  float* s_scale = (float*)(smem_gen + Plan::kScaleOffset);
  float* s_shift = s_scale + k2MaxN;
  for (int i = threadIdx.x; i < p.N; i += blockDim.x) {
    s_scale[i] = p.scale ? __ldg(p.scale + i) : p.alpha;
    s_shift[i] = p.shift ? __ldg(p.shift + i) : 0.f;
  }
  cluster_sync_all();                        // both CTAs resident before the paired TMEM allocation
  if (warp == 1) tmem_alloc_2cta(tmem_slot, 512);
  tc_fence_before();
  cluster_sync_all();                        // barriers initialised + TMEM allocated in both CTAs
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);   // warp-uniform for the compiler (MMA operands live in uniform registers)

  if (warp == 0) {
    // ===================== TMA producer (each CTA: its A rows + its half of B) =====================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        const int m0 = (tile / n_tiles) * (2 * kBM) + (int)rank * kBM;
        const int n0 = (tile % n_tiles) * p.bn_box;
        int n_cols = p.N - n0;
        if (n_cols > p.bn_box) n_cols = p.bn_box;
        n_cols = (n_cols + 15) & ~15;
        const int nb0 = n0 + (int)rank * (n_cols / 2);          // this CTA's half of the N tile
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          if (is_leader) mbar_expect_tx(full_bar(s), stage_tx);
          else mbar_arrive_remote(full_bar(s), 0);
          const uint32_t a_dst = smem_base + s * Plan::kStageBytes, b_dst = a_dst + Plan::kABytes;
          tma_load_2d_2cta(a_dst, &tmap_a, full_bar(s), kb * BLOCK_K, m0);
          tma_load_2d_2cta(b_dst, &tmap_b, full_bar(s), kb * BLOCK_K, nb0);
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: leader CTA; the whole warp walks the loop, one elected lane issues =====================
    if (is_leader) {
      int s = 0; uint32_t ph = 0;
      int as = 0; uint32_t aph = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        const int n0 = (tile % n_tiles) * p.bn_box;
        int n_cols = p.N - n0;
        if (n_cols > p.bn_box) n_cols = p.bn_box;
        n_cols = (n_cols + 15) & ~15;
        const uint32_t idesc = make_idesc(2 * kBM, n_cols);
        mbar_wait(tempty_bar(as), aph ^ 1u);
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * kBNMax);
#pragma unroll 1
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t a_src = smem_base + s * Plan::kStageBytes, b_src = a_src + Plan::kABytes;
          int ksteps = BLOCK_K / 16;
          if (kb == num_kb - 1) ksteps = (p.K - kb * BLOCK_K + 15) / 16;
          const uint64_t da = make_smem_desc<128>(a_src), db = make_smem_desc<128>(b_src);
          if (elect_one()) {
            umma_bf16_2cta(d_tmem, da, db, idesc, kb ? 1u : 0u);
            if (ksteps > 1) {
              umma_bf16_2cta(d_tmem, da + 2u, db + 2u, idesc, 1u);
            }
            if (ksteps > 2) {
              umma_bf16_2cta(d_tmem, da + 4u, db + 4u, idesc, 1u);
            }
            if (ksteps > 3) {
              umma_bf16_2cta(d_tmem, da + 6u, db + 6u, idesc, 1u);
            }
            umma_commit_2cta(empty_bar(s));
            if (kb == num_kb - 1) umma_commit_2cta(tfull_bar(as));
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        if (++as == ACC) { as = 0; aph ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue: each CTA drains its own 128 accumulator rows =====================
    // 8 warps: warp -> (TMEM lane quadrant = warp % 4, column half = (warp - 2) / 4); a chunk is 128 rows x 64 cols.
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row_in_tile = quad * 32 + lane;
    int as = 0; uint32_t aph = 0;
    const bool leader = (warp == 2 && lane == 0);
    const bool has_res = p.residual != nullptr;
    uint32_t cc = 0;
    // residual chunks are prefetched k2ResBufs-1 epilogue steps ahead (a chunk = 128 rows x 64 columns, in the order
    // the epilogue consumes them); `pf_*` is the prefetch cursor of the leader thread
    int pf_tile = cluster_id, pf_c = 0;
    uint32_t pf_cc = 0;
    auto prefetch_one = [&]() {
      if (pf_tile >= total_tiles) return;
      const int pn0 = (pf_tile % n_tiles) * p.bn_box;
      int pcols = p.N - pn0;
      if (pcols > p.bn_box) pcols = p.bn_box;
      const uint32_t b = pf_cc % k2ResBufs;
      mbar_expect_tx(res_bar(b), kEpiBytes);
      tma_load_2d(smem_base + Plan::kResOffset + b * kEpiBytes, &tmap_res, res_bar(b), pn0 + pf_c,
                  (pf_tile / n_tiles) * (2 * kBM) + (int)rank * kBM);
      ++pf_cc;
      pf_c += kEpiCols;
      if (pf_c >= pcols) { pf_c = 0; pf_tile += num_clusters; }
    };
    if (leader && has_res)
      for (int i = 0; i < k2ResBufs - 1; ++i) prefetch_one();
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
      const int m0 = (tile / n_tiles) * (2 * kBM) + (int)rank * kBM, n0 = (tile % n_tiles) * p.bn_box;
      int n_cols = p.N - n0;
      if (n_cols > p.bn_box) n_cols = p.bn_box;
      mbar_wait(tfull_bar(as), aph);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * kBNMax);
      for (int c = 0; c < n_cols; c += kEpiCols, ++cc) {
        const uint32_t buf = cc & 1u;
        const uint32_t rbuf = cc % k2ResBufs;
        uint8_t* out_st = smem_gen + Plan::kOutOffset + buf * kEpiBytes;
        const uint8_t* res_st = smem_gen + Plan::kResOffset + rbuf * kEpiBytes;
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_row + (uint32_t)(c + half * 32), v);
        tmem_ld_wait();
        if (has_res) mbar_wait(res_bar(rbuf), (cc / k2ResBufs) & 1u);
        const int n = n0 + c + half * 32;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float f[8];
          const int ng = n + g * 8;
          if (ng < p.N) {
            const float4 s0 = *(const float4*)(s_scale + ng), s1 = *(const float4*)(s_scale + ng + 4);
            const float4 h0 = *(const float4*)(s_shift + ng), h1 = *(const float4*)(s_shift + ng + 4);
            const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
            const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = __fadd_rn(__fmul_rn(__uint_as_float(v[g * 8 + j]), sc[j]), sh[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = 0.f;
          }
          const int chunk16 = half * 4 + g;                // 16-byte column group inside the 128-byte smem row
          const uint32_t sw_off = (uint32_t)row_in_tile * 128u + (uint32_t)((chunk16 ^ (row_in_tile & 7)) << 4);
          if (has_res) {
            const uint4 r = *(const uint4*)(res_st + sw_off);
            const __nv_bfloat162* rb = (const __nv_bfloat162*)&r;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 rf = __bfloat1622float2(rb[j]);
              f[2 * j] = __fadd_rn(f[2 * j], rf.x); f[2 * j + 1] = __fadd_rn(f[2 * j + 1], rf.y);
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.0f);
          }
          uint4 o;
          __nv_bfloat162* ob = (__nv_bfloat162*)&o;
#pragma unroll
          for (int j = 0; j < 4; ++j) ob[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
          *(uint4*)(out_st + sw_off) = o;
        }
        fence_async_smem();
        if (leader) tma_store_wait_read0();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (leader) {
          tma_store_2d(&tmap_out, smem_base + Plan::kOutOffset + buf * kEpiBytes, n0 + c, m0 + p.out_row_off);
          tma_store_commit();
          // the buffer of chunk cc-1 was fully read before the barrier above: refill it with chunk cc + k2ResBufs - 1
          if (has_res) prefetch_one();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(tempty_bar(as), 0);     // accumulator stage drained -> leader's barrier
      if (++as == ACC) { as = 0; aph ^= 1u; }
    }
    if (leader) tma_store_wait_all();
  }
  tc_fence_before();
  cluster_sync_all();        // no CTA may exit (or free TMEM) while its partner can still signal / read it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 512);
  }
}

// =====================================================================================================
// block1_conv2 (3x3 valid, 32 -> 64) as an INPUT-STATIONARY implicit GEMM.
// The first implementation re-fetched a 128x32 A tile per filter tap (nine L2 round trips per output tile, L2 bound at
// ~195 TFLOP/s).  Here the producer loads ONE window of the activation matrix per M tile -- the 128 virtual rows plus the
// 2*W+2 rows the taps reach forward to (4 TMA boxes of 128 rows, 64-byte swizzle) -- and every tap is the SAME smem window
// with the A descriptor start advanced by (ky*W + kx) rows of 64 bytes.  tcgen05 swizzles on absolute smem address bits, so
// row-shifted descriptors are exact (profiles/umma_probe.py).  The nine 64x32 weight tiles stay resident for the whole
// persistent CTA.  18 MMAs (128x64x16) per tile, four TMEM accumulator stages.
// =====================================================================================================
constexpr int kC2Stages = 3;
constexpr int kC2WinRows = 512;                          // 4 boxes x 128 rows >= 128 + 2*149 + 2
constexpr int kC2WinBytes = kC2WinRows * 64;             // 32 KB
constexpr int kC2BBytes = 9 * 64 * 64;                   // 36 KB: nine taps x [64 n][32 k] bf16
constexpr int kC2AccStages = 4;
constexpr int kC2OutTile = 32 * 64;                      // one epilogue warp's staging tile: [32 px][32 ch] bf16, SWIZZLE_64B
constexpr int kC2OutBytes = 8 * 2 * kC2OutTile;          // eight epilogue warps, double-buffered
constexpr int kC2Smem = kC2BBytes + kC2Stages * kC2WinBytes + kC2OutBytes + 256 + 1024;

// warp 0 TMA, warps 1 and 6 MMA issue (small MMAs are issue-latency bound), warps 2-5 and 7-10 epilogue: two warps per TMEM
// lane quadrant, 32 of the 64 output columns each (with one warp per quadrant the epilogue -- ~2700 cycles per 128-row
// tile against 576 tensor-pipe cycles -- was the critical role)
constexpr int kC2Threads = 11 * 32;
static __global__ void __launch_bounds__(kC2Threads, 1)
conv3x3_is_kernel(const __grid_constant__ CUtensorMap tmap_a /*[rows, 32] box [128 x 32] SW64*/,
                  const __grid_constant__ CUtensorMap tmap_b /*[64, 288] box [64 x 32] SW64*/,
                  const __grid_constant__ CUtensorMap tmap_c /*[n, 147, 147, 64] box [1, 1, 32, 32] SW64*/, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t b_smem = smem_base, win0 = smem_base + kC2BBytes;
  const uint32_t out0 = win0 + kC2Stages * kC2WinBytes;
  const uint32_t bar_base = out0 + kC2OutBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kC2Stages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kC2Stages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kC2Stages + kC2AccStages + a); };
  const uint32_t b_bar = bar_base + 8u * (2 * kC2Stages + 2 * kC2AccStages);
  const uint32_t tmem_slot = b_bar + 8u;
  volatile uint32_t* tmem_slot_ptr =
      (volatile uint32_t*)(smem_gen + kC2BBytes + kC2Stages * kC2WinBytes + kC2OutBytes + 8 * (2 * kC2Stages + 2 * kC2AccStages + 1));
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int total_tiles = (p.M + kBM - 1) / kBM;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_c);
    for (int s = 0; s < kC2Stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < kC2AccStages; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 8); }
    mbar_init(b_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(b_bar, (uint32_t)kC2BBytes);
      for (int t = 0; t < 9; ++t) tma_load_2d(b_smem + t * 4096, &tmap_b, b_bar, t * 32, 0);
      int s = 0; uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_expect_tx(full_bar(s), (uint32_t)kC2WinBytes);
        for (int b = 0; b < 4; ++b)
          tma_load_2d(win0 + s * kC2WinBytes + b * 128 * 64, &tmap_a, full_bar(s), 0, tile * kBM + b * 128);
        if (++s == kC2Stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1 || warp == 6) {
    // two issuing warps take alternate tiles; each walks its loop as a converged warp and one elected lane issues (an
    // `if (lane == 0)` region costs ~25 instructions per tcgen05.mma, see elect_one())
    {
      mbar_wait(b_bar, 0);
      const uint32_t idesc = make_idesc(kBM, 64);
      const uint64_t b_base = make_smem_desc<64>(b_smem);
      const int me = warp == 1 ? 0 : 1;
      int li = 0;                                           // local tile index of this CTA
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++li) {
        const int s = li % kC2Stages, as = li % kC2AccStages;
        const uint32_t ph = (uint32_t)(li / kC2Stages) & 1u, aph = (uint32_t)(li / kC2AccStages) & 1u;
        // the other warp's tile: still OBSERVE its window fill.  With 3 stages and 2 issuers a warp would otherwise see
        // only every other phase of a stage's barrier, and a parity wait cannot tell "one phase behind" from "done"
        // (the aliasing that bit the depthwise ring)
        if ((li & 1) != me) { mbar_wait(full_bar(s), ph); continue; }
        mbar_wait(tempty_bar(as), aph ^ 1u);
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint64_t a_base = make_smem_desc<64>(win0 + s * kC2WinBytes);
        const uint32_t d = tmem_base + (uint32_t)(as * 64);
        if (elect_one()) {
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const int roff = ((t / 3) * p.in_w + (t % 3)) * 64;           // tap = rows shifted inside the window
#pragma unroll
            for (int k = 0; k < 2; ++k)
              umma_bf16(d, a_base + (uint64_t)((roff + k * 32) >> 4), b_base + (uint64_t)((t * 4096 + k * 32) >> 4), idesc,
                        (t | k) ? 1u : 0u);
          }
          umma_commit(empty_bar(s));
          umma_commit(tfull_bar(as));
        }
        __syncwarp();
      }
    }
  } else if ((warp >= 2 && warp <= 5) || warp >= 7) {
    const int quad = warp & 3;                     // TMEM lanes a warp may read: 32 * (warp % 4)
    const int half = warp >= 7 ? 1 : 0;            // columns [32 * half, +32)
    int as = 0; uint32_t aph = 0;
    // Output path: the warp's 32 grid positions x 32 channels are staged as a [32 px][64 B] tile (SWIZZLE_64B, conflict-free
    // 16-byte stores) and written by TMA.  (Per-lane 64-byte global stores cost 32 partial-sector L2 requests per
    // instruction: the kernel ran at 2500 cycles per 128-pixel tile against 576 cycles of MMA work.)  The 32 positions are
    // consecutive in the 149-wide input grid, so they break into at most two output-row segments.  The first is one TMA
    // box at (x0, y): its rows past the end of the image row fall outside the 147-wide output and are clipped.  The lanes
    // behind a row break (one warp in five has some) store their 64 bytes directly: a box with a negative start
    // coordinate is rejected by the hardware (illegal instruction), and a box at x = 0 would write rows it does not own.
    // Grid rows / columns >= 147 are never written, as before.
    const int ew = (warp >= 7 ? warp - 3 : warp - 2);                 // 0..7
    uint32_t li = 0;
    // BN constants of this warp's 32 channels: lane-invariant, kept in registers
    float sc[32], sh[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 s4 = __ldg((const float4*)(p.scale + half * 32) + j), h4 = __ldg((const float4*)(p.shift + half * 32) + j);
      sc[4 * j] = s4.x; sc[4 * j + 1] = s4.y; sc[4 * j + 2] = s4.z; sc[4 * j + 3] = s4.w;
      sh[4 * j] = h4.x; sh[4 * j + 1] = h4.y; sh[4 * j + 2] = h4.z; sh[4 * j + 3] = h4.w;
    }
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++li) {
      const int m0 = tile * kBM + quad * 32;                          // warp-uniform: first grid position of this warp
      const int img0 = m0 / p.in_hw, rem0 = m0 - img0 * p.in_hw;
      const int y0 = rem0 / p.in_w, x0 = rem0 - y0 * p.in_w;
      const uint32_t stg = out0 + (uint32_t)((ew * 2 + (int)(li & 1u)) * kC2OutTile);
      mbar_wait(tfull_bar(as), aph);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * 64 + half * 32), v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
      if (elect_one()) tma_store_wait_read1();                        // the store issued from this buffer two tiles ago has read it
      __syncwarp();
      __nv_bfloat16* direct_ptr = nullptr;                            // lanes behind a row break of the input grid
      if (x0 + lane >= p.in_w && m0 + lane < p.M) {
        int y1 = y0 + 1, img1 = img0;
        if (y1 * p.in_w >= p.in_hw) { y1 = 0; ++img1; }
        if (y1 < p.out_h) direct_ptr = p.out + ((long long)img1 * p.out_hw + (long long)y1 * p.out_w + (x0 + lane - p.in_w)) * p.ldc + half * 32;
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = g * 8 + 2 * j;
          float f0 = __fadd_rn(__fmul_rn(__uint_as_float(v[c]), sc[c]), sh[c]);
          float f1 = __fadd_rn(__fmul_rn(__uint_as_float(v[c + 1]), sc[c + 1]), sh[c + 1]);
          if (p.relu) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(pk[j]) : "f"(f1), "f"(f0));
          else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[j]) : "f"(f1), "f"(f0));
        }
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + (uint32_t)(lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4))),
                     "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
        if (direct_ptr) *(uint4*)(direct_ptr + g * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      fence_async_smem();
      __syncwarp();
      if (elect_one()) {
        if (m0 < p.M && y0 < p.out_h && x0 < p.out_w) tma_store_4d(&tmap_c, stg, half * 32, x0, y0, img0);
        tma_store_commit();
      }
      __syncwarp();
      if (++as == kC2AccStages) { as = 0; aph ^= 1u; }
    }
    if (elect_one()) tma_store_wait_all();
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace sm100

// ---- debug-only SIMT GEMM with the same parameter block (BQ_GEMM=simt); slow, obviously correct ----
static __global__ void gemm_simt_kernel(const GemmParams p) {
  __shared__ float As[32][33];
  __shared__ float Bs[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int m = blockIdx.y * 32 + ty, n = blockIdx.x * 32 + tx;
  const int taps = p.conv_mode ? 9 : 1;
  const int kc = p.conv_mode ? p.K / 9 : p.K;
  float acc = 0.f;
  for (int t = 0; t < taps; ++t) {
    const long long roff = p.conv_mode ? (long long)(t / 3) * p.in_w + (t % 3) : 0;
    for (int k0 = 0; k0 < kc; k0 += 32) {
      const long long ar = (long long)(blockIdx.y * 32 + ty) + roff;
      const int ak = k0 + tx;
      As[ty][tx] = (ar < p.a_rows && ak < kc) ? __bfloat162float(p.a_ptr[ar * p.lda + ak]) : 0.f;
      const int bn = blockIdx.x * 32 + ty;
      Bs[ty][tx] = (bn < p.N && ak < kc) ? __bfloat162float(p.b_ptr[(long long)bn * p.ldb + t * kc + ak]) : 0.f;
      __syncthreads();
#pragma unroll 8
      for (int k = 0; k < 32; ++k) acc = fmaf(As[ty][k], Bs[tx][k], acc);
      __syncthreads();
    }
  }
  if (m >= p.M || n >= p.N) return;
  long long orow = m;
  if (p.conv_mode) {
    const int img = m / p.in_hw, rem = m - img * p.in_hw;
    const int y = rem / p.in_w, x = rem - y * p.in_w;
    if (y >= p.out_h || x >= p.out_w) return;
    orow = (long long)img * p.out_hw + (long long)y * p.out_w + x;
  }
  float v = __fmul_rn(acc, p.scale ? p.scale[n] : p.alpha);
  if (p.shift) v = __fadd_rn(v, p.shift[n]);
  if (p.residual) v = __fadd_rn(v, __bfloat162float(p.residual[orow * p.ldr + n]));
  if (p.relu) v = fmaxf(v, 0.f);
  p.out[(orow + p.out_row_off) * p.ldc + n] = __float2bfloat16_rn(v);
}

}  // namespace bq
