// Depthwise 3x3 on the tensor cores.
//
// A depthwise convolution has no reduction across channels, so it is normally CUDA-core work -- and at 9 FMA per
// output the smem/register-window kernel in layers.cuh is instruction-issue bound.  Here each 3x3 tap is instead ONE
// tcgen05.mma per 16-channel group:   D[128 px x 16 ch] += A_tap[128 px x 16 ch] * diag(w[tap, 16 ch])
// i.e. a 128x16x16 MMA whose B operand is a 16x16 diagonal matrix holding the tap's weights.  15/16 of the
// multiplies are by zero, but the tensor pipe is so much wider than the FMA pipe that this is still several times
// cheaper, and it removes the bf16->fp32 unpacking and all the FMA issue slots.
//
// What makes it cheap on the memory side: the (19+2)x(19+2) halo tile of a 64-channel chunk is loaded ONCE by a 4-D
// TMA (SWIZZLE_128B, zero fill outside the image); output pixel q (in halo-pitch order) needs input row
// q + dy*21 + dx for tap (dy, dx), so the A operand of a tap is the SAME smem tile with the descriptor start advanced
// by whole 128-byte rows.  tcgen05 applies the 128B swizzle on absolute smem address bits, so row-shifted windows
// need no base_offset (measured with profiles/umma_probe.py on B200: every shift exact).  The 36 diagonal B tiles
// of a channel chunk (9 taps x 4 groups, no-swizzle K-major layout, 18 KB) are prebuilt at weight-load time and
// arrive with one bulk copy.  Outputs: TMEM -> registers -> bf16 -> 128-byte rows straight to global.
//
// Numerics: bf16 activations x bf16 weights (exact products) accumulated in fp32 over the 9 taps.
#pragma once

#include "gemm_sm100.cuh"

namespace bq {
namespace dwtc {

using namespace sm100;

constexpr int kTile = 19;                       // output tile edge
constexpr int kHalo = kTile + 2;                // 21
constexpr int kQ0 = kHalo + 1;                  // 22: first output position (hy = 1, hx = 1) in halo-pitch order
constexpr int kMBlocks = 4;                     // 4 x 128 >= 19*21 - 2 = 397 positions
constexpr int kTileRows = 560;                                   // >= kQ0 + 4*128 + kQ0 = 556 rows the shifted windows touch; x128 B = multiple of 1024
constexpr int kTileBytes = kTileRows * 128;
constexpr int kLoadBytes = kHalo * kHalo * 128;                  // 56,448 B written by the TMA
constexpr int kBBytes = 36 * 512;                                // 9 taps x 4 groups x (16x16 bf16)
constexpr int kStages = 2;                                       // halo tiles in flight / TMEM accumulator stages
constexpr int kOutBytes = 47 * 1024;                             // 19*19 rows x 128 B staging for the TMA store (46,208 B)
constexpr int kSmem = kStages * kTileBytes + kOutBytes + kBBytes + 128 + 1024;
constexpr int kThreads = 352;     // warp 0: TMA, warps 1-4: MMA issue (one M-block each; warp 1 owns TMEM), warps 5-8: epilogue, warps 9-10: ReLU

// element (n, k) of a 16x16 K-major no-swizzle operand: 8x(16 B) core matrices, LBO = 128 B (k + 8), SBO = 256 B (n + 8)
__host__ __device__ inline int bdiag_index(int n, int k) { return (n / 8) * 128 + (k / 8) * 64 + (n % 8) * 8 + (k % 8); }

// Persistent and software pipelined: CTA c owns channel chunk (c % chunks) -- its 36 diagonal B tiles are loaded
// once -- and walks the (image, spatial tile) items of that chunk with stride gridDim.x / chunks.  While the MMA
// thread works on item i, the TMA warp is already loading item i+1 into the other smem buffer and the epilogue
// warps drain item i-1 from the other TMEM accumulator stage.
__global__ void __launch_bounds__(kThreads, 1)
depthwise3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap_in /*4-D [C, W, H, N] bf16, box [64, 21, 21, 1], SW128*/,
                       const __grid_constant__ CUtensorMap tmap_out /*4-D [C, W, H, N] bf16, box [64, 19, 19, 1], SW128*/,
                       const __nv_bfloat16* __restrict__ bdiag /*[chunks][36][256]*/,
                       int n_img, int H, int W, int C, int tiles_x, int chunks, int relu_in) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t out_st = smem_base + kStages * kTileBytes;       // 1024-aligned (kTileBytes is a multiple of 1024)
  const uint32_t bmat = out_st + kOutBytes;
  const uint32_t bar0 = bmat + kBBytes;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };            // TMA landed
  auto ready_bar = [&](int s) { return bar0 + 8u * (2 + s); };     // ReLU pre-pass done (only when relu_in)
  auto empty_bar = [&](int s) { return bar0 + 8u * (4 + s); };     // MMAs that read the buffer have retired
  auto afull_bar = [&](int s) { return bar0 + 8u * (6 + s); };     // accumulator stage complete
  auto aempty_bar = [&](int s) { return bar0 + 8u * (8 + s); };    // accumulator stage drained
  const uint32_t b_bar = bar0 + 8u * 10, tmem_slot = bar0 + 8u * 11;
  volatile uint32_t* tmem_slot_ptr = (volatile uint32_t*)(smem_gen + kStages * kTileBytes + kOutBytes + kBBytes + 8 * 11);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.x % chunks, c0 = chunk * 64;
  const int lane_of_chunk = blockIdx.x / chunks, stride = gridDim.x / chunks;
  const int tiles = tiles_x * tiles_x;
  const int n_items = n_img * tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1); mbar_init(ready_bar(s), 2); mbar_init(empty_bar(s), 4);
      mbar_init(afull_bar(s), 4); mbar_init(aempty_bar(s), 4);
    }
    mbar_init(b_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(b_bar, (uint32_t)kBBytes);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(bmat), "l"((uint64_t)(bdiag + (size_t)chunk * 36 * 256)), "r"((uint32_t)kBBytes), "r"(b_bar)
                   : "memory");
      int s = 0; uint32_t ph = 0;
      for (int it = lane_of_chunk; it < n_items; it += stride) {
        const int img = it / tiles, tl = it - img * tiles;
        const int ty0 = (tl / tiles_x) * kTile, tx0 = (tl % tiles_x) * kTile;
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_expect_tx(full_bar(s), (uint32_t)kLoadBytes);
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
            ::"r"(smem_base + s * kTileBytes), "l"((uint64_t)&tmap_in), "r"(full_bar(s)), "r"(c0), "r"(tx0 - 1),
              "r"(ty0 - 1), "r"(img)
            : "memory");
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp <= 4) {
    // ===================== MMA issuers: warp w issues the 36 MMAs of M-block (w - 1) =====================
    if (lane == 0) {
      const int mb = warp - 1;
      mbar_wait(b_bar, 0);
      const uint32_t idesc = make_idesc(128, 16);
      // B: no-swizzle K-major, LBO = 128 B, SBO = 256 B, version 1
      const uint64_t b_base = (uint64_t)((bmat & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 16) |
                              ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
      int s = 0; uint32_t ph = 0;
      for (int it = lane_of_chunk; it < n_items; it += stride) {
        mbar_wait(aempty_bar(s), ph ^ 1u);                   // epilogue drained this accumulator stage
        mbar_wait(relu_in ? ready_bar(s) : full_bar(s), ph);
        tc_fence_after();
        const uint64_t a_base = make_smem_desc<128>(smem_base + s * kTileBytes);
        // tap-outer / channel-group-inner: consecutive MMAs accumulate into DIFFERENT TMEM tiles, so the
        // accumulate-chain latency of one tile (~80 cycles, measured) is hidden behind the other three
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int row = kQ0 + mb * 128 + (t / 3 - 1) * kHalo + (t % 3 - 1);         // >= 0
#pragma unroll
          for (int cg = 0; cg < 4; ++cg) {
            const uint32_t d = tmem_base + (uint32_t)(s * 256 + mb * 64 + cg * 16);
            const uint64_t da = a_base + (uint64_t)((row * 128 + cg * 32) >> 4);
            const uint64_t db = b_base + (uint64_t)(((t * 4 + cg) * 512) >> 4);
            umma_bf16(d, da, db, idesc, t ? 1u : 0u);
          }
        }
        umma_commit(empty_bar(s));                           // smem buffer reusable (4 issuers)
        umma_commit(afull_bar(s));                           // accumulators complete (4 issuers)
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp <= 8) {
    // ===================== epilogue (TMEM lane quadrant = warp % 4) =====================
    // TMEM -> registers -> bf16 -> 128B-swizzled staging tile [19*19 px][64 ch] -> ONE 4-D TMA store per item
    // (coalesced 128-byte rows, image-border and channel-tail clipping done by the TMA unit).
    const int quad = warp & 3;
    const bool leader = (warp == 5 && lane == 0);
    uint8_t* stage = smem_gen + (size_t)kStages * kTileBytes;
    int s = 0; uint32_t ph = 0;
    for (int it = lane_of_chunk; it < n_items; it += stride) {
      const int img = it / tiles, tl = it - img * tiles;
      const int ty0 = (tl / tiles_x) * kTile, tx0 = (tl % tiles_x) * kTile;
      mbar_wait(afull_bar(s), ph);
      tc_fence_after();
      if (leader) tma_store_wait_read0();                    // previous item's store has finished reading the staging tile
      epi_barrier();
#pragma unroll 1
      for (int mb = 0; mb < kMBlocks; ++mb) {
        uint32_t v[64];
        {
          uint32_t (&v0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[0]);
          uint32_t (&v1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[32]);
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(s * 256 + mb * 64);
          tmem_ld_32x32b_x32(taddr, v0);
          tmem_ld_32x32b_x32(taddr + 32u, v1);
          tmem_ld_wait();
        }
        const int q = kQ0 + mb * 128 + quad * 32 + lane;
        const int hy = q / kHalo, hx = q - hy * kHalo;
        if (hx >= 1 && hx <= kTile && hy <= kTile) {
          const int pix = (hy - 1) * kTile + (hx - 1);          // row of the staging tile (x fastest, then y)
          uint8_t* dst = stage + (size_t)pix * 128;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            uint4 o;
            __nv_bfloat162* ob = (__nv_bfloat162*)&o;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              ob[j] = __floats2bfloat162_rn(__uint_as_float(v[g * 8 + 2 * j]), __uint_as_float(v[g * 8 + 2 * j + 1]));
            *(uint4*)(dst + ((g ^ (pix & 7)) << 4)) = o;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(aempty_bar(s));               // accumulator stage free for the item after next
      fence_async_smem();
      epi_barrier();
      if (leader) {
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                     ::"l"((uint64_t)&tmap_out), "r"(out_st), "r"(c0), "r"(tx0), "r"(ty0), "r"(img)
                     : "memory");
        tma_store_commit();
      }
      if (++s == kStages) { s = 0; ph ^= 1u; }
    }
    if (leader) tma_store_wait_all();
  } else if (relu_in) {
    // ===================== ReLU pre-pass (2 warps): once per element, in place, swizzle agnostic =====================
    const __nv_bfloat162 z2 = __floats2bfloat162_rn(0.f, 0.f);
    const int tid = threadIdx.x - 288;
    int s = 0; uint32_t ph = 0;
    for (int it = lane_of_chunk; it < n_items; it += stride) {
      mbar_wait(full_bar(s), ph);
      uint8_t* base = smem_gen + (size_t)s * kTileBytes;
      for (int i = tid; i < kLoadBytes / 16; i += 64) {
        uint4 v = *(uint4*)(base + (size_t)i * 16);
        __nv_bfloat162* b = (__nv_bfloat162*)&v;
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = __hmax2(b[j], z2);
        *(uint4*)(base + (size_t)i * 16) = v;
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(ready_bar(s));
      if (++s == kStages) { s = 0; ph ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace dwtc
}  // namespace bq
