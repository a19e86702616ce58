// Depthwise 3x3 'same' convolution, persistent + TMA-pipelined (production kernel for every depthwise layer that is not
// fused into a pointwise GEMM).  Replaces the one-tile-per-block kernel in layers.cuh (kept as BQ_DW=v2), whose ncu
// profile (profiles/ncu_dw_r1i.md) showed 2 resident blocks/SM each serialising  TMA fill -> in-place ReLU pass ->
// compute: 46 % issue utilisation, 40 % DRAM, 13.7 instructions per output element.
//
//   * one CTA per SM, 1 TMA warp + 3 compute groups of 5 warps; a 4-stage ring of (19+2)^2 x CC halo tiles
//     (CC = 64 or 56 channels; 728 = 13 x 56) filled by ONE 4-D TMA box each (out-of-image halo zero-filled by the
//     TMA unit), so fills of the next tiles overlap the arithmetic of the current ones;
//   * a thread owns 4 channels x TWO adjacent tile columns and walks down them with the 3x4 input window in
//     registers: 4 LDS.64 per row for two outputs (instead of 6), 36 packed FFMA2 (each lane an ordinary RN fma, in
//     the same tap order as the scalar kernel -> bit-identical), four independent accumulator chains;
//   * the pre-activation ReLU is a packed bf16 max on the loaded words (no extra smem pass, no block barrier).
//
// Keras semantics: SeparableConv2D's depthwise stage, no bias, 'same' zero padding (Appendix B of SURVEY.md).
#pragma once
#include "gemm_sm100.cuh"

namespace bq {
namespace dwp {

using namespace bq::sm100;

constexpr int kTile = 19;
constexpr int kHalo = kTile + 2;
constexpr int kPairs = (kTile + 1) / 2;                      // column pairs per tile
constexpr int kGroups = 3;
constexpr int kGroupThreads = 160;                           // >= (64/4) * kPairs, whole warps
constexpr int kThreads = 32 + kGroups * kGroupThreads;       // 512
constexpr int kStages = 4;
__host__ __device__ constexpr int stage_bytes(int CC) { return (kHalo * kHalo * CC * 2 + 127) & ~127; }
constexpr int kSmem = kStages * stage_bytes(64);             // 225,792 B dynamic (+ 64 B static barriers)

// four horizontally adjacent pixels x 4 channels -> fp32 float2 pairs (optionally ReLU'd while still packed bf16).
// The fma pipe (FFMA2 / IMAD, 2 cycles per warp instruction per SM sub-partition) is this kernel's scarce resource
// (ncu: math-pipe-throttle stalls sat on the IMAD.U32 x,0x10000 "shifts" the compiler picks for `<< 16`), so the
// bf16 -> fp32 widening is forced onto the ALU pipe: PRMT for the low half, LOP3 for the high half.  Pixel offsets
// are compile-time immediates of the LDS (CC is a template parameter).
template <int OFF>
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2+%3];" : "=r"(v.x), "=r"(v.y) : "r"(addr), "n"(OFF));
  return v;
}
template <bool RELU>
__device__ __forceinline__ void widen(uint2 v, float2 (&d)[2]) {
  if (RELU) {
    const __nv_bfloat162 z2 = __floats2bfloat162_rn(0.f, 0.f);
    __nv_bfloat162 a = __hmax2(*(__nv_bfloat162*)&v.x, z2), b = __hmax2(*(__nv_bfloat162*)&v.y, z2);
    v.x = *(uint32_t*)&a;
    v.y = *(uint32_t*)&b;
  }
  d[0] = make_float2(__uint_as_float(__byte_perm(v.x, 0u, 0x1044)), __uint_as_float(v.x & 0xFFFF0000u));
  d[1] = make_float2(__uint_as_float(__byte_perm(v.y, 0u, 0x1044)), __uint_as_float(v.y & 0xFFFF0000u));
}
template <bool RELU, int CC>
__device__ __forceinline__ void load4(uint32_t addr, bool has2, float2 (&d)[4][2]) {
  const uint2 v0 = lds64<0>(addr), v1 = lds64<CC * 2>(addr), v2 = lds64<2 * CC * 2>(addr);
  // a lone last column re-reads pixel 2 instead of running past the halo row
  const uint2 v3 = lds64<0>(addr + (has2 ? 3u : 2u) * (uint32_t)(CC * 2));
  widen<RELU>(v0, d[0]);
  widen<RELU>(v1, d[1]);
  widen<RELU>(v2, d[2]);
  widen<RELU>(v3, d[3]);
}

template <bool RELU, int CC>
__global__ void __launch_bounds__(kThreads, 1)
depthwise3x3_pipe_kernel(const __grid_constant__ CUtensorMap tmap_in /*4-D [C, W, H, N], box [CC, 21, 21, 1]*/,
                         const float* __restrict__ w /*[9][C]*/, bf16* __restrict__ out, int n_img, int H, int W, int C,
                         int tiles_x) {
  extern __shared__ __align__(128) uint8_t dwp_smem[];
  __shared__ __align__(8) uint64_t bars[2 * kStages];
  const uint32_t smem_base = smem_u32(dwp_smem);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  const int chunks = C / CC;
  const int tiles = tiles_x * tiles_x;
  const int n_items = n_img * tiles * chunks;
  constexpr int sbytes = stage_bytes(CC);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_in);
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), kGroupThreads / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == 0) {
    // ===================== TMA producer: one halo tile per item =====================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int chunk = it % chunks, t = (it / chunks) % tiles, img = it / (chunks * tiles);
        const int ty0 = (t / tiles_x) * kTile, tx0 = (t % tiles_x) * kTile;
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_expect_tx(full_bar(s), (uint32_t)(kHalo * kHalo * CC * 2));
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
            ::"r"(smem_base + (uint32_t)(s * sbytes)), "l"((uint64_t)&tmap_in), "r"(full_bar(s)), "r"(chunk * CC),
              "r"(tx0 - 1), "r"(ty0 - 1), "r"(img)
            : "memory");
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
    }
    return;
  }

  // ===================== compute groups: group g computes the CTA's items g, g + 3, g + 6, ... =====================
  const int g = (threadIdx.x - 32) / kGroupThreads, tg = (threadIdx.x - 32) % kGroupThreads;
  constexpr int cpc = CC >> 2;
  const int c4 = tg % cpc, pair = tg / cpc;
  constexpr uint32_t row_bytes = (uint32_t)kHalo * CC * 2;
  for (int j = 0;; ++j) {
    const int it = blockIdx.x + j * gridDim.x;
    if (it >= n_items) break;
    const int s = j % kStages;
    const uint32_t ph = (uint32_t)(j / kStages) & 1u;
    // Every warp observes EVERY fill of every stage in order, also those of the other groups' items.  A parity wait
    // only distinguishes "the previous phase" from "this phase": a warp that skipped a stage's intermediate fills
    // (3 groups share 4 stages) and ran ahead -- the fifth warp of a group has no active lane on 18-column tiles
    // with CC = 56 -- passed the wait while the barrier was still one phase behind, read a stage in flux and then
    // released it (seen as run-to-run differences in column 18 of ~5 % of the 37x37 tiles).
    if (j % kGroups != g) { mbar_wait(full_bar(s), ph); continue; }
    const int chunk = it % chunks, t = (it / chunks) % tiles, img = it / (chunks * tiles);
    const int ty0 = (t / tiles_x) * kTile, tx0 = (t % tiles_x) * kTile;
    const int th = min(kTile, H - ty0), tw = min(kTile, W - tx0);
    const int px0 = 2 * pair;
    const bool active = pair < kPairs && px0 < tw;
    const bool has2 = px0 + 1 < tw;
    const int c0 = chunk * CC;
    float2 wr[9][2];
    if (active) {
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const float4 wv = __ldg((const float4*)(w + (int64_t)k * C + c0 + c4 * 4));
        wr[k][0] = make_float2(wv.x, wv.y);
        wr[k][1] = make_float2(wv.z, wv.w);
      }
    }
    mbar_wait(full_bar(s), ph);
    if (active) {
      uint32_t src = smem_base + (uint32_t)(s * sbytes) + (uint32_t)(px0 * CC + c4 * 4) * 2u;   // halo row 0
      // 64-bit base once per item, 32-bit byte offsets inside it (an image is far below 4 GB)
      char* const dbase = (char*)(out + ((int64_t)img * H * W + (int64_t)ty0 * W + tx0 + px0) * C + c0 + c4 * 4);
      const uint32_t drow = (uint32_t)(W * C) * 2u, dcol = (uint32_t)C * 2u;
      uint32_t doff = 0;
      float2 ra[4][2], rb[4][2], rc[4][2];
      load4<RELU, CC>(src, has2, ra);
      load4<RELU, CC>(src + row_bytes, has2, rb);
      src += 2 * row_bytes;                                       // next row to load
      auto step = [&](const float2 (&r0)[4][2], const float2 (&r1)[4][2], float2 (&r2)[4][2]) {
        load4<RELU, CC>(src, has2, r2);
        src += row_bytes;
        float2 a[2][2];
        a[0][0] = a[0][1] = a[1][0] = a[1][1] = make_float2(0.f, 0.f);
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            a[0][h] = __ffma2_rn(r0[kx][h], wr[kx][h], a[0][h]);
            a[1][h] = __ffma2_rn(r0[kx + 1][h], wr[kx][h], a[1][h]);
          }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            a[0][h] = __ffma2_rn(r1[kx][h], wr[3 + kx][h], a[0][h]);
            a[1][h] = __ffma2_rn(r1[kx + 1][h], wr[3 + kx][h], a[1][h]);
          }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            a[0][h] = __ffma2_rn(r2[kx][h], wr[6 + kx][h], a[0][h]);
            a[1][h] = __ffma2_rn(r2[kx + 1][h], wr[6 + kx][h], a[1][h]);
          }
        uint2 o;
        __nv_bfloat162* ob = (__nv_bfloat162*)&o;
        ob[0] = __floats2bfloat162_rn(a[0][0].x, a[0][0].y);
        ob[1] = __floats2bfloat162_rn(a[0][1].x, a[0][1].y);
        *(uint2*)(dbase + doff) = o;
        if (has2) {
          ob[0] = __floats2bfloat162_rn(a[1][0].x, a[1][0].y);
          ob[1] = __floats2bfloat162_rn(a[1][1].x, a[1][1].y);
          *(uint2*)(dbase + (doff + dcol)) = o;
        }
        doff += drow;
      };
      for (int py = 0; py < th; py += 3) {
        step(ra, rb, rc);
        if (py + 1 < th) step(rb, rc, ra);
        if (py + 2 < th) step(rc, ra, rb);
      }
    }
    // The stage was read through the generic proxy (ld.shared) and will be overwritten through the async proxy (TMA):
    // every reading thread orders its loads ahead of that write before the warp releases the stage.
    fence_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_bar(s));                     // this warp no longer reads stage s
  }
}

}  // namespace dwp
}  // namespace bq
