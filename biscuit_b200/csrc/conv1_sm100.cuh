// block1_conv1 on the tensor cores without giving up fp32 accuracy: uint8 NHWC [n,299,299,3] -> bf16 NHWC [n,149,149,32],
// 3x3 stride 2 valid + per-image standardisation + BN + ReLU  (reference: tf.image.per_image_standardization,
// results.py:255, followed by the first Keras Xception layer).
//
// The CUDA-core kernel spent 432 FFMA2 per output pixel and ran at 13 % of the HBM roofline.  Here the layer is an implicit
// GEMM  D[128 px, 32] = A[128 px, 3 x (27 -> 32)] * B[32, 96]^T  issued as six tcgen05.mma (K = 16 each) per 128 pixels:
//   * A holds the uint8 pixels minus m0 = round(tile mean) as bf16 -- integers in [-255, 255] are exact in bf16 (removing
//     the common term up front keeps the accumulators small: exactly 0 on a constant tile);
//   * B holds the fp32 filter split into THREE bf16 terms  w = hi + mid + lo  (3 x 8 significand bits = fp32's 24) laid
//     along K: the same A stage is multiplied with the hi, mid and lo rows in turn and all three accumulate into ONE fp32
//     TMEM tile.  Every product is exact in fp32, so the accumulator holds the fp32 convolution of the mean-removed pixels up
//     to fp32 summation rounding;
//   * the standardisation is affine and the layer is bias-free and 'valid', so it moves behind the convolution together
//     with the folded BatchNorm:  BN(conv(W, (x - m) / sd)) = conv(W, x - m0) * g + h,  g = scale / sd,
//     h = shift - (m - m0) * sum(W) * g, computed per tile in fp64 by tile_stats_kernel; the epilogue is one FMA + ReLU.
//
// Persistent CTAs (two per SM), one work item = 128 consecutive output pixels of one image (<= 2 output rows):
//   warp 9     loader: the <= 5 input rows of an item are ONE contiguous byte range of the tile buffer -> one 1-D bulk copy
//              (cp.async.bulk, 16-byte aligned superset) into a 4-stage ring; with one register-prefetched item per CTA
//              the kernel was DRAM-latency-bound (2460 cycles per item)
//   warps 1-4  producers: thread = one output pixel = 3 x 9 bytes (three aligned word loads + funnel shifts per filter row)
//              -> fp32 by the 2^23 trick (byte placed in the low mantissa bits, minus 2^23: exact, no conversion pipe) ->
//              bf16 = upper halves -> four swizzled 16-byte stores (SWIZZLE_64B K-major A stage)
//   warp 0     tcgen05.mma issue (converged warp, elected lane), accumulators double-buffered in TMEM (2 x 32 columns)
//   warps 5-8  epilogue: thread = one pixel (TMEM lane): FMA + ReLU -> bf16 -> the warp's [32 px][32 ch] staging tile
//              (64-byte rows, SWIZZLE_64B: conflict-free 16-byte stores) -> one TMA store of 2 KB contiguous global
//              memory.  (Per-lane 64-byte global stores made 128 partial-sector L2 requests per warp and item and held
//              the kernel at 1.7 TB/s of writes.)
#pragma once

#include "gemm_sm100.cuh"

namespace bq {
namespace conv1tc {

using namespace sm100;

constexpr int kIn = 299, kOut = 149, kRowBytes = kIn * 3;          // 897
constexpr int kPx = kOut * kOut;                                   // 22,201 output pixels per image
constexpr int kItemsPerImg = (kPx + 127) / 128;                    // 174
constexpr int kN = 32;                                             // output channels
constexpr int kWRows = 96;                                         // hi | mid | lo rows of the weight tile
constexpr int kABytes = 128 * 64;                                  // 128 rows x 32 bf16
constexpr int kWBytes = kWRows * 64;                               // 6 KB (fits below kOffA)
constexpr int kInBytes = 5 * kRowBytes + 32;                       // 5 rows + alignment slack -> round up to 16
constexpr int kInSlots = (kInBytes + 15) / 16;                     // 284 uint4
constexpr int kOffW = 0;
constexpr int kOffA = 8192;                                        // 1024-aligned
static_assert(kWBytes <= kOffA, "weight tile");
constexpr int kInStages = 4;
constexpr int kOffIn = kOffA + 2 * kABytes;
constexpr int kOffOut = (kOffIn + kInStages * kInSlots * 16 + 1023) & ~1023;   // 4 epilogue warps x 2 x [32 px][64 B]
constexpr int kOutTile = 32 * 64;
constexpr int kOffBar = kOffOut + 4 * 2 * kOutTile;
constexpr int kSmem = kOffBar + 256 + 1024;
constexpr int kThreads = 10 * 32;

struct Conv1Params {
  const uint8_t* tiles;     // [n, 299, 299, 3]
  const float* affine;      // [n][64]: g[32] | h[32] per tile (tile_stats_kernel)
  const float* m0;          // [n]: round(tile mean), subtracted from every pixel by the producers
  const bf16* w;            // [96][32] bf16 K-major: row = part * 32 + channel, k = tap * 3 + ci (27..31 zero)
  bf16* out;                // [n, 149, 149, 32]
  int n_img;
};

__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

__global__ void __launch_bounds__(kThreads, 2)
conv1_tc_kernel(const __grid_constant__ CUtensorMap tmap_out /*[n * 22201, 32] box [32 x 32], SW64*/, const Conv1Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar0 = smem_base + kOffBar;
  auto a_full = [&](int s) { return bar0 + 8u * s; };          // [2] 4 producer warps
  auto a_empty = [&](int s) { return bar0 + 8u * (2 + s); };   // [2] commit
  auto acc_full = [&](int s) { return bar0 + 8u * (4 + s); };  // [2] commit
  auto acc_empty = [&](int s) { return bar0 + 8u * (6 + s); }; // [2] 4 epilogue warps
  auto in_full = [&](int s) { return bar0 + 8u * (8 + s); };   // [kInStages] bulk copy landed
  auto in_empty = [&](int s) { return bar0 + 8u * (8 + kInStages + s); };   // [kInStages] 4 producer warps
  const uint32_t tmem_slot = bar0 + 8u * (8 + 2 * kInStages);
  volatile uint32_t* tmem_slot_ptr = (volatile uint32_t*)(smem_gen + kOffBar + 8 * (8 + 2 * kInStages));

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int n_items = p.n_img * kItemsPerImg;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_out);
    for (int s = 0; s < 2; ++s) { mbar_init(a_full(s), 4); mbar_init(a_empty(s), 1); mbar_init(acc_full(s), 1); mbar_init(acc_empty(s), 4); }
    for (int s = 0; s < kInStages; ++s) { mbar_init(in_full(s), 1); mbar_init(in_empty(s), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // weights (already in the SWIZZLE_64B K-major layout's logical order: the swizzle is applied here) and constants
  for (int i = threadIdx.x; i < kWRows * 4; i += blockDim.x) {    // 16-byte chunks: row r, chunk c
    const int r = i >> 2, c = i & 3;
    *(uint4*)(smem_gen + kOffW + r * 64 + ((c ^ ((r >> 1) & 3)) << 4)) = __ldg((const uint4*)(p.w + r * 32 + c * 8));
  }
  fence_async_smem();
  if (warp == 0) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);

  // an item's input rows 2*oy0 .. 2*oy1 + 2 as one byte range of the tile buffer
  auto item_range = [&](int it, int& nbytes, int64_t& gstart) {
    const int img = it / kItemsPerImg;
    const int p0 = (it - img * kItemsPerImg) * 128;
    const int oy0 = p0 / kOut;
    int plast = p0 + 127;
    if (plast > kPx - 1) plast = kPx - 1;
    const int oy1 = plast / kOut;
    gstart = ((int64_t)img * kIn + 2 * oy0) * kRowBytes;
    nbytes = (2 * (oy1 - oy0) + 3) * kRowBytes;
  };

  if (warp == 0) {
    // ===================== MMA issue =====================
    const uint32_t idesc = make_idesc(128, kN);
    const uint64_t db = make_smem_desc<64>(smem_base + kOffW);
    uint32_t li = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++li) {
      const int s = (int)(li & 1u);
      const uint32_t ph = (li >> 1) & 1u;
      mbar_wait(a_full(s), ph);
      mbar_wait(acc_empty(s), ph ^ 1u);
      tc_fence_after();
      const uint64_t da = make_smem_desc<64>(smem_base + kOffA + s * kABytes);
      const uint32_t d = tmem_base + (uint32_t)(s * 32);
      if (elect_one()) {
#pragma unroll
        for (int part = 0; part < 3; ++part) {                      // hi, mid, lo rows of B against the same A stage
          const uint64_t dbp = db + (uint64_t)(part * ((32 * 64) >> 4));
          umma_bf16(d, da, dbp, idesc, part ? 1u : 0u);
          umma_bf16(d, da + 2u, dbp + 2u, idesc, 1u);
        }
        umma_commit(a_empty(s));
        umma_commit(acc_full(s));
      }
      __syncwarp();
    }
  } else if (warp == 9) {
    // ===================== loader =====================
    uint32_t li = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++li) {
      const int s = (int)(li % kInStages);
      const uint32_t ph = (li / kInStages) & 1u;
      int nbytes;
      int64_t gstart;
      item_range(it, nbytes, gstart);
      const uint64_t a0 = (uint64_t)(p.tiles + gstart);
      const uint64_t base = a0 & ~(uint64_t)15;                     // aligned superset: <= 15 bytes either side, never used
      const uint32_t size = (uint32_t)(((a0 - base) + (uint64_t)nbytes + 15) & ~(uint64_t)15);
      mbar_wait(in_empty(s), ph ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(in_full(s), size);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_base + kOffIn + s * (kInSlots * 16)), "l"(base), "r"(size), "r"(in_full(s))
                     : "memory");
      }
      __syncwarp();
    }
  } else if (warp <= 4) {
    // ===================== producers =====================
    const int ptid = threadIdx.x - 32;                              // 0..127: output pixel of the item
    uint32_t li = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++li) {
      const int s = (int)(li & 1u);
      const uint32_t ph = (li >> 1) & 1u;
      const int si = (int)(li % kInStages);
      const uint8_t* in_s = smem_gen + kOffIn + si * (kInSlots * 16);
      int nbytes;
      int64_t gstart;
      item_range(it, nbytes, gstart);
      const int delta = (int)((uint64_t)(p.tiles + gstart) & 15);
      const int img = it / kItemsPerImg;
      const int p0 = (it - img * kItemsPerImg) * 128;
      const int oy0 = p0 / kOut;
      const int pix = p0 + ptid;
      const float magic = __fadd_rn(8388608.f, __ldg(p.m0 + img));  // 2^23 + m0: exact (an integer below 2^24)
      mbar_wait(in_full(si), (li / kInStages) & 1u);
      mbar_wait(a_empty(s), ph ^ 1u);
      uint32_t packed[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) packed[k] = 0u;
      if (pix < kPx) {
        const int oy = pix / kOut, ox = pix - oy * kOut;
        const uint32_t o0 = (uint32_t)(delta + (2 * (oy - oy0)) * kRowBytes + ox * 6);
        uint32_t f[28];                                             // fp32 bit patterns of the 27 values x - m0 (+ 0)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const uint32_t o = o0 + (uint32_t)(ky * kRowBytes);
          const uint32_t* wp = (const uint32_t*)(in_s + (o & ~3u));
          const uint32_t sh = (o & 3u) * 8u;
          const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2];
          const uint32_t a[3] = {__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), w2 >> sh};
#pragma unroll
          for (int j = 0; j < 9; ++j)       // byte -> 0x4B0000bb = 2^23 + b; minus (2^23 + m0) is exact: b - m0
            f[ky * 9 + j] = __float_as_uint(__fadd_rn(__uint_as_float(__byte_perm(a[j >> 2], 0x4B000000u, 0x7440u | (uint32_t)(j & 3))), -magic));
        }
        f[27] = 0u;
#pragma unroll
        for (int k = 0; k < 14; ++k) packed[k] = __byte_perm(f[2 * k], f[2 * k + 1], 0x7632u);   // upper halves: exact bf16 of -255..255
      }
      uint8_t* a_row = smem_gen + kOffA + s * kABytes + ptid * 64;
      const int sw = (ptid >> 1) & 3;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *(uint4*)(a_row + ((c ^ sw) << 4)) = make_uint4(packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
      fence_async_smem();           // A rows -> tensor core (async proxy); also orders the ring reads ahead of the next bulk copy
      __syncwarp();
      if (lane == 0) { mbar_arrive(a_full(s)); mbar_arrive(in_empty(si)); }
    }
  } else {
    // ===================== epilogue =====================
    const int quad = warp & 3;
    const int quad_px = quad * 32;
    const int row = quad_px + lane;
    uint32_t li = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++li) {
      const int s = (int)(li & 1u);
      const uint32_t ph = (li >> 1) & 1u;
      const int img = it / kItemsPerImg;
      const int pix = (it - img * kItemsPerImg) * 128 + row;
      const float4* aff = (const float4*)(p.affine + (int64_t)img * 64);     // warp-uniform address: one broadcast per load
      float4 g4[4], h4[4];                                          // channels 0..15 now, 16..31 (same two L1 lines) later
#pragma unroll
      for (int j = 0; j < 4; ++j) { g4[j] = __ldg(aff + j); h4[j] = __ldg(aff + 8 + j); }
      mbar_wait(acc_full(s), ph);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(s * 32);
      uint32_t acc[32];
      tmem_ld_32x32b_x16(t_addr, *reinterpret_cast<uint32_t (*)[16]>(&acc[0]));
      tmem_ld_32x32b_x16(t_addr + 16u, *reinterpret_cast<uint32_t (*)[16]>(&acc[16]));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(s));
      const int p0 = (it - img * kItemsPerImg) * 128;
      const bool whole = p0 + 128 <= kPx;                           // the last item of an image is ragged: direct stores
      const uint32_t stg = smem_base + kOffOut + (uint32_t)(((warp - 5) * 2 + (int)(li & 1u)) * kOutTile);
      if (whole) {
        if (elect_one()) tma_store_wait_read1();                    // the store issued from this buffer two items ago has read it
        __syncwarp();
      }
      bf16* o = p.out + ((int64_t)img * kPx + pix) * 32;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        if (hf == 1) {
#pragma unroll
          for (int j = 0; j < 4; ++j) { g4[j] = __ldg(aff + 4 + j); h4[j] = __ldg(aff + 12 + j); }
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {                               // 8 channels = 16 bytes per store
          const float gg[8] = {g4[2 * q].x, g4[2 * q].y, g4[2 * q].z, g4[2 * q].w, g4[2 * q + 1].x, g4[2 * q + 1].y, g4[2 * q + 1].z, g4[2 * q + 1].w};
          const float hh[8] = {h4[2 * q].x, h4[2 * q].y, h4[2 * q].z, h4[2 * q].w, h4[2 * q + 1].x, h4[2 * q + 1].y, h4[2 * q + 1].z, h4[2 * q + 1].w};
          uint32_t pk[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = hf * 16 + q * 8 + 2 * j;
            const float y0 = __fmaf_rn(__uint_as_float(acc[c]), gg[2 * j], hh[2 * j]);
            const float y1 = __fmaf_rn(__uint_as_float(acc[c + 1]), gg[2 * j + 1], hh[2 * j + 1]);
            asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(pk[j]) : "f"(y1), "f"(y0));
          }
          const int chunk = hf * 2 + q;
          if (whole) {
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + (uint32_t)(lane * 64 + ((chunk ^ ((lane >> 1) & 3)) << 4))),
                         "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
          } else if (pix < kPx) {
            *(uint4*)(o + chunk * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
      }
      if (whole) {
        fence_async_smem();
        __syncwarp();
        if (elect_one()) {
          tma_store_2d(&tmap_out, stg, 0, img * kPx + p0 + quad_px);
          tma_store_commit();
        }
        __syncwarp();
      }
    }
    if (elect_one()) tma_store_wait_all();
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

}  // namespace conv1tc
}  // namespace bq
