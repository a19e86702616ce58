// block1_conv1 on the tensor cores without giving up fp32 accuracy: uint8 NHWC [n,299,299,3] -> bf16 NHWC [n,149,149,32],
// 3x3 stride 2 valid + per-image standardisation + BN + ReLU  (reference: tf.image.per_image_standardization,
// results.py:255, followed by the first Keras Xception layer).
//
// The CUDA-core kernel spent 432 FFMA2 per output pixel and ran at 13 % of the HBM roofline.  Here the layer is an implicit
// GEMM  D[128 px, 96] = A[128 px, K = 27 -> 32] * B[96, K]^T  issued as two tcgen05.mma (K = 16 each) per 128 pixels:
//   * A holds the RAW uint8 pixels converted to bf16 -- integers 0..255 are exact in bf16;
//   * B holds the fp32 filter split into THREE bf16 terms  w = hi + mid + lo  (3 x 8 mantissa bits = fp32's 24), as three
//     groups of 32 output columns; every product is exact in the fp32 accumulator, so hi + mid + lo recovers the fp32
//     convolution of the raw pixels up to fp32 summation rounding;
//   * the standardisation is affine and the layer is bias-free and 'valid', so it moves to the epilogue:
//        conv(W, (x - mu) / sigma) = (conv(W, x) - mu * sum(W)) * (1 / sigma)
//     followed by the folded BatchNorm (separately rounded multiply and add, as everywhere else), ReLU, bf16.
//
// Persistent CTAs, one work item = 128 consecutive output pixels of one image (<= 2 output rows):
//   warps 1-4  producers: the <= 5 input rows of the item are ONE contiguous byte range -> 16-byte coalesced loads into smem
//              (the next item's loads are in flight while this item's A rows are built); thread = one output pixel =
//              27 byte reads -> bf16 -> four swizzled 16-byte stores (SWIZZLE_64B K-major A stage)
//   warp 0     tcgen05.mma issue (converged warp, elected lane), accumulators double-buffered in TMEM (2 x 96 columns)
//   warps 5-8  epilogue: thread = one pixel (TMEM lane), 32 channels -> 64 contiguous bytes straight to global memory
#pragma once

#include "gemm_sm100.cuh"

namespace bq {
namespace conv1tc {

using namespace sm100;

constexpr int kIn = 299, kOut = 149, kRowBytes = kIn * 3;          // 897
constexpr int kPx = kOut * kOut;                                   // 22,201 output pixels per image
constexpr int kItemsPerImg = (kPx + 127) / 128;                    // 174
constexpr int kN = 96;                                             // hi | mid | lo groups of 32 output channels
constexpr int kABytes = 128 * 64;                                  // 128 rows x 32 bf16
constexpr int kWBytes = kN * 64;                                   // 6 KB
constexpr int kInBytes = 5 * kRowBytes + 32;                       // 5 rows + alignment slack -> round up to 16
constexpr int kInSlots = (kInBytes + 15) / 16;                     // 284 uint4
constexpr int kOffW = 0;
constexpr int kOffA = 8192;                                        // 1024-aligned
constexpr int kOffIn = kOffA + 2 * kABytes;
constexpr int kOffConst = kOffIn + 2 * kInSlots * 16;              // sumw[32], scale[32], shift[32]
constexpr int kOffBar = (kOffConst + 3 * 32 * 4 + 7) & ~7;
constexpr int kSmem = kOffBar + 128 + 1024;
constexpr int kThreads = 9 * 32;

struct Conv1Params {
  const uint8_t* tiles;     // [n, 299, 299, 3]
  const float* mean;        // [n]
  const float* inv_std;     // [n]
  const bf16* w;            // [96][32] bf16 K-major: row = part * 32 + channel, k = tap * 3 + ci (27..31 zero)
  const float* sumw;        // [32] sum over the 27 taps of the fp32 filter
  const float* scale;       // [32] folded BatchNorm
  const float* shift;
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

__global__ void __launch_bounds__(kThreads, 2) conv1_tc_kernel(const Conv1Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar0 = smem_base + kOffBar;
  auto a_full = [&](int s) { return bar0 + 8u * s; };          // [2] 4 producer warps
  auto a_empty = [&](int s) { return bar0 + 8u * (2 + s); };   // [2] commit
  auto acc_full = [&](int s) { return bar0 + 8u * (4 + s); };  // [2] commit
  auto acc_empty = [&](int s) { return bar0 + 8u * (6 + s); }; // [2] 4 epilogue warps
  const uint32_t tmem_slot = bar0 + 8u * 8;
  volatile uint32_t* tmem_slot_ptr = (volatile uint32_t*)(smem_gen + kOffBar + 8 * 8);
  float* s_const = (float*)(smem_gen + kOffConst);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int n_items = p.n_img * kItemsPerImg;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(a_full(s), 4); mbar_init(a_empty(s), 1); mbar_init(acc_full(s), 1); mbar_init(acc_empty(s), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // weights (already in the SWIZZLE_64B K-major layout's logical order: the swizzle is applied here) and constants
  for (int i = threadIdx.x; i < kN * 4; i += blockDim.x) {        // 16-byte chunks: row r, chunk c
    const int r = i >> 2, c = i & 3;
    *(uint4*)(smem_gen + kOffW + r * 64 + ((c ^ ((r >> 1) & 3)) << 4)) = __ldg((const uint4*)(p.w + r * 32 + c * 8));
  }
  for (int i = threadIdx.x; i < 96; i += blockDim.x)
    s_const[i] = i < 32 ? __ldg(p.sumw + i) : (i < 64 ? __ldg(p.scale + i - 32) : __ldg(p.shift + i - 64));
  fence_async_smem();
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);

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
      const uint32_t d = tmem_base + (uint32_t)(s * 128);
      if (elect_one()) {
        umma_bf16(d, da, db, idesc, 0u);
        umma_bf16(d, da + 2u, db + 2u, idesc, 1u);
        umma_commit(a_empty(s));
        umma_commit(acc_full(s));
      }
      __syncwarp();
    }
  } else if (warp <= 4) {
    // ===================== producers =====================
    const int ptid = threadIdx.x - 32;                              // 0..127: output pixel of the item
    auto item_range = [&](int it, int& img, int& p0, int& oy0, int& nbytes, int64_t& gstart) {
      img = it / kItemsPerImg;
      p0 = (it - img * kItemsPerImg) * 128;
      oy0 = p0 / kOut;
      int plast = p0 + 127;
      if (plast > kPx - 1) plast = kPx - 1;
      const int oy1 = plast / kOut;
      const int nrows = 2 * (oy1 - oy0) + 3;                        // input rows 2*oy0 .. 2*oy1 + 2
      gstart = ((int64_t)img * kIn + 2 * oy0) * kRowBytes;
      nbytes = nrows * kRowBytes;
    };
    auto fetch = [&](int it, uint4 (&r)[3], int& delta) {          // coalesced 16-byte loads of the item's input rows
      int img, p0, oy0, nbytes;
      int64_t gstart;
      item_range(it, img, p0, oy0, nbytes, gstart);
      const uint64_t a0 = (uint64_t)(p.tiles + gstart);
      const uint64_t base = a0 & ~(uint64_t)15;
      delta = (int)(a0 - base);
      const int slots = (delta + nbytes + 15) >> 4;
      const uint64_t end = (uint64_t)(p.tiles + (int64_t)p.n_img * kIn * kRowBytes);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int sidx = ptid + k * 128;
        r[k] = make_uint4(0u, 0u, 0u, 0u);
        if (sidx < slots) {
          const uint64_t a = base + (uint64_t)sidx * 16;
          if (a >= (uint64_t)p.tiles && a + 16 <= end) {
            r[k] = __ldg((const uint4*)a);
          } else {                                                  // first / last 16 bytes of the caller's buffer: never read outside it
            uint8_t* rb = (uint8_t*)&r[k];
            for (int b = 0; b < 16; ++b)
              if (a + b >= (uint64_t)p.tiles && a + b < end) rb[b] = __ldg((const uint8_t*)(a + b));
          }
        }
      }
    };
    uint4 nxt[3];
    int nxt_delta = 0;
    if ((int)blockIdx.x < n_items) fetch(blockIdx.x, nxt, nxt_delta);
    uint32_t li = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++li) {
      const int s = (int)(li & 1u);
      const uint32_t ph = (li >> 1) & 1u;
      uint8_t* in_s = smem_gen + kOffIn + s * (kInSlots * 16);
      const int delta = nxt_delta;
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (ptid + k * 128 < kInSlots) *(uint4*)(in_s + (ptid + k * 128) * 16) = nxt[k];
      if (it + (int)gridDim.x < n_items) fetch(it + gridDim.x, nxt, nxt_delta);      // in flight while this item is built
      asm volatile("bar.sync 2, 128;" ::: "memory");
      mbar_wait(a_empty(s), ph ^ 1u);
      int img, p0, oy0, nbytes;
      int64_t gstart;
      item_range(it, img, p0, oy0, nbytes, gstart);
      const int pix = p0 + ptid;
      uint32_t packed[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) packed[k] = 0u;
      if (pix < kPx) {
        const int oy = pix / kOut, ox = pix - oy * kOut;
        const uint8_t* src = in_s + delta + (2 * (oy - oy0)) * kRowBytes + ox * 6;
        float v[28];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int j = 0; j < 9; ++j)       // uint8 -> fp32 without the conversion pipe: 2^23 + b, minus 2^23 (exact)
            v[ky * 9 + j] = __fadd_rn(__uint_as_float(0x4B000000u | (uint32_t)src[ky * kRowBytes + j]), -8388608.f);
        v[27] = 0.f;
#pragma unroll
        for (int k = 0; k < 14; ++k) {
          const __nv_bfloat162 b = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);   // exact: integers 0..255
          packed[k] = *(const uint32_t*)&b;
        }
      }
      uint8_t* a_row = smem_gen + kOffA + s * kABytes + ptid * 64;
      const int sw = (ptid >> 1) & 3;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *(uint4*)(a_row + ((c ^ sw) << 4)) = make_uint4(packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full(s));
      // in_s[s] is rewritten two iterations from now, behind the next iteration's bar.sync: every thread has left this
      // iteration's reads by then
    }
  } else {
    // ===================== epilogue =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    uint32_t li = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++li) {
      const int s = (int)(li & 1u);
      const uint32_t ph = (li >> 1) & 1u;
      const int img = it / kItemsPerImg;
      const int pix = (it - img * kItemsPerImg) * 128 + row;
      const float mu = __ldg(p.mean + img), is = __ldg(p.inv_std + img);
      mbar_wait(acc_full(s), ph);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(s * 128);
      bf16* o = p.out + ((int64_t)img * kPx + pix) * 32;
#pragma unroll
      for (int h = 0; h < 2; ++h) {                                 // 16 channels at a time: hi | mid | lo
        uint32_t hi[16], mid[16], lo[16];
        tmem_ld_32x32b_x16(t_addr + (uint32_t)(h * 16), hi);
        tmem_ld_32x32b_x16(t_addr + (uint32_t)(32 + h * 16), mid);
        tmem_ld_32x32b_x16(t_addr + (uint32_t)(64 + h * 16), lo);
        tmem_ld_wait();
        if (h == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty(s));
        }
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int c = h * 16 + j;
          const float conv = __fadd_rn(__fadd_rn(__uint_as_float(hi[j]), __uint_as_float(mid[j])), __uint_as_float(lo[j]));
          const float x = __fmul_rn(__fadd_rn(conv, -__fmul_rn(mu, s_const[c])), is);      // standardisation moved behind the conv
          f[j] = fmaxf(__fadd_rn(__fmul_rn(x, s_const[32 + c]), s_const[64 + c]), 0.f);     // BN, ReLU
        }
        if (pix < kPx) {
          uint4 o0, o1;
          __nv_bfloat162* b0 = (__nv_bfloat162*)&o0;
          __nv_bfloat162* b1 = (__nv_bfloat162*)&o1;
#pragma unroll
          for (int j = 0; j < 4; ++j) { b0[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]); b1[j] = __floats2bfloat162_rn(f[8 + 2 * j], f[8 + 2 * j + 1]); }
          *(uint4*)(o + h * 16) = o0;
          *(uint4*)(o + h * 16 + 8) = o1;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace conv1tc
}  // namespace bq
