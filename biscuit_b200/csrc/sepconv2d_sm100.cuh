// Fused SeparableConv2D for the entry-flow layers with K <= 256 and N <= 256 (block2 / block3 sepconvs on the 147^2 and
// 74^2 maps -- the layers that are HBM-bound when the depthwise result makes a round trip through memory):
//
//     out[8x16 px, N] = epilogue( depthwise3x3(relu?(x))[8x16 px, K] @ Wpw[N, K]^T )
//
// A work item is an 8 x 16 pixel patch of one image.  The GEMM is issued TRANSPOSED (as in sepmid_sm100.cuh): the weights
// are the M-side operand and the 128 pixels the N-side operand, D^T[128 ch, 128 px] += W[128, 64] * dw[128, 64]^T per
// channel tile, so a TMEM lane is an output CHANNEL: the epilogue keeps its BatchNorm constants in four registers and
// transposes with stmatrix instead of fetching two constants per output from shared memory (these layers were
// epilogue-bound: 4.1 -> 2.6 instructions per output).  Per 64-channel k-block:
//   warp 0      one 4-D TMA load of the (8+2) x (16+2) x 64 halo patch (zero fill outside the image = 'same' padding) and
//               one TMA load of the [N x 64] pointwise-weight k-block;
//   warps 6-13  depthwise producers: thread = (4 channels, one patch column), vertical 3x3 register window walking down
//               the 8 rows (3 LDS.64 + 18 FFMA2 per 4 outputs, weights in registers) -> bf16 -> the 128B-swizzled A stage;
//   warp 1      tcgen05.mma 128 x 128 x 16 (x4 per k-block and channel tile), accumulators double-buffered in TMEM;
//   warps 2-9   epilogue, eight independent warp pipelines (no block barrier): warp = (32 channels, 64 pixels);
//               tcgen05.ld.16x256b accumulator fragments -> BN scale/shift (ReLU fused into the bf16 pack) ->
//               stmatrix.trans into the warp's [32 px][32 ch] staging tile -> 4-D TMA store of two patch rows (image
//               border clipped by the TMA unit).
// The depthwise output never touches L2/HBM (saves one tensor write + one tensor read per layer) and its CUDA-core work
// runs under the TMA / tensor-core work of the same kernel.
#pragma once

#include "gemm_sm100.cuh"

namespace bq {
namespace sep2d {

using namespace sm100;

constexpr int kPH = 8, kPW = 16;                               // output patch (rows x cols) = 128 pixels
constexpr int kHH = kPH + 2, kHW = kPW + 2;                    // halo patch 10 x 18
constexpr int kPatchBytes = 23 * 1024;                         // 10*18*128 = 23,040 B, padded to a multiple of 1024
constexpr int kABytes = 128 * 128;
constexpr int kOutTile = 32 * 64;                              // one epilogue warp's staging tile: [32 px][32 ch] bf16, SWIZZLE_64B
constexpr int kOutBufs = 2;                                    // rotating per warp: one store may still be reading (N = 256 leaves no room for a third)
constexpr int kOutBytes = 8 * kOutBufs * kOutTile;             // eight epilogue warps
constexpr int kMaxPStages = 5, kAStages = 2, kBStages = 2;
constexpr int kOffPatch = 0;
// runtime layout (all offsets multiples of 1024): [patch x P][A x 2][B x 2 (N*128 B each)][out][scale/shift][barriers]
__host__ __device__ inline int patch_stages(int N) { (void)N; return 4; }
__host__ __device__ inline int off_a(int N) { return patch_stages(N) * kPatchBytes; }
__host__ __device__ inline int off_b(int N) { return off_a(N) + kAStages * kABytes; }
__host__ __device__ inline int off_out(int N) { return off_b(N) + kBStages * N * 128; }
__host__ __device__ inline int off_scale(int N) { return off_out(N) + kOutBytes; }
__host__ __device__ inline int off_bar(int N) { return off_scale(N) + 2 * 256 * 4; }
__host__ __device__ inline int smem_bytes(int N) { return off_bar(N) + 512 + 1024; }
constexpr int kThreads = 576;                                 // warp 0 TMA, 1 MMA, 2-9 epilogue (8), 10-17 producers (8)
constexpr int kProducerWarps = 8;

struct Sep2dParams {
  int n_img, H, W;          // images in this launch, map size
  int K, N;                 // input / output channels (K % 64 == 0, K <= 256; N = 128 or 256)
  int relu_in, relu_out;
  const float* dw;          // [9][K]
  const float* scale;       // [N]
  const float* shift;
};

template <bool RELU_IN>
__global__ void __launch_bounds__(kThreads, 1)
sepconv2d_fused_kernel(const __grid_constant__ CUtensorMap tmap_x /*4-D [K, W, H, n] box [64, 18, 10, 1], no swizzle*/,
                       const __grid_constant__ CUtensorMap tmap_w /*2-D [N, K] box [64 x N], SW128*/,
                       const __grid_constant__ CUtensorMap tmap_out /*4-D [N, W, H, n] box [32, 16, 2, 1], SW64*/,
                       const Sep2dParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int kPStages = patch_stages(p.N);
  const int kOffA = off_a(p.N), kOffB = off_b(p.N), kOffOut = off_out(p.N), kOffScale = off_scale(p.N), kOffBar = off_bar(p.N);
  const int kBBytes = p.N * 128;
  const uint32_t bar0 = smem_base + kOffBar;
  auto patch_full = [&](int s) { return bar0 + 8u * s; };                       // [5]
  auto patch_empty = [&](int s) { return bar0 + 8u * (5 + s); };                // [5]
  auto a_full = [&](int s) { return bar0 + 8u * (10 + s); };                    // [2]
  auto a_empty = [&](int s) { return bar0 + 8u * (12 + s); };                   // [2]
  auto b_full = [&](int s) { return bar0 + 8u * (14 + s); };                    // [2]
  auto b_empty = [&](int s) { return bar0 + 8u * (16 + s); };                   // [2]
  auto acc_full = [&](int s) { return bar0 + 8u * (18 + s); };                  // [2]
  auto acc_empty = [&](int s) { return bar0 + 8u * (20 + s); };                 // [2]
  const uint32_t tmem_slot = bar0 + 8u * 22;
  volatile uint32_t* tmem_slot_ptr = (volatile uint32_t*)(smem_gen + kOffBar + 8 * 22);
  float* s_scale = (float*)(smem_gen + kOffScale);
  float* s_shift = s_scale + 256;

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int px_tiles = (p.W + kPW - 1) / kPW, py_tiles = (p.H + kPH - 1) / kPH;
  const int per_img = px_tiles * py_tiles;
  const int n_items = p.n_img * per_img;
  const int num_kb = p.K / 64;
  const uint32_t b_bytes = (uint32_t)p.N * 128u;
  // K = 64 / 128 (one / two k-blocks per item): stage s of the B ring always holds k-block s % num_kb, so the weights
  // are loaded once per CTA instead of once per item (16 KB of the 39 KB an item used to pull through TMA)
  const bool b_resident = num_kb <= kBStages;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_out);
    for (int s = 0; s < kPStages; ++s) { mbar_init(patch_full(s), 1); mbar_init(patch_empty(s), kProducerWarps / 2); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(a_full(s), kProducerWarps / 2); mbar_init(a_empty(s), 1);
      mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1);
      mbar_init(acc_full(s), 1); mbar_init(acc_empty(s), 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < p.N; i += blockDim.x) { s_scale[i] = __ldg(p.scale + i); s_shift[i] = __ldg(p.shift + i); }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);   // warp-uniform for the compiler

  if (warp == 0) {
    // ===================== TMA: halo patch + weight k-block =====================
    if (lane == 0) {
      int ps = 0; uint32_t pph = 0;
      int bs = 0; uint32_t bph = 0;
      if (b_resident) {                                          // K <= 128: the whole weight matrix fits the two B stages
        for (int st = 0; st < kBStages; ++st) {
          mbar_expect_tx(b_full(st), b_bytes);
          tma_load_2d(smem_base + kOffB + st * kBBytes, &tmap_w, b_full(st), (st % num_kb) * 64, 0);
        }
      }
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int img = it / per_img, t = it - img * per_img;
        const int y0 = (t / px_tiles) * kPH, x0 = (t % px_tiles) * kPW;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(patch_empty(ps), pph ^ 1u);
          mbar_expect_tx(patch_full(ps), (uint32_t)(kHH * kHW * 128));
          asm volatile(
              "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
              ::"r"(smem_base + kOffPatch + ps * kPatchBytes), "l"((uint64_t)&tmap_x), "r"(patch_full(ps)), "r"(kb * 64),
                "r"(x0 - 1), "r"(y0 - 1), "r"(img)
              : "memory");
          if (++ps == kPStages) { ps = 0; pph ^= 1u; }
          if (b_resident) continue;                              // weights were loaded once, above
          mbar_wait(b_empty(bs), bph ^ 1u);
          mbar_expect_tx(b_full(bs), b_bytes);
          tma_load_2d(smem_base + kOffB + bs * kBBytes, &tmap_w, b_full(bs), kb * 64, 0);
          if (++bs == kBStages) { bs = 0; bph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: converged warp, one elected lane issues (see elect_one()) =====================
    {
      const uint32_t idesc = make_idesc(128, 128);   // M = 128 output channels (weights), N = the patch's 128 pixels
      const int n_ct = p.N >> 7;
      int as = 0; uint32_t aph = 0;                // A / B rings advance together (one step per k-block)
      int cs = 0; uint32_t cph = 0;                // accumulator stage per item
      int gk = 0;                                  // k-blocks issued so far (resident weights: only the first two wait for B)
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        mbar_wait(acc_empty(cs), cph ^ 1u);
#pragma unroll 1
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(a_full(as), aph);
          if (!b_resident || gk < kBStages) mbar_wait(b_full(as), aph);
          ++gk;
          tc_fence_after();
          const uint64_t dpx = make_smem_desc<128>(smem_base + kOffA + as * kABytes);      // depthwise stage: [128 px][64 k]
          const uint64_t dw0 = make_smem_desc<128>(smem_base + kOffB + as * kBBytes);      // weights: [N ch][64 k]
          if (elect_one()) {
            for (int ct = 0; ct < n_ct; ++ct) {
              const uint64_t dw = dw0 + (uint64_t)(ct * ((128 * 128) >> 4));
              const uint32_t d = tmem_base + (uint32_t)(cs * 256 + ct * 128);
              umma_bf16(d, dw, dpx, idesc, kb ? 1u : 0u);
              umma_bf16(d, dw + 2u, dpx + 2u, idesc, 1u);
              umma_bf16(d, dw + 4u, dpx + 4u, idesc, 1u);
              umma_bf16(d, dw + 6u, dpx + 6u, idesc, 1u);
            }
            umma_commit(a_empty(as));
            if (!b_resident) umma_commit(b_empty(as));
            if (kb == num_kb - 1) umma_commit(acc_full(cs));
          }
          __syncwarp();
          if (++as == 2) { as = 0; aph ^= 1u; }
        }
        if (++cs == 2) { cs = 0; cph ^= 1u; }
      }
    }
  } else if (warp < 10) {
    // ===================== epilogue: eight independent warp pipelines =====================
    // warp = (TMEM lane quadrant: 32 channels of a channel tile, pixel half: patch rows 4 hh .. 4 hh + 3).  A step is
    // 32 pixels (two patch rows) x 32 channels: thread t holds channels t/4 + {0, 8, 16, 24} and the pixel pairs 2(t%4) of
    // every 8-pixel block (accumulator-fragment layout), so the BatchNorm constants are four scale/shift pairs per
    // thread and channel tile, a pixel pair packs into one bf16x2 register (cvt.rn.relu fuses the ReLU), and one
    // stmatrix.x4.trans writes an [8 px][32 ch] block of the staging tile (64-byte rows, SWIZZLE_64B).
    const int quad = warp & 3;
    const int hh = (warp - 2) >> 2;
    const int w8 = warp - 2;
    const int n_ct = p.N >> 7;
    const uint32_t my_out = smem_base + kOffOut + (uint32_t)(w8 * kOutBufs * kOutTile);
    const int mrow = lane & 7, mmat = lane >> 3;                  // this lane's row-address duty for stmatrix
    int cs = 0; uint32_t cph = 0;
    uint32_t g = 0;                                               // running step counter of this warp
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const int img = it / per_img, t = it - img * per_img;
      const int y0 = (t / px_tiles) * kPH, x0 = (t % px_tiles) * kPW;
      mbar_wait(acc_full(cs), cph);
      tc_fence_after();
      for (int ct = 0; ct < n_ct; ++ct) {
        float s4[4], h4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int c = ct * 128 + quad * 32 + (lane >> 2) + 8 * k;
          s4[k] = s_scale[c];
          h4[k] = s_shift[c];
        }
        const uint32_t t_lo = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(cs * 256 + ct * 128 + hh * 64);
        const uint32_t t_hi = t_lo + (16u << 16);
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2, ++g) {
          uint32_t lo[16], hi[16];
          tmem_ld_16x256b_x4(t_lo + (uint32_t)(s2 * 32), lo);
          tmem_ld_16x256b_x4(t_hi + (uint32_t)(s2 * 32), hi);
          tmem_ld_wait();
          if (ct == n_ct - 1 && s2 == 1) {                        // accumulators fully read -> the next item may start
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty(cs));
          }
          uint32_t pkv[16];
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            float f[8];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              f[2 * k] = __fadd_rn(__fmul_rn(__uint_as_float(lo[4 * b + 2 * k]), s4[k]), h4[k]);
              f[2 * k + 1] = __fadd_rn(__fmul_rn(__uint_as_float(lo[4 * b + 2 * k + 1]), s4[k]), h4[k]);
              f[4 + 2 * k] = __fadd_rn(__fmul_rn(__uint_as_float(hi[4 * b + 2 * k]), s4[2 + k]), h4[2 + k]);
              f[4 + 2 * k + 1] = __fadd_rn(__fmul_rn(__uint_as_float(hi[4 * b + 2 * k + 1]), s4[2 + k]), h4[2 + k]);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) pkv[4 * b + k] = pack_bf16x2(f[2 * k], f[2 * k + 1], p.relu_out != 0);
          }
          const uint32_t stg = my_out + (g % kOutBufs) * kOutTile;
          // buffer g % kOutBufs was last stored from kOutBufs steps ago: only now must that store have finished reading it
          if (elect_one()) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kOutBufs - 1) : "memory");
          __syncwarp();
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const int r = 8 * b + mrow;
            stmatrix_x4_trans(stg + r * 64 + ((mmat ^ ((r >> 1) & 3)) << 4), pkv[4 * b], pkv[4 * b + 1], pkv[4 * b + 2], pkv[4 * b + 3]);
          }
          fence_async_smem();
          __syncwarp();
          if (elect_one()) {
            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                         ::"l"((uint64_t)&tmap_out), "r"(stg), "r"(ct * 128 + quad * 32), "r"(x0), "r"(y0 + hh * 4 + s2 * 2), "r"(img)
                         : "memory");
            tma_store_commit();
          }
          __syncwarp();
        }
      }
      if (++cs == 2) { cs = 0; cph ^= 1u; }
    }
    if (elect_one()) tma_store_wait_all();
    __syncwarp();
  } else {
    // ===================== depthwise producers =====================
    // Two groups of four warps take alternate k-blocks (group g owns A stage g); a thread = (4 channels, TWO adjacent
    // patch columns) walks down the 8 rows with the 3x4 input window in registers: 4 LDS.64 per row for two outputs
    // instead of 6, four independent FFMA2 accumulator chains.
    const int ptid = threadIdx.x - 320;                         // 0..255
    const int grp = ptid >> 7, tg = ptid & 127;
    const int c4 = tg & 15, pair = tg >> 4;                     // channel group, column pair 0..7
    auto load4 = [&](const uint8_t* rowp, float2 (&d)[4][2]) {  // halo pixels 2*pair .. 2*pair+3 of one halo row
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint2 raw = *(const uint2*)(rowp + (size_t)k * 128);
        if (RELU_IN) {                                            // ReLU on the packed bf16 pairs
          const __nv_bfloat162 z2 = __floats2bfloat162_rn(0.f, 0.f);
          __nv_bfloat162* hb = (__nv_bfloat162*)&raw;
          hb[0] = __hmax2(hb[0], z2);
          hb[1] = __hmax2(hb[1], z2);
        }
        d[k][0] = make_float2(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xFFFF0000u));
        d[k][1] = make_float2(__uint_as_float(raw.y << 16), __uint_as_float(raw.y & 0xFFFF0000u));
      }
    };
    const int total_kb = ((n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * num_kb;   // this CTA's k-blocks
    for (int q = grp; q < total_kb; q += 2) {
      const int kb = q % num_kb;
      const int ps = q % kPStages;
      const uint32_t pph = (uint32_t)(q / kPStages) & 1u, aph = (uint32_t)(q >> 1) & 1u;
      const int c = kb * 64 + c4 * 4;
      float2 w[9][2];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float4 wv = __ldg((const float4*)(p.dw + (size_t)t * p.K + c));
        w[t][0] = make_float2(wv.x, wv.y);
        w[t][1] = make_float2(wv.z, wv.w);
      }
      // observe EVERY patch fill in order, also the other group's (q - 1): with an odd ring depth a group sees only
      // every other phase of a stage's barrier, and a parity wait cannot tell "one phase behind" from "done"
      if (q > 0) mbar_wait(patch_full((q - 1) % kPStages), (uint32_t)((q - 1) / kPStages) & 1u);
      mbar_wait(patch_full(ps), pph);
      mbar_wait(a_empty(grp), aph ^ 1u);
      const uint8_t* col = smem_gen + kOffPatch + ps * kPatchBytes + (size_t)(2 * pair) * 128 + c4 * 8;   // halo (row 0, col 2*pair)
      uint8_t* a_dst = smem_gen + kOffA + grp * kABytes;
      float2 ra[4][2], rb[4][2], rc[4][2];
      load4(col, ra);
      load4(col + (size_t)kHW * 128, rb);
      auto step = [&](const float2 (&r0)[4][2], const float2 (&r1)[4][2], float2 (&r2)[4][2], int py) {
        load4(col + (size_t)(py + 2) * kHW * 128, r2);
        float2 a[2][2];
        a[0][0] = a[0][1] = a[1][0] = a[1][1] = make_float2(0.f, 0.f);
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            a[0][h] = __ffma2_rn(r0[kx][h], w[kx][h], a[0][h]);
            a[1][h] = __ffma2_rn(r0[kx + 1][h], w[kx][h], a[1][h]);
          }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            a[0][h] = __ffma2_rn(r1[kx][h], w[3 + kx][h], a[0][h]);
            a[1][h] = __ffma2_rn(r1[kx + 1][h], w[3 + kx][h], a[1][h]);
          }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            a[0][h] = __ffma2_rn(r2[kx][h], w[6 + kx][h], a[0][h]);
            a[1][h] = __ffma2_rn(r2[kx + 1][h], w[6 + kx][h], a[1][h]);
          }
#pragma unroll
        for (int cx = 0; cx < 2; ++cx) {
          const int r = py * kPW + 2 * pair + cx;                 // A row of this output pixel
          uint2 o;
          __nv_bfloat162* ob = (__nv_bfloat162*)&o;
          ob[0] = __floats2bfloat162_rn(a[cx][0].x, a[cx][0].y);
          ob[1] = __floats2bfloat162_rn(a[cx][1].x, a[cx][1].y);
          *(uint2*)(a_dst + (size_t)r * 128 + (((c4 >> 1) ^ (r & 7)) << 4) + (c4 & 1) * 8) = o;
        }
      };
      step(ra, rb, rc, 0); step(rb, rc, ra, 1); step(rc, ra, rb, 2);
      step(ra, rb, rc, 3); step(rb, rc, ra, 4); step(rc, ra, rb, 5);
      step(ra, rb, rc, 6); step(rb, rc, ra, 7);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) { mbar_arrive(a_full(grp)); mbar_arrive(patch_empty(ps)); }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace sep2d
}  // namespace bq
