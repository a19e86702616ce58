// Fused SeparableConv2D for the 728 -> 728 layers (block4_sepconv2, the 24 middle-flow sepconvs, block13_sepconv1 =
// 56 % of the network's FLOPs):   out = epilogue( depthwise3x3(relu?(x)) @ Wpw^T )
//
// The depthwise result is the A operand of the pointwise GEMM and is produced INSIDE the GEMM kernel: eight producer warps
// compute a 128-pixel x 64-channel k-block from a TMA-loaded activation patch (CUDA cores, packed FFMA2) and write it
// straight into the 128B-swizzled smem stage that tcgen05.mma consumes, so the depthwise output never exists in HBM/L2
// (it was one full tensor write + one full tensor read per layer) and the CUDA-core work overlaps the tensor-core work.
//
// TMEM holds 128 lanes x 512 fp32 columns, less than the 728 output channels of one pixel tile, so a work item is
// (128-pixel M tile, one of two N ranges [0,384) / [384,736)): two independent CTAs each produce the A tile (the depthwise
// is recomputed 2x -- it is 1.6 % of the FLOPs) and accumulate up to 384 output channels (one N=256 and one N=128/96 MMA
// per k-step).  The M tile is 128 consecutive pixels of the flattened [image, y, x] index; the patch is the same matrix
// from W+1 rows before to W+1 rows after (TMA zero-fills outside the tensor), and per-pixel row / column validity bits
// remove the taps that fall outside the image (the zero padding of 'same').
//
// Roles (448 threads): warp 0 TMA (patch + weight k-blocks), warp 1 TMEM + MMA issue, warps 2-5 epilogue (BN, residual by
// TMA, ReLU, bf16, TMA store), warps 6-13 depthwise producers.
#pragma once

#include "gemm_sm100.cuh"

namespace bq {
namespace sepf {

using namespace sm100;

constexpr int kN0 = 384;                         // first N range [0, 384); second [384, 736)
constexpr int kPatchRows = 208;                  // >= 128 + 2*(W+1) + 2 for W <= 37; multiple of 8
constexpr int kPatchBytes = kPatchRows * 128;    // 26 KB (one 64-channel k-block of the patch), multiple of 1024
constexpr int kABytes = 128 * 128;               // 16 KB
constexpr int kBBytes = 3 * 128 * 128;           // 48 KB: three TMA boxes of 128 weight rows
constexpr int kIOBytes = 128 * 128;              // residual-in / output staging chunk (128 rows x 64 cols)
constexpr int kStages = 2;
constexpr int kOffPatch = 0;
constexpr int kOffA = kOffPatch + kStages * kPatchBytes;
constexpr int kOffB = kOffA + kStages * kABytes;
constexpr int kOffIO = kOffB + kStages * kBBytes;
constexpr int kOffScale = kOffIO + 2 * kIOBytes;             // float scale[384], shift[384]
constexpr int kOffMask = kOffScale + 2 * kN0 * 4;            // uint8 validity bits per row [128]
constexpr int kOffBar = kOffMask + 128;
constexpr int kSmem = kOffBar + 256 + 1024;
constexpr int kThreads = 448;
constexpr int kProducerWarps = 8;

struct SepParams {
  int M;                    // pixels (rows) in this launch
  int H, W;                 // image size
  int C;                    // 728 (input == output channels)
  int relu_in;              // ReLU on the block input before the depthwise
  int relu_out;
  int has_res;
  const float* dw;          // [9][C] depthwise weights (bf16-rounded values, fp32)
  const float* scale;       // [C] folded BatchNorm
  const float* shift;
};

__global__ void __launch_bounds__(kThreads, 1)
sepconv_fused_kernel(const __grid_constant__ CUtensorMap tmap_x /*[M, C] box [64 x kPatchRows] SW128*/,
                     const __grid_constant__ CUtensorMap tmap_w /*[C, C] box [64 x 128] SW128*/,
                     const __grid_constant__ CUtensorMap tmap_out /*[M, C] box [64 x 128] SW128*/,
                     const __grid_constant__ CUtensorMap tmap_res, const SepParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar0 = smem_base + kOffBar;
  auto patch_full = [&](int s) { return bar0 + 8u * s; };
  auto patch_empty = [&](int s) { return bar0 + 8u * (2 + s); };
  auto a_full = [&](int s) { return bar0 + 8u * (4 + s); };
  auto a_empty = [&](int s) { return bar0 + 8u * (6 + s); };
  auto b_full = [&](int s) { return bar0 + 8u * (8 + s); };
  auto b_empty = [&](int s) { return bar0 + 8u * (10 + s); };
  const uint32_t acc_full = bar0 + 8u * 12, acc_empty = bar0 + 8u * 13;
  auto res_bar = [&](int b) { return bar0 + 8u * (14 + b); };
  const uint32_t tmem_slot = bar0 + 8u * 16;
  volatile uint32_t* tmem_slot_ptr = (volatile uint32_t*)(smem_gen + kOffBar + 8 * 16);
  float* s_scale = (float*)(smem_gen + kOffScale);
  float* s_shift = s_scale + kN0;
  uint8_t* s_mask = smem_gen + kOffMask;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (p.M + 127) / 128;
  const int n_items = m_tiles * 2;
  const int num_kb = (p.C + 63) / 64;                    // 12 (last k-block: 24 valid channels)
  const int halo = p.W + 1;                              // rows before / after the tile that the taps reach

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_out);
    if (p.has_res) tma_prefetch_desc(&tmap_res);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(patch_full(s), 1); mbar_init(patch_empty(s), kProducerWarps);
      mbar_init(a_full(s), kProducerWarps); mbar_init(a_empty(s), 1);
      mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1);
    }
    mbar_init(acc_full, 1); mbar_init(acc_empty, 4);
    mbar_init(res_bar(0), 1); mbar_init(res_bar(1), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA: activation patch + weight k-blocks =====================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int m0 = (it >> 1) * 128, n0 = (it & 1) * kN0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(patch_empty(s), ph ^ 1u);
          mbar_expect_tx(patch_full(s), (uint32_t)kPatchBytes);
          tma_load_2d(smem_base + kOffPatch + s * kPatchBytes, &tmap_x, patch_full(s), kb * 64, m0 - halo);
          mbar_wait(b_empty(s), ph ^ 1u);
          mbar_expect_tx(b_full(s), (uint32_t)kBBytes);
          for (int b = 0; b < 3; ++b)
            tma_load_2d(smem_base + kOffB + s * kBBytes + b * 128 * 128, &tmap_w, b_full(s), kb * 64, n0 + b * 128);
          if (++s == kStages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc256 = make_idesc(128, 256);
      int s = 0; uint32_t ph = 0, acc_ph = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const uint32_t idesc2 = make_idesc(128, (it & 1) ? 96 : 128);       // [256,384) or [640,736)
        mbar_wait(acc_empty, acc_ph ^ 1u);
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(a_full(s), ph);
          mbar_wait(b_full(s), ph);
          tc_fence_after();
          const uint64_t da = make_smem_desc<128>(smem_base + kOffA + s * kABytes);
          const uint64_t db0 = make_smem_desc<128>(smem_base + kOffB + s * kBBytes);
          const uint64_t db1 = make_smem_desc<128>(smem_base + kOffB + s * kBBytes + 256 * 128);
          int ksteps = 4;
          if (kb == num_kb - 1) ksteps = (p.C - kb * 64 + 15) / 16;
          for (int k = 0; k < ksteps; ++k) {
            umma_bf16(tmem_base, da + (uint64_t)(2 * k), db0 + (uint64_t)(2 * k), idesc256, (kb | k) ? 1u : 0u);
            umma_bf16(tmem_base + 256u, da + (uint64_t)(2 * k), db1 + (uint64_t)(2 * k), idesc2, (kb | k) ? 1u : 0u);
          }
          umma_commit(a_empty(s));
          umma_commit(b_empty(s));
          if (kb == num_kb - 1) umma_commit(acc_full);
          if (++s == kStages) { s = 0; ph ^= 1u; }
        }
        acc_ph ^= 1u;
      }
    }
  } else if (warp < 6) {
    // ===================== epilogue: BN, residual, ReLU, bf16, TMA store =====================
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    const int etid = threadIdx.x - 64;
    const bool leader = (warp == 2 && lane == 0);
    uint32_t acc_ph = 0, cc = 0;
    int loaded_n0 = -1;
    // residual prefetch cursor (leader only): chunk order = (item, 64-column chunk), 6 chunks per item
    int pf_it = blockIdx.x, pf_c = 0;
    uint32_t pf_cc = 0;
    auto prefetch_res = [&]() {
      if (!p.has_res || pf_it >= n_items) return;
      const uint32_t b = pf_cc & 1u;
      mbar_expect_tx(res_bar(b), (uint32_t)kIOBytes);
      tma_load_2d(smem_base + kOffIO + b * kIOBytes, &tmap_res, res_bar(b), (pf_it & 1) * kN0 + pf_c, (pf_it >> 1) * 128);
      ++pf_cc;
      pf_c += 64;
      if (pf_c >= kN0) { pf_c = 0; pf_it += gridDim.x; }
    };
    if (leader) prefetch_res();
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const int m0 = (it >> 1) * 128, n0 = (it & 1) * kN0;
      if (n0 != loaded_n0) {                       // per-channel BN constants of this N range -> smem
        asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int i = etid; i < kN0; i += 128) {
          const int c = n0 + i;
          s_scale[i] = c < p.C ? __ldg(p.scale + c) : 0.f;
          s_shift[i] = c < p.C ? __ldg(p.shift + c) : 0.f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        loaded_n0 = n0;
      }
      mbar_wait(acc_full, acc_ph);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16);
      for (int c = 0; c < kN0; c += 64, ++cc) {
        const uint32_t buf = cc & 1u;
        uint8_t* io = smem_gen + kOffIO + buf * kIOBytes;
        uint32_t v[64];
        {
          uint32_t (&v0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[0]);
          uint32_t (&v1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[32]);
          tmem_ld_32x32b_x32(t_row + (uint32_t)c, v0);
          tmem_ld_32x32b_x32(t_row + (uint32_t)c + 32u, v1);     // columns of channels >= C are stale: masked below
          tmem_ld_wait();
        }
        if (c + 64 >= kN0) {                                    // accumulators fully read: release them early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty);
        }
        // buffer protocol: chunk cc stages in buffer cc & 1 (residual in, result out, in place).  Before chunk cc+1's
        // residual may land in the OTHER buffer, the store of chunk cc-1 must have finished reading it -- it has had the
        // whole previous chunk to drain, so the leader checks that first and prefetches one chunk ahead.
        if (leader) {
          tma_store_wait_read1();                               // store of chunk cc-2 (this buffer) has drained; cc-1 may be in flight
          if (p.has_res) { tma_store_wait_read0(); prefetch_res(); }   // the residual of chunk cc+1 lands in chunk cc-1's buffer
        }
        if (p.has_res) {
          mbar_wait(res_bar(buf), (cc >> 1) & 1u);
        } else {
          asm volatile("bar.sync 1, 128;" ::: "memory");        // leader has confirmed the buffer is free
        }
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const int cl = c + g * 8;                              // column inside this N range
          const uint32_t sw_off = (uint32_t)row_in_tile * 128u + (uint32_t)((g ^ (row_in_tile & 7)) << 4);
          float f[8];
          if (n0 + cl < p.C) {
            const float4 s0 = *(const float4*)(s_scale + cl), s1 = *(const float4*)(s_scale + cl + 4);
            const float4 h0 = *(const float4*)(s_shift + cl), h1 = *(const float4*)(s_shift + cl + 4);
            const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
            const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = __fadd_rn(__fmul_rn(__uint_as_float(v[g * 8 + j]), sc[j]), sh[j]);
            if (p.has_res) {
              const uint4 r = *(const uint4*)(io + sw_off);
              const __nv_bfloat162* rb = (const __nv_bfloat162*)&r;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 rf = __bfloat1622float2(rb[j]);
                f[2 * j] = __fadd_rn(f[2 * j], rf.x); f[2 * j + 1] = __fadd_rn(f[2 * j + 1], rf.y);
              }
            }
            if (p.relu_out) {
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = 0.f;
          }
          uint4 o;
          __nv_bfloat162* ob = (__nv_bfloat162*)&o;
#pragma unroll
          for (int j = 0; j < 4; ++j) ob[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
          *(uint4*)(io + sw_off) = o;
        }
        fence_async_smem();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (leader) {
          tma_store_2d(&tmap_out, smem_base + kOffIO + buf * kIOBytes, n0 + c, m0);    // columns >= C clipped by the TMA unit
          tma_store_commit();
        }
      }
      acc_ph ^= 1u;
    }
    if (leader) tma_store_wait_all();
  } else {
    // ===================== depthwise producers (8 warps): patch -> A k-block =====================
    const int ptid = threadIdx.x - 192;                         // 0..255
    const int c4 = ptid & 15;                                   // 4-channel group inside the 64-channel k-block
    const int r0 = (ptid >> 4) * 8;                             // this thread's 8 consecutive tile rows
    const int chunk16 = c4 >> 1, sub8 = (c4 & 1) * 8;           // 16-byte chunk / 8-byte half inside the 128-byte row
    const int hw = p.H * p.W;
    int s = 0; uint32_t ph = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const int m0 = (it >> 1) * 128;
      // validity bits of this item's 128 pixels: bits 0..2 = row y-1,y,y+1 inside the image, bits 3..5 = column x-1,x,x+1
      asm volatile("bar.sync 2, 256;" ::: "memory");             // everyone done with the previous item's masks
      if (ptid < 128) {
        const int m = m0 + ptid;
        uint32_t bits = 0;
        if (m < p.M) {
          const int rem = m % hw, y = rem / p.W, x = rem - y * p.W;
          bits = (y > 0 ? 1u : 0u) | 2u | (y + 1 < p.H ? 4u : 0u) | (x > 0 ? 8u : 0u) | 16u | (x + 1 < p.W ? 32u : 0u);
        }
        s_mask[ptid] = (uint8_t)bits;
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");
      uint32_t rowbits[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) rowbits[i] = s_mask[r0 + i];
      for (int kb = 0; kb < num_kb; ++kb) {
        const int c = kb * 64 + c4 * 4;
        mbar_wait(patch_full(s), ph);
        mbar_wait(a_empty(s), ph ^ 1u);
        const uint8_t* patch = smem_gen + kOffPatch + s * kPatchBytes;
        float2 acc[8][2];
#pragma unroll
        for (int i = 0; i < 8; ++i) { acc[i][0] = make_float2(0.f, 0.f); acc[i][1] = make_float2(0.f, 0.f); }
        if (c < p.C) {
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            // weights of the three taps of this kernel row, 4 channels each
            float2 w[3][2];
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              const float4 wv = __ldg((const float4*)(p.dw + (size_t)(dy * 3 + dx) * p.C + c));
              w[dx][0] = make_float2(wv.x, wv.y);
              w[dx][1] = make_float2(wv.z, wv.w);
            }
            // ten consecutive patch rows cover columns x-1..x+1 of all eight pixels of this thread
            const int pr0 = r0 + halo + (dy - 1) * p.W - 1;      // >= 0
            float2 v[10][2];
#pragma unroll
            for (int j = 0; j < 10; ++j) {
              const int pr = pr0 + j;
              const uint2 raw = *(const uint2*)(patch + (size_t)pr * 128 + ((chunk16 ^ (pr & 7)) << 4) + sub8);
              float2 a = make_float2(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xFFFF0000u));
              float2 b = make_float2(__uint_as_float(raw.y << 16), __uint_as_float(raw.y & 0xFFFF0000u));
              if (p.relu_in) {
                a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); b.x = fmaxf(b.x, 0.f); b.y = fmaxf(b.y, 0.f);
              }
              v[j][0] = a; v[j][1] = b;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const bool yok = (rowbits[i] >> dy) & 1u;
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
                if (yok && ((rowbits[i] >> (3 + dx)) & 1u)) {
                  acc[i][0] = __ffma2_rn(v[i + dx][0], w[dx][0], acc[i][0]);
                  acc[i][1] = __ffma2_rn(v[i + dx][1], w[dx][1], acc[i][1]);
                }
              }
            }
          }
        }
        uint8_t* a_dst = smem_gen + kOffA + s * kABytes;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = r0 + i;
          uint2 o;
          __nv_bfloat162* ob = (__nv_bfloat162*)&o;
          ob[0] = __floats2bfloat162_rn(acc[i][0].x, acc[i][0].y);
          ob[1] = __floats2bfloat162_rn(acc[i][1].x, acc[i][1].y);
          *(uint2*)(a_dst + (size_t)r * 128 + ((chunk16 ^ (r & 7)) << 4) + sub8) = o;
        }
        fence_async_smem();                                     // generic-proxy writes -> visible to tcgen05.mma
        __syncwarp();
        if (lane == 0) { mbar_arrive(a_full(s)); mbar_arrive(patch_empty(s)); }
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace sepf
}  // namespace bq
