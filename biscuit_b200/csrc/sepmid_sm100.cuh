// Fused SeparableConv2D for the 728 -> 728 layers on the 19 x 19 maps (the 24 middle-flow layers: 56 % of the
// network's FLOPs).  The depthwise result never exists in HBM and nothing is computed twice:
//
//     out[p, n] = epilogue( sum_c  depthwise3x3(relu?(x))[p, c] * Wpw[n, c] )        (+ residual[p, n])
//
// Two ideas make it fit the machine.
//
// (1) ZERO-PADDED FLATTENED LAYOUT.  The activations of these layers live in dedicated buffers with one zero column
//     and one zero row per image: pixel (y, x) of image b is row  b*400 + y*20 + x  of a [rows, 728] bf16 matrix, and
//     rows with x == 19 or y == 19 are zero and are never written.  In that layout the 3 x 3 'same' depthwise is nine
//     FIXED row offsets {-21,-20,-19,-1,0,+1,+19,+20,+21} with no border predicate anywhere (the neighbour of a border
//     pixel is a zero row), the halo of ANY run of consecutive rows is one 2-D TMA box, and a work item need not align
//     with an image.  Cost: 400 / 361 = 1.108 x the rows.
//
// (2) THE GEMM IS TRANSPOSED so the CTA pair shares the depthwise output in hardware.  728 fp32 accumulator columns do
//     not fit the 512 TMEM columns of one SM, and an N split across CTAs would need the produced A tile copied through
//     DSMEM (21 B/clk).  Instead the WEIGHTS are the M-side operand and the PIXELS the N-side operand of a
//     cta_group::2 MMA:   D^T[256 channels, 160 pixels] += W[256, 64] * dw[160, 64]^T.   With cta_group::2 each CTA
//     supplies HALF of the N-side tile from its own shared memory -- exactly the 80 pixels its own producer warps
//     computed -- and the tensor core reads both halves for both CTAs.  Three channel tiles (3 x 256 >= 728) keep
//     3 x 160 = 480 TMEM columns per CTA, so every depthwise k-block is produced once and used for all 728 outputs.
//
// Work item = 160 consecutive padded rows (8 padded image rows) per CTA pair; per 64-channel k-block and CTA:
//   warp 0        TMA: the 128 x 64 weight tiles (this CTA's half of each 256-row channel tile), a ring of five single
//                 tiles in exactly the order the MMAs consume them
//   warp 14       TMA: the (80 + 42)-row input window of this CTA's 80 pixels (its own warp: behind the weight ring's
//                 waits the window requests were issued late)
//   warps 10-13   depthwise producers: thread = 4 channels x (2 image rows x 5 columns); the 4 x 7 halo is read once
//                 (28 LDS.64 + 180 FFMA2 per 10 outputs, same tap order as depthwise3x3_pipe_kernel) -> bf16 ->
//                 this CTA's half of the 128B-swizzled N-side stage (4 stages)
//   warp 1        (leader CTA) 3 x 4 tcgen05.mma.cta_group::2 (M = 256, N = 160, K = 16) per k-block
//   warps 2-9     epilogue, eight independent warp pipelines on accumulator fragments: tcgen05.ld.16x256b -> BN
//                 scale/shift (four constant pairs per thread) (+ residual, fetched by the warp's own TMA load and added
//                 with FHADD.BF16) -> cvt.rn(.relu) pack -> stmatrix.trans into the warp's [40 px][32 ch] staging tile
//                 -> TMA stores of the VALID pixels only (one 19-pixel box per image row): the zero border is never touched.
// The accumulators are single-buffered (480 of 512 columns), so a channel tile cannot take the next item's MMAs before
// the epilogue has read it.  The three channel tiles are SKEWED by one k-block: in step t the MMA warp issues
// (tile 0, k-block t), (tile 1, k-block t-1), (tile 2, k-block t-2) of one continuous k-block stream over all items, so
// the three drains fall in different steps and a depthwise k-block lives for three steps.  Measured, the skew alone
// buys nothing (one in-order issuing thread still queues the runnable tiles behind a waiting one);
// it is kept as the precondition for per-tile issuers.
#pragma once

#include "gemm_sm100.cuh"

namespace bq {
namespace sepmid {

using namespace sm100;

constexpr int kC = 728;                         // channels in and out
constexpr int kMap = 19;                        // valid map size
constexpr int kPitch = 20;                      // padded row pitch (19 valid + 1 zero column)
constexpr int kImgRows = kPitch * kPitch;       // 400 padded rows per image
constexpr int kItemPx = 160;                    // rows per work item (per CTA pair)
constexpr int kCtaPx = 80;                      // rows produced per CTA
constexpr int kWinRows = kCtaPx + 2 * (kPitch + 1);   // 122: halo of 21 rows on either side
constexpr int kWinBytes = kWinRows * 128;       // 15,616
constexpr int kNumKb = 12;                      // ceil(728 / 64); the last k-block holds 24 channels (2 k-steps)
#ifndef BQ_SM_WST
#define BQ_SM_WST 5          // weight ring depth in TILES (128 x 64 bf16 = 16 KB per CTA and stage)
#endif
#ifndef BQ_SM_BST
#define BQ_SM_BST 4          // depthwise-output stages: a k-block lives for three MMA steps (the channel tiles are skewed)
#endif
constexpr int kWStages = BQ_SM_WST;
constexpr int kWTile = 128 * 128;               // 128 weight rows x 64 k
#ifndef BQ_SM_IST
#define BQ_SM_IST 2
#endif
#ifndef BQ_SM_EST
#define BQ_SM_EST 3
#endif
constexpr int kInStages = BQ_SM_IST;
constexpr int kBStages = BQ_SM_BST;
constexpr int kBBytes = kCtaPx * 128;           // 10,240 (multiple of 1024)
constexpr int kStepPx = 40;                     // epilogue step: 40 pixels x 128 channels per CTA
constexpr int kOutTile = kStepPx * 64;          // one epilogue warp's tile of a step: [40 px][32 ch] bf16
constexpr int kOutStep = 8 * kOutTile;          // the eight epilogue warps
constexpr int kOffW = 0;
constexpr int kOffB = kOffW + kWStages * kWTile;
constexpr int kOffOut = kOffB + kBStages * kBBytes;
constexpr int kEpiBufs = BQ_SM_EST;                     // rotating epilogue buffers: residual lands / in-place epilogue / store drains
constexpr int kOffIn = kOffOut + kEpiBufs * kOutStep;
constexpr int kOffBar = kOffIn + kInStages * kWinBytes;
constexpr int kSmem = kOffBar + 512 + 1024;     // + slack for the 1024 B alignment of the base
// developer tunables (profiles/build_variants.py builds A/B copies of the library with -D overrides)
constexpr int kProducerWarps = 4;               // 128 threads = 16 channel groups x (2 image-row pairs x 4 column strips)
constexpr int kColsPerStrip = 5;                // 4 strips x 5 columns = the 20-pixel padded row
constexpr int kThreads = 32 * (3 + 8 + kProducerWarps);   // warp 0 weight TMA, 1 MMA, 2-9 epilogue, 10.. producers, last: window TMA (registers are granted per 4 warps: 19 cost as 20)
constexpr int kEpiWarps = 8;
constexpr int kSlackRows = 2 * kItemPx;         // rows allocated past the last image (the last item may overhang)

struct SepMidParams {
  int n_rows;               // 400 * images in this launch
  int relu_out;
  const float* dw;          // [9][728] depthwise taps (bf16-rounded values held in fp32)
  const float* scale;       // [728] folded BatchNorm
  const float* shift;
  const bf16* residual;     // padded layout, nullable
};

// Cross-CTA hand-off of the depthwise stage (peer CTA's generic-proxy smem writes -> tcgen05.mma issued by the leader):
// the writer orders its stores ahead of async-proxy reads with fence.proxy.async.shared::cta (MEMBAR.ALL.CTA +
// FENCE.VIEW.ASYNC.S) and then arrives on the leader's mbarrier.  Cluster-scope release/acquire qualifiers are NOT used:
// ptxas lowers them to MEMBAR.ALL.GPU on every producer thread and CCTL.IVALL (an L1 flush) on the waiting MMA thread,
// per k-block -- measured 1.8x slower than the separate kernels.  Shared memory is not cached anywhere, the CTA-level
// membar completes the stores before the arrive is sent, and the arrive reaches the leader strictly later.

// Issue order inside MMA step t (shared by the TMA and the MMA warp): slot j -> channel tile, or -1.  Tile ct works on
// stream position t - ct; the one group that STARTS an item (k-block 0: it must wait for the epilogue) is issued last so
// that the other two never queue behind that wait.
__device__ __forceinline__ int step_tile(int t, int j, int U) {
  int first = -1;                                     // the tile whose k-block is 0 in this step, if any
#pragma unroll
  for (int ct = 0; ct < 3; ++ct) {
    const int u = t - ct;
    if (u >= 0 && u < U && u % kNumKb == 0) first = ct;
  }
  int n = 0;
#pragma unroll
  for (int ct = 0; ct < 3; ++ct) {
    const int u = t - ct;
    if (u < 0 || u >= U || ct == first) continue;
    if (n == j) return ct;
    ++n;
  }
  return (first >= 0 && n == j) ? first : -1;
}

template <bool RELU_IN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
sepconv_mid_kernel(const __grid_constant__ CUtensorMap tmap_in /*[rows, 728] box [122 x 64], no swizzle*/,
                   const __grid_constant__ CUtensorMap tmap_w /*[728, 728] box [128 x 64], SW128*/,
                   const __grid_constant__ CUtensorMap tmap_out /*[rows, 728] box [19 x 32], no swizzle*/,
                   const __grid_constant__ CUtensorMap tmap_res /*[rows, 728] box [40 x 32], no swizzle (residual, padded layout)*/,
                   const SepMidParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar0 = smem_base + kOffBar;
  constexpr int oWE = kWStages, oIF = 2 * kWStages, oIE = oIF + kInStages, oBF = oIE + kInStages, oBE = oBF + kBStages,
                oAF = oBE + kBStages, oAE = oAF + 3, oRF = oAE + 3, oTM = oRF + 8 * kEpiBufs;
  static_assert(8 * (oTM + 1) <= 512, "barrier area");
  auto w_full = [&](int s) { return bar0 + 8u * s; };                  // leader: 2 arrivals + tx
  auto w_empty = [&](int s) { return bar0 + 8u * (oWE + s); };         // both CTAs, multicast commit
  auto in_full = [&](int s) { return bar0 + 8u * (oIF + s); };         // local
  auto in_empty = [&](int s) { return bar0 + 8u * (oIE + s); };
  auto b_full = [&](int s) { return bar0 + 8u * (oBF + s); };          // leader: 2 x producer-warp arrivals
  auto b_empty = [&](int s) { return bar0 + 8u * (oBE + s); };         // both CTAs
  auto acc_full = [&](int t) { return bar0 + 8u * (oAF + t); };        // both CTAs
  auto acc_empty = [&](int t) { return bar0 + 8u * (oAE + t); };       // leader: 2 x epilogue-warp arrivals
  auto res_full = [&](int w8, int e) { return bar0 + 8u * (oRF + w8 * kEpiBufs + e); };   // local, per epilogue warp: residual tile landed
  const uint32_t tmem_slot = bar0 + 8u * oTM;
  volatile uint32_t* tmem_slot_ptr = (volatile uint32_t*)(smem_gen + kOffBar + 8 * oTM);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool is_leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int n_items = (p.n_rows + kItemPx - 1) / kItemPx;
  const int my_items = cluster_id < n_items ? (n_items - cluster_id + num_clusters - 1) / num_clusters : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_in);
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_out);
    tma_prefetch_desc(&tmap_res);
    for (int s = 0; s < kWStages; ++s) { mbar_init(w_full(s), 2); mbar_init(w_empty(s), 1); }
    for (int s = 0; s < kInStages; ++s) { mbar_init(in_full(s), 1); mbar_init(in_empty(s), kProducerWarps); }
    for (int s = 0; s < kBStages; ++s) { mbar_init(b_full(s), 2 * kProducerWarps); mbar_init(b_empty(s), 1); }
    for (int t = 0; t < 3; ++t) { mbar_init(acc_full(t), 1); mbar_init(acc_empty(t), 2 * kEpiWarps); }
    for (int w8 = 0; w8 < kEpiWarps; ++w8)
      for (int e = 0; e < kEpiBufs; ++e) mbar_init(res_full(w8, e), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync_all();
  if (warp == 1) tmem_alloc_2cta(tmem_slot, 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);   // warp-uniform for the compiler

  if (warp == 0) {
    // ===================== TMA: input window of this CTA's 80 pixels + its 128 rows of each weight tile =====================
    // (converged warp, one elected lane issues: TMA instructions take uniform-register operands like tcgen05.mma)
    {
      int ws = 0; uint32_t wph = 0;
      const int U = my_items * kNumKb;                 // length of this cluster's k-block stream
      for (int t = 0; t < U + 2; ++t) {
        // weight tiles in MMA issue order (see the MMA warp): the group that starts an item goes last
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int ct = step_tile(t, j, U);
          if (ct < 0) continue;
          const int kb = (t - ct) % kNumKb;
          mbar_wait(w_empty(ws), wph ^ 1u);
          if (elect_one()) {
            if (is_leader) mbar_expect_tx(w_full(ws), 2u * kWTile);
            else mbar_arrive_remote(w_full(ws), 0);
            tma_load_2d_2cta(smem_base + kOffW + ws * kWTile, &tmap_w, w_full(ws), kb * 64, ct * 256 + (int)rank * 128);
          }
          __syncwarp();
          if (++ws == kWStages) { ws = 0; wph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: leader CTA; the whole warp walks the loop, one elected lane issues =====================
    if (is_leader) {
      const uint32_t idesc = make_idesc(256, kItemPx);
      int ws = 0; uint32_t wph = 0;
      const int U = my_items * kNumKb;
#ifdef BQ_SM_DIAG_STALL
      long long st_b = 0, st_w = 0, st_a = 0;
      const long long st_t0 = clock64();
#endif
      for (int t = 0; t < U + 2; ++t) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int ct = step_tile(t, j, U);
          if (ct < 0) continue;
          const int u = t - ct;                        // position in the k-block stream
          const int li = u / kNumKb, kb = u - li * kNumKb;
          const int bs = u % kBStages;
#ifdef BQ_SM_DIAG_STALL         // TIMING DIAGNOSTIC ONLY: where the MMA issuer waits
          const long long c0 = clock64();
          mbar_wait(b_full(bs), (uint32_t)(u / kBStages) & 1u);
          const long long c1 = clock64();
          mbar_wait(w_full(ws), wph);
          const long long c2 = clock64();
          if (kb == 0) mbar_wait(acc_empty(ct), ((uint32_t)li & 1u) ^ 1u);
          const long long c3 = clock64();
          st_b += c1 - c0; st_w += c2 - c1; st_a += c3 - c2;
#else
          mbar_wait(b_full(bs), (uint32_t)(u / kBStages) & 1u);
          mbar_wait(w_full(ws), wph);
          if (kb == 0) mbar_wait(acc_empty(ct), ((uint32_t)li & 1u) ^ 1u);   // the epilogue has handed this tile back
#endif
          tc_fence_after();
          const uint64_t db = make_smem_desc<128>(smem_base + kOffB + bs * kBBytes);
          const uint64_t da = make_smem_desc<128>(smem_base + kOffW + ws * kWTile);
          const uint32_t d = tmem_base + (uint32_t)(ct * kItemPx);
          if (elect_one()) {
            umma_bf16_2cta(d, da, db, idesc, kb ? 1u : 0u);
            umma_bf16_2cta(d, da + 2u, db + 2u, idesc, 1u);
            if (kb != kNumKb - 1) {                    // the last k-block holds 24 channels: two k-steps
              umma_bf16_2cta(d, da + 4u, db + 4u, idesc, 1u);
              umma_bf16_2cta(d, da + 6u, db + 6u, idesc, 1u);
            }
            umma_commit_2cta(w_empty(ws));
            if (ct == 2) umma_commit_2cta(b_empty(bs));          // tile 2 is the last reader of a depthwise k-block
            if (kb == kNumKb - 1) umma_commit_2cta(acc_full(ct));
          }
          __syncwarp();
          if (++ws == kWStages) { ws = 0; wph ^= 1u; }
        }
      }
#ifdef BQ_SM_DIAG_STALL
      if (lane == 0 && (cluster_id == 0 || cluster_id == 37))
        printf("sepmid stall cluster %d items %d: total %lld  b_full %lld  w_full %lld  acc_empty %lld\n", cluster_id, my_items,
               clock64() - st_t0, st_b, st_w, st_a);
#endif
    }
  } else if (warp < 2 + kEpiWarps) {
    // ===================== epilogue: 128 channels (TMEM lanes) x 160 pixels (columns) per channel tile =====================
    // Eight INDEPENDENT warp pipelines (no block-level barrier, no single store leader): warp = (TMEM lane quadrant: 32
    // channels, pixel half: 80 columns).  A step = this warp's [40 px][32 ch] tile, finished IN PLACE in the warp's
    // buffer step % 3: the residual tile of the step was prefetched there by this warp's own TMA load, the lanes
    // overwrite it with the output, one elected lane stores it (one 19-pixel box per image row) and refills the buffer
    // whose stores have drained.  ncu / timing diagnostics: with a 256-thread barrier and one leader issuing all stores
    // per step the epilogue cost 1/3 of the kernel (the next item's MMAs wait for the channel tiles to be handed back).
    const int quad = warp & 3;                        // TMEM lane quadrant -> channels quad*32 .. +31 of this CTA's 128
    const int hh = (warp - 2) >> 2;                   // pixel half: columns hh*80 .. +79
    const int w8 = warp - 2;
    const bool has_res = p.residual != nullptr;
    const uint32_t total_steps = (uint32_t)my_items * 6u;
    const uint32_t my_out = smem_base + kOffOut + w8 * (kEpiBufs * kOutTile);
    auto step_coords = [&](uint32_t gg, int& row0, int& col0) {     // warp-local step index -> first row / first channel
      const uint32_t li2 = gg / 6u, r6 = gg % 6u;
      row0 = (cluster_id + (int)li2 * num_clusters) * kItemPx + hh * kCtaPx + (int)(r6 & 1u) * kStepPx;
      col0 = (int)(r6 >> 1) * 256 + (int)rank * 128 + quad * 32;
    };
    auto prefetch_res = [&](uint32_t gg) {                          // one elected lane
      if (gg >= total_steps) return;
      int row0, col0;
      step_coords(gg, row0, col0);
      const uint32_t e = gg % kEpiBufs;
      mbar_expect_tx(res_full(w8, e), (uint32_t)kOutTile);
      tma_load_2d(my_out + e * kOutTile, &tmap_res, res_full(w8, e), col0, row0);
    };
    if (has_res && elect_one()) {
      for (uint32_t i = 0; i + 1 < (uint32_t)kEpiBufs; ++i) prefetch_res(i);
    }
    __syncwarp();
    uint32_t g = 0;                                   // running step counter of this warp
    for (int li = 0; li < my_items; ++li) {
#pragma unroll
      for (int ct = 0; ct < 3; ++ct) {
        mbar_wait(acc_full(ct), (uint32_t)li & 1u);
        tc_fence_after();
#ifdef BQ_SM_DIAG_NOEPI          // TIMING DIAGNOSTIC ONLY: the epilogue only hands the tile back
        if (true) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (is_leader) mbar_arrive(acc_empty(ct)); else mbar_arrive_remote(acc_empty(ct), 0); }
          g += 2;
        } else
#endif
        {
          // Fragment epilogue.  Thread t holds channels cq = t/4 + {0, 8, 16, 24} of the warp's 32 and pixel pairs
          // 2(t%4) of every 8-pixel block: BN constants are four per channel tile, a pixel pair packs into one bf16x2
          // register (cvt.rn.relu fuses the ReLU), and ONE stmatrix.x4.trans writes an [8 px][32 ch] block of the
          // staging tile (64-byte rows, 16-byte chunks XOR-swizzled with (row >> 1) & 3 == TMA SWIZZLE_64B).
          float s4[4], h4[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int c = ct * 256 + (int)rank * 128 + quad * 32 + (lane >> 2) + 8 * k;
            s4[k] = c < kC ? __ldg(p.scale + c) : 0.f;
            h4[k] = c < kC ? __ldg(p.shift + c) : 0.f;
          }
          const uint32_t t_lo = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ct * kItemPx + hh * kCtaPx);
          const uint32_t t_hi = t_lo + (16u << 16);
          const int mrow = lane & 7, mmat = lane >> 3;             // this lane's row address duty for ld/stmatrix
          uint32_t pkv[20];
          auto load_step = [&](int s2, uint32_t (&lo)[20], uint32_t (&hi)[20]) {
            const uint32_t col = (uint32_t)(s2 * kStepPx);
            tmem_ld_16x256b_x4(t_lo + col, &lo[0]);
            tmem_ld_16x256b_x1(t_lo + col + 32u, &lo[16]);
            tmem_ld_16x256b_x4(t_hi + col, &hi[0]);
            tmem_ld_16x256b_x1(t_hi + col + 32u, &hi[16]);
            tmem_ld_wait();
          };
          auto finish2 = [&](const uint32_t (&lo)[20], const uint32_t (&hi)[20], uint32_t gg) {
            const uint32_t e = gg % kEpiBufs;
            if (has_res) mbar_wait(res_full(w8, e), (gg / kEpiBufs) & 1u);
#pragma unroll
            for (int b = 0; b < 5; ++b) {
              float f[8];
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                f[2 * k] = __fadd_rn(__fmul_rn(__uint_as_float(lo[4 * b + 2 * k]), s4[k]), h4[k]);
                f[2 * k + 1] = __fadd_rn(__fmul_rn(__uint_as_float(lo[4 * b + 2 * k + 1]), s4[k]), h4[k]);
                f[4 + 2 * k] = __fadd_rn(__fmul_rn(__uint_as_float(hi[4 * b + 2 * k]), s4[2 + k]), h4[2 + k]);
                f[4 + 2 * k + 1] = __fadd_rn(__fmul_rn(__uint_as_float(hi[4 * b + 2 * k + 1]), s4[2 + k]), h4[2 + k]);
              }
              if (has_res) {
                const int r = 8 * b + mrow;
                uint32_t q0, q1, q2, q3;
                ldmatrix_x4_trans(my_out + e * kOutTile + r * 64 + ((mmat ^ ((r >> 1) & 3)) << 4), q0, q1, q2, q3);
                const uint32_t qq[4] = {q0, q1, q2, q3};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  add_bf16x2_to_f32(f[2 * k], f[2 * k + 1], qq[k]);
                }
              }
#pragma unroll
              for (int k = 0; k < 4; ++k) pkv[4 * b + k] = pack_bf16x2(f[2 * k], f[2 * k + 1], p.relu_out != 0);
            }
          };
          auto store_step2 = [&](uint32_t gg) {
            const uint32_t e = gg % kEpiBufs;
            if (has_res) {
              __syncwarp();                                      // every lane's ldmatrix of this buffer precedes the overwrite
            } else {
              // buffer e was last stored from at step gg - 3: only now, two steps later, must that store have drained
              if (elect_one()) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kEpiBufs - 1) : "memory");
              __syncwarp();
            }
#pragma unroll
            for (int b = 0; b < 5; ++b) {
              const int r = 8 * b + mrow;
              stmatrix_x4_trans(my_out + e * kOutTile + r * 64 + ((mmat ^ ((r >> 1) & 3)) << 4), pkv[4 * b], pkv[4 * b + 1],
                                pkv[4 * b + 2], pkv[4 * b + 3]);
            }
            fence_async_smem();
            __syncwarp();
            if (elect_one()) {
              int row0, col0;
              step_coords(gg, row0, col0);
#pragma unroll
              for (int r2 = 0; r2 < 2; ++r2) {
                const int row = row0 + r2 * kPitch;
                const int y = (row / kPitch) % kPitch;
#ifndef BQ_SM_DIAG_NOSTORE       // TIMING DIAGNOSTIC ONLY: no output traffic
                if (y != kMap && row < p.n_rows && col0 < kC)
                  tma_store_2d(&tmap_out, my_out + e * kOutTile + r2 * kPitch * 64, col0, row);
#endif
              }
              tma_store_commit();
              if (has_res) {                                     // the residual of step gg + kEpiBufs - 1 lands in the buffer step gg - 1 stored from
                tma_store_wait_read1();
                prefetch_res(gg + kEpiBufs - 1);
              }
            }
            __syncwarp();
          };
          uint32_t lo[20], hi[20];
          load_step(0, lo, hi);
          finish2(lo, hi, g);
          load_step(1, lo, hi);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (is_leader) mbar_arrive(acc_empty(ct)); else mbar_arrive_remote(acc_empty(ct), 0); }
          store_step2(g);                                      // (the channel tile was handed back before its first staging store)
          finish2(lo, hi, g + 1);
          store_step2(g + 1);
          g += 2;
        }
      }
    }
    if (elect_one()) tma_store_wait_all();
    __syncwarp();
  } else if (warp == 2 + kEpiWarps + kProducerWarps) {
    // ===================== TMA: input window of this CTA's 80 pixels, one per k-block of the stream =====================
    // (its own warp: behind the weight ring's waits the window requests were issued late and the producers starved)
    int is = 0; uint32_t iph = 0;
    const int U = my_items * kNumKb;
    for (int t = 0; t < U; ++t) {
      const int p0 = (cluster_id + (t / kNumKb) * num_clusters) * kItemPx;
      mbar_wait(in_empty(is), iph ^ 1u);
      if (elect_one()) {
#ifdef BQ_SM_DIAG_NOWIN       // TIMING DIAGNOSTIC ONLY (wrong results): no input-window traffic
        mbar_arrive(in_full(is));
#else
        mbar_expect_tx(in_full(is), (uint32_t)kWinBytes);
        tma_load_2d(smem_base + kOffIn + is * kWinBytes, &tmap_in, in_full(is), (t % kNumKb) * 64,
                    p0 + (int)rank * kCtaPx - (kPitch + 1));
#endif
      }
      __syncwarp();
      if (++is == kInStages) { is = 0; iph ^= 1u; }
    }
  } else {
    // ===================== depthwise producers =====================
    // This CTA's 80 rows are four padded image rows of 20 pixels.  A thread owns 4 channels x (2 image rows x 5 columns):
    // it walks the four input rows of its 4 x 7 halo once, every loaded pixel feeds up to six outputs (28 window loads
    // and unpacks per 10 outputs; the earlier one-row strips needed 42), and each output still accumulates its nine
    // taps top row first, left to right -- the order of depthwise3x3_pipe_kernel, so the results are bit-identical.
    // (Diagnostic builds: the producers' shared-memory loads and stores are free, their INSTRUCTIONS are what slows the
    // kernel -- without the depthwise math it is 22 % faster, without the loads or without the stores not at all.)
    const int ptid = threadIdx.x - 32 * (2 + kEpiWarps);       // 0..127
    const int c4 = ptid & 15, sp = ptid >> 4;                  // channel group; strip 0..7
    const int rp = sp >> 2, cs = sp & 3;                       // image-row pair 0..1, column strip (columns 5 cs .. 5 cs + 4)
    const int wbase = (2 * rp) * kPitch + kColsPerStrip * cs;  // window row of this thread's halo corner (input row -1, column -1)
    const uint32_t total_kb = (uint32_t)my_items * kNumKb;
#ifdef BQ_SM_DIAG_NOLDS         // TIMING DIAGNOSTIC ONLY (wrong results): the producers do not read the window
    auto ldraw = [&](const uint8_t* rowp) { return make_uint2((uint32_t)(size_t)rowp * 2654435761u, (uint32_t)(size_t)rowp); };
#else
    auto ldraw = [&](const uint8_t* rowp) { return *(const uint2*)rowp; };
#endif
    auto unpack = [&](uint2 raw, float2 (&d)[2]) {
      if (RELU_IN) {
        const __nv_bfloat162 z2 = __floats2bfloat162_rn(0.f, 0.f);
        __nv_bfloat162* hb = (__nv_bfloat162*)&raw;
        hb[0] = __hmax2(hb[0], z2);
        hb[1] = __hmax2(hb[1], z2);
      }
      d[0] = make_float2(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xFFFF0000u));
      d[1] = make_float2(__uint_as_float(raw.y << 16), __uint_as_float(raw.y & 0xFFFF0000u));
    };
    for (uint32_t q = 0; q < total_kb; ++q) {
      const int kb = (int)(q % kNumKb);
      const int is = (int)(q % kInStages), bs = (int)(q % kBStages);
      const uint32_t iph = (q / kInStages) & 1u, bph = (q / kBStages) & 1u;
      const int c = kb * 64 + c4 * 4;
      const bool live = c < kC + 8;                            // the last k-block feeds 2 k-steps (channels 704..735)
      float2 w[9][2];
#pragma unroll
      for (int t = 0; t < 9; ++t) {                            // requested ahead of the two waits below
        float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < kC) wv = __ldg((const float4*)(p.dw + (size_t)t * kC + c));
        w[t][0] = make_float2(wv.x, wv.y);
        w[t][1] = make_float2(wv.z, wv.w);
      }
      mbar_wait(in_full(is), iph);
      mbar_wait(b_empty(bs), bph ^ 1u);
#ifdef BQ_SM_DIAG_NOPROD        // TIMING DIAGNOSTIC ONLY: producers skip the depthwise math
      if (false) {
#else
      if (live) {
#endif
        // window row of input (iy, ix) of this thread's halo: wbase + iy * 20 + ix   (the window starts 21 rows before the CTA's first row)
        const uint8_t* win = smem_gen + kOffIn + is * kWinBytes + (size_t)wbase * 128 + c4 * 8;
        uint8_t* bdst = smem_gen + kOffB + bs * kBBytes;
        float2 acc[2][kColsPerStrip][2];
#pragma unroll
        for (int ry = 0; ry < 2; ++ry)
#pragma unroll
          for (int cx = 0; cx < kColsPerStrip; ++cx) acc[ry][cx][0] = acc[ry][cx][1] = make_float2(0.f, 0.f);
#pragma unroll
        for (int iy = 0; iy < 4; ++iy) {
          float2 row[kColsPerStrip + 2][2];
#pragma unroll
          for (int ix = 0; ix < kColsPerStrip + 2; ++ix) unpack(ldraw(win + (size_t)(iy * kPitch + ix) * 128), row[ix]);
#pragma unroll
          for (int ry = 0; ry < 2; ++ry) {
            const int ky = iy - ry;                            // this input row is filter row ky of output row ry
            if (ky < 0 || ky > 2) continue;
#pragma unroll
            for (int cx = 0; cx < kColsPerStrip; ++cx)
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                acc[ry][cx][0] = __ffma2_rn(row[cx + kx][0], w[ky * 3 + kx][0], acc[ry][cx][0]);
                acc[ry][cx][1] = __ffma2_rn(row[cx + kx][1], w[ky * 3 + kx][1], acc[ry][cx][1]);
              }
          }
        }
#pragma unroll
        for (int ry = 0; ry < 2; ++ry)
#pragma unroll
          for (int cx = 0; cx < kColsPerStrip; ++cx) {
            const int pr = (2 * rp + ry) * kPitch + kColsPerStrip * cs + cx;   // row of the depthwise stage (this CTA's pixel)
            uint2 o;
            __nv_bfloat162* ob = (__nv_bfloat162*)&o;
            ob[0] = __floats2bfloat162_rn(acc[ry][cx][0].x, acc[ry][cx][0].y);
            ob[1] = __floats2bfloat162_rn(acc[ry][cx][1].x, acc[ry][cx][1].y);
#ifdef BQ_SM_DIAG_NOSTS         // TIMING DIAGNOSTIC ONLY (wrong results): the producers do not write the depthwise stage
            if (o.x == 0x12345678u && o.y == 0x9abcdef0u)
#endif
            *(uint2*)(bdst + (size_t)pr * 128 + (((c4 >> 1) ^ (pr & 7)) << 4) + (c4 & 1) * 8) = o;
          }
      }
      fence_async_smem();           // generic smem writes -> async proxy (tensor core), and window reads -> next TMA fill
      __syncwarp();
      if (lane == 0) {
        if (is_leader) mbar_arrive(b_full(bs)); else mbar_arrive_remote(b_full(bs), 0);
        mbar_arrive(in_empty(is));
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// layout converters: [n, 19, 19, 728] <-> padded [n * 400, 728] (valid pixels only; the zero border is never written)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pad_copy_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int n_img, int to_padded) {
  constexpr int cv = kC / 8;                                   // 91 x 16 B per pixel
  const int64_t total = (int64_t)n_img * kMap * kMap * cv;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(idx % cv);
    int64_t pix = idx / cv;
    const int x = (int)(pix % kMap);
    pix /= kMap;
    const int y = (int)(pix % kMap);
    const int img = (int)(pix / kMap);
    const int64_t dense = (((int64_t)img * kMap + y) * kMap + x) * kC + c8 * 8;
    const int64_t padded = ((int64_t)img * kImgRows + y * kPitch + x) * kC + c8 * 8;
    if (to_padded) *(uint4*)(out + padded) = __ldg((const uint4*)(in + dense));
    else *(uint4*)(out + dense) = __ldg((const uint4*)(in + padded));
  }
}

}  // namespace sepmid
}  // namespace bq
