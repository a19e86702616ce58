// Hardware probe (developer hook, not on the product path): checks how tcgen05.mma addresses an A operand whose
// smem start is shifted by whole rows inside a 128B-swizzled tile (descriptor base_offset semantics) and the
// no-swizzle K-major B layout.  Used to decide whether depthwise 3x3 / 3x3 convolutions can run as shifted-window
// MMAs on one resident halo tile.
#include <cuda.h>

#include "common.cuh"
#include "gemm_sm100.cuh"

namespace {
using namespace bq::sm100;

// D[128 x 16] = A[128 x 16] * B^T, A = rows [shift, shift+128) x channels [cg*16, +16) of a [rows x 64] bf16 tile that a
// TMA load placed in smem with SWIZZLE_128B; B = 16x16 identity in the no-swizzle K-major layout.
template <int SWZ>
__global__ void __launch_bounds__(128, 1)
umma_shift_probe_kernel(const __grid_constant__ CUtensorMap tmap_x, int rows, int shift, int cg, int base_offset_mode,
                        float* __restrict__ out /*[128][16]*/) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t tile = smem_base;                              // rows x 128 B (rows <= 384)
  const uint32_t bmat = smem_base + 384 * 128;                  // 512 B identity
  const uint32_t bar = bmat + 1024, mma_bar = bar + 8, tmem_slot = bar + 16;
  volatile uint32_t* tmem_slot_ptr = (volatile uint32_t*)(smem_gen + 384 * 128 + 1024 + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // identity B, no-swizzle K-major: element (n,k) at (n/8)*256 + (k/8)*128 + (n%8)*16 + (k%8)*2
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    const int n = i / 16, k = i % 16;
    __nv_bfloat16* p = (__nv_bfloat16*)(smem_gen + 384 * 128 + (n / 8) * 256 + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2);
    *p = __float2bfloat16_rn(n == k ? 1.0f : 0.0f);
  }
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(mma_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, 32);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, (uint32_t)rows * (uint32_t)SWZ);
    for (int r0 = 0; r0 < rows; r0 += 128) tma_load_2d(tile + r0 * SWZ, &tmap_x, bar, 0, r0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t a_addr = tile + (uint32_t)shift * (uint32_t)SWZ + (uint32_t)cg * 32u;
    uint64_t da = make_smem_desc<SWZ>(a_addr);
    if (base_offset_mode == 1) da |= (uint64_t)((a_addr >> 7) & 7u) << 49;
    // B: layout none, LBO = 128 B (next 8 k), SBO = 256 B (next 8 n)
    const uint64_t db = (uint64_t)((bmat & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
    umma_bf16(tmem_base, da, db, make_idesc(128, 16), 0u);
    umma_commit(mma_bar);
  }
  mbar_wait(mma_bar, 0);
  tc_fence_after();
  uint32_t v[32];
  tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld_wait();
  for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 16 + j] = __uint_as_float(v[j]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 32); }
}
}  // namespace

extern "C" int bq_debug_umma_probe(bq_ctx* ctx, int rows, int shift, int cg, int base_offset_mode, const uint16_t* x_bf16,
                                   float* out) {
  // base_offset_mode bit 1 selects the 64-byte-swizzle variant (rows of 32 bf16)
  const int swz = (base_offset_mode & 2) ? 64 : 128;
  const int cols = swz / 2;
  base_offset_mode &= 1;
  if (!ctx || !x_bf16 || !out || rows < 128 || rows > 384 || rows % 128 || shift < 0 || shift > 256 || cg < 0 || cg * 16 >= cols)
    return bq_fail(ctx, BQ_ERR_ARG, "bq_debug_umma_probe: bad argument");
  BQ_CUDA(ctx, cudaSetDevice(ctx->device));
  DevBuf x, o;
  int rc;
  if ((rc = bq_to_device(ctx, x, x_bf16, (size_t)rows * cols * 2)) || (rc = bq_alloc(ctx, o, 128 * 16 * 4))) return rc;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn)
    return bq_fail(ctx, BQ_ERR_CUDA, "no cuTensorMapEncodeTiled");
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)swz};
  cuuint32_t box[2] = {(cuuint32_t)cols, 128};
  cuuint32_t es[2] = {1, 1};
  if (((Enc)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, x.p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return bq_fail(ctx, BQ_ERR_CUDA, "tensor map encode failed");
  const int smem = 384 * 128 + 2048 + 1024;
  cudaFuncSetAttribute(umma_shift_probe_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(umma_shift_probe_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (swz == 128) umma_shift_probe_kernel<128><<<1, 128, smem, ctx->stream>>>(tm, rows, shift, cg, base_offset_mode, (float*)o.p);
  else umma_shift_probe_kernel<64><<<1, 128, smem, ctx->stream>>>(tm, rows, shift, cg, base_offset_mode, (float*)o.p);
  BQ_LAUNCH_CHECK(ctx);
  BQ_CUDA(ctx, cudaMemcpyAsync(out, o.p, 128 * 16 * 4, cudaMemcpyDeviceToHost, ctx->stream));
  BQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return BQ_OK;
}
