// C ABI of the stain normaliser (standalone entry point; the model path calls the same kernels from model.cu).
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/biscuit_b200.h"
#include "common.cuh"
#include "stain_sm100.cuh"

int bq_stain_launch(bq_ctx* ctx, const uint8_t* tiles_dev, int64_t n, int32_t px, const float* lut_dev, float* stats_dev,
                    const float target_means[3], const float target_stds[3], uint8_t* out_dev) {
  if (n <= 0) return BQ_OK;
  const int64_t ppt = (int64_t)px * px;
  bq::stain::lab_stats_kernel<<<(unsigned)n, 512, 0, ctx->stream>>>(tiles_dev, ppt, lut_dev, stats_dev);
  BQ_LAUNCH_CHECK(ctx);
  int64_t blocks = (n * ppt + 255) / 256;
  const int64_t cap = (int64_t)ctx->num_sms * 16;
  if (blocks > cap) blocks = cap;
  bq::stain::reinhard_apply_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(
      tiles_dev, out_dev, n, ppt, lut_dev, stats_dev, target_means[0], target_means[1], target_means[2], target_stds[0],
      target_stds[1], target_stds[2]);
  BQ_LAUNCH_CHECK(ctx);
  return BQ_OK;
}

extern "C" int bq_stain_normalize(bq_ctx* ctx, int32_t kind, const uint8_t* tiles, int64_t n, int32_t px,
                                  const float target_means[3], const float target_stds[3], uint8_t* out, float* lab_stats) {
  if (!ctx) return BQ_ERR_ARG;
  if (kind != BQ_NORM_REINHARD_FAST) return bq_fail(ctx, BQ_ERR_ARG, "bq_stain_normalize: unknown normaliser %d", kind);
  if (n < 0 || px <= 0 || !target_means || !target_stds || (n > 0 && (!tiles || !out)))
    return bq_fail(ctx, BQ_ERR_ARG, "bq_stain_normalize: bad argument");
  if (n == 0) return BQ_OK;
  BQ_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t bytes = (size_t)n * px * px * 3;
  DevBuf in, res, lut, stats;
  int rc;
  if ((rc = bq_to_device(ctx, in, tiles, bytes))) return rc;
  const bool out_dev = bq_is_device_ptr(out);
  if (out_dev) { res.p = out; res.bytes = bytes; res.owned = false; }
  else if ((rc = bq_alloc(ctx, res, bytes))) return rc;
  float h_lut[256];
  bq::stain::build_gamma_lut(h_lut);
  if ((rc = bq_alloc(ctx, lut, sizeof(h_lut))) || (rc = bq_alloc(ctx, stats, (size_t)n * 6 * sizeof(float)))) return rc;
  BQ_CUDA(ctx, cudaMemcpyAsync(lut.p, h_lut, sizeof(h_lut), cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = bq_stain_launch(ctx, (const uint8_t*)in.p, n, px, (const float*)lut.p, (float*)stats.p, target_means,
                            target_stds, (uint8_t*)res.p)))
    return rc;
  if (!out_dev && (rc = bq_from_device(ctx, out, res.p, bytes))) return rc;
  if (lab_stats && (rc = bq_from_device(ctx, lab_stats, stats.p, (size_t)n * 6 * sizeof(float)))) return rc;
  BQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // h_lut and the temporaries go out of scope
  return BQ_OK;
}
