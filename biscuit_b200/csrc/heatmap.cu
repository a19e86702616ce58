// Whole-slide tile-grid heat map: scatter of the per-tile MC-dropout predictions into the slide's grid and the
// uncertainty mask (reference results.py:216-227: `hm = sf.Heatmap(slide, model)`, `uq_mask = hm.uncertainty[:, :, 0] >
// thresh`, `hm.logits[uq_mask, :] = [-1, -1]`).  Bandwidth-trivial, but it keeps the per-tile results on the device between
// `bq_predict_uq` and the masked grid when the caller holds device buffers.
#include "common.cuh"

namespace {

__global__ void heatmap_fill_kernel(float* __restrict__ a, float* __restrict__ b, int64_t n, float v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    a[i] = v;
    b[i] = v;
  }
}

// one thread per (tile, class): grid cell (x, y) of the tile receives its mean / std
__global__ void heatmap_scatter_kernel(const float* __restrict__ mean, const float* __restrict__ stdv,
                                       const int32_t* __restrict__ grid_xy, int64_t n, int32_t n_classes, int32_t gx,
                                       float* __restrict__ logits, float* __restrict__ unc) {
  const int64_t total = n * n_classes;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = i / n_classes;
    const int c = (int)(i - t * n_classes);
    const int64_t cell = (int64_t)grid_xy[2 * t + 1] * gx + grid_xy[2 * t];
    logits[cell * n_classes + c] = mean[i];
    unc[cell * n_classes + c] = stdv[i];
  }
}

// mask[cell] = uncertainty[cell][0] > thresh (compared in float64; the caller applies NumPy's scalar promotion to
// `thresh`), masked cells get logits = -1 in every class
__global__ void heatmap_mask_kernel(const float* __restrict__ unc, int64_t cells, int32_t n_classes, double thresh,
                                    float* __restrict__ logits, uint8_t* __restrict__ mask) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < cells; i += (int64_t)gridDim.x * blockDim.x) {
    const bool m = (double)unc[i * n_classes] > thresh;
    mask[i] = m ? 1 : 0;
    if (m)
      for (int c = 0; c < n_classes; ++c) logits[i * n_classes + c] = -1.0f;
  }
}

int grid_for(int64_t total) {
  int64_t g = (total + 255) / 256;
  return (int)(g < 1 ? 1 : (g > 4096 ? 4096 : g));
}

}  // namespace

extern "C" {

int bq_heatmap_build(bq_ctx* ctx, int64_t n, int32_t n_classes, const float* mean, const float* stdv, const int32_t* grid_xy,
                     int32_t gx, int32_t gy, float* logits, float* uncertainty) {
  if (!ctx) return BQ_ERR_ARG;
  if (n < 0 || n_classes < 1 || gx < 0 || gy < 0 || (n > 0 && (!mean || !stdv || !grid_xy)) ||
      ((int64_t)gx * gy > 0 && (!logits || !uncertainty)))
    return bq_fail(ctx, BQ_ERR_ARG, "bq_heatmap_build: bad argument");
  BQ_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t cells = (int64_t)gx * gy, vals = cells * n_classes;
  if (vals == 0) return BQ_OK;
  DevBuf dm, ds, dg, dl, du;
  int rc;
  if ((rc = bq_to_device_pooled(ctx, dm, mean, (size_t)n * n_classes * 4)) || (rc = bq_to_device_pooled(ctx, ds, stdv, (size_t)n * n_classes * 4)) ||
      (rc = bq_to_device_pooled(ctx, dg, grid_xy, (size_t)n * 2 * 4)))
    return rc;
  const bool out_dev = bq_is_device_ptr(logits) && bq_is_device_ptr(uncertainty);
  float *pl = logits, *pu = uncertainty;
  if (!out_dev) {
    if ((rc = bq_alloc_pooled(ctx, dl, (size_t)vals * 4)) || (rc = bq_alloc_pooled(ctx, du, (size_t)vals * 4))) return rc;
    pl = (float*)dl.p;
    pu = (float*)du.p;
  }
  heatmap_fill_kernel<<<grid_for(vals), 256, 0, ctx->stream>>>(pl, pu, vals, -1.0f);
  BQ_LAUNCH_CHECK(ctx);
  if (n > 0) {
    heatmap_scatter_kernel<<<grid_for(n * n_classes), 256, 0, ctx->stream>>>((const float*)dm.p, (const float*)ds.p,
                                                                            (const int32_t*)dg.p, n, n_classes, gx, pl, pu);
    BQ_LAUNCH_CHECK(ctx);
  }
  if (!out_dev) {
    if ((rc = bq_from_device(ctx, logits, pl, (size_t)vals * 4)) || (rc = bq_from_device(ctx, uncertainty, pu, (size_t)vals * 4))) return rc;
  }
  BQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return BQ_OK;
}

int bq_heatmap_mask(bq_ctx* ctx, int64_t cells, int32_t n_classes, const float* uncertainty, double thresh, float* logits,
                    uint8_t* mask) {
  if (!ctx) return BQ_ERR_ARG;
  if (cells < 0 || n_classes < 1 || (cells > 0 && (!uncertainty || !logits || !mask)))
    return bq_fail(ctx, BQ_ERR_ARG, "bq_heatmap_mask: bad argument");
  if (cells == 0) return BQ_OK;
  BQ_CUDA(ctx, cudaSetDevice(ctx->device));
  DevBuf du, dl, dmk;
  int rc;
  const size_t vals = (size_t)cells * n_classes;
  if ((rc = bq_to_device_pooled(ctx, du, uncertainty, vals * 4)) || (rc = bq_to_device_pooled(ctx, dl, logits, vals * 4))) return rc;
  const bool mask_dev = bq_is_device_ptr(mask);
  uint8_t* pm = mask;
  if (!mask_dev) {
    if ((rc = bq_alloc_pooled(ctx, dmk, (size_t)cells))) return rc;
    pm = (uint8_t*)dmk.p;
  }
  heatmap_mask_kernel<<<grid_for(cells), 256, 0, ctx->stream>>>((const float*)du.p, cells, n_classes, thresh, (float*)dl.p, pm);
  BQ_LAUNCH_CHECK(ctx);
  if (!bq_is_device_ptr(logits) && (rc = bq_from_device(ctx, logits, dl.p, vals * 4))) return rc;
  if (!mask_dev && (rc = bq_from_device(ctx, mask, pm, (size_t)cells))) return rc;
  BQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return BQ_OK;
}

}  // extern "C"
