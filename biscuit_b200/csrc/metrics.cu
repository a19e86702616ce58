// Evaluation-metric kernels next to the hot path: the data-parallel parts of
// reference biscuit/utils.py:400-464 (`prediction_metrics`: 500 bootstrap confusion matrices of 150 samples) and
// biscuit/delong.py:36-73 (`fastDeLong`: per-example placement values of DeLong's AUC variance).
// Everything that is order-sensitive floating point (np.cov, statistics.mean / variance, scipy's normal quantiles) stays
// on the host in the wrapper, computed by the same library calls as the reference, so results are bit-identical.
#include <cuda_runtime.h>

#include <vector>

#include "../../include/biscuit_b200.h"
#include "common.cuh"

namespace {

// One thread per bootstrap replicate b: confusion counts of the S sampled rows idx[b][0..S).
__global__ void bootstrap_confusion_kernel(const uint8_t* __restrict__ y_true, const uint8_t* __restrict__ y_pred_bin,
                                           const int64_t* __restrict__ idx, int64_t n, int32_t n_boot, int32_t n_samp,
                                           int64_t* __restrict__ out /*[n_boot][4] = tp, fp, tn, fn*/, int* __restrict__ bad) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_boot) return;
  int64_t tp = 0, fp = 0, tn = 0, fn = 0;
  for (int s = 0; s < n_samp; ++s) {
    const int64_t i = idx[(int64_t)b * n_samp + s];
    if (i < 0 || i >= n) { atomicExch(bad, 1); continue; }
    const bool t = y_true[i] != 0, p = y_pred_bin[i] != 0;
    tp += t && p; fp += !t && p; tn += !t && !p; fn += t && !p;
  }
  out[b * 4 + 0] = tp; out[b * 4 + 1] = fp; out[b * 4 + 2] = tn; out[b * 4 + 3] = fn;
}

// DeLong placement values.  For a positive example i:  tz_i - tx_i = #{neg < s_i} + 0.5 #{neg == s_i}  (midrank among all
// minus midrank among positives, delong.py:63-67), v01_i = that / n_neg.  For a negative j: v10_j = 1 - (#{pos < s_j} +
// 0.5 #{pos == s_j}) / n_pos.  tz_sum accumulates the all-sample midranks of the positives (half-integers: the fp64 sum
// is exact in any order).  One warp per example, lanes stride over the other examples.
template <typename T>
__global__ void delong_placement_kernel(const T* __restrict__ score, const uint8_t* __restrict__ label, int64_t n,
                                        int64_t n_pos, int64_t n_neg, double* __restrict__ v, double* __restrict__ tz_sum) {
  const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const double si = (double)score[i];
  const bool pos = label[i] != 0;
  long long less_other = 0, eq_other = 0, less_all = 0, eq_all = 0;
  for (int64_t j = lane; j < n; j += 32) {
    const double sj = (double)score[j];
    const bool other = (label[j] != 0) != pos;
    const bool lt = sj < si, eq = sj == si;
    less_all += lt; eq_all += eq;
    less_other += lt && other; eq_other += eq && other;
  }
  for (int o = 16; o > 0; o >>= 1) {
    less_other += __shfl_down_sync(0xffffffffu, less_other, o);
    eq_other += __shfl_down_sync(0xffffffffu, eq_other, o);
    less_all += __shfl_down_sync(0xffffffffu, less_all, o);
    eq_all += __shfl_down_sync(0xffffffffu, eq_all, o);
  }
  if (lane == 0) {
    const double d = (double)less_other + 0.5 * (double)eq_other;      // exact (half-integer)
    v[i] = pos ? d / (double)n_neg : 1.0 - d / (double)n_pos;
    if (pos) atomicAdd(tz_sum, (double)less_all + 0.5 * ((double)eq_all + 1.0));   // midrank, 1-based (delong.py:25-27)
  }
}

}  // namespace

extern "C" int bq_bootstrap_confusion(bq_ctx* ctx, const uint8_t* y_true, const uint8_t* y_pred_bin, int64_t n,
                                      const int64_t* idx, int32_t n_boot, int32_t n_samp, int64_t* counts) {
  if (!ctx) return BQ_ERR_ARG;
  if (n <= 0 || n_boot <= 0 || n_samp <= 0 || !y_true || !y_pred_bin || !idx || !counts)
    return bq_fail(ctx, BQ_ERR_ARG, "bq_bootstrap_confusion: bad argument");
  BQ_CUDA(ctx, cudaSetDevice(ctx->device));
  DevBuf yt, yp, ix, out, bad;
  int rc;
  if ((rc = bq_to_device_pooled(ctx, yt, y_true, (size_t)n)) || (rc = bq_to_device_pooled(ctx, yp, y_pred_bin, (size_t)n)) ||
      (rc = bq_to_device_pooled(ctx, ix, idx, (size_t)n_boot * n_samp * 8)) ||
      (rc = bq_alloc_pooled(ctx, out, (size_t)n_boot * 4 * 8)) || (rc = bq_alloc_pooled(ctx, bad, 4)))
    return rc;
  BQ_CUDA(ctx, cudaMemsetAsync(bad.p, 0, 4, ctx->stream));
  bootstrap_confusion_kernel<<<(n_boot + 127) / 128, 128, 0, ctx->stream>>>(
      (const uint8_t*)yt.p, (const uint8_t*)yp.p, (const int64_t*)ix.p, n, n_boot, n_samp, (int64_t*)out.p, (int*)bad.p);
  BQ_LAUNCH_CHECK(ctx);
  int h_bad = 0;
  if ((rc = bq_from_device(ctx, counts, out.p, (size_t)n_boot * 4 * 8)) || (rc = bq_from_device(ctx, &h_bad, bad.p, 4))) return rc;
  BQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (h_bad) return bq_fail(ctx, BQ_ERR_ARG, "bq_bootstrap_confusion: sample index out of range");
  return BQ_OK;
}

extern "C" int bq_delong_placements(bq_ctx* ctx, const void* score, int32_t dtype, const uint8_t* label, int64_t n,
                                    double* v, double* tz_pos_sum, int64_t* n_pos_out) {
  if (!ctx) return BQ_ERR_ARG;
  if (n <= 0 || !score || !label || !v || !tz_pos_sum || (dtype != BQ_F32 && dtype != BQ_F64))
    return bq_fail(ctx, BQ_ERR_ARG, "bq_delong_placements: bad argument");
  if (n > (int64_t)1 << 20)
    return bq_fail(ctx, BQ_ERR_ARG, "bq_delong_placements: n = %lld exceeds 2^20 (all-pairs kernel; slide / patient level only)",
                   (long long)n);
  BQ_CUDA(ctx, cudaSetDevice(ctx->device));
  // the class counts are needed as divisors: count on the host copy or bring the labels back once
  DevBuf s, l, vd, tz;
  int rc;
  const size_t es = dtype == BQ_F32 ? 4 : 8;
  if ((rc = bq_to_device_pooled(ctx, s, score, (size_t)n * es)) || (rc = bq_to_device_pooled(ctx, l, label, (size_t)n)) ||
      (rc = bq_alloc_pooled(ctx, vd, (size_t)n * 8)) || (rc = bq_alloc_pooled(ctx, tz, 8)))
    return rc;
  std::vector<uint8_t> h_label((size_t)n);
  BQ_CUDA(ctx, cudaMemcpyAsync(h_label.data(), l.p, (size_t)n, cudaMemcpyDefault, ctx->stream));
  BQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  int64_t n_pos = 0;
  for (uint8_t x : h_label) n_pos += x != 0;
  const int64_t n_neg = n - n_pos;
  if (n_pos_out) *n_pos_out = n_pos;
  if (n_pos == 0 || n_neg == 0) return bq_fail(ctx, BQ_ERR_ARG, "bq_delong_placements: labels must contain both classes");
  BQ_CUDA(ctx, cudaMemsetAsync(tz.p, 0, 8, ctx->stream));
  const int64_t threads = n * 32;
  const unsigned blocks = (unsigned)((threads + 255) / 256);
  if (dtype == BQ_F32)
    delong_placement_kernel<float><<<blocks, 256, 0, ctx->stream>>>((const float*)s.p, (const uint8_t*)l.p, n, n_pos, n_neg,
                                                                    (double*)vd.p, (double*)tz.p);
  else
    delong_placement_kernel<double><<<blocks, 256, 0, ctx->stream>>>((const double*)s.p, (const uint8_t*)l.p, n, n_pos, n_neg,
                                                                     (double*)vd.p, (double*)tz.p);
  BQ_LAUNCH_CHECK(ctx);
  if ((rc = bq_from_device(ctx, v, vd.p, (size_t)n * 8)) || (rc = bq_from_device(ctx, tz_pos_sum, tz.p, 8))) return rc;
  BQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return BQ_OK;
}
