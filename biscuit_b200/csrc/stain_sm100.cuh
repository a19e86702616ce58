// Reinhard-fast stain normalisation (the pre-processing step in front of per-image standardisation:
// reference biscuit/hp.py:19 `normalizer='reinhard_fast'`, results.py:251-254 `wsi_normalizer.rgb_to_rgb`).
// The arithmetic lives in Slideflow (slideflow>=1.1.0rc1, slideflow/norm/tensorflow/{reinhard,color}.py -- third-party,
// not vendored); restated from its published algorithm (oracle/reinhard.py holds the CPU restatement):
//   lab = rgb_to_lab(u8 / 255)                     sRGB gamma -> XYZ (D65) -> CIE L*a*b*, all fp32
//   per tile and LAB channel: mean, population std over the 299 x 299 pixels
//   lab' = (lab - mean) * (target_std / std) + target_mean
//   out = clip(int32(lab_to_rgb(lab') * 255), 0, 255) as uint8     (the cast truncates toward zero)
// "fast" = no brightness standardisation (percentile rescale) in front.
//
// Two HBM-bound passes over 268 KB per tile (0.5 % of the step's traffic): lab_stats_kernel (one block per tile, fp64
// block reduction) and reinhard_apply_kernel (one thread per pixel).  The sRGB -> linear gamma is a 256-entry table
// (inputs are bytes); the inverse gamma and the cube roots are evaluated per pixel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bq {
namespace stain {

struct Lab { float L, a, b; };

__device__ __forceinline__ float f_xyz(float t) {            // color.py: epsilon = 6/29
  const float eps = 6.0f / 29.0f;
  return t <= eps * eps * eps ? __fadd_rn(__fdiv_rn(t, 3.0f * eps * eps), 4.0f / 29.0f) : cbrtf(t);
}

// linear RGB (already de-gamma'd through the table) -> LAB; products and sums in the matmul's row order
__device__ __forceinline__ Lab linear_to_lab(float r, float g, float b) {
  const float x = __fadd_rn(__fadd_rn(__fmul_rn(r, 0.412453f), __fmul_rn(g, 0.357580f)), __fmul_rn(b, 0.180423f));
  const float y = __fadd_rn(__fadd_rn(__fmul_rn(r, 0.212671f), __fmul_rn(g, 0.715160f)), __fmul_rn(b, 0.072169f));
  const float z = __fadd_rn(__fadd_rn(__fmul_rn(r, 0.019334f), __fmul_rn(g, 0.119193f)), __fmul_rn(b, 0.950227f));
  const float fx = f_xyz(__fmul_rn(x, 1.0f / 0.950456f)), fy = f_xyz(y), fz = f_xyz(__fmul_rn(z, 1.0f / 1.088754f));
  Lab o;
  o.L = __fadd_rn(__fmul_rn(fy, 116.0f), -16.0f);
  o.a = __fadd_rn(__fmul_rn(fx, 500.0f), __fmul_rn(fy, -500.0f));
  o.b = __fadd_rn(__fmul_rn(fy, 200.0f), __fmul_rn(fz, -200.0f));
  return o;
}

__device__ __forceinline__ float inv_f_xyz(float f) {
  const float eps = 6.0f / 29.0f;
  return f <= eps ? __fmul_rn(3.0f * eps * eps, __fadd_rn(f, -4.0f / 29.0f)) : __fmul_rn(__fmul_rn(f, f), f);
}

__device__ __forceinline__ float gamma_encode(float v) {      // linear [0,1] -> sRGB [0,1]
  v = fminf(fmaxf(v, 0.0f), 1.0f);
  return v <= 0.0031308f ? __fmul_rn(v, 12.92f) : __fadd_rn(__fmul_rn(powf(v, 1.0f / 2.4f), 1.055f), -0.055f);
}

__device__ __forceinline__ uint8_t to_u8(float srgb) {
  int v = (int)__fmul_rn(srgb, 255.0f);                       // tf.cast(float -> int32) truncates
  return (uint8_t)min(max(v, 0), 255);
}

__device__ __forceinline__ void lab_to_u8(float L, float a, float b, uint8_t* out) {
  const float fy = __fmul_rn(__fadd_rn(L, 16.0f), 1.0f / 116.0f);
  const float fx = __fadd_rn(fy, __fmul_rn(a, 1.0f / 500.0f));
  const float fz = __fadd_rn(fy, __fmul_rn(b, -1.0f / 200.0f));
  const float x = __fmul_rn(inv_f_xyz(fx), 0.950456f), y = inv_f_xyz(fy), z = __fmul_rn(inv_f_xyz(fz), 1.088754f);
  const float r = __fadd_rn(__fadd_rn(__fmul_rn(x, 3.2404542f), __fmul_rn(y, -1.5371385f)), __fmul_rn(z, -0.4985314f));
  const float g = __fadd_rn(__fadd_rn(__fmul_rn(x, -0.9692660f), __fmul_rn(y, 1.8760108f)), __fmul_rn(z, 0.0415560f));
  const float bl = __fadd_rn(__fadd_rn(__fmul_rn(x, 0.0556434f), __fmul_rn(y, -0.2040259f)), __fmul_rn(z, 1.0572252f));
  out[0] = to_u8(gamma_encode(r));
  out[1] = to_u8(gamma_encode(g));
  out[2] = to_u8(gamma_encode(bl));
}

// stats[tile] = {mean L, mean a, mean b, std L, std a, std b} (population std).  One block per tile.
static __global__ void __launch_bounds__(512)
lab_stats_kernel(const uint8_t* __restrict__ tiles, int64_t px_per_tile, const float* __restrict__ gamma_lut,
                 float* __restrict__ stats) {
  __shared__ float lut[256];
  __shared__ double red[6][16];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = gamma_lut[i];
  __syncthreads();
  const uint8_t* src = tiles + (int64_t)blockIdx.x * px_per_tile * 3;
  double s[6] = {0, 0, 0, 0, 0, 0};
  for (int64_t p = threadIdx.x; p < px_per_tile; p += blockDim.x) {
    const Lab v = linear_to_lab(lut[src[3 * p]], lut[src[3 * p + 1]], lut[src[3 * p + 2]]);
    s[0] += v.L; s[1] += v.a; s[2] += v.b;
    s[3] += (double)v.L * v.L; s[4] += (double)v.a * v.a; s[5] += (double)v.b * v.b;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    double v = s[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) red[k][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    const int c = threadIdx.x;
    double sum = 0, sq = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { sum += red[c][w]; sq += red[3 + c][w]; }
    const double mean = sum / (double)px_per_tile;
    double var = sq / (double)px_per_tile - mean * mean;
    if (var < 0) var = 0;
    stats[(int64_t)blockIdx.x * 6 + c] = (float)mean;
    stats[(int64_t)blockIdx.x * 6 + 3 + c] = (float)sqrt(var);
  }
}

static __global__ void __launch_bounds__(256)
reinhard_apply_kernel(const uint8_t* __restrict__ tiles, uint8_t* __restrict__ out, int64_t n_tiles, int64_t px_per_tile,
                      const float* __restrict__ gamma_lut, const float* __restrict__ stats, float tm0, float tm1,
                      float tm2, float ts0, float ts1, float ts2) {
  __shared__ float lut[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = gamma_lut[i];
  __syncthreads();
  const int64_t total = n_tiles * px_per_tile;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = p / px_per_tile;
    const float* st = stats + t * 6;
    const Lab v = linear_to_lab(lut[tiles[3 * p]], lut[tiles[3 * p + 1]], lut[tiles[3 * p + 2]]);
    // (I - mean) * (target_std / std) + target_mean, in that association (reinhard.py: transform)
    const float L = __fadd_rn(__fmul_rn(__fadd_rn(v.L, -st[0]), __fdiv_rn(ts0, st[3])), tm0);
    const float a = __fadd_rn(__fmul_rn(__fadd_rn(v.a, -st[1]), __fdiv_rn(ts1, st[4])), tm1);
    const float b = __fadd_rn(__fmul_rn(__fadd_rn(v.b, -st[2]), __fdiv_rn(ts2, st[5])), tm2);
    uint8_t o[3];
    lab_to_u8(L, a, b, o);
    out[3 * p] = o[0]; out[3 * p + 1] = o[1]; out[3 * p + 2] = o[2];
  }
}

// host: the 256-entry sRGB -> linear table, float32 arithmetic as in color.py (x/255, then the two branches)
inline void build_gamma_lut(float lut[256]) {
  for (int i = 0; i < 256; ++i) {
    const float x = (float)i / 255.0f;
    lut[i] = x <= 0.04045f ? x / 12.92f : powf((x + 0.055f) / 1.055f, 2.4f);
  }
}

}  // namespace stain
}  // namespace bq
