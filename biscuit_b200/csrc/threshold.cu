// UQ-thresholding kernels: tile pass, reference-order segmented slide reduction (K8), ROC / Youden (K9),
// group apply + confusion counts (K10).  All HBM-bound integer / compare / fp64-on-counts work.
//
// What each kernel reproduces (bit for bit) is restated library-free in oracle/threshold_oracle.py
// (tier 2) and follows: reference biscuit/threshold.py:125-245,297-348,411-460; sklearn 1.9.0
// metrics/_ranking.py:878-921,1020-1043,1317-1372 and :51-111; pandas 3.0.2 group_mean; numpy
// pairwise summation.
#include <cub/cub.cuh>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;

struct bq_table_impl {
  bq_ctx* ctx = nullptr;
  int64_t n = 0;
  int dtype = BQ_F32;
  DevBuf y_pred, unc, y_true, codes, incorrect;
  DevBuf perm, seg_begin, seg_end, sorted_codes;
  int32_t n_groups = 0;
  bool groups_ready = false;
  bool contiguous = true;
  bool filter_on = false;
  double tile_uq = 0.0;
  bool has_incorrect = false;
};

inline int grid_for(int64_t n, int threads, int num_sms, int per_sm = 16) {
  int64_t g = (n + threads - 1) / threads;
  int64_t cap = (int64_t)num_sms * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ---------------------------------------------------------------------------------------------
// validation + tile pass
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void validate_kernel(const T* __restrict__ yp, const T* __restrict__ unc,
                                const uint8_t* __restrict__ yt, int64_t n,
                                unsigned long long* __restrict__ flags) {
  unsigned long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    T p = yp[i], u = unc[i];
    c0 += (p != p);
    c1 += !isfinite((double)p);
    c2 += !isfinite((double)u);
    c3 += (yt[i] > 1);
  }
  for (int o = 16; o; o >>= 1) {
    c0 += __shfl_xor_sync(0xffffffffu, c0, o);
    c1 += __shfl_xor_sync(0xffffffffu, c1, o);
    c2 += __shfl_xor_sync(0xffffffffu, c2, o);
    c3 += __shfl_xor_sync(0xffffffffu, c3, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (c0) atomicAdd(flags + 0, c0);
    if (c1) atomicAdd(flags + 1, c1);
    if (c2) atomicAdd(flags + 2, c2);
    if (c3) atomicAdd(flags + 3, c3);
  }
}

template <typename T>
__global__ void tile_process_kernel(const T* __restrict__ yp, const uint8_t* __restrict__ yt, int64_t n,
                                    double thr, double* __restrict__ err, uint8_t* __restrict__ correct,
                                    uint8_t* __restrict__ incorrect, uint8_t* __restrict__ ybin) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    double p = (double)yp[i];
    int y = yt[i];
    bool lo = p < thr, hi = p >= thr;
    bool ok = (lo && y == 0) || (hi && y == 1);
    if (err) err[i] = fabs((double)y - p);
    if (correct) correct[i] = ok;
    incorrect[i] = !ok;
    if (ybin) ybin[i] = hi;
  }
}

// ---------------------------------------------------------------------------------------------
// K8: segmented, reference-order Kahan reduction
// ---------------------------------------------------------------------------------------------
// Run boundaries of the (possibly sorted) code array.  runs[c] counts how many separate runs code c
// has; >1 anywhere means the groups are not contiguous in row order and the sorted path is needed.
__global__ void seg_bounds_kernel(const int32_t* __restrict__ codes, int64_t n, int64_t* __restrict__ seg_begin,
                                  int64_t* __restrict__ seg_end, int32_t* __restrict__ runs,
                                  int32_t* __restrict__ noncontig) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int32_t c = codes[i];
    if (c < 0) continue;
    if (i == 0 || codes[i - 1] != c) {
      seg_begin[c] = i;
      if (atomicAdd(runs + c, 1) > 0) *noncontig = 1;
    }
    if (i == n - 1 || codes[i + 1] != c) seg_end[c] = i + 1;
  }
}

__global__ void iota_kernel(int32_t* p, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    p[i] = (int32_t)i;
}

__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }

// pandas group_mean step:  y = v - c; t = s + y; c = (t - s) - y; if (c != c) c = 0; s = t
template <typename T>
__device__ __forceinline__ void kahan_step(T& s, T& c, T v) {
  T y = add_rn(v, -c);
  T t = add_rn(s, y);
  T nc = add_rn(add_rn(t, -s), -y);
  c = (nc != nc) ? T(0) : nc;
  s = t;
}

// One warp per group: lanes fetch 32 rows coalesced (next chunk prefetched), then the warp replays the
// kept rows strictly in row order; every lane runs the same two Kahan chains (y_pred, uncertainty).
// NaN entries are skipped PER COLUMN with their own observation count, as pandas' group_mean does (a NaN
// uncertainty can only reach this kernel when the tile filter is off; NaN y_pred is rejected upstream).
template <typename T>
__global__ void __launch_bounds__(kThreads)
group_kahan_kernel(const T* __restrict__ yp, const T* __restrict__ unc, const uint8_t* __restrict__ yt,
                   const int32_t* __restrict__ perm, const int64_t* __restrict__ seg_begin,
                   const int64_t* __restrict__ seg_end, int filter_on, double tile_uq, int32_t n_groups,
                   T* __restrict__ g_pred, T* __restrict__ g_unc, double* __restrict__ g_true,
                   int64_t* __restrict__ g_count, int64_t* __restrict__ g_first) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int32_t g = blockIdx.x * warps_per_block + (threadIdx.x >> 5); g < n_groups;
       g += gridDim.x * warps_per_block) {
    const int64_t b = seg_begin[g], e = seg_end[g];
    T sp = 0, cp = 0, su = 0, cu = 0;
    int64_t cnt = 0, cnt_p = 0, cnt_u = 0, ysum = 0, first = -1;
    // prefetch chunk 0
    int64_t r = b + lane;
    bool valid = r < e;
    int64_t row = valid ? (perm ? (int64_t)perm[r] : r) : 0;
    T p = valid ? yp[row] : T(0);
    T u = valid ? unc[row] : T(0);
    int y = valid ? yt[row] : 0;
    for (int64_t base = b; base < e; base += 32) {
      // issue the next chunk's loads before the dependent arithmetic of this one
      int64_t r2 = base + 32 + lane;
      bool valid2 = r2 < e;
      int64_t row2 = valid2 ? (perm ? (int64_t)perm[r2] : r2) : 0;
      T p2 = valid2 ? yp[row2] : T(0);
      T u2 = valid2 ? unc[row2] : T(0);
      int y2 = valid2 ? yt[row2] : 0;

      bool keep = valid && (!filter_on || (double)u < tile_uq);
      unsigned m = __ballot_sync(0xffffffffu, keep);
      unsigned my = __ballot_sync(0xffffffffu, keep && y);
      const unsigned mp = __ballot_sync(0xffffffffu, keep && p == p);
      const unsigned mu = __ballot_sync(0xffffffffu, keep && u == u);
      cnt += __popc(m);
      cnt_p += __popc(mp);
      cnt_u += __popc(mu);
      ysum += __popc(my);
      if (first < 0 && m) first = __shfl_sync(0xffffffffu, row, __ffs(m) - 1);
      while (m) {
        int k = __ffs(m) - 1;
        m &= m - 1;
        T pv = __shfl_sync(0xffffffffu, p, k);
        T uv = __shfl_sync(0xffffffffu, u, k);
        if ((mp >> k) & 1u) kahan_step(sp, cp, pv);
        if ((mu >> k) & 1u) kahan_step(su, cu, uv);
      }
      valid = valid2; row = row2; p = p2; u = u2; y = y2;
    }
    if (lane == 0) {
      if (cnt > 0) {
        g_pred[g] = cnt_p > 0 ? div_rn(sp, (T)cnt_p) : (T)NAN;
        g_unc[g] = cnt_u > 0 ? div_rn(su, (T)cnt_u) : (T)NAN;
        g_true[g] = __ddiv_rn((double)ysum, (double)cnt);
      } else {
        g_pred[g] = (T)NAN;
        g_unc[g] = (T)NAN;
        g_true[g] = NAN;
      }
      g_count[g] = cnt;
      g_first[g] = first;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K9: ROC curve + Youden's J + AUC
// ---------------------------------------------------------------------------------------------
struct RocHeader {        // device-resident scalars shared by the pipeline stages
  unsigned long long n_eff;   // rows taking part (after `include`)
  long long m;                // distinct-score points (before pruning / prepending)
  long long n_pos;
  // results
  double threshold, youden_j, auc;
  long long n_points, best_index;
  int status, auc_exact;
};

template <typename T>
__global__ void roc_prep_kernel(const T* __restrict__ score, const uint8_t* __restrict__ label,
                                const uint8_t* __restrict__ include, int64_t n, T* __restrict__ keys,
                                uint8_t* __restrict__ vals, RocHeader* __restrict__ hdr) {
  unsigned long long c = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    bool inc = include ? include[i] != 0 : true;
    keys[i] = inc ? score[i] : -INFINITY;   // excluded rows sort to the tail (scores are finite)
    vals[i] = inc ? (label[i] == 1) : 0;
    c += inc;
  }
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&hdr->n_eff, c);
}

// packed[i] = (is_boundary << 32) | label  -> one inclusive scan yields both cumsum(label) and the
// output slot of every distinct-score boundary.
template <typename T>
__global__ void roc_pack_kernel(const T* __restrict__ keys, const uint8_t* __restrict__ vals, int64_t n,
                                const RocHeader* __restrict__ hdr, unsigned long long* __restrict__ packed) {
  const int64_t n_eff = (int64_t)hdr->n_eff;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    unsigned long long v = 0;
    if (i < n_eff) {
      bool boundary = (i == n_eff - 1) || (keys[i] != keys[i + 1]);
      v = ((unsigned long long)boundary << 32) | vals[i];
    }
    packed[i] = v;
  }
}

template <typename T>
__global__ void roc_scatter_kernel(const T* __restrict__ keys, const unsigned long long* __restrict__ scan,
                                   int64_t n, RocHeader* __restrict__ hdr, double* __restrict__ fps,
                                   double* __restrict__ tps, double* __restrict__ thr) {
  const int64_t n_eff = (int64_t)hdr->n_eff;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_eff;
       i += (int64_t)gridDim.x * blockDim.x) {
    unsigned long long s = scan[i];
    unsigned long long prev = i ? scan[i - 1] : 0ull;
    if ((s >> 32) != (prev >> 32)) {                 // row i closes a distinct-score run
      int64_t j = (int64_t)(s >> 32) - 1;
      double tp = (double)(uint32_t)s;               // cumsum(y)[i]          (_ranking.py:1034)
      tps[j] = tp;
      fps[j] = 1.0 + (double)i - tp;                 // 1 + idx - tps         (_ranking.py:1043)
      thr[j] = (double)keys[i];
    }
    if (i == n_eff - 1) {
      hdr->m = (long long)(s >> 32);
      hdr->n_pos = (long long)(uint32_t)s;
    }
  }
}

// keep flag of sklearn's drop_intermediate (second differences) + Youden J per point
__global__ void roc_points_kernel(const RocHeader* __restrict__ hdr, const double* __restrict__ fps,
                                  const double* __restrict__ tps, double* __restrict__ jval,
                                  uint8_t* __restrict__ keep) {
  const int64_t m = hdr->m;
  const double P = (double)hdr->n_pos, N = (double)((long long)hdr->n_eff - hdr->n_pos);
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < m;
       j += (int64_t)gridDim.x * blockDim.x) {
    bool k = true;
    if (m > 2 && j > 0 && j < m - 1) {
      double d2f = (fps[j + 1] - fps[j]) - (fps[j] - fps[j - 1]);
      double d2t = (tps[j + 1] - tps[j]) - (tps[j] - tps[j - 1]);
      k = (d2f != 0.0) || (d2t != 0.0);
    }
    keep[j] = k;
    jval[j] = k ? (__ddiv_rn(tps[j], P) - __ddiv_rn(fps[j], N)) : -INFINITY;
  }
}

// numpy's pairwise summation order, executed by one thread (iterative form of the recursion)
__device__ double np_pairwise_sum(const double* a, long long n) {
  struct Frame { long long lo, n; int state; double left; };
  Frame st[48];
  int sp = 0;
  st[0] = {0, n, 0, 0.0};
  double ret = 0.0;
  while (sp >= 0) {
    Frame& f = st[sp];
    if (f.state == 0) {
      if (f.n < 8) {
        double r = 0.0;
        for (long long i = 0; i < f.n; ++i) r = __dadd_rn(r, a[f.lo + i]);
        ret = r; --sp;
      } else if (f.n <= 128) {
        double r[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = a[f.lo + k];
        long long i = 8;
        for (; i < f.n - (f.n % 8); i += 8) {
#pragma unroll
          for (int k = 0; k < 8; ++k) r[k] = __dadd_rn(r[k], a[f.lo + i + k]);
        }
        double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                               __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
        for (; i < f.n; ++i) res = __dadd_rn(res, a[f.lo + i]);
        ret = res; --sp;
      } else {
        long long n2 = f.n / 2;
        n2 -= n2 % 8;
        f.state = 1;
        st[sp + 1] = {f.lo, n2, 0, 0.0};
        ++sp;
      }
    } else if (f.state == 1) {
      f.left = ret;
      long long n2 = f.n / 2;
      n2 -= n2 % 8;
      f.state = 2;
      st[sp + 1] = {f.lo + n2, f.n - n2, 0, 0.0};
      ++sp;
    } else {
      ret = __dadd_rn(f.left, ret);
      --sp;
    }
  }
  return __dadd_rn(0.0, ret);
}

constexpr int kFinalThreads = 1024;
constexpr long long kExactAucCap = 1 << 17;

// Single block: first-argmax of J (prepended (0,0,inf) point has J = 0 and index 0), then the AUC.
__global__ void __launch_bounds__(kFinalThreads)
roc_final_kernel(RocHeader* __restrict__ hdr, const double* __restrict__ fps, const double* __restrict__ tps,
                 const double* __restrict__ thr, const double* __restrict__ jval,
                 const uint8_t* __restrict__ keep, double* __restrict__ kf, double* __restrict__ kt,
                 double* __restrict__ terms) {
  __shared__ double s_j[kFinalThreads];
  __shared__ long long s_i[kFinalThreads];
  __shared__ long long s_base;
  __shared__ double s_red[kFinalThreads / 32];
  const int tid = threadIdx.x;
  const long long n_eff = (long long)hdr->n_eff;
  const long long m = hdr->m;
  const long long P = hdr->n_pos, N = n_eff - hdr->n_pos;

  if (n_eff == 0) {
    if (tid == 0) {
      hdr->status = 2; hdr->threshold = NAN; hdr->youden_j = NAN; hdr->auc = NAN;
      hdr->n_points = 0; hdr->best_index = -1; hdr->auc_exact = 1;
    }
    return;
  }
  const bool single = (P == 0) || (N == 0);

  // ---- Youden: lexicographic (max J, min index) over {-1 (prepended)} U kept points
  double bj = (tid == 0) ? 0.0 : -INFINITY;
  long long bi = (tid == 0) ? -1 : (1ll << 62);
  if (!single) {
    for (long long j = tid; j < m; j += kFinalThreads) {
      double v = jval[j];
      if (v > bj) { bj = v; bi = j; }
    }
  }
  s_j[tid] = bj; s_i[tid] = bi;
  __syncthreads();
  for (int o = kFinalThreads / 2; o; o >>= 1) {
    if (tid < o) {
      double vj = s_j[tid + o]; long long vi = s_i[tid + o];
      if (vj > s_j[tid] || (vj == s_j[tid] && vi < s_i[tid])) { s_j[tid] = vj; s_i[tid] = vi; }
    }
    __syncthreads();
  }
  const long long best = s_i[0];
  const double best_j = s_j[0];

  // ---- compaction of kept points (rank = position in the pruned curve, prepended point = 0)
  const bool exact = (m <= kExactAucCap);
  long long n_kept = 0;
  if (tid == 0) s_base = 0;
  __syncthreads();
  typedef cub::BlockScan<int, kFinalThreads> Scan;
  __shared__ typename Scan::TempStorage scan_tmp;
  __shared__ long long s_best_rank;
  if (tid == 0) s_best_rank = 0;
  for (long long c0 = 0; c0 < m; c0 += kFinalThreads) {
    long long j = c0 + tid;
    int k = (j < m) ? keep[j] : 0;
    int excl, total;
    Scan(scan_tmp).ExclusiveSum(k, excl, total);
    long long rank = s_base + excl;               // 0-based among kept points
    if (k) {
      if (exact) {
        kf[rank + 1] = single && N == 0 ? NAN : __ddiv_rn(fps[j], (double)N);
        kt[rank + 1] = single && P == 0 ? NAN : __ddiv_rn(tps[j], (double)P);
      }
      if (j == best) s_best_rank = rank + 1;
    }
    __syncthreads();
    if (tid == 0) s_base += total;
    __syncthreads();
  }
  n_kept = s_base;
  if (exact && tid == 0) {
    kf[0] = (N == 0) ? NAN : 0.0;
    kt[0] = (P == 0) ? NAN : 0.0;
  }
  __syncthreads();

  double auc = NAN;
  if (exact) {
    // np.trapezoid: d * (y[1:] + y[:-1]) / 2.0, then numpy's pairwise sum
    for (long long k = tid; k < n_kept; k += kFinalThreads)
      terms[k] = __ddiv_rn(__dmul_rn(__dadd_rn(kf[k + 1], -kf[k]), __dadd_rn(kt[k + 1], kt[k])), 2.0);
    __syncthreads();
    if (tid == 0) auc = np_pairwise_sum(terms, n_kept);
  } else if (!single) {
    // large curves (tile level): order-free parallel trapezoid over the unpruned points
    double acc = 0.0;
    for (long long j = tid; j < m; j += kFinalThreads) {
      double f0 = j ? fps[j - 1] : 0.0, t0 = j ? tps[j - 1] : 0.0;
      acc += (fps[j] - f0) / (double)N * ((tps[j] + t0) / (double)P) * 0.5;
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((tid & 31) == 0) s_red[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < kFinalThreads / 32; ++w) t += s_red[w];
      auc = t;
    }
  }
  if (tid == 0) {
    hdr->status = single ? 1 : 0;
    hdr->auc_exact = exact ? 1 : 0;
    hdr->auc = auc;
    hdr->n_points = n_kept + 1;
    if (single) {
      hdr->threshold = NAN; hdr->youden_j = NAN; hdr->best_index = -1;
    } else {
      hdr->threshold = (best < 0) ? INFINITY : thr[best];
      hdr->youden_j = best_j;
      hdr->best_index = (best < 0) ? 0 : s_best_rank;
    }
  }
}

template <typename T>
int roc_run(bq_ctx* ctx, const T* d_score, const uint8_t* d_label, const uint8_t* d_include, int64_t n,
            bq_roc_result* out) {
  if (n < 0 || n >= (1ll << 31)) return bq_fail(ctx, BQ_ERR_ARG, "bq_roc: n out of range");
  memset(out, 0, sizeof(*out));
  if (n == 0) {
    out->status = 2; out->threshold = NAN; out->youden_j = NAN; out->auc = NAN; out->best_index = -1;
    out->auc_exact = 1;
    return BQ_OK;
  }
  // scratch layout
  size_t sort_tmp = 0, scan_tmp = 0;
  cub::DeviceRadixSort::SortPairsDescending(nullptr, sort_tmp, (const T*)nullptr, (T*)nullptr,
                                            (const uint8_t*)nullptr, (uint8_t*)nullptr, (int)n, 0,
                                            (int)sizeof(T) * 8, ctx->stream);
  cub::DeviceScan::InclusiveSum(nullptr, scan_tmp, (const unsigned long long*)nullptr,
                                (unsigned long long*)nullptr, (int)n, ctx->stream);
  size_t cub_tmp = bq_align_up(sort_tmp > scan_tmp ? sort_tmp : scan_tmp, 256);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += bq_align_up(bytes, 256); return o; };
  size_t o_hdr = take(sizeof(RocHeader));
  size_t o_k0 = take((n + 1) * sizeof(T)), o_k1 = take((n + 1) * sizeof(T));
  size_t o_v0 = take(n), o_v1 = take(n);
  size_t o_packed = take(n * 8), o_scan = take(n * 8);
  size_t o_fps = take(n * 8), o_tps = take(n * 8), o_thr = take(n * 8), o_j = take(n * 8);
  size_t o_keep = take(n);
  size_t exact_n = (size_t)((n + 2 < kExactAucCap + 2) ? n + 2 : kExactAucCap + 2);
  size_t o_kf = take(exact_n * 8), o_kt = take(exact_n * 8), o_terms = take(exact_n * 8);
  size_t o_cub = take(cub_tmp);
  void* base = nullptr;
  int rc = bq_scratch(ctx, off, &base);
  if (rc) return rc;
  char* b = (char*)base;
  RocHeader* hdr = (RocHeader*)(b + o_hdr);
  T* k0 = (T*)(b + o_k0);
  T* k1 = (T*)(b + o_k1);
  uint8_t* v0 = (uint8_t*)(b + o_v0);
  uint8_t* v1 = (uint8_t*)(b + o_v1);
  unsigned long long* packed = (unsigned long long*)(b + o_packed);
  unsigned long long* scan = (unsigned long long*)(b + o_scan);
  double* fps = (double*)(b + o_fps);
  double* tps = (double*)(b + o_tps);
  double* thr = (double*)(b + o_thr);
  double* jv = (double*)(b + o_j);
  uint8_t* keep = (uint8_t*)(b + o_keep);

  BQ_CUDA(ctx, cudaMemsetAsync(hdr, 0, sizeof(RocHeader), ctx->stream));
  const int g = grid_for(n, kThreads, ctx->num_sms);
  roc_prep_kernel<T><<<g, kThreads, 0, ctx->stream>>>(d_score, d_label, d_include, n, k0, v0, hdr);
  BQ_LAUNCH_CHECK(ctx);
  BQ_CUDA(ctx, cub::DeviceRadixSort::SortPairsDescending(b + o_cub, sort_tmp, (const T*)k0, k1,
                                                         (const uint8_t*)v0, v1, (int)n, 0,
                                                         (int)sizeof(T) * 8, ctx->stream));
  ctx->launches += 4;   // onesweep: histogram, scan, 2-4 digit passes (counted conservatively)
  roc_pack_kernel<T><<<g, kThreads, 0, ctx->stream>>>(k1, v1, n, hdr, packed);
  BQ_LAUNCH_CHECK(ctx);
  BQ_CUDA(ctx, cub::DeviceScan::InclusiveSum(b + o_cub, scan_tmp, (const unsigned long long*)packed, scan,
                                             (int)n, ctx->stream));
  ctx->launches += 2;
  roc_scatter_kernel<T><<<g, kThreads, 0, ctx->stream>>>(k1, scan, n, hdr, fps, tps, thr);
  BQ_LAUNCH_CHECK(ctx);
  roc_points_kernel<<<g, kThreads, 0, ctx->stream>>>(hdr, fps, tps, jv, keep);
  BQ_LAUNCH_CHECK(ctx);
  roc_final_kernel<<<1, kFinalThreads, 0, ctx->stream>>>(hdr, fps, tps, thr, jv, keep,
                                                        (double*)(b + o_kf), (double*)(b + o_kt),
                                                        (double*)(b + o_terms));
  BQ_LAUNCH_CHECK(ctx);
  RocHeader h;
  BQ_CUDA(ctx, cudaMemcpyAsync(&h, hdr, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  BQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  out->threshold = h.threshold;
  out->youden_j = h.youden_j;
  out->auc = h.auc;
  out->n_pos = h.n_pos;
  out->n_neg = (int64_t)h.n_eff - h.n_pos;
  out->n_points = h.n_points;
  out->best_index = h.best_index;
  out->status = h.status;
  out->auc_exact = h.auc_exact;
  return BQ_OK;
}

// ---------------------------------------------------------------------------------------------
// K10: group-level apply
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void group_apply_kernel(int64_t L, const T* __restrict__ gp, const T* __restrict__ gu,
                                   const uint8_t* __restrict__ gt, double pred_thresh, double strict_thresh,
                                   int keep_mode, double slide_uq, T* __restrict__ err,
                                   uint8_t* __restrict__ correct, uint8_t* __restrict__ incorrect,
                                   uint8_t* __restrict__ ybin, uint8_t* __restrict__ include,
                                   unsigned long long* __restrict__ conf) {
  unsigned long long tp = 0, fp = 0, tn = 0, fn = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < L;
       i += (int64_t)gridDim.x * blockDim.x) {
    T p = gp[i];
    double pd = (double)p, ud = (double)gu[i];
    int y = gt[i];
    bool lo = pd < pred_thresh, hi = pd >= pred_thresh;
    // abs(uint8 - float) in the float dtype (threshold.py:237)
    err[i] = (T)fabs((double)y - pd);
    correct[i] = (lo && y == 0) || (hi && y == 1);
    incorrect[i] = (lo && y == 1) || (hi && y == 0);
    ybin[i] = hi;
    bool inc = keep_mode == BQ_KEEP_ALL ? true
               : keep_mode == BQ_KEEP_HIGH_CONFIDENCE ? (ud < slide_uq) : (ud >= slide_uq);
    include[i] = inc;
    if (inc) {
      bool t = y != 0, pp = pd > strict_thresh;
      tp += (t && pp); fp += (!t && pp); tn += (!t && !pp); fn += (t && !pp);
    }
  }
  for (int o = 16; o; o >>= 1) {
    tp += __shfl_xor_sync(0xffffffffu, tp, o);
    fp += __shfl_xor_sync(0xffffffffu, fp, o);
    tn += __shfl_xor_sync(0xffffffffu, tn, o);
    fn += __shfl_xor_sync(0xffffffffu, fn, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (tp) atomicAdd(conf + 0, tp);
    if (fp) atomicAdd(conf + 1, fp);
    if (tn) atomicAdd(conf + 2, tn);
    if (fn) atomicAdd(conf + 3, fn);
  }
}

size_t elem(int dtype) { return dtype == BQ_F64 ? 8 : 4; }

}  // namespace

struct bq_table : bq_table_impl {};

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int bq_table_create(bq_ctx* ctx, int64_t n, int dtype, const void* y_pred, const void* uncertainty,
                    const uint8_t* y_true, bq_table** out) {
  if (!ctx) return BQ_ERR_ARG;
  if (!out || n < 0 || (dtype != BQ_F32 && dtype != BQ_F64))
    return bq_fail(ctx, BQ_ERR_ARG, "bq_table_create: bad argument");
  if (n > 0 && (!y_pred || !uncertainty || !y_true))
    return bq_fail(ctx, BQ_ERR_ARG, "bq_table_create: null column");
  if (n >= (1ll << 31)) return bq_fail(ctx, BQ_ERR_ARG, "bq_table_create: n must be < 2^31");
  BQ_CUDA(ctx, cudaSetDevice(ctx->device));
  bq_table* t = new bq_table();
  t->ctx = ctx; t->n = n; t->dtype = dtype;
  int rc;
  if ((rc = bq_to_device_pooled(ctx, t->y_pred, y_pred, n * elem(dtype))) ||
      (rc = bq_to_device_pooled(ctx, t->unc, uncertainty, n * elem(dtype))) ||
      (rc = bq_to_device_pooled(ctx, t->y_true, y_true, n)) ||
      (rc = bq_alloc_pooled(ctx, t->incorrect, n > 0 ? n : 1))) {
    delete t;
    return rc;
  }
  // host staging buffers may be freed by the caller right after return
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) { delete t; return bq_fail(ctx, BQ_ERR_CUDA, "sync: %s", cudaGetErrorString(e)); }
  *out = t;
  return BQ_OK;
}

void bq_table_destroy(bq_table* t) {
  if (!t) return;
  cudaStreamSynchronize(t->ctx->stream);
  delete t;
}

int bq_table_set_groups(bq_table* t, const int32_t* codes, int32_t n_groups) {
  if (!t) return BQ_ERR_ARG;
  bq_ctx* ctx = t->ctx;
  if (n_groups < 0 || (t->n > 0 && !codes)) return bq_fail(ctx, BQ_ERR_ARG, "bq_table_set_groups: bad argument");
  BQ_CUDA(ctx, cudaSetDevice(ctx->device));
  t->groups_ready = false;
  t->n_groups = n_groups;
  int rc;
  if ((rc = bq_to_device_pooled(ctx, t->codes, codes, t->n * 4))) return rc;
  const size_t L = n_groups > 0 ? n_groups : 1;
  if ((rc = bq_alloc_pooled(ctx, t->seg_begin, L * 8)) || (rc = bq_alloc_pooled(ctx, t->seg_end, L * 8))) return rc;
  void* scr = nullptr;
  if ((rc = bq_scratch(ctx, L * 4 + 256, &scr))) return rc;
  int32_t* runs = (int32_t*)scr;
  int32_t* flag = (int32_t*)((char*)scr + bq_align_up(L * 4, 128));
  BQ_CUDA(ctx, cudaMemsetAsync(scr, 0, bq_align_up(L * 4, 128) + 4, ctx->stream));
  BQ_CUDA(ctx, cudaMemsetAsync(t->seg_begin.p, 0, L * 8, ctx->stream));
  BQ_CUDA(ctx, cudaMemsetAsync(t->seg_end.p, 0, L * 8, ctx->stream));
  t->contiguous = true;
  if (t->n > 0 && n_groups > 0) {
    const int g = grid_for(t->n, kThreads, ctx->num_sms);
    seg_bounds_kernel<<<g, kThreads, 0, ctx->stream>>>((const int32_t*)t->codes.p, t->n,
                                                      (int64_t*)t->seg_begin.p, (int64_t*)t->seg_end.p,
                                                      runs, flag);
    BQ_LAUNCH_CHECK(ctx);
    int32_t h_flag = 0;
    BQ_CUDA(ctx, cudaMemcpyAsync(&h_flag, flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
    BQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (h_flag) {
      // groups interleave in row order: stable radix sort of (code -> row) restores per-group row order
      t->contiguous = false;
      const int64_t n = t->n;
      if ((rc = bq_alloc_pooled(ctx, t->perm, n * 4)) || (rc = bq_alloc_pooled(ctx, t->sorted_codes, n * 4))) return rc;
      size_t tmp_bytes = 0;
      cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                      (const int32_t*)nullptr, (int32_t*)nullptr, (int)n, 0, 32, ctx->stream);
      size_t o_iota = 0, o_tmp = bq_align_up(n * 4, 256), o_runs = o_tmp + bq_align_up(tmp_bytes, 256);
      size_t total = o_runs + bq_align_up(L * 4, 128) + 256;
      if ((rc = bq_scratch(ctx, total, &scr))) return rc;
      int32_t* iota = (int32_t*)((char*)scr + o_iota);
      iota_kernel<<<g, kThreads, 0, ctx->stream>>>(iota, n);
      BQ_LAUNCH_CHECK(ctx);
      BQ_CUDA(ctx, cub::DeviceRadixSort::SortPairs((char*)scr + o_tmp, tmp_bytes, (const uint32_t*)t->codes.p,
                                                   (uint32_t*)t->sorted_codes.p, (const int32_t*)iota,
                                                   (int32_t*)t->perm.p, (int)n, 0, 32, ctx->stream));
      ctx->launches += 4;
      runs = (int32_t*)((char*)scr + o_runs);
      flag = (int32_t*)((char*)scr + o_runs + bq_align_up(L * 4, 128));
      BQ_CUDA(ctx, cudaMemsetAsync(runs, 0, bq_align_up(L * 4, 128) + 4, ctx->stream));
      seg_bounds_kernel<<<g, kThreads, 0, ctx->stream>>>((const int32_t*)t->sorted_codes.p, n,
                                                        (int64_t*)t->seg_begin.p, (int64_t*)t->seg_end.p,
                                                        runs, flag);
      BQ_LAUNCH_CHECK(ctx);
    }
  }
  BQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  t->groups_ready = true;
  return BQ_OK;
}

int bq_table_validate(bq_table* t, int64_t flags[4]) {
  if (!t || !flags) return BQ_ERR_ARG;
  bq_ctx* ctx = t->ctx;
  BQ_CUDA(ctx, cudaSetDevice(ctx->device));
  void* scr = nullptr;
  int rc = bq_scratch(ctx, 64, &scr);
  if (rc) return rc;
  BQ_CUDA(ctx, cudaMemsetAsync(scr, 0, 32, ctx->stream));
  if (t->n > 0) {
    const int g = grid_for(t->n, kThreads, ctx->num_sms);
    if (t->dtype == BQ_F32)
      validate_kernel<float><<<g, kThreads, 0, ctx->stream>>>((const float*)t->y_pred.p, (const float*)t->unc.p,
                                                             (const uint8_t*)t->y_true.p, t->n,
                                                             (unsigned long long*)scr);
    else
      validate_kernel<double><<<g, kThreads, 0, ctx->stream>>>((const double*)t->y_pred.p,
                                                              (const double*)t->unc.p,
                                                              (const uint8_t*)t->y_true.p, t->n,
                                                              (unsigned long long*)scr);
    BQ_LAUNCH_CHECK(ctx);
  }
  BQ_CUDA(ctx, cudaMemcpyAsync(flags, scr, 32, cudaMemcpyDeviceToHost, ctx->stream));
  BQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return BQ_OK;
}

int bq_tile_process(bq_table* t, double pred_thresh, double* error, uint8_t* correct, uint8_t* y_pred_bin) {
  if (!t) return BQ_ERR_ARG;
  bq_ctx* ctx = t->ctx;
  BQ_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t n = t->n;
  t->has_incorrect = true;
  if (n == 0) return BQ_OK;
  // outputs: device pointers are written in place, host pointers via scratch + one D2H copy each
  const bool e_dev = bq_is_device_ptr(error), c_dev = bq_is_device_ptr(correct), b_dev = bq_is_device_ptr(y_pred_bin);
  size_t off = 0;
  size_t o_e = off; if (error && !e_dev) off += bq_align_up(n * 8, 256);
  size_t o_c = off; if (correct && !c_dev) off += bq_align_up(n, 256);
  size_t o_b = off; if (y_pred_bin && !b_dev) off += bq_align_up(n, 256);
  void* scr = nullptr;
  int rc = bq_scratch(ctx, off + 256, &scr);
  if (rc) return rc;
  double* d_e = error ? (e_dev ? error : (double*)((char*)scr + o_e)) : nullptr;
  uint8_t* d_c = correct ? (c_dev ? correct : (uint8_t*)((char*)scr + o_c)) : nullptr;
  uint8_t* d_b = y_pred_bin ? (b_dev ? y_pred_bin : (uint8_t*)((char*)scr + o_b)) : nullptr;
  const int g = grid_for(n, kThreads, ctx->num_sms);
  if (t->dtype == BQ_F32)
    tile_process_kernel<float><<<g, kThreads, 0, ctx->stream>>>((const float*)t->y_pred.p,
                                                               (const uint8_t*)t->y_true.p, n, pred_thresh,
                                                               d_e, d_c, (uint8_t*)t->incorrect.p, d_b);
  else
    tile_process_kernel<double><<<g, kThreads, 0, ctx->stream>>>((const double*)t->y_pred.p,
                                                                (const uint8_t*)t->y_true.p, n, pred_thresh,
                                                                d_e, d_c, (uint8_t*)t->incorrect.p, d_b);
  BQ_LAUNCH_CHECK(ctx);
  if (error && !e_dev) BQ_CUDA(ctx, cudaMemcpyAsync(error, d_e, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (correct && !c_dev) BQ_CUDA(ctx, cudaMemcpyAsync(correct, d_c, n, cudaMemcpyDeviceToHost, ctx->stream));
  if (y_pred_bin && !b_dev) BQ_CUDA(ctx, cudaMemcpyAsync(y_pred_bin, d_b, n, cudaMemcpyDeviceToHost, ctx->stream));
  BQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return BQ_OK;
}

int bq_tile_roc(bq_table* t, int score_sel, int label_sel, bq_roc_result* out) {
  if (!t || !out) return BQ_ERR_ARG;
  bq_ctx* ctx = t->ctx;
  BQ_CUDA(ctx, cudaSetDevice(ctx->device));
  if (label_sel == BQ_LABEL_INCORRECT && !t->has_incorrect)
    return bq_fail(ctx, BQ_ERR_STATE, "bq_tile_roc: call bq_tile_process before using BQ_LABEL_INCORRECT");
  const void* score = score_sel == BQ_SCORE_Y_PRED ? t->y_pred.p : t->unc.p;
  const uint8_t* label = (const uint8_t*)(label_sel == BQ_LABEL_Y_TRUE ? t->y_true.p : t->incorrect.p);
  if (t->dtype == BQ_F32) return roc_run<float>(ctx, (const float*)score, label, nullptr, t->n, out);
  return roc_run<double>(ctx, (const double*)score, label, nullptr, t->n, out);
}

int bq_roc(bq_ctx* ctx, const void* score, int dtype, const uint8_t* label, const uint8_t* include,
           int64_t n, bq_roc_result* out) {
  if (!ctx) return BQ_ERR_ARG;
  if (!out || n < 0 || (dtype != BQ_F32 && dtype != BQ_F64) || (n > 0 && (!score || !label)))
    return bq_fail(ctx, BQ_ERR_ARG, "bq_roc: bad argument");
  BQ_CUDA(ctx, cudaSetDevice(ctx->device));
  DevBuf s, l, inc;
  int rc;
  if ((rc = bq_to_device_pooled(ctx, s, score, n * elem(dtype))) || (rc = bq_to_device_pooled(ctx, l, label, n))) return rc;
  if (include && (rc = bq_to_device_pooled(ctx, inc, include, n))) return rc;
  if (dtype == BQ_F32)
    rc = roc_run<float>(ctx, (const float*)s.p, (const uint8_t*)l.p, (const uint8_t*)inc.p, n, out);
  else
    rc = roc_run<double>(ctx, (const double*)s.p, (const uint8_t*)l.p, (const uint8_t*)inc.p, n, out);
  cudaStreamSynchronize(ctx->stream);
  return rc;
}

int bq_table_set_tile_filter(bq_table* t, int enabled, double tile_uq) {
  if (!t) return BQ_ERR_ARG;
  t->filter_on = enabled != 0;
  t->tile_uq = tile_uq;
  return BQ_OK;
}

int bq_group_reduce(bq_table* t, void* g_pred, void* g_unc, double* g_true_mean, int64_t* g_count,
                    int64_t* g_first_row) {
  if (!t) return BQ_ERR_ARG;
  bq_ctx* ctx = t->ctx;
  if (!t->groups_ready) return bq_fail(ctx, BQ_ERR_STATE, "bq_group_reduce: call bq_table_set_groups first");
  if (!g_pred || !g_unc || !g_true_mean || !g_count || !g_first_row)
    return bq_fail(ctx, BQ_ERR_ARG, "bq_group_reduce: null output");
  BQ_CUDA(ctx, cudaSetDevice(ctx->device));
  const int32_t L = t->n_groups;
  if (L == 0) return BQ_OK;
  const size_t es = elem(t->dtype);
  size_t o_p = 0, o_u = bq_align_up(L * es, 256), o_t = o_u + bq_align_up(L * es, 256),
         o_c = o_t + bq_align_up(L * 8, 256), o_f = o_c + bq_align_up(L * 8, 256),
         total = o_f + bq_align_up(L * 8, 256);
  void* scr = nullptr;
  int rc = bq_scratch(ctx, total, &scr);
  if (rc) return rc;
  char* b = (char*)scr;
  const int warps = kThreads / 32;
  int grid = (L + warps - 1) / warps;
  const int32_t* perm = t->contiguous ? nullptr : (const int32_t*)t->perm.p;
  if (t->dtype == BQ_F32)
    group_kahan_kernel<float><<<grid, kThreads, 0, ctx->stream>>>(
        (const float*)t->y_pred.p, (const float*)t->unc.p, (const uint8_t*)t->y_true.p, perm,
        (const int64_t*)t->seg_begin.p, (const int64_t*)t->seg_end.p, t->filter_on, t->tile_uq, L,
        (float*)(b + o_p), (float*)(b + o_u), (double*)(b + o_t), (int64_t*)(b + o_c), (int64_t*)(b + o_f));
  else
    group_kahan_kernel<double><<<grid, kThreads, 0, ctx->stream>>>(
        (const double*)t->y_pred.p, (const double*)t->unc.p, (const uint8_t*)t->y_true.p, perm,
        (const int64_t*)t->seg_begin.p, (const int64_t*)t->seg_end.p, t->filter_on, t->tile_uq, L,
        (double*)(b + o_p), (double*)(b + o_u), (double*)(b + o_t), (int64_t*)(b + o_c), (int64_t*)(b + o_f));
  BQ_LAUNCH_CHECK(ctx);
  if ((rc = bq_from_device(ctx, g_pred, b + o_p, L * es)) || (rc = bq_from_device(ctx, g_unc, b + o_u, L * es)) ||
      (rc = bq_from_device(ctx, g_true_mean, b + o_t, L * 8)) || (rc = bq_from_device(ctx, g_count, b + o_c, L * 8)) ||
      (rc = bq_from_device(ctx, g_first_row, b + o_f, L * 8)))
    return rc;
  BQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return BQ_OK;
}

int bq_group_apply(bq_ctx* ctx, int64_t L, int dtype, const void* g_pred, const void* g_unc,
                   const uint8_t* g_true, double pred_thresh, double slide_pred_strict, int keep_mode,
                   double slide_uq, void* error, uint8_t* correct, uint8_t* incorrect, uint8_t* y_pred_bin,
                   uint8_t* include, int64_t confusion[4]) {
  if (!ctx) return BQ_ERR_ARG;
  if (L < 0 || (dtype != BQ_F32 && dtype != BQ_F64) || keep_mode < 0 || keep_mode > 2 || !confusion ||
      (L > 0 && (!g_pred || !g_unc || !g_true || !error || !correct || !incorrect || !y_pred_bin || !include)))
    return bq_fail(ctx, BQ_ERR_ARG, "bq_group_apply: bad argument");
  BQ_CUDA(ctx, cudaSetDevice(ctx->device));
  memset(confusion, 0, 32);
  if (L == 0) return BQ_OK;
  const size_t es = elem(dtype);
  DevBuf p, u, y;
  int rc;
  if ((rc = bq_to_device_pooled(ctx, p, g_pred, L * es)) || (rc = bq_to_device_pooled(ctx, u, g_unc, L * es)) ||
      (rc = bq_to_device_pooled(ctx, y, g_true, L)))
    return rc;
  size_t o_e = 0, o_c = bq_align_up(L * es, 256), o_i = o_c + bq_align_up(L, 256), o_b = o_i + bq_align_up(L, 256),
         o_inc = o_b + bq_align_up(L, 256), o_conf = o_inc + bq_align_up(L, 256), total = o_conf + 256;
  void* scr = nullptr;
  if ((rc = bq_scratch(ctx, total, &scr))) return rc;
  char* b = (char*)scr;
  BQ_CUDA(ctx, cudaMemsetAsync(b + o_conf, 0, 32, ctx->stream));
  const int g = grid_for(L, kThreads, ctx->num_sms);
  if (dtype == BQ_F32)
    group_apply_kernel<float><<<g, kThreads, 0, ctx->stream>>>(
        L, (const float*)p.p, (const float*)u.p, (const uint8_t*)y.p, pred_thresh, slide_pred_strict, keep_mode,
        slide_uq, (float*)(b + o_e), (uint8_t*)(b + o_c), (uint8_t*)(b + o_i), (uint8_t*)(b + o_b),
        (uint8_t*)(b + o_inc), (unsigned long long*)(b + o_conf));
  else
    group_apply_kernel<double><<<g, kThreads, 0, ctx->stream>>>(
        L, (const double*)p.p, (const double*)u.p, (const uint8_t*)y.p, pred_thresh, slide_pred_strict, keep_mode,
        slide_uq, (double*)(b + o_e), (uint8_t*)(b + o_c), (uint8_t*)(b + o_i), (uint8_t*)(b + o_b),
        (uint8_t*)(b + o_inc), (unsigned long long*)(b + o_conf));
  BQ_LAUNCH_CHECK(ctx);
  if ((rc = bq_from_device(ctx, error, b + o_e, L * es)) || (rc = bq_from_device(ctx, correct, b + o_c, L)) ||
      (rc = bq_from_device(ctx, incorrect, b + o_i, L)) || (rc = bq_from_device(ctx, y_pred_bin, b + o_b, L)) ||
      (rc = bq_from_device(ctx, include, b + o_inc, L)))
    return rc;
  BQ_CUDA(ctx, cudaMemcpyAsync(confusion, b + o_conf, 32, cudaMemcpyDeviceToHost, ctx->stream));
  BQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return BQ_OK;
}

}  // extern "C"
