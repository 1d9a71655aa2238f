// Back-end reductions of the scoring recipe: speaker / global means of embedding tables and the
// detection-error metrics (EER, minDCF) of a scored trial list.  All of them are HBM-bound integer / streaming work.
//
//   * svk_sort_pairs_f64   stable LSD radix sort of (float64 score, int32 payload) pairs, ascending: the reference sorts
//                          Python floats with a stable sort (compute_eer.py:38-40), so ties keep their file order and the
//                          keys stay float64 (scores are parsed from text).  8 passes of 8 bits over an order-preserving
//                          uint64 image of the key: per-block digit histograms -> one exclusive scan -> stable scatter.
//   * svk_det_metrics      cumulative target / non-target counts along the sorted list (compute_eer.py:48-68), the first
//                          index minimising |fnr - fpr| (:99-100) and the first index minimising the detection cost
//                          (local/compute_min_dcf.py:93-102), in float64 with the reference's operation order.
//   * svk_segment_mean     per-speaker mean of embedding rows, accumulated in file order (compute_speaker_mean.py:16-27).
//   * svk_col_mean         mean over all rows (compute_mean.py:9-20).
#include "svk_common.cuh"

namespace {

constexpr int SORT_THREADS = 256;
constexpr int SORT_ROUNDS = 8;
constexpr int SORT_CHUNK = SORT_THREADS * SORT_ROUNDS;     // keys per block

__device__ __forceinline__ unsigned long long key_image(double v) {
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);     // ascending order of the doubles = ascending order of the images
}

__global__ void __launch_bounds__(SORT_THREADS) sort_hist_kernel(const double* __restrict__ keys, long long n, int shift,
                                                                 int nblk, unsigned int* __restrict__ hist) {
  pdl_prologue();
  __shared__ unsigned int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const long long base = (long long)blockIdx.x * SORT_CHUNK;
  for (int r = 0; r < SORT_ROUNDS; ++r) {
    const long long i = base + r * SORT_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[(key_image(keys[i]) >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];     // digit-major: one linear exclusive scan gives the offsets
}

// exclusive scan of `len` counters in place (one block; len = 256 * nblk is small: 131 k entries at 1 M keys)
__global__ void __launch_bounds__(1024) scan_kernel(unsigned int* __restrict__ a, long long len) {
  pdl_prologue();
  __shared__ unsigned long long part[1024];
  const long long per = (len + 1023) / 1024;
  const long long lo = (long long)threadIdx.x * per, hi = (lo + per < len) ? lo + per : len;
  unsigned long long s = 0;
  for (long long i = lo; i < hi; ++i) s += a[i];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    for (int t = 0; t < 1024; ++t) { const unsigned long long v = part[t]; part[t] = run; run += v; }
  }
  __syncthreads();
  unsigned int run = (unsigned int)part[threadIdx.x];
  for (long long i = lo; i < hi; ++i) { const unsigned int v = a[i]; a[i] = run; run += v; }
}

__global__ void __launch_bounds__(SORT_THREADS) sort_scatter_kernel(const double* __restrict__ keys, const int* __restrict__ vals,
                                                                    double* __restrict__ keys_out, int* __restrict__ vals_out,
                                                                    long long n, int shift, int nblk,
                                                                    const unsigned int* __restrict__ offs) {
  pdl_prologue();
  __shared__ unsigned int start[256];               // where this block's keys of each digit begin (+ those already placed)
  __shared__ unsigned int whist[SORT_THREADS / 32][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  start[threadIdx.x] = offs[(size_t)threadIdx.x * nblk + blockIdx.x];
  const long long base = (long long)blockIdx.x * SORT_CHUNK;
  for (int r = 0; r < SORT_ROUNDS; ++r) {
    for (int w = 0; w < SORT_THREADS / 32; ++w) whist[w][threadIdx.x] = 0;
    __syncthreads();
    const long long i = base + r * SORT_THREADS + threadIdx.x;
    const bool valid = i < n;
    double k = 0.0; int v = 0; unsigned int d = 0, rank = 0;
    const unsigned int vmask = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      k = keys[i]; v = vals[i];
      d = (unsigned int)((key_image(k) >> shift) & 255u);
      const unsigned int peers = __match_any_sync(vmask, d);
      rank = __popc(peers & ((1u << lane) - 1u));                      // earlier lanes with the same digit: stability
      if (rank == 0) whist[warp][d] = __popc(peers);
    }
    __syncthreads();
    if (valid) {
      unsigned int pos = start[d] + rank;
      for (int w = 0; w < warp; ++w) pos += whist[w][d];
      keys_out[pos] = k; vals_out[pos] = v;
    }
    __syncthreads();
    unsigned int add = 0;
    for (int w = 0; w < SORT_THREADS / 32; ++w) add += whist[w][threadIdx.x];
    start[threadIdx.x] += add;
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------- detection metrics
constexpr int DET_THREADS = 256;
constexpr int DET_PER = 8;
constexpr int DET_CHUNK = DET_THREADS * DET_PER;

__global__ void __launch_bounds__(DET_THREADS) det_blocksum_kernel(const int* __restrict__ lab, long long n,
                                                                   unsigned int* __restrict__ bsum) {
  pdl_prologue();
  __shared__ unsigned int red[DET_THREADS];
  const long long base = (long long)blockIdx.x * DET_CHUNK + (long long)threadIdx.x * DET_PER;
  unsigned int s = 0;
  for (int e = 0; e < DET_PER; ++e) if (base + e < n) s += (lab[base + e] != 0);
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = DET_THREADS / 2; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) bsum[blockIdx.x] = red[0];
}

struct DetBest { double v_eer; long long i_eer; double v_dcf; long long i_dcf; };

__device__ __forceinline__ void det_merge(double& v, long long& i, double v2, long long i2) {
  // first index of the minimum; NaN never wins (numpy.nanargmin, compute_eer.py:99)
  if (v2 < v || (v2 == v && i2 < i) || (v != v && v2 == v2)) { v = v2; i = i2; }
}

// bsum holds the EXCLUSIVE scan of the block sums; total targets = n_t.
__global__ void __launch_bounds__(DET_THREADS) det_eval_kernel(const int* __restrict__ lab, long long n,
                                                               const unsigned int* __restrict__ bsum, unsigned int n_t,
                                                               double p_target, double c_miss, double c_fa,
                                                               DetBest* __restrict__ out) {
  pdl_prologue();
  __shared__ unsigned int tsum[DET_THREADS];
  __shared__ DetBest red[DET_THREADS];
  const long long base = (long long)blockIdx.x * DET_CHUNK + (long long)threadIdx.x * DET_PER;
  int l[DET_PER]; unsigned int s = 0;
  for (int e = 0; e < DET_PER; ++e) { l[e] = (base + e < n) ? (lab[base + e] != 0) : 0; s += l[e]; }
  tsum[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) { unsigned int run = 0; for (int t = 0; t < DET_THREADS; ++t) { const unsigned int v = tsum[t]; tsum[t] = run; run += v; } }
  __syncthreads();
  unsigned long long cum_t = (unsigned long long)bsum[blockIdx.x] + tsum[threadIdx.x];
  const double nt = (double)n_t, nn = (double)(n - (long long)n_t);
  const double one_m_p = __dsub_rn(1.0, p_target);
  DetBest b; b.v_eer = __longlong_as_double(0x7ff8000000000000ll); b.i_eer = n; b.v_dcf = __longlong_as_double(0x7ff0000000000000ll); b.i_dcf = n;
  for (int e = 0; e < DET_PER; ++e) {
    const long long i = base + e;
    if (i >= n) break;
    cum_t += l[e];
    const unsigned long long cum_n = (unsigned long long)(i + 1) - cum_t;
    const double fnr = __ddiv_rn((double)cum_t, nt);                                   // compute_eer.py:60
    const double fpr = __dsub_rn(1.0, __ddiv_rn((double)cum_n, nn));                   // :65
    det_merge(b.v_eer, b.i_eer, fabs(__dsub_rn(fnr, fpr)), i);
    // c_miss * fnr * p_target + c_fa * fpr * (1 - p_target), left to right, no fused multiply-add (compute_min_dcf.py:98)
    const double c_det = __dadd_rn(__dmul_rn(__dmul_rn(c_miss, fnr), p_target), __dmul_rn(__dmul_rn(c_fa, fpr), one_m_p));
    if (c_det < b.v_dcf) { b.v_dcf = c_det; b.i_dcf = i; }                              // strict <: first minimum (:99)
  }
  red[threadIdx.x] = b;
  __syncthreads();
  for (int o = DET_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      DetBest& a = red[threadIdx.x]; const DetBest& c = red[threadIdx.x + o];
      det_merge(a.v_eer, a.i_eer, c.v_eer, c.i_eer);
      if (c.v_dcf < a.v_dcf || (c.v_dcf == a.v_dcf && c.i_dcf < a.i_dcf)) { a.v_dcf = c.v_dcf; a.i_dcf = c.i_dcf; }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = red[0];
}

// one block: merge the per-block candidates, then evaluate the rates at the two winning indices
__global__ void __launch_bounds__(DET_THREADS) det_final_kernel(const DetBest* __restrict__ cand, int nblk, const int* __restrict__ lab,
                                                                long long n, const unsigned int* __restrict__ bsum, unsigned int n_t,
                                                                double* __restrict__ out) {
  pdl_prologue();
  __shared__ DetBest red[DET_THREADS];
  DetBest b; b.v_eer = __longlong_as_double(0x7ff8000000000000ll); b.i_eer = n; b.v_dcf = __longlong_as_double(0x7ff0000000000000ll); b.i_dcf = n;
  for (int k = threadIdx.x; k < nblk; k += DET_THREADS) {
    det_merge(b.v_eer, b.i_eer, cand[k].v_eer, cand[k].i_eer);
    if (cand[k].v_dcf < b.v_dcf || (cand[k].v_dcf == b.v_dcf && cand[k].i_dcf < b.i_dcf)) { b.v_dcf = cand[k].v_dcf; b.i_dcf = cand[k].i_dcf; }
  }
  red[threadIdx.x] = b;
  __syncthreads();
  for (int o = DET_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      DetBest& a = red[threadIdx.x]; const DetBest& c = red[threadIdx.x + o];
      det_merge(a.v_eer, a.i_eer, c.v_eer, c.i_eer);
      if (c.v_dcf < a.v_dcf || (c.v_dcf == a.v_dcf && c.i_dcf < a.i_dcf)) { a.v_dcf = c.v_dcf; a.i_dcf = c.i_dcf; }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const DetBest w = red[0];
    // rates at idxE: cumulative count = exclusive block sum + labels of the block up to the index
    const long long i = w.i_eer < n ? w.i_eer : n - 1;
    const long long blk = i / DET_CHUNK;
    unsigned long long cum_t = bsum[blk];
    for (long long j = blk * DET_CHUNK; j <= i; ++j) cum_t += (lab[j] != 0);
    const unsigned long long cum_n = (unsigned long long)(i + 1) - cum_t;
    const double fnr = __ddiv_rn((double)cum_t, (double)n_t);
    const double fpr = __dsub_rn(1.0, __ddiv_rn((double)cum_n, (double)(n - (long long)n_t)));
    out[0] = fpr > fnr ? fpr : fnr;          // eer = max(fprs[idxE], fnrs[idxE])  (compute_eer.py:100)
    out[1] = (double)w.i_eer;
    out[2] = w.v_dcf;                        // min_c_det (not yet normalised by c_def)
    out[3] = (double)w.i_dcf;
    out[4] = (double)n_t;
    out[5] = (double)(n - (long long)n_t);
  }
}

// ---------------------------------------------------------------------------------------------- means
__global__ void __launch_bounds__(256) segment_mean_kernel(const float* __restrict__ X, const int* __restrict__ order,
                                                           const int* __restrict__ offsets, int D, float* __restrict__ out) {
  pdl_prologue();
  const int seg = blockIdx.x;
  const int lo = offsets[seg], hi = offsets[seg + 1];
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f;
    for (int k = lo; k < hi; ++k) acc += X[(size_t)order[k] * D + d];       // file order: speaker_mean[spk] += vec (:24)
    out[(size_t)seg * D + d] = hi > lo ? acc / (float)(hi - lo) : 0.f;       // speaker_mean[spk] /= spk2num[spk] (:27)
  }
}

// partial[b][d] = float64 sum of a slab of rows; a second launch (rows = number of slabs, scale = 1/n) finishes
__global__ void __launch_bounds__(256) col_sum_kernel(const float* __restrict__ X, long long n, int D, long long rows_per,
                                                      double* __restrict__ partial) {
  pdl_prologue();
  const long long lo = (long long)blockIdx.x * rows_per, hi = (lo + rows_per < n) ? lo + rows_per : n;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    double acc = 0.0;
    for (long long r = lo; r < hi; ++r) acc += (double)X[(size_t)r * D + d];
    partial[(size_t)blockIdx.x * D + d] = acc;
  }
}
__global__ void __launch_bounds__(256) col_mean_final_kernel(const double* __restrict__ partial, int nb, int D, long long n,
                                                             float* __restrict__ out) {
  pdl_prologue();
  for (int d = blockIdx.x * blockDim.x + threadIdx.x; d < D; d += gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (int b = 0; b < nb; ++b) acc += partial[(size_t)b * D + d];
    out[d] = (float)(acc / (double)n);
  }
}

inline size_t align256(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace

SVK_API size_t svk_sort_pairs_f64_workspace_bytes(long long n) {
  if (n <= 0) return 256;
  const long long nblk = (n + SORT_CHUNK - 1) / SORT_CHUNK;
  return align256((size_t)n * 8) + align256((size_t)n * 4) + align256((size_t)256 * nblk * 4);
}

SVK_API int svk_sort_pairs_f64(const double* keys, const int* vals, double* keys_out, int* vals_out, long long n,
                               void* workspace, size_t workspace_bytes, void* stream) {
  SVK_REQUIRE(n >= 0 && n < (1ll << 31), SVK_E_BADARG, "sort_pairs_f64: n=%lld out of range", n);
  if (n == 0) return 0;
  SVK_REQUIRE(keys && vals && keys_out && vals_out && workspace, SVK_E_BADARG, "sort_pairs_f64: null pointer");
  SVK_REQUIRE(workspace_bytes >= svk_sort_pairs_f64_workspace_bytes(n), SVK_E_BADARG, "sort_pairs_f64: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int nblk = (int)((n + SORT_CHUNK - 1) / SORT_CHUNK);
  uint8_t* w = (uint8_t*)workspace;
  double* tk = (double*)w; w += align256((size_t)n * 8);
  int* tv = (int*)w; w += align256((size_t)n * 4);
  unsigned int* hist = (unsigned int*)w;
  const double* sk = keys; const int* sv = vals;
  for (int pass = 0; pass < 8; ++pass) {
    double* dk = (pass & 1) ? keys_out : tk;
    int* dv = (pass & 1) ? vals_out : tv;
    svk_launch(sort_hist_kernel, nblk, SORT_THREADS, 0, st, sk, n, pass * 8, nblk, hist);
    SVK_LAUNCH_CHECK("sort_pairs_f64(hist)");
    svk_launch(scan_kernel, 1, 1024, 0, st, hist, (long long)256 * nblk);
    SVK_LAUNCH_CHECK("sort_pairs_f64(scan)");
    svk_launch(sort_scatter_kernel, nblk, SORT_THREADS, 0, st, sk, sv, dk, dv, n, pass * 8, nblk, hist);
    SVK_LAUNCH_CHECK("sort_pairs_f64(scatter)");
    sk = dk; sv = dv;
  }
  return 0;
}

SVK_API size_t svk_det_metrics_workspace_bytes(long long n) {
  const long long nblk = n > 0 ? (n + DET_CHUNK - 1) / DET_CHUNK : 1;
  return align256((size_t)(nblk + 1) * 4) + align256((size_t)nblk * sizeof(DetBest));
}

SVK_API int svk_det_metrics(const int* sorted_labels, long long n, long long n_target, double p_target, double c_miss,
                            double c_fa, double* out6, void* workspace, size_t workspace_bytes, void* stream) {
  SVK_REQUIRE(n > 0 && n < (1ll << 31), SVK_E_BADARG, "det_metrics: n=%lld out of range", n);
  SVK_REQUIRE(n_target > 0 && n_target < n, SVK_E_BADARG, "det_metrics: need at least one target and one non-target trial");
  SVK_REQUIRE(sorted_labels && out6 && workspace, SVK_E_BADARG, "det_metrics: null pointer");
  SVK_REQUIRE(workspace_bytes >= svk_det_metrics_workspace_bytes(n), SVK_E_BADARG, "det_metrics: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int nblk = (int)((n + DET_CHUNK - 1) / DET_CHUNK);
  uint8_t* w = (uint8_t*)workspace;
  unsigned int* bsum = (unsigned int*)w; w += align256((size_t)(nblk + 1) * 4);
  DetBest* cand = (DetBest*)w;
  svk_launch(det_blocksum_kernel, nblk, DET_THREADS, 0, st, sorted_labels, n, bsum);
  SVK_LAUNCH_CHECK("det_metrics(blocksum)");
  svk_launch(scan_kernel, 1, 1024, 0, st, bsum, nblk);
  SVK_LAUNCH_CHECK("det_metrics(scan)");
  svk_launch(det_eval_kernel, nblk, DET_THREADS, 0, st, sorted_labels, n, bsum, (unsigned int)n_target, p_target, c_miss, c_fa, cand);
  SVK_LAUNCH_CHECK("det_metrics(eval)");
  svk_launch(det_final_kernel, 1, DET_THREADS, 0, st, cand, nblk, sorted_labels, n, bsum, (unsigned int)n_target, out6);
  SVK_LAUNCH_CHECK("det_metrics(final)");
  return 0;
}

SVK_API int svk_segment_mean(const float* X, const int* order, const int* offsets, int n_seg, int D, float* out, void* stream) {
  SVK_REQUIRE(n_seg >= 0 && D > 0, SVK_E_BADARG, "segment_mean: bad sizes");
  if (n_seg == 0) return 0;
  SVK_REQUIRE(X && order && offsets && out, SVK_E_BADARG, "segment_mean: null pointer");
  svk_launch(segment_mean_kernel, n_seg, 256, 0, as_stream(stream), X, order, offsets, D, out);
  SVK_LAUNCH_CHECK("segment_mean");
  return 0;
}

SVK_API size_t svk_col_mean_workspace_bytes(long long n, int D) {
  (void)n;
  return (size_t)svk_num_sms() * 4 * (size_t)D * sizeof(double);
}

SVK_API int svk_col_mean(const float* X, long long n, int D, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  SVK_REQUIRE(n > 0 && D > 0, SVK_E_BADARG, "col_mean: bad sizes");
  SVK_REQUIRE(X && out && workspace, SVK_E_BADARG, "col_mean: null pointer");
  SVK_REQUIRE(workspace_bytes >= svk_col_mean_workspace_bytes(n, D), SVK_E_BADARG, "col_mean: workspace too small");
  long long nb = (long long)svk_num_sms() * 4;
  if (nb > n) nb = n;
  const long long rows_per = (n + nb - 1) / nb;
  nb = (n + rows_per - 1) / rows_per;
  cudaStream_t st = as_stream(stream);
  svk_launch(col_sum_kernel, (int)nb, 256, 0, st, X, n, D, rows_per, (double*)workspace);
  SVK_LAUNCH_CHECK("col_mean(sum)");
  svk_launch(col_mean_final_kernel, (D + 255) / 256, 256, 0, st, (const double*)workspace, (int)nb, D, n, out);
  SVK_LAUNCH_CHECK("col_mean(final)");
  return 0;
}
