// AAM-softmax head pieces (row L2-normalise, additive-angular margin, cross-entropy with top-k rank) and the
// trial-scoring kernels (cosine pairs, per-row top-k mean/std by radix select, adaptive s-norm).
#include "svk_common.cuh"

// ------------------------------------------------------------------------------------------ row L2 normalise
__global__ void __launch_bounds__(256) l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ xhat,
                                                         float* __restrict__ inv, int rows, int cols, float eps) {
  pdl_prologue();
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* p = x + (long long)warp * cols;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) { float v = p[c]; s = fmaf(v, v, s); }
  s = warp_sum(s);
  float r = 1.f / fmaxf(sqrtf(s), eps);
  if (lane == 0 && inv) inv[warp] = r;
  for (int c = lane; c < cols; c += 32) xhat[(long long)warp * cols + c] = p[c] * r;
}
SVK_API int svk_l2norm_rows_fwd(const float* x, float* xhat, float* inv, int rows, int cols, float eps, void* stream) {
  SVK_REQUIRE(x && xhat && rows > 0 && cols > 0, SVK_E_BADARG, "l2norm_rows_fwd: bad args");
  svk_launch(l2norm_fwd_kernel, (rows + 7) / 8, 256, 0, as_stream(stream), x, xhat, inv, rows, cols, eps);
  SVK_LAUNCH_CHECK("l2norm_rows_fwd");
  return 0;
}
__global__ void __launch_bounds__(256) l2norm_bwd_kernel(const float* __restrict__ dxhat, const float* __restrict__ xhat,
                                                         const float* __restrict__ inv, float* __restrict__ dx, int rows,
                                                         int cols) {
  pdl_prologue();
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* g = dxhat + (long long)warp * cols;
  const float* h = xhat + (long long)warp * cols;
  float d = 0.f;
  for (int c = lane; c < cols; c += 32) d = fmaf(g[c], h[c], d);
  d = warp_sum(d);
  float r = inv[warp];
  for (int c = lane; c < cols; c += 32) dx[(long long)warp * cols + c] = (g[c] - h[c] * d) * r;
}
SVK_API int svk_l2norm_rows_bwd(const float* dxhat, const float* xhat, const float* inv, float* dx, int rows, int cols,
                                void* stream) {
  SVK_REQUIRE(dxhat && xhat && inv && dx && rows > 0 && cols > 0, SVK_E_BADARG, "l2norm_rows_bwd: bad args");
  svk_launch(l2norm_bwd_kernel, (rows + 7) / 8, 256, 0, as_stream(stream), dxhat, xhat, inv, dx, rows, cols);
  SVK_LAUNCH_CHECK("l2norm_rows_bwd");
  return 0;
}

// A label outside [0, C) is a caller error (--spk-num mismatch, bad utt2spkid line).  torch's CrossEntropyLoss / scatter_
// raise a device-side assert in that situation; so do we: message + trap, so the failure surfaces as a CUDA error on the
// next synchronising call instead of an out-of-bounds read or a silently dropped target term.
__device__ __forceinline__ void require_label(long long lbl, int C, int row, const char* what) {
  if (lbl < 0 || lbl >= (long long)C) {
    printf("svk %s: label %lld of row %d is outside [0, %d)\n", what, lbl, row, C);
    __trap();
  }
}

// ------------------------------------------------------------------------------------------ AAM margin
__global__ void __launch_bounds__(256) aam_margin_fwd_kernel(float* __restrict__ z, const long long* __restrict__ label,
                                                             float* __restrict__ cos_t, int B, int C, float cos_m,
                                                             float sin_m, float th, float mm, float s) {
  pdl_prologue();
  long long n = (long long)B * C;
  if (blockIdx.x == 0) {
    for (int b = threadIdx.x; b < B; b += blockDim.x) require_label(label[b], C, b, "aam_margin_fwd");
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int b = (int)(i / C), c = (int)(i % C);
    float v = z[i];
    if ((long long)c == label[b]) {
      cos_t[b] = v;
      float sine = sqrtf(fminf(fmaxf(1.f - v * v, 0.f), 1.f));
      float phi = v * cos_m - sine * sin_m;
      v = (v - th > 0.f) ? phi : v - mm;
    }
    z[i] = v * s;
  }
}
SVK_API int svk_aam_margin_fwd(float* z, const long long* label, float* cos_t, int B, int C, float cos_m, float sin_m,
                               float th, float mm, float s, void* stream) {
  SVK_REQUIRE(z && label && cos_t && B > 0 && C > 0, SVK_E_BADARG, "aam_margin_fwd: bad args");
  long long n = (long long)B * C; long long b = (n + 255) / 256; long long cap = (long long)svk_num_sms() * 8; if (b > cap) b = cap;
  svk_launch(aam_margin_fwd_kernel, (int)b, 256, 0, as_stream(stream), z, label, cos_t, B, C, cos_m, sin_m, th, mm, s);
  SVK_LAUNCH_CHECK("aam_margin_fwd");
  return 0;
}
__global__ void __launch_bounds__(256) aam_margin_bwd_kernel(float* __restrict__ g, const long long* __restrict__ label,
                                                             const float* __restrict__ cos_t, int B, int C, int ld,
                                                             float cos_m, float sin_m, float th, float s) {
  pdl_prologue();
  long long n = (long long)B * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int b = (int)(i / C), c = (int)(i % C);
    float* gp = g + (long long)b * ld + c;
    float v = *gp * s;
    if ((long long)c == label[b]) {
      float ct = cos_t[b];
      if (ct - th > 0.f) {
        float u = 1.f - ct * ct;
        float dsine = (u > 0.f && u < 1.f) ? -ct / sqrtf(u) : 0.f;   // d sqrt(clamp(1-c^2,0,1)) / dc
        v *= cos_m - sin_m * dsine;
      }
    }
    *gp = v;
  }
}
SVK_API int svk_aam_margin_bwd(float* g, const long long* label, const float* cos_t, int B, int C, int ld, float cos_m,
                               float sin_m, float th, float s, void* stream) {
  SVK_REQUIRE(g && label && cos_t && B > 0 && C > 0 && ld >= C, SVK_E_BADARG, "aam_margin_bwd: bad args");
  long long n = (long long)B * C; long long b = (n + 255) / 256; long long cap = (long long)svk_num_sms() * 8; if (b > cap) b = cap;
  svk_launch(aam_margin_bwd_kernel, (int)b, 256, 0, as_stream(stream), g, label, cos_t, B, C, ld, cos_m, sin_m, th, s);
  SVK_LAUNCH_CHECK("aam_margin_bwd");
  return 0;
}

// ------------------------------------------------------------------------------------------ cross entropy
__device__ inline float block_reduce(float v, bool is_max, float* sm) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) sm[w] = v;
  __syncthreads();
  int nw = blockDim.x >> 5;
  float r = (threadIdx.x < nw) ? sm[threadIdx.x] : (is_max ? -INFINITY : 0.f);
  if (w == 0) { r = is_max ? warp_max(r) : warp_sum(r); if (lane == 0) sm[0] = r; }
  __syncthreads();
  return sm[0];
}
__global__ void __launch_bounds__(256) ce_fwd_kernel(const float* __restrict__ z, const long long* __restrict__ label,
                                                     float* __restrict__ loss_rows, float* __restrict__ lse,
                                                     int* __restrict__ rank, float* __restrict__ loss_mean, int B,
                                                     int C) {
  pdl_prologue();
  __shared__ float sm[32];
  int b = blockIdx.x;
  const float* p = z + (long long)b * C;
  require_label(label[b], C, b, "ce_fwd");
  float zt = p[label[b]];
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < C; c += blockDim.x) mx = fmaxf(mx, p[c]);
  mx = block_reduce(mx, true, sm);
  float s = 0.f, cnt = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) { float v = p[c]; s += expf(v - mx); cnt += (v > zt) ? 1.f : 0.f; }
  s = block_reduce(s, false, sm);
  cnt = block_reduce(cnt, false, sm);
  if (threadIdx.x == 0) {
    float l = mx + logf(s);
    lse[b] = l;
    loss_rows[b] = l - zt;
    if (rank) rank[b] = (int)(cnt + 0.5f);
    if (loss_mean) atomicAdd(loss_mean, (l - zt) / (float)B);
  }
}
SVK_API int svk_ce_fwd(const float* z, const long long* label, float* loss_rows, float* lse, int* rank,
                       float* loss_mean, int B, int C, void* stream) {
  SVK_REQUIRE(z && label && loss_rows && lse && B > 0 && C > 0, SVK_E_BADARG, "ce_fwd: bad args");
  svk_launch(ce_fwd_kernel, B, 256, 0, as_stream(stream), z, label, loss_rows, lse, rank, loss_mean, B, C);
  SVK_LAUNCH_CHECK("ce_fwd");
  return 0;
}
__global__ void __launch_bounds__(256) ce_bwd_kernel(const float* __restrict__ z, const long long* __restrict__ label,
                                                     const float* __restrict__ lse, const float* __restrict__ gout,
                                                     float mult, float* __restrict__ g, int B, int C) {
  pdl_prologue();
  long long n = (long long)B * C;
  float gs = *gout * mult;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int b = (int)(i / C), c = (int)(i % C);
    float p = expf(z[i] - lse[b]);
    g[i] = (p - ((long long)c == label[b] ? 1.f : 0.f)) * gs;
  }
}
SVK_API int svk_ce_bwd(const float* z, const long long* label, const float* lse, const float* gout, float mult,
                       float* g, int B, int C, void* stream) {
  SVK_REQUIRE(z && label && lse && gout && g && B > 0 && C > 0, SVK_E_BADARG, "ce_bwd: bad args");
  long long n = (long long)B * C; long long b = (n + 255) / 256; long long cap = (long long)svk_num_sms() * 8; if (b > cap) b = cap;
  svk_launch(ce_bwd_kernel, (int)b, 256, 0, as_stream(stream), z, label, lse, gout, mult, g, B, C);
  SVK_LAUNCH_CHECK("ce_bwd");
  return 0;
}

// ------------------------------------------------------------------------------------------ cosine pairs
// One warp per trial; both embeddings are mean-subtracted on the fly (cosine_score.py:52-56, 62-64).
__global__ void __launch_bounds__(256) cosine_pairs_kernel(const float* __restrict__ E, const float* __restrict__ T,
                                                           const float* __restrict__ mean, const int* __restrict__ ie,
                                                           const int* __restrict__ it, float* __restrict__ score,
                                                           long long ntrials, int D) {
  pdl_prologue();
  long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long t = warp; t < ntrials; t += nwarps) {
    const float* a = E + (long long)ie[t] * D;
    const float* b = T + (long long)it[t] * D;
    float dot = 0.f, na = 0.f, nb = 0.f;
    for (int c = lane; c < D; c += 32) {
      float m = mean ? mean[c] : 0.f;
      float x = a[c] - m, y = b[c] - m;
      dot = fmaf(x, y, dot); na = fmaf(x, x, na); nb = fmaf(y, y, nb);
    }
    dot = warp_sum(dot); na = warp_sum(na); nb = warp_sum(nb);
    // torch cosine_similarity: x.y / max(||x||*||y||, eps)  with eps = 1e-8
    if (lane == 0) score[t] = dot / fmaxf(sqrtf(na) * sqrtf(nb), 1e-8f);
  }
}
SVK_API int svk_cosine_score_pairs(const float* E, const float* T, const float* mean, const int* ie, const int* it,
                                   float* score, long long ntrials, int D, void* stream) {
  SVK_REQUIRE(E && T && ie && it && score && ntrials > 0 && D > 0, SVK_E_BADARG, "cosine_score_pairs: bad args");
  long long b = (ntrials + 7) / 8; long long cap = (long long)svk_num_sms() * 8; if (b > cap) b = cap;
  svk_launch(cosine_pairs_kernel, (int)b, 256, 0, as_stream(stream), E, T, mean, ie, it, score, ntrials, D);
  SVK_LAUNCH_CHECK("cosine_score_pairs");
  return 0;
}

// ------------------------------------------------------------------------------------------ top-k mean / std
// One block per row.  3-level radix select (11+11+10 bits of the order-preserving key) finds the k-th largest value,
// then one pass sums values above it (+ the needed copies of the threshold itself).
__device__ inline unsigned f2key(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ inline float key2f(unsigned k) {
  unsigned u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}
// General path: 3-level radix select (11 + 11 + 10 bits) + one summation pass — any k, any value distribution.
__device__ void topk_radix_row(const float* __restrict__ p, int ncoh, int topk, float* __restrict__ mean_out,
                               float* __restrict__ std_out) {
  __shared__ int hist[2048];
  __shared__ int part[256];
  __shared__ unsigned s_prefix;
  __shared__ int s_k;
  __shared__ double s_sum[8], s_sq[8];
  __shared__ int s_cnt[8];
  const int t = threadIdx.x;
  if (t == 0) { s_prefix = 0u; s_k = topk; }
  const int shifts[3] = {21, 10, 0};
  const int bits[3] = {11, 11, 10};
  unsigned mask_hi = 0u;   // bits already fixed
  for (int pass = 0; pass < 3; ++pass) {
    for (int i = t; i < 2048; i += 256) hist[i] = 0;
    __syncthreads();
    unsigned prefix = s_prefix;
    int nb = 1 << bits[pass];
    for (int i = t; i < ncoh; i += 256) {
      unsigned k = f2key(p[i]);
      if ((k & mask_hi) == prefix) atomicAdd(&hist[(k >> shifts[pass]) & (nb - 1)], 1);
    }
    __syncthreads();
    // suffix counts: thread t owns bins [8t, 8t+8) (nb <= 2048)
    int loc = 0;
    for (int j = 0; j < 8; ++j) { int bidx = t * 8 + j; if (bidx < nb) loc += hist[bidx]; }
    part[t] = loc;
    __syncthreads();
    for (int off = 1; off < 256; off <<= 1) {          // inclusive suffix scan
      int v = (t + off < 256) ? part[t + off] : 0;
      __syncthreads();
      part[t] += v;
      __syncthreads();
    }
    int above = (t + 1 < 256) ? part[t + 1] : 0;       // elements in bins owned by higher threads
    int k_need = s_k;
    __syncthreads();
    if (above < k_need && part[t] >= k_need) {          // the k-th largest lives in one of my bins
      int cum = above;
      for (int j = 7; j >= 0; --j) {
        int bidx = t * 8 + j;
        if (bidx >= nb) continue;
        int h = hist[bidx];
        if (cum + h >= k_need) {
          s_prefix = prefix | ((unsigned)bidx << shifts[pass]);
          s_k = k_need - cum;
          break;
        }
        cum += h;
      }
    }
    mask_hi |= ((unsigned)(nb - 1)) << shifts[pass];
    __syncthreads();
  }
  const unsigned kth = s_prefix;        // key of the k-th largest value
  const int n_eq = s_k;                 // how many copies of it belong to the top-k
  const float vth = key2f(kth);
  double sum = 0.0, sq = 0.0; int cnt = 0;
  for (int i = t; i < ncoh; i += 256) {
    float v = p[i];
    if (f2key(v) > kth) { sum += v; sq += (double)v * v; ++cnt; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    sq += __shfl_xor_sync(0xffffffffu, sq, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((t & 31) == 0) { s_sum[t >> 5] = sum; s_sq[t >> 5] = sq; s_cnt[t >> 5] = cnt; }
  __syncthreads();
  if (t == 0) {
    double S = 0.0, Q = 0.0;
    for (int w = 0; w < 8; ++w) { S += s_sum[w]; Q += s_sq[w]; }
    S += (double)n_eq * vth; Q += (double)n_eq * vth * vth;
    double m = S / topk;
    double var = (Q - S * m) / (double)(topk - 1);
    if (var < 0.0) var = 0.0;
    *mean_out = (float)m;
    *std_out = (float)sqrt(var);
  }
}

// Fast path for k <= 512 of a long row (the s-norm case: top-300 of 50,000 cosines).  The radix select above reads the row
// four times and funnels every element of a clustered distribution through a handful of shared-memory histogram bins;
// here the row is read twice and almost nothing is atomic:
//   1. every thread keeps the two largest keys of its strided slice (512 keys, all of them elements of the row);
//   2. tau = the k-th largest of those 512 (rank by counting, broadcast smem reads): a LOWER bound of the true k-th largest;
//   3. second pass: the elements >= tau are the only candidates (k <= count, typically ~1.3 k) -> smem list;
//   4. exact rank of every candidate inside the list (ties broken by list position), fp64 sums of ranks < k.
// A row whose candidate list overflows (adversarial data) takes the radix path.
constexpr int TK_SLOTS = 512, TK_CAND = 2048;
__global__ void __launch_bounds__(256) topk_meanstd_kernel(const float* __restrict__ scores, int ncoh, int topk,
                                                           float* __restrict__ mean, float* __restrict__ stdv) {
  pdl_prologue();
  const float* p = scores + (long long)blockIdx.x * ncoh;
  if (topk > TK_SLOTS || ncoh < 4 * TK_SLOTS) { topk_radix_row(p, ncoh, topk, mean + blockIdx.x, stdv + blockIdx.x); return; }
  __shared__ unsigned slot[TK_SLOTS];
  __shared__ unsigned cand[TK_CAND];
  __shared__ unsigned s_tau;
  __shared__ int s_n;
  __shared__ double r_sum[8], r_sq[8];
  const int t = threadIdx.x;
  unsigned k1 = 0u, k2 = 0u;                   // two largest keys of this thread's slice (key 0 sorts below every float)
  for (int i = t; i < ncoh; i += 256) {
    const unsigned k = f2key(p[i]);
    if (k > k1) { k2 = k1; k1 = k; } else if (k > k2) k2 = k;
  }
  slot[t] = k1; slot[256 + t] = k2;
  if (t == 0) s_n = 0;
  __syncthreads();
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int me = h * 256 + t;
    const unsigned k = slot[me];
    int rank = 0;
    for (int j = 0; j < TK_SLOTS; ++j) { const unsigned o = slot[j]; rank += (o > k || (o == k && j < me)) ? 1 : 0; }
    if (rank == topk - 1) s_tau = k;
  }
  __syncthreads();
  const unsigned tau = s_tau;
  for (int i = t; i < ncoh; i += 256) {
    const unsigned k = f2key(p[i]);
    if (k >= tau) { const int pos = atomicAdd(&s_n, 1); if (pos < TK_CAND) cand[pos] = k; }
  }
  __syncthreads();
  const int n = s_n;
  if (n > TK_CAND) { topk_radix_row(p, ncoh, topk, mean + blockIdx.x, stdv + blockIdx.x); return; }      // uniform per block
  double sum = 0.0, sq = 0.0;
  for (int me = t; me < n; me += 256) {
    const unsigned k = cand[me];
    int rank = 0;
    for (int j = 0; j < n; ++j) { const unsigned o = cand[j]; rank += (o > k || (o == k && j < me)) ? 1 : 0; }
    if (rank < topk) { const double v = (double)key2f(k); sum += v; sq += v * v; }
  }
  for (int o = 16; o > 0; o >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, o); sq += __shfl_xor_sync(0xffffffffu, sq, o); }
  if ((t & 31) == 0) { r_sum[t >> 5] = sum; r_sq[t >> 5] = sq; }
  __syncthreads();
  if (t == 0) {
    double S = 0.0, Q = 0.0;
    for (int w = 0; w < 8; ++w) { S += r_sum[w]; Q += r_sq[w]; }
    const double m = S / topk;
    double var = (Q - S * m) / (double)(topk - 1);
    if (var < 0.0) var = 0.0;
    mean[blockIdx.x] = (float)m;
    stdv[blockIdx.x] = (float)sqrt(var);
  }
}
SVK_API int svk_topk_meanstd(const float* scores, int rows, int ncoh, int topk, float* mean, float* stdv, void* stream) {
  SVK_REQUIRE(scores && mean && stdv && rows > 0, SVK_E_BADARG, "topk_meanstd: bad args");
  SVK_REQUIRE(topk >= 2 && topk <= ncoh, SVK_E_BADARG, "topk_meanstd: need 2 <= topk (%d) <= ncoh (%d)", topk, ncoh);
  svk_launch(topk_meanstd_kernel, rows, 256, 0, as_stream(stream), scores, ncoh, topk, mean, stdv);
  SVK_LAUNCH_CHECK("topk_meanstd");
  return 0;
}

// ------------------------------------------------------------------------------------------ adaptive s-norm
__global__ void __launch_bounds__(256) snorm_kernel(const float* __restrict__ score, const int* __restrict__ ie,
                                                    const int* __restrict__ it, const float* __restrict__ me,
                                                    const float* __restrict__ se, const float* __restrict__ mt,
                                                    const float* __restrict__ st, float* __restrict__ out, long long n) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float s = score[i];
    int e = ie[i], t = it[i];
    out[i] = (s - me[e]) / fmaxf(se[e], 1e-8f) / 2.f + (s - mt[t]) / fmaxf(st[t], 1e-8f) / 2.f;
  }
}
SVK_API int svk_snorm_apply(const float* score, const int* ie, const int* it, const float* me, const float* se,
                            const float* mt, const float* st, float* out, long long n, void* stream) {
  SVK_REQUIRE(score && ie && it && me && se && mt && st && out && n > 0, SVK_E_BADARG, "snorm_apply: bad args");
  long long b = (n + 255) / 256; long long cap = (long long)svk_num_sms() * 8; if (b > cap) b = cap;
  svk_launch(snorm_kernel, (int)b, 256, 0, as_stream(stream), score, ie, it, me, se, mt, st, out, n);
  SVK_LAUNCH_CHECK("snorm_apply");
  return 0;
}
