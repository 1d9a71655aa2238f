// HBM-bound per-channel / elementwise kernels over NHWC activations viewed as [M, C] (C innermost):
// BatchNorm statistics, finalise, apply(+residual)(+ReLU), backward reductions + apply, gradient merges,
// weight (un)packing, SGD, casts.  All accesses are 16-byte vectors; grids are sized in multiples of the SM count.
#include "svk_common.cuh"

static constexpr int EW_THREADS = 256;
static inline int ew_grid(long long nvec) {
  long long b = (nvec + EW_THREADS - 1) / EW_THREADS;
  long long cap = (long long)svk_num_sms() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}
static inline bool pow2(int c) { return c > 0 && (c & (c - 1)) == 0; }

// ---------------------------------------------------------------------------------- weight packing
template <typename T>
__global__ void pack_w_kernel(const float* __restrict__ w, T* __restrict__ wf, T* __restrict__ wd, int Cout, int Cin,
                              int taps) {
  pdl_prologue();
  long long n = (long long)Cout * Cin * taps;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int t = (int)(i % taps);
    long long r = i / taps;
    int ci = (int)(r % Cin);
    int co = (int)(r / Cin);
    float v = w[i];  // OIHW: ((co*Cin + ci)*taps + t)
    if (wf) wf[((long long)t * Cout + co) * Cin + ci] = from_f<T>(v);
    if (wd) wd[((long long)t * Cin + ci) * Cout + co] = from_f<T>(v);
  }
}
SVK_API int svk_pack_conv_weight(const float* w, void* wf, void* wd, int Cout, int Cin, int R, int dtype, void* stream) {
  SVK_REQUIRE(w && (wf || wd) && Cout > 0 && Cin > 0 && (R == 1 || R == 3), SVK_E_BADARG, "pack_conv_weight: bad args");
  long long n = (long long)Cout * Cin * R * R;
  SVK_DISPATCH_DTYPE(dtype, "pack_conv_weight",
    svk_launch(pack_w_kernel<T>, ew_grid(n), EW_THREADS, 0, as_stream(stream), w, (T*)wf, (T*)wd, Cout, Cin, R * R);)
  SVK_LAUNCH_CHECK("pack_conv_weight");
  return 0;
}
// ---------------------------------------------------------------------------------- per-channel reductions
// Thread (lane_c, lane_r): lane_c = fixed vector of V channels, rows strided.  NACC accumulators per channel.
template <typename T, int NACC, typename F>
__device__ inline void channel_reduce(long long M, int C, double* __restrict__ sums, F f) {
  constexpr int V = Vec<T>::N;
  const int lanes_c = C / V;
  const int lane_c = threadIdx.x % lanes_c;
  const int rows_blk = EW_THREADS / lanes_c;
  const int lane_r = threadIdx.x / lanes_c;
  float acc[NACC][V];
#pragma unroll
  for (int a = 0; a < NACC; ++a)
#pragma unroll
    for (int i = 0; i < V; ++i) acc[a][i] = 0.f;
  for (long long r = (long long)blockIdx.x * rows_blk + lane_r; r < M; r += (long long)gridDim.x * rows_blk)
    f(r * C + lane_c * V, lane_c * V, acc);
  __shared__ float red[EW_THREADS * V];
  for (int a = 0; a < NACC; ++a) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < V; ++i) red[threadIdx.x * V + i] = acc[a][i];
    __syncthreads();
    // thread t < C sums channel t over the rows_blk row-lanes
    for (int ch = threadIdx.x; ch < C; ch += EW_THREADS) {
      int lc = ch / V, i = ch % V;
      float s = 0.f;
      for (int rr = 0; rr < rows_blk; ++rr) s += red[(rr * lanes_c + lc) * V + i];
      atomicAdd(&sums[a * C + ch], (double)s);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(EW_THREADS) channel_stats_kernel(const T* __restrict__ x, long long M, int C,
                                                                  double* __restrict__ stats) {
  pdl_prologue();
  constexpr int V = Vec<T>::N;
  channel_reduce<T, 2>(M, C, stats, [&](long long off, int, float (&acc)[2][V]) {
    float v[V];
    Vec<T>::load(x + off, v);
#pragma unroll
    for (int i = 0; i < V; ++i) { acc[0][i] += v[i]; acc[1][i] += v[i] * v[i]; }
  });
}
static int check_mc(const char* name, long long M, int C, int dtype) {
  int V = dtype == SVK_BF16 ? 8 : 4;
  SVK_REQUIRE(M > 0 && pow2(C) && C >= V && C / V <= EW_THREADS, SVK_E_UNSUPPORTED,
              "%s: need power-of-two C in [%d, %d], got M=%lld C=%d", name, V, EW_THREADS * V, M, C);
  return 0;
}
static inline int red_grid(long long M, int C, int V) {
  int rows_blk = EW_THREADS / (C / V);
  long long b = (M + rows_blk - 1) / rows_blk;
  // enough rows per thread to amortise the block reduction; multiple of the SM count
  long long cap = (long long)svk_num_sms() * 4;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}
SVK_API int svk_channel_stats(const void* x, long long M, int C, int dtype, double* stats, void* stream) {
  SVK_REQUIRE(x && stats, SVK_E_BADARG, "channel_stats: null pointer");
  if (int e = check_mc("channel_stats", M, C, dtype)) return e;
  SVK_DISPATCH_DTYPE(dtype, "channel_stats",
    svk_launch(channel_stats_kernel<T>, red_grid(M, C, Vec<T>::N), EW_THREADS, 0, as_stream(stream), (const T*)x, M, C, stats);)
  SVK_LAUNCH_CHECK("channel_stats");
  return 0;
}

__global__ void bn_finalize_kernel(const double* __restrict__ stats, long long M, int C, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ rm, float* __restrict__ rv,
                                   float momentum, float eps, float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ save_mean, float* __restrict__ save_rstd) {
  pdl_prologue();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double mean = stats[c] / (double)M;
  double var = stats[C + c] / (double)M - mean * mean;
  if (var < 0.0) var = 0.0;
  float rstd = (float)(1.0 / sqrt(var + (double)eps));
  float sc = gamma[c] * rstd;
  scale[c] = sc;
  shift[c] = beta[c] - (float)mean * sc;
  if (save_mean) save_mean[c] = (float)mean;
  if (save_rstd) save_rstd[c] = rstd;
  if (rm) rm[c] = (1.f - momentum) * rm[c] + momentum * (float)mean;
  if (rv) {
    double unb = M > 1 ? var * ((double)M / (double)(M - 1)) : var;
    rv[c] = (1.f - momentum) * rv[c] + momentum * (float)unb;
  }
}
SVK_API int svk_bn_finalize(const double* stats, long long M, int C, const float* gamma, const float* beta, float* rm,
                            float* rv, float momentum, float eps, float* scale, float* shift, float* save_mean,
                            float* save_rstd, void* stream) {
  SVK_REQUIRE(stats && gamma && beta && scale && shift && M > 0 && C > 0, SVK_E_BADARG, "bn_finalize: bad args");
  svk_launch(bn_finalize_kernel, (C + 127) / 128, 128, 0, as_stream(stream), stats, M, C, gamma, beta, rm, rv, momentum, eps,
                                                                    scale, shift, save_mean, save_rstd);
  SVK_LAUNCH_CHECK("bn_finalize");
  return 0;
}
__global__ void bn_eval_coeffs_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ rm, const float* __restrict__ rv, float eps, int C,
                                      float* __restrict__ scale, float* __restrict__ shift) {
  pdl_prologue();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float sc = gamma[c] / sqrtf(rv[c] + eps);
  scale[c] = sc;
  shift[c] = beta[c] - rm[c] * sc;
}
SVK_API int svk_bn_eval_coeffs(const float* gamma, const float* beta, const float* rm, const float* rv, float eps,
                               int C, float* scale, float* shift, void* stream) {
  SVK_REQUIRE(gamma && beta && rm && rv && scale && shift && C > 0, SVK_E_BADARG, "bn_eval_coeffs: bad args");
  svk_launch(bn_eval_coeffs_kernel, (C + 127) / 128, 128, 0, as_stream(stream), gamma, beta, rm, rv, eps, C, scale, shift);
  SVK_LAUNCH_CHECK("bn_eval_coeffs");
  return 0;
}

// ---------------------------------------------------------------------------------- BN apply (+res)(+ReLU)
// grid-stride step is a multiple of C/V, so each thread's channel vector is fixed and its coefficients hoisted.
template <typename T, int RES /*0 none, 1 plain, 2 affine*/>
__global__ void __launch_bounds__(EW_THREADS)
bn_act_fwd_kernel(const T* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                  const T* __restrict__ res, const float* __restrict__ rscale, const float* __restrict__ rshift,
                  int relu, T* __restrict__ out, long long nvec, int C) {
  pdl_prologue();
  constexpr int V = Vec<T>::N;
  const int lanes_c = C / V;
  long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = (int)(i0 % lanes_c) * V;
  float sc[V], sh[V], rs[V], rh[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    sc[i] = scale[c0 + i]; sh[i] = shift[c0 + i];
    if (RES == 2) { rs[i] = rscale[c0 + i]; rh[i] = rshift[c0 + i]; }
  }
  for (long long iv = i0; iv < nvec; iv += (long long)gridDim.x * blockDim.x) {
    float v[V], r[V];
    Vec<T>::load(x + iv * V, v);
    if (RES) Vec<T>::load(res + iv * V, r);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float y = fmaf(v[i], sc[i], sh[i]);
      if (RES == 1) y += r[i];
      if (RES == 2) y += fmaf(r[i], rs[i], rh[i]);
      v[i] = relu ? fmaxf(y, 0.f) : y;
    }
    Vec<T>::store(out + iv * V, v);
  }
}
SVK_API int svk_bn_act_fwd(const void* x, const float* scale, const float* shift, const void* res, const float* rscale,
                           const float* rshift, int relu, void* out, long long M, int C, int dtype, void* stream) {
  SVK_REQUIRE(x && scale && shift && out, SVK_E_BADARG, "bn_act_fwd: null pointer");
  SVK_REQUIRE((rscale == nullptr) == (rshift == nullptr) && (!rscale || res), SVK_E_BADARG, "bn_act_fwd: bad residual args");
  if (int e = check_mc("bn_act_fwd", M, C, dtype)) return e;
  SVK_DISPATCH_DTYPE(dtype, "bn_act_fwd",
    long long nvec = M * C / Vec<T>::N;
    int g = ew_grid(nvec);
    if (!res) svk_launch(bn_act_fwd_kernel<T, 0>, g, EW_THREADS, 0, as_stream(stream), (const T*)x, scale, shift, nullptr, nullptr, nullptr, relu, (T*)out, nvec, C);
    else if (!rscale) svk_launch(bn_act_fwd_kernel<T, 1>, g, EW_THREADS, 0, as_stream(stream), (const T*)x, scale, shift, (const T*)res, nullptr, nullptr, relu, (T*)out, nvec, C);
    else svk_launch(bn_act_fwd_kernel<T, 2>, g, EW_THREADS, 0, as_stream(stream), (const T*)x, scale, shift, (const T*)res, rscale, rshift, relu, (T*)out, nvec, C);)
  SVK_LAUNCH_CHECK("bn_act_fwd");
  return 0;
}

// ---------------------------------------------------------------------------------- BN backward
template <typename T, bool MASK, bool TWO>
__global__ void __launch_bounds__(EW_THREADS)
bn_bwd_reduce_kernel(const T* __restrict__ dout, const T* __restrict__ out, const T* __restrict__ c,
                     const float* __restrict__ mean, const float* __restrict__ rstd, const T* __restrict__ cb,
                     const float* __restrict__ meanb, const float* __restrict__ rstdb, double* __restrict__ sums,
                     long long M, int C) {
  pdl_prologue();
  constexpr int V = Vec<T>::N;
  constexpr int NACC = TWO ? 3 : 2;
  channel_reduce<T, NACC>(M, C, sums, [&](long long off, int c0, float (&acc)[NACC][V]) {
    float g[V], o[V], x[V], xb[V];
    Vec<T>::load(dout + off, g);
    if (MASK) Vec<T>::load(out + off, o);
    Vec<T>::load(c + off, x);
    if (TWO) Vec<T>::load(cb + off, xb);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float gi = (MASK && !(o[i] > 0.f)) ? 0.f : g[i];
      acc[0][i] += gi;
      acc[1][i] += gi * (x[i] - mean[c0 + i]) * rstd[c0 + i];
      if (TWO) acc[2][i] += gi * (xb[i] - meanb[c0 + i]) * rstdb[c0 + i];
    }
  });
}
template <typename T>
static void launch_bn_bwd_reduce(const void* dout, const void* out, const void* c, const float* mean, const float* rstd,
                                 const void* cb, const float* meanb, const float* rstdb, double* sums, long long M, int C,
                                 cudaStream_t s) {
  int g = red_grid(M, C, Vec<T>::N);
#define SVK_BR(MASK_, TWO_) svk_launch(bn_bwd_reduce_kernel<T, MASK_, TWO_>, g, EW_THREADS, 0, s, (const T*)dout, (const T*)out, (const T*)c, mean, rstd, (const T*)cb, meanb, rstdb, sums, M, C)
  if (out && cb) SVK_BR(true, true); else if (out) SVK_BR(true, false); else if (cb) SVK_BR(false, true); else SVK_BR(false, false);
#undef SVK_BR
}
SVK_API int svk_bn_bwd_reduce(const void* dout, const void* out, const void* c, const float* mean, const float* rstd,
                              const void* cb, const float* meanb, const float* rstdb, double* sums, long long M, int C,
                              int dtype, void* stream) {
  SVK_REQUIRE(dout && c && mean && rstd && sums, SVK_E_BADARG, "bn_bwd_reduce: null pointer");
  SVK_REQUIRE(!cb || (meanb && rstdb), SVK_E_BADARG, "bn_bwd_reduce: second BN needs mean/rstd");
  if (int e = check_mc("bn_bwd_reduce", M, C, dtype)) return e;
  SVK_DISPATCH_DTYPE(dtype, "bn_bwd_reduce",
    launch_bn_bwd_reduce<T>(dout, out, c, mean, rstd, cb, meanb, rstdb, sums, M, C, as_stream(stream));)
  SVK_LAUNCH_CHECK("bn_bwd_reduce");
  return 0;
}

template <typename T, bool MASK, bool TWO>
__global__ void __launch_bounds__(EW_THREADS)
bn_bwd_apply_kernel(const T* __restrict__ dout, const T* __restrict__ out, const T* __restrict__ c,
                    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                    T* __restrict__ dc, const T* __restrict__ cb, const float* __restrict__ meanb,
                    const float* __restrict__ rstdb, const float* __restrict__ gammab, T* __restrict__ dcb,
                    const double* __restrict__ sums, float* __restrict__ dgamma, float* __restrict__ dbeta,
                    float* __restrict__ dgammab, float* __restrict__ dbetab, long long nvec, long long M, int C) {
  pdl_prologue();
  constexpr int V = Vec<T>::N;
  extern __shared__ float s_k[];             // [10][C]: mu, rs, k0, k1, k2 and the same for the second BN
  const int lanes_c = C / V;
  long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = (int)(i0 % lanes_c) * V;
  const float invM = 1.f / (float)M;
  // one thread per channel derives the coefficients of  dc = k0*g - k1 - xhat*k2  once per block; block 0 also
  // publishes the parameter gradients (dbeta = sum g, dgamma = sum g*xhat)
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float s1 = (float)sums[c], s2 = (float)sums[C + c];
    const float r_ = rstd[c], k = gamma[c] * r_;
    s_k[c] = mean[c]; s_k[C + c] = r_; s_k[2 * C + c] = k; s_k[3 * C + c] = k * s1 * invM; s_k[4 * C + c] = k * s2 * invM;
    if (blockIdx.x == 0) {
      if (dbeta) dbeta[c] = s1;
      if (dgamma) dgamma[c] = s2;
    }
    if (TWO) {
      const float s3 = (float)sums[2 * C + c];
      const float rb = rstdb[c], kb = gammab[c] * rb;
      s_k[5 * C + c] = meanb[c]; s_k[6 * C + c] = rb; s_k[7 * C + c] = kb; s_k[8 * C + c] = kb * s1 * invM;
      s_k[9 * C + c] = kb * s3 * invM;
      if (blockIdx.x == 0) {
        if (dbetab) dbetab[c] = s1;
        if (dgammab) dgammab[c] = s3;
      }
    }
  }
  __syncthreads();
  float mu[V], rs[V], k0[V], k1[V], k2[V];    // dc = k0*g - k1 - xhat*k2
  float mub[V], rsb[V], kb0[V], kb1[V], kb2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = c0 + i;
    mu[i] = s_k[c]; rs[i] = s_k[C + c]; k0[i] = s_k[2 * C + c]; k1[i] = s_k[3 * C + c]; k2[i] = s_k[4 * C + c];
    if (TWO) {
      mub[i] = s_k[5 * C + c]; rsb[i] = s_k[6 * C + c]; kb0[i] = s_k[7 * C + c]; kb1[i] = s_k[8 * C + c]; kb2[i] = s_k[9 * C + c];
    }
  }
  for (long long iv = i0; iv < nvec; iv += (long long)gridDim.x * blockDim.x) {
    float g[V], o[V], x[V], xb[V], r[V];
    Vec<T>::load(dout + iv * V, g);
    if (MASK) Vec<T>::load(out + iv * V, o);
    Vec<T>::load(c + iv * V, x);
    if (TWO) Vec<T>::load(cb + iv * V, xb);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float gi = (MASK && !(o[i] > 0.f)) ? 0.f : g[i];
      g[i] = gi;
      r[i] = k0[i] * gi - k1[i] - (x[i] - mu[i]) * rs[i] * k2[i];
    }
    Vec<T>::store(dc + iv * V, r);
    if (TWO) {
#pragma unroll
      for (int i = 0; i < V; ++i) r[i] = kb0[i] * g[i] - kb1[i] - (xb[i] - mub[i]) * rsb[i] * kb2[i];
      Vec<T>::store(dcb + iv * V, r);
    }
  }
}
template <typename T>
static void launch_bn_bwd_apply(const void* dout, const void* out, const void* c, const float* mean, const float* rstd,
                                const float* gamma, void* dc, const void* cb, const float* meanb, const float* rstdb,
                                const float* gammab, void* dcb, const double* sums, float* dgamma, float* dbeta,
                                float* dgammab, float* dbetab, long long M, int C, cudaStream_t s) {
  long long nvec = M * C / Vec<T>::N;
  int g = ew_grid(nvec);
  const size_t sm = (size_t)(cb ? 10 : 5) * C * sizeof(float);
#define SVK_BA(MASK_, TWO_) svk_launch(bn_bwd_apply_kernel<T, MASK_, TWO_>, g, EW_THREADS, sm, s, (const T*)dout, (const T*)out, (const T*)c, mean, rstd, gamma, (T*)dc, (const T*)cb, meanb, rstdb, gammab, (T*)dcb, sums, dgamma, dbeta, dgammab, dbetab, nvec, M, C)
  if (out && cb) SVK_BA(true, true); else if (out) SVK_BA(true, false); else if (cb) SVK_BA(false, true); else SVK_BA(false, false);
#undef SVK_BA
}
SVK_API int svk_bn_bwd_apply(const void* dout, const void* out, const void* c, const float* mean, const float* rstd,
                             const float* gamma, void* dc, const void* cb, const float* meanb, const float* rstdb,
                             const float* gammab, void* dcb, const double* sums, float* dgamma, float* dbeta,
                             float* dgammab, float* dbetab, long long M, int C, int dtype, void* stream) {
  SVK_REQUIRE(dout && c && mean && rstd && gamma && dc && sums, SVK_E_BADARG, "bn_bwd_apply: null pointer");
  SVK_REQUIRE(!cb || (meanb && rstdb && gammab && dcb), SVK_E_BADARG, "bn_bwd_apply: second BN args incomplete");
  if (int e = check_mc("bn_bwd_apply", M, C, dtype)) return e;
  SVK_DISPATCH_DTYPE(dtype, "bn_bwd_apply",
    launch_bn_bwd_apply<T>(dout, out, c, mean, rstd, gamma, dc, cb, meanb, rstdb, gammab, dcb, sums, dgamma, dbeta,
                           dgammab, dbetab, M, C, as_stream(stream));)
  SVK_LAUNCH_CHECK("bn_bwd_apply");
  return 0;
}

// ---------------------------------------------------------------------------------- gradient merges
template <typename T, bool MASK>
__global__ void __launch_bounds__(EW_THREADS)
add_masked_kernel(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ mask, T* __restrict__ out,
                  long long nvec) {
  pdl_prologue();
  constexpr int V = Vec<T>::N;
  for (long long iv = (long long)blockIdx.x * blockDim.x + threadIdx.x; iv < nvec; iv += (long long)gridDim.x * blockDim.x) {
    float x[V], y[V], m[V];
    Vec<T>::load(a + iv * V, x);
    Vec<T>::load(b + iv * V, y);
    if (MASK) Vec<T>::load(mask + iv * V, m);
#pragma unroll
    for (int i = 0; i < V; ++i) x[i] += (MASK && !(m[i] > 0.f)) ? 0.f : y[i];
    Vec<T>::store(out + iv * V, x);
  }
}
SVK_API int svk_add_masked(const void* a, const void* b, const void* mask, void* out, long long n, int dtype, void* stream) {
  SVK_REQUIRE(a && b && out && n > 0, SVK_E_BADARG, "add_masked: bad args");
  SVK_DISPATCH_DTYPE(dtype, "add_masked",
    SVK_REQUIRE(n % Vec<T>::N == 0, SVK_E_ALIGN, "add_masked: n=%lld not a multiple of %d", n, Vec<T>::N);
    long long nvec = n / Vec<T>::N;
    if (mask) svk_launch(add_masked_kernel<T, true>, ew_grid(nvec), EW_THREADS, 0, as_stream(stream), (const T*)a, (const T*)b, (const T*)mask, (T*)out, nvec);
    else svk_launch(add_masked_kernel<T, false>, ew_grid(nvec), EW_THREADS, 0, as_stream(stream), (const T*)a, (const T*)b, nullptr, (T*)out, nvec);)
  SVK_LAUNCH_CHECK("add_masked");
  return 0;
}

template <typename T>
__global__ void __launch_bounds__(EW_THREADS) relu_mask_kernel(T* __restrict__ g, const T* __restrict__ mask, long long nvec) {
  pdl_prologue();
  constexpr int V = Vec<T>::N;
  for (long long iv = (long long)blockIdx.x * blockDim.x + threadIdx.x; iv < nvec; iv += (long long)gridDim.x * blockDim.x) {
    float x[V], m[V];
    Vec<T>::load(g + iv * V, x);
    Vec<T>::load(mask + iv * V, m);
#pragma unroll
    for (int i = 0; i < V; ++i) x[i] = (m[i] > 0.f) ? x[i] : 0.f;
    Vec<T>::store(g + iv * V, x);
  }
}
SVK_API int svk_relu_mask_inplace(void* g, const void* mask, long long n, int dtype, void* stream) {
  SVK_REQUIRE(g && mask && n > 0, SVK_E_BADARG, "relu_mask_inplace: bad args");
  SVK_DISPATCH_DTYPE(dtype, "relu_mask_inplace",
    SVK_REQUIRE(n % Vec<T>::N == 0, SVK_E_ALIGN, "relu_mask_inplace: n=%lld not a multiple of %d", n, Vec<T>::N);
    long long nvec = n / Vec<T>::N;
    svk_launch(relu_mask_kernel<T>, ew_grid(nvec), EW_THREADS, 0, as_stream(stream), (T*)g, (const T*)mask, nvec);)
  SVK_LAUNCH_CHECK("relu_mask_inplace");
  return 0;
}

template <typename T>
__global__ void __launch_bounds__(EW_THREADS)
add_strided2_kernel(T* __restrict__ dx, const T* __restrict__ d, int N, int H, int W, int Ho, int Wo, int C) {
  pdl_prologue();
  constexpr int V = Vec<T>::N;
  const int lanes_c = C / V;
  long long nvec = (long long)N * Ho * Wo * lanes_c;
  for (long long iv = (long long)blockIdx.x * blockDim.x + threadIdx.x; iv < nvec; iv += (long long)gridDim.x * blockDim.x) {
    int cv = (int)(iv % lanes_c);
    long long p = iv / lanes_c;
    int j = (int)(p % Wo); p /= Wo;
    int i = (int)(p % Ho);
    int n = (int)(p / Ho);
    T* dst = dx + (((long long)n * H + 2 * i) * W + 2 * j) * C + cv * V;
    float x[V], y[V];
    Vec<T>::load(dst, x);
    Vec<T>::load(d + iv * V, y);
#pragma unroll
    for (int k = 0; k < V; ++k) x[k] += y[k];
    Vec<T>::store(dst, x);
  }
}
SVK_API int svk_add_strided2(void* dx, const void* d, int N, int H, int W, int Ho, int Wo, int C, int dtype, void* stream) {
  SVK_REQUIRE(dx && d && N > 0 && H > 0 && W > 0 && Ho == (H + 1) / 2 && Wo == (W + 1) / 2, SVK_E_BADARG, "add_strided2: bad args");
  if (int e = check_mc("add_strided2", 1, C, dtype)) return e;
  SVK_DISPATCH_DTYPE(dtype, "add_strided2",
    long long nvec = (long long)N * Ho * Wo * (C / Vec<T>::N);
    svk_launch(add_strided2_kernel<T>, ew_grid(nvec), EW_THREADS, 0, as_stream(stream), (T*)dx, (const T*)d, N, H, W, Ho, Wo, C);)
  SVK_LAUNCH_CHECK("add_strided2");
  return 0;
}

// ---------------------------------------------------------------------------------- SGD + casts
__global__ void __launch_bounds__(EW_THREADS)
sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, long long n, float lr,
           float mom, float wd, float gscale) {
  pdl_prologue();
  long long n4 = n / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i], gv = reinterpret_cast<const float4*>(g)[i], bv = reinterpret_cast<float4*>(buf)[i];
    float* pp = &pv.x; const float* gg = &gv.x; float* bb = &bv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float d = fmaf(wd, pp[k], gg[k] * gscale);
      bb[k] = fmaf(mom, bb[k], d);
      pp[k] -= lr * bb[k];
    }
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(buf)[i] = bv;
  }
  // tail
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float d = fmaf(wd, p[i], g[i] * gscale);
    float b = fmaf(mom, buf[i], d);
    buf[i] = b;
    p[i] -= lr * b;
  }
}
SVK_API int svk_sgd_step(float* p, const float* g, float* buf, long long n, float lr, float momentum, float wd,
                         float gscale, void* stream) {
  SVK_REQUIRE(p && g && buf && n > 0, SVK_E_BADARG, "sgd_step: bad args");
  SVK_REQUIRE(aligned16(p) && aligned16(g) && aligned16(buf), SVK_E_ALIGN, "sgd_step: buffers must be 16-byte aligned");
  svk_launch(sgd_kernel, ew_grid(n / 4 + 1), EW_THREADS, 0, as_stream(stream), p, g, buf, n, lr, momentum, wd, gscale);
  SVK_LAUNCH_CHECK("sgd_step");
  return 0;
}

template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ s, D* __restrict__ d, long long n) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    d[i] = from_f<D>(to_f(s[i]));
}
SVK_API int svk_cast(const void* src, void* dst, long long n, int sd, int dd, void* stream) {
  SVK_REQUIRE(src && dst && n > 0, SVK_E_BADARG, "cast: bad args");
  int g = ew_grid(n);
  cudaStream_t s = as_stream(stream);
  if (sd == SVK_F32 && dd == SVK_BF16) svk_launch(cast_kernel<float, __nv_bfloat16>, g, EW_THREADS, 0, s, (const float*)src, (__nv_bfloat16*)dst, n);
  else if (sd == SVK_BF16 && dd == SVK_F32) svk_launch(cast_kernel<__nv_bfloat16, float>, g, EW_THREADS, 0, s, (const __nv_bfloat16*)src, (float*)dst, n);
  else if (sd == SVK_F32 && dd == SVK_F32) svk_launch(cast_kernel<float, float>, g, EW_THREADS, 0, s, (const float*)src, (float*)dst, n);
  else if (sd == SVK_BF16 && dd == SVK_BF16) svk_launch(cast_kernel<__nv_bfloat16, __nv_bfloat16>, g, EW_THREADS, 0, s, (const __nv_bfloat16*)src, (__nv_bfloat16*)dst, n);
  else { svk_set_error("cast: bad dtypes %d -> %d", sd, dd); return SVK_E_BADARG; }
  SVK_LAUNCH_CHECK("cast");
  return 0;
}

// ---------------------------------------------------------------------------------- batched weight packing
// One launch packs every conv of the network: table rows = {w_off, p_off, Cout, Cin, taps, start} (int64), `start` =
// cumulative element index.  replaces 35 svk_pack_conv_weight launches per training step.
// Work unit = one 32 x 32 (Cout x Cin) tile of one conv, all taps: the OIHW rows are read as contiguous runs of
// 32*taps floats into shared memory and both packed layouts are written as 64-byte runs (the element-per-thread version
// wrote 2-byte values Cout*Cin elements apart: 0.2 ms per step for 21 MB).
template <typename T>
__global__ void __launch_bounds__(EW_THREADS)
pack_all_kernel(const float* __restrict__ flat, T* __restrict__ wf, T* __restrict__ wd, const long long* __restrict__ table,
                int nconv) {
  pdl_prologue();
  __shared__ float tile[32 * (32 * 9 + 1)];
  __shared__ int s_first[65];                 // first unit of conv k (nconv <= 64), s_first[nconv] = number of units
  if (threadIdx.x < nconv)      // tile count of conv k (one round trip for the whole table), then an exclusive scan
    s_first[threadIdx.x + 1] = (int)((table[threadIdx.x * 6 + 2] + 31) / 32) * (int)((table[threadIdx.x * 6 + 3] + 31) / 32);
  __syncthreads();
  if (threadIdx.x == 0) {
    int u = 0;
    for (int k = 0; k < nconv; ++k) { const int n = s_first[k + 1]; s_first[k] = u; u += n; }
    s_first[nconv] = u;
  }
  __syncthreads();
  const int units = s_first[nconv];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
    int k = 0;
    while (s_first[k + 1] <= unit) ++k;
    const long long* e = table + k * 6;
    const int Cout = (int)e[2], Cin = (int)e[3], taps = (int)e[4];
    const int tiles_ci = (Cin + 31) / 32;
    const int local = unit - s_first[k];
    const int co0 = (local / tiles_ci) * 32, ci0 = (local % tiles_ci) * 32;
    const int nci = min(32, Cin - ci0), nco = min(32, Cout - co0);
    const int pitch = 32 * taps + 1, run = nci * taps;
    const float* src = flat + e[0];
    __syncthreads();                           // previous unit's readers are done with the tile
    for (int r = warp; r < nco; r += EW_THREADS / 32) {
      const float* row = src + ((long long)(co0 + r) * Cin + ci0) * taps;
      for (int i = lane; i < run; i += 32) tile[r * pitch + i] = row[i];
    }
    __syncthreads();
    T* f = wf + e[1];
    T* d = wd + e[1];
    for (int j = warp; j < taps * 32; j += EW_THREADS / 32) {
      const int t = j / 32, r = j % 32;        // r = co for the fprop layout, ci for the dgrad layout
      if (r < nco && lane < nci) f[((long long)t * Cout + co0 + r) * Cin + ci0 + lane] = from_f<T>(tile[r * pitch + lane * taps + t]);
      if (r < nci && lane < nco) d[((long long)t * Cin + ci0 + r) * Cout + co0 + lane] = from_f<T>(tile[lane * pitch + r * taps + t]);
    }
  }
}
SVK_API int svk_pack_conv_weights_batched(const float* flat, void* wf, void* wd, const long long* table, int nconv,
                                          long long total, int dtype, void* stream) {
  SVK_REQUIRE(flat && wf && wd && table && nconv > 0 && nconv <= 64 && total > 0, SVK_E_BADARG,
              "pack_conv_weights_batched: bad args (at most 64 convs)");
  // one block per 32x32 tile when there are few, a grid-stride over them otherwise (taps >= 1: total/1024 bounds the count)
  long long nb = total / 1024 + nconv; long long cap = (long long)svk_num_sms() * 6; if (nb > cap) nb = cap;
  SVK_DISPATCH_DTYPE(dtype, "pack_conv_weights_batched",
    svk_launch(pack_all_kernel<T>, (int)nb, EW_THREADS, 0, as_stream(stream), flat, (T*)wf, (T*)wd, table, nconv);)
  SVK_LAUNCH_CHECK("pack_conv_weights_batched");
  return 0;
}

// ---------------------------------------------------------------------------------- training BN: finalise + apply fused
// Each block derives the coefficients of all C channels once from the fp64 sums into shared memory (one thread per
// channel), so the 1-block finalise launch disappears and the per-thread preamble is a few smem reads — with every
// thread deriving its own channels the preamble was ~15 us of a 21 us launch on the small late-stage tensors
// (profiles/r01_bn_small.md).  Block 0 also publishes scale/shift/mean/rstd (for backward) and updates the running
// statistics.
struct BnTrainArgs {
  const double* stats; const float* gamma; const float* beta; float* rm; float* rv; float* coef; 
};
__device__ inline void bn_train_coefs_block(const BnTrainArgs& a, long long M, int C, int cstride, float momentum, float eps,
                                            bool publish, float* __restrict__ s_sc, float* __restrict__ s_sh) {
  const double invM = 1.0 / (double)M;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    // fp64 only where cancellation matters (E[x^2] - E[x]^2); the division / square root run in fp32
    double mean = a.stats[c] * invM;
    double var = fma(-mean, mean, a.stats[C + c] * invM);
    if (var < 0.0) var = 0.0;
    float rstd = 1.0f / sqrtf((float)var + eps);
    float sc = a.gamma[c] * rstd;
    float sh = a.beta[c] - (float)mean * sc;
    s_sc[c] = sc; s_sh[c] = sh;
    if (publish) {
      a.coef[c] = sc; a.coef[cstride + c] = sh; a.coef[2 * cstride + c] = (float)mean; a.coef[3 * cstride + c] = rstd;
      if (a.rm) a.rm[c] = (1.f - momentum) * a.rm[c] + momentum * (float)mean;
      if (a.rv) {
        double unb = M > 1 ? var * ((double)M / (double)(M - 1)) : var;
        a.rv[c] = (1.f - momentum) * a.rv[c] + momentum * (float)unb;
      }
    }
  }
}
template <typename T, int RES /*0 none, 1 plain, 2 second BN*/>
__global__ void __launch_bounds__(EW_THREADS)
bn_train_act_kernel(const T* __restrict__ x, BnTrainArgs a, const T* __restrict__ res, BnTrainArgs b, float momentum,
                    float eps, int relu, T* __restrict__ out, long long nvec, long long M, int C, int cstride) {
  pdl_prologue();
  constexpr int V = Vec<T>::N;
  extern __shared__ float s_coef[];          // [4][C]: scale, shift, (second BN) scale, shift
  const int lanes_c = C / V;
  long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = (int)(i0 % lanes_c) * V;
  bn_train_coefs_block(a, M, C, cstride, momentum, eps, blockIdx.x == 0, s_coef, s_coef + C);
  if (RES == 2) bn_train_coefs_block(b, M, C, cstride, momentum, eps, blockIdx.x == 0, s_coef + 2 * C, s_coef + 3 * C);
  __syncthreads();
  float sc[V], sh[V], rs[V], rh[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    sc[i] = s_coef[c0 + i]; sh[i] = s_coef[C + c0 + i];
    if (RES == 2) { rs[i] = s_coef[2 * C + c0 + i]; rh[i] = s_coef[3 * C + c0 + i]; }
  }
  for (long long iv = i0; iv < nvec; iv += (long long)gridDim.x * blockDim.x) {
    float v[V], r[V];
    Vec<T>::load(x + iv * V, v);
    if (RES) Vec<T>::load(res + iv * V, r);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float y = fmaf(v[i], sc[i], sh[i]);
      if (RES == 1) y += r[i];
      if (RES == 2) y += fmaf(r[i], rs[i], rh[i]);
      v[i] = relu ? fmaxf(y, 0.f) : y;
    }
    Vec<T>::store(out + iv * V, v);
  }
}
SVK_API int svk_bn_train_act_fwd(const void* x, const double* stats, const float* gamma, const float* beta, float* rm,
                                 float* rv, float* coef, const void* res, const double* stats_b, const float* gamma_b,
                                 const float* beta_b, float* rm_b, float* rv_b, float* coef_b, int cstride, float momentum,
                                 float eps, int relu, void* out, long long M, int C, int dtype, void* stream) {
  SVK_REQUIRE(x && stats && gamma && beta && coef && out && cstride >= C, SVK_E_BADARG, "bn_train_act_fwd: bad args");
  SVK_REQUIRE(!stats_b || (res && gamma_b && beta_b && coef_b), SVK_E_BADARG, "bn_train_act_fwd: second BN args incomplete");
  if (int e = check_mc("bn_train_act_fwd", M, C, dtype)) return e;
  BnTrainArgs a{stats, gamma, beta, rm, rv, coef};
  BnTrainArgs b{stats_b, gamma_b, beta_b, rm_b, rv_b, coef_b};
  SVK_DISPATCH_DTYPE(dtype, "bn_train_act_fwd",
    long long nvec = M * C / Vec<T>::N;
    int g = ew_grid(nvec);
    cudaStream_t s = as_stream(stream);
    const size_t sm = (size_t)4 * C * sizeof(float);
    if (!res) svk_launch(bn_train_act_kernel<T, 0>, g, EW_THREADS, sm, s, (const T*)x, a, nullptr, b, momentum, eps, relu, (T*)out, nvec, M, C, cstride);
    else if (!stats_b) svk_launch(bn_train_act_kernel<T, 1>, g, EW_THREADS, sm, s, (const T*)x, a, (const T*)res, b, momentum, eps, relu, (T*)out, nvec, M, C, cstride);
    else svk_launch(bn_train_act_kernel<T, 2>, g, EW_THREADS, sm, s, (const T*)x, a, (const T*)res, b, momentum, eps, relu, (T*)out, nvec, M, C, cstride);)
  SVK_LAUNCH_CHECK("bn_train_act_fwd");
  return 0;
}
