// Temporal statistics pooling (fwd/bwd), a strided fp32 SGEMM (embedding FC, AAM cosine GEMMs, cohort scores)
// and a column sum.
#include "svk_common.cuh"

// ------------------------------------------------------------------------------------------ stats pooling
// One thread per (n, h, c); consecutive threads = consecutive channels => every load is a coalesced row of C values.
// Two passes over W (mean, then centred second moment), fp32; the second pass hits L1/L2.
template <typename T>
__global__ void __launch_bounds__(256) statspool_fwd_kernel(const T* __restrict__ x, float* __restrict__ out, int N,
                                                            int H, int W, int C, int mode,
                                                            const int* __restrict__ valid_w) {
  pdl_prologue();
  long long total = (long long)N * H * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C); long long q = i / C; int h = (int)(q % H); int n = (int)(q / H);
    int Wv = valid_w ? valid_w[n] : W;
    const T* p = x + (((long long)n * H + h) * W) * C + c;
    float s = 0.f;
    for (int w = 0; w < Wv; ++w) s += to_f(p[(long long)w * C]);
    float mean = s / (float)Wv;
    if (mode == 0) {
      out[(long long)n * C * H + (long long)c * H + h] = mean;
    } else {
      float m2 = 0.f;
      for (int w = 0; w < Wv; ++w) { float d = to_f(p[(long long)w * C]) - mean; m2 = fmaf(d, d, m2); }
      float* o = out + (long long)n * C * 2 * H + (long long)c * 2 * H;
      o[h] = m2 / (float)(Wv - 1);          // unbiased variance (W = 1 -> NaN, as torch.var_mean)
      o[H + h] = sqrtf(mean);               // sqrt of the mean: the reference's swapped unpack, model.py:450-452
    }
  }
}
SVK_API int svk_statspool_fwd(const void* x, float* out, int N, int H, int W, int C, int mode, const int* valid_w,
                              int dtype, void* stream) {
  SVK_REQUIRE(x && out && N > 0 && H > 0 && W > 0 && C > 0 && (mode == 0 || mode == 1), SVK_E_BADARG, "statspool_fwd: bad args");
  long long total = (long long)N * H * C;
  long long b = (total + 255) / 256; long long cap = (long long)svk_num_sms() * 8; if (b > cap) b = cap;
  SVK_DISPATCH_DTYPE(dtype, "statspool_fwd",
    svk_launch(statspool_fwd_kernel<T>, (int)b, 256, 0, as_stream(stream), (const T*)x, out, N, H, W, C, mode, valid_w);)
  SVK_LAUNCH_CHECK("statspool_fwd");
  return 0;
}

// dx[n,h,w,c] = dvar * 2 (x - mean) / (W-1) + dsm / (2 sqrt(mean) W)   [mode 1]   |   dmean / W   [mode 0]
// d sqrt(0) := 0 (the reference's inf is always multiplied by a zero ReLU mask).
// relu_mask != 0: x is the output of a ReLU and dx is additionally multiplied by (x > 0), i.e. the gradient leaves this
// kernel already masked for the BatchNorm backward of the last block.
template <typename T>
__global__ void __launch_bounds__(256) statspool_bwd_kernel(const T* __restrict__ x, const float* __restrict__ dout,
                                                            T* __restrict__ dx, int N, int H, int W, int C, int mode,
                                                            int relu_mask) {
  pdl_prologue();
  long long total = (long long)N * H * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C); long long q = i / C; int h = (int)(q % H); int n = (int)(q / H);
    long long base = (((long long)n * H + h) * W) * C + c;
    if (mode == 0) {
      float g = dout[(long long)n * C * H + (long long)c * H + h] / (float)W;
      for (int w = 0; w < W; ++w)
        dx[base + (long long)w * C] = from_f<T>((relu_mask && !(to_f(x[base + (long long)w * C]) > 0.f)) ? 0.f : g);
    } else {
      const T* p = x + base;
      float s = 0.f;
      for (int w = 0; w < W; ++w) s += to_f(p[(long long)w * C]);
      float mean = s / (float)W;
      const float* o = dout + (long long)n * C * 2 * H + (long long)c * 2 * H;
      float kv = 2.f * o[h] / (float)(W - 1);
      float km = mean > 0.f ? o[H + h] * 0.5f / (sqrtf(mean) * (float)W) : 0.f;
      for (int w = 0; w < W; ++w) {
        const float xv = to_f(p[(long long)w * C]);
        dx[base + (long long)w * C] = from_f<T>((relu_mask && !(xv > 0.f)) ? 0.f : fmaf(kv, xv - mean, km));
      }
    }
  }
}
SVK_API int svk_statspool_bwd(const void* x, const float* dout, void* dx, int N, int H, int W, int C, int mode,
                              int relu_mask, int dtype, void* stream) {
  SVK_REQUIRE(x && dout && dx && N > 0 && H > 0 && W > 0 && C > 0 && (mode == 0 || mode == 1), SVK_E_BADARG, "statspool_bwd: bad args");
  long long total = (long long)N * H * C;
  long long b = (total + 255) / 256; long long cap = (long long)svk_num_sms() * 8; if (b > cap) b = cap;
  SVK_DISPATCH_DTYPE(dtype, "statspool_bwd",
    svk_launch(statspool_bwd_kernel<T>, (int)b, 256, 0, as_stream(stream), (const T*)x, dout, (T*)dx, N, H, W, C, mode, relu_mask);)
  SVK_LAUNCH_CHECK("statspool_bwd");
  return 0;
}

// ------------------------------------------------------------------------------------------ SGEMM (strided)
namespace {
constexpr int GM = 64, GN = 64, GK = 16, GT = 256;

// AK: op(A) has unit stride along k (a_sk == 1); otherwise along m.  Same for B with BK_ (b_sk == 1) vs n.
template <bool AK, bool BK_>
__global__ void __launch_bounds__(GT) sgemm_kernel(const float* __restrict__ A, long long a_sm, long long a_sk,
                                                   const float* __restrict__ B, long long b_sk, long long b_sn,
                                                   float* __restrict__ C, long long ldc, int M, int N, int K,
                                                   float alpha, float beta, const float* __restrict__ bias) {
  pdl_prologue();
  __shared__ float As[GK][GM + 4];
  __shared__ float Bs[GK][GN + 4];
  const int t = threadIdx.x;
  const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
  const int ty = t >> 4, tx = t & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += GK) {
    float av[4], bv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = t + i * GT;                       // 1024 elements per tile
      int am, ak, bn, bk;
      if (AK) { ak = e % GK; am = e / GK; } else { am = e % GM; ak = e / GM; }
      if (BK_) { bk = e % GK; bn = e / GK; } else { bn = e % GN; bk = e / GN; }
      av[i] = (m0 + am < M && k0 + ak < K) ? A[(long long)(m0 + am) * a_sm + (long long)(k0 + ak) * a_sk] : 0.f;
      bv[i] = (n0 + bn < N && k0 + bk < K) ? B[(long long)(k0 + bk) * b_sk + (long long)(n0 + bn) * b_sn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = t + i * GT;
      int am, ak, bn, bk;
      if (AK) { ak = e % GK; am = e / GK; } else { am = e % GM; ak = e / GM; }
      if (BK_) { bk = e % GK; bn = e / GK; } else { bn = e % GN; bk = e / GN; }
      As[ak][am] = av[i];
      Bs[bk][bn] = bv[i];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      float4 x = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 y = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float xa[4] = {x.x, x.y, x.z, x.w}, ya[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xa[i], ya[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = alpha * acc[i][j];
      if (bias) v += bias[n];
      float* dst = C + (long long)m * ldc + n;
      if (beta != 0.f) v = fmaf(beta, *dst, v);
      *dst = v;
    }
  }
}
}  // namespace

SVK_API int svk_sgemm(const float* A, long long a_sm, long long a_sk, const float* B, long long b_sk, long long b_sn,
                      float* C, long long ldc, int M, int N, int K, float alpha, float beta, const float* bias,
                      void* stream) {
  SVK_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0 && ldc >= N, SVK_E_BADARG, "sgemm: bad args");
  dim3 grid((N + GN - 1) / GN, (M + GM - 1) / GM);
  cudaStream_t st = as_stream(stream);
  bool ak = (a_sk == 1), bk = (b_sk == 1);
  if (ak && bk) svk_launch(sgemm_kernel<true, true>, grid, GT, 0, st, A, a_sm, a_sk, B, b_sk, b_sn, C, ldc, M, N, K, alpha, beta, bias);
  else if (ak) svk_launch(sgemm_kernel<true, false>, grid, GT, 0, st, A, a_sm, a_sk, B, b_sk, b_sn, C, ldc, M, N, K, alpha, beta, bias);
  else if (bk) svk_launch(sgemm_kernel<false, true>, grid, GT, 0, st, A, a_sm, a_sk, B, b_sk, b_sn, C, ldc, M, N, K, alpha, beta, bias);
  else svk_launch(sgemm_kernel<false, false>, grid, GT, 0, st, A, a_sm, a_sk, B, b_sk, b_sn, C, ldc, M, N, K, alpha, beta, bias);
  SVK_LAUNCH_CHECK("sgemm");
  return 0;
}

__global__ void colsum_kernel(const float* __restrict__ x, float* __restrict__ out, int M, int N, long long ld) {
  pdl_prologue();
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int m = 0; m < M; ++m) s += x[(long long)m * ld + n];
  out[n] = s;
}
SVK_API int svk_colsum(const float* x, float* out, int M, int N, long long ld, void* stream) {
  SVK_REQUIRE(x && out && M > 0 && N > 0 && ld >= N, SVK_E_BADARG, "colsum: bad args");
  svk_launch(colsum_kernel, (N + 127) / 128, 128, 0, as_stream(stream), x, out, M, N, ld);
  SVK_LAUNCH_CHECK("colsum");
  return 0;
}
