// CUDA-core implicit-GEMM convolution (fprop / dgrad / wgrad), fp32 accumulate, fp32 or bf16 storage.
// This is the fp32 VALIDATION-mode path (north-star: "<=1e-4 relative in an fp32 validation mode") and the
// on-device cross-check for the tcgen05 kernels in conv_tc.cu; it is not the product path for bf16.
#include "svk_common.cuh"
#include <stdlib.h>

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

struct GatherArgs {
  const void* in;    // fprop: x [N,Hin,Win,Kc]; dgrad: dy [N,Hin,Win,Kc]
  const void* w;     // packed [taps][Nout][Kc]
  void* out;         // [N,Hout,Wout,Nout]
  int N, Hin, Win, Kc, Hout, Wout, Nout, R, stride;
  const float* scale; const float* shift;
  const void* res; const void* res_m; const void* mask;
  int relu;
  const int* valid_w;
};

template <typename T> __device__ inline void load4(const T* p, float (&v)[4]);
template <> __device__ inline void load4<float>(const float* p, float (&v)[4]) {
  float4 r = *reinterpret_cast<const float4*>(p); v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
}
template <> __device__ inline void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
  uint2 r = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
  float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <typename T> __device__ inline void store4(T* p, const float (&v)[4]);
template <> __device__ inline void store4<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ inline void store4<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[4]) {
  uint2 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
  h[0] = __floats2bfloat162_rn(v[0], v[1]); h[1] = __floats2bfloat162_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = r;
}

// out[m, n] = sum_{tap, k} in[pix(m, tap), k] * w[tap][n][k]
template <typename T, bool DGRAD>
__global__ void __launch_bounds__(NT) conv_gather_kernel(GatherArgs a) {
  pdl_prologue();
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const T* in = (const T*)a.in;
  const T* w = (const T*)a.w;
  const int t = threadIdx.x;
  const long long Mtot = (long long)a.N * a.Hout * a.Wout;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int pad = a.R / 2;
  // this thread's load row
  const int lrow = t >> 2, lk = (t & 3) * 4;
  long long lm = m0 + lrow;
  bool lvalid = lm < Mtot;
  int ln = 0, loh = 0, low = 0;
  if (lvalid) { low = (int)(lm % a.Wout); long long q = lm / a.Wout; loh = (int)(q % a.Hout); ln = (int)(q / a.Hout); }
  const int bn = n0 + lrow;  // weight row this thread loads
  const bool bvalid = bn < a.Nout;
  const int ty = t >> 4, tx = t & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int r = 0; r < a.R; ++r) {
    for (int s = 0; s < a.R; ++s) {
      int ih, iw; bool ok = lvalid;
      if (!DGRAD) {
        ih = loh * a.stride + r - pad; iw = low * a.stride + s - pad;
      } else {
        int th = loh + pad - r, tw = low + pad - s;
        ok = ok && th >= 0 && tw >= 0 && (th % a.stride == 0) && (tw % a.stride == 0);
        ih = th / a.stride; iw = tw / a.stride;
      }
      ok = ok && ih >= 0 && ih < a.Hin && iw >= 0 && iw < a.Win;
      const T* ap = in + (((long long)ln * a.Hin + ih) * a.Win + iw) * a.Kc;
      const T* bp = w + ((long long)(r * a.R + s) * a.Nout + bn) * a.Kc;
      for (int k0 = 0; k0 < a.Kc; k0 += BK) {
        float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
        if (ok) load4<T>(ap + k0 + lk, av);
        if (bvalid) load4<T>(bp + k0 + lk, bv);
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) { As[lk + i][lrow] = av[i]; Bs[lk + i][lrow] = bv[i]; }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
          float4 x = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
          float4 y = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
          const float xa[4] = {x.x, x.y, x.z, x.w}, ya[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xa[i], ya[j], acc[i][j]);
        }
      }
    }
  }
  // epilogue
  const int n = n0 + tx * 4;
  if (n >= a.Nout) return;
  T* out = (T*)a.out;
  float sc[4] = {1.f, 1.f, 1.f, 1.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
  if (a.scale) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { sc[j] = a.scale[n + j]; sh[j] = a.shift[n + j]; }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long m = m0 + ty * 4 + i;
    if (m >= Mtot) continue;
    long long idx = m * a.Nout + n;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = fmaf(acc[i][j], sc[j], sh[j]);
    if (a.res) { float r4[4]; load4<T>((const T*)a.res + idx, r4);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] += r4[j]; }
    if (a.res_m) { float r4[4], k4[4]; load4<T>((const T*)a.res_m + idx, r4); load4<T>((const T*)a.mask + idx, k4);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] += (k4[j] > 0.f) ? r4[j] : 0.f; }
    if (a.relu) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f); }
    if (a.valid_w) {
      int ow = (int)(m % a.Wout); int nimg = (int)(m / ((long long)a.Wout * a.Hout));
      if (ow >= a.valid_w[nimg]) { v[0] = v[1] = v[2] = v[3] = 0.f; }
    }
    store4<T>(out + idx, v);
  }
}

struct WgradArgs {
  const void* x; const void* dy; float* dw;   // dw: split-K partials [gridDim.z][taps][Cout][Cin]
  int N, H, W, Cin, Ho, Wo, Cout, R, stride;
  long long kslice;  // pixels per z-slice
};

// partial[z][tap][co][ci] = sum over this z-slice's pixels of dy[pix, co] * x[shift_tap(pix), ci]
template <typename T>
__global__ void __launch_bounds__(NT) conv_wgrad_simt_kernel(WgradArgs a) {
  pdl_prologue();
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const T* x = (const T*)a.x;
  const T* dy = (const T*)a.dy;
  const int t = threadIdx.x;
  const int co0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;            // column = tap*Cin + ci
  const int ncols = a.R * a.R * a.Cin;
  const long long Ktot = (long long)a.N * a.Ho * a.Wo;
  const long long kbeg = (long long)blockIdx.z * a.kslice;
  long long kend = kbeg + a.kslice; if (kend > Ktot) kend = Ktot;
  const int pad = a.R / 2;
  const int lkk = t >> 4, l4 = (t & 15) * 4;
  const bool avalid = co0 + l4 < a.Cout;
  const int col = n0 + l4;
  const bool bvalid = col < ncols;
  int tap = 0, ci = 0, r = 0, s = 0;
  if (bvalid) { tap = col / a.Cin; ci = col % a.Cin; r = tap / a.R; s = tap % a.R; }
  const int ty = t >> 4, tx = t & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (long long k0 = kbeg; k0 < kend; k0 += BK) {
    long long k = k0 + lkk;
    float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (k < kend) {
      if (avalid) load4<T>(dy + k * a.Cout + co0 + l4, av);
      if (bvalid) {
        int ow = (int)(k % a.Wo); long long q = k / a.Wo; int oh = (int)(q % a.Ho); int n = (int)(q / a.Ho);
        int ih = oh * a.stride + r - pad, iw = ow * a.stride + s - pad;
        if (ih >= 0 && ih < a.H && iw >= 0 && iw < a.W)
          load4<T>(x + (((long long)n * a.H + ih) * a.W + iw) * a.Cin + ci, bv);
      }
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&As[lkk][l4]) = make_float4(av[0], av[1], av[2], av[3]);
    *reinterpret_cast<float4*>(&Bs[lkk][l4]) = make_float4(bv[0], bv[1], bv[2], bv[3]);
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 xa4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 ya4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float xa[4] = {xa4.x, xa4.y, xa4.z, xa4.w}, ya[4] = {ya4.x, ya4.y, ya4.z, ya4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xa[i], ya[j], acc[i][j]);
    }
  }
  const int c = n0 + tx * 4;
  if (c >= ncols) return;
  const int otap = c / a.Cin, oci = c % a.Cin;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int co = co0 + ty * 4 + i;
    if (co >= a.Cout) continue;
    float* dst = a.dw + (long long)blockIdx.z * ((long long)a.R * a.R * a.Cout * a.Cin) +
                 ((long long)otap * a.Cout + co) * a.Cin + oci;
    *reinterpret_cast<float4*>(dst) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
}

}  // namespace

static int check_desc(const char* name, const svk_conv_desc* d) {
  SVK_REQUIRE(d, SVK_E_BADARG, "%s: null desc", name);
  SVK_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, SVK_E_BADARG, "%s: non-positive dims", name);
  SVK_REQUIRE((d->R == 1 || d->R == 3) && (d->stride == 1 || d->stride == 2), SVK_E_UNSUPPORTED,
              "%s: R=%d stride=%d unsupported", name, d->R, d->stride);
  SVK_REQUIRE(d->Ho == (d->H - 1) / d->stride + 1 && d->Wo == (d->W - 1) / d->stride + 1, SVK_E_BADARG,
              "%s: Ho/Wo inconsistent with H/W/stride", name);
  SVK_REQUIRE(d->Cin % 16 == 0 && d->Cout % 16 == 0, SVK_E_UNSUPPORTED, "%s: channels must be multiples of 16", name);
  return 0;
}

int svk_conv2d_fwd_simt(const svk_conv_desc* d, const void* x, const void* w, void* y, const float* scale,
                        const float* shift, const void* residual, int relu, const int* valid_wo, cudaStream_t st) {
  GatherArgs a{x, w, y, d->N, d->H, d->W, d->Cin, d->Ho, d->Wo, d->Cout, d->R, d->stride, scale, shift, residual,
               nullptr, nullptr, relu, valid_wo};
  long long M = (long long)d->N * d->Ho * d->Wo;
  dim3 grid((unsigned)((M + BM - 1) / BM), (d->Cout + BN - 1) / BN);
  if (d->dtype == SVK_F32) svk_launch(conv_gather_kernel<float, false>, grid, NT, 0, st, a);
  else svk_launch(conv_gather_kernel<__nv_bfloat16, false>, grid, NT, 0, st, a);
  SVK_LAUNCH_CHECK("conv2d_fwd(simt)");
  return 0;
}
int svk_conv2d_dgrad_simt(const svk_conv_desc* d, const void* dy, const void* w, void* dx, const void* res,
                          const void* res_m, const void* mask, cudaStream_t st) {
  GatherArgs a{dy, w, dx, d->N, d->Ho, d->Wo, d->Cout, d->H, d->W, d->Cin, d->R, d->stride, nullptr, nullptr, res,
               res_m, mask, 0, nullptr};
  long long M = (long long)d->N * d->H * d->W;
  dim3 grid((unsigned)((M + BM - 1) / BM), (d->Cin + BN - 1) / BN);
  if (d->dtype == SVK_F32) svk_launch(conv_gather_kernel<float, true>, grid, NT, 0, st, a);
  else svk_launch(conv_gather_kernel<__nv_bfloat16, true>, grid, NT, 0, st, a);
  SVK_LAUNCH_CHECK("conv2d_dgrad(simt)");
  return 0;
}
static void simt_wgrad_plan(const svk_conv_desc* d, long long* kslice, int* gz) {
  int gx = (d->Cout + BM - 1) / BM, gy = (d->R * d->R * d->Cin + BN - 1) / BN;
  long long K = (long long)d->N * d->Ho * d->Wo;
  long long want = (long long)svk_num_sms() * 4 / ((long long)gx * gy) + 1;   // ~4 CTAs per SM in total
  long long maxsl = (K + 255) / 256;                                          // at least 256 pixels per slice
  if (want > maxsl) want = maxsl;
  if (want < 1) want = 1;
  if (want > 1024) want = 1024;
  long long ks = (K + want - 1) / want;
  ks = (ks + BK - 1) / BK * BK;
  *kslice = ks;
  *gz = (int)((K + ks - 1) / ks);
}
int svk_conv2d_wgrad_simt(const svk_conv_desc* d, const void* x, const void* dy, float* ws, size_t ws_floats,
                          int* ksplit_out, cudaStream_t st) {
  WgradArgs a{x, dy, ws, d->N, d->H, d->W, d->Cin, d->Ho, d->Wo, d->Cout, d->R, d->stride, 0};
  int gx = (d->Cout + BM - 1) / BM, gy = (d->R * d->R * d->Cin + BN - 1) / BN, gz;
  simt_wgrad_plan(d, &a.kslice, &gz);
  SVK_REQUIRE((size_t)gz * d->R * d->R * d->Cout * d->Cin <= ws_floats, SVK_E_BADARG, "conv2d_wgrad(simt): workspace too small");
  dim3 grid(gx, gy, gz);
  if (d->dtype == SVK_F32) svk_launch(conv_wgrad_simt_kernel<float>, grid, NT, 0, st, a);
  else svk_launch(conv_wgrad_simt_kernel<__nv_bfloat16>, grid, NT, 0, st, a);
  SVK_LAUNCH_CHECK("conv2d_wgrad(simt)");
  *ksplit_out = gz;
  return 0;
}

// dw_oihw[co][ci][tap] = sum_k partial[k][tap][co][ci]: split-K reduction fused with the packed -> OIHW transpose.
// A block reduces 32 consecutive outputs at a time with 8 k-lanes of 32 threads (coalesced 128-byte rows of each
// partial), then folds the k-lanes through shared memory: with one thread per output the small early-stage filters
// (9,216 outputs, 148 partials) were a 15 us chain of dependent loads on 36 blocks.
__global__ void __launch_bounds__(256) wgrad_reduce_klane_kernel(const float* __restrict__ ws, int ksplit, long long stride,
                                                                 float* __restrict__ dw, int Cout, int Cin, int taps) {
  pdl_prologue();
  __shared__ float part[8][33];
  const int kl = threadIdx.x >> 5, li = threadIdx.x & 31;
  for (long long base = (long long)blockIdx.x * 32; base < stride; base += (long long)gridDim.x * 32) {
    const long long i = base + li;
    float s = 0.f;
    if (i < stride)
      for (int k = kl; k < ksplit; k += 8) s += ws[(long long)k * stride + i];
    part[kl][li] = s;
    __syncthreads();
    if (kl == 0 && i < stride) {
      float t = part[0][li];
#pragma unroll
      for (int j = 1; j < 8; ++j) t += part[j][li];
      int ci = (int)(i % Cin); long long r = i / Cin; int co = (int)(r % Cout); int tp = (int)(r / Cout);
      dw[((long long)co * Cin + ci) * taps + tp] = t;
    }
    __syncthreads();
  }
}

// Four consecutive outputs per thread (one float4 per partial, 8 partials in flight): the shape for many outputs and few
// partials (late stages: 0.6-2.4 M outputs, 12-48 partials = 28 MB that are still L2-resident).  Fixed summation order.
__global__ void __launch_bounds__(256) wgrad_reduce_flat_kernel(const float* __restrict__ ws, int ksplit, long long stride,
                                                                float* __restrict__ dw, int Cout, int Cin, int taps) {
  pdl_prologue();
  const long long nv = stride >> 2;                   // Cin % 4 == 0 on this path
  for (long long iv = (long long)blockIdx.x * blockDim.x + threadIdx.x; iv < nv; iv += (long long)gridDim.x * blockDim.x) {
    const float4* src = reinterpret_cast<const float4*>(ws) + iv;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    int k = 0;
    for (; k + 8 <= ksplit; k += 8) {
      float4 t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) t[u] = src[(long long)(k + u) * nv];
#pragma unroll
      for (int u = 0; u < 8; ++u) { s.x += t[u].x; s.y += t[u].y; s.z += t[u].z; s.w += t[u].w; }
    }
    for (; k < ksplit; ++k) { const float4 t = src[(long long)k * nv]; s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w; }
    const long long i = iv << 2;
    const int ci = (int)(i % Cin); const long long r = i / Cin; const int co = (int)(r % Cout); const int t = (int)(r / Cout);
    float* o = dw + ((long long)co * Cin + ci) * taps + t;
    o[0] = s.x; o[taps] = s.y; o[2 * taps] = s.z; o[3 * taps] = s.w;
  }
}

// implemented in conv_tc.cu
int svk_conv2d_fwd_tc(const svk_conv_desc* d, const void* x, const void* w, void* y, double* stats, const float* scale,
                      const float* shift, const void* residual, int relu, const int* valid_wo, cudaStream_t st);
int svk_conv2d_dgrad_tc(const svk_conv_desc* d, const void* dy, const void* w, void* dx, const void* res,
                        const void* res_m, const void* mask, const svk_bn_bwd_fuse* bn, cudaStream_t st);
int svk_downsample_dgrad_tc(const svk_conv_desc* d1, const void* dy1, const void* w1, const svk_conv_desc* dd,
                            const void* dyd, const void* wd, void* dx, const svk_bn_bwd_fuse* bn, cudaStream_t st);
int svk_conv2d_wgrad_tc(const svk_conv_desc* d, const void* x, const void* dy, float* ws, size_t ws_floats, int* ksplit_out,
                        cudaStream_t st);
size_t svk_conv2d_wgrad_tc_ws_floats(const svk_conv_desc* d);
// implemented in conv_tc_wgrad9.cu (3x3/s1, Cin == Cout in {32, 64}: all nine taps in one UMMA through shifted operand views)
bool svk_wgrad9_applicable(const svk_conv_desc* d);
size_t svk_conv2d_wgrad9_tc_ws_floats(const svk_conv_desc* d);
int svk_conv2d_wgrad9_tc(const svk_conv_desc* d, const void* x, const void* dy, float* ws, size_t ws_floats, int* ksplit_out,
                         cudaStream_t st);
// implemented in conv_tc_wgradr.cu (3x3/s1, channels multiples of 128: full-row tiles, filter columns stacked along N)
bool svk_wgradr_applicable(const svk_conv_desc* d);
size_t svk_conv2d_wgradr_tc_ws_floats(const svk_conv_desc* d);
int svk_conv2d_wgradr_tc(const svk_conv_desc* d, const void* x, const void* dy, float* ws, size_t ws_floats, int* ksplit_out,
                         cudaStream_t st);
static bool wgradr_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("SVK_DISABLE_WGRADR"); v = (e && e[0] == '1') ? 0 : 1; }
  return v == 1;
}
static bool wgrad9_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("SVK_DISABLE_WGRAD9"); v = (e && e[0] == '1') ? 0 : 1; }
  return v == 1;
}

SVK_API int svk_conv2d_fwd(const svk_conv_desc* d, const void* x, const void* w, void* y, double* stats,
                           const float* scale, const float* shift, const void* residual, int relu,
                           const int* valid_wo, void* stream) {
  if (int e = check_desc("conv2d_fwd", d)) return e;
  SVK_REQUIRE(x && w && y, SVK_E_BADARG, "conv2d_fwd: null pointer");
  SVK_REQUIRE((scale == nullptr) == (shift == nullptr), SVK_E_BADARG, "conv2d_fwd: scale and shift go together");
  SVK_REQUIRE(aligned16(x) && aligned16(w) && aligned16(y) && (!residual || aligned16(residual)), SVK_E_ALIGN,
              "conv2d_fwd: pointers must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  if (d->impl == SVK_IMPL_TCGEN05) {
    SVK_REQUIRE(d->dtype == SVK_BF16, SVK_E_UNSUPPORTED, "conv2d_fwd: tcgen05 path is bf16 only");
    return svk_conv2d_fwd_tc(d, x, w, y, stats, scale, shift, residual, relu, valid_wo, st);
  }
  SVK_REQUIRE(d->impl == SVK_IMPL_SIMT, SVK_E_BADARG, "conv2d_fwd: bad impl %d", d->impl);
  if (int e = svk_conv2d_fwd_simt(d, x, w, y, scale, shift, residual, relu, valid_wo, st)) return e;
  if (stats) return svk_channel_stats(y, (long long)d->N * d->Ho * d->Wo, d->Cout, d->dtype, stats, stream);
  return 0;
}

// finish the contract of the fused epilogue with the stand-alone kernels (validation path)
static int bn_fuse_compose(const svk_bn_bwd_fuse* bn, void* dx, long long M, int C, int dtype, void* stream) {
  if (!bn) return 0;
  if (int e = svk_relu_mask_inplace(dx, bn->mask, M * C, dtype, stream)) return e;
  if (bn->c)
    return svk_bn_bwd_reduce(dx, nullptr, bn->c, bn->mean, bn->rstd, nullptr, nullptr, nullptr, bn->sums, M, C, dtype, stream);
  return 0;
}

SVK_API int svk_conv2d_dgrad_bn(const svk_conv_desc* d, const void* dy, const void* w, void* dx, const void* res,
                                const void* res_m, const void* mask, const svk_bn_bwd_fuse* bn, void* stream) {
  if (int e = check_desc("conv2d_dgrad", d)) return e;
  SVK_REQUIRE(dy && w && dx, SVK_E_BADARG, "conv2d_dgrad: null pointer");
  SVK_REQUIRE((res_m == nullptr) == (mask == nullptr), SVK_E_BADARG, "conv2d_dgrad: res_m and mask go together");
  if (bn) {
    SVK_REQUIRE(bn->mask, SVK_E_BADARG, "conv2d_dgrad_bn: mask is required");
    SVK_REQUIRE(!bn->c || (bn->mean && bn->rstd && bn->sums), SVK_E_BADARG, "conv2d_dgrad_bn: c needs mean, rstd and sums");
    SVK_REQUIRE(!(d->R == 1 && d->stride == 2), SVK_E_UNSUPPORTED, "conv2d_dgrad_bn: not available for the 1x1/s2 accumulate form");
    SVK_REQUIRE(d->Cin <= 512, SVK_E_UNSUPPORTED, "conv2d_dgrad_bn: Cin=%d > 512", d->Cin);
    SVK_REQUIRE(res_m == nullptr, SVK_E_UNSUPPORTED,
                "conv2d_dgrad_bn: a masked residual cannot be combined with the BatchNorm fusion (the epilogue streams at "
                "most three tensors); pass an already-masked gradient as res");
  }
  cudaStream_t st = as_stream(stream);
  if (d->impl == SVK_IMPL_TCGEN05) {
    SVK_REQUIRE(d->dtype == SVK_BF16, SVK_E_UNSUPPORTED, "conv2d_dgrad: tcgen05 path is bf16 only");
    return svk_conv2d_dgrad_tc(d, dy, w, dx, res, res_m, mask, bn, st);
  }
  SVK_REQUIRE(d->impl == SVK_IMPL_SIMT, SVK_E_BADARG, "conv2d_dgrad: bad impl %d", d->impl);
  if (int e = svk_conv2d_dgrad_simt(d, dy, w, dx, res, res_m, mask, st)) return e;
  return bn_fuse_compose(bn, dx, (long long)d->N * d->H * d->W, d->Cin, d->dtype, stream);
}
SVK_API int svk_conv2d_dgrad(const svk_conv_desc* d, const void* dy, const void* w, void* dx, const void* res,
                             const void* res_m, const void* mask, void* stream) {
  return svk_conv2d_dgrad_bn(d, dy, w, dx, res, res_m, mask, nullptr, stream);
}

SVK_API int svk_downsample_dgrad_bn(const svk_conv_desc* d1, const void* dy1, const void* w1_dgrad, const svk_conv_desc* dd,
                                    const void* dyd, const void* wd_dgrad, void* dx, const svk_bn_bwd_fuse* bn, void* stream) {
  if (int e = check_desc("downsample_dgrad", d1)) return e;
  if (int e = check_desc("downsample_dgrad", dd)) return e;
  SVK_REQUIRE(dy1 && w1_dgrad && dyd && wd_dgrad && dx, SVK_E_BADARG, "downsample_dgrad: null pointer");
  SVK_REQUIRE(d1->R == 3 && d1->stride == 2 && dd->R == 1 && dd->stride == 2, SVK_E_BADARG,
              "downsample_dgrad: expects a 3x3/s2 conv and a 1x1/s2 shortcut conv");
  SVK_REQUIRE(d1->N == dd->N && d1->H == dd->H && d1->W == dd->W && d1->Cin == dd->Cin && d1->dtype == dd->dtype &&
              d1->impl == dd->impl, SVK_E_BADARG, "downsample_dgrad: the two convs must share their input");
  if (bn) {
    SVK_REQUIRE(bn->mask, SVK_E_BADARG, "downsample_dgrad: mask is required");
    SVK_REQUIRE(!bn->c || (bn->mean && bn->rstd && bn->sums), SVK_E_BADARG, "downsample_dgrad: c needs mean, rstd and sums");
    SVK_REQUIRE(d1->Cin <= 512, SVK_E_UNSUPPORTED, "downsample_dgrad: Cin=%d > 512", d1->Cin);
  }
  cudaStream_t st = as_stream(stream);
  if (d1->impl == SVK_IMPL_TCGEN05) {
    SVK_REQUIRE(d1->dtype == SVK_BF16, SVK_E_UNSUPPORTED, "downsample_dgrad: tcgen05 path is bf16 only");
    return svk_downsample_dgrad_tc(d1, dy1, w1_dgrad, dd, dyd, wd_dgrad, dx, bn, st);
  }
  SVK_REQUIRE(d1->impl == SVK_IMPL_SIMT, SVK_E_BADARG, "downsample_dgrad: bad impl %d", d1->impl);
  if (int e = svk_conv2d_dgrad_simt(dd, dyd, wd_dgrad, dx, nullptr, nullptr, nullptr, st)) return e;
  if (int e = svk_conv2d_dgrad_simt(d1, dy1, w1_dgrad, dx, dx, nullptr, nullptr, st)) return e;   // res aliases dx
  return bn_fuse_compose(bn, dx, (long long)d1->N * d1->H * d1->W, d1->Cin, d1->dtype, stream);
}

SVK_API size_t svk_conv2d_wgrad_workspace_bytes(const svk_conv_desc* d) {
  if (check_desc("conv2d_wgrad_workspace_bytes", d)) return 0;
  if (d->impl == SVK_IMPL_TCGEN05) {
    if (wgrad9_enabled() && svk_wgrad9_applicable(d)) return svk_conv2d_wgrad9_tc_ws_floats(d) * sizeof(float);
    if (wgradr_enabled() && svk_wgradr_applicable(d)) return svk_conv2d_wgradr_tc_ws_floats(d) * sizeof(float);
    return svk_conv2d_wgrad_tc_ws_floats(d) * sizeof(float);
  }
  long long ks; int gz;
  simt_wgrad_plan(d, &ks, &gz);
  return (size_t)gz * d->R * d->R * d->Cout * d->Cin * sizeof(float);
}

SVK_API int svk_conv2d_wgrad(const svk_conv_desc* d, const void* x, const void* dy, float* dw_oihw, void* workspace,
                             size_t workspace_bytes, void* stream) {
  if (int e = check_desc("conv2d_wgrad", d)) return e;
  SVK_REQUIRE(x && dy && dw_oihw && workspace, SVK_E_BADARG, "conv2d_wgrad: null pointer");
  SVK_REQUIRE(aligned16(workspace), SVK_E_ALIGN, "conv2d_wgrad: workspace must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  int ksplit = 0;
  if (d->impl == SVK_IMPL_TCGEN05) {
    SVK_REQUIRE(d->dtype == SVK_BF16, SVK_E_UNSUPPORTED, "conv2d_wgrad: tcgen05 path is bf16 only");
    if (wgrad9_enabled() && svk_wgrad9_applicable(d)) {
      if (int e = svk_conv2d_wgrad9_tc(d, x, dy, (float*)workspace, workspace_bytes / sizeof(float), &ksplit, st)) return e;
    } else if (wgradr_enabled() && svk_wgradr_applicable(d)) {
      if (int e = svk_conv2d_wgradr_tc(d, x, dy, (float*)workspace, workspace_bytes / sizeof(float), &ksplit, st)) return e;
    } else {
      if (int e = svk_conv2d_wgrad_tc(d, x, dy, (float*)workspace, workspace_bytes / sizeof(float), &ksplit, st)) return e;
    }
  } else {
    SVK_REQUIRE(d->impl == SVK_IMPL_SIMT, SVK_E_BADARG, "conv2d_wgrad: bad impl %d", d->impl);
    if (int e = svk_conv2d_wgrad_simt(d, x, dy, (float*)workspace, workspace_bytes / sizeof(float), &ksplit, st)) return e;
  }
  long long stride = (long long)d->R * d->R * d->Cout * d->Cin;
  const long long cap = (long long)svk_num_sms() * 8;
  if (ksplit >= 32 && stride <= 65536) {       // few outputs, many partials (stages 1-2): spread the k loop over 8 lanes
    long long b = (stride + 31) / 32; if (b > cap) b = cap;
    svk_launch(wgrad_reduce_klane_kernel, (int)b, 256, 0, st, (const float*)workspace, ksplit, stride, dw_oihw, d->Cout, d->Cin, d->R * d->R);
  } else {
    SVK_REQUIRE(d->Cin % 4 == 0 && stride % 4 == 0, SVK_E_UNSUPPORTED, "conv2d_wgrad: Cin=%d must be a multiple of 4", d->Cin);
    long long b = (stride / 4 + 255) / 256; if (b > cap) b = cap;
    svk_launch(wgrad_reduce_flat_kernel, (int)b, 256, 0, st, (const float*)workspace, ksplit, stride, dw_oihw, d->Cout, d->Cin, d->R * d->R);
  }
  SVK_LAUNCH_CHECK("conv2d_wgrad(reduce)");
  return 0;
}

// ------------------------------------------------------------------------------------------------ stem (Cin = 1)
// HBM-bound by construction: 4 B in, 2 * Cout B out per pixel (131 MB at batch 256).  Work item = (image row, run of 4
// consecutive pixels, group of V channels = one 16-byte piece of each output row).  A thread keeps its 9 x V filter taps in
// REGISTERS for the whole kernel (the first version re-read all 288 taps from shared memory for every pixel: LDS-bound,
// 28 % of the HBM roof), fetches the 3 x 6 input window of its run with 18 independent loads issued up front (a one-pixel
// item with 9 dependent loads was latency-bound at 16 warps per SM), and consecutive threads write consecutive 16-byte
// pieces.  Training forward: the per-channel sum / sum of squares of the values AS STORED are accumulated on the way (the
// separate svk_channel_stats pass re-read the whole tensor).
constexpr int STEM_RUN = 4;
template <typename T, bool AFFINE, bool STATS>
__global__ void __launch_bounds__(256, 2) stem_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          T* __restrict__ y, int N, int H, int W, int CO,
                                                          const float* __restrict__ scale, const float* __restrict__ shift,
                                                          int relu, const int* __restrict__ valid_w, double* __restrict__ stats) {
  pdl_prologue();
  constexpr int V = Vec<T>::N;
  const int groups = CO / V;
  const int cg = threadIdx.x % groups, lane_p = threadIdx.x / groups, lanes = blockDim.x / groups;
  float wr[9][V], sc[V], sh[V];
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int j = 0; j < V; ++j) wr[k][j] = w[(cg * V + j) * 9 + k];
#pragma unroll
  for (int j = 0; j < V; ++j) { sc[j] = AFFINE ? scale[cg * V + j] : 1.f; sh[j] = AFFINE ? shift[cg * V + j] : 0.f; }
  float s1[V], s2[V];
#pragma unroll
  for (int j = 0; j < V; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  const int runs_w = (W + STEM_RUN - 1) / STEM_RUN;
  const int items = N * H * runs_w, istep = (int)gridDim.x * lanes;
  for (int it = blockIdx.x * lanes + lane_p; it < items; it += istep) {
    const int row = it / runs_w, w0 = (it - row * runs_w) * STEM_RUN;
    const int n = row / H, h = row - n * H;
    const int vw = valid_w ? valid_w[n] : W;
    const float* xr = x + (long long)row * W;
    float v[3][STEM_RUN + 2];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int ih = h + r - 1;
      const bool rok = ih >= 0 && ih < H;
#pragma unroll
      for (int t = 0; t < STEM_RUN + 2; ++t) {
        const int iw = w0 + t - 1;
        v[r][t] = (rok && iw >= 0 && iw < W) ? xr[(r - 1) * W + iw] : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < STEM_RUN; ++u) {
      if (w0 + u >= W) break;
      float o[V];
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float a = 0.f;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int t = 0; t < 3; ++t) a = fmaf(v[r][u + t], wr[r * 3 + t][j], a);
        if (AFFINE) a = fmaf(a, sc[j], sh[j]);
        a = relu ? fmaxf(a, 0.f) : a;
        o[j] = (w0 + u >= vw) ? 0.f : a;
      }
      Vec<T>::store(y + ((long long)row * W + w0 + u) * CO + cg * V, o);
      if (STATS) {
#pragma unroll
        for (int j = 0; j < V; ++j) { const float r_ = round_to<T>(o[j]); s1[j] += r_; s2[j] = fmaf(r_, r_, s2[j]); }
      }
    }
  }
  if (STATS) {
    // lanes of a warp that share a channel group are `groups` apart (groups = 4 or 8 divides 32): butterfly over them first
#pragma unroll
    for (int j = 0; j < V; ++j)
      for (int o = groups; o < 32; o <<= 1) { s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o); s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o); }
    // fixed-order block reduction (no fp32 atomics: BatchNorm statistics that differ in the last bit between two runs flip
    // ReLU ties downstream and make a small-batch step irreproducible), then one fp64 atomic per channel per block
    __shared__ float red[8][2 * 64];
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) < groups) {
#pragma unroll
      for (int j = 0; j < V; ++j) { red[warp][cg * V + j] = s1[j]; red[warp][CO + cg * V + j] = s2[j]; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * CO; i += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int w_ = 0; w_ < 8; ++w_) t += red[w_][i];
      atomicAdd(&stats[i], (double)t);
    }
  }
}
SVK_API int svk_stem_conv_fwd(const float* x, const float* w, void* y, int N, int H, int W, int Cout, int dtype,
                              const float* scale, const float* shift, int relu, const int* valid_w, double* stats,
                              void* stream) {
  SVK_REQUIRE(x && w && y && N > 0 && H > 0 && W > 0, SVK_E_BADARG, "stem_conv_fwd: bad args");
  SVK_REQUIRE(Cout == 32 || Cout == 64, SVK_E_UNSUPPORTED, "stem_conv_fwd: Cout must be 32 or 64, got %d", Cout);
  SVK_REQUIRE((scale == nullptr) == (shift == nullptr), SVK_E_BADARG, "stem_conv_fwd: scale and shift go together");
  SVK_REQUIRE((long long)N * H * W < (1ll << 31), SVK_E_UNSUPPORTED, "stem_conv_fwd: more than 2^31 pixels");
  const int groups = Cout / (dtype == SVK_BF16 ? 8 : 4);
  SVK_REQUIRE(groups == 4 || groups == 8 || groups == 16, SVK_E_UNSUPPORTED, "stem_conv_fwd: %d channel groups", groups);
  const int lanes = 256 / groups;
  long long items = (long long)N * H * ((W + STEM_RUN - 1) / STEM_RUN);
  long long b = (items + lanes * 2 - 1) / (lanes * 2); long long cap = (long long)svk_num_sms() * 8; if (b > cap) b = cap; if (b < 1) b = 1;
  cudaStream_t st = as_stream(stream);
  SVK_DISPATCH_DTYPE(dtype, "stem_conv_fwd",
    if (scale && stats) svk_launch(stem_fwd_kernel<T, true, true>, (int)b, 256, 0, st, x, w, (T*)y, N, H, W, Cout, scale, shift, relu, valid_w, stats);
    else if (scale) svk_launch(stem_fwd_kernel<T, true, false>, (int)b, 256, 0, st, x, w, (T*)y, N, H, W, Cout, scale, shift, relu, valid_w, stats);
    else if (stats) svk_launch(stem_fwd_kernel<T, false, true>, (int)b, 256, 0, st, x, w, (T*)y, N, H, W, Cout, scale, shift, relu, valid_w, stats);
    else svk_launch(stem_fwd_kernel<T, false, false>, (int)b, 256, 0, st, x, w, (T*)y, N, H, W, Cout, scale, shift, relu, valid_w, stats);)
  SVK_LAUNCH_CHECK("stem_conv_fwd");
  return 0;
}

// dw[co][tap] = sum_p dy[p, co] * x[p + shift(tap)].  Same work items as the forward kernel: the 4 dy pieces and the 3 x 6
// input window of a run are 22 independent loads issued up front (HBM-bound on dy, 2 * Cout bytes per pixel: what matters is
// bytes in flight — with one 16-byte load in flight per thread the kernel ran at 14 % of the HBM roof), 9 * V accumulators
// in registers.  Reduction: butterfly over the lanes of a warp that share a channel group, shared-memory atomics per warp,
// one fp32 atomic per (channel, tap) per block.
template <typename T>
__global__ void __launch_bounds__(256, 2) stem_wgrad_kernel(const float* __restrict__ x, const T* __restrict__ dy,
                                                            float* __restrict__ dw, int N, int H, int W, int CO) {
  pdl_prologue();
  constexpr int V = Vec<T>::N;
  typedef typename Vec<T>::raw Raw;
  const int groups = CO / V;                                  // channel groups per pixel (4 or 8 for bf16)
  const int cg = threadIdx.x % groups, lane_p = threadIdx.x / groups, lanes = blockDim.x / groups;
  float acc[9][V];
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int i = 0; i < V; ++i) acc[k][i] = 0.f;
  const int runs_w = (W + STEM_RUN - 1) / STEM_RUN;
  const int items = N * H * runs_w, istep = (int)gridDim.x * lanes;
  for (int it = blockIdx.x * lanes + lane_p; it < items; it += istep) {
    const int row = it / runs_w, w0 = (it - row * runs_w) * STEM_RUN;
    const int h = row % H;
    Raw raw[STEM_RUN];
#pragma unroll
    for (int u = 0; u < STEM_RUN; ++u)
      if (w0 + u < W) raw[u] = *reinterpret_cast<const Raw*>(dy + ((long long)row * W + w0 + u) * CO + cg * V);
    const float* xr = x + (long long)row * W;
    float v[3][STEM_RUN + 2];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int ih = h + r - 1;
      const bool rok = ih >= 0 && ih < H;
#pragma unroll
      for (int t = 0; t < STEM_RUN + 2; ++t) {
        const int iw = w0 + t - 1;
        v[r][t] = (rok && iw >= 0 && iw < W) ? xr[(r - 1) * W + iw] : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < STEM_RUN; ++u) {
      if (w0 + u >= W) break;
      float g[V];
      Vec<T>::unpack(raw[u], g);
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int t = 0; t < 3; ++t)
#pragma unroll
          for (int i = 0; i < V; ++i) acc[r * 3 + t][i] = fmaf(g[i], v[r][u + t], acc[r * 3 + t][i]);
    }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int i = 0; i < V; ++i)
      for (int o = groups; o < 32; o <<= 1) acc[k][i] += __shfl_xor_sync(0xffffffffu, acc[k][i], o);
  __shared__ float red[64 * 9];                               // CO <= 64
  for (int i = threadIdx.x; i < CO * 9; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  if ((threadIdx.x & 31) < groups) {
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
      for (int i = 0; i < V; ++i) atomicAdd(&red[(cg * V + i) * 9 + k], acc[k][i]);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < CO * 9; o += blockDim.x) atomicAdd(&dw[o], red[o]);
}
SVK_API int svk_stem_conv_wgrad(const float* x, const void* dy, float* dw, int N, int H, int W, int Cout, int dtype,
                                void* stream) {
  SVK_REQUIRE(x && dy && dw && N > 0 && H > 0 && W > 0, SVK_E_BADARG, "stem_conv_wgrad: bad args");
  SVK_REQUIRE(Cout == 32 || Cout == 64, SVK_E_UNSUPPORTED, "stem_conv_wgrad: Cout must be 32 or 64, got %d", Cout);
  SVK_REQUIRE((long long)N * H * W < (1ll << 31), SVK_E_UNSUPPORTED, "stem_conv_wgrad: more than 2^31 pixels");
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * Cout * 9, st);
  SVK_REQUIRE(e == cudaSuccess, (int)e, "stem_conv_wgrad: memset failed: %s", cudaGetErrorString(e));
  const int groups = Cout / (dtype == SVK_BF16 ? 8 : 4);
  SVK_REQUIRE(groups == 4 || groups == 8 || groups == 16, SVK_E_UNSUPPORTED, "stem_conv_wgrad: %d channel groups", groups);
  const int lanes = 256 / groups;
  long long items = (long long)N * H * ((W + STEM_RUN - 1) / STEM_RUN);
  long long b = (items + lanes * 8 - 1) / (lanes * 8); long long cap = (long long)svk_num_sms() * 2; if (b > cap) b = cap; if (b < 1) b = 1;
  SVK_DISPATCH_DTYPE(dtype, "stem_conv_wgrad",
    svk_launch(stem_wgrad_kernel<T>, (int)b, 256, 0, st, x, (const T*)dy, dw, N, H, W, Cout);)
  SVK_LAUNCH_CHECK("stem_conv_wgrad");
  return 0;
}
