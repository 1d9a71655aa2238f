// Fused AAM-softmax head: row normalisation, cosine GEMM, additive angular margin, scale, cross-entropy (log-sum-exp, loss,
// target rank) and the whole backward, in FOUR launches (reference: AAMLayer.forward model.py:483-501, nn.CrossEntropyLoss
// train_resnet.py:201/317, accuracy.py:4-16; before: l2norm x4, gemm x6, margin x2, ce x2, split-K reduces = ~17 launches
// that materialised x_hat, W_hat, d_logits, d_x_hat and d_W_hat in HBM).
//
//   aam_ce_fwd_kernel     one CTA per 64 classes: W rows normalised on the way into smem, the batch streamed in 64-row
//                         chunks (x rows normalised on the way in), cos = x_hat . W_hat^T on the tensor cores
//                         (mma.sync tf32; "exact" = 3xTF32 split for the fp32 validation mode), margin on the target
//                         column, x s, logits stored once, per-(row, 32-class slab) max / sum-exp partials
//   aam_ce_finish_kernel  one block per row: partials -> log-sum-exp, loss row, mean loss, rank of the target class
//   aam_ce_bwd_kernel     one CTA per 64 classes: d_cos = s (softmax - onehot) dloss / B x margin derivative RECOMPUTED from
//                         logits and lse (d_logits never exists in memory); d_W_hat = d_cos^T x_hat accumulated over the
//                         batch and pushed through the W-normalisation Jacobian in the epilogue -> d_W; the partial
//                         d_x_hat = d_cos W_hat of these 64 classes -> workspace (deterministic, no atomics)
//   aam_ce_bwd_finish_kernel  one block per row: sum of the class-tile partials, x-normalisation Jacobian -> d_h
//
// The head is 2.4 GFLOP and ~20 MB per step — latency-bound, not throughput-bound — so these are warp-level mma.sync
// kernels (no TMEM / TMA set-up cost per launch); what matters is the launch count and the bytes that never reach HBM.
// E (embedding width) is fixed at 256 (model.py:355-357 hard-codes it).
#include "svk_common.cuh"

namespace {

constexpr int AF_E = 256;           // embedding width
constexpr int AF_T = 64;            // rows per chunk = classes per CTA
constexpr int AF_THREADS = 256;
constexpr int AF_LDF = AF_E + 4;    // smem row pitch, forward  (A / B^T fragments: bank = 4 g + t)
constexpr int AF_LDB = AF_E + 8;    // smem row pitch, backward (B fragments read k-major: bank = 8 t + g)
constexpr int AF_LDS = AF_T + 4;    // pitch of the 64 x 64 d_cos tiles

__device__ __forceinline__ uint32_t f2tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// hi / lo split of an fp32 value into two tf32 numbers (3xTF32: a b ~= a_lo b_hi + a_hi b_lo + a_hi b_hi)
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = f2tf32(x);
  lo = f2tf32(x - __uint_as_float(hi));
}

// acc[MT][NT][4] += A[m0 .. m0+16 MT) x K] * B[K x n0 .. n0+8 NT)] for one warp.
//   A(m, k) = As[m * lda + k];   B(k, n) = BKN ? Bs[k * ldb + n] : Bs[n * ldb + k]
template <bool EXACT, bool BKN, int MT, int NT>
__device__ __forceinline__ void warp_gemm(float (&acc)[MT][NT][4], const float* __restrict__ As, int lda, int m0,
                                          const float* __restrict__ Bs, int ldb, int n0, int K, int lane) {
  const int g = lane >> 2, t = lane & 3;
  for (int k0 = 0; k0 < K; k0 += 8) {
    uint32_t ah[MT][4], al[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const float* a = As + (m0 + mt * 16 + g) * lda + k0 + t;
      const float v[4] = {a[0], a[8 * lda], a[4], a[8 * lda + 4]};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (EXACT) split_tf32(v[i], ah[mt][i], al[mt][i]);
        else ah[mt][i] = f2tf32(v[i]);
      }
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int n = n0 + nt * 8 + g;
      float v0, v1;
      if (BKN) { v0 = Bs[(k0 + t) * ldb + n]; v1 = Bs[(k0 + t + 4) * ldb + n]; }
      else { v0 = Bs[n * ldb + k0 + t]; v1 = Bs[n * ldb + k0 + t + 4]; }
      uint32_t bh[2], bl[2];
      if (EXACT) { split_tf32(v0, bh[0], bl[0]); split_tf32(v1, bh[1], bl[1]); }
      else { bh[0] = f2tf32(v0); bh[1] = f2tf32(v1); }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        if (EXACT) { mma_tf32(acc[mt][nt], al[mt], bh); mma_tf32(acc[mt][nt], ah[mt], bl); }
        mma_tf32(acc[mt][nt], ah[mt], bh);
      }
    }
  }
}

// 64 rows of `src` (row pitch AF_E) starting at row0 -> L2-normalised rows in smem (pitch ld); rows >= nrows are zero.
// inv_s (smem, 64 floats) and inv_g (global, optional) receive 1 / max(||row||, eps).  All 256 threads call it.
__device__ __forceinline__ void load_norm_rows(const float* __restrict__ src, long long row0, long long nrows, float* dst, int ld,
                                               float* inv_s, float* inv_g, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < AF_T; r += AF_THREADS / 32) {
    const long long row = row0 + r;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (row < nrows) {
      const float4* p = reinterpret_cast<const float4*>(src + row * AF_E);
      a = p[lane]; b = p[32 + lane];
    }
    float s = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
    s = warp_sum(s);
    const float inv = 1.f / fmaxf(sqrtf(s), eps);           // F.normalize: x / max(||x||, eps)
    if (lane == 0) { inv_s[r] = inv; if (inv_g && row < nrows) inv_g[row] = inv; }
    float4* d = reinterpret_cast<float4*>(dst + r * ld);
    d[lane] = make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
    d[32 + lane] = make_float4(b.x * inv, b.y * inv, b.z * inv, b.w * inv);
  }
}

struct AamP {
  const float* h; const float* W; const long long* y;
  float* logits; float* cos_t; float* xinv; float* winv;
  float* pmax; float* psum;            // [2 * n_tiles][B]
  int B, C, n_tiles;
  float cos_m, sin_m, th, mm, s;
};

template <bool EXACT>
__global__ void __launch_bounds__(AF_THREADS, 1) aam_ce_fwd_kernel(const AamP p) {
  pdl_prologue();
  extern __shared__ float sm[];
  float* ws = sm;                              // [64][AF_LDF]  W_hat tile
  float* xs = ws + AF_T * AF_LDF;              // [64][AF_LDF]  x_hat chunk
  float* winv_s = xs + AF_T * AF_LDF;          // [64]
  float* xinv_s = winv_s + AF_T;               // [64]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int tile = blockIdx.x, c0 = tile * AF_T;
  if (tile == 0) {
    for (int b = threadIdx.x; b < p.B; b += AF_THREADS) {
      const long long l = p.y[b];
      if (l < 0 || l >= p.C) { printf("svk aam_ce_fwd: label %lld of row %d is outside [0, %d)\n", l, b, p.C); __trap(); }
    }
  }
  load_norm_rows(p.W, c0, p.C, ws, AF_LDF, winv_s, p.winv, 1e-12f);
  const int wm = warp >> 1, wn = warp & 1;     // 4 x 2 warps: 16 rows x 32 classes each
  for (int m0 = 0; m0 < p.B; m0 += AF_T) {
    __syncthreads();                           // previous chunk's fragments are consumed (and ws is complete)
    load_norm_rows(p.h, m0, p.B, xs, AF_LDF, xinv_s, tile == 0 ? p.xinv : nullptr, 1e-12f);
    __syncthreads();
    float acc[1][4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) { acc[0][nt][0] = acc[0][nt][1] = acc[0][nt][2] = acc[0][nt][3] = 0.f; }
    warp_gemm<EXACT, false, 1, 4>(acc, xs, AF_LDF, wm * 16, ws, AF_LDF, wn * 32, AF_E, lane);
    // epilogue: margin on the target column, scale, store, log-sum-exp partials of this warp's 32 classes
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int row = m0 + wm * 16 + g + half * 8;
      const bool rvalid = row < p.B;
      const long long lab = rvalid ? p.y[row] : -1;
      float z[8];
      float mx = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int cls = c0 + wn * 32 + nt * 8 + 2 * t + j;
          float v = acc[0][nt][half * 2 + j];
          const bool valid = rvalid && cls < p.C;
          if (valid && (long long)cls == lab) {
            p.cos_t[row] = v;
            const float sine = sqrtf(fminf(fmaxf(1.f - v * v, 0.f), 1.f));
            const float phi = v * p.cos_m - sine * p.sin_m;
            v = (v - p.th > 0.f) ? phi : v - p.mm;
          }
          v *= p.s;
          if (valid) p.logits[(long long)row * p.C + cls] = v;
          z[nt * 2 + j] = valid ? v : -INFINITY;
          mx = fmaxf(mx, z[nt * 2 + j]);
        }
      }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      float se = 0.f;
      if (mx > -INFINITY) {
#pragma unroll
        for (int i = 0; i < 8; ++i) se += (z[i] > -INFINITY) ? expf(z[i] - mx) : 0.f;
      }
      se += __shfl_xor_sync(0xffffffffu, se, 1);
      se += __shfl_xor_sync(0xffffffffu, se, 2);
      if (t == 0 && rvalid) {
        const long long o = (long long)(tile * 2 + wn) * p.B + row;
        p.pmax[o] = mx; p.psum[o] = se;
      }
    }
  }
}

__device__ __forceinline__ float af_block_reduce(float v, bool is_max, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : (is_max ? -INFINITY : 0.f);
  r = is_max ? warp_max(r) : warp_sum(r);
  return r;
}

__global__ void __launch_bounds__(128) aam_ce_finish_kernel(const float* __restrict__ logits, const long long* __restrict__ y,
                                                            const float* __restrict__ pmax, const float* __restrict__ psum,
                                                            int np, int B, int C, float* __restrict__ lse,
                                                            float* __restrict__ loss_rows, int* __restrict__ rank,
                                                            float* __restrict__ loss_mean) {
  pdl_prologue();
  __shared__ float red[32];
  const int r = blockIdx.x;
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < np; j += blockDim.x) mx = fmaxf(mx, pmax[(long long)j * B + r]);
  mx = af_block_reduce(mx, true, red);
  float s = 0.f;
  for (int j = threadIdx.x; j < np; j += blockDim.x) {
    const float m = pmax[(long long)j * B + r];
    if (m > -INFINITY) s += psum[(long long)j * B + r] * expf(m - mx);
  }
  s = af_block_reduce(s, false, red);
  const float* z = logits + (long long)r * C;
  const float zt = z[y[r]];
  float cnt = 0.f;
  if (rank) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) cnt += (z[c] > zt) ? 1.f : 0.f;
    cnt = af_block_reduce(cnt, false, red);
  }
  if (threadIdx.x == 0) {
    const float l = mx + logf(s);
    lse[r] = l;
    if (loss_rows) loss_rows[r] = l - zt;
    if (rank) rank[r] = (int)(cnt + 0.5f);
    if (loss_mean) atomicAdd(loss_mean, (l - zt) / (float)B);
  }
}

struct AamBwdP {
  const float* h; const float* W; const long long* y;
  const float* logits; const float* lse; const float* cos_t; const float* gout;
  float* dW; float* part;              // part: [n_tiles][B][E] partial d_x_hat
  int B, C, n_tiles;
  float cos_m, sin_m, th, s;
};

template <bool EXACT>
__global__ void __launch_bounds__(AF_THREADS, 1) aam_ce_bwd_kernel(const AamBwdP p) {
  pdl_prologue();
  extern __shared__ float sm[];
  float* ws = sm;                              // [64][AF_LDB]  W_hat tile (B operand of d_x_hat, k = class)
  float* xs = ws + AF_T * AF_LDB;              // [64][AF_LDB]  x_hat chunk (B operand of d_W_hat, k = row); later the d_W_hat tile
  float* ds_rc = xs + AF_T * AF_LDB;           // [64 rows][AF_LDS]   d_cos, row-major   (A of d_x_hat)
  float* ds_cr = ds_rc + AF_T * AF_LDS;        // [64 classes][AF_LDS] d_cos, transposed (A of d_W_hat)
  float* winv_s = ds_cr + AF_T * AF_LDS;       // [64]
  float* xinv_s = winv_s + AF_T;               // [64]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int tile = blockIdx.x, c0 = tile * AF_T;
  load_norm_rows(p.W, c0, p.C, ws, AF_LDB, winv_s, nullptr, 1e-12f);
  const float gscale = (p.gout ? *p.gout : 1.f) / (float)p.B;
  const int wm = warp >> 2, wn = warp & 3;     // 2 x 4 warps: 32 x 64 each
  float accw[2][8][4];                         // d_W_hat[64 classes][256], accumulated over the batch
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { accw[mt][nt][0] = accw[mt][nt][1] = accw[mt][nt][2] = accw[mt][nt][3] = 0.f; }
  for (int m0 = 0; m0 < p.B; m0 += AF_T) {
    __syncthreads();
    load_norm_rows(p.h, m0, p.B, xs, AF_LDB, xinv_s, nullptr, 1e-12f);
    // d_cos of this (64 rows x 64 classes) tile, recomputed from the logits and the row log-sum-exp
    for (int i = threadIdx.x; i < AF_T * AF_T; i += AF_THREADS) {
      const int rr = i >> 6, cc = i & 63;
      const int row = m0 + rr, cls = c0 + cc;
      float d = 0.f;
      if (row < p.B && cls < p.C) {
        const float z = p.logits[(long long)row * p.C + cls];
        const bool tgt = (long long)cls == p.y[row];
        d = (expf(z - p.lse[row]) - (tgt ? 1.f : 0.f)) * gscale * p.s;
        if (tgt) {
          const float ct = p.cos_t[row];
          if (ct - p.th > 0.f) {
            const float u = 1.f - ct * ct;
            const float dsine = (u > 0.f && u < 1.f) ? -ct / sqrtf(u) : 0.f;      // d sqrt(clamp(1 - c^2, 0, 1)) / dc
            d *= p.cos_m - p.sin_m * dsine;
          }
        }
      }
      ds_rc[rr * AF_LDS + cc] = d;
      ds_cr[cc * AF_LDS + rr] = d;
    }
    __syncthreads();
    {   // partial d_x_hat[64 rows][256] = d_cos[64 x 64] * W_hat[64 x 256]
      float acc[2][8][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) { acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f; }
      warp_gemm<EXACT, true, 2, 8>(acc, ds_rc, AF_LDS, wm * 32, ws, AF_LDB, wn * 64, AF_T, lane);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int row = m0 + wm * 32 + mt * 16 + g + half * 8;
          if (row < p.B) {
            float* dst = p.part + ((long long)tile * p.B + row) * AF_E + wn * 64 + 2 * t;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
              *reinterpret_cast<float2*>(dst + nt * 8) = make_float2(acc[mt][nt][half * 2], acc[mt][nt][half * 2 + 1]);
          }
        }
    }
    // d_W_hat[64 classes][256] += d_cos^T[64 x 64] * x_hat[64 x 256]
    warp_gemm<EXACT, true, 2, 8>(accw, ds_cr, AF_LDS, wm * 32, xs, AF_LDB, wn * 64, AF_T, lane);
  }
  // epilogue: d_W = (d_W_hat - W_hat (W_hat . d_W_hat)) / ||W||, rows of this class tile
  __syncthreads();
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float* dst = xs + (wm * 32 + mt * 16 + g + half * 8) * AF_LDB + wn * 64 + 2 * t;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
        *reinterpret_cast<float2*>(dst + nt * 8) = make_float2(accw[mt][nt][half * 2], accw[mt][nt][half * 2 + 1]);
    }
  __syncthreads();
  for (int r = warp; r < AF_T; r += AF_THREADS / 32) {
    const int cls = c0 + r;
    if (cls >= p.C) break;
    const float4* gq = reinterpret_cast<const float4*>(xs + r * AF_LDB);
    const float4* wq = reinterpret_cast<const float4*>(ws + r * AF_LDB);
    const float4 g0 = gq[lane], g1 = gq[32 + lane], w0 = wq[lane], w1 = wq[32 + lane];
    float d = g0.x * w0.x + g0.y * w0.y + g0.z * w0.z + g0.w * w0.w + g1.x * w1.x + g1.y * w1.y + g1.z * w1.z + g1.w * w1.w;
    d = warp_sum(d);
    const float inv = winv_s[r];
    float4* o = reinterpret_cast<float4*>(p.dW + (long long)cls * AF_E);
    o[lane] = make_float4((g0.x - w0.x * d) * inv, (g0.y - w0.y * d) * inv, (g0.z - w0.z * d) * inv, (g0.w - w0.w * d) * inv);
    o[32 + lane] = make_float4((g1.x - w1.x * d) * inv, (g1.y - w1.y * d) * inv, (g1.z - w1.z * d) * inv, (g1.w - w1.w * d) * inv);
  }
}

__global__ void __launch_bounds__(AF_E) aam_ce_bwd_finish_kernel(const float* __restrict__ h, const float* __restrict__ part,
                                                                 int n_tiles, int B, float* __restrict__ dh) {
  pdl_prologue();
  __shared__ float red[32];
  const int r = blockIdx.x, e = threadIdx.x;
  float v = 0.f;
  for (int tl = 0; tl < n_tiles; ++tl) v += part[((long long)tl * B + r) * AF_E + e];      // fixed order: deterministic
  const float x = h[(long long)r * AF_E + e];
  const float ss = af_block_reduce(x * x, false, red);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  const float xh = x * inv;
  const float dot = af_block_reduce(xh * v, false, red);
  dh[(long long)r * AF_E + e] = (v - xh * dot) * inv;
}

constexpr size_t AF_FWD_SMEM = (size_t)(2 * AF_T * AF_LDF + 2 * AF_T) * sizeof(float);
constexpr size_t AF_BWD_SMEM = (size_t)(2 * AF_T * AF_LDB + 2 * AF_T * AF_LDS + 2 * AF_T) * sizeof(float);

inline int af_tiles(int C) { return (C + AF_T - 1) / AF_T; }
inline size_t af_align(size_t n) { return (n + 255) / 256 * 256; }

}  // namespace

SVK_API size_t svk_aam_ce_workspace_bytes(int B, int E, int C) {
  if (B <= 0 || C <= 0 || E != AF_E) return 0;
  const size_t nt = (size_t)af_tiles(C);
  return af_align(2 * nt * B * sizeof(float)) * 2 + af_align(nt * (size_t)B * AF_E * sizeof(float));
}

SVK_API int svk_aam_ce_fwd(const float* h, const float* W, const long long* y, float* logits, float* cos_t, float* lse,
                           float* loss_rows, int* rank, float* loss_mean, int B, int E, int C, float cos_m, float sin_m,
                           float th, float mm, float s, int exact, void* workspace, size_t workspace_bytes, void* stream) {
  SVK_REQUIRE(h && W && y && logits && cos_t && lse && workspace && B > 0 && C > 0, SVK_E_BADARG, "aam_ce_fwd: bad args");
  SVK_REQUIRE(E == AF_E, SVK_E_UNSUPPORTED, "aam_ce_fwd: embedding width %d (only %d is built, model.py:355)", E, AF_E);
  SVK_REQUIRE(workspace_bytes >= svk_aam_ce_workspace_bytes(B, E, C), SVK_E_BADARG, "aam_ce_fwd: workspace too small");
  SVK_REQUIRE(aligned16(h) && aligned16(W) && aligned16(workspace), SVK_E_ALIGN, "aam_ce_fwd: pointers must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  const int nt = af_tiles(C);
  AamP p{};
  p.h = h; p.W = W; p.y = y; p.logits = logits; p.cos_t = cos_t; p.xinv = nullptr; p.winv = nullptr;
  p.pmax = reinterpret_cast<float*>(workspace);
  p.psum = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + af_align(2 * (size_t)nt * B * sizeof(float)));
  p.B = B; p.C = C; p.n_tiles = nt; p.cos_m = cos_m; p.sin_m = sin_m; p.th = th; p.mm = mm; p.s = s;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(aam_ce_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AF_FWD_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(aam_ce_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AF_FWD_SMEM);
    SVK_REQUIRE(e == cudaSuccess, (int)e, "aam_ce_fwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    configured = true;
  }
  if (loss_mean) {
    cudaError_t e = cudaMemsetAsync(loss_mean, 0, sizeof(float), st);
    SVK_REQUIRE(e == cudaSuccess, (int)e, "aam_ce_fwd: memset failed: %s", cudaGetErrorString(e));
  }
  if (exact) svk_launch(aam_ce_fwd_kernel<true>, nt, AF_THREADS, AF_FWD_SMEM, st, p);
  else svk_launch(aam_ce_fwd_kernel<false>, nt, AF_THREADS, AF_FWD_SMEM, st, p);
  SVK_LAUNCH_CHECK("aam_ce_fwd");
  svk_launch(aam_ce_finish_kernel, B, 128, 0, st, (const float*)logits, y, (const float*)p.pmax, (const float*)p.psum, 2 * nt, B, C,
             lse, loss_rows, rank, loss_mean);
  SVK_LAUNCH_CHECK("aam_ce_finish");
  return 0;
}

SVK_API int svk_aam_ce_bwd(const float* h, const float* W, const long long* y, const float* logits, const float* lse,
                           const float* cos_t, const float* gout, float* dh, float* dW, int B, int E, int C, float cos_m,
                           float sin_m, float th, float s, int exact, void* workspace, size_t workspace_bytes, void* stream) {
  SVK_REQUIRE(h && W && y && logits && lse && cos_t && dh && dW && workspace && B > 0 && C > 0, SVK_E_BADARG, "aam_ce_bwd: bad args");
  SVK_REQUIRE(E == AF_E, SVK_E_UNSUPPORTED, "aam_ce_bwd: embedding width %d (only %d is built)", E, AF_E);
  SVK_REQUIRE(workspace_bytes >= svk_aam_ce_workspace_bytes(B, E, C), SVK_E_BADARG, "aam_ce_bwd: workspace too small");
  SVK_REQUIRE(aligned16(h) && aligned16(W) && aligned16(dh) && aligned16(dW) && aligned16(workspace), SVK_E_ALIGN,
              "aam_ce_bwd: pointers must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  const int nt = af_tiles(C);
  AamBwdP p{};
  p.h = h; p.W = W; p.y = y; p.logits = logits; p.lse = lse; p.cos_t = cos_t; p.gout = gout; p.dW = dW;
  p.part = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 2 * af_align(2 * (size_t)nt * B * sizeof(float)));
  p.B = B; p.C = C; p.n_tiles = nt; p.cos_m = cos_m; p.sin_m = sin_m; p.th = th; p.s = s;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(aam_ce_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AF_BWD_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(aam_ce_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AF_BWD_SMEM);
    SVK_REQUIRE(e == cudaSuccess, (int)e, "aam_ce_bwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    configured = true;
  }
  if (exact) svk_launch(aam_ce_bwd_kernel<true>, nt, AF_THREADS, AF_BWD_SMEM, st, p);
  else svk_launch(aam_ce_bwd_kernel<false>, nt, AF_THREADS, AF_BWD_SMEM, st, p);
  SVK_LAUNCH_CHECK("aam_ce_bwd");
  svk_launch(aam_ce_bwd_finish_kernel, B, AF_E, 0, st, h, (const float*)p.part, nt, B, dh);
  SVK_LAUNCH_CHECK("aam_ce_bwd_finish");
  return 0;
}
