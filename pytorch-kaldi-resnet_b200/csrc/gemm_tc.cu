// fp32-in / fp32-out GEMM on the tensor cores (tcgen05.mma kind::tf32, fp32 accumulation in TMEM):
//     C[M, N] = A[M, K] * B[N, K]^T (+ bias[N])
// used by the product (bf16) mode for the embedding FC, the AAM cosine logits (x_hat * W_hat^T over the speaker-class
// matrix) and their gradients, and for the cohort score matrix of adaptive s-norm.  Each operand may be stored
// "K-major" (rows of K contiguous: one 128B-swizzled TMA box per tile) or "MN-major" (rows of M/N contiguous — the
// transposed views the backward GEMMs need: 32-element atoms, LBO = atom stride), so no transposed copies are made.
// Tile 128 x 128 x 32, 5-stage TMA ring, one accumulator; optional split-K writes partials that gemm_reduce_kernel sums
// in a fixed order (deterministic).  The fp32 validation mode keeps the CUDA-core svk_sgemm.
#include "tc_common.cuh"

namespace {

constexpr int GBM = 128, GBN = 128, GBK = 32;          // GBK fp32 = one 128-byte swizzle row
constexpr int G_STAGE = (GBM + GBN) * GBK * 4;         // 32 KB
constexpr int G_STAGES = 5;
constexpr int G_SMEM = G_STAGES * G_STAGE + 1024 + 1024;

struct GemmP {
  int M, N, K;
  int m_tiles, n_tiles, ksplit, ksteps_per;            // ksteps_per = K steps (of GBK) per split
  float* C; long long ldc;                             // ksplit == 1: final output; else partials [ksplit][M][ldc]
  long long part_stride;
  const float* bias;
};

__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (elect_one()) asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// tf32 operands: a_format = b_format = 2; c = f32
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <bool AK, bool BK_>   // operand stored K-major?
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ GemmP p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t aux = base + G_STAGES * G_STAGE;
  const uint32_t bar_full = aux, bar_empty = aux + 64, bar_done = aux + 128;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(gbase + G_STAGES * G_STAGE + 160);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < G_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "n"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_wait();             // the prologue above overlapped the previous kernel's tail; global memory is touched from here on
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  int wi = blockIdx.x;
  const int ks = wi % p.ksplit; wi /= p.ksplit;
  const int nt = wi % p.n_tiles;
  const int mt = wi / p.n_tiles;
  const int m0 = mt * GBM, n0 = nt * GBN;
  const int total_ksteps = (p.K + GBK - 1) / GBK;
  const int k_beg = ks * p.ksteps_per;
  const int k_end = (k_beg + p.ksteps_per) < total_ksteps ? (k_beg + p.ksteps_per) : total_ksteps;

  if (warp == 0) {
    int stage = 0; uint32_t ph = 0;
    for (int kk = k_beg; kk < k_end; ++kk) {
      const int k0 = kk * GBK;
      mbar_wait(bar_empty + 8 * stage, ph ^ 1u);
      const uint32_t sa = base + stage * G_STAGE;
      const uint32_t sb = sa + GBM * GBK * 4;
      mbar_expect_tx(bar_full + 8 * stage, G_STAGE);
      if (AK) tma_load_2d(sa, &tmA, bar_full + 8 * stage, k0, m0);                       // box {32 k, 128 m}
      else
#pragma unroll
        for (int a = 0; a < 4; ++a) tma_load_2d(sa + a * 4096, &tmA, bar_full + 8 * stage, m0 + a * 32, k0);   // box {32 m, 32 k}
      if (BK_) tma_load_2d(sb, &tmB, bar_full + 8 * stage, k0, n0);
      else
#pragma unroll
        for (int a = 0; a < 4; ++a) tma_load_2d(sb + a * 4096, &tmB, bar_full + 8 * stage, n0 + a * 32, k0);
      if (++stage == G_STAGES) { stage = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_tf32(GBM, GBN, AK ? 0 : 1, BK_ ? 0 : 1);
    int stage = 0; uint32_t ph = 0;
    for (int kk = k_beg; kk < k_end; ++kk) {
      mbar_wait(bar_full + 8 * stage, ph);
      tc_fence_after();
      const uint32_t sa = base + stage * G_STAGE;
      const uint32_t sb = sa + GBM * GBK * 4;
      // K-major: SWIZZLE_128B, rows of 128 B, 8-row groups 1024 B apart, UMMA K = 8 fp32 = 32 B inside the swizzle row.
      // MN-major: 32-bit operands can only be transposed from the SWIZZLE_128B_BASE32B layout (layout type 1, TMA
      // "128B_ATOM_32B"): 32-element atoms 4096 B apart (LBO), 4 k-rows per 512 B swizzle group (SBO), UMMA K = 8 rows = 1024 B.
      const uint64_t ad0 = AK ? make_desc(sa, 16, 1024, 2) : make_desc(sa, 4096, 512, 1);
      const uint64_t bd0 = BK_ ? make_desc(sb, 16, 1024, 2) : make_desc(sb, 4096, 512, 1);
#pragma unroll
      for (int k = 0; k < GBK / 8; ++k)
        tc_mma_tf32(tmem_base, ad0 + (uint64_t)(AK ? 2 * k : 64 * k), bd0 + (uint64_t)(BK_ ? 2 * k : 64 * k), idesc,
                    (kk != k_beg || k != 0) ? 1u : 0u);
      tc_commit(bar_empty + 8 * stage);
      if (++stage == G_STAGES) { stage = 0; ph ^= 1u; }
    }
    tc_commit(bar_done);
  } else {
    mbar_wait(bar_done, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    float* crow = p.C + (long long)ks * p.part_stride + (long long)m * p.ldc;
    const bool add_bias = p.bias != nullptr && p.ksplit == 1;
#pragma unroll
    for (int c = 0; c < GBN / 32; ++c) {
      uint32_t v[32];
      tc_ld32(taddr + c * 32, v);
      if (m < p.M) {
        const int nb = n0 + c * 32;
        if (nb + 32 <= p.N && ((p.ldc & 3) == 0)) {
          float4* dst = reinterpret_cast<float4*>(crow + nb);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float4 o = make_float4(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]), __uint_as_float(v[4 * e + 2]),
                                   __uint_as_float(v[4 * e + 3]));
            if (add_bias) { o.x += p.bias[nb + 4 * e]; o.y += p.bias[nb + 4 * e + 1]; o.z += p.bias[nb + 4 * e + 2]; o.w += p.bias[nb + 4 * e + 3]; }
            dst[e] = o;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (nb + e < p.N) crow[nb + e] = __uint_as_float(v[e]) + (add_bias ? p.bias[nb + e] : 0.f);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(128) : "memory");
  }
}

__global__ void __launch_bounds__(256) gemm_reduce_kernel(const float* __restrict__ part, int ksplit, long long stride,
                                                          float* __restrict__ C, long long ldc, long long ldp, int M, int N,
                                                          const float* __restrict__ bias) {
  pdl_prologue();
  long long total = (long long)M * N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int n = (int)(i % N); long long m = i / N;
    float s = bias ? bias[n] : 0.f;
    for (int k = 0; k < ksplit; ++k) s += part[(long long)k * stride + m * ldp + n];
    C[m * ldc + n] = s;
  }
}

// 2-D fp32 tensor map: dims {cols (contiguous), rows}, row pitch ld floats (must be a multiple of 4), 128B swizzle.
int make_f32_map(CUtensorMap* m, const float* ptr, long long rows, long long cols, long long ld, int box_cols, int box_rows,
                 bool atom32) {
  EncodeTiledFn enc = get_encode();
  SVK_REQUIRE(enc, SVK_E_DRIVER, "gemm_tf32: cuTensorMapEncodeTiled entry point not available");
  SVK_REQUIRE((ld % 4) == 0 && aligned16(ptr), SVK_E_ALIGN, "gemm_tf32: operand pitch (%lld floats) must be a multiple of 4 and "
              "the pointer 16-byte aligned", ld);
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SVK_REQUIRE(r == CUDA_SUCCESS, SVK_E_DRIVER, "gemm_tf32: cuTensorMapEncodeTiled failed with %d", (int)r);
  return 0;
}

template <bool AK, bool BK_>
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmP& p, int grid, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tf32_kernel<AK, BK_>, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM);
    SVK_REQUIRE(e == cudaSuccess, (int)e, "gemm_tf32: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    configured = true;
  }
  svk_launch(gemm_tf32_kernel<AK, BK_>, grid, TC_THREADS, G_SMEM, st, ta, tb, p);
  SVK_LAUNCH_CHECK("gemm_tf32");
  return 0;
}

void plan_split(int M, int N, int K, int* ksplit, int* ksteps_per) {
  const int tiles = ((M + GBM - 1) / GBM) * ((N + GBN - 1) / GBN);
  const int ksteps = (K + GBK - 1) / GBK;
  int ks = 1;
  if (tiles * 2 <= svk_num_sms() && ksteps >= 16) {
    ks = svk_num_sms() / tiles;
    if (ks > ksteps / 8) ks = ksteps / 8;       // at least 8 K steps per split
    if (ks < 1) ks = 1;
  }
  int per = (ksteps + ks - 1) / ks;
  *ksteps_per = per;
  *ksplit = (ksteps + per - 1) / per;
}

}  // namespace

SVK_API size_t svk_gemm_tf32_workspace_bytes(int M, int N, int K) {
  int ks, per;
  plan_split(M, N, K, &ks, &per);
  return ks > 1 ? (size_t)ks * M * ((N + 3) / 4 * 4) * sizeof(float) : 0;
}

SVK_API int svk_gemm_tf32(const float* A, long long lda, int a_kmajor, const float* B, long long ldb, int b_kmajor, float* C,
                          long long ldc, int M, int N, int K, const float* bias, void* workspace, size_t workspace_bytes,
                          void* stream) {
  SVK_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0 && ldc >= N, SVK_E_BADARG, "gemm_tf32: bad args");
  cudaStream_t st = as_stream(stream);
  GemmP p{};
  p.M = M; p.N = N; p.K = K;
  p.m_tiles = (M + GBM - 1) / GBM; p.n_tiles = (N + GBN - 1) / GBN;
  plan_split(M, N, K, &p.ksplit, &p.ksteps_per);
  const long long ldp = (N + 3) / 4 * 4;
  if (p.ksplit > 1) {
    SVK_REQUIRE(workspace && workspace_bytes >= (size_t)p.ksplit * M * ldp * sizeof(float) && aligned16(workspace), SVK_E_BADARG,
                "gemm_tf32: split-K workspace too small or misaligned");
    p.C = (float*)workspace; p.ldc = ldp; p.part_stride = (long long)M * ldp; p.bias = nullptr;
  } else {
    p.C = C; p.ldc = ldc; p.part_stride = 0; p.bias = bias;
  }
  CUtensorMap ta, tb;
  if (a_kmajor) { if (int e = make_f32_map(&ta, A, M, K, lda, GBK, GBM, false)) return e; }      // A[M][K]
  else { if (int e = make_f32_map(&ta, A, K, M, lda, 32, GBK, true)) return e; }                // A stored [K][M]
  if (b_kmajor) { if (int e = make_f32_map(&tb, B, N, K, ldb, GBK, GBN, false)) return e; }      // B[N][K]
  else { if (int e = make_f32_map(&tb, B, K, N, ldb, 32, GBK, true)) return e; }                // B stored [K][N]
  const int grid = p.m_tiles * p.n_tiles * p.ksplit;
  int rc;
  if (a_kmajor && b_kmajor) rc = launch_gemm<true, true>(ta, tb, p, grid, st);
  else if (a_kmajor) rc = launch_gemm<true, false>(ta, tb, p, grid, st);
  else if (b_kmajor) rc = launch_gemm<false, true>(ta, tb, p, grid, st);
  else rc = launch_gemm<false, false>(ta, tb, p, grid, st);
  if (rc) return rc;
  if (p.ksplit > 1) {
    long long total = (long long)M * N;
    long long b = (total + 255) / 256; long long cap = (long long)svk_num_sms() * 8; if (b > cap) b = cap;
    svk_launch(gemm_reduce_kernel, (int)b, 256, 0, st, (const float*)workspace, p.ksplit, p.part_stride, C, ldc, ldp, M, N, bias);
    SVK_LAUNCH_CHECK("gemm_tf32(reduce)");
  }
  return 0;
}
