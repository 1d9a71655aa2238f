// Micro-benchmark: issue rate of tcgen05.mma kind::f16 (bf16 in, fp32 accumulate in TMEM), SS mode, K-major SWIZZLE_128B
// operands, no TMA and no epilogue: how many SM cycles does one M x N x 16 instruction occupy the tensor pipe for
//   cta_group::1, M = 128, N in {32, 64, 128, 256}
//   cta_group::2, M = 256 (128 rows per CTA, each CTA holds N/2 rows of B), N in {64, 128, 256}
// and, optionally, while the other warps of the CTA hammer shared memory (the epilogue's transposes).
// Results are checked: A = 1, B[n][*] = (n % 7) + 1  =>  D[m][n] = 16 * (#UMMAs) * ((n % 7) + 1).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o bench_umma bench_umma.cu && ./bench_umma
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) return false;
  }
  return true;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn = 0, int b_mn = 0) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
template <int CG>
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (elect_one()) {
    if (CG == 1)
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    else
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  if (elect_one()) {
    if (CG == 1)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    else
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                   ::"r"(bar), "h"((uint16_t)3) : "memory");
  }
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

struct Result { unsigned long long cycles; unsigned long long ns; int bad; int timeout; };

// smem: nst stages of [A: 128 rows x 128 B][B: (N/CG) rows x 128 B], every stage 1024-aligned.
template <int N, int CG>
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(int iters, int nst, int hammer, Result* res) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  constexpr int BROWS = N / CG;
  constexpr int A_BYTES = 128 * 128, B_BYTES = BROWS * 128, STAGE = A_BYTES + B_BYTES;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  __shared__ __align__(8) uint64_t bar_done;
  __shared__ uint32_t tmem_ptr;
  __shared__ volatile int stop_flag;
  __shared__ float hammer_buf[4 * 32 * 36];

  // operands: A = 1.0, B row n (global n = rank * BROWS + row) = (n % 7) + 1
  for (int s = 0; s < nst; ++s) {
    __nv_bfloat16* a = reinterpret_cast<__nv_bfloat16*>(gbase + s * STAGE);
    for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) a[i] = __float2bfloat16(1.0f);
    __nv_bfloat16* b = reinterpret_cast<__nv_bfloat16*>(gbase + s * STAGE + A_BYTES);
    for (int i = threadIdx.x; i < BROWS * 64; i += blockDim.x) b[i] = __float2bfloat16((float)(((int)rank * BROWS + i / 64) % 7 + 1));
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar_done)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    stop_flag = 0;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> visible to the MMA (async proxy)
  if (warp == 0) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "n"(256) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "n"(256) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_ptr;

  bool ok = true;
  long long c0 = 0, c1 = 0;
  unsigned long long t0 = 0, t1 = 0;
  if (warp == 0) {
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc(128 * CG, N);
      const uint64_t a_desc0 = make_desc(base, 16, 1024, 2);
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      c0 = clock64();
      uint32_t accum = 0;
      for (int it = 0; it < iters; ++it) {
        uint64_t a_desc = a_desc0;
        for (int s = 0; s < nst; ++s) {
          const uint64_t b_desc = a_desc + (uint64_t)(A_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) { tc_mma<CG>(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, accum); accum = 1; }
          a_desc += (uint64_t)(STAGE >> 4);
        }
      }
      tc_commit<CG>(smem_u32(&bar_done));
    }
    ok = mbar_wait(smem_u32(&bar_done), 0);
    c1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    stop_flag = 1;
  } else if (hammer) {
    // the other three warps: 16-byte row writes + 4-byte column reads of a 36-word-pitch scratch, like the epilogue's
    // statistics transposes, until the MMA warp is done
    float* my = hammer_buf + warp * 32 * 36;
    const int lane = threadIdx.x & 31;
    float acc = 0.f;
    while (!stop_flag) {
#pragma unroll
      for (int g = 0; g < 8; ++g) *reinterpret_cast<float4*>(my + lane * 36 + g * 4) = make_float4(acc, 1.f, 2.f, 3.f);
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 32; ++r) acc += my[r * 36 + lane];
      __syncwarp();
    }
    if (acc == 12345.678f) res[0].bad = -1;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // check: every warp reads its 32 lanes, columns [0, N)
  int bad = 0;
  const float expect_unit = 16.0f * (float)iters * (float)nst * 4.0f;
  for (int c0col = 0; c0col < N; c0col += 8) {
    uint32_t r[8];
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0col;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int e = 0; e < 8; ++e) {
      const float want = expect_unit * (float)((c0col + e) % 7 + 1);
      if (__uint_as_float(r[e]) != want) ++bad;
    }
  }
  if (bad) atomicAdd(&res[blockIdx.x].bad, bad);
  if (threadIdx.x == 0) { res[blockIdx.x].cycles = (unsigned long long)(c1 - c0); res[blockIdx.x].ns = t1 - t0; res[blockIdx.x].timeout = ok ? 0 : 1; }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 0) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
  }
}

// Rate only (no result check): operand majors as in the wgrad kernels.  MN-major SW128 operands: atoms of 64 channels
// (128-byte rows = K index), atom stride ATOM_B bytes (A) or B_LBO bytes (B: 128 = views one pixel apart, as wgrad9/wgradr).
template <int N, int A_MN, int B_MN>
__global__ void __launch_bounds__(128, 1) umma_rate_mn_kernel(int iters, int b_lbo, Result* res) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5;
  __shared__ __align__(8) uint64_t bar_done;
  __shared__ uint32_t tmem_ptr;
  // A region: 32 KB at 0 ; B region: 64 KB at 32 KB.  Fill with small finite values.
  for (int i = threadIdx.x; i < 98304 / 2; i += blockDim.x) reinterpret_cast<__nv_bfloat16*>(gbase)[i] = __float2bfloat16(0.5f);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar_done)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "n"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_ptr;
  bool ok = true;
  long long c0 = 0, c1 = 0;
  unsigned long long t0 = 0, t1 = 0;
  if (warp == 0) {
    constexpr uint32_t idesc = make_idesc(128, N, A_MN, B_MN);
    // K-major: rows = M/N index, 128-byte rows, K advance = 32 B (4 steps per tile).
    // MN-major: rows = K index (64 rows per tile), K advance = 16 rows = 2 KB (4 steps per tile).
    const uint64_t a0 = A_MN ? make_desc(base, 8192, 1024, 2) : make_desc(base, 16, 1024, 2);
    const uint64_t b0 = B_MN ? make_desc(base + 32768, (uint32_t)b_lbo, 1024, 2) : make_desc(base + 32768, 16, 1024, 2);
    const uint64_t a_step = A_MN ? 128 : 2, b_step = B_MN ? 128 : 2;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    c0 = clock64();
    uint32_t accum = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int s = 0; s < 4; ++s) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { tc_mma<1>(tmem_base, a0 + a_step * k, b0 + b_step * k, idesc, accum); accum = 1; }
      }
    }
    tc_commit<1>(smem_u32(&bar_done));
    ok = mbar_wait(smem_u32(&bar_done), 0);
    c1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  }
  __syncthreads();
  if (threadIdx.x == 0) { res[blockIdx.x].cycles = (unsigned long long)(c1 - c0); res[blockIdx.x].ns = t1 - t0; res[blockIdx.x].timeout = ok ? 0 : 1; }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
}

// K-major SWIZZLE_64B operands (32-channel layers: 64-byte rows): 2 K steps per tile row.
template <int N>
__global__ void __launch_bounds__(128, 1) umma_rate_sw64_kernel(int iters, Result* res) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5;
  __shared__ __align__(8) uint64_t bar_done;
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < 98304 / 2; i += blockDim.x) reinterpret_cast<__nv_bfloat16*>(gbase)[i] = __float2bfloat16(0.5f);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar_done)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "n"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_ptr;
  bool ok = true;
  long long c0 = 0, c1 = 0;
  if (warp == 0) {
    constexpr uint32_t idesc = make_idesc(128, N);
    c0 = clock64();
    uint32_t accum = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int s = 0; s < 8; ++s) {          // 8 "stages" of 128 rows x 64 B (8 KB each), B behind at +64 KB
        const uint64_t a0 = make_desc(base + s * 8192, 16, 512, 4);
        const uint64_t b0 = make_desc(base + 65536 + s * 4096, 16, 512, 4);
#pragma unroll
        for (int k = 0; k < 2; ++k) { tc_mma<1>(tmem_base, a0 + 2 * k, b0 + 2 * k, idesc, accum); accum = 1; }
      }
    }
    tc_commit<1>(smem_u32(&bar_done));
    ok = mbar_wait(smem_u32(&bar_done), 0);
    c1 = clock64();
  }
  __syncthreads();
  if (threadIdx.x == 0) { res[blockIdx.x].cycles = (unsigned long long)(c1 - c0); res[blockIdx.x].timeout = ok ? 0 : 1; }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
}
template <int N>
void run_sw64(int grid, int iters) {
  const int smem = 98304 + 1024;
  CK(cudaFuncSetAttribute(umma_rate_sw64_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  Result* d; CK(cudaMalloc(&d, sizeof(Result) * grid)); CK(cudaMemset(d, 0, sizeof(Result) * grid));
  for (int rep = 0; rep < 2; ++rep) { umma_rate_sw64_kernel<N><<<grid, 128, smem>>>(iters, d); CK(cudaDeviceSynchronize()); }
  std::vector<Result> h(grid); CK(cudaMemcpy(h.data(), d, sizeof(Result) * grid, cudaMemcpyDeviceToHost));
  double cyc = 0; int to = 0;
  for (int i = 0; i < grid; ++i) { cyc += h[i].cycles; to += h[i].timeout; }
  cyc /= grid;
  const double n_umma = (double)iters * 16;
  printf("cta_group::1 M=128 N=%3d K-major SWIZZLE_64B (64-byte rows) : %7.1f cycles/UMMA  %6.1f MAC/clk/SM  timeout=%d\n", N,
         cyc / n_umma, n_umma * 128.0 * N * 16.0 / cyc, to);
  fflush(stdout);
  CK(cudaFree(d));
}


// SWIZZLE_64B operands as the stage-1 kernel uses them, with the traffic that shares the shared-memory port in the real
// kernel: warp 1 streams bulk copies global -> shared (two 8 KB copies in flight, back to back) when bg & 1, warps 2-3 do the
// epilogue-like row writes / column reads when bg & 2.  A_SHIFT: the A descriptor of UMMA pair s starts (s % 3) * 8 rows into
// the stage (tap offsets of a halo tile).  Reports cycles per UMMA and the bytes the background copies delivered per cycle.
template <int N, int commit_every = 0, int tile_len = 0, int WAIT = 0>
__global__ void __launch_bounds__(128, 1) umma_rate_sw64_bg_kernel(int iters, int bg, int a_shift,
                                                                    const uint8_t* __restrict__ src, Result* res) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ __align__(8) uint64_t bar_done;
  __shared__ __align__(8) uint64_t bar_ld[2];
  __shared__ __align__(8) uint64_t bar_dummy;
  __shared__ __align__(8) uint64_t bar_ready;
  __shared__ uint32_t tmem_ptr;
  __shared__ volatile int stop_flag;
  __shared__ float hammer_buf[2 * 32 * 36];
  for (int i = threadIdx.x; i < (81920 + 4 * N * 64 + 16384) / 2; i += blockDim.x) reinterpret_cast<__nv_bfloat16*>(gbase)[i] = __float2bfloat16(0.5f);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar_done)), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar_ld[0])), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar_ld[1])), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar_dummy)), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar_ready)), "r"(1));
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar_ready)) : "memory");   // phase 0 complete
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    stop_flag = 0;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "n"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_ptr;
  bool ok = true;
  long long c0 = 0, c1 = 0;
  unsigned long long copied = 0;
  if (warp == 0) {
    constexpr uint32_t idesc = make_idesc(128, N);
    c0 = clock64();
    uint32_t accum = 0;
    int cnt = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int s = 0; s < 8; ++s) {          // A: 8 regions of 10 KB (128 + 32 rows x 64 B) ; B at +80 KB, N rows x 64 B each
        const uint64_t a0 = make_desc(base + s * 10240 + (a_shift ? (s % 3) * 512 : 0), 16, 512, 4);
        const uint64_t b0 = make_desc(base + 81920 + (s & 3) * (N * 64), 16, 512, 4);
#pragma unroll
        bool early_ok = false;
        if (WAIT >= 4) early_ok = mbar_try(smem_u32(&bar_ready), 0);
        for (int k = 0; k < 2; ++k) {
          tc_mma<1>(tmem_base + (tile_len ? (uint32_t)(((cnt / tile_len) & 1) * N) : 0u), a0 + 2 * k, b0 + 2 * k, idesc,
                    (tile_len && cnt % tile_len == 0) ? 0u : accum);
          accum = 1; ++cnt;
        }
        // commit_every UMMAs: a tcgen05.commit to a barrier nobody waits on (the stage-free / tile-done signals of a pipeline);
        // WAIT: also poll an already-completed barrier + tcgen05.fence first, as a pipeline's "stage full" wait does
        if (commit_every && cnt % commit_every == 0) {
          tc_commit<1>(smem_u32(&bar_dummy));
          // WAIT 1: poll + fence, 2: poll only, 3: fence only, 4: the poll was issued BEFORE this group's MMAs (below) and is only
          // consumed here, + fence; 5: as 4 without the fence
          if (WAIT == 1 || WAIT == 2) mbar_wait(smem_u32(&bar_ready), 0);
          if (WAIT == 4 || WAIT == 5) { if (!early_ok) mbar_wait(smem_u32(&bar_ready), 0); }
          if (WAIT == 1 || WAIT == 3 || WAIT == 4) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
      }
    }
    tc_commit<1>(smem_u32(&bar_done));
    ok = mbar_wait(smem_u32(&bar_done), 0);
    c1 = clock64();
    stop_flag = 1;
  } else if (warp == 1 && (bg & 1)) {
    // destination: a 16 KB scratch behind the operands (never read by the MMAs)
    const uint32_t dst = base + 81920 + 4 * (N * 64);
    const uint8_t* my_src = src + (size_t)blockIdx.x * 65536;
    uint32_t ph[2] = {0, 0};
    int n = 0;
    if (lane == 0) {
      for (int b = 0; b < 2; ++b) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_ld[b])), "r"(8192) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst + b * 8192), "l"(my_src + b * 8192), "r"(8192), "r"(smem_u32(&bar_ld[b])) : "memory");
      }
      while (!stop_flag) {
        const int b = n & 1;
        mbar_wait(smem_u32(&bar_ld[b]), ph[b]); ph[b] ^= 1;
        copied += 8192;
        ++n;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_ld[b])), "r"(8192) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst + b * 8192), "l"(my_src + ((n + 1) & 7) * 8192), "r"(8192), "r"(smem_u32(&bar_ld[b])) : "memory");
      }
      mbar_wait(smem_u32(&bar_ld[0]), ph[0]);
      mbar_wait(smem_u32(&bar_ld[1]), ph[1]);
      res[blockIdx.x].ns = copied;
    }
  } else if (warp >= 2 && (bg & 2)) {
    float* my = hammer_buf + (warp - 2) * 32 * 36;
    float acc = 0.f;
    while (!stop_flag) {
#pragma unroll
      for (int g = 0; g < 8; ++g) *reinterpret_cast<float4*>(my + lane * 36 + g * 4) = make_float4(acc, 1.f, 2.f, 3.f);
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 32; ++r) acc += my[r * 36 + lane];
      __syncwarp();
    }
    if (acc == 12345.678f) res[0].bad = -1;
  }
  __syncthreads();
  if (threadIdx.x == 0) { res[blockIdx.x].cycles = (unsigned long long)(c1 - c0); res[blockIdx.x].timeout = ok ? 0 : 1; }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
}
template <int N, int commit_every = 0, int tile_len = 0, int WAIT = 0>
void run_sw64_bg(int grid, int iters, int bg, int a_shift) {
  const int smem = 81920 + 4 * N * 64 + 16384 + 1024;
  CK(cudaFuncSetAttribute(umma_rate_sw64_bg_kernel<N, commit_every, tile_len, WAIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  Result* d; CK(cudaMalloc(&d, sizeof(Result) * grid)); CK(cudaMemset(d, 0, sizeof(Result) * grid));
  uint8_t* src; CK(cudaMalloc(&src, (size_t)grid * 65536)); CK(cudaMemset(src, 0, (size_t)grid * 65536));
  for (int rep = 0; rep < 2; ++rep) { CK(cudaMemset(d, 0, sizeof(Result) * grid)); umma_rate_sw64_bg_kernel<N, commit_every, tile_len, WAIT><<<grid, 128, smem>>>(iters, bg, a_shift, src, d); CK(cudaDeviceSynchronize()); }
  std::vector<Result> h(grid); CK(cudaMemcpy(h.data(), d, sizeof(Result) * grid, cudaMemcpyDeviceToHost));
  double cyc = 0, bytes = 0; int to = 0;
  for (int i = 0; i < grid; ++i) { cyc += h[i].cycles; bytes += h[i].ns; to += h[i].timeout; }
  cyc /= grid; bytes /= grid;
  const double n_umma = (double)iters * 16;
  printf("SW64 M=128 N=%3d bulk-copy=%d hammer=%d tap-shift=%d commit-every=%2d tile=%2d wait=%d : %7.1f cycles/UMMA  %6.1f MAC/clk/SM  copies %5.1f B/clk  timeout=%d\n", N,
         bg & 1, (bg >> 1) & 1, a_shift, commit_every, tile_len, WAIT, cyc / n_umma, n_umma * 128.0 * N * 16.0 / cyc, bytes / cyc, to);
  fflush(stdout);
  CK(cudaFree(d)); CK(cudaFree(src));
}

template <int N, int A_MN, int B_MN>
void run_mn(int grid, int iters, int b_lbo) {
  const int smem = 98304 + 1024;
  CK(cudaFuncSetAttribute(umma_rate_mn_kernel<N, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  Result* d; CK(cudaMalloc(&d, sizeof(Result) * grid)); CK(cudaMemset(d, 0, sizeof(Result) * grid));
  for (int rep = 0; rep < 2; ++rep) {
    umma_rate_mn_kernel<N, A_MN, B_MN><<<grid, 128, smem>>>(iters, b_lbo, d);
    CK(cudaDeviceSynchronize());
  }
  std::vector<Result> h(grid); CK(cudaMemcpy(h.data(), d, sizeof(Result) * grid, cudaMemcpyDeviceToHost));
  const double n_umma = (double)iters * 16;
  double cyc = 0; int to = 0;
  for (int i = 0; i < grid; ++i) { cyc += h[i].cycles; to += h[i].timeout; }
  cyc /= grid;
  printf("cta_group::1 M=128 N=%3d A %s B %s (B atom stride %5d B) : %7.1f cycles/UMMA  %6.1f MAC/clk/SM  timeout=%d\n", N,
         A_MN ? "MN-major" : "K-major ", B_MN ? "MN-major" : "K-major ", b_lbo, cyc / n_umma, n_umma * 128.0 * N * 16.0 / cyc, to);
  fflush(stdout);
  CK(cudaFree(d));
}

template <int N, int CG>
void run(int grid, int iters, int nst, int hammer) {
  const int stage = 128 * 128 + (N / CG) * 128;
  const int smem = nst * stage + 1024;
  CK(cudaFuncSetAttribute(umma_rate_kernel<N, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  Result* d; CK(cudaMalloc(&d, sizeof(Result) * grid)); CK(cudaMemset(d, 0, sizeof(Result) * grid));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; ++rep) {      // rep 0 = warm-up
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, umma_rate_kernel<N, CG>, iters, nst, hammer, d));
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
  }
  float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<Result> h(grid); CK(cudaMemcpy(h.data(), d, sizeof(Result) * grid, cudaMemcpyDeviceToHost));
  const double n_umma = (double)iters * nst * 4;
  double cyc = 0, ns = 0; int cnt = 0, bad = 0, to = 0;
  for (int i = 0; i < grid; ++i) { bad += h[i].bad; to += h[i].timeout; if (CG == 2 && (i & 1)) continue; cyc += h[i].cycles; ns += h[i].ns; ++cnt; }
  cyc /= cnt; ns /= cnt;
  const double macs = n_umma * 128.0 * CG * N * 16.0;         // per issuing CTA (pair)
  const double tflops = 2.0 * macs * cnt / (ns * 1e-9) / 1e12;
  printf("cta_group::%d M=%3d N=%3d stages=%d grid=%3d hammer=%d : %7.1f cycles/UMMA  %6.1f MAC/clk/SM  %7.1f TFLOP/s (in-kernel)  kernel %.3f ms  bad=%d timeout=%d\n",
         CG, 128 * CG, N, nst, grid, hammer, cyc / n_umma, macs / cyc / CG, tflops, ms, bad, to);
  fflush(stdout);
  CK(cudaFree(d));
}

int main(int argc, char** argv) {
  int dev = 0; CK(cudaSetDevice(dev));
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = sms & ~1;
  const int iters = argc > 1 ? atoi(argv[1]) : 512;
  printf("SMs %d, %d UMMAs per CTA per launch\n", sms, iters * 4 * 4);
  if (argc > 4) {     // cost of tcgen05.commit / of a (satisfied) barrier wait / of the tcgen05 fence inside the instruction stream
    run_sw64_bg<32, 0, 0, 0>(grid, iters, 0, 1);
    run_sw64_bg<32, 2, 0, 0>(grid, iters, 0, 1);
    run_sw64_bg<32, 2, 0, 1>(grid, iters, 0, 1);
    run_sw64_bg<32, 2, 0, 2>(grid, iters, 0, 1);
    run_sw64_bg<32, 2, 0, 3>(grid, iters, 0, 1);
    run_sw64_bg<32, 2, 0, 4>(grid, iters, 0, 1);
    run_sw64_bg<32, 2, 0, 5>(grid, iters, 0, 1);
    run_sw64_bg<128, 2, 0, 1>(grid, iters, 0, 1);
    run_sw64_bg<128, 2, 0, 5>(grid, iters, 0, 1);
    return 0;
  }
  if (argc > 3) {     // SWIZZLE_64B operands under shared-memory-port contention (stage-1 conv configuration)
    for (int bg = 0; bg < 4; ++bg) {
      run_sw64_bg<32>(grid, iters, bg, 0);
      run_sw64_bg<32>(grid, iters, bg, 1);
      run_sw64_bg<64>(grid, iters, bg, 1);
      run_sw64_bg<96>(grid, iters, bg, 1);
      run_sw64_bg<192>(grid, iters, bg, 1);
    }
    return 0;
  }
  if (argc > 2) {     // operand-major sweep
    run_mn<128, 0, 0>(grid, iters, 16);
    run_mn<128, 1, 0>(grid, iters, 16);
    run_mn<128, 0, 1>(grid, iters, 8192);
    run_mn<128, 1, 1>(grid, iters, 8192);
    run_mn<128, 1, 1>(grid, iters, 128);
    run_mn<192, 0, 0>(grid, iters, 16);
    run_mn<192, 1, 0>(grid, iters, 16);
    run_mn<192, 0, 1>(grid, iters, 8192);
    run_mn<192, 1, 1>(grid, iters, 8192);
    run_mn<192, 1, 1>(grid, iters, 128);
    run_mn<256, 1, 1>(grid, iters, 8192);
    run_mn<96, 1, 1>(grid, iters, 128);
    run_mn<96, 0, 0>(grid, iters, 16);
    run_mn<32, 0, 0>(grid, iters, 16);
    run_mn<64, 0, 0>(grid, iters, 16);
    run_sw64<32>(grid, iters);
    run_sw64<64>(grid, iters);
    run_sw64<128>(grid, iters);
    return 0;
  }
  for (int hammer = 0; hammer < 2; ++hammer) {
    run<32, 1>(grid, iters, 4, hammer);
    run<64, 1>(grid, iters, 4, hammer);
    run<128, 1>(grid, iters, 4, hammer);
    run<256, 1>(grid, iters, 4, hammer);
    run<64, 2>(grid, iters, 4, hammer);
    run<128, 2>(grid, iters, 4, hammer);
    run<256, 2>(grid, iters, 4, hammer);
  }
  run<128, 1>(2, iters, 4, 0);
  run<256, 1>(2, iters, 4, 0);
  run<256, 2>(2, iters, 4, 0);
  return 0;
}
