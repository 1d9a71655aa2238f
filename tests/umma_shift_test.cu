// Hardware semantics probe: can a UMMA shared-memory descriptor start at (or step between atoms by) an offset that is NOT a
// multiple of the swizzle pattern (8 rows), i.e. "shift the operand view by one pixel row"?  Two cases the convolution
// kernels would use:
//   1. K-major SWIZZLE_128B operand A (fprop/dgrad): start address = tile + shift * 128 B  (a +-1 pixel tap shift)
//   2. MN-major SWIZZLE_64B / 128B operand B (wgrad): N = 3 atoms whose stride (LBO) is ONE row, i.e. three views of the
//      same tile shifted by 0/1/2 pixels stacked along N.
// Every case is tried with the descriptor's base-offset field = 0 and = (start >> 7) & 7.  Shared memory is filled exactly
// as TMA would fill it (swizzle = XOR of address bits), results are compared with the shifted logical operand.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_shift_test umma_shift_test.cu && ./umma_shift_test
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)layout << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct Params {
  int mode;        // 1: K-major SW128 A shift;  2: MN-major SW64 B with LBO = 1 row;  3: MN-major SW128 B with LBO = 1 row
  int shift;       // rows
  int use_base;    // 0: base offset 0;  1: (start >> 7) & 7
  int pass;        // 0: value = row;  1: value = column id (+ 64 * row%4 | + 32 * row%8)
  int lbo_rows;    // modes 2/3: rows between the N atoms (1 = the case of interest; 8 = aligned control)
};

// D[128][N] (fp32) out
__global__ void __launch_bounds__(128, 1) shift_kernel(Params P, float* D, int N) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* g = smem_raw + (base - smem_u32(smem_raw));
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  // regions: OPA at 0 (32 KB), OPB at 32 KB (64 KB)
  uint8_t* opa = g;
  uint8_t* opb = g + 32768;
  for (int i = threadIdx.x; i < 98304 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(g)[i] = 0u;
  __syncthreads();
  auto put = [](uint8_t* p, float v) { *reinterpret_cast<__nv_bfloat16*>(p) = __float2bfloat16(v); };
  if (P.mode == 1) {
    // A: 256 rows x 64 k, K-major SW128:  byte = row*128 + ((k/8) ^ (row%8))*16 + (k%8)*2
    for (int i = threadIdx.x; i < 256 * 64; i += blockDim.x) {
      const int row = i / 64, k = i % 64;
      const float v = P.pass == 0 ? (float)row : (float)(k + 64 * (row % 4));
      put(opa + row * 128 + (((k / 8) ^ (row % 8)) * 16) + (k % 8) * 2, v);
    }
    // B: N = 64 rows x 64 k one-hot (n == k), K-major SW128
    for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
      const int n = i / 64, k = i % 64;
      put(opb + n * 128 + (((k / 8) ^ (n % 8)) * 16) + (k % 8) * 2, n == k ? 1.f : 0.f);
    }
  } else {
    const int CKc = (P.mode == 2) ? 32 : 64;            // channels per atom row
    const int ROWB = CKc * 2;
    // A: one-hot [k][m] = (m == k), k < 16, m < 128, MN-major: atom a = m / CKc at a * (16 * ROWB), row k, swizzled chunk
    for (int i = threadIdx.x; i < 16 * 128; i += blockDim.x) {
      const int k = i / 128, m = i % 128;
      const int a = m / CKc, c = m % CKc;
      const int sw = (P.mode == 2) ? ((k >> 1) & 3) : (k & 7);
      put(opa + a * 16 * ROWB + k * ROWB + (((c / 8) ^ sw) * 16) + (c % 8) * 2, m == k ? 1.f : 0.f);
    }
    // B tile X: 256 rows (pixels) x CKc channels, as TMA writes it
    for (int i = threadIdx.x; i < 256 * CKc; i += blockDim.x) {
      const int row = i / CKc, c = i % CKc;
      const int sw = (P.mode == 2) ? ((row >> 1) & 3) : (row & 7);
      const float v = P.pass == 0 ? (float)row : (float)(c + CKc * (row % (256 / CKc)));
      put(opb + row * ROWB + (((c / 8) ^ sw) * 16) + (c % 8) * 2, v);
    }
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
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
  if (threadIdx.x == 0) {
    if (P.mode == 1) {
      const uint32_t idesc = make_idesc(128, 64, 0, 0);
      const uint32_t a_addr = base + P.shift * 128;
      const uint32_t boff = P.use_base ? ((a_addr >> 7) & 7) : 0;
      const uint64_t ad = make_desc(a_addr, 16, 1024, 2, boff);
      const uint64_t bd = make_desc(base + 32768, 16, 1024, 2, 0);
      for (int k = 0; k < 4; ++k) {
        const uint32_t acc = k ? 1u : 0u;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_base), "l"(ad + 2 * k), "l"(bd + 2 * k), "r"(idesc), "r"(acc) : "memory");
      }
    } else {
      const int CKc = (P.mode == 2) ? 32 : 64;
      const int ROWB = CKc * 2;
      const uint32_t layout = (P.mode == 2) ? 4u : 2u;
      const uint32_t idesc = make_idesc(128, 3 * CKc, 1, 1);
      const uint32_t b_addr = base + 32768 + P.shift * ROWB;
      const uint32_t boff = P.use_base ? ((b_addr >> 7) & 7) : 0;
      const uint64_t ad = make_desc(base, 16 * ROWB, 8 * ROWB, layout, 0);
      const uint64_t bd = make_desc(b_addr, P.lbo_rows * ROWB, 8 * ROWB, layout, boff);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem_base), "l"(ad), "l"(bd), "r"(idesc), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  {
    long long t0 = clock64();
    while (!mbar_try(smem_u32(&bar), 0)) { if (clock64() - t0 > 2000000000LL) break; }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t r[8];
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int e = 0; e < 8; ++e) D[threadIdx.x * N + c0 + e] = __uint_as_float(r[e]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
}

static int run_case(Params P) {
  const int N = (P.mode == 1) ? 64 : (P.mode == 2 ? 96 : 192);
  float* d; CK(cudaMalloc(&d, 128 * N * sizeof(float))); CK(cudaMemset(d, 0xff, 128 * N * sizeof(float)));
  CK(cudaFuncSetAttribute(shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  shift_kernel<<<1, 128, 100 * 1024>>>(P, d, N);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  kernel failed: %s\n", cudaGetErrorString(e)); exit(2); }
  std::vector<float> h(128 * N); CK(cudaMemcpy(h.data(), d, h.size() * sizeof(float), cudaMemcpyDeviceToHost));
  CK(cudaFree(d));
  int bad = 0, rows = (P.mode == 1) ? 128 : 16;
  for (int m = 0; m < rows; ++m)
    for (int n = 0; n < N; ++n) {
      float want;
      if (P.mode == 1) {
        const int row = m + P.shift;
        want = P.pass == 0 ? (float)row : (float)(n + 64 * (row % 4));
      } else {
        const int CKc = (P.mode == 2) ? 32 : 64;
        const int row = m + P.shift + (n / CKc) * P.lbo_rows, c = n % CKc;
        want = P.pass == 0 ? (float)row : (float)(c + CKc * (row % (256 / CKc)));
      }
      if (h[m * N + n] != want) {
        if (bad < 3) printf("    mismatch m=%d n=%d got %g want %g\n", m, n, h[m * N + n], want);
        ++bad;
      }
    }
  return bad;
}

int main() {
  CK(cudaSetDevice(0));
  const int shifts[] = {0, 1, 2, 5, 8, 9, 13};
  for (int mode = 1; mode <= 3; ++mode) {
    for (int lbo_rows = (mode == 1 ? 1 : 8); lbo_rows >= 1; lbo_rows -= 7) {
      for (int si = 0; si < 7; ++si) {
        for (int ub = 0; ub < 2; ++ub) {
          int bad = 0;
          for (int pass = 0; pass < 2; ++pass) {
            Params P{mode, shifts[si], ub, pass, lbo_rows};
            bad += run_case(P);
          }
          printf("mode %d (%s) lbo_rows=%d shift=%2d base_offset=%s : %s (%d mismatches)\n", mode,
                 mode == 1 ? "K-major SW128 A" : (mode == 2 ? "MN-major SW64 B, N=96" : "MN-major SW128 B, N=192"), lbo_rows,
                 shifts[si], ub ? "(addr>>7)&7" : "0", bad ? "WRONG" : "ok", bad);
          fflush(stdout);
        }
      }
      if (mode == 1) break;
    }
  }
  return 0;
}
