// Weight gradient of 3x3 / stride-1 convolutions with C = Cin = Cout in {32, 64} (stages 1-2 of ResNet-34).
//
// The generic wgrad kernel (conv_tc.cu) issues one M=128 UMMA per tap with only C (32/64) real rows of dy^T, so 50-75 %
// of every MMA is wasted and the tensor pipe is the bound (profiles/r01_conv_stage1.md: 61 cycles per UMMA, 46 % busy).
// Here the roles are turned around:
//     dw[co, (r,s), ci] = sum_q  dy[q - (r-1, s-1), co] * x[q, ci]          (q runs over the INPUT pixels)
// so x is read UNSHIFTED (operand B, N = C) and dy is read shifted (operand A).  For a filter column s one TMA load
// fetches a (bh+2) x bw halo tile of dy; its three row-shifted views r = 2,1,0 are three MN-major "atoms" of the SAME smem
// tile, bw rows apart — expressed to the tensor core simply as LBO = bw*ROWB — so ONE UMMA (M = 128 = up to 4 atoms of 32
// channels, or 2 atoms of 64) multiplies all stacked taps at once: 3 (C=32) or 6 (C=64) UMMAs per K step instead of 9.
// One CTA owns all 9 taps of a pixel range (split-K over all SMs); partials go to the workspace with plain stores and are
// reduced by wgrad_reduce_kernel (conv_simt.cu), exactly like the generic path.
#include "tc_common.cuh"

namespace {

struct Wgrad3P {
  int bh, bw, P;               // x pixel tile; P = bh*bw, multiple of 16, <= 128; bw % 8 == 0
  int tiles_h, tiles_w, num_pix_tiles;
  int C;                       // Cin = Cout
  int ksplit, tiles_per;
  int dy_stage_bytes;          // (P + 3*bw) * ROWB rounded up to 1024: halo tile + room for the deepest atom view
  int n_dy_stages;
  float* ws;                   // [ksplit][9][C][C]
  long long ws_stride;
};

template <int CK>   // CK = C: 32 -> SWIZZLE_64B atoms, 64 -> SWIZZLE_128B atoms
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_wgrad3_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX,
                      const __grid_constant__ Wgrad3P p) {
  constexpr int ROWB = CK * 2;
  constexpr uint32_t LAYOUT = (CK == 64) ? 2u : 4u;
  constexpr uint32_t SBO = 8 * ROWB;
  constexpr int MMAS = (CK == 32) ? 1 : 2;            // UMMAs per K step and filter column
  constexpr int ACC_COLS = CK;                        // one accumulator = 128 lanes x C columns
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t X_BYTES = (uint32_t)p.P * ROWB;      // multiple of 1024 (P % 16 == 0)
  const uint32_t x0 = base;                           // 2 x slots
  const uint32_t d0 = base + 2 * X_BYTES;             // dy halo ring
  const uint32_t auxoff = 2 * X_BYTES + (uint32_t)p.n_dy_stages * p.dy_stage_bytes;
  const uint32_t aux = base + auxoff;
  const uint32_t bar_xfull = aux, bar_xempty = aux + 16, bar_dfull = aux + 32, bar_dempty = aux + 96, bar_done = aux + 160;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(gbase + auxoff + 176);

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(bar_xfull + 8 * s, 1); mbar_init(bar_xempty + 8 * s, 1); }
    for (int s = 0; s < p.n_dy_stages; ++s) { mbar_init(bar_dfull + 8 * s, 1); mbar_init(bar_dempty + 8 * s, 1); }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int ks = blockIdx.x;
  const int t_beg = ks * p.tiles_per;
  const int t_end = (t_beg + p.tiles_per) < p.num_pix_tiles ? (t_beg + p.tiles_per) : p.num_pix_tiles;
  const uint32_t halo_bytes = (uint32_t)(p.bh + 2) * p.bw * ROWB;

  if (warp == 0) {
    {   // the WHOLE warp runs this loop (warp-uniform control flow lets ptxas keep descriptors and coordinates in
        // uniform registers); the issuing wrappers elect one lane
      int xs = 0; uint32_t xph = 0;
      int ds = 0; uint32_t dph = 0;
      for (int tile = t_beg; tile < t_end; ++tile) {
        int pt = tile;
        const int tw = pt % p.tiles_w; pt /= p.tiles_w;
        const int th = pt % p.tiles_h;
        const int n = pt / p.tiles_h;
        const int h0 = th * p.bh, w0 = tw * p.bw;
        mbar_wait(bar_xempty + 8 * xs, xph ^ 1u);
        mbar_expect_tx(bar_xfull + 8 * xs, X_BYTES);
        tma_load_4d(x0 + xs * X_BYTES, &tmX, bar_xfull + 8 * xs, 0, w0, h0, n);
        if (++xs == 2) { xs = 0; xph ^= 1u; }
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          mbar_wait(bar_dempty + 8 * ds, dph ^ 1u);
          mbar_expect_tx(bar_dfull + 8 * ds, halo_bytes);
          tma_load_4d(d0 + ds * p.dy_stage_bytes, &tmDy, bar_dfull + 8 * ds, 0, w0 + 1 - s, h0 - 1, n);
          if (++ds == p.n_dy_stages) { ds = 0; dph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    {   // the WHOLE warp runs this loop (warp-uniform control flow lets ptxas keep descriptors and coordinates in
        // uniform registers); the issuing wrappers elect one lane
      constexpr uint32_t idesc = make_idesc(128, CK, 1, 1);     // both operands MN-major
      int xs = 0; uint32_t xph = 0;
      int ds = 0; uint32_t dph = 0;
      const int ksteps = p.P / 16;
      const uint32_t lbo = (uint32_t)p.bw * ROWB;               // atom a = halo rows shifted by a*bw  (tap r = 2 - a)
      for (int tile = t_beg; tile < t_end; ++tile) {
        mbar_wait(bar_xfull + 8 * xs, xph);
        tc_fence_after();
        const uint64_t bd0 = make_desc(x0 + xs * X_BYTES, X_BYTES, SBO, LAYOUT);
        const uint32_t acc_flag = tile != t_beg ? 1u : 0u;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          mbar_wait(bar_dfull + 8 * ds, dph);
          tc_fence_after();
          const uint32_t dbase = d0 + ds * p.dy_stage_bytes;
#pragma unroll
          for (int j = 0; j < MMAS; ++j) {
            // C=32: one MMA, atoms a=0..3 -> taps r=2,1,0,(garbage).  C=64: MMA 0 = atoms (r=2, r=1), MMA 1 = (r=0, garbage)
            uint64_t ad = make_desc(dbase + (uint32_t)(j * 2) * lbo, lbo, SBO, LAYOUT);
            uint64_t bd = bd0;
            const uint32_t d_tmem = tmem_base + (uint32_t)((s * MMAS + j) * ACC_COLS);
            tc_mma(d_tmem, ad, bd, idesc, acc_flag);
            for (int k = 1; k < ksteps; ++k) {
              ad += ROWB; bd += ROWB;                            // 16 pixel rows = 16*ROWB bytes = ROWB 16-byte units
              tc_mma(d_tmem, ad, bd, idesc, 1u);
            }
          }
          tc_commit(bar_dempty + 8 * ds);
          if (++ds == p.n_dy_stages) { ds = 0; dph ^= 1u; }
        }
        tc_commit(bar_xempty + 8 * xs);
        if (++xs == 2) { xs = 0; xph ^= 1u; }
      }
      tc_commit(bar_done);
    }
  } else {
    mbar_wait(bar_done, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int m = q * 32 + lane;                        // accumulator row = atom * CK + co
    const int atom = m / CK, co = m % CK;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    float* wsl = p.ws + (long long)ks * p.ws_stride;
    for (int s = 0; s < 3; ++s) {
      for (int j = 0; j < MMAS; ++j) {
        const int r = 2 - (j * 2 + atom);               // filter row of this accumulator row
        for (int c = 0; c < CK / 32; ++c) {
          uint32_t v[32];
          tc_ld32(taddr + (uint32_t)((s * MMAS + j) * ACC_COLS + c * 32), v);
          if (r >= 0) {
            float4* dst = reinterpret_cast<float4*>(wsl + ((long long)(r * 3 + s) * p.C + co) * p.C + c * 32);
#pragma unroll
            for (int e = 0; e < 8; ++e)
              dst[e] = make_float4(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]), __uint_as_float(v[4 * e + 2]),
                                   __uint_as_float(v[4 * e + 3]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

int plan3(const svk_conv_desc* d, Wgrad3P* pp, size_t* smem_out) {
  Wgrad3P& p = *pp;
  const int C = d->Cin, ROWB = C * 2;
  p.C = C;
  // tile: bw % 8 == 0, P % 16 == 0, P <= 128; minimise tiles * (tensor cycles + load latency)
  double best = 1e30; p.bh = 0;
  for (int bw = 8; bw <= 128; bw += 8) {
    for (int bh = 1; bh * bw <= 128 && bh <= 250; ++bh) {
      if ((bh * bw) % 16) continue;
      int th = (d->H + bh - 1) / bh, tw = (d->W + bw - 1) / bw;
      double mma = 3.0 * (C == 32 ? 1 : 2) * (bh * bw / 16) * 64.0;
      double ld = ((bh * bw) + 3.0 * (bh + 2) * bw) * ROWB / 48.0 + 4 * 250.0;
      double cost = (double)th * tw * (mma > ld ? mma : ld);
      if (cost < best - 1e-9) { best = cost; p.bh = bh; p.bw = bw; }
    }
  }
  SVK_REQUIRE(p.bh > 0, SVK_E_UNSUPPORTED, "conv2d_wgrad3: no pixel tile for %dx%d", d->H, d->W);
  p.P = p.bh * p.bw;
  p.tiles_h = (d->H + p.bh - 1) / p.bh;
  p.tiles_w = (d->W + p.bw - 1) / p.bw;
  p.num_pix_tiles = d->N * p.tiles_h * p.tiles_w;
  int ks = svk_num_sms();
  if (ks > p.num_pix_tiles) ks = p.num_pix_tiles;
  p.tiles_per = (p.num_pix_tiles + ks - 1) / ks;
  p.ksplit = (p.num_pix_tiles + p.tiles_per - 1) / p.tiles_per;
  p.ws_stride = (long long)9 * C * C;
  p.dy_stage_bytes = ((p.P + 3 * p.bw) * ROWB + 1023) / 1024 * 1024;
  const size_t xb = (size_t)2 * p.P * ROWB;
  int ns = (int)((190 * 1024 - xb - 2048) / p.dy_stage_bytes);
  if (ns > 6) ns = 6;
  SVK_REQUIRE(ns >= 3, SVK_E_UNSUPPORTED, "conv2d_wgrad3: not enough shared memory");
  p.n_dy_stages = ns;
  *smem_out = xb + (size_t)ns * p.dy_stage_bytes + 2048;
  return 0;
}

}  // namespace

bool svk_wgrad3_applicable(const svk_conv_desc* d) {
  return d->R == 3 && d->stride == 1 && d->Cin == d->Cout && (d->Cin == 32 || d->Cin == 64);
}

size_t svk_conv2d_wgrad3_tc_ws_floats(const svk_conv_desc* d) {
  Wgrad3P p{}; size_t smem;
  if (plan3(d, &p, &smem)) return 0;
  return (size_t)p.ksplit * (size_t)p.ws_stride;
}

int svk_conv2d_wgrad3_tc(const svk_conv_desc* d, const void* x, const void* dy, float* ws, size_t ws_floats, int* ksplit_out,
                         cudaStream_t st) {
  Wgrad3P p{}; size_t smem;
  if (int e = plan3(d, &p, &smem)) return e;
  SVK_REQUIRE((size_t)p.ksplit * (size_t)p.ws_stride <= ws_floats, SVK_E_BADARG, "conv2d_wgrad3: workspace too small");
  p.ws = ws;
  const int C = d->Cin;
  CUtensorMap tdy, tx;
  if (int e = make_nhwc_map(&tdy, dy, d->N, d->Ho, d->Wo, C, C, p.bw, p.bh + 2, 1)) return e;
  if (int e = make_nhwc_map(&tx, x, d->N, d->H, d->W, C, C, p.bw, p.bh, 1)) return e;
  if (C == 64) {
    static bool cfg = false;
    if (!cfg) { cudaError_t e = cudaFuncSetAttribute(conv_tc_wgrad3_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      SVK_REQUIRE(e == cudaSuccess, (int)e, "conv_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); cfg = true; }
    conv_tc_wgrad3_kernel<64><<<p.ksplit, TC_THREADS, smem, st>>>(tdy, tx, p);
  } else {
    static bool cfg = false;
    if (!cfg) { cudaError_t e = cudaFuncSetAttribute(conv_tc_wgrad3_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      SVK_REQUIRE(e == cudaSuccess, (int)e, "conv_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); cfg = true; }
    conv_tc_wgrad3_kernel<32><<<p.ksplit, TC_THREADS, smem, st>>>(tdy, tx, p);
  }
  SVK_LAUNCH_CHECK("conv_tc_wgrad3");
  *ksplit_out = p.ksplit;
  return 0;
}
