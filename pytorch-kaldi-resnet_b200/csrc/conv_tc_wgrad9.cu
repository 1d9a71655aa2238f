// Weight gradient of 3x3 / stride-1 convolutions with C = Cin = Cout in {32, 64} (stages 1-2 of ResNet-34): ALL NINE taps
// in one tcgen05.mma per 16 pixels.
//
//     dw[co, (r,s), ci] = sum_{h,w}  dy[h, w, co] * x[h + r - 1, w + s - 1, ci]
//
// A UMMA shared-memory descriptor may start at ANY 16-byte-aligned address of a swizzled tile (the swizzle is a pure
// function of the absolute smem address bits — probed on the hardware, tests/umma_shift_test.cu), and so may the stride
// between the MN-atoms of an MN-major operand.  So both tap shifts become operand *views* of two TMA tiles:
//   * a tile is bh full image rows of PW >= W + 2 columns (the columns beyond the image are zero-filled by TMA);
//     K runs over its bh*PW smem rows, 16 per instruction;
//   * operand A = dy, halo tile of bh+2 image rows: MN-atom a (C channels) starts a*PW rows further down (LBO = one image
//     row): atoms a = 0,1,2 are the taps r = 2,1,0 stacked along M  (C = 32: M = 128 holds all three; C = 64: two MMAs);
//   * operand B = x tile loaded from column -1: MN-atom b starts b rows (pixels) further (LBO = ONE pixel): atoms
//     b = 0,1,2 are the taps s = 0,1,2 stacked along N = 3C.
//   D[(a, co), (b, ci)] += sum_k dyTile[k + a*PW][co] * xTile[k + b][ci]  is the whole 3x3 gradient: 1 (C=32) or 2 (C=64)
//   UMMAs per 16 pixels instead of 9 (conv_tc.cu).  Where a shifted view wraps into the
//   next image row it meets a zero of the other operand (dy columns >= W, x column -1), so no masking is needed.
// One CTA owns all 9 taps of a range of tiles (split-K over all SMs); partials go to the workspace with plain stores and
// are reduced by wgrad_reduce_*_kernel (conv_simt.cu), like the other wgrad paths.
#include "tc_common.cuh"

namespace {

struct Wgrad9P {
  int bh, PW;                  // tile: bh image rows x PW padded columns; bh*PW % 16 == 0
  int tiles_h, num_tiles;      // tiles per image, N * tiles_h
  int C;
  int ksplit, tiles_per;
  int x_bytes, dy_bytes;       // bytes the two TMA loads of a stage deliver
  int x_alloc, stage_bytes;    // smem: [x tile + 8 zero rows][dy halo tile + room for the deepest atom view], 1024-aligned
  int n_stages;
  float* ws;                   // [ksplit][9][C][C]
  long long ws_stride;
  unsigned long long* prof;    // SVK_PROF=1 cycle counters (tc_common.cuh), else NULL
};

template <int CK>   // CK = C: 32 -> SWIZZLE_64B atoms, 64 -> SWIZZLE_128B atoms
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_wgrad9_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX,
                      const __grid_constant__ Wgrad9P p) {
  constexpr int ROWB = CK * 2;
  constexpr uint32_t LAYOUT = (CK == 64) ? 2u : 4u;
  constexpr uint32_t SBO = 8 * ROWB;
  constexpr int MMAS = (CK == 32) ? 1 : 2;            // UMMAs per K step
  constexpr int ACC_COLS = 3 * CK;                    // one accumulator = 128 lanes x 3C columns
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t auxoff = (uint32_t)p.n_stages * p.stage_bytes;
  const uint32_t aux = base + auxoff;
  const uint32_t bar_full = aux, bar_empty = aux + 64, bar_done = aux + 128;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(gbase + auxoff + 144);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.n_stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // the 8 rows behind every x tile are read by the views b = 1, 2 of its last K step (times dy zeros): keep them finite
  for (int s = 0; s < p.n_stages; ++s) {
    uint32_t* tail = reinterpret_cast<uint32_t*>(gbase + (size_t)s * p.stage_bytes + p.x_bytes);
    for (int i = threadIdx.x; i < 8 * ROWB / 4; i += blockDim.x) tail[i] = 0u;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_wait();             // the prologue above overlapped the previous kernel's tail; global memory is touched from here on
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int ks = blockIdx.x;
  const int t_beg = ks * p.tiles_per;
  const int t_end = (t_beg + p.tiles_per) < p.num_tiles ? (t_beg + p.tiles_per) : p.num_tiles;

  long long et0 = 0, ew = 0;     // SVK_PROF: epilogue timing
  if (warp == 0) {
    {   // the WHOLE warp runs this loop (uniform control flow); the issuing wrappers elect one lane
      int st = 0; uint32_t ph = 0;
      const bool prof = p.prof != nullptr;
      long long pw = 0; const long long pt0 = prof ? clock64() : 0;
      for (int tile = t_beg; tile < t_end; ++tile) {
        const int th = tile % p.tiles_h;
        const int n = tile / p.tiles_h;
        const int h0 = th * p.bh;
        const uint32_t sb = base + (uint32_t)st * p.stage_bytes;
        mbar_wait_t(bar_empty + 8 * st, ph ^ 1u, prof, pw);
        mbar_expect_tx(bar_full + 8 * st, (uint32_t)(p.x_bytes + p.dy_bytes));
        tma_load_4d(sb, &tmX, bar_full + 8 * st, 0, -1, h0, n);
        tma_load_4d(sb + p.x_alloc, &tmDy, bar_full + 8 * st, 0, 0, h0 - 1, n);
        if (++st == p.n_stages) { st = 0; ph ^= 1u; }
      }
      if (prof) prof_flush(p.prof, 4, clock64() - pt0, pw, lane);
    }
  } else if (warp == 1) {
    {   // the WHOLE warp runs this loop; one elected lane issues
      constexpr uint32_t idesc = make_idesc(128, 3 * CK, 1, 1);     // both operands MN-major
      int st = 0; uint32_t ph = 0;
      const int ksteps = (p.bh * p.PW) / 16;
      const uint32_t lbo_a = (uint32_t)p.PW * ROWB;                 // dy atom a = rows shifted by a image rows (tap r = 2 - a)
      const bool prof = p.prof != nullptr;
      long long pwf = 0; const long long pt0 = prof ? clock64() : 0;
      unsigned long long gt0 = 0;
      if (prof) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0));
      for (int tile = t_beg; tile < t_end; ++tile) {
        mbar_wait_t(bar_full + 8 * st, ph, prof, pwf);
        tc_fence_after();
        const uint32_t sb = base + (uint32_t)st * p.stage_bytes;
        uint64_t bd = make_desc(sb, ROWB, SBO, LAYOUT);              // x atom b = rows shifted by b pixels (tap s = b)
        uint64_t ad = make_desc(sb + p.x_alloc, lbo_a, SBO, LAYOUT);
        const uint64_t ad_step2 = (uint64_t)((2u * lbo_a) >> 4);     // C = 64: second MMA starts two image rows down
        uint32_t acc_flag = tile != t_beg ? 1u : 0u;
        for (int k = 0; k < ksteps; ++k) {
          tc_mma(tmem_base, ad, bd, idesc, acc_flag);
          if (MMAS == 2) tc_mma(tmem_base + ACC_COLS, ad + ad_step2, bd, idesc, acc_flag);
          acc_flag = 1u;
          ad += ROWB; bd += ROWB;                                    // 16 rows = 16*ROWB bytes = ROWB 16-byte units
        }
        tc_commit(bar_empty + 8 * st);
        if (++st == p.n_stages) { st = 0; ph ^= 1u; }
      }
      tc_commit(bar_done);
      if (prof) {
        mbar_wait(bar_done, 0);        // include the drain of the last MMAs
        if (lane == 0) {
          atomicAdd(p.prof + 0, 1ull); atomicAdd(p.prof + 1, (unsigned long long)(clock64() - pt0));
          atomicAdd(p.prof + 2, (unsigned long long)pwf);
          unsigned long long gt1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
          atomicAdd(p.prof + 8, gt1 - gt0); atomicMax(p.prof + 9, ~gt0); atomicMax(p.prof + 10, gt1);
        }
      }
    }
  } else {
    et0 = p.prof ? clock64() : 0;
    mbar_wait_t(bar_done, 0, p.prof != nullptr, ew);
    tc_fence_after();
    const int q = warp & 3;
    const int m0 = q * 32;                              // first accumulator row of this warp: row = atom * CK + co
    const int atom = m0 / CK, co0 = m0 % CK;            // (uniform per warp: CK is a multiple of 32)
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    float* wsl = p.ws + (long long)ks * p.ws_stride;
    float* scr = reinterpret_cast<float*>(gbase) + q * 1024;      // the operand stages are idle now: 4 KB of scratch per warp
    for (int j = 0; j < MMAS; ++j) {
      const int r = 2 - (j * 2 + atom);                 // filter row of these accumulator rows (< 0: padding rows)
      for (int b = 0; b < 3; ++b) {
        for (int c = 0; c < CK / 32; ++c) {
          uint32_t v[32];
          tc_ld32(taddr + (uint32_t)(j * ACC_COLS + b * CK + c * 32), v);
          if (r >= 0) store_chunk_rows(scr, v, wsl + ((long long)(r * 3 + b) * p.C + co0) * p.C + c * 32, p.C, lane);
        }
      }
    }
  }
  if (p.prof && warp == 2) prof_flush(p.prof, 6, clock64() - et0, ew, lane);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

constexpr int W9_SMEM_MAX = 220 * 1024;

int plan9(const svk_conv_desc* d, Wgrad9P* pp, size_t* smem_out) {
  Wgrad9P& p = *pp;
  const int C = d->Cin, ROWB = C * 2;
  p.C = C;
  // bh image rows per tile; PW = smallest padded width >= W + 2 with bh*PW % 16 == 0.  Prefer the tallest tile (least
  // halo re-reading: (2*bh + 2) / (2*bh) rows fetched per row of work) that leaves room for 2 stages.
  p.bh = 0;
  for (int bh = 4; bh >= 1; --bh) {
    const int mult = (bh % 2 == 0) ? ((bh % 4 == 0) ? 4 : 8) : 16;
    const int PW = (d->W + 2 + mult - 1) / mult * mult;
    if (PW > 256 || bh + 2 > 256) continue;
    const int x_alloc = ((bh * PW + 8) * ROWB + 1023) / 1024 * 1024;
    const int dy_alloc = ((bh + 3) * PW * ROWB + 1023) / 1024 * 1024;
    const int stage = x_alloc + dy_alloc;
    const int ns = (W9_SMEM_MAX - 2048) / stage;
    if (ns < 2) continue;
    p.bh = bh; p.PW = PW; p.x_alloc = x_alloc; p.stage_bytes = stage; p.n_stages = ns > 4 ? 4 : ns;
    break;
  }
  SVK_REQUIRE(p.bh > 0, SVK_E_UNSUPPORTED, "conv2d_wgrad9: no tile for %dx%d", d->H, d->W);
  p.x_bytes = p.bh * p.PW * ROWB;
  p.dy_bytes = (p.bh + 2) * p.PW * ROWB;
  p.tiles_h = (d->H + p.bh - 1) / p.bh;
  p.num_tiles = d->N * p.tiles_h;
  int ks = svk_num_sms();
  if (ks > p.num_tiles) ks = p.num_tiles;
  p.tiles_per = (p.num_tiles + ks - 1) / ks;
  p.ksplit = (p.num_tiles + p.tiles_per - 1) / p.tiles_per;
  p.ws_stride = (long long)9 * C * C;
  *smem_out = (size_t)p.n_stages * p.stage_bytes + 2048;
  return 0;
}

}  // namespace

bool svk_wgrad9_applicable(const svk_conv_desc* d) {
  if (!(d->R == 3 && d->stride == 1 && d->Cin == d->Cout && (d->Cin == 32 || d->Cin == 64))) return false;
  Wgrad9P p{}; size_t smem;
  return plan9(d, &p, &smem) == 0;
}

size_t svk_conv2d_wgrad9_tc_ws_floats(const svk_conv_desc* d) {
  Wgrad9P p{}; size_t smem;
  if (plan9(d, &p, &smem)) return 0;
  return (size_t)p.ksplit * (size_t)p.ws_stride;
}

int svk_conv2d_wgrad9_tc(const svk_conv_desc* d, const void* x, const void* dy, float* ws, size_t ws_floats, int* ksplit_out,
                         cudaStream_t st) {
  Wgrad9P p{}; size_t smem;
  if (int e = plan9(d, &p, &smem)) return e;
  SVK_REQUIRE((size_t)p.ksplit * (size_t)p.ws_stride <= ws_floats, SVK_E_BADARG, "conv2d_wgrad9: workspace too small");
  p.ws = ws;
  p.prof = svk_prof_buffer();
  const int C = d->Cin;
  CUtensorMap tdy, tx;
  if (int e = make_nhwc_map(&tdy, dy, d->N, d->Ho, d->Wo, C, C, p.PW, p.bh + 2, 1)) return e;
  if (int e = make_nhwc_map(&tx, x, d->N, d->H, d->W, C, C, p.PW, p.bh, 1)) return e;
  if (C == 64) {
    static bool cfg = false;
    if (!cfg) { cudaError_t e = cudaFuncSetAttribute(conv_tc_wgrad9_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, W9_SMEM_MAX);
      SVK_REQUIRE(e == cudaSuccess, (int)e, "conv_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); cfg = true; }
    svk_launch(conv_tc_wgrad9_kernel<64>, p.ksplit, TC_THREADS, smem, st, tdy, tx, p);
  } else {
    static bool cfg = false;
    if (!cfg) { cudaError_t e = cudaFuncSetAttribute(conv_tc_wgrad9_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, W9_SMEM_MAX);
      SVK_REQUIRE(e == cudaSuccess, (int)e, "conv_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); cfg = true; }
    svk_launch(conv_tc_wgrad9_kernel<32>, p.ksplit, TC_THREADS, smem, st, tdy, tx, p);
  }
  SVK_LAUNCH_CHECK("conv_tc_wgrad9");
  *ksplit_out = p.ksplit;
  return 0;
}
