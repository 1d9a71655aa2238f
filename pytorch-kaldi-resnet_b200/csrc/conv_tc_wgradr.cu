// Weight gradient of 3x3 / stride-1 convolutions with Cin, Cout multiples of 128 (stages 3-4 of ResNet-34, every layer of
// the wide variant): full-image-row tiles, the three filter columns stacked along N through shifted operand views.
//
//     dw[co, (r,s), ci] = sum_{h,w}  dy[h, w, co] * x[h + r - 1, w + s - 1, ci]
//
// The late stages have tiny images (10x50, 5x25): pixel tiles whose width must be a multiple of 8 (the aligned-shift
// scheme of conv_tc.cu) waste up to half of every K step on padding.  A UMMA descriptor may start at any 16-byte-aligned
// address of a swizzled tile (tests/umma_shift_test.cu), so here
//   * a tile is bh full image rows of PW = W + 1 columns (the extra column is zero-filled by TMA); K runs over its
//     bh*PW smem rows, 16 per instruction, rounded up into a zeroed tail;
//   * a CTA owns one (128 Cout, filter row r, 128 Cin) block: dy is read unshifted (operand A, two 64-channel MN-atoms),
//     x is loaded from image row h0 + r - 1, column -1 (operand B); the taps s = 0,1,2 are three MN-atoms of the SAME x
//     tile ONE pixel apart (LBO = 128 bytes), stacked along N = 192: 2 UMMAs (one per 64-channel chunk of Cin) per 16 pixels
//     cover 3 taps at full tensor-pipe efficiency (M = 128, N = 192).
//   Where a shifted view wraps into the next image row it meets a zero of the other operand (dy column W, x column -1).
// Split-K over tiles; partials go to the workspace with plain stores, reduced by wgrad_reduce_*_kernel (conv_simt.cu).
#include "tc_common.cuh"

namespace {

struct WgradRP {
  int bh, PW, rows, ksteps;    // tile: bh image rows x PW columns = rows smem rows; ksteps = ceil(rows / 16)
  int tiles_h, num_tiles;      // tiles per image, N * tiles_h
  int Cin, Cout, n_n_blk;      // n_n_blk = Cin / 128
  int ksplit, tiles_per;
  int chunk_bytes;             // one 64-channel chunk of a tile in smem: (rows + 24 zero rows) * 128, 1024-aligned
  int tile_bytes;              // rows * 128: what one TMA box delivers
  int n_stages;                // a stage = dy chunk 0, dy chunk 1, x chunk 0, x chunk 1
  float* ws;                   // [ksplit][9][Cout][Cin]
  long long ws_stride;
  unsigned long long* prof;    // SVK_PROF=1 cycle counters (tc_common.cuh), else NULL
};

__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_wgradr_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX,
                      const __grid_constant__ WgradRP p) {
  constexpr int ROWB = 128;
  constexpr uint32_t LAYOUT = 2u;                     // SWIZZLE_128B
  constexpr uint32_t SBO = 8 * ROWB;
  constexpr int ACC_COLS = 192;                       // one accumulator = 128 Cout x (3 taps x 64 Cin)
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t stage_bytes = 4u * (uint32_t)p.chunk_bytes;
  const uint32_t auxoff = (uint32_t)p.n_stages * stage_bytes;
  const uint32_t aux = base + auxoff;
  const uint32_t bar_full = aux, bar_empty = aux + 64, bar_done = aux + 128;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(gbase + auxoff + 144);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.n_stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // the last K step of a tile runs into the 16 rows behind it, and the shifted x views read up to 2 rows further: those
  // rows are never written by TMA; zero them once (0 * stale NaN bits would poison the accumulators)
  for (int c = 0; c < 4 * p.n_stages; ++c) {
    uint32_t* tail = reinterpret_cast<uint32_t*>(gbase + (size_t)c * p.chunk_bytes + p.tile_bytes);
    const int nwords = (p.chunk_bytes - p.tile_bytes) / 4;
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) tail[i] = 0u;
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

  int wi = blockIdx.x;
  const int ks = wi % p.ksplit; wi /= p.ksplit;
  const int r = wi % 3; wi /= 3;
  const int nb = wi % p.n_n_blk;
  const int mb = wi / p.n_n_blk;
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
        const uint32_t sb = base + (uint32_t)st * stage_bytes;
        mbar_wait_t(bar_empty + 8 * st, ph ^ 1u, prof, pw);
        mbar_expect_tx(bar_full + 8 * st, 4u * (uint32_t)p.tile_bytes);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          tma_load_4d(sb + c * p.chunk_bytes, &tmDy, bar_full + 8 * st, mb * 128 + c * 64, 0, h0, n);
          tma_load_4d(sb + (2 + c) * p.chunk_bytes, &tmX, bar_full + 8 * st, nb * 128 + c * 64, -1, h0 + r - 1, n);
        }
        if (++st == p.n_stages) { st = 0; ph ^= 1u; }
      }
      if (prof) prof_flush(p.prof, 4, clock64() - pt0, pw, lane);
    }
  } else if (warp == 1) {
    {   // the WHOLE warp runs this loop; one elected lane issues
      constexpr uint32_t idesc = make_idesc(128, ACC_COLS, 1, 1);   // both operands MN-major
      int st = 0; uint32_t ph = 0;
      const uint64_t xstep = (uint64_t)((uint32_t)p.chunk_bytes >> 4);
      const bool prof = p.prof != nullptr;
      long long pwf = 0; const long long pt0 = prof ? clock64() : 0;
      unsigned long long gt0 = 0;
      if (prof) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0));
      for (int tile = t_beg; tile < t_end; ++tile) {
        mbar_wait_t(bar_full + 8 * st, ph, prof, pwf);
        tc_fence_after();
        const uint32_t sb = base + (uint32_t)st * stage_bytes;
        uint64_t ad = make_desc(sb, (uint32_t)p.chunk_bytes, SBO, LAYOUT);            // M = 2 atoms of 64 Cout
        uint64_t bd = make_desc(sb + 2u * p.chunk_bytes, ROWB, SBO, LAYOUT);          // N = 3 views one pixel apart
        uint32_t acc_flag = tile != t_beg ? 1u : 0u;
        for (int k = 0; k < p.ksteps; ++k) {
          tc_mma(tmem_base, ad, bd, idesc, acc_flag);
          tc_mma(tmem_base + ACC_COLS, ad, bd + xstep, idesc, acc_flag);
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
    const int co0 = mb * 128 + q * 32;                  // first Cout row of this warp
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    float* wsl = p.ws + (long long)ks * p.ws_stride;
    float* scr = reinterpret_cast<float*>(gbase) + q * 1024;      // the operand stages are idle now: 4 KB of scratch per warp
    for (int c = 0; c < 2; ++c) {
      for (int s = 0; s < 3; ++s) {
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          tc_ld32(taddr + (uint32_t)(c * ACC_COLS + s * 64 + h * 32), v);
          store_chunk_rows(scr, v, wsl + ((long long)(r * 3 + s) * p.Cout + co0) * p.Cin + nb * 128 + c * 64 + h * 32, p.Cin, lane);
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

constexpr int WR_SMEM_MAX = 220 * 1024;

int planr(const svk_conv_desc* d, WgradRP* pp, size_t* smem_out) {
  WgradRP& p = *pp;
  p.Cin = d->Cin; p.Cout = d->Cout; p.n_n_blk = d->Cin / 128;
  p.PW = d->W + 1;
  SVK_REQUIRE(p.PW <= 256, SVK_E_UNSUPPORTED, "conv2d_wgradr: W=%d too wide", d->W);
  // bh: least MMA time per image — K steps (with their round-up padding) x 192 cycles + ~600 cycles of per-tile barrier /
  // descriptor overhead (one-row tiles of a 5x25 image were issue-bound) — among the tiles that leave room for >= 3 stages,
  // else >= 2
  long long best = -1; int best_ns = 0;
  for (int want = 3; want >= 2 && best < 0; --want) {
    for (int bh = 1; bh <= d->H && bh <= 256; ++bh) {
      const int rows = bh * p.PW;
      const int chunk = ((rows + 24) * 128 + 1023) / 1024 * 1024;   // K round-up (<= 15 rows) + 2 rows of view shift
      const int ns = (WR_SMEM_MAX - 2048) / (4 * chunk);
      if (ns < want) break;
      const long long cost = (long long)((d->H + bh - 1) / bh) * (((rows + 15) / 16) * 192 + 600);
      if (best < 0 || cost < best) { best = cost; p.bh = bh; p.rows = rows; p.chunk_bytes = chunk; best_ns = ns; }
    }
  }
  SVK_REQUIRE(best >= 0, SVK_E_UNSUPPORTED, "conv2d_wgradr: no tile for %dx%d", d->H, d->W);
  p.n_stages = best_ns > 6 ? 6 : best_ns;
  p.ksteps = (p.rows + 15) / 16;
  p.tile_bytes = p.rows * 128;
  p.tiles_h = (d->H + p.bh - 1) / p.bh;
  p.num_tiles = d->N * p.tiles_h;
  const int items = (d->Cout / 128) * p.n_n_blk * 3;
  int ks = svk_num_sms() / items;
  if (ks < 1) ks = 1;
  if (ks > p.num_tiles) ks = p.num_tiles;
  p.tiles_per = (p.num_tiles + ks - 1) / ks;
  p.ksplit = (p.num_tiles + p.tiles_per - 1) / p.tiles_per;
  p.ws_stride = (long long)9 * d->Cout * d->Cin;
  *smem_out = (size_t)p.n_stages * 4 * p.chunk_bytes + 2048;
  return 0;
}

}  // namespace

bool svk_wgradr_applicable(const svk_conv_desc* d) {
  if (!(d->R == 3 && d->stride == 1 && d->Cin % 128 == 0 && d->Cout % 128 == 0)) return false;
  WgradRP p{}; size_t smem;
  return planr(d, &p, &smem) == 0;
}

size_t svk_conv2d_wgradr_tc_ws_floats(const svk_conv_desc* d) {
  WgradRP p{}; size_t smem;
  if (planr(d, &p, &smem)) return 0;
  return (size_t)p.ksplit * (size_t)p.ws_stride;
}

int svk_conv2d_wgradr_tc(const svk_conv_desc* d, const void* x, const void* dy, float* ws, size_t ws_floats, int* ksplit_out,
                         cudaStream_t st) {
  WgradRP p{}; size_t smem;
  if (int e = planr(d, &p, &smem)) return e;
  SVK_REQUIRE((size_t)p.ksplit * (size_t)p.ws_stride <= ws_floats, SVK_E_BADARG, "conv2d_wgradr: workspace too small");
  p.ws = ws;
  p.prof = svk_prof_buffer();
  CUtensorMap tdy, tx;
  if (int e = make_nhwc_map(&tdy, dy, d->N, d->Ho, d->Wo, d->Cout, 64, p.PW, p.bh, 1)) return e;
  if (int e = make_nhwc_map(&tx, x, d->N, d->H, d->W, d->Cin, 64, p.PW, p.bh, 1)) return e;
  static bool cfg = false;
  if (!cfg) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_wgradr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WR_SMEM_MAX);
    SVK_REQUIRE(e == cudaSuccess, (int)e, "conv_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    cfg = true;
  }
  const int grid = (d->Cout / 128) * p.n_n_blk * 3 * p.ksplit;
  svk_launch(conv_tc_wgradr_kernel, grid, TC_THREADS, smem, st, tdy, tx, p);
  SVK_LAUNCH_CHECK("conv_tc_wgradr");
  *ksplit_out = p.ksplit;
  return 0;
}
