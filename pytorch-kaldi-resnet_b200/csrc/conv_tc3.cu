// 3x3 / stride-1 fprop and dgrad for the SMALL-channel stages (Cout <= 64: stages 1-2 of ResNet-34, 45 % of the
// conv FLOPs), where the generic per-tap pipeline of conv_tc.cu is bound by per-stage latency and per-SM L2 ingest
// rather than by the tensor pipe (profiles/r01_conv_stage1.md).  Two changes of dataflow:
//   * the whole packed filter (9 * Cin * Cout bf16 <= 96 KB) is loaded into shared memory ONCE per persistent CTA;
//   * one TMA load fetches a (bh+2) x bw HALO tile for a filter column s; the three taps r = 0,1,2 of that column are
//     three UMMA A-descriptors into the same tile, offset by r*bw rows (bw % 8 == 0 keeps every offset a whole number of
//     swizzle atoms).  Per output tile: 3 loads and 0 weight loads instead of 9 + 9.
// Everything else (TMEM double buffering, epilogue, statistics) is shared with conv_tc.cu.
#include "tc_common.cuh"
#include "epilogue_v2.cuh"

namespace {

struct Gather3P {
  GatherP g;
  int col_dw[3];       // W offset of halo load l (filter column s = l): l-1 for fprop, 1-l for dgrad
  int a_stage_bytes;   // halo tile bytes rounded up to 1024
  int halo_bytes;      // bytes one halo TMA load delivers
  int n_stages;
  int single;          // 1: ONE (bh+2) x (bw+2) halo tile per output tile; tap (r, s) is the A descriptor offset r' * pitch + s'
                       //    (descriptors may start at any row: tests/umma_shift_test.cu), accumulator rows i * pitch + j with
                       //    j >= bw discarded.  0: one (bh+2) x bw halo tile per filter column (bw % 8 == 0)
  // staged epilogue (epilogue_v2.cuh): operand tiles of the epilogue arrive by TMA, the output leaves by TMA
  int n_acc;           // accumulator buffers in TMEM = epilogue warp groups (2: 320 threads, 3: 448 threads)
  int n_mma;           // MMA-issuing warps: 1, or 2 (one more warp at the end of the CTA; the two alternate tiles)
  int n_tm;            // accumulator buffers in TMEM (GatherP::n_tm): n_acc, or 2 * n_acc when n_mma == 2 and n_acc is odd
  int epi2;            // 0: first epilogue (tc_common.cuh)
  int aux_nbuf;        // aux buffers per epilogue group (2 when they fit)
  int aux_slots;       // tile slots per aux buffer: 1 (forward: staging only), 2 (mask, c), 3 (mask, c, shortcut gradient)
  int aux_box_bytes;   // bytes one operand-tile TMA load delivers (bh * bw * BN * 2)
  int coef_off, aux_off;   // smem offsets (from the 1024-aligned base) of the coefficient table and the aux region
};

// Kc == KC (one K chunk per tap: Cin in {32, 64}).  Tap (r = t, s = l) reads halo rows shifted by t (fprop) or 2 - t (dgrad)
// and the packed filter tap t*3 + l.
template <int KC, int BN, bool DGRAD>
__global__ void __launch_bounds__(GATHER3_THREADS + 32, 1)
conv_tc_gather3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmMask, const __grid_constant__ CUtensorMap tmC,
                       const __grid_constant__ CUtensorMap tmRes, const __grid_constant__ CUtensorMap tmOut,
                       const __grid_constant__ Gather3P q) {
  constexpr int ROWB = KC * 2;
  constexpr int B_BYTES = BN * ROWB;
  constexpr uint32_t LAYOUT = (KC == 64) ? 2u : 4u;
  constexpr uint32_t SBO = 8 * ROWB;
  constexpr int TMEM_COLS = 8 * BN;          // room for up to 8 accumulator buffers (q.n_tm of them are used; one CTA per SM)
  pdl_launch_dependents();
  const GatherP& p = q.g;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t a_bytes_all = (uint32_t)q.n_stages * q.a_stage_bytes;
  const uint32_t w_bytes_all = 9u * B_BYTES;
  const uint32_t stage0 = base;
  const uint32_t wsm = base + a_bytes_all;
  const uint32_t auxoff = a_bytes_all + w_bytes_all;
  const uint32_t aux = base + auxoff;
  // aux: wfull @160, tmem ptr @176, afull[8] @192, full[16] @320, empty[16] @448, tfull[8] @576, tempty[8] @640
  const uint32_t bar_full = aux + 320, bar_empty = aux + 448, bar_tfull = aux + 576, bar_tempty = aux + 640, bar_w = aux + 160;
  const uint32_t bar_afull = aux + 192;                               // staged epilogue: operand tiles of group g, buffer b at 2 g + b
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(gbase + auxoff + 176);
  float* scr = reinterpret_cast<float*>(gbase + auxoff + SMEM_AUX);
  float* coef = reinterpret_cast<float*>(gbase + q.coef_off);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // arrivals that free an accumulator buffer: one per epilogue warp draining it (8 when two groups split every tile)
  const uint32_t tempty_count = (p.bn_mask && BN >= 128 && blockDim.x == GATHER_THREADS) ? 8u : 4u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < q.n_stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int a = 0; a < 8; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, tempty_count); }
    mbar_init(bar_w, 1);
    for (int a = 0; a < 8; ++a) mbar_init(bar_afull + 8 * a, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_wait();             // everything above overlapped the previous kernel's tail; global memory is touched from here on
  if (p.scale) {
    for (int i = threadIdx.x; i < p.Nout; i += blockDim.x) { coef[i] = p.scale[i]; coef[512 + i] = p.shift[i]; }
  }
  if (p.bn_c) {
    for (int i = threadIdx.x; i < p.Nout; i += blockDim.x) { coef[i] = p.bn_mean[i]; coef[512 + i] = p.bn_rstd[i]; }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    {   // the WHOLE warp runs this loop (warp-uniform control flow lets ptxas keep descriptors and coordinates in
        // uniform registers); the issuing wrappers elect one lane
      mbar_expect_tx(bar_w, w_bytes_all);
      for (int t = 0; t < 9; ++t) tma_load_2d(wsm + t * B_BYTES, &tmB, bar_w, 0, t * p.Nout);
      int stage = 0; uint32_t ph = 0;
      const bool prof = p.prof != nullptr;
      long long pw = 0; const long long pt0 = prof ? clock64() : 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int pt = tile;
        const int tw = pt % p.tiles_w; pt /= p.tiles_w;
        const int th = pt % p.tiles_h;
        const int n = pt / p.tiles_h;
        const int h0 = th * p.bh, w0 = tw * p.bw;
        if (q.single) {
          mbar_wait_t(bar_empty + 8 * stage, ph ^ 1u, prof, pw);
          mbar_expect_tx(bar_full + 8 * stage, (uint32_t)q.halo_bytes);
          tma_load_4d(stage0 + stage * q.a_stage_bytes, &tmA, bar_full + 8 * stage, 0, w0 - 1, h0 - 1, n);
          if (++stage == q.n_stages) { stage = 0; ph ^= 1u; }
          continue;
        }
#pragma unroll
        for (int l = 0; l < 3; ++l) {
          mbar_wait_t(bar_empty + 8 * stage, ph ^ 1u, prof, pw);
#if SVK_DBG_NO_TMA      // diagnostic build: the stage is declared full without loading it
          if (elect_one()) mbar_arrive(bar_full + 8 * stage);
#else
          mbar_expect_tx(bar_full + 8 * stage, (uint32_t)q.halo_bytes);
          tma_load_4d(stage0 + stage * q.a_stage_bytes, &tmA, bar_full + 8 * stage, 0, w0 + (DGRAD ? 1 - l : l - 1), h0 - 1, n);
#endif
          if (++stage == q.n_stages) { stage = 0; ph ^= 1u; }
        }
      }
      if (prof) prof_flush(p.prof, 4, clock64() - pt0, pw, lane);
    }
  } else if (warp == 1 || (q.n_mma == 2 && warp == (int)(blockDim.x >> 5) - 1)) {
    // ===================== MMA issuer(s) =====================
    {   // the WHOLE warp runs this loop (warp-uniform control flow lets ptxas keep descriptors and coordinates in
        // uniform registers); the issuing wrappers elect one lane
      // The issuing thread is instruction-latency bound (profiles/r01: ~180 cycles per UMMA against a 64-cycle tensor
      // floor when descriptors are rebuilt per MMA), so everything is hoisted: a descriptor is base + (byte offset >> 4)
      // in its low 14-bit address field, and all tap offsets are compile-time multiples of loop-invariant registers.
      // Even so one warp needs ~18 SASS instructions (~80 cycles) per UMMA, twice what an N = 32 instruction occupies the
      // tensor pipe for (tests/bench_umma.cu: 40 cycles), and no background traffic changes that (profiles/r02_conv_notes.md);
      // each of a tile's four barrier polls adds ~100 cycles even when its phase completed long ago (same micro-benchmark).
      // Polling ahead of need (mbarrier.test_wait before the previous stage's MMAs) was tried: its bookkeeping cost more.
      // With q.n_mma == 2 a SECOND issuing warp (the CTA's last) takes every other tile of the CTA: its own accumulator
      // buffer(s), its own positions in the stage ring; tcgen05.commit only tracks the MMAs of the issuing thread.
      constexpr uint32_t idesc = make_idesc(128, BN, 0, 0);
      const int mma_id = (warp == 1) ? 0 : 1;
      const int loads_per_tile = q.single ? 1 : 3;
      mbar_wait(bar_w, 0);
      tc_fence_after();
      int stage = 0; uint32_t ph = 0;
      int acc = 0; uint32_t aph = 0;
      // The stage-ring / accumulator positions move on by one tile that the OTHER issuing warp handles.  Ownership keeps the
      // parity waits unambiguous: the host makes n_stages a multiple of 2 * loads_per_tile and the number of accumulator
      // buffers n_tm even, so a stage / an accumulator is only ever used by ONE of the two warps.  (With a buffer shared by
      // both, a warp that runs ahead — or falls behind — waits for a phase two away from the barrier's, which
      // try_wait.parity cannot tell from the right one: corrupted tiles and stalled pipelines in tests/stress_conv.py.)
      auto skip_tile = [&]() {
        stage += loads_per_tile;
        while (stage >= q.n_stages) { stage -= q.n_stages; ph ^= 1u; }
        if (++acc == q.n_tm) { acc = 0; aph ^= 1u; }
      };
      if (mma_id) skip_tile();
      const uint64_t row16 = (uint64_t)(((uint32_t)p.bw * ROWB) >> 4);      // one halo image row, in 16-byte units
      const uint64_t a_desc0 = make_desc(stage0, 16, SBO, LAYOUT);
      const uint64_t a_stage16 = (uint64_t)((uint32_t)q.a_stage_bytes >> 4);
      const uint64_t b_desc0 = make_desc(wsm, 16, SBO, LAYOUT);
      const bool prof = p.prof != nullptr && mma_id == 0;
      long long pwf = 0, pwt = 0; const long long pt0 = prof ? clock64() : 0;
      unsigned long long gt0 = 0;
      if (prof) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0));
      for (int tile = blockIdx.x + mma_id * gridDim.x; tile < p.total_tiles; tile += q.n_mma * gridDim.x) {
        mbar_wait_t(bar_tempty + 8 * acc, aph ^ 1u, prof, pwt);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        if (q.single) {
          // one halo tile: tap (r = t, s = l) reads rows shifted by r' * pitch + s' (fprop r' = t, s' = l; dgrad 2 - t, 2 - l)
          mbar_wait_t(bar_full + 8 * stage, ph, prof, pwf);
          tc_fence_after();
          const uint64_t a_desc = a_desc0 + (uint64_t)(uint32_t)stage * a_stage16;
          const uint64_t pix16 = (uint64_t)(ROWB >> 4);                               // one pixel row, in 16-byte units
          const uint64_t prow16 = (uint64_t)(((uint32_t)p.pitch * ROWB) >> 4);        // one halo image row
#pragma unroll
          for (int l = 0; l < 3; ++l) {
#pragma unroll
            for (int t = 0; t < 3; ++t) {
              const uint64_t ad_t = a_desc + (uint64_t)(DGRAD ? 2 - t : t) * prow16 + (uint64_t)(DGRAD ? 2 - l : l) * pix16;
              const uint64_t bd_t = b_desc0 + (uint64_t)(((t * 3 + l) * B_BYTES) >> 4);
#pragma unroll
              for (int k = 0; k < KC / 16; ++k)
                tc_mma(d_tmem, ad_t + 2 * k, bd_t + 2 * k, idesc, (l | t | k) != 0 ? 1u : 0u);
            }
          }
          tc_commit(bar_empty + 8 * stage);
          if (++stage == q.n_stages) { stage = 0; ph ^= 1u; }
        } else {
#pragma unroll
          for (int l = 0; l < 3; ++l) {
            mbar_wait_t(bar_full + 8 * stage, ph, prof, pwf);
            tc_fence_after();
            const uint64_t a_desc = a_desc0 + (uint64_t)(uint32_t)stage * a_stage16;
#pragma unroll
            for (int t = 0; t < 3; ++t) {
              const uint64_t ad_t = a_desc + (uint64_t)(DGRAD ? 2 - t : t) * row16;
              const uint64_t bd_t = b_desc0 + (uint64_t)(((t * 3 + l) * B_BYTES) >> 4);
#pragma unroll
              for (int k = 0; k < KC / 16; ++k)
                tc_mma(d_tmem, ad_t + 2 * k, bd_t + 2 * k, idesc, (l | t | k) != 0 ? 1u : 0u);
            }
            tc_commit(bar_empty + 8 * stage);
            if (++stage == q.n_stages) { stage = 0; ph ^= 1u; }
          }
        }
        tc_commit(bar_tfull + 8 * acc);
        if (++acc == q.n_tm) { acc = 0; aph ^= 1u; }
        if (q.n_mma == 2) skip_tile();
      }
      if (prof && lane == 0) {
        atomicAdd(p.prof + 0, 1ull); atomicAdd(p.prof + 1, (unsigned long long)(clock64() - pt0));
        atomicAdd(p.prof + 2, (unsigned long long)pwf); atomicAdd(p.prof + 3, (unsigned long long)pwt);
        unsigned long long gt1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
        atomicAdd(p.prof + 8, gt1 - gt0); atomicMax(p.prof + 9, ~gt0); atomicMax(p.prof + 10, gt1);   // ns; [9] = ~(earliest start)
        atomicMax(p.prof + 11, gt0); atomicMax(p.prof + 12, ~gt1);                                    // latest start, ~(earliest end)
      }
    }
  } else if (q.epi2) {
    gather_epilogue_v2<BN>(p, &tmOut, &tmMask, &tmC, &tmRes, tmem_base, bar_tfull, bar_tempty, bar_afull, gbase + q.aux_off,
                           base + (uint32_t)q.aux_off, q.aux_slots, q.aux_nbuf, q.aux_box_bytes, coef, warp, lane);
  } else {
    if (p.bn_mask) {
      if (BN >= 128 && blockDim.x == GATHER_THREADS)
        gather_epilogue_bn<BN, (BN >= 128)>(p, tmem_base, bar_tfull, bar_tempty, scr, coef, warp, lane);
      else
        gather_epilogue_bn<BN, false>(p, tmem_base, bar_tfull, bar_tempty, scr, coef, warp, lane);
    }
    else gather_epilogue<BN>(p, tmem_base, bar_tfull, bar_tempty, scr, coef, warp, lane);
  }
  tc_fence_before();
  __syncthreads();
  if (p.stats && stats_use_mailbox(p)) {       // the CTA's statistics: one atomic per channel and kind (tc_common.cuh)
    if (q.epi2) gather_epilogue_v2_finish<BN>(p, gbase + q.aux_off, q.n_acc, q.aux_slots, q.aux_nbuf);
    else if (!p.bn_mask) stats_mailbox_finish(p, scr, 32 * 33, 4 * q.n_acc, BN);
    else if (p.bn_c) stats_mailbox_finish(p, scr, 32 * SCR_STRIDE, 4 * q.n_acc, BN);
  }
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

constexpr int G3_SMEM_MAX = 227 * 1024;

template <int KC, int BN, bool DGRAD>
int launch3(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap* tx, const Gather3P& q, size_t smem, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_gather3_kernel<KC, BN, DGRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, G3_SMEM_MAX);
    SVK_REQUIRE(e == cudaSuccess, (int)e, "conv_tc3: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    configured = true;
  }
  int grid = q.g.total_tiles < svk_num_sms() ? q.g.total_tiles : svk_num_sms();
  svk_launch(conv_tc_gather3_kernel<KC, BN, DGRAD>, grid, (q.n_acc == 3 ? GATHER3_THREADS : GATHER_THREADS) + (q.n_mma == 2 ? 32 : 0), smem, st, ta, tb, tx[0], tx[1],
             tx[2], tx[3], q);
  SVK_LAUNCH_CHECK("conv_tc_gather3");
  return 0;
}

}  // namespace

// Is the resident-filter halo kernel applicable?  (3x3, stride 1, Cout in {32, 64}, filter fits next to >= 2 stages)
bool svk_gather3_applicable(int R, int stride, int Kc, int Nout) {
  if (R != 3 || stride != 1) return false;
  if (!(Nout == 32 || Nout == 64)) return false;
  return Kc == 32 || Kc == 64;        // one K chunk per tap; the filter (<= 73.7 KB) stays resident in smem
}

// `in` = gathered tensor [N, Hc, Wc, Kc] (x for fprop, dy for dgrad); output grid has the same Hc x Wc (stride 1).
// dgrad != 0 mirrors the tap offsets (dx[h,w] gathers dy[h+1-r, w+1-s]).
int svk_conv3x3s1_gather3_tc(const void* in, int N, int Hc, int Wc, int Kc, const void* w_packed, int Nout, int dgrad,
                             GatherP p, cudaStream_t st) {
  const int KC = Kc;
  const int BN = Nout;
  const int ROWB = KC * 2;
  Gather3P q{};
  // tile: bw % 8 == 0, bh*bw <= 128; minimise tiles * max(tensor cycles, load cycles + fixed latency)
  double best = 1e30; int bbh = 0, bbw = 0;
  const double tensor_cyc = 9.0 * (Kc / 16) * 64.0;
  for (int bw = 8; bw <= 128; bw += 8) {
    for (int bh = 1; bh * bw <= 128 && bh <= 254; ++bh) {
      int th = (Hc + bh - 1) / bh, tw = (Wc + bw - 1) / bw;
      double load_cyc = 3.0 * (Kc / KC) * ((bh + 2) * bw * ROWB / 48.0 + 250.0);
      double cost = (double)th * tw * (tensor_cyc > load_cyc ? tensor_cyc : load_cyc);
      if (cost < best - 1e-9) { best = cost; bbh = bh; bbw = bw; }
    }
  }
  static int single_on = -1;
  if (single_on < 0) { const char* e = getenv("SVK_DISABLE_SINGLE_HALO"); single_on = (e && e[0] == '1') ? 0 : 1; }
  // measured inside the training step (per step, 13 launches each): 64 channels — fprop 0.626 -> 0.568 ms, fused dgrad
  // 0.767 -> 0.719 ms; 32 channels — fprop unchanged, fused dgrad 0.908 -> 1.328 ms (its per-thread epilogue accesses do
  // not like the narrower tiles), so the 32-channel stage keeps one halo tile per filter column.
  static int single_fwd32 = -1;      // A/B: single halo tile also for the 32-channel FORWARD convolution
  if (single_fwd32 < 0) { const char* e = getenv("SVK_SINGLE_HALO_FWD32"); single_fwd32 = (e && e[0] == '1') ? 1 : 0; }
  q.single = (single_on && (Nout == 64 || (single_fwd32 && Nout == 32 && !dgrad))) ? 1 : 0;
  p.pitch = 0;
  if (q.single) {
    // single-halo tiles: accumulator rows i * (bw + 2) + j; fewest tiles, then the smallest halo tile (bytes loaded per tile)
    long long best_t = -1; int best_halo = 0;
    for (int bw = 1; bw <= 126 && bw <= Wc; ++bw) {
      const int P = bw + 2;
      if ((bw + 2) > 256) break;
      int bh = (128 - bw) / P + 1;                       // (bh - 1) * P + bw <= 128
      if (bh > Hc) bh = Hc;
      if (bh + 2 > 256) bh = 254;
      const long long tiles = (long long)((Hc + bh - 1) / bh) * ((Wc + bw - 1) / bw);
      const int halo = (bh + 2) * P;
      if (best_t < 0 || tiles < best_t || (tiles == best_t && halo < best_halo)) { best_t = tiles; best_halo = halo; bbh = bh; bbw = bw; }
    }
    p.pitch = bbw + 2;
  }
  p.bh = bbh; p.bw = bbw;
  p.tiles_h = (Hc + p.bh - 1) / p.bh;
  p.tiles_w = (Wc + p.bw - 1) / p.bw;
  p.num_pix_tiles = N * p.tiles_h * p.tiles_w;
  p.n_blocks = 1;
  p.total_tiles = p.num_pix_tiles;
  p.kchunks = Kc / KC;
  p.Nout = Nout;
  p.in_mul = 1;
  p.Hc = Hc; p.Wc = Wc;
  p.prof = svk_prof_buffer();
  q.g = p;
  for (int l = 0; l < 3; ++l) q.col_dw[l] = dgrad ? 1 - l : l - 1;
  q.halo_bytes = (p.bh + 2) * (q.single ? p.pitch : p.bw) * ROWB;
  // the tap views read rows [shift, shift + 128): keep every stage large enough for the deepest view
  int need_rows = q.single ? 2 * p.pitch + 2 + 128 : 2 * p.bw + 128;
  int rows = (p.bh + 2) * (q.single ? p.pitch : p.bw);
  if (rows < need_rows) rows = need_rows;
  q.a_stage_bytes = (rows * ROWB + 1023) / 1024 * 1024;
  // staged epilogue: training forward (statistics, no scale / residual / ReLU) and the fused BatchNorm-backward dgrad
  static int epi2_on = -1;
  if (epi2_on < 0) { const char* e = getenv("SVK_DISABLE_EPI2"); epi2_on = (e && e[0] == '1') ? 0 : 1; }
  const bool bwd_mode = dgrad && p.bn_mask && !p.res_m && !p.scale && !p.valid_w && !p.relu;
  // Measured inside the training step (profiles/r02_epilogue_v2.md): the staged epilogue wins where the first one is bound
  // by L1 wavefronts — the fused data gradient of the 64-channel stage (0.95 -> 0.74 ms per step) — and loses where a tile's
  // MMA time (32 channels: ~1,200 cycles) is shorter than its fixed barrier / store latencies, and in the forward pass.
  // SVK_EPI2_MODE (A/B): bit 0 = also the 32-channel fused dgrad, bit 1 = the 64-channel training forward, bit 2 = the
  // 32-channel training forward
  static int epi2_mode = -1;
  if (epi2_mode < 0) { const char* e = getenv("SVK_EPI2_MODE"); epi2_mode = e ? atoi(e) : 0; }
  const bool fwd_mode = !dgrad && p.stats && !p.bn_mask && !p.res && !p.res_m && !p.scale && !p.valid_w && !p.relu;
  q.epi2 = (epi2_on && bwd_mode && (Nout == 64 || (epi2_mode & 1))) ? 1 : 0;
  if (epi2_on && fwd_mode && ((Nout == 64 && (epi2_mode & 2)) || (Nout == 32 && (epi2_mode & 4)))) q.epi2 = 1;
  const size_t w_bytes = (size_t)9 * Kc * Nout * 2;
  // Issuing warps / epilogue groups / operand stages.  One warp issues a UMMA every ~80 cycles (about 18 SASS instructions
  // each), twice what an N = 32 instruction occupies the tensor pipe for (tests/bench_umma.cu: 40 cycles under any background
  // shared-memory traffic), so a second issuing warp taking every other tile doubles the issue rate; the loop then waits on
  // the epilogue (three groups instead of two) and on operands (TMA latency under load is ~2 tiles of MMA time: 12 stages
  // instead of 6).  Measured inside the training step, per step (profiles/r02_conv_notes.md): forward 32 channels 0.684 ->
  // 0.594 ms, 64 channels 0.545 -> 0.533 ms; the fused data gradients are bound by their four or five operand streams
  // (0.90 -> 0.93 ms for 32 channels with the same change, 0.71 -> 0.74 ms for 64), so they keep one warp / two groups / 6.
  // SVK_GATHER3_MMA / _GROUPS / _STAGES override the choice for every launch (A/B, tests/stress_conv.py).
  static int groups_env = -1, mma_env = -1, stages_env = -1;
  if (groups_env < 0) { const char* e = getenv("SVK_GATHER3_GROUPS"); groups_env = (e && (e[0] == '2' || e[0] == '3')) ? e[0] - '0' : 0; }
  if (mma_env < 0) { const char* e = getenv("SVK_GATHER3_MMA"); mma_env = (e && (e[0] == '1' || e[0] == '2')) ? e[0] - '0' : 0; }
  if (stages_env < 0) { const char* e = getenv("SVK_GATHER3_STAGES"); stages_env = e ? atoi(e) : 0; if (stages_env > 16) stages_env = 16; }
  const bool wide = !dgrad;
  const int groups_sel = groups_env ? groups_env : (wide ? 3 : 2);
  q.n_acc = q.epi2 ? 2 : groups_sel;
  q.n_mma = mma_env ? mma_env : (wide ? 2 : 1);
  const size_t scr_bytes = q.n_acc == 3 ? SCR3_BYTES : SCR_BYTES;
  size_t fixed = w_bytes + SMEM_AUX + scr_bytes + COEF_BYTES + 1024;
  size_t limit = q.n_acc == 3 ? 220 * 1024 : 200 * 1024;
  q.coef_off = 0; q.aux_off = 0;      // filled in below (they depend on the stage count)
  if (q.epi2) {
    q.aux_slots = p.res ? 3 : (p.bn_c ? 2 : 1);
    q.aux_box_bytes = p.bh * p.bw * Nout * 2;
    const size_t buf_bytes = (size_t)q.aux_slots * 128 * Nout * 2;
    const size_t fixed0 = w_bytes + SMEM_AUX + COEF_BYTES + 1024 /* align aux */ + 1024 /* align base */;
    // epilogue groups x aux buffers per group, in order of preference; each needs room for `want` operand stages next to it
    static int epi2_groups = -1;
    if (epi2_groups < 0) { const char* e = getenv("SVK_EPI2_GROUPS"); epi2_groups = (e && e[0] == '3') ? 3 : 2; }
    const int want = q.single ? 3 : 6;
    bool placed = false;
    for (int g = epi2_groups; g >= 2 && !placed; --g) {
      for (int nb = 2; nb >= 1 && !placed; --nb) {
        const size_t f = fixed0 + (size_t)g * nb * buf_bytes;
        if (f < (size_t)G3_SMEM_MAX && (int)((G3_SMEM_MAX - f) / q.a_stage_bytes) >= (nb == 2 ? want : (g == 3 ? 3 : 2))) {
          q.n_acc = g; q.aux_nbuf = nb; fixed = f; limit = G3_SMEM_MAX; placed = true;
        }
      }
    }
    if (!placed) {
      q.epi2 = 0;                     // not enough room next to the resident filter: first epilogue
      q.n_acc = groups_sel;
      fixed = w_bytes + SMEM_AUX + (q.n_acc == 3 ? SCR3_BYTES : SCR_BYTES) + COEF_BYTES + 1024;
      limit = q.n_acc == 3 ? 220 * 1024 : 200 * 1024;
    }
  }
  int ns = (int)((limit - fixed) / q.a_stage_bytes);
  const int ns_cap = stages_env >= 2 ? stages_env : (wide ? 12 : 6);
  if (ns > ns_cap) ns = ns_cap;
  if (q.n_mma == 2) {                // every stage belongs to one issuing warp (see the kernel's skip_tile)
    const int unit = 2 * (q.single ? 1 : 3);
    if (ns >= unit) ns = ns / unit * unit; else q.n_mma = 1;
  }
  SVK_REQUIRE(ns >= 2, SVK_E_UNSUPPORTED, "conv_tc3: not enough shared memory for 2 stages");
  q.n_stages = ns;
  q.n_tm = (q.n_mma == 2 && (q.n_acc & 1)) ? 2 * q.n_acc : q.n_acc;
  q.g.n_tm = q.n_tm;
  // what the kernel's barrier protocol relies on (see its skip_tile): fail here, loudly, rather than corrupt tiles there
  SVK_REQUIRE(q.n_mma == 1 || ((q.n_tm & 1) == 0 && ns % (2 * (q.single ? 1 : 3)) == 0), SVK_E_UNSUPPORTED,
              "conv_tc3: two issuing warps need an even number of accumulator buffers (%d) and whole tiles of stages per warp (%d)",
              q.n_tm, ns);
  SVK_REQUIRE(q.n_tm <= 8 && q.n_tm * BN <= 512 && ns <= 16 && q.n_tm % q.n_acc == 0, SVK_E_UNSUPPORTED,
              "conv_tc3: %d accumulator buffers of %d columns / %d stages exceed the kernel's TMEM or barrier tables", q.n_tm, BN, ns);
  const size_t smem = fixed + (size_t)ns * q.a_stage_bytes;
  {
    const size_t auxbar = (size_t)ns * q.a_stage_bytes + w_bytes;
    q.coef_off = (int)(auxbar + SMEM_AUX + (q.epi2 ? 0 : (q.n_acc == 3 ? SCR3_BYTES : SCR_BYTES)));
    q.aux_off = (int)((q.coef_off + COEF_BYTES + 1023) / 1024 * 1024);
  }
  CUtensorMap ta, tb, tx[4];
  if (int e = make_nhwc_map(&ta, in, N, Hc, Wc, Kc, KC, q.single ? p.pitch : p.bw, p.bh + 2, 1)) return e;
  if (int e = make_w_map(&tb, w_packed, (long long)9 * Nout, Kc, KC, BN)) return e;
  for (int k = 0; k < 4; ++k) tx[k] = ta;            // placeholders for the unused maps
  if (q.epi2) {
    if (p.bn_mask) { if (int e = make_nhwc_map(&tx[0], p.bn_mask, N, Hc, Wc, Nout, Nout, p.bw, p.bh, 1)) return e; }
    if (p.bn_c) { if (int e = make_nhwc_map(&tx[1], p.bn_c, N, Hc, Wc, Nout, Nout, p.bw, p.bh, 1)) return e; }
    if (p.res) { if (int e = make_nhwc_map(&tx[2], p.res, N, Hc, Wc, Nout, Nout, p.bw, p.bh, 1)) return e; }
    if (int e = make_nhwc_map(&tx[3], p.out, N, Hc, Wc, Nout, Nout, p.bw, p.bh, 1)) return e;
  }
#define SVK_L3(K_, N_) if (KC == K_ && BN == N_) return dgrad ? launch3<K_, N_, true>(ta, tb, tx, q, smem, st) : launch3<K_, N_, false>(ta, tb, tx, q, smem, st)
  SVK_L3(32, 32); SVK_L3(32, 64); SVK_L3(64, 32); SVK_L3(64, 64);
#undef SVK_L3
  svk_set_error("conv_tc3: no kernel for KC=%d BN=%d", KC, BN);
  return SVK_E_UNSUPPORTED;
}
