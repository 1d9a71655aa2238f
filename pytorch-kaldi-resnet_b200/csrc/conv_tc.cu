// bf16 implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05 + TMEM), operands staged by TMA.
//
//   fprop / dgrad  ("gather GEMM"):  D[128 pixels x BN channels] += A[128 pixels x KC] * B[BN x KC]^T   per (tap, k-chunk)
//       A = a bh x bw box of NHWC pixels shifted by the tap offset, fetched by ONE 4-D TMA box; out-of-image
//           coordinates are zero-filled by TMA (= the conv padding), stride-2 uses TMA element strides.
//       B = packed weights [tap][Nout][Kc] (K-major), one 2-D TMA box.
//       Both land in 128B/64B-swizzled K-major smem and feed tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16)
//       with fp32 accumulators in TMEM (double-buffered so the epilogue of tile i overlaps the MMAs of tile i+1).
//       Epilogue (4 warps): tcgen05.ld -> scale/shift -> +residual -> ReLU -> bf16 -> global, plus per-channel
//       sum / sum-of-squares of the stored values (BatchNorm batch statistics) reduced through smem.
//   wgrad:  D_tap[Cout x Cin_blk] += dy_tile^T[Cout x P] * x_tap_tile[P x Cin_blk], P = pixels of a tile (K dim).
//       Both operands are "MN-major" (channels contiguous): the same NHWC TMA boxes, UMMA descriptors with the
//       transpose bits set.  One TMEM accumulator region per filter tap; split-K over pixel tiles across CTAs, fp32
//       atomics into the packed gradient.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2-5 = epilogue.
// Every mbarrier wait is bounded (trap after ~4 s) so a protocol bug cannot hang the GPU.
#include "tc_common.cuh"
#include <stdlib.h>

namespace {

template <int KC, int BN>
struct GatherCfg {
  static constexpr int ROWB = KC * 2;
  static constexpr int A_BYTES = 128 * ROWB;
  static constexpr int B_BYTES = BN * ROWB;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int STAGES_RAW = (180 * 1024) / STAGE;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = (2 * BN) < 32 ? 32 : (2 * BN);   // power of two for BN in {32,64,128,256}
  static constexpr int SMEM = STAGES * STAGE + SMEM_AUX + SCR_BYTES + COEF_BYTES + 1024;
  static constexpr uint32_t LAYOUT = (KC == 64) ? 2u : 4u;
  static constexpr uint32_t SBO = 8 * ROWB;
};

template <int KC, int BN>
__global__ void __launch_bounds__(SVK_GATHER_BOUNDS(BN), 1)
conv_tc_gather_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ GatherP p) {
  typedef GatherCfg<KC, BN> Cfg;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t stage0 = base;
  const uint32_t aux = base + Cfg::STAGES * Cfg::STAGE;
  // aux layout: full[8] @0, empty[8] @64, tfull[2] @128, tempty[2] @144, tmem ptr @160
  const uint32_t bar_full = aux, bar_empty = aux + 64, bar_tfull = aux + 128, bar_tempty = aux + 144;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(gbase + Cfg::STAGES * Cfg::STAGE + 160);
  float* scr = reinterpret_cast<float*>(gbase + Cfg::STAGES * Cfg::STAGE + SMEM_AUX);
  float* coef = reinterpret_cast<float*>(gbase + Cfg::STAGES * Cfg::STAGE + SMEM_AUX + SCR_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // arrivals that free an accumulator buffer: one per epilogue warp draining it (8 when two groups split every tile)
  const uint32_t tempty_count = (p.bn_mask && BN >= 128 && blockDim.x == GATHER_THREADS) ? 8u : 4u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, tempty_count); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "n"(Cfg::TMEM_COLS) : "memory");
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

  const int ksteps = p.ntaps * p.kchunks;

  if (warp == 0) {
    // ===================== TMA producer =====================
    {   // the WHOLE warp runs this loop (warp-uniform control flow lets ptxas keep descriptors and coordinates in
        // uniform registers); the issuing wrappers elect one lane
      int stage = 0; uint32_t ph = 0;
      const uint32_t a_bytes = (uint32_t)(p.bh * p.bw) * Cfg::ROWB;
      const uint32_t b_bytes = p.sub_n ? (uint32_t)p.sub_n * Cfg::ROWB : (uint32_t)Cfg::B_BYTES;   // column-pair mode: one sub-accumulator's filter rows
      const bool prof = p.prof != nullptr;
      long long pw = 0; const long long pt0 = prof ? clock64() : 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int nblk = tile % p.n_blocks;
        int pt = tile / p.n_blocks;
        const int tw = pt % p.tiles_w; pt /= p.tiles_w;
        const int th = pt % p.tiles_h;
        const int n = pt / p.tiles_h;
        const int h0 = th * p.bh * p.in_mul, w0 = tw * p.bw * p.in_mul;
        for (int t = 0; t < p.ntaps; ++t) {
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait_t(bar_empty + 8 * stage, ph ^ 1u, prof, pw);
            const uint32_t sa = stage0 + stage * Cfg::STAGE;
            mbar_expect_tx(bar_full + 8 * stage, a_bytes + b_bytes);
            tma_load_4d(sa, &tmA, bar_full + 8 * stage, kc * KC, w0 + p.tap_dw[t], h0 + p.tap_dh[t], n);
            tma_load_2d(sa + Cfg::A_BYTES, &tmB, bar_full + 8 * stage, kc * KC, p.tap_w[t] * p.Nout + nblk * BN);
            if (++stage == Cfg::STAGES) { stage = 0; ph ^= 1u; }
          }
        }
      }
      if (prof) prof_flush(p.prof, 4, clock64() - pt0, pw, lane);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    {   // the WHOLE warp runs this loop (warp-uniform control flow lets ptxas keep descriptors and coordinates in
        // uniform registers); the issuing wrappers elect one lane
      // descriptors are base + (byte offset >> 4): nothing is rebuilt inside the loop (the issuing thread is
      // instruction-latency bound, profiles/r01_conv_stage1.md)
      constexpr uint32_t idesc_full = make_idesc(128, BN, 0, 0);
      const uint32_t idesc = p.sub_n ? make_idesc(128, p.sub_n, 0, 0) : idesc_full;     // column-pair mode: N = one sub-accumulator
      int stage = 0; uint32_t ph = 0;
      int acc = 0; uint32_t aph = 0;
      const uint64_t a_desc0 = make_desc(stage0, 16, Cfg::SBO, Cfg::LAYOUT);
      uint64_t a_desc = a_desc0;
      const bool prof = p.prof != nullptr;
      long long pwf = 0, pwt = 0; const long long pt0 = prof ? clock64() : 0;
      unsigned long long gt0 = 0;
      if (prof) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0));
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait_t(bar_tempty + 8 * acc, aph ^ 1u, prof, pwt);
        tc_fence_after();
        const uint32_t d_tile = tmem_base + (uint32_t)(acc * BN);
        int t = 0, kc = 0;
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait_t(bar_full + 8 * stage, ph, prof, pwf);
          tc_fence_after();
          const uint64_t b_desc = a_desc + (uint64_t)(Cfg::A_BYTES >> 4);
          // plain: one accumulator, the first k-step overwrites; column-pair: tap t feeds sub-accumulator tap_sub[t]
          const uint32_t d_tmem = p.sub_n ? d_tile + (uint32_t)(p.tap_sub[t] * p.sub_n) : d_tile;
          const uint32_t first = p.sub_n ? (uint32_t)(p.tap_first[t] && kc == 0) : (uint32_t)(ks == 0);
          tc_mma(d_tmem, a_desc, b_desc, idesc, first ? 0u : 1u);
#pragma unroll
          for (int k = 1; k < KC / 16; ++k) tc_mma(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, 1u);
          tc_commit(bar_empty + 8 * stage);
          a_desc += (uint64_t)(Cfg::STAGE >> 4);
          if (++stage == Cfg::STAGES) { stage = 0; ph ^= 1u; a_desc = a_desc0; }
          if (++kc == p.kchunks) { kc = 0; ++t; }
        }
        tc_commit(bar_tfull + 8 * acc);
        if (++acc == 2) { acc = 0; aph ^= 1u; }
      }
      if (prof && lane == 0) {
        atomicAdd(p.prof + 0, 1ull); atomicAdd(p.prof + 1, (unsigned long long)(clock64() - pt0));
        atomicAdd(p.prof + 2, (unsigned long long)pwf); atomicAdd(p.prof + 3, (unsigned long long)pwt);
        unsigned long long gt1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
        atomicAdd(p.prof + 8, gt1 - gt0); atomicMax(p.prof + 9, ~gt0); atomicMax(p.prof + 10, gt1);   // ns; [9] = ~(earliest start)
        atomicMax(p.prof + 11, gt0); atomicMax(p.prof + 12, ~gt1);                                    // latest start, ~(earliest end)
      }
    }
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
    const int n_epi = ((int)blockDim.x - 64) >> 5;
    if (!p.bn_mask) stats_mailbox_finish(p, scr, 32 * 33, n_epi, BN);
    else if (p.bn_c) stats_mailbox_finish(p, scr, 32 * SCR_STRIDE, n_epi, p.sub_n ? p.sub_n : BN);
  }
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ fprop / dgrad, CTA pair
// Same dataflow on a pair of CTAs (cluster of 2, tcgen05 cta_group::2): each CTA loads ITS pixel tile (A, 128 rows) and
// HALF of the filter block (B, BN/2 rows); the leader CTA issues M = 256 x N = BN x K = 16 instructions that read both
// CTAs' shared memory and write 128 accumulator rows into each CTA's TMEM.  The tensor pipe gains nothing from this
// (tests/bench_umma.cu: cta_group::1 already runs N >= 128 tiles at the full rate) — the point is shared-memory INGEST: the
// late stages stream single-use operands through shared memory (TMA write + UMMA read share 128 B/clk: 64 + N/2 cycles per
// instruction, profiles/r01b_what_bounds_the_convs.md), and a pair moves 25 % fewer bytes per MAC (A + B/2 instead of A + B per
// tile): 86 cycles per M=256 instruction measured at N = 128 against 143 per M=128 instruction of the single-CTA kernel.  Epilogues are the single-CTA ones: every CTA
// drains its own accumulator rows and signals the leader's accumulator-free barrier (GatherP::pair).
//   full[s]   (leader):   2 arrive.expect_tx (one per CTA's producer, default CTA-scope semantics: a .release.cluster arrive
//                         costs ~1,000 cycles here) + the bytes of both CTAs' TMA loads (cp.async.bulk.tensor .cta_group::2
//                         lets the peer's loads complete on the leader's barrier)
//   empty[s]  (each CTA): multicast tcgen05.commit of the leader when the MMAs that read stage s have completed
//   tfull[a]  (each CTA): multicast commit after the last MMA of a tile pair
//   tempty[a] (leader):   one arrive per epilogue warp of BOTH CTAs
template <int KC, int BN>
struct Gather2Cfg {
  static constexpr int ROWB = KC * 2;
  static constexpr int A_BYTES = 128 * ROWB;
  static constexpr int B_BYTES = (BN / 2) * ROWB;            // this CTA's half of the filter block
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int STAGES_RAW = (180 * 1024) / STAGE;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM = STAGES * STAGE + SMEM_AUX + SCR_BYTES + COEF_BYTES + 1024;
  static constexpr uint32_t LAYOUT = (KC == 64) ? 2u : 4u;
  static constexpr uint32_t SBO = 8 * ROWB;
};

__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2, int c3) {
  if (elect_one()) asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  if (elect_one()) asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {      // arrives on `bar` of BOTH CTAs of the pair
  if (elect_one()) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                                ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (elect_one()) asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

template <int KC, int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(SVK_GATHER_BOUNDS(BN), 1)
conv_tc_gather2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ GatherP p) {
  typedef Gather2Cfg<KC, BN> Cfg;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t stage0 = base;
  const uint32_t aux = base + Cfg::STAGES * Cfg::STAGE;
  const uint32_t bar_full = aux, bar_empty = aux + 64, bar_tfull = aux + 128, bar_tempty = aux + 144;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(gbase + Cfg::STAGES * Cfg::STAGE + 160);
  float* scr = reinterpret_cast<float*>(gbase + Cfg::STAGES * Cfg::STAGE + SMEM_AUX);
  float* coef = reinterpret_cast<float*>(gbase + Cfg::STAGES * Cfg::STAGE + SMEM_AUX + SCR_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();                 // 0 = leader (issues the MMAs)

  // arrivals that free an accumulator buffer: the epilogue warps of BOTH CTAs that drain it
  const uint32_t per_cta = (p.bn_mask && BN >= 128 && blockDim.x == GATHER_THREADS) ? 8u : 4u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(bar_full + 8 * s, 2); mbar_init(bar_empty + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 2 * per_cta); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "n"(Cfg::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
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
  cluster_sync_all();                                       // both CTAs' barriers are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int ksteps = p.ntaps * p.kchunks;
  const int n_pairs = (p.total_tiles + 1) >> 1;             // n_blocks == 1: tile = pixel tile; pair k = tiles (2k, 2k+1)
  const int pair0 = (int)(blockIdx.x >> 1), pair_step = (int)(gridDim.x >> 1);

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    int stage = 0; uint32_t ph = 0;
    const uint32_t a_bytes = (uint32_t)(p.bh * p.bw) * Cfg::ROWB;
    const bool prof = p.prof != nullptr;
    long long pw = 0; const long long pt0 = prof ? clock64() : 0;
    for (int k = pair0; k < n_pairs; k += pair_step) {
      int tile = 2 * k + (int)rank;
      if (tile >= p.total_tiles) tile = 2 * k;              // odd tile count: the last pair's second tile is a duplicate (never stored)
      int pt = tile;
      const int tw = pt % p.tiles_w; pt /= p.tiles_w;
      const int th = pt % p.tiles_h;
      const int n = pt / p.tiles_h;
      const int h0 = th * p.bh * p.in_mul, w0 = tw * p.bw * p.in_mul;
      for (int t = 0; t < p.ntaps; ++t) {
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait_t(bar_empty + 8 * stage, ph ^ 1u, prof, pw);
          const uint32_t sa = stage0 + stage * Cfg::STAGE;
          const uint32_t lfull = mapa_u32(bar_full + 8 * stage, 0);      // the leader's barrier of this stage
          mbar_expect_tx_cluster(lfull, a_bytes + (uint32_t)Cfg::B_BYTES);
          tma_load_4d_pair(sa, &tmA, lfull, kc * KC, w0 + p.tap_dw[t], h0 + p.tap_dh[t], n);
          tma_load_2d_pair(sa + Cfg::A_BYTES, &tmB, lfull, kc * KC, p.tap_w[t] * p.Nout + (int)rank * (BN / 2));
          if (++stage == Cfg::STAGES) { stage = 0; ph ^= 1u; }
        }
      }
    }
    if (prof && rank == 0) prof_flush(p.prof, 4, clock64() - pt0, pw, lane);
    if (prof && rank == 1 && lane == 0) atomicAdd(p.prof + 14, (unsigned long long)pw);      // peer producer: waiting for a free slot
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc(256, BN, 0, 0);
      int stage = 0; uint32_t ph = 0;
      int acc = 0; uint32_t aph = 0;
      const uint64_t a_desc0 = make_desc(stage0, 16, Cfg::SBO, Cfg::LAYOUT);
      uint64_t a_desc = a_desc0;
      const bool prof = p.prof != nullptr;
      long long pwf = 0, pwt = 0; const long long pt0 = prof ? clock64() : 0;
      for (int k = pair0; k < n_pairs; k += pair_step) {
        mbar_wait_t(bar_tempty + 8 * acc, aph ^ 1u, prof, pwt);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait_t(bar_full + 8 * stage, ph, prof, pwf);
          tc_fence_after();
          const uint64_t b_desc = a_desc + (uint64_t)(Cfg::A_BYTES >> 4);
          tc_mma_pair(d_tmem, a_desc, b_desc, idesc, ks != 0 ? 1u : 0u);
#pragma unroll
          for (int kk = 1; kk < KC / 16; ++kk) tc_mma_pair(d_tmem, a_desc + 2 * kk, b_desc + 2 * kk, idesc, 1u);
          tc_commit_pair(bar_empty + 8 * stage);
          a_desc += (uint64_t)(Cfg::STAGE >> 4);
          if (++stage == Cfg::STAGES) { stage = 0; ph ^= 1u; a_desc = a_desc0; }
        }
        tc_commit_pair(bar_tfull + 8 * acc);
        if (++acc == 2) { acc = 0; aph ^= 1u; }
      }
      if (prof && lane == 0) {
        atomicAdd(p.prof + 0, 1ull); atomicAdd(p.prof + 1, (unsigned long long)(clock64() - pt0));
        atomicAdd(p.prof + 2, (unsigned long long)pwf); atomicAdd(p.prof + 3, (unsigned long long)pwt);
      }
    }
  } else {
    // the single-CTA epilogues: tile = blockIdx.x + j * gridDim.x = 2 * pair + rank
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
  if (p.stats && stats_use_mailbox(p)) {       // this CTA's statistics (its 128 rows of every pair tile)
    const int n_epi = ((int)blockDim.x - 64) >> 5;
    if (!p.bn_mask) stats_mailbox_finish(p, scr, 32 * 33, n_epi, BN);
    else if (p.bn_c) stats_mailbox_finish(p, scr, 32 * SCR_STRIDE, n_epi, p.sub_n ? p.sub_n : BN);
  }
  cluster_sync_all();                                       // no CTA leaves while its pair may still signal its barriers
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ wgrad kernel
struct WgradP {
  int bh, bw, P;              // pixel tile; P = bh*bw (multiple of 16) = K extent of one tile
  int tiles_h, tiles_w, num_pix_tiles;
  int stride;                 // forward conv stride (input coord = out coord * stride + tap offset)
  int R;
  int Cin, Cout;
  int cin_blk, n_cin_blk, n_m_blk, n_tap_grp, taps_per_grp, ntaps;
  int ksplit;                 // CTAs per (m_blk, cin_blk, tap group); every one of them owns >= 1 pixel tile
  int tiles_per;              // pixel tiles per CTA
  int rshift;                 // 1: 3x3/s1 — a tap group is one filter COLUMN s; its three taps r=0..2 read the same
                              //    (bh+2) x bw halo tile of x at K offsets r*bw rows (one TMA load instead of three)
  int xrows;                  // rows of one x tile in smem: (bh+2)*bw (rshift) or P
  int b_stages;               // x-tile ring depth (2 or 3)
  float* ws;                  // split-K partials [ksplit][taps][Cout][Cin] fp32 (plain stores, no atomics)
  long long ws_stride;        // taps*Cout*Cin
};

// CK = channels per swizzle atom row: 64 (SW128) when both Cin and Cout are multiples of 64, else 32 (SW64).
template <int CK>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX,
                     const __grid_constant__ WgradP p) {
  constexpr int ROWB = CK * 2;
  constexpr uint32_t LAYOUT = (CK == 64) ? 2u : 4u;
  constexpr uint32_t SBO = 8 * ROWB;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int m_chunks = 128 / CK;                       // dy atoms per tile slot (only those < Cout are fetched)
  const int a_real = (p.Cout < 128 ? p.Cout : 128) / CK;
  const int b_chunks = p.cin_blk / CK;
  const uint32_t a_atom = (uint32_t)p.P * ROWB;        // one CK-channel chunk of a dy tile
  const uint32_t b_atom = (uint32_t)p.xrows * ROWB;    // one CK-channel chunk of an x tile
  const uint32_t A_BYTES = a_atom * m_chunks;
  const uint32_t B_BYTES = b_atom * b_chunks;
  const uint32_t a0 = base;                            // 2 dy slots
  const uint32_t b0 = base + 2 * A_BYTES;              // b_stages x slots
  const uint32_t auxoff = 2 * A_BYTES + p.b_stages * B_BYTES;
  const uint32_t aux = base + auxoff;
  const uint32_t bar_afull = aux, bar_aempty = aux + 16, bar_bfull = aux + 32, bar_bempty = aux + 64, bar_done = aux + 96;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(gbase + auxoff + 128);

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(bar_afull + 8 * s, 1); mbar_init(bar_aempty + 8 * s, 1); }
    for (int s = 0; s < p.b_stages; ++s) { mbar_init(bar_bfull + 8 * s, 1); mbar_init(bar_bempty + 8 * s, 1); }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // work item
  int wi = blockIdx.x;
  const int ks = wi % p.ksplit; wi /= p.ksplit;
  const int tg = wi % p.n_tap_grp; wi /= p.n_tap_grp;
  const int cb = wi % p.n_cin_blk;
  const int mb = wi / p.n_cin_blk;
  // taps of this group: rshift -> (r, s = tg) for r = 0..R-1;  else tap0 .. tap0+ntap-1
  const int tap0 = p.rshift ? tg : tg * p.taps_per_grp;
  const int tap_step = p.rshift ? p.R : 1;
  const int ntap = p.rshift ? p.R : ((p.ntaps - tap0) < p.taps_per_grp ? (p.ntaps - tap0) : p.taps_per_grp);
  const int loads = p.rshift ? 1 : ntap;               // x-tile loads per pixel tile
  const int taps_per_load = p.rshift ? ntap : 1;
  const int t_beg = ks * p.tiles_per;
  const int t_end = (t_beg + p.tiles_per) < p.num_pix_tiles ? (t_beg + p.tiles_per) : p.num_pix_tiles;
  const int pad = p.R / 2;

  if (warp == 0) {
    {   // the WHOLE warp runs this loop (warp-uniform control flow lets ptxas keep descriptors and coordinates in
        // uniform registers); the issuing wrappers elect one lane
      int bs = 0; uint32_t bph = 0;
      int as = 0; uint32_t aph = 0;
      for (int tile = t_beg; tile < t_end; ++tile) {
        int pt = tile;
        const int tw = pt % p.tiles_w; pt /= p.tiles_w;
        const int th = pt % p.tiles_h;
        const int n = pt / p.tiles_h;
        const int h0 = th * p.bh, w0 = tw * p.bw;
        mbar_wait(bar_aempty + 8 * as, aph ^ 1u);
        mbar_expect_tx(bar_afull + 8 * as, a_atom * a_real);
        for (int c = 0; c < a_real; ++c)
          tma_load_4d(a0 + as * A_BYTES + c * a_atom, &tmDy, bar_afull + 8 * as, mb * 128 + c * CK, w0, h0, n);
        if (++as == 2) { as = 0; aph ^= 1u; }
        for (int l = 0; l < loads; ++l) {
          int cw, chh;
          if (p.rshift) { cw = w0 + tg - pad; chh = h0 - pad; }
          else { const int tap = tap0 + l; cw = w0 * p.stride + tap % p.R - pad; chh = h0 * p.stride + tap / p.R - pad; }
          mbar_wait(bar_bempty + 8 * bs, bph ^ 1u);
          mbar_expect_tx(bar_bfull + 8 * bs, B_BYTES);
          for (int c = 0; c < b_chunks; ++c)
            tma_load_4d(b0 + bs * B_BYTES + c * b_atom, &tmX, bar_bfull + 8 * bs, cb * p.cin_blk + c * CK, cw, chh, n);
          if (++bs == p.b_stages) { bs = 0; bph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    {   // the WHOLE warp runs this loop (warp-uniform control flow lets ptxas keep descriptors and coordinates in
        // uniform registers); the issuing wrappers elect one lane
      const uint32_t idesc = make_idesc(128, p.cin_blk, 1, 1);
      int bs = 0; uint32_t bph = 0;
      int as = 0; uint32_t aph = 0;
      const int ksteps = p.P / 16;
      for (int tile = t_beg; tile < t_end; ++tile) {
        mbar_wait(bar_afull + 8 * as, aph);
        tc_fence_after();
        for (int l = 0; l < loads; ++l) {
          mbar_wait(bar_bfull + 8 * bs, bph);
          tc_fence_after();
          for (int t = 0; t < taps_per_load; ++t) {
            const int slot = p.rshift ? t : l;                           // accumulator region of this tap
            const uint32_t d_tmem = tmem_base + (uint32_t)(slot * p.cin_blk);
            const uint32_t boff = p.rshift ? (uint32_t)(t * p.bw) * ROWB : 0u;   // K offset of tap r inside the halo tile
            // MN-major operands: LBO = stride between CK-channel atoms, SBO = stride between 8-pixel groups; K
            // advances by 16 pixel rows = 16*ROWB bytes (whole swizzle atoms; r*bw rows too since bw % 8 == 0).
            // Descriptors are built once per tap and advanced by adding (16*ROWB) >> 4 = ROWB to the address field.
            uint64_t ad = make_desc(a0 + as * A_BYTES, a_atom, SBO, LAYOUT);
            uint64_t bd = make_desc(b0 + bs * B_BYTES + boff, b_atom, SBO, LAYOUT);
            tc_mma(d_tmem, ad, bd, idesc, tile != t_beg ? 1u : 0u);
            for (int k = 1; k < ksteps; ++k) {
              ad += ROWB; bd += ROWB;
              tc_mma(d_tmem, ad, bd, idesc, 1u);
            }
          }
          tc_commit(bar_bempty + 8 * bs);
          if (++bs == p.b_stages) { bs = 0; bph ^= 1u; }
        }
        tc_commit(bar_aempty + 8 * as);
        if (++as == 2) { as = 0; aph ^= 1u; }
      }
      tc_commit(bar_done);
    }
  } else {
    mbar_wait(bar_done, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int co = mb * 128 + q * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    float* wsl = p.ws + (long long)ks * p.ws_stride;
    for (int t = 0; t < ntap; ++t) {
      const int tap = tap0 + t * tap_step;
      for (int c = 0; c < p.cin_blk / 32; ++c) {
        uint32_t r[32];
        tc_ld32(taddr + t * p.cin_blk + c * 32, r);
        if (co < p.Cout) {
          float4* dst = reinterpret_cast<float4*>(wsl + ((long long)tap * p.Cout + co) * p.Cin + cb * p.cin_blk + c * 32);
#pragma unroll
          for (int e = 0; e < 8; ++e)
            dst[e] = make_float4(__uint_as_float(r[4 * e]), __uint_as_float(r[4 * e + 1]), __uint_as_float(r[4 * e + 2]),
                                 __uint_as_float(r[4 * e + 3]));
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

// ------------------------------------------------------------------------------------------ host side
// Pick the bh x bw (<= max_rows pixels, bw <= max_bw) tile that wastes the fewest accumulator rows on an Hc x Wc grid.
void pick_tile(int Hc, int Wc, int max_rows, int max_bw, int row_mult, int* bh_out, int* bw_out) {
  double best = -1.0; int bbh = 1, bbw = 1;
  for (int bw = 1; bw <= max_bw && bw <= max_rows; ++bw) {
    int bh = max_rows / bw;
    if (bh > Hc) bh = Hc;
    if (bh > 128) bh = 128;
    while (bh > 1 && (bh * bw) % row_mult != 0) --bh;
    if ((bh * bw) % row_mult != 0) continue;
    int th = (Hc + bh - 1) / bh, tw = (Wc + bw - 1) / bw;
    double eff = (double)Hc * Wc / ((double)th * tw * max_rows);
    if (row_mult > 1) eff = (double)Hc * Wc / ((double)th * tw * bh * bw);   // wgrad: cost ~ rows actually loaded
    if (eff > best + 1e-9 || (eff > best - 1e-9 && bw > bbw)) { best = eff; bbh = bh; bbw = bw; }
  }
  *bh_out = bbh; *bw_out = bbw;
}

template <int KC, int BN>
int launch_gather_t(const CUtensorMap& ta, const CUtensorMap& tb, const GatherP& p, cudaStream_t st) {
  typedef GatherCfg<KC, BN> Cfg;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_gather_kernel<KC, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    SVK_REQUIRE(e == cudaSuccess, (int)e, "conv_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    configured = true;
  }
  int grid = p.total_tiles < svk_num_sms() ? p.total_tiles : svk_num_sms();
  // 8 epilogue warps for narrow tiles (their epilogue outlasts their MMAs) and for the fused BatchNorm-backward epilogue
  svk_launch(conv_tc_gather_kernel<KC, BN>, grid, (BN <= 64 || p.bn_mask) ? GATHER_THREADS : TC_THREADS, Cfg::SMEM, st, ta, tb, p);
  SVK_LAUNCH_CHECK("conv_tc_gather");
  return 0;
}
template <int KC, int BN>
int launch_gather2_t(const CUtensorMap& ta, const CUtensorMap& tb, const GatherP& p, cudaStream_t st) {
  typedef Gather2Cfg<KC, BN> Cfg;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_gather2_kernel<KC, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    SVK_REQUIRE(e == cudaSuccess, (int)e, "conv_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    configured = true;
  }
  const int n_pairs = (p.total_tiles + 1) / 2;
  const int threads = p.bn_mask ? GATHER_THREADS : TC_THREADS;     // 8 epilogue warps for the plain epilogue measured slower (10.38 vs 10.28 ms per step)
  // a pair needs two SMs of one TPC: ask the driver how many pairs can be resident at once (fewer than SMs / 2 when TPCs
  // have a single enabled SM) — a persistent grid larger than that runs in two waves
  static int max_pairs[2] = {0, 0};
  int& mp = max_pairs[p.bn_mask ? 1 : 0];
  if (mp == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(svk_num_sms() & ~1); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = Cfg::SMEM;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nc = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, conv_tc_gather2_kernel<KC, BN>, &cfg);
    if (e != cudaSuccess || nc <= 0) { cudaGetLastError(); nc = svk_num_sms() / 2; }
    mp = nc;
    if (getenv("SVK_VERBOSE")) fprintf(stderr, "svk: conv_tc_gather2<%d,%d> threads %d: %d resident CTA pairs\n", KC, BN, threads, nc);
  }
  const int grid = 2 * (n_pairs < mp ? n_pairs : mp);
  svk_launch(conv_tc_gather2_kernel<KC, BN>, grid, threads, Cfg::SMEM, st, ta, tb, p);
  SVK_LAUNCH_CHECK("conv_tc_gather2");
  return 0;
}
// CTA-pair variant: one N block of 128 / 256 channels, 64-channel K chunks, at least one pair of tiles.  Stage 3 fprop / dgrad
// 54.8 / 40.8 us against 58.8 / 48.1 for the single-CTA kernel, stage 4 45.0 / 35.1 against 46.4 / 36.3 (tests/bench_conv.py);
// 10.34 -> 10.28 ms per training step.  SVK_DISABLE_PAIR=1 selects the single-CTA kernel.
bool gather2_applicable(int KC, int BN, const GatherP& p) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("SVK_DISABLE_PAIR"); on = (e && e[0] == '1') ? 0 : 1; }
  return on && KC == 64 && (BN == 128 || BN == 256) && p.n_blocks == 1 && p.total_tiles >= 2;
}

int launch_gather(int KC, int BN, const CUtensorMap& ta, const CUtensorMap& tb, const GatherP& p, cudaStream_t st) {
#define SVK_G(K_, N_) if (KC == K_ && BN == N_) return launch_gather_t<K_, N_>(ta, tb, p, st)
  SVK_G(64, 32); SVK_G(64, 64); SVK_G(64, 128); SVK_G(64, 256);
  SVK_G(32, 32); SVK_G(32, 64); SVK_G(32, 128); SVK_G(32, 256);
#undef SVK_G
  svk_set_error("conv_tc: no kernel for KC=%d BN=%d", KC, BN);
  return SVK_E_UNSUPPORTED;
}

// Common driver for fprop and the (per parity class) dgrad launches.
int run_gather(const bf16* in, int N, int Hin, int Win, int Kc,           // gathered tensor
               const bf16* w, int ntaps_total, int Nout,                  // packed weights [taps][Nout][Kc]
               GatherP p, int es, cudaStream_t st) {
  SVK_REQUIRE(Kc % 32 == 0 && Nout % 32 == 0 && Nout <= 512, SVK_E_UNSUPPORTED,
              "conv_tc: channels must be multiples of 32 (<= 512 out), got Kc=%d Nout=%d", Kc, Nout);
  const int KC = (Kc % 64 == 0) ? 64 : 32;
  int BN = p.sub_n ? 2 * Nout : Nout;          // column-pair mode: two sub-accumulators of Nout columns side by side
  if (BN > 256) BN = 256;
  SVK_REQUIRE(BN == 32 || BN == 64 || BN == 128 || BN == 256, SVK_E_UNSUPPORTED, "conv_tc: Nout=%d unsupported", Nout);
  SVK_REQUIRE(p.sub_n ? (BN == 2 * Nout) : (Nout % BN == 0), SVK_E_UNSUPPORTED, "conv_tc: Nout=%d not a multiple of %d", Nout, BN);
  pick_tile(p.Hc, p.Wc, 128, 128 / (es > 1 ? 1 : 1), 1, &p.bh, &p.bw);
  if (p.bw * es > 256) p.bw = 256 / es;
  p.tiles_h = (p.Hc + p.bh - 1) / p.bh;
  p.tiles_w = (p.Wc + p.bw - 1) / p.bw;
  p.num_pix_tiles = N * p.tiles_h * p.tiles_w;
  p.n_blocks = p.sub_n ? 1 : Nout / BN;
  p.total_tiles = p.num_pix_tiles * p.n_blocks;
  p.kchunks = Kc / KC;
  p.Nout = Nout;
  p.prof = svk_prof_buffer();
  CUtensorMap ta, tb;
  if (int e = make_nhwc_map(&ta, in, N, Hin, Win, Kc, KC, p.bw, p.bh, es)) return e;
  if (p.sub_n) {
    if (int e = make_w_map(&tb, w, (long long)ntaps_total * Nout, Kc, KC, Nout)) return e;       // one sub-accumulator's rows per tap
    return launch_gather(KC, BN, ta, tb, p, st);
  }
  if (gather2_applicable(KC, BN, p)) {
    p.pair = 1;
    if (int e = make_w_map(&tb, w, (long long)ntaps_total * Nout, Kc, KC, BN / 2)) return e;     // each CTA loads half a filter block
    return BN == 128 ? launch_gather2_t<64, 128>(ta, tb, p, st) : launch_gather2_t<64, 256>(ta, tb, p, st);
  }
  if (int e = make_w_map(&tb, w, (long long)ntaps_total * Nout, Kc, KC, BN)) return e;
  return launch_gather(KC, BN, ta, tb, p, st);
}

}  // namespace

// SVK_PROF=1 (debug): a 16-counter device buffer the gather kernels add their per-role cycle counts to.
unsigned long long* svk_prof_buffer() {
  static int on = -1;
  static unsigned long long* buf = nullptr;
  if (on < 0) {
    const char* e = getenv("SVK_PROF");
    on = (e && e[0] == '1') ? 1 : 0;
    if (on && (cudaMalloc(&buf, 16 * sizeof(unsigned long long)) != cudaSuccess ||
               cudaMemset(buf, 0, 16 * sizeof(unsigned long long)) != cudaSuccess)) { buf = nullptr; on = 0; }
  }
  return buf;
}
// Copies the 16 counters to `out` (host) and clears them; returns SVK_E_UNSUPPORTED unless SVK_PROF=1.
SVK_API int svk_debug_prof_read(unsigned long long* out) {
  unsigned long long* b = svk_prof_buffer();
  SVK_REQUIRE(b != nullptr, SVK_E_UNSUPPORTED, "svk_debug_prof_read: set SVK_PROF=1 before the first convolution call");
  cudaError_t e = cudaMemcpy(out, b, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  SVK_REQUIRE(e == cudaSuccess, (int)e, "svk_debug_prof_read: %s", cudaGetErrorString(e));
  e = cudaMemset(b, 0, 16 * sizeof(unsigned long long));
  SVK_REQUIRE(e == cudaSuccess, (int)e, "svk_debug_prof_read: %s", cudaGetErrorString(e));
  return 0;
}

// conv_tc3.cu: resident-filter halo kernel for 3x3/s1 with Cout <= 64
bool svk_gather3_applicable(int R, int stride, int Kc, int Nout);
int svk_conv3x3s1_gather3_tc(const void* in, int N, int Hc, int Wc, int Kc, const void* w_packed, int Nout, int dgrad,
                             GatherP p, cudaStream_t st);
static bool gather3_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("SVK_DISABLE_GATHER3"); v = (e && e[0] == '1') ? 0 : 1; }
  return v == 1;
}

int svk_conv2d_fwd_tc(const svk_conv_desc* d, const void* x, const void* w, void* y, double* stats, const float* scale,
                      const float* shift, const void* residual, int relu, const int* valid_wo, cudaStream_t st) {
  GatherP p{};
  p.Hc = d->Ho; p.Wc = d->Wo;
  p.in_mul = d->stride;
  p.ntaps = d->R * d->R;
  const int pad = d->R / 2;
  for (int r = 0; r < d->R; ++r)
    for (int s = 0; s < d->R; ++s) { int t = r * d->R + s; p.tap_dh[t] = r - pad; p.tap_dw[t] = s - pad; p.tap_w[t] = t; }
  p.Hout = d->Ho; p.Wout = d->Wo; p.o_mul = 1; p.o_off_h = 0; p.o_off_w = 0;
  p.out = (bf16*)y; p.scale = scale; p.shift = shift; p.res = (const bf16*)residual; p.res_m = nullptr; p.mask = nullptr;
  p.relu = relu; p.valid_w = valid_wo; p.stats = stats;
  if (gather3_enabled() && svk_gather3_applicable(d->R, d->stride, d->Cin, d->Cout))
    return svk_conv3x3s1_gather3_tc(x, d->N, d->H, d->W, d->Cin, w, d->Cout, 0, p, st);
  return run_gather((const bf16*)x, d->N, d->H, d->W, d->Cin, (const bf16*)w, d->R * d->R, d->Cout, p, d->stride, st);
}

// Stride-2 data gradient: one launch per output parity class; res00 is the additive tensor of class (0,0), res_rest
// of the other three.
static bool s2_pair_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("SVK_DISABLE_S2_PAIR"); v = (e && e[0] == '1') ? 0 : 1; }
  return v == 1;
}
static int dgrad_s2_classes(const svk_conv_desc* d, const void* dy, const void* w, void* dx, const void* res00,
                            const void* res_rest, const void* res_m, const void* mask, const svk_bn_bwd_fuse* bn,
                            cudaStream_t st) {
  const int pad = d->R / 2;
  if (s2_pair_enabled() && bn && d->R == 3 && !res_m && res_rest == nullptr && (d->Cin == 32 || d->Cin == 64 || d->Cin == 128)) {
    // Column-pair mode (GatherP::sub_n): one launch per output ROW parity; the two column classes of a tile share the launch,
    // the dy loads of their common taps excepted, and every epilogue access is a full 2 * Cin-channel run.
    for (int ph = 0; ph < 2; ++ph) {
      GatherP p{};
      p.Hc = (d->H - ph + 1) / 2; p.Wc = (d->W + 1) / 2;
      if (p.Hc <= 0 || p.Wc <= 0) continue;
      p.in_mul = 1; p.ntaps = 0; p.sub_n = d->Cin; p.res_sub0 = 1;
      for (int pw = 0; pw < 2; ++pw) {
        bool first = true;
        for (int r = 0; r < d->R; ++r) {
          if (((ph + pad - r) & 1) != 0) continue;
          for (int s = 0; s < d->R; ++s) {
            if (((pw + pad - s) & 1) != 0) continue;
            int t = p.ntaps++;
            p.tap_dh[t] = (ph + pad - r) / 2; p.tap_dw[t] = (pw + pad - s) / 2; p.tap_w[t] = r * d->R + s;
            p.tap_sub[t] = pw; p.tap_first[t] = first ? 1 : 0;
            first = false;
          }
        }
      }
      p.Hout = d->H; p.Wout = d->W; p.o_mul = 2; p.o_off_h = ph; p.o_off_w = 0;
      p.out = (bf16*)dx; p.res = (const bf16*)(ph ? nullptr : res00);
      p.bn_mask = (const bf16*)bn->mask;
      if (bn->c) { p.bn_c = (const bf16*)bn->c; p.bn_mean = bn->mean; p.bn_rstd = bn->rstd; p.stats = bn->sums; }
      if (int e = run_gather((const bf16*)dy, d->N, d->Ho, d->Wo, d->Cout, (const bf16*)w, d->R * d->R, d->Cin, p, 1, st)) return e;
    }
    return 0;
  }
  for (int ph = 0; ph < 2; ++ph) {
    for (int pw = 0; pw < 2; ++pw) {
      GatherP p{};
      p.Hc = (d->H - ph + 1) / 2; p.Wc = (d->W - pw + 1) / 2;
      if (p.Hc <= 0 || p.Wc <= 0) continue;
      p.in_mul = 1; p.ntaps = 0;
      for (int r = 0; r < d->R; ++r) {
        if (((ph + pad - r) & 1) != 0) continue;
        for (int s = 0; s < d->R; ++s) {
          if (((pw + pad - s) & 1) != 0) continue;
          int t = p.ntaps++;
          p.tap_dh[t] = (ph + pad - r) / 2; p.tap_dw[t] = (pw + pad - s) / 2; p.tap_w[t] = r * d->R + s;   // even numerators
        }
      }
      if (p.ntaps == 0) continue;   // 1x1/s2: only class (0,0) receives gradient
      p.Hout = d->H; p.Wout = d->W; p.o_mul = 2; p.o_off_h = ph; p.o_off_w = pw;
      p.out = (bf16*)dx; p.res = (const bf16*)((ph | pw) ? res_rest : res00);
      p.res_m = (const bf16*)res_m; p.mask = (const bf16*)mask;
      if (bn) {
        p.bn_mask = (const bf16*)bn->mask;
        if (bn->c) { p.bn_c = (const bf16*)bn->c; p.bn_mean = bn->mean; p.bn_rstd = bn->rstd; p.stats = bn->sums; }
      }
      if (int e = run_gather((const bf16*)dy, d->N, d->Ho, d->Wo, d->Cout, (const bf16*)w, d->R * d->R, d->Cin, p, 1, st)) return e;
    }
  }
  return 0;
}

int svk_conv2d_dgrad_tc(const svk_conv_desc* d, const void* dy, const void* w, void* dx, const void* res,
                        const void* res_m, const void* mask, const svk_bn_bwd_fuse* bn, cudaStream_t st) {
  const int pad = d->R / 2;
  auto set_bn = [&](GatherP& p) {
    if (!bn) return;
    p.bn_mask = (const bf16*)bn->mask;
    if (bn->c) { p.bn_c = (const bf16*)bn->c; p.bn_mean = bn->mean; p.bn_rstd = bn->rstd; p.stats = bn->sums; }
  };
  if (d->stride == 1) {
    GatherP p{};
    p.Hc = d->H; p.Wc = d->W; p.in_mul = 1; p.ntaps = d->R * d->R;
    for (int r = 0; r < d->R; ++r)
      for (int s = 0; s < d->R; ++s) { int t = r * d->R + s; p.tap_dh[t] = pad - r; p.tap_dw[t] = pad - s; p.tap_w[t] = t; }
    p.Hout = d->H; p.Wout = d->W; p.o_mul = 1;
    p.out = (bf16*)dx; p.res = (const bf16*)res; p.res_m = (const bf16*)res_m; p.mask = (const bf16*)mask;
    set_bn(p);
    if (gather3_enabled() && svk_gather3_applicable(d->R, 1, d->Cout, d->Cin))
      return svk_conv3x3s1_gather3_tc(dy, d->N, d->H, d->W, d->Cout, w, d->Cin, 1, p, st);
    return run_gather((const bf16*)dy, d->N, d->Ho, d->Wo, d->Cout, (const bf16*)w, d->R * d->R, d->Cin, p, 1, st);
  }
  // stride 2: one launch per output parity class (ph, pw); dx[2i+ph, 2j+pw] gathers dy[i+dh, j+dw] for the taps whose
  // offset has the right parity.
  if (d->R == 1) {
    SVK_REQUIRE(res_m == nullptr, SVK_E_UNSUPPORTED, "conv2d_dgrad(tc, 1x1/s2): masked residual unsupported");
    if (res != dx) {
      SVK_REQUIRE(res == nullptr, SVK_E_UNSUPPORTED, "conv2d_dgrad(tc, 1x1/s2): res must be NULL or alias dx (accumulate)");
      cudaError_t e = cudaMemsetAsync(dx, 0, (size_t)d->N * d->H * d->W * d->Cin * 2, st);
      SVK_REQUIRE(e == cudaSuccess, (int)e, "conv2d_dgrad: memset failed: %s", cudaGetErrorString(e));
    }
  }
  return dgrad_s2_classes(d, dy, w, dx, res, res, res_m, mask, bn, st);
}

// dx = dgrad(conv1: 3x3/s2) + dgrad(convd: 1x1/s2), the block-input gradient of a downsample block, with the optional
// BatchNorm-backward fusion of the layer below.  The 1x1 launch writes only the (even, even) pixels; the 3x3 launch of
// that parity class adds them back in (res = dx, same thread reads then writes), the other three classes write fresh.
int svk_downsample_dgrad_tc(const svk_conv_desc* d1, const void* dy1, const void* w1, const svk_conv_desc* dd,
                            const void* dyd, const void* wd, void* dx, const svk_bn_bwd_fuse* bn, cudaStream_t st) {
  if (int e = dgrad_s2_classes(dd, dyd, wd, dx, nullptr, nullptr, nullptr, nullptr, nullptr, st)) return e;
  return dgrad_s2_classes(d1, dy1, w1, dx, dx, nullptr, nullptr, nullptr, bn, st);
}

namespace {

// Pixel tile for wgrad: P = bh*bw a multiple of 16 (UMMA K), <= 128 rows; rshift additionally needs bw % 8 == 0 so that a
// shift by r*bw rows is a whole number of swizzle atoms.  Cost = smem rows fetched per pixel of work.
void pick_wgrad_tile(int Ho, int Wo, int es, int rshift, int taps_cta, int row_bytes_total, int* bh_out, int* bw_out) {
  // cost of one pixel tile = fixed pipeline latency + max(tensor cycles, smem-ingest cycles); total = tiles * cost.
  // (Without the fixed term every exact-fit tile ties and a 16-pixel tile wins: 7,600 cycles per tile of pure overhead.)
  double best = 1e30; int bbh = 0, bbw = 0;
  const int max_bw = 256 / es;
  for (int bw = rshift ? 8 : 1; bw <= max_bw && bw <= 128; bw += rshift ? 8 : 1) {
    for (int bh = 1; bh * bw <= 128 && bh <= Ho + 1 && (bh + 2) * es <= 256; ++bh) {
      const int P = bh * bw;
      if (P % 16 != 0) continue;
      int th = (Ho + bh - 1) / bh, tw = (Wo + bw - 1) / bw;
      double x_rows = rshift ? (double)(bh + 2) * bw : (double)taps_cta * P;
      double ingest = (P + x_rows) * row_bytes_total / 48.0;
      double mma = (double)taps_cta * (P / 16) * 64.0;
      double cost = (double)th * tw * (600.0 + (mma > ingest ? mma : ingest));
      if (cost < best - 1e-9) { best = cost; bbh = bh; bbw = bw; }
    }
  }
  *bh_out = bbh; *bw_out = bbw;
}

int plan_wgrad(const svk_conv_desc* d, WgradP* pp, int* ck_out, size_t* smem_out) {
  WgradP& p = *pp;
  SVK_REQUIRE(d->Cin % 32 == 0 && d->Cout % 32 == 0, SVK_E_UNSUPPORTED, "conv2d_wgrad(tc): channels must be multiples of 32");
  const int CK = (d->Cin % 64 == 0 && d->Cout % 64 == 0) ? 64 : 32;
  p.stride = d->stride; p.R = d->R; p.Cin = d->Cin; p.Cout = d->Cout; p.ntaps = d->R * d->R;
  p.cin_blk = d->Cin < 128 ? d->Cin : 128;
  SVK_REQUIRE(d->Cin % p.cin_blk == 0, SVK_E_UNSUPPORTED, "conv2d_wgrad(tc): Cin=%d unsupported", d->Cin);
  p.n_cin_blk = d->Cin / p.cin_blk;
  p.n_m_blk = (d->Cout + 127) / 128;
  p.rshift = (d->R == 3 && d->stride == 1) ? 1 : 0;
  if (p.rshift) {
    p.n_tap_grp = 3; p.taps_per_grp = 3;
  } else {
    int max_taps = 512 / p.cin_blk;
    p.n_tap_grp = (p.ntaps + max_taps - 1) / max_taps;
    p.taps_per_grp = (p.ntaps + p.n_tap_grp - 1) / p.n_tap_grp;
  }
  pick_wgrad_tile(d->Ho, d->Wo, d->stride, p.rshift, p.taps_per_grp, (p.cin_blk > 128 ? 128 : p.cin_blk) * 2 + 64, &p.bh, &p.bw);
  SVK_REQUIRE(p.bh > 0, SVK_E_UNSUPPORTED, "conv2d_wgrad(tc): no pixel tile for %dx%d", d->Ho, d->Wo);
  p.P = p.bh * p.bw;
  p.xrows = p.rshift ? (p.bh + 2) * p.bw : p.P;
  p.tiles_h = (d->Ho + p.bh - 1) / p.bh;
  p.tiles_w = (d->Wo + p.bw - 1) / p.bw;
  p.num_pix_tiles = d->N * p.tiles_h * p.tiles_w;
  const int items = p.n_m_blk * p.n_cin_blk * p.n_tap_grp;
  int ks = svk_num_sms() / items;
  if (ks < 1) ks = 1;
  if (ks > p.num_pix_tiles) ks = p.num_pix_tiles;
  p.tiles_per = (p.num_pix_tiles + ks - 1) / ks;
  p.ksplit = (p.num_pix_tiles + p.tiles_per - 1) / p.tiles_per;     // no empty CTA
  p.ws_stride = (long long)p.ntaps * d->Cout * d->Cin;
  const size_t a_bytes = (size_t)p.P * CK * 2 * (128 / CK) * 2;
  const size_t b_slot = (size_t)p.xrows * CK * 2 * (p.cin_blk / CK);
  p.b_stages = 3;
  if (a_bytes + 3 * b_slot + 2048 > 200 * 1024) p.b_stages = 2;
  SVK_REQUIRE(a_bytes + p.b_stages * b_slot + 2048 <= 200 * 1024, SVK_E_UNSUPPORTED, "conv2d_wgrad(tc): tile does not fit smem");
  *smem_out = a_bytes + p.b_stages * b_slot + 2048;
  *ck_out = CK;
  return 0;
}

}  // namespace

size_t svk_conv2d_wgrad_tc_ws_floats(const svk_conv_desc* d) {
  WgradP p{}; int ck; size_t smem;
  if (plan_wgrad(d, &p, &ck, &smem)) return 0;
  return (size_t)p.ksplit * (size_t)p.ws_stride;
}

// Writes ksplit partial gradients [ksplit][taps][Cout][Cin] into ws; *ksplit_out tells the caller how many to reduce.
int svk_conv2d_wgrad_tc(const svk_conv_desc* d, const void* x, const void* dy, float* ws, size_t ws_floats,
                        int* ksplit_out, cudaStream_t st) {
  WgradP p{}; int CK; size_t smem;
  if (int e = plan_wgrad(d, &p, &CK, &smem)) return e;
  SVK_REQUIRE((size_t)p.ksplit * (size_t)p.ws_stride <= ws_floats, SVK_E_BADARG,
              "conv2d_wgrad(tc): workspace too small (%zu floats, need %zu)", ws_floats, (size_t)p.ksplit * (size_t)p.ws_stride);
  p.ws = ws;
  CUtensorMap tdy, tx;
  if (int e = make_nhwc_map(&tdy, dy, d->N, d->Ho, d->Wo, d->Cout, CK, p.bw, p.bh, 1)) return e;
  if (int e = make_nhwc_map(&tx, x, d->N, d->H, d->W, d->Cin, CK, p.bw, p.rshift ? p.bh + 2 : p.bh, d->stride)) return e;
  const int grid = p.n_m_blk * p.n_cin_blk * p.n_tap_grp * p.ksplit;
  if (CK == 64) {
    static bool cfg64 = false;
    if (!cfg64) { cudaError_t e = cudaFuncSetAttribute(conv_tc_wgrad_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      SVK_REQUIRE(e == cudaSuccess, (int)e, "conv_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); cfg64 = true; }
    svk_launch(conv_tc_wgrad_kernel<64>, grid, TC_THREADS, smem, st, tdy, tx, p);
  } else {
    static bool cfg32 = false;
    if (!cfg32) { cudaError_t e = cudaFuncSetAttribute(conv_tc_wgrad_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      SVK_REQUIRE(e == cudaSuccess, (int)e, "conv_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); cfg32 = true; }
    svk_launch(conv_tc_wgrad_kernel<32>, grid, TC_THREADS, smem, st, tdy, tx, p);
  }
  SVK_LAUNCH_CHECK("conv_tc_wgrad");
  *ksplit_out = p.ksplit;
  return 0;
}
