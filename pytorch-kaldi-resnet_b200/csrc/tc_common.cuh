// Shared pieces of the tcgen05 convolution kernels: PTX wrappers (mbarrier, TMA, tcgen05), UMMA descriptors, the
// fprop/dgrad parameter block and epilogue, and the TMA tensor-map builders.
#pragma once
#include "svk_common.cuh"
#if SVK_DBG_NO_STAT_ATOMICS     // diagnostic build: what the statistics atomics at the end of a kernel cost
#define SVK_DBG_ATOMICS_ON (p.Nout == 12345)
#else
#define SVK_DBG_ATOMICS_ON true
#endif
#include <cuda.h>
#include <stdlib.h>

typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------------------------------ fprop / dgrad kernel
struct GatherP {
  int Hc, Wc;                 // output class-grid (tile space)
  int bh, bw, tiles_h, tiles_w, num_pix_tiles, n_blocks, total_tiles;
  int in_mul;                 // input coordinate = class coordinate * in_mul + tap offset
  int ntaps;
  int tap_dh[9], tap_dw[9], tap_w[9];
  int kchunks;                // Kc / KC
  int Nout;                   // output channels (row pitch of out)
  int Hout, Wout, o_mul, o_off_h, o_off_w;   // out pixel = (i*o_mul + o_off_h, j*o_mul + o_off_w)
  bf16* out;
  const float* scale; const float* shift;
  const bf16* res; const bf16* res_m; const bf16* mask;
  int relu;
  const int* valid_w;
  double* stats;
  // dgrad only: BatchNorm-backward fusion (svk_bn_bwd_fuse).  out is zeroed where bn_mask <= 0; with bn_c the statistics
  // become  stats[ch] += sum(out), stats[Nout + ch] += sum(out * (bn_c - mean[ch]) * rstd[ch])  (mean/rstd staged in coef).
  const bf16* bn_mask; const bf16* bn_c; const float* bn_mean; const float* bn_rstd;
  unsigned long long* prof;   // SVK_PROF=1: per-role cycle counters (svk_debug_prof_read), else NULL
  int pitch;                  // accumulator row m = i * pitch + j (0: pitch = bw).  pitch = bw + 2: rows with j >= bw are the
                              // pad columns of a single-halo tile (conv_tc3.cu) and are discarded
  int pair;                   // 1: CTA-pair kernel (cta_group::2): accumulator-free barriers live in the pair's leader CTA
  // Column-pair mode of the stride-2 fused data gradient (sub_n > 0): one launch per output ROW parity computes the two
  // column parity classes of a tile into two sub-accumulators of sub_n = Nout columns each, laid side by side in TMEM, so an
  // accumulator row of 2 * Nout values IS the two adjacent output pixels (2 j, 2 j + 1): 2 * Nout contiguous channels in
  // memory.  The epilogue then reads / writes full lines (the one-class-per-launch form touched 64 B of every 128 B line).
  int sub_n;                  // 0 = off; else Nout (the tile template's BN is 2 * Nout)
  int tap_sub[9];             // sub-accumulator of tap t (0: even output column, 1: odd)
  int tap_first[9];           // 1: first tap of its sub-accumulator (overwrites instead of accumulating)
  int res_sub0;               // 1: the additive tensor `res` applies to sub-accumulator 0 only (the 1x1 shortcut's even pixels)
  // Accumulator buffers in TMEM, used round-robin by the CTA's tiles (tile k of the CTA -> buffer k % n_tm).  0 = one per
  // epilogue group (two for a single group / a SPLIT epilogue).  conv_tc3.cu with two MMA-issuing warps and three epilogue
  // groups uses 6, so that every buffer has exactly ONE issuing warp (k % 2) and ONE draining group (k % 3): a barrier whose
  // waiters can be two phases apart cannot be waited on by parity.
  int n_tm;
};
// prof[0] CTAs | MMA warp: [1] loop cycles [2] waiting for operands [3] waiting for a free accumulator |
// producer: [4] loop cycles [5] waiting for a free stage | first epilogue warp: [6] loop cycles [7] waiting for an accumulator
unsigned long long* svk_prof_buffer();

namespace {


// ------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a converged warp (elect.sync): the role loops are executed by whole warps, the async instructions by one lane.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  if (elect_one()) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// ---- CTA-pair (cluster of 2) helpers: address of the same smem offset in CTA `rank`, remote arrives
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default semantics (.release at CTA scope), like cutlass::arch::ClusterBarrier::arrive(cta_id): a .release.cluster arrive
  // drains the SM's outstanding memory traffic first — ~1,000 cycles each inside a TMA-fed mainloop (measured)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
  if (elect_one()) asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
// accumulator-free signal of an epilogue warp: local barrier, or (CTA-pair kernels) the barrier of the pair's leader CTA
__device__ __forceinline__ void tempty_arrive(uint32_t bar, int pair) {
  if (pair) mbar_arrive_cluster(mapa_u32(bar, 0));
  else mbar_arrive(bar);
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    if (clock64() - t0 > 8000000000LL) {  // ~4 s at 2 GHz: protocol bug, do not hang the device
      printf("svk conv_tc: mbarrier wait timed out (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      asm volatile("trap;");
    }
  }
}
// mbar_wait that adds the cycles spent waiting to `acc` when profiling
__device__ __forceinline__ void mbar_wait_t(uint32_t bar, uint32_t parity, bool prof, long long& acc) {
  if (!prof) { mbar_wait(bar, parity); return; }
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  acc += clock64() - t0;
}
__device__ __forceinline__ void prof_flush(unsigned long long* prof, int slot, long long total, long long waited, int lane) {
  if (prof && lane == 0) {
    atomicAdd(prof + slot, (unsigned long long)total);
    atomicAdd(prof + slot + 1, (unsigned long long)waited);
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  if (elect_one()) asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  if (elect_one()) asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (elect_one()) asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout [61,64) (2 = SWIZZLE_128B, 4 = SWIZZLE_64B).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c=f32 [4,6)=1, a=bf16 [7,10)=1, b=bf16 [10,13)=1,
// a_major bit15, b_major bit16 (1 = MN-major), N>>3 [17,23), M>>4 [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}


constexpr int TC_THREADS = 192;                // wgrad kernels: producer warp, MMA warp, 4 epilogue warps
constexpr int GATHER_THREADS = 320;            // fprop/dgrad kernels: producer, MMA, 2 x 4 epilogue warps (one group per
                                               // TMEM accumulator buffer: a tile's epilogue is latency-bound, ~2x the MMA time)
#define SVK_GATHER_BOUNDS(BN) GATHER_THREADS
constexpr int GATHER3_THREADS = 448;           // resident-filter kernels (conv_tc3.cu, BN <= 64): producer, MMA, 3 x 4 epilogue warps —
                                               // three accumulator buffers / epilogue groups: a narrow tile's epilogue takes ~3x its MMA time
constexpr int SMEM_AUX = 1024;                 // barriers + tmem pointer
constexpr int SCR_BYTES = 8 * 32 * 36 * 4;     // per-epilogue-warp transpose scratch (33- or 36-word rows)
constexpr int SCR3_BYTES = 12 * 32 * 36 * 4;   // ... for the 12 epilogue warps of GATHER3_THREADS
constexpr int COEF_BYTES = 2 * 512 * 4;        // scale/shift staged in smem (Nout <= 512)

// ---- epilogue helpers
__device__ __forceinline__ void bf16x8_to_f32(const uint4& r, float (&v)[8]) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
// ---- statistics of a CTA leave through a "mailbox".  Every epilogue warp used to add its per-channel partial sums to the
// global fp64 accumulators itself: 8-12 warps x 148 CTAs = 1,200-1,800 same-address atomics per channel, which serialise in L2
// AFTER the last tile — 5-31 us per launch, ~0.5 ms per training step (tests/time_conv.py with -DSVK_DBG_NO_STAT_ATOMICS=1).
// Now (one n-block per CTA, i.e. always in ResNet-34) a warp leaves its partials in its own transpose scratch, which it no
// longer needs, and after the kernel's closing __syncthreads one thread per statistic adds them up in a FIXED order, in
// double, and issues ONE atomic per CTA (148 per channel).  Box layout: float[2 * bn_eff] — sums of the first kind for the
// bn_eff channels, then of the second.
__device__ __forceinline__ bool stats_use_mailbox(const GatherP& p) { return p.n_blocks == 1; }
__device__ __forceinline__ void stats_mailbox_finish(const GatherP& p, const float* scr, int warp_stride, int n_warps, int bn_eff) {
  for (int ch = threadIdx.x; ch < 2 * bn_eff; ch += blockDim.x) {
    double sum = 0.0;
    for (int w = 0; w < n_warps; ++w) sum += (double)scr[(size_t)w * warp_stride + ch];
    if (SVK_DBG_ATOMICS_ON) atomicAdd(&p.stats[ch < bn_eff ? ch : p.Nout + (ch - bn_eff)], sum);
  }
}

// Epilogue operands of one (pixel row, 32-channel chunk), as loaded: a = additive tensor (res or res_m), b = a mask tensor
// (the mask of res_m, else bn_mask), c = bn_c.  They are fetched one or two chunks AHEAD of their use so that the global
// load latency overlaps the MMAs / the previous chunk instead of stalling the 4 epilogue warps once per chunk.
struct EpiAux { uint4 a[4], b[4], c[4]; };
struct EpiTile { long long off; int nblk; bool valid, zero_out, valid1; };     // valid1: column-pair mode, the odd pixel exists

// Epilogue of the gather kernels (4 warps): TMEM -> registers -> scale/shift -> +residual -> ReLU -> bf16 -> global, plus the
// per-channel sum / sum-of-squares of the stored values (BatchNorm batch statistics) reduced through smem.
template <int BN>
__device__ __forceinline__ void gather_epilogue(const GatherP& p, uint32_t tmem_base, uint32_t bar_tfull, uint32_t bar_tempty,
                                                float* scr, const float* coef, int warp, int lane) {
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int ngroups = ((int)blockDim.x - 64) >> 7;   // 1 (192 threads), 2 (320) or 3 (448) epilogue warp groups
    const int group = (warp - 2) >> 2;      // with G >= 2 groups, group g drains accumulator buffer g = every G-th tile
    float* myscr = scr + (warp - 2) * 32 * 33;
    constexpr int NCH = BN / 32;
    float s1[NCH], s2[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) { s1[c] = 0.f; s2[c] = 0.f; }
    int stat_blk = -1;
    const int chl = 2 * (lane & 15) + (lane >> 4);    // channel (within a chunk) whose sums this lane keeps
    int acc = group; uint32_t aph = 0;
    const int astep = ngroups >= 2 ? ngroups : 1;                       // this group's next tile is `astep` buffers further on
    const int n_tm = p.n_tm ? p.n_tm : (ngroups >= 2 ? ngroups : 2);
    const bool prof = p.prof != nullptr && warp == 2;
    long long pw = 0; const long long pt0 = prof ? clock64() : 0;
    const int m = q * 32 + lane;            // accumulator row = pixel index inside the tile (loop invariant)
    const int pitch = p.pitch ? p.pitch : p.bw;
    const int i = m / pitch, j = m - i * pitch;
    for (int tile = blockIdx.x + group * gridDim.x; tile < p.total_tiles; tile += ngroups * gridDim.x) {
      const int nblk = tile % p.n_blocks;
      int pt = tile / p.n_blocks;
      const int tw = pt % p.tiles_w; pt /= p.tiles_w;
      const int th = pt % p.tiles_h;
      const int n = pt / p.tiles_h;
      if (p.stats && stat_blk != nblk) {
        if (stat_blk >= 0) {
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            if (SVK_DBG_ATOMICS_ON) atomicAdd(&p.stats[stat_blk * BN + c * 32 + chl], (double)s1[c]);
            if (SVK_DBG_ATOMICS_ON) atomicAdd(&p.stats[p.Nout + stat_blk * BN + c * 32 + chl], (double)s2[c]);
            s1[c] = 0.f; s2[c] = 0.f;
          }
        }
        stat_blk = nblk;
      }
      const int hc = th * p.bh + i, wc = tw * p.bw + j;
      bool valid = (i < p.bh) && (j < p.bw) && (hc < p.Hc) && (wc < p.Wc);
      const int oh = hc * p.o_mul + p.o_off_h, ow = wc * p.o_mul + p.o_off_w;
      const long long off = valid ? ((((long long)n * p.Hout + oh) * p.Wout + ow) * p.Nout + nblk * BN) : 0;
      bool zero_out = false;
      if (valid && p.valid_w) zero_out = ow >= p.valid_w[n];

      mbar_wait_t(bar_tfull + 8 * acc, aph, prof, pw);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        uint32_t r[32];
        tc_ld32(taddr + c * 32, r);
#if SVK_DBG_EPI_SKIP == 1   // diagnostic build: drain the accumulator and do nothing else
        s1[c] += __uint_as_float(r[0]) + __uint_as_float(r[31]);
        continue;
#endif
        float v[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r[e]);
        if (p.scale) {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = fmaf(v[e], coef[nblk * BN + c * 32 + e], coef[512 + nblk * BN + c * 32 + e]);
        }
        if (valid) {
          if (p.res) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float t8[8];
              Vec<bf16>::load(p.res + off + c * 32 + g * 8, t8);
#pragma unroll
              for (int e = 0; e < 8; ++e) v[g * 8 + e] += t8[e];
            }
          }
          if (p.res_m) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float t8[8], k8[8];
              Vec<bf16>::load(p.res_m + off + c * 32 + g * 8, t8);
              Vec<bf16>::load(p.mask + off + c * 32 + g * 8, k8);
#pragma unroll
              for (int e = 0; e < 8; ++e) v[g * 8 + e] += (k8[e] > 0.f) ? t8[e] : 0.f;
            }
          }
        }
        if (p.relu) {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], 0.f);
        }
        if (zero_out) {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = 0.f;
        }
        uint32_t packed[16];                 // the 32 output values of this row as stored: bf16x2 words
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
          packed[e] = valid ? *reinterpret_cast<uint32_t*>(&h) : 0u;
        }
#if SVK_DBG_EPI_SKIP == 2   // diagnostic build: no global stores
        if (valid && packed[3] == 0x12345678u) {
#else
        if (valid) {
#endif
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4*>(p.out + off + c * 32 + g * 8) =
                make_uint4(packed[g * 4], packed[g * 4 + 1], packed[g * 4 + 2], packed[g * 4 + 3]);
        }
        if (p.stats) {
          // statistics of the values as stored.  The packed rows are transposed through smem (the kernel is bound by
          // shared-memory bandwidth, so the bf16x2 words halve what this costs): row l starts at word 16 l + 4 (l / 2),
          // conflict-free for the 16-byte row writes and for the column reads below, where lane (h, j) sums the channel
          // pair j over the rows of parity h.
          uint32_t* row = reinterpret_cast<uint32_t*>(myscr) + 16 * lane + 4 * (lane >> 1);
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4*>(row + g * 4) = make_uint4(packed[g * 4], packed[g * 4 + 1], packed[g * 4 + 2], packed[g * 4 + 3]);
          __syncwarp();
          const int hh = lane >> 4, jj = lane & 15;
          const uint32_t* col = reinterpret_cast<const uint32_t*>(myscr) + jj;
          float a10 = 0.f, a11 = 0.f, a20 = 0.f, a21 = 0.f;
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const int rr = 2 * k + hh;
            const uint32_t wv = col[16 * rr + 4 * k];            // 4 * (rr / 2) == 4 * k
            const float x0 = __uint_as_float(wv << 16), x1 = __uint_as_float(wv & 0xffff0000u);
            a10 += x0; a20 = fmaf(x0, x0, a20);
            a11 += x1; a21 = fmaf(x1, x1, a21);
          }
          __syncwarp();
          a10 += __shfl_xor_sync(0xffffffffu, a10, 16); a11 += __shfl_xor_sync(0xffffffffu, a11, 16);
          a20 += __shfl_xor_sync(0xffffffffu, a20, 16); a21 += __shfl_xor_sync(0xffffffffu, a21, 16);
          // lane (h, j) keeps channel 2 j + h of this chunk
          s1[c] += hh ? a11 : a10;
          s2[c] += hh ? a21 : a20;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) tempty_arrive(bar_tempty + 8 * acc, p.pair);
      acc += astep;
      if (acc >= n_tm) { acc -= n_tm; aph ^= 1u; }
    }
    if (p.stats && stats_use_mailbox(p)) {
      __syncwarp();                         // the last column pass has read this warp's scratch
#pragma unroll
      for (int c = 0; c < NCH; ++c) { myscr[c * 32 + chl] = s1[c]; myscr[BN + c * 32 + chl] = s2[c]; }
    } else if (p.stats && stat_blk >= 0) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        if (SVK_DBG_ATOMICS_ON) atomicAdd(&p.stats[stat_blk * BN + c * 32 + chl], (double)s1[c]);
        if (SVK_DBG_ATOMICS_ON) atomicAdd(&p.stats[p.Nout + stat_blk * BN + c * 32 + chl], (double)s2[c]);
      }
    }
    if (prof) prof_flush(p.prof, 6, clock64() - pt0, pw, lane);
  }


// Data-gradient epilogue with the BatchNorm-backward fusion (svk_bn_bwd_fuse): out = (acc [+ res]) * (bn_mask > 0), and,
// with bn_c, stats[ch] += sum(out), stats[Nout + ch] += sum(out * (bn_c - mean[ch]) * rstd[ch]) over the stored values.
// Up to three operand streams (res, bn_mask, bn_c) are fetched one or two chunks AHEAD of their use, so their latency
// overlaps the MMAs instead of stalling the epilogue warps once per chunk.  The (g, c) pairs go through shared memory
// packed as bf16x2 in ONE transpose: the kernels are bound by shared-memory bandwidth (UMMA operand reads), so the
// epilogue's own smem traffic is what it costs.
constexpr int SCR_STRIDE = 36;      // words per scratch row: conflict-free for 16-byte row writes and 4-byte column reads
// SPLIT (BN >= 128, 8 epilogue warps): both warp groups drain EVERY tile, each one half of its 32-column chunks (four
// warps cannot issue the fused epilogue of a 128/256-wide tile within its MMA time); otherwise the groups alternate tiles.
template <int BN, bool SPLIT>
__device__ __forceinline__ void gather_epilogue_bn(const GatherP& p, uint32_t tmem_base, uint32_t bar_tfull,
                                                   uint32_t bar_tempty, float* scr, const float* coef, int warp, int lane) {
    const int q = warp & 3;
    const int ngroups = ((int)blockDim.x - 64) >> 7;
    const int group = (warp - 2) >> 2;
    uint32_t* myscr = reinterpret_cast<uint32_t*>(scr) + (warp - 2) * 32 * SCR_STRIDE;
    constexpr int NCH = SPLIT ? BN / 64 : BN / 32;     // chunks this warp handles per tile
    constexpr int DEPTH = 1;                           // chunks of look-ahead
    const int c_first = SPLIT ? group * NCH : 0;       // first 32-column chunk of this warp inside a tile
    const bf16* pa = p.res;
    const bf16* pb = p.bn_mask;
    const bf16* pc = p.bn_c;
    float s1[NCH], s2[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) { s1[c] = 0.f; s2[c] = 0.f; }
    int stat_blk = -1;
    int acc = SPLIT ? 0 : group; uint32_t aph = 0;
    const int astep = (!SPLIT && ngroups >= 2) ? ngroups : 1;
    const int n_tm = (!SPLIT && p.n_tm) ? p.n_tm : ((!SPLIT && ngroups >= 2) ? ngroups : 2);
    const bool prof = p.prof != nullptr && warp == 2;
    long long pw = 0; const long long pt0 = prof ? clock64() : 0;
    const int m = q * 32 + lane;
    const int pitch = p.pitch ? p.pitch : p.bw;
    const int i = m / pitch, j = m - i * pitch;
    const int tstep = SPLIT ? (int)gridDim.x : ngroups * (int)gridDim.x;

    auto tile_info = [&](int tile) {
      EpiTile t;
      t.nblk = tile % p.n_blocks;
      int pt = tile / p.n_blocks;
      const int tw = pt % p.tiles_w; pt /= p.tiles_w;
      const int th = pt % p.tiles_h;
      const int n = pt / p.tiles_h;
      const int hc = th * p.bh + i, wc = tw * p.bw + j;
      t.valid = (i < p.bh) && (j < p.bw) && (hc < p.Hc) && (wc < p.Wc);
      const int oh = hc * p.o_mul + p.o_off_h, ow = wc * p.o_mul + p.o_off_w;
      t.off = t.valid ? ((((long long)n * p.Hout + oh) * p.Wout + ow) * p.Nout + t.nblk * BN) : 0;
      t.zero_out = false;
      t.valid1 = t.valid && (ow + 1 < p.Wout);
      return t;
    };
    const int sub_chunks = p.sub_n ? (p.sub_n >> 5) : (1 << 30);    // chunks >= sub_chunks belong to the odd output pixel
    auto chunk_ok = [&](const EpiTile& t, int c) { return (c_first + c) < sub_chunks ? t.valid : t.valid1; };
    auto issue = [&](EpiAux& A, const EpiTile& t, int c) {
      if (!chunk_ok(t, c)) return;
      const long long o = t.off + (c_first + c) * 32;
      if (pa && !(p.res_sub0 && (c_first + c) >= sub_chunks)) {
#pragma unroll
        for (int g = 0; g < 4; ++g) A.a[g] = *reinterpret_cast<const uint4*>(pa + o + g * 8);
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) A.b[g] = *reinterpret_cast<const uint4*>(pb + o + g * 8);
      if (pc) {
#pragma unroll
        for (int g = 0; g < 4; ++g) A.c[g] = *reinterpret_cast<const uint4*>(pc + o + g * 8);
      }
    };

    int tile = SPLIT ? (int)blockIdx.x : (int)(blockIdx.x + group * gridDim.x);
    const int bn_eff = p.sub_n ? p.sub_n : BN;            // statistics channels of this CTA
    float* box = reinterpret_cast<float*>(myscr);
    if (tile >= p.total_tiles) {
      if (pc && stats_use_mailbox(p))
        for (int e = lane; e < 2 * bn_eff; e += 32) box[e] = 0.f;
      return;
    }
    EpiTile cur = tile_info(tile), nxt = cur;
    EpiAux aux[DEPTH];
    issue(aux[0], cur, 0);
    for (; tile < p.total_tiles; tile += tstep) {
      const bool has_next = tile + tstep < p.total_tiles;
      if (has_next) nxt = tile_info(tile + tstep);
      if (pc && stat_blk != cur.nblk) {
        if (stat_blk >= 0) {
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            const int sch = p.sub_n ? (((c_first + c) * 32) & (p.sub_n - 1)) + lane : stat_blk * BN + (c_first + c) * 32 + lane;
            if (SVK_DBG_ATOMICS_ON) atomicAdd(&p.stats[sch], (double)s1[c]);
            if (SVK_DBG_ATOMICS_ON) atomicAdd(&p.stats[p.Nout + sch], (double)s2[c]);
            s1[c] = 0.f; s2[c] = 0.f;
          }
        }
        stat_blk = cur.nblk;
      }
      mbar_wait_t(bar_tfull + 8 * acc, aph, prof, pw);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        EpiAux& A = aux[c % DEPTH];
        const bool valid = chunk_ok(cur, c);
        const bool has_res = pa && !(p.res_sub0 && (c_first + c) >= sub_chunks);
        uint32_t r[32];
        tc_ld32(taddr + (c_first + c) * 32, r);
        uint32_t packed[16];                 // the 32 output values of this row, bf16x2, as stored
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float v8[8], k8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v8[e] = __uint_as_float(r[g * 8 + e]);
          if (has_res) {
            float t8[8];
            bf16x8_to_f32(A.a[g], t8);
#pragma unroll
            for (int e = 0; e < 8; ++e) v8[e] += t8[e];
          }
          bf16x8_to_f32(A.b[g], k8);
#pragma unroll
          for (int e = 0; e < 8; ++e) v8[e] = (valid && k8[e] > 0.f) ? v8[e] : 0.f;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            __nv_bfloat162 h = __floats2bfloat162_rn(v8[2 * e], v8[2 * e + 1]);
            packed[g * 4 + e] = *reinterpret_cast<uint32_t*>(&h);
          }
        }
        if (valid) {
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4*>(p.out + cur.off + (c_first + c) * 32 + g * 8) =
                make_uint4(packed[g * 4], packed[g * 4 + 1], packed[g * 4 + 2], packed[g * 4 + 3]);
        }
        if (pc) {
          // word e of the scratch row = (g_e, c_e) as bf16x2; one transpose, the channel-owning lane does the arithmetic
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint32_t cw[4] = {A.c[g].x, A.c[g].y, A.c[g].z, A.c[g].w};
            if (!valid) { cw[0] = cw[1] = cw[2] = cw[3] = 0u; }      // rows outside the image were never loaded
            uint32_t o8[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const uint32_t gw = packed[g * 4 + e];
              o8[2 * e] = (gw & 0xffffu) | (cw[e] << 16);
              o8[2 * e + 1] = (gw >> 16) | (cw[e] & 0xffff0000u);
            }
            uint32_t* dst = myscr + lane * SCR_STRIDE + g * 8;
            *reinterpret_cast<uint4*>(dst) = make_uint4(o8[0], o8[1], o8[2], o8[3]);
            *reinterpret_cast<uint4*>(dst + 4) = make_uint4(o8[4], o8[5], o8[6], o8[7]);
          }
        }
        // the operands of this chunk are consumed: fetch the chunk DEPTH ahead into the same registers
        if (c + DEPTH < NCH) issue(A, cur, c + DEPTH);
        else if (has_next) issue(A, nxt, c + DEPTH - NCH);
        if (pc) {
          __syncwarp();
          const int ch = p.sub_n ? (((c_first + c) * 32) & (p.sub_n - 1)) + lane : cur.nblk * BN + (c_first + c) * 32 + lane;
          const float mu = coef[ch];
          float a1 = 0.f, a2 = 0.f;
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) {
            const uint32_t wv = myscr[rr * SCR_STRIDE + lane];
            const float gv = __uint_as_float(wv << 16), cv = __uint_as_float(wv & 0xffff0000u);
            a1 += gv;
            a2 = fmaf(gv, cv - mu, a2);
          }
          __syncwarp();
          s1[c] += a1; s2[c] += a2 * coef[512 + ch];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) tempty_arrive(bar_tempty + 8 * acc, p.pair);
      acc += astep;
      if (acc >= n_tm) { acc -= n_tm; aph ^= 1u; }
      cur = nxt;
    }
    if (pc && stats_use_mailbox(p)) {
      __syncwarp();
      for (int e = lane; e < 2 * bn_eff; e += 32) box[e] = 0.f;      // a SPLIT warp only has half of the chunks
      __syncwarp();
#pragma unroll
      for (int c = 0; c < NCH; ++c) {         // column-pair mode: two chunks of this lane can be the same channel
        const int sch = p.sub_n ? (((c_first + c) * 32) & (p.sub_n - 1)) + lane : (c_first + c) * 32 + lane;
        box[sch] += s1[c];
        box[bn_eff + sch] += s2[c];
      }
    } else if (pc && stat_blk >= 0) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int sch = p.sub_n ? (((c_first + c) * 32) & (p.sub_n - 1)) + lane : stat_blk * BN + (c_first + c) * 32 + lane;
        if (SVK_DBG_ATOMICS_ON) atomicAdd(&p.stats[sch], (double)s1[c]);
        if (SVK_DBG_ATOMICS_ON) atomicAdd(&p.stats[p.Nout + sch], (double)s2[c]);
      }
    }
    if (prof) prof_flush(p.prof, 6, clock64() - pt0, pw, lane);
  }

// Store a warp's 32 x 32 fp32 accumulator chunk (lane = row, v = its 32 columns) to global rows `row_stride` floats apart as
// full 128-byte lines: the rows go through a 4 KB XOR-swizzled smem scratch and come back 4 rows x 128 B per instruction
// (4 L1 wavefronts instead of the 32 of a lane-per-row float4 store).  All lanes must call it.
__device__ __forceinline__ void store_chunk_rows(float* scr, const uint32_t (&v)[32], float* dst_row0, long long row_stride, int lane) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<uint4*>(scr + lane * 32 + ((j ^ (lane & 7)) << 2)) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
  const int jr = lane & 7;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int row = it * 4 + (lane >> 3);
    const uint4 x = *reinterpret_cast<const uint4*>(scr + row * 32 + ((jr ^ (row & 7)) << 2));
    *reinterpret_cast<uint4*>(dst_row0 + (long long)row * row_stride + jr * 4) = x;
  }
  __syncwarp();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 4-D map over an NHWC bf16 tensor; box = {ck channels, bw*es, bh*es, 1}, element strides {1, es, es, 1}.
inline int make_nhwc_map(CUtensorMap* m, const void* ptr, int N, int H, int W, int C, int ck, int bw, int bh, int es) {
  EncodeTiledFn enc = get_encode();
  SVK_REQUIRE(enc, SVK_E_DRIVER, "conv_tc: cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)ck, (cuuint32_t)(bw * es), (cuuint32_t)(bh * es), 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)es, (cuuint32_t)es, 1};
  SVK_REQUIRE(box[1] <= 256 && box[2] <= 256, SVK_E_UNSUPPORTED, "conv_tc: TMA box %ux%u too large", box[1], box[2]);
  CUtensorMapSwizzle sw = (ck == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SVK_REQUIRE(r == CUDA_SUCCESS, SVK_E_DRIVER, "conv_tc: cuTensorMapEncodeTiled(NHWC) failed with %d", (int)r);
  return 0;
}
// 2-D map over packed weights [rows][K] bf16; box = {ck, bn}.
inline int make_w_map(CUtensorMap* m, const void* ptr, long long rows, int K, int ck, int bn) {
  EncodeTiledFn enc = get_encode();
  SVK_REQUIRE(enc, SVK_E_DRIVER, "conv_tc: cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)ck, (cuuint32_t)bn};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapSwizzle sw = (ck == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SVK_REQUIRE(r == CUDA_SUCCESS, SVK_E_DRIVER, "conv_tc: cuTensorMapEncodeTiled(weights) failed with %d", (int)r);
  return 0;
}

}  // namespace
