// "Staged" epilogue of the fprop / fused-dgrad tensor-core kernels (stride-1 output, BN <= 64 channels per tile).
//
// The first epilogue (tc_common.cuh) lets every thread own one pixel row of the accumulator tile and read / write its
// operands straight from / to global memory: a warp-wide 16-byte access then touches 32 different 128-byte lines, and the
// fused BatchNorm-backward data gradient (three operand streams + the output) becomes bound by L1 wavefronts, not by HBM
// or the tensor pipe (measured with SVK_PROF: the MMA warp waits for a free accumulator 30-50 % of the kernel).  Here all
// global traffic of the epilogue is TMA:
//   * the operand tiles of a tile (ReLU mask, raw conv output c, shortcut gradient) are TMA-loaded as [rows = pixels][BN
//     channels] boxes into a swizzled smem buffer owned by the epilogue group; the group's first thread issues the loads
//     for its NEXT tile as soon as the current one has left the buffer, two tiles of MMA time ahead of their use (when the
//     producer warp issued them it had to wait for the buffer and held back the activation loads behind it);
//   * an epilogue thread reads ITS row from those tiles (conflict-free thanks to the swizzle), combines it with the
//     accumulator row from TMEM and writes the bf16 result over the mask tile, in place;
//   * one thread stores the finished tile with a single TMA box store (out-of-image pixels are clipped by TMA);
//   * the BatchNorm sums (sum g, sum g*(c-mean) — or sum y, sum y^2 in the forward pass) are column sums of the staged
//     tiles: each thread owns a channel pair and walks down the rows (4-byte conflict-free reads).
#pragma once
#include "tc_common.cuh"

namespace {

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// single-thread forms (the wrappers of tc_common.cuh elect a lane of a converged warp)
__device__ __forceinline__ void mbar_expect_tx_1t(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d_1t(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// BN in {32, 64}.  Two epilogue groups (4 warps each); group g drains accumulator buffer g = every other tile of the CTA
// and owns nbuf (1 or 2) aux buffers of aux_slots tile slots each: [0] mask / output, [1] c, [2] shortcut gradient.
// aux = generic pointer to the aux region (1024-aligned), aux_s = its shared-space address.
// nbuf == 2: the operand tiles of the group's next tile are requested while the current tile is being stored and summed (a
// full tile time ahead of their use) and the read of store k-1 is awaited only inside iteration k; nbuf == 1 (no room next to
// a 64-channel resident filter) serialises store -> refill.
template <int BN>
__device__ __forceinline__ void gather_epilogue_v2(const GatherP& p, const CUtensorMap* tmOut, const CUtensorMap* tmMask,
                                                   const CUtensorMap* tmC, const CUtensorMap* tmRes, uint32_t tmem_base,
                                                   uint32_t bar_tfull, uint32_t bar_tempty, uint32_t bar_afull,
                                                   uint8_t* aux, uint32_t aux_s, int aux_slots, int nbuf, int aux_box_bytes,
                                                   const float* coef, int warp, int lane) {
  constexpr int ROWB = BN * 2;                         // bytes per pixel row of a tile
  constexpr int TILE_B = 128 * ROWB;
  constexpr int NP = BN / 2;                           // channel pairs
  constexpr int RG = 128 / NP;                         // row groups of the column pass (8 or 4)
  const int q = warp & 3;
  const int group = (warp - 2) >> 2;
  const int ngroups = ((int)blockDim.x - 64) >> 7;     // 2 or 3: group g drains accumulator buffer g = every ngroups-th tile
  const int tid_g = ((warp - 2) & 3) * 32 + lane;      // 0..127 inside the group
  const uint32_t bar_id = 1u + (uint32_t)group;
  const uint32_t buf_bytes = (uint32_t)aux_slots * TILE_B;
  uint8_t* g_aux = aux + (size_t)group * nbuf * buf_bytes;
  const uint32_t g_aux_s = aux_s + (uint32_t)group * nbuf * buf_bytes;
  const bool has_mask = p.bn_mask != nullptr, has_c = p.bn_c != nullptr, has_res = p.res != nullptr;
  const bool has_aux = has_mask;                       // forward pass: no operand tiles, the slot is only a staging buffer
  const bool stats = p.stats != nullptr;

  const int m = q * 32 + lane;                         // accumulator row m = i * pitch + j (pitch > bw: pad columns, discarded)
  const int rows_tile = p.bh * p.bw;
  const int pitch = p.pitch ? p.pitch : p.bw;
  const int i = m / pitch, j = m - i * pitch;
  const bool in_tile = (i < p.bh) && (j < p.bw);
  const int rb = i * p.bw + j;                         // its row in the staged [bh x bw] box tiles
  const uint32_t swz = (BN == 32) ? (uint32_t)((rb >> 1) & 3) : (uint32_t)(rb & 7);
  const uint32_t row_off = (uint32_t)rb * ROWB;

  // column pass: this thread sums channel pair cp over the rows rg, rg + RG, ...
  const int cp = tid_g % NP, rg = tid_g / NP;
  float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
  float mu_a = 0.f, mu_b = 0.f;
  if (has_c) { mu_a = coef[2 * cp]; mu_b = coef[2 * cp + 1]; }

  // operand tiles of `tile` -> buffer b of this group (one thread)
  auto issue_aux = [&](int tile, int b) {
    int pt = tile;
    const int tw = pt % p.tiles_w; pt /= p.tiles_w;
    const int th = pt % p.tiles_h;
    const int n = pt / p.tiles_h;
    const int h0 = th * p.bh, w0 = tw * p.bw;
    const uint32_t bar = bar_afull + 8 * (group * 2 + b);
    const uint32_t dst = g_aux_s + (uint32_t)b * buf_bytes;
    mbar_expect_tx_1t(bar, (uint32_t)((1 + (has_c ? 1 : 0) + (has_res ? 1 : 0)) * aux_box_bytes));
    tma_load_4d_1t(dst, tmMask, bar, 0, w0, h0, n);
    if (has_c) tma_load_4d_1t(dst + TILE_B, tmC, bar, 0, w0, h0, n);
    if (has_res) tma_load_4d_1t(dst + 2 * TILE_B, tmRes, bar, 0, w0, h0, n);
  };
  const int tile0 = blockIdx.x + group * gridDim.x, tstep = ngroups * gridDim.x;
  if (has_aux && tid_g == 0 && tile0 < p.total_tiles) issue_aux(tile0, 0);

  const int n_tm = p.n_tm ? p.n_tm : ngroups;          // accumulator buffers (GatherP::n_tm): this group's tiles use group, group + ngroups, ...
  int tbuf = group; uint32_t tph = 0;
  int k = 0;
  for (int tile = tile0; tile < p.total_tiles; tile += tstep, ++k) {
    int pt = tile;
    const int tw = pt % p.tiles_w; pt /= p.tiles_w;
    const int th = pt % p.tiles_h;
    const int n = pt / p.tiles_h;
    const int h0 = th * p.bh, w0 = tw * p.bw;
    const bool valid = in_tile && (h0 + i < p.Hc) && (w0 + j < p.Wc);
    const int b = (nbuf == 2) ? (k & 1) : 0;
    uint8_t* t_out = g_aux + (size_t)b * buf_bytes;
    const uint8_t* t_c = t_out + TILE_B;
    const uint8_t* t_res = t_out + 2 * TILE_B;

    mbar_wait(bar_tfull + 8 * tbuf, tph);
    tc_fence_after();
    if (has_aux) mbar_wait(bar_afull + 8 * (group * 2 + b), (uint32_t)((nbuf == 2 ? (k >> 1) : k) & 1));
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tbuf * BN);
#pragma unroll
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      tc_ld32(taddr + c * 32, r);
      if (in_tile) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t off = row_off + ((((uint32_t)(c * 4 + g)) ^ swz) << 4);
          float v8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v8[e] = __uint_as_float(r[g * 8 + e]);
          if (has_res) {
            float t8[8];
            bf16x8_to_f32(*reinterpret_cast<const uint4*>(t_res + off), t8);
#pragma unroll
            for (int e = 0; e < 8; ++e) v8[e] += t8[e];
          }
          if (has_mask) {
            float k8[8];
            bf16x8_to_f32(*reinterpret_cast<const uint4*>(t_out + off), k8);
#pragma unroll
            for (int e = 0; e < 8; ++e) v8[e] = (k8[e] > 0.f) ? v8[e] : 0.f;
          }
          uint32_t w4[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            __nv_bfloat162 h = __floats2bfloat162_rn(v8[2 * e], v8[2 * e + 1]);
            w4[e] = valid ? *reinterpret_cast<uint32_t*>(&h) : 0u;
          }
          *reinterpret_cast<uint4*>(t_out + off) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
        }
      }
    }
    // the accumulator is drained: hand it back to the MMA warp before the store / statistics
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_tempty + 8 * tbuf);
    tbuf += ngroups;
    if (tbuf >= n_tm) { tbuf -= n_tm; tph ^= 1u; }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // staged rows -> visible to the TMA store
    if (nbuf == 2 && tid_g == 0) tma_store_wait_read();              // store k-1 has read the OTHER buffer
    named_bar_sync(bar_id, 128);
    if (tid_g == 0) {
      tma_store_4d(tmOut, g_aux_s + (uint32_t)b * buf_bytes, 0, w0, h0, n);
      // after the barrier every thread is done with the other buffer (its column pass belongs to iteration k-1)
      if (nbuf == 2 && has_aux && tile + tstep < p.total_tiles) issue_aux(tile + tstep, b ^ 1);
    }
    if (stats) {
      // rows outside the image hold zeros (written above / zero-filled by TMA), rows >= rows_tile are not part of the tile
#pragma unroll 4
      for (int row = rg; row < rows_tile; row += RG) {
        const uint32_t sw = (BN == 32) ? (uint32_t)((row >> 1) & 3) : (uint32_t)(row & 7);
        const uint32_t off = (uint32_t)row * ROWB + ((((uint32_t)cp >> 2) ^ sw) << 4) + ((uint32_t)cp & 3u) * 4u;
        const uint32_t gw = *reinterpret_cast<const uint32_t*>(t_out + off);
        const float g0 = __uint_as_float(gw << 16), g1 = __uint_as_float(gw & 0xffff0000u);
        s1a += g0; s1b += g1;
        if (has_c) {
          const uint32_t cw = *reinterpret_cast<const uint32_t*>(t_c + off);
          s2a = fmaf(g0, __uint_as_float(cw << 16) - mu_a, s2a);
          s2b = fmaf(g1, __uint_as_float(cw & 0xffff0000u) - mu_b, s2b);
        } else {
          s2a = fmaf(g0, g0, s2a); s2b = fmaf(g1, g1, s2b);
        }
      }
    }
    if (nbuf == 1) {
      if (tid_g == 0) tma_store_wait_read();           // the store has read the tile: the buffer may be refilled
      named_bar_sync(bar_id, 128);
      if (tid_g == 0 && has_aux && tile + tstep < p.total_tiles) issue_aux(tile + tstep, 0);
    }
  }
  if (tid_g == 0) tma_store_wait_all();
  if (stats) {
    if (has_c) { s2a *= coef[512 + 2 * cp]; s2b *= coef[512 + 2 * cp + 1]; }
    // mailbox (tc_common.cuh): the group's first aux buffer is free once its last store has completed; thread (cp, rg)
    // leaves its four partial sums there, gather_epilogue_v2_finish adds them up after the kernel's closing barrier
    named_bar_sync(bar_id, 128);
    float* box = reinterpret_cast<float*>(g_aux);
    *reinterpret_cast<float4*>(box + 4 * tid_g) = make_float4(s1a, s1b, s2a, s2b);
  }
}

// After the CTA-wide barrier that follows the role loops: one thread per statistic sums the groups' mailboxes in a fixed
// order (double) and issues the CTA's single atomic for it.
template <int BN>
__device__ __forceinline__ void gather_epilogue_v2_finish(const GatherP& p, const uint8_t* aux, int ngroups, int aux_slots, int nbuf) {
  constexpr int NP = BN / 2, RG = 128 / NP;
  const size_t group_bytes = (size_t)nbuf * aux_slots * 128 * BN * 2;
  for (int ch = threadIdx.x; ch < 2 * BN; ch += blockDim.x) {
    const int kind = ch / BN, c = ch % BN, cp = c >> 1, e = c & 1;
    double sum = 0.0;
    for (int g = 0; g < ngroups; ++g) {
      const float* box = reinterpret_cast<const float*>(aux + (size_t)g * group_bytes);
      for (int rg = 0; rg < RG; ++rg) sum += (double)box[4 * (rg * NP + cp) + 2 * kind + e];
    }
    if (SVK_DBG_ATOMICS_ON) atomicAdd(&p.stats[kind ? p.Nout + c : c], sum);
  }
}

}  // namespace
