// Per-sample norms of a conv layer with the sample's backprops RESIDENT in shared memory and the filter taps read as
// SHIFTED WINDOWS of one shared-memory copy of the stride-residue plane (sm_100a, tcgen05 kind::f16).
//
// cl.cuh fetches every filter tap as its own TMA box: 25 boxes of 8 KB per 64-position k-block for a 5x5 layer, i.e.
// the space-to-depth planes (166 KB per sample for the 16x16 layer of the CelebA critics) cross the L2 -> SM path
// 4.8 times, and the backprop tile is re-fetched for each of the 7 column tiles.  ncu: that layer's norm launch is
// L2 -> SM bound with the tensor pipe 39 % active.  Here
//   * the sample's backprops (Q x M FP16 = 64 KB) are loaded ONCE and stay in shared memory for all column tiles
//     (double-buffered across samples);
//   * a column tile is one (plane, row shift) pair with all its column shifts: per k-block ONE TMA box brings the plane
//     rows the k-block's window rows touch, all Ws columns wide; a window row at column shift dw is then 16 CONSECUTIVE
//     rows of that box starting at row oh*Ws + ow0 + dw -- exactly one K = 16 MMA slab -- so the taps of the tile are
//     the same bytes seen through descriptors whose start address differs by dw rows.  The N dimension of one MMA
//     covers all column shifts at once: consecutive 64-channel chunks are LBO = 128 bytes (one row) apart.
// L2 -> SM bytes per sample: 64 KB + 10 tiles x 4 k-blocks x 9 KB = 0.43 MB instead of 1.34 MB.
#pragma once
#include "cl.cuh"

namespace cg {

constexpr int kResYStages = 4;
constexpr int kResMaxTiles = 32;
constexpr int kResXBytes = 64 * 1024;                // resident backprops of one sample (two buffers)
constexpr int kResYStride = 20 * 1024;               // one plane box (<= 160 rows of 128 B), 1024-byte aligned
constexpr int kResKb = 128;                          // window positions per k-block: 8 MMAs per barrier round trip (with 4
                                                     // the issuing thread, not the tensor pipe, set the pace: 52 % active)

struct ResParams {
  int M, Q, Wo, Ws;
  int nkb;                     // 128-position k-blocks per sample
  int kb_h;                    // window rows per k-block (128 / Wo)
  int n_tiles;
  int tile_plane[kResMaxTiles], tile_hoff[kResMaxTiles], tile_woff[kResMaxTiles], tile_ndw[kResMaxTiles];
  int slot_lo, n_groups;
  int y_bytes;                 // bytes of one plane box
  int n_cb;                    // chunks per plane (1)
  int flags;                   // unused (probe switches of the first version: the descriptor's base-offset field must
                               // stay 0 -- the swizzle is a function of the address -- and one MMA per tap is 1.6x slower)
  // clipped sum (cl_resident_sum_kernel): split-K over samples, two column tiles per item
  int tile_tap[kResMaxTiles][4];     // filter tap of column chunk j of a tile
  int n_tp;                          // tile pairs
  int slot_hi, spg;                  // samples [slot_lo, slot_hi) in groups of spg
  int C;                             // channels per tap (<= 64)
  long long ldT;                     // out[m][tap*C + c], row pitch
  const float* out_scale;
  float* out;
  const float* inv_x;
  const float* inv_y;
};

__global__ void __launch_bounds__(kClThreads, 1)
cl_resident_norm_kernel(const __grid_constant__ CUtensorMap tmap_xt, const __grid_constant__ CUtensorMap tmap_yt,
                        const __grid_constant__ ResParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* xs = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ys = xs + 2 * kResXBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ys + kResYStages * kResYStride);
  uint64_t* x_full = bars;
  uint64_t* x_empty = bars + 2;
  uint64_t* y_full = bars + 4;
  uint64_t* y_empty = y_full + kResYStages;
  uint64_t* acc_full = y_empty + kResYStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_xt);
    tma_prefetch_desc(&tmap_yt);
    for (int s = 0; s < 2; ++s) { mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], 1); }
    for (int s = 0; s < kResYStages; ++s) { mbar_init(&y_full[s], 1); mbar_init(&y_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, kClTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int x_kb_bytes = 2 * 64 * 128;               // one k-block of X: two 64-channel chunks x 64 rows

  if (warp == 0) {
    if (lane == 0) {
      int xb = 0; uint32_t xph = 0; int st = 0; uint32_t yph = 0;
      for (int g = blockIdx.x; g < p.n_groups; g += gridDim.x) {
        const int slot = p.slot_lo + g;
        mbar_wait(&x_empty[xb], xph ^ 1);
        const int n_xbox = p.Q / 64;
        mbar_expect_tx(&x_full[xb], static_cast<uint32_t>(n_xbox * x_kb_bytes));
        for (int kb = 0; kb < n_xbox; ++kb)
          tma_load_3d(xs + xb * kResXBytes + kb * x_kb_bytes, &tmap_xt, &x_full[xb], 0, slot * p.Q + kb * 64, 0);
        if (++xb == 2) { xb = 0; xph ^= 1; }
        for (int t = 0; t < p.n_tiles; ++t) {
          for (int kb = 0; kb < p.nkb; ++kb) {
            mbar_wait(&y_empty[st], yph ^ 1);
            mbar_expect_tx(&y_full[st], static_cast<uint32_t>(p.y_bytes));
            tma_load_5d(ys + st * kResYStride, &tmap_yt, &y_full[st], 0, 0, p.tile_hoff[t] + kb * p.kb_h, slot,
                        p.tile_plane[t] * p.n_cb);
            if (++st == kResYStages) { st = 0; yph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // The issuing thread is the critical path of this kernel (4 MMAs of ~96 cycles per k-block): everything that
      // does not change per k-block is hoisted -- slab rows (the only divisions), descriptor templates (a descriptor
      // is additive in its start address: +8 per 128-byte row), instruction descriptors per tile.
      int xb = 0; uint32_t xph = 0; int st = 0; uint32_t yph = 0; int acc = 0; uint32_t aph = 0;
      uint32_t slab_row[8];
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        const int ohl = (16 * s) / p.Wo, ow0 = (16 * s) - ohl * p.Wo;
        slab_row[s] = static_cast<uint32_t>(ohl * p.Ws + ow0);
      }
      const uint64_t adesc_x0 = umma_desc_mn_sw128_16b(smem_u32(xs), 64 * 128);
      const uint64_t bdesc_y0 = umma_desc_mn_sw128_16b(smem_u32(ys), 128u);    // chunks = column shifts: one row apart
      for (int g = blockIdx.x; g < p.n_groups; g += gridDim.x) {
        mbar_wait(&x_full[xb], xph);
        tc_fence_after();
        const uint64_t adesc_s = adesc_x0 + static_cast<uint64_t>(xb * (kResXBytes >> 4));
        for (int t = 0; t < p.n_tiles; ++t) {
          const uint32_t idesc = umma_idesc_f16(128, static_cast<uint32_t>(64 * p.tile_ndw[t]), 1u);
          const uint32_t woff = static_cast<uint32_t>(p.tile_woff[t]);
          uint64_t bs[8];
#pragma unroll
          for (int s = 0; s < 8; ++s) bs[s] = bdesc_y0 + static_cast<uint64_t>((slab_row[s] + woff) * 8u);
          mbar_wait(&acc_empty[acc], aph ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * 256);
          uint64_t adesc = adesc_s;
          for (int kb = 0; kb < p.nkb; ++kb) {
            mbar_wait(&y_full[st], yph);
            tc_fence_after();
            const uint64_t yo = static_cast<uint64_t>(st * (kResYStride >> 4));
#pragma unroll
            for (int s = 0; s < 8; ++s) {
              // slab s: X box s/4 (16 KB apart), rows 16*(s%4) of it
              umma_f16(tmem_d, adesc + static_cast<uint64_t>((s >> 2) * 1024 + (s & 3) * 128), bs[s] + yo, idesc,
                       (kb > 0 || s > 0) ? 1u : 0u);
            }
            umma_commit(&y_empty[st]);
            adesc += 2048;                           // next k-block of X: two 16 KB boxes
            if (++st == kResYStages) { st = 0; yph ^= 1; }
          }
          umma_commit(&acc_full[acc]);
          if (++acc == 2) { acc = 0; aph ^= 1; }
        }
        umma_commit(&x_empty[xb]);
        if (++xb == 2) { xb = 0; xph ^= 1; }
      }
    }
  } else {
    const int ew = warp & 3;
    int acc = 0; uint32_t aph = 0;
    for (int g = blockIdx.x; g < p.n_groups; g += gridDim.x) {
      const float gs = p.inv_x[p.slot_lo + g] * p.inv_y[p.slot_lo + g];
      float ss = 0.f;
      for (int t = 0; t < p.n_tiles; ++t) {
        mbar_wait(&acc_full[acc], aph);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(acc * 256);
        const int ncol = 64 * p.tile_ndw[t];
        ss = tmem_sumsq(taddr, ncol, ss);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[acc]);
        if (++acc == 2) { acc = 0; aph ^= 1; }
      }
      // rows >= M were zero-filled by TMA
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      if (lane == 0) atomicAdd(p.out + g, ss * gs * gs);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kClTmemCols); }
}


// Clipped sum of the same layers: out[m][tap][c] += sum over the samples of a group of Xc_n^T Y_n, the factor-scaled
// backprops of one sample resident in shared memory while TWO column tiles (<= 2 x 256 accumulator columns) consume
// them; an item = (tile pair, sample group), consecutive CTAs share the samples (L2 hits on the backprops).
// L2 -> SM bytes per sample: 5 x 64 KB + 10 tiles x 36 KB = 0.68 MB instead of 1.34 MB.
constexpr int kResSumEpiFloats = 32 * 33;

__global__ void __launch_bounds__(kClThreads, 1)
cl_resident_sum_kernel(const __grid_constant__ CUtensorMap tmap_xt, const __grid_constant__ CUtensorMap tmap_yt,
                       const __grid_constant__ ResParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* xs = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ys = xs + 2 * kResXBytes;
  float* epi_buf = reinterpret_cast<float*>(ys + kResYStages * kResYStride);
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_buf + 4 * kResSumEpiFloats);
  uint64_t* x_full = bars;
  uint64_t* x_empty = bars + 2;
  uint64_t* y_full = bars + 4;
  uint64_t* y_empty = y_full + kResYStages;
  uint64_t* acc_full = y_empty + kResYStages;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_xt);
    tma_prefetch_desc(&tmap_yt);
    for (int s = 0; s < 2; ++s) { mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], 1); }
    for (int s = 0; s < kResYStages; ++s) { mbar_init(&y_full[s], 1); mbar_init(&y_empty[s], 1); }
    mbar_init(acc_full, 1); mbar_init(acc_empty, 4);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, kClTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int x_kb_bytes = 2 * 64 * 128;
  const int n_items = p.n_tp * p.n_groups;

  if (warp == 0) {
    if (lane == 0) {
      int xb = 0; uint32_t xph = 0; int st = 0; uint32_t yph = 0;
      const int n_xbox = p.Q / 64;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int g = item / p.n_tp, tp = item - g * p.n_tp;
        const int t0 = 2 * tp, t1 = min(t0 + 2, p.n_tiles);
        const int s_lo = p.slot_lo + g * p.spg, s_hi = min(s_lo + p.spg, p.slot_hi);
        for (int slot = s_lo; slot < s_hi; ++slot) {
          mbar_wait(&x_empty[xb], xph ^ 1);
          mbar_expect_tx(&x_full[xb], static_cast<uint32_t>(n_xbox * x_kb_bytes));
          for (int kb = 0; kb < n_xbox; ++kb)
            tma_load_3d(xs + xb * kResXBytes + kb * x_kb_bytes, &tmap_xt, &x_full[xb], 0, slot * p.Q + kb * 64, 0);
          if (++xb == 2) { xb = 0; xph ^= 1; }
          for (int t = t0; t < t1; ++t) {
            for (int kb = 0; kb < p.nkb; ++kb) {
              mbar_wait(&y_empty[st], yph ^ 1);
              mbar_expect_tx(&y_full[st], static_cast<uint32_t>(p.y_bytes));
              tma_load_5d(ys + st * kResYStride, &tmap_yt, &y_full[st], 0, 0, p.tile_hoff[t] + kb * p.kb_h, slot,
                          p.tile_plane[t] * p.n_cb);
              if (++st == kResYStages) { st = 0; yph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int xb = 0; uint32_t xph = 0; int st = 0; uint32_t yph = 0; uint32_t aph = 0;
      uint32_t slab_row[8];
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        const int ohl = (16 * s) / p.Wo, ow0 = (16 * s) - ohl * p.Wo;
        slab_row[s] = static_cast<uint32_t>(ohl * p.Ws + ow0);
      }
      const uint64_t adesc_x0 = umma_desc_mn_sw128_16b(smem_u32(xs), 64 * 128);
      const uint64_t bdesc_y0 = umma_desc_mn_sw128_16b(smem_u32(ys), 128u);
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int g = item / p.n_tp, tp = item - g * p.n_tp;
        const int t0 = 2 * tp, t1 = min(t0 + 2, p.n_tiles);
        const int s_lo = p.slot_lo + g * p.spg, s_hi = min(s_lo + p.spg, p.slot_hi);
        uint64_t bs[2][8];
        uint32_t idesc[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int t = min(t0 + i, p.n_tiles - 1);
          idesc[i] = umma_idesc_f16(128, static_cast<uint32_t>(64 * p.tile_ndw[t]), 1u);
#pragma unroll
          for (int s = 0; s < 8; ++s)
            bs[i][s] = bdesc_y0 + static_cast<uint64_t>((slab_row[s] + static_cast<uint32_t>(p.tile_woff[t])) * 8u);
        }
        mbar_wait(acc_empty, aph ^ 1);
        tc_fence_after();
        for (int slot = s_lo; slot < s_hi; ++slot) {
          mbar_wait(&x_full[xb], xph);
          tc_fence_after();
          const uint64_t adesc_s = adesc_x0 + static_cast<uint64_t>(xb * (kResXBytes >> 4));
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            if (i < t1 - t0) {
              const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(i * 256);
              uint64_t adesc = adesc_s;
              for (int kb = 0; kb < p.nkb; ++kb) {
                mbar_wait(&y_full[st], yph);
                tc_fence_after();
                const uint64_t yo = static_cast<uint64_t>(st * (kResYStride >> 4));
#pragma unroll
                for (int s = 0; s < 8; ++s)
                  umma_f16(tmem_d, adesc + static_cast<uint64_t>((s >> 2) * 1024 + (s & 3) * 128), bs[i][s] + yo, idesc[i],
                           (slot > s_lo || kb > 0 || s > 0) ? 1u : 0u);
                umma_commit(&y_empty[st]);
                adesc += 2048;
                if (++st == kResYStages) { st = 0; yph ^= 1; }
              }
            }
          }
          umma_commit(&x_empty[xb]);
          if (++xb == 2) { xb = 0; xph ^= 1; }
        }
        umma_commit(acc_full);
        aph ^= 1;
      }
    }
  } else {
    // epilogue, once per item: T[m][tap*C + c] += tile (smem transpose -> coalesced red.add rows of 128 B)
    const int ew = warp & 3;
    float* tbuf = epi_buf + (warp - 2) * kResSumEpiFloats;
    uint32_t aph = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int g = item / p.n_tp, tp = item - g * p.n_tp;
      const int t0 = 2 * tp, t1 = min(t0 + 2, p.n_tiles);
      mbar_wait(acc_full, aph);
      tc_fence_after();
      const float gscale = p.out_scale ? p.out_scale[0] : 1.f;
      const int row0 = ew * 32;
      for (int i = 0; i < t1 - t0; ++i) {
        const int t = t0 + i;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(i * 256);
        for (int j = 0; j < 2 * p.tile_ndw[t]; ++j) {          // 32-column blocks: chunk j/2, half j%2
          float v[16];
          tmem_ld16(taddr + j * 32, v);
#pragma unroll
          for (int q = 0; q < 16; ++q) tbuf[lane * 33 + q] = v[q] * gscale;
          tmem_ld16(taddr + j * 32 + 16, v);
#pragma unroll
          for (int q = 0; q < 16; ++q) tbuf[lane * 33 + 16 + q] = v[q] * gscale;
          __syncwarp();
          const int tap = p.tile_tap[t][j >> 1];
          const int ch = (j & 1) * 32 + lane;
          if (ch < p.C) {
            float* o = p.out + static_cast<long long>(tap) * p.C + ch;
            for (int r = 0; r < 32; ++r)
              if (row0 + r < p.M) atomicAdd(o + static_cast<long long>(row0 + r) * p.ldT, tbuf[r * 33 + lane]);
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
      aph ^= 1;
      (void)g;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kClTmemCols); }
}

}  // namespace cg
