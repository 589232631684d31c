// Per-sample gradient contraction on tcgen05 tensor cores.
//
//   G_group[m][c][kh][kw] = sum over the group's slots, q:  X[m][slot,q] * Y_tap(kh)[kw*C+c][slot,q]
//
// One persistent, warp-specialised kernel (1 CTA / SM):
//   warp 0     TMA producer: 128 x 32 (X) and BN x 32 (Y) fp32 tiles, SWIZZLE_128B, into a
//              kStages-deep shared-memory ring guarded by full/empty mbarriers
//   warp 1     allocates TMEM, then one elected lane issues tcgen05.mma kind::tf32
//              (M = 128, N = BN, K = 8; four per 32-wide k-block) into one of two TMEM
//              accumulator stages, committing completion to the ring / accumulator barriers
//   warps 2-5  epilogue: tcgen05.ld the accumulator (each warp owns its 32 TMEM lanes) and
//              either reduce squares (per-sample norms), add coalesced rows into the
//              gradient-natural buffer (clipped sum) or store the per-sample gradient.
// Because the accumulator is double-buffered the epilogue of item i overlaps the main loop
// of item i+1, which matters here: norm items have as few as one k-block.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/cslgan_b200.h"
#include "ptx.cuh"

namespace cg {

constexpr int kBM = 128;          // UMMA M (rows of X per tile)
constexpr int kBK = 32;           // fp32/tf32 elements per 128-byte swizzle row
constexpr int kMaxBN = 256;       // widest MMA N (columns of one accumulator stage)
constexpr int kStages = 4;
constexpr int kAccStages = 2;
constexpr int kTmemCols = 512;    // 2 accumulator stages x 256 fp32 columns (all of TMEM; 1 CTA / SM)
constexpr int kEpiWarps = 4;
constexpr int kThreads = 32 * (2 + kEpiWarps);
constexpr int kXTileBytes = kBM * kBK * 4;       // 16 KB
constexpr int kYTileBytes = kMaxBN * kBK * 4;    // 32 KB (ks stacked tap tiles of BN rows, ks*BN <= 256)
constexpr int kStageBytes = kXTileBytes + kYTileBytes;
constexpr int kEpiBufFloats = 32 * 33;           // per epilogue warp transpose buffer
constexpr int kSmemBytes = 1024 /*align slack*/ + kStages * kStageBytes + kEpiWarps * kEpiBufFloats * 4 + 256;

struct ContractParams {
  int M, n_mtiles;
  int BN, n_rb, KWC;
  int ks, n_khg, NT;     // taps stacked per item, tap groups, MMA N = ks*BN
  int C, KH, KW;
  int tap_row0[CG_MAX_KH];
  int tap_coloff[CG_MAX_KH];
  int nkb;
  long long x_slot_stride, y_slot_stride;
  int group_mode, n_groups;
  int slot_lo, slot_hi, spg;
  int n_seg, seg_stride;
  int epi;
  float* out;
  long long out_group_stride;
  long long n_items;
};

struct ItemCoord {
  int mt, rb, kh, g;
  int base_slot, n_seg, seg_stride;
};

__device__ __forceinline__ ItemCoord decode_item(const ContractParams& p, long long item) {
  ItemCoord c;
  c.mt = static_cast<int>(item % p.n_mtiles);
  long long t = item / p.n_mtiles;
  c.rb = static_cast<int>(t % p.n_rb);
  t /= p.n_rb;
  c.kh = static_cast<int>(t % p.n_khg) * p.ks;      // first tap of the group
  c.g = static_cast<int>(t / p.n_khg);
  if (p.group_mode == CG_GROUP_SAMPLE) {
    c.base_slot = p.slot_lo + c.g;
    c.n_seg = p.n_seg;
    c.seg_stride = p.seg_stride;
  } else {
    c.base_slot = p.slot_lo + c.g * p.spg;
    int hi = c.base_slot + p.spg;
    if (hi > p.slot_hi) hi = p.slot_hi;
    c.n_seg = hi - c.base_slot;
    c.seg_stride = 1;
  }
  return c;
}

__global__ void __launch_bounds__(kThreads, 1)
contract_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_y,
                const __grid_constant__ ContractParams p) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tiles = smem;
  float* epi_buf = reinterpret_cast<float*>(smem + kStages * kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_buf + kEpiWarps * kEpiBufFloats);
  uint64_t* full_bar = bars;                       // [kStages]
  uint64_t* empty_bar = bars + kStages;            // [kStages]
  uint64_t* acc_full = bars + 2 * kStages;         // [kAccStages]
  uint64_t* acc_empty = acc_full + kAccStages;     // [kAccStages]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kAccStages);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_y);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t sub_bytes = static_cast<uint32_t>(p.BN * kBK * 4);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const ItemCoord c = decode_item(p, item);
        const int xrow = c.mt * kBM;
        const int ntap = min(p.ks, p.KH - c.kh);             // taps really present in this group
        const uint32_t stage_tx = static_cast<uint32_t>(kXTileBytes) + static_cast<uint32_t>(ntap) * sub_bytes;
        for (int s = 0; s < c.n_seg; ++s) {
          const long long slot = c.base_slot + static_cast<long long>(s) * c.seg_stride;
          const long long xcol0 = slot * p.x_slot_stride;
          const long long ycol0 = slot * p.y_slot_stride;
          for (int kb = 0; kb < p.nkb; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* xs = tiles + stage * kStageBytes;
            uint8_t* ys = xs + kXTileBytes;
            mbar_expect_tx(&full_bar[stage], stage_tx);
            tma_load_2d(xs, &tmap_x, &full_bar[stage], static_cast<int32_t>(xcol0 + kb * kBK), xrow);
            for (int t = 0; t < ntap; ++t) {
              // tap kh = a shifted window (column offset) of the plane group starting at tap_row0[kh]
              tma_load_2d(ys + t * sub_bytes, &tmap_y, &full_bar[stage],
                          static_cast<int32_t>(ycol0 + p.tap_coloff[c.kh + t] + kb * kBK),
                          p.tap_row0[c.kh + t] + c.rb * p.BN);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(kBM, static_cast<uint32_t>(p.NT));
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (long long item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const ItemCoord c = decode_item(p, item);
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * kMaxBN);
        const int n_it = c.n_seg * p.nkb;
        for (int it = 0; it < n_it; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t xs = smem_u32(tiles + stage * kStageBytes);
          const uint64_t adesc = umma_desc_k_sw128(xs);
          const uint64_t bdesc = umma_desc_k_sw128(xs + kXTileBytes);
#pragma unroll
          for (int k = 0; k < kBK / 8; ++k) {
            // advance 8 tf32 = 32 bytes inside the 128-byte swizzle row: +2 in 16-byte units
            umma_tf32(tmem_d, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc,
                      (it > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);      // frees the smem slot once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&acc_full[acc]);           // accumulator complete
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp & 3;                   // TMEM lane quarter this warp may touch
    float* tbuf = epi_buf + (warp - 2) * kEpiBufFloats;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (long long item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const ItemCoord c = decode_item(p, item);
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(acc * kMaxBN);
      const int row = c.mt * kBM + ew * 32 + lane;          // gradient row handled by this thread
      const int col_base = c.rb * p.BN;                     // first column inside each tap group
      const int ncols = min(p.BN, p.KWC - col_base);        // valid columns of each stacked sub-tile
      const int ntap = min(p.ks, p.KH - c.kh);
      const int nt_valid = ntap * p.BN;                     // accumulator columns that hold real taps

      if (p.epi == CG_EPI_SUMSQ) {
        float ss = 0.f;
        for (int c0 = 0; c0 < nt_valid; c0 += 16) {
          float v[16];
          tmem_ld16(taddr + c0, v);
          const int within = c0 % p.BN;                     // BN is a multiple of 16: a chunk never straddles taps
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (within + j < ncols) ss = fmaf(v[j], v[j], ss);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[acc]);
        if (row >= p.M) ss = 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if (lane == 0) atomicAdd(p.out + c.g, ss);
      } else if (p.epi == CG_EPI_ACCUM) {
        // out[m][kh*KWC + col]: transpose 32 rows x 32 cols through smem so each row is one
        // coalesced 128-byte reduction
        const long long ld = static_cast<long long>(p.KH) * p.KWC;
        const int row0 = c.mt * kBM + ew * 32;
        for (int c0 = 0; c0 < nt_valid; c0 += 32) {
          float v[16];
          tmem_ld16(taddr + c0, v);
#pragma unroll
          for (int j = 0; j < 16; ++j) tbuf[lane * 33 + j] = v[j];
          if (c0 + 16 < nt_valid) {
            tmem_ld16(taddr + c0 + 16, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) tbuf[lane * 33 + 16 + j] = v[j];
          }
          __syncwarp();
          const int cj = c0 + lane;                         // accumulator column of this lane
          const int sub = cj / p.BN, within = cj - sub * p.BN;
          if (cj < nt_valid && within < ncols) {
            float* o = p.out + static_cast<long long>(c.kh + sub) * p.KWC + col_base + within;
            for (int r = 0; r < 32; ++r) {
              if (row0 + r < p.M) atomicAdd(o + static_cast<long long>(row0 + r) * ld, tbuf[r * 33 + lane]);
            }
          }
          __syncwarp();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[acc]);
      } else {
        // CG_EPI_STORE: out[g][m][c][kh][kw]
        const int khkw = p.KH * p.KW;
        float* obase = p.out + static_cast<long long>(c.g) * p.out_group_stride +
                       static_cast<long long>(row) * p.C * khkw;
        for (int c0 = 0; c0 < nt_valid; c0 += 16) {
          float v[16];
          tmem_ld16(taddr + c0, v);
          const int sub = c0 / p.BN, within = c0 - sub * p.BN;
          if (row < p.M) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (within + j < ncols) {
                const int col = col_base + within + j;
                const int kw = col / p.C;
                const int ch = col - kw * p.C;
                obase[static_cast<long long>(ch) * khkw + (c.kh + sub) * p.KW + kw] = v[j];
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[acc]);
      }
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace cg
