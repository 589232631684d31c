// Ghost per-sample norms for layers with few window positions (Q = Ho*Wo divides 128):
//
//   ||G_n||_F^2 = sum_{q,q'} (Xn^T Xn)[q,q'] * (Un^T Un)[q,q']        G_n = Xn Un^T,  Xn [O,Q], Un [P,Q]
//
// Both Gram matrices are computed on tcgen05 for ns = 128/Q samples at once: the 128 rows of a tile
// are (sample, q) pairs, the contraction runs over output channels (Xt, NHWC backprops) and over
// (filter tap, input channel) (Yt, space-to-depth channels-last activations, one 5-D TMA box per tap
// and 32-channel chunk; the same staged tensors the channels-last contraction uses, read K-major here).  A tile is BOTH MMA operands (D += T T^T), so one 16 KB TMA load feeds a full
// 128x128x32 MMA block -- 4x less operand traffic than the direct contraction -- and the epilogue reads
// only the Q x Q diagonal blocks.  Cost per sample 2*Q*128*(O + P) FLOP instead of 2*O*P*Q plus an
// O*P-element epilogue; for the 8x8 and 4x4 layers of the CelebA critic that is 2-4x fewer FLOPs and
// ~100x less epilogue.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/cslgan_b200.h"
#include "ptx.cuh"

namespace cg {

constexpr int kGStages = 6;
constexpr int kGTileBytes = 128 * 32 * 4;                 // 16 KB
constexpr int kGThreads = 32 * 6;
constexpr int kGSmemBytes = 1024 + kGStages * kGTileBytes + 256;
constexpr int kGTmemCols = 512;                           // 2 item stages x (BB 128 + UU 128)

struct GhostParams {
  int Q, ns;                 // window positions per sample, samples per tile (ns * Q == 128)
  int O, C;                  // contraction extents
  int KH, KW;
  int tap_plane[CG_MAX_KH * CG_MAX_KH];   // plane index of tap (kh*KW + kw)
  int tap_hoff[CG_MAX_KH * CG_MAX_KH];    // hs offset of the tap window
  int tap_woff[CG_MAX_KH * CG_MAX_KH];    // ws offset of the tap window
  int slot0, n_slots;        // slots [slot0, slot0 + n_slots)
  int n_items;               // ceil(n_slots / ns)
  float* norm2;              // norm2[slot - slot0] += ||G_slot||^2
  const float* inv_x;        // FP16 operands: per-slot inverse staging scales (indexed by absolute slot), else NULL
  const float* inv_y;
};

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0,
                                            int32_t c1, int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0,
                                            int32_t c1, int32_t c2, int32_t c3, int32_t c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3), "r"(c4)
      : "memory");
}

// kHalf: FP16 tiles of 128 rows x 64 channels (the same 128-byte rows, SWIZZLE_128B and 16 KB per tile as the TF32
// tiles of 32 channels; K = 16 channels per instruction instead of 8)
template <bool kHalf>
__global__ void __launch_bounds__(kGThreads, 1)
ghost_norm_kernel(const __grid_constant__ CUtensorMap tmap_xt, const __grid_constant__ CUtensorMap tmap_yt,
                  const __grid_constant__ GhostParams p) {
  constexpr uint32_t kTileTx = kGTileBytes;
  constexpr int kCW = kHalf ? 64 : 32;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + kGStages * kGTileBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kGStages;
  uint64_t* acc_full = bars + 2 * kGStages;       // [2]
  uint64_t* acc_empty = acc_full + 2;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_xt);
    tma_prefetch_desc(&tmap_yt);
    for (int s = 0; s < kGStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, kGTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_ob = (p.O + kCW - 1) / kCW;          // k-blocks of the backprop Gram
  const int n_cb = (p.C + kCW - 1) / kCW;
  const int n_taps = p.KH * p.KW;
  const int n_ub = n_taps * n_cb;                  // k-blocks of the activation Gram

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int s0 = p.slot0 + item * p.ns;
        for (int kb = 0; kb < n_ob; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], kTileTx);
          tma_load_3d(tiles + stage * kGTileBytes, &tmap_xt, &full_bar[stage], 0, s0 * p.Q, kb);
          if (++stage == kGStages) { stage = 0; phase ^= 1; }
        }
        for (int t = 0; t < n_taps; ++t) {
          for (int cb = 0; cb < n_cb; ++cb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_expect_tx(&full_bar[stage], kTileTx);
            tma_load_5d(tiles + stage * kGTileBytes, &tmap_yt, &full_bar[stage], 0, p.tap_woff[t], p.tap_hoff[t], s0,
                        p.tap_plane[t] * n_cb + cb);
            if (++stage == kGStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = kHalf ? umma_idesc_f16(128, 128, 0u) : umma_idesc_tf32(128, 128);
      int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        for (int it = 0; it < n_ob + n_ub; ++it) {
          const bool is_u = it >= n_ob;
          const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * 256 + (is_u ? 128 : 0));
          const bool first = (it == 0) || (it == n_ob);
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t desc = umma_desc_k_sw128(smem_u32(tiles + stage * kGTileBytes));
#pragma unroll
          for (int k = 0; k < 4; ++k)                     // 32 bytes of K per instruction: 8 tf32 / 16 fp16 channels
            umma_op<kHalf>(tmem_d, desc + static_cast<uint64_t>(2 * k), desc + static_cast<uint64_t>(2 * k), idesc,
                           (!first || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == kGStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&acc_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    const int ew = warp & 3;
    int acc = 0; uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const int r = ew * 32 + lane;                 // tile row = (sample in group, q)
      const int sidx = r / p.Q;
      const int col_lo = sidx * p.Q;                // diagonal block of this row's sample
      const uint32_t t0 = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(acc * 256);
      // all lanes of a warp must read the same columns: cover the union of the warp's diagonal blocks
      const int wlo = ((ew * 32) / p.Q) * p.Q;      // first column any lane of this warp needs
      const int whi = ((ew * 32 + 31) / p.Q + 1) * p.Q;
      float dot = 0.f;
      for (int c0 = (wlo / 16) * 16; c0 < whi; c0 += 16) {
        float a[16], b[16];
        tmem_ld16(t0 + c0, a);
        tmem_ld16(t0 + 128 + c0, b);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int c = c0 + j;
          if (c >= col_lo && c < col_lo + p.Q) dot = fmaf(a[j], b[j], dot);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      // reduce over the Q rows of each sample
      const int span = p.Q < 32 ? p.Q : 32;
      for (int o = span >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      const int slot_rel = item * p.ns + sidx;
      if ((lane % span) == 0 && slot_rel < p.n_slots) {
        if (kHalf) {
          const float sc = p.inv_x[p.slot0 + slot_rel] * p.inv_y[p.slot0 + slot_rel];
          dot = dot * sc * sc;
        }
        atomicAdd(p.norm2 + slot_rel, dot);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kGTmemCols); }
}

}  // namespace cg
