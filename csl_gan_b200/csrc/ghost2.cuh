// Ghost per-sample norms from the Gram of the UN-SHIFTED activation planes.
//
//   ||G_n||_F^2 = sum_{q,q'} BB[q,q'] * UU[q,q'],   UU[q,q'] = sum_taps sum_c U[q,tap,c] U[q',tap,c]
//
// ghost.cuh builds UU with one Gram k-block per (tap, 32-channel chunk): 25 taps re-read the same space-to-depth
// plane through shifted windows.  Every U[q,tap,:] is a row of the plane the tap lives in, so
//
//   UU[q,q'] = sum_taps P_plane(tap)[ a_tap(q), a_tap(q') ],   P_pl[a,b] = sum_c Y_pl[a,c] Y_pl[b,c]
//   a_tap(q) = (oh + hoff_tap) * Ws + (ow + woff_tap)          (position inside the plane, q = (oh, ow))
//
// i.e. ONE Gram per stride-residue plane (K = C, 4 planes for a stride-2 layer) and a gather-sum over the taps
// in the epilogue.  For the 5x5/stride-2 layers of the CelebA critic that is 6x fewer MMA k-blocks and operand
// bytes (the 8x8 layer: 16 + 4 instead of 100 + 4 k-blocks per sample).
//
// Tiles (K-major, SWIZZLE_128B, 128 rows x 32 channels = 16 KB, the same tile is both MMA operands):
//   BB : rows = (sample, q) of ns = 128/Q samples             -> accumulator stage at TMEM column 0 / 128
//   P  : rows = (sample, plane position) of spp samples       -> accumulator stage at TMEM column 256 / 384
//        (one 5-D TMA box {32 ch, Ws, Hs, spp slots, 1 chunk}; rows past spp*Hs*Ws are stale and never read)
// Epilogue warps copy the diagonal (same-sample) blocks of BB and P from TMEM to shared memory -- TMEM cannot be
// gathered across lanes -- release the accumulator, and do the tap gather-sum from shared memory while the
// tensor core already works on the next plane.
#pragma once
#include <cstdint>
#include <type_traits>
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/cslgan_b200.h"
#include "ghost.cuh"
#include "ptx.cuh"

namespace cg {

constexpr int kG2Stages = 4;
constexpr int kG2EpiWarps = 8;                              // two warps per TMEM lane quarter
constexpr int kG2EpiThreads = 32 * kG2EpiWarps;
constexpr int kG2Threads = 32 * (2 + kG2EpiWarps);
constexpr int kG2PPitch = 129;                              // floats per row of the P copy (bank-conflict free)
constexpr int kG2PBytes = 128 * kG2PPitch * 4;
constexpr int kG2MaxTaps = CG_MAX_KH * CG_MAX_KH;

__host__ __device__ constexpr int g2_smem_bytes(int Q) {
  // tiles | P copy | BB copy (128 rows x (Q+1)) | plane-position table (Q ints) | per-sample sums (128) | barriers
  return 1024 + kG2Stages * kGTileBytes + kG2PBytes + 128 * (Q + 1) * 4 + Q * 4 + 128 * 4 + 256;
}

struct Ghost2Params {
  int Q, ns, Wo;               // window positions per sample, samples per item (ns * Q == 128), window row length
  int O, C;                    // contraction extents of the two Grams
  int n_planes, Hs, Ws, npos;  // stride-residue planes, plane extent, positions per plane (<= 128)
  int spp, n_sub;              // samples per P tile (spp * npos <= 128, spp <= ns), P tiles per plane and item
  int plane_tap0[CG_MAX_KH * CG_MAX_KH + 1];   // taps of plane pl: tap_shift[plane_tap0[pl] .. plane_tap0[pl+1])
  int tap_shift[kG2MaxTaps];   // hoff * Ws + woff of every tap, grouped by plane
  int slot0, n_slots;
  int n_items;                 // ceil(n_slots / ns)
  float* norm2;                // norm2[slot - slot0] += ||G_slot||^2
  const float* inv_x;          // FP16 operands: per-slot inverse staging scales (absolute slot index), else NULL
  const float* inv_y;
};

// explicit shared-space accesses with 32-bit addresses: through the lambdas below nvcc loses the address space of
// the dynamic shared-memory pointers and emits generic LD.E / ST.E with 64-bit address arithmetic (6 instructions
// per gathered product instead of 3, measured)
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kG2EpiThreads) : "memory"); }

template <bool kHalf>
__global__ void __launch_bounds__(kG2Threads, 1)
ghost2_norm_kernel(const __grid_constant__ CUtensorMap tmap_xt, const __grid_constant__ CUtensorMap tmap_yt,
                   const __grid_constant__ Ghost2Params p) {
  constexpr uint32_t kRowB = 128;                           // bytes per chunk row of a tile (32 tf32 / 64 fp16 channels)
  constexpr int kCW = kHalf ? 64 : 32;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* Ps = reinterpret_cast<float*>(tiles + kG2Stages * kGTileBytes);
  float* BBs = Ps + 128 * kG2PPitch;
  int* apos = reinterpret_cast<int*>(BBs + 128 * (p.Q + 1));
  float* nacc = reinterpret_cast<float*>(apos + p.Q);
  uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(nacc + 128) + 7) & ~uintptr_t(7));
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kG2Stages;
  uint64_t* bb_full = bars + 2 * kG2Stages;         // [2]
  uint64_t* bb_empty = bb_full + 2;                 // [2]
  uint64_t* p_full = bb_empty + 2;                  // [2]
  uint64_t* p_empty = p_full + 2;                   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_xt);
    tma_prefetch_desc(&tmap_yt);
    for (int s = 0; s < kG2Stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bb_full[s], 1); mbar_init(&bb_empty[s], kG2EpiWarps);
      mbar_init(&p_full[s], 1); mbar_init(&p_empty[s], kG2EpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  for (int q = threadIdx.x; q < p.Q; q += blockDim.x) apos[q] = (q / p.Wo) * p.Ws + (q % p.Wo);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_ob = (p.O + kCW - 1) / kCW;
  const int n_cb = (p.C + kCW - 1) / kCW;
  const uint32_t p_tile_bytes = static_cast<uint32_t>(p.spp * p.npos) * kRowB;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int s0 = p.slot0 + item * p.ns;
        for (int kb = 0; kb < n_ob; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], 128 * kRowB);
          tma_load_3d(tiles + stage * kGTileBytes, &tmap_xt, &full_bar[stage], 0, s0 * p.Q, kb);
          if (++stage == kG2Stages) { stage = 0; phase ^= 1; }
        }
        for (int sg = 0; sg < p.n_sub; ++sg)
          for (int pl = 0; pl < p.n_planes; ++pl)
            for (int cb = 0; cb < n_cb; ++cb) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              mbar_expect_tx(&full_bar[stage], p_tile_bytes);
              tma_load_5d(tiles + stage * kGTileBytes, &tmap_yt, &full_bar[stage], 0, 0, 0, s0 + sg * p.spp,
                          pl * n_cb + cb);
              if (++stage == kG2Stages) { stage = 0; phase ^= 1; }
            }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = kHalf ? umma_idesc_f16(128, 128, 0u) : umma_idesc_tf32(128, 128);
      int stage = 0; uint32_t phase = 0;
      int bb = 0; uint32_t bb_phase = 0; int pa = 0; uint32_t pa_phase = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        mbar_wait(&bb_empty[bb], bb_phase ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < n_ob; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t desc = umma_desc_k_sw128(smem_u32(tiles + stage * kGTileBytes));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_op<kHalf>(tmem_base + static_cast<uint32_t>(bb * 128), desc + static_cast<uint64_t>(2 * k),
                           desc + static_cast<uint64_t>(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == kG2Stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&bb_full[bb]);
        if (++bb == 2) { bb = 0; bb_phase ^= 1; }
        for (int t = 0; t < p.n_sub * p.n_planes; ++t) {
          mbar_wait(&p_empty[pa], pa_phase ^ 1);
          tc_fence_after();
          for (int cb = 0; cb < n_cb; ++cb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint64_t desc = umma_desc_k_sw128(smem_u32(tiles + stage * kGTileBytes));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_op<kHalf>(tmem_base + static_cast<uint32_t>(256 + pa * 128), desc + static_cast<uint64_t>(2 * k),
                             desc + static_cast<uint64_t>(2 * k), idesc, (cb > 0 || k > 0) ? 1u : 0u);
            umma_commit(&empty_bar[stage]);
            if (++stage == kG2Stages) { stage = 0; phase ^= 1; }
          }
          umma_commit(&p_full[pa]);
          if (++pa == 2) { pa = 0; pa_phase ^= 1; }
        }
      }
    }
  } else {
    const int ew = warp & 3;                       // TMEM lanes 32*ew .. 32*ew + 31
    const int half = (warp - 2) >> 2;              // the two warps of a lane quarter split the column chunks
    const int r = ew * 32 + lane;                  // accumulator row of this thread
    const int te = half * 128 + r;                 // epilogue thread index 0..255 for the gather work split
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
    const int Qp = p.Q + 1;
    constexpr int kShiftMul = kG2PPitch + 1;
    int bb = 0; uint32_t bb_phase = 0; int pa = 0; uint32_t pa_phase = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      // ---- BB: diagonal Q x Q blocks -> BBs[(sample, q)][q'] ----
      mbar_wait(&bb_full[bb], bb_phase);
      tc_fence_after();
      {
        const int s = r / p.Q;
        const uint32_t bb_w = smem_u32(BBs + r * Qp);
        const int wlo = ((ew * 32) / p.Q) * p.Q;                    // union of the warp's diagonal blocks
        const int whi = ((ew * 32 + 31) / p.Q + 1) * p.Q;
        for (int c0 = (wlo / 16) * 16 + 16 * half; c0 < whi; c0 += 32) {
          float v[16];
          tmem_ld16(lane_addr + static_cast<uint32_t>(bb * 128 + c0), v);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int c = c0 + j - s * p.Q;
            if (c >= 0 && c < p.Q) sts_f32(bb_w + 4u * static_cast<uint32_t>(c), v[j]);
          }
        }
      }
      if (te < 128) nacc[te] = 0.f;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bb_empty[bb]);
      if (++bb == 2) { bb = 0; bb_phase ^= 1; }
      epi_bar_sync();

      for (int sg = 0; sg < p.n_sub; ++sg) {
        const int m = min(p.spp, p.ns - sg * p.spp);                // samples in this P tile
        const int rows = m * p.npos;
        // gather work split: (sample, q) pairs x R ranges of q' (R a power of two, ranges of equal length)
        const int npairs = m * p.Q;
        int R = 1;
        while (2 * R * npairs <= kG2EpiThreads && 2 * R <= p.Q) R *= 2;
        const int pr = te % npairs, rr = te / npairs;
        const bool worker = rr < R;
        const int gs = pr / p.Q, gq = pr - gs * p.Q;
        const int len = p.Q / R;
        const int qa = rr * len;
        const float* brow = BBs + ((sg * p.spp + gs) * p.Q + gq) * Qp + qa;
        const int* arow = apos + qa;
        const float* pbase = Ps + (gs * p.npos + apos[gq]) * kG2PPitch + gs * p.npos;
        float acc = 0.f;
        // one plane: wait for its Gram, copy the diagonal npos x npos blocks to Ps[(sample, a)][(sample, b)],
        // release the accumulator
        auto copy_plane = [&]() {
          mbar_wait(&p_full[pa], pa_phase);
          tc_fence_after();
          if (ew * 32 < rows) {                                     // warp-uniform: some row of this warp is live
            const int sp = r / p.npos;
            const int lo_s = (ew * 32) / p.npos;
            const int hi_s = min((ew * 32 + 31) / p.npos, m - 1);
            const int clo = lo_s * p.npos, chi = (hi_s + 1) * p.npos;
            const int mylo = sp * p.npos, myhi = mylo + p.npos;
            const uint32_t prow_w = smem_u32(Ps + r * kG2PPitch);
            for (int c0 = (clo / 16) * 16 + 16 * half; c0 < chi; c0 += 32) {
              float v[16];
              tmem_ld16(lane_addr + static_cast<uint32_t>(256 + pa * 128 + c0), v);
              if (r < rows) {
                if (c0 >= mylo && c0 + 16 <= myhi) {                // interior chunk: no per-element predicates
#pragma unroll
                  for (int j = 0; j < 16; ++j) sts_f32(prow_w + 4u * static_cast<uint32_t>(c0 + j), v[j]);
                } else {
#pragma unroll
                  for (int j = 0; j < 16; ++j) {
                    const int c = c0 + j;
                    if (c >= mylo && c < myhi) sts_f32(prow_w + 4u * static_cast<uint32_t>(c), v[j]);
                  }
                }
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_empty[pa]);
          if (++pa == 2) { pa = 0; pa_phase ^= 1; }
          epi_bar_sync();
        };
        // gather-sum with the thread's whole q' range in registers: the N row pointers and the N running sums
        // live across the planes of the sample group, every tap is N independent shared-memory loads
        auto run_planes = [&](auto nconst) {
          constexpr int N = decltype(nconst)::value;
          uint32_t ptr[N];
          float sum[N];
#pragma unroll
          for (int i = 0; i < N; ++i) { ptr[i] = smem_u32(pbase + (worker ? arow[i] : 0)); sum[i] = 0.f; }
          for (int pl = 0; pl < p.n_planes; ++pl) {
            copy_plane();
            if (worker) {
              const int t0 = p.plane_tap0[pl], t1 = p.plane_tap0[pl + 1];
              for (int t = t0; t < t1; ++t) {
                const uint32_t sh = static_cast<uint32_t>(p.tap_shift[t] * kShiftMul * 4);
#pragma unroll
                for (int i = 0; i < N; ++i) sum[i] += lds_f32(ptr[i] + sh);
              }
            }
            epi_bar_sync();                                         // Ps may be overwritten by the next plane
          }
          if (worker) {
#pragma unroll
            for (int i = 0; i < N; ++i) acc = fmaf(brow[i], sum[i], acc);
          }
        };
        if (len == 16) run_planes(std::integral_constant<int, 16>{});
        else if (len == 8) run_planes(std::integral_constant<int, 8>{});
        else if (len == 4) run_planes(std::integral_constant<int, 4>{});
        else if (len == 2) run_planes(std::integral_constant<int, 2>{});
        else if (len == 1) run_planes(std::integral_constant<int, 1>{});
        else {
          // long ranges (few (sample, q) pairs per thread budget; the 8x8 layer: 64 values of q' per thread): blocks of
          // 16 values of q' with their row addresses and running sums in registers, explicit shared-memory loads (the
          // first version walked generic pointers 4 at a time: LD.E with 64-bit address arithmetic, ncu source page).
          // The product with BB is folded in per plane, so nothing but `acc` lives across planes.
          for (int pl = 0; pl < p.n_planes; ++pl) {
            copy_plane();
            if (worker) {
              const int t0 = p.plane_tap0[pl], t1 = p.plane_tap0[pl + 1];
              int i = 0;
              for (; i + 16 <= len; i += 16) {
                uint32_t ptr[16];
                float sum[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) { ptr[j] = smem_u32(pbase + arow[i + j]); sum[j] = 0.f; }
                for (int t = t0; t < t1; ++t) {
                  const uint32_t sh = static_cast<uint32_t>(p.tap_shift[t] * kShiftMul * 4);
#pragma unroll
                  for (int j = 0; j < 16; ++j) sum[j] += lds_f32(ptr[j] + sh);
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) acc = fmaf(brow[i + j], sum[j], acc);
              }
              for (; i < len; ++i) {
                const uint32_t p0 = smem_u32(pbase + arow[i]);
                float s0 = 0.f;
                for (int t = t0; t < t1; ++t) s0 += lds_f32(p0 + static_cast<uint32_t>(p.tap_shift[t] * kShiftMul * 4));
                acc = fmaf(brow[i], s0, acc);
              }
            }
            epi_bar_sync();
          }
        }
        // lanes [k*span, (k+1)*span) of a warp share (sample, range): shuffle-reduce, one shared atomic per group
        // (a float atomicAdd on shared memory is a CAS loop: 256 contending threads cost 30 % of the kernel)
        {
          const int span = p.Q < 32 ? p.Q : 32;
          for (int o = span >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
          if (worker && (lane & (span - 1)) == 0) atomicAdd(&nacc[sg * p.spp + gs], acc);
        }
      }
      epi_bar_sync();
      const int slot_rel = item * p.ns + te;
      if (te < p.ns && slot_rel < p.n_slots) {
        float v = nacc[te];
        if (kHalf) {
          const float sc = p.inv_x[p.slot0 + slot_rel] * p.inv_y[p.slot0 + slot_rel];
          v = v * sc * sc;
        }
        atomicAdd(p.norm2 + slot_rel, v);
      }
      epi_bar_sync();                                               // nacc / BBs are rewritten by the next item
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace cg
