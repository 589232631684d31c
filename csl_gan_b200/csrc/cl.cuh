// Channels-last per-sample contraction (the main path).
//
// Operands are staged once, channels innermost, TF32-rounded, with NO im2col blow-up:
//   Xt[m/32][slot*Q + q][m%32]               backprops (rows = window positions), 32-channel chunks outermost
//   Yt[plane][c/32][slot][hs][ws][c%32]      space-to-depth activations: every filter tap is a
//                                            unit-stride window, i.e. one 5-D TMA box per tap
// (chunk-major so that ONE TMA box fetches all the 32-channel chunks of a tile: box dims
//  {32 ch, window positions..., chunks} land in shared memory as [chunk][position][32 ch], which is the
//  MN-major operand layout; small boxes cost ~2x in TMA issue/latency when measured)
// Both are MN-major operands for tcgen05 (the contraction index -- window positions / samples --
// is the row index of the staged matrices), so a k-block of 32 positions is a [32 ch x 32 rows] TMA box
// per 32-channel chunk and no transposes, padding of K, or unfold planes are needed.  Linear layers
// are the same scheme with Q = 1 (a k-block = 32 samples).  This is the layout cuDNN produces for a
// channels_last critic, so capture is a single element-wise pass.
//
//   G[m][tap][c] = sum over k-blocks:  Xt[k][m] * Yt_tap[k][c]
//
// Kernel structure as contract.cuh: persistent, 1 CTA/SM, warp 0 = TMA, warp 1 = tcgen05.mma issuer
// (M = 128, N = 32*chunks <= 256, K = 8 per instruction, both operands MN-major, SWIZZLE_128B),
// warps 2-5 = epilogue over two 256-column TMEM accumulator stages.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../../include/cslgan_b200.h"
#include "ghost.cuh"   // tma_load_5d
#include "kernels.cuh" // block_sum
#include "ptx.cuh"

namespace cg {

// Operand containers: TF32 (fp32 words, 32 channels per 128-byte chunk row, SWIZZLE_128B_BASE32B) or FP16 (64 channels
// per 128-byte chunk row, SWIZZLE_128B).  Chunk rows are 128 bytes either way: the TMA unit and L2 work per row, so
// halving the row instead of doubling its channels would keep the request count and buy nothing (measured: FP16 in
// 64-byte rows ran exactly as fast as TF32).  FP16 has TF32's 10-bit mantissa; its 5-bit exponent is handled by an
// exact power-of-two scale per (sample, operand) chosen at staging time (ptx.cuh half_scale_for) and undone in the
// epilogues, so the arithmetic error is the same while operand bytes halve and the tensor-core rate doubles.
template <bool kHalf>
struct ClCfg {
  static constexpr int kCW = kHalf ? 64 : 32;                   // channels per chunk (128 bytes)
  static constexpr int kXChunks = 128 / kCW;                    // chunks of an M = 128 tile
  static constexpr int kYChunks = 256 / kCW;                    // chunks of an N <= 256 tile
  static constexpr int kMaxRows = kHalf ? 64 : 32;              // contraction rows of the largest k-block
  static constexpr int kBoxBytes = kMaxRows * 128;              // one [chunk x k-block rows] box: 4 KB / 8 KB
  static constexpr int kXBytes = kXChunks * kBoxBytes;
  static constexpr int kYBytes = kYChunks * kBoxBytes;
  static constexpr int kStageBytes = kXBytes + kYBytes;         // 48 KB for both types
  static constexpr int kStages = 4;
  static constexpr int kKRows = kHalf ? 16 : 8;                 // contraction rows per MMA instruction
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + 4 * 32 * 33 * 4 + 256;
};
constexpr int kClThreads = 32 * 6;
constexpr int kClEpiBufFloats = 32 * 33;
constexpr int kClTmemCols = 512;
constexpr int kClMaxTaps = CG_MAX_KH * CG_MAX_KH;

struct ClParams {
  int M, n_mtiles;
  int C, n_cb;                 // channels per tap (merged: KW*C) and 128-byte chunks (32 tf32 / 64 fp16 channels) per tap
  int n_taps;
  int tpt, cpt;                // N tile = tpt taps x cpt chunks (<= 256 columns); wide layers: tpt = 1
  int tiles_per_tap;           // layers wider than one tile per tap: ceil(n_cb / cpt); else 0 (tiles group whole taps)
  int n_nt;                    // N tiles
  int tap_plane[kClMaxTaps], tap_hoff[kClMaxTaps], tap_woff[kClMaxTaps];
  int Q, Wo;                   // window positions per slot, window row length
  int kb_rows;                 // contraction rows per k-block: 32, or Q (8 | Q < 32) for per-sample groups
  int kb_s;                    // slots per k-block (1 unless Q < 32 in split-K mode)
  int nkb_slot;                // k-blocks per slot (Q/32 when Q >= 32, else 1)
  int group_mode, n_groups;    // CG_GROUP_SAMPLE: group g = slot slot_lo+g ; CG_GROUP_SPLITK: units [u_lo + g*upg, ..)
  int slot_lo;
  int n_seg, seg_stride;       // CG_GROUP_SAMPLE: a group also covers slots slot_lo+g + s*seg_stride, s < n_seg
                               // (per-sample sum over passes: joint clipping of fake_i + real_i)
  long long u_lo, u_hi, upg;   // global k-block units
  int epi;
  float* out;
  long long out_group_stride, ldT;
  int KH, KW, Corig, merged;   // output index mapping for CG_EPI_STORE
  long long n_items;
  // FP16 operands: per-slot inverse staging scales (true value = staged * inv) and the scalar that undoes the
  // common power of two folded into the factor-scaled operand of the clipped sum; all NULL for TF32
  const float* inv_x;
  const float* inv_y;
  const float* out_scale;
};

struct ClItem {
  int mt, nt, g;
  long long u0, u1;            // k-block units of the first segment
  int n_seg;
  long long seg_units;         // unit distance between segments
  int tap0, ntap, cb0, ncb;    // the tile's taps [tap0, tap0+ntap) and chunks [cb0, cb0+ncb) of each
};

// (the host guarantees n_items < 2^31: all per-item arithmetic is 32-bit -- the single-thread TMA and MMA roles run
// this on their critical path, and 64-bit divisions there cost more than the k-blocks they schedule)
__device__ __forceinline__ ClItem cl_decode(const ClParams& p, long long item64) {
  ClItem c;
  const unsigned item = static_cast<unsigned>(item64);
  const unsigned t = item / static_cast<unsigned>(p.n_mtiles);
  c.mt = static_cast<int>(item - t * static_cast<unsigned>(p.n_mtiles));
  c.g = static_cast<int>(t / static_cast<unsigned>(p.n_nt));
  c.nt = static_cast<int>(t - static_cast<unsigned>(c.g) * static_cast<unsigned>(p.n_nt));
  if (p.tiles_per_tap > 0) {
    c.tap0 = c.nt / p.tiles_per_tap;
    c.ntap = 1;
    c.cb0 = (c.nt - c.tap0 * p.tiles_per_tap) * p.cpt;
    c.ncb = min(p.cpt, p.n_cb - c.cb0);
  } else {
    c.tap0 = c.nt * p.tpt;
    c.ntap = min(p.tpt, p.n_taps - c.tap0);
    c.cb0 = 0;
    c.ncb = p.n_cb;
  }
  c.n_seg = 1;
  c.seg_units = 0;
  if (p.group_mode == CG_GROUP_SAMPLE) {
    c.u0 = static_cast<long long>(p.slot_lo + c.g) * p.nkb_slot;
    c.u1 = c.u0 + p.nkb_slot;
    c.n_seg = p.n_seg;
    c.seg_units = static_cast<long long>(p.seg_stride) * p.nkb_slot;
  } else {
    c.u0 = p.u_lo + c.g * p.upg;
    c.u1 = c.u0 + p.upg;
    if (c.u1 > p.u_hi) c.u1 = p.u_hi;
  }
  return c;
}

// MN-major descriptor for 32-bit (TF32) operands.  The only shared-memory layout tcgen05 accepts for
// MN-major TF32 is SWIZZLE_128B_BASE32B: rows of 128 B (32 channels), 32-byte chunks XOR-swizzled with
// (row % 4), i.e. atoms of 4 contraction rows (512 B) -- what TMA writes with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  32-channel chunks along M/N are `lbo` bytes apart (one TMA box
// each); 4-row K groups are 512 B apart inside a box.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(1) << 61;                      // SWIZZLE_128B_BASE32B
  return d;
}
// MN-major descriptor for 16-bit operands in 64-channel chunks: rows of 128 B, SWIZZLE_128B (16-byte units
// XOR-swizzled with row % 8; what TMA writes with CU_TENSOR_MAP_SWIZZLE_128B), atoms of 8 contraction rows (1024 B);
// chunks along M/N are `lbo` bytes apart, 8-row K groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128_16b(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;                      // SWIZZLE_128B
  return d;
}
template <bool kHalf>
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  return kHalf ? umma_desc_mn_sw128_16b(smem_addr, lbo_bytes) : umma_desc_mn_sw128(smem_addr, lbo_bytes);
}
template <bool kHalf>
__host__ __device__ constexpr uint32_t umma_idesc_mn(uint32_t M, uint32_t N) {
  return kHalf ? umma_idesc_f16(M, N, 1u)
               : ((1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24));
}

__host__ __device__ constexpr uint32_t umma_idesc_tf32_mn(uint32_t M, uint32_t N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

template <bool kHalf>
__global__ void __launch_bounds__(kClThreads, 1)
cl_contract_kernel(const __grid_constant__ CUtensorMap tmap_xt, const __grid_constant__ CUtensorMap tmap_yt,
                   const __grid_constant__ ClParams p) {
  using Cfg = ClCfg<kHalf>;
  constexpr int kClStages = Cfg::kStages;
  constexpr int kClStageBytes = Cfg::kStageBytes;
  constexpr int kClXBytes = Cfg::kXBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* epi_buf = reinterpret_cast<float*>(tiles + kClStages * kClStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_buf + 4 * kClEpiBufFloats);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kClStages;
  uint64_t* acc_full = bars + 2 * kClStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_xt);
    tma_prefetch_desc(&tmap_yt);
    for (int s = 0; s < kClStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, kClTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    // lane 0 fetches the X box (all four 32-channel chunks in one op), lanes 1..ntap one tap each
    {
      int stage = 0; uint32_t phase = 0;
      for (long long item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const ClItem c = cl_decode(p, item);
        const uint32_t box_bytes = static_cast<uint32_t>(p.kb_rows * 128);
        // TMA counts whole boxes (out-of-range chunks are zero-filled but still complete_tx)
        const uint32_t tx = static_cast<uint32_t>(Cfg::kXChunks + c.ntap * p.cpt) * box_bytes;
        // per-lane tap constants (lane t+1 owns tap tap0+t)
        int my_plane = 0, my_hoff = 0, my_woff = 0;
        const bool y_lane = lane >= 1 && lane <= c.ntap;
        if (y_lane) {
          const int tap = c.tap0 + lane - 1;
          my_plane = p.tap_plane[tap] * p.n_cb + c.cb0;
          my_hoff = p.tap_hoff[tap];
          my_woff = p.tap_woff[tap];
        }
        const int n_u = static_cast<int>(c.u1 - c.u0);
        for (int sg = 0; sg < c.n_seg; ++sg) {
        // first k-block of the segment: one division; every later k-block advances (slot, q0, oh0, ow0) incrementally
        // (this loop is the producer's critical path: with divisions per k-block it ran slower than the MMAs it feeds)
        const long long ufirst = c.u0 + sg * c.seg_units;
        int slot, q0;
        if (p.kb_s > 1) { slot = static_cast<int>(ufirst) * p.kb_s; q0 = 0; }
        else {
          slot = static_cast<int>(ufirst / p.nkb_slot);
          q0 = static_cast<int>(ufirst - static_cast<long long>(slot) * p.nkb_slot) * p.kb_rows;
        }
        int oh0 = q0 / p.Wo, ow0 = q0 - oh0 * p.Wo;
        for (int iu = 0; iu < n_u; ++iu) {
          uint8_t* xs = tiles + stage * kClStageBytes;
          uint8_t* ys = xs + kClXBytes;
          if (lane == 0) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_expect_tx(&full_bar[stage], tx);
          }
          __syncwarp();
          if (lane == 0) {
            // chunks beyond M are zero-filled by TMA (no DRAM traffic)
            tma_load_3d(xs, &tmap_xt, &full_bar[stage], 0, slot * p.Q + q0, c.mt * Cfg::kXChunks);
          } else if (y_lane) {
            tma_load_5d(ys + static_cast<uint32_t>((lane - 1) * p.cpt) * box_bytes, &tmap_yt, &full_bar[stage], 0,
                        my_woff + ow0, my_hoff + oh0, slot, my_plane);
          }
          if (++stage == kClStages) { stage = 0; phase ^= 1; }
          if (p.kb_s > 1) {
            slot += p.kb_s;
          } else {
            q0 += p.kb_rows;
            if (q0 >= p.Q) { q0 = 0; oh0 = 0; ow0 = 0; ++slot; }
            else { ow0 += p.kb_rows; while (ow0 >= p.Wo) { ow0 -= p.Wo; ++oh0; } }
          }
        }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
      for (long long item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const ClItem c = cl_decode(p, item);
        const int nbv = c.ntap * c.ncb;
        const uint32_t idesc = umma_idesc_mn<kHalf>(128, static_cast<uint32_t>(Cfg::kCW * nbv));
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * 256);
        bool first = true;
        const long long n_it = (c.u1 - c.u0) * c.n_seg;
        for (long long it = 0; it < n_it; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t xs = smem_u32(tiles + stage * kClStageBytes);
          const uint32_t box_bytes = static_cast<uint32_t>(p.kb_rows * 128);
          const uint64_t adesc = umma_desc_mn<kHalf>(xs, box_bytes);
          const uint64_t bdesc = umma_desc_mn<kHalf>(xs + kClXBytes, box_bytes);
          const int n_k = p.kb_rows / Cfg::kKRows;
          // one instruction's contraction rows (8 tf32 / 16 fp16 rows of 128 B) inside every box, in 16-byte units
          constexpr uint64_t kAdv = Cfg::kKRows * 128 / 16;
          for (int k = 0; k < n_k; ++k) {
            umma_op<kHalf>(tmem_d, adesc + kAdv * k, bdesc + kAdv * k, idesc, (first && k == 0) ? 0u : 1u);
          }
          first = false;
          umma_commit(&empty_bar[stage]);
          if (++stage == kClStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&acc_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue =====================
    const int ew = warp & 3;
    float* tbuf = epi_buf + (warp - 2) * kClEpiBufFloats;
    int acc = 0; uint32_t acc_phase = 0;
    for (long long item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const ClItem c = cl_decode(p, item);
      const int nbv = c.ntap * c.ncb;
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(acc * 256);
      const int row0 = c.mt * 128 + ew * 32;
      const int row = row0 + lane;
      // FP16 operands: undo the per-sample staging scales (per-sample groups) / the common factor scale (split-K sum)
      float gscale = 1.f;
      if (kHalf) {
        if (p.group_mode == CG_GROUP_SAMPLE) gscale = p.inv_x[p.slot_lo + c.g] * p.inv_y[p.slot_lo + c.g];
        else if (p.out_scale) gscale = p.out_scale[0];
      }

      if (p.epi == CG_EPI_SUMSQ) {
        // rows >= M and channels >= C were zero-filled by TMA: no masking needed
        const float ss0 = tmem_sumsq(taddr, Cfg::kCW * nbv, 0.f);       // (kCW * nbv is a multiple of 32)
        float ss = ss0;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[acc]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if (lane == 0) atomicAdd(p.out + c.g, kHalf ? ss * gscale * gscale : ss);
      } else if (p.epi == CG_EPI_ACCUM || p.epi == CG_EPI_STORE_NATURAL) {
        // T[m][tap*C + c]: one 32-channel chunk = 128 contiguous bytes per row -> coalesced reductions / stores
        const bool store = p.epi == CG_EPI_STORE_NATURAL;
        float* const obase_nat = p.out + (store ? static_cast<long long>(c.g) * p.out_group_stride : 0);
        constexpr int kSub = Cfg::kCW / 32;                 // 32-column blocks per chunk
        for (int j = 0; j < nbv * kSub; ++j) {
          float v[16];
          tmem_ld16(taddr + j * 32, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) tbuf[lane * 33 + i] = kHalf ? v[i] * gscale : v[i];
          tmem_ld16(taddr + j * 32 + 16, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) tbuf[lane * 33 + 16 + i] = kHalf ? v[i] * gscale : v[i];
          __syncwarp();
          const int jc = j / kSub;                          // chunk of the tile
          const int tl = jc / c.ncb;
          const int tap = c.tap0 + tl, cb = c.cb0 + (jc - tl * c.ncb);
          const int ch = cb * Cfg::kCW + (j - jc * kSub) * 32 + lane;
          if (ch < p.C) {
            float* o = obase_nat + static_cast<long long>(tap) * p.C + ch;
            if (store) {
              for (int r = 0; r < 32; ++r)
                if (row0 + r < p.M) o[static_cast<long long>(row0 + r) * p.ldT] = tbuf[r * 33 + lane];
            } else {
              for (int r = 0; r < 32; ++r)
                if (row0 + r < p.M) atomicAdd(o + static_cast<long long>(row0 + r) * p.ldT, tbuf[r * 33 + lane]);
            }
          }
          __syncwarp();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[acc]);
      } else {
        // CG_EPI_STORE: out[g][m][c][kh][kw] (contiguous parameter layout)
        const int khkw = p.KH * p.KW;
        float* obase = p.out + static_cast<long long>(c.g) * p.out_group_stride +
                       static_cast<long long>(row) * p.Corig * khkw;
        for (int j = 0; j < nbv; ++j) {
          const int tl = j / c.ncb;
          const int tap = c.tap0 + tl, cb = c.cb0 + (j - tl * c.ncb);
          for (int h = 0; h < Cfg::kCW / 16; ++h) {
            float v[16];
            tmem_ld16(taddr + j * Cfg::kCW + h * 16, v);
            if (row < p.M) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int ch = cb * Cfg::kCW + h * 16 + i;
                if (ch < p.C) {
                  int kh, kw, cc;
                  if (p.merged == 2) { const int t = ch / p.Corig; cc = ch - t * p.Corig; kh = t / p.KW; kw = t - kh * p.KW; }
                  else if (p.merged) { kh = tap; kw = ch / p.Corig; cc = ch - kw * p.Corig; }
                  else { kh = tap / p.KW; kw = tap - kh * p.KW; cc = ch; }
                  obase[static_cast<long long>(cc) * khkw + kh * p.KW + kw] = kHalf ? v[i] * gscale : v[i];
                }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kClTmemCols); }
}

// ------------------------------------------------------------------------------------------
// staged element types: float (TF32-rounded fp32 words) or __half (scaled by the sample's power of two)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_elem(float* d, float v, float) { *d = round_tf32(v); }
__device__ __forceinline__ void st_elem(__half* d, float v, float sc) { *d = __float2half_rn(v * sc); }
__device__ __forceinline__ void st_elem2(float* d, float2 v, float) {
  v.x = round_tf32(v.x); v.y = round_tf32(v.y);
  *reinterpret_cast<float2*>(d) = v;
}
__device__ __forceinline__ void st_elem2(__half* d, float2 v, float sc) {
  *reinterpret_cast<uint32_t*>(d) = pack_half2(v.x * sc, v.y * sc);
}
__device__ __forceinline__ void st_elem4(float* d, float4 v, float) {
  v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w);
  *reinterpret_cast<float4*>(d) = v;
}
__device__ __forceinline__ void st_elem4(__half* d, float4 v, float sc) {
  uint2 o;
  o.x = pack_half2(v.x * sc, v.y * sc);
  o.y = pack_half2(v.z * sc, v.w * sc);
  *reinterpret_cast<uint2*>(d) = o;
}
// staging scale of sample slot: 1 for fp32 words; for fp16 the power of two that brings |mult| * absmax into
// [2^13, 2^14) (absmax: bit pattern of the sample's largest magnitude, written by the absmax kernels below)
template <typename T>
__device__ __forceinline__ float slot_scale(const unsigned int* amax, int slot, float mult) {
  if (sizeof(T) == 4) return 1.f;
  return half_scale_for(fabsf(mult) * __uint_as_float(amax[slot]));
}

// amax[slot0 + n] = max |src[n*sn + i]|, i < len: a sample that is dense in memory (NCHW- or NHWC-contiguous).
// grid (parts, B); non-negative floats order like their bit patterns, so the partial maxima meet in an atomicMax.
__global__ void absmax_dense_kernel(const float* __restrict__ src, long long sn, long long len,
                                    unsigned int* __restrict__ amax, int slot0) {
  const int n = blockIdx.y;
  const float* s = src + static_cast<long long>(n) * sn;
  const long long per = ((len + gridDim.x - 1) / gridDim.x + 3) & ~3LL;
  const long long lo = static_cast<long long>(blockIdx.x) * per;
  const long long hi = lo + per < len ? lo + per : len;
  float m = 0.f;
  if ((reinterpret_cast<uintptr_t>(s) & 15) == 0) {
    const long long hi4 = lo + ((hi - lo) & ~3LL);
    for (long long i = lo + 4LL * threadIdx.x; i < hi4; i += 4LL * blockDim.x) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(s + i));
      m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    for (long long i = hi4 + threadIdx.x; i < hi; i += blockDim.x) m = fmaxf(m, fabsf(s[i]));
  } else {
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) m = fmaxf(m, fabsf(s[i]));
  }
  m = warp_max(m);
  __shared__ float sh[32];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = m;
  __syncthreads();
  if (w == 0) {
    m = l < (blockDim.x >> 5) ? sh[l] : 0.f;
    m = warp_max(m);
    if (l == 0 && lo < hi) atomicMax(amax + slot0 + n, __float_as_uint(m));
  }
}

// the same through explicit strides (any layout; slow path)
__global__ void absmax_strided_kernel(const float* __restrict__ src, long long sn, long long sc, long long sh,
                                      long long sw, int C, int H, int W, unsigned int* __restrict__ amax, int slot0) {
  const int n = blockIdx.y;
  const float* s = src + static_cast<long long>(n) * sn;
  const long long len = static_cast<long long>(C) * H * W;
  float m = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < len;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int w = static_cast<int>(i % W);
    const long long t = i / W;
    const int h = static_cast<int>(t % H), c = static_cast<int>(t / H);
    m = fmaxf(m, fabsf(s[c * sc + h * sh + w * sw]));
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomicMax(amax + slot0 + n, __float_as_uint(m));
}

// ------------------------------------------------------------------------------------------
// channels-last staging (element-wise passes; sources addressed through explicit strides)
// ------------------------------------------------------------------------------------------
// Xt[m/32][(slot0+n)*Q + q][m%32] = tf32(scale * G[n][m][oh][ow]).  Optional per-sample column sums (bias
// gradients) and per-sample sum of squares (closed-form Linear norms).
// grid (B, ceil(Q/qpb)), block = channels rounded up to a warp (<= 256); every warp writes whole 128-byte
// chunk rows; a block walks `qpb` positions so the bias sums stay in registers.
// FP16 output (T = __half): `amax` holds the sample maxima, the staged value is v * slot_scale and inv[slot] receives
// the inverse scale (written by one thread per sample).
template <typename T>
__global__ void stage_xt_kernel(const float* __restrict__ src, long long sn, long long sm, long long sh, long long sw,
                                int M, int Wo, int Q, float scale, T* __restrict__ dst, long long rows_total,
                                int slot0, float* __restrict__ bias_rows, float* __restrict__ sumsq, int qpb,
                                const unsigned int* __restrict__ amax, float* __restrict__ inv) {
  const int n = blockIdx.x;                       // batch on grid.x: no 65535 limit
  const int q_lo = blockIdx.y * qpb, q_hi = min(q_lo + qpb, Q);
  const float* s = src + static_cast<long long>(n) * sn;
  constexpr int CW = 128 / sizeof(T);              // channels per 128-byte chunk row
  const int Mp = (M + CW - 1) / CW * CW;
  const float hsc = slot_scale<T>(amax, slot0 + n, scale);
  if (sizeof(T) == 2 && blockIdx.y == 0 && threadIdx.x == 0) inv[slot0 + n] = 1.0f / hsc;
  float ssq = 0.f;
  for (int m = threadIdx.x; m < Mp; m += blockDim.x) {
    T* d = dst + (static_cast<long long>(m / CW) * rows_total + static_cast<long long>(slot0 + n) * Q) * CW + (m % CW);
    const float* sm_ptr = s + static_cast<long long>(m) * sm;
    float bsum = 0.f;
    int oh = q_lo / Wo, ow = q_lo - oh * Wo;
    for (int q = q_lo; q < q_hi; ++q) {
      float v = 0.f;
      if (m < M) {
        v = scale * sm_ptr[oh * sh + ow * sw];
        bsum += v;
        ssq = fmaf(v, v, ssq);
      }
      st_elem(d + static_cast<long long>(q) * CW, v, hsc);
      if (++ow == Wo) { ow = 0; ++oh; }
    }
    if (bias_rows && m < M) atomicAdd(bias_rows + static_cast<long long>(slot0 + n) * M + m, bsum);
  }
  if (sumsq) {
    __shared__ float sh_red[32];
    ssq = block_sum(ssq, sh_red);
    if (threadIdx.x == 0) atomicAdd(sumsq + slot0 + n, ssq);
  }
}

// float4 variant for channels-fastest sources (sm == 1, M % 4 == 0): a thread owns 4 channels, 8 threads
// cover one 128-byte chunk row; the block walks positions in steps of (blockDim / (M/4)).
template <typename T>
__global__ void stage_xt_vec4_kernel(const float* __restrict__ src, long long sn, long long sh, long long sw, int M,
                                     int Wo, int Q, float scale, T* __restrict__ dst, long long rows_total,
                                     int slot0, float* __restrict__ bias_rows, float* __restrict__ sumsq, int qpb,
                                     const unsigned int* __restrict__ amax, float* __restrict__ inv) {
  const int n = blockIdx.x;                       // batch on grid.x: no 65535 limit
  const int q_lo = blockIdx.y * qpb, q_hi = min(q_lo + qpb, Q);
  const int mv = M >> 2;                         // channel vectors
  const int lanes_q = blockDim.x / mv;           // positions handled in parallel (>= 1)
  const int tq = threadIdx.x / mv, tm = threadIdx.x - tq * mv;
  const float* s = src + static_cast<long long>(n) * sn + 4 * tm;
  const int m = 4 * tm;
  constexpr int CW = 128 / sizeof(T);
  const float hsc = slot_scale<T>(amax, slot0 + n, scale);
  if (sizeof(T) == 2 && blockIdx.y == 0 && threadIdx.x == 0) inv[slot0 + n] = 1.0f / hsc;
  T* d = dst + (static_cast<long long>(m / CW) * rows_total + static_cast<long long>(slot0 + n) * Q) * CW + (m % CW);
  float4 bs = make_float4(0.f, 0.f, 0.f, 0.f);
  float ssq = 0.f;
  if (tq < lanes_q) {
#pragma unroll 4
    for (int q = q_lo + tq; q < q_hi; q += lanes_q) {
      const int oh = q / Wo, ow = q - oh * Wo;
      float4 v = __ldg(reinterpret_cast<const float4*>(s + oh * sh + ow * sw));
      v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
      bs.x += v.x; bs.y += v.y; bs.z += v.z; bs.w += v.w;
      ssq = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, ssq))));
      st_elem4(d + static_cast<long long>(q) * CW, v, hsc);
    }
    if (bias_rows) {
      float* b = bias_rows + static_cast<long long>(slot0 + n) * M + m;
      atomicAdd(b, bs.x); atomicAdd(b + 1, bs.y); atomicAdd(b + 2, bs.z); atomicAdd(b + 3, bs.w);
    }
  }
  if (sumsq) {
    __shared__ float sh_red[32];
    ssq = block_sum(ssq, sh_red);
    if (threadIdx.x == 0) atomicAdd(sumsq + slot0 + n, ssq);
  }
}

// Rows wider than 1024 channels with few positions (the 8192-wide Linear input): block-sized strips of
// channel vectors, every thread walks its strips over the block's positions.
template <typename T>
__global__ void stage_xt_vec4_wide_kernel(const float* __restrict__ src, long long sn, long long sh, long long sw, int M,
                                          int Wo, int Q, float scale, T* __restrict__ dst, long long rows_total,
                                          int slot0, float* __restrict__ bias_rows, float* __restrict__ sumsq, int qpb,
                                          const unsigned int* __restrict__ amax, float* __restrict__ inv) {
  const int n = blockIdx.x;                       // batch on grid.x: no 65535 limit
  const int q_lo = blockIdx.y * qpb, q_hi = min(q_lo + qpb, Q);
  const int mv = M >> 2;
  constexpr int CW = 128 / sizeof(T);
  const float hsc = slot_scale<T>(amax, slot0 + n, scale);
  if (sizeof(T) == 2 && blockIdx.y == 0 && threadIdx.x == 0) inv[slot0 + n] = 1.0f / hsc;
  float ssq = 0.f;
  for (int tm = threadIdx.x; tm < mv; tm += blockDim.x) {
    const int m = 4 * tm;
    const float* s = src + static_cast<long long>(n) * sn + m;
    T* d = dst + (static_cast<long long>(m / CW) * rows_total + static_cast<long long>(slot0 + n) * Q) * CW + (m % CW);
    float4 bs = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = q_lo; q < q_hi; ++q) {
      const int oh = q / Wo, ow = q - oh * Wo;
      float4 v = __ldg(reinterpret_cast<const float4*>(s + oh * sh + ow * sw));
      v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
      bs.x += v.x; bs.y += v.y; bs.z += v.z; bs.w += v.w;
      ssq = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, ssq))));
      st_elem4(d + static_cast<long long>(q) * CW, v, hsc);
    }
    if (bias_rows) {
      float* b = bias_rows + static_cast<long long>(slot0 + n) * M + m;
      atomicAdd(b, bs.x); atomicAdd(b + 1, bs.y); atomicAdd(b + 2, bs.z); atomicAdd(b + 3, bs.w);
    }
  }
  if (sumsq) {
    __shared__ float sh_red[32];
    ssq = block_sum(ssq, sh_red);
    if (threadIdx.x == 0) atomicAdd(sumsq + slot0 + n, ssq);
  }
}

// float2 variant for channels-fastest rows that are a multiple of 2 but not of 4 floats (the 794-wide
// conditional MNIST input): block-sized strips of channel pairs, 8-byte loads and stores.
template <typename T>
__global__ void stage_xt_vec2_kernel(const float* __restrict__ src, long long sn, long long sh, long long sw, int M,
                                     int Wo, int Q, float scale, T* __restrict__ dst, long long rows_total,
                                     int slot0, float* __restrict__ bias_rows, float* __restrict__ sumsq, int qpb,
                                     const unsigned int* __restrict__ amax, float* __restrict__ inv) {
  const int n = blockIdx.x;
  const int q_lo = blockIdx.y * qpb, q_hi = min(q_lo + qpb, Q);
  const int mv = M >> 1;
  constexpr int CW = 128 / sizeof(T);
  const float hsc = slot_scale<T>(amax, slot0 + n, scale);
  if (sizeof(T) == 2 && blockIdx.y == 0 && threadIdx.x == 0) inv[slot0 + n] = 1.0f / hsc;
  float ssq = 0.f;
  for (int tm = threadIdx.x; tm < mv; tm += blockDim.x) {
    const int m = 2 * tm;
    const float* s = src + static_cast<long long>(n) * sn + m;
    T* d = dst + (static_cast<long long>(m / CW) * rows_total + static_cast<long long>(slot0 + n) * Q) * CW + (m % CW);
    float2 bs = make_float2(0.f, 0.f);
    for (int q = q_lo; q < q_hi; ++q) {
      const int oh = q / Wo, ow = q - oh * Wo;
      float2 v = __ldg(reinterpret_cast<const float2*>(s + oh * sh + ow * sw));
      v.x *= scale; v.y *= scale;
      bs.x += v.x; bs.y += v.y;
      ssq = fmaf(v.x, v.x, fmaf(v.y, v.y, ssq));
      st_elem2(d + static_cast<long long>(q) * CW, v, hsc);
    }
    if (bias_rows) {
      float* b = bias_rows + static_cast<long long>(slot0 + n) * M + m;
      atomicAdd(b, bs.x); atomicAdd(b + 1, bs.y);
    }
  }
  // the chunk-row tail [M, round_up(M, 32)) must read as zero for the contraction: written once at allocation
  // (torch.zeros) and never touched by any staging kernel
  if (sumsq) {
    __shared__ float sh_red[32];
    ssq = block_sum(ssq, sh_red);
    if (threadIdx.x == 0) atomicAdd(sumsq + slot0 + n, ssq);
  }
}

// Q = 1 sources with contiguous rows (Linear layers: [B][M] activations / backprops): kTPR threads per row -- one
// thread (M <= 16), one warp, or a whole block (M >= 2048) -- coalesced 16 / 8 / 4-byte loads, shuffle reductions
// for the row maximum (FP16 scale) and the sum of squares, and the per-sample bias gradient -- which for Q = 1 is the
// scaled row itself -- stored directly: no atomics, no memsets, no separate absmax pass (the second sweep over the
// row hits L1/L2).
template <typename T, int kTPR>
__global__ void __launch_bounds__(256)
stage_rows_cl_kernel(const float* __restrict__ src, long long sn, int B, int M, float scale, T* __restrict__ dst,
                     long long rows_total, int slot0, float* __restrict__ bias_rows, float* __restrict__ sumsq,
                     float* __restrict__ inv) {
  __shared__ float sh_red[8];
  const long long gtid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int n = static_cast<int>(gtid / kTPR);
  const int lane = static_cast<int>(gtid - static_cast<long long>(n) * kTPR);      // position inside the row's team
  if (kTPR < 256 && n >= B) return;                 // (block-per-row: every thread of the block has the same n < B)
  const float* s = src + static_cast<long long>(n) * sn;
  const int slot = slot0 + n;
  constexpr int CW = 128 / sizeof(T);
  // vector width of the sweeps: 4 floats when rows are 16-byte aligned, 2 when 8-byte aligned, else 1
  const int vw = kTPR == 1 ? 1
               : ((M & 3) == 0 && (reinterpret_cast<uintptr_t>(s) & 15) == 0) ? 4
               : ((M & 1) == 0 && (reinterpret_cast<uintptr_t>(s) & 7) == 0) ? 2 : 1;
  auto team_reduce = [&](float v, bool is_max) -> float {
    if (kTPR == 1) return v;
    v = is_max ? warp_max(v) : warp_sum(v);
    if (kTPR == 32) return v;
    __syncthreads();                                // sh_red reuse
    if ((threadIdx.x & 31) == 0) sh_red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = sh_red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) r = is_max ? fmaxf(r, sh_red[i]) : r + sh_red[i];
    return r;
  };
  float hsc = 1.f;
  if (sizeof(T) == 2) {
    float mx = 0.f;
    if (vw == 4) {
      for (int m = 4 * lane; m < M; m += 4 * kTPR) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(s + m));
        mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
      }
    } else if (vw == 2) {
      for (int m = 2 * lane; m < M; m += 2 * kTPR) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(s + m));
        mx = fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y)));
      }
    } else {
      for (int m = lane; m < M; m += kTPR) mx = fmaxf(mx, fabsf(__ldg(s + m)));
    }
    mx = team_reduce(mx, true);
    hsc = half_scale_for(fabsf(scale) * mx);
    if (lane == 0) inv[slot] = 1.0f / hsc;
  }
  T* d = dst + static_cast<long long>(slot) * CW;
  const long long chunk = rows_total * CW;
  float* brow = bias_rows ? bias_rows + static_cast<long long>(slot) * M : nullptr;
  float ssq = 0.f;
  if (vw == 4) {
    const bool b4 = brow && (reinterpret_cast<uintptr_t>(brow) & 15) == 0;
    for (int m = 4 * lane; m < M; m += 4 * kTPR) {
      float4 v = __ldg(reinterpret_cast<const float4*>(s + m));
      v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
      ssq = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, ssq))));
      if (b4) *reinterpret_cast<float4*>(brow + m) = v;
      else if (brow) { brow[m] = v.x; brow[m + 1] = v.y; brow[m + 2] = v.z; brow[m + 3] = v.w; }
      st_elem4(d + static_cast<long long>(m / CW) * chunk + (m % CW), v, hsc);
    }
  } else if (vw == 2) {
    const bool b2 = brow && (reinterpret_cast<uintptr_t>(brow) & 7) == 0;
    for (int m = 2 * lane; m < M; m += 2 * kTPR) {
      float2 v = __ldg(reinterpret_cast<const float2*>(s + m));
      v.x *= scale; v.y *= scale;
      ssq = fmaf(v.x, v.x, fmaf(v.y, v.y, ssq));
      if (b2) *reinterpret_cast<float2*>(brow + m) = v;
      else if (brow) { brow[m] = v.x; brow[m + 1] = v.y; }
      st_elem2(d + static_cast<long long>(m / CW) * chunk + (m % CW), v, hsc);
    }
  } else {
    for (int m = lane; m < M; m += kTPR) {
      const float v = scale * __ldg(s + m);
      ssq = fmaf(v, v, ssq);
      if (brow) brow[m] = v;
      st_elem(d + static_cast<long long>(m / CW) * chunk + (m % CW), v, hsc);
    }
  }
  // the chunk-row tail [M, round_up(M, CW)) reads as zero: written once at allocation, never touched
  if (sumsq) {
    ssq = team_reduce(ssq, false);
    if (lane == 0) sumsq[slot] = ssq;
  }
}

// ------------------------------------------------------------------------------------------
// Single-pass FP16 capture of channels-last tensors: a sample is read ONCE (4 B per element), held in registers by a
// cluster of 1-8 CTAs (16 float4 per thread, 64 KB of fp32 per CTA), the sample maximum is exchanged through
// distributed shared memory, and the scaled FP16 values are written (2 B per element).  The two-pass route (absmax
// kernel + staging kernel) reads the tensor twice; at 134 MB per activation tensor the second read misses L2.
// ------------------------------------------------------------------------------------------
constexpr int kFusedVec = 16;                       // float4 per thread
constexpr int kFusedThreads = 256;
constexpr int kFusedPerCta = kFusedVec * kFusedThreads;   // float4 per CTA

// block maximum + exchange over the cluster; every thread returns the sample maximum
__device__ __forceinline__ float cluster_sample_max(float mx, int parts) {
  __shared__ float s_warp[kFusedThreads / 32];
  __shared__ float s_cta;
  mx = warp_max(mx);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) s_warp[w] = mx;
  __syncthreads();
  if (w == 0) {
    mx = l < kFusedThreads / 32 ? s_warp[l] : 0.f;
    mx = warp_max(mx);
    if (l == 0) s_cta = mx;
  }
  if (parts == 1) {
    __syncthreads();
    return s_cta;
  }
  cluster_sync_all();                               // every CTA's s_cta is written (and visible cluster-wide)
  float g = 0.f;
  for (int r = 0; r < parts; ++r) g = fmaxf(g, ld_shared_cluster_f32(mapa_u32(&s_cta, static_cast<uint32_t>(r))));
  cluster_sync_all();                               // nobody leaves (or reuses s_cta) while a peer still reads it
  return g;
}

// Xt[m/64][(slot0+n)*Q + q][m%64] = fp16(scale * src[n][q][m] * 2^e_n); src dense [Q][M] per sample (channels fastest).
// grid = B * parts CTAs in clusters of `parts`; 256 % (M/4) == 0 so a thread keeps ONE channel vector (bias sums in
// registers, one atomic per channel and CTA).
__global__ void __launch_bounds__(kFusedThreads)
stage_xt_fused_kernel(const float* __restrict__ src, long long sn, int M, int Q, float scale, __half* __restrict__ dst,
                      long long rows_total, int slot0, float* __restrict__ bias_rows, float* __restrict__ inv, int parts) {
  const int n = blockIdx.x / parts, part = blockIdx.x - n * parts;
  const int slot = slot0 + n;
  const int mv = M >> 2;
  const int len4 = Q * mv;
  const int base = part * kFusedPerCta + threadIdx.x;
  const float4* s4 = reinterpret_cast<const float4*>(src + static_cast<long long>(n) * sn);
  float4 v[kFusedVec];
  float mx = 0.f;
#pragma unroll
  for (int j = 0; j < kFusedVec; ++j) {
    const int i = base + j * kFusedThreads;
    v[j] = i < len4 ? __ldg(s4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v[j].x), fabsf(v[j].y))), fmaxf(fabsf(v[j].z), fabsf(v[j].w)));
  }
  mx = cluster_sample_max(mx, parts);
  const float hsc = half_scale_for(fabsf(scale) * mx);
  if (part == 0 && threadIdx.x == 0) inv[slot] = 1.0f / hsc;
  const int cv = base % mv;                          // the same for every j: kFusedThreads % mv == 0
  const int m = 4 * cv;
  __half* d = dst + (static_cast<long long>(m >> 6) * rows_total + static_cast<long long>(slot) * Q) * 64 + (m & 63);
  float4 bs = make_float4(0.f, 0.f, 0.f, 0.f);
  int q = base / mv;
  const int dq = kFusedThreads / mv;
#pragma unroll
  for (int j = 0; j < kFusedVec; ++j, q += dq) {
    const int i = base + j * kFusedThreads;
    if (i < len4) {
      float4 t = v[j];
      t.x *= scale; t.y *= scale; t.z *= scale; t.w *= scale;
      bs.x += t.x; bs.y += t.y; bs.z += t.z; bs.w += t.w;
      st_elem4(d + static_cast<long long>(q) * 64, t, hsc);
    }
  }
  if (bias_rows && base < len4) {
    float* b = bias_rows + static_cast<long long>(slot) * M + m;
    atomicAdd(b, bs.x); atomicAdd(b + 1, bs.y); atomicAdd(b + 2, bs.z); atomicAdd(b + 3, bs.w);
  }
}

struct YtParams {
  int B, C, H, W;              // source [B][C][H][W] through strides
  long long sn, sc, sh_, sw_;
  int Cs, n_cb;                // staged channels (merged: KW*C) and 128-byte chunks
  int merged, KW, dw, pw;      // merged-kw (1): channel = kw*C + c, column = ow*sw - pw + kw*dw
  int dh, ph;                  // merged-all (2): channel = (kh*KW + kw)*C + c, row = oh*sh - ph + kh*dh as well
  int Hs, Ws, n_rh, n_rw, sth, stw, ah_min, aw_min;
  int rho_h[CG_MAX_KH], rho_w[CG_MAX_KH];
  float scale;
  int slot0;
  long long slot_stride, chunk_stride;   // elements: Hs*Ws*CW and n_slots_total*slot_stride
};

// Yt[plane*n_cb + c/CW][slot][hs][ws][c%CW] = staged(scale * S[n][c][h][w]) (zero outside), CW = 32 (TF32) or 64 (FP16)
// channels per 128-byte chunk row.  grid (position chunks, B, planes*n_cb).  A lane owns 4 consecutive staged
// channels of one position and issues ONE 16-byte (TF32) or 8-byte (FP16) store; CW/4 lanes cover a chunk row, so
// a warp stages 4 (TF32) or 2 (FP16) positions per step.
//   kVec  : channels-fastest, un-merged source with C % 4 == 0 -> one float4 load per lane
//   !kVec : everything else (merged thin inputs c' = kw*C + c, C % 4 != 0, channel-strided sources) ->
//           4 scalar gathers through per-lane precomputed offsets
// These kernels are instruction-issue bound, not DRAM bound (ncu: 1600 warp instructions per 16 stores with
// 64-bit index arithmetic and a division per position), so all per-position arithmetic is 32-bit and
// incremental: (hs, ws) advance by a block-uniform (dh, dw) per step.  The host checks that one sample's
// source extent fits in 31 bits.
constexpr int kYtSamples = 4;     // samples per block: the same index arithmetic, 4 independent loads in flight

template <bool kVec, typename T>
__global__ void __launch_bounds__(512)
stage_yt_kernel(const float* __restrict__ src, const __grid_constant__ YtParams p, T* __restrict__ dst,
                int pos_per_block, const unsigned int* __restrict__ amax, float* __restrict__ inv) {
  const int n0 = blockIdx.y * kYtSamples;
  const int pl = blockIdx.z / p.n_cb, chi = blockIdx.z - pl * p.n_cb;
  const int jh = pl / p.n_rw, jw = pl - jh * p.n_rw;
  constexpr int CW = 128 / sizeof(T);              // channels per chunk row
  constexpr int LPP = CW / 4;                      // lanes per position (8 or 16)
  const int l8 = threadIdx.x & (LPP - 1);
  const int sh = static_cast<int>(p.sh_), sw = static_cast<int>(p.sw_), sc = static_cast<int>(p.sc);
  const int H = p.H, W = p.W, Ws = p.Ws, sth = p.sth, stw = p.stw;
  const float scale = p.scale;
  int off[4], wk[4], hk[4];
  bool ok[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int cs = chi * CW + 4 * l8 + j;
    int c = cs; wk[j] = 0; hk[j] = 0;
    if (p.merged == 1) { const int kw = cs / p.C; c = cs - kw * p.C; wk[j] = kw * p.dw - p.pw; }
    else if (p.merged == 2) {
      const int t = cs / p.C; c = cs - t * p.C;
      const int kh = t / p.KW, kw = t - kh * p.KW;
      wk[j] = kw * p.dw - p.pw; hk[j] = kh * p.dh - p.ph;
    }
    ok[j] = cs < p.Cs;
    off[j] = ok[j] ? c * sc + wk[j] * sw + hk[j] * sh : 0;
  }
  const float* s[kYtSamples];
  T* d[kYtSamples];
  bool live[kYtSamples];
  float hsc[kYtSamples];
#pragma unroll
  for (int k = 0; k < kYtSamples; ++k) {
    live[k] = n0 + k < p.B;
    const int n = live[k] ? n0 + k : n0;
    hsc[k] = slot_scale<T>(amax, p.slot0 + n, scale);
    if (sizeof(T) == 2 && live[k] && blockIdx.x == 0 && blockIdx.z == 0 && threadIdx.x == 0) inv[p.slot0 + n] = 1.0f / hsc[k];
    s[k] = src + static_cast<long long>(n) * p.sn;
    d[k] = dst + static_cast<long long>(blockIdx.z) * p.chunk_stride + static_cast<long long>(p.slot0 + n) * p.slot_stride + 4 * l8;
  }
  const int h_base = sth * p.ah_min + p.rho_h[jh];
  const int w_base = p.merged ? 0 : stw * p.aw_min + p.rho_w[jw];      // merged: + wk[j] per channel
  const int n_pos = p.Hs * Ws;
  const int p_lo = blockIdx.x * pos_per_block, p_hi = min(p_lo + pos_per_block, n_pos);
  const int step = blockDim.x / LPP;
  const int dh = step / Ws, dw = step - dh * Ws;
  int pos = p_lo + (threadIdx.x / LPP);
  int hs = pos / Ws, ws = pos - hs * Ws;
  for (; pos < p_hi; pos += step) {
    const int h = sth * hs + h_base, w0 = stw * ws + w_base;
    const bool h_ok = h >= 0 && h < H;
    const int base = h * sh + w0 * sw;
    float4 v[kYtSamples];
    if (kVec) {
      const bool in = ok[0] && h_ok && w0 >= 0 && w0 < W;
#pragma unroll
      for (int k = 0; k < kYtSamples; ++k)
        v[k] = in ? __ldg(reinterpret_cast<const float4*>(s[k] + base + off[0])) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      bool in[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int w = w0 + wk[j], hj = h + hk[j];
        in[j] = ok[j] && hj >= 0 && hj < H && w >= 0 && w < W;
      }
#pragma unroll
      for (int k = 0; k < kYtSamples; ++k) {
        v[k].x = in[0] ? __ldg(s[k] + base + off[0]) : 0.f;
        v[k].y = in[1] ? __ldg(s[k] + base + off[1]) : 0.f;
        v[k].z = in[2] ? __ldg(s[k] + base + off[2]) : 0.f;
        v[k].w = in[3] ? __ldg(s[k] + base + off[3]) : 0.f;
      }
    }
#pragma unroll
    for (int k = 0; k < kYtSamples; ++k) {
      if (!live[k]) continue;
      float4 o;
      o.x = v[k].x * scale; o.y = v[k].y * scale; o.z = v[k].z * scale; o.w = v[k].w * scale;
      st_elem4(d[k] + pos * CW, o, hsc[k]);
    }
    ws += dw; hs += dh;
    if (ws >= Ws) { ws -= Ws; ++hs; }
  }
}

// Space-to-depth FP16 capture in one pass (un-merged, dense NHWC sample [H][W][C], C % 4 == 0, 256 % (C/4) == 0,
// H, W <= 256): every source pixel belongs to exactly one stride-residue plane; padding positions are never
// written and stay zero from allocation.  The row / column part of the destination offset is tabulated once per
// block, so the per-element work is two shared-memory lookups and an add (the first version, with divisions and
// residue searches per element, was slower than the two-pass route).
constexpr int kYtFusedMaxDim = 256;

__global__ void __launch_bounds__(kFusedThreads)
stage_yt_fused_kernel(const float* __restrict__ src, const __grid_constant__ YtParams p, __half* __restrict__ dst,
                      float* __restrict__ inv, int parts) {
  __shared__ long long row_off[kYtFusedMaxDim];     // (jh*n_rw*n_cb)*chunk_stride + hs*Ws*64, or -1
  __shared__ long long col_off[kYtFusedMaxDim];     // (jw*n_cb)*chunk_stride + ws*64, or -1
  const int n = blockIdx.x / parts, part = blockIdx.x - n * parts;
  const int slot = p.slot0 + n;
  for (int t = threadIdx.x; t < p.H + p.W; t += kFusedThreads) {
    const bool is_row = t < p.H;
    const int x = is_row ? t : t - p.H;
    const int st = is_row ? p.sth : p.stw;
    const int x0 = x - st * (is_row ? p.ah_min : p.aw_min);
    const int r = ((x0 % st) + st) % st;
    const int nr = is_row ? p.n_rh : p.n_rw;
    int j = -1;
    for (int u = 0; u < nr; ++u) if ((is_row ? p.rho_h[u] : p.rho_w[u]) == r) j = u;
    const int xs = (x0 - r) / st;
    long long off = -1;
    if (j >= 0 && xs >= 0 && xs < (is_row ? p.Hs : p.Ws))
      off = is_row ? static_cast<long long>(j) * p.n_rw * p.n_cb * p.chunk_stride + static_cast<long long>(xs) * p.Ws * 64
                   : static_cast<long long>(j) * p.n_cb * p.chunk_stride + static_cast<long long>(xs) * 64;
    (is_row ? row_off : col_off)[x] = off;
  }
  const int cvn = p.C >> 2;
  const int len4 = p.H * p.W * cvn;
  const int base = part * kFusedPerCta + threadIdx.x;
  const float4* s4 = reinterpret_cast<const float4*>(src + static_cast<long long>(n) * p.sn);
  float4 v[kFusedVec];
  float mx = 0.f;
#pragma unroll
  for (int j = 0; j < kFusedVec; ++j) {
    const int i = base + j * kFusedThreads;
    v[j] = i < len4 ? __ldg(s4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v[j].x), fabsf(v[j].y))), fmaxf(fabsf(v[j].z), fabsf(v[j].w)));
  }
  mx = cluster_sample_max(mx, parts);               // (also orders the table writes before the reads below)
  const float scale = p.scale;
  const float hsc = half_scale_for(fabsf(scale) * mx);
  if (part == 0 && threadIdx.x == 0) inv[slot] = 1.0f / hsc;
  // this thread's channel vector is the same for every j (kFusedThreads % cvn == 0); its position advances by dpos
  const int pos0 = base / cvn, c = 4 * (base - pos0 * cvn);
  const int dpos = kFusedThreads / cvn;
  int h = pos0 / p.W, w = pos0 - h * p.W;
  __half* dbase = dst + static_cast<long long>(slot) * p.slot_stride + static_cast<long long>(c >> 6) * p.chunk_stride + (c & 63);
#pragma unroll
  for (int j = 0; j < kFusedVec; ++j) {
    const int i = base + j * kFusedThreads;
    if (i < len4) {
      const long long ro = row_off[h], co = col_off[w];
      if (ro >= 0 && co >= 0) {
        float4 t4 = v[j];
        t4.x *= scale; t4.y *= scale; t4.z *= scale; t4.w *= scale;
        st_elem4(dbase + ro + co, t4, hsc);
      }
    }
    w += dpos;
    while (w >= p.W) { w -= p.W; ++h; }
  }
}

// out[n][m][p] = Xt[m/32][slot0+n][m%32] * Yt[p/32][slot0+n][p%32]   (Linear layers, Q = 1; p fastest)
// FP16 operands: times inv_x[slot] * inv_y[slot]
template <typename T>
__global__ void outer_rows_cl_kernel(const T* __restrict__ Xt, long long x_rows, const T* __restrict__ Yt,
                                     long long y_rows, int M, int P, int slot0, int B, const float* __restrict__ inv_x,
                                     const float* __restrict__ inv_y, float* __restrict__ out) {
  const long long total = static_cast<long long>(B) * M * P;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int pp = static_cast<int>(i % P);
    const long long t = i / P;
    const int m = static_cast<int>(t % M);
    const int n = static_cast<int>(t / M);
    constexpr int CW = 128 / sizeof(T);
    float v = static_cast<float>(Xt[(static_cast<long long>(m / CW) * x_rows + slot0 + n) * CW + (m % CW)]) *
              static_cast<float>(Yt[(static_cast<long long>(pp / CW) * y_rows + slot0 + n) * CW + (pp % CW)]);
    if (sizeof(T) == 2) v = v * inv_x[slot0 + n] * inv_y[slot0 + n];
    out[i] = v;
  }
}

// mult[s] = factor[s] * inv_x[s] * inv_y[s] / 2^E, out_scale[0] = 2^E (one block: a step has a few thousand slots)
__global__ void clip_mult_kernel(const float* __restrict__ factor, const float* __restrict__ inv_x,
                                 const float* __restrict__ inv_y, int slot_lo, int slot_hi, float* __restrict__ mult,
                                 float* __restrict__ out_scale) {
  __shared__ float sh[32];
  __shared__ float s_down;
  float m = 0.f;
  for (int s = slot_lo + threadIdx.x; s < slot_hi; s += blockDim.x) m = fmaxf(m, factor[s] * inv_x[s] * inv_y[s]);
  m = warp_max(m);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = m;
  __syncthreads();
  if (w == 0) {
    m = l < (blockDim.x >> 5) ? sh[l] : 0.f;
    m = warp_max(m);
    if (l == 0) {
      float up = 1.f;
      if (m > 0.f && isfinite(m)) {
        int e;
        frexpf(m, &e);                                   // m = f * 2^e, f in [0.5, 1)  ->  m < 2^e
        e = e > 126 ? 126 : (e < -126 ? -126 : e);
        up = __int_as_float((e + 127) << 23);
      }
      out_scale[0] = up;
      s_down = 1.0f / up;
    }
  }
  __syncthreads();
  const float down = s_down;
  for (int s = slot_lo + threadIdx.x; s < slot_hi; s += blockDim.x) mult[s] = factor[s] * inv_x[s] * inv_y[s] * down;
}

// Many slots (MNIST at B = 65536 has 131072): the same in two multi-block stages.  Stage 1 writes the unnormalised
// products and meets in an atomicMax on raw_max (bit pattern of a non-negative float, zeroed by the caller); stage 2
// turns the maximum into 2^E, normalises and publishes out_scale.
__global__ void clip_mult_stage1_kernel(const float* __restrict__ factor, const float* __restrict__ inv_x,
                                        const float* __restrict__ inv_y, int slot_lo, int slot_hi,
                                        float* __restrict__ mult, unsigned int* __restrict__ raw_max) {
  float m = 0.f;
  for (int s = slot_lo + blockIdx.x * blockDim.x + threadIdx.x; s < slot_hi; s += gridDim.x * blockDim.x) {
    const float v = factor[s] * inv_x[s] * inv_y[s];
    mult[s] = v;
    m = fmaxf(m, v);
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f && isfinite(m)) atomicMax(raw_max, __float_as_uint(m));
}

__global__ void clip_mult_stage2_kernel(int slot_lo, int slot_hi, float* __restrict__ mult,
                                        const unsigned int* __restrict__ raw_max, float* __restrict__ out_scale) {
  const float m = __uint_as_float(raw_max[0]);
  float up = 1.f;
  if (m > 0.f) {
    int e;
    frexpf(m, &e);
    e = e > 126 ? 126 : (e < -126 ? -126 : e);
    up = __int_as_float((e + 127) << 23);
  }
  const float down = 1.0f / up;
  for (int s = slot_lo + blockIdx.x * blockDim.x + threadIdx.x; s < slot_hi; s += gridDim.x * blockDim.x) mult[s] *= down;
  if (blockIdx.x == 0 && threadIdx.x == 0) out_scale[0] = up;
}

// dst[r][slot*stride + q] = fp16(src * mult[slot]); 8 halfs (16 B) per thread; grid (col chunks, rows)
__global__ void scale_slots_half_kernel(const __half* __restrict__ src, __half* __restrict__ dst, int rows,
                                        long long pitch, int slot_stride, int slot_lo, int slot_hi,
                                        const float* __restrict__ mult) {
  const long long cols = static_cast<long long>(slot_hi - slot_lo) * slot_stride;
  const long long col0 = static_cast<long long>(slot_lo) * slot_stride;
  const long long n8 = cols >> 3;
  for (int r = blockIdx.y; r < rows; r += gridDim.y) {
    const uint4* s = reinterpret_cast<const uint4*>(src + static_cast<long long>(r) * pitch + col0);
    uint4* d = reinterpret_cast<uint4*>(dst + static_cast<long long>(r) * pitch + col0);
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
      const float f = __ldg(mult + slot_lo + static_cast<int>((i << 3) / slot_stride));
      const uint4 v = __ldg(s + i);
      uint4 o;
      const uint32_t in[4] = {v.x, v.y, v.z, v.w};
      uint32_t ov[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&in[j]));
        ov[j] = pack_half2(t.x * f, t.y * f);
      }
      o.x = ov[0]; o.y = ov[1]; o.z = ov[2]; o.w = ov[3];
      d[i] = o;
    }
  }
}


// the same for a table of layers in ONE launch (a step scales 4 operands; the small ones are launch-latency bound)
constexpr int kScaleMaxSegs = 8;
struct ScaleSeg {
  const __half* src;
  __half* dst;
  const float* mult;
  long long pitch;
  int rows, slot_stride, slot_lo, slot_hi;
  int blk0, nblk;
};
struct ScaleParams {
  ScaleSeg seg[kScaleMaxSegs];
  int n_segs;
};

__global__ void __launch_bounds__(256)
scale_slots_half_multi_kernel(const __grid_constant__ ScaleParams p) {
  int k = 0;
#pragma unroll 1
  while (k + 1 < p.n_segs && static_cast<int>(blockIdx.x) >= p.seg[k + 1].blk0) ++k;
  const ScaleSeg& sg = p.seg[k];
  const long long col0 = static_cast<long long>(sg.slot_lo) * sg.slot_stride;
  const long long n8 = (static_cast<long long>(sg.slot_hi - sg.slot_lo) * sg.slot_stride) >> 3;
  const long long total = n8 * sg.rows;
  const long long step = static_cast<long long>(sg.nblk) * 256;
  for (long long u = static_cast<long long>(blockIdx.x - sg.blk0) * 256 + threadIdx.x; u < total; u += step) {
    const long long r = u / n8, i = u - r * n8;
    const float f = __ldg(sg.mult + sg.slot_lo + static_cast<int>((i << 3) / sg.slot_stride));
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(sg.src + r * sg.pitch + col0) + i);
    const uint32_t in[4] = {v.x, v.y, v.z, v.w};
    uint32_t ov[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&in[j]));
      ov[j] = pack_half2(t.x * f, t.y * f);
    }
    uint4 o;
    o.x = ov[0]; o.y = ov[1]; o.z = ov[2]; o.w = ov[3];
    reinterpret_cast<uint4*>(sg.dst + r * sg.pitch + col0)[i] = o;
  }
}

}  // namespace cg
