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
#include <cuda_runtime.h>

#include "../../include/cslgan_b200.h"
#include "ghost.cuh"   // tma_load_5d
#include "kernels.cuh" // block_sum
#include "ptx.cuh"

namespace cg {

constexpr int kClStages = 4;
constexpr int kClBoxBytes = 32 * 32 * 4;              // one [32 ch x 32 rows] box
constexpr int kClXBytes = 4 * kClBoxBytes;            // M = 128
constexpr int kClYBytes = 8 * kClBoxBytes;            // N <= 256
constexpr int kClStageBytes = kClXBytes + kClYBytes;  // 48 KB
constexpr int kClThreads = 32 * 6;
constexpr int kClEpiBufFloats = 32 * 33;
constexpr int kClSmemBytes = 1024 + kClStages * kClStageBytes + 4 * kClEpiBufFloats * 4 + 256;
constexpr int kClTmemCols = 512;
constexpr int kClMaxTaps = CG_MAX_KH * CG_MAX_KH;

struct ClParams {
  int M, n_mtiles;
  int C, n_cb;                 // channels per tap (merged: KW*C) and 32-wide chunks per tap
  int n_taps;
  int tpt, cpt;                // N tile = tpt taps x cpt chunks (tpt*cpt <= 8); n_cb >= 8: tpt = 1, cpt = 8
  int tiles_per_tap;           // n_cb >= 8: ceil(n_cb/8); else 0 (tiles group whole taps)
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
};

struct ClItem {
  int mt, nt, g;
  long long u0, u1;            // k-block units of the first segment
  int n_seg;
  long long seg_units;         // unit distance between segments
  int tap0, ntap, cb0, ncb;    // the tile's taps [tap0, tap0+ntap) and chunks [cb0, cb0+ncb) of each
};

__device__ __forceinline__ ClItem cl_decode(const ClParams& p, long long item) {
  ClItem c;
  c.mt = static_cast<int>(item % p.n_mtiles);
  long long t = item / p.n_mtiles;
  c.nt = static_cast<int>(t % p.n_nt);
  c.g = static_cast<int>(t / p.n_nt);
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

__host__ __device__ constexpr uint32_t umma_idesc_tf32_mn(uint32_t M, uint32_t N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__global__ void __launch_bounds__(kClThreads, 1)
cl_contract_kernel(const __grid_constant__ CUtensorMap tmap_xt, const __grid_constant__ CUtensorMap tmap_yt,
                   const __grid_constant__ ClParams p) {
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
        const uint32_t tx = static_cast<uint32_t>(4 + c.ntap * p.cpt) * box_bytes;
        // per-lane tap constants (lane t+1 owns tap tap0+t)
        int my_plane = 0, my_hoff = 0, my_woff = 0;
        const bool y_lane = lane >= 1 && lane <= c.ntap;
        if (y_lane) {
          const int tap = c.tap0 + lane - 1;
          my_plane = p.tap_plane[tap] * p.n_cb + c.cb0;
          my_hoff = p.tap_hoff[tap];
          my_woff = p.tap_woff[tap];
        }
        for (int sg = 0; sg < c.n_seg; ++sg)
        for (long long u = c.u0 + sg * c.seg_units; u < c.u1 + sg * c.seg_units; ++u) {
          int slot, q0;
          if (p.kb_s > 1) { slot = static_cast<int>(u) * p.kb_s; q0 = 0; }
          else { slot = static_cast<int>(u / p.nkb_slot); q0 = static_cast<int>(u - static_cast<long long>(slot) * p.nkb_slot) * p.kb_rows; }
          const int oh0 = q0 / p.Wo, ow0 = q0 - oh0 * p.Wo;
          uint8_t* xs = tiles + stage * kClStageBytes;
          uint8_t* ys = xs + kClXBytes;
          if (lane == 0) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_expect_tx(&full_bar[stage], tx);
          }
          __syncwarp();
          if (lane == 0) {
            // chunks beyond M are zero-filled by TMA (no DRAM traffic)
            tma_load_3d(xs, &tmap_xt, &full_bar[stage], 0, slot * p.Q + q0, c.mt * 4);
          } else if (y_lane) {
            tma_load_5d(ys + static_cast<uint32_t>((lane - 1) * p.cpt) * box_bytes, &tmap_yt, &full_bar[stage], 0,
                        my_woff + ow0, my_hoff + oh0, slot, my_plane);
          }
          if (++stage == kClStages) { stage = 0; phase ^= 1; }
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
        const uint32_t idesc = umma_idesc_tf32_mn(128, static_cast<uint32_t>(32 * nbv));
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
          const uint64_t adesc = umma_desc_mn_sw128(xs, box_bytes);
          const uint64_t bdesc = umma_desc_mn_sw128(xs + kClXBytes, box_bytes);
          const int n_k = p.kb_rows >> 3;
          for (int k = 0; k < n_k; ++k) {
            // 8 contraction rows = one 1024-byte group inside every box: +64 in 16-byte units
            umma_tf32(tmem_d, adesc + static_cast<uint64_t>(64 * k), bdesc + static_cast<uint64_t>(64 * k), idesc,
                      (first && k == 0) ? 0u : 1u);
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

      if (p.epi == CG_EPI_SUMSQ) {
        // rows >= M and channels >= C were zero-filled by TMA: no masking needed
        float ss = 0.f;
        for (int c0 = 0; c0 < 32 * nbv; c0 += 16) {
          float v[16];
          tmem_ld16(taddr + c0, v);
#pragma unroll
          for (int j = 0; j < 16; ++j) ss = fmaf(v[j], v[j], ss);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[acc]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if (lane == 0) atomicAdd(p.out + c.g, ss);
      } else if (p.epi == CG_EPI_ACCUM || p.epi == CG_EPI_STORE_NATURAL) {
        // T[m][tap*C + c]: one 32-channel chunk = 128 contiguous bytes per row -> coalesced reductions / stores
        const bool store = p.epi == CG_EPI_STORE_NATURAL;
        float* const obase_nat = p.out + (store ? static_cast<long long>(c.g) * p.out_group_stride : 0);
        for (int j = 0; j < nbv; ++j) {
          float v[16];
          tmem_ld16(taddr + j * 32, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) tbuf[lane * 33 + i] = v[i];
          tmem_ld16(taddr + j * 32 + 16, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) tbuf[lane * 33 + 16 + i] = v[i];
          __syncwarp();
          const int tl = j / c.ncb;
          const int tap = c.tap0 + tl, cb = c.cb0 + (j - tl * c.ncb);
          const int ch = cb * 32 + lane;
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
          for (int h = 0; h < 2; ++h) {
            float v[16];
            tmem_ld16(taddr + j * 32 + h * 16, v);
            if (row < p.M) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int ch = cb * 32 + h * 16 + i;
                if (ch < p.C) {
                  int kh, kw, cc;
                  if (p.merged) { kh = tap; kw = ch / p.Corig; cc = ch - kw * p.Corig; }
                  else { kh = tap / p.KW; kw = tap - kh * p.KW; cc = ch; }
                  obase[static_cast<long long>(cc) * khkw + kh * p.KW + kw] = v[i];
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
// channels-last staging (element-wise passes; sources addressed through explicit strides)
// ------------------------------------------------------------------------------------------
// Xt[m/32][(slot0+n)*Q + q][m%32] = tf32(scale * G[n][m][oh][ow]).  Optional per-sample column sums (bias
// gradients) and per-sample sum of squares (closed-form Linear norms).
// grid (B, ceil(Q/qpb)), block = channels rounded up to a warp (<= 256); every warp writes whole 128-byte
// chunk rows; a block walks `qpb` positions so the bias sums stay in registers.
__global__ void stage_xt_kernel(const float* __restrict__ src, long long sn, long long sm, long long sh, long long sw,
                                int M, int Wo, int Q, float scale, float* __restrict__ dst, long long rows_total,
                                int slot0, float* __restrict__ bias_rows, float* __restrict__ sumsq, int qpb) {
  const int n = blockIdx.x;                       // batch on grid.x: no 65535 limit
  const int q_lo = blockIdx.y * qpb, q_hi = min(q_lo + qpb, Q);
  const float* s = src + static_cast<long long>(n) * sn;
  const int Mp = (M + 31) & ~31;
  float ssq = 0.f;
  for (int m = threadIdx.x; m < Mp; m += blockDim.x) {
    float* d = dst + (static_cast<long long>(m >> 5) * rows_total + static_cast<long long>(slot0 + n) * Q) * 32 + (m & 31);
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
      d[static_cast<long long>(q) * 32] = round_tf32(v);
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
__global__ void stage_xt_vec4_kernel(const float* __restrict__ src, long long sn, long long sh, long long sw, int M,
                                     int Wo, int Q, float scale, float* __restrict__ dst, long long rows_total,
                                     int slot0, float* __restrict__ bias_rows, float* __restrict__ sumsq, int qpb) {
  const int n = blockIdx.x;                       // batch on grid.x: no 65535 limit
  const int q_lo = blockIdx.y * qpb, q_hi = min(q_lo + qpb, Q);
  const int mv = M >> 2;                         // channel vectors
  const int lanes_q = blockDim.x / mv;           // positions handled in parallel (>= 1)
  const int tq = threadIdx.x / mv, tm = threadIdx.x - tq * mv;
  const float* s = src + static_cast<long long>(n) * sn + 4 * tm;
  const int m = 4 * tm;
  float* d = dst + (static_cast<long long>(m >> 5) * rows_total + static_cast<long long>(slot0 + n) * Q) * 32 + (m & 31);
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
      v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w);
      *reinterpret_cast<float4*>(d + static_cast<long long>(q) * 32) = v;
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
__global__ void stage_xt_vec4_wide_kernel(const float* __restrict__ src, long long sn, long long sh, long long sw, int M,
                                          int Wo, int Q, float scale, float* __restrict__ dst, long long rows_total,
                                          int slot0, float* __restrict__ bias_rows, float* __restrict__ sumsq, int qpb) {
  const int n = blockIdx.x;                       // batch on grid.x: no 65535 limit
  const int q_lo = blockIdx.y * qpb, q_hi = min(q_lo + qpb, Q);
  const int mv = M >> 2;
  float ssq = 0.f;
  for (int tm = threadIdx.x; tm < mv; tm += blockDim.x) {
    const int m = 4 * tm;
    const float* s = src + static_cast<long long>(n) * sn + m;
    float* d = dst + (static_cast<long long>(m >> 5) * rows_total + static_cast<long long>(slot0 + n) * Q) * 32 + (m & 31);
    float4 bs = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = q_lo; q < q_hi; ++q) {
      const int oh = q / Wo, ow = q - oh * Wo;
      float4 v = __ldg(reinterpret_cast<const float4*>(s + oh * sh + ow * sw));
      v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
      bs.x += v.x; bs.y += v.y; bs.z += v.z; bs.w += v.w;
      ssq = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, ssq))));
      v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w);
      *reinterpret_cast<float4*>(d + static_cast<long long>(q) * 32) = v;
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
__global__ void stage_xt_vec2_kernel(const float* __restrict__ src, long long sn, long long sh, long long sw, int M,
                                     int Wo, int Q, float scale, float* __restrict__ dst, long long rows_total,
                                     int slot0, float* __restrict__ bias_rows, float* __restrict__ sumsq, int qpb) {
  const int n = blockIdx.x;
  const int q_lo = blockIdx.y * qpb, q_hi = min(q_lo + qpb, Q);
  const int mv = M >> 1;
  float ssq = 0.f;
  for (int tm = threadIdx.x; tm < mv; tm += blockDim.x) {
    const int m = 2 * tm;
    const float* s = src + static_cast<long long>(n) * sn + m;
    float* d = dst + (static_cast<long long>(m >> 5) * rows_total + static_cast<long long>(slot0 + n) * Q) * 32 + (m & 31);
    float2 bs = make_float2(0.f, 0.f);
    for (int q = q_lo; q < q_hi; ++q) {
      const int oh = q / Wo, ow = q - oh * Wo;
      float2 v = __ldg(reinterpret_cast<const float2*>(s + oh * sh + ow * sw));
      v.x *= scale; v.y *= scale;
      bs.x += v.x; bs.y += v.y;
      ssq = fmaf(v.x, v.x, fmaf(v.y, v.y, ssq));
      v.x = round_tf32(v.x); v.y = round_tf32(v.y);
      *reinterpret_cast<float2*>(d + static_cast<long long>(q) * 32) = v;
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

struct YtParams {
  int B, C, H, W;              // source [B][C][H][W] through strides
  long long sn, sc, sh_, sw_;
  int Cs, n_cb;                // staged channels (merged: KW*C) and 32-wide chunks
  int merged, KW, dw, pw;      // merged-kw: channel = kw*C + c, column = ow*sw - pw + kw*dw
  int Hs, Ws, n_rh, n_rw, sth, stw, ah_min, aw_min;
  int rho_h[CG_MAX_KH], rho_w[CG_MAX_KH];
  float scale;
  int slot0;
  long long slot_stride, chunk_stride;   // floats: Hs*Ws*32 and n_slots_total*slot_stride
};

// Yt[plane*n_cb + c/32][slot][hs][ws][c%32] = tf32(scale * S[n][c][h][w]) (zero outside).
// grid (position chunks, B, planes*n_cb).  A lane owns 4 consecutive staged channels of one position and
// issues ONE 16-byte store, so a warp stages 4 positions (4 x 128-byte chunk rows) per step.
//   kVec  : channels-fastest, un-merged source with C % 4 == 0 -> one float4 load per lane
//   !kVec : everything else (merged thin inputs c' = kw*C + c, C % 4 != 0, channel-strided sources) ->
//           4 scalar gathers through per-lane precomputed offsets
// These kernels are instruction-issue bound, not DRAM bound (ncu: 1600 warp instructions per 16 stores with
// 64-bit index arithmetic and a division per position), so all per-position arithmetic is 32-bit and
// incremental: (hs, ws) advance by a block-uniform (dh, dw) per step.  The host checks that one sample's
// source extent fits in 31 bits.
constexpr int kYtSamples = 4;     // samples per block: the same index arithmetic, 4 independent loads in flight

template <bool kVec>
__global__ void __launch_bounds__(512)
stage_yt_kernel(const float* __restrict__ src, const __grid_constant__ YtParams p, float* __restrict__ dst,
                int pos_per_block) {
  const int n0 = blockIdx.y * kYtSamples;
  const int pl = blockIdx.z / p.n_cb, chi = blockIdx.z - pl * p.n_cb;
  const int jh = pl / p.n_rw, jw = pl - jh * p.n_rw;
  const int l8 = threadIdx.x & 7;
  const int sh = static_cast<int>(p.sh_), sw = static_cast<int>(p.sw_), sc = static_cast<int>(p.sc);
  const int H = p.H, W = p.W, Ws = p.Ws, sth = p.sth, stw = p.stw;
  const float scale = p.scale;
  int off[4], wk[4];
  bool ok[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int cs = chi * 32 + 4 * l8 + j;
    int c = cs; wk[j] = 0;
    if (p.merged) { const int kw = cs / p.C; c = cs - kw * p.C; wk[j] = kw * p.dw - p.pw; }
    ok[j] = cs < p.Cs;
    off[j] = ok[j] ? c * sc + wk[j] * sw : 0;
  }
  const float* s[kYtSamples];
  float* d[kYtSamples];
  bool live[kYtSamples];
#pragma unroll
  for (int k = 0; k < kYtSamples; ++k) {
    live[k] = n0 + k < p.B;
    const int n = live[k] ? n0 + k : n0;
    s[k] = src + static_cast<long long>(n) * p.sn;
    d[k] = dst + static_cast<long long>(blockIdx.z) * p.chunk_stride + static_cast<long long>(p.slot0 + n) * p.slot_stride + 4 * l8;
  }
  const int h_base = sth * p.ah_min + p.rho_h[jh];
  const int w_base = p.merged ? 0 : stw * p.aw_min + p.rho_w[jw];      // merged: + wk[j] per channel
  const int n_pos = p.Hs * Ws;
  const int p_lo = blockIdx.x * pos_per_block, p_hi = min(p_lo + pos_per_block, n_pos);
  const int step = blockDim.x >> 3;
  const int dh = step / Ws, dw = step - dh * Ws;
  int pos = p_lo + (threadIdx.x >> 3);
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
      for (int j = 0; j < 4; ++j) { const int w = w0 + wk[j]; in[j] = ok[j] && h_ok && w >= 0 && w < W; }
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
      o.x = round_tf32(v[k].x * scale); o.y = round_tf32(v[k].y * scale);
      o.z = round_tf32(v[k].z * scale); o.w = round_tf32(v[k].w * scale);
      *reinterpret_cast<float4*>(d[k] + pos * 32) = o;
    }
    ws += dw; hs += dh;
    if (ws >= Ws) { ws -= Ws; ++hs; }
  }
}

// out[n][m][p] = Xt[m/32][slot0+n][m%32] * Yt[p/32][slot0+n][p%32]   (Linear layers, Q = 1; p fastest)
__global__ void outer_rows_cl_kernel(const float* __restrict__ Xt, long long x_rows, const float* __restrict__ Yt,
                                     long long y_rows, int M, int P, int slot0, int B, float* __restrict__ out) {
  const long long total = static_cast<long long>(B) * M * P;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int pp = static_cast<int>(i % P);
    const long long t = i / P;
    const int m = static_cast<int>(t % M);
    const int n = static_cast<int>(t / M);
    out[i] = Xt[(static_cast<long long>(m >> 5) * x_rows + slot0 + n) * 32 + (m & 31)] *
             Yt[(static_cast<long long>(pp >> 5) * y_rows + slot0 + n) * 32 + (pp & 31)];
  }
}

}  // namespace cg
