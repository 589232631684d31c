// Per-sample gradients of a THIN first convolution straight from the critic's own tensors (sm_100a, tcgen05).
//
// The 3 -> 64 channel first conv of the CelebA critics has 4800 weights and 1024 window positions: 3 % of the FLOPs,
// but its operands are as large as any layer's, so staging them (capture kernels) and reading them back (contraction)
// cost 23 % of the DP step (ncu, profiles/r2_*).  This kernel never stages anything in HBM:
//
//   G_n[c'][m] = scale * sum_q  patch_n[q][c'] * bp_n[q][m]      c' = (kh*KW + kw)*C + c,  q = window position
//
//   * bp (the layer's grad_output, channels-last fp32 [B][Q][M]) is the MN-major B operand, TMA-loaded RAW from the
//     autograd tensor (tensor map dims {32 ch, rows, M/32 chunks} with the chunk stride SMALLER than the row stride,
//     so the box lands as [chunk][row][32 ch] = cl.cuh's MN-major layout, SWIZZLE_128B with 32-byte atoms).
//     kind::tf32 truncates fp32 words, so the builder warps round the landed tile to TF32 (round-to-nearest) in place
//     -- and sum its columns on the way: the per-sample bias gradient.
//   * the image (any strides: NCHW needs no layout pass) is copied once per sample into shared memory with cp.async
//     (planar, zero border = the conv's padding; double-buffered so sample n+1 loads under sample n's MMAs); the
//     builder warps gather the im2col tile from it as the K-MAJOR A operand (rows c', 32 positions per 128-byte
//     SWIZZLE_128B row; rows >= Cs are never written: they only produce accumulator rows nobody reads).
//   * D = A*B accumulates one sample in TMEM (128 lanes = c', M columns), two accumulator stages; the epilogue
//     warps scale, store G_n in the gradient-natural layout Gs[slot][m][c'] and reduce ||G_n||^2.
//
// TF32 operands (no per-sample scale, hence no sample maximum, hence ONE pass over the data): 10-bit mantissa like
// the FP16 containers of the other layers; the MMA rate does not matter for this layer.
// DRAM per sample: bp 4*Q*M + image 4*C*H*W + Gs 4*M*Cs  (330 KB for the CelebA layer; the staged route moved 1 MB).
#pragma once
#include "cl.cuh"

namespace cg {

constexpr int kThinKb = 64;                          // window positions per k-block
constexpr int kThinStages = 3;
constexpr int kThinBuilderWarps = 16;
constexpr int kThinBuilders = 32 * kThinBuilderWarps;
constexpr int kThinFirstBuilder = 32 * 6;            // warps 0 (TMA), 1 (MMA), 2-5 (epilogue), 6-21 (builders)
constexpr int kThinThreads = kThinFirstBuilder + kThinBuilders;

struct ThinParams {
  const float* act;                                  // image batch through strides (elements)
  long long a_sn, a_sc, a_sh, a_sw;
  int C, H, W;
  int KH, KW, sth, stw, ph, pw, dh, dw, Ho, Wo;
  int Cs, n_g;                                       // staged rows KH*KW*C (<= 128) and 8-row groups built
  int Hp, Wp, P;                                     // shared image: rows, columns, floats per plane
  int Hc, Wc;                                        // image rows / columns that fall inside the shared image
  int M, n_ch;                                       // backprop channels (MMA N) and their 32-channel chunks
  int Q, nkb;                                        // positions per sample, k-blocks per sample
  int B;
  float scale;
  float* Gs; long long gs_stride;                    // [B][M][Cs] (+ stride between samples)
  float* norm2;                                      // [B], zeroed by the launcher
  float* bias_rows;                                  // [B][M] or null
  int a_bytes, b_bytes, stage_bytes, img_floats;     // A tile (2 atoms of n_g KB), B tile, their sum, one image buffer
  int tmem_cols;
  // second segment (both passes of a step in ONE launch: 2B items on 148 CTAs waste less of the last round than B
  // twice): items [B0, B) read act2 / the second tensor map and write Gs2 / norm2_2 / bias_rows2; B0 == B: unused
  int B0;
  const float* act2;
  float* Gs2;
  float* norm2_2;
  float* bias_rows2;
};

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
// Round-to-nearest TF32 for an operand the tensor core will TRUNCATE to its upper 19 bits (kind::tf32 ignores the low
// 13 mantissa bits, measured): adding half an ulp to the bit pattern is all that is left to do (carries run into the
// exponent as they should; Inf becomes NaN, which a gradient containing Inf is anyway).  One instruction instead of
// the three cvt.rna.tf32.f32 expands to.
__device__ __forceinline__ uint32_t tf32_rn_bits(float x) { return __float_as_uint(x) + 0x1000u; }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void builder_barrier() {
  asm volatile("bar.sync 1, %0;" ::"n"(kThinBuilders) : "memory");
}

__global__ void __launch_bounds__(kThinThreads, 1)
thin_direct_kernel(const __grid_constant__ CUtensorMap tmap_bp, const __grid_constant__ CUtensorMap tmap_bp2,
                   const __grid_constant__ ThinParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* img = reinterpret_cast<float*>(tiles + kThinStages * p.stage_bytes);        // 2 buffers of img_floats
  float* s_bias = reinterpret_cast<float*>(img + 2 * p.img_floats);                   // [128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + 128);
  uint64_t* full_bar = bars;                         // TMA landed the raw bp tile
  uint64_t* ready_bar = bars + kThinStages;          // builders finished the stage (A built, B rounded)
  uint64_t* empty_bar = bars + 2 * kThinStages;      // MMAs of the stage retired
  uint64_t* acc_full = bars + 3 * kThinStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_ew = (p.Cs + 31) >> 5;                 // epilogue warps that own live accumulator rows
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_bp);
    tma_prefetch_desc(&tmap_bp2);
    for (int s = 0; s < kThinStages; ++s) {
      mbar_init(&full_bar[s], 1); mbar_init(&ready_bar[s], kThinBuilderWarps); mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], n_ew); }
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, p.tmem_cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: raw bp tiles =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int n = blockIdx.x; n < p.B; n += gridDim.x) {
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], static_cast<uint32_t>(p.b_bytes));
          tma_load_3d(tiles + stage * p.stage_bytes, n < p.B0 ? &tmap_bp : &tmap_bp2, &full_bar[stage], 0,
                      (n < p.B0 ? n : n - p.B0) * p.Q + kb * kThinKb, 0);
          if (++stage == kThinStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // A: K-major (im2col rows), B: MN-major (bp), TF32 -> FP32, M = 128 accumulator rows, N = p.M
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (1u << 16) |
                             ((static_cast<uint32_t>(p.M) >> 3) << 17) | ((128u >> 4) << 24);
      int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
      for (int n = blockIdx.x; n < p.B; n += gridDim.x) {
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * p.M);
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(&ready_bar[stage], phase);
          tc_fence_after();
          const uint32_t bs = smem_u32(tiles + stage * p.stage_bytes);
          const uint32_t as = bs + static_cast<uint32_t>(p.b_bytes);
          const uint64_t bdesc = umma_desc_mn_sw128(bs, static_cast<uint32_t>(kThinKb * 128));
#pragma unroll
          for (int j = 0; j < kThinKb / 8; ++j) {
            // K step j: positions [8j, 8j+8) = atom j/4, 32-byte column (j%4) of the K-major A rows; rows 8j.. of B
            const uint64_t adesc = umma_desc_k_sw128(as + static_cast<uint32_t>((j >> 2) * p.n_g * 1024)) +
                                   static_cast<uint64_t>(2 * (j & 3));
            umma_tf32(tmem_d, adesc, bdesc + static_cast<uint64_t>(64 * j), idesc, (kb > 0 || j > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == kThinStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&acc_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp < 6) {
    // ===================== epilogue: TMEM -> Gs[n][m][c'], ||G_n||^2 =====================
    const int ew = warp & 3;                         // TMEM lane quadrant this warp may read
    if (ew < n_ew) {
      const int row = ew * 32 + lane;                // c'
      const bool live = row < p.Cs;
      int acc = 0; uint32_t acc_phase = 0;
      for (int n = blockIdx.x; n < p.B; n += gridDim.x) {
        mbar_wait(&acc_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(acc * p.M);
        float* o = (n < p.B0 ? p.Gs + static_cast<long long>(n) * p.gs_stride
                             : p.Gs2 + static_cast<long long>(n - p.B0) * p.gs_stride) + row;
        float ss = 0.f;
        for (int c0 = 0; c0 < p.M; c0 += 16) {
          float v[16];
          tmem_ld16(taddr + c0, v);
          if (live) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float g = v[i] * p.scale;
              ss = fmaf(g, g, ss);
              o[static_cast<long long>(c0 + i) * p.Cs] = g;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[acc]);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, s);
        if (lane == 0) atomicAdd(n < p.B0 ? p.norm2 + n : p.norm2_2 + (n - p.B0), ss);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== builders =====================
    // (explicit 32-bit shared-space addresses throughout: through generic pointers the compiler emitted LD.E / ST.E
    //  with 64-bit address arithmetic and a loop-carried dependency per element -- ncu source page, 4000 cycles per
    //  k-block against ~600 of work)
    const int bt = threadIdx.x - kThinFirstBuilder;
    const int bw = bt >> 5;
    for (int i = bt; i < 2 * p.img_floats; i += kThinBuilders) img[i] = 0.f;      // borders stay zero for good
    if (bt < 128) s_bias[bt] = 0.f;
    builder_barrier();
    const uint32_t img_s = smem_u32(img);
    const bool wide = p.a_sw == 1 && (p.Wp & 1) == 0 && (p.P & 1) == 0 && (p.pw & 1) == 0 && (p.Wc & 1) == 0 &&
                      (p.a_sn & 1) == 0 && (p.a_sc & 1) == 0 && (p.a_sh & 1) == 0 &&
                      (reinterpret_cast<uintptr_t>(p.act) & 7) == 0 && (reinterpret_cast<uintptr_t>(p.act2) & 7) == 0;
    auto issue_image = [&](int n, uint32_t buf_s) {
      const float* s = n < p.B0 ? p.act + static_cast<long long>(n) * p.a_sn
                                : p.act2 + static_cast<long long>(n - p.B0) * p.a_sn;
      for (int line = bw; line < p.C * p.Hc; line += kThinBuilderWarps) {
        const int c = line / p.Hc, h = line - c * p.Hc;
        const float* sr = s + c * p.a_sc + h * p.a_sh;
        const uint32_t dr = buf_s + 4u * static_cast<uint32_t>(c * p.P + (h + p.ph) * p.Wp + p.pw);
        if (wide) {
          for (int w = 2 * lane; w < p.Wc; w += 64)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dr + 4u * w), "l"(sr + w) : "memory");
        } else {
          for (int w = lane; w < p.Wc; w += 32)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dr + 4u * w), "l"(sr + w * p.a_sw) : "memory");
        }
      }
      cp_async_commit();
    };
    // B-tile rounding: a thread owns one 16-byte unit of rows with a fixed (row % 4), hence fixed logical channels
    const int unit = bt & 7;
    const int chunk = (bt >> 3) % p.n_ch;
    const int rstep = kThinBuilders / (8 * p.n_ch);                    // a multiple of 4
    const int rbase = bt / (8 * p.n_ch);
    const int bch = chunk * 32 + (((unit >> 1) ^ (rbase & 3)) << 3) + ((unit & 1) << 2);   // first of its 4 channels
    const uint32_t b_lane = static_cast<uint32_t>(chunk * (kThinKb * 128) + rbase * 128 + unit * 16);
    // A-tile building: warp = (16-position segment, group residue mod 4); lane = (row in the 8-row group, position quad)
    const int seg = bw & 3, gq = bw >> 2;
    const int r8 = lane & 7, quad = lane >> 3;
    const uint32_t a_lane = static_cast<uint32_t>((seg >> 1) * p.n_g * 1024 + r8 * 128 + (((((seg & 1) << 2) | quad) ^ r8) << 4));
    uint32_t toff[4];                                  // byte offsets of this lane's staged rows inside the image
    bool tok[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int g = gq + 4 * i, cs = 8 * g + r8;
      tok[i] = g < p.n_g && cs < p.Cs;
      toff[i] = 0;
      if (tok[i]) {
        const int t = cs / p.C, c = cs - t * p.C;
        const int kh = t / p.KW, kw = t - kh * p.KW;
        toff[i] = 4u * static_cast<uint32_t>(c * p.P + kh * p.dh * p.Wp + kw * p.dw);
      }
    }
    const uint32_t s1 = 4u * static_cast<uint32_t>(p.stw);

    int stage = 0; uint32_t phase = 0;
    int k = 0;
    if (static_cast<int>(blockIdx.x) < p.B) issue_image(blockIdx.x, img_s);
    for (int n = blockIdx.x; n < p.B; n += gridDim.x, ++k) {
      cp_async_wait_all();
      builder_barrier();                             // image n complete everywhere; nobody still reads image n-1
      const uint32_t im_s = img_s + 4u * static_cast<uint32_t>((k & 1) * p.img_floats);
      if (n + static_cast<int>(gridDim.x) < p.B)
        issue_image(n + gridDim.x, img_s + 4u * static_cast<uint32_t>(((k + 1) & 1) * p.img_floats));
      float b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
      int pos = 16 * seg + 4 * quad;
      int oh = pos / p.Wo, ow = pos - oh * p.Wo;
      for (int kb = 0; kb < p.nkb; ++kb) {
        // gather first (the image does not depend on the stage), then wait for the stage and store
        const uint32_t src0 = im_s + 4u * static_cast<uint32_t>(oh * p.sth * p.Wp + ow * p.stw);
        float v[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (tok[i]) {
            const uint32_t a = src0 + toff[i];
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[i][0]) : "r"(a));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[i][1]) : "r"(a + s1));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[i][2]) : "r"(a + 2 * s1));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[i][3]) : "r"(a + 3 * s1));
          }
        }
        mbar_wait(&empty_bar[stage], phase ^ 1);
        const uint32_t bs_s = smem_u32(tiles + stage * p.stage_bytes);
        const uint32_t a_base = bs_s + static_cast<uint32_t>(p.b_bytes) + a_lane;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (tok[i]) {
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_base + static_cast<uint32_t>((gq + 4 * i) * 1024)),
                         "r"(tf32_rn_bits((v[i][0]))), "r"(tf32_rn_bits((v[i][1]))),
                         "r"(tf32_rn_bits((v[i][2]))), "r"(tf32_rn_bits((v[i][3]))) : "memory");
          }
        }
        mbar_wait(&full_bar[stage], phase);
        for (int r = rbase; r < kThinKb; r += 2 * rstep) {
          const uint32_t q0 = bs_s + b_lane + static_cast<uint32_t>((r - rbase) * 128);
          const uint32_t q1 = q0 + static_cast<uint32_t>(rstep * 128);
          const bool two = r + rstep < kThinKb;
          float4 x, y = make_float4(0.f, 0.f, 0.f, 0.f);
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(q0));
          if (two) asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(y.x), "=f"(y.y), "=f"(y.z), "=f"(y.w) : "r"(q1));
          b0 += x.x + y.x; b1 += x.y + y.y; b2 += x.z + y.z; b3 += x.w + y.w;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(q0), "r"(tf32_rn_bits((x.x))),
                       "r"(tf32_rn_bits((x.y))), "r"(tf32_rn_bits((x.z))),
                       "r"(tf32_rn_bits((x.w))) : "memory");
          if (two) asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(q1), "r"(tf32_rn_bits((y.x))),
                                "r"(tf32_rn_bits((y.y))), "r"(tf32_rn_bits((y.z))),
                                "r"(tf32_rn_bits((y.w))) : "memory");
        }
        fence_proxy_async();                         // generic-proxy writes -> visible to the tensor core's reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready_bar[stage]);
        if (++stage == kThinStages) { stage = 0; phase ^= 1; }
        ow += kThinKb;
        while (ow >= p.Wo) { ow -= p.Wo; ++oh; }
      }
      if (p.bias_rows) {
        atomicAdd(&s_bias[bch], b0); atomicAdd(&s_bias[bch + 1], b1);
        atomicAdd(&s_bias[bch + 2], b2); atomicAdd(&s_bias[bch + 3], b3);
        builder_barrier();
        if (bt < p.M) {
          float* br = n < p.B0 ? p.bias_rows + static_cast<long long>(n) * p.M
                               : p.bias_rows2 + static_cast<long long>(n - p.B0) * p.M;
          br[bt] = p.scale * s_bias[bt];
          s_bias[bt] = 0.f;
        }
      }
    }
    cp_async_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, p.tmem_cols); }
}

}  // namespace cg
