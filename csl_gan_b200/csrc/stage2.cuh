// FP16 capture kernels, second generation (sm_100a).
//
// The staged value of a sample is fp16(v * B-scale * 2^e_n) with e_n chosen from the SAMPLE maximum, so a sample must
// be seen completely before its first element can be written.  cl.cuh's single-pass kernels hold the sample in
// registers across that point (16 float4 per thread -> 2 CTAs per SM, load phase and store phase serialised: ncu
// shows 22 % occupancy, 16-31 % issue, 33-38 % of DRAM peak).  The kernels here read the sample TWICE instead: sweep 1
// is loads + fmax only, the sample maximum is exchanged over the cluster through distributed shared memory, sweep 2
// re-reads the same bytes -- which the CTA itself pulled into L2 microseconds ago -- converts and stores.  No sample
// is held in registers, so 5-8 CTAs per SM are resident at different phases and the SM always has loads in flight.
// DRAM traffic stays 4 B read + 2 B written per element; the second read is L2 -> SM traffic only.
//
//   stage_xt_sweep_kernel  : backprops / Linear-like dense [Q][M] samples  -> Xt[m/64][slot*Q + q][64]
//   stage_yt_sweep_kernel  : dense NHWC activations, space-to-depth planes -> Yt[plane*n_cb + c/64][slot][hs][ws][64]
//   stage_yt_window_kernel : thin inputs with the whole filter window folded into the channel axis (merged = 2): the
//                            rows of the image a CTA needs are copied to shared memory once (zero border included, any
//                            source strides, so an NCHW image needs no layout conversion pass), the im2col rows are
//                            gathered from there (8 shared loads + one 16-byte store per 8 staged channels)
#pragma once
#include "cl.cuh"

namespace cg {

constexpr int kSweepThreads = kFusedThreads;        // 256: cluster_sample_max() is written for this block size
constexpr int kSweepUnroll = 4;

// ---------------------------------------------------------------------------------------------------------------
// Xt[m/64][(slot0+n)*Q + q][m%64] = fp16(scale * src[n][q][m] * 2^e_n); src dense [Q][M] per sample (channels fastest).
// grid = B * parts CTAs in clusters of `parts`; a CTA owns `per` float4 (a multiple of 256) of the sample.
// 256 % (M/4) == 0: a thread keeps ONE channel vector, bias sums in registers -> shared memory -> one global atomic
// per channel and CTA.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSweepThreads, 4)
stage_xt_sweep_kernel(const float* __restrict__ src, long long sn, int M, int Q, float scale, __half* __restrict__ dst,
                      long long rows_total, int slot0, float* __restrict__ bias_rows, float* __restrict__ inv,
                      int parts, int per) {
  __shared__ float s_bias[1024];
  const int n = blockIdx.x / parts, part = blockIdx.x - n * parts;
  const int slot = slot0 + n;
  const int mv = M >> 2;
  const int len4 = Q * mv;
  const int lo = part * per + threadIdx.x;
  const int hi = min(part * per + per, len4);
  const float4* s4 = reinterpret_cast<const float4*>(src + static_cast<long long>(n) * sn);
  if (bias_rows) for (int t = threadIdx.x; t < M; t += kSweepThreads) s_bias[t] = 0.f;
  // sweep 1: the maximum
  float mx = 0.f;
  for (int i0 = lo; i0 < hi; i0 += kSweepUnroll * kSweepThreads) {
    float4 v[kSweepUnroll];
#pragma unroll
    for (int j = 0; j < kSweepUnroll; ++j) {
      const int i = i0 + j * kSweepThreads;
      v[j] = i < hi ? __ldg(s4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < kSweepUnroll; ++j)
      mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v[j].x), fabsf(v[j].y))), fmaxf(fabsf(v[j].z), fabsf(v[j].w)));
  }
  mx = cluster_sample_max(mx, parts);               // (its barriers also order the s_bias zeroing)
  const float hsc = half_scale_for(fabsf(scale) * mx);
  if (part == 0 && threadIdx.x == 0) inv[slot] = 1.0f / hsc;
  // sweep 2: the same elements again (L2), scaled, converted, stored
  const int cv = lo % mv;                           // the same for every iteration: 256 % mv == 0, per % 256 == 0
  const int m = 4 * cv;
  __half* d = dst + (static_cast<long long>(m >> 6) * rows_total + static_cast<long long>(slot) * Q) * 64 + (m & 63);
  const int dq = kSweepThreads / mv;
  int q = lo / mv;
  float4 bs = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i0 = lo; i0 < hi; i0 += kSweepUnroll * kSweepThreads) {
    float4 v[kSweepUnroll];
#pragma unroll
    for (int j = 0; j < kSweepUnroll; ++j) {
      const int i = i0 + j * kSweepThreads;
      v[j] = i < hi ? __ldg(s4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < kSweepUnroll; ++j, q += dq) {
      const int i = i0 + j * kSweepThreads;
      if (i < hi) {
        float4 t = v[j];
        t.x *= scale; t.y *= scale; t.z *= scale; t.w *= scale;
        bs.x += t.x; bs.y += t.y; bs.z += t.z; bs.w += t.w;
        st_elem4(d + static_cast<long long>(q) * 64, t, hsc);
      }
    }
  }
  if (bias_rows) {
    if (lo < hi) {
      atomicAdd(&s_bias[m], bs.x); atomicAdd(&s_bias[m + 1], bs.y);
      atomicAdd(&s_bias[m + 2], bs.z); atomicAdd(&s_bias[m + 3], bs.w);
    }
    __syncthreads();
    float* b = bias_rows + static_cast<long long>(slot) * M;
    for (int t = threadIdx.x; t < M; t += kSweepThreads) {
      if (parts == 1) b[t] = s_bias[t];             // the only contribution: plain store (b was zeroed by the caller)
      else atomicAdd(b + t, s_bias[t]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Space-to-depth FP16 capture (un-merged, dense NHWC sample [H][W][C], C % 4 == 0, 256 % (C/4) == 0, H, W <= 256):
// same two sweeps; the row / column part of the destination offset is tabulated once per block (cl.cuh's
// stage_yt_fused_kernel has the derivation), padding positions are never written and stay zero from allocation.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSweepThreads, 4)
stage_yt_sweep_kernel(const float* __restrict__ src, const __grid_constant__ YtParams p, __half* __restrict__ dst,
                      float* __restrict__ inv, int parts, int per) {
  __shared__ long long row_off[kYtFusedMaxDim];     // (jh*n_rw*n_cb)*chunk_stride + hs*Ws*64, or -1
  __shared__ long long col_off[kYtFusedMaxDim];     // (jw*n_cb)*chunk_stride + ws*64, or -1
  const int n = blockIdx.x / parts, part = blockIdx.x - n * parts;
  const int slot = p.slot0 + n;
  for (int t = threadIdx.x; t < p.H + p.W; t += kSweepThreads) {
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
  const int lo = part * per + threadIdx.x;
  const int hi = min(part * per + per, len4);
  const float4* s4 = reinterpret_cast<const float4*>(src + static_cast<long long>(n) * p.sn);
  float mx = 0.f;
  for (int i0 = lo; i0 < hi; i0 += kSweepUnroll * kSweepThreads) {
    float4 v[kSweepUnroll];
#pragma unroll
    for (int j = 0; j < kSweepUnroll; ++j) {
      const int i = i0 + j * kSweepThreads;
      v[j] = i < hi ? __ldg(s4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < kSweepUnroll; ++j)
      mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v[j].x), fabsf(v[j].y))), fmaxf(fabsf(v[j].z), fabsf(v[j].w)));
  }
  mx = cluster_sample_max(mx, parts);               // (also orders the table writes before the reads below)
  const float scale = p.scale;
  const float hsc = half_scale_for(fabsf(scale) * mx);
  if (part == 0 && threadIdx.x == 0) inv[slot] = 1.0f / hsc;
  const float mult = scale * hsc;                   // hsc is a power of two: one rounding either way
  const int pos0 = lo / cvn, c = 4 * (lo - pos0 * cvn);
  const int dpos = kSweepThreads / cvn;
  int h = pos0 / p.W, w = pos0 - h * p.W;
  __half* dbase = dst + static_cast<long long>(slot) * p.slot_stride + static_cast<long long>(c >> 6) * p.chunk_stride + (c & 63);
  for (int i0 = lo; i0 < hi; i0 += kSweepUnroll * kSweepThreads) {
    float4 v[kSweepUnroll];
#pragma unroll
    for (int j = 0; j < kSweepUnroll; ++j) {
      const int i = i0 + j * kSweepThreads;
      v[j] = i < hi ? __ldg(s4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < kSweepUnroll; ++j) {
      const int i = i0 + j * kSweepThreads;
      if (i < hi) {
        const long long ro = row_off[h], co = col_off[w];
        if (ro >= 0 && co >= 0) st_elem4(dbase + ro + co, v[j], mult);
      }
      w += dpos;
      while (w >= p.W) { w -= p.W; ++h; }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Thin inputs, whole window folded into the channel axis (plan->merged == 2):
//   Yt[c'/64][slot][oh][ow][c'%64] = fp16(scale * S[n][c][oh*sth - ph + kh*dh][ow*stw - pw + kw*dw] * 2^e_n),
//   c' = (kh*KW + kw)*C + c  (zero outside the image; channels c' >= Cs are never written and stay zero).
// grid = B * parts CTAs in clusters of `parts`; a CTA owns `rpp` output rows.  It copies the image rows those windows
// touch into shared memory as img[r][x][c] (x = w + pw, zero border), meets its peers in the sample maximum, and
// gathers: a thread owns ONE octet of staged channels (blockDim = 32 * n_oct, so its 8 tap offsets live in
// registers) and walks the CTA's positions 32 at a time.
// ---------------------------------------------------------------------------------------------------------------
struct YwParams {
  int B, C, H, W;
  long long sn, sc, sh_, sw_;
  int KH, KW, sth, stw, ph, pw, dh, dw, Ho, Wo;
  int Cs, n_oct;                // staged channels, octets that hold at least one of them
  int Wp, rows_max;             // shared image: columns ((Wo-1)*stw + (KW-1)*dw + 1), rows of the largest part
  int rpp;                      // output rows per CTA
  float scale;
  int slot0;
  long long slot_stride, chunk_stride;
};

__device__ __forceinline__ float block_max_any(float v, float* s_warp) {
  v = warp_max(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  if (l == 0) s_warp[w] = v;
  __syncthreads();
  float r = 0.f;
  for (int i = 0; i < nw; ++i) r = fmaxf(r, s_warp[i]);
  return r;
}

__global__ void __launch_bounds__(1024)
stage_yt_window_kernel(const float* __restrict__ src, const __grid_constant__ YwParams p, __half* __restrict__ dst,
                       float* __restrict__ inv, int parts) {
  extern __shared__ float img[];                    // [rows][Wp][C]
  __shared__ float s_warp[32];
  __shared__ float s_cta;
  const int n = blockIdx.x / parts, part = blockIdx.x - n * parts;
  const int slot = p.slot0 + n;
  const int oh0 = part * p.rpp, oh1 = min(oh0 + p.rpp, p.Ho);
  const int n_oh = max(oh1 - oh0, 0);
  const int h_lo = oh0 * p.sth - p.ph;              // image row of shared row 0
  const int rows = n_oh > 0 ? (n_oh - 1) * p.sth + (p.KH - 1) * p.dh + 1 : 0;
  const int C = p.C, Wp = p.Wp;
  const float* s = src + static_cast<long long>(n) * p.sn;
  float mx = 0.f;
  // one warp per image row segment: no per-element divisions
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
  if (p.sc == 1) {
    // channels fastest in memory: a shared row [Wp][C] is contiguous in the image (inside its bounds)
    const int rowlen = Wp * C;
    for (int r = warp; r < rows; r += n_warps) {
      const int h = h_lo + r;
      const bool h_ok = h >= 0 && h < p.H;
      const float* sr = s + static_cast<long long>(h) * p.sh_ - static_cast<long long>(p.pw) * p.sw_;
      float* ir = img + r * rowlen;
      for (int i = lane; i < rowlen; i += 32) {
        const int x = i / C;
        const int w = x - p.pw;
        float v = 0.f;
        if (h_ok && w >= 0 && w < p.W) v = __ldg(sr + static_cast<long long>(x) * p.sw_ + (i - x * C));
        ir[i] = v;
        mx = fmaxf(mx, fabsf(v));
      }
    }
  } else {
    // planar (NCHW-like): a warp takes one (channel, row) line, lanes walk the columns
    for (int cr = warp; cr < C * rows; cr += n_warps) {
      const int c = cr / rows, r = cr - c * rows;
      const int h = h_lo + r;
      const bool h_ok = h >= 0 && h < p.H;
      const float* sr = s + static_cast<long long>(c) * p.sc + static_cast<long long>(h) * p.sh_;
      float* ir = img + r * Wp * C + c;
      for (int x = lane; x < Wp; x += 32) {
        const int w = x - p.pw;
        float v = 0.f;
        if (h_ok && w >= 0 && w < p.W) v = __ldg(sr + static_cast<long long>(w) * p.sw_);
        ir[x * C] = v;
        mx = fmaxf(mx, fabsf(v));
      }
    }
  }
  mx = block_max_any(mx, s_warp);                   // (its barrier also publishes img)
  if (parts > 1) {
    if (threadIdx.x == 0) s_cta = mx;
    cluster_sync_all();
    float g = 0.f;
    for (int r = 0; r < parts; ++r) g = fmaxf(g, ld_shared_cluster_f32(mapa_u32(&s_cta, static_cast<uint32_t>(r))));
    cluster_sync_all();
    mx = g;
  }
  const float hsc = half_scale_for(fabsf(p.scale) * mx);
  if (part == 0 && threadIdx.x == 0) inv[slot] = 1.0f / hsc;
  const float mult = p.scale * hsc;
  // this thread's octet of staged channels: shared-memory offsets of its 8 taps relative to the window origin
  const int n_oct = p.n_oct;
  const int oct = threadIdx.x % n_oct, pl = threadIdx.x / n_oct;      // blockDim = 32 * n_oct -> pl in [0, 32)
  int off[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int cs = 8 * oct + e;
    if (cs < p.Cs) {
      const int t = cs / C, c = cs - t * C;
      const int kh = t / p.KW, kw = t - kh * p.KW;
      off[e] = (kh * p.dh * Wp + kw * p.dw) * C + c;
    } else {
      off[e] = -1;
    }
  }
  const int n_pos = n_oh * p.Wo;
  __half* dbase = dst + static_cast<long long>(oct >> 3) * p.chunk_stride + static_cast<long long>(slot) * p.slot_stride +
                  static_cast<long long>(oh0) * p.Wo * 64 + 8 * (oct & 7);
  for (int pos = pl; pos < n_pos; pos += 32) {
    const int ohl = pos / p.Wo, ow = pos - ohl * p.Wo;
    const float* b = img + (ohl * p.sth * Wp + ow * p.stw) * C;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = off[e] >= 0 ? b[off[e]] * mult : 0.f;
    uint4 o;
    o.x = pack_half2(v[0], v[1]); o.y = pack_half2(v[2], v[3]);
    o.z = pack_half2(v[4], v[5]); o.w = pack_half2(v[6], v[7]);
    *reinterpret_cast<uint4*>(dbase + static_cast<long long>(pos) * 64) = o;
  }
}

}  // namespace cg
