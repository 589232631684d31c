// Bandwidth kernels around the contraction: staging (capture), norms, clip factors, permutes,
// Philox noise, per-sample row norms and L2 clipping.  All coalesced / vectorised where the
// layout allows, warp-shuffle reductions, grid sized in multiples of the SM count.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <curand_kernel.h>

#include "../../include/cslgan_b200.h"
#include "ptx.cuh"

namespace cg {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum (blockDim.x multiple of 32, <= 1024). Result valid in thread 0.
__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.f;
  if (w == 0) v = warp_sum(v);
  __syncthreads();
  return v;
}

// ------------------------------------------------------------------------------------------
// stage_rows_t: src[B][R] -> dst[r][slot0+n] (32x32 shared-memory transpose tiles)
// grid (ceil(R/32), ceil(B/32)), block (32, 8)
// ------------------------------------------------------------------------------------------
__global__ void stage_rows_t_kernel(const float* __restrict__ src, int B, int R, float scale,
                                    float* __restrict__ dst, long long dst_pitch, int slot0,
                                    float* __restrict__ copy_out, float* __restrict__ sumsq) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int n = n0 + ty + i, r = r0 + tx;
    float v = 0.f;
    if (n < B && r < R) {
      v = scale * src[static_cast<long long>(n) * R + r];
      if (copy_out) copy_out[static_cast<long long>(slot0 + n) * R + r] = v;
    }
    tile[ty + i][tx] = v;
    if (sumsq) {
      float s = warp_sum(v * v);
      if (tx == 0 && n < B) atomicAdd(sumsq + slot0 + n, s);
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int r = r0 + ty + i, n = n0 + tx;
    if (r < R && n < B) dst[static_cast<long long>(r) * dst_pitch + slot0 + n] = round_tf32(tile[tx][ty + i]);
  }
}

// ------------------------------------------------------------------------------------------
// stage_rows: src[B][R][Q] -> dst[r][(slot0+n)*Qpad + q], zero padded; one warp per (n, r) row
// ------------------------------------------------------------------------------------------
__global__ void stage_rows_kernel(const float* __restrict__ src, int B, int R, int Q, int Wo, int Wop, int Qpad,
                                  float scale, float* __restrict__ dst, long long dst_pitch, int slot0,
                                  float* __restrict__ rowsum) {
  const long long nrows = static_cast<long long>(B) * R;
  const int lane = threadIdx.x & 31;
  const long long warps_total = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long w = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; w < nrows; w += warps_total) {
    const int n = static_cast<int>(w / R), r = static_cast<int>(w % R);
    const float* s = src + w * Q;
    float* d = dst + static_cast<long long>(r) * dst_pitch + static_cast<long long>(slot0 + n) * Qpad;
    float acc = 0.f;
    for (int t = lane; t < Qpad; t += 32) {
      float v = 0.f;
      int q = t;
      bool ok = t < Q;
      if (Wop != Wo) {                      // padded window rows: t = oh*Wop + ow
        const int oh = t / Wop, ow = t - oh * Wop;
        q = oh * Wo + ow;
        ok = ow < Wo && q < Q;
      }
      if (ok) {
        v = scale * s[q];
        acc += v;
      }
      d[t] = round_tf32(v);
    }
    if (rowsum) {
      acc = warp_sum(acc);
      if (lane == 0) rowsum[static_cast<long long>(slot0 + n) * R + r] = acc;
    }
  }
}

// ------------------------------------------------------------------------------------------
// stage_unfold: src[B][C][H][W] -> kw-plane matrix
//   dst[((j*KW + kw)*C + c)][ (slot0+n)*Hs*Wop + hs*Wop + ow ] = src[n][c][sh*(hs+a_min)+rho_j][ow*sw - pw + kw*dw]
// One thread per destination element; ow fastest so writes are coalesced and the strided reads
// of neighbouring kw planes hit the same lines in L1/L2.
// ------------------------------------------------------------------------------------------
struct UnfoldParams {
  int B, C, H, W, KW, sh, sw, pw, dw, Wo, Wop, Hs, n_rho, a_min;
  int rho[CG_MAX_KH];
  float scale;
  long long dst_pitch;
  int slot0;
};

__global__ void stage_unfold_kernel(const float* __restrict__ src, const __grid_constant__ UnfoldParams p,
                                    float* __restrict__ dst) {
  const long long per_row = static_cast<long long>(p.B) * p.Hs * p.Wop;
  const long long total = static_cast<long long>(p.n_rho) * p.KW * p.C * per_row;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int ow = static_cast<int>(i % p.Wop);
    long long t = i / p.Wop;
    const int hs = static_cast<int>(t % p.Hs);
    t /= p.Hs;
    const int n = static_cast<int>(t % p.B);
    t /= p.B;
    const int c = static_cast<int>(t % p.C);
    t /= p.C;
    const int kw = static_cast<int>(t % p.KW);
    const int j = static_cast<int>(t / p.KW);
    const int h = p.sh * (hs + p.a_min) + p.rho[j];
    const int w = ow * p.sw - p.pw + kw * p.dw;
    float v = 0.f;
    if (ow < p.Wo && h >= 0 && h < p.H && w >= 0 && w < p.W)
      v = p.scale * src[((static_cast<long long>(n) * p.C + c) * p.H + h) * p.W + w];
    const long long row = (static_cast<long long>(j) * p.KW + kw) * p.C + c;
    dst[row * p.dst_pitch + (static_cast<long long>(p.slot0 + n) * p.Hs + hs) * p.Wop + ow] = round_tf32(v);
  }
}

// Fast path: one block stages `cpb` input planes of one sample through shared memory (input read
// once, coalesced) and writes all n_rho*KW kw-planes of those channels (coalesced rows of Hs*Wop).
// grid (ceil(C/cpb), B), dynamic smem = cpb*H*W floats.
__global__ void stage_unfold_smem_kernel(const float* __restrict__ src, const __grid_constant__ UnfoldParams p,
                                         float* __restrict__ dst, int cpb) {
  extern __shared__ float sm[];
  const int n = blockIdx.y;
  const int c0 = blockIdx.x * cpb;
  const int nc = min(cpb, p.C - c0);
  const int hw = p.H * p.W;
  const float* base = src + (static_cast<long long>(n) * p.C + c0) * hw;
  const int tot_in = nc * hw;
  if ((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (tot_in & 3) == 0) {
    const float4* b4 = reinterpret_cast<const float4*>(base);
    float4* s4 = reinterpret_cast<float4*>(sm);
    for (int i = threadIdx.x; i < (tot_in >> 2); i += blockDim.x) s4[i] = __ldg(b4 + i);
  } else {
    for (int i = threadIdx.x; i < tot_in; i += blockDim.x) sm[i] = base[i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int per = p.Hs * p.Wop;
  const int n_out_rows = p.n_rho * p.KW * nc;
  const long long col0 = static_cast<long long>(p.slot0 + n) * per;
  for (int orow = warp; orow < n_out_rows; orow += nwarps) {
    const int cc = orow % nc;
    const int t = orow / nc;
    const int kw = t % p.KW, j = t / p.KW;
    const float* plane = sm + cc * hw;
    float* drow = dst + (static_cast<long long>(j * p.KW + kw) * p.C + c0 + cc) * p.dst_pitch + col0;
    const int hoff = p.sh * p.a_min + p.rho[j];
    const int woff = kw * p.dw - p.pw;
    for (int i = lane; i < per; i += 32) {
      const int hs = i / p.Wop, ow = i - hs * p.Wop;
      const int h = p.sh * hs + hoff, w = ow * p.sw + woff;
      float v = 0.f;
      if (ow < p.Wo && h >= 0 && h < p.H && w >= 0 && w < p.W) v = p.scale * plane[h * p.W + w];
      drow[i] = round_tf32(v);
    }
  }
}

// ------------------------------------------------------------------------------------------
// row reductions
// ------------------------------------------------------------------------------------------
// one block per row (grid-stride over rows); float4 path when the row is 16-byte aligned
__device__ __forceinline__ float row_sumsq_device(const float* __restrict__ row, long long cols) {
  float acc = 0.f;
  if ((reinterpret_cast<uintptr_t>(row) & 15) == 0) {
    const long long n4 = cols >> 2;
    const float4* r4 = reinterpret_cast<const float4*>(row);
    for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
      const float4 v = __ldg(r4 + i);
      acc = fmaf(v.x, v.x, acc);
      acc = fmaf(v.y, v.y, acc);
      acc = fmaf(v.z, v.z, acc);
      acc = fmaf(v.w, v.w, acc);
    }
    for (long long i = (n4 << 2) + threadIdx.x; i < cols; i += blockDim.x) acc = fmaf(row[i], row[i], acc);
  } else {
    for (long long i = threadIdx.x; i < cols; i += blockDim.x) acc = fmaf(row[i], row[i], acc);
  }
  return acc;
}

__global__ void row_sumsq_kernel(const float* __restrict__ src, long long rows, long long cols, long long ld,
                                 float* __restrict__ out, int accumulate, int take_sqrt) {
  __shared__ float sh[32];
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    float s = block_sum(row_sumsq_device(src + r * ld, cols), sh);
    if (threadIdx.x == 0) {
      if (take_sqrt) s = sqrtf(s);
      out[r] = accumulate ? out[r] + s : s;
    }
  }
}

// few long rows (the flat gradient of the immediate-sensitivity norm: rows = 1, cols = |theta|): every row is
// split over gridDim.x blocks, partial sums meet in out[r] (zeroed by the caller) through one atomic per block;
// sqrt_inplace_kernel finishes the norm.  grid (nsplit, rows); `per` is a multiple of 4.
__global__ void row_sumsq_split_kernel(const float* __restrict__ src, long long cols, long long ld, long long per,
                                       float* __restrict__ out) {
  __shared__ float sh[32];
  const long long r = blockIdx.y;
  const long long lo = static_cast<long long>(blockIdx.x) * per;
  if (lo >= cols) return;
  const long long len = (lo + per <= cols) ? per : cols - lo;
  const float s = block_sum(row_sumsq_device(src + r * ld + lo, len), sh);
  if (threadIdx.x == 0) atomicAdd(out + r, s);
}

__global__ void sqrt_inplace_kernel(float* __restrict__ v, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) v[i] = sqrtf(v[i]);
}

// small rows: one warp per row
__global__ void row_sumsq_warp_kernel(const float* __restrict__ src, long long rows, long long cols, long long ld,
                                      float* __restrict__ out, int accumulate, int take_sqrt) {
  const int lane = threadIdx.x & 31;
  const long long warps_total = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long r = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps_total) {
    const float* row = src + r * ld;
    float acc = 0.f;
    for (long long i = lane; i < cols; i += 32) acc = fmaf(row[i], row[i], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      if (take_sqrt) acc = sqrtf(acc);
      out[r] = accumulate ? out[r] + acc : acc;
    }
  }
}

// out[n][m][p] = X[m][slot0+n] * Y[p][slot0+n]; p fastest (coalesced writes)
__global__ void outer_rows_kernel(const float* __restrict__ X, long long x_pitch, const float* __restrict__ Y,
                                  long long y_pitch, int M, int P, int slot0, int B, float* __restrict__ out) {
  const long long total = static_cast<long long>(B) * M * P;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int pp = static_cast<int>(i % P);
    const long long t = i / P;
    const int m = static_cast<int>(t % M);
    const int n = static_cast<int>(t / M);
    out[i] = X[static_cast<long long>(m) * x_pitch + slot0 + n] * Y[static_cast<long long>(pp) * y_pitch + slot0 + n];
  }
}

// out[n] (+)= <T[:, row_a+n, :], T[:, row_b+n, :]> over a chunk-major matrix T[n_chunks][rows_total][32]
// (cross terms of the joint Linear norm: ||b1 a1^T + b2 a2^T||^2 = ... + 2 (a1.a2)(b1.b2)); one warp per n
__global__ void rowpair_dot_kernel(const float* __restrict__ T, long long rows_total, int n_chunks, int row_a,
                                   int row_b, int B, float* __restrict__ out, int accumulate) {
  const int lane = threadIdx.x & 31;
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (n >= B) return;
  float acc = 0.f;
  for (int c = 0; c < n_chunks; ++c) {
    const float* base = T + static_cast<long long>(c) * rows_total * 32;
    acc = fmaf(base[static_cast<long long>(row_a + n) * 32 + lane], base[static_cast<long long>(row_b + n) * 32 + lane], acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) out[n] = accumulate ? out[n] + acc : acc;
}

// out[n] = || sum_s rows_in[(slot_lo + n + s*seg_stride)][:] ||^2   (joint per-sample bias-gradient norms)
__global__ void joint_rows_sumsq_kernel(const float* __restrict__ rows_in, int R, int slot_lo, int seg_stride,
                                        int n_seg, int B, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (n >= B) return;
  float acc = 0.f;
  for (int r = lane; r < R; r += 32) {
    float v = 0.f;
    for (int s = 0; s < n_seg; ++s) v += rows_in[static_cast<long long>(slot_lo + n + s * seg_stride) * R + r];
    acc = fmaf(v, v, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) out[n] = acc;
}

__global__ void vec_mul_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                               long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = a[i] * b[i];
}

__global__ void vec_fma_kernel(const float* __restrict__ a, const float* __restrict__ b, float w, float* __restrict__ out,
                               long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = fmaf(w * a[i], b[i], out[i]);
}

__global__ void clip_factors_kernel(const float* __restrict__ norm2, int n_params, int n_slots, int per_layer,
                                    const float* __restrict__ C, float c_scale, int clip_lo, int clip_hi,
                                    float* __restrict__ factors, float* __restrict__ norms_out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  const bool clipped = (s >= clip_lo && s < clip_hi);
  if (per_layer) {
    for (int k = 0; k < n_params; ++k) {
      const float n = sqrtf(norm2[static_cast<long long>(k) * n_slots + s]);
      if (norms_out) norms_out[static_cast<long long>(k) * n_slots + s] = n;
      const float f = fminf((c_scale == 1.0f ? C[k] : C[k] * c_scale) / (n + 1e-6f), 1.0f);
      factors[static_cast<long long>(k) * n_slots + s] = clipped ? f : 1.0f;
    }
  } else {
    float tot = 0.f;
    for (int k = 0; k < n_params; ++k) tot += norm2[static_cast<long long>(k) * n_slots + s];
    const float n = sqrtf(tot);
    if (norms_out) norms_out[s] = n;
    const float f = fminf((c_scale == 1.0f ? C[0] : C[0] * c_scale) / (n + 1e-6f), 1.0f);
    factors[s] = clipped ? f : 1.0f;
  }
}

// dst[r][slot*stride + q] = tf32(src * factor[slot]); grid (col chunks, rows); float4 when 4-aligned
__global__ void scale_slots_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows,
                                   long long pitch, int slot_stride, int slot_lo, int slot_hi,
                                   const float* __restrict__ factor, int vec4) {
  const long long cols = static_cast<long long>(slot_hi - slot_lo) * slot_stride;
  const long long col0 = static_cast<long long>(slot_lo) * slot_stride;
  for (int r = blockIdx.y; r < rows; r += gridDim.y) {
    const float* s = src + static_cast<long long>(r) * pitch + col0;
    float* d = dst + static_cast<long long>(r) * pitch + col0;
    if (vec4) {
      const long long n4 = cols >> 2;
      for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
           i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float f = __ldg(factor + slot_lo + static_cast<int>((i << 2) / slot_stride));
        float4 v = __ldg(reinterpret_cast<const float4*>(s) + i);
        v.x = round_tf32(v.x * f); v.y = round_tf32(v.y * f); v.z = round_tf32(v.z * f); v.w = round_tf32(v.w * f);
        reinterpret_cast<float4*>(d)[i] = v;
      }
    } else {
      for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < cols;
           i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float f = __ldg(factor + slot_lo + static_cast<int>(i / slot_stride));
        d[i] = round_tf32(s[i] * f);
      }
    }
  }
}

// out[m][c][kh][kw] (+)= T[m][kh][kw*C + c]
__global__ void permute_accum_kernel(const float* __restrict__ T, float* __restrict__ out, int M, int C, int KH,
                                     int KW, int accumulate) {
  const long long total = static_cast<long long>(M) * C * KH * KW;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    // i enumerates the destination (coalesced writes); reads are strided but L2 resident
    const int kw = static_cast<int>(i % KW);
    long long t = i / KW;
    const int kh = static_cast<int>(t % KH);
    t /= KH;
    const int c = static_cast<int>(t % C);
    const int m = static_cast<int>(t / C);
    const float v = T[(static_cast<long long>(m) * KH + kh) * (static_cast<long long>(KW) * C) + static_cast<long long>(kw) * C + c];
    out[i] = accumulate ? out[i] + v : v;
  }
}

// out[r] (+)= sum_slot factor[slot] * rows_in[slot*R + r].  block (32 columns, 8 slot lanes); slots are also
// split over blockIdx.y so the (tiny) reduction is spread over many SMs; one atomic per column per block.
__global__ void weighted_colsum_kernel(const float* __restrict__ rows_in, const float* __restrict__ factor,
                                       int slot_lo, int slot_hi, int R, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int r = blockIdx.x * 32 + threadIdx.x;
  const int per = (slot_hi - slot_lo + gridDim.y - 1) / gridDim.y;
  const int lo = slot_lo + blockIdx.y * per;
  const int hi = min(lo + per, slot_hi);
  float acc = 0.f;
  if (r < R)
    for (int s = lo + threadIdx.y; s < hi; s += 8) acc = fmaf(__ldg(factor + s), __ldg(rows_in + static_cast<long long>(s) * R + r), acc);
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && r < R) {
#pragma unroll
    for (int k = 1; k < 8; ++k) acc += red[k][threadIdx.x];
    if (hi > lo) atomicAdd(out + r, acc);
  }
}

// ------------------------------------------------------------------------------------------
// Batched small operations: the per-layer scalar / bias work of a step (bias-gradient norms, Linear closed-form
// norms, clip multipliers, bias and thin-layer clipped sums) is ~25 launches of 3-5 us each when issued one by one --
// 8 % of the CelebA step.  One launch runs a table of them; a block finds its operation by the table's block prefix.
// ------------------------------------------------------------------------------------------
constexpr int kSmallMaxOps = 32;
struct SmallOp {
  int op;                      // CG_OP_*
  int R;                       // columns (ROW_SUMSQ, WCOLSUM)
  int lo;                      // first slot (WCOLSUM, CLIP_MULT)
  int blk0;                    // first block of this operation
  int nblk;                    // blocks of this operation
  int nx;                      // WCOLSUM: column tiles of 32
  long long n;                 // rows / elements / slots
  const float* a;
  const float* b;
  const float* c;
  float* out;
  float* out2;
};
struct SmallParams {
  SmallOp op[kSmallMaxOps];
  int n_ops;
};

__global__ void __launch_bounds__(256)
small_ops_kernel(const __grid_constant__ SmallParams p) {
  __shared__ float red[8][33];
  __shared__ float s_down;
  int i = 0;
#pragma unroll 1
  while (i + 1 < p.n_ops && static_cast<int>(blockIdx.x) >= p.op[i + 1].blk0) ++i;
  const SmallOp& o = p.op[i];
  const int bl = blockIdx.x - o.blk0;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  if (o.op == CG_OP_ROW_SUMSQ) {
    // out[r] = sum_j a[r*R + j]^2 : one warp per row
    for (long long r = static_cast<long long>(bl) * 8 + w; r < o.n; r += static_cast<long long>(o.nblk) * 8) {
      const float* row = o.a + r * o.R;
      float acc = 0.f;
      for (int j = lane; j < o.R; j += 32) acc = fmaf(row[j], row[j], acc);
      acc = warp_sum(acc);
      if (lane == 0) o.out[r] = acc;
    }
  } else if (o.op == CG_OP_COPY || o.op == CG_OP_MUL) {
    for (long long j = static_cast<long long>(bl) * 256 + t; j < o.n; j += static_cast<long long>(o.nblk) * 256)
      o.out[j] = o.op == CG_OP_COPY ? o.a[j] : o.a[j] * o.b[j];
  } else if (o.op == CG_OP_WCOLSUM) {
    // out[r] += sum_s b[lo + s] * a[(lo + s)*R + r]  (out zeroed by the caller); 32 columns x 8 slot lanes per block
    const int tx = t & 31, ty = t >> 5;
    const int ct = bl % o.nx, sy = bl / o.nx, ny = o.nblk / o.nx;
    const int r = ct * 32 + tx;
    const int per = static_cast<int>((o.n + ny - 1) / ny);
    const int s0 = o.lo + sy * per;
    const int s1 = min(s0 + per, o.lo + static_cast<int>(o.n));
    float acc = 0.f;
    if (r < o.R)
      for (int s = s0 + ty; s < s1; s += 8) acc = fmaf(__ldg(o.b + s), __ldg(o.a + static_cast<long long>(s) * o.R + r), acc);
    red[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && r < o.R) {
#pragma unroll
      for (int k = 1; k < 8; ++k) acc += red[k][tx];
      if (s1 > s0) atomicAdd(o.out + r, acc);
    }
  } else if (o.op == CG_OP_CLIP_MULT) {
    // out[s] = a[s]*b[s]*c[s] / 2^E, out2[0] = 2^E (cl.cuh clip_mult_kernel for one layer, one block)
    const int hi = o.lo + static_cast<int>(o.n);
    float m = 0.f;
    for (int s = o.lo + t; s < hi; s += 256) m = fmaxf(m, o.a[s] * o.b[s] * o.c[s]);
    m = warp_max(m);
    if (lane == 0) red[0][w] = m;
    __syncthreads();
    if (w == 0) {
      m = lane < 8 ? red[0][lane] : 0.f;
      m = warp_max(m);
      if (lane == 0) {
        float up = 1.f;
        if (m > 0.f && isfinite(m)) {
          int e;
          frexpf(m, &e);
          e = e > 126 ? 126 : (e < -126 ? -126 : e);
          up = __int_as_float((e + 127) << 23);
        }
        o.out2[0] = up;
        s_down = 1.0f / up;
      }
    }
    __syncthreads();
    const float down = s_down;
    for (int s = o.lo + t; s < hi; s += 256) o.out[s] = o.a[s] * o.b[s] * o.c[s] * down;
  }
}

__global__ void row_stat_kernel(const float* __restrict__ norms, int n_rows, int n_slots, int slot_lo, int slot_hi,
                                int stat, float scalar, float* __restrict__ out) {
  __shared__ float sh[32];
  const int k = blockIdx.x;
  const float* row = norms + static_cast<long long>(k) * n_slots;
  float acc = stat ? -INFINITY : 0.f;
  for (int s = slot_lo + threadIdx.x; s < slot_hi; s += blockDim.x) acc = stat ? fmaxf(acc, row[s]) : acc + row[s];
  if (stat) {
    acc = warp_max(acc);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
      acc = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : -INFINITY;
      acc = warp_max(acc);
    }
  } else {
    acc = block_sum(acc, sh);
    acc /= static_cast<float>(slot_hi - slot_lo);
  }
  if (threadIdx.x == 0) out[k] = acc * scalar;
}

// ------------------------------------------------------------------------------------------
// Philox noise, bit-compatible with torch's CUDA normal_ (see header comment of the ABI).
// torch launches, per tensor, block 256 and grid = min(SMs*(maxThreads/256), ceil(n/256)); thread `idx` runs
// curand_init(seed, idx, offset) and in loop trip k writes elements idx + nthreads*(4k + ii), ii < 4, from ONE
// curand_normal4.  Philox4_32_10 is counter based: that draw is philox(key = seed, ctr = {offset/4 + k, idx}), so
// the launch geometry is free.  One launch covers every tensor of the step: a block handles 256 consecutive `idx`
// of one (tensor, k) and finds its tensor in a table of at most kNoiseMaxSegs entries.
// ------------------------------------------------------------------------------------------
constexpr int kNoiseMaxSegs = 24;

struct NoiseSeg {
  const float* in;
  float* grad;
  long long n;
  const float* std_dev;
  float std_mult;
  unsigned int tgrid;                  // torch's grid for this tensor: nthreads = 256 * tgrid
  unsigned long long off4;             // (offset of this tensor inside the launch) / 4
  long long blk0;                      // first work block of this segment; work blocks = trips * tgrid
  int draws;                           // 0: no noise for this segment (std == 0 on the host path)
};

struct NoiseParams {
  NoiseSeg seg[kNoiseMaxSegs];
  int n_segs;
  long long n_blocks;
  float in_mul, noise_mul;             // fp32 reciprocals (<= 0: no division)
  const float* in_div_dev;             // device scalars overriding the two above
  const float* noise_div_dev;
  unsigned long long seed, offset;
  const unsigned long long* offset_dev;
  // fused allreduce (kAllReduce): every seg.in / seg.grad lies in ONE symmetric buffer that starts at `local_base` on
  // this rank and is mapped at the NVSwitch multicast address `mc_base` on all `world` ranks
  const float* local_base;
  float* mc_base;
  long long count_off;                 // element of the buffer that holds each rank's live sample count (-1: none)
  int rank, world;
  // peer-to-peer variant (mc_base == nullptr): the same buffer on every rank through its NVLink peer mapping; the sum
  // is formed in rank order from plain 16-byte loads and the result stored to every peer
  float* peer[8];
  int ld_mc, st_mc;                    // loads / stores through the multicast mapping (when mc_base is set); else peers
};



// NVLink SHARP (multimem) accesses on a multicast address: the load returns the SUM of the word over all ranks'
// copies (reduced inside the switch), the store writes the word into all ranks' copies.
__device__ __forceinline__ float multimem_ld_sum(const float* mc) {
  float v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];" : "=f"(v) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float* mc, float v) {
  asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mc), "f"(v) : "memory");
}
// 16-byte forms (one NVLink request per four words: the 4-byte forms above ran the exchange at a fifth of this rate)
__device__ __forceinline__ float4 multimem_ld_sum4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st4(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w) : "memory");
}

__device__ __forceinline__ float4 ar_load4(const NoiseParams& p, long long off) {
  if (p.ld_mc) return multimem_ld_sum4(p.mc_base + off);
  // all loads first (eight independent 16-byte requests in flight per thread), then the sum in rank order
  float4 b[8];
#pragma unroll
  for (int r = 0; r < 8; ++r)
    b[r] = r < p.world ? __ldcg(reinterpret_cast<const float4*>(p.peer[r] + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 a = b[0];
#pragma unroll
  for (int r = 1; r < 8; ++r) {
    if (r < p.world) { a.x += b[r].x; a.y += b[r].y; a.z += b[r].z; a.w += b[r].w; }
  }
  return a;
}
__device__ __forceinline__ float ar_load1(const NoiseParams& p, long long off) {
  if (p.ld_mc) return multimem_ld_sum(p.mc_base + off);
  float a = p.peer[0][off];
  for (int r = 1; r < p.world; ++r) a += p.peer[r][off];
  return a;
}
__device__ __forceinline__ void ar_store4(const NoiseParams& p, long long off, float4 v) {
  if (p.st_mc) { multimem_st4(p.mc_base + off, v); return; }
#pragma unroll
  for (int r = 0; r < 8; ++r)
    if (r < p.world) __stcg(reinterpret_cast<float4*>(p.peer[r] + off), v);
}
__device__ __forceinline__ void ar_store1(const NoiseParams& p, long long off, float v) {
  if (p.st_mc) { multimem_st(p.mc_base + off, v); return; }
  for (int r = 0; r < p.world; ++r) p.peer[r][off] = v;
}

// kAllReduce = false: grad = in/div + noise on this GPU's own buffers.
// kAllReduce = true : the data-parallel exchange and the noise in ONE kernel over NVSwitch multicast memory: rank r
//   owns the work blocks b = r (mod world); for its elements it loads the all-rank sum of the clipped gradients with
//   multimem.ld_reduce, divides by the all-rank sample count, adds the noise -- drawn from the same Philox counters
//   the single-GPU launch would use, so only 1/world of the normals is generated per rank -- and multicasts the result
//   into every rank's gradient (in place).  Callers bracket the launch with cross-rank barriers.
template <bool kAllReduce>
__global__ void __launch_bounds__(256, 4)
noise_multi_kernel(const __grid_constant__ NoiseParams p) {
  float in_mul = p.in_mul, noise_mul = p.noise_mul;
  if (kAllReduce) {
    if (p.count_off >= 0) {
      const float cnt = ar_load1(p, p.count_off);
      const float r = __fdiv_rn(1.0f, cnt);
      if (in_mul > 0.f || p.in_div_dev) in_mul = r;
      if (noise_mul > 0.f || p.noise_div_dev) noise_mul = r;
    }
  } else {
    if (p.in_div_dev) in_mul = __fdiv_rn(1.0f, p.in_div_dev[0]);
    if (p.noise_div_dev) noise_mul = __fdiv_rn(1.0f, p.noise_div_dev[0]);
  }
  unsigned long long base = p.offset;
  if (p.offset_dev) base += p.offset_dev[0];
  const unsigned long long base4 = base >> 2;
  const uint2 key = make_uint2(static_cast<unsigned int>(p.seed), static_cast<unsigned int>(p.seed >> 32));
  const long long b_first = kAllReduce ? static_cast<long long>(blockIdx.x) * p.world + p.rank : blockIdx.x;
  const long long b_step = kAllReduce ? static_cast<long long>(gridDim.x) * p.world : gridDim.x;
  for (long long b = b_first; b < p.n_blocks; b += b_step) {
    int s = 0;
#pragma unroll 1
    while (s + 1 < p.n_segs && b >= p.seg[s + 1].blk0) ++s;
    const NoiseSeg& sg = p.seg[s];
    const long long bl = b - sg.blk0;
    const long long k = bl / sg.tgrid;
    const long long idx = (bl - k * sg.tgrid) * 256 + threadIdx.x;
    const long long nthreads = 256LL * sg.tgrid;
    const long long li0 = idx + nthreads * 4 * k;
    if (!kAllReduce && li0 >= sg.n) continue;
    float zz[4] = {0.f, 0.f, 0.f, 0.f};
    if (sg.draws && li0 < sg.n) {
      float stdv = sg.std_mult;
      if (sg.std_dev) stdv = __fmul_rn(stdv, sg.std_dev[0]);
      const unsigned long long c = base4 + sg.off4 + static_cast<unsigned long long>(k);
      const uint4 ctr = make_uint4(static_cast<unsigned int>(c), static_cast<unsigned int>(c >> 32),
                                   static_cast<unsigned int>(idx), static_cast<unsigned int>(idx >> 32));
      const uint4 r = curand_Philox4x32_10(ctr, key);
      const float2 a = _curand_box_muller(r.x, r.y);
      const float2 bq = _curand_box_muller(r.z, r.w);
      zz[0] = a.x; zz[1] = a.y; zz[2] = bq.x; zz[3] = bq.y;
#pragma unroll
      for (int ii = 0; ii < 4; ++ii) {
        zz[ii] = __fmul_rn(zz[ii], stdv);                          // torch.normal(0, std): rand * std + 0
        if (noise_mul > 0.f) zz[ii] = __fmul_rn(zz[ii], noise_mul);   // noise /= batch_size (torch CUDA: * 1/B)
      }
    }
    if (kAllReduce) {
      // torch's element order gives a thread four words `nthreads` apart; the exchange wants 16 contiguous bytes per
      // thread.  The block's work is four runs of 256 consecutive elements: the normals meet in shared memory and
      // thread t then owns words [4q, 4q+4) of run t/64 (q = t%64) -- one multimem.ld_reduce.v4, one multimem.st.v4.
      __shared__ float s_z[4][256];
#pragma unroll
      for (int ii = 0; ii < 4; ++ii) s_z[ii][threadIdx.x] = zz[ii];
      __syncthreads();
      const int run = threadIdx.x >> 6, q4 = (threadIdx.x & 63) << 2;
      const long long e0 = (bl - k * sg.tgrid) * 256 + nthreads * (4 * k + run) + q4;     // element of the segment
      if (e0 < sg.n) {
        const long long oin = sg.in ? (sg.in - p.local_base) + e0 : 0;
        const long long oout = (sg.grad - p.local_base) + e0;
        const float4 z = *reinterpret_cast<const float4*>(&s_z[run][q4]);
        if (e0 + 3 < sg.n && ((oin | oout) & 3) == 0) {
          float4 v = z;
          if (sg.in) {
            v = ar_load4(p, oin);
            if (in_mul > 0.f) { v.x = __fmul_rn(v.x, in_mul); v.y = __fmul_rn(v.y, in_mul); v.z = __fmul_rn(v.z, in_mul); v.w = __fmul_rn(v.w, in_mul); }
            if (sg.draws) { v.x = __fadd_rn(v.x, z.x); v.y = __fadd_rn(v.y, z.y); v.z = __fadd_rn(v.z, z.z); v.w = __fadd_rn(v.w, z.w); }
          }
          ar_store4(p, oout, v);
        } else {
          const float zs[4] = {z.x, z.y, z.z, z.w};
          for (int j = 0; j < 4 && e0 + j < sg.n; ++j) {
            float v = zs[j];
            if (sg.in) {
              v = ar_load1(p, oin + j);
              if (in_mul > 0.f) v = __fmul_rn(v, in_mul);
              if (sg.draws) v = __fadd_rn(v, zs[j]);
            }
            ar_store1(p, oout + j, v);
          }
        }
      }
      __syncthreads();                              // s_z is rewritten by the next work block
      continue;
    }
    float g[4];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const long long li = li0 + nthreads * ii;
      g[ii] = (sg.in && li < sg.n) ? sg.in[li] : 0.f;
    }
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const long long li = li0 + nthreads * ii;
      if (li < sg.n) {
        float v = zz[ii];
        if (sg.in) {
          v = g[ii];
          if (in_mul > 0.f) v = __fmul_rn(v, in_mul);               // p.grad = summed_grad / batch_size
          if (sg.draws) v = __fadd_rn(v, zz[ii]);                   // p.grad += noise
        }
        sg.grad[li] = v;
      }
    }
  }
}

__global__ void philox_advance_kernel(unsigned long long* offset_dev, unsigned long long inc) {
  if (threadIdx.x == 0 && blockIdx.x == 0) offset_dev[0] += inc;
}

// ------------------------------------------------------------------------------------------
// row-norm backward, vector max, fused per-sample L2 clip
// ------------------------------------------------------------------------------------------
// grid (column chunks, row lanes): a row is spread over gridDim.x blocks, so a single long row (rows = 1,
// cols = |theta|) still fills the machine
__global__ void row_l2_norm_bwd_kernel(const float* __restrict__ g, const float* __restrict__ norms,
                                       const float* __restrict__ gout, long long rows, long long cols,
                                       float* __restrict__ gin) {
  const bool vec = (cols & 3) == 0 && ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(gin)) & 15) == 0;
  for (long long r = blockIdx.y; r < rows; r += gridDim.y) {
    const float n = norms[r];
    const float s = n > 0.f ? gout[r] / n : 0.f;
    const float* gr = g + r * cols;
    float* o = gin + r * cols;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const long long i0 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (vec) {
      const long long n4 = cols >> 2;
      for (long long i = i0; i < n4; i += stride) {
        float4 v = __ldg(reinterpret_cast<const float4*>(gr) + i);
        v.x *= s; v.y *= s; v.z *= s; v.w *= s;
        reinterpret_cast<float4*>(o)[i] = v;
      }
    } else {
      for (long long i = i0; i < cols; i += stride) o[i] = gr[i] * s;
    }
  }
}

__global__ void vec_max_kernel(const float* __restrict__ v, long long n, float* __restrict__ out) {
  __shared__ float sh[32];
  float acc = -INFINITY;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) acc = fmaxf(acc, v[i]);
  acc = warp_max(acc);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : -INFINITY;
    acc = warp_max(acc);
    if (threadIdx.x == 0) out[0] = acc;
  }
}

// one block per row: norm, then out = norm > C ? C * (t / norm) : t   (same op order as the reference)
__global__ void l2_clip_kernel(const float* __restrict__ t, long long rows, long long cols, float C,
                               float* __restrict__ out, float* __restrict__ norms_out) {
  __shared__ float sh[32];
  __shared__ float s_norm;
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    const float* row = t + r * cols;
    float s = block_sum(row_sumsq_device(row, cols), sh);
    if (threadIdx.x == 0) {
      s_norm = sqrtf(s);
      if (norms_out) norms_out[r] = s_norm;
    }
    __syncthreads();
    const float n = s_norm;
    float* o = out + r * cols;
    if (n > C) {
      for (long long i = threadIdx.x; i < cols; i += blockDim.x) o[i] = __fmul_rn(C, __fdiv_rn(row[i], n));
    } else {
      for (long long i = threadIdx.x; i < cols; i += blockDim.x) o[i] = row[i];
    }
    __syncthreads();
  }
}

}  // namespace cg
