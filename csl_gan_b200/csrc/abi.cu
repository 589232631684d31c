// extern "C" entry points of libcslgan_b200.so (see include/cslgan_b200.h for the contract).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/cslgan_b200.h"
#include "contract.cuh"
#include "ghost.cuh"
#include "ghost2.cuh"
#include "cl.cuh"
#include "cl_pair.cuh"
#include "stage2.cuh"
#include "thin.cuh"
#include "cl_res.cuh"
#include "kernels.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

#define CG_CHECK(expr)                                                                            \
  do {                                                                                            \
    cudaError_t e__ = (expr);                                                                     \
    if (e__ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

#define CG_LAUNCH_CHECK()                                                                         \
  do {                                                                                            \
    cudaError_t e__ = cudaGetLastError();                                                         \
    if (e__ != cudaSuccess) return fail("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

struct DevInfo {
  bool ok = false;
  int sm = 0, max_thr = 0, major = 0, minor = 0;
};

int dev_info(DevInfo* out) {
  static std::mutex mu;
  static DevInfo cache[64];
  int dev = 0;
  CG_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail("device ordinal %d out of range", dev);
  std::lock_guard<std::mutex> lk(mu);
  if (!cache[dev].ok) {
    CG_CHECK(cudaDeviceGetAttribute(&cache[dev].sm, cudaDevAttrMultiProcessorCount, dev));
    CG_CHECK(cudaDeviceGetAttribute(&cache[dev].max_thr, cudaDevAttrMaxThreadsPerMultiProcessor, dev));
    CG_CHECK(cudaDeviceGetAttribute(&cache[dev].major, cudaDevAttrComputeCapabilityMajor, dev));
    CG_CHECK(cudaDeviceGetAttribute(&cache[dev].minor, cudaDevAttrComputeCapabilityMinor, dev));
    cache[dev].ok = true;
  }
  *out = cache[dev];
  return 0;
}

inline cudaStream_t S(cg_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// grid for a grid-stride elementwise kernel: enough blocks to cover `n`, capped at 8 waves of the SMs
int ew_grid(long long n, int block, int sm) {
  long long g = (n + block - 1) / block;
  const long long cap = static_cast<long long>(sm) * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode(EncodeTiledFn* fn) {
  static EncodeTiledFn cached = nullptr;
  if (!cached) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CG_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess || !p) return fail("cuTensorMapEncodeTiled not available from the driver");
    cached = reinterpret_cast<EncodeTiledFn>(p);
  }
  *fn = cached;
  return 0;
}

// 2-D fp32 matrix [rows][cols] with row pitch `pitch` floats; box = 32 cols x box_rows, SWIZZLE_128B.
int make_tmap(CUtensorMap* tm, const float* base, long long rows, long long cols, long long pitch, int box_rows) {
  EncodeTiledFn enc;
  if (get_encode(&enc)) return 1;
  if (reinterpret_cast<uintptr_t>(base) & 15) return fail("operand base pointer must be 16-byte aligned");
  if ((pitch * 4) % 16) return fail("operand pitch (%lld floats) must be a multiple of 4", pitch);
  if (cols > pitch) return fail("operand cols (%lld) exceed pitch (%lld)", cols, pitch);
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(pitch) * 4};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(cg::kBK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld pitch=%lld box_rows=%d)",
                                     static_cast<int>(r), rows, cols, pitch, box_rows);
  return 0;
}

// fp32 reciprocal of a divisor, exactly as ATen's CUDA div-by-scalar computes it (1.0f / b); <= 0 disables
float recip(double div) { return div > 0.0 ? 1.0f / static_cast<float>(div) : 0.0f; }

int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// generic tiled tensor map (fp32, SWIZZLE_128B); dims/strides innermost first, strides in bytes for dims 1..rank-1
int make_tmap_nd(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                 const cuuint32_t* box, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B, bool half = false) {
  EncodeTiledFn enc;
  if (get_encode(&enc)) return 1;
  if (reinterpret_cast<uintptr_t>(base) & 15) return fail("tensor base pointer must be 16-byte aligned");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank,
                   const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled (rank %d) failed with CUresult %d", rank, static_cast<int>(r));
  return 0;
}

// residues / shifts of one spatial axis: index = s*(pos + a) + rho  for tap k  (k*d - pad = s*a + rho)
void axis_plan(int K, int s, int d, int pad, int* a, int* j_of, int* rho, int* n_rho, int* a_min, int* a_max) {
  *n_rho = 0; *a_min = 1 << 30; *a_max = -(1 << 30);
  for (int k = 0; k < K; ++k) {
    const int r = k * d - pad;
    a[k] = floor_div(r, s);
    const int rr = r - a[k] * s;
    if (a[k] < *a_min) *a_min = a[k];
    if (a[k] > *a_max) *a_max = a[k];
    int j = -1;
    for (int t = 0; t < *n_rho; ++t) if (rho[t] == rr) j = t;
    if (j < 0) { j = *n_rho; rho[(*n_rho)++] = rr; }
    j_of[k] = j;
  }
}

// amax[slot0 + n] = max |src[n][...]| of a [B][C][H][W] tensor addressed through strides (FP16 staging scale)
int sample_absmax(const float* src, long long sn, long long sc, long long sh, long long sw, int B, int C, int H, int W,
                  unsigned int* amax, int slot0, int sm_count, cg_stream_t stream) {
  CG_CHECK(cudaMemsetAsync(amax + slot0, 0, sizeof(unsigned int) * B, S(stream)));
  const long long len = static_cast<long long>(C) * H * W;
  long long extent = 1;
  if (C > 1) extent += static_cast<long long>(C - 1) * sc;
  if (H > 1) extent += static_cast<long long>(H - 1) * sh;
  if (W > 1) extent += static_cast<long long>(W - 1) * sw;
  if (B > 65535) return fail("batch too large for the absmax grid");
  if (sc >= 0 && sh >= 0 && sw >= 0 && extent == len) {
    // the sample is dense in memory (NCHW or NHWC contiguous): one linear float4 sweep, ~4 blocks per SM overall
    long long parts = (4LL * sm_count + B - 1) / B;
    const long long max_parts = (len + 2047) / 2048;
    if (parts > max_parts) parts = max_parts;
    if (parts < 1) parts = 1;
    dim3 grid(static_cast<unsigned>(parts), static_cast<unsigned>(B));
    cg::absmax_dense_kernel<<<grid, 256, 0, S(stream)>>>(src, sn, len, amax, slot0);
  } else {
    long long parts = (len + 255) / 256;
    if (parts > 64) parts = 64;
    dim3 grid(static_cast<unsigned>(parts), static_cast<unsigned>(B));
    cg::absmax_strided_kernel<<<grid, 256, 0, S(stream)>>>(src, sn, sc, sh, sw, C, H, W, amax, slot0);
  }
  CG_LAUNCH_CHECK();
  return 0;
}

// Tensor maps of the channels-last operands (MN-major tiles).  rb = bytes of one 32-channel chunk row.
//   Xt[m/32][row][m%32]: one box = kb_rows rows of four consecutive chunks -> smem [chunk][row][32]
//   Yt[plane*n_cb + c/32][slot][hs][ws][c%32]: one box = a tap window of `cpt` consecutive chunks
int make_cl_tmaps(CUtensorMap* tx, CUtensorMap* ty, const cg_cl_desc* d, const cg_cl_plan* plan, int n_cb, int kb_rows,
                  int kb_w, int kb_h, int kb_s, int cpt) {
  const bool half = d->half != 0;
  const cuuint32_t cw = half ? 64 : 32;                    // channels per 128-byte chunk row
  if (plan->cw != static_cast<int>(cw)) return fail("plan chunk width %d does not match the operand type", plan->cw);
  const cuuint64_t rb = 128;
  const CUtensorMapSwizzle sw = half ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  {
    cuuint64_t dims[3] = {cw, static_cast<cuuint64_t>(d->xt_rows), static_cast<cuuint64_t>((d->M + cw - 1) / cw)};
    cuuint64_t str[2] = {rb, static_cast<cuuint64_t>(d->xt_rows) * rb};
    cuuint32_t box[3] = {cw, static_cast<cuuint32_t>(kb_rows), 128 / cw};
    if (make_tmap_nd(tx, d->Xt, 3, dims, str, box, sw, half)) return 1;
  }
  {
    const cuuint64_t slot_bytes = static_cast<cuuint64_t>(plan->slot_stride) * (half ? 2 : 4);
    cuuint64_t dims[5] = {cw, static_cast<cuuint64_t>(plan->Ws), static_cast<cuuint64_t>(plan->Hs),
                          static_cast<cuuint64_t>(d->n_slots_total),
                          static_cast<cuuint64_t>(plan->n_rh * plan->n_rw) * n_cb};
    cuuint64_t str[4] = {rb, rb * plan->Ws, slot_bytes, slot_bytes * d->n_slots_total};
    cuuint32_t box[5] = {cw, static_cast<cuuint32_t>(kb_w), static_cast<cuuint32_t>(kb_h),
                         static_cast<cuuint32_t>(kb_s), static_cast<cuuint32_t>(cpt)};
    if (make_tmap_nd(ty, d->Yt, 5, dims, str, box, sw, half)) return 1;
  }
  return 0;
}

// CTA-pair launch of the split-K clipped sum (cl_pair.cuh); `p` is the single-CTA parameter block already filled
template <bool kHalf>
int launch_pair_t(const cg_cl_desc* d, const cg_cl_plan* plan, const cg::ClParams& p, int kb_rows, int kb_w, int kb_h,
                  int kb_s, cg_stream_t stream) {
  cg::ClPairParams q;
  memset(&q, 0, sizeof(q));
  q.M = d->M; q.n_mp = d->M / 256;
  q.C = p.C; q.n_cb = p.n_cb; q.n_taps = p.n_taps;
  constexpr int kHalfChunks = cg::PairCfg<kHalf>::kHalfChunks;      // chunks per 128 channels
  q.hpt = p.n_cb / kHalfChunks; q.n_ht = p.n_taps * q.hpt; q.n_nt = (q.n_ht + 1) / 2;
  for (int t = 0; t < p.n_taps; ++t) { q.tap_plane[t] = p.tap_plane[t]; q.tap_hoff[t] = p.tap_hoff[t]; q.tap_woff[t] = p.tap_woff[t]; }
  q.Q = p.Q; q.Wo = p.Wo; q.kb_s = p.kb_s; q.nkb_slot = p.nkb_slot; q.kb_rows = kb_rows;
  q.oob_chunk = plan->n_rh * plan->n_rw * p.n_cb;
  q.u_lo = p.u_lo; q.u_hi = p.u_hi; q.upg = p.upg; q.n_groups = p.n_groups;
  q.out = p.out; q.ldT = p.ldT; q.out_scale = p.out_scale;
  q.n_items = static_cast<long long>(q.n_groups) * q.n_nt * q.n_mp;
  CUtensorMap tx, ty;
  if (make_cl_tmaps(&tx, &ty, d, plan, p.n_cb, kb_rows, kb_w, kb_h, kb_s, kHalfChunks)) return 1;
  DevInfo dv;
  if (dev_info(&dv)) return 1;
  static int max_pairs[64] = {0};
  int dev = 0;
  CG_CHECK(cudaGetDevice(&dev));
  constexpr int kSmem = cg::PairCfg<kHalf>::kSmemBytes;
  if (!max_pairs[dev]) {
    CG_CHECK(cudaFuncSetAttribute(cg::cl_pair_kernel<kHalf>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * (dv.sm / 2)); cfg.blockDim = dim3(cg::kClThreads); cfg.dynamicSmemBytes = kSmem;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    int n = 0;
    CG_CHECK(cudaOccupancyMaxActiveClusters(&n, cg::cl_pair_kernel<kHalf>, &cfg));
    if (n < 1) return fail("cl_pair_kernel: no CTA pair can be resident on this device");
    max_pairs[dev] = n;
  }
  long long pairs = max_pairs[dev];
  if (d->max_ctas > 0 && d->max_ctas / 2 < pairs) pairs = d->max_ctas / 2 > 0 ? d->max_ctas / 2 : 1;
  if (pairs > q.n_items) pairs = q.n_items;
  cg::cl_pair_kernel<kHalf><<<static_cast<int>(2 * pairs), cg::kClThreads, kSmem, S(stream)>>>(tx, ty, q);
  CG_LAUNCH_CHECK();
  return 0;
}

template <bool kHalf>
int launch_cl_t(const cg_cl_desc* d, const cg_cl_plan* plan, const cg::ClParams& p, int kb_rows, int kb_w, int kb_h,
                int kb_s, int sm_count, cg_stream_t stream) {
  CUtensorMap tx, ty;
  if (make_cl_tmaps(&tx, &ty, d, plan, p.n_cb, kb_rows, kb_w, kb_h, kb_s, p.cpt)) return 1;
  static bool attr_set[64] = {false};
  int dev = 0;
  CG_CHECK(cudaGetDevice(&dev));
  constexpr int kSmem = cg::ClCfg<kHalf>::kSmemBytes;
  if (!attr_set[dev]) {
    CG_CHECK(cudaFuncSetAttribute(cg::cl_contract_kernel<kHalf>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    attr_set[dev] = true;
  }
  long long grid = d->max_ctas > 0 ? d->max_ctas : sm_count;
  if (grid > p.n_items) grid = p.n_items;
  cg::cl_contract_kernel<kHalf><<<static_cast<int>(grid), cg::kClThreads, kSmem, S(stream)>>>(tx, ty, p);
  CG_LAUNCH_CHECK();
  return 0;
}

// launch `kernel` on B * parts CTAs in clusters of `parts` (1, 2, 4 or 8)
template <typename... KArgs, typename... Args>
int launch_clustered(void (*kernel)(KArgs...), int n_ctas, int parts, cg_stream_t stream, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(static_cast<unsigned>(n_ctas));
  cfg.blockDim = dim3(cg::kFusedThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = S(stream);
  cudaLaunchAttribute at;
  at.id = cudaLaunchAttributeClusterDimension;
  at.val.clusterDim.x = static_cast<unsigned>(parts); at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
  cfg.attrs = &at; cfg.numAttrs = 1;
  CG_CHECK(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
  return 0;
}

// CTAs per sample of the single-pass FP16 capture (16 float4 per thread, 256 threads): 0 = sample too large
static int fused_parts(long long floats_per_sample) {
  const long long per = 4LL * cg::kFusedPerCta;
  for (int parts = 1; parts <= 8; parts *= 2)
    if (floats_per_sample <= per * parts) return parts;
  return 0;
}

static bool fused_enabled() {
  static const bool on = [] { const char* e = getenv("CSLGAN_FUSED_STAGE"); return !(e && e[0] == '0'); }();
  return on;
}

// FP16 capture route: 2 = two sweeps inside one kernel (stage2.cuh, default), 1 = sample held in registers (cl.cuh's
// single-pass kernels), 0 = separate absmax kernel + staging kernel.  CSLGAN_STAGE=sweep|fused|two is the A/B switch.
static int stage_route() {
  static const int r = [] {
    const char* e = getenv("CSLGAN_STAGE");
    if (e && !strcmp(e, "fused")) return 1;
    if (e && !strcmp(e, "two")) return 0;
    return 2;
  }();
  return fused_enabled() ? r : 0;
}

// CTAs per sample (a power of two <= 8, = the cluster size) and float4 per CTA (a multiple of 256) of the two-sweep
// capture: about CSLGAN_SWEEP_PER (default 4096 = 16 per thread per sweep;
// measured 1024 / 2048 / 4096 / 8192: 4096 is the fastest on every CelebA layer) float4 per CTA, more when 8 CTAs are not enough
static void sweep_split(long long len4, int* parts, int* per) {
  static const int target = [] { const char* e = getenv("CSLGAN_SWEEP_PER"); int v = e ? atoi(e) : 0; return v >= 256 ? v : 4096; }();
  int p = 1;
  while (p < 8 && static_cast<long long>(p) * target < len4) p *= 2;
  long long q = (len4 + p - 1) / p;
  q = (q + 255) / 256 * 256;
  *parts = p; *per = static_cast<int>(q);
}

// capture kernels of the channels-last path for either element type
template <typename T>
int stage_xt_t(const float* src, long long sn, long long sm, long long sh, long long sw, int B, int M, int Ho, int Wo,
               float scale, T* dst, long long rows_total, int slot0, float* bias_rows, float* sumsq,
               unsigned int* amax, float* inv, cg_stream_t stream) {
  if (B <= 0 || M <= 0) return 0;
  DevInfo d;
  if (dev_info(&d)) return 1;
  const int Q = Ho * Wo;
  if (Q == 1 && sm == 1 && sn >= M) {
    // Linear layers: rows [B][M]; one warp per row does the maximum, the scale, the sums and the staging
    if (static_cast<long long>(slot0 + B) > rows_total) return fail("Xt too small for slots [%d, %d)", slot0, slot0 + B);
    if (sizeof(T) == 2 && !inv) return fail("FP16 staging needs the inverse-scale output");
    // threads per row: one (tiny rows), a warp, or a whole block (wide rows, few of them)
    if (M <= 16) {
      const long long blocks = (static_cast<long long>(B) + 255) / 256;
      cg::stage_rows_cl_kernel<T, 1><<<static_cast<unsigned>(blocks), 256, 0, S(stream)>>>(src, sn, B, M, scale, dst,
                                                                                            rows_total, slot0, bias_rows, sumsq, inv);
    } else if (M >= 2048) {
      cg::stage_rows_cl_kernel<T, 256><<<static_cast<unsigned>(B), 256, 0, S(stream)>>>(src, sn, B, M, scale, dst,
                                                                                         rows_total, slot0, bias_rows, sumsq, inv);
    } else {
      const long long blocks = (static_cast<long long>(B) + 7) / 8;
      cg::stage_rows_cl_kernel<T, 32><<<static_cast<unsigned>(blocks), 256, 0, S(stream)>>>(src, sn, B, M, scale, dst,
                                                                                             rows_total, slot0, bias_rows, sumsq, inv);
    }
    CG_LAUNCH_CHECK();
    return 0;
  }
  if (bias_rows) CG_CHECK(cudaMemsetAsync(bias_rows + static_cast<long long>(slot0) * M, 0, sizeof(float) * B * M, S(stream)));
  if (sumsq) CG_CHECK(cudaMemsetAsync(sumsq + slot0, 0, sizeof(float) * B, S(stream)));
  if (static_cast<long long>(slot0 + B) * Q > rows_total) return fail("Xt too small for slots [%d, %d)", slot0, slot0 + B);
  if (sizeof(T) == 2) {
    if (!amax || !inv) return fail("FP16 staging needs the amax scratch and the inverse-scale output");
    // single pass when the sample is a dense channels-last block that a cluster of <= 8 CTAs can hold in registers
    const int mv = M / 4;
    const bool dense_cl = sm == 1 && (M % 4) == 0 && mv <= cg::kFusedThreads && (cg::kFusedThreads % mv) == 0 && sw == M &&
                          (Ho == 1 || sh == static_cast<long long>(Wo) * M) && (sn % 4) == 0 && B <= (1 << 24) &&
                          (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    if (stage_route() == 2 && dense_cl && M <= 1024 && static_cast<long long>(Q) * mv < (1LL << 30)) {
      int sp, per;
      sweep_split(static_cast<long long>(Q) * mv, &sp, &per);
      if (launch_clustered(cg::stage_xt_sweep_kernel, B * sp, sp, stream, src, sn, M, Q, scale,
                           reinterpret_cast<__half*>(dst), rows_total, slot0, bias_rows, inv, sp, per)) return 1;
      CG_LAUNCH_CHECK();
      return 0;
    }
    const int parts = fused_parts(static_cast<long long>(Q) * M);
    if (stage_route() == 1 && sm == 1 && (M % 4) == 0 && mv <= cg::kFusedThreads && (cg::kFusedThreads % mv) == 0 && sw == M &&
        (Ho == 1 || sh == static_cast<long long>(Wo) * M) && (sn % 4) == 0 && parts > 0 && B <= (1 << 24) &&
        (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
      if (launch_clustered(cg::stage_xt_fused_kernel, B * parts, parts, stream, src, sn, M, Q, scale,
                           reinterpret_cast<__half*>(dst), rows_total, slot0, bias_rows, inv, parts)) return 1;
      CG_LAUNCH_CHECK();
      return 0;
    }
    if (sample_absmax(src, sn, sm, sh, sw, B, M, Ho, Wo, amax, slot0, d.sm, stream)) return 1;
  }
  int block = ((M < 256 ? M : 256) + 31) / 32 * 32;
  // positions per block: keep ~8 blocks per SM in flight without shredding the bias sums into atomics
  int qchunks = static_cast<int>((8LL * d.sm + B - 1) / B);
  if (qchunks < 1) qchunks = 1;
  if (qchunks > Q) qchunks = Q;
  int qpb = (Q + qchunks - 1) / qchunks;
  dim3 grid(B, (Q + qpb - 1) / qpb);             // batch on grid.x (no 65535 limit), position chunks on grid.y
  const bool vec4 = sm == 1 && (M % 4) == 0 && (sn % 4) == 0 && (sh % 4) == 0 && (sw % 4) == 0 &&
                    (reinterpret_cast<uintptr_t>(src) & 15) == 0;
  if (vec4) {
    const int mv = M / 4;
    int vblock = 256;                                         // rows wider than 1024 channels: block-sized strips
    if (mv <= 256) {
      vblock = mv >= 128 ? ((mv + 31) / 32 * 32) : 128;       // at least 128 threads: several positions in parallel
      if (vblock % mv) vblock = (vblock / mv + 1) * mv;       // whole groups of channel vectors
      if (vblock > 1024) vblock = mv;
    }
    if (mv > 256)
      cg::stage_xt_vec4_wide_kernel<T><<<grid, vblock, 0, S(stream)>>>(src, sn, sh, sw, M, Wo, Q, scale, dst, rows_total,
                                                                        slot0, bias_rows, sumsq, qpb, amax, inv);
    else
      cg::stage_xt_vec4_kernel<T><<<grid, vblock, 0, S(stream)>>>(src, sn, sh, sw, M, Wo, Q, scale, dst, rows_total,
                                                                   slot0, bias_rows, sumsq, qpb, amax, inv);
  } else if (sm == 1 && (M % 2) == 0 && (sn % 2) == 0 && (sh % 2) == 0 && (sw % 2) == 0 &&
             (reinterpret_cast<uintptr_t>(src) & 7) == 0) {
    cg::stage_xt_vec2_kernel<T><<<grid, 256, 0, S(stream)>>>(src, sn, sh, sw, M, Wo, Q, scale, dst, rows_total, slot0,
                                                             bias_rows, sumsq, qpb, amax, inv);
  } else {
    cg::stage_xt_kernel<T><<<grid, block, 0, S(stream)>>>(src, sn, sm, sh, sw, M, Wo, Q, scale, dst, rows_total, slot0,
                                                          bias_rows, sumsq, qpb, amax, inv);
  }
  CG_LAUNCH_CHECK();
  return 0;
}

// Thin inputs with the whole window folded into the channel axis (plan->merged == 2), FP16: stage2.cuh's shared-memory
// im2col.  Returns 0 = launched, 1 = geometry not covered (caller falls back), -1 = error (message set).
int stage_yt_window(const float* src, long long sn, long long sc, long long sh, long long sw, int B, const cg_unfold_geom* g,
                    const cg_cl_plan* plan, float scale, __half* dst, int n_slots_total, int slot0, float* inv,
                    int sm_count, cg_stream_t stream) {
  const int n_oct = (plan->Cs + 7) / 8;
  if (plan->merged != 2 || plan->cw != 64 || n_oct > 32 || B > (1 << 24)) return 1;
  const long long span = static_cast<long long>(g->C - 1) * sc + static_cast<long long>(g->H - 1) * sh +
                         static_cast<long long>(g->W - 1) * sw;
  if (sc < 0 || sh < 0 || sw < 0 || span >= (1LL << 31)) return 1;
  cg::YwParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.C = g->C; p.H = g->H; p.W = g->W;
  p.sn = sn; p.sc = sc; p.sh_ = sh; p.sw_ = sw;
  p.KH = g->KH; p.KW = g->KW; p.sth = g->sh; p.stw = g->sw; p.ph = g->ph; p.pw = g->pw; p.dh = g->dh; p.dw = g->dw;
  p.Ho = g->Ho; p.Wo = g->Wo;
  p.Cs = plan->Cs; p.n_oct = n_oct;
  p.Wp = (g->Wo - 1) * g->sw + (g->KW - 1) * g->dw + 1;
  p.scale = scale; p.slot0 = slot0;
  p.slot_stride = plan->slot_stride;
  p.chunk_stride = plan->slot_stride * n_slots_total;
  // output rows per CTA: as few CTAs per sample as a 64 KB shared image and a covered machine allow (measured at
  // B = 512 on the 3 x 64 x 64 image: 60 / 64 / 82 / 111 us for 1 / 2 / 4 / 8 CTAs per sample -- halo rows and the
  // cluster exchange cost more than the extra CTAs bring)
  auto smem_for = [&](int parts) {
    const int rpp = (g->Ho + parts - 1) / parts;
    const long long rows = static_cast<long long>(rpp - 1) * g->sh + static_cast<long long>(g->KH - 1) * g->dh + 1;
    return rows * p.Wp * g->C * static_cast<long long>(sizeof(float));
  };
  static const int want = [] { const char* e = getenv("CSLGAN_WINDOW_PARTS"); return e ? atoi(e) : 0; }();
  int parts = 1;
  while (parts < 8 && g->Ho >= 2 * parts &&
         (static_cast<long long>(B) * parts < 2LL * sm_count || smem_for(parts) > 64 * 1024)) parts *= 2;
  if (want == 1 || want == 2 || want == 4 || want == 8) parts = want;
  const long long smem = smem_for(parts);
  if (smem > 200 * 1024) return 1;
  p.rpp = (g->Ho + parts - 1) / parts;
  p.rows_max = static_cast<int>(smem / (static_cast<long long>(p.Wp) * g->C * sizeof(float)));
  static long long attr_smem[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { fail("cudaGetDevice failed"); return -1; }
  if (smem > attr_smem[dev]) {
    if (cudaFuncSetAttribute(cg::stage_yt_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(smem)) != cudaSuccess) { fail("stage_yt_window: shared memory attribute"); return -1; }
    attr_smem[dev] = smem;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(static_cast<unsigned>(B) * parts);
  cfg.blockDim = dim3(32 * n_oct);
  cfg.dynamicSmemBytes = static_cast<size_t>(smem);
  cfg.stream = S(stream);
  cudaLaunchAttribute at;
  at.id = cudaLaunchAttributeClusterDimension;
  at.val.clusterDim.x = static_cast<unsigned>(parts); at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
  cfg.attrs = &at; cfg.numAttrs = 1;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, cg::stage_yt_window_kernel, src, p, dst, inv, parts);
  if (e != cudaSuccess) { fail("stage_yt_window launch: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}

template <typename T>
int stage_yt_t(const float* src, long long sn, long long sc, long long sh, long long sw, int B, const cg_unfold_geom* g,
               const cg_cl_plan* plan, float scale, T* dst, int n_slots_total, int slot0, unsigned int* amax,
               float* inv, cg_stream_t stream) {
  if (!g || !plan) return fail("null argument");
  if (B <= 0) return 0;
  DevInfo d;
  if (dev_info(&d)) return 1;
  const int n_planes = plan->n_rh * plan->n_rw;
  constexpr int kCW = 128 / sizeof(T), kLPP = kCW / 4;      // channels per chunk row, lanes per position
  if (plan->cw != kCW) return fail("plan chunk width %d does not match the staged element type", plan->cw);
  const int n_cb = plan->Cp / kCW;
  if ((B + cg::kYtSamples - 1) / cg::kYtSamples > 65535 || n_planes * n_cb > 65535)
    return fail("problem too large for the staging grid");
  if (sizeof(T) == 2 && (!amax || !inv)) return fail("FP16 staging needs the amax scratch and the inverse-scale output");
  cg::YtParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.C = g->C; p.H = g->H; p.W = g->W;
  p.sn = sn; p.sc = sc; p.sh_ = sh; p.sw_ = sw;
  p.Cs = plan->Cs; p.n_cb = n_cb; p.merged = plan->merged; p.KW = g->KW; p.dw = g->dw; p.pw = g->pw;
  p.dh = g->dh; p.ph = g->ph;
  p.Hs = plan->Hs; p.Ws = plan->Ws; p.n_rh = plan->n_rh; p.n_rw = plan->n_rw; p.sth = g->sh; p.stw = g->sw;
  p.ah_min = plan->ah_min; p.aw_min = plan->aw_min;
  for (int i = 0; i < CG_MAX_KH; ++i) { p.rho_h[i] = plan->rho_h[i]; p.rho_w[i] = plan->rho_w[i]; }
  p.scale = scale; p.slot0 = slot0;
  p.slot_stride = plan->slot_stride;
  p.chunk_stride = plan->slot_stride * n_slots_total;
  const int n_pos = plan->Hs * plan->Ws;
  const long long extent = static_cast<long long>(g->C - 1) * sc + static_cast<long long>(g->H - 1 + g->KH * g->dh) * sh +
                           static_cast<long long>(g->W - 1 + g->KW * g->dw) * sw;
  if (sc < 0 || sh < 0 || sw < 0 || extent >= (1LL << 31) || plan->slot_stride >= (1LL << 31))
    return fail("cg_stage_yt: one sample must span fewer than 2^31 elements with non-negative strides");
  const bool vec4 = !plan->merged && sc == 1 && (g->C % 4) == 0 && (sn % 4) == 0 && (sh % 4) == 0 && (sw % 4) == 0 &&
                    (reinterpret_cast<uintptr_t>(src) & 15) == 0;
  if (sizeof(T) == 2) {
    // single pass when the sample is a dense NHWC block that a cluster of <= 8 CTAs can hold in registers
    const int parts = fused_parts(static_cast<long long>(g->C) * g->H * g->W);
    const int cvn = g->C / 4;
    const bool dense_nhwc = vec4 && sw == g->C && sh == static_cast<long long>(g->W) * g->C && B <= (1 << 24) &&
                            cvn <= cg::kFusedThreads && (cg::kFusedThreads % cvn) == 0 && g->H <= cg::kYtFusedMaxDim &&
                            g->W <= cg::kYtFusedMaxDim;
    if (stage_route() == 2 && dense_nhwc && static_cast<long long>(g->H) * g->W * cvn < (1LL << 30)) {
      int sp, per;
      sweep_split(static_cast<long long>(g->H) * g->W * cvn, &sp, &per);
      if (launch_clustered(cg::stage_yt_sweep_kernel, B * sp, sp, stream, src, p, reinterpret_cast<__half*>(dst), inv,
                           sp, per)) return 1;
      CG_LAUNCH_CHECK();
      return 0;
    }
    if (stage_route() == 2 && plan->merged == 2) {
      const int rc = stage_yt_window(src, sn, sc, sh, sw, B, g, plan, scale, reinterpret_cast<__half*>(dst), n_slots_total,
                                     slot0, inv, d.sm, stream);
      if (rc <= 0) return -rc;                            // 0: launched, < 0: error; > 0: geometry not covered, fall through
    }
    if (stage_route() == 1 && vec4 && sw == g->C && sh == static_cast<long long>(g->W) * g->C && parts > 0 && B <= (1 << 24) &&
        cvn <= cg::kFusedThreads && (cg::kFusedThreads % cvn) == 0 && g->H <= cg::kYtFusedMaxDim &&
        g->W <= cg::kYtFusedMaxDim) {
      if (launch_clustered(cg::stage_yt_fused_kernel, B * parts, parts, stream, src, p, reinterpret_cast<__half*>(dst), inv,
                           parts)) return 1;
      CG_LAUNCH_CHECK();
      return 0;
    }
    if (sample_absmax(src, sn, sc, sh, sw, B, g->C, g->H, g->W, amax, slot0, d.sm, stream)) return 1;
  }
  // threads = kLPP lanes per position; small window grids get a block that covers them in k equal steps
  // (36 positions -> 288 threads x 1 step, 100 -> 416 x 2) instead of idling most of a 256-thread block
  int threads = 256, ppb = 128;
  if (n_pos < 128) {
    int k = 1;
    while (kLPP * ((n_pos + k - 1) / k) > 512) ++k;
    threads = (kLPP * ((n_pos + k - 1) / k) + 31) / 32 * 32;
    ppb = n_pos;
  }
  dim3 grid((n_pos + ppb - 1) / ppb, (B + cg::kYtSamples - 1) / cg::kYtSamples, n_planes * n_cb);
  if (vec4) cg::stage_yt_kernel<true, T><<<grid, threads, 0, S(stream)>>>(src, p, dst, ppb, amax, inv);
  else cg::stage_yt_kernel<false, T><<<grid, threads, 0, S(stream)>>>(src, p, dst, ppb, amax, inv);
  CG_LAUNCH_CHECK();
  return 0;
}

template <bool kHalf>
int ghost_norm_t(const cg_ghost_desc* d, const cg_unfold_geom* g, const cg_ghost_plan* plan, int sm_count,
                        cg_stream_t stream) {
  const int Q = g->Ho * g->Wo;
  cg::GhostParams p;
  memset(&p, 0, sizeof(p));
  p.Q = Q; p.ns = 128 / Q; p.O = d->O; p.C = g->C; p.KH = g->KH; p.KW = g->KW;
  for (int t = 0; t < g->KH * g->KW; ++t) {
    p.tap_plane[t] = plan->tap_plane[t]; p.tap_hoff[t] = plan->tap_hoff[t]; p.tap_woff[t] = plan->tap_woff[t];
  }
  p.slot0 = d->slot0; p.n_slots = d->n_slots;
  p.n_items = (d->n_slots + p.ns - 1) / p.ns;
  p.norm2 = d->norm2;
  p.inv_x = d->inv_x; p.inv_y = d->inv_y;
  // K-major tiles: 128 rows of one 128-byte chunk (32 TF32 / 64 FP16 channels), SWIZZLE_128B
  constexpr cuuint32_t cw = kHalf ? 64 : 32;
  if (plan->cw != static_cast<int>(cw)) return fail("plan chunk width %d does not match the operand type", plan->cw);
  const cuuint64_t rb = 128;
  const CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
  const cuuint64_t slot_bytes = static_cast<cuuint64_t>(plan->slot_stride) * (kHalf ? 2 : 4);

  CUtensorMap tx, ty;
  {
    // Xt[o/cw][row][o%cw]: a K-major tile = 128 rows of one chunk
    cuuint64_t dims[3] = {cw, static_cast<cuuint64_t>(d->xt_rows), static_cast<cuuint64_t>((d->O + cw - 1) / cw)};
    cuuint64_t str[2] = {rb, static_cast<cuuint64_t>(d->xt_rows) * rb};
    cuuint32_t box[3] = {cw, 128, 1};
    if (make_tmap_nd(&tx, d->Xt, 3, dims, str, box, sw, kHalf)) return 1;
  }
  const cuuint64_t n_cb = (g->C + cw - 1) / cw;
  cuuint64_t ydims[5] = {cw, static_cast<cuuint64_t>(plan->Ws), static_cast<cuuint64_t>(plan->Hs),
                         static_cast<cuuint64_t>(d->n_slots_total),
                         static_cast<cuuint64_t>(plan->n_rh * plan->n_rw) * n_cb};
  cuuint64_t ystr[4] = {rb, rb * plan->Ws, slot_bytes, slot_bytes * d->n_slots_total};
  const int npos = plan->Hs * plan->Ws;
  static const bool force_v1 = [] { const char* e = getenv("CSLGAN_GHOST_V1"); return e && e[0] == '1'; }();
  if (npos <= 128 && plan->Ws <= 256 && plan->Hs <= 256 && !force_v1) {
    // Gram of the un-shifted planes + tap gather in the epilogue (ghost2.cuh)
    cg::Ghost2Params q;
    memset(&q, 0, sizeof(q));
    q.Q = Q; q.ns = 128 / Q; q.Wo = g->Wo; q.O = d->O; q.C = g->C;
    q.n_planes = plan->n_rh * plan->n_rw; q.Hs = plan->Hs; q.Ws = plan->Ws; q.npos = npos;
    q.spp = 128 / npos < q.ns ? 128 / npos : q.ns;
    q.n_sub = (q.ns + q.spp - 1) / q.spp;
    int n = 0;
    for (int pl = 0; pl < q.n_planes; ++pl) {
      q.plane_tap0[pl] = n;
      for (int t = 0; t < g->KH * g->KW; ++t)
        if (plan->tap_plane[t] == pl) q.tap_shift[n++] = plan->tap_hoff[t] * plan->Ws + plan->tap_woff[t];
    }
    q.plane_tap0[q.n_planes] = n;
    q.slot0 = d->slot0; q.n_slots = d->n_slots; q.n_items = p.n_items; q.norm2 = d->norm2;
    q.inv_x = d->inv_x; q.inv_y = d->inv_y;
    cuuint32_t box[5] = {cw, static_cast<cuuint32_t>(plan->Ws), static_cast<cuuint32_t>(plan->Hs),
                         static_cast<cuuint32_t>(q.spp), 1};
    if (make_tmap_nd(&ty, d->Yt, 5, ydims, ystr, box, sw, kHalf)) return 1;
    const int smem = cg::g2_smem_bytes(Q);
    CG_CHECK(cudaFuncSetAttribute(cg::ghost2_norm_kernel<kHalf>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int grid2 = d->max_ctas > 0 ? d->max_ctas : sm_count;
    if (grid2 > q.n_items) grid2 = q.n_items;
    cg::ghost2_norm_kernel<kHalf><<<grid2, cg::kG2Threads, smem, S(stream)>>>(tx, ty, q);
    CG_LAUNCH_CHECK();
    return 0;
  }
  {
    cuuint32_t box[5] = {cw, static_cast<cuuint32_t>(g->Wo), static_cast<cuuint32_t>(g->Ho),
                         static_cast<cuuint32_t>(p.ns), 1};
    if (make_tmap_nd(&ty, d->Yt, 5, ydims, ystr, box, sw, kHalf)) return 1;
  }
  static bool attr_set[64] = {false};
  int dev = 0;
  CG_CHECK(cudaGetDevice(&dev));
  if (!attr_set[dev]) {
    CG_CHECK(cudaFuncSetAttribute(cg::ghost_norm_kernel<kHalf>, cudaFuncAttributeMaxDynamicSharedMemorySize, cg::kGSmemBytes));
    attr_set[dev] = true;
  }
  int grid = d->max_ctas > 0 ? d->max_ctas : sm_count;
  if (grid > p.n_items) grid = p.n_items;
  cg::ghost_norm_kernel<kHalf><<<grid, cg::kGThreads, cg::kGSmemBytes, S(stream)>>>(tx, ty, p);
  CG_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" {

int cg_version(void) { return 100; }

const char* cg_last_error(void) { return g_err; }

int cg_device_info(int* sm_count, int* max_threads_per_sm, int* cc_major, int* cc_minor) {
  DevInfo d;
  if (dev_info(&d)) return 1;
  if (sm_count) *sm_count = d.sm;
  if (max_threads_per_sm) *max_threads_per_sm = d.max_thr;
  if (cc_major) *cc_major = d.major;
  if (cc_minor) *cc_minor = d.minor;
  return 0;
}

int cg_plan_unfold(const cg_unfold_geom* g, cg_unfold_plan* plan) {
  if (!g || !plan) return fail("null argument");
  if (g->KH < 1 || g->KH > CG_MAX_KH || g->KW < 1) return fail("unsupported filter size %dx%d", g->KH, g->KW);
  if (g->sh < 1 || g->sw < 1 || g->dh < 1 || g->dw < 1) return fail("stride/dilation must be >= 1");
  memset(plan, 0, sizeof(*plan));
  int a[CG_MAX_KH], rho_of[CG_MAX_KH];
  int a_min = 1 << 30, a_max = -(1 << 30);
  int n_rho = 0;
  for (int kh = 0; kh < g->KH; ++kh) {
    const int r = kh * g->dh - g->ph;
    a[kh] = floor_div(r, g->sh);
    rho_of[kh] = r - a[kh] * g->sh;
    if (a[kh] < a_min) a_min = a[kh];
    if (a[kh] > a_max) a_max = a[kh];
    int j = -1;
    for (int t = 0; t < n_rho; ++t)
      if (plan->rho[t] == rho_of[kh]) j = t;
    if (j < 0) {
      j = n_rho;
      plan->rho[n_rho++] = rho_of[kh];
    }
    plan->tap_row0[kh] = j * g->KW * g->C;
  }
  plan->n_rho = n_rho;
  plan->a_min = a_min;
  plan->Hs = g->Ho + a_max - a_min;
  plan->Wop = (g->Wo + 3) / 4 * 4;
  plan->rows = n_rho * g->KW * g->C;
  plan->slot_stride = plan->Hs * plan->Wop;
  for (int kh = 0; kh < g->KH; ++kh) plan->tap_coloff[kh] = (a[kh] - a_min) * plan->Wop;
  return 0;
}

int cg_stage_rows_t(const float* src, int B, int R, float scale, float* dst, long long dst_pitch, int slot0,
                    float* copy_out, float* sumsq, cg_stream_t stream) {
  if (B <= 0 || R <= 0) return 0;
  if (sumsq) CG_CHECK(cudaMemsetAsync(sumsq + slot0, 0, sizeof(float) * B, S(stream)));
  dim3 grid((R + 31) / 32, (B + 31) / 32), block(32, 8);
  cg::stage_rows_t_kernel<<<grid, block, 0, S(stream)>>>(src, B, R, scale, dst, dst_pitch, slot0, copy_out, sumsq);
  CG_LAUNCH_CHECK();
  return 0;
}

int cg_stage_rows(const float* src, int B, int R, int Q, int Wo, int Wop, int Qpad, float scale, float* dst,
                  long long dst_pitch, int slot0, float* rowsum, cg_stream_t stream) {
  if (B <= 0 || R <= 0) return 0;
  if (Wo <= 0 || Wop < Wo || Q % Wo) return fail("bad window row geometry Q=%d Wo=%d Wop=%d", Q, Wo, Wop);
  if (Qpad < (Q / Wo) * Wop) return fail("Qpad (%d) too small for Q=%d Wo=%d Wop=%d", Qpad, Q, Wo, Wop);
  DevInfo d;
  if (dev_info(&d)) return 1;
  const long long nrows = static_cast<long long>(B) * R;
  const int block = 256;
  long long g = (nrows + 7) / 8;
  if (g > static_cast<long long>(d.sm) * 16) g = static_cast<long long>(d.sm) * 16;
  cg::stage_rows_kernel<<<static_cast<int>(g), block, 0, S(stream)>>>(src, B, R, Q, Wo, Wop, Qpad, scale, dst, dst_pitch,
                                                                      slot0, rowsum);
  CG_LAUNCH_CHECK();
  return 0;
}

int cg_stage_unfold(const float* src, int B, const cg_unfold_geom* g, const cg_unfold_plan* plan, float scale,
                    float* dst, long long dst_pitch, int slot0, cg_stream_t stream) {
  if (!g || !plan) return fail("null argument");
  if (B <= 0) return 0;
  DevInfo d;
  if (dev_info(&d)) return 1;
  cg::UnfoldParams p;
  p.B = B; p.C = g->C; p.H = g->H; p.W = g->W; p.KW = g->KW;
  p.sh = g->sh; p.sw = g->sw; p.pw = g->pw; p.dw = g->dw;
  p.Wo = g->Wo; p.Wop = plan->Wop; p.Hs = plan->Hs; p.n_rho = plan->n_rho; p.a_min = plan->a_min;
  for (int j = 0; j < CG_MAX_KH; ++j) p.rho[j] = plan->rho[j];
  p.scale = scale; p.dst_pitch = dst_pitch; p.slot0 = slot0;
  const long long hw = static_cast<long long>(g->H) * g->W;
  if (hw <= 8192 && B <= 65535) {
    // shared-memory path: each block reads `cpb` input planes once and emits all their kw-planes
    int cpb = static_cast<int>(8192 / hw);
    if (cpb > g->C) cpb = g->C;
    if (cpb < 1) cpb = 1;
    // keep enough blocks in flight: at least ~4 waves when the problem allows it
    while (cpb > 1 && static_cast<long long>((g->C + cpb - 1) / cpb) * B < 4LL * d.sm) cpb = (cpb + 1) / 2;
    dim3 grid((g->C + cpb - 1) / cpb, B);
    const size_t smem = static_cast<size_t>(cpb) * hw * sizeof(float);
    cg::stage_unfold_smem_kernel<<<grid, 256, smem, S(stream)>>>(src, p, dst, cpb);
  } else {
    const long long total = static_cast<long long>(plan->rows) * B * plan->slot_stride;
    cg::stage_unfold_kernel<<<ew_grid(total, 256, d.sm), 256, 0, S(stream)>>>(src, p, dst);
  }
  CG_LAUNCH_CHECK();
  return 0;
}

int cg_contract(const cg_contract_desc* d, cg_stream_t stream) {
  if (!d) return fail("null descriptor");
  DevInfo dv;
  if (dev_info(&dv)) return 1;
  if (dv.major != 10) return fail("cg_contract needs an sm_100-class device (found sm_%d%d)", dv.major, dv.minor);
  if (d->KH < 1 || d->KH > CG_MAX_KH) return fail("KH out of range");
  if (d->n_groups <= 0 || d->nkb <= 0) return 0;
  cg::ContractParams p;
  memset(&p, 0, sizeof(p));
  p.M = d->M;
  p.n_mtiles = (d->M + cg::kBM - 1) / cg::kBM;
  p.KWC = d->KW * d->C;
  int bn = d->block_n;
  if (bn <= 0) {
    // widest tile that wastes the fewest padded columns: ceil(KWC / ceil(KWC/256)) rounded up to 16
    const int parts = (p.KWC + cg::kMaxBN - 1) / cg::kMaxBN;
    bn = (((p.KWC + parts - 1) / parts) + 15) / 16 * 16;
  }
  if (bn % 16 || bn < 16 || bn > cg::kMaxBN) return fail("block_n must be a multiple of 16 in [16,256]");
  p.BN = bn;
  p.n_rb = (p.KWC + bn - 1) / bn;
  // narrow layers: stack several filter-row taps along N so the X tile is fetched once for all of them
  p.ks = 1;
  if (p.n_rb == 1) {
    p.ks = cg::kMaxBN / bn;
    if (p.ks > d->KH) p.ks = d->KH;
    if (p.ks < 1) p.ks = 1;
  }
  p.n_khg = (d->KH + p.ks - 1) / p.ks;
  p.NT = p.ks * bn;
  p.C = d->C; p.KH = d->KH; p.KW = d->KW;
  for (int i = 0; i < CG_MAX_KH; ++i) { p.tap_row0[i] = d->tap_row0[i]; p.tap_coloff[i] = d->tap_coloff[i]; }
  p.nkb = d->nkb;
  p.x_slot_stride = d->x_slot_stride; p.y_slot_stride = d->y_slot_stride;
  p.group_mode = d->group_mode; p.n_groups = d->n_groups;
  p.slot_lo = d->slot_lo; p.slot_hi = d->slot_hi; p.spg = d->spg;
  p.n_seg = d->n_seg; p.seg_stride = d->seg_stride;
  p.epi = d->epi; p.out = d->out; p.out_group_stride = d->out_group_stride;
  if (p.group_mode == CG_GROUP_SPLITK) {
    if (p.spg <= 0) return fail("spg must be positive for split-K groups");
    if (static_cast<long long>(p.n_groups - 1) * p.spg >= (p.slot_hi - p.slot_lo)) return fail("empty split-K group");
  } else if (p.n_seg <= 0) {
    return fail("n_seg must be positive");
  }
  p.n_items = static_cast<long long>(p.n_groups) * p.n_khg * p.n_rb * p.n_mtiles;

  CUtensorMap tx, ty;
  if (make_tmap(&tx, d->X, d->x_rows, d->x_cols, d->x_pitch, cg::kBM)) return 1;
  if (make_tmap(&ty, d->Y, d->y_rows, d->y_cols, d->y_pitch, bn)) return 1;

  static bool attr_set[64] = {false};
  int dev = 0;
  CG_CHECK(cudaGetDevice(&dev));
  if (!attr_set[dev]) {
    CG_CHECK(cudaFuncSetAttribute(cg::contract_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cg::kSmemBytes));
    attr_set[dev] = true;
  }
  long long grid = d->max_ctas > 0 ? d->max_ctas : dv.sm;
  if (grid > p.n_items) grid = p.n_items;
  cg::contract_kernel<<<static_cast<int>(grid), cg::kThreads, cg::kSmemBytes, S(stream)>>>(tx, ty, p);
  CG_LAUNCH_CHECK();
  return 0;
}

int cg_plan_cl(const cg_unfold_geom* g, int merged, cg_cl_plan* plan) { return cg_plan_cl_cw(g, merged, 32, plan); }

int cg_plan_cl_cw(const cg_unfold_geom* g, int merged, int cw, cg_cl_plan* plan) {
  if (!g || !plan) return fail("null argument");
  if (cw != 32 && cw != 64) return fail("chunk width must be 32 (TF32) or 64 (FP16) channels");
  if (g->KH < 1 || g->KH > CG_MAX_KH || g->KW < 1 || g->KW > CG_MAX_KH) return fail("unsupported filter size");
  if (g->sh < 1 || g->sw < 1 || g->dh < 1 || g->dw < 1) return fail("stride/dilation must be >= 1");
  memset(plan, 0, sizeof(*plan));
  int ah[CG_MAX_KH], aw[CG_MAX_KH], jh[CG_MAX_KH], jw[CG_MAX_KH];
  int ah_min, ah_max, aw_min, aw_max;
  axis_plan(g->KH, g->sh, g->dh, g->ph, ah, jh, plan->rho_h, &plan->n_rh, &ah_min, &ah_max);
  plan->ah_min = ah_min;
  plan->Hs = g->Ho + ah_max - ah_min;
  plan->merged = merged == 2 ? 2 : (merged ? 1 : 0);
  if (merged == 2) {
    // the whole filter window lives in the channel axis (c' = (kh*KW + kw)*C + c): plain im2col rows, ONE tap, one
    // TMA box per k-block; used for thin inputs with FP16 operands (3 x 5 x 5 = 75 staged channels in 2 chunks of 64)
    plan->n_rh = 1; plan->rho_h[0] = 0; plan->ah_min = 0; plan->Hs = g->Ho;
    plan->n_rw = 1; plan->rho_w[0] = 0; plan->aw_min = 0; plan->Ws = g->Wo;
    plan->Cs = g->KH * g->KW * g->C;
    plan->n_taps = 1;
    plan->tap_plane[0] = 0; plan->tap_hoff[0] = 0; plan->tap_woff[0] = 0;
  } else if (merged) {
    // filter columns live in the channel axis: one plane per row residue, window rows of exactly Wo
    plan->n_rw = 1; plan->rho_w[0] = 0; plan->aw_min = 0;
    plan->Ws = g->Wo;
    plan->Cs = g->KW * g->C;
    plan->n_taps = g->KH;
    for (int kh = 0; kh < g->KH; ++kh) {
      plan->tap_plane[kh] = jh[kh];
      plan->tap_hoff[kh] = ah[kh] - ah_min;
      plan->tap_woff[kh] = 0;
    }
  } else {
    axis_plan(g->KW, g->sw, g->dw, g->pw, aw, jw, plan->rho_w, &plan->n_rw, &aw_min, &aw_max);
    plan->aw_min = aw_min;
    plan->Ws = g->Wo + aw_max - aw_min;
    plan->Cs = g->C;
    plan->n_taps = g->KH * g->KW;
    for (int kh = 0; kh < g->KH; ++kh)
      for (int kw = 0; kw < g->KW; ++kw) {
        const int t = kh * g->KW + kw;
        plan->tap_plane[t] = jh[kh] * plan->n_rw + jw[kw];
        plan->tap_hoff[t] = ah[kh] - ah_min;
        plan->tap_woff[t] = aw[kw] - aw_min;
      }
  }
  plan->cw = cw;
  plan->Cp = (plan->Cs + cw - 1) / cw * cw;
  plan->slot_stride = static_cast<long long>(plan->Hs) * plan->Ws * cw;   // elements per slot inside one chunk
  return 0;
}

int cg_plan_ghost(const cg_unfold_geom* g, cg_ghost_plan* plan) {
  if (!g || !plan) return fail("null argument");
  const int Q = g->Ho * g->Wo;
  if (Q < 1 || Q > 128 || (128 % Q) != 0) return fail("ghost norms need Ho*Wo to divide 128 (got %d)", Q);
  if (g->Wo > 256 || g->Ho > 256) return fail("window extent too large for a TMA box");
  return cg_plan_cl_cw(g, 0, 32, plan);
}

int cg_ghost_norm(const cg_ghost_desc* d, const cg_unfold_geom* g, const cg_ghost_plan* plan, cg_stream_t stream) {
  if (!d || !g || !plan) return fail("null argument");
  DevInfo dv;
  if (dev_info(&dv)) return 1;
  if (dv.major != 10) return fail("cg_ghost_norm needs an sm_100-class device (found sm_%d%d)", dv.major, dv.minor);
  if (d->n_slots <= 0) return 0;
  const int Q = g->Ho * g->Wo;
  if (Q < 1 || (128 % Q) != 0) return fail("ghost norms need Ho*Wo to divide 128");
  if (d->half && (!d->inv_x || !d->inv_y)) return fail("cg_ghost_norm: FP16 operands need inv_x / inv_y");
  return d->half ? ghost_norm_t<true>(d, g, plan, dv.sm, stream) : ghost_norm_t<false>(d, g, plan, dv.sm, stream);
}

// k-block geometry of the channels-last path for k-blocks of KB contraction rows; non-zero = not tileable
static int cl_kblock_kb(const cg_unfold_geom* g, int per_sample, int KB, int min_rows, int* kb_rows, int* kb_w,
                        int* kb_h, int* kb_s) {
  const int Q = g->Ho * g->Wo;
  if (Q >= KB) {
    if (Q % KB) return 1;
    const int w = g->Wo < KB ? g->Wo : KB;
    if (g->Wo % w || KB % w) return 1;
    *kb_w = w; *kb_h = KB / w; *kb_s = 1; *kb_rows = KB;
    if (g->Ho % *kb_h) return 1;
    return 0;
  }
  if (KB % Q) return 1;
  *kb_w = g->Wo; *kb_h = g->Ho;
  if (per_sample) {
    if (Q % min_rows) return 1;
    *kb_s = 1; *kb_rows = Q;
  } else {
    if (KB / Q > 32) return 1;               // slot ranges are 32-aligned (Bpad), not more
    *kb_s = KB / Q; *kb_rows = KB;
  }
  return 0;
}

// FP16 operands use k-blocks of 64 rows where the window grid allows (the single-thread TMA / MMA roles pay a fixed
// cost per k-block that is comparable with 32 rows of FP16 MMA time), 32 rows otherwise; TF32 always 32.
static int cl_kblock(const cg_unfold_geom* g, int per_sample, int* kb_rows, int* kb_w, int* kb_h, int* kb_s,
                     int half = 0) {
  const int min_rows = half ? 16 : 8;
  if (half && cl_kblock_kb(g, per_sample, 64, min_rows, kb_rows, kb_w, kb_h, kb_s) == 0) return 0;
  if (cl_kblock_kb(g, per_sample, 32, min_rows, kb_rows, kb_w, kb_h, kb_s) == 0) return 0;
  const int Q = g->Ho * g->Wo;
  return fail("channels-last path cannot tile a %dx%d window grid (Q = %d) into k-blocks%s", g->Ho, g->Wo, Q,
              per_sample ? " (per-sample groups need a multiple of 8 / 16 positions)" : "");
}

int cg_cl_kblock_rows(const cg_unfold_geom* g, int half, int* kb_rows, int* kb_s) {
  if (!g || !kb_rows || !kb_s) return fail("null argument");
  int w, h;
  return cl_kblock(g, 0, kb_rows, &w, &h, kb_s, half);
}

int cg_stage_xt(const float* src, long long sn, long long sm, long long sh, long long sw, int B, int M, int Ho,
                int Wo, float scale, float* dst, long long rows_total, int slot0, float* bias_rows, float* sumsq,
                cg_stream_t stream) {
  return stage_xt_t<float>(src, sn, sm, sh, sw, B, M, Ho, Wo, scale, dst, rows_total, slot0, bias_rows, sumsq, nullptr,
                           nullptr, stream);
}

int cg_stage_xt_h(const float* src, long long sn, long long sm, long long sh, long long sw, int B, int M, int Ho,
                  int Wo, float scale, void* dst_half, long long rows_total, int slot0, float* bias_rows, float* sumsq,
                  unsigned int* amax, float* inv, cg_stream_t stream) {
  return stage_xt_t<__half>(src, sn, sm, sh, sw, B, M, Ho, Wo, scale, static_cast<__half*>(dst_half), rows_total, slot0,
                            bias_rows, sumsq, amax, inv, stream);
}

int cg_stage_yt(const float* src, long long sn, long long sc, long long sh, long long sw, int B,
                const cg_unfold_geom* g, const cg_cl_plan* plan, float scale, float* dst, int n_slots_total,
                int slot0, cg_stream_t stream) {
  return stage_yt_t<float>(src, sn, sc, sh, sw, B, g, plan, scale, dst, n_slots_total, slot0, nullptr, nullptr, stream);
}

int cg_stage_yt_h(const float* src, long long sn, long long sc, long long sh, long long sw, int B,
                  const cg_unfold_geom* g, const cg_cl_plan* plan, float scale, void* dst_half, int n_slots_total,
                  int slot0, unsigned int* amax, float* inv, cg_stream_t stream) {
  return stage_yt_t<__half>(src, sn, sc, sh, sw, B, g, plan, scale, static_cast<__half*>(dst_half), n_slots_total,
                            slot0, amax, inv, stream);
}

// ---- thin first convolution: per-sample gradients straight from the critic's tensors (thin.cuh) ------------------
namespace {
int thin_fill(const cg_unfold_geom* g, int M, cg::ThinParams* p) {
  memset(p, 0, sizeof(*p));
  const int Q = g->Ho * g->Wo;
  const int Cs = g->KH * g->KW * g->C;
  const int n_ch = M / 32;
  if (Cs > 128 || M % 32 || (n_ch != 1 && n_ch != 2 && n_ch != 4) || Q % cg::kThinKb || g->Wo % 4 || g->Wo > cg::kThinKb * 64) return 1;
  if (g->ph < 0 || g->pw < 0) return 1;
  p->C = g->C; p->H = g->H; p->W = g->W;
  p->KH = g->KH; p->KW = g->KW; p->sth = g->sh; p->stw = g->sw; p->ph = g->ph; p->pw = g->pw; p->dh = g->dh; p->dw = g->dw;
  p->Ho = g->Ho; p->Wo = g->Wo;
  p->Cs = Cs; p->n_g = (Cs + 7) / 8;
  p->Hp = (g->Ho - 1) * g->sh + (g->KH - 1) * g->dh + 1;
  p->Wp = (g->Wo - 1) * g->sw + (g->KW - 1) * g->dw + 1;
  p->Wp += p->Wp & 1;                                  // even rows and planes: 8-byte cp.async for dense image rows
  // plane stride: the channels of one tap land ~10 banks apart (the lanes of a builder warp read c' = 8g + lane%8)
  int P = p->Hp * p->Wp;
  while ((P % 32) != 10 && (P % 32) != 22) ++P;
  p->P = P;
  p->Hc = g->H < p->Hp - g->ph ? g->H : p->Hp - g->ph;
  p->Wc = g->W < p->Wp - g->pw ? g->W : p->Wp - g->pw;
  if (p->Hc <= 0 || p->Wc <= 0) return 1;
  p->M = M; p->n_ch = n_ch; p->Q = Q; p->nkb = Q / cg::kThinKb;
  p->b_bytes = n_ch * cg::kThinKb * 128;
  p->a_bytes = 2 * p->n_g * 1024;
  p->stage_bytes = p->b_bytes + p->a_bytes;
  p->img_floats = (g->C * P + 3) / 4 * 4;
  int cols = 32;
  while (cols < 2 * M) cols *= 2;
  p->tmem_cols = cols;
  return 0;
}
long long thin_smem(const cg::ThinParams& p) {
  // tiles (+ the 128-row reads of the last atom stay inside the image buffers) + images + tables + barriers
  // (small images: explicit slack so that those reads still end inside the allocation)
  const long long over = (16LL - p.n_g) * 1024 - 2LL * p.img_floats * 4;
  return 1024LL + static_cast<long long>(cg::kThinStages) * p.stage_bytes + 2LL * p.img_floats * 4 + 128 * 4 + 128 * 4 + 256 +
         (over > 0 ? over : 0);
}
}  // namespace

int cg_thin_direct_ok(const cg_unfold_geom* g, int M) {
  if (!g) return 0;
  cg::ThinParams p;
  if (thin_fill(g, M, &p)) return 0;
  return thin_smem(p) <= 227 * 1024 ? 1 : 0;
}

int cg_thin_capture2(const float* act, const float* act2, long long a_sn, long long a_sc, long long a_sh, long long a_sw,
                     const float* bp, const float* bp2, int B, int B2, const cg_unfold_geom* g, int M, float scale, float* Gs,
                     float* Gs2, long long gs_stride, float* norm2, float* norm2_2, float* bias_rows, float* bias_rows2,
                     cg_stream_t stream) {
  if (!act || !bp || !g || !Gs || !norm2) return fail("null argument");
  if (B2 > 0 && (!act2 || !bp2 || !Gs2 || !norm2_2)) return fail("null argument (second segment)");
  if (B2 > 0 && ((bias_rows == nullptr) != (bias_rows2 == nullptr))) return fail("bias rows for one segment only");
  if (B <= 0) return 0;
  if (B2 < 0) B2 = 0;
  cg::ThinParams p;
  if (!cg_thin_direct_ok(g, M) || thin_fill(g, M, &p)) return fail("cg_thin_capture: geometry not covered (cg_thin_direct_ok)");
  if (static_cast<long long>(B > B2 ? B : B2) * p.Q >= (1LL << 31)) return fail("cg_thin_capture: batch too large");
  p.act = act; p.a_sn = a_sn; p.a_sc = a_sc; p.a_sh = a_sh; p.a_sw = a_sw;
  p.B0 = B; p.B = B + B2; p.scale = scale;
  p.Gs = Gs; p.gs_stride = gs_stride; p.norm2 = norm2; p.bias_rows = bias_rows;
  p.act2 = B2 > 0 ? act2 : act; p.Gs2 = B2 > 0 ? Gs2 : Gs; p.norm2_2 = B2 > 0 ? norm2_2 : norm2;
  p.bias_rows2 = B2 > 0 ? bias_rows2 : bias_rows;
  DevInfo dv;
  if (dev_info(&dv)) return 1;
  CG_CHECK(cudaMemsetAsync(norm2, 0, sizeof(float) * B, S(stream)));
  if (B2 > 0) CG_CHECK(cudaMemsetAsync(norm2_2, 0, sizeof(float) * B2, S(stream)));
  // bp: dense channels-last [B*Q rows][M]; box = {32 ch, 64 rows, all chunks} lands as [chunk][row][32 ch]
  CUtensorMap tb, tb2;
  for (int sgi = 0; sgi < 2; ++sgi) {
    const int nb = sgi == 0 ? B : (B2 > 0 ? B2 : B);
    const float* base = sgi == 0 ? bp : (B2 > 0 ? bp2 : bp);
    cuuint64_t dims[3] = {32, static_cast<cuuint64_t>(nb) * p.Q, static_cast<cuuint64_t>(p.n_ch)};
    cuuint64_t str[2] = {static_cast<cuuint64_t>(M) * 4, 128};
    cuuint32_t box[3] = {32, static_cast<cuuint32_t>(cg::kThinKb), static_cast<cuuint32_t>(p.n_ch)};
    if (make_tmap_nd(sgi == 0 ? &tb : &tb2, base, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, false)) return 1;
  }
  const int smem = static_cast<int>(thin_smem(p));
  static bool attr_set[64] = {false};
  int dev = 0;
  CG_CHECK(cudaGetDevice(&dev));
  if (!attr_set[dev]) {
    CG_CHECK(cudaFuncSetAttribute(cg::thin_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set[dev] = true;
  }
  const int grid = p.B < dv.sm ? p.B : dv.sm;
  cg::thin_direct_kernel<<<grid, cg::kThinThreads, smem, S(stream)>>>(tb, tb2, p);
  CG_LAUNCH_CHECK();
  return 0;
}

int cg_thin_capture(const float* act, long long a_sn, long long a_sc, long long a_sh, long long a_sw, const float* bp,
                    int B, const cg_unfold_geom* g, int M, float scale, float* Gs, long long gs_stride, float* norm2,
                    float* bias_rows, cg_stream_t stream) {
  return cg_thin_capture2(act, nullptr, a_sn, a_sc, a_sh, a_sw, bp, nullptr, B, 0, g, M, scale, Gs, nullptr, gs_stride, norm2,
                          nullptr, bias_rows, nullptr, stream);
}

int cg_clip_mult(const float* factor, const float* inv_x, const float* inv_y, int slot_lo, int slot_hi, float* mult,
                 float* out_scale, cg_stream_t stream) {
  if (slot_hi <= slot_lo) return 0;
  if (!factor || !inv_x || !inv_y || !mult || !out_scale) return fail("null argument");
  const int n = slot_hi - slot_lo;
  if (n <= 8192) {
    cg::clip_mult_kernel<<<1, 1024, 0, S(stream)>>>(factor, inv_x, inv_y, slot_lo, slot_hi, mult, out_scale);
    CG_LAUNCH_CHECK();
    return 0;
  }
  // many slots: two multi-block stages meeting in out_scale[1] (raw maximum)
  unsigned int* raw = reinterpret_cast<unsigned int*>(out_scale + 1);
  CG_CHECK(cudaMemsetAsync(raw, 0, sizeof(unsigned int), S(stream)));
  const int blocks = (n + 1023) / 1024 < 592 ? (n + 1023) / 1024 : 592;
  cg::clip_mult_stage1_kernel<<<blocks, 256, 0, S(stream)>>>(factor, inv_x, inv_y, slot_lo, slot_hi, mult, raw);
  CG_LAUNCH_CHECK();
  cg::clip_mult_stage2_kernel<<<blocks, 256, 0, S(stream)>>>(slot_lo, slot_hi, mult, raw, out_scale);
  CG_LAUNCH_CHECK();
  return 0;
}

int cg_scale_slots_h(const void* src_half, void* dst_half, int rows, long long pitch, long long slot_stride, int slot_lo,
                     int slot_hi, const float* mult, cg_stream_t stream) {
  if (rows <= 0 || slot_hi <= slot_lo) return 0;
  DevInfo d;
  if (dev_info(&d)) return 1;
  if (slot_stride > 0x7fffffffLL) return fail("slot_stride too large");
  if (slot_stride % 8 || pitch % 8 || (reinterpret_cast<uintptr_t>(src_half) & 15) || (reinterpret_cast<uintptr_t>(dst_half) & 15))
    return fail("cg_scale_slots_h needs 16-byte aligned rows and slot strides");
  const long long work = static_cast<long long>(slot_hi - slot_lo) * slot_stride / 8;
  long long gx = (work + 255) / 256;
  const long long want = (static_cast<long long>(d.sm) * 8 + rows - 1) / rows;   // ~8 blocks per SM overall
  if (gx > want) gx = want;
  if (gx < 1) gx = 1;
  dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(rows < 65535 ? rows : 65535));
  cg::scale_slots_half_kernel<<<grid, 256, 0, S(stream)>>>(static_cast<const __half*>(src_half),
                                                           static_cast<__half*>(dst_half), rows, pitch,
                                                           static_cast<int>(slot_stride), slot_lo, slot_hi, mult);
  CG_LAUNCH_CHECK();
  return 0;
}

int cg_scale_slots_h_multi(const cg_scale_seg* segs, int n_segs, cg_stream_t stream) {
  if (n_segs <= 0) return 0;
  if (!segs) return fail("null segment table");
  if (n_segs > cg::kScaleMaxSegs) return fail("at most %d segments per cg_scale_slots_h_multi call", cg::kScaleMaxSegs);
  DevInfo d;
  if (dev_info(&d)) return 1;
  cg::ScaleParams p;
  memset(&p, 0, sizeof(p));
  long long total_units = 0;
  for (int i = 0; i < n_segs; ++i)
    if (segs[i].rows > 0 && segs[i].slot_hi > segs[i].slot_lo)
      total_units += static_cast<long long>(segs[i].slot_hi - segs[i].slot_lo) * segs[i].slot_stride / 8 * segs[i].rows;
  if (total_units <= 0) return 0;
  const long long budget = static_cast<long long>(d.sm) * 8;          // ~8 blocks per SM overall, shared by size
  int blk = 0, m = 0;
  for (int i = 0; i < n_segs; ++i) {
    const cg_scale_seg& in = segs[i];
    if (in.rows <= 0 || in.slot_hi <= in.slot_lo) continue;
    if (!in.src || !in.dst || !in.mult) return fail("segment %d: null pointer", i);
    if (in.slot_stride > 0x7fffffffLL || in.slot_stride % 8 || in.pitch % 8 || (reinterpret_cast<uintptr_t>(in.src) & 15) ||
        (reinterpret_cast<uintptr_t>(in.dst) & 15))
      return fail("segment %d: cg_scale_slots_h_multi needs 16-byte aligned rows and slot strides", i);
    cg::ScaleSeg& o = p.seg[m++];
    o.src = static_cast<const __half*>(in.src); o.dst = static_cast<__half*>(in.dst); o.mult = in.mult;
    o.pitch = in.pitch; o.rows = in.rows; o.slot_stride = static_cast<int>(in.slot_stride);
    o.slot_lo = in.slot_lo; o.slot_hi = in.slot_hi;
    const long long units = static_cast<long long>(in.slot_hi - in.slot_lo) * in.slot_stride / 8 * in.rows;
    long long nb = (units * budget + total_units - 1) / total_units;
    const long long cap = (units + 255) / 256;
    if (nb > cap) nb = cap;
    if (nb < 1) nb = 1;
    o.blk0 = blk; o.nblk = static_cast<int>(nb);
    blk += o.nblk;
  }
  p.n_segs = m;
  cg::scale_slots_half_multi_kernel<<<blk, 256, 0, S(stream)>>>(p);
  CG_LAUNCH_CHECK();
  return 0;
}

int cg_cl_pair_ok(int M, const cg_unfold_geom* g, const cg_cl_plan* plan) {
  if (!g || !plan) return 0;
  return (M >= 256 && M % 256 == 0 && plan->Cp >= 128 && plan->Cp % 128 == 0 && !plan->merged) ? 1 : 0;
}

// Per-sample norms with the sample's backprops resident in shared memory and the taps read as shifted windows of one
// plane box (cl_res.cuh).  Returns 0 = launched, 1 = geometry not covered (caller falls back), -1 = error.
static int res_mode() {
  // CSLGAN_RESIDENT: 0 = off; otherwise the flags of cl_res.cuh + 4 (bit 0: descriptor base offset, bit 1: one MMA
  // per tap); default on with the combination validated on B200
  static const int m = [] { const char* e = getenv("CSLGAN_RESIDENT"); return e ? atoi(e) : 4; }();
  return m;
}

// geometry + tile table shared by the two resident kernels; 0 = covered, 1 = not covered
static int res_fill(const cg_cl_desc* d, const cg_unfold_geom* g, const cg_cl_plan* plan, cg::ResParams* pp) {
  cg::ResParams& p = *pp;
  const int Q = g->Ho * g->Wo;
  if (!d->half || plan->merged || plan->cw != 64 || plan->Cp != 64 || d->M > 128 || (g->Wo != 16 && g->Wo != 32 && g->Wo != 64) ||
      Q % cg::kResKb || static_cast<long long>(Q) * 256 > cg::kResXBytes || (d->n_seg > 1))
    return 1;
  memset(&p, 0, sizeof(p));
  p.M = d->M; p.Q = Q; p.Wo = g->Wo; p.Ws = plan->Ws; p.nkb = Q / cg::kResKb; p.kb_h = cg::kResKb / g->Wo; p.n_cb = 1;
  p.y_bytes = 128 * plan->Ws * p.kb_h;
  if (p.y_bytes > cg::kResYStride) return 1;
  // tiles: taps grouped by (plane, row shift), column shifts consecutive
  const int n_taps = plan->n_taps;
  bool used[CG_MAX_KH * CG_MAX_KH] = {false};
  for (int t = 0; t < n_taps; ++t) {
    if (used[t]) continue;
    int lo = plan->tap_woff[t], hi = lo, cnt = 0;
    for (int u = 0; u < n_taps; ++u)
      if (plan->tap_plane[u] == plan->tap_plane[t] && plan->tap_hoff[u] == plan->tap_hoff[t]) {
        used[u] = true; ++cnt;
        if (plan->tap_woff[u] < lo) lo = plan->tap_woff[u];
        if (plan->tap_woff[u] > hi) hi = plan->tap_woff[u];
      }
    if (cnt != hi - lo + 1 || cnt > 4 || p.n_tiles >= cg::kResMaxTiles) return 1;     // shifts must be consecutive
    const int k = p.n_tiles++;
    p.tile_plane[k] = plan->tap_plane[t]; p.tile_hoff[k] = plan->tap_hoff[t];
    p.tile_woff[k] = lo; p.tile_ndw[k] = cnt;
    for (int u = 0; u < n_taps; ++u)
      if (plan->tap_plane[u] == plan->tap_plane[t] && plan->tap_hoff[u] == plan->tap_hoff[t])
        p.tile_tap[k][plan->tap_woff[u] - lo] = u;
  }
  // the last slab of a k-block must end inside the plane box
  for (int t = 0; t < p.n_tiles; ++t) {
    const int last = (p.kb_h - 1) * plan->Ws + (g->Wo - 16) + p.tile_woff[t] + p.tile_ndw[t] - 1 + 15;
    if (last >= plan->Ws * p.kb_h) return 1;
  }
  p.slot_lo = d->slot_lo; p.inv_x = d->inv_x; p.inv_y = d->inv_y;
  return 0;
}

typedef void (*ResKernel)(const CUtensorMap, const CUtensorMap, const cg::ResParams);
static int res_launch(ResKernel kernel, const cg_cl_desc* d, const cg_cl_plan* plan, const cg::ResParams& p, int smem, long long grid,
                      bool* attr_set, cg_stream_t stream) {
  CUtensorMap tx, ty;
  if (make_cl_tmaps(&tx, &ty, d, plan, 1, 64, plan->Ws, p.kb_h, 1, 1)) return -1;
  if (!*attr_set) {
    if (cudaFuncSetAttribute(reinterpret_cast<const void*>(kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
        cudaSuccess) {
      fail("resident kernel: shared memory attribute"); return -1;
    }
    *attr_set = true;
  }
  kernel<<<static_cast<int>(grid), cg::kClThreads, smem, S(stream)>>>(tx, ty, p);
  if (cudaGetLastError() != cudaSuccess) { fail("resident kernel launch failed"); return -1; }
  return 0;
}

int launch_resident_norm(const cg_cl_desc* d, const cg_unfold_geom* g, const cg_cl_plan* plan, int sm_count,
                         cg_stream_t stream) {
  if (!(res_mode() & 4)) return 1;
  cg::ResParams p;
  if (res_fill(d, g, plan, &p)) return 1;
  p.n_groups = d->n_groups; p.out = d->out;
  const int smem = 1024 + 2 * cg::kResXBytes + cg::kResYStages * cg::kResYStride + 512;
  static bool attr_set[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { fail("cudaGetDevice failed"); return -1; }
  long long grid = d->max_ctas > 0 ? d->max_ctas : sm_count;
  if (grid > p.n_groups) grid = p.n_groups;
  return res_launch(cg::cl_resident_norm_kernel, d, plan, p, smem, grid, &attr_set[dev], stream);
}

// split-K clipped sum with resident backprops (cl_res.cuh); out must be the gradient-natural layout
int launch_resident_sum(const cg_cl_desc* d, const cg_unfold_geom* g, const cg_cl_plan* plan, int sm_count,
                        cg_stream_t stream) {
  if (!(res_mode() & 4) || (res_mode() & 8)) return 1;             // CSLGAN_RESIDENT=12: norms only
  cg::ResParams p;
  if (res_fill(d, g, plan, &p)) return 1;
  if (plan->Cs > 64 || d->slot_hi <= d->slot_lo) return 1;
  p.slot_hi = d->slot_hi;
  p.n_tp = (p.n_tiles + 1) / 2;
  long long ctas = d->max_ctas > 0 ? d->max_ctas : sm_count;
  const int n_samples = d->slot_hi - d->slot_lo;
  int G = static_cast<int>(ctas / p.n_tp);
  if (G < 1) G = 1;
  if (G > n_samples) G = n_samples;
  p.spg = (n_samples + G - 1) / G;
  p.n_groups = (n_samples + p.spg - 1) / p.spg;
  p.C = plan->Cs; p.ldT = static_cast<long long>(plan->n_taps) * plan->Cs;
  p.out = d->out; p.out_scale = d->out_scale;
  const int smem = 1024 + 2 * cg::kResXBytes + cg::kResYStages * cg::kResYStride + 4 * cg::kResSumEpiFloats * 4 + 512;
  static bool attr_set[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { fail("cudaGetDevice failed"); return -1; }
  long long grid = ctas;
  if (grid > static_cast<long long>(p.n_tp) * p.n_groups) grid = static_cast<long long>(p.n_tp) * p.n_groups;
  return res_launch(cg::cl_resident_sum_kernel, d, plan, p, smem, grid, &attr_set[dev], stream);
}

int cg_cl_contract(const cg_cl_desc* d, const cg_unfold_geom* g, const cg_cl_plan* plan, cg_stream_t stream) {
  if (!d || !g || !plan) return fail("null argument");
  DevInfo dv;
  if (dev_info(&dv)) return 1;
  if (dv.major != 10) return fail("cg_cl_contract needs an sm_100-class device (found sm_%d%d)", dv.major, dv.minor);
  if (d->n_groups <= 0) return 0;
  const int Q = g->Ho * g->Wo;
  if (d->half && d->group_mode == CG_GROUP_SAMPLE && d->epi == CG_EPI_SUMSQ && !d->pair && d->inv_x && d->inv_y) {
    const int rc = launch_resident_norm(d, g, plan, dv.sm, stream);
    if (rc == 0) return 0;
    if (rc < 0) return 1;
  }
  if (d->half && d->group_mode == CG_GROUP_SPLITK && d->epi == CG_EPI_ACCUM && !d->pair) {
    const int rc = launch_resident_sum(d, g, plan, dv.sm, stream);
    if (rc == 0) return 0;
    if (rc < 0) return 1;
  }
  int kb_rows, kb_w, kb_h, kb_s;
  if (cl_kblock(g, d->group_mode == CG_GROUP_SAMPLE, &kb_rows, &kb_w, &kb_h, &kb_s, d->half)) return 1;
  cg::ClParams p;
  memset(&p, 0, sizeof(p));
  p.M = d->M; p.n_mtiles = (d->M + 127) / 128;
  const int cw = d->half ? 64 : 32;                  // channels per 128-byte chunk row
  const int maxc = 256 / cw;                         // chunks of a 256-column tile
  if (plan->cw != cw) return fail("plan chunk width %d does not match the operand type", plan->cw);
  p.C = plan->Cs; p.n_cb = plan->Cp / cw;
  p.n_taps = plan->n_taps;
  if (p.n_cb >= maxc) {
    // wide layers: a tile is up to 256 columns of ONE tap (one TMA box)
    const int parts = (p.n_cb + maxc - 1) / maxc;
    p.tpt = 1; p.cpt = (p.n_cb + parts - 1) / parts;
    p.tiles_per_tap = (p.n_cb + p.cpt - 1) / p.cpt;
    p.n_nt = p.n_taps * p.tiles_per_tap;
  } else {
    // narrow layers: a tile stacks whole taps (one TMA box each) up to 256 columns
    int tpt = maxc / p.n_cb;
    if (tpt > p.n_taps) tpt = p.n_taps;
    const int parts = (p.n_taps + tpt - 1) / tpt;
    p.tpt = (p.n_taps + parts - 1) / parts; p.cpt = p.n_cb;
    p.tiles_per_tap = 0;
    p.n_nt = (p.n_taps + p.tpt - 1) / p.tpt;
  }
  for (int t = 0; t < p.n_taps; ++t) {
    p.tap_plane[t] = plan->tap_plane[t]; p.tap_hoff[t] = plan->tap_hoff[t]; p.tap_woff[t] = plan->tap_woff[t];
  }
  p.Q = Q; p.Wo = g->Wo; p.kb_rows = kb_rows; p.kb_s = kb_s;
  p.nkb_slot = (Q >= kb_rows) ? Q / kb_rows : 1;
  p.group_mode = d->group_mode; p.n_groups = d->n_groups; p.slot_lo = d->slot_lo;
  p.n_seg = d->n_seg > 0 ? d->n_seg : 1;
  p.seg_stride = d->seg_stride;
  if (d->group_mode == CG_GROUP_SPLITK) {
    if (d->slot_hi <= d->slot_lo) return 0;
    if (kb_s > 1) {
      if (d->slot_lo % kb_s) return fail("slot_lo must be a multiple of %d for this layer", kb_s);
      p.u_lo = d->slot_lo / kb_s;
      p.u_hi = (d->slot_hi + kb_s - 1) / kb_s;
    } else {
      p.u_lo = static_cast<long long>(d->slot_lo) * p.nkb_slot;
      p.u_hi = static_cast<long long>(d->slot_hi) * p.nkb_slot;
    }
    const long long units = p.u_hi - p.u_lo;
    p.upg = (units + d->n_groups - 1) / d->n_groups;
    p.n_groups = static_cast<int>((units + p.upg - 1) / p.upg);
  }
  p.epi = d->epi; p.out = d->out; p.out_group_stride = d->out_group_stride;
  p.ldT = static_cast<long long>(p.n_taps) * p.C;
  p.KH = g->KH; p.KW = g->KW; p.Corig = g->C; p.merged = plan->merged;
  p.n_items = static_cast<long long>(p.n_groups) * p.n_nt * p.n_mtiles;

  if (d->half) {
    if (d->group_mode == CG_GROUP_SAMPLE) {
      if (!d->inv_x || !d->inv_y) return fail("cg_cl_contract: FP16 per-sample groups need inv_x / inv_y");
      if (p.n_seg > 1) return fail("cg_cl_contract: FP16 operands cannot sum passes per sample (staging scales differ)");
    }
    p.inv_x = d->inv_x; p.inv_y = d->inv_y; p.out_scale = d->out_scale;
  }
  if (d->pair) {
    if (!cg_cl_pair_ok(d->M, g, plan) || d->group_mode != CG_GROUP_SPLITK || d->epi != CG_EPI_ACCUM ||
        (kb_rows != 32 && kb_rows != 64))
      return fail("cg_cl_contract: pair = 1 needs a split-K clipped sum with M %% 256 == 0 and 128-channel multiples");
    return d->half ? launch_pair_t<true>(d, plan, p, kb_rows, kb_w, kb_h, kb_s, stream)
                   : launch_pair_t<false>(d, plan, p, kb_rows, kb_w, kb_h, kb_s, stream);
  }
  return d->half ? launch_cl_t<true>(d, plan, p, kb_rows, kb_w, kb_h, kb_s, dv.sm, stream)
                 : launch_cl_t<false>(d, plan, p, kb_rows, kb_w, kb_h, kb_s, dv.sm, stream);
}

int cg_rowpair_dot(const float* T, long long rows_total, int n_chunks, int row_a, int row_b, int B, float* out,
                   int accumulate, cg_stream_t stream) {
  if (B <= 0) return 0;
  cg::rowpair_dot_kernel<<<(B + 7) / 8, 256, 0, S(stream)>>>(T, rows_total, n_chunks, row_a, row_b, B, out, accumulate);
  CG_LAUNCH_CHECK();
  return 0;
}

int cg_joint_rows_sumsq(const float* rows_in, int R, int slot_lo, int seg_stride, int n_seg, int B, float* out,
                        cg_stream_t stream) {
  if (B <= 0) return 0;
  cg::joint_rows_sumsq_kernel<<<(B + 7) / 8, 256, 0, S(stream)>>>(rows_in, R, slot_lo, seg_stride, n_seg, B, out);
  CG_LAUNCH_CHECK();
  return 0;
}

int cg_outer_rows_cl(const float* Xt, long long x_rows, const float* Yt, long long y_rows, int M, int P, int slot0,
                     int B, float* out, cg_stream_t stream) {
  const long long total = static_cast<long long>(B) * M * P;
  if (total <= 0) return 0;
  DevInfo d;
  if (dev_info(&d)) return 1;
  cg::outer_rows_cl_kernel<float><<<ew_grid(total, 256, d.sm), 256, 0, S(stream)>>>(Xt, x_rows, Yt, y_rows, M, P, slot0,
                                                                                     B, nullptr, nullptr, out);
  CG_LAUNCH_CHECK();
  return 0;
}

int cg_outer_rows_cl_h(const void* Xt_half, long long x_rows, const void* Yt_half, long long y_rows, int M, int P,
                       int slot0, int B, const float* inv_x, const float* inv_y, float* out, cg_stream_t stream) {
  const long long total = static_cast<long long>(B) * M * P;
  if (total <= 0) return 0;
  if (!inv_x || !inv_y) return fail("null inverse scales");
  DevInfo d;
  if (dev_info(&d)) return 1;
  cg::outer_rows_cl_kernel<__half><<<ew_grid(total, 256, d.sm), 256, 0, S(stream)>>>(
      static_cast<const __half*>(Xt_half), x_rows, static_cast<const __half*>(Yt_half), y_rows, M, P, slot0, B, inv_x,
      inv_y, out);
  CG_LAUNCH_CHECK();
  return 0;
}

int cg_outer_rows(const float* X, long long x_pitch, const float* Y, long long y_pitch, int M, int P, int slot0,
                  int B, float* out, cg_stream_t stream) {
  const long long total = static_cast<long long>(B) * M * P;
  if (total <= 0) return 0;
  DevInfo d;
  if (dev_info(&d)) return 1;
  cg::outer_rows_kernel<<<ew_grid(total, 256, d.sm), 256, 0, S(stream)>>>(X, x_pitch, Y, y_pitch, M, P, slot0, B, out);
  CG_LAUNCH_CHECK();
  return 0;
}

int cg_row_sumsq(const float* src, long long rows, long long cols, long long ld, float* out, int accumulate,
                 cg_stream_t stream) {
  if (rows <= 0) return 0;
  DevInfo d;
  if (dev_info(&d)) return 1;
  if (cols <= 256) {
    long long g = (rows + 7) / 8;
    if (g > static_cast<long long>(d.sm) * 16) g = static_cast<long long>(d.sm) * 16;
    cg::row_sumsq_warp_kernel<<<static_cast<int>(g), 256, 0, S(stream)>>>(src, rows, cols, ld, out, accumulate, 0);
  } else {
    long long g = rows < static_cast<long long>(d.sm) * 8 ? rows : static_cast<long long>(d.sm) * 8;
    cg::row_sumsq_kernel<<<static_cast<int>(g), 256, 0, S(stream)>>>(src, rows, cols, ld, out, accumulate, 0);
  }
  CG_LAUNCH_CHECK();
  return 0;
}

int cg_vec_mul(const float* a, const float* b, float* out, long long n, cg_stream_t stream) {
  if (n <= 0) return 0;
  DevInfo d;
  if (dev_info(&d)) return 1;
  cg::vec_mul_kernel<<<ew_grid(n, 256, d.sm), 256, 0, S(stream)>>>(a, b, out, n);
  CG_LAUNCH_CHECK();
  return 0;
}

int cg_vec_fma(const float* a, const float* b, float w, float* out, long long n, cg_stream_t stream) {
  if (n <= 0) return 0;
  DevInfo d;
  if (dev_info(&d)) return 1;
  cg::vec_fma_kernel<<<ew_grid(n, 256, d.sm), 256, 0, S(stream)>>>(a, b, w, out, n);
  CG_LAUNCH_CHECK();
  return 0;
}

int cg_clip_factors(const float* norm2, int n_params, int n_slots, int per_layer, const float* C, float c_scale,
                    int clip_lo, int clip_hi, float* factors, float* norms_out, cg_stream_t stream) {
  if (n_slots <= 0 || n_params <= 0) return 0;
  if (!(c_scale > 0.f) || c_scale > 1.f) return fail("c_scale must be in (0, 1]");
  cg::clip_factors_kernel<<<(n_slots + 127) / 128, 128, 0, S(stream)>>>(norm2, n_params, n_slots, per_layer, C, c_scale,
                                                                       clip_lo, clip_hi, factors, norms_out);
  CG_LAUNCH_CHECK();
  return 0;
}

int cg_scale_slots(const float* src, float* dst, int rows, long long pitch, long long slot_stride, int slot_lo,
                   int slot_hi, const float* factor, cg_stream_t stream) {
  if (rows <= 0 || slot_hi <= slot_lo) return 0;
  DevInfo d;
  if (dev_info(&d)) return 1;
  if (slot_stride > 0x7fffffffLL) return fail("slot_stride too large");
  const long long cols = static_cast<long long>(slot_hi - slot_lo) * slot_stride;
  const int vec4 = (slot_stride % 4 == 0 && pitch % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(dst) & 15) == 0) ? 1 : 0;
  const long long work = vec4 ? cols / 4 : cols;
  long long gx = (work + 255) / 256;
  const long long want = (static_cast<long long>(d.sm) * 8 + rows - 1) / rows;   // ~8 blocks per SM overall
  if (gx > want) gx = want;
  if (gx < 1) gx = 1;
  dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(rows < 65535 ? rows : 65535));
  cg::scale_slots_kernel<<<grid, 256, 0, S(stream)>>>(src, dst, rows, pitch, static_cast<int>(slot_stride), slot_lo,
                                                      slot_hi, factor, vec4);
  CG_LAUNCH_CHECK();
  return 0;
}

int cg_permute_accum(const float* T, float* out, int M, int C, int KH, int KW, int accumulate, cg_stream_t stream) {
  const long long total = static_cast<long long>(M) * C * KH * KW;
  if (total <= 0) return 0;
  DevInfo d;
  if (dev_info(&d)) return 1;
  cg::permute_accum_kernel<<<ew_grid(total, 256, d.sm), 256, 0, S(stream)>>>(T, out, M, C, KH, KW, accumulate);
  CG_LAUNCH_CHECK();
  return 0;
}

int cg_weighted_colsum(const float* rows_in, const float* factor, int slot_lo, int slot_hi, int R, float* out,
                       int accumulate, cg_stream_t stream) {
  if (R <= 0) return 0;
  if (!accumulate) CG_CHECK(cudaMemsetAsync(out, 0, sizeof(float) * R, S(stream)));
  if (slot_hi <= slot_lo) return 0;
  int ny = (slot_hi - slot_lo + 31) / 32;             // >= 32 slots per block
  if (ny > 128) ny = 128;
  if (ny < 1) ny = 1;
  dim3 grid((R + 31) / 32, ny), block(32, 8);
  cg::weighted_colsum_kernel<<<grid, block, 0, S(stream)>>>(rows_in, factor, slot_lo, slot_hi, R, out);
  CG_LAUNCH_CHECK();
  return 0;
}

int cg_small_ops(const cg_small_op* ops, int n_ops, cg_stream_t stream) {
  if (n_ops <= 0) return 0;
  if (!ops) return fail("null operation table");
  if (n_ops > cg::kSmallMaxOps) return fail("at most %d operations per cg_small_ops call", cg::kSmallMaxOps);
  cg::SmallParams p;
  memset(&p, 0, sizeof(p));
  int blk = 0, m = 0;
  for (int i = 0; i < n_ops; ++i) {
    const cg_small_op& in = ops[i];
    if (in.n <= 0) continue;
    if (!in.a || !in.out) return fail("operation %d: null operand", i);
    cg::SmallOp& o = p.op[m];
    o.op = in.op; o.R = in.R; o.lo = in.lo; o.n = in.n; o.a = in.a; o.b = in.b; o.c = in.c; o.out = in.out; o.out2 = in.out2;
    o.blk0 = blk; o.nx = 1;
    long long nb = 1;
    switch (in.op) {
      case CG_OP_ROW_SUMSQ:
        if (in.R <= 0) return fail("operation %d: R must be positive", i);
        nb = (in.n + 7) / 8; if (nb > 256) nb = 256; break;
      case CG_OP_MUL:
        if (!in.b) return fail("operation %d: null operand", i);
        // fall through
      case CG_OP_COPY:
        nb = (in.n + 255) / 256; if (nb > 64) nb = 64; break;
      case CG_OP_WCOLSUM: {
        if (!in.b || in.R <= 0) return fail("operation %d: bad weighted column sum", i);
        if (in.lo + in.n > 0x7fffffffLL) return fail("operation %d: slot range too large", i);
        o.nx = (in.R + 31) / 32;
        long long ny = (in.n + 31) / 32; if (ny > 128) ny = 128; if (ny < 1) ny = 1;
        nb = static_cast<long long>(o.nx) * ny; break;
      }
      case CG_OP_CLIP_MULT:
        if (!in.b || !in.c || !in.out2) return fail("operation %d: null operand", i);
        if (in.n > 65536) return fail("operation %d: more than 65536 slots (use cg_clip_mult)", i);
        nb = 1; break;
      default:
        return fail("operation %d: unknown op %d", i, in.op);
    }
    o.nblk = static_cast<int>(nb);
    blk += o.nblk;
    ++m;
  }
  if (m == 0) return 0;
  p.n_ops = m;
  cg::small_ops_kernel<<<blk, 256, 0, S(stream)>>>(p);
  CG_LAUNCH_CHECK();
  return 0;
}

int cg_row_stat(const float* norms, int n_rows, int n_slots, int slot_lo, int slot_hi, int stat, float scalar,
                float* out, cg_stream_t stream) {
  if (n_rows <= 0) return 0;
  if (slot_hi <= slot_lo) return fail("empty slot range");
  cg::row_stat_kernel<<<n_rows, 256, 0, S(stream)>>>(norms, n_rows, n_slots, slot_lo, slot_hi, stat, scalar, out);
  CG_LAUNCH_CHECK();
  return 0;
}

namespace {
int noise_multi_impl(const cg_noise_seg* segs, int n_segs, double in_div, const float* in_div_dev,
                     double noise_div, const float* noise_div_dev, unsigned long long seed,
                     unsigned long long offset, const unsigned long long* offset_dev,
                     unsigned long long* offset_inc, const float* local_base, float* mc_base, float* const* peers,
                     long long buf_len, long long count_off, int rank, int world, cg_stream_t stream, int mc_mode = 3) {
  const bool ar = mc_base != nullptr || peers != nullptr;
  if (offset_inc) *offset_inc = 0;
  if (n_segs <= 0) return 0;
  if (!segs) return fail("null segment table");
  DevInfo d;
  if (dev_info(&d)) return 1;
  if (offset % 4) return fail("philox offset must be a multiple of 4");
  const unsigned long long tcap = static_cast<unsigned long long>(d.sm) * (d.max_thr / 256);
  unsigned long long off = 0;                    // advance inside this call (multiple of 4)
  int i = 0;
  while (i < n_segs) {
    cg::NoiseParams p;
    memset(&p, 0, sizeof(p));
    long long blk = 0;
    int m = 0;
    for (; i < n_segs && m < cg::kNoiseMaxSegs; ++i) {
      const cg_noise_seg& in = segs[i];
      if (in.n <= 0) continue;
      if (!in.grad) return fail("segment %d: null output", i);
      if (ar) {
        // both operands of every segment must lie inside the symmetric buffer
        const bool in_ok = !in.in || (in.in >= local_base && in.in + in.n <= local_base + buf_len);
        const bool out_ok = in.grad >= local_base && in.grad + in.n <= local_base + buf_len;
        if (!in_ok || !out_ok) return fail("segment %d lies outside the symmetric buffer", i);
      }
      cg::NoiseSeg& o = p.seg[m++];
      o.in = in.in; o.grad = in.grad; o.n = in.n; o.std_dev = in.std_dev;
      o.std_mult = static_cast<float>(in.std_mult);
      // upstream _generate_noise returns zeros when sigma*C == 0 and draws nothing from the generator
      o.draws = (in.std_dev || in.std_mult > 0.0) ? 1 : 0;
      unsigned long long tg = (static_cast<unsigned long long>(in.n) + 255) / 256;
      if (tg > tcap) tg = tcap;
      o.tgrid = static_cast<unsigned int>(tg);
      const unsigned long long trips = (static_cast<unsigned long long>(in.n) - 1) / (256 * tg * 4) + 1;
      o.off4 = off / 4;
      o.blk0 = blk;
      blk += static_cast<long long>(trips * tg);
      if (o.draws) off += trips * 4;
    }
    if (m == 0) break;
    p.n_segs = m; p.n_blocks = blk;
    p.in_mul = recip(in_div); p.noise_mul = recip(noise_div);
    p.in_div_dev = in_div_dev; p.noise_div_dev = noise_div_dev;
    p.seed = seed; p.offset = offset; p.offset_dev = offset_dev;
    p.local_base = local_base; p.mc_base = mc_base; p.count_off = count_off; p.rank = rank; p.world = world;
    if (peers) for (int r = 0; r < world && r < 8; ++r) p.peer[r] = peers[r];
    p.ld_mc = (mc_base && (mc_mode & 1)) ? 1 : 0; p.st_mc = (mc_base && (mc_mode & 2)) ? 1 : 0;
    long long grid = ar ? (blk + world - 1) / world : blk;
    const long long cap = static_cast<long long>(d.sm) * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    if (ar) cg::noise_multi_kernel<true><<<static_cast<unsigned int>(grid), 256, 0, S(stream)>>>(p);
    else cg::noise_multi_kernel<false><<<static_cast<unsigned int>(grid), 256, 0, S(stream)>>>(p);
    CG_LAUNCH_CHECK();
  }
  if (offset_inc) *offset_inc = off;
  return 0;
}
}  // namespace

int cg_noise_finalize_multi(const cg_noise_seg* segs, int n_segs, double in_div, const float* in_div_dev,
                            double noise_div, const float* noise_div_dev, unsigned long long seed,
                            unsigned long long offset, const unsigned long long* offset_dev,
                            unsigned long long* offset_inc, cg_stream_t stream) {
  return noise_multi_impl(segs, n_segs, in_div, in_div_dev, noise_div, noise_div_dev, seed, offset, offset_dev, offset_inc,
                          nullptr, nullptr, nullptr, 0, -1, 0, 1, stream);
}

int cg_noise_finalize_allreduce(const cg_noise_seg* segs, int n_segs, int mean, unsigned long long seed,
                                unsigned long long offset, const unsigned long long* offset_dev,
                                unsigned long long* offset_inc, const float* local_base, float* mc_base,
                                float* const* peers, long long buf_len, long long count_off, int rank, int world,
                                cg_stream_t stream) {
  if (!local_base || (!mc_base && !peers)) return fail("null symmetric base / neither a multicast nor peer mappings");
  if (world < 1 || rank < 0 || rank >= world) return fail("bad rank %d of %d", rank, world);
  if (!mc_base && world > 8) return fail("the peer-to-peer exchange covers up to 8 ranks");
  if (peers && peers[rank] != local_base) return fail("peers[rank] must be this rank's own buffer");
  // both mappings given: loads through the switch (one reduced reply per request instead of `world` round trips),
  // stores through the peer mappings -- or the other way round (CSLGAN_XFER_MIX=1: multicast loads, 2: multicast stores)
  const char* mix_env = getenv("CSLGAN_XFER_MIX");                     // (read per call: a probe switches it at run time)
  const int mix = mix_env ? atoi(mix_env) : 0;
  const int mc_mode = (mc_base && peers && (mix == 1 || mix == 2)) ? mix : 3;
  if (mc_base && peers && world > 8) return fail("the peer-to-peer exchange covers up to 8 ranks");
  if (count_off >= buf_len) return fail("count element outside the buffer");
  if (mean && count_off < 0) return fail("mean reduction needs the sample-count element");
  // mean: both divisors are the all-rank sample count the kernel reads through the switch (1.0 only marks them active)
  return noise_multi_impl(segs, n_segs, mean ? 1.0 : 0.0, nullptr, mean ? 1.0 : 0.0, nullptr, seed, offset, offset_dev,
                          offset_inc, local_base, mc_base, (mc_base && mc_mode == 3) ? nullptr : peers, buf_len,
                          mean ? count_off : -1, rank, world, stream, mc_mode);
}

int cg_noise_finalize(const float* in, float* grad, long long n, double in_div, double std, double noise_div,
                      unsigned long long seed, unsigned long long offset, unsigned long long* offset_inc,
                      cg_stream_t stream) {
  cg_noise_seg s = {in, grad, n, std, nullptr};
  return cg_noise_finalize_multi(&s, 1, in_div, nullptr, noise_div, nullptr, seed, offset, nullptr, offset_inc, stream);
}

int cg_noise_finalize_graph(const float* in, float* grad, long long n, double in_div, double std_mult,
                            const float* std_dev, double noise_div, unsigned long long seed,
                            const unsigned long long* offset_dev, unsigned long long intra_offset,
                            unsigned long long* offset_inc, cg_stream_t stream) {
  if (!offset_dev) return fail("offset_dev must not be null");
  cg_noise_seg s = {in, grad, n, std_mult, std_dev};
  return cg_noise_finalize_multi(&s, 1, in_div, nullptr, noise_div, nullptr, seed, intra_offset, offset_dev,
                                 offset_inc, stream);
}

int cg_philox_advance(unsigned long long* offset_dev, unsigned long long inc, cg_stream_t stream) {
  if (!offset_dev) return fail("offset_dev must not be null");
  cg::philox_advance_kernel<<<1, 32, 0, S(stream)>>>(offset_dev, inc);
  CG_LAUNCH_CHECK();
  return 0;
}

int cg_noise_finalize_dev(const float* in, float* grad, long long n, double in_div, double std_mult,
                          const float* std_dev, double noise_div, unsigned long long seed, unsigned long long offset,
                          unsigned long long* offset_inc, cg_stream_t stream) {
  if (!std_dev) return fail("std_dev must not be null");
  cg_noise_seg s = {in, grad, n, std_mult, std_dev};
  return cg_noise_finalize_multi(&s, 1, in_div, nullptr, noise_div, nullptr, seed, offset, nullptr, offset_inc, stream);
}

int cg_row_l2_norm(const float* src, long long rows, long long cols, float* norms, cg_stream_t stream) {
  if (rows <= 0) return 0;
  DevInfo d;
  if (dev_info(&d)) return 1;
  if (rows < 2LL * d.sm && rows * cols >= (1LL << 18) && rows <= 65535) {
    // few long rows: split every row over enough blocks to fill the machine (~4 blocks per SM in total)
    long long nsplit = (4LL * d.sm + rows - 1) / rows;
    long long per = ((cols + nsplit - 1) / nsplit + 1023) / 1024 * 1024;
    nsplit = (cols + per - 1) / per;
    CG_CHECK(cudaMemsetAsync(norms, 0, sizeof(float) * rows, S(stream)));
    dim3 grid(static_cast<unsigned>(nsplit), static_cast<unsigned>(rows));
    cg::row_sumsq_split_kernel<<<grid, 256, 0, S(stream)>>>(src, cols, cols, per, norms);
    CG_LAUNCH_CHECK();
    cg::sqrt_inplace_kernel<<<static_cast<unsigned>((rows + 255) / 256), 256, 0, S(stream)>>>(norms, rows);
    CG_LAUNCH_CHECK();
    return 0;
  }
  if (cols <= 256) {
    long long g = (rows + 7) / 8;
    if (g > static_cast<long long>(d.sm) * 16) g = static_cast<long long>(d.sm) * 16;
    cg::row_sumsq_warp_kernel<<<static_cast<int>(g), 256, 0, S(stream)>>>(src, rows, cols, cols, norms, 0, 1);
  } else {
    long long g = rows < static_cast<long long>(d.sm) * 8 ? rows : static_cast<long long>(d.sm) * 8;
    cg::row_sumsq_kernel<<<static_cast<int>(g), 256, 0, S(stream)>>>(src, rows, cols, cols, norms, 0, 1);
  }
  CG_LAUNCH_CHECK();
  return 0;
}

int cg_row_l2_norm_bwd(const float* g, const float* norms, const float* gout, long long rows, long long cols,
                       float* gin, cg_stream_t stream) {
  if (rows <= 0) return 0;
  DevInfo d;
  if (dev_info(&d)) return 1;
  // ~8 blocks per SM in total: row lanes on grid.y, column chunks of >= 1024 elements on grid.x
  long long gy = rows < static_cast<long long>(d.sm) * 8 ? rows : static_cast<long long>(d.sm) * 8;
  long long gx = (static_cast<long long>(d.sm) * 8 + gy - 1) / gy;
  const long long max_gx = (cols + 1023) / 1024;
  if (gx > max_gx) gx = max_gx;
  if (gx < 1) gx = 1;
  if (gy > 65535) gy = 65535;
  dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(gy));
  cg::row_l2_norm_bwd_kernel<<<grid, 256, 0, S(stream)>>>(g, norms, gout, rows, cols, gin);
  CG_LAUNCH_CHECK();
  return 0;
}

int cg_vec_max(const float* v, long long n, float* out, cg_stream_t stream) {
  if (n <= 0) return fail("cg_vec_max on an empty vector");
  cg::vec_max_kernel<<<1, 1024, 0, S(stream)>>>(v, n, out);
  CG_LAUNCH_CHECK();
  return 0;
}

int cg_l2_clip(const float* t, long long rows, long long cols, float C, float* out, float* norms_out,
               cg_stream_t stream) {
  if (rows <= 0) return 0;
  DevInfo d;
  if (dev_info(&d)) return 1;
  long long grid = rows < static_cast<long long>(d.sm) * 8 ? rows : static_cast<long long>(d.sm) * 8;
  cg::l2_clip_kernel<<<static_cast<int>(grid), 256, 0, S(stream)>>>(t, rows, cols, C, out, norms_out);
  CG_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
