/*
 * cslgan_b200.h -- C ABI of the B200-native DP discriminator-update hot path.
 *
 * The reference (twosixlabs/csl-gan) has no native boundary: train.py talks to a Python
 * object protocol (the twosixlabs/opacus fork's PrivacyEngine / ISPrivacyEngine,
 * reference requirements.txt:9, train.py:13-14).  This header is the native seam a
 * maintainer binds *underneath* that protocol: each entry point replaces one piece of
 * tensor arithmetic the fork performs with ATen ops, and the comment on each one cites
 * the reference call site (reference file:line) and the upstream-opacus op it replaces.
 * INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - plain C: raw device pointers (fp32 unless noted), sizes, a cudaStream_t passed as
 *     void*.  No torch types.  All work is enqueued on `stream`; nothing synchronises.
 *   - every function returns 0 on success, non-zero on failure; cg_last_error() returns a
 *     thread-local message.  Nothing throws across the boundary.
 *   - there is NO CPU fallback: without a CUDA device / the sm_100a cubin every call fails.
 *
 * Vocabulary
 *   slot      one (pass, sample) pair: slot = pass * B + n.  Passes are D forward calls in
 *             forward order (0 = fake, 1 = real in train.py:382-383).
 *   X operand "plain" side of a layer's per-sample contraction, staged K-major:
 *             X[r][slot * x_slot_stride + q]   (Conv2d/Linear: backprops, r = out channel)
 *   Y operand "unfolded" side, staged as kw-planes so that every filter tap is a shifted
 *             window of a plain 2-D matrix (no im2col blow-up beyond KW * n_rho / (sh*sw)):
 *             Y[(j*KW + kw)*C + c][slot * y_slot_stride + hs*Wop + ow]
 *   per-sample gradient of a layer  G_slot[m][c][kh][kw] = sum_q X[m][slot,q] * Y_tap[c][slot,q]
 */
#ifndef CSLGAN_B200_H
#define CSLGAN_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define CG_MAX_KH 16

typedef void* cg_stream_t; /* cudaStream_t */

int cg_version(void);
const char* cg_last_error(void);
/* SM count / max threads per SM of the current device (used for the torch-compatible Philox grid). */
int cg_device_info(int* sm_count, int* max_threads_per_sm, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------------------------------------
 * Geometry of the unfolded operand of one layer (Conv2d: the layer input; ConvTranspose2d: the
 * grad wrt the layer output; Linear: KH=KW=1, H=W=Ho=Wo=1, C = in_features).
 * Replaces F.unfold(A, kernel, padding, stride, dilation) in upstream opacus
 * `_compute_conv_grad_sample` (fired from reference train.py:387 d_loss.backward()).
 * ------------------------------------------------------------------------------------------- */
typedef struct cg_unfold_geom {
  int C, H, W;          /* source tensor [B][C][H][W]                                  */
  int KH, KW;           /* filter taps                                                 */
  int sh, sw, ph, pw, dh, dw; /* stride / padding / dilation                           */
  int Ho, Wo;           /* number of window positions (contraction length Q = Ho*Wo)   */
} cg_unfold_geom;

/* Derived staging layout, filled by cg_plan_unfold(). */
typedef struct cg_unfold_plan {
  int n_rho;                 /* distinct row residues (kh*dh - ph) mod sh               */
  int Hs;                    /* staged rows per plane = Ho + a_max - a_min              */
  int Wop;                   /* Wo rounded up to 4: TMA box starts must be 16-byte aligned */
  int rows;                  /* staged matrix rows = n_rho * KW * C                     */
  int slot_stride;           /* columns per slot = Hs * Wop                             */
  int tap_row0[CG_MAX_KH];   /* first staged row of tap-group kh  (= j(kh) * KW * C)    */
  int tap_coloff[CG_MAX_KH]; /* column offset of tap-group kh     (= (a(kh)-a_min)*Wop) */
  int rho[CG_MAX_KH];        /* residue value of plane j                                */
  int a_min;
} cg_unfold_plan;

int cg_plan_unfold(const cg_unfold_geom* g, cg_unfold_plan* plan);

/* ---------------------------------------------------------------------------------------------
 * Capture (replaces the fork's forward/backward hook bodies: upstream _capture_activations and
 * _compute_{linear,conv}_grad_sample; hooks are armed by enable_hooks(), reference train.py:373).
 * All staged values are rounded to TF32 (round-to-nearest) so the tensor cores see unbiased
 * operands.
 * ------------------------------------------------------------------------------------------- */

/* src [B][R] (Linear activations or backprops) -> dst[r][slot0+n] = tf32(scale*src[n][r]).
 * Optional: copy_out[(slot0+n)*R + r] = scale*src (fp32, un-rounded: per-sample bias gradients),
 *           sumsq[slot0+n] = sum_r (scale*src[n][r])^2 (closed-form Linear norms ||a||^2, ||b||^2). */
int cg_stage_rows_t(const float* src, int B, int R, float scale, float* dst, long long dst_pitch,
                    int slot0, float* copy_out, float* sumsq, cg_stream_t stream);

/* src [B][R][Q], Q = Ho*Wo (conv backprops; ConvTranspose2d activations) ->
 *   dst[r][(slot0+n)*Qpad + oh*Wop + ow] = tf32(scale*src[n][r][oh*Wo + ow]), zero elsewhere
 *   (Wop >= Wo is the padded window-row pitch of the matching unfolded operand; Wo = Wop = Q
 *   stages a flat row).  Optional rowsum[(slot0+n)*R + r] = scale * sum_q src[n][r][q]
 *   (per-sample bias gradients, upstream torch.sum(B, dim=2)). */
int cg_stage_rows(const float* src, int B, int R, int Q, int Wo, int Wop, int Qpad, float scale,
                  float* dst, long long dst_pitch, int slot0, float* rowsum, cg_stream_t stream);

/* src [B][C][H][W] -> kw-plane matrix (see cg_unfold_plan), zero padding materialised. */
int cg_stage_unfold(const float* src, int B, const cg_unfold_geom* g, const cg_unfold_plan* plan,
                    float scale, float* dst, long long dst_pitch, int slot0, cg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * The per-sample contraction on tcgen05 tensor cores (TMA-fed, TMEM accumulators).
 * One persistent kernel, three epilogues:
 *   CG_EPI_SUMSQ  norm2[group] += sum of squares of the group's gradient tile
 *                 -> per-sample squared norms without materialising B x |theta|
 *                 (replaces calc_sample_norms, reference train.py:311-314, and the norm pass
 *                  inside privacy_engine.clip(), train.py:399)
 *   CG_EPI_ACCUM  out[m][kh][kw*C+c] += tile  (clipped weighted sum as ONE split-K GEMM over
 *                 all slots with clip-factor-scaled X; replaces upstream _weighted_sum
 *                 einsum("i,i...", cf, grad_sample) inside clip(), train.py:399)
 *   CG_EPI_STORE  out[group][m][c][kh][kw] = tile (materialise grad_sample for the rare
 *                 consumers: train.py:233, 447 and tests)
 * ------------------------------------------------------------------------------------------- */
enum { CG_EPI_SUMSQ = 0, CG_EPI_ACCUM = 1, CG_EPI_STORE = 2,
       /* cg_cl_contract only: out[group][m][tap][c'] = tile in the gradient-natural layout, coalesced.  For THIN
          layers (few parameters, huge operands: the 3-channel first conv) materialising the per-sample gradients
          (|theta_layer| floats per sample) is far cheaper than contracting twice: norms and the clipped sum are
          then a row_sumsq and a weighted column sum over that small tensor. */
       CG_EPI_STORE_NATURAL = 3 };
enum { CG_GROUP_SAMPLE = 0, CG_GROUP_SPLITK = 1 };

typedef struct cg_contract_desc {
  const float* X; long long x_pitch; int x_rows; long long x_cols; /* [x_rows][x_cols], pitch in floats */
  const float* Y; long long y_pitch; int y_rows; long long y_cols;
  int M;                       /* valid rows of X (= x_rows)                               */
  int C, KH, KW;               /* output index mapping; KH*KW*C gradient columns per row   */
  int tap_row0[CG_MAX_KH];
  int tap_coloff[CG_MAX_KH];
  int nkb;                     /* 32-wide k-blocks per slot segment                        */
  long long x_slot_stride;     /* columns per slot in X                                    */
  long long y_slot_stride;     /* columns per slot in Y                                    */
  int group_mode;              /* CG_GROUP_SAMPLE: group g covers slots slot_lo+g + s*seg_stride, s<n_seg
                                  CG_GROUP_SPLITK: group g covers slots [slot_lo+g*spg, min(.., slot_hi)) */
  int n_groups;
  int slot_lo, slot_hi, spg;
  int n_seg, seg_stride;
  int epi;
  float* out;
  long long out_group_stride;  /* CG_EPI_STORE: floats between groups                      */
  int block_n;                 /* 0 = choose automatically; else multiple of 16, <= 256    */
  int max_ctas;                /* 0 = one CTA per SM                                       */
} cg_contract_desc;

int cg_contract(const cg_contract_desc* d, cg_stream_t stream);

/* Materialise Linear per-sample weight gradients (K = 1 outer products; upstream
 * einsum("n...i,n...j->nij", B, A)): out[n][m][p] = X[m][slot0+n] * Y[p][slot0+n]. */
int cg_outer_rows(const float* X, long long x_pitch, const float* Y, long long y_pitch, int M, int P,
                  int slot0, int B, float* out, cg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Ghost norms for layers with few window positions (Q = Ho*Wo divides 128):
 *   ||G_n||^2 = sum_{q,q'} (Xn^T Xn)[q,q'] (Un^T Un)[q,q']
 * Both Gram matrices run on tcgen05 for 128/Q samples per tile from the channels-last staging of the
 * main path (cg_stage_xt / cg_stage_yt, un-merged plan), read here as K-major tiles:
 *   Xt[o/32][slot*Q + q][o%32]
 *   Yt[plane*n_cb + c/32][slot][hs][ws][c%32]   (space-to-depth so every filter tap is a unit-stride
 *                                                window = one 5-D TMA box)
 * Same result as cg_contract(CG_EPI_SUMSQ) (replaces calc_sample_norms, reference train.py:311-314),
 * without the O x P-element epilogue per sample.
 * ------------------------------------------------------------------------------------------- */
typedef struct cg_ghost_plan {
  int n_rh, n_rw;            /* distinct row / column residues                               */
  int Hs, Ws;                /* staged plane extent                                          */
  int ah_min, aw_min;
  int Cp;                    /* staged channels rounded up to 32 (whole 128-byte chunks)     */
  int rho_h[CG_MAX_KH], rho_w[CG_MAX_KH];
  int tap_plane[CG_MAX_KH * CG_MAX_KH], tap_hoff[CG_MAX_KH * CG_MAX_KH], tap_woff[CG_MAX_KH * CG_MAX_KH];
  long long slot_stride;     /* floats per slot inside one 32-channel chunk = Hs*Ws*32      */
  int merged;                /* 1: filter columns folded into the channel axis (c' = kw*C + c) for thin
                                inputs such as the 3-channel image; taps then run over kh only        */
  int Cs;                    /* staged channels per tap: C, or KW*C when merged                       */
  int n_taps;                /* KH*KW, or KH when merged                                              */
  int cw;                    /* channels per 128-byte chunk row: 32 (TF32 words) or 64 (FP16); Cp and
                                slot_stride are in these units                                        */
} cg_ghost_plan;
typedef cg_ghost_plan cg_cl_plan;

/* returns non-zero (with a message) when the geometry is outside the ghost path's envelope */
int cg_plan_ghost(const cg_unfold_geom* g, cg_ghost_plan* plan);

typedef struct cg_ghost_desc {
  const float* Xt; long long xt_pitch; long long xt_rows;   /* xt_rows = n_slots_total*Q (xt_pitch unused) */
  const float* Yt; int n_slots_total;                       /* plane-major staged tensor   */
  int O;
  int slot0, n_slots;                                       /* slots to process            */
  float* norm2;                                             /* norm2[slot - slot0] +=      */
  int max_ctas;
  int half;                                                 /* 1: Xt / Yt hold FP16 (cg_stage_*_h), see below */
  const float* inv_x; const float* inv_y;                   /* half: per-slot inverse staging scales [n_slots_total] */
} cg_ghost_desc;

/* Two kernels behind one entry point.  When a stride-residue plane has at most 128 positions (Hs*Ws <= 128: the
 * 8x8 and 4x4 layers of the CelebA critics) the activation Gram is built ONCE PER PLANE,
 *     UU[q,q'] = sum_taps P_plane(tap)[a_tap(q), a_tap(q')],   P_pl = Y_pl Y_pl^T,
 * and the taps are a gather-sum in the epilogue (csrc/ghost2.cuh); otherwise one Gram k-block per tap and chunk
 * (csrc/ghost.cuh).  Environment CSLGAN_GHOST_V1=1 forces the latter (A/B measurements). */
int cg_ghost_norm(const cg_ghost_desc* d, const cg_unfold_geom* g, const cg_ghost_plan* plan, cg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Channels-last contraction (main path).  Same three epilogues and the same operations replaced as
 * cg_contract (calc_sample_norms / _weighted_sum / materialised grad_sample, reference
 * train.py:311-314, 399, 233), but both operands are staged channels-innermost with no unfold:
 *   Xt[m/32][slot*Q + q][m%32]                    cg_stage_xt (element-wise from a channels_last tensor)
 *   Yt[plane*n_cb + c'/32][slot][hs][ws][c'%32]   cg_stage_yt (space-to-depth; thin inputs fold kw into c')
 * (32-channel chunks outermost, so one TMA box fetches every chunk of a tile)
 * and fed to tcgen05 as MN-major operands: a k-block is 32 window positions (Q >= 32), 32/Q samples
 * (Q < 32, split-K mode) or the Q positions of one sample (8 | Q < 32, per-sample mode).
 * Linear layers are Q = 1.  CG_EPI_ACCUM writes the gradient-natural layout out[m][tap][c'] (which
 * is the memory order of a channels_last weight).
 * ------------------------------------------------------------------------------------------- */
int cg_plan_cl(const cg_unfold_geom* g, int merged, cg_cl_plan* plan);              /* TF32: 32-channel chunks */
int cg_plan_cl_cw(const cg_unfold_geom* g, int merged, int cw, cg_cl_plan* plan);   /* cw = 32 (TF32) / 64 (FP16) */

/* src addressed as src[n*sn + m*sm + oh*sh + ow*sw] (any layout; fastest when sm == 1).
 * Xt[m/32][(slot0+n)*Q + q][m%32] = tf32(scale*src), rows_total = rows of each chunk (n_slots_total*Q);
 * optional bias_rows[(slot0+n)*M + m] = scale*sum_q src (per-sample bias gradients) and
 * sumsq[slot0+n] = sum (scale*src)^2 (closed-form Linear norms). */
int cg_stage_xt(const float* src, long long sn, long long sm, long long sh, long long sw, int B, int M,
                int Ho, int Wo, float scale, float* dst, long long rows_total, int slot0, float* bias_rows,
                float* sumsq, cg_stream_t stream);

/* src addressed as src[n*sn + c*sc + h*sh + w*sw]; Yt as described above (zero padded). */
int cg_stage_yt(const float* src, long long sn, long long sc, long long sh, long long sw, int B,
                const cg_unfold_geom* g, const cg_cl_plan* plan, float scale, float* dst, int n_slots_total,
                int slot0, cg_stream_t stream);

/* FP16 operand containers (the default of the engine's channels-last path).  Same layouts with 2-byte elements
 * (64-byte chunk rows).  FP16 carries TF32's 10 explicit mantissa bits; its narrow exponent is handled by an EXACT
 * power-of-two scale per (slot, operand): staged = fp16(scale*src * 2^e), e chosen so that the sample's largest
 * magnitude lands in [2^13, 2^14); inv[slot0+n] receives 2^-e.  Everything downstream undoes it exactly:
 *   per-sample results (norms, stored gradients) are multiplied by inv_x[slot]*inv_y[slot] in the epilogue;
 *   the clipped sum folds factor[slot]*inv_x[slot]*inv_y[slot] / 2^E into the scaled operand (cg_clip_mult,
 *   cg_scale_slots_h) and multiplies the accumulated tile by 2^E (cg_cl_desc.out_scale).
 * amax: scratch [n_slots_total] words (zeroed and filled with the per-sample maxima by these calls). */
int cg_stage_xt_h(const float* src, long long sn, long long sm, long long sh, long long sw, int B, int M,
                  int Ho, int Wo, float scale, void* dst_half, long long rows_total, int slot0, float* bias_rows,
                  float* sumsq, unsigned int* amax, float* inv, cg_stream_t stream);
int cg_stage_yt_h(const float* src, long long sn, long long sc, long long sh, long long sw, int B,
                  const cg_unfold_geom* g, const cg_cl_plan* plan, float scale, void* dst_half, int n_slots_total,
                  int slot0, unsigned int* amax, float* inv, cg_stream_t stream);
/* Batched small operations: ONE launch for a table of the per-layer scalar / bias operations of a step (each is a
 * 3-5 us launch on its own; a CelebA step has ~25 of them).  Same arithmetic as the single-operation entry points:
 *   CG_OP_ROW_SUMSQ  out[r] = sum_j a[r*R + j]^2, r < n                      (cg_row_sumsq; bias-gradient norms)
 *   CG_OP_COPY       out[j] = a[j], j < n
 *   CG_OP_MUL        out[j] = a[j] * b[j], j < n                            (cg_vec_mul; Linear closed-form norms)
 *   CG_OP_WCOLSUM    out[r] += sum_{s in [lo, lo+n)} b[s] * a[s*R + r]      (cg_weighted_colsum, out zeroed by the caller)
 *   CG_OP_CLIP_MULT  out[s] = a[s]*b[s]*c[s] / 2^E, out2[0] = 2^E, s in [lo, lo+n), n <= 65536   (cg_clip_mult)
 * At most 32 operations per call. */
enum { CG_OP_ROW_SUMSQ = 0, CG_OP_COPY = 1, CG_OP_MUL = 2, CG_OP_WCOLSUM = 3, CG_OP_CLIP_MULT = 4 };
typedef struct {
  int op;
  int R;
  int lo;
  long long n;
  const float* a;
  const float* b;
  const float* c;
  float* out;
  float* out2;
} cg_small_op;
int cg_small_ops(const cg_small_op* ops, int n_ops, cg_stream_t stream);

/* Thin first convolution (few input channels, large window grid: the 3 -> 64 channel conv of the CelebA critics):
 * per-sample weight gradients, their squared norms and the per-sample bias gradients in ONE kernel that reads the
 * layer's own tensors -- no staged operands (csrc/thin.cuh).  Replaces, for such a layer, the fork's
 * _capture_activations / _capture_backprops + _compute_conv_grad_sample + calc_sample_norms (reference
 * train.py:382-387, 311-314).
 *   act   image batch [B][C][H][W] through strides (elements; any layout)
 *   bp    the layer's grad_output, dense channels-last [B][Ho*Wo][M] fp32, 16-byte aligned
 *   Gs    [B][M][KH*KW*C] (sample stride gs_stride floats): scale * gradient in the gradient-natural layout
 *         out[m][kh][kw][c] (= the memory of a channels_last conv weight)
 *   norm2 [B] ||Gs[n]||^2        bias_rows [B][M] scale * sum over positions of bp (may be NULL)
 * TF32 operands (rounded to nearest in shared memory), fp32 accumulation.  cg_thin_direct_ok: 1 when the geometry is
 * covered (KH*KW*C <= 128, M in {32, 64, 128}, Ho*Wo % 64 == 0, Wo % 4 == 0, image + tiles fit in shared memory). */
int cg_thin_direct_ok(const cg_unfold_geom* g, int M);
int cg_thin_capture(const float* act, long long a_sn, long long a_sc, long long a_sh, long long a_sw, const float* bp,
                    int B, const cg_unfold_geom* g, int M, float scale, float* Gs, long long gs_stride, float* norm2,
                    float* bias_rows, cg_stream_t stream);
/* The same for TWO batches in one launch (the fake and the real pass of a D step: 2B items share the 148 persistent
 * CTAs, which wastes less of the last round than B items twice).  Both image batches have the same strides; the second
 * segment's outputs are Gs2 / norm2_2 / bias_rows2.  B2 = 0 is cg_thin_capture. */
int cg_thin_capture2(const float* act, const float* act2, long long a_sn, long long a_sc, long long a_sh, long long a_sw,
                     const float* bp, const float* bp2, int B, int B2, const cg_unfold_geom* g, int M, float scale, float* Gs,
                     float* Gs2, long long gs_stride, float* norm2, float* norm2_2, float* bias_rows, float* bias_rows2,
                     cg_stream_t stream);

/* mult[s] = factor[s] * inv_x[s] * inv_y[s] / 2^E for s in [slot_lo, slot_hi), out_scale[0] = 2^E with 2^E the
 * smallest power of two above max_s factor*inv_x*inv_y (so mult <= 1 and the scaled operand stays in FP16 range;
 * samples far below the largest contribution lose low bits they could not contribute to the sum anyway).
 * out_scale points at TWO floats: [0] receives 2^E, [1] is scratch (the raw maximum of the multi-block path). */
int cg_clip_mult(const float* factor, const float* inv_x, const float* inv_y, int slot_lo, int slot_hi,
                 float* mult, float* out_scale, cg_stream_t stream);
/* dst[r][slot*stride + q] = fp16(src * mult[slot]) over FP16 rows (the factor-scaled operand of the clipped sum) */
int cg_scale_slots_h(const void* src_half, void* dst_half, int rows, long long pitch, long long slot_stride,
                     int slot_lo, int slot_hi, const float* mult, cg_stream_t stream);
/* cg_scale_slots_h for up to 8 layers in one launch (the factor-scaled operands of all layers of a step) */
typedef struct {
  const void* src;
  void* dst;
  const float* mult;
  long long pitch;
  long long slot_stride;
  int rows;
  int slot_lo, slot_hi;
} cg_scale_seg;
int cg_scale_slots_h_multi(const cg_scale_seg* segs, int n_segs, cg_stream_t stream);
/* cg_outer_rows_cl over FP16 operands: out = Xt*Yt * inv_x[slot]*inv_y[slot] */
int cg_outer_rows_cl_h(const void* Xt_half, long long x_rows, const void* Yt_half, long long y_rows, int M, int P,
                       int slot0, int B, const float* inv_x, const float* inv_y, float* out, cg_stream_t stream);

typedef struct cg_cl_desc {
  const float* Xt; long long xt_pitch; long long xt_rows; int M;   /* xt_rows = n_slots_total*Q (xt_pitch unused) */
  const float* Yt; int n_slots_total;
  int group_mode;              /* CG_GROUP_SAMPLE: one group per slot in [slot_lo, slot_lo+n_groups)
                                  CG_GROUP_SPLITK: slots [slot_lo, slot_hi) cut into n_groups K ranges */
  int n_groups;
  int slot_lo, slot_hi;
  int epi;
  float* out;
  long long out_group_stride;
  int max_ctas;
  int n_seg, seg_stride;       /* CG_GROUP_SAMPLE: a group also covers slots slot_lo+g + s*seg_stride, s < n_seg
                                  (per-sample sum over passes, accum_passes=True); 0/1 = single slot       */
  int pair;                    /* 1: run on CTA pairs (cluster of 2, tcgen05 cta_group::2, 256x256 tiles, each CTA
                                  loads half of the unfolded operand).  Only CG_GROUP_SPLITK + CG_EPI_ACCUM with
                                  M % 256 == 0 and 128-channel multiples (cg_cl_pair_ok); an error otherwise */
  int half;                    /* 1: Xt / Yt hold FP16 staged by cg_stage_xt_h / cg_stage_yt_h (tcgen05 kind::f16) */
  const float* inv_x;          /* half, CG_GROUP_SAMPLE: per-slot inverse staging scales [n_slots_total]; the     */
  const float* inv_y;          /*   epilogues multiply every per-sample result by inv_x[slot] * inv_y[slot]        */
  const float* out_scale;      /* half, CG_GROUP_SPLITK: device scalar multiplied into the accumulated tile
                                  (cg_clip_mult) */
} cg_cl_desc;

/* contraction rows per k-block (32, or 64 for FP16 operands where the window grid allows) and slots per k-block
 * (1 unless Ho*Wo < kb_rows) of the split-K clipped sum: what the host needs to pick a split-K group count */
int cg_cl_kblock_rows(const cg_unfold_geom* g, int half, int* kb_rows, int* kb_s);

/* 1 when cg_cl_contract accepts d->pair = 1 for this layer (split-K clipped sum), else 0 */
int cg_cl_pair_ok(int M, const cg_unfold_geom* g, const cg_cl_plan* plan);

/* Kernel selection inside cg_cl_contract (csrc/abi.cu): FP16 layers whose window grid is 16 / 32 / 64 wide with
 * C <= 64, M <= 128 and Ho*Wo*256 B <= 64 KB (the 16x16 layer of the CelebA critics) run csrc/cl_res.cuh -- the
 * sample's backprops resident in shared memory, the filter taps read as shifted windows of ONE shared-memory copy of
 * the stride-residue plane -- for CG_GROUP_SAMPLE + CG_EPI_SUMSQ and for CG_GROUP_SPLITK + CG_EPI_ACCUM; d->pair = 1
 * selects the CTA-pair kernel (csrc/cl_pair.cuh); everything else the tap-per-box kernel (csrc/cl.cuh).
 * CSLGAN_RESIDENT=0 switches the resident kernels off (A/B measurements). */
int cg_cl_contract(const cg_cl_desc* d, const cg_unfold_geom* g, const cg_cl_plan* plan, cg_stream_t stream);

/* Joint clipping (accum_passes=True: the per-sample gradients of all passes are summed before clipping).
 * Cross terms of the Linear closed form ||sum_p b_p a_p^T||^2 = sum_{p,p'} (a_p.a_p')(b_p.b_p'):
 *   out[n] (+)= < T[:, row_a+n, :], T[:, row_b+n, :] >   over a chunk-major matrix T[n_chunks][rows_total][32] */
int cg_rowpair_dot(const float* T, long long rows_total, int n_chunks, int row_a, int row_b, int B, float* out,
                   int accumulate, cg_stream_t stream);
/* out[n] = || sum_{s<n_seg} rows_in[slot_lo + n + s*seg_stride][0:R] ||^2   (joint per-sample bias norms) */
int cg_joint_rows_sumsq(const float* rows_in, int R, int slot_lo, int seg_stride, int n_seg, int B, float* out,
                        cg_stream_t stream);

/* out[n][m][p] = Xt[m/32][slot0+n][m%32] * Yt[p/32][slot0+n][p%32]  (materialised Linear per-sample
 * gradients; x_rows / y_rows = rows of each chunk) */
int cg_outer_rows_cl(const float* Xt, long long x_rows, const float* Yt, long long y_rows, int M, int P,
                     int slot0, int B, float* out, cg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Small reductions around the contraction
 * ------------------------------------------------------------------------------------------- */

/* out[i] = sum_j src[i*ld + j]^2 (accumulate!=0: +=).  Used for per-sample bias norms and as the
 * first half of row_l2_norm. */
int cg_row_sumsq(const float* src, long long rows, long long cols, long long ld, float* out,
                 int accumulate, cg_stream_t stream);

/* out[i] = a[i] * b[i]  (closed-form Linear norm ||b a^T||_F^2 = ||a||^2 ||b||^2). */
int cg_vec_mul(const float* a, const float* b, float* out, long long n, cg_stream_t stream);

/* out[i] += w * a[i] * b[i]  (pairwise terms of the joint Linear norm). */
int cg_vec_fma(const float* a, const float* b, float w, float* out, long long n, cg_stream_t stream);

/* Clip factors (replaces norm_clipper.calc_clipping_factors, reference train.py:324-328):
 *   norm2 [n_params][n_slots] squared per-parameter norms.
 *   per_layer == 0: total = sqrt(sum_k norm2[k][s]); factors[0][s] = min(1, C[0]/(total+1e-6));
 *                   norms_out[0][s] = total
 *   per_layer != 0: factors[k][s] = min(1, C[k]/(sqrt(norm2[k][s])+1e-6)); norms_out[k][s] = sqrt(..)
 *   clip_lo..clip_hi: slots outside this range get factor 1 (unclipped non-private pass switch).
 *   c_scale in (0, 1]: the thresholds are multiplied by it (1 = the reference's formula; 1 - 2^-9 makes
 *   ||factor * G|| <= C strict under the TF32 rounding of the norms).
 * C is a DEVICE array (so adaptive clipping never syncs the host). */
int cg_clip_factors(const float* norm2, int n_params, int n_slots, int per_layer, const float* C, float c_scale,
                    int clip_lo, int clip_hi, float* factors, float* norms_out, cg_stream_t stream);

/* dst[r][slot*slot_stride + q] = tf32(src[...] * factor[slot]) for slot in [slot_lo, slot_hi). */
int cg_scale_slots(const float* src, float* dst, int rows, long long pitch, long long slot_stride,
                   int slot_lo, int slot_hi, const float* factor, cg_stream_t stream);

/* out[m][c][kh][kw] (+)= T[m][kh][kw*C + c]   (gradient-natural -> parameter layout) */
int cg_permute_accum(const float* T, float* out, int M, int C, int KH, int KW, int accumulate,
                     cg_stream_t stream);

/* out[r] (+)= sum_slot factor[slot] * rows_in[slot*R + r], slot in [slot_lo, slot_hi)
 * (clipped sum of per-sample bias gradients). */
int cg_weighted_colsum(const float* rows_in, const float* factor, int slot_lo, int slot_hi, int R,
                       float* out, int accumulate, cg_stream_t stream);

/* stat over the slots of one pass of norms [n_rows][n_slots]: out[k] = mean or max of
 * norms[k][slot_lo:slot_hi] times `scalar` (adaptive clipping, reference train.py:230-243,
 * without the per-parameter .cpu().item() syncs).  stat: 0 = mean, 1 = max. */
int cg_row_stat(const float* norms, int n_rows, int n_slots, int slot_lo, int slot_hi, int stat,
                float scalar, float* out, cg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Gaussian noise + finalisation of the step (replaces the patched optimizer.step():
 * upstream privacy_engine.step / _generate_noise; reference train.py:484).
 *   grad[i] = in[i] * (1/in_div)  +  (z_i * std) * (1/noise_div)   (each op rounded in fp32;
 *             torch's CUDA `tensor / python_scalar` multiplies by the fp32 reciprocal, and
 *             bit-exactness is defined against that)
 * z_i is bit-identical to the stream torch.normal(0, std, shape, generator=<CUDA generator with
 * (seed, offset)>) would draw on this device: Philox4_32_10, curand_normal4, block 256,
 * unroll 4, grid = min(SMs * maxThreadsPerSM/256, ceil(n/256)).  *offset_inc receives the amount
 * the generator offset advances ( ((n-1)/(256*grid*4)+1)*4 ).  in_div / noise_div <= 0 disables
 * the division.  `in` may be NULL (pure noise) or equal to `grad` (in-place add).
 * ------------------------------------------------------------------------------------------- */
int cg_noise_finalize(const float* in, float* grad, long long n, double in_div, double std,
                      double noise_div, unsigned long long seed, unsigned long long offset,
                      unsigned long long* offset_inc, cg_stream_t stream);
/* Same, but the standard deviation is std_mult * std_dev[0] with std_dev a DEVICE scalar (adaptive
 * clipping thresholds / immediate sensitivities that never visit the host).  Always draws. */
int cg_noise_finalize_dev(const float* in, float* grad, long long n, double in_div, double std_mult,
                          const float* std_dev, double noise_div, unsigned long long seed,
                          unsigned long long offset, unsigned long long* offset_inc, cg_stream_t stream);

/* CUDA-graph variant: the generator offset is read from DEVICE memory (offset_dev[0] + intra_offset), so a
 * captured step draws fresh noise at every replay; std = std_mult * (std_dev ? std_dev[0] : 1).
 * cg_philox_advance bumps the device offset once per step (captured as the step's last node). */
int cg_noise_finalize_graph(const float* in, float* grad, long long n, double in_div, double std_mult,
                            const float* std_dev, double noise_div, unsigned long long seed,
                            const unsigned long long* offset_dev, unsigned long long intra_offset,
                            unsigned long long* offset_inc, cg_stream_t stream);
int cg_philox_advance(unsigned long long* offset_dev, unsigned long long inc, cg_stream_t stream);

/* Multi-tensor form: ONE launch noises every parameter tensor of the step (the patched optimizer.step() loops over
 * the parameters, reference train.py:484; upstream privacy_engine.step).  Equivalent, bit for bit, to calling
 * cg_noise_finalize / _dev / _graph on the segments in order with the generator offset advancing between them:
 * Philox4_32_10 is counter based, so the element torch's thread `idx` would draw in its k-th loop trip is
 * philox(seed, subsequence = idx, counter = offset/4 + k) whatever launch geometry computes it.  Blocks look the
 * segment up in a small table (torch's grid for that tensor, its offset inside the step).
 *   std of segment s = std_mult * (std_dev ? std_dev[0] : 1); a segment with std_dev == NULL and std_mult == 0
 *   draws nothing and does not advance the offset (upstream _generate_noise returns zeros for sigma*C == 0).
 *   in_div_dev / noise_div_dev: optional DEVICE scalars that override in_div / noise_div (the global batch size
 *   that arrives with the allreduce under data parallelism); the fp32 reciprocal is taken on the device.
 *   offset_dev: optional DEVICE generator offset (CUDA-graph replay), added to `offset`.
 * *offset_inc receives the total advance over all segments. */
typedef struct cg_noise_seg {
  const float* in;           /* summed gradient; may be NULL (pure noise) or == grad           */
  float* grad;
  long long n;
  double std_mult;
  const float* std_dev;      /* device scalar or NULL                                          */
} cg_noise_seg;
int cg_noise_finalize_multi(const cg_noise_seg* segs, int n_segs, double in_div, const float* in_div_dev,
                            double noise_div, const float* noise_div_dev, unsigned long long seed,
                            unsigned long long offset, const unsigned long long* offset_dev,
                            unsigned long long* offset_inc, cg_stream_t stream);

/* Data parallel: the allreduce of the clipped sums AND the noise in one kernel over NVSwitch multicast memory
 * (replaces NCCL allreduce + cg_noise_finalize_multi; reference: opacus' DistributedDataParallel gradient exchange
 * followed by the patched step, train.py:484).  Every seg.in / seg.grad must lie inside ONE symmetric buffer of
 * buf_len floats that starts at local_base on this rank and is mapped at the multicast address mc_base on all `world`
 * ranks (torch.distributed._symmetric_memory).  Rank r processes the work blocks b = r (mod world): it loads the
 * all-rank sum with multimem.ld_reduce, divides by the all-rank sum of element count_off (mean = 1; each rank stores
 * its live sample count there), adds noise from the SAME Philox counters cg_noise_finalize_multi would use -- every
 * rank generates 1/world of the normals -- and multicasts the result into every rank's buffer (in place).
 * mc_base == NULL selects the peer-to-peer variant: peers[r] is rank r's buffer through its NVLink peer mapping
 * (peers[rank] == local_base, world <= 8); the sum is formed in rank order from plain 16-byte loads and the result is
 * stored to every peer.
 * The caller brackets the launch with cross-rank barriers (all sums written before / all results visible after). */
int cg_noise_finalize_allreduce(const cg_noise_seg* segs, int n_segs, int mean, unsigned long long seed,
                                unsigned long long offset, const unsigned long long* offset_dev,
                                unsigned long long* offset_inc, const float* local_base, float* mc_base,
                                float* const* peers, long long buf_len, long long count_off, int rank, int world,
                                cg_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Per-sample row norms (reference gradient_penalty.py:52-53, 60-61; immediate sensitivity,
 * train.py:457/469) and per-sample L2 clipping (reference backprop_clip.py:18-22)
 * ------------------------------------------------------------------------------------------- */
/* norms[i] = ||src[i, :]||_2 */
int cg_row_l2_norm(const float* src, long long rows, long long cols, float* norms, cg_stream_t stream);
/* gin[i][j] = g[i][j] * (gout[i] / norms[i])   (backward of row_l2_norm; 0 where norms == 0) */
int cg_row_l2_norm_bwd(const float* g, const float* norms, const float* gout, long long rows,
                       long long cols, float* gin, cg_stream_t stream);
/* out[0] = max_i v[i] */
int cg_vec_max(const float* v, long long n, float* out, cg_stream_t stream);
/* out[i,:] = norm_i > C ? C * (t[i,:] / norm_i) : t[i,:] ; norms_out optional */
int cg_l2_clip(const float* t, long long rows, long long cols, float C, float* out, float* norms_out,
               cg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CSLGAN_B200_H */
