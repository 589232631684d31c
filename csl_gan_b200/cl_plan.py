"""Channels-last capture / contraction plan of one layer (the main path; `grad_sample.LayerPlan`
dispatches here whenever the layer's window grid can be tiled into 32-position k-blocks).

Capture is one element-wise pass per tensor when the critic runs in `torch.channels_last` (the layout
cuDNN prefers on Blackwell): backprops -> Xt[(slot, q)][m], activations -> space-to-depth
Yt[plane][slot][hs][ws][c].  Other layouts are read through their strides (slower, still correct).
Replaces the same fork operations as the legacy plan (upstream opacus `_capture_activations`,
`_compute_*_grad_sample`; reference train.py:382-387, 399).
"""
from __future__ import annotations

import os

import ctypes as C
from typing import Optional

import torch
from torch import nn

from . import _lib as L

MERGE_BELOW = 16       # inputs with fewer channels fold the filter columns into the channel axis


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def _strides4(t: torch.Tensor):
    """(sn, sc, sh, sw) of a [B, C, H, W] or [B, C] tensor, in elements."""
    if t.dim() == 2:
        return t.stride(0), t.stride(1), 0, 0
    return t.stride(0), t.stride(1), t.stride(2), t.stride(3)


def _channels_fastest(t: torch.Tensor) -> torch.Tensor:
    """The staging kernels coalesce when the channel stride is 1; convert anything else once."""
    if t.dim() == 4 and t.stride(1) != 1 and t.shape[1] > 1:
        return t.contiguous(memory_format=torch.channels_last)
    if t.dim() == 2 and t.stride(1) != 1:
        return t.contiguous()
    return t


class ClLayerPlan:
    def __init__(self, name: str, layer: nn.Module, kind: str, w_idx: int, b_idx: Optional[int], use_ghost: bool = True,
                 use_half: bool = True):
        self.name, self.layer, self.kind, self.w_idx, self.b_idx = name, layer, kind, w_idx, b_idx
        self.use_ghost = use_ghost
        self.use_half = use_half

    # ------------------------------------------------------------------ geometry / buffers
    @staticmethod
    def geometry(layer: nn.Module, kind: str, act_shape):
        """(Cn, H, W, kh, kw, sh, sw, ph, pw, dh, dw, Ho, Wo, M) of the layer's unfolded operand."""
        if kind == "linear":
            return (layer.in_features, 1, 1, 1, 1, 1, 1, 0, 0, 1, 1, 1, 1, layer.out_features)
        kh, kw = layer.kernel_size
        sh, sw = layer.stride
        ph, pw = layer.padding
        dh, dw = layer.dilation
        if kind == "conv":
            Cn, H, W = act_shape[1], act_shape[2], act_shape[3]
            Ho = (H + 2 * ph - dh * (kh - 1) - 1) // sh + 1
            Wo = (W + 2 * pw - dw * (kw - 1) - 1) // sw + 1
            M = layer.out_channels
        else:
            Hin, Win = act_shape[2], act_shape[3]
            oph, opw = layer.output_padding
            H = (Hin - 1) * sh - 2 * ph + dh * (kh - 1) + oph + 1
            W = (Win - 1) * sw - 2 * pw + dw * (kw - 1) + opw + 1
            Cn, Ho, Wo, M = layer.out_channels, Hin, Win, layer.in_channels
        return (Cn, H, W, kh, kw, sh, sw, ph, pw, dh, dw, Ho, Wo, M)

    def setup(self, act: torch.Tensor, Bpad: int, max_passes: int):
        dev = act.device
        self.Bpad, self.max_passes = Bpad, max_passes
        S = self.S = Bpad * max_passes
        (Cn, H, W, kh, kw, sh, sw, ph, pw, dh, dw, Ho, Wo, M) = self.geometry(self.layer, self.kind, act.shape)
        self.Cn, self.KH, self.KW, self.Ho, self.Wo, self.M = Cn, kh, kw, Ho, Wo, M
        self.Q = Ho * Wo
        self.geom = L.UnfoldGeom(Cn, H, W, kh, kw, sh, sw, ph, pw, dh, dw, Ho, Wo)
        ghost_ok = (self.use_ghost and self.kind != "linear" and self.Q <= 128 and 128 % self.Q == 0
                    and Wo <= 256 and Ho <= 256)
        # thin inputs (the 3-channel image) fold the filter columns into the channel axis so a 32-wide
        # channel chunk is not 90% padding; small-Q layers keep the plain layout the ghost norms need
        merged = self.kind != "linear" and Cn < MERGE_BELOW and kw > 1 and not ghost_ok
        # FP16 operand containers (kind::f16: 16 contraction rows per instruction) wherever a per-sample k-block is a
        # multiple of 16 rows; TF32 words otherwise (per-sample groups of 8 | Q < 16 positions)
        self.half = bool(self.use_half and (self.kind == "linear" or self.Q >= 32 or self.Q % 16 == 0))
        dt = torch.float16 if self.half else torch.float32
        cw = self.cw = 64 if self.half else 32            # channels per 128-byte chunk row
        # thin inputs: TF32 folds the filter COLUMNS into the channels (3 x 5 = 15 of 32, five taps over kh); with
        # FP16's 64-channel chunks the whole window is folded (3 x 5 x 5 = 75 of 128: plain im2col rows, one tap,
        # one TMA box per k-block) -- the staged bytes are the same, the contraction is a single tile
        self.plan = L.plan_cl(self.geom, (2 if self.half else 1) if merged else 0, cw)
        self.n_planes = self.plan.n_rh * self.plan.n_rw
        self.kblock = L.cl_kblock_rows(self.geom, self.half)      # (rows, slots) per k-block of the clipped sum
        self.ldT = self.plan.n_taps * self.plan.Cs
        # thin layers: materialise the (small) per-sample gradients once instead of contracting twice
        self.thin = (self.kind != "linear" and not ghost_ok and self.Q >= 256 and M * self.ldT <= 16384)
        # ... and, where csrc/thin.cuh covers the geometry, straight from the critic's own tensors at capture time:
        # nothing of this layer is staged in HBM (CSLGAN_THIN_DIRECT=0 is the A/B switch)
        self.direct = bool(self.thin and self.half and self.kind == "conv" and self.plan.merged == 2
                           and os.environ.get("CSLGAN_THIN_DIRECT", "1") != "0" and L.thin_direct_ok(self.geom, M))
        # chunk-major staging: Xt[m/cw][slot*Q + q][cw], Yt[plane*n_cb + c/cw][slot][hs][ws][cw]
        self.x_chunks = _round_up(M, cw) // cw
        self.x_rows = S * self.Q
        self.n_cb = self.plan.Cp // cw
        if self.direct:
            self.Xt = self.Xc = self.Yt = None
            self.gnorm2 = torch.zeros(S, device=dev)
            self._act = {}
            self._pending = []                     # (act, grad_output, slot0, scale) captures not launched yet
        else:
            self.Xt = torch.zeros((self.x_chunks, self.x_rows, cw), device=dev, dtype=dt)
            self.Xc = torch.zeros((self.x_chunks, self.x_rows, cw), device=dev, dtype=dt)
            self.Yt = torch.zeros(self.n_planes * self.n_cb * S * self.plan.slot_stride, device=dev, dtype=dt)
        if self.half:
            # per-slot inverse staging scales (true value = staged * inv), absmax scratch, clipped-sum multipliers
            self.inv_x = torch.ones(S, device=dev)
            self.inv_y = torch.ones(S, device=dev)
            self.amax = torch.zeros(S, device=dev, dtype=torch.int32)
            self.mult = torch.zeros(S, device=dev)
            self.out_scale = torch.ones(2, device=dev)       # [0] = 2^E, [1] = scratch of cg_clip_mult
        self.bias_len = layer_bias_len(self.layer, self.kind)
        self.bias_rows = torch.zeros((S, self.bias_len), device=dev) if self.b_idx is not None else None
        if self.kind == "linear":
            self.asq = torch.zeros(S, device=dev)
            self.bsq = torch.zeros(S, device=dev)
        # gradient-natural accumulation buffer T[m][tap][c'] (== parameter layout for Linear and for
        # channels_last conv weights when kw is not merged)
        self.T = torch.zeros((M, self.ldT), device=dev)
        # ghost norms when Q | 128 (needs the un-merged plan, which is what small-Q layers have)
        self.ghost = ghost_ok
        self.Gs = torch.zeros((S, M * self.ldT), device=dev) if self.thin else None
        self._gs_joint = 1
        # clipped-sum GEMM on CTA pairs (cta_group::2) where the layer is wide enough; CSLGAN_NO_PAIR=1 is the
        # A/B switch for measurements
        self.pair = (not self.thin and os.environ.get("CSLGAN_NO_PAIR", "0") != "1"
                     and L.cl_pair_ok(M, self.geom, self.plan))
        if self.kind == "convT" and self.b_idx is not None:
            self._bias_scratch = torch.empty((_round_up(self.bias_len, 32) // 32, Bpad * H * W, 32), device=dev)
            self._HW = (H, W)

    # ------------------------------------------------------------------ capture
    def _stage_x(self, t: torch.Tensor, slot0: int, scale: float, bias_rows, sumsq):
        t = _channels_fastest(t)
        sn, sm, sh, sw = _strides4(t)
        if self.half:
            L.call("cg_stage_xt_h", L.ptr(t), sn, sm, sh, sw, t.shape[0], self.M, self.Ho, self.Wo, scale,
                   L.ptr(self.Xt), self.x_rows, slot0, L.ptr(bias_rows), L.ptr(sumsq), L.ptr(self.amax),
                   L.ptr(self.inv_x), L.stream_ptr(t.device))
            return
        L.call("cg_stage_xt", L.ptr(t), sn, sm, sh, sw, t.shape[0], self.M, self.Ho, self.Wo, scale,
               L.ptr(self.Xt), self.x_rows, slot0, L.ptr(bias_rows), L.ptr(sumsq), L.stream_ptr(t.device))

    def _stage_y(self, t: torch.Tensor, slot0: int, scale: float):
        if not (self.half and self.plan.merged == 2):
            # (the window-folded capture copies the image rows to shared memory through any strides: an NCHW image
            # batch needs no layout conversion pass)
            t = _channels_fastest(t)
        st = L.stream_ptr(t.device)
        if self.kind == "linear":
            # Q = 1: Yt is [p/32][slot][p%32], the same chunked row layout as Xt; the row kernel also
            # yields ||a||^2 for the closed-form norms
            if self.half:
                L.call("cg_stage_xt_h", L.ptr(t), t.stride(0), t.stride(1), 0, 0, t.shape[0], self.Cn, 1, 1, scale,
                       L.ptr(self.Yt), self.S, slot0, None, L.ptr(self.asq), L.ptr(self.amax), L.ptr(self.inv_y), st)
            else:
                L.call("cg_stage_xt", L.ptr(t), t.stride(0), t.stride(1), 0, 0, t.shape[0], self.Cn, 1, 1, scale,
                       L.ptr(self.Yt), self.S, slot0, None, L.ptr(self.asq), st)
            return
        sn, sc, sh, sw = _strides4(t)
        if self.half:
            L.call("cg_stage_yt_h", L.ptr(t), sn, sc, sh, sw, t.shape[0], C.byref(self.geom), C.byref(self.plan), scale,
                   L.ptr(self.Yt), self.S, slot0, L.ptr(self.amax), L.ptr(self.inv_y), st)
            return
        L.call("cg_stage_yt", L.ptr(t), sn, sc, sh, sw, t.shape[0], C.byref(self.geom), C.byref(self.plan), scale,
               L.ptr(self.Yt), self.S, slot0, st)

    def capture_activation(self, act: torch.Tensor, pass_idx: int):
        slot0 = pass_idx * self.Bpad
        if self.direct:
            if pass_idx == 0:
                self._pending = []                     # a new step: nothing of an abandoned one may linger
            self._act[pass_idx] = act                  # read by cg_thin_capture when the backprops arrive
            return
        if self.kind == "convT":
            self._stage_x(act, slot0, 1.0, None, None)
        else:
            self._stage_y(act, slot0, 1.0)

    def flush_direct(self):
        """Launch the deferred thin-layer captures (cg_thin_capture2 for a pair with equal strides and scale)."""
        pend, self._pending = self._pending, []
        while pend:
            act, g, slot0, scale = pend.pop(0)
            sn, sc, sh, sw = _strides4(act)
            st = L.stream_ptr(g.device)
            bias = self.bias_rows
            if pend and _strides4(pend[0][0]) == (sn, sc, sh, sw) and pend[0][3] == scale:
                act2, g2, slot2, _ = pend.pop(0)
                L.call("cg_thin_capture2", L.ptr(act), L.ptr(act2), sn, sc, sh, sw, L.ptr(g), L.ptr(g2), g.shape[0], g2.shape[0],
                       C.byref(self.geom), self.M, scale, L.ptr(self.Gs[slot0:]), L.ptr(self.Gs[slot2:]), self.Gs.shape[1],
                       L.ptr(self.gnorm2[slot0:]), L.ptr(self.gnorm2[slot2:]),
                       L.ptr(bias[slot0:]) if bias is not None else None, L.ptr(bias[slot2:]) if bias is not None else None, st)
            else:
                L.call("cg_thin_capture", L.ptr(act), sn, sc, sh, sw, L.ptr(g), g.shape[0], C.byref(self.geom), self.M, scale,
                       L.ptr(self.Gs[slot0:]), self.Gs.shape[1], L.ptr(self.gnorm2[slot0:]),
                       L.ptr(bias[slot0:]) if bias is not None else None, st)

    def capture_backprop(self, g: torch.Tensor, pass_idx: int, scale: float):
        slot0 = pass_idx * self.Bpad
        if self.direct:
            act = self._act.pop(pass_idx, None)
            if act is None or act.shape[0] != g.shape[0]:
                raise L.CslGanCudaError(f"{self.name}: backprops of pass {pass_idx} arrived without their activation")
            g = g.contiguous(memory_format=torch.channels_last)       # dense [B][Ho*Wo][M]; a no-op for a channels_last critic
            if g.data_ptr() % 16:
                g = g.clone(memory_format=torch.channels_last)        # (a sliced view: the TMA base must be 16-byte aligned)
            # the launch is deferred until the other passes of the step have arrived (or the norms are asked for): both
            # passes in ONE launch waste less of the last round of the 148 persistent CTAs than two launches
            self._pending.append((act, g, slot0, float(scale)))
            if len(self._pending) >= min(2, self.max_passes):
                self.flush_direct()
            return
        if self.kind == "convT":
            self._stage_y(g, slot0, scale)
            if self.bias_rows is not None:
                gg = _channels_fastest(g)
                sn, sm, sh, sw = _strides4(gg)
                H, W = self._HW
                L.call("cg_stage_xt", L.ptr(gg), sn, sm, sh, sw, gg.shape[0], self.bias_len, H, W, scale,
                       L.ptr(self._bias_scratch), self._bias_scratch.shape[1], 0,
                       L.ptr(self.bias_rows[slot0:]), None, L.stream_ptr(g.device))
        else:
            self._stage_x(g, slot0, scale, self.bias_rows, self.bsq if self.kind == "linear" else None)

    # ------------------------------------------------------------------ launches
    def _desc(self, X: torch.Tensor) -> L.ClDesc:
        d = L.ClDesc()
        d.Xt, d.xt_pitch, d.xt_rows, d.M = L.ptr(X), 32, self.x_rows, self.M
        d.Yt, d.n_slots_total = L.ptr(self.Yt), self.S
        d.max_ctas = 0
        if self.half:
            d.half, d.inv_x, d.inv_y = 1, L.ptr(self.inv_x), L.ptr(self.inv_y)
        return d

    def weight_norm2(self, norm2_row: torch.Tensor, pass_idx: int, B: int, n_joint: int = 1, ops=None):
        """`ops`: a list that collects the small operations of this phase for ONE cg_small_ops launch (None: launch
        them here)."""
        slot0 = pass_idx * self.Bpad
        st = L.stream_ptr(norm2_row.device)
        if self.kind == "linear":
            if n_joint == 1:
                if ops is not None:
                    ops.append(L.small_op(L.OP_MUL, self.asq[slot0:], norm2_row[slot0:], B, b=self.bsq[slot0:]))
                    return
                L.call("cg_vec_mul", L.ptr(self.asq[slot0:]), L.ptr(self.bsq[slot0:]), L.ptr(norm2_row[slot0:]), B, st)
                return
            # ||sum_p b_p a_p^T||_F^2 = sum_{p,p'} (a_p . a_p') (b_p . b_p')
            if not hasattr(self, "_ta"):
                self._ta = torch.empty(self.Bpad, device=norm2_row.device)
                self._tb = torch.empty(self.Bpad, device=norm2_row.device)
            out = norm2_row[slot0:]
            out[:B].zero_()
            for p1 in range(n_joint):
                for p2 in range(p1, n_joint):
                    s1, s2 = slot0 + p1 * self.Bpad, slot0 + p2 * self.Bpad
                    if p1 == p2:
                        a, b, w = self.asq[s1:], self.bsq[s1:], 1.0
                    else:
                        L.call("cg_rowpair_dot", L.ptr(self.Yt), self.S, self.n_cb, s1, s2, B, L.ptr(self._ta), 0, st)
                        L.call("cg_rowpair_dot", L.ptr(self.Xt), self.x_rows, self.x_chunks, s1, s2, B, L.ptr(self._tb), 0, st)
                        a, b, w = self._ta, self._tb, 2.0
                    L.call("cg_vec_fma", L.ptr(a), L.ptr(b), w, L.ptr(out), B, st)
            return
        if self.ghost and n_joint == 1:
            gd = L.GhostDesc()
            gd.Xt, gd.xt_pitch, gd.xt_rows = L.ptr(self.Xt), 32, self.x_rows
            gd.Yt, gd.n_slots_total, gd.O = L.ptr(self.Yt), self.S, self.M
            gd.slot0, gd.n_slots = slot0, B
            gd.norm2, gd.max_ctas = L.ptr(norm2_row[slot0:]), 0
            if self.half:
                gd.half, gd.inv_x, gd.inv_y = 1, L.ptr(self.inv_x), L.ptr(self.inv_y)
            L.call("cg_ghost_norm", C.byref(gd), C.byref(self.geom), C.byref(self.plan), st)
            return
        d = self._desc(self.Xt)
        d.group_mode, d.n_groups, d.slot_lo, d.slot_hi = L.GROUP_SAMPLE, B, slot0, slot0 + B
        d.n_seg, d.seg_stride = n_joint, self.Bpad
        if self.direct:
            self.flush_direct()
            self._gs_joint = 1
            if ops is not None:
                ops.append(L.small_op(L.OP_COPY, self.gnorm2[slot0:], norm2_row[slot0:], B))
                return
            norm2_row[slot0:slot0 + B].copy_(self.gnorm2[slot0:slot0 + B])
            return
        if self.thin:
            # G[slot][m][tap][c'] once (joint mode: the per-sample sum over passes lands in pass 0's slots);
            # the norms are then a row reduction over |theta_layer| floats per sample
            R = self.Gs.shape[1]
            d.epi, d.out, d.out_group_stride = L.EPI_STORE_NATURAL, L.ptr(self.Gs[slot0:]), R
            L.call("cg_cl_contract", C.byref(d), C.byref(self.geom), C.byref(self.plan), st)
            L.call("cg_row_sumsq", L.ptr(self.Gs[slot0:]), B, R, R, L.ptr(norm2_row[slot0:]), 0, st)
            self._gs_joint = n_joint
            return
        d.epi, d.out, d.out_group_stride = L.EPI_SUMSQ, L.ptr(norm2_row[slot0:]), 0
        L.call("cg_cl_contract", C.byref(d), C.byref(self.geom), C.byref(self.plan), st)

    def bias_norm2(self, norm2_row: torch.Tensor, pass_idx: int, B: int, n_joint: int = 1, ops=None):
        slot0 = pass_idx * self.Bpad
        st = L.stream_ptr(norm2_row.device)
        if ops is not None and n_joint == 1:
            if self.kind == "linear":
                ops.append(L.small_op(L.OP_COPY, self.bsq[slot0:], norm2_row[slot0:], B))
            else:
                ops.append(L.small_op(L.OP_ROW_SUMSQ, self.bias_rows[slot0:], norm2_row[slot0:], B, R=self.bias_rows.shape[1]))
            return
        if n_joint > 1:
            R = self.bias_rows.shape[1]
            L.call("cg_joint_rows_sumsq", L.ptr(self.bias_rows), R, slot0, self.Bpad, n_joint, B,
                   L.ptr(norm2_row[slot0:]), st)
            return
        if self.kind == "linear":
            norm2_row[slot0:slot0 + B].copy_(self.bsq[slot0:slot0 + B])
            return
        R = self.bias_rows.shape[1]
        L.call("cg_row_sumsq", L.ptr(self.bias_rows[slot0:]), B, R, R, L.ptr(norm2_row[slot0:]), 0, st)

    def clip_mult_op(self, factor_row: torch.Tensor, slot_lo: int, slot_hi: int):
        """The clip-multiplier computation of scale_backprops() as a cg_small_ops entry (None when this layer has
        none: thin layers, TF32 operands, more than 65536 slots); scale_backprops(mult_ready=True) then skips it."""
        if self.thin or not self.half or slot_hi - slot_lo > 65536:
            return None
        return L.small_op(L.OP_CLIP_MULT, factor_row, self.mult, slot_hi - slot_lo, b=self.inv_x, c=self.inv_y,
                          out2=self.out_scale, lo=slot_lo)

    def scale_seg(self, slot_lo: int, slot_hi: int):
        """The cg_scale_slots_h launch of scale_backprops(mult_ready=True) as a cg_scale_slots_h_multi entry, or None."""
        if self.thin or not self.half:
            return None
        return (self.Xt, self.Xc, self.mult, self.x_rows * self.cw, self.Q * self.cw, self.x_chunks, slot_lo, slot_hi)

    def scale_backprops(self, factor_row: torch.Tensor, slot_lo: int, slot_hi: int, factor_shift: int = 0,
                        mult_ready: bool = False):
        """Xc = tf32(Xt * factor[slot - factor_shift]): every chunk of Xt is a row of slot-sized (Q*32) segments.
        factor_shift != 0 reuses pass 0's (joint) factors for a later pass."""
        if self.thin:
            return                      # the clipped sum comes from the materialised per-sample gradients
        if self.half:
            # factor * inv_x * inv_y, brought below 1 by a common power of two that the GEMM epilogue undoes
            if factor_shift:
                raise L.CslGanCudaError(f"{self.name}: joint clipping needs TF32 operands")
            st = L.stream_ptr(factor_row.device)
            if not mult_ready:
                L.call("cg_clip_mult", L.ptr(factor_row), L.ptr(self.inv_x), L.ptr(self.inv_y), slot_lo, slot_hi,
                       L.ptr(self.mult), L.ptr(self.out_scale), st)
            L.call("cg_scale_slots_h", L.ptr(self.Xt), L.ptr(self.Xc), self.x_chunks, self.x_rows * self.cw,
                   self.Q * self.cw, slot_lo, slot_hi, L.ptr(self.mult), st)
            return
        L.call("cg_scale_slots", L.ptr(self.Xt), L.ptr(self.Xc), self.x_chunks, self.x_rows * 32, self.Q * 32,
               slot_lo, slot_hi, L.ptr(factor_row) - 4 * factor_shift, L.stream_ptr(factor_row.device))

    def weighted_sum(self, out_w: torch.Tensor, slot_lo: int, slot_hi: int, sm_count: int, accumulate: bool,
                     factor_row: Optional[torch.Tensor] = None, prezeroed: bool = False, ops=None):
        """`prezeroed`: out_w already holds zeros (the engine clears the whole flat buffer once per step).
        `ops`: thin layers append their weighted column sum to this cg_small_ops table instead of launching it."""
        st = L.stream_ptr(out_w.device)
        if self.thin:
            return self._thin_weighted_sum(out_w, slot_lo, slot_hi, accumulate, factor_row, st, prezeroed, ops)
        d = self._desc(self.Xc)
        # tiles per K range (mirror of cg_cl_contract) -> split K so the grid covers the machine ~2x
        n_cb, n_taps = self.n_cb, self.plan.n_taps
        maxc = 256 // self.cw                              # chunks of a 256-column tile
        if n_cb >= maxc:
            parts = (n_cb + maxc - 1) // maxc
            cpt = (n_cb + parts - 1) // parts
            n_nt = n_taps * ((n_cb + cpt - 1) // cpt)
        else:
            tpt = min(maxc // n_cb, n_taps)
            parts = (n_taps + tpt - 1) // tpt
            tpt = (n_taps + parts - 1) // parts
            n_nt = (n_taps + tpt - 1) // tpt
        n_tiles = ((self.M + 127) // 128) * n_nt
        kb_rows, kb_s = self.kblock
        units = ((slot_hi - slot_lo) * (self.Q // kb_rows) if kb_s == 1 else (slot_hi - slot_lo + kb_s - 1) // kb_s)
        if self.pair:
            # CTA pairs: 256 x 256 tiles (two half tiles of one tap x four chunks), one pair per two SMs
            n_tiles = (self.M // 256) * ((n_taps * (self.plan.Cp // 128) + 1) // 2)
            n_groups = _pick_split_k(n_tiles, units, sm_count // 2)
        else:
            n_groups = _pick_split_k(n_tiles, units, sm_count)
        d.pair = 1 if self.pair else 0
        d.max_ctas = sm_count                              # (the engine passes fewer SMs while an allreduce is in flight)
        if self.half:
            d.out_scale = L.ptr(self.out_scale)
        d.group_mode, d.n_groups, d.slot_lo, d.slot_hi = L.GROUP_SPLITK, n_groups, slot_lo, slot_hi
        # where does the gradient-natural layout T[m][tap][c'] already equal the parameter's memory?
        natural = None
        if self.kind == "linear":
            natural = out_w
        elif (out_w.dim() == 4 and not out_w.is_contiguous()
              and out_w.is_contiguous(memory_format=torch.channels_last)):
            # channels_last weight memory is [m][kh][kw][c] = T[m][tap][c'] (merged: c' = kw*C + c)
            natural = out_w.permute(0, 2, 3, 1)
        target = natural if natural is not None else self.T
        if target is self.T or not (accumulate or prezeroed):
            target.zero_()
        d.epi, d.out, d.out_group_stride = L.EPI_ACCUM, L.ptr(target), 0
        L.call("cg_cl_contract", C.byref(d), C.byref(self.geom), C.byref(self.plan), st)
        if target is self.T:
            dst = out_w
            if not out_w.is_contiguous():
                if not (out_w.dim() == 4 and out_w.is_contiguous(memory_format=torch.channels_last)):
                    raise L.CslGanCudaError(f"{self.name}: unsupported weight memory layout {out_w.stride()}")
                dst = torch.empty(out_w.shape, device=out_w.device)      # contiguous scratch, copied below
            # T[m][kh][kw*C + c] (merged or not: the column index is (kh, kw, c) either way)
            L.call("cg_permute_accum", L.ptr(self.T), L.ptr(dst), self.M, self.Cn, self.KH, self.KW,
                   1 if (accumulate and dst is out_w) else 0, st)
            if dst is not out_w:
                out_w.add_(dst) if accumulate else out_w.copy_(dst)

    def _thin_weighted_sum(self, out_w, slot_lo, slot_hi, accumulate, factor_row, st, prezeroed=False, ops=None):
        """sum_slot factor[slot] * G[slot] over the materialised per-sample gradients (natural layout)."""
        if factor_row is None:
            raise L.CslGanCudaError(f"{self.name}: the thin-layer path needs the clip factors")
        if self._gs_joint > 1:
            slot_hi = min(slot_hi, slot_lo + self.Bpad)       # joint sums live in pass 0's slots only
        R = self.Gs.shape[1]
        natural = None
        if (out_w.dim() == 4 and not out_w.is_contiguous()
                and out_w.is_contiguous(memory_format=torch.channels_last)):
            natural = out_w.permute(0, 2, 3, 1)               # [m][kh][kw][c] == T[m][tap][c'] (merged or not)
        target = natural if natural is not None else self.T
        if ops is not None and natural is not None and (accumulate or prezeroed):
            ops.append(L.small_op(L.OP_WCOLSUM, self.Gs, target, slot_hi - slot_lo, b=factor_row, R=R, lo=slot_lo))
            return
        L.call("cg_weighted_colsum", L.ptr(self.Gs), L.ptr(factor_row), slot_lo, slot_hi, R, L.ptr(target),
               1 if ((accumulate or prezeroed) and natural is not None) else 0, st)
        if natural is None:
            if not out_w.is_contiguous():
                raise L.CslGanCudaError(f"{self.name}: unsupported weight memory layout {out_w.stride()}")
            L.call("cg_permute_accum", L.ptr(self.T), L.ptr(out_w), self.M, self.Cn, self.KH, self.KW,
                   1 if accumulate else 0, st)

    def bias_weighted_sum(self, out_b: torch.Tensor, factor_row: torch.Tensor, slot_lo: int, slot_hi: int,
                          accumulate: bool, factor_shift: int = 0, ops=None):
        R = self.bias_rows.shape[1]
        if ops is not None and factor_shift == 0 and accumulate:
            ops.append(L.small_op(L.OP_WCOLSUM, self.bias_rows, out_b, slot_hi - slot_lo, b=factor_row, R=R, lo=slot_lo))
            return
        L.call("cg_weighted_colsum", L.ptr(self.bias_rows), L.ptr(factor_row) - 4 * factor_shift, slot_lo, slot_hi, R,
               L.ptr(out_b), 1 if accumulate else 0, L.stream_ptr(out_b.device))

    def materialize(self, pass_idx: int, B: int) -> torch.Tensor:
        slot0 = pass_idx * self.Bpad
        w = self.layer.weight
        if self.direct:
            self.flush_direct()
            # Gs[slot][m][kh][kw][c] -> [B][m][c][kh][kw]
            return (self.Gs[slot0:slot0 + B].view(B, self.M, self.KH, self.KW, self.Cn).permute(0, 1, 4, 2, 3)
                    .contiguous())
        out = torch.zeros((B,) + tuple(w.shape), device=w.device)
        st = L.stream_ptr(w.device)
        if self.kind == "linear":
            if self.half:
                L.call("cg_outer_rows_cl_h", L.ptr(self.Xt), self.x_rows, L.ptr(self.Yt), self.S,
                       self.M, self.Cn, slot0, B, L.ptr(self.inv_x), L.ptr(self.inv_y), L.ptr(out), st)
            else:
                L.call("cg_outer_rows_cl", L.ptr(self.Xt), self.x_rows, L.ptr(self.Yt), self.S,
                       self.M, self.Cn, slot0, B, L.ptr(out), st)
            return out
        d = self._desc(self.Xt)
        d.group_mode, d.n_groups, d.slot_lo, d.slot_hi = L.GROUP_SAMPLE, B, slot0, slot0 + B
        d.epi, d.out, d.out_group_stride = L.EPI_STORE, L.ptr(out), w.numel()
        L.call("cg_cl_contract", C.byref(d), C.byref(self.geom), C.byref(self.plan), st)
        return out


def _pick_split_k(n_tiles: int, units: int, sm_count: int) -> int:
    """Split-K group count for the clipped-sum GEMM: items = n_tiles * groups are dealt round-robin to one
    persistent CTA per SM, so what matters is how full the last wave is (100 tiles x 2 groups on 148 SMs is
    1.35 waves = 68 % busy; x 10 groups is 6.76 waves = 96.5 %).  Every group costs one more split-K epilogue
    (a 128 x N tile of `red.add`) per tile, hence the small per-group penalty and the floor of 8 k-blocks."""
    g_max = max(1, min(64, units // 8 if units >= 8 else 1))
    best, best_score = 1, -1.0
    for g in range(1, g_max + 1):
        items = n_tiles * g
        waves = -(-items // sm_count)
        score = items / (waves * sm_count) - 0.004 * g
        if score > best_score + 1e-9:
            best, best_score = g, score
    return best


def layer_bias_len(layer: nn.Module, kind: str) -> int:
    if kind == "linear":
        return layer.out_features
    return layer.out_channels
