"""Per-layer capture and contraction plans (Linear / Conv2d / ConvTranspose2d).

Replaces the fork's forward/backward hook bodies and grad samplers (upstream opacus
`_capture_activations`, `_compute_linear_grad_sample`, `_compute_conv_grad_sample`; fired from
reference train.py:382-387) with staging kernels that lay the two operands of each layer's
per-sample contraction out K-major for the tcgen05 kernel (include/cslgan_b200.h).  The
B x |theta| `grad_sample` tensor is never written unless a caller explicitly materialises it.

slot = pass * Bpad + n, with Bpad = batch size rounded up to 32 so that every pass starts on a
k-block boundary of the Linear weighted-sum GEMM (whose contraction index is the slot itself).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch
from torch import nn

from . import _lib as L

KBLK = 32  # fp32 elements per 128-byte swizzle row = one k-block


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class LayerPlan:
    """Capture buffers + kernel launches for one supported layer."""

    def __init__(self, name: str, layer: nn.Module, w_idx: int, b_idx: Optional[int]):
        self.name = name
        self.layer = layer
        self.w_idx = w_idx
        self.b_idx = b_idx
        if isinstance(layer, nn.Linear):
            self.kind = "linear"
        elif isinstance(layer, nn.ConvTranspose2d):
            self.kind = "convT"
        elif isinstance(layer, nn.Conv2d):
            self.kind = "conv"
        else:  # pragma: no cover
            raise NotImplementedError(type(layer))
        if self.kind != "linear":
            if layer.groups != 1:
                raise NotImplementedError(f"{name}: grouped convolutions are not supported by the DP engine")
            if getattr(layer, "padding_mode", "zeros") != "zeros":
                raise NotImplementedError(f"{name}: only zero padding is supported")
            if isinstance(layer.padding, str):
                raise NotImplementedError(f"{name}: string padding modes are not supported")
        self.ready = False
        self.sig = None
        self.Bpad = 0
        self.max_passes = 0
        self.use_ghost = True
        self.use_half = L.default_operand_dtype() == "f16"   # FP16 operand containers on the channels-last path
        self.gplan = None
        self.force_legacy = False          # tests: exercise the kw-plane (legacy) path on any geometry
        self.impl = None                   # ClLayerPlan when the channels-last path applies

    # ------------------------------------------------------------------ geometry / buffers
    def _setup(self, act: torch.Tensor, Bpad: int, max_passes: int):
        dev = act.device
        lay = self.layer
        self.Bpad, self.max_passes = Bpad, max_passes
        S = Bpad * max_passes
        self.S = S
        if self.kind == "linear":
            if act.dim() != 2:
                raise NotImplementedError(
                    f"{self.name}: Linear layers with >2-D inputs are not supported (got {tuple(act.shape)})")
            self.M, self.Cn, self.KH, self.KW = lay.out_features, lay.in_features, 1, 1
            self.Q, self.Qpad = 1, 1
            self.x_slot_stride = self.y_slot_stride = 1
            self.tap_row0 = [0] * L.CG_MAX_KH
            self.tap_coloff = [0] * L.CG_MAX_KH
            pitch = _round_up(S, 4)
            self.X = torch.zeros((self.M, pitch), device=dev)
            self.Xc = torch.zeros((self.M, pitch), device=dev)
            self.Y = torch.zeros((self.Cn, pitch), device=dev)
            self.x_cols = self.y_cols = S
            self.asq = torch.zeros(S, device=dev)
            self.bsq = torch.zeros(S, device=dev)
            self.bias_rows = torch.zeros((S, self.M), device=dev) if self.b_idx is not None else None
        else:
            kh, kw = lay.kernel_size
            sh, sw = lay.stride
            ph, pw = lay.padding
            dh, dw = lay.dilation
            if self.kind == "conv":
                # unfolded operand = layer input; window positions = output pixels
                Cn, H, W = act.shape[1], act.shape[2], act.shape[3]
                Ho = (H + 2 * ph - dh * (kh - 1) - 1) // sh + 1
                Wo = (W + 2 * pw - dw * (kw - 1) - 1) // sw + 1
                self.M = lay.out_channels
                self.bias_len = lay.out_channels
            else:
                # ConvTranspose2d: unfolded operand = grad wrt output; window positions = input pixels
                Hin, Win = act.shape[2], act.shape[3]
                oph, opw = lay.output_padding
                H = (Hin - 1) * sh - 2 * ph + dh * (kh - 1) + oph + 1
                W = (Win - 1) * sw - 2 * pw + dw * (kw - 1) + opw + 1
                Cn = lay.out_channels
                Ho, Wo = Hin, Win
                self.M = lay.in_channels
                self.bias_len = lay.out_channels
            self.Cn, self.KH, self.KW = Cn, kh, kw
            self.geom, self.plan = L.plan_unfold(Cn, H, W, kh, kw, sh, sw, ph, pw, dh, dw, Ho, Wo)
            self.Q = Ho * Wo
            self.Wo, self.Wop = Wo, self.plan.Wop
            self.Qpad = _round_up(Ho * self.Wop, KBLK)
            self.x_slot_stride = self.Qpad
            self.y_slot_stride = self.plan.slot_stride
            self.tap_row0 = list(self.plan.tap_row0)
            self.tap_coloff = list(self.plan.tap_coloff)
            self.x_cols = S * self.Qpad
            self.y_cols = S * self.plan.slot_stride
            self.X = torch.zeros((self.M, _round_up(self.x_cols, 4)), device=dev)
            self.Xc = torch.zeros((self.M, _round_up(self.x_cols, 4)), device=dev)
            self.Y = torch.zeros((self.plan.rows, _round_up(self.y_cols, 4)), device=dev)
            self.bias_rows = torch.zeros((S, self.bias_len), device=dev) if self.b_idx is not None else None
            if self.kind == "convT" and self.b_idx is not None:
                self._bias_scratch = torch.empty((self.bias_len, _round_up(Bpad * H * W, 4)), device=dev)
        self.nkb = self.Qpad // KBLK if self.kind != "linear" else 1
        # gradient-natural accumulation buffer T[m][kh][kw*C + c] (Linear: == parameter layout)
        self.T = torch.zeros((self.M, self.KH * self.KW * self.Cn), device=dev)
        self.ready = True

    def _ensure(self, act: torch.Tensor, Bpad: int, max_passes: int):
        sig = (tuple(act.shape[1:]), Bpad, max_passes, act.device)
        if not self.ready or sig != self.sig:
            from .cl_plan import ClLayerPlan
            self.impl = None
            if self.kind == "linear" and act.dim() != 2:
                raise NotImplementedError(
                    f"{self.name}: Linear layers with >2-D inputs are not supported (got {tuple(act.shape)})")
            g = ClLayerPlan.geometry(self.layer, self.kind, act.shape)
            if not self.force_legacy and L.cl_supported(g[11], g[12]):
                self.impl = ClLayerPlan(self.name, self.layer, self.kind, self.w_idx, self.b_idx, self.use_ghost,
                                        self.use_half)
                self.impl.setup(act, Bpad, max_passes)
                self.Bpad, self.max_passes = Bpad, max_passes
                self.ready = True
            else:
                self._setup(act, Bpad, max_passes)
            self.sig = sig

    @property
    def path(self) -> str:
        return "channels_last" if self.impl is not None else "kw_planes"

    # ------------------------------------------------------------------ capture
    def capture_activation(self, act: torch.Tensor, pass_idx: int, Bpad: int, max_passes: int):
        act = act.detach()
        if not act.is_cuda or act.dtype != torch.float32:
            L.require_cuda_f32(act, f"{self.name}: activation")
        self._ensure(act, Bpad, max_passes)
        if self.impl is not None:
            return self.impl.capture_activation(act, pass_idx)
        act = act.contiguous()
        B = act.shape[0]
        slot0 = pass_idx * self.Bpad
        st = L.stream_ptr(act.device)
        if self.kind == "linear":
            L.call("cg_stage_rows_t", L.ptr(act), B, self.Cn, 1.0, L.ptr(self.Y), self.Y.stride(0), slot0,
                   None, L.ptr(self.asq), st)
        elif self.kind == "conv":
            L.call("cg_stage_unfold", L.ptr(act), B, C.byref(self.geom), C.byref(self.plan), 1.0,
                   L.ptr(self.Y), self.Y.stride(0), slot0, st)
        else:  # convT: the activation is the plain operand
            L.call("cg_stage_rows", L.ptr(act), B, self.M, self.Q, self.Wo, self.Wop, self.Qpad, 1.0,
                   L.ptr(self.X), self.X.stride(0), slot0, None, st)

    def capture_backprop(self, grad_out: torch.Tensor, pass_idx: int, scale: float):
        if self.impl is not None:
            g = grad_out.detach()
            if not g.is_cuda or g.dtype != torch.float32:
                L.require_cuda_f32(g, f"{self.name}: backprop")
            return self.impl.capture_backprop(g, pass_idx, scale)
        g = L.require_cuda_f32(grad_out.detach(), f"{self.name}: backprop")
        B = g.shape[0]
        slot0 = pass_idx * self.Bpad
        st = L.stream_ptr(g.device)
        if self.kind == "linear":
            L.call("cg_stage_rows_t", L.ptr(g), B, self.M, scale, L.ptr(self.X), self.X.stride(0), slot0,
                   L.ptr(self.bias_rows), L.ptr(self.bsq), st)
        elif self.kind == "conv":
            # rowsum is indexed by absolute slot inside the kernel -> pass the buffer base
            L.call("cg_stage_rows", L.ptr(g), B, self.M, self.Q, self.Wo, self.Wop, self.Qpad, scale,
                   L.ptr(self.X), self.X.stride(0), slot0, L.ptr(self.bias_rows), st)
        else:
            L.call("cg_stage_unfold", L.ptr(g), B, C.byref(self.geom), C.byref(self.plan), scale,
                   L.ptr(self.Y), self.Y.stride(0), slot0, st)
            if self.bias_rows is not None:
                hw = g.shape[2] * g.shape[3]
                L.call("cg_stage_rows", L.ptr(g), B, self.bias_len, hw, hw, hw, hw, scale,
                       L.ptr(self._bias_scratch), self._bias_scratch.stride(0), 0, L.ptr(self.bias_rows[slot0:]), st)

    def live_bias_rows(self) -> torch.Tensor:
        return self.impl.bias_rows if self.impl is not None else self.bias_rows

    # ------------------------------------------------------------------ contraction launches
    def _desc(self, X: torch.Tensor) -> L.ContractDesc:
        d = L.ContractDesc()
        d.X, d.x_pitch, d.x_rows, d.x_cols = L.ptr(X), X.stride(0), self.M, self.x_cols
        d.Y, d.y_pitch, d.y_rows, d.y_cols = L.ptr(self.Y), self.Y.stride(0), self.Y.shape[0], self.y_cols
        d.M, d.C, d.KH, d.KW = self.M, self.Cn, self.KH, self.KW
        for i in range(L.CG_MAX_KH):
            d.tap_row0[i] = self.tap_row0[i]
            d.tap_coloff[i] = self.tap_coloff[i]
        d.nkb = self.nkb
        d.x_slot_stride, d.y_slot_stride = self.x_slot_stride, self.y_slot_stride
        d.block_n = 0
        d.max_ctas = 0
        return d

    def weight_norm2(self, norm2_row: torch.Tensor, pass_idx: int, B: int, n_joint: int = 1, ops=None):
        """norm2_row[slot] += ||G_slot||_F^2 for the slots of one pass.  `ops` (channels-last plans): collect the
        small launches of this phase into one cg_small_ops table instead of issuing them."""
        if self.impl is not None:
            return self.impl.weight_norm2(norm2_row, pass_idx, B, n_joint, ops)
        slot0 = pass_idx * self.Bpad
        st = L.stream_ptr(norm2_row.device)
        if self.kind == "linear":
            if n_joint != 1:
                raise NotImplementedError("joint (accum_passes=True) norms for Linear layers")
            L.call("cg_vec_mul", L.ptr(self.asq[slot0:]), L.ptr(self.bsq[slot0:]), L.ptr(norm2_row[slot0:]), B, st)
            return
        d = self._desc(self.X)
        d.group_mode, d.n_groups = L.GROUP_SAMPLE, B
        d.slot_lo, d.slot_hi, d.spg = slot0, slot0 + B, 1
        d.n_seg, d.seg_stride = n_joint, self.Bpad
        d.epi, d.out, d.out_group_stride = L.EPI_SUMSQ, L.ptr(norm2_row[slot0:]), 0
        L.call("cg_contract", C.byref(d), st)

    def bias_norm2(self, norm2_row: torch.Tensor, pass_idx: int, B: int, n_joint: int = 1, ops=None):
        if self.impl is not None:
            return self.impl.bias_norm2(norm2_row, pass_idx, B, n_joint, ops)
        slot0 = pass_idx * self.Bpad
        st = L.stream_ptr(norm2_row.device)
        if n_joint > 1:
            R = self.bias_rows.shape[1]
            L.call("cg_joint_rows_sumsq", L.ptr(self.bias_rows), R, slot0, self.Bpad, n_joint, B,
                   L.ptr(norm2_row[slot0:]), st)
            return
        if self.kind == "linear":
            # per-sample bias gradient of a Linear layer is the (scaled) backprop itself
            norm2_row[slot0:slot0 + B].copy_(self.bsq[slot0:slot0 + B])
            return
        R = self.bias_rows.shape[1]
        L.call("cg_row_sumsq", L.ptr(self.bias_rows[slot0:]), B, R, R, L.ptr(norm2_row[slot0:]), 0, st)

    def clip_mult_op(self, factor_row: torch.Tensor, slot_lo: int, slot_hi: int):
        """cg_small_ops entry that prepares scale_backprops(mult_ready=True), or None."""
        if self.impl is not None:
            return self.impl.clip_mult_op(factor_row, slot_lo, slot_hi)
        return None

    def scale_seg(self, slot_lo: int, slot_hi: int):
        """cg_scale_slots_h_multi entry equivalent to scale_backprops(mult_ready=True), or None."""
        if self.impl is not None:
            return self.impl.scale_seg(slot_lo, slot_hi)
        return None

    def scale_backprops(self, factor_row: torch.Tensor, slot_lo: int, slot_hi: int, factor_shift: int = 0,
                        mult_ready: bool = False):
        """Xc = tf32(X * factor[slot - factor_shift]) over the slot range (clip factors folded into one operand)."""
        if self.impl is not None:
            return self.impl.scale_backprops(factor_row, slot_lo, slot_hi, factor_shift, mult_ready)
        st = L.stream_ptr(factor_row.device)
        L.call("cg_scale_slots", L.ptr(self.X), L.ptr(self.Xc), self.M, self.X.stride(0), self.x_slot_stride,
               slot_lo, slot_hi, L.ptr(factor_row) - 4 * factor_shift, st)

    def weighted_sum(self, out_w: torch.Tensor, slot_lo: int, slot_hi: int, sm_count: int, accumulate: bool,
                     factor_row: Optional[torch.Tensor] = None, prezeroed: bool = False, ops=None):
        """out_w (+)= sum_slot Xc[:, slot] (x) Y[:, slot]: ONE split-K GEMM over all slots (factor_row is only
        used by the thin-layer path of the channels-last plan, which sums materialised per-sample gradients).
        `prezeroed`: out_w already holds zeros; `ops`: see ClLayerPlan.weighted_sum."""
        if self.impl is not None:
            return self.impl.weighted_sum(out_w, slot_lo, slot_hi, sm_count, accumulate, factor_row, prezeroed, ops)
        if prezeroed:
            accumulate = True
        st = L.stream_ptr(out_w.device)
        d = self._desc(self.Xc)
        # mirror of the tile choice in cg_contract (csrc/abi.cu)
        kwc = self.KW * self.Cn
        parts = (kwc + 255) // 256
        bn = _round_up((kwc + parts - 1) // parts, 16)
        n_rb = (kwc + bn - 1) // bn
        ks = max(1, min(self.KH, 256 // bn)) if n_rb == 1 else 1
        n_tiles = ((self.M + 127) // 128) * ((self.KH + ks - 1) // ks) * n_rb
        if self.kind == "linear":
            # contraction index = slot; units of 32 slots
            u_lo, u_hi = slot_lo // KBLK, (slot_hi + KBLK - 1) // KBLK
            assert slot_lo % KBLK == 0
            d.x_slot_stride = d.y_slot_stride = KBLK
            d.x_cols = d.y_cols = slot_hi if slot_hi < self.x_cols else self.x_cols
            units = u_hi - u_lo
        else:
            u_lo, u_hi = slot_lo, slot_hi
            units = slot_hi - slot_lo
        # split K so that the grid covers the machine about twice, but keep >= 4 k-blocks per split
        want = max(1, (2 * sm_count) // max(1, n_tiles))
        spg = max(1, (units + want - 1) // want)
        min_units = max(1, (4 + self.nkb - 1) // self.nkb)
        spg = max(spg, min_units)
        n_groups = (units + spg - 1) // spg
        d.group_mode, d.n_groups = L.GROUP_SPLITK, n_groups
        d.slot_lo, d.slot_hi, d.spg = u_lo, u_hi, spg
        d.n_seg, d.seg_stride = 1, 1
        conv_like = self.kind != "linear"
        # the GEMM accumulates in the gradient-natural layout [m][kh][kw][c]; that IS the memory order
        # of a channels_last weight, so such parameters need no permute at all
        natural = None
        if conv_like and out_w.dim() == 4 and not out_w.is_contiguous() \
                and out_w.is_contiguous(memory_format=torch.channels_last):
            natural = out_w.permute(0, 2, 3, 1)
        if not conv_like:
            target = out_w
        elif natural is not None:
            target = natural
        else:
            target = self.T
        if target is self.T or not accumulate:
            target.zero_()
        d.epi, d.out, d.out_group_stride = L.EPI_ACCUM, L.ptr(target), 0
        L.call("cg_contract", C.byref(d), st)
        if target is self.T:
            if not out_w.is_contiguous():
                raise L.CslGanCudaError(f"{self.name}: unsupported weight memory layout {out_w.stride()}")
            L.call("cg_permute_accum", L.ptr(self.T), L.ptr(out_w), self.M, self.Cn, self.KH, self.KW,
                   1 if accumulate else 0, st)

    def bias_weighted_sum(self, out_b: torch.Tensor, factor_row: torch.Tensor, slot_lo: int, slot_hi: int,
                          accumulate: bool, factor_shift: int = 0, ops=None):
        if self.impl is not None:
            return self.impl.bias_weighted_sum(out_b, factor_row, slot_lo, slot_hi, accumulate, factor_shift, ops)
        st = L.stream_ptr(out_b.device)
        R = self.bias_rows.shape[1]
        L.call("cg_weighted_colsum", L.ptr(self.bias_rows), L.ptr(factor_row) - 4 * factor_shift, slot_lo, slot_hi, R,
               L.ptr(out_b), 1 if accumulate else 0, st)

    def materialize(self, pass_idx: int, B: int) -> torch.Tensor:
        """[B, *weight.shape] per-sample weight gradients of one pass (rare path: reference
        train.py:233, 447 and tests)."""
        if self.impl is not None:
            return self.impl.materialize(pass_idx, B)
        slot0 = pass_idx * self.Bpad
        w = self.layer.weight
        out = torch.zeros((B,) + tuple(w.shape), device=w.device)
        st = L.stream_ptr(w.device)
        d = self._desc(self.X)
        if self.kind == "linear":
            # K = 1: a per-sample outer product, pure output bandwidth -> dedicated kernel
            L.call("cg_outer_rows", L.ptr(self.X), self.X.stride(0), L.ptr(self.Y), self.Y.stride(0),
                   self.M, self.Cn, slot0, B, L.ptr(out), st)
            return out
        d.group_mode, d.n_groups = L.GROUP_SAMPLE, B
        d.slot_lo, d.slot_hi, d.spg = slot0, slot0 + B, 1
        d.n_seg, d.seg_stride = 1, self.Bpad
        d.epi, d.out, d.out_group_stride = L.EPI_STORE, L.ptr(out), w.numel()
        L.call("cg_contract", C.byref(d), st)
        return out
