"""Per-sample row norms, L2 clipping and the gradient penalties built on them.

Host-side mirror of reference gradient_penalty.py (calc_penalty :4, calc_WGAN_GP_penalty :31,
calc_lipschitz_penalty_WRT :43) and backprop_clip.py:18-22 (l2_clip), with the per-sample
L2 norm over [B, C*H*W] -- the bandwidth-bound piece -- running in the CUDA kernels
`cg_row_l2_norm` / `cg_row_l2_norm_bwd` / `cg_l2_clip`.  The norm stays differentiable
(`create_graph=True` flows through `RowL2Norm`), including one more level for the
double-backward the immediate-sensitivity engine takes.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
from torch import autograd

from . import _lib as L


def _rows_cols(t: torch.Tensor):
    rows = t.shape[0]
    cols = t.numel() // max(rows, 1)
    return rows, cols


class _RowNormBwd(autograd.Function):
    """gin[i, :] = g[i, :] * (gout[i] / n[i]) as its own differentiable node."""

    @staticmethod
    def forward(ctx, g, norms, gout):
        g = L.require_cuda_f32(g, "row_l2_norm_bwd: g")
        norms = L.require_cuda_f32(norms, "row_l2_norm_bwd: norms")
        gout = L.require_cuda_f32(gout, "row_l2_norm_bwd: gout")
        rows, cols = _rows_cols(g)
        gin = torch.empty_like(g)
        L.call("cg_row_l2_norm_bwd", L.ptr(g), L.ptr(norms), L.ptr(gout), rows, cols, L.ptr(gin),
               L.stream_ptr(g.device))
        ctx.save_for_backward(g, norms, gout)
        return gin

    @staticmethod
    def backward(ctx, dgin):
        # second-order terms are tiny elementwise/row-reduction expressions; plain autograd ops
        g, norms, gout = ctx.saved_tensors
        B = g.shape[0]
        g2, d2 = g.reshape(B, -1), dgin.reshape(B, -1)
        inv = torch.where(norms > 0, 1.0 / norms, torch.zeros_like(norms))
        s = gout * inv
        dot = (g2 * d2).sum(dim=1)
        dg = (d2 * s[:, None]).view_as(g)
        dn = -dot * gout * inv * inv
        dgout = dot * inv
        return dg, dn, dgout


class RowL2Norm(autograd.Function):
    """norms[i] = ||t[i, ...]||_2 (reference gradient_penalty.py:52-53: view(B,-1).norm(2, dim=1))."""

    @staticmethod
    def forward(ctx, t):
        tc = L.require_cuda_f32(t, "row_l2_norm input")
        rows, cols = _rows_cols(tc)
        out = torch.empty(rows, device=tc.device, dtype=torch.float32)
        L.call("cg_row_l2_norm", L.ptr(tc), rows, cols, L.ptr(out), L.stream_ptr(tc.device))
        ctx.save_for_backward(tc, out)
        ctx.shape = t.shape
        return out

    @staticmethod
    def backward(ctx, gout):
        t, norms = ctx.saved_tensors
        return _RowNormBwd.apply(t, norms, gout.contiguous()).view(ctx.shape)


def row_l2_norm(t: torch.Tensor) -> torch.Tensor:
    return RowL2Norm.apply(t)


def vec_max(v: torch.Tensor) -> torch.Tensor:
    v = L.require_cuda_f32(v, "vec_max input")
    out = torch.empty(1, device=v.device)
    L.call("cg_vec_max", L.ptr(v), v.numel(), L.ptr(out), L.stream_ptr(v.device))
    return out


class _L2Clip(autograd.Function):
    @staticmethod
    def forward(ctx, t, C):
        tc = L.require_cuda_f32(t, "l2_clip input")
        rows, cols = _rows_cols(tc)
        out = torch.empty_like(tc)
        norms = torch.empty(rows, device=tc.device)
        L.call("cg_l2_clip", L.ptr(tc), rows, cols, float(C), L.ptr(out), L.ptr(norms), L.stream_ptr(tc.device))
        ctx.save_for_backward(tc, norms)
        ctx.C = float(C)
        return out.view_as(t)

    @staticmethod
    def backward(ctx, go):
        t, norms = ctx.saved_tensors
        B = t.shape[0]
        shape = (B,) + (1,) * (t.dim() - 1)
        n = norms.view(shape)
        clipped = n > ctx.C
        # d/dt [C t / ||t||] = C/||t|| (I - t t^T / ||t||^2)
        dot = (go * t).flatten(1).sum(dim=1).view(shape)
        g_clip = ctx.C * (go / n - t * dot / (n * n * n))
        return torch.where(clipped, g_clip, go), None


def l2_clip(t: torch.Tensor, C: float) -> torch.Tensor:
    """Per-sample L2 clip over all non-batch dims (reference backprop_clip.py:18-22):
    where(norm > C, C * (t / norm), t)."""
    return _L2Clip.apply(t, C)


# ----------------------------------------------------------------------------------------------
# gradient penalties (same signatures as reference gradient_penalty.py)
# ----------------------------------------------------------------------------------------------
def calc_lipschitz_penalty_WRT(model, inputs, input_labels=None, device=None, per_sample=False,
                               one_sided=False, aux_penalty=True):
    inputs = inputs.detach()
    input_labels = None if input_labels is None else input_labels.detach()
    inputs.requires_grad_(True)
    out, aux_out = model(inputs, input_labels)
    grads = autograd.grad(outputs=out, inputs=inputs, grad_outputs=torch.ones_like(out),
                          create_graph=True, retain_graph=True, only_inputs=True)[0]
    norms = row_l2_norm(grads)
    pen = (norms - 1).clamp(min=0) ** 2 if one_sided else (norms - 1) ** 2
    if aux_penalty and aux_out is not None:
        for i in range(aux_out.size(1)):
            ga = autograd.grad(outputs=aux_out[:, i], inputs=inputs, grad_outputs=torch.ones_like(aux_out[:, i]),
                               create_graph=True, retain_graph=True, only_inputs=True)[0]
            na = row_l2_norm(ga)
            pen = pen + ((na - 1).clamp(min=0) ** 2 if one_sided else (na - 1) ** 2)
    return pen if per_sample else pen.mean()


def calc_WGAN_GP_penalty(model, real_data, real_labels, fake_data, fake_labels, device=None, per_sample=False,
                         one_sided=False, weight=10.0, aux_penalty=False, alpha: Optional[torch.Tensor] = None):
    """`alpha` [B,1] may be supplied; by default it is drawn on the data's device (the reference
    draws it with torch.rand on the CPU and copies it over, gradient_penalty.py:33-36)."""
    B = real_data.size(0)
    if alpha is None:
        alpha = torch.rand(B, 1, device=real_data.device)
    a = alpha.to(real_data.device).view(B, *([1] * (real_data.dim() - 1)))
    inter = a * real_data + (1 - a) * fake_data
    return weight * calc_lipschitz_penalty_WRT(model, inter, real_labels, device=device, per_sample=per_sample,
                                               one_sided=one_sided, aux_penalty=aux_penalty)


def calc_penalty(model, penalty_types: Sequence[str], real_data, real_labels, fake_data, fake_labels, device=None,
                 per_sample=False, weights=None, aux_penalty=False, alpha: Optional[torch.Tensor] = None):
    penalty = 0
    weights = [1 / len(penalty_types) for _ in penalty_types] if weights is None else weights
    for w, ptype in zip(weights, penalty_types):
        if ptype.startswith("WGAN-GP"):
            p = calc_WGAN_GP_penalty(model, real_data, real_labels, fake_data, fake_labels, device=device,
                                     per_sample=per_sample, one_sided=ptype[-1] == "1", aux_penalty=aux_penalty,
                                     alpha=alpha)
        elif ptype.startswith("DRAGAN"):
            # reference gradient_penalty.py:20-29 is broken (random_(0,1) is identically 0 and the
            # expand() shape does not match); SURVEY.md §3.4 marks it "skip".
            raise NotImplementedError("DRAGAN penalty is broken in the reference and not reproduced")
        else:
            raise Exception("Unknown penalty type: " + ptype)
        penalty = penalty + w * p
    return penalty
