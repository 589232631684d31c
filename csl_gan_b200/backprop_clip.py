"""Backprop clipping: per-sample L2 clipping of every layer's input (forward) and grad-input (backward),
which bounds each parameter's per-sample gradient norm a priori.

Host-side mirror of reference backprop_clip.py (`PGCWrapper` :49-103, `BackpropClipper` :105-158; wired in
at train.py:84-92), which the reference itself calls "experimental and not finished" (options.py:243-244).
The clipping arithmetic is the CUDA `l2_clip` kernel (`functional.l2_clip` -> `cg_l2_clip`, reference
backprop_clip.py:18-22).  Differences from the reference, on purpose: layer shapes come from one dry
forward pass with hooks instead of `torchinfo.summary` on a hard-coded (1,1,28,28) input (:123), so any
critic / input size works.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
from torch import nn

from .functional import l2_clip


def _prod(shape) -> int:
    out = 1
    for s in shape:
        out *= int(s)
    return out


def l2_size(n: int, scale: float) -> float:
    """L2 norm of an n-element tensor whose entries all equal `scale` (reference :14-16)."""
    return math.sqrt(n * scale ** 2)


def l2_to_l1(l2: float, n: int) -> float:
    return math.sqrt(n) * l2


class _Identity(nn.Module):
    def forward(self, x):
        return x


class ClippedLayer(nn.Module):
    """forward: layer(l2_clip(x, C_in)); backward: grad wrt the layer output's consumer is clipped to C_back
    (reference PGCWrapper :98-103)."""

    def __init__(self, clipper: "BackpropClipper", layer: nn.Module, input_clip: float, back_clip: float):
        super().__init__()
        self.module = layer
        self.clipper = clipper
        self.input_clip_param = float(input_clip)
        self.back_clip_param = float(back_clip)
        self.dummy = _Identity()
        self.dummy.register_full_backward_hook(self._clip_grad_input)

    def _clip_grad_input(self, module, grad_input, grad_output):
        if self.clipper.hooks_enabled:
            return tuple(None if g is None else l2_clip(g, self.back_clip_param) for g in grad_input)

    def forward(self, x):
        return self.dummy(self.module(l2_clip(x, self.input_clip_param)))


class BackpropClipper:
    """Derives the per-layer clip parameters and the per-parameter gradient L2 bounds
    (`grad_l2_bounds`, which train.py:89 turns into `clipping_param_per_layer`), then wraps every leaf
    layer that owns parameters."""

    def __init__(self, model: nn.Module, back_clip_params: Optional[Sequence[float]] = None,
                 input_clip_params: Optional[Sequence[float]] = None, auto_activation_scale: float = 0.5,
                 auto_weight_grad_scale: float = 1e-4, device="cpu", input_size=(1, 1, 28, 28), wrap: bool = True):
        self.hooks_enabled = True
        self.device = device
        self.auto_activation_scale = auto_activation_scale
        self.auto_weight_grad_scale = auto_weight_grad_scale
        auto = back_clip_params is None or input_clip_params is None
        self.back_clip_params: List[float] = [] if back_clip_params is None else list(back_clip_params)
        self.input_clip_params: List[float] = [] if input_clip_params is None else list(input_clip_params)
        self.grad_l2_bounds: List[float] = []

        layers = [(n, m) for n, m in model.named_modules()
                  if len(list(m.children())) == 0 and any(p.requires_grad for p in m.parameters())]
        shapes = self._trace_shapes(model, [m for _, m in layers], input_size)
        plan = []
        for li, (name, m) in enumerate(layers):
            in_shape, out_shape = shapes[m]
            p = list(m.parameters())
            n_in, n_out_sp = _prod(in_shape), _prod(out_shape[1:])
            if auto:
                c_in = l2_size(n_in, auto_activation_scale)
                w_bound = l2_size(p[0].numel(), auto_weight_grad_scale)
                if isinstance(m, nn.Linear):
                    c_back = w_bound / c_in
                    bounds = [w_bound] + ([c_back] if len(p) > 1 else [])
                elif isinstance(m, nn.Conv2d):
                    c_back = l2_to_l1(w_bound, n_out_sp) / c_in
                    bounds = [w_bound] + ([c_back * n_out_sp] if len(p) > 1 else [])
                else:
                    raise NotImplementedError(f"backprop clipping of {type(m).__name__}")
                self.input_clip_params.append(c_in)
                self.back_clip_params.append(c_back)
            else:
                c_in, c_back = self.input_clip_params[li], self.back_clip_params[li]
                if isinstance(m, nn.Linear):
                    bounds = [c_in * c_back] + ([c_back] if len(p) > 1 else [])
                elif isinstance(m, nn.Conv2d):
                    bounds = [c_in * l2_to_l1(c_back, n_out_sp)] + ([c_back * n_out_sp] if len(p) > 1 else [])
                else:
                    raise NotImplementedError(f"backprop clipping of {type(m).__name__}")
            self.grad_l2_bounds.extend(bounds)
            plan.append((name, m, c_in, c_back))
        if wrap:
            for name, m, c_in, c_back in plan:
                parent = model
                parts = name.split(".")
                for part in parts[:-1]:
                    parent = parent._modules[part]
                parent._modules[parts[-1]] = ClippedLayer(self, m, c_in, c_back).to(device)

    @staticmethod
    def _trace_shapes(model, leaves, input_size):
        shapes, handles = {}, []
        for m in leaves:
            handles.append(m.register_forward_hook(
                lambda mod, inp, out: shapes.__setitem__(mod, (tuple(inp[0].shape[1:]), tuple(out.shape[1:])))))
        was_training = model.training
        dev = next(model.parameters()).device
        with torch.no_grad():
            model(torch.zeros(input_size, device=dev))
        model.train(was_training)
        for h in handles:
            h.remove()
        return shapes

    def enable_hooks(self):
        self.hooks_enabled = True

    def disable_hooks(self):
        self.hooks_enabled = False
