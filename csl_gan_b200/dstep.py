"""The differentially-private discriminator update, `train_D` re-authored around the CUDA engines.

Mirrors reference train.py:360-500 (`train_D`), :95-138 (`setup_privacy_engine`), :204-245
(`update_adaptive_clipping_params`), :247-249 (`update_sens_moving_avg`) and :310-338 (grad / IS
logging), with the host synchronisations removed: adaptive thresholds, norm statistics and
sensitivities stay on the device; logging values are returned as device tensors and only converted
when the caller asks (`DStepResult.to_host()`).

Out of scope here: the generator, datasets, MeanSampler, file logging.  The caller
passes the fake batch (G's detached output) and, where a penalty or adaptive clipping needs it, a
public batch.
"""
from __future__ import annotations

import os

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch
from torch import autograd, nn

from .functional import calc_penalty
from .is_engine import ISPrivacyEngine
from .privacy_engine import PrivacyEngine, calc_sample_norms

ALPHAS = [1 + x / 10.0 for x in range(1, 100)] + list(range(12, 400))      # reference train.py:99


def setup_privacy_engine(opt, D: nn.Module, d_optimizer, **engine_kw):
    """reference train.py:95-138."""
    common = dict(batch_size=opt.batch_size, sample_size=opt.train_set_size, alphas=ALPHAS,
                  noise_multiplier=opt.sigma)
    if opt.dp_mode == "is":
        eng = ISPrivacyEngine(
            D, **common, per_param=opt.imm_sens_per_param,
            scaling_vec=None if opt.imm_sens_scaling_mode == "standard" else opt.imm_sens_scaling_vec, **engine_kw)
    elif opt.dp_mode == "gc":
        n_params = len(list(D.parameters()))
        if opt.clipping_param_per_layer is None:
            opt.clipping_param_per_layer = [1 for _ in range(n_params)]
        per_layer = opt.grad_clip_mode[-3:] == "-pl"
        eng = PrivacyEngine(
            D, **common, accum_passes=not opt.grad_clip_split,
            num_private_passes=1 if opt.grad_clip_split else None, auto_clip_and_accum_on_step=False,
            max_grad_norm=opt.clipping_param_per_layer if per_layer else opt.clipping_param, **engine_kw)
        eng.disable_hooks()
    else:
        raise NotImplementedError(f"dp_mode {opt.dp_mode!r}")
    eng.attach(d_optimizer)
    eng._set_seed(opt.manual_seed)
    return eng


@dataclass
class DStepResult:
    d_real_loss: torch.Tensor
    d_fake_loss: torch.Tensor
    d_real: torch.Tensor
    d_fake: torch.Tensor
    penalty: Optional[torch.Tensor] = None
    d_real_aux_loss: Optional[torch.Tensor] = None
    d_real_aux: Optional[torch.Tensor] = None
    stats: Dict[str, torch.Tensor] = field(default_factory=dict)

    def to_host(self) -> Dict[str, object]:
        """One synchronisation for everything the reference logs per step (train.py:487-500)."""
        out = {"D Real Loss": self.d_real_loss.item(), "D Fake Loss": self.d_fake_loss.item(),
               "D Real Acc": 100 * float((self.d_real > 0).float().mean()),
               "D Fake Acc": 100 * float((self.d_fake < 0).float().mean())}
        out["D Adv Loss"] = out["D Real Loss"] + out["D Fake Loss"]
        if self.penalty is not None:
            out["D Penalty"] = float(self.penalty)
        if self.d_real_aux_loss is not None:
            out["D Real Aux Loss"] = float(self.d_real_aux_loss)
        for k, v in self.stats.items():
            out[k] = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v
        return out


class DiscriminatorStep:
    def __init__(self, opt, D: nn.Module, d_optimizer, privacy_engine=None,
                 public_batch: Optional[Callable[[int, Optional[torch.Tensor]], Tuple[torch.Tensor, Optional[torch.Tensor]]]] = None,
                 collect_stats: bool = False, skip_weight_grads: bool = True):
        self.opt, self.D, self.opt_d, self.engine = opt, D, d_optimizer, privacy_engine
        self.public_batch = public_batch
        self.collect_stats = collect_stats
        self.skip_weight_grads = skip_weight_grads
        # a channels_last critic converts every NCHW batch it is given (and the capture kernels would convert it
        # once more): do it once, up front
        self._nhwc = any(p.dim() == 4 and not p.is_contiguous() and p.is_contiguous(memory_format=torch.channels_last)
                         for p in D.parameters())
        # ... unless the batch is thin (images: 1-4 channels): cuDNN's own conversion of such a batch costs the same as
        # ours, and the thin-layer kernel (csrc/thin.cuh) copies planar image rows with 8-byte cp.async but has to
        # gather a channels-last 3-channel image element by element.  CSLGAN_INPUT_NHWC=1 forces the conversion.
        self._nhwc_min_channels = 1 if os.environ.get("CSLGAN_INPUT_NHWC", "0") == "1" else 16

    # ------------------------------------------------------------------ losses (train.py:342-358)
    def _fake_loss(self, fake_img, y):
        opt, D = self.opt, self.D
        d_fake, d_fake_aux = D(fake_img, y, aux=opt.d_fake_aux_loss)
        loss = D.fake_loss(d_fake, fake_img.device)
        aux = D.aux_loss(d_fake_aux, y, fake_img.device, fake=True) if opt.use_aux_loss and opt.d_fake_aux_loss else 0
        return d_fake, loss, aux

    def _real_loss(self, img, labels):
        opt, D = self.opt, self.D
        d_real, d_real_aux = D(img, labels)
        loss = D.real_loss(d_real, img.device)
        aux = D.aux_loss(d_real_aux, labels, img.device, fake=False) if opt.use_aux_loss else 0
        return d_real, d_real_aux, loss, aux

    # ------------------------------------------------------------------ adaptive clipping (train.py:204-245)
    def update_adaptive_clipping_params(self, fake_img, fake_y):
        opt, eng = self.opt, self.engine
        if self.public_batch is None:
            raise RuntimeError("adaptive clipping needs a public batch source (mean samples or a public partition)")
        for p in self.D.parameters():
            p.grad = None
        img, labels = self.public_batch(opt.batch_size, None)
        if self._nhwc and img.dim() == 4:
            img = img.contiguous(memory_format=torch.channels_last)
        loss = 0
        if not opt.grad_clip_split:
            _, fl, fa = self._fake_loss(fake_img, fake_y)
            loss = fl + fa
        _, _, rl, ra = self._real_loss(img, labels)
        (loss + rl + ra).backward()
        th = eng.adaptive_thresholds(opt.adaptive_stat, 1.0, pass_idx=0)          # device tensor, no sync
        if opt.use_grad_clip_per_layer:
            eng.set_max_grad_norm(th * opt.adaptive_scalar)
        else:
            eng.set_max_grad_norm(th.norm(2) * opt.adaptive_scalar)
        self.opt_d.zero_grad()

    def update_sens_moving_avg(self):
        """train.py:247-249 (beta defaults to 0.9: the reference never defines opt.moving_avg_beta)."""
        eng, beta = self.engine, self.opt.moving_avg_beta
        norms = torch.stack([p.grad.reshape(-1).norm(2) for p in self.D.parameters()]).cpu().tolist()
        eng.set_scaling_vec([v * beta + n * (1 - beta) for v, n in zip(eng.scaling_vec, norms)])

    # ------------------------------------------------------------------ logging (train.py:310-338), device side
    def _grad_stats(self) -> Dict[str, torch.Tensor]:
        eng, opt = self.engine, self.opt
        all_norms = calc_sample_norms(named_params=eng.clipper._named_grad_samples(),
                                      flat=not eng.clipper.norm_clipper.is_per_layer)
        ps = 1 if opt.grad_clip_split else 0
        norms = torch.stack(all_norms)[:, min(ps, all_norms[0].shape[0] - 1)]
        factors = iter(eng.clipper.norm_clipper.calc_clipping_factors(all_norms))
        clipped = torch.stack([(next(factors)[min(ps, all_norms[0].shape[0] - 1)] < 0.999).float().mean()
                               for _ in range(len(all_norms))])
        return {"D Layer Grad Norm Means": norms.mean(dim=1), "D Layer Grad Norm Stds": norms.std(dim=1, unbiased=False),
                "D Layer Grad Norm Maxes": norms.max(dim=1).values, "Grads Clipped": clipped}

    # ------------------------------------------------------------------ the step (train.py:360-500)
    def __call__(self, img, labels, fake_img, fake_y, use_dp: bool = True) -> DStepResult:
        opt, D, eng = self.opt, self.D, self.engine
        for p in D.parameters():
            p.grad = None
        batch_size = img.size(0)
        use_gc = opt.dp_mode == "gc" and use_dp
        use_is = opt.dp_mode == "is" and use_dp
        fake_img = fake_img.detach()
        if self._nhwc and img.dim() == 4 and img.shape[1] >= self._nhwc_min_channels:
            img = img.contiguous(memory_format=torch.channels_last)
            fake_img = fake_img.contiguous(memory_format=torch.channels_last)

        if opt.per_sample_grad and use_dp:
            if use_gc and self.skip_weight_grads:
                eng.enable_hooks(freeze_weights=True)     # backward = dgrad chain only (privacy_engine.enable_hooks)
            else:
                eng.enable_hooks()
        if use_is:
            img = img.detach().requires_grad_(True)
        try:
            if use_gc and opt.grad_clip_mode[:8] == "adaptive":
                self.update_adaptive_clipping_params(fake_img, fake_y)

            d_fake, d_fake_loss, d_fake_aux_loss = self._fake_loss(fake_img, fake_y)
            d_real, d_real_aux, d_real_loss, d_real_aux_loss = self._real_loss(img, labels)
            d_loss = d_real_loss + d_fake_loss + d_real_aux_loss + d_fake_aux_loss
            res = DStepResult(d_real_loss.detach(), d_fake_loss.detach(), d_real.detach(), d_fake.detach())
            if opt.use_aux_loss:
                res.d_real_aux_loss, res.d_real_aux = d_real_aux_loss.detach(), d_real_aux.detach()

            if opt.per_sample_grad and use_dp:
                # grad_outputs of every captured layer, without the (unused) batch-summed weight gradients
                eng.backward(d_loss) if self.skip_weight_grads else d_loss.backward()
        finally:
            # also on an exception: the frozen-weights mode must hand the parameters back to autograd
            if opt.per_sample_grad and use_dp:
                eng.disable_hooks()
        if use_gc:
            if self.collect_stats:
                with torch.no_grad():
                    res.stats.update(self._grad_stats())
            if getattr(eng, "overlap_allreduce", False) and len(opt.penalty) > 0:
                # the penalty gradient is added to p.summed_grad AFTER accumulate_batch(): the allreduce must see it
                eng.overlap_allreduce = False
            eng.clip()
            if opt.grad_clip_split:
                eng.accum_grads_across_passes()

        has_penalty = len(opt.penalty) > 0
        if has_penalty:
            pen_real, pen_labels = img, labels
            if opt.penalty_use_public_data and self.public_batch is not None:
                pen_real, pen_labels = self.public_batch(batch_size, labels)
                if self._nhwc and pen_real.dim() == 4:
                    pen_real = pen_real.contiguous(memory_format=torch.channels_last)
            if use_dp and opt.per_sample_grad and not opt.penalty_use_public_data:
                # train.py:434-450: the penalty touches private data, so its per-sample gradients are added to the
                # per-sample gradients before (re-)clipping.  The slow path of the reference, kept as such: one
                # autograd.grad per sample into the materialised p.grad_sample[0, i] (without the reference's
                # create_graph=True, which is what leaks there, train.py:435-436)
                eng.expose_grad_sample_attrs()
                penalties = calc_penalty(D, opt.penalty, pen_real, pen_labels, fake_img, fake_y, device=img.device,
                                         per_sample=True, aux_penalty=opt.aux_penalty)
                penalty = penalties.mean(dim=0)
                params = list(D.parameters())
                for i in range(len(penalties)):
                    pg = autograd.grad(penalties[i], params, retain_graph=True, allow_unused=True)
                    with torch.no_grad():
                        for p, g in zip(params, pg):
                            if g is not None:
                                p.grad_sample[0, i] += g
                eng.clip()
                eng.accumulate_batch()
            elif use_dp and opt.per_sample_grad:
                eng.accumulate_batch()
                penalty = calc_penalty(D, opt.penalty, pen_real, pen_labels, fake_img, fake_y, device=img.device,
                                       aux_penalty=opt.aux_penalty)
                pgrad = autograd.grad(penalty, list(D.parameters()), allow_unused=True)
                with torch.no_grad():
                    for p, g in zip(D.parameters(), pgrad):
                        if g is not None:
                            p.summed_grad.add_(g, alpha=float(opt.batch_size))     # summed_grad is a sum (train.py:431)
            else:
                penalty = calc_penalty(D, opt.penalty, pen_real, pen_labels, fake_img, fake_y, device=img.device,
                                       aux_penalty=opt.aux_penalty)
                d_loss = d_loss + penalty
                if use_is:
                    eng.backward(d_loss, img)
                    if opt.imm_sens_scaling_mode == "moving-avg-pl":
                        self.update_sens_moving_avg()
                else:
                    d_loss.backward()
            res.penalty = penalty.detach()
        else:
            if use_gc:
                eng.accumulate_batch()
            elif use_is:
                eng.backward(d_loss, img)
                if opt.imm_sens_scaling_mode == "moving-avg-pl":
                    self.update_sens_moving_avg()
            else:
                d_loss.backward()
        if use_is and self.collect_stats:
            res.stats["IS"] = eng._sens_dev.detach().clone()

        self.opt_d.step()
        return res


class GraphedDiscriminatorStep:
    """The whole D step -- critic forward/backward, capture, norms, clip, (allreduce), noise, optimizer --
    captured once into a CUDA graph and replayed.  At MNIST sizes the step is a few
    hundred microseconds of GPU work behind ~100 kernel launches, so launch overhead, not the kernels,
    sets the speed; one graph launch per step removes it.

    Requirements: fixed batch shapes; an optimizer constructed with `capturable=True`; no host reads
    inside the step (the step keeps every statistic on the device; call `result.to_host()` afterwards).
    The engines switch to a device-resident Philox offset so every replay draws fresh noise.
    """

    def __init__(self, stepper: DiscriminatorStep, example_inputs, warmup: int = 3):
        self.stepper = stepper
        eng = stepper.engine
        if eng is not None:
            eng.enable_graph_safe_rng()
        img, labels, fake, fake_y = example_inputs
        self.s_img, self.s_fake = img.clone(), fake.clone()
        self.s_labels = None if labels is None else labels.clone()
        self.s_fake_y = None if fake_y is None else fake_y.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):                      # allocate plans / cuDNN workspaces outside the capture
                stepper(self.s_img, self.s_labels, self.s_fake, self.s_fake_y, use_dp=True)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        steps_before = eng.steps if eng is not None else 0
        with torch.cuda.graph(self.graph):
            self.result = stepper(self.s_img, self.s_labels, self.s_fake, self.s_fake_y, use_dp=True)
        self._engine = eng
        self.warmup_steps = warmup
        if eng is not None:
            eng.steps = steps_before                     # the capture itself executes nothing

    def __call__(self, img, labels, fake, fake_y) -> DStepResult:
        self.s_img.copy_(img, non_blocking=True)
        self.s_fake.copy_(fake, non_blocking=True)
        if self.s_labels is not None:
            self.s_labels.copy_(labels, non_blocking=True)
        if self.s_fake_y is not None and fake_y is not None:
            self.s_fake_y.copy_(fake_y, non_blocking=True)
        self.graph.replay()
        if self._engine is not None:
            self._engine.steps += 1
        return self.result
