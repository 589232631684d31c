"""PrivacyEngine: the gc (per-sample gradient clipping) engine behind `--dp_mode gc`.

Drop-in for the twosixlabs/opacus fork's `PrivacyEngine` as reference train.py drives it
(constructor train.py:110-116; attach/_set_seed :135-136; enable/disable_hooks :117, 373, 389;
clip :399; accum_grads_across_passes :402; accumulate_batch :417/450/467; set_max_grad_norm
:241-243; clipper.* and calc_sample_norms :311-328; patched optimizer.step :484; get_privacy_spent
:295, 588).  The arithmetic runs in hand-written sm_100a kernels behind the C ABI
(include/cslgan_b200.h); this module is host-side bookkeeping only and has no fallback path.

Pipeline per D step (all on the current CUDA stream, no host synchronisation):
  forward hooks   stage activations (TF32, K-major / kw-plane layout)
  tensor hooks    stage B * grad_output, per-sample bias gradients, Linear ||b||^2
  clip()          per-sample squared norms  (tcgen05 contraction with a sum-of-squares epilogue;
                  closed form ||a||^2 ||b||^2 for Linear)  ->  clip factors  ->  clip-factor-scaled
                  backprops  ->  ONE split-K tcgen05 GEMM per layer over all samples and passes
  accumulate_batch()  p.summed_grad (+)= clipped sums
  optimizer.step()    p.grad = summed/B + Philox N(0, (sigma C_k)^2)/B   (bit-compatible with torch)

Fork semantics that cannot be verified here are explicit switches: see
`split_clip_fake`, `noise_on`, and DESIGN.md.
"""
from __future__ import annotations

import os

import math
import types
from itertools import cycle
from typing import Dict, Iterable, List, Optional, Sequence, Tuple, Union

import ctypes as C

import torch
from torch import nn

from . import _lib as L
from .accountant import get_privacy_spent as _rdp_to_eps, compute_rdp
from .grad_sample import LayerPlan, _round_up

SUPPORTED_LAYERS = (nn.Linear, nn.Conv2d, nn.ConvTranspose2d)
_UNSUPPORTED_WITH_PARAMS = (nn.modules.batchnorm._BatchNorm,)


class GradSampleView:
    """Lazy stand-in for `p.grad_sample` ([n_passes, B, *p.shape], reference train.py:233, 388, 447).

    The tensor is never materialised for the accesses the reference makes on the hot path:
      p.grad_sample.size(1)                                    -> batch size
      p.grad_sample[k].view(B, -1).norm(2, dim=1)              -> fused per-sample norms
    Anything else (`.materialize()`, `torch.as_tensor`-style use) runs the store epilogue.
    """

    def __init__(self, engine: "PrivacyEngine", p_idx: int):
        self._e, self._k = engine, p_idx

    @property
    def shape(self):
        e = self._e
        return torch.Size((e._n_passes_view(), e._cur_B) + tuple(e._params[self._k].shape))

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def dim(self):
        return len(self.shape)

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, idx):
        if isinstance(idx, int):
            return _PassView(self, idx)
        # anything finer than a whole pass (reference train.py:447: `p.grad_sample[0, i] += penalty_grad`) needs the
        # real tensor: the engine materialises every parameter's per-sample gradients once, hands out views into
        # them, and the next clip() reads norms and sums from those tensors (the slow path, as in the reference)
        return self._e._materialized_for_write()[self._k][idx]

    def __setitem__(self, idx, value):
        t = self._e._materialized_for_write()[self._k]
        dst = t[idx]
        if not (isinstance(value, torch.Tensor) and value.data_ptr() == dst.data_ptr() and value.shape == dst.shape):
            t[idx] = value

    def materialize(self) -> torch.Tensor:
        if self._e._mat is not None:
            return self._e._mat[self._k]
        return self._e.materialize_grad_sample(self._k)

    def __torch_function__(self, func, types_, args=(), kwargs=None):  # pragma: no cover - convenience
        args = tuple(a.materialize() if isinstance(a, GradSampleView) else a for a in args)
        return func(*args, **(kwargs or {}))


class _PassView:
    def __init__(self, gs: GradSampleView, pass_idx: int):
        self._gs, self._p = gs, pass_idx

    def view(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        if len(shape) == 2 and shape[1] == -1:
            return _FlatView(self._gs, self._p)
        return self.materialize().view(*shape)

    reshape = view

    def materialize(self):
        return self._gs.materialize()[self._p]


class _FlatView:
    def __init__(self, gs: GradSampleView, pass_idx: int):
        self._gs, self._p = gs, pass_idx

    def norm(self, p=2, dim=1, **kw):
        if p != 2 or dim not in (1, -1):
            return self._gs.materialize()[self._p].flatten(1).norm(p, dim=dim, **kw)
        e = self._gs._e
        return e.per_sample_norms()[self._gs._k, self._p]


class _NormClipperShim:
    """`privacy_engine.clipper.norm_clipper` (reference train.py:313, 324)."""

    def __init__(self, engine: "PrivacyEngine"):
        self._e = engine

    @property
    def is_per_layer(self) -> bool:
        return self._e.is_per_layer

    @property
    def thresholds(self) -> torch.Tensor:
        return self._e._thresholds_dev.clone()

    def calc_clipping_factors(self, norms=None):
        """Iterable of [n_passes, B] factors, one per parameter (flat: the same vector cycled,
        like upstream ConstantFlatClipper)."""
        f = self._e.clipping_factors()
        if self._e.is_per_layer:
            return [f[k] for k in range(f.shape[0])]
        return cycle([f[0]])


class _ClipperShim:
    """`privacy_engine.clipper` (reference train.py:312-313)."""

    def __init__(self, engine: "PrivacyEngine"):
        self._e = engine
        self.norm_clipper = _NormClipperShim(engine)

    def _named_grad_samples(self):
        e = self._e
        return [(n, GradSampleView(e, k)) for k, n in enumerate(e._param_names)]

    def _named_params(self):
        return list(zip(self._e._param_names, self._e._params))


def calc_sample_norms(named_params, flat: bool = True) -> List[torch.Tensor]:
    """`opacus.utils.tensor_utils.calc_sample_norms` (reference train.py:311-314) with the fork's
    pass dimension: items are [n_passes, B]; flat -> one item with the norm across parameters."""
    named_params = list(named_params)
    views = [v for _, v in named_params]
    if views and all(isinstance(v, GradSampleView) for v in views):
        e = views[0]._e
        per = e.per_sample_norms()                      # [n_params, n_passes, B]
        ks = [v._k for v in views]
        if flat:
            if ks == list(range(per.shape[0])):
                return [e.flat_sample_norms()]
            return [per[ks].norm(2, dim=0)]
        return [per[k] for k in ks]
    norms = [p.reshape(p.shape[0], p.shape[1], -1).norm(2, dim=-1) for _, p in named_params]
    if flat:
        norms = [torch.stack(norms, dim=0).norm(2, dim=0)]
    return norms


class PrivacyEngine:
    def __init__(self, module: nn.Module, *, batch_size: int, sample_size: int,
                 alphas: Sequence[float] = tuple([1 + x / 10.0 for x in range(1, 100)] + list(range(12, 64))),
                 noise_multiplier: float, max_grad_norm: Union[float, Sequence[float], torch.Tensor],
                 accum_passes: bool = False, num_private_passes: Optional[int] = None,
                 auto_clip_and_accum_on_step: bool = True, loss_reduction: str = "mean",
                 split_clip_fake: bool = True, max_passes: int = 2, process_group=None,
                 data_parallel: bool = False, global_batch_size: Optional[int] = None,
                 per_layer_noise: str = "l2norm", clip_margin: float = 0.0,
                 operand_dtype: Optional[str] = None, overlap_allreduce: bool = False, nccl_reserve_sms: int = 8,
                 fused_allreduce: bool = True, **misc):
        """`batch_size` is THIS rank's batch.  Under data parallelism the accountant needs the global sampling
        rate: pass `global_batch_size`, or leave it None and the constructor sums the per-rank batch sizes with
        one tiny allreduce.
        `per_layer_noise`: with per-layer thresholds C_k the L2 sensitivity of the clipped sum is ||C||_2, so
        every parameter is noised with std sigma*||C||_2 ("l2norm", what upstream opacus >= 0.10 does and what
        `get_privacy_spent`, which accounts for noise_multiplier = sigma, assumes); "own" noises parameter k
        with sigma*C_k (weaker privacy than the accountant reports; kept as an explicit switch because the
        fork's choice cannot be verified).
        `clip_margin`: factor = min(1, C*(1 - margin)/(norm + 1e-6)).  The norms and the clipped sum come from
        TF32-rounded tensor-core operands (relative error <= 1e-3), so a clipped per-sample gradient can
        exceed C by that much; a margin of 2**-9 makes the bound strict.  Default 0 = the reference's formula.
        `operand_dtype`: "f16" (default) stages the tensor-core operands of the channels-last path as FP16 with an
        exact per-sample power-of-two scale (TF32's 10-bit mantissa, half the bytes, twice the MMA rate); "tf32"
        keeps fp32 words.  Joint clipping (accum_passes=True) sums two passes in one accumulator and therefore
        always uses TF32 (the passes' staging scales differ).
        `fused_allreduce` (data parallel, default on): where the ranks share an NVSwitch multicast mapping
        (torch.distributed._symmetric_memory), the clipped sums are written into symmetric memory and step() does the
        cross-rank sum, the noise and the broadcast of the finished gradient in ONE kernel (multimem.ld_reduce /
        multimem.st); every rank generates only 1/world of the normals.  Falls back to the NCCL allreduce when the
        mapping is unavailable, a parameter is not a view of the flat buffer, or `overlap_allreduce` is on.
        `overlap_allreduce` (data parallel): clip() walks the layers from the last to the first and all-reduces
        every finished bucket of the flat clipped-sum buffer asynchronously (NCCL stream), so the collective of the
        large late layers runs under the contractions of the early ones; the persistent GEMM kernels launched
        meanwhile leave `nccl_reserve_sms` SMs to NCCL.  step() only waits.  Because the reduction has then already
        happened, anything the caller adds to p.summed_grad after accumulate_batch() must be identical on every
        rank (DiscriminatorStep switches the overlap off for steps that add a penalty gradient)."""
        if loss_reduction not in ("mean", "sum"):
            raise ValueError("loss_reduction must be 'mean' or 'sum'")
        if per_layer_noise not in ("l2norm", "own"):
            raise ValueError("per_layer_noise must be 'l2norm' or 'own'")
        if not 0.0 <= clip_margin < 1.0:
            raise ValueError("clip_margin must be in [0, 1)")
        L.load()                                          # fail loudly, now, if the CUDA library is absent
        self.module = module
        self.batch_size = batch_size
        self.sample_size = sample_size
        self.per_layer_noise = per_layer_noise
        self.clip_margin = float(clip_margin)
        self.operand_dtype = operand_dtype or L.default_operand_dtype()
        if self.operand_dtype not in ("f16", "tf32"):
            raise ValueError("operand_dtype must be 'f16' or 'tf32'")
        if accum_passes:
            self.operand_dtype = "tf32"
        self.overlap_allreduce = bool(overlap_allreduce)
        self.fused_allreduce = bool(fused_allreduce)
        self.batch_small_ops = os.environ.get("CSLGAN_BATCH_SMALL", "1") != "0"
        self._symm = None
        self._symm_failed = False
        self._symm_last = 1
        self.nccl_reserve_sms = int(nccl_reserve_sms)
        self._reduce_works: list = []
        self._reduced_early = False
        self.alphas = list(alphas)
        self.noise_multiplier = float(noise_multiplier)
        self.accum_passes = accum_passes
        self.num_private_passes = num_private_passes
        self.auto_clip_and_accum_on_step = auto_clip_and_accum_on_step
        self.loss_reduction = loss_reduction
        self.split_clip_fake = split_clip_fake
        self.max_passes = max_passes
        self.process_group = process_group
        self.data_parallel = data_parallel or process_group is not None
        self.misc_settings = misc
        self.steps = 0
        self.hooks_enabled = True
        self.global_batch_size = global_batch_size
        self._frozen: list = []
        self._leaf_outputs: List[torch.Tensor] = []
        self.optimizer = None
        self.validate()

        self._params: List[nn.Parameter] = [p for p in module.parameters() if p.requires_grad]
        names = {id(p): n for n, p in module.named_parameters()}
        self._param_names = [names[id(p)] for p in self._params]
        pidx = {id(p): k for k, p in enumerate(self._params)}
        self.device = self._params[0].device
        if self.device.type != "cuda":
            raise L.CslGanCudaError(
                f"PrivacyEngine needs the module on a CUDA device (found {self.device}); there is no CPU path")
        if self.global_batch_size is None:
            self.global_batch_size = batch_size
            if self.data_parallel:
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized():
                    from .dist import global_batch_size as _gbs
                    self.global_batch_size = _gbs(batch_size, self.device, self.process_group)
        self._plans: List[LayerPlan] = []
        covered = set()
        for name, layer in module.named_modules():
            if isinstance(layer, SUPPORTED_LAYERS):
                own = [p for p in layer.parameters(recurse=False) if p.requires_grad]
                if not own:
                    continue
                w = pidx[id(layer.weight)]
                b = pidx[id(layer.bias)] if getattr(layer, "bias", None) is not None and layer.bias.requires_grad else None
                self._plans.append(LayerPlan(name, layer, w, b))
                self._plans[-1].use_half = self.operand_dtype == "f16"
                covered.update(id(p) for p in own)
        missing = [n for n, p in zip(self._param_names, self._params) if id(p) not in covered]
        if missing:
            raise NotImplementedError(f"parameters outside Linear/Conv2d/ConvTranspose2d layers: {missing}")
        self._handles = []
        for plan in self._plans:
            self._handles.append(plan.layer.register_forward_hook(self._make_fwd_hook(plan)))

        self.Bpad = _round_up(batch_size, 32)
        self._alloc_state()
        self.set_max_grad_norm(max_grad_norm)
        self.clipper = _ClipperShim(self)
        self._seed = 0
        self._philox_offset = 0
        self._set_seed(int(torch.initial_seed() & 0x7FFFFFFFFFFFFFFF))
        self._reset_capture()
        self._accum_bs = 0
        self._clipped: Optional[List[torch.Tensor]] = None
        self._clipped_flat: Optional[torch.Tensor] = None
        self._summed_flat: Optional[torch.Tensor] = None
        self._n_theta = sum(p.numel() for p in self._params)
        self._sm_count = L.device_info()[0]

    @property
    def sample_rate(self) -> float:
        """Sampling rate of the accountant: GLOBAL batch / sample size (the per-rank shard would under-report
        epsilon by about the world size)."""
        return self.global_batch_size / self.sample_size

    # ------------------------------------------------------------------ validation / attach
    def validate(self):
        """Upstream attach() refuses modules it cannot compute per-sample gradients for."""
        for name, m in self.module.named_modules():
            if isinstance(m, _UNSUPPORTED_WITH_PARAMS):
                raise NotImplementedError(f"{name}: BatchNorm is incompatible with per-sample gradients")

    def attach(self, optimizer: torch.optim.Optimizer):
        """Patch optimizer.step / zero_grad like upstream PrivacyEngine.attach (reference train.py:135)."""
        self.optimizer = optimizer
        engine = self

        def dp_step(opt_self, closure=None):
            engine.step()
            return opt_self.original_step(closure)

        def dp_zero_grad(opt_self, *a, **k):
            engine.zero_grad()
            return opt_self.original_zero_grad(*a, **k)

        def virtual_step(opt_self):
            engine.virtual_step()

        optimizer.privacy_engine = self
        optimizer.original_step = optimizer.step
        optimizer.step = types.MethodType(dp_step, optimizer)
        optimizer.original_zero_grad = optimizer.zero_grad
        optimizer.zero_grad = types.MethodType(dp_zero_grad, optimizer)
        optimizer.virtual_step = types.MethodType(virtual_step, optimizer)

    def detach(self):
        opt = self.optimizer
        if opt is not None:
            opt.step = opt.original_step
            opt.zero_grad = opt.original_zero_grad
            del opt.privacy_engine, opt.original_step, opt.original_zero_grad, opt.virtual_step
            self.optimizer = None
        for h in self._handles:
            h.remove()
        self._handles = []

    def _set_seed(self, seed: int):
        """Seed the engine's private Philox stream (upstream: a private torch.Generator on the
        device; reference train.py:136).  State = (seed, offset) exactly like a CUDA generator."""
        self._seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self._philox_offset = 0
        if getattr(self, "_offset_dev", None) is not None:
            self._offset_dev.zero_()

    def enable_graph_safe_rng(self):
        """Keep the Philox offset in device memory so a CUDA-graph-captured step draws fresh noise on every
        replay (the host counter would be frozen into the captured kernel arguments)."""
        if getattr(self, "_offset_dev", None) is None:
            self._offset_dev = torch.tensor([self._philox_offset], dtype=torch.int64, device=self.device)

    @property
    def philox_offset(self) -> int:
        od = getattr(self, "_offset_dev", None)
        return int(od.item()) if od is not None else self._philox_offset

    # checkpointing of the engine state the reference forgets (accountant restarts on resume)
    def state_dict(self) -> Dict:
        return {"steps": self.steps, "seed": self._seed, "philox_offset": self.philox_offset,
                "max_grad_norm": self.max_grad_norm}

    def load_state_dict(self, sd: Dict):
        self.steps = sd["steps"]
        self._seed, self._philox_offset = sd["seed"], sd["philox_offset"]
        if getattr(self, "_offset_dev", None) is not None:
            self._offset_dev.fill_(self._philox_offset)
        self.set_max_grad_norm(sd["max_grad_norm"])

    # ------------------------------------------------------------------ hooks
    def enable_hooks(self, freeze_weights: bool = False):
        """Reference train.py:373.  `freeze_weights=True` additionally takes the module's parameters out of
        autograd while the hooks are on: the per-sample machinery needs every layer's grad_output but never the
        batch-summed weight gradients (the reference computes them in d_loss.backward(), train.py:387, and then
        overwrites p.grad in step()), so the backward pass shrinks to the dgrad chain -- no cuDNN wgrad kernels
        and no gradient clones.  The output of a captured layer whose input carries no gradient (the first
        layer) becomes the leaf the chain starts from.  disable_hooks() restores requires_grad."""
        self.hooks_enabled = True
        if freeze_weights and not self._frozen:
            self._frozen = [(p, p.requires_grad) for p in self.module.parameters()]
            for p, _ in self._frozen:
                p.requires_grad_(False)

    def disable_hooks(self):
        self.hooks_enabled = False
        if self._frozen:
            for p, rg in self._frozen:
                p.requires_grad_(rg)
            self._frozen = []
        for o in self._leaf_outputs:
            o.grad = None
        self._leaf_outputs = []

    def _make_fwd_hook(self, plan: LayerPlan):
        def fwd_hook(layer, inputs, output):
            if not self.hooks_enabled or not torch.is_grad_enabled():
                return
            pass_idx = self._pass_count.get(plan, 0)
            if pass_idx >= self.max_passes:
                raise RuntimeError(
                    f"{plan.name}: more than max_passes={self.max_passes} forward passes captured before clip()")
            act = inputs[0]
            B = act.shape[0]
            if B > self.Bpad:
                self.Bpad = _round_up(B, 32)
                self._alloc_state()
            self._pass_count[plan] = pass_idx + 1
            self._pass_B[pass_idx] = B
            self._cur_B = B
            plan.capture_activation(act, pass_idx, self.Bpad, self.max_passes)
            self._norms_valid = False
            self._factors_valid = False
            if self._frozen and not output.requires_grad:
                output.requires_grad_(True)          # leaf: the dgrad chain of the frozen-weights mode starts here
                self._leaf_outputs.append(output)
            if output.requires_grad:
                scale = float(B) if self.loss_reduction == "mean" else 1.0
                self._outputs.append(output)

                def grad_hook(g, plan=plan, pass_idx=pass_idx, scale=scale):
                    if self.hooks_enabled:
                        plan.capture_backprop(g, pass_idx, scale)
                        self._bp_seen.add((plan, pass_idx))

                output.register_hook(grad_hook)
        return fwd_hook

    def ingest_captures(self, passes: Sequence[Dict[str, Tuple[torch.Tensor, torch.Tensor]]]):
        """Feed already-captured (activation, grad_output) pairs, one dict {layer name: (A, G)} per
        pass in forward order, exactly as the hooks would have staged them.  This is the hot path with
        its inputs resident in HBM (what bench.py times as t_dp) and a handy seam for tests."""
        self._reset_capture()
        by_name = {pl.name: pl for pl in self._plans}
        for ps, layers in enumerate(passes):
            for name, (act, gout) in layers.items():
                plan = by_name[name]
                B = act.shape[0]
                if B > self.Bpad:
                    self.Bpad = _round_up(B, 32)
                    self._alloc_state()
                self._pass_count[plan] = max(self._pass_count.get(plan, 0), ps + 1)
                self._pass_B[ps] = B
                self._cur_B = B
                plan.capture_activation(act, ps, self.Bpad, self.max_passes)
                if gout is not None:
                    plan.capture_backprop(gout, ps, float(B) if self.loss_reduction == "mean" else 1.0)
                    self._bp_seen.add((plan, ps))

    def backward(self, loss: torch.Tensor):
        """Backward pass that stops at the captured layer outputs: the per-sample machinery needs every
        layer's grad_output but never the batch-summed weight gradients autograd would also compute
        (the reference's d_loss.backward(), train.py:387, computes them and then overwrites p.grad in
        step()).  Skipping them removes the cuDNN wgrad kernels from the D step."""
        outs = self._outputs
        if not outs:
            raise RuntimeError("backward(loss) needs a forward pass with hooks enabled")
        if self._frozen:
            # enable_hooks(freeze_weights=True): the graph already holds the dgrad chain only
            loss.backward()
            for o in self._leaf_outputs:
                o.grad = None
            self._leaf_outputs = []
        else:
            # NOTE: autograd clones every captured gradient into o.grad here (measured: 0.54 ms per step on
            # the CelebA critic at B=512); the frozen-weights mode avoids that
            torch.autograd.backward(loss, inputs=outs)
            for o in outs:
                o.grad = None
        self._outputs = []

    def _reset_capture(self):
        self._outputs: List[torch.Tensor] = []
        for o in getattr(self, "_leaf_outputs", []):
            o.grad = None
        self._leaf_outputs: List[torch.Tensor] = []
        self._pass_count: Dict[LayerPlan, int] = {}
        self._pass_B: Dict[int, int] = {}
        self._bp_seen = set()
        self._mat: Optional[List[torch.Tensor]] = None     # materialised grad_sample tensors a caller has written to
        self._cur_B = self.batch_size
        self._norms_valid = False
        self._factors_valid = False

    def _n_passes(self) -> int:
        return max(self._pass_count.values(), default=0)

    def _n_passes_view(self) -> int:
        return 1 if self.accum_passes else max(self._n_passes(), 1)

    def _alloc_state(self):
        S = self.Bpad * self.max_passes
        n = len(self._params)
        dev = self.device
        self._S = S
        self._norm2 = torch.zeros((n, S), device=dev)
        self._norms = torch.zeros((n, S), device=dev)
        self._flat_norms = torch.zeros((1, S), device=dev)
        self._factors = torch.ones((n, S), device=dev)
        for plan in getattr(self, "_plans", []):
            plan.ready = False

    # ------------------------------------------------------------------ thresholds
    @property
    def is_per_layer(self) -> bool:
        return self._per_layer

    def set_max_grad_norm(self, v):
        """float -> flat clipping; list -> per-layer C_k; tensors stay on the device (reference
        train.py:241 passes a list, :243 a 0-dim tensor)."""
        n = len(self._params)
        self._factors_valid = False
        if isinstance(v, torch.Tensor):
            if v.dim() == 0:
                self._per_layer = False
                t = v.detach().to(self.device, torch.float32).reshape(1).clone()
            else:
                self._per_layer = True
                t = v.detach().to(self.device, torch.float32).reshape(-1)
                if t.numel() == 1:
                    t = t.repeat(n)
                if t.numel() != n:
                    raise ValueError(f"expected {n} per-layer thresholds, got {t.numel()}")
                t = t.clone()
            self._thresholds_dev = t
            self.max_grad_norm = v
            self._thresholds_host = None
            self._noise_c_host = None
            if not self._per_layer:
                self._noise_c_dev = t.expand(n).contiguous()
            elif self.per_layer_noise == "l2norm":
                self._noise_c_dev = t.norm(2).reshape(1).expand(n).contiguous()
            else:
                self._noise_c_dev = t
            return
        if isinstance(v, (list, tuple)):
            vals = [float(x) for x in v]
            if len(vals) == 1:
                vals = vals * n
            if len(vals) != n:
                raise ValueError(f"expected {n} per-layer thresholds, got {len(vals)}")
            self._per_layer = True
            self.max_grad_norm = list(v)
            if self.per_layer_noise == "l2norm":
                self._noise_c_host = [math.sqrt(sum(c * c for c in vals))] * n
            else:
                self._noise_c_host = list(vals)
        else:
            vals = [float(v)]
            self._per_layer = False
            self.max_grad_norm = float(v)
            self._noise_c_host = vals * n
        self._thresholds_host = vals
        self._thresholds_dev = torch.tensor(vals, dtype=torch.float32, device=self.device)
        self._noise_c_dev = None

    # ------------------------------------------------------------------ norms / factors
    def _check_captured(self):
        np_ = self._n_passes()
        if np_ == 0:
            raise RuntimeError("no per-sample gradients captured: run forward/backward with hooks enabled first")
        return np_

    def _compute_norms(self):
        if self._norms_valid:
            return
        n_passes = self._check_captured()
        self._norm2.zero_()
        if self._mat is not None:
            # a caller wrote into p.grad_sample (train.py:447): norms are row reductions over the real tensors
            st = L.stream_ptr(self.device)
            for k, t in enumerate(self._mat):
                R = self._params[k].numel()
                for ps in range(t.shape[0]):
                    B = t.shape[1]
                    L.call("cg_row_sumsq", L.ptr(t[ps]), B, R, R, L.ptr(self._norm2[k][ps * self.Bpad:]), 0, st)
            self._norms_valid = True
            self._factors_valid = False
            return
        joint = n_passes if (self.accum_passes and n_passes > 1) else 1
        if joint > 1 and len({self._pass_B[ps] for ps in range(n_passes)}) != 1:
            raise RuntimeError("accum_passes=True needs the same batch size in every pass")
        small = [] if self.batch_small_ops else None         # per-layer scalar work batched into one launch
        for plan in self._plans:
            live = [ps for ps in range(self._pass_count.get(plan, 0)) if (plan, ps) in self._bp_seen]
            if not live:
                continue                           # layer got no backprop at all -> zero gradient
            if joint > 1:
                # U2: the per-sample gradients of all passes are summed first; the joint norms live in
                # pass 0's slots.  A layer missing a backprop in some pass is not supported here.
                if len(live) != n_passes:
                    raise NotImplementedError(f"{plan.name}: accum_passes=True with a layer unused in one pass")
                plan.weight_norm2(self._norm2[plan.w_idx], 0, self._pass_B[0], joint)
                if plan.b_idx is not None:
                    plan.bias_norm2(self._norm2[plan.b_idx], 0, self._pass_B[0], joint)
                continue
            # when every pass fills its slots exactly, the passes are one contiguous slot range: one launch
            if len(live) == n_passes and all(self._pass_B[ps] == self.Bpad for ps in live):
                spans = [(0, n_passes * self.Bpad)]
            else:
                spans = [(ps, self._pass_B[ps]) for ps in live]
            for ps, nb in spans:
                plan.weight_norm2(self._norm2[plan.w_idx], ps, nb, 1, small)
                if plan.b_idx is not None:
                    plan.bias_norm2(self._norm2[plan.b_idx], ps, nb, 1, small)
        if small:
            L.small_ops(small, L.stream_ptr(self.device))     # bias norms, Linear closed forms: ONE launch
        self._norms_valid = True
        self._factors_valid = False

    def _compute_factors(self):
        self._compute_norms()
        if self._factors_valid:
            return
        n = len(self._params)
        S = self._S
        clip_lo, clip_hi = 0, S
        if (not self.accum_passes) and (not self.split_clip_fake) and self.num_private_passes is not None:
            clip_lo = (self._n_passes() - self.num_private_passes) * self.Bpad
        st = L.stream_ptr(self.device)
        c_scale = 1.0 - self.clip_margin
        if self._per_layer:
            L.call("cg_clip_factors", L.ptr(self._norm2), n, S, 1, L.ptr(self._thresholds_dev), c_scale, clip_lo, clip_hi,
                   L.ptr(self._factors), L.ptr(self._norms), st)
        else:
            L.call("cg_clip_factors", L.ptr(self._norm2), n, S, 0, L.ptr(self._thresholds_dev), c_scale, clip_lo, clip_hi,
                   L.ptr(self._factors), L.ptr(self._flat_norms), st)
        # slots past the live batch of a pass (batch smaller than the padded slot count) may hold the staged
        # rows of an earlier, larger step; clip() contracts whole slot ranges, so their factor must be 0
        fv = self._factors.view(self._factors.shape[0], self.max_passes, self.Bpad)
        live_b = [self._pass_B[ps] for ps in range(self._n_passes())]
        if len(set(live_b)) == 1:
            if live_b[0] < self.Bpad:
                fv[:, :len(live_b), live_b[0]:].zero_()
        else:
            for ps, b in enumerate(live_b):
                if b < self.Bpad:
                    fv[:, ps, b:].zero_()
        self._factors_valid = True

    def _slot_view(self, t: torch.Tensor) -> torch.Tensor:
        """[rows, S] -> [rows, n_passes, B] view over the live slots."""
        np_ = self._n_passes_view()
        return t.view(t.shape[0], self.max_passes, self.Bpad)[:, :np_, :self._cur_B]

    def per_sample_norms(self) -> torch.Tensor:
        """[n_params, n_passes, B] per-parameter per-sample L2 norms (device tensor, no sync)."""
        self._compute_norms()
        return self._slot_view(self._norm2).sqrt()

    def flat_sample_norms(self) -> torch.Tensor:
        """[n_passes, B] norms across all parameters."""
        self._compute_norms()
        return self._slot_view(self._norm2).sum(dim=0).sqrt()

    def clipping_factors(self) -> torch.Tensor:
        """[n_params or 1, n_passes, B]."""
        self._compute_factors()
        rows = self._factors if self._per_layer else self._factors[:1]
        return self._slot_view(rows)

    def adaptive_thresholds(self, stat: str = "mean", scalar: float = 1.5, pass_idx: int = 0) -> torch.Tensor:
        """Per-parameter mean/max of the per-sample norms of one pass times `scalar`, as a DEVICE
        tensor (reference train.py:230-243 does this with one .cpu().item() sync per parameter)."""
        self._compute_norms()
        n = len(self._params)
        norms = self._norm2.sqrt()
        out = torch.empty(n, device=self.device)
        lo = pass_idx * self.Bpad
        L.call("cg_row_stat", L.ptr(norms), n, self._S, lo, lo + self._pass_B.get(pass_idx, self._cur_B),
               1 if stat == "max" else 0, float(scalar), L.ptr(out), L.stream_ptr(self.device))
        return out

    def _materialized_for_write(self) -> List[torch.Tensor]:
        """Materialise every parameter's [n_passes, B, *shape] per-sample gradients (once per capture) so a caller
        can modify them in place; norms, factors and the clipped sum are then recomputed from these tensors."""
        if self._mat is None:
            mats = [self.materialize_grad_sample(k).contiguous() for k in range(len(self._params))]
            self._mat = mats
        self._norms_valid = False
        self._factors_valid = False
        return self._mat

    def _clip_from_materialized(self, outs: List[torch.Tensor]):
        st = L.stream_ptr(self.device)
        for k, t in enumerate(self._mat):
            R = self._params[k].numel()
            tmp = torch.empty(R, device=self.device)
            frow = self._factors[k if self._per_layer else 0]
            for ps in range(t.shape[0]):
                L.call("cg_weighted_colsum", L.ptr(t[ps]), L.ptr(frow[ps * self.Bpad:]), 0, t.shape[1], R, L.ptr(tmp),
                       1 if ps > 0 else 0, st)
            outs[k].copy_(tmp.view(self._params[k].shape))

    def materialize_grad_sample(self, p_idx: int) -> torch.Tensor:
        """[n_passes, B, *p.shape] (rare path; reference train.py:447 and tests)."""
        n_passes = self._check_captured()
        p = self._params[p_idx]
        outs = []
        for plan in self._plans:
            if plan.w_idx == p_idx:
                for ps in range(n_passes):
                    B = self._pass_B[ps]
                    if (plan, ps) in self._bp_seen:
                        outs.append(plan.materialize(ps, B))
                    else:
                        outs.append(torch.zeros((B,) + tuple(p.shape), device=self.device))
            elif plan.b_idx == p_idx:
                for ps in range(n_passes):
                    B = self._pass_B[ps]
                    lo = ps * self.Bpad
                    outs.append(plan.live_bias_rows()[lo:lo + B].clone() if (plan, ps) in self._bp_seen
                                else torch.zeros((B,) + tuple(p.shape), device=self.device))
        g = torch.stack(outs, dim=0)
        if self.accum_passes:
            g = g.sum(dim=0, keepdim=True)
        return g

    # ------------------------------------------------------------------ clip / accumulate / step
    def clip(self):
        """norms -> factors -> clipped weighted sum (reference train.py:399).  The per-pass sums of
        split mode are produced already added together; accum_grads_across_passes() is then a no-op."""
        n_passes = self._check_captured()
        self._compute_factors()
        # ONE flat buffer [|theta| + 1]: the parameters' clipped sums are views into it (same strides as the
        # parameter, so channels_last weights keep their layout) and the last element carries the live batch
        # count -> the data-parallel exchange is one allreduce of this buffer, no concatenation
        flat = self._new_flat()
        outs, off = [], 0
        for p in self._params:
            dense = p.is_contiguous() or (p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last))
            outs.append(flat[off:off + p.numel()].as_strided(p.shape, p.stride()) if dense else torch.empty_like(p))
            off += p.numel()
        self._clipped_flat = flat
        if self._mat is not None:
            self._clip_from_materialized(outs)
            self._clipped = outs
            return outs
        slot_hi = (n_passes - 1) * self.Bpad + self._pass_B[n_passes - 1]
        plans, buckets = self._plans, None
        if self.data_parallel and self.overlap_allreduce:
            buckets = self._reduce_buckets()
            if buckets is not None:
                plans = list(reversed(self._plans))          # last layers first: their (large) sums reduce under the rest
                flat[self._n_theta:].fill_(float(self._cur_B))
                self._reduced_early = True
        sms = self._sm_count
        st = L.stream_ptr(self.device)
        joint = self.accum_passes and n_passes > 1
        # the whole flat buffer is cleared ONCE (instead of one fill per layer), and the small per-layer launches that
        # depend on the factors only -- clip multipliers, bias sums, thin-layer sums -- go into ONE cg_small_ops table
        # ahead of the GEMMs.  Joint clipping and the overlapped allreduce keep the layer-by-layer order.
        batched = self.batch_small_ops and not joint and buckets is None
        if batched:
            flat.zero_()
            for o_, p_ in zip(outs, self._params):
                if o_.data_ptr() < flat.data_ptr() or o_.data_ptr() >= flat.data_ptr() + flat.numel() * 4:
                    o_.zero_()
        todo = []
        for plan in plans:
            live = [ps for ps in range(self._pass_count.get(plan, 0)) if (plan, ps) in self._bp_seen]
            if not live:
                outs[plan.w_idx].zero_()
                if plan.b_idx is not None:
                    outs[plan.b_idx].zero_()
                continue
            if joint:
                # every pass is scaled with pass 0's (joint) factors
                ranges = [(ps * self.Bpad, ps * self.Bpad + self._pass_B[ps], ps * self.Bpad) for ps in live]
            elif len(live) == n_passes:
                ranges = [(0, slot_hi, 0)]
            else:
                ranges = [(ps * self.Bpad, ps * self.Bpad + self._pass_B[ps], 0) for ps in live]
            todo.append((plan, ranges))
        small = [] if batched else None
        ready = set()
        if batched:
            for plan, ranges in todo:
                frow_w = self._factors[plan.w_idx if self._per_layer else 0]
                if len(ranges) == 1:
                    lo, hi, _ = ranges[0]
                    op = plan.clip_mult_op(frow_w, lo, _round_up(hi, 32))
                    if op is not None:
                        small.append(op)
                        ready.add(plan)
                if plan.b_idx is not None:
                    frow_b = self._factors[plan.b_idx if self._per_layer else 0]
                    for lo, hi, shift in ranges:
                        plan.bias_weighted_sum(outs[plan.b_idx], frow_b, lo, hi, accumulate=True, factor_shift=shift, ops=small)
                if getattr(plan.impl, "thin", False) and len(ranges) == 1:
                    plan.weighted_sum(outs[plan.w_idx], ranges[0][0], ranges[0][1], sms, accumulate=False,
                                      factor_row=frow_w, prezeroed=True, ops=small)
                    ready.add(("thin", plan))
            L.small_ops(small, st)
            # ... and the factor-scaled operands of all those layers in ONE launch
            segs, scaled = [], set()
            for plan, ranges in todo:
                if plan in ready and len(ranges) == 1:
                    sg = plan.scale_seg(ranges[0][0], _round_up(ranges[0][1], 32))
                    if sg is not None:
                        segs.append(sg)
                        scaled.add(plan)
            if segs:
                L.scale_slots_multi(segs, st)
        for plan, ranges in todo:
            if ("thin", plan) in ready:
                continue                                     # its clipped sum was one of the batched operations
            frow_w = self._factors[plan.w_idx if self._per_layer else 0]
            # scale up to the next 32-slot boundary (Bpad is a multiple of 32, dead slots have factor 0):
            # the contraction reads whole 32-row K blocks, which may reach past `hi` when Q < 32
            if joint or len(ranges) == 1:
                for lo, hi, shift in ranges:
                    if not (batched and plan in scaled):
                        plan.scale_backprops(frow_w, lo, _round_up(hi, 32), shift, mult_ready=plan in ready)
                # the scaled operand now covers every live slot: ONE GEMM over the whole range
                plan.weighted_sum(outs[plan.w_idx], ranges[0][0], ranges[-1][1], sms, accumulate=False,
                                  factor_row=frow_w, prezeroed=batched)
            else:
                # range by range: with FP16 operands every scale_backprops() call has its own common scale
                for i, (lo, hi, shift) in enumerate(ranges):
                    plan.scale_backprops(frow_w, lo, _round_up(hi, 32), shift)
                    plan.weighted_sum(outs[plan.w_idx], lo, hi, sms, accumulate=i > 0, factor_row=frow_w,
                                      prezeroed=batched and i == 0)
            if plan.b_idx is not None and not batched:
                frow_b = self._factors[plan.b_idx if self._per_layer else 0]
                for i, (lo, hi, shift) in enumerate(ranges):
                    plan.bias_weighted_sum(outs[plan.b_idx], frow_b, lo, hi, accumulate=i > 0, factor_shift=shift)
            if buckets is not None and plan in buckets:
                import torch.distributed as dist
                a, b = buckets[plan]
                self._reduce_works.append(dist.all_reduce(flat[a:b], op=dist.ReduceOp.SUM, group=self.process_group,
                                                          async_op=True))
                sms = max(2, self._sm_count - self.nccl_reserve_sms)    # leave NCCL its SMs from here on
        self._clipped = outs
        return outs

    def _reduce_buckets(self, min_bytes: int = 1 << 21):
        """{plan: (lo, hi)}: after that plan's sums are written (walking the layers last to first) the slice
        flat[lo:hi] -- a contiguous run of whole layers, at least `min_bytes` unless it is the last one -- is complete
        and can be all-reduced.  The live-sample count (last element) rides with the first bucket.  None when a
        parameter is not a view of the flat buffer or the layers are not laid out in parameter order."""
        offs, off = [], 0
        for p in self._params:
            if not (p.is_contiguous() or (p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last))):
                return None
            offs.append(off)
            off += p.numel()
        spans = []
        for plan in self._plans:
            idx = [plan.w_idx] + ([plan.b_idx] if plan.b_idx is not None else [])
            lo = min(offs[k] for k in idx)
            hi = max(offs[k] + self._params[k].numel() for k in idx)
            if hi - lo != sum(self._params[k].numel() for k in idx):
                return None
            spans.append((lo, hi))
        if any(spans[i][1] != spans[i + 1][0] for i in range(len(spans) - 1)) or spans[0][0] != 0 or spans[-1][1] != off:
            return None
        out, hi = {}, off + 1                                  # + the count element
        for i in range(len(self._plans) - 1, -1, -1):
            lo = spans[i][0]
            if (hi - lo) * 4 >= min_bytes or i == 0:
                out[self._plans[i]] = (lo, hi)
                hi = lo
        return out

    def _new_flat(self) -> torch.Tensor:
        """The flat clipped-sum buffer of this clip() call.  Data parallel over NVSwitch: one of two buffers in
        symmetric (multicast-mapped) memory, so step() can sum, noise and broadcast in one kernel; else plain memory."""
        if self.data_parallel and self.fused_allreduce and self._symm is None and not self._symm_failed:
            from .dist import SymmetricFlat
            try:
                if SymmetricFlat.available(self.process_group):
                    self._symm = SymmetricFlat(self._n_theta + 1, self.device, self.process_group)
                else:
                    self._symm_failed = True
            except Exception as e:                      # no symmetric allocator / no multicast: keep NCCL
                self._symm_failed = True
                self._symm_error = repr(e)
        if self._symm is not None:
            busy = self._symm.index_of(self._summed_flat)
            i = 1 - busy if busy >= 0 else 1 - self._symm_last
            self._symm_last = i
            return self._symm.bufs[i]
        return torch.empty(self._n_theta + 1, device=self.device)

    def accum_grads_across_passes(self):
        """Sum the per-pass clipped sums (reference train.py:402).  Already fused into clip()."""
        if self._clipped is None:
            raise RuntimeError("accum_grads_across_passes() called before clip()")

    def accumulate_batch(self):
        """p.summed_grad (+)= clipped sum; a SUM, not a mean (reference train.py:417, 431)."""
        if self._clipped is None:
            raise RuntimeError("accumulate_batch() called before clip()")
        fresh = all(getattr(p, "summed_grad", None) is None for p in self._params)
        if fresh:
            for p, c in zip(self._params, self._clipped):
                p.summed_grad = c
            self._summed_flat = self._clipped_flat
        elif self._summed_views_intact() and self._clipped_flat is not None:
            self._wait_reduces()
            self._summed_flat.add_(self._clipped_flat)          # gradient accumulation: one kernel, count included
        else:
            self._wait_reduces()
            self._summed_flat = None
            for p, c in zip(self._params, self._clipped):
                if getattr(p, "summed_grad", None) is None:
                    p.summed_grad = c
                else:
                    p.summed_grad.add_(c)
        # only the private pass counts towards the batch size (both passes hold the same B)
        self._accum_bs += self._cur_B
        self._clipped = None
        self._clipped_flat = None
        self._reset_capture()
        self._drop_grad_sample_attrs()

    def _wait_reduces(self):
        for w in self._reduce_works:
            w.wait()
        had = bool(self._reduce_works)
        self._reduce_works = []
        return had

    def _summed_views_intact(self) -> bool:
        """Are the p.summed_grad tensors still the views into `_summed_flat` this engine handed out?  (The caller
        may add to them in place, reference train.py:431; a caller that REPLACES them gets the slow path.)"""
        flat = self._summed_flat
        if flat is None:
            return False
        off = flat.data_ptr()
        for p in self._params:
            sg = getattr(p, "summed_grad", None)
            if sg is None or sg.data_ptr() != off or sg.stride() != p.stride():
                return False
            off += 4 * p.numel()
        return True

    def _drop_grad_sample_attrs(self):
        for p in self._params:
            if hasattr(p, "grad_sample"):
                del p.grad_sample

    def expose_grad_sample_attrs(self):
        """Attach the lazy `p.grad_sample` views (reference train.py:233, 388 read them)."""
        for k, p in enumerate(self._params):
            p.grad_sample = GradSampleView(self, k)

    def noise_stds(self) -> List[float]:
        """Per-parameter noise standard deviation sigma * c_k (host thresholds only): c_k = C (flat), ||C||_2
        (per-layer, default) or C_k (per_layer_noise="own")."""
        if self._noise_c_host is None:
            raise RuntimeError("thresholds live on the device")
        return [self.noise_multiplier * c for c in self._noise_c_host]

    def virtual_step(self):
        self.clip()
        self.accumulate_batch()

    def step(self):
        """Engine half of the patched optimizer.step() (reference train.py:484): p.grad = summed/B,
        noise = N(0, (sigma*c_k)^2) drawn per parameter tensor from the Philox stream, noise /= B
        (mean reduction), p.grad += noise.  ONE kernel launch for all parameter tensors."""
        if self.auto_clip_and_accum_on_step and self._n_passes() > 0:
            self.clip()
            self.accumulate_batch()
        if self._accum_bs == 0:
            raise ValueError("No accumulated gradients: call clip()/accumulate_batch() before step()")
        self.steps += 1
        bs = float(self._accum_bs)
        div_dev = None
        fused_i = -1
        if (self.data_parallel and self._symm is not None and not self._reduced_early and self._summed_views_intact()):
            fused_i = self._symm.index_of(self._summed_flat)
        if fused_i >= 0:
            pass                                             # exchange + noise in one kernel below
        elif self.data_parallel:
            if self._reduced_early:
                # clip() already reduced every bucket (overlap_allreduce): only wait; the count rode along
                self._wait_reduces()
                div_dev = self._summed_flat[self._n_theta:]
            else:
                div_dev = self._allreduce_summed(bs)        # device scalar: the GLOBAL number of samples
        mean = self.loss_reduction == "mean"
        st = L.stream_ptr(self.device)
        od = getattr(self, "_offset_dev", None)
        segs = []
        for k, p in enumerate(self._params):
            s = p.summed_grad
            if self._noise_c_host is not None:
                segs.append((s, s, self.noise_multiplier * self._noise_c_host[k], None))
            else:
                segs.append((s, s, self.noise_multiplier, self._noise_c_dev[k:k + 1]))
        div = bs if mean else 0.0
        if fused_i >= 0:
            # cross-rank sum, division by the global sample count, noise and broadcast in ONE kernel over the NVSwitch
            # multicast mapping of the flat buffer; every rank draws only its 1/world share of the normals
            sy = self._symm
            flat = self._summed_flat
            flat[self._n_theta:].fill_(bs)
            sy.barrier(fused_i)                              # every rank's sums (and count) are written
            inc = L.noise_multi_allreduce(segs, mean, self._seed, 0 if od is not None else self._philox_offset, od,
                                          flat, sy.mc_ptrs[fused_i] if sy.use_multicast else 0,
                                          sy.peer_ptrs[fused_i] if (sy.mix or not sy.use_multicast) else None,
                                          self._n_theta, sy.rank, sy.world, st)
            sy.barrier(fused_i)                              # every rank's share of the result has landed everywhere
        else:
            inc = L.noise_multi(segs, div, div_dev if mean else None, div, div_dev if mean else None, self._seed,
                                0 if od is not None else self._philox_offset, od, st)
        for p in self._params:
            p.grad = p.summed_grad                           # noised in place
            p.summed_grad = None
        self._summed_flat = None
        self._reduced_early = False
        if od is not None:
            if inc:
                L.call("cg_philox_advance", L.ptr(od), inc, st)
        else:
            self._philox_offset += inc
        self._accum_bs = 0

    def _allreduce_summed(self, bs: float) -> torch.Tensor:
        """Data parallel: one NCCL allreduce(SUM) of the flat clipped-sum buffer, whose last element is this
        rank's live sample count, so the divisor is the true global batch even with unequal shards; noise is then
        drawn identically on every rank from the shared Philox (seed, offset), so all replicas apply the same
        update.  Returns the device scalar holding the global count."""
        from .dist import allreduce_sum_and_count
        if self._summed_views_intact():
            flat = self._summed_flat
        else:
            flat = torch.cat([p.summed_grad.reshape(-1) for p in self._params]
                             + [torch.zeros(1, device=self.device)])
            off = 0
            for p in self._params:
                p.summed_grad = flat[off:off + p.numel()].view(p.shape)
                off += p.numel()
        flat[self._n_theta:].fill_(bs)
        allreduce_sum_and_count(flat, group=self.process_group)
        return flat[self._n_theta:]

    def zero_grad(self):
        """Patched optimizer.zero_grad (reference train.py:245): also drops captured state."""
        self._reset_capture()
        self._clipped = None
        self._clipped_flat = None
        self._drop_grad_sample_attrs()

    # ------------------------------------------------------------------ accountant
    def get_renyi_divergence(self):
        return compute_rdp(self.sample_rate, self.noise_multiplier, 1, self.alphas)

    def get_privacy_spent(self, target_delta: Optional[float] = None) -> Tuple[float, float]:
        """(epsilon, best alpha) after `self.steps` steps (reference train.py:295, 588;
        budget_analysis.py:79-80 also assigns `.steps` directly)."""
        if target_delta is None:
            target_delta = self.misc_settings.get("target_delta", 1e-6)
        rdp = compute_rdp(self.sample_rate, self.noise_multiplier, self.steps, self.alphas)
        return _rdp_to_eps(self.alphas, rdp, target_delta)
